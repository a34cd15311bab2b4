cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py base.en 32 1 > gpurun_out/profile_step.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 6 -f -o gpurun_out/prof_gemm_tc python tools/profile_step.py base.en 32 1 > gpurun_out/prof2.log 2>&1; tail -1 gpurun_out/prof2.log
