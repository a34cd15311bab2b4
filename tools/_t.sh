cd /root/repo
for pf in 0 2 4 6 9; do echo "== L2PF=$pf"; WB_HA_L2PF=$pf timeout 300 python tools/gpu_bringup.py bench 2>&1 | tail -1 | grep -o "| [0-9.]*ms timings.*" | tail -c 100; done
