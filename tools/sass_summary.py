#!/usr/bin/env python3
"""Per-kernel counts of the SASS opcodes that show which hardware path a kernel uses (run here, no GPU needed):
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = TMEM load / store, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR =
tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync, LDSM = ldmatrix, UCGABAR = cluster barrier, MUFU = special-function unit."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "openai-whisper-coreml_b200/libwhisper_b200.so"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "LDSM", "UCGABAR", "MUFU", "FFMA2", "LDG", "STG", "LDS", "STS", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("wb::", "")
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        base = op.split(".")[0]
        for k in KEYS:
            if base == k or base.startswith(k + "_"):
                counts[name][k] += 1
        counts[name]["total"] += 1
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
for n, c in sorted(counts.items()):
    print(f"{n[:58]:58s} {c['total']:6d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
