#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for the Blackwell-native claim (no GPU needed):
    python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.md
Counts, per kernel of libsvnet_b200.so, the tensor-core / tensor-memory / bulk-copy / popcount mnemonics
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "svnet_b200", "libsvnet_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "F2FP", "POPC", "HMMA", "IMMA", "LDGSTS", "MUFU", "FFMA", "FFMA2", "LDG", "STG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0].replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        hist[kern]["_total"] += 1
print("# SASS opcode counts per kernel of libsvnet_b200.so (cuobjdump -sass, sm_100a)\n")
print("| kernel | instructions | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for k, h in hist.items():
    if any(h[w] for w in WATCH[:7]) or "edge" in k or "knn" in k:
        print("| `%s` | %d | " % (k[:110], h["_total"]) + " | ".join(str(h[w]) if h[w] else "" for w in WATCH) + " |")
