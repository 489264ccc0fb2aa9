#!/usr/bin/env python
"""Turn the raw ncu CSV of one forward (+ launch list / bench json) into the tracked summaries under
profiles/.  On the GPU box:
    ncu --set full --clock-control none -o /tmp/prof ... python tools/time_forward.py --iters 1
    ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/<tag>_raw.csv
Here:  python tools/refresh_profiles.py gpurun_out/<tag>_raw.csv <tag>"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import summarise  # noqa: E402


def main():
    raw_csv, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw_csv)))
    md = summarise(rows)
    open(os.path.join(ROOT, "profiles", tag + "_ncu_full_all_kernels.md"), "w").write(md)
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def nbytes(r, key):
        return float(r[ix[key]].replace(",", "")) * mult[units[ix[key]]]
    knn, edge = [], []
    body = rows[2:]
    # the capture may start mid-forward: rotate so that it begins with layer 1's kNN (the pack kernel that
    # precedes the first-layer gate / edge kernels)
    names = [r[ix["Kernel Name"]] for r in body]
    first_xyz = next((i for i, n in enumerate(names) if "gate_xyz" in n or "edge_xyz" in n), None)
    if first_xyz is not None:
        start = max(i for i in range(first_xyz) if "knn_pack_kernel" in names[i] or "knn_kernel" in names[i]) \
            if any("knn_pack_kernel" in n or "knn_kernel" in n for n in names[:first_xyz]) else 0
        body = body[start:] + body[:start]
    for r in body:
        t = nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
        name = r[ix["Kernel Name"]]
        if "knn_pack_kernel" in name:
            knn.append(t)                      # a kNN call = pack + tcgen05 + finish kernels
        elif "knn_tc_kernel" in name or "knn_finish_kernel" in name:
            knn[-1] += t
        elif "knn_kernel" in name:
            knn.append(t)
        if "edge_bin_fast" in name or "svblock_edge_kernel" in name:
            edge.append(t)
    d = {"svnet_knn[layer%d]" % (i + 1): t for i, t in enumerate(knn[:4])}
    d.update({"svnet_svblock_edge_fwd[layer%d]" % (i + 2): t for i, t in enumerate(edge[:3])})
    json.dump(d, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    for src, dst in (("launches_r1.csv", tag + "_launches.csv"), ("bench_r1.json", tag + "_bench.json")):
        p = os.path.join(ROOT, "gpurun_out", src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(ROOT, "profiles", dst))
    print(md)


if __name__ == "__main__":
    main()
