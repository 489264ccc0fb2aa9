#!/usr/bin/env python
"""Turn gpurun_out/<rep>.ncu-rep (+ launch list / bench json) into the tracked summaries under profiles/.
Usage: python tools/refresh_profiles.py gpurun_out/prof_r1d.ncu-rep r1_v4"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    md = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", tag + "_ncu_full_all_kernels.md"), "w").write(md)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def nbytes(r, key):
        return float(r[ix[key]].replace(",", "")) * mult[units[ix[key]]]
    knn, edge = [], []
    for r in rows[2:]:
        t = nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
        if "knn_kernel" in r[ix["Kernel Name"]]:
            knn.append(t)
        if "edge_bin_fast" in r[ix["Kernel Name"]] or "svblock_edge_kernel" in r[ix["Kernel Name"]]:
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
