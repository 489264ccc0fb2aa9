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
    # A kNN call = pack + tcgen05 + finish kernels (or the CUDA-core knn_kernel); it belongs to the layer
    # of the next edge kernel in launch order.  Edge layers are told apart by their template shapes, so a
    # capture that starts mid-forward or covers more than one forward still maps correctly.
    body = rows[2:]
    d = {}
    pending = None
    edge_seen = []
    for r in body:
        t = nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
        name = r[ix["Kernel Name"]]
        if "knn_pack_kernel" in name or ("knn_kernel" in name and "tc" not in name):
            pending = t
        elif "knn_tc_kernel" in name or "knn_finish_kernel" in name:
            pending = t if pending is None else pending + t      # (a capture may start after the pack kernel)
        elif "edge_xyz" in name:
            if pending is not None:
                d["svnet_knn[layer1]"] = pending
            pending = None
        elif "edge_bin_fast" in name or "svblock_edge_kernel" in name or "edge_fp_fast" in name:
            shape = name.split("kernel<")[1].split(">")[0] if "kernel<" in name else "0, 0, 0, 0"
            d["_edge_" + shape] = (t, pending)
            pending = None
    shapes = sorted([k[6:] for k in d if k.startswith("_edge_")], key=lambda sh: [int(v) for v in sh.split(",")[:4]])
    for i, sh in enumerate(shapes):
        t, kn = d.pop("_edge_" + sh)
        d["svnet_svblock_edge_fwd[layer%d]" % (i + 2)] = t
        if kn is not None:
            d["svnet_knn[layer%d]" % (i + 2)] = kn
    d = dict(sorted(d.items()))
    json.dump(d, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    # what limits the edge kernels (bench.py quotes it next to the HBM fraction): issue-slot and DRAM utilisation
    lim = {}
    edge_rows = {}
    for r in body:
        name = r[ix["Kernel Name"]]
        if "edge_bin_fast" in name or "edge_fp_fast" in name:
            edge_rows[name.split("kernel<")[1].split(">")[0]] = r
    for i, sh in enumerate(sorted(edge_rows, key=lambda sh: [int(v) for v in sh.split(",")[:4]])):
        r = edge_rows[sh]
        lim["svnet_svblock_edge_fwd[layer%d]" % (i + 2)] = {
            "issue_slot_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
            "dram_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
            "warp_instructions": float(r[ix["smsp__inst_executed.sum"]].replace(",", "")),
            "source": "profiles/%s_ncu_full_all_kernels.md" % tag}
    json.dump(lim, open(os.path.join(ROOT, "profiles", "limiters.json"), "w"), indent=1)
    for src, dst in (("launches_r1.csv", tag + "_launches.csv"), ("bench_r1.json", tag + "_bench.json")):
        p = os.path.join(ROOT, "gpurun_out", src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(ROOT, "profiles", dst))
    print(md)


if __name__ == "__main__":
    main()
