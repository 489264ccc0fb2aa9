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

    rate = {"byte/s": 1, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}
    tunit = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}

    def nbytes(r, key):
        if key in ix:
            return float(r[ix[key]].replace(",", "")) * mult[units[ix[key]]]
        if key == "dram__bytes_read.sum":      # section-limited capture: total DRAM bytes = rate x duration (read + write)
            bi, ti = ix["dram__bytes.sum.per_second"], ix["gpu__time_duration.sum"]
            return float(r[bi].replace(",", "")) * rate[units[bi]] * float(r[ti].replace(",", "")) * tunit[units[ti]]
        return 0.0
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
        elif "edge_bin_fast" in name or "svblock_edge_kernel" in name or "edge_fp_fast" in name or "edge_bin_tc" in name:
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
    # what limits the hot kernels (bench.py quotes it next to the HBM fraction): the compute-side roofline --
    # issue-slot, tensor-pipe, L1 / L2 and DRAM utilisation as ncu measured them
    def col(r, key, default=None):
        try:
            return float(r[ix[key]].replace(",", ""))
        except (KeyError, ValueError):
            return default
    lim = {}
    edge_rows = {}
    knn_rows = []          # (tc kernel row, finish kernel row) per layer, in launch order
    cur = {}
    for r in body:
        name = r[ix["Kernel Name"]]
        if "edge_bin_fast" in name or "edge_fp_fast" in name or "edge_bin_tc" in name:
            edge_rows[name.split("kernel<")[1].split(">")[0]] = r
        if "knn_tc_kernel" in name:
            cur = {"tc": r}
        elif "knn_finish" in name and "tc" in cur:
            cur["fin"] = r
            knn_rows.append(cur)
            cur = {}

    def entry(r):
        return {"kernel": r[ix["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""),
                "time_us": col(r, "gpu__time_duration.sum"),
                "issue_slot_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "tensor_pipe_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "l1tex_pct": col(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                "l2_pct": col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                "dram_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "warp_instructions": col(r, "smsp__inst_executed.sum"),
                "source": "profiles/%s_ncu_full_all_kernels.md" % tag}
    for i, sh in enumerate(sorted(edge_rows, key=lambda sh: [int(v) for v in sh.split(",")[:4]])):
        lim["svnet_svblock_edge_fwd[layer%d]" % (i + 2)] = entry(edge_rows[sh])
    for i, kr in enumerate(knn_rows[:4]):
        lim["svnet_knn[layer%d]" % (i + 1)] = {"score_kernel": entry(kr["tc"]), "finish_kernel": entry(kr["fin"])}
    json.dump(lim, open(os.path.join(ROOT, "profiles", "limiters.json"), "w"), indent=1)
    for src, dst in ((tag + "_launches.csv", tag + "_launches.csv"), (tag + "_bench.json", tag + "_bench.json")):
        p = os.path.join(ROOT, "gpurun_out", src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(ROOT, "profiles", dst))
    print(md)


if __name__ == "__main__":
    main()
