#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) as a markdown table: one row per captured launch.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "warp-inst"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__bytes.sum.per_second", "dram B/s"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st.long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st.short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st.wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st.barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st.math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st.mio"),
]


def summarise(rows):
    """rows = parsed `ncu --page raw --csv` output -> markdown table (one row per captured launch)."""
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = ["| kernel | " + " | ".join(c[1] for c in COLS if c[0] in ix) + " |",
           "|---|" + "---|" * len([c for c in COLS if c[0] in ix])]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        cells = []
        for k, _ in COLS:
            if k not in ix:
                continue
            v, u = r[ix[k]], units[ix[k]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3g" % f) if abs(f) < 1e6 else ("%.3e" % f)
            except ValueError:
                pass
            cells.append("%s %s" % (v, u) if u and u not in ("%", "") else v)
        out.append("| %s | %s |" % (name, " | ".join(cells)))
    return "\n".join(out) + "\n"


def main():
    src = sys.argv[1]
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    print(summarise(list(csv.reader(io.StringIO(raw)))), end="")


if __name__ == "__main__":
    main()
