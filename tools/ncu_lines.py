"""Developer tool: aggregate ncu warp-stall samples per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, agg, tot = None, {}, 0
stall_cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        si, ie = r.index("# Samples"), r.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if r[0] != "":
        try: s, n = int(r[si]), int(r[ie])
        except Exception: continue
        key = (cur, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += s; a[1] += n; tot += s
        for i, c in stall_cols:
            try: v = int(r[i])
            except Exception: v = 0
            if v: a[2][c] = a[2].get(c, 0) + v
print("total samples", tot)
for (f, l, src), (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ",".join("%s:%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%6d %5.1f%% inst=%-8d %s:%d  %s   [%s]" % (s, 100 * s / tot, n, f, l, src, tops))
