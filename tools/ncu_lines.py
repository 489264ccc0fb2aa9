"""Developer tool: aggregate ncu warp-stall samples per CUDA source line.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [N] [kernel-substring]"""
import csv, sys
csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else None
cur, fn, agg, tot = None, None, {}, 0
si = ie = None
stall_cols = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No":
        si, ie = r.index("# Samples"), r.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if want and (fn is None or want not in fn): continue
    if not (cur or "").endswith((".cu", ".cuh")): continue
    try: line, s, n = int(r[0]), int(r[si]), int(r[ie])
    except Exception: continue
    key = (cur, line, r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += s; a[1] += n; tot += s
    for i, c in stall_cols:
        try: v = int(r[i])
        except Exception: v = 0
        if v: a[2][c] = a[2].get(c, 0) + v
print("total samples", tot, "instructions", sum(a[1] for a in agg.values()))
for (f, l, src), (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ",".join("%s:%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%6d %5.1f%% inst=%-9d %s:%d  %s   [%s]" % (s, 100 * s / max(tot, 1), n, f, l, src, tops))
