"""Developer tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
h = rows[0]; ki = h.index("Kernel Name"); gi = h.index("Grid Size"); vi = h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ki][:60], r[gi]), []).append(float(r[vi]) / 1e3)
for k, v in d.items():
    print("%-62s %-14s n=%-3d med %.1f us" % (k[0], k[1], len(v), sorted(v)[len(v) // 2]))
