"""Developer aid: kernel name, grid and duration of every launch of an `ncu --metrics gpu__time_duration.sum --csv` list."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if not h:
    print("".join(",".join(r) + "\n" for r in rows[:5])); sys.exit(0)
hd = rows[h[0]]; k = hd.index("Kernel Name"); v = hd.index("Metric Value"); g = hd.index("Grid Size"); u = hd.index("Metric Unit")
for r in rows[h[0] + 1:]:
    if len(r) > v: print(r[k][:40].ljust(42), r[g].ljust(16), r[v], r[u])
