"""Developer aid: source-line hot spots of one kernel from an .ncu-rep captured with --import-source on
(-lineinfo build).  usage: python scripts/ncu_lines.py x.ncu-rep [top]  -> file:line, % of stall samples,
warp instructions executed, source text."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        i_s, i_x = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(r) >= len(hdr) and r[0].isdigit():
        num = lambda v: float(v) if v not in ("", "-") else 0.0
        lines.append((fname, int(r[0]), num(r[i_s]), num(r[i_x]), r[1].strip()))
tot_s = sum(l[2] for l in lines) or 1.0
tot_x = sum(l[3] for l in lines) or 1.0
print("total samples %d, warp instructions %d" % (tot_s, tot_x))
for f, n, s, x, src in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%-16s %5d  %5.1f%% samples  %5.1f%% inst  %s" % (f, n, 100 * s / tot_s, 100 * x / tot_x, src[:100]))
