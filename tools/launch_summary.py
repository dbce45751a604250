"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...`) by kernel
name -> markdown table for profiles/.   usage: launch_summary.py launches.csv out.md "title line" """
import csv, sys
src, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
rows = [r for r in csv.reader(open(src)) if len(r) > 6]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
agg = {}
n = 0
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    except ValueError:
        continue
    name = r[ik].split("(")[0][:72]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v; n += 1
tot = sum(v for _, v in agg.values())
lines = ["# " + title, "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised:",
         "compare SHARES, not absolutes.  %d launches captured, total %.1f ms." % (n, tot), "",
         "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for name, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
    lines.append("| %s | %d | %.2f | %.1f%% |" % (name, c, v, 100 * v / tot))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:16]))
