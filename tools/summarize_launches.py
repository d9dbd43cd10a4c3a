"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches_iter.csv [skip_first_n_launches]"""
import collections, csv, io, sys
txt = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(txt))))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print("%d launches, %.1f us total (cold-cache, serialised)" % (len(rows), tot))
print("%7s %11s %9s %6s  %s" % ("count", "total us", "avg us", "share", "kernel"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%7d %11.1f %9.1f %5.1f%%  %s" % (v[0], v[1], v[1] / v[0], 100 * v[1] / tot, k))
