"""Per-kernel count / total / min / median / max of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys, collections, statistics
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = {k: i for i, k in enumerate(r)}
        continue
    if len(r) < len(hdr):
        continue
    name = r[hdr["Kernel Name"]].split("(")[0]
    unit = r[hdr["Metric Unit"]]
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    agg.setdefault(name, []).append(v_us)
if len(sys.argv) > 2:
    print(sys.argv[2])
print("%-64s %6s %12s %10s %10s %10s" % ("kernel", "count", "total_us", "min_us", "median_us", "max_us"))
for k, a in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-64s %6d %12.1f %10.1f %10.1f %10.1f" % (k[:64], len(a), sum(a), min(a), statistics.median(a), max(a)))
