"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections
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
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
print("%-60s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %6.1f%%" % (k[:60], a[0], a[1], 100 * a[1] / tot))
print("%-60s %8s %12.1f" % ("TOTAL", "", tot))
