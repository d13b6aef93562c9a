"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/ncu_agg.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
agg = collections.OrderedDict(); hdr = None
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg.setdefault(d["Kernel Name"][:80], []).append(float(d["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-82s n=%4d avg=%9.1f us  total=%9.1f us (%4.1f%%)" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
