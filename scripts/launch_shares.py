"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel.  usage: launch_shares.py file.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, agg, total = None, collections.OrderedDict(), 0.0
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    ms = v / 1e6 if d["Metric Unit"].startswith("n") else (v / 1e3 if d["Metric Unit"].startswith("u") else v)
    k = re.sub(r"\(.*", "", d["Kernel Name"].replace("void ", "").replace("tg::", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms; total += ms
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:40s} {n:6d} launches {ms:10.3f} ms {ms / total * 100:6.2f}%")
print(f"total {total:.3f} ms")
