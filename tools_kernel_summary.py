"""Summarise an `ncu --csv --log-file` launch list per kernel (mean of each metric).  usage: python tools_kernel_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[i0]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[i0 + 1:]:
    if len(r) < len(hdr):
        continue
    k = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("rsb::", "")
    agg[k][r[idx["Metric Name"]]].append(float(r[idx["Metric Value"]].replace(",", "")))
tot = sum(sum(m.get("gpu__time_duration.sum", [0])) for m in agg.values())
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [0])
    print("%-34s launches %4d  time share %5.1f%%" % (k[:34], len(t), 100 * sum(t) / max(tot, 1)))
    for name, v in m.items():
        print("      %-66s mean %14.3f" % (name, sum(v) / len(v)))
