"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of the training probe: time per kernel over the LAST
training step in the file (between the last two KLoss launches), and the GEMM time by grid size.
Usage: python tools/train_launch_summary.py gpurun_out/train_launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
rows = [r for r in rows[hdr + 1:] if len(r) >= 15]
names = [re.sub(r"\(.*", "", r[4].replace("void ", "").replace("catre_train::", "")) for r in rows]
loss = [i for i, n in enumerate(names) if "KLoss>" in n]
if len(loss) >= 2:
    lo, hi = loss[-2], loss[-1]  # one full step: loss .. backward .. forward .. loss
else:
    lo, hi = 0, len(rows)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r, n in zip(rows[lo:hi], names[lo:hi]):
    us = float(r[-1]) / 1e3
    agg[n][0] += 1
    agg[n][1] += us
    tot += us
print(f"one step: {hi - lo} launches, {tot:.1f} us of kernel time (serialised, cold caches)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} {v[0]:4d} launches {v[1]:9.1f} us")
