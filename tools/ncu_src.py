"""Summarise an `ncu --page source --csv` dump: per kernel, top stall reasons and hottest SASS lines.
usage: ncu_src.py dump.csv [n_lines] [kernel_index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for ki, (a, b) in enumerate(zip(starts, starts[1:])):
    if only is not None and ki != only:
        continue
    hdr = rows[a + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
    I = lambda r, k: int(float(r[idx[k]] or 0))
    tot = sum(I(r, "# Samples") for r in data)
    print(f"==== [{ki}]", rows[a][1][:90], "samples", tot, "sass lines", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(I(r, s) for r in data) for s in stalls}
    print([(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
    for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:n]:
        st = sorted(((s, I(r, s)) for s in stalls), key=lambda kv: -kv[1])[:2]
        print(I(r, "# Samples"), I(r, "Instructions Executed"), r[idx["Source"]][:100], st)
