import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
n_show = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tot = 0; n = 0
for r in rows[1:]:
    if r[idx['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[idx['Kernel Name']].replace('void catre::', '').replace('catre::', '')[:56]
    v = float(r[idx['Metric Value']].replace(',', '')); u = r[idx['Metric Unit']]
    us = v / 1e3 if u in ('ns', 'nsecond') else v
    tot += us; n += 1
    if n <= n_show: print(f"{us:8.1f} us  {name}  grid={r[idx['Grid Size']]}")
print("launches", n, "total us", round(tot, 1))
