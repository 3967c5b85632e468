"""Compact per-kernel table from an .ncu-rep (raw page): duration, DRAM bytes, tensor %, L2 %, registers."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tens%act"),
        ("sm__inst_executed_pipe_tensor.sum", "tensI"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__cycles_elapsed.avg", "cyc")]
print("id | kernel | " + " | ".join(c[1] for c in cols))
for r in data:
    name = r[idx["Kernel Name"]]
    name = name.replace("void catre::", "").split("(CUtensorMap")[0][:44]
    vals = []
    for k, _ in cols:
        v = r[idx[k]] if k in idx else ""
        u = units[idx[k]] if k in idx else ""
        try:
            f = float(v)
            if u == "byte": f /= 1e6
            elif u == "Kbyte": f /= 1e3
            elif u == "Gbyte": f *= 1e3
            elif u == "ms": f *= 1e3
            elif u == "ns": f /= 1e3
            v = f"{f:.1f}"
        except ValueError:
            pass
        vals.append(v)
    print(r[idx["ID"]], "|", name, "|", " | ".join(vals))
