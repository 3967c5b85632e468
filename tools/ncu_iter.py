"""Per-kernel table of ONE refinement iteration from an `ncu --set full` report (tensor-core modes), with the engine's
group names attached by launch order, and the DRAM traffic per launch merged into profiles/ncu_traffic.json (the source of
bench.py's roofline.traffic).

usage: python tools/ncu_iter.py report.ncu-rep <batch> [--update-traffic]
Capture:  ncu --set full --clock-control none --import-source on -s 21 -c 21 -o gpurun_out/prof_iter python tools/ncu_target.py f16x3 <batch>
(21 launches per iteration; the first iteration is skipped.)  Cold-cache, serialised replays: compare shares, not absolutes.
"""
import csv
import json
import os
import subprocess
import sys

ORDER = [("front3", "iter_head_kernel"), ("stn_conv3_max", "enc_fused"), ("tnet_fc", "fc_chain"),
         ("front3", "front3_split"), ("fstn_conv1", "tc_gemm_kernel<1, 2, 64"), ("fstn_conv3_max", "enc_fused"), ("tnet_fc", "fc_chain"),
         ("tnet_fc", "fc_tiled"), ("feat_transform", "tc_gemm_kernel<1, 5, 64"), ("conv2", "tc_gemm_kernel<1, 2, 128"),
         ("conv3", "tc_gemm_kernel<1, 2, 128"), ("conv4_max", "tc_gemm_kernel<0, 0, 256"), ("rot_gfeat", "fc_chain"), ("ts_pose", "ts_head"),
         ("rot_layer0", "tc_gemm_kernel<0, 3, 256"), ("gn_finalize", "gn_finalize_set"), ("rot_fused", "rot_fused"),
         ("ts_pose", "pose_update"), ("update_points", "update_points_kernel")]

rep, batch = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, k, scale_bytes=False):
    if k not in idx:
        return 0.0
    try:
        f = float(r[idx[k]])
    except ValueError:
        return 0.0
    u = units[idx[k]]
    if scale_bytes:
        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if k == "gpu__time_duration.sum":
        f *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return f


# align the captured launches with ORDER (the ts head runs on a side stream, so its position may float)
names = [r[idx["Kernel Name"]] for r in data]
seq, j = [], 0
for r, n in zip(data, names):
    grp = None
    for g, pat in ORDER:
        if pat in n:
            grp = g if pat not in ("tc_gemm_kernel<1, 2, 128", "fc_chain", "enc_fused", "front3_split") else None
            if pat == "iter_head_kernel":
                grp = "front3(pose update + points + stn.conv1)"
            break
    seq.append(grp)
# order-dependent ones: resolve by occurrence count
cnt = {}
for i, n in enumerate(names):
    if seq[i] is not None:
        continue
    for pat, groups in (("tc_gemm_kernel<1, 2, 128", ["conv2", "conv3"]), ("fc_chain", ["tnet_fc(stn)", "tnet_fc(fstn fc1+fc2)", "rot_gfeat(cset+ts0)"]),
                        ("enc_fused", ["stn_conv3_max", "fstn_conv3_max"]), ("front3_split", ["front3(T3 + conv1)"])):
        if pat in n:
            k = cnt.get(pat, 0)
            seq[i] = groups[k % len(groups)]
            cnt[pat] = k + 1
print(f"# one refinement iteration, B={batch}, N=1024, f16x3; ncu --set full (cold caches, serialised): shares, not absolutes")
print("group | kernel | us | dram rd MB | dram wr MB | tensor-pipe % | dram % | L2 % | ipc | regs | grid")
traffic, total_us = {}, 0.0
for r, n, g in zip(data, names, seq):
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
    total_us += us
    short = n.replace("void catre::", "").replace("catre::", "").split("(")[0][:40]
    print(f"{g or '?':24s} | {short:40s} | {us:7.1f} | {rd / 1e6:7.1f} | {wr / 1e6:7.1f} | "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} | "
          f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} | {val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} | "
          f"{val(r, 'sm__inst_executed.avg.per_cycle_elapsed'):4.2f} | {int(val(r, 'launch__registers_per_thread'))} | {int(val(r, 'launch__grid_size'))}")
    if g:
        key = g.split("(")[0]
        traffic[key] = traffic.get(key, 0.0) + rd + wr
print(f"# sum of kernel durations: {total_us:.1f} us; DRAM traffic of the iteration: {sum(traffic.values()) / 1e6:.1f} MB")
if "--update-traffic" in sys.argv:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    try:
        cur = json.load(open(p))
    except Exception:
        cur = {}
    for g, b in traffic.items():
        cur.setdefault(g, {})[str(batch)] = b
    json.dump(cur, open(p, "w"), indent=1)
