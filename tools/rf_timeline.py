"""Debug: phase timeline (clock64 deltas) of the fused rot kernel's CTA 0, first 8 work items."""
import os, sys
os.environ["CATRE_RF_DEBUG"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = engine.Engine(1024, B, "bf16x3", 0)
eng.load_weights(synth.load_weights())
b = synth.make_batch(B, 1024, seed=2).to("cuda")
for _ in range(2):
    eng.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 1)
torch.cuda.synchronize()
d = eng.debug_read("rf_dbg", (8, 32), torch.int64)
names = {0: "mma:start", 1: "mma:d0_empty", 2: "mma:L0 operands", 3: "mma:d1_empty", 16: "mma:L1 issued",
         17: "epi:d0_full", 18: "epi:slab0", 19: "epi:slab1", 20: "epi:slab2", 21: "epi:slab3", 22: "epi:d1_full", 23: "epi:done"}
for ks in range(4):
    names[4 + ks * 3] = f"mma:u_full{ks}"
    names[5 + ks * 3] = f"mma:w1[{ks}][0]"
    names[6 + ks * 3] = f"mma:w1[{ks}][1]"
t0 = int(d[0, 0])
for i in range(8):
    ev = sorted((int(d[i, k]) - t0, names[k]) for k in names if int(d[i, k]) != 0)
    print(f"--- work item {i}")
    prev = None
    for t, n in ev:
        print(f"  {t:8d}  (+{0 if prev is None else t - prev:6d})  {n}")
        prev = t
