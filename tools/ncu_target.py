"""Short target for `ncu --set full`: two refine calls of the bench workload (B=64, N=1024, K=4)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = engine.Engine(1024, B, prec, 0)
eng.load_weights(synth.load_weights())
b = synth.make_batch(B, 1024, seed=2).to("cuda")
for _ in range(2):
    eng.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 4)
torch.cuda.synchronize()
print("launches", eng.last_launch_count())
