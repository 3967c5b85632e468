"""Per-iteration divergence between the fp32 CUDA-core mode and the tensor-core mode on chosen objects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from catre_b200 import engine, synth
n_all, seed, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[4])
idx = torch.tensor([int(v) for v in sys.argv[3].split(",")])
w = synth.load_weights()
b = synth.make_batch(n_all, 1024, seed=seed)
sub = synth.Batch(*(getattr(b, f)[idx].contiguous() for f in ("pcl", "prior", "init_pose", "init_scale", "K", "obj_cls"))).to("cuda")
out = {}
for prec in ("fp32", "f16x3"):
    e = engine.Engine(1024, 16, prec, 0); e.load_weights(w)
    p, s = e.refine(sub.pcl, sub.prior, sub.init_pose, sub.init_scale, sub.K, K)
    torch.cuda.synchronize(); out[prec] = (p.cpu().double(), s.cpu().double()); e.close()
for it in range(1, K + 1):
    dr = (out["fp32"][0][it, :, :, :3] - out["f16x3"][0][it, :, :, :3]).abs().amax(dim=(1, 2))
    dt = (out["fp32"][0][it, :, :, 3] - out["f16x3"][0][it, :, :, 3]).abs().amax(dim=1)
    ds = (out["fp32"][1][it] - out["f16x3"][1][it]).abs().amax(dim=1)
    print(f"iter {it} dR {[f'{v:.1e}' for v in dr.tolist()]} dt {[f'{v:.1e}' for v in dt.tolist()]} ds {[f'{v:.1e}' for v in ds.tolist()]}")
# step-to-step motion of the fp32 result: has the refinement converged?
p = out["fp32"][0]
for it in range(1, K + 1):
    mv = (p[it, :, :, :3] - p[it - 1, :, :, :3]).abs().amax(dim=(1, 2))
    print(f"iter {it} |R_it - R_it-1| {[f'{v:.1e}' for v in mv.tolist()]}")
