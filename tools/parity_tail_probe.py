"""How far apart are the engine's arithmetic paths on a LARGE sample?  fp32 CUDA-core mode (<= 3e-6 from the
reference) vs f16x3 with the cluster FC kernels (launches < 128 objects) vs f16x3 with the tensor-core FC chain
(launches >= 128 objects).  Prints max |delta| on R, t, s per pair and the worst objects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from catre_b200 import engine, synth

n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = synth.load_weights()
b = synth.make_batch(n_obj, 1024, seed=77).to("cuda")
res = {}
for name, prec, mb in (("fp32", "fp32", 64), ("f16x3_cluster_fc", "f16x3", 64), ("f16x3_tc_fc", "f16x3", 256)):
    e = engine.Engine(1024, mb, prec, 0); e.load_weights(w)
    p, s = e.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, K)
    torch.cuda.synchronize()
    res[name] = (p.cpu(), s.cpu()); e.close()
names = list(res)
for i in range(3):
    for j in range(i + 1, 3):
        (pa, sa), (pb, sb) = res[names[i]], res[names[j]]
        dr = (pa[..., :3] - pb[..., :3]).abs().amax(dim=(0, 2, 3))
        dt = (pa[..., 3] - pb[..., 3]).abs().amax(dim=(0, 2))
        ds = (sa - sb).abs().amax(dim=(0, 2))
        worst = torch.topk(dr, 3)
        print(f"{names[i]:18s} vs {names[j]:18s} dR {dr.max():.2e} dt {dt.max():.2e} ds {ds.max():.2e}  "
              f"dR median {dr.median():.1e} p99 {dr.kthvalue(int(0.99 * n_obj)).values:.1e} worst objs {worst.indices.tolist()} {[f'{v:.1e}' for v in worst.values.tolist()]}")
