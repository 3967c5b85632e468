"""Where does the tensor-core mode drift from the fp32 CUDA-core mode?  One iteration on chosen objects, internal
taps of both engines compared buffer by buffer (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from catre_b200 import engine, synth

def keys2f(k):
    k = k.clone(); neg = k < 0; k[neg] = k[neg] ^ 0x7FFFFFFF
    return k.view(torch.float32)

n_all, seed = int(sys.argv[1]), int(sys.argv[2])
idx = torch.tensor([int(v) for v in sys.argv[3].split(",")])
w = synth.load_weights()
b = synth.make_batch(n_all, 1024, seed=seed)
sub = synth.Batch(*(getattr(b, f)[idx].contiguous() for f in ("pcl", "prior", "init_pose", "init_scale", "K", "obj_cls"))).to("cuda")
B = idx.numel(); S = 2 * B; P = 2048
taps = {}
for prec in ("fp32", "f16x3"):
    e = engine.Engine(1024, 16, prec, 0); e.load_weights(w)
    p, s = e.refine(sub.pcl, sub.prior, sub.init_pose, sub.init_scale, sub.K, 1)
    torch.cuda.synchronize()
    t = {"pose": p[1].cpu(), "scale": s[1].cpu()}
    for name, shape, isk in (("gmax_stn", (S, 1024), 1), ("t3", (S, 9), 0), ("gmax_fstn", (S, 1024), 1), ("gmax_pf", (S, 64), 1),
                             ("gmax_g", (S, 1024), 1), ("cset", (S, 512), 0), ("rot_partial", (B, 16, 6), 0)):
        try:
            v = e.debug_read(name, shape, torch.int32 if isk else torch.float32)
            t[name] = keys2f(v) if isk else v
        except Exception as ex:
            print("tap", name, "unavailable:", ex)
    taps[prec] = t
    e.close()
for name in taps["fp32"]:
    a, c = taps["fp32"][name].double(), taps["f16x3"][name].double()
    if name == "rot_partial":
        a, c = a.sum(1), c.sum(1) if c.shape == a.shape else c
    d = (a - c).abs().reshape(a.shape[0], -1)
    print(f"{name:12s} per-row max|diff| {[f'{v:.1e}' for v in d.amax(1).tolist()]}  (magnitude {a.abs().max():.2e})")
