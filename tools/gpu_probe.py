"""Ad-hoc GPU probe (not part of the product): times the engine per kernel group and the PyTorch eager
restatement (oracle code on CUDA tensors = the 'reference single-GPU PyTorch forward' of BASELINE.json)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402
from oracle import catre_oracle  # noqa: E402


def cuda_time(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--prec", default="fp32,f16x3")
    ap.add_argument("--ref", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/probe.json")
    args = ap.parse_args()
    res = {"batch": args.batch, "n": args.n, "k": args.k, "gpu": torch.cuda.get_device_name(0)}
    w = catre_oracle.resize_conv_p(synth.load_weights(), args.n)
    b = synth.make_batch(args.batch, args.n, seed=2).to("cuda")
    outs = {}
    for prec in args.prec.split(","):
        try:
            eng = engine.Engine(args.n, args.batch, prec, 0)
            eng.load_weights(w)
        except Exception as ex:  # noqa
            res[prec] = {"error": str(ex)}
            continue
        fn = lambda: eng.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, args.k)  # noqa
        ms = cuda_time(fn)
        p, s = fn()
        outs[prec] = (p.cpu(), s.cpu())
        eng.profile_enable(True)
        eng.profile_reset()
        fn()
        torch.cuda.synchronize()
        prof = eng.profile()
        eng.profile_enable(False)
        res[prec] = {"ms": ms, "obj_per_s": args.batch / ms * 1e3, "launches": eng.last_launch_count(),
                     "profile_ms": {k: round(v[0], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
        print(prec, json.dumps(res[prec]), flush=True)
        eng.close()
    if args.ref:
        wd = {k: v.cuda() for k, v in w.items()}
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False  # torch default
            fn = lambda: catre_oracle.refine(wd, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, args.k)  # noqa
            try:
                ms = cuda_time(fn, warm=1, reps=3)
                p, s = fn()
                key = "torch_eager_tf32conv" if tf32 else "torch_eager_fp32"
                res[key] = {"ms": ms, "obj_per_s": args.batch / ms * 1e3}
                outs[key] = (p.cpu(), s.cpu())
                print(key, res[key], flush=True)
            except Exception as ex:  # noqa
                res["torch_eager_error"] = str(ex)
    if "torch_eager_fp32" in outs:
        rp, rs = outs["torch_eager_fp32"]
        for k, (p, s) in outs.items():
            res.setdefault("max_abs_diff_vs_torch_fp32", {})[k] = max((p - rp).abs().max().item(), (s - rs).abs().max().item())
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
