"""Full-size golden OUTPUTS from the UNMODIFIED reference (round 2; BASELINE.json configs at the sizes they name).

Same reference import as make_golden.py (shim for the missing third-party packages, shipped checkpoint), but at
sizes whose inputs would be tens of MB: only the reference's outputs are committed, together with the seed and a
SHA-256 of the input bytes.  The inputs are regenerated at test time by ``catre_b200.synth.make_batch`` (pure seeded
CPU torch; ``tests/test_oracle.py::test_full_size_inputs_reproduce`` checks the digest, so a drifting generator fails
loudly instead of silently comparing different inputs).

Besides the fp32 reference output (the PIN) every case stores the fp64 oracle output on the same inputs (the
YARDSTICK for the per-mode regression thresholds: distance to the exact answer, free of the reference's own fp32
noise, which grows ~1.4x per iteration on the rotationally symmetric categories).

  golden_full_<case>.npz   poses [K+1,B,3,4] f32, scales [K+1,B,3] f32 (reference), poses64 / scales64 (fp64 oracle)
  golden_full_index.json   case list: batch, n_pts, n_iter, seed, round_robin, input digest, reference CPU seconds

Usage:  python tests/golden/make_golden_full.py [--ref /root/reference] [--only case,case]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

# (name, batch, n_pts, n_iter, seed, round_robin)
CASES = [
    ("headline_b256_n1024_k4", 256, 1024, 4, 21, False),   # north_star headline; runs the B >= 128 FC path
    ("c4_b256_n2048_k8", 256, 2048, 8, 24, False),         # BASELINE configs[3] at its full size
    ("c5_b384_n1024_k4_mixed", 384, 1024, 4, 25, True),    # BASELINE configs[4]: 64 objects per category, round robin
]


def batch_digest(batch) -> str:
    h = hashlib.sha256()
    for f in ("pcl", "prior", "init_pose", "init_scale", "K"):
        h.update(getattr(batch, f).contiguous().numpy().tobytes())
    h.update(batch.obj_cls.numpy().astype(np.int64).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    import make_golden as mg
    from catre_b200 import synth
    from oracle import catre_oracle

    sd = torch.load(os.path.join(args.ref, mg.CKPT_REL), map_location="cpu")
    mg.install_shim(args.ref)
    fx = synth.load_fixtures()
    idx_path = os.path.join(HERE, "golden_full_index.json")
    index = {"torch": torch.__version__, "threads": args.threads, "cases": {}}
    if os.path.exists(idx_path):
        with open(idx_path) as f:
            index["cases"] = json.load(f)["cases"]
    only = set(filter(None, args.only.split(",")))
    for name, b, n, k, seed, rr in CASES:
        if only and name not in only:
            continue
        batch = synth.make_batch(b, n, seed, rr, fx)
        w_n = catre_oracle.resize_conv_p({k_: v.clone() for k_, v in sd.items()}, n)
        cfg, model, load_msg = mg.build_reference_model(args.ref, n, w_n)
        t0 = time.perf_counter()
        poses, scales = [], []
        step = 32  # objects are independent: run the reference in slices to bound the [B,1088,2N] activations
        for b0 in range(0, b, step):
            sl = synth.Batch(*(getattr(batch, f)[b0:b0 + step] for f in ("pcl", "prior", "init_pose", "init_scale", "K", "obj_cls")))
            p, s = mg.run_reference(cfg, model, sl, k)
            poses.append(p)
            scales.append(s)
        poses, scales = torch.cat(poses, 1), torch.cat(scales, 1)
        dt = time.perf_counter() - t0
        w64 = catre_oracle.cast_weights(w_n, torch.float64)
        p64, s64 = [], []
        for b0 in range(0, b, step):
            a = [getattr(batch, f)[b0:b0 + step].double() for f in ("pcl", "prior", "init_pose", "init_scale", "K")]
            p, s = catre_oracle.refine(w64, *a, k)
            p64.append(p)
            s64.append(s)
        p64, s64 = torch.cat(p64, 1), torch.cat(s64, 1)
        np.savez_compressed(os.path.join(HERE, f"golden_full_{name}.npz"), poses=poses.numpy(), scales=scales.numpy(),
                            poses64=p64.numpy(), scales64=s64.numpy())
        gap = max(torch.nan_to_num((poses.double() - p64).abs(), nan=0.0).max().item(),
                  torch.nan_to_num((scales.double() - s64).abs(), nan=0.0).max().item())
        nan_objects = sorted(set(torch.nonzero(~torch.isfinite(poses))[:, 1].tolist()))  # the reference's own NaN outputs
        index["cases"][name] = {"batch": b, "n_pts": n, "n_iter": k, "seed": seed, "round_robin": rr,
                                "input_sha256": batch_digest(batch), "ref_cpu_seconds": round(dt, 2),
                                "ref_obj_per_s": round(b / dt, 3), "ref_fp32_vs_fp64_max_abs": gap, "nan_objects": nan_objects,
                                "load_state_dict": load_msg}
        print(f"{name}: reference CPU {dt:.1f}s ({b / dt:.2f} obj/s); reference fp32 vs fp64 oracle max|d| {gap:.2e}", flush=True)
        with open(idx_path, "w") as f:
            json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
