"""End-to-end training iteration on the GPU: the reference's loop shape (core/catre/engine/engine.py:293-352:
forward with losses -> sum -> backward -> NaN guard -> optimiser step -> zero_grad) with
 (a) the drop-in model (catre_train_step behind do_loss=True) + catre_b200.optim.FusedRanger, and
 (b) the pure-torch restatement of the reference modules with autograd + the reference Ranger's per-tensor op sequence
     (tools/optim_probe.torch_ranger_step),
on the same synthetic batch.  Wall clock around synchronised iterations; prints one JSON line each.  Measurement tool.
Usage (GPU box): python tools/train_loop_probe.py [B ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import dropin, synth  # noqa: E402
from oracle import catre_oracle as co, train_oracle as to  # noqa: E402  (baseline leg only)
from tools.optim_probe import torch_ranger_step  # noqa: E402

ZC = {"ZERO_CENTER_INPUT": True}  # the shipped config's value; the reference's base default (False) is refused


def timed(fn, n_warm=3, n=8):
    for i in range(n_warm):
        fn(i + 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(n_warm + i + 1)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16]
    dev = "cuda"
    w = synth.load_weights()
    rots = to.y_symmetry_rotations()
    for B in sizes:
        batch, tgt = synth.make_train_batch(B, 1024, 3, round_robin_cls=True)
        d = batch.to(dev)
        gt_pose, gt_scale = tgt.gt_pose.to(dev), tgt.gt_scale.to(dev)
        sym_info = [rots if s else None for s in tgt.sym_y]
        x, tfd = co.update_points(d.pcl, d.prior, d.init_pose, d.init_scale)

        # (a) drop-in + fused optimiser
        cfg = {"INPUT": ZC, "MODEL": {"DEVICE": dev}, "SOLVER": {"OPTIMIZER_CFG": {"type": "Ranger", "lr": 1e-4, "weight_decay": 0}}}
        model, opt = dropin.build_model_optimizer(cfg, is_test=False, max_batch=max(8, B))
        model.load_state_dict(w, strict=True)

        def ours(_):
            _, loss_dict = model(x, tfd, init_pose=d.init_pose, init_scale=d.init_scale, K_zoom=d.K, gt_ego_rot=gt_pose[:, :, :3],
                                 gt_trans=gt_pose[:, :, 3], gt_scale=gt_scale, obj_kps=d.prior, sym_info=sym_info, do_loss=True, cur_iter=1)
            sum(loss_dict.values()).backward()
            for p in model.parameters():
                if p.grad is not None:
                    torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
            opt.step()
            opt.zero_grad(set_to_none=True)

        ms_a = None
        for flat in ("0", "1"):  # per-tensor gradient copies (default) vs the opt-in one-copy hand-off
            os.environ["CATRE_TRAIN_FLAT_GRADS"] = flat
            ms = timed(ours)
            ms_a = ms if ms_a is None else min(ms_a, ms)
            print(json.dumps({"probe": "train_iteration", "impl": "catre_b200 drop-in + FusedRanger", "flat_grads": flat == "1", "B": B,
                              "N": 1024, "ms_per_iteration": ms, "objects_per_s": B / (ms / 1e3)}), flush=True)
        os.environ["CATRE_TRAIN_FLAT_GRADS"] = "1"
        # where the iteration goes: the same loop with a device synchronisation after every phase (the phases then cannot
        # overlap, so their sum is an upper bound of the iteration above)
        ph = {"forward(do_loss)": 0.0, "backward": 0.0, "nan_guard": 0.0, "optimizer.step": 0.0, "zero_grad": 0.0}
        n_ph = 6
        for it in range(n_ph):
            marks = [time.perf_counter()]

            def mark():
                torch.cuda.synchronize()
                marks.append(time.perf_counter())

            _, loss_dict = model(x, tfd, init_pose=d.init_pose, init_scale=d.init_scale, K_zoom=d.K, gt_ego_rot=gt_pose[:, :, :3],
                                 gt_trans=gt_pose[:, :, 3], gt_scale=gt_scale, obj_kps=d.prior, sym_info=sym_info, do_loss=True, cur_iter=1)
            mark()
            sum(loss_dict.values()).backward()
            mark()
            for p in model.parameters():
                if p.grad is not None:
                    torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
            mark()
            opt.step()
            mark()
            opt.zero_grad(set_to_none=True)
            mark()
            for k, a_, b_ in zip(ph, marks[:-1], marks[1:]):
                ph[k] += (b_ - a_) * 1e3 / n_ph
        print(json.dumps({"probe": "train_iteration_phases", "B": B, "ms": {k: round(v, 3) for k, v in ph.items()},
                          "sum_ms": round(sum(ph.values()), 3)}), flush=True)
        os.environ["CATRE_TRAIN_FLAT_GRADS"] = "0"

        # (b) torch modules (restatement) + per-tensor Ranger ops, TF32 as PyTorch defaults it and off
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ps = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in w.items()}
            used = [p for k, p in ps.items() if k not in to.UNUSED]
            state = {p: (torch.zeros_like(p), torch.zeros_like(p), p.detach().clone()) for p in used}

            def ref(step):
                pose, scale = co.forward_once(ps, x, tfd, d.init_pose, d.init_scale, d.K)
                losses = to.catre_loss(pose[:, :3, :3], pose[:, :3, 3], scale, gt_pose[:, :3, :3], gt_pose[:, :3, 3], gt_scale, d.prior, sym_info)
                sum(losses.values()).backward()
                with torch.no_grad():
                    torch_ranger_step(used, state, step)
                for p in used:
                    p.grad = None

            ms_b = timed(ref)
            print(json.dumps({"probe": "train_iteration", "impl": "torch eager + per-tensor Ranger", "tf32": tf32, "B": B, "N": 1024,
                              "ms_per_iteration": ms_b, "objects_per_s": B / (ms_b / 1e3), "speedup_of_catre_b200": ms_b / ms_a}), flush=True)


if __name__ == "__main__":
    main()
