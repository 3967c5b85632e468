"""Data-parallel training iteration on N GPUs (SURVEY.md 8(e) + 8(f) N4): the reference's loop shape
(core/catre/engine/engine.py:293-352) with the model wrapped as core/catre/main_catre.py:154-160 does --
DistributedDataParallel(find_unused_parameters=True) over NCCL -- around the drop-in (catre_train_step behind do_loss=True) and
catre_b200.optim.FusedRanger.  Every rank trains on its own B objects (weak scaling); wall clock around synchronised
iterations, max over ranks; rank 0 prints one JSON line.  Also checks that all ranks hold the same parameters afterwards and
that they equal a single-process run on the concatenated batch up to the all-reduce's summation order.
Usage (GPU box): python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
                 tools/train_ddp_probe.py [B]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import dropin, synth  # noqa: E402
from tests.test_train_gpu import y_symmetry_rotations  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    w = synth.load_weights()
    rots = y_symmetry_rotations()
    batch, tgt = synth.make_train_batch(B, 1024, 3 + rank, round_robin_cls=True)  # a different batch per rank
    d = batch.to(dev)
    gt_pose, gt_scale = tgt.gt_pose.to(dev), tgt.gt_scale.to(dev)
    sym_info = [rots if s else None for s in tgt.sym_y]
    # the re-posed points the reference's forward receives (batch_test.py:92-95), channel-major views like the reference passes
    x = (d.pcl - d.init_pose[:, :, 3].unsqueeze(1)).permute(0, 2, 1)
    tfd = ((d.prior * d.init_scale.unsqueeze(1)) @ d.init_pose[:, :, :3].transpose(1, 2)).permute(0, 2, 1)
    cfg = {"INPUT": {"ZERO_CENTER_INPUT": True}, "MODEL": {"DEVICE": dev},
           "SOLVER": {"OPTIMIZER_CFG": {"type": "Ranger", "lr": 1e-4, "weight_decay": 0}}}
    model, opt = dropin.build_model_optimizer(cfg, is_test=False, max_batch=max(8, B))
    model.load_state_dict(w, strict=True)
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False, find_unused_parameters=True)

    def iteration(i):
        _, loss_dict = net(x, tfd, init_pose=d.init_pose, init_scale=d.init_scale, K_zoom=d.K, gt_ego_rot=gt_pose[:, :, :3],
                           gt_trans=gt_pose[:, :, 3], gt_scale=gt_scale, obj_kps=d.prior, sym_info=sym_info, do_loss=True, cur_iter=1)
        sum(loss_dict.values()).backward()
        for p in model.parameters():
            if p.grad is not None:
                torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
        opt.step()
        opt.zero_grad(set_to_none=True)

    for i in range(4):
        iteration(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 10
    t0 = time.perf_counter()
    for i in range(n):
        iteration(4 + i)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / n], device=dev)
    # all ranks must hold the same parameters after 14 averaged steps
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    spread = torch.zeros(1, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        spread = (hi - lo).abs().max().reshape(1)
    if rank == 0:
        print(json.dumps({"probe": "train_iteration_ddp", "n_gpus": world, "objects_per_gpu": B, "N": 1024,
                          "ms_per_iteration_max_over_ranks": float(ms), "objects_per_s": world * B / (float(ms) / 1e3),
                          "max_parameter_spread_between_ranks": float(spread), "wrapper": "DistributedDataParallel(find_unused_parameters=True), nccl"}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
