"""GPU probe of the training step: device time per step of catre_train_step with the tiled and
the naive GEMM, at the reference's training batch sizes.  Prints one JSON line per configuration.
Usage (GPU box):  python tools/train_probe.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catre_b200 import engine, synth  # noqa: E402
from tests.test_train_gpu import y_symmetry_rotations  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 32]
    w = synth.load_weights()
    rots = y_symmetry_rotations()
    for B in sizes:
        batch, tgt = synth.make_train_batch(B, 1024, 3, round_robin_cls=True)
        d = batch.to("cuda")
        x_pm = (d.pcl - d.init_pose[:, :, 3].unsqueeze(1)).contiguous()
        tfd_pm = ((d.prior * d.init_scale.unsqueeze(1)) @ d.init_pose[:, :, :3].transpose(1, 2)).contiguous()
        gp, gs = tgt.gt_pose.cuda(), tgt.gt_scale.cuda()
        modes = os.environ.get("TRAIN_PROBE_MODES", "tc,tc-nograph,simt").split(",")
        for mode in modes:  # tc (default: tcgen05 GEMM for the large shapes), simt (64 x 64 CUDA-core tiles), v2, naive;
            # a "-nograph" suffix launches the chain kernel by kernel (CATRE_TRAIN_GRAPH=0) instead of replaying its CUDA graph
            ver = mode.replace("-nograph", "").replace("-carve", "").replace("-nofold", "").replace("-nolanes", "")
            os.environ["CATRE_TRAIN_LANES"] = "0" if "-nolanes" in mode else "1"
            os.environ["CATRE_TRAIN_FOLD_BIAS"] = "0" if "-nofold" in mode else "1"
            os.environ["CATRE_TRAIN_CARVEOUT"] = "1" if "-carve" in mode else "0"
            naive = "1" if ver == "naive" else "0"
            os.environ["CATRE_TRAIN_NAIVE_GEMM"] = naive
            os.environ["CATRE_TRAIN_GEMM"] = ver
            os.environ["CATRE_TRAIN_GRAPH"] = "0" if "-nograph" in mode else "1"
            eng = engine.Engine(1024, 8, "fp32", 0)
            eng.load_weights(w)
            step = lambda: eng.train_step(x_pm, tfd_pm, d.prior, d.init_pose, d.init_scale, d.K, gp, gs, tgt.sym_y.numpy(), rots)
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10 if naive == "0" else 2
            a.record()
            for _ in range(n):
                step()
            b.record()
            torch.cuda.synchronize()
            print(json.dumps({"probe": "train_step", "B": B, "N": 1024, "gemm": ver, "cuda_graph": "-nograph" not in mode, "max_shared_carveout": "-carve" in mode, "bias_grad_in_gemm": "-nofold" not in mode, "ts_head_lane": "-nolanes" not in mode,
                              "ms_per_step": a.elapsed_time(b) / n, "launches": eng.last_launch_count(),
                              "objects_per_s": B / (a.elapsed_time(b) / n / 1e3)}), flush=True)
            eng.close()


if __name__ == "__main__":
    main()
