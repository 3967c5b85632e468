"""Golden for NUM_PCL != NUM_KPS (VERDICT r1 missing #6): the UNMODIFIED reference model built with 512 observed and 1024
prior points per object (conv_p = Conv1d(1536, 1, 1), core/catre/models/heads/conv_out_per_rot_head.py:112), run through
the evaluator's K-loop on seeded inputs.  Weights: the shipped checkpoint with the two conv_p.weight re-sized per half
(synth.resize_conv_p(w, 512, 1024)).  Runs only in the build container.

Usage: python tests/golden/make_golden_uneven.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

N_OBS, N_PRIOR, BATCH, N_ITER, SEED = 512, 1024, 5, 3, 44


def main():
    ref = "/root/reference"
    mg.install_shim(ref)
    from core.catre.models import CATRE_disR_shared as ref_model

    from catre_b200 import synth

    sd = torch.load(os.path.join(ref, mg.CKPT_REL), map_location="cpu")
    w = synth.resize_conv_p({k: v.clone() for k, v in sd.items()}, N_OBS, N_PRIOR)
    cfg = mg._to_attr(mg.load_py_config(os.path.join(ref, mg.CFG_REL)))
    cfg.MODEL.DEVICE, cfg.MODEL.WEIGHTS = "cpu", "fixture"
    cfg.SOLVER.OPTIMIZER_NAME, cfg.SOLVER.BASE_LR = cfg.SOLVER.OPTIMIZER_CFG["type"], cfg.SOLVER.OPTIMIZER_CFG["lr"]
    cfg.MODEL.CATRE.ROT_HEAD.INIT_CFG.num_points = N_OBS + N_PRIOR
    cfg.MODEL.CATRE.PCLNET.INIT_CFG.num_points = N_OBS
    cfg.INPUT.NUM_PCL, cfg.INPUT.NUM_KPS = N_OBS, N_PRIOR
    model, _ = ref_model.build_model_optimizer(cfg, is_test=True)
    print(model.load_state_dict(w, strict=True))
    model.eval()
    batch = synth.make_batch(BATCH, N_OBS, SEED, n_prior=N_PRIOR)
    poses, scales = mg.run_reference(cfg, model, batch, N_ITER)
    np.savez_compressed(os.path.join(HERE, "golden_uneven_b5_no512_np1024_k3.npz"), poses=poses.numpy(), scales=scales.numpy(),
                        meta=np.array([N_OBS, N_PRIOR, BATCH, N_ITER, SEED]))
    print("R[-1,0]", poses[-1, 0, :, :3].flatten().tolist())


if __name__ == "__main__":
    main()
