"""Golden for the drop-in's parameter INITIALISATION: per-tensor SHA-256 of the state_dict the UNMODIFIED
reference builds with ``torch.manual_seed(1234); build_model_optimizer(cfg, is_test=True)`` (shipped config, N = 1024),
plus simple statistics.  ``catre_b200.dropin.reference_init`` must reproduce it bit for bit under the same seed
(tests/test_dropin_init.py).  Runs only where /root/reference exists.

Usage: python tests/golden/make_golden_init.py
"""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

SEED = 1234


def main():
    ref = "/root/reference"
    mg.install_shim(ref)
    from core.catre.models import CATRE_disR_shared as ref_model

    cfg = mg._to_attr(mg.load_py_config(os.path.join(ref, mg.CFG_REL)))
    cfg.MODEL.DEVICE = "cpu"
    cfg.MODEL.WEIGHTS = ""
    cfg.SOLVER.OPTIMIZER_NAME = cfg.SOLVER.OPTIMIZER_CFG["type"]
    cfg.SOLVER.BASE_LR = cfg.SOLVER.OPTIMIZER_CFG["lr"]
    torch.manual_seed(SEED)
    model, _ = ref_model.build_model_optimizer(cfg, is_test=True)
    out = {"seed": SEED, "torch": torch.__version__, "tensors": {}}
    for k, v in model.state_dict().items():
        out["tensors"][k] = {"sha256": hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest(), "shape": list(v.shape),
                             "mean": float(v.double().mean()), "std": float(v.double().std()) if v.numel() > 1 else 0.0}
    with open(os.path.join(HERE, "golden_init.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(out["tensors"]), "tensors")


if __name__ == "__main__":
    main()
