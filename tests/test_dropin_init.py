"""The drop-in's fresh parameters are the reference's: same distribution per tensor, drawn from
the global torch RNG in the reference's construction order, so ``torch.manual_seed`` governs them identically.
Golden: tests/golden/golden_init.json, made by the unmodified reference's build_model_optimizer (make_golden_init.py)."""
import hashlib
import json
import os

import torch

from catre_b200 import dropin, synth


def _golden():
    with open(os.path.join(synth.GOLDEN_DIR, "golden_init.json")) as f:
        return json.load(f)


def test_fresh_parameters_equal_the_references_bit_for_bit():
    g = _golden()
    torch.manual_seed(g["seed"])
    model = dropin.CatreB200(1024, 1024)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["tensors"].keys())  # names and ORDER of the reference's state_dict
    if torch.__version__ == g["torch"]:  # the bit pattern of torch's RNG streams is only pinned for the golden's version
        for k, meta in g["tensors"].items():
            assert list(sd[k].shape) == meta["shape"], k
            assert hashlib.sha256(sd[k].contiguous().numpy().tobytes()).hexdigest() == meta["sha256"], k
    for k, meta in g["tensors"].items():  # version-independent: the distributions
        v = sd[k].double()
        assert abs(float(v.mean()) - meta["mean"]) <= 1e-3 + 0.05 * abs(meta["mean"]), k
        if v.numel() > 64:
            assert abs(float(v.std()) - meta["std"]) <= 0.1 * meta["std"] + 1e-6, k


def test_seed_governs_the_initialisation():
    torch.manual_seed(7)
    a = dropin.CatreB200(1024, 1024).state_dict()
    torch.manual_seed(7)
    b = dropin.CatreB200(1024, 1024).state_dict()
    torch.manual_seed(8)
    c = dropin.CatreB200(1024, 1024).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["pcl_net.conv1.weight"], c["pcl_net.conv1.weight"])
    # the reference's scales: trunk ~ U(+-1/sqrt(fan_in)) (0.577 for conv1), heads N(0, 0.001), fc_t / fc_s N(0, 0.01)
    assert 0.25 < float(a["pcl_net.conv1.weight"].std()) < 0.4 and float(a["pcl_net.conv1.bias"].abs().max()) > 0.1
    assert 5e-4 < float(a["rot_head.rot_head_x.layers.0.weight"].std()) < 2e-3
    assert 5e-3 < float(a["ts_head.fc_t.weight"].std()) < 2e-2 and float(a["ts_head.fc_t.bias"].abs().max()) == 0.0
