"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol the header
declares, the weight registry equals the reference checkpoint, and the drop-in module has the
reference's state_dict."""
import ctypes
import os
import re

import pytest
import torch

from catre_b200 import build, dropin, engine, synth

ZC = {"ZERO_CENTER_INPUT": True}  # the shipped config's value; the reference's base default (False) is refused

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()  # nvcc cross-compiles sm_100a without a GPU
    return engine.load_library()


def header_symbols():
    src = open(os.path.join(REPO, "include", "catre_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(catre_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_header_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"libcatre_b200.so does not export {s}"
    assert sorted(engine.SYMBOLS) == syms  # the ctypes binding covers the whole header


def test_version_and_weight_registry(lib):
    assert b"sm_100a" in lib.catre_version()
    names = engine.Engine.weight_names()
    w = synth.load_weights()
    assert names == list(w.keys())  # same 74 names, checkpoint order


def test_create_rejects_bad_config(lib):
    h = ctypes.c_void_p()
    for kw in (dict(n_obs=1000, n_prior=1000), dict(n_obs=1024, n_prior=500), dict(n_obs=1024, n_prior=1024, max_batch=0),
               dict(n_obs=1024, n_prior=1024, precision=9)):
        full = dict(n_obs=1024, n_prior=1024, max_batch=4, precision=0, device=0)
        full.update(kw)
        cfg = engine.CatreCfg(**full)
        rc = lib.catre_create(ctypes.byref(h), ctypes.byref(cfg))
        assert rc < 0 and not h.value
        assert len(lib.catre_last_error(None)) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_device_fails_loudly(lib):
    with pytest.raises(engine.CatreError):
        engine.Engine(1024, 4, "fp32")


def test_dropin_state_dict_matches_checkpoint():
    w = synth.load_weights()
    m = dropin.CatreB200(1024, 1024)
    sd = m.state_dict()
    assert list(sd.keys()) == list(w.keys())
    for k in w:
        assert tuple(sd[k].shape) == tuple(w[k].shape), k
    res = m.load_state_dict(w, strict=True)
    assert str(res) == "<All keys matched successfully>"
    assert sum(p.numel() for p in m.parameters()) == 4298711


def test_dropin_refuses_cpu_and_training():
    m = dropin.CatreB200(1024, 1024)
    b = synth.known_answer_inputs()
    x = b.pcl.permute(0, 2, 1)
    with pytest.raises(ValueError):  # the training forward needs the ground truth, like the reference's assert
        m(x, x, b.init_pose, b.init_scale, K_zoom=b.K, do_loss=True)
    with pytest.raises(engine.CatreError):  # no CPU path, no silent fallback -- inference ...
        m(x, x, b.init_pose, b.init_scale, K_zoom=b.K)
    with pytest.raises(engine.CatreError):  # ... and training alike
        m(x, x, b.init_pose, b.init_scale, K_zoom=b.K, gt_ego_rot=b.init_pose[:, :, :3], gt_trans=b.init_pose[:, :, 3],
          gt_scale=b.init_scale, obj_kps=b.prior, sym_info=[None], do_loss=True)


def test_training_cfg_and_sym_info_checks():
    import numpy as np

    assert dropin.check_loss_cfg({}) == (1.0, 1.0, 1.0, 1.0)  # absent keys = the shipped values
    assert dropin.check_loss_cfg({"MODEL": {"CATRE": {"LOSS_CFG": {"PM_LOSS_TYPE": "l1", "ROT_LW": 2.5}}}}) == (1.0, 2.5, 1.0, 1.0)
    with pytest.raises(NotImplementedError):  # a zero weight removes the term from the reference's loss dict
        dropin.check_loss_cfg({"MODEL": {"CATRE": {"LOSS_CFG": {"SCALE_LW": 0}}}})
    with pytest.raises(NotImplementedError):
        dropin.check_loss_cfg({"MODEL": {"CATRE": {"LOSS_CFG": {"ROT_LOSS_TYPE": "L2"}}}})
    with pytest.raises(NotImplementedError):
        dropin.check_loss_cfg({"MODEL": {"CATRE": {"USE_MTL": True}}})
    r = np.stack([np.eye(3, dtype=np.float32)] * 4)
    is_sym, rots = dropin.split_sym_info([None, r, None, torch.from_numpy(r)])
    assert is_sym == [False, True, False, True] and rots.shape == (4, 3, 3)
    assert dropin.split_sym_info([None])[1].shape == (0, 3, 3)
    with pytest.raises(NotImplementedError):
        dropin.split_sym_info([r, 2 * r])
    model, opt = dropin.build_model_optimizer({"INPUT": ZC, "MODEL": {"DEVICE": "cpu"}, "SOLVER": {"OPTIMIZER_CFG": {"type": "SGD", "lr": 1e-3}}},
                                              is_test=False)
    assert isinstance(opt, torch.optim.SGD) and model.training
    from catre_b200 import optim

    _, opt = dropin.build_model_optimizer({"INPUT": ZC, "MODEL": {"DEVICE": "cpu"}, "SOLVER": {"OPTIMIZER_CFG": {"type": "Ranger", "lr": 1e-4,
                                                                                                   "weight_decay": 0}}}, is_test=False)
    assert isinstance(opt, optim.FusedRanger) and opt.param_groups[0]["lr"] == 1e-4  # the shipped config's optimiser
    assert [len(g["params"]) for g in opt.param_groups] == [32, 28, 14]  # pcl_net, rot_head, ts_head -- the reference's groups
    cfg = {"INPUT": ZC, "MODEL": {"DEVICE": "cpu", "CATRE": {"ROT_HEAD": {"LR_MULT": 0.5}, "TS_HEAD": {"FREEZE": True}}},
           "SOLVER": {"BASE_LR": 2e-4, "OPTIMIZER_CFG": {"type": "SGD", "lr": 1e-3}}}
    model, opt = dropin.build_model_optimizer(cfg, is_test=False)
    assert [g["lr"] for g in opt.param_groups] == [2e-4, 1e-4] and not model.ts_head.fc_t.weight.requires_grad


def test_check_cfg():
    cfg = {"INPUT": {"NUM_PCL": 1024, "NUM_KPS": 1024, "ZERO_CENTER_INPUT": True},
           "MODEL": {"REFINE_SCLAE": True, "CATRE": {"ROT_HEAD": {"ROT_TYPE": "ego_rot6d", "INIT_CFG": {"num_points": 2048}}}}}
    assert dropin.check_cfg(cfg) == (1024, 1024)
    cfg["MODEL"]["CATRE"]["ROT_HEAD"]["ROT_TYPE"] = "allo_rot6d"
    with pytest.raises(NotImplementedError):
        dropin.check_cfg(cfg)
    with pytest.raises(NotImplementedError):
        dropin.build_model_optimizer({"INPUT": ZC, "MODEL": {"DEVICE": "cpu"}}, is_test=False)
    model, opt = dropin.build_model_optimizer({"INPUT": ZC, "MODEL": {"DEVICE": "cpu"}}, is_test=True)
    assert opt is None and not model.training
    # the fused K-loop hard-codes ZERO_CENTER_INPUT=True; the reference's base default is False, so an absent
    # key (or False) is refused; so are the constructor switches the kernels assume and a PRETRAINED pcl_net
    for bad in ({"MODEL": {"DEVICE": "cpu"}}, {"INPUT": {"ZERO_CENTER_INPUT": False}, "MODEL": {"DEVICE": "cpu"}},
                {"INPUT": ZC, "MODEL": {"CATRE": {"ROT_HEAD": {"INIT_CFG": {"point_bias": False}}}}},
                {"INPUT": ZC, "MODEL": {"CATRE": {"ROT_HEAD": {"INIT_CFG": {"norm_input": True}}}}},
                {"INPUT": ZC, "MODEL": {"CATRE": {"TS_HEAD": {"INIT_CFG": {"num_gn_groups": 16}}}}},
                {"INPUT": ZC, "MODEL": {"CATRE": {"TS_HEAD": {"INIT_CFG": {"dropout": True}}}}}):
        with pytest.raises(NotImplementedError):
            dropin.check_cfg(bad)
    with pytest.raises(NotImplementedError):
        dropin.build_model_optimizer({"INPUT": ZC, "MODEL": {"DEVICE": "cpu", "WEIGHTS": "", "CATRE": {"PCLNET": {"PRETRAINED": "x.pth"}}}},
                                     is_test=True)
