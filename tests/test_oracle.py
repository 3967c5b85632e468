"""The oracle restatement against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the known-answer vector of SURVEY.md 8(c)."""
import pytest
import torch

from catre_b200 import synth
from oracle import catre_oracle
from tests import golden_util as gu


@pytest.fixture(scope="module")
def weights():
    return synth.load_weights()


def test_fixture_shapes(weights):
    fx = synth.load_fixtures()
    assert fx.priors.shape == (6, 1024, 3) and fx.init_pose.shape == (15374, 3, 4)
    assert len(weights) == 74
    assert sum(v.numel() for v in weights.values()) == 4298711  # SURVEY.md 8(a) parameter count
    assert weights["rot_head.rot_head_x.conv_p.weight"].shape == (1, 2048, 1)


def test_known_answer_vector(weights):
    """SURVEY.md 8(c) KAT, recorded from the reference during the survey (fp32, CPU)."""
    b = synth.known_answer_inputs()
    assert torch.allclose(b.prior[0, 0], torch.tensor([-0.1593079, -0.3997953, -0.3744943]), atol=1e-6)
    assert torch.allclose(b.pcl[0, 0], torch.tensor([-0.0218350, -0.0448422, 0.9800949]), atol=1e-6)
    poses, scales = catre_oracle.refine(weights, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 4)
    kat = {
        1: dict(t=(0.0175749, -0.0092993, 1.0167216), s=(0.1364689, 0.0875686, 0.1006906),
                R=(0.9818643, 0.0277701, 0.1875402, -0.0343420, 0.9989014, 0.0318844, -0.1864487, -0.0377467, 0.9817393)),
        4: dict(t=(0.0179817, -0.0090819, 1.0235741), s=(0.1317138, 0.0903642, 0.1021996),
                R=(0.9763225, 0.1059966, 0.1885720, -0.1304222, 0.9838986, 0.1222041, -0.1725824, -0.1439046, 0.9744264)),
    }
    for it, ref in kat.items():
        assert torch.allclose(poses[it, 0, :, 3], torch.tensor(ref["t"]), atol=2e-6)
        assert torch.allclose(scales[it, 0], torch.tensor(ref["s"]), atol=2e-6)
        assert torch.allclose(poses[it, 0, :, :3].flatten(), torch.tensor(ref["R"]), atol=2e-6)


@pytest.mark.parametrize("name", [n for n in gu.case_names() if n != "c2_b64_n1024_k4"])
def test_oracle_matches_reference_golden(weights, name):
    case = gu.load_case(name)
    w = catre_oracle.resize_conv_p(weights, case.n_pts)
    b = case.batch
    poses, scales = catre_oracle.refine(w, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, case.n_iter)
    e_r, e_t, e_s = gu.max_abs_err(poses, scales, case.poses, case.scales)
    # same torch ops on the same CPU: expected bit-close; 2e-6 leaves room for oneDNN blocking
    assert max(e_r, e_t, e_s) <= 2e-6, (e_r, e_t, e_s)


def test_oracle_fp64_close_to_fp32(weights):
    """fp32 noise floor of the path (~1.3e-6 on R at K=4)."""
    case = gu.load_case("ragged_b3_n1024_k2")
    b = case.batch
    w64 = catre_oracle.cast_weights(weights, torch.float64)
    poses, scales = catre_oracle.refine(w64, b.pcl.double(), b.prior.double(), b.init_pose.double(),
                                        b.init_scale.double(), b.K.double(), case.n_iter)
    e_r, e_t, e_s = gu.max_abs_err(poses, scales, case.poses, case.scales)
    assert max(e_r, e_t, e_s) <= 2e-5, (e_r, e_t, e_s)


def test_rot_is_orthonormal(weights):
    case = gu.load_case("c5s_b12_n1024_k4_mixed")
    r = case.poses[1:, :, :, :3].double()
    eye = torch.eye(3, dtype=torch.float64)
    assert (r @ r.transpose(-1, -2) - eye).abs().max() < 1e-5
    assert (torch.linalg.det(r) - 1).abs().max() < 1e-5


def test_seeded_generator_reproduces_the_golden_inputs():
    """synth.make_batch / make_train_batch are the SURVEY.md 8(d) generator: for the seeds the golden cases were made with they
    must keep producing the committed inputs byte for byte (bench.py, the probes and the training fixtures rely on it)."""
    import numpy as np

    from catre_b200 import synth
    from tests import golden_util as gu

    for name, meta in gu.index()["cases"].items():
        if meta["seed"] is None:
            continue
        z = np.load(f"{gu.GOLDEN_DIR}/golden_{name}.npz")
        b = synth.make_batch(meta["batch"], meta["n_pts"], meta["seed"], meta["round_robin"])
        assert np.array_equal(b.pcl.numpy(), z["pcl"]) and np.array_equal(b.init_pose.numpy(), z["init_pose"]), name
        assert np.array_equal(b.init_scale.numpy(), z["init_scale"]) and np.array_equal(b.obj_cls.numpy(), z["prior_cls"]), name
    zt = np.load(f"{gu.GOLDEN_DIR}/golden_train.npz")
    b, t = synth.make_train_batch(6, 1024, 11, round_robin_cls=True)
    assert np.array_equal(b.pcl.numpy(), zt["pcl"]) and np.array_equal(t.gt_pose.numpy(), zt["gt_pose"])
    assert np.array_equal(t.gt_scale.numpy(), zt["gt_scale"]) and np.array_equal(t.sym_y.numpy(), zt["sym_y"])


@pytest.mark.parametrize("name", gu.full_case_names())
def test_full_size_inputs_reproduce_and_oracle_matches_a_slice(weights, name):
    """The full-size goldens (tests/golden/make_golden_full.py) keep only the reference's outputs: the seeded generator
    must reproduce their inputs (SHA-256 recorded at generation time), and the oracle restatement must reproduce the
    reference's output on a slice of them (the whole case takes minutes on the CPU)."""
    case = gu.load_full_case(name)  # asserts the digest
    sl = slice(60, 68)  # the headline case's NaN object (index 64, initial t = 0) lies inside
    b = case.batch
    w = catre_oracle.resize_conv_p(weights, case.n_pts)
    k = min(case.n_iter, 2)
    poses, scales = catre_oracle.refine(w, b.pcl[sl], b.prior[sl], b.init_pose[sl], b.init_scale[sl], b.K[sl], k)
    e = gu.max_abs_err_nan_aware(poses, scales, case.poses[: k + 1, sl], case.scales[: k + 1, sl])
    assert max(e) <= 2e-6, e


def test_oracle_matches_reference_with_different_point_counts(weights):
    """NUM_PCL = 512 observed and NUM_KPS = 1024 prior points per object: the reference only ties conv_p to the SUM
    (conv_out_per_rot_head.py:112); golden from the unmodified reference built that way (make_golden_uneven.py)."""
    import numpy as np

    z = np.load(f"{gu.GOLDEN_DIR}/golden_uneven_b5_no512_np1024_k3.npz")
    n_obs, n_prior, batch, n_iter, seed = (int(v) for v in z["meta"])
    b = synth.make_batch(batch, n_obs, seed, n_prior=n_prior)
    assert b.pcl.shape == (batch, n_obs, 3) and b.prior.shape == (batch, n_prior, 3)
    w = catre_oracle.resize_conv_p(weights, n_obs, n_prior)
    assert w["rot_head.rot_head_x.conv_p.weight"].shape == (1, n_obs + n_prior, 1)
    poses, scales = catre_oracle.refine(w, b.pcl, b.prior, b.init_pose, b.init_scale, b.K, n_iter)
    e = gu.max_abs_err(poses, scales, torch.from_numpy(z["poses"]), torch.from_numpy(z["scales"]))
    assert max(e) <= 2e-6, e
