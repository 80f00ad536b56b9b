"""The numpy oracle vs the frozen outputs of the unmodified reference (tests/golden/*.npz,
made by oracle/gen_golden.py).  This is what pins the oracle (task rule ③)."""
import numpy as np
import pytest

from oracle import philox, rollout
from oracle.cases import CASES, EVAL_CASES, NOISE_SEED

# fp32 re-association noise floor between two fp32 implementations of the same T-step
# recursion (SURVEY §7 step 1 measured 7e-5 / 1.2e-4 at |rnd|~135): tolerance is
# |delta| <= ATOL + RTOL*|ref| elementwise.
RTOL, ATOL = 2e-4, 2e-4


def _close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"max violation {err.max():.3e}; max abs diff {np.abs(a-b).max():.3e}"


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_train(golden, name):
    g = golden(name)
    spec, x0 = g["spec"], g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = philox.normal_noise(NOISE_SEED, B, T, d)
    x_T, rnd, _ = rollout.rollout(spec, x0, noise=noise)
    _close(x_T, g["train"]["x_T"])
    _close(rnd, g["train"]["rnd"])
    loss, nf = rollout.loss_from_rnd(rnd, spec["loss"]["method"], spec["loss"]["max_rnd"],
                                     spec["loss"]["traj_per_sample"])
    assert nf == g["train"]["n_filtered"]
    ref = g["train"]["loss"]
    assert abs(loss - ref) <= 1e-3 * (1 + abs(ref)), (loss, ref)


@pytest.mark.parametrize("name", list(EVAL_CASES))
def test_oracle_matches_reference_eval(golden, name):
    g = golden(name)
    x0 = g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = philox.normal_noise(NOISE_SEED, B, T, d)
    for j, (cw, rt) in enumerate(EVAL_CASES[name]):
        e = g[f"eval{j}"]
        spec = dict(g["spec"])
        spec["loss"] = dict(spec["loss"], train=False, compute_ito=bool(cw), return_traj=bool(rt))
        x_T, rnd, xs = rollout.rollout(spec, x0, noise=noise)
        _close(x_T, e["x_T"])
        _close(rnd, e["rnd"])
        if rt:
            assert xs.shape == e["xs"].shape
            _close(xs, e["xs"])
        res = rollout.results_from_rnd(rnd, cw)
        for k, v in e["log_norm_const_preds"].items():
            assert abs(res[k] - v) <= 1e-3 * (1 + abs(v)), (k, res[k], v)
        if cw:
            assert abs(res["eval/lv_loss"] - e["metrics"]["eval/lv_loss"]) <= 1e-3 * (1 + abs(e["metrics"]["eval/lv_loss"]))
            _close(res["weights"], e["weights"], rtol=2e-3, atol=1e-6)


def test_oracle_fp64_close_to_fp32(golden):
    """The fp64 run of the same restatement bounds the fp32 round-off of the recursion."""
    g = golden("dis_gmm50_lv")
    x0 = g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = philox.normal_noise(NOISE_SEED, B, T, d)
    x32, r32, _ = rollout.rollout(g["spec"], x0, noise=noise)
    x64, r64, _ = rollout.rollout(g["spec"], x0, noise=noise, dtype=np.float64)
    _close(x32, x64)
    _close(r32, r64)


@pytest.mark.parametrize("name", list(CASES))
def test_torch_port_matches_reference_train(golden, name):
    """oracle/torch_port.py (the CPU baseline bench.py times) is pinned to the same golden vectors."""
    from oracle import torch_port

    g = golden(name)
    spec, x0 = g["spec"], g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = philox.normal_noise(NOISE_SEED, B, T, d)
    x_T, rnd, _ = torch_port.rollout(spec, x0, noise=noise)
    _close(x_T, g["train"]["x_T"])
    _close(rnd, g["train"]["rnd"])


def test_oracle_langevin_matches_reference(golden):
    """oracle.rollout.langevin_integrate vs EulerIntegrator.integrate(LangevinSDE) of the unmodified reference."""
    from oracle.cases import ULA_CASES

    for name in ULA_CASES:
        g = golden(name)
        B, d = g["x0"].shape
        noise = philox.normal_noise(NOISE_SEED, B, g["timesteps"].shape[0] - 1, d)
        xs = rollout.langevin_integrate(g["target"], g["x0"], g["timesteps"], g["ts"], g["diff_coeff"], g["clip_score"], noise)
        assert xs.shape == g["xs"].shape
        _close(xs, g["xs"])


def test_oracle_affine_integrate_matches_reference(golden):
    """oracle.rollout.affine_integrate vs EulerIntegrator.integrate of the unmodified reference on VP / ScaledBM / ConstOU and
    on the PIS ControlledSDE (tests/golden/ou_*.npz)."""
    from oracle.cases import OU_CASES

    for name, case in OU_CASES.items():
        g = golden(name)
        B, d = g["x0"].shape
        noise = philox.normal_noise(NOISE_SEED, B, g["timesteps"].shape[0] - 1, d)
        mu, sigma, ctrl = rollout.ou_coefficients(case, g["timesteps"])
        xs = rollout.affine_integrate(mu, sigma, ctrl, g["x0"], g["timesteps"], g["ts"], noise)
        assert xs.shape == g["xs"].shape
        _close(xs, g["xs"])
