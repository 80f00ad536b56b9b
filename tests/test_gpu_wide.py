"""-m gpu: the WIDE engine (d > 64 or a NICE target: layered tcgen05 GEMMs, csrc/sdes_wide.cu) beyond the golden
cases that tests/test_gpu_parity.py already runs for it (dds_nice16_lv, dds_nice196_lv, dis_gauss100_lv):
ragged batches against the numpy oracle, in-kernel noise == staged noise, shard invariance, the two engines
(tcgen05 GEMM / CUDA-core GEMM on the same operand images) against each other, every loss / control kind on a wide
state, and BASELINE cfg5's full layer widths (d = 784, mid = 1000, hidden = 5) against the oracle on a few steps.

Tolerance as in test_gpu_parity.py: |delta| <= 2e-4 + 2e-4 |ref| on x_T and rnd (fp32 path; the GEMMs run as
three bf16 passes over hi/lo-split operands, 16 significant bits per operand, fp32 accumulation)."""
import copy

import numpy as np
import pytest
import torch

from oracle import philox, rollout as oracle_rollout
from oracle.cases import NOISE_SEED
from sdes_test_helpers import assert_close, build_from_spec

pytestmark = pytest.mark.gpu
RTOL = ATOL = 2e-4
ENGINES = ["simt", "tcgen05"]


def _dev():
    return torch.device("cuda:0")


def _kw(b):
    return {"terminal_unnorm_log_prob": b["terminal"], b["second_name"]: b["second"]}


def _run(spec, x0, noise, engine, **sim_kw):
    b = build_from_spec(spec, _dev(), engine=engine)
    method = spec["loss"]["method"]
    return b["loss"].simulate(b["ts"], torch.from_numpy(x0).to(_dev()), compute_ito_int=method != "kl",
                              noise=None if noise is None else torch.from_numpy(noise).to(_dev()), **sim_kw, **_kw(b))


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name,B", [("dds_nice16_lv", 1), ("dds_nice16_lv", 130), ("dds_nice196_lv", 300), ("dis_gauss100_lv", 257)])
def test_wide_ragged_batches_match_oracle(golden, name, B, engine):
    g = golden(name)
    spec = g["spec"]
    d, T = spec["dim"], g["ts"].shape[0] - 1
    x0 = np.random.default_rng(B).standard_normal((B, d)).astype(np.float32)
    noise = philox.normal_noise(NOISE_SEED + 2, B, T, d)
    want_x, want_r, _ = oracle_rollout.rollout(spec, x0, noise=noise)
    x_T, rnd, _ = _run(spec, x0, noise, engine)
    assert_close(x_T.cpu().numpy(), want_x, RTOL, ATOL, "x_T")
    assert_close(rnd.cpu().numpy(), want_r, RTOL, ATOL, "rnd")


def _variant(spec, **changes):
    s = copy.deepcopy(spec)
    for k, v in changes.items():
        a, b = k.split("__")
        s[a][b] = v
    return s


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("variant", ["tr_eval_traj", "lerp_prior", "lerp_target", "clipped", "gmm3", "ref_sde_nice", "tr_kl",
                                     "multiwell", "funnel"])
def test_wide_loss_and_control_kinds_match_oracle(golden, variant, engine):
    """Every loss / control family on a wide state (the goldens pin DDS+ScoreCtrl on NICE and DIS+LerpCtrl on a Gaussian)."""
    g = golden("dis_gauss100_lv")
    spec = g["spec"]
    sim_kw = {}
    if variant == "tr_eval_traj":    # eval semantics: rnd -= int div, trajectory returned (losses/oc.py:210-211, :221-229)
        spec = _variant(spec, loss__train=False, loss__return_traj=True)
        sim_kw = dict(train=False, return_traj=True)
    elif variant == "lerp_prior":
        spec = _variant(spec, ctrl__kind="lerp_prior")
    elif variant == "lerp_target":
        spec = _variant(spec, ctrl__kind="lerp_target")
    elif variant == "clipped":
        spec = _variant(spec, ctrl__kind="clipped")
        spec["gate"] = None
    elif variant == "tr_kl":         # kl training: rnd0 = 0, no Ito term
        spec = _variant(spec, loss__method="kl", loss__compute_ito=False)
    elif variant == "gmm3":          # a genuine mixture on a wide state (3 components)
        rng = np.random.default_rng(0)
        d = spec["dim"]
        spec = copy.deepcopy(spec)
        spec["target"] = {"kind": "gmm", "loc": rng.uniform(-2, 2, (3, d)).astype(np.float32),
                          "scale": rng.uniform(0.7, 1.5, (3, d)).astype(np.float32),
                          "log_weights": np.log(np.array([0.2, 0.5, 0.3], np.float32)), "log_norm_const": 0.0,
                          "clip_target": None}
    elif variant == "multiwell":     # MultiWell on a wide state: 5 double wells + 95 Gaussian dims
        spec = copy.deepcopy(spec)
        spec["target"] = {"kind": "multiwell", "n_dw": 5, "separation": 2.0, "shift": 0.5, "clip_target": None}
    elif variant == "funnel":        # Funnel in d = 100 (variance 9).  Two steps only: in d = 100 the funnel score drives
        spec = copy.deepcopy(spec)   # x_0 to large negative values where clip(-x_j exp(-x_0)) acts like sign(x_j) with
        spec["target"] = {"kind": "funnel", "variance": 9.0, "log_norm_const": 0.0, "clip_target": None}  # slope 1e4 —
        spec["ts"] = np.asarray(spec["ts"], np.float32)[:3]  # round-off then decides trajectories in ANY fp32 code
    elif variant == "ref_sde_nice":  # ReferenceSDELoss (PIS-style, ScoreCtrl, ScaledBM) with a NICE target
        gn = golden("dds_nice196_lv")["spec"]
        spec = copy.deepcopy(gn)
        spec["loss"] = {"kind": "reference_sde", "method": "lv", "train": True, "compute_ito": True, "return_traj": False,
                        "max_rnd": None, "traj_per_sample": 1, "reference_ctrl": False}
        spec["sde"] = {"kind": "const_ou", "drift_coeff": 0.0, "diff_coeff": 0.7, "terminal_t": 1.0, "sign": 1.0}
        spec["ts"] = np.linspace(0.0, 1.0, 21).astype(np.float32)
        spec["prior"] = None
    d, T = spec["dim"], np.asarray(spec["ts"]).shape[0] - 1
    B = 140
    x0 = (0.5 * np.random.default_rng(3).standard_normal((B, d))).astype(np.float32)
    noise = philox.normal_noise(NOISE_SEED + 3, B, T, d)
    want_x, want_r, want_xs = oracle_rollout.rollout(spec, x0, noise=noise)
    x_T, rnd, xs = _run(spec, x0, noise, engine, **sim_kw)
    assert_close(x_T.cpu().numpy(), want_x, RTOL, ATOL, "x_T")
    assert_close(rnd.cpu().numpy(), want_r, RTOL, ATOL, "rnd")
    if want_xs is not None:
        assert_close(xs.cpu().numpy(), want_xs, RTOL, ATOL, "xs")


@pytest.mark.parametrize("engine", ENGINES)
def test_wide_fused_noise_equals_staged_noise_and_shards(golden, engine):
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.spec import extract_spec

    g = golden("dds_nice196_lv")
    b = build_from_spec(g["spec"], _dev(), engine=engine)
    B, d, T = 200, g["spec"]["dim"], g["ts"].shape[0] - 1
    x0 = torch.randn(B, d, device=_dev(), generator=torch.Generator(_dev()).manual_seed(3))
    spec = extract_spec(b["loss"], "exp_integrator", b["ts"], b["terminal"], b["second"], train=True, compute_ito=True)
    seed, off = 0xABCDEF0123, 1000
    xa, ra, _ = eng.rollout(spec, x0, seed=seed, traj_offset=off, engine=engine)
    noise = eng.philox_normal(seed, off, B, T, d, _dev())
    xb, rb, _ = eng.rollout(spec, x0, noise=noise, engine=engine)
    assert torch.equal(xa, xb) and torch.equal(ra, rb)
    h = 72  # shards that are not multiples of the 128-row tile
    x1, r1, _ = eng.rollout(spec, x0[:h], seed=seed, traj_offset=off, engine=engine)
    x2, r2, _ = eng.rollout(spec, x0[h:], seed=seed, traj_offset=off + h, engine=engine)
    assert torch.equal(torch.cat([x1, x2]), xa) and torch.equal(torch.cat([r1, r2]), ra)


def _cfg5_spec(golden, dim, mid, hidden, T):
    """BASELINE cfg5 shapes: NiceModel(coupling=4, in_out_dim=dim, mid_dim=mid, hidden=hidden), DDS + ScoreCtrl, lv."""
    spec = copy.deepcopy(golden("dds_nice196_lv")["spec"])
    rng = np.random.default_rng(dim + mid)
    half = dim // 2

    def lin(n_out, n_in):
        bound = 1.0 / np.sqrt(n_in)  # nn.Linear default init range
        return (rng.uniform(-bound, bound, (n_out, n_in)).astype(np.float32), rng.uniform(-bound, bound, n_out).astype(np.float32))

    spec["dim"] = dim
    spec["target"] = {"kind": "nice", "log_norm_const": 0.0, "clip_target": None,
                      "scale": (0.2 * rng.standard_normal(dim)).astype(np.float32),
                      "couplings": [{"mask_config": (1 + c) % 2,
                                     "layers": [lin(mid, half)] + [lin(mid, mid) for _ in range(hidden - 1)] + [lin(half, mid)]}
                                    for c in range(4)]}
    m = spec["mlp"]
    m["in_w"] = (rng.standard_normal((64, dim)) / np.sqrt(dim)).astype(np.float32)
    m["out_w"] = (0.05 * rng.standard_normal((dim, 64))).astype(np.float32)
    m["out_b"] = (0.05 * rng.standard_normal(dim)).astype(np.float32)
    for k in ("prior", "ref"):
        if spec.get(k) is not None:
            spec[k] = {"loc": np.zeros(dim, np.float32), "scale": np.ones(dim, np.float32)}
    spec["ts"] = np.asarray(spec["ts"], np.float32)[: T + 1]
    return spec


@pytest.mark.parametrize("engine", ENGINES)
def test_cfg5_full_layer_widths_match_oracle(golden, engine):
    """d = 784, mid = 1000, hidden = 5 (scripts/train_nice.py:67-78 with --resize 28): 896-wide planar state, 1024-wide
    padded layers, K loops of 16 pipeline stages — three steps of 200 trajectories against the numpy oracle."""
    spec = _cfg5_spec(golden, 784, 1000, 5, T=3)
    B, d, T = 200, 784, 3
    x0 = np.random.default_rng(1).standard_normal((B, d)).astype(np.float32)
    noise = philox.normal_noise(NOISE_SEED + 4, B, T, d)
    want_x, want_r, _ = oracle_rollout.rollout(spec, x0, noise=noise)
    x_T, rnd, _ = _run(spec, x0, noise, engine)
    assert_close(x_T.cpu().numpy(), want_x, RTOL, ATOL, "x_T")
    assert_close(rnd.cpu().numpy(), want_r, RTOL, ATOL, "rnd")


def test_wide_engines_agree_at_scale(golden):
    """4 096 trajectories (32 row tiles) on d = 196, mid = 256: tcgen05 GEMMs vs CUDA-core GEMMs on the same images,
    deterministic reruns, finite outputs."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.spec import extract_spec

    spec_d = _cfg5_spec(golden, 196, 256, 3, T=6)
    outs = {}
    for engine in ENGINES:
        b = build_from_spec(spec_d, _dev(), engine=engine)
        x0 = torch.randn(4096, 196, device=_dev(), generator=torch.Generator(_dev()).manual_seed(5))
        spec = extract_spec(b["loss"], "exp_integrator", b["ts"], b["terminal"], b["second"], train=True, compute_ito=True)
        outs[engine] = eng.rollout(spec, x0, seed=77, engine=engine)
        again = eng.rollout(spec, x0, seed=77, engine=engine)
        assert torch.equal(outs[engine][0], again[0]) and torch.equal(outs[engine][1], again[1])
        assert torch.isfinite(outs[engine][0]).all() and torch.isfinite(outs[engine][1]).all()
    assert_close(outs["tcgen05"][0].cpu().numpy(), outs["simt"][0].cpu().numpy(), 1e-4, 1e-4, "x_T engines")
    assert_close(outs["tcgen05"][1].cpu().numpy(), outs["simt"][1].cpu().numpy(), 1e-4, 1e-4, "rnd engines")
