"""-m gpu: every compiled state-dimension bucket of both engines (DPAD 4..64 / 8..64), odd step counts and
hidden-layer counts, against the numpy oracle on synthetic specs (random weights; the oracle itself is
pinned to the reference by tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import philox, rollout as oracle_rollout
from sdes_test_helpers import assert_close, build_from_spec

pytestmark = pytest.mark.gpu


def synth_spec(d, T, n_hidden, target, loss_kind, ctrl, seed):
    rng = np.random.default_rng(seed)
    C = 64

    def lin(o, i, s=None):
        s = s if s is not None else 1.0 / np.sqrt(i)
        return (rng.uniform(-s, s, (o, i)).astype(np.float32), rng.uniform(-s, s, (o,)).astype(np.float32))

    def time_embed(n_hid, out, out_bias=0.0):
        hidden = [lin(C, 2 * C)] + [lin(C, C) for _ in range(n_hid - 1)]
        ow, ob = lin(out, C, 0.05)
        return {"phase": rng.standard_normal(C).astype(np.float32), "hidden": hidden, "out_w": ow, "out_b": ob + np.float32(out_bias)}

    in_w, in_b = lin(C, d)
    out_w, out_b = lin(d, C, 0.15)
    mlp = {"in_w": in_w, "in_b": in_b, "hidden": [lin(C, C) for _ in range(n_hidden)], "out_w": out_w, "out_b": out_b,
           "time_embed": time_embed(1, C)}
    spec = {"dim": d, "mlp": mlp, "gate": None if ctrl == "clipped" else time_embed(3, 1, 1.0)}
    if target == "gmm":
        K = 7
        tg = {"kind": "gmm", "loc": (rng.uniform(-3, 3, (K, d))).astype(np.float32),
              "scale": rng.uniform(0.6, 1.5, (K, d)).astype(np.float32),
              "log_weights": np.log(np.full(K, 1.0 / K)).astype(np.float32), "log_norm_const": 0.0}
    elif target == "gmm_shared":  # only the first 3 dims differ between components
        K = 6
        loc = np.tile(rng.uniform(-1, 1, (1, d)), (K, 1))
        loc[:, :3] = rng.uniform(-4, 4, (K, 3))
        sc = np.tile(rng.uniform(0.7, 1.4, (1, d)), (K, 1))
        tg = {"kind": "gmm", "loc": loc.astype(np.float32), "scale": sc.astype(np.float32),
              "log_weights": np.log(np.full(K, 1.0 / K)).astype(np.float32), "log_norm_const": 0.0}
    elif target == "funnel":
        tg = {"kind": "funnel", "variance": 4.0, "log_norm_const": 0.0}
    else:
        tg = {"kind": "multiwell", "n_dw": max(1, d // 3), "separation": 1.5, "shift": 0.25}
    tg["clip_target"] = None
    spec["target"] = tg
    gauss = {"loc": np.zeros(d, np.float32), "scale": np.ones(d, np.float32)}
    spec["prior"] = gauss
    spec["ref"] = None if loss_kind == "time_reversal" else {"loc": np.zeros(d, np.float32), "scale": np.full(d, 1.2, np.float32)}
    spec["ctrl"] = {"kind": ctrl, "clip_model": 10.0}
    if ctrl != "clipped":
        spec["ctrl"].update(scale_score=1.0, clip_score=10.0)
    if loss_kind == "exp_integrator":
        spec["sde"] = None
        ts = np.cumsum(np.concatenate([[0.0], rng.uniform(0.02, 0.06, T)])).astype(np.float32)
    else:
        # without a score term nothing pulls x back: keep the clipped-control case mild
        spec["sde"] = {"kind": "vp", "beta_min": 0.1, "beta_max": 2.0 if ctrl == "clipped" else 8.0, "scale": 1.0,
                       "terminal_t": 1.0, "sign": 1.0}
        ts = np.linspace(0, 1, T + 1).astype(np.float32)
    spec["ts"] = ts
    spec["loss"] = {"kind": loss_kind, "method": "lv", "train": True, "compute_ito": True, "return_traj": False,
                    "max_rnd": 1e8, "traj_per_sample": 1}
    if loss_kind == "exp_integrator":
        spec["loss"].update(alpha=1.0, sigma=1.0)
    if loss_kind == "reference_sde":
        spec["loss"]["reference_ctrl"] = False
    return spec


CASES = [
    # d, T, n_hidden, target, loss, ctrl
    (64, 23, 2, "gmm", "time_reversal", "lerp"),
    (57, 11, 2, "gmm_shared", "time_reversal", "lerp"),
    (40, 17, 2, "gmm", "time_reversal", "lerp_target"),
    (33, 30, 1, "multiwell", "time_reversal", "lerp"),  # (T=9 is ill-conditioned: the fp32 and fp64 oracle runs differ by 8e-4)
    (20, 31, 3, "gmm_shared", "reference_sde", "score"),
    (16, 8, 0, "funnel", "exp_integrator", "score"),
    (12, 13, 2, "gmm", "time_reversal", "clipped"),
    (8, 1, 2, "gmm", "time_reversal", "lerp"),
    (7, 40, 4, "funnel", "reference_sde", "score"),
    (49, 19, 2, "funnel", "exp_integrator", "score"),
]


@pytest.mark.parametrize("engine", ["tcgen05", "simt"])
@pytest.mark.parametrize("d,T,nh,target,loss_kind,ctrl", CASES)
def test_dimension_buckets_match_oracle(d, T, nh, target, loss_kind, ctrl, engine):
    spec = synth_spec(d, T, nh, target, loss_kind, ctrl, seed=d * 1000 + T)
    B = 301
    rng = np.random.default_rng(7)
    x0 = rng.standard_normal((B, d)).astype(np.float32)
    noise = philox.normal_noise(99, B, T, d)
    want_x, want_r, _ = oracle_rollout.rollout(spec, x0, noise=noise)
    dev = torch.device("cuda:0")
    b = build_from_spec(spec, dev, engine=engine)
    kw = {"terminal_unnorm_log_prob": b["terminal"], b["second_name"]: b["second"]}
    x_T, rnd, _ = b["loss"].simulate(b["ts"], torch.from_numpy(x0).to(dev), compute_ito_int=True,
                                     noise=torch.from_numpy(noise).to(dev), **kw)
    assert_close(x_T.cpu().numpy(), want_x, 2e-4, 2e-4, "x_T")
    assert_close(rnd.cpu().numpy(), want_r, 2e-4, 2e-4, "rnd")
