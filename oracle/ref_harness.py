"""TEST INFRASTRUCTURE — builds the *unmodified reference* objects for the rollout path.

Only usable where `/root/reference` exists (the build container).  Used by
`oracle/gen_golden.py` to freeze golden vectors under `tests/golden/` and by the
`not gpu` tests that pin the numpy oracle against the live reference.  Nothing in
the product (`sde_sampler_b200/`) imports this module.

Recipe: SURVEY.md Appendix B (stub the absent third-party imports, build the
objects with the YAML values directly, inject noise through `torch.randn_like`).
"""
from __future__ import annotations

import contextlib
import os
import sys
from functools import partial
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SDES_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "torchquad", "torchsde", "plotly", "plotly.graph_objects", "plotly.express",
    "matplotlib", "matplotlib.pyplot", "pykeops", "pykeops.torch", "hydra",
    "hydra.utils", "omegaconf", "torch_ema",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sde_sampler"))


def import_reference():
    """Put the reference on sys.path with MagicMock stubs for absent deps."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for m in _STUBS:
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sde_sampler  # noqa: F401

    return sde_sampler


@contextlib.contextmanager
def injected_noise(noise):
    """Replace torch.randn_like by a pop from `noise` (T,B,d) — exactly one call per
    step in every simulate() (reference losses/oc.py:214, :326, :432)."""
    import torch

    it = iter(noise)
    orig = torch.randn_like

    def fake(x, *a, **k):
        n = next(it)
        assert n.shape == x.shape, (n.shape, x.shape)
        return n.to(x.dtype)

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig


def fab_loc(dim: int):
    """GMM-40 'fab' locations (reference distr/gauss.py:42-47), zero-padded to `dim`
    (the reference's own dim>2 padding is broken for 40 modes)."""
    import torch

    g = torch.Generator()
    g.manual_seed(42)
    loc2 = (torch.rand((40, 2), generator=g) - 0.5) * 2 * 40
    if dim == 2:
        return loc2
    if dim == 1:
        return loc2[:, :1].clone()
    return torch.cat([loc2, torch.zeros(40, dim - 2)], dim=1)


def build_target(kind: str, dim: int):
    import torch

    import_reference()
    from sde_sampler.distr.double_well import DoubleWell, MultiWell
    from sde_sampler.distr.funnel import Funnel
    from sde_sampler.distr.gauss import GMM, IsotropicGauss

    if kind == "gmm40":
        loc = fab_loc(dim)
        scale = torch.nn.functional.softplus(torch.tensor(1.0)) * torch.ones_like(loc)
        return GMM(dim=dim, loc=loc, scale=scale, mixture_weights=torch.ones(40),
                   name=None, domain_tol=None, n_reference_samples=1000)
    if kind == "gmm_rand":  # heterogeneous loc/scale/weights: exercises the general form
        g = torch.Generator()
        g.manual_seed(7)
        K = 5
        loc = (torch.rand((K, dim), generator=g) - 0.5) * 6
        scale = 0.5 + torch.rand((K, dim), generator=g)
        w = 0.2 + torch.rand((K,), generator=g)
        return GMM(dim=dim, loc=loc, scale=scale, mixture_weights=w, name=None,
                   domain_tol=None, n_reference_samples=1000)
    if kind.startswith("gmm_dense"):
        # "gmm_dense:<K>" — K modes that differ in EVERY dimension (random loc, per-(k, j) scale, non-uniform weights): the
        # general MixtureSameFamily of distr/gauss.py:119-140 with no dimension shared between components
        g = torch.Generator()
        K = int(kind.split(":")[1])
        g.manual_seed(1000 + K)
        loc = (torch.rand((K, dim), generator=g) - 0.5) * (80.0 if K >= 40 else 8.0)
        scale = 0.8 + 1.2 * torch.rand((K, dim), generator=g)
        w = 0.5 + torch.rand((K,), generator=g)
        return GMM(dim=dim, loc=loc, scale=scale, mixture_weights=w, name=None,
                   domain_tol=None, n_reference_samples=1000)
    if kind == "dw_shift":
        assert dim == 1
        return DoubleWell(separation=2.0, shift=1.5)
    if kind == "multiwell":
        return MultiWell(dim=dim, n_double_wells=max(1, dim // 2), separation=2.0, shift=0.5)
    if kind == "funnel":
        return Funnel(dim=dim, n_reference_samples=1000)
    if kind == "gauss":
        return IsotropicGauss(dim=dim, loc=0.7, scale=1.3)
    if kind.startswith("nice"):
        # "nice:<mid_dim>:<hidden>" — the reference's NiceModel (distr/nice.py:127-216) with seeded random
        # weights wrapped exactly like Nice.unnorm_log_prob (:262-263); the shipped Nice class needs
        # data/nice.pt (d = 196 only), which does not exist.
        from sde_sampler.distr.base import Distribution
        from sde_sampler.distr.nice import NiceModel, StandardLogistic

        _, mid, hidden = kind.split(":")
        torch.manual_seed(4242)
        model = NiceModel(prior=StandardLogistic(), coupling=4, in_out_dim=dim, mid_dim=int(mid),
                          hidden=int(hidden), mask_config=1)
        with torch.no_grad():
            model.scaling.scale.normal_(0.0, 0.2)
        model.eval()
        for p in model.parameters():
            p.requires_grad_(False)

        class Nice(Distribution):
            def __init__(self):
                super().__init__(dim=dim, log_norm_const=0.0, n_reference_samples=1000)
                self.model = model

            def unnorm_log_prob(self, x):
                return self.model.log_prob(x).unsqueeze(-1) + self.log_norm_const

        return Nice()
    raise ValueError(kind)


def build_base_models(dim: int, seed: int, gate_bias: float, gate_dim: int = 1):
    """FourierMLP + TimeEmbed gate as in conf/model/base/{fouriermlp,time_embed}.yaml,
    with the zero-initialised out layers RE-RANDOMISED (default zero init
    makes NN == 0 and parity vacuous)."""
    import torch
    from torch import nn

    import_reference()
    from sde_sampler.models.mlp import FourierMLP, TimeEmbed

    torch.manual_seed(seed)
    base = FourierMLP(dim=dim, activation=nn.GELU(), num_layers=4, channels=64)
    gate = TimeEmbed(dim_out=gate_dim, activation=nn.GELU(), num_layers=4, channels=64,
                     last_bias_init=partial(nn.init.constant_, val=gate_bias))
    with torch.no_grad():
        base.out_layer.weight.normal_(0.0, 0.15)
        base.out_layer.bias.normal_(0.0, 0.1)
        gate.out_layer.weight.normal_(0.0, 0.05)
    return base, gate


def build_case(case: dict):
    """Build (loss, ts, callables) of the unmodified reference for a case dict
    (see oracle/cases.py)."""
    import torch

    import_reference()
    from sde_sampler.distr.delta import Delta
    from sde_sampler.distr.gauss import IsotropicGauss
    from sde_sampler.eq.sdes import VP, ConstOU, ScaledBM
    from sde_sampler.losses.oc import (ExponentialIntegratorSDELoss, ReferenceSDELoss,
                                       TimeReversalLoss)
    from sde_sampler.models.reparam import (ClippedCtrl, LerpCtrl, LerpPriorCtrl,
                                            LerpTargetCtrl, ScoreCtrl)
    from sde_sampler.utils.common import get_timesteps

    d = case["dim"]
    target = build_target(case["target"], d)
    base, gate = build_base_models(d, case.get("seed", 1), case.get("gate_bias", 1.0),
                                   case.get("gate_dim", 1))

    sde_kind = case.get("sde")
    if sde_kind == "vp":
        sde = VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=case.get("beta_max", 10.0),
                 scale_diff_coeff=1.0, terminal_t=1.0)
    elif sde_kind == "bm_pis":
        sde = ScaledBM(diff_coeff=0.4472135954999579, terminal_t=5.0)
    elif sde_kind == "const_ou":
        sde = ConstOU(drift_coeff=4.5, diff_coeff=3.0, terminal_t=1.0)
    elif sde_kind is None:
        sde = None
    else:
        raise ValueError(sde_kind)

    prior_kind = case["prior"]
    if prior_kind == "gauss":
        prior = IsotropicGauss(dim=d, scale=case.get("prior_scale", 1.0))
    elif prior_kind == "delta":
        prior = Delta(dim=d)
    else:
        raise ValueError(prior_kind)

    ctrl_kind = case["ctrl"]
    kw = dict(base_model=base, clip_model=case.get("clip_model"),)
    skw = dict(target_score=target.score, score_model=gate, detach_score=False,
               scale_score=case.get("scale_score", 1.0), clip_score=case.get("clip_score"))
    if ctrl_kind == "clipped":
        ctrl = ClippedCtrl(**kw)
    elif ctrl_kind == "score":
        ctrl = ScoreCtrl(**kw, **skw)
    elif ctrl_kind in ("lerp", "lerp_prior", "lerp_target"):
        cls = {"lerp": LerpCtrl, "lerp_prior": LerpPriorCtrl, "lerp_target": LerpTargetCtrl}[ctrl_kind]
        ctrl = cls(**kw, **skw, sde=sde, prior_score=prior.score)
    else:
        raise ValueError(ctrl_kind)

    loss_kind = case["loss"]
    lkw = dict(generative_ctrl=ctrl, sde=sde, method=case["method"],
               max_rnd=case.get("max_rnd"), traj_per_sample=case.get("traj_per_sample", 1))
    clip_target = case.get("clip_target")

    def terminal(x):  # == TrainableDiff.clipped_target_unnorm_log_prob (solver/oc.py:48-54)
        out = target.unnorm_log_prob(x)
        if clip_target is not None:
            out = out.clip(min=-clip_target, max=clip_target)
        return out

    if loss_kind == "time_reversal":
        loss = TimeReversalLoss(**lkw)
        second = prior.log_prob
    elif loss_kind == "reference_sde":
        ref_ctrl = None
        if case.get("euler_dds"):
            ref_ctrl = lambda t, x: sde.diff(t, x) * prior.score(x)  # solver/oc.py:305-306
            ref_distr = sde.marginal_distr(sde.terminal_t, x_init=prior.loc,
                                           var_init=prior.scale ** 2)
        else:
            ref_distr = sde.marginal_distr(t=sde.terminal_t, x_init=prior.loc)
        loss = ReferenceSDELoss(**lkw, reference_ctrl=ref_ctrl)
        second = ref_distr.log_prob
    elif loss_kind == "exp_integrator":
        loss = ExponentialIntegratorSDELoss(**lkw, alpha=case.get("alpha", 1.0),
                                            sigma=case.get("sigma", 1.0))
        second = prior.log_prob
    else:
        raise ValueError(loss_kind)

    tk = case["timesteps"]
    if tk.get("rescale_t") == "cosine":
        ts = get_timesteps(0.0, tk["end"], dt=tk["dt"], rescale_t="cosine")
    else:
        end = sde.terminal_t if sde is not None else tk["end"]
        ts = get_timesteps(0.0, end, steps=tk["steps"])
    return dict(loss=loss, ts=ts, terminal=terminal, second=second, target=target,
                prior=prior, sde=sde, ctrl=ctrl, base=base, gate=gate)
