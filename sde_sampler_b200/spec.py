"""Host-side introspection of the reference's plugin objects -> a flat `RolloutSpec`.

The reference hands its losses *objects and bound methods* (SURVEY §8b "Callables are
bound methods"): `generative_ctrl` (ClippedCtrl / ScoreCtrl / Lerp*Ctrl wrapping a
FourierMLP + TimeEmbed gate, reference models/reparam.py, models/mlp.py), `sde`
(VP / ConstOU / ScaledBM, eq/sdes.py), `solver.clipped_target_unnorm_log_prob`,
`prior.log_prob`, `reference_distr.log_prob` (solver/oc.py:158-163, :213-215, :257-259).

`extract_spec` recovers the raw parameters from those objects by duck typing on class
name + attribute names, so that it accepts both the reference's classes and the
parameter-holder mirrors in `sde_sampler_b200.plugins`.  Anything it cannot express in
the kernel's descriptor raises `NotImplementedError` — there is no fallback path.
"""
from __future__ import annotations

import math
import weakref
from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np
import torch

CTRL_KINDS = {"ClippedCtrl": "clipped", "ScoreCtrl": "score", "LerpCtrl": "lerp",
              "LerpPriorCtrl": "lerp_prior", "LerpTargetCtrl": "lerp_target"}
LOSS_KINDS = ("time_reversal", "reference_sde", "exp_integrator")
METHODS = ("kl", "kl_ito", "lv", "lv_traj")


def _np(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        return t.detach().to("cpu", torch.float32).numpy().copy()
    return np.asarray(t, dtype=np.float32)


def _t(t) -> torch.Tensor:
    """Parameters stay torch tensors on their device (no copy, no host sync); the kernel
    reads them through their data pointers, the oracle through `RolloutSpec.to_dict()`."""
    if isinstance(t, torch.Tensor):
        return t.detach()
    return torch.as_tensor(t, dtype=torch.float32)


# 0-dim *buffers* (SDE coefficients, DoubleWell.shift, ...) are constants of the reference
# objects; reading one is a device->host sync, so remember (tensor identity, version) -> value.
_SCALAR_CACHE: dict[int, tuple] = {}
_CHECKED_MULTIWELL: "weakref.WeakSet" = weakref.WeakSet()


def _f(v) -> float:
    if isinstance(v, torch.Tensor):
        key = id(v)
        hit = _SCALAR_CACHE.get(key)
        if hit is not None and hit[0]() is v and hit[1] == v._version:
            return hit[2]
        val = float(v.detach().cpu())
        if len(_SCALAR_CACHE) > 4096:
            _SCALAR_CACHE.clear()
        _SCALAR_CACHE[key] = (weakref.ref(v), v._version, val)
        return val
    return float(v)


def _owner(fn: Callable, what: str):
    owner = getattr(fn, "__self__", None)
    if owner is None:
        raise NotImplementedError(
            f"{what} must be a bound method of a supported object (got {fn!r}); the fused "
            "rollout evaluates densities in-kernel and cannot call arbitrary Python.")
    return owner


def _cls(obj) -> str:
    return type(obj).__name__


# ----------------------------------------------------------------------------- networks
def _check_gelu(act, where):
    if _cls(act) != "GELU" or getattr(act, "approximate", "none") != "none":
        raise NotImplementedError(f"{where}: only exact-erf nn.GELU() is supported (got {act!r})")


def _time_embed_params(te) -> dict:
    """TimeEmbed (reference models/mlp.py:43-82)."""
    _check_gelu(te.activation, "TimeEmbed")
    if int(te.channels) != 64:
        raise NotImplementedError(f"TimeEmbed with channels={te.channels} (the fused kernel is built for 64)")
    hidden = [(_t(l.weight), _t(l.bias)) for l in te.hidden_layer]
    return {"phase": _t(te.timestep_phase).reshape(-1), "hidden": hidden,
            "out_w": _t(te.out_layer.weight), "out_b": _t(te.out_layer.bias)}


def _fourier_mlp_params(net) -> dict:
    """FourierMLP (reference models/mlp.py:85-122)."""
    if _cls(net) != "FourierMLP":
        raise NotImplementedError(f"base_model {_cls(net)} is not supported (FourierMLP only)")
    _check_gelu(net.activation, "FourierMLP")
    if int(net.channels) != 64:
        raise NotImplementedError(f"FourierMLP with channels={net.channels} (the fused kernel is built for 64)")
    return {"in_w": _t(net.input_embed.weight), "in_b": _t(net.input_embed.bias),
            "hidden": [(_t(l.weight), _t(l.bias)) for l in net.hidden_layer],
            "out_w": _t(net.out_layer.weight), "out_b": _t(net.out_layer.bias),
            "time_embed": _time_embed_params(net.timestep_embed)}


def ctrl_parameters(ctrl) -> list:
    """The control's trainable tensors in the order of the parameter blob (include/sdes_b200.h) — the live
    nn.Parameters, not detached copies: inputs of the autograd node of the lv loss (sde_sampler_b200/autograd.py)."""
    net = ctrl.base_model
    te = net.timestep_embed
    ps = [net.input_embed.weight, net.input_embed.bias, te.timestep_phase]
    for l in te.hidden_layer:
        ps += [l.weight, l.bias]
    ps += [te.out_layer.weight, te.out_layer.bias]
    for l in net.hidden_layer:
        ps += [l.weight, l.bias]
    ps += [net.out_layer.weight, net.out_layer.bias]
    gate = getattr(ctrl, "score_model", None)
    if gate is not None:
        ps += [gate.timestep_phase]
        for l in gate.hidden_layer:
            ps += [l.weight, l.bias]
        ps += [gate.out_layer.weight, gate.out_layer.bias]
    return ps


# ------------------------------------------------------------------------- distributions
def _row(t, dim) -> torch.Tensor:
    """(1,d) / (d,) / scalar parameter -> contiguous (d,) float32 view (copy only if broadcast)."""
    t = _t(t).reshape(-1)
    if t.numel() == 1 and dim != 1:
        t = t.expand(dim)
    if t.numel() != dim:
        raise NotImplementedError(f"parameter of {t.numel()} elements for dim={dim}")
    return t.to(torch.float32).contiguous()


def _same_gauss(a: dict, b: dict) -> bool:
    if a.get("owner") is b.get("owner"):
        return True
    return bool(torch.equal(a["loc"], b["loc"]) and torch.equal(a["scale"], b["scale"]))


def _diag_gauss(distr, dim) -> dict:
    """Gauss / IsotropicGauss / Delta (reference distr/gauss.py:158-242, distr/delta.py)."""
    if _cls(distr) not in ("Gauss", "IsotropicGauss", "Delta"):
        raise NotImplementedError(f"{_cls(distr)} is not a supported Gaussian prior/reference")
    lnc = getattr(distr, "log_norm_const", 0.0) or 0.0
    if abs(float(lnc)) > 0:
        raise NotImplementedError("Gaussian prior/reference with log_norm_const != 0")
    return {"loc": _row(distr.loc, dim), "scale": _row(distr.scale, dim), "owner": distr}


def _target_params(target, dim) -> dict:
    name = _cls(target)
    lnc = getattr(target, "log_norm_const", None)
    if name == "GMM":
        # distr/gauss.py:66-140
        loc, scale = _t(target.loc), _t(target.scale)
        if loc.ndim != 2 or loc.shape != scale.shape or loc.shape[1] != dim:
            raise NotImplementedError(f"GMM loc/scale of shape {tuple(loc.shape)}/{tuple(scale.shape)}")
        # Categorical(probs=w) normalises (distr/gauss.py:119-128); the kernel does the same
        w = None if target.mixture_weights is None else _t(target.mixture_weights).reshape(-1)
        return {"kind": "gmm", "loc": loc.to(torch.float32).contiguous(),
                "scale": scale.to(torch.float32).contiguous(),
                "weights": None if w is None else w.to(torch.float32).contiguous(),
                "log_norm_const": float(lnc or 0.0)}
    if name in ("Gauss", "IsotropicGauss"):
        return {"kind": "gauss", "loc": _row(target.loc, dim), "scale": _row(target.scale, dim),
                "log_norm_const": float(lnc or 0.0)}
    if name == "DoubleWell":
        # distr/double_well.py:14-45 — unnorm_log_prob has no constant
        return {"kind": "multiwell", "n_dw": 1, "separation": _f(target.separation),
                "shift": _f(target.shift)}
    if name == "MultiWell":
        # distr/double_well.py:103-179; Gaussian part: loc=shift, scale=1, constant folded to 0
        dw = target.double_well
        if target.gauss is not None and target not in _CHECKED_MULTIWELL:
            gs = _np(target.gauss.scale).reshape(-1)  # host read: once per target object
            gl = _np(target.gauss.loc).reshape(-1)
            if not (np.allclose(gs, 1.0) and np.allclose(gl, _f(dw.shift))):
                raise NotImplementedError("MultiWell with a non-standard Gaussian part")
            _CHECKED_MULTIWELL.add(target)
        return {"kind": "multiwell", "n_dw": int(target.n_double_wells),
                "separation": _f(dw.separation), "shift": _f(dw.shift)}
    if name == "Funnel":
        # distr/funnel.py:11-80
        return {"kind": "funnel", "variance": float(target.variance),
                "log_norm_const": float(lnc or 0.0)}
    if name == "Nice" or (hasattr(target, "model") and _cls(target.model) == "NiceModel"):
        return _nice_params(target, dim, lnc)
    raise NotImplementedError(
        f"target {name} is not implemented in the fused rollout (supported: GMM, Gauss, "
        "IsotropicGauss, DoubleWell, MultiWell, Funnel, Nice)")


def _nice_params(target, dim, lnc) -> dict:
    """Nice / NiceModel (distr/nice.py:127-263): additive couplings with ReLU MLPs, a log-scale
    vector and a standard-logistic latent prior."""
    model = target.model
    if int(model.in_out_dim) != dim or dim % 2:
        raise NotImplementedError(f"NICE with in_out_dim={model.in_out_dim} for dim={dim}")
    if _cls(model.prior) != "StandardLogistic":
        raise NotImplementedError(f"NICE latent prior {_cls(model.prior)} (StandardLogistic only)")
    couplings = []
    for c in model.coupling:
        lins = [c.in_block[0]] + [b[0] for b in c.mid_block] + [c.out_block]
        for blk in [c.in_block] + list(c.mid_block):
            if _cls(blk[1]) != "ReLU":
                raise NotImplementedError("NICE coupling with a non-ReLU activation")
        couplings.append({"mask_config": int(c.mask_config),
                          "layers": [(_t(l.weight), _t(l.bias)) for l in lins]})
    if len({len(c["layers"]) for c in couplings}) != 1 or len({tuple(c["layers"][0][0].shape) for c in couplings}) != 1:
        raise NotImplementedError("NICE couplings of different shapes")
    return {"kind": "nice", "couplings": couplings, "scale": _t(model.scaling.scale).reshape(-1),
            "log_norm_const": float(lnc or 0.0)}


def _sde_params(sde) -> dict | None:
    if sde is None:
        return None
    name = _cls(sde)
    if getattr(sde, "noise_type", "diagonal") not in ("diagonal", "scalar"):
        raise NotImplementedError(f"sde noise type {sde.noise_type}")
    if name == "VP":
        return {"kind": "vp", "beta_min": _f(sde.diff_coeff_sq_min), "beta_max": _f(sde.diff_coeff_sq_max),
                "scale": _f(sde.scale_diff_coeff), "terminal_t": _f(sde.terminal_t), "sign": float(sde.sign)}
    if name in ("ConstOU", "ScaledBM"):
        return {"kind": "const_ou", "drift_coeff": _f(sde.drift_coeff), "diff_coeff": _f(sde.diff_coeff),
                "terminal_t": _f(sde.terminal_t), "sign": float(sde.sign)}
    raise NotImplementedError(f"sde {name} is not supported (VP, ConstOU, ScaledBM)")


# ------------------------------------------------------------------------------ the spec
@dataclass
class RolloutSpec:
    """Raw, framework-free description of one rollout call (numpy arrays + scalars)."""
    dim: int
    ts: Any
    loss: dict
    ctrl: dict
    mlp: dict
    gate: dict | None
    sde: dict | None
    prior: dict | None
    ref: dict | None
    target: dict
    extras: dict = field(default_factory=dict)

    def to_dict(self) -> dict:
        """Plain numpy / scalar copy (what the oracle and the golden fixtures consume)."""
        def conv(o):
            if isinstance(o, torch.Tensor):
                return _np(o)
            if isinstance(o, dict):
                return {k: conv(v) for k, v in o.items() if k != "owner"}
            if isinstance(o, (list, tuple)):
                return type(o)(conv(v) for v in o)
            return o

        target = conv(self.target)
        if target["kind"] == "gmm":
            w = target.pop("weights")
            if w is None:
                target["log_weights"] = np.zeros((1,), np.float32)
            else:
                w = w.astype(np.float64)
                target["log_weights"] = np.log(w / w.sum()).astype(np.float32)
        return {"dim": self.dim, "ts": conv(self.ts), "loss": dict(self.loss), "ctrl": dict(self.ctrl),
                "mlp": conv(self.mlp), "gate": conv(self.gate), "sde": conv(self.sde),
                "prior": conv(self.prior), "ref": conv(self.ref), "target": target}


def extract_spec(loss_obj, loss_kind: str, ts: torch.Tensor, terminal_unnorm_log_prob: Callable,
                 second_log_prob: Callable | None, *, train: bool, compute_ito: bool,
                 return_traj: bool = False) -> RolloutSpec:
    """Build the spec for one `simulate` call of a fused loss.

    `second_log_prob` is `initial_log_prob` (TimeReversalLoss, losses/oc.py:160) or
    `reference_log_prob` (ReferenceSDELoss :291, ExponentialIntegratorSDELoss :405)."""
    assert loss_kind in LOSS_KINDS
    if loss_obj.method not in METHODS:
        raise ValueError("Unknown loss method.")
    if getattr(loss_obj, "inference_ctrl", None) is not None:
        raise NotImplementedError("learned inference_ctrl (Bridge) needs div_x autograd; out of scope (SURVEY §8a1)")
    if loss_obj.sde_ctrl_noise is not None or loss_obj.sde_ctrl_dropout is not None:
        raise NotImplementedError("sde_ctrl_noise / sde_ctrl_dropout are not implemented (SURVEY §8a4)")

    ctrl = loss_obj.generative_ctrl
    kind = CTRL_KINDS.get(_cls(ctrl))
    if kind is None:
        raise NotImplementedError(f"control {_cls(ctrl)} is not supported {tuple(CTRL_KINDS)}")
    if getattr(ctrl, "hard_constrain", False):
        raise NotImplementedError("hard_constrain=True")
    mlp = _fourier_mlp_params(ctrl.base_model)
    dim = int(mlp["in_w"].shape[1])
    if mlp["out_w"].shape[0] != dim:
        raise NotImplementedError("FourierMLP with dim_out != dim")

    cd: dict[str, Any] = {"kind": kind, "clip_model": ctrl.clip_model}
    gate = None
    target_obj = None
    prior = None
    sde = _sde_params(loss_obj.sde)
    if kind != "clipped":
        cd["scale_score"] = float(ctrl.scale_score)
        cd["clip_score"] = ctrl.clip_score
        if ctrl.score_model is not None:
            if _cls(ctrl.score_model) != "TimeEmbed":
                raise NotImplementedError("score_model must be a TimeEmbed (x-independent gate)")
            gate = _time_embed_params(ctrl.score_model)
            if gate["out_w"].shape[0] not in (1, dim):
                raise NotImplementedError("gate dim_out must be 1 or dim")
        target_obj = _owner(ctrl.target_score, "target_score")
        if kind in ("lerp", "lerp_prior", "lerp_target"):
            if sde is None:
                raise ValueError("Lerp controls need an sde")
            if _sde_params(ctrl.sde) != sde:
                raise NotImplementedError("control and loss use different SDEs")
            prior = _diag_gauss(_owner(ctrl.prior_score, "prior_score"), dim)

    # terminal density: solver.clipped_target_unnorm_log_prob (solver/oc.py:48-54) or target.unnorm_log_prob
    owner = _owner(terminal_unnorm_log_prob, "terminal_unnorm_log_prob")
    clip_target = None
    if hasattr(owner, "target") and hasattr(owner, "clip_target"):
        clip_target = owner.clip_target
        term_obj = owner.target
    else:
        term_obj = owner
    if target_obj is not None and term_obj is not target_obj:
        raise NotImplementedError("control's target_score and terminal_unnorm_log_prob refer to different targets")
    target = _target_params(term_obj, dim)
    target["clip_target"] = clip_target
    # filter_samples (target.filter, solver/oc.py:152) acts on x_T after the rollout: applied by the statistics kernel's
    # sample mask in compute_loss (losses/oc.py:50-58), nothing for the descriptor to carry

    ld: dict[str, Any] = {"kind": loss_kind, "method": loss_obj.method, "train": bool(train),
                          "compute_ito": bool(compute_ito), "return_traj": bool(return_traj),
                          "max_rnd": loss_obj.max_rnd, "traj_per_sample": int(loss_obj.traj_per_sample)}
    ref = None
    if loss_kind == "time_reversal":
        if sde is None:
            raise ValueError("TimeReversalLoss needs an sde")
        if not (train and loss_obj.method in ("kl", "kl_ito")):
            p = _diag_gauss(_owner(second_log_prob, "initial_log_prob"), dim)
            if prior is not None and not _same_gauss(p, prior):
                raise NotImplementedError("initial_log_prob and prior_score refer to different priors")
            prior = p
    else:
        ref = _diag_gauss(_owner(second_log_prob, "reference_log_prob"), dim)
        if loss_kind == "reference_sde":
            if sde is None:
                raise ValueError("ReferenceSDELoss needs an sde")
            rc = getattr(loss_obj, "reference_ctrl", None)
            ld["reference_ctrl"] = rc is not None
            if rc is not None:
                # EulerDDS.reference_ctrl = sde.diff * prior.score (solver/oc.py:305-306)
                rc_owner = _owner(rc, "reference_ctrl")
                if not hasattr(rc_owner, "prior"):
                    raise NotImplementedError("reference_ctrl must be EulerDDS-style (owner with .prior)")
                p = _diag_gauss(rc_owner.prior, dim)
                if prior is not None and not _same_gauss(p, prior):
                    raise NotImplementedError("reference_ctrl prior differs from control prior")
                prior = p
        else:
            ld["alpha"] = float(loss_obj.alpha)
            ld["sigma"] = float(loss_obj.sigma)

    ts_t = _t(ts).reshape(-1).to(torch.float32)
    if ts_t.shape[0] < 2:
        raise ValueError("need at least one time step")
    # only the kl gradient looks at this: with detach_score the score term is evaluated on x.detach() (reparam.py:58,:134)
    extras = {"detach_score": kind != "clipped" and bool(getattr(ctrl, "detach_score", False))}
    return RolloutSpec(dim=dim, ts=ts_t, loss=ld, ctrl=cd, mlp=mlp, gate=gate, sde=sde,
                       prior=prior, ref=ref, target=target, extras=extras)


def inf_if_none(v) -> float:
    return math.inf if v is None else float(v)
