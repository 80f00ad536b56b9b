"""Parameter-holder mirrors of the reference plug-ins the rollout path touches.

The fused losses accept the reference's own objects (`sde_sampler.models.reparam.LerpCtrl`,
`sde_sampler.eq.sdes.VP`, `sde_sampler.distr.gauss.GMM`, ...) and introspect them by class name
and attribute name (sde_sampler_b200/spec.py).  Where the reference package is not installed
(the GPU box, bench.py, the -m gpu tests) these mirrors carry the same names, constructor
arguments, parameter shapes and initialisation, so the same introspection applies.

They hold parameters only.  Their per-step arithmetic (`forward`, `score`, `log_prob`, `drift`,
...) is evaluated inside the CUDA kernel; calling it on a mirror raises — there is no
PyTorch implementation of the path in the product.  What stays here is what the caller runs
OUTSIDE the rollout: `prior.sample` (solver/oc.py:71) and `get_timesteps` (utils/common.py:18-55).
"""
from __future__ import annotations

import math
from typing import Callable

import torch
from torch import nn


def _in_kernel(name: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__}.{name} is evaluated inside the fused CUDA rollout; this mirror only "
            "holds its parameters (pass the bound method to a Fused*Loss, do not call it)")
    method.__name__ = name
    return method


# ------------------------------------------------------------------------------ networks
class TimeEmbed(nn.Module):
    """models/mlp.py:43-82"""

    def __init__(self, dim_out: int, activation: Callable | None = None, num_layers: int = 2, channels: int = 64,
                 last_bias_init: Callable | None = None, last_weight_init: Callable | None = None):
        super().__init__()
        self.dim, self.dim_out, self.channels = 1, dim_out, channels
        self.activation = activation if activation is not None else nn.GELU()
        self.register_buffer("timestep_coeff", torch.linspace(start=0.1, end=100, steps=channels).unsqueeze(0),
                             persistent=False)
        self.timestep_phase = nn.Parameter(torch.randn(1, channels))
        self.hidden_layer = nn.ModuleList([nn.Linear(2 * channels, channels)])
        self.hidden_layer += [nn.Linear(channels, channels) for _ in range(num_layers - 2)]
        self.out_layer = nn.Linear(channels, dim_out)
        if last_bias_init:
            last_bias_init(self.out_layer.bias)
        if last_weight_init:
            last_weight_init(self.out_layer.weight)

    forward = _in_kernel("forward")


class FourierMLP(nn.Module):
    """models/mlp.py:85-122"""

    def __init__(self, dim: int, activation: Callable | None = None, num_layers: int = 4, channels: int = 64,
                 last_bias_init: Callable | None = None, last_weight_init: Callable | None = None, dim_out=None):
        super().__init__()
        self.dim, self.dim_out, self.channels = dim, dim_out or dim, channels
        self.activation = activation if activation is not None else nn.GELU()
        self.input_embed = nn.Linear(dim, channels)
        self.timestep_embed = TimeEmbed(dim_out=channels, activation=self.activation, num_layers=2, channels=channels)
        self.hidden_layer = nn.ModuleList([nn.Linear(channels, channels) for _ in range(num_layers - 2)])
        self.out_layer = nn.Linear(channels, self.dim_out)
        if last_bias_init:
            last_bias_init(self.out_layer.bias)
        if last_weight_init:
            last_weight_init(self.out_layer.weight)

    forward = _in_kernel("forward")


# ------------------------------------------------------------------------------ controls
class ClippedCtrl(nn.Module):
    """models/reparam.py:13-36"""

    def __init__(self, base_model: nn.Module, clip_model: float | None = None, name: str = "ctrl", **kwargs):
        super().__init__()
        self.base_model = base_model
        self.clip_model = clip_model
        self.name = name

    forward = _in_kernel("forward")


class ScoreCtrl(ClippedCtrl):
    """models/reparam.py:39-83"""

    def __init__(self, *args, target_score: Callable, score_model: nn.Module | None = None, detach_score: bool = True,
                 scale_score: float = 1.0, clip_score: float | None = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.score_model = score_model
        self.target_score = target_score
        self.detach_score = detach_score
        self.scale_score = scale_score
        self.clip_score = clip_score


class LerpCtrl(ScoreCtrl):
    """models/reparam.py:113-162"""

    def __init__(self, *args, sde, prior_score: Callable, hard_constrain: bool = False, scale_lerp: bool = False,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.sde = sde
        self.prior_score = prior_score
        self.hard_constrain = hard_constrain
        self.scale_lerp = scale_lerp


class LerpPriorCtrl(LerpCtrl):
    """models/reparam.py:165-181"""


class LerpTargetCtrl(LerpCtrl):
    """models/reparam.py:184-200"""


# ---------------------------------------------------------------------------------- SDEs
class _OU(nn.Module):
    """eq/sdes.py:68-122 (+ TorchSDE base :14-35): terminal_t, sign, noise_type."""

    noise_type = "diagonal"

    def __init__(self, terminal_t: float = 1.0, generative: bool = True, **kwargs):
        super().__init__()
        self.register_buffer("terminal_t", torch.tensor(terminal_t, dtype=torch.float), persistent=False)
        self.generative = generative
        self.sign = 1.0 if generative else -1.0

    drift = _in_kernel("drift")
    diff = _in_kernel("diff")
    drift_div_int = _in_kernel("drift_div_int")
    # The x-INDEPENDENT coefficient functions stay callable: FusedEulerIntegrator evaluates them once on the whole grid
    # (they are the caller's object's own functions there; these mirror eq/sdes.py for the tests on the GPU box).


class ControlledSDE(nn.Module):
    """eq/sdes.py:272-305 — holder of (sde, ctrl); drift = sde.drift + sde.diff * ctrl(t or terminal_t - t, x)."""

    def __init__(self, sde, ctrl=None):
        super().__init__()
        self.sde, self.ctrl = sde, ctrl
        self.register_buffer("terminal_t", sde.terminal_t.clone(), persistent=False)

    drift = _in_kernel("drift")
    diff = _in_kernel("diff")
    f_and_g = _in_kernel("f_and_g")


class ConstOU(_OU):
    """eq/sdes.py:125-172"""

    def __init__(self, drift_coeff: float = 2.0, diff_coeff: float = 2.0, **kwargs):
        if drift_coeff < 0 or diff_coeff <= 0:
            raise ValueError("Choose non-negative drift_coeff and positive diff_coeff.")
        super().__init__(**kwargs)
        self.register_buffer("drift_coeff", torch.tensor(drift_coeff, dtype=torch.float), persistent=False)
        self.register_buffer("diff_coeff", torch.tensor(diff_coeff, dtype=torch.float), persistent=False)

    def drift_coeff_t(self, t):  # eq/sdes.py:141-142
        return self.sign * self.drift_coeff

    def diff_coeff_t(self, t):  # :144-145
        return self.diff_coeff

    def marginal_params(self, t, x_init, var_init=None):  # :157-172
        drift_coeff = self.sign * self.drift_coeff
        loc = torch.exp(drift_coeff * t)
        var = -self.diff_coeff ** 2 / (2 * drift_coeff) * (1 - torch.exp(2 * drift_coeff * t))
        if var_init is not None:
            var = var + loc ** 2 * var_init
        return loc * x_init, var


class ScaledBM(ConstOU):
    """eq/sdes.py:175-188"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, drift_coeff=0.0, **kwargs)

    def marginal_params(self, t, x_init, var_init=None):  # eq/sdes.py:179-188
        var = self.diff_coeff ** 2 * t
        if var_init is not None:
            var = var + var_init
        return x_init, var

    def marginal_distr(self, t, x_init, var_init=None) -> "Gauss":
        var = self.diff_coeff ** 2 * t
        if var_init is not None:
            var = var + var_init
        return Gauss(dim=x_init.shape[-1], loc=x_init, scale=(var * torch.ones_like(x_init)).sqrt())


class VP(_OU):
    """eq/sdes.py:191-269"""

    def __init__(self, diff_coeff_sq_min: float = 0.1, diff_coeff_sq_max: float = 20.0, scale_diff_coeff: float = 1.0,
                 **kwargs):
        super().__init__(**kwargs)
        self.register_buffer("scale_diff_coeff", torch.tensor(scale_diff_coeff, dtype=torch.float), persistent=False)
        self.register_buffer("diff_coeff_sq_min", torch.tensor(diff_coeff_sq_min, dtype=torch.float), persistent=False)
        self.register_buffer("diff_coeff_sq_max", torch.tensor(diff_coeff_sq_max, dtype=torch.float), persistent=False)

    def _diff_coeff_sq_t(self, t):  # eq/sdes.py:222-229
        if self.generative:
            return torch.lerp(self.diff_coeff_sq_max, self.diff_coeff_sq_min, t / self.terminal_t)
        return torch.lerp(self.diff_coeff_sq_min, self.diff_coeff_sq_max, t / self.terminal_t)

    def drift_coeff_t(self, t):  # :231-232
        return self.sign * 0.5 * self._diff_coeff_sq_t(t)

    def diff_coeff_t(self, t):  # :234-235
        return self.scale_diff_coeff * torch.sqrt(self._diff_coeff_sq_t(t))


class LangevinSDE(nn.Module):
    """eq/sdes.py:38-65 — drift = clip(target_score(x) * diff_coeff^2 / 2, clip_score), constant diffusion."""

    def __init__(self, target_score: Callable, diff_coeff: float = 1.0, clip_score: float | None = None, terminal_t: float = 1.0):
        super().__init__()
        self.target_score = target_score
        self.clip_score = clip_score
        self.register_buffer("diff_coeff", torch.tensor(diff_coeff, dtype=torch.float), persistent=False)
        self.register_buffer("terminal_t", torch.tensor(terminal_t, dtype=torch.float), persistent=False)

    drift = _in_kernel("drift")
    diff = _in_kernel("diff")


# ------------------------------------------------------------------------- distributions
class _Distribution(nn.Module):
    def __init__(self, dim: int, log_norm_const: float | None = None):
        super().__init__()
        self.dim = dim
        self.log_norm_const = log_norm_const

    unnorm_log_prob = _in_kernel("unnorm_log_prob")
    log_prob = _in_kernel("log_prob")
    score = _in_kernel("score")


class GMM(_Distribution):
    """distr/gauss.py:66-155 (explicit loc / scale / mixture_weights; `name="fab"` builds the
    40-mode mixture of distr/gauss.py:42-47 for dim=2)."""

    def __init__(self, dim: int = 2, loc: torch.Tensor | None = None, scale: torch.Tensor | None = None,
                 mixture_weights: torch.Tensor | None = None, name: str | None = None,
                 log_norm_const: float = 0.0, **kwargs):
        super().__init__(dim=dim, log_norm_const=log_norm_const)
        if name is not None:
            if name != "fab":
                raise NotImplementedError(f"GMM name {name!r}")
            loc, scale, mixture_weights = fab_gmm_params(dim)
        n = loc.shape[0]
        if not loc.shape == scale.shape == (n, dim):
            raise ValueError("Shape missmatch between loc and scale.")
        if mixture_weights is None and n > 1:
            raise ValueError("Require mixture weights.")
        self.register_buffer("loc", loc.float(), persistent=False)
        self.register_buffer("scale", scale.float(), persistent=False)
        self.register_buffer("mixture_weights", None if mixture_weights is None else mixture_weights.float(),
                             persistent=False)


def fab_gmm_params(dim: int):
    """GMM-40 of the FAB paper as the reference builds it (distr/gauss.py:42-47): 40 modes,
    loc = (U[0,1)^{40x2} - 0.5) * 80 from torch.Generator seed 42, scale = softplus(1).  For dim > 2
    the locations are zero-padded (the reference's own dim>2 padding is broken for 40 modes)."""
    g = torch.Generator()
    g.manual_seed(42)
    loc = (torch.rand((40, 2), generator=g) - 0.5) * 2 * 40
    if dim == 1:
        loc = loc[:, :1].clone()
    elif dim > 2:
        loc = torch.cat([loc, torch.zeros(40, dim - 2)], dim=1)
    scale = torch.nn.functional.softplus(torch.tensor(1.0)) * torch.ones_like(loc)
    return loc, scale, torch.ones(40)


class Gauss(GMM):
    """distr/gauss.py:158-183"""

    def __init__(self, dim: int = 1, loc=0.0, scale=1.0, **kwargs):
        def prep(p):
            if not isinstance(p, torch.Tensor):
                p = torch.tensor(p, dtype=torch.float)
            p = torch.atleast_2d(p)
            if p.numel() == 1:
                p = p.repeat(1, dim)
            return p

        super().__init__(dim=dim, loc=prep(loc), scale=prep(scale), **kwargs)


class IsotropicGauss(Gauss):
    """distr/gauss.py:186-242"""

    def __init__(self, dim: int = 1, loc: float = 0.0, scale: float = 1.0, truncate_quartile: float | None = None,
                 **kwargs):
        super().__init__(dim=dim, loc=loc, scale=scale, **kwargs)
        if truncate_quartile is not None:
            q = torch.tensor([truncate_quartile / 2, 1 - truncate_quartile / 2])
            truncate_quartile = torch.distributions.Normal(float(loc), float(scale)).icdf(q).tolist()
        self.truncate_quartile = truncate_quartile

    def sample(self, shape: tuple | None = None) -> torch.Tensor:
        """Outside the rollout (solver/oc.py:71): x0 ~ prior, optionally truncated (gauss.py:235-242)."""
        shape = tuple() if shape is None else tuple(shape)
        dev = self.loc.device
        if self.truncate_quartile is None:
            return self.loc[0, 0] + self.scale[0, 0] * torch.randn(*shape, self.dim, device=dev)
        out = torch.empty(*shape, self.dim, device=dev)
        return nn.init.trunc_normal_(out, mean=float(self.loc[0, 0]), std=float(self.scale[0, 0]),
                                     a=self.truncate_quartile[0], b=self.truncate_quartile[1])


class Delta(Gauss):
    """distr/delta.py:8-28"""

    def __init__(self, dim: int = 1, loc=0.0, approx_scale: float = 1e-3, **kwargs):
        super().__init__(dim=dim, loc=loc, scale=approx_scale, **kwargs)

    def sample(self, shape: tuple | None = None) -> torch.Tensor:
        shape = tuple() if shape is None else tuple(shape)
        return self.loc.repeat(*shape, 1)


class DoubleWell(_Distribution):
    """distr/double_well.py:14-100"""

    def __init__(self, dim: int = 1, separation: float = 2.0, shift: float = 0.0, **kwargs):
        if not dim == 1:
            raise ValueError("`dim` needs to be `1`. Consider using `MultiWell`.")
        super().__init__(dim=1)
        self.register_buffer("separation", torch.tensor(separation), persistent=False)
        self.register_buffer("shift", torch.tensor(shift), persistent=False)


class MultiWell(_Distribution):
    """distr/double_well.py:103-193"""

    def __init__(self, dim: int = 2, n_double_wells: int = 1, separation: float = 2.0, shift: float = 0.0, **kwargs):
        super().__init__(dim=dim)
        if n_double_wells > dim or n_double_wells == 0:
            raise ValueError(f"Please specify between 1 and {dim} double wells.")
        self.separation = separation
        self.n_double_wells = n_double_wells
        self.n_gauss = dim - n_double_wells
        self.double_well = DoubleWell(separation=separation, shift=shift)
        self.gauss = None
        if self.n_gauss > 0:
            self.gauss = IsotropicGauss(dim=self.n_gauss, loc=shift,
                                        log_norm_const=0.5 * math.log(2.0 * math.pi) * self.n_gauss)


class Funnel(_Distribution):
    """distr/funnel.py:11-96"""

    def __init__(self, dim: int = 10, variance: float | None = None, log_norm_const: float = 0.0, **kwargs):
        super().__init__(dim=dim, log_norm_const=log_norm_const)
        self.variance = variance if variance is not None else dim - 1


class Coupling(nn.Module):
    """distr/nice.py:43-95 — additive coupling: `on <- on + out_block(relu-MLP(off))`."""

    def __init__(self, in_out_dim: int, mid_dim: int, hidden: int, mask_config: int):
        super().__init__()
        self.mask_config = mask_config
        self.in_block = nn.Sequential(nn.Linear(in_out_dim // 2, mid_dim), nn.ReLU())
        self.mid_block = nn.ModuleList([nn.Sequential(nn.Linear(mid_dim, mid_dim), nn.ReLU()) for _ in range(hidden - 1)])
        self.out_block = nn.Linear(mid_dim, in_out_dim // 2)

    forward = _in_kernel("forward")


class Scaling(nn.Module):
    """distr/nice.py:98-124"""

    def __init__(self, dim: int):
        super().__init__()
        self.scale = nn.Parameter(torch.zeros((1, dim)), requires_grad=True)

    forward = _in_kernel("forward")


class StandardLogistic:
    """distr/nice.py:17-40 (the latent prior; its log-density is evaluated in-kernel)."""


class NiceModel(nn.Module):
    """distr/nice.py:127-216"""

    def __init__(self, prior, coupling: int, in_out_dim: int, mid_dim: int, hidden: int, mask_config: int):
        super().__init__()
        self.prior = prior
        self.in_out_dim = in_out_dim
        self.coupling = nn.ModuleList([Coupling(in_out_dim, mid_dim, hidden, (mask_config + i) % 2) for i in range(coupling)])
        self.scaling = Scaling(in_out_dim)

    log_prob = _in_kernel("log_prob")
    forward = _in_kernel("forward")


class Nice(_Distribution):
    """distr/nice.py:219-263 — a NICE flow as target density.  The reference loads a checkpoint trained on
    14x14 MNIST (dim 196); no checkpoint can travel here, so the model is always passed in (any even dim)."""

    def __init__(self, model: nn.Module, dim: int | None = None, log_norm_const: float = 0.0, **kwargs):
        dim = int(model.in_out_dim) if dim is None else dim
        super().__init__(dim=dim, log_norm_const=log_norm_const)
        if dim != int(model.in_out_dim):
            raise ValueError(f"Dimension is {dim} but the model has {model.in_out_dim}.")
        self.model = model
        for p in self.model.parameters():
            p.requires_grad_(False)


# ---------------------------------------------------------------------------- time grids
def get_timesteps(start, end, dt=None, steps: int | None = None, rescale_t: str | None = None, device=None):
    """utils/common.py:18-55 — the (T+1,) fp32 grid the caller hands to the loss."""
    if (steps is None) is (dt is None):
        raise ValueError("Exactly one of `dt` and `steps` should be defined.")
    if steps is None:
        steps = int(math.ceil((end - start) / dt))
    if rescale_t is None:
        return torch.linspace(start, end, steps=steps + 1, device=device)
    if rescale_t == "quad":
        end_t = torch.as_tensor(end, dtype=torch.float)
        return torch.sqrt(torch.linspace(start, float(end_t.square()), steps=steps + 1, device=device)).clip(max=float(end))
    if rescale_t == "cosine":
        s = 0.008
        pre_phase = torch.linspace(start, end, steps + 1, device=device) / end
        phase = ((pre_phase + s) / (1 + s)) * torch.pi * 0.5
        dts = torch.cos(phase) ** 4
        dts /= dts.sum()
        dts *= end
        return torch.concat((torch.tensor([start], device=device), torch.cumsum(dts, -1)))
    raise ValueError("Unkown timestep rescaling method.")
