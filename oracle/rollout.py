"""TEST INFRASTRUCTURE — CPU (numpy) restatement of the reference rollout path.

This is the parity oracle for the B200 kernels.  It is NOT the product: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  Parity status: PINNED — `oracle/gen_golden.py` runs the unmodified
reference (`/root/reference/sde_sampler/losses/oc.py` `simulate`) with injected noise and
freezes its outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this file
against those vectors (and against the live reference where it is present).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Input is a *spec* dict of plain numpy arrays / scalars (see
`sde_sampler_b200/spec.py::RolloutSpec.to_dict`), i.e. raw model / SDE / target
parameters — all derived tables are recomputed here independently of the product.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf as _erf

from . import philox

F32 = np.float32


# --------------------------------------------------------------------------------------
# control network  (models/mlp.py)
# --------------------------------------------------------------------------------------
def gelu(x):
    """torch.nn.GELU() default = exact erf form (conf/model/base/fouriermlp.yaml:5-6)."""
    dt = x.dtype
    return (0.5 * x * (1.0 + _erf(x.astype(np.float64) / math.sqrt(2.0)).astype(dt))).astype(dt)


def linear(x, w, b):
    return x @ w.T + b


def time_embed(t, net, dtype):
    """TimeEmbed.forward, models/mlp.py:71-82.  t: (n,) -> (n, dim_out).
    timestep_coeff = linspace(0.1, 100, channels) (models/mlp.py:58)."""
    phase = net["phase"].astype(dtype).reshape(1, -1)
    C = phase.shape[1]
    # torch.linspace computes in fp32 for fp32 output; replicate its symmetric formula
    coeff = _torch_linspace(0.1, 100.0, C).astype(dtype).reshape(1, -1)
    t = np.asarray(t, dtype=dtype).reshape(-1, 1)
    arg = coeff * t + phase
    h = np.concatenate([np.sin(arg), np.cos(arg)], axis=1).astype(dtype)
    for w, b in net["hidden"]:
        h = gelu(linear(h, w.astype(dtype), b.astype(dtype)))
    return linear(h, net["out_w"].astype(dtype), net["out_b"].astype(dtype))


def _torch_linspace(start, end, steps):
    """torch.linspace fp32 kernel: step=(end-start)/(steps-1); first half counts up from
    start, second half counts down from end (ATen RangeFactories)."""
    start32, end32 = F32(start), F32(end)
    step = F32((end32 - start32) / F32(steps - 1))
    idx = np.arange(steps)
    half = steps // 2
    up = (start32 + step * idx.astype(F32)).astype(F32)
    down = (end32 - step * (steps - 1 - idx).astype(F32)).astype(F32)
    return np.where(idx < half, up, down).astype(F32)


def fourier_mlp(x, emb_t, mlp, dtype):
    """FourierMLP.forward, models/mlp.py:114-122, with the x-independent time-embedding
    row `emb_t` (= timestep_embed(t), :116) supplied by the caller."""
    h = linear(x, mlp["in_w"].astype(dtype), mlp["in_b"].astype(dtype)) + emb_t
    for w, b in mlp["hidden"]:
        h = linear(gelu(h), w.astype(dtype), b.astype(dtype))
    return linear(gelu(h), mlp["out_w"].astype(dtype), mlp["out_b"].astype(dtype))


# --------------------------------------------------------------------------------------
# distributions  (distr/*.py)
# --------------------------------------------------------------------------------------
LOG_2PI = math.log(2.0 * math.pi)


def diag_gauss_log_prob(x, loc, scale):
    """Gauss / IsotropicGauss log_prob (distr/gauss.py:131-140, :215-220; log_norm_const=0):
    sum_j [-((x-m)/s)^2/2 - log s - log(2 pi)/2]."""
    z = (x - loc) / scale
    return (-0.5 * z * z - np.log(scale) - x.dtype.type(0.5 * LOG_2PI)).sum(axis=-1, keepdims=True)


def diag_gauss_score(x, loc, scale):
    """Gauss.score distr/gauss.py:182-183, IsotropicGauss.score :222-223."""
    return (loc - x) / (scale * scale)


def target_log_prob_and_score(tg, x, need_score=True):
    """unnorm_log_prob (B,1) and score (B,d) of the supported targets."""
    dt = x.dtype
    kind = tg["kind"]
    if kind == "gmm":
        # MixtureSameFamily(Categorical(w), Independent(Normal(loc, scale),1)).log_prob
        # (distr/gauss.py:119-140); the reference obtains the score by autograd
        # (distr/base.py:130-137) — analytic form: SURVEY App. A.4.
        loc = tg["loc"].astype(dt)
        scale = tg["scale"].astype(dt)
        logw = tg["log_weights"].astype(dt)  # log_softmax of the mixture logits
        diff = x[:, None, :] - loc[None]  # (B,K,d)
        z = diff / scale[None]
        comp = (-0.5 * z * z - np.log(scale)[None] - dt.type(0.5 * LOG_2PI)).sum(-1) + logw[None]
        m = comp.max(axis=1, keepdims=True)
        e = np.exp(comp - m)
        s = e.sum(axis=1, keepdims=True)
        logp = m + np.log(s) + dt.type(tg.get("log_norm_const", 0.0))
        score = None
        if need_score:
            p = e / s
            score = -(p[:, :, None] * diff / (scale * scale)[None]).sum(axis=1)
        return logp.astype(dt), score
    if kind == "multiwell":
        # DoubleWell distr/double_well.py:39-45 on the first n_dw dims, N(shift, 1) (constant
        # folded away, :128-133) on the rest: MultiWell :165-179.  DoubleWell == n_dw = d = 1.
        n = int(tg["n_dw"])
        sep, shift = dt.type(tg["separation"]), dt.type(tg["shift"])
        y = x - shift
        a = y[:, :n]
        logp = -((a * a - sep) ** 2).sum(-1, keepdims=True)
        score = None
        if n < x.shape[1]:
            logp = logp - 0.5 * (y[:, n:] ** 2).sum(-1, keepdims=True)
        if need_score:
            score = np.concatenate([-4.0 * (a * a - sep) * a, -y[:, n:]], axis=1)
        return logp.astype(dt), None if score is None else score.astype(dt)
    if kind == "funnel":
        # Funnel distr/funnel.py:57-80: x0 ~ N(0, var), x_{1:} | x0 ~ N(0, exp(x0) I)
        var = dt.type(tg["variance"])
        d = x.shape[1]
        x0 = x[:, :1]
        xo = x[:, 1:]
        sq = (xo * xo).sum(-1, keepdims=True)
        inv = np.exp(-x0)
        lp_first = dt.type(-0.5 * math.log(2.0 * math.pi * float(var))) - 0.5 * x0 * x0 / var
        lp_other = -(d - 1) * (x0 + dt.type(LOG_2PI)) / 2.0 - 0.5 * sq * inv
        logp = lp_first + lp_other + dt.type(tg.get("log_norm_const", 0.0))
        score = None
        if need_score:
            s0 = -x0 / var - 0.5 * (d - 1) + 0.5 * sq * inv
            score = np.concatenate([s0, -xo * inv], axis=1)
        return logp.astype(dt), None if score is None else score.astype(dt)
    if kind == "gauss":
        loc, scale = tg["loc"].astype(dt), tg["scale"].astype(dt)
        logp = diag_gauss_log_prob(x, loc, scale) + dt.type(tg.get("log_norm_const", 0.0))
        return logp.astype(dt), diag_gauss_score(x, loc, scale).astype(dt) if need_score else None
    if kind == "nice":
        return nice_log_prob_and_score(tg, x, need_score)
    raise ValueError(f"oracle: unsupported target kind {kind!r}")


def _softplus(z):
    return np.logaddexp(z, z.dtype.type(0.0))


def nice_log_prob_and_score(tg, x, need_score=True):
    """NiceModel.log_prob (distr/nice.py:178-190) wrapped as Nice.unnorm_log_prob (:262-263), and its
    score.  The reference differentiates log_prob with autograd (distr/base.py:130-137); here the
    backward pass of the additive couplings is written out:
      Coupling.forward :64-95   x.reshape(B, W/2, 2): mask_config=1 -> on = [:, :, 0], off = [:, :, 1];
                                on <- on + out_block(relu-MLP(off))
      Scaling.forward  :109-124 z = h * exp(scale), log|det J| = sum(scale)
      StandardLogistic.log_prob :21-29  -(softplus(z) + softplus(-z)), d/dz = -tanh(z/2)
    tg: {"couplings": [{"mask_config": 0|1, "layers": [(w, b), ...]}], "scale": (d,), "log_norm_const"}."""
    dt = x.dtype
    B, W = x.shape
    h = x.reshape(B, W // 2, 2).copy()
    saved = []
    for c in tg["couplings"]:
        on_i = 0 if int(c["mask_config"]) else 1
        off = h[:, :, 1 - on_i]
        a = off
        masks = []
        layers = c["layers"]
        for w, b in layers[:-1]:
            a = np.maximum(linear(a, np.asarray(w, dt), np.asarray(b, dt)), dt.type(0))
            masks.append(a > 0)
        w, b = layers[-1]
        h[:, :, on_i] = h[:, :, on_i] + linear(a, np.asarray(w, dt), np.asarray(b, dt))
        saved.append((on_i, masks))
    scale = np.asarray(tg["scale"], dt).reshape(1, W)
    z = h.reshape(B, W) * np.exp(scale)
    logp = (-(_softplus(z) + _softplus(-z))).sum(-1, keepdims=True) + scale.sum() + dt.type(tg.get("log_norm_const", 0.0))
    if not need_score:
        return logp.astype(dt), None
    g = (-np.tanh(z * dt.type(0.5)) * np.exp(scale)).astype(dt).reshape(B, W // 2, 2)
    for c, (on_i, masks) in zip(reversed(tg["couplings"]), reversed(saved)):
        layers = c["layers"]
        delta = g[:, :, on_i] @ np.asarray(layers[-1][0], dt)           # through out_block
        for (w, _), m in zip(reversed(layers[1:-1]), reversed(masks[1:])):
            delta = (delta * m) @ np.asarray(w, dt)
        delta = (delta * masks[0]) @ np.asarray(layers[0][0], dt)        # through in_block
        g[:, :, 1 - on_i] = g[:, :, 1 - on_i] + delta
    return logp.astype(dt), g.reshape(B, W).astype(dt)


# --------------------------------------------------------------------------------------
# SDE coefficient functions  (eq/sdes.py)
# --------------------------------------------------------------------------------------
def _lerp(a, b, w):
    """torch.lerp(start, end, weight): start + w*(end-start) if w < 0.5 else end - (end-start)*(1-w)."""
    w = np.asarray(w)
    lo = a + w * (b - a)
    hi = b - (b - a) * (1 - w)
    return np.where(w < 0.5, lo, hi)


def sde_coeffs(sde, s, t, dim, dtype):
    """Per-step scalars: mu_c (drift = mu_c * x), sigma, int_s^t div(mu).
    VP eq/sdes.py:222-245 (generative: beta runs max->min, drift +beta/2 x), ConstOU/ScaledBM
    :141-154,:175-177; drift_div_int :88-91."""
    dt_ = (t - s).astype(dtype)
    if sde is None or sde["kind"] == "none":
        z = np.zeros_like(dt_)
        return z, z, z
    if sde["kind"] == "vp":
        bmin, bmax = dtype(sde["beta_min"]), dtype(sde["beta_max"])
        T_end = dtype(sde["terminal_t"])
        sign = dtype(sde.get("sign", 1.0))
        if sign > 0:
            beta_s, beta_t = _lerp(bmax, bmin, s / T_end), _lerp(bmax, bmin, t / T_end)
        else:
            beta_s, beta_t = _lerp(bmin, bmax, s / T_end), _lerp(bmin, bmax, t / T_end)
        beta_s, beta_t = beta_s.astype(dtype), beta_t.astype(dtype)
        mu = sign * dtype(0.5) * beta_s
        sigma = dtype(sde.get("scale", 1.0)) * np.sqrt(beta_s)
        div_int = sign * dtype(0.25) * (beta_t + beta_s) * dt_ * dtype(dim)
        return mu.astype(dtype), sigma.astype(dtype), div_int.astype(dtype)
    if sde["kind"] == "const_ou":
        sign = dtype(sde.get("sign", 1.0))
        a, b = dtype(sde["drift_coeff"]), dtype(sde["diff_coeff"])
        mu = np.full_like(dt_, sign * a)
        sigma = np.full_like(dt_, b)
        return mu, sigma, (sign * a * dt_ * dtype(dim)).astype(dtype)
    raise ValueError(sde["kind"])


# --------------------------------------------------------------------------------------
# the control  (models/reparam.py)
# --------------------------------------------------------------------------------------
def _clip(v, c):
    """clip_and_log's clip (utils/common.py:83-84); None = no clip."""
    if c is None or not np.isfinite(c):
        return v
    return np.clip(v, -c, c)


def control(spec, i, s, x, emb_row, gate_row, sigma_s, dtype):
    """g = generative_ctrl(s, x).  ClippedCtrl reparam.py:35-36, ScoreCtrl :78-83,
    LerpCtrl :131-162, LerpPriorCtrl :165-181, LerpTargetCtrl :184-200."""
    c = spec["ctrl"]
    nn_out = fourier_mlp(x, emb_row, spec["mlp"], dtype)
    g = _clip(nn_out, c.get("clip_model"))
    kind = c["kind"]
    if kind == "clipped":
        return g.astype(dtype)
    _, tscore = target_log_prob_and_score(spec["target"], x)
    if kind == "score":
        inner = tscore
    else:
        w = dtype(s / dtype(spec["sde"]["terminal_t"]))
        pscore = diag_gauss_score(x, spec["prior"]["loc"].astype(dtype), spec["prior"]["scale"].astype(dtype))
        if kind == "lerp":
            inner = _lerp(pscore, tscore, w)
        elif kind == "lerp_prior":
            inner = (dtype(1.0) - w) * pscore
        elif kind == "lerp_target":
            inner = w * tscore
        else:
            raise ValueError(kind)
    score = dtype(c.get("scale_score", 1.0)) * _clip(inner.astype(dtype), c.get("clip_score"))
    if gate_row is not None:
        score = score * _clip(gate_row, c.get("clip_model"))
    if kind == "score":
        return (g + score).astype(dtype)
    return (g + sigma_s * score).astype(dtype)


# --------------------------------------------------------------------------------------
# the rollout  (losses/oc.py)
# --------------------------------------------------------------------------------------
def rollout(spec, x0, noise=None, seed=None, traj_offset=0, dtype=np.float32):
    """simulate() of TimeReversalLoss (losses/oc.py:156-230), ReferenceSDELoss (:286-343),
    ExponentialIntegratorSDELoss (:400-457) — selected by spec["loss"]["kind"].
    Returns x_T (B,d), rnd (B,1), xs (T+1,B,d) or None."""
    dtype = np.dtype(dtype).type
    ls = spec["loss"]
    ts = np.asarray(spec["ts"], dtype=dtype)
    T = ts.shape[0] - 1
    x = np.asarray(x0, dtype=dtype).copy()
    B, d = x.shape
    kind = ls["kind"]
    train = bool(ls["train"])
    method = ls["method"]
    compute_ito = bool(ls["compute_ito"])

    s_all, t_all = ts[:-1], ts[1:]
    emb = time_embed(s_all, spec["mlp"]["time_embed"], dtype)  # (T,C)  models/mlp.py:116
    gate = None
    if spec.get("gate") is not None:
        gate = time_embed(s_all, spec["gate"], dtype)  # (T,1|d)  models/reparam.py:68-76
    mu_c, sigma, div_int = sde_coeffs(spec.get("sde"), s_all, t_all, d, dtype)

    # initial cost (losses/oc.py:168-172, :296, :410)
    if kind == "time_reversal" and not (train and method in ("kl", "kl_ito")):
        rnd = diag_gauss_log_prob(x, spec["prior"]["loc"].astype(dtype), spec["prior"]["scale"].astype(dtype))
    else:
        rnd = np.zeros((B, 1), dtype=dtype)

    xs = [x.copy()] if ls.get("return_traj") else None
    ref_ctrl = bool(ls.get("reference_ctrl"))
    for i in range(T):
        s, t = s_all[i], t_all[i]
        dt_ = dtype(t - s)
        g = control(spec, i, s, x, emb[i:i + 1], None if gate is None else gate[i:i + 1], sigma[i], dtype)
        if noise is not None:
            eps = np.asarray(noise[i], dtype=dtype)
        else:
            eps = philox.normal_block(seed, np.arange(traj_offset, traj_offset + B), i, d).astype(dtype)
        if kind == "exp_integrator":
            # losses/oc.py:429-443
            alpha, sig = dtype(ls["alpha"]), dtype(ls["sigma"])
            beta_k = dtype(np.clip(alpha * np.sqrt(dt_), 0, 1))
            alpha_k = dtype(np.sqrt(dtype(1.0) - beta_k * beta_k))
            rnd = rnd + beta_k * beta_k * sig * sig * (dtype(0.5) * (g * g).sum(-1, keepdims=True))
            x_new = x * alpha_k + (beta_k * beta_k) * (sig * sig) * g + sig * beta_k * eps
            if compute_ito:
                rnd = rnd + (sig * g * eps * beta_k).sum(-1, keepdims=True)
            x = x_new.astype(dtype)
        else:
            if kind == "reference_sde" and ref_ctrl:
                # EulerDDS.reference_ctrl solver/oc.py:305-306
                r = sigma[i] * diag_gauss_score(x, spec["prior"]["loc"].astype(dtype), spec["prior"]["scale"].astype(dtype))
                gm = g - r
            else:
                gm = g
            # running cost: kl 1/2|gm|^2 dt (oc.py:208, :323); lv: gm*(u - (g+r)/2) dt with u == g
            # numerically (:204-206, :320-321) — identical value.
            rnd = rnd + dtype(0.5) * (gm * gm).sum(-1, keepdims=True) * dt_
            if kind == "time_reversal" and not train:
                rnd = rnd - div_int[i]  # oc.py:210-211
            db = eps * np.sqrt(dt_)
            x_new = x + (mu_c[i] * x + sigma[i] * g) * dt_ + sigma[i] * db  # oc.py:214-215
            if compute_ito:
                rnd = rnd + (gm * db).sum(-1, keepdims=True)  # oc.py:218-219, :330-331
            x = x_new.astype(dtype)
        rnd = rnd.astype(dtype)
        if xs is not None:
            xs.append(x.copy())

    # terminal cost (oc.py:225, :337, :449-450)
    logp, _ = target_log_prob_and_score(spec["target"], x, need_score=False)
    logp = _clip(logp, spec["target"].get("clip_target"))
    if kind == "time_reversal":
        rnd = rnd - logp
    else:
        rnd = rnd + diag_gauss_log_prob(x, spec["ref"]["loc"].astype(dtype), spec["ref"]["scale"].astype(dtype)) - logp
    return x, rnd.astype(dtype), (np.stack(xs) if xs is not None else None)


# --------------------------------------------------------------------------------------
# Langevin dynamics  (eq/integrator.py:79-127 on eq/sdes.py:38-65; solver/langevin.py:34-63)
# --------------------------------------------------------------------------------------
def langevin_integrate(tg, x0, timesteps, out_ts, diff_coeff, clip_score, noise, eps=1e-8, dtype=np.float32):
    """EulerIntegrator.integrate(LangevinSDE(target.score, diff_coeff, clip_score), ts=out_ts, x_init=x0,
    timesteps=timesteps) with the normal draws `noise` (n_steps, B, d).  Returns xs (len(out_ts), B, d)."""
    dtype = np.dtype(dtype).type
    xs = np.asarray(x0, dtype=dtype).copy()
    timesteps = np.asarray(timesteps, dtype=dtype)
    out_ts = np.asarray(out_ts, dtype=dtype)
    out, cnt = [], 0
    for i, (s, t) in enumerate(zip(timesteps[:-1], timesteps[1:])):
        _, score = target_log_prob_and_score(tg, xs)
        drift = _clip(score * dtype(diff_coeff) ** 2 / dtype(2.0), clip_score)               # eq/sdes.py:54-61
        nz = np.asarray(noise[i], dtype=dtype) * np.sqrt(dtype(t - s))                        # integrator.py:116
        xt = (xs + drift * dtype(t - s) + dtype(diff_coeff) * nz).astype(dtype)              # :119
        if cnt < out_ts.shape[0] and out_ts[cnt] <= t + dtype(eps):                           # :121-123
            ind = int(np.searchsorted(out_ts[cnt:], t + dtype(eps), side="right"))
            w = ((out_ts[cnt:cnt + ind].reshape(-1, 1, 1) - s) / (t - s)).astype(dtype)
            out.append(_lerp(xs[None], xt[None], w).astype(dtype))                            # interpolate(), :66-77
            cnt += ind
        xs = xt
    return np.concatenate(out, axis=0)


def ou_coefficients(case, timesteps):
    """Per-step (mu, sigma) of the OU family (eq/sdes.py:125-269) and, for the PIS inference process (solver/oc.py:200-208;
    ControlledSDE eq/sdes.py:296-305), the control's (diff, loc factor, 1/var) at the reversed time — fp32 like the reference."""
    f = np.float32
    s = np.asarray(timesteps, dtype=f)[:-1]
    sign = f(1.0 if case["generative"] else -1.0)
    if case["sde"] == "vp":
        bmin, bmax, T_end = f(0.1), f(10.0), f(1.0)
        w = (s / T_end).astype(f)
        a, b = (bmax, bmin) if case["generative"] else (bmin, bmax)
        beta = _lerp(np.full_like(s, a), np.full_like(s, b), w).astype(f)                      # VP._diff_coeff_sq_t :222-229
        mu, sigma = (sign * f(0.5) * beta).astype(f), np.sqrt(beta).astype(f)
    elif case["sde"] == "bm_pis":
        T_end = f(5.0)
        mu, sigma = np.zeros_like(s), np.full_like(s, f(0.4472135954999579))
    else:
        T_end = f(1.0)
        mu, sigma = np.full_like(s, sign * f(4.5)), np.full_like(s, f(3.0))
    ctrl = None
    if case.get("ctrl") == "pis":
        tt = s if case["generative"] else (T_end - s).astype(f)                                # eq/sdes.py:302-303
        var = (sigma * sigma * tt).astype(f)                                                    # ScaledBM.marginal_params :179-188
        ctrl = dict(csig=sigma.copy(), cinv=(f(1.0) / var).astype(f), cmax=f(1e5))              # Delta prior at the origin: loc = 0
    return mu, sigma, ctrl


def affine_integrate(mu, sigma, ctrl, x0, timesteps, out_ts, noise, eps=1e-8, increments=False, dtype=np.float32):
    """EulerIntegrator.integrate (eq/integrator.py:93-127) for x-affine SDEs: drift mu_i x (+ sigma_i csig_i min((0 - x) cinv_i,
    cmax) for the PIS bridge control), diffusion sigma_i; `noise` standard normals (or Brownian increments)."""
    dtype = np.dtype(dtype).type
    xs = np.asarray(x0, dtype=dtype).copy()
    timesteps = np.asarray(timesteps, dtype=dtype)
    out_ts = np.asarray(out_ts, dtype=dtype)
    out, cnt = [], 0
    for i, (s, t) in enumerate(zip(timesteps[:-1], timesteps[1:])):
        drift = dtype(mu[i]) * xs
        if ctrl is not None:
            score = np.minimum((dtype(0.0) - xs) * dtype(ctrl["cinv"][i]), dtype(ctrl["cmax"]))
            drift = drift + dtype(sigma[i]) * (dtype(ctrl["csig"][i]) * score)
        nz = np.asarray(noise[i], dtype=dtype) * (dtype(1.0) if increments else np.sqrt(dtype(t - s)))
        xt = (xs + drift * dtype(t - s) + dtype(sigma[i]) * nz).astype(dtype)
        if cnt < out_ts.shape[0] and out_ts[cnt] <= t + dtype(eps):
            ind = int(np.searchsorted(out_ts[cnt:], t + dtype(eps), side="right"))
            w = ((out_ts[cnt:cnt + ind].reshape(-1, 1, 1) - s) / (t - s)).astype(dtype)
            out.append(_lerp(xs[None], xt[None], w).astype(dtype))
            cnt += ind
        xs = xt
    return np.concatenate(out, axis=0)


# --------------------------------------------------------------------------------------
# reductions  (losses/oc.py:50-123)
# --------------------------------------------------------------------------------------
def loss_from_rnd(rnd, method, max_rnd=None, traj_per_sample=1, sample_mask=None):
    """BaseOCLoss.filter + compute_loss (losses/oc.py:50-92). Returns (loss, n_filtered)."""
    rnd = np.asarray(rnd, dtype=np.float64).reshape(-1)
    mask = np.isfinite(rnd) if max_rnd is None else (rnd < max_rnd)
    if sample_mask is not None:
        mask = mask & sample_mask.reshape(-1)
    if method == "lv_traj":
        r = rnd.reshape(traj_per_sample, -1)
        m = mask.reshape(traj_per_sample, -1).all(axis=0)
        n_filtered = traj_per_sample * int(m.size - m.sum())
        return float(r[:, m].var(axis=0, ddof=1).mean()), n_filtered
    n_filtered = int(mask.size - mask.sum())
    kept = rnd[mask]
    if method == "lv":
        return float(kept.var(ddof=1)), n_filtered
    return float(kept.mean()), n_filtered


def results_from_rnd(rnd, compute_weights):
    """BaseOCLoss.compute_results (losses/oc.py:94-123)."""
    rnd = np.asarray(rnd, dtype=np.float64).reshape(-1)
    neg = -rnd
    if compute_weights:
        mx = neg.max()
        w = np.exp(neg - mx)
        return {
            "log_norm_const_lb_ito": float(neg.mean()),
            "log_norm_const_is": float(np.log(w.mean()) + mx),
            "eval/lv_loss": float(rnd.var(ddof=1)),
            "weights": w.reshape(-1, 1),
        }
    return {"log_norm_const_lb": float(neg.mean())}
