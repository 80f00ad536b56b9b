"""TEST INFRASTRUCTURE / CPU BASELINE — the reference's rollout restated with the SAME library
ops the reference issues on the host (torch eager: `F.linear`, `nn.GELU`, `torch.distributions`
MixtureSameFamily + an autograd score for the GMM, one `torch.randn_like` per step), driven
from the same spec dict as `oracle/rollout.py`.

Why it exists: the unmodified reference is Python and cannot travel to the GPU box
(`/root/reference` is absent there), and the numpy oracle is single-threaded.  This port is what
`bench.py` times as `cpu_baseline` (kind "port") and as `--impl reference`: per time step it
performs the reference's op sequence (losses/oc.py:176-222, models/mlp.py:71-122 including the
per-row recomputation of the time embedding, models/reparam.py:78-162, distr/base.py:130-137)
on all host threads.  Pinned like the numpy oracle: tests/test_oracle_golden.py checks it
against the golden vectors frozen from the unmodified reference.

Never imported by the product (`sde_sampler_b200/`).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import distributions


_DEVICE = [torch.device("cpu")]  # set by rollout(..., device=...): the same op sequence runs on the host cores or, as
                                  # bench.py's `gpu_eager_baseline`, as stock eager PyTorch kernels on the GPU


def _t(a):
    if isinstance(a, torch.Tensor):
        return a.to(device=_DEVICE[0], dtype=torch.float32)
    return torch.as_tensor(np.asarray(a, dtype=np.float32), device=_DEVICE[0])


class _TimeEmbed:
    """models/mlp.py:43-82"""

    def __init__(self, p):
        self.phase = _t(p["phase"]).reshape(1, -1)
        c = self.phase.shape[1]
        self.coeff = torch.linspace(start=0.1, end=100, steps=c).unsqueeze(0).to(_DEVICE[0])
        self.hidden = [(_t(w), _t(b)) for w, b in p["hidden"]]
        self.out_w, self.out_b = _t(p["out_w"]), _t(p["out_b"])

    def __call__(self, t):
        t = t.view(-1, 1).float()
        e = torch.cat([torch.sin(self.coeff * t + self.phase), torch.cos(self.coeff * t + self.phase)], dim=1)
        for w, b in self.hidden:
            e = F.gelu(F.linear(e, w, b))
        return F.linear(e, self.out_w, self.out_b)


class _FourierMLP:
    """models/mlp.py:85-122 — the time embedding is recomputed on B identical rows, as the reference does."""

    def __init__(self, p):
        self.in_w, self.in_b = _t(p["in_w"]), _t(p["in_b"])
        self.te = _TimeEmbed(p["time_embed"])
        self.hidden = [(_t(w), _t(b)) for w, b in p["hidden"]]
        self.out_w, self.out_b = _t(p["out_w"]), _t(p["out_b"])

    def __call__(self, t, x):
        t = t.view(-1, 1).expand(x.shape[0], 1).float()
        h = F.linear(x, self.in_w, self.in_b) + self.te(t)
        for w, b in self.hidden:
            h = F.linear(F.gelu(h), w, b)
        return F.linear(F.gelu(h), self.out_w, self.out_b)


def _clip(v, c):
    if c is None or not math.isfinite(c):
        return v
    return v.clip(-c, c)


class _Target:
    def __init__(self, tg, dim):
        self.kind, self.dim = tg["kind"], dim
        self.lnc = float(tg.get("log_norm_const", 0.0) or 0.0)
        self.clip_target = tg.get("clip_target")
        if self.kind == "gmm":
            loc, scale = _t(tg["loc"]), _t(tg["scale"])
            logits = _t(tg["log_weights"]).expand(loc.shape[0]).contiguous()
            self.distr = distributions.MixtureSameFamily(
                distributions.Categorical(logits=logits),
                distributions.Independent(distributions.Normal(loc, scale), 1))  # distr/gauss.py:119-128
        elif self.kind == "gauss":
            self.loc, self.scale = _t(tg["loc"]), _t(tg["scale"])
        elif self.kind == "multiwell":
            self.n, self.sep, self.shift = int(tg["n_dw"]), float(tg["separation"]), float(tg["shift"])
        elif self.kind == "funnel":
            self.var = float(tg["variance"])
        elif self.kind == "nice":
            self.couplings = [(int(c["mask_config"]), [(_t(w), _t(b)) for w, b in c["layers"]]) for c in tg["couplings"]]
            self.scale = _t(tg["scale"]).reshape(1, -1)

    def _nice_log_prob(self, x):
        # NiceModel.log_prob with the reference's own op sequence (distr/nice.py:64-95, :109-124, :178-190)
        B, W = x.shape
        for mc, layers in self.couplings:
            xr = x.reshape(B, W // 2, 2)
            on, off = (xr[:, :, 0], xr[:, :, 1]) if mc else (xr[:, :, 1], xr[:, :, 0])
            a = off
            for w, b in layers[:-1]:
                a = F.relu(F.linear(a, w, b))
            on = on + F.linear(a, *layers[-1])
            x = (torch.stack((on, off), dim=2) if mc else torch.stack((off, on), dim=2)).reshape(B, W)
        z = x * torch.exp(self.scale)
        return torch.sum(-(F.softplus(z) + F.softplus(-z)), dim=1) + torch.sum(self.scale)

    def unnorm_log_prob(self, x):
        if self.kind == "nice":
            return self._nice_log_prob(x).unsqueeze(-1) + self.lnc
        if self.kind == "gmm":
            return self.distr.log_prob(x).unsqueeze(-1) + self.lnc
        if self.kind == "gauss":
            z = (x - self.loc) / self.scale
            return (-0.5 * z * z - self.scale.log() - 0.5 * math.log(2 * math.pi)).sum(-1, keepdim=True) + self.lnc
        if self.kind == "multiwell":
            y = x - self.shift
            a = y[:, : self.n]
            lp = -((a ** 2 - self.sep) ** 2).sum(-1, keepdim=True)
            if self.n < x.shape[1]:
                lp = lp - 0.5 * (y[:, self.n:] ** 2).sum(-1, keepdim=True)
            return lp
        x0, xo = x[:, :1], x[:, 1:]  # funnel, distr/funnel.py:57-69
        lp0 = -0.5 * math.log(2 * math.pi * self.var) - 0.5 * x0 * x0 / self.var
        lpo = -(self.dim - 1) * (x0 + math.log(2 * math.pi)) / 2 - 0.5 * (xo * xo).sum(-1, keepdim=True) * torch.exp(-x0)
        return lp0 + lpo + self.lnc

    def score(self, x):
        if self.kind in ("gmm", "nice"):  # Distribution.score: autograd of the log-density (distr/base.py:130-137)
            x = x.detach().requires_grad_(True)
            with torch.enable_grad():
                lr = self.unnorm_log_prob(x).sum()
                return torch.autograd.grad(lr, x)[0]
        if self.kind == "gauss":
            return (self.loc - x) / self.scale ** 2
        if self.kind == "multiwell":
            y = x - self.shift
            a = y[:, : self.n]
            return torch.cat([-4.0 * (a ** 2 - self.sep) * a, -y[:, self.n:]], dim=1)
        x0, xo = x[:, :1], x[:, 1:]
        inv = torch.exp(-x0)
        s0 = -x0 / self.var - 0.5 * (self.dim - 1) + 0.5 * (xo * xo).sum(-1, keepdim=True) * inv
        return torch.cat([s0, -xo * inv], dim=1)


def _gauss_logp(x, g):
    loc, scale = _t(g["loc"]), _t(g["scale"])
    z = (x - loc) / scale
    return (-0.5 * z * z - scale.log() - 0.5 * math.log(2 * math.pi)).sum(-1, keepdim=True)


def _gauss_score(x, g):
    return (_t(g["loc"]) - x) / _t(g["scale"]) ** 2


def _sde(sde, s, t, dim):
    """mu coefficient, sigma, int div — eq/sdes.py (scalar tensors, computed per step like the reference)."""
    if sde is None:
        z = torch.zeros((), device=_DEVICE[0])
        return z, z, z
    dt = t - s
    if sde["kind"] == "vp":
        bmin, bmax, T_end = torch.tensor(sde["beta_min"], device=_DEVICE[0]), torch.tensor(sde["beta_max"], device=_DEVICE[0]), sde["terminal_t"]
        sign = float(sde.get("sign", 1.0))
        a, b = (bmax, bmin) if sign > 0 else (bmin, bmax)
        beta_s, beta_t = torch.lerp(a, b, s / T_end), torch.lerp(a, b, t / T_end)
        return sign * 0.5 * beta_s, float(sde.get("scale", 1.0)) * torch.sqrt(beta_s), sign * 0.25 * (beta_t + beta_s) * dt * dim
    sign = float(sde.get("sign", 1.0))
    return (torch.tensor(sign * sde["drift_coeff"], device=_DEVICE[0]), torch.tensor(float(sde["diff_coeff"]), device=_DEVICE[0]),
            sign * sde["drift_coeff"] * dt * dim)


@torch.no_grad()
def rollout(spec, x0, noise=None, generator=None, device="cpu", as_numpy=True):
    """Same contract as oracle.rollout.rollout (noise (T,B,d) injected, or drawn with torch.randn).
    `device="cuda"` runs the identical eager op sequence on the GPU (the "existing Blackwell path": stock PyTorch
    kernels, ~370 launches per time step, fp32 with TF32 off); `as_numpy=False` leaves the results on the device."""
    _DEVICE[0] = torch.device(device)
    try:
        return _rollout(spec, x0, noise, generator, as_numpy)
    finally:
        _DEVICE[0] = torch.device("cpu")


def _rollout(spec, x0, noise, generator, as_numpy):
    ls, cd = spec["loss"], spec["ctrl"]
    ts = _t(spec["ts"])
    x = _t(x0).clone()
    B, d = x.shape
    net = _FourierMLP(spec["mlp"])
    gate = _TimeEmbed(spec["gate"]) if spec.get("gate") is not None else None
    target = _Target(spec["target"], d)
    kind, train, method = ls["kind"], bool(ls["train"]), ls["method"]
    compute_ito = bool(ls["compute_ito"])
    cm, cs = cd.get("clip_model"), cd.get("clip_score")
    if kind == "time_reversal" and not (train and method in ("kl", "kl_ito")):
        rnd = _gauss_logp(x, spec["prior"])
    else:
        rnd = torch.zeros(B, 1, device=_DEVICE[0])
    xs = [x] if ls.get("return_traj") else None
    for i, (s, t) in enumerate(zip(ts[:-1], ts[1:])):
        dt = t - s
        mu, sigma, div_int = _sde(spec.get("sde"), s, t, d)
        g = _clip(net(s, x), cm)
        ck = cd["kind"]
        if ck != "clipped":
            if ck == "score":
                inner = target.score(x)
            else:
                w = s / spec["sde"]["terminal_t"]
                if ck == "lerp":
                    inner = torch.lerp(_gauss_score(x, spec["prior"]), target.score(x), w)
                elif ck == "lerp_prior":
                    inner = (1 - w) * _gauss_score(x, spec["prior"])
                else:
                    inner = w * target.score(x)
            sc = float(cd.get("scale_score", 1.0)) * _clip(inner, cs)
            if gate is not None:
                sc = sc * _clip(gate(s), cm)
            g = g + (sc if ck == "score" else sigma * sc)
        eps = _t(noise[i]) if noise is not None else torch.randn(x.shape, generator=generator, device=_DEVICE[0])
        if kind == "exp_integrator":
            alpha, sg = float(ls["alpha"]), float(ls["sigma"])
            bk = (alpha * dt.sqrt()).clip(0, 1)
            ak = (1 - bk ** 2).sqrt()
            rnd = rnd + bk ** 2 * sg ** 2 * (0.5 * (g ** 2).sum(-1, keepdim=True))
            x_new = x * ak + bk ** 2 * sg ** 2 * g + sg * bk * eps
            if compute_ito:
                rnd = rnd + (sg * g * eps * bk).sum(-1, keepdim=True)
            x = x_new
        else:
            gm = g
            if kind == "reference_sde" and ls.get("reference_ctrl"):
                gm = g - sigma * _gauss_score(x, spec["prior"])
            rnd = rnd + 0.5 * (gm ** 2).sum(-1, keepdim=True) * dt
            if kind == "time_reversal" and not train:
                rnd = rnd - div_int
            db = eps * dt.sqrt()
            x_new = x + (mu * x + sigma * g) * dt + sigma * db
            if compute_ito:
                rnd = rnd + (gm * db).sum(-1, keepdim=True)
            x = x_new
        if xs is not None:
            xs.append(x)
    lp = _clip(target.unnorm_log_prob(x), spec["target"].get("clip_target"))
    if kind == "time_reversal":
        rnd = rnd - lp
    else:
        rnd = rnd + _gauss_logp(x, spec["ref"]) - lp
    if not as_numpy:
        return x, rnd, (torch.stack(xs) if xs is not None else None)
    return x.cpu().numpy(), rnd.cpu().numpy(), (torch.stack(xs).cpu().numpy() if xs is not None else None)
