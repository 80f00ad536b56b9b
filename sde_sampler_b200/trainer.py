"""The caller's side of the rollout: the tail of `Trainable.step` (solver/base.py:409-439) as ONE C-ABI
call — `sdes_trainer_step`: loss / gradient checks, `clip_grad_norm_`, `torch.optim.Adam` and the EMA update on a flat
parameter buffer, two kernel launches, the skip decision taken on the device (the reference synchronises the host for
each of them: `.item()`, `all(p.grad.isfinite().all() ...)`).

    optim = FusedAdamEMA(ctrl.parameters(), lr=0.005, weight_decay=1e-7,          # conf/solver/oc_base.yaml:26-29
                         grad_clip_norm=1.0,                                        # conf/utils/grad_clip.yaml
                         ema=dict(decay=0.9999, inv_gamma=1, power=0.9, update_after_step=..., update_every=5))
    optim.zero_grad(); loss, _ = fused_loss(ts, x0, ...); (scale_loss * loss).backward(); optim.step(loss=loss)

`FusedAdamEMA` is a `torch.optim.Optimizer`; the parameters are re-pointed to views of one flat fp32 buffer so the kernel
updates them in place.  There is no CPU path.

LR schedulers: the reference calls `scheduler.step()` only when the optimizer stepped (`loss_ok and grad_ok`,
solver/base.py:423-436).  Here that decision is taken on the DEVICE, so a scheduler (or `MultiStepParams`) attached to this
optimizer must be gated on it — `if optim.step(loss=loss, sync_skip=True): scheduler.step()` (one host read of the flag,
the same sync the reference pays), or `optim.stepped()` later.  Stepping a scheduler unconditionally advances the
schedule on skipped steps and drifts from the reference.

Checkpoints: `state_dict()` uses the `torch.optim.Adam` layout (`state` / `param_groups`, per-parameter `step`, `exp_avg`,
`exp_avg_sq`) plus a `fused` entry (device counters, EMA shadow); `load_state_dict` accepts that, a plain
`torch.optim.Adam` state dict of the reference's `Trainable` checkpoint, and — through `load_ema_state_dict` — a
`torch_ema.ExponentialMovingAverage` state dict (`shadow_params`, `num_updates`)."""
from __future__ import annotations

import contextlib
import ctypes as C
import math

import torch

from . import _cabi


def _inf(v) -> float:
    return math.inf if v is None else float(v)


class FusedAdamEMA(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, *,
                 grad_clip_norm: float | None = None, max_grad: float | None = None, max_loss: float | None = None,
                 ema: dict | None = None):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("Invalid lr / eps / weight_decay")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedAdamEMA updates one flat parameter group")
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps:
            raise ValueError("no trainable parameters")
        dev = ps[0].device
        if dev.type != "cuda":
            raise _cabi.SdesError("the fused trainer step runs on a CUDA device only; there is no CPU path")
        if any(p.device != dev or p.dtype != torch.float32 for p in ps):
            raise NotImplementedError("parameters must be fp32 on one device")
        _cabi.lib()
        self._params = ps
        self.grad_clip_norm, self.max_grad, self.max_loss = grad_clip_norm, max_grad, max_loss
        self.ema = None if ema is None else dict(decay=0.9999, inv_gamma=1.0, power=2 / 3, update_after_step=100,
                                                  update_every=10, min_value=0.0) | dict(ema)
        with torch.no_grad():
            self.flat = torch.cat([p.detach().reshape(-1) for p in ps]).contiguous()
            o = 0
            for p in ps:  # parameters become views of the flat buffer: the kernel's in-place update is what the modules see
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                o += p.numel()
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema_shadow = self.flat.clone() if self.ema is not None else None  # torch_ema: shadow starts as a copy
        self.dev_state = torch.zeros(8, dtype=torch.float64, device=dev)
        self._ws = torch.empty(int(_cabi.lib().sdes_trainer_workspace_bytes()), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------------------------------
    def _flat_grads(self) -> torch.Tensor:
        missing = [i for i, p in enumerate(self._params) if p.grad is None]
        if missing:
            # torch.optim.Adam skips such parameters entirely (no weight decay, no moment decay); the flat kernel updates
            # every element, so a silent zero gradient would let them drift under weight_decay
            raise RuntimeError(f"parameters {missing} have no gradient: FusedAdamEMA updates the whole flat buffer (freeze them with "
                               "requires_grad=False before constructing the optimizer, or pass `grads=`)")
        return torch.cat([p.grad.reshape(-1) for p in self._params])

    def stepped(self) -> bool:
        """Did the last `step()` update the parameters (loss_ok and grad_ok, solver/base.py:423-436)?  One host read."""
        return bool(self.dev_state[5].item())

    @torch.no_grad()
    def step(self, closure=None, loss: torch.Tensor | None = None, grads: torch.Tensor | None = None, sync_skip: bool = False):
        """One `Trainable.step` tail.  `loss` (0-dim device tensor, optional) feeds the max_loss / isfinite check;
        `grads` (flat, blob order) may be passed instead of reading `p.grad`.  Returns None (no host sync), or with
        `sync_skip=True` whether the step was applied — gate `scheduler.step()` on it (module docstring)."""
        if closure is not None:
            raise NotImplementedError("closure")
        g = self._flat_grads() if grads is None else grads.detach().reshape(-1).to(torch.float32).contiguous()
        if g.numel() != self.flat.numel():
            raise ValueError("gradient size does not match the parameters")
        grp = self.param_groups[0]
        d = _cabi.TrainerStepDesc()
        d.struct_bytes = C.sizeof(_cabi.TrainerStepDesc)
        d.n = self.flat.numel()
        d.params, d.grads = self.flat.data_ptr(), g.data_ptr()
        d.exp_avg, d.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        d.ema_shadow = None if self.ema_shadow is None else self.ema_shadow.data_ptr()
        lt = None
        if loss is not None:
            lt = loss.detach().reshape(-1)[:1].to(torch.float32).contiguous()
            d.loss = lt.data_ptr()
        d.lr, (d.beta1, d.beta2), d.eps, d.weight_decay = float(grp["lr"]), grp["betas"], float(grp["eps"]), float(grp["weight_decay"])
        d.max_loss, d.max_grad, d.grad_clip_norm = _inf(self.max_loss), _inf(self.max_grad), _inf(self.grad_clip_norm)
        if self.ema is not None:
            e = self.ema
            d.ema_decay, d.ema_inv_gamma, d.ema_power, d.ema_min_value = e["decay"], e["inv_gamma"], e["power"], e["min_value"]
            d.ema_update_after_step, d.ema_update_every = int(e["update_after_step"]), int(e["update_every"])
        d.state = self.dev_state.data_ptr()
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        with torch.cuda.device(self.flat.device):
            stream = torch.cuda.current_stream(self.flat.device).cuda_stream
            _cabi.check(_cabi.lib().sdes_trainer_step(C.byref(d), C.c_void_p(stream)), "sdes_trainer_step")
        return self.stepped() if sync_skip else None

    def metrics(self) -> dict:
        """One host read of the device-side counters (the reference's train/* metrics, solver/base.py:421-451)."""
        s = self.dev_state.tolist()
        out = {"train/optim_steps": int(s[0]), "train/skipped_steps": int(s[1]), "train/grad_norm": s[3], "train/max_grad": s[4],
               "train/stepped": bool(s[5]), "train/grad_clip_coef": s[7]}
        if self.ema is not None:
            out["train/ema_num_updates"] = int(s[2])
            if s[6] >= 0:
                out["train/ema_decay"] = s[6]
        return out

    @contextlib.contextmanager
    def average_parameters(self):
        """torch_ema's `average_parameters()` (used by Trainable.evaluate, solver/base.py:342-346): the EMA weights are
        swapped in for the duration of the block."""
        if self.ema_shadow is None:
            yield
            return
        backup = self.flat.clone()
        self.flat.copy_(self.ema_shadow)
        try:
            yield
        finally:
            self.flat.copy_(backup)

    # ------------------------------------------------------------------------------------------ checkpoints
    def _segments(self):
        o = 0
        for i, p in enumerate(self._params):
            yield i, p, o, o + p.numel()
            o += p.numel()

    def state_dict(self) -> dict:
        """torch.optim.Adam layout + a `fused` entry (module docstring)."""
        st = self.dev_state.tolist()
        step = torch.tensor(float(st[0]))
        state = {i: {"step": step.clone(), "exp_avg": self.exp_avg[a:b].view(p.shape).clone(),
                     "exp_avg_sq": self.exp_avg_sq[a:b].view(p.shape).clone()} for i, p, a, b in self._segments()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self._params)))
        return {"state": state, "param_groups": [group],
                "fused": {"dev_state": self.dev_state.clone(), "ema_shadow": None if self.ema_shadow is None else self.ema_shadow.clone(),
                          "ema": None if self.ema is None else dict(self.ema)}}

    def load_state_dict(self, sd: dict):
        if "state" not in sd or "param_groups" not in sd:
            raise KeyError("FusedAdamEMA.load_state_dict expects the torch.optim layout: keys 'state' and 'param_groups' "
                           f"(got {sorted(sd)})")
        groups = sd["param_groups"]
        if len(groups) != 1:
            raise ValueError(f"checkpoint has {len(groups)} parameter groups; FusedAdamEMA holds one")
        ids = list(groups[0].get("params", range(len(self._params))))
        if len(ids) != len(self._params):
            raise ValueError(f"checkpoint holds {len(ids)} parameters, the optimizer {len(self._params)}")
        self.param_groups[0].update({k: v for k, v in groups[0].items() if k != "params"})
        steps = []
        for (i, p, a, b), pid in zip(self._segments(), ids):
            ent = sd["state"].get(pid, sd["state"].get(str(pid)))
            if ent is None:  # torch.optim.Adam creates state lazily: a parameter that never stepped has none
                self.exp_avg[a:b].zero_()
                self.exp_avg_sq[a:b].zero_()
                continue
            for key, buf in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                t = ent[key]
                if tuple(t.shape) != tuple(p.shape):
                    raise ValueError(f"state[{pid}][{key!r}] has shape {tuple(t.shape)}, parameter {i} has {tuple(p.shape)}")
                buf[a:b].copy_(t.reshape(-1).to(buf.dtype))
            steps.append(float(ent["step"]))
        if steps and max(steps) != min(steps):
            raise ValueError("per-parameter Adam step counts differ; FusedAdamEMA keeps one step counter")
        fused = sd.get("fused")
        if fused is not None:
            self.dev_state.copy_(fused["dev_state"])
            if self.ema_shadow is not None:
                if fused.get("ema_shadow") is not None:
                    if fused["ema_shadow"].numel() != self.ema_shadow.numel():
                        raise ValueError("EMA shadow size does not match the parameters")
                    self.ema_shadow.copy_(fused["ema_shadow"])
                else:  # no shadow in the checkpoint: restart the average from the current weights, counter at 0
                    self.ema_shadow.copy_(self.flat)
                    self.dev_state[2] = 0.0
        else:  # a plain torch.optim.Adam checkpoint (the reference's Trainable): step counter from the entries, rest fresh
            self.dev_state.zero_()
            self.dev_state[0] = steps[0] if steps else 0.0
            if self.ema_shadow is not None:
                self.ema_shadow.copy_(self.flat)

    def load_ema_state_dict(self, sd: dict):
        """A `torch_ema.ExponentialMovingAverage.state_dict()` (the reference's checkpoint entry `ema`): `shadow_params`
        (one tensor per parameter, in parameter order) and `num_updates`."""
        if self.ema_shadow is None:
            raise ValueError("this optimizer was built without ema=...")
        shadow = sd["shadow_params"]
        if len(shadow) != len(self._params):
            raise ValueError(f"EMA checkpoint holds {len(shadow)} tensors, the optimizer {len(self._params)} parameters")
        for (i, p, a, b), t in zip(self._segments(), shadow):
            if tuple(t.shape) != tuple(p.shape):
                raise ValueError(f"shadow_params[{i}] has shape {tuple(t.shape)}, parameter has {tuple(p.shape)}")
            self.ema_shadow[a:b].copy_(t.reshape(-1).to(self.ema_shadow.dtype))
        self.dev_state[2] = float(sd.get("num_updates") or 0)
        if sd.get("decay") is not None:
            self.ema["decay"] = float(sd["decay"])


_PRIOR_CALLS = [0]


def sample_gauss_prior(batch: int, dim: int, *, mean: float = 0.0, std: float = 1.0, truncate: tuple | None = None,
                       seed: int | None = None, traj_offset: int = 0, device="cuda", uniforms: torch.Tensor | None = None) -> torch.Tensor:
    """x0 ~ IsotropicGauss prior on the device (`sdes_sample_gauss_prior`): `truncate=(a, b)` = the bounds
    `IsotropicGauss.truncate_quartile` holds (distr/gauss.py:206-213).  `seed=None` (default) draws a fresh stream per call
    from torch's seed and a call counter — like `prior.sample`, repeated calls return different samples; pass `seed` for a
    reproducible draw."""
    if seed is None:
        from .losses import mix_key

        seed = mix_key(torch.initial_seed(), 0x9817, _PRIOR_CALLS[0])
        _PRIOR_CALLS[0] += 1
    device = torch.device(device)
    if device.type != "cuda":
        raise _cabi.SdesError("prior sampling runs on a CUDA device only; there is no CPU path")
    out = torch.empty((batch, dim), dtype=torch.float32, device=device)
    a, b = (0.0, 0.0) if truncate is None else (float(truncate[0]), float(truncate[1]))
    u = None
    if uniforms is not None:
        u = uniforms.to(device=device, dtype=torch.float32).contiguous()
        if tuple(u.shape) != (batch, dim):
            raise ValueError("uniforms must be (batch, dim)")
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _cabi.check(_cabi.lib().sdes_sample_gauss_prior(out.data_ptr(), batch, dim, float(mean), float(std), int(truncate is not None),
                                                        a, b, seed & 0xFFFFFFFFFFFFFFFF, traj_offset,
                                                        None if u is None else u.data_ptr(), C.c_void_p(stream)),
                    "sdes_sample_gauss_prior")
    return out


def eval_moments(samples: torch.Tensor, weights: torch.Tensor | None = None) -> dict:
    """ESS and per-dimension mean / stddev of get_metrics (eval/metrics.py:120-131) from one kernel pass and one host read."""
    x = samples.detach().to(torch.float32).contiguous()
    if not x.is_cuda:
        raise _cabi.SdesError("eval moments run on a CUDA device only; there is no CPU path")
    B, d = x.shape
    w = None if weights is None else weights.detach().reshape(-1).to(torch.float32).contiguous()
    out = torch.empty(4 + 2 * d, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _cabi.check(_cabi.lib().sdes_eval_moments(x.data_ptr(), None if w is None else w.data_ptr(), B, d, out.data_ptr(),
                                                  C.c_void_p(stream)), "sdes_eval_moments")
    h = out.cpu()
    s1, s2 = h[4:4 + d], h[4 + d:4 + 2 * d]
    mean = s1 / B
    std = ((s2 - s1 * s1 / B) / max(B - 1, 1)).clamp_min(0).sqrt()  # samples.std(dim=0): unbiased
    m = {"eval/avg_stddev": float(std.mean()), "means": mean, "stddevs": std}
    if w is not None:
        ess = float(h[0] ** 2 / h[1])
        m["eval/effective_sample_size"] = ess
        m["eval/norm_effective_sample_size"] = ess / B
    return m
