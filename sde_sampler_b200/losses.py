"""Drop-in replacements for the reference's optimal-control losses (sde_sampler/losses/oc.py).

    loss._target_: sde_sampler_b200.FusedTimeReversalLoss            (conf/loss/time_reversal*.yaml:2)
                   sde_sampler_b200.FusedReferenceSDELoss            (conf/loss/reference_sde*.yaml:2)
                   sde_sampler_b200.FusedExponentialIntegratorSDELoss (conf/loss/exponential_sde*.yaml:2)

Same constructor arguments, same `__call__` / `eval` / `simulate` / `compute_loss` /
`compute_results` / `state_dict` surface and error behaviour as `BaseOCLoss` and its three
subclasses (losses/oc.py:13-137, :139-278, :281-392, :395-483) — but `simulate` is ONE call
into the CUDA library instead of a Python loop over time steps, and the reductions run in a
CUDA kernel whose partial statistics are combined across ranks when a process group is given.

What is not implemented raises (`NotImplementedError`) instead of falling back: a learned
`inference_ctrl`, `sde_ctrl_noise` / `sde_ctrl_dropout`, targets other than GMM / Gauss /
DoubleWell / MultiWell / Funnel / Nice.  Gradients: `loss.method = "lv"` with trainable control
parameters returns a loss whose `.backward()` runs on the tensor cores (autograd.py, csrc/sdes_grad.cu); `"kl"` /
`"kl_ito"` do the same after a reverse sweep over the stored trajectory (csrc/sdes_adjoint.cu; d <= 64, analytic
targets — on the wide engine the kl losses return a plain value).
"""
from __future__ import annotations

import logging
from collections import namedtuple
from typing import Callable

import torch

from . import _cabi, engine
from .autograd import wants_grad
from .dist import combine_stats
from .spec import ctrl_parameters, extract_spec

_STREAM_IDS = [0]


def mix_key(seed: int, stream: int, call: int) -> int:
    """splitmix64-style mix of (seed, stream id, call counter) into one 64-bit Philox key."""
    m = 0xFFFFFFFFFFFFFFFF
    z = (seed & m) ^ ((stream & m) * 0x9E3779B97F4A7C15 & m) ^ ((call & m) * 0xBF58476D1CE4E5B9 & m)
    for _ in range(2):
        z = (z + 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        z ^= z >> 31
    return z


# utils/common.py:9-13
Results = namedtuple(
    "Results",
    "samples weights log_norm_const_preds expectation_preds ts xs metrics plots",
    defaults=[{}, {}, None, None, None, None, {}, {}],
)


class FusedOCLoss:
    """BaseOCLoss (losses/oc.py:13-137) on the fused kernel."""

    loss_kind: str = ""

    def __init__(
        self,
        generative_ctrl: Callable,
        sde=None,
        method: str = "kl",
        traj_per_sample: int = 1,
        filter_samples: Callable | None = None,
        max_rnd: float | None = None,
        sde_ctrl_dropout: float | None = None,
        sde_ctrl_noise: float | None = None,
        *,
        seed: int | None = None,
        engine: str = "auto",
        process_group=None,
        sync_metrics: bool = True,
        **kwargs,
    ):
        self.generative_ctrl = generative_ctrl
        self.sde = sde
        if method not in ["kl", "kl_ito", "lv", "lv_traj"]:
            raise ValueError("Unknown loss method.")
        self.method = method
        if traj_per_sample == 1 and self.method == "lv_traj":
            raise ValueError("Cannot compute variance over a single trajectory.")
        self.traj_per_sample = traj_per_sample
        self.filter_samples = filter_samples
        self.max_rnd = max_rnd
        self.sde_ctrl_noise = sde_ctrl_noise
        self.sde_ctrl_dropout = sde_ctrl_dropout
        if sde_ctrl_noise is not None or sde_ctrl_dropout is not None:
            raise NotImplementedError("sde_ctrl_noise / sde_ctrl_dropout are not implemented in the fused rollout")
        self._n_filtered = 0
        # fused-path extras (keyword-only, absent from the reference signature)
        self.engine = engine
        self.process_group = process_group
        # True (default, the reference's behaviour, oc.py:86): n_filtered is a Python int updated with a host
        # sync every call.  False: the count stays on the device and the metric is returned as a 0-dim tensor,
        # so consecutive calls queue back to back on the stream (read `loss.n_filtered` to materialise it).
        self.sync_metrics = sync_metrics
        self._n_filtered_dev = None
        self._seed = seed
        self._calls = 0
        _STREAM_IDS[0] += 1
        self._stream_id = _STREAM_IDS[0]  # per-object noise stream (see _next_seed)
        self._workspace = engine_workspace()
        self._grad_workspace = engine_workspace()
        self._traj_buffer = engine_workspace()  # trajectory of the last training forward (row-tiled), reused across steps
        self._gate_cot_buffer = engine_workspace()  # (T, B) gate cotangent sums of the last training forward (lv losses)
        self._score_keep_buffer = engine_workspace()  # ungated score part per (step, trajectory, dim) of the last kl / kl_ito forward
        self._traj_version = 0
        self._spec_cache: dict = {}
        _cabi.lib()  # fail now, not at the first step, if the CUDA library is missing

    # ------------------------------------------------------------------ noise stream
    def _next_seed(self) -> int:
        """Philox key for the next rollout: a 64-bit mix of the FULL base seed (torch.initial_seed() unless given), a
        per-object stream id and the number of rollouts this loss has run — every call draws fresh noise (the reference
        advances torch's global generator, losses/oc.py:214), and two loss objects (train / eval, or a loss and an
        integrator) under the same torch seed do not replay each other's stream."""
        base = torch.initial_seed() if self._seed is None else self._seed
        key = mix_key(base, self._stream_id, self._calls)
        self._calls += 1
        return key

    def _rank_offset(self, batch: int) -> int:
        """Global index of this rank's first trajectory: ranks hold contiguous shards of the
        global batch, so the noise a trajectory sees does not depend on the number of GPUs."""
        pg = self.process_group
        if pg is None:
            return 0
        import torch.distributed as dist

        return dist.get_rank(pg) * batch

    # ---------------------------------------------------------------------- rollout
    def _spec(self, ts, terminal_unnorm_log_prob, second_log_prob, *, train, compute_ito, return_traj):
        """`extract_spec` with its object introspection cached: the spec only holds REFERENCES to the caller's live
        tensors (values are re-read by the kernels every call), so it stays valid as long as the same objects and the
        same parameter tensors are handed in; the scalars a scheduler may change between calls (clip values, ts) are
        refreshed here on every call.  Distribution / SDE buffers and the Python scalars read from them are part of the
        cache key (`_buffer_ptrs`: pointers, in-place version counters, scalar values), because the spec may hold converted
        copies of those."""
        ctrl = self.generative_ctrl
        live = ctrl_parameters(ctrl)
        key = (train, compute_ito, return_traj, int(ts.shape[0]), id(ctrl), id(self.sde),
               id(getattr(terminal_unnorm_log_prob, "__self__", terminal_unnorm_log_prob)),
               id(getattr(second_log_prob, "__self__", second_log_prob)), self.method, self.traj_per_sample)
        # storage fingerprint: parameters / distribution buffers that were re-allocated (Module.to, p.data = ...) miss
        fp = tuple(p.data_ptr() for p in live) + _buffer_ptrs(getattr(terminal_unnorm_log_prob, "__self__", None)) + \
            _buffer_ptrs(getattr(second_log_prob, "__self__", None)) + _buffer_ptrs(self.sde)
        hit = self._spec_cache.get(key)
        if hit is not None and hit[1] == fp:
            spec = hit[0]
            spec.ts = ts.detach().reshape(-1).to(torch.float32)
            spec.ctrl["clip_model"] = ctrl.clip_model
            if spec.ctrl["kind"] != "clipped":
                spec.ctrl["scale_score"] = float(ctrl.scale_score)
                spec.ctrl["clip_score"] = ctrl.clip_score
                spec.extras["detach_score"] = bool(getattr(ctrl, "detach_score", False))
            owner = getattr(terminal_unnorm_log_prob, "__self__", None)
            if hasattr(owner, "clip_target"):
                spec.target["clip_target"] = owner.clip_target
            spec.loss["max_rnd"] = self.max_rnd
            if self.loss_kind == "exp_integrator":
                spec.loss["alpha"], spec.loss["sigma"] = float(self.alpha), float(self.sigma)
            return spec
        spec = extract_spec(self, self.loss_kind, ts, terminal_unnorm_log_prob, second_log_prob,
                            train=train, compute_ito=compute_ito, return_traj=return_traj)
        if len(self._spec_cache) > 16:
            self._spec_cache.clear()
        self._spec_cache[key] = (spec, fp)
        return spec

    def _simulate(self, ts, x, terminal_unnorm_log_prob, second_log_prob, *, train, compute_ito_int,
                  return_traj, noise=None):
        spec = self._spec(ts, terminal_unnorm_log_prob, second_log_prob, train=train, compute_ito=compute_ito_int,
                          return_traj=return_traj)
        x_T, rnd, xs = engine.rollout(spec, x, noise=noise, seed=self._next_seed(),
                                      traj_offset=self._rank_offset(x.shape[0]), engine=self.engine,
                                      workspace=self._workspace)
        if engine.is_wide(spec):
            # the wide engine's gradient reads what a training forward kept INSIDE this workspace: any later rollout
            # through it invalidates a pending backward (LvLoss.backward checks the version)
            self._traj_version += 1
        return x_T, rnd, xs

    # ------------------------------------------------------------------- reductions
    def filter(self, rnd: torch.Tensor, samples: torch.Tensor | None = None) -> torch.Tensor:
        """losses/oc.py:50-58 (kept for API parity; compute_loss uses the stats kernel)."""
        mask = True
        if samples is not None and self.filter_samples is not None:
            mask = self.filter_samples(samples)
        if self.max_rnd is None:
            return mask & rnd.isfinite()
        return mask & (rnd < self.max_rnd)

    def _stats(self, rnd, samples=None, mask_mode=None):
        smask = None
        if samples is not None and self.filter_samples is not None:
            smask = self.filter_samples(samples)
        if mask_mode is None:
            mask_mode = _cabi.MASK_ISFINITE if self.max_rnd is None else _cabi.MASK_MAX_RND
        st = engine.rnd_stats(rnd, mask_mode, 0.0 if self.max_rnd is None else self.max_rnd, smask)
        return combine_stats(st, self.process_group)

    def compute_loss(self, rnd: torch.Tensor, samples: torch.Tensor | None = None) -> tuple[torch.Tensor, dict]:
        """losses/oc.py:72-92.  lv: unbiased variance of the kept rnd; kl: their mean."""
        if self.method == "lv_traj":
            return self._compute_loss_lv_traj(rnd, samples)
        return self._loss_from_stats(self._stats(rnd, samples))

    def _loss_from_stats(self, st) -> tuple[torch.Tensor, dict]:
        # st[6] / st[7]: lv / kl loss of the kept entries, computed by the statistics kernel (or by merge_stats)
        if self.sync_metrics:
            self.n_filtered += int((st[5] - st[0]).item())  # the reference syncs here too (.item(), oc.py:86)
            count = self.n_filtered
        else:
            dropped = st[5] - st[0]
            self._n_filtered_dev = dropped if self._n_filtered_dev is None else self._n_filtered_dev + dropped
            count = self._n_filtered_dev + self._n_filtered
        loss = st[6] if self.method == "lv" else st[7]
        return loss.to(torch.float32), {"train/n_filtered_cumulative": count}

    def _call_with_grad(self, ts, x, terminal_unnorm_log_prob, second_log_prob, noise=None):
        """Training call whose result carries a gradient: value by the fused rollout (trajectory kept), gradient by
        `sdes_rollout_lv_grad` (lv) or `sdes_rollout_kl_grad` (kl / kl_ito: backpropagation through time as a discrete
        adjoint) — sde_sampler_b200/autograd.py."""
        from .autograd import LvLoss

        params = ctrl_parameters(self.generative_ctrl)
        net = self.generative_ctrl.base_model
        gate = getattr(self.generative_ctrl, "score_model", None)
        out = {}

        def run():
            spec = self._spec(ts, terminal_unnorm_log_prob, second_log_prob, train=True, compute_ito=self.method != "kl",
                              return_traj=True)
            seed, off = self._next_seed(), self._rank_offset(x.shape[0])
            extra = {}
            x_T, rnd, xs = engine.rollout(spec, x, noise=noise, seed=seed, traj_offset=off, engine=self.engine,
                                          workspace=self._workspace, traj_tiled=True, traj_buffer=self._traj_buffer,
                                          keep_for_grad=True, keep_score=self.method in ("kl", "kl_ito"),
                                          gate_cot=self._gate_cot_buffer if self.method in ("lv", "lv_traj") else None, out=extra,
                                          score_keep=self._score_keep_buffer if self.method in ("kl", "kl_ito") else None)
            self._traj_version += 1
            if self.method == "lv_traj":
                loss, metrics = self._compute_loss_lv_traj(rnd, x_T)
                st = self._lv_traj_stats
            else:
                st = self._stats(rnd, x_T)
                loss, metrics = self._loss_from_stats(st)
            smask = None if self.filter_samples is None else self.filter_samples(x_T)
            out.update(loss=loss, metrics=metrics, stats=st, rnd=rnd, smask=smask, xs=xs, spec=spec, seed=seed,
                       traj_offset=off, noise=noise, samples=x_T, traj_version=self._traj_version, gate_cot=extra.get("gate_cot"),
                       score_keep=extra.get("score_keep"))
            return out

        loss = LvLoss.apply(self, run, len(net.timestep_embed.hidden_layer), len(net.hidden_layer),
                            0 if gate is None else len(gate.hidden_layer), *params)
        return loss, out["metrics"]

    @property
    def n_filtered(self) -> int:
        if self._n_filtered_dev is not None:
            self._n_filtered += int(self._n_filtered_dev.item())
            self._n_filtered_dev = None
        return self._n_filtered

    @n_filtered.setter
    def n_filtered(self, value: int):
        self._n_filtered = int(value)
        self._n_filtered_dev = None

    def _compute_loss_lv_traj(self, rnd, samples):
        # variance over the traj_per_sample copies of each x0 (oc.py:78-84), reduced by `sdes_lv_traj_stats`; every rank
        # holds all copies of its own samples, so ranks only add three numbers
        smask = None
        if samples is not None and self.filter_samples is not None:
            smask = self.filter_samples(samples)
        mode = _cabi.MASK_ISFINITE if self.max_rnd is None else _cabi.MASK_MAX_RND
        st = engine.lv_traj_stats(rnd, self.traj_per_sample, mode, 0.0 if self.max_rnd is None else self.max_rnd, smask)
        if self.process_group is not None:
            import torch.distributed as dist

            dist.all_reduce(st, group=self.process_group)
        self._lv_traj_stats = st  # (rank-combined) [sum of variances, kept samples, samples]: the backward's weights need [1]
        self.n_filtered += self.traj_per_sample * int((st[2] - st[1]).item())
        return (st[0] / st[1]).to(torch.float32), {"train/n_filtered_cumulative": self.n_filtered}

    def compute_results(self, rnd: torch.Tensor, compute_weights: bool = False, ts=None, samples=None, xs=None):
        """losses/oc.py:94-123 (a staticmethod there; an instance method here because the
        statistics may span ranks — `FusedOCLoss.compute_results_local` is the static form)."""
        return _compute_results(rnd, compute_weights, ts, samples, xs, self.process_group)

    @staticmethod
    def compute_results_local(rnd, compute_weights=False, ts=None, samples=None, xs=None):
        return _compute_results(rnd, compute_weights, ts, samples, xs, None)

    def __call__(self, ts: torch.Tensor, x: torch.Tensor, *args, **kwargs) -> tuple[torch.Tensor, dict]:
        raise NotImplementedError

    def eval(self, ts: torch.Tensor, x: torch.Tensor, *args, **kwargs) -> Results:
        raise NotImplementedError

    def load_state_dict(self, state_dict: dict):
        self.n_filtered = state_dict["n_filtered"]
        self._calls = int(state_dict.get("noise_calls", self._calls))  # absent in the reference's checkpoints

    def state_dict(self) -> dict:
        # `n_filtered` is the reference's key (losses/oc.py:133-137); `noise_calls` resumes the Philox stream where it stopped
        return {"n_filtered": self.n_filtered, "noise_calls": self._calls}

    def _repeat(self, x):
        if self.traj_per_sample != 1:
            x = x.repeat(self.traj_per_sample, 1, 1).reshape(-1, x.shape[-1])
        return x


_SCALAR_ATTRS = ("variance", "separation", "shift", "n_double_wells", "log_norm_const", "clip_target", "dim", "sign", "generative",
                 "truncate_quartile")


def _buffer_ptrs(obj) -> tuple:
    """Fingerprint of what an introspected owner (solver shim -> target, prior, reference, sde) holds: data pointers AND
    in-place version counters of its buffers (the spec may hold fp32 / broadcast COPIES of them), plus the Python scalars
    the extraction reads — so an in-place edit or a changed scalar re-extracts instead of being silently ignored."""
    if obj is None:
        return ()
    out = []
    for o in (obj, getattr(obj, "target", None), getattr(obj, "prior", None)):
        if o is None:
            continue
        if isinstance(o, torch.nn.Module):
            out += [(t.data_ptr(), t._version) for t in o.buffers()]
        for name in _SCALAR_ATTRS:
            v = getattr(o, name, None)
            if isinstance(v, (int, float, bool)):
                out.append((name, v))
            elif isinstance(v, (list, tuple)) and all(isinstance(e, (int, float)) for e in v):
                out.append((name, tuple(v)))
    return tuple(out)


def engine_workspace():
    return engine.Workspace()


def _compute_results(rnd, compute_weights, ts, samples, xs, process_group) -> Results:
    metrics = {}
    st = combine_stats(engine.rnd_stats(rnd, _cabi.MASK_ALL), process_group)
    host = st.tolist()  # one sync; the reference does three .item() calls here
    n, s1, s2, mx, se = host[0], host[1], host[2], host[3], host[4]
    neg_mean = -s1 / n
    if compute_weights:
        weights = engine.importance_weights(rnd, st)
        import math

        log_norm_const_preds = {
            "log_norm_const_lb_ito": neg_mean,
            "log_norm_const_is": math.log(se / n) + mx if se > 0 else float("nan") if se != se else -math.inf,
        }
        metrics["eval/lv_loss"] = (s2 - s1 * s1 / n) / (n - 1.0) if n > 1 else float("nan")
    else:
        weights = None
        log_norm_const_preds = {"log_norm_const_lb": neg_mean}
    return Results(samples=samples, weights=weights, log_norm_const_preds=log_norm_const_preds, ts=ts, xs=xs,
                   metrics=metrics)


class FusedTimeReversalLoss(FusedOCLoss):
    """TimeReversalLoss (losses/oc.py:139-278): DIS / Bridge without a learned inference control."""

    loss_kind = "time_reversal"

    def __init__(self, *args, inference_ctrl: Callable | None = None, div_estimator: str | None = None, **kwargs):
        super().__init__(*args, **kwargs)
        if inference_ctrl is not None:
            raise NotImplementedError(
                "a learned inference_ctrl needs div_x of a network (utils/autograd.py) — not on the fused path")
        self.inference_ctrl = inference_ctrl
        self.div_estimator = div_estimator
        if self.div_estimator is not None and self.inference_ctrl is None:
            logging.warning("Without inference control the divergence estimator has no effect.")

    def simulate(self, ts, x, terminal_unnorm_log_prob, initial_log_prob=None, train=True, compute_ito_int=False,
                 change_sde_ctrl=False, return_traj=False, noise=None):
        # change_sde_ctrl only changes what the autograd graph sees (sde_ctrl = g.detach()), not values
        return self._simulate(ts, x, terminal_unnorm_log_prob, initial_log_prob, train=train,
                              compute_ito_int=compute_ito_int, return_traj=return_traj, noise=noise)

    def __call__(self, ts, x, terminal_unnorm_log_prob, initial_log_prob, *, noise=None):
        x = self._repeat(x)
        if wants_grad(self):
            return self._call_with_grad(ts, x, terminal_unnorm_log_prob, initial_log_prob, noise=noise)
        samples, rnd, _ = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, initial_log_prob=initial_log_prob,
            compute_ito_int=self.method != "kl", change_sde_ctrl=self.method in ["lv", "lv_traj"],
            train=True, return_traj=False, noise=noise)
        return self.compute_loss(rnd, samples=samples)

    def eval(self, ts, x, terminal_unnorm_log_prob, initial_log_prob=None, compute_weights=True, return_traj=True):
        samples, rnd, xs = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, initial_log_prob=initial_log_prob,
            compute_ito_int=compute_weights, train=False, return_traj=return_traj)
        return self.compute_results(rnd, compute_weights=compute_weights, ts=ts, samples=samples, xs=xs)


class FusedReferenceSDELoss(FusedOCLoss):
    """ReferenceSDELoss (losses/oc.py:281-392): PIS and Euler-DDS."""

    loss_kind = "reference_sde"

    def __init__(self, *args, reference_ctrl: Callable | None = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.reference_ctrl = reference_ctrl

    def simulate(self, ts, x, terminal_unnorm_log_prob, reference_log_prob, compute_ito_int=False,
                 change_sde_ctrl=False, return_traj=False, noise=None):
        return self._simulate(ts, x, terminal_unnorm_log_prob, reference_log_prob, train=True,
                              compute_ito_int=compute_ito_int, return_traj=return_traj, noise=noise)

    def __call__(self, ts, x, terminal_unnorm_log_prob, reference_log_prob, *, noise=None):
        x = self._repeat(x)
        if wants_grad(self):
            return self._call_with_grad(ts, x, terminal_unnorm_log_prob, reference_log_prob, noise=noise)
        samples, rnd, _ = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, reference_log_prob=reference_log_prob,
            compute_ito_int=self.method != "kl", change_sde_ctrl=self.method in ["lv", "lv_traj"],
            return_traj=False, noise=noise)
        return self.compute_loss(rnd, samples=samples)

    def eval(self, ts, x, terminal_unnorm_log_prob, reference_log_prob=None, compute_weights=True, return_traj=True):
        samples, rnd, xs = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, reference_log_prob=reference_log_prob,
            compute_ito_int=compute_weights, change_sde_ctrl=False, return_traj=return_traj)
        return self.compute_results(rnd, compute_weights=compute_weights, ts=ts, samples=samples, xs=xs)


class FusedExponentialIntegratorSDELoss(FusedOCLoss):
    """ExponentialIntegratorSDELoss (losses/oc.py:395-483): DDS."""

    loss_kind = "exp_integrator"

    def __init__(self, *args, alpha: float, sigma: float, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = alpha
        self.sigma = sigma

    def simulate(self, ts, x, terminal_unnorm_log_prob, reference_log_prob, compute_ito_int=False,
                 change_sde_ctrl=False, return_traj=False, noise=None):
        return self._simulate(ts, x, terminal_unnorm_log_prob, reference_log_prob, train=True,
                              compute_ito_int=compute_ito_int, return_traj=return_traj, noise=noise)

    def __call__(self, ts, x, terminal_unnorm_log_prob, reference_log_prob, *, noise=None):
        x = self._repeat(x)
        if wants_grad(self):
            return self._call_with_grad(ts, x, terminal_unnorm_log_prob, reference_log_prob, noise=noise)
        samples, rnd, _ = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, reference_log_prob=reference_log_prob,
            compute_ito_int=self.method != "kl", change_sde_ctrl=self.method in ["lv", "lv_traj"],
            return_traj=False, noise=noise)
        return self.compute_loss(rnd, samples=samples)

    def eval(self, ts, x, terminal_unnorm_log_prob, reference_log_prob=None, compute_weights=True, return_traj=True):
        samples, rnd, xs = self.simulate(
            ts, x, terminal_unnorm_log_prob=terminal_unnorm_log_prob, reference_log_prob=reference_log_prob,
            compute_ito_int=compute_weights, change_sde_ctrl=False, return_traj=return_traj)
        return self.compute_results(rnd, compute_weights=compute_weights, ts=ts, samples=samples, xs=xs)
