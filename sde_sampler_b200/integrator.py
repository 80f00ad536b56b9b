"""Drop-in for `sde_sampler.eq.integrator.EulerIntegrator` (eq/integrator.py:79-127), SURVEY §8f-3.

    integrator._target_: sde_sampler_b200.FusedEulerIntegrator          (conf/integrator/euler.yaml:2)

`integrate(sde, ts, x_init, timesteps=None, bm=None)` keeps the reference's signature and returns xs (len(ts), B, d); the
whole chain is ONE kernel launch with Philox noise (or the caller's `bm` increments).  Served SDE classes:

* `LangevinSDE` (eq/sdes.py:38-65) — the unadjusted Langevin sampler of `LangevinSolver.run` (solver/langevin.py:34-63;
  10 000 steps in conf/solver/langevin.yaml): state in registers, analytic target score in place; `expectations()` gives
  the burn-in expectation estimates of `run` (:50-54) in one reduction launch;
* the OU family `VP / ConstOU / ScaledBM` (eq/sdes.py:66-269), generative or not, and `ControlledSDE` (:272-305) over one
  of them with no control or with the Gaussian-marginal score control of PIS (solver/oc.py:204-208) — the inference
  processes `TrainableDiff.compute_results` integrates with `timesteps=ts` (solver/oc.py:100-110).  The x-independent
  coefficient functions of the caller's own SDE object are evaluated once on the whole grid.

A `ControlledSDE` around a learned network control is the rollout of the fused losses (`simulate(..., return_traj=True)`);
anything else raises `NotImplementedError`."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi, engine
from .spec import _cls, _f, _owner, _target_params


class FusedEulerIntegrator:
    def __init__(self, dt: float | None = 0.01, steps: int | None = None, rescale_t: str | None = None, eps: float = 1e-8,
                 *, seed: int | None = None):
        self.dt, self.steps, self.rescale_t, self.eps = dt, steps, rescale_t, eps
        self._seed = seed
        self._calls = 0
        from .losses import _STREAM_IDS

        _STREAM_IDS[0] += 1
        self._stream_id = _STREAM_IDS[0]  # per-object noise stream: a loss and an integrator under one torch seed do not share noise
        self._workspace = engine.Workspace()
        _cabi.lib()

    _OU = ("VP", "ConstOU", "ScaledBM")

    def integrate(self, sde, ts: torch.Tensor, x_init: torch.Tensor, timesteps: torch.Tensor | None = None, bm=None,
                  *, noise: torch.Tensor | None = None) -> torch.Tensor:
        kind = _cls(sde)
        if kind != "LangevinSDE" and kind not in self._OU and kind != "ControlledSDE":
            raise NotImplementedError(f"FusedEulerIntegrator integrates LangevinSDE, VP / ConstOU / ScaledBM and ControlledSDE over "
                                      f"them (got {kind})")
        if not x_init.is_cuda:
            raise _cabi.SdesError("the fused integrator runs on a CUDA device only; there is no CPU path")
        device = x_init.device
        B, dim = x_init.shape
        if timesteps is None:
            # eq/integrator.py:105-113 builds the integration grid with the reference's own get_timesteps; the drop-in
            # runs next to the reference package, so that function is used as is (bit-identical grid).  Without the
            # reference importable the caller passes `timesteps`.
            try:
                from sde_sampler.utils.common import get_timesteps
            except ImportError as exc:
                raise ValueError("FusedEulerIntegrator.integrate needs `timesteps` when sde_sampler.utils.common.get_timesteps "
                                 "is not importable") from exc
            timesteps = get_timesteps(ts[0], ts[-1], dt=self.dt, steps=self.steps, rescale_t=self.rescale_t, device=device)
        timesteps = timesteps.to(device=device, dtype=torch.float32).contiguous()
        out_ts = ts.to(device=device, dtype=torch.float32).contiguous()
        increments = False
        if bm is not None:
            # eq/integrator.py:116-119: `noise = bm(s, t)` per step.  The Brownian path is the caller's object; its increments
            # are evaluated here (one call per step, as the reference does) and handed to the kernel as they are.
            if noise is not None:
                raise ValueError("pass either `bm` or `noise`")
            noise = torch.stack([bm(s, t) for s, t in zip(timesteps[:-1], timesteps[1:])])
            increments = True
        if kind != "LangevinSDE":
            return self._integrate_affine(sde, kind, timesteps, out_ts, x_init, noise, increments)
        lib = _cabi.lib()
        tg = _target_params(_owner(sde.target_score, "target_score"), dim)
        if tg["kind"] == "nice":
            raise NotImplementedError("Langevin dynamics on a NICE target")
        d = _cabi.new_desc()
        keep = []
        d.dim, d.batch = dim, B
        d.log_norm_const = float(tg.get("log_norm_const", 0.0) or 0.0)
        if tg["kind"] in ("gmm", "gauss"):
            d.target_kind = _cabi.TARGET_GMM
            d.n_components = int(tg["loc"].shape[0]) if tg["kind"] == "gmm" else 1
            loc, scale = tg["loc"].to(device), tg["scale"].to(device)
            w = tg.get("weights")
            d.gmm_loc, d.gmm_scale = loc.data_ptr(), scale.data_ptr()
            d.gmm_weights = None if w is None else w.to(device).data_ptr()
            keep += [loc, scale, w]
        elif tg["kind"] == "multiwell":
            d.target_kind = _cabi.TARGET_MULTIWELL
            d.n_double_wells, d.separation, d.shift = int(tg["n_dw"]), float(tg["separation"]), float(tg["shift"])
        else:
            d.target_kind = _cabi.TARGET_FUNNEL
            d.variance = float(tg["variance"])
        d.seed = self._next_seed()
        n_steps = int(timesteps.shape[0]) - 1
        if noise is not None:
            if tuple(noise.shape) != (n_steps, B, dim):
                raise ValueError(f"noise must be {(n_steps, B, dim)}")
            noise = noise.to(device=device, dtype=torch.float32).contiguous()
            d.noise = noise.data_ptr()
            d.flags |= _cabi.F_NOISE_FROM_HBM
        g = _cabi.IntegrateDesc()
        g.struct_bytes = C.sizeof(_cabi.IntegrateDesc)
        g.n_steps, g.n_out = n_steps, int(out_ts.shape[0])
        g.diff_coeff = _f(sde.diff_coeff)
        g.clip_score = math.inf if sde.clip_score is None else float(sde.clip_score)
        g.eps = float(self.eps)
        g.noise_is_increment = int(increments)
        x0 = x_init.detach().to(torch.float32).contiguous()
        xs = torch.empty((g.n_out, B, dim), dtype=torch.float32, device=device)
        g.timesteps, g.out_ts, g.x_init, g.xs_out = timesteps.data_ptr(), out_ts.data_ptr(), x0.data_ptr(), xs.data_ptr()
        with torch.cuda.device(device):
            need = lib.sdes_integrate_workspace_bytes(C.byref(d))
            if need == 0:
                raise _cabi.SdesError("integrator: " + lib.sdes_last_error().decode())
            wsbuf = self._workspace.get(need, device)
            d.workspace, d.workspace_bytes = wsbuf.data_ptr(), wsbuf.numel()
            stream = torch.cuda.current_stream(device).cuda_stream
            _cabi.check(lib.sdes_langevin_integrate(C.byref(d), C.byref(g), C.c_void_p(stream)), "sdes_langevin_integrate")
        return xs

    def _next_seed(self) -> int:
        from .losses import mix_key

        seed = mix_key(torch.initial_seed() if self._seed is None else self._seed, self._stream_id, self._calls)
        self._calls += 1
        return seed

    def _integrate_affine(self, sde, kind, timesteps, out_ts, x_init, noise, increments):
        """OU family / ControlledSDE (module docstring).  Per-step coefficient table from the SDE object's own functions."""
        lib = _cabi.lib()
        device = x_init.device
        B, dim = x_init.shape
        base, ctrl = (sde.sde, sde.ctrl) if kind == "ControlledSDE" else (sde, None)
        if _cls(base) not in self._OU:
            raise NotImplementedError(f"ControlledSDE over {_cls(base)} (VP / ConstOU / ScaledBM only)")
        n_steps = int(timesteps.shape[0]) - 1
        s = timesteps[:-1]
        one = torch.ones(n_steps, device=device, dtype=torch.float32)
        tab = torch.zeros((n_steps, 8), device=device, dtype=torch.float32)
        tab[:, 0] = base.drift_coeff_t(s) * one   # OU.drift = drift_coeff_t(t) x   (eq/sdes.py:94-95)
        tab[:, 1] = base.diff_coeff_t(s) * one    # OU.diff                          (eq/sdes.py:98-99)
        cloc, cmax = None, math.inf
        if ctrl is not None:
            owner = getattr(ctrl, "__self__", None)
            if owner is None or getattr(ctrl, "__name__", "") != "inference_ctrl" or not hasattr(owner, "sde") or not hasattr(owner, "prior"):
                raise NotImplementedError("ControlledSDE.ctrl must be None or the Gaussian-marginal score control of PIS "
                                          "(solver/oc.py:206-208); a learned control is rolled out by the fused losses' simulate()")
            # ControlledSDE.f_and_g (eq/sdes.py:296-305): the control sees terminal_t - t for a non-generative SDE;
            # PIS.inference_ctrl: sde.diff(t, x) * sde.marginal_distr(t, x_init=prior.loc).score(x).clip(max=1e5)
            tt = s if base.generative else float(sde.terminal_t) - s
            loc, var = owner.sde.marginal_params(tt.view(-1, 1), owner.prior.loc.to(device).view(1, -1))
            tab[:, 2] = owner.sde.diff_coeff_t(tt) * one
            tab[:, 3] = 1.0 / (var.to(torch.float32).reshape(n_steps, -1)[:, 0])
            cloc = (loc.to(torch.float32) * torch.ones((n_steps, dim), device=device)).contiguous()
            cmax = 1e5
        g = _cabi.AffineIntegrateDesc()
        g.struct_bytes = C.sizeof(_cabi.AffineIntegrateDesc)
        g.dim, g.batch, g.n_steps, g.n_out = dim, B, n_steps, int(out_ts.shape[0])
        g.eps, g.cmax = float(self.eps), cmax
        x0 = x_init.detach().to(torch.float32).contiguous()
        xs = torch.empty((g.n_out, B, dim), dtype=torch.float32, device=device)
        g.timesteps, g.out_ts, g.tab, g.x_init, g.xs_out = timesteps.data_ptr(), out_ts.data_ptr(), tab.data_ptr(), x0.data_ptr(), xs.data_ptr()
        g.cloc = None if cloc is None else cloc.data_ptr()
        if noise is not None:
            if tuple(noise.shape) != (n_steps, B, dim):
                raise ValueError(f"noise must be {(n_steps, B, dim)}")
            noise = noise.to(device=device, dtype=torch.float32).contiguous()
            g.noise = noise.data_ptr()
        g.noise_is_increment = int(increments)
        g.seed, g.traj_offset = self._next_seed(), 0
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _cabi.check(lib.sdes_affine_integrate(C.byref(g), C.c_void_p(stream)), "sdes_affine_integrate")
        return xs

    @staticmethod
    def expectations(xs: torch.Tensor, burn_steps: int = 0) -> dict:
        """`expectation_preds` of LangevinSolver.run (solver/langevin.py:50-54): means of EXPECTATION_FNS (distr/base.py:12-17)
        over xs[burn_steps:] flattened to rows, as 0-dim device tensors (one reduction launch, no host sync)."""
        if not xs.is_cuda:
            raise _cabi.SdesError("expectations() runs on a CUDA device only")
        rows = xs[burn_steps:].to(torch.float32).contiguous()
        dim = int(rows.shape[-1])
        out = torch.empty(4, dtype=torch.float64, device=xs.device)
        with torch.cuda.device(xs.device):
            _cabi.check(_cabi.lib().sdes_expectations(rows.data_ptr(), rows.numel() // dim, dim, out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream(xs.device).cuda_stream)), "sdes_expectations")
        return {"square": out[0], "abs": out[1], "sum": out[2], "square_minus_sum": out[3]}
