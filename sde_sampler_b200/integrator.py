"""Drop-in for `sde_sampler.eq.integrator.EulerIntegrator` (eq/integrator.py:79-127) on the path the reference's
`LangevinSolver.run` takes (solver/langevin.py:34-63): Euler–Maruyama on a `LangevinSDE` (eq/sdes.py:38-65), SURVEY §8f-3.

    integrator._target_: sde_sampler_b200.FusedEulerIntegrator          (conf/integrator/euler.yaml:2)

`integrate(sde, ts, x_init, timesteps=None, bm=None)` keeps the reference's signature and returns xs (len(ts), B, d); the
whole chain (10 000 steps in conf/solver/langevin.yaml) is ONE kernel launch with the state in registers, the analytic
target score in place and Philox noise.  Other SDE classes raise `NotImplementedError` — the controlled SDEs of the
samplers are served by the fused losses' `simulate(..., return_traj=True)`."""
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
        self._workspace = engine.Workspace()
        _cabi.lib()

    def integrate(self, sde, ts: torch.Tensor, x_init: torch.Tensor, timesteps: torch.Tensor | None = None, bm=None,
                  *, noise: torch.Tensor | None = None) -> torch.Tensor:
        if bm is not None:
            raise NotImplementedError("a torchsde Brownian path (bm) cannot be consumed by the fused integrator")
        if _cls(sde) != "LangevinSDE":
            raise NotImplementedError(f"FusedEulerIntegrator integrates LangevinSDE only (got {_cls(sde)})")
        if not x_init.is_cuda:
            raise _cabi.SdesError("the fused integrator runs on a CUDA device only; there is no CPU path")
        lib = _cabi.lib()
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
        base = torch.initial_seed() if self._seed is None else self._seed
        d.seed = (((self._calls & 0xFFFFFFFF) << 32) | (base & 0xFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
        self._calls += 1
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
