"""RolloutSpec -> SdesRolloutDesc -> one call of the CUDA library.

PyTorch is plumbing here: device memory for inputs/outputs/workspace, the current stream and
one `torch.cat` that packs the caller's parameters into the flat blob the C ABI expects.
Every arithmetic operation of the rollout happens inside `sdes_rollout_fwd`.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi
from .spec import RolloutSpec

_ENGINES = ("auto", "tcgen05", "simt")


def _inf(v) -> float:
    return math.inf if v is None else float(v)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def pack_params(spec: RolloutSpec) -> torch.Tensor:
    """Flat fp32 parameter blob in the order documented in include/sdes_b200.h (values re-read on every call; the list
    of flat VIEWS of the live parameters is kept with the spec)."""
    views = spec.extras.get("flat_views")
    if views is not None:
        return torch.cat(views)
    m = spec.mlp
    te = m["time_embed"]
    parts = [m["in_w"], m["in_b"], te["phase"]]
    for w, b in te["hidden"]:
        parts += [w, b]
    parts += [te["out_w"], te["out_b"]]
    for w, b in m["hidden"]:
        parts += [w, b]
    parts += [m["out_w"], m["out_b"]]
    if spec.gate is not None:
        g = spec.gate
        parts += [g["phase"]]
        for w, b in g["hidden"]:
            parts += [w, b]
        parts += [g["out_w"], g["out_b"]]
    flat = [p.reshape(-1).to(torch.float32) for p in parts]
    if all(f.data_ptr() == p.data_ptr() for f, p in zip(flat, parts)):  # views, not copies: safe to reuse
        spec.extras["flat_views"] = flat
    return torch.cat(flat)


def pack_nice_params(tg: dict) -> torch.Tensor:
    """Flat fp32 NICE blob (include/sdes_b200.h): per coupling the Linear layers in order, then the log-scale."""
    parts = []
    for c in tg["couplings"]:
        for w, b in c["layers"]:
            parts += [w, b]
    parts.append(tg["scale"])
    return torch.cat([p.reshape(-1).to(torch.float32) for p in parts])


def pack_params_numel(spec: RolloutSpec) -> int:
    m = spec.mlp
    te = m["time_embed"]
    ts = [m["in_w"], m["in_b"], te["phase"], te["out_w"], te["out_b"], m["out_w"], m["out_b"]]
    ts += [t for pair in te["hidden"] for t in pair] + [t for pair in m["hidden"] for t in pair]
    if spec.gate is not None:
        g = spec.gate
        ts += [g["phase"], g["out_w"], g["out_b"]] + [t for pair in g["hidden"] for t in pair]
    return sum(int(t.numel()) for t in ts)


class Workspace:
    """Grow-only device scratch, one per loss object and device (allocations are cached across calls;
    nothing else is)."""

    def __init__(self):
        self.buf: torch.Tensor | None = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        return self.buf


def fill_desc(spec: RolloutSpec, *, batch: int, engine: str = "auto") -> tuple[_cabi.RolloutDesc, list]:
    """Everything of the descriptor except the per-call pointers x0/noise/outputs/workspace.
    Returns (desc, keepalive tensors)."""
    if engine not in _ENGINES:
        raise ValueError(f"engine must be one of {_ENGINES}")
    d = _cabi.new_desc()
    keep = []
    ls, cd, tg = spec.loss, spec.ctrl, spec.target
    d.loss_kind = _cabi.LOSS[ls["kind"]]
    d.ctrl_kind = _cabi.CTRL[cd["kind"]]
    flags = 0
    if ls["kind"] == "time_reversal" and ls["train"] and ls["method"] in ("kl", "kl_ito"):
        flags |= _cabi.F_RND0_ZERO
    if ls["compute_ito"]:
        flags |= _cabi.F_COMPUTE_ITO
    if ls["kind"] == "time_reversal" and not ls["train"]:
        flags |= _cabi.F_SUB_DIV_INT
    if ls.get("return_traj"):
        flags |= _cabi.F_RETURN_TRAJ
    if ls.get("reference_ctrl"):
        flags |= _cabi.F_REFERENCE_CTRL
    if spec.gate is not None:
        flags |= _cabi.F_HAS_GATE
    if engine == "simt":
        flags |= _cabi.F_MLP_SIMT
    want_auto = engine == "auto"
    d.dim = spec.dim
    d.n_steps = int(spec.ts.shape[0]) - 1
    d.n_hidden = len(spec.mlp["hidden"])
    d.te_hidden = len(spec.mlp["time_embed"]["hidden"])
    if spec.gate is not None:
        d.gate_hidden = len(spec.gate["hidden"])
        d.gate_dim = int(spec.gate["out_w"].shape[0])
    d.batch = batch
    d.clip_model = _inf(cd.get("clip_model"))
    d.clip_score = _inf(cd.get("clip_score"))
    d.clip_target = _inf(tg.get("clip_target"))
    d.scale_score = float(cd.get("scale_score", 1.0))
    d.alpha = float(ls.get("alpha", 0.0))
    d.sigma = float(ls.get("sigma", 0.0))
    sde = spec.sde
    if sde is None:
        d.sde_kind = _cabi.SDE_NONE
    elif sde["kind"] == "vp":
        d.sde_kind = _cabi.SDE_VP
        d.beta_min, d.beta_max = sde["beta_min"], sde["beta_max"]
        d.scale_diff, d.terminal_t, d.sde_sign = sde["scale"], sde["terminal_t"], sde["sign"]
    else:
        d.sde_kind = _cabi.SDE_CONST_OU
        d.drift_coeff, d.diff_coeff = sde["drift_coeff"], sde["diff_coeff"]
        d.terminal_t, d.sde_sign = sde["terminal_t"], sde["sign"]
    d.log_norm_const = float(tg.get("log_norm_const", 0.0) or 0.0)
    if tg["kind"] == "gmm":
        d.target_kind = _cabi.TARGET_GMM
        d.n_components = int(tg["loc"].shape[0])
        d.gmm_loc, d.gmm_scale, d.gmm_weights = _ptr(tg["loc"]), _ptr(tg["scale"]), _ptr(tg["weights"])
        keep += [tg["loc"], tg["scale"], tg["weights"]]
    elif tg["kind"] == "gauss":  # a diagonal Gaussian target is the K=1 mixture
        d.target_kind = _cabi.TARGET_GMM
        d.n_components = 1
        d.gmm_loc, d.gmm_scale, d.gmm_weights = _ptr(tg["loc"]), _ptr(tg["scale"]), None
        keep += [tg["loc"], tg["scale"]]
    elif tg["kind"] == "multiwell":
        d.target_kind = _cabi.TARGET_MULTIWELL
        d.n_double_wells = int(tg["n_dw"])
        d.separation, d.shift = float(tg["separation"]), float(tg["shift"])
    elif tg["kind"] == "funnel":
        d.target_kind = _cabi.TARGET_FUNNEL
        d.variance = float(tg["variance"])
    elif tg["kind"] == "nice":
        d.target_kind = _cabi.TARGET_NICE
        cps = tg["couplings"]
        d.nice_couplings = len(cps)
        d.nice_hidden = len(cps[0]["layers"]) - 1
        d.nice_mid = int(cps[0]["layers"][0][0].shape[0])
        d.nice_mask_config = int(cps[0]["mask_config"])
        for i, c in enumerate(cps):
            if int(c["mask_config"]) != (d.nice_mask_config + i) % 2:
                raise NotImplementedError("NICE couplings must alternate their mask (distr/nice.py:146-157)")
        blob = pack_nice_params(tg)
        d.nice_params, d.n_nice_params = blob.data_ptr(), blob.numel()
        keep.append(blob)
    else:
        raise NotImplementedError(tg["kind"])
    if spec.prior is not None:
        d.prior_loc, d.prior_scale = _ptr(spec.prior["loc"]), _ptr(spec.prior["scale"])
        keep += [spec.prior["loc"], spec.prior["scale"]]
    if spec.ref is not None:
        d.ref_loc, d.ref_scale = _ptr(spec.ref["loc"]), _ptr(spec.ref["scale"])
        keep += [spec.ref["loc"], spec.ref["scale"]]
    d.flags = flags
    if want_auto:
        # same hardware, two kernels: the tensor-core engine where its tile shapes apply, else the
        # fp32-FFMA engine.  Decided here, visibly, from the descriptor — never inside the library.
        d.n_params = pack_params_numel(spec)
        if not _cabi.lib().sdes_tcgen05_supported(C.byref(d)):
            d.flags |= _cabi.F_MLP_SIMT
    return d, keep


def _check_device(name: str, t: torch.Tensor | None, device):
    if t is not None and t.device != device:
        raise ValueError(f"{name} lives on {t.device}, the rollout runs on {device}")


def is_wide(spec: RolloutSpec) -> bool:
    """Served by the wide engine (csrc/sdes_wide.cu): state wider than the fused kernels hold, or a NICE target."""
    return spec.dim > _cabi.MAX_DIM or spec.target["kind"] == "nice"


def _mma_pad_dim(d: int) -> int:
    return 8 if d <= 8 else 16 if d <= 16 else 32 if d <= 32 else 48 if d <= 48 else 56 if d <= 56 else 64


def tiled_traj_numel(T: int, B: int, dim: int) -> int:
    """Floats of the row-tiled trajectory layout (SDES_F_TRAJ_TILED, include/sdes_b200.h)."""
    return (T + 1) * ((B + 127) // 128) * _mma_pad_dim(dim) * 128


def rollout(spec: RolloutSpec, x0: torch.Tensor, *, noise: torch.Tensor | None = None, seed: int = 0,
            traj_offset: int = 0, engine: str = "auto", workspace: Workspace | None = None,
            params: torch.Tensor | None = None, traj_tiled: bool = False, traj_buffer: Workspace | None = None,
            keep_for_grad: bool = False, keep_score: bool = False, gate_cot: Workspace | None = None, out: dict | None = None,
            score_keep: Workspace | None = None):
    """One fused rollout on x0's device.  Returns (x_T (B,d), rnd (B,1), xs (T+1,B,d) | None); with `traj_tiled`
    the trajectory comes back as a flat buffer in the row-tiled layout that `lv_grad` consumes."""
    lib = _cabi.lib()
    if not x0.is_cuda:
        raise _cabi.SdesError("the fused rollout runs on a CUDA device only (x is on %s); there is no CPU path" % x0.device)
    if x0.ndim != 2 or x0.shape[1] != spec.dim:
        raise ValueError(f"x must be (B, {spec.dim}), got {tuple(x0.shape)}")
    device = x0.device
    B, dim = x0.shape
    T = int(spec.ts.shape[0]) - 1
    x0c = x0.detach().to(torch.float32).contiguous()
    ts = spec.ts.to(device=device, dtype=torch.float32).contiguous()
    d, keep = fill_desc(spec, batch=B, engine=engine)
    if params is None:
        params = pack_params(spec)
    _check_device("model parameters", params, device)
    for k in keep:
        _check_device("distribution parameters", k, device)
    d.ts, d.params, d.n_params = ts.data_ptr(), params.data_ptr(), params.numel()
    d.seed, d.traj_offset = seed & 0xFFFFFFFFFFFFFFFF, traj_offset
    if keep_for_grad and is_wide(spec):
        # wide engine: what the gradient needs stays inside the workspace (state image per step, gate cotangent sums)
        d.flags |= _cabi.F_KEEP_FOR_GRAD
        if keep_score:  # kl / kl_ito: the reverse sweep also needs the target score of every step (sdes_rollout_kl_grad)
            d.flags |= _cabi.F_KEEP_SCORE
        d.flags &= ~_cabi.F_RETURN_TRAJ
    x_T = torch.empty_like(x0c)
    rnd = torch.empty((B, 1), dtype=torch.float32, device=device)
    xs = None
    if d.flags & _cabi.F_RETURN_TRAJ:
        if traj_tiled and dim <= _cabi.MAX_DIM and spec.target["kind"] != "nice":
            d.flags |= _cabi.F_TRAJ_TILED
            n = tiled_traj_numel(T, B, dim)
            if traj_buffer is not None:  # grow-only buffer owned by the loss: no 1.5 GB allocation per training step
                xs = traj_buffer.get(4 * n, device)[: 4 * n].view(torch.float32)
            else:
                xs = torch.empty(n, dtype=torch.float32, device=device)
        else:
            xs = torch.empty((T + 1, B, dim), dtype=torch.float32, device=device)
        d.xs = xs.data_ptr()
    if noise is not None:
        if tuple(noise.shape) != (T, B, dim):
            raise ValueError(f"noise must be {(T, B, dim)}, got {tuple(noise.shape)}")
        noise = noise.to(device=device, dtype=torch.float32).contiguous()
        d.noise = noise.data_ptr()
        d.flags |= _cabi.F_NOISE_FROM_HBM
    d.x0, d.x_T, d.rnd = x0c.data_ptr(), x_T.data_ptr(), rnd.data_ptr()
    if gate_cot is not None and spec.gate is not None and int(spec.gate["out_w"].shape[0]) == 1 and not is_wide(spec) \
            and not (d.flags & _cabi.F_MLP_SIMT) and spec.ctrl["kind"] != "clipped":
        # training forward on the tensor-core engine: keep d rnd / d gate per (step, trajectory) so that the lv gradient
        # does not have to re-evaluate the target score (SdesRolloutDesc.gate_cot)
        gc = gate_cot.get(4 * T * B, device)[: 4 * T * B].view(torch.float32)
        d.gate_cot = gc.data_ptr()
        if out is not None:
            out["gate_cot"] = gc
    if score_keep is not None and xs is not None and not is_wide(spec) and not (d.flags & _cabi.F_MLP_SIMT) \
            and spec.ctrl["kind"] != "clipped":
        # kl / kl_ito training forward on the tensor-core engine: keep the ungated score part of every (step, trajectory,
        # dimension) in the layout of xs, so that the reverse sweep runs as one kernel (SdesRolloutDesc.score_keep)
        sk = score_keep.get(4 * xs.numel(), device)[: 4 * xs.numel()].view(torch.float32)
        d.score_keep = sk.data_ptr()
        if out is not None:
            out["score_keep"] = sk
    with torch.cuda.device(device):
        need = lib.sdes_workspace_bytes(C.byref(d))
        if need == 0:
            raise _cabi.SdesError("invalid descriptor: " + lib.sdes_last_error().decode())
        wsbuf = (workspace or Workspace()).get(need, device)
        d.workspace, d.workspace_bytes = wsbuf.data_ptr(), wsbuf.numel()
        stream = torch.cuda.current_stream(device).cuda_stream
        _cabi.check(lib.sdes_rollout_fwd(C.byref(d), C.c_void_p(stream)), "sdes_rollout_fwd")
    return x_T, rnd, xs


def kl_grad_flags(spec: RolloutSpec) -> int:
    """SDES_GRAD_* for `sdes_rollout_kl_grad`: which parts of the control's score term carry a graph in the reference.
    A `GMM` target's score is `Distribution.score` — autograd WITHOUT create_graph (distr/base.py:130-137, called with
    create_graph=detach_score=False from models/reparam.py:60,:135) — a constant of the graph; Gauss / DoubleWell /
    MultiWell / Funnel scores are analytic functions of x; `detach_score=True` detaches the whole score term."""
    flags = 0
    if spec.target["kind"] in ("gmm", "nice"):  # Distribution.score (autograd, no create_graph) — Nice inherits it too
        flags |= _cabi.GRAD_TARGET_SCORE_CONST
    if spec.extras.get("detach_score"):
        flags |= _cabi.GRAD_SCORE_DETACHED
    return flags


def kl_grad(spec: RolloutSpec, xs: torch.Tensor, w: torch.Tensor, **kw):
    """d loss / d theta of the kl / kl_ito loss (backpropagation through time, `sdes_rollout_kl_grad`); arguments and
    results as `lv_grad`."""
    return lv_grad(spec, xs, w, bptt=True, **kw)


def lv_grad(spec: RolloutSpec, xs: torch.Tensor, w: torch.Tensor, *, noise: torch.Tensor | None = None, seed: int = 0,
            traj_offset: int = 0, engine: str = "auto", workspace: Workspace | None = None,
            params: torch.Tensor | None = None, chunk_rows: int = 0, bptt: bool = False, grad_flags: int = 0,
            gate_cot: torch.Tensor | None = None, score_keep: torch.Tensor | None = None):
    """d loss / d theta of the log-variance loss for the rollout that produced `xs` (same spec / seed / traj_offset /
    noise).  Returns (grad_params blob, grad_emb (T,64), grad_gate (T,gate_dim) | None) — see include/sdes_b200.h
    `sdes_rollout_lv_grad`.  `bptt=True`: the kl / kl_ito gradient (`sdes_rollout_kl_grad`)."""
    lib = _cabi.lib()
    wide = is_wide(spec)
    fn_bytes, fn_grad, what = ((lib.sdes_kl_grad_workspace_bytes, lib.sdes_rollout_kl_grad, "sdes_rollout_kl_grad") if bptt else
                               (lib.sdes_lv_grad_workspace_bytes, lib.sdes_rollout_lv_grad, "sdes_rollout_lv_grad"))
    if not w.is_cuda:
        raise _cabi.SdesError("the fused gradient runs on a CUDA device only; there is no CPU path")
    device = w.device
    T = int(spec.ts.shape[0]) - 1
    w = w.detach().reshape(-1).to(torch.float32).contiguous()
    B = w.numel()
    tiled = False
    if wide:
        if workspace is None or workspace.buf is None:
            raise ValueError("the wide-engine gradient needs the workspace of the forward call (keep_for_grad=True)")
    else:
        tiled = xs.ndim == 1  # the flat row-tiled buffer of rollout(..., traj_tiled=True)
        if tiled:
            if xs.numel() != tiled_traj_numel(T, B, spec.dim):
                raise ValueError("tiled xs does not match (T, B, dim)")
        elif tuple(xs.shape) != (T + 1, B, spec.dim):
            raise ValueError(f"xs must be {(T + 1, B, spec.dim)}, got {tuple(xs.shape)}")
        xs = xs.detach().to(torch.float32).contiguous()
    ts = spec.ts.to(device=device, dtype=torch.float32).contiguous()
    d, keep = fill_desc(spec, batch=B, engine=engine)
    d.flags &= ~_cabi.F_RETURN_TRAJ
    if tiled:
        d.flags |= _cabi.F_TRAJ_TILED
    if wide:
        d.flags |= _cabi.F_KEEP_FOR_GRAD | (_cabi.F_KEEP_SCORE if bptt else 0)
    if params is None:
        params = pack_params(spec)
    d.ts, d.params, d.n_params = ts.data_ptr(), params.data_ptr(), params.numel()
    d.seed, d.traj_offset = seed & 0xFFFFFFFFFFFFFFFF, traj_offset
    if noise is not None:
        noise = noise.to(device=device, dtype=torch.float32).contiguous()
        d.noise = noise.data_ptr()
        d.flags |= _cabi.F_NOISE_FROM_HBM
    g = _cabi.LvGradDesc()
    g.struct_bytes = C.sizeof(_cabi.LvGradDesc)
    g.flags = (kl_grad_flags(spec) if bptt else 0) | grad_flags
    grad_params = torch.empty_like(params)
    grad_emb = torch.empty((T, _cabi.CHANNELS), dtype=torch.float32, device=device)
    grad_gate = None
    if spec.gate is not None:
        grad_gate = torch.empty((T, int(spec.gate["out_w"].shape[0])), dtype=torch.float32, device=device)
    g.xs, g.w, g.grad_params, g.grad_emb, g.grad_gate = _ptr(None if wide else xs), w.data_ptr(), grad_params.data_ptr(), grad_emb.data_ptr(), _ptr(grad_gate)
    g.chunk_rows = chunk_rows
    if gate_cot is not None and not bptt and not wide:
        if gate_cot.numel() != T * B:
            raise ValueError("gate_cot must be (T, B)")
        g.gate_cot = gate_cot.data_ptr()
    if score_keep is not None and bptt and not wide:
        if score_keep.numel() != xs.numel():
            raise ValueError("score_keep must have the layout (and size) of xs")
        g.score_keep = score_keep.data_ptr()
    with torch.cuda.device(device):
        need = fn_bytes(C.byref(d), C.byref(g))
        if need == 0:
            raise _cabi.SdesError(what + ": " + lib.sdes_last_error().decode())
        if wide:
            wsbuf = workspace.buf  # must be the buffer the keep-mode forward wrote (never re-allocated here)
            if wsbuf.numel() < need or wsbuf.device != device:
                raise _cabi.SdesError("the workspace does not hold a keep-mode wide rollout of this configuration")
        else:
            wsbuf = (workspace or Workspace()).get(need, device)
        d.workspace, d.workspace_bytes = wsbuf.data_ptr(), wsbuf.numel()
        stream = torch.cuda.current_stream(device).cuda_stream
        _cabi.check(fn_grad(C.byref(d), C.byref(g), C.c_void_p(stream)), what)
    return grad_params, grad_emb, grad_gate


def rnd_stats(rnd: torch.Tensor, mask_mode: int, max_rnd: float = 0.0,
              sample_mask: torch.Tensor | None = None) -> torch.Tensor:
    """8 doubles on the device (layout: include/sdes_b200.h `sdes_rnd_stats`)."""
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    out = torch.empty(8, dtype=torch.float64, device=r.device)
    m = None
    if sample_mask is not None:
        m = sample_mask.reshape(-1).to(torch.uint8).contiguous()
        if m.numel() != r.numel():
            raise ValueError("filter_samples must return one flag per sample")
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_rnd_stats(r.data_ptr(), r.numel(), mask_mode, float(max_rnd), _ptr(m),
                                       out.data_ptr(), C.c_void_p(stream)), "sdes_rnd_stats")
    return out


def lv_traj_stats(rnd: torch.Tensor, traj_per_sample: int, mask_mode: int, max_rnd: float = 0.0,
                  sample_mask: torch.Tensor | None = None) -> torch.Tensor:
    """[sum of per-sample variances, kept samples, samples] on the device (`sdes_lv_traj_stats`)."""
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    if r.numel() % traj_per_sample:
        raise ValueError("rnd does not hold traj_per_sample trajectories per sample")
    out = torch.empty(3, dtype=torch.float64, device=r.device)
    m = None if sample_mask is None else sample_mask.reshape(-1).to(torch.uint8).contiguous()
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_lv_traj_stats(r.data_ptr(), r.numel() // traj_per_sample, traj_per_sample, mask_mode, float(max_rnd),
                                           _ptr(m), out.data_ptr(), C.c_void_p(stream)), "sdes_lv_traj_stats")
    return out


def lv_weights(rnd: torch.Tensor, stats: torch.Tensor, mask_mode: int, max_rnd: float = 0.0,
               sample_mask: torch.Tensor | None = None, upstream: torch.Tensor | None = None) -> torch.Tensor:
    """d (lv loss) / d rnd on the device (include/sdes_b200.h `sdes_lv_weights`)."""
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    w = torch.empty_like(r)
    m = None if sample_mask is None else sample_mask.reshape(-1).to(torch.uint8).contiguous()
    up = None if upstream is None else upstream.detach().reshape(-1)[:1].to(torch.float32).contiguous()
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_lv_weights(r.data_ptr(), r.numel(), mask_mode, float(max_rnd), _ptr(m), stats.data_ptr(), _ptr(up),
                                        w.data_ptr(), C.c_void_p(stream)), "sdes_lv_weights")
    return w


def lv_traj_weights(rnd: torch.Tensor, stats3: torch.Tensor, traj_per_sample: int, mask_mode: int, max_rnd: float = 0.0,
                    sample_mask: torch.Tensor | None = None, upstream: torch.Tensor | None = None) -> torch.Tensor:
    """d (lv_traj loss) / d rnd on the device (include/sdes_b200.h `sdes_lv_traj_weights`)."""
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    w = torch.empty_like(r)
    m = None if sample_mask is None else sample_mask.reshape(-1).to(torch.uint8).contiguous()
    up = None if upstream is None else upstream.detach().reshape(-1)[:1].to(torch.float32).contiguous()
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_lv_traj_weights(r.data_ptr(), r.numel() // traj_per_sample, traj_per_sample, mask_mode, float(max_rnd),
                                             _ptr(m), stats3.data_ptr(), _ptr(up), w.data_ptr(), C.c_void_p(stream)),
                    "sdes_lv_traj_weights")
    return w


def kl_weights(rnd: torch.Tensor, stats: torch.Tensor, mask_mode: int, max_rnd: float = 0.0,
               sample_mask: torch.Tensor | None = None, upstream: torch.Tensor | None = None) -> torch.Tensor:
    """d (kl loss) / d rnd on the device (include/sdes_b200.h `sdes_kl_weights`)."""
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    w = torch.empty_like(r)
    m = None if sample_mask is None else sample_mask.reshape(-1).to(torch.uint8).contiguous()
    up = None if upstream is None else upstream.detach().reshape(-1)[:1].to(torch.float32).contiguous()
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_kl_weights(r.data_ptr(), r.numel(), mask_mode, float(max_rnd), _ptr(m), stats.data_ptr(), _ptr(up),
                                        w.data_ptr(), C.c_void_p(stream)), "sdes_kl_weights")
    return w


def importance_weights(rnd: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    lib = _cabi.lib()
    r = rnd.detach().reshape(-1).to(torch.float32).contiguous()
    w = torch.empty_like(r)
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream(r.device).cuda_stream
        _cabi.check(lib.sdes_weights(r.data_ptr(), r.numel(), stats.data_ptr(), w.data_ptr(), C.c_void_p(stream)),
                    "sdes_weights")
    return w.reshape(-1, 1)


def philox_normal(seed: int, traj_offset: int, batch: int, n_steps: int, dim: int, device) -> torch.Tensor:
    """The in-kernel noise stream written out to HBM (test hook)."""
    lib = _cabi.lib()
    out = torch.empty((n_steps, batch, dim), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _cabi.check(lib.sdes_philox_normal(seed & 0xFFFFFFFFFFFFFFFF, traj_offset, batch, n_steps, dim,
                                           out.data_ptr(), C.c_void_p(stream)), "sdes_philox_normal")
    return out
