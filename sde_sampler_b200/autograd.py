"""`loss.backward()` for the log-variance losses and the kl / kl_ito losses (reference:
`Trainable.step`, solver/base.py:404-407).

The forward is the fused rollout (with the trajectory kept), the backward is ONE C-ABI call, `sdes_rollout_lv_grad`:
a pass of the control MLP's backward over all (trajectory, step) rows on the tensor cores plus the backward of the two
x-independent TimeEmbed networks (csrc/sdes_grad.cu explains why no backpropagation through time is needed —
losses/oc.py:60-64: the state is driven by the detached control).  It returns the gradient of every control parameter
in the layout of the parameter blob; this module only slices that blob back onto the caller's `nn.Parameter`s.

kl / kl_ito: the state carries the graph, so `sdes_rollout_kl_grad` first runs a reverse sweep over the stored
trajectory (csrc/sdes_adjoint.cu: discrete adjoint, control cotangent per (trajectory, step)) and then the same batched
tensor-core pass."""
from __future__ import annotations

import torch

from . import _cabi, engine
from .spec import ctrl_parameters


class LvLoss(torch.autograd.Function):
    """value = fused rollout + statistics kernel; gradient = sdes_rollout_lv_grad."""

    @staticmethod
    def forward(ctx, loss_obj, run, n_te_hidden, n_hidden, n_gate_hidden, *params):
        # `run()` performs the rollout with return_traj=True and the reductions; it returns what backward needs
        out = run()
        ctx.loss_obj, ctx.meta = loss_obj, out
        ctx.counts = (n_te_hidden, n_hidden, n_gate_hidden)
        ctx.params = params
        return out["loss"]

    @staticmethod
    def backward(ctx, grad_out):
        m, lo = ctx.meta, ctx.loss_obj
        params = ctx.params
        if m["traj_version"] != lo._traj_version:
            raise RuntimeError("the trajectory of this loss value was overwritten by a later training call of the same "
                               "loss object; call backward() before the next forward (as Trainable.step does)")
        mode = _cabi.MASK_ISFINITE if lo.max_rnd is None else _cabi.MASK_MAX_RND
        bptt = lo.method in ("kl", "kl_ito")  # the state carries the graph -> reverse sweep first (sdes_rollout_kl_grad)
        max_rnd = 0.0 if lo.max_rnd is None else lo.max_rnd
        if lo.method == "lv_traj":  # variance across each sample's trajectories (losses/oc.py:78-84); state detached like lv
            w = engine.lv_traj_weights(m["rnd"], m["stats"], lo.traj_per_sample, mode, max_rnd, m["smask"], grad_out)
        else:
            weights = engine.kl_weights if bptt else engine.lv_weights
            w = weights(m["rnd"], m["stats"], mode, max_rnd, m["smask"], grad_out)
        blob = torch.cat([p.detach().reshape(-1).float() for p in params])
        wide = engine.is_wide(m["spec"])  # wide engine: the forward kept what is needed inside its own workspace
        g_blob, g_emb, g_gate = engine.lv_grad(m["spec"], m["xs"], w, noise=m["noise"], seed=m["seed"],
                                               traj_offset=m["traj_offset"], engine=lo.engine,
                                               workspace=lo._workspace if wide else lo._grad_workspace, params=blob,
                                               bptt=bptt, gate_cot=m.get("gate_cot"), score_keep=m.get("score_keep"))
        grads, o = [], 0
        for p in params:  # blob order (include/sdes_b200.h) == ctrl_parameters order
            grads.append(g_blob[o:o + p.numel()].reshape(p.shape))
            o += p.numel()
        if lo.process_group is not None:  # every rank holds the gradient of the global loss w.r.t. its shard's rows
            import torch.distributed as dist

            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, group=lo.process_group)
            o = 0
            for i, g in enumerate(grads):
                grads[i] = flat[o:o + g.numel()].reshape(g.shape)
                o += g.numel()
        grads = [g if p.requires_grad else None for g, p in zip(grads, params)]
        return (None, None, None, None, None, *grads)


def wants_grad(loss_obj) -> bool:
    """True when the training call should return a loss with a grad_fn: grad mode on, trainable control parameters, and a
    configuration the gradient kernels cover — lv / lv_traj: `sdes_rollout_lv_grad`; kl / kl_ito: `sdes_rollout_kl_grad`
    (every engine).  A configuration the kernels do not cover raises `NotImplementedError` here, like every other
    unsupported case of the package, instead of returning a value without a grad_fn."""
    if not torch.is_grad_enabled():
        return False
    ctrl = loss_obj.generative_ctrl
    try:
        params = ctrl_parameters(ctrl)
    except AttributeError:
        return False
    dim = int(ctrl.base_model.input_embed.weight.shape[1])
    target = getattr(getattr(ctrl, "target_score", None), "__self__", None)
    wide = dim > _cabi.MAX_DIM or (target is not None and hasattr(target, "model"))
    if not any(p.requires_grad for p in params):
        return False
    gate = getattr(ctrl, "score_model", None)
    if wide and gate is not None and int(gate.out_layer.weight.shape[0]) != 1:
        # a silent grad-less value would surface as an unrelated autograd error in the caller's backward()
        raise NotImplementedError("training with a per-dimension gate (conf/model/*_dim.yaml) on the wide engine (d > 64 or a NICE "
                                  "target) is not implemented; evaluate under torch.no_grad() or use a scalar gate")
    return True
