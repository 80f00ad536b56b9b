"""`loss.backward()` for the log-variance losses (reference: `Trainable.step`, solver/base.py:404-407).

The forward is the fused rollout (with the trajectory kept), the backward is `sdes_rollout_lv_grad`: one pass of
the control MLP's backward over all (trajectory, step) rows on the tensor cores — see csrc/sdes_grad.cu for why
no backpropagation through time is needed (losses/oc.py:60-64: the state is driven by the detached control).

What stays in PyTorch is plumbing on T rows: the CUDA call returns d loss / d emb (T, 64) and d loss / d gate
(T, gate_dim) for the two x-independent TimeEmbed networks (models/mlp.py:43-82); their few thousand parameters
receive their gradient by chaining those cotangents through `_time_embed` below (T = 100 rows, autograd)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _cabi, engine
from .spec import ctrl_parameters


def _time_embed(t, phase, hidden, out_w, out_b):
    """TimeEmbed.forward (models/mlp.py:71-82) on the (T, 1) grid — used only to chain cotangents to its parameters."""
    coeff = torch.linspace(0.1, 100, _cabi.CHANNELS, device=t.device).unsqueeze(0)
    arg = coeff * t + phase.reshape(1, -1)
    h = torch.cat([arg.sin(), arg.cos()], dim=1)
    for w, b in hidden:
        h = F.gelu(F.linear(h, w, b))
    return F.linear(h, out_w, out_b)


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
        st = m["stats"]
        n, mean = st[0], st[1] / st[0]
        rnd = m["rnd"].reshape(-1).double()
        w = torch.where(m["keep"].reshape(-1), 2.0 * (rnd - mean) / (n - 1.0), torch.zeros_like(rnd)) * grad_out.double()
        blob = torch.cat([p.detach().reshape(-1).float() for p in params])
        wide = engine.is_wide(m["spec"])  # wide engine: the forward kept what is needed inside its own workspace
        g_blob, g_emb, g_gate = engine.lv_grad(m["spec"], m["xs"], w.float(), noise=m["noise"], seed=m["seed"],
                                               traj_offset=m["traj_offset"], engine=lo.engine,
                                               workspace=lo._workspace if wide else lo._grad_workspace, params=blob)
        n_te, n_h, n_g = ctx.counts
        grads, o = [], 0
        for p in params:
            grads.append(g_blob[o:o + p.numel()].reshape(p.shape))
            o += p.numel()
        # blob order (include/sdes_b200.h): in_w, in_b, te_phase, te hidden (w,b)*, te out (w,b), hidden (w,b)*, out (w,b), gate...
        grads[1] = g_emb.sum(dim=0)  # in_b enters through emb = timestep_embed(s) + in_b
        ts = m["spec"].ts.to(g_emb.device).float()[:-1].reshape(-1, 1)
        te_idx = list(range(2, 2 + 1 + 2 * n_te + 2))
        with torch.enable_grad():
            te_p = [params[i] for i in te_idx]
            emb = _time_embed(ts, te_p[0], [(te_p[1 + 2 * k], te_p[2 + 2 * k]) for k in range(n_te)], te_p[-2], te_p[-1])
            te_g = torch.autograd.grad(emb, te_p, grad_outputs=g_emb, allow_unused=True)
        for i, g in zip(te_idx, te_g):
            grads[i] = g if g is not None else torch.zeros_like(params[i])
        if g_gate is not None:
            g0 = 2 + 1 + 2 * n_te + 2 + 2 * n_h + 2
            g_idx = list(range(g0, g0 + 1 + 2 * n_g + 2))
            with torch.enable_grad():
                gp = [params[i] for i in g_idx]
                gate = _time_embed(ts, gp[0], [(gp[1 + 2 * k], gp[2 + 2 * k]) for k in range(n_g)], gp[-2], gp[-1])
                gg = torch.autograd.grad(gate, gp, grad_outputs=g_gate, allow_unused=True)
            for i, g in zip(g_idx, gg):
                grads[i] = g if g is not None else torch.zeros_like(params[i])
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
    """True when the training call should return a loss with a grad_fn: grad mode on, log-variance loss, trainable
    control parameters, and a configuration `sdes_rollout_lv_grad` covers."""
    if not torch.is_grad_enabled() or loss_obj.method != "lv":
        return False
    ctrl = loss_obj.generative_ctrl
    try:
        params = ctrl_parameters(ctrl)
    except AttributeError:
        return False
    dim = int(ctrl.base_model.input_embed.weight.shape[1])
    target = getattr(getattr(ctrl, "target_score", None), "__self__", None)
    wide = dim > _cabi.MAX_DIM or (target is not None and hasattr(target, "model"))
    gate = getattr(ctrl, "score_model", None)
    if wide and gate is not None and int(gate.out_layer.weight.shape[0]) != 1:
        return False  # the wide engine has a scalar gate only
    return any(p.requires_grad for p in params)
