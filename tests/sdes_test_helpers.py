"""Test helpers: rebuild plug-in objects (the parameter-holder mirrors in tests/ref_mirrors.py) from the raw
spec dict stored in a golden fixture, so the -m gpu tests drive the fused losses through the
same object/bound-method interface the reference's solver uses — without /root/reference."""
from __future__ import annotations

import numpy as np
import torch

from sde_sampler_b200 import (FusedExponentialIntegratorSDELoss, FusedReferenceSDELoss,
                              FusedTimeReversalLoss)
import ref_mirrors as plugins


def _load_linear(layer, w, b):
    with torch.no_grad():
        layer.weight.copy_(torch.as_tensor(w))
        layer.bias.copy_(torch.as_tensor(b))


def _load_time_embed(te, p):
    with torch.no_grad():
        te.timestep_phase.copy_(torch.as_tensor(p["phase"]).reshape(1, -1))
    for layer, (w, b) in zip(te.hidden_layer, p["hidden"]):
        _load_linear(layer, w, b)
    _load_linear(te.out_layer, p["out_w"], p["out_b"])


class SolverShim:
    """Stands in for TrainableDiff: owner of clipped_target_unnorm_log_prob (solver/oc.py:48-54)
    and, for Euler-DDS, of reference_ctrl (solver/oc.py:305-306)."""

    def __init__(self, target, clip_target, prior=None, sde=None):
        self.target, self.clip_target, self.prior, self.sde = target, clip_target, prior, sde

    def clipped_target_unnorm_log_prob(self, x):
        raise RuntimeError("introspected, never called")

    def reference_ctrl(self, t, x):
        raise RuntimeError("introspected, never called")


def build_from_spec(spec: dict, device, engine: str = "simt", **loss_kw):
    dim = int(spec["dim"])
    m = spec["mlp"]
    base = plugins.FourierMLP(dim=dim, num_layers=len(m["hidden"]) + 2)
    _load_linear(base.input_embed, m["in_w"], m["in_b"])
    _load_time_embed(base.timestep_embed, m["time_embed"])
    for layer, (w, b) in zip(base.hidden_layer, m["hidden"]):
        _load_linear(layer, w, b)
    _load_linear(base.out_layer, m["out_w"], m["out_b"])
    gate = None
    if spec.get("gate") is not None:
        g = spec["gate"]
        gate = plugins.TimeEmbed(dim_out=int(np.asarray(g["out_w"]).shape[0]), num_layers=len(g["hidden"]) + 1)
        _load_time_embed(gate, g)

    tg = spec["target"]
    if tg["kind"] == "gmm":
        lw = np.asarray(tg["log_weights"], np.float64)
        K = np.asarray(tg["loc"]).shape[0]
        w = torch.as_tensor(np.exp(lw)).float() if lw.shape[0] == K and K > 1 else (torch.ones(K) if K > 1 else None)
        target = plugins.GMM(dim=dim, loc=torch.as_tensor(tg["loc"]), scale=torch.as_tensor(tg["scale"]),
                             mixture_weights=w, log_norm_const=float(tg.get("log_norm_const", 0.0)))
    elif tg["kind"] == "gauss":
        target = plugins.Gauss(dim=dim, loc=torch.as_tensor(tg["loc"]), scale=torch.as_tensor(tg["scale"]),
                               log_norm_const=float(tg.get("log_norm_const", 0.0)))
    elif tg["kind"] == "multiwell":
        if dim == 1:
            target = plugins.DoubleWell(separation=float(tg["separation"]), shift=float(tg["shift"]))
        else:
            target = plugins.MultiWell(dim=dim, n_double_wells=int(tg["n_dw"]), separation=float(tg["separation"]),
                                       shift=float(tg["shift"]))
    elif tg["kind"] == "funnel":
        target = plugins.Funnel(dim=dim, variance=float(tg["variance"]), log_norm_const=float(tg.get("log_norm_const", 0.0)))
    elif tg["kind"] == "nice":
        cps = tg["couplings"]
        mid, half = np.asarray(cps[0]["layers"][0][0]).shape
        model = plugins.NiceModel(plugins.StandardLogistic(), coupling=len(cps), in_out_dim=dim, mid_dim=int(mid),
                                  hidden=len(cps[0]["layers"]) - 1, mask_config=int(cps[0]["mask_config"]))
        for cm, c in zip(model.coupling, cps):
            lins = [cm.in_block[0]] + [b[0] for b in cm.mid_block] + [cm.out_block]
            for lin, (w, b) in zip(lins, c["layers"]):
                _load_linear(lin, w, b)
        with torch.no_grad():
            model.scaling.scale.copy_(torch.as_tensor(tg["scale"]).reshape(1, -1))
        target = plugins.Nice(model=model, log_norm_const=float(tg.get("log_norm_const", 0.0)))
    else:
        raise ValueError(tg["kind"])

    sd = spec.get("sde")
    sde = None
    if sd is not None:
        if sd["kind"] == "vp":
            sde = plugins.VP(diff_coeff_sq_min=float(sd["beta_min"]), diff_coeff_sq_max=float(sd["beta_max"]),
                             scale_diff_coeff=float(sd["scale"]), terminal_t=float(sd["terminal_t"]),
                             generative=float(sd["sign"]) > 0)
        else:
            sde = plugins.ConstOU(drift_coeff=float(sd["drift_coeff"]), diff_coeff=float(sd["diff_coeff"]),
                                  terminal_t=float(sd["terminal_t"]), generative=float(sd["sign"]) > 0)

    def gauss(p):
        return None if p is None else plugins.Gauss(dim=dim, loc=torch.as_tensor(p["loc"]), scale=torch.as_tensor(p["scale"]))

    prior, ref = gauss(spec.get("prior")), gauss(spec.get("ref"))
    cd = spec["ctrl"]
    kw = dict(base_model=base, clip_model=cd.get("clip_model"))
    skw = dict(target_score=target.score, score_model=gate, detach_score=False,
               scale_score=cd.get("scale_score", 1.0), clip_score=cd.get("clip_score"))
    kind = cd["kind"]
    if kind == "clipped":
        ctrl = plugins.ClippedCtrl(**kw)
    elif kind == "score":
        ctrl = plugins.ScoreCtrl(**kw, **skw)
    else:
        cls = {"lerp": plugins.LerpCtrl, "lerp_prior": plugins.LerpPriorCtrl, "lerp_target": plugins.LerpTargetCtrl}[kind]
        ctrl = cls(**kw, **skw, sde=sde, prior_score=prior.score)

    ls = spec["loss"]
    shim = SolverShim(target, tg.get("clip_target"), prior=prior, sde=sde)
    lkw = dict(generative_ctrl=ctrl, sde=sde, method=ls["method"], max_rnd=ls.get("max_rnd"),
               traj_per_sample=int(ls.get("traj_per_sample", 1)), engine=engine, **loss_kw)
    if ls["kind"] == "time_reversal":
        loss = FusedTimeReversalLoss(**lkw)
        second = prior.log_prob if prior is not None else None
        second_name = "initial_log_prob"
    elif ls["kind"] == "reference_sde":
        loss = FusedReferenceSDELoss(**lkw, reference_ctrl=shim.reference_ctrl if ls.get("reference_ctrl") else None)
        second, second_name = ref.log_prob, "reference_log_prob"
    else:
        loss = FusedExponentialIntegratorSDELoss(**lkw, alpha=float(ls["alpha"]), sigma=float(ls["sigma"]))
        second, second_name = ref.log_prob, "reference_log_prob"
    for mod in (base, gate, target, sde, prior, ref):
        if mod is not None:
            mod.to(device)
    ts = torch.as_tensor(np.asarray(spec["ts"], np.float32)).to(device)
    return dict(loss=loss, ts=ts, terminal=shim.clipped_target_unnorm_log_prob, second=second,
                second_name=second_name, target=target, prior=prior, ref=ref, sde=sde, ctrl=ctrl)


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{what}: max violation {err.max():.3e}; max abs diff {np.abs(a - b).max():.3e}"
