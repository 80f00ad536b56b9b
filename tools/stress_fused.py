"""Repeat the one-kernel gradients (lv, kl, kl with Hessian) many times at sizes that fill the GPU and compare every run with
the first: catches rare synchronisation bugs (a phase-overrun race in the first version only showed under load).
usage: python tools/stress_fused.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from oracle import specio
from oracle.cases import CASES
from sdes_test_helpers import build_from_spec
from sde_sampler_b200 import engine as eng
from sde_sampler_b200.engine import Workspace
from sde_sampler_b200.spec import extract_spec

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
dev = torch.device("cuda:0")
for name, B in (("dis_gmm50_lv", 65536), ("dis_gmm50_kl", 65536), ("dis_gmm50_kl", 33000), ("pis_funnel10_kl", 65536), ("dis_lerp_multiwell5_klito", 40000)):
    g = specio.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    b = build_from_spec(g["spec"], dev, engine="tcgen05")
    bptt = CASES[name]["method"] in ("kl", "kl_ito")
    d = g["x0"].shape[1]
    x0 = torch.randn(B, d, device=dev, generator=torch.Generator(dev).manual_seed(1))
    spec = extract_spec(b["loss"], CASES[name]["loss"], b["ts"], b["terminal"], b["second"], train=True,
                        compute_ito=CASES[name]["method"] != "kl", return_traj=True)
    key = "score_keep" if bptt else "gate_cot"
    out = {}
    _, rnd, xs = eng.rollout(spec, x0, seed=3, engine="tcgen05", traj_tiled=True, out=out, **{key: Workspace()})
    r = rnd.reshape(-1).double()
    w = torch.full((B,), 1.0 / B, device=dev) if bptt else (2.0 * (r - r.mean()) / (B - 1)).float()
    ws = Workspace()
    first, worst = None, 0.0
    for k in range(reps):
        got = eng.lv_grad(spec, xs, w, seed=3, engine="tcgen05", bptt=bptt, workspace=ws, **{key: out.get(key)})
        torch.cuda.synchronize()
        if first is None:
            first = [None if a is None else a.clone() for a in got]
            continue
        for a, c in zip(first, got):
            if a is None:
                continue
            assert torch.isfinite(c).all(), (name, k)
            worst = max(worst, ((a - c).abs().max() / (a.abs().max() + 1e-30)).item())
    print(f"{name:28s} B={B:6d}  {reps} runs, worst relative deviation from the first run {worst:.2e}")
    assert worst < 1e-4, name
print("ok")
