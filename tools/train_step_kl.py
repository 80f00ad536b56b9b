"""kl training step (BASELINE cfg 3: funnel d=10, PIS, loss.method=kl, T=200): forward rollout keeping xs, reverse sweep
(csrc/sdes_adjoint.cu) and the tensor-core gradient passes — timing and launch lists:
    python tools/train_step_kl.py [--batch B] [--reps N] [--case pis_funnel10_kl]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from oracle import specio
from sde_sampler_b200.spec import ctrl_parameters
from sdes_test_helpers import build_from_spec

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--case", default="pis_funnel10_kl")
args = ap.parse_args()
dev = torch.device("cuda:0")
g = specio.load(os.path.join(ROOT, "tests", "golden", args.case + ".npz"))
b = build_from_spec(g["spec"], dev, engine="auto", sync_metrics=False)
d = int(g["spec"]["dim"])
x0 = torch.zeros(args.batch, d, device=dev) if g["spec"]["prior"] is None else torch.randn(args.batch, d, device=dev)
T = b["ts"].shape[0] - 1
out = []
for k in range(args.reps):
    for p in ctrl_parameters(b["ctrl"]):
        p.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    v, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"])
    e[1].record()
    v.backward()
    e[2].record()
    torch.cuda.synchronize()
    out.append((round(e[0].elapsed_time(e[1]), 3), round(e[1].elapsed_time(e[2]), 3)))
fw, bw = out[-1]
print("case", args.case, "B", args.batch, "T", T, "loss", float(v))
print("forward ms, backward ms per rep:", out)
print("train step traj-steps/s: %.4g" % (args.batch * T / ((fw + bw) * 1e-3)))
