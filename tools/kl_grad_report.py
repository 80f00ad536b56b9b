"""Worst per-tensor relative error of the kl / kl_ito gradients against the reference autograd goldens, per case and engine
(the numbers quoted in DESIGN §4.5), and the truncated-prior transform error:  python tools/kl_grad_report.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
from oracle import philox, specio
from oracle.cases import CASES, NOISE_SEED
from sde_sampler_b200 import sample_gauss_prior
from sde_sampler_b200.spec import ctrl_parameters
from sdes_test_helpers import build_from_spec

dev = torch.device("cuda:0")
for name, c in CASES.items():
    if c["method"] not in ("kl", "kl_ito"):
        continue
    g = specio.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    x0 = torch.from_numpy(g["x0"]).to(dev)
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, T, d)).to(dev)
    for engine in ("simt", "tcgen05"):
        b = build_from_spec(g["spec"], dev, engine=engine)
        params = ctrl_parameters(b["ctrl"])
        val, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"], noise=noise)
        val.backward()
        ref = np.asarray(g["train"]["grad_blob"], np.float64)
        o, worst = 0, 0.0
        for p in params:
            r = ref[o:o + p.numel()].reshape(tuple(p.shape)); o += p.numel()
            got = (torch.zeros_like(p) if p.grad is None else p.grad).double().cpu().numpy()
            if np.abs(r).max() > 0:
                worst = max(worst, np.abs(got - r).max() / np.abs(r).max())
        print(f"{name:32s} {engine:8s} loss {float(val.detach()):+.6e} (ref {g['train']['loss']:+.6e})  worst rel err per tensor {worst:.2e}")
B, d, q = 4096, 50, 1e-4
a, b_ = torch.distributions.Normal(0.0, 1.0).icdf(torch.tensor([q / 2, 1 - q / 2])).tolist()
u = torch.rand(B, d, generator=torch.Generator().manual_seed(5))
got = sample_gauss_prior(B, d, truncate=(a, b_), uniforms=u, device=dev).cpu()
ncdf = lambda x: (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
lo, hi = ncdf(a), ncdf(b_)
ref = (u * (2 * hi - 1 - (2 * lo - 1)) + (2 * lo - 1)).erfinv().mul(math.sqrt(2.0)).clamp(a, b_)
print("truncated prior vs trunc_normal_ transform: max abs err %.2e" % (got - ref).abs().max().item())
