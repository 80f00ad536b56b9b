"""Headline size, both engines on the same Philox stream: distribution of |x_T(tcgen05) - x_T(simt)| over all rows, and the
worst rows against the numpy oracle in fp32 and fp64 (which engine is off, and is the row simply ill-conditioned?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
from oracle import philox, rollout as oracle_rollout, specio
from sdes_test_helpers import build_from_spec
from sde_sampler_b200 import engine as eng
from sde_sampler_b200.spec import extract_spec

dev = torch.device("cuda:0")
spec_d = specio.load(os.path.join(ROOT, "tests", "golden", "dis_gmm50_lv.npz"))["spec"]
B, d, T = 65536, 50, 100
x0 = torch.randn(B, d, device=dev, generator=torch.Generator(dev).manual_seed(12))
out = {}
for e in ("simt", "tcgen05"):
    b = build_from_spec(spec_d, dev, engine=e)
    spec = extract_spec(b["loss"], "time_reversal", b["ts"], b["terminal"], b["second"], train=True, compute_ito=True)
    out[e] = [t.cpu().numpy() for t in eng.rollout(spec, x0, seed=77, engine=e)[:2]]
dx = np.abs(out["tcgen05"][0] - out["simt"][0]).max(axis=1)
dr = np.abs(out["tcgen05"][1] - out["simt"][1]).reshape(-1)
print("rows with max|dx_T| > 2e-4:", int((dx > 2e-4).sum()), " > 1e-3:", int((dx > 1e-3).sum()), " median %.2e  p99 %.2e  p99.9 %.2e  max %.2e" % (np.median(dx), np.quantile(dx, .99), np.quantile(dx, .999), dx.max()))
print("rnd: rows > 2e-4(1+|r|):", int((dr > 2e-4 * (1 + np.abs(out['simt'][1].reshape(-1)))).sum()), " max %.2e" % dr.max())
rows = np.argsort(-dx)[:8]
noise = np.stack([philox.normal_block(77, rows, i, d) for i in range(T)])
x32, r32, _ = oracle_rollout.rollout(spec_d, x0[rows].cpu().numpy(), noise=noise)
x64, r64, _ = oracle_rollout.rollout(spec_d, x0[rows].cpu().numpy(), noise=noise, dtype=np.float64)
for k, r in enumerate(rows):
    print(f"row {r}: |tc-simt| {dx[r]:.2e}  |tc-o64| {np.abs(out['tcgen05'][0][r]-x64[k]).max():.2e}  |simt-o64| {np.abs(out['simt'][0][r]-x64[k]).max():.2e}  |o32-o64| {np.abs(x32[k]-x64[k]).max():.2e}  x_T[0:2]={out['simt'][0][r][:2]}")
