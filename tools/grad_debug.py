"""Debug helper: lv gradient vs the reference-autograd goldens, per parameter tensor, both engines."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import philox, specio
from oracle.cases import NOISE_SEED
from sde_sampler_b200.spec import ctrl_parameters
from sdes_test_helpers import build_from_spec

dev = torch.device("cuda:0")
names = sys.argv[1:] or ["dis_gmm50_lv", "dis_dw1_lv", "dds_funnel10_lv", "dis_lerptarget_gmmrand3_dimgate", "eulerdds_gmm2_lv", "dis_noscore_constou_gauss5", "dis_gmm2_lv"]
for name in names:
    g = specio.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    spec, x0 = g["spec"], g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, T, d)).to(dev)
    ref = np.asarray(g["train"]["grad_blob"], np.float64)
    for engine in ("simt", "tcgen05"):
        try:
            b = build_from_spec(spec, dev, engine=engine)
            params = ctrl_parameters(b["ctrl"])
            val, _ = b["loss"](b["ts"], torch.from_numpy(x0).to(dev), b["terminal"], b["second"], noise=noise)
            val.backward()
            torch.cuda.synchronize()
            o = 0; rows = []
            for i, p in enumerate(params):
                r = ref[o:o + p.numel()]; o += p.numel()
                got = (p.grad if p.grad is not None else torch.zeros_like(p)).double().cpu().numpy().reshape(-1)
                rows.append(f"{i}:{np.abs(got - r).max() / (np.abs(r).max() + 1e-30):.1e}")
            print(f"{name:32s} {engine:8s} loss {float(val):.5e} (ref {g['train']['loss']:.5e}) rel err per param: " + " ".join(rows), flush=True)
        except Exception as e:
            import traceback; traceback.print_exc()
            print(f"{name:32s} {engine:8s} FAILED: {type(e).__name__}: {e}", flush=True)
