"""Debug helper: run the wide-engine golden cases on cuda:0 with both engines and print the error maxima."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import philox, specio
from oracle.cases import NOISE_SEED
from sdes_test_helpers import build_from_spec

dev = torch.device("cuda:0")
names = sys.argv[1:] or ["dis_gauss100_lv", "dds_nice16_lv", "dds_nice196_lv"]
for name in names:
    g = specio.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    spec, x0 = g["spec"], g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, T, d)).to(dev)
    for engine in ("simt", "tcgen05"):
        try:
            b = build_from_spec(spec, dev, engine=engine)
            method = spec["loss"]["method"]
            kw = {"terminal_unnorm_log_prob": b["terminal"], b["second_name"]: b["second"]}
            x_T, rnd, _ = b["loss"].simulate(b["ts"], torch.from_numpy(x0).to(dev), compute_ito_int=method != "kl",
                                             change_sde_ctrl=True, return_traj=False, noise=noise, **kw)
            torch.cuda.synchronize()
            ex = np.abs(x_T.cpu().numpy() - g["train"]["x_T"]); er = np.abs(rnd.cpu().numpy() - g["train"]["rnd"])
            print(f"{name:24s} {engine:8s} max|dx_T|={np.nanmax(ex):.3e} (nan {np.isnan(ex).sum()}) max|drnd|={np.nanmax(er):.3e} (nan {np.isnan(er).sum()}) |rnd|max={np.abs(g['train']['rnd']).max():.2e}", flush=True)
        except Exception as e:
            print(f"{name:24s} {engine:8s} FAILED: {type(e).__name__}: {e}", flush=True)
