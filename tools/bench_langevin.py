"""conf/solver/langevin.yaml at full size: LangevinSDE(diff_coeff=1, clip_score=1e5, terminal_t=100), EulerIntegrator(dt=0.01)
-> 10 000 steps, eval_timesteps steps=1000 -> 1001 outputs, eval_batch_size=6000, target GMM-40 (d=2).
Prints GPU time of FusedEulerIntegrator.integrate and, with --cpu, the numpy-oracle time on a small sample."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from sde_sampler_b200 import FusedEulerIntegrator
import ref_mirrors as plugins  # parameter-holder mirrors of the reference classes (tests/ref_mirrors.py)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=6000)
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
loc, scale, w = plugins.fab_gmm_params(args.dim)
target = plugins.GMM(dim=args.dim, loc=loc, scale=scale, mixture_weights=w).to(dev)
sde = plugins.LangevinSDE(target_score=target.score, diff_coeff=1.0, clip_score=1e5, terminal_t=100.0).to(dev)
integ = FusedEulerIntegrator(dt=0.01, seed=1)
ts = plugins.get_timesteps(0.0, 100.0, steps=1000).to(dev)
grid = plugins.get_timesteps(0.0, 100.0, dt=0.01).to(dev)  # eq/integrator.py:105-113
x0 = torch.randn(args.batch, args.dim, device=dev)
xs = integ.integrate(sde, ts=ts, x_init=x0, timesteps=grid)
torch.cuda.synchronize()
ms = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); xs = integ.integrate(sde, ts=ts, x_init=x0, timesteps=grid); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
out = {"workload": f"ULA GMM-40 d={args.dim} B={args.batch} steps=10000 outputs=1001", "ms": sorted(ms)[1],
       "traj_steps_per_s": args.batch * 10000 / (sorted(ms)[1] * 1e-3), "finite": bool(torch.isfinite(xs).all())}
if args.cpu:
    from oracle import rollout, philox
    import numpy as np
    Bc, steps = 256, 200
    tg = {"kind": "gmm", "loc": loc.numpy(), "scale": scale.numpy(), "log_weights": np.zeros(1, np.float32)}
    tsn = np.linspace(0, 2.0, steps + 1).astype(np.float32)
    noise = philox.normal_noise(1, Bc, steps, args.dim)
    t0 = time.perf_counter()
    rollout.langevin_integrate(tg, x0[:Bc].cpu().numpy(), tsn, tsn[::10], 1.0, 1e5, noise)
    out["cpu_numpy_traj_steps_per_s"] = Bc * steps / (time.perf_counter() - t0)
print(json.dumps(out))
