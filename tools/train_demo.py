"""A complete training run through the drop-in pieces — what `python scripts/main.py solver=basic_dis target=gmm
loss.method=lv` does in the reference (`Trainable.run`, solver/base.py:456-498), here without Hydra: mirror objects
(`tests/ref_mirrors.py`), the fused loss, `loss.backward()` on the tensor cores, the fused optimizer tail and the fused
prior sampler; every `--eval-every` iterations `loss.eval` (EMA weights swapped in) gives the importance-sampling estimate
of log Z (0 for the normalised GMM) and the ESS.

    python tools/train_demo.py [--iters 400] [--batch 8192] [--method lv|kl] [--dim 2]
"""
import argparse
import os
import sys
import time
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from torch import nn

from sde_sampler_b200 import FusedAdamEMA, FusedTimeReversalLoss, eval_moments, sample_gauss_prior
import ref_mirrors as plugins  # parameter-holder mirrors of the reference classes (tests/ref_mirrors.py)
from sde_sampler_b200.spec import ctrl_parameters

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=400)
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--eval-batch", type=int, default=65536)
ap.add_argument("--eval-every", type=int, default=100)
ap.add_argument("--method", default="lv", choices=["lv", "kl"])
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--steps", type=int, default=100)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(1)                                                     # conf/base.yaml:8
d = args.dim

# target/gmm.yaml ("fab" GMM-40, distr/gauss.py:42-62; zero-padded to d > 2 as SURVEY §8d prescribes), prior/gauss.yaml, sde/vp.yaml
g = torch.Generator().manual_seed(42)
loc = (torch.rand((40, 2), generator=g) - 0.5) * 2 * 40
if d > 2:
    loc = torch.cat([loc, torch.zeros(40, d - 2)], dim=1)
scale = torch.nn.functional.softplus(torch.tensor(1.0)) * torch.ones_like(loc)
target = plugins.GMM(dim=d, loc=loc, scale=scale, mixture_weights=torch.ones(40), log_norm_const=0.0).to(dev)
prior = plugins.IsotropicGauss(dim=d, loc=0.0, scale=1.0).to(dev)
sde = plugins.VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, scale_diff_coeff=1.0, terminal_t=1.0, generative=True).to(dev)
# model/lerp.yaml over model/base/fouriermlp.yaml + time_embed.yaml
base = plugins.FourierMLP(dim=d, num_layers=4).to(dev)
gate = plugins.TimeEmbed(dim_out=1, num_layers=4, last_bias_init=partial(nn.init.constant_, val=1.0)).to(dev)
ctrl = plugins.LerpCtrl(base_model=base, clip_model=1e4, target_score=target.score, score_model=gate, detach_score=False,
                        scale_score=1.0, clip_score=1e4, sde=sde, prior_score=prior.score)


class Solver:  # stands in for TrainableDiff: owner of clipped_target_unnorm_log_prob (solver/oc.py:48-54)
    def __init__(self):
        self.target, self.clip_target = target, None

    def clipped_target_unnorm_log_prob(self, x):
        return self.target.unnorm_log_prob(x)


solver = Solver()
loss_fn = FusedTimeReversalLoss(generative_ctrl=ctrl, sde=sde, method=args.method, max_rnd=None, sync_metrics=False)
ts = torch.linspace(0.0, 1.0, args.steps + 1, device=dev)                # get_timesteps(0, T, steps) (utils/common.py:18-55)
opt = FusedAdamEMA(ctrl_parameters(ctrl), lr=0.005, weight_decay=1e-7, grad_clip_norm=1.0,
                   ema=dict(decay=0.9999, inv_gamma=1.0, power=0.9, update_after_step=max(args.iters - 1500, 0), update_every=5))

t0, ev0 = time.time(), torch.cuda.Event(enable_timing=True)
ev0.record()
for it in range(1, args.iters + 1):
    opt.zero_grad()
    x0 = sample_gauss_prior(args.batch, d, seed=it, device=dev)
    loss, _ = loss_fn(ts, x0, solver.clipped_target_unnorm_log_prob, prior.log_prob)
    (loss / d).backward()                                                # scale_loss = 1 / dim
    opt.step(loss=loss)
    if it % args.eval_every == 0 or it == args.iters:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.eval_every
        with opt.average_parameters(), torch.no_grad():
            xe = sample_gauss_prior(args.eval_batch, d, seed=10_000 + it, device=dev)
            res = loss_fn.eval(ts, xe, solver.clipped_target_unnorm_log_prob, prior.log_prob, compute_weights=True, return_traj=False)
        mom = eval_moments(res.samples, res.weights)
        m = opt.metrics()
        print(f"iter {it:5d}  loss {float(loss.detach()):10.4f}  log Z (is) {res.log_norm_const_preds['log_norm_const_is']:+.4f}  "
              f"lb {res.log_norm_const_preds['log_norm_const_lb_ito']:+.4f}  ESS/N {mom['eval/norm_effective_sample_size']:.3f}  "
              f"avg stddev {mom['eval/avg_stddev']:.2f}  |grad| {m['train/grad_norm']:.2e}  skipped {m['train/skipped_steps']}  "
              f"{ms:.2f} ms/iter = {args.batch * args.steps / ms * 1e3:.3g} traj-steps/s", flush=True)
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
print(f"wall {time.time() - t0:.1f} s for {args.iters} iterations (B={args.batch}, T={args.steps}, d={d}, method={args.method})")
