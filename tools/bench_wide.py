"""BASELINE cfg5 on one GPU: nice/mnist-shaped target (NiceModel coupling=4, in_out_dim=d, mid_dim, hidden=5 as in
scripts/train_nice.py:67-78 of the reference, seeded random weights — no checkpoint can travel), DDS
(ExponentialIntegratorSDELoss + ScoreCtrl, conf/solver/dds.yaml), lv, cosine grid dt=0.05 end=12.8 -> T=257.

    python tools/bench_wide.py [--dim 784] [--mid 1000] [--hidden 5] [--batch 4096] [--steps 257] [--reps 3] [--engine tcgen05]

Prints one JSON line: trajectory-steps/s and algorithmic TFLOP/s (x-dependent Linear layers of the control MLP and
of the NICE forward + input-gradient backward: 2 FLOP per multiply-add, no split-precision passes counted)."""
import argparse, json, os, sys, statistics
from functools import partial
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from torch import nn
import ref_mirrors as plugins  # parameter-holder mirrors of the reference classes (tests/ref_mirrors.py)

SPEC_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_cfg5_spec.npz")


def nice_model(dim, mid, hidden):
    """NiceModel(coupling=4, in_out_dim, mid_dim, hidden, mask_config=1) of scripts/train_nice.py:67-78 with seeded random
    weights (no checkpoint can travel).  Constructed FIRST after the seed, so every caller gets the same weights."""
    torch.manual_seed(1)
    model = plugins.NiceModel(plugins.StandardLogistic(), coupling=4, in_out_dim=dim, mid_dim=mid, hidden=hidden, mask_config=1)
    with torch.no_grad():
        model.scaling.scale.normal_(0.0, 0.2)
    return model


def nice_target_dict(dim, mid, hidden) -> dict:
    """The `target` entry of a rollout spec dict for nice_model(...) (same layout as sde_sampler_b200.spec._nice_params),
    built from the mirror alone: bench.py's reference arm completes the committed cfg5 spec fixture with it (the 76 MB of
    coupling weights are regenerated from the seed instead of being committed) without importing the product."""
    model = nice_model(dim, mid, hidden)
    couplings = []
    for c in model.coupling:
        lins = [c.in_block[0]] + [b[0] for b in c.mid_block] + [c.out_block]
        couplings.append({"mask_config": int(c.mask_config),
                          "layers": [(l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()) for l in lins]})
    return {"kind": "nice", "couplings": couplings, "scale": model.scaling.scale.detach().reshape(-1).numpy().copy(), "log_norm_const": 0.0}


def build(device, dim, mid, hidden, engine, T_end=12.8, dt=0.05, steps=None):
    from sde_sampler_b200 import FusedExponentialIntegratorSDELoss

    model = nice_model(dim, mid, hidden)
    target = plugins.Nice(model=model)
    prior = plugins.IsotropicGauss(dim=dim)
    base = plugins.FourierMLP(dim=dim, num_layers=4, channels=64)
    gate = plugins.TimeEmbed(dim_out=1, num_layers=4, channels=64, last_bias_init=partial(nn.init.constant_, val=0.01))
    with torch.no_grad():
        base.out_layer.weight.normal_(0.0, 0.05)
        base.out_layer.bias.normal_(0.0, 0.05)
    ctrl = plugins.ScoreCtrl(base_model=base, clip_model=10.0, target_score=target.score, score_model=gate,
                             detach_score=False, scale_score=1.0, clip_score=10.0)
    for m in (target, prior, base, gate):
        m.to(device)
    loss = FusedExponentialIntegratorSDELoss(generative_ctrl=ctrl, sde=None, method="lv", max_rnd=1e8, alpha=1.0, sigma=1.0,
                                             engine=engine, seed=7, sync_metrics=False)

    class Solver:
        def __init__(self):
            self.target, self.clip_target = target, None

        def clipped_target_unnorm_log_prob(self, x):
            raise RuntimeError("introspected, never called")

    ts = plugins.get_timesteps(0.0, T_end, dt=dt, rescale_t="cosine").to(device)
    if steps is not None:
        ts = ts[: steps + 1].contiguous()
    return dict(loss=loss, ts=ts, terminal=Solver().clipped_target_unnorm_log_prob, second=prior.log_prob, prior=prior)


def flops_per_traj_step(dim, mid, hidden, couplings=4):
    half = dim // 2
    mlp = 2 * 64 * (dim + 2 * 64 + dim)
    nice_fwd = couplings * 2 * (half * mid + (hidden - 1) * mid * mid + mid * half)
    return mlp + 2 * nice_fwd  # forward + input-gradient backward


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=784)
    ap.add_argument("--mid", type=int, default=1000)
    ap.add_argument("--hidden", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=None, help="truncate the T=257 grid (profiling)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--engine", default="tcgen05")
    ap.add_argument("--train", action="store_true", help="also time loss(...) with grad + loss.backward()")
    args = ap.parse_args()
    from sde_sampler_b200 import _cabi

    dev = torch.device("cuda:0")
    o = build(dev, args.dim, args.mid, args.hidden, args.engine, steps=args.steps)
    T = o["ts"].shape[0] - 1
    x0 = o["prior"].sample((args.batch,))
    lib = _cabi.lib()
    with torch.no_grad():
        val, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])  # warm-up
    torch.cuda.synchronize()
    n0 = lib.sdes_launch_count()
    ms = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            val, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    launches = (lib.sdes_launch_count() - n0) // args.reps
    train_ms = None
    if args.train:
        from sde_sampler_b200.spec import ctrl_parameters
        tm = []
        for _ in range(3):
            for p_ in ctrl_parameters(o["loss"].generative_ctrl):
                p_.grad = None
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            v, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
            v.backward()
            b.record()
            torch.cuda.synchronize()
            tm.append(a.elapsed_time(b))
        train_ms = sorted(tm)[1]
    t = statistics.median(ms) * 1e-3
    f = flops_per_traj_step(args.dim, args.mid, args.hidden)
    print(json.dumps({"workload": f"NICE d={args.dim} mid={args.mid} hidden={args.hidden} DDS lv T={T} batch={args.batch}",
                      "engine": args.engine, "traj_steps_per_s": args.batch * T / t, "ms_per_rollout": t * 1e3,
                      "ms_per_time_step": t * 1e3 / T, "algorithmic_tflops": args.batch * T * f / t / 1e12,
                      "flops_per_traj_step": f, "launches_per_rollout": int(launches), "loss": float(val), "train_step_ms": train_ms}))


if __name__ == "__main__":
    main()
