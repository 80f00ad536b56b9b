"""-m gpu: the caller-side kernels (csrc/sdes_trainer.cu) against plain PyTorch on the same inputs:
`FusedAdamEMA.step` vs clip_grad_norm_ + torch.optim.Adam + the reference's EMA rule (solver/base.py:620-684, restated
below from torch_ema's update with the reference's warm-up decay), `sample_gauss_prior` vs nn.init.trunc_normal_'s
transform on the same uniforms, `eval_moments` vs torch reductions.  Tolerances are fp32 round-off (stated per test)."""
import math

import pytest
import torch

from sde_sampler_b200 import FusedAdamEMA, eval_moments, sample_gauss_prior

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _RefEMA:
    """EMA of the reference: torch_ema.ExponentialMovingAverage.update with EMA.get_current_decay (solver/base.py:620-684)."""

    def __init__(self, params, decay, inv_gamma, power, update_after_step, update_every, min_value):
        self.shadow = [p.detach().clone() for p in params]
        self.n = 0
        self.c = dict(decay=decay, inv_gamma=inv_gamma, power=power, after=update_after_step, every=update_every, minv=min_value)

    def decay(self):
        c = self.c
        epoch = max(self.n - c["after"] - 1, 0.0)
        value = 1 - (1 + epoch / c["inv_gamma"]) ** -c["power"]
        return 0.0 if epoch <= 0 else min(max(value, c["minv"]), c["decay"])

    def update(self, params):
        self.n += 1
        if self.n % self.c["every"] != 0:
            return
        if self.n <= self.c["after"]:
            self.shadow = [p.detach().clone() for p in params]
            return
        omd = 1.0 - self.decay()
        with torch.no_grad():
            for s, p in zip(self.shadow, params):
                tmp = s - p
                tmp.mul_(omd)
                s.sub_(tmp)


def _make_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 10), (64,), (64, 128), (64, 64), (10, 64), (10,), (1, 64), (1,)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]


def test_trainer_step_matches_torch_adam_clip_ema():
    ema_cfg = dict(decay=0.9999, inv_gamma=1.0, power=0.9, update_after_step=6, update_every=2, min_value=0.0)
    ref_p, fus_p = _make_params(0), _make_params(0)
    ref = torch.optim.Adam(ref_p, lr=0.005, weight_decay=1e-7)
    ref_ema = _RefEMA(ref_p, **ema_cfg)
    fus = FusedAdamEMA(fus_p, lr=0.005, weight_decay=1e-7, grad_clip_norm=1.0, max_grad=1e6, ema=ema_cfg)
    sched_r = torch.optim.lr_scheduler.StepLR(ref, step_size=5, gamma=0.5)
    sched_f = torch.optim.lr_scheduler.StepLR(fus, step_size=5, gamma=0.5)  # the reference's schedulers attach unchanged
    g = torch.Generator().manual_seed(1)
    skipped = 0
    for it in range(16):
        grads = [torch.randn(*p.shape, generator=g).to(DEV) * (3.0 if it % 3 == 0 else 0.05) for p in ref_p]
        if it == 7:
            grads[2][3, 5] = float("nan")   # a non-finite gradient: the step must be skipped (solver/base.py:413-439)
        if it == 11:
            grads[0][0, 0] = 1e7            # inf-norm above max_grad
        for p, q, gr in zip(ref_p, fus_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        flat = torch.cat([gr.reshape(-1) for gr in grads])
        ok = bool(torch.isfinite(flat).all()) and float(flat.abs().max()) <= 1e6
        if ok:
            torch.nn.utils.clip_grad_norm_(ref_p, max_norm=1.0, norm_type=2.0, error_if_nonfinite=False)
            ref.step()
            sched_r.step()
            ref_ema.update(ref_p)
        else:
            skipped += 1
        if fus.step(sync_skip=True):   # the reference steps its scheduler only when the optimizer stepped (solver/base.py:423-436)
            sched_f.step()
        for p, q in zip(ref_p, fus_p):
            assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), it
    m = fus.metrics()
    assert m["train/skipped_steps"] == skipped == 2 and m["train/optim_steps"] == 14
    assert m["train/ema_num_updates"] == ref_ema.n
    o = 0
    for s in ref_ema.shadow:
        got = fus.ema_shadow[o:o + s.numel()].view(s.shape)
        o += s.numel()
        assert torch.allclose(s, got, rtol=2e-6, atol=2e-7)
    assert abs(m["train/ema_decay"] - ref_ema.decay()) < 1e-12  # device pow() in double vs Python float pow
    # EMA weights swap in and out (Trainable.evaluate, solver/base.py:342-346)
    before = fus_p[0].detach().clone()
    with fus.average_parameters():
        assert torch.equal(fus_p[0], fus.ema_shadow[: before.numel()].view(before.shape))
    assert torch.equal(fus_p[0], before)


def test_trainer_step_loss_check_and_state_roundtrip():
    p = _make_params(3)
    opt = FusedAdamEMA(p, lr=1e-2, max_loss=10.0)
    for q in p:
        q.grad = torch.ones_like(q)
    w0 = p[0].detach().clone()
    opt.step(loss=torch.tensor(50.0, device=DEV))      # |loss| > max_loss: skipped
    assert torch.equal(p[0], w0) and opt.metrics()["train/skipped_steps"] == 1
    opt.step(loss=torch.tensor(float("nan"), device=DEV))
    assert torch.equal(p[0], w0) and opt.metrics()["train/skipped_steps"] == 2
    opt.step(loss=torch.tensor(5.0, device=DEV))
    assert not torch.equal(p[0], w0) and opt.metrics()["train/optim_steps"] == 1
    sd = opt.state_dict()
    p2 = _make_params(3)
    opt2 = FusedAdamEMA(p2, lr=1e-2, max_loss=10.0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and opt2.metrics()["train/optim_steps"] == 1


def test_checkpoint_interop_with_torch_adam_and_torch_ema():
    """A checkpoint written by the reference's Trainable (torch.optim.Adam + torch_ema state dicts) loads into FusedAdamEMA
    and training continues bit-compatibly; the fused state dict has the torch.optim layout; bad shapes raise clearly."""
    ref_p, fus_p = _make_params(4), _make_params(4)
    ref = torch.optim.Adam(ref_p, lr=0.003, weight_decay=1e-7)
    g = torch.Generator().manual_seed(2)
    for _ in range(3):
        for p in ref_p:
            p.grad = torch.randn(*p.shape, generator=g).to(DEV) * 0.1
        ref.step()
    adam_sd = ref.state_dict()
    ema_sd = {"decay": 0.999, "num_updates": 3, "shadow_params": [p.detach().clone() * 0.5 for p in ref_p], "collected_params": None}
    with torch.no_grad():
        for p, q in zip(ref_p, fus_p):
            q.copy_(p)
    fus = FusedAdamEMA(fus_p, lr=0.003, weight_decay=1e-7, ema=dict(decay=0.999, update_after_step=0, update_every=1))
    fus.load_state_dict(adam_sd)
    fus.load_ema_state_dict(ema_sd)
    assert fus.metrics()["train/optim_steps"] == 3 and fus.metrics()["train/ema_num_updates"] == 3
    assert torch.equal(fus.ema_shadow[: ref_p[0].numel()].view(ref_p[0].shape), ref_p[0].detach() * 0.5)
    for _ in range(2):
        grads = [torch.randn(*p.shape, generator=g).to(DEV) * 0.1 for p in ref_p]
        for p, q, gr in zip(ref_p, fus_p, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        ref.step()
        fus.step()
        for p, q in zip(ref_p, fus_p):
            assert torch.allclose(p, q, rtol=2e-6, atol=2e-7)
    sd = fus.state_dict()
    assert set(sd) >= {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(fus_p)))
    assert tuple(sd["state"][2]["exp_avg"].shape) == tuple(fus_p[2].shape) and float(sd["state"][0]["step"]) == 5
    torch.optim.Adam(_make_params(4), lr=0.003).load_state_dict({"state": sd["state"], "param_groups": sd["param_groups"]})  # torch accepts it
    bad = {"state": {0: {"step": torch.tensor(1.0), "exp_avg": torch.zeros(3, 3), "exp_avg_sq": torch.zeros(3, 3)}},
           "param_groups": [{"lr": 0.1, "params": list(range(len(fus_p)))}]}
    with pytest.raises(ValueError, match="shape"):
        fus.load_state_dict(bad)
    with pytest.raises(KeyError, match="torch.optim layout"):
        fus.load_state_dict({"exp_avg": None})


def test_missing_gradient_raises_and_prior_streams_differ():
    p = _make_params(5)
    opt = FusedAdamEMA(p, lr=1e-2)
    for q in p[1:]:
        q.grad = torch.ones_like(q)
    with pytest.raises(RuntimeError, match="no gradient"):
        opt.step()
    a = sample_gauss_prior(256, 4, device=DEV)
    b = sample_gauss_prior(256, 4, device=DEV)
    assert not torch.equal(a, b)                                     # like prior.sample: a fresh draw per call
    assert torch.equal(sample_gauss_prior(256, 4, seed=9, device=DEV), sample_gauss_prior(256, 4, seed=9, device=DEV))


def test_truncated_prior_matches_trunc_normal_transform():
    """Same uniforms through nn.init.trunc_normal_'s op sequence (torch/nn/init.py) — tolerance 1e-4 absolute: erfinv is
    ill-conditioned in the tails (the bounds sit at +-3.89 sigma); measured max |delta| 3.8e-5, 1e-6 in the bulk."""
    B, d, mean, std, q = 4096, 50, 0.0, 1.0, 1e-4
    a, b = torch.distributions.Normal(mean, std).icdf(torch.tensor([q / 2, 1 - q / 2])).tolist()   # distr/gauss.py:206-213
    u = torch.rand(B, d, generator=torch.Generator().manual_seed(5))
    got = sample_gauss_prior(B, d, mean=mean, std=std, truncate=(a, b), uniforms=u, device=DEV).cpu()
    ncdf = lambda x: (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0  # noqa: E731
    lo, hi = ncdf((a - mean) / std), ncdf((b - mean) / std)
    ref = (u * (2 * hi - 1 - (2 * lo - 1)) + (2 * lo - 1)).erfinv().mul(std * math.sqrt(2.0)).add(mean).clamp(a, b)
    assert (got - ref).abs().max().item() <= 1e-4
    bulk = ref.abs() < 2.5
    assert (got - ref)[bulk].abs().max().item() <= 5e-6
    assert got.min().item() >= a and got.max().item() <= b


def test_prior_stream_is_shard_invariant_and_standard_normal():
    B, d = 1 << 16, 10
    full = sample_gauss_prior(B, d, mean=0.5, std=2.0, seed=9, device=DEV)
    half = sample_gauss_prior(B // 2, d, mean=0.5, std=2.0, seed=9, traj_offset=B // 2, device=DEV)
    assert torch.equal(full[B // 2:], half)
    z = (full.double() - 0.5) / 2.0
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1) < 5e-3
    assert abs((z ** 4).mean().item() - 3.0) < 0.05
    tr = sample_gauss_prior(B, d, truncate=(-1.0, 2.0), seed=9, device=DEV)
    assert tr.min().item() >= -1.0 and tr.max().item() <= 2.0
    # truncated-normal mean on [-1, 2]: (phi(-1) - phi(2)) / (Phi(2) - Phi(-1))
    phi = lambda x: math.exp(-x * x / 2) / math.sqrt(2 * math.pi)  # noqa: E731
    Phi = lambda x: (1 + math.erf(x / math.sqrt(2))) / 2  # noqa: E731
    assert abs(tr.double().mean().item() - (phi(-1) - phi(2)) / (Phi(2) - Phi(-1))) < 5e-3


def test_eval_moments_match_torch():
    B, d = 10000, 37
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(B, d, generator=g) * 3 + 1).to(DEV)
    w = torch.rand(B, 1, generator=g).to(DEV)
    m = eval_moments(x, w)
    ess = (w.double().sum() ** 2 / (w.double() ** 2).sum()).item()
    assert abs(m["eval/effective_sample_size"] - ess) <= 1e-9 * ess
    assert torch.allclose(m["stddevs"].float(), x.std(dim=0).cpu(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(m["means"].float(), x.mean(dim=0).cpu(), rtol=1e-5, atol=1e-6)
    assert abs(m["eval/avg_stddev"] - x.std(dim=0).mean().item()) < 1e-5
    assert "eval/effective_sample_size" not in eval_moments(x)
