"""-m gpu: the whole training iteration of `Trainable.step` (solver/base.py:399-454) through the public API — x0 from the
fused prior sampler, loss with grad (lv: sdes_rollout_lv_grad, kl: sdes_rollout_kl_grad), `loss.backward()`, the fused
optimizer tail (`FusedAdamEMA.step`) — run for a few dozen iterations on two golden configurations.  The objective must go
down: the log-variance loss of DIS on GMM-40 d=2 and the kl loss of PIS on the funnel.  (The gradients themselves are pinned
against the reference's autograd in test_gpu_grad.py; this test is about the pieces working together, in place, with the
parameters living in the optimizer's flat buffer.)"""
import pytest
import torch

from sde_sampler_b200 import FusedAdamEMA, sample_gauss_prior
from sde_sampler_b200.spec import ctrl_parameters
from sdes_test_helpers import build_from_spec

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _train(golden, name, iters, batch, lr, x0_fn):
    g = golden(name)
    b = build_from_spec(g["spec"], DEV, engine="auto", seed=1234, sync_metrics=False)
    params = ctrl_parameters(b["ctrl"])
    opt = FusedAdamEMA(params, lr=lr, weight_decay=1e-7, grad_clip_norm=1.0,
                       ema=dict(decay=0.9999, inv_gamma=1.0, power=0.9, update_after_step=5, update_every=2))
    d = int(g["spec"]["dim"])
    losses = []
    for it in range(iters):
        opt.zero_grad()
        x0 = x0_fn(batch, d, it)
        loss, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"])
        (loss / d).backward()            # scale_loss = 1 / dim (conf/solver/oc_base.yaml:22)
        opt.step(loss=loss)
        losses.append(loss.detach())
    losses = torch.stack(losses).cpu()
    m = opt.metrics()
    assert m["train/optim_steps"] == iters and m["train/skipped_steps"] == 0
    assert torch.isfinite(losses).all()
    return losses, b, opt


def test_lv_training_reduces_the_log_variance_loss(golden):
    x0_fn = lambda B, d, it: sample_gauss_prior(B, d, seed=it, device=DEV)  # noqa: E731
    losses, b, opt = _train(golden, "dis_gmm2_lv", iters=60, batch=4096, lr=5e-3, x0_fn=x0_fn)
    first, last = losses[:8].median().item(), losses[-8:].median().item()
    assert last < 0.7 * first, (first, last, losses.tolist())
    # evaluation with the EMA weights swapped in (Trainable.evaluate, solver/base.py:342-346) runs and restores the weights
    w0 = ctrl_parameters(b["ctrl"])[0].detach().clone()
    with opt.average_parameters(), torch.no_grad():
        res = b["loss"].eval(b["ts"], x0_fn(2048, 2, 999), b["terminal"], b["second"])
    assert torch.isfinite(res.samples).all() and "log_norm_const_is" in res.log_norm_const_preds
    assert torch.equal(ctrl_parameters(b["ctrl"])[0], w0)


def test_kl_training_reduces_the_kl_loss(golden):
    x0_fn = lambda B, d, it: torch.zeros(B, d, device=DEV)  # noqa: E731  (PIS: Delta prior at 0, distr/delta.py:25-28)
    losses, _, _ = _train(golden, "pis_funnel10_kl", iters=40, batch=4096, lr=5e-3, x0_fn=x0_fn)
    first, last = losses[:6].median().item(), losses[-6:].median().item()
    assert last < first - 0.2, (first, last, losses.tolist())
