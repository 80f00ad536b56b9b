"""-m gpu: FusedEulerIntegrator on the OU family and the PIS ControlledSDE against the golden vectors frozen
from the unmodified reference's EulerIntegrator.integrate (tests/golden/ou_*.npz, oracle/gen_golden.py), the `bm` argument,
ragged batches against the numpy oracle, and the burn-in expectation reduction of LangevinSolver.run."""
import numpy as np
import pytest
import torch

import ref_mirrors as plugins
from oracle import philox, rollout as oracle_rollout
from oracle.cases import NOISE_SEED, OU_CASES
from sde_sampler_b200 import FusedEulerIntegrator
from sdes_test_helpers import assert_close

pytestmark = pytest.mark.gpu

RTOL = ATOL = 2e-4


def _dev():
    return torch.device("cuda:0")


def _build(case):
    mk = {"vp": lambda g: plugins.VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, scale_diff_coeff=1.0, terminal_t=1.0, generative=g),
          "bm_pis": lambda g: plugins.ScaledBM(diff_coeff=0.4472135954999579, terminal_t=5.0, generative=g),
          "const_ou": lambda g: plugins.ConstOU(drift_coeff=4.5, diff_coeff=3.0, terminal_t=1.0, generative=g)}[case["sde"]]
    sde = mk(case["generative"]).to(_dev())
    if case["ctrl"] == "pis":
        class PIS:  # what solver.oc.PIS exposes to its inference process (solver/oc.py:200-208)
            def __init__(self):
                self.sde, self.prior = mk(True).to(_dev()), plugins.Delta(dim=case["dim"]).to(_dev())

            def inference_ctrl(self, t, x):
                raise RuntimeError("introspected, never called")

        sde = plugins.ControlledSDE(sde=sde, ctrl=PIS().inference_ctrl).to(_dev())
    integ = FusedEulerIntegrator(seed=5) if case["grid"] == "ts" else FusedEulerIntegrator(dt=case["grid"], seed=5)
    return integ, sde


@pytest.fixture(autouse=True)
def _reference_get_timesteps(monkeypatch):
    """integrate(..., timesteps=None) uses the reference's get_timesteps; stand the mirror in where the reference is absent."""
    import sys
    import types

    try:
        import sde_sampler.utils.common  # noqa: F401
    except ImportError:
        pkg, utils, common = types.ModuleType("sde_sampler"), types.ModuleType("sde_sampler.utils"), types.ModuleType("sde_sampler.utils.common")
        common.get_timesteps = plugins.get_timesteps
        pkg.utils, utils.common = utils, common
        for name, mod in (("sde_sampler", pkg), ("sde_sampler.utils", utils), ("sde_sampler.utils.common", common)):
            monkeypatch.setitem(sys.modules, name, mod)
    yield


@pytest.mark.parametrize("name", list(OU_CASES))
def test_affine_integrate_matches_reference_golden(golden, name):
    g, case = golden(name), OU_CASES[name]
    B, d = g["x0"].shape
    n_steps = g["timesteps"].shape[0] - 1
    integ, sde = _build(case)
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, n_steps, d)).to(_dev())
    ts = torch.from_numpy(g["ts"]).to(_dev())
    x0 = torch.from_numpy(g["x0"]).to(_dev())
    grid = torch.from_numpy(g["timesteps"]).to(_dev())
    xs = integ.integrate(sde, ts=ts, x_init=x0, timesteps=grid if case["grid"] == "ts" else None, noise=noise)
    assert tuple(xs.shape) == g["xs"].shape
    assert_close(xs.cpu().numpy(), g["xs"], RTOL, ATOL * (1 + np.abs(g["xs"]).max() * 0.0), "xs")
    # the `bm` argument of the reference signature (eq/integrator.py:116-119): a Brownian path object called once per step
    calls = []

    def bm(s, t):
        calls.append((float(s), float(t)))
        return noise[len(calls) - 1] * torch.sqrt(t - s)

    xs_bm = integ.integrate(sde, ts=ts, x_init=x0, timesteps=grid, bm=bm)
    assert len(calls) == n_steps
    assert_close(xs_bm.cpu().numpy(), g["xs"], RTOL, ATOL, "xs (bm)")


def test_affine_integrate_ragged_batch_philox_and_wide_state():
    case = OU_CASES["ou_pis_bridge4"]
    integ, sde = _build(case)
    B, d = 333, 4
    ts_np = np.linspace(0.0, 5.0, 31, dtype=np.float32)
    x0 = (np.random.default_rng(3).standard_normal((B, d)) * 1.5).astype(np.float32)
    noise = philox.normal_noise(NOISE_SEED + 3, B, 30, d)
    mu, sigma, ctrl = oracle_rollout.ou_coefficients(case, ts_np)
    want = oracle_rollout.affine_integrate(mu, sigma, ctrl, x0, ts_np, ts_np, noise)
    ts = torch.from_numpy(ts_np).to(_dev())
    got = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()), timesteps=ts, noise=torch.from_numpy(noise).to(_dev()))
    assert_close(got.cpu().numpy(), want, RTOL, ATOL, "xs")
    a = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()), timesteps=ts)
    b = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()), timesteps=ts)
    assert torch.isfinite(a).all() and torch.equal(a[0].cpu(), torch.from_numpy(x0)) and not torch.equal(a[-1], b[-1])
    # the bridge pins the process to the origin: the last Euler step leaves exactly sigma sqrt(dt) eps
    assert abs(float(a[-1].std()) - 0.4472135954999579 * np.sqrt(5.0 / 30)) < 0.02 and abs(float(a[-1].mean())) < 0.02
    # an elementwise SDE has no dimension limit: d = 784 (the cfg5 state), VP noising process, moments of the marginal
    vp = plugins.VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, terminal_t=1.0, generative=False).to(_dev())
    big = FusedEulerIntegrator(seed=1)
    tsb = torch.linspace(0, 1, 201, device=_dev())
    xb = big.integrate(vp, ts=tsb[[0, -1]], x_init=torch.full((512, 784), 2.0, device=_dev()), timesteps=tsb)
    m = float(np.exp(-0.25 * (0.1 + 10.0)))  # exp(int drift_coeff): mean factor of the VP marginal at t = 1
    assert abs(float(xb[-1].mean()) - 2.0 * m) < 0.02 and abs(float(xb[-1].var()) - (1 - m * m)) < 0.02


def test_unsupported_sdes_raise():
    integ = FusedEulerIntegrator()
    vp = plugins.VP().to(_dev())
    learned = plugins.ControlledSDE(sde=vp, ctrl=lambda t, x: x)
    with pytest.raises(NotImplementedError):
        integ.integrate(learned, ts=torch.linspace(0, 1, 3, device=_dev()), x_init=torch.zeros(4, 2, device=_dev()),
                        timesteps=torch.linspace(0, 1, 3, device=_dev()))
    with pytest.raises(NotImplementedError):
        integ.integrate(torch.nn.Identity(), ts=torch.linspace(0, 1, 3, device=_dev()), x_init=torch.zeros(4, 2, device=_dev()))


def test_burn_in_expectations_match_torch():
    """expectation_preds of LangevinSolver.run (solver/langevin.py:50-54) with EXPECTATION_FNS (distr/base.py:12-17)."""
    g = torch.Generator(_dev()).manual_seed(4)
    xs = torch.randn(37, 501, 10, device=_dev(), generator=g) * 2 + 0.3
    burn = 5
    got = FusedEulerIntegrator.expectations(xs, burn_steps=burn)
    s = xs[burn:].reshape(-1, 10).double()
    want = {"square": (s ** 2).sum(-1).mean(), "abs": s.abs().sum(-1).mean(), "sum": s.sum(-1).mean(), "square_minus_sum": (s ** 2 - s).sum(-1).mean()}
    for k, v in want.items():
        assert float(got[k]) == pytest.approx(float(v), rel=1e-9), k
