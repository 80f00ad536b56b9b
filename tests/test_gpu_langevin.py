"""-m gpu: FusedEulerIntegrator (the reference's EulerIntegrator.integrate on a LangevinSDE — the ULA sampler of
solver/langevin.py) through the plug-in interface / C ABI, against the outputs frozen from the unmodified
reference (tests/golden/ula_*.npz) and against the numpy oracle on a ragged batch; tolerance 2e-4 + 2e-4 |ref|."""
import numpy as np
import pytest
import torch

from oracle import philox, rollout as oracle_rollout
from oracle.cases import NOISE_SEED, ULA_CASES
import ref_mirrors as plugins
from sde_sampler_b200 import FusedEulerIntegrator
from sdes_test_helpers import assert_close

pytestmark = pytest.mark.gpu
RTOL = ATOL = 2e-4


def _dev():
    return torch.device("cuda:0")


def _target(tg, dim):
    if tg["kind"] == "gmm":
        K = np.asarray(tg["loc"]).shape[0]
        lw = np.asarray(tg["log_weights"], np.float64)
        w = torch.as_tensor(np.exp(lw)).float() if lw.shape[0] == K and K > 1 else torch.ones(K)
        return plugins.GMM(dim=dim, loc=torch.as_tensor(tg["loc"]), scale=torch.as_tensor(tg["scale"]), mixture_weights=w)
    if tg["kind"] == "multiwell":
        return plugins.MultiWell(dim=dim, n_double_wells=int(tg["n_dw"]), separation=float(tg["separation"]), shift=float(tg["shift"]))
    return plugins.Funnel(dim=dim, variance=float(tg["variance"]))


def _build(g, case):
    dim = g["x0"].shape[1]
    target = _target(g["target"], dim).to(_dev())
    sde = plugins.LangevinSDE(target_score=target.score, diff_coeff=g["diff_coeff"], clip_score=g["clip_score"],
                              terminal_t=case["terminal_t"]).to(_dev())
    return FusedEulerIntegrator(dt=case["dt"], seed=3), sde


@pytest.fixture(autouse=True)
def _reference_get_timesteps(monkeypatch):
    """`integrate(..., timesteps=None)` builds its grid with the reference's `sde_sampler.utils.common.get_timesteps`
    (the drop-in runs next to the reference package).  On the GPU box the reference is absent: stand its module in
    with the mirror of that one function so the default path is exercised too."""
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


@pytest.mark.parametrize("name", list(ULA_CASES))
def test_langevin_matches_reference_golden(golden, name):
    g, case = golden(name), ULA_CASES[name]
    B, d = g["x0"].shape
    n_steps = g["timesteps"].shape[0] - 1
    integ, sde = _build(g, case)
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, n_steps, d)).to(_dev())
    ts = torch.from_numpy(g["ts"]).to(_dev())
    xs = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(g["x0"]).to(_dev()), noise=noise)
    assert tuple(xs.shape) == g["xs"].shape  # (len(ts), B, d), as LangevinSolver.run consumes it
    assert_close(xs.cpu().numpy(), g["xs"], RTOL, ATOL, "xs")
    # explicit integration grid (the `timesteps=` argument of the reference signature) gives the same result
    xs2 = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(g["x0"]).to(_dev()), timesteps=torch.from_numpy(g["timesteps"]).to(_dev()),
                          noise=noise)
    assert torch.equal(xs, xs2)


def test_langevin_ragged_batch_and_philox(golden):
    g, case = golden("ula_gmm2"), ULA_CASES["ula_gmm2"]
    integ, sde = _build(g, case)
    B, d = 333, 2
    n_steps = g["timesteps"].shape[0] - 1
    x0 = np.random.default_rng(0).standard_normal((B, d)).astype(np.float32)
    noise = philox.normal_noise(NOISE_SEED + 9, B, n_steps, d)
    want = oracle_rollout.langevin_integrate(g["target"], x0, g["timesteps"], g["ts"], g["diff_coeff"], g["clip_score"], noise)
    ts = torch.from_numpy(g["ts"]).to(_dev())
    got = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()), noise=torch.from_numpy(noise).to(_dev()))
    assert_close(got.cpu().numpy(), want, RTOL, ATOL, "xs")
    # in-kernel Philox noise: finite, the right shape, first output = x_init, and different calls draw different noise
    a = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()))
    b = integ.integrate(sde, ts=ts, x_init=torch.from_numpy(x0).to(_dev()))
    assert torch.isfinite(a).all() and torch.equal(a[0].cpu(), torch.from_numpy(x0)) and not torch.equal(a[-1], b[-1])


def test_unsupported_sde_raises():
    integ = FusedEulerIntegrator()
    with pytest.raises(NotImplementedError):
        integ.integrate(torch.nn.Identity(), ts=torch.linspace(0, 1, 3, device=_dev()), x_init=torch.zeros(4, 2, device=_dev()))
