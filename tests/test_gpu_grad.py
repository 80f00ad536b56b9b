"""-m gpu: `loss.backward()` of the log-variance losses and of the kl / kl_ito losses (SURVEY §8f-2:
backpropagation through time — reverse sweep csrc/sdes_adjoint.cu + the same tensor-core passes; the goldens hold the
reference's autograd gradients for those too) — the tensor-core gradient path
(csrc/sdes_grad.cu through sde_sampler_b200/autograd.py) against the gradients the UNMODIFIED reference produces with
its own autograd on the same x0 / noise (`train/grad_blob` in tests/golden, frozen by oracle/gen_golden.py:
`loss.simulate(...)`, `loss.compute_loss(rnd)`, `.backward()`; parameters flattened in blob order).

Tolerance: per parameter tensor, max |delta| <= 2e-3 * max |ref| of that tensor (+1e-7): the gradient is a sum over
B*T rows of products evaluated with 16-bit-split operands and fp32 accumulation; the reference sums the same terms
in a different order in fp32."""
import numpy as np
import pytest
import torch

from oracle import philox
from oracle.cases import CASES, NOISE_SEED
from sde_sampler_b200 import _cabi
from sde_sampler_b200.spec import ctrl_parameters
from sdes_test_helpers import build_from_spec

pytestmark = pytest.mark.gpu
ENGINES = ["simt", "tcgen05"]
GRAD_CASES = [n for n, c in CASES.items() if c["method"] in ("lv", "lv_traj")]  # incl. the wide engine: NICE targets, d = 100 / 196
KL_GRAD_CASES = [n for n, c in CASES.items() if c["method"] in ("kl", "kl_ito")]


def _dev():
    return torch.device("cuda:0")


def _grads(b, x0, noise):
    params = ctrl_parameters(b["ctrl"])
    for p in params:
        p.grad = None
    val, metrics = b["loss"](b["ts"], x0, b["terminal"], b["second"], noise=noise)
    assert val.requires_grad
    val.backward()
    return val, params, [torch.zeros_like(p) if p.grad is None else p.grad.detach().clone() for p in params]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", KL_GRAD_CASES)
def test_kl_gradient_matches_reference_autograd(golden, name, engine):
    """Backpropagation through time: every loss kind (time reversal, reference SDE with and without the Euler-DDS
    reference control, exponential integrator), kl and kl_ito, active clips, per-dimension gate, and each target's
    second derivative; the GMM score enters as a constant exactly as in the reference (engine.kl_grad_flags)."""
    test_lv_gradient_matches_reference_autograd(golden, name, engine)


def test_kl_gradient_ignores_filtered_trajectories_and_detach_score(golden):
    """rnd[mask].mean(): a filtered trajectory contributes nothing, whatever its state holds; detach_score=True removes the
    score term's x-dependence from the adjoint (models/reparam.py:58) and must change the gradient."""
    g = golden("pis_funnel10_kl")
    x0 = torch.from_numpy(g["x0"]).to(_dev())
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, T, d)).to(_dev())
    b = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    _, params, ref = _grads(b, x0, noise)
    # (1) max_rnd below the largest rnd: those trajectories are dropped from value and gradient
    with torch.no_grad():
        _, rnd, _ = b["loss"].simulate(b["ts"], x0, b["terminal"], b["second"], noise=noise)
    thr = float(rnd.reshape(-1).sort().values[-8])
    b2 = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    b2["loss"].max_rnd = thr
    keep = (rnd.reshape(-1) < thr)
    val2, _, got2 = _grads(b2, x0, noise)
    b3 = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    val3, _, got3 = _grads(b3, x0[keep], noise[:, keep])
    assert abs(float(val2) - float(val3)) <= 1e-5 * (1 + abs(float(val3)))
    for u, v in zip(got2, got3):
        assert (u - v).abs().max().item() <= 2e-3 * v.abs().max().item() + 1e-7
    # (2) detach_score
    b4 = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    b4["ctrl"].detach_score = True
    _, _, got4 = _grads(b4, x0, noise)
    diff = max((u - v).abs().max().item() / (v.abs().max().item() + 1e-30) for u, v in zip(got4, ref))
    assert diff > 1e-2


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", GRAD_CASES)
def test_lv_gradient_matches_reference_autograd(golden, name, engine):
    g = golden(name)
    spec, x0 = g["spec"], g["x0"]
    T, (B, d) = g["ts"].shape[0] - 1, x0.shape
    noise = torch.from_numpy(philox.normal_noise(NOISE_SEED, B, T, d)).to(_dev())
    b = build_from_spec(spec, _dev(), engine=engine)
    # lv_traj fixtures hold the repeated batch; the training call is given the initial points and repeats them itself
    tps = int(spec["loss"].get("traj_per_sample", 1))
    val, params, grads = _grads(b, torch.from_numpy(x0[: B // tps]).to(_dev()), noise)
    ref_loss = g["train"]["loss"]
    assert abs(float(val.detach()) - ref_loss) <= 1e-3 * (1 + abs(ref_loss))
    ref = np.asarray(g["train"]["grad_blob"], np.float64)
    o = 0
    worst = 0.0
    for i, (p, gr) in enumerate(zip(params, grads)):
        r = ref[o:o + p.numel()].reshape(tuple(p.shape))
        o += p.numel()
        got = gr.double().cpu().numpy()
        scale = np.abs(r).max()
        err = np.abs(got - r).max()
        worst = max(worst, err / (scale + 1e-30) if scale > 0 else err)
        assert err <= 2e-3 * scale + 1e-7, f"parameter {i} {tuple(p.shape)}: max|delta|={err:.3e}, max|ref|={scale:.3e}"
    assert o == ref.size


def test_lv_gradient_engines_agree_and_chunking_is_invariant(golden):
    """1 000 trajectories of the headline configuration: tcgen05 GEMMs vs CUDA-core GEMMs on the same operand images,
    and the result must not depend on how the (trajectory, step) rows are cut into chunks."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.spec import extract_spec

    g = golden("dis_gmm50_lv")
    outs = []
    for engine, chunk in (("tcgen05", 0), ("tcgen05", 3000), ("simt", 0)):
        b = build_from_spec(g["spec"], _dev(), engine=engine)
        B, d, T = 1000, 50, g["ts"].shape[0] - 1
        x0 = torch.randn(B, d, device=_dev(), generator=torch.Generator(_dev()).manual_seed(2))
        spec = extract_spec(b["loss"], "time_reversal", b["ts"], b["terminal"], b["second"], train=True, compute_ito=True,
                            return_traj=True)
        x_T, rnd, xs = eng.rollout(spec, x0, seed=5, engine=engine)
        r = rnd.reshape(-1).double()
        w = (2.0 * (r - r.mean()) / (B - 1)).float()
        outs.append(eng.lv_grad(spec, xs, w, seed=5, engine=engine, chunk_rows=chunk))
    for other in outs[1:]:
        for a, c in zip(outs[0], other):
            if a is None:
                assert c is None
                continue
            scale = a.abs().max().item()
            assert (a - c).abs().max().item() <= 2e-3 * scale + 1e-7


@pytest.mark.parametrize("name,B", [("dis_gmm50_lv", 1000), ("dis_gmm50_lv", 40000), ("dds_funnel10_lv", 3000), ("eulerdds_gmm2_lv", 777), ("dis_dw1_lv", 130)])
def test_lv_fused_kernel_matches_layerwise_passes(golden, name, B):
    """The one-kernel lv gradient (csrc/sdes_grad_fused.cuh: replayed forward, dgrad chain and weight-gradient MMAs per row
    tile, operands in shared memory, accumulators in TMEM) against the layer-by-layer GEMM passes on the same stored
    trajectory: ragged last tile, several steps per CTA, all three loss kinds, in-kernel Philox noise."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.engine import Workspace
    from sde_sampler_b200.spec import extract_spec

    if name not in CASES:
        pytest.skip("no such golden")
    g = golden(name)
    b = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    d, T = g["x0"].shape[1], g["ts"].shape[0] - 1
    x0 = torch.randn(B, d, device=_dev(), generator=torch.Generator(_dev()).manual_seed(4))
    kind = CASES[name]["loss"]
    spec = extract_spec(b["loss"], kind, b["ts"], b["terminal"], b["second"], train=True, compute_ito=True, return_traj=True)
    out = {}
    x_T, rnd, xs = eng.rollout(spec, x0, seed=5, engine="tcgen05", gate_cot=Workspace(), out=out)
    r = rnd.reshape(-1).double()
    w = (2.0 * (r - r.mean()) / (B - 1)).float()
    gc = out.get("gate_cot")
    fused = eng.lv_grad(spec, xs, w, seed=5, engine="tcgen05", gate_cot=gc)
    layer = eng.lv_grad(spec, xs, w, seed=5, engine="tcgen05", gate_cot=gc, grad_flags=_cabi.GRAD_LAYERWISE_SWEEP)
    for i, (a, c) in enumerate(zip(layer, fused)):
        if a is None:
            assert c is None
            continue
        scale = a.abs().max().item()
        assert (a - c).abs().max().item() <= 1e-3 * scale + 1e-7, f"output {i}: {(a - c).abs().max().item():.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("name,B", [("dis_gmm50_kl", 1000), ("dis_gmm50_kl", 33000), ("dis_lerpprior_multiwell4", 3000), ("dis_gmm2_kl", 130),
                                    ("pis_funnel10_kl", 20000), ("dds_funnel10_kl", 700), ("eulerdds_gauss3_kl", 2000),
                                    ("dis_lerp_multiwell5_klito", 5000)])
def test_kl_fused_kernel_matches_stepwise_sweep(golden, name, B):
    """kl / kl_ito with the forward's score_keep: the whole reverse sweep as one persistent kernel (adjoint in registers, a
    tile's steps walked backwards by one CTA) against the step-by-step sweep (one elementwise kernel + dgrad chain per step)
    on the same stored trajectory — ragged last tile, more tiles than SMs, kl and kl_ito, Lerp / LerpPrior / Score controls, all three loss kinds, targets
    whose Hessian enters the sweep (funnel, Gauss, multi-well) and a per-dimension gate."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.engine import Workspace
    from sde_sampler_b200.spec import extract_spec

    g = golden(name)
    b = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    d, T = g["x0"].shape[1], g["ts"].shape[0] - 1
    x0 = torch.randn(B, d, device=_dev(), generator=torch.Generator(_dev()).manual_seed(4))
    spec = extract_spec(b["loss"], CASES[name]["loss"], b["ts"], b["terminal"], b["second"], train=True,
                        compute_ito=CASES[name]["method"] == "kl_ito", return_traj=True)
    out = {}
    x_T, rnd, xs = eng.rollout(spec, x0, seed=5, engine="tcgen05", traj_tiled=True, score_keep=Workspace(), out=out)
    w = torch.full((B,), 1.0 / B, device=_dev())
    w[::7] = 0.0  # filtered trajectories
    assert out.get("score_keep") is not None
    fused = eng.kl_grad(spec, xs, w, seed=5, engine="tcgen05", score_keep=out["score_keep"])
    step = eng.kl_grad(spec, xs, w, seed=5, engine="tcgen05")
    for i, (a, c) in enumerate(zip(step, fused)):
        if a is None:
            assert c is None
            continue
        scale = a.abs().max().item()
        assert (a - c).abs().max().item() <= 1e-3 * scale + 1e-7, f"output {i}: {(a - c).abs().max().item():.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("bptt", [False, True])
def test_fused_gradient_full_size_properties(golden, bptt):
    """BASELINE's headline size (65 536 trajectories x T steps, d = 50) through the one-kernel gradient — too large for the
    oracle, so the size-independent properties: the gradient is linear in the per-trajectory weights (w = w1 + w2 gives the
    sum of the two gradients) and additive over shards addressed through the global Philox counters (traj_offset)."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.engine import Workspace
    from sde_sampler_b200.spec import extract_spec

    name = "dis_gmm50_kl" if bptt else "dis_gmm50_lv"
    g = golden(name)
    b = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    B, d = 65536, 50
    x0 = torch.randn(B, d, device=_dev(), generator=torch.Generator(_dev()).manual_seed(9))
    spec = extract_spec(b["loss"], "time_reversal", b["ts"], b["terminal"], b["second"], train=True, compute_ito=not bptt, return_traj=True)
    key = "score_keep" if bptt else "gate_cot"

    def fwd(x, off):
        out = {}
        _, rnd, xs = eng.rollout(spec, x, seed=5, traj_offset=off, engine="tcgen05", traj_tiled=True, out=out, **{key: Workspace()})
        return rnd.reshape(-1), xs, out[key]

    def grad(xs, w, aux, off):
        return eng.lv_grad(spec, xs, w, seed=5, traj_offset=off, engine="tcgen05", bptt=bptt, **{key: aux})

    rnd, xs, aux = fwd(x0, 0)
    r = rnd.double()
    w = (2.0 * (r - r.mean()) / (B - 1)).float() if not bptt else torch.full((B,), 1.0 / B, device=_dev())
    full = grad(xs, w, aux, 0)
    gen = torch.Generator(_dev()).manual_seed(1)
    w1 = w * torch.rand(B, device=_dev(), generator=gen)
    parts = [grad(xs, w1, aux, 0), grad(xs, w - w1, aux, 0)]
    halves = []
    for h in range(2):
        sl = slice(h * B // 2, (h + 1) * B // 2)
        _, xs_h, aux_h = fwd(x0[sl], h * B // 2)
        halves.append(grad(xs_h, w[sl], aux_h, h * B // 2))
    for what, pair in (("linearity in w", parts), ("shard additivity", halves)):
        for i, (a, p0, p1) in enumerate(zip(full, *pair)):
            if a is None:
                continue
            scale = a.abs().max().item()
            assert (a - (p0 + p1)).abs().max().item() <= 1e-3 * scale + 1e-7, f"{what}, output {i}"
            assert torch.isfinite(a).all()


def test_kl_gradient_is_additive_over_shards_and_engine_independent(golden):
    """Size-independent properties of the BPTT gradient at 4 096 trajectories of the cfg-3 configuration (funnel d=10, PIS,
    kl): the gradient is linear in the per-trajectory weights, so two half-batch calls (global Philox counters via
    traj_offset) add up to the full-batch call; the tensor-core sweep (fused dgrad chain, or one GEMM launch per layer) and
    the thread-per-trajectory fp32 sweep agree; the row chunking of the GEMM passes does not matter."""
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.spec import extract_spec

    g = golden("pis_funnel10_kl")
    B, d, T = 4096, 10, g["ts"].shape[0] - 1
    x0 = torch.zeros(B, d, device=_dev())
    outs = {}
    for engine in ("tcgen05", "simt"):
        b = build_from_spec(g["spec"], _dev(), engine=engine)
        spec = extract_spec(b["loss"], "reference_sde", b["ts"], b["terminal"], b["second"], train=True, compute_ito=False,
                            return_traj=True)
        x_T, rnd, xs = eng.rollout(spec, x0, seed=11, engine=engine)
        w = torch.full((B,), 1.0 / B, device=_dev())
        outs[engine] = eng.kl_grad(spec, xs, w, seed=11, engine=engine)
        if engine == "tcgen05":
            outs["chunked"] = eng.kl_grad(spec, xs, w, seed=11, engine=engine, chunk_rows=50000)
            outs["layerwise"] = eng.kl_grad(spec, xs, w, seed=11, engine=engine, grad_flags=_cabi.GRAD_LAYERWISE_SWEEP)
            parts = []
            for h in range(2):
                sl = slice(h * B // 2, (h + 1) * B // 2)
                _, _, xs_h = eng.rollout(spec, x0[sl], seed=11, traj_offset=h * B // 2, engine=engine)
                parts.append(eng.kl_grad(spec, xs_h, w[sl], seed=11, traj_offset=h * B // 2, engine=engine))
            outs["shards"] = tuple(None if a is None else a + c for a, c in zip(*parts))
    ref = outs["tcgen05"]
    for key in ("simt", "chunked", "layerwise", "shards"):
        for a, c in zip(ref, outs[key]):
            if a is None:
                assert c is None
                continue
            assert (a - c).abs().max().item() <= 2e-3 * a.abs().max().item() + 1e-7, key


def test_no_grad_call_returns_plain_value(golden):
    g = golden("dis_gmm2_lv")
    b = build_from_spec(g["spec"], _dev(), engine="tcgen05")
    x0 = torch.from_numpy(g["x0"]).to(_dev())
    with torch.no_grad():
        val, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"])
    assert not val.requires_grad
    val2, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"])
    assert val2.requires_grad and val2.grad_fn is not None
