"""TEST INFRASTRUCTURE — freeze outputs of the UNMODIFIED reference rollout as golden vectors.

    python -m oracle.gen_golden            # writes tests/golden/<case>.npz

Runs only where /root/reference exists (the build container; it cannot travel to the GPU
box).  For every case in `oracle/cases.py` it builds the reference objects without Hydra
(oracle/ref_harness.py), draws x0 from the reference prior, injects the Philox noise stream
of `oracle/philox.py` through `torch.randn_like`, calls the reference's own
`loss.simulate` / `loss.__call__` / `loss.eval` (sde_sampler/losses/oc.py) and stores

    spec/...   raw parameters as extracted by sde_sampler_b200.spec.extract_spec
    x0, train/{x_T, rnd, loss, n_filtered}, eval<i>/{x_T, rnd, xs?, log_norm_const_*, lv_loss}

The reference has no rollout fixtures of its own — these are the pin.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import philox, ref_harness, specio  # noqa: E402
from oracle.cases import CASES, EVAL_CASES, NOISE_SEED  # noqa: E402


def reference_spec(built: dict, case: dict, compute_ito: bool) -> dict:
    """sde_sampler_b200.spec.extract_spec applied to the UNMODIFIED reference objects of a case, the way the solver hands
    them to the loss: bound methods of an owner object (solver/oc.py:158-163, :213-215, :305-306)."""
    from sde_sampler_b200.spec import extract_spec

    class _SolverShim:  # owner of clipped_target_unnorm_log_prob: introspected, never called
        def __init__(self, target, clip_target):
            self.target = target
            self.clip_target = clip_target

        def clipped_target_unnorm_log_prob(self, x):
            raise RuntimeError("shim is introspected, never called")

    loss = built["loss"]
    shim = _SolverShim(built["target"], case.get("clip_target"))
    if case.get("euler_dds"):
        class _EulerShim:
            def __init__(self, prior, sde):
                self.prior, self.sde = prior, sde

            def reference_ctrl(self, t, x):
                return self.sde.diff(t, x) * self.prior.score(x)
        loss.reference_ctrl = _EulerShim(built["prior"], built["sde"]).reference_ctrl
    spec = extract_spec(loss, case["loss"], built["ts"], shim.clipped_target_unnorm_log_prob, built["second"],
                        train=True, compute_ito=compute_ito)
    return spec.to_dict()


def run_case(name: str, case: dict) -> dict:
    import torch

    from sde_sampler_b200.spec import extract_spec

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    built = ref_harness.build_case(case)
    loss, ts = built["loss"], built["ts"]
    B, d = case["batch"], case["dim"]
    T = ts.shape[0] - 1
    torch.manual_seed(1000 + case.get("seed", 1))
    tps = case.get("traj_per_sample", 1)
    x0 = built["prior"].sample((B // tps,)).float()
    if tps != 1:
        # lv_traj: the fixture holds the batch as `__call__` repeats it (losses/oc.py:240-241): tps stacked copies of B / tps
        # initial points, so that `simulate` / `compute_loss` see exactly what the reference's training call sees
        x0 = x0.repeat(tps, 1, 1).reshape(-1, x0.shape[-1])
    # spread the initial points a little for Delta priors? No: keep the reference semantics.
    noise = philox.normal_noise(NOISE_SEED, B, T, d)
    noise_t = torch.from_numpy(noise)

    method = case["method"]
    compute_ito = method != "kl"
    change = method in ("lv", "lv_traj")
    out = {"x0": x0.numpy().copy(), "ts": ts.numpy().copy(), "torch_version": torch.__version__}

    kind = case["loss"]
    kw = dict(terminal_unnorm_log_prob=built["terminal"])
    if kind == "time_reversal":
        kw["initial_log_prob"] = built["second"]
        sim_kw = dict(train=True)
    else:
        kw["reference_log_prob"] = built["second"]
        sim_kw = {}

    # --- train-mode rollout exactly as __call__ drives it (losses/oc.py:232-256 etc.)
    with ref_harness.injected_noise(noise_t):
        x_T, rnd, _ = loss.simulate(ts, x0.clone(), compute_ito_int=compute_ito,
                                    change_sde_ctrl=change, return_traj=False, **kw, **sim_kw)
    n_before = loss.n_filtered
    lval, metrics = loss.compute_loss(rnd.detach(), samples=x_T.detach())
    out["train"] = {"x_T": x_T.detach().numpy().copy(), "rnd": rnd.detach().numpy().copy(),
                    "loss": float(lval), "n_filtered": int(loss.n_filtered - n_before)}

    # --- gradient of the loss w.r.t. every control parameter by the reference's own autograd
    #     (`loss.backward()` of Trainable.step, solver/base.py:404-407), flattened in parameter-blob order.
    #     lv: the state is detached (one MLP backward over all rows); kl / kl_ito: backpropagation through time.
    if method in ("lv", "lv_traj", "kl", "kl_ito"):
        from sde_sampler_b200.spec import ctrl_parameters

        params = ctrl_parameters(built["ctrl"])
        for q in params:
            q.grad = None
        n_keep = loss.n_filtered
        with ref_harness.injected_noise(noise_t):
            x_g, rnd_g, _ = loss.simulate(ts, x0.clone(), compute_ito_int=compute_ito, change_sde_ctrl=change,
                                          return_traj=False, **kw, **sim_kw)
        lg, _ = loss.compute_loss(rnd_g, samples=x_g)
        lg.backward()
        loss.n_filtered = n_keep
        out["train"]["grad_blob"] = np.concatenate(
            [(q.grad if q.grad is not None else torch.zeros_like(q)).detach().reshape(-1).numpy() for q in params]).astype(np.float32)
        for q in params:
            q.grad = None

    out["spec"] = reference_spec(built, case, compute_ito)

    # --- eval-mode rollouts (losses/oc.py:258-278): train=False, change_sde_ctrl=False
    for j, (cw, rt) in enumerate(EVAL_CASES.get(name, [])):
        with torch.no_grad(), ref_harness.injected_noise(noise_t):
            res = loss.eval(ts, x0.clone(), compute_weights=cw, return_traj=rt, **kw)
        with torch.no_grad(), ref_harness.injected_noise(noise_t):
            ekw = dict(train=False) if kind == "time_reversal" else dict(change_sde_ctrl=False)
            _, rnd_e, _ = loss.simulate(ts, x0.clone(), compute_ito_int=cw, return_traj=False, **kw, **ekw)
        e = {"compute_weights": bool(cw), "return_traj": bool(rt),
             "x_T": res.samples.numpy().copy(), "rnd": rnd_e.numpy().copy(),
             "log_norm_const_preds": {k: float(v) for k, v in res.log_norm_const_preds.items()},
             "metrics": {k: float(v) for k, v in res.metrics.items()}}
        if cw:
            e["weights"] = res.weights.numpy().copy()
        if rt:
            e["xs"] = res.xs.numpy().copy()
        out[f"eval{j}"] = e
    return out


def run_ula_case(name: str, case: dict) -> dict:
    """EulerIntegrator.integrate on a LangevinSDE of the unmodified reference (solver/langevin.py:34-63) with the
    Philox stream injected through `torch.randn` (eq/integrator.py:116 draws with torch.randn, not randn_like)."""
    import torch

    from sde_sampler_b200.spec import _target_params

    ref_harness.import_reference()
    from sde_sampler.eq.integrator import EulerIntegrator
    from sde_sampler.eq.sdes import LangevinSDE
    from sde_sampler.utils.common import get_timesteps

    d, B = case["dim"], case["batch"]
    target = ref_harness.build_target(case["target"], d)
    sde = LangevinSDE(target_score=target.score, diff_coeff=case["diff_coeff"], clip_score=case["clip_score"],
                      terminal_t=case["terminal_t"])
    integ = EulerIntegrator(dt=case["dt"])
    ts = get_timesteps(0.0, case["terminal_t"], steps=case["eval_steps"])
    timesteps = get_timesteps(ts[0], ts[-1], dt=case["dt"])
    n_steps = timesteps.shape[0] - 1
    torch.manual_seed(77)
    x0 = torch.randn(B, d)
    noise = philox.normal_noise(NOISE_SEED, B, n_steps, d)
    it = iter(torch.from_numpy(noise))
    orig = torch.randn

    def fake(*shape, **kw):
        n = next(it)
        assert tuple(n.shape) == tuple(shape), (n.shape, shape)
        return n

    torch.randn = fake
    try:
        xs = integ.integrate(sde, ts=ts, x_init=x0.clone())
    finally:
        torch.randn = orig
    tg = _target_params(target, d)
    from sde_sampler_b200.spec import RolloutSpec

    tgd = RolloutSpec(dim=d, ts=ts, loss={}, ctrl={}, mlp={}, gate=None, sde=None, prior=None, ref=None, target=tg).to_dict()["target"]
    return {"x0": x0.numpy().copy(), "ts": ts.numpy().copy(), "timesteps": timesteps.numpy().copy(), "target": tgd,
            "diff_coeff": float(case["diff_coeff"]), "clip_score": float(case["clip_score"]), "xs": xs.numpy().copy(),
            "torch_version": torch.__version__}


def run_ou_case(name: str, case: dict) -> dict:
    """EulerIntegrator.integrate of the unmodified reference on an OU-family SDE or a ControlledSDE around one
    (eq/integrator.py:93-127, eq/sdes.py:272-305), Philox stream injected through `torch.randn`."""
    import torch

    ref_harness.import_reference()
    from sde_sampler.distr.delta import Delta
    from sde_sampler.eq.integrator import EulerIntegrator
    from sde_sampler.eq.sdes import VP, ConstOU, ControlledSDE, ScaledBM
    from sde_sampler.utils.common import get_timesteps

    d, B = case["dim"], case["batch"]
    gen = case["generative"]
    mk = {"vp": lambda g: VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, scale_diff_coeff=1.0, terminal_t=1.0, generative=g),
          "bm_pis": lambda g: ScaledBM(diff_coeff=0.4472135954999579, terminal_t=5.0, generative=g),
          "const_ou": lambda g: ConstOU(drift_coeff=4.5, diff_coeff=3.0, terminal_t=1.0, generative=g)}[case["sde"]]
    sde = mk(gen)
    if case["ctrl"] == "pis":
        class _PIS:  # the two attributes and the one method of solver.oc.PIS that the inference process uses (:204-208)
            def __init__(self):
                self.sde, self.prior = mk(True), Delta(dim=d)

            def inference_ctrl(self, t, x):
                reference_distr = self.sde.marginal_distr(t=t, x_init=self.prior.loc)
                return self.sde.diff(t, x) * reference_distr.score(x).clip(max=1e5)

        sde = ControlledSDE(sde=sde, ctrl=_PIS().inference_ctrl)
    T_end = float(sde.terminal_t)
    ts = get_timesteps(0.0, T_end, steps=case["eval_steps"])
    if case["grid"] == "ts":
        integ, timesteps = EulerIntegrator(), ts
    else:
        integ = EulerIntegrator(dt=case["grid"])
        timesteps = get_timesteps(ts[0], ts[-1], dt=case["grid"])
    n_steps = timesteps.shape[0] - 1
    torch.manual_seed(78)
    x0 = torch.randn(B, d) * 1.5
    noise = philox.normal_noise(NOISE_SEED, B, n_steps, d)
    it = iter(torch.from_numpy(noise))
    orig = torch.randn

    def fake(*shape, **kw):
        n = next(it)
        assert tuple(n.shape) == tuple(shape), (n.shape, shape)
        return n

    torch.randn = fake
    try:
        xs = integ.integrate(sde, ts=ts, x_init=x0.clone(), timesteps=timesteps if case["grid"] == "ts" else None)
    finally:
        torch.randn = orig
    return {"x0": x0.numpy().copy(), "ts": ts.numpy().copy(), "timesteps": timesteps.numpy().copy(), "xs": xs.numpy().copy(),
            "torch_version": torch.__version__}


def main():
    if not ref_harness.available():
        raise SystemExit("reference not available; golden vectors can only be generated in the build container")
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    from oracle.cases import OU_CASES, ULA_CASES

    names = sys.argv[1:] or list(CASES) + list(ULA_CASES) + list(OU_CASES)
    for name in names:
        if name in OU_CASES:
            out = run_ou_case(name, OU_CASES[name])
            path = os.path.join(outdir, f"{name}.npz")
            specio.save(path, out)
            print(f"{name:36s} B={out['x0'].shape[0]:4d} d={out['x0'].shape[1]:3d} steps={out['timesteps'].shape[0]-1:5d} "
                  f"outputs={out['xs'].shape[0]} |xs|max={np.abs(out['xs']).max():.3e} size={os.path.getsize(path)/1024:.0f} KiB")
            continue
        if name in ULA_CASES:
            out = run_ula_case(name, ULA_CASES[name])
            path = os.path.join(outdir, f"{name}.npz")
            specio.save(path, out)
            print(f"{name:36s} B={out['x0'].shape[0]:4d} d={out['x0'].shape[1]:3d} steps={out['timesteps'].shape[0]-1:5d} "
                  f"outputs={out['xs'].shape[0]} |xs|max={np.abs(out['xs']).max():.3e} size={os.path.getsize(path)/1024:.0f} KiB")
            continue
        out = run_case(name, CASES[name])
        path = os.path.join(outdir, f"{name}.npz")
        specio.save(path, out)
        tr = out["train"]
        print(f"{name:36s} B={out['x0'].shape[0]:4d} d={out['x0'].shape[1]:3d} T={out['ts'].shape[0]-1:4d} "
              f"loss={tr['loss']:+.6e} |rnd|max={np.abs(tr['rnd']).max():.3e} "
              f"|x_T|max={np.abs(tr['x_T']).max():.3e} size={os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
