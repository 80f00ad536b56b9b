"""Freeze the cfg5 bench workload's rollout spec WITHOUT the NICE coupling weights (tests/golden/bench_cfg5_spec.npz).

bench.py's reference arm may not import the product; the fused-engine workloads read the golden fixtures, and cfg5 (whose
76 MB of coupling weights are seeded random numbers) reads this file and regenerates the weights with
tools/bench_wide.nice_target_dict.  Run here (CPU; needs the built library only to construct the loss object):

    python tools/gen_bench_specs.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]

import torch  # noqa: E402

import bench_wide  # noqa: E402
from oracle import specio  # noqa: E402
from sde_sampler_b200.spec import extract_spec  # noqa: E402


def main():
    dim, mid, hidden = 784, 1000, 5
    o = bench_wide.build(torch.device("cpu"), dim, mid, hidden, "simt")
    spec = extract_spec(o["loss"], "exp_integrator", o["ts"], o["terminal"], o["second"], train=True, compute_ito=True).to_dict()
    want = bench_wide.nice_target_dict(dim, mid, hidden)
    got = spec["target"]
    for cw, cg in zip(want["couplings"], got["couplings"]):  # the regenerated weights ARE the ones the GPU arm builds
        assert cw["mask_config"] == cg["mask_config"]
        for (w1, b1), (w2, b2) in zip(cw["layers"], cg["layers"]):
            assert (w1 == w2).all() and (b1 == b2).all()
    assert (want["scale"] == got["scale"]).all()
    spec["target"] = {"kind": "nice", "regenerate": {"dim": dim, "mid": mid, "hidden": hidden}}
    specio.save(bench_wide.SPEC_FIXTURE, spec)
    print(bench_wide.SPEC_FIXTURE, os.path.getsize(bench_wide.SPEC_FIXTURE) // 1024, "KiB")


if __name__ == "__main__":
    main()
