"""Ad-hoc timing of the other BASELINE.json configurations and of eval mode (not a bench line):
python tools/probe_configs.py   (needs a GPU)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import bench


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    o = bench.build_objects(dev, "auto")
    loss, ts = o["loss"], o["ts"]
    B = 65536
    x0 = o["prior"].sample((B,))
    ms = timeit(lambda: loss(ts, x0, o["terminal"], o["second"]))
    print(f"cfg4/headline train call: {ms:.3f} ms  {B*100/ms/1e3:.3e} traj-steps/s")
    ms = timeit(lambda: loss.eval(ts, x0, o["terminal"], o["second"], compute_weights=False, return_traj=False))
    print(f"eval (no traj, no weights): {ms:.3f} ms  {B*100/ms/1e3:.3e} traj-steps/s")
    ms = timeit(lambda: loss.eval(ts, x0, o["terminal"], o["second"], compute_weights=True, return_traj=True))
    print(f"eval (return_traj, weights): {ms:.3f} ms  {B*100/ms/1e3:.3e} traj-steps/s  (xs = {101*B*50*4/1e9:.2f} GB)")
    # the golden-case configurations at full batch
    from oracle import specio
    from sdes_test_helpers import build_from_spec

    for name, Bc in [("dis_gmm2_lv", 65536), ("pis_funnel10_kl", 65536), ("dds_funnel10_lv", 65536), ("dis_dw1_lv", 65536)]:
        g = specio.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        for engine in ("tcgen05", "simt"):
            b = build_from_spec(g["spec"], dev, engine=engine)
            d = g["spec"]["dim"]
            T = g["ts"].shape[0] - 1
            x = torch.randn(Bc, d, device=dev) if name != "pis_funnel10_kl" else torch.zeros(Bc, d, device=dev)
            ms = timeit(lambda: b["loss"](b["ts"], x, b["terminal"], b["second"]))
            print(f"{name:20s} {engine:8s} B={Bc} T={T}: {ms:.3f} ms  {Bc*T/ms/1e3:.3e} traj-steps/s")


if __name__ == "__main__":
    main()
