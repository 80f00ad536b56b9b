"""One or more lv training steps (forward rollout keeping xs + tensor-core backward) of the headline workload — for
launch lists / profiles:  ncu --metrics gpu__time_duration.sum ... python tools/train_step.py [--batch B] [--reps N]"""
import argparse, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from sdes_test_helpers import build_from_spec
from sde_sampler_b200.spec import ctrl_parameters

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--method", default="lv")
args = ap.parse_args()
dev = torch.device("cuda:0")
W = bench.WORKLOADS["gmm50"]
o = build_from_spec(bench.load_spec(W), dev, engine="auto", seed=1234, sync_metrics=False)
x0 = bench.sample_x0(W["x0"], args.batch, 50, dev, 100)
o["loss"].method = args.method
ms = []
for k in range(args.reps):
    for p in ctrl_parameters(o["ctrl"]):
        p.grad = None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    v, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
    m = torch.cuda.Event(enable_timing=True); m.record()
    v.backward()
    b.record()
    torch.cuda.synchronize()
    ms.append((a.elapsed_time(m), m.elapsed_time(b)))
print("forward ms, backward ms:", ms)
