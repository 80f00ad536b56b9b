"""Where an lv training step's wall time goes, phase by phase (host clock around device syncs), to chase intermittent stalls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from sdes_test_helpers import build_from_spec
from sde_sampler_b200.spec import ctrl_parameters
from sde_sampler_b200 import engine as eng

dev = torch.device("cuda:0")
W = bench.WORKLOADS["gmm50"]
o = build_from_spec(bench.load_spec(W), dev, engine="auto", seed=1234, sync_metrics=False)
x0 = bench.sample_x0(W["x0"], 65536, 50, dev, 100)
if len(sys.argv) > 1:
    o["loss"].method = sys.argv[1]
orig = eng.rollout
def timed_rollout(*a, **k):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = orig(*a, **k)
    torch.cuda.synchronize(); print("   rollout %.2f ms" % ((time.perf_counter() - t) * 1e3))
    return r
eng.rollout = timed_rollout
orig_g = eng.lv_grad
def timed_grad(*a, **k):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = orig_g(*a, **k)
    torch.cuda.synchronize(); print("   lv_grad %.2f ms" % ((time.perf_counter() - t) * 1e3))
    return r
eng.lv_grad = timed_grad
for k in range(8):
    for p in ctrl_parameters(o["ctrl"]):
        p.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    v, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    v.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep %d: forward %.2f ms, backward %.2f ms" % (k, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
