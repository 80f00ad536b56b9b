"""cProfile of the host side of training steps (method from argv[1], default lv)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from sdes_test_helpers import build_from_spec
from sde_sampler_b200.spec import ctrl_parameters

dev = torch.device("cuda:0")
W = bench.WORKLOADS["gmm50"]
o = build_from_spec(bench.load_spec(W), dev, engine="auto", seed=1234, sync_metrics=False)
x0 = bench.sample_x0(W["x0"], 65536, 50, dev, 100)
if len(sys.argv) > 1:
    o["loss"].method = sys.argv[1]
def step():
    for p in ctrl_parameters(o["ctrl"]):
        p.grad = None
    v, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
    v.backward()
    torch.cuda.synchronize()
for _ in range(3):
    step()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
