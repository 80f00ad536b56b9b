"""Runs two rollouts of the headline workload so that a library built with -DSDES_TC_TIMELINE prints its per-step timelines
(lines `TLSTEP <step> cta <c> g <group>: cumulative cycles ...`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, bench
from sdes_test_helpers import build_from_spec
dev = torch.device("cuda:0")
W = bench.WORKLOADS["gmm50"]
o = build_from_spec(bench.load_spec(W), dev, engine="auto", seed=1234, sync_metrics=False)
x0 = bench.sample_x0(W["x0"], 65536, 50, dev, 100)
with torch.no_grad():
    for _ in range(2):
        v, _ = o["loss"](o["ts"], x0, o["terminal"], o["second"])
torch.cuda.synchronize()
