"""Rollout kernel time of the headline workload in its training variants: plain, + trajectory (tiled / reference layout),
+ gate_cot (lv), + score_keep (kl)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import bench
from sdes_test_helpers import build_from_spec
from sde_sampler_b200 import engine as eng
from sde_sampler_b200.engine import Workspace
from sde_sampler_b200.spec import extract_spec

dev = torch.device("cuda:0")
W = bench.WORKLOADS["gmm50"]
o = build_from_spec(bench.load_spec(W), dev, engine="auto", seed=1234, sync_metrics=False)
x0 = bench.sample_x0(W["x0"], 65536, 50, dev, 100)
ws, tb, gc, sk = Workspace(), Workspace(), Workspace(), Workspace()
def run(name, **kw):
    train = kw.pop("train", True)
    spec = extract_spec(o["loss"], "time_reversal", o["ts"], o["terminal"], o["second"], train=train, compute_ito=True, return_traj=kw.pop("traj", False))
    ms = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.rollout(spec, x0, seed=5, engine="tcgen05", workspace=ws, **kw)
        b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    print(f"{name:40s} {statistics.median(ms[1:]):.3f} ms")
run("plain")
run("traj reference layout", traj=True)
run("traj tiled", traj=True, traj_tiled=True, traj_buffer=tb)
run("traj tiled + gate_cot", traj=True, traj_tiled=True, traj_buffer=tb, gate_cot=gc, out={})
run("traj tiled + score_keep", traj=True, traj_tiled=True, traj_buffer=tb, score_keep=sk, out={})
# kl spec (no Ito term), many repetitions: every time listed (looking for outliers)
o["loss"].method = "kl"
spec = extract_spec(o["loss"], "time_reversal", o["ts"], o["terminal"], o["second"], train=True, compute_ito=False, return_traj=True)
ms = []
for _ in range(40):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.rollout(spec, x0, seed=5, engine="tcgen05", workspace=ws, traj_tiled=True, traj_buffer=tb, score_keep=sk, out={})
    b.record(); torch.cuda.synchronize(); ms.append(round(a.elapsed_time(b), 2))
print("kl forward with score_keep, 40 reps:", ms)
