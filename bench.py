#!/usr/bin/env python
"""bench.py — trajectory-steps/sec of the fused rollout on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload gmm50|cfg2|cfg3|cfg4|cfg5|gmm50dense]
                    [--scaling weak|strong] [--global-batch B] [--engine auto|tcgen05|simt]

Default workload `gmm50` (north-star headline): GMM-40 d=50, DIS + log-variance loss, T=100, 65 536 trajectories per GPU
(weak scaling: every rank carries its own shard of the global batch; the only exchange is the 8-double statistics
all-gather inside the loss).  One "step" = one training-mode call of the loss plug-in,
`loss(ts, x0, clipped_target_unnorm_log_prob, prior.log_prob)`: prologue kernel + persistent rollout kernel over all T time
steps + statistics kernel.  The other workloads are BASELINE.json's configs (cfg2, cfg3, cfg4 = gmm50 at 32 768 per GPU,
cfg5 on the wide engine) and `gmm50dense`, the headline with a 40-mode mixture that differs in every dimension.

`value`   : N*B*T / time with x0 resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`     : same call, but every step's x0 starts in pinned host memory and the loss scalar is read back each step.
`roofline`: MLP FLOPs (F(d) = 256 d + 16 384 per trajectory-step) of one rollout launch over its CUDA-event
            duration, against the measured dense bf16 peak.
`cpu_baseline` / `--impl reference`: oracle/torch_port.py — the reference's per-step op sequence in torch eager on the
            host cores — on a bounded sample of the same workload.  The reference arm imports NOTHING from the product
            package: the workload's spec comes from the golden fixture frozen from the unmodified reference.
`gpu_eager_baseline`: the same op sequence as stock eager PyTorch kernels on the same GPU (fp32, TF32 off) — the
            "existing Blackwell path" of SURVEY §8d.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# NCCL prints its version banner on stdout (C-level write at the first communicator); the contract is ONE JSON line
# there, so file descriptor 1 points at stderr until the line is printed.
_SAVED_STDOUT = None


def _stdout_to_stderr():
    global _SAVED_STDOUT
    if _SAVED_STDOUT is None:
        sys.stdout.flush()
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    global _SAVED_STDOUT
    sys.stdout.flush()
    if _SAVED_STDOUT is not None:
        os.dup2(_SAVED_STDOUT, 1)
        os.close(_SAVED_STDOUT)
        _SAVED_STDOUT = None
    print(json.dumps(line), flush=True)



ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "trajectory-steps/sec"
UNIT = "traj-steps/s"

# Every fused-engine workload is a golden fixture's spec (tests/golden/<golden>.npz: the raw parameters extracted from the
# UNMODIFIED reference objects by oracle/gen_golden.py) run at the BASELINE batch size; both arms read the same file.
WORKLOADS = {
    "gmm50": dict(golden="dis_gmm50_lv", batch=65536, x0="trunc_gauss",
                  what="GMM-40 d=50 solver=dis loss=lv T=100", note="north-star headline (BASELINE configs[3] at 65 536 per GPU)"),
    "cfg2": dict(golden="dis_gmm2_lv", batch=65536, x0="gauss", what="GMM-40 d=2 solver=basic_dis loss=lv T=100", note="BASELINE configs[1]"),
    "cfg3": dict(golden="pis_funnel10_kl", batch=65536, x0="zeros", what="funnel d=10 solver=basic_pis loss=kl T=200", note="BASELINE configs[2]"),
    "cfg4": dict(golden="dis_gmm50_lv", batch=32768, x0="trunc_gauss", what="GMM-40 d=50 solver=dis loss=lv T=100",
                 note="BASELINE configs[3] as written: 262 144 trajectories over 8 GPUs = 32 768 per GPU"),
    "gmm50dense": dict(golden="dis_gmmdense50_lv", batch=65536, x0="trunc_gauss", what="dense GMM-40 (modes differ in all 50 dims) d=50 solver=dis loss=lv T=100",
                       note="the headline without GMM-40's zero-padded structure: no dimension factors out of the mixture"),
}


def flops_per_traj_step(d: int) -> int:
    return 256 * d + 16384  # x-dependent FourierMLP layers only, C=64, 4 layers


def workload_name(w: dict, batch: int) -> str:
    return f"{w['what']} batch={batch}/GPU"


def load_spec(w: dict) -> dict:
    from oracle import specio

    return specio.load(os.path.join(ROOT, "tests", "golden", w["golden"] + ".npz"))["spec"]


def sample_x0(kind: str, batch: int, dim: int, device, seed: int):
    """The caller-side prior draw (solver/oc.py:71): conf/prior/gauss_truncate.yaml (trunc_normal_ at the 1e-4 quantiles),
    conf/prior/gauss.yaml, or the Delta prior of PIS (distr/delta.py:25-28)."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "zeros":
        return torch.zeros(batch, dim, device=device)
    x = torch.randn(batch, dim, generator=g)
    if kind == "trunc_gauss":
        x = x.clamp_(-3.7190, 3.7190)  # standard normal quantiles of 1e-4 / 1 - 1e-4: same support as trunc_normal_
    return x.to(device)


def recorded_traffic(workload: str, engine: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel from the committed `ncu --set full` capture of
    this same command (profiles/traffic.json; one capture per (workload, engine, batch)), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = rec.get(f"{workload}:{engine}:B={batch}")
        return (None, None) if e is None else (e["dram_bytes_per_launch"], e.get("source"))
    except Exception:
        return None, None


# ------------------------------------------------------------------------------ CPU / eager baselines
def time_port(spec: dict, x0, device: str, repeats: int):
    """oracle/torch_port.py: the reference's op sequence (losses/oc.py:176-222 and what it calls) in torch eager."""
    import torch

    from oracle import torch_port

    times = []
    for _ in range(repeats):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch_port.rollout(spec, x0, device=device, as_numpy=False)
        if device != "cpu":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return times


def cpu_sample_size(spec, w, dim, budget_s: float, total_steps: int, batch_cap: int) -> int:
    T = len(spec["ts"]) - 1
    probe_b = 256
    probe = min(time_port(spec, sample_x0(w["x0"], probe_b, dim, "cpu", 0), "cpu", 2))
    speed = probe_b * T / probe
    sample = int(speed * budget_s / max(1, total_steps) / T)
    return max(64, min(batch_cap, sample // 64 * 64))


def reference_arm(args, w):
    """--impl reference: the CPU port on all host threads, bounded sample per step.  No product import."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    assert "sde_sampler_b200" not in sys.modules
    torch.set_num_threads(os.cpu_count() or 1)
    spec = load_spec(w)
    dim, T = int(spec["dim"]), len(spec["ts"]) - 1
    B = args.batch or w["batch"]
    total_steps = args.steps + args.warmup
    sample = cpu_sample_size(spec, w, dim, 150.0, total_steps, B)
    x0 = sample_x0(w["x0"], sample, dim, "cpu", 0)
    times = time_port(spec, x0, "cpu", total_steps)[args.warmup:]
    dt = sum(times) / len(times)
    value = sample * T / dt
    assert "sde_sampler_b200" not in sys.modules, "the reference arm must not import the product"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(w, B), "sample": f"{sample} of {B} trajectories x T={T} per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"oracle/torch_port.py (reference op sequence incl. the autograd target score, torch eager fp32, under "
                                   f"no_grad: faster than the reference's train-mode call, so ratios against it are conservative) on {sample} "
                                   f"trajectories x T={T}, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms DURING the timed region (NVML in a thread;
    falls back to `nvidia-smi -lms` when pynvml is unavailable)."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_clock = None
        self._stop = False
        self._thread = None
        self._smi = None

    def _nvml_loop(self, pynvml, handle):
        R = {
            getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop:
            try:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for bit, name in R.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the local CUDA index to the NVML index via the PCI bus id
            import torch

            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            handle = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        handle = h
                        break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_clock = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._nvml_loop, args=(pynvml, handle), daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            self._smi_f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                              "-i", str(self.idx)], stdout=self._smi_f, stderr=subprocess.DEVNULL)
            except Exception:
                self._smi = None

    def stop(self) -> dict:
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2)
        elif self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            self._smi_f.flush()
            self._smi_f.seek(0)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self._smi_f.read().strip().splitlines():
                parts = [x.strip() for x in ln.split(",")]
                try:
                    self.samples.append(float(parts[0]))
                    self.max_clock = float(parts[1])
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            os.unlink(self._smi_f.name)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_clock, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_clock, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}




# ------------------------------------------------------------------------------ cfg5 (wide engine)
def cfg5_reference_spec(steps: int) -> dict:
    """cfg5's spec for the CPU port, without the product: committed fixture + regenerated NICE weights (tools/gen_bench_specs.py)."""
    import bench_wide
    from oracle import specio

    spec = specio.load(bench_wide.SPEC_FIXTURE)
    spec["target"] = bench_wide.nice_target_dict(**spec["target"]["regenerate"])
    spec["ts"] = spec["ts"][: steps + 1]
    return spec


def main_cfg5(args):
    """BASELINE configs[4] on ONE GPU's shard: nice/mnist d=784, DDS + lv, T=257 (cosine grid), 4 096 trajectories
    per GPU (32 768 over 8).  Same JSON keys as the headline line; `roofline` counts the Linear layers of the control
    MLP and of the NICE forward + input-gradient backward (7.68e7 FLOP per trajectory-step, 2 per multiply-add; the
    three bf16 passes of the split-precision GEMM are NOT counted) against the measured dense bf16 peak."""
    import torch

    import bench_wide

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch or 4096
    dim, mid, hidden = 784, 1000, 5
    name = f"NICE d={dim} mid={mid} hidden={hidden} solver=dds loss=lv T=257 batch={B}/GPU"
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import torch_port

        torch.set_num_threads(os.cpu_count() or 1)
        sample, T = 64, 4
        spec = cfg5_reference_spec(T)
        x0 = sample_x0("gauss", sample, dim, "cpu", 0)
        times = []
        for _ in range(args.steps + args.warmup):
            t0 = time.perf_counter()
            torch_port.rollout(spec, x0, generator=torch.Generator().manual_seed(0))
            times.append(time.perf_counter() - t0)
        dt = sum(times[args.warmup:]) / max(1, len(times[args.warmup:]))
        value = sample * T / dt
        assert "sde_sampler_b200" not in sys.modules, "the reference arm must not import the product"
        emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
              "dtype": "f32", "data": "synthetic",
              "config": {"workload": name, "sample": f"{sample} trajectories x {T} of 257 time steps per step"},
              "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"oracle/torch_port.py on {sample} trajectories x {T} time steps"},
              "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    import torch.distributed as dist

    from sde_sampler_b200 import _cabi

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    lib = _cabi.lib()
    o = bench_wide.build(device, dim, mid, hidden, "auto" if args.engine == "auto" else args.engine)
    loss, ts = o["loss"], o["ts"]
    loss.process_group = pg
    T = ts.shape[0] - 1
    torch.manual_seed(100 + rank)
    x0 = o["prior"].sample((B,))
    x0_host = x0.cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    steps = min(args.steps, 5)
    torch.set_grad_enabled(False)  # the metric is the forward rollout; with grad the loss keeps per-step state images
    for _ in range(3):
        loss(ts, x0, o["terminal"], o["second"])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = lib.sdes_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        val, _m = loss(ts, x0, o["terminal"], o["second"])
    b.record()
    barrier()
    launches = lib.sdes_launch_count() - n0
    clocks = sampler.stop()
    total_ms = a.elapsed_time(b)
    a.record()
    for _ in range(steps):
        v, _m = loss(ts, x0_host.to(device, non_blocking=True), o["terminal"], o["second"])
        host_loss = v.item()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    times = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = times.tolist()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
        f = bench_wide.flops_per_traj_step(dim, mid, hidden)
        ts_per_s = world * B * T * steps / (total_ms * 1e-3)
        ach = B * T * steps * f / (total_ms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": ts_per_s, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 3,
                "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (every Linear as 3 bf16 tcgen05 passes over hi/lo-split operands, fp32 accumulate; f32 elsewhere)",
                "data": "synthetic",
                "config": {"workload": name, "engine": "wide/tcgen05",
                           "global_batch": world * B, "l2": "working set per step (weights 153 MB + activations) exceeds the 126 MB L2",
                           "noise": "in-kernel Philox4x32-10", "loss_value": float(val)},
                "clocks": clocks,
                "e2e": {"value": world * B * T * steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * dim * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms / steps, "loss_value": host_loss},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                             "flops_per_traj_step": f, "peak_source": "MEASURED_PEAKS.json bf16 dense, sustained (a step is thousands of GEMM launches)"}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import torch_port

            torch.set_num_threads(os.cpu_count() or 1)
            sample, Tc = 64, 4
            spec = cfg5_reference_spec(Tc)
            t0 = time.perf_counter()
            torch_port.rollout(spec, sample_x0("gauss", sample, dim, "cpu", 0), generator=torch.Generator().manual_seed(0))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample * Tc / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"oracle/torch_port.py on {sample} trajectories x {Tc} of 257 time steps, one pass, {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "tcgen05", "simt"])
    ap.add_argument("--batch", type=int, default=None, help="trajectories per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch trajectories per GPU; strong: --global-batch trajectories split over the ranks")
    ap.add_argument("--global-batch", type=int, default=65536, help="total trajectories with --scaling strong")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--workload", default="gmm50", choices=list(WORKLOADS) + ["cfg5"],
                    help="gmm50 = north-star headline (default, the driver's line); cfg2 / cfg3 / cfg4 / cfg5 = BASELINE configs[1..4] "
                         "per-GPU shards; gmm50dense = the headline with a mixture that differs in every dimension")
    args = ap.parse_args()
    _stdout_to_stderr()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "cfg5":
        return main_cfg5(args)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, w)
        return

    import torch
    import torch.distributed as dist

    from sde_sampler_b200 import _cabi
    from sdes_test_helpers import build_from_spec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    if args.gpus != world:
        if rank == 0:
            print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    lib = _cabi.lib()
    if args.scaling == "strong":
        if args.global_batch % world:
            raise SystemExit("--global-batch must be divisible by the number of ranks")
        B = args.global_batch // world
    else:
        B = args.batch or w["batch"]
    spec_dict = load_spec(w)
    dim, T = int(spec_dict["dim"]), len(spec_dict["ts"]) - 1
    # resident-input leg: the filtered-trajectory count stays on the device (no host sync per call), so calls
    # queue back to back; the e2e leg below reads the loss scalar back every step.
    o = build_from_spec(spec_dict, device, engine=args.engine, process_group=pg, seed=1234, sync_metrics=False)
    loss, ts = o["loss"], o["ts"]
    x0 = sample_x0(w["x0"], B, dim, device, 100 + rank)
    x0_host = [x0.cpu().pin_memory(), x0.cpu().pin_memory()]
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def step(x):
        # the metric is the forward rollout; the training step with its backward is timed separately below
        with torch.no_grad():
            return loss(ts, x, o["terminal"], o["second"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # which engine did the descriptor resolve to?
    from sde_sampler_b200 import engine as eng
    from sde_sampler_b200.spec import extract_spec
    method = spec_dict["loss"]["method"]
    spec = extract_spec(loss, spec_dict["loss"]["kind"], ts, o["terminal"], o["second"], train=True, compute_ito=method != "kl")
    desc, _ = eng.fill_desc(spec, batch=B, engine=args.engine)
    engine_used = "simt" if desc.flags & _cabi.F_MLP_SIMT else "tcgen05"

    for _ in range(args.warmup):
        val, _m = step(x0)
        flush.zero_()
    barrier()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.sdes_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        val, _m = step(x0)
        ev[k][1].record()
        flush.zero_()  # L2 flush between timed iterations (inside the region; ~0.03 ms)
    t_end.record()
    barrier()
    launches = lib.sdes_launch_count() - launches0
    clocks = sampler.stop()
    total_ms = t_start.elapsed_time(t_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    loss_value = float(val)

    # ---- rollout kernel alone (prologue + persistent kernel; CUDA events around the C-ABI call)
    k_ms = []
    for _ in range(min(args.steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        eng.rollout(spec, x0, seed=99, engine=args.engine, workspace=loss._workspace)
        b.record()
        torch.cuda.synchronize(device)
        k_ms.append(a.elapsed_time(b))
    kernel_ms = statistics.median(k_ms)

    # ---- training step (headline only): loss(...) with grad + loss.backward() (csrc/sdes_grad.cu)
    from sde_sampler_b200.spec import ctrl_parameters
    train_ms = kl_ms = full_ms = None
    headline = args.workload == "gmm50"
    if headline:
        try:
            tm = []
            for k in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for prm in ctrl_parameters(o["ctrl"]):
                    prm.grad = None
                a.record()
                v, _m = loss(ts, x0, o["terminal"], o["second"])
                v.backward()
                b.record()
                torch.cuda.synchronize(device)
                tm.append(a.elapsed_time(b))
            train_ms = statistics.median(tm[2:])
        except Exception as exc:  # reported, never hidden
            train_ms = f"failed: {type(exc).__name__}: {exc}"

    # ---- timed region 2: end to end through the plug-in with host buffers.  The loop a user writes: the NEXT step's x0 goes
    #      host -> device on a copy stream while the current rollout runs; every step still copies its own input inside the
    #      region and reads its own loss back.
    barrier()
    copy_stream = torch.cuda.Stream(device)
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    host_loss = 0.0
    with torch.cuda.stream(copy_stream):
        xd_next = x0_host[0].to(device, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(copy_stream)
    # every step's loss goes device -> pinned host memory with a non-blocking copy and is READ one step later (after the next
    # step has been queued), the last one before the region ends: each step's result reaches the host inside the region, but
    # the host never idles the GPU waiting for it — the loop a training script with a logging callback runs
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [None, None]
    pending = None
    for k in range(args.steps):
        torch.cuda.current_stream(device).wait_event(ready)
        xd = xd_next
        v, _m = step(xd)
        loss_host[k & 1].copy_(v.detach().reshape(()).float(), non_blocking=True)
        loss_done[k & 1] = torch.cuda.Event()
        loss_done[k & 1].record()
        if k + 1 < args.steps:
            with torch.cuda.stream(copy_stream):
                xd_next = x0_host[(k + 1) & 1].to(device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
        xd.record_stream(torch.cuda.current_stream(device))
        if pending is not None:  # the previous step's loss: on the host by now
            loss_done[pending].synchronize()
            host_loss = float(loss_host[pending])
        pending = k & 1
    loss_done[pending].synchronize()
    host_loss = float(loss_host[pending])  # device -> host read of the last step's result
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)

    # ---- after the headline regions (the optimizer below changes the weights): the same workload with loss.method = kl
    #      (backpropagation through time), and the complete training iteration of Trainable.step (solver/base.py:399-454)
    #      with the fused optimizer tail (csrc/sdes_trainer.cu)
    if headline:
        try:
            loss.method = "kl"
            tm = []
            for k in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for prm in ctrl_parameters(o["ctrl"]):
                    prm.grad = None
                a.record()
                v, _m = loss(ts, x0, o["terminal"], o["second"])
                v.backward()
                b.record()
                torch.cuda.synchronize(device)
                tm.append(a.elapsed_time(b))
            kl_ms = statistics.median(tm[2:])
        except Exception as exc:
            kl_ms = f"failed: {type(exc).__name__}: {exc}"
        finally:
            loss.method = "lv"
        try:
            from sde_sampler_b200 import FusedAdamEMA
            opt = FusedAdamEMA(ctrl_parameters(o["ctrl"]), lr=0.005, weight_decay=1e-7, grad_clip_norm=1.0,
                               ema=dict(decay=0.9999, inv_gamma=1.0, power=0.9, update_after_step=2, update_every=1))
            tm = []
            for k in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                opt.zero_grad()
                v, _m = loss(ts, x0, o["terminal"], o["second"])
                v = v * (1.0 / dim)  # scale_loss (conf/solver/oc_base.yaml:22)
                v.backward()
                opt.step(loss=v)
                b.record()
                torch.cuda.synchronize(device)
                tm.append(a.elapsed_time(b))
            full_ms = statistics.median(tm[1:])
        except Exception as exc:
            full_ms = f"failed: {type(exc).__name__}: {exc}"

    times = torch.tensor([total_ms, e2e_ms, kernel_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms = times.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops (burst), of measured" if "bf16_tflops" in peaks else "1590 TFLOP/s, of fallback"
        traj_steps = B * T
        value = world * traj_steps * args.steps / (total_ms * 1e-3)
        e2e_value = world * traj_steps * args.steps / (e2e_ms * 1e-3)
        achieved_tf = traj_steps * flops_per_traj_step(dim) / (kernel_ms * 1e-3) / 1e12
        traffic, traffic_src = recorded_traffic(args.workload, engine_used, B)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if engine_used == "simt" else "f32 (MLP as 3 bf16 tcgen05 MMA passes over hi/lo-split operands with fp32 accumulate, fp32-equivalent; f32 elsewhere)",
            "data": "synthetic",
            "config": {"workload": workload_name(w, B), "note": w["note"], "engine": engine_used, "global_batch": world * B,
                       "l2": "flushed between timed steps (192 MiB memset inside the region)",
                       "noise": "in-kernel Philox4x32-10", "loss_value": loss_value,
                       "spec": f"tests/golden/{w['golden']}.npz (parameters extracted from the unmodified reference objects)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * dim * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps, "loss_value": host_loss,
                    "how": "x0 of step k+1 copied from pinned host memory on a side stream while step k runs; every step's loss copied to pinned host memory (non-blocking) and read one step later, the last one before the region ends"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": kernel_ms,
                         "flops_per_traj_step": flops_per_traj_step(dim), "peak_source": peak_src,
                         "traj_steps_per_s_kernel": traj_steps / (kernel_ms * 1e-3),
                         "hbm_algorithmic_bytes": B * (8 * dim + 4),
                         "hbm_gbs": B * (8 * dim + 4) / (kernel_ms * 1e-3) / 1e9},
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
        }
        if headline:
            line["train_step"] = {
                "what": "loss(ts, x0, ...) with grad + loss.backward(): rollout keeping xs, then the lv gradient (one persistent kernel)", "ms": train_ms,
                "traj_steps_per_s": (world * traj_steps / (train_ms * 1e-3)) if isinstance(train_ms, float) else None,
                "kl_ms": kl_ms, "kl_what": "same workload with loss.method=kl: rollout keeping xs and the score part, then backpropagation through time as a discrete adjoint (one persistent kernel)",
                "full_iteration_ms": full_ms,
                "full_iteration_what": "zero_grad + lv loss + backward + sdes_trainer_step (grad check, clip_grad_norm_, Adam, EMA), no host sync inside"}
        if world == 1 and not args.no_cpu_baseline:
            # the existing-Blackwell bar: the reference's op sequence as stock eager PyTorch kernels on this GPU (fp32, TF32 off)
            try:
                torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.allow_tf32 = False
                be = min(B, 65536)
                xe = x0[:be]
                te = time_port(spec_dict, xe, str(device), 2)
                line["gpu_eager_baseline"] = {
                    "value": be * T / te[-1], "unit": UNIT, "ms": te[-1] * 1e3, "batch": be,
                    "what": "oracle/torch_port.py (losses/oc.py:176-222 op sequence incl. the autograd GMM score; torch eager, fp32, "
                            "TF32 off, under no_grad) on the same B200, second of two passes, wall clock around a device sync"}
            except Exception as exc:
                line["gpu_eager_baseline"] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
            torch.set_num_threads(os.cpu_count() or 1)
            sample = cpu_sample_size(spec_dict, w, dim, 15.0, 1, B)
            dt = min(time_port(spec_dict, sample_x0(w["x0"], sample, dim, "cpu", 0), "cpu", 1))
            line["cpu_baseline"] = {"value": sample * T / dt, "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": f"oracle/torch_port.py (reference op sequence, torch eager fp32, under no_grad: a lower bound on the "
                                              f"reference's train-mode cost) on {sample} of {B} trajectories x T={T}, one pass, {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
