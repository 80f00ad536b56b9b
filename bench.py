#!/usr/bin/env python
"""bench.py — trajectory-steps/sec of the fused rollout on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--engine auto|tcgen05|simt]

Workload (north-star headline): GMM-40 d=50, DIS + log-variance loss, T=100, 65 536 trajectories
per GPU (weak scaling: every rank carries its own shard of the global batch; the only exchange is
the 8-double statistics all-gather inside the loss).  One "step" = one training-mode call of the
loss plug-in, `loss(ts, x0, clipped_target_unnorm_log_prob, prior.log_prob)`: prologue kernel +
persistent rollout kernel over all T time steps + statistics kernel.

`value`   : N*B*T / time with x0 resident in HBM (CUDA events on the launching stream, max over ranks).
`e2e`     : same call, but x0 starts in pinned host memory and the loss scalar is read back each step.
`roofline`: MLP FLOPs (F(d) = 256 d + 16 384 per trajectory-step) of one rollout launch
            over its CUDA-event duration, against the measured dense bf16 peak.
`cpu_baseline` / `--impl reference`: oracle/torch_port.py — the reference's per-step op sequence in
            torch eager on the host cores — on a bounded sample of the same workload.
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
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

DIM, T_STEPS, BATCH_PER_GPU, N_MODES = 50, 100, 65536, 40
METRIC = "trajectory-steps/sec"
UNIT = "traj-steps/s"


def flops_per_traj_step(d: int) -> int:
    return 256 * d + 16384  # x-dependent FourierMLP layers only, C=64, 4 layers


def workload_name(batch: int) -> str:
    return f"GMM-40 d={DIM} solver=dis loss=lv T={T_STEPS} batch={batch}/GPU"


def recorded_traffic(engine: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the rollout kernel from the committed
    `ncu --set full` capture of this same command (profiles/traffic.json), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = rec.get(f"{engine}:B={batch}")
        return None if e is None else e["dram_bytes_per_launch"]
    except Exception:
        return None


# ----------------------------------------------------------------------------- the objects
def build_objects(device, engine: str, process_group=None, seed: int = 1, sync_metrics: bool = True):
    """conf/solver/dis.yaml with target GMM-40 d=50 (explicit loc), random-init
    weights of the reference architecture, out layers re-randomised (the default zero init makes
    NN == 0 and the MLP trivial)."""
    import torch
    from torch import nn
    from functools import partial

    import ref_mirrors as plugins  # parameter-holder mirrors of the reference classes (tests/ref_mirrors.py)
    from sde_sampler_b200 import FusedTimeReversalLoss

    torch.manual_seed(seed)
    loc, scale, w = plugins.fab_gmm_params(DIM)
    target = plugins.GMM(dim=DIM, loc=loc, scale=scale, mixture_weights=w)
    prior = plugins.IsotropicGauss(dim=DIM, truncate_quartile=1e-4)  # conf/prior/gauss_truncate.yaml
    sde = plugins.VP(diff_coeff_sq_min=0.1, diff_coeff_sq_max=10.0, terminal_t=1.0)  # conf/sde/vp_10.yaml
    base = plugins.FourierMLP(dim=DIM, num_layers=4, channels=64)
    gate = plugins.TimeEmbed(dim_out=1, num_layers=4, channels=64, last_bias_init=partial(nn.init.constant_, val=1.0))
    with torch.no_grad():
        base.out_layer.weight.normal_(0.0, 0.15)
        base.out_layer.bias.normal_(0.0, 0.1)
        gate.out_layer.weight.normal_(0.0, 0.05)
    ctrl = plugins.LerpCtrl(base_model=base, clip_model=10.0, target_score=target.score, score_model=gate,
                            detach_score=False, scale_score=1.0, clip_score=10.0, sde=sde, prior_score=prior.score)
    for m in (target, prior, sde, base, gate):
        m.to(device)
    loss = FusedTimeReversalLoss(generative_ctrl=ctrl, sde=sde, method="lv", max_rnd=1e8, engine=engine,
                                 process_group=process_group, seed=1234, sync_metrics=sync_metrics)

    class Solver:  # owner of clipped_target_unnorm_log_prob (solver/oc.py:48-54)
        def __init__(self):
            self.target, self.clip_target = target, None

        def clipped_target_unnorm_log_prob(self, x):
            raise RuntimeError("introspected, never called")

    ts = plugins.get_timesteps(0.0, 1.0, steps=T_STEPS).to(device)
    return dict(loss=loss, ts=ts, terminal=Solver().clipped_target_unnorm_log_prob, second=prior.log_prob,
                prior=prior, target=target, sde=sde, ctrl=ctrl)


# ------------------------------------------------------------------------------ CPU baseline
def cpu_port_setup():
    """Spec dict of the same workload for oracle/torch_port.py (CPU tensors)."""
    import torch

    from sde_sampler_b200.spec import extract_spec

    o = build_objects_cpu()
    spec = extract_spec(o["loss"], "time_reversal", o["ts"], o["terminal"], o["second"], train=True, compute_ito=True)
    return spec.to_dict(), o


def build_objects_cpu():
    import torch

    from sde_sampler_b200 import _cabi

    # the mirrors hold parameters only; building them on the CPU touches no kernel
    return build_objects(torch.device("cpu"), "simt")


def time_cpu_port(batch: int, repeats: int = 1):
    import torch

    from oracle import torch_port

    spec, o = cpu_port_setup()
    torch.manual_seed(0)
    x0 = o["prior"].sample((batch,))
    gen = torch.Generator().manual_seed(0)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        torch_port.rollout(spec, x0.numpy(), generator=gen)
        times.append(time.perf_counter() - t0)
    return times


def reference_arm(args):
    """--impl reference: the CPU port on all host threads, bounded sample per step."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe_b = 256
    probe = min(time_cpu_port(probe_b, repeats=2))
    speed = probe_b * T_STEPS / probe
    budget_s = 150.0
    total_steps = args.steps + args.warmup
    sample = int(speed * budget_s / total_steps / T_STEPS)
    sample = max(64, min(BATCH_PER_GPU, sample // 64 * 64))
    times = time_cpu_port(sample, repeats=total_steps)[args.warmup:]
    dt = sum(times) / len(times)
    value = sample * T_STEPS / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(BATCH_PER_GPU), "sample": f"{sample} of {BATCH_PER_GPU} trajectories x T={T_STEPS} per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"oracle/torch_port.py (reference op sequence, torch eager fp32, no_grad) on {sample} trajectories x T={T_STEPS}, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


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
def main_cfg5(args):
    """BASELINE configs[4] on ONE GPU's shard: nice/mnist d=784, DDS + lv, T=257 (cosine grid), 4 096 trajectories
    per GPU (32 768 over 8).  Same JSON keys as the headline line; `roofline` counts the Linear layers of the control
    MLP and of the NICE forward + input-gradient backward (7.68e7 FLOP per trajectory-step, 2 per multiply-add; the
    three bf16 passes of the split-precision GEMM are NOT counted) against the measured dense bf16 peak."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_wide

    from sde_sampler_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = 4096 if args.batch == BATCH_PER_GPU else args.batch
    dim, mid, hidden = 784, 1000, 5
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import torch_port
        from sde_sampler_b200.spec import extract_spec

        torch.set_num_threads(os.cpu_count() or 1)
        o = bench_wide.build(torch.device("cpu"), dim, mid, hidden, "simt", steps=4)
        spec = extract_spec(o["loss"], "exp_integrator", o["ts"], o["terminal"], o["second"], train=True, compute_ito=True).to_dict()
        sample, T = 64, 4
        x0 = o["prior"].sample((sample,))
        times = []
        for _ in range(args.steps + args.warmup):
            t0 = time.perf_counter()
            torch_port.rollout(spec, x0.numpy(), generator=torch.Generator().manual_seed(0))
            times.append(time.perf_counter() - t0)
        dt = sum(times[args.warmup:]) / max(1, len(times[args.warmup:]))
        value = sample * T / dt
        emit({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"NICE d={dim} mid={mid} hidden={hidden} solver=dds loss=lv T=257 batch={B}/GPU",
                                     "sample": f"{sample} trajectories x {T} of 257 time steps per step"},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                           "sample": f"oracle/torch_port.py on {sample} trajectories x {T} time steps"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
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
    for _ in range(2):
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
        line = {"metric": METRIC, "value": ts_per_s, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 2,
                "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (every Linear as 3 bf16 tcgen05 passes over hi/lo-split operands, fp32 accumulate; f32 elsewhere)",
                "data": "synthetic",
                "config": {"workload": f"NICE d={dim} mid={mid} hidden={hidden} solver=dds loss=lv T={T} batch={B}/GPU", "engine": "wide/tcgen05",
                           "global_batch": world * B, "l2": "working set per step (weights 153 MB + activations) exceeds the 126 MB L2",
                           "noise": "in-kernel Philox4x32-10", "loss_value": float(val)},
                "clocks": clocks,
                "e2e": {"value": world * B * T * steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * dim * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms / steps, "loss_value": host_loss},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                             "flops_per_traj_step": f, "peak_source": "MEASURED_PEAKS.json bf16 dense, sustained (a step is thousands of GEMM launches)"}}
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
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="gmm50", choices=["gmm50", "cfg5"],
                    help="gmm50 = north-star headline (default, the driver's line); cfg5 = BASELINE configs[4] per-GPU shard "
                         "(NICE d=784 mid=1000, DDS+lv, T=257, 4096 trajectories per GPU) on the wide engine")
    args = ap.parse_args()
    _stdout_to_stderr()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "cfg5":
        return main_cfg5(args)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from sde_sampler_b200 import _cabi

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
    B = args.batch
    # resident-input leg: the filtered-trajectory count stays on the device (no host sync per call), so calls
    # queue back to back; the e2e leg below reads the loss scalar back every step.
    o = build_objects(device, args.engine, process_group=pg, sync_metrics=False)
    loss, ts = o["loss"], o["ts"]
    torch.manual_seed(100 + rank)
    x0 = o["prior"].sample((B,))
    x0_host = x0.cpu().pin_memory()
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
    spec = extract_spec(loss, "time_reversal", ts, o["terminal"], o["second"], train=True, compute_ito=True)
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
        flush.zero_()  # L2 flush between timed iterations (inside the region; ~0.1 ms)
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

    # ---- training step: loss(...) with grad + loss.backward() (lv gradient on the tensor cores, csrc/sdes_grad.cu)
    from sde_sampler_b200.spec import ctrl_parameters
    train_ms = None
    try:
        tm = []
        for k in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for prm in ctrl_parameters(o["ctrl"]):
                prm.grad = None
            a.record()
            v, _m = loss(ts, x0, o["terminal"], o["second"])
            v.backward()
            b.record()
            torch.cuda.synchronize(device)
            tm.append(a.elapsed_time(b))
        train_ms = statistics.median(tm[1:])
    except Exception as exc:  # reported, never hidden
        train_ms = f"failed: {type(exc).__name__}: {exc}"

    # ---- timed region 2: end to end through the plug-in with host buffers
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    host_loss = 0.0
    for k in range(args.steps):
        xd = x0_host.to(device, non_blocking=True)
        v, _m = step(xd)
        host_loss = v.item()  # device -> host read of the step's result
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)

    # ---- after the headline regions (the optimizer below changes the weights): the same workload with loss.method = kl
    #      (backpropagation through time: reverse sweep csrc/sdes_adjoint.cu + the GEMM passes), and the complete training
    #      iteration of Trainable.step (solver/base.py:399-454) with the fused optimizer tail (csrc/sdes_trainer.cu)
    kl_ms, full_ms = None, None
    try:
        loss.method = "kl"
        tm = []
        for k in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for prm in ctrl_parameters(o["ctrl"]):
                prm.grad = None
            a.record()
            v, _m = loss(ts, x0, o["terminal"], o["second"])
            v.backward()
            b.record()
            torch.cuda.synchronize(device)
            tm.append(a.elapsed_time(b))
        kl_ms = statistics.median(tm[1:])
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
            v = v * (1.0 / DIM)  # scale_loss (conf/solver/oc_base.yaml:22)
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
        traj_steps = B * T_STEPS
        value = world * traj_steps * args.steps / (total_ms * 1e-3)
        e2e_value = world * traj_steps * args.steps / (e2e_ms * 1e-3)
        achieved_tf = traj_steps * flops_per_traj_step(DIM) / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if engine_used == "simt" else "f32 (MLP as 3 bf16 tcgen05 MMA passes over hi/lo-split operands with fp32 accumulate, fp32-equivalent; f32 elsewhere)",
            "data": "synthetic",
            "config": {"workload": workload_name(B), "engine": engine_used, "global_batch": world * B,
                       "l2": "flushed between timed steps (192 MiB memset inside the region)",
                       "noise": "in-kernel Philox4x32-10", "loss_value": loss_value},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * DIM * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps, "loss_value": host_loss},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": recorded_traffic(engine_used, B), "kernel_ms": kernel_ms,
                         "flops_per_traj_step": flops_per_traj_step(DIM), "peak_source": peak_src,
                         "traj_steps_per_s_kernel": traj_steps / (kernel_ms * 1e-3),
                         "hbm_algorithmic_bytes": B * (8 * DIM + 4),
                         "hbm_gbs": B * (8 * DIM + 4) / (kernel_ms * 1e-3) / 1e9},
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)},
            "train_step": {"what": "loss(ts, x0, ...) with grad + loss.backward(): rollout keeping xs, then the lv gradient "
                                   "(forward + dgrad + wgrad GEMMs over all B*T rows)", "ms": train_ms,
                           "traj_steps_per_s": (world * traj_steps / (train_ms * 1e-3)) if isinstance(train_ms, float) else None,
                           "kl_ms": kl_ms, "kl_what": "same workload with loss.method=kl: rollout keeping xs, backpropagation through time as a "
                                                      "discrete adjoint (per-step cotangent kernel + fused tcgen05 dgrad chain) "
                                                      "inside the gradient's GEMM passes",
                           "full_iteration_ms": full_ms,
                           "full_iteration_what": "zero_grad + lv loss + backward + sdes_trainer_step (grad check, clip_grad_norm_, "
                                                  "Adam, EMA), no host sync inside"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            probe = min(time_cpu_port(256, repeats=2))
            sample = int(256 * T_STEPS / probe * 15.0 / T_STEPS)
            sample = max(64, min(B, sample // 64 * 64))
            dt = min(time_cpu_port(sample, repeats=1))
            line["cpu_baseline"] = {"value": sample * T_STEPS / dt, "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": f"oracle/torch_port.py (reference op sequence, torch eager fp32) on {sample} of {B} trajectories x T={T_STEPS}, one pass, {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
