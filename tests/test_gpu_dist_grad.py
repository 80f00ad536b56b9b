"""-m gpu, needs 2 GPUs (skipped otherwise): the training step on two ranks over NCCL — each rank rolls out its contiguous
shard with the global Philox counters, the rnd statistics are merged (`dist.combine_stats`: all_gather_into_tensor +
`sdes_merge_stats`), each rank's `loss.backward()` produces the gradient of the GLOBAL loss restricted to its rows and the
autograd function all-reduces it (reference semantics `losses/oc.py:88-90`, `:106-112` on the whole batch).
The result must equal the single-GPU step on the full batch: loss value and every parameter's gradient, for lv and kl.
Tolerance 2e-3 of each tensor's max (the bound of tests/test_gpu_grad.py; measured ~1e-5)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
B_GLOBAL = 2048


def _one_step(name, method, device, x0, pg):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from oracle import specio
    from sdes_test_helpers import build_from_spec
    from sde_sampler_b200.spec import ctrl_parameters

    g = specio.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
    b = build_from_spec(g["spec"], device, engine="tcgen05", seed=77, process_group=pg)
    b["loss"].method = method
    params = ctrl_parameters(b["ctrl"])
    val, _ = b["loss"](b["ts"], x0, b["terminal"], b["second"])
    val.backward()
    return float(val.detach()), [torch.zeros_like(p) if p.grad is None else p.grad.detach().clone().cpu() for p in params]


def _worker(rank, world, port, name, method, q):
    """world = 1: the single-GPU step on the full batch, in its own fresh process like the ranks (the loss objects' noise
    streams are numbered per process)."""
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    x_full = torch.randn(B_GLOBAL, 50, generator=torch.Generator().manual_seed(3))
    if world == 1:
        val, grads = _one_step(name, method, device, x_full.to(device), None)
        q.put((-1, val, [g.numpy() for g in grads]))
        return
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    per = B_GLOBAL // world
    val, grads = _one_step(name, method, device, x_full[rank * per:(rank + 1) * per].to(device), dist.group.WORLD)
    q.put((rank, val, [g.numpy() for g in grads]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("method", ["lv", "kl"])
def test_two_rank_training_step_equals_single_gpu(method):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    name = "dis_gmm50_lv"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, method, q)) for r in range(2)]
    procs.append(ctx.Process(target=_worker, args=(0, 1, port, name, method, q)))
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = [t for t in got if t[0] == -1][0]
    res = sorted((t for t in got if t[0] >= 0), key=lambda t: t[0])
    val1, grads1 = single[1], [torch.from_numpy(a) for a in single[2]]
    for rank, val, grads in res:
        assert abs(val - val1) <= 1e-4 * (1 + abs(val1)), (rank, val, val1)
        for i, (a, c) in enumerate(zip(grads1, grads)):
            scale = a.abs().max().item()
            err = (a - torch.from_numpy(c)).abs().max().item()
            assert err <= 2e-3 * scale + 1e-7, f"rank {rank} parameter {i}: {err:.3e} vs scale {scale:.3e}"
    # both ranks hold the same (all-reduced) gradient
    for a, c in zip(res[0][2], res[1][2]):
        assert (torch.from_numpy(a) - torch.from_numpy(c)).abs().max().item() == 0.0
