"""Host-side logic of the plug-ins that needs no GPU: the cached spec introspection (losses.FusedOCLoss._spec), the
parameter ordering shared by the blob and the autograd node, the derived slots of merged statistics, layout helpers."""
import numpy as np
import pytest
import torch

from oracle import specio
from sde_sampler_b200 import engine
from sde_sampler_b200.dist import merge_stats
from sde_sampler_b200.spec import ctrl_parameters, extract_spec


@pytest.fixture(scope="module")
def objs():
    import os

    from conftest import GOLDEN_DIR
    from sdes_test_helpers import build_from_spec

    # the headline workload's objects (mirrors on the CPU hold parameters only; no kernel is touched)
    return build_from_spec(specio.load(os.path.join(GOLDEN_DIR, "dis_gmm50_lv.npz"))["spec"], torch.device("cpu"), engine="simt")


def _spec(o, **kw):
    return o["loss"]._spec(o["ts"], o["terminal"], o["second"], train=True, compute_ito=True, return_traj=False, **kw)


def test_spec_cache_hits_and_refreshes_scalars(objs):
    o = objs
    a = _spec(o)
    b = _spec(o)
    assert a is b                                        # same objects, same storages: introspection is reused
    o["ctrl"].clip_model, o["ctrl"].clip_score = 50.0, 20.0   # what MultiStepParams does between steps
    c = _spec(o)
    assert c is a and c.ctrl["clip_model"] == 50.0 and c.ctrl["clip_score"] == 20.0
    ts2 = o["ts"].clone() * 0.5                          # a new grid tensor every step (solver.train_ts())
    d = o["loss"]._spec(ts2, o["terminal"], o["second"], train=True, compute_ito=True, return_traj=False)
    assert d is a and torch.equal(d.ts, ts2)
    o["ctrl"].clip_model, o["ctrl"].clip_score = 10.0, 10.0


def test_spec_cache_misses_when_storage_is_replaced(objs):
    o = objs
    a = _spec(o)
    w = o["ctrl"].base_model.out_layer.weight
    w.data = w.data.clone()                              # re-allocated parameter storage (e.g. Module.to, load)
    b = _spec(o)
    assert b is not a
    assert b.mlp["out_w"].data_ptr() == w.data_ptr()


def test_cached_spec_sees_in_place_updates(objs):
    o = objs
    a = _spec(o)
    blob0 = engine.pack_params(a).clone()
    with torch.no_grad():
        o["ctrl"].base_model.input_embed.bias.add_(1.0)  # optimizer step / EMA copy_: in place
    blob1 = engine.pack_params(_spec(o))
    n_w = o["ctrl"].base_model.input_embed.weight.numel()
    assert torch.allclose(blob1[n_w:n_w + 64], blob0[n_w:n_w + 64] + 1.0)
    with torch.no_grad():
        o["ctrl"].base_model.input_embed.bias.sub_(1.0)


def test_blob_order_matches_parameter_list(objs):
    o = objs
    spec = extract_spec(o["loss"], "time_reversal", o["ts"], o["terminal"], o["second"], train=True, compute_ito=True)
    flat = torch.cat([p.detach().reshape(-1) for p in ctrl_parameters(o["ctrl"])])
    assert torch.equal(flat, engine.pack_params(spec)) and flat.numel() == engine.pack_params_numel(spec)


def test_merged_stats_carry_the_losses():
    rng = np.random.default_rng(1)
    r = rng.standard_normal(1000) * 2 + 5

    def st(x):
        return torch.tensor([x.size, x.sum(), (x * x).sum(), (-x).max(), np.exp(-x - (-x).max()).sum(), x.size, 0, 0], dtype=torch.float64)

    m = merge_stats(torch.stack([st(r[:300]), st(r[300:])]))
    assert m[6].item() == pytest.approx(r.var(ddof=1), rel=1e-10) and m[7].item() == pytest.approx(r.mean(), rel=1e-12)


def test_tiled_trajectory_size():
    assert engine.tiled_traj_numel(100, 65536, 50) == 101 * 512 * 56 * 128
    assert engine.tiled_traj_numel(3, 1, 2) == 4 * 1 * 8 * 128
