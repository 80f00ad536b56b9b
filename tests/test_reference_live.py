"""Live check of the plug-in introspection against the UNMODIFIED reference (only where /root/reference exists — the build
container; skipped on the GPU box): `extract_spec` applied to freshly built reference objects must reproduce the spec
frozen in the golden fixtures, i.e. the path a real solver takes (reference classes, bound methods) is what the GPU tests
replay through the mirrors."""
import numpy as np
import pytest

from oracle import ref_harness
from oracle.cases import CASES

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="the reference package is not present on this machine")


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and set(a) == set(b), (path, sorted(a), sorted(b) if isinstance(b, dict) else b)
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}/{i}")
    elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b), err_msg=path)
    elif isinstance(a, float) or isinstance(b, float):
        assert (a is None) == (b is None) and (a is None or a == pytest.approx(b, rel=0, abs=0) or (a != a and b != b)), (path, a, b)
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("name", ["dis_gmm50_lv", "dis_gmm2_lv", "pis_funnel10_kl", "eulerdds_gmm2_lv", "dds_nice16_lv",
                                  "dis_lerptarget_gmmrand3_dimgate", "dis_lerpprior_multiwell4", "dis_noscore_constou_gauss5"])
def test_extract_spec_on_live_reference_objects(golden, name):
    from oracle.gen_golden import reference_spec

    case = CASES[name]
    built = ref_harness.build_case(case)
    live = reference_spec(built, case, compute_ito=case["method"] != "kl")
    from oracle import specio
    import io, tempfile, os

    # through the same (de)serialisation as the fixture so that scalar / array types compare like for like
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "live.npz")
        specio.save(path, {"spec": live})
        live = specio.load(path)["spec"]
    _same(live, golden(name)["spec"])
