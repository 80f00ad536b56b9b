"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the header
declares, agrees with the ctypes mirror of the descriptor, validates descriptors, and the
plug-in introspection reproduces the golden specs."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from sde_sampler_b200 import _cabi, build as sdes_build
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEADER = os.path.join(ROOT, "include", "sdes_b200.h")


@pytest.fixture(scope="module")
def lib():
    sdes_build.build()
    return _cabi.lib()


def test_exports_every_declared_symbol(lib):
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(sdes_[a-z_0-9]+)\s*\(", text))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SYMBOLS), (declared ^ set(_cabi.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None


def test_desc_layout_matches_c(tmp_path, lib):
    """sizeof / offsetof of the C struct, compiled with gcc from the header, equal the ctypes mirror."""
    src = tmp_path / "probe.c"
    fields = [f[0] for f in _cabi.RolloutDesc._fields_]
    body = "\n".join(f'printf("{f} %zu\\n", offsetof(SdesRolloutDesc, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdes_b200.h"\nint main(){'
                   'printf("sizeof %zu\\n", sizeof(SdesRolloutDesc));' + body + "return 0;}")
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    assert int(out["sizeof"]) == C.sizeof(_cabi.RolloutDesc)
    for f in fields:
        assert int(out[f]) == getattr(_cabi.RolloutDesc, f).offset, f
    assert lib.sdes_version() == _cabi.ABI_VERSION


def _valid_desc():
    d = _cabi.new_desc()
    d.loss_kind, d.ctrl_kind, d.sde_kind, d.target_kind = 0, 2, 1, 0
    d.dim, d.n_steps, d.n_hidden, d.te_hidden, d.gate_hidden, d.gate_dim = 50, 100, 2, 1, 3, 1
    d.n_components, d.batch = 40, 1024
    d.flags = _cabi.F_HAS_GATE | _cabi.F_MLP_SIMT
    c, dim = 64, 50
    d.n_params = (c * dim + c) + (c + c * 2 * c + c + c * c + c) + 2 * (c * c + c) + (dim * c + dim) \
        + (c + c * 2 * c + c + 2 * (c * c + c) + c + 1)
    return d


def test_descriptor_validation_without_gpu(lib):
    d = _valid_desc()
    assert lib.sdes_workspace_bytes(C.byref(d)) > 0, lib.sdes_last_error()
    bad = _valid_desc()
    bad.dim = _cabi.MAX_WIDE_DIM + 1
    assert lib.sdes_workspace_bytes(C.byref(bad)) == 0
    assert b"dim" in lib.sdes_last_error()
    # d > 64 selects the wide engine: its workspace holds the state, the operand images and the layer activations
    wide = _valid_desc()
    wide.dim = 100
    wide.n_params += (100 - 50) * 2 * 64 + (100 - 50)
    assert lib.sdes_workspace_bytes(C.byref(wide)) > lib.sdes_workspace_bytes(C.byref(d))
    assert lib.sdes_tcgen05_supported(C.byref(wide)) == 1
    bad = _valid_desc()
    bad.dim, bad.target_kind = 50, _cabi.TARGET_NICE  # NICE needs its parameter blob and an even dim
    assert lib.sdes_workspace_bytes(C.byref(bad)) == 0
    assert b"wide engine" in lib.sdes_last_error()
    bad = _valid_desc()
    bad.n_params += 1
    assert lib.sdes_workspace_bytes(C.byref(bad)) == 0
    assert b"n_params" in lib.sdes_last_error()
    bad = _valid_desc()
    bad.struct_bytes -= 8
    assert lib.sdes_workspace_bytes(C.byref(bad)) == 0
    # pointers are validated before anything is launched: no GPU is touched here
    rc = lib.sdes_rollout_fwd(C.byref(d), None)
    assert rc == -5 and b"NULL" in lib.sdes_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libsdes_b200.so")
    with pytest.raises(_cabi.SdesError, match="no CPU or PyTorch fallback"):
        _cabi.lib()


def test_cpu_tensor_is_rejected(lib, golden):
    """The product has no CPU path: a rollout on a CPU tensor raises instead of computing."""
    from sdes_test_helpers import build_from_spec

    g = golden("dis_dw1_lv")
    b = build_from_spec(g["spec"], "cpu")
    with pytest.raises(_cabi.SdesError, match="CUDA device only"):
        b["loss"](b["ts"], torch.as_tensor(g["x0"]), b["terminal"], b["second"])


@pytest.mark.parametrize("name", ["dis_gmm50_lv", "pis_funnel10_kl", "dds_funnel10_lv", "eulerdds_gmm2_lv",
                                  "dis_lerptarget_gmmrand3_dimgate", "dis_lerpprior_multiwell4",
                                  "dis_noscore_constou_gauss5", "dis_dw1_lv"])
def test_introspection_roundtrip(lib, golden, name):
    """mirror objects built from a golden spec -> extract_spec -> the same spec (the golden spec was
    extracted from the reference's own objects by the same code, oracle/gen_golden.py)."""
    from sde_sampler_b200.spec import extract_spec
    from sdes_test_helpers import build_from_spec

    g = golden(name)
    spec = g["spec"]
    b = build_from_spec(spec, "cpu")
    ls = spec["loss"]
    got = extract_spec(b["loss"], ls["kind"], b["ts"], b["terminal"], b["second"], train=ls["train"],
                       compute_ito=ls["compute_ito"]).to_dict()

    def cmp(a, c, path=""):
        if isinstance(a, dict):
            assert set(a) == set(c), (path, set(a) ^ set(c))
            for k in a:
                cmp(a[k], c[k], f"{path}/{k}")
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(c), path
            for i, (u, v) in enumerate(zip(a, c)):
                cmp(u, v, f"{path}/{i}")
        elif isinstance(a, np.ndarray) or isinstance(c, np.ndarray):
            np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(c, np.float64), rtol=1e-6, atol=1e-7, err_msg=path)
        elif isinstance(a, float) or isinstance(c, float):
            if a is None or c is None:
                assert a == c, path
            else:
                assert a == pytest.approx(c, rel=1e-6), path
        else:
            assert a == c, (path, a, c)

    cmp(spec, got)


def test_aux_desc_layouts_match_c(tmp_path):
    """SdesLvGradDesc / SdesIntegrateDesc: sizeof and offsets of the C structs equal the ctypes mirrors."""
    for cname, cls in (("SdesLvGradDesc", _cabi.LvGradDesc), ("SdesIntegrateDesc", _cabi.IntegrateDesc)):
        fields = [f[0] for f in cls._fields_]
        body = "\n".join(f'printf("{f} %zu\\n", offsetof({cname}, {f}));' for f in fields)
        src = tmp_path / f"probe_{cname}.c"
        src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdes_b200.h"\nint main(){'
                       f'printf("sizeof %zu\\n", sizeof({cname}));' + body + "return 0;}")
        exe = tmp_path / f"probe_{cname}"
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
        out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
        assert int(out["sizeof"]) == C.sizeof(cls), cname
        for f in fields:
            assert int(out[f]) == getattr(cls, f).offset, (cname, f)


def test_gradient_and_integrator_entry_points_validate_without_gpu(lib):
    d = _valid_desc()
    g = _cabi.LvGradDesc()
    g.struct_bytes = C.sizeof(_cabi.LvGradDesc)
    assert lib.sdes_lv_grad_workspace_bytes(C.byref(d), C.byref(g)) > lib.sdes_workspace_bytes(C.byref(d))
    assert lib.sdes_rollout_lv_grad(C.byref(d), C.byref(g), None) == -5 and b"NULL" in lib.sdes_last_error()
    g.struct_bytes -= 4
    assert lib.sdes_lv_grad_workspace_bytes(C.byref(d), C.byref(g)) == 0
    t = _cabi.new_desc()   # the integrator reads only the target part of the descriptor
    t.target_kind, t.dim, t.batch, t.variance = _cabi.TARGET_FUNNEL, 10, 64, 9.0
    assert lib.sdes_integrate_workspace_bytes(C.byref(t)) > 0, lib.sdes_last_error()
    ig = _cabi.IntegrateDesc()
    ig.struct_bytes = C.sizeof(_cabi.IntegrateDesc)
    ig.n_steps, ig.n_out = 10, 3
    assert lib.sdes_langevin_integrate(C.byref(t), C.byref(ig), None) == -5
    t.dim = 100
    assert lib.sdes_integrate_workspace_bytes(C.byref(t)) == 0 and b"Langevin" in lib.sdes_last_error()


def test_aux_struct_layouts_match_c_trainer_and_grad(tmp_path, lib):
    """SdesTrainerStepDesc / SdesLvGradDesc: sizeof and offsetof from gcc equal the ctypes mirrors."""
    src = tmp_path / "probe2.c"
    lines = []
    for cname, cls in (("SdesTrainerStepDesc", _cabi.TrainerStepDesc), ("SdesLvGradDesc", _cabi.LvGradDesc)):
        lines.append(f'printf("{cname}.sizeof %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sdes_b200.h"\nint main(){' + "\n".join(lines) + "return 0;}")
    exe = tmp_path / "probe2"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    for cname, cls in (("SdesTrainerStepDesc", _cabi.TrainerStepDesc), ("SdesLvGradDesc", _cabi.LvGradDesc)):
        assert int(out[f"{cname}.sizeof"]) == C.sizeof(cls)
        for f, _ in cls._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_kl_grad_and_trainer_entry_points_validate(lib):
    d = _valid_desc()
    g = _cabi.LvGradDesc()
    g.struct_bytes = C.sizeof(_cabi.LvGradDesc)
    # a 40-component GMM score inside a Lerp control: the reference treats it as a constant of the graph
    assert lib.sdes_kl_grad_workspace_bytes(C.byref(d), C.byref(g)) == 0
    assert b"SDES_GRAD_TARGET_SCORE_CONST" in lib.sdes_last_error()
    g.flags = _cabi.GRAD_TARGET_SCORE_CONST
    need_kl = lib.sdes_kl_grad_workspace_bytes(C.byref(d), C.byref(g))
    need_lv = lib.sdes_lv_grad_workspace_bytes(C.byref(d), C.byref(g))
    assert need_kl >= need_lv + 4 * d.n_steps * d.batch * d.dim   # the control cotangent of every (trajectory, step)
    d.dim = 100  # wide engine: the sweep needs what a keep-mode forward left in the workspace
    d.n_params += 2 * 64 * 50 + 50
    assert lib.sdes_kl_grad_workspace_bytes(C.byref(d), C.byref(g)) == 0
    assert b"SDES_F_KEEP_FOR_GRAD" in lib.sdes_last_error()
    d.flags |= _cabi.F_KEEP_FOR_GRAD
    assert lib.sdes_kl_grad_workspace_bytes(C.byref(d), C.byref(g)) == 0
    assert b"SDES_F_KEEP_SCORE" in lib.sdes_last_error()
    need_lv_wide = lib.sdes_lv_grad_workspace_bytes(C.byref(d), C.byref(g))
    d.flags |= _cabi.F_KEEP_SCORE  # + the kept scores of every step, the fp32 adjoint and one cotangent image per layer
    assert lib.sdes_kl_grad_workspace_bytes(C.byref(d), C.byref(g)) > need_lv_wide > 0
    t = _cabi.TrainerStepDesc()
    assert lib.sdes_trainer_step(C.byref(t), None) == -2
    t.struct_bytes = C.sizeof(_cabi.TrainerStepDesc)
    t.n = 8
    assert lib.sdes_trainer_step(C.byref(t), None) == -5
    assert lib.sdes_trainer_workspace_bytes() > 0
    assert lib.sdes_sample_gauss_prior(None, 4, 2, 0.0, 1.0, 0, 0.0, 0.0, 0, 0, None, None) == -5
    assert lib.sdes_eval_moments(None, None, 4, 2, None, None) == -5


def test_trainer_has_no_cpu_path(lib):
    from sde_sampler_b200 import FusedAdamEMA, eval_moments, sample_gauss_prior

    with pytest.raises(_cabi.SdesError, match="CUDA device only"):
        FusedAdamEMA([torch.nn.Parameter(torch.zeros(3))])
    with pytest.raises(_cabi.SdesError, match="CUDA device only"):
        sample_gauss_prior(4, 2, device="cpu")
    with pytest.raises(_cabi.SdesError, match="CUDA device only"):
        eval_moments(torch.zeros(4, 2))
