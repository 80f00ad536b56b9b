"""-m gpu: the tensor-core building block on its own — one bf16 hi/lo split tcgen05 layer (A staged in TMEM,
W images in shared memory, three kind::f16 passes) against a float64 matmul.  Runs before the rollout parity tests so a
descriptor / layout bug shows up here, in isolation."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("K,N", [(8, 16), (8, 64), (16, 64), (56, 64), (64, 64), (64, 16), (64, 48), (32, 32)])
def test_split_layer_matches_fp64(K, N, mode):
    from sde_sampler_b200 import _cabi

    lib = _cabi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(K * 100 + N)
    A = (torch.randn(128, K, generator=g) * 3).to(dev)
    W = torch.randn(N, K, generator=g).to(dev)
    D = torch.full((128, N), float("nan"), device=dev)
    rc = lib.sdes_tcgen05_selftest(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, mode,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.sdes_last_error()
    torch.cuda.synchronize()
    want = A.double() @ W.double().T
    scale = (A.double().abs() @ W.double().abs().T)
    err = ((D.double() - want).abs() / scale).max().item()
    # fp32 SGEMM itself sits at ~1e-7 of sum|a||w|; a single bf16 pass would be ~4e-3, single-pass TF32 ~5e-4.
    # 16 significant bits per operand (the dropped lo*lo term and the lo roundings are ~2^-17 each)
    tol = 2e-5
    print(f"mode {mode} K={K} N={N}: max relative error {err:.3e}")
    assert err < tol, f"relative error {err:.3e} (K={K}, N={N}, mode={mode})"


def test_fast_gelu_matches_exact_erf_gelu():
    """The epilogue's GELU (erfc via A&S 7.1.26 on MUFU) against float64 exact-erf GELU, and against
    torch's own fp32 GELU on the same points: it must be at least as close to the exact function."""
    from scipy.special import erf

    from sde_sampler_b200 import _cabi

    lib = _cabi.lib()
    dev = torch.device("cuda:0")
    x = torch.cat([torch.linspace(-12, 12, 2_000_001), torch.tensor([0.0, -0.0, 1e-30, -1e-30, 40.0, -40.0])]).to(dev)
    y = torch.empty_like(x)
    assert lib.sdes_gelu_probe(x.data_ptr(), y.data_ptr(), x.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    torch.cuda.synchronize()
    x64 = x.double().cpu().numpy()
    exact = 0.5 * x64 * (1 + erf(x64 / np.sqrt(2)))
    ours = np.abs(y.double().cpu().numpy() - exact)
    theirs = np.abs(torch.nn.functional.gelu(x).double().cpu().numpy() - exact)
    print(f"max |gelu_fast - exact| = {ours.max():.3e}; torch fp32 gelu: {theirs.max():.3e}")
    assert ours.max() < 6e-7
    assert np.isfinite(y.cpu().numpy()).all()


def test_packed_gelu_matches_exact_erf_gelu():
    """The persistent rollout kernel's epilogue GELU (logistic form, packed f32x2) against float64 exact-erf GELU:
    absolute error below 7e-7 + 1.2e-7 |x| (one ulp-class relative term from MUFU.EX2 / MUFU.RCP on large |x|), tails
    saturate to x and 0, no NaN / inf on finite inputs."""
    from scipy.special import erf

    from sde_sampler_b200 import _cabi

    lib = _cabi.lib()
    dev = torch.device("cuda:0")
    x = torch.cat([torch.linspace(-12, 12, 2_000_001), torch.tensor([0.0, -0.0, 1e-30, -1e-30, 40.0, -40.0, 1e4, -1e4, 3e19])]).to(dev)
    y = torch.empty_like(x)
    assert lib.sdes_gelu_pair_probe(x.data_ptr(), y.data_ptr(), x.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0, lib.sdes_last_error()
    torch.cuda.synchronize()
    x64 = x.double().cpu().numpy()
    exact = 0.5 * x64 * (1 + erf(x64 / np.sqrt(2)))
    err = np.abs(y.double().cpu().numpy() - exact)
    bound = 7e-7 + 1.2e-7 * np.abs(x64)
    theirs = np.abs(torch.nn.functional.gelu(x).double().cpu().numpy() - exact)
    core = np.abs(x64) <= 12
    print(f"max |gelu_fast2 - exact| on [-12, 12] = {err[core].max():.3e}; worst err/bound = {(err / bound).max():.3f}; torch fp32 gelu: {theirs[core].max():.3e}")
    assert (err <= bound).all()
    assert np.isfinite(y.cpu().numpy()).all()
