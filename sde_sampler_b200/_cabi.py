"""ctypes binding of `include/sdes_b200.h` — the stub a maintainer of the reference would add
(see INTEGRATION.md).  There is no fallback: if the shared library is missing or stale the
import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libsdes_b200.so")

ABI_VERSION = 3
CHANNELS = 64
MAX_DIM = 64
MAX_WIDE_DIM = 4096
MAX_HIDDEN = 6
MAX_COMPONENTS = 64

LOSS_TIME_REVERSAL, LOSS_REFERENCE_SDE, LOSS_EXP_INTEGRATOR = 0, 1, 2
CTRL = {"clipped": 0, "score": 1, "lerp": 2, "lerp_prior": 3, "lerp_target": 4}
LOSS = {"time_reversal": 0, "reference_sde": 1, "exp_integrator": 2}
SDE_NONE, SDE_VP, SDE_CONST_OU = 0, 1, 2
TARGET_GMM, TARGET_MULTIWELL, TARGET_FUNNEL, TARGET_NICE = 0, 1, 2, 3

F_RND0_ZERO = 1 << 0
F_COMPUTE_ITO = 1 << 1
F_SUB_DIV_INT = 1 << 2
F_RETURN_TRAJ = 1 << 3
F_NOISE_FROM_HBM = 1 << 4
F_REFERENCE_CTRL = 1 << 5
F_HAS_GATE = 1 << 6
F_MLP_SIMT = 1 << 7
F_TRAJ_TILED = 1 << 8
F_KEEP_FOR_GRAD = 1 << 9
F_KEEP_SCORE = 1 << 10

GRAD_TARGET_SCORE_CONST = 1 << 0
GRAD_SCORE_DETACHED = 1 << 1
GRAD_LAYERWISE_SWEEP = 1 << 2

MASK_ISFINITE, MASK_MAX_RND, MASK_ALL = 0, 1, 2

_fp = C.c_void_p  # device pointers travel as integers


class RolloutDesc(C.Structure):
    """struct SdesRolloutDesc (include/sdes_b200.h) — field order and types must match."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("abi_version", C.c_uint32),
        ("loss_kind", C.c_int32), ("ctrl_kind", C.c_int32), ("sde_kind", C.c_int32), ("target_kind", C.c_int32),
        ("flags", C.c_uint32),
        ("dim", C.c_int32), ("n_steps", C.c_int32), ("n_hidden", C.c_int32), ("te_hidden", C.c_int32),
        ("gate_hidden", C.c_int32), ("gate_dim", C.c_int32), ("n_components", C.c_int32),
        ("n_double_wells", C.c_int32),
        ("batch", C.c_int64), ("traj_offset", C.c_uint64), ("seed", C.c_uint64),
        ("clip_model", C.c_float), ("clip_score", C.c_float), ("clip_target", C.c_float), ("scale_score", C.c_float),
        ("alpha", C.c_float), ("sigma", C.c_float),
        ("beta_min", C.c_float), ("beta_max", C.c_float), ("scale_diff", C.c_float), ("terminal_t", C.c_float),
        ("sde_sign", C.c_float), ("drift_coeff", C.c_float), ("diff_coeff", C.c_float),
        ("separation", C.c_float), ("shift", C.c_float), ("variance", C.c_float), ("log_norm_const", C.c_float),
        ("ts", _fp), ("params", _fp), ("n_params", C.c_int64),
        ("gmm_loc", _fp), ("gmm_scale", _fp), ("gmm_weights", _fp),
        ("prior_loc", _fp), ("prior_scale", _fp), ("ref_loc", _fp), ("ref_scale", _fp),
        ("x0", _fp), ("noise", _fp), ("x_T", _fp), ("rnd", _fp), ("xs", _fp),
        ("workspace", _fp), ("workspace_bytes", C.c_size_t),
        ("nice_couplings", C.c_int32), ("nice_mid", C.c_int32), ("nice_hidden", C.c_int32),
        ("nice_mask_config", C.c_int32), ("nice_params", _fp), ("n_nice_params", C.c_int64),
        ("gate_cot", _fp), ("score_keep", _fp),
    ]


class LvGradDesc(C.Structure):
    """struct SdesLvGradDesc (include/sdes_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("flags", C.c_uint32),
        ("xs", _fp), ("w", _fp), ("grad_params", _fp), ("grad_emb", _fp), ("grad_gate", _fp),
        ("chunk_rows", C.c_int64), ("gate_cot", _fp), ("score_keep", _fp),
    ]


class TrainerStepDesc(C.Structure):
    """struct SdesTrainerStepDesc (include/sdes_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("reserved", C.c_uint32), ("n", C.c_int64),
        ("params", _fp), ("grads", _fp), ("exp_avg", _fp), ("exp_avg_sq", _fp), ("ema_shadow", _fp), ("loss", _fp),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("max_loss", C.c_float), ("max_grad", C.c_float), ("grad_clip_norm", C.c_float),
        ("ema_decay", C.c_double), ("ema_inv_gamma", C.c_double), ("ema_power", C.c_double), ("ema_min_value", C.c_double),
        ("ema_update_after_step", C.c_int32), ("ema_update_every", C.c_int32),
        ("state", _fp), ("workspace", _fp), ("workspace_bytes", C.c_size_t),
    ]


class IntegrateDesc(C.Structure):
    """struct SdesIntegrateDesc (include/sdes_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("n_steps", C.c_int32), ("n_out", C.c_int32),
        ("diff_coeff", C.c_float), ("clip_score", C.c_float), ("eps", C.c_float),
        ("timesteps", _fp), ("out_ts", _fp), ("x_init", _fp), ("xs_out", _fp),
        ("noise_is_increment", C.c_int32), ("reserved", C.c_int32),
    ]


class AffineIntegrateDesc(C.Structure):
    """struct SdesAffineIntegrateDesc (include/sdes_b200.h)."""
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("dim", C.c_int32), ("batch", C.c_int64), ("n_steps", C.c_int32), ("n_out", C.c_int32),
        ("eps", C.c_float), ("cmax", C.c_float),
        ("timesteps", _fp), ("out_ts", _fp), ("tab", _fp), ("cloc", _fp), ("x_init", _fp), ("noise", _fp),
        ("noise_is_increment", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64), ("traj_offset", C.c_uint64),
        ("xs_out", _fp),
    ]


# every symbol include/sdes_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sdes_integrate_workspace_bytes": (C.c_size_t, [C.POINTER(RolloutDesc)]),
    "sdes_langevin_integrate": (C.c_int, [C.POINTER(RolloutDesc), C.POINTER(IntegrateDesc), C.c_void_p]),
    "sdes_affine_integrate": (C.c_int, [C.POINTER(AffineIntegrateDesc), C.c_void_p]),
    "sdes_expectations": (C.c_int, [_fp, C.c_int64, C.c_int32, _fp, C.c_void_p]),
    "sdes_lv_grad_workspace_bytes": (C.c_size_t, [C.POINTER(RolloutDesc), C.POINTER(LvGradDesc)]),
    "sdes_rollout_lv_grad": (C.c_int, [C.POINTER(RolloutDesc), C.POINTER(LvGradDesc), C.c_void_p]),
    "sdes_kl_grad_workspace_bytes": (C.c_size_t, [C.POINTER(RolloutDesc), C.POINTER(LvGradDesc)]),
    "sdes_rollout_kl_grad": (C.c_int, [C.POINTER(RolloutDesc), C.POINTER(LvGradDesc), C.c_void_p]),
    "sdes_lv_traj_weights": (C.c_int, [_fp, C.c_int64, C.c_int32, C.c_int, C.c_float, _fp, _fp, _fp, _fp, C.c_void_p]),
    "sdes_kl_weights": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_float, _fp, _fp, _fp, _fp, C.c_void_p]),
    "sdes_sample_gauss_prior": (C.c_int, [_fp, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_float,
                                          C.c_uint64, C.c_uint64, _fp, C.c_void_p]),
    "sdes_trainer_workspace_bytes": (C.c_size_t, []),
    "sdes_trainer_step": (C.c_int, [C.POINTER(TrainerStepDesc), C.c_void_p]),
    "sdes_eval_moments": (C.c_int, [_fp, _fp, C.c_int64, C.c_int32, _fp, C.c_void_p]),
    "sdes_version": (C.c_int, []),
    "sdes_last_error": (C.c_char_p, []),
    "sdes_workspace_bytes": (C.c_size_t, [C.POINTER(RolloutDesc)]),
    "sdes_rollout_fwd": (C.c_int, [C.POINTER(RolloutDesc), C.c_void_p]),
    "sdes_tcgen05_supported": (C.c_int, [C.POINTER(RolloutDesc)]),
    "sdes_rnd_stats": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_float, _fp, _fp, C.c_void_p]),
    "sdes_merge_stats": (C.c_int, [_fp, C.c_int32, _fp, C.c_void_p]),
    "sdes_weights": (C.c_int, [_fp, C.c_int64, _fp, _fp, C.c_void_p]),
    "sdes_lv_traj_stats": (C.c_int, [_fp, C.c_int64, C.c_int32, C.c_int, C.c_float, _fp, _fp, C.c_void_p]),
    "sdes_lv_weights": (C.c_int, [_fp, C.c_int64, C.c_int, C.c_float, _fp, _fp, _fp, _fp, C.c_void_p]),
    "sdes_philox_normal": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_int32, _fp, C.c_void_p]),
    "sdes_tcgen05_selftest": (C.c_int, [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "sdes_gelu_probe": (C.c_int, [_fp, _fp, C.c_int64, C.c_void_p]),
    "sdes_gelu_pair_probe": (C.c_int, [_fp, _fp, C.c_int64, C.c_void_p]),
    "sdes_launch_count": (C.c_int64, []),
}

_lib = None


class SdesError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load csrc/libsdes_b200.so (built in-tree by `python -m sde_sampler_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdesError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(run `python -m sde_sampler_b200.build`). There is no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if handle.sdes_version() != ABI_VERSION:
            raise SdesError(f"{LIB_PATH} has ABI {handle.sdes_version()}, binding expects {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise SdesError(f"{what} failed ({rc}): {lib().sdes_last_error().decode()}")


def new_desc() -> RolloutDesc:
    d = RolloutDesc()
    d.struct_bytes = C.sizeof(RolloutDesc)
    d.abi_version = ABI_VERSION
    return d
