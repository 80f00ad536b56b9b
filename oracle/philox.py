"""TEST INFRASTRUCTURE — numpy restatement of the in-kernel noise stream.

The reference draws `torch.randn_like(x)` once per step (losses/oc.py:214, :326, :432);
bitwise RNG parity with torch is not a goal (SURVEY §7 "RNG parity").  The B200 path
defines its own counter-based stream so that results are invariant to how the
trajectory batch is sharded over GPUs:

    Philox4x32-10( counter = (traj_global_idx, step, dim_chunk, stream_hi),
                   key     = (seed_lo, seed_hi) )  ->  4 x uint32
    u   = float32(r) * 2^-32 + 2^-33                      (in (0, 1])
    rad = sqrt(-2 ln u_even) ; theta = 2*pi*u_odd - pi
    eps[4c+0], eps[4c+1] = rad0*cos(theta0), rad0*sin(theta0)   (from r0, r1)
    eps[4c+2], eps[4c+3] = rad1*cos(theta1), rad1*sin(theta1)   (from r2, r3)

This file is the checker for `csrc/philox.cuh`; integer part is bit-exact, the
Box-Muller part agrees to the accuracy of the GPU fast-math intrinsics (~1e-6).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)
STREAM_HI = np.uint32(0x5DE5A301)  # 4th counter word: tags the "rollout noise" stream
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds (Salmon et al. 2011). All args uint32 arrays
    (broadcastable). Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _uniform(r):
    # fp32: r * 2^-32 + 2^-33  (exactly what the kernel does, one FMA; the product is exact)
    return (r.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)


def _box_muller(ua, ub):
    ua64 = ua.astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(ua64))
    theta = (np.float32(6.283185307179586) * ub + np.float32(-3.141592653589793)).astype(np.float32)
    th64 = theta.astype(np.float64)
    return (rad * np.cos(th64)).astype(np.float32), (rad * np.sin(th64)).astype(np.float32)


def normal_block(seed: int, traj_idx, step: int, dim: int):
    """eps[b, j] for global trajectory indices `traj_idx` (array) at time step `step`."""
    traj_idx = np.asarray(traj_idx, dtype=np.uint32)
    nchunk = (dim + 3) // 4
    chunks = np.arange(nchunk, dtype=np.uint32)
    r = philox4x32_10(traj_idx[:, None], np.uint32(step), chunks[None, :], STREAM_HI,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [_uniform(x) for x in r]
    n0, n1 = _box_muller(u[0], u[1])
    n2, n3 = _box_muller(u[2], u[3])
    out = np.stack([n0, n1, n2, n3], axis=-1).reshape(traj_idx.shape[0], nchunk * 4)
    return np.ascontiguousarray(out[:, :dim])


def normal_noise(seed: int, n_traj: int, n_steps: int, dim: int, traj_offset: int = 0):
    """(T, B, d) float32 noise tensor of the stream — what the kernel draws in-register."""
    idx = np.arange(traj_offset, traj_offset + n_traj, dtype=np.uint64).astype(np.uint32)
    return np.stack([normal_block(seed, idx, i, dim) for i in range(n_steps)], axis=0)
