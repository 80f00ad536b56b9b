"""In-tree build of csrc/libsdes_b200.so with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object that links only the static CUDA runtime).

    python -m sde_sampler_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libsdes_b200.so")
OBJ_DIR = os.path.join(CSRC, "_obj")
SOURCES = ["sdes_api.cu", "sdes_prepare.cu", "sdes_rollout_simt.cu", "sdes_rollout_tc_api.cu", "sdes_wide.cu", "sdes_grad.cu", "sdes_adjoint.cu", "sdes_integrate.cu", "sdes_trainer.cu"]
# the tensor-core rollout kernel is compiled once per padded state dimension (parallel builds, one object each)
TC_DPADS = [8, 16, 32, 48, 56, 64]
TC_SOURCE = "sdes_rollout_mma.cu"
HEADERS = ["sdes_common.cuh", "sdes_step.cuh", "sdes_tc.cuh", "sdes_timeembed.cuh", "sdes_linear.cuh", "sdes_grad_fused.cuh", os.path.join("..", "..", "include", "sdes_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return nvcc


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + [TC_SOURCE] + HEADERS:
        path = os.path.join(CSRC, name)
        if os.path.exists(path):
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ_DIR, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == digest:
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(job):
        src, tag, extra = job
        obj = os.path.join(OBJ_DIR, src.replace(".cu", tag + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    jobs = [(TC_SOURCE, f"_{dp}", [f"-DSDES_TC_DPAD={dp}"]) for dp in reversed(TC_DPADS)] + [(s, "", []) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, len(jobs))) as ex:
        objs = list(ex.map(compile_one, jobs))
    r = subprocess.run([nvcc, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
