"""Builds libnfft_b200.so in-tree with plain nvcc for sm_100a (no torch headers, seconds).

The reference builds a torch CUDAExtension without arch flags (reference setup.py:7-19); this
engine is a torch-free C-ABI library, so the build is a single nvcc invocation.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnfft_b200.so")
_SRC_DIR = os.path.join(_HERE, "csrc")
_SOURCES = ["nfft_b200.cu"]
_DEPS = ["nfft_b200.cu", "common.cuh", "sort.cuh", "window.cuh", "window1d.cuh", "window_reg.cuh", "window_reg2d.cuh",
         "spectral.cuh",
         os.path.join("..", "..", "include", "nfft_b200.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(_SRC_DIR, d)) > t for d in _DEPS)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compile the CUDA library if it is missing or older than its sources.
    `defines` / `out` build an experimental variant next to the default library."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("torch_nfft_b200: nvcc not found; cannot build libnfft_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", out or LIB_PATH] + ["-D" + d for d in defines] + _SOURCES + ["-lcufft"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=_SRC_DIR, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("torch_nfft_b200: nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out or LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
