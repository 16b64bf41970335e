"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference.

The reference (dominikbuenger/torch_nfft) ships no golden vectors; its NFFT ops need a CUDA
device.  This script is run on a B200 box (`gpurun -- python tests/golden/make_golden.py`) with
the reference built into baseline/_ref; it writes gpurun_out/golden/*.npz, which
are then committed under tests/golden/.  Inputs are seeded; every file stores inputs, parameters
and the reference outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torch_nfft as ref  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda")


def points(rng, n, d, B, scale=1.0):
    pos = ((rng.random((n * B, d), dtype=np.float32) - 0.5) * scale).astype(np.float32)
    batch = np.repeat(np.arange(B, dtype=np.int64), n)
    return pos, batch


def values(rng, shape, cplx):
    v = rng.standard_normal(shape).astype(np.float32)
    if cplx:
        v = (v + 1j * rng.standard_normal(shape)).astype(np.complex64)
    return v


def t(a):
    return None if a is None else torch.from_numpy(a).to(dev)


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items()})


rng = np.random.default_rng(20240229)
NEW_ONLY = "--new-only" in sys.argv  # round 2: write only the fixtures added below (the older ones are committed)

# ---- round-2 additions: the instantiations the BASELINE configs run ---------------------------------
# (own generator, so the round-1 fixtures below stay reproducible bit for bit)
rng2 = np.random.default_rng(20261018)
for name, d, N, m, B, n, C, cplx, ro in [
    ("adjoint_3d_m4_n32", 3, 32, 4, 2, 2500, 1, False, False),      # spread_reg_kernel<10,4,4,2>: the c4 kernel
    ("adjoint_2d_m4_c8_n64", 2, 64, 4, 2, 3000, 8, False, False),   # spread_reg2d_kernel<10,8>: the c3 kernel
    ("adjoint_1d_m8_n256", 1, 256, 8, 3, 2000, 1, False, False),    # spread1d_kernel, m = 8: the c2 kernel
]:
    pos, batch = points(rng2, n, d, B)
    x = values(rng2, (n * B, C), cplx)
    y = ref.nfft_adjoint(t(x), t(pos), t(batch), N, m, ro)
    y2 = ref.nfft_adjoint(t(x), t(pos), t(batch), N, m, ro)
    save(name, op="adjoint", pos=pos, batch=batch, x=x, N=N, m=m, real_output=ro, y=y.cpu().numpy(),
         run_to_run=float((y - y2).abs().max().item()))
for name, d, N, m, B, n, C, cplx, ro in [
    ("forward_3d_m4_n32_realout", 3, 32, 4, 2, 2500, 1, True, True),     # gather_reg_kernel<10,4,4,2>
    ("forward_2d_m4_c8_n64_realout", 2, 64, 4, 2, 3000, 8, True, True),  # gather_reg2d_kernel<10,8>
    ("forward_1d_m8_n256_realout", 1, 256, 8, 3, 2000, 1, True, True),   # gather1d_kernel, m = 8
]:
    pos, batch = points(rng2, n, d, B)
    xh = values(rng2, (B,) + (N,) * d + (C,), cplx)
    y = ref.nfft_forward(t(xh), t(pos), t(batch), m, ro)
    save(name, op="forward", pos=pos, batch=batch, x=xh, m=m, real_output=ro, y=y.cpu().numpy())
src, sb = points(rng2, 3000, 3, 1, scale=0.5)
x = values(rng2, (3000, 1), False)
co = ref.gaussian_interpolated_coeffs(0.1, 3, 32)
y = ref.nfft_fastsum(t(x), co, t(src), batch=t(sb), cutoff=4)
save("fastsum_3d_m4_n32_sym_interp", op="fastsum", sources=src, source_batch=sb, x=x, coeffs=co.cpu().numpy(), m=4,
     y=y.cpu().numpy())
if NEW_ONLY:
    print("golden (round-2 additions) done")
    sys.exit(0)

# ---- adjoint -------------------------------------------------------------------------------
for name, d, N, m, B, n, C, cplx, ro in [
    ("adjoint_1d_real", 1, 32, 8, 2, 200, 1, False, False),
    ("adjoint_2d_real", 2, 16, 4, 3, 300, 2, False, False),
    ("adjoint_2d_cplx_realout", 2, 16, 3, 2, 250, 2, True, True),
    ("adjoint_3d_real", 3, 16, 3, 2, 200, 1, False, False),
    ("adjoint_3d_cplx", 3, 8, 2, 1, 150, 2, True, False),
]:
    pos, batch = points(rng, n, d, B)
    x = values(rng, (n * B, C), cplx)
    y = ref.nfft_adjoint(t(x), t(pos), t(batch), N, m, ro)
    y2 = ref.nfft_adjoint(t(x), t(pos), t(batch), N, m, ro)
    save(name, op="adjoint", pos=pos, batch=batch, x=x, N=N, m=m, real_output=ro, y=y.cpu().numpy(),
         run_to_run=float((y - y2).abs().max().item()))

# no batch vector, 1-D x (no channel dim)
pos, _ = points(rng, 300, 2, 1)
x = values(rng, (300,), False)
y = ref.nfft_adjoint(t(x), t(pos), None, 16, 3, False)
save("adjoint_2d_nobatch_1dx", op="adjoint", pos=pos, x=x, N=16, m=3, real_output=False, y=y.cpu().numpy())

# ---- forward -------------------------------------------------------------------------------
for name, d, N, m, B, n, C, cplx, ro in [
    ("forward_1d_cplx", 1, 32, 8, 2, 200, 1, True, False),
    ("forward_2d_real", 2, 16, 4, 1, 100, 1, False, False),  # real xhat as in reference test_forward.py:32
    ("forward_2d_cplx_realout", 2, 16, 4, 3, 300, 2, True, True),
    ("forward_2d_real_realout", 2, 16, 3, 2, 200, 3, False, True),
    ("forward_3d_cplx", 3, 16, 3, 2, 200, 1, True, False),
    ("forward_3d_cplx_realout", 3, 8, 2, 2, 200, 2, True, True),
]:
    pos, batch = points(rng, n, d, B)
    xh = values(rng, (B,) + (N,) * d + (C,), cplx)
    y = ref.nfft_forward(t(xh), t(pos), t(batch), m, ro)
    save(name, op="forward", pos=pos, batch=batch, x=xh, m=m, real_output=ro, y=y.cpu().numpy())

# ---- fastsum -------------------------------------------------------------------------------
# NOTE: only the symmetric case (targets is sources) can be generated.  The reference's
# non-symmetric path calls cudaFree(&point_shifts) (core_cuda.cu:782-783), the following
# CHECK_ERRORS() sees "invalid argument" and exit()s the process (observed on B200:
# "GPUassert: invalid argument .../core_cuda.cu 794").  sources != targets is therefore pinned by
# the oracle and by ndft_fastsum only.
for name, d, N, m, B, ns, nt, C, cplx, ckind in [
    ("fastsum_2d_sym_analytic", 2, 16, 3, 2, 200, 0, 2, False, "analytic"),
    ("fastsum_2d_sym_interp", 2, 16, 4, 2, 200, 0, 2, False, "interp"),
    ("fastsum_3d_sym_interp", 3, 16, 3, 1, 300, 0, 1, False, "interp"),
    ("fastsum_3d_sym_cplx", 3, 8, 2, 2, 150, 0, 1, True, "analytic"),
    ("fastsum_1d_sym_cplx_interp", 1, 32, 4, 1, 200, 0, 2, True, "interp0"),
]:
    src, sb = points(rng, ns, d, B, scale=0.5)
    x = values(rng, (ns * B, C), cplx)
    if ckind == "analytic":
        co = ref.gaussian_analytic_coeffs(0.15, d, N)
    elif ckind == "interp":
        co = ref.gaussian_interpolated_coeffs(0.15, d, N)
    else:
        co = ref.gaussian_interpolated_coeffs(0.15, d, N, 0)
    if nt:
        tgt, tb = points(rng, nt, d, B, scale=0.5)
        y = ref.nfft_fastsum(t(x), co, t(src), t(tgt), t(sb), t(tb), cutoff=m)
        save(name, op="fastsum", sources=src, source_batch=sb, targets=tgt, target_batch=tb, x=x,
             coeffs=co.cpu().numpy(), m=m, y=y.cpu().numpy())
    else:
        y = ref.nfft_fastsum(t(x), co, t(src), batch=t(sb), cutoff=m)
        save(name, op="fastsum", sources=src, source_batch=sb, x=x, coeffs=co.cpu().numpy(), m=m,
             y=y.cpu().numpy())

# ---- coefficient helpers ---------------------------------------------------------------------
save("coeffs",
     analytic_2d=ref.gaussian_analytic_coeffs(0.2, 2, 8).cpu().numpy(),
     analytic_3d=ref.gaussian_analytic_coeffs(0.1, 3, 8).cpu().numpy(),
     interp_2d=ref.gaussian_interpolated_coeffs(0.2, 2, 8).cpu().numpy(),
     interp_2d_p0=ref.gaussian_interpolated_coeffs(0.2, 2, 8, 0).cpu().numpy(),
     interp_3d=ref.gaussian_interpolated_coeffs(0.1, 3, 8).cpu().numpy(),
     grid_2d=ref.interpolation_grid(2, 8).cpu().numpy(),
     radial_3d=ref.radial_interpolation_grid(3, 4).cpu().numpy(),
     kernel_coeffs_2d=ref.interpolated_kernel_coeffs(
         torch.exp(-ref.radial_interpolation_grid(2, 8) ** 2 / 0.04)).cpu().numpy())
print("golden done")
