"""CPU tests of the oracle itself: against the exact direct sums, against its own fp64 mode and
against the golden fixtures produced by the compiled reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import nfft_oracle as O
from conftest import GOLDEN_DIR, golden_files

# NFFT-vs-NDFT relative L2 error of the reference's algorithm per cutoff,
# with a 2x safety margin: the oracle must be exactly this accurate, not better, not worse.
ERR_BOUND = {2: 2.5e-2, 3: 2.5e-3, 4: 3e-4, 6: 6e-6, 8: 6e-6}


def _case(rng, d, N, B, n, C, cplx=False):
    pos = rng.random((n * B, d), dtype=np.float32) - 0.5
    batch = np.repeat(np.arange(B), n)
    x = rng.standard_normal((n * B, C)).astype(np.float32)
    if cplx:
        x = (x + 1j * rng.standard_normal((n * B, C))).astype(np.complex64)
    return pos, batch, x


@pytest.mark.parametrize("d,N,m", [(1, 64, 2), (1, 64, 4), (1, 64, 8), (2, 32, 3), (2, 32, 4), (2, 16, 6), (3, 16, 3), (3, 16, 4)])
def test_adjoint_and_forward_match_ndft(d, N, m):
    rng = np.random.default_rng(d * 100 + m)
    pos, batch, x = _case(rng, d, N, 2, 150, 2)
    ya = O.nfft_adjoint(x, pos, batch, N, m)
    assert ya.shape == (2,) + (N,) * d + (2,) and ya.dtype == np.complex64
    assert O.rel_l2(ya, O.ndft_adjoint(x, pos, batch, N)) < ERR_BOUND[m]
    xh = (rng.standard_normal(ya.shape) + 1j * rng.standard_normal(ya.shape)).astype(np.complex64)
    yf = O.nfft_forward(xh, pos, batch, m)
    assert yf.shape == x.shape
    assert O.rel_l2(yf, O.ndft_forward(xh, pos, batch)) < ERR_BOUND[m]


def test_error_decreases_with_cutoff():
    rng = np.random.default_rng(5)
    pos, batch, x = _case(rng, 2, 32, 1, 300, 1)
    exact = O.ndft_adjoint(x, pos, batch, 32)
    errs = [O.rel_l2(O.nfft_adjoint(x, pos, batch, 32, m), exact) for m in (2, 3, 4, 6)]
    assert all(a > 3 * b for a, b in zip(errs, errs[1:])), errs


@pytest.mark.parametrize("d,N,m", [(1, 32, 4), (2, 16, 3), (3, 8, 2)])
def test_fp32_and_fp64_modes_agree(d, N, m):
    rng = np.random.default_rng(7)
    pos, batch, x = _case(rng, d, N, 2, 200, 2, cplx=True)
    a32 = O.nfft_adjoint(x, pos, batch, N, m, prec="f32")
    a64 = O.nfft_adjoint(x, pos, batch, N, m, prec="f64")
    assert O.rel_l2(a32, a64) < 2e-6
    f32 = O.nfft_forward(a32, pos, batch, m, prec="f32")
    f64 = O.nfft_forward(a32, pos, batch, m, prec="f64")
    assert O.rel_l2(f32, f64) < 2e-6


def test_real_output_is_real_part():
    rng = np.random.default_rng(8)
    pos, batch, x = _case(rng, 2, 16, 2, 100, 3)
    full = O.nfft_adjoint(x, pos, batch, 16, 3)
    assert np.allclose(O.nfft_adjoint(x, pos, batch, 16, 3, real_output=True), full.real, atol=1e-6)
    ff = O.nfft_forward(full, pos, batch, 3)
    assert np.allclose(O.nfft_forward(full, pos, batch, 3, real_output=True), ff.real, atol=1e-4)


def test_adjointness():
    """<adjoint(x), yhat> == <x, forward(yhat)> : the two transforms are exact transposes."""
    rng = np.random.default_rng(9)
    pos, batch, x = _case(rng, 2, 16, 2, 120, 2, cplx=True)
    yh = (rng.standard_normal((2, 16, 16, 2)) + 1j * rng.standard_normal((2, 16, 16, 2))).astype(np.complex64)
    lhs = np.vdot(yh.astype(np.complex128), O.nfft_adjoint(x, pos, batch, 16, 4, prec="f64"))
    rhs = np.vdot(O.nfft_forward(yh, pos, batch, 4, prec="f64"), x.astype(np.complex128))
    assert abs(lhs - rhs) / abs(lhs) < 1e-10


def test_periodic_wrap_and_no_batch():
    rng = np.random.default_rng(10)
    pos = rng.random((50, 2), dtype=np.float32) - 0.5
    x = rng.standard_normal(50).astype(np.float32)  # 1-D x: no channel dimension
    y = O.nfft_adjoint(x, pos, None, 16, 3)
    assert y.shape == (1, 16, 16)
    shifted = pos.copy()
    shifted[:10] += 1.0  # outside [-1/2, 1/2): wraps periodically
    assert O.rel_l2(O.nfft_adjoint(x, shifted, None, 16, 3), y) < 1e-5


def test_fastsum_matches_dense_gaussian():
    rng = np.random.default_rng(11)
    n, d, N, m, sigma = 150, 2, 32, 4, 0.1
    pos = ((rng.random((n, d), dtype=np.float32) - 0.5) * 0.5).astype(np.float32)
    dense = np.exp(-((pos[None] - pos[:, None]) ** 2).sum(-1) / sigma ** 2)
    eye = np.eye(n, dtype=np.float32)
    for co in (O.gaussian_analytic_coeffs(sigma, d, N), O.gaussian_interpolated_coeffs(sigma, d, N),
               O.gaussian_interpolated_coeffs(sigma, d, N, p=0)):
        assert np.abs(O.nfft_fastsum(eye, co, pos, m=m) - dense).max() < 1e-5
    x = rng.standard_normal((n, 2)).astype(np.float32)
    co = O.gaussian_analytic_coeffs(sigma, d, N)
    assert O.rel_l2(O.nfft_fastsum(x, co, pos, m=m), O.ndft_fastsum(x, co, pos)) < 5e-6


def test_binning_oracle():
    rng = np.random.default_rng(12)
    pos = rng.random((1000, 3), dtype=np.float32) - 0.5
    batch = np.repeat(np.arange(2), 500)
    keys = O.tile_keys(pos, batch, 16, (8, 8, 16))
    assert keys.min() >= 0 and keys.max() < 2 * 4 * 4 * 2
    perm = O.stable_permutation(keys)
    assert np.all(np.diff(keys[perm]) >= 0)
    same = np.diff(keys[perm]) == 0
    assert np.all(np.diff(perm)[same] > 0)  # stability


# ---------------------------------------------------------------------------------------------
# golden fixtures = outputs of the compiled reference on a B200 (pins the oracle to the reference)
# ---------------------------------------------------------------------------------------------
GOLDEN_TOL = 1e-5  # north-star parity tolerance (relative L2, fp32)


def _none(z, k):
    return z[k] if k in z.files else None


@pytest.mark.parametrize("fname", golden_files("adjoint") + golden_files("forward") + golden_files("fastsum"))
def test_oracle_matches_reference_golden(fname):
    z = np.load(os.path.join(GOLDEN_DIR, fname))
    op = str(z["op"])
    if op == "adjoint":
        y = O.nfft_adjoint(z["x"], z["pos"], _none(z, "batch"), int(z["N"]), int(z["m"]), bool(z["real_output"]))
    elif op == "forward":
        y = O.nfft_forward(z["x"], z["pos"], _none(z, "batch"), int(z["m"]), bool(z["real_output"]))
    else:
        y = O.nfft_fastsum(z["x"], z["coeffs"], z["sources"], _none(z, "targets"), _none(z, "source_batch"),
                           _none(z, "target_batch") if "targets" in z.files else _none(z, "source_batch"),
                           m=int(z["m"]))
    assert y.shape == z["y"].shape and y.dtype == z["y"].dtype
    assert O.rel_l2(y, z["y"]) < GOLDEN_TOL


def test_golden_fixtures_present():
    """The fixtures are part of the repository (generated once on a B200 by make_golden.py)."""
    assert len(golden_files()) >= 10, "tests/golden/*.npz missing: run tests/golden/make_golden.py on a GPU box"


@pytest.mark.skipif(not golden_files("coeffs"), reason="golden fixtures not generated yet")
def test_coefficient_helpers_match_reference_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "coeffs.npz"))
    assert np.allclose(O.gaussian_analytic_coeffs(0.2, 2, 8), z["analytic_2d"], rtol=1e-5, atol=1e-9)
    assert np.allclose(O.gaussian_analytic_coeffs(0.1, 3, 8), z["analytic_3d"], rtol=1e-5, atol=1e-9)
    assert np.allclose(O.gaussian_interpolated_coeffs(0.2, 2, 8), z["interp_2d"], atol=1e-6)
    assert np.allclose(O.gaussian_interpolated_coeffs(0.2, 2, 8, p=0), z["interp_2d_p0"], atol=1e-6)
    assert np.allclose(O.gaussian_interpolated_coeffs(0.1, 3, 8), z["interp_3d"], atol=1e-6)
    assert np.allclose(O.interpolation_grid(2, 8), z["grid_2d"], atol=1e-7)
    assert np.allclose(O.radial_interpolation_grid(3, 4), z["radial_3d"], atol=1e-6)
    vals = np.exp(-O.radial_interpolation_grid(2, 8).astype(np.float64) ** 2 / 0.04)
    assert np.allclose(O.interpolated_kernel_coeffs(vals), z["kernel_coeffs_2d"], atol=1e-6)
