"""GPU parity tests: the sm_100a engine (through the C ABI) against the numpy oracle on seeded
inputs, against the golden fixtures produced by the compiled reference, and -- at the full
BASELINE sizes -- through size-independent properties.

Tolerance: relative L2 <= 1e-5 in fp32 (BASELINE.json north_star).  Integer work (tile keys,
sort permutation) must be bit-exact.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, golden_files
from oracle import nfft_oracle as O
import torch_nfft_b200 as T
from torch_nfft_b200 import _lib

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north-star parity tolerance, relative L2, fp32
DEV = "cuda"


def cuda(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def make_points(rng, d, B, n, ragged=False):
    if ragged:
        sizes = rng.integers(0, 2 * n, size=B)
        sizes[B // 2] = 0  # an empty point set in the middle
        sizes[-1] = max(sizes[-1], 1)  # batch[-1] must be B-1 (reference README.md:44-46)
    else:
        sizes = np.full(B, n)
    batch = np.repeat(np.arange(B, dtype=np.int64), sizes)
    pos = rng.random((batch.size, d), dtype=np.float32) - 0.5
    return pos, batch


def make_values(rng, shape, cplx):
    v = rng.standard_normal(shape).astype(np.float32)
    if cplx:
        v = (v + 1j * rng.standard_normal(shape)).astype(np.complex64)
    return v


CASES = [
    # d, N, m, B, n, C
    (1, 64, 4, 2, 500, 2),
    (1, 256, 8, 3, 2000, 1),
    (2, 32, 4, 3, 700, 3),
    (2, 64, 3, 2, 1500, 8),
    (2, 16, 2, 1, 300, 10),
    (2, 32, 4, 2, 900, 10),  # 2D register kernels: passes of 8 + 2 channels
    (2, 32, 3, 1, 5000, 5),  # passes of 4 + 1, dense (5 points per cell)
    (3, 16, 3, 2, 600, 1),
    (3, 32, 4, 2, 2500, 1),
    (3, 16, 2, 1, 500, 3),
    (3, 12, 3, 1, 400, 2),   # bandwidth not a power of two
    (2, 4, 1, 2, 50, 1),     # smallest legal grid: N = 4, m = 1
]


@pytest.mark.parametrize("d,N,m,B,n,C", CASES)
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("real_output", [False, True])
def test_adjoint_matches_oracle(d, N, m, B, n, C, cplx, real_output):
    rng = np.random.default_rng(hash((d, N, m, C, cplx)) % 2 ** 31)
    pos, batch = make_points(rng, d, B, n)
    x = make_values(rng, (pos.shape[0], C), cplx)
    y = T.nfft_adjoint(cuda(x), cuda(pos), cuda(batch), N, m, real_output=real_output)
    ref = O.nfft_adjoint(x, pos, batch, N, m, real_output=real_output)
    assert y.shape == ref.shape and y.cpu().numpy().dtype == ref.dtype
    assert O.rel_l2(y.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("d,N,m,B,n,C", CASES)
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("real_output", [False, True])
def test_forward_matches_oracle(d, N, m, B, n, C, cplx, real_output):
    rng = np.random.default_rng(hash((d, N, m, C, cplx, 1)) % 2 ** 31)
    pos, batch = make_points(rng, d, B, n)
    xh = make_values(rng, (B,) + (N,) * d + (C,), cplx)
    y = T.nfft_forward(cuda(xh), cuda(pos), cuda(batch), m, real_output=real_output)
    ref = O.nfft_forward(xh, pos, batch, m, real_output=real_output)
    assert y.shape == ref.shape and y.cpu().numpy().dtype == ref.dtype
    assert O.rel_l2(y.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("d,N,m,B,n,C", [(1, 64, 4, 2, 400, 2), (2, 32, 4, 2, 600, 3), (3, 16, 3, 2, 500, 1), (3, 32, 4, 1, 1500, 2)])
@pytest.mark.parametrize("xc,cc", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("symmetric", [True, False])
def test_fastsum_matches_oracle(d, N, m, B, n, C, xc, cc, symmetric):
    rng = np.random.default_rng(hash((d, N, m, C, xc, cc)) % 2 ** 31)
    src, sb = make_points(rng, d, B, n)
    src = (src * 0.5).astype(np.float32)
    x = make_values(rng, (src.shape[0], C), xc)
    co = O.gaussian_interpolated_coeffs(0.15, d, N) if cc else O.gaussian_analytic_coeffs(0.15, d, N)
    if cc:
        co = (co * np.exp(0.3j)).astype(np.complex64)  # deliberately not Hermitian-even
    if symmetric:
        y = T.nfft_fastsum(cuda(x), cuda(co), cuda(src), batch=cuda(sb), cutoff=m)
        ref = O.nfft_fastsum(x, co, src, None, sb, sb, m=m)
    else:
        tgt, tb = make_points(rng, d, B, n // 2 + 1)
        tgt = (tgt * 0.5).astype(np.float32)
        y = T.nfft_fastsum(cuda(x), cuda(co), cuda(src), cuda(tgt), cuda(sb), cuda(tb), cutoff=m)
        ref = O.nfft_fastsum(x, co, src, tgt, sb, tb, m=m)
    assert y.shape == ref.shape and y.cpu().numpy().dtype == ref.dtype
    assert O.rel_l2(y.cpu().numpy(), ref) < TOL


# ---------------------------------------------------------------------------------------------
# golden vectors of the compiled reference
# ---------------------------------------------------------------------------------------------
def _opt(z, k):
    return cuda(z[k]) if k in z.files else None


@pytest.mark.parametrize("fname", golden_files("adjoint") + golden_files("forward") + golden_files("fastsum"))
def test_engine_matches_reference_golden(fname):
    z = np.load(os.path.join(GOLDEN_DIR, fname))
    op = str(z["op"])
    if op == "adjoint":
        y = T.nfft_adjoint(cuda(z["x"]), cuda(z["pos"]), _opt(z, "batch"), int(z["N"]), int(z["m"]), bool(z["real_output"]))
    elif op == "forward":
        y = T.nfft_forward(cuda(z["x"]), cuda(z["pos"]), _opt(z, "batch"), int(z["m"]), bool(z["real_output"]))
    else:
        y = T.nfft_fastsum(cuda(z["x"]), cuda(z["coeffs"]), cuda(z["sources"]), batch=_opt(z, "source_batch"), cutoff=int(z["m"]))
    assert tuple(y.shape) == z["y"].shape
    assert O.rel_l2(y.cpu().numpy(), z["y"]) < TOL


# ---------------------------------------------------------------------------------------------
# integer work: bit-exact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,N,m,B,n", [(1, 1024, 8, 5, 3000), (2, 256, 4, 3, 20000), (3, 128, 4, 2, 150000), (3, 16, 3, 2, 999), (2, 64, 3, 300, 50),
                                       (3, 16, 4, 1, 40000),   # dense (> 1 point per cell): all 7 fine key bits
                                       (3, 64, 4, 1, 3000)])   # 9 tile bits: 7 spare bits in the second radix pass
def test_binning_is_bit_exact(d, N, m, B, n):
    rng = np.random.default_rng(d + N)
    pos, batch = make_points(rng, d, B, n, ragged=True)
    pos[::7] += 1.0  # some points outside [-1/2, 1/2): wrapped cell
    nn = pos.shape[0]
    L = _lib.lib()
    keys = torch.zeros(nn, dtype=torch.int32, device=DEV)
    perm = torch.zeros(nn, dtype=torch.int32, device=DEV)
    tile = (ctypes.c_int32 * 3)()
    nb = L.nfftb200_workspace_bytes(_lib.OP_SORT, nn, 0, d, N, m, B, 1, 0)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    tp, tb = cuda(pos), cuda(batch)
    for _ in range(2):  # twice: the permutation must be reproducible
        _lib.check(L.nfftb200_sort_points(tp.data_ptr(), tb.data_ptr(), keys.data_ptr(), perm.data_ptr(),
                                          ctypes.cast(tile, ctypes.c_void_p), nn, d, N, m, B, 1, 0, ws.data_ptr(),
                                          ws.numel(), torch.cuda.current_stream().cuda_stream), "sort")
        torch.cuda.synchronize()
        geo = _lib.geometry(d, N, m, B, 1, 0, nn)
        okeys = O.sort_keys(pos, batch, N, list(tile)[:d][::-1], geo["fine_bits"], (geo["scz"], geo["scy"], geo["scx"]))
        assert np.array_equal(keys.cpu().numpy().astype(np.int64), okeys)
        assert np.array_equal(perm.cpu().numpy().astype(np.int64), O.stable_permutation(okeys))


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def test_empty_and_ragged_inputs():
    rng = np.random.default_rng(3)
    # no points at all: adjoint is zero, forward is empty
    y = T.nfft_adjoint(torch.zeros(0, 2, device=DEV), torch.zeros(0, 2, device=DEV), None, 16, 3)
    assert y.shape == (1, 16, 16, 2) and float(y.abs().max()) == 0.0
    f = T.nfft_forward(torch.zeros(1, 16, 16, dtype=torch.complex64, device=DEV), torch.zeros(0, 2, device=DEV), None, 3)
    assert f.shape == (0,)
    # ragged batch with an empty point set
    pos, batch = make_points(rng, 2, 5, 300, ragged=True)
    x = make_values(rng, (pos.shape[0], 2), False)
    y = T.nfft_adjoint(cuda(x), cuda(pos), cuda(batch), 16, 4)
    ref = O.nfft_adjoint(x, pos, batch, 16, 4)
    assert O.rel_l2(y.cpu().numpy(), ref) < TOL
    assert float(y[2].abs().max()) == 0.0  # the empty set
    yf = T.nfft_forward(y, cuda(pos), cuda(batch), 4)
    assert O.rel_l2(yf.cpu().numpy(), O.nfft_forward(ref, pos, batch, 4)) < TOL


def test_one_dimensional_x_and_no_batch():
    rng = np.random.default_rng(4)
    pos = rng.random((400, 3), dtype=np.float32) - 0.5
    x = make_values(rng, (400,), False)
    y = T.nfft_adjoint(cuda(x), cuda(pos), None, 16, 3)
    assert y.shape == (1, 16, 16, 16)
    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, None, 16, 3)) < TOL
    f = T.nfft_forward(y, cuda(pos), None, 3, real_output=True)
    assert f.shape == (400,)
    assert O.rel_l2(f.cpu().numpy(), O.nfft_forward(y.cpu().numpy(), pos, None, 3, real_output=True)) < TOL
    # multi-dimensional channel shape [n, 2, 3]
    x3 = make_values(rng, (400, 2, 3), True)
    y3 = T.nfft_adjoint(cuda(x3), cuda(pos), None, 8, 2)
    assert y3.shape == (1, 8, 8, 8, 2, 3)
    assert O.rel_l2(y3.cpu().numpy(), O.nfft_adjoint(x3, pos, None, 8, 2)) < TOL


def test_periodic_wrap_and_boundary_points():
    rng = np.random.default_rng(5)
    pos = rng.random((600, 2), dtype=np.float32) - 0.5
    pos[0] = [-0.5, -0.5]
    pos[1] = [np.nextafter(np.float32(0.5), np.float32(0)), 0.0]
    pos[2] = [0.0, -0.5]
    x = make_values(rng, (600, 1), False)
    y = T.nfft_adjoint(cuda(x), cuda(pos), None, 32, 4)
    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, None, 32, 4)) < TOL
    moved = pos.copy()
    moved[:100] += 1.0
    moved[100:150] -= 1.0
    y2 = T.nfft_adjoint(cuda(x), cuda(moved), None, 32, 4)
    assert O.rel_l2(y2.cpu().numpy(), y.cpu().numpy()) < TOL
    f2 = T.nfft_forward(y, cuda(moved), None, 4)
    assert O.rel_l2(f2.cpu().numpy(), T.nfft_forward(y, cuda(pos), None, 4).cpu().numpy()) < TOL


def test_all_points_in_one_cell():
    """Maximal collision: every point in the same oversampled cell (one tile, many chunks)."""
    rng = np.random.default_rng(6)
    n = 20000
    pos = (0.123 + 1e-4 * rng.random((n, 3))).astype(np.float32)
    x = make_values(rng, (n, 1), False)
    y = T.nfft_adjoint(cuda(x), cuda(pos), None, 16, 3)
    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, None, 16, 3)) < TOL
    f = T.nfft_forward(y, cuda(pos), None, 3)
    assert O.rel_l2(f.cpu().numpy(), O.nfft_forward(y.cpu().numpy(), pos, None, 3)) < TOL


def test_clustered_points_split_heavy_columns():
    """A tight Gaussian cluster: a few supercell columns of the 3D register-stencil sweep hold most of a
    tile's points, so they are split into z-range work units (window_reg.cuh make_units)."""
    rng = np.random.default_rng(16)
    n, N, m = 30000, 32, 4
    pos = (np.array([0.11, -0.2, 0.31]) + 0.03 * rng.standard_normal((n, 3))).astype(np.float32)
    pos = (((pos + 0.5) % 1.0) - 0.5).astype(np.float32)
    x = make_values(rng, (n, 1), False)
    y = T.nfft_adjoint(cuda(x), cuda(pos), None, N, m)
    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, None, N, m)) < TOL
    f = T.nfft_forward(y, cuda(pos), None, m, real_output=True)
    assert O.rel_l2(f.cpu().numpy(), O.nfft_forward(y.cpu().numpy(), pos, None, m, real_output=True)) < TOL


def test_non_contiguous_inputs_and_side_stream():
    rng = np.random.default_rng(7)
    pos = rng.random((500, 2), dtype=np.float32) - 0.5
    x = make_values(rng, (500, 4), False)
    tp = cuda(np.concatenate([pos, pos], axis=1))[:, :2]  # strided view
    tx = cuda(np.concatenate([x, x], axis=1))[:, ::2][:, :2]
    assert not tp.is_contiguous() and not tx.is_contiguous()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y = T.nfft_adjoint(tx, tp, None, 16, 3)
    s.synchronize()
    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(tx.cpu().numpy(), pos, None, 16, 3)) < TOL


def test_input_validation_raises():
    pos = torch.rand(10, 2, device=DEV) - 0.5
    x = torch.rand(10, device=DEV)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x, pos.double(), None, 16, 3)  # pos must be float32 (core_cuda.cu:49)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x, torch.rand(10, 4, device=DEV), None, 16, 3)  # d <= 3 (core_cuda.cu:53)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x[:5], pos, None, 16, 3)  # x.size(0) == n (core_cuda.cu:83)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x.double(), pos, None, 16, 3)  # float32 | complex64 (core_cuda.cu:77-79)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x, pos, torch.zeros(10, device=DEV, dtype=torch.int32), 16, 3)  # batch int64
    with pytest.raises(RuntimeError):
        T.nfft_forward(torch.rand(1, 16, 8, device=DEV), pos, None, 3)  # all frequency dims == N
    with pytest.raises(RuntimeError):
        T.nfft_forward(torch.rand(2, 16, 16, device=DEV), pos, None, 3)  # x.size(0) == batch size
    with pytest.raises(RuntimeError):
        T.nfft_fastsum(x, torch.rand(16, device=DEV), pos)  # coeffs must be d-dimensional
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x, pos, None, 16, 9)  # cutoff range


# ---------------------------------------------------------------------------------------------
# approximation error against the exact NDFT: no worse than the reference's algorithm at each m
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,N", [(1, 64), (2, 32), (3, 16)])
@pytest.mark.parametrize("m", [2, 3, 4, 6])
def test_error_vs_ndft_not_worse_than_reference_algorithm(d, N, m):
    if m >= N // 2:
        pytest.skip("m < N/2")
    rng = np.random.default_rng(d * 10 + m)
    pos, batch = make_points(rng, d, 2, 300)
    x = make_values(rng, (600, 2), False)
    exact = O.ndft_adjoint(x, pos, batch, N)
    ours = O.rel_l2(T.nfft_adjoint(cuda(x), cuda(pos), cuda(batch), N, m).cpu().numpy(), exact)
    theirs = O.rel_l2(O.nfft_adjoint(x, pos, batch, N, m, prec="f64"), exact)
    assert ours <= theirs * 1.05 + 2e-6
    xh = make_values(rng, (2,) + (N,) * d + (2,), True)
    exact = O.ndft_forward(xh, pos, batch)
    ours = O.rel_l2(T.nfft_forward(cuda(xh), cuda(pos), cuda(batch), m).cpu().numpy(), exact)
    theirs = O.rel_l2(O.nfft_forward(xh, pos, batch, m, prec="f64"), exact)
    assert ours <= theirs * 1.05 + 2e-6


# ---------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------------------------
FULL = {
    "c2": (1, 1024, 8, 2 ** 20, 64, 1),
    "c3": (2, 256, 4, 2 ** 23, 16, 8),
    "c4": (3, 128, 4, 2 ** 24, 4, 1),
}


def _full_inputs(name, clustered=False):
    d, N, m, n, B, C = FULL[name]
    gen = torch.Generator(device=DEV)
    gen.manual_seed(0)
    if clustered:
        centers = torch.rand(64, d, device=DEV, generator=gen) * 0.8 - 0.4
        ids = torch.randint(0, 64, (n,), device=DEV, generator=gen)
        pos = centers[ids] + 0.02 * torch.randn(n, d, device=DEV, generator=gen)
        pos = ((pos + 0.5) % 1.0) - 0.5
    else:
        pos = torch.rand(n, d, device=DEV, generator=gen) - 0.5
    x = torch.randn(n, C, device=DEV, generator=gen)
    batch = torch.arange(n, device=DEV) // (n // B)
    return d, N, m, n, B, C, pos.contiguous(), x, batch


@pytest.mark.parametrize("name,clustered", [("c2", False), ("c3", False), ("c4", False), ("c4", True)])
def test_full_size_properties(name, clustered):
    d, N, m, n, B, C, pos, x, batch = _full_inputs(name, clustered)
    y = T.nfft_adjoint(x, pos, batch, N, m, batch_size=B)
    assert y.shape == (B,) + (N,) * d + (C,)
    # (1) zero frequency = plain sum of the values of each point set (exact for the NDFT)
    zero = y[(slice(None),) + (N // 2,) * d]
    sums = torch.zeros(B, C, device=DEV, dtype=torch.float64).index_add_(0, batch, x.double())
    scale = x.double().pow(2).sum().sqrt()
    assert float((zero.real.double() - sums).abs().max() / scale) < 1e-3
    # (2) Hermitian symmetry of the spectrum of real data: y[-k] = conj(y[k]) for |k| < N/2
    inner = (slice(None),) + (slice(1, None),) * d
    flipped = torch.flip(y[inner], dims=list(range(1, d + 1))).conj()
    assert float((y[inner] - flipped).abs().max() / y.abs().max()) < 1e-4
    # (3) linearity in x
    gen = torch.Generator(device=DEV)
    gen.manual_seed(1)
    x2 = torch.randn(n, C, device=DEV, generator=gen)
    y2 = T.nfft_adjoint(x2, pos, batch, N, m, batch_size=B)
    y12 = T.nfft_adjoint(2.5 * x - x2, pos, batch, N, m, batch_size=B)
    assert float(torch.linalg.vector_norm(y12 - (2.5 * y - y2)) / torch.linalg.vector_norm(y12)) < TOL
    del y12, y2
    # (4) forward of a single frequency is a plane wave: exp(-2 pi i k . pos)
    xh = torch.zeros_like(y)
    k = (1, -2, 3)[:d]
    xh[(slice(None),) + tuple(N // 2 + kk for kk in k)] = 1.0
    f = T.nfft_forward(xh, pos, batch, m, batch_size=B)
    phase = -2 * np.pi * sum(kk * pos[:, a].double() for a, kk in enumerate(k))
    wave = torch.polar(torch.ones_like(phase), phase)
    bound = 5e-4 if m <= 4 else 2e-5  # the NFFT's own error at this cutoff
    assert float((f.reshape(n, C).to(torch.complex128) - wave[:, None]).abs().max()) < bound
    del f, xh
    # (5) adjointness: <adjoint(x), yh> = <x, forward(yh)>, and the C2R path equals Re of the C2C path
    gen.manual_seed(2)
    yh = torch.complex(torch.randn(y.shape, device=DEV, generator=gen), torch.randn(y.shape, device=DEV, generator=gen))
    fy = T.nfft_forward(yh, pos, batch, m, batch_size=B)
    lhs = torch.sum(yh.conj().to(torch.complex128) * y.to(torch.complex128))
    rhs = torch.sum(fy.reshape(n, C).conj().to(torch.complex128) * x.to(torch.complex128))
    assert float((lhs - rhs).abs() / lhs.abs()) < 1e-5
    fr = T.nfft_forward(yh, pos, batch, m, real_output=True, batch_size=B)
    assert float(torch.linalg.vector_norm(fr - fy.real) / torch.linalg.vector_norm(fr)) < TOL


# ---------------------------------------------------------------------------------------------
# point plans: the binning of a point set is made once and reused
# ---------------------------------------------------------------------------------------------
BINNING_LAUNCHES = 5  # the binning alone is at least this many kernels (keys, scans, scatter, items)


def test_plan_reuses_the_binning_between_transforms():
    rng = np.random.default_rng(11)
    pos, batch = make_points(rng, 3, 2, 3000)
    x = make_values(rng, (pos.shape[0], 1), False)
    tp, tb, tx = cuda(pos), cuda(batch), cuda(x)
    T.clear_caches()
    ref_y = O.nfft_adjoint(x, pos, batch, 32, 4)
    ref_f = O.nfft_forward(ref_y, pos, batch, 4, real_output=True)
    # without a plan every transform bins the points
    before = _lib.launch_count()
    y0 = T.nfft_adjoint(tx, tp, tb, 32, 4)
    unplanned = _lib.launch_count() - before
    # with a plan: the first transform bins, the following ones do not
    plan = T.NfftPlan(tp, tb)
    before = _lib.launch_count()
    y = T.nfft_adjoint(tx, plan=plan, N=32, m=4)
    first = _lib.launch_count() - before
    f = T.nfft_forward(y, plan=plan, m=4, real_output=True)  # same points, same tiling: no second binning
    second = _lib.launch_count() - before - first
    y2 = T.nfft_adjoint(tx, tp, tb, 32, 4, plan=plan)        # pos / batch may be passed along with their plan
    third = _lib.launch_count() - before - first - second
    assert first >= unplanned and second <= 4 and third <= 4 and first - third >= BINNING_LAUNCHES, (unplanned, first, second, third)
    assert plan.sorts == 1 and 4 * pos.shape[0] <= plan.nbytes < 8 * pos.shape[0] + (1 << 20)
    assert O.rel_l2(y0.cpu().numpy(), ref_y) < TOL and O.rel_l2(y.cpu().numpy(), ref_y) < TOL
    assert O.rel_l2(y2.cpu().numpy(), ref_y) < TOL and O.rel_l2(f.cpu().numpy(), ref_f) < TOL
    assert plan.flags() == {"dropped": 0, "tma_timeouts": 0, "clustered": 0}
    # a plan refuses other tensors
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(tx, cuda(pos), tb, 32, 4, plan=plan)


def test_stale_plan_is_detected_and_never_writes_out_of_bounds():
    """The caller owns a plan's validity.  If the positions change behind its back (a write through
    `.data`, another CUDA graph, a custom kernel: nothing the engine can see), the transforms drop the points
    they find outside their tile and count them instead of indexing shared memory out of bounds."""
    rng = np.random.default_rng(12)
    for d, N, m in [(3, 32, 4), (2, 32, 4), (1, 1024, 8), (3, 32, 6), (2, 128, 2)]:  # reg 3D / reg 2D / 1D / team 3D / team 2D, several tiles each
        pos, batch = make_points(rng, d, 2, 3000)
        x = make_values(rng, (pos.shape[0], 1), False)
        tp, tb, tx = cuda(pos), cuda(batch), cuda(x)
        plan = T.NfftPlan(tp, tb)
        y = T.nfft_adjoint(tx, plan=plan, N=N, m=m)
        assert plan.dropped_points() == 0
        assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, batch, N, m)) < TOL
        tp.data.copy_(cuda(rng.random(pos.shape, dtype=np.float32) - 0.5))  # new positions, same tensor
        T.nfft_adjoint(tx, plan=plan, N=N, m=m)
        T.nfft_forward(y, plan=plan, m=m)
        torch.cuda.synchronize()
        assert plan.dropped_points() > 0, (d, N, m)
        # a fresh plan of the new positions is exact again
        fresh = T.NfftPlan(tp, tb)
        y2 = T.nfft_adjoint(tx, plan=fresh, N=N, m=m)
        assert O.rel_l2(y2.cpu().numpy(), O.nfft_adjoint(x, tp.cpu().numpy(), batch, N, m)) < TOL
        assert fresh.dropped_points() == 0


def test_gram_matrix_bins_its_points_once():
    """`A @ x` of the kernel-matrix layer: the first product bins sources and targets, products 2..k launch
    no binning kernel (the reference recomputes its per-point scratch every time, core_cuda.cu:188-211)."""
    rng = np.random.default_rng(13)
    src = cuda((rng.random((4000, 3), dtype=np.float32) - 0.5) * 0.5)
    tgt = cuda((rng.random((2500, 3), dtype=np.float32) - 0.5) * 0.5)
    co = T.gaussian_analytic_coeffs(0.1, 3, 32)
    for A, n_in in [(T.GramMatrix(co, src, cutoff=4), 4000), (T.GramMatrix(co, src, tgt, cutoff=4), 4000)]:
        x = cuda(make_values(rng, (n_in, 2), False))
        before = _lib.launch_count()
        y1 = A @ x
        first = _lib.launch_count() - before
        counts = []
        for _ in range(3):
            before = _lib.launch_count()
            y = A @ x
            counts.append(_lib.launch_count() - before)
        assert max(counts) == min(counts) and first - counts[0] >= BINNING_LAUNCHES, (first, counts)
        assert torch.equal(y, y1) or O.rel_l2(y.cpu().numpy(), y1.cpu().numpy()) < 1e-6
        exact = T.ndft_fastsum(x, co, src, None if A.is_symmetric() else tgt)
        assert O.rel_l2(y.cpu().numpy(), exact.cpu().numpy()) < 5e-4  # the NFFT's own error at m = 4
        # the transposed matrix shares the binnings
        before = _lib.launch_count()
        A.T @ cuda(make_values(rng, (A.shape[1], 2), False))
        assert _lib.launch_count() - before == counts[0]


def test_batch_offsets_equal_batch_vector():
    """`batch_ptr` (B + 1 offsets) instead of the per-point int64 vector: identical binning, identical result."""
    rng = np.random.default_rng(14)
    for d, N, m in [(3, 32, 4), (2, 32, 4), (1, 256, 8)]:
        pos, batch = make_points(rng, d, 5, 700, ragged=True)
        x = make_values(rng, (pos.shape[0], 2), False)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(batch, minlength=5))]).astype(np.int64)
        tp, tb, tx, tptr = cuda(pos), cuda(batch), cuda(x), cuda(ptr)
        y1 = T.nfft_adjoint(tx, tp, tb, N, m)
        y2 = T.nfft_adjoint(tx, tp, batch_ptr=tptr, N=N, m=m)
        assert O.rel_l2(y2.cpu().numpy(), O.nfft_adjoint(x, pos, batch, N, m)) < TOL
        assert O.rel_l2(y2.cpu().numpy(), y1.cpu().numpy()) < 2e-6
        f2 = T.nfft_forward(y1, plan=T.NfftPlan(tp, batch_ptr=tptr), m=m, real_output=True)
        assert O.rel_l2(f2.cpu().numpy(), O.nfft_forward(y1.cpu().numpy(), pos, batch, m, real_output=True)) < TOL


def test_int64_index_variants_of_the_spectral_kernels():
    """The spectral kernels switch to 64-bit index arithmetic when B*C*M^d >= 2^31 (e.g. N=512 in 3D with 8
    grids); the variants are forced here on small transforms and must reproduce the oracle."""
    rng = np.random.default_rng(15)
    L = _lib.lib()
    L.nfftb200_debug_force_int64(1)
    try:
        for d, N, m, C in [(1, 64, 4, 2), (2, 32, 4, 3), (3, 16, 3, 1)]:
            pos, batch = make_points(rng, d, 2, 500)
            for cplx in (False, True):
                x = make_values(rng, (pos.shape[0], C), cplx)
                for ro in (False, True):
                    y = T.nfft_adjoint(cuda(x), cuda(pos), cuda(batch), N, m, real_output=ro)
                    assert O.rel_l2(y.cpu().numpy(), O.nfft_adjoint(x, pos, batch, N, m, real_output=ro)) < TOL
                    xh = make_values(rng, (2,) + (N,) * d + (C,), cplx)
                    f = T.nfft_forward(cuda(xh), cuda(pos), cuda(batch), m, real_output=ro)
                    assert O.rel_l2(f.cpu().numpy(), O.nfft_forward(xh, pos, batch, m, real_output=ro)) < TOL
                co = O.gaussian_interpolated_coeffs(0.15, d, N)
                src = (pos * 0.5).astype(np.float32)
                s_ = T.nfft_fastsum(cuda(x), cuda(co), cuda(src), batch=cuda(batch), cutoff=m)
                assert O.rel_l2(s_.cpu().numpy(), O.nfft_fastsum(x, co, src, None, batch, batch, m=m)) < TOL
    finally:
        L.nfftb200_debug_force_int64(0)


def test_fft_work_area_comes_from_the_callers_workspace():
    """Two streams run the same cached cuFFT handle at once without sharing scratch, and the hidden
    allocations of the handles stay small (the work areas live in the per-stream torch workspaces)."""
    rng = np.random.default_rng(17)
    d, N, m, B = 3, 64, 4, 2
    pos, batch = make_points(rng, d, B, 20000)
    x1, x2 = make_values(rng, (pos.shape[0], 1), False), make_values(rng, (pos.shape[0], 1), False)
    tp, tb, t1, t2 = cuda(pos), cuda(batch), cuda(x1), cuda(x2)
    T.nfft_adjoint(t1, tp, tb, N, m)  # make the handle
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(4):
        with torch.cuda.stream(s1):
            a = T.nfft_adjoint(t1, tp, tb, N, m, batch_size=B)
        with torch.cuda.stream(s2):
            b = T.nfft_adjoint(t2, tp, tb, N, m, batch_size=B)
        outs.append((a, b))
    torch.cuda.synchronize()
    r1, r2 = O.nfft_adjoint(x1, pos, batch, N, m), O.nfft_adjoint(x2, pos, batch, N, m)
    for a, b in outs:
        assert O.rel_l2(a.cpu().numpy(), r1) < TOL and O.rel_l2(b.cpu().numpy(), r2) < TOL
    assert _lib.lib().nfftb200_plan_cache_size() >= 1


# ---------------------------------------------------------------------------------------------
# CUDA graphs: the path makes no host synchronisation, allocation or plan creation after its first
# call (the reference syncs the device after every kernel, csrc/cuda/cuda_utils.cu:7-14), so a
# transform pair can be captured once and replayed on new data in the same buffers
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,N,m,B,n", [(1, 256, 8, 4, 3000), (3, 32, 4, 2, 4000)])
def test_cuda_graph_capture_and_replay(d, N, m, B, n):
    rng = np.random.default_rng(21)
    pos, batch = make_points(rng, d, B, n)
    x = make_values(rng, (pos.shape[0], 1), False)
    tp, tb, tx = cuda(pos), cuda(batch), cuda(x)

    def pair():
        plan = T.NfftPlan(tp, tb, batch_size=B)  # batch_size: no batch[-1].item() sync; binned inside the capture
        y = T.nfft_adjoint(tx, plan=plan, N=N, m=m)
        return y, T.nfft_forward(y, plan=plan, m=m, real_output=True)

    graphed = T.GraphedTransforms(pair)  # warm-up on a side stream (plans, workspace), then the capture
    gy, gf = graphed.outputs
    assert graphed.kernels_per_replay > 0
    # new points and values in the captured buffers; the replay bins the new points on the device
    pos2 = rng.random(pos.shape, dtype=np.float32) - 0.5
    x2 = make_values(rng, x.shape, False)
    tp.copy_(cuda(pos2))
    tx.copy_(cuda(x2))
    before = _lib.launch_count()
    assert graphed.replay()[0] is gy
    torch.cuda.synchronize()
    assert _lib.launch_count() == before  # no host-side launches: the whole pair is one graph launch
    ref_y = O.nfft_adjoint(x2, pos2, batch, N, m)
    assert O.rel_l2(gy.cpu().numpy(), ref_y) < TOL
    assert O.rel_l2(gf.cpu().numpy(), O.nfft_forward(ref_y, pos2, batch, m, real_output=True)) < TOL
    # eager calls after the replay bin the changed tensors themselves
    y3 = T.nfft_adjoint(tx, tp, tb, N, m)
    assert O.rel_l2(y3.cpu().numpy(), ref_y) < TOL
    # while the graph lives the cuFFT handle cache is pinned; close() releases it
    with pytest.warns(RuntimeWarning):
        T.clear_caches()
    graphed.close()
    T.clear_caches()


# ---------------------------------------------------------------------------------------------
# mixed density (large 3D point sets with the CLUSTERED hint): whether the set really is clustered is decided on
# the device from a sample of the keys; heavy tiles are then swept with 2 x 2 x 2 supercells, the rest with 4 x 4 x 2
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def mixed_mode():
    L = _lib.lib()
    L.nfftb200_debug_mixed(-1, 0, 600)  # hinted sets of every size; a 16^3 tile is heavy from 600 points
    yield
    L.nfftb200_debug_mixed(-1, -1, 0)


def clustered_points(rng, n, B, sigma=0.03, uniform_share=0.2):
    """B point sets: three Gaussian clusters each plus a uniform background, wrapped into [-1/2, 1/2)."""
    nu = int(n * uniform_share)
    per = (n - nu) // 3
    sets = []
    for _ in range(B):
        centres = rng.random((3, 3)) - 0.5
        pts = [centres[k] + sigma * rng.standard_normal((per, 3)) for k in range(3)]
        pts.append(rng.random((n - 3 * per, 3)) - 0.5)
        p = np.concatenate(pts)
        sets.append(((p + 0.5) % 1.0 - 0.5).astype(np.float32))
    pos = np.concatenate(sets)
    pos[pos >= 0.5] = -0.5
    batch = np.repeat(np.arange(B, dtype=np.int64), n)
    return pos, batch


@pytest.mark.parametrize("m,cplx", [(4, False), (3, True)])
def test_mixed_density_transforms(mixed_mode, m, cplx):
    rng = np.random.default_rng(40 + m)
    N, B, n = 32, 2, 15000
    geo = _lib.geometry(3, N, m, B, 1, _lib.CLUSTERED, B * n)
    assert geo["mixed"] == 1 and geo["refine_pass"] == 1 and geo["dense_tile_pts"] == 600
    assert _lib.geometry(3, N, m, B, 1, 0, B * n)["mixed"] == 0  # no hint: the single-sweep path
    for clustered in (True, False):
        if clustered:
            pos, batch = clustered_points(rng, n, B)
        else:
            pos, batch = make_points(rng, 3, B, n)
        x = make_values(rng, (pos.shape[0],), cplx)
        tp, tb, tx = cuda(pos), cuda(batch), cuda(x)
        plan = T.NfftPlan(tp, tb, clustered=True)
        y = T.nfft_adjoint(tx, plan=plan, N=N, m=m)
        ref_y = O.nfft_adjoint(x, pos, batch, N, m)
        assert O.rel_l2(y.cpu().numpy(), ref_y) < TOL
        f = T.nfft_forward(y, plan=plan, m=m)
        assert O.rel_l2(f.cpu().numpy(), O.nfft_forward(ref_y, pos, batch, m)) < TOL
        # the device-side decision: heavy tiles exist only in the clustered set
        assert plan.flags() == {"dropped": 0, "tma_timeouts": 0, "clustered": 1 if clustered else 0}
        # against the single-sweep path (no hint)
        y0 = T.nfft_adjoint(tx, tp, tb, N, m)
        assert O.rel_l2(y0.cpu().numpy(), y.cpu().numpy()) < 2e-6
        # fastsum through the same plan (symmetric: one binning for spread and gather)
        if not cplx and clustered:
            coeffs = cuda(make_values(rng, (N, N, N), False))
            s1 = T.nfft_fastsum(tx, coeffs, tp, batch=tb, cutoff=m, source_plan=plan)
            s0 = T.nfft_fastsum(tx, coeffs, tp, batch=tb, cutoff=m)
            assert O.rel_l2(s1.cpu().numpy(), s0.cpu().numpy()) < 2e-6


def test_mixed_density_binning_is_bit_exact(mixed_mode):
    """The conditional low radix pass: a clustered set is sorted by the full key, a uniform one by the key
    without its low digit (tile + the top fine bits), both stable."""
    rng = np.random.default_rng(77)
    N, m, B, n = 32, 4, 2, 20000
    L = _lib.lib()
    for clustered in (True, False):
        pos, batch = clustered_points(rng, n, B) if clustered else make_points(rng, 3, B, n)
        nn = pos.shape[0]
        keys = torch.zeros(nn, dtype=torch.int32, device=DEV)
        perm = torch.zeros(nn, dtype=torch.int32, device=DEV)
        tile = (ctypes.c_int32 * 3)()
        nb = L.nfftb200_workspace_bytes(_lib.OP_SORT, nn, 0, 3, N, m, B, 1, _lib.CLUSTERED)
        ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
        tp, tb = cuda(pos), cuda(batch)
        _lib.check(L.nfftb200_sort_points(tp.data_ptr(), tb.data_ptr(), keys.data_ptr(), perm.data_ptr(),
                                          ctypes.cast(tile, ctypes.c_void_p), nn, 3, N, m, B, 1, _lib.CLUSTERED,
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "sort")
        torch.cuda.synchronize()
        geo = _lib.geometry(3, N, m, B, 1, _lib.CLUSTERED, nn)
        assert geo["mixed"] == 1 and geo["fine_bits"] == 9 and (geo["scx"], geo["scy"], geo["scz"]) == (2, 2, 2)
        okeys = O.sort_keys(pos, batch, N, list(tile)[::-1], geo["fine_bits"], (2, 2, 2))
        assert np.array_equal(keys.cpu().numpy().astype(np.int64), okeys)
        expect = O.stable_permutation(okeys if clustered else okeys >> 8)
        assert np.array_equal(perm.cpu().numpy().astype(np.int64), expect)


# ---------------------------------------------------------------------------------------------
# pruned real transforms (hand-written X pass with the crop + cuFFT C2C over the kept kx planes) against the plain
# cuFFT R2C / C2R path of the same library
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,N,B,C", [(3, 128, 2, 1), (2, 128, 3, 2), (2, 256, 2, 1), (2, 256, 1, 8)])
def test_pruned_fft_matches_cufft_path(d, N, B, C):
    rng = np.random.default_rng(100 * d + N + C)
    m, n = 4, 20000
    pos, batch = make_points(rng, d, B, n)
    x = make_values(rng, (pos.shape[0], C) if C > 1 else (pos.shape[0],), False)
    spec = make_values(rng, (B,) + (N,) * d + ((C,) if C > 1 else ()), True)
    tp, tb, tx, ts = cuda(pos), cuda(batch), cuda(x), cuda(spec)
    L = _lib.lib()
    out = {}
    try:
        for mode in (1, 0):
            L.nfftb200_debug_pruned_fft(mode)
            out[mode] = (T.nfft_adjoint(tx, tp, tb, N, m).cpu().numpy(),                     # real grid: R2C
                         T.nfft_adjoint(tx, tp, tb, N, m, real_output=True).cpu().numpy(),
                         T.nfft_forward(ts, tp, tb, m, real_output=True).cpu().numpy())      # real grid: C2R
    finally:
        L.nfftb200_debug_pruned_fft(-1)
    for a, b in zip(out[1], out[0]):
        assert a.shape == b.shape and O.rel_l2(a, b) < 2e-6, O.rel_l2(a, b)
    # and against the oracle on the smallest of them
    if d == 2 and N == 128 and C == 2:
        assert O.rel_l2(out[1][0], O.nfft_adjoint(x, pos, batch, N, m)) < TOL
        assert O.rel_l2(out[1][2], O.nfft_forward(spec, pos, batch, m, real_output=True)) < TOL
