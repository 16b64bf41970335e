"""CPU tests of the host-side Python layer that needs no GPU: exact direct sums, coefficient
helpers, point scaling utilities and the implicit-matrix classes."""
import numpy as np
import pytest
import torch

from oracle import nfft_oracle as O
import torch_nfft_b200 as T
from torch_nfft_b200 import utils
from torch_nfft_b200.matrices import AbstractMatrix, AdjacencyMatrix


def test_ndft_matches_oracle():
    rng = np.random.default_rng(0)
    pos = rng.random((60, 2), dtype=np.float32) - 0.5
    batch = np.repeat(np.arange(2), 30)
    x = rng.standard_normal((60, 3)).astype(np.float32)
    y = T.ndft_adjoint(torch.from_numpy(x), torch.from_numpy(pos), torch.from_numpy(batch), N=8)
    assert O.rel_l2(y.numpy(), O.ndft_adjoint(x, pos, batch, 8)) < 5e-6
    f = T.ndft_forward(y, torch.from_numpy(pos), torch.from_numpy(batch))
    assert O.rel_l2(f.numpy(), O.ndft_forward(y.numpy(), pos, batch)) < 5e-6
    co = T.gaussian_analytic_coeffs(0.2, 2, 8, device="cpu")
    s = T.ndft_fastsum(torch.from_numpy(x), co, torch.from_numpy(pos), batch=torch.from_numpy(batch))
    assert O.rel_l2(s.numpy(), O.ndft_fastsum(x, co.numpy(), pos, None, batch, batch)) < 5e-6
    s1 = T.ndft_fastsum(torch.from_numpy(x[:, 0].copy()), co, torch.from_numpy(pos))  # 1-D x accepted
    assert s1.shape == (60,)


def test_exact_matrices():
    rng = np.random.default_rng(1)
    pos = torch.from_numpy((rng.random((40, 2), dtype=np.float32) - 0.5) * 0.5)
    dense = T.exact_gaussian_matrix(0.1, pos)
    ref = torch.exp(-(pos[None] - pos[:, None]).pow(2).sum(-1) / 0.01)
    assert torch.allclose(dense, ref, atol=1e-5)
    co = T.gaussian_analytic_coeffs(0.1, 2, 32, device="cpu")
    trig = T.exact_trigonometric_matrix(co, pos).real
    assert (trig - ref).abs().max() < 1e-4
    batch = torch.arange(40) // 20
    blk = T.exact_gaussian_matrix(0.2, pos, batch=batch)
    assert blk.shape == (40, 40) and blk[:20, 20:].abs().max() == 0


@pytest.mark.parametrize("dim,N", [(1, 16), (2, 8), (3, 4)])
def test_coefficient_helpers_match_oracle(dim, N):
    assert np.allclose(T.gaussian_analytic_coeffs(0.2, dim, N, device="cpu").numpy(), O.gaussian_analytic_coeffs(0.2, dim, N), rtol=1e-5)
    for p in (-1, 0):
        a = T.gaussian_interpolated_coeffs(0.2, dim, N, p, device="cpu").numpy()
        assert a.dtype == np.complex64 and np.allclose(a, O.gaussian_interpolated_coeffs(0.2, dim, N, p), atol=1e-6)
    assert np.allclose(T.interpolation_grid(dim, N, device="cpu").numpy(), O.interpolation_grid(dim, N))
    assert np.allclose(T.radial_interpolation_grid(dim, N, device="cpu").numpy(), O.radial_interpolation_grid(dim, N), atol=1e-6)
    with pytest.raises(RuntimeError):
        T.gaussian_interpolated_coeffs(0.2, dim, N, 1, device="cpu")  # p > 0 unsupported (core_cuda.cu:890)


def test_point_utils_batched_equals_per_set():
    torch.manual_seed(0)
    src = torch.randn(50, 3) * 3 + 1
    tgt = torch.randn(30, 3) * 2 - 1
    sb = torch.sort(torch.randint(0, 3, (50,))).values
    tb = torch.sort(torch.randint(0, 3, (30,))).values
    sb[-1] = tb[-1] = 2
    c = utils.compute_points_center(src, tgt, sb, tb)
    r = utils.compute_points_radius(src, tgt, sb, tb, norm="infinity")
    for b in range(3):
        cb = utils.compute_points_center(src[sb == b], tgt[tb == b])
        assert torch.allclose(c[b], cb)
        assert abs(r[b].item() - utils.compute_points_radius(src[sb == b], tgt[tb == b], norm="infinity")) < 1e-6
    s2, t2 = utils.shift_points_by_center(src, tgt, sb, tb)
    s3, t3 = utils.scale_points_by_norm(s2, t2, sb, tb, factor=0.25, norm="euclidean")
    for b in range(3):
        assert abs(max(s3[sb == b].norm(dim=1).max(), t3[tb == b].norm(dim=1).max()) - 0.25) < 1e-6
    s4, none = utils.scale_points_by_norm(src, factor=0.25, norm="infinity")
    assert none is None and abs(s4.abs().max().item() - 0.25) < 1e-6
    with pytest.raises(ValueError):
        utils.compute_points_radius(src, norm="manhattan")


class _Dense(AbstractMatrix):
    def __init__(self, A):
        super().__init__(A.shape, A.device)
        self.A = A

    def apply(self, x):
        return self.A @ x

    def is_symmetric(self):
        return True


@pytest.mark.parametrize("normalization", [None, "sym", "left", "rw", "right"])
@pytest.mark.parametrize("shift", [None, "laplacian", "signless"])
def test_adjacency_matrix_against_dense(normalization, shift):
    torch.manual_seed(1)
    W = torch.rand(12, 12)
    W = W + W.T
    A = AdjacencyMatrix(_Dense(W), diagonal_offset=0.5, normalization=normalization, shift=shift)
    Wd = W + 0.5 * torch.eye(12)
    deg = Wd.sum(1)
    norm = {"rw": "left"}.get(normalization, normalization)
    if norm == "sym":
        E = Wd / deg.sqrt()[:, None] / deg.sqrt()[None, :]
    elif norm == "left":
        E = Wd / deg[:, None]
    elif norm == "right":
        E = Wd / deg[None, :]
    else:
        E = Wd
    if shift is not None:
        D = torch.diag(deg) if norm is None else torch.eye(12)
        E = D + E if shift == "signless" else D - E
    assert torch.allclose(A.to_dense(), E, atol=1e-5)
    x = torch.rand(12, 3)
    assert torch.allclose(A @ x, E @ x, atol=1e-5)
    assert torch.allclose(A.T.to_dense(), E.T, atol=1e-5)
    assert torch.allclose(A.column_sums(), E.sum(0), atol=1e-4)


def test_adjacency_matrix_rejects_bad_arguments():
    W = _Dense(torch.eye(3))
    with pytest.raises(ValueError):
        AdjacencyMatrix(W, normalization="bogus")
    with pytest.raises(ValueError):
        AdjacencyMatrix(W, shift="bogus")
    with pytest.warns(RuntimeWarning):
        AdjacencyMatrix(_Dense(-torch.eye(3)), normalization="sym")


def test_graphed_transforms_need_a_gpu():
    """No CPU fallback anywhere on the hot path: without a CUDA device the graph helper refuses to run."""
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError):
        T.GraphedTransforms(lambda: None)


def test_release_stream_workspace_forgets_only_that_stream():
    from torch_nfft_b200 import nfft
    saved = dict(nfft._workspaces)
    try:
        nfft._workspaces[(0, 111)], nfft._workspaces[(0, 222)] = "a", "b"
        nfft.release_stream_workspace(0, 111)
        assert (0, 111) not in nfft._workspaces
        assert nfft._workspaces[(0, 222)] == "b"
        nfft.release_stream_workspace(0, 333)  # unknown stream: no error
    finally:
        nfft._workspaces.clear(); nfft._workspaces.update(saved)


def test_plan_and_batch_ptr_argument_checks_need_no_gpu():
    """NfftPlan / batch_ptr validate like check_point_input (core_cuda.cu:38-66): CPU tensors raise."""
    pos = torch.rand(10, 2) - 0.5
    with pytest.raises(RuntimeError):
        T.NfftPlan(pos)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(torch.rand(10), pos, batch_ptr=torch.tensor([0, 10]))
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(torch.rand(10))  # neither pos nor plan


def test_plan_cache_pin_blocks_clear():
    """Captured CUDA graphs pin the cuFFT handle cache (their kernels reference the handles' twiddle
    tables): clearing is refused with a message until the pin is released.  Host-only bookkeeping."""
    from torch_nfft_b200 import _lib
    L = _lib.lib()
    assert L.nfftb200_plan_cache_clear() == 0
    assert L.nfftb200_plan_cache_pin(1) == 1
    try:
        assert L.nfftb200_plan_cache_clear() != 0 and b"pinned" in L.nfftb200_last_error()
        with pytest.warns(RuntimeWarning):
            T.clear_caches()
    finally:
        assert L.nfftb200_plan_cache_pin(-1) == 0
    assert L.nfftb200_plan_cache_clear() == 0


def test_point_limits_are_rejected():
    """n is int64 across the ABI but permutation / bin offsets are 32-bit: oversize inputs are refused."""
    from torch_nfft_b200 import _lib
    L = _lib.lib()
    assert L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, 2 ** 32, 0, 3, 128, 4, 4, 1, 0) == 0
    assert b"points" in L.nfftb200_last_error()
    assert L.nfftb200_plan_bytes(2 ** 33, 0, 1, 1024, 8, 1, 1, 0) == 0
