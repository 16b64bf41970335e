"""Public functional API + autograd of the B200 NFFT engine.

Drop-in for reference `torch_nfft/nfft.py:11-179`: the same three functions with the same
argument meaning, the same backward rules (adjoint <-> forward, fastsum <-> swapped fastsum) and
the same input checks as `csrc/cuda/core_cuda.cu:38-115`, but the work is done by the C-ABI
library `libnfft_b200.so` (hand-written sm_100a kernels + cached cuFFT plans) on the current
stream, with outputs and workspace owned by PyTorch.

Additions over the reference (all optional, defaults reproduce the reference):
  * `N=` / `m=` keyword aliases of `bandwidth=` / `cutoff=` (the reference's own test scripts
    use them: test/test_adjoint.py:32, test/test_forward.py:34);
  * `batch_size=` to skip the device->host read of `batch[-1]` (core_cuda.cu:60);
  * gradients w.r.t. `pos` for nfft_forward / nfft_adjoint (the reference returns None,
    nfft.py:28,54) -- see `pos_grad`.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import _lib

_workspaces = {}
# What the sort region of each workspace currently holds: the adjoint and forward transforms of the
# same point set (forward + backward of autograd, iterative solvers) then skip the binning pass.
# The reference recomputes its per-point scratch on every call (core_cuda.cu:188-211, 461-484).
_sorted_points = {}
_PLAN_REUSE = os.environ.get("NFFTB200_NO_PLAN_REUSE") is None


def _workspace(nbytes: int, device: torch.device, keep_sorted: bool = False) -> torch.Tensor:
    """Grow-only scratch tensor per (device, stream); allocated by the torch caching allocator.
    Unless `keep_sorted`, the caller is about to overwrite the sort region: forget what it held."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    if not keep_sorted:
        _sorted_points.pop(key, None)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspaces.pop(key, None)
        _sorted_points.pop(key, None)
        ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def clear_caches():
    """Drop cached workspaces, remembered point sorts and cuFFT plans."""
    _workspaces.clear()
    _sorted_points.clear()
    _lib.lib().nfftb200_plan_cache_clear()


def release_stream_workspace(device_index: int, cuda_stream: int):
    """Drop the scratch tensor (and remembered sort) kept for one (device, stream): called when the
    owner of a private stream goes away (`GraphedTransforms.close`)."""
    key = (device_index, cuda_stream)
    _workspaces.pop(key, None)
    _sorted_points.pop(key, None)


def forget_sorted_points():
    """Forget which point sets the workspaces hold a sort for (the next transform bins again)."""
    _sorted_points.clear()


def _points_identity(pos, batch, geometry):
    """Identity of a sorted point set: the tensor objects (weakly held) at their current version plus
    the tiling the sort was made for.  In-place writes through torch bump `_version`."""
    return (weakref.ref(pos), pos._version, None if batch is None else weakref.ref(batch),
            None if batch is None else batch._version, geometry)


def _same_points(ident, pos, batch, geometry):
    if ident is None or not _PLAN_REUSE:
        return False
    rpos, vpos, rbatch, vbatch, geom = ident
    if rpos() is not pos or vpos != pos._version or geom != geometry:
        return False
    if batch is None:
        return rbatch is None
    return rbatch is not None and rbatch() is batch and vbatch == batch._version


def _ws_key(device):
    return (device.index, torch.cuda.current_stream(device).cuda_stream)


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _check(cond, msg):
    if not cond:
        raise RuntimeError("torch_nfft_b200: " + msg)


def _check_points(pos, batch, batch_size=None):
    """check_point_input (core_cuda.cu:38-66)."""
    _check(isinstance(pos, torch.Tensor) and pos.is_cuda, "pos must be a CUDA tensor")
    _check(pos.dim() == 2, "pos must have shape [n, d]")
    _check(pos.dtype == torch.float32, "pos must be float32")
    n, d = pos.shape
    _check(1 <= d <= 3, "spatial dimension must be 1, 2 or 3")
    if batch is not None:
        _check(batch.is_cuda and batch.device == pos.device, "batch must be a CUDA tensor on the device of pos")
        _check(batch.dim() == 1 and batch.dtype == torch.int64, "batch must be a 1-D int64 tensor")
        _check(batch.numel() == n, "batch must have one entry per point")
        if batch_size is None:
            batch_size = int(batch[-1].item()) + 1 if n > 0 else 1  # core_cuda.cu:60
        batch = batch.contiguous()
    else:
        batch_size = 1
    return pos.contiguous(), batch, int(n), int(d), int(batch_size)


def _check_cutoff(m, N):
    _check(isinstance(m, int) and 1 <= m <= 8, "cutoff must be an integer in [1, 8]")
    _check(N >= 2 and N % 2 == 0, "bandwidth must be even and >= 2")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# raw operators (same argument order as torch.ops.torch_nfft.*, reference core.cpp:43-121)
# --------------------------------------------------------------------------------------
def _op_adjoint(pos, x, batch, N, m, real_output, batch_size=None):
    pos, batch, n, d, B = _check_points(pos, batch, batch_size)
    _check(x.is_cuda and x.device == pos.device, "x must be a CUDA tensor on the device of pos")
    _check(x.dtype in (torch.float32, torch.complex64), "x must be float32 or complex64")
    _check(x.dim() >= 1 and x.size(0) == n, "x.size(0) must equal the number of points")
    N = int(N)
    _check_cutoff(m, N)
    cols = tuple(x.shape[1:])
    C = 1
    for s in cols:
        C *= s
    x = x.contiguous()
    flags = (_lib.X_COMPLEX if x.is_complex() else 0) | (_lib.Y_REAL if real_output else 0)
    y = torch.empty((B,) + (N,) * d + cols, dtype=torch.float32 if real_output else torch.complex64, device=pos.device)
    if C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(pos.device):
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, n, 0, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, pos.device, keep_sorted=True)
        geometry = tuple(_lib.geometry(d, N, m, B, C, flags & _lib.X_COMPLEX, n).values())
        key = _ws_key(pos.device)
        if _same_points(_sorted_points.get(key), pos, batch, geometry):
            flags |= _lib.PRESORTED
        _sorted_points.pop(key, None)
        _lib.check(L.nfftb200_adjoint(_ptr(pos), _ptr(x), _ptr(batch), _ptr(y), n, d, N, m, B, C, flags,
                                      ws.data_ptr(), ws.numel(), _stream_ptr(pos.device)), "nfft_adjoint")
        _sorted_points[key] = _points_identity(pos, batch, geometry)
    return y


def _op_forward(pos, xhat, batch, m, real_output, batch_size=None):
    pos, batch, n, d, B = _check_points(pos, batch, batch_size)
    # check_spectral_coeffs_input (core_cuda.cu:89-115)
    _check(xhat.is_cuda and xhat.device == pos.device, "x must be a CUDA tensor on the device of pos")
    _check(xhat.dtype in (torch.float32, torch.complex64), "x must be float32 or complex64")
    _check(xhat.dim() >= d + 1, "x must have shape [batch_size, N, ..., N, *columns]")
    _check(xhat.size(0) == B, "x.size(0) must equal the batch size")
    N = int(xhat.size(1))
    _check(all(xhat.size(a) == N for a in range(1, d + 1)), "all frequency dimensions of x must have size N")
    _check_cutoff(m, N)
    cols = tuple(xhat.shape[1 + d:])
    C = 1
    for s in cols:
        C *= s
    xhat = xhat.contiguous()
    flags = (_lib.X_COMPLEX if xhat.is_complex() else 0) | (_lib.Y_REAL if real_output else 0)
    y = torch.empty((n,) + cols, dtype=torch.float32 if real_output else torch.complex64, device=pos.device)
    if n == 0 or C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(pos.device):
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_FORWARD, 0, n, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, pos.device, keep_sorted=True)
        # the gather grid is complex unless real_output: same tiling rule as a complex adjoint
        geometry = tuple(_lib.geometry(d, N, m, B, C, 0 if real_output else _lib.X_COMPLEX, n).values())
        key = _ws_key(pos.device)
        if _same_points(_sorted_points.get(key), pos, batch, geometry):
            flags |= _lib.PRESORTED
        _sorted_points.pop(key, None)
        _lib.check(L.nfftb200_forward(_ptr(pos), _ptr(xhat), _ptr(batch), _ptr(y), n, d, N, m, B, C, flags,
                                      ws.data_ptr(), ws.numel(), _stream_ptr(pos.device)), "nfft_forward")
        _sorted_points[key] = _points_identity(pos, batch, geometry)
    return y


def _op_fastsum(sources, targets, x, coeffs, source_batch, target_batch, m, batch_size=None):
    symmetric = targets is sources  # core_cuda.cu:552
    sources_c, source_batch, n_src, d, B = _check_points(sources, source_batch, batch_size)
    if symmetric:
        targets_c, target_batch, n_tgt = sources_c, source_batch, n_src
    else:
        targets_c, target_batch, n_tgt, d_t, B_t = _check_points(targets, target_batch, batch_size)
        _check(d_t == d, "sources and targets must have the same dimension")
        _check(B_t == B, "sources and targets must have the same batch size")
    _check(x.is_cuda and x.dtype in (torch.float32, torch.complex64), "x must be a float32/complex64 CUDA tensor")
    _check(x.dim() >= 1 and x.size(0) == n_src, "x.size(0) must equal the number of source points")
    _check(coeffs.is_cuda and coeffs.dim() == d, "coeffs must be a d-dimensional CUDA tensor")
    N = int(coeffs.size(0))
    _check(all(coeffs.size(a) == N for a in range(d)), "coeffs must have size N in every dimension")
    _check(coeffs.dtype in (torch.float32, torch.complex64), "coeffs must be float32 or complex64")
    _check_cutoff(m, N)
    cols = tuple(x.shape[1:])
    C = 1
    for s in cols:
        C *= s
    x = x.contiguous()
    coeffs = coeffs.contiguous()
    flags = ((_lib.X_COMPLEX if x.is_complex() else 0) | (_lib.COEFFS_COMPLEX if coeffs.is_complex() else 0)
             | (_lib.SYMMETRIC if symmetric else 0))
    y = torch.empty((n_tgt,) + cols, dtype=x.dtype, device=x.device)
    if n_tgt == 0 or C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(x.device):
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_FASTSUM, n_src, n_tgt, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, x.device)
        _lib.check(L.nfftb200_fastsum(_ptr(sources_c), _ptr(targets_c), _ptr(x), _ptr(coeffs), _ptr(source_batch),
                                      _ptr(target_batch), _ptr(y), n_src, n_tgt, d, N, m, B, C, flags,
                                      ws.data_ptr(), ws.numel(), _stream_ptr(x.device)), "nfft_fastsum")
    return y


# --------------------------------------------------------------------------------------
# gradient w.r.t. the point positions (new capability; the reference returns None)
# --------------------------------------------------------------------------------------
def _freq_ramp(N, d, a, device):
    """-2 pi i k_a on the API frequency grid, broadcastable to [B, N..N, C]."""
    k = torch.arange(-(N // 2), N - N // 2, device=device, dtype=torch.float32)
    shape = [1] * (d + 2)
    shape[1 + a] = N
    return (-2j * torch.pi * k).to(torch.complex64).reshape(shape)


def _forward_pos_grad(pos, xhat, batch, m, dy, batch_size):
    """d/dpos of sum(Re(conj(dy) * forward(xhat))): forward NFFT of (-2 pi i k_a) xhat."""
    n, d = pos.shape
    B, N = xhat.shape[0], xhat.shape[1]
    xh = xhat.reshape(B, *(N,) * d, -1).to(torch.complex64)
    grads = []
    for a in range(d):
        fa = _op_forward(pos, (xh * _freq_ramp(N, d, a, pos.device)).contiguous(), batch, m, False, batch_size)
        g = (fa.reshape(n, -1) * dy.reshape(n, -1).conj()).real.sum(-1) if dy.is_complex() else \
            (fa.reshape(n, -1).real * dy.reshape(n, -1)).sum(-1)
        grads.append(g)
    return torch.stack(grads, dim=-1)


def _adjoint_pos_grad(pos, x, batch, N, m, dy, batch_size):
    """d/dpos of sum(Re(conj(dy) * adjoint(x))): y_k = sum_i x_i e^{+2 pi i k p_i}."""
    n, d = pos.shape
    B = dy.shape[0]
    dyc = dy.reshape(B, *(N,) * d, -1).to(torch.complex64)
    xs = x.reshape(n, -1)
    grads = []
    for a in range(d):
        # sum_k conj(dy_k) (2 pi i k_a) e^{+2 pi i k p}  = conj( forward( (-2 pi i k_a)^* ... ) )
        fa = _op_forward(pos, (dyc * _freq_ramp(N, d, a, pos.device)).contiguous(), batch, m, False, batch_size)
        # fa_i = sum_k dy_k (-2 pi i k_a) e^{-2 pi i k p_i};  conj(fa_i) = sum_k conj(dy_k) (2 pi i k_a) e^{+...}
        g = (fa.reshape(n, -1).conj() * xs).real.sum(-1)
        grads.append(g)
    return torch.stack(grads, dim=-1)


# --------------------------------------------------------------------------------------
# autograd wrappers (reference torch_nfft/nfft.py:11-88)
# --------------------------------------------------------------------------------------
class NfftAdjointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, batch, bandwidth, cutoff, real_output, batch_size):
        y = _op_adjoint(pos, x, batch, bandwidth, cutoff, real_output, batch_size)
        ctx.save_for_backward(pos, batch, x if pos.requires_grad else None)
        ctx.cutoff = cutoff
        ctx.bandwidth = bandwidth
        ctx.real_input = not x.is_complex()
        ctx.batch_size = y.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        pos, batch, x = ctx.saved_tensors
        dx = dpos = None
        if ctx.needs_input_grad[0]:
            # reference nfft.py:26: forward NFFT of dy, real output iff x was real
            dx = _op_forward(pos, dy, batch, ctx.cutoff, ctx.real_input, ctx.batch_size)
        if ctx.needs_input_grad[1]:
            dpos = _adjoint_pos_grad(pos, x, batch, ctx.bandwidth, ctx.cutoff, dy, ctx.batch_size)
        return dx, dpos, None, None, None, None, None


class NfftForwardFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, batch, cutoff, real_output, batch_size):
        y = _op_forward(pos, x, batch, cutoff, real_output, batch_size)
        ctx.save_for_backward(pos, batch, x if pos.requires_grad else None)
        ctx.cutoff = cutoff
        ctx.bandwidth = x.size(1)
        ctx.real_input = not x.is_complex()
        ctx.batch_size = x.shape[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        pos, batch, x = ctx.saved_tensors
        dx = dpos = None
        if ctx.needs_input_grad[0]:
            # reference nfft.py:52: adjoint NFFT of dy, real output iff x was real
            dx = _op_adjoint(pos, dy, batch, ctx.bandwidth, ctx.cutoff, ctx.real_input, ctx.batch_size)
        if ctx.needs_input_grad[1]:
            dpos = _forward_pos_grad(pos, x, batch, ctx.cutoff, dy, ctx.batch_size)
        return dx, dpos, None, None, None, None


class NfftFastsumFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coeffs, sources, targets, source_batch, target_batch, cutoff, batch_size):
        # reference nfft.py:66-73
        assert not coeffs.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. coefficients is not possible"
        assert not sources.requires_grad and not targets.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. sources and targets is not possible"
        assert source_batch is None or not source_batch.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. batches is not possible"
        assert target_batch is None or not target_batch.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. batches is not possible"
        y = _op_fastsum(sources, targets, x, coeffs, source_batch, target_batch, cutoff, batch_size)
        ctx.save_for_backward(sources, targets, coeffs, source_batch, target_batch)
        ctx.symmetric = targets is sources
        ctx.cutoff = cutoff
        ctx.batch_size = batch_size
        return y

    @staticmethod
    def backward(ctx, dy):
        sources, targets, coeffs, source_batch, target_batch = ctx.saved_tensors
        if ctx.symmetric:
            targets = sources
        # reference nfft.py:86: fastsum with sources and targets swapped
        dx = _op_fastsum(targets, sources, dy.contiguous(), coeffs, target_batch, source_batch, ctx.cutoff,
                         ctx.batch_size)
        return dx, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------
# functional API (reference torch_nfft/nfft.py:31,57,91)
# --------------------------------------------------------------------------------------
def nfft_adjoint(x, pos, batch=None, bandwidth=16, cutoff=3, real_output=False, *, N=None, m=None,
                 batch_size=None):
    """Adjoint NFFT  y[b, k + N/2, ...] ~= sum_{i in b} x[i, ...] exp(+2 pi i k . pos[i]).

    x: [n, *cols] float32|complex64, pos: [n, d] float32 in [-1/2, 1/2), batch: [n] int64 sorted.
    Returns [batch_size, N, ..., N, *cols] complex64 (float32 real part if real_output)."""
    if N is not None:
        bandwidth = N
    if m is not None:
        cutoff = m
    return NfftAdjointFunction.apply(x, pos, batch, int(bandwidth), int(cutoff), bool(real_output), batch_size)


def nfft_forward(x, pos, batch=None, cutoff=3, real_output=False, *, m=None, batch_size=None):
    """Forward NFFT  y[i, ...] ~= sum_k x[b_i, k + N/2, ...] exp(-2 pi i k . pos[i]).

    x: [batch_size, N, ..., N, *cols] float32|complex64.  Returns [n, *cols]."""
    if m is not None:
        cutoff = m
    return NfftForwardFunction.apply(x, pos, batch, int(cutoff), bool(real_output), batch_size)


def nfft_fastsum(x, coeffs, sources, targets=None, source_batch=None, target_batch=None, /, batch=None,
                 cutoff=3, *, m=None, batch_size=None, differentiable_points=False):
    """Fast multiplication with the trigonometric kernel matrix
    A[t, s] = sum_l coeffs[l + N/2] exp(2 pi i l . (sources[s] - targets[t])).

    Variants (reference nfft.py:97-103):
        nfft_fastsum(x, coeffs, sources)
        nfft_fastsum(x, coeffs, sources, targets)
        nfft_fastsum(x, coeffs, sources, batch=batch)
        nfft_fastsum(x, coeffs, sources, targets, batch=batch)
        nfft_fastsum(x, coeffs, sources, targets, source_batch, target_batch)
    Real x gives the real part (reference core_cuda.cu:814-818).

    `differentiable_points=True` (new; the reference asserts, nfft.py:66-69) evaluates the same
    product as forward(coeffs * adjoint(x, sources), targets), which autograd can differentiate
    w.r.t. x, coeffs, sources and targets."""
    if differentiable_points:
        if m is not None:
            cutoff = m
        if targets is None:
            targets, target_batch = sources, source_batch
        if batch is not None:
            source_batch = target_batch = batch
        N = coeffs.shape[0]
        d = sources.shape[1]
        spec = nfft_adjoint(x, sources, source_batch, N, int(cutoff), batch_size=batch_size)
        cshape = (1,) + tuple(coeffs.shape) + (1,) * (spec.dim() - 1 - d)
        out = nfft_forward(spec * coeffs.reshape(cshape), targets, target_batch, int(cutoff),
                           real_output=not x.is_complex(), batch_size=spec.shape[0])
        return out
    if targets is None:
        targets = sources
        target_batch = source_batch
    if batch is not None:
        source_batch = batch
        target_batch = batch
    if m is not None:
        cutoff = m
    return NfftFastsumFunction.apply(x, coeffs, sources, targets, source_batch, target_batch, int(cutoff),
                                     batch_size)


# --------------------------------------------------------------------------------------
# optional: the reference's TorchScript operator surface
# --------------------------------------------------------------------------------------
_OPS_REGISTERED = {}


def register_torch_ops(namespace: str = "torch_nfft"):
    """Registers `torch.ops.<namespace>.nfft_adjoint / nfft_forward / nfft_fastsum` with the reference's
    schemas and argument order (reference csrc/core.cpp:43-121,176-179: `(pos, x, batch, N, m,
    real_output)`), backed by this engine, for code that calls the raw operators.  The reference
    registers the same names, so never do this in a process that also loads the reference's core.so.
    Returns the torch.ops namespace."""
    if namespace in _OPS_REGISTERED:
        return getattr(torch.ops, namespace)
    lib = torch.library.Library(namespace, "DEF")
    lib.define("nfft_adjoint(Tensor pos, Tensor x, Tensor? batch, int N, int m, int real_output) -> Tensor")
    lib.define("nfft_forward(Tensor pos, Tensor x, Tensor? batch, int m, int real_output) -> Tensor")
    lib.define("nfft_fastsum(Tensor sources, Tensor targets, Tensor x, Tensor coeffs, Tensor? source_batch, "
               "Tensor? target_batch, int m) -> Tensor")
    lib.impl("nfft_adjoint", lambda pos, x, batch, N, m, real_output: _op_adjoint(pos, x, batch, N, m, bool(real_output)),
             "CUDA")
    lib.impl("nfft_forward", lambda pos, x, batch, m, real_output: _op_forward(pos, x, batch, m, bool(real_output)),
             "CUDA")
    lib.impl("nfft_fastsum", lambda sources, targets, x, coeffs, sb, tb, m: _op_fastsum(sources, targets, x, coeffs, sb, tb, m),
             "CUDA")
    _OPS_REGISTERED[namespace] = lib
    return getattr(torch.ops, namespace)
