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
  * `batch_ptr=`: batch_size + 1 int64 offsets of the point sets instead of one int64 per point;
  * `plan=NfftPlan(pos, batch)`: the binning of a point set, made once and reused by every transform of
    it (the reference recomputes its per-point scratch in every call, core_cuda.cu:188-211, 461-484);
  * gradients w.r.t. `pos` for nfft_forward / nfft_adjoint (the reference returns None,
    nfft.py:28,54): ONE forward transform of the d ramped spectra (-2 pi i k_a) xhat as extra channels.
"""
from __future__ import annotations

import torch

from . import _lib

_workspaces = {}
_ws_small_calls = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Scratch tensor per (device, stream) from the torch caching allocator: sort scratch, oversampled grid,
    half spectrum and the cuFFT work area of one call.  It grows on demand and is given back when 32
    consecutive calls on the stream needed less than a quarter of it (one c4-sized call would otherwise
    pin 1 GB per stream for the life of the process)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is not None and ws.numel() >= nbytes:
        if nbytes * 4 < ws.numel() and ws.numel() > (64 << 20) and not torch.cuda.is_current_stream_capturing():
            _ws_small_calls[key] = _ws_small_calls.get(key, 0) + 1
            if _ws_small_calls[key] < 32:
                return ws
        else:
            _ws_small_calls[key] = 0
            return ws
    _ws_small_calls[key] = 0
    _workspaces.pop(key, None)
    ws = None
    ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=device)
    _workspaces[key] = ws
    return ws


def clear_caches():
    """Drop cached workspaces and cuFFT plans.  While CUDA graphs captured by `GraphedTransforms` are alive
    the cuFFT handles stay (their kernels reference the handles' twiddle tables); a warning says so."""
    _workspaces.clear()
    _ws_small_calls.clear()
    if _lib.lib().nfftb200_plan_cache_clear() != 0:
        import warnings
        warnings.warn("torch_nfft_b200.clear_caches: " + _lib.lib().nfftb200_last_error().decode(), RuntimeWarning,
                      stacklevel=2)


def release_stream_workspace(device_index: int, cuda_stream: int):
    """Drop the scratch tensor kept for one (device, stream): called when the owner of a private stream
    goes away (`GraphedTransforms.close`)."""
    _workspaces.pop((device_index, cuda_stream), None)
    _ws_small_calls.pop((device_index, cuda_stream), None)


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _check(cond, msg):
    if not cond:
        raise RuntimeError("torch_nfft_b200: " + msg)


def _check_points(pos, batch, batch_size=None, batch_ptr=None):
    """check_point_input (core_cuda.cu:38-66).  `batch_ptr` (new): B + 1 ascending int64 offsets of the point
    sets instead of one int64 per point -- 8 (B + 1) bytes instead of 8 n, and no device->host read of
    batch[-1].  Returns (pos, batch-or-offsets tensor, n, d, B, offsets?)."""
    _check(isinstance(pos, torch.Tensor) and pos.is_cuda, "pos must be a CUDA tensor")
    _check(pos.dim() == 2, "pos must have shape [n, d]")
    _check(pos.dtype == torch.float32, "pos must be float32")
    n, d = pos.shape
    _check(1 <= d <= 3, "spatial dimension must be 1, 2 or 3")
    _check(n < 2 ** 32 - 1, "at most 2^32 - 2 points per call")
    if batch_ptr is not None:
        _check(batch is None, "pass either batch or batch_ptr")
        _check(batch_ptr.is_cuda and batch_ptr.device == pos.device, "batch_ptr must be a CUDA tensor on the device of pos")
        _check(batch_ptr.dim() == 1 and batch_ptr.dtype == torch.int64 and batch_ptr.numel() >= 2,
               "batch_ptr must be a 1-D int64 tensor of batch_size + 1 offsets")
        B = batch_ptr.numel() - 1
        _check(batch_size is None or int(batch_size) == B, "batch_size does not match batch_ptr")
        return pos.contiguous(), batch_ptr.contiguous(), int(n), int(d), int(B), True
    if batch is not None:
        _check(batch.is_cuda and batch.device == pos.device, "batch must be a CUDA tensor on the device of pos")
        _check(batch.dim() == 1 and batch.dtype == torch.int64, "batch must be a 1-D int64 tensor")
        _check(batch.numel() == n, "batch must have one entry per point")
        if batch_size is None:
            batch_size = int(batch[-1].item()) + 1 if n > 0 else 1  # core_cuda.cu:60
        batch = batch.contiguous()
    else:
        batch_size = 1
    return pos.contiguous(), batch, int(n), int(d), int(batch_size), False


def _check_cutoff(m, N):
    _check(isinstance(m, int) and 1 <= m <= 8, "cutoff must be an integer in [1, 8]")
    _check(N >= 2 and N % 2 == 0, "bandwidth must be even and >= 2")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# persistent point plan (the reference recomputes its per-point scratch in
# every call, core_cuda.cu:188-211, 461-484)
# --------------------------------------------------------------------------------------
_TILING_FIELDS = ("dim", "M", "Tx", "Ty", "Tz", "ntx", "nty", "ntz", "pmax", "fine_bits", "scx", "scy", "scz", "mixed",
                  "dense_tile_pts")


class NfftPlan:
    """The binning of one point set, made once and reused by every transform of these points.

        plan = NfftPlan(pos, batch, batch_size=B)            # or batch_ptr=offsets
        y = nfft_adjoint(x, plan=plan, N=128, m=4)
        f = nfft_forward(y, plan=plan, m=4, real_output=True)

    The plan keeps `pos` / `batch` (contiguous) and, per tiling the engine uses for the requested
    transforms, the stable permutation of the points by grid tile with its bin offsets and work items
    (4 bytes per point, made lazily on the stream of the first transform that needs it).  The caller owns
    the validity: positions must not change while the plan is used.  Transforms never index out of bounds
    with a stale plan -- they drop the points they find outside their tile and count them;
    `dropped_points()` reads that counter (it synchronises).

    `GramMatrix`, `AdjacencyMatrix` and the backward passes of the three autograd functions use plans, so
    `A @ x` in an iterative solver bins the points once.

    `clustered=True` is a hint for large 3D point sets (cutoff 3 or 4) that are far from uniform: the binning
    then samples the points on the device and, if it finds heavy grid tiles, orders their points more finely
    and marks them for the sweep made for dense tiles (2 x 2 x 2-cell supercells).  Results are the same; a
    Gaussian-clustered c4 runs 5 % faster with it, a uniform one 1.5 % slower (`flags()["clustered"]` tells
    what the device found)."""

    def __init__(self, pos, batch=None, *, batch_size=None, batch_ptr=None, clustered=False):
        self.pos, self.batch, self.n, self.d, self.batch_size, self.offsets = _check_points(pos, batch, batch_size, batch_ptr)
        self.device = self.pos.device
        self.clustered = bool(clustered)
        self._sorts = {}

    @property
    def op_flags(self):
        """Flag bits every C-ABI call that uses this plan must carry (they are part of the tiling)."""
        return _lib.PLANNED | (_lib.CLUSTERED if self.clustered else 0)

    def _sorted(self, N, m, C, cplx_flag, n_geom=0):
        """(plan buffer, flags to add) for the tiling of a transform with these parameters."""
        L = _lib.lib()
        hint = _lib.CLUSTERED if self.clustered else 0
        geo = _lib.geometry(self.d, N, m, self.batch_size, C, cplx_flag | hint, max(self.n, n_geom))
        key = tuple(geo[f] for f in _TILING_FIELDS)
        entry = self._sorts.get(key)
        if entry is None:
            flags = cplx_flag | hint | (_lib.BATCH_OFFSETS if self.offsets else 0)
            args = (self.n, n_geom, self.d, N, m, self.batch_size, C, flags)
            nbytes = L.nfftb200_plan_bytes(*args)
            _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
            with torch.cuda.device(self.device):
                buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                ws = _workspace(L.nfftb200_workspace_bytes(_lib.OP_PLAN, self.n, n_geom, self.d, N, m, self.batch_size,
                                                           C, flags), self.device)
                _lib.check(L.nfftb200_plan_points(self.pos.data_ptr(), _ptr(self.batch), buf.data_ptr(), buf.numel(),
                                                  self.n, n_geom, self.d, N, m, self.batch_size, C, flags, ws.data_ptr(),
                                                  ws.numel(), _stream_ptr(self.device)), "plan_points")
            entry = (buf, args)
            self._sorts[key] = entry
        return entry[0]

    def matches(self, pos, batch=None):
        """True if `pos` / `batch` are the tensors this plan was made from (same memory and shape)."""
        if pos.shape != self.pos.shape or pos.data_ptr() != self.pos.data_ptr():
            return False
        return batch is None or self.batch is None or self.offsets or batch.data_ptr() == self.batch.data_ptr()

    @property
    def nbytes(self):
        """Device bytes held by the binnings made so far."""
        return sum(buf.numel() for buf, _ in self._sorts.values())

    @property
    def sorts(self):
        """Number of distinct tilings the points have been binned for."""
        return len(self._sorts)

    def flags(self):
        """Device-side counters of the binnings made so far, summed: `dropped` = points that transforms using this
        plan found outside their tile (evidence that the positions changed after the plan was made), `tma_timeouts`
        = TMA tile loads that did not complete (must be 0), `clustered` = binnings that found the point set clustered
        and marked its heavy tiles for the 2 x 2 x 2 sweep (large 3D sets).  Reads device memory: synchronises the
        current stream."""
        import ctypes
        L = _lib.lib()
        dropped = timeouts = clustered = 0
        out = (ctypes.c_uint32 * 8)()
        with torch.cuda.device(self.device):
            for buf, args in self._sorts.values():
                _lib.check(L.nfftb200_plan_flags(buf.data_ptr(), *args, ctypes.cast(out, ctypes.c_void_p),
                                                 _stream_ptr(self.device)), "plan_flags")
                dropped += int(out[0])
                timeouts += int(out[1])
                clustered += int(out[2])
        return {"dropped": dropped, "tma_timeouts": timeouts, "clustered": clustered}

    def dropped_points(self):
        """Points that transforms using this plan found outside their tile (see `flags`).  Synchronises."""
        return self.flags()["dropped"]


def _resolve_points(pos, batch, batch_size, batch_ptr, plan):
    """Common front end: returns (plan-or-None, pos, batch tensor, n, d, B, offsets?)."""
    if isinstance(pos, NfftPlan):
        _check(plan is None or plan is pos, "two different plans passed")
        plan, pos = pos, None
    if plan is not None:
        _check(isinstance(plan, NfftPlan), "plan must be an NfftPlan")
        _check(pos is None or plan.matches(pos, batch), "plan was made for other pos / batch tensors")
        _check(batch_size is None or int(batch_size) == plan.batch_size, "batch_size does not match the plan")
        return plan, plan.pos, plan.batch, plan.n, plan.d, plan.batch_size, plan.offsets
    return (None,) + _check_points(pos, batch, batch_size, batch_ptr)


# --------------------------------------------------------------------------------------
# raw operators (same argument order as torch.ops.torch_nfft.*, reference core.cpp:43-121)
# --------------------------------------------------------------------------------------
def _op_adjoint(pos, x, batch, N, m, real_output, batch_size=None, plan=None, batch_ptr=None):
    plan, pos, batch, n, d, B, offsets = _resolve_points(pos, batch, batch_size, batch_ptr, plan)
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
    cplx = _lib.X_COMPLEX if x.is_complex() else 0
    flags = cplx | (_lib.Y_REAL if real_output else 0) | (_lib.BATCH_OFFSETS if offsets else 0)
    y = torch.empty((B,) + (N,) * d + cols, dtype=torch.float32 if real_output else torch.complex64, device=pos.device)
    if C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(pos.device):
        pbuf = None
        if plan is not None:
            pbuf = plan._sorted(N, m, C, cplx)
            flags |= plan.op_flags
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, n, 0, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, pos.device)
        _lib.check(L.nfftb200_adjoint_planned(_ptr(pos), _ptr(x), _ptr(batch), _ptr(pbuf), 0 if pbuf is None else pbuf.numel(),
                                              _ptr(y), n, d, N, m, B, C, flags, ws.data_ptr(), ws.numel(),
                                              _stream_ptr(pos.device)), "nfft_adjoint")
    return y


def _op_forward(pos, xhat, batch, m, real_output, batch_size=None, plan=None, batch_ptr=None):
    plan, pos, batch, n, d, B, offsets = _resolve_points(pos, batch, batch_size, batch_ptr, plan)
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
    flags = ((_lib.X_COMPLEX if xhat.is_complex() else 0) | (_lib.Y_REAL if real_output else 0)
             | (_lib.BATCH_OFFSETS if offsets else 0))
    y = torch.empty((n,) + cols, dtype=torch.float32 if real_output else torch.complex64, device=pos.device)
    if n == 0 or C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(pos.device):
        pbuf = None
        if plan is not None:
            # the gather grid is complex unless real_output: same tiling rule as a complex adjoint
            pbuf = plan._sorted(N, m, C, 0 if real_output else _lib.X_COMPLEX)
            flags |= plan.op_flags
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_FORWARD, 0, n, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, pos.device)
        _lib.check(L.nfftb200_forward_planned(_ptr(pos), _ptr(xhat), _ptr(batch), _ptr(pbuf),
                                              0 if pbuf is None else pbuf.numel(), _ptr(y), n, d, N, m, B, C, flags,
                                              ws.data_ptr(), ws.numel(), _stream_ptr(pos.device)), "nfft_forward")
    return y


def _op_fastsum(sources, targets, x, coeffs, source_batch, target_batch, m, batch_size=None, source_plan=None,
                target_plan=None):
    symmetric = targets is sources or (source_plan is not None and target_plan is source_plan)  # core_cuda.cu:552
    source_plan, sources_c, source_batch, n_src, d, B, s_off = _resolve_points(sources, source_batch, batch_size, None,
                                                                               source_plan)
    if symmetric:
        target_plan, targets_c, target_batch, n_tgt, t_off = source_plan, sources_c, source_batch, n_src, s_off
    else:
        target_plan, targets_c, target_batch, n_tgt, d_t, B_t, t_off = _resolve_points(targets, target_batch, batch_size,
                                                                                        None, target_plan)
        _check(d_t == d, "sources and targets must have the same dimension")
        _check(B_t == B, "sources and targets must have the same batch size")
        _check(t_off == s_off, "sources and targets must describe their point sets the same way (batch or batch_ptr)")
        _check((source_plan is None) == (target_plan is None), "pass plans for both sources and targets, or for neither")
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
    cplx = _lib.X_COMPLEX if x.is_complex() else 0
    flags = (cplx | (_lib.COEFFS_COMPLEX if coeffs.is_complex() else 0) | (_lib.SYMMETRIC if symmetric else 0)
             | (_lib.BATCH_OFFSETS if s_off else 0))
    y = torch.empty((n_tgt,) + cols, dtype=x.dtype, device=x.device)
    if n_tgt == 0 or C == 0:
        return y
    L = _lib.lib()
    with torch.cuda.device(x.device):
        sbuf = tbuf = None
        if source_plan is not None:
            n_geom = max(n_src, n_tgt)
            sbuf = source_plan._sorted(N, m, C, cplx, n_geom)
            tbuf = sbuf if symmetric else target_plan._sorted(N, m, C, cplx, n_geom)
            _check(symmetric or source_plan.clustered == target_plan.clustered,
                   "source_plan and target_plan must agree on the clustered hint")
            flags |= source_plan.op_flags
        nbytes = L.nfftb200_workspace_bytes(_lib.OP_FASTSUM, n_src, n_tgt, d, N, m, B, C, flags)
        _check(nbytes > 0, "invalid arguments: " + L.nfftb200_last_error().decode())
        ws = _workspace(nbytes, x.device)
        _lib.check(L.nfftb200_fastsum_planned(_ptr(sources_c), _ptr(targets_c), _ptr(x), _ptr(coeffs), _ptr(source_batch),
                                              _ptr(target_batch), _ptr(sbuf), 0 if sbuf is None else sbuf.numel(),
                                              _ptr(tbuf), 0 if tbuf is None else tbuf.numel(), _ptr(y), n_src, n_tgt,
                                              d, N, m, B, C, flags, ws.data_ptr(), ws.numel(), _stream_ptr(x.device)),
                   "nfft_fastsum")
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


def _ramped_spectra(xh, N, d):
    """[B, N..N, d * C]: the spectra (-2 pi i k_a) xhat of all d directions as extra channels, so that ONE
    forward transform (one pack, one FFT batch, one gather sweep over the points) yields all d derivative
    components instead of d separate transforms."""
    return torch.cat([xh * _freq_ramp(N, d, a, xh.device) for a in range(d)], dim=-1).contiguous()


def _forward_pos_grad(plan, xhat, m, dy):
    """d/dpos of sum(Re(conj(dy) * forward(xhat))): forward NFFT of (-2 pi i k_a) xhat."""
    n, d = plan.n, plan.d
    B, N = xhat.shape[0], xhat.shape[1]
    xh = xhat.reshape(B, *(N,) * d, -1).to(torch.complex64)
    C = xh.shape[-1]
    fa = _op_forward(None, _ramped_spectra(xh, N, d), None, m, False, plan=plan).reshape(n, d, C)
    dyr = dy.reshape(n, 1, C)
    if dy.is_complex():
        return (fa * dyr.conj()).real.sum(-1)
    return (fa.real * dyr).sum(-1)


def _adjoint_pos_grad(plan, x, N, m, dy):
    """d/dpos of sum(Re(conj(dy) * adjoint(x))): y_k = sum_i x_i e^{+2 pi i k p_i}."""
    n, d = plan.n, plan.d
    B = dy.shape[0]
    dyc = dy.reshape(B, *(N,) * d, -1).to(torch.complex64)
    C = dyc.shape[-1]
    # fa_i = sum_k dy_k (-2 pi i k_a) e^{-2 pi i k p_i};  conj(fa_i) = sum_k conj(dy_k) (2 pi i k_a) e^{+2 pi i k p_i}
    fa = _op_forward(None, _ramped_spectra(dyc, N, d), None, m, False, plan=plan).reshape(n, d, C)
    return (fa.conj() * x.reshape(n, 1, C)).real.sum(-1)


# --------------------------------------------------------------------------------------
# autograd wrappers (reference torch_nfft/nfft.py:11-88)
# --------------------------------------------------------------------------------------
def _plan_for_backward(plan, pos, batch, batch_size, batch_ptr, needs_grad):
    """The backward pass transforms the same points again: bin them once.  (autograd's version check of the
    saved tensors guards the positions against in-place changes between forward and backward.)"""
    if plan is None and needs_grad:
        plan = NfftPlan(pos, batch, batch_size=batch_size, batch_ptr=batch_ptr)
    return plan


class NfftAdjointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, batch, bandwidth, cutoff, real_output, batch_size, plan, batch_ptr):
        needs_grad = any(ctx.needs_input_grad)
        if isinstance(pos, NfftPlan):
            plan, pos = pos, pos.pos
        plan = _plan_for_backward(plan, pos, batch, batch_size, batch_ptr, needs_grad)
        y = _op_adjoint(pos, x, batch, bandwidth, cutoff, real_output, batch_size, plan, batch_ptr)
        ctx.save_for_backward(pos, x if pos.requires_grad else None)
        ctx.plan = plan
        ctx.cutoff = cutoff
        ctx.bandwidth = bandwidth
        ctx.real_input = not x.is_complex()
        return y

    @staticmethod
    def backward(ctx, dy):
        pos, x = ctx.saved_tensors
        dx = dpos = None
        if ctx.needs_input_grad[0]:
            # reference nfft.py:26: forward NFFT of dy, real output iff x was real
            dx = _op_forward(None, dy, None, ctx.cutoff, ctx.real_input, plan=ctx.plan)
        if ctx.needs_input_grad[1]:
            dpos = _adjoint_pos_grad(ctx.plan, x, ctx.bandwidth, ctx.cutoff, dy)
        return dx, dpos, None, None, None, None, None, None, None


class NfftForwardFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, batch, cutoff, real_output, batch_size, plan, batch_ptr):
        needs_grad = any(ctx.needs_input_grad)
        if isinstance(pos, NfftPlan):
            plan, pos = pos, pos.pos
        plan = _plan_for_backward(plan, pos, batch, batch_size, batch_ptr, needs_grad)
        y = _op_forward(pos, x, batch, cutoff, real_output, batch_size, plan, batch_ptr)
        ctx.save_for_backward(pos, x if pos.requires_grad else None)
        ctx.plan = plan
        ctx.cutoff = cutoff
        ctx.bandwidth = x.size(1)
        ctx.real_input = not x.is_complex()
        return y

    @staticmethod
    def backward(ctx, dy):
        pos, x = ctx.saved_tensors
        dx = dpos = None
        if ctx.needs_input_grad[0]:
            # reference nfft.py:52: adjoint NFFT of dy, real output iff x was real
            dx = _op_adjoint(None, dy, None, ctx.bandwidth, ctx.cutoff, ctx.real_input, plan=ctx.plan)
        if ctx.needs_input_grad[1]:
            dpos = _forward_pos_grad(ctx.plan, x, ctx.cutoff, dy)
        return dx, dpos, None, None, None, None, None, None


class NfftFastsumFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coeffs, sources, targets, source_batch, target_batch, cutoff, batch_size, source_plan,
                target_plan):
        # reference nfft.py:66-73
        assert not coeffs.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. coefficients is not possible"
        assert not sources.requires_grad and not targets.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. sources and targets is not possible"
        assert source_batch is None or not source_batch.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. batches is not possible"
        assert target_batch is None or not target_batch.requires_grad, \
            "NfftFastsum: Gradient computation w.r.t. batches is not possible"
        symmetric = targets is sources
        if source_plan is None and ctx.needs_input_grad[0]:
            # the backward pass is the same product with sources and targets swapped: bin both sets once
            source_plan = NfftPlan(sources, source_batch, batch_size=batch_size)
            target_plan = source_plan if symmetric else NfftPlan(targets, target_batch, batch_size=batch_size)
        y = _op_fastsum(sources, targets, x, coeffs, source_batch, target_batch, cutoff, batch_size, source_plan,
                        target_plan)
        ctx.save_for_backward(sources, targets, coeffs, source_batch, target_batch)
        ctx.symmetric = symmetric
        ctx.cutoff = cutoff
        ctx.batch_size = batch_size
        ctx.plans = (source_plan, target_plan)
        return y

    @staticmethod
    def backward(ctx, dy):
        sources, targets, coeffs, source_batch, target_batch = ctx.saved_tensors
        if ctx.symmetric:
            targets = sources
        source_plan, target_plan = ctx.plans
        # reference nfft.py:86: fastsum with sources and targets swapped
        dx = _op_fastsum(targets, sources, dy.contiguous(), coeffs, target_batch, source_batch, ctx.cutoff,
                         ctx.batch_size, target_plan, source_plan)
        return dx, None, None, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------
# functional API (reference torch_nfft/nfft.py:31,57,91)
# --------------------------------------------------------------------------------------
def nfft_adjoint(x, pos=None, batch=None, bandwidth=16, cutoff=3, real_output=False, *, N=None, m=None,
                 batch_size=None, batch_ptr=None, plan=None):
    """Adjoint NFFT  y[b, k + N/2, ...] ~= sum_{i in b} x[i, ...] exp(+2 pi i k . pos[i]).

    x: [n, *cols] float32|complex64, pos: [n, d] float32 in [-1/2, 1/2), batch: [n] int64 sorted.
    Returns [batch_size, N, ..., N, *cols] complex64 (float32 real part if real_output).
    New, optional: `plan=NfftPlan(pos, batch)` (or the plan passed as `pos`) reuses the binning of the
    points; `batch_ptr` = batch_size + 1 int64 offsets instead of the per-point `batch` vector."""
    if N is not None:
        bandwidth = N
    if m is not None:
        cutoff = m
    if pos is None:
        _check(plan is not None, "nfft_adjoint needs pos or plan")
        pos = plan
    return NfftAdjointFunction.apply(x, pos, batch, int(bandwidth), int(cutoff), bool(real_output), batch_size, plan,
                                     batch_ptr)


def nfft_forward(x, pos=None, batch=None, cutoff=3, real_output=False, *, m=None, batch_size=None, batch_ptr=None,
                 plan=None):
    """Forward NFFT  y[i, ...] ~= sum_k x[b_i, k + N/2, ...] exp(-2 pi i k . pos[i]).

    x: [batch_size, N, ..., N, *cols] float32|complex64.  Returns [n, *cols].  `plan`, `batch_ptr`: see
    nfft_adjoint."""
    if m is not None:
        cutoff = m
    if pos is None:
        _check(plan is not None, "nfft_forward needs pos or plan")
        pos = plan
    return NfftForwardFunction.apply(x, pos, batch, int(cutoff), bool(real_output), batch_size, plan, batch_ptr)


def nfft_fastsum(x, coeffs, sources, targets=None, source_batch=None, target_batch=None, /, batch=None,
                 cutoff=3, *, m=None, batch_size=None, differentiable_points=False, source_plan=None,
                 target_plan=None):
    """Fast multiplication with the trigonometric kernel matrix
    A[t, s] = sum_l coeffs[l + N/2] exp(2 pi i l . (sources[s] - targets[t])).

    Variants (reference nfft.py:97-103):
        nfft_fastsum(x, coeffs, sources)
        nfft_fastsum(x, coeffs, sources, targets)
        nfft_fastsum(x, coeffs, sources, batch=batch)
        nfft_fastsum(x, coeffs, sources, targets, batch=batch)
        nfft_fastsum(x, coeffs, sources, targets, source_batch, target_batch)
    Real x gives the real part (reference core_cuda.cu:814-818).

    `source_plan` / `target_plan` (new): NfftPlans of sources / targets (the same object for the symmetric
    product) so that repeated products bin the points once.

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
        if target_plan is None:
            target_plan = source_plan
    if batch is not None:
        source_batch = batch
        target_batch = batch
    if m is not None:
        cutoff = m
    return NfftFastsumFunction.apply(x, coeffs, sources, targets, source_batch, target_batch, int(cutoff),
                                     batch_size, source_plan, target_plan)


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
