"""Multi-GPU execution of the NFFT hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink 5 / NVSwitch on B200; gloo on CPU for the host-logic tests).

The reference has no multi-device path at all.  Two shardings:

* batch sharding  -- point sets (batch entries) are independent transforms
  (reference spatial_window_operations.cu:146: grid index (b*C + c)), so each rank transforms a
  contiguous range of batch entries.  No communication.
* point sharding  -- a single huge point set is split by points.  Spreading is linear in the
  points, so each rank spreads its slice into a *partial* oversampled grid and the partial grids
  are summed over the ranks.  Adjoint with B a multiple of the world size: ONE NCCL reduce-scatter
  leaves whole summed grids on each rank, which runs FFT + unpack for its batch entries only.
  Otherwise (a single grid, fastsum): all-reduce (in-switch with NVLS), the small FFT / spectral
  stage runs redundantly on every rank, and each rank gathers at its own target points.  Outputs
  indexed by points stay sharded.

The compute stages are the split entry points of the C ABI (nfftb200_spread / _adjoint_finish /
_forward_begin / _gather / _fastsum_middle).  They are injected through an `engine` object so the
sharding logic can be tested on CPU (gloo, world_size 2) with a numpy stand-in engine.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from . import nfft as _nfft


# --------------------------------------------------------------------------------------
# shard planning (pure host logic)
# --------------------------------------------------------------------------------------
def split_range(total: int, world: int, rank: int):
    """Contiguous, balanced split of range(total): the first total % world parts get one more."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batches(batch: torch.Tensor, batch_size: int, world: int, rank: int):
    """Batch sharding plan for a sorted batch vector.

    Returns (b_lo, b_hi, p_lo, p_hi): this rank owns batch entries [b_lo, b_hi) = points
    [p_lo, p_hi) (contiguous because `batch` is sorted ascending, reference README.md:44-46)."""
    b_lo, b_hi = split_range(batch_size, world, rank)
    bounds = torch.searchsorted(batch, torch.tensor([b_lo, b_hi], dtype=batch.dtype, device=batch.device))
    return b_lo, b_hi, int(bounds[0]), int(bounds[1])


def shard_points(n: int, world: int, rank: int):
    """Point sharding plan: this rank owns points [p_lo, p_hi)."""
    return split_range(n, world, rank)


# --------------------------------------------------------------------------------------
# engines
# --------------------------------------------------------------------------------------
class CudaEngine:
    """Stages of the transform on the current CUDA device through the C ABI.  `plan` (an `NfftPlan` of exactly the
    `pos` / `batch` passed along) lets spread and gather of the same points share one binning."""

    supports_plans = True

    @staticmethod
    def _geom(pos, batch, batch_size):
        pos, batch, n, d, B, _ = _nfft._check_points(pos, batch, batch_size)
        return pos, batch, n, d, B

    def spread(self, x, pos, batch, B, N, m, plan=None):
        """x [n, *cols] -> partial grid [B*C, (2N)^d] (float32, or complex64 for complex x)."""
        pos, batch, n, d, B = self._geom(pos, batch, B)
        x = x.contiguous()
        C = max(1, x.numel() // max(n, 1)) if n > 0 else int(torch.tensor(x.shape[1:]).prod()) if x.dim() > 1 else 1
        flags = _lib.X_COMPLEX if x.is_complex() else 0
        grid = torch.empty((B * C,) + (2 * N,) * d, dtype=x.dtype, device=pos.device)
        L = _lib.lib()
        with torch.cuda.device(pos.device):
            pbuf = None
            if plan is not None and n > 0:
                pbuf = plan._sorted(N, m, C, flags)
                flags |= plan.op_flags
            ws = _nfft._workspace(L.nfftb200_workspace_bytes(_lib.OP_SPREAD, n, 0, d, N, m, B, C, flags), pos.device)
            _lib.check(L.nfftb200_spread(pos.data_ptr(), x.data_ptr(), _nfft._ptr(batch), _nfft._ptr(pbuf),
                                         0 if pbuf is None else pbuf.numel(), grid.data_ptr(), n, d, N, m, B, C, flags,
                                         ws.data_ptr(), ws.numel(), _nfft._stream_ptr(pos.device)), "spread")
        return grid

    def adjoint_finish(self, grid, d, B, cols, N, m, real_output):
        C = grid.shape[0] // B
        flags = (_lib.X_COMPLEX if grid.is_complex() else 0) | (_lib.Y_REAL if real_output else 0)
        y = torch.empty((B,) + (N,) * d + tuple(cols), dtype=torch.float32 if real_output else torch.complex64,
                        device=grid.device)
        L = _lib.lib()
        with torch.cuda.device(grid.device):
            ws = _nfft._workspace(L.nfftb200_workspace_bytes(_lib.OP_SPECTRAL, 0, 0, d, N, m, B, C, flags), grid.device)
            _lib.check(L.nfftb200_adjoint_finish(grid.data_ptr(), y.data_ptr(), d, N, m, B, C, flags, ws.data_ptr(),
                                                 ws.numel(), _nfft._stream_ptr(grid.device)), "adjoint_finish")
        return y

    def forward_begin(self, xhat, d, m, real_output):
        xhat = xhat.contiguous()
        B, N = xhat.shape[0], xhat.shape[1]
        C = xhat.numel() // (B * N ** d)
        flags = (_lib.X_COMPLEX if xhat.is_complex() else 0) | (_lib.Y_REAL if real_output else 0)
        grid = torch.empty((B * C,) + (2 * N,) * d, dtype=torch.float32 if real_output else torch.complex64,
                           device=xhat.device)
        L = _lib.lib()
        with torch.cuda.device(xhat.device):
            ws = _nfft._workspace(L.nfftb200_workspace_bytes(_lib.OP_SPECTRAL, 0, 0, d, N, m, B, C, flags), xhat.device)
            _lib.check(L.nfftb200_forward_begin(xhat.data_ptr(), grid.data_ptr(), d, N, m, B, C, flags, ws.data_ptr(),
                                                ws.numel(), _nfft._stream_ptr(xhat.device)), "forward_begin")
        return grid

    def gather(self, grid, pos, batch, B, cols, N, m, plan=None):
        pos, batch, n, d, B = self._geom(pos, batch, B)
        C = grid.shape[0] // B
        flags = _lib.X_COMPLEX if grid.is_complex() else 0
        y = torch.empty((n,) + tuple(cols), dtype=grid.dtype, device=grid.device)
        if n == 0:
            return y
        L = _lib.lib()
        with torch.cuda.device(grid.device):
            pbuf = None
            if plan is not None:
                pbuf = plan._sorted(N, m, C, flags)
                flags |= plan.op_flags
            ws = _nfft._workspace(L.nfftb200_workspace_bytes(_lib.OP_GATHER, 0, n, d, N, m, B, C, flags), grid.device)
            _lib.check(L.nfftb200_gather(pos.data_ptr(), _nfft._ptr(batch), _nfft._ptr(pbuf),
                                         0 if pbuf is None else pbuf.numel(), grid.data_ptr(), y.data_ptr(), n, d, N, m, B,
                                         C, flags, ws.data_ptr(), ws.numel(), _nfft._stream_ptr(grid.device)), "gather")
        return y

    def fastsum_middle(self, grid, coeffs, d, B, N, m):
        C = grid.shape[0] // B
        coeffs = coeffs.contiguous()
        flags = (_lib.X_COMPLEX if grid.is_complex() else 0) | (_lib.COEFFS_COMPLEX if coeffs.is_complex() else 0)
        L = _lib.lib()
        with torch.cuda.device(grid.device):
            ws = _nfft._workspace(L.nfftb200_workspace_bytes(_lib.OP_SPECTRAL, 0, 0, d, N, m, B, C, flags), grid.device)
            _lib.check(L.nfftb200_fastsum_middle(grid.data_ptr(), coeffs.data_ptr(), d, N, m, B, C, flags,
                                                 ws.data_ptr(), ws.numel(), _nfft._stream_ptr(grid.device)),
                       "fastsum_middle")
        return grid


_default_engine = CudaEngine()


def _sum_over_ranks(grid, group):
    """Sum the partial grids of all ranks; every rank ends up with the full grid.
    NCCL implements this as reduce-scatter + all-gather over NVLink (in-switch with NVLS)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if grid.is_complex():
            dist.all_reduce(torch.view_as_real(grid), op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    return grid


def _reduce_scatter_rows(grid, group):
    """Sum the partial grids over the ranks and leave rank r with rows [r*R, (r+1)*R) of the sum,
    R = rows / world (whole grids per rank: the FFT stage then needs no further exchange).
    NCCL: one reduce-scatter over NVLink; gloo (CPU tests) has none, so all-reduce + slice."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows = grid.shape[0] // world
    flat = torch.view_as_real(grid) if grid.is_complex() else grid
    if dist.get_backend(group) == "nccl":
        out = torch.empty((rows,) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
        dist.reduce_scatter_tensor(out, flat.contiguous(), op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        out = flat[rank * rows:(rank + 1) * rows].contiguous()
    return torch.view_as_complex(out) if grid.is_complex() else out


def _all_gather_batches(y_local, group):
    world = dist.get_world_size(group)
    flat = torch.view_as_real(y_local) if y_local.is_complex() else y_local
    out = torch.empty((world * flat.shape[0],) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat.contiguous(), group=group)
    return torch.view_as_complex(out) if y_local.is_complex() else out


# --------------------------------------------------------------------------------------
# point-sharded transforms: every rank passes ITS slice of the points
# --------------------------------------------------------------------------------------
def _local_plan(eng, pos, batch, B, plan):
    """The binning of this rank's point slice: the caller's, or (for engines that take plans) a fresh one."""
    if not getattr(eng, "supports_plans", False):
        return None
    return plan if plan is not None else _nfft.NfftPlan(pos, batch, batch_size=B)


def nfft_adjoint_point_sharded(x, pos, batch=None, bandwidth=16, cutoff=3, real_output=False, *, batch_size=None,
                               group=None, engine=None, scatter_output=False, plan=None):
    """Adjoint NFFT of a point set that is split across ranks.

    Every rank spreads its points into a partial oversampled grid.  If the number of batch entries is
    a multiple of the world size the partial grids are summed with ONE reduce-scatter into whole
    grids per rank, each rank runs FFT + unpack for its B / world entries only,
    and the (2^d times smaller) spectra are all-gathered unless `scatter_output`; otherwise the grids
    are all-reduced and the spectral stage runs redundantly.

    Returns the full spectrum [B, N..N, *cols] on every rank, or with scatter_output=True
    `(y_local, (b_lo, b_hi))`, the spectra of the batch entries this rank owns.
    """
    eng = engine or _default_engine
    d = pos.shape[1]
    B = int(batch_size) if batch_size is not None else (1 if batch is None else None)
    if B is None:
        raise RuntimeError("point-sharded transforms with a batch vector need batch_size= (the global value)")
    cols = tuple(x.shape[1:])
    grid = eng.spread(x, pos, batch, B, bandwidth, cutoff, **({"plan": plan} if plan is not None else {}))
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1 and B % world == 0:
        rank = dist.get_rank(group)
        grid = _reduce_scatter_rows(grid, group)
        y_local = eng.adjoint_finish(grid, d, B // world, cols, bandwidth, cutoff, real_output)
        if scatter_output:
            return y_local, (rank * (B // world), (rank + 1) * (B // world))
        return _all_gather_batches(y_local, group)
    grid = _sum_over_ranks(grid, group)
    y = eng.adjoint_finish(grid, d, B, cols, bandwidth, cutoff, real_output)
    if scatter_output:
        lo, hi = split_range(B, world, dist.get_rank(group) if world > 1 else 0)
        return y[lo:hi], (lo, hi)
    return y


def nfft_forward_point_sharded(xhat, pos, batch=None, cutoff=3, real_output=False, *, batch_size=None, group=None,
                               engine=None, plan=None):
    """Forward NFFT at this rank's slice of the points; xhat is replicated.  Returns [n_local, *cols].
    `plan`: the NfftPlan of this rank's slice (e.g. the one the adjoint of the same points used)."""
    eng = engine or _default_engine
    d = pos.shape[1]
    B, N = xhat.shape[0], xhat.shape[1]
    grid = eng.forward_begin(xhat, d, cutoff, real_output)
    return eng.gather(grid, pos, batch, B, tuple(xhat.shape[1 + d:]), N, cutoff,
                      **({"plan": plan} if plan is not None else {}))


def nfft_fastsum_point_sharded(x, coeffs, sources, targets=None, source_batch=None, target_batch=None, *,
                               cutoff=3, batch_size=None, group=None, engine=None, source_plan=None, target_plan=None):
    """Fastsum with sources and targets split across ranks (each rank passes its slices of both).
    One all-reduce of the oversampled grid; the result rows stay sharded like `targets`.
    `source_plan` / `target_plan`: NfftPlans of this rank's slices, kept by callers that multiply repeatedly."""
    eng = engine or _default_engine
    if targets is None:
        targets, target_batch = sources, source_batch
    d = sources.shape[1]
    N = coeffs.shape[0]
    B = int(batch_size) if batch_size is not None else (1 if source_batch is None else None)
    if B is None:
        raise RuntimeError("point-sharded transforms with a batch vector need batch_size= (the global value)")
    # the symmetric product spreads from and gathers at the same points: one binning serves both stages
    symmetric = targets is sources and target_batch is source_batch
    src_plan = _local_plan(eng, sources, source_batch, B, source_plan) if (symmetric or source_plan is not None) else None
    tgt_plan = src_plan if symmetric else target_plan
    grid = eng.spread(x, sources, source_batch, B, N, cutoff, **({"plan": src_plan} if src_plan is not None else {}))
    grid = _sum_over_ranks(grid, group)
    grid = eng.fastsum_middle(grid, coeffs, d, B, N, cutoff)
    return eng.gather(grid, targets, target_batch, B, tuple(x.shape[1:]), N, cutoff,
                      **({"plan": tgt_plan} if tgt_plan is not None else {}))


# --------------------------------------------------------------------------------------
# batch-sharded transforms: every rank passes the FULL inputs, computes its batch entries
# --------------------------------------------------------------------------------------
def nfft_adjoint_batch_sharded(x, pos, batch, bandwidth=16, cutoff=3, real_output=False, *, batch_size, group=None,
                               gather_output=False, adjoint_fn=None):
    """Adjoint NFFT of the batch entries owned by this rank.  Returns (y_local, (b_lo, b_hi));
    with gather_output=True the full [B, ...] tensor is assembled on every rank (all_gather)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    b_lo, b_hi, p_lo, p_hi = shard_batches(batch, batch_size, world, rank)
    fn = adjoint_fn or _nfft.nfft_adjoint
    nb = b_hi - b_lo
    if nb > 0:
        y = fn(x[p_lo:p_hi], pos[p_lo:p_hi], batch[p_lo:p_hi] - b_lo, bandwidth, cutoff, real_output, batch_size=nb)
    else:
        d = pos.shape[1]
        y = torch.zeros((0,) + (bandwidth,) * d + tuple(x.shape[1:]),
                        dtype=torch.float32 if real_output else torch.complex64, device=x.device)
    if not gather_output or world == 1:
        return y, (b_lo, b_hi)
    # ranks may own different numbers of entries: pad to the maximum, gather, trim
    most = -(-batch_size // world)
    padded = torch.zeros((most,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
    padded[:nb] = y
    parts = [torch.empty_like(padded) for _ in range(world)]
    if y.is_complex():
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(padded), group=group)
    else:
        dist.all_gather(parts, padded, group=group)
    out = torch.cat([parts[r][: split_range(batch_size, world, r)[1] - split_range(batch_size, world, r)[0]]
                     for r in range(world)])
    return out, (b_lo, b_hi)


def nfft_forward_batch_sharded(xhat, pos, batch, cutoff=3, real_output=False, *, group=None, forward_fn=None):
    """Forward NFFT at the points of the batch entries owned by this rank.
    Returns (y_local, (p_lo, p_hi)): rows p_lo..p_hi of the full result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = xhat.shape[0]
    b_lo, b_hi, p_lo, p_hi = shard_batches(batch, B, world, rank)
    fn = forward_fn or _nfft.nfft_forward
    if b_hi > b_lo and p_hi > p_lo:
        y = fn(xhat[b_lo:b_hi], pos[p_lo:p_hi], batch[p_lo:p_hi] - b_lo, cutoff, real_output, batch_size=b_hi - b_lo)
    else:
        d = pos.shape[1]
        y = torch.zeros((0,) + tuple(xhat.shape[1 + d:]), dtype=torch.float32 if real_output else torch.complex64,
                        device=xhat.device)
    return y, (p_lo, p_hi)
