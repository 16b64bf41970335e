"""Exact (non-approximating) direct sums in plain PyTorch, device agnostic.

Same semantics as reference `torch_nfft/ndft.py:5-117`; this is the accuracy oracle for the NFFT
(O(n N^d) work, so small sizes only) and the reference's only host (CPU) path.  Not accelerated
on purpose.
"""
import math

import torch


def _frequency_grid(N, d, device):
    k = torch.arange(-(N // 2), N - N // 2, dtype=torch.float32, device=device)
    mesh = torch.meshgrid(*([k] * d), indexing="ij")
    return torch.stack(mesh, dim=-1)  # [N]*d + [d]


def _batch_size(batch):
    return int(batch.max().item()) + 1


def ndft_adjoint(x, pos, batch=None, N=16):
    """y[b, k + N/2, ...] = sum_{i in b} x[i, ...] exp(+2 pi i k . pos[i])   (reference ndft.py:5-23)."""
    d = pos.shape[1]
    x = x.to(torch.cfloat)
    grid = _frequency_grid(N, d, pos.device)

    def single(xp, pp):
        phase = torch.tensordot(grid, pp, dims=([-1], [-1]))  # [N]*d + [n_b]
        return torch.tensordot(torch.exp(2j * math.pi * phase), xp, dims=1)[None]

    if batch is None:
        return single(x, pos)
    return torch.cat([single(x[batch == b], pos[batch == b]) for b in range(_batch_size(batch))])


def ndft_forward(x, pos, batch=None):
    """y[i, ...] = sum_k x[b_i, k + N/2, ...] exp(-2 pi i k . pos[i])   (reference ndft.py:26-44)."""
    d = pos.shape[1]
    x = x.to(torch.cfloat)
    N = x.shape[1]
    grid = _frequency_grid(N, d, pos.device)

    def single(xb, pp):
        phase = torch.tensordot(pp, grid, dims=([-1], [-1]))  # [n_b] + [N]*d
        return torch.tensordot(torch.exp(-2j * math.pi * phase), xb, dims=d)

    if batch is None:
        return single(x[0], pos)
    return torch.cat([single(x[b], pos[batch == b]) for b in range(_batch_size(batch))])


def ndft_fastsum(x, coeffs, sources, targets=None, source_batch=None, target_batch=None, batch=None, N=16):
    """ndft_forward(coeffs * ndft_adjoint(x))   (reference ndft.py:48-62).  Unlike the reference,
    a 1-D x (no channel dimension) is accepted."""
    if targets is None:
        targets, target_batch = sources, source_batch
    if batch is not None:
        source_batch = target_batch = batch
    N = coeffs.shape[0]
    y = ndft_adjoint(x, sources, source_batch, N=N)
    y = y * coeffs.reshape((1,) + tuple(coeffs.shape) + (1,) * (x.dim() - 1))
    y = ndft_forward(y, targets, target_batch)
    return y if x.is_complex() else y.real


def exact_trigonometric_matrix(coeffs, sources, targets=None, source_batch=None, target_batch=None, /, batch=None):
    """A[t, s] = sum_l coeffs[l + N/2] exp(2 pi i l . (sources[s] - targets[t]))   (reference ndft.py:66-95)."""
    if targets is None:
        targets, target_batch = sources, source_batch
    if batch is not None:
        source_batch = target_batch = batch
    d = coeffs.dim()
    N = coeffs.size(0)
    coeffs = coeffs.to(torch.cfloat)
    grid = _frequency_grid(N, d, coeffs.device)

    def single(sp, tp):
        diff = sp.reshape(1, -1, d) - tp.reshape(-1, 1, d)
        phase = torch.tensordot(grid, diff, dims=([-1], [-1]))
        return torch.tensordot(coeffs, torch.exp(2j * math.pi * phase), dims=d)

    if source_batch is None:
        return single(sources, targets)
    blocks = [single(sources[source_batch == b], targets[target_batch == b]) for b in range(_batch_size(source_batch))]
    return torch.block_diag(*blocks)


def exact_gaussian_matrix(sigma, sources, targets=None, source_batch=None, target_batch=None, batch=None):
    """A[t, s] = exp(-|sources[s] - targets[t]|^2 / sigma^2)   (reference ndft.py:98-117)."""
    if targets is None:
        targets, target_batch = sources, source_batch
    if batch is not None:
        source_batch = target_batch = batch

    def single(sp, tp):
        return torch.exp(-torch.cdist(tp, sp).pow(2) / (sigma ** 2))

    if source_batch is None:
        return single(sources, targets)
    blocks = [single(sources[source_batch == b], targets[target_batch == b]) for b in range(_batch_size(source_batch))]
    return torch.block_diag(*blocks)
