"""Trigonometric kernel coefficients (one-off O(N^d) setup; host-side PyTorch, not kernels).

Same names, argument order and results as reference `torch_nfft/coeffs.py:10-27`, whose CUDA
implementation lives in `csrc/cuda/kernel_coeffs.cu:6-202` / `csrc/cuda/core_cuda.cu:855-1064`.
Like the reference, tensors are created on the current CUDA device by default; `device=` can
override that (e.g. "cpu" for tests).
"""
import math

import torch


def _dev(device):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def gaussian_analytic_coeffs(sigma, dim=3, N=16, device=None):
    """b_l = prod_a sqrt(pi) sigma exp(-sigma^2 pi^2 l_a^2) at index l + N/2; float32 [N]^dim
    (kernel_coeffs.cu:6-30)."""
    dev = _dev(device)
    l = torch.arange(N, dtype=torch.float32, device=dev) - (N // 2)
    v = math.sqrt(math.pi) * sigma * torch.exp(-(sigma * sigma) * (math.pi ** 2) * l * l)
    out = torch.ones((N,) * dim, dtype=torch.float32, device=dev)
    for a in range(dim):
        shape = [1] * dim
        shape[a] = N
        out = out * v.reshape(shape)
    return out


def interpolation_grid(dim=3, N=16, device=None):
    """grid[i_0, ..., i_{d-1}, a] = i_a / N - 1/2; float32 [N]^dim x dim (kernel_coeffs.cu:76-97)."""
    dev = _dev(device)
    g = torch.arange(N, dtype=torch.float32, device=dev) / N - 0.5
    return torch.stack(torch.meshgrid(*([g] * dim), indexing="ij"), dim=-1)


def radial_interpolation_grid(dim=3, N=16, device=None):
    """Euclidean norm of interpolation_grid; float32 [N]^dim (kernel_coeffs.cu:99-123)."""
    return interpolation_grid(dim, N, device).pow(2).sum(-1).sqrt()


def interpolated_kernel_coeffs(grid_values):
    """Trigonometric interpolation of kernel samples on interpolation_grid:
    fftshift(fftn(ifftshift(values))) / N^d, complex64 [N]^dim
    (kernel_coeffs.cu:126-202, core_cuda.cu:995-1064)."""
    v = grid_values.to(torch.complex64)
    out = torch.fft.fftshift(torch.fft.fftn(torch.fft.ifftshift(v))) / v.numel()
    return out.to(torch.complex64)


def gaussian_interpolated_coeffs(sigma, dim=3, N=16, p=-1, eps=0.0, device=None):
    """Interpolated coefficients of exp(-r^2 / sigma^2); for p >= 0 the samples are held constant
    outside the ball r <= 1/2 (kernel_coeffs.cu:33-73).  Only p <= 0 and eps == 0 are implemented,
    as in the reference (core_cuda.cu:890-891)."""
    if p > 0 or eps != 0.0:
        raise RuntimeError("gaussian_interpolated_coeffs: only p <= 0 and eps == 0 are implemented")
    r2 = interpolation_grid(dim, N, device).pow(2).sum(-1)
    vals = torch.exp(-r2 / (sigma * sigma))
    if p >= 0:
        vals = torch.where(r2 <= 0.25, vals, torch.full_like(vals, math.exp(-0.25 / (sigma * sigma))))
    return interpolated_kernel_coeffs(vals)
