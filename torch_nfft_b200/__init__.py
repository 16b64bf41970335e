"""torch_nfft_b200 -- B200-native NFFT engine, drop-in for `torch_nfft` (dominikbuenger/torch_nfft).

    import torch_nfft_b200 as torch_nfft

exports the names of reference `torch_nfft/__init__.py:14-20`.  The hot path (nfft_adjoint,
nfft_forward, nfft_fastsum) runs in `libnfft_b200.so`, hand-written sm_100a CUDA behind the C ABI
of `include/nfft_b200.h`; there is no CPU or PyTorch fallback for it.
"""
from .nfft import nfft_forward, nfft_adjoint, nfft_fastsum, NfftPlan, clear_caches, register_torch_ops
from .ndft import ndft_forward, ndft_adjoint, ndft_fastsum, \
    exact_trigonometric_matrix, exact_gaussian_matrix
from .coeffs import gaussian_analytic_coeffs, gaussian_interpolated_coeffs, \
    interpolation_grid, radial_interpolation_grid, interpolated_kernel_coeffs
from .matrices import GramMatrix, AdjacencyMatrix
from .kernel import GaussianKernel
from .graph import GraphedTransforms

__version__ = "0.2.0"
