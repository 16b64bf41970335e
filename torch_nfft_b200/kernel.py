"""Gaussian kernel front end: scales points into the torus and builds fastsum Gram matrices.

API of reference `torch_nfft/kernel.py:9-126`.
"""
import math

from .coeffs import gaussian_analytic_coeffs, gaussian_interpolated_coeffs
from .matrices import AdjacencyMatrix, GramMatrix
from .utils import scale_points_by_norm, shift_points_by_center


class GaussianKernel:
    r"""Approximation of K(z) = exp(-|z|^2 / sigma^2) with fast Gram-matrix products.

        kernel = GaussianKernel(sigma, dim=3, bandwidth=16, cutoff=3)
        A = kernel(sources, targets, batch=batch)        # GramMatrix
        y = A @ x                                        # nfft_fastsum

    Points must end up in a ball of radius 1/4 (minus half the regularisation width) so that all
    differences lie in the torus cell.  Either the radius rho of the data is given up front
    (`max_euclidean_norm` / `max_infinity_norm`): then points are multiplied by a fixed factor and
    the kernel above is approximated.  Or it is not: then every point set is scaled by its own
    radius rho and the kernel is exp(-|z|^2 / (rho sigma)^2).

    Parameters: `sigma`, `dim`, `bandwidth` (N, a small power of two), `cutoff` (m),
    `shift_by_center` (translate each point set to the origin first), `analytic` (closed-form
    coefficients instead of interpolated ones), `reg_degree` / `reg_width` (p and eps of
    gaussian_interpolated_coeffs).
    """

    def __init__(self, sigma, dim=3, bandwidth=16, cutoff=3, shift_by_center=True, max_euclidean_norm=None,
                 max_infinity_norm=None, analytic=False, reg_degree=-1, reg_width=0.0):
        self.cutoff = cutoff
        self.shift_by_center = shift_by_center
        self.factor = 0.25 - 0.5 * reg_width
        if reg_degree < 0:
            # no regularisation: only the box matters, the infinity norm is enough
            radius, fallback = (max_infinity_norm or max_euclidean_norm), "infinity"
        else:
            radius = max_euclidean_norm
            if radius is None and max_infinity_norm is not None:
                radius = max_infinity_norm * math.sqrt(dim)
            fallback = "euclidean"
        self.scale_by_norm = fallback if radius is None else None
        if radius is not None:
            self.factor /= radius
        if analytic:
            self.coeffs = gaussian_analytic_coeffs(self.factor * sigma, dim, bandwidth)
        else:
            self.coeffs = gaussian_interpolated_coeffs(self.factor * sigma, dim, bandwidth, reg_degree, reg_width)

    def gram_matrix(self, sources, targets=None, source_batch=None, target_batch=None, /, batch=None):
        if batch is not None:
            source_batch = target_batch = batch
        if self.shift_by_center:
            sources, targets = shift_points_by_center(sources, targets, source_batch, target_batch)
        if self.scale_by_norm is None:
            sources = self.factor * sources
            targets = None if targets is None else self.factor * targets
        else:
            sources, targets = scale_points_by_norm(sources, targets, source_batch, target_batch,
                                                    factor=self.factor, norm=self.scale_by_norm)
        return GramMatrix(self.coeffs, sources, targets, source_batch, target_batch, cutoff=self.cutoff)

    __call__ = gram_matrix

    def adjacency_matrix(self, sources, batch=None, loop_weight=1, normalization=None, shift=None,
                         degree_threshold=0):
        return AdjacencyMatrix(self.gram_matrix(sources, batch=batch), diagonal_offset=loop_weight - 1,
                               normalization=normalization, shift=shift, degree_threshold=degree_threshold)
