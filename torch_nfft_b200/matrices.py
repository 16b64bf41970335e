"""Implicit kernel matrices whose products are evaluated with nfft_fastsum.

API of reference `torch_nfft/matrices.py:5-175` (`GramMatrix`, `AdjacencyMatrix`).  Two defects of
the reference are not reproduced: `GramMatrix.is_symmetric` compared `sources` with itself
(matrices.py:65) and `AdjacencyMatrix.apply_shift` read an undefined name `shift` (matrices.py:149).
"""
import warnings

import torch

from .nfft import NfftPlan, nfft_fastsum


class AbstractMatrix:
    """A linear operator known only through its action `apply(x)`."""

    def __init__(self, shape, device):
        self.shape = shape
        self.device = device

    def apply(self, x):
        raise NotImplementedError()

    def __matmul__(self, x):
        return self.apply(x)

    def is_symmetric(self):
        return False

    def transpose(self):
        if not self.is_symmetric():
            raise NotImplementedError()
        return self

    @property
    def T(self):
        return self.transpose()

    def row_sums(self):
        return self.apply(torch.ones(self.shape[0], device=self.device))

    def column_sums(self):
        return self.T.row_sums()

    def to_dense(self):
        return self.apply(torch.eye(self.shape[0], device=self.device))


class GramMatrix(AbstractMatrix):
    """A[t, s] = K(sources[s] - targets[t]) for the trigonometric kernel given by `coeffs`."""

    def __init__(self, coeffs, sources, targets=None, source_batch=None, target_batch=None, /, batch=None, cutoff=3):
        if targets is None:
            targets, target_batch = sources, source_batch
        if batch is not None:
            source_batch = target_batch = batch
        super().__init__((sources.size(0), targets.size(0)), sources.device)
        self.coeffs, self.cutoff = coeffs, cutoff
        self.sources, self.targets = sources, targets
        self.source_batch, self.target_batch = source_batch, target_batch
        # the points of a matrix are fixed: they are binned by the first product and never again (the reference
        # recomputes its per-point scratch in every product, core_cuda.cu:188-211); CPU tensors raise in apply
        self._plans = None

    def _point_plans(self):
        if self._plans is None:
            src = NfftPlan(self.sources, self.source_batch)
            same = self.sources is self.targets and self.source_batch is self.target_batch
            self._plans = (src, src if same else NfftPlan(self.targets, self.target_batch, batch_size=src.batch_size))
        return self._plans

    def apply(self, x):
        source_plan, target_plan = self._point_plans()
        return nfft_fastsum(x, self.coeffs, self.sources, self.targets, self.source_batch, self.target_batch,
                            cutoff=self.cutoff, batch_size=source_plan.batch_size, source_plan=source_plan,
                            target_plan=target_plan)

    def is_symmetric(self):
        return self.sources is self.targets and self.source_batch is self.target_batch

    def transpose(self):
        if self.is_symmetric():
            return self
        flipped = GramMatrix(self.coeffs, self.targets, self.sources, self.target_batch, self.source_batch,
                             cutoff=self.cutoff)
        if self._plans is not None:  # the same binnings serve the transposed product
            flipped._plans = (self._plans[1], self._plans[0])
        return flipped


def _col(v, x):
    """Broadcast a per-node vector against x of shape [n, ...]."""
    return v.reshape(v.shape + (1,) * (x.dim() - 1))


class AdjacencyMatrix(AbstractMatrix):
    """Graph matrix built on a symmetric Gram matrix W (+ diagonal_offset * I):

    normalization: None | "sym" (D^-1/2 W D^-1/2) | "left"/"rw" (D^-1 W) | "right" (W D^-1)
    shift:         None | "laplacian" (D or I minus the above) | "signless" (plus)
    """

    _NORMALIZATIONS = ("none", "sym", "left", "right")

    def __init__(self, gram_matrix, diagonal_offset=0, normalization=None, shift=None, degree_threshold=0):
        if not gram_matrix.is_symmetric():
            raise ValueError("The underlying Gram matrix of an AdjacencyMatrix must be symmetric")
        super().__init__(gram_matrix.shape, gram_matrix.device)
        self.gram_matrix = gram_matrix
        self.diagonal_offset = diagonal_offset

        normalization = "none" if normalization is None else normalization.lower()
        if normalization == "rw":
            normalization = "left"
        if normalization not in self._NORMALIZATIONS:
            raise ValueError(f"Unknown AdjacencyMatrix normalization type: {normalization}")
        shift = "none" if shift is None else shift.lower()
        if shift not in ("none", "laplacian", "signless"):
            raise ValueError(f"Unknown AdjacencyMatrix shift type: {shift}")
        self.normalization, self.shift = normalization, shift

        if normalization == "none" and shift == "none":
            return
        degrees = gram_matrix.row_sums()
        if diagonal_offset != 0:
            degrees = degrees + diagonal_offset
        if normalization == "none":
            self.degrees = degrees
            return
        small = degrees < degree_threshold
        if torch.any(small):
            warnings.warn("AdjacencyMatrix with normalization: {} out of {} node degrees are smaller than the "
                          "threshold {:.4g}".format(int(small.sum()), degrees.numel(), degree_threshold),
                          RuntimeWarning, stacklevel=2)
            degrees = torch.where(small, torch.full_like(degrees, float("inf")), degrees)
        if normalization == "sym":
            self.d_inv_sqrt = torch.rsqrt(degrees)
        else:
            self.d_inv = 1.0 / degrees

    def _scale(self, x, side):
        if self.normalization == "sym":
            return _col(self.d_inv_sqrt, x) * x
        if self.normalization == side:
            return _col(self.d_inv, x) * x
        return x

    def apply_left_normalization(self, x):
        return self._scale(x, "left")

    def apply_right_normalization(self, x):
        return self._scale(x, "right")

    def apply_shift(self, x, y):
        if self.shift == "none":
            return y
        if self.normalization == "none":
            x = _col(self.degrees, x) * x
        return x + y if self.shift == "signless" else x - y

    def apply(self, x):
        scaled = self.apply_right_normalization(x)
        y = self.gram_matrix @ scaled
        if self.diagonal_offset != 0:
            y = y + self.diagonal_offset * scaled
        return self.apply_shift(x, self.apply_left_normalization(y))

    def is_symmetric(self):
        return self.normalization not in ("left", "right")

    def transpose(self):
        if self.is_symmetric():
            return self
        # share the degree data instead of recomputing the row sums
        flipped = AdjacencyMatrix(self.gram_matrix, self.diagonal_offset)
        flipped.normalization = "right" if self.normalization == "left" else "left"
        flipped.shift = self.shift
        flipped.d_inv = self.d_inv
        return flipped
