"""CUDA-graph capture of transform sequences (new).

The engine enqueues everything on the caller's stream and never synchronises, allocates or creates
plans after the first call with a given shape (the reference device-synchronises after every kernel,
reference csrc/cuda/cuda_utils.cu:7-14, and creates its cuFFT plan inside every call,
csrc/cuda/core_cuda.cu:254-272), so a fixed sequence of transforms on fixed tensors can be captured
once and replayed with new *contents* in those tensors.  This removes the per-launch host cost, which
dominates small transforms: BASELINE config c2 (1D, N=1024, 2^20 points, 14 kernels per
adjoint+forward pair) runs in 0.12 ms per pair replayed against 0.21 ms launched eagerly on a B200.

    pos, x, batch = ...                                  # static CUDA tensors
    def pair():
        y = torch_nfft_b200.nfft_adjoint(x, pos, batch, N=1024, m=8, batch_size=64)
        return y, torch_nfft_b200.nfft_forward(y, pos, batch, m=8, real_output=True, batch_size=64)
    g = torch_nfft_b200.GraphedTransforms(pair)
    pos.copy_(new_pos); x.copy_(new_x)                   # new data, same buffers
    y, f = g.replay()                                    # the tensors returned at capture time, refilled

Rules (those of CUDA graphs): the captured function must not synchronise with the host, so pass
`batch_size=` or `batch_ptr=` (otherwise `batch[-1].item()` is read, reference core_cuda.cu:60); shapes,
dtypes and tensor addresses are frozen.  Points are binned inside the captured function -- by the
transforms themselves or by an `NfftPlan` created inside it -- so every replay bins the positions it
finds in the tensors; a plan made OUTSIDE the function is baked in and only valid while `pos` is unchanged.

cuFFT: the captured FFT kernels take their work area from the capture stream's workspace (kept alive by
this object) and reference the twiddle tables of the cached cuFFT handles, so the plan cache is pinned
(`nfftb200_plan_cache_pin`) until `close()`: `clear_caches()` keeps the handles and warns meanwhile.
"""
from __future__ import annotations

import torch

from . import _lib
from . import nfft as _nfft
from .nfft import release_stream_workspace


class GraphedTransforms:
    def __init__(self, fn, warmup: int = 1, device=None):
        self.graph = None
        self.outputs = None
        if not torch.cuda.is_available():
            raise RuntimeError("torch_nfft_b200: GraphedTransforms needs a CUDA device")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            current = torch.cuda.current_stream()
            self._stream = torch.cuda.Stream()
            self._stream.wait_stream(current)
            # plans, kernel attributes and the capture stream's workspace are created outside the capture
            with torch.cuda.stream(self._stream):
                for _ in range(max(1, warmup)):
                    fn()
            current.wait_stream(self._stream)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            _lib.lib().nfftb200_plan_cache_pin(1)
            self._pinned = True
            with torch.cuda.graph(self.graph, stream=self._stream):
                self.outputs = fn()
            self.kernels_per_replay = _lib.launch_count() - before
            # the captured kernels hold raw pointers into the capture stream's workspace: keep it alive even if
            # the cache entry is replaced
            self._workspace = _nfft._workspaces.get((self.device.index, self._stream.cuda_stream))

    def replay(self):
        """Runs the captured transforms on the current stream; returns the output tensors of the capture."""
        if self.graph is None:
            raise RuntimeError("torch_nfft_b200: this GraphedTransforms has been closed")
        self.graph.replay()
        return self.outputs

    def close(self):
        """Frees the graph, its outputs and the capture stream's workspace (0.9 GB at BASELINE config c4)."""
        if getattr(self, "graph", None) is None:
            return
        torch.cuda.synchronize(self.device)  # replays still in flight use the workspace
        self.graph = None
        self.outputs = None
        self._workspace = None
        if getattr(self, "_pinned", False):
            _lib.lib().nfftb200_plan_cache_pin(-1)
            self._pinned = False
        release_stream_workspace(self.device.index, self._stream.cuda_stream)

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: modules may already be gone
            pass
