"""ctypes binding of the C ABI declared in include/nfft_b200.h.

There is NO fallback: if the CUDA library is missing or a call fails, a RuntimeError is raised
(the reference likewise hard-requires CUDA tensors, reference csrc/core.cpp:52).
"""
import ctypes
import os
import threading

from ._build import LIB_PATH as _DEFAULT_LIB_PATH

# development aid: NFFTB200_LIB selects an alternative build of the same library (A/B experiments)
LIB_PATH = os.environ.get("NFFTB200_LIB", _DEFAULT_LIB_PATH)

OP_ADJOINT, OP_FORWARD, OP_FASTSUM, OP_SPREAD, OP_GATHER, OP_SORT, OP_SPECTRAL, OP_PLAN = range(8)
X_COMPLEX, Y_REAL, COEFFS_COMPLEX, SYMMETRIC, PLANNED, BATCH_OFFSETS, CLUSTERED = 1, 2, 4, 8, 16, 32, 64

_lock = threading.Lock()
_lib = None

_i64, _i32, _vp, _sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t

_SIGNATURES = {
    "nfftb200_version": (ctypes.c_int, []),
    "nfftb200_last_error": (ctypes.c_char_p, []),
    "nfftb200_launch_count": (_i64, []),
    "nfftb200_plan_cache_clear": (ctypes.c_int, []),
    "nfftb200_plan_cache_pin": (ctypes.c_int, [_i32]),
    "nfftb200_plan_cache_size": (ctypes.c_int, []),
    "nfftb200_debug_force_int64": (None, [_i32]),
    "nfftb200_debug_mixed": (None, [_i32, _i64, _i32]),
    "nfftb200_debug_min_resident_ctas": (_i32, []),
    "nfftb200_debug_pruned_fft": (None, [_i32]),
    # (n, n_geom, d, N, m, B, C, flags)
    "nfftb200_plan_bytes": (_sz, [_i64, _i64, _i32, _i64, _i32, _i64, _i64, _i32]),
    # (pos, batch, plan, plan_bytes, n, n_geom, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_plan_points": (_i32, [_vp, _vp, _vp, _sz, _i64, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (plan, n, n_geom, d, N, m, B, C, flags, flags_out_host, stream)
    "nfftb200_plan_flags": (_i32, [_vp, _i64, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _vp]),
    # (pos, x, batch, plan, plan_bytes, y, n, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_adjoint_planned": (_i32, [_vp, _vp, _vp, _vp, _sz, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    "nfftb200_forward_planned": (_i32, [_vp, _vp, _vp, _vp, _sz, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (src, tgt, x, coeffs, sb, tb, src_plan, src_plan_bytes, tgt_plan, tgt_plan_bytes, y, n_src, n_tgt, d, N, m, B, C,
    #  flags, ws, ws_bytes, stream)
    "nfftb200_fastsum_planned": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _i64, _i64, _i32, _i64, _i32,
                                        _i64, _i64, _i32, _vp, _sz, _vp]),
    "nfftb200_profile_enable": (None, [_i32]),
    "nfftb200_profile_read": (_i32, [_vp, _vp]),
    "nfftb200_debug_geometry": (_i32, [_i32, _i64, _i32, _i64, _i64, _i32, _i64, _vp]),
    "nfftb200_workspace_bytes": (_sz, [_i32, _i64, _i64, _i32, _i64, _i32, _i64, _i64, _i32]),
    # (pos, x, batch, y, n, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_adjoint": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    "nfftb200_forward": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (src, tgt, x, coeffs, sb, tb, y, n_src, n_tgt, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_fastsum": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i64, _i32, _i64, _i64, _i32,
                                _vp, _sz, _vp]),
    # (pos, x, batch, plan, plan_bytes, grid, n, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_spread": (_i32, [_vp, _vp, _vp, _vp, _sz, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (pos, batch, plan, plan_bytes, grid, y, n, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_gather": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (grid, y, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_adjoint_finish": (_i32, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (xhat, grid, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_forward_begin": (_i32, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (grid, coeffs, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_fastsum_middle": (_i32, [_vp, _vp, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
    # (pos, batch, keys_out, perm_out, tile_out_host, n, d, N, m, B, C, flags, ws, ws_bytes, stream)
    "nfftb200_sort_points": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _i32, _vp, _sz, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded C-ABI library; raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "torch_nfft_b200: CUDA library %s is missing. Build it with "
                        "`python -m torch_nfft_b200._build` (needs nvcc); there is no CPU fallback." % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().nfftb200_last_error()
        raise RuntimeError("torch_nfft_b200.%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(lib().nfftb200_launch_count())


GEOMETRY_FIELDS = ("dim", "N", "M", "m", "L", "Tx", "Ty", "Tz", "ntx", "nty", "ntz", "Px", "Py", "Pz", "sY", "sZ",
                   "tile_elems", "ncomp", "pmax", "spread_threads", "use_reg", "fine_bits", "scx", "scy", "scz", "mixed",
                   "refine_pass", "dense_tile_pts")


def geometry(d, N, m, B=1, C=1, flags=0, n=0):
    """Tiling the engine uses for a transform (host-only query)."""
    out = (ctypes.c_int32 * 28)()
    check(lib().nfftb200_debug_geometry(d, N, m, B, C, flags, n, ctypes.cast(out, ctypes.c_void_p)), "geometry")
    return dict(zip(GEOMETRY_FIELDS, list(out)))


STAGES = ("sort", "spread", "fft", "unpack", "pack", "gather", "multiply", "memset")


def profile_enable(on: bool):
    lib().nfftb200_profile_enable(1 if on else 0)


def profile_read():
    """{stage: (total_ms, calls)} since the last read; synchronise the stream first."""
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_int64 * 8)()
    check(lib().nfftb200_profile_read(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(cnt, ctypes.c_void_p)), "profile_read")
    return {s: (ms[i], int(cnt[i])) for i, s in enumerate(STAGES)}
