"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/nfft_b200.h, validates arguments without a GPU, and the Python layer fails loudly
(no CPU fallback) for CPU tensors or a missing library."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from torch_nfft_b200 import _build, _lib
import torch_nfft_b200 as T


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nfft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfftb200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared_symbols()
    assert len(names) >= 14
    handle = ctypes.CDLL(_build.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/nfft_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes binding and header disagree"


def test_version_and_error_string():
    L = _lib.lib()
    assert L.nfftb200_version() >= 100
    assert isinstance(L.nfftb200_last_error(), bytes)


@pytest.mark.parametrize("d,N,m,B,C,n", [(1, 1024, 8, 64, 1, 2 ** 20), (2, 256, 4, 16, 8, 2 ** 23), (3, 128, 4, 4, 1, 2 ** 24)])
def test_workspace_bytes_at_baseline_configs(d, N, m, B, C, n):
    L = _lib.lib()
    M = 2 * N
    nbytes = L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, n, 0, d, N, m, B, C, 0)
    grid = B * C * M ** d * 4
    half = B * C * M ** (d - 1) * (M // 2 + 1) * 8
    assert nbytes >= grid + half + 8 * n  # grid + spectrum + keys/permutation
    assert nbytes < grid + half + 64 * n + (1 << 26)  # and no per-point psi scratch (reference: 4*d*(2m+3) B/pt)


def test_invalid_arguments_are_rejected_without_a_gpu():
    L = _lib.lib()
    assert L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, 10, 0, 4, 16, 3, 1, 1, 0) == 0  # d = 4
    assert b"dimension" in L.nfftb200_last_error()
    assert L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, 10, 0, 2, 15, 3, 1, 1, 0) == 0  # odd N
    assert L.nfftb200_workspace_bytes(_lib.OP_ADJOINT, 10, 0, 2, 16, 9, 1, 1, 0) == 0  # m > 8
    # null pointers: status code + message, no crash, no CUDA call
    st = L.nfftb200_adjoint(0, 0, 0, 0, 10, 2, 16, 3, 1, 1, 0, 0, 0, 0)
    assert st == -1 and b"null" in L.nfftb200_last_error()


def test_geometry_query():
    g = _lib.geometry(3, 128, 4, 4, 1, 0, 2 ** 24)
    assert (g["M"], g["L"]) == (256, 10)
    assert g["Px"] % 4 == 0 and g["Px"] >= g["Tx"] + g["L"] - 1
    assert g["tile_elems"] * g["ncomp"] * 4 < 227 * 1024
    assert g["spread_threads"] >= g["L"] ** 2


def test_clustered_hint_changes_only_the_binning_geometry():
    """NFFTB200_CLUSTERED (NfftPlan(clustered=True)) enables the per-tile density decision for large 3D sets with
    cutoff 3 or 4: the keys carry the 2 x 2 x 2 hierarchy, left-aligned so that the lowest radix pass is fine bits
    only; everything else of the tiling is unchanged.  Host-only query."""
    plain = _lib.geometry(3, 128, 4, 4, 1, 0, 2 ** 24)
    hinted = _lib.geometry(3, 128, 4, 4, 1, _lib.CLUSTERED, 2 ** 24)
    assert plain["mixed"] == 0 and (plain["scx"], plain["scy"], plain["scz"], plain["fine_bits"]) == (4, 4, 2, 2)
    assert hinted["mixed"] == 1 and hinted["refine_pass"] == 1 and hinted["dense_tile_pts"] == 2048
    assert (hinted["scx"], hinted["scy"], hinted["scz"], hinted["fine_bits"]) == (2, 2, 2, 10)  # 14 tile bits + 10 = 3 passes
    for k in ("M", "Tx", "Ty", "Tz", "ntx", "Px", "Py", "Pz", "sY", "sZ", "tile_elems", "pmax", "use_reg"):
        assert plain[k] == hinted[k]
    # small point sets, other cutoffs, 2D and dense sets ignore the hint
    assert _lib.geometry(3, 128, 4, 4, 1, _lib.CLUSTERED, 1000)["mixed"] == 0
    assert _lib.geometry(3, 128, 2, 4, 1, _lib.CLUSTERED, 2 ** 24)["mixed"] == 0
    assert _lib.geometry(2, 256, 4, 16, 8, _lib.CLUSTERED, 2 ** 23)["mixed"] == 0
    dense = _lib.geometry(3, 64, 4, 1, 1, _lib.CLUSTERED, 2 ** 26)
    assert dense["mixed"] == 0 and (dense["scx"], dense["scy"], dense["scz"]) == (2, 2, 2)
    # the test hook switches the mode for every call and lowers the size threshold
    L = _lib.lib()
    try:
        L.nfftb200_debug_mixed(1, 0, 600)
        forced = _lib.geometry(3, 32, 4, 2, 1, 0, 30000)
        assert forced["mixed"] == 1 and forced["dense_tile_pts"] == 600 and forced["fine_bits"] == 9
        L.nfftb200_debug_mixed(0, 0, 0)
        assert _lib.geometry(3, 128, 4, 4, 1, _lib.CLUSTERED, 2 ** 24)["mixed"] == 0
    finally:
        L.nfftb200_debug_mixed(-1, -1, 0)
    assert _lib.geometry(3, 128, 4, 4, 1, _lib.CLUSTERED, 2 ** 24)["mixed"] == 1


def test_cpu_tensors_raise_like_the_reference():
    """reference csrc/core.cpp:52 asserts x.device().is_cuda(); so do we -- no CPU fallback."""
    pos = torch.rand(10, 2) - 0.5
    x = torch.rand(10)
    with pytest.raises(RuntimeError):
        T.nfft_adjoint(x, pos, None, 16, 3)
    with pytest.raises(RuntimeError):
        T.nfft_forward(torch.rand(1, 16, 16), pos, None, 3)
    with pytest.raises(RuntimeError):
        T.nfft_fastsum(x, torch.rand(16, 16), pos)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libnfft_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_path_does_not_import_the_oracle():
    import subprocess
    import sys
    code = "import sys; import torch_nfft_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for fn in os.listdir(os.path.join(ROOT, "torch_nfft_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "torch_nfft_b200", fn)).read().replace("accuracy oracle", "")
