"""GPU diagnostic: stage-by-stage comparison of the CUDA engine with the numpy oracle.

Usage: python scripts/diag.py [case ...]   (each case runs in its own subprocess so that a CUDA
fault in one case does not poison the others).  Prints one line per check.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (d, N, m, B, n_per_batch, C)
    "1d": (1, 64, 4, 2, 500, 2),
    "1d_m8": (1, 1024, 8, 3, 3000, 1),
    "2d": (2, 32, 4, 3, 700, 3),
    "2d_c8": (2, 64, 3, 2, 2000, 8),
    "3d": (3, 16, 3, 2, 600, 1),
    "3d_m4": (3, 32, 4, 2, 3000, 1),
    "3d_c3": (3, 16, 2, 1, 500, 3),
    "3d_big": (3, 64, 4, 1, 40000, 1),
}


def run_case(name):
    import ctypes
    import numpy as np
    import torch
    from oracle import nfft_oracle as O
    import torch_nfft_b200 as T
    from torch_nfft_b200 import _lib

    d, N, m, B, n, C = CASES[name]
    rng = np.random.default_rng(1)
    pos = (rng.random((n * B, d), dtype=np.float32) - 0.5)
    batch = np.repeat(np.arange(B), n)
    dev = torch.device("cuda")
    tpos = torch.from_numpy(pos).to(dev)
    tbatch = torch.from_numpy(batch).to(dev)
    L = _lib.lib()
    M = 2 * N
    print(f"[{name}] d={d} N={N} m={m} B={B} n={n*B} C={C} geom={_lib.geometry(d, N, m, B, C, 0, n*B)}", flush=True)

    def ws_for(op, ns, nt, flags):
        nb = L.nfftb200_workspace_bytes(op, ns, nt, d, N, m, B, C, flags)
        return torch.empty(nb, dtype=torch.uint8, device=dev)

    # ---- sort
    nn = n * B
    keys = torch.zeros(nn, dtype=torch.int32, device=dev)
    perm = torch.zeros(nn, dtype=torch.int32, device=dev)
    tile = (ctypes.c_int32 * 3)()
    ws = ws_for(_lib.OP_SORT, nn, 0, 0)
    st = L.nfftb200_sort_points(tpos.data_ptr(), tbatch.data_ptr(), keys.data_ptr(), perm.data_ptr(),
                                ctypes.cast(tile, ctypes.c_void_p), nn, d, N, m, B, C, 0, ws.data_ptr(), ws.numel(), 0)
    torch.cuda.synchronize()
    print(f"[{name}] sort status {st} tile {list(tile)}")
    tl = list(tile)[:d][::-1]  # API order (dim 0 = slowest slot)
    okeys = O.tile_keys(pos, batch, N, tl)
    operm = O.stable_permutation(okeys)
    kk = keys.cpu().numpy().astype(np.int64) & 0xffffffff
    pp = perm.cpu().numpy().astype(np.int64) & 0xffffffff
    print(f"[{name}] keys equal: {np.array_equal(kk, okeys)}  perm equal: {np.array_equal(pp, operm)}", flush=True)

    for cplx in (False, True):
        flags = _lib.X_COMPLEX if cplx else 0
        x = rng.standard_normal((nn, C)).astype(np.float32)
        if cplx:
            x = (x + 1j * rng.standard_normal((nn, C))).astype(np.complex64)
        tx = torch.from_numpy(x).to(dev)
        # ---- spread
        grid = torch.full((B * C * M ** d,), 7.0, dtype=torch.complex64 if cplx else torch.float32, device=dev)
        ws = ws_for(_lib.OP_SPREAD, nn, 0, flags)
        st = L.nfftb200_spread(tpos.data_ptr(), tx.data_ptr(), tbatch.data_ptr(), grid.data_ptr(), nn, d, N, m, B, C,
                               flags, ws.data_ptr(), ws.numel(), 0)
        torch.cuda.synchronize()
        og = O.spread(pos, x.reshape(nn, C), batch, B, N, m)
        gg = grid.cpu().numpy().reshape(og.shape)
        print(f"[{name}] cplx={cplx} spread status {st} rel_l2 {O.rel_l2(gg, og if cplx else og.real):.3e}", flush=True)
        # ---- gather (from the oracle grid)
        gin = torch.from_numpy(og.astype(np.complex64) if cplx else og.real.astype(np.float32)).to(dev).contiguous()
        y = torch.zeros((nn, C), dtype=torch.complex64 if cplx else torch.float32, device=dev)
        ws = ws_for(_lib.OP_GATHER, 0, nn, flags)
        st = L.nfftb200_gather(tpos.data_ptr(), tbatch.data_ptr(), gin.data_ptr(), y.data_ptr(), nn, d, N, m, B, C,
                               flags, ws.data_ptr(), ws.numel(), 0)
        torch.cuda.synchronize()
        oy = O.gather(gin.cpu().numpy().reshape(og.shape).astype(np.complex64), pos, batch, N, m)
        print(f"[{name}] cplx={cplx} gather status {st} rel_l2 {O.rel_l2(y.cpu().numpy(), oy if cplx else oy.real):.3e}", flush=True)
        # ---- full transforms
        for ro in (False, True):
            ya = T.nfft_adjoint(tx, tpos, tbatch, N, m, real_output=ro)
            oa = O.nfft_adjoint(x, pos, batch, N, m, real_output=ro)
            print(f"[{name}] cplx={cplx} real_out={ro} adjoint rel_l2 {O.rel_l2(ya.cpu().numpy(), oa):.3e}", flush=True)
        xh = rng.standard_normal((B,) + (N,) * d + (C,)).astype(np.float32)
        if cplx:
            xh = (xh + 1j * rng.standard_normal(xh.shape)).astype(np.complex64)
        txh = torch.from_numpy(xh).to(dev)
        for ro in (False, True):
            yf = T.nfft_forward(txh, tpos, tbatch, m, real_output=ro)
            of = O.nfft_forward(xh, pos, batch, m, real_output=ro)
            print(f"[{name}] xhat cplx={cplx} real_out={ro} forward rel_l2 {O.rel_l2(yf.cpu().numpy(), of):.3e}", flush=True)
        # ---- fastsum (symmetric and not)
        for ccplx in (False, True):
            co = O.gaussian_interpolated_coeffs(0.15, d, N) if ccplx else O.gaussian_analytic_coeffs(0.15, d, N)
            if ccplx:
                co = (co * np.exp(0.3j)).astype(np.complex64)  # not Hermitian-even on purpose
            tco = torch.from_numpy(co).to(dev)
            spos = (pos * 0.5).astype(np.float32)
            tsp = torch.from_numpy(spos).to(dev)
            yfs = T.nfft_fastsum(tx, tco, tsp, batch=tbatch, cutoff=m)
            ofs = O.nfft_fastsum(x, co, spos, None, batch, batch, m=m)
            print(f"[{name}] x cplx={cplx} coeffs cplx={ccplx} fastsum(sym) rel_l2 {O.rel_l2(yfs.cpu().numpy(), ofs):.3e}", flush=True)
            tpos2 = (rng.random((nn // 2 * 1, d), dtype=np.float32) - 0.5) * 0.5
            b2 = np.sort(rng.integers(0, B, size=tpos2.shape[0])).astype(np.int64)
            b2[-1] = B - 1
            yfs = T.nfft_fastsum(tx, tco, tsp, torch.from_numpy(tpos2).to(dev), tbatch, torch.from_numpy(b2).to(dev), cutoff=m)
            ofs = O.nfft_fastsum(x, co, spos, tpos2, batch, b2, m=m)
            print(f"[{name}] x cplx={cplx} coeffs cplx={ccplx} fastsum(s!=t) rel_l2 {O.rel_l2(yfs.cpu().numpy(), ofs):.3e}", flush=True)
    print(f"[{name}] launches so far {_lib.launch_count()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    rc = 0
    for nm in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", nm], timeout=600)
            if r.returncode != 0:
                print(f"[{nm}] FAILED rc={r.returncode}", flush=True)
                rc = 1
        except subprocess.TimeoutExpired:
            print(f"[{nm}] TIMEOUT", flush=True)
            rc = 1
    sys.exit(rc)
