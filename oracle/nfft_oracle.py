"""CPU oracle for the NFFT hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm implemented by the reference
(dominikbuenger/torch_nfft, CUDA).  It exists to *check* the sm_100a engine in
`torch_nfft_b200`; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product path never does and
fails loudly when its CUDA library is missing.

Parity pinning: the reference ships NO golden vectors (its tests print norms of unseeded
random runs, reference `test/*.py`).  The oracle is therefore pinned two ways:
  * against outputs of the compiled reference CUDA ops on seeded inputs, produced on a
    B200 by `tests/golden/make_golden.py` and committed under `tests/golden/*.npz`
    (`tests/test_oracle.py::test_oracle_matches_reference_golden`);
  * against the exact direct sums `ndft_*` below (restating reference
    `torch_nfft/ndft.py:5-62`) to the NFFT's own approximation error per cutoff m.

All citations are relative to the reference repository root.

Numerics: `prec="f32"` evaluates window taps, roll-off factors and tap products in
float32 exactly as the reference's kernels do, but accumulates grid cells / outputs in
float64 (the reference accumulates with float atomics in a non-deterministic order, so
no sequential fp32 order is "the" reference order).  `prec="f64"` evaluates everything
in float64 and is the tie-breaker when two fp32 engines disagree by round-off.
"""
from __future__ import annotations

import itertools
import numpy as np

# csrc/cuda/spatial_window_operations.cu:1-6
THREE_QUARTER_PI = np.float32(2.356194490192344928846982537459627163147877049531)
# csrc/cuda/spectral_window_operations.cu:1-2
PI_THIRD = np.float32(1.047197551196597746154214461093167628065723133125)


# --------------------------------------------------------------------------------------
# window pieces
# --------------------------------------------------------------------------------------
def compute_cells(pos: np.ndarray, M: int) -> np.ndarray:
    """floorf(pos*M) per coordinate, in float32 like compute_shifts_kernel
    (csrc/cuda/spatial_window_operations.cu:50).  Returns int64 [n, d]."""
    p = np.asarray(pos, dtype=np.float32)
    return np.floor(p * np.float32(M)).astype(np.int64)


def compute_shifts(pos: np.ndarray, M: int, m: int) -> np.ndarray:
    """shift = (int)floorf(pos*M) - m  (spatial_window_operations.cu:50)."""
    return compute_cells(pos, M) - int(m)


def window_params(N: int, m: int, prec: str = "f32"):
    """(inv_b, inv_sqrt_b_pi) = (0.75*pi/m, sqrt(0.75/m))  (spatial_window_operations.cu:3-4)."""
    if prec == "f32":
        inv_b = np.float32(THREE_QUARTER_PI / np.float32(m))
        s = np.float32(np.sqrt(np.float32(0.75) / np.float32(m)))
    else:
        inv_b = 0.75 * np.pi / m
        s = np.sqrt(0.75 / m)
    return inv_b, s


def compute_psi(pos: np.ndarray, shifts: np.ndarray, N: int, m: int, prec: str = "f32") -> np.ndarray:
    """psi[i, a, l] = phi(pos*2N - shift - l), l in [0, 2m+2)
    (spatial_window_operations.cu:84-86 with eval_phi at :24-28).  The argument is formed in
    double (the literal 2.0 promotes) and narrowed to float at the eval_phi call."""
    L = 2 * m + 2
    p64 = np.asarray(pos, dtype=np.float32).astype(np.float64)
    arg = p64[:, :, None] * 2.0 * N - shifts[:, :, None].astype(np.float64) - np.arange(L, dtype=np.float64)[None, None, :]
    inv_b, s = window_params(N, m, prec)
    if prec == "f32":
        t = arg.astype(np.float32)
        return (np.exp(-(t * t) * inv_b) * s).astype(np.float32)
    return np.exp(-(arg * arg) * inv_b) * s


def phi_hat_inv(N: int, m: int, prec: str = "f32") -> np.ndarray:
    """phi_hat_inv[k] = expf(float(k*k) * (pi/3) m / N^2), k = 0..N/2
    (spectral_window_operations.cu:2,14-18,27-43)."""
    k = np.arange(N // 2 + 1, dtype=np.int64)
    if prec == "f32":
        c = np.float32(np.float32(PI_THIRD * np.float32(m)) / np.float32(N * N))
        return np.exp((k * k).astype(np.float32) * c).astype(np.float32)
    c = (np.pi / 3.0) * m / (N * N)
    return np.exp((k * k).astype(np.float64) * c)


def rolloff_factors(N: int, m: int, d: int, prec: str = "f32") -> np.ndarray:
    """prod_a phi_hat_inv[|k_a|] laid out on the API frequency grid [N]^d (index k+N/2),
    multiplied dimension 0 first as the reference loop does
    (spectral_window_operations.cu:78-99)."""
    ph = phi_hat_inv(N, m, prec)
    k = np.abs(np.arange(N) - N // 2)
    f1 = ph[k]
    out = np.ones((N,) * d, dtype=ph.dtype)
    for a in range(d):
        shape = [1] * d
        shape[a] = N
        out = (out * f1.reshape(shape)).astype(ph.dtype)
    return out


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _as_points(pos, batch):
    pos = np.asarray(pos, dtype=np.float32)
    assert pos.ndim == 2 and 1 <= pos.shape[1] <= 3  # core_cuda.cu:47-53
    n = pos.shape[0]
    if batch is None:
        batch = np.zeros(n, dtype=np.int64)  # core_cuda.cu:63
        B = 1
    else:
        batch = np.asarray(batch, dtype=np.int64)
        B = int(batch[-1]) + 1 if n > 0 else 1  # core_cuda.cu:60
    return pos, batch, B


def _columns(x, n):
    """view x as [n, C] like core_cuda.cu:82-84,219."""
    x = np.asarray(x)
    cols = x.shape[1:]
    C = int(np.prod(cols)) if len(cols) else 1
    return x.reshape(n, C), cols, C


def spread(pos, x2, batch, B, N, m, prec="f32"):
    """Adjoint window convolution: g[(b*C+c), j] += x[i,c] * prod_a psi[i,a,l_a],
    j_a = (shift_a + l_a + M) mod M  (spatial_window_operations.cu:146-156, 197-206).
    Returns complex128 grid [B, C, M, ..., M] (real inputs leave imag = 0)."""
    n, d = pos.shape
    C = x2.shape[1]
    M, L = 2 * N, 2 * m + 2
    sh = compute_shifts(pos, M, m)
    psi = compute_psi(pos, sh, N, m, prec)
    g = np.zeros((B * C * M ** d,), dtype=np.complex128)
    xt = x2.astype(np.complex64 if prec == "f32" else np.complex128)
    wdt = np.float32 if prec == "f32" else np.float64
    bc = batch[:, None] * C + np.arange(C)[None, :]  # [n, C]
    for ls in itertools.product(range(L), repeat=d):
        idx = bc.copy()
        val = xt.copy()
        for a in range(d):
            ja = np.mod(sh[:, a] + ls[a] + M, M)
            idx = idx * M + ja[:, None]
            # value *= psi (float * float / complex<float> * float), dimension 0 first
            w = psi[:, a, ls[a]].astype(wdt)[:, None]
            val = (val * w).astype(val.dtype)
        flat = idx.ravel()
        g.real += np.bincount(flat, weights=val.real.ravel().astype(np.float64), minlength=g.size)
        if np.iscomplexobj(x2):
            g.imag += np.bincount(flat, weights=val.imag.ravel().astype(np.float64), minlength=g.size)
    return g.reshape((B, C) + (M,) * d)


def gather(g, pos, batch, N, m, prec="f32"):
    """Forward window convolution: y[i,c] = sum_l prod_a psi[i,a,l_a] * g[(b,c), j]
    (spatial_window_operations.cu:257-267).  g: [B, C, M..M] complex.  Returns complex128 [n, C]."""
    n, d = pos.shape
    B, C = g.shape[:2]
    M, L = 2 * N, 2 * m + 2
    sh = compute_shifts(pos, M, m)
    psi = compute_psi(pos, sh, N, m, prec)
    gf = g.reshape(B * C, -1)
    wdt = np.float32 if prec == "f32" else np.float64
    gt = gf.astype(np.complex64 if prec == "f32" else np.complex128)
    y = np.zeros((n, C), dtype=np.complex128)
    for ls in itertools.product(range(L), repeat=d):
        idx = np.zeros(n, dtype=np.int64)
        fac = np.ones(n, dtype=wdt)
        for a in range(d):
            idx = idx * M + np.mod(sh[:, a] + ls[a] + M, M)
            fac = (fac * psi[:, a, ls[a]].astype(wdt)).astype(wdt)
        for c in range(C):
            vals = gt[batch * C + c, idx]
            y[:, c] += (vals * fac).astype(np.complex128)
    return y


def _band_index(N, M):
    """grid index k mod M for API index i = k + N/2 (spectral_window_operations.cu:80-97)."""
    k = np.arange(N) - N // 2
    return np.mod(k, M)


# --------------------------------------------------------------------------------------
# the three transforms
# --------------------------------------------------------------------------------------
def adjoint_finish(g, N, m, cols, real_output=False, prec="f32"):
    """grid [B, C, M..M] -> y [B, N..N, *cols]: FFT(+) (core_cuda.cu:267) and roll-off/crop (:298-326)."""
    B, C = g.shape[:2]
    d = g.ndim - 2
    M = 2 * N
    axes = tuple(range(2, 2 + d))
    ghat = np.fft.ifftn(g, axes=axes) * float(M ** d)  # sum_j g_j e^{+2 pi i jk/M}
    if prec == "f32":
        ghat = ghat.astype(np.complex64)
    bi = _band_index(N, M)
    sub = ghat[(slice(None), slice(None)) + np.ix_(*([bi] * d))]
    f = rolloff_factors(N, m, d, prec)
    y = sub * f[None, None]
    y = np.moveaxis(y, 1, -1)  # planar -> channels-last (core_cuda.cu:308)
    y = y.reshape((B,) + (N,) * d + tuple(cols))
    if real_output:
        return y.real.astype(np.float32 if prec == "f32" else np.float64)
    return y.astype(np.complex64 if prec == "f32" else np.complex128)


def nfft_adjoint(x, pos, batch=None, N=16, m=3, real_output=False, prec="f32"):
    """Restates nfft_adjoint_cuda (csrc/cuda/core_cuda.cu:144-336):
    spread -> unnormalised FFT with sign + (CUFFT_INVERSE, :267) -> crop/fftshift/deconvolve
    into y[B, N..N, *cols] (:298-326)."""
    pos, batch, B = _as_points(pos, batch)
    n, d = pos.shape
    x2, cols, C = _columns(x, n)
    g = spread(pos, x2, batch, B, N, m, prec)
    return adjoint_finish(g, N, m, cols, real_output, prec)


def forward_begin(xhat, d, m, prec="f32"):
    """xhat [B, N..N, *cols] -> complex grid [B, C, M..M]: roll-off + zero-pad (core_cuda.cu:403-420)
    and FFT(-) (:445)."""
    xhat = np.asarray(xhat)
    B, N = xhat.shape[0], xhat.shape[1]
    assert xhat.ndim >= d + 1 and all(s == N for s in xhat.shape[1:1 + d])  # core_cuda.cu:104-114
    cols = xhat.shape[1 + d:]
    C = int(np.prod(cols)) if len(cols) else 1
    M = 2 * N
    xh = np.moveaxis(xhat.reshape((B,) + (N,) * d + (C,)), -1, 1)  # [B, C, N..N]
    f = rolloff_factors(N, m, d, prec)
    cdt = np.complex64 if prec == "f32" else np.complex128
    vals = (xh.astype(cdt) * f[None, None]).astype(cdt)
    ghat = np.zeros((B, C) + (M,) * d, dtype=np.complex128)
    bi = _band_index(N, M)
    ghat[(slice(None), slice(None)) + np.ix_(*([bi] * d))] = vals
    g = np.fft.fftn(ghat, axes=tuple(range(2, 2 + d)))
    return g.astype(np.complex64) if prec == "f32" else g


def nfft_forward(xhat, pos, batch=None, m=3, real_output=False, prec="f32"):
    """Restates nfft_forward_cuda (core_cuda.cu:340-531): deconvolve + zero-pad into the
    oversampled grid (:403-420), unnormalised FFT with sign - (CUFFT_FORWARD, :445), gather."""
    pos, batch, B = _as_points(pos, batch)
    n, d = pos.shape
    xhat = np.asarray(xhat)
    assert xhat.ndim >= d + 1 and xhat.shape[0] == B  # core_cuda.cu:104-106
    N = xhat.shape[1]
    cols = xhat.shape[1 + d:]
    cdt = np.complex64 if prec == "f32" else np.complex128
    g = forward_begin(xhat, d, m, prec)
    y = gather(g, pos, batch, N, m, prec)
    y = y.reshape((n,) + tuple(cols))
    if real_output:
        return y.real.astype(np.float32 if prec == "f32" else np.float64)
    return y.astype(cdt)


def fastsum_middle(g, coeffs, m, prec="f32"):
    """grid -> grid: FFT(+), multiply in-band entries by (prod phi_hat_inv)^2 * coeffs[k+N/2] and zero
    the rest (spectral_window_operations.cu:292-331), FFT(-)  (core_cuda.cu:683-765)."""
    coeffs = np.asarray(coeffs)
    d = coeffs.ndim
    N = coeffs.shape[0]
    M = 2 * N
    axes = tuple(range(2, 2 + d))
    ghat = np.fft.ifftn(g, axes=axes) * float(M ** d)
    cdt = np.complex64 if prec == "f32" else np.complex128
    if prec == "f32":
        ghat = ghat.astype(cdt)
    f = rolloff_factors(N, m, d, prec)
    f2 = (f * f).astype(f.dtype)  # factor *= factor (spectral_window_operations.cu:326)
    bi = _band_index(N, M)
    sel = (slice(None), slice(None)) + np.ix_(*([bi] * d))
    out = np.zeros_like(ghat, dtype=np.complex128)
    out[sel] = (ghat[sel] * coeffs.astype(cdt)[None, None]).astype(cdt) * f2[None, None]
    g2 = np.fft.fftn(out, axes=axes)
    return g2.astype(cdt) if prec == "f32" else g2


def nfft_fastsum(x, coeffs, sources, targets=None, source_batch=None, target_batch=None,
                 m=3, prec="f32"):
    """Restates nfft_fastsum_cuda (core_cuda.cu:535-852): spread sources, fastsum_middle, gather at
    targets.  Output dtype follows x (real x -> real part, core_cuda.cu:814-818)."""
    if targets is None:
        targets, target_batch = sources, source_batch
    sources, source_batch, B = _as_points(sources, source_batch)
    targets, target_batch, Bt = _as_points(targets, target_batch)
    assert B == Bt and sources.shape[1] == targets.shape[1]
    n, d = sources.shape
    x2, cols, C = _columns(x, n)
    coeffs = np.asarray(coeffs)
    assert coeffs.ndim == d
    N = coeffs.shape[0]
    cdt = np.complex64 if prec == "f32" else np.complex128
    g = spread(sources, x2, source_batch, B, N, m, prec)
    g2 = fastsum_middle(g, coeffs, m, prec)
    y = gather(g2, targets, target_batch, N, m, prec).reshape((targets.shape[0],) + tuple(cols))
    if np.iscomplexobj(x):
        return y.astype(cdt)
    return y.real.astype(np.float32 if prec == "f32" else np.float64)


# --------------------------------------------------------------------------------------
# exact direct sums (accuracy oracle / CPU baseline port), float64
# --------------------------------------------------------------------------------------
def _freq_grid(N, d):
    k1 = np.arange(-N // 2, N // 2, dtype=np.float64)  # ndft.py:10
    return np.stack(np.meshgrid(*([k1] * d), indexing="ij"), axis=-1).reshape(-1, d)


def ndft_adjoint(x, pos, batch=None, N=16):
    """y[b, k+N/2, c] = sum_{i in b} x[i,c] e^{+2 pi i k.pos_i}  (torch_nfft/ndft.py:5-23)."""
    pos, batch, B = _as_points(pos, batch)
    n, d = pos.shape
    x2, cols, C = _columns(x, n)
    K = _freq_grid(N, d)
    out = np.zeros((B, K.shape[0], C), dtype=np.complex128)
    for b in range(B):
        sel = batch == b
        F = np.exp(2j * np.pi * (K @ pos[sel].astype(np.float64).T))
        out[b] = F @ x2[sel].astype(np.complex128)
    return out.reshape((B,) + (N,) * d + tuple(cols))


def ndft_forward(xhat, pos, batch=None):
    """y[i, c] = sum_k xhat[b_i, k+N/2, c] e^{-2 pi i k.pos_i}  (torch_nfft/ndft.py:26-44)."""
    pos, batch, B = _as_points(pos, batch)
    n, d = pos.shape
    xhat = np.asarray(xhat)
    N = xhat.shape[1]
    cols = xhat.shape[1 + d:]
    C = int(np.prod(cols)) if len(cols) else 1
    xh = xhat.reshape(B, N ** d, C).astype(np.complex128)
    K = _freq_grid(N, d)
    y = np.zeros((n, C), dtype=np.complex128)
    for b in range(B):
        sel = batch == b
        F = np.exp(-2j * np.pi * (pos[sel].astype(np.float64) @ K.T))
        y[sel] = F @ xh[b]
    return y.reshape((n,) + tuple(cols))


def ndft_fastsum(x, coeffs, sources, targets=None, source_batch=None, target_batch=None):
    """ndft_forward(coeffs * ndft_adjoint(x))  (torch_nfft/ndft.py:48-62)."""
    if targets is None:
        targets, target_batch = sources, source_batch
    coeffs = np.asarray(coeffs)
    N = coeffs.shape[0]
    x = np.asarray(x)
    xx = x if x.ndim > 1 else x[:, None]
    y = ndft_adjoint(xx, sources, source_batch, N=N)
    y = y * coeffs.reshape((1,) + coeffs.shape + (1,) * (xx.ndim - 1))
    y = ndft_forward(y, targets, target_batch)
    y = y.reshape((np.asarray(targets).shape[0],) + x.shape[1:])
    return y if np.iscomplexobj(x) else y.real


# --------------------------------------------------------------------------------------
# coefficient helpers (setup only)
# --------------------------------------------------------------------------------------
def gaussian_analytic_coeffs(sigma, dim=3, N=16):
    """b_l = prod_a sqrt(pi) sigma exp(-sigma^2 pi^2 l_a^2), stored at l+N/2
    (csrc/cuda/kernel_coeffs.cu:6-30)."""
    l = np.arange(N, dtype=np.float64) - N // 2
    v = np.sqrt(np.pi) * sigma * np.exp(-(sigma ** 2) * (np.pi ** 2) * l * l)
    out = np.ones((N,) * dim)
    for a in range(dim):
        shape = [1] * dim
        shape[a] = N
        out = out * v.reshape(shape)
    return out.astype(np.float32)


def interpolation_grid(dim=3, N=16):
    """grid[i_0..i_{d-1}, a] = i_a/N - 1/2  (kernel_coeffs.cu:76-97)."""
    g1 = np.arange(N, dtype=np.float32) / np.float32(N) - np.float32(0.5)
    return np.stack(np.meshgrid(*([g1] * dim), indexing="ij"), axis=-1).astype(np.float32)


def radial_interpolation_grid(dim=3, N=16):
    """Euclidean norm of interpolation_grid  (kernel_coeffs.cu:99-123)."""
    g = interpolation_grid(dim, N).astype(np.float64)
    return np.sqrt((g * g).sum(-1)).astype(np.float32)


def interpolated_kernel_coeffs(values):
    """fftshift(fftn(ifftshift(values))) / N^d  (kernel_coeffs.cu:126-202, core_cuda.cu:995-1064)."""
    v = np.asarray(values)
    out = np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(v.astype(np.complex128)))) / v.size
    return out.astype(np.complex64)


def gaussian_interpolated_coeffs(sigma, dim=3, N=16, p=-1, eps=0.0):
    """Samples exp(-r^2/sigma^2) on i/N - 1/2 (clamped to r = 1/2 outside the ball when p >= 0)
    and interpolates  (kernel_coeffs.cu:33-73; p <= 0 and eps == 0 only, core_cuda.cu:890-891)."""
    assert p <= 0 and eps == 0.0
    g = interpolation_grid(dim, N).astype(np.float64)
    r2 = (g * g).sum(-1)
    vals = np.exp(-r2 / sigma ** 2)
    if p >= 0:
        vals = np.where(r2 <= 0.25, vals, np.exp(-0.25 / sigma ** 2))
    return interpolated_kernel_coeffs(vals)


# --------------------------------------------------------------------------------------
# binning oracle for the engine's deterministic counting sort (integer work: bit-exact)
# --------------------------------------------------------------------------------------
def tile_keys(pos, batch, N, tile):
    """key = ((b*nt_0 + t_0)*nt_1 + t_1)... with t_a = (floorf(pos_a*M) mod M) // tile_a.
    The cell rule is the reference's (spatial_window_operations.cu:50); the tiling is the
    engine's own."""
    pos, batch, B = _as_points(pos, batch)
    d = pos.shape[1]
    M = 2 * N
    cw = np.mod(compute_cells(pos, M), M)
    key = batch.copy()
    for a in range(d):
        nt = -(-M // tile[a])
        key = key * nt + cw[:, a] // tile[a]
    return key.astype(np.int64)


def sort_keys(pos, batch, N, tile, fine_bits=0, supercell=(2, 4, 4)):
    """The engine's full sort key: tile key << fine_bits | top fine_bits bits of the hierarchical index of the
    point's supercell inside its 16^3 tile -- per level from halves down to single supercells a (y bit, x bit)
    pair, then the z supercell (the engine's own binning rule, torch_nfft_b200/csrc/sort.cuh: fine_index; 3D only).
    `tile` / `supercell` are in API dimension order (dim 0 = slot Z ... dim 2 = slot X)."""
    key = tile_keys(pos, batch, N, tile)
    if fine_bits == 0:
        return key
    pos32, _, _ = _as_points(pos, batch)
    assert pos32.shape[1] == 3
    M = 2 * N
    cw = np.mod(compute_cells(pos32, M), M)
    inside = cw - (cw // np.asarray(tile)) * np.asarray(tile)
    bz, by, bx = (inside[:, a] // supercell[a] for a in range(3))
    levels = int(np.log2(tile[2] // supercell[2]))       # supercells per tile edge in x (= y): 2^levels
    zbits = int(np.log2(tile[0] // supercell[0]))
    fine = np.zeros_like(key)
    for b in range(levels - 1, -1, -1):
        fine = (fine << 2) | (((by >> b) & 1) << 1) | ((bx >> b) & 1)
    fine = (fine << zbits) | bz
    total = 2 * levels + zbits
    if fine_bits >= total:  # mixed-density keys: the index left-aligned in a wider field
        return ((key << fine_bits) | (fine << (fine_bits - total))).astype(np.int64)
    return ((key << fine_bits) | (fine >> (total - fine_bits))).astype(np.int64)


def stable_permutation(keys):
    """Stable sort permutation: what the engine's counting sort must reproduce bit-for-bit."""
    return np.argsort(keys, kind="stable").astype(np.int64)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))
