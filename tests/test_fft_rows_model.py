"""CPU model of the hand-written FFT row pass (torch_nfft_b200/csrc/fft_rows.cuh): the same index maps in numpy.

Not the kernel itself (that is compared with the cuFFT path and the reference on the GPU,
tests/test_parity_gpu.py::test_pruned_fft_matches_cufft_path) -- this pins the decomposition the kernel
implements: a pair of real rows packed as one complex sequence, M = 16 * R2, x = R2 x1 + x2, k = k1 + 16 k2,
a radix-2 decimation-in-time FFT per stage, only the frequencies k <= M/4 kept (adjoint) or given (forward)."""
import numpy as np
import pytest


def bit_reverse(i, bits):
    r = 0
    for b in range(bits):
        r |= ((i >> b) & 1) << (bits - 1 - b)
    return r


def fft_reg(v, sign):
    """fft_reg<R, SIGN>: in-place radix-2 DIT, v[k] <- sum_n v[n] exp(sign 2 pi i n k / R)."""
    v = list(v)
    R = len(v)
    log = R.bit_length() - 1
    for i in range(R):
        j = bit_reverse(i, log)
        if j > i:
            v[i], v[j] = v[j], v[i]
    for s in range(1, log + 1):
        m, h = 1 << s, 1 << (s - 1)
        for k in range(0, R, m):
            for j in range(h):
                t = np.exp(sign * 2j * np.pi * (j * (32 // m)) / 32) * v[k + j + h]
                a = v[k + j]
                v[k + j], v[k + j + h] = a + t, a - t
    return v


def rows_r2c_crop(ga, gb, R2):
    """rows_r2c_crop_kernel for one pair of rows: the kept half spectra X_a[0..M/4], X_b[0..M/4]."""
    M, KX = 16 * R2, 16 * R2 // 4 + 1
    tw = np.exp(-2j * np.pi * np.arange(M) / M)
    stage = np.zeros((16, R2), complex)
    for x2 in range(R2):  # stage A: thread <-> x2
        v = fft_reg([complex(ga[R2 * x1 + x2], gb[R2 * x1 + x2]) for x1 in range(16)], -1)
        for k1 in range(16):
            stage[k1, x2] = v[k1] * tw[x2 * k1]
    Z = {}
    for t in range(16):  # stage B: thread <-> k1
        u = fft_reg(list(stage[t]), -1)
        for k2 in range(R2):
            k = t + 16 * k2
            if (k2 <= R2 // 4 and k < KX) or k2 >= 3 * R2 // 4:
                Z[k] = u[k2]
    xa, xb = np.zeros(KX, complex), np.zeros(KX, complex)
    for kx in range(KX):
        zk, zm = Z[kx], Z[kx] if kx == 0 else Z[M - kx]
        xa[kx] = complex(0.5 * (zk.real + zm.real), 0.5 * (zk.imag - zm.imag))
        xb[kx] = complex(0.5 * (zk.imag + zm.imag), -0.5 * (zk.real - zm.real))
    return xa, xb


def rows_c2r_pad(Za, Zb, R2):
    """rows_c2r_pad_kernel for one pair of rows: the two real rows of length M."""
    M, KX = 16 * R2, 16 * R2 // 4 + 1
    tw = np.exp(2j * np.pi * np.arange(M) / M)
    stage = np.zeros((16, R2), complex)
    for t in range(16):  # stage B': thread <-> k1
        u = [0j] * R2
        for k2 in range(R2):
            k = t + 16 * k2
            if k2 <= R2 // 4 and k < KX:
                a, b = Za[k], Zb[k]
                u[k2] = complex(a.real, b.real) if k == 0 else complex(a.real - b.imag, a.imag + b.real)
            elif k2 >= 3 * R2 // 4:
                a, b = Za[M - k], Zb[M - k]
                u[k2] = complex(a.real + b.imag, b.real - a.imag)
        u = fft_reg(u, 1)
        for x2 in range(R2):
            stage[t, x2] = u[x2] * tw[t * x2]
    ga, gb = np.zeros(M), np.zeros(M)
    for x2 in range(R2):  # stage A': thread <-> x2
        v = fft_reg(list(stage[:, x2]), 1)
        for x1 in range(16):
            ga[R2 * x1 + x2], gb[R2 * x1 + x2] = v[x1].real, v[x1].imag
    return ga, gb


@pytest.mark.parametrize("R2", [16, 32])
def test_row_pass_model_equals_numpy_fft(R2):
    M, KX = 16 * R2, 16 * R2 // 4 + 1
    rng = np.random.default_rng(R2)
    ga, gb = rng.standard_normal(M), rng.standard_normal(M)
    xa, xb = rows_r2c_crop(ga, gb, R2)
    assert np.abs(xa - np.fft.rfft(ga)[:KX]).max() < 1e-11 and np.abs(xb - np.fft.rfft(gb)[:KX]).max() < 1e-11
    Za = rng.standard_normal(KX) + 1j * rng.standard_normal(KX)
    Zb = rng.standard_normal(KX) + 1j * rng.standard_normal(KX)
    ra, rb = rows_c2r_pad(Za, Zb, R2)
    for Z, r in ((Za, ra), (Zb, rb)):
        full = np.zeros(M // 2 + 1, complex)
        full[:KX] = Z  # C2R semantics (unnormalised, Im Z[0] ignored): numpy's irfft times M
        assert np.abs(r - np.fft.irfft(full, n=M) * M).max() < 1e-10
