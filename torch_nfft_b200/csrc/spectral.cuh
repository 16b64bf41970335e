// Spectral kernels: roll-off correction (deconvolution by the window's Fourier coefficients),
// cropping / zero-padding, fftshift and the planar <-> channels-last transposition, fused into one
// pack (pre-FFT) or unpack (post-FFT) pass over the data; plus the fastsum kernel multiply.
//
// Replaces compute_phi_hat_inv_kernel, {real,complex}_adjoint_rolloff_correction_kernel,
// {real,complex}_forward_rolloff_correction_kernel and {real,complex}_kernel_convolution_kernel
// (reference csrc/cuda/spectral_window_operations.cu:27-43, 51-153, 158-265, 269-402).
//
// Real data takes the half-spectrum path: the oversampled grid is real, cuFFT R2C/C2R replaces
// the reference's C2C (core_cuda.cu:254-267, 432-445).  cuFFT's R2C has sign -, C2R sign +, the
// adjoint needs + and the forward -, hence the conjugations below:
//   adjoint : ghat+[k] = conj(R[k])          (k_last >= 0),   = R[-k]   (k_last < 0),  R = R2C(g)
//   forward : Re(FFT-(ghat)) = C2R(in),  in[k] = 1/2 (conj(ghat[k]) + ghat[-k])   (Hermitian part)
// phi_hat_inv(k) = expf(float(k*k) * c_hat) is evaluated in-kernel with the reference's
// expression (spectral_window_operations.cu:14-18) instead of a table that is re-allocated and
// re-computed on every call (core_cuda.cu:283-290).
#pragma once
#include "common.cuh"

namespace nfftb200 {

__device__ __forceinline__ float phi_hat_inv(int k, float c_hat) { return expf((float)(k * k) * c_hat); }

__device__ __forceinline__ int wrapM(int k, int M) { return k < 0 ? k + M : k; }

// ---------------------------------------------------------------------------------------
// adjoint unpack: spectrum (planar) -> y[b, i_0..i_{d-1}, c]
//   HALF: spec is the R2C half spectrum [BC][M]..[M][M/2+1];  else full C2C (+ sign) [BC][M]^d
// ---------------------------------------------------------------------------------------
// I = int when every index fits 31 bits (64-bit divisions are ~10x dearer), long long otherwise
// PRUNED (with HALF): spec is P[BC][N/2+1][M]..[M] -- the kept kx outermost, see fft_rows.cuh
template <int DIM, bool HALF, bool REAL_OUT, typename I, bool PRUNED = false>
__global__ void __launch_bounds__(256)
unpack_kernel(const float2* __restrict__ spec, float* __restrict__ y, Geom g) {
    const I total = (I)g.B * g.C;
    I nd = 1;
    for (int a = 0; a < DIM; ++a) nd *= g.N;
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total * nd) return;
    // index decode: shifts when N and C are powers of two (the kernel is instruction-bound otherwise: five integer
    // divisions per output, ncu: issue slots 79 % busy, profiles/r03c_misc_kernels.txt)
    int c;
    I f;
    int k[3] = {0, 0, 0};
    if (((g.N & (g.N - 1)) | (g.C & (g.C - 1))) == 0) {
        const int cs = __ffs(g.C) - 1, ns = __ffs(g.N) - 1;
        c = (int)(idx & (I)(g.C - 1));
        f = idx >> cs;
#pragma unroll
        for (int a = DIM - 1; a >= 0; --a) {
            k[a] = (int)(f & (I)(g.N - 1)) - g.N / 2;
            f >>= ns;
        }
    } else {
        c = (int)(idx % g.C);
        f = idx / g.C;
#pragma unroll
        for (int a = DIM - 1; a >= 0; --a) {
            k[a] = (int)(f % g.N) - g.N / 2;
            f /= g.N;
        }
    }
    const I bc = f * g.C + c;
    float factor = 1.0f;
#pragma unroll
    for (int a = 0; a < DIM; ++a) factor *= phi_hat_inv(k[a] < 0 ? -k[a] : k[a], g.c_hat);

    float2 v;
    if (HALF) {
        const int H = g.M / 2 + 1;
        const bool neg = k[DIM - 1] < 0;
        I s = bc;
        if (PRUNED) s = s * (g.N / 2 + 1) + (neg ? -k[DIM - 1] : k[DIM - 1]);
#pragma unroll
        for (int a = 0; a < DIM - 1; ++a) s = s * g.M + wrapM(neg ? -k[a] : k[a], g.M);
        if (!PRUNED) s = s * H + (neg ? -k[DIM - 1] : k[DIM - 1]);
        v = spec[s];
        if (!neg) v.y = -v.y;
    } else {
        I s = bc;
#pragma unroll
        for (int a = 0; a < DIM; ++a) s = s * g.M + wrapM(k[a], g.M);
        v = spec[s];
    }
    if (REAL_OUT) {
        y[idx] = v.x * factor;
    } else {
        reinterpret_cast<float2*>(y)[idx] = make_float2(v.x * factor, v.y * factor);
    }
}

// ---------------------------------------------------------------------------------------
// forward pack: xhat[b, i.., c] -> spectrum (planar), deconvolved and zero padded
// ---------------------------------------------------------------------------------------
// decode a planar spectrum index into (bc, signed frequencies); returns false if idx out of range
template <int DIM, bool HALF, typename I>
__device__ __forceinline__ bool decode_spec(I idx, const Geom& g, I& bc, int kap[3]) {
    const int H = g.M / 2 + 1;
    const int last = HALF ? H : g.M;
    I per = last;
    for (int a = 0; a < DIM - 1; ++a) per *= g.M;
    if (idx >= per * g.B * g.C) return false;
    I r = idx;
    int j = (int)(r % last);
    r /= last;
    kap[DIM - 1] = (HALF || j < g.M / 2) ? j : j - g.M;
    if (!HALF && j == g.M / 2) kap[DIM - 1] = g.M / 2;
#pragma unroll
    for (int a = DIM - 2; a >= 0; --a) {
        j = (int)(r % g.M);
        r /= g.M;
        kap[a] = j < g.M / 2 ? j : j - g.M;
        if (j == g.M / 2) kap[a] = g.M / 2;  // +-N: out of band either way
    }
    bc = r;
    return true;
}

template <int DIM>
__device__ __forceinline__ bool in_band(const int kap[3], int sign, int N) {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        const int k = sign * kap[a];
        ok = ok && (k >= -N / 2) && (k <= N / 2 - 1);
    }
    return ok;
}

// index into channels-last [B][N]^d[C] for signed frequency sign*kap
template <int DIM, typename I>
__device__ __forceinline__ I api_index(const int kap[3], int sign, I b, int c, const Geom& g) {
    I s = b;
#pragma unroll
    for (int a = 0; a < DIM; ++a) s = s * g.N + (sign * kap[a] + g.N / 2);
    return s * g.C + c;
}

template <int DIM>
__device__ __forceinline__ float rolloff(const int kap[3], float c_hat) {
    float factor = 1.0f;
#pragma unroll
    for (int a = 0; a < DIM; ++a) factor *= phi_hat_inv(kap[a] < 0 ? -kap[a] : kap[a], c_hat);
    return factor;
}

// HALF = true : input of the C2R transform (Hermitian part, conjugated);  false: input of C2C(-).
// The spectrum is zero outside the band, so the host zero-fills it with one memset and this kernel
// visits only the band box: kappa_a in [-N/2, N/2] (HALF: last dim [0, N/2]) resp. [-N/2, N/2-1].
template <int DIM, bool HALF, bool XREAL, typename I, bool PRUNED = false>
__global__ void __launch_bounds__(256)
pack_kernel(const float* __restrict__ xhat, float2* __restrict__ spec, Geom g) {
    const int ext = HALF ? g.N + 1 : g.N;            // band extent of the leading dims
    const int last = HALF ? g.N / 2 + 1 : g.N;       // band extent of the last dim
    I per = last;
    for (int a = 0; a < DIM - 1; ++a) per *= ext;
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per * g.B * g.C) return;
    int kap[3] = {0, 0, 0};
    I r = idx;
    kap[DIM - 1] = (int)(r % last) - (HALF ? 0 : g.N / 2);
    r /= last;
#pragma unroll
    for (int a = DIM - 2; a >= 0; --a) {
        kap[a] = (int)(r % ext) - g.N / 2;
        r /= ext;
    }
    const I bc = r;
    const I b = bc / g.C;
    const int c = (int)(bc % g.C);
    // position in the planar spectrum
    I dst = bc;
    if (PRUNED) dst = dst * (g.N / 2 + 1) + kap[DIM - 1];  // P[bc][kx][..] (HALF only)
#pragma unroll
    for (int a = 0; a < DIM - 1; ++a) dst = dst * g.M + wrapM(kap[a], g.M);
    if (!PRUNED) dst = dst * (HALF ? g.M / 2 + 1 : g.M) + (HALF ? kap[DIM - 1] : wrapM(kap[DIM - 1], g.M));

    const bool pos_in = in_band<DIM>(kap, 1, g.N);
    float2 out = make_float2(0.f, 0.f);
    if (HALF) {
        const bool neg_in = in_band<DIM>(kap, -1, g.N);
        if (pos_in || neg_in) {
            const float factor = rolloff<DIM>(kap, g.c_hat);
            float re = 0.f, im = 0.f;
            if (pos_in) {  // conj(xhat[k])
                const I s = api_index<DIM, I>(kap, 1, b, c, g);
                if (XREAL) {
                    re += xhat[s];
                } else {
                    const float2 v = reinterpret_cast<const float2*>(xhat)[s];
                    re += v.x;
                    im -= v.y;
                }
            }
            if (neg_in) {  // xhat[-k]
                const I s = api_index<DIM, I>(kap, -1, b, c, g);
                if (XREAL) {
                    re += xhat[s];
                } else {
                    const float2 v = reinterpret_cast<const float2*>(xhat)[s];
                    re += v.x;
                    im += v.y;
                }
            }
            out = make_float2(0.5f * (re * factor), 0.5f * (im * factor));
        }
    } else if (pos_in) {
        const float factor = rolloff<DIM>(kap, g.c_hat);
        const I s = api_index<DIM, I>(kap, 1, b, c, g);
        if (XREAL) {
            out = make_float2(xhat[s] * factor, 0.f);
        } else {
            const float2 v = reinterpret_cast<const float2*>(xhat)[s];
            out = make_float2(v.x * factor, v.y * factor);
        }
    }
    spec[dst] = out;
}

// ---------------------------------------------------------------------------------------
// fastsum: multiply the spectrum of the spread sources by the kernel coefficients, in place.
//   full (C2C, spectrum computed with sign +): G[k] *= f(k)^2 * b_k in band, 0 elsewhere
//     (spectral_window_operations.cu:292-331).
//   half (R = R2C(g), output feeds C2R):  in[k] = R[k] * f(k)^2 * 1/2 (conj(b_k)[k in band] + b_{-k}[-k in band])
// ---------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ long long coeff_index(const int kap[3], int sign, int N) {
    long long s = 0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) s = s * N + (sign * kap[a] + N / 2);
    return s;
}

template <int DIM, bool HALF, bool CREAL, typename I>
__global__ void __launch_bounds__(256)
kernel_multiply_kernel(float2* __restrict__ spec, const float* __restrict__ coeffs, Geom g) {
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    I bc;
    int kap[3] = {0, 0, 0};
    if (!decode_spec<DIM, HALF, I>(idx, g, bc, kap)) return;
    const bool pos_in = in_band<DIM>(kap, 1, g.N);
    const bool neg_in = HALF && in_band<DIM>(kap, -1, g.N);
    if (!pos_in && !neg_in) {
        spec[idx] = make_float2(0.f, 0.f);
        return;
    }
    float factor = rolloff<DIM>(kap, g.c_hat);
    factor *= factor;  // spectral_window_operations.cu:326
    float2 beta = make_float2(0.f, 0.f);
    if (pos_in) {
        const long long s = coeff_index<DIM>(kap, 1, g.N);
        if (CREAL) {
            beta.x += coeffs[s];
        } else {
            const float2 c = reinterpret_cast<const float2*>(coeffs)[s];
            beta.x += c.x;
            beta.y += HALF ? -c.y : c.y;
        }
    }
    if (neg_in) {
        const long long s = coeff_index<DIM>(kap, -1, g.N);
        if (CREAL) {
            beta.x += coeffs[s];
        } else {
            const float2 c = reinterpret_cast<const float2*>(coeffs)[s];
            beta.x += c.x;
            beta.y += c.y;
        }
    }
    if (HALF) {
        beta.x *= 0.5f;
        beta.y *= 0.5f;
    }
    const float2 v = spec[idx];
    spec[idx] = make_float2(factor * (v.x * beta.x - v.y * beta.y), factor * (v.x * beta.y + v.y * beta.x));
}

}  // namespace nfftb200
