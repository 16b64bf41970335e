// Pruned real FFTs along the fastest grid axis (X), for the real (half-spectrum) transforms on grids with
// M = 256 or 512 cells per dimension.
//
// An NFFT keeps only the N = M/2 central frequencies of the oversampled grid's spectrum (adjoint) or starts from
// them (forward): 1/2^d of what a full R2C / C2R transform writes or reads.  cuFFT cannot crop on write, so its 3D
// transform moves the whole half spectrum (B C M^(d-1) (M/2+1) complex numbers) two or three times.  Here the X
// pass is hand-written and fused with the crop: it reads a real row of M cells and writes only the M/4 + 1
// frequencies kx = 0 .. N/2 that survive (adjoint), resp. reads those and writes the real row (forward).  The
// remaining (d-1)-dimensional complex transforms (cuFFT C2C, batched, contiguous) then work on
//     P[bc][kx][z][y]          (bc = batch entry x channel; kx outermost so that every (bc, kx) plane is contiguous)
// which is a quarter of the half spectrum.  At c4 the spectral stages move 1.15 GB instead of 1.95 GB per
// transform.  (Replaces the X pass of the reference's cufftPlanMany C2C transforms, core_cuda.cu:254-272, 432-445.)
//
// A pair of rows (packed as one complex sequence) is transformed by 16 threads in two stages, M = 16 * R2 (R2 = 16 or 32), x = R2 x1 + x2, k = k1 + 16 k2:
//     X[k1 + 16 k2] = sum_x2 w_R2^(x2 k2) * w_M^(x2 k1) * [ sum_x1 g[R2 x1 + x2] w_16^(x1 k1) ]
//   stage A: thread <-> x2, a 16-point DFT over x1 in registers, twiddle, transposed through shared memory
//   stage B: thread <-> k1, an R2-point DFT over x2 in registers; only k2 <= R2/4 is kept (k <= M/4)
// and the inverse runs the same two stages in the opposite order with conjugated twiddles.
// tests/test_fft_rows_model.py restates these index maps in numpy (CPU); the kernels themselves are compared with
// the cuFFT path in tests/test_parity_gpu.py::test_pruned_fft_matches_cufft_path.
#pragma once
#include "common.cuh"

namespace nfftb200 {

constexpr int kFftPairsPerCta = 16;                  // row pairs per CTA (16 threads each)
constexpr int kFftRowsPerCta = 2 * kFftPairsPerCta;  // rows per CTA
constexpr int kFftThreads = 16 * kFftPairsPerCta;

__host__ __device__ constexpr int bit_reverse(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }

// compile-time loop: f(IntC<I>{}) for I = BEGIN .. END-1 (indices into register arrays must be constants)
template <int K>
struct FftIdx { static constexpr int value = K; };
template <int BEGIN, int END, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (BEGIN < END) {
        f(FftIdx<BEGIN>{});
        static_for<BEGIN + 1, END>(f);
    }
}

// In-register radix-2 FFT of R = 8, 16 or 32 complex values (every index is a compile-time constant, the twiddles
// are immediates): v[k] <- sum_n v[n] exp(SIGN 2 pi i n k / R).
template <int R, int SIGN>
__device__ __forceinline__ void fft_reg(float2 (&v)[R]) {
    constexpr int LOG = ilog2(R);
    // bit-reversal permutation (register renaming)
    static_for<0, R>([&](auto ic) {
        constexpr int i = decltype(ic)::value, j = bit_reverse(i, LOG);
        if constexpr (j > i) {
            const float2 t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    });
    static_for<1, LOG + 1>([&](auto sc) {
        constexpr int m = 1 << decltype(sc)::value, h = m >> 1;
        static_for<0, R / m>([&](auto kc) {
            constexpr int k = decltype(kc)::value * m;
            static_for<0, h>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr int tj = j * (32 / m);  // w_m^j = w_32^(j 32/m), tj < 16
                // cos / sin (2 pi tj / 32)
                constexpr float kCos32[16] = {1.f,           0.980785251f,  0.923879504f,  0.831469595f,
                                              0.707106769f,  0.555570245f,  0.382683426f,  0.195090324f,
                                              0.f,           -0.195090324f, -0.382683426f, -0.555570245f,
                                              -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f};
                constexpr float kSin32[16] = {0.f,          0.195090324f, 0.382683426f, 0.555570245f,
                                              0.707106769f, 0.831469595f, 0.923879504f, 0.980785251f,
                                              1.f,          0.980785251f, 0.923879504f, 0.831469595f,
                                              0.707106769f, 0.555570245f, 0.382683426f, 0.195090324f};
                const float2 b = v[k + j + h];
                float2 t;
                if constexpr (tj == 0) {
                    t = b;
                } else if constexpr (tj == 8) {  // multiplication by SIGN * i
                    t = SIGN > 0 ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
                } else {
                    constexpr float c = kCos32[tj], sn = SIGN > 0 ? kSin32[tj] : -kSin32[tj];
                    t = make_float2(b.x * c - b.y * sn, b.x * sn + b.y * c);
                }
                const float2 a = v[k + j];
                v[k + j] = make_float2(a.x + t.x, a.y + t.y);
                v[k + j + h] = make_float2(a.x - t.x, a.y - t.y);
            });
        });
    });
}

// Two real rows share one complex transform (z = g_a + i g_b): 16 threads transform a PAIR of rows, a CTA 16 pairs.
// shared memory of the row kernels: twiddles w_M^j | stage buffer [pair][k1][R2 + 1]; the same buffer also stages
// the spectra: Z of a pair at the kept frequencies [pair][M/2 + 2] (R2C), resp. the input planes [kx][32 + 1] (C2R).
template <int R2>
constexpr size_t fft_rows_smem_bytes() {
    return (size_t)(16 * R2) * sizeof(float2) + (size_t)kFftPairsPerCta * 16 * (R2 + 1) * sizeof(float2);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// grid [rows][M] real  ->  P[bc][kx][row in bc], kx = 0 .. M/4, = the R2C (sign -) transform of every row, cropped.
// rows_per_bc = M^(d-1) (a multiple of 32: a CTA's 32 rows belong to one bc and are consecutive in P).
// With Z = FFT(g_a + i g_b):  X_a[k] = (Z[k] + conj Z[M-k]) / 2,  X_b[k] = (Z[k] - conj Z[M-k]) / (2i).
template <int R2>
__global__ void __launch_bounds__(kFftThreads)
rows_r2c_crop_kernel(const float* __restrict__ grid, float2* __restrict__ P, long long rows_per_bc) {
    constexpr int M = 16 * R2, KX = M / 4 + 1, XPT = R2 / 16;  // XPT: x2 values per thread in stage A
    constexpr int ZS = M / 2 + 2;                               // kept Z entries per pair: k <= M/4 | k >= 3M/4 (+ pad)
    extern __shared__ __align__(16) float2 fsm[];
    float2* tw = fsm;           // tw[j] = exp(-2 pi i j / M)
    float2* stage = fsm + M;    // [pair][k1][R2 + 1]
    const int tid = threadIdx.x, pair = tid >> 4, t = tid & 15;
    for (int j = tid; j < M; j += kFftThreads) {
        float sn, cs;
        sincospif((float)(2 * j) / (float)M, &sn, &cs);  // the argument is exact (M is a power of two)
        tw[j] = make_float2(cs, -sn);
    }
    const long long row0 = (long long)blockIdx.x * (2 * kFftPairsPerCta);
    const float* src = grid + (row0 + 2 * pair) * M;
    float2 v[XPT][16];
#pragma unroll
    for (int e = 0; e < XPT; ++e)
#pragma unroll
        for (int x1 = 0; x1 < 16; ++x1) v[e][x1] = make_float2(src[R2 * x1 + t + 16 * e], src[M + R2 * x1 + t + 16 * e]);
    __syncthreads();  // twiddles
    float2* srow = stage + (size_t)pair * 16 * (R2 + 1);
#pragma unroll
    for (int e = 0; e < XPT; ++e) {
        const int x2 = t + 16 * e;
        fft_reg<16, -1>(v[e]);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) srow[k1 * (R2 + 1) + x2] = k1 == 0 ? v[e][0] : cmul(v[e][k1], tw[x2 * k1]);
    }
    __syncthreads();
    float2 u[R2];
#pragma unroll
    for (int x2 = 0; x2 < R2; ++x2) u[x2] = srow[t * (R2 + 1) + x2];
    fft_reg<R2, -1>(u);
    __syncthreads();  // everybody has read the stage buffer: reuse it for the kept Z values, zs[pair][ZS]
    float2* zs = stage;
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
        const int k = t + 16 * k2;  // u[k2] = Z[k]
        if (k2 <= R2 / 4) {
            if (k < KX) zs[pair * ZS + k] = u[k2];
        } else if (k2 >= 3 * R2 / 4) {
            zs[pair * ZS + KX + (k - 3 * M / 4)] = u[k2];  // k >= 3M/4
        }
    }
    __syncthreads();
    const long long bc = row0 / rows_per_bc, r_in = row0 - bc * rows_per_bc;
    float2* dst = P + (bc * KX) * rows_per_bc + r_in;
    for (int i = tid; i < KX * 2 * kFftPairsPerCta; i += kFftThreads) {
        const int kx = i >> 5, r = i & 31, pr = r >> 1;
        const float2 zk = zs[pr * ZS + kx];
        const float2 zm = kx == 0 ? zk : zs[pr * ZS + KX + (M / 4 - kx)];  // Z[M - kx]
        float2 o;
        if ((r & 1) == 0) o = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        else o = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
        dst[(long long)kx * rows_per_bc + r] = o;
    }
}

// P[bc][kx][row in bc] (kx = 0 .. M/4; all higher frequencies are zero)  ->  grid [rows][M] real, = the C2R (sign +)
// transform of every row: g[x] = Re Z[0] + 2 Re sum_{k >= 1} Z[k] exp(+2 pi i k x / M).
// Two rows per transform: F[k] = Z_a[k] + i Z_b[k] (k <= M/4), F[M-k] = conj Z_a[k] + i conj Z_b[k]; then
// g_a + i g_b = sum_k F[k] exp(+2 pi i k x / M).  Im Z[0] is ignored, as cuFFT's C2R does.
template <int R2>
__global__ void __launch_bounds__(kFftThreads)
rows_c2r_pad_kernel(const float2* __restrict__ P, float* __restrict__ grid, long long rows_per_bc) {
    constexpr int M = 16 * R2, KX = M / 4 + 1, XPT = R2 / 16;
    constexpr int NR = 2 * kFftPairsPerCta, NRP = NR + 1;  // rows per CTA, padded pitch of the input staging
    extern __shared__ __align__(16) float2 fsm[];
    float2* tw = fsm;           // tw[j] = exp(+2 pi i j / M)
    float2* stage = fsm + M;
    const int tid = threadIdx.x, pair = tid >> 4, t = tid & 15;
    for (int j = tid; j < M; j += kFftThreads) {
        float sn, cs;
        sincospif((float)(2 * j) / (float)M, &sn, &cs);
        tw[j] = make_float2(cs, sn);
    }
    const long long row0 = (long long)blockIdx.x * NR;
    const long long bc = row0 / rows_per_bc, r_in = row0 - bc * rows_per_bc;
    const float2* src = P + (bc * KX) * rows_per_bc + r_in;
    float2* in = stage;  // [kx][NRP]
    for (int i = tid; i < KX * NR; i += kFftThreads) {
        const int kx = i >> 5, r = i & 31;
        in[kx * NRP + r] = src[(long long)kx * rows_per_bc + r];
    }
    __syncthreads();
    // stage B': thread <-> k1, R2-point inverse DFT over k2 of F
    float2 u[R2];
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
        const int k = t + 16 * k2;
        u[k2] = make_float2(0.f, 0.f);
        if (k2 <= R2 / 4) {
            if (k < KX) {
                const float2 a = in[k * NRP + 2 * pair], b = in[k * NRP + 2 * pair + 1];
                u[k2] = k == 0 ? make_float2(a.x, b.x) : make_float2(a.x - b.y, a.y + b.x);
            }
        } else if (k2 >= 3 * R2 / 4) {
            const int kk = M - k;  // 1 .. M/4
            const float2 a = in[kk * NRP + 2 * pair], b = in[kk * NRP + 2 * pair + 1];
            u[k2] = make_float2(a.x + b.y, b.x - a.y);
        }
    }
    fft_reg<R2, 1>(u);
    __syncthreads();  // the input staging buffer is reused as the stage buffer
    float2* srow = stage + (size_t)pair * 16 * (R2 + 1);
#pragma unroll
    for (int x2 = 0; x2 < R2; ++x2) srow[t * (R2 + 1) + x2] = t == 0 ? u[x2] : cmul(u[x2], tw[t * x2]);
    __syncthreads();
    // stage A': thread <-> x2, 16-point inverse DFT over k1
    float* dst = grid + (row0 + 2 * pair) * M;
#pragma unroll
    for (int e = 0; e < XPT; ++e) {
        const int x2 = t + 16 * e;
        float2 v[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) v[k1] = srow[k1 * (R2 + 1) + x2];
        fft_reg<16, 1>(v);
#pragma unroll
        for (int x1 = 0; x1 < 16; ++x1) {
            dst[R2 * x1 + x2] = v[x1].x;
            dst[M + R2 * x1 + x2] = v[x1].y;
        }
    }
}

}  // namespace nfftb200
