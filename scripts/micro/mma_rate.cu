// Microbenchmark: what the legacy tensor path (mma.sync.m16n8k8 tf32, SASS HMMA.1688.F32.TF32) delivers on
// sm_100a, and whether a 3xTF32 split product (hi*hi + lo*hi + hi*lo) reproduces fp32 products closely enough
// for the spread's 1e-5 parity budget.
//   mode 0: pure MMA issue rate, 24 independent accumulator tiles x 3 dependent MMAs per round
//   mode 1: one "k-step" of a tensor-core spread per round: fragments of 8 points' tap windows from shared
//           memory, B = psi_y * psi_z split into hi + lo, 24 tiles x 3 MMAs
//   accuracy: D = A * B for random fp32 A (16 x 8), B (8 x 8), against a double product
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// hi = the value with its low 13 mantissa bits cleared (what the tensor core reads of an fp32 register),
// lo = value - hi (exact)
__device__ __forceinline__ void split(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

constexpr int kTiles = 24;

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, float seed, unsigned long long* clk) {
    __shared__ float s_w[8][8 * 48];  // per warp: 8 points x (x 16 | y 16 | z 12 | pad)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int i = lane; i < 8 * 48; i += 32) s_w[warp][i] = seed + 0.001f * (float)(i * 7 % 13);
    __syncwarp();
    unsigned long long c0 = clock64();
    float acc[kTiles][4];
#pragma unroll
    for (int j = 0; j < kTiles; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;
    uint32_t ahi[4], alo[4], bhi[2], blo[2];
#pragma unroll
    for (int r = 0; r < 4; ++r) split(seed + lane + r, ahi[r], alo[r]);
    split(seed * 3.f + lane, bhi[0], blo[0]);
    split(seed * 5.f + lane, bhi[1], blo[1]);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < kTiles; ++j) {
                mma_tf32(acc[j], ahi, bhi);
                mma_tf32(acc[j], alo, bhi);
                mma_tf32(acc[j], ahi, blo);
            }
        } else {
            const float* w = s_w[warp];
            // A: x windows (value folded in) of points t, t + 4 at rows g, g + 8
            const float2 x0 = *reinterpret_cast<const float2*>(w + t * 48 + 2 * g);
            const float2 x1 = *reinterpret_cast<const float2*>(w + (t + 4) * 48 + 2 * g);
            split(x0.x, ahi[0], alo[0]);
            split(x0.y, ahi[1], alo[1]);
            split(x1.x, ahi[2], alo[2]);
            split(x1.y, ahi[3], alo[3]);
            // y windows: [dy = g >> 1][4 quads], z windows: [dz = g & 1][6 pairs]
            const float4 y0 = *reinterpret_cast<const float4*>(w + t * 48 + 16 + 4 * (g >> 1));
            const float4 y1 = *reinterpret_cast<const float4*>(w + (t + 4) * 48 + 16 + 4 * (g >> 1));
            const float* z0p = w + t * 48 + 32 + 8 * (g & 1);
            const float* z1p = w + (t + 4) * 48 + 32 + 8 * (g & 1);
            const float4 z0a = *reinterpret_cast<const float4*>(z0p);
            const float2 z0b = *reinterpret_cast<const float2*>(z0p + 4);
            const float4 z1a = *reinterpret_cast<const float4*>(z1p);
            const float2 z1b = *reinterpret_cast<const float2*>(z1p + 4);
            const float wy0[4] = {y0.x, y0.y, y0.z, y0.w}, wy1[4] = {y1.x, y1.y, y1.z, y1.w};
            const float wz0[6] = {z0a.x, z0a.y, z0a.z, z0a.w, z0b.x, z0b.y};
            const float wz1[6] = {z1a.x, z1a.y, z1a.z, z1a.w, z1b.x, z1b.y};
#pragma unroll
            for (int jy = 0; jy < 4; ++jy)
#pragma unroll
                for (int jz = 0; jz < 6; ++jz) {
                    split(wy0[jy] * wz0[jz], bhi[0], blo[0]);
                    split(wy1[jy] * wz1[jz], bhi[1], blo[1]);
                    mma_tf32(acc[jy * 6 + jz], ahi, bhi);
                    mma_tf32(acc[jy * 6 + jz], alo, bhi);
                    mma_tf32(acc[jy * 6 + jz], ahi, blo);
                }
            __syncwarp();
            if (lane == 0) s_w[warp][it & 7] += 1e-6f;  // keeps the loads inside the loop
            __syncwarp();
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kTiles; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - c0;
}

// D (16 x 8) = A (16 x 8) * B (8 x 8), one warp; variant 0: one TF32 MMA, 1: 3xTF32 with truncated hi,
// 2: 3xTF32 with round-to-nearest hi (cvt.rna)
__global__ void accuracy(const float* A, const float* B, float* D, int variant) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const float av[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
    const float bv[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
    uint32_t ahi[4], alo[4], bhi[2], blo[2];
    for (int r = 0; r < 4; ++r) {
        if (variant == 2) {
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ahi[r]) : "f"(av[r]));
            alo[r] = __float_as_uint(av[r] - __uint_as_float(ahi[r]));
        } else {
            split(av[r], ahi[r], alo[r]);
        }
    }
    for (int r = 0; r < 2; ++r) {
        if (variant == 2) {
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bhi[r]) : "f"(bv[r]));
            blo[r] = __float_as_uint(bv[r] - __uint_as_float(bhi[r]));
        } else {
            split(bv[r], bhi[r], blo[r]);
        }
    }
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (variant == 0) {
        uint32_t a[4], b[2];
        for (int r = 0; r < 4; ++r) a[r] = __float_as_uint(av[r]);
        for (int r = 0; r < 2; ++r) b[r] = __float_as_uint(bv[r]);
        mma_tf32(c, a, b);
    } else {
        mma_tf32(c, alo, bhi);  // small terms first
        mma_tf32(c, ahi, blo);
        mma_tf32(c, ahi, bhi);
    }
    D[g * 8 + 2 * t] = c[0];
    D[g * 8 + 2 * t + 1] = c[1];
    D[(g + 8) * 8 + 2 * t] = c[2];
    D[(g + 8) * 8 + 2 * t + 1] = c[3];
}

template <int MODE>
void run(const char* name, int sms, double ghz_hint) {
    float* out;
    unsigned long long* clk;
    const int ctas = sms * 2, iters = 2000;
    cudaMalloc(&out, ctas * 256 * 4);
    cudaMalloc(&clk, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(out, 10, 1.25f, clk);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, iters, 1.25f, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h;
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    // 4 warps per sub-partition, each iters * 72 MMAs
    const double mma_per_smsp = 4.0 * iters * 3 * kTiles;
    printf("%-28s %8.3f ms, CTA0 %llu clk (%.2f GHz): %.2f clk per MMA per sub-partition, %.1f clk per k-step per SM\n", name, ms,
           h, h / (ms * 1e6), (double)h / mma_per_smsp, (double)h / (4.0 * iters) /* 16 warps -> 4 k-steps in flight per SMSP */);
    (void)ghz_hint;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error: %s\n", cudaGetErrorString(err));
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("pure mma (24 tiles x 3)", p.multiProcessorCount, 1.9);
    run<1>("k-step (frags + 72 mma)", p.multiProcessorCount, 1.9);
    // accuracy
    std::vector<float> A(128), B(64), D(128);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX;
    for (auto& v : B) v = (float)rand() / RAND_MAX;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 512);
    cudaMalloc(&dB, 256);
    cudaMalloc(&dD, 512);
    cudaMemcpy(dA, A.data(), 512, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), 256, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 3; ++variant) {
        accuracy<<<1, 32>>>(dA, dB, dD, variant);
        cudaMemcpy(D.data(), dD, 512, cudaMemcpyDeviceToHost);
        double emax = 0, e32 = 0;
        for (int i = 0; i < 16; ++i)
            for (int j = 0; j < 8; ++j) {
                double ref = 0;
                float f32 = 0.f;
                for (int kk = 0; kk < 8; ++kk) {
                    ref += (double)A[i * 8 + kk] * (double)B[kk * 8 + j];
                    f32 = fmaf(A[i * 8 + kk], B[kk * 8 + j], f32);
                }
                emax = fmax(emax, fabs(D[i * 8 + j] - ref) / fabs(ref));
                e32 = fmax(e32, fabs((double)f32 - ref) / fabs(ref));
            }
        printf("accuracy variant %d: max rel err %.3e (fp32 fma chain %.3e)\n", variant, emax, e32);
    }
    return 0;
}
