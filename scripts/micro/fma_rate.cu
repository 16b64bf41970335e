// Microbenchmark: issue rate of FFMA, FFMA2 (packed fp32x2), FMUL and IMAD per SM sub-partition at
// the occupancy of the window kernels (16 warps / SM).  Prints warp-instructions per clock per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, float a, float b, unsigned long long* clk) {
    unsigned long long t0 = 0, c0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    float2 acc[12];
    float2 w[6];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = make_float2(a + i, b - i);
    int ia = threadIdx.x, ib = (int)a + 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                if (MODE == 0) {  // FFMA x2 (scalar)
                    acc[i].x = fmaf(w[i % 6].x, w[(i + r) % 6].y, acc[i].x);
                    acc[i].y = fmaf(w[i % 6].y, w[(i + r) % 6].x, acc[i].y);
                } else if (MODE == 1) {  // FFMA2
                    acc[i] = __ffma2_rn(w[i % 6], w[(i + r) % 6], acc[i]);
                } else if (MODE == 2) {  // FFMA2 + IMAD interleaved (address-like integer work)
                    acc[i] = __ffma2_rn(w[i % 6], w[(i + r) % 6], acc[i]);
                    ia = ia * ib + i;
                } else if (MODE == 3) {  // FFMA2 + IADD3/LOP
                    acc[i] = __ffma2_rn(w[i % 6], w[(i + r) % 6], acc[i]);
                    ia = (ia + ib) ^ i;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)ia;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[0] = t1 - t0;
        clk[1] = clock64() - c0;
    }
}

// the spread kernel's inner block: 36 accumulators, acc[q][kp] += v[q] (broadcast) * wz[kp]
__global__ void __launch_bounds__(256, 2) kblock(float* out, int iters, float a, float b, unsigned long long* clk) {
    unsigned long long t0 = 0, c0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        c0 = clock64();
    }
    float2 acc[6][6];
    float2 wz[6];
    float v[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        v[q] = a + q + threadIdx.x * 0.01f;
        wz[q] = make_float2(b + q, b - q);
#pragma unroll
        for (int kp = 0; kp < 6; ++kp) acc[q][kp] = make_float2(0.f, 0.f);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const float2 vv = make_float2(v[q], v[q]);
#pragma unroll
            for (int kp = 0; kp < 6; ++kp) acc[q][kp] = __ffma2_rn(vv, wz[kp], acc[q][kp]);
        }
        // keep the operands loop-variant without adding FMA-pipe work
        v[it % 6 == 0 ? 0 : 1] = __int_as_float(__float_as_int(v[0]) ^ 1);
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int kp = 0; kp < 6; ++kp) s += acc[q][kp].x + acc[q][kp].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        clk[0] = t1 - t0;
        clk[1] = clock64() - c0;
    }
}

void run_block() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 256);
    unsigned long long* clk;
    cudaMallocManaged(&clk, 16);
    const int iters = 40000;
    kblock<<<sms * 2, 256>>>(out, 100, 1.0f, 2.0f, clk);
    cudaDeviceSynchronize();
    kblock<<<sms * 2, 256>>>(out, iters, 1.0f, 2.0f, clk);
    cudaDeviceSynchronize();
    // 16 warps / SM = 4 per SMSP, both CTAs resident (one wave)
    const double per_smsp = 36.0 * iters * 4.0;
    printf("%-28s %.3f cycles per FFMA2 per SMSP (block 0 ran %llu cycles, %.0f MHz)\n", "spread block 6x6 FFMA2",
           (double)clk[1] / per_smsp, clk[1], (double)clk[1] / (double)clk[0] * 1e3);
}

template <int MODE>
void run(const char* name, double inst_per_iter) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 256);
    const int iters = 20000;
    unsigned long long* clk;
    cudaMallocManaged(&clk, 16);
    k<MODE><<<sms * 2, 256>>>(out, 100, 1.0f, 2.0f, clk);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * 2, 256>>>(out, iters, 1.0f, 2.0f, clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mhz = (double)clk[1] / (double)clk[0] * 1e3;
    const double cycles = (double)clk[1];
    const double warp_inst_per_smsp = inst_per_iter * iters * 16.0 / 4.0;  // 16 warps / SM over 4 SMSPs
    printf("%-28s %.3f ms  %.3f warp-inst/clk/SMSP (SM clock %.0f MHz measured, %d nominal)\n", name, ms, warp_inst_per_smsp / cycles, mhz, khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("FFMA (96 per iter)", 96);
    run<1>("FFMA2 (48 per iter)", 48);
    run<2>("FFMA2 + IMAD (48+48)", 96);
    run<3>("FFMA2 + IADD3/LOP (48+96)", 144);
    run_block();
    return 0;
}
