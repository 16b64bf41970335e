// 1D window kernels.
//
// The team kernels of window.cuh give one thread per residue class of tile cells, which in 1D means
// L = 2m + 2 busy threads per CTA.  In one dimension a work item is small enough for a simpler
// scheme with full parallelism and no read-modify-write conflicts at all:
//   spread : the chunk's points are bucketed by grid cell in shared memory; one thread per PADDED
//            TILE CELL then sums the taps that land on it (the points of the L cells below it).
//            Every (point, tap) pair is evaluated exactly once, by the thread that owns its cell.
//   gather : one thread per point evaluates its L taps against the staged tile.
// Taps are evaluated on the fly with the reference's expression (spatial_window_operations.cu:24-28,
// 84-86): t = (float)((double)pos * M - shift - l), psi = expf(-t^2 * 0.75 pi / m) * sqrtf(0.75 / m);
// for power-of-two M the argument is formed exactly in fp32 (see window_reg.cuh).
//
// Replaces, for d = 1, real_/complex_adjoint_window_convolution_kernel and
// real_/complex_forward_window_convolution_kernel (spatial_window_operations.cu:103-332).
#pragma once
#include "window.cuh"

namespace nfftb200 {

constexpr int kW1Threads = 256;        // gather
constexpr int kW1SpreadThreads = 288;  // 9 warps: the 512 + 20 padded cells of a tile take 2 rounds, not 3
constexpr int kW1MaxPts = 2048;        // points per work item (chunk)

inline size_t w1_spread_smem_bytes(const Geom& g, int ncomp) {
    // tile | point positions | values | cell start[T+1], cursor[T] | cell of each point (u16)
    return (size_t)ncomp * g.tile_elems * 4 + (size_t)kW1MaxPts * 4 + (size_t)kW1MaxPts * ncomp * 4 +
           (size_t)(2 * g.T[0] + 4) * 4 + (size_t)kW1MaxPts * 2;
}
inline size_t w1_gather_smem_bytes(const Geom& g, int ncomp) { return (size_t)ncomp * g.tile_elems * 4; }

// tap l of a point at position p whose reference cell is fl = floorf(p * M)
template <bool POW2>
__device__ __forceinline__ float tap_1d(const Geom& g, float p, float pm, float fl, int l) {
    float tt;
    if (POW2) {
        tt = (pm - fl) + (float)(g.m - l);  // exact fraction, one rounding: = the reference's double evaluation
    } else {
        tt = (float)((double)p * (double)g.M - (double)((int)fl - g.m) - (double)l);
    }
    return expf(-(tt * tt) * g.inv_b) * g.inv_sqrt_b_pi;  // eval_phi
}

template <int NCOMP, bool POW2>
__global__ void __launch_bounds__(kW1SpreadThreads)
spread1d_kernel(const Geom g, const WindowArgs a) {
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    const int T = g.T[0];
    float* tile = smem;
    float* s_p = tile + (size_t)NCOMP * g.tile_elems;
    float* s_x = s_p + kW1MaxPts;
    int* s_start = reinterpret_cast<int*>(s_x + (size_t)kW1MaxPts * NCOMP);
    int* s_cur = s_start + T + 2;
    unsigned short* s_c = reinterpret_cast<unsigned short*>(s_cur + T + 2);

    for (int i = threadIdx.x; i < T; i += kW1SpreadThreads) s_cur[i] = 0;
    __syncthreads();

    // bucket the chunk's points by cell of the tile
    constexpr int kPer = (kW1MaxPts + kW1SpreadThreads - 1) / kW1SpreadThreads;
    const int cnt = (int)(t.p_hi - t.p_lo);
    const int lo0 = t.org[0] + g.org[0];  // first cell of the tile
    const float Mf = (float)g.M;
    float pt[kPer];
    uint32_t src[kPer];
    int cell[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int e = threadIdx.x + k * kW1SpreadThreads;
        cell[k] = -1;
        if (e < cnt) {
            src[k] = a.perm[t.p_lo + e];
            pt[k] = a.pos[src[k]];
            const int c = wrap_mod((int)floorf(pt[k] * Mf), g.M) - lo0;
            if (c >= 0 && c < T) {  // anything else can only come from a stale / foreign sort
                cell[k] = c;
                atomicAdd(&s_cur[c], 1);
            } else {
                note_dropped_point(a);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the cell counts
        int running = 0;
        for (int base = 0; base < T; base += 32) {
            const int idx = base + (int)threadIdx.x;
            const int v = idx < T ? s_cur[idx] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int nb = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += nb;
            }
            if (idx < T) s_start[idx] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (threadIdx.x == 0) s_start[T] = running;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += kW1SpreadThreads) s_cur[i] = s_start[i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (cell[k] >= 0) {
            const int dst = atomicAdd(&s_cur[cell[k]], 1);
            const float pm = pt[k] * Mf;
            s_p[dst] = POW2 ? pm - floorf(pm) : pt[k];  // exact fraction when M is a power of two
            s_c[dst] = (unsigned short)cell[k];
#pragma unroll
            for (int c = 0; c < NCOMP; ++c)
                s_x[dst * NCOMP + c] = a.k0 + c < g.K ? a.xin[(size_t)src[k] * g.K + a.k0 + c] : 0.f;
        }
    }
    __syncthreads();

    // one thread per padded cell j: taps of core cells c with l = j - (c + org - m) in [0, L)
    const int shift = g.org[0] - g.m;
    for (int j = threadIdx.x; j < g.tile_elems; j += kW1SpreadThreads) {
        float acc[NCOMP];
#pragma unroll
        for (int c = 0; c < NCOMP; ++c) acc[c] = 0.f;
        int c_hi = j - shift, c_lo = c_hi - (g.L - 1);
        c_lo = c_lo < 0 ? 0 : c_lo;
        c_hi = c_hi > T - 1 ? T - 1 : c_hi;
        // one flat loop over the points of cells c_lo .. c_hi (contiguous after the bucketing): nested
        // per-cell loops diverge badly, the lanes of a warp own cells with different point counts
        const int q_end = c_lo <= c_hi ? s_start[c_hi + 1] : 0;
        for (int q = c_lo <= c_hi ? s_start[c_lo] : 0; q < q_end; ++q) {
            const int l = j - shift - (int)s_c[q];  // tap of point q that lands on cell j
            float w;
            if (POW2) {
                const float tt = s_p[q] + (float)(g.m - l);  // s_p holds the exact fraction: one rounding
                w = expf(-(tt * tt) * g.inv_b) * g.inv_sqrt_b_pi;
            } else {
                const float p = s_p[q];
                const float pm = p * Mf;
                w = tap_1d<false>(g, p, pm, floorf(pm), l);
            }
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) acc[k] = fmaf(s_x[q * NCOMP + k], w, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < NCOMP; ++k) tile[(size_t)k * g.tile_elems + j] = acc[k];
    }
    __syncthreads();
    flush_tile<1, NCOMP>(g, t, a, tile);
}

template <int NCOMP, bool POW2>
__global__ void __launch_bounds__(kW1Threads)
gather1d_kernel(const Geom g, const WindowArgs a) {
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    float* tile = smem;
    load_tile<1, NCOMP>(g, t, a, tile);
    __syncthreads();
    const int cnt = (int)(t.p_hi - t.p_lo);
    const int lo0 = t.org[0] + g.org[0];
    const int shift = g.org[0] - g.m;
    const float Mf = (float)g.M;
    for (int e = threadIdx.x; e < cnt; e += kW1Threads) {
        const uint32_t i = a.perm[t.p_lo + e];
        const float p = a.pos[i];
        const float pm = p * Mf;
        const float fl = floorf(pm);
        const int c = wrap_mod((int)fl, g.M) - lo0;
        float acc[NCOMP];
#pragma unroll
        for (int k = 0; k < NCOMP; ++k) acc[k] = 0.f;
        if (c < 0 || c >= g.T[0]) note_dropped_point(a);  // stale / foreign plan
        if (c >= 0 && c < g.T[0]) {
            const float* row = tile + c + shift;
            for (int l = 0; l < g.L; ++l) {
                const float w = tap_1d<POW2>(g, p, pm, fl, l);
#pragma unroll
                for (int k = 0; k < NCOMP; ++k) acc[k] = fmaf(row[(size_t)k * g.tile_elems + l], w, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < NCOMP; ++k)
            if (a.k0 + k < g.K) a.yout[(size_t)i * g.K + a.k0 + k] = acc[k];
    }
}

}  // namespace nfftb200
