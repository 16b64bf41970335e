// Register-stencil window kernels (2D, several channels): the 2D counterpart of window_reg.cuh.
//
// The chunk's points are bucketed by supercell (4 x 4 oversampled cells).  One warp owns a
// supercell at a time and keeps the (L+3) x (L+3) block of grid cells it touches, for all NCOMP
// channels of the pass, in registers: lane <-> (x, y) positions (c = lane + 32 q), NCOMP
// accumulators per position.  The window weights psi_y * psi_x are channel independent
// (reference spatial_window_operations.cu:146-156 recomputes them per channel), so a point costs
// CPL weight products plus CPL * NCOMP FFMAs.  The shared-memory tile (NCOMP planes) is touched
// once per supercell: spread adds the block out under row-band locks, gather loads it.
#pragma once
#include "window_reg.cuh"

#ifndef NFFT_REG2_FFMA2
#define NFFT_REG2_FFMA2 1
#endif
#ifndef NFFT_REG2_BANDED
#define NFFT_REG2_BANDED 1
#endif

namespace nfftb200 {

constexpr int kReg2Threads = 256;
constexpr int kReg2Warps = kReg2Threads / 32;
constexpr int kReg2MaxPts = 1280;  // points per work item held in shared memory
constexpr int kReg2Group = 8;      // points staged per warp round
constexpr int kReg2S = 4;          // supercell edge (cells)

template <int LC, int NCOMP>
struct Reg2Cfg {
    static constexpr int W = LC + kReg2S - 1;               // block edge
    static constexpr int COLS = W * W;
    static constexpr int CPL = (COLS + 31) / 32;            // positions per lane
    static constexpr int XYP = W + 1 + ((W + 1) % 2 == 0);  // odd window pitch, entry XYP-1 always zero
    static constexpr int WIN_FLOATS = (2 * kReg2Group * XYP + 3) / 4 * 4;
    // point record in shared memory (floats): spread [x_0..x_{NCOMP-1}, px, py, pad], gather [idx, px, py, pad]
    static constexpr int PITCH_SPREAD = NCOMP >= 4 ? NCOMP + 4 : 4;
    static constexpr int POS_SPREAD = NCOMP == 1 ? 1 : NCOMP;
};

inline size_t reg2_smem_bytes(const Geom& g, int ncomp, bool spread, int nsc, int win_floats) {
    const int pitch = spread ? (ncomp >= 4 ? ncomp + 4 : 4) : 4;
    return (size_t)ncomp * g.tile_elems * 4 + (size_t)kReg2MaxPts * pitch * 4 + (size_t)kReg2Warps * win_floats * 4 +
           (size_t)(2 * nsc + 4) * 4 + (size_t)kReg2MaxPts + 64;
}

// Loads the chunk's points, buckets them by supercell; record layout see Reg2Cfg.
template <int NCOMP, bool SPREAD, int PITCH, int POS>
__device__ __forceinline__ void bucket_points_2d(const Geom& g, const WindowArgs& a, const TileCtx& t, int cnt, int nsx,
                                                 int nsy, float* s_rec, unsigned char* s_off, int* s_start, int* s_cur) {
    constexpr int kPer = (kReg2MaxPts + kReg2Threads - 1) / kReg2Threads;
    const int nsc = nsx * nsy;
    float px[kPer], py[kPer];
    uint32_t id[kPer];
    int sc[kPer];
    const int lo0 = t.org[0] + g.org[0], lo1 = t.org[1] + g.org[1];
    const float Mf = (float)g.M;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int e = threadIdx.x + k * kReg2Threads;
        sc[k] = -1;
        if (e < cnt) {
            const uint32_t i = a.perm[t.p_lo + e];
            id[k] = i;
            py[k] = a.pos[(size_t)i * 2 + 0];  // API dim 0 = slot Y
            px[k] = a.pos[(size_t)i * 2 + 1];
            const int cy = wrap_mod((int)floorf(py[k] * Mf), g.M) - lo1;
            const int cx = wrap_mod((int)floorf(px[k] * Mf), g.M) - lo0;
            const int bx = cx / kReg2S, by = cy / kReg2S;
            if (cx >= 0 && cy >= 0 && bx < nsx && by < nsy) {
                sc[k] = (by * nsx + bx) | ((cx - bx * kReg2S) | (cy - by * kReg2S) << 2) << 24;
                atomicAdd(&s_cur[sc[k] & 0xffffff], 1);
            } else {
                note_dropped_point(a);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int running = 0;
        for (int base = 0; base < nsc; base += 32) {
            const int idx = base + (int)threadIdx.x;
            const int v = idx < nsc ? s_cur[idx] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += n;
            }
            if (idx < nsc) s_start[idx] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (threadIdx.x == 0) s_start[nsc] = running;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsc; i += kReg2Threads) s_cur[i] = s_start[i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (sc[k] >= 0) {
            const int dst = atomicAdd(&s_cur[sc[k] & 0xffffff], 1);
            float* r = s_rec + (size_t)dst * PITCH;
            if (SPREAD) {
#pragma unroll
                for (int c = 0; c < NCOMP; ++c)
                    r[c] = (a.k0 + c < g.K) ? a.xin[(size_t)id[k] * g.K + a.k0 + c] : 0.f;
            } else {
                r[0] = __int_as_float((int)id[k]);
            }
            r[POS] = px[k];
            r[POS + 1] = py[k];
            s_off[dst] = (unsigned char)((unsigned)sc[k] >> 24);
        }
    }
    __syncthreads();
}

// taps of up to kReg2Group points: lane <-> (point, dimension), L unrolled expf chains per lane
template <typename Cfg, int LC, int PITCH, int POS>
__device__ __forceinline__ void stage_windows_2d(const Geom& g, const float* s_rec, const unsigned char* s_off, int base,
                                                 int npts, float* win, int lane, bool pow2) {
    constexpr int kQuads = Cfg::WIN_FLOATS / 4;
#pragma unroll
    for (int k = 0; k < (kQuads + 31) / 32; ++k) {
        const int qd = lane + 32 * k;
        if (qd < kQuads) reinterpret_cast<float4*>(win)[qd] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const int pt = lane >> 1, slot = lane & 1;  // slot 0 = X (API dim 1), slot 1 = Y (API dim 0)
    if (pt < npts) {
        const float p = s_rec[(size_t)(base + pt) * PITCH + POS + slot];
        const int off = (s_off[base + pt] >> (2 * slot)) & 3;
        float* dst = win + (2 * pt + slot) * Cfg::XYP + off;
        const float pm = p * (float)g.M;
        const float fl = floorf(pm);  // reference cell (spatial_window_operations.cu:50)
        if (pow2) {
            const float frac = pm - fl;  // exact; one rounding per tap below, see window_reg.cuh
#if NFFT_WINDOW_RECUR
            window_taps_recur<LC>(g, frac, g.inv_sqrt_b_pi, [&](int l, float v) { dst[l] = v; });
#else
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = frac + (float)((LC - 2) / 2 - l);  // m - l, m = (L - 2) / 2
                dst[l] = window_exp(-(tt * tt) * g.inv_b) * g.inv_sqrt_b_pi;  // eval_phi, :24-28
            }
#endif
        } else {
            const double bd = (double)p * (double)g.M - (double)((int)fl - g.m);
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = (float)(bd - (double)l);
                dst[l] = window_exp(-(tt * tt) * g.inv_b) * g.inv_sqrt_b_pi;
            }
        }
    }
    __syncwarp();
}

// ======================================================================================
// spread
// ======================================================================================
template <int LC, int NCOMP>
__global__ void __launch_bounds__(kReg2Threads, 2)
spread_reg2d_kernel(const Geom g, const WindowArgs a) {
    using Cfg = Reg2Cfg<LC, NCOMP>;
    constexpr int W = Cfg::W, CPL = Cfg::CPL, PITCH = Cfg::PITCH_SPREAD, POS = Cfg::POS_SPREAD;
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;

    const int nsx = (g.T[0] + kReg2S - 1) / kReg2S, nsy = (g.T[1] + kReg2S - 1) / kReg2S;
    const int nsc = nsx * nsy;
    float* tile = smem;
    float* s_rec = tile + (size_t)NCOMP * g.tile_elems;
    float* s_win = s_rec + (size_t)kReg2MaxPts * PITCH;
    int* s_start = reinterpret_cast<int*>(s_win + kReg2Warps * Cfg::WIN_FLOATS);
    int* s_cur = s_start + nsc + 2;
    unsigned char* s_off = reinterpret_cast<unsigned char*>(s_cur + nsc + 2);
    __shared__ int s_next;
    __shared__ int s_lock[32];  // one lock per band of kReg2S tile rows
    __shared__ int s_order[64];
    if (threadIdx.x < 32) s_lock[threadIdx.x] = 0;

    for (int i = threadIdx.x; i < NCOMP * g.tile_elems; i += kReg2Threads) tile[i] = 0.f;
    for (int i = threadIdx.x; i < nsc; i += kReg2Threads) s_cur[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    const int cnt = (int)(t.p_hi - t.p_lo);
    bucket_points_2d<NCOMP, true, PITCH, POS>(g, a, t, cnt, nsx, nsy, s_rec, s_off, s_start, s_cur);
    order_columns(s_start, nsc, 1, s_order);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* win = s_win + warp * Cfg::WIN_FLOATS;
    const int padx = g.org[0] - g.m;
    const bool pow2 = (g.M & (g.M - 1)) == 0;
    int wi[CPL], wj[CPL], coff[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        const bool ok = c < Cfg::COLS;
        wi[q] = ok ? c % W : Cfg::XYP - 1;
        wj[q] = ok ? Cfg::XYP + c / W : 2 * Cfg::XYP - 1;
        coff[q] = ok ? (c / W) * g.sY + (c % W) : 0;
    }

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(&s_next, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= nsc) break;
        const int sc = s_order[tk];
        const int lo = s_start[sc], hi = s_start[sc + 1];
        if (lo == hi) break;  // supercells are ordered by size: the rest is empty
        const int scx = sc % nsx, scy = sc / nsx;

        // channel pairs share one packed fp32x2 FMA (scalar-broadcast weight)
        constexpr int NC2 = (NCOMP + 1) / 2;
        float2 acc2[CPL][NC2];
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
            for (int c = 0; c < NC2; ++c) acc2[q][c] = make_float2(0.f, 0.f);

        for (int base = lo; base < hi; base += kReg2Group) {
            const int npts = hi - base < kReg2Group ? hi - base : kReg2Group;
            stage_windows_2d<Cfg, LC, PITCH, POS>(g, s_rec, s_off, base, npts, win, lane, pow2);
            const float* wv0 = win;
            const float* rec0 = s_rec + (size_t)base * PITCH;
            // one copy of the point body per slot of the round: the 12 window loads get immediate offsets
            // instead of 12 address additions per point (60 -> 44 instructions per point with 8 channels)
#pragma unroll
            for (int gp = 0; gp < kReg2Group; ++gp) {
                if (gp >= npts) break;
                const float* wv = wv0 + gp * 2 * Cfg::XYP;
                const float* rec = rec0 + gp * PITCH;
                float xs[NCOMP];
                if (NCOMP >= 4) {
#pragma unroll
                    for (int c4 = 0; c4 < NCOMP / 4; ++c4) {
                        const float4 x4 = reinterpret_cast<const float4*>(rec)[c4];
                        xs[4 * c4] = x4.x; xs[4 * c4 + 1] = x4.y; xs[4 * c4 + 2] = x4.z; xs[4 * c4 + 3] = x4.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCOMP; ++c) xs[c] = rec[c];
                }
                float2 xs2[NC2];
#pragma unroll
                for (int c = 0; c < NC2; ++c) xs2[c] = make_float2(xs[2 * c], 2 * c + 1 < NCOMP ? xs[2 * c + 1] : 0.f);
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const float v = wv[wj[q]] * wv[wi[q]];  // psi(dim 0 = Y) * psi(dim 1 = X)
                    const float2 vv = make_float2(v, v);
#pragma unroll
                    for (int c = 0; c < NC2; ++c) {
                        if (NCOMP >= 2 && NFFT_REG2_FFMA2) {
                            acc2[q][c] = __ffma2_rn(vv, xs2[c], acc2[q][c]);
                        } else {
                            acc2[q][c].x = fmaf(v, xs2[c].x, acc2[q][c].x);
                            acc2[q][c].y = fmaf(v, xs2[c].y, acc2[q][c].y);
                        }
                    }
                }
            }
            __syncwarp();
        }
        // add the block into the shared tile planes, one band of kReg2S tile rows at a time: rows
        // [4 scy, 4 scy + W) live in bands scy .. scy + (W-1)/4, each guarded by its own lock, so warps
        // working on overlapping supercells pipeline through the bands instead of excluding each other
        // for the whole block
        constexpr int kBands = (W + kReg2S - 1) / kReg2S;
        float* bbase = tile + (scy * kReg2S) * g.sY + scx * kReg2S + padx;
#if NFFT_REG2_BANDED
#pragma unroll
        for (int b = 0; b < kBands; ++b) {
            if (lane == 0) {
                lock_acquire(&s_lock[scy + b]);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int row = (lane + 32 * q) / W;
                if (lane + 32 * q < Cfg::COLS && row / kReg2S == b) {
                    float cur[NCOMP];
#pragma unroll
                    for (int c = 0; c < NCOMP; ++c) cur[c] = bbase[(size_t)c * g.tile_elems + coff[q]];
#pragma unroll
                    for (int c = 0; c < NCOMP; ++c) bbase[(size_t)c * g.tile_elems + coff[q]] = cur[c] + ((c & 1) ? acc2[q][c >> 1].y : acc2[q][c >> 1].x);
                }
            }
            release_fence();
            __syncwarp();
            if (lane == 0) atomicExch(&s_lock[scy + b], 0);
        }
#else
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < kBands; ++b)
                lock_acquire(&s_lock[scy + b]);  // ascending order: no deadlock
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            if (lane + 32 * q < Cfg::COLS) {
                float cur[NCOMP];
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) cur[c] = bbase[(size_t)c * g.tile_elems + coff[q]];
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) bbase[(size_t)c * g.tile_elems + coff[q]] = cur[c] + ((c & 1) ? acc2[q][c >> 1].y : acc2[q][c >> 1].x);
            }
        }
        release_fence();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < kBands; ++b) atomicExch(&s_lock[scy + b], 0);
        }
#endif
    }
    __syncthreads();

    // flush: vector reductions into the global grid; untouched (== 0) quads are skipped
    for_each_quad<2>(g, t, [&](int so, long long cell) {
#pragma unroll
        for (int k = 0; k < NCOMP; ++k) {
            if (a.k0 + k < g.K) {
                const float* s = tile + (size_t)k * g.tile_elems + so;
                const float4 val = make_float4(s[0], s[1], s[2], s[3]);
                if (val.x != 0.f || val.y != 0.f || val.z != 0.f || val.w != 0.f)
                    reduce_quad(g, a.grid, t.b, a.k0 + k, cell, val);
            }
        }
    });
}

// ======================================================================================
// gather
// ======================================================================================
template <int LC, int NCOMP>
__global__ void __launch_bounds__(kReg2Threads, 2)
gather_reg2d_kernel(const Geom g, const WindowArgs a) {
    using Cfg = Reg2Cfg<LC, NCOMP>;
    constexpr int W = Cfg::W, CPL = Cfg::CPL, PITCH = 4, POS = 1;
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;

    const int nsx = (g.T[0] + kReg2S - 1) / kReg2S, nsy = (g.T[1] + kReg2S - 1) / kReg2S;
    const int nsc = nsx * nsy;
    float* tile = smem;
    float* s_rec = tile + (size_t)NCOMP * g.tile_elems;
    float* s_win = s_rec + (size_t)kReg2MaxPts * PITCH;
    int* s_start = reinterpret_cast<int*>(s_win + kReg2Warps * Cfg::WIN_FLOATS);
    int* s_cur = s_start + nsc + 2;
    unsigned char* s_off = reinterpret_cast<unsigned char*>(s_cur + nsc + 2);
    __shared__ int s_next;
    __shared__ int s_order[64];

    for (int i = threadIdx.x; i < nsc; i += kReg2Threads) s_cur[i] = 0;
    if (threadIdx.x == 0) s_next = 0;
    // the tile planes travel with asynchronous copies (16 bytes per quad when the tile strides keep the
    // quads aligned, else 4 x 4 bytes) while the points are loaded and bucketed
    {
        const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
        const bool vec = !g.cplx && ((g.sY | g.tile_elems) & 3) == 0;
        const int cs = g.cplx ? 2 : 1;
        for_each_quad<2>(g, t, [&](int so, long long cell) {
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) {
                const uint32_t dst = tile_s + 4u * (uint32_t)(k * g.tile_elems + so);
                if (a.k0 + k < g.K) {
                    const float* src = a.grid + grid_plane(g, t.b, a.k0 + k) + (g.cplx ? ((a.k0 + k) & 1) : 0) + cs * cell;
                    if (vec) {
                        cp_async16(dst, src);
                    } else {
                        cp_async4(dst, src);
                        cp_async4(dst + 4, src + cs);
                        cp_async4(dst + 8, src + 2 * cs);
                        cp_async4(dst + 12, src + 3 * cs);
                    }
                } else {
                    float* s = tile + (size_t)k * g.tile_elems + so;
                    s[0] = 0.f; s[1] = 0.f; s[2] = 0.f; s[3] = 0.f;
                }
            }
        });
    }
    __syncthreads();
    const int cnt = (int)(t.p_hi - t.p_lo);
    bucket_points_2d<NCOMP, false, PITCH, POS>(g, a, t, cnt, nsx, nsy, s_rec, s_off, s_start, s_cur);
    cp_async_wait_all();  // order_columns ends with a barrier, which publishes the tile
    order_columns(s_start, nsc, 1, s_order);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* win = s_win + warp * Cfg::WIN_FLOATS;
    const int padx = g.org[0] - g.m;
    const bool pow2 = (g.M & (g.M - 1)) == 0;
    int wi[CPL], wj[CPL], coff[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        const bool ok = c < Cfg::COLS;
        wi[q] = ok ? c % W : Cfg::XYP - 1;
        wj[q] = ok ? Cfg::XYP + c / W : 2 * Cfg::XYP - 1;
        coff[q] = ok ? (c / W) * g.sY + (c % W) : 0;
    }
    constexpr int kLanesPerChannel = 32 / NCOMP;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomicAdd(&s_next, 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= nsc) break;
        const int sc = s_order[tk];
        const int lo = s_start[sc], hi = s_start[sc + 1];
        if (lo == hi) break;
        const int scx = sc % nsx, scy = sc / nsx;
        const float* bbase = tile + (scy * kReg2S) * g.sY + scx * kReg2S + padx;

        float blk[CPL][NCOMP];
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
            for (int c = 0; c < NCOMP; ++c) blk[q][c] = bbase[(size_t)c * g.tile_elems + coff[q]];

        for (int base = lo; base < hi; base += kReg2Group) {
            const int npts = hi - base < kReg2Group ? hi - base : kReg2Group;
            stage_windows_2d<Cfg, LC, PITCH, POS>(g, s_rec, s_off, base, npts, win, lane, pow2);
            const float* wv0 = win;
#pragma unroll
            for (int gp = 0; gp < kReg2Group; ++gp) {
                if (gp >= npts) break;
                const float* wv = wv0 + gp * 2 * Cfg::XYP;
                constexpr int NC2 = (NCOMP + 1) / 2;
                float2 part2[NC2];
#pragma unroll
                for (int c = 0; c < NC2; ++c) part2[c] = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const float v = wv[wj[q]] * wv[wi[q]];  // psi(Y) * psi(X); zero for unused positions
                    const float2 vv = make_float2(v, v);
#pragma unroll
                    for (int c = 0; c < NC2; ++c) {
                        const float2 b2 = make_float2(blk[q][2 * c], 2 * c + 1 < NCOMP ? blk[q][2 * c + 1] : 0.f);
                        if (NCOMP >= 2 && NFFT_REG2_FFMA2) {
                            part2[c] = __ffma2_rn(vv, b2, part2[c]);
                        } else {
                            part2[c].x = fmaf(v, b2.x, part2[c].x);
                            part2[c].y = fmaf(v, b2.y, part2[c].y);
                        }
                    }
                }
                float part[NCOMP];
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) part[c] = (c & 1) ? part2[c >> 1].y : part2[c >> 1].x;
                warp_reduce_channels<NCOMP>(part, lane);
                if ((lane & (kLanesPerChannel - 1)) == 0) {
                    const int c = lane / kLanesPerChannel;
                    if (a.k0 + c < g.K) {
                        const uint32_t i = (uint32_t)__float_as_int(s_rec[(size_t)(base + gp) * PITCH]);
                        a.yout[(size_t)i * g.K + a.k0 + c] = part[0];
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace nfftb200
