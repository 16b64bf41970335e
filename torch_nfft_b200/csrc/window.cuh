// Spatial window kernels: adjoint spreading and forward interpolation on shared-memory tiles.
//
// Replaces real_/complex_adjoint_window_convolution_kernel and real_/complex_forward_window_
// convolution_kernel (reference csrc/cuda/spatial_window_operations.cu:103-332), which issue one
// global float atomic per (point, channel, tap).  Here every work item = (grid tile, chunk of the
// tile's sorted points) owns a zero-initialised shared-memory copy of the padded tile:
//
//  * spread: the CTA is a "team" in which every thread owns a residue class of tile cells
//    (3D: (y mod L, z mod L), 2D: (x mod L, y mod L), 1D: x mod L).  A point's (2m+2)^d stencil
//    contains exactly one row / cell of every class, so all threads walk the same point list and
//    each accumulates its own cells with plain LDS/FFMA/STS -- no shared-memory float atomics
//    (a CAS loop on sm_100a) and a fixed accumulation order inside the tile.  The tile is then
//    flushed to HBM with 16-byte vector reductions (RED.E.ADD.F32x4), skipping untouched cells.
//  * gather: the padded tile is staged into shared memory with 16-byte loads, one warp per
//    point, lanes over stencil rows, warp-shuffle reduction, one plain store per point/channel
//    (no atomics on y, unlike spatial_window_operations.cu:267,317).
//
// Window taps are recomputed in registers/shared memory from pos (never stored to HBM):
//   shift = (int)floorf(pos*M) - m,  psi_l = expf(-t*t*inv_b)*s,  t = (float)((double)pos*M - shift - l)
// exactly as spatial_window_operations.cu:50,84-86,24-28 (the argument is formed in double there).
#pragma once
#include "common.cuh"
#include "sort.cuh"

namespace nfftb200 {

struct WindowArgs {
    const float* pos;        // [n, dim]
    const float* xin;        // spread: values [n, K] (float components)
    float* yout;             // gather: values [n, K]
    float* grid;             // planar grid, see grid_addr()
    const uint32_t* perm;
    const uint32_t* bin_start;
    const uint32_t* chunk_start;
    const uint4* items;
    long long nbins;
    int k0;                  // first component handled by this pass
    uint32_t* flags;         // point plan flags: flags[0] counts points found outside their tile (stale plan),
                             // flags[1] TMA transfers that did not complete (must stay 0)
    int use_tma;             // 3D register-stencil kernels: tile planes move by TMA (window_reg.cuh)
};

// a point of a work item lies outside the item's tile: the plan was made for other positions
__device__ __forceinline__ void note_dropped_point(const WindowArgs& a) {
    if (a.flags) atomicAdd(a.flags, 1u);
}

__device__ __forceinline__ int fast_div(int x, int d, float inv) {
    int q = (int)((float)x * inv);
    if (q * d > x) --q;
    if ((q + 1) * d <= x) ++q;
    return q;
}

// ------------------------------------------------------------------------------------------
// Point staging.  Every round the CTA stages kSubBatch sorted points into shared memory:
//   s_psi[(q*DIM + slot)*LP + l]   window taps of point q along slot (X, Y, Z), zero padded to LP
//   s_rec[q] = { sx | sxmod << 16, sy | symod << 16, sz | szmod << 16, w }
//     s*  = coordinate of the first tap inside the padded tile, s*mod = s* mod L
//     w   = original point index (gather) or the bits of x[i, k0] (spread)
// Global loads (perm -> pos / x, two dependent HBM accesses) are issued one round ahead into
// registers (prefetch) so that their latency overlaps the previous round's tap loop.
// ------------------------------------------------------------------------------------------
constexpr int kMinThreads = 64;
template <int DIM>
struct StageRegs {
    static constexpr int kItems = (kSubBatch * DIM + kMinThreads - 1) / kMinThreads;
    float p[kItems];
    uint32_t idx[kItems];
};

template <int DIM>
__device__ __forceinline__ void stage_prefetch(StageRegs<DIM>& r, const WindowArgs& a, long long p0, int cnt) {
#pragma unroll
    for (int k = 0; k < StageRegs<DIM>::kItems; ++k) {
        const int w = threadIdx.x + k * blockDim.x;
        if (w < cnt * DIM) {
            const int q = w / DIM, slot = w - q * DIM;
            const uint32_t i = a.perm[p0 + q];
            r.idx[k] = i;
            r.p[k] = a.pos[(size_t)i * DIM + (DIM - 1 - slot)];
        }
    }
}

template <int DIM, int LC>
__device__ __forceinline__ void stage_store(const StageRegs<DIM>& r, const Geom& g, const WindowArgs& a, int cnt,
                                            const int* tile_org, float* s_psi, int* s_rec, bool store_index) {
    const int L = LC ? LC : g.L, LP = g.LP;
#pragma unroll
    for (int k = 0; k < StageRegs<DIM>::kItems; ++k) {
        const int w = threadIdx.x + k * blockDim.x;
        if (w < cnt * DIM) {
            const int q = w / DIM, slot = w - q * DIM;
            const float p = r.p[k];
            const int c = (int)floorf(p * (float)g.M);  // reference cell (spatial_window_operations.cu:50)
            const int sh = c - g.m;                      // reference shift
            const double base = (double)p * (double)g.M - (double)sh;
            float* ps = s_psi + (q * DIM + slot) * LP;
            if (LC) {
#pragma unroll
                for (int l = 0; l < (LC + 3) / 4 * 4; ++l) {
                    const float t = (float)(base - (double)l);
                    ps[l] = l < LC ? expf(-(t * t) * g.inv_b) * g.inv_sqrt_b_pi : 0.f;  // eval_phi, :24-28
                }
            } else {
                for (int l = 0; l < LP; ++l) {
                    const float t = (float)(base - (double)l);
                    ps[l] = l < L ? expf(-(t * t) * g.inv_b) * g.inv_sqrt_b_pi : 0.f;
                }
            }
            // first tap in padded-tile coordinates: wrapped cell - m - (tile origin)
            int s0 = wrap_mod(c, g.M) - g.m - tile_org[slot];
            if (s0 < 0 || s0 > g.P[slot] - L) {
                // only a stale plan (positions changed after the binning) gets here: stay inside the tile
                // and report it through the plan's flag word instead of writing out of bounds
                s0 = s0 < 0 ? 0 : g.P[slot] - L;
                note_dropped_point(a);
            }
            s_rec[q * 4 + slot] = s0 | ((s0 % L) << 16);
            if (slot == 0) {
                if (DIM < 3) s_rec[q * 4 + 2] = 0;
                if (DIM < 2) s_rec[q * 4 + 1] = 0;
                if (store_index) s_rec[q * 4 + 3] = (int)r.idx[k];
            }
        }
    }
}

// address (in floats) of component k of grid cell `cell` for batch entry b
__device__ __forceinline__ size_t grid_plane(const Geom& g, int b, int k) {
    // real: plane (b*C + k); complex: plane (b*C + k/2), interleaved re/im
    return g.cplx ? (size_t)((long long)b * g.C + (k >> 1)) * (size_t)g.Md * 2 : (size_t)((long long)b * g.C + k) * (size_t)g.Md;
}

// One quad (4 consecutive X cells) of component k: real grids are planar (one 16-byte vector
// reduction / load), complex grids interleave re/im (component k = 2 * channel + part, stride 2).
__device__ __forceinline__ void reduce_quad(const Geom& g, float* grid, int b, int k, long long cell, float4 val) {
    if (!g.cplx) {
        atomicAdd(reinterpret_cast<float4*>(grid + grid_plane(g, b, k) + cell), val);
    } else {
        float* p = grid + grid_plane(g, b, k) + 2 * cell + (k & 1);
        if (val.x != 0.f) atomicAdd(p, val.x);
        if (val.y != 0.f) atomicAdd(p + 2, val.y);
        if (val.z != 0.f) atomicAdd(p + 4, val.z);
        if (val.w != 0.f) atomicAdd(p + 6, val.w);
    }
}

__device__ __forceinline__ float4 load_quad(const Geom& g, const float* grid, int b, int k, long long cell) {
    if (!g.cplx) return __ldg(reinterpret_cast<const float4*>(grid + grid_plane(g, b, k) + cell));
    const float* p = grid + grid_plane(g, b, k) + 2 * cell;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    return (k & 1) ? make_float4(v0.y, v0.w, v1.y, v1.w) : make_float4(v0.x, v0.z, v1.x, v1.z);
}

struct TileCtx {
    int b, org[3];
    long long p_lo, p_hi;
    int cls;  // class of the tile (items[].w; Geom::mixed: 1 = heavy tile)
};

__device__ __forceinline__ bool decode_item(const Geom& g, const WindowArgs& a, TileCtx& t) {
    const uint4 it = __ldg(a.items + blockIdx.x);
    if (it.y == it.z) return false;  // beyond the last work item
    const int bin = (int)it.x;
    t.cls = (int)it.w;
    t.p_lo = it.y;
    t.p_hi = it.z;
    t.b = bin / g.tiles_per_batch;
    int r = bin - t.b * g.tiles_per_batch;
    const int tx = r % g.nt[0];
    r /= g.nt[0];
    const int ty = r % g.nt[1];
    const int tz = r / g.nt[1];
    t.org[0] = tx * g.T[0] - g.org[0];
    t.org[1] = ty * g.T[1] - g.org[1];
    t.org[2] = tz * g.T[2] - g.org[2];
    return true;
}

// iterate over the padded tile in groups of 4 consecutive X cells; f(smem_offset, global_cell)
template <int DIM, typename F>
__device__ __forceinline__ void for_each_quad(const Geom& g, const TileCtx& t, F f) {
    const int nx4 = g.P[0] >> 2;
    const int rows = (DIM >= 2 ? g.P[1] : 1) * (DIM >= 3 ? g.P[2] : 1);
    const int total = rows * nx4;
    const float inv_nx4 = 1.0f / (float)nx4;
    const float inv_p1 = 1.0f / (float)g.P[1];
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int row = fast_div(idx, nx4, inv_nx4);
        const int x = (idx - row * nx4) << 2;
        int y = 0, z = 0;
        if (DIM >= 3) {
            z = fast_div(row, g.P[1], inv_p1);
            y = row - z * g.P[1];
        } else if (DIM == 2) {
            y = row;
        }
        const int gx = wrap_mod(t.org[0] + x, g.M);
        long long cell = gx;
        int so = x;
        if (DIM >= 2) {
            cell += (long long)wrap_mod(t.org[1] + y, g.M) * g.M;
            so += y * g.sY;
        }
        if (DIM >= 3) {
            cell += (long long)wrap_mod(t.org[2] + z, g.M) * g.M * g.M;
            so += z * g.sZ;
        }
        f(so, cell);
    }
}

// Adds the CTA's padded tile (NCOMP component planes in shared memory) into the global grid.
template <int DIM, int NCOMP>
__device__ __forceinline__ void flush_tile(const Geom& g, const TileCtx& t, const WindowArgs& a, const float* tile) {
    // vector reductions into the global grid; untouched (== 0) quads are skipped
    if (!g.cplx) {
        for_each_quad<DIM>(g, t, [&](int so, long long cell) {
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) {
                if (a.k0 + k < g.K) {
                    const float* s = tile + (size_t)k * g.tile_elems + so;
                    const float4 val = make_float4(s[0], s[1], s[2], s[3]);
                    if (val.x != 0.f || val.y != 0.f || val.z != 0.f || val.w != 0.f) {
                        float4* dst = reinterpret_cast<float4*>(a.grid + grid_plane(g, t.b, a.k0 + k) + cell);
                        atomicAdd(dst, val);
                    }
                }
            }
        });
    } else {
        for_each_quad<DIM>(g, t, [&](int so, long long cell) {
#pragma unroll
            for (int k = 0; k + 1 < NCOMP; k += 2) {
                if (a.k0 + k < g.K) {
                    const float* re = tile + (size_t)k * g.tile_elems + so;
                    const float* im = tile + (size_t)(k + 1) * g.tile_elems + so;
                    float* dst = a.grid + grid_plane(g, t.b, a.k0 + k) + 2 * cell;
                    const float4 v0 = make_float4(re[0], im[0], re[1], im[1]);
                    const float4 v1 = make_float4(re[2], im[2], re[3], im[3]);
                    if (v0.x != 0.f || v0.y != 0.f || v0.z != 0.f || v0.w != 0.f)
                        atomicAdd(reinterpret_cast<float4*>(dst), v0);
                    if (v1.x != 0.f || v1.y != 0.f || v1.z != 0.f || v1.w != 0.f)
                        atomicAdd(reinterpret_cast<float4*>(dst + 4), v1);
                }
            }
        });
    }
}

// Stages the CTA's padded tile from the global grid into shared memory.
template <int DIM, int NCOMP>
__device__ __forceinline__ void load_tile(const Geom& g, const TileCtx& t, const WindowArgs& a, float* tile) {
    // periodic wrap resolved per quad
    if (!g.cplx) {
        for_each_quad<DIM>(g, t, [&](int so, long long cell) {
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.k0 + k < g.K)
                    val = __ldg(reinterpret_cast<const float4*>(a.grid + grid_plane(g, t.b, a.k0 + k) + cell));
                float* s = tile + (size_t)k * g.tile_elems + so;
                s[0] = val.x; s[1] = val.y; s[2] = val.z; s[3] = val.w;
            }
        });
    } else {
        for_each_quad<DIM>(g, t, [&](int so, long long cell) {
#pragma unroll
            for (int k = 0; k < NCOMP; k += 2) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (a.k0 + k < g.K) {
                    const float* src = a.grid + grid_plane(g, t.b, a.k0 + k) + 2 * cell;
                    v0 = __ldg(reinterpret_cast<const float4*>(src));
                    v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                }
                float* re = tile + (size_t)k * g.tile_elems + so;
                float* im = tile + (size_t)(k + 1 < NCOMP ? k + 1 : k) * g.tile_elems + so;
                re[0] = v0.x; re[1] = v0.z; re[2] = v1.x; re[3] = v1.z;
                if (k + 1 < NCOMP) { im[0] = v0.y; im[1] = v0.w; im[2] = v1.y; im[3] = v1.w; }
            }
        });
    }
}

// ======================================================================================
// adjoint spreading
// ======================================================================================
template <int DIM, int NCOMP, int LC>
__global__ void __launch_bounds__(384)
spread_kernel(const Geom g, const WindowArgs a) {
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;

    const int L = LC ? LC : g.L;
    const int LP = g.LP;
    float* tile = smem;
    float* s_psi = tile + (size_t)NCOMP * g.tile_elems;
    float* s_xv = s_psi + kSubBatch * DIM * LP;
    int* s_rec = (int*)(s_xv + kSubBatch * NCOMP);
    __shared__ int s_org[3];
    if (threadIdx.x < 3) s_org[threadIdx.x] = t.org[threadIdx.x];

    for (int i = threadIdx.x; i < NCOMP * g.tile_elems; i += blockDim.x) tile[i] = 0.f;

    // residue classes owned by this thread
    const int tid = threadIdx.x;
    int c0, c1;  // DIM3: (y class, z class); DIM2: (x class, y class); DIM1: x class
    bool active;
    if (DIM == 1) {
        c0 = tid; c1 = 0; active = tid < L;
    } else {
        c1 = tid / L; c0 = tid - c1 * L; active = tid < L * L;
    }

    // values of the staged points: x[i, k0 .. k0+NCOMP) prefetched with the positions
    constexpr int kXItems = (kSubBatch * NCOMP + kMinThreads - 1) / kMinThreads;
    StageRegs<DIM> regs;
    float xreg[kXItems];
    auto prefetch = [&](long long p0, int cnt) {
        stage_prefetch<DIM>(regs, a, p0, cnt);
#pragma unroll
        for (int k = 0; k < kXItems; ++k) {
            const int w = tid + k * blockDim.x;
            if (w < cnt * NCOMP) {
                const int q = w / NCOMP, kk = w - q * NCOMP;
                const uint32_t i = a.perm[p0 + q];
                xreg[k] = (a.k0 + kk < g.K) ? a.xin[(size_t)i * g.K + a.k0 + kk] : 0.f;
            }
        }
    };
    {
        const long long left = t.p_hi - t.p_lo;
        prefetch(t.p_lo, (int)(left < kSubBatch ? left : kSubBatch));
    }
    __syncthreads();

    for (long long p0 = t.p_lo; p0 < t.p_hi; p0 += kSubBatch) {
        const int cnt = (int)((t.p_hi - p0) < kSubBatch ? (t.p_hi - p0) : kSubBatch);
        stage_store<DIM, LC>(regs, g, a, cnt, s_org, s_psi, s_rec, false);
#pragma unroll
        for (int k = 0; k < kXItems; ++k) {
            const int w = tid + k * blockDim.x;
            if (w < cnt * NCOMP) {
                if (NCOMP == 1) s_rec[w * 4 + 3] = __float_as_int(xreg[k]);
                else s_xv[w] = xreg[k];
            }
        }
        __syncthreads();
        {
            const long long left = t.p_hi - (p0 + kSubBatch);
            if (left > 0) prefetch(p0 + kSubBatch, (int)(left < kSubBatch ? left : kSubBatch));
        }
        if (active) {
            for (int q = 0; q < cnt; ++q) {
                const int4 rec = *reinterpret_cast<const int4*>(s_rec + q * 4);
                const float* ps = s_psi + q * DIM * LP;
                float v[NCOMP];
                if (NCOMP == 1) {
                    v[0] = __int_as_float(rec.w);
                } else {
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) v[k] = s_xv[q * NCOMP + k];
                }
                if (DIM == 3) {
                    int ay = c0 - (rec.y >> 16); ay += ay < 0 ? L : 0;
                    int az = c1 - (rec.z >> 16); az += az < 0 ? L : 0;
                    const float wz = ps[2 * LP + az], wy = ps[LP + ay];
                    // reference product order: x * psi(dim 0 = Z) * psi(dim 1 = Y) * psi(dim 2 = X)
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) v[k] = (v[k] * wz) * wy;
                    float* row = tile + ((rec.z & 0xffff) + az) * g.sZ + ((rec.y & 0xffff) + ay) * g.sY + (rec.x & 0xffff);
                    if (LC) {
                        constexpr int LQ = LC ? (LC + 3) / 4 : 1;
                        float wx[LQ * 4];
#pragma unroll
                        for (int l4 = 0; l4 < LQ; ++l4) {
                            const float4 w4 = reinterpret_cast<const float4*>(ps)[l4];
                            wx[4 * l4] = w4.x; wx[4 * l4 + 1] = w4.y; wx[4 * l4 + 2] = w4.z; wx[4 * l4 + 3] = w4.w;
                        }
#pragma unroll
                        for (int k = 0; k < NCOMP; ++k) {
                            float* r = row + (size_t)k * g.tile_elems;
                            float acc[LC ? LC : 1];
#pragma unroll
                            for (int l = 0; l < LC; ++l) acc[l] = r[l];
#pragma unroll
                            for (int l = 0; l < LC; ++l) acc[l] = fmaf(v[k], wx[l], acc[l]);
#pragma unroll
                            for (int l = 0; l < LC; ++l) r[l] = acc[l];
                        }
                    } else {
                        for (int l = 0; l < L; ++l) {
                            const float wx = ps[l];
#pragma unroll
                            for (int k = 0; k < NCOMP; ++k) {
                                float* r = row + (size_t)k * g.tile_elems + l;
                                *r = fmaf(v[k], wx, *r);
                            }
                        }
                    }
                } else if (DIM == 2) {
                    int ax = c0 - (rec.x >> 16); ax += ax < 0 ? L : 0;
                    int ay = c1 - (rec.y >> 16); ay += ay < 0 ? L : 0;
                    const float wy = ps[LP + ay], wx = ps[ax];
                    float* cellp = tile + ((rec.y & 0xffff) + ay) * g.sY + (rec.x & 0xffff) + ax;
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) {
                        float* r = cellp + (size_t)k * g.tile_elems;
                        *r = fmaf(v[k] * wy, wx, *r);
                    }
                } else {
                    int ax = c0 - (rec.x >> 16); ax += ax < 0 ? L : 0;
                    const float wx = ps[ax];
                    float* cellp = tile + (rec.x & 0xffff) + ax;
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) {
                        float* r = cellp + (size_t)k * g.tile_elems;
                        *r = fmaf(v[k], wx, *r);
                    }
                }
            }
        }
        __syncthreads();
    }

    flush_tile<DIM, NCOMP>(g, t, a, tile);
}

// ======================================================================================
// forward interpolation
// ======================================================================================
constexpr int kGatherThreads = 256;

template <int DIM, int NCOMP, int LC>
__global__ void __launch_bounds__(kGatherThreads)
gather_kernel(const Geom g, const WindowArgs a) {
    extern __shared__ __align__(16) float smem[];
    TileCtx t;
    if (!decode_item(g, a, t)) return;

    const int L = LC ? LC : g.L;
    const int LP = g.LP;
    float* tile = smem;
    float* s_psi = tile + (size_t)NCOMP * g.tile_elems;
    int* s_rec = (int*)(s_psi + kSubBatch * DIM * LP);
    __shared__ int s_org[3];
    if (threadIdx.x < 3) s_org[threadIdx.x] = t.org[threadIdx.x];

    StageRegs<DIM> regs;
    {
        const long long left = t.p_hi - t.p_lo;
        stage_prefetch<DIM>(regs, a, t.p_lo, (int)(left < kSubBatch ? left : kSubBatch));
    }

    load_tile<DIM, NCOMP>(g, t, a, tile);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // stencil rows (3D) / taps (2D) handled by this lane: r = lane + 32 j
    constexpr int kRows = LC ? (LC * LC + 31) / 32 : 1;
    int roff[kRows], ri0[kRows], ri1[kRows];
    if (LC && DIM >= 2) {
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
            const int r = lane + 32 * j;
            const int i1 = r / (LC ? LC : 1), i0 = r - i1 * LC;
            ri0[j] = i0;
            ri1[j] = r < LC * LC ? i1 : -1;
            roff[j] = DIM == 3 ? i1 * g.sZ + i0 * g.sY : i1 * g.sY + i0;
        }
    }

    for (long long p0 = t.p_lo; p0 < t.p_hi; p0 += kSubBatch) {
        const int cnt = (int)((t.p_hi - p0) < kSubBatch ? (t.p_hi - p0) : kSubBatch);
        stage_store<DIM, LC>(regs, g, a, cnt, s_org, s_psi, s_rec, true);
        __syncthreads();
        {
            const long long left = t.p_hi - (p0 + kSubBatch);
            if (left > 0) stage_prefetch<DIM>(regs, a, p0 + kSubBatch, (int)(left < kSubBatch ? left : kSubBatch));
        }
        for (int q = warp; q < cnt; q += nwarps) {
            const int4 rec = *reinterpret_cast<const int4*>(s_rec + q * 4);
            const float* ps = s_psi + q * DIM * LP;
            float acc[NCOMP];
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) acc[k] = 0.f;
            if (DIM == 3) {
                const float* base = tile + (rec.z & 0xffff) * g.sZ + (rec.y & 0xffff) * g.sY + (rec.x & 0xffff);
                if (LC) {
                    constexpr int LQ = LC ? (LC + 3) / 4 : 1;
                    float wx[LQ * 4];
#pragma unroll
                    for (int l4 = 0; l4 < LQ; ++l4) {
                        const float4 w4 = reinterpret_cast<const float4*>(ps)[l4];
                        wx[4 * l4] = w4.x; wx[4 * l4 + 1] = w4.y; wx[4 * l4 + 2] = w4.z; wx[4 * l4 + 3] = w4.w;
                    }
#pragma unroll
                    for (int j = 0; j < kRows; ++j) {
                        if (ri1[j] >= 0) {
                            const float w = ps[2 * LP + ri1[j]] * ps[LP + ri0[j]];  // psi(Z) * psi(Y)
                            const float* row = base + roff[j];
#pragma unroll
                            for (int k = 0; k < NCOMP; ++k) {
                                const float* rk = row + (size_t)k * g.tile_elems;
                                float inner = 0.f;
#pragma unroll
                                for (int l = 0; l < LC; ++l) inner = fmaf(wx[l], rk[l], inner);
                                acc[k] = fmaf(w, inner, acc[k]);
                            }
                        }
                    }
                } else {
                    int ay = lane % L, az = lane / L;
                    for (int r = lane; r < L * L; r += 32) {
                        const float w = ps[2 * LP + az] * ps[LP + ay];
                        const float* row = base + az * g.sZ + ay * g.sY;
#pragma unroll
                        for (int k = 0; k < NCOMP; ++k) {
                            const float* rk = row + (size_t)k * g.tile_elems;
                            float inner = 0.f;
                            for (int l = 0; l < L; ++l) inner = fmaf(ps[l], rk[l], inner);
                            acc[k] = fmaf(w, inner, acc[k]);
                        }
                        ay += 32;
                        while (ay >= L) { ay -= L; ++az; }
                    }
                }
            } else if (DIM == 2) {
                const float* base = tile + (rec.y & 0xffff) * g.sY + (rec.x & 0xffff);
                int ax = lane % L, ay = lane / L;
                for (int r = lane; r < L * L; r += 32) {
                    const float w = ps[LP + ay] * ps[ax];
                    const float* cellp = base + ay * g.sY + ax;
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) acc[k] = fmaf(w, cellp[(size_t)k * g.tile_elems], acc[k]);
                    ax += 32;
                    while (ax >= L) { ax -= L; ++ay; }
                }
            } else {
                if (lane < L) {
                    const float w = ps[lane];
#pragma unroll
                    for (int k = 0; k < NCOMP; ++k) acc[k] = w * tile[(size_t)k * g.tile_elems + (rec.x & 0xffff) + lane];
                }
            }
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            if (lane == 0) {
                float* dst = a.yout + (size_t)(uint32_t)rec.w * g.K + a.k0;
#pragma unroll
                for (int k = 0; k < NCOMP; ++k)
                    if (a.k0 + k < g.K) dst[k] = acc[k];
            }
        }
        __syncthreads();
    }
}

}  // namespace nfftb200
