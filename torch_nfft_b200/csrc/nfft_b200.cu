// libnfft_b200.so -- host orchestration and C ABI of the sm_100a NFFT engine.
//
// Mirrors the three host pipelines of the reference (csrc/cuda/core_cuda.cu:144-336 adjoint,
// :340-531 forward, :535-852 fastsum) with a different execution plan:
//   sort points by tile -> spread on shared-memory tiles -> cuFFT (cached plan, R2C/C2R for real
//   data, caller's stream) -> fused unpack;  pack -> cuFFT -> tile gather;  no per-call
//   cudaMalloc / plan creation / device synchronisation.
#include "common.cuh"
#include "sort.cuh"
#include "spectral.cuh"
#include "fft_rows.cuh"
#include "window.cuh"
#include "window_reg.cuh"
#include "window_reg2d.cuh"
#include "window1d.cuh"
#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>


namespace nfftb200 {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// ----------------------------------------------------------------------------------------
// optional per-stage timing with CUDA events on the caller's stream (bench.py roofline)
// ----------------------------------------------------------------------------------------
enum Stage { ST_SORT = 0, ST_SPREAD, ST_FFT, ST_UNPACK, ST_PACK, ST_GATHER, ST_MULTIPLY, ST_MEMSET, ST_COUNT };
struct ProfSpan { int stage; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::mutex g_prof_mutex;
static std::vector<ProfSpan> g_prof_spans;
static std::vector<cudaEvent_t> g_prof_pool;

static const char* const kStageNames[] = {"nfftb200:sort",   "nfftb200:spread", "nfftb200:fft",      "nfftb200:unpack",
                                          "nfftb200:pack",   "nfftb200:gather", "nfftb200:multiply", "nfftb200:memset"};

// One pipeline stage: an NVTX range around its enqueue (visible to Nsight tools; a no-op without one attached)
// and, while profiling is enabled, a pair of CUDA events on the caller's stream.
struct ProfScope {
    ProfSpan span{};
    cudaStream_t st;
    bool on;
    ProfScope(int stage, cudaStream_t s) : st(s), on(g_prof_on) {
        nvtxRangePushA(kStageNames[stage]);
        if (!on) return;
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        for (cudaEvent_t* e : {&span.a, &span.b}) {
            if (!g_prof_pool.empty()) { *e = g_prof_pool.back(); g_prof_pool.pop_back(); }
            else if (cudaEventCreate(e) != cudaSuccess) { on = false; return; }
        }
        span.stage = stage;
        cudaEventRecord(span.a, st);
    }
    ~ProfScope() {
        nvtxRangePop();
        if (!on) return;
        cudaEventRecord(span.b, st);
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        g_prof_spans.push_back(span);
    }
};

// ----------------------------------------------------------------------------------------
// geometry
// ----------------------------------------------------------------------------------------
// Extra shared-memory wavefronts (bank conflicts) per warp-wide tile access, averaged over the
// access patterns of the two kernels:
//   gather: lane t reads offset (t % L) * s0 + (t / L) * s1                      (stencil rows)
//   spread: lane t owns class (t % L, t / L) and touches offset
//           ((t % L - u) mod L) * s0 + ((t / L - v) mod L) * s1 for a point with first-tap
//           residues (u, v) -- all L^2 residues are equally likely.
// (3D: s0 = sY, s1 = sZ over (y, z); 2D: s0 = 1, s1 = sY over (x, y).)
static double conflict_score(int dim, int L, int sY, int sZ) {
    const int s0 = dim == 3 ? sY : 1, s1 = dim == 3 ? sZ : sY;
    const int team = L * L;
    auto warp_extra = [&](int u, int v) {
        int extra = 0;
        for (int w0 = 0; w0 < team; w0 += 32) {
            int cnt[32] = {0};
            for (int t = w0; t < w0 + 32 && t < team; ++t) {
                const int a0 = ((t % L) - u + L) % L, a1 = ((t / L) - v + L) % L;
                cnt[(a0 * s0 + a1 * s1) & 31]++;
            }
            int mx = 0;
            for (int b = 0; b < 32; ++b) mx = cnt[b] > mx ? cnt[b] : mx;
            extra += mx - 1;
        }
        return extra;
    };
    double spread = 0.0;
    for (int u = 0; u < L; ++u)
        for (int v = 0; v < L; ++v) spread += warp_extra(u, v);
    spread /= (double)(L * L);
    const double gather = warp_extra(0, 0);
    const int nwarps = (team + 31) / 32;
    return (spread + gather) / (2.0 * nwarps);
}

struct StrideKey {
    int dim, L, P0, P1, P2;
    bool operator<(const StrideKey& o) const {
        return std::tie(dim, L, P0, P1, P2) < std::tie(o.dim, o.L, o.P0, o.P1, o.P2);
    }
};
static std::mutex g_stride_mutex;
static std::map<StrideKey, std::pair<int, int>> g_strides;

// shared-memory strides: trade padding against bank conflicts; cost = tile floats * (1 + extra
// wavefronts per access).  Cached: the search is a few million integer ops.
static void choose_strides(Geom& g) {
    g.sY = g.P[0];
    g.sZ = g.P[0] * g.P[1];
    if (g.dim == 1) return;
#ifndef NFFT_REG_OLD_STRIDES
#define NFFT_REG_OLD_STRIDES 0  // bisect aid: the searched strides (no TMA then: planes are not 128-byte aligned)
#endif
    if (g.use_reg == 1 && !NFFT_REG_OLD_STRIDES) {
        // 3D register-stencil kernels: rows dense (a TMA box plane is P0 x P1 floats, row pitch P0), planes
        // 128-byte aligned (TMA shared-memory address).  Bank conflicts of the sweep's (x, y)-position accesses
        // are the same for sY = 28 as for the previous 29 (1.83 wavefronts per access at L = 10).
        g.sZ = (g.P[0] * g.P[1] + 31) / 32 * 32;
        return;
    }
    std::lock_guard<std::mutex> lock(g_stride_mutex);
    const StrideKey key{g.dim, g.L, g.P[0], g.P[1], g.P[2]};
    auto it = g_strides.find(key);
    if (it == g_strides.end()) {
        double best = 1e300;
        int bY = g.sY, bZ = g.sZ;
        for (int sy = g.P[0]; sy < g.P[0] + 32; ++sy) {
            if (g.dim == 2) {
                const double cost = (double)sy * g.P[1] * (1.0 + conflict_score(2, g.L, sy, 0));
                if (cost < best) { best = cost; bY = sy; bZ = sy * g.P[1]; }
            } else {
                for (int pad = 0; pad < 32; ++pad) {
                    const int sz = sy * g.P[1] + pad;
                    const double cost = (double)sz * g.P[2] * (1.0 + conflict_score(3, g.L, sy, sz));
                    if (cost < best) { best = cost; bY = sy; bZ = sz; }
                }
            }
        }
        it = g_strides.emplace(key, std::make_pair(bY, bZ)).first;
    }
    g.sY = it->second.first;
    g.sZ = it->second.second;
}

// SM count of the current device (148 on B200), cached per device; 148 without a usable device (host-only
// size queries)
static int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return 148;
    }
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
        (void)cudaGetLastError();
        return 148;
    }
    cached[dev].store(v, std::memory_order_relaxed);
    return v;
}

// Mixed-density mode of the 3D register-stencil path (make_geom): -1 = default (on for calls that carry
// NFFTB200_CLUSTERED, or all calls with the environment variable NFFTB200_MIXED; never with NFFTB200_NO_MIXED /
// NFFTB200_NO_DENSE), 0 = off, 1 = on; point sets below g_mixed_min_points never use it (the keys are sampled and
// a conditional radix pass is launched: not worth it for small sets).  Test hook: nfftb200_debug_mixed.
static std::atomic<int> g_mixed_mode{-1};
static std::atomic<long long> g_mixed_min_points{1ll << 18};
static std::atomic<int> g_mixed_dense_pts{0};  // 0: kDenseTilePts / NFFTB200_DENSE_TILE_PTS
constexpr int kDenseTilePts = 2048;  // half a point per oversampled cell of a 16^3 tile (profiles/r02r_ab.txt)

// gflags: bit 0 = the grid is complex, bit 1 = the caller expects CLUSTERED points (NFFTB200_CLUSTERED)
static inline int geom_flags(bool grid_cplx, int flags) { return (grid_cplx ? 1 : 0) | ((flags & NFFTB200_CLUSTERED) ? 2 : 0); }
static int make_geom(Geom& g, int d, int64_t N, int m, int64_t B, int64_t C, int gflags, int64_t n_points) {
    const bool cplx = gflags & 1, clustered_hint = gflags & 2;
    if (d < 1 || d > 3) NF_FAIL(NFFTB200_ERR_INVALID, "dimension d=%d not in [1,3]", d);
    if (N < 2 || (N & 1)) NF_FAIL(NFFTB200_ERR_INVALID, "bandwidth N=%lld must be even and >= 2", (long long)N);
    if (m < 1 || m > kMaxCutoff) NF_FAIL(NFFTB200_ERR_INVALID, "cutoff m=%d not in [1,%d]", m, kMaxCutoff);
    if (B < 1 || C < 1) NF_FAIL(NFFTB200_ERR_INVALID, "batch size %lld / columns %lld must be >= 1", (long long)B, (long long)C);
    if (2 * N > (1 << 20)) NF_FAIL(NFFTB200_ERR_INVALID, "bandwidth N=%lld too large", (long long)N);
    // the permutation, bin offsets and work items are 32-bit, launch grids are unsigned
    if (n_points < 0 || n_points >= (1ll << 32) - 1)
        NF_FAIL(NFFTB200_ERR_INVALID, "number of points %lld not in [0, 2^32 - 1)", (long long)n_points);
    g = Geom{};
    g.dim = d;
    g.N = (int)N;
    g.M = (int)(2 * N);
    g.m = m;
    g.L = 2 * m + 2;
    g.LP = (g.L + 3) / 4 * 4;
    g.B = (int)B;
    g.C = (int)C;
    g.cplx = cplx ? 1 : 0;
    g.K = (int)C * (cplx ? 2 : 1);
    g.Md = 1;
    for (int a = 0; a < d; ++a) g.Md *= g.M;
    // window parameters, evaluated like the reference's host macros
    g.inv_b = kThreeQuarterPi / (float)m;                 // WINDOW_FORWARD_PARAM1
    g.inv_sqrt_b_pi = sqrtf(0.75f / (float)m);            // WINDOW_FORWARD_PARAM2
    for (int j = 0; j < kMaxCutoff + 2; ++j) g.kexp[j] = (float)exp(-(double)(j * j) * (double)g.inv_b);
    g.c_hat = kPiThird * (float)m / (float)(N * N);       // WINDOW_ADJOINT_PARAM

    // components per pass and tile extents
    int ncomp = 1;
    const int maxc = d == 3 ? 2 : 8;
    while (ncomp * 2 <= g.K && ncomp * 2 <= maxc) ncomp *= 2;
    if (cplx && ncomp < 2) ncomp = 2;
    // register-stencil kernels: 3D, m <= 4, one float component per pass (complex: re, im passes)
    static const bool no_reg = getenv("NFFTB200_NO_REG") != nullptr;
    g.use_reg = (d == 3 && m <= 4 && !no_reg) ? 1 : 0;
    if (g.use_reg) ncomp = 1;
    // 2D register-stencil kernels: m = 3 or 4, up to 8 float components per pass
    if (d == 2 && (m == 3 || m == 4) && !no_reg) g.use_reg = 2;
    // 1D: cell-owner spread / point-owner gather (window1d.cuh), any m, real and complex
    if (d == 1 && !no_reg) g.use_reg = 3;
    g.ncomp = ncomp;
    int T[3] = {1, 1, 1};
    if (d == 1) {
        T[0] = 512;
    } else if (d == 2 && g.use_reg == 2) {
        T[0] = 32;
        T[1] = 16;
    } else if (d == 2) {
        T[0] = ncomp >= 8 ? 32 : 64;
        T[1] = ncomp >= 4 ? 32 : 64;
    } else {
        T[0] = 16;
        T[1] = ncomp >= 2 ? 8 : 16;
        T[2] = ncomp >= 2 ? 8 : 16;
    }
    for (int s = 0; s < 3; ++s) {
        if (s >= d) {
            g.T[s] = 1; g.nt[s] = 1; g.P[s] = 1; g.org[s] = 0;
            continue;
        }
        g.T[s] = T[s] < g.M ? T[s] : g.M;
        g.nt[s] = (g.M + g.T[s] - 1) / g.T[s];
        g.org[s] = m;
        g.P[s] = g.T[s] + g.L - 1;
    }
    // X: origin and extent aligned to 4 cells (16-byte vector flush / stage)
    g.org[0] = (m + 3) / 4 * 4;
    g.P[0] = ((g.org[0] - m) + g.T[0] + g.L - 1 + 3) / 4 * 4;
    g.tiles_per_batch = g.nt[0] * g.nt[1] * g.nt[2];
    if ((long long)g.tiles_per_batch * B >= (1ll << 31)) NF_FAIL(NFFTB200_ERR_INVALID, "too many tiles");

    // Supercells and fine sort-key bits of the 3D register-stencil kernels (full 16^3 tiles only).
    // * A dense point set (>= 1 point per oversampled cell on average, known on the host: BASELINE c5) is swept
    //   with 2 x 2 x 2 supercells: the register block shrinks from 13 x 13 to 11 x 11 positions (4 instead of 6
    //   per lane, 24 instead of 36 FFMA2 per point); its add-outs, 4x as frequent per cell, are amortised over
    //   the many points of a supercell.  Everything else uses 4 x 4 x 2.
    // * Fine key bits (sort.cuh: fine_index): the bits that fit the radix passes the tile key needs anyway are
    //   free; a dense set gets at least 4 of them even if that adds a pass, because its tiles are cut into many
    //   chunks and compact chunks are what lets the points of a chunk share register blocks.
    // * Mixed density (large point sets the caller marks NFFTB200_CLUSTERED; Python: NfftPlan(clustered=True)): which
    //   tiles are heavy is not known on the host, so it is decided on the device.  The hint only enables the
    //   machinery -- it costs a uniform set ~1.5 % (measured, profiles/r02r_ab.txt), which is why it is not the
    //   default.  The keys carry the 2 x 2 x 2 hierarchy, laid out so
    //   that the lowest radix pass covers fine bits only; that pass runs only when a sample of the keys finds a
    //   clustered set (sort.cuh: sort_points), and the work items of heavy tiles (>= dense_tile_pts points) are
    //   then marked for the 2 x 2 x 2 sweep: launch_window starts both sweeps, each takes its class of tiles.
    g.fine_bits = g.fine_xy_levels = g.fine_z_bits = 0;
    g.sc[0] = g.sc[1] = g.sc[2] = 1;
    g.mixed = g.refine_pass = g.dense_tile_pts = 0;
    static const bool no_fine = getenv("NFFTB200_NO_FINE_SORT") != nullptr;
    static const bool no_dense = getenv("NFFTB200_NO_DENSE") != nullptr;
    static const bool no_mixed = getenv("NFFTB200_NO_MIXED") != nullptr;
    static const bool env_mixed = getenv("NFFTB200_MIXED") != nullptr;  // experiments: as if every call carried the hint
    static const int env_dense_pts = getenv("NFFTB200_DENSE_TILE_PTS") ? atoi(getenv("NFFTB200_DENSE_TILE_PTS")) : 0;
    if (g.use_reg == 1) {
        const bool full_tiles = g.T[0] == 16 && g.T[1] == 16 && g.T[2] == 16;
        const bool dense = (double)n_points >= (double)B * (double)g.Md;
        g.sc[0] = kRegSX, g.sc[1] = kRegSY, g.sc[2] = kRegSZ;
        const bool small_kernels = full_tiles && (m == 3 || m == 4);  // spread/gather_reg_kernel<L, 2, 2, 2> exist
        if (dense && small_kernels && !no_dense) g.sc[0] = g.sc[1] = g.sc[2] = 2;
        int tile_bits = 0;
        while ((1ll << tile_bits) < (long long)g.tiles_per_batch * B) ++tile_bits;
        const int spare = (tile_bits + 7) / 8 * 8 - tile_bits;
        const int mixed_mode = g_mixed_mode.load();
        const long long mixed_min = g_mixed_min_points.load();
        if (g.sc[0] != 2 && small_kernels && !no_fine && tile_bits > 0 && n_points >= mixed_min &&
            (mixed_mode == 1 || (mixed_mode < 0 && (clustered_hint || env_mixed) && !no_mixed && !no_dense))) {
            const int total = 9;  // 2 x 2 x 2 supercells in a 16^3 tile: 3 (y, x) levels + 3 z bits
            const int fb = spare >= total ? total : spare + 8;
            if (tile_bits + fb <= 31) {
                g.mixed = 1;
                g.refine_pass = spare < total;
                const int hook_pts = g_mixed_dense_pts.load();
                g.dense_tile_pts = hook_pts > 0 ? hook_pts : (env_dense_pts > 0 ? env_dense_pts : kDenseTilePts);
                g.sc[0] = g.sc[1] = g.sc[2] = 2;  // (of the key hierarchy; the sweep is chosen per tile)
                g.fine_xy_levels = g.fine_z_bits = 3;
                g.fine_bits = fb;
            }
        }
        auto log2i = [](int v) { int b = 0; while ((1 << b) < v) ++b; return b; };
        const bool pow2_cells = g.sc[0] == g.sc[1] && (g.sc[0] & (g.sc[0] - 1)) == 0 && (g.sc[2] & (g.sc[2] - 1)) == 0;
        if (!g.mixed && full_tiles && pow2_cells && !no_fine) {
            g.fine_xy_levels = log2i(16 / g.sc[0]);
            g.fine_z_bits = log2i(16 / g.sc[2]);
            const int total = 2 * g.fine_xy_levels + g.fine_z_bits;
            int k = spare < total ? spare : total;
            // a dense set whose tile key leaves fewer than 4 spare bits pays one more radix pass for compact chunks
            // (c5: 9 tile bits, 7 spare: its 2^26-point sort costs 2.6 ms per pass, profiles/r02f_c5.txt)
            if (dense && k < 4) k = spare + 8 < total ? spare + 8 : total;
            if (tile_bits == 0 && !dense) k = 0;  // a single tile and few points: no radix pass at all
            if (k > 31 - tile_bits) k = 31 - tile_bits;
            g.fine_bits = k < 0 ? 0 : k;
        }
    }

    const int team = d == 1 ? g.L : g.L * g.L;
    choose_strides(g);
    long long te = d == 1 ? g.P[0] : (d == 2 ? (long long)g.sY * g.P[1] : (long long)g.sZ * g.P[2]);
    g.tile_elems = (int)((te + 3) / 4 * 4);

    long long pm = n_points / (sm_count() * 8);
    g.pmax = (int)(pm < 256 ? 256 : (pm > 2048 ? 2048 : pm));
    if (g.use_reg == 1) g.pmax = kRegMaxPts;
    if (g.use_reg == 2) g.pmax = kReg2MaxPts;
    if (g.use_reg == 3) {
        pm = n_points / (sm_count() * 4);
        g.pmax = (int)(pm < 512 ? 512 : (pm > kW1MaxPts ? kW1MaxPts : pm));
    }
    int threads = (team + 31) / 32 * 32;
    g.spread_threads = threads < 64 ? 64 : threads;
    // work items (one CTA each) and radix blocks must fit a launch grid
    const long long nbins = (long long)g.B * g.tiles_per_batch;
    if (n_points / g.pmax + (nbins < n_points ? nbins : n_points) + 1 >= (1ll << 31))
        NF_FAIL(NFFTB200_ERR_INVALID, "too many work items for one launch");
    return NFFTB200_OK;
}

static size_t spread_smem_bytes(const Geom& g, int ncomp) {
    return (size_t)ncomp * g.tile_elems * 4 + (size_t)kSubBatch * (g.dim * g.LP * 4 + ncomp * 4 + 32);
}
static size_t gather_smem_bytes(const Geom& g, int ncomp) {
    return (size_t)ncomp * g.tile_elems * 4 + (size_t)kSubBatch * (g.dim * g.LP * 4 + 32);
}

// ----------------------------------------------------------------------------------------
// kernel dispatch tables
// ----------------------------------------------------------------------------------------
typedef void (*WindowKernel)(const Geom, const WindowArgs);

template <int DIM, int LC>
static WindowKernel pick_spread(int ncomp) {
    switch (ncomp) {
        case 1: return spread_kernel<DIM, 1, LC>;
        case 2: return spread_kernel<DIM, 2, LC>;
        case 4: return spread_kernel<DIM, 4, LC>;
        default: return spread_kernel<DIM, 8, LC>;
    }
}
template <int DIM, int LC>
static WindowKernel pick_gather(int ncomp) {
    switch (ncomp) {
        case 1: return gather_kernel<DIM, 1, LC>;
        case 2: return gather_kernel<DIM, 2, LC>;
        case 4: return gather_kernel<DIM, 4, LC>;
        default: return gather_kernel<DIM, 8, LC>;
    }
}
static WindowKernel get_spread(int dim, int ncomp, int L) {
    if (dim == 3) return L == 10 ? pick_spread<3, 10>(ncomp) : (L == 8 ? pick_spread<3, 8>(ncomp) : pick_spread<3, 0>(ncomp));
    if (dim == 2) return pick_spread<2, 0>(ncomp);
    return pick_spread<1, 0>(ncomp);
}
static WindowKernel get_gather(int dim, int ncomp, int L) {
    if (dim == 3) return L == 10 ? pick_gather<3, 10>(ncomp) : (L == 8 ? pick_gather<3, 8>(ncomp) : pick_gather<3, 0>(ncomp));
    if (dim == 2) return pick_gather<2, 0>(ncomp);
    return pick_gather<1, 0>(ncomp);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel) and size, not per launch
static std::mutex g_attr_mutex;
static std::map<std::pair<int, const void*>, size_t> g_attr_smem;
// threads > 0: the kernel is built for TWO resident CTAs per SM (the 3D register-stencil sweeps); the smallest
// number the runtime reports for any such launch configuration is kept for nfftb200_debug_min_resident_ctas --
// the sweeps sit within a few hundred bytes of the shared-memory limit and losing the second CTA costs 45 %.
static std::atomic<int> g_min_resident{1 << 30};
static int ensure_dynamic_smem(const void* kern, size_t smem, int threads = 0) {
    if (smem > 227 * 1024) NF_FAIL(NFFTB200_ERR_INVALID, "tile needs %zu bytes of shared memory", smem);
    int dev = 0;
    NF_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    size_t& have = g_attr_smem[std::make_pair(dev, kern)];
    if (smem > have) {
        NF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
        if (threads > 0) {
            int ctas = 0;
            NF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, threads, smem));
            if (ctas < g_min_resident.load()) g_min_resident.store(ctas);
        }
    }
    return NFFTB200_OK;
}

// ----------------------------------------------------------------------------------------
// TMA tensor map of the real oversampled grid (window_reg.cuh).  cuTensorMapEncodeTiled is a driver entry
// point: it is looked up through the runtime (no link dependency on libcuda).
// ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static std::mutex mu;
    static bool looked_up = false;
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!looked_up) {
        looked_up = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    }
    return fn;
}

// true (and *map filled) if the 3D register-stencil kernels can move this grid's tile planes by TMA
static bool make_grid_tensor_map(const Geom& g, float* grid, CUtensorMap* map) {
    static const bool disabled = getenv("NFFTB200_NO_TMA") != nullptr;
    memset(map, 0, sizeof(*map));
    if (disabled || g.use_reg != 1 || g.cplx || g.dim != 3) return false;
    if (g.P[0] > g.M || g.P[1] > g.M || g.P[2] > g.M) return false;  // tiny grids: a tile wraps more than once
    if (g.P[0] > 256 || g.P[1] > 256 || (g.P[0] * 4) % 16 != 0 || (g.sZ * 4) % 128 != 0 || g.sY != g.P[0]) return false;
    if (((uintptr_t)grid & 15) != 0 || (g.M * 4) % 16 != 0) return false;
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t M = (cuuint64_t)g.M;
    const cuuint64_t dims[4] = {M, M, M, (cuuint64_t)g.B * g.C};
    const cuuint64_t strides[3] = {M * 4, M * M * 4, M * M * M * 4};
    const cuuint32_t box[4] = {(cuuint32_t)g.P[0], (cuuint32_t)g.P[1], 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, grid, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

typedef void (*WindowKernelTma)(const Geom, const WindowArgs, const CUtensorMap);

static int launch_window(bool spread, const Geom& g, WindowArgs a, const SortPlan& sp, cudaStream_t st) {
    a.perm = sp.perm;
    a.bin_start = sp.bin_start;
    a.chunk_start = sp.chunk_start;
    a.items = sp.items;
    a.nbins = sp.nbins;
    a.flags = sp.flags;
    if (g.use_reg == 2) {
        // 2D: supercell 4 x 4 cells, up to 8 channels per pass (remaining channels in smaller passes)
        int ncomp = g.ncomp;
        for (int k0 = 0; k0 < g.K; k0 += ncomp) {
            ncomp = g.ncomp;
            while (ncomp > g.K - k0) ncomp >>= 1;
            a.k0 = k0;
            WindowKernel kern = nullptr;
            int win_floats = 0;
#define NF_REG2_CASE(L_, C_)                                                                              \
            if (g.L == L_ && ncomp == C_) {                                                               \
                kern = spread ? spread_reg2d_kernel<L_, C_> : gather_reg2d_kernel<L_, C_>;                \
                win_floats = Reg2Cfg<L_, C_>::WIN_FLOATS;                                                 \
            }
            NF_REG2_CASE(8, 1) NF_REG2_CASE(8, 2) NF_REG2_CASE(8, 4) NF_REG2_CASE(8, 8)
            NF_REG2_CASE(10, 1) NF_REG2_CASE(10, 2) NF_REG2_CASE(10, 4) NF_REG2_CASE(10, 8)
#undef NF_REG2_CASE
            if (!kern) NF_FAIL(NFFTB200_ERR_INVALID, "no 2D register-stencil kernel for L=%d ncomp=%d", g.L, ncomp);
            const int nsc = ((g.T[0] + 3) / 4) * ((g.T[1] + 3) / 4);
            const size_t smem = reg2_smem_bytes(g, ncomp, spread, nsc, win_floats);
            NF_TRY(ensure_dynamic_smem((const void*)kern, smem));
            NF_LAUNCH(kern, (unsigned)sp.max_items, kReg2Threads, smem, st, g, a);
        }
        return NFFTB200_OK;
    }
    if (g.use_reg == 3) {
        const bool pow2 = (g.M & (g.M - 1)) == 0;
        int ncomp = g.ncomp;
        for (int k0 = 0; k0 < g.K; k0 += ncomp) {
            ncomp = g.ncomp;
            while (ncomp > g.K - k0) ncomp >>= 1;
            if (g.cplx && ncomp < 2) ncomp = 2;
            a.k0 = k0;
            WindowKernel kern = nullptr;
#define NF_W1_CASE(C_)                                                                                     \
            if (ncomp == C_)                                                                               \
                kern = spread ? (pow2 ? spread1d_kernel<C_, true> : spread1d_kernel<C_, false>)            \
                              : (pow2 ? gather1d_kernel<C_, true> : gather1d_kernel<C_, false>);
            NF_W1_CASE(1) NF_W1_CASE(2) NF_W1_CASE(4) NF_W1_CASE(8)
#undef NF_W1_CASE
            if (!kern) NF_FAIL(NFFTB200_ERR_INVALID, "no 1D kernel for ncomp=%d", ncomp);
            const size_t smem = spread ? w1_spread_smem_bytes(g, ncomp) : w1_gather_smem_bytes(g, ncomp);
            NF_TRY(ensure_dynamic_smem((const void*)kern, smem));
            NF_LAUNCH(kern, (unsigned)sp.max_items, spread ? kW1SpreadThreads : kW1Threads, smem, st, g, a);
        }
        return NFFTB200_OK;
    }
    if (g.use_reg == 1) {
        // supercell kRegSX x kRegSY x kRegSZ = 4 x 4 x 2 cells; for m = 4 the register block is 13 x 13 x 12:
        // 6 positions x 6 float2 accumulators per lane
        WindowKernelTma kern = nullptr;
        int win_floats = 0;
        // 2 x 2 x 2 supercells: dense point sets (m = 3, 4).  Mixed density: ONE launch whose CTAs choose the sweep
        // by the class of their work item (spread/gather_reg_mixed_kernel); shared memory for the larger of the two.
        const bool small_cells = g.sc[0] == 2 && g.sc[1] == 2 && g.sc[2] == 2;
        const int scx = small_cells ? 2 : kRegSX, scy = small_cells ? 2 : kRegSY, scz = small_cells ? 2 : kRegSZ;
        switch (g.m) {
#define NF_REG_CASE(M_, L_)                                                                          \
            case M_:                                                                                 \
                kern = spread ? spread_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>                        \
                              : gather_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>;                       \
                win_floats = RegCfg<L_, kRegSX, kRegSY, kRegSZ>::WIN_FLOATS;                         \
                break;
#define NF_REG_CASE_DENSE(M_, L_)                                                                    \
            case M_:                                                                                 \
                if (g.mixed) {                                                                       \
                    kern = spread ? spread_reg_mixed_kernel<L_> : gather_reg_mixed_kernel<L_>;       \
                    win_floats = RegCfg<L_, 2, 2, 2>::WIN_FLOATS;                                    \
                    if (win_floats < RegCfg<L_, kRegSX, kRegSY, kRegSZ>::WIN_FLOATS)                 \
                        win_floats = RegCfg<L_, kRegSX, kRegSY, kRegSZ>::WIN_FLOATS;                 \
                } else if (small_cells) {                                                            \
                    kern = spread ? spread_reg_kernel<L_, 2, 2, 2> : gather_reg_kernel<L_, 2, 2, 2>; \
                    win_floats = RegCfg<L_, 2, 2, 2>::WIN_FLOATS;                                    \
                } else {                                                                             \
                    kern = spread ? spread_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>                    \
                                  : gather_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>;                   \
                    win_floats = RegCfg<L_, kRegSX, kRegSY, kRegSZ>::WIN_FLOATS;                     \
                }                                                                                    \
                break;
            NF_REG_CASE(1, 4)
            NF_REG_CASE(2, 6)
            NF_REG_CASE_DENSE(3, 8)
            NF_REG_CASE_DENSE(4, 10)
#undef NF_REG_CASE
#undef NF_REG_CASE_DENSE
            default: NF_FAIL(NFFTB200_ERR_INVALID, "register-stencil kernels need m <= 4");
        }
        const int nsc = ((g.T[0] + scx - 1) / scx) * ((g.T[1] + scy - 1) / scy) * ((g.T[2] + scz - 1) / scz);
        // experiment switch: NFFTB200_SMEM_PAD=<bytes> requests more shared memory per CTA (fewer resident CTAs)
        static const size_t smem_pad = getenv("NFFTB200_SMEM_PAD") ? (size_t)atoll(getenv("NFFTB200_SMEM_PAD")) : 0;
        const size_t smem = reg_smem_bytes(g, nsc, win_floats) + smem_pad;
        NF_TRY(ensure_dynamic_smem((const void*)kern, smem, smem_pad ? 0 : kRegThreads));
        CUtensorMap tmap;
        a.use_tma = make_grid_tensor_map(g, a.grid, &tmap) ? 1 : 0;
        for (int k0 = 0; k0 < g.K; ++k0) {
            a.k0 = k0;
            NF_LAUNCH(kern, (unsigned)sp.max_items, kRegThreads, smem, st, g, a, tmap);
        }
        return NFFTB200_OK;
    }
    // components are processed in passes of <= g.ncomp (real: channels, complex: re/im pairs)
    int ncomp = g.ncomp;
    for (int k0 = 0; k0 < g.K; k0 += ncomp) {
        ncomp = g.ncomp;
        while (ncomp > g.K - k0) ncomp >>= 1;
        a.k0 = k0;
        WindowKernel kern = spread ? get_spread(g.dim, ncomp, g.L) : get_gather(g.dim, ncomp, g.L);
        const size_t smem = spread ? spread_smem_bytes(g, ncomp) : gather_smem_bytes(g, ncomp);
        NF_TRY(ensure_dynamic_smem((const void*)kern, smem));
        const unsigned grid = (unsigned)sp.max_items;
        const unsigned block = spread ? (unsigned)g.spread_threads : (unsigned)kGatherThreads;
        NF_LAUNCH(kern, grid, block, smem, st, g, a);
    }
    return NFFTB200_OK;
}

// ----------------------------------------------------------------------------------------
// cuFFT plan cache (the reference creates and destroys a plan per call, core_cuda.cu:254-272)
//
// Plans are made with cufftSetAutoAllocation(0): their work area is NOT owned by cuFFT but handed in per
// call from the caller's workspace (cufftSetWorkArea), so two streams (or an eager call next to a graph
// replay) never share scratch, captured graphs only reference caller-owned memory plus the plan's own
// twiddle tables, and the memory shows up in the caller's allocator.  The handle state (stream, work area)
// is set and the transform enqueued under g_plan_mutex.  The cache is an LRU of kMaxPlans handles;
// nfftb200_plan_cache_pin() forbids destroying handles while captured graphs may still use their tables.
// ----------------------------------------------------------------------------------------
struct PlanKey {
    int dev, dim, M, type;
    long long batch;
    bool operator<(const PlanKey& o) const {
        return std::tie(dev, dim, M, type, batch) < std::tie(o.dev, o.dim, o.M, o.type, o.batch);
    }
};
struct PlanEntry {
    cufftHandle handle;
    size_t work;
    unsigned long long tick;
};
constexpr size_t kMaxPlans = 24;
static std::mutex g_plan_mutex;
static std::map<PlanKey, PlanEntry> g_plans;
static unsigned long long g_plan_tick = 0;
static int g_plan_pins = 0;

// caller holds g_plan_mutex
static int get_plan_locked(int dim, int M, long long batch, cufftType type, PlanEntry** out) {
    int dev = 0;
    NF_CUDA(cudaGetDevice(&dev));
    PlanKey key{dev, dim, M, (int)type, batch};
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        it->second.tick = ++g_plan_tick;
        *out = &it->second;
        return NFFTB200_OK;
    }
    if (g_plans.size() >= kMaxPlans && g_plan_pins == 0) {  // evict the least recently used handle
        auto lru = g_plans.begin();
        for (auto jt = g_plans.begin(); jt != g_plans.end(); ++jt)
            if (jt->second.tick < lru->second.tick) lru = jt;
        cufftDestroy(lru->second.handle);
        g_plans.erase(lru);
    }
    cufftHandle plan;
    NF_CUFFT(cufftCreate(&plan));
    cufftResult r = cufftSetAutoAllocation(plan, 0);
    long long n[3] = {M, M, M};
    long long real_dist = 1, half_dist = 1;
    for (int a = 0; a < dim; ++a) real_dist *= M;
    for (int a = 0; a < dim - 1; ++a) half_dist *= M;
    half_dist *= (M / 2 + 1);
    long long idist = real_dist, odist = real_dist;
    if (type == CUFFT_R2C) odist = half_dist;
    if (type == CUFFT_C2R) idist = half_dist;
    size_t work = 0;
    if (r == CUFFT_SUCCESS)
        r = cufftMakePlanMany64(plan, dim, n, nullptr, 1, idist, nullptr, 1, odist, type, batch, &work);
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(plan);
        NF_FAIL(NFFTB200_ERR_CUFFT, "cufftMakePlanMany64(dim=%d, M=%d, batch=%lld, type=%d) -> %d", dim, M, batch,
                (int)type, (int)r);
    }
    PlanEntry e{plan, work, ++g_plan_tick};
    *out = &g_plans.emplace(key, e).first->second;
    return NFFTB200_OK;
}

// bytes of cuFFT work area the transforms of one op need (the maximum over the plans it runs); 0 when no
// device is usable (host-only size queries): the op itself then fails with a clear message
// Pruned real transforms (fft_rows.cuh): hand-written X pass + cuFFT C2C over the other d - 1 dimensions of the
// kept kx planes.  -1 = default (on unless NFFTB200_NO_PRUNED_FFT is set), 0 = off, 1 = on; test hook
// nfftb200_debug_pruned_fft.  The fastsum keeps the plain R2C / C2R pair (its spectrum is not cropped in between).
static std::atomic<int> g_pruned_mode{-1};
static bool pruned_fft_ok(const Geom& g) {
    static const bool env_off = getenv("NFFTB200_NO_PRUNED_FFT") != nullptr;
    const int mode = g_pruned_mode.load();
    if (mode == 0 || (mode < 0 && env_off)) return false;
    // (2D and 3D real transforms on 256- and 512-cell grids: c4 FFT stages 0.71 -> 0.40 ms, c3 0.29 -> 0.20 ms,
    // profiles/r02x_ab.txt, r02y_ab.txt)
    return (g.dim == 2 || g.dim == 3) && (g.M == 256 || g.M == 512) && g.N == g.M / 2;
}
static long long pruned_batch(const Geom& g) { return (long long)g.B * g.C * (g.N / 2 + 1); }

static size_t fft_work_bytes(const Geom& g, bool grid_cplx) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    size_t need = 0;
    const long long batch = (long long)g.B * g.C;
    PlanEntry* e = nullptr;
    if (grid_cplx) {
        if (get_plan_locked(g.dim, g.M, batch, CUFFT_C2C, &e) == NFFTB200_OK) need = e->work;
    } else {
        if (get_plan_locked(g.dim, g.M, batch, CUFFT_R2C, &e) == NFFTB200_OK) need = e->work;
        if (get_plan_locked(g.dim, g.M, batch, CUFFT_C2R, &e) == NFFTB200_OK && e->work > need) need = e->work;
        if (pruned_fft_ok(g) && get_plan_locked(g.dim - 1, g.M, pruned_batch(g), CUFFT_C2C, &e) == NFFTB200_OK &&
            e->work > need)
            need = e->work;
    }
    (void)cudaGetLastError();
    if (e == nullptr) g_err[0] = 0;  // no usable device: not an error of a size query
    return align_up(need);
}

struct FftWork {
    void* ptr;
    size_t bytes;
};

enum FftKind { FFT_R2C, FFT_C2R, FFT_C2C_INVERSE, FFT_C2C_FORWARD };

// one batched transform (dim dimensions of M cells, `batch` contiguous arrays) on the caller's stream with the
// caller's work area
static int fft_exec_dims(int dim, int M, long long batch, FftKind kind, void* in, void* out, const FftWork& work,
                         cudaStream_t st) {
    const cufftType type = kind == FFT_R2C ? CUFFT_R2C : (kind == FFT_C2R ? CUFFT_C2R : CUFFT_C2C);
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    PlanEntry* e = nullptr;
    NF_TRY(get_plan_locked(dim, M, batch, type, &e));
    if (e->work > work.bytes)
        NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small for the cuFFT work area: %zu < %zu", work.bytes, e->work);
    NF_CUFFT(cufftSetStream(e->handle, st));
    if (e->work > 0) NF_CUFFT(cufftSetWorkArea(e->handle, work.ptr));
    ProfScope ps(ST_FFT, st);
    switch (kind) {
        case FFT_R2C: NF_CUFFT(cufftExecR2C(e->handle, (cufftReal*)in, (cufftComplex*)out)); break;
        case FFT_C2R: NF_CUFFT(cufftExecC2R(e->handle, (cufftComplex*)in, (cufftReal*)out)); break;
        case FFT_C2C_INVERSE:  // sign +, core_cuda.cu:267
            NF_CUFFT(cufftExecC2C(e->handle, (cufftComplex*)in, (cufftComplex*)out, CUFFT_INVERSE));
            break;
        default:               // sign -, core_cuda.cu:445
            NF_CUFFT(cufftExecC2C(e->handle, (cufftComplex*)in, (cufftComplex*)out, CUFFT_FORWARD));
            break;
    }
    return NFFTB200_OK;
}
// one transform of all B*C grids
static int fft_exec(const Geom& g, FftKind kind, void* in, void* out, const FftWork& work, cudaStream_t st) {
    return fft_exec_dims(g.dim, g.M, (long long)g.B * g.C, kind, in, out, work, st);
}

// ----------------------------------------------------------------------------------------
// workspace layout
// ----------------------------------------------------------------------------------------
static size_t grid_bytes(const Geom& g, bool cplx) { return align_up((size_t)g.B * g.C * g.Md * (cplx ? 8 : 4)); }
static size_t half_bytes(const Geom& g) {
    size_t h = (size_t)g.B * g.C * (g.M / 2 + 1);
    for (int a = 0; a < g.dim - 1; ++a) h *= g.M;
    return align_up(h * 8);
}
static long long half_elems(const Geom& g) {
    long long h = (long long)g.B * g.C * (g.M / 2 + 1);
    for (int a = 0; a < g.dim - 1; ++a) h *= g.M;
    return h;
}

static unsigned blocks_for(long long total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }

// ----------------------------------------------------------------------------------------
// stage implementations
// ----------------------------------------------------------------------------------------
static int do_spread(const Geom& g, const float* pos, const float* x, float* grid, long long n, const SortPlan& sp,
                     cudaStream_t st) {
    {
        ProfScope ps(ST_MEMSET, st);
        NF_CUDA(cudaMemsetAsync(grid, 0, (size_t)g.B * g.C * g.Md * (g.cplx ? 8 : 4), st));
    }
    if (n == 0) return NFFTB200_OK;
    WindowArgs a{};
    a.pos = pos;
    a.xin = x;
    a.grid = grid;
    ProfScope ps(ST_SPREAD, st);
    return launch_window(true, g, a, sp, st);
}

static int do_gather(const Geom& g, const float* pos, const float* grid, float* y, long long n, const SortPlan& sp,
                     cudaStream_t st) {
    if (n == 0) return NFFTB200_OK;
    WindowArgs a{};
    a.pos = pos;
    a.yout = y;
    a.grid = const_cast<float*>(grid);
    ProfScope ps(ST_GATHER, st);
    return launch_window(false, g, a, sp, st);
}

static int do_sort(const Geom& g, const float* pos, const int64_t* batch, bool offsets, long long n, char* scratch,
                   char* plan_mem, SortPlan* sp, cudaStream_t st) {
    ProfScope ps(ST_SORT, st);
    return sort_points(pos, batch, offsets, n, g, scratch, plan_mem, sp, st);
}

// spectral kernels use 32-bit index arithmetic whenever every element index fits 31 bits
template <int DIM, typename I>
static int launch_unpack_t(const Geom& g, bool half, bool real_out, const float2* spec, float* y, cudaStream_t st,
                           bool pruned) {
    long long total = (long long)g.B * g.C;
    for (int a = 0; a < DIM; ++a) total *= g.N;
    const unsigned grid = blocks_for(total);
    if (half && pruned) {
        if (real_out) NF_LAUNCH((unpack_kernel<DIM, true, true, I, true>), grid, 256, 0, st, spec, y, g);
        else NF_LAUNCH((unpack_kernel<DIM, true, false, I, true>), grid, 256, 0, st, spec, y, g);
    } else if (half) {
        if (real_out) NF_LAUNCH((unpack_kernel<DIM, true, true, I>), grid, 256, 0, st, spec, y, g);
        else NF_LAUNCH((unpack_kernel<DIM, true, false, I>), grid, 256, 0, st, spec, y, g);
    } else {
        if (real_out) NF_LAUNCH((unpack_kernel<DIM, false, true, I>), grid, 256, 0, st, spec, y, g);
        else NF_LAUNCH((unpack_kernel<DIM, false, false, I>), grid, 256, 0, st, spec, y, g);
    }
    return NFFTB200_OK;
}

template <int DIM, typename I>
static int launch_pack_t(const Geom& g, bool half, bool xreal, const float* xhat, float2* spec, cudaStream_t st,
                         bool pruned) {
    // zero fill (out-of-band 7/8 of a 3D spectrum) with one memset, then write the band box
    long long total = half ? half_elems(g) : (long long)g.B * g.C * g.Md;
    if (half && pruned) total = pruned_batch(g) * (g.Md / g.M);  // P[bc][kx][M^(d-1)]
    NF_CUDA(cudaMemsetAsync(spec, 0, (size_t)total * sizeof(float2), st));
    long long band = (long long)g.B * g.C * (half ? g.N / 2 + 1 : g.N);
    for (int a = 0; a < DIM - 1; ++a) band *= half ? g.N + 1 : g.N;
    const unsigned grid = blocks_for(band);
    if (half && pruned) {
        if (xreal) NF_LAUNCH((pack_kernel<DIM, true, true, I, true>), grid, 256, 0, st, xhat, spec, g);
        else NF_LAUNCH((pack_kernel<DIM, true, false, I, true>), grid, 256, 0, st, xhat, spec, g);
    } else if (half) {
        if (xreal) NF_LAUNCH((pack_kernel<DIM, true, true, I>), grid, 256, 0, st, xhat, spec, g);
        else NF_LAUNCH((pack_kernel<DIM, true, false, I>), grid, 256, 0, st, xhat, spec, g);
    } else {
        if (xreal) NF_LAUNCH((pack_kernel<DIM, false, true, I>), grid, 256, 0, st, xhat, spec, g);
        else NF_LAUNCH((pack_kernel<DIM, false, false, I>), grid, 256, 0, st, xhat, spec, g);
    }
    return NFFTB200_OK;
}

template <int DIM, typename I>
static int launch_multiply_t(const Geom& g, bool half, bool creal, float2* spec, const float* coeffs, cudaStream_t st) {
    const long long total = half ? half_elems(g) : (long long)g.B * g.C * g.Md;
    const unsigned grid = blocks_for(total);
    if (half) {
        if (creal) NF_LAUNCH((kernel_multiply_kernel<DIM, true, true, I>), grid, 256, 0, st, spec, coeffs, g);
        else NF_LAUNCH((kernel_multiply_kernel<DIM, true, false, I>), grid, 256, 0, st, spec, coeffs, g);
    } else {
        if (creal) NF_LAUNCH((kernel_multiply_kernel<DIM, false, true, I>), grid, 256, 0, st, spec, coeffs, g);
        else NF_LAUNCH((kernel_multiply_kernel<DIM, false, false, I>), grid, 256, 0, st, spec, coeffs, g);
    }
    return NFFTB200_OK;
}

// tests: force the 64-bit index variants of the spectral kernels (nfftb200_debug_force_int64)
static std::atomic<int> g_force_int64{0};

static bool fits_int32(const Geom& g) {
    // largest index any spectral kernel forms: complex grid elements plus one block of slack
    if (g_force_int64.load(std::memory_order_relaxed)) return false;
    return (long long)g.B * g.C * g.Md < (1ll << 31) - 1024;
}
template <int DIM>
static int launch_unpack(const Geom& g, bool half, bool real_out, const float2* spec, float* y, cudaStream_t st,
                         bool pruned = false) {
    return fits_int32(g) ? launch_unpack_t<DIM, int>(g, half, real_out, spec, y, st, pruned)
                         : launch_unpack_t<DIM, long long>(g, half, real_out, spec, y, st, pruned);
}
template <int DIM>
static int launch_pack(const Geom& g, bool half, bool xreal, const float* xhat, float2* spec, cudaStream_t st,
                       bool pruned = false) {
    return fits_int32(g) ? launch_pack_t<DIM, int>(g, half, xreal, xhat, spec, st, pruned)
                         : launch_pack_t<DIM, long long>(g, half, xreal, xhat, spec, st, pruned);
}

// the hand-written X pass of the pruned real transforms (fft_rows.cuh)
static int launch_fft_rows(const Geom& g, bool forward_r2c, float* grid, float2* P, cudaStream_t st) {
    const long long rows_per_bc = g.Md / g.M, rows = (long long)g.B * g.C * rows_per_bc;
    const unsigned ctas = (unsigned)(rows / kFftRowsPerCta);
    ProfScope ps(ST_FFT, st);
    if (g.M == 256) {
        const size_t smem = fft_rows_smem_bytes<16>();
        if (forward_r2c) {
            NF_TRY(ensure_dynamic_smem((const void*)rows_r2c_crop_kernel<16>, smem));
            NF_LAUNCH(rows_r2c_crop_kernel<16>, ctas, kFftThreads, smem, st, grid, P, rows_per_bc);
        } else {
            NF_TRY(ensure_dynamic_smem((const void*)rows_c2r_pad_kernel<16>, smem));
            NF_LAUNCH(rows_c2r_pad_kernel<16>, ctas, kFftThreads, smem, st, P, grid, rows_per_bc);
        }
    } else {
        const size_t smem = fft_rows_smem_bytes<32>();
        if (forward_r2c) {
            NF_TRY(ensure_dynamic_smem((const void*)rows_r2c_crop_kernel<32>, smem));
            NF_LAUNCH(rows_r2c_crop_kernel<32>, ctas, kFftThreads, smem, st, grid, P, rows_per_bc);
        } else {
            NF_TRY(ensure_dynamic_smem((const void*)rows_c2r_pad_kernel<32>, smem));
            NF_LAUNCH(rows_c2r_pad_kernel<32>, ctas, kFftThreads, smem, st, P, grid, rows_per_bc);
        }
    }
    return NFFTB200_OK;
}
template <int DIM>
static int launch_multiply(const Geom& g, bool half, bool creal, float2* spec, const float* coeffs, cudaStream_t st) {
    return fits_int32(g) ? launch_multiply_t<DIM, int>(g, half, creal, spec, coeffs, st)
                         : launch_multiply_t<DIM, long long>(g, half, creal, spec, coeffs, st);
}

#define NF_DIM_DISPATCH(fn, ...)                                   \
    (g.dim == 1 ? fn<1>(__VA_ARGS__) : (g.dim == 2 ? fn<2>(__VA_ARGS__) : fn<3>(__VA_ARGS__)))

// grid (real: [BC][M^d] float, complex: [BC][M^d] float2) -> y.  spec: scratch for the half spectrum.
static int do_adjoint_finish(const Geom& g, float* grid, float* y, bool real_out, float2* spec, const FftWork& fw,
                             cudaStream_t st) {
    if (!g.cplx && pruned_fft_ok(g)) {
        // R2C along X with the crop to kx <= N/2, then C2C (sign -) over the other dimensions of the kept planes:
        // P[bc][kx][..] holds what the full R2C transform holds at those frequencies
        NF_TRY(launch_fft_rows(g, true, grid, spec, st));
        NF_TRY(fft_exec_dims(g.dim - 1, g.M, pruned_batch(g), FFT_C2C_FORWARD, spec, spec, fw, st));
        ProfScope ps(ST_UNPACK, st);
        return NF_DIM_DISPATCH(launch_unpack, g, true, real_out, spec, y, st, true);
    }
    if (!g.cplx) {
        NF_TRY(fft_exec(g, FFT_R2C, grid, spec, fw, st));
        ProfScope ps(ST_UNPACK, st);
        return NF_DIM_DISPATCH(launch_unpack, g, true, real_out, spec, y, st);
    }
    NF_TRY(fft_exec(g, FFT_C2C_INVERSE, grid, grid, fw, st));
    ProfScope ps(ST_UNPACK, st);
    return NF_DIM_DISPATCH(launch_unpack, g, false, real_out, reinterpret_cast<const float2*>(grid), y, st);
}

// xhat -> grid.  real_out: grid is float (C2R), else float2 (C2C sign -).
static int do_forward_begin(const Geom& g, const float* xhat, bool xreal, bool real_out, float* grid, float2* spec,
                            const FftWork& fw, cudaStream_t st) {
    if (real_out && pruned_fft_ok(g)) {
        // the C2R input restricted to kx <= N/2 (everything above is zero): C2C (sign +) over the other
        // dimensions of those planes, then the C2R X pass from the kept frequencies
        {
            ProfScope ps(ST_PACK, st);
            NF_TRY(NF_DIM_DISPATCH(launch_pack, g, true, xreal, xhat, spec, st, true));
        }
        NF_TRY(fft_exec_dims(g.dim - 1, g.M, pruned_batch(g), FFT_C2C_INVERSE, spec, spec, fw, st));
        return launch_fft_rows(g, false, grid, spec, st);
    }
    if (real_out) {
        {
            ProfScope ps(ST_PACK, st);
            NF_TRY(NF_DIM_DISPATCH(launch_pack, g, true, xreal, xhat, spec, st));
        }
        return fft_exec(g, FFT_C2R, spec, grid, fw, st);
    }
    {
        ProfScope ps(ST_PACK, st);
        NF_TRY(NF_DIM_DISPATCH(launch_pack, g, false, xreal, xhat, reinterpret_cast<float2*>(grid), st));
    }
    return fft_exec(g, FFT_C2C_FORWARD, grid, grid, fw, st);
}

static int do_fastsum_middle(const Geom& g, float* grid, const float* coeffs, bool creal, float2* spec,
                             const FftWork& fw, cudaStream_t st) {
    if (!g.cplx) {
        NF_TRY(fft_exec(g, FFT_R2C, grid, spec, fw, st));
        {
            ProfScope ps(ST_MULTIPLY, st);
            NF_TRY(NF_DIM_DISPATCH(launch_multiply, g, true, creal, spec, coeffs, st));
        }
        return fft_exec(g, FFT_C2R, spec, grid, fw, st);
    }
    NF_TRY(fft_exec(g, FFT_C2C_INVERSE, grid, grid, fw, st));
    {
        ProfScope ps(ST_MULTIPLY, st);
        NF_TRY(NF_DIM_DISPATCH(launch_multiply, g, false, creal, reinterpret_cast<float2*>(grid), coeffs, st));
    }
    return fft_exec(g, FFT_C2C_FORWARD, grid, grid, fw, st);
}

// workspace carving ------------------------------------------------------------------------
// [sort scratch | point plan(s) | grid | half spectrum | cuFFT work area]; pieces an op does not need are empty
struct Workspace {
    size_t scratch, plan_a, plan_b, grid, spec, fft, fft_bytes, total;
};

static Workspace ws_layout(const Geom& g, long long n_sort_a, long long n_sort_b, bool need_grid, bool grid_cplx,
                           bool need_spec, bool need_fft) {
    Workspace w{};
    size_t off = 0;
    w.scratch = off;
    if (n_sort_a >= 0 || n_sort_b >= 0) {
        const size_t a = n_sort_a >= 0 ? sort_layout(n_sort_a, g).total : 0;
        const size_t b = n_sort_b >= 0 ? sort_layout(n_sort_b, g).total : 0;
        off += align_up(a > b ? a : b);
    }
    w.plan_a = off;
    if (n_sort_a >= 0) off += align_up(plan_layout(n_sort_a, g).total);
    w.plan_b = off;
    if (n_sort_b >= 0) off += align_up(plan_layout(n_sort_b, g).total);
    w.grid = off;
    if (need_grid) off += grid_bytes(g, grid_cplx);
    w.spec = off;
    if (need_spec) off += half_bytes(g);
    w.fft = off;
    w.fft_bytes = need_fft ? fft_work_bytes(g, grid_cplx) : 0;
    off += w.fft_bytes;
    w.total = off;
    return w;
}

}  // namespace nfftb200

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace nfftb200;

extern "C" {

int nfftb200_version(void) { return 200; }
const char* nfftb200_last_error(void) { return g_err; }
int64_t nfftb200_launch_count(void) { return (int64_t)g_launches.load(); }

// out[0..24] = dim,N,M,m,L, T[3], nt[3], P[3], sY,sZ, tile_elems, ncomp, pmax, spread_threads, use_reg, fine_bits, sc[3]
int nfftb200_debug_geometry(int d, int64_t N, int m, int64_t B, int64_t C, int flags, int64_t n, int32_t* out) {
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(flags & NFFTB200_X_COMPLEX, flags), n));
    int v[28] = {g.dim, g.N, g.M, g.m, g.L, g.T[0], g.T[1], g.T[2], g.nt[0], g.nt[1], g.nt[2], g.P[0], g.P[1], g.P[2],
                 g.sY, g.sZ, g.tile_elems, g.ncomp, g.pmax, g.spread_threads, g.use_reg, g.fine_bits, g.sc[0], g.sc[1],
                 g.sc[2], g.mixed, g.refine_pass, g.dense_tile_pts};
    for (int i = 0; i < 28; ++i) out[i] = v[i];
    return NFFTB200_OK;
}

void nfftb200_debug_force_int64(int on) { g_force_int64.store(on ? 1 : 0); }

void nfftb200_debug_pruned_fft(int mode) { g_pruned_mode.store(mode < 0 ? -1 : (mode ? 1 : 0)); }

int nfftb200_debug_min_resident_ctas(void) {
    const int v = g_min_resident.load();
    return v == (1 << 30) ? -1 : v;
}

void nfftb200_debug_mixed(int mode, int64_t min_points, int dense_tile_pts) {
    g_mixed_mode.store(mode < 0 ? -1 : (mode ? 1 : 0));
    g_mixed_min_points.store(min_points < 0 ? (1ll << 18) : (long long)min_points);
    g_mixed_dense_pts.store(dense_tile_pts > 0 ? dense_tile_pts : 0);
}

#ifdef NFFT_PHASE_TIMING
// debug build only: out[2][24] = accumulated clock64() phase lengths of the register-stencil kernels
// (window_reg.cuh); reset = 1 clears the counters afterwards
int nfftb200_debug_phase_read(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 48) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[48] = {0};
        cudaMemcpyToSymbol(g_phase, z, sizeof(z));
    }
    return 0;
}
#endif

void nfftb200_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_on = on != 0;
}

// Accumulates the elapsed milliseconds and call counts of all recorded stage spans into
// ms_out[8] / count_out[8] (stage order: sort, spread, fft, unpack, pack, gather, multiply, memset)
// and clears the record.  The caller must have synchronised the stream(s).
int nfftb200_profile_read(double* ms_out, int64_t* count_out) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (int i = 0; i < ST_COUNT; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
    for (const ProfSpan& sp : g_prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            ms_out[sp.stage] += ms;
            count_out[sp.stage] += 1;
        }
        g_prof_pool.push_back(sp.a);
        g_prof_pool.push_back(sp.b);
    }
    g_prof_spans.clear();
    (void)cudaGetLastError();
    return NFFTB200_OK;
}

int nfftb200_plan_cache_clear(void) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    if (g_plan_pins > 0)
        NF_FAIL(NFFTB200_ERR_INVALID, "cuFFT plan cache is pinned by %d live CUDA graph(s): not cleared", g_plan_pins);
    for (auto& kv : g_plans) cufftDestroy(kv.second.handle);
    g_plans.clear();
    return NFFTB200_OK;
}

int nfftb200_plan_cache_pin(int delta) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    g_plan_pins += delta;
    if (g_plan_pins < 0) g_plan_pins = 0;
    return g_plan_pins;
}

int nfftb200_plan_cache_size(void) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    return (int)g_plans.size();
}

// how each public op uses the grid: (grid is complex?, needs half-spectrum scratch?)
static void op_modes(int op, int flags, bool* grid_cplx, bool* need_spec) {
    const bool xc = flags & NFFTB200_X_COMPLEX, yr = flags & NFFTB200_Y_REAL;
    switch (op) {
        case NFFTB200_OP_ADJOINT: *grid_cplx = xc; *need_spec = !xc; break;
        case NFFTB200_OP_FORWARD: *grid_cplx = !yr; *need_spec = yr; break;
        case NFFTB200_OP_FASTSUM: *grid_cplx = xc; *need_spec = !xc; break;
        default: *grid_cplx = xc; *need_spec = false; break;
    }
}

// the layout of a public op (shared by nfftb200_workspace_bytes and the op itself)
static int op_layout(int op, long long n_src, long long n_tgt, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                     Geom* g, Workspace* w) {
    bool gc, ns;
    op_modes(op, flags, &gc, &ns);
    const long long np = n_src > n_tgt ? n_src : n_tgt;
    NF_TRY(make_geom(*g, d, N, m, B, C, geom_flags(gc, flags), np));
    const bool planned = flags & NFFTB200_PLANNED;  // the caller brings the point plan(s): no sort regions
    const bool sym = (flags & NFFTB200_SYMMETRIC) && n_src == n_tgt;
    switch (op) {
        case NFFTB200_OP_ADJOINT: *w = ws_layout(*g, planned ? -1 : n_src, -1, true, gc, ns, true); break;
        case NFFTB200_OP_FORWARD: *w = ws_layout(*g, planned ? -1 : n_tgt, -1, true, gc, ns, true); break;
        case NFFTB200_OP_FASTSUM:
            *w = ws_layout(*g, planned ? -1 : n_src, planned || sym ? -1 : n_tgt, true, gc, ns, true);
            break;
        case NFFTB200_OP_SPREAD:
        case NFFTB200_OP_GATHER: *w = ws_layout(*g, planned ? -1 : np, -1, false, gc, false, false); break;
        case NFFTB200_OP_SORT: *w = ws_layout(*g, np, -1, false, gc, false, false); break;
        case NFFTB200_OP_PLAN: {  // nfftb200_plan_points: scratch only, the plan lives in the caller's plan buffer
            *w = Workspace{};
            w->total = align_up(sort_layout(np, *g).total);
            break;
        }
        case NFFTB200_OP_SPECTRAL: {  // adjoint_finish / forward_begin / fastsum_middle: spectrum + FFT work area
            Geom gr = *g;
            *w = ws_layout(gr, -1, -1, false, flags & NFFTB200_X_COMPLEX, true, false);
            // the three stage helpers run R2C / C2R / C2C depending on their own flags: provide for all of them
            size_t f = fft_work_bytes(*g, false);
            Geom gcx = *g;
            gcx.cplx = 1;
            const size_t f2 = fft_work_bytes(gcx, true);
            w->fft = w->total;
            w->fft_bytes = f > f2 ? f : f2;
            w->total += w->fft_bytes;
            break;
        }
        default: NF_FAIL(NFFTB200_ERR_INVALID, "unknown op %d", op);
    }
    return NFFTB200_OK;
}

size_t nfftb200_workspace_bytes(int op, int64_t n_src, int64_t n_tgt, int d, int64_t N, int m, int64_t B, int64_t C,
                                int flags) {
    Geom g;
    Workspace w;
    if (op_layout(op, n_src, n_tgt, d, N, m, B, C, flags, &g, &w) != NFFTB200_OK) return 0;
    return w.total + 256;
}

size_t nfftb200_plan_bytes(int64_t n, int64_t n_geom, int d, int64_t N, int m, int64_t B, int64_t C, int flags) {
    Geom g;
    if (make_geom(g, d, N, m, B, C, geom_flags(flags & NFFTB200_X_COMPLEX, flags), n_geom > n ? n_geom : n) != NFFTB200_OK) return 0;
    return align_up(plan_layout(n, g).total) + 256;
}

#define NF_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) NF_FAIL(NFFTB200_ERR_INVALID, "%s", msg); \
    } while (0)

static char* align_ptr(const void* p) { return (char*)(((uintptr_t)p + 255) / 256 * 256); }

#define NF_CHECK_WS(w, bytes)                                                                                     \
    do {                                                                                                          \
        if ((bytes) < (w).total + 256)                                                                            \
            NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small: %zu < %zu", (size_t)(bytes), (w).total + 256);  \
    } while (0)

// resolves the point plan of an op: the caller's kept plan (NFFTB200_PLANNED) or a fresh sort into the workspace
static int resolve_plan(const Geom& g, int flags, const float* pos, const int64_t* batch, long long n, const void* plan,
                        size_t plan_bytes, char* ws, size_t scratch_off, size_t plan_off, SortPlan* sp, cudaStream_t st) {
    if (flags & NFFTB200_PLANNED) {
        NF_REQUIRE(plan != nullptr, "NFFTB200_PLANNED without a plan buffer");
        if (plan_bytes < plan_layout(n, g).total)
            NF_FAIL(NFFTB200_ERR_WORKSPACE, "plan buffer too small: %zu < %zu", plan_bytes, plan_layout(n, g).total);
        sort_plan_pointers(n, g, align_ptr(plan), sp);
        return NFFTB200_OK;
    }
    return do_sort(g, pos, batch, flags & NFFTB200_BATCH_OFFSETS, n, ws + scratch_off, ws + plan_off, sp, st);
}

int nfftb200_plan_points(const float* pos, const int64_t* batch, void* plan, size_t plan_bytes, int64_t n, int64_t n_geom,
                         int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                         size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && plan && workspace && (n == 0 || pos), "nfftb200_plan_points: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(flags & NFFTB200_X_COMPLEX, flags), n_geom > n ? n_geom : n));
    if (plan_bytes < align_up(plan_layout(n, g).total) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "plan buffer too small");
    if (workspace_bytes < align_up(sort_layout(n, g).total) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    SortPlan sp{};
    return do_sort(g, pos, batch, flags & NFFTB200_BATCH_OFFSETS, n, align_ptr(workspace), align_ptr(plan), &sp,
                   (cudaStream_t)stream);
}

// flags_out[8] (host): flags_out[0] = points the window kernels found outside their tile since the plan was made
// (non-zero: the positions changed after nfftb200_plan_points).  Synchronises the stream.
int nfftb200_plan_flags(const void* plan, int64_t n, int64_t n_geom, int d, int64_t N, int m, int64_t B, int64_t C,
                        int flags, uint32_t* flags_out, void* stream) {
    NF_REQUIRE(plan && flags_out, "nfftb200_plan_flags: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(flags & NFFTB200_X_COMPLEX, flags), n_geom > n ? n_geom : n));
    SortPlan sp{};
    sort_plan_pointers(n, g, align_ptr(plan), &sp);
    cudaStream_t st = (cudaStream_t)stream;
    NF_CUDA(cudaMemcpyAsync(flags_out, sp.flags, kPlanFlagWords * 4, cudaMemcpyDeviceToHost, st));
    NF_CUDA(cudaStreamSynchronize(st));
    return NFFTB200_OK;
}

int nfftb200_adjoint(const float* pos, const void* x, const int64_t* batch, void* y, int64_t n, int d, int64_t N, int m,
                     int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    return nfftb200_adjoint_planned(pos, x, batch, nullptr, 0, y, n, d, N, m, B, C, flags & ~NFFTB200_PLANNED, workspace,
                                    workspace_bytes, stream);
}

int nfftb200_adjoint_planned(const float* pos, const void* x, const int64_t* batch, const void* plan, size_t plan_bytes,
                             void* y, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                             void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && y && workspace && (n == 0 || (pos && x)), "nfftb200_adjoint: null pointer");
    const bool yr = flags & NFFTB200_Y_REAL;
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_ADJOINT, n, 0, d, N, m, B, C, flags, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    SortPlan sp{};
    NF_TRY(resolve_plan(g, flags, pos, batch, n, plan, plan_bytes, ws, w.scratch, w.plan_a, &sp, st));
    NF_TRY(do_spread(g, pos, (const float*)x, grid, n, sp, st));
    return do_adjoint_finish(g, grid, (float*)y, yr, (float2*)(ws + w.spec), FftWork{ws + w.fft, w.fft_bytes}, st);
}

int nfftb200_forward(const float* pos, const void* xhat, const int64_t* batch, void* y, int64_t n, int d, int64_t N,
                     int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    return nfftb200_forward_planned(pos, xhat, batch, nullptr, 0, y, n, d, N, m, B, C, flags & ~NFFTB200_PLANNED,
                                    workspace, workspace_bytes, stream);
}

int nfftb200_forward_planned(const float* pos, const void* xhat, const int64_t* batch, const void* plan,
                             size_t plan_bytes, void* y, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C,
                             int flags, void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && xhat && workspace && (n == 0 || (pos && y)), "nfftb200_forward: null pointer");
    const bool xc = flags & NFFTB200_X_COMPLEX, yr = flags & NFFTB200_Y_REAL;
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_FORWARD, 0, n, d, N, m, B, C, flags, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    if (n == 0) return NFFTB200_OK;
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    NF_TRY(do_forward_begin(g, (const float*)xhat, !xc, yr, grid, (float2*)(ws + w.spec),
                            FftWork{ws + w.fft, w.fft_bytes}, st));
    SortPlan sp{};
    NF_TRY(resolve_plan(g, flags, pos, batch, n, plan, plan_bytes, ws, w.scratch, w.plan_a, &sp, st));
    return do_gather(g, pos, grid, (float*)y, n, sp, st);
}

int nfftb200_fastsum(const float* sources, const float* targets, const void* x, const void* coeffs,
                     const int64_t* source_batch, const int64_t* target_batch, void* y, int64_t n_src, int64_t n_tgt,
                     int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes,
                     void* stream) {
    return nfftb200_fastsum_planned(sources, targets, x, coeffs, source_batch, target_batch, nullptr, 0, nullptr, 0, y,
                                    n_src, n_tgt, d, N, m, B, C, flags & ~NFFTB200_PLANNED, workspace, workspace_bytes,
                                    stream);
}

int nfftb200_fastsum_planned(const float* sources, const float* targets, const void* x, const void* coeffs,
                             const int64_t* source_batch, const int64_t* target_batch, const void* source_plan,
                             size_t source_plan_bytes, const void* target_plan, size_t target_plan_bytes, void* y,
                             int64_t n_src, int64_t n_tgt, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                             void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n_src >= 0 && n_tgt >= 0 && coeffs && workspace, "nfftb200_fastsum: null pointer");
    NF_REQUIRE(n_src == 0 || (sources && x), "nfftb200_fastsum: null sources/x");
    NF_REQUIRE(n_tgt == 0 || (targets && y), "nfftb200_fastsum: null targets/y");
    const bool creal = !(flags & NFFTB200_COEFFS_COMPLEX);
    const bool sym = (flags & NFFTB200_SYMMETRIC) && n_src == n_tgt;
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_FASTSUM, n_src, n_tgt, d, N, m, B, C, flags, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    if (n_tgt == 0) return NFFTB200_OK;
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    SortPlan sp_src{}, sp_tgt{};
    NF_TRY(resolve_plan(g, flags, sources, source_batch, n_src, source_plan, source_plan_bytes, ws, w.scratch, w.plan_a,
                        &sp_src, st));
    NF_TRY(do_spread(g, sources, (const float*)x, grid, n_src, sp_src, st));
    NF_TRY(do_fastsum_middle(g, grid, (const float*)coeffs, creal, (float2*)(ws + w.spec),
                             FftWork{ws + w.fft, w.fft_bytes}, st));
    if (sym) {
        sp_tgt = sp_src;
    } else {
        NF_TRY(resolve_plan(g, flags, targets, target_batch, n_tgt, target_plan, target_plan_bytes, ws, w.scratch,
                            w.plan_b, &sp_tgt, st));
    }
    return do_gather(g, targets, grid, (float*)y, n_tgt, sp_tgt, st);
}

int nfftb200_spread(const float* pos, const void* x, const int64_t* batch, const void* plan, size_t plan_bytes,
                    void* grid, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                    size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && grid && workspace && (n == 0 || (pos && x)), "nfftb200_spread: null pointer");
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_SPREAD, n, 0, d, N, m, B, C, flags, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    SortPlan sp{};
    NF_TRY(resolve_plan(g, flags, pos, batch, n, plan, plan_bytes, ws, w.scratch, w.plan_a, &sp, st));
    return do_spread(g, pos, (const float*)x, (float*)grid, n, sp, st);
}

int nfftb200_gather(const float* pos, const int64_t* batch, const void* plan, size_t plan_bytes, const void* grid,
                    void* y, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                    size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && grid && workspace && (n == 0 || (pos && y)), "nfftb200_gather: null pointer");
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_GATHER, 0, n, d, N, m, B, C, flags, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    SortPlan sp{};
    NF_TRY(resolve_plan(g, flags, pos, batch, n, plan, plan_bytes, ws, w.scratch, w.plan_a, &sp, st));
    return do_gather(g, pos, (const float*)grid, (float*)y, n, sp, st);
}

// the three spectral stage helpers share one layout: [half spectrum | cuFFT work area]
static int spectral_layout(const Geom& g, void* workspace, size_t workspace_bytes, float2** spec, FftWork* fw) {
    const size_t hb = half_bytes(g);
    if (workspace_bytes < hb + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    char* ws = align_ptr(workspace);
    *spec = (float2*)ws;
    fw->ptr = ws + hb;
    const size_t used = (size_t)(ws - (char*)workspace) + hb;
    fw->bytes = workspace_bytes > used ? workspace_bytes - used : 0;
    return NFFTB200_OK;
}

int nfftb200_adjoint_finish(void* grid, void* y, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                            void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(grid && y && workspace, "nfftb200_adjoint_finish: null pointer");
    const bool xc = flags & NFFTB200_X_COMPLEX;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(xc, 0), 0));
    float2* spec;
    FftWork fw;
    NF_TRY(spectral_layout(g, workspace, workspace_bytes, &spec, &fw));
    return do_adjoint_finish(g, (float*)grid, (float*)y, flags & NFFTB200_Y_REAL, spec, fw, (cudaStream_t)stream);
}

int nfftb200_forward_begin(const void* xhat, void* grid, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                           void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(xhat && grid && workspace, "nfftb200_forward_begin: null pointer");
    const bool yr = flags & NFFTB200_Y_REAL;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(!yr, 0), 0));
    float2* spec;
    FftWork fw;
    NF_TRY(spectral_layout(g, workspace, workspace_bytes, &spec, &fw));
    return do_forward_begin(g, (const float*)xhat, !(flags & NFFTB200_X_COMPLEX), yr, (float*)grid, spec, fw,
                            (cudaStream_t)stream);
}

int nfftb200_fastsum_middle(void* grid, const void* coeffs, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                            void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(grid && coeffs && workspace, "nfftb200_fastsum_middle: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, geom_flags(flags & NFFTB200_X_COMPLEX, 0), 0));
    float2* spec;
    FftWork fw;
    NF_TRY(spectral_layout(g, workspace, workspace_bytes, &spec, &fw));
    return do_fastsum_middle(g, (float*)grid, (const float*)coeffs, !(flags & NFFTB200_COEFFS_COMPLEX), spec, fw,
                             (cudaStream_t)stream);
}

int nfftb200_sort_points(const float* pos, const int64_t* batch, uint32_t* keys_out, uint32_t* perm_out,
                         int32_t* tile_out_host, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                         void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && workspace && (n == 0 || (pos && keys_out && perm_out)), "nfftb200_sort_points: null pointer");
    Geom g;
    Workspace w;
    NF_TRY(op_layout(NFFTB200_OP_SORT, n, 0, d, N, m, B, C, flags & ~NFFTB200_PLANNED, &g, &w));
    NF_CHECK_WS(w, workspace_bytes);
    if (tile_out_host) {
        for (int s = 0; s < 3; ++s) tile_out_host[s] = g.T[s];
    }
    SortPlan sp{};
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = align_ptr(workspace);
    NF_TRY(do_sort(g, pos, batch, flags & NFFTB200_BATCH_OFFSETS, n, ws + w.scratch, ws + w.plan_a, &sp, st));
    if (n > 0) {
        NF_CUDA(cudaMemcpyAsync(keys_out, sp.keys, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        NF_CUDA(cudaMemcpyAsync(perm_out, sp.perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return NFFTB200_OK;
}

}  // extern "C"
