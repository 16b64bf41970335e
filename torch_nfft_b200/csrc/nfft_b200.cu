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
#include "window.cuh"
#include "window_reg.cuh"
#include "window_reg2d.cuh"
#include "window1d.cuh"
#include <stdlib.h>

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

struct ProfScope {
    ProfSpan span{};
    cudaStream_t st;
    bool on;
    ProfScope(int stage, cudaStream_t s) : st(s), on(g_prof_on) {
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

static int make_geom(Geom& g, int d, int64_t N, int m, int64_t B, int64_t C, bool cplx, int64_t n_points) {
    if (d < 1 || d > 3) NF_FAIL(NFFTB200_ERR_INVALID, "dimension d=%d not in [1,3]", d);
    if (N < 2 || (N & 1)) NF_FAIL(NFFTB200_ERR_INVALID, "bandwidth N=%lld must be even and >= 2", (long long)N);
    if (m < 1 || m > kMaxCutoff) NF_FAIL(NFFTB200_ERR_INVALID, "cutoff m=%d not in [1,%d]", m, kMaxCutoff);
    if (B < 1 || C < 1) NF_FAIL(NFFTB200_ERR_INVALID, "batch size %lld / columns %lld must be >= 1", (long long)B, (long long)C);
    if (2 * N > (1 << 20)) NF_FAIL(NFFTB200_ERR_INVALID, "bandwidth N=%lld too large", (long long)N);
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

    const int team = d == 1 ? g.L : g.L * g.L;
    choose_strides(g);
    long long te = d == 1 ? g.P[0] : (d == 2 ? (long long)g.sY * g.P[1] : (long long)g.sZ * g.P[2]);
    g.tile_elems = (int)((te + 3) / 4 * 4);

    long long pm = n_points / (148 * 8);
    g.pmax = (int)(pm < 256 ? 256 : (pm > 2048 ? 2048 : pm));
    if (g.use_reg == 1) g.pmax = kRegMaxPts;
    if (g.use_reg == 2) g.pmax = kReg2MaxPts;
    if (g.use_reg == 3) {
        pm = n_points / (148 * 4);
        g.pmax = (int)(pm < 512 ? 512 : (pm > kW1MaxPts ? kW1MaxPts : pm));
    }
    int threads = (team + 31) / 32 * 32;
    g.spread_threads = threads < 64 ? 64 : threads;
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

static int launch_window(bool spread, const Geom& g, WindowArgs a, const SortPlan& sp, cudaStream_t st) {
    a.perm = sp.perm;
    a.bin_start = sp.bin_start;
    a.chunk_start = sp.chunk_start;
    a.items = sp.items;
    a.nbins = sp.nbins;
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
            if (smem > 227 * 1024) NF_FAIL(NFFTB200_ERR_INVALID, "tile needs %zu bytes of shared memory", smem);
            NF_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
            if (smem > 227 * 1024) NF_FAIL(NFFTB200_ERR_INVALID, "tile needs %zu bytes of shared memory", smem);
            NF_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NF_LAUNCH(kern, (unsigned)sp.max_items, spread ? kW1SpreadThreads : kW1Threads, smem, st, g, a);
        }
        return NFFTB200_OK;
    }
    if (g.use_reg == 1) {
        // supercell kRegSX x kRegSY x kRegSZ = 4 x 4 x 2 cells; for m = 4 the register block is 13 x 13 x 12:
        // 6 positions x 6 float2 accumulators per lane
        WindowKernel kern = nullptr;
        int win_floats = 0;
        switch (g.m) {
#define NF_REG_CASE(M_, L_)                                                                          \
            case M_:                                                                                 \
                kern = spread ? spread_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>                        \
                              : gather_reg_kernel<L_, kRegSX, kRegSY, kRegSZ>;                       \
                win_floats = RegCfg<L_, kRegSX, kRegSY, kRegSZ>::WIN_FLOATS;                         \
                break;
            NF_REG_CASE(1, 4)
            NF_REG_CASE(2, 6)
            NF_REG_CASE(3, 8)
            NF_REG_CASE(4, 10)
#undef NF_REG_CASE
            default: NF_FAIL(NFFTB200_ERR_INVALID, "register-stencil kernels need m <= 4");
        }
        const int nsc = ((g.T[0] + kRegSX - 1) / kRegSX) * ((g.T[1] + kRegSY - 1) / kRegSY) * ((g.T[2] + kRegSZ - 1) / kRegSZ);
        const size_t smem = reg_smem_bytes(g, nsc, win_floats);
        if (smem > 227 * 1024) NF_FAIL(NFFTB200_ERR_INVALID, "tile needs %zu bytes of shared memory", smem);
        NF_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int k0 = 0; k0 < g.K; ++k0) {
            a.k0 = k0;
            NF_LAUNCH(kern, (unsigned)sp.max_items, kRegThreads, smem, st, g, a);
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
        if (smem > 227 * 1024) NF_FAIL(NFFTB200_ERR_INVALID, "tile needs %zu bytes of shared memory", smem);
        NF_CUDA(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)sp.max_items;
        const unsigned block = spread ? (unsigned)g.spread_threads : (unsigned)kGatherThreads;
        NF_LAUNCH(kern, grid, block, smem, st, g, a);
    }
    return NFFTB200_OK;
}

// ----------------------------------------------------------------------------------------
// cuFFT plan cache (the reference creates and destroys a plan per call, core_cuda.cu:254-272)
// ----------------------------------------------------------------------------------------
struct PlanKey {
    int dev, dim, M, type;
    long long batch;
    bool operator<(const PlanKey& o) const {
        return std::tie(dev, dim, M, type, batch) < std::tie(o.dev, o.dim, o.M, o.type, o.batch);
    }
};
static std::mutex g_plan_mutex;
static std::map<PlanKey, cufftHandle> g_plans;

static int get_plan(int dim, int M, long long batch, cufftType type, cufftHandle* out) {
    int dev = 0;
    NF_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    PlanKey key{dev, dim, M, (int)type, batch};
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        *out = it->second;
        return NFFTB200_OK;
    }
    cufftHandle plan;
    NF_CUFFT(cufftCreate(&plan));
    long long n[3] = {M, M, M};
    long long real_dist = 1, half_dist = 1;
    for (int a = 0; a < dim; ++a) real_dist *= M;
    for (int a = 0; a < dim - 1; ++a) half_dist *= M;
    half_dist *= (M / 2 + 1);
    long long idist = real_dist, odist = real_dist;
    if (type == CUFFT_R2C) odist = half_dist;
    if (type == CUFFT_C2R) idist = half_dist;
    size_t work = 0;
    cufftResult r = cufftMakePlanMany64(plan, dim, n, nullptr, 1, idist, nullptr, 1, odist, type, batch, &work);
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(plan);
        NF_FAIL(NFFTB200_ERR_CUFFT, "cufftMakePlanMany64(dim=%d, M=%d, batch=%lld, type=%d) -> %d", dim, M, batch,
                (int)type, (int)r);
    }
    g_plans[key] = plan;
    *out = plan;
    return NFFTB200_OK;
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
static int do_spread(const Geom& g, const float* pos, const float* x, const int64_t* batch, float* grid, long long n,
                     char* sort_ws, SortPlan* sp_out, cudaStream_t st, bool presorted = false) {
    {
        ProfScope ps(ST_MEMSET, st);
        NF_CUDA(cudaMemsetAsync(grid, 0, (size_t)g.B * g.C * g.Md * (g.cplx ? 8 : 4), st));
    }
    SortPlan sp{};
    if (presorted) {
        sort_plan_pointers(n, g, sort_ws, &sp);
    } else {
        ProfScope ps(ST_SORT, st);
        NF_TRY(sort_points(pos, batch, n, g, sort_ws, &sp, st));
    }
    if (sp_out) *sp_out = sp;
    if (n == 0) return NFFTB200_OK;
    WindowArgs a{};
    a.pos = pos;
    a.xin = x;
    a.grid = grid;
    ProfScope ps(ST_SPREAD, st);
    return launch_window(true, g, a, sp, st);
}

static int do_gather(const Geom& g, const float* pos, const int64_t* batch, const float* grid, float* y, long long n,
                     char* sort_ws, const SortPlan* presorted, cudaStream_t st, bool presorted_ws = false) {
    if (n == 0) return NFFTB200_OK;
    SortPlan sp{};
    if (presorted) {
        sp = *presorted;
    } else if (presorted_ws) {
        sort_plan_pointers(n, g, sort_ws, &sp);
    } else {
        ProfScope ps(ST_SORT, st);
        NF_TRY(sort_points(pos, batch, n, g, sort_ws, &sp, st));
    }
    WindowArgs a{};
    a.pos = pos;
    a.yout = y;
    a.grid = const_cast<float*>(grid);
    ProfScope ps(ST_GATHER, st);
    return launch_window(false, g, a, sp, st);
}

// spectral kernels use 32-bit index arithmetic whenever every element index fits 31 bits
template <int DIM, typename I>
static int launch_unpack_t(const Geom& g, bool half, bool real_out, const float2* spec, float* y, cudaStream_t st) {
    long long total = (long long)g.B * g.C;
    for (int a = 0; a < DIM; ++a) total *= g.N;
    const unsigned grid = blocks_for(total);
    if (half) {
        if (real_out) NF_LAUNCH((unpack_kernel<DIM, true, true, I>), grid, 256, 0, st, spec, y, g);
        else NF_LAUNCH((unpack_kernel<DIM, true, false, I>), grid, 256, 0, st, spec, y, g);
    } else {
        if (real_out) NF_LAUNCH((unpack_kernel<DIM, false, true, I>), grid, 256, 0, st, spec, y, g);
        else NF_LAUNCH((unpack_kernel<DIM, false, false, I>), grid, 256, 0, st, spec, y, g);
    }
    return NFFTB200_OK;
}

template <int DIM, typename I>
static int launch_pack_t(const Geom& g, bool half, bool xreal, const float* xhat, float2* spec, cudaStream_t st) {
    // zero fill (out-of-band 7/8 of a 3D spectrum) with one memset, then write the band box
    const long long total = half ? half_elems(g) : (long long)g.B * g.C * g.Md;
    NF_CUDA(cudaMemsetAsync(spec, 0, (size_t)total * sizeof(float2), st));
    long long band = (long long)g.B * g.C * (half ? g.N / 2 + 1 : g.N);
    for (int a = 0; a < DIM - 1; ++a) band *= half ? g.N + 1 : g.N;
    const unsigned grid = blocks_for(band);
    if (half) {
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

static bool fits_int32(const Geom& g) {
    // largest index any spectral kernel forms: complex grid elements plus one block of slack
    return (long long)g.B * g.C * g.Md < (1ll << 31) - 1024;
}
template <int DIM>
static int launch_unpack(const Geom& g, bool half, bool real_out, const float2* spec, float* y, cudaStream_t st) {
    return fits_int32(g) ? launch_unpack_t<DIM, int>(g, half, real_out, spec, y, st)
                         : launch_unpack_t<DIM, long long>(g, half, real_out, spec, y, st);
}
template <int DIM>
static int launch_pack(const Geom& g, bool half, bool xreal, const float* xhat, float2* spec, cudaStream_t st) {
    return fits_int32(g) ? launch_pack_t<DIM, int>(g, half, xreal, xhat, spec, st)
                         : launch_pack_t<DIM, long long>(g, half, xreal, xhat, spec, st);
}
template <int DIM>
static int launch_multiply(const Geom& g, bool half, bool creal, float2* spec, const float* coeffs, cudaStream_t st) {
    return fits_int32(g) ? launch_multiply_t<DIM, int>(g, half, creal, spec, coeffs, st)
                         : launch_multiply_t<DIM, long long>(g, half, creal, spec, coeffs, st);
}

#define NF_DIM_DISPATCH(fn, ...)                                   \
    (g.dim == 1 ? fn<1>(__VA_ARGS__) : (g.dim == 2 ? fn<2>(__VA_ARGS__) : fn<3>(__VA_ARGS__)))

// grid (real: [BC][M^d] float, complex: [BC][M^d] float2) -> y.  spec: scratch for the half spectrum.
static int do_adjoint_finish(const Geom& g, float* grid, float* y, bool real_out, float2* spec, cudaStream_t st) {
    cufftHandle plan;
    if (!g.cplx) {
        NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_R2C, &plan));
        NF_CUFFT(cufftSetStream(plan, st));
        {
            ProfScope ps(ST_FFT, st);
            NF_CUFFT(cufftExecR2C(plan, grid, reinterpret_cast<cufftComplex*>(spec)));
        }
        ProfScope ps(ST_UNPACK, st);
        return NF_DIM_DISPATCH(launch_unpack, g, true, real_out, spec, y, st);
    }
    NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_C2C, &plan));
    NF_CUFFT(cufftSetStream(plan, st));
    {
        ProfScope ps(ST_FFT, st);
        NF_CUFFT(cufftExecC2C(plan, reinterpret_cast<cufftComplex*>(grid), reinterpret_cast<cufftComplex*>(grid),
                              CUFFT_INVERSE));  // sign +, core_cuda.cu:267
    }
    ProfScope ps(ST_UNPACK, st);
    return NF_DIM_DISPATCH(launch_unpack, g, false, real_out, reinterpret_cast<const float2*>(grid), y, st);
}

// xhat -> grid.  real_out: grid is float (C2R), else float2 (C2C sign -).
static int do_forward_begin(const Geom& g, const float* xhat, bool xreal, bool real_out, float* grid, float2* spec,
                            cudaStream_t st) {
    cufftHandle plan;
    if (real_out) {
        {
            ProfScope ps(ST_PACK, st);
            NF_TRY(NF_DIM_DISPATCH(launch_pack, g, true, xreal, xhat, spec, st));
        }
        NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_C2R, &plan));
        NF_CUFFT(cufftSetStream(plan, st));
        ProfScope ps(ST_FFT, st);
        NF_CUFFT(cufftExecC2R(plan, reinterpret_cast<cufftComplex*>(spec), grid));
        return NFFTB200_OK;
    }
    {
        ProfScope ps(ST_PACK, st);
        NF_TRY(NF_DIM_DISPATCH(launch_pack, g, false, xreal, xhat, reinterpret_cast<float2*>(grid), st));
    }
    NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_C2C, &plan));
    NF_CUFFT(cufftSetStream(plan, st));
    ProfScope ps(ST_FFT, st);
    NF_CUFFT(cufftExecC2C(plan, reinterpret_cast<cufftComplex*>(grid), reinterpret_cast<cufftComplex*>(grid),
                          CUFFT_FORWARD));  // sign -, core_cuda.cu:445
    return NFFTB200_OK;
}

static int do_fastsum_middle(const Geom& g, float* grid, const float* coeffs, bool creal, float2* spec,
                             cudaStream_t st) {
    cufftHandle plan;
    if (!g.cplx) {
        NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_R2C, &plan));
        NF_CUFFT(cufftSetStream(plan, st));
        {
            ProfScope ps(ST_FFT, st);
            NF_CUFFT(cufftExecR2C(plan, grid, reinterpret_cast<cufftComplex*>(spec)));
        }
        {
            ProfScope ps(ST_MULTIPLY, st);
            NF_TRY(NF_DIM_DISPATCH(launch_multiply, g, true, creal, spec, coeffs, st));
        }
        NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_C2R, &plan));
        NF_CUFFT(cufftSetStream(plan, st));
        ProfScope ps(ST_FFT, st);
        NF_CUFFT(cufftExecC2R(plan, reinterpret_cast<cufftComplex*>(spec), grid));
        return NFFTB200_OK;
    }
    NF_TRY(get_plan(g.dim, g.M, (long long)g.B * g.C, CUFFT_C2C, &plan));
    NF_CUFFT(cufftSetStream(plan, st));
    cufftComplex* gc = reinterpret_cast<cufftComplex*>(grid);
    {
        ProfScope ps(ST_FFT, st);
        NF_CUFFT(cufftExecC2C(plan, gc, gc, CUFFT_INVERSE));
    }
    {
        ProfScope ps(ST_MULTIPLY, st);
        NF_TRY(NF_DIM_DISPATCH(launch_multiply, g, false, creal, reinterpret_cast<float2*>(grid), coeffs, st));
    }
    ProfScope ps(ST_FFT, st);
    NF_CUFFT(cufftExecC2C(plan, gc, gc, CUFFT_FORWARD));
    return NFFTB200_OK;
}

// workspace carving ------------------------------------------------------------------------
struct Workspace {
    size_t sort, grid, spec, total;
};

// which pieces an op needs: sort scratch for max(n_src, n_tgt) points, a grid, a half spectrum
static Workspace ws_layout(int op, const Geom& g, long long n_src, long long n_tgt, bool grid_cplx, bool need_spec) {
    Workspace w{};
    size_t off = 0;
    w.sort = off;
    const bool need_sort = op != -1;
    if (need_sort) {
        Geom gs = g;
        size_t a = sort_layout(n_src, gs).total, b = sort_layout(n_tgt, gs).total;
        off += align_up(a > b ? a : b);
    }
    w.grid = off;
    if (op == NFFTB200_OP_ADJOINT || op == NFFTB200_OP_FORWARD || op == NFFTB200_OP_FASTSUM) off += grid_bytes(g, grid_cplx);
    w.spec = off;
    if (need_spec) off += half_bytes(g);
    w.total = off;
    return w;
}

}  // namespace nfftb200

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace nfftb200;

extern "C" {

int nfftb200_version(void) { return 100; }
const char* nfftb200_last_error(void) { return g_err; }
int64_t nfftb200_launch_count(void) { return (int64_t)g_launches.load(); }

// out[0..20] = dim,N,M,m,L, T[3], nt[3], P[3], sY,sZ, tile_elems, ncomp, pmax, spread_threads, use_reg
int nfftb200_debug_geometry(int d, int64_t N, int m, int64_t B, int64_t C, int flags, int64_t n, int32_t* out) {
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, flags & NFFTB200_X_COMPLEX, n));
    int v[21] = {g.dim, g.N, g.M, g.m, g.L, g.T[0], g.T[1], g.T[2], g.nt[0], g.nt[1], g.nt[2], g.P[0], g.P[1], g.P[2],
                 g.sY, g.sZ, g.tile_elems, g.ncomp, g.pmax, g.spread_threads, g.use_reg};
    for (int i = 0; i < 21; ++i) out[i] = v[i];
    return NFFTB200_OK;
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
    for (auto& kv : g_plans) cufftDestroy(kv.second);
    g_plans.clear();
    return NFFTB200_OK;
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

size_t nfftb200_workspace_bytes(int op, int64_t n_src, int64_t n_tgt, int d, int64_t N, int m, int64_t B, int64_t C,
                                int flags) {
    bool gc, ns;
    op_modes(op, flags, &gc, &ns);
    Geom g;
    const long long np = n_src > n_tgt ? n_src : n_tgt;
    if (make_geom(g, d, N, m, B, C, gc, np) != NFFTB200_OK) return 0;
    if (op == NFFTB200_OP_SPREAD || op == NFFTB200_OP_GATHER || op == NFFTB200_OP_SORT) {
        return align_up(sort_layout(np, g).total) + 256;
    }
    if (op == NFFTB200_OP_SPECTRAL) return half_bytes(g) + 256;
    // stage-only helpers: adjoint_finish / forward_begin / fastsum_middle need only the spectrum
    return ws_layout(op, g, n_src, n_tgt, gc, true).total + 256;
}

#define NF_REQUIRE(cond, msg)                                  \
    do {                                                       \
        if (!(cond)) NF_FAIL(NFFTB200_ERR_INVALID, "%s", msg); \
    } while (0)

static char* align_ptr(void* p) { return (char*)(((uintptr_t)p + 255) / 256 * 256); }

int nfftb200_adjoint(const float* pos, const void* x, const int64_t* batch, void* y, int64_t n, int d, int64_t N, int m,
                     int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && y && workspace && (n == 0 || (pos && x)), "nfftb200_adjoint: null pointer");
    const bool xc = flags & NFFTB200_X_COMPLEX, yr = flags & NFFTB200_Y_REAL;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, xc, n));
    const Workspace w = ws_layout(NFFTB200_OP_ADJOINT, g, n, 0, xc, true);
    if (workspace_bytes < w.total + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total + 256);
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    NF_TRY(do_spread(g, pos, (const float*)x, batch, grid, n, ws + w.sort, nullptr, st, flags & NFFTB200_PRESORTED));
    return do_adjoint_finish(g, grid, (float*)y, yr, (float2*)(ws + w.spec), st);
}

int nfftb200_forward(const float* pos, const void* xhat, const int64_t* batch, void* y, int64_t n, int d, int64_t N,
                     int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && xhat && workspace && (n == 0 || (pos && y)), "nfftb200_forward: null pointer");
    const bool xc = flags & NFFTB200_X_COMPLEX, yr = flags & NFFTB200_Y_REAL;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, !yr, n));
    const Workspace w = ws_layout(NFFTB200_OP_FORWARD, g, 0, n, !yr, true);
    if (workspace_bytes < w.total + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total + 256);
    if (n == 0) return NFFTB200_OK;
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    NF_TRY(do_forward_begin(g, (const float*)xhat, !xc, yr, grid, (float2*)(ws + w.spec), st));
    return do_gather(g, pos, batch, grid, (float*)y, n, ws + w.sort, nullptr, st, flags & NFFTB200_PRESORTED);
}

int nfftb200_fastsum(const float* sources, const float* targets, const void* x, const void* coeffs,
                     const int64_t* source_batch, const int64_t* target_batch, void* y, int64_t n_src, int64_t n_tgt,
                     int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes,
                     void* stream) {
    NF_REQUIRE(n_src >= 0 && n_tgt >= 0 && coeffs && workspace, "nfftb200_fastsum: null pointer");
    NF_REQUIRE(n_src == 0 || (sources && x), "nfftb200_fastsum: null sources/x");
    NF_REQUIRE(n_tgt == 0 || (targets && y), "nfftb200_fastsum: null targets/y");
    const bool xc = flags & NFFTB200_X_COMPLEX;
    const bool creal = !(flags & NFFTB200_COEFFS_COMPLEX);
    const bool sym = (flags & NFFTB200_SYMMETRIC) && n_src == n_tgt;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, xc, n_src > n_tgt ? n_src : n_tgt));
    const Workspace w = ws_layout(NFFTB200_OP_FASTSUM, g, n_src, n_tgt, xc, true);
    if (workspace_bytes < w.total + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total + 256);
    if (n_tgt == 0) return NFFTB200_OK;
    char* ws = align_ptr(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    float* grid = (float*)(ws + w.grid);
    SortPlan sp{};
    NF_TRY(do_spread(g, sources, (const float*)x, source_batch, grid, n_src, ws + w.sort, &sp, st));
    NF_TRY(do_fastsum_middle(g, grid, (const float*)coeffs, creal, (float2*)(ws + w.spec), st));
    return do_gather(g, targets, target_batch, grid, (float*)y, n_tgt, ws + w.sort, sym ? &sp : nullptr, st);
}

int nfftb200_spread(const float* pos, const void* x, const int64_t* batch, void* grid, int64_t n, int d, int64_t N,
                    int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && grid && workspace && (n == 0 || (pos && x)), "nfftb200_spread: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, flags & NFFTB200_X_COMPLEX, n));
    if (workspace_bytes < align_up(sort_layout(n, g).total) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    return do_spread(g, pos, (const float*)x, batch, (float*)grid, n, align_ptr(workspace), nullptr, (cudaStream_t)stream);
}

int nfftb200_gather(const float* pos, const int64_t* batch, const void* grid, void* y, int64_t n, int d, int64_t N,
                    int m, int64_t B, int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && grid && workspace && (n == 0 || (pos && y)), "nfftb200_gather: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, flags & NFFTB200_X_COMPLEX, n));
    if (workspace_bytes < align_up(sort_layout(n, g).total) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    return do_gather(g, pos, batch, (const float*)grid, (float*)y, n, align_ptr(workspace), nullptr, (cudaStream_t)stream);
}

int nfftb200_adjoint_finish(void* grid, void* y, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                            void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(grid && y && workspace, "nfftb200_adjoint_finish: null pointer");
    const bool xc = flags & NFFTB200_X_COMPLEX;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, xc, 0));
    if (workspace_bytes < half_bytes(g) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    return do_adjoint_finish(g, (float*)grid, (float*)y, flags & NFFTB200_Y_REAL, (float2*)align_ptr(workspace),
                             (cudaStream_t)stream);
}

int nfftb200_forward_begin(const void* xhat, void* grid, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                           void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(xhat && grid && workspace, "nfftb200_forward_begin: null pointer");
    const bool yr = flags & NFFTB200_Y_REAL;
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, !yr, 0));
    if (workspace_bytes < half_bytes(g) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    return do_forward_begin(g, (const float*)xhat, !(flags & NFFTB200_X_COMPLEX), yr, (float*)grid,
                            (float2*)align_ptr(workspace), (cudaStream_t)stream);
}

int nfftb200_fastsum_middle(void* grid, const void* coeffs, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                            void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(grid && coeffs && workspace, "nfftb200_fastsum_middle: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, flags & NFFTB200_X_COMPLEX, 0));
    if (workspace_bytes < half_bytes(g) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    return do_fastsum_middle(g, (float*)grid, (const float*)coeffs, !(flags & NFFTB200_COEFFS_COMPLEX),
                             (float2*)align_ptr(workspace), (cudaStream_t)stream);
}

int nfftb200_sort_points(const float* pos, const int64_t* batch, uint32_t* keys_out, uint32_t* perm_out,
                         int32_t* tile_out_host, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C, int flags,
                         void* workspace, size_t workspace_bytes, void* stream) {
    NF_REQUIRE(n >= 0 && workspace && (n == 0 || (pos && keys_out && perm_out)), "nfftb200_sort_points: null pointer");
    Geom g;
    NF_TRY(make_geom(g, d, N, m, B, C, flags & NFFTB200_X_COMPLEX, n));
    if (workspace_bytes < align_up(sort_layout(n, g).total) + 256) NF_FAIL(NFFTB200_ERR_WORKSPACE, "workspace too small");
    if (tile_out_host) {
        for (int s = 0; s < 3; ++s) tile_out_host[s] = g.T[s];
    }
    SortPlan sp{};
    cudaStream_t st = (cudaStream_t)stream;
    NF_TRY(sort_points(pos, batch, n, g, align_ptr(workspace), &sp, st));
    if (n > 0) {
        NF_CUDA(cudaMemcpyAsync(keys_out, sp.keys, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        NF_CUDA(cudaMemcpyAsync(perm_out, sp.perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return NFFTB200_OK;
}

}  // extern "C"
