// Shared definitions for the sm_100a NFFT engine.  Torch-free.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/nfft_b200.h"

namespace nfftb200 {

// ----------------------------------------------------------------------------------------
// error plumbing: status codes + thread-local message, never exit()/throw across the ABI
// (the reference exits the process on CUDA errors, csrc/cuda/cuda_utils.cu:7-14).
// ----------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

#define NF_FAIL(code, ...)                      \
    do {                                        \
        snprintf(g_err, sizeof(g_err), __VA_ARGS__); \
        return (code);                          \
    } while (0)

#define NF_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            NF_FAIL(NFFTB200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                    cudaGetErrorString(_e));                                            \
    } while (0)

#define NF_CUFFT(expr)                                                                  \
    do {                                                                                \
        cufftResult _r = (expr);                                                        \
        if (_r != CUFFT_SUCCESS)                                                        \
            NF_FAIL(NFFTB200_ERR_CUFFT, "%s:%d %s -> cufft error %d", __FILE__, __LINE__, \
                    #expr, (int)_r);                                                    \
    } while (0)

#define NF_TRY(expr)              \
    do {                          \
        int _s = (expr);          \
        if (_s != NFFTB200_OK) return _s; \
    } while (0)

// kernel launch + launch-error check + launch accounting
#define NF_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);               \
        g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        NF_CUDA(cudaGetLastError());                                              \
    } while (0)

// Window constants, written exactly as the reference evaluates them on the host
// (spatial_window_operations.cu:1-4, spectral_window_operations.cu:1-2).
constexpr float kThreeQuarterPi = 2.356194490192344928846982537459627163147877049531f;
constexpr float kPiThird = 1.047197551196597746154214461093167628065723133125f;

constexpr int kMaxCutoff = 8;             // m <= 8  (L = 2m+2 <= 18)
constexpr int kMaxL = 2 * kMaxCutoff + 2; // 18
constexpr int kSubBatch = 64;             // points staged in shared memory per round

// ----------------------------------------------------------------------------------------
// Geometry of one transform: oversampled grid, tiling, shared-memory tile layout.
// "Slots" are internal coordinates: slot 0 = X = fastest grid axis (API dim d-1),
// slot 1 = Y (API dim d-2), slot 2 = Z (API dim d-3).
// ----------------------------------------------------------------------------------------
struct Geom {
    int dim;
    int N, M, m, L, LP;  // LP = L rounded up to a multiple of 4 (psi row pitch in smem)
    int T[3];            // tile core extent per slot (cells that own points)
    int nt[3];           // tiles per slot
    int P[3];            // padded tile extent per slot (P[0] is a multiple of 4)
    int org[3];          // origin of padded tile = t*T - org  (org[0] multiple of 4)
    int sY, sZ;          // shared-memory strides of slot 1 / slot 2 (floats)
    int tile_elems;      // floats per component in the shared-memory tile (multiple of 4)
    int tiles_per_batch;
    long long Md;        // M^dim
    int B, C, K, cplx;   // K = C * (1 + cplx) float components per point
    int ncomp;           // components handled per kernel pass (1,2,4,8)
    int pmax;            // max points per work item
    int spread_threads;  // block size of the spread kernel
    int use_reg;         // 1: register-stencil kernels (window_reg.cuh), 0: team kernels (window.cuh)
    int fine_bits;       // low bits of a sort key: position of the point's supercell inside its tile (sort.cuh)
    int sc[3];           // supercell extent per slot (3D register-stencil kernels: 4 x 4 x 2, dense point sets 2 x 2 x 2)
    int fine_xy_levels, fine_z_bits;  // log2 of the supercells per tile edge in X / Y, and in Z
    int mixed;           // 1: density decided per TILE on the device (3D register-stencil kernels, sort.cuh: refine pass +
                         //    work-item classes); sc[] then describes the 2 x 2 x 2 hierarchy of the fine key bits
    int refine_pass;     // mixed: the lowest radix pass covers fine key bits only and is conditional (sort_points)
    int dense_tile_pts;  // mixed: a tile with at least this many points is swept with 2 x 2 x 2 supercells
    float inv_b, inv_sqrt_b_pi, c_hat;
    float kexp[kMaxCutoff + 2];  // exp(-j^2 inv_b), j = 0 .. m + 1 (tap recurrence of the register-stencil kernels)
};

struct SortPlan {
    uint32_t* keys;         // [n]  tile key per point (input order); sort scratch, null for a kept plan
    uint32_t* perm;         // [n]  stable permutation (sorted position -> input index)
    uint32_t* bin_start;    // [nbins+1]
    uint32_t* chunk_start;  // [nbins+1]; chunk_start[nbins] = number of work items
    uint4* items;           // [max_items] {bin, first point, one past last point, 0}; zero beyond the last item
    uint32_t* flags;        // [kPlanFlagWords] flags[0] = points a window kernel found outside their tile (stale plan)
    long long nbins;
    long long max_items;
};

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

}  // namespace nfftb200
