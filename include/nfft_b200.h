/*
 * nfft_b200.h -- C ABI of the B200-native NFFT engine (libnfft_b200.so).
 *
 * This is the drop-in boundary for the hot path of dominikbuenger/torch_nfft: the three
 * operators the reference registers as torch.ops.torch_nfft.{nfft_adjoint,nfft_forward,
 * nfft_fastsum} (reference csrc/core.cpp:43-55, 94-105, 108-121, 176-179), i.e. the host
 * functions nfft_adjoint_cuda / nfft_forward_cuda / nfft_fastsum_cuda
 * (reference csrc/cuda/core_cuda.cu:144-336, 340-531, 535-852).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All data pointers are DEVICE pointers
 *     on the current CUDA device; the caller owns every buffer including the workspace.
 *   - every entry point returns 0 on success, a negative NFFTB200_ERR_* code otherwise and
 *     never throws or exits (the reference calls exit() on CUDA errors, cuda_utils.cu:7-14).
 *     nfftb200_last_error() returns a thread-local message for the last failure.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), including cuFFT;
 *     no host synchronisation happens inside (the reference synchronises the device after
 *     every kernel, cuda_utils.cu:16).
 *   - layouts are the reference's: pos [n,d] float32 row-major in [-1/2,1/2) (wrapped
 *     periodically otherwise); batch [n] int64 sorted ascending or NULL (= one point set);
 *     spatial values x / y [n, C] float32 or complex64 (interleaved), channels last;
 *     spectral values [B, N,...,N, C] channels last with frequency k stored at k + N/2
 *     (core_cuda.cu:298-308); coeffs [N]^d float32 or complex64 with b_l at l + N/2
 *     (spectral_window_operations.cu:303-318).
 *   - window: Gaussian, oversampling 2 (M = 2N), 2m+2 taps per dimension -- the reference's
 *     (spatial_window_operations.cu:3-28, spectral_window_operations.cu:2-18).
 *   - N must be even, 1 <= d <= 3, 1 <= m <= 8, C >= 1, B >= 1.
 */
#ifndef NFFT_B200_H
#define NFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFFTB200_OK 0
#define NFFTB200_ERR_INVALID (-1)   /* bad argument (shape, range, null pointer)            */
#define NFFTB200_ERR_WORKSPACE (-2) /* workspace too small                                  */
#define NFFTB200_ERR_CUDA (-3)      /* CUDA runtime error (message in nfftb200_last_error)  */
#define NFFTB200_ERR_CUFFT (-4)     /* cuFFT error                                          */

/* operation ids for nfftb200_workspace_bytes */
#define NFFTB200_OP_ADJOINT 0
#define NFFTB200_OP_FORWARD 1
#define NFFTB200_OP_FASTSUM 2
#define NFFTB200_OP_SPREAD 3 /* split entry points below */
#define NFFTB200_OP_GATHER 4
#define NFFTB200_OP_SORT 5
#define NFFTB200_OP_SPECTRAL 6 /* adjoint_finish / forward_begin / fastsum_middle */
#define NFFTB200_OP_PLAN 7     /* nfftb200_plan_points (sort scratch only)            */

/* flags */
#define NFFTB200_X_COMPLEX 1      /* input values are complex64                              */
#define NFFTB200_Y_REAL 2         /* adjoint/forward: write only the real part (real_output) */
#define NFFTB200_COEFFS_COMPLEX 4 /* fastsum: coeffs are complex64                           */
#define NFFTB200_SYMMETRIC 8      /* fastsum: targets are the sources (reuse the sort)       */
#define NFFTB200_PLANNED 16       /* *_planned / spread / gather: the caller passes the point */
                                  /* plan(s) made by nfftb200_plan_points for exactly these   */
                                  /* points and this geometry; the workspace then holds no    */
                                  /* sort regions (also a flag of nfftb200_workspace_bytes)   */
#define NFFTB200_BATCH_OFFSETS 32 /* the `batch` pointers are B+1 ascending int64 OFFSETS of  */
                                  /* the point sets (CSR style) instead of one entry per point */
#define NFFTB200_CLUSTERED 64     /* hint: the points are clustered (large 3D sets, m = 3, 4).  */
                                  /* The binning then samples the keys on the device and, if it  */
                                  /* finds heavy tiles, refines the sort and marks their work    */
                                  /* items for the 2 x 2 x 2-supercell sweep.  Part of the        */
                                  /* geometry: a plan and the calls that use it must agree on it */

int nfftb200_version(void);
const char* nfftb200_last_error(void);

/* Bytes of device workspace the given call needs (0 on invalid arguments).
 * n_src / n_tgt: number of points spread / gathered (adjoint: n_src, forward: n_tgt).
 * The workspace is [sort scratch | point plan(s) | grid | half spectrum | cuFFT work area]: cuFFT plans are
 * cached per (device, shape) WITHOUT a work area of their own; it is taken from the caller's workspace on
 * every call, so concurrent streams and captured CUDA graphs never share FFT scratch.  (Without a usable
 * device the work-area term is 0; the figure is then a lower bound.) */
size_t nfftb200_workspace_bytes(int op, int64_t n_src, int64_t n_tgt, int d, int64_t N, int m,
                                int64_t B, int64_t C, int flags);

/* Replaces torch_nfft::nfft_adjoint (core.cpp:43-55 -> core_cuda.cu:144-336).
 * x [n,C] (float32, or complex64 with NFFTB200_X_COMPLEX) -> y [B, N^d, C]
 * (complex64, or float32 real part with NFFTB200_Y_REAL).  y is fully overwritten. */
int nfftb200_adjoint(const float* pos, const void* x, const int64_t* batch, void* y, int64_t n,
                     int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Replaces torch_nfft::nfft_forward (core.cpp:94-105 -> core_cuda.cu:340-531).
 * xhat [B, N^d, C] (float32 or complex64) -> y [n, C] (complex64 or float32). */
int nfftb200_forward(const float* pos, const void* xhat, const int64_t* batch, void* y, int64_t n,
                     int d, int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Replaces torch_nfft::nfft_fastsum (core.cpp:108-121 -> core_cuda.cu:535-852).
 * x [n_src, C] -> y [n_tgt, C], same dtype as x (real x: real part, core_cuda.cu:814-818). */
int nfftb200_fastsum(const float* sources, const float* targets, const void* x, const void* coeffs,
                     const int64_t* source_batch, const int64_t* target_batch, void* y,
                     int64_t n_src, int64_t n_tgt, int d, int64_t N, int m, int64_t B, int64_t C,
                     int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ---- point plans -------------------------------------------------------------------------
 * The binning of a point set (stable permutation by grid tile, bin offsets, work items) can be made once
 * and reused by every transform of the same points with the same tiling -- adjoint -> forward, forward ->
 * backward, iterative solvers -- where the reference recomputes its per-point scratch in every call
 * (core_cuda.cu:188-211, 461-484).  The plan lives in a caller-owned device buffer of nfftb200_plan_bytes
 * (about 4 bytes per point).  n_geom: the point count the TILING is chosen for (0 or n for adjoint / forward;
 * max(n_src, n_tgt) for both plans of a fastsum).  A plan is valid for one (d, N, m, B, C, real/complex
 * values) geometry and the positions it was made from; transforms count the points they find outside their
 * tile in the plan's flag word (nfftb200_plan_flags) instead of writing out of bounds. */
size_t nfftb200_plan_bytes(int64_t n, int64_t n_geom, int d, int64_t N, int m, int64_t B, int64_t C,
                           int flags);
int nfftb200_plan_points(const float* pos, const int64_t* batch, void* plan, size_t plan_bytes,
                         int64_t n, int64_t n_geom, int d, int64_t N, int m, int64_t B, int64_t C,
                         int flags, void* workspace, size_t workspace_bytes, void* stream);
/* flags_out[8] on the HOST; flags_out[0] = points found outside their tile so far, [1] = TMA tile loads that did
 * not complete (must be 0), [2] = 1 if the binning (with NFFTB200_CLUSTERED) found the point set clustered and marked
 * its heavy tiles for the 2 x 2 x 2 sweep.  Synchronises. */
int nfftb200_plan_flags(const void* plan, int64_t n, int64_t n_geom, int d, int64_t N, int m, int64_t B,
                        int64_t C, int flags, uint32_t* flags_out, void* stream);

/* The three operators with caller-kept plans (flags must contain NFFTB200_PLANNED; `batch` is unused then
 * and may be NULL).  Without NFFTB200_PLANNED they bin the points into the workspace like the plain calls. */
int nfftb200_adjoint_planned(const float* pos, const void* x, const int64_t* batch, const void* plan,
                             size_t plan_bytes, void* y, int64_t n, int d, int64_t N, int m, int64_t B,
                             int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream);
int nfftb200_forward_planned(const float* pos, const void* xhat, const int64_t* batch, const void* plan,
                             size_t plan_bytes, void* y, int64_t n, int d, int64_t N, int m, int64_t B,
                             int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream);
int nfftb200_fastsum_planned(const float* sources, const float* targets, const void* x, const void* coeffs,
                             const int64_t* source_batch, const int64_t* target_batch,
                             const void* source_plan, size_t source_plan_bytes, const void* target_plan,
                             size_t target_plan_bytes, void* y, int64_t n_src, int64_t n_tgt, int d,
                             int64_t N, int m, int64_t B, int64_t C, int flags, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---- split entry points (multi-GPU point sharding, tests) --------------------------------
 * grid: oversampled grid, planar per (b,c): float32 [B*C, M^d] for real values, complex64
 * [B*C, M^d] with NFFTB200_X_COMPLEX.  M = 2N.  */

/* adjoint stage 1 (compute_shifts/compute_psi/adjoint_window_convolution kernels,
 * spatial_window_operations.cu:38-211): grid = sum of window contributions (grid is zeroed).
 * plan / plan_bytes: a kept point plan with NFFTB200_PLANNED, else NULL / 0. */
int nfftb200_spread(const float* pos, const void* x, const int64_t* batch, const void* plan,
                    size_t plan_bytes, void* grid, int64_t n, int d, int64_t N, int m, int64_t B,
                    int64_t C, int flags, void* workspace, size_t workspace_bytes, void* stream);

/* adjoint stage 2 (cuFFT + adjoint_rolloff_correction, core_cuda.cu:254-326): grid -> y.
 * The grid is consumed (may be overwritten). */
int nfftb200_adjoint_finish(void* grid, void* y, int d, int64_t N, int m, int64_t B, int64_t C,
                            int flags, void* workspace, size_t workspace_bytes, void* stream);

/* forward stage 1 (forward_rolloff_correction + cuFFT, core_cuda.cu:397-450): xhat -> grid.
 * With NFFTB200_Y_REAL the grid is float32 (C2R path), otherwise complex64. */
int nfftb200_forward_begin(const void* xhat, void* grid, int d, int64_t N, int m, int64_t B,
                           int64_t C, int flags, void* workspace, size_t workspace_bytes,
                           void* stream);

/* forward stage 2 (forward_window_convolution, spatial_window_operations.cu:214-332):
 * y[i,c] = sum of window-weighted grid values.  NFFTB200_X_COMPLEX: grid and y complex64. */
int nfftb200_gather(const float* pos, const int64_t* batch, const void* plan, size_t plan_bytes,
                    const void* grid, void* y, int64_t n, int d, int64_t N, int m, int64_t B, int64_t C,
                    int flags, void* workspace, size_t workspace_bytes, void* stream);

/* fastsum middle stage (FFT, kernel_convolution, FFT; core_cuda.cu:683-765): grid -> grid. */
int nfftb200_fastsum_middle(void* grid, const void* coeffs, int d, int64_t N, int m, int64_t B,
                            int64_t C, int flags, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Deterministic binning only: keys_out[n] (uint32 sort key per point, input order: tile key <<
 * fine_bits | position of the point's supercell inside the tile, see debug_geometry), perm_out[n]
 * (uint32 stable sort permutation by that key), tile_out[3] (tile extents X,Y,Z used). */
int nfftb200_sort_points(const float* pos, const int64_t* batch, uint32_t* keys_out,
                         uint32_t* perm_out, int32_t* tile_out_host, int64_t n, int d, int64_t N,
                         int m, int64_t B, int64_t C, int flags, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Host-only: the tiling the engine would use.  out[28] = dim,N,M,m,L, T[3], nt[3], P[3], sY, sZ,
 * tile_elems, ncomp, pmax, spread_threads, use_reg, fine_bits, supercell[3] (slot order X,Y,Z; see
 * DESIGN.md), mixed (density decided per tile on the device), refine_pass, dense_tile_pts. */
int nfftb200_debug_geometry(int d, int64_t N, int m, int64_t B, int64_t C, int flags, int64_t n,
                            int32_t* out);

/* Optional per-stage timing with CUDA events recorded on the caller's stream around each stage.
 * read: ms_out[8], count_out[8] in stage order sort, spread, fft, unpack, pack, gather, multiply,
 * memset; call after synchronising the stream.  Used by bench.py for the roofline numbers. */
void nfftb200_profile_enable(int on);
int nfftb200_profile_read(double* ms_out, int64_t* count_out);

/* Tests: force the 64-bit index variants of the spectral kernels (normally chosen when B*C*M^d >= 2^31). */
void nfftb200_debug_force_int64(int on);

/* Tests / experiments: the pruned real transforms (hand-written X pass + cuFFT C2C over the kept kx planes; grids
 * with 256 or 512 cells per dimension, d = 2, 3).  mode -1 = default (on), 0 = plain cuFFT R2C / C2R, 1 = on. */
void nfftb200_debug_pruned_fft(int mode);

/* Tests: the smallest number of resident CTAs per SM the runtime reported for any launch configuration of the 3D
 * register-stencil sweeps so far (they are built for 2; -1 = none launched yet). */
int nfftb200_debug_min_resident_ctas(void);

/* Tests / experiments: mixed-density mode of the 3D register-stencil path (heavy tiles swept with 2 x 2 x 2
 * supercells, decided on the device).  mode -1 = default, 0 = off, 1 = on; min_points = smallest point set
 * that uses it (-1 = default 2^18); dense_tile_pts = points that make a 16^3 tile heavy (<= 0 = default 2048). */
void nfftb200_debug_mixed(int mode, int64_t min_points, int dense_tile_pts);

/* cuFFT plan cache: an LRU of handles per (device, dimension, M, type, B*C).  _clear destroys them (all
 * devices) and fails with NFFTB200_ERR_INVALID while the cache is pinned; _pin(+1/-1) is called by owners of
 * captured CUDA graphs, whose FFT kernels reference the handles' twiddle tables, and returns the pin count;
 * _size returns the number of cached handles. */
int nfftb200_plan_cache_clear(void);
int nfftb200_plan_cache_pin(int delta);
int nfftb200_plan_cache_size(void);

/* Number of kernels this library launched so far in this process (for bench.py). */
int64_t nfftb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NFFT_B200_H */
