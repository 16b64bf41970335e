// Deterministic binning of non-uniform points by (batch entry, grid tile).
//
// Replaces the reference's per-point scratch (compute_shifts_kernel / compute_psi_kernel,
// csrc/cuda/spatial_window_operations.cu:38-97 -- 4*d*(2m+3) bytes per point written to and
// re-read from HBM) with a 4-byte key and a 4-byte permutation entry per point.
//
// The cell rule is the reference's: cell_a = (int)floorf(pos_a * M)  (:50).  Keys are sorted
// with a stable LSD radix sort (8-bit digits, per-block digit histograms -> exclusive scan ->
// in-order scatter), so the permutation is a pure function of the keys: bit-reproducible and
// equal to numpy's argsort(kind="stable") (tests/test_parity_gpu.py::test_binning_is_bit_exact).
#pragma once
#include "common.cuh"

namespace nfftb200 {

constexpr int kPlanFlagWords = 8;  // plan flags: [0] points dropped by a window kernel (stale plan), [1] TMA timeouts,
                                   // [2] Geom::mixed: the point set was found clustered and its keys were refined

// ------------------------------------------------------------------------- block scan helpers
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread over the block (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *total receives the block sum.  s_warp: >= 33 uint32.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nw ? s_warp[lane] : 0u;
        uint32_t wi = warp_incl_scan(w);
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    uint32_t res = s_warp[warp] + incl - v;
    *total = s_warp[32];
    __syncthreads();
    return res;
}

// ------------------------------------------------------------------------- device-wide scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, long long count, uint32_t* __restrict__ partial,
                   const uint32_t* __restrict__ run_if = nullptr) {
    __shared__ uint32_t s_warp[33];
    if (run_if && *run_if == 0) return;  // (uniform over the grid: a skipped conditional radix pass)
    const long long base = (long long)blockIdx.x * kScanTile;
    uint32_t s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + (long long)k * kScanThreads + threadIdx.x;
        if (i < count) s += in[i];
    }
    uint32_t total;
    block_excl_scan(s, s_warp, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single block: in-place exclusive scan of partial[0..nparts), total written to partial[nparts]
__global__ void __launch_bounds__(1024)
scan_partials_kernel(uint32_t* partial, int nparts, const uint32_t* __restrict__ run_if = nullptr) {
    __shared__ uint32_t s_warp[33];
    if (run_if && *run_if == 0) return;
    uint32_t running = 0;
    for (int base = 0; base < nparts; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = i < nparts ? partial[i] : 0u;
        uint32_t total;
        uint32_t ex = block_excl_scan(v, s_warp, &total);
        if (i < nparts) partial[i] = running + ex;
        running += total;
    }
    if (threadIdx.x == 0) partial[nparts] = running;
}

// out[i] = exclusive prefix of in; out[count] = total.  in may alias out.
__global__ void __launch_bounds__(kScanThreads)
scan_final_kernel(const uint32_t* in, uint32_t* out, long long count, const uint32_t* __restrict__ partial,
                  int nparts, const uint32_t* __restrict__ run_if = nullptr) {
    __shared__ uint32_t s_warp[33];
    if (run_if && *run_if == 0) return;
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + k;
        v[k] = i < count ? in[i] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(s, s_warp, &total) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        long long i = base + k;
        if (i < count) out[i] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[count] = partial[nparts];
}

// Small arrays (bin counts, chunk counts, the digit tables of small sorts): one block, one launch
// instead of three.  Every thread owns a contiguous run of `per` items; in may alias out.
constexpr long long kScanSmallMax = 1024 * 8;  // (a 64K-entry digit table through one block costs 85 us)
__global__ void __launch_bounds__(1024)
scan_small_kernel(const uint32_t* in, uint32_t* out, int count, int per, const uint32_t* __restrict__ run_if = nullptr) {
    __shared__ uint32_t s_warp[33];
    if (run_if && *run_if == 0) return;
    const int lo = threadIdx.x * per;
    const int hi = lo + per < count ? lo + per : count;
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += in[i];
    uint32_t total;
    uint32_t ex = block_excl_scan(s, s_warp, &total);
    for (int i = lo; i < hi; ++i) {
        const uint32_t v = in[i];
        out[i] = ex;
        ex += v;
    }
    if (threadIdx.x == 0) out[count] = total;
}

inline size_t scan_scratch_bytes(long long count) {
    long long nparts = (count + kScanTile - 1) / kScanTile;
    if (nparts < 1) nparts = 1;
    return align_up((size_t)(nparts + 1) * sizeof(uint32_t));
}

// exclusive scan of `count` uint32 (count >= 0); writes count+1 entries to out.
// run_if (device, optional): the scan is skipped when *run_if == 0.
inline int scan_exclusive(const uint32_t* in, uint32_t* out, long long count, uint32_t* scratch,
                          cudaStream_t st, const uint32_t* run_if = nullptr) {
    if (count <= kScanSmallMax) {
        const int per = (int)((count + 1023) / 1024);
        NF_LAUNCH(scan_small_kernel, 1, 1024, 0, st, in, out, (int)count, per < 1 ? 1 : per, run_if);
        return NFFTB200_OK;
    }
    int nparts = (int)((count + kScanTile - 1) / kScanTile);
    if (nparts < 1) nparts = 1;
    NF_LAUNCH(scan_reduce_kernel, nparts, kScanThreads, 0, st, in, count, scratch, run_if);
    NF_LAUNCH(scan_partials_kernel, 1, 1024, 0, st, scratch, nparts, run_if);
    NF_LAUNCH(scan_final_kernel, nparts, kScanThreads, 0, st, in, out, count, scratch, nparts, run_if);
    return NFFTB200_OK;
}

// ------------------------------------------------------------------------- keys + bin counts
__device__ __forceinline__ int wrap_mod(int v, int M) {
    v %= M;
    return v < 0 ? v + M : v;
}

// key = ((b*nt[2] + tz)*nt[1] + ty)*nt[0] + tx, tile index from the wrapped reference cell.
// Powers of two (the usual M and tile edges) take mask / shift paths: the generic integer modulo and
// division cost ~25 instructions each, six per point.
struct KeyFast {
    bool m_pow2;
    int tshift[3];    // log2(T[slot]) or -1
    int sc_shift[3];  // log2(sc[slot]) (supercell extents are powers of two whenever fine key bits are used)
};

__device__ __forceinline__ KeyFast key_fast(const Geom& g) {
    KeyFast f;
    f.m_pow2 = (g.M & (g.M - 1)) == 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        f.tshift[s] = (g.T[s] & (g.T[s] - 1)) == 0 ? __ffs(g.T[s]) - 1 : -1;
        f.sc_shift[s] = __ffs(g.sc[s]) - 1;
    }
    return f;
}

// The batch entry of point i: from the per-point vector batch[n] (the reference's layout, README.md:42-46) or,
// with nfftb200's NFFTB200_BATCH_OFFSETS, from the B + 1 ascending offsets of the (sorted) point sets.
struct BatchRef {
    const int64_t* data;  // nullptr: one point set
    int offsets;          // 0: data[i] is the entry of point i;  B > 0: data[0..B] are offsets
};

__device__ __forceinline__ long long batch_of(const BatchRef& br, long long i) {
    if (!br.data) return 0;
    if (!br.offsets) return br.data[i];
    int lo = 0, hi = br.offsets;  // largest b with data[b] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(br.data + mid) <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Fine part of a key (3D register-stencil tiling: 16^3 tiles of sc[0] x sc[1] x sc[2] supercells, all counts powers
// of two): the supercell of the point inside its tile as a HIERARCHICAL index, most significant first
//   [ y bit, x bit ] per level from halves down to single supercells, then the z supercell
// (4 x 4 x 2 supercells: 7 bits, 2 x 2 x 2: 9 bits) of which the top g.fine_bits bits are kept.  Sorting by
// (tile, fine) makes the chunks a heavy tile is cut into spatially compact -- a quadrant, a supercell column,
// a z-range of it -- instead of random samples of the whole tile, so the points of a chunk share register
// blocks in the sweep (window_reg.cuh).
__device__ __forceinline__ uint32_t fine_index(int cx, int cy, int cz, const Geom& g, const KeyFast& kf) {
    // supercell extents are powers of two whenever fine bits are used (make_geom): shifts, not divisions
    const int bx = cx >> kf.sc_shift[0], by = cy >> kf.sc_shift[1], bz = cz >> kf.sc_shift[2];
    uint32_t f = 0;
    for (int b = g.fine_xy_levels - 1; b >= 0; --b) f = (f << 2) | (uint32_t)((((by >> b) & 1) << 1) | ((bx >> b) & 1));
    return (f << g.fine_z_bits) | (uint32_t)bz;
}

// known_b >= 0: the batch entry of the point is already known (a radix tile that lies inside one point set)
__device__ __forceinline__ uint32_t point_key(const float* __restrict__ pos, const BatchRef& batch,
                                              long long i, const Geom& g, const KeyFast& f, long long known_b = -1) {
    long long b = known_b >= 0 ? known_b : batch_of(batch, i);
    b = b < 0 ? 0 : (b >= g.B ? g.B - 1 : b);
    uint32_t key = (uint32_t)b;
    const float Mf = (float)g.M;
    const float* p = pos + i * g.dim;
    int in_tile[3] = {0, 0, 0};
#pragma unroll
    for (int slot = 2; slot >= 0; --slot) {
        if (slot < g.dim) {
            const int c = (int)floorf(p[g.dim - 1 - slot] * Mf);  // spatial_window_operations.cu:50
            const int cw = f.m_pow2 ? (c & (g.M - 1)) : wrap_mod(c, g.M);
            const int tile = f.tshift[slot] >= 0 ? (cw >> f.tshift[slot]) : cw / g.T[slot];
            in_tile[slot] = cw - tile * g.T[slot];
            key = key * (uint32_t)g.nt[slot] + (uint32_t)tile;
        }
    }
    if (g.fine_bits > 0) {
        // the top fine_bits bits of the hierarchical index (Geom::mixed: left-aligned in a field that may be wider)
        const int total = 2 * g.fine_xy_levels + g.fine_z_bits;
        const uint32_t fi = fine_index(in_tile[0], in_tile[1], in_tile[2], g, f);
        key = (key << g.fine_bits) | (g.fine_bits >= total ? fi << (g.fine_bits - total) : fi >> (total - g.fine_bits));
    }
    return key;
}

// The same key with the dimension and "M and the tile edges are powers of two" known at compile time: no runtime
// dimension tests, no integer divisions (ncu of the generic version at c4: 173 instructions per point, issue
// slots 81 % busy -- the key kernel was instruction-bound, profiles/r03c_misc_kernels.txt).
template <int DIM, bool POW2>
__device__ __forceinline__ uint32_t point_key_t(const float* __restrict__ pos, const BatchRef& batch, long long i,
                                                const Geom& g, const KeyFast& f, long long known_b) {
    long long b = known_b >= 0 ? known_b : batch_of(batch, i);
    b = b < 0 ? 0 : (b >= g.B ? g.B - 1 : b);
    uint32_t key = (uint32_t)b;
    const float Mf = (float)g.M;
    const float* p = pos + i * DIM;
    int in_tile[3] = {0, 0, 0};
#pragma unroll
    for (int slot = DIM - 1; slot >= 0; --slot) {
        const int c = (int)floorf(p[DIM - 1 - slot] * Mf);  // spatial_window_operations.cu:50
        int tile;
        if (POW2) {
            const int cw = c & (g.M - 1);
            tile = cw >> f.tshift[slot];
            in_tile[slot] = cw & (g.T[slot] - 1);
        } else {
            const int cw = f.m_pow2 ? (c & (g.M - 1)) : wrap_mod(c, g.M);
            tile = f.tshift[slot] >= 0 ? (cw >> f.tshift[slot]) : cw / g.T[slot];
            in_tile[slot] = cw - tile * g.T[slot];
        }
        key = key * (uint32_t)g.nt[slot] + (uint32_t)tile;
    }
    if (g.fine_bits > 0) {
        const int total = 2 * g.fine_xy_levels + g.fine_z_bits;
        const uint32_t fi = fine_index(in_tile[0], in_tile[1], in_tile[2], g, f);
        key = (key << g.fine_bits) | (g.fine_bits >= total ? fi << (g.fine_bits - total) : fi >> (total - g.fine_bits));
    }
    return key;
}

__global__ void __launch_bounds__(256)
key_hist_kernel(const float* __restrict__ pos, const BatchRef batch, long long n, Geom g,
                uint32_t* __restrict__ keys, uint32_t* __restrict__ bin_count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = point_key(pos, batch, i, g, key_fast(g));
    keys[i] = key;
    if (bin_count) atomicAdd(&bin_count[key >> g.fine_bits], 1u);  // only when no radix pass follows (single bin)
}

// Bin sizes from the SORTED keys: one atomicAdd per run of equal keys per warp (a global atomic per
// point on 2^14 hot addresses costs 0.45 ms at 2^24 points; this is ~30x fewer atomics).
__global__ void __launch_bounds__(256)
count_sorted_kernel(const uint32_t* __restrict__ keys_sorted, long long n, int fine_bits,
                    uint32_t* __restrict__ bin_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    const uint32_t key = valid ? (keys_sorted[i] >> fine_bits) : 0xffffffffu;  // bin = tile part of the key
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = valid && (lane == 0 || prev != key);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t valids = __ballot_sync(0xffffffffu, valid);
    if (head) {
        // run = [lane, next head or end of the valid lanes)
        const uint32_t above = heads & ~((2u << lane) - 1u);
        const int end = above ? __ffs(above) - 1 : 32 - __clz(valids);
        atomicAdd(&bin_count[key], (uint32_t)(end - lane));
    }
}

// The same from a binary search per bin: bin_start[b] = first sorted position whose tile key is >= b (b = 0 .. nbins).
// 24 dependent loads per bin instead of a pass over all keys plus a scan: used when there are far fewer bins than
// points (c4: 16 384 bins, 2^24 points -- 12 us instead of 56 + 15 us).
__global__ void __launch_bounds__(256)
bin_search_kernel(const uint32_t* __restrict__ keys_sorted, long long n, int fine_bits, long long nbins,
                  uint32_t* __restrict__ bin_start) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nbins) return;
    long long lo = 0, hi = n;  // first position in [0, n] with key >= b
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)(__ldg(keys_sorted + mid) >> fine_bits) < b) lo = mid + 1;
        else hi = mid;
    }
    bin_start[b] = (uint32_t)lo;
}

// ------------------------------------------------------------------------- stable radix sort
#ifndef NFFT_SORT_MINB
#define NFFT_SORT_MINB 5  // resident CTAs per SM the scatter kernel is compiled for (register cap)
#endif
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsIpt = 16;  // items per thread
constexpr int kRsTile = kRsThreads * kRsIpt;
constexpr int kRsBins = 256;

// table[digit * nblocks + block] = number of keys of `block` whose digit is `digit`
__global__ void __launch_bounds__(kRsThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys, long long n, int shift, uint32_t* __restrict__ table,
                  int nblocks, const uint32_t* __restrict__ keys_alt = nullptr, const uint32_t* __restrict__ cond = nullptr) {
    __shared__ uint32_t hist[kRsBins];
    if (cond && *cond == 0) keys = keys_alt;  // the conditional pass before this one did not run
    hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kRsTile;
    for (int k = 0; k < kRsIpt; ++k) {
        long long i = base + (long long)k * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[(long long)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Keys of one radix tile and, in the same pass, the tile's row of the first radix pass's digit table
// (saves re-reading the keys in radix_hist_kernel).
template <int DIM, bool POW2>
__global__ void __launch_bounds__(kRsThreads)
key_tile_kernel(const float* __restrict__ pos, const BatchRef batch, long long n, Geom g,
                uint32_t* __restrict__ keys, uint32_t* __restrict__ table, int nblocks,
                uint32_t* __restrict__ sample_counts = nullptr, int sample_mask = 0) {
    __shared__ uint32_t hist[kRsBins];
    __shared__ long long s_b[2];
    hist[threadIdx.x] = 0;
    const long long base = (long long)blockIdx.x * kRsTile;
    // offsets: the point sets are contiguous, so almost every radix tile lies inside ONE of them -- look the batch
    // entry up once per tile instead of with a binary search per point
    if (threadIdx.x < 2 && batch.data && batch.offsets) {
        const long long last = base + kRsTile - 1 < n ? base + kRsTile - 1 : n - 1;
        s_b[threadIdx.x] = batch_of(batch, threadIdx.x == 0 ? base : last);
    }
    __syncthreads();
    const KeyFast f = key_fast(g);
    const long long tile_b = (batch.data && batch.offsets && s_b[0] == s_b[1]) ? s_b[0] : -1;
#pragma unroll 4
    for (int k = 0; k < kRsIpt; ++k) {
        const long long i = base + (long long)k * kRsThreads + threadIdx.x;
        if (i < n) {
            const uint32_t key = point_key_t<DIM, POW2>(pos, batch, i, g, f, tile_b);
            keys[i] = key;
            atomicAdd(&hist[key & 255u], 1u);
            // Geom::mixed: every (sample_mask + 1)-th point is counted into its tile (see density_flag_kernel)
            if (sample_counts && (i & sample_mask) == 0) atomicAdd(&sample_counts[key >> g.fine_bits], 1u);
        }
    }
    __syncthreads();
    table[(long long)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// In-order scatter.  Warp w of a block owns the contiguous items [base + w*512, +512) and walks
// them 32 at a time, so (block, warp, round, lane) order == input order and equal digits keep
// their relative order (stable).  idx_in == nullptr means the identity payload.
// The block first sorts its 4096 items by digit in shared memory and then writes them out in that
// order, so consecutive threads store to consecutive addresses of each digit's run (full sectors
// instead of one 4-byte store per 32-byte sector).
__global__ void __launch_bounds__(kRsThreads, NFFT_SORT_MINB)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out, long long n, int shift,
                     const uint32_t* __restrict__ table_scanned, int nblocks,
                     const uint32_t* __restrict__ cond = nullptr, int cond_mode = 0,
                     const uint32_t* __restrict__ keys_alt = nullptr) {
    // Conditional passes (Geom::mixed, see sort_points): cond_mode 1 = this pass runs only if *cond != 0;
    // cond_mode 2 = the pass before this one was conditional: if it did not run, read its input instead
    // (keys_alt, identity payload).
    if (cond_mode == 1 && *cond == 0) return;
    if (cond_mode == 2 && *cond == 0) {
        keys_in = keys_alt;
        idx_in = nullptr;
    }
    __shared__ uint32_t wcnt[kRsWarps][kRsBins];
    __shared__ uint32_t s_key[kRsTile];
    __shared__ uint32_t s_idx[kRsTile];
    __shared__ uint32_t s_loc[kRsBins];    // first position of a digit in the block-sorted order
    __shared__ uint32_t s_gbase[kRsBins];  // first global position of this block's items of a digit
    __shared__ uint32_t s_warp[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < kRsWarps * kRsBins; k += kRsThreads) (&wcnt[0][0])[k] = 0;
    __syncthreads();

    const long long blk = (long long)blockIdx.x * kRsTile;
    const long long base = blk + (long long)warp * (32 * kRsIpt);
    // the keys are read again in the second phase (L2 hits) instead of being held in 16 more registers:
    // the kernel is latency-bound and 5 resident CTAs per SM beat 2
    uint32_t rank[kRsIpt];
#pragma unroll
    for (int s = 0; s < kRsIpt; ++s) {
        const long long i = base + s * 32 + lane;
        const bool valid = i < n;
        const uint32_t key = valid ? keys_in[i] : 0xffffffffu;
        const uint32_t digit = valid ? ((key >> shift) & 255u) : 256u;
        const uint32_t mask = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(mask) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = wcnt[warp][digit];
            wcnt[warp][digit] = old + __popc(mask);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[s] = old + __popc(mask & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    uint32_t total_d = 0;
    {
        // exclusive prefix over the warps of this block (block-local), per digit
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t t = wcnt[w][d];
            wcnt[w][d] = total_d;
            total_d += t;
        }
        s_gbase[d] = table_scanned[(long long)d * nblocks + blockIdx.x];
    }
    uint32_t block_total;
    const uint32_t loc = block_excl_scan(total_d, s_warp, &block_total);  // ends with __syncthreads
    s_loc[threadIdx.x] = loc;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kRsIpt; ++s) {
        const long long i = base + s * 32 + lane;
        if (i < n) {
            const uint32_t key = keys_in[i];
            const uint32_t digit = (key >> shift) & 255u;
            const uint32_t lp = s_loc[digit] + wcnt[warp][digit] + rank[s];
            s_key[lp] = key;
            s_idx[lp] = idx_in ? idx_in[i] : (uint32_t)i;
        }
    }
    __syncthreads();
    for (uint32_t lp = threadIdx.x; lp < block_total; lp += kRsThreads) {
        const uint32_t k = s_key[lp];
        const uint32_t digit = (k >> shift) & 255u;
        const uint32_t dst = s_gbase[digit] + (lp - s_loc[digit]);
        keys_out[dst] = k;
        idx_out[dst] = s_idx[lp];
    }
}

// Geom::mixed -- is the point set clustered?  Every `stride`-th key is counted into its tile (key_tile_kernel); the set is
// "clustered" (flag = 1) when at least 1/8 of the sampled points lie in tiles that hold >= dense_pts points.
// Only then does the radix sort run its low pass over the fine key bits (which makes the chunks of a heavy tile
// compact) and only then are heavy tiles marked for the 2 x 2 x 2 sweep (fill_items_kernel).
__global__ void __launch_bounds__(1024)
density_flag_kernel(const uint32_t* __restrict__ counts, long long nbins, int stride, int dense_pts, long long n,
                    uint32_t* __restrict__ flag) {
    __shared__ uint32_t s_warp[33];
    uint32_t heavy = 0;
    for (long long b = threadIdx.x; b < nbins; b += 1024) {
        const uint32_t c = counts[b];
        if ((unsigned long long)c * stride >= (unsigned long long)dense_pts) heavy += c;
    }
    uint32_t total;
    block_excl_scan(heavy, s_warp, &total);
    if (threadIdx.x == 0) *flag = (unsigned long long)total * stride * 8ull >= (unsigned long long)n ? 1u : 0u;
}

__global__ void __launch_bounds__(256) iota_kernel(uint32_t* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------- work items
// number of chunks per bin: ceil(count / pmax)
__global__ void __launch_bounds__(256)
chunk_count_kernel(const uint32_t* __restrict__ bin_start, long long nbins, int pmax, uint32_t* __restrict__ nch) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    uint32_t c = bin_start[b + 1] - bin_start[b];
    nch[b] = (c + (uint32_t)pmax - 1u) / (uint32_t)pmax;
}

// items[w] = {bin, first point, one past the last point, 0}: the points of a bin are split evenly
// over its chunks.  Entries beyond the last work item stay zero (empty range), so a CTA needs one
// 16-byte load to know its work.
// items[].w = class of the tile (Geom::mixed): 1 = heavy (>= dense_pts points, and the keys were refined:
// *refined != 0 or refined == nullptr), else 0.
__global__ void __launch_bounds__(256)
fill_items_kernel(const uint32_t* __restrict__ bin_start, const uint32_t* __restrict__ chunk_start, long long nbins,
                  uint4* __restrict__ items, int dense_pts = 0, const uint32_t* __restrict__ refined = nullptr) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const uint32_t lo = chunk_start[b], hi = chunk_start[b + 1];
    const uint32_t p0 = bin_start[b];
    const unsigned long long cnt = bin_start[b + 1] - p0;
    const uint32_t nch = hi - lo;
    const uint32_t cls = dense_pts > 0 && cnt >= (unsigned long long)dense_pts && (!refined || *refined != 0) ? 1u : 0u;
    for (uint32_t w = lo; w < hi; ++w) {
        const uint32_t c = w - lo;
        items[w] = make_uint4((uint32_t)b, p0 + (uint32_t)(cnt * c / nch), p0 + (uint32_t)(cnt * (c + 1) / nch), cls);
    }
}

// Small problems (one radix pass whose digit IS the bin: <= 256 bins, no fine key bits -- e.g. the 1D workload
// c2): bin offsets come straight from the scanned digit table (entry [digit][block 0] = number of keys with a
// smaller digit), so one block derives bin_start, the chunk counts, their scan and the work items: one launch
// instead of five (count_sorted, scan, chunk_count, scan, fill_items) on a launch-bound path.
__global__ void __launch_bounds__(kRsBins)
finish_single_pass_kernel(const uint32_t* __restrict__ table_scanned, int nblocks, int nbins, uint32_t n, int pmax,
                          uint32_t* __restrict__ bin_start, uint32_t* __restrict__ chunk_start,
                          uint4* __restrict__ items, long long max_items, uint32_t* __restrict__ flags) {
    __shared__ uint32_t s_warp[33];
    const int b = threadIdx.x;
    if (b < kPlanFlagWords) flags[b] = 0;  // (this launch also stands in for the memsets of the general path)
    uint32_t lo = 0, hi = 0;
    if (b < nbins) {
        lo = table_scanned[(long long)b * nblocks];
        hi = b + 1 < kRsBins ? table_scanned[(long long)(b + 1) * nblocks] : n;
        bin_start[b] = lo;
        if (b == nbins - 1) bin_start[nbins] = n;
    }
    const uint32_t cnt = hi - lo;
    const uint32_t nch = b < nbins ? (cnt + (uint32_t)pmax - 1u) / (uint32_t)pmax : 0u;
    uint32_t total;
    const uint32_t first = block_excl_scan(nch, s_warp, &total);
    if (b < nbins) {
        chunk_start[b] = first;
        if (b == nbins - 1) chunk_start[nbins] = total;
        for (uint32_t c = 0; c < nch; ++c)
            items[first + c] = make_uint4((uint32_t)b, lo + (uint32_t)((unsigned long long)cnt * c / nch),
                                          lo + (uint32_t)((unsigned long long)cnt * (c + 1) / nch), 0u);
    }
    // entries beyond the last work item: empty ranges (the window kernels are launched over max_items CTAs)
    for (long long w = (long long)total + b; w < max_items; w += kRsBins) items[w] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------- host orchestration
// The binning result that the window kernels read ("point plan", persistent) and the scratch the sort
// needs while it runs (transient) are separate regions, so that a caller can keep the plan of a point
// set (NfftPlan in nfft.py: 4 bytes per point + the bin / work-item tables) and drop the 16 bytes per
// point of sort scratch.
struct PlanLayout {   // persistent: perm | bin_start | chunk_start | items | flags
    size_t perm, bin_start, chunk_start, items, flags, total;
    long long nbins, max_items;
};
struct SortLayout {   // transient scratch
    size_t keys0, keysA, keysB, idxT, bin_count, nch, table, scan, total;
    long long nblocks;
};

inline PlanLayout plan_layout(long long n, const Geom& g) {
    PlanLayout L{};
    L.nbins = (long long)g.B * g.tiles_per_batch;
    L.max_items = n / g.pmax + (L.nbins < n ? L.nbins : n) + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align_up(bytes);
        return o;
    };
    const size_t nn = (size_t)(n > 0 ? n : 1);
    L.perm = take(nn * 4);
    L.bin_start = take((size_t)(L.nbins + 1) * 4);
    L.chunk_start = take((size_t)(L.nbins + 1) * 4);
    L.items = take((size_t)L.max_items * sizeof(uint4));
    L.flags = take(kPlanFlagWords * 4);
    L.total = off;
    return L;
}

inline SortLayout sort_layout(long long n, const Geom& g) {
    SortLayout L{};
    const long long nbins = (long long)g.B * g.tiles_per_batch;
    L.nblocks = (n + kRsTile - 1) / kRsTile;
    if (L.nblocks < 1) L.nblocks = 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align_up(bytes);
        return o;
    };
    const size_t nn = (size_t)(n > 0 ? n : 1);
    L.keys0 = take(nn * 4);
    L.keysA = take(nn * 4);
    L.keysB = take(nn * 4);
    L.idxT = take(nn * 4);
    L.bin_count = take((size_t)(nbins + 1) * 4);
    L.nch = take((size_t)(nbins + 1) * 4);
    L.table = take((size_t)(kRsBins * L.nblocks + 1) * 4);
    size_t s1 = scan_scratch_bytes(nbins + 1);
    size_t s2 = scan_scratch_bytes(kRsBins * L.nblocks + 1);
    L.scan = take(s1 > s2 ? s1 : s2);
    L.total = off;
    return L;
}

inline int tile_key_bits(const Geom& g) {
    const long long nbins = (long long)g.B * g.tiles_per_batch;
    int bits = 0;
    while (bits < 32 && (1ll << bits) < nbins) ++bits;
    return bits;
}
inline int sort_passes(const Geom& g) { return (tile_key_bits(g) + g.fine_bits + 7) / 8; }

// Pointers of a point plan inside its persistent region (a pure function of n and the geometry).
inline void sort_plan_pointers(long long n, const Geom& g, char* plan_mem, SortPlan* plan) {
    const PlanLayout L = plan_layout(n, g);
    plan->keys = nullptr;
    plan->perm = (uint32_t*)(plan_mem + L.perm);
    plan->bin_start = (uint32_t*)(plan_mem + L.bin_start);
    plan->chunk_start = (uint32_t*)(plan_mem + L.chunk_start);
    plan->items = (uint4*)(plan_mem + L.items);
    plan->flags = (uint32_t*)(plan_mem + L.flags);
    plan->nbins = L.nbins;
    plan->max_items = L.max_items;
}

// Bins n points: the plan is written to `plan_mem` (plan_layout(n,g).total bytes), `scratch` must hold
// sort_layout(n,g).total bytes and is dead afterwards (plan->keys points into it: the unsorted tile keys).
inline int sort_points(const float* pos, const int64_t* batch, bool batch_is_offsets, long long n, const Geom& g,
                       char* scratch, char* plan_mem, SortPlan* plan, cudaStream_t st) {
    const SortLayout L = sort_layout(n, g);
    sort_plan_pointers(n, g, plan_mem, plan);
    const long long nbins = plan->nbins;
    uint32_t* keys0 = (uint32_t*)(scratch + L.keys0);
    plan->keys = keys0;
    const int passes = sort_passes(g);
    // the last radix pass writes its permutation straight into the plan
    uint32_t* kbuf[2] = {(uint32_t*)(scratch + L.keysA), (uint32_t*)(scratch + L.keysB)};
    uint32_t* ibuf[2] = {(uint32_t*)(scratch + L.idxT), (uint32_t*)(scratch + L.idxT)};
    ibuf[passes == 0 ? 0 : ((passes - 1) & 1)] = plan->perm;
    uint32_t* bin_count = (uint32_t*)(scratch + L.bin_count);
    uint32_t* bin_start = plan->bin_start;
    uint32_t* nch = (uint32_t*)(scratch + L.nch);
    uint32_t* chunk_start = plan->chunk_start;
    uint32_t* table = (uint32_t*)(scratch + L.table);
    uint32_t* scan = (uint32_t*)(scratch + L.scan);
    const BatchRef bref{batch, batch_is_offsets ? g.B : 0};
    const int sample_stride = n >= (1ll << 22) ? 32 : 8;  // power of two

    // single radix pass whose digit is the bin (small problems): one finishing launch, no memsets
    const bool single_pass = n > 0 && passes == 1 && g.fine_bits == 0 && nbins <= kRsBins;
    // Geom::mixed with a refine pass: pass 0 covers fine key bits only and runs only for clustered point sets
    // (device flag plan->flags[2], from a sample of the keys counted per tile into `nch`); pass 1 then reads
    // either its output or, if it did not run, the unsorted keys with the identity payload.  A uniform point set
    // pays the sampling and a few empty launches, not the pass (0.2 ms at 2^24 points).
    const bool refine = g.mixed && g.refine_pass && passes >= 2 && n > 0;
    // bin offsets by binary search in the sorted keys when there are far fewer bins than points
    const bool search_bins = passes > 0 && !single_pass && nbins * 8 <= n;
    if (!single_pass) {
        // bin_count = 0 (unless the offsets come from the search), the sample counts in nch = 0 (refine)
        if (!search_bins) NF_CUDA(cudaMemsetAsync(bin_count, 0, (size_t)(nbins + 1) * 4, st));
        if (refine) NF_CUDA(cudaMemsetAsync(nch, 0, (size_t)(nbins + 1) * 4, st));
        NF_CUDA(cudaMemsetAsync(plan->flags, 0, kPlanFlagWords * 4, st));
    }
    // stable LSD radix sort of (key, index) over the key bits that can be set
    const uint32_t* kin = keys0;
    const uint32_t* iin = nullptr;  // identity payload on the first pass
    if (n > 0) {
        if (passes == 0) {
            NF_LAUNCH(key_hist_kernel, (unsigned)((n + 255) / 256), 256, 0, st, pos, bref, n, g, keys0, bin_count);
            NF_LAUNCH(iota_kernel, (unsigned)((n + 255) / 256), 256, 0, st, plan->perm, n);
        } else {
            bool pow2 = (g.M & (g.M - 1)) == 0;
            for (int sl = 0; sl < g.dim; ++sl) pow2 = pow2 && (g.T[sl] & (g.T[sl] - 1)) == 0;
#define NF_KEY_TILE(D_, P_)                                                                                    \
            NF_LAUNCH((key_tile_kernel<D_, P_>), (unsigned)L.nblocks, kRsThreads, 0, st, pos, bref, n, g, keys0, \
                      table, (int)L.nblocks, refine ? nch : nullptr, sample_stride - 1)
            if (g.dim == 1) { if (pow2) NF_KEY_TILE(1, true); else NF_KEY_TILE(1, false); }
            else if (g.dim == 2) { if (pow2) NF_KEY_TILE(2, true); else NF_KEY_TILE(2, false); }
            else { if (pow2) NF_KEY_TILE(3, true); else NF_KEY_TILE(3, false); }
#undef NF_KEY_TILE
        }
        uint32_t* refined = plan->flags + 2;
        if (refine)
            NF_LAUNCH(density_flag_kernel, 1, 1024, 0, st, nch, nbins, sample_stride, g.dense_tile_pts, n, refined);
        for (int p = 0; p < passes; ++p) {
            uint32_t* kout = kbuf[p & 1];
            uint32_t* iout = ibuf[p & 1];
            const bool cond_run = refine && p == 0, cond_in = refine && p == 1;
            if (p > 0) {
                NF_LAUNCH(radix_hist_kernel, (unsigned)L.nblocks, kRsThreads, 0, st, kin, n, 8 * p, table,
                          (int)L.nblocks, cond_in ? keys0 : nullptr, cond_in ? refined : nullptr);
            }
            NF_TRY(scan_exclusive(table, table, kRsBins * L.nblocks, scan, st, cond_run ? refined : nullptr));
            NF_LAUNCH(radix_scatter_kernel, (unsigned)L.nblocks, kRsThreads, 0, st, kin, iin, kout, iout, n, 8 * p,
                      table, (int)L.nblocks, (cond_run || cond_in) ? refined : nullptr, cond_run ? 1 : (cond_in ? 2 : 0),
                      cond_in ? keys0 : nullptr);
            kin = kout;
            iin = iout;
        }
        if (single_pass) {
            // the digit is the bin: everything after the scatter in one small launch
            NF_LAUNCH(finish_single_pass_kernel, 1, kRsBins, 0, st, table, (int)L.nblocks, (int)nbins, (uint32_t)n,
                      g.pmax, bin_start, chunk_start, plan->items, plan->max_items, plan->flags);
            return NFFTB200_OK;
        }
        if (search_bins) {
            NF_LAUNCH(bin_search_kernel, (unsigned)((nbins + 256) / 256), 256, 0, st, kin, n, g.fine_bits, nbins, bin_start);
        } else if (passes > 0) {
            NF_LAUNCH(count_sorted_kernel, (unsigned)((n + 255) / 256), 256, 0, st, kin, n, g.fine_bits, bin_count);
        }
    }
    // bin offsets, chunks of at most pmax points, work items
    if (!search_bins) NF_TRY(scan_exclusive(bin_count, bin_start, nbins, scan, st));
    NF_LAUNCH(chunk_count_kernel, (unsigned)((nbins + 255) / 256), 256, 0, st, bin_start, nbins, g.pmax, nch);
    NF_TRY(scan_exclusive(nch, chunk_start, nbins, scan, st));
    NF_CUDA(cudaMemsetAsync(plan->items, 0, (size_t)plan->max_items * sizeof(uint4), st));
    NF_LAUNCH(fill_items_kernel, (unsigned)((nbins + 255) / 256), 256, 0, st, bin_start, chunk_start, nbins,
              plan->items, g.mixed ? g.dense_tile_pts : 0,
              g.mixed && g.refine_pass && sort_passes(g) >= 2 ? plan->flags + 2 : nullptr);
    return NFFTB200_OK;
}

}  // namespace nfftb200
