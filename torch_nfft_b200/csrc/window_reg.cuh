// Register-stencil window kernels (3D): the tap loop runs entirely in registers.
//
// The team kernels of window.cuh pay one shared-memory read-modify-write (spread) or read
// (gather) per window tap, which makes them LSU-bound (ncu: l1tex data-pipe wavefronts ~89 %).
// Here the chunk's points are first bucketed (in shared memory) by "supercell" = SX x SY x SZ
// oversampled cells, ordered as columns along Z.  One warp sweeps a column of supercells bottom
// to top and keeps the (L+SX-1) x (L+SY-1) x (L+SZ-1) block of grid cells its current supercell
// touches in registers: lane <-> (x, y) positions of the block, WZ consecutive z cells per
// position.  Each point costs one FFMA per block cell with its zero-padded, shifted tap vectors;
// the shared-memory tile is touched only when the sweep advances (SZ finished planes are added
// out / SZ new planes are loaded, the rest of the block slides inside the register file).
//
// Products are formed as ((x * psi_y) * psi_x) * psi_z (spread) and (psi_y * psi_x) * sum_z
// (gather); the reference multiplies dimension 0 first (spatial_window_operations.cu:146-156,
// 257-267).  The difference is a rounding of ~6e-8 per tap, far below the 1e-5 parity budget.
#pragma once
#include <cuda.h>  // CUtensorMap (type only: the encoder is looked up at run time, nfft_b200.cu)

#include "window.cuh"

namespace nfftb200 {

// ----------------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) on the real oversampled grid, described by ONE rank-4 tensor map per launch:
//   dims {M, M, M, B*C} (X fastest), box {P0, P1, 1, 1} = one plane of a CTA's padded tile, dense in shared
//   memory (row pitch P0 floats, planes 128-byte aligned: Geom::sY = P0, sZ = roundup(P0 * P1, 32)).
// * spread: a finished plane pair of the shared-memory tile is ADDED into the grid by the TMA unit
//   (cp.reduce.async.bulk.tensor .add, SASS UTMAREDG) as soon as the last warp has added into it, while the
//   warps go on sweeping: no register staging, no LSU traffic, no flush phase after the sweep.  Tiles that
//   cross the periodic boundary in X or Y keep the vector-reduction flush (negative box coordinates fault).
// * gather: the tile's planes are LOADED by the TMA unit (UTMALDG) into the shared-memory tile, completion on
//   an mbarrier, while the CTA buckets its points.  Out-of-range elements of a load are zero-FILLED, which
//   would overwrite the other half of a wrapped plane, so tiles that cross the grid boundary in X or Y (2 of
//   16 tile rows per dimension at c4) keep the LDGSTS path; the Z wrap is per plane and free.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_reduce_add_plane(const CUtensorMap* tmap, uint32_t smem_src, int x, int y, int z,
                                                     int plane) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
        "r"(smem_src), "r"(x), "r"(y), "r"(z), "r"(plane)
        : "memory");
}
__device__ __forceinline__ void tma_load_plane(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t mbar, int x, int y,
                                               int z, int plane) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_dst),
        "l"(tmap), "r"(mbar), "r"(x), "r"(y), "r"(z), "r"(plane)
        : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have READ their shared-memory source (the CTA may exit / reuse it)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory (the warps' STS) become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// Bounded: a transfer that never completes (a bug) must not hang the GPU; returns false then.
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
// first byte of the 128-byte aligned tile inside the dynamic shared memory (TMA needs 128-byte alignment)
__device__ __forceinline__ float* align_tile(float* smem) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    return smem + (((128u - (base & 127u)) & 127u) >> 2);
}
constexpr int kTmaParamWords = 8;  // s_tma: x0, y0, dx, dy, z0, plane (per CTA, read by the flushing lanes)

// Hands one finished plane pair of the spread tile to the TMA unit (executed by ONE lane).  Out of line on
// purpose: the sweep has seven add-out sites, and the inlined copies (110 instructions each) pushed the kernel
// from 65 KB to 98 KB of code -- instruction-cache misses in the hot loop cost more than the flush saved.
__device__ __noinline__ void tma_flush_plane_pair(const CUtensorMap* tmap, uint32_t tile_s, const int* s_tma, int pr,
                                                  int nplanes, int plane_floats, int M) {
    fence_acq_rel_cta();       // the other units' tile updates (ordered before their completion counts)
    fence_proxy_async_smem();  // ... become visible to the async proxy
    const int x0 = s_tma[0], y0 = s_tma[1], dx = s_tma[2], dy = s_tma[3], plane = s_tma[5];
    for (int zz = 2 * pr; zz < 2 * pr + 2 && zz < nplanes; ++zz) {
        int gz = s_tma[4] + zz;
        gz = gz < 0 ? gz + M : (gz >= M ? gz - M : gz);
        const uint32_t src = tile_s + 4u * (uint32_t)(zz * plane_floats);
        tma_reduce_add_plane(tmap, src, x0, y0, gz, plane);
        if (dx) tma_reduce_add_plane(tmap, src, x0 + dx, y0, gz, plane);
        if (dy) tma_reduce_add_plane(tmap, src, x0, y0 + dy, gz, plane);
        if (dx && dy) tma_reduce_add_plane(tmap, src, x0 + dx, y0 + dy, gz, plane);
    }
    bulk_commit_group();
}

#ifndef NFFT_REG_THREADS
#define NFFT_REG_THREADS 256
#endif
constexpr int kRegThreads = NFFT_REG_THREADS;
constexpr int kRegWarps = kRegThreads / 32;
#ifndef NFFT_REG_MAXPTS
#define NFFT_REG_MAXPTS 1536
#endif
constexpr int kRegMaxPts = NFFT_REG_MAXPTS;  // points per work item (chunk) held in shared memory
#ifndef NFFT_REG_GROUP
#define NFFT_REG_GROUP 8
#endif
#ifndef NFFT_REG_FFMA2
#define NFFT_REG_FFMA2 1
#endif
// supercell of the 3D sweep (oversampled cells along X, Y, Z); 2 bits per offset in s_off
#ifndef NFFT_REG_SX
#define NFFT_REG_SX 4
#endif
#ifndef NFFT_REG_SY
#define NFFT_REG_SY 4
#endif
#ifndef NFFT_REG_SZ
#define NFFT_REG_SZ 2
#endif
constexpr int kRegSX = NFFT_REG_SX, kRegSY = NFFT_REG_SY, kRegSZ = NFFT_REG_SZ;
static_assert(kRegSX <= 4 && kRegSY <= 4 && kRegSZ <= 4, "cell offsets inside a supercell are stored in 2 bits");
constexpr int kRegGroup = NFFT_REG_GROUP;  // points staged per warp round: one lane per (point, dimension), 3 * 8 <= 32
static_assert(kRegGroup == 8, "the sweeps have one point body per slot of an 8-point round");
constexpr int kGatherSlots = 8;  // point slots of a gather round (= kRegGroup)

// Debug build (-DNFFT_PHASE_TIMING): thread 0 of every CTA adds the clock64() length of its phases to
// g_phase[kernel][phase]; read back through nfftb200_debug_phase_read.  Phases: 0 zero + bucket +
// order, 1 column sweep (until this warp is done), 2 wait for the other warps, 3 flush / store,
// 4 number of CTAs.
#ifdef NFFT_PHASE_TIMING
__device__ unsigned long long g_phase[2][24];  // [8 + w]: sweep length of warp w
#define NFFT_PHASE_MARK(var) const long long var = clock64()
#define NFFT_PHASE_ADD(kern, ph, t0, t1) \
    if (threadIdx.x == 0) atomicAdd(&g_phase[kern][ph], (unsigned long long)((t1) - (t0)))
#define NFFT_PHASE_WARP(kern, t0) \
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase[kern][8 + (threadIdx.x >> 5)], (unsigned long long)(clock64() - (t0)))
// [16] longest CTA (cycles), [17] / [18] first CTA start / last CTA end (globaltimer ns, [17] stored negated
// so that a zeroed counter works with atomicMax), [19..22] CTAs by points: <= 1/4, 1/2, 3/4, 1 of kRegMaxPts
__device__ __forceinline__ unsigned long long phase_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NFFT_PHASE_BEGIN(kern) \
    if (threadIdx.x == 0) atomicMax(&g_phase[kern][17], ~phase_ns())
#define NFFT_PHASE_END(kern, t0, t1, cnt)                                                  \
    if (threadIdx.x == 0) {                                                                \
        atomicMax(&g_phase[kern][16], (unsigned long long)((t1) - (t0)));                  \
        atomicMax(&g_phase[kern][18], phase_ns());                                         \
        atomicAdd(&g_phase[kern][19 + min(3, (int)((cnt) * 4 / (kRegMaxPts + 1)))], 1ull); \
    }
#else
#define NFFT_PHASE_MARK(var)
#define NFFT_PHASE_ADD(kern, ph, t0, t1)
#define NFFT_PHASE_WARP(kern, t0)
#define NFFT_PHASE_BEGIN(kern)
#define NFFT_PHASE_END(kern, t0, t1, cnt)
#endif

template <int LC, int SX, int SY, int SZ>
struct RegCfg {
    static constexpr int WX = LC + SX - 1, WY = LC + SY - 1, WZ = LC + SZ - 1;
    static constexpr int WMAX = WX > WY ? (WX > WZ ? WX : WZ) : (WY > WZ ? WY : WZ);
    static constexpr int WP = (WMAX + 1 + 3) / 4 * 4;  // window pitch; entry WP-1 is always zero
    static constexpr int COLS = WX * WY;               // (x, y) positions of the register block
    static constexpr int CPL = (COLS + 31) / 32;       // positions per lane
    static constexpr int ZQ = (WZ + 3) / 4;            // float4 loads per z window
    static constexpr int ZP = ZQ * 2;                  // float2 (packed fp32x2) accumulators per position
    static_assert(SZ % 2 == 0, "the block slides by whole float2 pairs");
    // tap windows of one staging round in shared memory (per warp):
    //   x / y windows: 2 * kRegGroup rows of XYP floats (odd pitch: conflict-free lane-per-window stores,
    //                  entry XYP-1 is always zero), read as scalars
    //   z windows    : kRegGroup rows of ZWP floats (16-byte aligned rows, read as float4)
    static constexpr int XYP = (WX > WY ? WX : WY) + 1 + (((WX > WY ? WX : WY) + 1) % 2 == 0 ? 1 : 0);
    static constexpr int ZWP = ZQ * 4 + 8;  // 20 floats: rows 16-byte aligned, row starts spread over the banks
    // Row skipping.  The (x, y) positions c = lane + 32 q of group q = 0 lie in rows 0 .. ROW_FIRST_MAX, those of the
    // last group in rows ROW_LAST_MIN .. WY-1, and a point whose y offset inside its supercell is oy has taps in rows
    // [oy, oy + LC) only.  For 4 x 4 x 2 supercells and L = 10 (13 rows: group 0 = rows 0-2, group 5 = row 12) EXACTLY
    // ONE of the two groups is all zero for every point (oy <= 2: the last, oy = 3: the first), so the point body
    // exists in two variants of CPL - 1 groups each (30 instead of 36 FFMA2), chosen per point by a warp-uniform
    // mask the tap staging returns (a warp ballot over the y lanes of the round: no shared-memory round trip in
    // front of the branch).  Two whole bodies, not guards around one group: ptxas if-converts small guarded blocks
    // into predicated FFMA2s that still issue.
    static constexpr int ROW_FIRST_MAX = 31 / WX;
    static constexpr int ROW_LAST_MIN = (32 * (CPL - 1)) / WX;
#ifndef NFFT_REG_ROWSKIP
#define NFFT_REG_ROWSKIP 1
#endif
// Measured on the B200 at c4 (profiles/r02i_ab.txt): gather 3.23 -> 3.05 ms, but the SPREAD gets slower with the two
// variants (4.02 -> 4.94 ms; no spills, fewer instructions per point -- the accumulators live across points there and
// the second body order costs more than the six FFMA2s save), so only the gather uses them.
#ifndef NFFT_REG_ROWSKIP_SPREAD
#define NFFT_REG_ROWSKIP_SPREAD 0
#endif
    static constexpr bool ROWSKIP = NFFT_REG_ROWSKIP && CPL > 2 && ROW_FIRST_MAX + 1 == ROW_LAST_MIN - LC + 1 &&
                                    ROW_FIRST_MAX + 1 == SY - 1;  // oy <= ROW_FIRST_MAX <=> last group empty; else first
    static constexpr int WIN_FLOATS = (2 * kRegGroup * XYP + 3) / 4 * 4 + kRegGroup * ZWP;
};

// exp(x) of the window taps, x in [-(m + 1)^2 * 0.75 pi / m, 0].
//   0: expf (<= 1 ulp, ~10 instructions)
//   1: 2^(x log2 e) with the product split into hi + lo parts, MUFU.EX2 (2 ulp, 3 instructions)
//   2: 2^(x * log2 e), MUFU.EX2 (2 instructions)
// Measured against the fp64 oracle (scripts/parity_probe.py): 2.0e-7 relative L2 with 0, 2.4e-7 with 1 or 2
// (fp32 accumulation dominates; tolerance 1e-5); c4 pair 9.77 -> 9.55 ms with 1.
#ifndef NFFT_WINDOW_EXP
#define NFFT_WINDOW_EXP 1
#endif
__device__ __forceinline__ float window_exp(float x) {
#if NFFT_WINDOW_EXP == 0
    return expf(x);
#else
#if NFFT_WINDOW_EXP == 1
    const float t = fmaf(x, 1.4426950216293335f, x * 1.9259629911266175e-8f);
#else
    const float t = x * 1.4426950408889634f;
#endif
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
#endif
}

// Tap recurrence (pow2 grids): the L taps phi(frac + j), j = m .. -(m + 1), of one point and dimension are
//   amp exp(-(frac + j)^2 inv_b) = [amp exp(-frac^2 inv_b)] * exp(-2 frac inv_b)^j * exp(-j^2 inv_b)
// = e0 * u^j * k_j: THREE exponentials (e0, u, 1/u) and ~2.5 multiplications per tap instead of one exponential
// and six other instructions per tap (k_j = Geom::kexp, constant-bank operands).  The powers are built by
// multiplication, so the rounding of u enters a tap |j| times: <= ~20 ulp on the outermost taps (values <= 1e-6
// of the centre), ~4 ulp on the central ones.  Measured (profiles/r02p_ab.txt): error against the fp64 oracle
// unchanged (2.3e-7 adjoint, 1.9e-7 forward at N = 32, m = 4), c4 spread 4.03 -> 3.96 ms, gather 3.05 -> 3.01 ms.
#ifndef NFFT_WINDOW_RECUR
#define NFFT_WINDOW_RECUR 1
#endif
template <int LC, typename Store>
__device__ __forceinline__ void window_taps_recur(const Geom& g, float frac, float amp, Store store) {
    constexpr int m = (LC - 2) / 2;
    const float a = frac * g.inv_b;
    const float e0 = window_exp(-frac * a) * amp;
    float up[m + 2], vp[m + 2], ek[m + 2];
    up[1] = window_exp(-2.f * a);
    vp[1] = window_exp(2.f * a);
#pragma unroll
    for (int j = 2; j <= m + 1; ++j) {
        up[j] = up[j / 2] * up[j - j / 2];
        vp[j] = vp[j / 2] * vp[j - j / 2];
    }
#pragma unroll
    for (int j = 0; j <= m + 1; ++j) ek[j] = e0 * g.kexp[j];
#pragma unroll
    for (int l = 0; l < LC; ++l) {
        const int j = m - l;  // tap l sits at frac + (m - l)
        store(l, j == 0 ? ek[0] : (j > 0 ? ek[j] * up[j] : ek[-j] * vp[-j]));
    }
}

// Shared-memory spin lock of the add-out (lane 0 of the warp).  Experiment switches for the next A/B run
// (the r01e capture counts 11 CAS attempts per acquisition): NFFT_REG_LOCK_TTAS=1 polls the lock word
// with a plain volatile load and tries the CAS only when it reads free; NFFT_REG_LOCK_NS is the back-off.
#ifndef NFFT_REG_LOCK_TTAS
#define NFFT_REG_LOCK_TTAS 0
#endif
#ifndef NFFT_REG_LOCK_NS
#define NFFT_REG_LOCK_NS 0
#endif
#ifndef NFFT_REG_LOCK_EXCH
#define NFFT_REG_LOCK_EXCH 0
#endif
// Experiment: warp w of a CTA starts its sweep w * NFFT_REG_STAGGER_NS later, so that the warps (which all start
// at the bottom of their columns after the same barrier) do not meet at the same plane-pair lock.
#ifndef NFFT_REG_STAGGER_NS
#define NFFT_REG_STAGGER_NS 0
#endif
__device__ __forceinline__ void lock_acquire(int* lk) {
#if NFFT_REG_LOCK_TTAS
    for (;;) {
        if (*reinterpret_cast<volatile int*>(lk) == 0 && atomicCAS(lk, 0, 1) == 0) break;
        __nanosleep(NFFT_REG_LOCK_NS);
    }
#elif NFFT_REG_LOCK_EXCH
    while (atomicExch(lk, 1) != 0) __nanosleep(NFFT_REG_LOCK_NS);  // ATOMS.EXCH instead of the CAS form
#else
    while (atomicCAS(lk, 0, 1) != 0) __nanosleep(NFFT_REG_LOCK_NS);
#endif
}

// orders the tile updates of a critical section before the lock release (CTA scope).
// fence.acq_rel is enough; __threadfence_block() is the sequentially consistent fence.sc.cta.
__device__ __forceinline__ void release_fence() {
#ifdef NFFT_REG_SC_FENCE
    __threadfence_block();
#else
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}

// packed fp32x2 FMA (sm_100 FFMA2): d = a * b + c on both halves
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if NFFT_REG_FFMA2
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// Shared-memory accesses of the sweeps go through opaque 32-bit shared-window addresses
// (`asm volatile("" : "+r"(addr))` + ld/st.shared): at 128 registers ptxas otherwise rematerialises the
// per-warp window base inside every point slot (S2R / S2UR / LDC / ULEA / IMAD: 9 of 69 instructions)
// and rebuilds every tile address of the add-out inside its critical section (175 -> 113 instructions).
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
template <int OFF>
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
__device__ __forceinline__ float lds_at(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts_at(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// 4-byte asynchronous global -> shared copy (LDGSTS): no register staging, completion is awaited with
// cp_async_wait_all() before the CTA barrier that publishes the tile
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src) {  // both 16-byte aligned
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
template <int BASE, int ZQ, int I = 0>
__device__ __forceinline__ void lds_window(uint32_t addr, float2* wz) {  // ZQ quads -> 2 ZQ packed pairs
    if constexpr (I < ZQ) {
        const float4 w4 = lds_f32x4<BASE + 16 * I>(addr);
        wz[2 * I] = make_float2(w4.x, w4.y);
        wz[2 * I + 1] = make_float2(w4.z, w4.w);
        lds_window<BASE, ZQ, I + 1>(addr, wz);
    }
}

template <int K>
struct IntC { static constexpr int value = K; };

// 3D tile walk for the register-stencil kernels: f(smem_offset, global_cell) for every group of 4
// consecutive X cells of the padded tile, like for_each_quad<3>, but a thread keeps one X quad and
// steps through the (y, z) rows, so that an iteration costs two wraps and four multiply-adds instead of
// two divisions and three integer modulo operations (ncu source view of the previous flush: 145
// instructions per quad, 11.7 % of all warp instructions of the spread kernel).
#ifndef NFFT_REG_WALK_UNROLL
#define NFFT_REG_WALK_UNROLL 4
#endif
constexpr int kWalkUnroll = NFFT_REG_WALK_UNROLL;
template <typename F>
__device__ __forceinline__ void for_each_quad3_rows(const Geom& g, const TileCtx& t, F f) {
    const int nx4 = g.P[0] >> 2;
    int shift = 0;
    while ((1 << shift) < nx4) ++shift;                    // lanes per row: the next power of two
    if ((blockDim.x >> shift) == 0) {                      // rows wider than the CTA: generic walk
        for_each_quad<3>(g, t, f);
        return;
    }
    const int xq = threadIdx.x & ((1 << shift) - 1);
    if (xq >= nx4) return;
    const int M = g.M, P1 = g.P[1];
    const int x = xq << 2;
    const int rows = P1 * g.P[2], rstep = blockDim.x >> shift;
    const int row0 = threadIdx.x >> shift;
    const int z0 = fast_div(row0, P1, 1.0f / (float)P1), y0 = row0 - z0 * P1;
    const int dz = rstep / P1, dy = rstep - dz * P1;
    // two copies of the walk (power-of-two grids wrap with a mask) so that the modulo path is not
    // evaluated and then discarded by a select
    auto walk = [&](auto p2) {
        constexpr bool kPow2 = decltype(p2)::value != 0;
        auto wrap = [&](int v) { return kPow2 ? (v & (M - 1)) : wrap_mod(v, M); };
        const int gx = wrap(t.org[0] + x);
        int y = y0, z = z0;
        // unrolled: consecutive quads use different registers, so a quad's loads do not wait for the
        // previous quad's global reduction to have read its operands (long-scoreboard stalls of the flush)
#if NFFT_REG_WALK_UNROLL > 1
#pragma unroll kWalkUnroll
#endif
        for (int row = row0; row < rows; row += rstep) {
            const int gy = wrap(t.org[1] + y), gz = wrap(t.org[2] + z);
            const long long cell = (long long)(gz * M + gy) * M + gx;
            f(x + y * g.sY + z * g.sZ, cell);
            y += dy;
            z += dz;
            if (y >= P1) {
                y -= P1;
                ++z;
            }
        }
    };
    if ((M & (M - 1)) == 0) walk(IntC<1>{});
    else walk(IntC<0>{});
}

inline size_t reg_smem_bytes(const Geom& g, int nsc, int win_floats) {
    // tile | points (float4) | offsets (u8) | per-warp windows | supercell start[nsc+2], cursor[nsc+2]
    return (size_t)g.tile_elems * 4 + (size_t)kRegMaxPts * 16 + (size_t)kRegWarps * win_floats * 4 +
           (size_t)(2 * nsc + 4) * 4 + (size_t)kRegMaxPts + 64 + 128;  // + 128: tile aligned for TMA
}

// Loads the chunk's points, buckets them by supercell (column-major: z fastest) and leaves them in
// s_pts as (pos0, pos1, pos2, w), w = x value (spread) or original index (gather), together with
// the cell offsets inside the supercell s_off = ox | oy << 2 | oz << 4.
template <int SX, int SY, int SZ, bool SPREAD>
__device__ __forceinline__ void bucket_points(const Geom& g, const WindowArgs& a, const TileCtx& t, int cnt,
                                              int nsx, int nsy, int nsz, float4* s_pts, unsigned char* s_off,
                                              int* s_start, int* s_cur) {
    constexpr int kPer = (kRegMaxPts + kRegThreads - 1) / kRegThreads;
    const int nsc = nsx * nsy * nsz;
    float4 pt[kPer];
    int sc[kPer];
    // tile-local cell 0 in wrapped grid coordinates
    const int lo0 = t.org[0] + g.org[0], lo1 = t.org[1] + g.org[1], lo2 = t.org[2] + g.org[2];
    const float Mf = (float)g.M;
    // (Issuing all index loads, then all position loads, before the first use was measured and is
    // slower: this phase overlaps the other resident CTA's sweep, see DESIGN.md.)
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int e = threadIdx.x + k * kRegThreads;
        sc[k] = -1;
        if (e < cnt) {
            const uint32_t i = a.perm[t.p_lo + e];
            const float p0 = a.pos[(size_t)i * 3 + 0], p1 = a.pos[(size_t)i * 3 + 1], p2 = a.pos[(size_t)i * 3 + 2];
            float w;
            if (SPREAD) w = a.xin[(size_t)i * g.K + a.k0];
            else w = __int_as_float((int)i);
            pt[k] = make_float4(p0, p1, p2, w);
            const int cz = wrap_mod((int)floorf(p0 * Mf), g.M) - lo2;  // API dim 0 = slot Z
            const int cy = wrap_mod((int)floorf(p1 * Mf), g.M) - lo1;
            const int cx = wrap_mod((int)floorf(p2 * Mf), g.M) - lo0;
            const int bx = cx / SX, by = cy / SY, bz = cz / SZ;
            // a point outside its tile can only come from a stale / foreign plan: drop it rather
            // than index shared memory out of bounds, and count it in the plan's flag word
            if (cx >= 0 && cy >= 0 && cz >= 0 && bx < nsx && by < nsy && bz < nsz) {
                sc[k] = ((by * nsx + bx) * nsz + bz) | ((cx - bx * SX) | (cy - by * SY) << 2 | (cz - bz * SZ) << 4) << 24;
                atomicAdd(&s_cur[sc[k] & 0xffffff], 1);
            } else {
                note_dropped_point(a);
            }
        }
    }
    __syncthreads();
    // exclusive scan of the counts by warp 0
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
    for (int i = threadIdx.x; i < nsc; i += kRegThreads) s_cur[i] = s_start[i];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (sc[k] >= 0) {
            const int dst = atomicAdd(&s_cur[sc[k] & 0xffffff], 1);
            s_pts[dst] = pt[k];
            s_off[dst] = (unsigned char)((unsigned)sc[k] >> 24);
        }
    }
    __syncthreads();
}

// Sum NCOMP per-lane values over the warp; afterwards lane l with (l & (32/NCOMP - 1)) == 0 holds the
// total of channel l / (32/NCOMP) in v[0].  log2(NCOMP) halving exchanges + plain butterflies.
template <int NCOMP>
__device__ __forceinline__ void warp_reduce_channels(float (&v)[NCOMP], int lane) {
    int bit = 16;
#pragma unroll
    for (int n = NCOMP; n > 1; n >>= 1, bit >>= 1) {
        const bool hi = lane & bit;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
            const float send = hi ? v[k] : v[k + n / 2];
            const float keep = hi ? v[k + n / 2] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
#pragma unroll
    for (; bit > 0; bit >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
}

// Columns are processed longest first (LPT): with dynamic hand-out the last warp to finish then
// holds a short column.  s_order[rank] = column; ncols <= 64.
__device__ __forceinline__ void order_columns(const int* s_start, int ncols, int nsz, int* s_order) {
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        const int cnt = s_start[(c + 1) * nsz] - s_start[c * nsz];
        int rank = 0;
        for (int o = 0; o < ncols; ++o) {
            const int oc = s_start[(o + 1) * nsz] - s_start[o * nsz];
            rank += (oc > cnt || (oc == cnt && o < c)) ? 1 : 0;
        }
        s_order[rank] = c;
    }
    __syncthreads();
}

// Work units of the 3D sweep: the points of one column of supercells, or - when a column holds more than its
// share of the chunk's points (clustered / dense data, where the fine sort keys make a chunk spatially compact)
// - equal POINT ranges of it, so that one heavy column (or a single heavy supercell) is swept by several warps;
// every unit adds its own partial register blocks into the tile.  Unit = col | lo << 8 | hi << 20 (positions
// [lo, hi) of the bucketed point list, all inside the column); s_units[rank] is ordered longest first (LPT),
// *s_nunits counts the units (<= kRegMaxUnits) and must have been zeroed before the caller's last barrier.
#ifndef NFFT_REG_SPLIT
#define NFFT_REG_SPLIT 1
#endif
#ifndef NFFT_REG_UNITS
#define NFFT_REG_UNITS (2 * kRegWarps)  // units a chunk is cut into when its columns are uneven
#endif
constexpr int kRegMaxUnits = 96;  // >= NFFT_REG_UNITS + columns of a tile (16 with 4 x 4, 64 with 2 x 2 supercells)
static_assert(kRegMaxUnits >= NFFT_REG_UNITS + 64, "unit list: one per column plus the extra segments");
// Static shared memory of the 3D sweeps, declared ONCE in a non-template function: a kernel that carries two sweeps
// (spread/gather_reg_mixed_kernel) must not allocate it twice -- at c4 two CTAs of 114 KB dynamic shared memory
// share an SM with 2.7 KB to spare.
struct RegStatic {
    unsigned long long mbar;                             // gather: TMA tile load
    int expect[32], done[32], tma[kTmaParamWords];       // spread: TMA flush bookkeeping per plane pair
    int lock[64];                                        // spread: one lock per pair of tile planes
    int order[kRegMaxUnits], nunits, next;               // work units of the CTA, hand-out counter
};
__device__ __forceinline__ RegStatic& reg_static() {
    __shared__ __align__(8) RegStatic s;
    return s;
}
static_assert(kRegMaxPts < 4096, "unit encoding: 12 bits per point position");
static_assert(kRegMaxPts % 16 == 0, "the tap windows behind the u8 offsets must stay 16-byte aligned");
// s_expect (spread with the TMA flush, else nullptr): s_expect[p] = number of units that will add into plane
// pair p of the tile -- a unit whose first / last point lies in supercell s0 / s1 of its column adds into the
// pairs s0 * sp .. min(s1 * sp + zp, npairs) - 1, each exactly once (see advance() in the spread kernel).
// units: how many units a batch is cut into when its columns are uneven (NFFT_REG_UNITS = 2 per warp for the
// 4 x 4 x 2 sweep; one per warp for the dense 2 x 2 x 2 sweep, where every unit ends with an add-out of all its
// plane pairs under the same few locks: c5 24.5 -> 23.0 ms, profiles/r02l_c5.txt)
__device__ __forceinline__ void make_units(const int* s_start, int ncols, int nsz, int* s_units, int* s_nunits,
                                           int* s_expect = nullptr, int sp = 0, int zp = 0, int npairs = 0,
                                           int units = NFFT_REG_UNITS) {
    __shared__ int s_raw[kRegMaxUnits], s_rawcnt[kRegMaxUnits];
    const int total = s_start[ncols * nsz] - s_start[0];
    const int maxseg = NFFT_REG_SPLIT && ncols <= 64 ? 16 : 1;  // sum of segments <= NFFT_REG_UNITS + ncols <= 80
    if ((int)threadIdx.x < ncols) {
        const int c0 = threadIdx.x * nsz;
        const int lo = s_start[c0], cnt = s_start[c0 + nsz] - lo;
        int nseg = (int)(((long long)cnt * units + total / 2) / (total > 0 ? total : 1));  // round(cnt / share)
        nseg = nseg < 1 ? 1 : (nseg > maxseg ? maxseg : nseg);
        if (cnt < 2 * kRegGroup * nseg) nseg = cnt / (2 * kRegGroup) > 0 ? cnt / (2 * kRegGroup) : 1;  // >= 2 rounds each
        for (int j = 0; j < nseg && cnt > 0; ++j) {
            const int ulo = lo + (int)((long long)cnt * j / nseg), uhi = lo + (int)((long long)cnt * (j + 1) / nseg);
            if (uhi > ulo) {
                const int k = atomicAdd(s_nunits, 1);
                s_raw[k] = (int)threadIdx.x | ulo << 8 | uhi << 20;
                s_rawcnt[k] = uhi - ulo;
            }
        }
    }
    __syncthreads();
    const int n = *s_nunits;
    if ((int)threadIdx.x < n) {
        const int cnt = s_rawcnt[threadIdx.x];
        int rank = 0;
        for (int o = 0; o < n; ++o) {
            const int oc = s_rawcnt[o];
            rank += (oc > cnt || (oc == cnt && o < (int)threadIdx.x)) ? 1 : 0;
        }
        s_units[rank] = s_raw[threadIdx.x];
        if (s_expect) {
            const int unit = s_raw[threadIdx.x];
            const int c0 = (unit & 0xff) * nsz, lo = (unit >> 8) & 0xfff, hi = (int)((unsigned)unit >> 20);
            int s0 = 0;
            while (s_start[c0 + s0 + 1] <= lo) ++s0;
            int s1 = s0;
            while (s_start[c0 + s1 + 1] < hi) ++s1;  // supercell of the unit's last point hi - 1
            const int last = s1 * sp + zp < npairs ? s1 * sp + zp : npairs;
            for (int p = s0 * sp; p < last; ++p) atomicAdd(&s_expect[p], 1);
        }
    }
    __syncthreads();
}

// Phase A of a warp round: the taps of up to kRegGroup points are evaluated and stored at their
// shifted positions inside zero-initialised windows.  Lane <-> (point, dimension): each lane runs
// L independent exp chains (unrolled), so the latency of one tap hides behind the others.
// SCALE_Z: the z taps carry the point's value (spread), so the sweep multiplies only psi(Y) * psi(X).
// All shared-memory traffic goes through two opaque shared-window addresses -- `wbase` (this warp's windows)
// and `pts_sh` (the CTA's point records, with the u8 cell offsets kRegMaxPts * 16 bytes behind them): at 128
// registers ptxas otherwise rebuilds the three array bases from the kernel parameters in every round
// (S2R / S2UR / LDC / ULEA / UIMAD chains: 45 of the ~100 instructions a round spent before its first tap).
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "f"(0.f) : "memory");
}
// Returns the row-skip mask of the round (RegCfg::ROWSKIP): bit 3 p + 1 set <=> the FIRST position group of point
// p is all zero (else the last one is); 0 for configurations without row skipping.
template <typename Cfg, int LC, bool SCALE_Z = false>
__device__ __forceinline__ unsigned stage_windows(const Geom& g, uint32_t pts_sh, int base, int npts, uint32_t wbase,
                                                  int lane, bool pow2) {
    constexpr int kQuads = Cfg::WIN_FLOATS / 4;
#pragma unroll
    for (int k = 0; k < (kQuads + 31) / 32; ++k) {
        const int qd = lane + 32 * k;
        if (qd < kQuads) sts_zero16(wbase + 16u * (uint32_t)qd);
    }
    __syncwarp();
    bool skip_first = false;
    const int pt = lane / 3, api = lane - pt * 3;  // API dim 0,1,2 <-> slot Z,Y,X
    if (pt < npts) {
        const int slot = 2 - api;
        const uint32_t rec = pts_sh + 16u * (uint32_t)(base + pt);
        const float p = lds_at(rec + 4u * (uint32_t)api);
        const int off = (int)(lds_u8(pts_sh + (uint32_t)(kRegMaxPts * 16 + base + pt)) >> (2 * slot)) & 3;
        constexpr int kXY = (2 * kRegGroup * Cfg::XYP + 3) / 4 * 4;
        const uint32_t dst = wbase + 4u * (uint32_t)((slot == 2 ? kXY + pt * Cfg::ZWP : (2 * pt + slot) * Cfg::XYP) + off);
        const float pm = p * (float)g.M;
        const float fl = floorf(pm);  // reference cell (spatial_window_operations.cu:50)
        float amp = g.inv_sqrt_b_pi;
        if (SCALE_Z && slot == 2) amp *= lds_at(rec + 12u);
        skip_first = slot == 1 && off > Cfg::ROW_FIRST_MAX;
        if (pow2) {
            // M is a power of two: p*M, its fractional part and frac + (m - l) are exact or correctly
            // rounded, i.e. identical to the reference's double evaluation (:84-86)
            // frac = pm - fl is exact; frac + (m - l) is ONE rounding of the exact value, i.e. bit-identical
            // to the reference's (float)((double)pos * 2N - shift - l)
            const float frac = pm - fl;
#if NFFT_WINDOW_RECUR
            window_taps_recur<LC>(g, frac, amp, [&](int l, float v) { sts_at(dst + 4u * l, v); });
#else
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = frac + (float)((LC - 2) / 2 - l);  // m - l, m = (L - 2) / 2
                sts_at(dst + 4u * l, window_exp(-(tt * tt) * g.inv_b) * amp);  // eval_phi, :24-28
            }
#endif
        } else {
            const double bd = (double)p * (double)g.M - (double)((int)fl - g.m);
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = (float)(bd - (double)l);
                sts_at(dst + 4u * l, window_exp(-(tt * tt) * g.inv_b) * amp);
            }
        }
    }
    __syncwarp();
    return Cfg::ROWSKIP ? __ballot_sync(0xffffffffu, skip_first) : 0u;
}

// The tap staging through generic pointers is the default: the variant above (opaque shared addresses, 22 fewer
// instructions per round) measured SLOWER on the B200 (c4: spread 4.08 vs 4.04 ms, gather 3.35 vs 3.29 ms,
// profiles/r02d_ab.txt) -- its volatile accesses pin the order of the loads and stores of a round.
#ifndef NFFT_REG_OLD_STAGE
#define NFFT_REG_OLD_STAGE 1
#endif
template <typename Cfg, int LC, bool SCALE_Z = false>
__device__ __forceinline__ unsigned stage_windows_generic(const Geom& g, const float4* s_pts, const unsigned char* s_off,
                                                          int base, int npts, float* win, int lane, bool pow2) {
    constexpr int kQuads = Cfg::WIN_FLOATS / 4;
#pragma unroll
    for (int k = 0; k < (kQuads + 31) / 32; ++k) {
        const int qd = lane + 32 * k;
        if (qd < kQuads) reinterpret_cast<float4*>(win)[qd] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    bool skip_first = false;
    const int pt = lane / 3, api = lane - pt * 3;  // API dim 0,1,2 <-> slot Z,Y,X
    if (pt < npts) {
        const int slot = 2 - api;
        const float p = reinterpret_cast<const float*>(s_pts + base + pt)[api];
        const int off = (s_off[base + pt] >> (2 * slot)) & 3;
        constexpr int kXY = (2 * kRegGroup * Cfg::XYP + 3) / 4 * 4;
        float* dst = slot == 2 ? win + kXY + pt * Cfg::ZWP + off : win + (2 * pt + slot) * Cfg::XYP + off;
        const float pm = p * (float)g.M;
        const float fl = floorf(pm);  // reference cell (spatial_window_operations.cu:50)
        float amp = g.inv_sqrt_b_pi;
        if (SCALE_Z && slot == 2) amp *= s_pts[base + pt].w;
        skip_first = slot == 1 && off > Cfg::ROW_FIRST_MAX;
        if (pow2) {
            // M is a power of two: p*M, its fractional part and frac + (m - l) are exact or correctly
            // rounded, i.e. identical to the reference's double evaluation (:84-86)
            // frac = pm - fl is exact; frac + (m - l) is ONE rounding of the exact value, i.e. bit-identical
            // to the reference's (float)((double)pos * 2N - shift - l)
            const float frac = pm - fl;
#if NFFT_WINDOW_RECUR
            window_taps_recur<LC>(g, frac, amp, [&](int l, float v) { dst[l] = v; });
#else
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = frac + (float)((LC - 2) / 2 - l);  // m - l, m = (L - 2) / 2
                dst[l] = window_exp(-(tt * tt) * g.inv_b) * amp;  // eval_phi, :24-28
            }
#endif
        } else {
            const double bd = (double)p * (double)g.M - (double)((int)fl - g.m);
#pragma unroll
            for (int l = 0; l < LC; ++l) {
                const float tt = (float)(bd - (double)l);
                dst[l] = window_exp(-(tt * tt) * g.inv_b) * amp;
            }
        }
    }
    __syncwarp();
    return Cfg::ROWSKIP ? __ballot_sync(0xffffffffu, skip_first) : 0u;
}

// ======================================================================================
// spread
// ======================================================================================
template <int LC, int SX, int SY, int SZ>
__device__ __forceinline__ void spread_reg_body(const Geom& g, const WindowArgs& a, const CUtensorMap* tmap,
                                                const TileCtx& t) {
    using Cfg = RegCfg<LC, SX, SY, SZ>;
    constexpr int WX = Cfg::WX, CPL = Cfg::CPL, ZP = Cfg::ZP, SP = SZ / 2;
    extern __shared__ __align__(128) float smem_reg[];
    NFFT_PHASE_MARK(ph0);
    NFFT_PHASE_BEGIN(0);

    const int nsx = (g.T[0] + SX - 1) / SX, nsy = (g.T[1] + SY - 1) / SY, nsz = (g.T[2] + SZ - 1) / SZ;
    const int nsc = nsx * nsy * nsz;
    float* tile = align_tile(smem_reg);
    // TMA flush: plane pairs leave for the grid as soon as every unit that adds into them has done so
    RegStatic& S = reg_static();
    int *s_expect = S.expect, *s_done = S.done, *s_tma = S.tma, *s_lock = S.lock, *s_order = S.order;
    int &s_next = S.next, &s_nunits = S.nunits;
    const int npairs = (g.P[2] + 1) / 2;
    if (threadIdx.x < 32) s_expect[threadIdx.x] = 0, s_done[threadIdx.x] = 0;
    // Only tiles that lie inside the grid in X and Y: a box with a NEGATIVE start coordinate raises "illegal
    // instruction" in cp.reduce.async.bulk.tensor on this driver (scripts/micro/tma_probe.cu test 2,
    // profiles/r02d_tma_probe.txt), so the 2 of 16 tile rows per dimension that cross the periodic boundary keep
    // the vector-reduction flush after the sweep.  The Z wrap is per plane and free.
    const bool tma_tile = a.use_tma && t.org[0] >= 0 && t.org[0] + g.P[0] <= g.M && t.org[1] >= 0 && t.org[1] + g.P[1] <= g.M;
    if (tma_tile && threadIdx.x == 0) {
        s_tma[0] = t.org[0];
        s_tma[1] = t.org[1];
        s_tma[2] = 0;  // (second box of a wrapped tile: unused, see above)
        s_tma[3] = 0;
        s_tma[4] = t.org[2];
        s_tma[5] = t.b * g.C + a.k0;
    }
    // tile | points (float4) | cell offsets (u8, right behind the points: one base address serves both in the
    // tap staging) | per-warp tap windows | supercell start[nsc + 2], cursor[nsc + 2]
    float4* s_pts = reinterpret_cast<float4*>(tile + g.tile_elems);
    unsigned char* s_off = reinterpret_cast<unsigned char*>(s_pts + kRegMaxPts);
    float* s_win = reinterpret_cast<float*>(s_off + kRegMaxPts);
    int* s_start = reinterpret_cast<int*>(s_win + kRegWarps * Cfg::WIN_FLOATS);
    int* s_cur = s_start + nsc + 2;
    if (threadIdx.x < 64) s_lock[threadIdx.x] = 0;

    for (int i = threadIdx.x; i < (g.tile_elems >> 2); i += kRegThreads)  // tile_elems is a multiple of 4
        reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < nsc; i += kRegThreads) s_cur[i] = 0;
    if (threadIdx.x == 0) s_next = 0, s_nunits = 0;
    __syncthreads();
    NFFT_PHASE_MARK(pha);
    const int cnt = (int)(t.p_hi - t.p_lo);
    bucket_points<SX, SY, SZ, true>(g, a, t, cnt, nsx, nsy, nsz, s_pts, s_off, s_start, s_cur);
    NFFT_PHASE_MARK(phb);
    make_units(s_start, nsx * nsy, nsz, s_order, &s_nunits, tma_tile ? s_expect : nullptr, SP, ZP, npairs,
               SX == 2 ? kRegWarps : NFFT_REG_UNITS);
    NFFT_PHASE_MARK(ph1);
    NFFT_PHASE_ADD(0, 5, ph0, pha);
    NFFT_PHASE_ADD(0, 6, pha, phb);
    NFFT_PHASE_ADD(0, 7, phb, ph1);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* win = s_win + warp * Cfg::WIN_FLOATS;
    constexpr int kXY = (2 * kRegGroup * Cfg::XYP + 3) / 4 * 4;  // z windows start here
    const int padx = g.org[0] - g.m;
    const bool pow2 = (g.M & (g.M - 1)) == 0;
    // (x, y) positions of the register block owned by this lane: c = lane + 32 q -> (c % WX, c / WX);
    // wi / wj index the x / y tap windows (invalid positions read the always-zero entry WP-1)
    int wi[CPL], wj[CPL], coff[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        const bool ok = c < Cfg::COLS;
        wi[q] = ok ? c % WX : Cfg::XYP - 1;               // x window row (slot 0) of the point
        wj[q] = ok ? Cfg::XYP + c / WX : 2 * Cfg::XYP - 1;  // y window row (slot 1)
        coff[q] = ok ? (c / WX) * g.sY + (c % WX) : 0;
    }
    uint32_t wbase = (uint32_t)__cvta_generic_to_shared(win);
    asm volatile("" : "+r"(wbase));
    uint32_t pts_sh = (uint32_t)__cvta_generic_to_shared(s_pts);
    asm volatile("" : "+r"(pts_sh));
    uint32_t awi[CPL], awj[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) awi[q] = wbase + 4u * wi[q], awj[q] = wbase + 4u * wj[q];

    // work units (columns of supercells or z-ranges of heavy columns) are handed out dynamically
    const int nunits = s_nunits;
#if NFFT_REG_STAGGER_NS > 0
    if (warp > 0) __nanosleep(warp * NFFT_REG_STAGGER_NS);
#endif
    for (;;) {
        int col = 0;
        if (lane == 0) col = atomicAdd(&s_next, 1);
        col = __shfl_sync(0xffffffffu, col, 0);
        if (col >= nunits) break;
        const int unit = s_order[col];
        col = unit & 0xff;
        const int lo_seg = (unit >> 8) & 0xfff, hi_seg = (int)((unsigned)unit >> 20);  // never empty
        const int c0 = col * nsz;
        const int scx = col % nsx, scy = col / nsx;
        float* cbase = tile + (scy * SY) * g.sY + scx * SX + padx;

        float2 acc[CPL][ZP];  // acc[q][kp] = block cells (k = 2 kp, 2 kp + 1) at position q
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
            for (int kp = 0; kp < ZP; ++kp) acc[q][kp] = make_float2(0.f, 0.f);
        // shared-window byte addresses of this lane's positions in plane 0 of the unit's column: the
        // add-out then costs one IADD per tile access instead of rebuilding every address from
        // (lane, coff, scz, sZ) inside the critical section
        uint32_t aq[CPL];
        {
            uint32_t cb = (uint32_t)__cvta_generic_to_shared(cbase);
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                aq[q] = cb + 4u * (uint32_t)coff[q];
                asm volatile("" : "+r"(aq[q]));
            }
        }
        const uint32_t sz4 = 4u * (uint32_t)g.sZ;

        // Planes 0 .. SZ-1 of the block at supercell `scz` are complete once its points are done:
        // add them out and slide the block up by SZ (all planes when `last`).  Other warps' blocks
        // overlap this one in x/y, so each pair of tile planes is guarded by a shared-memory lock;
        // inside it the update is a plain pipelined LDS / FADD / STS (a shared-memory float
        // atomicAdd is a CAS loop per element on sm_100a).
        auto advance = [&](int scz, bool last) {
            const int zlim = g.P[2] - scz * SZ;  // planes of the padded tile above the block origin
#pragma unroll
            for (int kp = 0; kp < ZP; ++kp) {
                if ((kp < SP || last) && 2 * kp < zlim) {
                    const bool two = 2 * kp + 1 < zlim;
                    int* lk = s_lock + scz * SP + kp;
                    if (lane == 0) {
                        lock_acquire(lk);
                    }
                    __syncwarp();
                    float2 cur[CPL];
                    const uint32_t po = (uint32_t)(scz * SZ + 2 * kp) * sz4;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const bool ok = lane + 32 * q < Cfg::COLS;
                        cur[q] = make_float2(0.f, 0.f);
                        if (ok) {
                            cur[q].x = lds_at(aq[q] + po);
                            if (two) cur[q].y = lds_at(aq[q] + po + sz4);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        if (lane + 32 * q < Cfg::COLS) {
                            sts_at(aq[q] + po, cur[q].x + acc[q][kp].x);
                            if (two) sts_at(aq[q] + po + sz4, cur[q].y + acc[q][kp].y);
                        }
                    }
                    release_fence();
                    __syncwarp();
                    if (lane == 0) {
                        atomicExch(lk, 0);
                        // the unit that completes a plane pair hands it to the TMA unit (no other warp will touch
                        // it again); the sweep goes on while the reduction drains
                        const int pr = scz * SP + kp;
                        if (tma_tile && atomicAdd(&s_done[pr], 1) + 1 == s_expect[pr])
                            tma_flush_plane_pair(tmap, (uint32_t)__cvta_generic_to_shared(tile), s_tma, pr, g.P[2], g.sZ,
                                                 g.M);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
#pragma unroll
                for (int kp = 0; kp < ZP; ++kp) acc[q][kp] = kp + SP < ZP ? acc[q][kp + SP] : make_float2(0.f, 0.f);
            }
        };

        // Within the unit the points are staged in rounds of kRegGroup regardless of supercell
        // boundaries.
        {
            int scz = 0, next_end = s_start[c0 + 1];
            while (next_end <= lo_seg) {  // the unit's first point lies in supercell scz: the block is still zero
                ++scz;
                next_end = s_start[c0 + scz + 1];
            }
            for (int base = lo_seg; base < hi_seg; base += kRegGroup) {
                const int npts = hi_seg - base < kRegGroup ? hi_seg - base : kRegGroup;
#if NFFT_REG_OLD_STAGE
                const unsigned skipmask = stage_windows_generic<Cfg, LC, true>(g, s_pts, s_off, base, npts, win, lane, pow2);
#else
                const unsigned skipmask = stage_windows<Cfg, LC, true>(g, pts_sh, base, npts, wbase, lane, pow2);
#endif
                (void)skipmask;
                // One copy of the point body per slot of the round (window loads with immediate offsets),
                // entered through a switch; a slot that ends a supercell leaves the switch so that the ONE
                // copy of the add-out code above it runs, and the switch is re-entered at the next slot.
                // The z window carries the point's value (stage_windows<.., true>).
                static_assert(kRegGroup == 8, "slotted sweep: 8 slots");
                for (int gp = 0; gp < npts;) {
                    while (base + gp >= next_end) {  // warp-uniform: the sweep reaches the next supercell
                        advance(scz, false);
                        ++scz;
                        next_end = s_start[c0 + scz + 1];
                    }
                    const int stop = npts < next_end - base ? npts : next_end - base;  // slots gp .. stop-1: supercell scz
                    // groups [Q0, Q1) of the point in slot K; REV: descending (keeps ptxas from merging the two
                    // row-skip variants back into one body with a predicated group)
                    auto body = [&](auto kc, auto q0c, auto q1c, auto revc) {
                        constexpr int K = decltype(kc)::value, Q0 = decltype(q0c)::value, Q1 = decltype(q1c)::value;
                        constexpr bool REV = decltype(revc)::value != 0;
                        float w0[CPL], w1[CPL];
#pragma unroll
                        for (int i = Q0; i < Q1; ++i) {
                            const int q = REV ? Q1 - 1 - (i - Q0) : i;
                            w1[q] = lds_f32<K * 2 * Cfg::XYP * 4>(awj[q]);
                            w0[q] = lds_f32<K * 2 * Cfg::XYP * 4>(awi[q]);
                        }
                        float2 wz[ZP];
                        lds_window<(kXY + K * Cfg::ZWP) * 4, Cfg::ZQ>(wbase, wz);
#pragma unroll
                        for (int i = Q0; i < Q1; ++i) {
                            const int q = REV ? Q1 - 1 - (i - Q0) : i;
                            const float v = w1[q] * w0[q];
                            const float2 vv = make_float2(v, v);
#pragma unroll
                            for (int kp = 0; kp < ZP; ++kp) acc[q][kp] = ffma2(vv, wz[kp], acc[q][kp]);
                        }
                    };
                    auto point = [&](auto kc) {
                        constexpr int K = decltype(kc)::value;
                        if constexpr (Cfg::ROWSKIP && NFFT_REG_ROWSKIP_SPREAD) {
                            // warp-uniform (a ballot): set = the first group of this point is all zero, else the last
                            if (skipmask & (1u << (3 * K + 1)))
                                body(kc, IntC<1>{}, IntC<CPL>{}, IntC<1>{});
                            else
                                body(kc, IntC<0>{}, IntC<CPL - 1>{}, IntC<0>{});
                        } else {
                            body(kc, IntC<0>{}, IntC<CPL>{}, IntC<0>{});
                        }
                    };
#define NFFT_SLOT(K)                                                                       \
                    case K:                                                                \
                        point(IntC<K>{});                                                \
                        gp = K + 1;                                                        \
                        if (K + 1 >= stop) break;
                    switch (gp) {
                        NFFT_SLOT(0) NFFT_SLOT(1) NFFT_SLOT(2) NFFT_SLOT(3)
                        NFFT_SLOT(4) NFFT_SLOT(5) NFFT_SLOT(6) NFFT_SLOT(7)
                        default: break;
                    }
#undef NFFT_SLOT
                }
                __syncwarp();
            }
            // the unit's last point lies in supercell scz: everything the block holds goes out
            advance(scz, true);
        }
    }
    NFFT_PHASE_WARP(0, ph1);
    NFFT_PHASE_MARK(ph2);
    if (tma_tile) {
        // every plane pair has been handed to the TMA unit by the unit that completed it; the shared memory
        // must stay allocated until the reductions this lane issued have read it
        if (lane == 0) bulk_wait_read_all();
        return;
    }
    __syncthreads();
    NFFT_PHASE_MARK(ph3);

    // flush: vector reductions into the global grid; untouched (== 0) quads are skipped
    for_each_quad3_rows(g, t, [&](int so, long long cell) {
        const float* s = tile + so;
        const float4 val = make_float4(s[0], s[1], s[2], s[3]);
        if (val.x != 0.f || val.y != 0.f || val.z != 0.f || val.w != 0.f) reduce_quad(g, a.grid, t.b, a.k0, cell, val);
    });
    NFFT_PHASE_MARK(ph4);
    NFFT_PHASE_ADD(0, 0, ph0, ph1);
    NFFT_PHASE_ADD(0, 1, ph1, ph2);
    NFFT_PHASE_ADD(0, 2, ph2, ph3);
    NFFT_PHASE_ADD(0, 3, ph3, ph4);
    NFFT_PHASE_ADD(0, 4, 0, 1);
    NFFT_PHASE_END(0, ph0, ph4, cnt);
}

template <int LC, int SX, int SY, int SZ>
__global__ void __launch_bounds__(kRegThreads, 2)
spread_reg_kernel(const Geom g, const WindowArgs a, const __grid_constant__ CUtensorMap tmap) {
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    spread_reg_body<LC, SX, SY, SZ>(g, a, &tmap, t);
}
// Geom::mixed: one launch, the sweep is chosen per work item -- heavy tiles (class 1, marked by the binning when
// the point set was found clustered) with 2 x 2 x 2 supercells, the others with the default supercell.
template <int LC>
__global__ void __launch_bounds__(kRegThreads, 2)
spread_reg_mixed_kernel(const Geom g, const WindowArgs a, const __grid_constant__ CUtensorMap tmap) {
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    if (t.cls) spread_reg_body<LC, 2, 2, 2>(g, a, &tmap, t);
    else spread_reg_body<LC, kRegSX, kRegSY, kRegSZ>(g, a, &tmap, t);
}

// ======================================================================================
// gather
// ======================================================================================
template <int LC, int SX, int SY, int SZ>
__device__ __forceinline__ void gather_reg_body(const Geom& g, const WindowArgs& a, const CUtensorMap* tmap,
                                                const TileCtx& t) {
    using Cfg = RegCfg<LC, SX, SY, SZ>;
    constexpr int WX = Cfg::WX, WZ = Cfg::WZ, CPL = Cfg::CPL, ZP = Cfg::ZP, SP = SZ / 2;
    extern __shared__ __align__(128) float smem_reg[];
    NFFT_PHASE_MARK(ph0);
    NFFT_PHASE_BEGIN(1);

    const int nsx = (g.T[0] + SX - 1) / SX, nsy = (g.T[1] + SY - 1) / SY, nsz = (g.T[2] + SZ - 1) / SZ;
    const int nsc = nsx * nsy * nsz;
    float* tile = align_tile(smem_reg);
    RegStatic& S = reg_static();
    unsigned long long& s_mbar = S.mbar;
    int* s_order = S.order;
    int &s_next = S.next, &s_nunits = S.nunits;
    // planes by TMA unless the tile crosses the periodic boundary in X or Y (a load zero-fills out-of-range
    // elements, which would overwrite the wrapped half; the Z wrap is per plane)
    const bool tma_tile = a.use_tma && t.org[0] >= 0 && t.org[0] + g.P[0] <= g.M && t.org[1] >= 0 && t.org[1] + g.P[1] <= g.M;
    // tile | points (float4) | cell offsets (u8, right behind the points: one base address serves both in the
    // tap staging) | per-warp tap windows | supercell start[nsc + 2], cursor[nsc + 2]
    float4* s_pts = reinterpret_cast<float4*>(tile + g.tile_elems);
    unsigned char* s_off = reinterpret_cast<unsigned char*>(s_pts + kRegMaxPts);
    float* s_win = reinterpret_cast<float*>(s_off + kRegMaxPts);
    int* s_start = reinterpret_cast<int*>(s_win + kRegWarps * Cfg::WIN_FLOATS);
    int* s_cur = s_start + nsc + 2;

    for (int i = threadIdx.x; i < nsc; i += kRegThreads) s_cur[i] = 0;
    if (threadIdx.x == 0) s_next = 0, s_nunits = 0;
    // stage the padded tile (periodic wrap resolved per quad)
    // the tile travels with asynchronous copies while the points are loaded and bucketed (both phases
    // are latency-bound); complex grids: this pass's component of the interleaved pairs
    if (tma_tile) {
        if (threadIdx.x == 0) {
            const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
            const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, (uint32_t)(g.P[2] * g.P[1] * g.P[0] * 4));
            const int plane = t.b * g.C + a.k0;
            for (int zz = 0; zz < g.P[2]; ++zz) {
                int gz = t.org[2] + zz;
                gz = gz < 0 ? gz + g.M : (gz >= g.M ? gz - g.M : gz);
                tma_load_plane(tile_s + 4u * (uint32_t)(zz * g.sZ), tmap, mbar, t.org[0], t.org[1], gz, plane);
            }
        }
    } else {
        const int cs = g.cplx ? 2 : 1;
        const float* gsrc = a.grid + grid_plane(g, t.b, a.k0) + (g.cplx ? (a.k0 & 1) : 0);
        const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
        for_each_quad3_rows(g, t, [&](int so, long long cell) {
            const float* src = gsrc + cs * cell;
            const uint32_t dst = tile_s + 4u * (uint32_t)so;
            if (cs == 1 && ((g.sY | g.sZ) & 3) == 0) {  // real grid, strides multiples of 4: 16-byte aligned quads
                cp_async16(dst, src);
            } else {
                cp_async4(dst, src);
                cp_async4(dst + 4, src + cs);
                cp_async4(dst + 8, src + 2 * cs);
                cp_async4(dst + 12, src + 3 * cs);
            }
        });
    }
    __syncthreads();
    NFFT_PHASE_MARK(pha);
    const int cnt = (int)(t.p_hi - t.p_lo);
    bucket_points<SX, SY, SZ, false>(g, a, t, cnt, nsx, nsy, nsz, s_pts, s_off, s_start, s_cur);
    NFFT_PHASE_MARK(phb);
    if (tma_tile) {
        // every thread observes the completion itself (the mbarrier makes the TMA writes visible to its waiters)
        if (!mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar), 0) && a.flags) atomicAdd(a.flags + 1, 1u);
    } else {
        cp_async_wait_all();  // make_units' barriers publish the tile to the other threads
    }
    make_units(s_start, nsx * nsy, nsz, s_order, &s_nunits, nullptr, 0, 0, 0, SX == 2 ? kRegWarps : NFFT_REG_UNITS);
    NFFT_PHASE_MARK(ph1);
    NFFT_PHASE_ADD(1, 5, ph0, pha);
    NFFT_PHASE_ADD(1, 6, pha, phb);
    NFFT_PHASE_ADD(1, 7, phb, ph1);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* win = s_win + warp * Cfg::WIN_FLOATS;
    constexpr int kXY = (2 * kRegGroup * Cfg::XYP + 3) / 4 * 4;  // z windows start here
    const int padx = g.org[0] - g.m;
    const bool pow2 = (g.M & (g.M - 1)) == 0;
    int wi[CPL], wj[CPL], coff[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        const bool ok = c < Cfg::COLS;
        wi[q] = ok ? c % WX : Cfg::XYP - 1;               // x window row (slot 0) of the point
        wj[q] = ok ? Cfg::XYP + c / WX : 2 * Cfg::XYP - 1;  // y window row (slot 1)
        coff[q] = ok ? (c / WX) * g.sY + (c % WX) : 0;
    }
    uint32_t wbase = (uint32_t)__cvta_generic_to_shared(win);
    asm volatile("" : "+r"(wbase));
    uint32_t pts_sh = (uint32_t)__cvta_generic_to_shared(s_pts);
    asm volatile("" : "+r"(pts_sh));
    uint32_t awi[CPL], awj[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) awi[q] = wbase + 4u * wi[q], awj[q] = wbase + 4u * wj[q];
    // planes above the padded tile are never weighted (their taps are zero) but must stay in bounds
    const int zmax = g.P[2] - 1;

    const int nunits = s_nunits;
    for (;;) {
        int col = 0;
        if (lane == 0) col = atomicAdd(&s_next, 1);
        col = __shfl_sync(0xffffffffu, col, 0);
        if (col >= nunits) break;
        const int unit = s_order[col];
        col = unit & 0xff;
        const int lo_col = (unit >> 8) & 0xfff, hi_col = (int)((unsigned)unit >> 20);  // never empty
        const int c0 = col * nsz;
        const int scx = col % nsx, scy = col / nsx;
        const float* cbase = tile + (scy * SY) * g.sY + scx * SX + padx;
        int scz = 0, next_end = s_start[c0 + 1];
        while (next_end <= lo_col) {  // the unit's first point lies in supercell scz
            ++scz;
            next_end = s_start[c0 + scz + 1];
        }

        // register block: planes [scz*SZ, scz*SZ + 2 ZP) of the column; loaded at the unit's first
        // populated supercell, then slid
        float2 blk[CPL][ZP];
        // shared-window byte addresses of this lane's positions in plane 0 of the unit's column (see the
        // spread kernel): a block load is one IADD + LDS per cell
        uint32_t aq[CPL];
        {
            const uint32_t cb = (uint32_t)__cvta_generic_to_shared(cbase);
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                aq[q] = cb + 4u * (uint32_t)coff[q];
                asm volatile("" : "+r"(aq[q]));
            }
        }
        const uint32_t sz4 = 4u * (uint32_t)g.sZ;
        auto load_pair = [&](int kp, int scz) {  // planes scz*SZ + 2 kp, + 1 of every position (clamped)
            const int z0 = scz * SZ + 2 * kp;
            const uint32_t oa = (uint32_t)(z0 < zmax ? z0 : zmax) * sz4, ob = (uint32_t)(z0 + 1 < zmax ? z0 + 1 : zmax) * sz4;
#pragma unroll
            for (int q = 0; q < CPL; ++q) blk[q][kp] = make_float2(lds_at(aq[q] + oa), lds_at(aq[q] + ob));
        };
#pragma unroll
        for (int kp = 0; kp < ZP; ++kp) load_pair(kp, scz);
        auto advance = [&](int scz) {
#pragma unroll
            for (int kp = 0; kp + SP < ZP; ++kp)
#pragma unroll
                for (int q = 0; q < CPL; ++q) blk[q][kp] = blk[q][kp + SP];
#pragma unroll
            for (int kp = ZP - SP; kp < ZP; ++kp) load_pair(kp, scz);
        };

        for (int base = lo_col; base < hi_col; base += kRegGroup) {
            const int npts = hi_col - base < kRegGroup ? hi_col - base : kRegGroup;
#if NFFT_REG_OLD_STAGE
            const unsigned skipmask = stage_windows_generic<Cfg, LC>(g, s_pts, s_off, base, npts, win, lane, pow2);
#else
            const unsigned skipmask = stage_windows<Cfg, LC>(g, pts_sh, base, npts, wbase, lane, pow2);
#endif
            // As in the spread: one copy of the point body per slot (window loads with immediate offsets, the
            // partial sum in a static register), entered through a switch; a slot that ends a supercell
            // leaves the switch so that the ONE copy of the block slide runs.  (An unrolled loop carries a
            // copy of the slide in every slot: 8 x 140 instructions, instruction-cache misses.)
            static_assert(kGatherSlots == 8 && kRegGroup == 8, "slotted gather: 8 slots");
            float part[kGatherSlots];
#pragma unroll
            for (int sl = 0; sl < kGatherSlots; ++sl) part[sl] = 0.f;
            for (int gp = 0; gp < npts;) {
                while (base + gp >= next_end) {  // warp-uniform: the sweep reaches the next supercell
                    ++scz;
                    advance(scz);
                    next_end = s_start[c0 + scz + 1];
                }
                const int stop = npts < next_end - base ? npts : next_end - base;  // slots gp .. stop-1: supercell scz
                // groups [Q0, Q1) of the point in slot K (see the spread kernel: row skipping)
                auto gbody = [&](auto kc, auto q0c, auto q1c, auto revc) {
                    constexpr int K = decltype(kc)::value, Q0 = decltype(q0c)::value, Q1 = decltype(q1c)::value;
                    constexpr bool REV = decltype(revc)::value != 0;
                    float w0[CPL], w1[CPL];
#pragma unroll
                    for (int i = Q0; i < Q1; ++i) {
                        const int q = REV ? Q1 - 1 - (i - Q0) : i;
                        w1[q] = lds_f32<K * 2 * Cfg::XYP * 4>(awj[q]);
                        w0[q] = lds_f32<K * 2 * Cfg::XYP * 4>(awi[q]);
                    }
                    float2 wz[ZP];
                    lds_window<(kXY + K * Cfg::ZWP) * 4, Cfg::ZQ>(wbase, wz);
                    float2 zsum[ZP];
#pragma unroll
                    for (int kp = 0; kp < ZP; ++kp) zsum[kp] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = Q0; i < Q1; ++i) {
                        const int q = REV ? Q1 - 1 - (i - Q0) : i;
                        const float w = w1[q] * w0[q];  // psi(Y) * psi(X)
                        const float2 ww = make_float2(w, w);
#pragma unroll
                        for (int kp = 0; kp < ZP; ++kp) zsum[kp] = ffma2(ww, blk[q][kp], zsum[kp]);
                    }
                    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
                    for (int kp = 0; kp < ZP; ++kp) sum = ffma2(wz[kp], zsum[kp], sum);
                    return sum.x + sum.y;
                };
                auto gpoint = [&](auto kc) {
                    constexpr int K = decltype(kc)::value;
                    if constexpr (Cfg::ROWSKIP) {
                        if (skipmask & (1u << (3 * K + 1))) return gbody(kc, IntC<1>{}, IntC<CPL>{}, IntC<1>{});
                        return gbody(kc, IntC<0>{}, IntC<CPL - 1>{}, IntC<0>{});
                    } else {
                        return gbody(kc, IntC<0>{}, IntC<CPL>{}, IntC<0>{});
                    }
                };
#define NFFT_GSLOT(K)                                                                      \
                case K:                                                                    \
                    part[K] = gpoint(IntC<K>{});                                         \
                    gp = K + 1;                                                            \
                    if (K + 1 >= stop) break;
                switch (gp) {
                    NFFT_GSLOT(0) NFFT_GSLOT(1) NFFT_GSLOT(2) NFFT_GSLOT(3)
                    NFFT_GSLOT(4) NFFT_GSLOT(5) NFFT_GSLOT(6) NFFT_GSLOT(7)
                    default: break;
                }
#undef NFFT_GSLOT
            }
            // sum the partials over the warp (transpose reduction: 9 shuffles for 8 values);
            // lane (32 / kGatherSlots) p then holds the value of point p of the round
            warp_reduce_channels<kGatherSlots>(part, lane);
            constexpr int kLanesPerPoint = 32 / kGatherSlots;
            if ((lane & (kLanesPerPoint - 1)) == 0 && lane / kLanesPerPoint < npts) {
                const int gp = lane / kLanesPerPoint;
                const uint32_t i = (uint32_t)__float_as_int(s_pts[base + gp].w);
                a.yout[(size_t)i * g.K + a.k0] = part[0];
            }
            __syncwarp();
        }
    }
    (void)WZ;
    NFFT_PHASE_WARP(1, ph1);
    NFFT_PHASE_MARK(ph2);
#ifdef NFFT_PHASE_TIMING
    __syncthreads();
#endif
    NFFT_PHASE_MARK(ph3);
    NFFT_PHASE_ADD(1, 0, ph0, ph1);
    NFFT_PHASE_ADD(1, 1, ph1, ph2);
    NFFT_PHASE_ADD(1, 2, ph2, ph3);
    NFFT_PHASE_ADD(1, 4, 0, 1);
    NFFT_PHASE_END(1, ph0, ph3, cnt);
}

template <int LC, int SX, int SY, int SZ>
__global__ void __launch_bounds__(kRegThreads, 2)
gather_reg_kernel(const Geom g, const WindowArgs a, const __grid_constant__ CUtensorMap tmap) {
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    gather_reg_body<LC, SX, SY, SZ>(g, a, &tmap, t);
}
template <int LC>
__global__ void __launch_bounds__(kRegThreads, 2)
gather_reg_mixed_kernel(const Geom g, const WindowArgs a, const __grid_constant__ CUtensorMap tmap) {  // see the spread
    TileCtx t;
    if (!decode_item(g, a, t)) return;
    if (t.cls) gather_reg_body<LC, 2, 2, 2>(g, a, &tmap, t);
    else gather_reg_body<LC, kRegSX, kRegSY, kRegSZ>(g, a, &tmap, t);
}

}  // namespace nfftb200
