// mel -> linear lift on the 5th-generation tensor cores:  S = max(0, pinv(basis) . delog(mel)) ^ power
// (step 1 of griffin_lim::GriffinLim::infer; reference call site
// /root/reference src/lib.rs:141, parameters from src/tacotron2/mod.rs:453-456).
//
// The product is a [bins x n_mels] . [n_mels x frames] GEMM with a tiny inner dimension (80): 82 kFLOP and
// 2.4 kB per frame, i.e. HBM-bound -- IF the multiply-adds are off the CUDA cores (on the fp32 pipe the 1.3 G
// FMAs of a 32 x 1000 batch alone take 35 us at 100% issue; the fp64 kernel of round 1 took 220-316 us) AND the
// de-logged mel tile is produced once per frame, not once per bin tile.  The second condition is a shared-memory
// question: the pseudo-inverse P must stay resident next to the ring of mel tiles.
//
// Precision: P has 36% negative entries and the gate is 1e-5 of full scale, so one low-precision
// pass is not enough (tf32: 9e-4).  Both operands are split in two fp16 halves, x = hi + lo with hi = fp16(x),
// lo = fp16(x - hi): 22 significant bits, and D = Phi Ehi + Phi Elo + Plo Ehi in three passes of
// tcgen05.mma kind::f16 into one fp32 accumulator -- 3e-7 of full scale, the same as the tf32 split this kernel
// used before, at HALF the shared-memory bytes per element (4 instead of 8).  That is what lets ONE CTA hold all
// four 128-bin tiles of P at n_fft 1024 (160 KB) beside a two-stage ring of 64-frame mel tiles (40 KB): before,
// each of the four bin-tile CTAs of a frame tile repeated the exp / split / store of the same mel tile and the
// kernel was issue-bound on exactly that work (45 us; now one conversion per frame).
// fp16 has 5 exponent bits, so both operands are scaled by exact powers of two: P once on the host (largest entry
// into [4, 8)), every frame of E by its own 2^-ceil(log2 max_m E[m]) (a frame is a column of the GEMM, so a per-frame
// factor commutes with it).  The accumulator is then O(1) -- its log2 keeps full precision -- and ^power turns the two
// exponents into an addend of the epilogue's 2^(power * log2 x): one FFMA.
//
// One tile of the GEMM is D[bin 0..127][frame 0..63] = sum_m P[bin][m] E[frame][m] (UMMA M = 128 bins, N = 64
// frames, K = 16 per instruction), both operands K-major in the un-swizzled core-matrix layout (8 rows x 16 bytes
// contiguous; K-direction stride LBO, row-group stride SBO = 128 B): any n_mels that is a multiple of 16 is a
// whole number of instructions (80 = 5), no padding to a swizzle atom, and a producer thread's 8 consecutive mel
// rows of one frame are one conflict-free 16-byte shared-memory store.
// Bins as the TMEM lane dimension make the epilogue's stores coalesced: for one frame (TMEM column) the 32 lanes
// of a warp hold 32 consecutive bins = 128 contiguous bytes of the frame-major state record the iteration kernel reads.
//
// Warp roles (864 threads, one CTA per SM; CTA (part, g) owns bin tiles part*nbt .. +nbt-1 and frame tiles g, g+G, ...):
//     warps 0-15   epilogue     tcgen05.ld (lane quarter = warp & 3) -> clamp -> ^power -> S
//     warps 16-25  producers    mel [n_mels][T] -> per-frame scale -> delog -> fp16 hi / lo -> E tile (loads one tile ahead)
//     warp  26     TMEM alloc + one thread issuing the bulk copy of P, the MMAs and the commits
// The Nyquist bin (bin M, a 129th row of the last tile otherwise) is summed by the producers of part 0 on the
// CUDA cores, from the de-logged values they hold anyway.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gl_host.h"

namespace xdtts {

namespace {

// timing experiments only (tools/build_lift_variants.py): bit 0 skips the MMAs, bit 1 the epilogue's math and stores,
// bit 2 the producers' conversion; bits 3 / 4 / 5 only the epilogue's stores / special-function math / TMEM loads.  0 in the product.
#ifndef XDTTS_LIFT_SKIP
#define XDTTS_LIFT_SKIP 0
#endif

#ifdef XDTTS_LIFT_TRACE
__device__ __forceinline__ long long lt_gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LT_GSTAMP(role, k) do { p.trace[(((size_t)blockIdx.x * 4 + (role)) * 32 + 31) * 2 + (k)] = lt_gtime(); } while (0)
#define LT_STAMP(role, it_, k) do { if ((it_) < 31) p.trace[(((size_t)blockIdx.x * 4 + (role)) * 32 + (it_)) * 2 + (k)] = clock64(); } while (0)
#else
#define LT_STAMP(role, it_, k) do { } while (0)
#define LT_GSTAMP(role, k) do { } while (0)
#endif

constexpr int LT_BM = 128;       // bins per tile
constexpr int LT_BN = 64;        // frames per tile
constexpr int LT_STAGES = 2;     // E-tile ring and TMEM accumulator sets
constexpr int LT_MAX_BT = 4;     // bin tiles resident per CTA (TMEM: LT_STAGES * 4 * 64 = 512 columns)
constexpr int LT_TILE_SM = 64;   // tile records staged in shared memory per CTA
constexpr int LT_MAX_KC = 16;    // K chunks of 8 mel rows: n_mels <= 128; wider bases take gl_lift_f32_kernel
constexpr int LT_EPI_WARPS = 16;
constexpr int LT_PRO_WARPS = 10;
constexpr int LT_PGROUPS = LT_PRO_WARPS * 32 / LT_BN;   // producer thread (frame f, chunk group cg): chunks cg, cg + 5, ...
constexpr int LT_THREADS = 32 * (LT_EPI_WARPS + LT_PRO_WARPS + 1);
constexpr int LT_CF_RING = 2 * LT_STAGES;   // per-frame exponents: written up to 2 stages ahead of the epilogue that reads them
constexpr int LT_E_SHIFT = 0;    // a frame's largest de-logged value lands in (1/2, 1]
constexpr int LT_SMEM_MAX = 232448;   // 227 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"   // %2: suspend-time hint (ns): sleep instead of spinning
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(100000u)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate.  The two shared-memory descriptors share their high word.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory descriptors (K-major operand in the un-swizzled layout): a core matrix is 8 rows x 16 bytes, contiguous
// (128 B); the two core matrices one instruction reads along K are LBO bytes apart, consecutive 8-row groups SBO bytes
// apart.  Low word: address >> 4 (bits 0-13), LBO >> 4 (bits 16-29); high word: SBO >> 4 (bits 0-13), version 1 (bit 14),
// layout type 0 = no swizzle (bits 29-31).

// The de-log as 2^t: t = v * c with the rounding of the product carried into a first-order correction --
// t = fl(v c_hi), r = (v c_hi - t) + v c_lo exactly, e = 2^t (1 + r ln 2).  ~1.5 ulp (ex2.approx is 2 ulp), seven
// instructions with the frame's power-of-two scale against ~15 for expf().
template <int DELOG>
struct Delog {
    static constexpr float c_hi = DELOG == 0 ? 1.4426950216293335f : 3.3219280242919922f;       // log2(e), log2(10) rounded to fp32
    static constexpr float c_lo = DELOG == 0 ? 1.9259629911266175e-8f : 7.0595369550985533e-8f;  // ... and what the rounding dropped
};
template <int DELOG>
__device__ __forceinline__ float delog_scaled(float v, float sc) {   // delog(v) * sc, sc a power of two (exact); -inf (no sample) -> 0
    if (DELOG == 2) return v == -INFINITY ? 0.f : v * sc;
    const float t = v * Delog<DELOG>::c_hi;
    float r = fmaf(v, Delog<DELOG>::c_hi, -t);
    r = fmaf(v, Delog<DELOG>::c_lo, r);
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(t));
    p *= sc;   // not 2^(t - shift): that subtraction rounds the exponent at 2^-21 for the very values that matter
    return v == -INFINITY ? 0.f : fmaf(p, r * 0.69314718055994531f, p);
}
template <int DELOG>
__device__ __forceinline__ float delog_value(float v) { return delog_scaled<DELOG>(v, 1.f); }

// s ^ power for s > 0 on the special-function unit: 2^(power * log2 s), two MUFU operations and a multiply.  ~1e-6
// relative at full scale (the gate on S is 1e-5 of full scale, tests/test_gpu_gl.py::test_lift_matches_oracle); powf
// costs ~40 instructions per value -- more issue time than the rest of the kernel.
__device__ __forceinline__ float pow_pos(float s, float power) {
    float l, r;   // max(s, 0) -> log2 = -inf at 0 -> 2^-inf = 0: the clamp needs no select
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 0.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * power));
    return r;
}
// (s 2^e) ^ power with the addend a = power * e supplied: 2^(power * log2 s + a), the factor rides in the FFMA
__device__ __forceinline__ float pow_pos_add(float s, float power, float a) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 0.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(l, power, a)));
    return r;
}
__device__ __forceinline__ float exp2_int(int e) {   // 2^e, e clamped to the normal range
    return __int_as_float(min(max(e + 127, 1), 254) << 23);
}

struct LiftParams {
    const float* mel_arena;      // utterance u: row-major [n_mels][T_u] at float offset foff[u] * n_mels
    const uint8_t* a_image;      // [n_mt][2 planes][kchunks][128 rows][8 fp16]: the scaled pseudo-inverse, hi / lo halves
    const float* pinv_nyq;       // the pseudo-inverse's row of bin M: entry m at pinv_nyq[m * pinv_ld]
    int pinv_ld;
    const int4* tiles;           // frame tiles: (first frame row of the utterance, its frame count T, first frame of the tile, frames in the tile <= 64)
    float* S;                    // frame-major state records: S[(foff + t) * ld + k], k < M
    int n_tiles, n_mt, nbt, n_part, groups, n_mels, kchunks, ld;
    int mma_n;                   // UMMA N: the largest tile extent of the plan, rounded up to 16
    float power;
    float p_exp;                 // the image holds P * 2^p_exp
#ifdef XDTTS_LIFT_TRACE
    long long* trace;            // timing experiments: [cta][role 4][tile 32][2] clock64 stamps
#endif
};

struct LiftSmem {
    // [P: nbt tiles x (hi | lo)] then the E ring (per stage [hi | lo]), then barriers and the small tables
    __host__ __device__ static int a_tile(int kchunks) { return 2 * kchunks * LT_BM * 16; }
    __host__ __device__ static int e_plane(int kchunks) { return kchunks * LT_BN * 16; }
    __host__ __device__ static int bar_off(int nbt, int kchunks) { return nbt * a_tile(kchunks) + LT_STAGES * 2 * e_plane(kchunks); }
    static constexpr int N_BARS = 3 * LT_STAGES + LT_STAGES * LT_MAX_BT + LT_MAX_BT;
    static constexpr int BAR_SLOTS = (N_BARS + 1 + 1) & ~1;   // barriers + the TMEM address, padded to 16 bytes
    static constexpr int TAIL = 8 * BAR_SLOTS + 4 * (8 * LT_MAX_KC + 2 * 2 * LT_PGROUPS * LT_BN + LT_CF_RING * LT_BN) + 16 * LT_TILE_SM;
    __host__ __device__ static int total(int nbt, int kchunks) { return bar_off(nbt, kchunks) + TAIL + 16; }
};

// CPT: K chunks per producer thread (ceil(kchunks / LT_PGROUPS)); LD: floats between the records of consecutive frames
// (0: p.ld at run time; as a constant, the epilogue's 64 stores per warp and bin tile are one instruction each)
template <int DELOG, int CPT, int LD>
__global__ void __launch_bounds__(LT_THREADS, 1) gl_lift_tc_kernel(const LiftParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 15) & ~(uintptr_t)15);
    const int a_tile = LiftSmem::a_tile(p.kchunks), e_plane = LiftSmem::e_plane(p.kchunks);
    uint8_t* sa = smem;
    uint8_t* se = smem + p.nbt * a_tile;
    uint64_t* bars = (uint64_t*)(smem + LiftSmem::bar_off(p.nbt, p.kchunks));
    uint64_t* full = bars;                               // [S] E stage written (one arrival per producer warp)
    uint64_t* empty = bars + LT_STAGES;                  // [S] E stage consumed (tcgen05.commit)
    uint64_t* tempty = bars + 2 * LT_STAGES;             // [S] accumulator set drained (the epilogue warps)
    uint64_t* tfull = bars + 3 * LT_STAGES;              // [S][LT_MAX_BT] accumulator of one bin tile complete
    uint64_t* a_bar = tfull + LT_STAGES * LT_MAX_BT;     // [LT_MAX_BT] bin tile of P landed
    uint32_t* tmem_ptr = (uint32_t*)(a_bar + LT_MAX_BT);
    float* wn = (float*)(bars + LiftSmem::BAR_SLOTS);       // [8 LT_MAX_KC] the pseudo-inverse's Nyquist row (bin M)
    float* pmax = wn + 8 * LT_MAX_KC;                    // [2][LT_PGROUPS][LT_BN] per-group maxima of a frame's mel column
    float* nqp = pmax + 2 * LT_PGROUPS * LT_BN;          // [2][LT_PGROUPS][LT_BN] per-group partial sums of the Nyquist bin
    float* cfac = nqp + 2 * LT_PGROUPS * LT_BN;          // [LT_CF_RING][LT_BN] power of two the epilogue owes each frame
    int4* tile_sm = (int4*)(cfac + LT_CF_RING * LT_BN);  // this CTA's first LT_TILE_SM tile records (a global read per tile and
                                                         // role is a serialised L2 round trip)

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // the shuffle tells the compiler it is warp-uniform
    const int part = blockIdx.x % p.n_part, g = blockIdx.x / p.n_part;
    const int nbt = p.nbt;
    const int ld = LD ? LD : p.ld;
    if (threadIdx.x == 0) LT_GSTAMP(0, 0);
    const uint32_t tmem_cols = nbt * LT_STAGES * LT_BN <= 128 ? 128u : (nbt * LT_STAGES * LT_BN <= 256 ? 256u : 512u);

    const int my_tiles = g < p.n_tiles ? (p.n_tiles - g + p.groups - 1) / p.groups : 0;   // tiles g, g + G, ... of this CTA

    // producer thread (f, cg): frame f of a tile, K chunks cg, cg + 5, ... (8 mel rows each).  The loads of the CTA's first tile go
    // out before anything else: their ~3 us (cold HBM, one line per mel row) then overlap the set-up below instead of following it.
    const int ptid = threadIdx.x - 32 * LT_EPI_WARPS;
    const int f = ptid & (LT_BN - 1), cg = ptid / LT_BN;
    constexpr int NX = CPT * 8;
    constexpr bool AHEAD = CPT <= 2;                 // the prefetched tile lives in registers
    auto issue_loads_rec = [&](const int4 tl, float* x) {
        const int T = tl.y;
        const bool fv = f < tl.w && tl.z + f < T;
        const float* src = p.mel_arena + (size_t)tl.x * p.n_mels + tl.z + f;   // row 0 of the utterance, this thread's frame
#pragma unroll
        for (int j = 0; j < CPT; j++)
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int m = (cg + LT_PGROUPS * j) * 8 + r;   // 32-bit offsets: a plan holds fewer than 2^31 / K frames
                x[8 * j + r] = (fv && m < p.n_mels) ? __ldg(src + m * T) : -INFINITY;   // -inf: no sample
            }
    };
    float xq[NX];
    if (AHEAD && my_tiles > 0 && warp >= LT_EPI_WARPS && warp < LT_EPI_WARPS + LT_PRO_WARPS) issue_loads_rec(__ldg(&p.tiles[g]), xq);

    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; s++) {
            mbar_init(&full[s], LT_PRO_WARPS);
            mbar_init(&empty[s], 1);
            mbar_init(&tempty[s], LT_EPI_WARPS);
            for (int b = 0; b < LT_MAX_BT; b++) mbar_init(&tfull[s * LT_MAX_BT + b], 1);
        }
        for (int b = 0; b < LT_MAX_BT; b++) mbar_init(&a_bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == LT_EPI_WARPS + LT_PRO_WARPS) {   // LT_STAGES sets of nbt accumulators, 128 lanes x LT_BN fp32 columns each
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8 * LT_MAX_KC; i += LT_THREADS) wn[i] = i < p.n_mels ? p.pinv_nyq[(size_t)i * p.pinv_ld] : 0.f;
    for (int i = threadIdx.x; i < LT_TILE_SM && g + i * p.groups < p.n_tiles; i += LT_THREADS) tile_sm[i] = p.tiles[g + i * p.groups];
    auto tile_rec = [&](int it_) -> int4 { return it_ < LT_TILE_SM ? tile_sm[it_] : p.tiles[g + it_ * p.groups]; };
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    if (threadIdx.x == 0) LT_GSTAMP(1, 0);

    if (warp == LT_EPI_WARPS + LT_PRO_WARPS) {
        // ===================== MMA issuer.  The whole warp walks the loop (warp-uniform control flow and operands: the
        // descriptors live in uniform registers, ~3 instructions per MMA; issued from inside a one-lane branch each MMA cost
        // ~20 dependent instructions of one warp = 120 cycles against the 32 the tensor pipe needs); one elected lane issues.
        const bool leader = elect_one();
        if (leader) {
            for (int b = 0; b < nbt; b++) {   // one bulk copy per bin tile: the first MMAs start when a quarter of P has landed
                mbar_expect_tx(&a_bar[b], (uint32_t)a_tile);
                bulk_g2s(sa + b * a_tile, p.a_image + (size_t)(part * nbt + b) * a_tile, (uint32_t)a_tile, &a_bar[b]);
            }
        }
        // instruction descriptor: D fp32, A/B fp16, both K-major, N = the plan's tile extent, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.mma_n >> 3) << 17) | ((uint32_t)(LT_BM >> 4) << 24);
        constexpr uint32_t A_LBO = LT_BM * 16, E_LBO = LT_BN * 16, SBO = 128;
        constexpr uint32_t A_HI = ((A_LBO >> 4) << 16), E_HI = ((E_LBO >> 4) << 16);   // low words' LBO fields
        constexpr uint32_t D_HI = (SBO >> 4) | (1u << 14);                               // high word: SBO, version 1
        const int ksteps = p.kchunks >> 1;
        const uint32_t sa_u = smem_u32(sa), se_u = smem_u32(se);
        for (int it = 0; it < my_tiles; it++) {
            const int s = it % LT_STAGES;
            const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
            if (lane == 0) {   // one lane polls (32 spinning lanes cost this scheduler 15% of its issue slots), the rest wait at the warp barrier
                mbar_wait(&tempty[s], ph ^ 1u);   // epilogue has drained this accumulator set
                mbar_wait(&full[s], ph);          // producers have written this E stage
            }
            __syncwarp();
            tc_fence_after();
            if (lane == 0) LT_STAMP(1, it, 0);
            const uint32_t e_hi = ((se_u + (uint32_t)(s * 2 * e_plane)) >> 4) | E_HI, e_lo = e_hi + (uint32_t)(e_plane >> 4);
            for (int b = 0; b < nbt; b++) {
                if (it == 0) {
                    if (lane == 0) mbar_wait(&a_bar[b], 0);
                    __syncwarp();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)((s * nbt + b) * LT_BN);
                const uint32_t a_hi = ((sa_u + (uint32_t)(b * a_tile)) >> 4) | A_HI, a_lo = a_hi + (uint32_t)(a_tile >> 5);
                if (leader) {
#pragma unroll
                    for (int pass = 0; pass < 3; pass++) {   // Phi Ehi, Phi Elo, Plo Ehi
                        const uint32_t ab = pass == 2 ? a_lo : a_hi, eb = pass == 1 ? e_lo : e_hi;
#pragma unroll
                        for (int k = 0; k < LT_MAX_KC / 2; k++)   // UMMA_K = 16 fp16 = two K chunks: the address field moves by 2 LBO
                            if (k < ksteps && !(XDTTS_LIFT_SKIP & 1))
                                tc_mma_f16(d_tmem, ab + (uint32_t)k * (2 * A_LBO >> 4), eb + (uint32_t)k * (2 * E_LBO >> 4), D_HI, idesc, (pass | k) ? 1u : 0u);
                    }
                    tc_commit(&tfull[s * LT_MAX_BT + b]);   // this bin tile's accumulator complete -> epilogue
                }
            }
            if (leader) tc_commit(&empty[s]);    // E stage free when these MMAs retire
            if (lane == 0) LT_STAMP(1, it, 1);
            __syncwarp();
        }
    } else if (warp >= LT_EPI_WARPS) {
        // ===================== producers (thread (f, cg), see the top of the kernel).
        // A warp's 32 lanes are 32 consecutive frames of one mel row: every global load is one 128-byte line, every
        // shared-memory store 32 consecutive 16-byte core-matrix rows.  Loads run one tile ahead of the conversion.
        float wreg[NX];                                  // this thread's entries of the pseudo-inverse's Nyquist row
#pragma unroll
        for (int j = 0; j < CPT; j++)
#pragma unroll
            for (int r = 0; r < 8; r++) wreg[8 * j + r] = (cg + LT_PGROUPS * j) < p.kchunks ? wn[(cg + LT_PGROUPS * j) * 8 + r] : 0.f;
        auto issue_loads = [&](int it_, float* x) { issue_loads_rec(tile_rec(it_), x); };
        float nq_prev = 0.f, sh_prev = 0.f;   // cg == 0: the Nyquist partial sum and 2^shift of the previous tile
        int4 tl_prev = make_int4(0, 0, 0, 0);
        auto finish_nyquist = [&](int it_prev) {   // after the producers' barrier that follows tile it_prev
            if (part == 0 && cg == 0 && f < tl_prev.w && tl_prev.z + f < tl_prev.y) {
                float v = nq_prev;
                const float* pp = nqp + (it_prev & 1) * LT_PGROUPS * LT_BN + f;
#pragma unroll
                for (int c = 1; c < LT_PGROUPS; c++) v += pp[c * LT_BN];
                p.S[((size_t)tl_prev.x + tl_prev.z + f) * ld + p.n_mt * LT_BM] =
                    p.power == 1.0f ? fmaxf(v * sh_prev, 0.f) : pow_pos(v * sh_prev, p.power);
            }
        };
        for (int it = 0; it < my_tiles; it++) {
            const int s = it % LT_STAGES;
            float x[NX];
            if (AHEAD) {
#pragma unroll
                for (int i = 0; i < NX; i++) x[i] = xq[i];
                if (it + 1 < my_tiles) issue_loads(it + 1, xq);
            } else {
                issue_loads(it, x);
            }
            // the frame's exponent: every value of the column is scaled by 2^-shift so that the largest lands in (2^7, 2^8]
            float mx = DELOG == 2 ? 0.f : -INFINITY;
#pragma unroll
            for (int i = 0; i < NX; i++) mx = DELOG == 2 ? fmaxf(mx, x[i] == -INFINITY ? 0.f : fabsf(x[i])) : fmaxf(mx, x[i]);
            float* pm = pmax + (it & 1) * LT_PGROUPS * LT_BN + f;
            pm[cg * LT_BN] = mx;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LT_PRO_WARPS) : "memory");
#pragma unroll
            for (int c = 0; c < LT_PGROUPS; c++) mx = fmaxf(mx, pm[c * LT_BN]);
            float shift;
            if (DELOG == 2) shift = (float)((int)((__float_as_uint(mx) >> 23) & 0xFFu) - 126);   // mx < 2^shift
            else shift = ceilf(mx * Delog<DELOG>::c_hi);
            shift = fminf(fmaxf(shift, -100.f), 118.f) - (float)LT_E_SHIFT;
            const float sc = __int_as_float((127 - (int)shift) << 23);   // 2^-shift
            if (it > 0) finish_nyquist(it - 1);
            if (ptid == 0) LT_STAMP(0, it, 0);
            if (ptid == 0 && it == 0) LT_GSTAMP(1, 1);
            mbar_wait(&empty[s], (uint32_t)((it / LT_STAGES) & 1) ^ 1u);
            if (ptid == 0) LT_STAMP(3, it, 0);
            uint8_t* ehi = se + s * 2 * e_plane + f * 16;
            float nq = 0.f;
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const int c = cg + LT_PGROUPS * j;
                if (!(XDTTS_LIFT_SKIP & 4) && c < p.kchunks) {
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int r = 0; r < 8; r += 2) {
                        const float e0 = delog_scaled<DELOG>(x[8 * j + r], sc), e1 = delog_scaled<DELOG>(x[8 * j + r + 1], sc);
                        nq = fmaf(wreg[8 * j + r], e0, nq);
                        nq = fmaf(wreg[8 * j + r + 1], e1, nq);
                        const __half2 h = __floats2half2_rn(e0, e1);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
                        hw[r >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                        lw[r >> 1] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    *reinterpret_cast<uint4*>(ehi + c * (LT_BN * 16)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(ehi + e_plane + c * (LT_BN * 16)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
            }
            // D = (P 2^p_exp)(E 2^-shift): the accumulator owes 2^(shift - p_exp).  ^power turns that into the ADDEND power * (shift -
            // p_exp) of the epilogue's 2^(power * log2 x): one FFMA instead of two multiplies per value.  (The operands are scaled
            // so that the accumulator is O(1): the log2 of a value around 2^20 is rounded at 2^-19, 1e-6 of the result.)
            if (cg == 0) cfac[(it % LT_CF_RING) * LT_BN + f] = p.power == 1.0f ? exp2_int((int)shift - (int)p.p_exp) : p.power * (shift - p.p_exp);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (ptid == 0) LT_STAMP(0, it, 1);
            if (cg == 0) {
                nq_prev = nq;
                sh_prev = exp2_int((int)shift);
                tl_prev = tile_rec(it);
            } else {
                nqp[((it & 1) * LT_PGROUPS + cg) * LT_BN + f] = nq;
            }
        }
        if (my_tiles > 0) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * LT_PRO_WARPS) : "memory");
            finish_nyquist(my_tiles - 1);
        }
    } else {
        // ===================== epilogue: TMEM lane quarter q <-> bins bt*128 + 32 q + lane.  A work item is (bin tile, half of the
        // tile's 64 columns); the four warps of a quarter take items j, j + 4, ...  Four epilogue warps per scheduler: six
        // instructions per value over 16.4 M values make this role the one that sets the tile time (it is issue-bound).
        const int q = warp & 3, j = warp >> 2;
        constexpr int CW = 32;
        const bool plain = (XDTTS_LIFT_SKIP & 16) ? true : p.power == 1.0f;
        for (int it = 0; it < my_tiles; it++) {
            const int s = it % LT_STAGES;
            const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
            const int4 tl = tile_rec(it);
            for (int item = j; item < 2 * nbt; item += LT_EPI_WARPS / 4) {
                const int b = item >> 1, ch = item & 1;
                const int nf = min(tl.w, tl.y - tl.z) - ch * CW;    // frames of these columns that exist
                if (nf <= 0 || (XDTTS_LIFT_SKIP & 2)) continue;
                const bool full_tile = nf >= CW;
                float* out = p.S + ((size_t)tl.x + tl.z + ch * CW) * ld + (part * nbt + b) * LT_BM + q * 32 + lane;
                const float4* cf4 = reinterpret_cast<const float4*>(cfac + (it % LT_CF_RING) * LT_BN + ch * CW);
                mbar_wait(&tfull[s * LT_MAX_BT + b], ph);
                tc_fence_after();
                if (threadIdx.x == 0 && item == j) LT_STAMP(2, it, 0);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((s * nbt + b) * LT_BN + ch * CW);
                uint32_t v[CW];
                if (XDTTS_LIFT_SKIP & 32) {
#pragma unroll
                    for (int i = 0; i < CW; i++) v[i] = 0x3f800000u + lane + i;
                } else {
                    tc_ld32(taddr, v);
                    tc_wait_ld();
                }
#pragma unroll
                for (int c0 = 0; c0 < CW; c0 += 4) {
                    const float4 c = cf4[c0 >> 2];   // the frames' powers of two (a broadcast read)
                    const float cf[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float r = plain ? fmaxf(__uint_as_float(v[c0 + i]) * cf[i], 0.f) : pow_pos_add(__uint_as_float(v[c0 + i]), p.power, cf[i]);
                        // 32 lanes = 32 consecutive bins of one frame: every store is one 128-byte line
                        if ((XDTTS_LIFT_SKIP & 8) ? r == 123456.789f : (full_tile || c0 + i < nf)) __stcs(out + (c0 + i) * ld, r);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[s]);
            if (threadIdx.x == 0) LT_STAMP(2, it, 1);
        }
    }

    if (threadIdx.x == 0) LT_GSTAMP(2, 0);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) LT_GSTAMP(0, 1);
    if (warp == LT_EPI_WARPS + LT_PRO_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// CUDA-core fp32 form of the same lift (all K bins, Nyquist included): the path for mel bases wider than 96 rows,
// whose P tile does not fit beside the E ring, and the on-device cross-check of the tensor path (XDTTS_LIFT_F32=1).
// grid (1, ceil(maxT/16), n_utt); thread <-> bins k, k + 128, ...; 16 frames per block; the plain fp32 FMA sum is
// 5e-7 of full scale from the fp64 result (the gate is 1e-5).
constexpr int LF_TT = 16;
template <int DELOG>
__global__ void __launch_bounds__(128) gl_lift_f32_kernel(const float* __restrict__ mel_arena, const float* __restrict__ pinvT,
                                                          const int* __restrict__ utt_T, const int* __restrict__ utt_foff,
                                                          int n_mels, int K, int ld, float power, float* __restrict__ S) {
    extern __shared__ __align__(16) float e_sm[];   // [n_mels][LF_TT]
    const int u = blockIdx.z, T = utt_T[u], t0 = blockIdx.y * LF_TT;
    if (t0 >= T) return;
    const size_t foff = (size_t)utt_foff[u];
    const float* mel = mel_arena + foff * n_mels;
    for (int i = threadIdx.x; i < n_mels * LF_TT; i += 128) {
        const int m = i / LF_TT, tt = i % LF_TT;
        e_sm[i] = (t0 + tt < T) ? delog_value<DELOG>(mel[(size_t)m * T + t0 + tt]) : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 128) {
        float acc[LF_TT];
#pragma unroll
        for (int tt = 0; tt < LF_TT; tt++) acc[tt] = 0.f;
#pragma unroll 4
        for (int m = 0; m < n_mels; m++) {
            const float a = pinvT[(size_t)m * K + k];
            const float4* er = reinterpret_cast<const float4*>(e_sm + m * LF_TT);
#pragma unroll
            for (int q = 0; q < LF_TT / 4; q++) {
                const float4 x = er[q];
                acc[4 * q + 0] = fmaf(a, x.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(a, x.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(a, x.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(a, x.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int tt = 0; tt < LF_TT; tt++)
            if (t0 + tt < T) S[(foff + t0 + tt) * ld + k] = power == 1.0f ? fmaxf(acc[tt], 0.f) : pow_pos(acc[tt], power);
    }
}

}  // namespace

// host: the device image of the pseudo-inverse for gl_lift_tc_kernel.  pinv: [K][n_mels] row-major (K = M + 1; the
// last row is the Nyquist bin and is not part of the image).  Layout [bin tile][plane hi, lo][K chunk][row 0..127][8 fp16]
// (core matrices of 8 rows x 16 bytes, un-swizzled); values are pinv * 2^(*p_exp), the power of two that brings the
// largest entry into [4, 8) -- fp16 has 5 exponent bits, and the low halves should stay out of its subnormals.
std::vector<float> gl_lift_build_image(const float* pinv, int K, int n_mels, float* p_exp) {
    const int M = K - 1, n_mt = (M + LT_BM - 1) / LT_BM, kchunks = 2 * ((n_mels + 15) / 16);
    float amax = 0.f;
    for (size_t i = 0; i < (size_t)M * n_mels; i++)
        if (std::isfinite(pinv[i])) amax = std::max(amax, std::fabs(pinv[i]));
    int ex = 0;
    if (amax > 0.f) std::frexp(amax, &ex);   // amax in [2^(ex-1), 2^ex)
    const int pe = amax > 0.f ? 3 - ex : 0;
    *p_exp = (float)pe;
    std::vector<float> img((size_t)n_mt * 2 * kchunks * LT_BM * 8 / 2, 0.f);   // fp16 pairs in float-sized slots
    uint16_t* h16 = reinterpret_cast<uint16_t*>(img.data());
    for (int bin = 0; bin < M; bin++) {
        const int mt = bin / LT_BM, r = bin % LT_BM;
        for (int m = 0; m < n_mels; m++) {
            const float x = std::ldexp(pinv[(size_t)bin * n_mels + m], pe);
            const __half hi = __float2half_rn(x);
            const __half lo = __float2half_rn(x - __half2float(hi));
            const size_t at = ((((size_t)mt * 2 + 0) * kchunks + (m >> 3)) * LT_BM + r) * 8 + (m & 7);
            memcpy(&h16[at], &hi, 2);
            memcpy(&h16[at + (size_t)kchunks * LT_BM * 8], &lo, 2);
        }
    }
    return img;
}

static int lift_bin_tiles_per_cta(int n_mt, int kchunks);
bool gl_lift_uses_tensor_cores(int n_mels, int K);

// frames per tile for a plan: the extent (a multiple of 8, at most 64) whose tile count fills the CTAs' last round best.
// 32 x 1000 frames in tiles of 64 are 500 tiles = 3.4 rounds of 148 CTAs, paid as 4; tiles of 56 are 576 = 3.9 rounds, 12% less
// work on the longest CTA.  The cost model is rounds x (extent + 8): a tile has a fixed cost of about 8 frames' worth.
int gl_lift_tile_frames(const int* Ts, int B, int n_mels, int K, int sm_count) {
    if (!gl_lift_uses_tensor_cores(n_mels, K)) return LT_BN;
    const int n_mt = (K - 1) / LT_BM, kchunks = 2 * ((n_mels + 15) / 16);
    const int n_part = n_mt / lift_bin_tiles_per_cta(n_mt, kchunks);
    const int groups = std::max(1, sm_count / n_part);
    int best = LT_BN;
    long best_cost = -1;
    for (int bn = LT_BN; bn >= 16; bn -= 8) {
        long tiles = 0;
        for (int b = 0; b < B; b++) tiles += (Ts[b] + bn - 1) / bn;
        const long rounds = (tiles + groups - 1) / groups, cost = rounds * (bn + 8);
        if (best_cost < 0 || cost < best_cost) best = bn, best_cost = cost;
    }
    return best;
}

// bin tiles one CTA keeps resident: the largest of 4, 2, 1 that divides the tile count and fits shared memory
static int lift_bin_tiles_per_cta(int n_mt, int kchunks) {
    for (int nbt = LT_MAX_BT; nbt > 1; nbt >>= 1)
        if (n_mt % nbt == 0 && LiftSmem::total(nbt, kchunks) <= LT_SMEM_MAX) return nbt;
    return 1;
}

// false: the tensor path does not apply (n_mels > 128, or bins not a multiple of 128): gl_launch_lift takes the fp32 kernel
bool gl_lift_uses_tensor_cores(int n_mels, int K) {
    static const bool force_f32 = getenv("XDTTS_LIFT_F32") != nullptr;
    return !force_f32 && (n_mels + 7) / 8 <= LT_MAX_KC && (K - 1) % LT_BM == 0;
}

template <int DELOG, int CPT>
static cudaError_t lift_attr(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(gl_lift_tc_kernel<DELOG, CPT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gl_lift_tc_kernel<DELOG, CPT, 772>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gl_lift_tc_kernel<DELOG, CPT, 1540>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gl_lift_tc_kernel<DELOG, CPT, 3076>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

cudaError_t gl_lift_prepare(int n_mels) {
    if ((n_mels + 7) / 8 > LT_MAX_KC) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = lift_attr<0, 2>(LT_SMEM_MAX);
    if (e == cudaSuccess) e = lift_attr<1, 2>(LT_SMEM_MAX);
    if (e == cudaSuccess) e = lift_attr<2, 2>(LT_SMEM_MAX);
    if (e == cudaSuccess) e = lift_attr<0, 4>(LT_SMEM_MAX);
    if (e == cudaSuccess) e = lift_attr<1, 4>(LT_SMEM_MAX);
    if (e == cudaSuccess) e = lift_attr<2, 4>(LT_SMEM_MAX);
    return e;
}

template <int DELOG, int CPT>
static void lift_launch(const LiftParams& p, int grid, int sm, cudaStream_t s) {
    switch (p.ld) {   // the record strides of n_fft 512 / 1024 / 2048 (3 n_fft / 2 + 4 floats) as compile-time constants
        case 772: gl_lift_tc_kernel<DELOG, CPT, 772><<<grid, LT_THREADS, sm, s>>>(p); break;
        case 1540: gl_lift_tc_kernel<DELOG, CPT, 1540><<<grid, LT_THREADS, sm, s>>>(p); break;
        case 3076: gl_lift_tc_kernel<DELOG, CPT, 3076><<<grid, LT_THREADS, sm, s>>>(p); break;
        default: gl_lift_tc_kernel<DELOG, CPT, 0><<<grid, LT_THREADS, sm, s>>>(p);
    }
}

// S[(foff + t) * ld + k] = max(0, sum_m pinv[k][m] delog(mel[m][t])) ^ power for k = 0..M (k = M: the Nyquist slot).
// pinvT: [n_mels][K] (the transposed pseudo-inverse); a_image, p_exp: gl_lift_build_image; tiles: records of at most
// tile_frames frames (gl_lift_tile_frames).  *n_kernels: launches made.
cudaError_t gl_launch_lift(const float* mel_arena, const float* a_image, float p_exp, const float* pinvT, const int4* tiles, int n_tiles, int tile_frames,
                           const int* utt_T, const int* utt_foff, int n_utt, int max_T, int n_mels, int K, int ld, float power,
                           int delog, int sm_count, float* S, cudaStream_t s, int* n_kernels) {
    const int M = K - 1;
    if (n_kernels) *n_kernels = 1;
    if (!gl_lift_uses_tensor_cores(n_mels, K)) {
        if (n_mels > 256) return cudaErrorInvalidValue;
        dim3 grid(1, (max_T + LF_TT - 1) / LF_TT, n_utt);
        const size_t sm = (size_t)n_mels * LF_TT * sizeof(float);
        if (delog == 0) gl_lift_f32_kernel<0><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        else if (delog == 1) gl_lift_f32_kernel<1><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        else gl_lift_f32_kernel<2><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        return cudaGetLastError();
    }
    LiftParams p;
    p.mel_arena = mel_arena; p.a_image = reinterpret_cast<const uint8_t*>(a_image); p.tiles = tiles; p.S = S;
    p.n_tiles = n_tiles; p.n_mt = M / LT_BM; p.n_mels = n_mels; p.kchunks = 2 * ((n_mels + 15) / 16); p.ld = ld;
    p.power = power; p.p_exp = p_exp;
    p.nbt = lift_bin_tiles_per_cta(p.n_mt, p.kchunks);
    p.n_part = p.n_mt / p.nbt;
    p.groups = sm_count / p.n_part;
    if (p.groups < 1) p.groups = 1;
    if (p.groups > n_tiles) p.groups = n_tiles;
    p.pinv_nyq = pinvT + M; p.pinv_ld = K;
    const int grid = p.n_part * p.groups, sm = LiftSmem::total(p.nbt, p.kchunks);
    const bool wide = p.kchunks > 2 * LT_PGROUPS;   // more than two K chunks per producer thread
    p.mma_n = std::min(LT_BN, (std::max(tile_frames, 1) + 15) & ~15);
#ifdef XDTTS_LIFT_TRACE
    static long long* d_trace = nullptr;
    const size_t tr_n = (size_t)grid * 4 * 32 * 2;
    if (!d_trace) cudaMalloc(&d_trace, 296 * 4 * 32 * 2 * 8);
    cudaMemsetAsync(d_trace, 0, tr_n * 8, s);
    p.trace = d_trace;
#endif
    if (delog == 0) wide ? lift_launch<0, 4>(p, grid, sm, s) : lift_launch<0, 2>(p, grid, sm, s);
    else if (delog == 1) wide ? lift_launch<1, 4>(p, grid, sm, s) : lift_launch<1, 2>(p, grid, sm, s);
    else wide ? lift_launch<2, 4>(p, grid, sm, s) : lift_launch<2, 2>(p, grid, sm, s);
#ifdef XDTTS_LIFT_TRACE
    if (getenv("XDTTS_LIFT_TRACE_PRINT")) {
        cudaStreamSynchronize(s);
        std::vector<long long> tr(tr_n);
        cudaMemcpy(tr.data(), d_trace, tr_n * 8, cudaMemcpyDeviceToHost);
        for (int cta : {0, grid - 1}) {
            long long t0 = 0;
            for (size_t i = 0; i < 4 * 32 * 2; i++) { long long v = tr[(size_t)cta * 256 + i]; if (v && (!t0 || v < t0)) t0 = v; }
            fprintf(stderr, "lift trace cta %d (cycles from first stamp): it | prod start, slot free, done | mma start, issued | epi start, done\n", cta);
            {
                auto G = [&](int role, int k) { return tr[(((size_t)cta * 4 + role) * 32 + 31) * 2 + k]; };
                long long e0 = 0, e1 = 0;
                for (int c = 0; c < grid; c++) { long long a = tr[(((size_t)c * 4 + 0) * 32 + 31) * 2], b = tr[(((size_t)c * 4 + 0) * 32 + 31) * 2 + 1]; if (!e0 || a < e0) e0 = a; if (b > e1) e1 = b; }
                fprintf(stderr, "  ns: all CTAs entry..exit %lld | this CTA entry +%lld, setup done +%lld, first tile loaded +%lld, epilogue warp 0 done +%lld, exit +%lld\n", e1 - e0, G(0, 0) - e0, G(1, 0) - e0, G(1, 1) - e0, G(2, 0) - e0, G(0, 1) - e0);
            }
            for (int it = 0; it < 31; it++) {
                auto g = [&](int role, int k) { long long v = tr[(((size_t)cta * 4 + role) * 32 + it) * 2 + k]; return v ? v - t0 : -1; };
                if (g(0, 0) < 0) break;
                fprintf(stderr, "  %2d | %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld\n", it, g(0, 0), g(3, 0), g(0, 1), g(1, 0), g(1, 1), g(2, 0), g(2, 1));
            }
        }
    }
#endif
    return cudaGetLastError();
}

}  // namespace xdtts
