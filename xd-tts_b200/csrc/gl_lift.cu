// mel -> linear lift on the 5th-generation tensor cores:  S = max(0, pinv(basis) . delog(mel)) ^ power
// (step 1 of griffin_lim::GriffinLim::infer; reference call site
// /root/reference src/lib.rs:141, parameters from src/tacotron2/mod.rs:453-456).
//
// The product is a [bins x n_mels] . [n_mels x frames] GEMM with a tiny inner dimension (80): 82 kFLOP and
// 2.4 kB per frame, i.e. HBM-bound -- IF the multiply-adds are off the CUDA cores (on the fp32 pipe the 1.3 G
// FMAs of a 32 x 1000 batch alone take 35 us at 100% issue; the fp64 kernel of round 1 took 220-316 us).
// One tile of the GEMM is
//
//     D[bin 0..127][frame 0..31] = sum_m P[bin][m] * E[frame][m]          (UMMA M = 128 bins, N = 32 frames)
//
// with both operands K-major (mel index contiguous) in 128-byte-swizzled shared memory and fp32 accumulation
// in TMEM.  Precision: the pseudo-inverse has 36% negative entries and the gate is 1e-5 of full
// scale; a single tf32 pass gives 9e-4, bf16 split in three 2e-5, tf32 split in three 3e-7 (the plain fp32 FMA
// sum gives 5e-7) -- hence "tf32x3": x = hi + lo, D = Phi Ehi + Phi Elo + Plo Ehi, three passes of
// tcgen05.mma kind::tf32 into the same accumulator.
//
// Work split: CTA (mt, g) owns bin tile mt (its P tile, hi and lo planes, stays in shared memory for the whole
// kernel: one bulk copy of a pre-swizzled host-built image) and walks frame tiles g, g + G, ...  Bins as the
// TMEM lane dimension make the epilogue's stores coalesced: for one frame (TMEM column) the 32 lanes of a warp
// hold 32 consecutive bins = 128 contiguous bytes of the frame-major state record the iteration kernel reads.
//
// Warp roles (416 threads, one CTA per SM):
//     warps 0-3   epilogue     tcgen05.ld (lane quarter = warp) -> clamp -> ^power -> S
//     warps 4-11  producers    mel [n_mels][T] -> delog -> tf32 hi / lo -> swizzled E tile (4-stage ring, loads 3 tiles ahead)
//     warp  12    TMEM alloc + one thread issuing the MMAs and commits
// The Nyquist bin (bin M, a 129th row of the last tile otherwise) is summed by the producers of bin tile 0 on the
// CUDA cores, from the de-logged values they hold anyway.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gl_host.h"

namespace xdtts {

namespace {

// timing experiments only (tools/build_lift_variants.py): bit 0 skips the MMAs, bit 1 the epilogue's math and stores,
// bit 2 the producers' conversion and shared-memory stores.  0 in the product.
#ifndef XDTTS_LIFT_SKIP
#define XDTTS_LIFT_SKIP 0
#endif

constexpr int LT_BM = 128;       // bins per tile
constexpr int LT_BN = 64;        // frames per tile
constexpr int LT_STAGES = 2;     // E-tile ring and TMEM accumulators
constexpr int LT_AHEAD = 1;      // tiles whose mel loads are in flight ahead of the conversion
constexpr int LT_EPI_WARPS = 4;   // epilogue warps: TMEM lane quarter = warp (13 warps keep 128 registers per thread)
constexpr int LT_PRO_WARPS = 8;   // producer warps: frame groups (warp & 3, + 4), mel half = warp >> 2
constexpr int LT_THREADS = 32 * (LT_EPI_WARPS + LT_PRO_WARPS + 1);
constexpr int LT_TILE_SM = 128;  // tile records staged in shared memory per CTA
constexpr int LT_MAX_KB = 3;     // shared memory holds the P tile (2 planes) + the E ring for n_mels <= 96; wider bases take gl_lift_f32_kernel

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
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs (fp32 containers, low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major tile with 128-byte rows (32 tf32) in the 128-byte swizzle: 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// round to tf32 (10 mantissa bits), nearest with ties away from zero -- cvt.rna.tf32.f32, which sm_100 expands
// into ~15 instructions; for the finite values of this kernel the two integer operations below are the same function
// (and the host builds the pseudo-inverse image with them, tf32_rna_bits)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// e^v (DELOG 0) / 10^v (DELOG 1) in six instructions: 2^t with t = v * c split so that the rounding of the product
// is carried into a first-order correction -- t = fl(v c_hi), r = (v c_hi - t) + v c_lo exactly, e = 2^t (1 + r ln 2).
// ~1.5 ulp (ex2.approx is 2 ulp), against ~15 instructions for expf(); exp(-inf) = 0 without a branch.
template <int DELOG>
__device__ __forceinline__ float delog_value(float v) {
    if (DELOG == 2) return v;
    const float c_hi = DELOG == 0 ? 1.4426950216293335f : 3.3219280242919922f;      // log2(e), log2(10) rounded to fp32
    const float c_lo = DELOG == 0 ? 1.9259629911266175e-8f : 7.0595369550985533e-8f; // ... and what the rounding dropped
    const float t = v * c_hi;
    float r = fmaf(v, c_hi, -t);
    r = fmaf(v, c_lo, r);
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(t));
    return v == -INFINITY ? 0.f : fmaf(p, r * 0.69314718055994531f, p);
}

// s ^ power for s > 0 on the special-function unit: 2^(power * log2 s), two MUFU operations and a multiply.  ~1e-6
// relative at full scale (the gate on S is 1e-5 of full scale, tests/test_gpu_gl.py::test_lift_matches_oracle); powf
// costs ~40 instructions per value -- more issue time than the rest of the kernel.
__device__ __forceinline__ float pow_pos(float s, float power) {
    float l, r;   // max(s, 0) -> log2 = -inf at 0 -> 2^-inf = 0: the clamp needs no select
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 0.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * power));
    return r;
}

struct LiftParams {
    const float* mel_arena;      // utterance u: row-major [n_mels][T_u] at float offset foff[u] * n_mels
    const float* a_image;        // [n_mt][2 planes][kblocks][128 rows][32] tf32 values, rows pre-swizzled
    const float* pinv_nyq;       // the pseudo-inverse's row of bin M: entry m at pinv_nyq[m * pinv_ld]
    int pinv_ld;
    const int4* tiles;           // frame tiles: (first frame row of the utterance, its frame count T, first frame of the tile, 0)
    float* S;                    // frame-major state records: S[(foff + t) * ld + k], k < M
    int n_tiles, n_mt, groups, n_mels, kblocks, ld;
    float power;
};

struct LiftSmem {
    // [A hi | A lo] then the E ring (per stage [E hi | E lo]); every tile 1024-byte aligned
    __host__ __device__ static int a_plane(int kblocks) { return kblocks * LT_BM * 128; }
    __host__ __device__ static int e_plane(int kblocks) { return kblocks * LT_BN * 128; }
    __host__ __device__ static int bar_off(int kblocks) { return 2 * a_plane(kblocks) + LT_STAGES * 2 * e_plane(kblocks); }
    __host__ __device__ static int total(int kblocks) { return bar_off(kblocks) + 8 * (4 * LT_STAGES + 4) + 4 * 32 * LT_MAX_KB + 4 * LT_STAGES * LT_BN + 16 * LT_TILE_SM + 1024; }
};

template <int DELOG>
__global__ void __launch_bounds__(LT_THREADS, 1) gl_lift_tc_kernel(const LiftParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_plane = LiftSmem::a_plane(p.kblocks), e_plane = LiftSmem::e_plane(p.kblocks);
    uint8_t* sa = smem;
    uint8_t* se = smem + 2 * a_plane;
    uint64_t* bars = (uint64_t*)(smem + LiftSmem::bar_off(p.kblocks));
    uint64_t* full = bars;                       // [S] E stage written (all producer threads arrive)
    uint64_t* empty = bars + LT_STAGES;          // [S] E stage consumed (tcgen05.commit)
    uint64_t* tfull = bars + 2 * LT_STAGES;      // [S] accumulator complete
    uint64_t* tempty = bars + 3 * LT_STAGES;     // [S] accumulator drained (the epilogue warps)
    uint64_t* a_bar = bars + 4 * LT_STAGES;      // P tile landed
    uint32_t* tmem_ptr = (uint32_t*)(bars + 4 * LT_STAGES + 2);
    float* wn = (float*)(bars + 4 * LT_STAGES + 4);        // [32 LT_MAX_KB] the pseudo-inverse's Nyquist row (bin M)
    float* nq_sm = wn + 32 * LT_MAX_KB;                     // [S][LT_BN] Nyquist partial sums of the upper mel half
    int4* tile_sm = (int4*)(nq_sm + LT_STAGES * LT_BN);     // this CTA's first LT_TILE_SM tile records (a global read per tile and
                                                            // role is a serialised L2 round trip: a third of all stall samples before)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % p.n_mt, g = blockIdx.x / p.n_mt;

    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; s++) {
            mbar_init(&full[s], 32 * LT_PRO_WARPS);
            mbar_init(&empty[s], 1);
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], LT_EPI_WARPS);
        }
        mbar_init(a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == LT_EPI_WARPS + LT_PRO_WARPS) {   // LT_STAGES accumulators of 128 lanes x LT_BN fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)(LT_STAGES * LT_BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 32 * LT_MAX_KB; i += LT_THREADS) wn[i] = i < p.n_mels ? p.pinv_nyq[(size_t)i * p.pinv_ld] : 0.f;
    for (int i = threadIdx.x; i < LT_TILE_SM && g + i * p.groups < p.n_tiles; i += LT_THREADS) tile_sm[i] = p.tiles[g + i * p.groups];
    auto tile_rec = [&](int it_) -> int4 { return it_ < LT_TILE_SM ? tile_sm[it_] : p.tiles[g + it_ * p.groups]; };
    // the zero padding of the E tiles (mel indices n_mels .. 32 kblocks - 1) is written once: producers only touch k < n_mels
    for (int i = threadIdx.x; i < LT_STAGES * 2 * e_plane / 16; i += LT_THREADS) reinterpret_cast<uint4*>(se)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int my_tiles = g < p.n_tiles ? (p.n_tiles - g + p.groups - 1) / p.groups : 0;   // tiles g, g + G, ... of this CTA

    if (warp == LT_EPI_WARPS + LT_PRO_WARPS) {
        // ===================== MMA issuer
        if (lane == 0) {
            mbar_expect_tx(a_bar, (uint32_t)(2 * a_plane));
            bulk_g2s(sa, p.a_image + (size_t)mt * (2 * a_plane / 4), (uint32_t)(2 * a_plane), a_bar);
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = LT_BN, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(LT_BN >> 3) << 17) | ((uint32_t)(LT_BM >> 4) << 24);
            mbar_wait(a_bar, 0);
            for (int it = 0; it < my_tiles; it++) {
                const int s = it % LT_STAGES;
                const uint32_t ph = (uint32_t)((it / LT_STAGES) & 1);
                mbar_wait(&tempty[s], ph ^ 1u);   // epilogue has drained this accumulator
                mbar_wait(&full[s], ph);          // producers have written this E stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * LT_BN);
                const uint32_t e_hi = smem_u32(se + s * 2 * e_plane), e_lo = e_hi + e_plane;
                const uint32_t a_hi = smem_u32(sa), a_lo = a_hi + a_plane;
                uint32_t acc = 0;
                for (int pass = 0; pass < 3; pass++) {   // Phi Ehi, Phi Elo, Plo Ehi
                    const uint32_t ab = pass == 2 ? a_lo : a_hi, eb = pass == 1 ? e_lo : e_hi;
                    for (int kb = 0; kb < p.kblocks; kb++) {
                        const uint64_t da = umma_desc_sw128(ab + kb * (LT_BM * 128)), db = umma_desc_sw128(eb + kb * (LT_BN * 128));
#pragma unroll
                        for (int k = 0; k < 4; k++) {   // UMMA_K = 8 tf32 = 32 B along the swizzled row: +2 in the address field
                            if (!(XDTTS_LIFT_SKIP & 1)) tc_mma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
                            acc = 1;
                        }
                    }
                }
                tc_commit(&empty[s]);    // E stage free when these MMAs retire
                tc_commit(&tfull[s]);    // accumulator complete -> epilogue
            }
        }
    } else if (warp >= LT_EPI_WARPS) {
        // ===================== producers.  A warp covers 4 mel rows x 8 frames per step (conflict-free swizzled stores,
        // four full 32-byte sectors per global load); producer warp w owns frame group w & 3 of the tile (a lane: frame
        // f = 8 (w & 3) + lane / 4) and the mel-row groups of half w >> 2 (a lane: rows 4 mg + lane % 4).  Loads run
        // LT_AHEAD tiles ahead of the conversion, so DRAM latency hides behind the tiles in flight.
        const int pw = (warp - LT_EPI_WARPS) & 3, mh = (warp - LT_EPI_WARPS) >> 2;
        const int m4 = lane & 3, f8 = lane >> 2;
        constexpr int FG = LT_BN / 32;                      // frame groups of 8 per producer warp: frames 8 pw + f8 + 32 h
        const int f = pw * 8 + f8;
        constexpr int MGT = LT_MAX_KB * 8;                  // mel groups of 4 in a padded tile (24)
        constexpr int MG = MGT / (LT_PRO_WARPS / 4);        // ... of which this warp converts MG, starting at mg0
        const int mg0 = mh * MG;
        const int mgroups = (p.n_mels + 3) / 4;
        float wreg[MG];                                     // this lane's entries of the pseudo-inverse's Nyquist row
#pragma unroll
        for (int mg = 0; mg < MG; mg++) wreg[mg] = wn[(mg0 + mg) * 4 + m4];
        float xq[LT_AHEAD][FG * MG];
        auto issue_loads = [&](int it_, float* x) {
            const int4 tl = tile_rec(it_);
            const int T = tl.y;
            const float* src = p.mel_arena + (size_t)tl.x * p.n_mels + tl.z + f;   // row 0 of the utterance, this lane's first frame
            int off = (mg0 * 4 + m4) * T;   // 32-bit: a plan holds fewer than 2^31 / K frames
#pragma unroll
            for (int mg = 0; mg < MG; mg++) {
                const bool mv = (mg0 + mg) * 4 + m4 < p.n_mels;
#pragma unroll
                for (int h = 0; h < FG; h++) x[FG * mg + h] = (mv && tl.z + f + 32 * h < T) ? __ldg(src + off + 32 * h) : -INFINITY;   // -inf: no sample
                off += 4 * T;
            }
        };
#pragma unroll
        for (int a = 0; a < LT_AHEAD; a++)
            if (a < my_tiles) issue_loads(a, xq[a]);
        for (int it = 0; it < my_tiles; it++) {
            const int s = it % LT_STAGES;
            float x[FG * MG];
#pragma unroll
            for (int i = 0; i < FG * MG; i++) x[i] = xq[0][i];
#pragma unroll
            for (int a = 0; a + 1 < LT_AHEAD; a++)
#pragma unroll
                for (int i = 0; i < FG * MG; i++) xq[a][i] = xq[a + 1][i];
            if (it + LT_AHEAD < my_tiles) issue_loads(it + LT_AHEAD, xq[LT_AHEAD - 1]);
            mbar_wait(&empty[s], (uint32_t)((it / LT_STAGES) & 1) ^ 1u);
            uint8_t* ehi = se + s * 2 * e_plane + f * 128 + m4 * 4;
            float nq[FG];   // Nyquist-bin partial sums of this lane's frames
#pragma unroll
            for (int h = 0; h < FG; h++) nq[h] = 0.f;
#pragma unroll
            for (int mg = 0; mg < MG; mg++) {
                const int mgg = mg0 + mg;
                if (!(XDTTS_LIFT_SKIP & 4) && mgg < mgroups && mgg * 4 + m4 < p.n_mels) {
                    // row f of K-block mgg >> 3, 16-byte chunk (mgg & 7) ^ (f & 7) (f & 7 == f8 for every h), element m4
                    const int off = (mgg >> 3) * (LT_BN * 128) + (((mgg & 7) ^ f8) << 4);
#pragma unroll
                    for (int h = 0; h < FG; h++) {
                        const float v = x[FG * mg + h];
                        float e = delog_value<DELOG>(v);               // -inf (no sample) -> 0
                        if (DELOG == 2) e = v == -INFINITY ? 0.f : e;
                        const float hi = to_tf32(e);
                        const float lo = to_tf32(e - hi);
                        *reinterpret_cast<float*>(ehi + off + h * (32 * 128)) = hi;
                        *reinterpret_cast<float*>(ehi + e_plane + off + h * (32 * 128)) = lo;
                        nq[h] = fmaf(wreg[mg], e, nq[h]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
            mbar_arrive(&full[s]);
            if (mt == 0) {   // the Nyquist bin (slot M of the frame's record): sum over the four lanes that share a frame,
                             // then over the two warps that share the frame (one per half of the mel rows)
                float* part = nq_sm + s * LT_BN;   // the upper half's partial sums of this stage
#pragma unroll
                for (int h = 0; h < FG; h++) {
                    nq[h] += __shfl_xor_sync(0xffffffffu, nq[h], 1);
                    nq[h] += __shfl_xor_sync(0xffffffffu, nq[h], 2);
                    if (mh == 1 && m4 == 0) part[f + 32 * h] = nq[h];
                }
                asm volatile("bar.sync %0, %1;" ::"r"(1 + pw), "r"(64) : "memory");   // the two warps of frame group pw
                const int4 tl = tile_rec(it);
#pragma unroll
                for (int h = 0; h < FG; h++)
                    if (mh == 0 && m4 == 0 && tl.z + f + 32 * h < tl.y) {
                        const float v = nq[h] + part[f + 32 * h];
                        p.S[((size_t)tl.x + tl.z + f + 32 * h) * p.ld + p.n_mt * LT_BM] = p.power == 1.0f ? fmaxf(v, 0.f) : pow_pos(v, p.power);
                    }
            }
        }
    } else {
        // ===================== epilogue: TMEM lane quarter q <-> bins mt*128 + 32 q + lane
        const int q = warp & 3;
        const int bin = mt * LT_BM + q * 32 + lane;
        const bool plain = p.power == 1.0f;
        for (int it = 0; it < my_tiles; it++) {
            const int s = it % LT_STAGES;
            const int4 tl = tile_rec(it);
            const int nf = min(LT_BN, tl.y - tl.z);   // frames of this tile that exist
            float* out = p.S + ((size_t)tl.x + tl.z) * p.ld + bin;
            mbar_wait(&tfull[s], (uint32_t)((it / LT_STAGES) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * LT_BN);
            const bool full_tile = nf >= LT_BN;   // all but an utterance's last tile
#pragma unroll
            for (int c0 = 0; c0 < ((XDTTS_LIFT_SKIP & 2) ? 0 : LT_BN); c0 += 16) {
                uint32_t v[16];
                tc_ld16(taddr + (uint32_t)c0, v);
                tc_wait_ld();
                float r[16];
#pragma unroll
                for (int i = 0; i < 16; i++) r[i] = plain ? fmaxf(__uint_as_float(v[i]), 0.f) : pow_pos(__uint_as_float(v[i]), p.power);
                // 32 lanes = 32 consecutive bins of one frame: every store is one 128-byte line
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 16; i++) __stcs(out + (c0 + i) * p.ld, r[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (c0 + i < nf) __stcs(out + (c0 + i) * p.ld, r[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[s]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == LT_EPI_WARPS + LT_PRO_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(LT_STAGES * LT_BN)) : "memory");
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

uint32_t tf32_rna_bits(float x) {   // cvt.rna.tf32.f32 on the host: nearest, ties away from zero, 10 mantissa bits
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return u;
    return (u + 0x1000u) & 0xFFFFE000u;
}

}  // namespace

// host: the device image of the pseudo-inverse for gl_lift_tc_kernel.  pinv: [K][n_mels] row-major (K = M + 1; the
// last row is the Nyquist bin and is not part of the image).  Layout [mt][plane hi, lo][kblock][row 0..127][32 floats],
// each 128-byte row stored with its 16-byte chunks XOR-swizzled by (row & 7) -- what TMA's SWIZZLE_128B would write.
std::vector<float> gl_lift_build_image(const float* pinv, int K, int n_mels) {
    const int M = K - 1, n_mt = (M + LT_BM - 1) / LT_BM, kblocks = (n_mels + 31) / 32;
    std::vector<float> img((size_t)n_mt * 2 * kblocks * LT_BM * 32, 0.f);
    for (int mt = 0; mt < n_mt; mt++)
        for (int r = 0; r < LT_BM; r++) {
            const int bin = mt * LT_BM + r;
            if (bin >= M) continue;
            for (int m = 0; m < n_mels; m++) {
                const float x = pinv[(size_t)bin * n_mels + m];
                const uint32_t hb = tf32_rna_bits(x);
                float hi, lo;
                memcpy(&hi, &hb, 4);
                const uint32_t lb = tf32_rna_bits(x - hi);
                memcpy(&lo, &lb, 4);
                const int kb = m >> 5, kk = m & 31;
                const size_t row = (size_t)r * 32 + (size_t)((((kk >> 2) ^ (r & 7)) << 2) + (kk & 3));
                img[(((size_t)mt * 2 + 0) * kblocks + kb) * LT_BM * 32 + row] = hi;
                img[(((size_t)mt * 2 + 1) * kblocks + kb) * LT_BM * 32 + row] = lo;
            }
        }
    return img;
}

int gl_lift_tile_frames() { return LT_BN; }

// false: the tensor path does not apply (n_mels > 96, or bins not a multiple of 128): gl_launch_lift takes the fp32 kernel
bool gl_lift_uses_tensor_cores(int n_mels, int K) {
    static const bool force_f32 = getenv("XDTTS_LIFT_F32") != nullptr;
    return !force_f32 && (n_mels + 31) / 32 <= LT_MAX_KB && (K - 1) % LT_BM == 0;
}

cudaError_t gl_lift_prepare(int n_mels) {
    const int kblocks = (n_mels + 31) / 32;
    if (kblocks > LT_MAX_KB) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(gl_lift_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LiftSmem::total(kblocks));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gl_lift_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LiftSmem::total(kblocks));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gl_lift_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LiftSmem::total(kblocks));
    return e;
}

// S[(foff + t) * ld + k] = max(0, sum_m pinv[k][m] delog(mel[m][t])) ^ power for k = 0..M (k = M: the Nyquist slot).
// pinvT: [n_mels][K] (the transposed pseudo-inverse); a_image: gl_lift_build_image.  *n_kernels: launches made.
cudaError_t gl_launch_lift(const float* mel_arena, const float* a_image, const float* pinvT, const int4* tiles, int n_tiles,
                           const int* utt_T, const int* utt_foff, int n_utt, int max_T, int n_mels, int K, int ld, float power,
                           int delog, int sm_count, float* S, cudaStream_t s, int* n_kernels) {
    const int M = K - 1;
    if (!gl_lift_uses_tensor_cores(n_mels, K)) {
        if (n_mels > 256) return cudaErrorInvalidValue;
        dim3 grid(1, (max_T + LF_TT - 1) / LF_TT, n_utt);
        const size_t sm = (size_t)n_mels * LF_TT * sizeof(float);
        if (delog == 0) gl_lift_f32_kernel<0><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        else if (delog == 1) gl_lift_f32_kernel<1><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        else gl_lift_f32_kernel<2><<<grid, 128, sm, s>>>(mel_arena, pinvT, utt_T, utt_foff, n_mels, K, ld, power, S);
        if (n_kernels) *n_kernels = 1;
        return cudaGetLastError();
    }
    LiftParams p;
    p.mel_arena = mel_arena; p.a_image = a_image; p.tiles = tiles; p.S = S;
    p.n_tiles = n_tiles; p.n_mt = M / LT_BM; p.n_mels = n_mels; p.kblocks = (n_mels + 31) / 32; p.ld = ld;
    p.power = power;
    p.groups = sm_count / p.n_mt;
    if (p.groups < 1) p.groups = 1;
    if (p.groups > n_tiles) p.groups = n_tiles;
    p.pinv_nyq = pinvT + M; p.pinv_ld = K;
    const int grid = p.n_mt * p.groups, sm = LiftSmem::total(p.kblocks);
    if (delog == 0) gl_lift_tc_kernel<0><<<grid, LT_THREADS, sm, s>>>(p);
    else if (delog == 1) gl_lift_tc_kernel<1><<<grid, LT_THREADS, sm, s>>>(p);
    else gl_lift_tc_kernel<2><<<grid, LT_THREADS, sm, s>>>(p);
    if (n_kernels) *n_kernels = 1;
    return cudaGetLastError();
}

}  // namespace xdtts
