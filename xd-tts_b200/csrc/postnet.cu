// Tacotron2 postnet for sm_100a: out = mel + Postnet(mel), Postnet = 5 x [Conv1d(k=5, pad=2) -> BatchNorm(eval)]
// with tanh after the first four (the ONNX session the reference runs at /root/reference
// src/tacotron2/mod.rs:344-357, loaded at :256-259; graph = NVIDIA Tacotron2 `Postnet`, SURVEY.md
// section 8 row a3 and appendix B).  BatchNorm is folded into the convolution on the host.
//
// Each layer is ONE implicit GEMM on the 5th-generation tensor cores:
//
//     D[r][co] = sum_{j<5} sum_{ci} X[r + j][ci] * W[j][co][ci]          (r = frame row, halo of 2)
//
// with frames as the UMMA M dimension (128 rows per tile = 128 TMEM lanes), output channels as N
// (up to 256 fp32 TMEM columns, two accumulator stages = all 512 columns) and (tap, ci) as K in
// 64-element steps.  Activations live in HBM time-major [row][channel] bf16, so a tap is the same
// TMA box shifted down by j rows -- the im2col matrix is never materialised.  Utterances are stacked
// along the row axis with two zero rows between them, which is the convolution's zero padding.
//
// Warp roles (256 threads, one CTA per SM, persistent over tiles):
//     warp 0   TMA producer   cp.async.bulk.tensor -> 4-stage smem ring (A 16 KB + B <= 32 KB per stage, SWIZZLE_128B)
//     warp 1   MMA issuer     one thread, tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), 4 per stage, tcgen05.commit
//     warp 2   TMEM allocator
//     warps 4-7 epilogue      tcgen05.ld -> + bias -> tanh -> bf16 (hi, lo) -> HBM; the last layer adds the
//                             residual and writes fp32 [80][T] (the reference's ndarray layout)
//
// Precision.  The reference computes in fp32 (ONNX Runtime CPU).  A single bf16 pass costs ~1e-2 on
// the ln-mel; the default therefore runs the split scheme x = hi + lo (both bf16):
// D = Ahi Bhi + Ahi Blo + Alo Bhi, three passes over K on the tensor cores into the same fp32
// accumulator ("bf16x3", error ~2^-17 per operand).  n_split = 1 is the plain bf16 GEMM.
// A CUDA-core fp32 kernel (pn_conv_f32_kernel) restates the layer for strict parity runs and as the
// on-device cross-check of the tensor path.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "postnet.h"

namespace xdtts {

// ------------------------------------------------------------------ PTX wrappers
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of TMEM -> 16 registers per thread (lane = row)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major tile whose rows are 128 B (64 bf16) and which was
// written by TMA with SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), LBO unused, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------ tensor-core layer kernel
struct PnSmem {
    // stage ring first (1024-byte aligned for the 128 B swizzle), then the bookkeeping
    static constexpr int A_BYTES = PN_BM * PN_BK * 2;            // 16 KB
    static constexpr int B_BYTES = PN_MAX_BN * PN_BK * 2;        // 32 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RING_BYTES = PN_STAGES * STAGE_BYTES;   // 192 KB
    static constexpr int BAR_OFF = RING_BYTES;                   // full[4] empty[4] tfull[2] tempty[2]
    static constexpr int TMEM_PTR_OFF = BAR_OFF + 8 * (2 * PN_STAGES + 4);
    static constexpr int BIAS_OFF = TMEM_PTR_OFF + 16;
    static constexpr int TOTAL = BIAS_OFF + 4 * PN_MAX_COUT + 1024;   // + slack for the manual 1024 B alignment
};

__global__ void __launch_bounds__(PN_THREADS, 1)
    pn_conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const PnLayer p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + PnSmem::BAR_OFF);
    uint64_t* empty = full + PN_STAGES;
    uint64_t* tfull = empty + PN_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_ptr = (uint32_t*)(smem + PnSmem::TMEM_PTR_OFF);
    float* bias_s = (float*)(smem + PnSmem::BIAS_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_total = p.m_tiles * p.n_tiles;
    const int k_blocks = p.n_split * p.taps * p.cin_chunks;

    for (int i = threadIdx.x; i < p.cout; i += PN_THREADS) bias_s[i] = p.bias[i];
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_hi) : "memory");
        if (p.n_split > 1) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b_lo) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PN_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // 512 columns: two 128 x 256 fp32 accumulator stages
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer (one thread)
        if (lane == 0) {
            const uint32_t stage_tx = (uint32_t)(PnSmem::A_BYTES + p.block_n * PN_BK * 2);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
                const int m = tile / p.n_tiles, n = tile % p.n_tiles;
                for (int sp = 0; sp < p.n_split; sp++) {
                    const CUtensorMap* ma = (sp == 2) ? &tm_a_lo : &tm_a_hi;
                    const CUtensorMap* mb = (sp == 1) ? &tm_b_lo : &tm_b_hi;
                    for (int j = 0; j < p.taps; j++) {
                        for (int c = 0; c < p.cin_chunks; c++) {
                            mbar_wait(&empty[stage], phase ^ 1u);
                            uint8_t* sa = smem + stage * PnSmem::STAGE_BYTES;
                            mbar_expect_tx(&full[stage], stage_tx);
                            // buffer row of tile row r is r + 2, so tap j reads buffer rows m*128 + j ...
                            tma_load_2d(sa, ma, &full[stage], c * PN_BK, m * PN_BM + j);
                            tma_load_2d(sa + PnSmem::A_BYTES, mb, &full[stage], c * PN_BK, j * p.cout + n * p.block_n);
                            if (++stage == PN_STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread issues and commits)
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = block_n, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(PN_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, it++) {
                const int as = it & 1;
                mbar_wait(&tempty[as], (uint32_t)((it >> 1) & 1) ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * PN_MAX_BN);
                for (int kb = 0; kb < k_blocks; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * PnSmem::STAGE_BYTES);
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + PnSmem::A_BYTES);
#pragma unroll
                    for (int k = 0; k < PN_BK / 16; k++)   // UMMA_K = 16 bf16 = 32 B along the swizzled row: +2 in the address field
                        tc_mma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    tc_commit(&empty[stage]);                       // frees the stage when these MMAs retire
                    if (++stage == PN_STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(&tfull[as]);                              // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM lane quarter q <-> tile rows 32 q .. 32 q + 31
        const int q = warp - 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, it++) {
            const int m = tile / p.n_tiles, n = tile % p.n_tiles;
            const int as = it & 1;
            const int row = m * PN_BM + q * 32 + lane;              // tile-space row; buffer row is row + 2
            const int t = p.row_t[row];                             // -1: gap / tail row
            const bool valid = t >= 0;
            // last layer: the residual of the whole row (up to PN_RES_MAX channels) is requested BEFORE waiting for the accumulator.
            // out and resid may be one buffer, so the compiler kept each load behind the previous store and the 80 loads of a row
            // were 80 serialised DRAM round trips (29 us per tile; the layer took 121 us for 16% of a 512 -> 512 layer's flops)
            constexpr int PN_RES_MAX = 96;
            long res_base = 0;
            int res_T = 0;
            float rs[PN_RES_MAX];
            auto load_res = [&](int c) -> float {
                const int co = n * p.block_n + c;
                return (valid && co < p.cout && c < p.block_n) ? __ldcg(p.resid + res_base + (long)co * res_T) : 0.f;
            };
            if (p.last) {
                if (valid) {
                    const int u = p.row_u[row];
                    res_T = p.utt_T[u];
                    res_base = (long)p.utt_foff[u] * p.cout + t;
                }
#pragma unroll
                for (int c = 0; c < PN_RES_MAX; c++) rs[c] = load_res(c);
            }
            mbar_wait(&tfull[as], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PN_MAX_BN);
            if (!p.last) {
                __nv_bfloat16* orow_hi = p.out_hi + (size_t)(row + 2) * p.out_ld + n * p.block_n;
                __nv_bfloat16* orow_lo = p.out_lo ? p.out_lo + (size_t)(row + 2) * p.out_ld + n * p.block_n : nullptr;
                for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(taddr + (uint32_t)c0, v);
                    tc_wait_ld();
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float x0 = __uint_as_float(v[2 * i]) + bias_s[n * p.block_n + c0 + 2 * i];
                        float x1 = __uint_as_float(v[2 * i + 1]) + bias_s[n * p.block_n + c0 + 2 * i + 1];
                        if (p.apply_tanh) { x0 = tanhf(x0); x1 = tanhf(x1); }
                        if (!valid) { x0 = 0.f; x1 = 0.f; }            // gap rows stay zero: they are the next layer's padding
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                        const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                        const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                        hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                    uint4* dh = reinterpret_cast<uint4*>(orow_hi + c0);
                    dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    if (orow_lo) {
                        uint4* dl = reinterpret_cast<uint4*>(orow_lo + c0);
                        dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
            } else {
                // last layer: out[co][t] = mel[co][t] + D + bias, fp32, the reference's [C, T] layout
#pragma unroll
                for (int c0 = 0; c0 < PN_MAX_BN; c0 += 16) {
                    if (c0 >= p.block_n) break;
                    uint32_t v[16];
                    tc_ld16(taddr + (uint32_t)c0, v);
                    tc_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int co = n * p.block_n + c0 + i;
                            const float r = c0 + i < PN_RES_MAX ? rs[c0 + i < PN_RES_MAX ? c0 + i : 0] : load_res(c0 + i);   // wider layers: the tail is loaded in place
                            if (co < p.cout) p.out_f32[res_base + (long)co * res_T] = r + (__uint_as_float(v[i]) + bias_s[co]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

cudaError_t pn_prepare() {
    return cudaFuncSetAttribute(pn_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PnSmem::TOTAL);
}

cudaError_t pn_launch_conv_tc(const CUtensorMap* maps4, const PnLayer& p, int sm_count, cudaStream_t s) {
    int grid = p.m_tiles * p.n_tiles;
    if (grid > sm_count) grid = sm_count;
    pn_conv_tc_kernel<<<grid, PN_THREADS, PnSmem::TOTAL, s>>>(maps4[0], maps4[1], maps4[2], maps4[3], p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ input staging
// mel arena (utterance u: row-major [C][T_u] fp32 at float offset foff[u]*C) -> time-major rows
// [row + 2][ld] as bf16 hi/lo (tensor path) and/or fp32 (CUDA-core path).  grid (ceil(maxT/32), B).
__global__ void __launch_bounds__(256) pn_stage_input_kernel(const float* __restrict__ mel, const int* __restrict__ utt_T,
                                                             const int* __restrict__ utt_foff, const int* __restrict__ utt_roff,
                                                             int C, int ld, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, float* __restrict__ f32) {
    __shared__ float tile[PN_MAX_CIN0][33];
    const int u = blockIdx.y, T = utt_T[u], t0 = blockIdx.x * 32;
    if (t0 >= T) return;
    const float* src = mel + (long)utt_foff[u] * C;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int c = ty; c < C; c += 8) tile[c][tx] = (t0 + tx < T) ? src[(long)c * T + t0 + tx] : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * C; i += 256) {
        const int tt = i / C, c = i % C;
        if (t0 + tt >= T) continue;
        const float x = tile[c][tt];
        const size_t o = (size_t)(utt_roff[u] + t0 + tt + 2) * ld + c;
        if (hi) {
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            hi[o] = h;
            if (lo) lo[o] = __float2bfloat16_rn(x - __bfloat162float(h));
        }
        if (f32) f32[o] = x;
    }
}

cudaError_t pn_launch_stage_input(const float* mel, const int* utt_T, const int* utt_foff, const int* utt_roff, int n_utt,
                                  int max_T, int C, int ld, __nv_bfloat16* hi, __nv_bfloat16* lo, float* f32, cudaStream_t s) {
    if (C > PN_MAX_CIN0) return cudaErrorInvalidValue;
    dim3 grid((max_T + 31) / 32, n_utt);
    pn_stage_input_kernel<<<grid, 256, 0, s>>>(mel, utt_T, utt_foff, utt_roff, C, ld, hi, lo, f32);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ CUDA-core fp32 layer (strict parity / cross-check)
// out[r][co] = act(b[co] + sum_{j,ci} x[r + j][ci] * w[j][ci][co]); 64 x 64 tile, 4 x 4 per thread, K step 16.
__global__ void __launch_bounds__(256) pn_conv_f32_kernel(const float* __restrict__ x, int ld_in, const float* __restrict__ w,
                                                          const PnLayer p, int cin, float* __restrict__ out, int ld_out) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
    for (int j = 0; j < p.taps; j++) {
        for (int k0 = 0; k0 < cin; k0 += 16) {
            for (int i = threadIdx.x; i < 64 * 16; i += 256) {
                const int rr = i >> 4, kk = i & 15;   // consecutive threads -> consecutive ci
                As[kk][rr] = (k0 + kk < cin) ? x[(size_t)(r0 + rr + j) * ld_in + k0 + kk] : 0.f;
            }
            for (int i = threadIdx.x; i < 16 * 64; i += 256) {
                const int kk = i >> 6, cc = i & 63;   // consecutive threads -> consecutive co
                Bs[kk][cc] = (k0 + kk < cin && c0 + cc < p.cout) ? w[((size_t)j * cin + k0 + kk) * p.cout + c0 + cc] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; kk++) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) acc[i][jj] = fmaf(a[i], b[jj], acc[i][jj]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = r0 + ty * 4 + i;
        const int t = p.row_t[row];
        const bool valid = t >= 0;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int co = c0 + tx * 4 + jj;
            if (co >= p.cout) continue;
            float v = acc[i][jj] + p.bias[co];
            if (!p.last) {
                if (p.apply_tanh) v = tanhf(v);
                out[(size_t)(row + 2) * ld_out + co] = valid ? v : 0.f;
            } else if (valid) {
                const int u = p.row_u[row];
                const long idx = (long)p.utt_foff[u] * p.cout + (long)co * p.utt_T[u] + t;
                p.out_f32[idx] = p.resid[idx] + v;
            }
        }
    }
}

cudaError_t pn_launch_conv_f32(const float* x, int ld_in, const float* w, const PnLayer& p, int cin, float* out, int ld_out,
                               cudaStream_t s) {
    dim3 grid(p.m_tiles * (PN_BM / 64), (p.cout + 63) / 64);
    pn_conv_f32_kernel<<<grid, 256, 0, s>>>(x, ld_in, w, p, cin, out, ld_out);
    return cudaGetLastError();
}

}  // namespace xdtts
