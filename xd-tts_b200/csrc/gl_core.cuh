// Griffin-Lim iteration core: the per-lane program of one warp that walks a run of
// consecutive STFT frames of one utterance.
//
// Replaces the loop body of griffin_lim::GriffinLim::infer (external crate
// griffin-lim 0.2.0 @ e6415314, called at /root/reference src/lib.rs:141, configured
// at src/tacotron2/mod.rs:453-456; algorithm per SURVEY.md section 3.4 / appendix B):
//
//     rebuilt = stft(istft(S * angles));  angles = unit(rebuilt - alpha * tprev)
//
// fused "STFT-first" so that ONE kernel launch is one iteration:
//
//     R_i  = STFT(y_{i-1})                       Hann window, real FFT of N = 128*R3
//     Y    = S * unit(R_i - alpha * R_{i-1})     momentum + projection onto |S|
//     y_i  = ISTFT(Y)                            inverse real FFT, window, overlap-add
//
// Mapping.  One warp owns one frame at a time; the N-point real FFT is an M = N/2
// point complex FFT on z[n] = x[2n] + i x[2n+1], factored M = 8 * 8 * R3 with all
// butterflies in registers (VPL = M/32 complex values per lane) and two
// shared-memory transposes per direction.  The last forward pass is arranged so
// that every lane holds both bins k and M-k of each of its values, hence the
// real-FFT split, the projection and the inverse split need no data exchange, and
// the inverse transform is the exact transpose of the forward one (same lane
// mapping, conjugated twiddles).  The windowed output of consecutive frames is
// overlap-added in registers (hop = N/4: a lane owns the same sample offsets in
// every hop block), so a finished hop block leaves the warp exactly once.
//
// Everything here is `__host__ __device__`: tests/emu/ runs the identical lane
// program on the CPU (lanes looped, warp syncs between phases) to check the index
// algebra without a GPU.  The product only ever calls it from gl_iter.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define XD_HD __host__ __device__ __forceinline__
#else
#define XD_HD inline
#endif

// 1: the frame's samples stay in place in L.raw and the window multiply is instantiated per rotation (phase_f1_rot);
// 0: L.raw is shifted by one hop per frame (phase_f1 + prefetch_next_block).  Measured on B200 (round 2):
// n_fft 1024 86.1 -> 85.8 us per launch with rotation, n_fft 2048 367.8 -> 375.7 us (four copies of a window
// multiply of 64 values cost more instruction-cache misses than the 56 moves they save), hence per geometry.
#ifndef XDTTS_GL_ROT
#define XDTTS_GL_ROT(R3) ((R3) != 16)
#endif
// keep the 7 pass-2 twiddles in registers for the whole run instead of re-reading them from shared memory twice per
// frame: n_fft 2048 383 -> 376 us, n_fft 1024 85.8 -> 84.0 us (round 2; in round 1, before the sample registers
// stopped moving, the extra 14 registers cost more than the loads at n_fft 1024)
#ifndef XDTTS_GL_TW2_REGS
#define XDTTS_GL_TW2_REGS(R3) 1
#endif

namespace xdtts {

// ------------------------------------------------------------------ geometry
template <int R3_>
struct Geo {
    static constexpr int R3 = R3_;            // radix of the last pass: 4, 8 or 16
    static constexpr int M = 64 * R3;         // complex FFT length
    static constexpr int N = 2 * M;           // n_fft
    static constexpr int H = N / 4;           // hop (the fused path needs hop == n_fft/4)
    static constexpr int VPL = 2 * R3;        // complex values per lane
    static constexpr int NB = R3 / 4;         // radix-8 butterflies per lane in passes 1 and 2
    static constexpr int LN3 = (R3 == 4) ? 2 : (R3 == 8 ? 3 : 4);
    static constexpr int S1 = 8 * R3 + (R3 == 4 ? 4 : 8);  // exchange-1 row stride (per k1), bank-conflict free
    static constexpr int S2 = 64 + 16 / R3;                // exchange-2 row stride (per n3), bank-conflict free
    static constexpr int EX1 = 8 * S1;        // exchange-1 size, complex elements
    static constexpr int EX2 = R3 * S2;       // exchange-2 size
    static constexpr int EXW = EX1 + EX2;     // per-warp exchange scratch
    // constant tables, float2 units.  Rows that a lane reads together are PAIRED into 16-byte entries ([pair][32 lanes]
    // of float4), so that they arrive with one 128-bit shared load: per butterfly i the pass-1 twiddles k1 = (1,2), (3,4),
    // (5,6) then k1 = 7 alone; the window rows n1 = (0,1), (2,3); the split twiddles j = (0,1), (2,3), ...
    static constexpr int TW1_OFF = 0;                     // W_M^{(lane+32 i) k1}: block i at 7*32*i = [3 pairs][32] float4 + [32] float2
    static constexpr int TW2_OFF = TW1_OFF + 7 * NB * 32; // W_{8 R3}^{(lane & (R3-1)) k2}, row k2-1
    static constexpr int RTW_OFF = TW2_OFF + 7 * 32;      // -i W_N^{k(lane,j)} / 2, row j
    static constexpr int WIN_OFF = RTW_OFF + R3 * 32;     // (w[s], w[s+1]), row i*4 + n1, n1 = 0..3 only:
                                                          // w[s + N/2] = 1 - w[s] (periodic Hann), so the rows n1+4 are
                                                          // applied as x - x w with one FFMA and never loaded
    static constexpr int TAB = WIN_OFF + (VPL / 2) * 32;
    static constexpr int REC_F = 3 * M + 4;   // floats per frame state record: R (2M) | S (M) | S_nyq + 3 pad
    static constexpr int REC = 4 * REC_F;     // bytes; a multiple of 16
    static constexpr bool TW2_REGS = XDTTS_GL_TW2_REGS(R3_);
    static constexpr bool ROT = XDTTS_GL_ROT(R3_);
};

// Exchange 1 holds, per k1 row, the 8 R3 values m = lane + 32 i of pass 1.  A lane's values of one butterfly pair
// (i even, i odd) are adjacent, so pass 1 stores and the last inverse pass loads them as ONE 128-bit access; pass 2
// finds its inputs n2 and n2 + 16/R3*... (m and m + 32) adjacent in the same way.  Conflict-free: a quarter warp
// touches 8 consecutive 16-byte units.
template <int R3>
XD_HD constexpr int ex1_pos(int m) {
    return (R3 == 4) ? m : 2 * (m & 31) + ((m >> 5) & 1) + 64 * (m >> 6);
}
XD_HD float2 mk2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
XD_HD float4 pack4(float2 a, float2 b) { float4 r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y; return r; }

// paired table rows (see Geo): `base` = float2 offset of the block, entry `pair` of this lane -> the two float2 rows
XD_HD void tab_pair(const float2* tab, int base, int pair, int lane, float2* a, float2* b) {
    const float4 q = *reinterpret_cast<const float4*>(&tab[base + 64 * pair + 2 * lane]);
    *a = mk2(q.x, q.y);
    *b = mk2(q.z, q.w);
}
// the seven pass-1 twiddles of butterfly i: tw[k1 - 1]
template <int R3>
XD_HD void load_tw1(const float2* tab, int i, int lane, float2* tw) {
    const int base = Geo<R3>::TW1_OFF + i * 7 * 32;
    tab_pair(tab, base, 0, lane, &tw[0], &tw[1]);
    tab_pair(tab, base, 1, lane, &tw[2], &tw[3]);
    tab_pair(tab, base, 2, lane, &tw[4], &tw[5]);
    tw[6] = tab[base + 6 * 32 + lane];
}

// pass 1 -> exchange 1: v[i*8 + k1] (already multiplied by its twiddle) goes to row k1, element lane + 32 i
template <int R3>
XD_HD void ex1_store_rows(const float2* v, int lane, float2* ex1) {
    typedef Geo<R3> G;
    if constexpr (R3 == 4) {
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) ex1[k1 * G::S1 + lane] = v[k1];
    } else {
#pragma unroll
        for (int ip = 0; ip < G::NB / 2; ip++)
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++)
                *reinterpret_cast<float4*>(&ex1[k1 * G::S1 + 2 * lane + 64 * ip]) = pack4(v[(2 * ip) * 8 + k1], v[(2 * ip + 1) * 8 + k1]);
    }
}
// exchange 1 -> last inverse pass: the transpose of ex1_store_rows
template <int R3>
XD_HD void ex1_load_rows(float2* v, int lane, const float2* ex1) {
    typedef Geo<R3> G;
    if constexpr (R3 == 4) {
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) v[k1] = ex1[k1 * G::S1 + lane];
    } else {
#pragma unroll
        for (int ip = 0; ip < G::NB / 2; ip++)
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) {
                const float4 q = *reinterpret_cast<const float4*>(&ex1[k1 * G::S1 + 2 * lane + 64 * ip]);
                v[(2 * ip) * 8 + k1] = mk2(q.x, q.y);
                v[(2 * ip + 1) * 8 + k1] = mk2(q.z, q.w);
            }
    }
}
// exchange 1 -> pass 2: butterfly i of a lane reads row k1, elements R3 n2 + n3, n2 = 0..7
template <int R3>
XD_HD void ex1_load_cols(float2* v8, int k1, int n3, const float2* ex1) {
    typedef Geo<R3> G;
    if constexpr (R3 == 4) {
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) v8[n2] = ex1[k1 * G::S1 + R3 * n2 + n3];
    } else {
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++)
            if ((((R3 * n2) >> 5) & 1) == 0) {
                const float4 q = *reinterpret_cast<const float4*>(&ex1[k1 * G::S1 + ex1_pos<R3>(R3 * n2) + 2 * n3]);
                v8[n2] = mk2(q.x, q.y);
                v8[n2 + 32 / R3] = mk2(q.z, q.w);
            }
    }
}
template <int R3>
XD_HD void ex1_store_cols(const float2* v8, int k1, int n3, float2* ex1) {
    typedef Geo<R3> G;
    if constexpr (R3 == 4) {
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) ex1[k1 * G::S1 + R3 * n2 + n3] = v8[n2];
    } else {
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++)
            if ((((R3 * n2) >> 5) & 1) == 0)
                *reinterpret_cast<float4*>(&ex1[k1 * G::S1 + ex1_pos<R3>(R3 * n2) + 2 * n3]) = pack4(v8[n2], v8[n2 + 32 / R3]);
    }
}

// bin held in pair slot j of a lane after the last forward pass (its partner is M - k)
template <int R3>
XD_HD int kslot(int lane, int j) {
    if (lane) return lane + 64 * j;
    return (j < R3 / 2) ? 64 * j : 32 + 64 * j;   // lane 0: column 0, then column 32 entered from its upper half (lane0_fix)
}

// ------------------------------------------------------------------ complex helpers
// Complex add / subtract.  sm_100 has packed fp32x2 instructions (FADD2 / FFMA2 on a 64-bit register
// pair; a - b = fma(b, -1, a) is exact).  Measured on B200 they cut the FP instruction count of this
// kernel by 17% but make it SLOWER (cfg2 96.1 -> 106.8 us, cfg5 459 -> 501 us): the packed forms issue
// at a lower rate than two scalar FADDs on the two FP32 sub-pipes.  Kept behind a macro, default off.
#ifndef XDTTS_GL_PACKED
#define XDTTS_GL_PACKED 0
#endif
XD_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && XDTTS_GL_PACKED
    return __fadd2_rn(a, b);
#else
    return mk2(a.x + b.x, a.y + b.y);
#endif
}
XD_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && XDTTS_GL_PACKED
    return __ffma2_rn(b, mk2(-1.f, -1.f), a);
#else
    return mk2(a.x - b.x, a.y - b.y);
#endif
}
XD_HD float2 cmul(float2 a, float2 b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
XD_HD float2 cmulc(float2 a, float2 b) { return mk2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)

// cos(2 pi k / 16), sin(2 pi k / 16) as literals (no constexpr libm on device)
XD_HD constexpr float cos16(int k) {
    return k == 0 ? 1.f : k == 1 ? 0.92387953251128674f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.38268343236508977f
         : k == 4 ? 0.f : k == 5 ? -0.38268343236508977f : k == 6 ? -0.70710678118654752f : k == 7 ? -0.92387953251128674f : -1.f;
}
XD_HD constexpr float sin16(int k) {
    return k == 0 ? 0.f : k == 1 ? 0.38268343236508977f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.92387953251128674f
         : k == 4 ? 1.f : k == 5 ? 0.92387953251128674f : k == 6 ? 0.70710678118654752f : k == 7 ? 0.38268343236508977f : 0.f;
}

// z * W_R^k (forward, e^{-2 pi i k / R}) or z * conj(W_R^k) (inverse); k < R/2, R <= 16
template <int R, int K, bool INV>
XD_HD float2 mulw(float2 z) {
    constexpr float r = 0.70710678118654752f;
    if constexpr (K == 0) {
        return z;
    } else if constexpr (4 * K == R) {
        return INV ? mk2(-z.y, z.x) : mk2(z.y, -z.x);
    } else if constexpr (8 * K == R) {
        return INV ? mk2(r * (z.x - z.y), r * (z.x + z.y)) : mk2(r * (z.x + z.y), r * (z.y - z.x));
    } else if constexpr (8 * K == 3 * R) {
        return INV ? mk2(-r * (z.x + z.y), r * (z.x - z.y)) : mk2(r * (z.y - z.x), -r * (z.x + z.y));
    } else {
        constexpr int k16 = K * (16 / R);
        constexpr float c = cos16(k16), s = sin16(k16);
        return INV ? mk2(z.x * c - z.y * s, z.x * s + z.y * c) : mk2(z.x * c + z.y * s, z.y * c - z.x * s);
    }
}

// in-register DFT of R points, natural order in and out (decimation in time, fully unrolled)
template <int R, bool INV>
struct Dft {
    template <int K>
    static XD_HD void combine(float2* x, const float2* e, const float2* o) {
        if constexpr (K < R / 2) {
            const float2 t = mulw<R, K, INV>(o[K]);
            x[K] = cadd(e[K], t);
            x[K + R / 2] = csub(e[K], t);
            combine<K + 1>(x, e, o);
        }
    }
    static XD_HD void run(float2* x) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; k++) { e[k] = x[2 * k]; o[k] = x[2 * k + 1]; }
        Dft<R / 2, INV>::run(e);
        Dft<R / 2, INV>::run(o);
        combine<0>(x, e, o);
    }
};
template <bool INV>
struct Dft<2, INV> {
    static XD_HD void run(float2* x) {
        const float2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};

// ------------------------------------------------------------------ parameters
enum { GL_MODE_INIT = 0, GL_MODE_FIRST = 1, GL_MODE_MID = 2 };  // + flags below
enum { GL_PAD_REFLECT = 0, GL_PAD_CONSTANT = 1 };

struct GlRun {   // one warp's work: frames [ta, tb) of utterance utt
    int utt, ta, tb, pad;
};

struct GlParams {
    int n_runs;
    const GlRun* runs;
    const int* utt_T;        // frames per utterance
    const int* utt_foff;     // first frame row of the utterance in S / R; its samples start at foff * H
    const float* y_in;       // waveform of the previous iteration
    float* y_out;            // waveform of this iteration
    // per-frame state record, `Geo::REC` bytes, contiguous so that ONE bulk copy stages a frame:
    //   [R: M float2, rebuilt spectrum, updated in place; slot 0 = (Re R[0], Re R[M])][S: M floats, |S| of bins 0..M-1]
    //   [S_nyq: |S| of bin M][3 floats of padding]
    float* state;            // record of frame f at state + f * Geo::REC_F floats
    float* halo;             // [n_runs][2][3*H] partial overlap-add sums at run boundaries
    unsigned* flags;         // [n_runs] arrival counters of the boundary to the right of each run
    unsigned* amax;          // [n_utt] max |y| as float bits (last iteration only)
    const float* edge_scale; // [2][H] 1 / window-sum-square of the first / last hop block
    const float2* tables;    // Geo::TAB constants
    const float* turns;      // INIT: [frames][M + 1] initial phase in turns (bin M last), or null -> hashed from seed
    // INIT, seeded phase: both live in device memory so that a captured CUDA graph sees the values of the call
    // that replays it (the seed advances per call; a multi-GPU pool numbers utterances globally)
    const unsigned long long* seed;   // seed of the counter-based phase generator
    const int* utt_seed_id;           // [n_utt] stream index of each utterance (its position in the caller's batch)
    float* ybuf[2];          // persistent kernel: the two waveform buffers (iteration i reads [(i-1)&1], writes [i&1])
    unsigned* done;          // persistent kernel: [n_runs] number of iterations each run has completed
    int n_iter;              // persistent kernel: iterations after the initial inverse transform
    float alpha;             // momentum / (1 + momentum)
    float inv_n;             // 1 / n_fft
    int pad_mode;
};

// counter-based uniform in [0,1): top 24 bits of splitmix64 (bit-identical to oracle phase_turns)
XD_HD float phase_turn(unsigned long long seed, int utt, int k_bins, int t, int k) {
    const unsigned long long G = 0x9E3779B97F4A7C15ull;
    unsigned long long z = seed + G * (unsigned long long)(utt + 1);
    z += G * ((unsigned long long)t * (unsigned long long)k_bins + (unsigned long long)k + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * 5.9604644775390625e-8f;
}

XD_HD void sincos_turns(float u, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincospif(2.0f * u, s, c);
#else
    const double th = 6.283185307179586476925286766559 * (double)u;
    *s = (float)sin(th);
    *c = (float)cos(th);
#endif
}

// 1 / sqrt(m2) for the unit-modulus projection u / |u| (librosa: angles / (|angles| + tiny)).  One MUFU.RSQ on the
// device: rsqrtf() wraps the same instruction in a denormal-rescaling sequence (FSETP + 2 predicated FMUL per call,
// 16 calls per lane and frame), and the m2 > 0 guard cost another FSETP + FSEL.  The floor keeps u = 0 -> 0 (0 * 1e15)
// and only changes values with |u| < 1e-15, far below the rounding noise of the transform.
XD_HD float rsqrt_pos(float m2) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(m2, 1e-30f)));
    return r;
#else
    return 1.0f / sqrtf(fmaxf(m2, 1e-30f));
#endif
}

// streaming global accesses (state that is touched once per launch must not evict the waveform from L1)
template <typename T>
XD_HD T ld_stream(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T>
XD_HD void st_stream(T* p, T v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}
template <typename T>
XD_HD T ld_l2(const T* p) {   // bypass L1: data written by another SM during this launch
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}

// ------------------------------------------------------------------ per-lane state
template <int R3>
struct Lane {
    float2 v[Geo<R3>::VPL];          // FFT working set
    float2 raw[Geo<R3>::VPL];        // un-windowed samples of the current frame; three quarters carry over to the next
    float2 acc[3][2 * Geo<R3>::NB];  // overlap-add sums of the three unfinished hop blocks
    float2 ynew[2 * Geo<R3>::NB];    // newest hop block of the NEXT frame, fetched one frame ahead (unused when Geo::ROT)
    float2 tw2[7];                   // pass-2 twiddles W_{8 R3}^{(lane & (R3-1)) k2}, k2 = 1..7 (only when Geo::TW2_REGS)
    float amax;
};

XD_HD int reflect_index(int j, int len) {
    if (j < 0) j = -j;
    if (j >= len) j = 2 * (len - 1) - j;
    return j;
}

// ------------------------------------------------------------------ state staging
// S and the previous rebuilt spectrum of ONE frame are contiguous rows (4M and 8M bytes).  The warp
// stages them in shared memory with two bulk asynchronous copies (cp.async.bulk, the TMA engine)
// issued by lane 0 one frame ahead and signalled on the warp's mbarrier, so no lane holds state in
// registers while the FFT passes run and the DRAM latency is off the warp's critical path.
// The CPU emulator performs the same copy synchronously.
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_addr_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"(smem_addr_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
                 : "memory");
}
#endif

// Issue the copy of `frame`'s state record into the warp's staging area (call after every lane has finished
// reading the previous contents: the caller puts a warp sync in between).  The record is contiguous in HBM and in
// the staging area -- [R | S | S_nyq] -- so a frame is ONE bulk copy (INIT / FIRST, which have no previous
// spectrum, copy the [S | S_nyq] tail only): each cp.async.bulk costs ~20 instructions of uniform-register set-up
// around the one that does the work, and three of them per frame were 4% of the kernel's issue slots.
template <int R3, int MODE>
XD_HD void stage_issue(int lane, const GlParams& p, long frame, float* stg, unsigned long long* bar) {
    typedef Geo<R3> G;
    if (lane == 0) {
        const float* rec = p.state + frame * G::REC_F;
#if defined(__CUDA_ARCH__)
        if (MODE == GL_MODE_MID) {
            mbar_expect_tx(bar, (unsigned)G::REC);
            bulk_g2s(stg, rec, (unsigned)G::REC, bar);
        } else {
            mbar_expect_tx(bar, (unsigned)(G::M * 4 + 16));
            bulk_g2s(stg + 2 * G::M, rec + 2 * G::M, (unsigned)(G::M * 4 + 16), bar);
        }
#else
        (void)bar;
        for (int k = (MODE == GL_MODE_MID ? 0 : 2 * G::M); k < G::REC_F; k++) stg[k] = rec[k];
#endif
    }
}

XD_HD void stage_wait(unsigned long long* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
    mbar_wait(bar, parity);
#else
    (void)bar;
    (void)parity;
#endif
}

XD_HD bool frame_is_edge(int t, int T) { return (t < 2) || (t >= T - 2); }

// quarter qi (one hop block, H samples) of frame t of the padded previous waveform -> q[2 NB]
// element (n1 & 1) * NB + i  <->  sample 16 R3 n1 + 2 (lane + 32 i), n1 = 2 qi + (n1 & 1)
// L2LOAD: read through L2 (ld.global.cg) -- required when the waveform was written by other SMs during
// the same launch (persistent kernel), where an L1 line of the same buffer may be stale.
template <typename T>
XD_HD T ld_y(const T* p, bool l2) {
#if defined(__CUDA_ARCH__)
    return l2 ? __ldcg(p) : *p;
#else
    (void)l2;
    return *p;
#endif
}

template <int R3, bool L2LOAD = false>
XD_HD void load_quarter(float2* q, int lane, const float* y, int T, int t, int qi, int pad_mode) {
    typedef Geo<R3> G;
    const int g = t - 2 + qi;                  // hop block of the trimmed signal
    const int base = g * G::H;
    if (g >= 0 && g <= T - 2) {                // warp-uniform: entirely inside the signal
        const float* yb = y + base + 2 * lane;
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++)
            q[e] = ld_y(reinterpret_cast<const float2*>(yb + 16 * R3 * (e / G::NB) + 64 * (e % G::NB)), L2LOAD);
    } else {
        const int len = G::H * (T - 1);
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++) {
            const int j0 = base + 16 * R3 * (e / G::NB) + 64 * (e % G::NB) + 2 * lane, j1 = j0 + 1;
            float2 x;
            if (pad_mode == GL_PAD_REFLECT) {
                x.x = ld_y(y + reflect_index(j0, len), L2LOAD);
                x.y = ld_y(y + reflect_index(j1, len), L2LOAD);
            } else {
                x.x = (j0 >= 0 && j0 < len) ? ld_y(y + j0, L2LOAD) : 0.f;
                x.y = (j1 >= 0 && j1 < len) ? ld_y(y + j1, L2LOAD) : 0.f;
            }
            q[e] = x;
        }
    }
}

// F1: frame t of the previous waveform -> window -> pass 1 (radix 8 over n1) -> exchange 1.
// The frame's samples live in L.raw: a frame shares three of its four hop blocks with its
// predecessor, so only the newest block comes from memory (first: the run's first frame, which
// loads all four; have_pref: L.ynew already holds the newest block, fetched during the previous frame).
template <int R3, bool L2LOAD = false>
XD_HD void phase_f1(Lane<R3>& L, int lane, const float* y, int T, int t, int pad_mode, bool first, bool have_pref,
                    const float2* tab, float2* ex1) {
    typedef Geo<R3> G;
    if (first) {
#pragma unroll
        for (int qi = 0; qi < 4; qi++) {
            float2 q[2 * G::NB];
            load_quarter<R3, L2LOAD>(q, lane, y, T, t, qi, pad_mode);
#pragma unroll
            for (int e = 0; e < 2 * G::NB; e++) L.raw[(e % G::NB) * 8 + 2 * qi + e / G::NB] = q[e];
        }
    } else {
#pragma unroll
        for (int i = 0; i < G::NB; i++)
#pragma unroll
            for (int n1 = 0; n1 < 6; n1++) L.raw[i * 8 + n1] = L.raw[i * 8 + n1 + 2];
        float2 q[2 * G::NB];
        if (have_pref) {
#pragma unroll
            for (int e = 0; e < 2 * G::NB; e++) q[e] = L.ynew[e];
        } else {
            load_quarter<R3, L2LOAD>(q, lane, y, T, t, 3, pad_mode);
        }
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++) L.raw[(e % G::NB) * 8 + 6 + e / G::NB] = q[e];
    }
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
#pragma unroll
        for (int np = 0; np < 2; np++) {
            float2 wp[2];
            tab_pair(tab, G::WIN_OFF + i * 4 * 32, np, lane, &wp[0], &wp[1]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n1 = 2 * np + h;
                const float2 w = wp[h];
                const float2 lo = L.raw[i * 8 + n1], hi = L.raw[i * 8 + n1 + 4];
                L.v[i * 8 + n1] = mk2(lo.x * w.x, lo.y * w.y);
                L.v[i * 8 + n1 + 4] = mk2(fmaf(-hi.x, w.x, hi.x), fmaf(-hi.y, w.y, hi.y));   // x (1 - w)
            }
        }
        Dft<8, false>::run(&L.v[i * 8]);
#pragma unroll
        float2 tw[7];
        load_tw1<R3>(tab, i, lane, tw);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) L.v[i * 8 + k1] = cmul(L.v[i * 8 + k1], tw[k1 - 1]);
    }
    ex1_store_rows<R3>(L.v, lane, ex1);
}

// ---- F1 with ROTATING sample registers.  Shifting L.raw by one hop costs 6 NB register moves per frame (plus
// 2 NB to bring in the prefetched block); instead the registers stay put and the frame-relative position n1
// lives in register (n1 + 2 rho) & 7, rho = (t - ta) & 3.  Register indices must be static, so the window
// multiply is instantiated four times and selected by a warp-uniform switch.  The group that holds the oldest
// hop block is dead after the multiply -- the newest block of frame t+1 is loaded straight into it (no ynew).
template <int R3, int RHO, bool L2LOAD>
XD_HD void f1_window_rot(Lane<R3>& L, int lane, const float2* tab, const float* y, int t, bool insert, const float2* q,
                         bool fetch_next) {
    typedef Geo<R3> G;
    constexpr int P0 = (2 * RHO) & 7;   // register slot of frame-relative n1 = 0
    if (insert) {                       // newest hop block of THIS frame was not prefetched (edge frame / first frame)
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++) L.raw[(e % G::NB) * 8 + ((6 + P0) & 7) + e / G::NB] = q[e];
    }
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
#pragma unroll
        for (int np = 0; np < 2; np++) {
            float2 wp[2];
            tab_pair(tab, G::WIN_OFF + i * 4 * 32, np, lane, &wp[0], &wp[1]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n1 = 2 * np + h;
                const float2 w = wp[h];
                const float2 lo = L.raw[i * 8 + ((n1 + P0) & 7)], hi = L.raw[i * 8 + ((n1 + 4 + P0) & 7)];
                L.v[i * 8 + n1] = mk2(lo.x * w.x, lo.y * w.y);
                L.v[i * 8 + n1 + 4] = mk2(fmaf(-hi.x, w.x, hi.x), fmaf(-hi.y, w.y, hi.y));   // x (1 - w)
            }
        }
    }
    if (fetch_next) {   // hop block t+2 of the signal = newest block of frame t+1, entirely inside the signal
        const float* yb = y + (long)t * G::H + 2 * G::H + 2 * lane;
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++)
            L.raw[(e % G::NB) * 8 + P0 + e / G::NB] =
                ld_y(reinterpret_cast<const float2*>(yb + 16 * R3 * (e / G::NB) + 64 * (e % G::NB)), L2LOAD);
    }
}

template <int R3, bool L2LOAD = false>
XD_HD void phase_f1_rot(Lane<R3>& L, int lane, const float* y, int T, int t, int pad_mode, bool first, bool have_pref,
                        bool fetch_next, int rho, const float2* tab, float2* ex1) {
    typedef Geo<R3> G;
    float2 q[2 * G::NB];
    const bool insert = first || !have_pref;
    if (first) {   // rho == 0
#pragma unroll
        for (int qi = 0; qi < 3; qi++) {
            float2 q3[2 * G::NB];
            load_quarter<R3, L2LOAD>(q3, lane, y, T, t, qi, pad_mode);
#pragma unroll
            for (int e = 0; e < 2 * G::NB; e++) L.raw[(e % G::NB) * 8 + 2 * qi + e / G::NB] = q3[e];
        }
    }
    if (insert) load_quarter<R3, L2LOAD>(q, lane, y, T, t, 3, pad_mode);
    switch (rho) {
        case 0: f1_window_rot<R3, 0, L2LOAD>(L, lane, tab, y, t, insert, q, fetch_next); break;
        case 1: f1_window_rot<R3, 1, L2LOAD>(L, lane, tab, y, t, insert, q, fetch_next); break;
        case 2: f1_window_rot<R3, 2, L2LOAD>(L, lane, tab, y, t, insert, q, fetch_next); break;
        default: f1_window_rot<R3, 3, L2LOAD>(L, lane, tab, y, t, insert, q, fetch_next); break;
    }
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        Dft<8, false>::run(&L.v[i * 8]);
#pragma unroll
        float2 tw[7];
        load_tw1<R3>(tab, i, lane, tw);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) L.v[i * 8 + k1] = cmul(L.v[i * 8 + k1], tw[k1 - 1]);
    }
    ex1_store_rows<R3>(L.v, lane, ex1);
}

// Fetch the newest hop block of frame t+1 (inside the signal by construction: fast path) one frame
// ahead.  Issued AFTER the warp sync that follows F1: the loads then share no scoreboard wait with
// the window multiply of frame t, which consumes the block fetched a frame earlier.
template <int R3, bool L2LOAD = false>
XD_HD void prefetch_next_block(Lane<R3>& L, int lane, const float* y, int T, int t, int pad_mode) {
    load_quarter<R3, L2LOAD>(L.ynew, lane, y, T, t + 1, 3, pad_mode);
}

// F2: pass 2 (radix 8 over n2) -> exchange 2.  Split into its read half and its write half so that
// the two exchange buffers may share storage (a warp sync between the halves then orders them).
template <int R3>
XD_HD void phase_f2_load(Lane<R3>& L, int lane, const float2* ex1) {
    typedef Geo<R3> G;
    const int n3 = lane & (R3 - 1);
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        const int k1 = (lane >> G::LN3) + (32 / R3) * i;
        ex1_load_cols<R3>(&L.v[i * 8], k1, n3, ex1);
    }
}
template <int R3>
XD_HD void phase_f2_store(Lane<R3>& L, int lane, const float2* tab, float2* ex2) {
    typedef Geo<R3> G;
    const int n3 = lane & (R3 - 1);
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        const int k1 = (lane >> G::LN3) + (32 / R3) * i;
        Dft<8, false>::run(&L.v[i * 8]);
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            float2 z = L.v[i * 8 + k2];
            if (k2) z = cmul(z, G::TW2_REGS ? L.tw2[k2 - 1] : tab[G::TW2_OFF + (k2 - 1) * 32 + lane]);
            ex2[n3 * G::S2 + k1 + 8 * k2] = z;
        }
    }
}

// Lane 0 holds the two self-paired columns (k mod 64 == 0 and == 32): column 0 pairs A[j] with A[R3-j], column 32
// pairs B[j] with B[R3-1-j].  Every other lane pairs A[j] with B[R3-1-j].  Lane 0 is brought to that form with the
// FEWEST register moves: the lower half of A and the lower half of B stay where they are -- slots j < R3/2 keep
// a = A[j] and only need their partner in the b register; slots j >= R3/2 keep b = B[R3-1-j] and process the pair
// from the other side (a = B[j], the partner of B[R3-1-j]; kslot() names that bin).  So only the upper halves trade
// places: R3 selects per direction instead of the 3/2 R3 moves of a full re-sort.
template <int R3>
XD_HD void lane0_fix(float2* v, bool l0) {
    float2 Au[R3 / 2], Bu[R3 / 2];   // upper halves
#pragma unroll
    for (int i = 0; i < R3 / 2; i++) { Au[i] = v[R3 / 2 + i]; Bu[i] = v[R3 + R3 / 2 + i]; }
#pragma unroll
    for (int i = 0; i < R3 / 2; i++)
        if (l0) v[R3 / 2 + i] = Bu[i];                            // A'[j] = B[j], j >= R3/2
#pragma unroll
    for (int i = 0; i < R3 / 2 - 1; i++)
        if (l0) v[R3 + R3 / 2 + i] = Au[i + 1];                   // B'[m] = A[m+1], m = R3/2 .. R3-2
    if (l0) v[R3 + R3 - 1] = Au[0];                               // B'[R3-1] = A[R3/2]
}
template <int R3>
XD_HD void lane0_unfix(float2* v, bool l0) {
    float2 Au[R3 / 2], Bu[R3 / 2];   // upper halves of the primed arrays
#pragma unroll
    for (int i = 0; i < R3 / 2; i++) { Au[i] = v[R3 / 2 + i]; Bu[i] = v[R3 + R3 / 2 + i]; }
#pragma unroll
    for (int i = 0; i < R3 / 2; i++)
        if (l0) v[R3 + R3 / 2 + i] = Au[i];                       // B[j] = A'[j], j >= R3/2
#pragma unroll
    for (int i = 0; i < R3 / 2 - 1; i++)
        if (l0) v[R3 / 2 + i + 1] = Bu[i];                        // A[m+1] = B'[m]
    if (l0) v[R3 / 2] = Bu[R3 / 2 - 1];                           // A[R3/2] = B'[R3-1]
}

// F3: pass 3 (radix R3 over n3) -> real-FFT split -> momentum + projection -> inverse split
//     -> inverse pass 1 -> exchange 2 (in place)
template <int R3, int MODE, bool STORE_R>
XD_HD void phase_f3(Lane<R3>& L, int lane, const GlParams& p, int utt, int T, int t, long frame, const float2* tab,
                    float2* ex2, const float* stg) {
    typedef Geo<R3> G;
    const bool l0 = (lane == 0);
    const int qA = lane, qB = lane ? 64 - lane : 32;
    float2* A = &L.v[0];
    float2* B = &L.v[R3];
    float2* Rrow = reinterpret_cast<float2*>(p.state + frame * G::REC_F);
    const float2* r_stg = reinterpret_cast<const float2*>(stg);   // staged record: [R | S | S_nyq]
    const float* s_stg = stg + 2 * G::M;

    // The 1/N of the inverse transform is NOT applied here: Y = S unit(u) goes through the unnormalised inverse FFT
    // and the factor is folded into the 1/window-sum-square scale of the store (store_block), which every sample
    // passes exactly once.
    if (MODE != GL_MODE_INIT) {
#pragma unroll
        for (int n3 = 0; n3 < R3; n3++) {
            A[n3] = ex2[n3 * G::S2 + qA];
            B[n3] = ex2[n3 * G::S2 + qB];
        }
        Dft<R3, false>::run(A);
        Dft<R3, false>::run(B);
        lane0_fix<R3>(L.v, l0);
    }

    float2 Cpair[2];
#pragma unroll
    for (int j = 0; j < R3; j++) {
        const int ka = kslot<R3>(lane, j);
        const int kb = (l0 && j == 0) ? G::M / 2 : G::M - ka;
        if ((j & 1) == 0) tab_pair(tab, G::RTW_OFF, j >> 1, lane, &Cpair[0], &Cpair[1]);
        const float2 C = Cpair[j & 1];   // -i W_N^k / 2
        const float sa_j = s_stg[ka], sb_j = s_stg[kb];      // staged |S| at bins k and M-k
        float2 Ya, Yb;
        if (MODE != GL_MODE_INIT) {
            const float2 a = A[j], b = B[R3 - 1 - j];
            // X[k] = E + W^k O,  X[M-k] = conj(E - W^k O),  E = (a + conj b)/2,  O = (a - conj b)/(2i)
            const float2 s = mk2(a.x + b.x, a.y - b.y);
            const float2 d = mk2(a.x - b.x, a.y + b.y);
            const float2 tt = cmul(C, d);
            float2 Ra = mk2(0.5f * s.x + tt.x, 0.5f * s.y + tt.y);
            float2 Rb = mk2(0.5f * s.x - tt.x, tt.y - 0.5f * s.y);
            if (j == 0 && l0) {   // bins 0 and M are real and share one slot; bin M/2 pairs with itself
                Ra = mk2(a.x + a.y, a.x - a.y);
                Rb = mk2(b.x, -b.y);
            }
            float2 ua = Ra, ub = Rb;
            if (MODE == GL_MODE_MID) {
                const float2 pa_j = r_stg[ka], pb_j = r_stg[kb];   // staged previous rebuilt spectrum
                ua = mk2(Ra.x - p.alpha * pa_j.x, Ra.y - p.alpha * pa_j.y);
                ub = mk2(Rb.x - p.alpha * pb_j.x, Rb.y - p.alpha * pb_j.y);
            }
            if (STORE_R) {
                st_stream(Rrow + ka, Ra);
                st_stream(Rrow + kb, Rb);
            }
            const float ga = sa_j * rsqrt_pos(ua.x * ua.x + ua.y * ua.y);
            const float gb = sb_j * rsqrt_pos(ub.x * ub.x + ub.y * ub.y);
            Ya = mk2(ua.x * ga, ua.y * ga);
            Yb = mk2(ub.x * gb, ub.y * gb);
            if (j == 0 && l0) {
                const float y0 = ua.x > 0.f ? 1.f : (ua.x < 0.f ? -1.f : 0.f);
                const float ym = ua.y > 0.f ? 1.f : (ua.y < 0.f ? -1.f : 0.f);
                Ya = mk2(y0 * sa_j, ym * s_stg[G::M]);
            }
        } else {
            float ta_, tb_, tn_ = 0.f;
            if (p.turns) {
                ta_ = p.turns[frame * (G::M + 1) + ka];
                tb_ = p.turns[frame * (G::M + 1) + kb];
                if (j == 0 && l0) tn_ = p.turns[frame * (G::M + 1) + G::M];
            } else {
                const unsigned long long seed = *p.seed;
                const int sid = p.utt_seed_id[utt];
                ta_ = phase_turn(seed, sid, G::M + 1, t, ka);
                tb_ = phase_turn(seed, sid, G::M + 1, t, kb);
                if (j == 0 && l0) tn_ = phase_turn(seed, sid, G::M + 1, t, G::M);
            }
            float sn, cs;
            sincos_turns(ta_, &sn, &cs);
            Ya = mk2(sa_j * cs, sa_j * sn);
            sincos_turns(tb_, &sn, &cs);
            Yb = mk2(sb_j * cs, sb_j * sn);
            if (j == 0 && l0) {   // the inverse real FFT ignores the imaginary part of bins 0 and M
                sincos_turns(tn_, &sn, &cs);
                Ya = mk2(Ya.x, s_stg[G::M] * cs);
            }
        }
        // Z[k] = E' + i O',  Z[M-k] = conj(E' - i O'),  E' = Ya + conj Yb,  O' = conj(W^k) (Ya - conj Yb)
        const float2 s2 = mk2(Ya.x + Yb.x, Ya.y - Yb.y);
        const float2 d2 = mk2(Ya.x - Yb.x, Ya.y + Yb.y);
        const float2 t2 = cmulc(d2, C);                    // d2 * i conj(W_N^k) / 2  (C = -i W_N^k / 2; the 2 is exact in the FFMAs below)
        float2 Za = mk2(fmaf(2.f, t2.x, s2.x), fmaf(2.f, t2.y, s2.y));
        float2 Zb = mk2(fmaf(-2.f, t2.x, s2.x), fmaf(2.f, t2.y, -s2.y));
        if (j == 0 && l0) {
            Za = mk2(Ya.x + Ya.y, Ya.x - Ya.y);
            Zb = mk2(2.f * Yb.x, -2.f * Yb.y);
        }
        A[j] = Za;
        B[R3 - 1 - j] = Zb;
    }
    lane0_unfix<R3>(L.v, l0);
    Dft<R3, true>::run(A);
    Dft<R3, true>::run(B);
#pragma unroll
    for (int n3 = 0; n3 < R3; n3++) {
        ex2[n3 * G::S2 + qA] = A[n3];
        ex2[n3 * G::S2 + qB] = B[n3];
    }
}

// F4: inverse pass 2 (radix 8 over k2) -> exchange 1 (read half / write half, as F2)
template <int R3>
XD_HD void phase_f4_load(Lane<R3>& L, int lane, const float2* tab, const float2* ex2) {
    typedef Geo<R3> G;
    const int n3 = lane & (R3 - 1);
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        const int k1 = (lane >> G::LN3) + (32 / R3) * i;
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            float2 z = ex2[n3 * G::S2 + k1 + 8 * k2];
            if (k2) z = cmulc(z, G::TW2_REGS ? L.tw2[k2 - 1] : tab[G::TW2_OFF + (k2 - 1) * 32 + lane]);
            L.v[i * 8 + k2] = z;
        }
    }
}
template <int R3>
XD_HD void phase_f4_store(Lane<R3>& L, int lane, float2* ex1) {
    typedef Geo<R3> G;
    const int n3 = lane & (R3 - 1);
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        const int k1 = (lane >> G::LN3) + (32 / R3) * i;
        Dft<8, true>::run(&L.v[i * 8]);
        ex1_store_cols<R3>(&L.v[i * 8], k1, n3, ex1);
    }
}

// F5a: inverse pass 3 (radix 8 over k1) -> window; leaves frame samples in L.v
template <int R3>
XD_HD void phase_f5(Lane<R3>& L, int lane, const float2* tab, const float2* ex1) {
    typedef Geo<R3> G;
    ex1_load_rows<R3>(L.v, lane, ex1);
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
#pragma unroll
        float2 tw[7];
        load_tw1<R3>(tab, i, lane, tw);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) L.v[i * 8 + k1] = cmulc(L.v[i * 8 + k1], tw[k1 - 1]);
    }
#pragma unroll
    for (int i = 0; i < G::NB; i++) {
        Dft<8, true>::run(&L.v[i * 8]);
#pragma unroll
        for (int np = 0; np < 2; np++) {
            float2 wp[2];
            tab_pair(tab, G::WIN_OFF + i * 4 * 32, np, lane, &wp[0], &wp[1]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n1 = 2 * np + h;
                const float2 w = wp[h];
                const float2 lo = L.v[i * 8 + n1], hi = L.v[i * 8 + n1 + 4];
                L.v[i * 8 + n1] = mk2(lo.x * w.x, lo.y * w.y);
                L.v[i * 8 + n1 + 4] = mk2(fmaf(-hi.x, w.x, hi.x), fmaf(-hi.y, w.y, hi.y));
            }
        }
    }
}

// element e of a hop block held by a lane: e = p * NB + i  <->  sample offset p*16*R3 + 64 i + 2 lane
template <int R3>
XD_HD int blk_off(int lane, int e) {
    return (e / Geo<R3>::NB) * 16 * R3 + 64 * (e % Geo<R3>::NB) + 2 * lane;
}

// store one finished hop block, scaled by 1/(N window-sum-square) -- `scale` alone in the interior, where the
// window-sum-square is 3/2 exactly, `scale_tab[] * scale` at the first / last block -- and track max |y|
template <int R3, bool TRACK_MAX>
XD_HD void store_block(Lane<R3>& L, int lane, const float2* blk, float* ydst, const float* scale_tab, float scale) {
    typedef Geo<R3> G;
#pragma unroll
    for (int e = 0; e < 2 * G::NB; e++) {
        const int off = blk_off<R3>(lane, e);
        float2 o = blk[e];
        if (scale_tab) {
            o.x *= scale_tab[off] * scale;
            o.y *= scale_tab[off + 1] * scale;
        } else {
            o.x *= scale;
            o.y *= scale;
        }
        if (TRACK_MAX) L.amax = fmaxf(L.amax, fmaxf(fabsf(o.x), fabsf(o.y)));
        *reinterpret_cast<float2*>(ydst + off) = o;
    }
}

// F5b: overlap-add frame t into the register accumulators; returns in `out` the hop block
// t-2, which has now received every frame of this run that touches it
template <int R3>
XD_HD void ola_shift(Lane<R3>& L, float2* out) {
    typedef Geo<R3> G;
    // frame sample s = 16 R3 n1 + 2 m + e lies in quarter n1 >> 1 at element (n1 & 1) * NB + i
#pragma unroll
    for (int e = 0; e < 2 * G::NB; e++) {
        const int pbit = e / G::NB, i = e % G::NB;
        const float2 q0 = L.v[i * 8 + 0 + pbit], q1 = L.v[i * 8 + 2 + pbit];
        const float2 q2 = L.v[i * 8 + 4 + pbit], q3 = L.v[i * 8 + 6 + pbit];
        out[e] = cadd(L.acc[0][e], q0);
        L.acc[0][e] = cadd(L.acc[1][e], q1);
        L.acc[1][e] = cadd(L.acc[2][e], q2);
        L.acc[2][e] = q3;
    }
}

// ------------------------------------------------------------------ run-level bookkeeping
// A run [ta, tb) emits hop block b = t-2 after adding frame t.  The block has then seen frames
// max(ta, t-3)..t; it is complete iff ta == 0 or t-3 >= ta.  The three incomplete blocks at the
// head of a run (ta-2, ta-1, ta) and the three at its tail (tb-2, tb-1, tb) are written as raw
// partial sums to halo[boundary][side]; whichever of the two neighbouring warps arrives second
// adds left + right (fixed order -> deterministic) and stores the finished blocks.

template <int R3>
XD_HD float* halo_ptr(const GlParams& p, int boundary, int side, int blk) {
    return p.halo + ((size_t)(boundary * 2 + side) * 3 + blk) * Geo<R3>::H;
}

template <int R3>
XD_HD void store_partial(int lane, const float2* blk, float* dst) {
#pragma unroll
    for (int e = 0; e < 2 * Geo<R3>::NB; e++) *reinterpret_cast<float2*>(dst + blk_off<R3>(lane, e)) = blk[e];
}

// block t-2 after frame t (called by every lane); returns true when the caller must signal the
// boundary to its left (third partial block written)
template <int R3, bool TRACK_MAX>
XD_HD bool emit_block(Lane<R3>& L, int lane, const GlParams& p, int run_idx, const GlRun& r, long yoff, int t,
                      const float2* out) {
    typedef Geo<R3> G;
    const int b = t - 2;
    if (b < 0) return false;
    const bool complete = (r.ta == 0) || (t - 3 >= r.ta);
    if (complete) {
        store_block<R3, TRACK_MAX>(L, lane, out, p.y_out + yoff + (long)b * G::H, b == 0 ? p.edge_scale : nullptr,
                                   b == 0 ? p.inv_n : p.inv_n * (2.0f / 3.0f));
        return false;
    }
    store_partial<R3>(lane, out, halo_ptr<R3>(p, run_idx - 1, 1, t - r.ta));
    return t == r.ta + 2;
}

// end of a run whose right neighbour has already published its three partial blocks (the usual order: a
// run reaches its third frame long before its left neighbour reaches its last): finish blocks tb-2..tb from
// the register accumulators + the neighbour's partial sums.  Same operands in the same order as
// combine_boundary (left + right), so both paths give the same bits.
template <int R3, bool TRACK_MAX>
XD_HD void finish_tail_from_registers(Lane<R3>& L, int lane, const GlParams& p, int run_idx, const GlRun& r, long yoff) {
    typedef Geo<R3> G;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float* rp = halo_ptr<R3>(p, run_idx, 1, q);
        float2 blk[2 * G::NB];
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++)
            blk[e] = cadd(L.acc[q][e], ld_l2(reinterpret_cast<const float2*>(rp + blk_off<R3>(lane, e))));
        store_block<R3, TRACK_MAX>(L, lane, blk, p.y_out + yoff + (long)(r.tb - 2 + q) * G::H, nullptr, p.inv_n * (2.0f / 3.0f));
    }
}

// end of run; returns true when the caller must signal the boundary to its right.  neighbour_ready: the
// run to the right has already published its partial blocks (see finish_tail_from_registers) -- nothing is
// written to the halo then and the caller only restores the parity of the arrival counter.
template <int R3, bool TRACK_MAX>
XD_HD bool emit_tail(Lane<R3>& L, int lane, const GlParams& p, int run_idx, const GlRun& r, long yoff, int T,
                     bool neighbour_ready = false) {
    typedef Geo<R3> G;
    if (r.tb == T) {   // block T-2 is the last one; it is complete (frames T-3..T-1)
        store_block<R3, TRACK_MAX>(L, lane, L.acc[0], p.y_out + yoff + (long)(T - 2) * G::H, p.edge_scale + G::H, p.inv_n);
        return false;
    }
    if (neighbour_ready) {
        finish_tail_from_registers<R3, TRACK_MAX>(L, lane, p, run_idx, r, yoff);
        return false;
    }
#pragma unroll
    for (int q = 0; q < 3; q++) store_partial<R3>(lane, L.acc[q], halo_ptr<R3>(p, run_idx, 0, q));
    return true;
}

// second arrival at `boundary` (between runs boundary and boundary+1): finish blocks tb-2..tb
template <int R3, bool TRACK_MAX>
XD_HD void combine_boundary(Lane<R3>& L, int lane, const GlParams& p, int boundary) {
    typedef Geo<R3> G;
    const GlRun r = p.runs[boundary];
    const long yoff = (long)p.utt_foff[r.utt] * G::H;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float* lp = halo_ptr<R3>(p, boundary, 0, q);
        const float* rp = halo_ptr<R3>(p, boundary, 1, q);
        float2 blk[2 * G::NB];
#pragma unroll
        for (int e = 0; e < 2 * G::NB; e++) {
            const int off = blk_off<R3>(lane, e);
            const float2 a = ld_l2(reinterpret_cast<const float2*>(lp + off));
            const float2 c = ld_l2(reinterpret_cast<const float2*>(rp + off));
            blk[e] = cadd(a, c);
        }
        store_block<R3, TRACK_MAX>(L, lane, blk, p.y_out + yoff + (long)(r.tb - 2 + q) * G::H, nullptr, p.inv_n * (2.0f / 3.0f));
    }
}

// after the constant tables have landed in shared memory
template <int R3>
XD_HD void lane_load_constants(Lane<R3>& L, int lane, const float2* tab) {
    if (Geo<R3>::TW2_REGS) {
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) L.tw2[k2 - 1] = tab[Geo<R3>::TW2_OFF + (k2 - 1) * 32 + lane];
    }
}

template <int R3>
XD_HD void lane_reset(Lane<R3>& L) {
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
        for (int e = 0; e < 2 * Geo<R3>::NB; e++) L.acc[q][e] = mk2(0.f, 0.f);
    L.amax = 0.f;
}

}  // namespace xdtts
