// Tacotron2 decoder loop for sm_100a: ONE persistent cooperative kernel runs every decoder step of a
// batch of utterances, replacing the per-frame `decoder.run(inputs)` round trip of
// Tacotron2::run_decoder (/root/reference src/tacotron2/mod.rs:272-342: up to 1000 sequential ONNX Runtime
// calls, 11 tensors shuttled per call, `concatenate` of the growing spectrogram per frame :315).
//
// One step of the exported graph (NVIDIA Tacotron2 Decoder.decode, oracle/decoder_oracle.py):
//   prenet (2 x linear+relu+dropout) -> attention LSTM -> location-sensitive attention -> decoder LSTM
//   -> linear projection (mel frame) + gate.
// At batch 1 this is a chain of matrix-vector products: 18.2 M weights = 72.7 MB fp32 are read once per step (they stay
// in the 126 MB L2, a third of a CTA's share in its shared memory) and everything else is latency, so the design is about
// (1) streaming the two LSTM matrices with every SM at once, (2) keeping all state on chip or in L2 between steps,
// (3) making the seven hand-overs between CTAs that a step needs as short as the hardware allows, with the weight
// streaming placed so that it overlaps the small serial stages:
//
//   stage     on the critical path (published for the next stage)               streamed while the cells travel
//   P   prenet layer 1 (every CTA, redundantly) + layer 2 rows (one per CTA)    att. LSTM, h_att columns (2nd half)
//   A2  attention LSTM: the 256 prenet columns + cell update                    dec. LSTM, h_dec columns (2nd half)
//   Q   query rows (one per CTA)                                                dec. LSTM, h_att columns [0, 384)
//   E   attention energies, one encoder position per CTA (or per warp)          dec. LSTM, h_att columns [384, 768)
//   C   softmax (every CTA, redundantly) + context chunks (32 dims per CTA)     dec. LSTM, h_att columns [768, 1024)
//   D2  decoder LSTM: the 512 context columns + cell update                     NEXT step's att. LSTM, ctx columns
//   R   projection + gate rows (one per CTA)                                    NEXT step: att. LSTM h_att (1st half),
//                                                                               dec. LSTM h_dec (1st half)
// There is NO grid barrier: a stage publishes each of its results as a 64-bit cell {value, step tag} -- one store that
// carries the data and its own "ready" flag -- then streams LSTM weight columns whose input vectors are already known
// (82% of the bytes of a step), and the next stage polls exactly the cells it consumes (cell_put / cell_get below).
// Round 1 and most of round 2 used split-phase grid barriers (arrive -> stream -> wait); stamps around them showed
// ~1000 cycles per arrival spent waiting for the CTA's stores to be acknowledged before the release could go out, the
// arrival's own trip to the L2, the poll, and then one more L2 round trip to load the data: 19.2 us per step, 17.3 with cells.
//
// CTA c owns hidden units [7c, 7c+7) of both LSTMs: their four gate rows, the partial gate sums (per lane, in
// registers, across slices and stages) and the cell state never leave the SM.  A warp computes two weight rows
// at a time against up to 8 utterances' input vectors held in shared memory (float4 loads, all loads of a slice in
// flight; one shuffle reduction per row and step), so the weights are read once per step however many utterances
// decode in lockstep.  What bounds the step is dependent instruction chains and L2 latency in the serial stages, NOT
// weight traffic: keeping a third of a CTA's LSTM rows in shared memory for the whole loop (dec_slice below) moved
// the step by 2.5%.
// Attention weights / cumulative weights are kept by every CTA in shared memory (same instructions, same
// bits), hence never travel.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "decoder.h"
#include "gl_core.cuh"   // phase_turn: the counter-based generator shared with the vocoder's phase init

namespace xdtts {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Cross-CTA hand-over without barriers.  Every value one stage publishes for the next (a prenet output, a hidden unit, a
// query entry, an energy, a context dimension, a mel bin) travels as ONE 64-bit cell {value, tag}, tag = step + 1: a single
// store, atomic by size, carries the data and its own "ready" flag, and a consumer polls the cells it needs until their
// tag is this step's.  Against a grid barrier this removes, per stage, the release (~700 cycles waiting for the CTA's
// stores to be acknowledged before the arrival may be sent), the arrival's trip to the L2, and the separate load of the
// data after the wait: store -> L2 -> poll is the whole critical path (measured: 19.2 -> see DESIGN.md 3.5 us per step).
// A cell is rewritten one step later; by then every reader of the old value has published something that needed it
// (the stages of a step form one dependence chain through all CTAs), so one buffer per stage is enough.
// Polling is bounded: a cooperative launch keeps all CTAs resident, so a cell is only ever late, never lost, but a bound
// turns any mistake into an error code instead of a hung GPU.
constexpr unsigned DC_SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ void cell_put(dc_cell* c, float v, unsigned tag) {
    const dc_cell w = ((dc_cell)tag << 32) | (dc_cell)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(c), "l"(w) : "memory");
}
__device__ __forceinline__ dc_cell cell_peek(const dc_cell* c) {
    dc_cell w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(c) : "memory");
    return w;
}
__device__ __forceinline__ float cell_wait(const dc_cell* c, dc_cell w, unsigned tag, int* err) {   // w: a first peek
    unsigned spins = 0;
    while ((unsigned)(w >> 32) != tag) {
        if (++spins > DC_SPIN_LIMIT) {
            *err = 1;
            break;
        }
        w = cell_peek(c);
    }
    return __uint_as_float((unsigned)w);
}
// POLL = false: the batch-of-4-and-more form, where a grid barrier stands between producer and consumer and the cells are
// read once, unchecked (see hand_over below)
template <bool POLL>
__device__ __forceinline__ float cell_get(const dc_cell* c, unsigned tag, int* err) {
    const dc_cell w = cell_peek(c);
    return POLL ? cell_wait(c, w, tag, err) : __uint_as_float((unsigned)w);
}

// n <= PER * DC_THREADS cells -> shared memory, dst[(i / row) * ld + i % row].  All of a thread's peeks are in flight before
// the first is checked (one L2 round trip for the whole vector when the values are already there); cells that are not,
// are peeked again together after a short sleep -- 75 k threads re-polling without one flood the L2 the producers write through.
template <int PER, bool POLL>
__device__ __forceinline__ void cells_to_smem(const dc_cell* src, int n, int row, float* dst, int ld, unsigned tag, int* err) {
    if (!POLL) {
#pragma unroll 4
        for (int i = threadIdx.x; i < n; i += DC_THREADS) dst[(i / row) * ld + i % row] = __uint_as_float((unsigned)cell_peek(src + i));
        return;
    }
    unsigned pending = 0;
#pragma unroll
    for (int k = 0; k < PER; k++)
        if ((int)threadIdx.x + k * DC_THREADS < n) pending |= 1u << k;
    unsigned spins = 0;
    while (pending) {
        dc_cell w[PER];
#pragma unroll
        for (int k = 0; k < PER; k++)
            if (pending >> k & 1) w[k] = cell_peek(src + threadIdx.x + k * DC_THREADS);
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if ((pending >> k & 1) && (unsigned)(w[k] >> 32) == tag) {
                const int i = threadIdx.x + k * DC_THREADS;
                dst[(i / row) * ld + i % row] = __uint_as_float((unsigned)w[k]);
                pending &= ~(1u << k);
            }
        }
        if (pending) {
            if (++spins > (DC_SPIN_LIMIT >> 4)) {
                *err = 1;
                break;
            }
            if (PER > 4) __nanosleep(64);   // (batch 1-2: a few cells per thread, re-polled at once)
        }
    }
}

// The grid barrier of the batch >= 4 form (split-phase: arrive -> stream weight columns -> wait).  With 4 or 8 utterances in
// lockstep a stage hands over 4-8 k values per vector; polling each of them costs more (register-resident peeks, twice the
// bytes in flight) than the barrier's fixed ~1 us, so large batches keep the barrier and read the cells once, unchecked.
__device__ __forceinline__ void bar_arrive(unsigned* counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release at gpu scope: cumulative over the CTA's stores ordered before it by the barrier above.  (A
        // __threadfence() here compiles to MEMBAR.SC + an L1 invalidate and costs a fifth of the barrier.)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
}
__device__ __forceinline__ void bar_wait(unsigned* counter, unsigned& epoch, unsigned n_ctas) {
    epoch += n_ctas;
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < epoch);
    }
    __syncthreads();
}


// dot products of `ncols` consecutive weights (a multiple of 128) with NB vectors in shared memory
// (vector b starts at zs + b * zld); every lane ends up with the full sums added to acc[]
template <int NB>
__device__ __forceinline__ void row_dot(const float* __restrict__ wrow, const float* zs, int zld, int ncols, float (&acc)[NB]) {
    const int lane = threadIdx.x & 31;
    float a[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) a[b] = 0.f;
#pragma unroll 8
    for (int c = lane * 4; c < ncols; c += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + c));
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const float4 z = *reinterpret_cast<const float4*>(zs + b * zld + c);
            a[b] = fmaf(w.x, z.x, fmaf(w.y, z.y, fmaf(w.z, z.z, fmaf(w.w, z.w, a[b]))));
        }
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) a[b] += __shfl_xor_sync(0xffffffffu, a[b], o);
        acc[b] += a[b];
    }
}

// The gate sums of a warp's two rows (r = warp and warp + 16; r = gate * 7 + unit), accumulated PER LANE across the
// column slices of a step and reduced over the warp only once, when the row is complete (lstm_finish): a slice
// costs loads and FMAs, no shuffles.  The registers live across the grid barriers of the persistent kernel.
template <int NB>
struct GateAcc {
    float a0[NB], a1[NB];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int b = 0; b < NB; b++) a0[b] = a1[b] = 0.f;
    }
};

// adds weight columns [col0, col0 + 128 NC4) times the vectors that START at zs (the caller offsets zs to the
// segment).  The loads of BOTH rows are issued before either is used (these products are latency-bound: a slice is
// 2-4 float4 per lane and row, so everything that can be in flight must be).  Rows past the end (last CTAs, second
// row of warps 12..15) are clamped to a valid row and dropped in lstm_finish.
// cache: this CTA's 28 rows of exactly these columns in shared memory ([row][128 NC4], filled once in the prologue), or
// null -- the same values in the same order either way, so the result does not depend on what is cached.
template <int NB, int NC4>
__device__ __forceinline__ void lstm_partial(GateAcc<NB>& acc, const float* __restrict__ W, int ld, int col0, const float* zs,
                                             int zld, int unit0, const float* cache) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = DC_THREADS / 32;
    const int r0 = warp, r1 = warp + WARPS < 4 * DC_UNITS ? warp + WARPS : 4 * DC_UNITS - 1;
    const int u0 = min(unit0 + r0 % DC_UNITS, DC_RNN - 1), u1 = min(unit0 + r1 % DC_UNITS, DC_RNN - 1);
    const float* wr0 = W + (size_t)((r0 / DC_UNITS) * DC_RNN + u0) * ld + col0 + lane * 4;
    const float* wr1 = W + (size_t)((r1 / DC_UNITS) * DC_RNN + u1) * ld + col0 + lane * 4;
    float4 w0[NC4], w1[NC4];
    if (cache) {   // CTA-uniform
#pragma unroll
        for (int c = 0; c < NC4; c++) w0[c] = *reinterpret_cast<const float4*>(cache + r0 * (128 * NC4) + lane * 4 + 128 * c);
#pragma unroll
        for (int c = 0; c < NC4; c++) w1[c] = *reinterpret_cast<const float4*>(cache + r1 * (128 * NC4) + lane * 4 + 128 * c);
    } else {
#pragma unroll
        for (int c = 0; c < NC4; c++) w0[c] = __ldg(reinterpret_cast<const float4*>(wr0 + 128 * c));
#pragma unroll
        for (int c = 0; c < NC4; c++) w1[c] = __ldg(reinterpret_cast<const float4*>(wr1 + 128 * c));
    }
#pragma unroll
    for (int c = 0; c < NC4; c++) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const float4 z = *reinterpret_cast<const float4*>(zs + b * zld + lane * 4 + 128 * c);
            acc.a0[b] = fmaf(w0[c].x, z.x, fmaf(w0[c].y, z.y, fmaf(w0[c].z, z.z, fmaf(w0[c].w, z.w, acc.a0[b]))));
            acc.a1[b] = fmaf(w1[c].x, z.x, fmaf(w1[c].y, z.y, fmaf(w1[c].z, z.z, fmaf(w1[c].w, z.w, acc.a1[b]))));
        }
    }
}

// rows complete: reduce over the warp, add the bias, hand the gate sums to the cell update through shared memory
template <int NB>
__device__ __forceinline__ void lstm_finish(GateAcc<NB>& acc, int unit0, float* part, const float* __restrict__ bias) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = DC_THREADS / 32;
    const int r0 = warp, r1 = warp + WARPS < 4 * DC_UNITS ? warp + WARPS : 4 * DC_UNITS - 1;
    const bool ok0 = unit0 + r0 % DC_UNITS < DC_RNN, ok1 = warp + WARPS < 4 * DC_UNITS && unit0 + r1 % DC_UNITS < DC_RNN;
#pragma unroll
    for (int b = 0; b < NB; b++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            acc.a0[b] += __shfl_xor_sync(0xffffffffu, acc.a0[b], o);
            acc.a1[b] += __shfl_xor_sync(0xffffffffu, acc.a1[b], o);
        }
    }
    if (lane == 0) {
        if (ok0) {
            const float bv = bias[(r0 / DC_UNITS) * DC_RNN + unit0 + r0 % DC_UNITS];
#pragma unroll
            for (int b = 0; b < NB; b++) part[r0 * NB + b] = bv + acc.a0[b];
        }
        if (ok1) {
            const float bv = bias[(r1 / DC_UNITS) * DC_RNN + unit0 + r1 % DC_UNITS];
#pragma unroll
            for (int b = 0; b < NB; b++) part[r1 * NB + b] = bv + acc.a1[b];
        }
    }
    acc.clear();
}

// LSTM cell update of this CTA's units from the finished gate sums; h goes to global memory
template <int NB>
__device__ __forceinline__ void lstm_cell(const float* part, float* cst, int unit0, int nb, dc_cell* h_out, unsigned tag) {
    const int t = threadIdx.x;
    if (t < DC_UNITS * NB) {
        const int u = t / NB, b = t % NB, unit = unit0 + u;
        if (unit < DC_RNN && b < nb) {
            const float gi = part[(0 * DC_UNITS + u) * NB + b], gf = part[(1 * DC_UNITS + u) * NB + b];
            const float gg = part[(2 * DC_UNITS + u) * NB + b], go = part[(3 * DC_UNITS + u) * NB + b];
            const float c = sigmoidf_(gf) * cst[u * NB + b] + sigmoidf_(gi) * tanhf(gg);
            cst[u * NB + b] = c;
            cell_put(h_out + (size_t)b * DC_RNN + unit, sigmoidf_(go) * tanhf(c), tag);
        }
    }
}

__device__ __forceinline__ float keep_scale(const DecParams& p, int b, int step, int layer, int unit) {
    if (!p.dropout) return 1.0f;
    return phase_turn(p.seed, p.utt_base + b, 2 * DC_PRE, step, layer * DC_PRE + unit) <= 0.5f ? 2.0f : 0.0f;
}

// shared memory of one CTA (floats)
// LSTM weight slices a CTA may keep in shared memory for the whole loop instead of re-reading them from the L2 every step, in
// the order they are granted while shared memory lasts (dec_cache_mask): first the two slices that sit ON the critical path
// of a step (their inputs arrive with the barrier they follow, so their L2 latency cannot hide behind one), then slices
// streamed behind barriers (less L2 traffic per step: at batch 1 the step is bound by the 72.7 MB it reads from the L2).
// A slice is this CTA's 28 gate rows x 128 nc4 columns.
struct DecSlice { int mat, col0, nc4; };   // mat 0: attention LSTM (Wa), 1: decoder LSTM (Wd)
constexpr int DC_NSLICES = 6;
__host__ __device__ constexpr DecSlice dec_slice(int id) {
    return id == 0 ? DecSlice{0, DC_ENC + DC_RNN, 2}     // stage A2: prenet columns of the attention LSTM
         : id == 1 ? DecSlice{1, 2 * DC_RNN, 4}          // stage D2: context columns of the decoder LSTM
         : id == 2 ? DecSlice{0, 0, 4}                   // behind D2's barrier: context columns of the attention LSTM
         : id == 3 ? DecSlice{1, 768, 2}                 // behind C's barrier
         : id == 4 ? DecSlice{1, 384, 3}                 // behind E's barrier
         :           DecSlice{1, 0, 3};                  // behind Q's barrier
}
__host__ __device__ constexpr int dec_slice_floats(int id) { return 4 * DC_UNITS * 128 * dec_slice(id).nc4; }

template <int NB>
struct DecSmem {
    __host__ __device__ static constexpr int a4(int x) { return (x + 3) & ~3; }  // regions start on 16-byte boundaries (float4 reads)
    static constexpr int ZA_LD = DC_ENC + DC_RNN;                  // [ctx | h_att]: attention-LSTM columns streamed ahead
    static constexpr int ZD_LD = 2 * DC_RNN;                       // [h_att | h_dec]: decoder-LSTM columns streamed ahead
    static constexpr int ZA = 0;
    static constexpr int ZD = ZA + NB * ZA_LD;
    static constexpr int X1 = ZD + NB * ZD_LD;                     // [NB][256] prenet layer 1 output
    static constexpr int X2 = X1 + NB * DC_PRE;                    // [NB][256] prenet output (copy of the global one)
    static constexpr int WEFF = X2 + NB * DC_PRE;                  // [128][63]
    static constexpr int PART_A = a4(WEFF + DC_ATT * DC_WEFF_LD);  // [28][NB] attention-LSTM gate sums
    static constexpr int PART_D = a4(PART_A + 4 * DC_UNITS * NB);
    static constexpr int CST_A = a4(PART_D + 4 * DC_UNITS * NB);   // [7][NB] cell states
    static constexpr int CST_D = a4(CST_A + DC_UNITS * NB);
    static constexpr int XIN = a4(CST_D + DC_UNITS * NB);          // [NB][80]
    static constexpr int VS = a4(XIN + NB * DC_MEL);               // [128]
    static constexpr int DYN = a4(VS + DC_ATT);                    // then [NB][t_enc] new weights, [NB][2][t_enc + 30] padded w / w_cum
    __host__ __device__ static constexpr int cache_off(int t_enc) { return a4(DYN + NB * t_enc + NB * 2 * (t_enc + 30) + 8); }   // the weight cache follows
    static size_t bytes(int t_enc, unsigned cache_mask) {
        size_t f = (size_t)cache_off(t_enc);
        for (int id = 0; id < DC_NSLICES; id++)
            if (cache_mask >> id & 1) f += (size_t)dec_slice_floats(id);
        return sizeof(float) * f;
    }
};

// timing experiments (-DXDTTS_DEC_TRACE): clock stamps of one CTA's thread 0 at every hand-over of steps 100..102
#ifdef XDTTS_DEC_TRACE
__device__ long long g_dec_trace[2][4][32];
#define DC_STAMP(i) do { if ((blockIdx.x == 0 || blockIdx.x == 140) && threadIdx.x == 0 && step >= 100 && step < 104) g_dec_trace[blockIdx.x ? 1 : 0][step - 100][i] = clock64(); } while (0)
#else
#define DC_STAMP(i) do { } while (0)
#endif

template <int NB>
__global__ void __launch_bounds__(DC_THREADS, 1) dec_persist_kernel(const DecParams p) {
    extern __shared__ __align__(16) float sm[];
    typedef DecSmem<NB> S;
    constexpr int ZA_LD = S::ZA_LD, ZD_LD = S::ZD_LD;
    float* zA = sm + S::ZA;          // [NB][ctx(512) | h_att(1024)]
    float* zD = sm + S::ZD;          // [NB][h_att(1024) | h_dec(1024)]
    float* x1 = sm + S::X1;
    float* x2s = sm + S::X2;
    float* part_a = sm + S::PART_A;
    float* part_d = sm + S::PART_D;
    float* cst_a = sm + S::CST_A;
    float* cst_d = sm + S::CST_D;
    float* xin = sm + S::XIN;
    float* vs = sm + S::VS;
    float* weff = sm + S::WEFF;
    const int t_enc = p.t_enc, wld = t_enc + 2 * (DC_LOCK / 2);
    float* wnew = sm + S::DYN;                 // [NB][t_enc]
    float* wpad = wnew + NB * t_enc;           // [NB][2][wld]
    const float* wc[DC_NSLICES];               // cached weight slices (null: streamed from the L2)
    {
        float* cp = sm + S::cache_off(t_enc);
#pragma unroll
        for (int id = 0; id < DC_NSLICES; id++) {
            wc[id] = (p.cache_mask >> id & 1) ? cp : nullptr;
            if (p.cache_mask >> id & 1) cp += dec_slice_floats(id);
        }
    }

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x, nb = p.nb;
    const int unit0 = cta * DC_UNITS;
    constexpr int WARPS = DC_THREADS / 32;
    constexpr bool POLL = NB <= 2;             // hand-over by polled cells (batch 1-2) or by grid barrier (batch 4-8)
    unsigned epoch = 0;
    // between publishing a stage's results and consuming the next stage's inputs: a CTA-wide sync (the stages reuse shared
    // scratch) and, in the barrier form, the split-phase grid barrier around the streamed weight columns
    auto hand_over_publish = [&]() {
        if (POLL) __syncthreads();
        else bar_arrive(p.barrier);
    };
    auto hand_over_consume = [&]() {
        if (POLL) __syncthreads();
        else bar_wait(p.barrier, epoch, G);
    };
    constexpr int LIGHT_WARP = WARPS - 1;      // has one LSTM row where most warps have two

    // ---- prologue: constants and the all-zero DecoderState (src/tacotron2/mod.rs:212-236)
    for (int i = tid; i < DC_ATT * (2 * DC_LOCK); i += DC_THREADS)
        weff[(i / (2 * DC_LOCK)) * DC_WEFF_LD + i % (2 * DC_LOCK)] = p.Weff[i];
    for (int i = tid; i < DC_ATT; i += DC_THREADS) vs[i] = p.v[i];
#pragma unroll
    for (int id = 0; id < DC_NSLICES; id++) {
        if (!(p.cache_mask >> id & 1)) continue;
        const DecSlice sl = dec_slice(id);
        const float* W = sl.mat ? p.Wd : p.Wa;
        const int ld = sl.mat ? DC_ZD : DC_ZA;
        float* dst = const_cast<float*>(wc[id]);
        for (int i = tid; i < 4 * DC_UNITS * 32 * sl.nc4; i += DC_THREADS) {   // row r = gate * 7 + unit, as lstm_partial indexes it
            const int r = i / (32 * sl.nc4), c4 = i % (32 * sl.nc4);
            const int u = min(cta * DC_UNITS + r % DC_UNITS, DC_RNN - 1);
            *reinterpret_cast<float4*>(dst + r * (128 * sl.nc4) + 4 * c4) =
                __ldg(reinterpret_cast<const float4*>(W + (size_t)((r / DC_UNITS) * DC_RNN + u) * ld + sl.col0 + 4 * c4));
        }
    }
    for (int i = tid; i < NB * 2 * wld; i += DC_THREADS) wpad[i] = 0.f;
    for (int i = tid; i < NB * (ZA_LD + ZD_LD); i += DC_THREADS) zA[i] = 0.f;   // zA and zD are adjacent
    for (int i = tid; i < DC_UNITS * NB; i += DC_THREADS) { cst_a[i] = 0.f; cst_d[i] = 0.f; }
    GateAcc<NB> acc_a, acc_d;   // zero state: the columns streamed ahead of step 0 contribute nothing
    acc_a.clear();
    acc_d.clear();
    bool done[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) done[b] = b >= nb;
    __syncthreads();

    int step = 0;
    for (; step < p.max_steps; step++) {
        const int cur = step & 1, nxt = cur ^ 1;
        const unsigned tag = (unsigned)step + 1u;   // of everything this step publishes
        DC_STAMP(30);
        // ================= stage P: prenet of this step  ||  attention LSTM, second half of its h_att columns
        for (int i = tid; i < nb * DC_MEL; i += DC_THREADS) {
            const int b = i / DC_MEL, k = i % DC_MEL;
            xin[b * DC_MEL + k] = step ? cell_get<POLL>(p.melt + b * (DC_MEL + 1) + k, tag - 1u, p.err) : 0.f;   // the previous step's frame
        }
        __syncthreads();
        {   // layer 1: thread = (output row, half of the 80 inputs); the 40 weight loads of a thread are all in flight
            const int r = tid & (DC_PRE - 1), half = tid >> 8;
            float wv[DC_MEL / 2];
#pragma unroll
            for (int k = 0; k < DC_MEL / 2; k++) wv[k] = __ldg(p.p1T + (half * (DC_MEL / 2) + k) * DC_PRE + r);
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
#pragma unroll
            for (int k = 0; k < DC_MEL / 2; k++) {
#pragma unroll
                for (int b = 0; b < NB; b++) acc[b] = fmaf(wv[k], xin[b * DC_MEL + half * (DC_MEL / 2) + k], acc[b]);
            }
            if (half) {
#pragma unroll
                for (int b = 0; b < NB; b++) x2s[b * DC_PRE + r] = acc[b];   // x2s is free until stage A2
            }
            __syncthreads();
            if (!half) {
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) x1[b * DC_PRE + r] = fmaxf(acc[b] + x2s[b * DC_PRE + r], 0.f) * keep_scale(p, b, step, 0, r);
            }
        }
        __syncthreads();
        for (int r = cta + G * warp; r < DC_PRE; r += G * WARPS) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            row_dot<NB>(p.p2 + (size_t)r * DC_PRE, x1, DC_PRE, DC_PRE, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) cell_put(p.x2 + b * DC_PRE + r, fmaxf(acc[b], 0.f) * keep_scale(p, b, step, 1, r), tag);
            }
        }
        DC_STAMP(0);
        hand_over_publish();
        if (step) lstm_partial<NB, 4>(acc_a, p.Wa, DC_ZA, DC_ENC + DC_RNN / 2, zA + DC_ENC + DC_RNN / 2, ZA_LD, unit0, nullptr);
        DC_STAMP(1);
        hand_over_consume();
        DC_STAMP(2);

        // ================= stage A2: attention LSTM, prenet columns + cell update  ||  decoder LSTM, second half of h_dec columns
        cells_to_smem<(NB * DC_PRE + DC_THREADS - 1) / DC_THREADS, POLL>(p.x2, nb * DC_PRE, DC_PRE, x2s, DC_PRE, tag, p.err);
        __syncthreads();
        lstm_partial<NB, 2>(acc_a, p.Wa, DC_ZA, DC_ENC + DC_RNN, x2s, DC_PRE, unit0, wc[0]);
        lstm_finish<NB>(acc_a, unit0, part_a, p.ba);
        __syncthreads();
        lstm_cell<NB>(part_a, cst_a, unit0, nb, p.h_a + (size_t)nxt * nb * DC_RNN, tag);
        DC_STAMP(3);
        hand_over_publish();
        if (step) lstm_partial<NB, 4>(acc_d, p.Wd, DC_ZD, DC_RNN + DC_RNN / 2, zD + DC_RNN + DC_RNN / 2, ZD_LD, unit0, nullptr);
        DC_STAMP(4);
        hand_over_consume();
        DC_STAMP(5);

        // ================= stage Q: query rows  ||  decoder LSTM, h_att columns [0, 384)
        cells_to_smem<(NB * DC_RNN + DC_THREADS - 1) / DC_THREADS, POLL>(p.h_a + (size_t)nxt * nb * DC_RNN, nb * DC_RNN, DC_RNN, zD, ZD_LD, tag, p.err);
        __syncthreads();
        if (warp == LIGHT_WARP && cta < DC_ATT) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            row_dot<NB>(p.Wq + (size_t)cta * DC_RNN, zD, ZD_LD, DC_RNN, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) cell_put(p.pq + b * DC_ATT + cta, acc[b], tag);
            }
        }
        DC_STAMP(6);
        hand_over_publish();
        lstm_partial<NB, 3>(acc_d, p.Wd, DC_ZD, 0, zD, ZD_LD, unit0, wc[5]);
        DC_STAMP(7);
        hand_over_consume();
        DC_STAMP(8);

        // ================= stage E: energies e[b][t] = v . tanh(pq + Weff * [w; w_cum](t-15..t+15) + pm[t])  ||  h_att columns [384, 768)
        if (nb * t_enc <= G) {
            // at most one encoder position per CTA (the reference's shape: one utterance, 100 positions): the whole
            // CTA computes it -- thread = (attention dim, quarter of the 62 taps) -- instead of one warp grinding
            // through 250 dependent FMAs.  Same partial sums, same order of additions as the warp path below.
            const int b = cta / t_enc, t = cta % t_enc;
            const bool have = cta < nb * t_enc && t < p.t_len[b < nb ? b : 0];   // CTA-uniform
            float* scratch = x1;   // [4][128] partial sums, then [128] tanh values (x1 and x2s are adjacent and free here)
            if (have) {
                const int a = tid & (DC_ATT - 1), g = tid >> 7;            // g: 0/1 even/odd taps of w, 2/3 of w_cum
                const float* f = weff + a * DC_WEFF_LD + (g >> 1) * DC_LOCK;
                const float* wv = wpad + (b * 2 + (g >> 1)) * wld + t;
                float pa = 0.f;
#pragma unroll
                for (int k = 0; k < DC_LOCK; k += 2)
                    if (k + (g & 1) < DC_LOCK) pa = fmaf(f[k + (g & 1)], wv[k + (g & 1)], pa);
                scratch[g * DC_ATT + a] = pa;
            }
            float q = 0.f, m = 0.f;
            if (have && tid < DC_ATT) {
                q = cell_get<POLL>(p.pq + b * DC_ATT + tid, tag, p.err);
                m = __ldg(p.pm + ((size_t)b * t_enc + t) * DC_ATT + tid);
            }
            __syncthreads();
            float th = 0.f;
            if (have && tid < DC_ATT)
                th = tanhf(q + ((scratch[tid] + scratch[DC_ATT + tid]) + (scratch[2 * DC_ATT + tid] + scratch[3 * DC_ATT + tid])) + m);
            __syncthreads();
            if (have && tid < DC_ATT) scratch[tid] = th;
            __syncthreads();
            if (have && warp == 0) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < DC_ATT / 32; j++) s = fmaf(vs[lane + 32 * j], scratch[lane + 32 * j], s);
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) cell_put(p.e + b * t_enc + t, s, tag);
            }
        } else {
            for (int item = cta + G * warp; item < nb * t_enc; item += G * WARPS) {
                const int b = item / t_enc, t = item % t_enc;
                if (t >= p.t_len[b]) continue;
                const float* w0 = wpad + (b * 2 + 0) * wld + t;   // padded by 15 on both sides: index t <-> position t - 15
                const float* w1 = wpad + (b * 2 + 1) * wld + t;
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < DC_ATT / 32; j++) {
                    const int a = lane + 32 * j;
                    const float* f = weff + a * DC_WEFF_LD;
                    // four independent partial sums: a single chain of 62 dependent FMAs is pure latency for the one warp
                    // that owns this encoder position
                    float pa0 = 0.f, pa1 = 0.f, pa2 = 0.f, pa3 = 0.f;
#pragma unroll
                    for (int k = 0; k < DC_LOCK; k += 2) {
                        pa0 = fmaf(f[k], w0[k], pa0);
                        pa2 = fmaf(f[DC_LOCK + k], w1[k], pa2);
                        if (k + 1 < DC_LOCK) {
                            pa1 = fmaf(f[k + 1], w0[k + 1], pa1);
                            pa3 = fmaf(f[DC_LOCK + k + 1], w1[k + 1], pa3);
                        }
                    }
                    const float pa = (pa0 + pa1) + (pa2 + pa3);
                    const float q = cell_get<POLL>(p.pq + b * DC_ATT + a, tag, p.err);
                    const float m = __ldg(p.pm + ((size_t)b * t_enc + t) * DC_ATT + a);
                    s = fmaf(vs[a], tanhf(q + pa + m), s);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) cell_put(p.e + b * t_enc + t, s, tag);
            }
        }
        DC_STAMP(9);
        hand_over_publish();
        lstm_partial<NB, 3>(acc_d, p.Wd, DC_ZD, 384, zD + 384, ZD_LD, unit0, wc[4]);
        DC_STAMP(10);
        hand_over_consume();
        DC_STAMP(11);

        // ================= stage C: softmax (every CTA keeps w / w_cum itself) + context chunks  ||  h_att columns [768, 1024)
        if (NB <= 2) {
            // one or two utterances: thread = encoder position, so the exponentials are one instruction deep instead
            // of a loop in a single warp.  Same sums in the same order as the warp path below (maximum is exact;
            // the sum is taken per lane over t = lane + 32 j, then by the same butterfly).
            float* red = x1;   // [17] scratch, free in this stage
            const int nj = (t_enc + 31) / 32;
            for (int b = 0; b < nb; b++) {
                const int tl = p.t_len[b];
                const float ev = tid < tl ? cell_get<POLL>(p.e + b * t_enc + tid, tag, p.err) : -INFINITY;
                float m = ev;
#pragma unroll
                for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0) red[warp] = m;
                __syncthreads();
                m = red[0];
#pragma unroll
                for (int w = 1; w < WARPS; w++) m = fmaxf(m, red[w]);
                const float ex = tid < tl ? expf(ev - m) : 0.f;
                if (tid < t_enc) wnew[b * t_enc + tid] = ex;
                __syncthreads();
                if (warp == 0) {
                    float sum = 0.f;
                    for (int j = 0; j < nj; j++) sum += (lane + 32 * j) < t_enc ? wnew[b * t_enc + lane + 32 * j] : 0.f;
#pragma unroll
                    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    if (lane == 0) red[WARPS] = sum;
                }
                __syncthreads();
                if (tid < t_enc) {
                    const float w = ex / red[WARPS];
                    wnew[b * t_enc + tid] = w;
                    wpad[(b * 2 + 0) * wld + DC_LOCK / 2 + tid] = w;
                    wpad[(b * 2 + 1) * wld + DC_LOCK / 2 + tid] += w;
                    if (p.align_out && cta == 0) p.align_out[((size_t)b * p.max_steps + step) * t_enc + tid] = w;
                }
            }
        } else         if (warp < nb) {
            const int b = warp, tl = p.t_len[b];
            float ev[DC_MAX_TENC / 32];
            float m = -INFINITY;
            const int nj = (t_enc + 31) / 32;   // the loops below are unrolled to 16 but stop at the encoder length
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
                if (j >= nj) break;
                const int t = lane + 32 * j;
                ev[j] = t < tl ? cell_get<POLL>(p.e + b * t_enc + t, tag, p.err) : -INFINITY;
                m = fmaxf(m, ev[j]);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
                if (j >= nj) break;
                ev[j] = (lane + 32 * j) < tl ? expf(ev[j] - m) : 0.f;
                sum += ev[j];
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
                if (j >= nj) break;
                const int t = lane + 32 * j;
                if (t < t_enc) {
                    const float w = ev[j] / sum;
                    wnew[b * t_enc + t] = w;
                    wpad[(b * 2 + 0) * wld + DC_LOCK / 2 + t] = w;
                    wpad[(b * 2 + 1) * wld + DC_LOCK / 2 + t] += w;
                    if (p.align_out && cta == 0) p.align_out[((size_t)b * p.max_steps + step) * t_enc + t] = w;
                }
            }
        }
        __syncthreads();
        {   // context: a CTA owns 32 dims of one utterance (lane = dim); its 16 warps split the encoder positions
            // (t = warp, warp + 16, ...: independent loads, all in flight), then the 16 partial sums are added in warp
            // order -- the same summation order whatever the batch, so a batch returns the bits of single calls
            float* scratch = x1;   // [WARPS][32], free in this stage
            for (int item = cta; item < nb * (DC_ENC / 32); item += G) {
                const int b = item / (DC_ENC / 32), d0 = (item % (DC_ENC / 32)) * 32, tl = p.t_len[b];
                const float* mp = p.memory + (size_t)b * t_enc * DC_ENC + d0 + lane;
                float acc = 0.f;
#pragma unroll 8
                for (int t = warp; t < tl; t += WARPS) acc = fmaf(wnew[b * t_enc + t], __ldg(mp + (size_t)t * DC_ENC), acc);
                scratch[warp * 32 + lane] = acc;
                __syncthreads();
                if (tid < 32) {
                    float sum = 0.f;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) sum += scratch[w * 32 + tid];
                    cell_put(p.ctx + b * DC_ENC + d0 + tid, sum, tag);   // read by every CTA in stage D2
                }
                __syncthreads();
            }
        }
        DC_STAMP(12);
        hand_over_publish();
        lstm_partial<NB, 2>(acc_d, p.Wd, DC_ZD, 768, zD + 768, ZD_LD, unit0, wc[3]);
        DC_STAMP(13);
        hand_over_consume();
        DC_STAMP(14);

        // ================= stage D2: decoder LSTM, context columns + cell update  ||  attention LSTM of the NEXT step, context columns
        cells_to_smem<(NB * DC_ENC + DC_THREADS - 1) / DC_THREADS, POLL>(p.ctx, nb * DC_ENC, DC_ENC, zA, ZA_LD, tag, p.err);
        for (int i = tid; i < nb * DC_RNN; i += DC_THREADS)
            zA[(i / DC_RNN) * ZA_LD + DC_ENC + i % DC_RNN] = zD[(i / DC_RNN) * ZD_LD + i % DC_RNN];   // h_att of this step
        __syncthreads();
        lstm_partial<NB, 4>(acc_d, p.Wd, DC_ZD, 2 * DC_RNN, zA, ZA_LD, unit0, wc[1]);
        lstm_finish<NB>(acc_d, unit0, part_d, p.bd);
        __syncthreads();
        lstm_cell<NB>(part_d, cst_d, unit0, nb, p.h_d + (size_t)nxt * nb * DC_RNN, tag);
        DC_STAMP(15);
        hand_over_publish();
        lstm_partial<NB, 4>(acc_a, p.Wa, DC_ZA, 0, zA, ZA_LD, unit0, wc[2]);
        DC_STAMP(16);
        hand_over_consume();
        DC_STAMP(17);

        // ================= stage R: projection + gate rows  ||  next step: attention LSTM h_att columns (first half),
        //                   decoder LSTM h_dec columns (first half)
        cells_to_smem<(NB * DC_RNN + DC_THREADS - 1) / DC_THREADS, POLL>(p.h_d + (size_t)nxt * nb * DC_RNN, nb * DC_RNN, DC_RNN, zD + DC_RNN, ZD_LD, tag, p.err);
        __syncthreads();
        if (warp == LIGHT_WARP && cta <= DC_MEL) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            const float* wrow = p.Wp + (size_t)cta * DC_ZP;              // columns [ctx | h_dec]
            row_dot<NB>(wrow, zA, ZA_LD, DC_ENC, acc);
            row_dot<NB>(wrow + DC_ENC, zD + DC_RNN, ZD_LD, DC_RNN, acc);
            if (lane == 0) {
                const float bias = p.bp[cta];
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (b >= nb) continue;
                    if (cta < DC_MEL) p.mel_out[((size_t)b * p.max_steps + step) * DC_MEL + cta] = acc[b] + bias;
                    else p.gate_out[(size_t)b * p.max_steps + step] = acc[b] + bias;
                    cell_put(p.melt + b * (DC_MEL + 1) + cta, acc[b] + bias, tag);   // the in-loop copy: next step's prenet, the stop rule
                }
            }
        }
        DC_STAMP(18);
        hand_over_publish();
        lstm_partial<NB, 4>(acc_a, p.Wa, DC_ZA, DC_ENC, zA + DC_ENC, ZA_LD, unit0, nullptr);
        lstm_partial<NB, 4>(acc_d, p.Wd, DC_ZD, DC_RNN, zD + DC_RNN, ZD_LD, unit0, nullptr);
        DC_STAMP(19);
        hand_over_consume();
        DC_STAMP(20);

        // ================= stop rule (src/tacotron2/mod.rs:319-324): every CTA takes the same decision
        bool all = true;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (b < nb && !done[b]) {
                const float g = cell_get<POLL>(p.melt + b * (DC_MEL + 1) + DC_MEL, tag, p.err);
                // the reference's sigmoid (src/tacotron2/mod.rs:126-133)
                const float sg = g >= 0.f ? 1.0f / (1.0f + expf(-g)) : expf(g) / (1.0f + expf(g));
                if (sg > p.gate_threshold) {
                    done[b] = true;
                    if (cta == 0 && tid == 0) p.n_frames[b] = step + 1;
                }
            }
            all = all && done[b];
        }
        if (all) break;
    }
    if (cta == 0 && tid == 0) {
#pragma unroll
        for (int b = 0; b < NB; b++)
            if (b < nb && !done[b]) p.n_frames[b] = p.max_steps;
    }
}

// ------------------------------------------------------------------ output layout
// decoder frames [nb][max_steps][80] -> per utterance [80][n_frames] row-major at dst + b * 80 * max_steps
// (the `mel_spec.t()` of src/tacotron2/mod.rs:345, so that the postnet / the caller see ndarray's [80, T])
__global__ void __launch_bounds__(256) dec_transpose_kernel(const float* __restrict__ src, const int* __restrict__ n_frames,
                                                            int max_steps, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, T = n_frames[b];
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if (t0 >= T) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, c = c0 + tx;
        tile[r][tx] = (t < T && c < DC_MEL) ? src[((size_t)b * max_steps + t) * DC_MEL + c] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, t = t0 + tx;
        if (c < DC_MEL && t < T) dst[(size_t)b * DC_MEL * max_steps + (size_t)c * T + t] = tile[tx][r];
    }
}

cudaError_t dec_launch_transpose(const float* mel_frames, const int* n_frames, int nb, int max_steps, float* dst, cudaStream_t s) {
    dim3 grid((max_steps + 31) / 32, (DC_MEL + 31) / 32, nb);
    dec_transpose_kernel<<<grid, 256, 0, s>>>(mel_frames, n_frames, max_steps, dst);
    return cudaGetLastError();
}

constexpr int DC_SMEM_MAX = 232448;   // 227 KB

// the weight slices that fit beside the kernel's own shared memory, in dec_slice's order of preference
template <int NB>
static unsigned cache_mask_nb(int t_enc) {
    static const bool off = getenv("XDTTS_DEC_NO_CACHE") != nullptr;   // measurement: stream everything from the L2
    unsigned mask = 0;
    if (off) return 0;
    for (int id = 0; id < DC_NSLICES; id++)
        if (DecSmem<NB>::bytes(t_enc, mask | 1u << id) <= (size_t)DC_SMEM_MAX) mask |= 1u << id;
    return mask;
}

template <int NB>
static cudaError_t prepare_nb() {
    return cudaFuncSetAttribute(dec_persist_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM_MAX);
}

cudaError_t dec_prepare(int* grid_out) {
    cudaError_t e = prepare_nb<1>();
    if (e == cudaSuccess) e = prepare_nb<2>();
    if (e == cudaSuccess) e = prepare_nb<4>();
    if (e == cudaSuccess) e = prepare_nb<8>();
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dec_persist_kernel<8>, DC_THREADS, DC_SMEM_MAX);
    if (e != cudaSuccess) return e;
    const int need = (DC_RNN + DC_UNITS - 1) / DC_UNITS;   // CTAs that own hidden units
    if (per_sm < 1 || sms < need) return cudaErrorInvalidConfiguration;
    *grid_out = sms;   // one CTA per SM; cooperative launch keeps them co-resident
    return cudaSuccess;
}

template <int NB>
static cudaError_t launch_nb(const DecParams& p, int grid, cudaStream_t s) {
    DecParams q = p;
    q.cache_mask = cache_mask_nb<NB>(p.t_enc);
    void* args[1] = {&q};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)dec_persist_kernel<NB>, dim3(grid), dim3(DC_THREADS), args,
                                                DecSmem<NB>::bytes(p.t_enc, q.cache_mask), s);
#ifdef XDTTS_DEC_TRACE
    if (e == cudaSuccess && p.max_steps > 104) {
        cudaStreamSynchronize(s);
        static long long tr[2][4][32];
        cudaMemcpyFromSymbol(tr, g_dec_trace, sizeof(tr));
        static const char* names[7] = {"P", "A2", "Q", "E", "C", "D2", "R"};
        for (int c = 0; c < 2; c++) {
            fprintf(stderr, "decoder trace, CTA %d, step 101 (cycles): stage: inputs + work until published | streamed columns | wait\n", c ? 140 : 0);
            long long t = tr[c][1][30];
            for (int k = 0; k < 7; k++) {
                fprintf(stderr, "  %-2s %6lld | %6lld | %6lld\n", names[k], tr[c][1][3 * k] - t, tr[c][1][3 * k + 1] - tr[c][1][3 * k],
                        tr[c][1][3 * k + 2] - tr[c][1][3 * k + 1]);
                t = tr[c][1][3 * k + 2];
            }
            fprintf(stderr, "  stop rule %lld, step total %lld\n", tr[c][2][30] - t, tr[c][2][30] - tr[c][1][30]);
        }
    }
#endif
    return e;
}

// how a launch for nb utterances at t_enc hands over and what it keeps on chip: info[0] = LSTM / projection / prenet weight bytes a
// step reads (grid-wide), [1] = of those, bytes resident in shared memory for the whole loop, [2] = hand-overs between CTAs per
// step, [3] = 1 when they are polled cells, 0 when grid barriers
void dec_info(int nb, int t_enc, int grid, long long* info) {
    unsigned mask = nb == 1 ? cache_mask_nb<1>(t_enc) : nb == 2 ? cache_mask_nb<2>(t_enc) : nb <= 4 ? cache_mask_nb<4>(t_enc) : cache_mask_nb<8>(t_enc);
    long long cached = 0;
    for (int id = 0; id < DC_NSLICES; id++)
        if (mask >> id & 1) cached += (long long)dec_slice_floats(id) * 4;
    const long long owners = (DC_RNN + DC_UNITS - 1) / DC_UNITS;   // CTAs that own hidden units (the last one fewer than 7)
    info[0] = 4ll * ((long long)DC_MEL * DC_PRE + (long long)DC_PRE * DC_PRE + 4ll * DC_RNN * (DC_ZA + DC_ZD) + (long long)DC_ATT * DC_RNN +
                     (long long)(DC_MEL + 1) * DC_ZP);
    info[1] = cached * (owners < grid ? owners : grid);
    info[2] = 7;
    info[3] = nb <= 2 ? 1 : 0;
}

cudaError_t dec_launch(const DecParams& p, int grid, cudaStream_t s) {
    if (p.nb < 1 || p.nb > DC_MAX_NB || p.t_enc < 1 || p.t_enc > DC_MAX_TENC) return cudaErrorInvalidValue;
    if (p.nb == 1) return launch_nb<1>(p, grid, s);
    if (p.nb == 2) return launch_nb<2>(p, grid, s);
    if (p.nb <= 4) return launch_nb<4>(p, grid, s);
    return launch_nb<8>(p, grid, s);
}

}  // namespace xdtts
