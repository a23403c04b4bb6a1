// Tacotron2 decoder loop for sm_100a: ONE persistent cooperative kernel runs every decoder step of a
// batch of utterances, replacing the per-frame `decoder.run(inputs)` round trip of
// Tacotron2::run_decoder (/root/reference src/tacotron2/mod.rs:272-342: up to 1000 sequential ONNX Runtime
// calls, 11 tensors shuttled per call, `concatenate` of the growing spectrogram per frame :315).
//
// One step of the exported graph (NVIDIA Tacotron2 Decoder.decode, oracle/decoder_oracle.py):
//   prenet (2 x linear+relu+dropout) -> attention LSTM -> location-sensitive attention -> decoder LSTM
//   -> linear projection (mel frame) + gate.
// At batch 1 this is a chain of matrix-vector products: 18.4 M weights = 73.5 MB fp32 are read once per
// step and everything else is latency, so the design is about (1) streaming the two LSTM matrices with
// every SM at once, (2) keeping all state on chip or in L2 between steps, (3) as few grid-wide barriers as
// the data dependences allow, with the weight streaming placed so that it overlaps the small serial stages:
//
//   stage P   prenet layer 1 (every CTA, redundantly) + layer 2 rows (one per CTA)               | barrier
//   stage A2  attention LSTM: the 256 prenet columns + cell update                               | barrier
//   stage Q   query rows (one per CTA)  +  decoder LSTM over its [h_att | h_dec] columns (D1)     | barrier
//   stage E   attention energies, one warp per encoder position                                  | barrier
//   stage C   softmax (every CTA, redundantly) + context chunks (one warp per 32 dims)            | barrier
//   stage D2  decoder LSTM: the 512 context columns + cell update                                | barrier
//   stage R   projection + gate rows (one per CTA)  +  attention LSTM of the NEXT step over its
//             [ctx | h_att] columns (A1)                                                        | barrier
//
// CTA c owns hidden units [7c, 7c+7) of both LSTMs: their four gate rows, the partial gate sums (which live
// in shared memory across barriers) and the cell state never leave the SM.  A warp computes one weight row
// at a time against up to 8 utterances' input vectors held in shared memory (float4 loads, shuffle
// reduction), so the weights are read once per step however many utterances decode in lockstep.
// Attention weights / cumulative weights are kept by every CTA in shared memory (same instructions, same
// bits), hence never travel.  All cross-CTA state is read through L2 (ld.global.cg).
#include <cuda_runtime.h>

#include "decoder.h"
#include "gl_core.cuh"   // phase_turn: the counter-based generator shared with the vocoder's phase init

namespace xdtts {

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned n_ctas) {
    __syncthreads();
    epoch += n_ctas;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < epoch);
    }
    __syncthreads();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// dot products of `ncols` consecutive weights (a multiple of 128) with NB vectors in shared memory
// (vector b starts at zs + b * zld); every lane ends up with the full sums added to acc[]
template <int NB>
__device__ __forceinline__ void row_dot(const float* __restrict__ wrow, const float* zs, int zld, int ncols, float (&acc)[NB]) {
    const int lane = threadIdx.x & 31;
    float a[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) a[b] = 0.f;
#pragma unroll 4
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

// partial gate sums of this CTA's units over weight columns [col0, col0 + ncols) against the vectors that START at
// zs (the caller offsets zs to the segment): part[r][b], r = gate * 7 + unit
template <int NB>
__device__ __forceinline__ void lstm_partial(const float* __restrict__ W, int ld, int col0, int ncols, const float* zs, int zld,
                                             int unit0, float* part, const float* __restrict__ bias_or_null) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < 4 * DC_UNITS; r += DC_THREADS / 32) {
        const int unit = unit0 + r % DC_UNITS;
        if (unit >= DC_RNN) continue;
        const int row = (r / DC_UNITS) * DC_RNN + unit;
        float acc[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) acc[b] = 0.f;
        row_dot<NB>(W + (size_t)row * ld + col0, zs, zld, ncols, acc);
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; b++) part[r * NB + b] = (bias_or_null ? bias_or_null[row] : part[r * NB + b]) + acc[b];
        }
    }
}

// LSTM cell update of this CTA's units from the finished gate sums; h goes to global memory
template <int NB>
__device__ __forceinline__ void lstm_cell(const float* part, float* cst, int unit0, int nb, float* h_out) {
    const int t = threadIdx.x;
    if (t < DC_UNITS * NB) {
        const int u = t / NB, b = t % NB, unit = unit0 + u;
        if (unit < DC_RNN && b < nb) {
            const float gi = part[(0 * DC_UNITS + u) * NB + b], gf = part[(1 * DC_UNITS + u) * NB + b];
            const float gg = part[(2 * DC_UNITS + u) * NB + b], go = part[(3 * DC_UNITS + u) * NB + b];
            const float c = sigmoidf_(gf) * cst[u * NB + b] + sigmoidf_(gi) * tanhf(gg);
            cst[u * NB + b] = c;
            h_out[(size_t)b * DC_RNN + unit] = sigmoidf_(go) * tanhf(c);
        }
    }
}

__device__ __forceinline__ float keep_scale(const DecParams& p, int b, int step, int layer, int unit) {
    if (!p.dropout) return 1.0f;
    return phase_turn(p.seed, p.utt_base + b, 2 * DC_PRE, step, layer * DC_PRE + unit) <= 0.5f ? 2.0f : 0.0f;
}

// shared memory of one CTA (floats)
template <int NB>
struct DecSmem {
    static constexpr int a4(int x) { return (x + 3) & ~3; }  // regions start on 16-byte boundaries (float4 reads)
    static constexpr int ZLD = DC_ZD;
    static constexpr int Z = 0;                                    // [NB][2560] input vectors of the current stage
    static constexpr int X1 = Z + NB * ZLD;                        // [NB][256] prenet layer 1 output
    static constexpr int WEFF = X1 + NB * DC_PRE;                  // [128][63]
    static constexpr int PART_A = a4(WEFF + DC_ATT * DC_WEFF_LD);  // [28][NB] attention-LSTM gate sums
    static constexpr int PART_D = a4(PART_A + 4 * DC_UNITS * NB);
    static constexpr int CST_A = a4(PART_D + 4 * DC_UNITS * NB);   // [7][NB] cell states
    static constexpr int CST_D = a4(CST_A + DC_UNITS * NB);
    static constexpr int XIN = a4(CST_D + DC_UNITS * NB);          // [NB][80]
    static constexpr int VS = a4(XIN + NB * DC_MEL);               // [128]
    static constexpr int DYN = a4(VS + DC_ATT);                    // then [NB][t_enc] new weights, [NB][2][t_enc + 30] padded w / w_cum
    static size_t bytes(int t_enc) { return sizeof(float) * (size_t)(DYN + NB * t_enc + NB * 2 * (t_enc + 30) + 8); }
};

template <int NB>
__global__ void __launch_bounds__(DC_THREADS, 1) dec_persist_kernel(const DecParams p) {
    extern __shared__ __align__(16) float sm[];
    typedef DecSmem<NB> S;
    constexpr int ZLD = S::ZLD;
    float* zs = sm + S::Z;
    float* part_a = sm + S::PART_A;
    float* part_d = sm + S::PART_D;
    float* cst_a = sm + S::CST_A;
    float* cst_d = sm + S::CST_D;
    float* x1 = sm + S::X1;
    float* xin = sm + S::XIN;
    float* vs = sm + S::VS;
    float* weff = sm + S::WEFF;
    const int t_enc = p.t_enc, wld = t_enc + 2 * (DC_LOCK / 2);
    float* wnew = sm + S::DYN;                 // [NB][t_enc]
    float* wpad = wnew + NB * t_enc;           // [NB][2][wld]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x, nb = p.nb;
    const int unit0 = cta * DC_UNITS;
    constexpr int LIGHT_WARP = DC_THREADS / 32 - 1;   // has one LSTM row where most warps have two
    unsigned epoch = 0;

    // ---- prologue: constants and the all-zero DecoderState (src/tacotron2/mod.rs:212-236)
    for (int i = tid; i < DC_ATT * (2 * DC_LOCK); i += DC_THREADS)
        weff[(i / (2 * DC_LOCK)) * DC_WEFF_LD + i % (2 * DC_LOCK)] = p.Weff[i];
    for (int i = tid; i < DC_ATT; i += DC_THREADS) vs[i] = p.v[i];
    for (int i = tid; i < NB * 2 * wld; i += DC_THREADS) wpad[i] = 0.f;
    for (int i = tid; i < NB * ZLD; i += DC_THREADS) zs[i] = 0.f;
    for (int i = tid; i < DC_UNITS * NB; i += DC_THREADS) { cst_a[i] = 0.f; cst_d[i] = 0.f; }
    for (int i = tid; i < 4 * DC_UNITS * NB; i += DC_THREADS) {   // A1 of step 0: ctx = h_att = 0 -> bias only
        const int r = i / NB, unit = unit0 + r % DC_UNITS;
        part_a[i] = unit < DC_RNN ? p.ba[(r / DC_UNITS) * DC_RNN + unit] : 0.f;
    }
    bool done[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) done[b] = b >= nb;
    __syncthreads();

    int step = 0;
    for (; step < p.max_steps; step++) {
        const int cur = step & 1, nxt = cur ^ 1;
        // ================= stage P: prenet
        for (int i = tid; i < nb * DC_MEL; i += DC_THREADS) {
            const int b = i / DC_MEL, k = i % DC_MEL;
            xin[b * DC_MEL + k] = step ? __ldcg(p.mel_out + ((size_t)b * p.max_steps + (step - 1)) * DC_MEL + k) : 0.f;
        }
        __syncthreads();
        if (tid < DC_PRE) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            for (int k = 0; k < DC_MEL; k++) {
                const float w = __ldg(p.p1T + k * DC_PRE + tid);
#pragma unroll
                for (int b = 0; b < NB; b++) acc[b] = fmaf(w, xin[b * DC_MEL + k], acc[b]);
            }
#pragma unroll
            for (int b = 0; b < NB; b++)
                if (b < nb) x1[b * DC_PRE + tid] = fmaxf(acc[b], 0.f) * keep_scale(p, b, step, 0, tid);
        }
        __syncthreads();
        for (int r = cta + G * warp; r < DC_PRE; r += G * (DC_THREADS / 32)) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            row_dot<NB>(p.p2 + (size_t)r * DC_PRE, x1, DC_PRE, DC_PRE, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) p.x2[b * DC_PRE + r] = fmaxf(acc[b], 0.f) * keep_scale(p, b, step, 1, r);
            }
        }
        grid_barrier(p.barrier, epoch, G);

        // ================= stage A2: attention LSTM, prenet columns + cell update
        for (int i = tid; i < nb * DC_PRE; i += DC_THREADS) zs[(i / DC_PRE) * ZLD + i % DC_PRE] = __ldcg(p.x2 + i);
        __syncthreads();
        lstm_partial<NB>(p.Wa, DC_ZA, DC_ENC + DC_RNN, DC_PRE, zs, ZLD, unit0, part_a, nullptr);
        __syncthreads();
        lstm_cell<NB>(part_a, cst_a, unit0, nb, p.h_a + (size_t)nxt * nb * DC_RNN);
        grid_barrier(p.barrier, epoch, G);

        // ================= stage Q: query rows + decoder LSTM over [h_att | h_dec]
        for (int i = tid; i < nb * DC_RNN; i += DC_THREADS) {
            const int b = i / DC_RNN, k = i % DC_RNN;
            zs[b * ZLD + k] = __ldcg(p.h_a + (size_t)nxt * nb * DC_RNN + i);
            zs[b * ZLD + DC_RNN + k] = __ldcg(p.h_d + (size_t)cur * nb * DC_RNN + i);
        }
        __syncthreads();
        if (warp == LIGHT_WARP && cta < DC_ATT) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            row_dot<NB>(p.Wq + (size_t)cta * DC_RNN, zs, ZLD, DC_RNN, acc);
            if (lane == 0) {
#pragma unroll
                for (int b = 0; b < NB; b++)
                    if (b < nb) p.pq[b * DC_ATT + cta] = acc[b];
            }
        }
        lstm_partial<NB>(p.Wd, DC_ZD, 0, 2 * DC_RNN, zs, ZLD, unit0, part_d, p.bd);
        grid_barrier(p.barrier, epoch, G);

        // ================= stage E: energies e[b][t] = v . tanh(pq + Weff * [w; w_cum](t-15..t+15) + pm[t])
        for (int item = cta * (DC_THREADS / 32) + warp; item < nb * t_enc; item += G * (DC_THREADS / 32)) {
            const int b = item / t_enc, t = item % t_enc;
            if (t >= p.t_len[b]) continue;
            const float* w0 = wpad + (b * 2 + 0) * wld + t;   // padded by 15 on both sides: index t <-> position t - 15
            const float* w1 = wpad + (b * 2 + 1) * wld + t;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < DC_ATT / 32; j++) {
                const int a = lane + 32 * j;
                const float* f = weff + a * DC_WEFF_LD;
                float pa = 0.f;
#pragma unroll
                for (int k = 0; k < DC_LOCK; k++) pa = fmaf(f[k], w0[k], pa);
#pragma unroll
                for (int k = 0; k < DC_LOCK; k++) pa = fmaf(f[DC_LOCK + k], w1[k], pa);
                const float q = __ldcg(p.pq + b * DC_ATT + a);
                const float m = __ldg(p.pm + ((size_t)b * t_enc + t) * DC_ATT + a);
                s = fmaf(vs[a], tanhf(q + pa + m), s);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) p.e[b * t_enc + t] = s;
        }
        grid_barrier(p.barrier, epoch, G);

        // ================= stage C: softmax (every CTA keeps w / w_cum itself) + context chunks
        if (warp < nb) {
            const int b = warp, tl = p.t_len[b];
            float ev[DC_MAX_TENC / 32];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
                const int t = lane + 32 * j;
                ev[j] = t < tl ? __ldcg(p.e + b * t_enc + t) : -INFINITY;
                m = fmaxf(m, ev[j]);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
                ev[j] = (lane + 32 * j) < tl ? expf(ev[j] - m) : 0.f;
                sum += ev[j];
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
            for (int j = 0; j < DC_MAX_TENC / 32; j++) {
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
        for (int item = cta + G * warp; item < nb * (DC_ENC / 32); item += G * (DC_THREADS / 32)) {
            const int b = item / (DC_ENC / 32), d = (item % (DC_ENC / 32)) * 32 + lane, tl = p.t_len[b];
            const float* mp = p.memory + (size_t)b * t_enc * DC_ENC + d;
            float acc = 0.f;
            for (int t = 0; t < tl; t++) acc = fmaf(wnew[b * t_enc + t], __ldg(mp + (size_t)t * DC_ENC), acc);
            p.ctx[b * DC_ENC + d] = acc;   // read by every CTA after the barrier
        }
        grid_barrier(p.barrier, epoch, G);

        // ================= stage D2: decoder LSTM, context columns + cell update
        for (int i = tid; i < nb * DC_ENC; i += DC_THREADS)
            zs[(i / DC_ENC) * ZLD + 2 * DC_RNN + i % DC_ENC] = __ldcg(p.ctx + i);
        __syncthreads();
        lstm_partial<NB>(p.Wd, DC_ZD, 2 * DC_RNN, DC_ENC, zs + 2 * DC_RNN, ZLD, unit0, part_d, nullptr);
        __syncthreads();
        lstm_cell<NB>(part_d, cst_d, unit0, nb, p.h_d + (size_t)nxt * nb * DC_RNN);
        grid_barrier(p.barrier, epoch, G);

        // ================= stage R: projection + gate rows, and A1 of the next step
        // zs: [0, 512) ctx | [512, 1536) h_att (new) | [1536, 2560) h_dec (new)
        for (int i = tid; i < nb * DC_ENC; i += DC_THREADS) zs[(i / DC_ENC) * ZLD + i % DC_ENC] = __ldcg(p.ctx + i);
        for (int i = tid; i < nb * DC_RNN; i += DC_THREADS) {
            const int b = i / DC_RNN, k = i % DC_RNN;
            zs[b * ZLD + DC_ENC + k] = __ldcg(p.h_a + (size_t)nxt * nb * DC_RNN + i);
            zs[b * ZLD + DC_ENC + DC_RNN + k] = __ldcg(p.h_d + (size_t)nxt * nb * DC_RNN + i);
        }
        __syncthreads();
        if (warp == LIGHT_WARP && cta <= DC_MEL) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) acc[b] = 0.f;
            const float* wrow = p.Wp + (size_t)cta * DC_ZP;              // columns [ctx | h_dec]
            row_dot<NB>(wrow, zs, ZLD, DC_ENC, acc);
            row_dot<NB>(wrow + DC_ENC, zs + DC_ENC + DC_RNN, ZLD, DC_RNN, acc);
            if (lane == 0) {
                const float bias = p.bp[cta];
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (b >= nb) continue;
                    if (cta < DC_MEL) p.mel_out[((size_t)b * p.max_steps + step) * DC_MEL + cta] = acc[b] + bias;
                    else p.gate_out[(size_t)b * p.max_steps + step] = acc[b] + bias;
                }
            }
        }
        lstm_partial<NB>(p.Wa, DC_ZA, 0, DC_ENC + DC_RNN, zs, ZLD, unit0, part_a, p.ba);
        grid_barrier(p.barrier, epoch, G);

        // ================= stop rule (src/tacotron2/mod.rs:319-324): every CTA takes the same decision
        bool all = true;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (b < nb && !done[b]) {
                const float g = __ldcg(p.gate_out + (size_t)b * p.max_steps + step);
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

template <int NB>
static cudaError_t prepare_nb() {
    return cudaFuncSetAttribute(dec_persist_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)DecSmem<NB>::bytes(DC_MAX_TENC));
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
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dec_persist_kernel<8>, DC_THREADS, DecSmem<8>::bytes(DC_MAX_TENC));
    if (e != cudaSuccess) return e;
    const int need = (DC_RNN + DC_UNITS - 1) / DC_UNITS;   // CTAs that own hidden units
    if (per_sm < 1 || sms < need) return cudaErrorInvalidConfiguration;
    *grid_out = sms;   // one CTA per SM; cooperative launch keeps them co-resident
    return cudaSuccess;
}

template <int NB>
static cudaError_t launch_nb(const DecParams& p, int grid, cudaStream_t s) {
    DecParams q = p;
    void* args[1] = {&q};
    return cudaLaunchCooperativeKernel((const void*)dec_persist_kernel<NB>, dim3(grid), dim3(DC_THREADS), args,
                                       DecSmem<NB>::bytes(p.t_enc), s);
}

cudaError_t dec_launch(const DecParams& p, int grid, cudaStream_t s) {
    if (p.nb < 1 || p.nb > DC_MAX_NB || p.t_enc < 1 || p.t_enc > DC_MAX_TENC) return cudaErrorInvalidValue;
    if (p.nb == 1) return launch_nb<1>(p, grid, s);
    if (p.nb == 2) return launch_nb<2>(p, grid, s);
    if (p.nb <= 4) return launch_nb<4>(p, grid, s);
    return launch_nb<8>(p, grid, s);
}

}  // namespace xdtts
