// Griffin-Lim for the geometries the fused kernel does not cover: any hop = n_fft - noverlap in [1, n_fft] and any
// power-of-two n_fft in [64, 4096] -- the full contract of griffin_lim::GriffinLim::new(mel_basis, noverlap, power, iter,
// momentum) (/root/reference src/tacotron2/mod.rs:456; the shipped call passes noverlap = 768 of n_fft = 1024, which takes
// gl_iter_kernel).  Same arithmetic as the fused path and the oracle (librosa 0.9.2 stft / istft / griffinlim: centred
// frames, reflect or zero padding, periodic Hann, overlap-add in ascending frame order divided by the window sum-square
// where > tiny, alpha = m / (1 + m), eps = tiny), un-fused: per iteration
//     glg_frame_kernel   one CTA per frame: gather + window -> forward FFT (radix-4 Stockham in shared memory, the real transform
//                        as a half-length complex one) -> momentum / projection onto S -> inverse FFT -> window -> the frame's samples
//     glg_ola_kernel     one thread per output sample: the frames that cover it, ascending, / window sum-square
// and n_iter + 1 of each per vocode.  The per-frame state lives in the same records as the fused path's (R packed with
// (Re R[0], Re R[M]) in slot 0, then S, then S_nyq), so the lift, the magnitude / phase transposes, peek and the download
// are shared.  A general hop has no fixed number of frames per sample, which is what the fused kernel's register
// overlap-add is built on; this path trades ~2.5x the HBM traffic and a shared-memory FFT for having no such assumption
// (cfg2 batch: 22.8 ms per 60-iteration step against the fused kernel's 4.8).
#include <cuda_runtime.h>
#include <math.h>

#include <vector>

#include "gl_core.cuh"
#include "gl_generic.h"

namespace xdtts {

namespace {

constexpr int GG_THREADS = 256;   // overlap-add, finish
constexpr int GF_THREADS = 64;    // one frame per CTA; 32 CTAs per SM: a frame is a chain of ~16 CTA-wide barriers, concurrency hides it

// Stockham autosort FFT of m = 2^k complex points in shared memory, radix 4 (one radix-2 stage last when k is odd), ping-pong
// between a and b; returns the buffer that holds the result (natural order in, natural order out).  tw[j] = exp(-2 pi i j / (2 m)),
// j < m, i.e. the table of the REAL transform of length n = 2 m that this complex one is half of; INV conjugates (no 1 / m).
// Radix 4 halves the passes through shared memory of a radix-2 form, and the real transform as a half-length complex one
// halves them again: the un-fused path is bound by exactly that traffic.
template <bool INV>
__device__ float2* fft_pow2(float2* a, float2* b, int m, const float2* __restrict__ tw) {
    int lg = 0;   // log2 of ns
    for (int ns = 1; ns < m;) {
        if ((m / ns) % 4 == 0) {
            const int q = m >> 2, s = m / (4 * ns);
            for (int j = threadIdx.x; j < q; j += GF_THREADS) {
                const int k = j & (ns - 1);
                float2 w1 = __ldg(&tw[2 * k * s]);   // exp(-2 pi i k / (4 ns))
                if (INV) w1.y = -w1.y;
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                const float2 v0 = a[j], v1 = cmul(a[j + q], w1), v2 = cmul(a[j + 2 * q], w2), v3 = cmul(a[j + 3 * q], w3);
                const float2 t0 = mk2(v0.x + v2.x, v0.y + v2.y), t1 = mk2(v0.x - v2.x, v0.y - v2.y);
                const float2 t2 = mk2(v1.x + v3.x, v1.y + v3.y), d = mk2(v1.x - v3.x, v1.y - v3.y);
                const float2 t3 = INV ? mk2(-d.y, d.x) : mk2(d.y, -d.x);   // (v1 - v3) * (+-i)
                const int j0 = ((j >> lg) << (lg + 2)) + k;
                b[j0] = mk2(t0.x + t2.x, t0.y + t2.y);
                b[j0 + ns] = mk2(t1.x + t3.x, t1.y + t3.y);
                b[j0 + 2 * ns] = mk2(t0.x - t2.x, t0.y - t2.y);
                b[j0 + 3 * ns] = mk2(t1.x - t3.x, t1.y - t3.y);
            }
            ns <<= 2;
            lg += 2;
        } else {   // the last stage of an odd power of two: ns = m / 2, k = j
            const int h = m >> 1;
            for (int j = threadIdx.x; j < h; j += GF_THREADS) {
                float2 w = __ldg(&tw[2 * j]);
                if (INV) w.y = -w.y;
                const float2 v0 = a[j], v1 = cmul(a[j + h], w);
                b[j] = mk2(v0.x + v1.x, v0.y + v1.y);
                b[j + h] = mk2(v0.x - v1.x, v0.y - v1.y);
            }
            ns <<= 1;
            lg += 1;
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
    }
    return a;
}

// index into a signal of `len` samples for the centred frame's sample j (may be outside [0, len)): numpy 'reflect'
// (no edge repeat, as many folds as it takes), or -1 = zero for constant padding / an empty signal
__device__ __forceinline__ int padded_index(int j, int len, int pad_mode) {
    if (j >= 0 && j < len) return j;
    if (pad_mode != GL_PAD_REFLECT || len < 1) return -1;
    if (len == 1) return 0;
    const int period = 2 * (len - 1);
    j %= period;
    if (j < 0) j += period;
    return j < len ? j : period - j;
}

// One frame per CTA.  The real transform of n samples is the complex transform of m = n / 2 points z[i] = x[2i] + i x[2i+1]
// followed by the split X[k] = (Z[k] + conj Z[m-k]) / 2 - (i / 2) W_n^k (Z[k] - conj Z[m-k]); the inverse merges
// Z[k] = (Y[k] + conj Y[m-k]) + i conj(W_n^k) (Y[k] - conj Y[m-k]) and transforms back (x n in total, folded into 1 / n).
template <int MODE>   // 0: initial spectrum S e^{i phase}; 1: one iteration from the waveform
__global__ void __launch_bounds__(GF_THREADS) glg_frame_kernel(const GlgParams p) {
    extern __shared__ __align__(16) float2 gsm[];
    const int N = p.n_fft, M = N >> 1;
    float2* a = gsm;                // M + 1 entries each
    float2* b = gsm + (M + 1);
    __shared__ int s_u;
    __shared__ float s_dc, s_nyq;
    const int f = blockIdx.x;   // frame row of the batch
    if (threadIdx.x < 32) {   // the utterance of frame f: the last one that starts at or before it -- 32 offsets per load latency, not one
        int u = 0;
        for (int base = 1; base < p.n_utt; base += 32) {
            const int i = base + (int)threadIdx.x;
            const unsigned m = __ballot_sync(0xffffffffu, i < p.n_utt && __ldg(p.utt_foff + i) <= f);
            u += __popc(m);
            if (m != 0xffffffffu) break;
        }
        if (threadIdx.x == 0) s_u = u;
    }
    __syncthreads();
    const int u = s_u, T = p.utt_T[u], foff = p.utt_foff[u], t = f - foff;
    float* rec = p.state + (size_t)f * p.rec_f;
    float2* R = reinterpret_cast<float2*>(rec);   // previous rebuilt spectrum, slot 0 = (Re R[0], Re R[M])
    const float* S = rec + 2 * M;                 // S[0..M-1], S[M] = the Nyquist magnitude
    float2* Y;                                    // Y[0..M]: the spectrum to invert (DC and Nyquist real)
    if (MODE == 0) {
        Y = b;
        const unsigned long long seed = *p.seed;
        const int sid = p.utt_seed_id[u];
        for (int k = threadIdx.x; k <= M; k += GF_THREADS) {
            const float turn = p.turns ? p.turns[(size_t)f * (M + 1) + k] : phase_turn(seed, sid, M + 1, t, k);
            float sn, cs;
            sincos_turns(turn, &sn, &cs);
            const float mag = S[k];
            Y[k] = mk2(mag * cs, (k == 0 || k == M) ? 0.f : mag * sn);   // irfft reads only the real part of DC / Nyquist
            if (k < M) R[k] = mk2(0.f, 0.f);   // tprev of the first iteration is zero
        }
        __syncthreads();
    } else {
        const int len = p.hop * (T - 1);
        const float* y = p.y + (size_t)foff * p.hop;
        for (int i = threadIdx.x; i < M; i += GF_THREADS) {
            const int i0 = padded_index(t * p.hop + 2 * i - M, len, p.pad_mode), i1 = padded_index(t * p.hop + 2 * i + 1 - M, len, p.pad_mode);
            a[i] = mk2(i0 < 0 ? 0.f : y[i0] * p.win[2 * i], i1 < 0 ? 0.f : y[i1] * p.win[2 * i + 1]);
        }
        __syncthreads();
        float2* Z = fft_pow2<false>(a, b, M, p.tw);
        Y = Z == a ? b : a;
        for (int k = threadIdx.x; k <= M; k += GF_THREADS) {
            const float2 zk = Z[k == M ? 0 : k], zm = Z[k == 0 ? 0 : M - k];
            const float2 zc = mk2(zm.x, -zm.y);
            const float2 w = k == M ? mk2(-1.f, 0.f) : __ldg(&p.tw[k]);   // W_n^k
            const float2 e = mk2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y)), o = mk2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            const float2 wo = cmul(w, o);
            float2 rn = mk2(e.x + wo.y, e.y - wo.x);   // e - i w o
            if (k == 0 || k == M) rn.y = 0.f;          // a real signal's DC / Nyquist bins are real
            float2 rp;
            if (k == 0) rp = mk2(R[0].x, 0.f);
            else if (k == M) rp = mk2(R[0].y, 0.f);
            else rp = R[k];
            const float2 uu = mk2(rn.x - p.alpha * rp.x, rn.y - p.alpha * rp.y);
            const float g = S[k] / (sqrtf(uu.x * uu.x + uu.y * uu.y) + 1.17549435e-38f);
            Y[k] = mk2(g * uu.x, (k == 0 || k == M) ? 0.f : g * uu.y);
            if (k == 0) s_dc = rn.x;
            else if (k == M) s_nyq = rn.x;
            else R[k] = rn;
        }
        __syncthreads();   // every Z and R[0] has been read, every Y written
        if (threadIdx.x == 0) R[0] = mk2(s_dc, s_nyq);
    }
    float2* Zi = Y == a ? b : a;
    for (int k = threadIdx.x; k < M; k += GF_THREADS) {
        const float2 yk = Y[k], ym = Y[M - k];
        const float2 yc = mk2(ym.x, -ym.y);
        float2 w = __ldg(&p.tw[k]);
        w.y = -w.y;                                // conj(W_n^k)
        const float2 e = mk2(yk.x + yc.x, yk.y + yc.y), o = mk2(yk.x - yc.x, yk.y - yc.y);
        const float2 wo = cmul(w, o);
        Zi[k] = mk2(e.x - wo.y, e.y + wo.x);       // e + i w o
    }
    __syncthreads();
    float2* z = fft_pow2<true>(Zi, Zi == a ? b : a, M, p.tw);
    float2* out = reinterpret_cast<float2*>(p.frames + (size_t)f * N);
    const float inv_n = 1.0f / (float)N;
    for (int i = threadIdx.x; i < M; i += GF_THREADS) out[i] = mk2(z[i].x * inv_n * p.win[2 * i], z[i].y * inv_n * p.win[2 * i + 1]);
}

// y[j] of every utterance: sum of the frames that cover padded position n = j + n_fft / 2, ascending frame order
__global__ void __launch_bounds__(GG_THREADS) glg_ola_kernel(const GlgParams p, int last) {
    const int u = blockIdx.y, T = p.utt_T[u], foff = p.utt_foff[u];
    const int N = p.n_fft, hop = p.hop, len = hop * (T - 1);
    const float* fr = p.frames + (size_t)foff * N;
    float* y = p.y + (size_t)foff * hop;
    float amax = 0.f;
    for (int j = blockIdx.x * GG_THREADS + threadIdx.x; j < len; j += gridDim.x * GG_THREADS) {
        const int n = j + (N >> 1);
        const int t_hi = min(T - 1, n / hop);
        const int t_lo = n - N + 1 > 0 ? (n - N + hop) / hop : 0;
        float acc = 0.f, wss = 0.f;
        for (int t = t_lo; t <= t_hi; t++) {
            const int o = n - t * hop;
            acc += fr[(size_t)t * N + o];
            const float w = p.win[o];
            wss += w * w;
        }
        const float v = wss > 1.17549435e-38f ? acc / wss : acc;
        y[j] = v;
        amax = fmaxf(amax, fabsf(v));
    }
    if (last) {
#pragma unroll
        for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(p.amax + u, __float_as_uint(amax));
    }
}

// out arena <- y, peak-normalised (the scalar form of gl_finish_kernel: a general hop gives no 16-byte alignment)
__global__ void __launch_bounds__(GG_THREADS) glg_finish_kernel(const float* __restrict__ y, const int* __restrict__ utt_T,
                                                                 const int* __restrict__ utt_foff, const long long* __restrict__ out_off,
                                                                 const unsigned* __restrict__ amax, int hop, int normalise,
                                                                 float* __restrict__ out) {
    const int u = blockIdx.y;
    const long len = (long)hop * (utt_T[u] - 1);
    const float* src = y + (long)utt_foff[u] * hop;
    float* dst = out + out_off[u];
    const float m = __uint_as_float(amax[u]);
    const bool norm = normalise && m > 0.f;
    for (long i = (long)blockIdx.x * GG_THREADS + threadIdx.x; i < len; i += (long)gridDim.x * GG_THREADS) dst[i] = norm ? src[i] / m : src[i];
}

}  // namespace

std::vector<float2> glg_build_twiddles(int n_fft) {
    std::vector<float2> tw(n_fft / 2);
    for (int j = 0; j < n_fft / 2; j++) {
        const double th = -2.0 * M_PI * (double)j / (double)n_fft;
        tw[j] = make_float2((float)cos(th), (float)sin(th));
    }
    return tw;
}

std::vector<float> glg_build_window(int n_fft) {   // periodic Hann, rounded to fp32 from fp64 like the oracle's hann_periodic
    std::vector<float> w(n_fft);
    for (int i = 0; i < n_fft; i++) w[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)n_fft));
    return w;
}

cudaError_t glg_prepare(int n_fft) {
    const int bytes = 2 * (n_fft / 2 + 1) * (int)sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(glg_frame_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(glg_frame_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

cudaError_t glg_launch_frames(const GlgParams& p, int mode, int total_frames, cudaStream_t s) {
    const size_t sm = 2 * ((size_t)p.n_fft / 2 + 1) * sizeof(float2);
    if (mode == 0) glg_frame_kernel<0><<<total_frames, GF_THREADS, sm, s>>>(p);
    else glg_frame_kernel<1><<<total_frames, GF_THREADS, sm, s>>>(p);
    return cudaGetLastError();
}

cudaError_t glg_launch_ola(const GlgParams& p, bool last, int max_T, cudaStream_t s) {
    long blocks = ((long)p.hop * max_T + GG_THREADS - 1) / GG_THREADS;
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    glg_ola_kernel<<<dim3((unsigned)blocks, p.n_utt), GG_THREADS, 0, s>>>(p, last ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t glg_launch_finish(const float* y, const int* utt_T, const int* utt_foff, const long long* out_off, const unsigned* amax,
                              int n_utt, int max_T, int hop, int normalise, float* out, cudaStream_t s) {
    long blocks = ((long)hop * max_T + GG_THREADS - 1) / GG_THREADS;
    if (blocks > 256) blocks = 256;
    if (blocks < 1) blocks = 1;
    glg_finish_kernel<<<dim3((unsigned)blocks, n_utt), GG_THREADS, 0, s>>>(y, utt_T, utt_foff, out_off, amax, hop, normalise, out);
    return cudaGetLastError();
}

}  // namespace xdtts
