// Griffin-Lim for the geometries the fused kernel does not cover: any hop = n_fft - noverlap in [1, n_fft] and any
// power-of-two n_fft in [64, 4096] -- the full contract of griffin_lim::GriffinLim::new(mel_basis, noverlap, power, iter,
// momentum) (/root/reference src/tacotron2/mod.rs:456; the shipped call passes noverlap = 768 of n_fft = 1024, which takes
// gl_iter_kernel).  Same arithmetic as the fused path and the oracle (librosa 0.9.2 stft / istft / griffinlim: centred
// frames, reflect or zero padding, periodic Hann, overlap-add in ascending frame order divided by the window sum-square
// where > tiny, alpha = m / (1 + m), eps = tiny), un-fused: per iteration
//     glg_frame_kernel   one CTA per frame: gather + window -> forward FFT -> momentum / projection onto S -> inverse FFT
//                        -> window -> the frame's n_fft samples
//     glg_ola_kernel     one thread per output sample: the frames that cover it, ascending, / window sum-square
// and n_iter + 1 of each per vocode.  The per-frame state lives in the same records as the fused path's (R packed with
// (Re R[0], Re R[M]) in slot 0, then S, then S_nyq), so the lift, the magnitude / phase transposes, peek and the download
// are shared.  A general hop has no fixed number of frames per sample, which is what the fused kernel's register
// overlap-add is built on; this path trades ~6x the HBM traffic for having no such assumption.
#include <cuda_runtime.h>
#include <math.h>

#include <vector>

#include "gl_core.cuh"
#include "gl_generic.h"

namespace xdtts {

namespace {

constexpr int GG_THREADS = 256;

// Stockham radix-2 autosort FFT of n = 2^logn points in shared memory, ping-pong between a and b; returns the buffer
// that holds the result (natural order).  tw[j] = exp(-2 pi i j / n), j < n / 2; INV conjugates (no 1/n).
template <bool INV>
__device__ float2* fft_pow2(float2* a, float2* b, int n, int logn, const float2* __restrict__ tw) {
    const int half = n >> 1;
    for (int s = 0; s < logn; s++) {
        const int m = 1 << s;   // butterflies of one group; n / (2 m) groups
        for (int i = threadIdx.x; i < half; i += GG_THREADS) {
            const int j = i >> s, k = i & (m - 1);
            const float2 c0 = a[k + j * m], c1 = a[k + j * m + half];
            float2 w = __ldg(&tw[j * m]);
            if (INV) w.y = -w.y;
            const float2 d = mk2(c0.x - c1.x, c0.y - c1.y);
            b[k + 2 * j * m] = mk2(c0.x + c1.x, c0.y + c1.y);
            b[k + 2 * j * m + m] = mk2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
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

template <int MODE>   // 0: initial spectrum S e^{i phase}; 1: one iteration from the waveform
__global__ void __launch_bounds__(GG_THREADS) glg_frame_kernel(const GlgParams p) {
    extern __shared__ __align__(16) float2 gsm[];
    const int N = p.n_fft, M = N >> 1;
    float2* a = gsm;
    float2* b = gsm + N;
    __shared__ int s_u;
    const int f = blockIdx.x;   // frame row of the batch
    if (threadIdx.x == 0) {
        int u = 0;
        while (u + 1 < p.n_utt && p.utt_foff[u + 1] <= f) u++;
        s_u = u;
    }
    __syncthreads();
    const int u = s_u, T = p.utt_T[u], foff = p.utt_foff[u], t = f - foff;
    float* rec = p.state + (size_t)f * p.rec_f;
    float2* R = reinterpret_cast<float2*>(rec);   // previous rebuilt spectrum, slot 0 = (Re R[0], Re R[M])
    const float* S = rec + 2 * M;                 // S[0..M-1], S[M] = the Nyquist magnitude
    float2* Z;                                    // the full Hermitian spectrum to invert
    if (MODE == 0) {
        Z = a;
        const unsigned long long seed = *p.seed;
        const int sid = p.utt_seed_id[u];
        for (int k = threadIdx.x; k <= M; k += GG_THREADS) {
            const float turn = p.turns ? p.turns[(size_t)f * (M + 1) + k] : phase_turn(seed, sid, M + 1, t, k);
            float sn, cs;
            sincos_turns(turn, &sn, &cs);
            const float mag = S[k];
            const float2 y = mk2(mag * cs, (k == 0 || k == M) ? 0.f : mag * sn);   // irfft reads only the real part of DC / Nyquist
            Z[k] = y;
            if (k > 0 && k < M) Z[N - k] = mk2(y.x, -y.y);
            if (k < M) R[k] = mk2(0.f, 0.f);   // tprev of the first iteration is zero
        }
        __syncthreads();
    } else {
        const int len = p.hop * (T - 1);
        const float* y = p.y + (size_t)foff * p.hop;
        for (int n = threadIdx.x; n < N; n += GG_THREADS) {
            const int idx = padded_index(t * p.hop + n - M, len, p.pad_mode);
            a[n] = mk2(idx < 0 ? 0.f : y[idx] * p.win[n], 0.f);
        }
        __syncthreads();
        float2* X = fft_pow2<false>(a, b, N, p.log_n, p.tw);
        Z = X == a ? b : a;
        for (int k = threadIdx.x; k <= M; k += GG_THREADS) {
            float2 rn = X[k];
            if (k == 0 || k == M) rn.y = 0.f;   // a real signal's DC / Nyquist bins are real (rfft returns exactly 0 there)
            float2 rp;
            if (k == 0) rp = mk2(R[0].x, 0.f);
            else if (k == M) rp = mk2(R[0].y, 0.f);
            else rp = R[k];
            const float2 uu = mk2(rn.x - p.alpha * rp.x, rn.y - p.alpha * rp.y);
            const float g = S[k] / (sqrtf(uu.x * uu.x + uu.y * uu.y) + 1.17549435e-38f);
            const float2 yk = mk2(g * uu.x, g * uu.y);
            Z[k] = (k == 0 || k == M) ? mk2(yk.x, 0.f) : yk;
            if (k > 0 && k < M) {
                Z[N - k] = mk2(yk.x, -yk.y);
                R[k] = rn;
            }
        }
        __syncthreads();   // every X[0] / X[M] has been read
        if (threadIdx.x == 0) R[0] = mk2(X[0].x, X[M].x);
        __syncthreads();
    }
    float2* z = fft_pow2<true>(Z, Z == a ? b : a, N, p.log_n, p.tw);
    float* out = p.frames + (size_t)f * N;
    const float inv_n = 1.0f / (float)N;
    for (int n = threadIdx.x; n < N; n += GG_THREADS) out[n] = z[n].x * inv_n * p.win[n];
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
    const int bytes = 2 * n_fft * (int)sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(glg_frame_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(glg_frame_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}

cudaError_t glg_launch_frames(const GlgParams& p, int mode, int total_frames, cudaStream_t s) {
    const size_t sm = 2 * (size_t)p.n_fft * sizeof(float2);
    if (mode == 0) glg_frame_kernel<0><<<total_frames, GG_THREADS, sm, s>>>(p);
    else glg_frame_kernel<1><<<total_frames, GG_THREADS, sm, s>>>(p);
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
