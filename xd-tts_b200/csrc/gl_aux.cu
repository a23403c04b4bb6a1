// Kernels either side of the Griffin-Lim iteration loop (all HBM-bound, one pass each):
//   * the non-negative least-squares refinement of the mel -> linear lift (the lift itself is gl_lift.cu)
//   * [K,T] row-major (the reference's ndarray layout) -> frame-major [T][M] + Nyquist column
//   * peak normalisation of the final waveform (row a8; the caller scales by i16::MAX,
//     src/lib.rs:153-155)
#include <cuda_runtime.h>

#include "gl_host.h"

namespace xdtts {

constexpr int LIFT_MAX_MELS = 256;

// ---------------------------------------------------------------- NNLS lift
// The crate behind griffin_lim::GriffinLim::infer ports librosa, whose mel -> linear step is
// `nnls(mel_basis, mel)`: start from the clipped least-squares solution and minimise
// 0.5 |A x - b|^2 over x >= 0 (librosa.util.nnls).  librosa hands that to L-BFGS-B; the
// minimiser is not unique (80 equations, 513 unknowns), so there is no trajectory to be bit-faithful to -- what is
// defined is the optimisation problem and its starting point.  Here every frame is solved by accelerated
// projected gradient (FISTA, step 1/L with L = sigma_max(A)^2, gradient-based adaptive restart), a fixed,
// deterministic recurrence: one warp per frame, x and the extrapolated point y in registers (bin k = lane + 32 j),
// the mel residual in shared memory.  A is used in both sparse forms (rows for A y, columns for A^T r): a mel
// filterbank has <= 2 non-zeros per column, so an iteration costs ~1.5 k multiply-adds instead of 82 k.
// In: S / S_nyq hold x0 = max(0, pinv(A) b) (gl_lift_kernel with power 1).  Out: S = x^power, in place.
struct NnlsMat {
    const int* row_ptr;    // [n_mels + 1]
    const int* row_col;    // [nnz] bin of each non-zero, row by row
    const float* row_val;
    const int* col_ptr;    // [K + 1]
    const int* col_row;    // [nnz] mel row of each non-zero, column by column
    const float* col_val;
};

template <int KJ>
__global__ void __launch_bounds__(128) gl_nnls_kernel(const float* __restrict__ mel_arena, NnlsMat A, const int* __restrict__ utt_T,
                                                      const int* __restrict__ utt_foff, int n_mels, int K, float power, int delog,
                                                      float inv_L, int max_iter, float pgtol_over_L, float* __restrict__ S,
                                                      int ld) {
    extern __shared__ __align__(16) float nn_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.y, T = utt_T[u], t = blockIdx.x * 4 + warp;
    const int M = K - 1, kpad = 32 * KJ;
    if (t >= T) return;
    const long foff = utt_foff[u];
    float* ys = nn_sm + warp * (kpad + 2 * LIFT_MAX_MELS);   // [kpad] extrapolated point
    float* rs = ys + kpad;                                   // [n_mels] residual A y - b
    float* bs = rs + LIFT_MAX_MELS;                          // [n_mels] de-logged mel frame
    const float* mel = mel_arena + foff * n_mels;
    for (int m = lane; m < n_mels; m += 32) {
        const float v = mel[(long)m * T + t];
        bs[m] = delog == 0 ? expf(v) : (delog == 1 ? powf(10.f, v) : v);
    }
    float x[KJ], y[KJ];
    const long frame = foff + t;
#pragma unroll
    for (int j = 0; j < KJ; j++) {
        const int k = lane + 32 * j;
        x[j] = k <= M ? S[frame * ld + k] : 0.f;   // k == M: the Nyquist slot of the frame's record
        y[j] = x[j];
    }
    float tk = 1.f;
    for (int it = 0; it < max_iter; it++) {
#pragma unroll
        for (int j = 0; j < KJ; j++) ys[lane + 32 * j] = y[j];
        __syncwarp();
        for (int m = lane; m < n_mels; m += 32) {
            float acc = -bs[m];
            const int p1 = A.row_ptr[m + 1];
            for (int p = A.row_ptr[m]; p < p1; p++) acc = fmaf(__ldg(A.row_val + p), ys[__ldg(A.row_col + p)], acc);
            rs[m] = acc;
        }
        __syncwarp();
        float dot = 0.f, step = 0.f;
#pragma unroll
        for (int j = 0; j < KJ; j++) {
            const int k = lane + 32 * j;
            if (k < K) {
                float g = 0.f;
                const int p1 = A.col_ptr[k + 1];
                for (int p = A.col_ptr[k]; p < p1; p++) g = fmaf(__ldg(A.col_val + p), rs[__ldg(A.col_row + p)], g);
                const float xn = fmaxf(y[j] - g * inv_L, 0.f);
                dot = fmaf(g, xn - x[j], dot);
                step = fmaxf(step, fabsf(xn - y[j]));
                y[j] = xn;   // x[j] keeps the previous iterate until the momentum is known
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            step = fmaxf(step, __shfl_xor_sync(0xffffffffu, step, o));
        }
        // gradient restart (O'Donoghue & Candes): drop the momentum when it points uphill
        float beta = 0.f;
        if (dot > 0.f) {
            tk = 1.f;
        } else {
            const float tn = 0.5f * (1.f + sqrtf(1.f + 4.f * tk * tk));
            beta = (tk - 1.f) / tn;
            tk = tn;
        }
#pragma unroll
        for (int j = 0; j < KJ; j++) {
            const float xn = y[j];
            y[j] = fmaf(beta, xn - x[j], xn);
            x[j] = xn;
        }
        __syncwarp();
        if (step < pgtol_over_L) break;   // projected-gradient step below pgtol / L (warp-uniform)
    }
#pragma unroll
    for (int j = 0; j < KJ; j++) {
        const int k = lane + 32 * j;
        const float v = x[j] > 0.f ? (power == 1.0f ? x[j] : powf(x[j], power)) : 0.f;
        if (k <= M) S[frame * ld + k] = v;
    }
}

// Banded form of the same solver, used when every row of the basis is a short run of consecutive bins and every
// column has at most 4 non-zeros -- i.e. for a mel filterbank (triangles: rows of <= 27 bins at n_fft 1024, columns
// of <= 2 filters).  Rows are stored as dense bands, transposed ([i][m]: lanes read consecutive addresses), columns
// in ELL form ([c][k]); both loops have fixed trip counts, so the loads are independent of each other and the
// multiply-adds run as four interleaved chains instead of one pointer-chasing loop per row (measured on B200:
// 15.7 -> 11.6 ms for 32 x 1000 worst-case frames at n_fft 1024, 104 -> 50 ms for 8 x 8000 at n_fft 2048).
struct NnlsBand {
    const int* row_lo;      // [n_mels] first bin of the row's band
    const float* bandT;     // [rw][n_mels] A[m][row_lo[m] + i], zero past the band
    const int* col_row;     // [CW][K] mel row of the c-th non-zero of column k (0 when absent)
    const float* col_val;   // [CW][K] its value (0 when absent)
    int rw;                 // band width, a multiple of 4
};

template <int KJ, int CW>
__global__ void __launch_bounds__(128) gl_nnls_band_kernel(const float* __restrict__ mel_arena, NnlsBand A, const int* __restrict__ utt_T,
                                                           const int* __restrict__ utt_foff, int n_mels, int K, float power, int delog,
                                                           float inv_L, int max_iter, float pgtol_over_L, float* __restrict__ S,
                                                           int ld) {
    extern __shared__ __align__(16) float nn_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.y, T = utt_T[u], t = blockIdx.x * 4 + warp;
    if (t >= T) return;
    const long foff = utt_foff[u];
    const int M = K - 1;
    constexpr int YS = 32 * KJ + 64;                         // the extrapolated point, zero padded past K for the bands
    float* ys = nn_sm + warp * (YS + 2 * LIFT_MAX_MELS);
    float* rs = ys + YS;
    float* bs = rs + LIFT_MAX_MELS;
    const float* mel = mel_arena + foff * n_mels;
    for (int m = lane; m < n_mels; m += 32) {
        const float v = mel[(long)m * T + t];
        bs[m] = delog == 0 ? expf(v) : (delog == 1 ? powf(10.f, v) : v);
    }
    ys[32 * KJ + lane] = 0.f;
    ys[32 * KJ + 32 + lane] = 0.f;
    float x[KJ], y[KJ];
    const long frame = foff + t;
#pragma unroll
    for (int j = 0; j < KJ; j++) {
        const int k = lane + 32 * j;
        x[j] = k <= M ? S[frame * ld + k] : 0.f;   // k == M: the Nyquist slot of the frame's record
        y[j] = x[j];
    }
    float tk = 1.f;
    for (int it = 0; it < max_iter; it++) {
#pragma unroll
        for (int j = 0; j < KJ; j++) ys[lane + 32 * j] = y[j];
        __syncwarp();
        for (int m = lane; m < n_mels; m += 32) {
            const float* yb = ys + __ldg(A.row_lo + m);
            const float* ab = A.bandT + m;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int i = 0; i < A.rw; i += 4) {
                a0 = fmaf(__ldg(ab + (size_t)(i + 0) * n_mels), yb[i + 0], a0);
                a1 = fmaf(__ldg(ab + (size_t)(i + 1) * n_mels), yb[i + 1], a1);
                a2 = fmaf(__ldg(ab + (size_t)(i + 2) * n_mels), yb[i + 2], a2);
                a3 = fmaf(__ldg(ab + (size_t)(i + 3) * n_mels), yb[i + 3], a3);
            }
            rs[m] = ((a0 + a1) + (a2 + a3)) - bs[m];
        }
        __syncwarp();
        float dot = 0.f, step = 0.f;
#pragma unroll
        for (int j = 0; j < KJ; j++) {
            const int k = lane + 32 * j;
            if (k < K) {
                float g = 0.f;
#pragma unroll
                for (int c = 0; c < CW; c++) g = fmaf(__ldg(A.col_val + (size_t)c * K + k), rs[__ldg(A.col_row + (size_t)c * K + k)], g);
                const float xn = fmaxf(y[j] - g * inv_L, 0.f);
                dot = fmaf(g, xn - x[j], dot);
                step = fmaxf(step, fabsf(xn - y[j]));
                y[j] = xn;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            step = fmaxf(step, __shfl_xor_sync(0xffffffffu, step, o));
        }
        float beta = 0.f;
        if (dot > 0.f) {
            tk = 1.f;
        } else {
            const float tn = 0.5f * (1.f + sqrtf(1.f + 4.f * tk * tk));
            beta = (tk - 1.f) / tn;
            tk = tn;
        }
#pragma unroll
        for (int j = 0; j < KJ; j++) {
            const float xn = y[j];
            y[j] = fmaf(beta, xn - x[j], xn);
            x[j] = xn;
        }
        __syncwarp();
        if (step < pgtol_over_L) break;
    }
#pragma unroll
    for (int j = 0; j < KJ; j++) {
        const int k = lane + 32 * j;
        const float v = x[j] > 0.f ? (power == 1.0f ? x[j] : powf(x[j], power)) : 0.f;
        if (k <= M) S[frame * ld + k] = v;
    }
}

// band: [n_mels] row_lo | then nothing else (ints); band_val: [rw][n_mels]; colrow: [cw][K]; colval: [cw][K]
cudaError_t gl_launch_nnls_band(const float* mel_arena, const int* row_lo, const float* bandT, int rw, const int* col_row,
                                const float* col_val, int cw, const int* utt_T, const int* utt_foff, int n_utt, int max_T, int n_mels,
                                int K, float power, int delog, float L, int max_iter, float pgtol, float* S, int ld,
                                cudaStream_t s) {
    if (n_mels > LIFT_MAX_MELS || cw < 1 || cw > 4 || rw < 4 || rw > 64 || (rw & 3)) return cudaErrorInvalidValue;
    NnlsBand A;
    A.row_lo = row_lo; A.bandT = bandT; A.col_row = col_row; A.col_val = col_val; A.rw = rw;
    dim3 grid((max_T + 3) / 4, n_utt);
    const int kj = (K + 31) / 32;
    const float inv_L = 1.0f / L, tol = pgtol / L;
#define XD_BAND_LAUNCH(KJ_, CW_)                                                                                                  \
    gl_nnls_band_kernel<KJ_, CW_><<<grid, 128, 4 * (size_t)(32 * KJ_ + 64 + 2 * LIFT_MAX_MELS) * sizeof(float), s>>>(             \
        mel_arena, A, utt_T, utt_foff, n_mels, K, power, delog, inv_L, max_iter, tol, S, ld)
#define XD_BAND_CW(KJ_)                                                                                                           \
    do {                                                                                                                          \
        if (cw <= 2) XD_BAND_LAUNCH(KJ_, 2);                                                                                      \
        else XD_BAND_LAUNCH(KJ_, 4);                                                                                              \
    } while (0)
    if (kj <= 9) XD_BAND_CW(9);
    else if (kj <= 17) XD_BAND_CW(17);
    else if (kj <= 33) XD_BAND_CW(33);
    else return cudaErrorInvalidValue;
#undef XD_BAND_CW
#undef XD_BAND_LAUNCH
    return cudaGetLastError();
}

cudaError_t gl_launch_nnls(const float* mel_arena, const int* csr, const float* csr_val, const int* csc, const float* csc_val,
                           const int* utt_T, const int* utt_foff, int n_utt, int max_T, int n_mels, int K, float power, int delog,
                           float L, int max_iter, float pgtol, float* S, int ld, cudaStream_t s) {
    if (n_mels > LIFT_MAX_MELS) return cudaErrorInvalidValue;
    NnlsMat A;
    A.row_ptr = csr; A.row_col = csr + n_mels + 1; A.row_val = csr_val;
    A.col_ptr = csc; A.col_row = csc + K + 1; A.col_val = csc_val;
    dim3 grid((max_T + 3) / 4, n_utt);
    const int kj = (K + 31) / 32;
    const float inv_L = 1.0f / L, tol = pgtol / L;
#define XD_NNLS_LAUNCH(KJ_)                                                                                                   \
    gl_nnls_kernel<KJ_><<<grid, 128, 4 * (size_t)(32 * KJ_ + 2 * LIFT_MAX_MELS) * sizeof(float), s>>>(                          \
        mel_arena, A, utt_T, utt_foff, n_mels, K, power, delog, inv_L, max_iter, tol, S, ld)
    if (kj <= 9) XD_NNLS_LAUNCH(9);
    else if (kj <= 17) XD_NNLS_LAUNCH(17);
    else if (kj <= 33) XD_NNLS_LAUNCH(33);
    else return cudaErrorInvalidValue;
#undef XD_NNLS_LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------- [K,T] -> frame-major
// src arena: utterance u is a row-major [K][T_u] block at float offset foff[u]*K.  dst: frame-major rows of `ld` floats
// (bin k of frame f at dst[f * ld + k], k = 0..K-1).
__global__ void __launch_bounds__(256) gl_to_frame_major_kernel(const float* __restrict__ src_arena,
                                                                const int* __restrict__ utt_T,
                                                                const int* __restrict__ utt_foff, int K,
                                                                float* __restrict__ dst, int ld) {
    __shared__ float tile[32][33];
    const int u = blockIdx.z;
    const int T = utt_T[u];
    const int t0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    if (t0 >= T) return;
    const long foff = utt_foff[u];
    const float* src = src_arena + foff * K;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, t = t0 + tx;
        tile[r][tx] = (k < K && t < T) ? src[(long)k * T + t] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, k = k0 + tx;
        if (t < T && k < K) dst[(foff + t) * ld + k] = tile[tx][r];   // k == K-1 lands in the Nyquist slot
    }
}

cudaError_t gl_launch_to_frame_major(const float* src_arena, const int* utt_T, const int* utt_foff, int n_utt, int max_T,
                                     int K, float* dst, int ld, cudaStream_t s) {
    dim3 grid((K + 31) / 32, (max_T + 31) / 32, n_utt);
    gl_to_frame_major_kernel<<<grid, 256, 0, s>>>(src_arena, utt_T, utt_foff, K, dst, ld);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- finish
// out arena: utterance u holds hop*(T_u-1) samples at out_off[u]; y holds them at foff[u]*hop.
__global__ void __launch_bounds__(256) gl_finish_kernel(const float* __restrict__ y, const int* __restrict__ utt_T,
                                                        const int* __restrict__ utt_foff,
                                                        const long long* __restrict__ out_off,
                                                        const unsigned* __restrict__ amax, int hop, int normalise,
                                                        float* __restrict__ out) {
    const int u = blockIdx.y;
    const long len = (long)hop * (utt_T[u] - 1);
    const float* src = y + (long)utt_foff[u] * hop;
    float* dst = out + out_off[u];
    const float m = __uint_as_float(amax[u]);
    const bool norm = normalise && m > 0.f;
    // 4 samples per thread; both offsets are multiples of hop (>= 128) so float4 is aligned
    for (long i = ((long)blockIdx.x * 256 + threadIdx.x) * 4; i < len; i += (long)gridDim.x * 256 * 4) {
        float4 v = *reinterpret_cast<const float4*>(src + i);
        if (norm) { v.x /= m; v.y /= m; v.z /= m; v.w /= m; }
        *reinterpret_cast<float4*>(dst + i) = v;
    }
}

cudaError_t gl_launch_finish(const float* y, const int* utt_T, const int* utt_foff, const long long* out_off,
                             const unsigned* amax, int n_utt, int max_T, int hop, int normalise, float* out,
                             cudaStream_t s) {
    int gx = (int)(((long)hop * max_T / 4 + 255) / 256);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid(gx, n_utt);
    gl_finish_kernel<<<grid, 256, 0, s>>>(y, utt_T, utt_foff, out_off, amax, hop, normalise, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- f32 -> 16-bit PCM
// The caller's per-sample loop `(sample * i16::MAX as f32) as i16` (/root/reference src/lib.rs:153-157;
// Rust float -> int `as` casts truncate toward zero, saturate, and map NaN to 0).  n is a multiple of 4
// per utterance only up to alignment, so the tail is handled scalar.
__device__ __forceinline__ short pcm16_of(float v) {
    const float x = v * 32767.0f;
    if (!(x == x)) return 0;
    if (x >= 32767.0f) return 32767;
    if (x <= -32768.0f) return -32768;
    return (short)(int)x;   // truncation toward zero
}

__global__ void __launch_bounds__(256) gl_pcm16_kernel(const float* __restrict__ src, long long n, short* __restrict__ dst) {
    const long long stride = (long long)gridDim.x * 256;
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        short4 o;
        o.x = pcm16_of(v.x); o.y = pcm16_of(v.y); o.z = pcm16_of(v.z); o.w = pcm16_of(v.w);
        reinterpret_cast<short4*>(dst)[i] = o;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) dst[i] = pcm16_of(src[i]);
}

cudaError_t gl_launch_pcm16(const float* src, long long n, short* dst, int sm_count, cudaStream_t s) {
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    if (blocks < 1) blocks = 1;
    gl_pcm16_kernel<<<(int)blocks, 256, 0, s>>>(src, n, dst);
    return cudaGetLastError();
}

}  // namespace xdtts
