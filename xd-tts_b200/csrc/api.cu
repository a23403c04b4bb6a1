// C ABI of libxdtts_b200 (include/xdtts_b200.h): handles, plans, CUDA-graph capture, error
// plumbing.  Host-side equivalent of what griffin_lim::GriffinLim owns in the reference
// (/root/reference src/tacotron2/mod.rs:441-458 builds it, src/lib.rs:141 calls it).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "api_internal.h"
#include "gl_generic.h"
#include "gl_host.h"
#include "gl_tables.h"
#include "host_copy.h"

namespace xdtts {
// gl_lift.cu
std::vector<float> gl_lift_build_image(const float* pinv, int K, int n_mels, float* p_exp);
int gl_lift_tile_frames(const int* Ts, int B, int n_mels, int K, int sm_count);
cudaError_t gl_lift_prepare(int n_mels);
cudaError_t gl_launch_lift(const float* mel_arena, const float* a_image, float p_exp, const float* pinvT, const int4* tiles, int n_tiles, int tile_frames,
                           const int* utt_T, const int* utt_foff, int n_utt, int max_T, int n_mels, int K, int ld, float power,
                           int delog, int sm_count, float* S, cudaStream_t s, int* n_kernels);
bool gl_lift_uses_tensor_cores(int n_mels, int K);
// gl_aux.cu (S: bin k of frame f at S[f * ld + k], k = 0..K-1)
cudaError_t gl_launch_nnls(const float*, const int*, const float*, const int*, const float*, const int*, const int*, int, int, int,
                           int, float, int, float, int, float, float*, int, cudaStream_t);
cudaError_t gl_launch_nnls_band(const float*, const int*, const float*, int, const int*, const float*, int, const int*, const int*, int,
                                int, int, int, float, int, float, int, float, float*, int, cudaStream_t);
cudaError_t gl_launch_to_frame_major(const float*, const int*, const int*, int, int, int, float*, int, cudaStream_t);
cudaError_t gl_launch_finish(const float*, const int*, const int*, const long long*, const unsigned*, int, int, int, int,
                             float*, cudaStream_t);
cudaError_t gl_launch_pcm16(const float*, long long, short*, int, cudaStream_t);
std::atomic<unsigned long long> g_launches{0};
}  // namespace xdtts

using namespace xdtts;

// ------------------------------------------------------------------ errors
static thread_local std::string tl_error;

namespace xdtts {
int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tl_error = buf;
    return code;
}

bool debug_on() {
    static int v = -1;
    if (v < 0) v = getenv("XDTTS_DEBUG") ? 1 : 0;
    return v == 1;
}

// A CUDA error left behind by earlier work in this process (another library, an ignored return
// in a destroy path) must not be blamed on -- or break -- the next launch: clear it on entry.
void clear_stale_error(const char* where) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && debug_on()) fprintf(stderr, "[xdtts] %s: cleared stale CUDA error %s\n", where, cudaGetErrorName(e));
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace xdtts

#define fail xdtts::set_error

extern "C" const char* xdtts_last_error(void) { return tl_error.c_str(); }
extern "C" unsigned long long xdtts_kernel_launches(void) { return g_launches.load(); }
extern "C" const char* xdtts_version(void) { return "xdtts_b200 0.1.0 sm_100a"; }

extern "C" void* xdtts_host_alloc(unsigned long long bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(XDTTS_ERR_OOM, "cudaHostAlloc(%llu) failed", bytes);
        return nullptr;
    }
    return p;
}
extern "C" void xdtts_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ host math
// Slaney mel scale (librosa hz_to_mel / mel_to_hz, htk=False)
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

extern "C" int xdtts_mel_filter_bank(float sr, int n_fft, int n_mels, float fmin, float fmax, float* out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "mel_filter_bank: out is null");
    if (!(sr > 0.f) || n_fft < 2 || n_mels < 1 || !(fmin >= 0.f)) return fail(XDTTS_ERR_BAD_ARG, "mel_filter_bank: bad parameter");
    const double hi = fmax < 0.f ? 0.5 * (double)sr : (double)fmax;
    if (!(hi > (double)fmin)) return fail(XDTTS_ERR_BAD_ARG, "mel_filter_bank: fmax <= fmin");
    const int K = n_fft / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel(fmin), m_hi = hz_to_mel(hi);
    for (int i = 0; i < n_mels + 2; i++) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * (double)i / (double)(n_mels + 1));
    for (int i = 0; i < n_mels; i++) {
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < K; k++) {
            const double f = 0.5 * (double)sr * (double)k / (double)(K - 1);
            const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
            const double upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
            const double w = std::fmax(0.0, std::fmin(lower, upper));
            out[(size_t)i * K + k] = (float)(w * enorm);
        }
    }
    return XDTTS_OK;
}

// Moore-Penrose pseudo-inverse of B [m][n] (m <= n typical) by one-sided Jacobi (Hestenes) on the
// rows of B, fp64: G B = W with orthogonal rows, |w_i| = sigma_i, G orthogonal, hence
// pinv(B) = sum_i w_i^T g_i / sigma_i^2 over sigma_i > rcond * sigma_max.  Replaces the MKL
// lstsq/SVD the crate reaches through ndarray-linalg (Cargo.lock:1006).  out: [n][m].
static void pinv_rows(const float* B, int m, int n, std::vector<double>* out) {
    std::vector<double> W((size_t)m * n), G((size_t)m * m, 0.0);
    for (size_t i = 0; i < (size_t)m * n; i++) W[i] = (double)B[i];
    for (int i = 0; i < m; i++) G[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < m - 1; p++)
            for (int q = p + 1; q < m; q++) {
                double* wp = &W[(size_t)p * n];
                double* wq = &W[(size_t)q * n];
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < n; k++) { a += wp[k] * wp[k]; b += wq[k] * wq[k]; c += wp[k] * wq[k]; }
                if (c == 0.0 || std::fabs(c) <= 1e-17 * std::sqrt(a * b)) continue;
                off = std::fmax(off, std::fabs(c) / std::sqrt(a * b));
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < n; k++) {
                    const double x = wp[k], y = wq[k];
                    wp[k] = cs * x - sn * y;
                    wq[k] = sn * x + cs * y;
                }
                double* gp = &G[(size_t)p * m];
                double* gq = &G[(size_t)q * m];
                for (int k = 0; k < m; k++) {
                    const double x = gp[k], y = gq[k];
                    gp[k] = cs * x - sn * y;
                    gq[k] = sn * x + cs * y;
                }
            }
        if (off < 1e-15) break;
    }
    std::vector<double> s2(m);
    double smax2 = 0.0;
    for (int i = 0; i < m; i++) {
        double a = 0;
        for (int k = 0; k < n; k++) a += W[(size_t)i * n + k] * W[(size_t)i * n + k];
        s2[i] = a;
        smax2 = std::fmax(smax2, a);
    }
    const double rcond = 1e-15 * (double)(m > n ? m : n);   // numpy.linalg.pinv default
    out->assign((size_t)n * m, 0.0);
    for (int i = 0; i < m; i++) {
        if (!(s2[i] > rcond * rcond * smax2) || s2[i] == 0.0) continue;
        const double inv = 1.0 / s2[i];
        for (int k = 0; k < n; k++) {
            const double wk = W[(size_t)i * n + k] * inv;
            if (wk == 0.0) continue;
            for (int j = 0; j < m; j++) (*out)[(size_t)k * m + j] += wk * G[(size_t)i * m + j];
        }
    }
}

extern "C" int xdtts_pinv(const float* a, int rows, int cols, float* out) {
    if (!a || !out || rows < 1 || cols < 1) return fail(XDTTS_ERR_BAD_ARG, "pinv: bad argument");
    std::vector<double> pv;
    pinv_rows(a, rows, cols, &pv);
    for (size_t i = 0; i < pv.size(); i++) out[i] = (float)pv[i];
    return XDTTS_OK;
}

extern "C" int xdtts_gl_create(const float* mel_basis, int n_mels, int K, int noverlap, float power, int n_iter,
                               float momentum, const xdtts_gl_opts* opts, int device, xdtts_gl** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "gl_create: out is null");
    *out = nullptr;
    if (!mel_basis) return fail(XDTTS_ERR_BAD_ARG, "gl_create: mel_basis is null");
    if (n_mels < 1 || K < 2) return fail(XDTTS_ERR_SHAPE, "gl_create: mel_basis must be [n_mels >= 1, K >= 2], got [%d, %d]", n_mels, K);
    if (n_mels > 256) return fail(XDTTS_ERR_UNSUPPORTED, "gl_create: n_mels = %d > 256", n_mels);
    const int n_fft = 2 * (K - 1);
    if (n_fft < 64 || n_fft > 4096 || (n_fft & (n_fft - 1)))
        return fail(XDTTS_ERR_UNSUPPORTED, "gl_create: n_fft = 2*(K-1) = %d, supported: the powers of two from 64 to 4096", n_fft);
    if (noverlap < 0 || noverlap >= n_fft) return fail(XDTTS_ERR_BAD_ARG, "gl_create: noverlap = %d not in [0, n_fft = %d)", noverlap, n_fft);
    const int hop = n_fft - noverlap;
    // the fused one-launch-per-iteration kernel covers hop == n_fft / 4 at n_fft 512 / 1024 / 2048 (the shipped configuration and
    // its neighbours); every other geometry of GriffinLim::new's signature takes the un-fused kernels of gl_generic.cu
    const bool generic = !((n_fft == 512 || n_fft == 1024 || n_fft == 2048) && hop * 4 == n_fft) || getenv("XDTTS_GL_GENERIC") != nullptr;
    if (!(power > 0.f) || !std::isfinite(power)) return fail(XDTTS_ERR_BAD_ARG, "gl_create: power must be finite and > 0");
    if (n_iter < 0) return fail(XDTTS_ERR_BAD_ARG, "gl_create: iter must be >= 0");
    if (!(momentum >= 0.f) || !std::isfinite(momentum)) return fail(XDTTS_ERR_BAD_ARG, "gl_create: momentum must be finite and >= 0");
    for (size_t i = 0; i < (size_t)n_mels * K; i++)
        if (!std::isfinite(mel_basis[i])) return fail(XDTTS_ERR_BAD_ARG, "gl_create: mel_basis has a non-finite entry");
    if (opts && (opts->delog < 0 || opts->delog > 2 || opts->pad_mode < 0 || opts->pad_mode > 1 || opts->normalise < 0 ||
                 opts->normalise > 1 || opts->run_frames < 0 || opts->persistent < 0 || opts->persistent > 1 || opts->lift < 0 ||
                 opts->lift > 1 || opts->nnls_iters < 0 || opts->fixed_seed < 0 || opts->fixed_seed > 1 || opts->exponent < 0 ||
                 opts->exponent > 1))
        return fail(XDTTS_ERR_BAD_ARG, "gl_create: option out of range");

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(XDTTS_ERR_CUDA, "gl_create: no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= n_dev) return fail(XDTTS_ERR_BAD_ARG, "gl_create: device %d of %d", device, n_dev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(XDTTS_ERR_CUDA, "gl_create: device %d is sm_%d%d, this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));

    xdtts_gl* h = new (std::nothrow) xdtts_gl();
    if (!h) return fail(XDTTS_ERR_OOM, "gl_create: out of host memory");
    h->device = device; h->n_mels = n_mels; h->K = K; h->n_fft = n_fft; h->hop = hop; h->n_iter = n_iter;
    h->generic = generic;
    h->power = power; h->momentum = momentum; h->sm_count = prop.multiProcessorCount;
    if (opts) h->opts = *opts;
    if (h->opts.exponent == 1) h->power = 1.0f / power;   // librosa's mel_to_stft convention: S = x ^ (1 / power)

    std::vector<double> pv;
    pinv_rows(mel_basis, n_mels, K, &pv);
    h->pinv.resize((size_t)K * n_mels);
    std::vector<float> pT((size_t)n_mels * K);
    for (int k = 0; k < K; k++)
        for (int m = 0; m < n_mels; m++) {
            const float v = (float)pv[(size_t)k * n_mels + m];
            h->pinv[(size_t)k * n_mels + m] = v;
            pT[(size_t)m * K + k] = v;
        }
    // sparse forms of the basis + its largest singular value squared (power iteration on A A^T, fp64) for the NNLS lift
    std::vector<int> csr(n_mels + 1, 0), csc(K + 1, 0);
    std::vector<float> csr_val, csc_val;
    for (int m = 0; m < n_mels; m++) {
        for (int k = 0; k < K; k++)
            if (mel_basis[(size_t)m * K + k] != 0.f) { csr.push_back(k); csr_val.push_back(mel_basis[(size_t)m * K + k]); }
        csr[m + 1] = (int)csr_val.size();
    }
    for (int k = 0; k < K; k++) {
        for (int m = 0; m < n_mels; m++)
            if (mel_basis[(size_t)m * K + k] != 0.f) { csc.push_back(m); csc_val.push_back(mel_basis[(size_t)m * K + k]); }
        csc[k + 1] = (int)csc_val.size();
    }
    {   // row_ptr / col_ptr were reserved at the front; the indices follow them
        std::vector<int> rp(csr.begin(), csr.begin() + n_mels + 1), ri(csr.begin() + n_mels + 1, csr.end());
        std::vector<double> v(n_mels, 1.0), w(K), v2(n_mels);
        double lam = 0.0;
        for (int it = 0; it < 300; it++) {
            for (int k = 0; k < K; k++) w[k] = 0.0;
            for (int m = 0; m < n_mels; m++)
                for (int p = rp[m]; p < rp[m + 1]; p++) w[ri[p]] += (double)csr_val[p] * v[m];
            double nrm = 0.0;
            for (int m = 0; m < n_mels; m++) {
                double a = 0.0;
                for (int p = rp[m]; p < rp[m + 1]; p++) a += (double)csr_val[p] * w[ri[p]];
                v2[m] = a;
                nrm += a * a;
            }
            nrm = std::sqrt(nrm);
            if (!(nrm > 0.0)) break;
            lam = nrm;
            for (int m = 0; m < n_mels; m++) v[m] = v2[m] / nrm;
        }
        h->lipschitz = (float)(lam * 1.001);   // small margin above the estimate keeps the step a descent step
    }
    // banded form (a mel filterbank: each row one short run of bins, each column <= 4 filters)
    std::vector<int> band_lo(n_mels, 0), ell_row;
    std::vector<float> bandT, ell_val;
    {
        int rw = 0, cw = 0;
        bool banded = true;
        for (int m = 0; m < n_mels; m++) {
            int lo = -1, hi = -1;
            for (int k = 0; k < K; k++)
                if (mel_basis[(size_t)m * K + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
            band_lo[m] = lo < 0 ? 0 : lo;
            rw = std::max(rw, lo < 0 ? 0 : hi - lo + 1);
        }
        for (int k = 0; k < K; k++) cw = std::max(cw, csc[k + 1] - csc[k]);
        rw = (std::max(rw, 1) + 3) & ~3;
        banded = rw <= 64 && cw >= 1 && cw <= 4;
        if (banded) {
            cw = cw <= 2 ? 2 : 4;
            bandT.assign((size_t)rw * n_mels, 0.f);
            for (int m = 0; m < n_mels; m++)
                for (int i = 0; i < rw && band_lo[m] + i < K; i++) bandT[(size_t)i * n_mels + m] = mel_basis[(size_t)m * K + band_lo[m] + i];
            ell_row.assign((size_t)cw * K, 0);
            ell_val.assign((size_t)cw * K, 0.f);
            for (int k = 0; k < K; k++)
                for (int p = csc[k], c = 0; p < csc[k + 1]; p++, c++) {
                    ell_row[(size_t)c * K + k] = csc[K + 1 + p];
                    ell_val[(size_t)c * K + k] = csc_val[p];
                }
            h->band_rw = rw;
            h->band_cw = cw;
        }
    }
    if (h->opts.lift == 1 && !(h->lipschitz > 0.f)) {
        delete h;
        return fail(XDTTS_ERR_BAD_ARG, "gl_create: the NNLS lift needs a non-zero mel basis");
    }
    const bool fused_size = n_fft == 512 || n_fft == 1024 || n_fft == 2048;
    std::vector<float2> tab = !fused_size ? std::vector<float2>(1) : (n_fft == 512 ? build_tables<4>() : (n_fft == 1024 ? build_tables<8>() : build_tables<16>()));
    std::vector<float> edge = fused_size ? build_edge_scale(n_fft) : std::vector<float>(1);
    std::vector<float2> gtw = glg_build_twiddles(n_fft);
    std::vector<float> gwin = glg_build_window(n_fft);
    cudaError_t e = cudaSuccess;
    std::vector<float> img = gl_lift_build_image(h->pinv.data(), K, n_mels, &h->lift_p_exp);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_lift_img, img.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_lift_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = gl_lift_prepare(n_mels);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_pinvT, pT.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_tables, tab.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_edge, edge.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_csr, csr.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_csc, csc.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_csr_val, (csr_val.size() + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_csc_val, (csc_val.size() + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_csr, csr.data(), csr.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_csc, csc.data(), csc.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !csr_val.empty()) e = cudaMemcpy(h->d_csr_val, csr_val.data(), csr_val.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !csc_val.empty()) e = cudaMemcpy(h->d_csc_val, csc_val.data(), csc_val.size() * 4, cudaMemcpyHostToDevice);
    if (h->band_rw > 0) {
        if (e == cudaSuccess) e = cudaMalloc(&h->d_band_lo, band_lo.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_bandT, bandT.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_ell_row, ell_row.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_ell_val, ell_val.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_band_lo, band_lo.data(), band_lo.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_bandT, bandT.data(), bandT.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_ell_row, ell_row.data(), ell_row.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_ell_val, ell_val.data(), ell_val.size() * 4, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemcpy(h->d_pinvT, pT.data(), pT.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_tables, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_edge, edge.data(), edge.size() * 4, cudaMemcpyHostToDevice);
    {   // the un-fused path's tables, for every handle: utterances shorter than 4 frames take it on any geometry
        if (e == cudaSuccess) e = cudaMalloc(&h->d_gtw, gtw.size() * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&h->d_gwin, gwin.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_gtw, gtw.data(), gtw.size() * sizeof(float2), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_gwin, gwin.data(), gwin.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = glg_prepare(n_fft);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && !generic) e = gl_prepare(n_fft);
    if (e != cudaSuccess) {
        xdtts_gl_destroy(h);
        return fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "gl_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return XDTTS_OK;
}

extern "C" void xdtts_gl_destroy(xdtts_gl* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (xdtts_gl_plan* p : h->cache) xdtts_gl_plan_destroy(p);
    cudaFree(h->d_pinvT);
    cudaFree(h->d_lift_img);
    cudaFree(h->d_tables);
    cudaFree(h->d_edge);
    cudaFree(h->d_gtw);
    cudaFree(h->d_gwin);
    cudaFree(h->d_csr); cudaFree(h->d_csc); cudaFree(h->d_csr_val); cudaFree(h->d_csc_val);
    cudaFree(h->d_band_lo); cudaFree(h->d_bandT); cudaFree(h->d_ell_row); cudaFree(h->d_ell_val);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int xdtts_gl_out_len(const xdtts_gl* h, int T) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "gl_out_len: handle is null");
    if (T < 2) return fail(XDTTS_ERR_SHAPE, "gl_out_len: T = %d, need >= 2 frames (one frame is hop * (T - 1) = 0 samples)", T);
    return h->hop * (T - 1);
}

extern "C" int xdtts_gl_get_pinv(const xdtts_gl* h, float* out) {
    if (!h || !out) return fail(XDTTS_ERR_BAD_ARG, "gl_get_pinv: null argument");
    memcpy(out, h->pinv.data(), h->pinv.size() * 4);
    return XDTTS_OK;
}

// ------------------------------------------------------------------ plan
extern "C" void xdtts_gl_plan_destroy(xdtts_gl_plan* p) {
    if (!p) return;
    cudaSetDevice(p->h->device);
    for (auto& g : p->graphs)
        if (g) cudaGraphExecDestroy(g);
    for (auto& e : p->ev)
        if (e) cudaEventDestroy(e);
    cudaFree(p->d_runs); cudaFree(p->d_T); cudaFree(p->d_foff); cudaFree(p->d_out_off);
    cudaFree(p->d_mel); cudaFree(p->d_in_mag); cudaFree(p->d_in_phase); cudaFree(p->d_turns); cudaFree(p->d_lift_tiles); cudaFree(p->d_seed); cudaFree(p->d_frames);
    if (p->h_seed) cudaFreeHost(p->h_seed);
    for (cudaEvent_t e : p->out_ev)
        if (e) cudaEventDestroy(e);
    if (p->h_in_busy) cudaEventDestroy(p->h_in_busy);
    cudaFree(p->d_state); cudaFree(p->d_y[0]); cudaFree(p->d_y[1]); cudaFree(p->d_halo);
    cudaFree(p->d_out); cudaFree(p->d_flags); cudaFree(p->d_amax); cudaFree(p->d_pcm); cudaFree(p->d_done);
    if (p->h_pcm) cudaFreeHost(p->h_pcm);
    if (p->h_in) cudaFreeHost(p->h_in);
    if (p->h_out) cudaFreeHost(p->h_out);
    delete p;
}

int xdtts::gl_plan_build(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out) {
    *out = nullptr;
    if (!Ts || B < 1) return fail(XDTTS_ERR_BAD_ARG, "plan: need B >= 1 utterances");
    long long total = 0;
    for (int b = 0; b < B; b++) {
        if (Ts[b] < 2) return fail(XDTTS_ERR_SHAPE, "plan: utterance %d has T = %d frames, need >= 2", b, Ts[b]);
        total += Ts[b];
    }
    if (total * (long long)(h->K - 1) >= (1ll << 31)) return fail(XDTTS_ERR_SHAPE, "plan: %lld frames in one batch is too many (split it)", total);
    CU(cudaSetDevice(h->device));
    clear_stale_error(__func__);
    xdtts_gl_plan* p = new (std::nothrow) xdtts_gl_plan();
    if (!p) return fail(XDTTS_ERR_OOM, "plan: out of host memory");
    p->h = h; p->B = B; p->Ts.assign(Ts, Ts + B); p->total_T = (int)total;
    // the fused kernel splits an utterance into runs of >= 4 frames (a hop block is shared by at most two runs): a batch with a
    // shorter utterance runs the un-fused kernels, like the geometries the fused kernel does not cover
    p->generic = h->generic;
    for (int b = 0; b < B; b++) p->generic = p->generic || Ts[b] < 4;
    for (int b = 0; b < B; b++) p->max_T = Ts[b] > p->max_T ? Ts[b] : p->max_T;
    // Runs: one warp each.  A fixed run length (option / environment) gives results that do not depend on
    // the batch; the automatic choice fills the resident warp slots of the device a whole number of times
    // (one wave for up to 64 frames per slot, else the fewest waves that keep runs at <= 64 frames;
    // utterances split in proportion to their length): a partial last wave leaves SMs idle for a run's
    // duration (256 x 1000 frames as 4000 runs of 64 = 2.25 waves; as 5328 runs of 48 = 3 full waves).
    int rf = h->opts.run_frames;
    if (const char* env = getenv("XDTTS_GL_RUN_FRAMES")) rf = atoi(env);
    if (p->generic) rf = 64;   // the un-fused path has no runs; the table only carries the frame offsets
    if (rf > 0) {
        if (rf < 4) rf = 4;
        build_runs(Ts, B, rf, &p->runs, &p->foff);
    } else {
        const long long resident = (long long)h->sm_count * gl_resident_warps_per_sm(h->n_fft);
        const long long waves = (total + resident * 64 - 1) / (resident * 64);
        const long long target = resident * (waves < 1 ? 1 : waves);
        std::vector<int> counts(B);
        long long used = 0;
        for (int b = 0; b < B; b++) {
            long long n = (long long)Ts[b] * target / total;   // floor: the sum never exceeds the target
            if (n > Ts[b] / 4) n = Ts[b] / 4;                   // a hop block may be shared by at most two runs
            counts[b] = n < 1 ? 1 : (int)n;
            used += counts[b];
        }
        // hand the slots the floors left over to the utterances with the longest runs (largest frames per run first,
        // ties by index: deterministic), so that a resident wave is filled exactly
        {
            for (long long left = target - used; left > 0; left--) {
                int best = -1;
                for (int b = 0; b < B; b++) {
                    if (counts[b] + 1 > Ts[b] / 4) continue;
                    if (best < 0 || (long long)Ts[b] * counts[best] > (long long)Ts[best] * counts[b]) best = b;
                }
                if (best < 0) break;
                counts[best]++;
            }
        }
        build_runs_counts(Ts, B, counts.data(), &p->runs, &p->foff);
        rf = 0;
        for (const GlRun& r : p->runs) rf = std::max(rf, r.tb - r.ta);
    }
    p->run_frames = rf;
    {   // persistent single-launch path: every run must be resident at once (checked again at launch)
        const char* env = getenv("XDTTS_GL_PERSISTENT");
        const long long resident = p->generic ? 0 : (long long)h->sm_count * gl_resident_warps_per_sm(h->n_fft);
        const bool want = env ? atoi(env) != 0 : h->opts.persistent != 0;
        p->use_persistent = !p->generic && want && (long long)p->runs.size() <= resident;
    }
    p->out_off.resize(B);
    for (int b = 0; b < B; b++) {
        p->out_off[b] = p->out_total;
        p->out_total += (long long)h->hop * (Ts[b] - 1);
    }
    const size_t M = (size_t)h->K - 1, H = (size_t)h->hop, TT = (size_t)total, nr = p->runs.size();
    cudaError_t e = cudaSuccess;
#define ALLOC(ptr, bytes) if (e == cudaSuccess) e = cudaMalloc((void**)&(ptr), (bytes))
    ALLOC(p->d_runs, nr * sizeof(GlRun));
    ALLOC(p->d_T, B * sizeof(int));
    ALLOC(p->d_foff, B * sizeof(int));
    ALLOC(p->d_out_off, B * sizeof(long long));
    // one state record per frame: [R: M float2 | S: M floats | S_nyq + 3 floats of padding] (gl_core.cuh Geo::REC)
    p->rec_f = (int)(3 * M + 4);
    ALLOC(p->d_state, TT * (size_t)p->rec_f * 4);
    {   // frame tiles of the lift: (frame row of the utterance, its T, first frame, frames in the tile), never across utterances;
        // the extent is chosen per plan so that the tile count fills the CTAs' last round (gl_lift_tile_frames)
        const int tf = p->lift_tile_frames = gl_lift_tile_frames(Ts, B, h->n_mels, h->K, h->sm_count);
        for (int b = 0; b < B; b++)
            for (int t0 = 0; t0 < Ts[b]; t0 += tf) p->lift_tiles.push_back(make_int4(p->foff[b], Ts[b], t0, std::min(tf, Ts[b] - t0)));
    }
    ALLOC(p->d_lift_tiles, p->lift_tiles.size() * sizeof(int4));
    ALLOC(p->d_seed, 8 + (size_t)B * sizeof(int));   // [u64 phase seed][int stream index of each utterance]
    if (p->generic) ALLOC(p->d_frames, TT * (size_t)h->n_fft * 4);
    ALLOC(p->d_y[0], TT * H * 4);
    ALLOC(p->d_y[1], TT * H * 4);
    ALLOC(p->d_halo, nr * 6 * H * 4);
    ALLOC(p->d_flags, nr * sizeof(unsigned));
    ALLOC(p->d_done, nr * sizeof(unsigned));
    ALLOC(p->d_amax, B * sizeof(unsigned));
    ALLOC(p->d_out, (size_t)p->out_total * 4);
#undef ALLOC
    if (e == cudaSuccess) e = cudaMemcpy(p->d_runs, p->runs.data(), nr * sizeof(GlRun), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_T, Ts, B * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_foff, p->foff.data(), B * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_out_off, p->out_off.data(), B * sizeof(long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->d_flags, 0, nr * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(p->d_state, 0, TT * (size_t)p->rec_f * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_lift_tiles, p->lift_tiles.data(), p->lift_tiles.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&p->h_seed, 8 + (size_t)B * sizeof(int), cudaHostAllocDefault);
    for (int i = 0; i < 6 && e == cudaSuccess; i++) e = cudaEventCreate(&p->ev[i]);
    if (e != cudaSuccess) {
        xdtts_gl_plan_destroy(p);
        return fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "plan: %s", cudaGetErrorString(e));
    }
    *out = p;
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_create(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out) {
    if (!h || !out) return fail(XDTTS_ERR_BAD_ARG, "plan_create: null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    return gl_plan_build(h, Ts, B, out);
}

extern "C" int xdtts_gl_plan_lift_ms(const xdtts_gl_plan* p, float* ms) {
    if (!p || !ms) return fail(XDTTS_ERR_BAD_ARG, "plan_lift_ms: null argument");
    if (p->lift_ms < 0.f) return fail(XDTTS_ERR_BAD_ARG, "plan_lift_ms: no kernel-by-kernel (XDTTS_RUN_NO_GRAPH) pass has run on this plan");
    *ms = p->lift_ms;
    return XDTTS_OK;
}

static int plan_enqueue_lift(xdtts_gl_plan* p, int flags, cudaStream_t s, unsigned long long* launched);

extern "C" int xdtts_gl_plan_time_lift(xdtts_gl_plan* p, int reps, float* ms_per_launch) {
    if (!p || !ms_per_launch) return fail(XDTTS_ERR_BAD_ARG, "plan_time_lift: null argument");
    if (reps < 1 || reps > 10000) return fail(XDTTS_ERR_BAD_ARG, "plan_time_lift: reps must be in 1..10000");
    xdtts_gl* h = p->h;
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    unsigned long long launched = 0;
    int rc = plan_enqueue_lift(p, 0, s, &launched);   // warm-up
    CU(cudaEventRecord(p->ev[4], s));
    for (int i = 0; i < reps && !rc; i++) rc = plan_enqueue_lift(p, 0, s, &launched);
    if (rc) return rc;
    CU(cudaEventRecord(p->ev[5], s));
    CU(cudaStreamSynchronize(s));
    g_launches += launched;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, p->ev[4], p->ev[5]));
    *ms_per_launch = ms / (float)reps;
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_is_persistent(const xdtts_gl_plan* p) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "plan_is_persistent: plan is null");
    return p->use_persistent ? 1 : 0;
}

extern "C" int xdtts_gl_plan_info(const xdtts_gl_plan* p, int* info4) {
    if (!p || !info4) return fail(XDTTS_ERR_BAD_ARG, "plan_info: null argument");
    info4[0] = (int)p->runs.size();
    info4[1] = p->run_frames;
    info4[2] = ((int)p->runs.size() + gl_warps_per_cta(p->h->n_fft) - 1) / gl_warps_per_cta(p->h->n_fft);
    info4[3] = p->total_T;
    return XDTTS_OK;
}

int xdtts::gl_plan_upload_locked(xdtts_gl_plan* p, int kind, const float* const* srcs, cudaStream_t s) {
    xdtts_gl* h = p->h;
    if (!s) s = h->stream;
    clear_stale_error(__func__);
    if (!srcs) return fail(XDTTS_ERR_BAD_ARG, "plan_upload: srcs is null");
    if (kind < 0 || kind > 2) return fail(XDTTS_ERR_BAD_ARG, "plan_upload: kind %d", kind);
    for (int b = 0; b < p->B; b++)
        if (!srcs[b]) return fail(XDTTS_ERR_BAD_ARG, "plan_upload: srcs[%d] is null", b);
    CU(cudaSetDevice(h->device));
    const size_t rows = kind == 0 ? (size_t)h->n_mels : (size_t)h->K;
    float** dst = kind == 0 ? &p->d_mel : (kind == 1 ? &p->d_in_mag : &p->d_in_phase);
    if (!*dst) CU(cudaMalloc((void**)dst, rows * (size_t)p->total_T * 4));
    if (kind == 2 && !p->d_turns) CU(cudaMalloc((void**)&p->d_turns, (size_t)p->total_T * h->K * 4));   // [frames][K]
    // pageable sources go through the plan's pinned staging buffer: helper threads copy utterance by utterance, and an
    // utterance goes on the bus (async DMA) as soon as it is staged, while the next ones are still being copied
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(srcs[b]);
    if (!all_pinned) {
        const size_t need = rows * (size_t)p->total_T;
        if (p->h_in_busy) CU(cudaEventSynchronize(p->h_in_busy));   // the previous upload's DMA still reads the staging buffer
        if (p->h_in_floats < need) {
            if (p->h_in) cudaFreeHost(p->h_in);
            p->h_in = nullptr; p->h_in_floats = 0;
            CU(cudaHostAlloc((void**)&p->h_in, need * 4, cudaHostAllocDefault));
            p->h_in_floats = need;
        }
        if (!p->h_in_busy) CU(cudaEventCreateWithFlags(&p->h_in_busy, cudaEventDisableTiming));
        std::vector<std::atomic<int>> pending(p->B);
        for (int b = 0; b < p->B; b++) {
            pending[b].store(0);
            host_copy_async(p->h_in + rows * (size_t)p->foff[b], srcs[b], rows * (size_t)p->Ts[b] * 4, &pending[b], true);
        }
        cudaError_t ce = cudaSuccess;
        for (int b = 0; b < p->B; b++) {
            host_copy_wait(&pending[b]);      // every chunk must be done before this function returns, error or not
            if (ce == cudaSuccess)
                ce = cudaMemcpyAsync(*dst + rows * (size_t)p->foff[b], p->h_in + rows * (size_t)p->foff[b], rows * (size_t)p->Ts[b] * 4,
                                     cudaMemcpyHostToDevice, s);
        }
        CU(ce);
        CU(cudaEventRecord(p->h_in_busy, s));
    } else {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(*dst + rows * (size_t)p->foff[b], srcs[b], rows * (size_t)p->Ts[b] * 4, cudaMemcpyHostToDevice, s));
    }
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_upload(xdtts_gl_plan* p, int kind, const float* const* srcs) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "plan_upload: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return gl_plan_upload_locked(p, kind, srcs, nullptr);
}

// The reference draws new random phases on every call.  Here the phase field is a
// counter-based generator; the seed of call c on a handle is opts.seed + c * odd constant (call 0 uses opts.seed
// itself), unless opts.fixed_seed asks for the same field on every call.
unsigned long long xdtts::gl_draw_seed(xdtts_gl* h) {
    const unsigned long long c = h->opts.fixed_seed ? 0ull : h->calls.fetch_add(1);
    return h->opts.seed + c * 0xD1B54A32D192ED03ull;
}

// seed + stream indices of the next pass of this plan -> device (stream-ordered before the launches that read them).
// seed_ids == null: utterance b draws stream b.
int xdtts::gl_plan_set_seed(xdtts_gl_plan* p, unsigned long long seed, const int* seed_ids, cudaStream_t s) {
    memcpy(p->h_seed, &seed, 8);
    int* ids = reinterpret_cast<int*>(p->h_seed + 8);
    for (int b = 0; b < p->B; b++) ids[b] = seed_ids ? seed_ids[b] : b;
    CU(cudaMemcpyAsync(p->d_seed, p->h_seed, 8 + (size_t)p->B * sizeof(int), cudaMemcpyHostToDevice, s));
    return XDTTS_OK;
}

// the mel -> linear step of a pass (lift, or the transpose of caller-supplied magnitudes; + NNLS when enabled); *launched += kernels
static int plan_enqueue_lift(xdtts_gl_plan* p, int flags, cudaStream_t s, unsigned long long* launched) {
    xdtts_gl* h = p->h;
    const int M = h->K - 1;
    const bool from_mag = flags & XDTTS_RUN_FROM_MAG;
    int lift_kernels = 1;
    float* d_S = p->d_state + 2 * M;   // the S part of frame 0's record; records are rec_f floats apart
    if (from_mag) {
        CU(gl_launch_to_frame_major(p->d_in_mag, p->d_T, p->d_foff, p->B, p->max_T, h->K, d_S, p->rec_f, s));
    } else {
        const bool nnls = h->opts.lift == 1;
        CU(gl_launch_lift(p->d_mel, h->d_lift_img, h->lift_p_exp, h->d_pinvT, p->d_lift_tiles, (int)p->lift_tiles.size(), p->lift_tile_frames, p->d_T, p->d_foff, p->B,
                          p->max_T, h->n_mels, h->K, p->rec_f, nnls ? 1.0f : h->power, h->opts.delog, h->sm_count, d_S, s, &lift_kernels));
        *launched += (unsigned long long)(lift_kernels - 1);
        if (nnls) {   // refine the clipped least-squares start in place, then apply the exponent
            const int iters = h->opts.nnls_iters > 0 ? h->opts.nnls_iters : 300;
            if (h->band_rw > 0 && !getenv("XDTTS_NNLS_GENERIC"))
                CU(gl_launch_nnls_band(p->d_mel, h->d_band_lo, h->d_bandT, h->band_rw, h->d_ell_row, h->d_ell_val, h->band_cw, p->d_T,
                                       p->d_foff, p->B, p->max_T, h->n_mels, h->K, h->power, h->opts.delog, h->lipschitz, iters, 3e-6f,
                                       d_S, p->rec_f, s));
            else
                CU(gl_launch_nnls(p->d_mel, h->d_csr, h->d_csr_val, h->d_csc, h->d_csc_val, p->d_T, p->d_foff, p->B, p->max_T, h->n_mels,
                                  h->K, h->power, h->opts.delog, h->lipschitz, iters, 3e-6f, d_S, p->rec_f, s));
            (*launched)++;
        }
    }
    (*launched)++;
    return XDTTS_OK;
}

// enqueue the whole pass on the handle's stream; ev[1]/ev[2] bracket the steady-state launches when timed
static int plan_enqueue(xdtts_gl_plan* p, int flags, bool timed, int* n_mid, cudaStream_t s, bool capturing = false) {
    unsigned long long launched = 0;   // added to the process-wide counter at the end -- not at all while capturing a graph
    xdtts_gl* h = p->h;
    const int M = h->K - 1;
    const bool from_mag = flags & XDTTS_RUN_FROM_MAG, use_phase = flags & XDTTS_RUN_USE_PHASE;
    CU(cudaMemsetAsync(p->d_amax, 0, p->B * sizeof(unsigned), s));
    if (timed) CU(cudaEventRecord(p->ev[4], s));
    {
        int rc = plan_enqueue_lift(p, flags, s, &launched);
        if (rc) return rc;
    }
    if (timed) CU(cudaEventRecord(p->ev[5], s));
    if (use_phase) {
        CU(gl_launch_to_frame_major(p->d_in_phase, p->d_T, p->d_foff, p->B, p->max_T, h->K, p->d_turns, h->K, s));
        launched++;
    }
    if (p->generic) {   // any hop / any power-of-two n_fft / utterances of 2-3 frames: un-fused kernels, two launches per iteration (gl_generic.cu)
        GlgParams q;
        memset(&q, 0, sizeof(q));
        q.n_fft = h->n_fft; q.hop = h->hop; q.n_utt = p->B;
        q.utt_T = p->d_T; q.utt_foff = p->d_foff; q.state = p->d_state; q.rec_f = p->rec_f; q.frames = p->d_frames; q.y = p->d_y[0];
        q.turns = use_phase ? p->d_turns : nullptr;
        q.seed = reinterpret_cast<const unsigned long long*>(p->d_seed);
        q.utt_seed_id = reinterpret_cast<const int*>(p->d_seed + 8);
        q.tw = h->d_gtw; q.win = h->d_gwin; q.amax = p->d_amax;
        q.alpha = h->momentum / (1.0f + h->momentum);
        q.pad_mode = h->opts.pad_mode;
        int mids = 0;
        for (int it = 0; it <= h->n_iter; it++) {
            if (timed && it == 2) CU(cudaEventRecord(p->ev[1], s));
            CU(glg_launch_frames(q, it == 0 ? 0 : 1, p->total_T, s));
            CU(glg_launch_ola(q, it == h->n_iter, p->max_T, s));
            launched += 2;
            if (it >= 2 && it < h->n_iter) mids++;
            if (timed && it == h->n_iter - 1 && it >= 2) CU(cudaEventRecord(p->ev[2], s));
        }
        CU(glg_launch_finish(p->d_y[0], p->d_T, p->d_foff, p->d_out_off, p->d_amax, p->B, p->max_T, h->hop, h->opts.normalise == 0, p->d_out, s));
        launched++;
        if (n_mid) *n_mid = mids;
        if (!capturing) g_launches += launched;
        return XDTTS_OK;
    }
    GlParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.n_runs = (int)p->runs.size();
    gp.runs = p->d_runs; gp.utt_T = p->d_T; gp.utt_foff = p->d_foff;
    gp.state = p->d_state; gp.halo = p->d_halo; gp.flags = p->d_flags; gp.amax = p->d_amax;
    gp.edge_scale = h->d_edge; gp.tables = h->d_tables;
    gp.turns = use_phase ? p->d_turns : nullptr;
    gp.seed = reinterpret_cast<const unsigned long long*>(p->d_seed);
    gp.utt_seed_id = reinterpret_cast<const int*>(p->d_seed + 8);
    gp.alpha = h->momentum / (1.0f + h->momentum);
    gp.inv_n = 1.0f / (float)h->n_fft;
    gp.pad_mode = h->opts.pad_mode;
    int mids = 0;
    bool done = false;
    if (p->use_persistent && !(flags & XDTTS_RUN_PER_LAUNCH)) {
        gp.ybuf[0] = p->d_y[0]; gp.ybuf[1] = p->d_y[1]; gp.done = p->d_done; gp.n_iter = h->n_iter;
        CU(cudaMemsetAsync(p->d_done, 0, p->runs.size() * sizeof(unsigned), s));
        if (timed) CU(cudaEventRecord(p->ev[1], s));
        bool fits = false;
        CU(gl_launch_persistent(h->n_fft, gp, h->sm_count, s, &fits));
        if (fits) {
            if (timed) CU(cudaEventRecord(p->ev[2], s));
            launched++;
            mids = 1;
            done = true;
        } else {
            p->use_persistent = false;   // does not fit this device after all: launch per iteration from now on
        }
    }
    if (!done) {
        gp.y_in = p->d_y[1]; gp.y_out = p->d_y[0];
        CU(gl_launch_iteration(h->n_fft, GL_MODE_INIT, h->n_iter == 0, gp, s));
        launched++;
        for (int it = 1; it <= h->n_iter; it++) {
            gp.y_in = p->d_y[(it - 1) & 1];
            gp.y_out = p->d_y[it & 1];
            const bool last = it == h->n_iter;
            if (timed && it == 2) CU(cudaEventRecord(p->ev[1], s));
            CU(gl_launch_iteration(h->n_fft, it == 1 ? GL_MODE_FIRST : GL_MODE_MID, last, gp, s));
            launched++;
            if (it >= 2 && !last) mids++;
            if (timed && it == h->n_iter - 1 && it >= 2) CU(cudaEventRecord(p->ev[2], s));
        }
    }
    CU(gl_launch_finish(p->d_y[h->n_iter & 1], p->d_T, p->d_foff, p->d_out_off, p->d_amax, p->B, p->max_T, h->hop,
                        h->opts.normalise == 0, p->d_out, s));
    launched++;
    if (n_mid) *n_mid = mids;
    if (!capturing) g_launches += launched;
    return XDTTS_OK;
}

// launch the plan's CUDA graph of the whole pass (captured on first use) on stream s, without waiting
static int plan_launch_graph(xdtts_gl_plan* p, int flags, cudaStream_t s) {
    xdtts_gl* h = p->h;
    const int gi = flags & 3;   // (PER_LAUNCH does not change what is captured: this is the per-launch path)
    if (!p->graphs[gi]) {
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = plan_enqueue(p, flags, false, nullptr, s, true);   // captured, not launched: not counted
        cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return fail(XDTTS_ERR_CUDA, "plan_run: graph capture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&p->graphs[gi], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(XDTTS_ERR_CUDA, "plan_run: graph instantiate: %s", cudaGetErrorString(e));
    }
    CU(cudaGraphLaunch(p->graphs[gi], s));
    g_launches += (unsigned long long)((p->generic ? 2 * h->n_iter + 4 : h->n_iter + 3) + ((!(flags & XDTTS_RUN_FROM_MAG) && h->opts.lift == 1) ? 1 : 0) + ((flags & XDTTS_RUN_USE_PHASE) ? 1 : 0));
    return XDTTS_OK;
}

// enqueue the pass on stream s and return without waiting (xdtts_pipe): CUDA graph, or the persistent kernel
int xdtts::gl_plan_launch_async(xdtts_gl_plan* p, int flags, cudaStream_t s, const unsigned long long* seed, const int* seed_ids) {
    clear_stale_error(__func__);
    CU(cudaSetDevice(p->h->device));
    if (!(flags & XDTTS_RUN_USE_PHASE)) {
        int rc = gl_plan_set_seed(p, seed ? *seed : gl_draw_seed(p->h), seed_ids, s);
        if (rc) return rc;
    }
    if (p->use_persistent && !(flags & XDTTS_RUN_PER_LAUNCH)) return plan_enqueue(p, flags, false, nullptr, s);
    return plan_launch_graph(p, flags, s);
}

// enqueue the device -> host copies of the result on stream s; pinned destinations receive the data
// directly, pageable ones go through the plan's pinned buffer (*staged = true: call
// gl_plan_download_finish after the stream has been synchronised)
int xdtts::gl_plan_download_async(xdtts_gl_plan* p, float* const* outs, cudaStream_t s, bool* staged) {
    xdtts_gl* h = p->h;
    clear_stale_error(__func__);
    if (!outs) return fail(XDTTS_ERR_BAD_ARG, "plan_download: outs is null");
    for (int b = 0; b < p->B; b++)
        if (!outs[b]) return fail(XDTTS_ERR_BAD_ARG, "plan_download: outs[%d] is null", b);
    CU(cudaSetDevice(h->device));
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(outs[b]);
    *staged = !all_pinned;
    if (all_pinned) {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(outs[b], p->d_out + p->out_off[b], (size_t)h->hop * (p->Ts[b] - 1) * 4, cudaMemcpyDeviceToHost, s));
    } else {
        // utterance by utterance into the pinned staging buffer, an event after each: gl_plan_download_finish hands an
        // utterance to the copy threads as soon as its DMA is done, while the following ones are still on the bus
        if (!p->h_out) CU(cudaHostAlloc((void**)&p->h_out, (size_t)p->out_total * 4, cudaHostAllocDefault));
        if (p->out_ev.empty()) {
            p->out_ev.resize(p->B, nullptr);
            for (int b = 0; b < p->B; b++) CU(cudaEventCreateWithFlags(&p->out_ev[b], cudaEventDisableTiming));
        }
        for (int b = 0; b < p->B; b++) {
            CU(cudaMemcpyAsync(p->h_out + p->out_off[b], p->d_out + p->out_off[b], (size_t)h->hop * (p->Ts[b] - 1) * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaEventRecord(p->out_ev[b], s));
        }
    }
    return XDTTS_OK;
}

void xdtts::gl_plan_download_finish(xdtts_gl_plan* p, float* const* outs) {
    std::atomic<int> pending{0};
    for (int b = 0; b < p->B; b++) {
        if (!p->out_ev.empty()) cudaEventSynchronize(p->out_ev[b]);
        host_copy_async(outs[b], p->h_out + p->out_off[b], (size_t)p->h->hop * (p->Ts[b] - 1) * 4, &pending);
    }
    host_copy_wait(&pending);
}

int xdtts::gl_plan_run_locked(xdtts_gl_plan* p, int flags, float* ms_total, float* ms_iter, int* n_iter_launches,
                              const unsigned long long* seed, const int* seed_ids) {
    xdtts_gl* h = p->h;
    clear_stale_error(__func__);
    CU(cudaSetDevice(h->device));
    if (!(flags & XDTTS_RUN_USE_PHASE)) {
        int rc = gl_plan_set_seed(p, seed ? *seed : gl_draw_seed(h), seed_ids, h->stream);
        if (rc) return rc;
    }
    if ((flags & XDTTS_RUN_FROM_MAG) && !p->d_in_mag) return fail(XDTTS_ERR_BAD_ARG, "plan_run: FROM_MAG without uploaded magnitudes");
    if (!(flags & XDTTS_RUN_FROM_MAG) && !p->d_mel) return fail(XDTTS_ERR_BAD_ARG, "plan_run: no mels uploaded");
    if ((flags & XDTTS_RUN_USE_PHASE) && !p->d_in_phase) return fail(XDTTS_ERR_BAD_ARG, "plan_run: USE_PHASE without uploaded phase");
    cudaStream_t s = h->stream;
    int mids = 0;
    if (ms_iter) *ms_iter = 0.f;
    if (n_iter_launches) *n_iter_launches = 0;
    const bool persistent = p->use_persistent && !(flags & XDTTS_RUN_PER_LAUNCH);
    if ((flags & XDTTS_RUN_NO_GRAPH) || persistent) {   // the persistent path is 3-4 launches: no graph needed
        CU(cudaEventRecord(p->ev[0], s));
        int rc = plan_enqueue(p, flags, (flags & XDTTS_RUN_NO_GRAPH) != 0, &mids, s);
        if (rc) return rc;
        CU(cudaEventRecord(p->ev[3], s));
        CU(cudaStreamSynchronize(s));
        if ((flags & XDTTS_RUN_NO_GRAPH) && ms_iter && mids > 0) CU(cudaEventElapsedTime(ms_iter, p->ev[1], p->ev[2]));
        if ((flags & XDTTS_RUN_NO_GRAPH) && n_iter_launches) *n_iter_launches = mids;
        p->lift_ms = -1.f;
        if (flags & XDTTS_RUN_NO_GRAPH) CU(cudaEventElapsedTime(&p->lift_ms, p->ev[4], p->ev[5]));
    } else {
        CU(cudaEventRecord(p->ev[0], s));
        int rc = plan_launch_graph(p, flags, s);
        if (rc) return rc;
        CU(cudaEventRecord(p->ev[3], s));
        CU(cudaStreamSynchronize(s));
    }
    if (ms_total) CU(cudaEventElapsedTime(ms_total, p->ev[0], p->ev[3]));
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_run(xdtts_gl_plan* p, int flags, float* ms_total, float* ms_iter, int* n_iter_launches) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "plan_run: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return gl_plan_run_locked(p, flags, ms_total, ms_iter, n_iter_launches, nullptr, nullptr);
}

int xdtts::gl_plan_download_locked(xdtts_gl_plan* p, float* const* outs) {
    xdtts_gl* h = p->h;
    clear_stale_error(__func__);
    if (!outs) return fail(XDTTS_ERR_BAD_ARG, "plan_download: outs is null");
    for (int b = 0; b < p->B; b++)
        if (!outs[b]) return fail(XDTTS_ERR_BAD_ARG, "plan_download: outs[%d] is null", b);
    bool staged = false;
    int rc = gl_plan_download_async(p, outs, h->stream, &staged);
    if (rc) return rc;
    if (staged) gl_plan_download_finish(p, outs);   // overlaps the host copies with the DMA of the following utterances
    CU(cudaStreamSynchronize(h->stream));
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_download(xdtts_gl_plan* p, float* const* outs) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "plan_download: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return gl_plan_download_locked(p, outs);
}

// the caller's f32 -> i16 loop (src/lib.rs:153-157) on the device: halves the bytes that cross PCIe
static int plan_download_pcm16_locked(xdtts_gl_plan* p, short* const* outs) {
    xdtts_gl* h = p->h;
    clear_stale_error(__func__);
    if (!outs) return fail(XDTTS_ERR_BAD_ARG, "plan_download_pcm16: outs is null");
    for (int b = 0; b < p->B; b++)
        if (!outs[b]) return fail(XDTTS_ERR_BAD_ARG, "plan_download_pcm16: outs[%d] is null", b);
    CU(cudaSetDevice(h->device));
    if (!p->d_pcm) CU(cudaMalloc((void**)&p->d_pcm, (size_t)p->out_total * 2));
    CU(gl_launch_pcm16(p->d_out, p->out_total, p->d_pcm, h->sm_count, h->stream));
    g_launches++;
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(outs[b]);
    if (all_pinned) {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(outs[b], p->d_pcm + p->out_off[b], (size_t)h->hop * (p->Ts[b] - 1) * 2, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else {
        if (!p->h_pcm) CU(cudaHostAlloc((void**)&p->h_pcm, (size_t)p->out_total * 2, cudaHostAllocDefault));
        CU(cudaMemcpyAsync(p->h_pcm, p->d_pcm, (size_t)p->out_total * 2, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int b = 0; b < p->B; b++) memcpy(outs[b], p->h_pcm + p->out_off[b], (size_t)h->hop * (p->Ts[b] - 1) * 2);
    }
    return XDTTS_OK;
}

extern "C" int xdtts_gl_plan_download_pcm16(xdtts_gl_plan* p, short* const* outs) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "plan_download_pcm16: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return plan_download_pcm16_locked(p, outs);
}

extern "C" int xdtts_gl_plan_peek(xdtts_gl_plan* p, int what, float* out, long long n_floats) {
    if (!p || !out) return fail(XDTTS_ERR_BAD_ARG, "plan_peek: null argument");
    std::lock_guard<std::mutex> lk(p->h->mu);
    const long long M = p->h->K - 1, TT = p->total_T;
    const long long n = what == 0 ? TT * M : (what == 1 ? TT : TT * M * 2);
    if (what < 0 || what > 2 || n != n_floats) return fail(XDTTS_ERR_SHAPE, "plan_peek: what=%d expects %lld floats, got %lld", what, n, n_floats);
    CU(cudaSetDevice(p->h->device));
    CU(cudaStreamSynchronize(p->h->stream));
    // gather the requested part of every frame's state record [R: 2M floats | S: M floats | S_nyq | pad]
    const size_t off = what == 0 ? 2 * M : (what == 1 ? 3 * M : 0), width = (what == 0 ? M : (what == 1 ? 1 : 2 * M)) * 4;
    CU(cudaMemcpy2D(out, width, p->d_state + off, (size_t)p->rec_f * 4, width, (size_t)TT, cudaMemcpyDeviceToHost));
    return XDTTS_OK;
}

// plan of this exact batch shape from the handle's small cache (built on a miss); h->mu held by the caller
int xdtts::gl_cached_plan(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out) {
    xdtts_gl_plan* p = nullptr;
    for (xdtts_gl_plan* c : h->cache)
        if (c->B == B && memcmp(c->Ts.data(), Ts, sizeof(int) * B) == 0) p = c;
    if (!p) {
        if (h->cache.size() >= 4) {   // small FIFO of shapes; buffers are large
            xdtts_gl_plan_destroy(h->cache.front());
            h->cache.erase(h->cache.begin());
        }
        int rc = gl_plan_build(h, Ts, B, &p);
        if (rc) return rc;
        h->cache.push_back(p);
    }
    *out = p;
    return XDTTS_OK;
}

// device arena the lift reads: utterance u is a row-major [n_mels][T_u] block at float offset foff[u] * n_mels
int xdtts::gl_plan_mel_arena(xdtts_gl_plan* p, float** out) {
    if (!p->d_mel) {
        CU(cudaSetDevice(p->h->device));
        CU(cudaMalloc((void**)&p->d_mel, (size_t)p->h->n_mels * (size_t)p->total_T * 4));
    }
    *out = p->d_mel;
    return XDTTS_OK;
}

// ------------------------------------------------------------------ batch entry points
int xdtts::gl_batch_common(xdtts_gl* h, int kind, const float* const* ins, const int* Ts, int B, const float* const* phases,
                           float* const* outs, short* const* pcm_outs, const unsigned long long* seed, const int* seed_ids) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "infer: handle is null");
    if (!ins || !Ts || (!outs && !pcm_outs)) return fail(XDTTS_ERR_BAD_ARG, "infer: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "infer: B = %d", B);
    std::lock_guard<std::mutex> lk(h->mu);
    xdtts_gl_plan* p = nullptr;
    {
        int rc = gl_cached_plan(h, Ts, B, &p);
        if (rc) return rc;
    }
    int rc = gl_plan_upload_locked(p, kind, ins, nullptr);
    if (rc) return rc;
    int flags = kind == 1 ? XDTTS_RUN_FROM_MAG : 0;
    if (phases) {
        rc = gl_plan_upload_locked(p, 2, phases, nullptr);
        if (rc) return rc;
        flags |= XDTTS_RUN_USE_PHASE;
    }
    rc = gl_plan_run_locked(p, flags, nullptr, nullptr, nullptr, seed, seed_ids);
    if (rc) return rc;
    return pcm_outs ? plan_download_pcm16_locked(p, pcm_outs) : gl_plan_download_locked(p, outs);
}

extern "C" int xdtts_gl_infer_batch_pcm16(xdtts_gl* h, const float* const* mels, const int* Ts, int B,
                                          const float* const* init_phases, short* const* outs) {
    return gl_batch_common(h, 0, mels, Ts, B, init_phases, nullptr, outs, nullptr, nullptr);
}

extern "C" int xdtts_gl_infer_batch(xdtts_gl* h, const float* const* mels, const int* Ts, int B,
                                    const float* const* init_phases, float* const* outs) {
    return gl_batch_common(h, 0, mels, Ts, B, init_phases, outs, nullptr, nullptr, nullptr);
}

extern "C" int xdtts_gl_from_mag_batch(xdtts_gl* h, const float* const* mags, const int* Ts, int B,
                                       const float* const* init_phases, float* const* outs) {
    return gl_batch_common(h, 1, mags, Ts, B, init_phases, outs, nullptr, nullptr, nullptr);
}

extern "C" int xdtts_gl_infer(xdtts_gl* h, const float* mel, int T, const float* init_phase, float* out, int out_len) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "infer: handle is null");
    if (!mel || !out) return fail(XDTTS_ERR_BAD_ARG, "infer: null argument");
    if (T < 2) return fail(XDTTS_ERR_SHAPE, "infer: T = %d, need >= 2 frames", T);
    if (out_len != h->hop * (T - 1)) return fail(XDTTS_ERR_SHAPE, "infer: out_len = %d, expected hop*(T-1) = %d", out_len, h->hop * (T - 1));
    const float* mels[1] = {mel};
    const float* ph[1] = {init_phase};
    float* outs[1] = {out};
    return gl_batch_common(h, 0, mels, &T, 1, init_phase ? ph : nullptr, outs, nullptr, nullptr, nullptr);
}
