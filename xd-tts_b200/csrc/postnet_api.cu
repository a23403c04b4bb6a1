// C ABI of the postnet (include/xdtts_b200.h, xdtts_postnet_*): replaces the `postnet` ONNX session of
// the reference's Tacotron2 (/root/reference src/tacotron2/mod.rs:256-259 loads it, :344-357 runs it
// and extracts "mel_outputs_postnet").  Host side: BatchNorm folding, weight re-layout (bf16 hi/lo,
// K-major per tap), TMA descriptors, device-resident plans; plus the fused tail
// postnet -> mel-to-linear lift -> Griffin-Lim that XdTts::infer performs with two calls
// (src/lib.rs:123 and :141).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "api_internal.h"
#include "host_copy.h"
#include "postnet.h"

using namespace xdtts;
#define fail xdtts::set_error

// ------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor [rows][cols] with row pitch ld (elements), boxes of [box_rows][64], 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, int cols, long long rows, int ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(XDTTS_ERR_CUDA, "postnet: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)PN_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(XDTTS_ERR_CUDA, "postnet: cuTensorMapEncodeTiled failed (%d) for [%lld x %d] ld %d box %d", (int)r, rows, cols, ld, box_rows);
    return XDTTS_OK;
}

// ------------------------------------------------------------------ handle
struct xdtts_postnet {
    int device = 0, n_layers = 0, taps = 5, sm_count = 148, precision = 0;
    int ch[PN_MAX_LAYERS + 1] = {0};
    int cin_pad[PN_MAX_LAYERS] = {0}, block_n[PN_MAX_LAYERS] = {0}, n_tiles[PN_MAX_LAYERS] = {0};
    int act_ld = 0;   // row pitch of the hidden activation buffers
    __nv_bfloat16 *w_hi[PN_MAX_LAYERS] = {nullptr}, *w_lo[PN_MAX_LAYERS] = {nullptr};   // [taps][cout][cin_pad]
    float* w_f32[PN_MAX_LAYERS] = {nullptr};                                            // [taps][cin][cout]
    float* bias[PN_MAX_LAYERS] = {nullptr};
    CUtensorMap tm_b_hi[PN_MAX_LAYERS], tm_b_lo[PN_MAX_LAYERS];
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::vector<xdtts_postnet_plan*> cache;
};

struct xdtts_postnet_plan {
    xdtts_postnet* h = nullptr;
    int B = 0, max_T = 0, total_T = 0, m_tiles = 0, rows = 0;
    std::vector<int> Ts, foff, roff;
    int *d_T = nullptr, *d_foff = nullptr, *d_roff = nullptr, *d_row_t = nullptr, *d_row_u = nullptr;
    float *d_mel = nullptr, *d_out = nullptr;                      // caller-layout arenas
    __nv_bfloat16 *xin_hi = nullptr, *xin_lo = nullptr;             // [rows + 4][cin_pad[0]]
    __nv_bfloat16 *act_hi[2] = {nullptr, nullptr}, *act_lo[2] = {nullptr, nullptr};   // [rows + 4][act_ld]
    float *xin_f32 = nullptr, *act_f32[2] = {nullptr, nullptr};     // CUDA-core path
    CUtensorMap tm_a_hi[PN_MAX_LAYERS], tm_a_lo[PN_MAX_LAYERS];
    float *h_in = nullptr, *h_out = nullptr;                        // pinned staging
    cudaEvent_t h_in_busy = nullptr;                                // recorded after the DMA that reads h_in
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

extern "C" void xdtts_postnet_plan_destroy(xdtts_postnet_plan* p);

extern "C" void xdtts_postnet_destroy(xdtts_postnet* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (xdtts_postnet_plan* p : h->cache) xdtts_postnet_plan_destroy(p);
    for (int l = 0; l < PN_MAX_LAYERS; l++) {
        cudaFree(h->w_hi[l]); cudaFree(h->w_lo[l]); cudaFree(h->w_f32[l]); cudaFree(h->bias[l]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int xdtts_postnet_create(int n_layers, const int* channels, int ksize, const float* const* conv_w,
                                    const float* const* conv_b, const float* const* bn_gamma, const float* const* bn_beta,
                                    const float* const* bn_mean, const float* const* bn_var, float eps,
                                    const xdtts_postnet_opts* opts, int device, xdtts_postnet** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: out is null");
    *out = nullptr;
    if (!channels || !conv_w) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: null argument");
    if (n_layers < 1 || n_layers > PN_MAX_LAYERS) return fail(XDTTS_ERR_UNSUPPORTED, "postnet_create: %d layers, supported 1..%d", n_layers, PN_MAX_LAYERS);
    if (ksize != 5) return fail(XDTTS_ERR_UNSUPPORTED, "postnet_create: kernel size %d, the Tacotron2 postnet uses 5", ksize);
    if (!(eps >= 0.f) || !std::isfinite(eps)) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: eps must be finite and >= 0");
    const int precision = opts ? opts->precision : 0;
    if (precision < 0 || precision > 2) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: precision %d not in {0,1,2}", precision);
    for (int l = 0; l <= n_layers; l++)
        if (channels[l] < 1) return fail(XDTTS_ERR_SHAPE, "postnet_create: channels[%d] = %d", l, channels[l]);
    if (channels[0] > PN_MAX_CIN0) return fail(XDTTS_ERR_UNSUPPORTED, "postnet_create: %d input channels > %d", channels[0], PN_MAX_CIN0);
    if (channels[n_layers] != channels[0]) return fail(XDTTS_ERR_SHAPE, "postnet_create: residual needs channels[last] == channels[0] (%d vs %d)", channels[n_layers], channels[0]);
    for (int l = 0; l < n_layers; l++) {
        const int co = channels[l + 1];
        if (co % 16 || co > PN_MAX_COUT) return fail(XDTTS_ERR_UNSUPPORTED, "postnet_create: layer %d has %d output channels; need a multiple of 16, <= %d", l, co, PN_MAX_COUT);
        if (co > PN_MAX_BN && co % 128) return fail(XDTTS_ERR_UNSUPPORTED, "postnet_create: layer %d has %d output channels; above 256 need a multiple of 128", l, co);
        if (!conv_w[l]) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: conv_w[%d] is null", l);
        const bool any_bn = bn_gamma && bn_gamma[l];
        if (any_bn && !(bn_beta && bn_beta[l] && bn_mean && bn_mean[l] && bn_var && bn_var[l]))
            return fail(XDTTS_ERR_BAD_ARG, "postnet_create: layer %d has a partial BatchNorm (gamma without beta/mean/var)", l);
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(XDTTS_ERR_CUDA, "postnet_create: no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= n_dev) return fail(XDTTS_ERR_BAD_ARG, "postnet_create: device %d of %d", device, n_dev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(XDTTS_ERR_CUDA, "postnet_create: device %d is sm_%d%d, this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));

    xdtts_postnet* h = new (std::nothrow) xdtts_postnet();
    if (!h) return fail(XDTTS_ERR_OOM, "postnet_create: out of host memory");
    h->device = device; h->n_layers = n_layers; h->taps = ksize; h->sm_count = prop.multiProcessorCount; h->precision = precision;
    for (int l = 0; l <= n_layers; l++) h->ch[l] = channels[l];
    for (int l = 1; l < n_layers; l++) h->act_ld = std::max(h->act_ld, round_up(channels[l], 64));
    if (h->act_ld == 0) h->act_ld = 64;

    int rc = XDTTS_OK;
    cudaError_t e = cudaSuccess;
    for (int l = 0; l < n_layers && rc == XDTTS_OK && e == cudaSuccess; l++) {
        const int ci = channels[l], co = channels[l + 1], cp = round_up(ci, 64);
        h->cin_pad[l] = cp;
        h->block_n[l] = co <= PN_MAX_BN ? co : (co % PN_MAX_BN == 0 ? PN_MAX_BN : 128);
        h->n_tiles[l] = co / h->block_n[l];
        // fold BatchNorm (eval): W' = W g / sqrt(var + eps), b' = (b - mean) g / sqrt(var + eps) + beta   (fp64)
        std::vector<float> bias(co);
        std::vector<__nv_bfloat16> whi((size_t)ksize * co * cp, __float2bfloat16(0.f)), wlo(whi);
        std::vector<float> wf((size_t)ksize * ci * co);
        const bool bn = bn_gamma && bn_gamma[l];
        for (int o = 0; o < co && rc == XDTTS_OK; o++) {
            double scale = 1.0, shift = 0.0, b = (conv_b && conv_b[l]) ? (double)conv_b[l][o] : 0.0;
            if (bn) {
                const double var = (double)bn_var[l][o];
                if (!(var + (double)eps > 0.0)) { rc = fail(XDTTS_ERR_BAD_ARG, "postnet_create: layer %d channel %d: var + eps <= 0", l, o); break; }
                scale = (double)bn_gamma[l][o] / std::sqrt(var + (double)eps);
                shift = (double)bn_beta[l][o] - (double)bn_mean[l][o] * scale;
            }
            const double bf = b * scale + shift;
            if (!std::isfinite(bf)) { rc = fail(XDTTS_ERR_BAD_ARG, "postnet_create: layer %d channel %d: non-finite folded bias", l, o); break; }
            bias[o] = (float)bf;
            for (int i = 0; i < ci; i++)
                for (int j = 0; j < ksize; j++) {
                    const double w = (double)conv_w[l][((size_t)o * ci + i) * ksize + j] * scale;
                    if (!std::isfinite(w)) { rc = fail(XDTTS_ERR_BAD_ARG, "postnet_create: layer %d: non-finite weight", l); break; }
                    const float wfl = (float)w;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(wfl);
                    whi[((size_t)j * co + o) * cp + i] = hi;
                    wlo[((size_t)j * co + o) * cp + i] = __float2bfloat16_rn(wfl - __bfloat162float(hi));
                    wf[((size_t)j * ci + i) * co + o] = wfl;
                }
        }
        if (rc != XDTTS_OK) break;
        if (e == cudaSuccess) e = cudaMalloc((void**)&h->w_hi[l], whi.size() * 2);
        if (e == cudaSuccess) e = cudaMalloc((void**)&h->w_lo[l], wlo.size() * 2);
        if (e == cudaSuccess) e = cudaMalloc((void**)&h->w_f32[l], wf.size() * 4);
        if (e == cudaSuccess) e = cudaMalloc((void**)&h->bias[l], bias.size() * 4);
        if (e == cudaSuccess) e = cudaMemcpy(h->w_hi[l], whi.data(), whi.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->w_lo[l], wlo.data(), wlo.size() * 2, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->w_f32[l], wf.data(), wf.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->bias[l], bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && precision != 2) {
            rc = make_map(&h->tm_b_hi[l], h->w_hi[l], cp, (long long)ksize * co, cp, h->block_n[l]);
            if (rc == XDTTS_OK) rc = make_map(&h->tm_b_lo[l], h->w_lo[l], cp, (long long)ksize * co, cp, h->block_n[l]);
        }
    }
    if (rc == XDTTS_OK && e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (rc == XDTTS_OK && e == cudaSuccess) e = pn_prepare();
    if (rc != XDTTS_OK || e != cudaSuccess) {
        xdtts_postnet_destroy(h);
        if (rc != XDTTS_OK) return rc;
        return fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "postnet_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return XDTTS_OK;
}

// ------------------------------------------------------------------ plan
extern "C" void xdtts_postnet_plan_destroy(xdtts_postnet_plan* p) {
    if (!p) return;
    cudaSetDevice(p->h->device);
    cudaFree(p->d_T); cudaFree(p->d_foff); cudaFree(p->d_roff); cudaFree(p->d_row_t); cudaFree(p->d_row_u);
    cudaFree(p->d_mel); cudaFree(p->d_out); cudaFree(p->xin_hi); cudaFree(p->xin_lo); cudaFree(p->xin_f32);
    for (int i = 0; i < 2; i++) {
        cudaFree(p->act_hi[i]); cudaFree(p->act_lo[i]); cudaFree(p->act_f32[i]);
        if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    }
    if (p->h_in) cudaFreeHost(p->h_in);
    if (p->h_in_busy) cudaEventDestroy(p->h_in_busy);
    if (p->h_out) cudaFreeHost(p->h_out);
    delete p;
}

static int pn_plan_build(xdtts_postnet* h, const int* Ts, int B, xdtts_postnet_plan** out) {
    *out = nullptr;
    if (!Ts || B < 1) return fail(XDTTS_ERR_BAD_ARG, "postnet plan: need B >= 1 utterances");
    long long total = 0, rows_ll = 0;
    for (int b = 0; b < B; b++) {
        if (Ts[b] < 1) return fail(XDTTS_ERR_SHAPE, "postnet plan: utterance %d has T = %d frames, need >= 1", b, Ts[b]);
        total += Ts[b];
        rows_ll += Ts[b] + 2;
    }
    if (rows_ll * (long long)h->act_ld >= (1ll << 31)) return fail(XDTTS_ERR_SHAPE, "postnet plan: %lld frames in one batch is too many (split it)", total);
    CU(cudaSetDevice(h->device));
    clear_stale_error(__func__);
    xdtts_postnet_plan* p = new (std::nothrow) xdtts_postnet_plan();
    if (!p) return fail(XDTTS_ERR_OOM, "postnet plan: out of host memory");
    p->h = h; p->B = B; p->Ts.assign(Ts, Ts + B); p->total_T = (int)total;
    p->m_tiles = (int)((rows_ll + PN_BM - 1) / PN_BM);
    p->rows = p->m_tiles * PN_BM;
    std::vector<int> row_t(p->rows, -1), row_u(p->rows, 0);
    int f = 0, r = 0;
    for (int b = 0; b < B; b++) {
        p->foff.push_back(f);
        p->roff.push_back(r);
        for (int t = 0; t < Ts[b]; t++) { row_t[r + t] = t; row_u[r + t] = b; }
        f += Ts[b];
        r += Ts[b] + 2;   // two zero rows = the convolution's padding between utterances
        p->max_T = std::max(p->max_T, Ts[b]);
    }
    const size_t C0 = h->ch[0], brow = (size_t)p->rows + 4;
    const bool tc = h->precision != 2, split = h->precision == 0;
    cudaError_t e = cudaSuccess;
#define ALLOC(ptr, bytes) if (e == cudaSuccess) e = cudaMalloc((void**)&(ptr), (bytes))
#define ZALLOC(ptr, bytes) ALLOC(ptr, bytes); if (e == cudaSuccess) e = cudaMemset((ptr), 0, (bytes))
    ALLOC(p->d_T, B * sizeof(int));
    ALLOC(p->d_foff, B * sizeof(int));
    ALLOC(p->d_roff, B * sizeof(int));
    ALLOC(p->d_row_t, p->rows * sizeof(int));
    ALLOC(p->d_row_u, p->rows * sizeof(int));
    ALLOC(p->d_mel, C0 * (size_t)total * 4);
    ALLOC(p->d_out, C0 * (size_t)total * 4);
    if (tc) {
        ZALLOC(p->xin_hi, brow * h->cin_pad[0] * 2);
        ZALLOC(p->act_hi[0], brow * h->act_ld * 2);
        ZALLOC(p->act_hi[1], brow * h->act_ld * 2);
        if (split) {
            ZALLOC(p->xin_lo, brow * h->cin_pad[0] * 2);
            ZALLOC(p->act_lo[0], brow * h->act_ld * 2);
            ZALLOC(p->act_lo[1], brow * h->act_ld * 2);
        }
    } else {
        ZALLOC(p->xin_f32, brow * h->cin_pad[0] * 4);
        ZALLOC(p->act_f32[0], brow * h->act_ld * 4);
        ZALLOC(p->act_f32[1], brow * h->act_ld * 4);
    }
#undef ZALLOC
#undef ALLOC
    if (e == cudaSuccess) e = cudaMemcpy(p->d_T, Ts, B * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_foff, p->foff.data(), B * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_roff, p->roff.data(), B * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_row_t, row_t.data(), p->rows * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_row_u, row_u.data(), p->rows * sizeof(int), cudaMemcpyHostToDevice);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&p->ev[i]);
    int rc = XDTTS_OK;
    if (e == cudaSuccess && tc) {
        for (int l = 0; l < h->n_layers && rc == XDTTS_OK; l++) {
            const __nv_bfloat16* hi = l == 0 ? p->xin_hi : p->act_hi[(l - 1) & 1];
            const __nv_bfloat16* lo = l == 0 ? p->xin_lo : p->act_lo[(l - 1) & 1];
            const int ld = l == 0 ? h->cin_pad[0] : h->act_ld;
            rc = make_map(&p->tm_a_hi[l], hi, h->cin_pad[l], (long long)brow, ld, PN_BM);
            if (rc == XDTTS_OK) rc = make_map(&p->tm_a_lo[l], split ? lo : hi, h->cin_pad[l], (long long)brow, ld, PN_BM);
        }
    }
    if (e != cudaSuccess || rc != XDTTS_OK) {
        xdtts_postnet_plan_destroy(p);
        if (rc != XDTTS_OK) return rc;
        return fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "postnet plan: %s", cudaGetErrorString(e));
    }
    *out = p;
    return XDTTS_OK;
}

extern "C" int xdtts_postnet_plan_create(xdtts_postnet* h, const int* Ts, int B, xdtts_postnet_plan** out) {
    if (!h || !out) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_create: null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    return pn_plan_build(h, Ts, B, out);
}

static int pn_plan_upload_locked(xdtts_postnet_plan* p, const float* const* srcs, cudaStream_t s) {
    xdtts_postnet* h = p->h;
    clear_stale_error(__func__);
    if (!srcs) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_upload: srcs is null");
    for (int b = 0; b < p->B; b++)
        if (!srcs[b]) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_upload: srcs[%d] is null", b);
    CU(cudaSetDevice(h->device));
    const size_t C0 = h->ch[0];
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(srcs[b]);
    if (!all_pinned) {   // pageable: staged by the copy threads with non-temporal stores, utterance by utterance (host_copy.h)
        if (p->h_in_busy) CU(cudaEventSynchronize(p->h_in_busy));   // the previous upload's DMA still reads the staging buffer
        if (!p->h_in) CU(cudaHostAlloc((void**)&p->h_in, C0 * (size_t)p->total_T * 4, cudaHostAllocDefault));
        if (!p->h_in_busy) CU(cudaEventCreateWithFlags(&p->h_in_busy, cudaEventDisableTiming));
        std::vector<std::atomic<int>> pending(p->B);
        for (int b = 0; b < p->B; b++) {
            pending[b].store(0);
            host_copy_async(p->h_in + C0 * (size_t)p->foff[b], srcs[b], C0 * (size_t)p->Ts[b] * 4, &pending[b], true);
        }
        cudaError_t ce = cudaSuccess;
        for (int b = 0; b < p->B; b++) {
            host_copy_wait(&pending[b]);
            if (ce == cudaSuccess)
                ce = cudaMemcpyAsync(p->d_mel + C0 * (size_t)p->foff[b], p->h_in + C0 * (size_t)p->foff[b], C0 * (size_t)p->Ts[b] * 4,
                                     cudaMemcpyHostToDevice, s);
        }
        CU(ce);
        CU(cudaEventRecord(p->h_in_busy, s));
    } else {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(p->d_mel + C0 * (size_t)p->foff[b], srcs[b], C0 * (size_t)p->Ts[b] * 4, cudaMemcpyHostToDevice, s));
    }
    return XDTTS_OK;
}

extern "C" int xdtts_postnet_plan_upload(xdtts_postnet_plan* p, const float* const* mels) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_upload: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return pn_plan_upload_locked(p, mels, p->h->stream);
}

// enqueue input staging + the layers on stream s; the result goes to `dst` (caller layout, same
// frame offsets as the plan: the plan's own d_out, or the mel arena of a Griffin-Lim plan)
static int pn_enqueue(xdtts_postnet_plan* p, float* dst, cudaStream_t s) {
    xdtts_postnet* h = p->h;
    const bool tc = h->precision != 2, split = h->precision == 0;
    CU(pn_launch_stage_input(p->d_mel, p->d_T, p->d_foff, p->d_roff, p->B, p->max_T, h->ch[0], h->cin_pad[0],
                             tc ? p->xin_hi : nullptr, split ? p->xin_lo : nullptr, tc ? nullptr : p->xin_f32, s));
    g_launches++;
    for (int l = 0; l < h->n_layers; l++) {
        PnLayer q;
        memset(&q, 0, sizeof(q));
        q.m_tiles = p->m_tiles; q.n_tiles = h->n_tiles[l]; q.block_n = h->block_n[l];
        q.cin_chunks = h->cin_pad[l] / PN_BK; q.cout = h->ch[l + 1];
        q.n_split = split ? 3 : 1; q.taps = h->taps;
        q.last = (l == h->n_layers - 1); q.apply_tanh = !q.last;
        q.bias = h->bias[l];
        q.out_hi = p->act_hi[l & 1]; q.out_lo = split ? p->act_lo[l & 1] : nullptr; q.out_ld = h->act_ld;
        q.row_t = p->d_row_t; q.row_u = p->d_row_u; q.utt_T = p->d_T; q.utt_foff = p->d_foff;
        q.resid = p->d_mel; q.out_f32 = dst;
        if (tc) {
            CUtensorMap maps[4] = {p->tm_a_hi[l], p->tm_a_lo[l], h->tm_b_hi[l], split ? h->tm_b_lo[l] : h->tm_b_hi[l]};
            CU(pn_launch_conv_tc(maps, q, h->sm_count, s));
        } else {
            const float* x = l == 0 ? p->xin_f32 : p->act_f32[(l - 1) & 1];
            CU(pn_launch_conv_f32(x, l == 0 ? h->cin_pad[0] : h->act_ld, h->w_f32[l], q, h->ch[l], p->act_f32[l & 1], h->act_ld, s));
        }
        g_launches++;
    }
    return XDTTS_OK;
}

static int pn_plan_run_locked(xdtts_postnet_plan* p, float* dst, cudaStream_t s, float* ms_total) {
    clear_stale_error(__func__);
    CU(cudaSetDevice(p->h->device));
    CU(cudaEventRecord(p->ev[0], s));
    int rc = pn_enqueue(p, dst, s);
    if (rc) return rc;
    CU(cudaEventRecord(p->ev[1], s));
    CU(cudaStreamSynchronize(s));
    if (ms_total) CU(cudaEventElapsedTime(ms_total, p->ev[0], p->ev[1]));
    return XDTTS_OK;
}

extern "C" int xdtts_postnet_plan_run(xdtts_postnet_plan* p, xdtts_gl_plan* feed, float* ms_total) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_run: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    if (!feed) return pn_plan_run_locked(p, p->d_out, p->h->stream, ms_total);
    // write "mel_outputs_postnet" straight into the vocoder plan's mel arena, on the vocoder's stream
    xdtts_gl* g = feed->h;
    std::lock_guard<std::mutex> lk2(g->mu);
    if (g->device != p->h->device) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_run: the vocoder plan lives on device %d, the postnet on %d", g->device, p->h->device);
    if (g->n_mels != p->h->ch[0] || feed->B != p->B || memcmp(feed->Ts.data(), p->Ts.data(), sizeof(int) * p->B) != 0)
        return fail(XDTTS_ERR_SHAPE, "postnet_plan_run: the vocoder plan has a different batch shape");
    float* arena = nullptr;
    int rc = gl_plan_mel_arena(feed, &arena);
    if (rc) return rc;
    CU(cudaStreamSynchronize(p->h->stream));   // uploads of this plan ran on the postnet's stream
    return pn_plan_run_locked(p, arena, g->stream, ms_total);
}

static int pn_plan_download_locked(xdtts_postnet_plan* p, const float* src, float* const* outs, cudaStream_t s) {
    xdtts_postnet* h = p->h;
    clear_stale_error(__func__);
    if (!outs) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_download: outs is null");
    for (int b = 0; b < p->B; b++)
        if (!outs[b]) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_download: outs[%d] is null", b);
    CU(cudaSetDevice(h->device));
    const size_t C0 = h->ch[0];
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(outs[b]);
    if (all_pinned) {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(outs[b], src + C0 * (size_t)p->foff[b], C0 * (size_t)p->Ts[b] * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    } else {
        if (!p->h_out) CU(cudaHostAlloc((void**)&p->h_out, C0 * (size_t)p->total_T * 4, cudaHostAllocDefault));
        CU(cudaMemcpyAsync(p->h_out, src, C0 * (size_t)p->total_T * 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        for (int b = 0; b < p->B; b++) memcpy(outs[b], p->h_out + C0 * (size_t)p->foff[b], C0 * (size_t)p->Ts[b] * 4);
    }
    return XDTTS_OK;
}

extern "C" int xdtts_postnet_plan_download(xdtts_postnet_plan* p, float* const* outs) {
    if (!p) return fail(XDTTS_ERR_BAD_ARG, "postnet_plan_download: plan is null");
    std::lock_guard<std::mutex> lk(p->h->mu);
    return pn_plan_download_locked(p, p->d_out, outs, p->h->stream);
}

// ------------------------------------------------------------------ batch entry points
static int pn_cached_plan(xdtts_postnet* h, const int* Ts, int B, xdtts_postnet_plan** out) {
    xdtts_postnet_plan* p = nullptr;
    for (xdtts_postnet_plan* c : h->cache)
        if (c->B == B && memcmp(c->Ts.data(), Ts, sizeof(int) * B) == 0) p = c;
    if (!p) {
        if (h->cache.size() >= 4) {
            xdtts_postnet_plan_destroy(h->cache.front());
            h->cache.erase(h->cache.begin());
        }
        int rc = pn_plan_build(h, Ts, B, &p);
        if (rc) return rc;
        h->cache.push_back(p);
    }
    *out = p;
    return XDTTS_OK;
}

extern "C" int xdtts_postnet_infer_batch(xdtts_postnet* h, const float* const* mels, const int* Ts, int B, float* const* outs) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "postnet_infer: handle is null");
    if (!mels || !Ts || !outs) return fail(XDTTS_ERR_BAD_ARG, "postnet_infer: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "postnet_infer: B = %d", B);
    std::lock_guard<std::mutex> lk(h->mu);
    xdtts_postnet_plan* p = nullptr;
    int rc = pn_cached_plan(h, Ts, B, &p);
    if (rc) return rc;
    rc = pn_plan_upload_locked(p, mels, h->stream);
    if (rc) return rc;
    rc = pn_plan_run_locked(p, p->d_out, h->stream, nullptr);
    if (rc) return rc;
    return pn_plan_download_locked(p, p->d_out, outs, h->stream);
}

extern "C" int xdtts_postnet_infer(xdtts_postnet* h, const float* mel, int T, float* out) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "postnet_infer: handle is null");
    if (!mel || !out) return fail(XDTTS_ERR_BAD_ARG, "postnet_infer: null argument");
    if (T < 1) return fail(XDTTS_ERR_SHAPE, "postnet_infer: T = %d, need >= 1 frame", T);
    const float* mels[1] = {mel};
    float* outs[1] = {out};
    return xdtts_postnet_infer_batch(h, mels, &T, 1, outs);
}

// postnet -> lift -> Griffin-Lim for B utterances without the intermediate mel leaving HBM
// (XdTts::infer: self.model.infer(..) tail + self.vocoder.infer(&spectrogram), src/lib.rs:123,141)
extern "C" int xdtts_tail_infer_batch(xdtts_postnet* pn, xdtts_gl* gl, const float* const* mels, const int* Ts, int B,
                                      const float* const* init_phases, float* const* out_mels, float* const* out_waves) {
    if (!pn || !gl) return fail(XDTTS_ERR_BAD_ARG, "tail_infer: handle is null");
    if (!mels || !Ts || !out_waves) return fail(XDTTS_ERR_BAD_ARG, "tail_infer: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "tail_infer: B = %d", B);
    if (pn->device != gl->device) return fail(XDTTS_ERR_BAD_ARG, "tail_infer: postnet on device %d, vocoder on %d", pn->device, gl->device);
    if (pn->ch[0] != gl->n_mels) return fail(XDTTS_ERR_SHAPE, "tail_infer: postnet has %d mel channels, vocoder %d", pn->ch[0], gl->n_mels);
    std::lock_guard<std::mutex> lk(pn->mu);
    std::lock_guard<std::mutex> lk2(gl->mu);
    xdtts_postnet_plan* pp = nullptr;
    xdtts_gl_plan* gp = nullptr;
    int rc = pn_cached_plan(pn, Ts, B, &pp);
    if (rc) return rc;
    rc = gl_cached_plan(gl, Ts, B, &gp);      // also rejects T < 2
    if (rc) return rc;
    float* arena = nullptr;
    rc = gl_plan_mel_arena(gp, &arena);
    if (rc) return rc;
    cudaStream_t s = gl->stream;
    rc = pn_plan_upload_locked(pp, mels, s);
    if (rc) return rc;
    CU(cudaSetDevice(pn->device));
    rc = pn_enqueue(pp, arena, s);
    if (rc) return rc;
    int flags = 0;
    if (init_phases) {
        rc = gl_plan_upload_locked(gp, 2, init_phases, nullptr);
        if (rc) return rc;
        flags |= XDTTS_RUN_USE_PHASE;
    }
    rc = gl_plan_run_locked(gp, flags, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (out_mels) {
        rc = pn_plan_download_locked(pp, arena, out_mels, s);
        if (rc) return rc;
    }
    return gl_plan_download_locked(gp, out_waves);
}

// ------------------------------------------------------------------ streaming pipeline
// XdTts::infer handles one utterance after another (src/lib.rs:83-104 loops over the chunks of the
// input text) and every call ends with the samples on the host.  Served in batches, that pattern leaves
// the GPU idle during each batch's PCIe copies.  A pipe owns `depth` device-resident slots of one batch
// shape, each with its own stream: push() enqueues H2D -> [postnet] -> lift -> Griffin-Lim -> D2H of one
// batch and returns; the copies of one batch then overlap the kernels of its neighbours.
struct xdtts_pipe {
    xdtts_gl* gl = nullptr;
    xdtts_postnet* pn = nullptr;
    int B = 0, depth = 0, head = 0;
    std::vector<int> Ts;
    struct Slot {
        xdtts_gl_plan* gp = nullptr;
        xdtts_postnet_plan* pp = nullptr;
        cudaStream_t s = nullptr;
        cudaEvent_t computed = nullptr;   // recorded after the slot's kernels, before its device -> host copies
        bool busy = false, staged_wave = false, staged_mel = false;
        std::vector<float*> out_waves, out_mels;
    };
    std::vector<Slot> slots;
    cudaEvent_t last_computed = nullptr;   // the most recently pushed batch's `computed`
    std::mutex mu;
};

// device -> host copy of "mel_outputs_postnet" on stream s without waiting (pinned: direct; pageable: staged)
static int pn_download_async(xdtts_postnet_plan* p, const float* src, float* const* outs, cudaStream_t s, bool* staged) {
    xdtts_postnet* h = p->h;
    const size_t C0 = h->ch[0];
    for (int b = 0; b < p->B; b++)
        if (!outs[b]) return fail(XDTTS_ERR_BAD_ARG, "pipe_push: out_mels[%d] is null", b);
    bool all_pinned = true;
    for (int b = 0; b < p->B; b++) all_pinned = all_pinned && is_pinned(outs[b]);
    *staged = !all_pinned;
    if (all_pinned) {
        for (int b = 0; b < p->B; b++)
            CU(cudaMemcpyAsync(outs[b], src + C0 * (size_t)p->foff[b], C0 * (size_t)p->Ts[b] * 4, cudaMemcpyDeviceToHost, s));
    } else {
        if (!p->h_out) CU(cudaHostAlloc((void**)&p->h_out, C0 * (size_t)p->total_T * 4, cudaHostAllocDefault));
        CU(cudaMemcpyAsync(p->h_out, src, C0 * (size_t)p->total_T * 4, cudaMemcpyDeviceToHost, s));
    }
    return XDTTS_OK;
}

// wait for the batch in `sl` and finish the staged copies of pageable destinations
static int pipe_collect(xdtts_pipe* q, xdtts_pipe::Slot& sl) {
    if (!sl.busy) return XDTTS_OK;
    CU(cudaSetDevice(q->gl->device));
    CU(cudaStreamSynchronize(sl.s));     // on failure the slot stays busy: the batch is not silently dropped
    sl.busy = false;
    if (sl.staged_wave) gl_plan_download_finish(sl.gp, sl.out_waves.data());
    if (sl.staged_mel) {
        const size_t C0 = q->pn->ch[0];
        for (int b = 0; b < q->B; b++)
            memcpy(sl.out_mels[b], sl.pp->h_out + C0 * (size_t)sl.pp->foff[b], C0 * (size_t)sl.pp->Ts[b] * 4);
    }
    return XDTTS_OK;
}

extern "C" void xdtts_pipe_destroy(xdtts_pipe* q) {
    if (!q) return;
    cudaSetDevice(q->gl->device);
    for (auto& sl : q->slots) {
        if (sl.s) cudaStreamSynchronize(sl.s);
        if (sl.gp) xdtts_gl_plan_destroy(sl.gp);
        if (sl.pp) xdtts_postnet_plan_destroy(sl.pp);
        if (sl.computed) cudaEventDestroy(sl.computed);
        if (sl.s) cudaStreamDestroy(sl.s);
    }
    cudaGetLastError();
    delete q;
}

extern "C" int xdtts_pipe_create(xdtts_gl* gl, xdtts_postnet* pn, const int* Ts, int B, int depth, xdtts_pipe** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "pipe_create: out is null");
    *out = nullptr;
    if (!gl || !Ts) return fail(XDTTS_ERR_BAD_ARG, "pipe_create: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "pipe_create: B = %d", B);
    if (depth < 1 || depth > 8) return fail(XDTTS_ERR_BAD_ARG, "pipe_create: depth = %d, supported: 1..8", depth);
    if (pn && pn->device != gl->device) return fail(XDTTS_ERR_BAD_ARG, "pipe_create: postnet on device %d, vocoder on %d", pn->device, gl->device);
    if (pn && pn->ch[0] != gl->n_mels) return fail(XDTTS_ERR_SHAPE, "pipe_create: postnet has %d mel channels, vocoder %d", pn->ch[0], gl->n_mels);
    xdtts_pipe* q = new (std::nothrow) xdtts_pipe();
    if (!q) return fail(XDTTS_ERR_OOM, "pipe_create: out of host memory");
    q->gl = gl; q->pn = pn; q->B = B; q->depth = depth; q->Ts.assign(Ts, Ts + B);
    q->slots.resize(depth);
    int rc = XDTTS_OK;
    for (int i = 0; i < depth && rc == XDTTS_OK; i++) {
        xdtts_pipe::Slot& sl = q->slots[i];
        {
            std::lock_guard<std::mutex> lk(gl->mu);
            rc = gl_plan_build(gl, Ts, B, &sl.gp);      // also rejects T < 2
        }
        if (rc == XDTTS_OK && pn) {
            std::lock_guard<std::mutex> lk(pn->mu);
            rc = pn_plan_build(pn, Ts, B, &sl.pp);
        }
        if (rc == XDTTS_OK) {
            float* arena = nullptr;
            rc = gl_plan_mel_arena(sl.gp, &arena);
        }
        if (rc == XDTTS_OK && (cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) != cudaSuccess ||
                               cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming) != cudaSuccess))
            rc = fail(XDTTS_ERR_CUDA, "pipe_create: cudaStreamCreate / cudaEventCreate failed");
        sl.out_waves.resize(B);
        sl.out_mels.resize(B);
    }
    if (rc != XDTTS_OK) {
        xdtts_pipe_destroy(q);
        return rc;
    }
    *out = q;
    return XDTTS_OK;
}

extern "C" int xdtts_pipe_push(xdtts_pipe* q, const float* const* mels, const float* const* init_phases,
                               float* const* out_mels, float* const* out_waves) {
    if (!q) return fail(XDTTS_ERR_BAD_ARG, "pipe_push: pipe is null");
    if (!mels || !out_waves) return fail(XDTTS_ERR_BAD_ARG, "pipe_push: null argument");
    if (out_mels && !q->pn) return fail(XDTTS_ERR_BAD_ARG, "pipe_push: out_mels needs a pipe with a postnet");
    for (int b = 0; b < q->B; b++)
        if (!mels[b] || !out_waves[b] || (init_phases && !init_phases[b]))
            return fail(XDTTS_ERR_BAD_ARG, "pipe_push: null buffer for utterance %d", b);
    std::lock_guard<std::mutex> lk(q->mu);
    xdtts_pipe::Slot& sl = q->slots[q->head];
    int rc = pipe_collect(q, sl);      // the oldest batch in flight owns this slot: wait for it
    if (rc) return rc;
    CU(cudaSetDevice(q->gl->device));
    clear_stale_error(__func__);
    float* arena = nullptr;
    rc = gl_plan_mel_arena(sl.gp, &arena);
    if (rc) return rc;
    int flags = 0;
    // From here on work may already be queued on the slot's stream: every failure goes through the clean-up below
    // (wait for the stream) before returning, so nothing of a half-submitted batch touches the caller's buffers later.
    auto cuda_rc = [&](cudaError_t e, const char* what) {
        return e == cudaSuccess ? XDTTS_OK : fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "pipe_push: %s: %s", what, cudaGetErrorString(e));
    };
    // The copies of neighbouring batches overlap kernels, but the KERNELS of two batches must not interleave: every
    // Griffin-Lim launch is sized to fill the device in one wave, and two such grids sharing the SMs finish later than
    // the same two grids back to back.  So a slot's kernels wait for the previous batch's kernels (not for its copies).
    if (q->pn) {
        std::lock_guard<std::mutex> lk2(q->pn->mu);
        rc = pn_plan_upload_locked(sl.pp, mels, sl.s);
        if (rc == XDTTS_OK && q->last_computed) rc = cuda_rc(cudaStreamWaitEvent(sl.s, q->last_computed, 0), "cudaStreamWaitEvent");
        if (rc == XDTTS_OK) rc = pn_enqueue(sl.pp, arena, sl.s);
    } else {
        std::lock_guard<std::mutex> lk2(q->gl->mu);
        rc = gl_plan_upload_locked(sl.gp, 0, mels, sl.s);
    }
    if (rc == XDTTS_OK) {
        std::lock_guard<std::mutex> lk2(q->gl->mu);
        if (init_phases) {
            rc = gl_plan_upload_locked(sl.gp, 2, init_phases, sl.s);
            flags |= XDTTS_RUN_USE_PHASE;
        }
        if (rc == XDTTS_OK && !q->pn && q->last_computed) rc = cuda_rc(cudaStreamWaitEvent(sl.s, q->last_computed, 0), "cudaStreamWaitEvent");
        if (rc == XDTTS_OK) rc = gl_plan_launch_async(sl.gp, flags, sl.s);
        if (rc == XDTTS_OK) rc = cuda_rc(cudaEventRecord(sl.computed, sl.s), "cudaEventRecord");
        if (rc == XDTTS_OK) q->last_computed = sl.computed;
        if (rc == XDTTS_OK) rc = gl_plan_download_async(sl.gp, out_waves, sl.s, &sl.staged_wave);
    }
    sl.staged_mel = false;
    if (rc == XDTTS_OK && out_mels) rc = pn_download_async(sl.pp, arena, out_mels, sl.s, &sl.staged_mel);
    if (rc) {
        const std::string msg = xdtts_last_error();
        cudaStreamSynchronize(sl.s);   // leave nothing of a half-submitted batch in flight
        cudaGetLastError();
        return fail(rc, "%s", msg.c_str());
    }
    for (int b = 0; b < q->B; b++) {
        sl.out_waves[b] = out_waves[b];
        sl.out_mels[b] = out_mels ? out_mels[b] : nullptr;
    }
    sl.busy = true;
    q->head = (q->head + 1) % q->depth;
    return XDTTS_OK;
}

extern "C" int xdtts_pipe_flush(xdtts_pipe* q) {
    if (!q) return fail(XDTTS_ERR_BAD_ARG, "pipe_flush: pipe is null");
    std::lock_guard<std::mutex> lk(q->mu);
    int rc = XDTTS_OK;
    for (int i = 0; i < q->depth; i++) {   // oldest first
        int r = pipe_collect(q, q->slots[(q->head + i) % q->depth]);
        if (r && !rc) rc = r;
    }
    return rc;
}

extern "C" int xdtts_pipe_pop(xdtts_pipe* q) {
    if (!q) return fail(XDTTS_ERR_BAD_ARG, "pipe_pop: pipe is null");
    std::lock_guard<std::mutex> lk(q->mu);
    for (int i = 0; i < q->depth; i++) {   // oldest first
        xdtts_pipe::Slot& sl = q->slots[(q->head + i) % q->depth];
        if (!sl.busy) continue;
        int rc = pipe_collect(q, sl);
        return rc ? rc : 1;
    }
    return 0;
}

extern "C" int xdtts_pipe_pending(xdtts_pipe* q) {
    if (!q) return fail(XDTTS_ERR_BAD_ARG, "pipe_pending: pipe is null");
    std::lock_guard<std::mutex> lk(q->mu);
    int n = 0;
    for (auto& sl : q->slots) n += sl.busy ? 1 : 0;
    return n;
}
