// C ABI of the Tacotron2 decoder loop (include/xdtts_b200.h, xdtts_decoder_*): replaces the `decoder`
// ort::Session and the loop around it in Tacotron2::run_decoder (/root/reference src/tacotron2/mod.rs:272-342;
// state shapes DecoderState::new :204-238; session load :251-254).  Host side: weight re-layout for the
// persistent kernel (decoder.cu), workspaces, batching in groups of DC_MAX_NB utterances, output transposition.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "api_internal.h"
#include "decoder.h"

using namespace xdtts;
#define fail xdtts::set_error

struct xdtts_decoder {
    int device = 0, grid = 0, max_steps = 1000, dropout = 1;
    float gate_threshold = 0.6f;
    unsigned long long seed = 0;
    // weights
    float *p1T = nullptr, *p2 = nullptr, *Wa = nullptr, *ba = nullptr, *Wq = nullptr, *v = nullptr, *Weff = nullptr,
          *Wd = nullptr, *bd = nullptr, *Wp = nullptr, *bp = nullptr;
    // workspace for one group of DC_MAX_NB utterances
    int ws_t_enc = 0;
    dc_cell* state = nullptr;
    float *memory = nullptr, *pm = nullptr, *mel_out = nullptr, *gate_out = nullptr, *align_out = nullptr,
          *mel_T = nullptr;
    int *t_len = nullptr, *n_frames = nullptr;
    int* err = nullptr;             // [0] poll error flag, [1] grid barrier counter
    float *h_stage = nullptr;          // pinned staging for pageable callers / results
    size_t h_stage_floats = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0.f;
    int last_steps = 0;
    std::mutex mu;
};

static cudaError_t upload(float** dst, const std::vector<float>& src) {
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice);
    return e;
}

extern "C" void xdtts_decoder_destroy(xdtts_decoder* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    float* bufs[] = {h->p1T, h->p2, h->Wa, h->ba, h->Wq, h->v, h->Weff, h->Wd, h->bd, h->Wp, h->bp,
                     h->memory, h->pm, h->mel_out, h->gate_out, h->align_out, h->mel_T};
    for (float* b : bufs) cudaFree(b);
    cudaFree(h->state);
    cudaFree(h->t_len); cudaFree(h->n_frames); cudaFree(h->err);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

extern "C" int xdtts_decoder_create(const xdtts_decoder_weights* w, const xdtts_decoder_opts* opts, int device, xdtts_decoder** out) {
    if (!out) return fail(XDTTS_ERR_BAD_ARG, "decoder_create: out is null");
    *out = nullptr;
    if (!w) return fail(XDTTS_ERR_BAD_ARG, "decoder_create: weights is null");
    struct { const float* p; size_t n; const char* name; } ts[] = {
        {w->prenet1, (size_t)DC_PRE * DC_MEL, "prenet1"}, {w->prenet2, (size_t)DC_PRE * DC_PRE, "prenet2"},
        {w->att_w_ih, (size_t)4 * DC_RNN * (DC_PRE + DC_ENC), "att_w_ih"}, {w->att_w_hh, (size_t)4 * DC_RNN * DC_RNN, "att_w_hh"},
        {w->att_b_ih, (size_t)4 * DC_RNN, "att_b_ih"}, {w->att_b_hh, (size_t)4 * DC_RNN, "att_b_hh"},
        {w->query, (size_t)DC_ATT * DC_RNN, "query"}, {w->v, (size_t)DC_ATT, "v"},
        {w->loc_conv, (size_t)DC_LOCF * 2 * DC_LOCK, "loc_conv"}, {w->loc_dense, (size_t)DC_ATT * DC_LOCF, "loc_dense"},
        {w->dec_w_ih, (size_t)4 * DC_RNN * (DC_RNN + DC_ENC), "dec_w_ih"}, {w->dec_w_hh, (size_t)4 * DC_RNN * DC_RNN, "dec_w_hh"},
        {w->dec_b_ih, (size_t)4 * DC_RNN, "dec_b_ih"}, {w->dec_b_hh, (size_t)4 * DC_RNN, "dec_b_hh"},
        {w->proj_w, (size_t)DC_MEL * DC_ZP, "proj_w"}, {w->proj_b, (size_t)DC_MEL, "proj_b"},
        {w->gate_w, (size_t)DC_ZP, "gate_w"}, {w->gate_b, (size_t)1, "gate_b"}};
    for (auto& t : ts) {
        if (!t.p) return fail(XDTTS_ERR_BAD_ARG, "decoder_create: weights.%s is null", t.name);
        for (size_t i = 0; i < t.n; i++)
            if (!std::isfinite(t.p[i])) return fail(XDTTS_ERR_BAD_ARG, "decoder_create: weights.%s has a non-finite entry", t.name);
    }
    if (opts && (opts->max_steps < 0 || opts->prenet_dropout < 0 || opts->prenet_dropout > 1 || !(opts->gate_threshold >= 0.f) ||
                 opts->gate_threshold >= 1.f))
        return fail(XDTTS_ERR_BAD_ARG, "decoder_create: option out of range");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(XDTTS_ERR_CUDA, "decoder_create: no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= n_dev) return fail(XDTTS_ERR_BAD_ARG, "decoder_create: device %d of %d", device, n_dev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(XDTTS_ERR_CUDA, "decoder_create: device %d is sm_%d%d, this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    clear_stale_error(__func__);

    xdtts_decoder* h = new (std::nothrow) xdtts_decoder();
    if (!h) return fail(XDTTS_ERR_OOM, "decoder_create: out of host memory");
    h->device = device;
    if (opts) {
        if (opts->gate_threshold > 0.f) h->gate_threshold = opts->gate_threshold;
        if (opts->max_steps > 0) h->max_steps = opts->max_steps;
        h->dropout = opts->prenet_dropout == 0;
        h->seed = opts->seed;
    }
    // ---- re-layout (see DecParams)
    std::vector<float> p1T((size_t)DC_MEL * DC_PRE), Wa((size_t)4 * DC_RNN * DC_ZA), ba((size_t)4 * DC_RNN),
        Weff((size_t)DC_ATT * 2 * DC_LOCK), Wd((size_t)4 * DC_RNN * DC_ZD), bd((size_t)4 * DC_RNN),
        Wp((size_t)(DC_MEL + 1) * DC_ZP), bp(DC_MEL + 1);
    for (int r = 0; r < DC_PRE; r++)
        for (int k = 0; k < DC_MEL; k++) p1T[(size_t)k * DC_PRE + r] = w->prenet1[(size_t)r * DC_MEL + k];
    const int ih_a = DC_PRE + DC_ENC, ih_d = DC_RNN + DC_ENC;
    for (int row = 0; row < 4 * DC_RNN; row++) {
        float* a = &Wa[(size_t)row * DC_ZA];
        memcpy(a, w->att_w_ih + (size_t)row * ih_a + DC_PRE, DC_ENC * sizeof(float));                 // ctx
        memcpy(a + DC_ENC, w->att_w_hh + (size_t)row * DC_RNN, DC_RNN * sizeof(float));               // h_att
        memcpy(a + DC_ENC + DC_RNN, w->att_w_ih + (size_t)row * ih_a, DC_PRE * sizeof(float));        // prenet
        ba[row] = w->att_b_ih[row] + w->att_b_hh[row];
        float* d = &Wd[(size_t)row * DC_ZD];
        memcpy(d, w->dec_w_ih + (size_t)row * ih_d, DC_RNN * sizeof(float));                          // h_att
        memcpy(d + DC_RNN, w->dec_w_hh + (size_t)row * DC_RNN, DC_RNN * sizeof(float));               // h_dec
        memcpy(d + 2 * DC_RNN, w->dec_w_ih + (size_t)row * ih_d + DC_RNN, DC_ENC * sizeof(float));    // ctx
        bd[row] = w->dec_b_ih[row] + w->dec_b_hh[row];
    }
    for (int a = 0; a < DC_ATT; a++)          // location_dense . location_conv fused into one [128][2][31] filter
        for (int c = 0; c < 2; c++)
            for (int k = 0; k < DC_LOCK; k++) {
                double s = 0.0;
                for (int f = 0; f < DC_LOCF; f++)
                    s += (double)w->loc_dense[(size_t)a * DC_LOCF + f] * (double)w->loc_conv[((size_t)f * 2 + c) * DC_LOCK + k];
                Weff[(size_t)a * 2 * DC_LOCK + c * DC_LOCK + k] = (float)s;
            }
    for (int r = 0; r <= DC_MEL; r++) {       // [ctx | h_dec]
        const float* src = r < DC_MEL ? w->proj_w + (size_t)r * DC_ZP : w->gate_w;
        memcpy(&Wp[(size_t)r * DC_ZP], src + DC_RNN, DC_ENC * sizeof(float));
        memcpy(&Wp[(size_t)r * DC_ZP + DC_ENC], src, DC_RNN * sizeof(float));
        bp[r] = r < DC_MEL ? w->proj_b[r] : w->gate_b[0];
    }
    std::vector<float> p2(w->prenet2, w->prenet2 + (size_t)DC_PRE * DC_PRE), Wq(w->query, w->query + (size_t)DC_ATT * DC_RNN),
        v(w->v, w->v + DC_ATT);
    cudaError_t e = upload(&h->p1T, p1T);
    if (e == cudaSuccess) e = upload(&h->p2, p2);
    if (e == cudaSuccess) e = upload(&h->Wa, Wa);
    if (e == cudaSuccess) e = upload(&h->ba, ba);
    if (e == cudaSuccess) e = upload(&h->Wq, Wq);
    if (e == cudaSuccess) e = upload(&h->v, v);
    if (e == cudaSuccess) e = upload(&h->Weff, Weff);
    if (e == cudaSuccess) e = upload(&h->Wd, Wd);
    if (e == cudaSuccess) e = upload(&h->bd, bd);
    if (e == cudaSuccess) e = upload(&h->Wp, Wp);
    if (e == cudaSuccess) e = upload(&h->bp, bp);
    const size_t nb = DC_MAX_NB, ms = (size_t)h->max_steps;
    const size_t state_cells = nb * (DC_PRE + 4 * DC_RNN + DC_ENC + DC_ATT + DC_MAX_TENC + DC_MEL + 1);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->state, state_cells * sizeof(dc_cell));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->mel_out, nb * ms * DC_MEL * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->mel_T, nb * ms * DC_MEL * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->gate_out, nb * ms * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->t_len, nb * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->n_frames, nb * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&h->err, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreate(&h->ev[i]);
    if (e == cudaSuccess) e = dec_prepare(&h->grid);
    if (e != cudaSuccess) {
        xdtts_decoder_destroy(h);
        if (e == cudaErrorInvalidConfiguration) return fail(XDTTS_ERR_UNSUPPORTED, "decoder_create: the device cannot keep one decoder CTA per SM resident on >= 147 SMs");
        return fail(e == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "decoder_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return XDTTS_OK;
}

extern "C" int xdtts_decoder_max_steps(const xdtts_decoder* h) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "decoder_max_steps: handle is null");
    return h->max_steps;
}

extern "C" int xdtts_decoder_last_timing(const xdtts_decoder* h, float* ms, int* steps) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "decoder_last_timing: handle is null");
    std::lock_guard<std::mutex> lk(const_cast<xdtts_decoder*>(h)->mu);
    if (ms) *ms = h->last_ms;
    if (steps) *steps = h->last_steps;
    return XDTTS_OK;
}

extern "C" int xdtts_decoder_info(const xdtts_decoder* h, int nb, int t_enc, long long* info) {
    if (!h || !info) return fail(XDTTS_ERR_BAD_ARG, "decoder_info: null argument");
    if (nb < 1 || nb > DC_MAX_NB || t_enc < 1 || t_enc > DC_MAX_TENC) return fail(XDTTS_ERR_SHAPE, "decoder_info: nb = %d, t_enc = %d out of range", nb, t_enc);
    dec_info(nb, t_enc, h->grid, info);
    return XDTTS_OK;
}

static int ensure_stage(xdtts_decoder* h, size_t floats) {
    if (h->h_stage_floats >= floats) return XDTTS_OK;
    if (h->h_stage) cudaFreeHost(h->h_stage);
    h->h_stage = nullptr; h->h_stage_floats = 0;
    CU(cudaHostAlloc((void**)&h->h_stage, floats * 4, cudaHostAllocDefault));
    h->h_stage_floats = floats;
    return XDTTS_OK;
}

extern "C" int xdtts_decoder_infer_batch(xdtts_decoder* h, const float* const* memory, const float* const* processed_memory,
                                         int t_enc, const int* unpadded_len, int B, float* const* out_mels, int* n_frames,
                                         float* const* out_gates, float* const* out_align) {
    if (!h) return fail(XDTTS_ERR_BAD_ARG, "decoder_infer: handle is null");
    if (!memory || !processed_memory || !unpadded_len || !out_mels || !n_frames) return fail(XDTTS_ERR_BAD_ARG, "decoder_infer: null argument");
    if (B < 1) return fail(XDTTS_ERR_BAD_ARG, "decoder_infer: B = %d", B);
    if (t_enc < 1 || t_enc > DC_MAX_TENC) return fail(XDTTS_ERR_SHAPE, "decoder_infer: t_enc = %d, supported: 1..%d", t_enc, DC_MAX_TENC);
    for (int b = 0; b < B; b++) {
        if (!memory[b] || !processed_memory[b] || !out_mels[b] || (out_gates && !out_gates[b]) || (out_align && !out_align[b]))
            return fail(XDTTS_ERR_BAD_ARG, "decoder_infer: null buffer for utterance %d", b);
        if (unpadded_len[b] < 1 || unpadded_len[b] > t_enc)
            return fail(XDTTS_ERR_SHAPE, "decoder_infer: unpadded_len[%d] = %d not in 1..t_enc = %d", b, unpadded_len[b], t_enc);
    }
    std::lock_guard<std::mutex> lk(h->mu);
    CU(cudaSetDevice(h->device));
    clear_stale_error(__func__);
    const size_t nbmax = DC_MAX_NB, ms = (size_t)h->max_steps;
    if (h->ws_t_enc < t_enc) {
        cudaFree(h->memory); cudaFree(h->pm);
        h->memory = h->pm = nullptr; h->ws_t_enc = 0;
        CU(cudaMalloc((void**)&h->memory, nbmax * t_enc * DC_ENC * 4));
        CU(cudaMalloc((void**)&h->pm, nbmax * t_enc * DC_ATT * 4));
        h->ws_t_enc = t_enc;
        cudaFree(h->align_out);
        h->align_out = nullptr;
    }
    if (out_align && !h->align_out) CU(cudaMalloc((void**)&h->align_out, nbmax * ms * (size_t)h->ws_t_enc * 4));
    const size_t in_floats = nbmax * (size_t)t_enc * (DC_ENC + DC_ATT);
    const size_t out_floats = nbmax * ms * (DC_MEL + 1 + (out_align ? (size_t)t_enc : 0)) + 16;
    int rc = ensure_stage(h, in_floats > out_floats ? in_floats : out_floats);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    h->last_ms = 0.f;
    h->last_steps = 0;
    for (int b0 = 0; b0 < B; b0 += DC_MAX_NB) {
        const int nb = B - b0 < DC_MAX_NB ? B - b0 : DC_MAX_NB;
        // ---- inputs: one pinned staging copy, one DMA per tensor
        float* sm = h->h_stage;
        float* sp = sm + (size_t)nb * t_enc * DC_ENC;
        for (int b = 0; b < nb; b++) {
            memcpy(sm + (size_t)b * t_enc * DC_ENC, memory[b0 + b], (size_t)t_enc * DC_ENC * 4);
            memcpy(sp + (size_t)b * t_enc * DC_ATT, processed_memory[b0 + b], (size_t)t_enc * DC_ATT * 4);
        }
        CU(cudaMemcpyAsync(h->memory, sm, (size_t)nb * t_enc * DC_ENC * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->pm, sp, (size_t)nb * t_enc * DC_ATT * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->t_len, unpadded_len + b0, nb * sizeof(int), cudaMemcpyHostToDevice, s));
        const size_t state_cells = nbmax * (DC_PRE + 4 * DC_RNN + DC_ENC + DC_ATT + DC_MAX_TENC + DC_MEL + 1);
        CU(cudaMemsetAsync(h->state, 0, state_cells * sizeof(dc_cell), s));   // tag 0: nothing published yet
        CU(cudaMemsetAsync(h->err, 0, 2 * sizeof(int), s));
        CU(cudaMemsetAsync(h->n_frames, 0, nbmax * sizeof(int), s));
        DecParams p;
        memset(&p, 0, sizeof(p));
        p.p1T = h->p1T; p.p2 = h->p2; p.Wa = h->Wa; p.ba = h->ba; p.Wq = h->Wq; p.v = h->v; p.Weff = h->Weff;
        p.Wd = h->Wd; p.bd = h->bd; p.Wp = h->Wp; p.bp = h->bp;
        p.nb = nb; p.t_enc = t_enc; p.memory = h->memory; p.pm = h->pm; p.t_len = h->t_len;
        dc_cell* st = h->state;
        p.x2 = st; st += (size_t)nb * DC_PRE;
        p.h_a = st; st += (size_t)2 * nb * DC_RNN;
        p.h_d = st; st += (size_t)2 * nb * DC_RNN;
        p.ctx = st; st += (size_t)nb * DC_ENC;
        p.pq = st; st += (size_t)nb * DC_ATT;
        p.melt = st; st += (size_t)nb * (DC_MEL + 1);
        p.e = st;
        p.mel_out = h->mel_out; p.gate_out = h->gate_out; p.align_out = out_align ? h->align_out : nullptr;
        p.n_frames = h->n_frames; p.err = h->err; p.barrier = reinterpret_cast<unsigned*>(h->err + 1);
        p.max_steps = h->max_steps; p.gate_threshold = h->gate_threshold; p.seed = h->seed; p.utt_base = b0; p.dropout = h->dropout;
        CU(cudaEventRecord(h->ev[0], s));
        CU(dec_launch(p, h->grid, s));
        g_launches++;
        CU(cudaEventRecord(h->ev[1], s));
        CU(dec_launch_transpose(h->mel_out, h->n_frames, nb, h->max_steps, h->mel_T, s));
        g_launches++;
        // ---- results
        int nf[DC_MAX_NB], poll_err = 0;
        CU(cudaMemcpyAsync(nf, h->n_frames, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&poll_err, h->err, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (poll_err) return fail(XDTTS_ERR_CUDA, "decoder_infer: a stage waited for another CTA's result beyond its bound (internal error)");
        float ms_chunk = 0.f;
        CU(cudaEventElapsedTime(&ms_chunk, h->ev[0], h->ev[1]));
        h->last_ms += ms_chunk;
        float* so = h->h_stage;
        size_t off = 0;
        for (int b = 0; b < nb; b++) {
            n_frames[b0 + b] = nf[b];
            h->last_steps = nf[b] > h->last_steps ? nf[b] : h->last_steps;
            CU(cudaMemcpyAsync(so + off, h->mel_T + (size_t)b * DC_MEL * ms, (size_t)DC_MEL * nf[b] * 4, cudaMemcpyDeviceToHost, s));
            off += (size_t)DC_MEL * nf[b];
            if (out_gates) {
                CU(cudaMemcpyAsync(so + off, h->gate_out + (size_t)b * ms, (size_t)nf[b] * 4, cudaMemcpyDeviceToHost, s));
                off += nf[b];
            }
            if (out_align) {
                CU(cudaMemcpyAsync(so + off, h->align_out + (size_t)b * ms * t_enc, (size_t)nf[b] * t_enc * 4, cudaMemcpyDeviceToHost, s));
                off += (size_t)nf[b] * t_enc;
            }
        }
        CU(cudaStreamSynchronize(s));
        off = 0;
        for (int b = 0; b < nb; b++) {
            memcpy(out_mels[b0 + b], so + off, (size_t)DC_MEL * nf[b] * 4);
            off += (size_t)DC_MEL * nf[b];
            if (out_gates) {
                memcpy(out_gates[b0 + b], so + off, (size_t)nf[b] * 4);
                off += nf[b];
            }
            if (out_align) {
                memcpy(out_align[b0 + b], so + off, (size_t)nf[b] * t_enc * 4);
                off += (size_t)nf[b] * t_enc;
            }
        }
    }
    return XDTTS_OK;
}
