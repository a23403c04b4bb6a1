// Internals shared by the translation units that implement the C ABI (api.cu, postnet_api.cu):
// error plumbing, the Griffin-Lim handle and plan layouts.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xdtts_b200.h"
#include "gl_core.cuh"

namespace xdtts {
extern std::atomic<unsigned long long> g_launches;
// records the thread-local message returned by xdtts_last_error() and returns `code`
int set_error(int code, const char* fmt, ...);
bool debug_on();
void clear_stale_error(const char* where);
bool is_pinned(const void* p);
}  // namespace xdtts

#define CU(expr)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (xdtts::debug_on()) {                                                                      \
            cudaError_t pe_ = cudaPeekAtLastError();                                                  \
            if (pe_ != cudaSuccess || e_ != cudaSuccess)                                              \
                fprintf(stderr, "[xdtts] %s:%d %s -> %s (last error: %s)\n", __FILE__, __LINE__, #expr, \
                        cudaGetErrorName(e_), cudaGetErrorName(pe_));                                 \
        }                                                                                             \
        if (e_ != cudaSuccess)                                                                        \
            return xdtts::set_error(e_ == cudaErrorMemoryAllocation ? XDTTS_ERR_OOM : XDTTS_ERR_CUDA, "%s: %s", #expr, \
                                    cudaGetErrorString(e_));                                          \
    } while (0)

// ------------------------------------------------------------------ Griffin-Lim handle and plan
struct xdtts_gl {
    int device = 0, n_mels = 0, K = 0, n_fft = 0, hop = 0, n_iter = 0, sm_count = 148;
    float power = 1.f, momentum = 0.f;
    xdtts_gl_opts opts{};
    std::vector<float> pinv;        // [K][n_mels] host copy
    float* d_pinvT = nullptr;       // [n_mels][K]
    float* d_lift_img = nullptr;    // fp16 hi/lo tiles of the scaled pseudo-inverse in the MMA's operand layout (gl_lift.cu)
    float2* d_gtw = nullptr;        // un-fused path (gl_generic.cu): FFT twiddles and the Hann window
    float* d_gwin = nullptr;
    bool generic = false;           // geometry outside the fused kernel's: any hop, any power-of-two n_fft
    float lift_p_exp = 0.f;         // ... which holds pinv * 2^lift_p_exp
    float2* d_tables = nullptr;
    float* d_edge = nullptr;
    int *d_csr = nullptr, *d_csc = nullptr;        // sparse forms of the mel basis for the NNLS lift (rows / columns)
    float *d_csr_val = nullptr, *d_csc_val = nullptr;
    int band_rw = 0, band_cw = 0;                   // > 0: the basis is banded (a filterbank): fast form of the NNLS lift
    int* d_band_lo = nullptr;
    float* d_bandT = nullptr;
    int* d_ell_row = nullptr;
    float* d_ell_val = nullptr;
    float lipschitz = 0.f;                          // sigma_max(basis)^2
    cudaStream_t stream = nullptr;
    std::atomic<unsigned long long> calls{0};   // seeded-phase calls made on this handle (advances the phase seed)
    std::mutex mu;
    std::vector<xdtts_gl_plan*> cache;   // plans owned by the batch entry points
};

struct xdtts_gl_plan {
    xdtts_gl* h = nullptr;
    int B = 0, max_T = 0, total_T = 0, run_frames = 0;
    std::vector<int> Ts, foff;
    std::vector<long long> out_off;
    long long out_total = 0;
    std::vector<xdtts::GlRun> runs;
    // device
    xdtts::GlRun* d_runs = nullptr;
    int *d_T = nullptr, *d_foff = nullptr;
    long long* d_out_off = nullptr;
    float *d_mel = nullptr, *d_in_mag = nullptr, *d_in_phase = nullptr, *d_turns = nullptr;
    float* d_state = nullptr;        // per-frame records [R | S | S_nyq | pad], rec_f floats each (gl_core.cuh Geo::REC_F)
    int rec_f = 0;
    float *d_y[2] = {nullptr, nullptr}, *d_halo = nullptr, *d_out = nullptr;
    std::vector<int4> lift_tiles;    // (frame row of the utterance, its T, first frame, 0) of every frame tile of the lift
    int4* d_lift_tiles = nullptr;
    int lift_tile_frames = 64;
    float* d_frames = nullptr;       // un-fused path: [total frames][n_fft] windowed inverse transforms
    bool generic = false;            // this plan runs the un-fused kernels (the handle's geometry, or an utterance of 2-3 frames)
    unsigned char *d_seed = nullptr, *h_seed = nullptr;   // [u64 phase seed][int stream index per utterance], device + pinned
    short* d_pcm = nullptr;          // 16-bit PCM copy of d_out (allocated on first use)
    short* h_pcm = nullptr;
    unsigned *d_flags = nullptr, *d_amax = nullptr, *d_done = nullptr;
    bool use_persistent = false;     // the run table fits the resident warps: one cooperative launch per vocode
    // pinned staging for pageable callers
    float *h_in = nullptr, *h_out = nullptr;
    size_t h_in_floats = 0;
    cudaEvent_t h_in_busy = nullptr;         // recorded after the DMA that reads h_in
    std::vector<cudaEvent_t> out_ev;         // per utterance: its device -> h_out copy is done
    cudaGraphExec_t graphs[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float lift_ms = -1.f;            // device time of the mel -> linear step of the last kernel-by-kernel pass
};


namespace xdtts {
// the *_locked functions expect h->mu to be held
int gl_plan_build(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out);
int gl_plan_upload_locked(xdtts_gl_plan* p, int kind, const float* const* srcs, cudaStream_t s);   // s == null: the handle's stream
// seed / seed_ids: null = the handle's next seed / utterance b draws stream b (a pool passes one seed and global indices)
int gl_plan_launch_async(xdtts_gl_plan* p, int flags, cudaStream_t s, const unsigned long long* seed = nullptr, const int* seed_ids = nullptr);
unsigned long long gl_draw_seed(xdtts_gl* h);
int gl_plan_set_seed(xdtts_gl_plan* p, unsigned long long seed, const int* seed_ids, cudaStream_t s);
int gl_batch_common(xdtts_gl* h, int kind, const float* const* ins, const int* Ts, int B, const float* const* phases,
                    float* const* outs, short* const* pcm_outs, const unsigned long long* seed, const int* seed_ids);
int gl_plan_download_async(xdtts_gl_plan* p, float* const* outs, cudaStream_t s, bool* staged);
void gl_plan_download_finish(xdtts_gl_plan* p, float* const* outs);
int gl_plan_run_locked(xdtts_gl_plan* p, int flags, float* ms_total, float* ms_iter, int* n_iter_launches,
                       const unsigned long long* seed = nullptr, const int* seed_ids = nullptr);
int gl_plan_download_locked(xdtts_gl_plan* p, float* const* outs);
int gl_cached_plan(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out);
int gl_plan_mel_arena(xdtts_gl_plan* p, float** out);
}  // namespace xdtts
