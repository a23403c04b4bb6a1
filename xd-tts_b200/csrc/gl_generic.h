// Host-side declarations of the un-fused Griffin-Lim path (gl_generic.cu): any hop, any power-of-two n_fft in [64, 4096].
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace xdtts {

struct GlgParams {
    int n_fft, hop, n_utt;
    const int* utt_T;        // [n_utt] frames per utterance
    const int* utt_foff;     // [n_utt] first frame row of each utterance
    float* state;            // per-frame records [R: M float2, slot 0 = (Re R[0], Re R[M]) | S: M floats | S_nyq | pad], rec_f floats apart
    int rec_f;
    float* frames;           // [total frames][n_fft] windowed inverse transforms of the current iteration
    float* y;                // waveforms: utterance u at foff[u] * hop, hop * (T_u - 1) samples
    const float* turns;      // caller-supplied initial phase in turns, [frame][K], or null (counter-based generator)
    const unsigned long long* seed;
    const int* utt_seed_id;
    const float2* tw;        // exp(-2 pi i j / n_fft), j < n_fft / 2
    const float* win;        // periodic Hann
    unsigned* amax;          // [n_utt] bit pattern of the peak |y| (written by the last overlap-add)
    float alpha;             // momentum / (1 + momentum)
    int pad_mode;
};

std::vector<float2> glg_build_twiddles(int n_fft);
std::vector<float> glg_build_window(int n_fft);
cudaError_t glg_prepare(int n_fft);
cudaError_t glg_launch_frames(const GlgParams& p, int mode, int total_frames, cudaStream_t s);   // mode 0: initial spectrum, 1: iteration
cudaError_t glg_launch_ola(const GlgParams& p, bool last, int max_T, cudaStream_t s);
cudaError_t glg_launch_finish(const float* y, const int* utt_T, const int* utt_foff, const long long* out_off, const unsigned* amax,
                              int n_utt, int max_T, int hop, int normalise, float* out, cudaStream_t s);

}  // namespace xdtts
