// Internal declarations of the postnet kernels (postnet.cu) shared with the C ABI (postnet_api.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace xdtts {

constexpr int PN_BM = 128;        // frame rows per tile = UMMA M = TMEM lanes
constexpr int PN_BK = 64;         // K elements per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int PN_MAX_BN = 256;    // output channels per tile = UMMA N = fp32 TMEM columns per accumulator stage
constexpr int PN_STAGES = 4;
constexpr int PN_THREADS = 256;
constexpr int PN_MAX_COUT = 1024;
constexpr int PN_MAX_CIN0 = 128;  // channels of the network input (mel bins)
constexpr int PN_MAX_LAYERS = 8;

// one convolution layer over one plan (batch)
struct PnLayer {
    int m_tiles, n_tiles, block_n;   // tile grid; block_n: channels per tile (multiple of 16, <= 256)
    int cin_chunks;                  // padded input channels / 64
    int cout;                        // output channels
    int n_split;                     // 1: bf16, 3: bf16x3
    int taps;                        // kernel size (5)
    int apply_tanh;
    int last;                        // 1: fp32 [C][T] output + residual instead of a hidden activation
    const float* bias;               // [cout], BatchNorm folded in
    __nv_bfloat16* out_hi;           // hidden output [rows + 4][out_ld]
    __nv_bfloat16* out_lo;           // null when n_split == 1
    int out_ld;
    const int* row_t;                // [m_tiles * 128] frame index of a tile row inside its utterance, -1 for gap rows
    const int* row_u;                // [m_tiles * 128] utterance of a tile row
    const int* utt_T;
    const int* utt_foff;             // first frame of the utterance in the caller-layout arenas (x C floats)
    const float* resid;              // last layer: the network input, caller layout
    float* out_f32;                  // last layer: output, caller layout
};

cudaError_t pn_prepare();
// maps4: A hi, A lo, B hi, B lo (lo = copies of hi when n_split == 1)
cudaError_t pn_launch_conv_tc(const CUtensorMap* maps4, const PnLayer& p, int sm_count, cudaStream_t s);
cudaError_t pn_launch_stage_input(const float* mel, const int* utt_T, const int* utt_foff, const int* utt_roff, int n_utt,
                                  int max_T, int C, int ld, __nv_bfloat16* hi, __nv_bfloat16* lo, float* f32, cudaStream_t s);
// x: [rows + 4][ld_in] fp32 time-major, w: [taps][cin][cout] fp32
cudaError_t pn_launch_conv_f32(const float* x, int ld_in, const float* w, const PnLayer& p, int cin, float* out, int ld_out,
                               cudaStream_t s);

}  // namespace xdtts
