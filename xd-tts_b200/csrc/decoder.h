// Internal declarations of the Tacotron2 decoder-loop kernel (decoder.cu) shared with the C ABI (decoder_api.cu).
#pragma once
#include <cuda_runtime.h>

namespace xdtts {

// NVIDIA Tacotron2 decoder dimensions (DecoderState::new, /root/reference src/tacotron2/mod.rs:207-210:
// attention_rnn_dim 1024, decoder_rnn_dim 1024, encoder_embedding_dim 512, n_mel_channels 80; prenet 256,
// attention 128, location filters 32 x kernel 31 are the model's published hyper-parameters)
constexpr int DC_MEL = 80, DC_PRE = 256, DC_ENC = 512, DC_RNN = 1024, DC_ATT = 128, DC_LOCF = 32, DC_LOCK = 31;
constexpr int DC_ZA = DC_ENC + DC_RNN + DC_PRE;    // attention-LSTM input  [ctx | h_att | prenet]  (1792)
constexpr int DC_ZD = DC_RNN + DC_RNN + DC_ENC;    // decoder-LSTM input    [h_att | h_dec | ctx]   (2560)
constexpr int DC_ZP = DC_RNN + DC_ENC;             // projection input      [h_dec | ctx]           (1536)
constexpr int DC_THREADS = 512;
constexpr int DC_UNITS = 7;                        // hidden units per CTA: 147 CTAs x 7 >= 1024
constexpr int DC_MAX_TENC = 512;
constexpr int DC_MAX_NB = 8;                       // utterances decoded in lockstep by one launch
constexpr int DC_WEFF_LD = 2 * DC_LOCK + 1;        // padded row of the fused location filter (bank-conflict free)

typedef unsigned long long dc_cell;   // {value, tag}: a float and the step that published it (decoder.cu)

struct DecParams {
    // weights, re-laid out on the host (decoder_api.cu)
    const float* p1T;    // [80][256]     prenet layer 1, transposed
    const float* p2;     // [256][256]    prenet layer 2
    const float* Wa;     // [4096][1792]  attention LSTM [W_ih | W_hh] with columns ordered [ctx | h_att | prenet]
    const float* ba;     // [4096]        b_ih + b_hh
    const float* Wq;     // [128][1024]   query layer
    const float* v;      // [128]
    const float* Weff;   // [128][63]     location dense x location conv, fused: [a][c * 31 + k]
    const float* Wd;     // [4096][2560]  decoder LSTM with columns ordered [h_att | h_dec | ctx]
    const float* bd;     // [4096]
    const float* Wp;     // [81][1536]    rows 0..79 linear projection, row 80 gate layer; columns [ctx | h_dec]
    const float* bp;     // [81]
    // inputs of this launch
    int nb;                 // utterances in this launch (<= DC_MAX_NB)
    int t_enc;              // padded encoder length (common to the batch)
    const float* memory;    // [nb][t_enc][512]
    const float* pm;        // [nb][t_enc][128]  processed_memory
    const int* t_len;       // [nb] unpadded lengths (mask: t >= t_len)
    // cross-CTA state in global memory: cells {value, tag = step + 1}, all zero at launch
    dc_cell* x2;            // [nb][256]           prenet output of the current step
    dc_cell* h_a;           // [2][nb][1024]       attention LSTM hidden state, double buffered by step parity
    dc_cell* h_d;           // [2][nb][1024]
    dc_cell* ctx;           // [nb][512]           attention context of the current step
    dc_cell* pq;            // [nb][128]           processed query
    dc_cell* e;             // [nb][t_enc]         attention energies
    dc_cell* melt;          // [nb][81]            the step's mel frame and gate logit (the in-loop copy of mel_out / gate_out)
    int* err;               // set to 1 if a poll ran into its bound (never, unless the kernel is wrong)
    unsigned* barrier;      // grid barrier counter of the batch >= 4 form, zero at launch
    float* mel_out;         // [nb][max_steps][80] decoder outputs, frame major (also the next step's input)
    float* gate_out;        // [nb][max_steps]     gate logits
    float* align_out;       // [nb][max_steps][t_enc] attention weights, or null
    int* n_frames;          // [nb] frames produced (the frame whose gate fired is kept)
    int max_steps;
    float gate_threshold;   // stop when sigmoid(gate) > threshold
    unsigned long long seed;
    int utt_base;           // global index of utterance 0 of this launch (dropout stream)
    int dropout;            // 1: prenet dropout on (the exported graph's behaviour), 0: off
    unsigned cache_mask;    // set by dec_launch: LSTM weight slices kept in shared memory (decoder.cu dec_slice)
};

cudaError_t dec_prepare(int* grid_out);
// one cooperative launch: the whole decoder loop for p.nb utterances
cudaError_t dec_launch(const DecParams& p, int grid, cudaStream_t s);
void dec_info(int nb, int t_enc, int grid, long long* info);
// [nb][max_steps][80] frame-major decoder output -> per utterance [80][n_frames[b]] row-major at dst + b * 80 * max_steps
cudaError_t dec_launch_transpose(const float* mel_frames, const int* n_frames, int nb, int max_steps, float* dst, cudaStream_t s);

}  // namespace xdtts
