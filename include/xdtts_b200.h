/* xdtts_b200.h -- C ABI of the B200 (sm_100a) vocoding + postnet back end for xd-tts.
 *
 * Drop-in boundary for the two calls XdTts::infer makes on its hot path
 * (/root/reference src/lib.rs:110-159): `self.model.infer(input)` -> postnet tail
 * (src/tacotron2/mod.rs:344-357) and `self.vocoder.infer(&spectrogram)` (src/lib.rs:141,
 * griffin_lim::GriffinLim, external crate griffin-lim 0.2.0 @ e6415314, Cargo.lock:666-680).
 * A Rust shim crate named `griffin-lim` binds these symbols (INTEGRATION.md shows it); the
 * Python mirror in xd-tts_b200/xdtts_b200/ binds them through ctypes.
 *
 * Conventions: plain pointers and sizes only.  Host arrays are C-contiguous float32 in the
 * reference's ndarray layouts: mel [n_mels, T], linear magnitude / phase [K, T] with
 * K = n_fft/2 + 1, waveform [hop * (T - 1)].  The caller owns every host buffer, inputs and
 * outputs; the library returns only opaque handles.  Every function returns 0 or a negative
 * XDTTS_ERR_* code and never throws or aborts; xdtts_last_error() gives the message of the
 * last failure on the calling thread (maps to anyhow::bail! in the shim).  Handles may be
 * shared between threads (calls on one handle are serialised internally).
 * There is NO CPU fallback: every entry point fails with XDTTS_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef XDTTS_B200_H
#define XDTTS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define XDTTS_OK 0
#define XDTTS_ERR_BAD_ARG (-1)     /* null pointer, non-finite or out-of-range parameter */
#define XDTTS_ERR_SHAPE (-2)       /* T < 2, K != n_fft/2+1, wrong out_len ... */
#define XDTTS_ERR_CUDA (-3)        /* CUDA runtime error, no usable sm_100 device */
#define XDTTS_ERR_OOM (-4)
#define XDTTS_ERR_UNSUPPORTED (-5) /* n_fft not a power of two in [64, 4096], n_mels > 256, a graph that is not the expected network */

/* Options whose value the un-vendored crate fixes internally (SURVEY.md section 7 "unknowns");
 * defaults (all zero) are the librosa-0.9.2 behaviour the crate ports -- WITH ONE DELIBERATE EXCEPTION:
 *
 *   lift = 0, the default, is the mel-filterbank PSEUDO-INVERSE lift this back end was specified with.  The crate behind
 *   GriffinLim::infer most likely does what librosa's mel_to_stft does (non-negative least squares: least-squares start +
 *   L-BFGS-B; its dependencies lbfgsb and ndarray-linalg are in Cargo.lock:888,1006).  The two give different magnitudes
 *   (about 2% apart on speech-like mels, up to 70% on random ones), i.e. with the defaults this library does NOT reproduce
 *   the reference's audio sample for sample.  Set lift = 1 for the librosa-faithful lift (costs +0.2 ms per 32 x 1000
 *   frames on speech-like input); the Rust shim exposes it as the cargo feature `nnls-lift` / XDTTS_B200_LIFT=nnls.
 *   Likewise `exponent` selects between S = x^power (default, the call site's reading) and librosa's x^(1/power).
 *   Which combination the crate really uses is what tools/capture_reference_fixtures.sh + test_reference_fixtures decide. */
typedef struct xdtts_gl_opts {
    int delog;               /* 0: exp (Tacotron2 ln-mel, default)  1: 10^x  2: mel is already linear */
    int pad_mode;            /* 0: reflect (librosa 0.9.2 stft)     1: constant zero */
    int normalise;           /* 0: peak-normalise to [-1,1] (caller scales by i16::MAX, src/lib.rs:155)  1: none */
    int run_frames;          /* 0: auto.  Frames per warp run (tuning knob, >= 4) */
    unsigned long long seed; /* seed of the random initial phase used when the caller passes none */
    int persistent;          /* 0: one launch per iteration, CUDA graph (default; fastest at full occupancy)
                                1: whole vocode in one cooperative launch when the batch fits one resident wave */
    int lift;                /* mel -> linear magnitude.  0: S = max(0, pinv(basis) . delog(mel)) ^ power (default)
                                1: non-negative least squares as librosa's mel_to_stft does it (util.nnls: start from the
                                   clipped least-squares solution, minimise |basis . x - delog(mel)|^2 over x >= 0), solved
                                   per frame by accelerated projected gradient on the device; S = x ^ power */
    int nnls_iters;          /* lift = 1: iteration cap per frame (0: 300); a frame stops early once its projected gradient
                                falls below 3e-6 (librosa passes pgtol = 1e-5 to L-BFGS-B; 3e-6 here reaches that
                                solver's objective, tests/test_oracle.py) */
    int fixed_seed;          /* when the caller passes no initial phase: 0 (default): every call on the handle draws a NEW phase
                                field, as the reference does (call c uses seed + c * 0xD1B54A32D192ED03; call 0 uses `seed`)
                                1: every call draws the field of `seed` (reproducible runs, tests) */
    int exponent;            /* 0: S = x ^ power (default: the Tacotron convention the call site describes, "tuned by ear ...
                                1.2-1.7", src/tacotron2/mod.rs:446-449)   1: S = x ^ (1 / power), librosa's mel_to_stft */
} xdtts_gl_opts;

typedef struct xdtts_gl xdtts_gl;           /* replaces griffin_lim::GriffinLim */
typedef struct xdtts_gl_plan xdtts_gl_plan; /* device-resident batch: buffers + captured CUDA graph */

/* griffin_lim::mel::create_mel_filter_bank(sr, n_fft, n_mels, fmin, fmax) (call site
 * src/tacotron2/mod.rs:453): Slaney-scale, Slaney-normalised triangular filters (librosa
 * filters.mel, htk=False, norm='slaney').  fmax < 0 means None (sr/2).  out: [n_mels, n_fft/2+1]. */
int xdtts_mel_filter_bank(float sr, int n_fft, int n_mels, float fmin, float fmax, float* out);

/* Moore-Penrose pseudo-inverse of a [rows, cols] row-major matrix -> out [cols, rows] (host, fp64
 * one-sided Jacobi; stands in for the MKL lstsq/SVD the crate uses, Cargo.lock:1006).  This is
 * the matrix the mel -> linear lift multiplies by. */
int xdtts_pinv(const float* a, int rows, int cols, float* out);

/* GriffinLim::new(mel_basis, noverlap, power, iter, momentum) (call site src/tacotron2/mod.rs:456).
 * mel_basis: [n_mels, K] row-major; n_fft = 2 (K-1), a power of two in [64, 4096]; hop = n_fft - noverlap, any value in
 * [1, n_fft].  hop == n_fft / 4 at n_fft 512 / 1024 / 2048 (the shipped call: 1024 / 768) runs the fused kernel -- one
 * launch per iteration -- every other geometry the un-fused kernels (two launches per iteration, same results contract).
 * Builds the pseudo-inverse of the basis once (host, fp64) and uploads constants to `device`. */
int xdtts_gl_create(const float* mel_basis, int n_mels, int K, int noverlap, float power, int n_iter, float momentum,
                    const xdtts_gl_opts* opts_or_null, int device, xdtts_gl** out);
void xdtts_gl_destroy(xdtts_gl* h);

/* hop * (T - 1), or a negative error */
int xdtts_gl_out_len(const xdtts_gl* h, int T);
/* copies the [K, n_mels] pseudo-inverse the lift uses (for parity checks) */
int xdtts_gl_get_pinv(const xdtts_gl* h, float* out);

/* GriffinLim::infer(&mel) (call site src/lib.rs:141): mel [n_mels, T] -> out [hop (T-1)].
 * init_phase_or_null: [K, T] initial phase in turns, u in [0,1) (angle = 2 pi u); null draws
 * u from the counter-based generator (splitmix64 of seed, utterance index, t*K + k). */
int xdtts_gl_infer(xdtts_gl* h, const float* mel, int T, const float* init_phase_or_null, float* out, int out_len);
/* B independent utterances in one pass (the reference loops; src/lib.rs:83-104).  init_phases
 * may be null, or an array of B pointers (all non-null). */
int xdtts_gl_infer_batch(xdtts_gl* h, const float* const* mels, const int* Ts, int B,
                         const float* const* init_phases_or_null, float* const* outs);
/* Same, starting from linear magnitudes [K, T] (skips the mel -> linear lift; rows a6-a8 only) */
int xdtts_gl_from_mag_batch(xdtts_gl* h, const float* const* mags, const int* Ts, int B,
                            const float* const* init_phases_or_null, float* const* outs);

/* As xdtts_gl_infer_batch, but returns 16-bit PCM: the caller's loop
 * `wav.write_sample((sample * i16::MAX as f32) as i16)` (src/lib.rs:153-157: truncate toward zero,
 * saturate, NaN -> 0) runs on the device, which halves the device -> host bytes. outs[b]: hop (T_b - 1) samples. */
int xdtts_gl_infer_batch_pcm16(xdtts_gl* h, const float* const* mels, const int* Ts, int B,
                               const float* const* init_phases_or_null, short* const* outs);

/* ---- device-resident plan: what the batch calls above use internally, exposed so that a
 * pipeline (or the benchmark) can keep inputs and outputs in HBM and time the device work. */
int xdtts_gl_plan_create(xdtts_gl* h, const int* Ts, int B, xdtts_gl_plan** out);
void xdtts_gl_plan_destroy(xdtts_gl_plan* p);
/* kind: 0 mel [n_mels,T], 1 magnitude [K,T], 2 initial phase [K,T]; srcs: B host pointers */
int xdtts_gl_plan_upload(xdtts_gl_plan* p, int kind, const float* const* srcs);
#define XDTTS_RUN_FROM_MAG 1    /* start from the uploaded magnitudes instead of lifting the mels */
#define XDTTS_RUN_USE_PHASE 2   /* use the uploaded initial phase instead of the seeded generator */
#define XDTTS_RUN_NO_GRAPH 4    /* launch kernel by kernel (needed for ms_iter) */
#define XDTTS_RUN_PER_LAUNCH 16 /* one launch per iteration even when the persistent single-launch kernel applies */
/* Runs lift/transposes, Griffin-Lim and the normalisation on the plan's stream and waits: one launch per
 * iteration, captured in a CUDA graph.  With opts.persistent = 1 and a batch that fits the device's resident
 * warps in one wave (xdtts_gl_plan_is_persistent) the initial inverse transform and all n_iter iterations are
 * instead ONE cooperative launch in which runs hand hop blocks to their neighbours through L2 flags
 * (bit-identical results; measured equal or slower on B200); PER_LAUNCH overrides it per call.
 * ms_total: device time of the whole pass (CUDA events).  With NO_GRAPH: ms_iter / n_iter_launches =
 * device time and count of the Griffin-Lim launches it covers (the n_iter-2 steady-state launches, or the
 * single persistent launch). */
int xdtts_gl_plan_run(xdtts_gl_plan* p, int flags, float* ms_total, float* ms_iter, int* n_iter_launches);
/* device time (CUDA events) of the mel -> linear step (lift, or magnitude transpose, + NNLS when enabled) of the last
 * XDTTS_RUN_NO_GRAPH pass of this plan */
int xdtts_gl_plan_lift_ms(const xdtts_gl_plan* p, float* ms);
/* measurement: the same step of this plan launched reps times back to back (after one warm-up), device time per launch --
 * CUDA events around the whole train, so the ~7 us of launch and event latency a single bracketed launch carries is amortised */
int xdtts_gl_plan_time_lift(xdtts_gl_plan* p, int reps, float* ms_per_launch);
int xdtts_gl_plan_download(xdtts_gl_plan* p, float* const* outs);
int xdtts_gl_plan_download_pcm16(xdtts_gl_plan* p, short* const* outs);   /* after a run: 16-bit PCM of the same waveforms */
/* debugging / parity: copy device state to host. what: 0 S [T_total][M] frame-major, 1 S Nyquist [T_total],
 * 2 R [T_total][M][2]; n_floats must match. */
int xdtts_gl_plan_peek(xdtts_gl_plan* p, int what, float* out, long long n_floats);
/* 1 when xdtts_gl_plan_run uses the persistent single-launch kernel for this plan, 0 otherwise */
int xdtts_gl_plan_is_persistent(const xdtts_gl_plan* p);
/* geometry the plan chose: info[0]=n_runs, [1]=frames per run (max), [2]=CTAs, [3]=total frames */
int xdtts_gl_plan_info(const xdtts_gl_plan* p, int* info4);

/* ---- Tacotron2 postnet: replaces the `postnet` ort::Session (src/tacotron2/mod.rs:256-259 load,
 * :344-357 run; output "mel_outputs_postnet").  out = mel + Postnet(mel), Postnet = n_layers x
 * [Conv1d(k=5, pad=2, bias) -> BatchNorm1d(eval)], tanh after every layer but the last. */
typedef struct xdtts_postnet_opts {
    int precision; /* 0: bf16x3 split on the tensor cores (fp32-class, default)  1: single bf16 pass
                      2: fp32 on the CUDA cores (strict parity / cross-check, slow) */
} xdtts_postnet_opts;

typedef struct xdtts_postnet xdtts_postnet;           /* the weights, folded and laid out for the device */
typedef struct xdtts_postnet_plan xdtts_postnet_plan; /* device-resident batch */

/* channels: n_layers+1 entries (Tacotron2: 80,512,512,512,512,80); ksize must be 5.
 * conv_w[l]: [channels[l+1], channels[l], ksize] row-major (the ONNX initializer layout); conv_b[l]:
 * [channels[l+1]] or null.  bn_*[l]: [channels[l+1]] each, or bn_gamma == null / bn_gamma[l] == null for
 * a layer without BatchNorm.  BatchNorm is folded on the host in fp64. */
int xdtts_postnet_create(int n_layers, const int* channels, int ksize, const float* const* conv_w,
                         const float* const* conv_b, const float* const* bn_gamma, const float* const* bn_beta,
                         const float* const* bn_mean, const float* const* bn_var, float eps,
                         const xdtts_postnet_opts* opts_or_null, int device, xdtts_postnet** out);
void xdtts_postnet_destroy(xdtts_postnet* h);
/* mel [C, T] -> out [C, T] (C = channels[0]), host buffers, T >= 1 */
int xdtts_postnet_infer(xdtts_postnet* h, const float* mel, int T, float* out);
int xdtts_postnet_infer_batch(xdtts_postnet* h, const float* const* mels, const int* Ts, int B, float* const* outs);

int xdtts_postnet_plan_create(xdtts_postnet* h, const int* Ts, int B, xdtts_postnet_plan** out);
void xdtts_postnet_plan_destroy(xdtts_postnet_plan* p);
int xdtts_postnet_plan_upload(xdtts_postnet_plan* p, const float* const* mels);
/* feed_or_null: a vocoder plan of the same batch shape; the result is then written into that plan's
 * mel arena (on the vocoder's stream) so that xdtts_gl_plan_run(feed, 0, ...) vocodes it without a
 * host round trip.  ms_total: device time of the postnet (CUDA events). */
int xdtts_postnet_plan_run(xdtts_postnet_plan* p, xdtts_gl_plan* feed_or_null, float* ms_total);
int xdtts_postnet_plan_download(xdtts_postnet_plan* p, float* const* outs);

/* Tacotron2::load for the postnet (src/tacotron2/mod.rs:256-259 opens postnet.onnx as an ort::Session):
 * reads the Conv / BatchNormalization initializers out of the ONNX file (own protobuf wire-format
 * reader, no ONNX Runtime) and builds the device postnet from them.  The graph is matched by
 * structure (Conv nodes in order, the BatchNormalization consuming each Conv), not by tensor names. */
int xdtts_postnet_create_from_onnx(const char* path, const xdtts_postnet_opts* opts_or_null, int device, xdtts_postnet** out);
/* host-only access to the same reader (no GPU needed): layer shapes and tensors of a postnet.onnx */
typedef struct xdtts_onnx_postnet xdtts_onnx_postnet;
int xdtts_onnx_postnet_open(const char* path, xdtts_onnx_postnet** out);
void xdtts_onnx_postnet_close(xdtts_onnx_postnet* m);
int xdtts_onnx_postnet_n_layers(const xdtts_onnx_postnet* m);   /* >= 1, or a negative error */
int xdtts_onnx_postnet_layer_info(const xdtts_onnx_postnet* m, int layer, int* cout, int* cin, int* ksize, int* has_bias,
                                  int* has_bn, float* eps);
/* which: 0 conv weight [cout, cin, k], 1 conv bias, 2 BN scale, 3 BN bias, 4 BN running mean, 5 BN running var */
int xdtts_onnx_postnet_layer_copy(const xdtts_onnx_postnet* m, int layer, int which, float* out);

/* The tail of XdTts::infer in one call (src/lib.rs:123 postnet tail of model.infer + :141 vocoder.infer):
 * decoder mels [C, T_b] -> postnet -> mel-to-linear lift -> Griffin-Lim -> waveforms [hop (T_b - 1)].
 * out_mels_or_null: B buffers [C, T_b] that receive "mel_outputs_postnet" (the --output-spectrogram dump). */
int xdtts_tail_infer_batch(xdtts_postnet* pn, xdtts_gl* gl, const float* const* mels, const int* Ts, int B,
                           const float* const* init_phases_or_null, float* const* out_mels_or_null,
                           float* const* out_waves);

/* ---- Tacotron2 decoder loop: replaces the `decoder` ort::Session and the per-frame loop around it in
 * Tacotron2::run_decoder (src/tacotron2/mod.rs:272-342; session load :251-254; state DecoderState::new
 * :204-238).  The graph is NVIDIA Tacotron2's Decoder.decode as exported by export_tacotron2_onnx.py
 * (named at src/tacotron2/mod.rs:137-138): prenet -> attention LSTM -> location-sensitive attention ->
 * decoder LSTM -> linear projection + gate, fp32.  The whole loop -- every step of up to 8 utterances in
 * lockstep -- is ONE persistent cooperative kernel; nothing returns to the host between frames.
 * Weights are passed in the PyTorch layouts of the checkpoint the reference's ONNX files were exported
 * from (row-major float32; LSTM gate order i, f, g, o). */
typedef struct xdtts_decoder_weights {
    const float* prenet1;   /* [256, 80]    decoder.prenet.layers.0.linear_layer.weight */
    const float* prenet2;   /* [256, 256]   decoder.prenet.layers.1.linear_layer.weight */
    const float* att_w_ih;  /* [4096, 768]  decoder.attention_rnn.weight_ih, input = [prenet(256) | context(512)] */
    const float* att_w_hh;  /* [4096, 1024] */
    const float* att_b_ih;  /* [4096] */
    const float* att_b_hh;  /* [4096] */
    const float* query;     /* [128, 1024]  decoder.attention_layer.query_layer.linear_layer.weight */
    const float* v;         /* [128]        decoder.attention_layer.v.linear_layer.weight */
    const float* loc_conv;  /* [32, 2, 31]  ...location_layer.location_conv.conv.weight */
    const float* loc_dense; /* [128, 32]    ...location_layer.location_dense.linear_layer.weight */
    const float* dec_w_ih;  /* [4096, 1536] decoder.decoder_rnn.weight_ih, input = [attention_hidden(1024) | context(512)] */
    const float* dec_w_hh;  /* [4096, 1024] */
    const float* dec_b_ih;  /* [4096] */
    const float* dec_b_hh;  /* [4096] */
    const float* proj_w;    /* [80, 1536]   decoder.linear_projection, input = [decoder_hidden(1024) | context(512)] */
    const float* proj_b;    /* [80] */
    const float* gate_w;    /* [1536]       decoder.gate_layer */
    const float* gate_b;    /* [1] */
} xdtts_decoder_weights;

typedef struct xdtts_decoder_opts {
    float gate_threshold;    /* 0: 0.6, the reference's constant (src/tacotron2/mod.rs:279); stop when sigmoid(gate) > it */
    int max_steps;           /* 0: 1000 (src/tacotron2/mod.rs:280) */
    int prenet_dropout;      /* 0: on -- the exported prenet draws a Bernoulli(0.5) mask inside the graph at every step, so
                                the reference's decoder is stochastic; here the mask is the counter-based generator
                                splitmix64(seed, utterance, step * 512 + layer * 256 + unit) <= 0.5.  1: off (deterministic) */
    unsigned long long seed; /* seed of the dropout stream */
} xdtts_decoder_opts;

typedef struct xdtts_decoder xdtts_decoder;
int xdtts_decoder_create(const xdtts_decoder_weights* w, const xdtts_decoder_opts* opts_or_null, int device, xdtts_decoder** out);
void xdtts_decoder_destroy(xdtts_decoder* h);
int xdtts_decoder_max_steps(const xdtts_decoder* h);   /* frame capacity the output buffers must have */
/* run_decoder for B utterances (the reference decodes one at a time).  memory[b]: [t_enc, 512],
 * processed_memory[b]: [t_enc, 128] -- the encoder session's outputs without their batch axis, padded to a
 * common t_enc <= 512 (the reference pads every chunk to 100, src/tacotron2/mod.rs:366-368); unpadded_len[b]
 * in 1..t_enc is the mask boundary of DecoderState::new (:228-229).  out_mels[b] must hold 80 * max_steps
 * floats and receives the spectrogram as [80, n_frames[b]] row-major (the transposed layout of :345, what the
 * postnet consumes); the frame whose gate fires is kept (:312-324).  out_gates_or_null[b]: [n_frames[b]] gate
 * logits (capacity max_steps); out_align_or_null[b]: [n_frames[b], t_enc] attention weights
 * (capacity max_steps * t_enc). */
int xdtts_decoder_infer_batch(xdtts_decoder* h, const float* const* memory, const float* const* processed_memory, int t_enc,
                              const int* unpadded_len, int B, float* const* out_mels, int* n_frames,
                              float* const* out_gates_or_null, float* const* out_align_or_null);
/* Tacotron2::load for the decoder (src/tacotron2/mod.rs:251-254 opens decoder_iter.onnx as an ort::Session): reads the
 * weights out of the ONNX file with the library's own protobuf reader and builds the device decoder from them.  Tensors are
 * found by their role in the dataflow between the graph's named inputs and outputs (the names run_decoder feeds and reads,
 * src/tacotron2/mod.rs:285-341), not by initializer names; an LSTM cell may be an ONNX LSTM operator (gate order i, o, f, c,
 * re-ordered here) or its Gemm / Add / Split decomposition.  Fails with XDTTS_ERR_UNSUPPORTED when the graph is not that
 * decoder step or its dimensions are not Tacotron2's. */
int xdtts_decoder_create_from_onnx(const char* path, const xdtts_decoder_opts* opts_or_null, int device, xdtts_decoder** out);
/* host-only access to the same reader (no GPU needed) */
typedef struct xdtts_onnx_decoder xdtts_onnx_decoder;
int xdtts_onnx_decoder_open(const char* path, xdtts_onnx_decoder** out);
void xdtts_onnx_decoder_close(xdtts_onnx_decoder* m);
/* dims10: mel channels, prenet, encoder embedding, attention rnn, decoder rnn, attention dim, location filters, location
 * kernel, LSTM encoding (0: Gemm decomposition, 1: LSTM operators), 1 when the graph draws the prenet dropout mask itself */
int xdtts_onnx_decoder_dims(const xdtts_onnx_decoder* m, int* dims10);
/* which: index of the field in the decoder weight struct above (0 prenet1 ... 17 gate_b), in its layouts.  out == null:
 * returns the number of floats; else copies them (capacity in floats) and returns the count; negative on error */
long long xdtts_onnx_decoder_tensor(const xdtts_onnx_decoder* m, int which, float* out_or_null, long long capacity);
/* device time (CUDA events around the persistent kernel launches) and steps executed by the last call */
int xdtts_decoder_last_timing(const xdtts_decoder* h, float* ms, int* steps);
/* measurement: what a launch for nb utterances in lockstep at encoder length t_enc reads and how its CTAs hand results over.
 * info[0] = weight bytes one decoder step reads, [1] = of those, bytes kept in shared memory for the whole loop (the rest
 * comes from the L2 every step), [2] = hand-overs between CTAs per step, [3] = 1: polled {value, tag} cells, 0: grid barriers */
int xdtts_decoder_info(const xdtts_decoder* h, int nb, int t_enc, long long* info);

/* ---- Streaming: the same tail for a sequence of batches of one shape, copies overlapped with kernels.
 * XdTts::infer returns host samples per call (src/lib.rs:141-157) and the binary loops over chunks
 * (src/lib.rs:83-104); a server doing the same batch after batch would leave the GPU idle during every
 * PCIe copy.  A pipe owns `depth` (1..8; 2 is enough) device-resident slots with one stream each.
 * xdtts_pipe_push enqueues H2D -> [postnet, when pn != null] -> lift -> Griffin-Lim -> D2H for one batch and
 * returns without waiting; it blocks only when every slot is in flight (it then waits for the oldest).  The copies of
 * neighbouring batches overlap kernels; the kernels of consecutive batches run back to back, in push order.
 * All host buffers of a batch -- inputs and outputs -- must stay valid and untouched until that batch has
 * been collected: by xdtts_pipe_pop / xdtts_pipe_flush, or by the push that reuses its slot (`depth` pushes later).
 * Pinned buffers (xdtts_host_alloc) make the copies asynchronous; pageable ones are staged (inputs then
 * cost a wait for that slot's stream inside push).  Results are bit-identical to
 * xdtts_gl_infer_batch / xdtts_tail_infer_batch on the same batch. */
typedef struct xdtts_pipe xdtts_pipe;
int xdtts_pipe_create(xdtts_gl* gl, xdtts_postnet* pn_or_null, const int* Ts, int B, int depth, xdtts_pipe** out);
int xdtts_pipe_push(xdtts_pipe* q, const float* const* mels, const float* const* init_phases_or_null,
                    float* const* out_mels_or_null, float* const* out_waves);
int xdtts_pipe_pop(xdtts_pipe* q);      /* wait for the OLDEST batch in flight: 1 = collected, 0 = nothing in flight, < 0 error */
int xdtts_pipe_flush(xdtts_pipe* q);    /* wait until every pushed batch is in its host buffers */
int xdtts_pipe_pending(xdtts_pipe* q);  /* batches in flight (>= 0), or a negative error */
void xdtts_pipe_destroy(xdtts_pipe* q); /* waits for batches in flight; destroy it before its gl / postnet handles */

/* ---- Multi-GPU: one vocoder per device behind one handle.  The reference is a single process that vocodes one
 * utterance at a time (src/lib.rs:83-104; "could run in parallel", src/phonemes.rs:677-680); utterances are
 * independent, so a batch shards across the GPUs of a box without any data-path exchange.
 * xdtts_pool_create builds one vocoder handle (the arguments of xdtts_gl_create) and one worker thread per device;
 * devices == null / n_devices == 0 means every visible device.  xdtts_pool_infer_batch assigns the utterances
 * longest-first to the least-loaded device (ties to the first listed device), runs the sub-batches concurrently and
 * returns when all are in the caller's buffers.  Every utterance draws the phase stream of its index in THIS call's
 * batch and all devices use the call's seed, so the output does not depend on how many devices there are (bit for
 * bit when opts.run_frames fixes the run length; the automatic run length depends on a device's share of the batch). */
typedef struct xdtts_pool xdtts_pool;
int xdtts_pool_create(const float* mel_basis, int n_mels, int K, int noverlap, float power, int n_iter, float momentum,
                      const xdtts_gl_opts* opts_or_null, const int* devices_or_null, int n_devices, xdtts_pool** out);
void xdtts_pool_destroy(xdtts_pool* p);
int xdtts_pool_n_devices(const xdtts_pool* p);
int xdtts_pool_out_len(const xdtts_pool* p, int T);
/* the assignment rule alone (host only, no GPU needed): slot 0..n_slots-1 of every utterance -- longest first to the
 * least-loaded slot, ties to the lowest slot; what xdtts_b200/shard.py applies across processes */
int xdtts_shard_assign(const int* Ts, int B, int n_slots, int* slot_of_utt);
/* the device each utterance of a batch with these frame counts is sent to */
int xdtts_pool_assignment(const xdtts_pool* p, const int* Ts, int B, int* device_of_utt);
int xdtts_pool_infer_batch(xdtts_pool* p, const float* const* mels, const int* Ts, int B,
                           const float* const* init_phases_or_null, float* const* outs);
int xdtts_pool_from_mag_batch(xdtts_pool* p, const float* const* mags, const int* Ts, int B,
                              const float* const* init_phases_or_null, float* const* outs);

/* .npy I/O for [rows, cols] float32 arrays -- the format of the reference's spectrogram dump
 * (ndarray_npy::write_npy, src/lib.rs:125-139, --output-spectrogram src/bin/app.rs:12-14).
 * read: call with out == null to get the shape, then with a buffer of capacity >= rows*cols floats
 * (1-D files read as [1, n]; Fortran-ordered files are transposed into C order). */
int xdtts_npy_write_f32(const char* path, const float* data, int rows, int cols);
int xdtts_npy_read_f32(const char* path, float* out_or_null, long long capacity, int* rows, int* cols);

/* pinned host memory for callers that want copy/compute overlap and full PCIe speed */
void* xdtts_host_alloc(unsigned long long bytes);
void xdtts_host_free(void* p);

/* thread-local message of the last error returned on this thread ("" if none) */
const char* xdtts_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long xdtts_kernel_launches(void);
/* "xdtts_b200 <version> sm_100a" */
const char* xdtts_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XDTTS_B200_H */
