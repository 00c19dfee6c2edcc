/*
 * b200codec.h -- C ABI of the B200-native (sm_100a) xcodec2-compatible codec DECODE path.
 *
 * This is the drop-in boundary for tts-max's `tts.core.codec.decoding` /
 * `tts.core.codec.decoder.Decoder.forward`. The reference is pure Python and has no FFI;
 * each entry point below names the reference interface it replaces (paths relative to the
 * reference repo root). The Python mirror of the reference interface
 * (tts_max_b200/codec/{decoding,decoder}.py) is a thin ctypes client of this header.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every function returns 0 on success, non-zero on failure; the message is available
 *     from b200codec_last_error() (thread-local). The library never aborts the process:
 *     the reference's callers rely on catchable exceptions
 *     (tts/training/rlhf/rewards.py:86-97).
 *   - the caller owns every input/output buffer and the CUDA stream; the library owns only
 *     the prepared weights and its activation workspace.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are
 *     ordinary (ideally pinned) host memory.
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels.
 *   - Threading: a handle owns ONE plan, workspace and staging area, so it decodes one batch at a
 *     time. The decode entry points take a per-handle mutex (concurrent callers serialise), and a
 *     handle must be used with one stream at a time: work a previous decode left in flight on
 *     ANOTHER stream is not waited for before buffers are rebuilt. Several handles per process are
 *     independent (the reference makes up to three decoders per process, rewards.py:36-38).
 */
#ifndef B200CODEC_H_
#define B200CODEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define B200CODEC_ABI_VERSION 3

/* arithmetic type of the tensor-core operands (accumulation is always fp32; the residual
 * stream, norms, softmax, FSQ lookup and ISTFT are always fp32). */
enum B200CodecPrecision {
    B200CODEC_BF16 = 0, /* tcgen05 kind::f16, bf16 operands (BASELINE config 2 "bf16 decode") */
    B200CODEC_FP16 = 1  /* tcgen05 kind::f16, fp16 operands (same rate, 3 more mantissa bits): the
                           high-precision mode, >= 55 dB against the fp32 reference */
};

/* element type of the ids passed to the lookup / decode entry points */
enum B200CodecIdType { B200CODEC_IDS_I32 = 0, B200CODEC_IDS_I64 = 1 };

/* dtype tags for b200codec_load_tensor */
enum B200CodecDType { B200CODEC_F32 = 0, B200CODEC_F16 = 1, B200CODEC_BF16_T = 2, B200CODEC_F64 = 3 };

/* Mirrors the constructor arguments of `Decoder` (tts/core/codec/decoder.py:17-37) and
 * `DecoderConfig` (tts/core/codec/decoding.py:13-35). */
typedef struct B200CodecConfig {
    int32_t abi_version;      /* must be B200CODEC_ABI_VERSION */
    int32_t sample_rate;      /* 16000 for xcodec2 */
    int32_t hop_length;       /* 320 (xcodec2), 160 (48 kHz), also 240 and 80; n_fft = win = 4 * hop (decoder_modules.py:426-431) */
    int32_t n_upsample;       /* len(upsample_factors): 0 (xcodec2) .. 3 (UpSamplerBlock, upsampler.py:9-69) */
    int32_t precision;        /* enum B200CodecPrecision */
    int32_t device;           /* CUDA device ordinal */
    int32_t hidden_dim;       /* 1024 */
    int32_t depth;            /* 12 transformer blocks */
    int32_t heads;            /* 16 */
    int32_t vq_dim;           /* 2048 */
    int32_t upsample_factors[3]; /* e.g. {3, 2, 0}: ConvTranspose1d strides (decoder.py:48-53) */
    int32_t kernel_sizes[3];     /* e.g. {7, 6, 0}: ConvTranspose1d kernel sizes, padding (k - u) / 2 */
} B200CodecConfig;

typedef struct B200Codec B200Codec;

/* thread-local message of the last failing call ("" if none) */
const char* b200codec_last_error(void);

/* number of state-dict tensors the decoder expects (117 for the xcodec2 config) and the
 * i-th expected key ("decoder.quantizer.project_out.weight", ...), in the order of
 * `Decoder.state_dict()`. Used by the Python shim for strict loading
 * (tts/core/codec/decoder.py:109-110,119). */
int b200codec_num_tensors(const B200Codec* h);
const char* b200codec_tensor_key(const B200Codec* h, int i);
/* shape of the i-th expected tensor; returns ndim (<= 4) */
int b200codec_tensor_shape(const B200Codec* h, int i, int64_t shape_out[4]);

/* Replaces Decoder.__init__ (tts/core/codec/decoder.py:17-67): validates the 50 Hz
 * constraint (:31-37), allocates the module. Weights are NOT initialised. */
int b200codec_create(const B200CodecConfig* cfg, B200Codec** out);
void b200codec_destroy(B200Codec* h);

/* Replaces load_state_dict for one tensor (tts/core/codec/decoder.py:91-119 keeps the two
 * checkpoint layouts in Python and calls this once per `Decoder.state_dict()` key).
 * `host_ptr` is contiguous host memory of `dtype`; unknown keys and shape mismatches fail
 * (strict=True semantics). */
int b200codec_load_tensor(B200Codec* h, const char* key, const void* host_ptr, int dtype,
                          const int64_t* shape, int ndim);

/* Reads the fp32 master copy of a tensor back (state_dict() round trip). */
int b200codec_read_tensor(const B200Codec* h, const char* key, float* host_out, size_t n_elems);

/* Fails unless every expected tensor has been loaded; repacks weights for the tensor cores
 * (operand dtype, conv taps -> K-major slabs, head-indexed RoPE folded into c_attn
 * (decoder_modules.py:280-281), TMA descriptors). Idempotent. */
int b200codec_finalize_weights(B200Codec* h, void* stream);

/* ---- the hot path ------------------------------------------------------------------- */

/* Replaces Decoder.forward (tts/core/codec/decoder.py:69-89) for a VARLEN batch: utterance
 * u has seqlens_host[u] tokens, ids are packed back to back, waveforms are packed back to
 * back with hop_length * prod(upsample_factors) * seqlens[u] samples each
 * (b200codec_samples_per_token: 320 for xcodec2, 960 for the 48 kHz config). Every utterance is decoded with exactly
 * the single-utterance semantics of the reference (per-utterance GroupNorm statistics,
 * unmasked attention within the utterance, zero conv padding at the utterance edges), so an
 * equal-length batch reproduces the reference's batched forward row for row.
 * Asynchronous on `stream`. */
int b200codec_decode_varlen(B200Codec* h, const void* ids_dev, int id_type,
                            const int32_t* seqlens_host, int n_utts, float* wav_dev,
                            void* stream);

/* Same, HOST buffers in and out: the call a serving loop makes
 * (AudioDecoder.decode, tts/core/codec/decoding.py:84-89: ids .to(device) ... .cpu()).
 * H2D of the ids, the decode, D2H of the PCM and the final synchronisation all happen
 * inside. ids are range-checked ([0, 65535]) on the host before anything is launched. */
int b200codec_decode_host(B200Codec* h, const void* ids_host, int id_type,
                          const int32_t* seqlens_host, int n_utts, float* wav_host,
                          void* stream);

/* The same without the final synchronisation, for callers that keep several batches in flight (a dataset sweep:
 * tts/data/data_vectorizer.py's code store decoded bucket by bucket). wav_host_pinned must be page-locked,
 * device-mapped memory: the last kernel stores the PCM straight into it. The call returns after the enqueue;
 * the caller waits on `stream` (or an event recorded behind the call) before reading the PCM, and keeps ids_host
 * valid until then if it is page-locked (pageable ids are staged before the call returns). */
int b200codec_decode_host_async(B200Codec* h, const void* ids_host, int id_type,
                                const int32_t* seqlens_host, int n_utts, float* wav_host_pinned,
                                void* stream);

/* ---- cached streaming (BASELINE config 5's shape) -----------------------------------------
 * The reference has no streaming decoder (tools/serving/inference.py:155-170 decodes once), and chunking
 * changes what the model computes. Two definitions are offered:
 *   "window"  (tts_max_b200/codec/streaming.py, round 1): audio(new) = Decoder.forward(cat(context, new))
 *             trimmed -- every push recomputes the whole [context | new] window (3x the FLOPs of the new audio
 *             at 100 + 50 tokens); its oracle is the reference forward on that window;
 *   "cached"  (this API): a push runs the model on [overlap | new] rows only -- `overlap` (~8) previous tokens
 *             are recomputed so that the convolutions (receptive field 7 rows before the transformer, 4 after)
 *             and the ISTFT overlap-add (2 frames) have their left halo, GroupNorm normalises over these rows,
 *             and attention sees the keys / values of the last `left_context` tokens AS THEY WERE COMPUTED
 *             WHEN THOSE TOKENS WERE NEW (a per-layer key / value ring; the softmax is order-free and the
 *             head-indexed rotary embedding carries no position). Its oracle is the CPU restatement of exactly
 *             this algorithm (oracle/streaming_oracle.py); its quality against the one-shot decode is reported
 *             next to the window method's in tests/test_streaming.py.
 * b200codec_stream_push: ids_dev = [n_streams][overlap + new_tokens] packed, wav_dev receives
 * [n_streams][(overlap + new_tokens) * 320] samples (the caller keeps the last new_tokens * 320 of each stream).
 * `overlap` <= tokens pushed so far. One state per set of lock-step streams; pushes of one state are serial. */
typedef struct B200Stream B200Stream;
int b200codec_stream_create(B200Codec* h, int n_streams, int new_tokens, int left_context, B200Stream** out);
void b200codec_stream_destroy(B200Stream* st);
int b200codec_stream_reset(B200Stream* st);
int b200codec_stream_capacity(const B200Stream* st);  /* tokens per key / value ring */
int64_t b200codec_stream_tokens(const B200Stream* st); /* tokens pushed per stream since the last reset */
int b200codec_stream_push(B200Codec* h, B200Stream* st, const void* ids_dev, int id_type, int overlap,
                          float* wav_dev, void* stream);

/* decode_varlen is asynchronous and takes DEVICE ids, so it cannot range-check them before
 * launching: the FSQ kernel flags ids outside [0, 65535] (and masks them to 16 bits).
 * After synchronising the stream, this returns 1 if any decode since the last call saw such
 * an id (and clears the flag). decode_host checks on the host instead and fails up front. */
int b200codec_take_id_error(B200Codec* h);

/* Monotonic counter, bumped whenever a decode rebuilt or reallocated the handle's plan, workspace or
 * statistics buffers (a new batch shape, a larger batch). A CUDA graph captured from decode_varlen
 * bakes those device pointers and the plan contents in: replay it only while the value is the one
 * seen right after capture, otherwise decode eagerly and capture again
 * (tts_max_b200/codec/streaming.py does exactly that). */
int64_t b200codec_plan_generation(const B200Codec* h);

/* Debug taps for stage-level parity (tests/test_gpu_decode.py::test_stage_taps_*): when on, every
 * decode keeps device copies of named stage tensors -- "embed", "prior_net", "tblock0",
 * "transformers" (fp32 residual stream after backbone.embed / prior_net / transformers[0] / all
 * transformers, decoder_modules.py:392-396), "backbone" (final_layer_norm output, operand dtype,
 * :399), "head_linear" (head.out output, :131) and, with an upsampler, "up<i>", "res<i>",
 * "upsampled" (upsampler.py:62-69). b200codec_read_stage copies one of them to the host as packed
 * token-major fp32 [rows, width]; b200codec_stage_rows / _width return its shape
 * (rows = sum(T) x the upsampling reached at that stage; -1: unknown name). Costs device copies per decode; never enable it in a timed run. */
int b200codec_set_stage_taps(B200Codec* h, int on);
int b200codec_stage_width(B200Codec* h, const char* name);
int64_t b200codec_stage_rows(B200Codec* h, const char* name);
int b200codec_read_stage(B200Codec* h, const char* name, float* host_out, size_t n_elems, void* stream);

/* Front end (ids -> embed output). project_out, fc_post_a and the backbone's embed conv (k = 7) are
 * linear maps back to back, so the conv output is a linear function of the 7 neighbouring codes whose
 * coefficients are folded in fp64 at load time. mode 1 (default): im2col of the codes (exact in 16 bits)
 * + one K = 128 tensor-core GEMM against the coefficients split hi + lo; mode 2: the same fold as an fp32
 * FMA lookup kernel; mode 0: 8 -> 1024 lookup (project_out o fc_post_a) + conv7 GEMM on 16-bit operands.
 * A/B switch; all agree with the fp32 reference, the folds more closely. */
int b200codec_set_frontend_fold(int mode);

/* GEMM tile width (default on): when M is small (B = 1 serving, BASELINE config 1) the 256-wide tiling
 * would give only N / 256 of the 74 CTA pairs a tile; 256 x 64 tiles spread the same work over four
 * times as many. Every output element sees the same K order, so results are bit-identical. A/B switch:
 * 0 = 256-wide tiles only, 1 = default (also: 128- / 192-wide tiles for the encoder's N % 256 != 0 convs),
 * 2 = 64-wide tiles for every N % 256 != 0. */
int b200codec_set_gemm_narrow_tiles(int mode);

/* Default on: a GEMM launch issues the WEIGHT halves of its first pipeline stages before
 * griddepcontrol.wait (weights are never written by a kernel of the decode), so their DRAM latency overlaps
 * the predecessor's tail; the activation halves follow after the wait. b200codec_gemm callers whose `w_dev`
 * is produced by the kernel directly before on the same stream WITH programmatic dependent launch must
 * switch this off. A/B switch. */
int b200codec_set_gemm_early_weights(int on);

/* GEMM chains (default OFF): c_proj -> fc1 -> fc2 -> the next block's c_attn of every transformer block run
 * as one persistent launch whose tiles wait for the 256-row block they read (csrc/gemm_tc05_2cta.cuh)
 * instead of four launches that each drain the grid. Same arithmetic per element. Measured: no faster than
 * the programmatic-dependent-launch chain at config 2 and slower at small M; A/B switch. */
int b200codec_set_gemm_chain(int on);

/* Output hops per ISTFT CTA: 12 (8 warps, two CTAs per SM) or 28 (16 warps, one CTA per SM, less halo
 * recomputation: a tile of H hops transforms H + 4 frames); 0 (default) picks 28 when that still gives every
 * SM a CTA. Same samples either way. A/B switch. */
int b200codec_set_istft_tile(int hops);

/* decode_host with a PINNED output buffer lets the last kernel store the PCM straight into host
 * memory (default on; pageable buffers always take the staged device buffer + copy). A/B switch. */
int b200codec_set_zero_copy_output(int on);

/* Process-wide: launch the kernel chain with programmatic dependent launch (1, default) or as
 * plain stream-ordered launches (0; for A/B measurements). */
int b200codec_set_pdl(int on);

/* output samples per input token: hop_length * prod(upsample_factors) */
int b200codec_samples_per_token(const B200Codec* h);

/* number of kernels the library launched since creation (bench.py's gpu_launches) */
int64_t b200codec_launch_count(const B200Codec* h);

/* Per-stage timing hook: when enabled (on != 0) decode calls record CUDA events around
 * every stage; b200codec_stage_times returns the accumulated device milliseconds per stage
 * name since the last reset. Adds synchronisation; never enable it inside a timed run. */
int b200codec_profile(B200Codec* h, int on);
int b200codec_stage_times(B200Codec* h, int max_stages, const char** names_out,
                          float* ms_out, int* n_out);

/* ---- per-stage entry points (unit parity against the oracle) ------------------------- */

/* K1, ResidualFSQ.get_output_from_indices (vector-quantize-pytorch 1.17.8; called at
 * tts/core/codec/decoder.py:77): n ids -> [n, 2048] fp32, bit-exact with torch CPU
 * (acc = 0; acc += code_k * W[c,k], k = 0..7; + bias[c]). */
int b200codec_fsq_lookup(B200Codec* h, const void* ids_dev, int id_type, int64_t n,
                         float* out_dev, void* stream);

/* LLM token ids -> FSQ code ids on the GPU. Replaces detokenise -> tokenise -> parse "<|s_N|>"
 * (tts/training/rlhf/rewards.py:70-73, tts/inference/inferencing.py:53-63). The speech tokens enter the tokenizer
 * in SORTED order (tts/core/tokenization.py:36-49), so vocabulary id -> N is a permutation: table_dev[v] = N for
 * "<|s_N|>", -1 for every other token. Sequence s is tok_dev[seq_off_dev[s] .. seq_off_dev[s + 1]); its speech
 * tokens are written in order to codes_dev[seq_off_dev[s] ..] and counted in out_len_dev[s]; other tokens
 * (text, <|speech_end|>, padding) are dropped, as extract_speech_ids does. Asynchronous on `stream`. */
int b200codec_map_speech_tokens(const int32_t* table_dev, int vocab, const int64_t* tok_dev,
                                const int32_t* seq_off_dev, int n_seq, int32_t* codes_dev, int32_t* out_len_dev,
                                void* stream);

/* Encode-direction FSQ: ResidualFSQ.forward as called by Encoder.quantize
 * (tts/core/codec/encoder.py:73-78; vector-quantize-pytorch 1.17.8, one quantizer, levels [4]*8):
 * feats_dev [n_tokens, ld] fp32 token-major (ld >= 2048) -> ids (id_type 0: int32, 1: int64):
 * z = project_in(x); digit_d = rint(tanh(z_d + shift) * half_l - 0.5) + 2; id = sum_d digit_d * 4^d.
 * z_dev (optional, [n_tokens, 8]) receives the projected values. pre_bound != 0 applies FSQ.bound
 * once more before the layer (library releases differ; see oracle/codec_oracle.py::fsq_quantize).
 * Asynchronous on `stream`. */
int b200codec_fsq_quantize(B200Codec* h, const float* feats_dev, int ld, int64_t n_tokens,
                           void* ids_dev, int id_type, float* z_dev, int pre_bound, void* stream);

/* K13+K14, ISTFTHead.forward after the Linear + ISTFT.forward "same"
 * (tts/core/codec/decoder_modules.py:131-148, 35-93): x_pred_dev is the head Linear output,
 * packed [sum(T), ld] fp32 (cols 0..640 log-magnitude, 641..1281 phase); writes
 * hop*T samples per utterance. */
int b200codec_istft(B200Codec* h, const float* x_pred_dev, int ld, const int32_t* seqlens_host,
                    int n_utts, float* wav_dev, void* stream);

/* Generic tensor-core GEMM / implicit conv1d used by every dense layer of the path
 * (see csrc/gemm_tc05.cuh, gemm_tc05_2cta.cuh). a: [M, Cin] operand dtype (per `precision`), Cin a
 * multiple of 64; w: [N, taps*Cin] same dtype; out: [M, ldc] (out_dtype: 0 = fp32, 1 = operand dtype),
 * ldc >= N rounded up to 64 (columns [N, ldc) of that range are written as the product with zero weights);
 * bias / residual need N % 64 == 0. out = act(conv(a, w) + bias) + residual. */
int b200codec_gemm(int precision, const void* a_dev, const void* w_dev, int M, int N, int Cin,
                   int taps, void* out_dev, int out_dtype, int ldc, const float* bias_dev,
                   const float* residual_dev, int ld_res, int act, void* stream);

/* RMSNorm / LayerNorm / GroupNorm(32)+swish on token-major [rows, 1024] fp32 input, output
 * in the operand dtype of `precision` (decoder_modules.py:226-236, 373, 151-159). */
int b200codec_rmsnorm(int precision, const float* x_dev, const float* w_dev, int rows, int dim,
                      float eps, void* out_dev, void* stream);
int b200codec_layernorm(int precision, const float* x_dev, const float* w_dev, const float* b_dev,
                        int rows, int dim, float eps, void* out_dev, void* stream);
int b200codec_groupnorm_swish(int precision, const float* x_dev, const float* gamma_dev,
                              const float* beta_dev, const int32_t* seqlens_host, int n_utts,
                              int dim, float eps, void* out_dev, void* stream);

/* Unmasked multi-head attention over packed varlen utterances (decoder_modules.py:283-285):
 * qkv [sum(T), 3*H*64] operand dtype, rows "(r h d)"; out [sum(T), H*64]. */
int b200codec_attention(int precision, const void* qkv_dev, const int32_t* seqlens_host,
                        int n_utts, int heads, void* out_dev, void* stream);

/* ---- encode direction ------------------------------------------------------------------
 * Replaces tts.core.codec.encoder.Encoder.forward (tts/core/codec/encoder.py:58-78) minus the w2v-BERT model:
 * AcousticEncoder (encoder_modules.py:128-191: weight-normed Conv1d stack with dilations 1/3/9, strides
 * 2/2/4/4/5, Activation1d(SnakeBeta) -- activations.py:47-110, filters.py:87-135), SemanticEncoder
 * (encoder_modules.py:72-125), fusion_layer (encoder.py:42,69) and ResidualFSQ.forward (encoder.py:73-78).
 * The handle takes the keys of Encoder.state_dict() (`semantic_encoder.*`, `acoustic_encoder.*`,
 * `fusion_layer.*`, `quantizer.project_{in,out}.*`); the two checkpoint layouts of
 * Encoder.load_from_checkpoint (encoder.py:80-112) stay in Python. Same conventions as above: 0 on success,
 * b200codec_last_error() for the message, no CPU fallback, one encode at a time per handle (mutex). */
typedef struct B200Enc B200Enc;
int b200enc_create(int precision, int device, B200Enc** out);
void b200enc_destroy(B200Enc* h);
int b200enc_num_tensors(const B200Enc* h);
const char* b200enc_tensor_key(const B200Enc* h, int i);
int b200enc_tensor_shape(const B200Enc* h, int i, int64_t shape_out[4]);
/* fp32 host tensor for one state-dict key (strict: unknown keys and shape mismatches fail) */
int b200enc_load_tensor(B200Enc* h, const char* key, const float* host_ptr, const int64_t* shape, int ndim);
/* folds weight_norm (g * v / ||v||), lays the strided convs out per super-row, casts to the operand dtype */
int b200enc_finalize_weights(B200Enc* h, void* stream);
/* ONE utterance: wav_dev [n_samples] fp32 with n_samples a positive multiple of 320 (Encoder.encode pads to
 * that, encoder.py:116-120); w2v_dev [T][1024] fp32 = the w2v-BERT `hidden_states[16]` of the same audio,
 * T = n_samples / 320. Writes ids [T] (id_type 0: int32 like the reference, 1: int64; NULL: skip) and, when
 * non-NULL, hidden_dev [T][2048], acoustic_dev [T][1024], semantic_dev [T][1024] (fp32, token-major).
 * pre_bound: see b200codec_fsq_quantize. Asynchronous on `stream`. */
int b200enc_encode(B200Enc* h, const float* wav_dev, int64_t n_samples, const float* w2v_dev, void* ids_dev,
                   int id_type, int pre_bound, float* hidden_dev, float* acoustic_dev, float* semantic_dev,
                   void* stream);
/* A batch of n_clips equally long clips in ONE launch sequence (80 kernels; reference: Encoder.forward on a
 * (B, 1, S) tensor, encoder.py:58-71): wav_dev [n_clips][n_samples], w2v_dev [n_clips][T][1024], ids
 * [n_clips][T], hidden_dev [n_clips][T][2048], acoustic_dev / semantic_dev [n_clips][T][1024]. The clips share a
 * padded row space; zero gap rows between them are the convolutions' zero padding, so every clip's result is
 * the one b200enc_encode gives for it alone. n_clips x (n_samples + 1920) must stay below 2^31. */
int b200enc_encode_batch(B200Enc* h, const float* wav_dev, int n_clips, int64_t n_samples, const float* w2v_dev,
                         void* ids_dev, int id_type, int pre_bound, float* hidden_dev, float* acoustic_dev,
                         float* semantic_dev, void* stream);
/* stage parity: with taps on, every encode keeps conv_blocks[0 .. 5] outputs ("conv0", "block1" .. "block5");
 * read_stage returns one as fp32 token-major [clips][rows][C] (rows = n_samples / stride so far, C = 48 * 2^i;
 * clips = those of the last encode call) */
int b200enc_set_stage_taps(B200Enc* h, int on);
int b200enc_read_stage(B200Enc* h, const char* name, int64_t n_samples, float* host_out, size_t n_elems,
                       void* stream);
int64_t b200enc_launch_count(const B200Enc* h);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200CODEC_H_ */
