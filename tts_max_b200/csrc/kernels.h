// Host-callable launchers of the sm_100a kernels (one .cu per family).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// precision codes == enum B200CodecPrecision
constexpr int kPrecBf16 = 0;
constexpr int kPrecFp16 = 1;

inline size_t operand_bytes(int /*precision*/) { return 2; }

// The padded row space all activations live in: utterance u occupies rows
// [utt_row0[u], utt_row0[u] + utt_len[u]); `gap` rows that stay zero in every conv
// operand separate consecutive utterances (conv halo). Device arrays.
struct RowSpace {
    int rows = 0;          // total padded rows R
    int n_utts = 0;
    int total_tokens = 0;  // sum(T)
    int max_len = 0;
    const int32_t* row_tok = nullptr;    // [R] packed token index, -1 on halo rows
    const int32_t* row_utt = nullptr;    // [R] utterance index, -1 on halo rows
    const uint8_t* row_valid = nullptr;  // [R] 1 on token rows
    const int32_t* utt_row0 = nullptr;   // [n]
    const int32_t* utt_len = nullptr;    // [n]
    const int32_t* utt_tok0 = nullptr;   // [n] packed token offset (== sample offset / hop)
    const int4* attn_work = nullptr;     // [n_attn_work] {row0, T, q0, 0}
    int n_attn_work = 0;
    const int4* attn128_work = nullptr;  // [n_attn128_work] {q_row, n_q, kv_row0, T_kv}, 128-query tiles
    int n_attn128_work = 0;
    const int4* istft_work = nullptr;    // [n_istft_work] {utt, b0, 0, 0}
    int n_istft_work = 0;
    int istft_hops = 12;                 // output hops per ISTFT work item
};

constexpr int kAttnBlockQ = 64;     // query rows per attention CTA
// output hops per ISTFT CTA (+4 halo frames): 12 (8 warps, 2 CTAs per SM) or 28 (16 warps, 1 CTA per SM)
extern int g_istft_hops;

// ---- fsq.cu ----
// out_prec: -1 -> fp32 output, else operand dtype of that precision.
// row_tok == nullptr means rows are the packed tokens themselves.
int launch_fsq_lookup(const void* ids, int id_type, const int32_t* row_tok, int rows,
                      const float* w_out /*[C,8]*/, const float* b_out /*[C]*/, int channels,
                      void* out, int ld, int out_prec, int* err_flag, cudaStream_t stream);
// fused front end: ids -> embed(fc_post_a(project_out(code))) as one 7-tap x 8-digit lookup
// (m_fold [7][C][8], cb_fold [7][C], b_embed [C]; x fp32 [rows, C], halo rows zero)
int launch_fsq_frontend(const void* ids, int id_type, const int32_t* row_tok, int rows, const float* m_fold,
                        const float* cb_fold, const float* b_embed, int C, float* x, int* err_flag,
                        cudaStream_t stream);
// the folded front end as a GEMM operand: ids -> a [rows, 128] operand dtype (codes of the 7 neighbouring
// frames, "frame exists" flags, constant 1; columns 64..127 repeat 0..63 for the hi/lo coefficient split)
int launch_fsq_im2col(const void* ids, int id_type, const int32_t* row_tok, int rows, int prec, void* a,
                      int* err_flag, cudaStream_t stream);
// encode direction (encoder.py:73-78): features [n_tokens, ld] fp32 -> ids (id_type 0: int32, 1: int64);
// z_out (optional) receives the eight projected values per token
int launch_fsq_quantize(const float* x, int ld, int n_tokens, const float* w_in, const float* b_in,
                        int dim, int pre_bound, void* ids, int id_type, float* z_out, cudaStream_t stream);

// LLM vocabulary ids -> FSQ code ids: table[v] = N for "<|s_N|>", -1 otherwise; sequence s is tok[seq_off[s] ..
// seq_off[s + 1]); its speech tokens are written, in order, to codes[seq_off[s] ..] and counted in out_len[s]
int launch_map_speech_tokens(const int32_t* table, int vocab, const long long* tok, const int32_t* seq_off, int n_seq,
                             int32_t* codes, int32_t* out_len, cudaStream_t stream);

// ---- norms.cu ----
// w == nullptr: no elementwise weight (it is folded into the consumer GEMM's weight columns)
int launch_rmsnorm(int prec, const float* x, const float* w, int rows, int dim, float eps,
                   void* out, cudaStream_t stream);
// row_valid (optional): rows flagged 0 are written as zeros (operand of a following conv)
int launch_layernorm(int prec, const float* x, const float* w, const float* b, int rows, int dim,
                     float eps, void* out, cudaStream_t stream, const uint8_t* row_valid = nullptr);
// stats: double [n_utts][32][2] (sum, sumsq), must be zero on entry
int launch_groupnorm_stats(const float* x, const RowSpace& rs, int dim, double* stats,
                           cudaStream_t stream);
// mean / rstd per (utterance, group) are formed from `stats` (fp64 sums) inside the kernel
int launch_groupnorm_apply_swish(int prec, const float* x, const RowSpace& rs, int dim,
                                 const double* stats, const float* gamma, const float* beta,
                                 float eps, void* out, cudaStream_t stream);

// ---- attention_tc05.cu (tcgen05 / TMEM) ----
int launch_attention_tc05(int prec, const void* qkv, const RowSpace& rs, int heads, void* out,
                          cudaStream_t stream);
// general form: queries from `q` ([q_rows][q_ld], head h at column 64 h), keys / values from `kv`
// ([kv_rows][kv_ld], K of head h at k_col0 + 64 h, V at v_col0 + 64 h); work[i] = {q_row, n_q, kv_row0, T_kv}
int launch_attention_tc05_ex(int prec, const void* q, int q_rows, int q_ld, const void* kv, int kv_rows, int kv_ld,
                             int k_col0, int v_col0, const int4* work, int n_work, int heads, void* out,
                             cudaStream_t stream);

// cached streaming: copy the keys / values of every stream's n_new newest rows (row utt_row0[u] + overlap + i of
// the c_attn output [rows][3 D]) into ring [n_streams * cap][2 D] at slot (wpos + i) % cap of stream u
int launch_kv_scatter(int prec, const void* qkv, const int32_t* utt_row0, int n_streams, int overlap, int n_new, int D,
                      void* ring, int cap, int wpos, cudaStream_t stream);

// ---- istft.cu ----
struct IstftTables {
    const float2* twiddle = nullptr;  // istft_table_words(hop) float2, filled by istft_fill_tables(hop, ...)
    const float* window = nullptr;    // [n_fft] from the checkpoint (decoder.head.istft.window)
};
int istft_table_words(int hop);
void istft_fill_tables(int hop, float2* host);
int launch_istft(const float* x_pred, int ld, const RowSpace& rs, const IstftTables& tab,
                 int hop, float* wav, cudaStream_t stream);

// ---- gemm_tc05.cu ----
enum GemmAct : int { kActNone = 0, kActSilu = 1, kActRelu = 2 };
struct GemmCall {
    int precision;       // operand dtype
    const void* a;       // [a_rows, Cin]
    int a_rows;
    int Cin;
    const void* w;       // [N, taps*Cin]
    int N;               // logical out features (rows of w)
    int taps;
    int tap_pad = -1;    // A row = m + tap * tap_dil - tap_pad; -1 -> (taps / 2) * tap_dil ("same" conv)
    int tap_dil = 1;     // dilation of the conv taps, in rows
    void* out;
    int out_fp32;        // 1 -> fp32 output, 0 -> operand dtype
    int ldc;
    int n_store;         // multiple of 64; columns >= N get bias only (zeros)
    // Narrower-than-tile matrices (the encoder's 48 / 96-channel levels): a_cols > 0 = the columns of A that exist
    // in memory (row pitch a_cols; TMA out-of-bounds fill pads every row with zeros up to Cin); out_cols > 0 = the
    // columns of out / residual that exist (the TMA stores drop the rest; ldc >= out_cols instead of n_store)
    int a_cols = 0;
    int out_cols = 0;
    const float* bias;
    const float* residual;
    int ld_res;
    const uint8_t* row_valid;
    int act;
    // RMSNorm fusion (see GemmParams); all optional
    const float* ss_in = nullptr;
    float ss_inv_dim = 0.f;
    float ss_eps = 0.f;
    void* out16 = nullptr;
    int ld16 = 0;
    float* ss_out = nullptr;
    float out16_scale = 1.f;  // power of two applied to the 16-bit copy (fp16 operands), see GemmParams
    float ss_in_scale = 1.f;  // its inverse, applied with the consumer's row scale
    // optional pre-encoded TMA descriptors (CUtensorMap, 128 B each, 64-byte aligned);
    // when null they are encoded on the fly.
    const void* tmap_a = nullptr;
    const void* tmap_b = nullptr;
    // GroupNorm(32) statistics of the fp32 result, reduced in the epilogue (CTA-pair kernel, N == 1024):
    // gn_stats[(utt * 32 + group) * 2 + {0, 1}] += {sum, sum of squares}; gn_row_utt[row] = utterance or -1
    double* gn_stats = nullptr;
    const int32_t* gn_row_utt = nullptr;
};
constexpr size_t kTmapBytes = 128;
int launch_gemm(const GemmCall& c, cudaStream_t stream);
// GEMM chain: calls[0..n) are plain linears over the same rows where calls[i + 1] reads (as its A operand,
// row scale or residual) only what calls[0..i] wrote in the SAME 256-row block. One persistent launch;
// `counters` = n * ceil(rows / 256) uint32, zero on entry (see gemm_tc05_2cta.cuh).
bool gemm_chain_supported(const GemmCall* calls, int n);
int launch_gemm_chain(const GemmCall* calls, int n, uint32_t* counters, cudaStream_t stream);
// true when launch_gemm accepts the shape (N a multiple of 64; the fused-norm epilogue is always available)
bool gemm_uses_cta_pairs(int a_rows, int n_store);

// ---- weight repack helpers (codec.cu uses them at finalize) ----
// dst[n, tap*Cin + c] = src[n, c, tap]  (src is torch Conv1d weight [Cout, Cin, taps]), cast
// col_scale (optional, [Cin], taps == 1): dst[n, c] = src[n, c] * col_scale[c]  (norm weight fold)
int launch_repack_weight(int prec, const float* src, void* dst, int N, int Cin, int taps,
                         cudaStream_t stream, const float* col_scale = nullptr);
// weight_norm (dim 0): scale[c] = g[c] / ||v[c, :]||_2 over the `inner` trailing elements
int launch_weightnorm_scale(const float* g, const float* v, int C0, int inner, float* scale, cudaStream_t stream);
// one ConvTranspose1d output phase: dst[co, t * Cin + ci] = v[ci, co, tap_ids[t]] * scale[ci]
// (v is the torch ConvTranspose1d weight_v [Cin, Cout, k])
int launch_repack_convT_phase(int prec, const float* v, const float* scale, void* dst, int Cin, int Cout, int k,
                              int taps, const int* tap_ids, cudaStream_t stream);
// in-place on fp32 c_attn weight [3*H*64, K]: rotate q,k row pairs by the head-indexed angle
int launch_fold_rope(float* w_qkv, int heads, int head_dim, int K, const float* cos_tab,
                     const float* sin_tab, cudaStream_t stream);

}  // namespace b200
