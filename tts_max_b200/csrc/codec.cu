// b200codec: handle, weights, workspace, the decode forward pass and the C ABI (include/b200codec.h).
//
// Forward pass == Decoder.forward (tts/core/codec/decoder.py:69-89) ->
// VocosBackbone.forward (tts/core/codec/decoder_modules.py:390-400) ->
// ISTFTHead.forward (:118-148), restructured for B200:
//   * activations are token-major [rows, C] in one PADDED ROW SPACE for the whole varlen batch
//     (3 zero rows between utterances), so every Linear is a plain GEMM over all rows, every
//     Conv1d is the same GEMM kernel with row-shifted K-slabs, and the transposes of
//     :391,395,397,399 disappear;
//   * GEMM operands are 16-bit (bf16 / fp16) with fp32 accumulation in TMEM; the residual
//     stream, all norm statistics, softmax, the FSQ lookup and the ISTFT stay fp32;
//   * bias / SiLU / residual-add / halo masking live in the GEMM epilogue.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200codec.h"
#include "common.cuh"
#include "kernels.h"

namespace b200 {
extern int g_gemm_narrow_tiles;  // gemm_tc05.cu: 256 x 64 tiles for small M (default on)
extern int g_gemm_early_weights;  // gemm_tc05.cu: weight loads before griddepcontrol.wait (default on)

int g_use_pdl = 1;
// ids -> embed output: 1 = folded, im2col of the codes + one K = 128 tensor-core GEMM (default);
// 2 = folded, fp32 FMA lookup kernel; 0 = 8 -> 1024 lookup + conv7 GEMM on 16-bit operands
static int g_frontend_fold = 1;
static int g_zero_copy_out = 1;  // decode_host: write PCM directly into pinned host memory
// c_proj -> fc1 -> fc2 -> next c_attn of a transformer block as ONE persistent launch (gemm_tc05_2cta.cuh).
// OFF by default: measured on config 2 it is within noise of the programmatic-dependent-launch chain of four
// kernels (3.22-3.30 vs 3.25-3.28 ms per step) and it is slower at small M, where a chain serialises on its
// single m-block (config 1: 1.41 vs 1.06 ms; config 5: 4.88 vs 4.20 ms). Kept as an A/B switch with its tests.
static int g_gemm_chain = 0;

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

int run_attention(int prec, const void* qkv, const RowSpace& rs, int heads, void* out,
                  cudaStream_t s) {
    return launch_attention_tc05(prec, qkv, rs, heads, out, s);
}

constexpr int kGap = 3;          // zero rows between utterances (conv7 halo)
constexpr int kMaxUp = 3;        // upsampler stages (UpSamplerBlock, tts/core/codec/upsampler.py)

struct TensorSpec {
    std::string key;
    std::vector<int64_t> shape;
    size_t numel() const {
        size_t n = 1;
        for (auto d : shape) n *= static_cast<size_t>(d);
        return n;
    }
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        // grow geometrically so varlen batches of similar size do not thrash the allocator
        size_t want = need + need / 4;
        B200_CUDA_OK(cudaMalloc(&p, want));
        bytes = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T* as() const {
        return static_cast<T*>(p);
    }
};

struct ResBlockW {
    const float *gn1_w, *gn1_b, *b1, *gn2_w, *gn2_b, *b2;
    void *w1, *w2;  // operand dtype [1024, 3*1024]
};
struct LayerW {
    const float *att_norm, *ffn_norm;
    void *qkv, *proj, *fc1, *fc2;
};

// one ConvTranspose1d output phase as an ordinary row-shifted conv (see finalize)
struct UpPhase {
    void* w = nullptr;  // operand dtype [Cout, taps * Cin]
    int taps = 0;
    int pad = 0;        // A row = m + tap - pad
};
struct UpStageW {
    int Cin = 0, Cout = 0, stride = 0, k = 0;
    UpPhase phase[8];
    const float* bias = nullptr;
    ResBlockW res;
};

struct StageTimer {
    std::string name;
    cudaEvent_t a, b;
};

}  // namespace
}  // namespace b200

using namespace b200;

// Cached streaming state: per transformer layer, a ring of the keys and values of each
// stream's last `cap` tokens, as computed when those tokens were new. See b200codec.h for the semantics.
struct B200Stream {
    B200Codec* owner = nullptr;
    int n_streams = 0, new_tokens = 0, left_context = 0, cap = 0;
    int64_t seen = 0;            // tokens pushed per stream so far
    void* kv_slab = nullptr;     // [layers][n_streams * cap + 128][2 C] operand dtype
    size_t layer_stride = 0;     // bytes between layers
    int4* work_dev = nullptr;    // attention work items of the current push
    int4* work_host = nullptr;   // pinned
    int work_cap = 0;
    // set by push for forward_impl
    int overlap = 0, wpos = 0, t_kv = 0, n_work = 0;
    int list_len = -1, list_tkv = -1;  // what the uploaded work list was built for
};

struct B200Codec {
    B200CodecConfig cfg;
    int C, H, L, V, hop, n_fft, n_bins;
    int n_up = 0, total_up = 1, head_ld = 0;  // upsampler stages, prod(factors), padded head.out width
    int up_f[kMaxUp] = {0, 0, 0}, up_k[kMaxUp] = {0, 0, 0};
    std::vector<TensorSpec> specs;
    std::map<std::string, int> index;
    std::vector<float*> master;  // fp32 device copies, nullptr until loaded
    bool finalized = false;

    // prepared weights
    DevBuf wbuf;        // one slab for all operand-dtype weights
    void* w_embed = nullptr;
    ResBlockW res[4];
    LayerW layers[64];
    void* w_head = nullptr;
    UpStageW up[kMaxUp];
    void* w_out_proj = nullptr;
    float* head_bias_pad = nullptr;
    float* w_pre = nullptr;  // fc_post_a o project_out folded: [C, 8] codebook projection ...
    float* b_pre = nullptr;  // ... and [C] bias (fp32)
    float* m_fold = nullptr;   // embed o fc_post_a o project_out: [7][C][8] ...
    float* cb_fold = nullptr;  // ... and the per-tap bias part [7][C] (fp32 FMA form, A/B path 2)
    void* w_front = nullptr;   // the same coefficients as a GEMM operand [C][128] = [hi | lo], operand dtype
    float2* twiddle = nullptr;
    float *rope_cos = nullptr, *rope_sin = nullptr;

    // workspace (grow-only)
    DevBuf ws;
    int ws_rows = 0;
    void *xc, *an, *qkv, *y, *f, *xb;  // operand dtype
    float *x, *hbuf, *ho, *ss;
    // upsampler stage i lives in row space plan[i + 1] at up[i].Cout channels
    float *u_x[kMaxUp], *u_h[kMaxUp];
    void *u_an[kMaxUp], *u_a16[kMaxUp];
    void* u_fin = nullptr;  // out_proj + swish output, operand dtype [rows_last, C]
    double* gn_stats = nullptr;
    size_t gn_stats_bytes = 0;
    uint32_t* chain_ctr = nullptr;   // GEMM-chain m-block counters, behind the statistics (zeroed with them)
    size_t stats_zero_bytes = 0;     // bytes of gn_stats + chain_ctr to clear at the start of a decode
    int gn_slots = 8;  // GroupNorm layers: 8 in the backbone + 2 per upsampler stage
    // plan (row space) cache; plan i > 0 is the row space after upsampler stage i - 1
    // (every length, offset and gap of plan 0 multiplied by the product of the strides so far)
    DevBuf plan_dev, plan_up_dev[kMaxUp];
    // page-locked staging for plan uploads: a ring, so that building the plan of the next shape does not wait for
    // the stream (the device-side plan is rewritten in stream order; only the host staging needs a guard)
    struct PlanStage {
        void* host = nullptr;
        size_t bytes = 0;
        cudaEvent_t copied = nullptr;
        bool pending = false;
    };
    static constexpr int kPlanStages = 4;
    PlanStage plan_stage[kPlanStages];
    int plan_stage_next = 0;
    std::vector<int32_t> plan_key;
    int plan_gap = -1;
    int plan_istft_mode = -1;
    RowSpace rs, rs_up[kMaxUp];
    // io staging for decode_host
    DevBuf io_ids, io_wav;
    int* err_flag_host = nullptr;  // mapped pinned
    int* err_flag_dev = nullptr;

    size_t l2_persist_bytes = 0;  // persisting-L2 carve-out granted at create (0 = unsupported)
    size_t l2_window_max = 0;

    // Debug taps (b200codec_set_stage_taps): device copies of named stage tensors of the LAST decode,
    // in the padded row space they were computed in (b200codec_read_stage compacts them to token rows).
    struct StageTap {
        DevBuf buf;
        int width = 0;      // logical columns
        int ld = 0;         // elements per row as stored
        int elem = 4;       // 4: fp32, 2: operand dtype
        int space = 0;      // 0: plan 0 (token rate), i > 0: row space after upsampler stage i - 1
        bool valid = false;
    };
    bool taps_on = false;
    std::map<std::string, StageTap> taps;
    // One decode at a time per handle: plan, workspace, statistics and staging buffers are single
    // instances that a decode rewrites and may reallocate (see b200codec.h "Threading").
    std::mutex mu;
    // bumped whenever the plan, the workspace or the statistics buffer is rebuilt or reallocated: a
    // CUDA graph captured from a decode bakes those pointers and contents in (b200codec_plan_generation)
    int64_t generation = 0;

    B200Stream* cur_stream = nullptr;  // non-null only inside b200codec_stream_push

    int64_t launches = 0;
    bool profiling = false;
    std::vector<StageTimer> timers;
    std::map<std::string, float> stage_ms;
    std::vector<std::string> stage_order;

    const float* m(const std::string& key) const {
        auto it = index.find(key);
        return it == index.end() ? nullptr : master[it->second];
    }
};

namespace {

void add_spec(B200Codec* h, const std::string& key, std::vector<int64_t> shape) {
    h->index[key] = static_cast<int>(h->specs.size());
    h->specs.push_back({key, std::move(shape)});
}

void build_specs(B200Codec* h) {
    const int64_t C = h->C, V = h->V;
    const std::string g = "decoder.";
    add_spec(h, g + "quantizer.project_in.weight", {8, V});
    add_spec(h, g + "quantizer.project_in.bias", {8});
    add_spec(h, g + "quantizer.project_out.weight", {V, 8});
    add_spec(h, g + "quantizer.project_out.bias", {V});
    add_spec(h, g + "backbone.embed.weight", {C, C, 7});
    add_spec(h, g + "backbone.embed.bias", {C});
    auto resnet = [&](const std::string& p) {
        add_spec(h, p + "norm1.weight", {C});
        add_spec(h, p + "norm1.bias", {C});
        add_spec(h, p + "conv1.weight", {C, C, 3});
        add_spec(h, p + "conv1.bias", {C});
        add_spec(h, p + "norm2.weight", {C});
        add_spec(h, p + "norm2.bias", {C});
        add_spec(h, p + "conv2.weight", {C, C, 3});
        add_spec(h, p + "conv2.bias", {C});
    };
    resnet(g + "backbone.prior_net.0.");
    resnet(g + "backbone.prior_net.1.");
    for (int l = 0; l < h->L; ++l) {
        const std::string p = g + "backbone.transformers." + std::to_string(l) + ".";
        add_spec(h, p + "att_norm.weight", {C});
        add_spec(h, p + "ffn_norm.weight", {C});
        add_spec(h, p + "att.c_attn.weight", {3 * C, C});
        add_spec(h, p + "att.c_proj.weight", {C, C});
        add_spec(h, p + "mlp.fc1.weight", {4 * C, C});
        add_spec(h, p + "mlp.fc2.weight", {C, 4 * C});
    }
    add_spec(h, g + "backbone.final_layer_norm.weight", {C});
    add_spec(h, g + "backbone.final_layer_norm.bias", {C});
    resnet(g + "backbone.post_net.0.");
    resnet(g + "backbone.post_net.1.");
    add_spec(h, g + "head.out.weight", {h->n_fft + 2, C});
    add_spec(h, g + "head.out.bias", {h->n_fft + 2});
    add_spec(h, g + "head.istft.window", {h->n_fft});
    // UpSamplerBlock (upsampler.py:27-60), registered before fc_post_a (decoder.py:48-63)
    for (int i = 0; i < h->n_up; ++i) {
        const std::string p = "upsampler.upsample_layers." + std::to_string(i) + ".";
        const int64_t cin = C >> i, cout = C >> (i + 1);
        add_spec(h, p + "bias", {cout});
        add_spec(h, p + "weight_g", {cin, 1, 1});
        add_spec(h, p + "weight_v", {cin, cout, h->up_k[i]});
    }
    for (int i = 0; i < h->n_up; ++i) {
        const std::string p = "upsampler.resnet_blocks." + std::to_string(i) + ".";
        const int64_t c = C >> (i + 1);
        add_spec(h, p + "norm1.weight", {c});
        add_spec(h, p + "norm1.bias", {c});
        add_spec(h, p + "conv1.weight", {c, c, 3});
        add_spec(h, p + "conv1.bias", {c});
        add_spec(h, p + "temb_proj.weight", {c, 512});  // unused at inference (temb=None), but in the state dict
        add_spec(h, p + "temb_proj.bias", {c});
        add_spec(h, p + "norm2.weight", {c});
        add_spec(h, p + "norm2.bias", {c});
        add_spec(h, p + "conv2.weight", {c, c, 3});
        add_spec(h, p + "conv2.bias", {c});
    }
    if (h->n_up > 0) {
        add_spec(h, "upsampler.out_proj.weight", {C, C >> h->n_up});
        add_spec(h, "upsampler.out_proj.bias", {C});
    }
    add_spec(h, "fc_post_a.weight", {C, V});
    add_spec(h, "fc_post_a.bias", {C});
}

// ---------------------------------------------------------------------------
// row-space plan
// ---------------------------------------------------------------------------
struct PlanLayout {
    size_t R = 0;
    int n_utts = 0;
    int64_t toks = 0;
    int max_len = 0;
    int64_t n_attn = 0, n_attn128 = 0, n_istft = 0;
    int istft_hops = 12;
    size_t off_row_tok, off_row_utt, off_u0, off_ul, off_ut, off_attn, off_attn128, off_istft, off_valid;
    size_t total_bytes = 0;
};

int plan_layout(const int32_t* seqlens, int n_utts, int gap, PlanLayout* L) {
    B200_CHECK(n_utts > 0, "decode: empty batch (no utterances)");
    // ISTFT tile: 28 output hops (16 warps, 32 frames: 1.14x halo recomputation) once that still gives every
    // SM a CTA, else 12 hops (8 warps, 1.33x) so that short batches spread over more SMs
    L->istft_hops = g_istft_hops;
    if (L->istft_hops == 0) {
        int64_t tiles28 = 0;
        for (int u = 0; u < n_utts; ++u) tiles28 += (seqlens[u] + 27) / 28;
        L->istft_hops = tiles28 >= kNumSMs ? 28 : 12;
    }
    int64_t rows = 0;
    for (int u = 0; u < n_utts; ++u) {
        const int T = seqlens[u];
        B200_CHECK(T > 0, "decode: utterance %d is empty (length %d); the reference's callers "
                   "guard T == 0 (rewards.py:76-82)", u, T);
        rows += T + (u + 1 < n_utts ? gap : 0);
        L->toks += T;
        L->max_len = T > L->max_len ? T : L->max_len;
        L->n_attn += (T + kAttnBlockQ - 1) / kAttnBlockQ;
        L->n_attn128 += (T + 127) / 128;
        L->n_istft += (T + L->istft_hops - 1) / L->istft_hops;
    }
    B200_CHECK(rows < (1 << 30), "decode: batch too large (%lld rows)", (long long)rows);
    const size_t R = static_cast<size_t>(rows);
    L->R = R;
    L->n_utts = n_utts;
    // layout (int32 units): row_tok[R] row_utt[R] utt_row0[n] utt_len[n] utt_tok0[n]
    //                       attn_work[4*n_attn] istft_work[4*n_istft] row_valid[R bytes]
    L->off_row_tok = 0;
    L->off_row_utt = R;
    L->off_u0 = 2 * R;
    L->off_ul = 2 * R + n_utts;
    L->off_ut = 2 * R + 2 * n_utts;
    L->off_attn = (2 * R + 3 * n_utts + 3) & ~static_cast<size_t>(3);  // int4 aligned
    L->off_attn128 = L->off_attn + 4 * L->n_attn;
    L->off_istft = L->off_attn128 + 4 * L->n_attn128;
    L->off_valid = L->off_istft + 4 * L->n_istft;
    L->total_bytes = (L->off_valid * 4 + R + 255) & ~static_cast<size_t>(255);
    return 0;
}

void plan_fill(const PlanLayout& L, const int32_t* seqlens, int gap, void* host) {
    int32_t* hp = static_cast<int32_t*>(host);
    uint8_t* hv = reinterpret_cast<uint8_t*>(hp + L.off_valid);
    // work items of one utterance are contiguous, which keeps its K/V in L2 while its
    // query tiles run.
    int64_t r = 0, t0 = 0, ia = 0, ia2 = 0, ii = 0;
    for (int u = 0; u < L.n_utts; ++u) {
        const int T = seqlens[u];
        hp[L.off_u0 + u] = static_cast<int32_t>(r);
        hp[L.off_ul + u] = T;
        hp[L.off_ut + u] = static_cast<int32_t>(t0);
        for (int t = 0; t < T; ++t) {
            hp[L.off_row_tok + r + t] = static_cast<int32_t>(t0 + t);
            hp[L.off_row_utt + r + t] = u;
            hv[r + t] = 1;
        }
        for (int q0 = 0; q0 < T; q0 += kAttnBlockQ) {
            int32_t* w = hp + L.off_attn + 4 * ia++;
            w[0] = static_cast<int32_t>(r); w[1] = T; w[2] = q0; w[3] = 0;
        }
        for (int q0 = 0; q0 < T; q0 += 128) {
            int32_t* w = hp + L.off_attn128 + 4 * ia2++;  // {q_row, n_q, kv_row0, T_kv}
            w[0] = static_cast<int32_t>(r) + q0; w[1] = T - q0 < 128 ? T - q0 : 128; w[2] = static_cast<int32_t>(r); w[3] = T;
        }
        for (int b0 = 0; b0 < T; b0 += L.istft_hops) {
            int32_t* w = hp + L.off_istft + 4 * ii++;
            w[0] = u; w[1] = b0; w[2] = 0; w[3] = 0;
        }
        r += T;
        t0 += T;
        if (u + 1 < L.n_utts)
            for (int gi = 0; gi < gap; ++gi, ++r) {
                hp[L.off_row_tok + r] = -1;
                hp[L.off_row_utt + r] = -1;
                hv[r] = 0;
            }
    }
}

void plan_bind(const PlanLayout& L, const void* dev, RowSpace* rs) {
    const int32_t* dp = static_cast<const int32_t*>(dev);
    rs->rows = static_cast<int>(L.R);
    rs->n_utts = L.n_utts;
    rs->total_tokens = static_cast<int>(L.toks);
    rs->max_len = L.max_len;
    rs->row_tok = dp + L.off_row_tok;
    rs->row_utt = dp + L.off_row_utt;
    rs->utt_row0 = dp + L.off_u0;
    rs->utt_len = dp + L.off_ul;
    rs->utt_tok0 = dp + L.off_ut;
    rs->attn_work = reinterpret_cast<const int4*>(dp + L.off_attn);
    rs->n_attn_work = static_cast<int>(L.n_attn);
    rs->attn128_work = reinterpret_cast<const int4*>(dp + L.off_attn128);
    rs->n_attn128_work = static_cast<int>(L.n_attn128);
    rs->istft_work = reinterpret_cast<const int4*>(dp + L.off_istft);
    rs->n_istft_work = static_cast<int>(L.n_istft);
    rs->istft_hops = L.istft_hops;
    rs->row_valid = reinterpret_cast<const uint8_t*>(dp + L.off_valid);
}

int build_plan(B200Codec* h, const int32_t* seqlens, int n_utts, int gap, cudaStream_t stream) {
    B200_CHECK(n_utts > 0, "decode: empty batch (no utterances)");
    bool same = h->plan_gap == gap && h->plan_istft_mode == g_istft_hops && static_cast<int>(h->plan_key.size()) == n_utts &&
                std::memcmp(h->plan_key.data(), seqlens, sizeof(int32_t) * n_utts) == 0;
    if (same) return 0;
    h->generation++;
    // layouts of the token row space and of every upsampled row space (out row = stride * in row + phase, so
    // everything scales by the stride)
    PlanLayout L, Lu[kMaxUp];
    if (plan_layout(seqlens, n_utts, gap, &L)) return 1;
    std::vector<int32_t> scaled[kMaxUp];
    int up_gap[kMaxUp];
    auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    size_t need = al(L.total_bytes);
    {
        int factor = 1;
        for (int i = 0; i < h->n_up; ++i) {
            factor *= h->up_f[i];
            scaled[i].resize(n_utts);
            for (int u = 0; u < n_utts; ++u) scaled[i][u] = seqlens[u] * factor;
            up_gap[i] = gap * factor;
            if (plan_layout(scaled[i].data(), n_utts, up_gap[i], &Lu[i])) return 1;
            need += al(Lu[i].total_bytes);
        }
    }
    // No stream synchronisation: the device plan is rewritten by copies in stream order (behind whatever still
    // reads the previous plan); the host staging buffer comes from a ring and is only refilled once the copy that
    // last read it has executed. (Growing a device buffer frees the old one, which waits for the device.)
    B200Codec::PlanStage& st = h->plan_stage[h->plan_stage_next];
    h->plan_stage_next = (h->plan_stage_next + 1) % B200Codec::kPlanStages;
    if (st.pending) {
        B200_CUDA_OK(cudaEventSynchronize(st.copied));
        st.pending = false;
    }
    if (need > st.bytes) {
        if (st.host) cudaFreeHost(st.host);
        st.host = nullptr;
        st.bytes = 0;
        B200_CUDA_OK(cudaMallocHost(&st.host, need * 2));
        st.bytes = need * 2;
    }
    if (st.copied == nullptr) B200_CUDA_OK(cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming));
    uint8_t* hp = static_cast<uint8_t*>(st.host);
    if (h->plan_dev.ensure(L.total_bytes)) return 1;
    plan_fill(L, seqlens, gap, hp);
    B200_CUDA_OK(cudaMemcpyAsync(h->plan_dev.p, hp, L.total_bytes, cudaMemcpyHostToDevice, stream));
    plan_bind(L, h->plan_dev.p, &h->rs);
    hp += al(L.total_bytes);
    for (int i = 0; i < h->n_up; ++i) {
        if (h->plan_up_dev[i].ensure(Lu[i].total_bytes)) return 1;
        plan_fill(Lu[i], scaled[i].data(), up_gap[i], hp);
        B200_CUDA_OK(cudaMemcpyAsync(h->plan_up_dev[i].p, hp, Lu[i].total_bytes, cudaMemcpyHostToDevice, stream));
        plan_bind(Lu[i], h->plan_up_dev[i].p, &h->rs_up[i]);
        hp += al(Lu[i].total_bytes);
    }
    B200_CUDA_OK(cudaEventRecord(st.copied, stream));
    st.pending = true;
    h->plan_key.assign(seqlens, seqlens + n_utts);
    h->plan_gap = gap;
    h->plan_istft_mode = g_istft_hops;
    return 0;
}

// throw-away packed (gap = 0) row space for the per-stage entry points
struct TempPlan {
    void* dev = nullptr;
    RowSpace rs;
    int build(const int32_t* seqlens, int n_utts, cudaStream_t s) {
        PlanLayout L;
        if (plan_layout(seqlens, n_utts, 0, &L)) return 1;
        std::vector<uint8_t> host(L.total_bytes, 0);
        plan_fill(L, seqlens, 0, host.data());
        B200_CUDA_OK(cudaMalloc(&dev, L.total_bytes));
        B200_CUDA_OK(cudaMemcpyAsync(dev, host.data(), L.total_bytes, cudaMemcpyHostToDevice, s));
        B200_CUDA_OK(cudaStreamSynchronize(s));
        plan_bind(L, dev, &rs);
        return 0;
    }
    ~TempPlan() {
        if (dev) cudaFree(dev);
    }
};

int ensure_workspace(B200Codec* h, int rows) {
    if (rows <= h->ws_rows) return 0;
    const size_t R = static_cast<size_t>(rows) + rows / 4 + 128;
    const size_t es = operand_bytes(h->cfg.precision);
    const size_t C = h->C;
    auto al = [](size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); };
    const size_t Rlast = R * h->total_up;  // rows of the last (upsampled) row space
    size_t sz_c = al(R * C * es), sz_qkv = al(R * 3 * C * es),
           sz_f = al(R * 4 * C * es), sz_x = al(R * C * 4), sz_ho = al(Rlast * h->head_ld * 4);
    size_t sz_ss = al(R * 32 * 4);  // kGemmSsSlots row partials
    size_t total = 4 * sz_c + sz_qkv + sz_f + 2 * sz_x + sz_ho + sz_ss;
    size_t up_rows[kMaxUp], sz_up32[kMaxUp], sz_up16[kMaxUp];
    {
        size_t f = 1;
        for (int i = 0; i < h->n_up; ++i) {
            f *= h->up_f[i];
            up_rows[i] = R * f;
            sz_up32[i] = al(up_rows[i] * h->up[i].Cout * 4);
            sz_up16[i] = al(up_rows[i] * h->up[i].Cout * es);
            total += 2 * sz_up32[i] + 2 * sz_up16[i];
        }
        if (h->n_up > 0) total += al(Rlast * C * es);
    }
    h->generation++;
    h->ws.release();
    h->ws_rows = 0;
    if (h->ws.ensure(total)) return 1;
    B200_CUDA_OK(cudaMemset(h->ws.p, 0, h->ws.bytes));  // no NaN garbage in halo rows
    uint8_t* p = h->ws.as<uint8_t>();
    h->xc = p; p += sz_c;
    h->an = p; p += sz_c;
    h->y = p; p += sz_c;
    h->xb = p; p += sz_c;
    h->qkv = p; p += sz_qkv;
    h->f = p; p += sz_f;
    h->x = reinterpret_cast<float*>(p); p += sz_x;
    h->hbuf = reinterpret_cast<float*>(p); p += sz_x;
    h->ho = reinterpret_cast<float*>(p); p += sz_ho;
    h->ss = reinterpret_cast<float*>(p); p += sz_ss;
    for (int i = 0; i < h->n_up; ++i) {
        h->u_x[i] = reinterpret_cast<float*>(p); p += sz_up32[i];
        h->u_h[i] = reinterpret_cast<float*>(p); p += sz_up32[i];
        h->u_an[i] = p; p += sz_up16[i];
        h->u_a16[i] = p; p += sz_up16[i];
    }
    if (h->n_up > 0) { h->u_fin = p; p += al(Rlast * C * es); }
    h->ws_rows = static_cast<int>(R);
    return 0;
}

// ---------------------------------------------------------------------------
// profiling helpers
// ---------------------------------------------------------------------------
struct Stage {
    B200Codec* h;
    cudaStream_t s;
    bool on;
    Stage(B200Codec* h_, const char* name, cudaStream_t s_) : h(h_), s(s_), on(h_->profiling) {
        if (!on) return;
        StageTimer t;
        t.name = name;
        cudaEventCreate(&t.a);
        cudaEventCreate(&t.b);
        cudaEventRecord(t.a, s);
        h->timers.push_back(t);
    }
    ~Stage() {
        if (on) cudaEventRecord(h->timers.back().b, s);
    }
};

void collect_timers(B200Codec* h) {
    for (auto& t : h->timers) {
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.a, t.b);
        if (h->stage_ms.find(t.name) == h->stage_ms.end()) h->stage_order.push_back(t.name);
        h->stage_ms[t.name] += ms;
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    h->timers.clear();
}

// ---------------------------------------------------------------------------
// the forward pass
// ---------------------------------------------------------------------------
#define RUN(expr)                \
    do {                         \
        if ((expr) != 0) return 1; \
        h->launches++;           \
    } while (0)

// RMSNorm fusion hooks of one GEMM call (all optional)
struct NormFuse {
    const float* ss_in = nullptr;  // consume: scale rows by rsqrt(mean(x^2) + eps)
    void* out16 = nullptr;         // produce: 16-bit copy of the fp32 result ...
    float* ss_out = nullptr;       // ... and its per-row sum-of-squares partials
    double* gn_stats = nullptr;    // produce: GroupNorm statistics of the fp32 result (N == 1024 only)
    float out16_scale = 1.f;       // produce: power-of-two scale of the 16-bit copy (fp16 operands)
    float ss_in_scale = 1.f;       // consume: its inverse
};

GemmCall gemm_call(B200Codec* h, const void* a, int Cin, const void* w, int N, int taps, void* out,
                   bool out_fp32, int ldc, int n_store, const float* bias, const float* residual, int act,
                   bool mask_rows, const NormFuse& nf = NormFuse()) {
    GemmCall c;
    c.precision = h->cfg.precision;
    c.a = a;
    c.a_rows = h->rs.rows;
    c.Cin = Cin;
    c.w = w;
    c.N = N;
    c.taps = taps;
    c.out = out;
    c.out_fp32 = out_fp32 ? 1 : 0;
    c.ldc = ldc;
    c.n_store = n_store;
    c.bias = bias;
    c.residual = residual;
    c.ld_res = ldc;
    c.row_valid = mask_rows ? h->rs.row_valid : nullptr;
    c.act = act;
    c.ss_in = nf.ss_in;
    c.ss_inv_dim = 1.f / static_cast<float>(h->C);
    c.ss_eps = 1e-6f;
    c.out16 = nf.out16;
    c.ld16 = h->C;
    c.ss_out = nf.ss_out;
    c.gn_stats = nf.gn_stats;
    c.gn_row_utt = h->rs.row_utt;
    c.out16_scale = nf.out16_scale;
    c.ss_in_scale = nf.ss_in_scale;
    return c;
}

int gemm(B200Codec* h, const void* a, int Cin, const void* w, int N, int taps, void* out,
         bool out_fp32, int ldc, int n_store, const float* bias, const float* residual, int act,
         bool mask_rows, cudaStream_t s, const NormFuse& nf = NormFuse()) {
    return launch_gemm(gemm_call(h, a, Cin, w, N, taps, out, out_fp32, ldc, n_store, bias, residual, act, mask_rows, nf), s);
}

// ResnetBlock (decoder_modules.py:201-223) on one row space at C channels:
//   x <- x + conv2(swish(GN2(conv1(swish(GN1(x))))))
static bool g_fuse_rmsnorm = true;  // RMSNorm inside the GEMM epilogues (false: stand-alone kernel)
static bool g_gn_in_gemm = true;  // GroupNorm statistics reduced in the producing GEMM's epilogue
static bool gn_stats_in_gemm(int rows, int C) {
    return g_gn_in_gemm && C == 1024 && gemm_uses_cta_pairs(rows, C);
}

struct ResCtx {
    const RowSpace* rs;
    int C;
    float* x;        // fp32 [rows, C], updated in place
    float* hb;       // fp32 scratch [rows, C]
    void* an;        // operand scratch [rows, C]
    bool mask_out;   // write zeros on halo rows of x / out16 (x feeds a (transposed) conv directly)
};

// gn1_done: the producer of x already reduced this block's first GroupNorm statistics into its
// slot (a GEMM epilogue, see NormFuse::gn_stats). next_gn: where conv2 should leave the statistics
// of the block output for the GroupNorm that consumes it next (nullptr: nobody).
int resnet_block_ex(B200Codec* h, const ResCtx& cx, const ResBlockW& w, int stats_slot, cudaStream_t s,
                    const NormFuse& nf, bool gn1_done = false, double* next_gn = nullptr) {
    const int C = cx.C, prec = h->cfg.precision;
    const RowSpace& rs = *cx.rs;
    double* st1 = h->gn_stats + static_cast<size_t>(stats_slot) * rs.n_utts * 64;
    double* st2 = st1 + static_cast<size_t>(rs.n_utts) * 64;
    auto conv3 = [&](const void* a, const void* wt, float* out, const float* bias, const float* residual,
                     bool mask, const NormFuse& f) {
        GemmCall c;
        c.precision = prec;
        c.a = a;
        c.a_rows = rs.rows;
        c.Cin = C;
        c.w = wt;
        c.N = C;
        c.taps = 3;
        c.out = out;
        c.out_fp32 = 1;
        c.ldc = C;
        c.n_store = C;
        c.bias = bias;
        c.residual = residual;
        c.ld_res = C;
        c.row_valid = mask ? rs.row_valid : nullptr;
        c.act = kActNone;
        c.ss_in = f.ss_in;
        c.ss_inv_dim = 1.f / static_cast<float>(C);
        c.ss_eps = 1e-6f;
        c.out16 = f.out16;
        c.ld16 = C;
        c.ss_out = f.ss_out;
        c.gn_stats = f.gn_stats;
        c.gn_row_utt = rs.row_utt;
        c.out16_scale = f.out16_scale;
        c.ss_in_scale = f.ss_in_scale;
        return launch_gemm(c, s);
    };
    // the conv GEMMs reduce the statistics of what they write when a chunk of their epilogue is a
    // group (C == 1024 on the CTA-pair kernel); otherwise a stand-alone pass over the activations does
    const bool fused_stats = gn_stats_in_gemm(rs.rows, C);
    {
        Stage t(h, "groupnorm_swish", s);
        if (!gn1_done) RUN(launch_groupnorm_stats(cx.x, rs, C, st1, s));
        RUN(launch_groupnorm_apply_swish(prec, cx.x, rs, C, st1, w.gn1_w, w.gn1_b, 1e-6f, cx.an, s));
    }
    {
        Stage t(h, "conv3_gemm", s);
        NormFuse f1;
        if (fused_stats) f1.gn_stats = st2;
        RUN(conv3(cx.an, w.w1, cx.hb, w.b1, nullptr, false, f1));
    }
    {
        Stage t(h, "groupnorm_swish", s);
        if (!fused_stats) RUN(launch_groupnorm_stats(cx.hb, rs, C, st2, s));
        RUN(launch_groupnorm_apply_swish(prec, cx.hb, rs, C, st2, w.gn2_w, w.gn2_b, 1e-6f, cx.an, s));
    }
    {
        Stage t(h, "conv3_gemm", s);
        NormFuse f2 = nf;
        if (fused_stats) f2.gn_stats = next_gn;
        RUN(conv3(cx.an, w.w2, cx.x, w.b2, cx.x, cx.mask_out, f2));
    }
    return 0;
}

int resnet_block(B200Codec* h, const ResBlockW& w, int stats_slot, cudaStream_t s,
                 const NormFuse& nf = NormFuse(), bool gn1_done = false, double* next_gn = nullptr) {
    ResCtx cx{&h->rs, h->C, h->x, h->hbuf, h->an, false};
    return resnet_block_ex(h, cx, w, stats_slot, s, nf, gn1_done, next_gn);
}

int tap_stage(B200Codec* h, const char* name, const void* src, int elem, int width, int ld, int rows, int space,
              cudaStream_t s);

// UpSamplerBlock.forward (upsampler.py:62-69): per stage, ConvTranspose1d as `stride` row-shifted convs
// (one per output phase, written with a row stride of `stride`), then a ResnetBlock; finally
// out_proj + swish. Input: LayerNorm output h->an (halo rows zero); output: h->u_fin.
int upsample_path(B200Codec* h, cudaStream_t s) {
    const int prec = h->cfg.precision;
    const void* cur = h->an;
    const RowSpace* rs_in = &h->rs;
    for (int i = 0; i < h->n_up; ++i) {
        const UpStageW& st = h->up[i];
        const RowSpace& rs_out = h->rs_up[i];
        {
            Stage t(h, "upsample_convT_gemm", s);
            for (int ph = 0; ph < st.stride; ++ph) {
                GemmCall c;
                c.precision = prec;
                c.a = cur;
                c.a_rows = rs_in->rows;
                c.Cin = st.Cin;
                c.w = st.phase[ph].w;
                c.N = st.Cout;
                c.taps = st.phase[ph].taps;
                c.tap_pad = st.phase[ph].pad;
                c.out = h->u_x[i] + static_cast<size_t>(ph) * st.Cout;  // out row = stride * m + ph
                c.out_fp32 = 1;
                c.ldc = st.stride * st.Cout;
                c.n_store = st.Cout;
                c.bias = st.bias;
                c.residual = nullptr;
                c.ld_res = 0;
                c.row_valid = nullptr;
                c.act = kActNone;
                RUN(launch_gemm(c, s));
            }
        }
        const std::string tag = std::to_string(i);
        if (tap_stage(h, ("up" + tag).c_str(), h->u_x[i], 4, st.Cout, st.Cout, rs_out.rows, i + 1, s)) return 1;
        NormFuse nf;
        nf.out16 = h->u_a16[i];  // operand copy of the block output: input of the next stage / out_proj
        ResCtx cx{&rs_out, st.Cout, h->u_x[i], h->u_h[i], h->u_an[i], true};
        if (resnet_block_ex(h, cx, st.res, 8 + 2 * i, s, nf)) return 1;
        if (tap_stage(h, ("res" + tag).c_str(), h->u_x[i], 4, st.Cout, st.Cout, rs_out.rows, i + 1, s)) return 1;
        cur = h->u_a16[i];
        rs_in = &rs_out;
    }
    {
        Stage t(h, "out_proj_gemm", s);
        GemmCall c;
        c.precision = prec;
        c.a = cur;
        c.a_rows = rs_in->rows;
        c.Cin = h->C >> h->n_up;
        c.w = h->w_out_proj;
        c.N = h->C;
        c.taps = 1;
        c.out = h->u_fin;
        c.out_fp32 = 0;
        c.ldc = h->C;
        c.n_store = h->C;
        c.bias = h->m("upsampler.out_proj.bias");
        c.residual = nullptr;
        c.ld_res = 0;
        c.row_valid = nullptr;
        c.act = kActSilu;  // nonlinearity(out_proj(x)) (upsampler.py:69)
        RUN(launch_gemm(c, s));
    }
    if (tap_stage(h, "upsampled", h->u_fin, static_cast<int>(operand_bytes(prec)), h->C, h->C, rs_in->rows, h->n_up, s))
        return 1;
    return 0;
}

// The fp32 residual stream x ([rows, 1024]) is re-read by every residual epilogue and GroupNorm after
// ~150 MB of other activations have gone through L2; pin it with a persisting access-policy window
// on the caller's stream for the duration of the decode (restored afterwards).
void set_l2_window(B200Codec* h, cudaStream_t s, bool on, cudaStreamAttrValue* saved) {
    if (h->l2_persist_bytes == 0) return;
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (on) {
        // the caller's own window (if any) comes back after the decode
        if (cudaStreamGetAttribute(s, cudaStreamAttributeAccessPolicyWindow, saved) != cudaSuccess) {
            (void)cudaGetLastError();
            std::memset(saved, 0, sizeof(*saved));
        }
        size_t bytes = static_cast<size_t>(h->rs.rows) * h->C * sizeof(float);
        if (bytes > h->l2_window_max) bytes = h->l2_window_max;
        attr.accessPolicyWindow.base_ptr = h->x;
        attr.accessPolicyWindow.num_bytes = bytes;
        const double ratio = static_cast<double>(h->l2_persist_bytes) / static_cast<double>(bytes);
        attr.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : static_cast<float>(ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr = *saved;
    }
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
}

int forward_impl(B200Codec* h, const void* ids_dev, int id_type, float* wav_dev, cudaStream_t s);

// debug tap: keep a device copy of a stage tensor (whole padded row space) for b200codec_read_stage
int tap_stage(B200Codec* h, const char* name, const void* src, int elem, int width, int ld, int rows, int space,
              cudaStream_t s) {
    if (!h->taps_on) return 0;
    B200Codec::StageTap& t = h->taps[name];
    const size_t bytes = static_cast<size_t>(rows) * ld * elem;
    if (t.buf.ensure(bytes)) return 1;
    B200_CUDA_OK(cudaMemcpyAsync(t.buf.p, src, bytes, cudaMemcpyDeviceToDevice, s));
    t.width = width;
    t.ld = ld;
    t.elem = elem;
    t.space = space;
    t.valid = true;
    return 0;
}

int forward(B200Codec* h, const void* ids_dev, int id_type, float* wav_dev, cudaStream_t s) {
    cudaStreamAttrValue saved;
    std::memset(&saved, 0, sizeof(saved));
    set_l2_window(h, s, true, &saved);
    const int rc = forward_impl(h, ids_dev, id_type, wav_dev, s);
    set_l2_window(h, s, false, &saved);
    return rc;
}

int forward_impl(B200Codec* h, const void* ids_dev, int id_type, float* wav_dev, cudaStream_t s) {
    const int C = h->C, prec = h->cfg.precision;
    const RowSpace& rs = h->rs;
    B200_CUDA_OK(cudaMemsetAsync(h->gn_stats, 0, h->stats_zero_bytes, s));
    // GroupNorm statistics slot k (k = 2 * block + {0: norm1, 1: norm2}) of utterance u, group g:
    // gn_stats[((k * n_utts + u) * 32 + g) * 2 + {sum, sumsq}]
    const bool gn_fused = gn_stats_in_gemm(rs.rows, C);
    auto gn_slot = [&](int k) { return gn_fused ? h->gn_stats + static_cast<size_t>(k) * rs.n_utts * 64 : nullptr; };
    bool gn0_done = false;  // prior_net.0's first GroupNorm statistics already reduced
    // project_out (8 -> 2048, decoder.py:77), fc_post_a (2048 -> 1024, decoder.py:79) and the backbone's
    // embed Conv1d (k = 7, decoder_modules.py:340,392) are linear maps back to back: the conv output is
    // a linear function of the seven neighbouring codes, with coefficients folded in fp64 at load time.
    // No 2048-wide intermediate, no K = 2048 / K = 7168 GEMM.
    if (g_frontend_fold == 1) {
        {
            Stage t(h, "fsq_lookup", s);
            RUN(launch_fsq_im2col(ids_dev, id_type, rs.row_tok, rs.rows, prec, h->xc, h->err_flag_dev, s));
        }
        {
            // x = A [hi | lo]^T on the tensor cores (K = 128); the epilogue also reduces the first
            // GroupNorm's statistics
            Stage t(h, "frontend_gemm", s);
            NormFuse nf0;
            nf0.gn_stats = gn_slot(0);
            gn0_done = gn_fused;
            RUN(gemm(h, h->xc, 128, h->w_front, C, 1, h->x, true, C, C, nullptr, nullptr, kActNone, false, s, nf0));
        }
    } else if (g_frontend_fold == 2) {
        Stage t(h, "fsq_lookup", s);
        RUN(launch_fsq_frontend(ids_dev, id_type, rs.row_tok, rs.rows, h->m_fold, h->cb_fold,
                                h->m("decoder.backbone.embed.bias"), C, h->x, h->err_flag_dev, s));
    } else {
        {
            // fc_post_a(project_out(code)) = W_pre code + b_pre: the same digit-unpack / 8-term lookup
            // kernel as the stand-alone K1. Halo rows are written as zeros: this is the conv7 operand.
            Stage t(h, "fsq_lookup", s);
            RUN(launch_fsq_lookup(ids_dev, id_type, rs.row_tok, rs.rows, h->w_pre, h->b_pre, C, h->xc, C, prec,
                                  h->err_flag_dev, s));
        }
        {
            Stage t(h, "embed_conv7_gemm", s);
            RUN(gemm(h, h->xc, C, h->w_embed, C, 7, h->x, true, C, C,
                     h->m("decoder.backbone.embed.bias"), nullptr, kActNone, false, s));
        }
    }
    // RMSNorm fusion (CTA-pair GEMM): the GEMM that produces the residual stream x also emits a 16-bit
    // copy of x and per-row sum-of-squares partials; the norm weight lives in the columns of c_attn /
    // fc1 and rstd[row] is applied in their epilogues. With fp16 operands the copy is x / 64 (exact;
    // un-normalised activations up to 4e6 stay inside fp16's range) and the consumers' row scale
    // carries the factor 64 back -- RMSNorm does not care about the scale of its input.
    const bool fuse_rms = g_fuse_rmsnorm && gemm_uses_cta_pairs(rs.rows, C);
    NormFuse produce, consume;
    if (fuse_rms) {
        produce.out16 = h->xb;
        produce.ss_out = h->ss;
        consume.ss_in = h->ss;
        if (prec == kPrecFp16) {
            produce.out16_scale = 1.f / 64.f;
            consume.ss_in_scale = 64.f;
        }
    }
    if (tap_stage(h, "embed", h->x, 4, C, C, rs.rows, 0, s)) return 1;  // backbone.embed output
    if (resnet_block(h, h->res[0], 0, s, NormFuse(), gn0_done, gn_slot(2))) return 1;
    if (resnet_block(h, h->res[1], 2, s, produce, gn_fused, nullptr)) return 1;
    if (tap_stage(h, "prior_net", h->x, 4, C, C, rs.rows, 0, s)) return 1;
    // Transformer blocks (decoder_modules.py:311-314). With the fused RMSNorm every linear of a block reads,
    // per 256-row block, only what the previous linear wrote in that block, so c_proj -> fc1 -> fc2 -> the
    // next block's c_attn run as ONE persistent launch (a GEMM chain); the last block's fc2 also reduces
    // GroupNorm statistics and stays a launch of its own.
    const bool chain = fuse_rms && g_gemm_chain != 0;
    const size_t num_m = (static_cast<size_t>(rs.rows) + 255) / 256;
    auto qkv_call = [&](int l) {
        return gemm_call(h, fuse_rms ? h->xb : h->an, C, h->layers[l].qkv, 3 * C, 1, h->qkv, false, 3 * C, 3 * C, nullptr,
                         nullptr, kActNone, false, consume);
    };
    for (int l = 0; l < h->L; ++l) {
        const LayerW& w = h->layers[l];
        const bool last = l + 1 == h->L;
        if (!fuse_rms) {
            Stage t(h, "rmsnorm", s);
            RUN(launch_rmsnorm(prec, h->x, nullptr, rs.rows, C, 1e-6f, h->an, s));
        }
        if (!chain || l == 0) {
            Stage t(h, "qkv_gemm", s);
            RUN(launch_gemm(qkv_call(l), s));
        }
        if (h->cur_stream != nullptr) {
            // cached streaming: this layer's keys / values of the new rows join the ring, then every row of the
            // push attends to the ring (the cached tokens as they were computed when new + the new ones)
            Stage t(h, "attention", s);
            B200Stream* st = h->cur_stream;
            void* ring = static_cast<uint8_t*>(st->kv_slab) + static_cast<size_t>(l) * st->layer_stride;
            RUN(launch_kv_scatter(prec, h->qkv, rs.utt_row0, st->n_streams, st->overlap, st->new_tokens, C, ring, st->cap,
                                  st->wpos, s));
            RUN(launch_attention_tc05_ex(prec, h->qkv, rs.rows, 3 * C, ring, st->n_streams * st->cap, 2 * C, 0, C,
                                         st->work_dev, st->n_work, h->H, h->y, s));
        } else {
            Stage t(h, "attention", s);
            RUN(run_attention(prec, h->qkv, rs, h->H, h->y, s));
        }
        const GemmCall proj = gemm_call(h, h->y, C, w.proj, C, 1, h->x, true, C, C, nullptr, h->x, kActNone, false, produce);
        const GemmCall fc1 = gemm_call(h, fuse_rms ? h->xb : h->an, C, w.fc1, 4 * C, 1, h->f, false, 4 * C, 4 * C, nullptr,
                                       nullptr, kActSilu, false, consume);
        NormFuse f2 = produce;
        if (last) {
            f2 = NormFuse();
            f2.gn_stats = gn_slot(4);  // post_net.0's first GroupNorm reads this x
        }
        const GemmCall fc2 = gemm_call(h, h->f, 4 * C, w.fc2, C, 1, h->x, true, C, C, nullptr, h->x, kActNone, false, f2);
        if (chain) {
            GemmCall calls[4] = {proj, fc1, fc2, fc2};
            int n = 2;
            if (!last) {
                calls[3] = qkv_call(l + 1);
                n = 4;
            }
            {
                Stage t(h, "chain_gemm", s);
                RUN(launch_gemm_chain(calls, n, h->chain_ctr + static_cast<size_t>(l) * 4 * num_m, s));
            }
            if (last) {
                Stage t(h, "fc2_gemm", s);
                RUN(launch_gemm(fc2, s));
            }
        } else {
            {
                Stage t(h, "proj_gemm", s);
                RUN(launch_gemm(proj, s));
            }
            if (!fuse_rms) {
                Stage t(h, "rmsnorm", s);
                RUN(launch_rmsnorm(prec, h->x, nullptr, rs.rows, C, 1e-6f, h->an, s));
            }
            {
                Stage t(h, "fc1_gemm", s);
                RUN(launch_gemm(fc1, s));
            }
            {
                Stage t(h, "fc2_gemm", s);
                RUN(launch_gemm(fc2, s));
            }
        }
        if (l == 0 && tap_stage(h, "tblock0", h->x, 4, C, C, rs.rows, 0, s)) return 1;
    }
    if (tap_stage(h, "transformers", h->x, 4, C, C, rs.rows, 0, s)) return 1;
    if (resnet_block(h, h->res[2], 4, s, NormFuse(), gn_fused && h->L > 0, gn_slot(6))) return 1;
    if (resnet_block(h, h->res[3], 6, s, NormFuse(), gn_fused, nullptr)) return 1;
    {
        Stage t(h, "layernorm", s);
        // with an upsampler the LayerNorm output feeds a transposed conv: halo rows must be zero
        RUN(launch_layernorm(prec, h->x, h->m("decoder.backbone.final_layer_norm.weight"),
                             h->m("decoder.backbone.final_layer_norm.bias"), rs.rows, C, 1e-6f,
                             h->an, s, h->n_up > 0 ? rs.row_valid : nullptr));
    }
    // VocosBackbone.forward output (final_layer_norm), stored in the operand dtype
    if (tap_stage(h, "backbone", h->an, static_cast<int>(operand_bytes(prec)), C, C, rs.rows, 0, s)) return 1;
    const RowSpace& rs_last = h->n_up > 0 ? h->rs_up[h->n_up - 1] : rs;
    const void* head_in = h->an;
    if (h->n_up > 0) {
        if (upsample_path(h, s)) return 1;
        head_in = h->u_fin;
    }
    {
        Stage t(h, "head_gemm", s);
        GemmCall c;
        c.precision = prec;
        c.a = head_in;
        c.a_rows = rs_last.rows;
        c.Cin = C;
        c.w = h->w_head;
        c.N = h->n_fft + 2;
        c.taps = 1;
        c.out = h->ho;
        c.out_fp32 = 1;
        c.ldc = h->head_ld;
        c.n_store = h->head_ld;
        c.bias = h->head_bias_pad;
        c.residual = nullptr;
        c.ld_res = 0;
        c.row_valid = nullptr;
        c.act = kActNone;
        RUN(launch_gemm(c, s));
    }
    if (tap_stage(h, "head_linear", h->ho, 4, h->n_fft + 2, h->head_ld, rs_last.rows, h->n_up, s)) return 1;
    {
        Stage t(h, "istft", s);
        IstftTables tab;
        tab.twiddle = h->twiddle;
        tab.window = h->m("decoder.head.istft.window");
        RUN(launch_istft(h->ho, h->head_ld, rs_last, tab, h->hop, wav_dev, s));
    }
    return 0;
}

int check_ids_host(const void* ids, int id_type, int64_t n) {
    if (id_type == B200CODEC_IDS_I64) {
        const int64_t* p = static_cast<const int64_t*>(ids);
        for (int64_t i = 0; i < n; ++i)
            B200_CHECK(p[i] >= 0 && p[i] <= 65535, "speech id %lld at position %lld is outside "
                       "[0, 65535]", (long long)p[i], (long long)i);
    } else {
        const int32_t* p = static_cast<const int32_t*>(ids);
        for (int64_t i = 0; i < n; ++i)
            B200_CHECK(p[i] >= 0 && p[i] <= 65535, "speech id %d at position %lld is outside "
                       "[0, 65535]", p[i], (long long)i);
    }
    return 0;
}

int prepare(B200Codec* h, const int32_t* seqlens, int n_utts, cudaStream_t s) {
    B200_CHECK(h != nullptr, "null handle");
    B200_CHECK(h->finalized, "decode called before b200codec_finalize_weights");
    B200_CHECK(seqlens != nullptr, "null seqlens");
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    if (build_plan(h, seqlens, n_utts, kGap, s)) return 1;
    if (ensure_workspace(h, h->rs.rows)) return 1;
    return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* b200codec_last_error(void) { return g_err; }

int b200codec_create(const B200CodecConfig* cfg, B200Codec** out) {
    g_err[0] = 0;
    B200_CHECK(cfg != nullptr && out != nullptr, "b200codec_create: null argument");
    B200_CHECK(cfg->abi_version == B200CODEC_ABI_VERSION, "ABI version mismatch: caller %d, library %d",
               cfg->abi_version, B200CODEC_ABI_VERSION);
    B200_CHECK(cfg->n_upsample >= 0 && cfg->n_upsample <= 3,
               "upsample_factors: %d stages are not supported (0 .. 3; channels halve per stage and the "
               "kernels are instantiated for 512, 256 and 128)", cfg->n_upsample);
    int total_up = 1;
    for (int i = 0; i < cfg->n_upsample; ++i) {
        const int u = cfg->upsample_factors[i], k = cfg->kernel_sizes[i];
        B200_CHECK(u >= 1 && u <= 8 && k >= u && k <= 16 && (k - u) % 2 == 0,
                   "unsupported upsampler stage %d: factor %d kernel %d (need factor <= kernel, kernel - factor even)",
                   i, u, k);
        total_up *= u;
    }
    B200_CHECK(cfg->hop_length > 0 && cfg->sample_rate / cfg->hop_length / total_up == 50,
               "Current hop length %d and upsample factors (product %d) do not match the target sample "
               "rate %d.", cfg->hop_length, total_up, cfg->sample_rate);  // decoder.py:31-37
    B200_CHECK(cfg->hop_length == 320 || cfg->hop_length == 240 || cfg->hop_length == 160 || cfg->hop_length == 80,
               "hop_length %d is not instantiated (320, 240, 160, 80: n_fft = 4 hop = 64 x {20, 15, 10, 5})", cfg->hop_length);
    B200_CHECK(cfg->precision == B200CODEC_BF16 || cfg->precision == B200CODEC_FP16,
               "precision %d is not available (bf16 = 0, fp16 = 1)", cfg->precision);
    B200_CHECK(cfg->hidden_dim == 1024 && cfg->heads == 16 && cfg->vq_dim == 2048 &&
               cfg->depth >= 1 && cfg->depth <= 64,
               "unsupported architecture (hidden %d heads %d vq %d depth %d)", cfg->hidden_dim,
               cfg->heads, cfg->vq_dim, cfg->depth);
    int ndev = 0;
    B200_CUDA_OK(cudaGetDeviceCount(&ndev));
    B200_CHECK(cfg->device >= 0 && cfg->device < ndev, "CUDA device %d not present (%d devices); "
               "this library has no CPU fallback", cfg->device, ndev);
    B200_CUDA_OK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    B200_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    B200_CHECK(prop.major == 10, "device %d is sm_%d%d; b200codec kernels are built for sm_100a only",
               cfg->device, prop.major, prop.minor);

    B200Codec* h = new B200Codec();
    h->cfg = *cfg;
    h->C = cfg->hidden_dim;
    h->H = cfg->heads;
    h->L = cfg->depth;
    h->V = cfg->vq_dim;
    h->hop = cfg->hop_length;
    h->n_fft = 4 * cfg->hop_length;
    h->n_bins = h->n_fft / 2 + 1;
    // head.out columns padded to a multiple of 256 (1282 -> 1536, 642 -> 768): the padded columns cost
    // ~17 % extra MMA work but let the ragged GEMM run on the CTA-pair kernel
    h->head_ld = (h->n_fft + 2 + 255) / 256 * 256;
    h->n_up = cfg->n_upsample;
    h->total_up = total_up;
    h->gn_slots = 8 + 2 * h->n_up;
    for (int i = 0; i < h->n_up; ++i) {
        h->up_f[i] = cfg->upsample_factors[i];
        h->up_k[i] = cfg->kernel_sizes[i];
        h->up[i].Cin = h->C >> i;
        h->up[i].Cout = h->C >> (i + 1);
        h->up[i].stride = h->up_f[i];
        h->up[i].k = h->up_k[i];
    }
    build_specs(h);
    h->master.assign(h->specs.size(), nullptr);
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->err_flag_host), sizeof(int),
                      cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_flag_dev), h->err_flag_host, 0) !=
            cudaSuccess) {
        set_error("cannot allocate the mapped error flag");
        delete h;
        return 1;
    }
    *h->err_flag_host = 0;
    // persisting L2 carve-out for the residual stream (best effort)
    if (prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        size_t want = static_cast<size_t>(prop.persistingL2CacheMaxSize);
        const size_t cap = static_cast<size_t>(prop.l2CacheSize) / 2;  // leave half of L2 to everything else
        if (want > cap) want = cap;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            h->l2_persist_bytes = want;
            h->l2_window_max = static_cast<size_t>(prop.accessPolicyMaxWindowSize);
        } else {
            cudaGetLastError();
        }
    }
    *out = h;
    return 0;
}

void b200codec_destroy(B200Codec* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (float* p : h->master)
        if (p) cudaFree(p);
    h->wbuf.release();
    h->ws.release();
    h->plan_dev.release();
    for (int i = 0; i < kMaxUp; ++i) h->plan_up_dev[i].release();
    h->io_ids.release();
    h->io_wav.release();
    for (auto& st : h->plan_stage) {
        if (st.host) cudaFreeHost(st.host);
        if (st.copied) cudaEventDestroy(st.copied);
    }
    if (h->head_bias_pad) cudaFree(h->head_bias_pad);
    if (h->twiddle) cudaFree(h->twiddle);
    if (h->m_fold) cudaFree(h->m_fold);
    if (h->cb_fold) cudaFree(h->cb_fold);
    if (h->w_pre) cudaFree(h->w_pre);
    if (h->b_pre) cudaFree(h->b_pre);
    if (h->rope_cos) cudaFree(h->rope_cos);
    if (h->rope_sin) cudaFree(h->rope_sin);
    if (h->gn_stats) cudaFree(h->gn_stats);
    if (h->err_flag_host) cudaFreeHost(h->err_flag_host);
    delete h;
}

int b200codec_num_tensors(const B200Codec* h) { return h ? static_cast<int>(h->specs.size()) : 0; }

const char* b200codec_tensor_key(const B200Codec* h, int i) {
    if (!h || i < 0 || i >= static_cast<int>(h->specs.size())) return nullptr;
    return h->specs[i].key.c_str();
}

int b200codec_tensor_shape(const B200Codec* h, int i, int64_t shape_out[4]) {
    if (!h || i < 0 || i >= static_cast<int>(h->specs.size())) return -1;
    const auto& s = h->specs[i].shape;
    for (size_t d = 0; d < s.size() && d < 4; ++d) shape_out[d] = s[d];
    return static_cast<int>(s.size());
}

int b200codec_load_tensor(B200Codec* h, const char* key, const void* host_ptr, int dtype,
                          const int64_t* shape, int ndim) {
    B200_CHECK(h && key && host_ptr && shape, "b200codec_load_tensor: null argument");
    auto it = h->index.find(key);
    B200_CHECK(it != h->index.end(), "Unexpected key(s) in state_dict: \"%s\"", key);
    const TensorSpec& sp = h->specs[it->second];
    bool ok = static_cast<int>(sp.shape.size()) == ndim;
    for (int d = 0; ok && d < ndim; ++d) ok = sp.shape[d] == shape[d];
    B200_CHECK(ok, "size mismatch for %s: checkpoint tensor does not match the model shape", key);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    const size_t n = sp.numel();
    std::vector<float> tmp;
    const float* src = nullptr;
    if (dtype == B200CODEC_F32) {
        src = static_cast<const float*>(host_ptr);
    } else {
        tmp.resize(n);
        if (dtype == B200CODEC_F64) {
            const double* p = static_cast<const double*>(host_ptr);
            for (size_t i = 0; i < n; ++i) tmp[i] = static_cast<float>(p[i]);
        } else if (dtype == B200CODEC_F16) {
            const __half* p = static_cast<const __half*>(host_ptr);
            for (size_t i = 0; i < n; ++i) tmp[i] = __half2float(p[i]);
        } else if (dtype == B200CODEC_BF16_T) {
            const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(host_ptr);
            for (size_t i = 0; i < n; ++i) tmp[i] = __bfloat162float(p[i]);
        } else {
            set_error("b200codec_load_tensor: unsupported dtype %d for %s", dtype, key);
            return 1;
        }
        src = tmp.data();
    }
    float*& dst = h->master[it->second];
    if (dst == nullptr) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dst), n * sizeof(float)));
    B200_CUDA_OK(cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    h->finalized = false;
    return 0;
}

int b200codec_read_tensor(const B200Codec* h, const char* key, float* host_out, size_t n_elems) {
    B200_CHECK(h && key && host_out, "b200codec_read_tensor: null argument");
    auto it = h->index.find(key);
    B200_CHECK(it != h->index.end(), "unknown tensor \"%s\"", key);
    const TensorSpec& sp = h->specs[it->second];
    B200_CHECK(sp.numel() == n_elems, "read_tensor %s: expected %zu elements, got %zu", key,
               sp.numel(), n_elems);
    B200_CHECK(h->master[it->second] != nullptr, "tensor \"%s\" has not been loaded", key);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    B200_CUDA_OK(cudaMemcpy(host_out, h->master[it->second], n_elems * sizeof(float),
                            cudaMemcpyDeviceToHost));
    return 0;
}

int b200codec_finalize_weights(B200Codec* h, void* stream) {
    B200_CHECK(h != nullptr, "null handle");
    if (h->finalized) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    std::string missing;
    for (size_t i = 0; i < h->specs.size(); ++i)
        if (h->master[i] == nullptr) {
            if (missing.size() < 600) missing += (missing.empty() ? "\"" : ", \"") + h->specs[i].key + "\"";
        }
    B200_CHECK(missing.empty(), "Missing key(s) in state_dict: %s", missing.c_str());

    const int C = h->C, V = h->V, L = h->L, prec = h->cfg.precision;
    const size_t es = operand_bytes(prec);
    const size_t n_head = static_cast<size_t>(h->n_fft + 2) * C;
    size_t elems = static_cast<size_t>(C) * 128 + static_cast<size_t>(C) * C * 7 +
                   8 * static_cast<size_t>(C) * C * 3 +
                   static_cast<size_t>(L) * (3ull * C * C + 1ull * C * C + 8ull * C * C) + n_head;
    for (int i = 0; i < h->n_up; ++i) {
        const size_t ci = h->up[i].Cin, co = h->up[i].Cout;
        elems += ci * co * (h->up[i].k + h->up[i].stride * 2) + 2 * co * co * 3 + 4096;
    }
    if (h->n_up > 0) elems += static_cast<size_t>(C) * (C >> h->n_up) + 4096;
    if (h->wbuf.ensure(elems * es + 64 * 1024)) return 1;
    uint8_t* wp = h->wbuf.as<uint8_t>();
    auto take = [&](size_t n) {
        void* r = wp;
        wp += (n * es + 255) & ~static_cast<size_t>(255);
        return r;
    };
    // RoPE tables: torchtune.modules.RotaryPositionalEmbeddings(dim=64, base=10000) evaluated
    // at "position" = head index (decoder_modules.py:276-281 passes [b, h, t, d]).
    {
        const int half = 32;
        std::vector<float> c(h->H * half), sn(h->H * half);
        for (int hd = 0; hd < h->H; ++hd)
            for (int j = 0; j < half; ++j) {
                const float theta = 1.0f / powf(10000.0f, static_cast<float>(2 * j) / 64.0f);
                const float ang = static_cast<float>(hd) * theta;
                c[hd * half + j] = cosf(ang);
                sn[hd * half + j] = sinf(ang);
            }
        if (!h->rope_cos) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->rope_cos), c.size() * 4));
        if (!h->rope_sin) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->rope_sin), c.size() * 4));
        B200_CUDA_OK(cudaMemcpy(h->rope_cos, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
        B200_CUDA_OK(cudaMemcpy(h->rope_sin, sn.data(), c.size() * 4, cudaMemcpyHostToDevice));
    }
    float* scratch = nullptr;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&scratch), 3ull * C * C * sizeof(float)));

    {
        // W_pre = W_fc W_out, b_pre = W_fc b_out + b_fc (fp64 on the host, stored fp32)
        std::vector<float> wfc(static_cast<size_t>(C) * V), bfc(C), wo(static_cast<size_t>(V) * 8), bo(V);
        B200_CUDA_OK(cudaMemcpy(wfc.data(), h->m("fc_post_a.weight"), wfc.size() * 4, cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(bfc.data(), h->m("fc_post_a.bias"), bfc.size() * 4, cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(wo.data(), h->m("decoder.quantizer.project_out.weight"), wo.size() * 4,
                                cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(bo.data(), h->m("decoder.quantizer.project_out.bias"), bo.size() * 4,
                                cudaMemcpyDeviceToHost));
        std::vector<float> wp(static_cast<size_t>(C) * 8), bp(C);
        std::vector<double> wp64(static_cast<size_t>(C) * 8), bp64(C);
        for (int c = 0; c < C; ++c) {
            double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            double b = bfc[c];
            const float* wr = &wfc[static_cast<size_t>(c) * V];
            for (int v = 0; v < V; ++v) {
                const double wv = wr[v];
                for (int d = 0; d < 8; ++d) acc[d] += wv * wo[static_cast<size_t>(v) * 8 + d];
                b += wv * bo[v];
            }
            for (int d = 0; d < 8; ++d) {
                wp[static_cast<size_t>(c) * 8 + d] = static_cast<float>(acc[d]);
                wp64[static_cast<size_t>(c) * 8 + d] = acc[d];
            }
            bp[c] = static_cast<float>(b);
            bp64[c] = b;
        }
        // second fold: M[tap] = W_e[:, :, tap] W_pre, cb[tap] = W_e[:, :, tap] b_pre
        std::vector<float> we(static_cast<size_t>(C) * C * 7);
        B200_CUDA_OK(cudaMemcpy(we.data(), h->m("decoder.backbone.embed.weight"), we.size() * 4,
                                cudaMemcpyDeviceToHost));
        std::vector<float> mf(static_cast<size_t>(7) * C * 8), cbf(static_cast<size_t>(7) * C);
        for (int c = 0; c < C; ++c) {
            double acc[7][9];
            for (int t = 0; t < 7; ++t)
                for (int d = 0; d < 9; ++d) acc[t][d] = 0.0;
            for (int k = 0; k < C; ++k) {
                const float* wr = &we[(static_cast<size_t>(c) * C + k) * 7];
                const double* pk = &wp64[static_cast<size_t>(k) * 8];
                for (int t = 0; t < 7; ++t) {
                    const double wv = wr[t];
                    for (int d = 0; d < 8; ++d) acc[t][d] += wv * pk[d];
                    acc[t][8] += wv * bp64[k];
                }
            }
            for (int t = 0; t < 7; ++t) {
                for (int d = 0; d < 8; ++d)
                    mf[(static_cast<size_t>(t) * C + c) * 8 + d] = static_cast<float>(acc[t][d]);
                cbf[static_cast<size_t>(t) * C + c] = static_cast<float>(acc[t][8]);
            }
        }
        {
            // GEMM form: B[c][8 * tap + d] = M, B[c][56 + tap] = cb, B[c][63] = embed bias; columns
            // 64..127 hold the part the 16-bit rounding of columns 0..63 lost (hi + lo ~ fp32)
            std::vector<float> be(C);
            B200_CUDA_OK(cudaMemcpy(be.data(), h->m("decoder.backbone.embed.bias"), be.size() * 4,
                                    cudaMemcpyDeviceToHost));
            std::vector<uint16_t> wf(static_cast<size_t>(C) * 128);
            auto split = [&](float v, uint16_t& hi, uint16_t& lo) {
                if (prec == kPrecBf16) {
                    const __nv_bfloat16 a = __float2bfloat16_rn(v);
                    const __nv_bfloat16 b = __float2bfloat16_rn(v - __bfloat162float(a));
                    hi = *reinterpret_cast<const uint16_t*>(&a);
                    lo = *reinterpret_cast<const uint16_t*>(&b);
                } else {
                    const __half a = __float2half_rn(v);
                    const __half b = __float2half_rn(v - __half2float(a));
                    hi = *reinterpret_cast<const uint16_t*>(&a);
                    lo = *reinterpret_cast<const uint16_t*>(&b);
                }
            };
            for (int c = 0; c < C; ++c)
                for (int col = 0; col < 64; ++col) {
                    float v;
                    if (col < 56) v = mf[(static_cast<size_t>(col >> 3) * C + c) * 8 + (col & 7)];
                    else if (col < 63) v = cbf[static_cast<size_t>(col - 56) * C + c];
                    else v = be[c];
                    split(v, wf[static_cast<size_t>(c) * 128 + col], wf[static_cast<size_t>(c) * 128 + 64 + col]);
                }
            h->w_front = take(static_cast<size_t>(C) * 128);
            B200_CUDA_OK(cudaMemcpy(h->w_front, wf.data(), wf.size() * 2, cudaMemcpyHostToDevice));
        }
        if (!h->m_fold) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->m_fold), mf.size() * 4));
        if (!h->cb_fold) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->cb_fold), cbf.size() * 4));
        B200_CUDA_OK(cudaMemcpy(h->m_fold, mf.data(), mf.size() * 4, cudaMemcpyHostToDevice));
        B200_CUDA_OK(cudaMemcpy(h->cb_fold, cbf.data(), cbf.size() * 4, cudaMemcpyHostToDevice));
        if (!h->w_pre) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->w_pre), wp.size() * 4));
        if (!h->b_pre) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->b_pre), bp.size() * 4));
        B200_CUDA_OK(cudaMemcpy(h->w_pre, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
        B200_CUDA_OK(cudaMemcpy(h->b_pre, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    }
    h->w_embed = take(static_cast<size_t>(C) * C * 7);
    if (launch_repack_weight(prec, h->m("decoder.backbone.embed.weight"), h->w_embed, C, C, 7, s)) return 1;
    const char* nets[4] = {"decoder.backbone.prior_net.0.", "decoder.backbone.prior_net.1.",
                           "decoder.backbone.post_net.0.", "decoder.backbone.post_net.1."};
    for (int i = 0; i < 4; ++i) {
        const std::string p = nets[i];
        ResBlockW& r = h->res[i];
        r.gn1_w = h->m(p + "norm1.weight");
        r.gn1_b = h->m(p + "norm1.bias");
        r.b1 = h->m(p + "conv1.bias");
        r.gn2_w = h->m(p + "norm2.weight");
        r.gn2_b = h->m(p + "norm2.bias");
        r.b2 = h->m(p + "conv2.bias");
        r.w1 = take(static_cast<size_t>(C) * C * 3);
        r.w2 = take(static_cast<size_t>(C) * C * 3);
        if (launch_repack_weight(prec, h->m(p + "conv1.weight"), r.w1, C, C, 3, s)) return 1;
        if (launch_repack_weight(prec, h->m(p + "conv2.weight"), r.w2, C, C, 3, s)) return 1;
    }
    for (int l = 0; l < L; ++l) {
        const std::string p = "decoder.backbone.transformers." + std::to_string(l) + ".";
        LayerW& w = h->layers[l];
        w.att_norm = h->m(p + "att_norm.weight");
        w.ffn_norm = h->m(p + "ffn_norm.weight");
        w.qkv = take(3ull * C * C);
        w.proj = take(1ull * C * C);
        w.fc1 = take(4ull * C * C);
        w.fc2 = take(4ull * C * C);
        // fold the head-indexed rotary embedding into the q and k rows of c_attn (fp32), then cast
        B200_CUDA_OK(cudaMemcpyAsync(scratch, h->m(p + "att.c_attn.weight"), 3ull * C * C * 4,
                                     cudaMemcpyDeviceToDevice, s));
        if (launch_fold_rope(scratch, h->H, C / h->H, C, h->rope_cos, h->rope_sin, s)) return 1;
        // RMSNorm weights are folded into the columns of the consuming Linear:
        // (x * rstd * g) W^T == rstd * (x (W diag(g))^T)   (decoder_modules.py:233-236, 312-313)
        if (launch_repack_weight(prec, scratch, w.qkv, 3 * C, C, 1, s, w.att_norm)) return 1;
        if (launch_repack_weight(prec, h->m(p + "att.c_proj.weight"), w.proj, C, C, 1, s)) return 1;
        if (launch_repack_weight(prec, h->m(p + "mlp.fc1.weight"), w.fc1, 4 * C, C, 1, s, w.ffn_norm)) return 1;
        if (launch_repack_weight(prec, h->m(p + "mlp.fc2.weight"), w.fc2, C, 4 * C, 1, s)) return 1;
    }
    // ---- upsampler (48 kHz variant): weight-norm fold + one K-major slab set per output phase ----
    // ConvTranspose1d(stride u, kernel k, padding p = (k - u) / 2): out[u q + ph] = sum_e x[q + e] W[:, :, ph + p - u e]
    // over the e with 0 <= ph + p - u e < k, i.e. an ordinary conv over consecutive input rows per phase.
    float* wn_scale = nullptr;
    if (h->n_up > 0) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&wn_scale), C * sizeof(float)));
    for (int i = 0; i < h->n_up; ++i) {
        UpStageW& st = h->up[i];
        const std::string pl = "upsampler.upsample_layers." + std::to_string(i) + ".";
        const std::string pr = "upsampler.resnet_blocks." + std::to_string(i) + ".";
        const int u = st.stride, k = st.k, pad = (k - u) / 2;
        st.bias = h->m(pl + "bias");
        if (launch_weightnorm_scale(h->m(pl + "weight_g"), h->m(pl + "weight_v"), st.Cin, st.Cout * k, wn_scale, s)) return 1;
        for (int ph = 0; ph < u; ++ph) {
            int e_min = 1 << 20, e_max = -(1 << 20);
            for (int e = -16; e <= 16; ++e) {
                const int tap = ph + pad - u * e;
                if (tap >= 0 && tap < k) {
                    e_min = e < e_min ? e : e_min;
                    e_max = e > e_max ? e : e_max;
                }
            }
            B200_CHECK(e_min <= e_max && e_min <= 0 && e_max - e_min + 1 <= 8, "upsampler stage %d phase %d has no taps", i, ph);
            int tap_ids[8] = {0};
            const int taps = e_max - e_min + 1;
            for (int t = 0; t < taps; ++t) tap_ids[t] = ph + pad - u * (e_min + t);
            st.phase[ph].taps = taps;
            st.phase[ph].pad = -e_min;
            st.phase[ph].w = take(static_cast<size_t>(st.Cout) * taps * st.Cin);
            if (launch_repack_convT_phase(prec, h->m(pl + "weight_v"), wn_scale, st.phase[ph].w, st.Cin, st.Cout, k,
                                          taps, tap_ids, s))
                return 1;
        }
        ResBlockW& r = st.res;
        r.gn1_w = h->m(pr + "norm1.weight");
        r.gn1_b = h->m(pr + "norm1.bias");
        r.b1 = h->m(pr + "conv1.bias");
        r.gn2_w = h->m(pr + "norm2.weight");
        r.gn2_b = h->m(pr + "norm2.bias");
        r.b2 = h->m(pr + "conv2.bias");
        r.w1 = take(static_cast<size_t>(st.Cout) * st.Cout * 3);
        r.w2 = take(static_cast<size_t>(st.Cout) * st.Cout * 3);
        if (launch_repack_weight(prec, h->m(pr + "conv1.weight"), r.w1, st.Cout, st.Cout, 3, s)) return 1;
        if (launch_repack_weight(prec, h->m(pr + "conv2.weight"), r.w2, st.Cout, st.Cout, 3, s)) return 1;
    }
    if (h->n_up > 0) {
        h->w_out_proj = take(static_cast<size_t>(C) * (C >> h->n_up));
        if (launch_repack_weight(prec, h->m("upsampler.out_proj.weight"), h->w_out_proj, C, C >> h->n_up, 1, s)) return 1;
    }
    h->w_head = take(n_head);
    if (launch_repack_weight(prec, h->m("decoder.head.out.weight"), h->w_head, h->n_fft + 2, C, 1, s)) return 1;
    if (!h->head_bias_pad)
        B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->head_bias_pad), h->head_ld * sizeof(float)));
    B200_CUDA_OK(cudaMemsetAsync(h->head_bias_pad, 0, h->head_ld * sizeof(float), s));
    B200_CUDA_OK(cudaMemcpyAsync(h->head_bias_pad, h->m("decoder.head.out.bias"),
                                 (h->n_fft + 2) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    {
        std::vector<float2> tw(istft_table_words(h->hop));
        istft_fill_tables(h->hop, tw.data());
        if (!h->twiddle) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->twiddle), tw.size() * sizeof(float2)));
        B200_CUDA_OK(cudaMemcpy(h->twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    B200_CUDA_OK(cudaStreamSynchronize(s));
    B200_CUDA_OK(cudaFree(scratch));
    if (wn_scale) B200_CUDA_OK(cudaFree(wn_scale));
    h->finalized = true;
    return 0;
}

static int ensure_stats(B200Codec* h, int n_utts);

static int decode_varlen_locked(B200Codec* h, const void* ids_dev, int id_type,
                                const int32_t* seqlens_host, int n_utts, float* wav_dev, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (prepare(h, seqlens_host, n_utts, s)) return 1;
    B200_CHECK(ids_dev && wav_dev, "decode: null device buffer");
    if (ensure_stats(h, n_utts)) return 1;
    const int rc = forward(h, ids_dev, id_type, wav_dev, s);
    if (h->profiling) {
        cudaStreamSynchronize(s);
        collect_timers(h);
    }
    return rc;
}

static int ensure_stats(B200Codec* h, int n_utts) {
    // 8 GroupNorm layers x [n_utts][32 groups][sum, sumsq] fp64
    // + one float2 [n_utts][32] (mean, rstd) scratch per GroupNorm layer
    // + per transformer block, 4 GEMMs x ceil(rows / 256) m-block counters of the GEMM chains
    const size_t gn_bytes = (sizeof(double) * h->gn_slots * static_cast<size_t>(n_utts) * 64 + 255) & ~static_cast<size_t>(255);
    const size_t num_m = (static_cast<size_t>(h->rs.rows) + 255) / 256;
    const size_t ctr_bytes = sizeof(uint32_t) * static_cast<size_t>(h->L) * 4 * num_m;
    const size_t need = gn_bytes + ctr_bytes;
    h->stats_zero_bytes = need;
    if (h->gn_stats_bytes < need) {
        h->generation++;
        if (h->gn_stats) B200_CUDA_OK(cudaFree(h->gn_stats));
        h->gn_stats = nullptr;
        h->gn_stats_bytes = 0;
        B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&h->gn_stats), need * 2));
        h->gn_stats_bytes = need * 2;
    }
    h->chain_ctr = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(h->gn_stats) + gn_bytes);
    return 0;
}

int b200codec_decode_varlen(B200Codec* h, const void* ids_dev, int id_type,
                            const int32_t* seqlens_host, int n_utts, float* wav_dev, void* stream) {
    B200_CHECK(h != nullptr, "null handle");
    std::lock_guard<std::mutex> lock(h->mu);
    return decode_varlen_locked(h, ids_dev, id_type, seqlens_host, n_utts, wav_dev, stream);
}

// ---- cached streaming -----------------------------------------------------------------
int b200codec_stream_create(B200Codec* h, int n_streams, int new_tokens, int left_context, B200Stream** out) {
    B200_CHECK(h && out, "stream_create: null argument");
    B200_CHECK(h->finalized, "stream_create called before b200codec_finalize_weights");
    B200_CHECK(n_streams > 0 && new_tokens > 0 && left_context >= 0 && n_streams <= 65535, "stream_create: bad sizes");
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    B200Stream* st = new B200Stream();
    st->owner = h;
    st->n_streams = n_streams;
    st->new_tokens = new_tokens;
    st->left_context = left_context;
    // ring capacity: the context rounded up to whole pushes, plus the push itself (a push never wraps)
    st->cap = (left_context + new_tokens - 1) / new_tokens * new_tokens + new_tokens;
    const size_t es = operand_bytes(h->cfg.precision);
    st->layer_stride = ((static_cast<size_t>(n_streams) * st->cap + 128) * 2 * h->C * es + 1023) & ~static_cast<size_t>(1023);
    if (cudaMalloc(&st->kv_slab, st->layer_stride * h->L) != cudaSuccess) {
        set_error("stream_create: cannot allocate %zu bytes of key / value rings", st->layer_stride * h->L);
        delete st;
        return 1;
    }
    cudaMemset(st->kv_slab, 0, st->layer_stride * h->L);
    *out = st;
    return 0;
}

void b200codec_stream_destroy(B200Stream* st) {
    if (!st) return;
    if (st->owner) {
        cudaSetDevice(st->owner->cfg.device);
        cudaDeviceSynchronize();
    }
    if (st->kv_slab) cudaFree(st->kv_slab);
    if (st->work_dev) cudaFree(st->work_dev);
    if (st->work_host) cudaFreeHost(st->work_host);
    delete st;
}

int b200codec_stream_reset(B200Stream* st) {
    B200_CHECK(st != nullptr, "null stream state");
    st->seen = 0;  // rows beyond the valid count are never attended to: no need to clear the rings
    st->list_len = st->list_tkv = -1;
    return 0;
}

int b200codec_stream_capacity(const B200Stream* st) { return st ? st->cap : 0; }
int64_t b200codec_stream_tokens(const B200Stream* st) { return st ? st->seen : 0; }

int b200codec_stream_push(B200Codec* h, B200Stream* st, const void* ids_dev, int id_type, int overlap,
                          float* wav_dev, void* stream) {
    B200_CHECK(h && st && ids_dev && wav_dev, "stream_push: null argument");
    B200_CHECK(st->owner == h, "stream_push: the stream state belongs to another decoder handle");
    B200_CHECK(h->n_up == 0, "stream_push: the cached streaming path is built for the 16 kHz decoder (no upsampler)");
    B200_CHECK(overlap >= 0 && overlap <= st->seen && overlap <= st->cap - st->new_tokens,
               "stream_push: overlap %d must be within the tokens pushed so far (%lld) and the ring's context (%d)", overlap,
               (long long)st->seen, st->cap - st->new_tokens);
    std::lock_guard<std::mutex> lock(h->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int len = overlap + st->new_tokens;
    std::vector<int32_t> seqlens(st->n_streams, len);
    if (prepare(h, seqlens.data(), st->n_streams, s)) return 1;
    if (ensure_stats(h, st->n_streams)) return 1;
    // attention work items: every row of the push is a query; keys = the stream's ring
    const int tiles = (len + 127) / 128;
    const int n_work = tiles * st->n_streams;
    if (n_work > st->work_cap) {
        B200_CUDA_OK(cudaStreamSynchronize(s));
        if (st->work_dev) cudaFree(st->work_dev);
        if (st->work_host) cudaFreeHost(st->work_host);
        st->work_dev = nullptr;
        st->work_host = nullptr;
        B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&st->work_dev), sizeof(int4) * n_work * 2));
        B200_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&st->work_host), sizeof(int4) * n_work * 2));
        st->work_cap = n_work * 2;
    }
    st->overlap = overlap;
    st->wpos = static_cast<int>(st->seen % st->cap);
    const int64_t total = st->seen + st->new_tokens;
    st->t_kv = static_cast<int>(total < st->cap ? total : st->cap);
    if (st->list_len != len || st->list_tkv != st->t_kv) {  // steady state: the list of the last push still holds
        B200_CUDA_OK(cudaStreamSynchronize(s));  // the previous upload may still be reading the pinned list
        for (int u = 0, w = 0; u < st->n_streams; ++u)
            for (int q0 = 0; q0 < len; q0 += 128, ++w)
                st->work_host[w] = make_int4(u * (len + kGap) + q0, len - q0 < 128 ? len - q0 : 128, u * st->cap, st->t_kv);
        B200_CUDA_OK(cudaMemcpyAsync(st->work_dev, st->work_host, sizeof(int4) * n_work, cudaMemcpyHostToDevice, s));
        st->list_len = len;
        st->list_tkv = st->t_kv;
    }
    st->n_work = n_work;
    h->cur_stream = st;
    const int rc = forward(h, ids_dev, id_type, wav_dev, s);
    h->cur_stream = nullptr;
    if (rc == 0) st->seen += st->new_tokens;
    return rc;
}

int b200codec_take_id_error(B200Codec* h) {
    if (!h || !h->err_flag_host) return 0;
    const int v = *reinterpret_cast<volatile int*>(h->err_flag_host);
    *reinterpret_cast<volatile int*>(h->err_flag_host) = 0;
    return v;
}

int b200codec_decode_host(B200Codec* h, const void* ids_host, int id_type,
                          const int32_t* seqlens_host, int n_utts, float* wav_host, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CHECK(h && ids_host && wav_host && seqlens_host, "decode_host: null argument");
    B200_CHECK(n_utts > 0, "decode: empty batch (no utterances)");
    int64_t toks = 0;
    for (int u = 0; u < n_utts; ++u) {
        B200_CHECK(seqlens_host[u] > 0, "decode: utterance %d is empty", u);
        toks += seqlens_host[u];
    }
    if (check_ids_host(ids_host, id_type, toks)) return 1;
    std::lock_guard<std::mutex> lock(h->mu);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    const size_t id_bytes = static_cast<size_t>(toks) * (id_type == B200CODEC_IDS_I64 ? 8 : 4);
    const size_t wav_bytes = static_cast<size_t>(toks) * h->hop * h->total_up * sizeof(float);
    if (h->io_ids.ensure(id_bytes)) return 1;
    B200_CUDA_OK(cudaMemcpyAsync(h->io_ids.p, ids_host, id_bytes, cudaMemcpyHostToDevice, s));
    // Pinned (page-locked, device-mapped) output: the ISTFT kernel stores the PCM straight into host
    // memory, so the transfer overlaps the kernel instead of following it as a separate copy.
    float* wav_mapped = nullptr;
    if (g_zero_copy_out) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, wav_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            wav_mapped = static_cast<float*>(attr.devicePointer);
        else
            (void)cudaGetLastError();  // pageable memory: not an error, take the staged path
    }
    if (wav_mapped != nullptr) {
        if (decode_varlen_locked(h, h->io_ids.p, id_type, seqlens_host, n_utts, wav_mapped, s)) return 1;
    } else {
        if (h->io_wav.ensure(wav_bytes)) return 1;
        if (decode_varlen_locked(h, h->io_ids.p, id_type, seqlens_host, n_utts, h->io_wav.as<float>(), s))
            return 1;
        B200_CUDA_OK(cudaMemcpyAsync(wav_host, h->io_wav.p, wav_bytes, cudaMemcpyDeviceToHost, s));
    }
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

// Same with page-locked buffers and WITHOUT the final synchronisation: the PCM is stored straight into the
// mapped host buffer by the last kernel, the call returns after the enqueue, and the caller waits on the stream
// (or on an event recorded behind the call) before reading wav_host. A dataset sweep keeps two buffers in flight,
// so the host-side handling of batch k (copies, resampling, file writes) overlaps the decode of batch k + 1.
int b200codec_decode_host_async(B200Codec* h, const void* ids_host, int id_type, const int32_t* seqlens_host,
                                int n_utts, float* wav_host_pinned, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CHECK(h && ids_host && wav_host_pinned && seqlens_host, "decode_host_async: null argument");
    B200_CHECK(n_utts > 0, "decode: empty batch (no utterances)");
    int64_t toks = 0;
    for (int u = 0; u < n_utts; ++u) {
        B200_CHECK(seqlens_host[u] > 0, "decode: utterance %d is empty", u);
        toks += seqlens_host[u];
    }
    if (check_ids_host(ids_host, id_type, toks)) return 1;
    std::lock_guard<std::mutex> lock(h->mu);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    cudaPointerAttributes attr;
    const bool mapped = cudaPointerGetAttributes(&attr, wav_host_pinned) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
                        attr.devicePointer != nullptr;
    if (!mapped) (void)cudaGetLastError();
    B200_CHECK(mapped, "decode_host_async: wav_host must be page-locked, device-mapped memory (cudaHostAlloc / pin_memory)");
    const size_t id_bytes = static_cast<size_t>(toks) * (id_type == B200CODEC_IDS_I64 ? 8 : 4);
    if (id_bytes > h->io_ids.bytes) B200_CUDA_OK(cudaStreamSynchronize(s));  // a previous call may still read the old buffer
    if (h->io_ids.ensure(id_bytes)) return 1;
    B200_CUDA_OK(cudaMemcpyAsync(h->io_ids.p, ids_host, id_bytes, cudaMemcpyHostToDevice, s));
    return decode_varlen_locked(h, h->io_ids.p, id_type, seqlens_host, n_utts, static_cast<float*>(attr.devicePointer), s);
}

int b200codec_samples_per_token(const B200Codec* h) { return h ? h->hop * h->total_up : 0; }

int64_t b200codec_launch_count(const B200Codec* h) { return h ? h->launches : 0; }

int64_t b200codec_plan_generation(const B200Codec* h) { return h ? h->generation : 0; }

int b200codec_set_stage_taps(B200Codec* h, int on) {
    B200_CHECK(h != nullptr, "null handle");
    std::lock_guard<std::mutex> lock(h->mu);
    h->taps_on = on != 0;
    if (!h->taps_on) {
        for (auto& kv : h->taps) kv.second.buf.release();
        h->taps.clear();
    }
    return 0;
}

int b200codec_stage_width(B200Codec* h, const char* name) {
    if (!h || !name) return -1;
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->taps.find(name);
    return it == h->taps.end() || !it->second.valid ? -1 : it->second.width;
}

int64_t b200codec_stage_rows(B200Codec* h, const char* name) {
    if (!h || !name) return -1;
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->taps.find(name);
    if (it == h->taps.end() || !it->second.valid) return -1;
    int64_t factor = 1, toks = 0;
    for (int i = 0; i < it->second.space; ++i) factor *= h->up_f[i];
    for (int32_t T : h->plan_key) toks += static_cast<int64_t>(T) * factor;
    return toks;
}

int b200codec_read_stage(B200Codec* h, const char* name, float* host_out, size_t n_elems, void* stream) {
    B200_CHECK(h && name && host_out, "read_stage: null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->taps.find(name);
    B200_CHECK(it != h->taps.end() && it->second.valid,
               "read_stage: no tap named \"%s\" (enable b200codec_set_stage_taps, then decode)", name);
    const B200Codec::StageTap& t = it->second;
    int factor = 1;
    for (int i = 0; i < t.space; ++i) factor *= h->up_f[i];
    int64_t toks = 0;
    for (int32_t T : h->plan_key) toks += static_cast<int64_t>(T) * factor;
    B200_CHECK(static_cast<size_t>(toks) * t.width == n_elems, "read_stage %s: expected %lld x %d elements, got %zu",
               name, (long long)toks, t.width, n_elems);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CUDA_OK(cudaStreamSynchronize(s));
    size_t rows = 0;
    for (size_t u = 0; u < h->plan_key.size(); ++u)
        rows += static_cast<size_t>(h->plan_key[u]) * factor + (u + 1 < h->plan_key.size() ? kGap * factor : 0);
    std::vector<uint8_t> raw(rows * t.ld * t.elem);
    B200_CUDA_OK(cudaMemcpy(raw.data(), t.buf.p, raw.size(), cudaMemcpyDeviceToHost));
    const int prec = h->cfg.precision;
    size_t r = 0, o = 0;
    for (size_t u = 0; u < h->plan_key.size(); ++u) {
        const size_t T = static_cast<size_t>(h->plan_key[u]) * factor;
        for (size_t k = 0; k < T; ++k, ++r) {
            const uint8_t* src = raw.data() + r * t.ld * t.elem;
            for (int c = 0; c < t.width; ++c) {
                float v;
                if (t.elem == 4) v = reinterpret_cast<const float*>(src)[c];
                else if (prec == kPrecBf16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[c]);
                else v = __half2float(reinterpret_cast<const __half*>(src)[c]);
                host_out[o++] = v;
            }
        }
        r += static_cast<size_t>(kGap) * factor;
    }
    return 0;
}

int b200codec_set_frontend_fold(int mode) {
    B200_CHECK(mode >= 0 && mode <= 2, "front-end mode must be 0 (lookup + conv7 GEMM), 1 (folded, tensor cores) "
                                       "or 2 (folded, fp32 FMA kernel)");
    g_frontend_fold = mode;
    return 0;
}

int b200codec_set_gemm_early_weights(int on) {
    g_gemm_early_weights = on != 0;
    return 0;
}

int b200codec_set_gemm_chain(int on) {
    g_gemm_chain = on != 0;
    return 0;
}

int b200codec_set_istft_tile(int hops) {
    B200_CHECK(hops == 0 || hops == 12 || hops == 28, "istft tile must be 0 (auto), 12 (8 warps, 2 CTAs per SM) or 28 (16 warps) output hops");
    g_istft_hops = hops;
    return 0;
}

int b200codec_set_gemm_narrow_tiles(int mode) {
    B200_CHECK(mode >= 0 && mode <= 2, "narrow-tile mode must be 0 (256-wide only), 1 (default) or 2 (64-wide for every N %% 256 != 0)");
    g_gemm_narrow_tiles = mode;
    return 0;
}

int b200codec_set_zero_copy_output(int on) {
    g_zero_copy_out = on != 0;
    return 0;
}

int b200codec_set_pdl(int on) {
    g_use_pdl = on != 0;
    return 0;
}

int b200codec_profile(B200Codec* h, int on) {
    B200_CHECK(h != nullptr, "null handle");
    h->profiling = on != 0;
    h->stage_ms.clear();
    h->stage_order.clear();
    return 0;
}

int b200codec_stage_times(B200Codec* h, int max_stages, const char** names_out, float* ms_out,
                          int* n_out) {
    B200_CHECK(h && names_out && ms_out && n_out, "stage_times: null argument");
    int n = 0;
    for (const auto& name : h->stage_order) {
        if (n >= max_stages) break;
        names_out[n] = name.c_str();
        ms_out[n] = h->stage_ms[name];
        ++n;
    }
    *n_out = n;
    return 0;
}

// ---- per-stage entry points -------------------------------------------------------------

int b200codec_fsq_lookup(B200Codec* h, const void* ids_dev, int id_type, int64_t n, float* out_dev,
                         void* stream) {
    B200_CHECK(h && ids_dev && out_dev, "fsq_lookup: null argument");
    const float* w = h->m("decoder.quantizer.project_out.weight");
    const float* b = h->m("decoder.quantizer.project_out.bias");
    B200_CHECK(w && b, "fsq_lookup: quantizer.project_out has not been loaded");
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    if (launch_fsq_lookup(ids_dev, id_type, nullptr, static_cast<int>(n), w, b, h->V, out_dev, h->V,
                          -1, h->err_flag_dev, static_cast<cudaStream_t>(stream)))
        return 1;
    h->launches++;
    return 0;
}

int b200codec_map_speech_tokens(const int32_t* table_dev, int vocab, const int64_t* tok_dev, const int32_t* seq_off_dev,
                                int n_seq, int32_t* codes_dev, int32_t* out_len_dev, void* stream) {
    B200_CHECK(table_dev && tok_dev && seq_off_dev && codes_dev && out_len_dev, "map_speech_tokens: null argument");
    B200_CHECK(vocab > 0 && n_seq >= 0, "map_speech_tokens: bad sizes");
    return launch_map_speech_tokens(table_dev, vocab, reinterpret_cast<const long long*>(tok_dev), seq_off_dev, n_seq,
                                    codes_dev, out_len_dev, static_cast<cudaStream_t>(stream));
}

int b200codec_fsq_quantize(B200Codec* h, const float* feats_dev, int ld, int64_t n_tokens, void* ids_dev,
                           int id_type, float* z_dev, int pre_bound, void* stream) {
    B200_CHECK(h && feats_dev && ids_dev, "fsq_quantize: null argument");
    B200_CHECK(id_type == 0 || id_type == 1, "fsq_quantize: id_type must be 0 (int32) or 1 (int64)");
    B200_CHECK(n_tokens >= 0 && n_tokens <= 0x7fffffff, "fsq_quantize: bad token count");
    const float* w = h->m("decoder.quantizer.project_in.weight");
    const float* b = h->m("decoder.quantizer.project_in.bias");
    B200_CHECK(w && b, "fsq_quantize: quantizer.project_in has not been loaded");
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    if (launch_fsq_quantize(feats_dev, ld, static_cast<int>(n_tokens), w, b, h->V, pre_bound, ids_dev, id_type,
                            z_dev, static_cast<cudaStream_t>(stream)))
        return 1;
    h->launches++;
    return 0;
}

int b200codec_istft(B200Codec* h, const float* x_pred_dev, int ld, const int32_t* seqlens_host,
                    int n_utts, float* wav_dev, void* stream) {
    B200_CHECK(h && x_pred_dev && wav_dev && seqlens_host, "istft: null argument");
    B200_CHECK(h->finalized, "istft called before b200codec_finalize_weights");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CUDA_OK(cudaSetDevice(h->cfg.device));
    TempPlan tp;
    if (tp.build(seqlens_host, n_utts, s)) return 1;
    IstftTables tab;
    tab.twiddle = h->twiddle;
    tab.window = h->m("decoder.head.istft.window");
    if (launch_istft(x_pred_dev, ld, tp.rs, tab, h->hop, wav_dev, s)) return 1;
    h->launches++;
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int b200codec_gemm(int precision, const void* a_dev, const void* w_dev, int M, int N, int Cin,
                   int taps, void* out_dev, int out_dtype, int ldc, const float* bias_dev,
                   const float* residual_dev, int ld_res, int act, void* stream) {
    B200_CHECK(a_dev && w_dev && out_dev, "gemm: null argument");
    GemmCall c;
    c.precision = precision;
    c.a = a_dev;
    c.a_rows = M;
    c.Cin = Cin;
    c.w = w_dev;
    c.N = N;
    c.taps = taps;
    c.out = out_dev;
    c.out_fp32 = out_dtype == 0 ? 1 : 0;
    c.ldc = ldc;
    c.n_store = (N + 63) / 64 * 64;  // weight rows past N read as zeros (TMA out-of-bounds fill)
    c.bias = bias_dev;
    c.residual = residual_dev;
    c.ld_res = ld_res;
    c.row_valid = nullptr;
    c.act = act;
    B200_CHECK(c.n_store <= ldc, "gemm: ldc (%d) must cover N rounded up to 64 (%d)", ldc, c.n_store);
    B200_CHECK(N % 64 == 0 || (bias_dev == nullptr && residual_dev == nullptr),
               "gemm: bias / residual need N (%d) to be a multiple of 64", N);
    return launch_gemm(c, static_cast<cudaStream_t>(stream));
}

int b200codec_rmsnorm(int precision, const float* x_dev, const float* w_dev, int rows, int dim,
                      float eps, void* out_dev, void* stream) {
    B200_CHECK(x_dev && w_dev && out_dev, "rmsnorm: null argument");
    return launch_rmsnorm(precision, x_dev, w_dev, rows, dim, eps, out_dev,
                          static_cast<cudaStream_t>(stream));
}

int b200codec_layernorm(int precision, const float* x_dev, const float* w_dev, const float* b_dev,
                        int rows, int dim, float eps, void* out_dev, void* stream) {
    B200_CHECK(x_dev && w_dev && b_dev && out_dev, "layernorm: null argument");
    return launch_layernorm(precision, x_dev, w_dev, b_dev, rows, dim, eps, out_dev,
                            static_cast<cudaStream_t>(stream));
}

int b200codec_groupnorm_swish(int precision, const float* x_dev, const float* gamma_dev,
                              const float* beta_dev, const int32_t* seqlens_host, int n_utts,
                              int dim, float eps, void* out_dev, void* stream) {
    B200_CHECK(x_dev && gamma_dev && beta_dev && out_dev && seqlens_host, "groupnorm: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TempPlan tp;
    if (tp.build(seqlens_host, n_utts, s)) return 1;
    double* stats = nullptr;
    const size_t bytes = sizeof(double) * 64 * static_cast<size_t>(n_utts);
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&stats), bytes));
    int rc = 0;
    if (cudaMemsetAsync(stats, 0, bytes, s) != cudaSuccess) rc = 1;
    if (!rc) rc = launch_groupnorm_stats(x_dev, tp.rs, dim, stats, s);
    if (!rc) rc = launch_groupnorm_apply_swish(precision, x_dev, tp.rs, dim, stats, gamma_dev,
                                               beta_dev, eps, out_dev, s);
    cudaStreamSynchronize(s);
    cudaFree(stats);
    return rc;
}

int b200codec_attention(int precision, const void* qkv_dev, const int32_t* seqlens_host, int n_utts,
                        int heads, void* out_dev, void* stream) {
    B200_CHECK(qkv_dev && out_dev && seqlens_host, "attention: null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TempPlan tp;
    if (tp.build(seqlens_host, n_utts, s)) return 1;
    int rc = run_attention(precision, qkv_dev, tp.rs, heads, out_dev, s);
    cudaStreamSynchronize(s);
    return rc;
}

}  // extern "C"
