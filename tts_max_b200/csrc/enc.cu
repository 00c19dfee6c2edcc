// b200enc: the ENCODE direction up to the FSQ ids -- AcousticEncoder, SemanticEncoder,
// fusion layer and quantise of tts/core/codec/encoder.py:58-78; the w2v-BERT model that produces the semantic
// features stays in HuggingFace and its hidden state is an input.
//
// Replaces (reference, all ATen there):
//   AcousticEncoder.forward        tts/core/codec/encoder_modules.py:128-191
//     ResidualUnit / EncoderBlock   :20-69   (weight-normed Conv1d k = 7 dilated 1/3/9, k = 1, k = 2s stride s)
//     Activation1d(SnakeBeta)       tts/core/codec/activations.py:47-110, filters.py:87-135
//   SemanticEncoder.forward        tts/core/codec/encoder_modules.py:72-125
//   fusion_layer, Encoder.quantize tts/core/codec/encoder.py:42, 66-78
//
// B200 mapping. Activations are token-major (channels-last) [rows, C] like the decoder's, dense (row pitch C, also
// for C = 48 and 96: the GEMM tiles are 64 wide, TMA out-of-bounds fill pads the operand rows with zeros in shared
// memory and the TMA stores drop the pad columns, so the padding costs FLOPs but no memory traffic), the clips of a
// call in one launch sequence:
//   * every Conv1d is the tcgen05 GEMM of gemm_tc05*.cuh: k = 7 dilated = 7 row-shifted K-slabs with a row
//     stride of `dilation` (zero padding = TMA out-of-bounds fill at the array ends); the strided
//     down-sampling conv (k = 2s, stride s) is a 3-tap conv over the SAME buffer viewed as [rows / s, s * C]
//     with the kernel laid out (and zero-filled) per super-row at load time; weight-norm is folded at load;
//   * Activation1d (2x FIR up-sampling, SnakeBeta, 2x FIR down-sampling, replicate padding) is one fused
//     CUDA-core kernel: fp32 in, 16-bit GEMM operand out, nothing at the doubled rate ever reaches memory;
//   * bias / ReLU / residual adds ride the GEMM epilogues; the fusion Linear reads the two encoders' 16-bit
//     outputs from one [T, 2048] buffer they were written into side by side (the concat is free).
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
namespace {

constexpr int kStages = 5;
constexpr int kStrides[kStages] = {2, 2, 4, 4, 5};
constexpr int kDil[3] = {1, 3, 9};
constexpr int kGenFeatures = 48;
constexpr int kOutDim = 1024;
constexpr int kHopTotal = 320;

inline int pad64(int c) { return (c + 63) / 64 * 64; }

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        B200_CUDA_OK(cudaMalloc(&p, need + need / 4));
        bytes = need + need / 4;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// conv_blocks[0]: Conv1d(1 -> 48, k = 7, padding 3), weight-normed: x0[t, c] = b[c] + sum_k w[c][k] wav[t + k - 3].
// blockIdx.y = clip: wav is compact [clips][S], out is the slotted row space (clip b at row b * slot).
// A thread owns 4 channels (its 28 taps and 4 biases live in registers; 128-bit stores) and walks kConv0Rows
// consecutive samples with a sliding window of 7 input samples; C / 4 threads share a sample; out rows are dense.
constexpr int kConv0Rows = 32;

__global__ void __launch_bounds__(256)
enc_conv0_kernel(const float* __restrict__ wav, int S, int slot, const float* __restrict__ w /*[C][7]*/,
                 const float* __restrict__ bias, int C, float* __restrict__ out /*[clips * slot][C]*/) {
    pdl_launch_dependents();
    pdl_wait();
    const int quads = C >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (idx % quads) * 4;
    const int t0 = (idx / quads) * kConv0Rows;
    if (t0 >= S) return;
    wav += static_cast<size_t>(blockIdx.y) * S;
    out += (static_cast<size_t>(blockIdx.y) * slot + t0) * C + c0;
    float wr[4][7], br[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        br[j] = bias[c0 + j];
#pragma unroll
        for (int k = 0; k < 7; ++k) wr[j][k] = w[(c0 + j) * 7 + k];
    }
    auto at = [&](int i) { return (i >= 0 && i < S) ? __ldg(wav + i) : 0.f; };
    float xin[7];
#pragma unroll
    for (int k = 0; k < 6; ++k) xin[k + 1] = at(t0 + k - 3);
#pragma unroll 8
    for (int r = 0; r < kConv0Rows; ++r) {
        if (t0 + r >= S) break;
#pragma unroll
        for (int k = 0; k < 6; ++k) xin[k] = xin[k + 1];
        xin[6] = at(t0 + r + 3);
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = br[j];
#pragma unroll
            for (int k = 0; k < 7; ++k) a = fmaf(wr[j][k], xin[k], a);
            acc[j] = a;
        }
        *reinterpret_cast<float4*>(out + static_cast<size_t>(r) * C) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// Activation1d(SnakeBeta) fused (activations.py:90-110, filters.py:87-135 with ratio 2, 12-tap kaiser-sinc
// filters, replicate padding):
//   u[2t]   = 2 sum_j fu[2j+1] x[t+2-j]       u[2t+1] = 2 sum_j fu[2j] x[t+3-j]      (j = 0..5, rows clamped;
//                                                                                    the 2 is folded into f.up)
//   s[n]    = u[n] + sin^2(u[n] e^alpha) / (e^beta + 1e-9)                           (n in [0, 2T))
//   y[t]    = sum_k fd[k] s[clamp(2t + k - 5, 0, 2T - 1)]                            (k = 0..11)
// A thread owns TWO adjacent channels (packed fp32x2 arithmetic: FFMA2 / FMUL2, 64-bit loads, 32-bit stores; a
// warp covers 64 channels of a row) and walks kM * 6 - 5 consecutive rows with two sliding windows in registers:
// step t loads nothing new but row t + 5 (prefetched a macro-step of 6 rows ahead), computes the two new snake
// values s[2t + 5], s[2t + 6] from rows t .. t + 5 into a 12-entry circular window and emits y[t] from the window.
// Five warm-up steps fill the window, so a chunk costs (rows + 5) steps: 8 % over the minimum for 67 rows (short
// inputs use 19-row chunks to keep the SMs busy). ~22 instructions per output instead of 53 for the first version
// (one channel per thread, a 16-row chunk recomputing 42 snake values, Cody-Waite + FRND range reduction);
// sin is MUFU.SIN on the raw argument: |u e^alpha| stays below ~10^2, where the approximation's phase error
// (~|z| 2^-24 revolutions) is three orders of magnitude under the 16-bit rounding of the result.
// Batches: blockIdx.y = clip, clip b occupies rows [b * slot, b * slot + T) of x and out; the slot - T gap rows
// behind every clip are written as ZEROS -- they are the zero padding of the convolution that reads `out`
// (its taps reach at most 27 rows across a clip edge; the gap is at least 30 rows at every level).
struct SnakeFilters {
    float up[12];
    float down[12];
};

template <typename OutT, int kM, bool kEdge>
__device__ __forceinline__ void snake_chunk(const float* __restrict__ x, int T, int slot, int P, int c, int t0, float2 a,
                                            float2 ib, const SnakeFilters& f, OutT* __restrict__ out) {
    // interior chunks step two pointers (rows are loaded and stored in increasing order) instead of deriving a
    // 64-bit address per row; edge chunks clamp the row index
    const float* px = x + static_cast<ptrdiff_t>(t0 - 5) * P + c;
    OutT* po = out + static_cast<size_t>(t0) * P + c;
    auto ld = [&](int t) -> float2 {
        if (kEdge) {
            t = min(max(t, 0), T - 1);
            return __ldg(reinterpret_cast<const float2*>(x + static_cast<size_t>(t) * P + c));
        }
        const float2 v = __ldg(reinterpret_cast<const float2*>(px));
        px += P;
        return v;
    };
    auto snake2 = [&](float2 u) -> float2 {
        const float2 z = __fmul2_rn(u, a);
        const float sx = __sinf(z.x), sy = __sinf(z.y);
        const float2 sn = make_float2(sx, sy);
        return __ffma2_rn(__fmul2_rn(sn, sn), ib, u);
    };
    float2 s_first = make_float2(0.f, 0.f), s_last = s_first;  // s[0], s[2T - 1]: the replicate padding of the doubled signal
    if (kEdge) {
        float2 u0 = make_float2(0.f, 0.f), u1 = u0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            u0 = __ffma2_rn(make_float2(f.up[2 * j + 1], f.up[2 * j + 1]), ld(2 - j), u0);
            u1 = __ffma2_rn(make_float2(f.up[2 * j], f.up[2 * j]), ld(T - 1 + 3 - j), u1);
        }
        s_first = snake2(u0);
        s_last = snake2(u1);
    }
    float2 xb[11];  // rows tm .. tm + 10 of the current macro-step (steps tm .. tm + 5)
    float2 xn[6];   // rows tm + 11 .. tm + 16: the next macro-step's new rows, in flight
    float2 sw[12];  // circular: after step i of a macro-step the oldest value sits at (2 i + 2) % 12
#pragma unroll
    for (int k = 0; k < 12; ++k) sw[k] = make_float2(0.f, 0.f);
    int tm = t0 - 5;
#pragma unroll
    for (int r = 0; r < 11; ++r) xb[r] = ld(tm + r);
#pragma unroll 1
    for (int m = 0; m < kM; ++m, tm += 6) {
        if (m + 1 < kM) {
#pragma unroll
            for (int r = 0; r < 6; ++r) xn[r] = ld(tm + 11 + r);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int t = tm + i;
            float2 uo = make_float2(0.f, 0.f), ue = uo;  // u[2t + 5] (odd: 2 (t + 2) + 1), u[2t + 6] (even: 2 (t + 3))
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                uo = __ffma2_rn(make_float2(f.up[2 * j], f.up[2 * j]), xb[i + 5 - j], uo);
                ue = __ffma2_rn(make_float2(f.up[2 * j + 1], f.up[2 * j + 1]), xb[i + 5 - j], ue);
            }
            float2 so = snake2(uo), se = snake2(ue);
            if (kEdge) {
                const int n = 2 * t + 5;
                if (n < 0) so = s_first;
                if (n + 1 < 0) se = s_first;
                if (n > 2 * T - 1) so = s_last;
                if (n + 1 > 2 * T - 1) se = s_last;
            }
            sw[(2 * i) % 12] = so;
            sw[(2 * i + 1) % 12] = se;
            if (m > 0 || i == 5) {  // the first five steps of a chunk only fill the window
                float2 y = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 12; ++k) y = __ffma2_rn(make_float2(f.down[k], f.down[k]), sw[(2 * i + 2 + k) % 12], y);
                if (!kEdge || t < T) *reinterpret_cast<uint32_t*>(po) = Half16<OutT>::pack(y.x, y.y);
                else if (t < slot) *reinterpret_cast<uint32_t*>(po) = 0u;
                po += P;
            }
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) xb[r] = xb[r + 6];
#pragma unroll
        for (int r = 0; r < 6; ++r) xb[5 + r] = xn[r];
    }
}

template <typename OutT, int kM>
__global__ void __launch_bounds__(256)
snake_aa_kernel(const float* __restrict__ x, int T, int slot, int P, const float* __restrict__ alpha /*[P] e^alpha*/,
                const float* __restrict__ inv_beta /*[P] 1 / (e^beta + 1e-9)*/, SnakeFilters f,
                OutT* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int K = kM * 6 - 5;  // rows per chunk
    // thread = (chunk of K rows, channel pair), pairs fastest: for P a multiple of 64 a warp is 64 channels of one
    // chunk; for P = 48 / 96 a warp may straddle two chunks (two coalesced segments per access)
    const int pairs = P >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = idx / pairs;
    const int c = (idx - chunk * pairs) * 2;
    const int t0 = chunk * K;
    if (t0 >= slot) return;
    x += static_cast<size_t>(blockIdx.y) * slot * P;
    out += static_cast<size_t>(blockIdx.y) * slot * P;
    if (t0 >= T) {  // a chunk of gap rows only
        for (int t = t0; t < t0 + K && t < slot; ++t) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(t) * P + c) = 0u;
        return;
    }
    const float2 a = *reinterpret_cast<const float2*>(alpha + c), ib = *reinterpret_cast<const float2*>(inv_beta + c);
    if (t0 >= 5 && t0 + K + 4 < T) snake_chunk<OutT, kM, false>(x, T, slot, P, c, t0, a, ib, f, out);
    else snake_chunk<OutT, kM, true>(x, T, slot, P, c, t0, a, ib, f, out);
}

// w2v hidden state: compact fp32 [clips][T][C] -> operand dtype in the slotted row space [clips * slot][C], gap rows
// zero; also writes the row-validity bytes the token-level GEMM epilogues use to keep gap rows zero
template <typename OutT>
__global__ void cast_slotted_kernel(const float* __restrict__ x, int T, int slot, int C, OutT* __restrict__ out,
                                    uint8_t* __restrict__ row_valid) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;               // slotted row
    const int clip = row / slot, pos = row - clip * slot;
    const bool ok = pos < T;
    if (threadIdx.x == 0) row_valid[row] = ok ? 1 : 0;
    const float4* src = reinterpret_cast<const float4*>(x + (static_cast<size_t>(clip) * T + pos) * C);
    OutT* dst = out + static_cast<size_t>(row) * C;
    for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
        const float4 v = ok ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 pk;
        pk.x = Half16<OutT>::pack(v.x, v.y);
        pk.y = Half16<OutT>::pack(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + 4 * i) = pk;
    }
}

// weight-normed Conv1d weight [Cout][Cin][k] (scale[Cout] = g / ||v||) -> GEMM operand [Npad][k * Cinpad], zero padded
template <typename T>
__global__ void enc_repack_conv_kernel(const float* __restrict__ v, const float* __restrict__ scale, T* __restrict__ dst,
                                       int Cout, int Cin, int k, int Npad, int Cinpad) {
    const size_t total = static_cast<size_t>(Npad) * k * Cinpad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cinpad);
        const size_t rest = i / Cinpad;
        const int tap = static_cast<int>(rest % k);
        const int n = static_cast<int>(rest / k);
        float w = 0.f;
        if (n < Cout && c < Cin) w = v[(static_cast<size_t>(n) * Cin + c) * k + tap] * (scale ? scale[n] : 1.f);
        dst[i] = Half16<T>::from_float(w);
    }
}

// strided conv (k = 2s, stride s, padding p) as a 3-tap conv over super-rows of s input rows (s * Cin dense columns,
// padded as a whole to Kt = a multiple of 64):
// dst[n][tap' * Kt + r * Cin + c] = w[n][c][s (tap' - 1) + r + p] when that kernel index exists, else 0
template <typename T>
__global__ void enc_repack_strided_kernel(const float* __restrict__ v, const float* __restrict__ scale, T* __restrict__ dst,
                                          int Cout, int Cin, int s, int p, int Npad, int Kt) {
    const int k = 2 * s;
    const size_t total = static_cast<size_t>(Npad) * 3 * Kt;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int col = static_cast<int>(i % Kt);
        const size_t rest = i / Kt;
        const int tp = static_cast<int>(rest % 3);
        const int n = static_cast<int>(rest / 3);
        const int r = col / Cin, c = col - r * Cin;
        const int kk = s * (tp - 1) + r + p;
        float w = 0.f;
        if (n < Cout && r < s && kk >= 0 && kk < k) w = v[(static_cast<size_t>(n) * Cin + c) * k + kk] * scale[n];
        dst[i] = Half16<T>::from_float(w);
    }
}

__global__ void enc_pad_vec_kernel(const float* __restrict__ src, int n, int npad, int mode, float* __restrict__ dst) {
    // mode 0: copy, zero pad; 1: exp(src), pad exp(0) = 1; 2: 1 / (exp(src) + 1e-9), pad 1 / (1 + 1e-9)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const float v = i < n ? src[i] : 0.f;
    dst[i] = mode == 0 ? (i < n ? v : 0.f) : mode == 1 ? expf(v) : 1.f / (expf(v) + 1e-9f);
}

struct TensorSpec {
    std::string key;
    std::vector<int64_t> shape;
    size_t numel() const {
        size_t n = 1;
        for (auto d : shape) n *= static_cast<size_t>(d);
        return n;
    }
};

struct ActW {           // Activation1d(SnakeBeta)
    float* alpha = nullptr;     // [P] e^alpha
    float* inv_beta = nullptr;  // [P] 1 / (e^beta + 1e-9)
    SnakeFilters f;
};
struct ConvW {
    void* w = nullptr;          // operand dtype [Npad][K]
    float* bias = nullptr;      // [Npad]
};
struct UnitW {
    ActW act0, act2;
    ConvW conv7, conv1;
};
struct StageW {
    UnitW unit[3];
    ActW act;
    ConvW down;
};

}  // namespace
}  // namespace b200

using namespace b200;

struct B200Enc {
    int precision = 0, device = 0;
    std::vector<TensorSpec> specs;
    std::map<std::string, int> index;
    std::vector<float*> master;
    bool finalized = false;
    std::mutex mu;

    Buf wbuf;  // operand-dtype weights
    Buf fbuf;  // fp32 vectors (padded biases, exp(alpha), 1 / exp(beta))
    float* conv0_w = nullptr;
    float* conv0_b = nullptr;
    StageW stage[kStages];
    ActW act_final;
    ConvW conv_final;
    ConvW sem_init, sem_rb1, sem_rb3, sem_final, fusion;

    // workspace for a batch of up to ws_samples slotted samples (clips x (samples + gap))
    Buf ws;
    int64_t ws_samples = 0;
    uint8_t* row_valid = nullptr;  // [token rows] 1 on a clip's tokens, 0 on gap rows
    void* ids_tmp = nullptr;       // [token rows] ids in the slotted row space (int64 sized)
    float* x[kStages + 1];   // fp32 [R_i][P_i]
    float* hb[kStages];      // fp32 conv7 output
    void* a[kStages + 1];    // operand [R_i][P_i]
    void *s16a, *s16b;       // semantic encoder operands [T][1024]
    float *sr, *sx;          // semantic fp32 [T][1024]
    void* cat16;             // [T][2048] operand: [semantic | acoustic]
    float *ac32, *sem32, *hid32;  // fp32 [T][1024], [T][1024], [T][2048]

    int64_t launches = 0;
    // debug taps (b200enc_set_stage_taps): copies of conv_blocks[0 .. 5] outputs taken when they are produced
    // (the residual units update x[i] in place afterwards)
    bool taps_on = false;
    Buf tap[kStages + 1];
    int64_t tap_samples = 0;
    int tap_clips = 0;

    const float* m(const std::string& key) const {
        auto it = index.find(key);
        return it == index.end() ? nullptr : master[it->second];
    }
};

namespace {

void add_spec(B200Enc* h, const std::string& key, std::vector<int64_t> shape) {
    h->index[key] = static_cast<int>(h->specs.size());
    h->specs.push_back({key, std::move(shape)});
}

// Encoder.state_dict() keys (encoder.py:28-44) in module order: semantic_encoder, acoustic_encoder,
// fusion_layer, quantizer (project_in / project_out are the only persistent tensors of ResidualFSQ)
void build_specs(B200Enc* h) {
    const std::string s = "semantic_encoder.";
    add_spec(h, s + "initial_conv.weight", {1024, 1024, 3});
    add_spec(h, s + "residual_blocks.1.weight", {1024, 1024, 3});
    add_spec(h, s + "residual_blocks.1.bias", {1024});
    add_spec(h, s + "residual_blocks.3.weight", {1024, 1024, 3});
    add_spec(h, s + "residual_blocks.3.bias", {1024});
    add_spec(h, s + "final_conv.weight", {1024, 1024, 3});
    auto act = [&](const std::string& p, int64_t c) {
        add_spec(h, p + "act.alpha", {c});
        add_spec(h, p + "act.beta", {c});
        add_spec(h, p + "upsample.filter", {1, 1, 12});
        add_spec(h, p + "downsample.lowpass.filter", {1, 1, 12});
    };
    auto wn = [&](const std::string& p, int64_t cout, int64_t cin, int64_t k) {
        add_spec(h, p + "bias", {cout});
        add_spec(h, p + "weight_g", {cout, 1, 1});
        add_spec(h, p + "weight_v", {cout, cin, k});
    };
    const std::string a = "acoustic_encoder.";
    wn(a + "conv_blocks.0.", kGenFeatures, 1, 7);
    int64_t d = kGenFeatures;
    for (int i = 0; i < kStages; ++i) {
        const std::string p = a + "conv_blocks." + std::to_string(i + 1) + ".";
        for (int u = 0; u < 3; ++u) {
            const std::string q = p + "block." + std::to_string(u) + ".";
            act(q + "block.0.", d);
            wn(q + "block.1.", d, d, 7);
            act(q + "block.2.", d);
            wn(q + "block.3.", d, d, 1);
        }
        act(p + "block.3.", d);
        wn(p + "block.4.", 2 * d, d, 2 * kStrides[i]);
        d *= 2;
    }
    act(a + "conv_final_block.0.", d);
    wn(a + "conv_final_block.1.", kOutDim, d, 3);
    add_spec(h, "fusion_layer.weight", {2048, 2048});
    add_spec(h, "fusion_layer.bias", {2048});
    add_spec(h, "quantizer.project_in.weight", {8, 2048});
    add_spec(h, "quantizer.project_in.bias", {8});
    add_spec(h, "quantizer.project_out.weight", {2048, 8});
    add_spec(h, "quantizer.project_out.bias", {2048});
}

#define ENC_RUN(expr)              \
    do {                           \
        if ((expr) != 0) return 1; \
        h->launches++;             \
    } while (0)

template <typename T>
int repack_conv(const float* v, const float* scale, void* dst, int Cout, int Cin, int k, int Npad, int Cinpad, cudaStream_t s) {
    const size_t total = static_cast<size_t>(Npad) * k * Cinpad;
    const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, 8192));
    enc_repack_conv_kernel<T><<<grid, 256, 0, s>>>(v, scale, static_cast<T*>(dst), Cout, Cin, k, Npad, Cinpad);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
template <typename T>
int repack_strided(const float* v, const float* scale, void* dst, int Cout, int Cin, int st, int p, int Npad, int Kt,
                   cudaStream_t s) {
    const size_t total = static_cast<size_t>(Npad) * 3 * Kt;
    const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, 8192));
    enc_repack_strided_kernel<T><<<grid, 256, 0, s>>>(v, scale, static_cast<T*>(dst), Cout, Cin, st, p, Npad, Kt);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int kM>
int snake_launch(B200Enc* h, const float* x, int T, int slot, int clips, int P, const ActW& w, void* out, cudaStream_t s) {
    constexpr int K = kM * 6 - 5;
    const int64_t threads = static_cast<int64_t>((slot + K - 1) / K) * (P / 2);
    dim3 grid(static_cast<unsigned>((threads + 255) / 256), clips);
    if (h->precision == kPrecBf16)
        B200_CUDA_OK(launch_kernel(snake_aa_kernel<__nv_bfloat16, kM>, grid, dim3(256), 0, s, x, T, slot, P, w.alpha, w.inv_beta, w.f,
                                   static_cast<__nv_bfloat16*>(out)));
    else
        B200_CUDA_OK(launch_kernel(snake_aa_kernel<__half, kM>, grid, dim3(256), 0, s, x, T, slot, P, w.alpha, w.inv_beta, w.f,
                                   static_cast<__half*>(out)));
    return 0;
}

int snake(B200Enc* h, const float* x, int T, int slot, int clips, int P, const ActW& w, void* out, cudaStream_t s) {
    // 67-row chunks (8 % warm-up overhead) once they still give every SM a full set of warps, else 19-row chunks
    const int64_t warps67 = static_cast<int64_t>((slot + 66) / 67) * (P / 2) * clips / 32;
    if (warps67 >= 64 * kNumSMs) return snake_launch<12>(h, x, T, slot, clips, P, w, out, s);
    return snake_launch<4>(h, x, T, slot, clips, P, w, out, s);
}

// Conv / Linear over dense [rows, a_cols] -> [rows, out_cols] matrices; the GEMM sees them padded to multiples of 64
// (weights and biases are laid out padded at load time)
GemmCall conv_call(B200Enc* h, const void* a, int rows, int a_cols, const ConvW& w, int out_cols, int taps, int dil, void* out,
                   bool out_fp32, const float* residual, int act, const uint8_t* row_valid = nullptr) {
    GemmCall c{};
    c.precision = h->precision;
    c.a = a;
    c.a_rows = rows;
    c.Cin = pad64(a_cols);
    c.a_cols = a_cols % 64 ? a_cols : 0;
    c.w = w.w;
    c.N = pad64(out_cols);
    c.taps = taps;
    c.tap_pad = -1;
    c.tap_dil = dil;
    c.out = out;
    c.out_fp32 = out_fp32 ? 1 : 0;
    c.ldc = out_cols;
    c.n_store = c.N;
    c.out_cols = out_cols % 64 ? out_cols : 0;
    c.bias = w.bias;
    c.residual = residual;
    c.ld_res = out_cols;
    c.row_valid = row_valid;
    c.act = act;
    c.out16_scale = 1.f;
    c.ss_in_scale = 1.f;
    return c;
}

// Batches: the clips of one call (all n_samples long) sit in ONE slotted row space. Clip b owns rows
// [b * slot_l, b * slot_l + T_l) at level l; behind every clip there are kGapTokens tokens' worth of gap rows
// (6 tokens = 1920 samples: 30 rows at the 250 Hz level, where a dilation-9 conv reaches 27 rows) that the
// snake kernel writes as zeros, so a convolution's taps read zero padding across a clip edge. slot_l is a
// multiple of every later stride, so the [rows / s, s * C] super-row view of the strided convs stays aligned.
constexpr int kGapTokens = 6;

int ensure_ws(B200Enc* h, int64_t S) {  // S: slotted samples of the whole batch
    if (S <= h->ws_samples) return 0;
    const size_t es = 2;
    const int64_t T = S / kHopTotal;
    auto al = [](size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); };
    size_t total = 0;
    int64_t rows = S;
    int c = kGenFeatures;
    size_t off_x[kStages + 1], off_h[kStages], off_a[kStages + 1];
    for (int i = 0; i <= kStages; ++i) {
        const size_t P = c;  // dense rows
        off_x[i] = total; total += al(static_cast<size_t>(rows) * P * 4);
        if (i < kStages) { off_h[i] = total; total += al(static_cast<size_t>(rows) * P * 4); }
        off_a[i] = total; total += al(static_cast<size_t>(rows) * P * es);
        if (i < kStages) { rows /= kStrides[i]; c *= 2; }
    }
    const size_t sz16 = al(static_cast<size_t>(T) * 1024 * es), sz32 = al(static_cast<size_t>(T) * 1024 * 4);
    const size_t o_s16a = total; total += sz16;
    const size_t o_s16b = total; total += sz16;
    const size_t o_sr = total; total += sz32;
    const size_t o_sx = total; total += sz32;
    const size_t o_cat = total; total += 2 * sz16;
    const size_t o_ac = total; total += sz32;
    const size_t o_sem = total; total += sz32;
    const size_t o_hid = total; total += 2 * sz32;
    const size_t o_valid = total; total += al(static_cast<size_t>(T));
    const size_t o_ids = total; total += al(static_cast<size_t>(T) * 8);
    h->ws_samples = 0;
    if (h->ws.ensure(total)) return 1;
    B200_CUDA_OK(cudaMemset(h->ws.p, 0, h->ws.bytes));
    uint8_t* p = static_cast<uint8_t*>(h->ws.p);
    for (int i = 0; i <= kStages; ++i) {
        h->x[i] = reinterpret_cast<float*>(p + off_x[i]);
        if (i < kStages) h->hb[i] = reinterpret_cast<float*>(p + off_h[i]);
        h->a[i] = p + off_a[i];
    }
    h->s16a = p + o_s16a; h->s16b = p + o_s16b;
    h->sr = reinterpret_cast<float*>(p + o_sr); h->sx = reinterpret_cast<float*>(p + o_sx);
    h->cat16 = p + o_cat;
    h->ac32 = reinterpret_cast<float*>(p + o_ac); h->sem32 = reinterpret_cast<float*>(p + o_sem);
    h->hid32 = reinterpret_cast<float*>(p + o_hid);
    h->row_valid = p + o_valid;
    h->ids_tmp = p + o_ids;
    h->ws_samples = S;
    return 0;
}

struct EncTap {
    std::string name;
    const float* ptr;
    int64_t rows;
    int C, P;
};

}  // namespace

extern "C" {

int b200enc_create(int precision, int device, B200Enc** out) {
    B200_CHECK(out != nullptr, "b200enc_create: null argument");
    B200_CHECK(precision == B200CODEC_BF16 || precision == B200CODEC_FP16, "precision %d is not available (bf16 = 0, fp16 = 1)",
               precision);
    int ndev = 0;
    B200_CUDA_OK(cudaGetDeviceCount(&ndev));
    B200_CHECK(device >= 0 && device < ndev, "CUDA device %d not present (%d devices); this library has no CPU fallback", device, ndev);
    cudaDeviceProp prop;
    B200_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    B200_CHECK(prop.major == 10, "device %d is sm_%d%d; b200codec kernels are built for sm_100a only", device, prop.major, prop.minor);
    B200Enc* h = new B200Enc();
    h->precision = precision;
    h->device = device;
    build_specs(h);
    h->master.assign(h->specs.size(), nullptr);
    *out = h;
    return 0;
}

void b200enc_destroy(B200Enc* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (float* p : h->master)
        if (p) cudaFree(p);
    h->wbuf.release();
    h->fbuf.release();
    h->ws.release();
    for (auto& t : h->tap) t.release();
    delete h;
}

int b200enc_num_tensors(const B200Enc* h) { return h ? static_cast<int>(h->specs.size()) : 0; }
const char* b200enc_tensor_key(const B200Enc* h, int i) {
    if (!h || i < 0 || i >= static_cast<int>(h->specs.size())) return nullptr;
    return h->specs[i].key.c_str();
}
int b200enc_tensor_shape(const B200Enc* h, int i, int64_t shape_out[4]) {
    if (!h || i < 0 || i >= static_cast<int>(h->specs.size())) return -1;
    const auto& s = h->specs[i].shape;
    for (size_t d = 0; d < s.size() && d < 4; ++d) shape_out[d] = s[d];
    return static_cast<int>(s.size());
}

int b200enc_load_tensor(B200Enc* h, const char* key, const float* host_ptr, const int64_t* shape, int ndim) {
    B200_CHECK(h && key && host_ptr && shape, "b200enc_load_tensor: null argument");
    auto it = h->index.find(key);
    B200_CHECK(it != h->index.end(), "Unexpected key(s) in state_dict: \"%s\"", key);
    const TensorSpec& sp = h->specs[it->second];
    bool ok = static_cast<int>(sp.shape.size()) == ndim;
    for (int d = 0; ok && d < ndim; ++d) ok = sp.shape[d] == shape[d];
    B200_CHECK(ok, "size mismatch for %s: checkpoint tensor does not match the model shape", key);
    B200_CUDA_OK(cudaSetDevice(h->device));
    float*& dst = h->master[it->second];
    if (dst == nullptr) B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dst), sp.numel() * sizeof(float)));
    B200_CUDA_OK(cudaMemcpy(dst, host_ptr, sp.numel() * sizeof(float), cudaMemcpyHostToDevice));
    h->finalized = false;
    return 0;
}

int b200enc_finalize_weights(B200Enc* h, void* stream) {
    B200_CHECK(h != nullptr, "null handle");
    if (h->finalized) return 0;
    std::lock_guard<std::mutex> lock(h->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CUDA_OK(cudaSetDevice(h->device));
    std::string missing;
    for (size_t i = 0; i < h->specs.size(); ++i)
        if (h->master[i] == nullptr && missing.size() < 600) missing += (missing.empty() ? "\"" : ", \"") + h->specs[i].key + "\"";
    B200_CHECK(missing.empty(), "Missing key(s) in state_dict: %s", missing.c_str());

    // sizes
    size_t welems = 0, felems = 2 * 64 + 7 * 64;
    {
        int c = kGenFeatures;
        for (int i = 0; i < kStages; ++i) {
            const size_t P = pad64(c), P2 = pad64(2 * c);
            welems += 3 * (P * 7 * P + P * P) + P2 * 3 * kStrides[i] * P + 8 * 256;
            felems += 3 * (2 * P + 2 * P + 2 * P) + 2 * P + P2 + 64;
            c *= 2;
        }
        welems += static_cast<size_t>(kOutDim) * 3 * pad64(c) + 4ull * 1024 * 3 * 1024 + 2048ull * 2048 + 16 * 256;
        felems += 2 * pad64(c) + kOutDim + 4 * 1024 + 2048 + 64;
    }
    if (h->wbuf.ensure(welems * 2 + 64 * 1024)) return 1;
    (void)felems;
    if (h->fbuf.ensure(static_cast<size_t>(4) << 20)) return 1;  // ~0.2 M floats of padded vectors; 4 MB is ample
    uint8_t* wp = static_cast<uint8_t*>(h->wbuf.p);
    float* fp = static_cast<float*>(h->fbuf.p);
    auto takew = [&](size_t n) { void* r = wp; wp += (n * 2 + 255) & ~static_cast<size_t>(255); return r; };
    auto takef = [&](size_t n) { float* r = fp; fp += (n + 63) & ~static_cast<size_t>(63); return r; };
    float* scale = nullptr;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&scale), 4096 * sizeof(float)));
    const bool bf = h->precision == kPrecBf16;
    auto padvec = [&](const float* src, int n, int npad, int mode) -> float* {
        float* dst = takef(npad);
        enc_pad_vec_kernel<<<(npad + 255) / 256, 256, 0, s>>>(src, n, npad, mode, dst);
        return dst;
    };
    auto make_act = [&](const std::string& p, int C, ActW* w) -> int {
        const int P = pad64(C);
        w->alpha = padvec(h->m(p + "act.alpha"), C, P, 1);
        w->inv_beta = padvec(h->m(p + "act.beta"), C, P, 2);
        float fu[12], fd[12];
        B200_CUDA_OK(cudaMemcpy(fu, h->m(p + "upsample.filter"), sizeof(fu), cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(fd, h->m(p + "downsample.lowpass.filter"), sizeof(fd), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 12; ++i) {
            w->f.up[i] = 2.f * fu[i];  // UpSample1d multiplies its output by the ratio (filters.py:111)
            w->f.down[i] = fd[i];
        }
        return 0;
    };
    // weight-normed conv: scale = g / ||v||, operand [Npad][k * Cinpad]
    auto make_wn_conv = [&](const std::string& p, int Cout, int Cin, int k, ConvW* w) -> int {
        const int Np = pad64(Cout), Cp = pad64(Cin);
        if (launch_weightnorm_scale(h->m(p + "weight_g"), h->m(p + "weight_v"), Cout, Cin * k, scale, s)) return 1;
        w->w = takew(static_cast<size_t>(Np) * k * Cp);
        if (bf ? repack_conv<__nv_bfloat16>(h->m(p + "weight_v"), scale, w->w, Cout, Cin, k, Np, Cp, s)
               : repack_conv<__half>(h->m(p + "weight_v"), scale, w->w, Cout, Cin, k, Np, Cp, s))
            return 1;
        w->bias = padvec(h->m(p + "bias"), Cout, Np, 0);
        return 0;
    };
    auto make_plain_conv = [&](const float* wsrc, const float* bsrc, int Cout, int Cin, int k, ConvW* w) -> int {
        w->w = takew(static_cast<size_t>(Cout) * k * Cin);
        if (bf ? repack_conv<__nv_bfloat16>(wsrc, nullptr, w->w, Cout, Cin, k, Cout, Cin, s)
               : repack_conv<__half>(wsrc, nullptr, w->w, Cout, Cin, k, Cout, Cin, s))
            return 1;
        w->bias = bsrc ? padvec(bsrc, Cout, Cout, 0) : nullptr;
        return 0;
    };
    const std::string a = "acoustic_encoder.";
    {
        // conv_blocks.0: fp32 FIR weights [48][7] with the weight-norm scale folded on the host
        std::vector<float> g(kGenFeatures), v(kGenFeatures * 7), w(kGenFeatures * 7);
        B200_CUDA_OK(cudaMemcpy(g.data(), h->m(a + "conv_blocks.0.weight_g"), g.size() * 4, cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(v.data(), h->m(a + "conv_blocks.0.weight_v"), v.size() * 4, cudaMemcpyDeviceToHost));
        for (int c = 0; c < kGenFeatures; ++c) {
            double n2 = 0.0;
            for (int k = 0; k < 7; ++k) n2 += static_cast<double>(v[c * 7 + k]) * v[c * 7 + k];
            const float sc = static_cast<float>(g[c] / std::sqrt(n2));
            for (int k = 0; k < 7; ++k) w[c * 7 + k] = v[c * 7 + k] * sc;
        }
        h->conv0_w = takef(w.size());
        B200_CUDA_OK(cudaMemcpyAsync(h->conv0_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice, s));
        B200_CUDA_OK(cudaStreamSynchronize(s));
        h->conv0_b = padvec(h->m(a + "conv_blocks.0.bias"), kGenFeatures, 64, 0);
        static_assert(kGenFeatures % 4 == 0, "enc_conv0_kernel owns channel quads");
    }
    int c = kGenFeatures;
    for (int i = 0; i < kStages; ++i) {
        const std::string p = a + "conv_blocks." + std::to_string(i + 1) + ".";
        StageW& st = h->stage[i];
        for (int u = 0; u < 3; ++u) {
            const std::string q = p + "block." + std::to_string(u) + ".";
            if (make_act(q + "block.0.", c, &st.unit[u].act0)) return 1;
            if (make_wn_conv(q + "block.1.", c, c, 7, &st.unit[u].conv7)) return 1;
            if (make_act(q + "block.2.", c, &st.unit[u].act2)) return 1;
            if (make_wn_conv(q + "block.3.", c, c, 1, &st.unit[u].conv1)) return 1;
        }
        if (make_act(p + "block.3.", c, &st.act)) return 1;
        {
            const int sdn = kStrides[i], Np = pad64(2 * c), Kt = pad64(sdn * c);
            const std::string q = p + "block.4.";
            if (launch_weightnorm_scale(h->m(q + "weight_g"), h->m(q + "weight_v"), 2 * c, c * 2 * sdn, scale, s)) return 1;
            st.down.w = takew(static_cast<size_t>(Np) * 3 * Kt);
            const int pd = sdn / 2 + sdn % 2;
            if (bf ? repack_strided<__nv_bfloat16>(h->m(q + "weight_v"), scale, st.down.w, 2 * c, c, sdn, pd, Np, Kt, s)
                   : repack_strided<__half>(h->m(q + "weight_v"), scale, st.down.w, 2 * c, c, sdn, pd, Np, Kt, s))
                return 1;
            st.down.bias = padvec(h->m(q + "bias"), 2 * c, Np, 0);
        }
        c *= 2;
    }
    if (make_act(a + "conv_final_block.0.", c, &h->act_final)) return 1;
    if (make_wn_conv(a + "conv_final_block.1.", kOutDim, c, 3, &h->conv_final)) return 1;
    const std::string se = "semantic_encoder.";
    if (make_plain_conv(h->m(se + "initial_conv.weight"), nullptr, 1024, 1024, 3, &h->sem_init)) return 1;
    if (make_plain_conv(h->m(se + "residual_blocks.1.weight"), h->m(se + "residual_blocks.1.bias"), 1024, 1024, 3, &h->sem_rb1)) return 1;
    if (make_plain_conv(h->m(se + "residual_blocks.3.weight"), h->m(se + "residual_blocks.3.bias"), 1024, 1024, 3, &h->sem_rb3)) return 1;
    if (make_plain_conv(h->m(se + "final_conv.weight"), nullptr, 1024, 1024, 3, &h->sem_final)) return 1;
    if (make_plain_conv(h->m("fusion_layer.weight"), h->m("fusion_layer.bias"), 2048, 2048, 1, &h->fusion)) return 1;
    B200_CUDA_OK(cudaGetLastError());
    B200_CUDA_OK(cudaStreamSynchronize(s));
    B200_CUDA_OK(cudaFree(scale));
    h->finalized = true;
    return 0;
}

// Encoder.forward for a batch of equally long clips (encoder.py:58-78):
//   wav_dev [n_clips][n_samples] fp32 (n_samples a positive multiple of 320: Encoder.encode pads to that, :116-120),
//   w2v_dev [n_clips][T][1024] fp32 token-major, T = n_samples / 320  ->  ids [n_clips][T].
// hidden_dev [n_clips][T][2048] / acoustic_dev, semantic_dev [n_clips][T][1024] (fp32, optional) receive the fused
// hidden state and the two encoders' outputs. One launch sequence for the whole batch (80 kernels), asynchronous
// on `stream`.
int b200enc_encode_batch(B200Enc* h, const float* wav_dev, int n_clips, int64_t n_samples, const float* w2v_dev,
                         void* ids_dev, int id_type, int pre_bound, float* hidden_dev, float* acoustic_dev,
                         float* semantic_dev, void* stream) {
    B200_CHECK(h && wav_dev && w2v_dev, "b200enc_encode: null argument");
    B200_CHECK(h->finalized, "b200enc_encode called before b200enc_finalize_weights");
    B200_CHECK(n_clips > 0 && n_clips <= 65535, "encode: n_clips (%d) must be in [1, 65535]", n_clips);
    B200_CHECK(n_samples > 0 && n_samples % kHopTotal == 0 && n_samples < (1ll << 30),
               "encode: n_samples (%lld) must be a positive multiple of 320", (long long)n_samples);
    B200_CHECK(ids_dev == nullptr || id_type == 0 || id_type == 1, "encode: id_type must be 0 (int32) or 1 (int64)");
    const int S = static_cast<int>(n_samples), T = S / kHopTotal;
    const int slot_tok = T + kGapTokens;
    const int64_t total = static_cast<int64_t>(n_clips) * slot_tok * kHopTotal;  // slotted samples of the batch
    B200_CHECK(total < (1ll << 31), "encode: batch too large (%d clips x %lld samples); split it", n_clips, (long long)n_samples);
    std::lock_guard<std::mutex> lock(h->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    B200_CUDA_OK(cudaSetDevice(h->device));
    if (total > h->ws_samples) {
        B200_CUDA_OK(cudaStreamSynchronize(s));  // the old workspace may still be in use
        if (ensure_ws(h, total)) return 1;
    }
    const bool bf = h->precision == kPrecBf16;
    const int Rtok = n_clips * slot_tok;  // token-level rows

    // ---- acoustic encoder ----
    int rows = S, slot = slot_tok * kHopTotal, c = kGenFeatures;  // valid rows per clip / rows per slot at this level
    {
        const int threads = ((S + kConv0Rows - 1) / kConv0Rows) * (kGenFeatures / 4);
        B200_CUDA_OK(launch_kernel(enc_conv0_kernel, dim3((threads + 255) / 256, n_clips), dim3(256), 0, s, wav_dev, S, slot,
                                   (const float*)h->conv0_w, (const float*)h->conv0_b, kGenFeatures, h->x[0]));
    }
    h->launches++;
    auto take_tap = [&](int idx, int64_t slot_rows, int Pc) -> int {
        if (!h->taps_on) return 0;
        const size_t bytes = static_cast<size_t>(n_clips) * slot_rows * Pc * 4;
        if (h->tap[idx].ensure(bytes)) return 1;
        B200_CUDA_OK(cudaMemcpyAsync(h->tap[idx].p, h->x[idx], bytes, cudaMemcpyDeviceToDevice, s));
        h->tap_samples = n_samples;
        h->tap_clips = n_clips;
        return 0;
    };
    if (take_tap(0, slot, kGenFeatures)) return 1;
    for (int i = 0; i < kStages; ++i) {
        const StageW& st = h->stage[i];
        const int P = c;  // dense rows
        const int R = n_clips * slot;
        for (int u = 0; u < 3; ++u) {
            ENC_RUN(snake(h, h->x[i], rows, slot, n_clips, P, st.unit[u].act0, h->a[i], s));
            ENC_RUN(launch_gemm(conv_call(h, h->a[i], R, P, st.unit[u].conv7, P, 7, kDil[u], h->hb[i], true, nullptr, kActNone), s));
            ENC_RUN(snake(h, h->hb[i], rows, slot, n_clips, P, st.unit[u].act2, h->a[i], s));
            ENC_RUN(launch_gemm(conv_call(h, h->a[i], R, P, st.unit[u].conv1, P, 1, 1, h->x[i], true, h->x[i], kActNone), s));
        }
        ENC_RUN(snake(h, h->x[i], rows, slot, n_clips, P, st.act, h->a[i], s));
        const int sdn = kStrides[i], P2 = 2 * c;
        // the same operand buffer viewed as [rows / s][s * P]: a 3-tap conv over super-rows
        ENC_RUN(launch_gemm(conv_call(h, h->a[i], R / sdn, sdn * P, st.down, P2, 3, 1, h->x[i + 1], true, nullptr, kActNone), s));
        rows /= sdn;
        slot /= sdn;
        c *= 2;
        if (take_tap(i + 1, slot, P2)) return 1;
    }
    const size_t es = 2;
    // ---- semantic encoder input: cast into the slotted token rows (also writes the row-validity bytes) ----
    if (bf) B200_CUDA_OK(launch_kernel(cast_slotted_kernel<__nv_bfloat16>, dim3(Rtok), dim3(256), 0, s, w2v_dev, T, slot_tok, 1024,
                                       static_cast<__nv_bfloat16*>(h->s16a), h->row_valid));
    else B200_CUDA_OK(launch_kernel(cast_slotted_kernel<__half>, dim3(Rtok), dim3(256), 0, s, w2v_dev, T, slot_tok, 1024,
                                    static_cast<__half*>(h->s16a), h->row_valid));
    h->launches++;
    {
        const int P = c;  // 1536
        ENC_RUN(snake(h, h->x[kStages], rows, slot, n_clips, P, h->act_final, h->a[kStages], s));
        GemmCall g = conv_call(h, h->a[kStages], Rtok, P, h->conv_final, kOutDim, 3, 1, h->ac32, true, nullptr, kActNone);
        g.out16 = static_cast<uint8_t*>(h->cat16) + 1024 * es;  // right half of [semantic | acoustic]
        g.ld16 = 2048;
        ENC_RUN(launch_gemm(g, s));
    }
    // ---- semantic encoder (encoder_modules.py:121-125; ReLU(inplace=True) makes the skip carry relu(x)). The
    // outputs that feed the next 3-tap conv keep their gap rows zero (row_valid), like the decoder's halo rows ----
    {
        const uint8_t* rv = h->row_valid;
        GemmCall g0 = conv_call(h, h->s16a, Rtok, 1024, h->sem_init, 1024, 3, 1, h->sr, true, nullptr, kActRelu, rv);
        g0.out16 = h->s16b;  // r = relu(initial_conv(x)) as fp32 (skip) and as operand
        g0.ld16 = 1024;
        ENC_RUN(launch_gemm(g0, s));
        ENC_RUN(launch_gemm(conv_call(h, h->s16b, Rtok, 1024, h->sem_rb1, 1024, 3, 1, h->s16a, false, nullptr, kActRelu, rv), s));
        GemmCall g2 = conv_call(h, h->s16a, Rtok, 1024, h->sem_rb3, 1024, 3, 1, h->sx, true, h->sr, kActNone, rv);
        g2.out16 = h->s16b;
        g2.ld16 = 1024;
        ENC_RUN(launch_gemm(g2, s));
        GemmCall g3 = conv_call(h, h->s16b, Rtok, 1024, h->sem_final, 1024, 3, 1, h->sem32, true, nullptr, kActNone);
        g3.out16 = h->cat16;  // left half of [semantic | acoustic]
        g3.ld16 = 2048;
        ENC_RUN(launch_gemm(g3, s));
    }
    // ---- fusion + quantise ----
    ENC_RUN(launch_gemm(conv_call(h, h->cat16, Rtok, 2048, h->fusion, 2048, 1, 1, h->hid32, true, nullptr, kActNone), s));
    // results leave the slotted row space: one strided copy per output (clip b: rows [b * slot_tok, b * slot_tok + T))
    auto compact = [&](void* dst, const void* src, size_t row_bytes) -> int {
        B200_CUDA_OK(cudaMemcpy2DAsync(dst, static_cast<size_t>(T) * row_bytes, src, static_cast<size_t>(slot_tok) * row_bytes,
                                       static_cast<size_t>(T) * row_bytes, static_cast<size_t>(n_clips), cudaMemcpyDeviceToDevice, s));
        return 0;
    };
    if (ids_dev != nullptr) {
        ENC_RUN(launch_fsq_quantize(h->hid32, 2048, Rtok, h->m("quantizer.project_in.weight"), h->m("quantizer.project_in.bias"),
                                    2048, pre_bound, h->ids_tmp, id_type, nullptr, s));
        if (compact(ids_dev, h->ids_tmp, id_type == 1 ? 8 : 4)) return 1;
    }
    if (hidden_dev && compact(hidden_dev, h->hid32, 2048 * 4)) return 1;
    if (acoustic_dev && compact(acoustic_dev, h->ac32, 1024 * 4)) return 1;
    if (semantic_dev && compact(semantic_dev, h->sem32, 1024 * 4)) return 1;
    return 0;
}

// One clip: Encoder.forward on [1, 1, n_samples] (kept for callers of ABI 3's first encode entry).
int b200enc_encode(B200Enc* h, const float* wav_dev, int64_t n_samples, const float* w2v_dev, void* ids_dev,
                   int id_type, int pre_bound, float* hidden_dev, float* acoustic_dev, float* semantic_dev, void* stream) {
    return b200enc_encode_batch(h, wav_dev, 1, n_samples, w2v_dev, ids_dev, id_type, pre_bound, hidden_dev, acoustic_dev,
                                semantic_dev, stream);
}

// Debug / stage parity: fp32 copy of an acoustic-encoder stage of the LAST encode, [clips][rows][C] token-major
// (pad channels and gap rows dropped): "conv0" (S rows, 48), "block1".."block5" (EncoderBlock outputs: rows S/2 .. S/320).
int b200enc_read_stage(B200Enc* h, const char* name, int64_t n_samples, float* host_out, size_t n_elems, void* stream) {
    B200_CHECK(h && name && host_out, "b200enc_read_stage: null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    B200_CHECK(h->taps_on && n_samples > 0 && n_samples == h->tap_samples,
               "read_stage: enable b200enc_set_stage_taps and encode %lld samples first", (long long)n_samples);
    int idx = -1;
    if (std::strcmp(name, "conv0") == 0) idx = 0;
    else if (std::strncmp(name, "block", 5) == 0 && name[5] >= '1' && name[5] <= '5' && name[6] == 0) idx = name[5] - '0';
    B200_CHECK(idx >= 0, "read_stage: unknown stage \"%s\"", name);
    int64_t rows = n_samples, slot = (n_samples / kHopTotal + kGapTokens) * kHopTotal;
    int c = kGenFeatures;
    for (int i = 0; i < idx; ++i) {
        rows /= kStrides[i];
        slot /= kStrides[i];
        c *= 2;
    }
    const int P = c;  // dense rows
    const size_t per_clip = static_cast<size_t>(rows) * c;
    B200_CHECK(per_clip * h->tap_clips == n_elems, "read_stage %s: expected %d x %lld x %d elements, got %zu", name,
               h->tap_clips, (long long)rows, c, n_elems);
    B200_CUDA_OK(cudaSetDevice(h->device));
    B200_CUDA_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    for (int b = 0; b < h->tap_clips; ++b)
        B200_CUDA_OK(cudaMemcpy2D(host_out + b * per_clip, static_cast<size_t>(c) * 4,
                                  static_cast<const float*>(h->tap[idx].p) + static_cast<size_t>(b) * slot * P,
                                  static_cast<size_t>(P) * 4, static_cast<size_t>(c) * 4, static_cast<size_t>(rows),
                                  cudaMemcpyDeviceToHost));
    return 0;
}

int b200enc_set_stage_taps(B200Enc* h, int on) {
    B200_CHECK(h != nullptr, "null handle");
    std::lock_guard<std::mutex> lock(h->mu);
    h->taps_on = on != 0;
    if (!h->taps_on) {
        for (auto& t : h->tap) t.release();
        h->tap_samples = 0;
        h->tap_clips = 0;
    }
    return 0;
}

int64_t b200enc_launch_count(const B200Enc* h) { return h ? h->launches : 0; }

}  // extern "C"
