// tcgen05 GEMM / implicit-GEMM conv1d for sm_100a.
//
//   out[m, n] = epilogue( sum_{tap, c} A[m + tap - tap_pad, c] * W[n, tap*Cin + c] )
//
// A is a token-major (channels-last) activation matrix [rows, Cin]; W is the
// torch Linear weight [N, K] (K-major) or a Conv1d weight repacked to
// [Cout, taps*Cin]. With taps == 1 this is a plain Linear; with taps == 3/7 it is
// Conv1d(k=3/7, padding="same") evaluated as `taps` row-shifted K-slabs, the
// halo coming from TMA out-of-bounds zero fill (array edges) or from zero rows
// kept between utterances in the padded row space (see codec.cu).
//
// Replaces (reference, all dispatched to ATen/cuBLAS/cuDNN there):
//   fc_post_a            tts/core/codec/decoder.py:63,79
//   backbone.embed       tts/core/codec/decoder_modules.py:340,392
//   ResnetBlock conv1/2  tts/core/codec/decoder_modules.py:185-193,206,215
//   c_attn / c_proj      tts/core/codec/decoder_modules.py:272-273,276,288
//   mlp.fc1 / fc2        tts/core/codec/decoder_modules.py:243-251
//   head.out             tts/core/codec/decoder_modules.py:112,130
//
// This header holds the contract (GemmParams) and the trace hooks; the kernel is the CTA-pair one in
// gemm_tc05_2cta.cuh (the 1-CTA 128 x 128 kernel of round 1 is gone: every shape of the decode, the
// 48 kHz variant and the encoder tiles as 256 x {64, 128, 192, 256}).
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace b200 {

struct GemmParams {
    int M;                 // rows of A / out
    int n_store;           // columns [0, n_store) are written; multiple of 32
    int k_blocks_per_tap;  // Cin / BLOCK_K
    int taps;              // 1 (linear), 3 or 7 (conv)
    int tap_pad;           // A row = m + tap * tap_dil - tap_pad
    int tap_dil;           // dilation (rows between taps)
    void* out;
    int ldc;               // elements
    const float* bias;     // [n_store] or nullptr
    const float* residual; // fp32 [M, ld_res] or nullptr (added after activation)
    int ld_res;
    const uint8_t* row_valid;  // [M] or nullptr; rows flagged 0 are written as zeros
    int act;
    // ---- RMSNorm fusion (CTA-pair kernel only) ----
    // consumer: out = act(rstd[m] * acc + bias), rstd[m] = rsqrt(sum_s ss_in[m][s] * ss_inv_dim + ss_eps)
    const float* ss_in;    // [M][32] partial sums of x^2 (one per 32 columns of the 1024-wide x) or nullptr
    float ss_inv_dim;
    float ss_eps;
    // producer (fp32-output GEMMs): also write the 16-bit copy of the result and its row partial sums
    void* out16;           // [M, ld16] operand dtype or nullptr
    int ld16;
    float* ss_out;         // [M][32] or nullptr; slot = column / 32 (requires N == 1024)
    // fp16 operands: the copy is stored as x * out16_scale (a power of two, so un-normalised activations
    // cannot leave fp16's range) and the consumer multiplies its row scale by ss_in_scale = 1 / out16_scale
    // (RMSNorm is scale-invariant; both factors are exact). 1 for bf16.
    float out16_scale;
    float ss_in_scale;
    // CTA-pair kernel: which outputs exist (their addresses travel in tensor maps)
    int has32;             // fp32 `out` (+ optional residual)
    int has16;             // 16-bit `out` (has32 == 0) or the 16-bit copy `out16` (has32 == 1)
    // CTA-pair kernel, GroupNorm(32 groups) statistics of the fp32 result (N == 1024: a 32-column
    // epilogue chunk is one group): stats[(utt * 32 + group) * 2 + {0, 1}] += {sum, sum of squares}
    double* gn_stats;            // nullptr = off
    const int32_t* gn_row_utt;   // [M] utterance of each row, -1 on halo rows
#ifdef B200_GEMM_TRACE
    unsigned long long* trace;  // tools/gemm_trace.cu only: [gridDim.x][128] timestamps
#endif
};

// in-kernel timeline for tools/gemm_trace.cu; compiles to nothing in the product build
#ifdef B200_GEMM_TRACE
#define B200_TRACE(slot)                                                                     \
    do {                                                                                     \
        if (p.trace != nullptr) {                                                            \
            unsigned long long t_;                                                           \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                            \
            p.trace[static_cast<size_t>(blockIdx.x) * 128 + (slot)] = t_;                     \
        }                                                                                    \
    } while (0)
#else
#define B200_TRACE(slot) do { } while (0)
#endif

constexpr int kGemmSsSlots = 32;  // row sum-of-squares partials per row (fused RMSNorm)
constexpr int kGemmBlockM = 128;  // rows of A per CTA (a CTA pair covers 256)

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
enum TmapDtype : int { kTmapBf16 = 0, kTmapF16 = 1, kTmapF32 = 2 };

// Encode a 2-D row-major [rows, cols] tensor map with a {128 bytes, box_rows} box
// and the 128-byte swizzle. Returns 0 on success.
int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows);
// General form: a {box_cols elements, box_rows} box whose row is 32, 64 or 128 bytes, swizzled
// with the pattern of the same width. Encodings are memoised per thread (the decoder re-issues
// the same few hundred maps every step; cuTensorMapEncodeTiled costs about a microsecond).
int make_tmap_box(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                  uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows);

}  // namespace b200
