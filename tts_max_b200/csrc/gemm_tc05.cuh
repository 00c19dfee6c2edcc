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
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      MMA issuer    (one elected thread, tcgen05.mma, fp32 accum in TMEM)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue      (tcgen05.ld -> bias/act/residual/mask -> global)
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the
// mainloop of tile i+1.
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
constexpr int kGemmBlockM = 128;
constexpr int kGemmThreads = 256;

template <int BLOCK_N, int kStages>
struct GemmSmem {
    static constexpr int kABytes = kGemmBlockM * 128;  // 128 rows x 128 B (one swizzle span)
    static constexpr int kBBytes = BLOCK_N * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16 + 1024;  // + align slack
};

template <int BLOCK_N, typename InT, typename OutT, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using SM = GemmSmem<BLOCK_N, kStages>;
    constexpr int BLOCK_K = 128 / sizeof(InT);   // elements per 128-byte swizzle row
    constexpr int UMMA_K = 32 / sizeof(InT);     // 16 for bf16 / fp16
    static_assert(sizeof(InT) == 2, "operands are bf16 or fp16");
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int num_n = (p.n_store + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.taps * p.k_blocks_per_tap;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    pdl_wait();  // everything above overlaps the predecessor's tail
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                // n fastest: the clusters that share one A (activation) tile run concurrently, so
                // it is fetched from HBM once; the weight tiles are few and stay in L2
                const int n_blk = tile % num_n;
                const int m_blk = tile / num_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], SM::kStageBytes);
                    const int tap = kb / p.k_blocks_per_tap;
                    const int kc = kb - tap * p.k_blocks_per_tap;
                    uint8_t* sa = smem + stage * SM::kStageBytes;
                    tma_load_2d(sa, &tmap_a, &full_bar[stage], kc * BLOCK_K,
                                m_blk * kGemmBlockM + tap * p.tap_dil - p.tap_pad);
                    tma_load_2d(sa + SM::kABytes, &tmap_b, &full_bar[stage], kb * BLOCK_K,
                                n_blk * BLOCK_N);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------- MMA issuer -------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(UmmaFmt<InT>::value, kGemmBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
                    const uint64_t a_desc = umma_desc_k_sw128(sa);
                    const uint64_t b_desc = umma_desc_k_sw128(sa + SM::kABytes);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 32 bytes (= UMMA_K elements) inside the swizzle span: +2 in >>4 units
                        umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // -------------------------------- epilogue --------------------------------
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_blk = tile % num_n;
            const int m_blk = tile / num_n;
            const int row = m_blk * kGemmBlockM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const bool row_zero = row_ok && p.row_valid != nullptr && p.row_valid[row] == 0;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc05_fence_after();
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                const int n0 = n_blk * BLOCK_N + c * 32;
                if (n0 >= p.n_store) break;
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                  static_cast<uint32_t>(acc * BLOCK_N + c * 32),
                              r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x;
                        v[4 * j + 1] += b.y;
                        v[4 * j + 2] += b.z;
                        v[4 * j + 3] += b.w;
                    }
                }
                if (p.act == kActSilu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));  // MUFU.EX2 + MUFU.RCP
                } else if (p.act == kActRelu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (row_ok) {
                    if (p.residual != nullptr) {
                        const float4* r4 = reinterpret_cast<const float4*>(
                            p.residual + static_cast<size_t>(row) * p.ld_res + n0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = r4[j];
                            v[4 * j + 0] += b.x;
                            v[4 * j + 1] += b.y;
                            v[4 * j + 2] += b.z;
                            v[4 * j + 3] += b.w;
                        }
                    }
                    if (row_zero) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                    OutT* dst = out + static_cast<size_t>(row) * p.ldc + n0;
                    if constexpr (sizeof(OutT) == 4) {
                        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            u.x = Half16<OutT>::pack(v[8 * j + 0], v[8 * j + 1]);
                            u.y = Half16<OutT>::pack(v[8 * j + 2], v[8 * j + 3]);
                            u.z = Half16<OutT>::pack(v[8 * j + 4], v[8 * j + 5]);
                            u.w = Half16<OutT>::pack(v[8 * j + 6], v[8 * j + 7]);
                            d4[j] = u;
                        }
                    }
                }
            }
            tc05_fence_before();
            mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

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

template <int BLOCK_N, typename InT, typename OutT, int kStages>
int launch_gemm_tc05(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                     cudaStream_t stream) {
    using SM = GemmSmem<BLOCK_N, kStages>;
    auto kern = gemm_tc05_kernel<BLOCK_N, InT, OutT, kStages>;
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SM::kTotal));
    }
    const int num_m = (p.M + kGemmBlockM - 1) / kGemmBlockM;
    const int num_n = (p.n_store + BLOCK_N - 1) / BLOCK_N;
    int grid = num_m * num_n;
    if (grid > kNumSMs) grid = kNumSMs;
    if (grid < 1) return 0;
    B200_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(kGemmThreads), SM::kTotal, stream, ta, tb, p));
    return 0;
}

}  // namespace b200
