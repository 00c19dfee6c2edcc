// K8 on the 5th-generation tensor cores: unmasked, non-causal varlen attention with S and O in TMEM.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0, is_causal=False)
// at tts/core/codec/decoder_modules.py:283-285 (rearranges of :276-278, :287 folded into the
// addressing; head-indexed RoPE of :280-281 folded into c_attn at load time).
//
// One CTA = one (utterance, 128-query tile, head); two CTAs are co-resident per SM (80 KB smem,
// 256 TMEM columns each) so one CTA's softmax (MUFU-bound: 128x128 exponentials per tile) overlaps
// the other's MMAs.
//
//   warp 8 (lane 0)  control: TMA loads (Q once; K,V tiles of 128 keys, 2-stage ring) and
//                    tcgen05.mma issue:  S = Q K^T   (M128 N128 K64,  A,B K-major from smem)
//                                        O_j = P V   (M128 N64  K128, A = P from TENSOR MEMORY,
//                                                     B = V MN-major straight from its TMA tile)
//                    A tcgen05.mma costs ~60 ns to issue (tools/attn_trace.cu): 12 per tile. With
//                    the issuing thread inside a softmax warp that was 0.8 of a 2.25 us tile on the
//                    critical path (its warp stalled, the other seven waited at the P barrier); a
//                    warp of its own issues S_{j+1} first and P V_j while the softmax of tile j+1 runs.
//   warps 0-7        two threads per query row (= TMEM lane), 64 keys each: tcgen05.ld S, online
//                    softmax in fp32 (one smem exchange of the row max per tile), P -> bf16/fp16 -> tcgen05.st into TMEM, fold
//                    the per-tile O_j from TMEM into register accumulators with the max correction.
// Tried and measured slower: half of the exponentials as a degree-3 polynomial on the FMA pipe
// (the softmax phase is as issue-bound as it is MUFU-bound: c2 55.5 -> 58.3 us, c4 306 -> 344 us).
// Tried and measured equal: persistent CTAs (2 per SM walking the item list on one global tile
// counter, the control warp prefetching the next item's Q/K/V): 6 % faster alone on c2, nothing on
// c4, and 0.5-1 % SLOWER inside the decode step, where short-lived CTAs let the next GEMM's CTAs
// (programmatic dependent launch) start on SMs as they drain.
// TMEM -> register bandwidth is the scarce resource (ncu: identical time for very different softmax
// instruction counts): S is read once per tile, O is accumulated by the tensor core in TMEM and a
// row rescales its accumulator (tcgen05.ld / st) only when its running max actually changes.
#include "common.cuh"
#include "kernels.h"

#include <cudaTypedefs.h>

namespace b200 {

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows);

// in-kernel timeline for tools/attn_trace.cu; compiles to nothing in the product build
#ifdef B200_ATTN_TRACE
__device__ unsigned long long* g_attn_trace = nullptr;  // [CTAs][64] globaltimer stamps of thread 0 (softmax) / 256 (control)
#define ATTN_TRACE(slot)                                                                          \
    do {                                                                                          \
        if ((threadIdx.x == 0 || threadIdx.x == 256) && g_attn_trace != nullptr) {                \
            unsigned long long t_;                                                                \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                 \
            g_attn_trace[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 64 + (slot)] = t_; \
        }                                                                                         \
    } while (0)
#else
#define ATTN_TRACE(slot) do { } while (0)
#endif

namespace {

constexpr int kD = 64;            // head dim
constexpr int kBQ = 128;          // queries per CTA
constexpr int kBK = 128;          // keys per tile
constexpr int kSoftmaxThreads = 256;  // two threads per query row
constexpr int kThreads = kSoftmaxThreads + 32;  // + the control warp (TMA / MMA issue)
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 halfs
constexpr uint32_t kTmemCols = 256;    // S: 128 columns, O_j: 64, P (16-bit pairs): 64

struct AttnSmem {
    static constexpr int kQ = 0;
    static constexpr int kK = kTileBytes;                  // 2 stages
    static constexpr int kV = kK + 2 * kTileBytes;         // 2 stages
    static constexpr int kBar = kV + 2 * kTileBytes;
    static constexpr int kRed = kBar + 16 * 8 + 16;           // row-max / row-sum exchange, 2 KB
    static constexpr int kTotal = kRed + 2 * 2 * 128 * 4 + 1024;  // + alignment slack
};

// B operand descriptor for V: [128 keys x 64 d] tile as TMA SWIZZLE_128B wrote it (one 128-byte
// row per key). As the (N = d) x (K = keys) operand it is MN-major: d is contiguous, 8 keys form
// one 1024-byte swizzle atom (stride byte offset), a single atom spans all 64 d (LBO unused).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // LBO (unused: one atom along MN)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO: next group of 8 keys
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // not volatile: free to schedule
    return y;
}

// Work item {q_row, n_q, kv_row0, T_kv}: queries are rows [q_row, q_row + n_q) of the Q tensor (n_q <= 128), keys /
// values rows [kv_row0, kv_row0 + T_kv) of the K/V tensor. One-shot decode: both are the c_attn output
// (K at column D, V at 2 D) and kv_row0 / T_kv are the utterance; cached streaming (codec.cu, B200Stream): K/V
// come from a per-layer ring of the last tokens' keys and values (K at column 0, V at D) -- the softmax does
// not care about the order of the keys, and the head-indexed rotary embedding carries no position.
template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
attention_tc05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                      T* __restrict__ out, const int4* __restrict__ work, int heads, int k_col0, int v_col0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + AttnSmem::kBar);
    uint64_t* bar_kv_full = bar_q + 1;   // [2]
    uint64_t* bar_s_full = bar_q + 3;
    uint64_t* bar_p_full = bar_q + 4;
    uint64_t* bar_o_full = bar_q + 5;
    uint64_t* bar_s_read = bar_q + 6;    // every softmax thread holds S_j in registers: S may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 8);
    float* red = reinterpret_cast<float*>(smem + AttnSmem::kRed);  // [2 parities][2 halves][128 rows]

    pdl_launch_dependents();
    ATTN_TRACE(0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int half = warp >> 2;               // which 64 keys of a tile / which 32 output columns
    const int r = (warp & 3) * 32 + lane;     // query row == TMEM lane
    const bool ctrl_warp = warp == 8;         // control warp; warps 0-7 are softmax warps
    const bool ctrl = ctrl_warp && lane == 0;
    const int4 wk = work[blockIdx.x];
    const int q_row = wk.x, n_q = wk.y, row0 = wk.z, T_utt = wk.w;  // row0 / T_utt: the keys
    const int head = blockIdx.y;
    const int D = heads * kD;
    const int n_kv = (T_utt + kBK - 1) / kBK;

    if (ctrl) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        mbar_init(bar_q, 1);
        mbar_init(&bar_kv_full[0], 1);
        mbar_init(&bar_kv_full[1], 1);
        mbar_init(bar_s_full, 1);
        mbar_init(bar_p_full, kSoftmaxThreads);
        mbar_init(bar_o_full, 1);
        mbar_init(bar_s_read, kSoftmaxThreads);
        fence_barrier_init();
    }
    __syncwarp();
    if (ctrl_warp) tmem_alloc<kTmemCols>(tmem_slot);
    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    pdl_wait();  // prologue above overlaps the predecessor's tail; no global access before this point
    ATTN_TRACE(1);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_addr + half * 64;        // this thread's 64 score columns
    const uint32_t tmem_o = tmem_base + lane_addr + 128 + half * 32;  // this thread's 32 output columns
    const uint32_t tmem_p = tmem_base + lane_addr + 192 + half * 32;  // its 64 keys as 32 packed columns

    uint8_t* sQ = smem + AttnSmem::kQ;
    uint8_t* sK = smem + AttnSmem::kK;
    uint8_t* sV = smem + AttnSmem::kV;

    constexpr uint32_t idesc_s = umma_idesc(UmmaFmt<T>::value, 128, kBK);              // K-major B
    constexpr uint32_t idesc_o = umma_idesc(UmmaFmt<T>::value, 128, kD) | (1u << 16);  // MN-major B
    auto load_kv = [&](int j) {
        const int s = j & 1;
        mbar_arrive_expect_tx(&bar_kv_full[s], 2 * kTileBytes);
        tma_load_2d(sK + s * kTileBytes, &tmap_kv, &bar_kv_full[s], k_col0 + head * kD, row0 + j * kBK);
        tma_load_2d(sV + s * kTileBytes, &tmap_kv, &bar_kv_full[s], v_col0 + head * kD, row0 + j * kBK);
    };
    auto issue_s = [&](int j) {
        const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sQ));
        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sK + (j & 1) * kTileBytes));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
            umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc_s, k != 0);
        umma_commit(bar_s_full);
    };
    if (ctrl_warp) {
        if (ctrl) {
            mbar_arrive_expect_tx(bar_q, kTileBytes);
            tma_load_2d(sQ, &tmap_q, bar_q, head * kD, q_row);
            load_kv(0);
            if (n_kv > 1) load_kv(1);
            mbar_wait(bar_q, 0);
            mbar_wait(&bar_kv_full[0], 0);
            tc05_fence_after();
            ATTN_TRACE(2);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                // S_{j+1} only needs S_j out of tensor memory (it sits in the softmax threads'
                // registers as soon as their tcgen05.ld completes), not P_j: issue it during the
                // exponentials, so that it is ready when the softmax warps come back for it
                if (j + 1 < n_kv) {
                    mbar_wait(bar_s_read, j & 1);
                    tc05_fence_after();
                    mbar_wait(&bar_kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    tc05_fence_after();
                    issue_s(j + 1);
                }
                // every row's P_j is in TMEM and O rescaled: O += P_j V_j
                mbar_wait(bar_p_full, j & 1);
                tc05_fence_after();
                if (j < 6) ATTN_TRACE(8 + j * 8 + 5);  // every row's P_j stored
                const uint32_t v_base = smem_u32(sV + (j & 1) * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    // P from TMEM (16 keys = 8 columns per step); V MN-major (16 rows of 128 B per step)
                    const uint64_t b_desc = umma_desc_mn_sw128(v_base + k * 16 * 128);
                    umma_f16_ts(tmem_base + 128, tmem_base + 192 + k * 8, b_desc, idesc_o, (j | k) != 0);
                }
                umma_commit(bar_o_full);
                if (j < 6) ATTN_TRACE(8 + j * 8 + 6);  // S_{j+1} and P V_j issued
                if (j + 2 < n_kv) {
                    // K/V stage j & 1 is free once P V_j has retired: refill it with tile j + 2
                    mbar_wait(bar_o_full, j & 1);
                    load_kv(j + 2);
                }
            }
        }
        __syncwarp();
    } else {
    // ---------------- softmax: two threads per query row, 64 keys each ----------------
    // TMEM -> register bandwidth (~64 B/clk/SM) is the scarce resource: S is read exactly once per
    // tile and kept in registers, and O is accumulated by the tensor core inside TMEM; a row only
    // touches its O when its running max changes (rare after the first tiles) to rescale it.
    const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m_run = -INFINITY;  // running max of the raw scores (identical in both threads of a row)
    float l_run = 0.f;        // this thread's share of the running sum
    for (int j = 0; j < n_kv; ++j) {
        const int n_valid = min(kBK, T_utt - j * kBK) - half * 64;  // valid keys among this thread's 64
        const bool full = n_valid >= 64;
        mbar_wait(bar_s_full, j & 1);
        tc05_fence_after();
        if (j < 6) ATTN_TRACE(8 + j * 8 + 0);  // S_j ready
        uint32_t sr[64];
        tmem_ld_32x32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
        tmem_ld_32x32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
        tmem_ld_wait();
        tc05_fence_before();
        mbar_arrive(bar_s_read);
        if (j < 6) ATTN_TRACE(8 + j * 8 + 1);  // S_j in registers
        if (!full) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
                if (i >= n_valid) sr[i] = 0xff800000u;  // -inf: exp2 gives exactly 0
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
            mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
        }
        float* red_j = red + (j & 1) * 256;
        red_j[half * 128 + r] = fmaxf(mx0, mx1);
        // only the two warps that share this lane quarter exchange data: 4 independent 64-thread
        // named barriers instead of one CTA-wide barrier per tile
        asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
        if (j < 6) ATTN_TRACE(8 + j * 8 + 2);  // row max exchanged
        const float m_new = fmaxf(m_run, fmaxf(red_j[r], red_j[128 + r]));
        const float corr = ex2_approx((m_run - m_new) * c);  // first tile: ex2(-inf) = 0
        const float m_scaled = m_new * c;
        if (j > 0) {
            // P V_{j-1} retired: P, its K/V stage and the O accumulator are quiescent
            mbar_wait(bar_o_full, (j - 1) & 1);
            tc05_fence_after();
            // tcgen05.ld/st are .sync.aligned: the decision must be warp-uniform, so a warp rescales
            // when ANY of its 32 rows moved its max (the others multiply by exactly 1)
            if (__any_sync(0xffffffffu, corr != 1.f)) {
                // two 16-column halves: only 16 registers live next to the 64 scores (2 CTAs per SM
                // leave 96 registers per thread)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t orr[16];
                    tmem_ld_32x16(tmem_o + hf * 16, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                    tmem_st_32x16(tmem_o + hf * 16, orr);
                }
            }
        }
        if (j < 6) ATTN_TRACE(8 + j * 8 + 3);  // P V_{j-1} retired, O rescaled
        l_run *= corr;
        m_run = m_new;
        // P = exp2(s * c - m * c) -> 16-bit pairs -> TMEM (A operand of P V)
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(sr[ch * 16 + 2 * i]), c, -m_scaled));
                const float p1 = ex2_approx(fmaf(__uint_as_float(sr[ch * 16 + 2 * i + 1]), c, -m_scaled));
                if (i & 1) { l2 += p0; l3 += p1; } else { l0 += p0; l1 += p1; }
                pk[i] = Half16<T>::pack(p0, p1);
            }
            // 16 keys -> 8 packed columns of this row's lane
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                    tmem_p + ch * 8),
                "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                : "memory");
        }
        l_run += (l0 + l1) + (l2 + l3);
        tmem_st_wait();
        tc05_fence_before();
        mbar_arrive(bar_p_full);
        if (j < 6) ATTN_TRACE(8 + j * 8 + 4);  // own P_j stored
    }
    // the finished accumulator and the two halves of the row sum
    mbar_wait(bar_o_full, (n_kv - 1) & 1);
    tc05_fence_after();
    ATTN_TRACE(3);  // last P V retired
    uint32_t orr[32];
    tmem_ld_32x32(tmem_o, orr);
    tmem_ld_wait();
    float* red_l = red + (n_kv & 1) * 256;  // the parity not used by the last tile's max exchange
    red_l[half * 128 + r] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
    if (r < n_q) {
        const float inv = 1.f / (red_l[r] + red_l[128 + r]);
        T* dst = out + static_cast<size_t>(q_row + r) * D + head * kD + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = Half16<T>::pack(__uint_as_float(orr[8 * i + 0]) * inv, __uint_as_float(orr[8 * i + 1]) * inv);
            u.y = Half16<T>::pack(__uint_as_float(orr[8 * i + 2]) * inv, __uint_as_float(orr[8 * i + 3]) * inv);
            u.z = Half16<T>::pack(__uint_as_float(orr[8 * i + 4]) * inv, __uint_as_float(orr[8 * i + 5]) * inv);
            u.w = Half16<T>::pack(__uint_as_float(orr[8 * i + 6]) * inv, __uint_as_float(orr[8 * i + 7]) * inv);
            reinterpret_cast<uint4*>(dst)[i] = u;
        }
    }

    }  // softmax warps
    ATTN_TRACE(4);  // output stored
    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    if (ctrl_warp) tmem_dealloc<kTmemCols>(tmem_base);
    ATTN_TRACE(5);
}

}  // namespace

namespace {
// keys / values of each stream's NEW rows -> its slots of the layer's ring (cached streaming): row
// utt_row0[u] + overlap + i of the c_attn output, columns [D, 3 D), goes to ring row u * cap + (wpos + i) % cap
template <typename T>
__global__ void __launch_bounds__(256)
kv_scatter_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ utt_row0, int overlap, int n_new, int D,
                  T* __restrict__ ring, int cap, int wpos) {
    pdl_launch_dependents();
    pdl_wait();
    const int u = blockIdx.y, i = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(utt_row0[u] + overlap + i) * 3 * D + D);
    uint4* dst = reinterpret_cast<uint4*>(ring + (static_cast<size_t>(u) * cap + (wpos + i) % cap) * 2 * D);
    for (int k = threadIdx.x; k < 2 * D / 8; k += blockDim.x) dst[k] = src[k];
}
}  // namespace

int launch_kv_scatter(int prec, const void* qkv, const int32_t* utt_row0, int n_streams, int overlap, int n_new, int D,
                      void* ring, int cap, int wpos, cudaStream_t stream) {
    if (n_streams <= 0 || n_new <= 0) return 0;
    dim3 grid(n_new, n_streams);
    if (prec == kPrecBf16)
        B200_CUDA_OK(launch_kernel(kv_scatter_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream,
                                   static_cast<const __nv_bfloat16*>(qkv), utt_row0, overlap, n_new, D,
                                   static_cast<__nv_bfloat16*>(ring), cap, wpos));
    else
        B200_CUDA_OK(launch_kernel(kv_scatter_kernel<__half>, grid, dim3(256), 0, stream, static_cast<const __half*>(qkv),
                                   utt_row0, overlap, n_new, D, static_cast<__half*>(ring), cap, wpos));
    return 0;
}

int launch_attention_tc05_ex(int prec, const void* q, int q_rows, int q_ld, const void* kv, int kv_rows, int kv_ld,
                             int k_col0, int v_col0, const int4* work, int n_work, int heads, void* out,
                             cudaStream_t stream) {
    if (n_work <= 0) return 0;
    B200_CHECK(prec == kPrecBf16 || prec == kPrecFp16, "attention: unsupported precision %d", prec);
    alignas(64) CUtensorMap tq, tkv;
    if (make_tmap_2d(&tq, q, prec == kPrecBf16 ? 0 : 1, q_rows, q_ld, q_ld, 128)) return 1;
    if (make_tmap_2d(&tkv, kv, prec == kPrecBf16 ? 0 : 1, kv_rows, kv_ld, kv_ld, 128)) return 1;
    dim3 grid(n_work, heads);
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__nv_bfloat16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__half>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
        // two CTAs per SM only fit with the full 228 KB shared-memory carveout
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__nv_bfloat16>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__half>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
    }
    if (prec == kPrecBf16)
        B200_CUDA_OK(launch_kernel(attention_tc05_kernel<__nv_bfloat16>, grid, dim3(kThreads), AttnSmem::kTotal, stream,
                                   tq, tkv, static_cast<__nv_bfloat16*>(out), work, heads, k_col0, v_col0));
    else
        B200_CUDA_OK(launch_kernel(attention_tc05_kernel<__half>, grid, dim3(kThreads), AttnSmem::kTotal, stream, tq, tkv,
                                   static_cast<__half*>(out), work, heads, k_col0, v_col0));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_attention_tc05(int prec, const void* qkv, const RowSpace& rs, int heads, void* out,
                          cudaStream_t stream) {
    const int D = heads * kD;
    return launch_attention_tc05_ex(prec, qkv, rs.rows, 3 * D, qkv, rs.rows, 3 * D, D, 2 * D, rs.attn128_work,
                                    rs.n_attn128_work, heads, out, stream);
}

}  // namespace b200
