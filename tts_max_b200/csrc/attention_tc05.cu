// K8 on the 5th-generation tensor cores: unmasked, non-causal varlen attention with S and O in TMEM.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0, is_causal=False)
// at tts/core/codec/decoder_modules.py:283-285 (rearranges of :276-278, :287 folded into the
// addressing; head-indexed RoPE of :280-281 folded into c_attn at load time).
//
// One CTA = one (utterance, 128-query tile, head); two CTAs are co-resident per SM (112 KB smem,
// 256 TMEM columns each) so one CTA's softmax (MUFU-bound: 128x128 exponentials per tile) overlaps
// the other's MMAs.
//
//   warp 4, lane 0   control thread: TMA loads (Q once; K,V tiles of 128 keys, 2-stage ring) and
//                    tcgen05.mma issue:  S = Q K^T   (M128 N128 K64,  A,B K-major from smem)
//                                        O_j = P V   (M128 N64  K128, A = P K-major from smem,
//                                                     B = V MN-major straight from its TMA tile)
//   warps 0-3        one thread per query row (= TMEM lane): tcgen05.ld S, online softmax in fp32
//                    with no cross-thread reduction, P -> bf16/fp16 -> 128B-swizzled smem, fold
//                    the per-tile O_j from TMEM into register accumulators with the max correction.
// O is not accumulated in TMEM across tiles: every P V product lands in TMEM fresh and is folded
// into registers (64 fp32 per thread) with the running-max rescale, so no TMEM read-modify-write.
#include "common.cuh"
#include "kernels.h"

#include <cudaTypedefs.h>

namespace b200 {

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows);

namespace {

constexpr int kD = 64;            // head dim
constexpr int kBQ = 128;          // queries per CTA
constexpr int kBK = 128;          // keys per tile
constexpr int kThreads = 160;     // 4 softmax warps + 1 control warp
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 halfs
constexpr uint32_t kTmemCols = 256;    // S: 128, O_j: 64 (power of two >= 192)

struct AttnSmem {
    static constexpr int kQ = 0;
    static constexpr int kK = kTileBytes;                  // 2 stages
    static constexpr int kV = kK + 2 * kTileBytes;         // 2 stages
    static constexpr int kP = kV + 2 * kTileBytes;         // 2 K-atoms of [128 rows x 128 B]
    static constexpr int kBar = kP + 2 * kTileBytes;
    // no alignment slack: two CTAs must fit in one SM's 228 KB; the kernel checks the base instead
    static constexpr int kTotal = kBar + 16 * 8 + 16;
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
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
attention_tc05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, T* __restrict__ out,
                      const int4* __restrict__ work, int heads) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte alignment
        if (threadIdx.x == 0) printf("b200codec: attention smem base is not 1024-byte aligned\n");
        __trap();
    }
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + AttnSmem::kBar);
    uint64_t* bar_kv_full = bar_q + 1;   // [2]
    uint64_t* bar_kv_free = bar_q + 3;   // [2]
    uint64_t* bar_s_full = bar_q + 5;
    uint64_t* bar_p_full = bar_q + 6;
    uint64_t* bar_o_full = bar_q + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int4 wk = work[blockIdx.x];
    const int row0 = wk.x, T_utt = wk.y, q0 = wk.z;
    const int head = blockIdx.y;
    const int D = heads * kD;
    const int n_kv = (T_utt + kBK - 1) / kBK;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_q, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_kv_full[s], 1);
            mbar_init(&bar_kv_free[s], 1);
        }
        mbar_init(bar_s_full, 1);
        mbar_init(bar_p_full, 128);
        mbar_init(bar_o_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;        // 128 columns
    const uint32_t tmem_o = tmem_base + 128;  // 64 columns

    uint8_t* sQ = smem + AttnSmem::kQ;
    uint8_t* sK = smem + AttnSmem::kK;
    uint8_t* sV = smem + AttnSmem::kV;
    uint8_t* sP = smem + AttnSmem::kP;

    if (warp == 4) {
        // ------------------------------ control thread ------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc(UmmaFmt<T>::value, 128, kBK);          // K-major B
            constexpr uint32_t idesc_o = umma_idesc(UmmaFmt<T>::value, 128, kD) | (1u << 16);  // MN-major B
            auto load_kv = [&](int j) {
                const int s = j & 1;
                mbar_arrive_expect_tx(&bar_kv_full[s], 2 * kTileBytes);
                tma_load_2d(sK + s * kTileBytes, &tmap_qkv, &bar_kv_full[s], D + head * kD, row0 + j * kBK);
                tma_load_2d(sV + s * kTileBytes, &tmap_qkv, &bar_kv_full[s], 2 * D + head * kD, row0 + j * kBK);
            };
            auto issue_s = [&](int j) {
                const int s = j & 1;
                const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sQ));
                const uint64_t b_desc = umma_desc_k_sw128(smem_u32(sK + s * kTileBytes));
#pragma unroll
                for (int k = 0; k < kD / 16; ++k)
                    umma_f16_ss(tmem_s, a_desc + 2 * k, b_desc + 2 * k, idesc_s, k != 0);
                umma_commit(bar_s_full);
            };
            mbar_arrive_expect_tx(bar_q, kTileBytes);
            tma_load_2d(sQ, &tmap_qkv, bar_q, head * kD, row0 + q0);
            load_kv(0);
            if (n_kv > 1) load_kv(1);
            mbar_wait(bar_q, 0);
            mbar_wait(&bar_kv_full[0], 0);
            tc05_fence_after();
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                mbar_wait(bar_p_full, j & 1);  // P_j in smem, S_j fully read
                tc05_fence_after();
                if (j + 1 < n_kv) {
                    mbar_wait(&bar_kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    tc05_fence_after();
                    issue_s(j + 1);  // overlaps the softmax threads' fold of O_j
                }
                // O_j = P_j V_j : K = 128 keys = 8 steps of 16
                const uint32_t p_base = smem_u32(sP);
                const uint32_t v_base = smem_u32(sV + s * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    // P: K-major, two 64-key atoms of 16 KB; 32 B per step inside an atom
                    const uint64_t a_desc = umma_desc_k_sw128(p_base + (k >> 2) * kTileBytes + (k & 3) * 32);
                    // V: MN-major, 16 keys = 16 rows of 128 B per step
                    const uint64_t b_desc = umma_desc_mn_sw128(v_base + k * 16 * 128);
                    umma_f16_ss(tmem_o, a_desc, b_desc, idesc_o, k != 0);
                }
                umma_commit(bar_o_full);
                umma_commit(&bar_kv_free[s]);
                if (j + 2 < n_kv) {
                    mbar_wait(&bar_kv_free[s], (j >> 1) & 1);  // P V_j retired: stage s is free
                    load_kv(j + 2);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------ softmax: one thread per query row ------------------------
        const int r = threadIdx.x;  // 0..127 == TMEM lane == row of the Q tile
        const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
        const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        float o[kD];
#pragma unroll
        for (int i = 0; i < kD; ++i) o[i] = 0.f;
        float m_run = -INFINITY;  // running max of the raw scores
        float l_run = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            const int kv0 = j * kBK;
            const int n_valid = min(kBK, T_utt - kv0);
            mbar_wait(bar_s_full, j & 1);
            tc05_fence_after();
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t sr[32];
                tmem_ld_32x32(tmem_s + lane_addr + ch * 32, sr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = __uint_as_float(sr[i]);
                    if (ch * 32 + i < n_valid) mx = fmaxf(mx, v);
                }
            }
            const float m_new = fmaxf(m_run, mx);
            const float corr = ex2_approx((m_run - m_new) * c);  // first tile: ex2(-inf) = 0
            const float m_scaled = m_new * c;
            // fold the previous tile's O_{j-1} (computed against m_run) before P_{j} overwrites
            // the P buffer: waiting for it also guarantees P V_{j-1} has finished reading P.
            if (j > 0) {
                mbar_wait(bar_o_full, (j - 1) & 1);
                tc05_fence_after();
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    uint32_t orr[32];
                    tmem_ld_32x32(tmem_o + lane_addr + ch * 32, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[ch * 32 + i] += __uint_as_float(orr[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < kD; ++i) o[i] *= corr;
            l_run *= corr;
            m_run = m_new;
            // pass 2: P = exp2(s * c - m * c) -> 16-bit -> swizzled smem (A operand of P V)
            float l_add = 0.f;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t sr[32];
                tmem_ld_32x32(tmem_s + lane_addr + ch * 32, sr);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(sr[2 * i]), c, -m_scaled));
                    float p1 = ex2_approx(fmaf(__uint_as_float(sr[2 * i + 1]), c, -m_scaled));
                    if (ch * 32 + 2 * i >= n_valid) p0 = 0.f;
                    if (ch * 32 + 2 * i + 1 >= n_valid) p1 = 0.f;
                    l_add += p0 + p1;
                    pk[i] = Half16<T>::pack(p0, p1);
                }
                // 32 keys = 4 chunks of 16 B; key chunk index within its 64-key atom: (ch & 1) * 4 + q
                uint8_t* prow = sP + (ch >> 1) * kTileBytes + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((ch & 1) * 4 + q) ^ (r & 7);
                    *reinterpret_cast<uint4*>(prow + chunk * 16) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run += l_add;
            fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core
            tc05_fence_before();
            mbar_arrive(bar_p_full);
        }
        // last tile's O
        mbar_wait(bar_o_full, (n_kv - 1) & 1);
        tc05_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            uint32_t orr[32];
            tmem_ld_32x32(tmem_o + lane_addr + ch * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[ch * 32 + i] += __uint_as_float(orr[i]);
        }
        if (q0 + r < T_utt) {
            const float inv = 1.f / l_run;
            T* dst = out + static_cast<size_t>(row0 + q0 + r) * D + head * kD;
#pragma unroll
            for (int i = 0; i < kD / 8; ++i) {
                uint4 u;
                u.x = Half16<T>::pack(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
                u.y = Half16<T>::pack(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
                u.z = Half16<T>::pack(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
                u.w = Half16<T>::pack(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
                reinterpret_cast<uint4*>(dst)[i] = u;
            }
        }
    }

    tc05_fence_before();
    __syncthreads();
    tc05_fence_after();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

int launch_attention_tc05(int prec, const void* qkv, const RowSpace& rs, int heads, void* out,
                          cudaStream_t stream) {
    if (rs.n_attn128_work <= 0) return 0;
    B200_CHECK(prec == kPrecBf16 || prec == kPrecFp16, "attention: unsupported precision %d", prec);
    alignas(64) CUtensorMap tm;
    const int D = heads * kD;
    if (make_tmap_2d(&tm, qkv, prec == kPrecBf16 ? 0 : 1, rs.rows, 3 * D, 3 * D, 128)) return 1;
    dim3 grid(rs.n_attn128_work, heads);
    static bool configured = false;
    if (!configured) {
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__nv_bfloat16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
        B200_CUDA_OK(cudaFuncSetAttribute(attention_tc05_kernel<__half>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
        configured = true;
    }
    if (prec == kPrecBf16)
        attention_tc05_kernel<__nv_bfloat16><<<grid, kThreads, AttnSmem::kTotal, stream>>>(
            tm, static_cast<__nv_bfloat16*>(out), rs.attn128_work, heads);
    else
        attention_tc05_kernel<__half><<<grid, kThreads, AttnSmem::kTotal, stream>>>(
            tm, static_cast<__half*>(out), rs.attn128_work, heads);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
