// K8: unmasked, non-causal multi-head attention over packed varlen utterances.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0,
// is_causal=False) at tts/core/codec/decoder_modules.py:283-285, with the einops
// rearranges of :276-278 and :287 folded into the addressing: q/k/v are read straight
// from the c_attn output [rows, 3*H*64] ("(r h d)" column order) and the result is written
// as [rows, H*64] ("b t (h d)"). The head-indexed RoPE of :280-281 is folded into the
// c_attn weights at load time (codec.cu), so q and k arrive already rotated.
//
// v1 kernel: flash-style online softmax, one CTA per (utterance, 64-query tile, head),
// 4 warps x 16 query rows, K/V tiles of 64 keys double-buffered through shared memory with
// cp.async, QK^T and PV on mma.sync.m16n8k16 (fp32 accumulate), softmax in fp32 registers.
// Scores are never materialised, so T = 3000 (long-form) needs no extra memory.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int kHeadDim = 64;
constexpr int kBlockQ = kAttnBlockQ;  // 64
constexpr int kBlockKV = 64;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr, bool valid) {
    const int src_bytes = valid ? 16 : 0;  // 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                            uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1,
                                                  uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                          uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4],
                                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4],
                                                  uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
        "{%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile [64 rows][64 halfs]; 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return static_cast<uint32_t>((row * 8 + (chunk ^ (row & 7))) * 16);
}

// copy up to 64 rows x 128 B from global (row stride ld elements) into a swizzled tile
template <typename T>
__device__ __forceinline__ void load_tile_async(uint32_t smem_base, const T* gbase, int ld,
                                                int rows_valid) {
#pragma unroll
    for (int i = 0; i < (64 * 8) / kAttnThreads; ++i) {
        const int c = threadIdx.x + i * kAttnThreads;
        const int row = c >> 3, chunk = c & 7;
        const bool ok = row < rows_valid;
        const T* src = gbase + static_cast<size_t>(ok ? row : 0) * ld + chunk * 8;
        cp_async_16(smem_base + tile_off(row, chunk), src, ok);
    }
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, const int4* __restrict__ work,
                 int heads) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(128) uint8_t smem[(1 + 2 + 2) * 64 * 64 * 2];  // Q, K[2], V[2] = 40 KB
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + 8192;
    const uint32_t sV = sK + 2 * 8192;

    const int4 wk = work[blockIdx.x];
    const int row0 = wk.x, T_utt = wk.y, q0 = wk.z;
    const int head = blockIdx.y;
    const int D = heads * kHeadDim;  // 1024
    const int ld = 3 * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const T* q_base = qkv + static_cast<size_t>(row0 + q0) * ld + head * kHeadDim;
    const T* k_base = qkv + static_cast<size_t>(row0) * ld + D + head * kHeadDim;
    const T* v_base = k_base + D;
    const int q_valid = min(kBlockQ, T_utt - q0);
    const int n_kv = (T_utt + kBlockKV - 1) / kBlockKV;

    load_tile_async<T>(sQ, q_base, ld, q_valid);
    load_tile_async<T>(sK, k_base, ld, min(kBlockKV, T_utt));
    load_tile_async<T>(sV, v_base, ld, min(kBlockKV, T_utt));
    cp_async_commit();

    // per-thread state: rows r0 = lane/4 and r0 + 8 of this warp's 16-row slab
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    uint32_t qf[4][4];
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

    for (int j = 0; j < n_kv; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_kv) {
            const int kv_next = (j + 1) * kBlockKV;
            const int nvalid = min(kBlockKV, T_utt - kv_next);
            load_tile_async<T>(sK + (buf ^ 1) * 8192, k_base + static_cast<size_t>(kv_next) * ld,
                               ld, nvalid);
            load_tile_async<T>(sV + (buf ^ 1) * 8192, v_base + static_cast<size_t>(kv_next) * ld,
                               ld, nvalid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (j == 0) {
            // Q fragments: A operand, 16 rows x (4 k-steps of 16)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = ks * 2 + (lane >> 4);
                ldmatrix_x4(sQ + tile_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }

        // ---- S = Q K^T : 16 x 64 per warp ----
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
        const uint32_t kb = sK + buf * 8192;
#pragma unroll
        for (int n = 0; n < 8; ++n) {        // 8 keys per n-tile
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {  // two k-steps (32 of d) per ldmatrix.x4
                uint32_t b0, b1, b2, b3;
                const int row = n * 8 + (lane & 7);
                const int chunk = kp * 4 + (lane >> 3);
                ldmatrix_x4(kb + tile_off(row, chunk), b0, b1, b2, b3);
                mma_16816<T>(s[n], qf[kp * 2 + 0], b0, b1);
                mma_16816<T>(s[n], qf[kp * 2 + 1], b2, b3);
            }
        }

        // ---- mask keys beyond the utterance, online softmax ----
        const int kv0 = j * kBlockKV;
        if (kv0 + kBlockKV > T_utt) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int key = kv0 + n * 8 + (lane & 3) * 2;
                if (key >= T_utt) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
                if (key + 1 >= T_utt) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            mx[0] = fmaxf(mx[0], fmaxf(s[n][0], s[n][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[n][2], s[n][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], m_new[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            m_new[r] = fmaxf(m_run[r], mx[r] * scale_log2);
            corr[r] = exp2f(m_run[r] - m_new[r]);  // first tile: exp2(-inf) = 0
            m_run[r] = m_new[r];
            l_run[r] *= corr[r];
        }
        uint32_t pf[4][4];  // P as A operand for 4 k-steps of 16 keys
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = exp2f(fmaf(s[n][0], scale_log2, -m_new[0]));
            const float p1 = exp2f(fmaf(s[n][1], scale_log2, -m_new[0]));
            const float p2 = exp2f(fmaf(s[n][2], scale_log2, -m_new[1]));
            const float p3 = exp2f(fmaf(s[n][3], scale_log2, -m_new[1]));
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[n >> 1][(n & 1) * 2 + 0] = Half16<T>::pack(p0, p1);
            pf[n >> 1][(n & 1) * 2 + 1] = Half16<T>::pack(p2, p3);
        }
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            o[d][0] *= corr[0];
            o[d][1] *= corr[0];
            o[d][2] *= corr[1];
            o[d][3] *= corr[1];
        }

        // ---- O += P V : 16 x 64 per warp ----
        const uint32_t vb = sV + buf * 8192;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {      // 16 keys per k-step
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {  // two d n-tiles (16 of d) per ldmatrix.x4.trans
                uint32_t b0, b1, b2, b3;
                const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = dp * 2 + (lane >> 4);
                ldmatrix_x4_trans(vb + tile_off(row, chunk), b0, b1, b2, b3);
                mma_16816<T>(o[dp * 2 + 0], pf[ks], b0, b1);
                mma_16816<T>(o[dp * 2 + 1], pf[ks], b2, b3);
            }
        }
        __syncthreads();  // all warps done with `buf` before it is refilled
    }

    // ---- finalize: O / l, write [row, head*64 + d] ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    const int qr0 = warp * 16 + (lane >> 2);
    T* o_base = out + static_cast<size_t>(row0 + q0) * D + head * kHeadDim + (lane & 3) * 2;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        if (qr0 < q_valid)
            *reinterpret_cast<uint32_t*>(o_base + static_cast<size_t>(qr0) * D + d * 8) =
                Half16<T>::pack(o[d][0] * inv0, o[d][1] * inv0);
        if (qr0 + 8 < q_valid)
            *reinterpret_cast<uint32_t*>(o_base + static_cast<size_t>(qr0 + 8) * D + d * 8) =
                Half16<T>::pack(o[d][2] * inv1, o[d][3] * inv1);
    }
}

}  // namespace

int launch_attention(int prec, const void* qkv, const RowSpace& rs, int heads, void* out,
                     cudaStream_t stream) {
    if (rs.n_attn_work <= 0) return 0;
    dim3 grid(rs.n_attn_work, heads);
    if (prec == kPrecBf16)
        B200_CUDA_OK(launch_kernel(attention_kernel<__nv_bfloat16>, grid, dim3(kAttnThreads), 0, stream,
                                   static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                   rs.attn_work, heads));
    else if (prec == kPrecFp16)
        B200_CUDA_OK(launch_kernel(attention_kernel<__half>, grid, dim3(kAttnThreads), 0, stream,
                                   static_cast<const __half*>(qkv), static_cast<__half*>(out), rs.attn_work, heads));
    else {
        set_error("attention: unsupported precision %d", prec);
        return 1;
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
