// tcgen05 GEMM / implicit conv1d, CTA-pair version (cta_group::2) -- the main dense kernel.
//
// Same contract as gemm_tc05.cuh (see there for the reference call sites it replaces); what
// changes is the mapping onto the machine. ncu on the 1-CTA kernel (profiles/r01_gemm_ncu_full_
// summary.md) showed the tensor pipe idling 30-50 % with DRAM and L2 far from saturated: a single
// CTA must both read 96 B/clk of operands out of shared memory and let TMA write another 96 B/clk
// into it. Here two CTAs on the two SMs of a TPC form one cluster and issue ONE 256x256x16
// tcgen05.mma.cta_group::2 per K-step: each CTA stages its own 128 rows of A and only HALF of the
// B tile (128 of the 256 weight rows), the tensor cores read the other half from the peer's
// shared memory. Per SM that is 64 B/clk read + 64 B/clk written, and 1.5x fewer L2->SM bytes.
//
//   cluster (2 CTAs) tile: 256 (M) x 256 (N), K-step 64; 6-stage TMA ring of 32 KB per CTA
//   warp 0      TMA producer (both CTAs; loads signal the LEADER CTA's full barrier)
//   warp 1      MMA issuer   (leader CTA only; commits multicast to both CTAs' barriers)
//   warp 2      TMEM allocator (cta_group::2 alloc / dealloc, both CTAs)
//   warps 4-11  epilogue: 2 warps per TMEM lane quarter, each owning 128 of the 256 columns;
//               the fp32 residual and the next TMEM chunk are prefetched one chunk ahead.
#pragma once

#include "common.cuh"
#include "gemm_tc05.cuh"
#include "kernels.h"

namespace b200 {

constexpr int kGemm2Threads = 384;
constexpr int kGemm2BlockN = 256;
constexpr int kGemm2Stages = 6;

struct Gemm2Smem {
    static constexpr int kABytes = 128 * 128;  // this CTA's 128 rows of A, one 128-byte swizzle span
    static constexpr int kBBytes = 128 * 128;  // this CTA's half (128 rows) of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    // per epilogue warp: a 32-row x 32-word transpose buffer (row pitch 33 words: conflict-free
    // for both the row-per-lane writes and the 4-rows-per-instruction coalesced read-back)
    static constexpr int kStagePitch = 33;
    static constexpr int kStagingBytes = 8 * 32 * kStagePitch * 4;
    static constexpr int kStagingOffset = kGemm2Stages * kStageBytes;
    static constexpr int kBarOffset = kStagingOffset + kStagingBytes;
    static constexpr int kTotal = kBarOffset + (2 * kGemm2Stages + 4) * 8 + 16 + 1024;
};
static_assert(Gemm2Smem::kTotal <= 232448, "exceeds the 227 KB dynamic shared memory limit");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
                 : "memory");
}
// TMA load into THIS CTA's shared memory; completion bytes are signalled on `mbar_cluster_addr`
// (the leader CTA's full barrier), which .cta_group::2 permits to live in the peer CTA.
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* m,
                                                 uint32_t mbar_cluster_addr, int32_t c0,
                                                 int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0),
        "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the issued MMAs retire
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
                 : "memory");
}

template <typename InT, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_tc05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    using SM = Gemm2Smem;
    constexpr int BLOCK_N = kGemm2BlockN;
    constexpr int BLOCK_K = 64;
    constexpr int UMMA_K = 16;
    constexpr int kStages = kGemm2Stages;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages of 256 fp32 columns
    static_assert(sizeof(InT) == 2, "2-CTA kernel is instantiated for bf16 / fp16 operands");

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
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const int num_m = (p.M + 255) / 256;
    const int num_n = p.n_store / BLOCK_N;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.taps * p.k_blocks_per_tap;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);   // leader's: one arrive.expect_tx by the leader's producer
            mbar_init(&empty_bar[s], 1);  // one multicast commit per use
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 16);  // leader's: 8 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta<kTmemCols>(tmem_slot);
    tc05_fence_before();
    cluster_sync_all();  // peer barriers initialised, TMEM allocated in both CTAs
    tc05_fence_after();
    pdl_wait();  // barrier init, TMEM allocation and the cluster sync overlap the predecessor's tail
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer (both CTAs) ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                // n fastest: the clusters that share one A (activation) tile run concurrently, so
                // it is fetched from HBM once; the weight tiles are few and stay in L2
                const int n_blk = tile % num_n;
                const int m_blk = tile / num_n;
                const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
                const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    // both CTAs' bytes complete on the leader's barrier
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * SM::kStageBytes);
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    const int tap = kb / p.k_blocks_per_tap;
                    const int kc = kb - tap * p.k_blocks_per_tap;
                    const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
                    tma_load_2d_2cta(sa, &tmap_a, full_leader, kc * BLOCK_K, m0 + tap - p.tap_pad);
                    tma_load_2d_2cta(sa + SM::kABytes, &tmap_b, full_leader, kb * BLOCK_K, n0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader) -------------------------------
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc(UmmaFmt<InT>::value, 256, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
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
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                         (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[stage]);  // frees the stage in both CTAs
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(&tmem_full[acc]);  // accumulators of both CTAs complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------ epilogue (both CTAs) ------------------------------
        // Each warp drains 32 accumulator rows (its TMEM lane quarter) x 128 columns. TMEM hands
        // every lane one ROW; global memory wants one instruction to cover whole 128-byte lines.
        // So each chunk of 32 output words per row is transposed through a per-warp smem buffer:
        // bias / activation are applied row-per-lane, then 8 lanes cover 128 contiguous bytes of
        // one row for the residual load and the store (4 rows per instruction, 4 lines instead
        // of 32).
        const int q = warp & 3;          // TMEM lane quarter
        const int hh = (warp - 4) >> 2;  // which 128-column half of the tile
        constexpr bool kOut32 = sizeof(OutT) == 4;
        constexpr int kColsPerChunk = kOut32 ? 32 : 64;  // 32 words of output per row either way
        constexpr int kChunks = 128 / kColsPerChunk;
        constexpr int kPitch = SM::kStagePitch;
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + SM::kStagingOffset) + (warp - 4) * 32 * kPitch;
        const int sub_row = lane >> 3;       // coalesced phase: row within a group of 4
        const int sub_w = (lane & 7) * 4;    // coalesced phase: first of 4 words
        int acc = 0;
        uint32_t acc_phase = 0;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        const uint32_t tmem_empty_leader0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
        const uint32_t tmem_empty_leader1 = mapa_shared(smem_u32(&tmem_empty[1]), 0);
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int n_blk = tile % num_n;
            const int m_blk = tile / num_n;
            const int row_base = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
            const int row = row_base + lane;  // row-per-lane phase
            const bool row_zero = row < p.M && p.row_valid != nullptr && p.row_valid[row] == 0;
            const int ncol0 = n_blk * BLOCK_N + hh * 128;
            const bool has_res = kOut32 && p.residual != nullptr;
            float4 res[2][8];
            auto load_res = [&](int c, float4 (&dst)[8]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int grow = row_base + i * 4 + sub_row;
                    if (grow < p.M)
                        dst[i] = *reinterpret_cast<const float4*>(
                            p.residual + static_cast<size_t>(grow) * p.ld_res + ncol0 + c * 32 + sub_w);
                }
            };
            if (has_res) load_res(0, res[0]);  // in flight while the mainloop finishes
            // fused RMSNorm, consumer side: per-row 1/rms from the producer's eight partial sums
            float rscale = 1.f;
            if (p.ss_in != nullptr && row < p.M) {
                const float4* sp = reinterpret_cast<const float4*>(p.ss_in + static_cast<size_t>(row) * 8);
                const float4 s0 = sp[0], s1 = sp[1];
                const float tot = ((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w));
                rscale = rsqrtf(tot * p.ss_inv_dim + p.ss_eps);
            }
            float ssq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ssq[i] = 0.f;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc05_fence_after();
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                    static_cast<uint32_t>(acc * BLOCK_N + hh * 128);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int cur = c & 1, nxt = cur ^ 1;
                const int n0 = ncol0 + c * kColsPerChunk;
                uint32_t r[kColsPerChunk];
                tmem_ld_32x32(t_base + c * kColsPerChunk, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                if constexpr (!kOut32)
                    tmem_ld_32x32(t_base + c * kColsPerChunk + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                if (c + 1 < kChunks && has_res) load_res(c + 1, res[nxt]);
                tmem_ld_wait();
                float v[kColsPerChunk];
#pragma unroll
                for (int j = 0; j < kColsPerChunk; ++j) v[j] = __uint_as_float(r[j]) * rscale;
                if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
                    for (int j = 0; j < kColsPerChunk / 4; ++j) {
                        const float4 b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x;
                        v[4 * j + 1] += b.y;
                        v[4 * j + 2] += b.z;
                        v[4 * j + 3] += b.w;
                    }
                }
                if (p.act == kActSilu) {
#pragma unroll
                    for (int j = 0; j < kColsPerChunk; ++j) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));
                }
                if (row_zero) {
#pragma unroll
                    for (int j = 0; j < kColsPerChunk; ++j) v[j] = 0.f;
                }
                // row-per-lane -> smem (lane stride 33 words: conflict-free)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if constexpr (kOut32) stg[lane * kPitch + j] = __float_as_uint(v[j]);
                    else stg[lane * kPitch + j] = Half16<OutT>::pack(v[2 * j], v[2 * j + 1]);
                }
                __syncwarp();
                // smem -> global, 8 lanes per 128-byte row segment
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub_row;
                    const int grow = row_base + rr;
                    uint4 w;
                    w.x = stg[rr * kPitch + sub_w + 0];
                    w.y = stg[rr * kPitch + sub_w + 1];
                    w.z = stg[rr * kPitch + sub_w + 2];
                    w.w = stg[rr * kPitch + sub_w + 3];
                    if (grow < p.M) {
                        if constexpr (kOut32) {
                            float4 o = make_float4(__uint_as_float(w.x), __uint_as_float(w.y),
                                                   __uint_as_float(w.z), __uint_as_float(w.w));
                            if (has_res) {
                                o.x += res[cur][i].x;
                                o.y += res[cur][i].y;
                                o.z += res[cur][i].z;
                                o.w += res[cur][i].w;
                                // halo rows stay zero even when a residual is added
                                if (p.row_valid != nullptr && p.row_valid[grow] == 0) o = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            *reinterpret_cast<float4*>(out + static_cast<size_t>(grow) * p.ldc + n0 + sub_w) = o;
                            // fused RMSNorm, producer side: 16-bit copy + row sum of squares
                            if (p.out16 != nullptr) {
                                uint2 h2;
                                h2.x = Half16<InT>::pack(o.x, o.y);
                                h2.y = Half16<InT>::pack(o.z, o.w);
                                *reinterpret_cast<uint2*>(reinterpret_cast<InT*>(p.out16) +
                                                          static_cast<size_t>(grow) * p.ld16 + n0 + sub_w) = h2;
                            }
                            ssq[i] = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, ssq[i]))));
                        } else {
                            *reinterpret_cast<uint4*>(out + static_cast<size_t>(grow) * p.ldc + n0 + 2 * sub_w) = w;
                        }
                    }
                }
                __syncwarp();  // staging buffer is reused by the next chunk
            }
            tc05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc == 0 ? tmem_empty_leader0 : tmem_empty_leader1);
            if constexpr (kOut32) {
                if (p.ss_out != nullptr) {
                    // 8 lanes hold the 128 columns of one row: fixed-order butterfly, one slot per warp
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t = ssq[i];
                        t += __shfl_xor_sync(0xffffffffu, t, 1);
                        t += __shfl_xor_sync(0xffffffffu, t, 2);
                        t += __shfl_xor_sync(0xffffffffu, t, 4);
                        const int grow = row_base + i * 4 + sub_row;
                        if ((lane & 7) == 0 && grow < p.M)
                            p.ss_out[static_cast<size_t>(grow) * 8 + (ncol0 >> 7)] = t;
                    }
                }
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    // the leader's MMAs read the peer's shared memory: neither CTA may exit early
    tc05_fence_before();
    cluster_sync_all();
    tc05_fence_after();
    if (warp == 2) tmem_dealloc_2cta<kTmemCols>(tmem_base);
}

template <typename InT, typename OutT>
int launch_gemm_tc05_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          cudaStream_t stream) {
    auto kern = gemm_tc05_2cta_kernel<InT, OutT>;
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Gemm2Smem::kTotal));
    }
    const int num_m = (p.M + 255) / 256;
    const int num_n = p.n_store / kGemm2BlockN;
    int clusters = num_m * num_n;
    if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
    if (clusters < 1) return 0;
    B200_CUDA_OK(launch_kernel(kern, dim3(2 * clusters), dim3(kGemm2Threads), Gemm2Smem::kTotal, stream, ta, tb, p));
    return 0;
}

}  // namespace b200
