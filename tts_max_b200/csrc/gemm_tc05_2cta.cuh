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
//   cluster (2 CTAs) tile: 256 (M) x 256 (N), K-step 64; 5-stage TMA ring of 32 KB per CTA
//   warp 0      TMA producer (both CTAs; loads signal the LEADER CTA's full barrier)
//   warp 1      MMA issuer   (leader CTA only; commits multicast to both CTAs' barriers)
//   warp 2      TMEM allocator (cta_group::2 alloc / dealloc, both CTAs)
//   warps 4-11  epilogue: 2 warps per TMEM lane quarter, each owning 128 of the 256 columns.
//
// Epilogue. An in-kernel timeline (tools/gemm_trace.cu) showed the mainloop at ~6.3 us per tile
// (about 1600 TFLOP/s) but an LDG/STG epilogue at 8-13 us per fp32 tile: two warps per scheduler,
// ~650 dependent instructions per 32-column chunk (address arithmetic, bound checks, a transpose
// through shared memory), plus ~1 us per tile in the cluster-scope release of the tmem_empty
// arrive. Now: per 32-column chunk a lane (= one accumulator row) adds the fp32 residual it
// fetched with four 256-bit loads one chunk earlier (whole sectors, so row-per-lane access is
// not wasteful), writes the fp32 result and the 16-bit copy into TMA-swizzled staging boxes,
// and one elected lane issues the two TMA stores. The warp computes no store address, rows
// beyond M are clipped by the tensor maps, and the next TMEM chunk is always in flight.
#pragma once

#include "common.cuh"
#include "gemm_tc05.cuh"
#include "kernels.h"

namespace b200 {

constexpr int kGemm2Threads = 384;
constexpr int kGemm2BlockN = 256;
constexpr int kGemm2ChunkCols = 32;  // accumulator columns per epilogue step

// Shared memory of one CTA for tiles of 256 x BLOCK_N: a stage = this CTA's 128 rows of A + its half of the B
// tile. Narrower tiles have smaller stages, so more of them fit: 8 stages at BLOCK_N = 64 instead of 5 -- the
// K loop of a narrow tile is bounded by TMA latency x loads in flight (encoder convs with K = 448 ... 1344,
// B = 1 decode), not by the tensor pipe.
template <int BLOCK_N>
struct Gemm2SmemT {
    static constexpr int kStages = BLOCK_N == 64 ? 8 : BLOCK_N == 128 ? 6 : 5;
    static constexpr int kABytes = 128 * 128;            // this CTA's 128 rows of A, one 128-byte swizzle span
    static constexpr int kBBytes = (BLOCK_N / 2) * 128;  // this CTA's half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static_assert(kStageBytes % 1024 == 0, "SWIZZLE_128B tiles start on 1024-byte boundaries");
    // per epilogue warp: one 32-row x 32-column fp32 staging box (128-byte rows, SWIZZLE_128B)
    // and two 16-bit boxes (64-byte rows, SWIZZLE_64B). A TMA store has read its box well within
    // one chunk period (measured), so the fp32 box is not double-buffered.
    static constexpr int kBox32Bytes = 32 * kGemm2ChunkCols * 4;
    static constexpr int kBox16Bytes = 32 * kGemm2ChunkCols * 2;
    static constexpr int kEpiWarpBytes = kBox32Bytes + 2 * kBox16Bytes;
    static constexpr int kEpiOffset = kStages * kStageBytes;
    static constexpr int kBarOffset = kEpiOffset + 8 * kEpiWarpBytes;
    // full[stages], empty[stages], tmem_full[2], tmem_empty[2]
    static constexpr int kNumBars = 2 * kStages + 4;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
    static_assert(kTotal <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};
using Gemm2Smem = Gemm2SmemT<256>;
constexpr int kGemm2Stages = Gemm2Smem::kStages;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only cluster barrier. The release form costs a MEMBAR.GPU (~0.3-0.5 us, tools/gemm_trace.cu)
// and this kernel never needs it: mbarrier initialisation is published by fence.mbarrier_init, the TMEM
// slot is CTA-local (a __syncthreads orders it), and at exit only "nobody leaves early" matters.
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
// The only data this arrive orders is TMEM traffic, which tcgen05.fence::before_thread_sync
// already covers; a .release at cluster scope would add a MEMBAR.GPU that waits ~1 us for every
// outstanding global access of the warp (tools/gemm_trace.cu).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
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

// tmap_o32: fp32 [M, n_store] output, box 32 x 32, SWIZZLE_128B (used when p.has32);
// tmap_o16: 16-bit [M, n_store] output or operand copy, box 32 x 32, SWIZZLE_64B (p.has16).
// kGnStats: the instantiation that also reduces GroupNorm statistics (p.gn_stats), kept apart so the
// extra registers do not touch the common kernel.
// BLOCK_N: 256, or 64 for small M: with one to a few m-blocks only N / 256 of the 74 CTA pairs would
// have a tile and each would walk the whole K loop alone; 256 x 64 tiles put four times as many pairs
// to work (the K loop of a narrow tile is bounded by the A-tile load, ~0.15 us per K-block instead of
// 0.26). 128 / 192: the widths of the encoder's 96(128)-, 192- and 384-channel convs -- with 64-wide tiles
// every n-tile re-reads the A operand from L2 once per tap (a 7-tap conv at N = 384: 42 times), which
// made those launches L2-bandwidth-bound. Every output element sees the same K order, so the tile width
// never changes results. The shared-memory stage keeps its 256-wide size; TMEM holds 2 x BLOCK_N columns.
// ---------------------------------------------------------------------------------------------------
// GEMM chains (kChain). Inside a transformer block c_proj -> fc1 -> fc2 -> (next block's) c_attn are
// row-block local: RMSNorm is per row and already travels as sum-of-squares partials, so a 256-row block
// of GEMM g+1 needs nothing but the same 256 rows of GEMM g. As separate launches every boundary drains
// the grid (griddepcontrol.wait), pays ~5 us of prologue / exposed epilogue / teardown and rounds each
// GEMM up to whole waves of 74 tiles (86 % for the N = 1024 GEMMs). Here ONE persistent launch walks the
// tiles of up to four GEMMs in a single global order (GEMM by GEMM, m-block major, n fastest; cluster c
// takes tiles c, c + 74, ...). A tile of GEMM g > 0 waits until every tile of GEMM g-1 in its m-block has
// been stored: the epilogue warps count themselves into counters[g][m_blk] with a gpu-scope release after
// their TMA stores completed, the TMA producer (and the epilogue warps, for the residual / row-scale
// reads) acquire it before touching that m-block. Tiles only ever wait for tiles EARLIER in the order and
// all clusters are co-resident (grid <= 74 clusters of one CTA per SM), so the smallest unfinished tile can
// always run: no deadlock. Weights are constants: their loads never wait.
// ---------------------------------------------------------------------------------------------------
constexpr int kChainMax = 4;

struct ChainMaps {
    CUtensorMap a[kChainMax], b[kChainMax], o32[kChainMax], o16[kChainMax];
};
struct ChainParams {
    int n_gemm;                  // 1 = plain GEMM
    int num_m;                   // m-blocks of 256 rows (all GEMMs of a chain share M)
    int tile_end[kChainMax];     // cumulative tile counts in the global order
    int num_n[kChainMax];        // n-tiles per m-block
    uint32_t full[kChainMax];    // counter value of a finished m-block: 16 epilogue warps x num_n
    uint32_t* counters;          // [n_gemm][num_m], zero on entry (chains only)
    // chains: the tiles each cluster walks, in increasing global order, -1 terminated:
    // sched[cluster * sched_stride + k]. Built on the host by in-order list scheduling on the tiles' K
    // lengths, because a plain round robin gives some clusters two of fc2's long (K = 4096) tiles and
    // others one -- 4 of ~21 tile units of imbalance (measured: the chain was then no faster in the step).
    const int32_t* sched;
    int sched_stride;
    int early_weights;  // 1: weight loads of the first stages are issued before griddepcontrol.wait (A/B)
    int dbg;  // timing experiments only (tools): 1 no store-completion wait, 2 no epilogue acquire, 4 no producer fence, 8 no publish
    GemmParams g[kChainMax];
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// orders async-proxy (TMA) accesses to global memory against generic-proxy accesses of this thread
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Bounded wait for an m-block of the previous GEMM of the chain (a protocol bug must trap, not hang).
__device__ __forceinline__ void chain_wait(const uint32_t* ctr, uint32_t full) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(ctr) < full) {
        __nanosleep(64);
        if (++spins > (1u << 24)) {
            printf("b200codec: GEMM chain dependency timed out (block %d thread %d: %u of %u)\n", blockIdx.x,
                   threadIdx.x, ld_acquire_gpu(ctr), full);
            __trap();
        }
    }
}

template <typename InT, bool kGnStats, int BLOCK_N, bool kChain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_tc05_2cta_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams cp) {
    const GemmParams& p = cp.g[0];  // plain GEMM: the only one; chains: per-tile parameters are cp.g[g]
    using SM = Gemm2SmemT<BLOCK_N>;
    static_assert(BLOCK_N == 256 || BLOCK_N == 192 || BLOCK_N == 128 || BLOCK_N == 64, "tile width");
    constexpr int kHalfN = BLOCK_N / 2;  // B rows staged by each CTA = accumulator columns per epilogue warp
    constexpr uint32_t kStageTx = 2 * (SM::kABytes + kHalfN * 128);  // bytes both CTAs land per stage
    constexpr int BLOCK_K = 64;
    constexpr int UMMA_K = 16;
    constexpr int kStages = SM::kStages;
    // two accumulator stages of BLOCK_N fp32 columns; allocations are powers of two (2 x 192 -> 512)
    constexpr uint32_t kTmemCols = 2 * BLOCK_N <= 128 ? 128 : 2 * BLOCK_N <= 256 ? 256 : 512;
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
    if (threadIdx.x == 0) B200_TRACE(0);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    const int num_tiles = cp.tile_end[kChain ? cp.n_gemm - 1 : 0];
    const int cluster_id = blockIdx.x >> 1;
    // the tiles of this cluster: a host-built list (chains) or every num_clusters-th tile. next_tile is
    // called at the TOP of a tile's body, so the list entry (an L2 hit, ~0.6 us) is in flight during the tile
    const int num_clusters = gridDim.x >> 1;
    auto first_tile = [&]() -> int {
        if constexpr (kChain) return __ldg(cp.sched + static_cast<size_t>(cluster_id) * cp.sched_stride);
        return cluster_id < num_tiles ? cluster_id : -1;
    };
    auto next_tile = [&](int it, int tile) -> int {
        if constexpr (kChain) return __ldg(cp.sched + static_cast<size_t>(cluster_id) * cp.sched_stride + it);
        const int t = tile + num_clusters;
        return t < num_tiles ? t : -1;
    };
    // global tile index -> (GEMM of the chain, m-block, n-block); n fastest: the clusters that share one
    // A (activation) tile run concurrently, so it is fetched from HBM once; the weight tiles are few and
    // stay in L2
    auto decode_tile = [&](int tile, int& g, int& m_blk, int& n_blk) {
        g = 0;
        int local = tile;
        if constexpr (kChain) {
            while (g + 1 < cp.n_gemm && tile >= cp.tile_end[g]) ++g;
            if (g > 0) local = tile - cp.tile_end[g - 1];
        }
        const int nn = cp.num_n[g];
        m_blk = local / nn;
        n_blk = local - m_blk * nn;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b[0]);
    }
    if (warp == 3 && lane == 0) {
        if (p.has32) tma_prefetch_desc(&maps.o32[0]);
        if (p.has16) tma_prefetch_desc(&maps.o16[0]);
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
        B200_TRACE(21);
    }
    if (warp == 2) {
        if (lane == 0) B200_TRACE(22);
        tmem_alloc_2cta<kTmemCols>(tmem_slot);
        if (lane == 0) B200_TRACE(23);
    }
    tc05_fence_before();
    __syncthreads();         // this CTA's barriers and TMEM slot
    cluster_sync_relaxed();  // peer barriers initialised, TMEM allocated in both CTAs
    tc05_fence_after();
    if (threadIdx.x == 0) B200_TRACE(1);
    // Barrier init, TMEM allocation and the cluster sync overlap the predecessor's tail (programmatic
    // dependent launch). griddepcontrol.wait is per thread and only the threads that read what the predecessor
    // wrote need it: the producer (A operand) and the epilogue warps (residual, row scale). The producer
    // first puts the WEIGHT halves of its first stages in flight -- weights are written at load time, never by
    // a kernel of the decode -- so their DRAM latency (~2.4 us from prologue to first operands in the
    // in-kernel timeline, cold L2) overlaps the predecessor's tail as well.
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer (both CTAs) ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // weight halves of the first tile's first stages, before waiting for the predecessor (the ring is
            // empty: no empty-barrier wait; the leader arms each stage's barrier for all of its bytes)
            int early_b = 0;
            {
                const int tile0 = first_tile();
                if (tile0 >= 0 && cp.early_weights) {
                    int g, m_blk, n_blk;
                    decode_tile(tile0, g, m_blk, n_blk);
                    const int num_kb = cp.g[g].taps * cp.g[g].k_blocks_per_tap;
                    const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * kHalfN;
                    early_b = num_kb < kStages ? num_kb : kStages;
                    for (int kb = 0; kb < early_b; ++kb) {
                        if (leader) mbar_arrive_expect_tx(&full_bar[kb], kStageTx);
                        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[kb]), 0);
                        tma_load_2d_2cta(smem_u32(smem + kb * SM::kStageBytes) + SM::kABytes, &maps.b[g], full_leader,
                                         kb * BLOCK_K, n0);
                    }
                }
            }
            pdl_wait();
            for (int it = 0, tile = first_tile(), tile_next; tile >= 0; tile = tile_next) {
                tile_next = next_tile(++it, tile);
                int g, m_blk, n_blk;
                decode_tile(tile, g, m_blk, n_blk);
                const GemmParams& pg = cp.g[g];
                const CUtensorMap* tmap_a = &maps.a[g];
                const CUtensorMap* tmap_b = &maps.b[g];
                const int num_kb = pg.taps * pg.k_blocks_per_tap;
                const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
                const int n0 = n_blk * BLOCK_N + static_cast<int>(rank) * kHalfN;
                bool dep_pending = kChain && g > 0;  // this m-block of the previous GEMM must be stored first
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    const int tap = kb / pg.k_blocks_per_tap;
                    const int kc = kb - tap * pg.k_blocks_per_tap;
                    const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
                    if (early_b > 0) {
                        --early_b;  // this stage is armed and its weight half is in flight
                    } else {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        // both CTAs' bytes complete on the leader's barrier
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], kStageTx);
                        tma_load_2d_2cta(sa + SM::kABytes, tmap_b, full_leader, kb * BLOCK_K, n0);  // weights never wait
                    }
                    if constexpr (kChain) {
                        if (dep_pending) {
                            if (!(cp.dbg & 8)) chain_wait(cp.counters + (g - 1) * cp.num_m + m_blk, cp.full[g - 1]);
                            if (!(cp.dbg & 4)) fence_proxy_async_all();  // the acquired rows are read through the async proxy (TMA)
                            dep_pending = false;
                        }
                    }
                    tma_load_2d_2cta(sa, tmap_a, full_leader, kc * BLOCK_K, m0 + tap * pg.tap_dil - pg.tap_pad);
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
            int trace_tile = 0;
            (void)trace_tile;
            for (int it = 0, tile = first_tile(), tile_next; tile >= 0; tile = tile_next) {
                tile_next = next_tile(++it, tile);
                int g, m_blk_unused, n_blk_unused;
                decode_tile(tile, g, m_blk_unused, n_blk_unused);
                const int num_kb = cp.g[g].taps * cp.g[g].k_blocks_per_tap;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc05_fence_after();
                    if (kb == 0 && trace_tile < 4) B200_TRACE(2 + 2 * trace_tile);  // first operands landed
                    if (kChain && kb == 0 && trace_tile < 24) B200_TRACE(32 + 4 * trace_tile);  // tools/chain_trace.cu
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
                if (trace_tile < 4) B200_TRACE(3 + 2 * trace_tile);  // last MMA of the tile issued
                if (kChain && trace_tile < 24) B200_TRACE(32 + 4 * trace_tile + 1);
                ++trace_tile;
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------ epilogue (both CTAs) ------------------------------
        // Each warp drains 32 accumulator rows (its TMEM lane quarter) x 128 columns in four
        // 32-column steps; a lane owns one row. Staging boxes use the TMA swizzles, so the
        // row-per-lane 128-bit writes are bank-conflict-free: the 16-byte group j of row r sits
        // at slot j ^ (r & 7) of its 128-byte fp32 row and at slot j ^ ((r >> 1) & 3) of its
        // 64-byte 16-bit row.
        pdl_wait();  // residual, row scale, GroupNorm maps: written by predecessors
        constexpr int CH = kGemm2ChunkCols;
        constexpr int kChunks = kHalfN / CH;
        const int ew = warp - 4;
        const int q = warp & 3;   // TMEM lane quarter
        const int hh = ew >> 2;   // which 128-column half of the tile
        const uint32_t box32 = smem_u32(smem + SM::kEpiOffset + ew * SM::kEpiWarpBytes);
        const uint32_t ring16 = box32 + SM::kBox32Bytes;
        const uint32_t row32 = box32 + lane * 128, swz32 = lane & 7;
        const uint32_t row16 = lane * 64, swz16 = (lane >> 1) & 3;
        int slot16 = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t tmem_empty_leader0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
        const uint32_t tmem_empty_leader1 = mapa_shared(smem_u32(&tmem_empty[1]), 0);
        int trace_tile = 0;
        (void)trace_tile;
        for (int it = 0, tile = first_tile(), tile_next; tile >= 0; tile = tile_next) {
            tile_next = next_tile(++it, tile);
            int g, m_blk, n_blk;
            decode_tile(tile, g, m_blk, n_blk);
            const GemmParams& p = cp.g[g];  // shadows the kernel-level alias: this tile's GEMM
            const CUtensorMap* tmap_o32 = &maps.o32[g];
            const CUtensorMap* tmap_o16 = &maps.o16[g];
            const bool has32 = p.has32 != 0;
            const bool has16 = p.has16 != 0;
            const bool has_res = has32 && p.residual != nullptr;
            if constexpr (kChain) {
                // the residual and the row scale of this m-block come from earlier GEMMs of the chain: acquire
                // before the prefetches below (the counter is long complete by now; the producer waited for it)
                if (g > 0 && !(cp.dbg & (2 | 8))) {
                    if (lane == 0) chain_wait(cp.counters + (g - 1) * cp.num_m + m_blk, cp.full[g - 1]);
                    __syncwarp();
                }
            }
            const int row_base = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32;
            const int row = row_base + lane;
            const bool row_ok = row < p.M;
            const bool row_zero = row_ok && p.row_valid != nullptr && p.row_valid[row] == 0;
            const bool any_zero = __any_sync(0xffffffffu, row_zero);  // halo rows are rare
            int gn_utt = -1;         // this lane's utterance (GroupNorm statistics), -1: not a frame
            bool gn_uniform = true;  // every frame row of the warp belongs to one utterance
            int gn_u0 = -1;
            if constexpr (kGnStats) {
                if (row_ok) gn_utt = p.gn_row_utt[row];
                gn_u0 = gn_utt;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gn_u0 = max(gn_u0, __shfl_xor_sync(0xffffffffu, gn_u0, o));
                gn_uniform = __all_sync(0xffffffffu, gn_utt < 0 || gn_utt == gn_u0);
            }
            const int ncol0 = n_blk * BLOCK_N + hh * kHalfN;
            // this lane's residual row: 128 contiguous bytes per chunk, fetched as whole sectors
            const float* res_row = has_res && row_ok
                                       ? p.residual + static_cast<size_t>(row) * p.ld_res + ncol0
                                       : nullptr;
            float res[2][CH];  // two chunks of lookahead
            auto load_res = [&](int c, float (&dst)[CH]) {
                if (res_row != nullptr) {
#pragma unroll
                    for (int j = 0; j < CH / 8; ++j) ldg_256(res_row + c * CH + j * 8, &dst[j * 8]);
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) dst[j] = 0.f;
                }
            };
            if (has_res) {  // in flight while the mainloop finishes
                load_res(0, res[0]);
                load_res(1, res[1]);
            }
            // fused RMSNorm, consumer side: per-row 1/rms from the producer's eight partial sums
            float rscale = 1.f;
            if (p.ss_in != nullptr && row_ok) {
                // 32 partial sums (one per 32-column chunk of the 1024-wide x), added in a fixed order that
                // does not depend on the tile width of the GEMM that produced them
                const float4* sp = reinterpret_cast<const float4*>(p.ss_in + static_cast<size_t>(row) * kGemmSsSlots);
                float tot = 0.f;
#pragma unroll
                for (int i = 0; i < kGemmSsSlots / 4; ++i) {
                    // chains: an SM may have read this row's partials for an earlier GEMM of the same launch
                    const float4 s4 = kChain ? __ldcg(sp + i) : sp[i];
                    tot += (s4.x + s4.y) + (s4.z + s4.w);
                }
                rscale = rsqrtf(tot * p.ss_inv_dim + p.ss_eps) * p.ss_in_scale;
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc05_fence_after();
            if (warp == 4 && lane == 0 && trace_tile < 4) B200_TRACE(10 + 2 * trace_tile);  // accumulator ready
            if (kChain && warp == 4 && lane == 0 && trace_tile < 24) B200_TRACE(32 + 4 * trace_tile + 2);
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                    static_cast<uint32_t>(acc * BLOCK_N + hh * kHalfN);
            uint32_t r_next[CH];
            tmem_ld_32x32(t_base, r_next);  // chunk c+1 is in flight while chunk c is processed
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                const int n0 = ncol0 + c * CH;
#ifdef B200_GEMM_TRACE
                const bool tr = !kChain && warp == 4 && lane == 0 && trace_tile < 2;
                const int tslot = 32 + trace_tile * 48 + c * 6;
                if (tr) B200_TRACE(tslot + 0);
#endif
                const uint32_t buf16 = ring16 + slot16 * SM::kBox16Bytes;
                tmem_ld_wait();
                float v[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r_next[j]);
                if (c + 1 < kChunks) {
                    tmem_ld_32x32(t_base + (c + 1) * CH, r_next);
                } else {
                    // the accumulator stage is drained: hand it back before finishing the chunk
                    tc05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc == 0 ? tmem_empty_leader0 : tmem_empty_leader1);
                }
#ifdef B200_GEMM_TRACE
                if (tr) B200_TRACE(tslot + 2);  // accumulator chunk in registers
#endif
                if (p.ss_in != nullptr) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] *= rscale;
                }
                if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
                    for (int j = 0; j < CH / 4; ++j) {
                        const float4 b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x;
                        v[4 * j + 1] += b.y;
                        v[4 * j + 2] += b.z;
                        v[4 * j + 3] += b.w;
                    }
                }
                if (p.act == kActSilu) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));
                } else if (p.act == kActRelu) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] += res[c & 1][j];
                    if (c + 2 < kChunks) load_res(c + 2, res[c & 1]);
                }
                if (any_zero) {  // halo rows stay zero, with or without a residual
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = row_zero ? 0.f : v[j];
                }
                if constexpr (kGnStats) {
                    // GroupNorm statistics of this chunk = one group of 32 channels: fp32 per row in
                    // a fixed order (position-independent), fp64 across rows (order-insensitive at
                    // fp32 resolution), so the statistics do not depend on the batch an utterance is in
                    float gs = 0.f, gq = 0.f;
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        gs += v[j];
                        gq = fmaf(v[j], v[j], gq);
                    }
                    if (gn_utt < 0) gs = gq = 0.f;
                    const int group = n0 >> 5;
                    if (gn_uniform) {
                        if (gn_u0 >= 0) {
                            double ds = static_cast<double>(gs), dq = static_cast<double>(gq);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                ds += __shfl_xor_sync(0xffffffffu, ds, o);
                                dq += __shfl_xor_sync(0xffffffffu, dq, o);
                            }
                            if (lane == 0) {
                                atomicAdd(p.gn_stats + (static_cast<size_t>(gn_u0) * 32 + group) * 2 + 0, ds);
                                atomicAdd(p.gn_stats + (static_cast<size_t>(gn_u0) * 32 + group) * 2 + 1, dq);
                            }
                        }
                    } else if (gn_utt >= 0) {  // a warp that straddles two utterances (rare)
                        atomicAdd(p.gn_stats + (static_cast<size_t>(gn_utt) * 32 + group) * 2 + 0, static_cast<double>(gs));
                        atomicAdd(p.gn_stats + (static_cast<size_t>(gn_utt) * 32 + group) * 2 + 1, static_cast<double>(gq));
                    }
                }
                // the boxes this chunk writes were last read by the stores of chunk g-1 (fp32)
                // and g-2 (16-bit); those are normally long done
                if (lane == 0) bulk_wait_group_read<0>();
                __syncwarp();
#ifdef B200_GEMM_TRACE
                if (tr) B200_TRACE(tslot + 1);  // staging boxes free
#endif
                if (has32) {
#pragma unroll
                    for (int j = 0; j < CH / 4; ++j)
                        sts_128(row32 + ((static_cast<uint32_t>(j) ^ swz32) << 4),
                                __float_as_uint(v[4 * j + 0]), __float_as_uint(v[4 * j + 1]),
                                __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    // fused RMSNorm, producer side: row sum of squares (fixed order)
                    if (p.ss_out != nullptr) {
                        float cs = 0.f;
#pragma unroll
                        for (int j = 0; j < CH; ++j) cs = fmaf(v[j], v[j], cs);
                        if (row_ok) p.ss_out[static_cast<size_t>(row) * kGemmSsSlots + (n0 >> 5)] = cs;  // one slot per chunk
                    }
                }
                if (has16) {
                    // the 16-bit COPY of an fp32 result may carry a power-of-two scale (see out16_scale)
                    const float s16 = has32 ? p.out16_scale : 1.f;
#pragma unroll
                    for (int j = 0; j < CH / 8; ++j)
                        sts_128(buf16 + row16 + ((static_cast<uint32_t>(j) ^ swz16) << 4),
                                Half16<InT>::pack(v[8 * j + 0] * s16, v[8 * j + 1] * s16),
                                Half16<InT>::pack(v[8 * j + 2] * s16, v[8 * j + 3] * s16),
                                Half16<InT>::pack(v[8 * j + 4] * s16, v[8 * j + 5] * s16),
                                Half16<InT>::pack(v[8 * j + 6] * s16, v[8 * j + 7] * s16));
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
#ifdef B200_GEMM_TRACE
                if (tr) B200_TRACE(tslot + 3);  // staged + fenced
#endif
                if (lane == 0) {
                    if (has32) tma_store_2d(tmap_o32, box32, n0, row_base);
                    if (has16) tma_store_2d(tmap_o16, buf16, n0, row_base);
                    bulk_commit_group();
#ifdef B200_GEMM_TRACE
                    if (tr) B200_TRACE(tslot + 4);  // stores issued
#endif
                }
                slot16 ^= 1;
            }
            if constexpr (kChain) {
                // publish this warp's share of the tile to the next GEMM of the chain: its TMA stores have
                // completed (not merely been read out of the staging boxes), its lanes' row partial sums
                // are ordered before lane 0 by the warp barrier, then a gpu-scope release
                if (g + 1 < cp.n_gemm && !(cp.dbg & 8)) {
                    if (lane == 0 && !(cp.dbg & 1)) {
                        bulk_wait_group<0>();
                        fence_proxy_async_all();
                    }
                    __syncwarp();
                    if (lane == 0) red_release_gpu_add(cp.counters + g * cp.num_m + m_blk, 1u);
                }
            }
            if (warp == 4 && lane == 0 && trace_tile < 4) B200_TRACE(11 + 2 * trace_tile);  // tile drained
            if (kChain && warp == 4 && lane == 0 && trace_tile < 24) B200_TRACE(32 + 4 * trace_tile + 3);
            ++trace_tile;
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        // the staging boxes must outlive their readers; completion of the writes themselves is
        // ordered by the end of the grid
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
    }

    // the leader's MMAs read the peer's shared memory: neither CTA may exit early
    tc05_fence_before();
    cluster_sync_relaxed();
    tc05_fence_after();
    if (threadIdx.x == 0) B200_TRACE(20);
    if (warp == 2) tmem_dealloc_2cta<kTmemCols>(tmem_base);
}

template <typename InT, bool kGnStats, int BLOCK_N, bool kChain>
int launch_gemm_tc05_2cta(const ChainMaps& maps, const ChainParams& cp, cudaStream_t stream) {
    auto kern = gemm_tc05_2cta_kernel<InT, kGnStats, BLOCK_N, kChain>;
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Gemm2SmemT<BLOCK_N>::kTotal));
    }
    int clusters = cp.tile_end[cp.n_gemm - 1];
    if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
    if (clusters < 1) return 0;
    B200_CUDA_OK(launch_kernel(kern, dim3(2 * clusters), dim3(kGemm2Threads), Gemm2SmemT<BLOCK_N>::kTotal, stream, maps, cp));
    return 0;
}

}  // namespace b200
