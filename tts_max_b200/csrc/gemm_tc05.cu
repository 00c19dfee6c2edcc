// Host side of the tcgen05 GEMM: TMA descriptor encoding, dispatch, weight repacking.
#include "gemm_tc05.cuh"
#include "gemm_tc05_2cta.cuh"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <map>
#include <mutex>
#include <queue>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace b200 {

namespace {

// cuTensorMapEncodeTiled is a driver API; fetch it through the runtime so the library has
// no link-time dependency on libcuda (it must load on a CPU-only box for the symbol tests).
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

template <typename T>
__global__ void repack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int N,
                                     int Cin, int taps, const float* __restrict__ col_scale) {
    // dst[n, tap*Cin + c] = src[n, c, tap]
    const size_t total = static_cast<size_t>(N) * Cin * taps;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cin);
        const size_t rest = i / Cin;
        const int tap = static_cast<int>(rest % taps);
        const size_t n = rest / taps;
        float v = src[(n * Cin + c) * taps + tap];
        if (col_scale != nullptr) v *= col_scale[c];
        dst[i] = Half16<T>::from_float(v);
    }
}

__global__ void fold_rope_kernel(float* __restrict__ w, int heads, int head_dim, int K,
                                 const float* __restrict__ cos_tab,
                                 const float* __restrict__ sin_tab) {
    // rows: r in {q, k}, head h, pair j -> rows (2j, 2j+1); angle index [h, j]
    const int half = head_dim / 2;
    const int pairs = 2 * heads * half;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x) {
        const int j = pr % half;
        const int h = (pr / half) % heads;
        const int r = pr / (half * heads);
        const float c = cos_tab[h * half + j], s = sin_tab[h * half + j];
        float* row0 = w + static_cast<size_t>((r * heads + h) * head_dim + 2 * j) * K;
        float* row1 = row0 + K;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float a = row0[k], b = row1[k];
            row0[k] = a * c - b * s;  // torchtune: x0*cos - x1*sin
            row1[k] = b * c + a * s;  //            x1*cos + x0*sin
        }
    }
}

__global__ void weightnorm_scale_kernel(const float* __restrict__ g, const float* __restrict__ v, int inner,
                                        float* __restrict__ scale) {
    const int c = blockIdx.x;
    float ss = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) {
        const float x = v[static_cast<size_t>(c) * inner + i];
        ss = fmaf(x, x, ss);
    }
    __shared__ float red[32];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) scale[c] = g[c] / sqrtf(t);
    }
}

struct TapIds {
    int id[8];
};

template <typename T>
__global__ void repack_convT_phase_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                          T* __restrict__ dst, int Cin, int Cout, int k, int taps, TapIds ids) {
    const size_t total = static_cast<size_t>(Cout) * taps * Cin;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % Cin);
        const size_t rest = i / Cin;
        const int t = static_cast<int>(rest % taps);
        const size_t co = rest / taps;
        const float x = v[(static_cast<size_t>(ci) * Cout + co) * k + ids.id[t]] * scale[ci];
        dst[i] = Half16<T>::from_float(x);
    }
}

}  // namespace

namespace {
struct TmapKey {
    const void* base;
    uint64_t rows, cols, ld;
    uint32_t dtype, box_cols, box_rows;
    bool operator==(const TmapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && dtype == o.dtype &&
               box_cols == o.box_cols && box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
        auto mix = [&h](uint64_t v) { h = (h ^ v) * 0x100000001B3ull + (h >> 29); };
        mix(k.rows);
        mix(k.cols);
        mix(k.ld);
        mix((static_cast<uint64_t>(k.dtype) << 40) | (static_cast<uint64_t>(k.box_cols) << 20) | k.box_rows);
        return static_cast<size_t>(h);
    }
};
struct alignas(64) TmapSlot {
    CUtensorMap map;
};
}  // namespace

int make_tmap_box(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                  uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows) {
    // a tensor map is a pure function of the key (it holds the address, not the data)
    static thread_local std::unordered_map<TmapKey, TmapSlot, TmapKeyHash> cache;
    const TmapKey key{base, rows, cols, ld_elems, static_cast<uint32_t>(dtype), box_cols, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *map = it->second.map;
        return 0;
    }
    auto fn = get_encode_fn();
    B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    CUtensorMapDataType dt;
    uint64_t esz;
    switch (dtype) {
        case kTmapBf16: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; esz = 2; break;
        case kTmapF16: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; esz = 2; break;
        case kTmapF32: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; esz = 4; break;
        default: set_error("make_tmap_box: bad dtype %d", dtype); return 1;
    }
    CUtensorMapSwizzle swz;
    switch (box_cols * esz) {
        case 128: swz = CU_TENSOR_MAP_SWIZZLE_128B; break;
        case 64: swz = CU_TENSOR_MAP_SWIZZLE_64B; break;
        case 32: swz = CU_TENSOR_MAP_SWIZZLE_32B; break;
        default: set_error("make_tmap_box: box row of %u bytes", (unsigned)(box_cols * esz)); return 1;
    }
    B200_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    B200_CHECK((ld_elems * esz) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
    B200_CHECK(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * esz};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    if (cache.size() >= 8192) cache.clear();  // shapes churn (varlen batches): bound the memo
    cache[key].map = *map;
    return 0;
}

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows) {
    const uint32_t esz = dtype == kTmapF32 ? 4 : 2;
    return make_tmap_box(map, base, dtype, rows, cols, ld_elems, 128 / esz, box_rows);
}

// Every GEMM runs on the CTA-pair kernel; N must tile by 64 (the callers pad: head.out 1282 -> 1536, the
// encoder's 48 / 96 channels -> 64 / 128). For every M, so an utterance sees the same arithmetic alone or
// inside a batch.
bool gemm_uses_cta_pairs(int /*a_rows*/, int n_store) { return n_store >= 64 && n_store % 64 == 0; }

#ifdef B200_GEMM_TRACE
unsigned long long* g_gemm_trace = nullptr;  // set by tools/gemm_trace.cu
#endif

int g_gemm_narrow_tiles = 1;  // A/B: 0 = 256-wide tiles only
int g_gemm_early_weights = 1; // A/B: weight loads of a launch's first stages before griddepcontrol.wait

namespace {

// validation and GemmParams of one call (shared by plain launches and chains)
int fill_params(const GemmCall& c, GemmParams* out) {
    B200_CHECK(c.precision == kPrecBf16 || c.precision == kPrecFp16,
               "gemm: unsupported precision %d", c.precision);
    const int block_k = 64;
    B200_CHECK(c.Cin % block_k == 0, "gemm: Cin (%d) must be a multiple of %d", c.Cin, block_k);
    B200_CHECK(c.n_store % 32 == 0 && (c.out_cols > 0 ? c.out_cols : c.n_store) <= c.ldc, "gemm: bad n_store %d (ldc %d)",
               c.n_store, c.ldc);
    B200_CHECK(c.a_cols >= 0 && c.a_cols <= c.Cin && c.out_cols >= 0 && c.out_cols <= c.n_store && c.a_cols % 8 == 0 &&
                   c.out_cols % 8 == 0, "gemm: bad a_cols %d / out_cols %d", c.a_cols, c.out_cols);
    B200_CHECK(c.out_cols == 0 || (c.out16 == nullptr && c.ss_out == nullptr && c.gn_stats == nullptr),
               "gemm: out_cols is for plain epilogues");
    B200_CHECK(c.taps >= 1 && (c.tap_pad >= 0 || c.taps % 2 == 1), "gemm: even tap counts need an explicit tap_pad");
    B200_CHECK(c.residual == nullptr || c.out_fp32, "gemm: a residual requires fp32 output");
    GemmParams p;
    p.M = c.a_rows;
    p.n_store = c.n_store;
    p.k_blocks_per_tap = c.Cin / block_k;
    p.taps = c.taps;
    B200_CHECK(c.tap_dil >= 1, "gemm: tap_dil must be >= 1");
    p.tap_dil = c.tap_dil;
    p.tap_pad = c.tap_pad >= 0 ? c.tap_pad : (c.taps / 2) * c.tap_dil;
    p.out = c.out;
    p.ldc = c.ldc;
    p.bias = c.bias;
    p.residual = c.residual;
    p.ld_res = c.ld_res;
    p.row_valid = c.row_valid;
    p.act = c.act;
    p.ss_in = c.ss_in;
    p.ss_inv_dim = c.ss_inv_dim;
    p.ss_eps = c.ss_eps;
    p.out16 = c.out16;
    p.ld16 = c.ld16;
    p.ss_out = c.ss_out;
    p.out16_scale = c.out16_scale;
    p.ss_in_scale = c.ss_in_scale;
#ifdef B200_GEMM_TRACE
    p.trace = g_gemm_trace;
#endif
    B200_CHECK(c.n_store >= 64 && c.n_store % 64 == 0, "gemm: N (%d stored columns) must be a multiple of 64", c.n_store);
    B200_CHECK(c.out16 == nullptr || c.out_fp32, "gemm: out16 requires fp32 output");
    B200_CHECK(c.ss_out == nullptr || (c.out_fp32 && c.n_store == 1024),
               "gemm: ss_out requires fp32 output with N == 1024");
    B200_CHECK(c.residual == nullptr || ((reinterpret_cast<uintptr_t>(c.residual) & 31) == 0 && c.ld_res % 8 == 0),
               "gemm: the residual must be 32-byte aligned with a row pitch that is a multiple of 8");
    p.gn_stats = c.gn_stats;
    p.gn_row_utt = c.gn_row_utt;
    B200_CHECK(c.gn_stats == nullptr || (c.out_fp32 && c.n_store == 1024 && c.gn_row_utt != nullptr),
               "gemm: GroupNorm statistics need fp32 output, N == 1024 and a row -> utterance map");
    p.has32 = c.out_fp32 ? 1 : 0;
    p.has16 = (!c.out_fp32 || c.out16 != nullptr) ? 1 : 0;
    *out = p;
    return 0;
}

// Small M: 256 x 64 tiles when the 256-wide tiling would leave most CTA pairs without a tile
// (B = 1 serving, config 1). A narrow tile costs ~0.6 of a wide one (its K loop is bounded by
// the A-tile load), so it pays when four times the tiles still fit in fewer weighted rounds.
bool want_narrow(int a_rows, int n_store) {
    const int pairs = kNumSMs / 2;
    const int tiles256 = ((a_rows + 255) / 256) * (n_store / 256);
    const int rounds256 = (tiles256 + pairs - 1) / pairs;
    const int rounds64 = (4 * tiles256 + pairs - 1) / pairs;
    return g_gemm_narrow_tiles && rounds64 * 60 < rounds256 * 95;
}

// tensor maps of one CTA-pair GEMM: A, B (box of block_n / 2 weight rows), fp32 / 16-bit output boxes
int fill_maps_2cta(const GemmCall& c, int block_n, CUtensorMap* ta, CUtensorMap* tb, CUtensorMap* to32,
                   CUtensorMap* to16) {
    const int dt16 = c.precision == kPrecBf16 ? kTmapBf16 : kTmapF16;
    const int a_cols = c.a_cols > 0 ? c.a_cols : c.Cin;
    const int o_cols = c.out_cols > 0 ? c.out_cols : c.n_store;
    if (c.tmap_a != nullptr) *ta = *static_cast<const CUtensorMap*>(c.tmap_a);
    else if (make_tmap_2d(ta, c.a, dt16, c.a_rows, a_cols, a_cols, kGemmBlockM)) return 1;
    if (c.tmap_b != nullptr && block_n == 256) *tb = *static_cast<const CUtensorMap*>(c.tmap_b);
    else if (make_tmap_2d(tb, c.w, dt16, c.N, static_cast<uint64_t>(c.taps) * c.Cin,
                          static_cast<uint64_t>(c.taps) * c.Cin, block_n / 2))
        return 1;
    // the epilogue stores with TMA: fp32 boxes of 32 x 32 (128-byte rows), 16-bit output
    // (or operand copy) boxes of 32 x 32 (64-byte rows)
    *to32 = *ta;  // placeholders for the map this GEMM does not use
    *to16 = *ta;
    if (c.out_fp32) {
        if (make_tmap_box(to32, c.out, kTmapF32, c.a_rows, o_cols, c.ldc, kGemm2ChunkCols, 32)) return 1;
        if (c.out16 != nullptr &&
            make_tmap_box(to16, c.out16, dt16, c.a_rows, c.n_store, c.ld16, kGemm2ChunkCols, 32))
            return 1;
    } else {
        if (make_tmap_box(to16, c.out, dt16, c.a_rows, o_cols, c.ldc, kGemm2ChunkCols, 32)) return 1;
    }
    return 0;
}

template <bool kChain>
int launch_2cta(int precision, bool gn, int block_n, const ChainMaps& maps, const ChainParams& cp, cudaStream_t s) {
    const bool bf = precision == kPrecBf16;
    if (block_n == 64) {
        if (gn) return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, true, 64, kChain>(maps, cp, s)
                          : launch_gemm_tc05_2cta<__half, true, 64, kChain>(maps, cp, s);
        return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, false, 64, kChain>(maps, cp, s)
                  : launch_gemm_tc05_2cta<__half, false, 64, kChain>(maps, cp, s);
    }
    if constexpr (!kChain) {  // the encoder's channel widths: plain launches without GroupNorm statistics
        if (block_n == 128 && !gn)
            return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, false, 128, false>(maps, cp, s)
                      : launch_gemm_tc05_2cta<__half, false, 128, false>(maps, cp, s);
        if (block_n == 192 && !gn)
            return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, false, 192, false>(maps, cp, s)
                      : launch_gemm_tc05_2cta<__half, false, 192, false>(maps, cp, s);
    }
    B200_CHECK(block_n == 256, "gemm: no %d-wide instantiation for this launch", block_n);
    if (gn) return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, true, 256, kChain>(maps, cp, s)
                      : launch_gemm_tc05_2cta<__half, true, 256, kChain>(maps, cp, s);
    return bf ? launch_gemm_tc05_2cta<__nv_bfloat16, false, 256, kChain>(maps, cp, s)
              : launch_gemm_tc05_2cta<__half, false, 256, kChain>(maps, cp, s);
}

// Tile width of a CTA-pair GEMM whose N is not a multiple of 256 (the encoder's convs): the widest of
// 192 / 128 / 64 that divides N, unless that leaves CTA pairs idle that a narrower tiling would use.
// Cost per launch ~ rounds x (K blocks x time per K block + fixed), floored by the L2 traffic of re-reading
// the A operand once per n-tile.
int pick_block_n(const GemmCall& c) {
    const int pairs = kNumSMs / 2;
    const int m_blocks = (c.a_rows + 255) / 256;
    const double kb = static_cast<double>(c.taps) * (c.Cin / 64);
    int best = 64;
    double best_cost = 1e30;
    for (int bn : {64, 128, 192}) {
        if (c.n_store % bn != 0) continue;
        if (bn != 64 && (c.gn_stats != nullptr)) continue;
        const int tiles = m_blocks * (c.n_store / bn);
        const int rounds = (tiles + pairs - 1) / pairs;
        const double t_kb = bn == 64 ? 0.15 : bn == 128 ? 0.16 : 0.20;          // us per 64-deep K block
        const double t_tile = kb * t_kb + 2.0 + 0.012 * bn;                       // + prologue / epilogue
        const double a_bytes = static_cast<double>(c.a_rows) * c.Cin * 2.0 * c.taps * (c.n_store / bn);
        const double cost = std::max(rounds * t_tile, a_bytes / 6.0e6);           // ~6 TB/s from L2
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace

int launch_gemm(const GemmCall& c, cudaStream_t stream) {
    GemmParams p;
    if (fill_params(c, &p)) return 1;
    if (c.a_rows <= 0) return 0;
    int block_n = 256;
    if (c.n_store % 256 != 0) block_n = g_gemm_narrow_tiles == 2 ? 64 : pick_block_n(c);
    else if (c.tmap_b == nullptr && want_narrow(c.a_rows, c.n_store)) block_n = 64;
    alignas(64) ChainMaps maps;
    ChainParams cp{};
    if (fill_maps_2cta(c, block_n, &maps.a[0], &maps.b[0], &maps.o32[0], &maps.o16[0])) return 1;
    cp.n_gemm = 1;
    cp.num_m = (c.a_rows + 255) / 256;
    cp.num_n[0] = c.n_store / block_n;
    cp.tile_end[0] = cp.num_m * cp.num_n[0];
    cp.full[0] = 0;
    cp.counters = nullptr;
    cp.early_weights = g_gemm_early_weights;
    cp.g[0] = p;
    return launch_2cta<false>(c.precision, p.gn_stats != nullptr, block_n, maps, cp, stream);
}

namespace {

// Per-cluster tile lists of a chain (see ChainParams::sched): tiles are handed out in the global order,
// each to the cluster that becomes free first under the cost model "K blocks + a fixed turnaround". Every
// list is increasing, which is what the chain's no-deadlock argument needs. Lists depend on the shapes
// only, so they are built once per shape and cached on the device (never freed: a few KB each).
struct SchedKey {
    int dev, n, clusters, num_m, num_n[kChainMax], kb[kChainMax];
    bool operator<(const SchedKey& o) const { return std::memcmp(this, &o, sizeof(SchedKey)) < 0; }
};
struct SchedBuf {
    int32_t* dev = nullptr;
    int stride = 0;
};

int chain_schedule(const ChainParams& cp, const int* kb, int clusters, SchedBuf* out) {
    static std::map<SchedKey, SchedBuf> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    SchedKey key;
    std::memset(&key, 0, sizeof(key));
    cudaGetDevice(&key.dev);
    key.n = cp.n_gemm;
    key.clusters = clusters;
    key.num_m = cp.num_m;
    for (int i = 0; i < cp.n_gemm; ++i) {
        key.num_n[i] = cp.num_n[i];
        key.kb[i] = kb[i];
    }
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return 0;
    }
    const int total = cp.tile_end[cp.n_gemm - 1];
    std::vector<std::vector<int32_t>> lists(clusters);
    // (time the cluster becomes free, cluster): smallest first, ties to the lower cluster id
    using Slot = std::pair<long long, int>;
    std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> free_at;
    for (int c = 0; c < clusters; ++c) free_at.push({0, c});
    int g = 0;
    for (int t = 0; t < total; ++t) {
        while (t >= cp.tile_end[g]) ++g;
        Slot sl = free_at.top();
        free_at.pop();
        lists[sl.second].push_back(t);
        free_at.push({sl.first + kb[g] + 2, sl.second});
    }
    size_t longest = 0;
    for (auto& l : lists) longest = std::max(longest, l.size());
    SchedBuf buf;
    buf.stride = static_cast<int>(longest) + 1;
    std::vector<int32_t> host(static_cast<size_t>(clusters) * buf.stride, -1);
    for (int c = 0; c < clusters; ++c) std::copy(lists[c].begin(), lists[c].end(), host.begin() + static_cast<size_t>(c) * buf.stride);
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&buf.dev), host.size() * sizeof(int32_t)));
    B200_CUDA_OK(cudaMemcpy(buf.dev, host.data(), host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (cache.size() >= 4096) cache.clear();  // the buffers stay allocated (a few KB each); only the memo is bounded
    cache[key] = buf;
    *out = buf;
    return 0;
}

}  // namespace

bool gemm_chain_supported(const GemmCall* calls, int n) {
    if (n < 2 || n > kChainMax) return false;
    for (int i = 0; i < n; ++i) {
        const GemmCall& c = calls[i];
        if (c.n_store % 256 != 0 || c.taps != 1 || c.gn_stats != nullptr || c.a_rows != calls[0].a_rows ||
            c.precision != calls[0].precision || c.row_valid != nullptr)
            return false;
    }
    return true;
}

int launch_gemm_chain(const GemmCall* calls, int n, uint32_t* counters, cudaStream_t stream) {
    B200_CHECK(gemm_chain_supported(calls, n), "gemm chain: 2 to %d plain linears (no taps, no GroupNorm statistics) "
               "on the CTA-pair kernel with a common row count", kChainMax);
    B200_CHECK(counters != nullptr, "gemm chain: null counters");
    if (calls[0].a_rows <= 0) return 0;
    const bool narrow = want_narrow(calls[0].a_rows, 1024);
    const int block_n = narrow ? 64 : 256;
    alignas(64) ChainMaps maps;
    ChainParams cp{};
    cp.n_gemm = n;
    cp.num_m = (calls[0].a_rows + 255) / 256;
    cp.counters = counters;
    int end = 0;
    for (int i = 0; i < n; ++i) {
        if (fill_params(calls[i], &cp.g[i])) return 1;
        if (fill_maps_2cta(calls[i], block_n, &maps.a[i], &maps.b[i], &maps.o32[i], &maps.o16[i])) return 1;
        cp.num_n[i] = calls[i].n_store / block_n;
        end += cp.num_m * cp.num_n[i];
        cp.tile_end[i] = end;
        cp.full[i] = 16u * static_cast<uint32_t>(cp.num_n[i]);  // 8 epilogue warps x 2 CTAs per tile
    }
    int kb[kChainMax] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) kb[i] = cp.g[i].taps * cp.g[i].k_blocks_per_tap;
    int clusters = end < kNumSMs / 2 ? end : kNumSMs / 2;
    SchedBuf sched;
    if (chain_schedule(cp, kb, clusters, &sched)) return 1;
    cp.sched = sched.dev;
    cp.sched_stride = sched.stride;
    cp.early_weights = g_gemm_early_weights;
    {
        static const char* e = getenv("B200_CHAIN_DBG");  // timing experiments (tools), never set in production
        cp.dbg = e ? atoi(e) : 0;
    }
    for (int i = n; i < kChainMax; ++i) {  // unused slots: valid descriptors, never indexed
        maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; maps.o32[i] = maps.o32[0]; maps.o16[i] = maps.o16[0];
        cp.tile_end[i] = end;
        cp.num_n[i] = 1;
    }
    return launch_2cta<true>(calls[0].precision, false, narrow ? 64 : 256, maps, cp, stream);
}

int launch_repack_weight(int prec, const float* src, void* dst, int N, int Cin, int taps,
                         cudaStream_t stream, const float* col_scale) {
    const size_t total = static_cast<size_t>(N) * Cin * taps;
    if (total == 0) return 0;
    const int grid = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (prec == kPrecBf16)
        repack_weight_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
            src, static_cast<__nv_bfloat16*>(dst), N, Cin, taps, col_scale);
    else if (prec == kPrecFp16)
        repack_weight_kernel<__half><<<grid, 256, 0, stream>>>(src, static_cast<__half*>(dst), N,
                                                               Cin, taps, col_scale);
    else {
        set_error("repack: unsupported precision %d", prec);
        return 1;
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_weightnorm_scale(const float* g, const float* v, int C0, int inner, float* scale, cudaStream_t stream) {
    weightnorm_scale_kernel<<<C0, 256, 0, stream>>>(g, v, inner, scale);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_repack_convT_phase(int prec, const float* v, const float* scale, void* dst, int Cin, int Cout, int k,
                              int taps, const int* tap_ids, cudaStream_t stream) {
    B200_CHECK(taps >= 1 && taps <= 8, "convT repack: bad tap count %d", taps);
    TapIds ids;
    for (int t = 0; t < 8; ++t) ids.id[t] = t < taps ? tap_ids[t] : 0;
    const size_t total = static_cast<size_t>(Cout) * taps * Cin;
    const int grid = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (prec == kPrecBf16)
        repack_convT_phase_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(v, scale, static_cast<__nv_bfloat16*>(dst),
                                                                           Cin, Cout, k, taps, ids);
    else if (prec == kPrecFp16)
        repack_convT_phase_kernel<__half><<<grid, 256, 0, stream>>>(v, scale, static_cast<__half*>(dst), Cin, Cout, k,
                                                                    taps, ids);
    else {
        set_error("convT repack: unsupported precision %d", prec);
        return 1;
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_fold_rope(float* w_qkv, int heads, int head_dim, int K, const float* cos_tab,
                     const float* sin_tab, cudaStream_t stream) {
    const int pairs = 2 * heads * (head_dim / 2);
    fold_rope_kernel<<<pairs, 256, 0, stream>>>(w_qkv, heads, head_dim, K, cos_tab, sin_tab);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
