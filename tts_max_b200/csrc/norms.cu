// K5 / K11 / K4-norm: RMSNorm, LayerNorm and GroupNorm(32)+swish on token-major
// activations [rows, dim] fp32 -> tensor-core operand dtype.
//
// Replaces (reference):
//   RMSNorm.forward                tts/core/codec/decoder_modules.py:233-236
//   final_layer_norm (LayerNorm)   tts/core/codec/decoder_modules.py:373,399
//   Normalize (GroupNorm 32, eps 1e-6, affine) + nonlinearity (swish)
//                                  tts/core/codec/decoder_modules.py:151-159, 204-205, 212-213
//
// All three are HBM/L2-bound (4 B in + 2 B out per element): one warp per row with
// 128-bit loads and stores for the row norms; GroupNorm statistics span (32 channels x T)
// per utterance, so it is two-phase: a partial-sum kernel with fp64 atomics per
// (utterance, group), then an apply+swish kernel that also writes the zero halo rows the
// following implicit-GEMM conv relies on.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int kNormWarps = 8;

template <typename OutT>
__device__ __forceinline__ void store8_out(OutT* dst, const float* v) {
    uint4 u;
    u.x = Half16<OutT>::pack(v[0], v[1]);
    u.y = Half16<OutT>::pack(v[2], v[3]);
    u.z = Half16<OutT>::pack(v[4], v[5]);
    u.w = Half16<OutT>::pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
}

// dim == kChunks * 256; lane owns elements [lane*8 + i*256, +8) for i < kChunks
template <typename OutT, int kChunks, bool kLayerNorm>
__global__ void __launch_bounds__(kNormWarps * 32)
rownorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
               const float* __restrict__ b, int rows, float eps, OutT* __restrict__ out,
               const uint8_t* __restrict__ row_valid) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int dim = kChunks * 256;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * kNormWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (row_valid != nullptr && row_valid[row] == 0) {  // halo row: zeros for the conv that follows
        OutT* zrow = out + static_cast<size_t>(row) * dim;
#pragma unroll
        for (int i = 0; i < kChunks; ++i)
            *reinterpret_cast<uint4*>(zrow + i * 256 + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const float* xr = x + static_cast<size_t>(row) * dim;
    float v[kChunks][8];
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
        const float4* p = reinterpret_cast<const float4*>(xr + i * 256 + lane * 8);
        const float4 a = p[0], c = p[1];
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = c.x; v[i][5] = c.y; v[i][6] = c.z; v[i][7] = c.w;
    }
    float mean = 0.f;
    if (kLayerNorm) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kChunks; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        mean = warp_sum(s) * (1.f / dim);
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kChunks; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = v[i][j] - mean;
            ss = fmaf(d, d, ss);
        }
    const float var = warp_sum(ss) * (1.f / dim);
    const float rstd = rsqrtf(var + eps);
    OutT* orow = out + static_cast<size_t>(row) * dim;
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
        const int c0 = i * 256 + lane * 8;
        float wv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (w != nullptr) {
            const float4* wp = reinterpret_cast<const float4*>(w + c0);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
            wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
        }
        float o[8];
        if (kLayerNorm) {
            const float4* bp = reinterpret_cast<const float4*>(b + c0);
            const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * wv[j] + bv[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rstd * wv[j];
        }
        store8_out<OutT>(orow + c0, o);
    }
}

template <bool kLayerNorm>
int launch_rownorm(int prec, const float* x, const float* w, const float* b, int rows, int dim,
                   float eps, void* out, cudaStream_t stream, const uint8_t* row_valid) {
    B200_CHECK(dim == 1024, "row norm: only dim == 1024 is instantiated (got %d)", dim);
    if (rows <= 0) return 0;
    const int grid = (rows + kNormWarps - 1) / kNormWarps;
    if (prec == kPrecBf16)
        B200_CUDA_OK(launch_kernel(rownorm_kernel<__nv_bfloat16, 4, kLayerNorm>, dim3(grid), dim3(kNormWarps * 32), 0,
                                   stream, x, w, b, rows, eps, static_cast<__nv_bfloat16*>(out), row_valid));
    else if (prec == kPrecFp16)
        B200_CUDA_OK(launch_kernel(rownorm_kernel<__half, 4, kLayerNorm>, dim3(grid), dim3(kNormWarps * 32), 0, stream,
                                   x, w, b, rows, eps, static_cast<__half*>(out), row_valid));
    else {
        set_error("row norm: unsupported precision %d", prec);
        return 1;
    }
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// GroupNorm: 32 groups; thread t of 256 owns channels [4t, 4t+4) -> group t / (gsz/4)
// ---------------------------------------------------------------------------
constexpr int kGnThreads = 256;
constexpr int kGnRowsPerBlock = 16;

// One CTA reduces rows [q0 + 16*y, +16) of one utterance (work item {row0, T, q0} shared with
// the attention tile list). The chunking is relative to the utterance start, so the fp32
// partial sums -- and with them the whole decode -- do not depend on where in a batch the
// utterance sits (row-of-batch == single decode).
// DIM channels, 32 groups of DIM / 32 channels; a thread owns 4 channels, DIM / 4 threads cover a row,
// a 256-thread CTA handles 1024 / DIM rows per step.
template <int DIM>
__global__ void __launch_bounds__(kGnThreads)
groupnorm_stats_kernel(const float* __restrict__ x, const int4* __restrict__ work,
                       double* __restrict__ stats, const int32_t* __restrict__ row_utt) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int kTpr = DIM / 4;            // threads per row
    constexpr int kRows = kGnThreads / kTpr;  // rows per step
    constexpr int kTpg = DIM / 32 / 4;        // threads per group (8, 4, 2)
    const int4 wk = work[blockIdx.x];
    const int t0 = wk.z + blockIdx.y * kGnRowsPerBlock;
    const int t1 = min(t0 + kGnRowsPerBlock, min(wk.z + kAttnBlockQ, wk.y));
    if (t0 >= t1) return;
    const int utt = row_utt[wk.x];
    const int tc = threadIdx.x % kTpr;  // channel thread
    const int tr = threadIdx.x / kTpr;  // row slot
    const int group = tc / kTpg;
    float s = 0.f, ss = 0.f;
    for (int t = t0 + tr; t < t1; t += kRows) {
        const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(wk.x + t) * DIM + tc * 4);
        s += (v.x + v.y) + (v.z + v.w);
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
        ss = fmaf(v.z, v.z, ss);
        ss = fmaf(v.w, v.w, ss);
    }
#pragma unroll
    for (int o = kTpg / 2; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((tc % kTpg) == 0) {
        atomicAdd(stats + (static_cast<size_t>(utt) * 32 + group) * 2 + 0, static_cast<double>(s));
        atomicAdd(stats + (static_cast<size_t>(utt) * 32 + group) * 2 + 1, static_cast<double>(ss));
    }
}

// y = swish((x - mean) * rstd * gamma + beta) -> operand dtype; halo rows are written as zeros
// (the conv that follows reads them as its padding). DIM / 8 threads per row, 8 channels per thread:
// 2 x 128-bit loads, one 128-bit store; a 256-thread CTA streams 2048 / DIM rows per iteration.
template <typename OutT, int DIM>
__global__ void __launch_bounds__(kGnThreads)
groupnorm_apply_swish_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_utt,
                             int rows, const double* __restrict__ stats,
                             const int32_t* __restrict__ utt_len, float eps,
                             const float* __restrict__ gamma, const float* __restrict__ beta,
                             OutT* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int kTpr = DIM / 8;             // threads per row (128, 64, 32)
    constexpr int kRows = kGnThreads / kTpr;  // rows per iteration
    const int t = threadIdx.x % kTpr;
    const int sub = threadIdx.x / kTpr;
    const int c0 = t * 8;
    // group size DIM / 32 >= 8: a thread's 8 channels sit in one group; DIM = 128 (groups of 4): in two
    constexpr bool kTwoGroups = DIM / 32 < 8;
    static_assert(DIM / 32 >= 4, "GroupNorm(32) kernels: at least 128 channels");
    const int group = c0 / (DIM / 32);
    float g[8], b[8];
    {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    }
    // a CTA walks a contiguous slab of rows, so (mean, rstd) of (utterance, group) -- fp64 from the
    // accumulated (sum, sumsq), exactly what the former finalize kernel computed -- is refreshed
    // only when the utterance changes
    const int slab = (rows + gridDim.x - 1) / gridDim.x;
    const int r_lo = blockIdx.x * slab;
    const int r_hi = min(r_lo + slab, rows);
    int cur_u = -1;
    float2 mr = make_float2(0.f, 0.f), mr2 = mr;  // (mean, rstd) of the group(s) of this thread's channels
    // two rows per step with all four 128-bit loads issued before anything depends on them (the
    // pass is HBM-bound: bytes in flight per thread are what it needs)
    auto finish_row = [&](int r, int u, const float4& v0, const float4& v1) {
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
        if (u >= 0) {
            if (u != cur_u) {
                cur_u = u;
                const double cnt = static_cast<double>(utt_len[u]) * (DIM / 32);
                auto mean_rstd = [&](int grp) {
                    const double m = stats[(static_cast<size_t>(u) * 32 + grp) * 2 + 0] / cnt;
                    double var = stats[(static_cast<size_t>(u) * 32 + grp) * 2 + 1] / cnt - m * m;
                    var = var < 0.0 ? 0.0 : var;
                    return make_float2(static_cast<float>(m), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
                };
                mr = mean_rstd(group);
                mr2 = kTwoGroups ? mean_rstd(group + 1) : mr;
            }
            float y[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 q = (kTwoGroups && j >= 4) ? mr2 : mr;
                const float n = (y[j] - q.x) * q.y * g[j] + b[j];
                y[j] = __fdividef(n, 1.f + __expf(-n));  // x * sigmoid(x)
            }
            packed.x = Half16<OutT>::pack(y[0], y[1]);
            packed.y = Half16<OutT>::pack(y[2], y[3]);
            packed.z = Half16<OutT>::pack(y[4], y[5]);
            packed.w = Half16<OutT>::pack(y[6], y[7]);
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * DIM + c0) = packed;
    };
    for (int r = r_lo + sub; r < r_hi; r += 2 * kRows) {
        const int r2 = r + kRows;
        const bool has2 = r2 < r_hi;
        const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * DIM + c0);
        const float4* xq = reinterpret_cast<const float4*>(x + static_cast<size_t>(has2 ? r2 : r) * DIM + c0);
        const float4 a0 = xp[0], a1 = xp[1], b0 = xq[0], b1 = xq[1];
        const int u1 = row_utt[r];
        const int u2 = has2 ? row_utt[r2] : -1;
        finish_row(r, u1, a0, a1);
        if (has2) finish_row(r2, u2, b0, b1);
    }
}

template <int DIM>
int launch_gn_typed(int prec, const float* x, const RowSpace& rs, const double* stats, const float* gamma,
                    const float* beta, float eps, void* out, cudaStream_t stream,
                    bool stats_only, double* stats_out) {
    if (stats_only) {
        if (rs.n_attn_work <= 0) return 0;
        static_assert(kAttnBlockQ % kGnRowsPerBlock == 0, "GroupNorm chunks must tile the work item");
        dim3 grid(rs.n_attn_work, kAttnBlockQ / kGnRowsPerBlock);
        B200_CUDA_OK(launch_kernel(groupnorm_stats_kernel<DIM>, grid, dim3(kGnThreads), 0, stream, x,
                                   rs.attn_work, stats_out, rs.row_utt));
        return 0;
    }
    if (rs.rows <= 0) return 0;
    constexpr int kRows = kGnThreads / (DIM / 8);
    // one wave of 4 CTAs per SM (64 registers): a CTA's slab is long enough to amortise its
    // gamma / beta / statistics prologue (ncu: 1184 short CTAs ran as two half-empty waves)
    int grid = (rs.rows + kRows - 1) / kRows;
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    if (prec == kPrecBf16)
        B200_CUDA_OK(launch_kernel(groupnorm_apply_swish_kernel<__nv_bfloat16, DIM>, dim3(grid), dim3(kGnThreads), 0,
                                   stream, x, rs.row_utt, rs.rows, stats, rs.utt_len, eps, gamma,
                                   beta, static_cast<__nv_bfloat16*>(out)));
    else if (prec == kPrecFp16)
        B200_CUDA_OK(launch_kernel(groupnorm_apply_swish_kernel<__half, DIM>, dim3(grid), dim3(kGnThreads), 0, stream,
                                   x, rs.row_utt, rs.rows, stats, rs.utt_len, eps, gamma, beta,
                                   static_cast<__half*>(out)));
    else {
        set_error("groupnorm: unsupported precision %d", prec);
        return 1;
    }
    return 0;
}

int launch_gn_dispatch(int dim, int prec, const float* x, const RowSpace& rs, const double* stats,
                       const float* gamma, const float* beta, float eps, void* out, cudaStream_t stream,
                       bool stats_only, double* stats_out) {
    switch (dim) {
        case 1024: return launch_gn_typed<1024>(prec, x, rs, stats, gamma, beta, eps, out, stream, stats_only, stats_out);
        case 512: return launch_gn_typed<512>(prec, x, rs, stats, gamma, beta, eps, out, stream, stats_only, stats_out);
        case 256: return launch_gn_typed<256>(prec, x, rs, stats, gamma, beta, eps, out, stream, stats_only, stats_out);
        case 128: return launch_gn_typed<128>(prec, x, rs, stats, gamma, beta, eps, out, stream, stats_only, stats_out);
        default:
            set_error("groupnorm: dim %d is not instantiated (1024, 512, 256, 128)", dim);
            return 1;
    }
}

}  // namespace

int launch_rmsnorm(int prec, const float* x, const float* w, int rows, int dim, float eps,
                   void* out, cudaStream_t stream) {
    return launch_rownorm<false>(prec, x, w, nullptr, rows, dim, eps, out, stream, nullptr);
}

int launch_layernorm(int prec, const float* x, const float* w, const float* b, int rows, int dim,
                     float eps, void* out, cudaStream_t stream, const uint8_t* row_valid) {
    return launch_rownorm<true>(prec, x, w, b, rows, dim, eps, out, stream, row_valid);
}

int launch_groupnorm_stats(const float* x, const RowSpace& rs, int dim, double* stats,
                           cudaStream_t stream) {
    return launch_gn_dispatch(dim, kPrecBf16, x, rs, nullptr, nullptr, nullptr, 0.f, nullptr, stream, true, stats);
}

int launch_groupnorm_apply_swish(int prec, const float* x, const RowSpace& rs, int dim,
                                 const double* stats, const float* gamma, const float* beta,
                                 float eps, void* out, cudaStream_t stream) {
    return launch_gn_dispatch(dim, prec, x, rs, stats, gamma, beta, eps, out, stream, false, nullptr);
}

}  // namespace b200
