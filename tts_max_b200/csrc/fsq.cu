// K1: FSQ index unpack + codebook / project_out lookup (integer kernel, bit-exact).
//
// Replaces vector_quantize_pytorch.ResidualFSQ.get_output_from_indices
// (vector-quantize-pytorch 1.17.8, not vendored in the reference; built at
// tts/core/codec/decoder_modules.py:418-420, called at tts/core/codec/decoder.py:77):
//   digit_d = (id // 4^d) % 4,  code_d = (digit_d - 2) / 2   in {-1, -0.5, 0, 0.5}
//   out[c]  = project_out(code)[c] = (sum_{d=0..7} code_d * W[c, d]) + b[c]
// Evaluation order (matches torch CPU fp32 bit for bit):
//   acc = 0; for d = 0..7: acc = acc + code_d * W[c, d]; out = acc + b[c]
// The products are exact in fp32 (codes are 0 or +-2^k), so fma(code, w, acc) rounds
// exactly like (code * w) + acc; no fast-math, no reassociation.
//
// HBM-bound: 4 or 8 B in, channels * sizeof(out) B out per token. Each thread keeps the
// 8x8 weight block of its 8 output channels in registers and streams rows; a 256-thread
// CTA writes one full 2048-channel row per iteration with 128-bit stores.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int kFsqThreads = 256;
constexpr int kFsqChanPerThread = 8;
constexpr int kFsqRowsPerBlock = 32;

template <typename OutT>
__device__ __forceinline__ void store8(OutT* dst, const float (&v)[8]);

template <>
__device__ __forceinline__ void store8<float>(float* dst, const float (&v)[8]) {
    float4* d = reinterpret_cast<float4*>(dst);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[8]) {
    uint4 u;
    u.x = Half16<__nv_bfloat16>::pack(v[0], v[1]);
    u.y = Half16<__nv_bfloat16>::pack(v[2], v[3]);
    u.z = Half16<__nv_bfloat16>::pack(v[4], v[5]);
    u.w = Half16<__nv_bfloat16>::pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
}
template <>
__device__ __forceinline__ void store8<__half>(__half* dst, const float (&v)[8]) {
    uint4 u;
    u.x = Half16<__half>::pack(v[0], v[1]);
    u.y = Half16<__half>::pack(v[2], v[3]);
    u.z = Half16<__half>::pack(v[4], v[5]);
    u.w = Half16<__half>::pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
}

template <typename OutT, typename IdT>
__global__ void __launch_bounds__(kFsqThreads)
fsq_lookup_kernel(const IdT* __restrict__ ids, const int32_t* __restrict__ row_tok, int rows,
                  const float* __restrict__ w_out, const float* __restrict__ b_out, int channels,
                  OutT* __restrict__ out, int ld, int* __restrict__ err_flag) {
    pdl_launch_dependents();
    pdl_wait();
    // channel block handled by this thread (grid.y covers channels > 2048 if ever needed)
    const int c0 = (blockIdx.y * kFsqThreads + threadIdx.x) * kFsqChanPerThread;  // < channels (checked by the launcher)

    float w[kFsqChanPerThread][8];
    float b[kFsqChanPerThread];
#pragma unroll
    for (int i = 0; i < kFsqChanPerThread; ++i) {
        const float4* wr = reinterpret_cast<const float4*>(w_out + static_cast<size_t>(c0 + i) * 8);
        const float4 lo = __ldg(wr), hi = __ldg(wr + 1);
        w[i][0] = lo.x; w[i][1] = lo.y; w[i][2] = lo.z; w[i][3] = lo.w;
        w[i][4] = hi.x; w[i][5] = hi.y; w[i][6] = hi.z; w[i][7] = hi.w;
        b[i] = __ldg(b_out + c0 + i);
    }

    // each CTA owns a contiguous slab of rows, so the 72 KB weight block it holds in registers is
    // fetched once per kFsqRowsPerBlock rows. The slab's ids are gathered into shared memory first:
    // the row -> token -> id chain is two dependent global loads, which must not sit inside the
    // serial row loop.
    __shared__ int s_id[kFsqRowsPerBlock];  // -1: halo row, otherwise the 16-bit code id
    const int r_begin = blockIdx.x * kFsqRowsPerBlock;
    const int r_end = min(r_begin + kFsqRowsPerBlock, rows);
    if (threadIdx.x < kFsqRowsPerBlock) {
        const int r = r_begin + threadIdx.x;
        int v = -1;
        if (r < r_end) {
            const int tok = row_tok ? row_tok[r] : r;
            if (tok >= 0) {
                const long long id = static_cast<long long>(ids[tok]);
                if (id < 0 || id > 65535) {
                    if (blockIdx.y == 0) atomicExch(err_flag, 1);
                }
                v = static_cast<int>(static_cast<unsigned>(id) & 0xFFFFu);
            }
        }
        s_id[threadIdx.x] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = r_begin; r < r_end; ++r) {
        const int sid = s_id[r - r_begin];
        float v[kFsqChanPerThread];
        if (sid < 0) {
            // halo row of the padded row space: conv operands must see zeros here
#pragma unroll
            for (int i = 0; i < kFsqChanPerThread; ++i) v[i] = 0.f;
        } else {
            const unsigned uid = static_cast<unsigned>(sid);
            float code[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int digit = (uid >> (2 * d)) & 3;         // (id // 4^d) % 4
                code[d] = static_cast<float>(digit - 2) * 0.5f;  // (digit - half_width) / half_width
            }
#pragma unroll
            for (int i = 0; i < kFsqChanPerThread; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < 8; ++d) acc = __fmaf_rn(code[d], w[i][d], acc);
                v[i] = __fadd_rn(acc, b[i]);
            }
        }
        store8<OutT>(out + static_cast<size_t>(r) * ld + c0, v);
    }
}

template <typename OutT>
int launch_typed(const void* ids, int id_type, const int32_t* row_tok, int rows,
                 const float* w_out, const float* b_out, int channels, void* out, int ld,
                 int* err_flag, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const int chan_blocks = (channels + kFsqThreads * kFsqChanPerThread - 1) /
                            (kFsqThreads * kFsqChanPerThread);
    dim3 grid((rows + kFsqRowsPerBlock - 1) / kFsqRowsPerBlock, chan_blocks);
    if (id_type == 1)
        B200_CUDA_OK(launch_kernel(fsq_lookup_kernel<OutT, long long>, grid, dim3(kFsqThreads), 0, stream,
                                   static_cast<const long long*>(ids), row_tok, rows, w_out, b_out, channels,
                                   static_cast<OutT*>(out), ld, err_flag));
    else
        B200_CUDA_OK(launch_kernel(fsq_lookup_kernel<OutT, int>, grid, dim3(kFsqThreads), 0, stream,
                                   static_cast<const int*>(ids), row_tok, rows, w_out, b_out, channels,
                                   static_cast<OutT*>(out), ld, err_flag));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_fsq_lookup(const void* ids, int id_type, const int32_t* row_tok, int rows,
                      const float* w_out, const float* b_out, int channels, void* out, int ld,
                      int out_prec, int* err_flag, cudaStream_t stream) {
    B200_CHECK(channels % (kFsqChanPerThread * kFsqThreads) == 0, "fsq: channels must be a multiple of %d",
               kFsqChanPerThread * kFsqThreads);
    if (out_prec < 0)
        return launch_typed<float>(ids, id_type, row_tok, rows, w_out, b_out, channels, out, ld,
                                   err_flag, stream);
    if (out_prec == kPrecBf16)
        return launch_typed<__nv_bfloat16>(ids, id_type, row_tok, rows, w_out, b_out, channels,
                                           out, ld, err_flag, stream);
    if (out_prec == kPrecFp16)
        return launch_typed<__half>(ids, id_type, row_tok, rows, w_out, b_out, channels, out, ld,
                                    err_flag, stream);
    set_error("fsq: unsupported output precision %d", out_prec);
    return 1;
}

}  // namespace b200
