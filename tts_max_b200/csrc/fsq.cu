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
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * kFsqChanPerThread;  // < channels (checked by the launcher)

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

// ---------------------------------------------------------------------------
// Fused front end: ids -> embed(fc_post_a(project_out(code)))  (decoder.py:77-79, decoder_modules.py:340,392)
//
// Nothing non-linear separates the codebook projection (8 -> 2048), fc_post_a (2048 -> 1024) and the
// backbone's first Conv1d (k = 7, "same"): frame t of the conv output is a linear function of the
// seven neighbouring codes,
//   y_t[c] = b_e[c] + sum_{tap : frame t+tap-3 exists} ( cb[tap][c] + sum_d M[tap][c][d] * code_{t+tap-3}[d] )
// with M[tap] = W_e[:, :, tap] W_fc W_out  ([1024, 8]) and cb[tap] = W_e[:, :, tap] (W_fc b_out + b_fc)
// folded in fp64 at load time. Zero padding at the utterance edges drops the whole tap term (the
// reference pads fc_post_a's OUTPUT with zeros, so the bias part disappears with it).
// One thread owns one output channel (56 + 7 coefficients in registers) and streams a slab of
// rows; the slab's codes (+3 rows each side) are decoded once into shared memory. Fixed
// evaluation order (tap-major, digit-minor): batch-invariant. HBM-bound: 8 B in, 4 096 B out.
constexpr int kFrontRows = 63;  // a multiple of the 7-row register window
constexpr int kFrontThreads = 256;

template <typename IdT>
__global__ void __launch_bounds__(kFrontThreads)
fsq_frontend_kernel(const IdT* __restrict__ ids, const int32_t* __restrict__ row_tok, int rows,
                    const float* __restrict__ m_fold,   // [7][C][8]
                    const float* __restrict__ cb_fold,  // [7][C]
                    const float* __restrict__ b_embed,  // [C]
                    int C, float* __restrict__ x, int* __restrict__ err_flag) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.y * kFrontThreads + threadIdx.x;
    float m[7][8], cb[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
        const float4* mr = reinterpret_cast<const float4*>(m_fold + (static_cast<size_t>(t) * C + c) * 8);
        const float4 lo = __ldg(mr), hi = __ldg(mr + 1);
        m[t][0] = lo.x; m[t][1] = lo.y; m[t][2] = lo.z; m[t][3] = lo.w;
        m[t][4] = hi.x; m[t][5] = hi.y; m[t][6] = hi.z; m[t][7] = hi.w;
        cb[t] = __ldg(cb_fold + static_cast<size_t>(t) * C + c);
    }
    const float be = __ldg(b_embed + c);

    // codes of rows [r_begin - 3, r_end + 3): eight digits + a 0/1 "frame exists" factor that
    // multiplies the tap's bias part, so edge handling needs no branch (absent frames have zero codes)
    __shared__ __align__(16) float s_code[kFrontRows + 6][12];
    const int r_begin = blockIdx.x * kFrontRows;
    const int r_end = min(r_begin + kFrontRows, rows);
    for (int i = threadIdx.x; i < kFrontRows + 6; i += kFrontThreads) {
        const int r = r_begin - 3 + i;
        int valid = 0;
        unsigned uid = 0;
        if (r >= 0 && r < rows) {
            const int tok = row_tok ? row_tok[r] : r;
            if (tok >= 0) {
                const long long id = static_cast<long long>(ids[tok]);
                if ((id < 0 || id > 65535) && blockIdx.y == 0) atomicExch(err_flag, 1);
                uid = static_cast<unsigned>(id) & 0xFFFFu;
                valid = 1;
            }
        }
#pragma unroll
        for (int d = 0; d < 8; ++d)
            s_code[i][d] = valid ? static_cast<float>(static_cast<int>((uid >> (2 * d)) & 3) - 2) * 0.5f : 0.f;
        s_code[i][8] = valid ? 1.f : 0.f;
        s_code[i][9] = s_code[i][10] = s_code[i][11] = 0.f;
    }
    __syncthreads();
    // Sliding 7-row window in registers: code row n of the slab lives in slot n % 7, so every row is
    // read from shared memory once (3 loads) instead of once per tap (21).
    float cw[7][9];
    auto load_slot = [&](int slot, int n) {
        const float4 lo = *reinterpret_cast<const float4*>(&s_code[n][0]);
        const float4 hi = *reinterpret_cast<const float4*>(&s_code[n][4]);
        cw[slot][0] = lo.x; cw[slot][1] = lo.y; cw[slot][2] = lo.z; cw[slot][3] = lo.w;
        cw[slot][4] = hi.x; cw[slot][5] = hi.y; cw[slot][6] = hi.z; cw[slot][7] = hi.w;
        cw[slot][8] = s_code[n][8];
    };
#pragma unroll
    for (int n = 0; n < 6; ++n) load_slot(n, n);
    const int n_rows = r_end - r_begin;
    for (int rr = 0; rr < n_rows; rr += 7) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            load_slot((j + 6) % 7, rr + j + 6);  // row r + 3 of output row r = r_begin + rr + j
            // five independent chains (digit pairs + bias), tap-major inside each: fixed order
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, ab = 0.f;
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                const float* cv = cw[(j + t) % 7];
                a0 = fmaf(cv[1], m[t][1], fmaf(cv[0], m[t][0], a0));
                a1 = fmaf(cv[3], m[t][3], fmaf(cv[2], m[t][2], a1));
                a2 = fmaf(cv[5], m[t][5], fmaf(cv[4], m[t][4], a2));
                a3 = fmaf(cv[7], m[t][7], fmaf(cv[6], m[t][6], a3));
                ab = fmaf(cv[8], cb[t], ab);
            }
            const float y = (be + ab) + ((a0 + a1) + (a2 + a3));
            if (rr + j < n_rows)  // halo rows: zeros
                x[static_cast<size_t>(r_begin + rr + j) * C + c] = cw[(j + 3) % 7][8] != 0.f ? y : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------
// The same folded front end as a GEMM: ids -> A [rows, 128] (operand dtype), the im2col of the seven
// neighbouring codes. Column 8 * tap + d = digit d of frame t + tap - 3 (0 when that frame does not
// exist), 56 + tap = "frame exists", 63 = 1; columns 64..127 repeat 0..63. Every value is 0, +-0.5
// or +-1, exact in bf16 and fp16, so x = A B^T with B = [hi(coef) | lo(coef)] (the folded
// coefficients split into two 16-bit parts, launch_fsq_frontend_gemm's caller) reproduces the fp32
// fold to ~2^-17 on the tensor cores, and the GroupNorm statistics of x ride the GEMM epilogue.
// One warp per row: lane l writes columns 4l .. 4l+3 (8 bytes), 256 bytes per row.
template <typename OutT, typename IdT>
__global__ void __launch_bounds__(256)
fsq_im2col_kernel(const IdT* __restrict__ ids, const int32_t* __restrict__ row_tok, int rows,
                  OutT* __restrict__ a, int* __restrict__ err_flag) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int col = (lane & 15) * 4;            // 0..60; lanes 16-31 write the copy at +64
    const int tap = col < 56 ? col >> 3 : 3;    // flag columns are handled below
    auto frame_id = [&](int rr, bool& valid) -> unsigned {
        valid = false;
        if (rr < 0 || rr >= rows) return 0u;
        const int tok = row_tok ? row_tok[rr] : rr;
        if (tok < 0) return 0u;
        const long long id = static_cast<long long>(ids[tok]);
        if ((id < 0 || id > 65535) && lane == 0) atomicExch(err_flag, 1);
        valid = true;
        return static_cast<unsigned>(id) & 0xFFFFu;
    };
    bool centre;
    (void)frame_id(r, centre);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (centre) {  // halo rows stay all-zero: their x row is zero
        if (col < 56) {
            bool ok;
            const unsigned uid = frame_id(r + tap - 3, ok);
            const int d0 = col & 7;  // 0 or 4
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = ok ? static_cast<float>(static_cast<int>((uid >> (2 * (d0 + i))) & 3) - 2) * 0.5f : 0.f;
        } else {
            // columns 56..63: "frame exists" of taps 0..6, then the constant 1
            const int t0 = col - 56;  // 0 or 4
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int t = t0 + i;
                bool ok = true;
                if (t < 7) (void)frame_id(r + t - 3, ok);
                v[i] = ok ? 1.f : 0.f;
            }
        }
    }
    uint2 w;
    w.x = Half16<OutT>::pack(v[0], v[1]);
    w.y = Half16<OutT>::pack(v[2], v[3]);
    *reinterpret_cast<uint2*>(a + static_cast<size_t>(r) * 128 + lane * 4) = w;
}

// ---------------------------------------------------------------------------
// FSQ quantise (encode direction): the mirror image of the lookup above.
//
// Replaces vector_quantize_pytorch.ResidualFSQ.forward as called by Encoder.quantize
// (tts/core/codec/encoder.py:73-78; one quantizer, levels [4]*8, dim 2048):
//   z      = project_in(x)                       Linear 2048 -> 8 (+ bias)
//   b      = tanh(z + shift) * half_l - offset   FSQ.bound: half_l = 3 * 1.001 / 2, offset = 0.5,
//                                                shift = atanh(offset / half_l)
//   digit  = rint(b) + 2  in {0, 1, 2, 3}        round_ste; half_width = 2
//   id     = sum_d digit_d * 4^d                 codes_to_indices, basis = cumprod([1, 4, ...])
// `pre_bound`: releases of the library differ in whether ResidualFSQ bounds the projected input
// once more before its first layer (residual = layers[0].bound(x)); 1 applies it.
//
// HBM-bound: 8 KB of fp32 features in, one id out per token. A warp reduces four tokens at a
// time against the 64 KB projection (L1-resident); the summation order is fixed (lane-serial
// over 16 float4 slabs, then an xor butterfly), so a token's id does not depend on the batch.
constexpr int kFsqQTokensPerWarp = 4;

__device__ __forceinline__ float fsq_bound(float z, float half_l, float shift) {
    return tanhf(z + shift) * half_l - 0.5f;
}

template <typename IdT>
__global__ void __launch_bounds__(kFsqThreads)
fsq_quantize_kernel(const float* __restrict__ x, int ld, int n_tokens, const float* __restrict__ w_in,
                    const float* __restrict__ b_in, int dim, float half_l, float shift, int pre_bound,
                    IdT* __restrict__ ids, float* __restrict__ z_out) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * kFsqThreads + threadIdx.x) >> 5;
    const int tok0 = warp_global * kFsqQTokensPerWarp;
    if (tok0 >= n_tokens) return;
    float acc[kFsqQTokensPerWarp][8];
#pragma unroll
    for (int t = 0; t < kFsqQTokensPerWarp; ++t)
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[t][d] = 0.f;
    for (int c = lane * 4; c < dim; c += 128) {
        float4 w[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) w[d] = __ldg(reinterpret_cast<const float4*>(w_in + static_cast<size_t>(d) * dim + c));
#pragma unroll
        for (int t = 0; t < kFsqQTokensPerWarp; ++t) {
            const int tok = min(tok0 + t, n_tokens - 1);
            const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(tok) * ld + c);
#pragma unroll
            for (int d = 0; d < 8; ++d)
                acc[t][d] = fmaf(v.w, w[d].w, fmaf(v.z, w[d].z, fmaf(v.y, w[d].y, fmaf(v.x, w[d].x, acc[t][d]))));
        }
    }
#pragma unroll
    for (int t = 0; t < kFsqQTokensPerWarp; ++t)
#pragma unroll
        for (int d = 0; d < 8; ++d)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[t][d] += __shfl_xor_sync(0xffffffffu, acc[t][d], o);
    // lane l finishes (token l / 8, dimension l % 8)
    const int t_sel = lane >> 3, d_sel = lane & 7;
    float z = 0.f;
#pragma unroll
    for (int t = 0; t < kFsqQTokensPerWarp; ++t)
#pragma unroll
        for (int d = 0; d < 8; ++d)
            if (t == t_sel && d == d_sel) z = acc[t][d];
    z += __ldg(b_in + d_sel);
    const int tok = tok0 + t_sel;
    if (z_out != nullptr && tok < n_tokens) z_out[static_cast<size_t>(tok) * 8 + d_sel] = z;
    float b = z;
    if (pre_bound) b = fsq_bound(b, half_l, shift);
    b = fsq_bound(b, half_l, shift);
    int digit = static_cast<int>(rintf(b)) + 2;  // torch.round: half to even
    digit = min(max(digit, 0), 3);
    unsigned id = static_cast<unsigned>(digit) << (2 * d_sel);
    id |= __shfl_xor_sync(0xffffffffu, id, 1);
    id |= __shfl_xor_sync(0xffffffffu, id, 2);
    id |= __shfl_xor_sync(0xffffffffu, id, 4);
    if (d_sel == 0 && tok < n_tokens) ids[tok] = static_cast<IdT>(id);
}

template <typename OutT>
int launch_typed(const void* ids, int id_type, const int32_t* row_tok, int rows,
                 const float* w_out, const float* b_out, int channels, void* out, int ld,
                 int* err_flag, cudaStream_t stream) {
    if (rows <= 0) return 0;
    // one thread per 8 channels: 256 threads for the 2048-wide codebook output, 128 for the folded
    // 1024-wide one; wider outputs tile over grid.y
    const int threads = channels / kFsqChanPerThread < kFsqThreads ? channels / kFsqChanPerThread : kFsqThreads;
    const int chan_blocks = channels / (threads * kFsqChanPerThread);
    dim3 grid((rows + kFsqRowsPerBlock - 1) / kFsqRowsPerBlock, chan_blocks);
    if (id_type == 1)
        B200_CUDA_OK(launch_kernel(fsq_lookup_kernel<OutT, long long>, grid, dim3(threads), 0, stream,
                                   static_cast<const long long*>(ids), row_tok, rows, w_out, b_out, channels,
                                   static_cast<OutT*>(out), ld, err_flag));
    else
        B200_CUDA_OK(launch_kernel(fsq_lookup_kernel<OutT, int>, grid, dim3(threads), 0, stream,
                                   static_cast<const int*>(ids), row_tok, rows, w_out, b_out, channels,
                                   static_cast<OutT*>(out), ld, err_flag));
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_fsq_lookup(const void* ids, int id_type, const int32_t* row_tok, int rows,
                      const float* w_out, const float* b_out, int channels, void* out, int ld,
                      int out_prec, int* err_flag, cudaStream_t stream) {
    B200_CHECK(channels >= 256 && (channels % (kFsqChanPerThread * kFsqThreads) == 0 ||
                                   (channels < kFsqChanPerThread * kFsqThreads && channels % 256 == 0)),
               "fsq: channels (%d) must be 256..2048 in steps of 256, or a multiple of %d", channels,
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

int launch_fsq_frontend(const void* ids, int id_type, const int32_t* row_tok, int rows, const float* m_fold,
                        const float* cb_fold, const float* b_embed, int C, float* x, int* err_flag,
                        cudaStream_t stream) {
    if (rows <= 0) return 0;
    B200_CHECK(C % kFrontThreads == 0, "fsq front end: channels must be a multiple of %d", kFrontThreads);
    dim3 grid((rows + kFrontRows - 1) / kFrontRows, C / kFrontThreads);
    if (id_type == 1)
        B200_CUDA_OK(launch_kernel(fsq_frontend_kernel<long long>, grid, dim3(kFrontThreads), 0, stream,
                                   static_cast<const long long*>(ids), row_tok, rows, m_fold, cb_fold, b_embed, C, x,
                                   err_flag));
    else
        B200_CUDA_OK(launch_kernel(fsq_frontend_kernel<int>, grid, dim3(kFrontThreads), 0, stream,
                                   static_cast<const int*>(ids), row_tok, rows, m_fold, cb_fold, b_embed, C, x,
                                   err_flag));
    return 0;
}

int launch_fsq_im2col(const void* ids, int id_type, const int32_t* row_tok, int rows, int prec, void* a,
                      int* err_flag, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const dim3 grid((rows + 7) / 8), block(256);
#define B200_IM2COL(OUT, ID)                                                                               \
    B200_CUDA_OK(launch_kernel(fsq_im2col_kernel<OUT, ID>, grid, block, 0, stream, static_cast<const ID*>(ids), \
                               row_tok, rows, static_cast<OUT*>(a), err_flag))
    if (prec == kPrecBf16) {
        if (id_type == 1) B200_IM2COL(__nv_bfloat16, long long);
        else B200_IM2COL(__nv_bfloat16, int);
    } else if (prec == kPrecFp16) {
        if (id_type == 1) B200_IM2COL(__half, long long);
        else B200_IM2COL(__half, int);
    } else {
        set_error("fsq im2col: unsupported precision %d", prec);
        return 1;
    }
#undef B200_IM2COL
    return 0;
}

int launch_fsq_quantize(const float* x, int ld, int n_tokens, const float* w_in, const float* b_in,
                        int dim, int pre_bound, void* ids, int id_type, float* z_out, cudaStream_t stream) {
    if (n_tokens <= 0) return 0;
    B200_CHECK(dim % 128 == 0 && ld >= dim && ld % 4 == 0, "fsq_quantize: bad dim %d / ld %d", dim, ld);
    B200_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0, "fsq_quantize: features must be 16-byte aligned");
    // the constants of FSQ.bound for levels = 4, evaluated in fp32 as the library does
    const float half_l = 3.0f * 1.001f / 2.0f;
    const float shift = atanhf(0.5f / half_l);
    const int tokens_per_block = (kFsqThreads / 32) * kFsqQTokensPerWarp;
    const int grid = (n_tokens + tokens_per_block - 1) / tokens_per_block;
    if (id_type == 1)
        fsq_quantize_kernel<long long><<<grid, kFsqThreads, 0, stream>>>(
            x, ld, n_tokens, w_in, b_in, dim, half_l, shift, pre_bound, static_cast<long long*>(ids), z_out);
    else
        fsq_quantize_kernel<int><<<grid, kFsqThreads, 0, stream>>>(x, ld, n_tokens, w_in, b_in, dim, half_l, shift,
                                                                  pre_bound, static_cast<int*>(ids), z_out);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LLM token ids -> FSQ code ids. The reference detokenises a completion to a string, tokenises
// it again and parses "<|s_N|>" (rewards.py:70-73, inferencing.py:53-63). The speech tokens are added to the
// tokenizer in SORTED (lexicographic) order (tokenization.py:36-49), so vocabulary id -> N is a permutation, not an
// offset: `table[v]` holds N for a speech token and -1 for anything else. One CTA per sequence keeps the speech
// tokens in order (ballot + prefix sums) and writes them packed at the sequence's own offset.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
map_speech_tokens_kernel(const int32_t* __restrict__ table, int vocab, const long long* __restrict__ tok,
                         const int32_t* __restrict__ seq_off, int32_t* __restrict__ codes, int32_t* __restrict__ out_len) {
    __shared__ int warp_cnt[8];
    __shared__ int base;
    const int s = blockIdx.x;
    const int lo = seq_off[s], hi = seq_off[s + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = lo; i0 < hi; i0 += 256) {
        const int i = i0 + threadIdx.x;
        int code = -1;
        if (i < hi) {
            const long long v = tok[i];
            if (v >= 0 && v < vocab) code = table[v];
        }
        const unsigned m = __ballot_sync(0xffffffffu, code >= 0);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        if (code >= 0) codes[lo + before + __popc(m & ((1u << lane) - 1u))] = code;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
            base += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_len[s] = base;
}
}  // namespace

int launch_map_speech_tokens(const int32_t* table, int vocab, const long long* tok, const int32_t* seq_off, int n_seq,
                             int32_t* codes, int32_t* out_len, cudaStream_t stream) {
    if (n_seq <= 0) return 0;
    map_speech_tokens_kernel<<<n_seq, 256, 0, stream>>>(table, vocab, tok, seq_off, codes, out_len);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
