// K13 + K14: complex spectrum + inverse STFT ("same" padding), register-resident FFT, one warp per PAIR
// of frames.
//
// Replaces (reference):
//   ISTFTHead.forward after self.out   tts/core/codec/decoder_modules.py:131-148
//       mag = clip(exp(m), max=100); S = mag * (cos p + i sin p)
//   ISTFT.forward, padding == "same"   tts/core/codec/decoder_modules.py:59-93
//       irfft(S, n_fft, norm="backward") * window -> overlap-add (fold) -> trim (win-hop)/2
//       -> divide by the overlap-added squared window
//
// Input is the head Linear output, token-major [rows, ld] fp32: columns [0, n_bins) are the
// log-magnitudes and [n_bins, 2*n_bins) the phases of one frame (that is what
// transpose(1,2).chunk(2, dim=1) selects). n_fft = N = 4 * hop, n_bins = N/2 + 1.
//
// Why this shape. The first version (a shared-memory Stockham FFT, radix 4-4-4-10, four frames per
// 256-thread group) ran at 0.08 of the HBM roofline: ncu showed no wasted DRAM traffic but 5.6 M
// shared-memory bank conflicts, short-scoreboard / MIO / barrier stalls and seven CTA-wide barriers per
// four frames (profiles/r02a_hbm_kernels_ncu_full_summary.md). Here
//   * two REAL frames a, b ride one COMPLEX inverse FFT of length N: Y = X~a + i X~b (X~ = the Hermitian
//     extension irfft implies) gives z = N (x_a + i x_b), so there is no even/odd packing pass and the
//     mirrored bins are formed locally from the bins a lane already needs;
//   * N = N1 * 8 * 8 (N1 = 20 for n_fft 1280, 10 for 640) is done in THREE register-resident steps per
//     warp -- a prime-factor DFT-N1 (4x5 or 2x5, no inner twiddles) and two radix-8 steps -- with two
//     transposes through a private, padded (bank-conflict-free) scratch of 11.5 KB per warp; only
//     __syncwarp inside the transform, twiddles from two small tables laid out like their consumers;
//   * one CTA = kWarps warps = 2 * kWarps frames = 2 * kWarps - 4 output hops: a single __syncthreads,
//     then every thread overlap-adds (ascending frame order: deterministic) and normalises the samples it
//     owns straight out of the warps' scratch and stores them coalesced.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// packed fp32 pairs (FADD2 / FFMA2 on sm_100): one instruction per complex add / scale
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cfma(float s, float2 a, float2 c) { return __ffma2_rn(make_float2(s, s), a, c); }
__device__ __forceinline__ float2 cmuli(float2 a) { return make_float2(-a.y, a.x); }  // i * a

// ---- small inverse DFTs (sign +), fully unrolled, register resident ----
__device__ __forceinline__ void idft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    // out[q] = sum_r x[r] (+i)^(q r)
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmuli(csub(x1, x3));
    x0 = cadd(t0, t2);
    x1 = cadd(t1, t3);
    x2 = csub(t0, t2);
    x3 = csub(t1, t3);
}
__device__ __forceinline__ void idft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
    constexpr float c1 = 0.30901699437494745f;   // cos(2 pi / 5)
    constexpr float c2 = -0.8090169943749475f;   // cos(4 pi / 5)
    constexpr float s1 = 0.9510565162951535f;    // sin(2 pi / 5)
    constexpr float s2 = 0.5877852522924731f;    // sin(4 pi / 5)
    const float2 a1 = cadd(x1, x4), a2 = cadd(x2, x3), b1 = csub(x1, x4), b2 = csub(x2, x3);
    const float2 e1 = cfma(c2, a2, cfma(c1, a1, x0));
    const float2 e2 = cfma(c1, a2, cfma(c2, a1, x0));
    const float2 d1 = cmuli(cfma(s2, b2, __fmul2_rn(make_float2(s1, s1), b1)));
    const float2 d2 = cmuli(cfma(-s1, b2, __fmul2_rn(make_float2(s2, s2), b1)));
    x0 = cadd(x0, cadd(a1, a2));
    x1 = cadd(e1, d1);
    x4 = csub(e1, d1);
    x2 = cadd(e2, d2);
    x3 = csub(e2, d2);
}
// in place, natural order in and out
__device__ __forceinline__ void idft8(float2 (&v)[8]) {
    constexpr float r = 0.70710678118654752f;
    idft4(v[0], v[2], v[4], v[6]);  // E[k] in v[0], v[2], v[4], v[6]
    idft4(v[1], v[3], v[5], v[7]);  // O[k] in v[1], v[3], v[5], v[7]
    const float2 o0 = v[1];
    const float2 o1 = make_float2((v[3].x - v[3].y) * r, (v[3].x + v[3].y) * r);    // * (1 + i) / sqrt 2
    const float2 o2 = cmuli(v[5]);                                                  // * i
    const float2 o3 = make_float2((-v[7].x - v[7].y) * r, (v[7].x - v[7].y) * r);   // * (-1 + i) / sqrt 2
    const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
__device__ __forceinline__ void idft3(float2& x0, float2& x1, float2& x2) {
    // out[q] = sum_r x[r] w^(q r), w = exp(+2 pi i / 3) = -1/2 + i sqrt(3)/2
    constexpr float h = 0.8660254037844386f;
    const float2 sum = cadd(x1, x2);
    const float2 d = cmuli(__fmul2_rn(make_float2(h, h), csub(x1, x2)));
    const float2 t = cfma(-0.5f, sum, x0);
    x0 = cadd(x0, sum);
    x1 = cadd(t, d);
    x2 = csub(t, d);
}
// Prime-factor inverse DFT of length N1 = NA * 5 (NA = 4, 3, 2 or 1; gcd(NA, 5) = 1, so no twiddles):
// input index n = (5 a + NA b) mod N1, output index k = (KA ka + KB kb) mod N1 with KA = 5 (5^-1 mod NA),
// KB = NA (NA^-1 mod 5) (the CRT map).
template <int N1>
__device__ __forceinline__ void idft_pfa(float2 (&v)[N1]) {
    static_assert(N1 == 20 || N1 == 15 || N1 == 10 || N1 == 5, "DFT-20 (4 x 5), DFT-15 (3 x 5), DFT-10 (2 x 5), DFT-5");
    constexpr int NA = N1 / 5;
    constexpr int KA = NA == 3 ? 10 : 5;
    constexpr int KB = NA == 4 ? 16 : NA == 3 ? 6 : NA == 2 ? 6 : 1;
    float2 y[NA][5];
#pragma unroll
    for (int a = 0; a < NA; ++a) {
#pragma unroll
        for (int b = 0; b < 5; ++b) y[a][b] = v[(5 * a + NA * b) % N1];
        idft5(y[a][0], y[a][1], y[a][2], y[a][3], y[a][4]);
    }
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) {
        if constexpr (NA == 4) {
            idft4(y[0][kb], y[1][kb], y[2][kb], y[3][kb]);
        } else if constexpr (NA == 3) {
            idft3(y[0][kb], y[1][kb], y[2][kb]);
        } else if constexpr (NA == 2) {
            const float2 s = cadd(y[0][kb], y[1][kb]), d = csub(y[0][kb], y[1][kb]);
            y[0][kb] = s;
            y[1][kb] = d;
        }
#pragma unroll
        for (int ka = 0; ka < NA; ++ka) v[(KA * ka + KB * kb) % N1] = y[ka][kb];
    }
}

constexpr int kT1Stride = 72;  // float2 words per k1 row of the first transpose (64 + 8: see header)
constexpr int kT2Stride = 9;   // float2 words per (k1, k2) row of the second transpose (8 + 1)

template <int HOP>
struct IstftCfg {
    static constexpr int kN = 4 * HOP;        // n_fft = complex FFT length (two real frames per transform)
    static constexpr int kBins = 2 * HOP + 1;
    static constexpr int kN1 = kN / 64;       // 20, 15, 10 or 5 (hop 320, 240, 160, 80)
    static_assert(kN % 64 == 0 && kN1 % 5 == 0, "n_fft = 4 hop must be 64 x (a multiple of 5)");
    static constexpr int kCombos = 8 * kN1;   // (k1, n3) / (k1, k2) combinations of the radix-8 steps
    static constexpr int kRounds = (kCombos + 31) / 32;
    static constexpr int kScratch = kN1 * kT1Stride;  // float2 words per warp: T1 == T2 size >= 2 * kBins, >= kN
    static constexpr int kTw1 = 8 * kN1;              // tw1[n2][k1]
    static constexpr int kTw2 = kCombos * kT2Stride;  // tw2[(k1 + N1 k2)][n3], padded like T2
    static_assert(kScratch >= 2 * kBins && kScratch >= kN && kScratch == kCombos * kT2Stride, "scratch layout");
};

// kWarps warps per CTA; frames [b0 - 2, b0 + 2 kWarps - 2) -> output hops [b0, b0 + 2 kWarps - 4)
template <int HOP, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 512 / (kWarps * 32))
istft_kernel(const float* __restrict__ x_pred, int ld, const int4* __restrict__ work,
             const int32_t* __restrict__ utt_row0, const int32_t* __restrict__ utt_len,
             const int32_t* __restrict__ utt_tok0, const float2* __restrict__ tw_tables,
             const float* __restrict__ window, float* __restrict__ wav) {
    using C = IstftCfg<HOP>;
    constexpr int kN = C::kN, kBins = C::kBins, kN1 = C::kN1;
    constexpr int kHops = 2 * kWarps - 4;
    constexpr int kPad = (kN - HOP) / 2;  // "same" trim on each side
    constexpr int kThreads = kWarps * 32;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float2* scratch_all = reinterpret_cast<float2*>(smem_raw);          // [kWarps][kScratch]
    float2* tw1 = scratch_all + kWarps * C::kScratch;                   // [8][kN1]
    float2* tw2 = tw1 + C::kTw1;                                        // [kCombos][9]
    float* win = reinterpret_cast<float*>(tw2 + C::kTw2);               // [kN]
    float* env_hop = win + kN;                                          // [HOP] interior envelope

    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // tables do not depend on the predecessor kernel: fill them before waiting for it
    for (int i = threadIdx.x; i < C::kTw1 + C::kTw2; i += kThreads) tw1[i] = tw_tables[i];
    for (int i = threadIdx.x; i < kN; i += kThreads) win[i] = window[i];
    pdl_wait();
    __syncthreads();  // the twiddle tables are read by every warp long before the barrier that ends the transforms

    const int4 wk = work[blockIdx.x];
    const int utt = wk.x, b0 = wk.y;
    const int T = utt_len[utt];
    const int row0 = utt_row0[utt];
    float2* sc = scratch_all + warp * C::kScratch;

    const int ta = b0 - 2 + 2 * warp, tb = ta + 1;  // this warp's two frames
    const bool va = ta >= 0 && ta < T, vb = tb >= 0 && tb < T;
    if (va || vb) {
        // ---- P0: spectra X_a, X_b -> scratch [0, kBins) and [kBins, 2 kBins); Im of bins 0 and N/2 dropped ----
        // all 4 x kIt loads of both frames are in flight before the first MUFU instruction
        constexpr int kIt = (kBins + 31) / 32;
        float lm[2][kIt], ph[2][kIt];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            const bool ok = f == 0 ? va : vb;
            const float* row = x_pred + static_cast<size_t>(row0 + (ok ? (f == 0 ? ta : tb) : 0)) * ld;
#pragma unroll
            for (int i = 0; i < kIt; ++i) {
                const int k = lane + 32 * i;
                const bool in = ok && k < kBins;
                lm[f][i] = in ? __ldg(row + k) : -INFINITY;   // exp(-inf) = 0: an absent frame has a zero spectrum
                ph[f][i] = in ? __ldg(row + kBins + k) : 0.f;
            }
        }
#pragma unroll
        for (int f = 0; f < 2; ++f) {
#pragma unroll
            for (int i = 0; i < kIt; ++i) {
                const int k = lane + 32 * i;
                if (k >= kBins) continue;
                // exp via MUFU.EX2, sin/cos via MUFU after a two-term Cody-Waite reduction to
                // [-pi, pi] (abs error ~5e-7, three orders below the GEMM operand rounding)
                const float mag = fminf(__expf(lm[f][i]), 100.f);
                const float p0 = ph[f][i];
                const float kq = rintf(p0 * 0.15915494309189535f);
                float rr = fmaf(kq, -6.2831854820251465f, p0);   // 2*pi (fp32 high part)
                rr = fmaf(kq, 1.7484556000744487e-07f, rr);      // 2*pi low part: 2*pi = hi - 1.748e-7
                float sn, cs;
                __sincosf(rr, &sn, &cs);
                sc[f * kBins + k] = make_float2(mag * cs, (k == 0 || k == kBins - 1) ? 0.f : mag * sn);
            }
        }
        __syncwarp();

        // ---- P1: Y[n] = X~a[n] + i X~b[n] for n = 64 n1 + c, both of this lane's offsets c; DFT-N1 over n1 ----
        // offsets: lo c = lane; hi c = 64 - lane (lane 0: 32) -- the mirror N - n of a lane's bins falls on its
        // own other offset, and every shared-memory access below is to consecutive words across the warp
        const int c_lo = lane, c_hi = lane == 0 ? 32 : 64 - lane;
        float2 y[2][kN1];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int c = o == 0 ? c_lo : c_hi;
#pragma unroll
            for (int n1 = 0; n1 < kN1; ++n1) {
                const int n = 64 * n1 + c;
                // N / 2 = 32 N1: for odd N1 it falls inside block n1 = N1 / 2, whose low offsets (c < 32) are bins
                // and whose high offsets (c >= 32) are mirrors
                if (n1 < kN1 / 2 || (kN1 % 2 == 1 && n1 == kN1 / 2 && o == 0)) {  // n < N/2: the bin itself
                    const float2 xa = sc[n], xb = sc[kBins + n];
                    y[o][n1] = make_float2(xa.x - xb.y, xa.y + xb.x);
                } else {             // n >= N/2: conj of bin N - n
                    const float2 xa = sc[kN - n], xb = sc[kBins + kN - n];
                    y[o][n1] = make_float2(xa.x + xb.y, xb.x - xa.y);
                }
            }
        }
        __syncwarp();  // every lane has read its spectra: the scratch becomes T1
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int c = o == 0 ? c_lo : c_hi;
            idft_pfa<kN1>(y[o]);
            const float2* t1 = tw1 + (c >> 3) * kN1;  // W_N^(8 n2 k1), n2 = c / 8
#pragma unroll
            for (int k1 = 0; k1 < kN1; ++k1) sc[k1 * kT1Stride + c] = k1 == 0 ? y[o][0] : cmul(y[o][k1], t1[k1]);
        }
        __syncwarp();

        // ---- P2: radix-8 over n2 for (k1, n3) = divmod(q, 8), q = lane + 32 j; twiddle W_N^(n3 (k1 + N1 k2)) ----
        float2 v[C::kRounds][8];
#pragma unroll
        for (int j = 0; j < C::kRounds; ++j) {
            const int q = lane + 32 * j;
            if (C::kCombos % 32 != 0 && q >= C::kCombos) continue;
            const int k1 = q >> 3, n3 = q & 7;
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) v[j][n2] = sc[k1 * kT1Stride + 8 * n2 + n3];
        }
        __syncwarp();  // T1 consumed: the scratch becomes T2
#pragma unroll
        for (int j = 0; j < C::kRounds; ++j) {
            const int q = lane + 32 * j;
            if (C::kCombos % 32 != 0 && q >= C::kCombos) continue;
            const int k1 = q >> 3, n3 = q & 7;
            idft8(v[j]);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int m = k1 + kN1 * k2;
                sc[m * kT2Stride + n3] = cmul(v[j][k2], tw2[m * kT2Stride + n3]);
            }
        }
        __syncwarp();

        // ---- P3: radix-8 over n3 for p = k1 + N1 k2 = lane + 32 j; output sample t = p + 8 N1 k3 ----
#pragma unroll
        for (int j = 0; j < C::kRounds; ++j) {
            const int p = lane + 32 * j;
            if (C::kCombos % 32 != 0 && p >= C::kCombos) continue;
#pragma unroll
            for (int n3 = 0; n3 < 8; ++n3) v[j][n3] = sc[p * kT2Stride + n3];
        }
        __syncwarp();  // T2 consumed: the scratch becomes the two time-domain frames [2][kN] floats
        float* fa = reinterpret_cast<float*>(sc);
        float* fb = fa + kN;
#pragma unroll
        for (int j = 0; j < C::kRounds; ++j) {
            const int p = lane + 32 * j;
            if (C::kCombos % 32 != 0 && p >= C::kCombos) continue;
            idft8(v[j]);
#pragma unroll
            for (int k3 = 0; k3 < 8; ++k3) {
                fa[p + C::kCombos * k3] = v[j][k3].x;  // z = N (x_a + i x_b); 1 / N rides the window
                fb[p + C::kCombos * k3] = v[j][k3].y;
            }
        }
    }
    // interior envelope: a hop whose five frames b-2 .. b+2 all exist sees the same sum of squared windows
    for (int j = threadIdx.x; j < HOP; j += kThreads) {
        float env = 0.f;
#pragma unroll
        for (int dt = -2; dt <= 2; ++dt) {
            const int m = j + kPad - HOP * dt;
            if (m >= 0 && m < kN) env = fmaf(win[m], win[m], env);
        }
        env_hop[j] = env;
    }
    __syncthreads();

    // ---- overlap-add (ascending frame order), envelope normalisation, coalesced store ----
    // frame slot f (0 .. 2 kWarps) = frame b0 - 2 + f covers tile samples [HOP (f - 2) - pad, + N)
    float* wav_u = wav + static_cast<size_t>(utt_tok0[utt]) * HOP;
    const int n_out = min(kHops, T - b0) * HOP;
    const float* frames = reinterpret_cast<const float*>(scratch_all);
    for (int n = threadIdx.x; n < n_out; n += kThreads) {
        const int hb = n / HOP;
        const int b = b0 + hb;
        float acc = 0.f, env = 0.f;
        const bool interior = b >= 2 && b + 2 < T;
#pragma unroll
        for (int dt = -2; dt <= 2; ++dt) {
            const int f = hb + 2 + dt;           // slot of frame b + dt
            const int t = b + dt;
            const int m = n - (HOP * (f - 2) - kPad);
            if (t >= 0 && t < T && m >= 0 && m < kN) {
                const float w = win[m];
                acc = fmaf(frames[(f >> 1) * (2 * C::kScratch) + (f & 1) * kN + m], w, acc);
                env = fmaf(w, w, env);
            }
        }
        if (interior) env = env_hop[n - hb * HOP];
        // z = N x (norm="backward" irfft divides by N): sum(x w) / sum(w^2) = acc / (env N)
        wav_u[static_cast<size_t>(b0) * HOP + n] = acc / (env * kN);
    }
}

template <int HOP, int kWarps>
int launch_istft_typed(const float* x_pred, int ld, const RowSpace& rs, const IstftTables& tab, float* wav,
                       cudaStream_t stream) {
    using C = IstftCfg<HOP>;
    B200_CHECK(ld >= 2 * C::kBins, "istft: ld %d < %d", ld, 2 * C::kBins);
    constexpr size_t smem = sizeof(float2) * (static_cast<size_t>(kWarps) * C::kScratch + C::kTw1 + C::kTw2) +
                            sizeof(float) * (C::kN + HOP);
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(istft_kernel<HOP, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
        B200_CUDA_OK(cudaFuncSetAttribute(istft_kernel<HOP, kWarps>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
    }
    B200_CUDA_OK(launch_kernel(istft_kernel<HOP, kWarps>, dim3(rs.n_istft_work), dim3(kWarps * 32), smem, stream,
                               x_pred, ld, rs.istft_work, rs.utt_row0, rs.utt_len, rs.utt_tok0,
                               tab.twiddle, tab.window, wav));
    return 0;
}

}  // namespace

// twiddle tables in the kernel's shared-memory layout: tw1[n2][k1] = W_N^(8 n2 k1), then
// tw2[(k1 + N1 k2) * 9 + n3] = W_N^(n3 (k1 + N1 k2)) with W_N = exp(+2 pi i / N); computed in double
int g_istft_hops = 0;  // 0: chosen per batch (codec.cu plan_layout), 12 / 28: forced (A/B)

int istft_table_words(int hop) {
    const int n1 = 4 * hop / 64;
    return 8 * n1 + 8 * n1 * kT2Stride;
}
void istft_fill_tables(int hop, float2* host) {
    const int N = 4 * hop, n1 = N / 64;
    const double w = 2.0 * 3.14159265358979323846 / N;
    for (int n2 = 0; n2 < 8; ++n2)
        for (int k1 = 0; k1 < n1; ++k1) {
            const double a = w * ((8 * n2 * k1) % N);
            host[n2 * n1 + k1] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
        }
    float2* t2 = host + 8 * n1;
    for (int m = 0; m < 8 * n1; ++m)
        for (int n3 = 0; n3 < kT2Stride; ++n3) {
            const double a = w * ((n3 * m) % N);
            t2[m * kT2Stride + n3] = n3 < 8 ? make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)))
                                            : make_float2(0.f, 0.f);
        }
}

int launch_istft(const float* x_pred, int ld, const RowSpace& rs, const IstftTables& tab, int hop,
                 float* wav, cudaStream_t stream) {
    if (rs.n_istft_work <= 0) return 0;
    B200_CHECK(rs.istft_hops == 12 || rs.istft_hops == 28, "istft: tile of %d hops is not instantiated", rs.istft_hops);
    if (hop == 320) {
        if (rs.istft_hops == 12) return launch_istft_typed<320, 8>(x_pred, ld, rs, tab, wav, stream);
        return launch_istft_typed<320, 16>(x_pred, ld, rs, tab, wav, stream);
    }
    if (hop == 160) {
        if (rs.istft_hops == 12) return launch_istft_typed<160, 8>(x_pred, ld, rs, tab, wav, stream);
        return launch_istft_typed<160, 16>(x_pred, ld, rs, tab, wav, stream);
    }
    if (hop == 240) {
        if (rs.istft_hops == 12) return launch_istft_typed<240, 8>(x_pred, ld, rs, tab, wav, stream);
        return launch_istft_typed<240, 16>(x_pred, ld, rs, tab, wav, stream);
    }
    if (hop == 80) {
        if (rs.istft_hops == 12) return launch_istft_typed<80, 8>(x_pred, ld, rs, tab, wav, stream);
        return launch_istft_typed<80, 16>(x_pred, ld, rs, tab, wav, stream);
    }
    set_error("istft: hop_length %d is not instantiated (320, 240, 160, 80: n_fft = 4 hop = 64 x {20, 15, 10, 5})", hop);
    return 1;
}

}  // namespace b200
