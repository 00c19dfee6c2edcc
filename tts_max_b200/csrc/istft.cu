// K13 + K14: complex spectrum + inverse STFT ("same" padding) as one shared-memory-staged kernel.
//
// Replaces (reference):
//   ISTFTHead.forward after self.out   tts/core/codec/decoder_modules.py:131-148
//       mag = clip(exp(m), max=100); S = mag * (cos p + i sin p)
//   ISTFT.forward, padding == "same"   tts/core/codec/decoder_modules.py:59-93
//       irfft(S, n_fft, norm="backward") * hann -> overlap-add (fold) -> trim (win-hop)/2
//       -> divide by the overlap-added squared window
//
// Input is the head Linear output, token-major [rows, ld] fp32: columns [0, n_bins) are the
// log-magnitudes and [n_bins, 2*n_bins) the phases of one frame (that is what
// transpose(1,2).chunk(2, dim=1) selects). n_fft = 4 * hop = 1280, n_bins = 641.
//
// One CTA produces kIstftOutHops * hop consecutive output samples of one utterance. Output hop b
// receives frames b-2 .. b+2, so the CTA transforms kIstftOutHops + 4 frames (in groups of 4)
// and overlap-adds them in shared memory; nothing but the finished samples goes back to HBM.
// The length-1280 real inverse FFT is a length-640 complex inverse FFT of the packed
// even/odd spectrum (Stockham autosort, radix 4-4-4-10) -- irfft ignores Im(S[0]) and
// Im(S[n_fft/2]), and so does the packing below.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int kGroup = 4;    // frames transformed together
constexpr int kIstftThreads = 256;
constexpr int kTileFrames = kIstftOutHops + 4;
static_assert(kTileFrames % kGroup == 0, "tile frames must be a multiple of the group");

// HOP = 320 (xcodec2, 16 kHz: n_fft 1280, complex FFT 640 = 4*4*4*10) or 160 (48 kHz upsampler
// variant: n_fft 640, complex FFT 320 = 4*4*4*5)
template <int HOP>
struct IstftSmem {
    static constexpr int kNfft = 4 * HOP;
    static constexpr int kHalf = 2 * HOP;
    float2 buf_a[kGroup][kHalf];
    float2 buf_b[kGroup][kHalf];
    float2 tw[kNfft];
    float win[kNfft];
    float ola[kIstftOutHops * HOP];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// one Stockham pass of radix R over `kGroup` independent length-kHalf transforms
template <int R, int kHalf>
__device__ __forceinline__ void stockham_pass(const float2 (*src)[kHalf], float2 (*dst)[kHalf],
                                              const float2* tw, int Ns) {
    constexpr int kNfft = 2 * kHalf;
    constexpr int kButterflies = kHalf / R;
    const int tw_stride = kNfft / (Ns * R);
    for (int idx = threadIdx.x; idx < kGroup * kButterflies; idx += kIstftThreads) {
        const int f = idx / kButterflies;
        const int j = idx - f * kButterflies;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = src[f][j + r * kButterflies];
            if (r > 0) v[r] = cmul(v[r], tw[r * k * tw_stride]);
        }
        const int j0 = (j / Ns) * Ns * R + k;
        if constexpr (R == 4) {
            // inverse DFT-4: out[q] = sum_r v[r] * (+i)^(q r)
            const float2 t0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
            const float2 t1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 t2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
            const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            const float2 t3 = make_float2(-d.y, d.x);  // i * d
            dst[f][j0 + 0 * Ns] = make_float2(t0.x + t2.x, t0.y + t2.y);
            dst[f][j0 + 1 * Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
            dst[f][j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
            dst[f][j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        } else if constexpr (R == 5) {
            // inverse DFT-5: X_k = sum_n x_n exp(+2 pi i n k / 5)
            constexpr float c1 = 0.30901699437494745f;   // cos(2 pi / 5)
            constexpr float c2 = -0.8090169943749475f;   // cos(4 pi / 5)
            constexpr float s1 = 0.9510565162951535f;    // sin(2 pi / 5)
            constexpr float s2 = 0.5877852522924731f;    // sin(4 pi / 5)
            const float2 x0 = v[0], x1 = v[1], x2 = v[2], x3 = v[3], x4 = v[4];
            const float2 a1 = make_float2(x1.x + x4.x, x1.y + x4.y);
            const float2 a2 = make_float2(x2.x + x3.x, x2.y + x3.y);
            const float2 b1 = make_float2(x1.x - x4.x, x1.y - x4.y);
            const float2 b2 = make_float2(x2.x - x3.x, x2.y - x3.y);
            const float2 e1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
            const float2 e2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
            const float2 t1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
            const float2 t2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
            const float2 d1 = make_float2(-t1.y, t1.x);  // i * t1
            const float2 d2 = make_float2(-t2.y, t2.x);
            dst[f][j0 + 0 * Ns] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
            dst[f][j0 + 1 * Ns] = make_float2(e1.x + d1.x, e1.y + d1.y);
            dst[f][j0 + 4 * Ns] = make_float2(e1.x - d1.x, e1.y - d1.y);
            dst[f][j0 + 2 * Ns] = make_float2(e2.x + d2.x, e2.y + d2.y);
            dst[f][j0 + 3 * Ns] = make_float2(e2.x - d2.x, e2.y - d2.y);
        } else {
            static_assert(R == 10, "only radix 4, 5 and 10 are instantiated");
            // inverse DFT-10 by the prime-factor map (no inner twiddles): n = (5 n1 + 2 n2) % 10,
            // k = (5 k1 + 6 k2) % 10; two DFT-5 over n2, then five DFT-2 over n1.
            constexpr float c1 = 0.30901699437494745f;   // cos(2 pi / 5)
            constexpr float c2 = -0.8090169943749475f;   // cos(4 pi / 5)
            constexpr float s1 = 0.9510565162951535f;    // sin(2 pi / 5)
            constexpr float s2 = 0.5877852522924731f;    // sin(4 pi / 5)
            float2 y[2][5];
#pragma unroll
            for (int n1 = 0; n1 < 2; ++n1) {
                const float2 x0 = v[(5 * n1 + 0) % 10], x1 = v[(5 * n1 + 2) % 10],
                             x2 = v[(5 * n1 + 4) % 10], x3 = v[(5 * n1 + 6) % 10],
                             x4 = v[(5 * n1 + 8) % 10];
                const float2 a1 = make_float2(x1.x + x4.x, x1.y + x4.y);
                const float2 a2 = make_float2(x2.x + x3.x, x2.y + x3.y);
                const float2 b1 = make_float2(x1.x - x4.x, x1.y - x4.y);
                const float2 b2 = make_float2(x2.x - x3.x, x2.y - x3.y);
                y[n1][0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
                const float2 e1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
                const float2 e2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
                // d = i * (s * b): (re, im) -> (-im, re)
                const float2 t1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
                const float2 t2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
                const float2 d1 = make_float2(-t1.y, t1.x);
                const float2 d2 = make_float2(-t2.y, t2.x);
                y[n1][1] = make_float2(e1.x + d1.x, e1.y + d1.y);
                y[n1][4] = make_float2(e1.x - d1.x, e1.y - d1.y);
                y[n1][2] = make_float2(e2.x + d2.x, e2.y + d2.y);
                y[n1][3] = make_float2(e2.x - d2.x, e2.y - d2.y);
            }
#pragma unroll
            for (int k2 = 0; k2 < 5; ++k2) {
                dst[f][j0 + ((6 * k2) % 10) * Ns] = make_float2(y[0][k2].x + y[1][k2].x, y[0][k2].y + y[1][k2].y);
                dst[f][j0 + ((5 + 6 * k2) % 10) * Ns] = make_float2(y[0][k2].x - y[1][k2].x, y[0][k2].y - y[1][k2].y);
            }
        }
    }
}

template <int HOP>
__global__ void __launch_bounds__(kIstftThreads)
istft_kernel(const float* __restrict__ x_pred, int ld, const int4* __restrict__ work,
             const int32_t* __restrict__ utt_row0, const int32_t* __restrict__ utt_len,
             const int32_t* __restrict__ utt_tok0, const float2* __restrict__ twiddle,
             const float* __restrict__ window, float* __restrict__ wav) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int kHop = HOP;
    constexpr int kNfft = 4 * HOP;
    constexpr int kHalf = 2 * HOP;
    constexpr int kBins = kHalf + 1;
    constexpr int kPad = (kNfft - kHop) / 2;  // "same" trim on each side
    extern __shared__ __align__(16) uint8_t smem_raw[];
    IstftSmem<HOP>& sm = *reinterpret_cast<IstftSmem<HOP>*>(smem_raw);

    const int4 wk = work[blockIdx.x];
    const int utt = wk.x, b0 = wk.y;
    const int T = utt_len[utt];
    const int row0 = utt_row0[utt];
    float* wav_u = wav + static_cast<size_t>(utt_tok0[utt]) * kHop;

    for (int i = threadIdx.x; i < kNfft; i += kIstftThreads) {
        sm.tw[i] = twiddle[i];
        sm.win[i] = window[i];
    }
    for (int i = threadIdx.x; i < kIstftOutHops * kHop; i += kIstftThreads) sm.ola[i] = 0.f;
    __syncthreads();

    // The head output of a frame group is fetched into registers one group ahead (kGroup * kKi
    // (log-magnitude, phase) pairs per thread), so the HBM/L2 latency hides behind the FFT passes
    // of the previous group instead of stalling the spectrum phase.
    constexpr int kKi = (kBins + kIstftThreads - 1) / kIstftThreads;
    float lm[kGroup][kKi], ph[kGroup][kKi];
    auto fetch_group = [&](int g) {
        const int t_first = b0 - 2 + g * kGroup;
#pragma unroll
        for (int f = 0; f < kGroup; ++f) {
            const int t = t_first + f;
            const bool t_ok = g < kTileFrames / kGroup && t >= 0 && t < T;
            const float* row = x_pred + static_cast<size_t>(row0 + (t_ok ? t : 0)) * ld;
#pragma unroll
            for (int i = 0; i < kKi; ++i) {
                const int k = threadIdx.x + i * kIstftThreads;
                const bool ok = t_ok && k < kBins;
                lm[f][i] = ok ? row[k] : 0.f;
                ph[f][i] = ok ? row[kBins + k] : 0.f;
            }
        }
    };
    fetch_group(0);

    for (int g = 0; g < kTileFrames / kGroup; ++g) {
        const int t_first = b0 - 2 + g * kGroup;  // absolute frame index of group slot 0
        if (t_first >= T || t_first + kGroup <= 0) {  // uniform: whole group outside
            fetch_group(g + 1);
            continue;
        }

        // 1. spectrum: X[k] = min(exp(m_k), 100) * (cos p_k, sin p_k); Re X[640] parked in X[0].y
#pragma unroll
        for (int f = 0; f < kGroup; ++f) {
            const int t = t_first + f;
            const bool t_ok = t >= 0 && t < T;
#pragma unroll
            for (int i = 0; i < kKi; ++i) {
                const int k = threadIdx.x + i * kIstftThreads;
                if (k >= kBins) continue;
                float2 X = make_float2(0.f, 0.f);
                if (t_ok) {
                    // exp via MUFU.EX2, sin/cos via MUFU after a two-term Cody-Waite reduction to
                    // [-pi, pi] (abs error ~5e-7, three orders below the GEMM operand rounding)
                    const float mag = fminf(__expf(lm[f][i]), 100.f);
                    const float p0 = ph[f][i];
                    const float kq = rintf(p0 * 0.15915494309189535f);
                    float rr = fmaf(kq, -6.2831854820251465f, p0);   // 2*pi (fp32 high part)
                    rr = fmaf(kq, 1.7484556000744487e-07f, rr);      // 2*pi low part: 2*pi = hi - 1.748e-7
                    float sn, cs;
                    __sincosf(rr, &sn, &cs);
                    X = make_float2(mag * cs, mag * sn);
                }
                if (k == 0) sm.buf_b[f][0].x = X.x;            // Im X[0] ignored by irfft
                else if (k == kHalf) sm.buf_b[f][0].y = X.x;   // Im X[640] ignored by irfft
                else sm.buf_b[f][k] = X;
            }
        }
        fetch_group(g + 1);  // lands during the pack / FFT / overlap-add phases below
        __syncthreads();

        // 2. pack: Z[k] = (X[k] + conj X[640-k]) + i * (X[k] - conj X[640-k]) * exp(+2 pi i k / 1280)
        for (int idx = threadIdx.x; idx < kGroup * kHalf; idx += kIstftThreads) {
            const int f = idx / kHalf;
            const int k = idx - f * kHalf;
            float2 xk, xr;
            if (k == 0) {
                xk = make_float2(sm.buf_b[f][0].x, 0.f);
                xr = make_float2(sm.buf_b[f][0].y, 0.f);
            } else {
                xk = sm.buf_b[f][k];
                const float2 m = sm.buf_b[f][kHalf - k];
                xr = make_float2(m.x, -m.y);
            }
            const float2 e = make_float2(xk.x + xr.x, xk.y + xr.y);
            const float2 o = cmul(make_float2(xk.x - xr.x, xk.y - xr.y), sm.tw[k]);
            sm.buf_a[f][k] = make_float2(e.x - o.y, e.y + o.x);
        }
        __syncthreads();

        // 3. length-640 inverse complex FFT, radix 4-4-4-10, ping-pong a -> b -> a -> b -> a
        stockham_pass<4, kHalf>(sm.buf_a, sm.buf_b, sm.tw, 1);
        __syncthreads();
        stockham_pass<4, kHalf>(sm.buf_b, sm.buf_a, sm.tw, 4);
        __syncthreads();
        stockham_pass<4, kHalf>(sm.buf_a, sm.buf_b, sm.tw, 16);
        __syncthreads();
        stockham_pass<kHalf / 64, kHalf>(sm.buf_b, sm.buf_a, sm.tw, 64);  // radix 10 (640) or 5 (320)
        __syncthreads();

        // 4. window, 1/n_fft and overlap-add. A thread owns the output samples n = tid (mod 256) in every
        //    frame (no write conflicts, no extra barrier) and walks, per frame, only the samples that
        //    frame covers; frames are added in ascending order, as before.
        //    buf_a[f] viewed as 1280 floats is the time-domain frame: x[2n] = Re z[n], x[2n+1] = Im z[n].
#pragma unroll
        for (int f = 0; f < kGroup; ++f) {
            const int t = t_first + f;
            if (t < 0 || t >= T) continue;  // uniform
            // frame t covers output samples [hop * t - pad, hop * t - pad + n_fft); o = that start
            // relative to this tile's first sample hop * b0
            const int o = kHop * (t - b0) - kPad;
            const int lo = max(o, 0), hi = min(o + kNfft, kIstftOutHops * kHop);
            const float* xf = reinterpret_cast<const float*>(sm.buf_a[f]);
            // first n >= lo with n = tid (mod 256)
            int n = lo + ((static_cast<int>(threadIdx.x) - lo) & (kIstftThreads - 1));
            for (; n < hi; n += kIstftThreads) {
                const int m = n - o;
                sm.ola[n] += xf[m] * (1.f / kNfft) * sm.win[m];
            }
        }
        __syncthreads();
    }

    // 5. normalise by the overlap-added squared window and store. A hop whose five frames b-2 .. b+2 all
    //    exist sees the same envelope (a function of n mod hop, summed in the same dt order); only the
    //    two hops at either end of an utterance take the general path.
    float* env_hop = reinterpret_cast<float*>(sm.buf_b);  // buf_b is free after the last group
    for (int j = threadIdx.x; j < kHop; j += kIstftThreads) {
        float env = 0.f;
#pragma unroll
        for (int dt = -2; dt <= 2; ++dt) {
            const int m = j + kPad - kHop * dt;
            if (m >= 0 && m < kNfft) env = fmaf(sm.win[m], sm.win[m], env);
        }
        env_hop[j] = env;
    }
    __syncthreads();
    const int n_out = min(kIstftOutHops, T - b0) * kHop;
    for (int n = threadIdx.x; n < n_out; n += kIstftThreads) {
        const int hb = n / kHop;
        const int b = b0 + hb;
        float env;
        if (b >= 2 && b + 2 < T) {
            env = env_hop[n - hb * kHop];
        } else {
            env = 0.f;
#pragma unroll
            for (int dt = -2; dt <= 2; ++dt) {
                const int t = b + dt;
                const int m = n + kPad - kHop * (t - b0);
                if (t >= 0 && t < T && m >= 0 && m < kNfft) env = fmaf(sm.win[m], sm.win[m], env);
            }
        }
        wav_u[static_cast<size_t>(b0) * kHop + n] = sm.ola[n] / env;
    }
}

template <int HOP>
int launch_istft_typed(const float* x_pred, int ld, const RowSpace& rs, const IstftTables& tab, float* wav,
                       cudaStream_t stream) {
    B200_CHECK(ld >= 2 * (2 * HOP + 1), "istft: ld %d < %d", ld, 2 * (2 * HOP + 1));
    static PerDeviceOnce once;
    if (once.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(istft_kernel<HOP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(sizeof(IstftSmem<HOP>))));
        B200_CUDA_OK(cudaFuncSetAttribute(istft_kernel<HOP>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
    }
    B200_CUDA_OK(launch_kernel(istft_kernel<HOP>, dim3(rs.n_istft_work), dim3(kIstftThreads),
                               sizeof(IstftSmem<HOP>), stream, x_pred, ld, rs.istft_work, rs.utt_row0, rs.utt_len,
                               rs.utt_tok0, tab.twiddle, tab.window, wav));
    return 0;
}

}  // namespace

int launch_istft(const float* x_pred, int ld, const RowSpace& rs, const IstftTables& tab, int hop,
                 float* wav, cudaStream_t stream) {
    if (rs.n_istft_work <= 0) return 0;
    if (hop == 320) return launch_istft_typed<320>(x_pred, ld, rs, tab, wav, stream);
    if (hop == 160) return launch_istft_typed<160>(x_pred, ld, rs, tab, wav, stream);
    set_error("istft: hop_length %d is not instantiated (320: n_fft 1280, 160: n_fft 640)", hop);
    return 1;
}

}  // namespace b200
