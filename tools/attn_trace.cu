// In-kernel timeline of the tcgen05 attention kernel (development tool, not part of the product).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200_ATTN_TRACE --expt-relaxed-constexpr \
//        -I tts_max_b200/csrc -I include tools/attn_trace.cu -lcuda -o build/attn_trace
//   build/attn_trace [n_utts T heads]
//
// Prints when (us after the first CTA started) thread 0 of a few CTAs reached each phase.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_tc05.cu"       // make_tmap_2d
#include "attention_tc05.cu"

namespace b200 {
int g_use_pdl = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}
}  // namespace b200

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint32_t x = static_cast<uint32_t>(i) * 2654435761u + seed;
        x ^= x >> 15;
        x *= 2246822519u;
        x ^= x >> 13;
        p[i] = __float2bfloat16((static_cast<float>(x & 0xffff) / 65536.f - 0.5f) * 2.f);
    }
}

int main(int argc, char** argv) {
    const int n_utts = argc > 1 ? atoi(argv[1]) : 16;
    const int T = argc > 2 ? atoi(argv[2]) : 500;
    const int heads = argc > 3 ? atoi(argv[3]) : 16;
    const int D = heads * 64;
    const int pitch = T + 3;
    const int rows = n_utts * pitch - 3;
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, static_cast<size_t>(rows) * 3 * D * 2);
    cudaMalloc(&out, static_cast<size_t>(rows) * D * 2);
    fill_kernel<<<1024, 256>>>(qkv, static_cast<size_t>(rows) * 3 * D, 7);
    std::vector<int4> work;
    for (int u = 0; u < n_utts; ++u)
        for (int q0 = 0; q0 < T; q0 += 128)  // {q_row, n_q, kv_row0, T_kv}
            work.push_back(make_int4(u * pitch + q0, T - q0 < 128 ? T - q0 : 128, u * pitch, T));
    int4* work_dev;
    cudaMalloc(&work_dev, work.size() * sizeof(int4));
    cudaMemcpy(work_dev, work.data(), work.size() * sizeof(int4), cudaMemcpyHostToDevice);
    b200::RowSpace rs;
    rs.rows = rows;
    rs.n_utts = n_utts;
    rs.attn128_work = work_dev;
    rs.n_attn128_work = static_cast<int>(work.size());
    const size_t n_cta = work.size() * heads;
    unsigned long long* trace;
    cudaMalloc(&trace, n_cta * 64 * 8);
    cudaMemset(trace, 0, n_cta * 64 * 8);

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) b200::launch_attention_tc05(b200::kPrecBf16, qkv, rs, heads, out, nullptr);
    float best = 1e9f;
    for (int it = 0; it < 10; ++it) {
        cudaEventRecord(e0);
        b200::launch_attention_tc05(b200::kPrecBf16, qkv, rs, heads, out, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    printf("%d utterances x %d tokens, %d heads: %zu CTAs, untraced best %.2f us\n", n_utts, T, heads, n_cta,
           best * 1e3f);
    cudaMemcpyToSymbol(b200::g_attn_trace, &trace, sizeof(trace));
    b200::launch_attention_tc05(b200::kPrecBf16, qkv, rs, heads, out, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    std::vector<unsigned long long> h(n_cta * 64);
    cudaMemcpy(h.data(), trace, n_cta * 64 * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t1 = 0;
    for (size_t b = 0; b < n_cta; ++b) {
        t0 = std::min(t0, h[b * 64]);
        t1 = std::max(t1, h[b * 64 + 5]);
    }
    printf("first CTA start -> last CTA exit: %.2f us\n", (t1 - t0) * 1e-3);
    // CTA lifetime statistics
    std::vector<double> life;
    for (size_t b = 0; b < n_cta; ++b) life.push_back((h[b * 64 + 5] - h[b * 64]) * 1e-3);
    std::sort(life.begin(), life.end());
    printf("CTA lifetime us: min %.2f  median %.2f  p90 %.2f  max %.2f; sum / 296 slots = %.2f us\n", life.front(),
           life[life.size() / 2], life[life.size() * 9 / 10], life.back(),
           [&] { double s = 0; for (double v : life) s += v; return s / 296.0; }());
    static const char* ph[7] = {"S ready", "S in regs", "max xchg", "PV_j-1 done", "own P stored", "all P stored", "MMAs issued"};
    for (size_t b : {static_cast<size_t>(0), n_cta / 2, n_cta - 1}) {
        const unsigned long long* e = &h[b * 64];
        const unsigned long long s0 = e[0];
        printf("CTA %zu: start %.2f us | prologue +%.2f, operands landed +%.2f, last PV done +%.2f, stored +%.2f, exit +%.2f\n",
               b, (s0 - t0) * 1e-3, (e[1] - s0) * 1e-3, (e[2] - s0) * 1e-3, (e[3] - s0) * 1e-3, (e[4] - s0) * 1e-3,
               (e[5] - s0) * 1e-3);
        for (int j = 0; j < 6; ++j) {
            if (e[8 + j * 8] == 0) continue;
            printf("   tile %d:", j);
            for (int k = 0; k < 7; ++k) printf("  %s +%.2f", ph[k], e[8 + j * 8 + k] ? (e[8 + j * 8 + k] - s0) * 1e-3 : -1.0);
            printf("\n");
        }
    }
    return 0;
}
