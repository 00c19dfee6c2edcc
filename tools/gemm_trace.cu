// In-kernel timeline of the CTA-pair GEMM (development tool, not part of the product library).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200_GEMM_TRACE \
//        --expt-relaxed-constexpr -I tts_max_b200/csrc -I include tools/gemm_trace.cu -lcuda \
//        -o build/gemm_trace
//   build/gemm_trace [M N K taps residual(0/1) out_fp32(0/1) flush(0/1) in_place(0/1) out16(0/1)]
//
// Prints, for the slowest / median CTA pair, when (ns after the first CTA started) each phase of
// the kernel was reached: prologue done, operands of tile i landed, last MMA of tile i issued,
// accumulator i ready, tile i drained, exit.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_tc05.cu"

namespace b200 {
int g_use_pdl = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}
extern unsigned long long* g_gemm_trace;
}  // namespace b200

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint32_t x = static_cast<uint32_t>(i) * 2654435761u + seed;
        x ^= x >> 15;
        x *= 2246822519u;
        x ^= x >> 13;
        p[i] = __float2bfloat16((static_cast<float>(x & 0xffff) / 65536.f - 0.5f) * 0.1f);
    }
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 8045;
    const int N = argc > 2 ? atoi(argv[2]) : 1024;
    const int K = argc > 3 ? atoi(argv[3]) : 1024;
    const int taps = argc > 4 ? atoi(argv[4]) : 1;
    const int with_res = argc > 5 ? atoi(argv[5]) : 1;
    const int out_fp32 = argc > 6 ? atoi(argv[6]) : 1;
    const int do_flush = argc > 7 ? atoi(argv[7]) : 1;   // 0: operands stay L2-resident between launches
    const int in_place = argc > 8 ? atoi(argv[8]) : 0;   // 1: residual == out (x += ..., as the decoder does)
    const int with_16 = argc > 9 ? atoi(argv[9]) : 1;    // 0: no 16-bit copy / sum of squares
    __nv_bfloat16 *a, *w, *o16;
    float *out, *res, *ss;
    unsigned long long* trace;
    const int halo = 8;
    cudaMalloc(&a, static_cast<size_t>(M + 2 * halo) * K * 2);
    cudaMalloc(&w, static_cast<size_t>(N) * K * taps * 2);
    cudaMalloc(&out, static_cast<size_t>(M) * N * 4);
    cudaMalloc(&res, static_cast<size_t>(M) * N * 4);
    cudaMalloc(&o16, static_cast<size_t>(M) * N * 2);
    cudaMalloc(&ss, static_cast<size_t>(M) * b200::kGemmSsSlots * 4);
    const int grid = 148;
    cudaMalloc(&trace, grid * 128 * 8);
    fill_kernel<<<1024, 256>>>(a, static_cast<size_t>(M + 2 * halo) * K, 1);
    fill_kernel<<<1024, 256>>>(w, static_cast<size_t>(N) * K * taps, 2);
    cudaMemset(res, 0, static_cast<size_t>(M) * N * 4);
    char* flush;
    cudaMalloc(&flush, 256u << 20);

    b200::GemmCall c{};
    c.precision = b200::kPrecBf16;
    c.a = a + static_cast<size_t>(halo) * K;
    c.a_rows = M;
    c.Cin = K;
    c.w = w;
    c.N = N;
    c.taps = taps;
    c.out = out;
    c.out_fp32 = out_fp32;
    c.ldc = N;
    c.n_store = N;
    c.bias = nullptr;
    c.residual = with_res && out_fp32 ? (in_place ? out : res) : nullptr;
    c.ld_res = N;
    c.row_valid = nullptr;
    c.act = b200::kActNone;
    if (with_16 && with_res && out_fp32 && N == 1024) {
        c.out16 = o16;
        c.ld16 = N;
        c.ss_out = ss;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) b200::launch_gemm(c, nullptr);
    float best = 1e9f;
    for (int it = 0; it < 10; ++it) {
        if (do_flush) cudaMemsetAsync(flush, it, 256u << 20);
        cudaEventRecord(e0);
        b200::launch_gemm(c, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    printf("M=%d N=%d K=%d taps=%d res=%d fp32=%d flush=%d inplace=%d out16=%d: untraced best %.2f us\n", M, N, K, taps,
           with_res, out_fp32, do_flush, in_place, with_16, best * 1e3f);
    cudaMemset(trace, 0, grid * 128 * 8);
    b200::g_gemm_trace = trace;
    if (do_flush) cudaMemsetAsync(flush, 1, 256u << 20);
    cudaEventRecord(e0);
    b200::launch_gemm(c, nullptr);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("traced launch: %.2f us\n", ms * 1e3f);
    std::vector<unsigned long long> h(grid * 128);
    cudaMemcpy(h.data(), trace, grid * 128 * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t_end = 0;
    for (int b = 0; b < grid; ++b) {
        if (h[b * 128 + 0]) t0 = std::min(t0, h[b * 128 + 0]);
        t_end = std::max(t_end, h[b * 128 + 20]);
    }
    int missing = 0;
    for (int b = 0; b < grid; ++b) missing += h[b * 128 + 20] == 0;
    if (missing) printf("%d CTAs left no exit stamp\n", missing);
    printf("first CTA start -> last CTA exit: %.2f us\n", (t_end - t0) * 1e-3);
    static const char* names[32] = {"start", "prologue", "ops0", "mma0", "ops1", "mma1", "ops2", "mma2",
                                    "ops3", "mma3", "acc0", "drain0", "acc1", "drain1", "acc2", "drain2",
                                    "acc3", "drain3", "", "", "exit", "bar_init", "alloc_in", "alloc_out"};
    // order leader CTAs by exit time; print the fastest, the median and the slowest pair
    std::vector<int> leaders;
    for (int b = 0; b < grid; b += 2) leaders.push_back(b);
    std::sort(leaders.begin(), leaders.end(), [&](int x, int y) { return h[x * 128 + 20] < h[y * 128 + 20]; });
    const int picks[1] = {leaders.back()};
    for (int b : picks) {
        printf("CTA %3d (leader) / %3d (peer):\n", b, b + 1);
        for (int s = 0; s <= 23; ++s) {
            if (names[s][0] == 0) continue;
            const unsigned long long tl = h[b * 128 + s], tp = h[(b + 1) * 128 + s];
            if (tl == 0 && tp == 0) continue;
            printf("   %-9s leader %8.2f us   peer %8.2f us\n", names[s], tl ? (tl - t0) * 1e-3 : -1.0,
                   tp ? (tp - t0) * 1e-3 : -1.0);
        }
        // per-chunk phases of epilogue warp 4 (leader CTA), tiles 0 and 1: ns since the chunk began
        for (int t = 0; t < 2; ++t) {
            if (h[b * 128 + 32 + t * 48] == 0) continue;
            printf("   tile %d chunks (start | +boxes free, +acc in regs, +staged, +stores issued):\n", t);
            for (int c = 0; c < 4; ++c) {
                const unsigned long long* e = &h[b * 128 + 32 + t * 48 + c * 6];
                if (e[0] == 0) continue;
                printf("     c%d %8.2f us |", c, (e[0] - t0) * 1e-3);
                for (int k = 1; k < 5; ++k) printf(" %5lld", e[k] ? (long long)(e[k] - e[0]) : -1ll);
                printf("\n");
            }
        }
    }
    return 0;
}
