// In-kernel timeline of a GEMM chain (c_proj -> fc1 -> fc2 -> next c_attn) and of the same four GEMMs as
// separate launches (development tool, not part of the product library).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200_GEMM_TRACE --expt-relaxed-constexpr \
//        -I tts_max_b200/csrc -I include tools/chain_trace.cu -lcuda -o build/chain_trace
//   build/chain_trace [M] [dbg]
//
// Prints per cluster the time of every tile (first operands landed, last MMA issued, accumulator seen by the
// epilogue, tile drained) relative to the first CTA's start, plus the busy fraction of the tensor pipe.
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
        p[i] = __float2bfloat16((static_cast<float>(x & 0xffff) / 65536.f - 0.5f) * 0.05f);
    }
}

int main(int argc, char** argv) {
    using namespace b200;
    const int M = argc > 1 ? atoi(argv[1]) : 8045;
    const int C = 1024;
    __nv_bfloat16 *y, *xb, *f, *qkv, *wproj, *wfc1, *wfc2, *wqkv;
    float *x, *ss;
    uint32_t* ctr;
    unsigned long long* trace;
    cudaMalloc(&y, (size_t)M * C * 2);
    cudaMalloc(&xb, (size_t)M * C * 2);
    cudaMalloc(&f, (size_t)M * 4 * C * 2);
    cudaMalloc(&qkv, (size_t)M * 3 * C * 2);
    cudaMalloc(&wproj, (size_t)C * C * 2);
    cudaMalloc(&wfc1, (size_t)4 * C * C * 2);
    cudaMalloc(&wfc2, (size_t)4 * C * C * 2);
    cudaMalloc(&wqkv, (size_t)3 * C * C * 2);
    cudaMalloc(&x, (size_t)M * C * 4);
    cudaMalloc(&ss, (size_t)M * kGemmSsSlots * 4);
    const int num_m = (M + 255) / 256;
    cudaMalloc(&ctr, 4 * num_m * 4);
    const int grid = 148;
    cudaMalloc(&trace, grid * 128 * 8);
    fill_kernel<<<1024, 256>>>(y, (size_t)M * C, 1);
    fill_kernel<<<1024, 256>>>(wproj, (size_t)C * C, 2);
    fill_kernel<<<1024, 256>>>(wfc1, (size_t)4 * C * C, 3);
    fill_kernel<<<1024, 256>>>(wfc2, (size_t)4 * C * C, 4);
    fill_kernel<<<1024, 256>>>(wqkv, (size_t)3 * C * C, 5);
    cudaMemset(x, 0, (size_t)M * C * 4);
    auto call = [&](const void* a, int Cin, const void* w, int N, void* out, bool fp32, const float* res, int act,
                    bool produce, bool consume) {
        GemmCall c{};
        c.precision = kPrecBf16;
        c.a = a; c.a_rows = M; c.Cin = Cin; c.w = w; c.N = N; c.taps = 1; c.out = out; c.out_fp32 = fp32;
        c.ldc = N; c.n_store = N; c.bias = nullptr; c.residual = res; c.ld_res = N; c.row_valid = nullptr; c.act = act;
        c.ss_inv_dim = 1.f / C; c.ss_eps = 1e-6f; c.out16_scale = 1.f; c.ss_in_scale = 1.f;
        if (produce) { c.out16 = xb; c.ld16 = C; c.ss_out = ss; }
        if (consume) c.ss_in = ss;
        return c;
    };
    GemmCall calls[4] = {call(y, C, wproj, C, x, true, x, kActNone, true, false),
                         call(xb, C, wfc1, 4 * C, f, false, nullptr, kActSilu, false, true),
                         call(f, 4 * C, wfc2, C, x, true, x, kActNone, true, false),
                         call(xb, C, wqkv, 3 * C, qkv, false, nullptr, kActNone, false, true)};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    char* flush;
    cudaMalloc(&flush, 256u << 20);
    float best_chain = 1e9f, best_sep = 1e9f;
    for (int it = 0; it < 6; ++it) {
        cudaMemsetAsync(flush, it, 256u << 20);
        cudaMemsetAsync(ctr, 0, 4 * num_m * 4);
        cudaEventRecord(e0);
        if (launch_gemm_chain(calls, 4, ctr, nullptr)) return 1;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 1) best_chain = std::min(best_chain, ms);
        cudaMemsetAsync(flush, it, 256u << 20);
        cudaEventRecord(e0);
        for (int g = 0; g < 4; ++g)
            if (launch_gemm(calls[g], nullptr)) return 1;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 1) best_sep = std::min(best_sep, ms);
    }
    printf("M=%d: chain %.1f us, four launches (no PDL) %.1f us\n", M, best_chain * 1e3f, best_sep * 1e3f);
    cudaMemset(trace, 0, grid * 128 * 8);
    g_gemm_trace = trace;
    cudaMemsetAsync(flush, 1, 256u << 20);
    cudaMemsetAsync(ctr, 0, 4 * num_m * 4);
    if (launch_gemm_chain(calls, 4, ctr, nullptr)) return 1;
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    std::vector<unsigned long long> h(grid * 128);
    cudaMemcpy(h.data(), trace, grid * 128 * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t_end = 0;
    for (int b = 0; b < grid; ++b) {
        if (h[b * 128 + 0]) t0 = std::min(t0, h[b * 128 + 0]);
        t_end = std::max(t_end, h[b * 128 + 20]);
    }
    printf("traced chain: first CTA start -> last exit %.2f us\n", (t_end - t0) * 1e-3);
    double busy_sum = 0;
    int shown = 0;
    for (int b = 0; b < grid; b += 2) {
        double busy = 0;
        int tiles = 0;
        for (int i = 0; i < 24; ++i) {
            const unsigned long long a = h[b * 128 + 32 + 4 * i], m = h[b * 128 + 32 + 4 * i + 1];
            if (a == 0) break;
            busy += (m - a) * 1e-3;
            ++tiles;
        }
        busy_sum += busy;
        if (b % 24 == 0 || b == grid - 2) {
            printf("cluster %2d: %2d tiles, mainloop busy %.1f us, exit %.1f us\n   ", b / 2, tiles, busy, (h[b * 128 + 20] - t0) * 1e-3);
            for (int i = 0; i < tiles; ++i) {
                const unsigned long long* e = &h[b * 128 + 32 + 4 * i];
                printf("[%d: ops %.1f mma %.1f acc %.1f drain %.1f] ", i, (e[0] - t0) * 1e-3, (e[1] - t0) * 1e-3, e[2] ? (e[2] - t0) * 1e-3 : -1.0,
                       e[3] ? (e[3] - t0) * 1e-3 : -1.0);
            }
            printf("\n");
            ++shown;
        }
    }
    printf("mean mainloop-busy per cluster %.1f us of %.1f us (%.0f %%)\n", busy_sum / (grid / 2), (t_end - t0) * 1e-3,
           100.0 * busy_sum / (grid / 2) / ((t_end - t0) * 1e-3));
    return 0;
}
