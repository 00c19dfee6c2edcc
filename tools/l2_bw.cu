// L2 / HBM store and load throughput of plain 128-bit accesses (development microbenchmark).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/l2_bw.cu -o build/l2_bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void store_kernel(float4* p, size_t n4, int reps) {
    const float4 v = make_float4(1.f, 2.f, 3.f, static_cast<float>(threadIdx.x));
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
             i += static_cast<size_t>(gridDim.x) * blockDim.x)
            p[i] = v;
}
__global__ void load_kernel(const float4* p, size_t n4, int reps, float* sink) {
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
             i += static_cast<size_t>(gridDim.x) * blockDim.x) {
            const float4 v = p[i];
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}
__global__ void rmw_kernel(float4* p, size_t n4, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
             i += static_cast<size_t>(gridDim.x) * blockDim.x) {
            float4 v = p[i];
            v.x += 1.f;
            p[i] = v;
        }
}
int main() {
    float* sink;
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (size_t mb : {16, 32, 64, 512}) {
        float4* buf;
        const size_t bytes = mb << 20;
        cudaMalloc(&buf, bytes);
        cudaMemset(buf, 0, bytes);
        const int reps = mb >= 512 ? 4 : 40;
        for (int threads : {256, 1024}) {
            const int grid = 148 * (2048 / threads);
            for (int mode = 0; mode < 3; ++mode) {
                float best = 1e9f;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    if (mode == 0) store_kernel<<<grid, threads>>>(buf, bytes / 16, reps);
                    if (mode == 1) load_kernel<<<grid, threads>>>(buf, bytes / 16, reps, sink);
                    if (mode == 2) rmw_kernel<<<grid, threads>>>(buf, bytes / 16, reps);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double gb = static_cast<double>(bytes) * reps * (mode == 2 ? 2 : 1) / 1e9;
                printf("%4zu MiB  threads %4d  %-5s %8.1f GB/s\n", mb, threads,
                       mode == 0 ? "store" : mode == 1 ? "load" : "rmw", gb / (best * 1e-3));
            }
        }
        cudaFree(buf);
    }
    return 0;
}
