// Microbenchmark: throughput of scattered row updates to a [N, 200] fp32 matrix (1.9 GB, >> L2) with
//   (a) plain 128-bit stores, (b) red.global.add.f32 (32 lanes x 4 B), (c) red.global.add.v4.f32 (32 lanes x 16 B),
//   (d) 128-bit loads (gather) of the same rows.   nvcc -arch=sm_100a -O3 atomic_bw.cu -o atomic_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_rows(float* M, const int* rows, long nrows, int ld, int mode, float* sink) {
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (long i = w; i < nrows; i += nw) {
        const int r = __ldg(rows + i);
        float* p = M + (long)r * ld;
        if (mode == 0) { *reinterpret_cast<float4*>(p + 4 * lane) = make_float4(1, 2, 3, 4); }
        else if (mode == 1) { for (int c = 0; c < 4; ++c) atomicAdd(p + c * 32 + lane, 1.0f); }
        else if (mode == 2) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * lane), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory"); }
        else { float4 v = __ldg(reinterpret_cast<const float4*>(p + 4 * lane)); acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 12345.f) *sink = acc;
}
int main() {
    const long N = 2400000; const int ld = 200; const long nrows = 20000000;
    float* M; int* rows; float* sink;
    cudaMalloc(&M, N * ld * 4); cudaMemset(M, 0, N * ld * 4); cudaMalloc(&sink, 4);
    int* h = (int*)malloc(nrows * 4); unsigned s = 1; for (long i = 0; i < nrows; ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)((s >> 8) % N); }
    cudaMalloc(&rows, nrows * 4); cudaMemcpy(rows, h, nrows * 4, cudaMemcpyHostToDevice);
    const char* names[4] = {"store.v4", "red.f32 x4", "red.v4.f32", "load.v4"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k_rows<<<148 * 16, 256>>>(M, rows, nrows, ld, mode, sink); cudaDeviceSynchronize();
        cudaEventRecord(a); k_rows<<<148 * 16, 256>>>(M, rows, nrows, ld, mode, sink); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-12s %8.3f ms  %7.1f GB/s of 512-B row payload   %6.2f G lane-ops/s\n", names[mode], ms, nrows * 512.0 / ms / 1e6, nrows * (mode == 1 ? 128.0 : 32.0) / ms / 1e6);
    }
    return 0;
}
