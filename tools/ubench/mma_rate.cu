// Micro-benchmark: legacy warp-level mma.sync m16n8k8 tf32 issue rate on sm_100a (per SM), vs FFMA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mma_kernel(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_kernel(float* out, int iters) {
    float c[32];
    for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 0.001f + i;
    float a = 1.0001f, b = 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < 32; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        mma_kernel<<<148, warps * 32>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        mma_kernel<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double macs = 148.0 * warps * iters * 8.0 * (16 * 8 * 8);
        printf("mma.sync tf32 m16n8k8: warps/SM=%d  %.1f TMAC/s  -> %.0f MAC/clk/SM @1.965GHz (dense tf32 TFLOP/s %.0f)\n", warps,
               macs / ms / 1e9, macs / (ms * 1e-3) / 148 / 1.965e9, 2 * macs / ms / 1e9);
        ffma_kernel<<<148, warps * 32>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        ffma_kernel<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fm = 148.0 * warps * 32 * iters * 32.0;
        printf("ffma: warps/SM=%d %.0f FMA/clk/SM (%.1f TFLOP/s)\n", warps, fm / (ms * 1e-3) / 148 / 1.965e9, 2 * fm / ms / 1e9);
    }
    return 0;
}
