// tf32_probe.cu -- throughput of legacy mma.sync TF32 (m16n8k8) and BF16 (m16n8k16) on sm_100a, for the 3xTF32 Float32 plan.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tf32_kernel(float* out, int iters) {
    float c[8][4];
    for (int q = 0; q < 8; q++) for (int e = 0; e < 4; e++) c[q][e] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x * 5, b1 = 11;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[q][0]), "+f"(c[q][1]), "+f"(c[q][2]), "+f"(c[q][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int q = 0; q < 8; q++) for (int e = 0; e < 4; e++) s += c[q][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    for (int nt : {256, 512, 1024}) {
        const int iters = 4096, blocks = 148 * (1024 / nt);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        tf32_kernel<<<blocks, nt>>>(out, iters);
        cudaEventRecord(e0);
        tf32_kernel<<<blocks, nt>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * (double)blocks * (nt / 32);
        printf("mma.sync m16n8k8 tf32, %4d threads/CTA, %d CTAs: %.1f TFLOP/s dense (3xTF32 effective: %.1f)\n", nt, blocks, flops / ms * 1e-9,
               flops / ms * 1e-9 / 3);
    }
    return 0;
}
