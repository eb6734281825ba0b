// Microbenchmark: FP64 tensor (DMMA m8n8k4) rate on B200, alone and interleaved with vector DFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DMMA only, 1: DFMA only, 2: both interleaved (8 DMMA + 16 DFMA per trip)
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b) {
    double c[8][2], x[16];
#pragma unroll
    for (int q = 0; q < 8; q++) { c[q][0] = threadIdx.x; c[q][1] = q; }
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = threadIdx.x + q;
    for (int i = 0; i < iters; i++) {
        if (MODE != 1) {
#pragma unroll
            for (int q = 0; q < 8; q++) dmma(c[q][0], c[q][1], a, b);
        }
        if (MODE != 0) {
#pragma unroll
            for (int q = 0; q < 16; q++) x[q] = fma(x[q], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += c[q][0] + c[q][1];
#pragma unroll
    for (int q = 0; q < 16; q++) s += x[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter) {
    int sms = 148, blocks = sms * 8, iters = 1 << 14;
    double* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    double fmas = (double)blocks * 256 * iters * fma_per_thread_iter;
    printf("%-28s %8.3f ms  %.3e FMA/s  (%.1f per clk per SM at 1965 MHz)\n", name, best, fmas / (best * 1e-3),
           fmas / (best * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main() {
    run<0>("DMMA m8n8k4 only", 8 * 8.0);          // 8 DMMA x 256 FMA / 32 lanes
    run<1>("DFMA only", 16.0);
    run<2>("DMMA + DFMA interleaved", 8 * 8.0 + 16.0);
    return 0;
}
