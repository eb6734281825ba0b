// Microbenchmark: what packed FP32 (FFMA2 / FADD2 / FMUL2, sm_100 f32x2) and MUFU.EX2 cost on a B200 SM sub-partition, alone and mixed:
// the numbers the Float32 value kernels (gram_mvm_f32p.cuh, gram_mvm_tc5.cuh) are designed against.
//   MODE 0: FFMA   x = fma(x, y, z)                  scalar, three registers
//   MODE 1: FFMA2  x2 = fma(x2, y2, z2)              packed, three register pairs
//   MODE 2: FFMA2  x2 = fma(x2, y2, c2)              packed, accumulate form with a loop-invariant pair
//   MODE 3: FADD2  x2 = x2 + y2
//   MODE 4: MUFU.EX2 only
//   MODE 5: 1 MUFU.EX2 + K FFMA2 (independent), K = NK
//   MODE 6: 1 MUFU.EX2 + K FFMA (scalar)
//   MODE 7: FFMA2 + IMAD alternating (does the integer multiply-add share the FMA pipe with packed ops?)
//   MODE 8: FFMA2 + FMNMX alternating (alu pipe)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

template <int MODE, int NCH, int NK>
__global__ void __launch_bounds__(256, 2) probe(float* out, int iters, float ca, int ia) {
    float x[NCH], y[NCH], z[NCH], e[NCH];
    uint64_t x2[NCH], y2[NCH], z2[NCH];
    int w[NCH];
#pragma unroll
    for (int q = 0; q < NCH; q++) {
        x[q] = 1e-3f * (threadIdx.x + q); y[q] = 0.99999f + 1e-7f * (threadIdx.x + q); z[q] = 1e-7f * (q + 1); e[q] = -1e-3f * (threadIdx.x + q);
        x2[q] = pk(x[q], x[q] + 1.f); y2[q] = pk(y[q], y[q]); z2[q] = pk(z[q], z[q]); w[q] = threadIdx.x + q;
    }
    const uint64_t c2 = pk(ca, ca);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                if (MODE == 0) x[q] = fmaf(x[q], y[q], z[q]);
                if (MODE == 1) x2[q] = fma2(x2[q], y2[q], z2[q]);
                if (MODE == 2) x2[q] = fma2(y2[q], c2, x2[q]);
                if (MODE == 3) x2[q] = add2(x2[q], y2[q]);
                if (MODE == 4) e[q] = ex2(e[q]);
                if (MODE == 5) {
                    e[q] = ex2(e[q]);
#pragma unroll
                    for (int k = 0; k < NK; k++) x2[(q + k) % NCH] = fma2(x2[(q + k) % NCH], y2[q], z2[q]);
                }
                if (MODE == 6) {
                    e[q] = ex2(e[q]);
#pragma unroll
                    for (int k = 0; k < NK; k++) x[(q + k) % NCH] = fmaf(x[(q + k) % NCH], y[q], z[q]);
                }
                if (MODE == 7) { x2[q] = fma2(x2[q], y2[q], z2[q]); w[q] = w[q] * ia + 12345; }
                if (MODE == 8) { x2[q] = fma2(x2[q], y2[q], z2[q]); z[q] = fminf(z[q], x[q]); x[q] = fmaxf(z[q], y[q]); }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < NCH; q++) { float a, b; upk(x2[q], a, b); s += x[q] + y[q] + z[q] + e[q] + a + b + w[q]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NCH, int NK>
void run(const char* name, double instr_per_slot, double lane_ops_per_slot) {
    const int ctas_per_sm = 2, sms = 148, blocks = sms * ctas_per_sm, iters = 1 << 12;
    float* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        probe<MODE, NCH, NK><<<blocks, 256>>>(out, iters, 0.999f, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    // warps per SMSP = ctas_per_sm * 8 / 4 = 4; "slots" per warp = iters * 4 * NCH
    const double slots = (double)iters * 4 * NCH * (ctas_per_sm * 2.0);  // per SMSP
    const double cycles = best * 1e-3 * 1.965e9;
    printf("%-52s %8.3f ms  %6.2f cycles per slot per SMSP  = %5.2f cycles per warp instruction; %6.1f fp32 lane-ops/clk/SM\n", name, best,
           cycles / slots, cycles / slots / instr_per_slot, lane_ops_per_slot * 4 * 32.0 / (cycles / slots));
    cudaFree(out);
}

int main() {
    run<0, 16, 0>("FFMA  x = fma(x, y, z)", 1, 1);
    run<1, 16, 0>("FFMA2 x2 = fma(x2, y2, z2)", 1, 2);
    run<2, 16, 0>("FFMA2 x2 = fma(y2, c2, x2)", 1, 2);
    run<3, 16, 0>("FADD2 x2 = x2 + y2", 1, 2);
    run<4, 16, 0>("MUFU.EX2", 1, 1);
    run<5, 8, 1>("1 MUFU.EX2 + 1 FFMA2", 2, 2);
    run<5, 8, 2>("1 MUFU.EX2 + 2 FFMA2", 3, 4);
    run<5, 8, 3>("1 MUFU.EX2 + 3 FFMA2", 4, 6);
    run<5, 8, 4>("1 MUFU.EX2 + 4 FFMA2", 5, 8);
    run<5, 8, 6>("1 MUFU.EX2 + 6 FFMA2", 7, 12);
    run<6, 8, 4>("1 MUFU.EX2 + 4 FFMA", 5, 4);
    run<6, 8, 7>("1 MUFU.EX2 + 7 FFMA", 8, 7);
    run<7, 8, 0>("FFMA2 + IMAD", 2, 2);
    run<8, 8, 0>("FFMA2 + 2 FMNMX", 3, 2);
    return 0;
}
