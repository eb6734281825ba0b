// Microbenchmark: what an FP64 instruction costs on a B200 SM sub-partition, by operand form and neighbourhood.
//   MODE 0: x = fma(x, c, c)        one register operand (constant-bank multiplier and addend)
//   MODE 1: x = fma(x, y, c)        two register operands
//   MODE 2: x = fma(x, y, z)        three distinct register operands
//   MODE 3: x = fma(y, z, x)        three, accumulator form
//   MODE 4: MODE 2 + one IMAD per DFMA (independent integer chain)
//   MODE 5: MODE 0 + one IMAD per DFMA
//   MODE 6: x = x + y (DADD), MODE 7: x = x * y (DMUL)
//   MODE 12: x = fma(x, x, y)  MODE 13: x = fma(y, y, x)  MODE 14: x = fma(x, y, x)   (the same register in two operand slots: the
//            Newton steps of the square root have this form)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_probe fp64_issue_probe.cu && ./fp64_issue_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NCH>
__global__ void __launch_bounds__(256, 2) probe(double* out, int iters, double ca, double cb, int ia) {
    double x[NCH], y[NCH], z[NCH];
    int w[NCH];
#pragma unroll
    for (int q = 0; q < NCH; q++) { x[q] = threadIdx.x + q; y[q] = 0.999999 + 1e-9 * (threadIdx.x + q); z[q] = 1e-7 * (q + 1) + 1e-12 * threadIdx.x; w[q] = threadIdx.x + q; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                if (MODE == 0 || MODE == 5) x[q] = fma(x[q], ca, cb);
                if (MODE == 1) x[q] = fma(x[q], y[q], cb);
                if (MODE == 2 || MODE == 4) x[q] = fma(x[q], y[q], z[q]);
                if (MODE == 3) x[q] = fma(y[q], z[q], x[q]);
                if (MODE == 6) x[q] = x[q] + y[q];
                if (MODE == 7) x[q] = x[q] * y[q];
                if (MODE == 8 || MODE == 11) x[q] = fma(x[q], y[q], cb);
                if (MODE == 10) x[q] = x[q] + y[q];
                if (MODE == 12) x[q] = fma(x[q], x[q], y[q]);
                if (MODE == 13) x[q] = fma(y[q], y[q], x[q]);
                if (MODE == 14) x[q] = fma(x[q], y[q], x[q]);
                if (MODE == 4 || MODE == 5 || MODE == 8 || MODE == 10 || MODE == 11) w[q] = w[q] * ia + 12345;
                if (MODE == 11) w[q] = w[q] * ia + 777;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < NCH; q++) s += x[q] + y[q] + z[q] + w[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NCH>
void run(const char* name, int ctas_per_sm) {
    int sms = 148, blocks = sms * ctas_per_sm, iters = 1 << 12;
    double* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        probe<MODE, NCH><<<blocks, 256>>>(out, iters, 0.999999, 1e-7, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    // warps per SMSP = ctas_per_sm * 8 / 4; FP64 instructions per warp = iters * 4 * NCH
    const double warp_instr = (double)iters * 4 * NCH * (ctas_per_sm * 2.0);   // per SMSP
    const double cycles = best * 1e-3 * 1.965e9;
    printf("%-44s NCH=%2d CTAs/SM=%d %8.3f ms  %.2f cycles per FP64 warp-instruction per SMSP  (%.1f lane-ops/clk/SM)\n", name, NCH, ctas_per_sm,
           best, cycles / warp_instr, 4 * 32.0 / (cycles / warp_instr));
    cudaFree(out);
}

int main() {
    run<0, 16>("DFMA x = fma(x, c, c)", 2);
    run<1, 16>("DFMA x = fma(x, y, c)", 2);
    run<2, 16>("DFMA x = fma(x, y, z)", 2);
    run<3, 16>("DFMA x = fma(y, z, x)", 2);
    run<2, 8>("DFMA x = fma(x, y, z)", 2);
    run<2, 16>("DFMA x = fma(x, y, z)", 1);
    run<6, 16>("DADD x = x + y", 2);
    run<7, 16>("DMUL x = x * y", 2);
    run<4, 8>("DFMA(x,y,z) + 1 IMAD each", 2);
    run<5, 8>("DFMA(x,c,c) + 1 IMAD each", 2);
    run<8, 8>("DFMA(x,y,c) + 1 IMAD each", 2);
    run<10, 8>("DADD(x,y) + 1 IMAD each", 2);
    run<11, 8>("DFMA(x,y,c) + 2 IMAD each", 2);
    run<12, 16>("DFMA x = fma(x, x, y)  (one register twice)", 2);
    run<13, 16>("DFMA x = fma(y, y, x)  (one register twice)", 2);
    run<14, 16>("DFMA x = fma(x, y, x)  (one register twice)", 2);
    run<1, 8>("DFMA x = fma(x, y, c)", 2);
    return 0;
}
