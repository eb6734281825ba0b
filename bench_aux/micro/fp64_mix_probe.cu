// Microbenchmark 2: cost of integer / shared-memory instructions issued between FP64 instructions on a B200 SM sub-partition.
// FK: 0 DFMA(x,c,c)  1 DFMA(x,y,c)  2 DFMA(x,y,z)  3 DFMA(x,Ys,Zs) (Ys, Zs shared by all chains: operand reuse)
// IK: 0 none  1 LOP3 (1 reg + imm)  2 IMAD (1 reg + imm)  3 LDS.64 (per-lane address)  4 LOP3 with 2 regs  5 IADD3 3 regs
// NI: integer instructions per FP64 instruction (independent chains)
#include <cstdio>
#include <cuda_runtime.h>

template <int FK, int IK, int NI, int NCH>
__global__ void __launch_bounds__(256, 2) probe(double* out, int iters, double ca, double cb, int ia, const double* src) {
    __shared__ double sm[512];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 256] = 1.0;
    __syncthreads();
    double x[NCH], y[NCH], z[NCH], acc = 0;
    int w[NCH][NI > 0 ? NI : 1];
    const double Ys = src[threadIdx.x & 7], Zs = src[8 + (threadIdx.x & 7)];
#pragma unroll
    for (int q = 0; q < NCH; q++) {
        x[q] = threadIdx.x + q; y[q] = 0.999999 + 1e-9 * (threadIdx.x + q); z[q] = 1e-7 * (q + 1) + 1e-12 * threadIdx.x;
#pragma unroll
        for (int u = 0; u < (NI > 0 ? NI : 1); u++) w[q][u] = threadIdx.x * 8 + q + u;
    }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u4 = 0; u4 < 4; u4++) {
#pragma unroll
            for (int q = 0; q < NCH; q++) {
                if (FK == 0) x[q] = fma(x[q], ca, cb);
                if (FK == 1) x[q] = fma(x[q], y[q], cb);
                if (FK == 2) x[q] = fma(x[q], y[q], z[q]);
                if (FK == 3) x[q] = fma(x[q], Ys, Zs);
#pragma unroll
                for (int u = 0; u < NI; u++) {
                    if (IK == 1) asm volatile("lop3.b32 %0, %0, 0x7f8, 0x128, 0x6a;" : "+r"(w[q][u]));
                    if (IK == 2) asm volatile("mad.lo.s32 %0, %0, 5, 12345;" : "+r"(w[q][u]));
                    if (IK == 3) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(w[q][u] & 0xff8)); asm volatile("" ::"d"(v)); }
                    if (IK == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(w[q][u]) : "r"(ia), "r"(w[q][(u + 1) % (NI > 0 ? NI : 1)]));
                    if (IK == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w[q][u]) : "r"(ia), "r"(w[(q + 1) % NCH][u]));
                    if (IK == 6) asm volatile("mad.lo.s32 %0, %1, 32, %0;" : "+r"(w[q][u]) : "r"(__double2loint(x[q])));
                }
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int q = 0; q < NCH; q++) {
        s += x[q] + y[q] + z[q];
#pragma unroll
        for (int u = 0; u < (NI > 0 ? NI : 1); u++) s += w[q][u];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sm[(threadIdx.x * 7) & 511];
}

template <int FK, int IK, int NI, int NCH>
void run(const char* name) {
    int sms = 148, blocks = sms * 2, iters = 1 << 12;
    double *out, *src;
    cudaMalloc(&out, (size_t)blocks * 256 * 8);
    cudaMalloc(&src, 128);
    double h[16];
    for (int i = 0; i < 16; i++) h[i] = i < 8 ? 0.999999 + 1e-8 * i : 1e-7 * i;
    cudaMemcpy(src, h, 128, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        probe<FK, IK, NI, NCH><<<blocks, 256>>>(out, iters, 0.999999, 1e-7, 0xff0, src);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    const double warp_instr = (double)iters * 4 * NCH * 4.0;  // FP64 warp-instructions per SMSP (4 warps each)
    const double cycles = best * 1e-3 * 1.965e9;
    printf("%-52s %8.3f ms  %.2f cycles per (1 FP64 + %d other)\n", name, best, cycles / warp_instr, NI);
    cudaFree(out); cudaFree(src);
}

int main() {
    run<1, 0, 0, 8>("DFMA(x,y,c)");
    run<3, 0, 0, 8>("DFMA(x,Ys,Zs) shared operands (reuse)");
    run<1, 1, 1, 8>("DFMA(x,y,c) + 1 LOP3(r,imm)");
    run<1, 1, 2, 8>("DFMA(x,y,c) + 2 LOP3(r,imm)");
    run<1, 2, 1, 8>("DFMA(x,y,c) + 1 IMAD(r,imm,imm)");
    run<1, 2, 2, 8>("DFMA(x,y,c) + 2 IMAD(r,imm,imm)");
    run<1, 3, 1, 8>("DFMA(x,y,c) + 1 (LDS.64 + LOP)");
    run<0, 1, 1, 8>("DFMA(x,c,c) + 1 LOP3(r,imm)");
    run<0, 1, 2, 8>("DFMA(x,c,c) + 2 LOP3(r,imm)");
    run<0, 2, 2, 8>("DFMA(x,c,c) + 2 IMAD(r,imm,imm)");
    run<2, 1, 1, 8>("DFMA(x,y,z) + 1 LOP3(r,imm)");
    run<2, 0, 0, 8>("DFMA(x,y,z)");
    run<1, 4, 1, 8>("DFMA(x,y,c) + 1 LOP3(r,r,r)");
    run<1, 5, 1, 8>("DFMA(x,y,c) + 1 IMAD(r,r,r)");
    run<1, 3, 2, 8>("DFMA(x,y,c) + 2 LDS.64");
    run<1, 6, 1, 8>("DFMA(x,y,c) + 1 IMAD(lo(x),imm,r)");
    run<1, 1, 3, 8>("DFMA(x,y,c) + 3 LOP3(r,imm)");
    run<1, 1, 4, 8>("DFMA(x,y,c) + 4 LOP3(r,imm)");
    run<3, 1, 1, 8>("DFMA(x,Ys,Zs) + 1 LOP3(r,imm)");
    run<3, 1, 2, 8>("DFMA(x,Ys,Zs) + 2 LOP3(r,imm)");
    return 0;
}
