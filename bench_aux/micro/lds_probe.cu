// lds_probe.cu -- cost of shared-memory loads on sm_100a: broadcast vs per-lane, 64- vs 128-bit, alone and mixed with DFMA.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_probe lds_probe.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NF>
__global__ void probe(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int q = threadIdx.x; q < 8192; q += blockDim.x) sm[q] = 1.0 + 1e-9 * q;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // MODE 0: LDS.128 broadcast; 1: LDS.64 broadcast; 2: LDS.128 per-lane (conflict-free, stride 16 B); 3: LDS.64 per lane
    int base = (MODE == 0 || MODE == 1) ? 0 : (MODE == 2 ? lane * 2 : lane);
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double f0 = 1.0 + 1e-12 * lane, f1 = 0.5;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int off = (base + ((it * 8 + u) * 64)) & 4095;
            if (MODE == 0 || MODE == 2) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[off]);
                acc[u] += v.x; acc[(u + 1) & 7] += v.y;
            } else {
                acc[u] += sm[off];
            }
#pragma unroll
            for (int f = 0; f < NF; f++) acc[(u + f + 2) & 7] = fma(acc[(u + f + 2) & 7], f0, f1);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int u = 0; u < 8; u++) s += acc[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int NF>
void run(const char* name, int nthreads) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    cudaFuncSetAttribute(probe<MODE, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<MODE, NF><<<148, nthreads, 65536>>>(out, cyc, iters);
    probe<MODE, NF><<<148, nthreads, 65536>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    const double nlds = (double)iters * 8 * (nthreads / 32);
    // the accumulate per load adds 1 (64-bit) or 2 (128-bit) DADD per load in addition to NF DFMA
    printf("%-34s threads %4d NF %d: %.2f SM-cycles per warp-LDS (%.2f cycles per warp-LDS per scheduler)\n", name, nthreads, NF,
           c / nlds, c / (nlds / 4));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int nt : {256, 512, 1024}) {
        run<0, 0>("LDS.128 broadcast", nt);
        run<1, 0>("LDS.64 broadcast", nt);
        run<2, 0>("LDS.128 per-lane", nt);
        run<3, 0>("LDS.64 per-lane", nt);
    }
    run<0, 2>("LDS.128 broadcast + 2 DFMA", 256);
    run<0, 4>("LDS.128 broadcast + 4 DFMA", 256);
    run<2, 4>("LDS.128 per-lane + 4 DFMA", 256);
    run<0, 2>("LDS.128 broadcast + 2 DFMA", 512);
    run<0, 4>("LDS.128 broadcast + 4 DFMA", 512);
    return 0;
}
