// Probe: one tcgen05.mma kind::tf32 GEMM D(128 x 64) = A(128 x 32) B(64 x 32)^T from shared memory in the canonical K-major
// no-swizzle layout, accumulator in TMEM, read back with tcgen05.ld 32x32b -- validates the descriptor conventions used by
// csrc/gram_mm_tc5.cuh.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc5_probe tc5_probe.cu && ./tc5_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (r, k) of an R x K tf32 tile, K-major canonical layout without swizzle: core matrix = 8 rows x 16 bytes
__host__ __device__ inline int canon(int r, int k, int R) { return ((k >> 2) * (R >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // layout type 0 = no swizzle, base offset 0
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* As = reinterpret_cast<float*>(smem);               // 128 x 32
    float* Bs = As + 128 * 32;                                // 64 x 32
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + 64 * 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int q = tid; q < 128 * 32; q += 128) As[canon(q / 32, q % 32, 128)] = A[q];
    for (int q = tid; q < 64 * 32; q += 128) Bs[canon(q / 32, q % 32, 64)] = B[q];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        // instruction descriptor: c_format F32 (1) at bit 4, a/b format TF32 (2) at bits 7 / 10, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 4; ks++) {  // K = 32 = 4 steps of 8
            const uint64_t da = make_desc(smem_u32(As) + ks * 2 * 2048, 2048, 128);
            const uint64_t db = make_desc(smem_u32(Bs) + ks * 2 * 1024, 1024, 128);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // everyone waits for the MMAs
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // thread tid reads TMEM lane tid (warp w may touch lanes 32 w .. 32 w + 31), 64 columns in two loads of 32
    uint32_t v[64];
    for (int h = 0; h < 2; h++) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + h * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[h * 32 + 0]), "=r"(v[h * 32 + 1]), "=r"(v[h * 32 + 2]), "=r"(v[h * 32 + 3]), "=r"(v[h * 32 + 4]), "=r"(v[h * 32 + 5]), "=r"(v[h * 32 + 6]), "=r"(v[h * 32 + 7]),
                       "=r"(v[h * 32 + 8]), "=r"(v[h * 32 + 9]), "=r"(v[h * 32 + 10]), "=r"(v[h * 32 + 11]), "=r"(v[h * 32 + 12]), "=r"(v[h * 32 + 13]), "=r"(v[h * 32 + 14]), "=r"(v[h * 32 + 15]),
                       "=r"(v[h * 32 + 16]), "=r"(v[h * 32 + 17]), "=r"(v[h * 32 + 18]), "=r"(v[h * 32 + 19]), "=r"(v[h * 32 + 20]), "=r"(v[h * 32 + 21]), "=r"(v[h * 32 + 22]), "=r"(v[h * 32 + 23]),
                       "=r"(v[h * 32 + 24]), "=r"(v[h * 32 + 25]), "=r"(v[h * 32 + 26]), "=r"(v[h * 32 + 27]), "=r"(v[h * 32 + 28]), "=r"(v[h * 32 + 29]), "=r"(v[h * 32 + 30]), "=r"(v[h * 32 + 31])
                     : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 64; c++) D[tid * 64 + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
    float hA[128 * 32], hB[64 * 32], hD[128 * 64];
    srand(3);
    for (auto& v : hA) v = (float)((rand() % 17) - 8) * 0.25f;   // exactly representable in tf32
    for (auto& v : hB) v = (float)((rand() % 13) - 6) * 0.5f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, sizeof(hD));
    const int smem = (128 * 32 + 64 * 32) * 4 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < 64; j++) {
            double s = 0;
            for (int k = 0; k < 32; k++) s += (double)hA[i * 32 + k] * hB[j * 32 + k];
            worst = fmax(worst, fabs(s - hD[i * 64 + j]));
        }
    printf("tcgen05 kind::tf32 128x64x32: max |error| = %g  (D[0][0..3] = %g %g %g %g)  %s\n", worst, hD[0], hD[1], hD[2], hD[3], worst == 0 ? "OK" : "MISMATCH");
    return worst == 0 ? 0 : 2;
}
