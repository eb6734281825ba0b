// gram_mm_tc5.cuh -- K4u: Float32 multi-RHS product  B <- alpha K A + beta B  on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in tensor memory), 3xTF32 split precision.
//
// Replaces mul!(B::AbstractMatrix, G::Gramian{Float32}, A::AbstractMatrix, alpha, beta) (reference src/gramian.jl:89-99) for
// well-scaled points of dimension d >= 8, like gram_mm_tf32.cuh, whose legacy mma.sync fragments it supersedes: there the warps
// that evaluate the kernel also load fragments and issue ~1700 mma.sync per tile; here ONE thread issues 36 asynchronous MMAs per
// 128 x 64 tile, operands are read by the tensor core straight from shared memory, accumulators never touch the register file,
// and the eight evaluation warps do nothing but the kernel program.
//
//   per row tile of 128 rows (one CTA), per column tile of TJ = 64 points:
//   phase A (tensor core)  Dot (128 x 64) = Xhi Yhi^T + Xlo Yhi^T + Xhi Ylo^T          12 MMAs M128 N64 K8 -> TMEM dot[t & 1]
//   evaluation (8 warps)   tcgen05.ld Dot -> r2 = |x|^2 + |y|^2 - 2 dot -> k (program interpreter, 8 entries at a time)
//                          -> K tile hi / lo written to TENSOR MEMORY (tcgen05.st): phase B takes its A operand from TMEM
//   phase B (tensor core)  Out (128 x 64) = Khi Ahi + Klo Ahi + Khi Alo                   24 MMAs M128 N64 K8 -> TMEM out
//   accumulate (8 warps)   tcgen05.ld Out -> running sums in registers with round-to-nearest adds (the tensor core truncates when it
//                          adds to its accumulator, so a tile starts from zero: see gram_mm_tf32.cuh)
// hi = the raw fp32 word (the tensor core reads only the TF32 bits, i.e. truncates), lo = v - trunc(v): the dropped lo.lo term
// is 2^-22 relative.  The column-side operands arrive by TMA bulk copies from per-handle / per-call images that are ALREADY in the
// tensor core's canonical shared-memory layout (cf_canon_* kernels below), so a stage is five 1-D copies and no thread touches it.
//
// Warp roles: CF_MMU_EW = 16 evaluation warps in two groups of 8 that take alternate column tiles (within a group warp w owns TMEM lanes
// 32 (w % 4) .. +31, i.e. tile rows, and columns 32 (w / 4) .. +31), one TMA producer warp, one phase-A and one phase-B issuing warp.  mbarriers: full / empty per stage (TMA <-> MMA), dotfull / dotfree per TMEM dot buffer,
// kfull (K tile written), outfull (phase B done: Out readable, K tile and stage free), outfree.
#pragma once
#include "gram_mm_tf32.cuh"

#define CF_MMU_TI 128
#define CF_MMU_TJ 64
#define CF_MMU_PC 64
#define CF_MMU_NS 3
#ifndef CF_MMU_EW
#define CF_MMU_EW 16                       // evaluation warps: two groups of 8 that take alternate column tiles
#endif
#define CF_MMU_GW (CF_MMU_EW / 2)          // warps per evaluation group
#define CF_MMU_CW (64 / (CF_MMU_GW / 4))   // tile columns (and right-hand sides) per evaluation warp
#define CF_MMU_THREADS (32 * CF_MMU_EW + 96)   // + TMA producer warp, phase-A issuer warp, phase-B issuer warp

// element (r, k) of an R x K tile of 4-byte values in the K-major canonical layout without swizzle (core matrix = 8 rows x 16 bytes,
// core matrices of one K chunk contiguous): LBO (next K chunk) = R / 8 * 128 bytes, SBO (next 8 rows) = 128 bytes
__host__ __device__ constexpr int cf_canon(int r, int k, int R) { return ((k >> 2) * (R >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3); }

template <int D>
struct cf_mmu_layout {
    static constexpr int dk = ((D + 7) / 8) * 8;
    static constexpr int x_bytes = CF_MMU_TI * dk * 4;            // each of hi, lo
    static constexpr int y_bytes = CF_MMU_TJ * dk * 4;            // each of hi, lo
    static constexpr int a_bytes = CF_MMU_PC * CF_MMU_TJ * 4;     // each of hi, lo
    static constexpr int n_bytes = CF_MMU_TJ * 4;
    static constexpr int k_bytes = CF_MMU_TI * CF_MMU_TJ * 4;     // each of hi, lo
    static constexpr int stage_bytes = 2 * y_bytes + 2 * a_bytes + ((n_bytes + 127) / 128) * 128;
    static constexpr int bar_bytes = 256;
    static constexpr int total = bar_bytes + 2 * x_bytes + CF_MMU_NS * stage_bytes + 1024;  // + alignment slack (the K tile lives in TMEM)
};

// ---- one-off images in the canonical layout ------------------------------------------------------------------------------------
// column points: per tile of TJ points an image [TJ x dk] of hi words and one of lo words (zero beyond m and beyond D)
static __global__ void cf_canon_points_kernel(const float* __restrict__ Y, int D, int dk, int64_t m, int64_t mpad, float* __restrict__ hi,
                                              float* __restrict__ lo, const float* __restrict__ yn, float* __restrict__ ynpad) {
    const int64_t total = mpad * dk;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = q / dk;
        const int k = (int)(q - j * dk);
        const float v = (j < m && k < D) ? Y[j * D + k] : 0.f;
        const int64_t dst = (j / CF_MMU_TJ) * (CF_MMU_TJ * dk) + cf_canon((int)(j % CF_MMU_TJ), k, CF_MMU_TJ);
        hi[dst] = v;
        lo[dst] = __uint_as_float(cf_tf32_lo(v));
        if (k == 0) ynpad[j] = (j < m) ? yn[j] : 0.f;
    }
}
// right-hand sides: per tile of TJ points an image [PC x TJ] (row = rhs column c, K index = point) of hi and lo words
static __global__ void cf_canon_rhs_kernel(const float* __restrict__ A, int64_t lda, int64_t m, int64_t mpad, int nrhs, float* __restrict__ hi,
                                           float* __restrict__ lo) {
    const int64_t total = mpad * CF_MMU_PC;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(q / mpad);
        const int64_t j = q - (int64_t)c * mpad;  // consecutive threads: consecutive points of one column (coalesced reads)
        const float v = (j < m && c < nrhs) ? A[j + lda * c] : 0.f;
        const int64_t dst = (j / CF_MMU_TJ) * (CF_MMU_PC * CF_MMU_TJ) + cf_canon(c, (int)(j % CF_MMU_TJ), CF_MMU_PC);
        hi[dst] = v;
        lo[dst] = __uint_as_float(cf_tf32_lo(v));
    }
}

struct cf_mmu_params {
    cf_mm_params mm;       // X (rows as uploaded, stride D), xn, B, ldb, row0, nrows, m, nrhs, alpha, beta, sop
    const float* yhi;      // canonical column-point images
    const float* ylo;
    const float* ynpad;    // squared norms of the columns, zero padded to a multiple of TJ
    const float* ahi;      // canonical right-hand-side images
    const float* alo;
    int64_t ntiles;        // column tiles
};

// ---- tcgen05 / TMEM wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t cf_umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46);  // version 1 (sm_100), no swizzle, base offset 0
}
// D[tmem] (+)= A[smem] B[smem]^T, M = 128, K = 8, kind::tf32, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void cf_umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the A operand in tensor memory (128 lanes x 8 columns of 32-bit words at tmem_a)
__device__ __forceinline__ void cf_umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cf_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {  // 32 consecutive columns of this thread's TMEM lane
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
                   "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ void cf_umma_commit(uint64_t* bar) {  // arrives on bar when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cf_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {  // 32 consecutive columns of this thread's TMEM lane
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cf_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cf_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void cf_tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
template <int N> __device__ __forceinline__ void cf_tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
    if constexpr (N == 32) cf_tmem_ld32(taddr, v); else cf_tmem_ld16(taddr, v);
}
template <int N> __device__ __forceinline__ void cf_tmem_st(uint32_t taddr, const uint32_t (&v)[N]) {
    if constexpr (N == 32) cf_tmem_st32(taddr, v); else cf_tmem_st16(taddr, v);
}
// one lane of a converged warp (elect.sync): the issuing warps run their loops warp-uniformly, so that descriptors live in uniform
// registers, and only the tcgen05.mma / commit instructions sit under the elected predicate
__device__ __forceinline__ bool cf_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cf_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cf_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cf_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cf_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int D>
__global__ void __launch_bounds__(CF_MMU_THREADS, 1) gram_mm_tc5_kernel(const __grid_constant__ cf_mmu_params PP) {
    using S = cf_mmu_layout<D>;
    constexpr int DK = S::dk, TI = CF_MMU_TI, TJ = CF_MMU_TJ, PC = CF_MMU_PC, NS = CF_MMU_NS;
    const cf_mm_params& P = PP.mm;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (cf_smem_u32(smem_raw) & 1023u)) & 1023u);  // 1 KB alignment inside the shared window
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t *full = bars, *empty = bars + NS, *dotfull = bars + 2 * NS, *dotfree = dotfull + 2, *kfull = dotfree + 2, *outfull = kfull + 2,
             *outfree = outfull + 2;  // dot / K / out buffers and their barriers are indexed by the tile parity = evaluation group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    float* Xhi = reinterpret_cast<float*>(smem + S::bar_bytes);
    float* Xlo = Xhi + TI * DK;
    unsigned char* stages = reinterpret_cast<unsigned char*>(Xlo + TI * DK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = (int)PP.ntiles;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { cf_mbar_init(&full[s], 1); cf_mbar_init(&empty[s], 1 + 32 * CF_MMU_GW); }  // phase-B commit + every thread of the tile's evaluation group
        for (int b = 0; b < 2; b++) {  // one arrival per evaluation warp of the group
            cf_mbar_init(&dotfull[b], 1); cf_mbar_init(&dotfree[b], CF_MMU_GW);
            cf_mbar_init(&kfull[b], CF_MMU_GW); cf_mbar_init(&outfull[b], 1); cf_mbar_init(&outfree[b], CF_MMU_GW);
        }
        cf_fence_barrier_init();
    }
    if (warp == CF_MMU_EW + 1) {  // 512 TMEM columns (320 used; the allocation must be a power of two)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(cf_smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * TI;
    const int64_t rend = P.row0 + P.nrows;
    // the row tile's points, hi / lo, canonical layout (rows past the end: clamped, never stored)
    const float* __restrict__ Xg = static_cast<const float*>(P.X);
    for (int q = tid; q < TI * (DK / 4); q += CF_MMU_THREADS) {
        const int row = q % TI, kc = q / TI;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        float4 h, l;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (4 * kc + e < D) ? Xg[ir * D + 4 * kc + e] : 0.f;
        h = make_float4(v[0], v[1], v[2], v[3]);
        l = make_float4(__uint_as_float(cf_tf32_lo(v[0])), __uint_as_float(cf_tf32_lo(v[1])), __uint_as_float(cf_tf32_lo(v[2])),
                        __uint_as_float(cf_tf32_lo(v[3])));
        *reinterpret_cast<float4*>(&Xhi[cf_canon(row, 4 * kc, TI)]) = h;
        *reinterpret_cast<float4*>(&Xlo[cf_canon(row, 4 * kc, TI)]) = l;
    }
    cf_fence_async_smem();
    cf_tc_fence_before();
    __syncthreads();
    cf_tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns (64 each): dot[2], out[2], K hi[2], K lo[2], indexed by tile parity.  The per-tile chain
    //   dot ready -> tcgen05.ld -> kernel values -> tcgen05.st K -> phase B -> tcgen05.ld Out
    // is latency-bound (each TMEM access and mbarrier hand-over costs a few hundred cycles), so the evaluation warps form TWO groups that
    // take alternate tiles, each with its own dot / K / out buffers: while one group waits on a hand-over the other computes.
    // (Measured alternatives: 8 or 16 warps in ONE group 84 / 85 ms; accumulators split by k-step parity 90 ms; one issuing thread
    // for both phases 111 ms.)
    const uint32_t tm_dot[2] = {tmem, tmem + 64}, tm_out[2] = {tmem + 128, tmem + 192}, tm_khi[2] = {tmem + 256, tmem + 320},
                   tm_klo[2] = {tmem + 384, tmem + 448};

    if (warp == CF_MMU_EW) {
        // ---- TMA producer ------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            for (int t = 0; t < ntiles; t++) {
                const int s = t % NS;
                if (t >= NS) cf_mbar_wait(&empty[s], (uint32_t)(((t / NS) - 1) & 1));
                unsigned char* st = stages + (size_t)s * S::stage_bytes;
                cf_mbar_expect_tx(&full[s], (uint32_t)(2 * S::y_bytes + 2 * S::a_bytes + S::n_bytes));
                cf_tma_load_1d(st, PP.yhi + (int64_t)t * TJ * DK, (uint32_t)S::y_bytes, &full[s]);
                cf_tma_load_1d(st + S::y_bytes, PP.ylo + (int64_t)t * TJ * DK, (uint32_t)S::y_bytes, &full[s]);
                cf_tma_load_1d(st + 2 * S::y_bytes, PP.ahi + (int64_t)t * PC * TJ, (uint32_t)S::a_bytes, &full[s]);
                cf_tma_load_1d(st + 2 * S::y_bytes + S::a_bytes, PP.alo + (int64_t)t * PC * TJ, (uint32_t)S::a_bytes, &full[s]);
                cf_tma_load_1d(st + 2 * S::y_bytes + 2 * S::a_bytes, PP.ynpad + (int64_t)t * TJ, (uint32_t)S::n_bytes, &full[s]);
            }
        }
    } else if (warp == CF_MMU_EW + 1) {
        // ---- phase-A issuer: the distance GEMM of tile t as soon as its stage has landed and TMEM dot[t & 1] has been read -----------------
        // (two issuing warps, one per phase: a single thread spends ~50-100 cycles per tcgen05.mma on descriptor set-up, 36 MMAs per tile)
        {
            // instruction descriptor: D fp32, A / B tf32, both K-major, N = 64, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t LBO_X = (TI / 8) * 128, LBO_Y = (TJ / 8) * 128;
            const uint64_t dxh = cf_umma_desc(cf_smem_u32(Xhi), LBO_X, 128), dxl = cf_umma_desc(cf_smem_u32(Xlo), LBO_X, 128);
            for (int t = 0; t < ntiles; t++) {
                const int s = t % NS;
                cf_mbar_wait(&full[s], (uint32_t)((t / NS) & 1));
                if (t >= 2) cf_mbar_wait(&dotfree[t & 1], (uint32_t)(((t >> 1) - 1) & 1));
                cf_tc_fence_after();
                const uint64_t dyh = cf_umma_desc(cf_smem_u32(stages + (size_t)s * S::stage_bytes), LBO_Y, 128), dyl = dyh + (S::y_bytes >> 4);
                if (cf_elect_one()) {
#pragma unroll
                    for (int pr = 0; pr < 3; pr++) {  // small terms first: lo.hi, hi.lo, hi.hi
                        const uint64_t xa = pr == 0 ? dxl : dxh, yb = pr == 1 ? dyl : dyh;
#pragma unroll
                        for (int ks = 0; ks < DK / 8; ks++) {  // the start-address field advances by two K chunks per step
                            cf_umma_tf32(tm_dot[t & 1], xa + (uint64_t)((ks * 2 * LBO_X) >> 4), yb + (uint64_t)((ks * 2 * LBO_Y) >> 4), idesc,
                                         (pr > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    cf_umma_commit(&dotfull[t & 1]);
                }
                __syncwarp();
            }
        }
    } else if (warp == CF_MMU_EW + 2) {
        // ---- phase-B issuer: Out = K A as soon as the evaluation warps have written K tile t to tensor memory ----------------------------------
        {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t LBO_A = (PC / 8) * 128;
            for (int t = 0; t < ntiles; t++) {
                const int s = t % NS;
                const int g = t & 1;
                cf_mbar_wait(&kfull[g], (uint32_t)((t >> 1) & 1));                            // K tile t written (TMEM) by evaluation group g
                if (t >= 2) cf_mbar_wait(&outfree[g], (uint32_t)(((t >> 1) - 1) & 1));        // Out of tile t - 2 read
                cf_tc_fence_after();
                const uint64_t dah = cf_umma_desc(cf_smem_u32(stages + (size_t)s * S::stage_bytes + 2 * S::y_bytes), LBO_A, 128), dal = dah + (S::a_bytes >> 4);
                if (cf_elect_one()) {
#pragma unroll
                    for (int pr = 0; pr < 3; pr++) {
                        const uint32_t ka = pr == 0 ? tm_klo[g] : tm_khi[g];
                        const uint64_t ab = pr == 1 ? dal : dah;
#pragma unroll
                        for (int ks = 0; ks < TJ / 8; ks++) {
                            cf_umma_tf32_ta(tm_out[g], ka + 8 * ks, ab + (uint64_t)((ks * 2 * LBO_A) >> 4), idesc, (pr > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    cf_umma_commit(&outfull[g]);  // Out readable, K tile free
                    cf_umma_commit(&empty[s]);    // stage s free for the producer
                }
                __syncwarp();
            }
        }
    } else {
        // ---- evaluation warps ----------------------------------------------------------------------------------------------------------
        constexpr int CW = CF_MMU_CW;
        const int grp = warp / CF_MMU_GW, wg = warp % CF_MMU_GW;  // evaluation group (tile parity), warp within the group
        const int q4 = wg & 3, cq = wg >> 2;                      // TMEM lane quarter (tile rows 32 q4 ..), column group (CW cq ..)
        const int row = 32 * q4 + lane;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        const float xnorm = static_cast<const float*>(P.xn)[ir];
        const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
        float acc[CW];
#pragma unroll
        for (int c = 0; c < CW; c++) acc[c] = 0.f;
        auto drain_out = [&](int t) {  // Out of tile t (of this group) -> running sums
            cf_mbar_wait(&outfull[grp], (uint32_t)((t >> 1) & 1));
            cf_tc_fence_after();
            uint32_t o[CW];
            cf_tmem_ld<CW>(tm_out[grp] + lane_base + CW * cq, o);
            cf_tc_fence_before();
            __syncwarp();
            if (lane == 0) cf_mbar_arrive(&outfree[grp]);
#pragma unroll
            for (int c = 0; c < CW; c++) acc[c] += __uint_as_float(o[c]);
        };
        int last = -1;
        for (int t = grp; t < ntiles; t += 2) {
            const int s = t % NS;
            // |y|^2 of the tile is read from the stage below: observe the TMA completion directly (it happened long ago -- the MMA issuer
            // waited for it before the distance GEMM -- but only a wait on full[s] orders the bulk copy's writes before THIS thread's reads)
            cf_mbar_wait(&full[s], (uint32_t)((t / NS) & 1));
            cf_mbar_wait(&dotfull[grp], (uint32_t)((t >> 1) & 1));
            cf_tc_fence_after();
            uint32_t dv[CW];
            cf_tmem_ld<CW>(tm_dot[grp] + lane_base + CW * cq, dv);
            cf_tc_fence_before();
            __syncwarp();
            if (lane == 0) cf_mbar_arrive(&dotfree[grp]);
            // the stage of tile t is still resident (freed only when phase B of tile t completes): |y|^2 of this warp's columns
            const float* yns = reinterpret_cast<const float*>(stages + (size_t)s * S::stage_bytes + 2 * S::y_bytes + 2 * S::a_bytes) + CW * cq;
            float kv[CW];
#pragma unroll
            for (int g = 0; g < CW / 8; g++) {
                float r2[8], dt[8], k8[8];
                const float4 n0 = *reinterpret_cast<const float4*>(yns + 8 * g), n1 = *reinterpret_cast<const float4*>(yns + 8 * g + 4);
                const float yn8[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    dt[u] = __uint_as_float(dv[8 * g + u]);
                    r2[u] = fmaxf(fmaf(-2.f, dt[u], xnorm + yn8[u]), 0.f);
                }
                cf_sop_value_f32_n<8>(r2, dt, P.sop, k8);
#pragma unroll
                for (int u = 0; u < 8; u++) kv[8 * g + u] = k8[u];
            }
            if (last >= 0) drain_out(last);  // phase B of this group's previous tile is complete: its Out is added, the K buffer may be overwritten
            // K tile t, hi / lo, into tensor memory: lane = tile row, column = tile column (the A operand of phase B)
            uint32_t kh[CW], kl[CW];
#pragma unroll
            for (int c = 0; c < CW; c++) { kh[c] = __float_as_uint(kv[c]); kl[c] = cf_tf32_lo(kv[c]); }
            cf_tmem_st<CW>(tm_khi[grp] + lane_base + CW * cq, kh);
            cf_tmem_st<CW>(tm_klo[grp] + lane_base + CW * cq, kl);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            cf_tc_fence_before();
            __syncwarp();
            if (lane == 0) cf_mbar_arrive(&kfull[grp]);
            // this thread has read the stage's |y|^2 (the tensor core's share arrives with the phase-B commit).  Every lane arrives itself:
            // lane 0 arriving after __syncwarp orders the same accesses, but compute-sanitizer's racecheck only follows a thread's OWN
            // arrival (profiles/r2_sanitizer.txt), and a warp-wide arrive costs the same one instruction.
            cf_mbar_arrive(&empty[s]);
            last = t;
        }
        if (last >= 0) drain_out(last);
        // the two groups hold the sums over the even and the odd tiles: group 1 hands its sums to group 0 through shared memory
        float* xch = reinterpret_cast<float*>(stages);  // all stages are free once every phase B has completed
        asm volatile("bar.sync 1, %0;" ::"r"(CF_MMU_EW * 32) : "memory");  // all evaluation warps: every tile drained
        if (grp == 1) {
#pragma unroll
            for (int c = 0; c < CW; c++) xch[(CW * cq + c) * TI + row] = acc[c];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(CF_MMU_EW * 32) : "memory");
        if (grp == 0) {
#pragma unroll
            for (int c = 0; c < CW; c++) acc[c] += xch[(CW * cq + c) * TI + row];
        }
        float* Bg = static_cast<float*>(P.B);
        const int64_t i = rbase + row;
        if (grp == 0 && i < rend) {
#pragma unroll
            for (int c = 0; c < CW; c++) {
                const int col = CW * cq + c;
                if (col < P.nrhs) {
                    float* o = Bg + (i - P.row0) + P.ldb * col;
                    double v = P.alpha * (double)acc[c];
                    if (P.beta != 0.0) v += P.beta * (double)(*o);
                    *o = (float)v;
                }
            }
        }
    }
    cf_tc_fence_before();
    __syncthreads();
    if (warp == CF_MMU_EW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

#ifndef __CUDACC_RTC__
typedef cudaError_t (*cf_mmu_launch_fn)(const cf_mmu_params& P, int row_tiles, cudaStream_t stream);
template <int D>
cudaError_t cf_mmu_launch(const cf_mmu_params& P, int row_tiles, cudaStream_t stream) {
    using S = cf_mmu_layout<D>;
    auto kern = gram_mm_tc5_kernel<D>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<row_tiles, CF_MMU_THREADS, S::total, stream>>>(P);
    return cudaGetLastError();
}
template <int D, bool OK = (D >= 8)>
struct cf_mmu_entry {
    static constexpr cf_mmu_launch_fn fn = nullptr;
    static constexpr int dk = 0, smem = 0;
};
template <int D>
struct cf_mmu_entry<D, true> {
    static constexpr cf_mmu_launch_fn fn = &cf_mmu_launch<D>;
    static constexpr int dk = cf_mmu_layout<D>::dk, smem = cf_mmu_layout<D>::total;
};
#endif
