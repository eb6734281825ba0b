// gram_mm_tf32.cuh -- K4t: Float32 multi-RHS product  B <- alpha K A + beta B  with both GEMM-shaped phases on the tensor
// cores in 3xTF32 split precision.
//
// Replaces mul!(B::AbstractMatrix, G::Gramian{Float32}, A::AbstractMatrix, alpha, beta) (reference src/gramian.jl:89-99) for
// well-scaled points of dimension d >= 8.  The scalar Float32 kernel (cf_extra.cuh) is bound by shared-memory operand
// delivery at 18 TFLOP/s -- slower than the Float64 DMMA kernel.  TF32 tensor cores alone (10-bit mantissa) would miss the
// 1e-5 tolerance, so every operand is split in registers into hi = tf32(v), lo = tf32(v - hi) and each product is formed as
// lo.hi + hi.lo + hi.hi (fp32 accumulate; the dropped lo.lo term is 2^-22 relative): error ~ 2^-21 per product, inside the
// tolerance, at a third of the TF32 rate: 92 TFLOP/s effective with legacy mma.sync m16n8k8 (profiles/r1_tf32_probe.txt).
//   phase A   Dot (128 x 32) = Xs . Ys^T, r2 from the norms, program on the C fragments -> K tile (row-major, raw fp32)
//   phase B   Out (128 x 64) += Ks (128 x 32) . As (32 x 64)
// Fragment layout of mma.m16n8k8 (g = lane / 4, t = lane % 4):  A: a0 (m = g, k = t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B: b0 (k = t, n = g) b1 (k = t+4, n = g);  C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).  Which matrix element a
// slot holds is free as long as A, B and C agree, so the slots are RELABELLED: k slots (t, t+4) hold k = (2t, 2t+1) and row
// slots (g, g+8) hold tile rows (2g, 2g+1) of each 16-row block.  With the A operands stored k-major (XsT[k][row],
// KsT[k][row]) every register pair of a fragment -- (a0,a1), (a2,a3), (b0,b1), and the C pairs (c0,c2), (c1,c3) when they are
// stored as K values -- is then ONE aligned 8-byte shared-memory access, and the loaded registers ARE the fragment: the
// "hi" operand of the split is the raw value (the tensor core ignores the 13 low mantissa bits, i.e. truncates), only
// lo = v - trunc(v) costs two ALU instructions.  (A first version assembled fragments from scattered registers: ptxas
// inserted ~450 register moves per warp-tile.)
#pragma once
#include "cf_extra.cuh"

#define CF_MMT_SK (CF_MM_TI + 4)  // row stride of KsT[k][i] and XsT[k][i] (floats): 2 SK = 8 mod 32 -> conflict-free 8-byte accesses
#define CF_MMT_SA (CF_MM_PC + 4)  // row stride of At[j][c] (floats, global and shared): 2 SA = 8 mod 32

// row stride of the padded point copies (floats): smallest value >= D that is 8 mod 16 (conflict-free 8-byte fragment loads)
__host__ __device__ constexpr int cf_mmt_sx(int D) { return (D % 16 <= 8) ? (D / 16) * 16 + 8 : (D / 16) * 16 + 24; }

template <int D>
struct cf_mmt_layout {
    static constexpr int sx = cf_mmt_sx(D);
    static constexpr int bar_bytes = 128;
    static constexpr int dk = ((D + 7) / 8) * 8;  // k extent of the distance GEMM (zero padded)
    static constexpr int ks_bytes = CF_MM_TJ * CF_MMT_SK * 4;
    static constexpr int xs_bytes = dk * CF_MMT_SK * 4;
    static constexpr int y_bytes = CF_MM_TJ * sx * 4;
    static constexpr int n_bytes = CF_MM_TJ * 4;
    static constexpr int a_bytes = CF_MM_TJ * CF_MMT_SA * 4;
    static constexpr int stage_bytes = ((y_bytes + n_bytes + a_bytes + 127) / 128) * 128;
    static constexpr int total = bar_bytes + ks_bytes + xs_bytes + CF_MM_NS * stage_bytes;
};

// Xp[i][c] = c < D ? X[i][c] : 0 with row stride sx (one-off per handle)
static __global__ void cf_pad_rows_f32_kernel(const float* __restrict__ X, int D, int sx, int64_t n, float* __restrict__ Xp) {
    const int64_t total = n * sx;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q / sx;
        const int c = (int)(q - i * sx);
        Xp[q] = (c < D) ? X[i * D + c] : 0.f;
    }
}

// v = hi + lo with hi = v truncated to TF32's 10 mantissa bits and lo = v - hi (exact in fp32).  The tensor core reads only
// the TF32 bits of an operand, so the RAW value serves as hi and lo is truncated the same way; what is dropped is below
// 2^-20 |v|.  (cvt.rna.tf32.f32 compiles to a 6-instruction software rounding sequence on sm_100a.)
__device__ __forceinline__ uint32_t cf_tf32_lo(float v) {
    return __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
}
__device__ __forceinline__ void cf_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c += A . B in 3xTF32: small terms first
__device__ __forceinline__ void cf_mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                              const uint32_t (&bl)[2]) {
    cf_mma_tf32(c, al, bh);
    cf_mma_tf32(c, ah, bl);
    cf_mma_tf32(c, ah, bh);
}

// P.X: the points as uploaded (row stride D); P.Y: the padded Float32 copy (row stride sx); P.At has row stride CF_MMT_SA
template <int D>
__global__ void __launch_bounds__(256, 2) gram_mm_tf32_kernel(const __grid_constant__ cf_mm_params P) {
    using S = cf_mmt_layout<D>;
    constexpr int SX = S::sx, SK = CF_MMT_SK, SA = CF_MMT_SA, NTB = 256, DK = S::dk;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* Ks = reinterpret_cast<float*>(smem + S::bar_bytes);
    float* Xs = reinterpret_cast<float*>(smem + S::bar_bytes + S::ks_bytes);
    unsigned char* stages = smem + S::bar_bytes + S::ks_bytes + S::xs_bytes;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const float* __restrict__ Xg = static_cast<const float*>(P.X);
    const float* __restrict__ Yg = static_cast<const float*>(P.Y);
    const float* __restrict__ yng = static_cast<const float*>(P.yn);
    const float* __restrict__ Atg = static_cast<const float*>(P.At);
    if (tid == 0) {
        for (int s = 0; s < CF_MM_NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    const int nfull = (int)(P.m / CF_MM_TJ);
    auto issue = [&](int tile) {
        const int s = tile % CF_MM_NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = (int64_t)tile * CF_MM_TJ;
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + S::n_bytes + S::a_bytes));
        cf_tma_load_1d(st, Yg + j0 * SX, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, yng + j0, (uint32_t)S::n_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes + S::n_bytes, Atg + j0 * SA, (uint32_t)S::a_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < CF_MM_NS && t < nfull; t++) issue(t);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * CF_MM_TI;
    const int64_t rend = P.row0 + P.nrows;
    for (int q = tid; q < CF_MM_TI * DK; q += NTB) {  // XsT[k][row]: the row tile's points, transposed (rows past the end: clamped)
        const int row = q / DK, c = q - row * DK;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        Xs[c * SK + row] = (c < D) ? Xg[ir * D + c] : 0.f;
    }
    float xnorm[2];  // phase A rows of this lane: 16 w + 2 g + h  (row slots g, g + 8 of the 16-row block)
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int64_t i = rbase + 16 * w + 2 * g + h;
        if (i >= rend) i = rend - 1;
        xnorm[h] = static_cast<const float*>(P.xn)[i];
    }
    __syncthreads();
    // phase B tile of this warp: rows 32 (w / 2) + 16 mb + 2 g + h, columns 32 (w % 2) + 8 cb + 2 t4 + e
    const int brow = 32 * (w >> 1), bcol = 32 * (w & 1);
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[a][b][e] = 0.f;

    auto tile_compute = [&](const float* __restrict__ ys, const float* __restrict__ yns, const float* __restrict__ As, int cnt) {
        {
            float c[4][4];
#pragma unroll
            for (int cb = 0; cb < 4; cb++)
#pragma unroll
                for (int e = 0; e < 4; e++) c[cb][e] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < DK; k0 += 8) {
                // (a0, a1) = rows (2g, 2g+1) at k = k0 + 2 t4, (a2, a3) the same rows at k + 1
                const float2 x0 = *reinterpret_cast<const float2*>(&Xs[(k0 + 2 * t4) * SK + 16 * w + 2 * g]);
                const float2 x1 = *reinterpret_cast<const float2*>(&Xs[(k0 + 2 * t4 + 1) * SK + 16 * w + 2 * g]);
                const uint32_t ah[4] = {__float_as_uint(x0.x), __float_as_uint(x0.y), __float_as_uint(x1.x), __float_as_uint(x1.y)};
                const uint32_t al[4] = {cf_tf32_lo(x0.x), cf_tf32_lo(x0.y), cf_tf32_lo(x1.x), cf_tf32_lo(x1.y)};
#pragma unroll
                for (int cb = 0; cb < 4; cb++) {
                    const float2 yv = *reinterpret_cast<const float2*>(&ys[(8 * cb + g) * SX + k0 + 2 * t4]);
                    const uint32_t bh[2] = {__float_as_uint(yv.x), __float_as_uint(yv.y)};
                    const uint32_t bl[2] = {cf_tf32_lo(yv.x), cf_tf32_lo(yv.y)};
                    cf_mma_3xtf32(c[cb], ah, al, bh, bl);
                }
            }
            float yn8[8];  // this lane's columns: 8 cb + 2 t4 + e
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const float2 v = *reinterpret_cast<const float2*>(&yns[8 * cb + 2 * t4]);
                yn8[2 * cb] = v.x; yn8[2 * cb + 1] = v.y;
            }
            float kv[2][8];
#pragma unroll
            for (int h = 0; h < 2; h++) {  // row slots g (c0, c1) and g + 8 (c2, c3)
                float r2[8], dt[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    dt[u] = c[u >> 1][2 * h + (u & 1)];
                    const float v = fmaf(-2.f, dt[u], xnorm[h] + yn8[u]);
                    r2[u] = fmaxf(v, 0.f);
                }
                cf_sop_value_f32_n<8>(r2, dt, P.sop, kv[h]);
            }
            // KsT[col][row]: the two rows of a lane are adjacent -> one 8-byte store per column
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                float2 o;
                o.x = (col < cnt) ? kv[0][u] : 0.f;  // past the end: no contribution
                o.y = (col < cnt) ? kv[1][u] : 0.f;
                *reinterpret_cast<float2*>(&Ks[col * SK + 16 * w + 2 * g]) = o;
            }
        }
        __syncthreads();
        // the tensor core truncates when it adds to the accumulator: a sum over all column tiles kept in the MMA accumulator drifts
        // (3e-6 after 200 accumulations, measured).  Each tile therefore starts from zero and is added to the running sums with
        // round-to-nearest FADDs.
        float tacc[2][4][4];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int e = 0; e < 4; e++) tacc[a][b][e] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < CF_MM_TJ; k0 += 8) {
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mb = 0; mb < 2; mb++) {
                const float2 k0v = *reinterpret_cast<const float2*>(&Ks[(k0 + 2 * t4) * SK + brow + 16 * mb + 2 * g]);
                const float2 k1v = *reinterpret_cast<const float2*>(&Ks[(k0 + 2 * t4 + 1) * SK + brow + 16 * mb + 2 * g]);
                ah[mb][0] = __float_as_uint(k0v.x); ah[mb][1] = __float_as_uint(k0v.y);
                ah[mb][2] = __float_as_uint(k1v.x); ah[mb][3] = __float_as_uint(k1v.y);
                al[mb][0] = cf_tf32_lo(k0v.x); al[mb][1] = cf_tf32_lo(k0v.y);
                al[mb][2] = cf_tf32_lo(k1v.x); al[mb][3] = cf_tf32_lo(k1v.y);
            }
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const float b0 = As[(k0 + 2 * t4) * SA + bcol + 8 * cb + g];
                const float b1 = As[(k0 + 2 * t4 + 1) * SA + bcol + 8 * cb + g];
                const uint32_t bh[2] = {__float_as_uint(b0), __float_as_uint(b1)};
                const uint32_t bl[2] = {cf_tf32_lo(b0), cf_tf32_lo(b1)};
#pragma unroll
                for (int mb = 0; mb < 2; mb++) cf_mma_3xtf32(tacc[mb][cb], ah[mb], al[mb], bh, bl);
            }
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[a][b][e] += tacc[a][b][e];
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % CF_MM_NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / CF_MM_NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        tile_compute(reinterpret_cast<const float*>(st), reinterpret_cast<const float*>(st + S::y_bytes),
                     reinterpret_cast<const float*>(st + S::y_bytes + S::n_bytes), CF_MM_TJ);
        __syncthreads();  // Ks and stage s are free again
        if (tid == 0 && t + CF_MM_NS < nfull) issue(t + CF_MM_NS);
    }
    if ((int64_t)nfull * CF_MM_TJ < P.m) {  // ragged last tile: cooperative loads, zero fill
        const int64_t j0 = (int64_t)nfull * CF_MM_TJ;
        const int cnt = (int)(P.m - j0);
        float* ys = reinterpret_cast<float*>(stages);
        float* yns = reinterpret_cast<float*>(stages + S::y_bytes);
        float* As = reinterpret_cast<float*>(stages + S::y_bytes + S::n_bytes);
        __syncthreads();
        for (int q = tid; q < CF_MM_TJ * SX; q += NTB) ys[q] = (q < cnt * SX) ? Yg[j0 * SX + q] : 0.f;
        for (int q = tid; q < CF_MM_TJ; q += NTB) yns[q] = (q < cnt) ? yng[j0 + q] : 0.f;
        for (int q = tid; q < CF_MM_TJ * SA; q += NTB) As[q] = (q < cnt * SA) ? Atg[j0 * SA + q] : 0.f;
        __syncthreads();
        tile_compute(ys, yns, As, cnt);
    }
    float* Bg = static_cast<float*>(P.B);
#pragma unroll
    for (int cb = 0; cb < 4; cb++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = bcol + 8 * cb + 2 * t4 + e;
            if (c >= P.nrhs) continue;
#pragma unroll
            for (int mb = 0; mb < 2; mb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t i = rbase + brow + 16 * mb + 2 * g + h;
                    if (i >= rend) continue;
                    float* o = Bg + (i - P.row0) + P.ldb * c;
                    double v = P.alpha * (double)acc[mb][cb][2 * h + e];
                    if (P.beta != 0.0) v += P.beta * (double)(*o);
                    *o = (float)v;
                }
        }
}

#ifndef __CUDACC_RTC__ // host side
template <int D>
cudaError_t cf_mmt_launch(const cf_mm_params& P, int row_tiles, cudaStream_t stream) {
    using S = cf_mmt_layout<D>;
    auto kern = gram_mm_tf32_kernel<D>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<row_tiles, 256, S::total, stream>>>(P);
    return cudaGetLastError();
}

// registry hook: padded dimensions >= 8 (the k loop runs over multiples of 8; the padded copies are zero filled)
template <int D, bool OK = (D >= 8)>
struct cf_mmt_entry {
    static constexpr cf_mm_launch_fn fn = nullptr;
    static constexpr int sx = 0, smem = 0;
};
template <int D>
struct cf_mmt_entry<D, true> {
    static constexpr cf_mm_launch_fn fn = &cf_mmt_launch<D>;
    static constexpr int sx = cf_mmt_sx(D), smem = cf_mmt_layout<D>::total;
};
#endif // !__CUDACC_RTC__
