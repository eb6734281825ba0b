// gram_mm_dmma.cuh -- K4d: Float64 multi-RHS product  B <- alpha K A + beta B  on the FP64 tensor-core path (DMMA m8n8k4).
//
// Replaces mul!(B::AbstractMatrix, G::Gramian, A::AbstractMatrix, alpha, beta) (reference src/gramian.jl:89-99) for
// well-scaled points of dimension d >= 8 (padded D a multiple of 4), where r^2 = |x|^2 + |y|^2 - 2 x.y is safe.
//
// Why tensor cores although DMMA has no more FLOP/s than DFMA on B200 (profiles/r1_dmma_probe.txt): the scalar
// kernel (cf_extra.cuh gram_mm_kernel) is bound by SHARED-MEMORY OPERAND DELIVERY, not by the FP64 pipe -- a broadcast
// LDS.128 costs 2.1 SM-cycles and a per-lane LDS.64 2.0 (bench_aux/micro/lds_probe.cu), and an 8x4 register tile needs
// 0.375 operand doubles per FMA.  One DMMA performs 256 FMAs from 2 operand doubles per lane (0.008 per FMA), so both
// GEMM-shaped phases of the tile run at the FP64 pipe rate:
//   phase A   Dot (128 x 32)  = Xs (128 x D) . Ys^T (D x 32)         D/4 k-steps, 8 DMMA each per warp
//             K_ij = k(r2_ij = xn_i + yn_j - 2 Dot_ij, Dot_ij)        program evaluated on the C fragments (8 entries a time)
//   phase B   Out (128 x 64) += Ks (128 x 32) . As (32 x 64)          8 k-steps, 16 DMMA each per warp
// Fragment layout of mma.sync.m8n8k4.row.col.f64 (g = lane / 4, t = lane % 4):  A[m = g][k = t],  B[k = t][n = g],
// C[m = g][n = 2 t + {0, 1}].  Every operand array has a row stride = 4 (mod 8) doubles, which makes the per-lane
// LDS.64 of a fragment conflict-free (the 4 k-values of a half-warp land on 4 disjoint groups of 8 banks).
#pragma once
#include "cf_extra.cuh"

#define CF_MMD_SK (CF_MM_TI + 4)  // row stride of Ks[k][i]
#define CF_MMD_SA (CF_MM_PC + 4)  // row stride of At[j][c] (global and shared)

template <int D>
struct cf_mmd_smem {
    // row stride of the padded point copies: = 4 (mod 8) and >= the next multiple of 8 (the gradient kernel reads 8-wide column blocks)
    static constexpr int sx = (D % 8 == 4) ? D + 8 : D + 4;
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int ks_bytes = CF_MM_TJ * CF_MMD_SK * 8;
    static constexpr int xs_bytes = CF_MM_TI * sx * 8;
    static constexpr int y_bytes = CF_MM_TJ * sx * 8;
    static constexpr int n_bytes = CF_MM_TJ * 8;
    static constexpr int a_bytes = CF_MM_TJ * CF_MMD_SA * 8;
    static constexpr int stage_bytes = ((y_bytes + n_bytes + a_bytes + 127) / 128) * 128;
    static constexpr int total = tbl_bytes + bar_bytes + ks_bytes + xs_bytes + CF_MM_NS * stage_bytes;
};

// Xp[i][c] = c < D ? X[i][c] : 0  with row stride sx (one-off per handle)
static __global__ void cf_pad_rows_kernel(const double* __restrict__ X, int D, int sx, int64_t n, double* __restrict__ Xp) {
    const int64_t total = n * sx;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q / sx;
        const int c = (int)(q - i * sx);
        Xp[q] = (c < D) ? X[i * D + c] : 0.0;
    }
}

__device__ __forceinline__ void cf_dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// P.X / P.Y point at the PADDED copies (row stride sx); P.At has row stride CF_MMD_SA
template <int D>
__global__ void __launch_bounds__(256, 1) gram_mm_dmma_kernel(const __grid_constant__ cf_mm_params P) {
    using S = cf_mmd_smem<D>;
    constexpr int SX = S::sx, SK = CF_MMD_SK, SA = CF_MMD_SA, NTB = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    double* Ks = reinterpret_cast<double*>(smem + S::tbl_bytes + S::bar_bytes);
    double* Xs = reinterpret_cast<double*>(smem + S::tbl_bytes + S::bar_bytes + S::ks_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes + S::ks_bytes + S::xs_bytes;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t4 = lane & 3;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const double* __restrict__ Xg = static_cast<const double*>(P.X);
    const double* __restrict__ Yg = static_cast<const double*>(P.Y);
    const double* __restrict__ yng = static_cast<const double*>(P.yn);
    const double* __restrict__ Atg = static_cast<const double*>(P.At);
    cf_fill_exp_table(tbl, P.exp2_tbl, tid, NTB);
    if (tid == 0) {
        for (int s = 0; s < CF_MM_NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    cf_tbl_publish(tbl_lane);
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
    for (int q = tid; q < CF_MM_TI * SX; q += NTB) {  // the row tile's points (rows past the end: clamped, never stored)
        const int row = q / SX;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        Xs[q] = Xg[ir * SX + (q - row * SX)];
    }
    // phase A rows of this lane: 16 w + 8 rb + g
    double xnorm[2];
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        int64_t i = rbase + 16 * w + 8 * rb + g;
        if (i >= rend) i = rend - 1;
        xnorm[rb] = static_cast<const double*>(P.xn)[i];
    }
    __syncthreads();
    // phase B tile of this warp: rows 32 (w / 2) + 8 rb + g, columns 32 (w % 2) + 8 cb + 2 t4 + e
    const int brow = 32 * (w >> 1), bcol = 32 * (w & 1);
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto tile_compute = [&](const double* __restrict__ ys, const double* __restrict__ yns, const double* __restrict__ As, int cnt) {
        {
            double c[2][4][2];
#pragma unroll
            for (int rb = 0; rb < 2; rb++)
#pragma unroll
                for (int cb = 0; cb < 4; cb++) c[rb][cb][0] = c[rb][cb][1] = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < D; k0 += 4) {
                double a[2], b[4];
#pragma unroll
                for (int rb = 0; rb < 2; rb++) a[rb] = Xs[(16 * w + 8 * rb + g) * SX + k0 + t4];
#pragma unroll
                for (int cb = 0; cb < 4; cb++) b[cb] = ys[(8 * cb + g) * SX + k0 + t4];
#pragma unroll
                for (int rb = 0; rb < 2; rb++)
#pragma unroll
                    for (int cb = 0; cb < 4; cb++) cf_dmma884(c[rb][cb], a[rb], b[cb]);
            }
            double yn8[8];
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const double2 v = *reinterpret_cast<const double2*>(&yns[8 * cb + 2 * t4]);
                yn8[2 * cb] = v.x; yn8[2 * cb + 1] = v.y;
            }
#pragma unroll
            for (int rb = 0; rb < 2; rb++) {
                double r2[8], dt[8], kv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    dt[u] = c[rb][u >> 1][u & 1];
                    const double v = fma(-2.0, dt[u], xnorm[rb] + yn8[u]);
                    r2[u] = (__double2hiint(v) < 0) ? 0.0 : v;  // rounding can leave a tiny negative value: sign-bit test, no FP64 compare
                }
                cf_sop_value_n<8>(r2, dt, P.sop, tbl_lane, kv);
                const int row = 16 * w + 8 * rb + g;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                    Ks[col * SK + row] = (col < cnt) ? kv[u] : 0.0;  // past the end: no contribution
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k0 = 0; k0 < CF_MM_TJ; k0 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int rb = 0; rb < 4; rb++) a[rb] = Ks[(k0 + t4) * SK + brow + 8 * rb + g];
#pragma unroll
            for (int cb = 0; cb < 4; cb++) b[cb] = As[(k0 + t4) * SA + bcol + 8 * cb + g];
#pragma unroll
            for (int rb = 0; rb < 4; rb++)
#pragma unroll
                for (int cb = 0; cb < 4; cb++) cf_dmma884(acc[rb][cb], a[rb], b[cb]);
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % CF_MM_NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / CF_MM_NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        tile_compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                     reinterpret_cast<const double*>(st + S::y_bytes + S::n_bytes), CF_MM_TJ);
        __syncthreads();  // Ks and stage s are free again
        if (tid == 0 && t + CF_MM_NS < nfull) issue(t + CF_MM_NS);
    }
    if ((int64_t)nfull * CF_MM_TJ < P.m) {  // ragged last tile: cooperative loads, zero fill
        const int64_t j0 = (int64_t)nfull * CF_MM_TJ;
        const int cnt = (int)(P.m - j0);
        double* ys = reinterpret_cast<double*>(stages);
        double* yns = reinterpret_cast<double*>(stages + S::y_bytes);
        double* As = reinterpret_cast<double*>(stages + S::y_bytes + S::n_bytes);
        __syncthreads();
        for (int q = tid; q < CF_MM_TJ * SX; q += NTB) ys[q] = (q < cnt * SX) ? Yg[j0 * SX + q] : 0.0;
        for (int q = tid; q < CF_MM_TJ; q += NTB) yns[q] = (q < cnt) ? yng[j0 + q] : 0.0;
        for (int q = tid; q < CF_MM_TJ * SA; q += NTB) As[q] = (q < cnt * SA) ? Atg[j0 * SA + q] : 0.0;
        __syncthreads();
        tile_compute(ys, yns, As, cnt);
    }
    double* Bg = static_cast<double*>(P.B);
#pragma unroll
    for (int cb = 0; cb < 4; cb++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = bcol + 8 * cb + 2 * t4 + e;
            if (c >= P.nrhs) continue;
#pragma unroll
            for (int rb = 0; rb < 4; rb++) {
                const int64_t i = rbase + brow + 8 * rb + g;
                if (i >= rend) continue;
                double* o = Bg + (i - P.row0) + P.ldb * c;
                double v = P.alpha * acc[rb][cb][e];
                if (P.beta != 0.0) v += P.beta * (*o);
                *o = v;
            }
        }
}

#ifndef __CUDACC_RTC__ // host side: not part of run-time specialised builds
template <int D>
cudaError_t cf_mmd_launch(const cf_mm_params& P, int row_tiles, cudaStream_t stream) {
    using S = cf_mmd_smem<D>;
    auto kern = gram_mm_dmma_kernel<D>;
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

// registry hook: the DMMA variant exists for padded dimensions that are multiples of 4 and at least 8
template <int D, bool OK = (D >= 8 && D % 4 == 0)>
struct cf_mmd_entry {
    static constexpr cf_mm_launch_fn fn = nullptr;
    static constexpr int sx = 0, smem = 0;
};
template <int D>
struct cf_mmd_entry<D, true> {
    static constexpr cf_mm_launch_fn fn = &cf_mmd_launch<D>;
    static constexpr int sx = cf_mmd_smem<D>::sx, smem = cf_mmd_smem<D>::total;
};
#endif // !__CUDACC_RTC__
