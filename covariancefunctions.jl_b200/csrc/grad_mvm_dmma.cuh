// grad_mvm_dmma.cuh -- K5d: isotropic GradientKernel O(n^2 d) matrix-vector product with every d-dependent operation on the
// FP64 tensor cores (DMMA m8n8k4).  Float64, padded D in {8, 12, 16, 24, 32}, well-scaled points (same host check as K1d / K4d).
//
// Replaces blockmul!(y, G::Gramian, x, alpha, beta) (reference src/gramian.jl:241-253) with the lazy
// IsotropicGradientKernelElement product (reference src/gradient.jl:86-92)
//     b_i += -2 (k1 a_j + 2 k2 r (r.a_j)),   r = x_i - y_j,  (k1, k2) = (k'(r2), k''(r2)).
// The scalar kernel K5 (grad_mvm.cuh) needs 5 D FP64 instructions per block and 2 D operand doubles from shared memory.
// Written with ca = -2 k1, cw = -4 k2 (r.a_j):
//     r2_ij   = |x_i|^2 + |y_j|^2 - 2 (X Y^T)_ij                       GEMM 1  (D FMA per block)
//     r.a_j   = (X A^T)_ij - y_j.a_j                                    GEMM 2  (D FMA per block; y_j.a_j once per column)
//     b_i     = (Ca A)_i - (Cw Y)_i + x_i sum_j cw_ij                    GEMM 3  ([Ca | -Cw] (128 x 64) . [A; Y] (64 x D): 2 D FMA)
// i.e. 4 D FMAs per block, all on DMMA, which also removes the shared-memory operand traffic.  Per 128 x 32 tile: phase A
// (GEMM 1 + 2 with shared X fragments, jets on the C fragments, coefficient tiles to shared memory), phase B (GEMM 3).
#pragma once
#include "grad_mvm.cuh"
#include "gram_mm_dmma.cuh"

#define CF_GD_TI 128
#define CF_GD_TJ 32
#define CF_GD_NS 2
#define CF_GD_SC (CF_GD_TI + 4)  // row stride of the coefficient tile Cc[k][i]

struct cf_gradd_params {
    cf_grad_params g;   // X, Y, a point at the copies with the padded row stride; partial as in K5
    const double* xn;   // squared norms
    const double* yn;
    const double* q;    // q_j = y_j . a_j
};

template <int D>
struct cf_gd_smem {
    static constexpr int sx = cf_mmd_smem<D>::sx;
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int xs_bytes = CF_GD_TI * sx * 8;
    static constexpr int cc_bytes = 2 * CF_GD_TJ * CF_GD_SC * 8;  // Cc[0..31] = ca, Cc[32..63] = -cw
    static constexpr int y_bytes = CF_GD_TJ * sx * 8;
    static constexpr int n_bytes = CF_GD_TJ * 8;
    static constexpr int stage_bytes = ((2 * y_bytes + 3 * n_bytes + 127) / 128) * 128;  // y | a | yn | q | a0 (value weights, VG)
    static constexpr int total = tbl_bytes + bar_bytes + xs_bytes + cc_bytes + CF_GD_NS * stage_bytes;
};

// q[j] = y_j . a_j over the padded rows
static __global__ void cf_rowdot_kernel(const double* __restrict__ Y, const double* __restrict__ A, int sx, int64_t m, double* __restrict__ q) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < sx; c++) s = fma(Y[j * sx + c], A[j * sx + c], s);
        q[j] = s;
    }
}

// VG = true: ValueGradientKernel, blocks (d+1) x (d+1) with entry 0 the value observation (reference src/gradient.jl:442-463):
//   b_g += ca a_g + cw r with cw = -4 k2 (r.a_g) + 2 k1 a_0,   b_0 += k a_0 - 2 k1 (r.a_g)   -- two more per-entry FMAs and a row sum
// MODE = CF_GRAD_DOT: DotProductInput programs (reference src/gradient.jl:109-115): the variable is t = x.y (GEMM 1 as it is, no
//   norms, nothing to cancel), s = x_i.a_g (GEMM 2 as it is), b_g += k1 a_g + (k2 s + k1 a_0) y,  b_0 += k a_0 + k1 s
//   -> [Ca | Cw] . [A; Y] with ca = k1, cw = k2 s (+ k1 a_0), no row-sum term.
template <int D, int KIND, bool VG, int MODE = CF_GRAD_ISO>
__global__ void __launch_bounds__(256, 1) grad_mvm_dmma_kernel(const __grid_constant__ cf_gradd_params PP) {
    using S = cf_gd_smem<D>;
    constexpr int SX = S::sx, SC = CF_GD_SC, NTB = 256, TJ = CF_GD_TJ, TI = CF_GD_TI, NS = CF_GD_NS, NCB = (D + 7) / 8;
    const cf_grad_params& P = PP.g;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    double* Xs = reinterpret_cast<double*>(smem + S::tbl_bytes + S::bar_bytes);
    double* Cc = reinterpret_cast<double*>(smem + S::tbl_bytes + S::bar_bytes + S::xs_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes + S::xs_bytes + S::cc_bytes;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t4 = lane & 3;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);

    const int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    const int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    const int nfull = (int)((c1 - c0) / TJ);
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;

    cf_fill_exp_table(tbl, P.exp2_tbl, tid, NTB);
    if (tid == 0) {
        for (int s = 0; s < NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    cf_tbl_publish(tbl_lane);
    auto issue = [&](int tile) {
        const int s = tile % NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = c0 + (int64_t)tile * TJ;
        cf_mbar_expect_tx(&bars[s], (uint32_t)(2 * S::y_bytes + (VG ? 3 : 2) * S::n_bytes));
        cf_tma_load_1d(st, P.Y + j0 * SX, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, P.a + j0 * SX, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + 2 * S::y_bytes, PP.yn + j0, (uint32_t)S::n_bytes, &bars[s]);
        cf_tma_load_1d(st + 2 * S::y_bytes + S::n_bytes, PP.q + j0, (uint32_t)S::n_bytes, &bars[s]);
        if (VG) cf_tma_load_1d(st + 2 * S::y_bytes + 2 * S::n_bytes, P.a0 + j0, (uint32_t)S::n_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * TI;
    const int64_t rend = P.row0 + P.nrows;
    for (int q = tid; q < TI * SX; q += NTB) {  // the row tile's points (rows past the end: clamped, never stored)
        const int row = q / SX;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        Xs[q] = P.X[ir * SX + (q - row * SX)];
    }
    double xnorm[2], cwsum[2] = {0.0, 0.0}, b0sum[2] = {0.0, 0.0};  // this lane's rows in both phases: 16 w + 8 rb + g
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        int64_t i = rbase + 16 * w + 8 * rb + g;
        if (i >= rend) i = rend - 1;
        xnorm[rb] = PP.xn[i];
    }
    double out[2][NCB][2];  // output fragments: rows 16 w + 8 rb + g, coordinates 8 cb + 2 t4 + e
#pragma unroll
    for (int rb = 0; rb < 2; rb++)
#pragma unroll
        for (int cb = 0; cb < NCB; cb++) out[rb][cb][0] = out[rb][cb][1] = 0.0;
    __syncthreads();

    const double eq_ca = -2.0 * P.atom.v.e.c, eq_cw = -4.0 * P.atom.v.e.c * P.atom.v.e.c;  // EQ: ca = -2 c k, cw = -4 c^2 k (r.a)
    auto tile_compute = [&](const double* __restrict__ ys, const double* __restrict__ as, const double* __restrict__ yns,
                            const double* __restrict__ qs, const double* __restrict__ a0s, int cnt, const bool ragged) {
        {   // phase A: Dot = Xs . Ys^T and Pa = Xs . As^T share the X fragments
            double c[2][4][2], p[2][4][2];
#pragma unroll
            for (int rb = 0; rb < 2; rb++)
#pragma unroll
                for (int cb = 0; cb < 4; cb++) c[rb][cb][0] = c[rb][cb][1] = p[rb][cb][0] = p[rb][cb][1] = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < D; k0 += 4) {
                double a[2], by[4], ba[4];
#pragma unroll
                for (int rb = 0; rb < 2; rb++) a[rb] = Xs[(16 * w + 8 * rb + g) * SX + k0 + t4];
#pragma unroll
                for (int cb = 0; cb < 4; cb++) {
                    by[cb] = ys[(8 * cb + g) * SX + k0 + t4];
                    ba[cb] = as[(8 * cb + g) * SX + k0 + t4];
                }
#pragma unroll
                for (int rb = 0; rb < 2; rb++)
#pragma unroll
                    for (int cb = 0; cb < 4; cb++) {
                        cf_dmma884(c[rb][cb], a[rb], by[cb]);
                        cf_dmma884(p[rb][cb], a[rb], ba[cb]);
                    }
            }
            double yn8[8], q8[8], a08[VG ? 8 : 1];  // this lane's columns: 8 cb + 2 t4 + e
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const double2 v = *reinterpret_cast<const double2*>(&yns[8 * cb + 2 * t4]);
                const double2 z = *reinterpret_cast<const double2*>(&qs[8 * cb + 2 * t4]);
                yn8[2 * cb] = v.x; yn8[2 * cb + 1] = v.y;
                q8[2 * cb] = z.x; q8[2 * cb + 1] = z.y;
                if constexpr (VG) {
                    const double2 z0 = *reinterpret_cast<const double2*>(&a0s[8 * cb + 2 * t4]);
                    a08[2 * cb] = z0.x; a08[2 * cb + 1] = z0.y;
                }
            }
            if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
                for (int rb = 0; rb < 2; rb++) {
                    const int row = 16 * w + 8 * rb + g;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                        const double v = fma(-2.0, c[rb][u >> 1][u & 1], xnorm[rb] + yn8[u]);
                        const double r2 = (__double2hiint(v) < 0) ? 0.0 : v;
                        const double sdot = p[rb][u >> 1][u & 1] - q8[u];  // r . a_j
                        // k = exp(c r2), k1 = c k, k2 = c^2 k: constants folded
                        const double kval = cf_exp_cv(r2, P.atom.v.e, tbl_lane);
                        double ca = eq_ca * kval;
                        double cw = (eq_cw * kval) * sdot;
                        if constexpr (VG) {  // ca = -2 k1 also multiplies the value weight into cw and r.a_g into the value row
                            double v0 = fma(kval, a08[u], ca * sdot);
                            if (ragged && col >= cnt) v0 = 0.0;
                            b0sum[rb] += v0;
                            cw = fma(-ca, a08[u], cw);
                        }
                        if (ragged && col >= cnt) { ca = 0.0; cw = 0.0; }  // past the end of a ragged tile: no contribution
                        cwsum[rb] += cw;
                        Cc[col * SC + row] = ca;
                        Cc[(TJ + col) * SC + row] = -cw;
                    }
                }
            } else if constexpr (KIND == CF_ATOM_MATERN) {  // one MaternP(p >= 2) atom: 8-wide jets per row block
#pragma unroll
                for (int rb = 0; rb < 2; rb++) {
                    const int row = 16 * w + 8 * rb + g;
                    double r2[8], kv[8], k1[8], k2[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const double v = fma(-2.0, c[rb][u >> 1][u & 1], xnorm[rb] + yn8[u]);
                        r2[u] = (__double2hiint(v) < 0) ? 0.0 : v;
                    }
                    cf_matern_jet_n<8>(r2, P.atom, tbl_lane, kv, k1, k2);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                        const double sdot = p[rb][u >> 1][u & 1] - q8[u];
                        double ca = -2.0 * k1[u];
                        double cw = -4.0 * k2[u] * sdot;
                        if constexpr (VG) {
                            double v0 = fma(kv[u], a08[u], ca * sdot);
                            if (ragged && col >= cnt) v0 = 0.0;
                            b0sum[rb] += v0;
                            cw = fma(-ca, a08[u], cw);
                        }
                        if (ragged && col >= cnt) { ca = 0.0; cw = 0.0; }
                        cwsum[rb] += cw;
                        Cc[col * SC + row] = ca;
                        Cc[(TJ + col) * SC + row] = -cw;
                    }
                }
            } else {
                // generic programs: 8-wide jets (the program is decoded once per 8 entries); the two row blocks go through ONE
                // copy of the code in a rolled loop -- r2 and r.a_g are parked in the entries' own slots of the coefficient tile
                // (nobody else touches them before the barrier) and overwritten by ca and -cw.  Sixteen inlined scalar copies of
                // the jet interpreter thrashed the instruction cache (1.8x slower than the scalar kernel).
#pragma unroll
                for (int rb = 0; rb < 2; rb++) {
                    const int row = 16 * w + 8 * rb + g;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                        if constexpr (MODE == CF_GRAD_DOT) {
                            Cc[col * SC + row] = c[rb][u >> 1][u & 1];
                            Cc[(TJ + col) * SC + row] = p[rb][u >> 1][u & 1];
                        } else {
                            const double v = fma(-2.0, c[rb][u >> 1][u & 1], xnorm[rb] + yn8[u]);
                            Cc[col * SC + row] = (__double2hiint(v) < 0) ? 0.0 : v;
                            Cc[(TJ + col) * SC + row] = p[rb][u >> 1][u & 1] - q8[u];
                        }
                    }
                }
#pragma unroll 1
                for (int rb = 0; rb < 2; rb++) {
                    const int row = 16 * w + 8 * rb + g;
                    double r2[8], kv[8], k1[8], k2[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) r2[u] = Cc[(8 * (u >> 1) + 2 * t4 + (u & 1)) * SC + row];
                    if (P.single) cf_atom_jet_n<8>(r2, P.atom, tbl_lane, kv, k1, k2);
                    else cf_sop_jet_n<8>(r2, P.sop, tbl_lane, kv, k1, k2);
                    double cws = 0.0, b0s = 0.0;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int col = 8 * (u >> 1) + 2 * t4 + (u & 1);
                        const double sdot = Cc[(TJ + col) * SC + row];
                        double ca, cw;
                        if constexpr (MODE == CF_GRAD_DOT) { ca = k1[u]; cw = k2[u] * sdot; }
                        else { ca = -2.0 * k1[u]; cw = -4.0 * k2[u] * sdot; }
                        if constexpr (VG) {
                            double v0 = fma(kv[u], a08[u], ca * sdot);
                            if (ragged && col >= cnt) v0 = 0.0;
                            b0s += v0;
                            cw = (MODE == CF_GRAD_DOT) ? fma(ca, a08[u], cw) : fma(-ca, a08[u], cw);
                        }
                        if (ragged && col >= cnt) { ca = 0.0; cw = 0.0; }
                        cws += cw;
                        Cc[col * SC + row] = ca;
                        Cc[(TJ + col) * SC + row] = (MODE == CF_GRAD_DOT) ? cw : -cw;
                    }
                    if (rb == 0) { cwsum[0] += cws; b0sum[0] += b0s; }
                    else { cwsum[1] += cws; b0sum[1] += b0s; }
                }
            }
        }
        __syncthreads();
        // phase B: out (128 x D) += [Ca | -Cw] (128 x 64) . [As; Ys] (64 x D)
#pragma unroll
        for (int k0 = 0; k0 < 2 * TJ; k0 += 4) {
            const double* __restrict__ bsrc = (k0 < TJ) ? as + (k0 + t4) * SX : ys + (k0 - TJ + t4) * SX;
            double a[2], b[NCB];
#pragma unroll
            for (int rb = 0; rb < 2; rb++) a[rb] = Cc[(k0 + t4) * SC + 16 * w + 8 * rb + g];
#pragma unroll
            for (int cb = 0; cb < NCB; cb++) b[cb] = bsrc[8 * cb + g];
#pragma unroll
            for (int rb = 0; rb < 2; rb++)
#pragma unroll
                for (int cb = 0; cb < NCB; cb++) cf_dmma884(out[rb][cb], a[rb], b[cb]);
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        tile_compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                     reinterpret_cast<const double*>(st + 2 * S::y_bytes),
                     reinterpret_cast<const double*>(st + 2 * S::y_bytes + S::n_bytes),
                     reinterpret_cast<const double*>(st + 2 * S::y_bytes + 2 * S::n_bytes), TJ, false);
        __syncthreads();  // Cc and stage s are free again
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {  // ragged tail: cooperative loads, zero fill
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes);
        double* yns = reinterpret_cast<double*>(stages + 2 * S::y_bytes);
        double* qs = reinterpret_cast<double*>(stages + 2 * S::y_bytes + S::n_bytes);
        double* a0s = reinterpret_cast<double*>(stages + 2 * S::y_bytes + 2 * S::n_bytes);
        __syncthreads();
        for (int q = tid; q < TJ * SX; q += NTB) {
            ys[q] = (q < cnt * SX) ? P.Y[j0 * SX + q] : 0.0;
            as[q] = (q < cnt * SX) ? P.a[j0 * SX + q] : 0.0;
        }
        for (int q = tid; q < TJ; q += NTB) {
            yns[q] = (q < cnt) ? PP.yn[j0 + q] : 0.0;
            qs[q] = (q < cnt) ? PP.q[j0 + q] : 0.0;
            if (VG) a0s[q] = (q < cnt) ? P.a0[j0 + q] : 0.0;
        }
        __syncthreads();
        tile_compute(ys, as, yns, qs, a0s, cnt, true);
        __syncthreads();
    }

    // b_i += x_i sum_j cw_ij: the four lanes of a quad hold partial sums of the same rows
    const double coef = P.single ? P.coef : 1.0;
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        cwsum[rb] += cf_shfl_xor_f64(cwsum[rb], 1);
        cwsum[rb] += cf_shfl_xor_f64(cwsum[rb], 2);
        if constexpr (VG) {
            b0sum[rb] += cf_shfl_xor_f64(b0sum[rb], 1);
            b0sum[rb] += cf_shfl_xor_f64(b0sum[rb], 2);
        }
        const int row = 16 * w + 8 * rb + g;
        const int64_t i = rbase + row;
        if (i >= rend) continue;
        double* o = P.partial + ((int64_t)blockIdx.y * P.nrows + (i - P.row0)) * D;
#pragma unroll
        for (int cb = 0; cb < NCB; cb++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int cidx = 8 * cb + 2 * t4 + e;
                if (cidx < D) o[cidx] = (MODE == CF_GRAD_DOT) ? coef * out[rb][cb][e] : coef * fma(Xs[row * SX + cidx], cwsum[rb], out[rb][cb][e]);
            }
        if (VG && t4 == 0) P.partial0[(int64_t)blockIdx.y * P.nrows + (i - P.row0)] = coef * b0sum[rb];
    }
}

#ifndef __CUDACC_RTC__ // host side
typedef cudaError_t (*cf_gradd_launch_fn)(const cf_gradd_params& P, dim3 grid, cudaStream_t stream);
template <int D, int KIND, bool VG, int MODE = CF_GRAD_ISO>
cudaError_t cf_gradd_launch(const cf_gradd_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_gd_smem<D>;
    auto kern = grad_mvm_dmma_kernel<D, KIND, VG, MODE>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<grid, 256, S::total, stream>>>(P);
    return cudaGetLastError();
}

// registry hook: padded dimensions >= 8 that are multiples of 4 (D = 12: the second 8-wide output block is half padding)
template <int D, bool OK = (D >= 8 && D % 4 == 0)>
struct cf_gradd_entry {
    // [value_gradient][0 EQ specialised, 1 generic isotropic, 2 single MaternP(p >= 2), 3 dot-product programs]
    static constexpr cf_gradd_launch_fn fn[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    static constexpr cf_mvm_config cfg = {CF_GD_TI, CF_GD_TJ, 0, 1};
};
template <int D>
struct cf_gradd_entry<D, true> {
    static constexpr cf_gradd_launch_fn fn[2][4] = {
        {&cf_gradd_launch<D, CF_ATOM_EQ, false>, &cf_gradd_launch<D, CF_ATOM_SOP, false>, &cf_gradd_launch<D, CF_ATOM_MATERN, false>,
         &cf_gradd_launch<D, CF_ATOM_SOP, false, CF_GRAD_DOT>},
        {&cf_gradd_launch<D, CF_ATOM_EQ, true>, &cf_gradd_launch<D, CF_ATOM_SOP, true>, &cf_gradd_launch<D, CF_ATOM_MATERN, true>,
         &cf_gradd_launch<D, CF_ATOM_SOP, true, CF_GRAD_DOT>}};
    static constexpr cf_mvm_config cfg = {CF_GD_TI, CF_GD_TJ, cf_gd_smem<D>::total, 1};
};
#endif // !__CUDACC_RTC__
