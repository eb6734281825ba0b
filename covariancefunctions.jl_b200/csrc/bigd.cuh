// bigd.cuh -- point dimension d > 32 (any d): the per-pair reductions over d no longer fit a thread's registers, so the
// products are organised as tiled contractions over the coordinate index (GEMM-shaped loops with the reference's own
// inner operations -- direct differences, not the norm trick):
//   bigd_pair_kernel      T[i][j] = |x_i - y_j|^2 (ISO) or x_i . y_j (DOT),   S[i][j] = w . a_j  (derivative operators: the
//                         reference's r'a / x'a, src/gradient.jl:90,113)  or  x_i . y_j (value kernels that need both)
//   bigd_value_kernel     y_i = alpha sum_j k(T_ij, S_ij) a_j + beta y_i                      (src/gramian.jl:78-87)
//   bigd_jet_kernel       (T, S) -> (ca, cw) in place: ISO ca = -2 k', cw = -4 k'' S;  DOT ca = k', cw = k'' S
//   bigd_jet_vg_kernel    the same for the ValueGradientKernel ((d + 1)-blocks, entry 0 = value: src/gradient.jl:400-474): the value weight
//                         a0_j adds 2 k' a0_j (ISO) / k' a0_j (DOT) to cw, and the value row b0_i = sum_j k a0_j + ca S is reduced per row
//   bigd_update_kernel    b_i[c] = sum_j ca_ij a_j[c] + cw_ij w_ij[c],  w = x_i - y_j (ISO) or y_j (DOT)   (src/gradient.jl:91,114)
// The host walks row blocks so that the two [rows][m] scratch matrices stay bounded.  Points are padded to a multiple of
// 16 coordinates.  Float64 only (Float32 handles: Float64 shadow).  64 x 64 output tiles, 16-deep chunks staged through shared memory, 4 x 4 register tiles.
#pragma once
#include "grad_mvm.cuh"

#define CF_BD_T 64   /* tile edge */
#define CF_BD_K 16   /* chunk depth */

// MODE: CF_GRAD_ISO / CF_GRAD_DOT.  WITH_A: S = w . a_j (a is m x D); otherwise S = x . y.
template <int MODE, bool WITH_A>
__global__ void __launch_bounds__(256) bigd_pair_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                        const double* __restrict__ A, int D, int64_t i0, int64_t nrows, int64_t m,
                                                        double* __restrict__ Tm, double* __restrict__ Sm) {
    // rows padded by 2 doubles: the transposing stores of the tile loads hit 8 banks instead of 1, loads stay 16-byte aligned
    __shared__ __align__(16) double xs[CF_BD_K][CF_BD_T + 2], ys[CF_BD_K][CF_BD_T + 2], as[CF_BD_K][CF_BD_T + 2];
    const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    const int64_t ib = (int64_t)blockIdx.y * CF_BD_T, jb = (int64_t)blockIdx.x * CF_BD_T;
    double t[4][4], s[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { t[a][b] = 0.0; s[a][b] = 0.0; }
    for (int c0 = 0; c0 < D; c0 += CF_BD_K) {
        __syncthreads();
        for (int q = tid; q < CF_BD_T * CF_BD_K; q += 256) {  // consecutive threads read consecutive coordinates of one point
            const int p = q / CF_BD_K, c = q - p * CF_BD_K;
            const int64_t i = ib + p, j = jb + p;
            xs[c][p] = (i < nrows) ? X[(i0 + i) * D + c0 + c] : 0.0;
            ys[c][p] = (j < m) ? Y[j * D + c0 + c] : 0.0;
            if (WITH_A) as[c][p] = (j < m) ? A[j * D + c0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CF_BD_K; c++) {
            const double2 x01 = *reinterpret_cast<const double2*>(&xs[c][4 * ti]), x23 = *reinterpret_cast<const double2*>(&xs[c][4 * ti + 2]);
            const double2 y01 = *reinterpret_cast<const double2*>(&ys[c][4 * tj]), y23 = *reinterpret_cast<const double2*>(&ys[c][4 * tj + 2]);
            const double xv[4] = {x01.x, x01.y, x23.x, x23.y}, yv[4] = {y01.x, y01.y, y23.x, y23.y};
            double av[4] = {0, 0, 0, 0};
            if (WITH_A) {
                const double2 a01 = *reinterpret_cast<const double2*>(&as[c][4 * tj]), a23 = *reinterpret_cast<const double2*>(&as[c][4 * tj + 2]);
                av[0] = a01.x; av[1] = a01.y; av[2] = a23.x; av[3] = a23.y;
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (MODE == CF_GRAD_ISO) {
                        const double df = xv[a] - yv[b];
                        t[a][b] = fma(df, df, t[a][b]);
                        s[a][b] = WITH_A ? fma(df, av[b], s[a][b]) : fma(xv[a], yv[b], s[a][b]);
                    } else {
                        t[a][b] = fma(xv[a], yv[b], t[a][b]);
                        if (WITH_A) s[a][b] = fma(xv[a], av[b], s[a][b]);
                    }
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int64_t i = ib + 4 * ti + a;
        if (i >= nrows) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t j = jb + 4 * tj + b;
            if (j >= m) continue;
            Tm[i * m + j] = t[a][b];
            Sm[i * m + j] = (MODE == CF_GRAD_DOT && !WITH_A) ? t[a][b] : s[a][b];
        }
    }
}

// y_i = alpha sum_j k(r2_ij, dot_ij) a_j + beta y_i; one CTA per row of the block (T holds r2, S holds x.y)
static __global__ void __launch_bounds__(256) bigd_value_kernel(const double* __restrict__ Tm, const double* __restrict__ Sm,
                                                                const double* __restrict__ a, int64_t nrows, int64_t m,
                                                                const __grid_constant__ cf_sop_val prog,
                                                                const double* __restrict__ exp2_tbl, double* __restrict__ y,
                                                                const double* __restrict__ yin, double alpha, double beta) {
    extern __shared__ __align__(128) unsigned char bd_smem[];
    double* tbl = reinterpret_cast<double*>(bd_smem);
    __shared__ double red[8];
    cf_fill_exp_table(tbl, exp2_tbl, threadIdx.x, 256);
    __syncthreads();
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, threadIdx.x);
    cf_tbl_publish(tbl_lane);
    for (int64_t i = blockIdx.x; i < nrows; i += gridDim.x) {
        double acc = 0.0;
        for (int64_t j = threadIdx.x; j < m; j += 256)
            acc = fma(cf_sop_value(Tm[i * m + j], Sm[i * m + j], prog, tbl_lane), a[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += cf_shfl_xor_f64(acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < 8; w++) v += red[w];
            v *= alpha;
            if (beta != 0.0) v += beta * yin[i];
            y[i] = v;
        }
        __syncthreads();
    }
}

// (T, S) -> (ca, cw) in place
template <int MODE>
__global__ void __launch_bounds__(256) bigd_jet_kernel(double* __restrict__ Tm, double* __restrict__ Sm, int64_t total,
                                                       const __grid_constant__ cf_sop_grad prog, const double* __restrict__ exp2_tbl) {
    extern __shared__ __align__(128) unsigned char bd_smem[];
    double* tbl = reinterpret_cast<double*>(bd_smem);
    cf_fill_exp_table(tbl, exp2_tbl, threadIdx.x, 256);
    __syncthreads();
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, threadIdx.x);
    cf_tbl_publish(tbl_lane);
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total; q += (int64_t)gridDim.x * 256) {
        double k, k1, k2;
        cf_sop_jet(Tm[q], prog, tbl_lane, k, k1, k2);
        const double sd = Sm[q];
        if (MODE == CF_GRAD_ISO) { Tm[q] = -2.0 * k1; Sm[q] = -4.0 * k2 * sd; }
        else { Tm[q] = k1; Sm[q] = k2 * sd; }
    }
}

// ValueGradientKernel: one CTA per row of the block.  (T, S) -> (ca, cw) in place with the value weight's share in cw, and
// out0[i * ostride] = alpha sum_j (k a0_j + ca S_ij) + beta yin0[i * ostride]   (the value entry of row block i; fixed reduction order)
template <int MODE>
__global__ void __launch_bounds__(256) bigd_jet_vg_kernel(double* __restrict__ Tm, double* __restrict__ Sm, const double* __restrict__ a0,
                                                          int64_t nrows, int64_t m, const __grid_constant__ cf_sop_grad prog,
                                                          const double* __restrict__ exp2_tbl, double* __restrict__ out0,
                                                          const double* __restrict__ yin0, int64_t ostride, double alpha, double beta) {
    extern __shared__ __align__(128) unsigned char bd_smem[];
    double* tbl = reinterpret_cast<double*>(bd_smem);
    __shared__ double red[8];
    cf_fill_exp_table(tbl, exp2_tbl, threadIdx.x, 256);
    __syncthreads();
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, threadIdx.x);
    cf_tbl_publish(tbl_lane);
    for (int64_t i = blockIdx.x; i < nrows; i += gridDim.x) {
        double acc = 0.0;
        for (int64_t j = threadIdx.x; j < m; j += 256) {
            const int64_t q = i * m + j;
            double k, k1, k2;
            cf_sop_jet(Tm[q], prog, tbl_lane, k, k1, k2);
            const double sd = Sm[q], aj = a0[j];
            const double ca = (MODE == CF_GRAD_ISO) ? -2.0 * k1 : k1;
            Tm[q] = ca;
            Sm[q] = (MODE == CF_GRAD_ISO) ? fma(-4.0 * k2, sd, 2.0 * k1 * aj) : fma(k2, sd, k1 * aj);
            acc = fma(k, aj, fma(ca, sd, acc));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += cf_shfl_xor_f64(acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < 8; w++) v += red[w];
            v *= alpha;
            if (beta != 0.0) v += beta * yin0[i * ostride];
            out0[i * ostride] = v;
        }
        __syncthreads();
    }
}

// out[i * ostride + c] = alpha sum_j (ca_ij a_j[c] + cw_ij w_ij[c]) + beta yin[i * ostride + c];  out / yin are unpadded flat vectors with
// blocks of ostride entries (d for the GradientKernel; d + 1 for the ValueGradientKernel, whose callers pass pointers to entry 1)
template <int MODE>
__global__ void __launch_bounds__(256) bigd_update_kernel(const double* __restrict__ X, const double* __restrict__ Y,
                                                          const double* __restrict__ A, int D, int d, int64_t i0, int64_t nrows,
                                                          int64_t m, const double* __restrict__ CA, const double* __restrict__ CW,
                                                          double* __restrict__ out, const double* __restrict__ yin, double alpha,
                                                          double beta, int64_t ostride) {
    __shared__ __align__(16) double cas[CF_BD_K][CF_BD_T + 2], cws[CF_BD_K][CF_BD_T + 2], as[CF_BD_K][CF_BD_T], ys[CF_BD_K][CF_BD_T];
    const int tid = threadIdx.x, ti = tid & 15, tc = tid >> 4;
    const int64_t ib = (int64_t)blockIdx.y * CF_BD_T;
    const int cb = blockIdx.x * CF_BD_T;
    double xr[4][4], acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int64_t i = ib + 4 * ti + a;
            const int c = cb + 4 * tc + b;
            xr[a][b] = (MODE == CF_GRAD_ISO && i < nrows && c < D) ? X[(i0 + i) * D + c] : 0.0;
            acc[a][b] = 0.0;
        }
    for (int64_t j0 = 0; j0 < m; j0 += CF_BD_K) {
        __syncthreads();
        for (int q = tid; q < CF_BD_T * CF_BD_K; q += 256) {
            {   // coefficient tiles: consecutive threads read consecutive j of one row
                const int p = q / CF_BD_K, jj = q - p * CF_BD_K;
                const int64_t i = ib + p, j = j0 + jj;
                const bool ok = (i < nrows) && (j < m);
                cas[jj][p] = ok ? CA[i * m + j] : 0.0;
                cws[jj][p] = ok ? CW[i * m + j] : 0.0;
            }
            {   // point tiles: consecutive threads read consecutive coordinates of one column point
                const int jj = q / CF_BD_T, c = q - jj * CF_BD_T;
                const int64_t j = j0 + jj;
                const bool ok = (j < m) && (cb + c < D);
                as[jj][c] = ok ? A[j * D + cb + c] : 0.0;
                ys[jj][c] = ok ? Y[j * D + cb + c] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < CF_BD_K; jj++) {
            const double2 ca01 = *reinterpret_cast<const double2*>(&cas[jj][4 * ti]), ca23 = *reinterpret_cast<const double2*>(&cas[jj][4 * ti + 2]);
            const double2 cw01 = *reinterpret_cast<const double2*>(&cws[jj][4 * ti]), cw23 = *reinterpret_cast<const double2*>(&cws[jj][4 * ti + 2]);
            const double2 a01 = *reinterpret_cast<const double2*>(&as[jj][4 * tc]), a23 = *reinterpret_cast<const double2*>(&as[jj][4 * tc + 2]);
            const double2 y01 = *reinterpret_cast<const double2*>(&ys[jj][4 * tc]), y23 = *reinterpret_cast<const double2*>(&ys[jj][4 * tc + 2]);
            const double cav[4] = {ca01.x, ca01.y, ca23.x, ca23.y}, cwv[4] = {cw01.x, cw01.y, cw23.x, cw23.y};
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, yv[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double w = (MODE == CF_GRAD_ISO) ? (xr[a][b] - yv[b]) : yv[b];
                    acc[a][b] = fma(cwv[a], w, fma(cav[a], av[b], acc[a][b]));
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int64_t i = ib + 4 * ti + a;
        if (i >= nrows) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int c = cb + 4 * tc + b;
            if (c >= d) continue;
            const int64_t o = i * ostride + c;
            double v = alpha * acc[a][b];
            if (beta != 0.0) v += beta * yin[o];
            out[o] = v;
        }
    }
}
