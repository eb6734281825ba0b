// gram_mvm_sym.cuh -- OPTIONAL symmetric variant of K1 for y === x (K = K^T), Float64.
//
// Every unordered pair {i, j} in different row blocks is evaluated ONCE and used twice:
//     b_i += k(x_i, x_j) a_j      (row sums: registers, exactly as in gram_mvm_kernel)
//     b_j += k(x_i, x_j) a_i      (column sums: reduced across the CTA's rows and added to b with red.global.add.f64)
// which halves the exp / distance work of the symmetric Gramians (BASELINE configs 1, 2, 5).  The reference does not
// do this (src/gramian.jl:78-87 evaluates all n*m entries); results agree to rounding but the summation order of the
// column part depends on the order in which CTAs retire (floating-point atomics), so the variant is OFF by default
// (cf_gramian_set_option(g, CF_OPT_SYMMETRIC, 1) / COVFN_SYMMETRIC=1) and the bench headline does not use it.
//
// Work decomposition: row tiles of NT*R rows.  For row tile I the columns of its own row block are handled by the
// plain kernel in "diagonal block" mode; the columns beyond the block, [(I+1) NT R, n), are cut into chunks and every
// (row tile, chunk) pair is one CTA of this kernel (the host builds the item list; the triangular shape is balanced by
// having ~64 K similar-sized items).  Column sums: each thread forms, for 8 columns at a time, the R-row partial
// sums; a 5-stage transpose-reduce (xor 16, 8, 4 with halving payload, then 2, 1) leaves one column total per lane
// quad, 9 shuffle-exchanges per 8 columns instead of 40.
#pragma once
#include "gram_mvm.cuh"

struct cf_sym_item {
    int64_t col0, col1;  // column range (col0 is a multiple of TJ)
    int32_t row_tile;
    int32_t pad_;
};

struct cf_sym_params {
    const double* X;       // padded AoS, stride D (rows and columns: the same point set)
    const double* a;       // weights, length n
    double* bsym;          // accumulation target, length n, zero-initialised by the host
    const double* exp2_tbl;
    const cf_sym_item* items;
    int64_t n;
    int use_tma;
    cf_atom_val atom;
    cf_sop_val sop;
};

template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) gram_mvm_sym_kernel(const __grid_constant__ cf_sym_params P) {
    using S = cf_mvm_smem<double, D, TJ, NS>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;
    const int tid = threadIdx.x, lane = tid & 31;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const cf_sym_item it = P.items[blockIdx.x];
    const int64_t c0 = it.col0, c1 = it.col1;
    const int nfull = P.use_tma ? (int)((c1 - c0) / TJ) : 0;
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;

    cf_fill_exp_table(tbl, P.exp2_tbl, tid, NT);
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
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + S::a_bytes));
        cf_tma_load_1d(st, P.X + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, P.a + j0, (uint32_t)S::a_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    // this thread's rows and their weights (rows past the end: clamped point, zero weight, never stored)
    double x[R][D], ai[R], tot[R];
    const int64_t rbase = (int64_t)it.row_tile * (NT * R);
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        const bool ok = i < P.n;
        if (!ok) i = P.n - 1;
#pragma unroll
        for (int c = 0; c < D; c++) x[r][c] = P.X[i * D + c];
        ai[r] = ok ? P.a[i] : 0.0;
        tot[r] = 0.0;
    }

    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ as, int cnt, int64_t j0) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        for (int jb = 0; jb < TJ; jb += 8) {
            if (jb >= cnt) break;
            double cv[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int j = jb + jj;
                double yj[D];
#pragma unroll
                for (int c = 0; c < D; c++) yj[c] = ys[j * D + c];
                const double aj = as[j];
                double kv[R];
                cf_rows_value<double, D, KIND, R>(x, yj, P.atom, P.sop, tbl_lane, kv);
                double cs = kv[0] * ai[0];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    acc[r] = fma(kv[r], aj, acc[r]);
                    if (r > 0) cs = fma(kv[r], ai[r], cs);
                }
                cv[jj] = cs;
            }
            // transpose-reduce 8 columns over the 32 lanes
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
            double c4[4], c2[2];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double send = b16 ? cv[q] : cv[q + 4];
                const double keep = b16 ? cv[q + 4] : cv[q];
                c4[q] = keep + cf_shfl_xor_f64(send, 16);
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const double send = b8 ? c4[q] : c4[q + 2];
                const double keep = b8 ? c4[q + 2] : c4[q];
                c2[q] = keep + cf_shfl_xor_f64(send, 8);
            }
            double c1v;
            {
                const double send = b4 ? c2[0] : c2[1];
                const double keep = b4 ? c2[1] : c2[0];
                c1v = keep + cf_shfl_xor_f64(send, 4);
            }
            c1v += cf_shfl_xor_f64(c1v, 2);
            c1v += cf_shfl_xor_f64(c1v, 1);
            const int col = jb + (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0);
            if ((lane & 3) == 0 && col < cnt) atomicAdd(&P.bsym[j0 + col], c1v);
        }
#pragma unroll
        for (int r = 0; r < R; r++) tot[r] += acc[r];
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes), TJ, c0 + (int64_t)t * TJ);
        __syncthreads();
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes);
        __syncthreads();
        for (int q = tid; q < TJ * D; q += NT) ys[q] = (q < cnt * D) ? P.X[j0 * D + q] : 0.0;
        for (int q = tid; q < TJ; q += NT) as[q] = (q < cnt) ? P.a[j0 + q] : 0.0;
        __syncthreads();
        compute(ys, as, cnt, j0);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < P.n) atomicAdd(&P.bsym[i], tot[r]);
    }
}

// y[o] = alpha * (diag[o] + bsym[o]) + beta * y[o]
static __global__ void gram_sym_combine(const double* __restrict__ diag, const double* __restrict__ bsym, int64_t n,
                                        double* __restrict__ y, const double* __restrict__ yin, double alpha, double beta) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        double v = alpha * (diag[o] + bsym[o]);
        if (beta != 0.0) v += beta * yin[o];
        y[o] = v;
    }
}

typedef cudaError_t (*cf_sym_launch_fn)(const cf_sym_params& P, int nitems, cudaStream_t stream);
template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_sym_launch(const cf_sym_params& P, int nitems, cudaStream_t stream) {
    using S = cf_mvm_smem<double, D, TJ, NS>;
    auto kern = gram_mvm_sym_kernel<D, KIND, R, NT, TJ, NS, MINB>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<nitems, NT, S::total, stream>>>(P);
    return cudaGetLastError();
}
