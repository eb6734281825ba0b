// grad_mvm.cuh -- K5: isotropic GradientKernel O(n^2 d) matrix-vector product on sm_100a (FP64).
//
// Replaces  blockmul!(y, G::Gramian, x, alpha, beta)  (reference src/gramian.jl:241-253) with the lazy
// IsotropicGradientKernelElement product (reference src/gradient.jl:86-92):
//     r = x_i - y_j;  r2 = |r|^2;  (k1, k2) = (k'(r2), k''(r2));   b_i += -2 (k1 a_j + 2 k2 r (r.a_j))
// k', k'' come from closed forms per base kernel (cf_math.cuh) instead of nested ForwardDiff duals
// (reference src/gradient.jl:589-600).  Vectors are flat, entry i*d + c (BlockFactorization isstrided).
//
// Same skeleton as gram_mvm.cuh: a CTA owns NT*R points x_i with their d-vector accumulators in registers and
// streams (y_j, a_j) tiles through a TMA/mbarrier ring; all shared-memory reads in the inner loop are
// broadcasts.  Per block: 5 D + ~15 FP64 issue slots.  Output always goes through the partial-sum buffer;
// grad_reduce_partials applies alpha / beta and removes the padding of d to the template D.
#pragma once
#include "gram_mvm.cuh"

struct cf_grad_params {
    const double* X;   // padded AoS, stride D
    const double* Y;
    const double* a;   // padded, m x D
    double* partial;   // [chunks][nrows * D]
    const double* exp2_tbl;
    int64_t row0, nrows, m, cols_per_chunk;
    int single;
    cf_atom atom;      // single atom
    cf_sop_grad sop;   // composite isotropic kernels (!single)
    double coef;       // leading constant of a single-atom program
};

template <int D, int TJ, int NS>
struct cf_grad_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int y_bytes = TJ * D * 8;
    static constexpr int stage_bytes = 2 * y_bytes;
    static constexpr int total = tbl_bytes + bar_bytes + NS * stage_bytes;
};

template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) grad_mvm_kernel(const __grid_constant__ cf_grad_params P) {
    using S = cf_grad_smem<D, TJ, NS>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;
    const int tid = threadIdx.x;
    const cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);

    const int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    const int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    const int nfull = (int)((c1 - c0) / TJ);
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;

    cf_fill_exp_table(tbl, P.exp2_tbl, tid, NT);
    if (tid == 0) {
        for (int s = 0; s < NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int tile) {
        const int s = tile % NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = c0 + (int64_t)tile * TJ;
        cf_mbar_expect_tx(&bars[s], (uint32_t)(2 * S::y_bytes));
        cf_tma_load_1d(st, P.Y + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, P.a + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    double x[R][D], b[R][D];
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * (NT * R);
    const int64_t rend = P.row0 + P.nrows;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        if (i >= rend) i = rend - 1;
#pragma unroll
        for (int c = 0; c < D; c++) { x[r][c] = P.X[i * D + c]; b[r][c] = 0.0; }
    }

    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ as, int cnt) {
        for (int j = 0; j < cnt; j++) {
            const double* __restrict__ yj = ys + j * D;
            const double* __restrict__ aj = as + j * D;
#pragma unroll
            for (int r = 0; r < R; r++) {
                // r2 and r.a with NP independent partial sums each (FP64 latency would otherwise serialise 2 D FMAs)
                constexpr int NP = (D >= 8) ? 4 : ((D >= 2) ? 2 : 1);
                double r2p[NP], drp[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) { r2p[q] = 0.0; drp[q] = 0.0; }
                double rr[D];
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const double df = x[r][c] - yj[c];
                    rr[c] = df;
                    r2p[c % NP] = fma(df, df, r2p[c % NP]);
                    drp[c % NP] = fma(df, aj[c], drp[c % NP]);
                }
                double r2 = r2p[0], dra = drp[0];
#pragma unroll
                for (int q = 1; q < NP; q++) { r2 += r2p[q]; dra += drp[q]; }
                double k, k1, k2;
                if constexpr (KIND == CF_ATOM_EQ) cf_atom_jet_t<CF_ATOM_EQ>(r2, P.atom, tbl_lane, k, k1, k2);
                else if (P.single) cf_atom_jet(r2, P.atom, tbl_lane, k, k1, k2);
                else cf_sop_jet(r2, P.sop, tbl_lane, k, k1, k2);
                const double ca = -2.0 * k1, cr = -4.0 * k2 * dra;
#pragma unroll
                for (int c = 0; c < D; c++) {
                    b[r][c] = fma(cr, rr[c], fma(ca, aj[c], b[r][c]));
                }
            }
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes), TJ);
        __syncthreads();
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes);
        __syncthreads();
        for (int q = tid; q < cnt * D; q += NT) { ys[q] = P.Y[j0 * D + q]; as[q] = P.a[j0 * D + q]; }
        __syncthreads();
        compute(ys, as, cnt);
    }

    const double coef = P.single ? P.coef : 1.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < rend) {
            double* o = P.partial + ((int64_t)blockIdx.y * P.nrows + (i - P.row0)) * D;
#pragma unroll
            for (int c = 0; c < D; c++) o[c] = coef * b[r][c];
        }
    }
}

// y[i*d + c] = alpha * sum_s partial[s][i*D + c] + beta * y[i*d + c]   (reference src/gramian.jl:245)
static __global__ void grad_reduce_partials(const double* __restrict__ partial, int chunks, int64_t nrows, int D, int d,
                                     double* __restrict__ y, const double* __restrict__ yin, int64_t ldy_unused, double alpha, double beta) {
    const int64_t total = nrows * d;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q / d;
        const int c = (int)(q - i * d);
        double s = 0.0;
        for (int ch = 0; ch < chunks; ch++) s += partial[((int64_t)ch * nrows + i) * D + c];
        double v = alpha * s;
        if (beta != 0.0) v += beta * yin[q];
        y[q] = v;
    }
}

// dst[i*D + c] = c < d ? src[i*lds + c] : 0
template <typename T>
__global__ void cf_pad_points(const T* __restrict__ src, int64_t lds, int d, T* __restrict__ dst, int D, int64_t n) {
    const int64_t total = n * D;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q / D;
        const int c = (int)(q - i * D);
        dst[q] = (c < d) ? src[i * lds + c] : (T)0;
    }
}

typedef cudaError_t (*cf_grad_launch_fn)(const cf_grad_params& P, dim3 grid, cudaStream_t stream);

template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_grad_launch(const cf_grad_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_grad_smem<D, TJ, NS>;
    auto kern = grad_mvm_kernel<D, KIND, R, NT, TJ, NS, MINB>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<grid, NT, S::total, stream>>>(P);
    return cudaGetLastError();
}
