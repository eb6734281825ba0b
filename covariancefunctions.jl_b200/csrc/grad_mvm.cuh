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
    const double* a;   // gradient weights, padded, m x D
    const double* a0;  // value weights, m (ValueGradientKernel only)
    double* partial;   // [chunks][nrows * D]
    double* partial0;  // [chunks][nrows]       (ValueGradientKernel only)
    const double* exp2_tbl;
    int64_t row0, nrows, m, cols_per_chunk;
    int single;
    cf_atom atom;      // single atom
    cf_sop_grad sop;   // composite kernels (!single)
    double coef;       // leading constant of a single-atom program
};

// MODE: which scalar the kernel is a function of (reference src/properties.jl:31-45)
#define CF_GRAD_ISO 0  /* IsotropicInput:  t = |x - y|^2, block = -2 (k1 I + 2 k2 r r^T)      src/gradient.jl:86-92   */
#define CF_GRAD_DOT 1  /* DotProductInput: t = x . y,     block = k1 I + k2 y x^T             src/gradient.jl:109-115 */

template <int D, int TJ, int NS, bool VG>
struct cf_grad_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int y_bytes = TJ * D * 8;
    static constexpr int a0_bytes = VG ? TJ * 8 : 0;
    static constexpr int stage_bytes = ((2 * y_bytes + a0_bytes + 127) / 128) * 128;
    static constexpr int total = tbl_bytes + bar_bytes + NS * stage_bytes;
};

// VG = false: GradientKernel (blocks d x d);  VG = true: ValueGradientKernel (blocks (d+1) x (d+1), entry 0 = value):
//   isotropic   [k0, (-2 k1 r)^T; 2 k1 r, GG]      dot product   [k0, (k1 x)^T; k1 y, GG]     (src/gradient.jl:442-463)
// Both modes reduce to   b_g += ca a_g + cw w,   b_0 += k0 a_0 + cs s   with
//   ISO: w = r = x - y, s = r.a_g, ca = -2 k1, cw = -4 k2 s + 2 k1 a_0, cs = -2 k1
//   DOT: w = y,         s = x.a_g, ca = k1,    cw = k2 s + k1 a_0,      cs = k1
template <int D, int KIND, int MODE, bool VG, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) grad_mvm_kernel(const __grid_constant__ cf_grad_params P) {
    using S = cf_grad_smem<D, TJ, NS, VG>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;
    const int tid = threadIdx.x;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);

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
    cf_tbl_publish(tbl_lane);
    auto issue = [&](int tile) {
        const int s = tile % NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = c0 + (int64_t)tile * TJ;
        cf_mbar_expect_tx(&bars[s], (uint32_t)(2 * S::y_bytes + S::a0_bytes));
        cf_tma_load_1d(st, P.Y + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, P.a + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        if (VG) cf_tma_load_1d(st + 2 * S::y_bytes, P.a0 + j0, (uint32_t)S::a0_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    double x[R][D], b[R][D], b0[R];
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * (NT * R);
    const int64_t rend = P.row0 + P.nrows;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        if (i >= rend) i = rend - 1;
        b0[r] = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) { x[r][c] = P.X[i * D + c]; b[r][c] = 0.0; }
    }

    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ as, const double* __restrict__ a0s, int cnt) {
        for (int j = 0; j < cnt; j++) {
            const double* __restrict__ yj = ys + j * D;
            const double* __restrict__ aj = as + j * D;
            const double a0 = VG ? a0s[j] : 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                // t and s with NP independent partial sums each (FP64 latency would otherwise serialise 2 D FMAs)
                constexpr int NP = (D >= 8) ? 4 : ((D >= 2) ? 2 : 1);
                double tp[NP], sp[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) { tp[q] = 0.0; sp[q] = 0.0; }
                double w[D];
#pragma unroll
                for (int c = 0; c < D; c++) {
                    if (MODE == CF_GRAD_ISO) {
                        const double df = x[r][c] - yj[c];
                        w[c] = df;
                        tp[c % NP] = fma(df, df, tp[c % NP]);
                        sp[c % NP] = fma(df, aj[c], sp[c % NP]);
                    } else {
                        w[c] = yj[c];
                        tp[c % NP] = fma(x[r][c], yj[c], tp[c % NP]);
                        sp[c % NP] = fma(x[r][c], aj[c], sp[c % NP]);
                    }
                }
                double t = tp[0], sdot = sp[0];
#pragma unroll
                for (int q = 1; q < NP; q++) { t += tp[q]; sdot += sp[q]; }
                double k, k1, k2;
                if constexpr (KIND == CF_ATOM_EQ) cf_atom_jet_t<CF_ATOM_EQ>(t, P.atom, tbl_lane, k, k1, k2);
                else if (P.single) cf_atom_jet(t, P.atom, tbl_lane, k, k1, k2);
                else cf_sop_jet(t, P.sop, tbl_lane, k, k1, k2);
                double ca, cw, cs;
                if (MODE == CF_GRAD_ISO) { ca = -2.0 * k1; cs = ca; cw = fma(-4.0 * k2, sdot, VG ? 2.0 * k1 * a0 : 0.0); }
                else { ca = k1; cs = k1; cw = fma(k2, sdot, VG ? k1 * a0 : 0.0); }
#pragma unroll
                for (int c = 0; c < D; c++) b[r][c] = fma(cw, w[c], fma(ca, aj[c], b[r][c]));
                if (VG) b0[r] = fma(k, a0, fma(cs, sdot, b0[r]));
            }
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                reinterpret_cast<const double*>(st + 2 * S::y_bytes), TJ);
        __syncthreads();
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes);
        double* a0s = reinterpret_cast<double*>(stages + 2 * S::y_bytes);
        __syncthreads();
        for (int q = tid; q < cnt * D; q += NT) { ys[q] = P.Y[j0 * D + q]; as[q] = P.a[j0 * D + q]; }
        if (VG) for (int q = tid; q < cnt; q += NT) a0s[q] = P.a0[j0 + q];
        __syncthreads();
        compute(ys, as, a0s, cnt);
    }

    const double coef = P.single ? P.coef : 1.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < rend) {
            double* o = P.partial + ((int64_t)blockIdx.y * P.nrows + (i - P.row0)) * D;
#pragma unroll
            for (int c = 0; c < D; c++) o[c] = coef * b[r][c];
            if (VG) P.partial0[(int64_t)blockIdx.y * P.nrows + (i - P.row0)] = coef * b0[r];
        }
    }
}

// y[i*bs + e] = alpha * sum_s partial[...] + beta * y[i*bs + e], bs = d + vg; entry 0 of a ValueGradient block is the
// value part (reference src/gramian.jl:245 for the beta handling)
static __global__ void grad_reduce_partials(const double* __restrict__ partial, const double* __restrict__ partial0, int chunks,
                                            int64_t nrows, int D, int d, int vg, double* __restrict__ y,
                                            const double* __restrict__ yin, double alpha, double beta, const cf_peer_out peers) {
    const int bs = d + vg;
    const int64_t total = nrows * bs;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q / bs;
        const int e = (int)(q - i * bs);
        double s = 0.0;
        if (vg && e == 0) {
            for (int ch = 0; ch < chunks; ch++) s += partial0[(int64_t)ch * nrows + i];
        } else {
            for (int ch = 0; ch < chunks; ch++) s += partial[((int64_t)ch * nrows + i) * D + (e - vg)];
        }
        double v = alpha * s;
        if (beta != 0.0) v += beta * yin[q];
        y[q] = v;
        for (int p = 0; p < peers.n; p++) static_cast<double*>(peers.ptr[p])[q] = v; // NVLink peer stores
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

#ifndef __CUDACC_RTC__ // host side: not part of run-time specialised builds
typedef cudaError_t (*cf_grad_launch_fn)(const cf_grad_params& P, dim3 grid, cudaStream_t stream);

template <int D, int KIND, int MODE, bool VG, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_grad_launch(const cf_grad_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_grad_smem<D, TJ, NS, VG>;
    auto kern = grad_mvm_kernel<D, KIND, MODE, VG, R, NT, TJ, NS, MINB>;
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
#endif // !__CUDACC_RTC__
