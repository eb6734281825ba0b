// cf_extra.cuh -- kernels beside the MVM hot loop: dense tile instantiation (Matrix!), the first multi-RHS
// product, vector helpers for on-device CG (K6), and the pipe-peak probes used by bench.py.
#pragma once
#include "gram_mvm.cuh"

// y = beta * yin (beta == 0 -> 0), the m == 0 product
template <typename T>
__global__ void cf_scale_kernel(T* __restrict__ y, const T* __restrict__ yin, int64_t n, double beta) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
        y[q] = (beta == 0.0) ? (T)0 : (T)(beta * (double)yin[q]);
}

// ---- Matrix!(M, G): M[(i - r0) + ld (j - j0)] = k(x_i, y_j)   (reference src/gramian.jl:107-114) ------------------
// one thread per entry, generic sum-of-products evaluation; not a hot path.
template <typename T>
__global__ void gram_dense_kernel(const T* __restrict__ X, const T* __restrict__ Y, int D, const cf_program* __restrict__ prog,
                                  const double* __restrict__ exp2_tbl, int64_t r0, int64_t nrows, int64_t j0, int64_t ncols,
                                  T* __restrict__ M, int64_t ld) {
    __shared__ double tbl[CF_EXP_TBL_DOUBLES];
    cf_fill_exp_table(tbl, exp2_tbl, threadIdx.x, blockDim.x);
    __syncthreads();
    const double* tbl_lane = tbl + (threadIdx.x & 15);
    const int64_t total = nrows * ncols;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jj = q / nrows, ii = q - jj * nrows;
        const T* x = X + (r0 + ii) * D;
        const T* y = Y + (j0 + jj) * D;
        T r2 = 0, dt = 0;
        for (int c = 0; c < D; c++) {
            T df = x[c] - y[c];
            r2 = fma(df, df, r2);
            dt = fma(x[c], y[c], dt);
        }
        T v;
        if constexpr (sizeof(T) == 8) v = cf_sop_value(r2, dt, prog, tbl_lane);
        else v = cf_sop_value_f32(r2, dt, prog);
        M[ii + ld * jj] = v;
    }
}

// ---- K4 (first version): B <- alpha K A + beta B with A m x p, every kernel entry evaluated ONCE -------------------
// (the reference re-evaluates it for each of the p columns: src/gramian.jl:89-99).
// CTA: TI = 64 rows, tiles of TJ = 64 columns, PC = 64 right-hand sides per pass.
//   phase A: K tile (TI x TJ) -> shared memory; thread (i = t & 63, jg = t >> 6) evaluates 16 entries of its row with
//            x_i in registers and y_j read as shared-memory broadcasts.
//   phase B: B tile (TI x PC) += K tile . A tile with 4 x 4 register tiles per thread (FP64 FMA pipe).
#define CF_MM_TI 64
#define CF_MM_TJ 64
#define CF_MM_PC 64
#define CF_MM_LDA (CF_MM_PC + 2)

template <typename T, int D>
struct cf_mm_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int ks_bytes = CF_MM_TJ * CF_MM_TI * 8;     // Ks[j][i], double
    static constexpr int as_bytes = CF_MM_TJ * CF_MM_LDA * 8;    // As[j][c], double
    static constexpr int ys_bytes = CF_MM_TJ * D * (int)sizeof(T);
    static constexpr int total = tbl_bytes + ks_bytes + as_bytes + ys_bytes;
};

struct cf_mm_params {
    const void* X; const void* Y; const void* A; void* B;
    const double* exp2_tbl; const cf_program* prog;
    int64_t row0, nrows, m, lda, ldb;
    int nrhs;      // columns in this pass (<= CF_MM_PC)
    double alpha, beta;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) gram_mm_kernel(const __grid_constant__ cf_mm_params P) {
    using S = cf_mm_smem<T, D>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    double* Ks = reinterpret_cast<double*>(smem + S::tbl_bytes);
    double* As = reinterpret_cast<double*>(smem + S::tbl_bytes + S::ks_bytes);
    T* ys = reinterpret_cast<T*>(smem + S::tbl_bytes + S::ks_bytes + S::as_bytes);
    const int tid = threadIdx.x;
    const double* tbl_lane = tbl + (tid & 15);
    const T* __restrict__ Xg = static_cast<const T*>(P.X);
    const T* __restrict__ Yg = static_cast<const T*>(P.Y);
    const T* __restrict__ Ag = static_cast<const T*>(P.A);
    cf_fill_exp_table(tbl, P.exp2_tbl, tid, 256);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * CF_MM_TI;
    const int64_t rend = P.row0 + P.nrows;
    const int li = tid & 63, jg = tid >> 6;
    T x[D];
    {
        int64_t i = rbase + li;
        if (i >= rend) i = rend - 1;
#pragma unroll
        for (int c = 0; c < D; c++) x[c] = Xg[i * D + c];
    }
    const int ri = tid & 15, ci = tid >> 4; // phase B tile: rows 4 ri.., cols 4 ci..
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;

    for (int64_t j0 = 0; j0 < P.m; j0 += CF_MM_TJ) {
        const int cnt = (int)((P.m - j0 < CF_MM_TJ) ? P.m - j0 : CF_MM_TJ);
        __syncthreads();
        for (int q = tid; q < cnt * D; q += 256) ys[q] = Yg[j0 * D + q];
        for (int q = tid; q < CF_MM_TJ * CF_MM_PC; q += 256) {
            const int c = q / CF_MM_TJ, k = q - c * CF_MM_TJ; // consecutive threads: consecutive k (coalesced in A)
            As[k * CF_MM_LDA + c] = (k < cnt && c < P.nrhs) ? (double)Ag[(j0 + k) + P.lda * c] : 0.0;
        }
        __syncthreads();
        // phase A
#pragma unroll 2
        for (int q = 0; q < 16; q++) {
            const int j = jg * 16 + q;
            double kv = 0.0;
            if (j < cnt) {
                T r2 = 0, dt = 0;
#pragma unroll
                for (int c = 0; c < D; c++) {
                    T yv = ys[j * D + c];
                    T df = x[c] - yv;
                    r2 = (c == 0) ? df * df : fma(df, df, r2);
                    dt = (c == 0) ? x[c] * yv : fma(x[c], yv, dt);
                }
                if constexpr (sizeof(T) == 8) kv = cf_sop_value(r2, dt, P.prog, tbl_lane);
                else kv = (double)cf_sop_value_f32(r2, dt, P.prog);
            }
            Ks[j * CF_MM_TI + li] = kv;
        }
        __syncthreads();
        // phase B
#pragma unroll 4
        for (int k = 0; k < CF_MM_TJ; k++) {
            const double2 k01 = *reinterpret_cast<const double2*>(&Ks[k * CF_MM_TI + 4 * ri]);
            const double2 k23 = *reinterpret_cast<const double2*>(&Ks[k * CF_MM_TI + 4 * ri + 2]);
            const double2 a01 = *reinterpret_cast<const double2*>(&As[k * CF_MM_LDA + 4 * ci]);
            const double2 a23 = *reinterpret_cast<const double2*>(&As[k * CF_MM_LDA + 4 * ci + 2]);
            const double kr[4] = {k01.x, k01.y, k23.x, k23.y};
            const double ac[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(kr[a], ac[b], acc[a][b]);
        }
    }
    T* Bg = static_cast<T*>(P.B);
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int64_t i = rbase + 4 * ri + a;
        if (i >= rend) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int c = 4 * ci + b;
            if (c >= P.nrhs) continue;
            T* o = Bg + (i - P.row0) + P.ldb * c;
            double v = P.alpha * acc[a][b];
            if (P.beta != 0.0) v += P.beta * (double)(*o);
            *o = (T)v;
        }
    }
}

typedef cudaError_t (*cf_mm_launch_fn)(const cf_mm_params& P, int row_tiles, cudaStream_t stream);
template <typename T, int D>
cudaError_t cf_mm_launch(const cf_mm_params& P, int row_tiles, cudaStream_t stream) {
    using S = cf_mm_smem<T, D>;
    auto kern = gram_mm_kernel<T, D>;
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

// ---- K6: vector helpers for conjugate gradients (Float64) -------------------------------------------------------------
// z = a x + b y
static __global__ void cf_axpby_kernel(double* __restrict__ z, double a, const double* __restrict__ x, double b,
                                const double* __restrict__ y, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
        z[q] = a * x[q] + b * y[q];
}
// deterministic dot product: one CTA, fixed summation tree.  out[0] = sum x[q] y[q]
static __global__ void __launch_bounds__(1024) cf_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                                      double* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t q = threadIdx.x; q < n; q += 1024) s = fma(x[q], y[q], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// ---- pipe peak probes ---------------------------------------------------------------------------------------------------
// dependent-chain-free FMAs: 8 independent chains per thread, register resident.
static __global__ void __launch_bounds__(256) cf_peak_dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
static __global__ void __launch_bounds__(256) cf_peak_ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
static __global__ void __launch_bounds__(256) cf_peak_mufu_kernel(float* out, int iters, float a) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f, x4 = x0 + .4f, x5 = x0 + .5f, x6 = x0 + .6f, x7 = x0 + .7f;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = cf_ex2f(x0) ; x1 = cf_ex2f(x1); x2 = cf_ex2f(x2); x3 = cf_ex2f(x3);
        x4 = cf_ex2f(x4); x5 = cf_ex2f(x5); x6 = cf_ex2f(x6); x7 = cf_ex2f(x7);
        x0 -= a; x1 -= a; x2 -= a; x3 -= a; x4 -= a; x5 -= a; x6 -= a; x7 -= a;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
