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

// element type conversion (Float32 handles run their derivative operators, solves and d > 32 products on a Float64 shadow)
template <typename S, typename T>
__global__ void cf_convert_kernel(const S* __restrict__ src, T* __restrict__ dst, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) dst[q] = (T)src[q];
}

// ---- Matrix!(M, G): M[(i - r0) + ld (j - j0)] = k(x_i, y_j)   (reference src/gramian.jl:107-114) ------------------
// one thread per entry, generic sum-of-products evaluation; not a hot path.
template <typename T>
__global__ void gram_dense_kernel(const T* __restrict__ X, const T* __restrict__ Y, int D, const __grid_constant__ cf_sop_val prog,
                                  const double* __restrict__ exp2_tbl, int64_t r0, int64_t nrows, int64_t j0, int64_t ncols,
                                  T* __restrict__ M, int64_t ld) {
    __shared__ double tbl[CF_EXP_TBL_DOUBLES];
    cf_fill_exp_table(tbl, exp2_tbl, threadIdx.x, blockDim.x);
    __syncthreads();
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, threadIdx.x);
    cf_tbl_publish(tbl_lane);
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

// ---- K4: B <- alpha K A + beta B with A m x p: every kernel entry is evaluated ONCE and contracted with a p-wide slice ----
// (the reference re-evaluates it for each of the p columns: src/gramian.jl:89-99).
// CTA = 256 threads, TI = 128 rows, tiles of TJ = 32 columns, PC = 64 right-hand sides per pass.
//   stage  : (y tile, |y|^2 tile, A^T tile) arrive by three 1-D TMA bulk copies into a 2-stage mbarrier ring
//            (A is transposed once per call to At[j][c] so that a tile is one contiguous block).
//   phase A: K tile (TI x TJ) -> shared memory.  Thread (i = t & 127, jh = t >> 7) evaluates 16 entries of its row with
//            x_i in registers; y_j is read as shared-memory broadcasts.  For d >= 8, and when the data is well scaled
//            (host check: (d + 2) eps max|x|^2 < 1e-13), r2 = |x|^2 + |y|^2 - 2 x.y shares the d FMAs of the dot product
//            ("r2 from norms"); otherwise direct differences as the reference does.
//   phase B: B tile (TI x PC) += K tile . A tile with 8 x 4 register tiles per thread: per k, 4 + 2 LDS.128 feed 32 FMAs.
#define CF_MM_TI 128
#define CF_MM_TJ 32
#define CF_MM_PC 64
#define CF_MM_NS 2

#ifndef CF_MM_AG
#define CF_MM_AG 8
#endif
template <typename T, int D>
struct cf_mm_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int ks_bytes = CF_MM_TJ * CF_MM_TI * (int)sizeof(T);   // Ks[j][i]
    static constexpr int y_bytes = CF_MM_TJ * D * (int)sizeof(T);
    static constexpr int n_bytes = CF_MM_TJ * (int)sizeof(T);
    static constexpr int a_bytes = CF_MM_TJ * CF_MM_PC * (int)sizeof(T);    // As[j][c]
    static constexpr int stage_bytes = ((y_bytes + n_bytes + a_bytes + 127) / 128) * 128;
    // fp64, d >= 8: the row tile's points live in shared memory, row-major with stride D + 2 (LDS.128 of a warp's 32
    // consecutive rows is conflict-free for every even D), which frees 2 D registers for wider phase-A groups
    static constexpr bool xs = (sizeof(T) == 8 && D >= 8 && D % 2 == 0);
    static constexpr int xstr = D + 2;
    static constexpr int xs_bytes = xs ? CF_MM_TI * xstr * (int)sizeof(T) : 0;
    static constexpr int total_nox = tbl_bytes + bar_bytes + ks_bytes + CF_MM_NS * stage_bytes;
    static constexpr int total = total_nox + xs_bytes;
};

struct cf_mm_params {
    const void* X; const void* Y; const void* xn; const void* yn; // points and squared norms
    const void* At;   // transposed weights, m x PC (row j holds the PC columns of this pass, zero padded)
    void* B;
    const double* exp2_tbl;
    int64_t row0, nrows, m, ldb;
    int nrhs;         // columns in this pass (<= CF_MM_PC)
    int use_norms;
    int ldat;         // row stride of At (elements)
    double alpha, beta;
    cf_sop_val sop;
};

// At[j][c] = c < nrhs ? A[j + lda c] : 0   (one pass of <= CF_MM_PC columns)
template <typename T>
__global__ void cf_transpose_rhs(const T* __restrict__ A, int64_t lda, int64_t m, int nrhs, T* __restrict__ At, int ldat) {
    __shared__ T tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int64_t j = j0 + threadIdx.x;
        const int c = c0 + cc;
        tile[cc][threadIdx.x] = (j < m && c < nrhs) ? A[j + lda * c] : (T)0;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int64_t j = j0 + jj;
        const int c = c0 + threadIdx.x;
        if (j < m && c < CF_MM_PC) At[j * ldat + c] = tile[threadIdx.x][jj];
    }
}

// squared norms of padded points + validation: flags[0] += number of non-finite coordinates,
// flags[1] = max squared norm (bit pattern of a non-negative double: integer order == floating-point order)
// ARD metric (reference src/transformation.jl:42-45): coordinate c of every point times scale[c] = 1/sqrt(l_c), in place,
// on the padded device copy (padding columns stay 0)
template <typename T>
__global__ void cf_scale_coords_kernel(T* __restrict__ X, int D, int d, int64_t n, const double* __restrict__ scale) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n * D; q += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(q % D);
        if (c < d) X[q] = (T)((double)X[q] * scale[c]);
    }
}

template <typename T>
__global__ void cf_sqnorm_validate_kernel(const T* __restrict__ X, int D, int64_t n, T* __restrict__ out, double* flags) {
    unsigned long long bad = 0;
    double mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T s = 0;
        for (int c = 0; c < D; c++) {
            const T v = X[i * D + c];
            if (!isfinite(v)) bad++;
            s = fma(v, v, s);
        }
        out[i] = s;
        if ((double)s > mx) mx = (double)s;
    }
    unsigned long long* f = reinterpret_cast<unsigned long long*>(flags);
    if (bad) atomicAdd(&f[0], bad);
    if (mx > 0.0 && isfinite(mx)) atomicMax(&f[1], (unsigned long long)__double_as_longlong(mx));
}

// NTB = 256: 8 x 4 register tiles, 16 entries per thread in phase A; NTB = 512: 4 x 4 tiles, 8 entries (<= 128 registers,
// twice the warps to hide FP64 and shared-memory latency)
template <typename T, int D, int NTB, int AGP = CF_MM_AG>
__global__ void __launch_bounds__(NTB, 1) gram_mm_kernel(const __grid_constant__ cf_mm_params P) {
    constexpr int JQ = NTB / 128;            // column groups in phase A
    constexpr int EPT = CF_MM_TJ / JQ;        // entries per thread in phase A
    constexpr int RH = (NTB == 256) ? 4 : 2;  // LDS.128 of K per k step in phase B; rows per thread = 2 RH
    constexpr bool XS = cf_mm_smem<T, D>::xs; // fp64, d >= 8: x_i lives in shared memory instead of 2 D registers
    constexpr int XSTR = cf_mm_smem<T, D>::xstr;
    using S = cf_mm_smem<T, D>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    T* Ks = reinterpret_cast<T*>(smem + S::tbl_bytes + S::bar_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes + S::ks_bytes;
    const int tid = threadIdx.x;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const T* __restrict__ Xg = static_cast<const T*>(P.X);
    const T* __restrict__ Yg = static_cast<const T*>(P.Y);
    const T* __restrict__ yng = static_cast<const T*>(P.yn);
    const T* __restrict__ Atg = static_cast<const T*>(P.At);
    if (sizeof(T) == 8) cf_fill_exp_table(tbl, P.exp2_tbl, tid, NTB);
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
        cf_tma_load_1d(st, Yg + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, yng + j0, (uint32_t)S::n_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes + S::n_bytes, Atg + j0 * CF_MM_PC, (uint32_t)S::a_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < CF_MM_NS && t < nfull; t++) issue(t);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * CF_MM_TI;
    const int64_t rend = P.row0 + P.nrows;
    const int li = tid & 127, jh = tid >> 7;
    T x[XS ? 1 : D], xnorm;
    T* Xs = reinterpret_cast<T*>(smem + S::total_nox);
    {
        int64_t i = rbase + li;
        if (i >= rend) i = rend - 1;
        if constexpr (XS) {
            for (int q = tid; q < CF_MM_TI * D; q += NTB) {
                const int row = q / D, c = q - row * D;
                int64_t ir = rbase + row;
                if (ir >= rend) ir = rend - 1;
                Xs[row * XSTR + c] = Xg[ir * D + c];
            }
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) x[c] = Xg[i * D + c];
        }
        xnorm = static_cast<const T*>(P.xn)[i];
    }
    if constexpr (XS) __syncthreads();
    // phase B tile: rows {rbase_b + 32 h + 2 rg + b : h < RH, b < 2}, columns 4 cg .. 4 cg + 3
    const int rg = tid & 15;
    const int rbase_b = (NTB == 256) ? 0 : 64 * ((tid >> 4) & 1);
    const int cg = (NTB == 256) ? (tid >> 4) : (tid >> 5);
    T acc[2 * RH][4];
#pragma unroll
    for (int a = 0; a < 2 * RH; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0;

    auto tile_compute = [&](const T* __restrict__ ys, const T* __restrict__ yns, const T* __restrict__ As, int cnt) {
        // phase A: AG entries at a time (independent FMA chains hide the FP64 latency with only 2 warps per scheduler;
        // the program decode of cf_sop_value_n is paid once per group)
        constexpr int AG = XS ? (AGP < EPT ? AGP : EPT) : 4;
        for (int q0 = 0; q0 < EPT; q0 += AG) {
            const int jb = jh * EPT + q0;
            T r2[AG], dt[AG];
#pragma unroll
            for (int u = 0; u < AG; u++) { r2[u] = 0; dt[u] = 0; }
            if constexpr (XS) {
                if (P.use_norms) {
#pragma unroll
                    for (int c = 0; c < D; c += 2) {
                        const double2 xv = *reinterpret_cast<const double2*>(&Xs[li * XSTR + c]);
#pragma unroll
                        for (int u = 0; u < AG; u++) {
                            const double2 yv = *reinterpret_cast<const double2*>(&ys[(jb + u) * D + c]);
                            dt[u] = fma(xv.x, yv.x, dt[u]);
                            dt[u] = fma(xv.y, yv.y, dt[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < AG; u++) {
                        const T v = fma((T)-2, dt[u], xnorm + yns[jb + u]);
                        r2[u] = (v > (T)0) ? v : (T)0;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < D; c += 2) {
                        const double2 xv = *reinterpret_cast<const double2*>(&Xs[li * XSTR + c]);
#pragma unroll
                        for (int u = 0; u < AG; u++) {
                            const double2 yv = *reinterpret_cast<const double2*>(&ys[(jb + u) * D + c]);
                            const T d0 = xv.x - yv.x, d1 = xv.y - yv.y;
                            r2[u] = fma(d0, d0, r2[u]);
                            r2[u] = fma(d1, d1, r2[u]);
                            dt[u] = fma(xv.x, yv.x, dt[u]);
                            dt[u] = fma(xv.y, yv.y, dt[u]);
                        }
                    }
                }
            } else if (P.use_norms) {
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const T xc = x[XS ? 0 : c];
#pragma unroll
                    for (int u = 0; u < AG; u++) dt[u] = fma(xc, ys[(jb + u) * D + c], dt[u]);
                }
#pragma unroll
                for (int u = 0; u < AG; u++) {
                    const T v = fma((T)-2, dt[u], xnorm + yns[jb + u]);
                    r2[u] = (v > (T)0) ? v : (T)0;
                }
            } else {
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const T xc = x[XS ? 0 : c];
#pragma unroll
                    for (int u = 0; u < AG; u++) {
                        const T yv = ys[(jb + u) * D + c];
                        const T df = xc - yv;
                        r2[u] = fma(df, df, r2[u]);
                        dt[u] = fma(xc, yv, dt[u]);
                    }
                }
            }
            T kv[AG];
            if constexpr (sizeof(T) == 8) cf_sop_value_n<AG>(r2, dt, P.sop, tbl_lane, kv);
            else cf_sop_value_f32_n<AG>(r2, dt, P.sop, kv);
#pragma unroll
            for (int u = 0; u < AG; u++) Ks[(jb + u) * CF_MM_TI + li] = (jb + u < cnt) ? kv[u] : (T)0; // past the end: no contribution
        }
        __syncthreads();
        // phase B: this thread's rows are {32 a2 + 2 rg + b}: the 16 lanes of a half-warp read 256 contiguous bytes of Ks.
        // Software pipelined: the operands of step k+1 are loaded while the 32 FMAs of step k issue.
        T kr[2][2 * RH], ac[2][4];
        auto loadk = [&](int k, T (&krr)[2 * RH], T (&acc_)[4]) {
            if constexpr (sizeof(T) == 8) {
#pragma unroll
                for (int h = 0; h < RH; h++) {
                    const double2 v = *reinterpret_cast<const double2*>(&Ks[k * CF_MM_TI + rbase_b + 32 * h + 2 * rg]);
                    krr[2 * h] = v.x; krr[2 * h + 1] = v.y;
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const double2 v = *reinterpret_cast<const double2*>(&As[k * CF_MM_PC + 4 * cg + 2 * h]);
                    acc_[2 * h] = v.x; acc_[2 * h + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int h = 0; h < RH; h++) {
                    const float2 v = *reinterpret_cast<const float2*>(&Ks[k * CF_MM_TI + rbase_b + 32 * h + 2 * rg]);
                    krr[2 * h] = v.x; krr[2 * h + 1] = v.y;
                }
                const float4 v = *reinterpret_cast<const float4*>(&As[k * CF_MM_PC + 4 * cg]);
                acc_[0] = v.x; acc_[1] = v.y; acc_[2] = v.z; acc_[3] = v.w;
            }
        };
        loadk(0, kr[0], ac[0]);
#pragma unroll 1
        for (int k = 0; k < CF_MM_TJ; k += 2) {
            loadk(k + 1, kr[1], ac[1]);
#pragma unroll
            for (int a = 0; a < 2 * RH; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(kr[0][a], ac[0][b], acc[a][b]);
            if (k + 2 < CF_MM_TJ) loadk(k + 2, kr[0], ac[0]);
#pragma unroll
            for (int a = 0; a < 2 * RH; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(kr[1][a], ac[1][b], acc[a][b]);
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % CF_MM_NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / CF_MM_NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        tile_compute(reinterpret_cast<const T*>(st), reinterpret_cast<const T*>(st + S::y_bytes),
                     reinterpret_cast<const T*>(st + S::y_bytes + S::n_bytes), CF_MM_TJ);
        __syncthreads(); // Ks and stage s are free again
        if (tid == 0 && t + CF_MM_NS < nfull) issue(t + CF_MM_NS);
    }
    if ((int64_t)nfull * CF_MM_TJ < P.m) { // ragged last tile: cooperative loads, zero fill
        const int64_t j0 = (int64_t)nfull * CF_MM_TJ;
        const int cnt = (int)(P.m - j0);
        T* ys = reinterpret_cast<T*>(stages);
        T* yns = reinterpret_cast<T*>(stages + S::y_bytes);
        T* As = reinterpret_cast<T*>(stages + S::y_bytes + S::n_bytes);
        __syncthreads();
        for (int q = tid; q < CF_MM_TJ * D; q += NTB) ys[q] = (q < cnt * D) ? Yg[j0 * D + q] : (T)0;
        for (int q = tid; q < CF_MM_TJ; q += NTB) yns[q] = (q < cnt) ? yng[j0 + q] : (T)0;
        for (int q = tid; q < CF_MM_TJ * CF_MM_PC; q += NTB) As[q] = (q < cnt * CF_MM_PC) ? Atg[j0 * CF_MM_PC + q] : (T)0;
        __syncthreads();
        tile_compute(ys, yns, As, cnt);
    }
    T* Bg = static_cast<T*>(P.B);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int c = 4 * cg + b;
        if (c >= P.nrhs) continue;
#pragma unroll
        for (int a = 0; a < 2 * RH; a++) {
            const int64_t i = rbase + rbase_b + 32 * (a >> 1) + 2 * rg + (a & 1);
            if (i >= rend) continue;
            T* o = Bg + (i - P.row0) + P.ldb * c;
            double v = P.alpha * (double)acc[a][b];
            if (P.beta != 0.0) v += P.beta * (double)(*o);
            *o = (T)v;
        }
    }
}

#ifndef __CUDACC_RTC__ // host side: not part of run-time specialised builds
typedef cudaError_t (*cf_mm_launch_fn)(const cf_mm_params& P, int row_tiles, cudaStream_t stream);
template <typename T, int D, int NTB, int AGP = CF_MM_AG>
cudaError_t cf_mm_launch(const cf_mm_params& P, int row_tiles, cudaStream_t stream) {
    using S = cf_mm_smem<T, D>;
    constexpr int smem_bytes = S::total;
    auto kern = gram_mm_kernel<T, D, NTB, AGP>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<row_tiles, NTB, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}
#endif // !__CUDACC_RTC__

// ---- K6: vector helpers for conjugate gradients (Float64) -------------------------------------------------------------
// z = a x + b y
// (no __restrict__: the CG driver updates in place, z aliases x or y)
static __global__ void cf_axpby_kernel(double* z, double a, const double* x, double b, const double* y, int64_t n) {
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
// dependent-chain-free FMAs: 16 independent chains per thread, register resident, 128 FMAs per loop trip so that the
// loop overhead (3 non-FMA instructions) is < 2.5 % of the issue slots.  Operand form x = fma(x, y, b) with y a per-thread
// register and b a constant: an FP64 instruction with TWO constant operands needs an extra cycle and one with THREE register
// operands needs three (bench_aux/micro/fp64_issue_probe.cu: 56.9, 41.6 and, for this form, 63.5 lane-FMA per clock per SM).
#define CF_PROBE_BODY(T, FMA)                                                                                         \
    T x[16], y[16];                                                                                                   \
    _Pragma("unroll") for (int q = 0; q < 16; q++) { x[q] = (T)(threadIdx.x + q); y[q] = a + (T)1e-9 * (T)(threadIdx.x + q); } \
    for (int i = 0; i < iters; i += 8) {                                                                              \
        _Pragma("unroll") for (int u = 0; u < 8; u++) {                                                               \
            _Pragma("unroll") for (int q = 0; q < 16; q++) x[q] = FMA(x[q], y[q], b);                                 \
        }                                                                                                             \
    }                                                                                                                 \
    T s = 0;                                                                                                          \
    _Pragma("unroll") for (int q = 0; q < 16; q++) s += x[q] + y[q];                                                  \
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;

static __global__ void __launch_bounds__(256) cf_peak_dfma_kernel(double* out, int iters, double a, double b) { CF_PROBE_BODY(double, fma) }
static __global__ void __launch_bounds__(256) cf_peak_ffma_kernel(float* out, int iters, float a, float b) { CF_PROBE_BODY(float, fmaf) }
static __global__ void __launch_bounds__(256) cf_peak_mufu_kernel(float* out, int iters, float a) {
    float x[16];
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = threadIdx.x * 1e-3f + 0.05f * q;
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int q = 0; q < 16; q++) x[q] = cf_ex2f(x[q]);
        }
#pragma unroll
        for (int q = 0; q < 16; q++) x[q] -= a; // keeps the values bounded; 16 FADD per 64 MUFU
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < 16; q++) s += x[q];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
