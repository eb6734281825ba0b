// gram_mvm.cuh -- K1/K2/K3: lazy-Gramian matrix-vector product  y <- alpha K a + beta y  on sm_100a.
//
// Replaces the reference hot loop  (src/gramian.jl:78-87)
//     @threads for i in 1:n;  @simd for j in 1:m;  y[i] += alpha * G[i, j] * x[j]
// with G[i,j] = k(x_i, y_j) evaluated on the fly (src/gramian.jl:37-40, src/stationary.jl:9, src/mercer.jl:3).
//
// Decomposition.  grid = (row tiles, column chunks).  A CTA owns NT*R rows: every thread keeps R points
// x_i (R*D registers) and R accumulators for the whole sweep.  It streams its column chunk of (y_j, a_j)
// through shared memory in tiles of TJ points: one elected thread issues two 1-D TMA bulk copies
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx) per tile into an NS-stage ring, every thread
// waits on the stage's mbarrier, then all threads read y_j / a_j as shared-memory BROADCASTS (every lane
// the same address) -- one LDS serves 32*R pair evaluations, so the inner loop is pure FP64-pipe work:
// D subtractions + D FMAs (r2), the exp / sqrt / polynomial of the kernel, one FMA into the accumulator.
// Per-tile partial sums are folded into a running total (two-level summation: error ~ sqrt(TJ)+sqrt(m/TJ)
// ulps instead of sqrt(m)).  With several column chunks each CTA writes a partial row sum and
// gram_reduce_partials applies alpha/beta; with one chunk the epilogue writes y directly.
// HBM traffic: X, Y, a once (they live in the 126 MB L2 afterwards) + y: negligible next to the arithmetic
// (SURVEY.md section 8d), the bound is the FP64 FMA pipe.
#pragma once
#include "cf_math.cuh"

// Fused all-gather: when chained MVMs run on several GPUs of one process, the epilogue that produces a row block of the
// result also stores it into every peer's copy of the vector through peer-mapped pointers (NVLink P2P stores); ptr[p]
// already points at this row block inside peer p's vector.  n == 0: no peers.
struct cf_peer_out {
    int32_t n;
    int32_t pad_;
    void* ptr[8];
};

struct cf_mvm_params {
    const void* X;        // rows: padded AoS, stride D elements
    const void* Y;        // columns: padded AoS, stride D elements
    const void* a;        // weights, length m
    void* out;            // y (direct) or partial sums [chunks][nrows]
    const void* yin;      // y for the beta term (direct mode)
    const double* exp2_tbl;
    int64_t row0, nrows;  // rows [row0, row0 + nrows) are computed; out index = i - row0
    int64_t m;            // number of columns
    int64_t cols_per_chunk; // multiple of TJ
    double alpha, beta;
    int direct;           // 1: write alpha*sum + beta*y, 0: write the raw partial sum
    int use_tma;          // 0: a is not 16-byte aligned -> cooperative loads
    int64_t diag_block;   // > 0: each CTA sweeps only the columns of its own row block of this size (symmetric variant)
    cf_atom_val atom;     // the single atom (specialised kinds)
    cf_sop_val sop;       // generic sum of products (KIND == CF_ATOM_SOP)
    cf_peer_out peers;    // direct mode: also store the finished rows into these peer vectors
    const void* xn;       // squared norms of the rows / columns (tensor-core variant gram_mvm_dmma.cuh only; X, Y then point at
    const void* yn;       // the point copies with the padded row stride)
    double eqc[4];        // gram_mvm_eq.cuh: polynomial constants of exp(f ln2 / 32768) (constant-bank operands)
};

// ---- mbarrier / TMA 1-D bulk copy wrappers (PTX ISA: cp.async.bulk, mbarrier) ------------------------------
__device__ __forceinline__ void cf_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cf_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void cf_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void cf_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cf_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cf_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(cf_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cf_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "CF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra CF_DONE;\n"
        "bra CF_WAIT;\n"
        "CF_DONE:\n"
        "}" ::"r"(cf_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ double cf_shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// shared memory carve-up (bytes)
template <typename T, int D, int TJ, int NS>
struct cf_mvm_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8; // also kept for fp32 (unused, simplifies layout)
    static constexpr int bar_bytes = 128;
    static constexpr int y_bytes = TJ * D * (int)sizeof(T);
    static constexpr int a_bytes = TJ * (int)sizeof(T);
    static constexpr int stage_bytes = ((y_bytes + a_bytes + 127) / 128) * 128;
    static constexpr int total = tbl_bytes + bar_bytes + NS * stage_bytes;
};

// k(x_r, y_j) for the R rows of a thread ----------------------------------------------------------------------
template <typename T, int D, int KIND, int R>
__device__ __forceinline__ void cf_rows_value(const T (&x)[R][D], const T (&yj)[D], const cf_atom_val& atom, const cf_sop_val& sop,
                                              cf_tbl_t tbl_lane, T (&kv)[R]) {
    T r2[R], dt[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        r2[r] = 0; dt[r] = 0;
        if (KIND != CF_ATOM_LINE) {
#pragma unroll
            for (int c = 0; c < D; c++) {
                T df = x[r][c] - yj[c];
                r2[r] = (c == 0) ? df * df : fma(df, df, r2[r]);
            }
        }
        if (KIND == CF_ATOM_LINE || KIND == CF_ATOM_SOP) {
#pragma unroll
            for (int c = 0; c < D; c++) dt[r] = (c == 0) ? x[r][c] * yj[c] : fma(x[r][c], yj[c], dt[r]);
        }
    }
    if constexpr (sizeof(T) == 8) {
        if constexpr (KIND == CF_ATOM_SOP) cf_sop_value_n<R>(r2, dt, sop, tbl_lane, kv);
        else if constexpr (KIND == CF_ATOM_MATERN) cf_atom_matern_n<R>(r2, atom, tbl_lane, kv);
        else if constexpr (KIND == CF_ATOM_RQ_INT) cf_atom_rq_int_n<R>(r2, atom, kv);
        else {
#pragma unroll
            for (int r = 0; r < R; r++) kv[r] = cf_atom_value<KIND>(r2[r], dt[r], atom, tbl_lane);
        }
    } else {
        if constexpr (KIND == CF_ATOM_SOP) cf_sop_value_f32_n<R>(r2, dt, sop, kv);
        else if constexpr (KIND == CF_ATOM_MATERN || KIND == CF_ATOM_RQ_INT) cf_atom_value_f32_n<R>(r2, dt, atom, kv); // one loop over p for R values
        else {
#pragma unroll
            for (int r = 0; r < R; r++) kv[r] = cf_atom_value_f32<KIND>(r2[r], dt[r], atom);
        }
    }
}

template <typename T, int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) gram_mvm_kernel(const __grid_constant__ cf_mvm_params P) {
    using S = cf_mvm_smem<T, D, TJ, NS>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;

    const int tid = threadIdx.x;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const T* __restrict__ Xg = static_cast<const T*>(P.X);
    const T* __restrict__ Yg = static_cast<const T*>(P.Y);
    const T* __restrict__ ag = static_cast<const T*>(P.a);

    int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    if (P.diag_block > 0) { // columns of this CTA's own row block only
        c0 = ((P.row0 + (int64_t)blockIdx.x * (NT * R)) / P.diag_block) * P.diag_block;
        c1 = (c0 + P.diag_block < P.m) ? c0 + P.diag_block : P.m;
    }
    const int64_t ncols = c1 - c0;
    const int nfull = P.use_tma ? (int)(ncols / TJ) : 0;  // tiles streamed by TMA
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;          // first column handled by cooperative loads

    if (sizeof(T) == 8) cf_fill_exp_table(tbl, P.exp2_tbl, tid, NT);
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
        cf_tma_load_1d(st, Yg + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, ag + j0, (uint32_t)S::a_bytes, &bars[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < NS && t < nfull; t++) issue(t);
    }

    // this thread's rows
    T x[R][D];
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * (NT * R);
    const int64_t rend = P.row0 + P.nrows;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        if (i >= rend) i = rend - 1; // clamp: computed but never stored
#pragma unroll
        for (int c = 0; c < D; c++) x[r][c] = Xg[i * D + c];
    }

    double tot[R];
#pragma unroll
    for (int r = 0; r < R; r++) tot[r] = 0.0;

    auto compute = [&](const T* __restrict__ ys, const T* __restrict__ as, int cnt) {
        T acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0;
#pragma unroll 2
        for (int j = 0; j < cnt; j++) {
            T yj[D];
#pragma unroll
            for (int c = 0; c < D; c++) yj[c] = ys[j * D + c];
            const T aj = as[j];
            T kv[R];
            cf_rows_value<T, D, KIND, R>(x, yj, P.atom, P.sop, tbl_lane, kv);
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = fma(kv[r], aj, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; r++) tot[r] += (double)acc[r];
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        const uint32_t parity = (uint32_t)((t / NS) & 1);
        cf_mbar_wait(&bars[s], parity);
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const T*>(st), reinterpret_cast<const T*>(st + S::y_bytes), TJ);
        __syncthreads(); // every thread is done reading stage s
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    // remainder (partial last tile, or everything when the weight vector is not TMA-aligned)
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        T* ys = reinterpret_cast<T*>(stages);
        T* as = reinterpret_cast<T*>(stages + S::y_bytes);
        __syncthreads();
        for (int q = tid; q < cnt * D; q += NT) ys[q] = Yg[j0 * D + q];
        for (int q = tid; q < cnt; q += NT) as[q] = ag[j0 + q];
        __syncthreads();
        compute(ys, as, cnt);
    }

    // epilogue
    T* out = static_cast<T*>(P.out);
    const T* yin = static_cast<const T*>(P.yin);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < rend) {
            const int64_t o = i - P.row0;
            if (P.direct) {
                double v = P.alpha * tot[r];
                if (P.beta != 0.0) v += P.beta * (double)yin[o];
                out[o] = (T)v;
                for (int p = 0; p < P.peers.n; p++) static_cast<T*>(P.peers.ptr[p])[o] = (T)v; // NVLink peer stores
            } else {
                reinterpret_cast<double*>(P.out)[(int64_t)blockIdx.y * P.nrows + o] = tot[r];
            }
        }
    }
}

// y[o] = alpha * sum_s partial[s][o] + beta * y[o]   (beta == 0 overwrites: reference src/gramian.jl:80)
template <typename T>
__global__ void gram_reduce_partials(const double* __restrict__ partial, int chunks, int64_t nrows, T* __restrict__ y,
                                     const T* __restrict__ yin, double alpha, double beta, const cf_peer_out peers) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nrows; o += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < chunks; c++) s += partial[(int64_t)c * nrows + o];
        double v = alpha * s;
        if (beta != 0.0) v += beta * (double)yin[o];
        y[o] = (T)v;
        for (int p = 0; p < peers.n; p++) static_cast<T*>(peers.ptr[p])[o] = (T)v; // NVLink peer stores
    }
}

// launcher table entry ----------------------------------------------------------------------------------------
struct cf_mvm_config {
    int rows_per_cta; // NT * R
    int tj;
    int smem_bytes;
    int min_blocks;
};
#ifndef __CUDACC_RTC__ // host side: not part of run-time specialised builds
typedef cudaError_t (*cf_mvm_launch_fn)(const cf_mvm_params& P, dim3 grid, cudaStream_t stream);

template <typename T, int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_mvm_launch(const cf_mvm_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvm_smem<T, D, TJ, NS>;
    auto kern = gram_mvm_kernel<T, D, KIND, R, NT, TJ, NS, MINB>;
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
