// gram_mvm_f32p.cuh -- K1p: Float32 value MVM  b <- alpha K a + beta b  for small point dimension with the pair arithmetic in the
// packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: one issue slot, two lanes of work).
//
// Replaces mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta) (reference src/gramian.jl:78-87) for single
// isotropic atoms (EQ, MaternP, RQ with integer alpha) at padded D <= 8; same decomposition, pipeline, summation scheme and epilogue as
// gram_mvm.cuh (K1), direct differences (no norm expansion: exact for any scaling of the points).
//
// Why.  The scalar Float32 K1 (EQ, d = 3) executes 10.4 instructions per pair and is ISSUE bound (issue slots 84 % active, 12.6 cycles
// per warp-pair on a sub-partition: profiles/r2_ncu_c2_f32.md) although the pipes behind it need less: the FMA pipe 3 FADD + 1 FMUL +
// 2 FFMA + 1 FMUL + 1 FFMA = 8 cycles and the MUFU pipe 8 cycles (one ex2, 4 lanes per clock).  Measured (bench_aux/micro/f32x2_probe.cu,
// profiles/r2_f32x2_probe.txt): a packed instruction occupies the FMA pipe for 2 cycles but only ONE issue slot, and a MUFU.EX2 overlaps
// completely with four FFMA2 (8.2 cycles for the five).  So here a thread keeps R rows as scalars and walks the columns in PAIRS:
// the column tile sits in shared memory structure-of-arrays ([coordinate][column], filled by one TMA bulk copy per coordinate from a
// transposed copy of the column points), one LDS.64 delivers coordinate c of columns (j, j + 1) as a packed operand, the row coordinate
// enters FADD2 as a broadcast scalar, and the accumulator pair holds the even / odd column sums.  Per two pairs: 3 FADD2 + FMUL2 + 2 FFMA2
// (distance) + FMUL2 (scale) + 2 MUFU.EX2 + FFMA2 (accumulate) = 10 issue slots instead of 20; both pipes stay at 8 cycles per pair, which
// is the floor of this formulation (2^40 pairs: 236 ms on 148 SMs at 1.965 GHz).
#pragma once
#include "gram_mvm_tf32.cuh"

__device__ __forceinline__ uint64_t cf_pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void cf_upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t cf_sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t cf_add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t cf_mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t cf_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ float cf_sqrtf_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 2^x for two values WITHOUT the MUFU pipe: n = round(x) by the magic-number addition, f = x - n in [-1/2, 1/2], degree-5 minimax polynomial
// of 2^f (relative error 7.5e-8; 2.4e-7 with Float32 Horner rounding -- the same 2 ulp as ex2.approx), n added to the exponent field.
// 3 FADD2 + 5 FFMA2 + 2 LEA for the pair.  Valid for -126 < x < 128 (no clamp: callers bound x).  Used for a FRACTION of the entries of
// kernels whose MUFU pipe is saturated while the FMA pipe idles (gram_mvm_tc5.cuh), the softmax trick of recent attention kernels.
__device__ __forceinline__ uint64_t cf_ex2_poly2(uint64_t x2) {
    const uint64_t magic2 = cf_pk2(12582912.f, 12582912.f);  // 1.5 * 2^23
    const uint64_t t2 = cf_add2(x2, magic2);
    const uint64_t f2 = cf_sub2(x2, cf_sub2(t2, magic2));
    uint64_t p2 = cf_fma2(cf_pk2(0.001327647129073739f, 0.001327647129073739f), f2, cf_pk2(0.009675541892647743f, 0.009675541892647743f));
    p2 = cf_fma2(p2, f2, cf_pk2(0.05550713092088699f, 0.05550713092088699f));
    p2 = cf_fma2(p2, f2, cf_pk2(0.24022120237350464f, 0.24022120237350464f));
    p2 = cf_fma2(p2, f2, cf_pk2(0.6931469440460205f, 0.6931469440460205f));
    p2 = cf_fma2(p2, f2, cf_pk2(1.0000001192092896f, 1.0000001192092896f));
    float plo, phi, tlo, thi;
    cf_upk2(p2, plo, phi);
    cf_upk2(t2, tlo, thi);
    return cf_pk2(__uint_as_float(__float_as_uint(plo) + (__float_as_uint(tlo) << 23)), __uint_as_float(__float_as_uint(phi) + (__float_as_uint(thi) << 23)));
}

// k(r2) of a single isotropic atom for N PAIRS of entries (packed halves), the kind dispatched once: every stage is issued for all N
// pairs, the Horner / power loops over the atom's integer parameter run once.  r2 >= 0 up to rounding; NONNEG = false clamps at 0
// (r2 from the norm expansion), CLAMP = true bounds g = sqrt(r2) so that M(g) never overflows before exp(c g) has flushed to 0.
// MaternP: one MUFU.SQRT and one MUFU.EX2 per entry, everything else packed (src/stationary.jl:148-157 in Float32).
template <int KIND, int N, bool NONNEG, bool CLAMP>
__device__ __forceinline__ void cf_atom_value_f32x2_n(const uint64_t (&r2)[N], const cf_atom_val& A, uint64_t (&kv)[N]) {
    if constexpr (KIND == CF_ATOM_EQ) {
        const uint64_t cl2 = cf_pk2(A.f_clog2e, A.f_clog2e);
#pragma unroll
        for (int u = 0; u < N; u++) {
            float lo, hi;
            cf_upk2(cf_mul2(r2[u], cl2), lo, hi);
            kv[u] = cf_pk2(cf_ex2f(lo), cf_ex2f(hi));
        }
    } else if constexpr (KIND == CF_ATOM_MATERN) {
        const uint64_t cl2 = cf_pk2(A.f_clog2e, A.f_clog2e);
        uint64_t g2[N], e2[N];
#pragma unroll
        for (int u = 0; u < N; u++) {
            float lo, hi;
            cf_upk2(r2[u], lo, hi);
            if (!NONNEG) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            lo = cf_sqrtf_approx(lo); hi = cf_sqrtf_approx(hi);
            if (CLAMP) { lo = fminf(lo, A.f_gmax); hi = fminf(hi, A.f_gmax); }
            g2[u] = cf_pk2(lo, hi);
            cf_upk2(cf_mul2(g2[u], cl2), lo, hi);
            e2[u] = cf_pk2(cf_ex2f(lo), cf_ex2f(hi));
        }
        const int p = A.p;
        if (p == 0) {
#pragma unroll
            for (int u = 0; u < N; u++) kv[u] = e2[u];
        } else {
            uint64_t mp[N];
            const uint64_t top = cf_pk2(A.f_mat[p], A.f_mat[p]);
#pragma unroll
            for (int u = 0; u < N; u++) mp[u] = top;
#pragma unroll 1
            for (int i = p - 1; i >= 0; i--) {
                const uint64_t ci = cf_pk2(A.f_mat[i], A.f_mat[i]);
#pragma unroll
                for (int u = 0; u < N; u++) mp[u] = cf_fma2(mp[u], g2[u], ci);
            }
#pragma unroll
            for (int u = 0; u < N; u++) kv[u] = cf_mul2(mp[u], e2[u]);
        }
    } else {  // CF_ATOM_RQ_INT: (1 + w r2)^-p
        const uint64_t w2 = cf_pk2(A.f_w, A.f_w), one2 = cf_pk2(1.f, 1.f);
        uint64_t ib[N];
#pragma unroll
        for (int u = 0; u < N; u++) {
            float lo, hi;
            cf_upk2(cf_fma2(r2[u], w2, one2), lo, hi);
            if (!NONNEG) { lo = fmaxf(lo, 1.f); hi = fmaxf(hi, 1.f); }
            ib[u] = cf_pk2(cf_rcpf(lo), cf_rcpf(hi));
            kv[u] = ib[u];
        }
#pragma unroll 1
        for (int i = 1; i < A.p; i++) {
#pragma unroll
            for (int u = 0; u < N; u++) kv[u] = cf_mul2(kv[u], ib[u]);
        }
    }
}

#ifndef CF_MVP_UNROLL
#define CF_MVP_UNROLL 2  // column pairs per loop body
#endif
template <int D, int TJ, int NS>
struct cf_mvp_smem {
    static constexpr int bar_bytes = 128;
    static constexpr int row_bytes = TJ * 4;                       // one coordinate of the tile's columns (or the weights)
    static constexpr int stage_bytes = (D + 1) * row_bytes;        // [D][TJ] coordinates | [TJ] weights
    static constexpr int total = bar_bytes + NS * stage_bytes;
};

// transposed copy of the column points: Yt[c * ldt + j] = Y[j * D + c], zero beyond m (ldt = m rounded up to the tile)
static __global__ void cf_transpose_points_f32_kernel(const float* __restrict__ Y, int D, int64_t m, int64_t ldt, float* __restrict__ Yt) {
    const int64_t total = ldt * D;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(q / ldt);
        const int64_t j = q - (int64_t)c * ldt;
        Yt[q] = (j < m) ? Y[j * D + c] : 0.f;
    }
}

// P.X: row points as uploaded (stride D); P.Y: the TRANSPOSED column points, leading dimension P.diag_block (reused: this kernel has no
// symmetric mode); P.a: weights
template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) gram_mvm_f32p_kernel(const __grid_constant__ cf_mvm_params P) {
    using S = cf_mvp_smem<D, TJ, NS>;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    unsigned char* stages = smem + S::bar_bytes;
    const int tid = threadIdx.x;
    const float* __restrict__ Xg = static_cast<const float*>(P.X);
    const float* __restrict__ Yt = static_cast<const float*>(P.Y);
    const float* __restrict__ ag = static_cast<const float*>(P.a);
    const int64_t ldt = P.diag_block;

    const int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    const int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    const int nfull = P.use_tma ? (int)((c1 - c0) / TJ) : 0;  // tiles streamed by TMA
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;              // first column handled by cooperative loads

    if (tid == 0) {
        for (int s = 0; s < NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int tile) {
        const int s = tile % NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = c0 + (int64_t)tile * TJ;
        cf_mbar_expect_tx(&bars[s], (uint32_t)S::stage_bytes);
#pragma unroll
        for (int c = 0; c < D; c++) cf_tma_load_1d(st + c * S::row_bytes, Yt + (int64_t)c * ldt + j0, (uint32_t)S::row_bytes, &bars[s]);
        cf_tma_load_1d(st + D * S::row_bytes, ag + j0, (uint32_t)S::row_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    // this thread's rows (scalars: they enter the packed subtraction as broadcast operands)
    float x[R][D];
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * (NT * R);
    const int64_t rend = P.row0 + P.nrows;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        if (i >= rend) i = rend - 1;  // clamp: computed but never stored
#pragma unroll
        for (int c = 0; c < D; c++) x[r][c] = Xg[i * D + c];
    }
    double tot[R];
#pragma unroll
    for (int r = 0; r < R; r++) tot[r] = 0.0;
    // one tile: TJ columns as TJ / 2 packed pairs; columns past the end of a ragged tile carry a_j = 0 (and zero points)
    auto compute = [&](const unsigned char* __restrict__ st) {
        const uint64_t* as2 = reinterpret_cast<const uint64_t*>(st + D * S::row_bytes);
        uint64_t acc[R];
        constexpr int UNR = CF_MVP_UNROLL;
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ull;
#pragma unroll UNR
        for (int jp = 0; jp < TJ / 2; jp++) {
            uint64_t y2[D];
#pragma unroll
            for (int c = 0; c < D; c++) y2[c] = reinterpret_cast<const uint64_t*>(st + c * S::row_bytes)[jp];
            const uint64_t a2 = as2[jp];
            uint64_t r2[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int c = 0; c < D; c++) {
                    const uint64_t df = cf_sub2(cf_pk2(x[r][c], x[r][c]), y2[c]);
                    r2[r] = (c == 0) ? cf_mul2(df, df) : cf_fma2(df, df, r2[r]);
                }
            }
            uint64_t kv[R];
            cf_atom_value_f32x2_n<KIND, R, true, true>(r2, P.atom, kv);
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = cf_fma2(kv[r], a2, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {  // two-level summation as in K1: Float32 within a tile, Float64 across tiles
            float lo, hi;
            cf_upk2(acc[r], lo, hi);
            tot[r] += (double)(lo + hi);
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        compute(stages + (size_t)s * S::stage_bytes);
        __syncthreads();  // every thread is done reading stage s
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {  // ragged tail (or everything when a is not TMA-aligned): cooperative loads, zero filled
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        float* ys = reinterpret_cast<float*>(stages);
        __syncthreads();
        for (int q = tid; q < (D + 1) * TJ; q += NT) {
            const int c = q / TJ, j = q - c * TJ;
            ys[q] = (j < cnt) ? (c < D ? Yt[(int64_t)c * ldt + j0 + j] : ag[j0 + j]) : 0.f;
        }
        __syncthreads();
        compute(stages);
    }

    float* out = static_cast<float*>(P.out);
    const float* yin = static_cast<const float*>(P.yin);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < rend) {
            const int64_t o = i - P.row0;
            if (P.direct) {
                double v = P.alpha * tot[r];
                if (P.beta != 0.0) v += P.beta * (double)yin[o];
                out[o] = (float)v;
                for (int p = 0; p < P.peers.n; p++) static_cast<float*>(P.peers.ptr[p])[o] = (float)v;  // NVLink peer stores
            } else {
                reinterpret_cast<double*>(P.out)[(int64_t)blockIdx.y * P.nrows + o] = tot[r];
            }
        }
    }
}

#ifndef __CUDACC_RTC__ // host side
template <int D, int KIND, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_mvp_launch(const cf_mvm_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvp_smem<D, TJ, NS>;
    auto kern = gram_mvm_f32p_kernel<D, KIND, R, NT, TJ, NS, MINB>;
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

// tuning: rows per thread, threads, tile, stages, CTAs per SM
#ifndef CF_MVP_R
#define CF_MVP_R 8
#endif
#ifndef CF_MVP_NT
#define CF_MVP_NT 128
#endif
#ifndef CF_MVP_MINB
#define CF_MVP_MINB 4
#endif
template <int D, bool OK = (D <= 8)>
struct cf_mvp_entry {
    static constexpr cf_mvm_launch_fn fn[3] = {nullptr, nullptr, nullptr};
    static constexpr cf_mvm_config cfg = {0, 0, 0, 0};
};
template <int D>
struct cf_mvp_entry<D, true> {
    static constexpr int R = (D <= 4) ? CF_MVP_R : 4, NT = CF_MVP_NT, TJ = 128, NS = 3, MINB = CF_MVP_MINB;
    static constexpr cf_mvm_launch_fn fn[3] = {&cf_mvp_launch<D, CF_ATOM_EQ, R, NT, TJ, NS, MINB>, &cf_mvp_launch<D, CF_ATOM_MATERN, R, NT, TJ, NS, MINB>,
                                               &cf_mvp_launch<D, CF_ATOM_RQ_INT, R, NT, TJ, NS, MINB>};
    static constexpr cf_mvm_config cfg = {NT * R, TJ, cf_mvp_smem<D, TJ, NS>::total, MINB};
};
#endif // !__CUDACC_RTC__
