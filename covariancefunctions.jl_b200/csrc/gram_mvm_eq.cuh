// gram_mvm_eq.cuh -- K1e: the Float64 EQ lazy-Gramian matrix-vector product with the exponent formed in the scaled domain.
//
// Same operator, decomposition, TMA ring and epilogue as gram_mvm.cuh (reference hot loop src/gramian.jl:78-87 with
// k = exp(-r2 / (2 l^2)), src/stationary.jl:42, src/transformation.jl:19); what changes is the per-pair instruction sequence.
// An SM sub-partition issues one instruction per cycle and an FP64 instruction holds the port for two (profiles/README.md),
// so the pair costs 2 * #FP64 + #other cycles.  K1 spends 15 FP64 + ~8 other (38 cycles); this kernel spends 12 + ~4 (28):
//
//   exp(c r2) = 2^(T / 256),  T = c1 r2,  c1 = 256 c / ln2,  r2 = |x|^2 + (|y|^2 - 2 x.y)
//   i0 = |y_j|^2 - 2 x_i . y_j                     D FMAs on (-2 x_i) held in registers, |y_j|^2 from the tile      (K1: 2D)
//   t  = fma(i0, c1, M_i)                          M_i = 1.5 2^52 + rint(c1 |x_i|^2): ONE FMA rounds T to an integer kk = lo(t)
//   kd = t - M_i ;  f = fma(i0, c1, -kd)           reduced argument, |f| <= 1/2 (an exact FMA: no Cody-Waite constant)
//   p  = ((g3 f + g2) f + g1) f + g0               exp(f ln2 / 256) = 1 + f p, kernel-independent constants
//   s  = 2^k 2^(j/256),  kk = 256 k + j            AND + multiply-add (table address), LDS.64, one multiply-add on the high
//                                                  word (pre-compensated table entries, capi.cu get_ctx)
//   e  = fma(s f, p, s) ;  acc = fma(e, a_j, acc)
// and the row's leftover factor exp((c1 |x_i|^2 - rint(c1 |x_i|^2)) ln2 / 256) is applied once per row in the epilogue.
// No clamp: the host only selects this kernel when |c| (|x| + |y|)^2 < 700 for every pair (exponent field cannot wrap) and when
// the cancellation error of the norm expansion, eps (|x|^2 + |y|^2) |c|, is below 1e-13 -- the same scale check that guards the
// tensor-core kernels (capi.cu); otherwise K1 (direct differences, clamped) runs.
//
// FAST = 2 (round 2): the same skeleton for a single MaternP atom with p >= 1 (src/stationary.jl:134-158), "K1m".  What carries over is
// the distance from the norm expansion (D FMAs + one add instead of 2 D), the clamp-free exp with the FMA-rounded exponent and the
// two-instruction table scaling; what is new is a 6-instruction square root (MUFU.RSQ64H seed, one coupled Newton step, one residual
// correction with the UNREFINED half-reciprocal, whose 2^-22 error only multiplies the 2^-44 residual) with the |.| and the underflow
// guard as two integer instructions on the high word.  Per pair: D + 1 (r2) + 6 (sqrt) + 8 (exp) + p (Horner) + 2 = D + p + 17 FP64
// instructions and ~6 others, against 2 D + 7 + 8 + p + 2 and ~15 in K1 (d = 3, p = 2: 22 + 6 against 25 + 15).  The host selects it
// under the same kind of scale check as the EQ form (capi.cu set_norm_flags: mat_fast).  A slightly NEGATIVE r2 (cancellation at
// (nearly) coincident points) is harmless: M(g) e^{-g} is even in g to third order for p >= 1, so sqrt(|r2|) ~ 1e-8 gives 1 - O(1e-16).
#pragma once
#include "gram_mvm.cuh"

// MaternP values M(g) exp(c g), g = sqrt(|r2|), for R rows at once; r2 from the norm expansion.  No clamps (host-checked ranges).
// c1 = 256 c / ln2 (A.e.c1), g0..g2 = the exp polynomial constants of this file (P.eqc), A.mat = Horner coefficients of M in g.
// `magic` (= CF_MAGIC) and `top` (= A.mat[A.p]) are passed in REGISTERS the caller pins once per kernel: an FP64 instruction takes one
// non-register operand, so fma(w, c1, CF_MAGIC) with both as constants made the compiler re-materialise one of them per use (4 extra
// uniform-to-register moves per pair in the first version).
template <int R>
__device__ __forceinline__ void cf_matern_fast_n(const double (&r2)[R], const cf_atom_val& A, const double c1, const double g0, const double g1,
                                                 const double g2, const double magic, const double top, cf_tbl_t tbl_lane, double (&kv)[R]) {
    constexpr double g3 = 0x1.3b2ab00000000p-39;
    double w[R], h[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        // the seed instruction (MUFU.RSQ64H) reads only the HIGH word: it gets max(|r2|, 2^-1007) formed with two integer instructions (so
        // that r2 = 0 yields a finite y and g = 0 * y = 0), everything else uses |r2| as an operand modifier of the FP64 instructions
        const int hi = max(__double2hiint(r2[r]) & 0x7fffffff, 0x01000000);
        const double v = fabs(r2[r]);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(hi, 0)));
        double g = v * y;
        h[r] = 0.5 * y;
        const double e = fma(-h[r], g, 0.5);
        g = fma(g, e, g);                    // relative error ~2^-44
        const double dd = fma(-g, g, v);
        w[r] = fma(dd, h[r], g);             // h is the unrefined y / 2: its 2^-22 error multiplies dd ~ 2^-44 v
    }
    double t[R], f[R], s[R], q[R];
#pragma unroll
    for (int r = 0; r < R; r++) t[r] = fma(w[r], c1, magic);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double kd = t[r] - magic;
        f[r] = fma(w[r], c1, -kd);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int kk = __double2loint(t[r]);
        int addr;
        asm("mad.lo.s32 %0, %1, 128, %2;" : "=r"(addr) : "r"(kk & (CF_EXP_TBL - 1)), "r"((int)tbl_lane));
        double tj;
        asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));
        int hi;
        asm("mad.lo.s32 %0, %1, 4096, %2;" : "=r"(hi) : "r"(kk), "r"(__double2hiint(tj)));
        s[r] = __hiloint2double(hi, __double2loint(tj));
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        double p = fma(g3, f[r], g2);
        p = fma(p, f[r], g1);
        p = fma(p, f[r], g0);
        q[r] = fma(f[r], p, 1.0);
    }
    const int pm = A.p;  // >= 1
    double mp[R];
    {
        const double c = A.mat[pm - 1];
#pragma unroll
        for (int r = 0; r < R; r++) mp[r] = fma(top, w[r], c);
    }
#pragma unroll 1
    for (int i = pm - 2; i >= 0; i--) {
        const double ci = A.mat[i];
#pragma unroll
        for (int r = 0; r < R; r++) mp[r] = fma(mp[r], w[r], ci);
    }
#pragma unroll
    for (int r = 0; r < R; r++) kv[r] = (mp[r] * s[r]) * q[r];
}


template <int D, int TJ, int NS>
struct cf_mvme_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;        // 32 KB
    static constexpr int bar_bytes = 128;
    static constexpr int y_bytes = TJ * D * 8;
    static constexpr int v_bytes = TJ * 8;                          // |y|^2 tile, a tile
    static constexpr int stage_bytes = ((y_bytes + 2 * v_bytes + 127) / 128) * 128;
    static constexpr int ring_bytes = bar_bytes + NS * stage_bytes;
    static constexpr int total = tbl_bytes + ring_bytes;
};

template <int D, int R, int NT, int TJ, int NS, int MINB, int FAST = 1>
__global__ void __launch_bounds__(NT, MINB) gram_mvm_eq_kernel(const __grid_constant__ cf_mvm_params P) {
    using S = cf_mvme_smem<D, TJ, NS>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;

    const int tid = threadIdx.x;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const double* __restrict__ Xg = static_cast<const double*>(P.X);
    const double* __restrict__ Yg = static_cast<const double*>(P.Y);
    const double* __restrict__ xng = static_cast<const double*>(P.xn);
    const double* __restrict__ yng = static_cast<const double*>(P.yn);
    const double* __restrict__ ag = static_cast<const double*>(P.a);

    int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    int64_t c1c = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    if (P.diag_block > 0) {
        c0 = ((P.row0 + (int64_t)blockIdx.x * (NT * R)) / P.diag_block) * P.diag_block;
        c1c = (c0 + P.diag_block < P.m) ? c0 + P.diag_block : P.m;
    }
    const int64_t ncols = c1c - c0;
    const int nfull = P.use_tma ? (int)(ncols / TJ) : 0;
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
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + 2 * S::v_bytes));
        cf_tma_load_1d(st, Yg + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, yng + j0, (uint32_t)S::v_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes + S::v_bytes, ag + j0, (uint32_t)S::v_bytes, &bars[s]);
    };
    if (tid == 0) {
        for (int t = 0; t < NS && t < nfull; t++) issue(t);
    }

    // this thread's rows: -2 x_i, the rounding constant M_i, (the leftover row factor is recomputed in the epilogue)
    const double c1s = P.atom.e.c1; // 256 c / ln2
    double xs[R][D], mrow[R];
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * (NT * R);
    const int64_t rend = P.row0 + P.nrows;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        if (i >= rend) i = rend - 1; // clamp: computed but never stored
#pragma unroll
        for (int c = 0; c < D; c++) xs[r][c] = -2.0 * Xg[i * D + c];
        mrow[r] = (FAST == 2) ? xng[i] : rint(P.atom.e.c1 * xng[i]) + CF_MAGIC;  // FAST = 2: |x_i|^2 itself
    }

    double tot[R];
#pragma unroll
    for (int r = 0; r < R; r++) tot[r] = 0.0;

    // exp(f ln2 / 256) = 1 + f (g0 + g1 f + g2 f^2 + g3 f^3), |f| <= 1/2: truncation error < 4e-17.  The constants travel in the
    // kernel parameters (P.eqc, filled by the host) so that they are constant-bank operands of the FMAs, not registers.
    // An FP64 instruction takes at most ONE non-register operand (constant, uniform register or 32-bit immediate), so g3 f + g2 would
    // read two coefficient registers.  g3 only needs 11 bits (its term is < 1.4e-13 of the result): it is rounded to a double whose low
    // word is zero, which the instruction encodes as an immediate; g2 then is the one uniform operand.
    const double g0 = P.eqc[0], g1 = P.eqc[1], g2 = P.eqc[2];
    constexpr double g3 = 0x1.3b2ab00000000p-39; // (ln2 / 256)^4 / 24 to 3.4e-7
    double magic = CF_MAGIC, mtop = (FAST == 2) ? P.atom.mat[P.atom.p] : 0.0;  // FAST = 2: pinned in registers (see cf_matern_fast_n)
    if (FAST == 2) asm volatile("" : "+d"(magic), "+d"(mtop));

    double acc[R];
    // one column against the R rows of this thread
    auto column = [&](const double (&yj)[D], const double nj, const double aj) {
        double i0[R], t[R], f[R], p[R], s[R];
        // coordinate-major: the R consecutive FMAs of a coordinate share y_c (and |y|^2) in the same operand slot, so that all but the
        // first CAN be served by the operand-reuse cache and read two registers.  ptxas interleaves them with other work, though: in the
        // shipped SASS ~16 of the 24 distance FMAs per column still fetch three registers (82 of 384 FP64 instructions per 32 pairs;
        // bench_aux/k1e_model.py).  Source order does not change that (volatile asm statements in this order give identical SASS);
        // grouping them would be worth ~2 of the 31 cycles per pair.
#pragma unroll
        for (int r = 0; r < R; r++) i0[r] = fma(xs[r][D - 1], yj[D - 1], nj);
#pragma unroll
        for (int c = D - 2; c >= 0; c--) {
#pragma unroll
            for (int r = 0; r < R; r++) i0[r] = fma(xs[r][c], yj[c], i0[r]);
        }
        if constexpr (FAST == 2) {
            double r2[R], kv[R];
#pragma unroll
            for (int r = 0; r < R; r++) r2[r] = i0[r] + mrow[r];
            cf_matern_fast_n<R>(r2, P.atom, c1s, g0, g1, g2, magic, mtop, tbl_lane, kv);
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = fma(kv[r], aj, acc[r]);
        } else {
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = fma(i0[r], c1s, mrow[r]);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const double kd = t[r] - mrow[r];
            f[r] = fma(i0[r], c1s, -kd);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int kk = __double2loint(t[r]);
            int addr;
            asm("mad.lo.s32 %0, %1, 128, %2;" : "=r"(addr) : "r"(kk & (CF_EXP_TBL - 1)), "r"((int)tbl_lane));
            double tj;
            asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));
            int hi;
            asm("mad.lo.s32 %0, %1, 4096, %2;" : "=r"(hi) : "r"(kk), "r"(__double2hiint(tj))); // + kk << 12
            s[r] = __hiloint2double(hi, __double2loint(tj));
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            p[r] = fma(g3, f[r], g2);
            p[r] = fma(p[r], f[r], g1);
            p[r] = fma(p[r], f[r], g0);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            // every 64-bit REGISTER source operand of an FP64 instruction costs an operand-delivery cycle (bench_aux/micro/
            // fp64_issue_probe.cu: a three-register DFMA takes 3 cycles, not 2): s (1 + f p) as a two-register FMA and a multiply
            const double q = fma(f[r], p[r], 1.0);
            const double e = s[r] * q;
            acc[r] = fma(e, aj, acc[r]);
        }
        }  // FAST == 1
    };
    // a tile: columns two at a time, every operand load a 16-byte shared-memory broadcast (stage buffers are 128-byte aligned)
    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ ns, const double* __restrict__ as, int cnt) {
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        const double2* __restrict__ ys2 = reinterpret_cast<const double2*>(ys);
        const double2* __restrict__ ns2 = reinterpret_cast<const double2*>(ns);
        const double2* __restrict__ as2 = reinterpret_cast<const double2*>(as);
        const int npair = cnt >> 1;
#ifndef CF_EQ_UNROLL
#define CF_EQ_UNROLL 2
#endif
        constexpr int UNR = CF_EQ_UNROLL;
#pragma unroll UNR
        for (int jj = 0; jj < npair; jj++) {
            double yy[2 * D];
#pragma unroll
            for (int q = 0; q < D; q++) {
                const double2 v = ys2[jj * D + q];
                yy[2 * q] = v.x; yy[2 * q + 1] = v.y;
            }
            const double2 n2 = ns2[jj], a2 = as2[jj];
            double y0[D], y1[D];
#pragma unroll
            for (int c = 0; c < D; c++) { y0[c] = yy[c]; y1[c] = yy[D + c]; }
            column(y0, n2.x, a2.x);
            column(y1, n2.y, a2.y);
        }
        if (cnt & 1) {
            const int j = cnt - 1;
            double yj[D];
#pragma unroll
            for (int c = 0; c < D; c++) yj[c] = ys[j * D + c];
            column(yj, ns[j], as[j]);
        }
#pragma unroll
        for (int r = 0; r < R; r++) tot[r] += acc[r];
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        const uint32_t parity = (uint32_t)((t / NS) & 1);
        cf_mbar_wait(&bars[s], parity);
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                reinterpret_cast<const double*>(st + S::y_bytes + S::v_bytes), TJ);
        __syncthreads();
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1c; j0 += TJ) {
        const int cnt = (int)((c1c - j0 < TJ) ? c1c - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* ns = reinterpret_cast<double*>(stages + S::y_bytes);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes + S::v_bytes);
        __syncthreads();
        for (int q = tid; q < cnt * D; q += NT) ys[q] = Yg[j0 * D + q];
        for (int q = tid; q < cnt; q += NT) { ns[q] = yng[j0 + q]; as[q] = ag[j0 + q]; }
        __syncthreads();
        compute(ys, ns, as, cnt);
    }

    // epilogue: the row's leftover factor, then alpha / beta or the raw partial sum
    double* out = static_cast<double*>(P.out);
    const double* yin = static_cast<const double*>(P.yin);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < rend) {
            const double nx = xng[i];
            const double w = P.atom.e.c1 * nx;
            const double lo = (w - rint(w)) + fma(P.atom.e.c1, nx, -w); // c1 |x|^2 - rint(c1 |x|^2), with the product's rounding error
            const double z = lo * (0.693147180559945309417232121458 / 256.0);    // |z| <= 0.00136
            const double ez = 1.0 + z * (1.0 + z * (0.5 + z * (1.0 / 6.0 + z * (1.0 / 24.0 + z * (1.0 / 120.0)))));
            const double v0 = (FAST == 2) ? tot[r] : tot[r] * ez;
            const int64_t o = i - P.row0;
            if (P.direct) {
                double v = P.alpha * v0;
                if (P.beta != 0.0) v += P.beta * yin[o];
                out[o] = v;
                for (int q = 0; q < P.peers.n; q++) static_cast<double*>(P.peers.ptr[q])[o] = v; // NVLink peer stores
            } else {
                out[(int64_t)blockIdx.y * P.nrows + o] = v0;
            }
        }
    }
}

#ifndef __CUDACC_RTC__
template <int D, int R, int NT, int TJ, int NS, int MINB, int FAST = 1>
cudaError_t cf_mvme_launch(const cf_mvm_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvme_smem<D, TJ, NS>;
    auto kern = gram_mvm_eq_kernel<D, R, NT, TJ, NS, MINB, FAST>;
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

// available for the small padded dimensions (the tensor-core kernel takes D >= 8 under the same scale check).
// Tuning (bench_aux/micro/k1e_variants.cu, n = 2^19): 8 rows per thread in CTAs of 128 threads, 2 CTAs per SM: 1.23e12 pairs/s
// (R = 4, 256 threads: 1.17e12; R = 6: 1.20e12); more rows amortise the column operands, whose first use per column is a
// three-register FMA (3 cycles instead of 2).
template <int D, bool OK = (D <= 6)>
struct cf_mvme_entry {
    static constexpr int R = (D <= 4) ? 8 : 4, NT = 128, TJ = 128, NS = 3, MINB = 2;
    static constexpr cf_mvm_launch_fn fn = &cf_mvme_launch<D, R, NT, TJ, NS, MINB>;
    static constexpr cf_mvm_config cfg = {NT * R, TJ, cf_mvme_smem<D, TJ, NS>::total, MINB};
};
template <int D>
struct cf_mvme_entry<D, false> {
    static constexpr cf_mvm_launch_fn fn = nullptr;
    static constexpr cf_mvm_config cfg = {0, 0, 0, 0};
};
// the MaternP form (FAST = 2), padded D <= 8 (D = 8 serves the symmetric variant: the plain product at D >= 8 takes the tensor-core kernel).
// Rows per thread measured with MaternP(2) (bench_aux/micro/k1m_variants.sh; d = 3 at n = 16384 / 131072, d = 8 at n = 131072):
// R = 8 / 4 (254 and 242 registers, no spills): 0.453 / 14.46 / 18.92 ms; R = 6 / 3: 0.486 / 14.77 / 19.74; R = 4 / 2: 0.455 / 15.72 / 21.87.
#ifndef CF_MVMM_R
#define CF_MVMM_R 8
#endif
#ifndef CF_MVMM_R8
#define CF_MVMM_R8 4
#endif
template <int D, bool OK = (D <= 8)>
struct cf_mvmm_entry {
    static constexpr int R = (D <= 4) ? CF_MVMM_R : (D <= 6 ? 4 : CF_MVMM_R8), NT = 128, TJ = 128, NS = 3, MINB = 2;
    static constexpr cf_mvm_launch_fn fn = &cf_mvme_launch<D, R, NT, TJ, NS, MINB, 2>;
    static constexpr cf_mvm_config cfg = {NT * R, TJ, cf_mvme_smem<D, TJ, NS>::total, MINB};
};
template <int D>
struct cf_mvmm_entry<D, false> {
    static constexpr cf_mvm_launch_fn fn = nullptr;
    static constexpr cf_mvm_config cfg = {0, 0, 0, 0};
};
#endif
