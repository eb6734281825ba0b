// cf_math.cuh -- device arithmetic for the Gramian kernels (sm_100a).
//
// FP64: there is no FP64 SFU path on the chip (MUFU.EX2 is FP32-only), so exp is built from the FP64
// FMA pipe: one magic-number rounding, a one-step Cody-Waite reduction against a 256-entry 2^(j/256)
// table held in shared memory (replicated 16x so that every lane of a half-warp reads its own bank
// pair: zero bank conflicts for any index pattern) and a degree-4 polynomial -- 9 FP64 issue slots
// instead of the ~16-20 of a table-free exp.  sqrt and reciprocal take their seed from the SFU
// (MUFU.RSQ64H / MUFU.RCP64H through rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64) and are finished with
// FMA Newton steps.  Accuracy is measured in tests/test_gpu_math.py: sqrt / reciprocal <= 1.5 ulp, exp <= 1.4 ulp with the
// two-step reduction (CF_EXP_ACCURATE) and <= (2 + 0.35 |argument|) ulp with the one-step reduction of the hot loop.
// FP32: ex2.approx / rsqrt.approx / rcp.approx on the SFU.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "cf_program.h"

#define CF_EXP_TBL_BITS 8
#define CF_EXP_TBL (1 << CF_EXP_TBL_BITS)  /* 256 entries: |reduced argument| <= ln2/512, degree-4 polynomial */
#define CF_EXP_TBL_REP 16
#define CF_EXP_TBL_DOUBLES (CF_EXP_TBL * CF_EXP_TBL_REP)  /* 32 KB of shared memory */
#define CF_MAGIC 6755399441055744.0 /* 1.5 * 2^52 */

typedef uint32_t cf_tbl_t; // shared-space byte address of this lane's replica column of the exp table
__device__ __forceinline__ uint32_t cf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ cf_tbl_t cf_tbl_lane(const double* tbl, int tid) { return cf_smem_u32(tbl + (tid & 15)); }
// The table is read with `ld.shared` inside non-volatile asm (so that the loads can be scheduled freely among the FMA
// chains).  The compiler therefore does not know they read memory and could hoist them above the barrier that publishes the
// table (compute-sanitizer racecheck caught exactly that).  Every kernel calls this right after that barrier: it makes the
// table address -- an input of every table load -- the output of a volatile asm that cannot cross the barrier.
__device__ __forceinline__ void cf_tbl_publish(cf_tbl_t& tbl_lane) { asm volatile("" : "+r"(tbl_lane) : : "memory"); }

// Copy the 2^(j/256) table (256 doubles in global memory, written once per device by the host, correctly
// rounded from long double) into shared memory, 16 replicas per entry: entry j for lane l lives at
// tbl[j*16 + (l & 15)], so the 16 lanes of a half-warp always hit 16 different bank pairs.
__device__ __forceinline__ void cf_fill_exp_table(double* tbl, const double* __restrict__ g_tbl, int tid, int nthreads) {
    for (int i = tid; i < CF_EXP_TBL_DOUBLES; i += nthreads) tbl[i] = g_tbl[i >> 4];
}

// exp(c*v), v >= 0 (c < 0 folded into the constants).  tbl_lane = table + (lane & 15).
// The SM issues one instruction per cycle per sub-partition and an FP64 instruction holds the dispatch port for two
// (measured: cycles = 2 * #FP64 + #other, profiles/), so the non-FP64 work is kept to 5 instructions per call:
// the clamp of v is ONE integer min on the high word (v >= 0, so integer order of high words == floating-point
// order; the clamped value lies within 2^-20 relative of E.vmax, where the result is ~1e-304, i.e. 0), the table
// address is AND + multiply-add, the 2^k scaling is one multiply-add on the high word of the (pre-compensated) table entry.
__device__ __forceinline__ double cf_exp_cv(double v, const cf_exp_consts& E, cf_tbl_t tbl_lane) {
    v = __hiloint2double(min(__double2hiint(v), E.vmax_hi), __double2loint(v));
    double t = fma(v, E.c1, CF_MAGIC);
    const int kk = __double2loint(t);
    double kd = t - CF_MAGIC;
#ifdef CF_EXP_ACCURATE
    // two-step Cody-Waite (kd * c2_hi is exact): <= ~1 ulp for every argument; used by the dense-instantiation and d > 32
    // translation unit.  The one-step form below adds a relative error of about 0.3 |c v| ulp (the size of the rounding
    // error r2 itself carries), i.e. an absolute error <= 4e-17 relative to k = 1 -- invisible in a sum, and 2 cycles cheaper.
    double u = fma(kd, E.c2_lo, fma(kd, E.c2_hi, v));
#else
    double u = fma(kd, E.c2, v);
#endif
    double p = fma(E.q[3], u, E.q[2]);
    p = fma(p, u, E.q[1]);
    p = fma(p, u, E.q[0]);
    int off;
    double tj;
    asm("mad.lo.s32 %0, %1, 128, %2;" : "=r"(off) : "r"(kk & (CF_EXP_TBL - 1)), "r"((int)tbl_lane));
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(off)); // table is written once, before the first barrier
    // 2^k scaling in ONE integer instruction: the table stores 2^(j/256) with (j << 12) already subtracted from its high word, so that
    // adding kk << 12 = (k << 20) + (j << 12) leaves exactly k in the exponent field (capi.cu get_ctx builds the table)
    int hi;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(kk), "r"(0x100000 >> CF_EXP_TBL_BITS), "r"(__double2hiint(tj)));
    const double s = __hiloint2double(hi, __double2loint(tj)); // 2^k 2^(j/256)
    const double su = s * u;
    return fma(su, p, s); // s * (1 + u p)
}

// sqrt(v) for v >= 0; v below 2^-1007 (incl. 0 and subnormals) returns ~2^-504 (i.e. 0 for our purposes).
__device__ __forceinline__ double cf_sqrt_pos(double v) {
    int hi = __double2hiint(v);
    if (hi < 0x01000000) v = __hiloint2double(0x01000000, 0);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    double g = v * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double dd = fma(-g, g, v);
    return fma(dd, h, g);
}

// 1/b for normal b (here b >= 1)
__device__ __forceinline__ double cf_rcp(double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    return y;
}

__device__ __forceinline__ double cf_powi(double b, int p) { // p >= 1, repeated multiplication (uniform p)
    double r = b;
    for (int i = 1; i < p; i++) r *= b;
    return r;
}

// ---- atom values (FP64) ------------------------------------------------------------------------
__device__ __forceinline__ double cf_atom_eq(double r2, const cf_atom_val& A, cf_tbl_t tbl_lane) {
    return cf_exp_cv(r2, A.e, tbl_lane);
}
// M(g) exp(c g), g = sqrt(r2).  The reference's Taylor branch (src/stationary.jl:139-146) differs from this
// closed form by < 1e-17 relative for every p, so the value path does not branch.
__device__ __forceinline__ double cf_clamp_v(double v, const cf_exp_consts& E) {
    return __hiloint2double(min(__double2hiint(v), E.vmax_hi), __double2loint(v));
}
__device__ __forceinline__ double cf_atom_matern(double r2, const cf_atom_val& A, cf_tbl_t tbl_lane) {
    double g = cf_clamp_v(cf_sqrt_pos(r2), A.e); // clamp BEFORE the polynomial: M(g) e^{cg} with g ~ 1e150 must be 0, not M(g) * 1e-304
    double e = cf_exp_cv(g, A.e, tbl_lane);
    int p = A.p;
    if (p == 0) return e;
    double mp = A.mat[p];
    for (int i = p - 1; i >= 0; i--) mp = fma(mp, g, A.mat[i]);
    return mp * e;
}
// N pairs at once: every stage (sqrt, exp, Horner step) is issued for all N values before the next one, so the N
// dependent chains interleave and the runtime loop over p costs one branch per N values
// PS >= 0: the integer parameter is a compile-time constant (run-time specialised builds, cf_jit.h): loops unroll fully
template <int N, int PS = -1>
__device__ __forceinline__ void cf_atom_matern_n(const double (&r2)[N], const cf_atom_val& A, cf_tbl_t tbl_lane, double (&out)[N]) {
    double g[N], e[N];
#pragma unroll
    for (int u = 0; u < N; u++) g[u] = cf_clamp_v(cf_sqrt_pos(r2[u]), A.e);
#pragma unroll
    for (int u = 0; u < N; u++) e[u] = cf_exp_cv(g[u], A.e, tbl_lane);
    const int p = (PS >= 0) ? PS : A.p;
    if (p == 0) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = e[u];
        return;
    }
    double mp[N];
#pragma unroll
    for (int u = 0; u < N; u++) mp[u] = A.mat[p];
    if constexpr (PS >= 0) {
#pragma unroll
        for (int i = PS - 1; i >= 0; i--) {
            const double ci = A.mat[i];
#pragma unroll
            for (int u = 0; u < N; u++) mp[u] = fma(mp[u], g[u], ci);
        }
    } else {
#pragma unroll 1
        for (int i = p - 1; i >= 0; i--) {
            const double ci = A.mat[i];
#pragma unroll
            for (int u = 0; u < N; u++) mp[u] = fma(mp[u], g[u], ci);
        }
    }
#pragma unroll
    for (int u = 0; u < N; u++) out[u] = mp[u] * e[u];
}
// (1 + w r2)^-a, a a positive Int: one reciprocal of base^a (a - 1 multiplications), clamped so that an overflowing power
// returns ~0 instead of feeding Inf to the Newton steps
__device__ __forceinline__ double cf_atom_rq_int(double r2, const cf_atom_val& A) {
    double base = fma(r2, A.w, 1.0);
    return cf_rcp(fmin(cf_powi(base, A.p), 1e300));
}
template <int N, int PS = -1>
__device__ __forceinline__ void cf_atom_rq_int_n(const double (&r2)[N], const cf_atom_val& A, double (&out)[N]) {
    double base[N], pw[N];
#pragma unroll
    for (int u = 0; u < N; u++) { base[u] = fma(r2[u], A.w, 1.0); pw[u] = base[u]; }
    const int p = (PS >= 0) ? PS : A.p;
    if constexpr (PS >= 0) {
#pragma unroll
        for (int i = 1; i < PS; i++) {
#pragma unroll
            for (int u = 0; u < N; u++) pw[u] *= base[u];
        }
    } else {
#pragma unroll 1
        for (int i = 1; i < p; i++) {
#pragma unroll
            for (int u = 0; u < N; u++) pw[u] *= base[u];
        }
    }
#pragma unroll
    for (int u = 0; u < N; u++) out[u] = cf_rcp(fmin(pw[u], 1e300));
}
// pow() is a large routine: keep one out-of-line copy per kernel instead of one per call site (instruction cache)
static __device__ __noinline__ double cf_pow_outlined(double base, double e) { return pow(base, e); }
__device__ __forceinline__ double cf_atom_rq_real(double r2, const cf_atom_val& A) {
    double base = fma(r2, A.w, 1.0);
    return cf_pow_outlined(base, -A.alpha);
}

template <int KIND>
__device__ __forceinline__ double cf_atom_value(double r2, double dt, const cf_atom_val& A, cf_tbl_t tbl_lane) {
    if (KIND == CF_ATOM_EQ) return cf_atom_eq(r2, A, tbl_lane);
    if (KIND == CF_ATOM_MATERN) return cf_atom_matern(r2, A, tbl_lane);
    if (KIND == CF_ATOM_RQ_INT) return cf_atom_rq_int(r2, A);
    if (KIND == CF_ATOM_RQ_REAL) return cf_atom_rq_real(r2, A);
    return dt + A.sigma; // LINE
}

__device__ __forceinline__ double cf_atom_value_dyn(double r2, double dt, const cf_atom_val& A, cf_tbl_t tbl_lane) {
    switch (A.kind) {
        case CF_ATOM_EQ: return cf_atom_eq(r2, A, tbl_lane);
        case CF_ATOM_MATERN: return cf_atom_matern(r2, A, tbl_lane);
        case CF_ATOM_RQ_INT: return cf_atom_rq_int(r2, A);
        case CF_ATOM_RQ_REAL: return cf_atom_rq_real(r2, A);
        default: return dt + A.sigma;
    }
}

// generic sum of products: the program lives in the kernel parameters (constant bank) and control flow is warp-uniform.
// N pairs are evaluated together so that the interpreter's branches and loop counters are paid once per N values.
template <int N>
__device__ __forceinline__ void cf_atom_value_dyn_n(const double (&r2)[N], const double (&dt)[N], const cf_atom_val& A,
                                                    cf_tbl_t tbl_lane, double (&out)[N]) {
    switch (A.kind) {
        case CF_ATOM_EQ:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = cf_atom_eq(r2[u], A, tbl_lane);
            break;
        case CF_ATOM_MATERN:
            cf_atom_matern_n<N>(r2, A, tbl_lane, out);
            break;
        case CF_ATOM_RQ_INT:
            cf_atom_rq_int_n<N>(r2, A, out);
            break;
        case CF_ATOM_RQ_REAL:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = cf_atom_rq_real(r2[u], A);
            break;
        default:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = dt[u] + A.sigma;
    }
}
// out = atom^PW with the atom kind, its integer parameter and the power known at compile time
template <int N, int KIND, int PS, int PW>
__device__ __forceinline__ void cf_atom_pow_s(const double (&r2)[N], const double (&dt)[N], const cf_atom_val& A, cf_tbl_t tbl_lane,
                                              double (&out)[N]) {
    if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = cf_atom_eq(r2[u], A, tbl_lane);
    } else if constexpr (KIND == CF_ATOM_MATERN) {
        cf_atom_matern_n<N, PS>(r2, A, tbl_lane, out);
    } else if constexpr (KIND == CF_ATOM_RQ_INT) {
        cf_atom_rq_int_n<N, PS>(r2, A, out);
    } else if constexpr (KIND == CF_ATOM_RQ_REAL) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = cf_atom_rq_real(r2[u], A);
    } else {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = dt[u] + A.sigma;
    }
    if constexpr (PW > 1) {
        double a[N];
#pragma unroll
        for (int u = 0; u < N; u++) a[u] = out[u];
#pragma unroll
        for (int q = 1; q < PW; q++) {
#pragma unroll
            for (int u = 0; u < N; u++) out[u] *= a[u];
        }
    }
}

#ifdef CF_JIT_SHAPE
// run-time specialised build (capi.cu, cf_jit.h): the generated header defines cf_sop_value_n<N> for ONE program structure --
// atom kinds, integer parameters, powers and the term list are compile-time constants, only the coefficients and atom
// parameters stay in the kernel arguments.  Same signature and same results as the interpreter below.
#define CF_JIT_PART 1  // the Float64 evaluator
#include "cf_jit_shape.h"
#undef CF_JIT_PART
#else
// out = atom^pw for N pairs (pw >= 1; the common pw = 1, 2 cost no copies)
template <int N>
__device__ __forceinline__ void cf_atom_pow_n(const double (&r2)[N], const double (&dt)[N], const cf_atom_val& A, int pw,
                                              cf_tbl_t tbl_lane, double (&out)[N]) {
    cf_atom_value_dyn_n<N>(r2, dt, A, tbl_lane, out);
    if (pw == 2) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] *= out[u];
    } else if (pw > 2) {
        double a[N];
#pragma unroll
        for (int u = 0; u < N; u++) a[u] = out[u];
#pragma unroll 1
        for (int q = 1; q < pw; q++) {
#pragma unroll
            for (int u = 0; u < N; u++) out[u] *= a[u];
        }
    }
}
// value = sum_t coef_t prod_f atom^pw: the coefficient is applied with the final FMA, the first factor initialises the product
template <int N>
__device__ __forceinline__ void cf_sop_value_n(const double (&r2)[N], const double (&dt)[N], const cf_sop_val& P,
                                               cf_tbl_t tbl_lane, double (&val)[N]) {
#pragma unroll
    for (int u = 0; u < N; u++) val[u] = 0.0;
    const int nt = P.nterms;
    for (int t = 0; t < nt; t++) {
        const cf_sop_term& T = P.terms[t];
        const int nf = T.nfac;
        if (nf == 0) { // constant term
#pragma unroll
            for (int u = 0; u < N; u++) val[u] += T.coef;
            continue;
        }
        double prod[N];
        cf_atom_pow_n<N>(r2, dt, P.atoms[T.atom[0]], T.power[0], tbl_lane, prod);
        for (int f = 1; f < nf; f++) {
            double a[N];
            cf_atom_pow_n<N>(r2, dt, P.atoms[T.atom[f]], T.power[f], tbl_lane, a);
#pragma unroll
            for (int u = 0; u < N; u++) prod[u] *= a[u];
        }
#pragma unroll
        for (int u = 0; u < N; u++) val[u] = fma(T.coef, prod[u], val[u]);
    }
}
#endif // CF_JIT_SHAPE
__device__ __forceinline__ double cf_sop_value(double r2, double dt, const cf_sop_val& P, cf_tbl_t tbl_lane) {
    double a[1] = {r2}, b[1] = {dt}, v[1];
    cf_sop_value_n<1>(a, b, P, tbl_lane, v);
    return v[0];
}

// ---- derivatives with respect to r2 (gradient kernel; reference src/gradient.jl:589-600) -------------
// returns k, k1 = dk/dr2, k2 = d2k/dr2^2 of one isotropic atom
__device__ __forceinline__ void cf_atom_jet(double r2, const cf_atom& A, cf_tbl_t tbl_lane, double& k, double& k1, double& k2) {
    switch (A.v.kind) {
        case CF_ATOM_EQ: {
            k = cf_exp_cv(r2, A.v.e, tbl_lane);
            k1 = A.v.e.c * k;
            k2 = A.v.e.c * k1;
            return;
        }
        case CF_ATOM_MATERN: {
            const int p = A.v.p;
            const double s = r2 * A.inv_l2; // the inner kernel sees r2 / l^2 (reference src/transformation.jl:19)
            if (s < A.taylor_bound) { // reference src/stationary.jl:139-146 (differentiated through by ForwardDiff)
                double v = 0, d1 = 0, d2 = 0;
                for (int i = p; i >= 0; i--) { // Horner with derivatives
                    d2 = fma(d2, s, 2.0 * d1);
                    d1 = fma(d1, s, v);
                    v = fma(v, s, A.tay[i]);
                }
                k = v; k1 = d1 * A.inv_l2; k2 = d2 * A.inv_l2 * A.inv_l2;
                return;
            }
            double g = cf_clamp_v(cf_sqrt_pos(r2), A.v.e);
            double e = cf_exp_cv(g, A.v.e, tbl_lane);
            double m = A.v.mat[p], a = (p >= 1) ? A.matA[p - 1] : 0.0, b = (p >= 2) ? A.matB[p - 2] : 0.0;
            for (int i = p - 1; i >= 0; i--) m = fma(m, g, A.v.mat[i]);
            for (int i = p - 2; i >= 0; i--) a = fma(a, g, A.matA[i]);
            for (int i = p - 3; i >= 0; i--) b = fma(b, g, A.matB[i]);
            if (A.am1 != 0.0 || A.bm[0] != 0.0 || A.bm[1] != 0.0 || A.bm[2] != 0.0) { // p <= 1: singular terms
                // coincident points (r2 == 0 exactly, p = 0: Exp has no Taylor branch): the reference differentiates exp(-sqrt(r2)) at 0
                // with ForwardDiff and gets k' = -Inf, so that its block -2 (k' a + 2 k'' r (r.a)) is NaN (Inf * 0); 1 / 0 = Inf here
                // reproduces that instead of the large finite value the clamped square root would give
                double gi = (r2 == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : 1.0 / g;
                a = fma(A.am1, gi, a);
                b += gi * (A.bm[0] + gi * (A.bm[1] + gi * A.bm[2]));
            }
            k = m * e; k1 = a * e; k2 = b * e;
            return;
        }
        case CF_ATOM_RQ_INT:
        case CF_ATOM_RQ_REAL: {
            double base = fma(r2, A.v.w, 1.0);
            double ib = 1.0 / base;
            k = (A.v.kind == CF_ATOM_RQ_INT) ? cf_powi(ib, A.v.p) : cf_pow_outlined(base, -A.v.alpha);
            k1 = -A.v.alpha * A.v.w * k * ib;
            k2 = -(A.v.alpha + 1.0) * A.v.w * k1 * ib;
            return;
        }
        default: k = r2 + A.v.sigma; k1 = 1.0; k2 = 0.0; return; // LINE: the variable is t = x.y (DotProductInput programs)
    }
}

// element u of a register array, u a run-time index: a chain of selects (register arrays cannot be indexed dynamically)
template <int N>
__device__ __forceinline__ double cf_select_n(const double (&a)[N], int u) {
    double v = a[0];
#pragma unroll
    for (int q = 1; q < N; q++) v = (u == q) ? a[q] : v;
    return v;
}
template <int N>
__device__ __forceinline__ void cf_store_n(double (&a)[N], int u, double v) {
#pragma unroll
    for (int q = 0; q < N; q++) a[q] = (u == q) ? v : a[q];
}

// jets of N pairs of ONE Matern atom with p >= 2 (k' and k'' have no singular terms): every stage is issued for all N values, the
// loops over p are uniform.  The reference's Taylor branch for r2 / l^2 < eps^(1/p) (src/stationary.jl:139-146) is evaluated
// only by the lanes that have such an entry (coincident points) and selected per entry.
template <int N>
__device__ __forceinline__ void cf_matern_jet_n(const double (&r2)[N], const cf_atom& A, cf_tbl_t tbl_lane, double (&k)[N],
                                                double (&k1)[N], double (&k2)[N]) {
    double g[N], e[N], m[N], a[N], b[N];
#pragma unroll
    for (int u = 0; u < N; u++) g[u] = cf_clamp_v(cf_sqrt_pos(r2[u]), A.v.e);
#pragma unroll
    for (int u = 0; u < N; u++) e[u] = cf_exp_cv(g[u], A.v.e, tbl_lane);
    const int p = A.v.p;
#pragma unroll
    for (int u = 0; u < N; u++) { m[u] = A.v.mat[p]; a[u] = A.matA[p - 1]; b[u] = A.matB[p - 2]; }
#pragma unroll 1
    for (int i = p - 1; i >= 0; i--) {
        const double ci = A.v.mat[i];
#pragma unroll
        for (int u = 0; u < N; u++) m[u] = fma(m[u], g[u], ci);
    }
#pragma unroll 1
    for (int i = p - 2; i >= 0; i--) {
        const double ci = A.matA[i];
#pragma unroll
        for (int u = 0; u < N; u++) a[u] = fma(a[u], g[u], ci);
    }
#pragma unroll 1
    for (int i = p - 3; i >= 0; i--) {
        const double ci = A.matB[i];
#pragma unroll
        for (int u = 0; u < N; u++) b[u] = fma(b[u], g[u], ci);
    }
    bool any_small = false;
#pragma unroll
    for (int u = 0; u < N; u++) {
        k[u] = m[u] * e[u]; k1[u] = a[u] * e[u]; k2[u] = b[u] * e[u];
        any_small |= (r2[u] * A.inv_l2 < A.taylor_bound);
    }
    if (any_small) {
        double s[N], v[N], d1[N], d2[N];
#pragma unroll
        for (int u = 0; u < N; u++) { s[u] = r2[u] * A.inv_l2; v[u] = 0.0; d1[u] = 0.0; d2[u] = 0.0; }
#pragma unroll 1
        for (int i = p; i >= 0; i--) {  // Horner with derivatives, as in cf_atom_jet
            const double ti = A.tay[i];
#pragma unroll
            for (int u = 0; u < N; u++) {
                d2[u] = fma(d2[u], s[u], 2.0 * d1[u]);
                d1[u] = fma(d1[u], s[u], v[u]);
                v[u] = fma(v[u], s[u], ti);
            }
        }
#pragma unroll
        for (int u = 0; u < N; u++)
            if (s[u] < A.taylor_bound) { k[u] = v[u]; k1[u] = d1[u] * A.inv_l2; k2[u] = d2[u] * A.inv_l2 * A.inv_l2; }
    }
}

// jets of N pairs of one isotropic atom, dispatching on the kind ONCE for the N values.  MaternP(p < 2) (singular k'') is
// handled entry by entry through cf_atom_jet: the tensor-core gradient kernel never sees it, the scalar one may.
__device__ __forceinline__ void cf_atom_jet(double r2, const cf_atom& A, cf_tbl_t tbl_lane, double& k, double& k1, double& k2);
template <int N>
__device__ __forceinline__ void cf_atom_jet_n(const double (&r2)[N], const cf_atom& A, cf_tbl_t tbl_lane, double (&k)[N],
                                              double (&k1)[N], double (&k2)[N]) {
    const int kind = A.v.kind;
    if (kind == CF_ATOM_EQ) {
#pragma unroll
        for (int u = 0; u < N; u++) {
            k[u] = cf_exp_cv(r2[u], A.v.e, tbl_lane);
            k1[u] = A.v.e.c * k[u];
            k2[u] = A.v.e.c * k1[u];
        }
    } else if (kind == CF_ATOM_MATERN && A.v.p >= 2) {
        cf_matern_jet_n<N>(r2, A, tbl_lane, k, k1, k2);
    } else if (kind == CF_ATOM_RQ_INT) {
        double ib[N];
#pragma unroll
        for (int u = 0; u < N; u++) { ib[u] = cf_rcp(fma(r2[u], A.v.w, 1.0)); k[u] = ib[u]; }
#pragma unroll 1
        for (int i = 1; i < A.v.p; i++) {
#pragma unroll
            for (int u = 0; u < N; u++) k[u] *= ib[u];
        }
        const double c1 = -A.v.alpha * A.v.w, c2 = -(A.v.alpha + 1.0) * A.v.w;
#pragma unroll
        for (int u = 0; u < N; u++) {
            k1[u] = c1 * k[u] * ib[u];
            k2[u] = c2 * k1[u] * ib[u];
        }
    } else {
#pragma unroll 1
        for (int u = 0; u < N; u++) {  // rolled: one copy of the scalar code (real-power RQ, MaternP(p < 2), LINE)
            double kk, kk1, kk2;
            cf_atom_jet(cf_select_n<N>(r2, u), A, tbl_lane, kk, kk1, kk2);
            cf_store_n<N>(k, u, kk); cf_store_n<N>(k1, u, kk1); cf_store_n<N>(k2, u, kk2);
        }
    }
}

// compile-time specialised jets for the common single-atom gradient kernels (no switch, no inlined dead paths)
template <int KIND>
__device__ __forceinline__ void cf_atom_jet_t(double r2, const cf_atom& A, cf_tbl_t tbl_lane, double& k, double& k1, double& k2) {
    if constexpr (KIND == CF_ATOM_EQ) {
        k = cf_exp_cv(r2, A.v.e, tbl_lane);
        k1 = A.v.e.c * k;
        k2 = A.v.e.c * k1;
    } else {
        cf_atom_jet(r2, A, tbl_lane, k, k1, k2);
    }
}

// jets of the generic isotropic sum of products: product rule over factors, powers by repeated multiplication
__device__ __forceinline__ void cf_sop_jet(double r2, const cf_sop_grad& P, cf_tbl_t tbl_lane, double& k, double& k1, double& k2) {
    double sv = 0, s1 = 0, s2 = 0;
    for (int t = 0; t < P.nterms; t++) {
        const cf_sop_term& T = P.terms[t];
        double pv = T.coef, p1 = 0, p2 = 0;
        for (int f = 0; f < T.nfac; f++) {
            double av, a1, a2;
            cf_atom_jet(r2, P.atoms[T.atom[f]], tbl_lane, av, a1, a2);
            for (int q = 0; q < T.power[f]; q++) {
                double nv = pv * av;
                double n1 = fma(p1, av, pv * a1);
                double n2 = fma(p2, av, fma(2.0 * p1, a1, pv * a2));
                pv = nv; p1 = n1; p2 = n2;
            }
        }
        sv += pv; s1 += p1; s2 += p2;
    }
    k = sv; k1 = s1; k2 = s2;
}

// jets of N pairs of one atom whose kind (and integer parameter) is a compile-time constant (run-time specialised builds)
template <int N, int KIND, int PS>
__device__ __forceinline__ void cf_atom_jet_s(const double (&r2)[N], const cf_atom& A, cf_tbl_t tbl_lane, double (&k)[N],
                                              double (&k1)[N], double (&k2)[N]) {
    if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
        for (int u = 0; u < N; u++) {
            k[u] = cf_exp_cv(r2[u], A.v.e, tbl_lane);
            k1[u] = A.v.e.c * k[u];
            k2[u] = A.v.e.c * k1[u];
        }
    } else if constexpr (KIND == CF_ATOM_MATERN && PS >= 2) {
        cf_matern_jet_n<N>(r2, A, tbl_lane, k, k1, k2);
    } else if constexpr (KIND == CF_ATOM_RQ_INT) {
        const double c1 = -A.v.alpha * A.v.w, c2 = -(A.v.alpha + 1.0) * A.v.w;
#pragma unroll
        for (int u = 0; u < N; u++) {
            const double ib = cf_rcp(fma(r2[u], A.v.w, 1.0));
            double kk = ib;
#pragma unroll
            for (int i = 1; i < PS; i++) kk *= ib;
            k[u] = kk;
            k1[u] = c1 * kk * ib;
            k2[u] = c2 * k1[u] * ib;
        }
    } else {
        cf_atom_jet_n<N>(r2, A, tbl_lane, k, k1, k2);  // real-power RQ, MaternP(p < 2), LINE: the generic N-wide form
    }
}

#ifdef CF_JIT_SHAPE
#define CF_JIT_PART 3  // the jets of the derivative program (or only a declaration when the specialised kernel has none)
#include "cf_jit_shape.h"
#undef CF_JIT_PART
#else
// the same for N pairs at a time: the program is decoded once per N values
template <int N>
__device__ __forceinline__ void cf_sop_jet_n(const double (&r2)[N], const cf_sop_grad& P, cf_tbl_t tbl_lane, double (&k)[N],
                                             double (&k1)[N], double (&k2)[N]) {
#pragma unroll
    for (int u = 0; u < N; u++) { k[u] = 0.0; k1[u] = 0.0; k2[u] = 0.0; }
    for (int t = 0; t < P.nterms; t++) {
        const cf_sop_term& T = P.terms[t];
        double pv[N], p1[N], p2[N];
#pragma unroll
        for (int u = 0; u < N; u++) { pv[u] = T.coef; p1[u] = 0.0; p2[u] = 0.0; }
        for (int f = 0; f < T.nfac; f++) {
            double av[N], a1[N], a2[N];
            cf_atom_jet_n<N>(r2, P.atoms[T.atom[f]], tbl_lane, av, a1, a2);
#pragma unroll 1
            for (int q = 0; q < T.power[f]; q++) {
#pragma unroll
                for (int u = 0; u < N; u++) {
                    const double nv = pv[u] * av[u];
                    const double n1 = fma(p1[u], av[u], pv[u] * a1[u]);
                    const double n2 = fma(p2[u], av[u], fma(2.0 * p1[u], a1[u], pv[u] * a2[u]));
                    pv[u] = nv; p1[u] = n1; p2[u] = n2;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < N; u++) { k[u] += pv[u]; k1[u] += p1[u]; k2[u] += p2[u]; }
    }
}
#endif // CF_JIT_SHAPE (jets)

// ---- FP32 ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cf_ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cf_rsqrtf(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cf_rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cf_lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#define CF_LOG2E_F 1.4426950408889634f

// KIND < 0: dispatch on A.kind at run time (generic sum of products)
template <int KIND>
__device__ __forceinline__ float cf_atom_value_f32(float r2, float dt, const cf_atom_val& A) {
    const int kind = (KIND >= 0) ? KIND : A.kind;
    if (kind == CF_ATOM_EQ) return cf_ex2f(r2 * A.f_clog2e);
    if (kind == CF_ATOM_MATERN) {
        float g = r2 * cf_rsqrtf(fmaxf(r2, 1e-37f)); // sqrt(r2); 0 at r2 = 0
        g = fminf(g, A.f_gmax);                       // far points: M(g) e^{cg} must underflow to 0
        const float e = cf_ex2f(g * A.f_clog2e);
        const int p = A.p;
        if (p == 0) return e;
        float mp = A.f_mat[p];
        for (int i = p - 1; i >= 0; i--) mp = fmaf(mp, g, A.f_mat[i]);
        return mp * e;
    }
    if (kind == CF_ATOM_RQ_INT) {
        const float ib = cf_rcpf(fmaf(r2, A.f_w, 1.0f));
        float r = ib;
        for (int i = 1; i < A.p; i++) r *= ib;
        return r;
    }
    if (kind == CF_ATOM_RQ_REAL) return cf_ex2f(-A.f_alpha * cf_lg2f(fmaf(r2, A.f_w, 1.0f)));
    return dt + A.f_sigma;
}
// N pairs at once, dispatching on the atom kind ONCE (uniform branch): only the code of the kind that is present runs, every
// stage is issued for all N values.  (Per-entry dispatch made ptxas if-convert the chain: every entry executed the rsqrt /
// lg2 / ex2 of kinds that were not there -- profiles/r1_ncu_gram_mm_tf32_c3.md.)
template <int N>
__device__ __forceinline__ void cf_atom_value_f32_n(const float (&r2)[N], const float (&dt)[N], const cf_atom_val& A, float (&out)[N]) {
    switch (A.kind) {
        case CF_ATOM_EQ:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = cf_ex2f(r2[u] * A.f_clog2e);
            break;
        case CF_ATOM_MATERN: {
            float g[N], e[N];
#pragma unroll
            for (int u = 0; u < N; u++) {
                g[u] = fminf(r2[u] * cf_rsqrtf(fmaxf(r2[u], 1e-37f)), A.f_gmax);  // sqrt(r2), 0 at r2 = 0; far points underflow to 0
                e[u] = cf_ex2f(g[u] * A.f_clog2e);
            }
            const int p = A.p;
            float mp[N];
#pragma unroll
            for (int u = 0; u < N; u++) mp[u] = A.f_mat[p];
#pragma unroll 1
            for (int i = p - 1; i >= 0; i--) {
                const float ci = A.f_mat[i];
#pragma unroll
                for (int u = 0; u < N; u++) mp[u] = fmaf(mp[u], g[u], ci);
            }
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = (p == 0) ? e[u] : mp[u] * e[u];
            break;
        }
        case CF_ATOM_RQ_INT: {
            float ib[N];
#pragma unroll
            for (int u = 0; u < N; u++) { ib[u] = cf_rcpf(fmaf(r2[u], A.f_w, 1.0f)); out[u] = ib[u]; }
#pragma unroll 1
            for (int i = 1; i < A.p; i++) {
#pragma unroll
                for (int u = 0; u < N; u++) out[u] *= ib[u];
            }
            break;
        }
        case CF_ATOM_RQ_REAL:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = cf_ex2f(-A.f_alpha * cf_lg2f(fmaf(r2[u], A.f_w, 1.0f)));
            break;
        default:
#pragma unroll
            for (int u = 0; u < N; u++) out[u] = dt[u] + A.f_sigma;
    }
}
// out = atom^PW in Float32 with the atom kind, its integer parameter and the power known at compile time (run-time specialised builds)
template <int N, int KIND, int PS, int PW>
__device__ __forceinline__ void cf_atom_pow_f32_s(const float (&r2)[N], const float (&dt)[N], const cf_atom_val& A, float (&out)[N]) {
    if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = cf_ex2f(r2[u] * A.f_clog2e);
    } else if constexpr (KIND == CF_ATOM_MATERN) {
#pragma unroll
        for (int u = 0; u < N; u++) {
            const float g = fminf(r2[u] * cf_rsqrtf(fmaxf(r2[u], 1e-37f)), A.f_gmax);
            const float e = cf_ex2f(g * A.f_clog2e);
            float mp = A.f_mat[PS];
#pragma unroll
            for (int i = PS - 1; i >= 0; i--) mp = fmaf(mp, g, A.f_mat[i]);
            out[u] = (PS == 0) ? e : mp * e;
        }
    } else if constexpr (KIND == CF_ATOM_RQ_INT) {
#pragma unroll
        for (int u = 0; u < N; u++) {
            const float ib = cf_rcpf(fmaf(r2[u], A.f_w, 1.0f));
            float r = ib;
#pragma unroll
            for (int i = 1; i < PS; i++) r *= ib;
            out[u] = r;
        }
    } else if constexpr (KIND == CF_ATOM_RQ_REAL) {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = cf_ex2f(-A.f_alpha * cf_lg2f(fmaf(r2[u], A.f_w, 1.0f)));
    } else {
#pragma unroll
        for (int u = 0; u < N; u++) out[u] = dt[u] + A.f_sigma;
    }
    if constexpr (PW > 1) {
#pragma unroll
        for (int u = 0; u < N; u++) {
            const float a = out[u];
#pragma unroll
            for (int q = 1; q < PW; q++) out[u] *= a;
        }
    }
}
#ifdef CF_JIT_SHAPE
#define CF_JIT_PART 2  // the Float32 evaluator (same generated header, second part)
#include "cf_jit_shape.h"
#undef CF_JIT_PART
#else
// value = sum_t coef_t prod_f atom^pw (same structure as the Float64 interpreter: first factor initialises the product,
// coefficient in the final FMA)
template <int N>
__device__ __forceinline__ void cf_sop_value_f32_n(const float (&r2)[N], const float (&dt)[N], const cf_sop_val& P, float (&val)[N]) {
#pragma unroll
    for (int u = 0; u < N; u++) val[u] = 0.f;
    const int nt = P.nterms;
    for (int t = 0; t < nt; t++) {
        const cf_sop_term& T = P.terms[t];
        const float coef = (float)T.coef;
        const int nf = T.nfac;
        if (nf == 0) {
#pragma unroll
            for (int u = 0; u < N; u++) val[u] += coef;
            continue;
        }
        float prod[N];
        for (int f = 0; f < nf; f++) {
            float a[N];
            cf_atom_value_f32_n<N>(r2, dt, P.atoms[T.atom[f]], a);
            const int pw = T.power[f];
            if (pw >= 2) {
                float b[N];
#pragma unroll
                for (int u = 0; u < N; u++) b[u] = a[u];
#pragma unroll 1
                for (int q = 1; q < pw; q++) {
#pragma unroll
                    for (int u = 0; u < N; u++) a[u] *= b[u];
                }
            }
#pragma unroll
            for (int u = 0; u < N; u++) prod[u] = (f == 0) ? a[u] : prod[u] * a[u];
        }
#pragma unroll
        for (int u = 0; u < N; u++) val[u] = fmaf(coef, prod[u], val[u]);
    }
}
#endif // CF_JIT_SHAPE (Float32)
__device__ __forceinline__ float cf_sop_value_f32(float r2, float dt, const cf_sop_val& P) {
    float a[1] = {r2}, b[1] = {dt}, v[1];
    cf_sop_value_f32_n<1>(a, b, P, v);
    return v[0];
}
