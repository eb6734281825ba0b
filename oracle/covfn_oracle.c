/*
 * covfn_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm (CovarianceFunctions.jl v0.3.5, /root/reference)
 * for the lazy-Gramian multiply path.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load this; the product (libcovfn_b200.so and the Python
 * mirror) never does.
 *
 * PINNING STATUS.  The reference is Julia and Julia is not installed here, so the reference binary
 * cannot be run, and the reference's tests hold no stored golden vectors and no RNG seeds
 * (SURVEY.md section 8c).  The oracle is therefore pinned on the *relations* those tests assert
 * (tests/test_oracle_relations.py reproduces each: test/gramian.jl:56-72, test/stationary.jl:39-42,
 * 53-69, test/mercer.jl:12-19, test/algebra.jl:28-51, test/gradient.jl:47-52,66-70) plus a long-double
 * "truth" evaluator.  Bit-level parity with the Julia binary is UNPINNED (differences: libm exp/pow
 * vs Julia Base, BLAS dot order, @simd re-association; all <= a few ulp per term).  The CG restatement
 * follows IterativeSolvers.jl 0.9.2 (Manifest.toml:353-357), whose source is not under /root/reference:
 * iterate-level parity is unpinned.
 *
 * Every function cites the reference file:line it follows.  The arithmetic is deliberately written
 * the way the reference writes it (per-leaf recomputation of r2 / dot, sequential sums, alpha applied
 * per term) -- it is a checker, not a fast implementation.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t op;
    int32_t iparam;
    double fparam;
} orc_knode_t; /* same layout and op numbering as cf_knode_t (include/covfn_b200.h) */

enum {
    OP_EQ = 1, OP_EXP = 2, OP_RQ = 3, OP_MATERNP = 4, OP_DOT = 5, OP_CONST = 6,
    OP_SUM = 7, OP_PROD = 8, OP_POW = 9, OP_LENGTHSCALE = 10, OP_ARDSCALE = 11, OP_ARD = 12
};

#define ORC_MAXSTACK 64
#define ORC_MAXP 20

/* ------------------------------------------------------------------------------------------
 * Typed scalars: Julia promotes Float32 op Float64 -> Float64, Float32 op Int -> Float32.
 * A num is a double holding either a Float64 or an (exactly representable) Float32.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double v; int f64; } num;

static inline double r32(double x) { return (double)(float)x; }
static inline num mk(double v, int f64) { num r; r.v = f64 ? v : r32(v); r.f64 = f64; return r; }
static inline num n_add(num a, num b) { return mk(a.v + b.v, a.f64 | b.f64); }
static inline num n_sub(num a, num b) { return mk(a.v - b.v, a.f64 | b.f64); }
static inline num n_mul(num a, num b) { return mk(a.v * b.v, a.f64 | b.f64); }
static inline num n_div(num a, num b) { return mk(a.v / b.v, a.f64 | b.f64); }
static inline num n_muli(num a, double i) { return mk(a.v * i, a.f64); }   /* Float * Int */
static inline num n_divi(num a, double i) { return mk(a.v / i, a.f64); }   /* Float / Int */
static inline num n_addi(num a, double i) { return mk(a.v + i, a.f64); }
static inline num n_neg(num a) { num r = a; r.v = -a.v; return r; }
static inline num n_exp(num a) { return a.f64 ? mk(exp(a.v), 1) : mk((double)expf((float)a.v), 0); }
static inline num n_sqrt(num a) { return a.f64 ? mk(sqrt(a.v), 1) : mk((double)sqrtf((float)a.v), 0); }
/* Julia 1.8 ^(x::Float64, y::Integer) (Base math.jl): special-cases -1,0,1,2,3 then llvm.pow */
static inline num n_powi(num a, int p) {
    if (p == -1) return mk(1.0 / a.v, a.f64);
    if (p == 0) return mk(1.0, a.f64);
    if (p == 1) return a;
    if (p == 2) return mk(a.v * a.v, a.f64);
    if (p == 3) { num t = mk(a.v * a.v, a.f64); return mk(t.v * a.v, a.f64); }
    return a.f64 ? mk(pow(a.v, (double)p), 1) : mk((double)powf((float)a.v, (float)p), 0);
}
static inline num n_powf(num a, num b) { /* Float ^ Float */
    int f = a.f64 | b.f64;
    return f ? mk(pow(a.v, b.v), 1) : mk((double)powf((float)a.v, (float)b.v), 0);
}

static double dfactorial(int n) { double f = 1; for (int i = 2; i <= n; i++) f *= i; return f; }
static double dbinomial(int n, int k) { return dfactorial(n) / (dfactorial(k) * dfactorial(n - k)); }

/* reference src/stationary.jl:184-191  MaternP_coefficients(p): reverse of binomial(p,i)*(p+i)!/p! */
static void maternp_coefficients(int p, double* c) {
    for (int i = 1; i <= p; i++) c[p - i] = dbinomial(p, i) * (dfactorial(p + i) / dfactorial(p));
}
/* reference src/stationary.jl:172-182  MaternP_derivatives_at_zero(p): the reference differentiates the
 * naive closed form symbolically (SymEngine) and evaluates at 0.  The i-th derivative of the
 * Matern(nu = p + 1/2) kernel with respect to r2 at 0 has the closed form
 *     d_i = (nu/2)^i / prod_{m=1..i} (m - nu)
 * (series of the modified Bessel function K_nu); tests/test_oracle_relations.py checks it against
 * symbolic differentiation with sympy for p <= 6. */
static void maternp_derivatives_at_zero(int p, double* d) {
    double nu = p + 0.5, num_ = 1, den = 1;
    for (int i = 1; i <= p; i++) { num_ *= nu / 2; den *= (i - nu); d[i - 1] = num_ / den; }
}

/* ------------------------------------------------------------------------------------------
 * scalar pieces
 * ------------------------------------------------------------------------------------------ */
/* reference src/util.jl:40-47  euclidean2: val += (x[i]-y[i])^2, sequential, in the data type */
static num euclidean2(int d, const double* x, const double* y, int f64) {
    num val = mk(0, f64);
    for (int i = 0; i < d; i++) {
        num t = mk(x[i] - y[i], f64);
        val = n_add(val, n_mul(t, t));
    }
    return val;
}
/* LinearAlgebra.dot (BLAS ddot on Vector{Float64}; order unspecified there) -- sequential here */
static num dotxy(int d, const double* x, const double* y, int f64) {
    num val = mk(0, f64);
    for (int i = 0; i < d; i++) val = n_add(val, mk(x[i] * y[i], f64));
    return val;
}

/* isotropic leaves as functions of r2 */
static num leaf_eq(num r2) { return n_exp(n_divi(n_neg(r2), 2)); }              /* stationary.jl:42 */
static num leaf_exp(num r2) { return n_exp(n_neg(n_sqrt(r2))); }                /* stationary.jl:60 */
static num leaf_rq(num r2, double alpha, int alpha_is_int) {                    /* stationary.jl:53 */
    if (alpha_is_int) {
        num base = n_addi(n_divi(r2, 2 * alpha), 1);
        return n_powi(base, -(int)alpha);
    }
    num a = mk(alpha, 1);
    num base = n_addi(n_div(r2, n_muli(a, 2)), 1);
    return n_powf(base, n_neg(a));
}
/* reference src/stationary.jl:134-158 */
static num leaf_maternp(num r2, int p) {
    double coef[ORC_MAXP], der[ORC_MAXP];
    double epsT = r2.f64 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
    double taylor_bound = (p == 0) ? 0.0 : pow(epsT, 1.0 / p); /* eps^(1/0) = eps^Inf = 0 */
    if (r2.v < taylor_bound) {                                                  /* :139-146 */
        maternp_derivatives_at_zero(p, der);
        num y = mk(1, r2.f64);
        num r2i = r2;
        for (int i = 1; i <= p; i++) {
            y = n_add(y, n_divi(n_mul(mk(der[i - 1], 1), r2i), dfactorial(i)));
            r2i = n_mul(r2i, r2);
        }
        return y;
    }
    maternp_coefficients(p, coef);                                              /* :148-157 */
    num y = mk(0, r2.f64);
    num r = n_sqrt(n_muli(r2, 2 * p + 1));
    num ri = mk(1, r.f64);
    for (int i = 1; i <= p; i++) {
        y = n_add(y, n_mul(mk(coef[i - 1], 1), ri));
        ri = n_mul(ri, n_muli(r, 2));
    }
    y = n_add(y, ri);
    double norm = dfactorial(2 * p) / dfactorial(p);
    return n_mul(y, n_divi(n_exp(n_neg(r)), norm));
}

/* ------------------------------------------------------------------------------------------
 * k(x, y): postfix evaluation with the reference's semantics: every leaf recomputes its own
 * euclidean2 / dot (reference src/algebra.jl:17,40,62, src/stationary.jl:9, src/mercer.jl:3).
 * A LENGTHSCALE node rescales the r2 seen by the isotropic leaf below it
 * (reference src/transformation.jl:19), so leaves are evaluated lazily: a leaf pushes a marker
 * and is finalised when the next node is not a LENGTHSCALE.
 * ------------------------------------------------------------------------------------------ */
static num eval_iso_leaf(const orc_knode_t* nd, num r2) {
    switch (nd->op) {
        case OP_EQ: return leaf_eq(r2);
        case OP_EXP: return leaf_exp(r2);
        case OP_RQ: return leaf_rq(r2, nd->fparam, nd->iparam);
        case OP_MATERNP: return leaf_maternp(r2, nd->iparam);
        default: return mk(NAN, 1);
    }
}
static int is_iso_leaf(int op) { return op == OP_EQ || op == OP_EXP || op == OP_RQ || op == OP_MATERNP; }

/* ARD(k, l) = Normed(k, tau -> enorm2(Diagonal(inv.(l)), tau))            reference src/transformation.jl:25-45
 * (m::Normed)(x, y) = m.k(m.n2(difference(x, y)));  enorm2(A, x) = dot(x, A, x)       src/transformation.jl:38-39, src/util.jl:52
 * LinearAlgebra.dot(x, D::Diagonal, y) sums conj(x_c) * d_c * y_c = (tau_c * (1 / l_c)) * tau_c, sequentially for short vectors;
 * inv.(l) is Float64, so the squared norm is Float64 whatever the data type. */
static num ard_norm2(int d, const double* x, const double* y, int f64, const orc_knode_t* scales) {
    num val = mk(0, 1);
    for (int c = 0; c < d; c++) {
        num tau = mk(x[c] - y[c], f64);
        num invl = mk(1.0 / scales[c].fparam, 1);
        val = n_add(val, n_mul(n_mul(tau, invl), tau));
    }
    return val;
}
/* postfix layout: <child nodes> ARDSCALE(l_1) ... ARDSCALE(l_d) ARD(d).  For the leaf at index t, the ARDSCALE block of the ARD
 * node whose child subtree contains t, or NULL. */
static int node_arity(const orc_knode_t* nd) {
    switch (nd->op) {
        case OP_SUM: case OP_PROD: return nd->iparam;
        case OP_POW: case OP_LENGTHSCALE: return 1;
        case OP_ARD: return nd->iparam + 1;
        default: return 0;
    }
}
static const orc_knode_t* ard_scales_for(const orc_knode_t* prog, int nnodes, int t) {
    for (int u = t + 1; u < nnodes; u++) {
        if (prog[u].op != OP_ARD) continue;
        int k = prog[u].iparam, root = u - k - 1, need = 1, i = root;
        while (need > 0 && i >= 0) { need += node_arity(&prog[i]) - 1; i--; }
        if (t > i && t <= root) return &prog[u - k];
    }
    return NULL;
}

static num kernel_eval(const orc_knode_t* prog, int nnodes, int d, const double* x, const double* y, int f64) {
    num st[ORC_MAXSTACK];
    int sp = 0, has_ard = 0;
    for (int t = 0; t < nnodes; t++) has_ard |= (prog[t].op == OP_ARD);
    for (int t = 0; t < nnodes; t++) {
        const orc_knode_t* nd = &prog[t];
        if (is_iso_leaf(nd->op)) {
            const orc_knode_t* sc = has_ard ? ard_scales_for(prog, nnodes, t) : NULL;
            num r2 = sc ? ard_norm2(d, x, y, f64, sc) : euclidean2(d, x, y, f64);
            /* Lengthscale wrappers directly above the leaf: k(r2 / l^2), outermost first */
            int u = t + 1, nls = 0;
            while (u < nnodes && prog[u].op == OP_LENGTHSCALE) { u++; nls++; }
            for (int w = t + nls; w > t; w--) {
                num l = mk(prog[w].fparam, 1);
                r2 = n_div(r2, n_powi(l, 2));
            }
            st[sp++] = eval_iso_leaf(nd, r2);
            t += nls;
        } else if (nd->op == OP_DOT) {
            st[sp++] = dotxy(d, x, y, f64);
        } else if (nd->op == OP_CONST) {
            num cst; /* Constant{Int} (iparam = 1) never promotes; Constant{Float64} does */
            cst.v = nd->fparam; cst.f64 = nd->iparam ? 0 : 1;
            st[sp++] = cst;
        } else if (nd->op == OP_SUM || nd->op == OP_PROD) {
            int k = nd->iparam;
            num acc = st[sp - k];
            for (int q = 1; q < k; q++) acc = (nd->op == OP_SUM) ? n_add(acc, st[sp - k + q]) : n_mul(acc, st[sp - k + q]);
            sp -= k;
            st[sp++] = acc;
        } else if (nd->op == OP_POW) {
            st[sp - 1] = n_powi(st[sp - 1], nd->iparam);
        } else if (nd->op == OP_ARDSCALE || nd->op == OP_ARD) {
            /* the metric was applied at the leaves; the child's value stays on the stack */
        } else {
            return mk(NAN, 1);
        }
    }
    return st[0];
}

int orc_force_interpreter = 0; /* tests set this to compare the fast loops with the interpreter */
void orc_set_force_interpreter(int v) { orc_force_interpreter = v; }

/* data access: points are d x n column-major in the data dtype; promote each point to doubles */
static void load_point(int dtype64, const void* X, int64_t ldx, int64_t i, int d, double* out) {
    if (dtype64) { const double* p = (const double*)X + ldx * i; for (int c = 0; c < d; c++) out[c] = p[c]; }
    else { const float* p = (const float*)X + ldx * i; for (int c = 0; c < d; c++) out[c] = p[c]; }
}
static inline double ldv(int dtype64, const void* a, int64_t i) { return dtype64 ? ((const double*)a)[i] : (double)((const float*)a)[i]; }
static inline void stv(int dtype64, void* a, int64_t i, double v) { if (dtype64) ((double*)a)[i] = v; else ((float*)a)[i] = (float)v; }

/* G[i,j] = k(x[i], y[j])  (reference src/gramian.jl:37-40) */
double orc_getindex(const orc_knode_t* prog, int nnodes, int dtype64, int d, const void* X, int64_t ldx,
                    int64_t i, const void* Y, int64_t ldy, int64_t j) {
    double xi[d > 0 ? d : 1], yj[d > 0 ? d : 1];
    load_point(dtype64, X, ldx, i, d, xi);
    load_point(dtype64, Y, ldy, j, d, yj);
    return kernel_eval(prog, nnodes, d, xi, yj, dtype64).v;
}

/* Matrix!(M, G)  (reference src/gramian.jl:107-114), threads over columns */
void orc_matrix(const orc_knode_t* prog, int nnodes, int dtype64, int d, int64_t n, const void* X, int64_t ldx,
                int64_t m, const void* Y, int64_t ldy, void* M, int64_t ldm) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m; j++) {
        double xi[d > 0 ? d : 1], yj[d > 0 ? d : 1];
        load_point(dtype64, Y, ldy, j, d, yj);
        for (int64_t i = 0; i < n; i++) {
            load_point(dtype64, X, ldx, i, d, xi);
            stv(dtype64, M, i + ldm * j, kernel_eval(prog, nnodes, d, xi, yj, dtype64).v);
        }
    }
}

/*
 * Specialised Float64 row loops for single-leaf programs.  Julia compiles a specialised mul! for every
 * kernel type; the generic postfix interpreter above would make the timed CPU baseline unfairly slow.
 * The arithmetic is IDENTICAL to the interpreter (same operations in the same order, coefficients
 * hoisted like the reference's MaternP struct fields) -- tests/test_oracle_relations.py checks the two
 * agree bit for bit.  Returns 0 if the program is not a single Float64 leaf.
 */
static int mul_vec_fast(const orc_knode_t* prog, int nnodes, int d, int64_t m, const double* Yd, int64_t i0,
                        int64_t i1, const double* X, int64_t ldx, double* y, const double* x, double alpha, double beta) {
    if (nnodes != 1 || !is_iso_leaf(prog[0].op)) return 0;
    const int op = prog[0].op, p = prog[0].iparam;
    const double al = prog[0].fparam;
    double coef[ORC_MAXP], der[ORC_MAXP];
    double taylor_bound = 0, norm = 1;
    if (op == OP_MATERNP) {
        maternp_coefficients(p, coef);
        maternp_derivatives_at_zero(p, der);
        taylor_bound = (p == 0) ? 0.0 : pow(2.220446049250313e-16, 1.0 / p);
        norm = dfactorial(2 * p) / dfactorial(p);
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        const double* xi = X + ldx * i;
        double yi = beta == 0 ? 0.0 : beta * y[i - i0];
        for (int64_t j = 0; j < m; j++) {
            const double* yj = Yd + j * d;
            double r2 = 0;
            for (int c = 0; c < d; c++) { double t = xi[c] - yj[c]; r2 += t * t; }
            double g;
            if (op == OP_EQ) g = exp(-r2 / 2);
            else if (op == OP_EXP) g = exp(-sqrt(r2));
            else if (op == OP_RQ) {
                if (p) { double base = r2 / (2 * al) + 1; int q = -(int)al;
                         g = (q == -1) ? 1.0 / base : pow(base, (double)q); }
                else g = pow(r2 / (al * 2) + 1, -al);
            } else {
                if (r2 < taylor_bound) {
                    double yv = 1, r2i = r2;
                    for (int q = 1; q <= p; q++) { yv += der[q - 1] * r2i / dfactorial(q); r2i *= r2; }
                    g = yv;
                } else {
                    double yv = 0, r = sqrt(r2 * (2 * p + 1)), ri = 1;
                    for (int q = 1; q <= p; q++) { yv += coef[q - 1] * ri; ri *= r * 2; }
                    yv += ri;
                    g = yv * (exp(-r) / norm);
                }
            }
            yi += alpha * g * x[j];
        }
        y[i - i0] = yi;
    }
    return 1;
}

/*
 * mul!(y::AbstractVector, G::Gramian, x::AbstractVector, alpha, beta)   reference src/gramian.jl:78-87
 *     @. y = iszero(beta) ? 0 : beta * y
 *     @threads for i in 1:n;  @simd for j in 1:m;  y[i] += alpha * G[i, j] * x[j]
 * rows [i0, i1) only (i0=0, i1=n for the full product) so that big configurations can be checked on
 * a row subset; y addresses row i0.
 */
void orc_gramian_mul_vec(const orc_knode_t* prog, int nnodes, int dtype64, int d, int64_t n, const void* X,
                         int64_t ldx, int64_t m, const void* Y, int64_t ldy, int64_t i0, int64_t i1,
                         void* y, const void* x, double alpha, double beta) {
    (void)n;
    double* Yd = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * (size_t)(d > 0 ? d : 1));
    for (int64_t j = 0; j < m; j++) load_point(dtype64, Y, ldy, j, d, Yd + j * d);
    if (dtype64 && !orc_force_interpreter &&
        mul_vec_fast(prog, nnodes, d, m, Yd, i0, i1, (const double*)X, ldx, (double*)y, (const double*)x, alpha, beta)) {
        free(Yd);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        double xi[d > 0 ? d : 1];
        load_point(dtype64, X, ldx, i, d, xi);
        num yi = mk(beta == 0 ? 0.0 : beta * ldv(dtype64, y, i - i0), dtype64);
        num al = mk(alpha, dtype64);
        for (int64_t j = 0; j < m; j++) {
            num g = kernel_eval(prog, nnodes, d, xi, Yd + j * d, dtype64);
            num term = n_mul(n_mul(al, g), mk(ldv(dtype64, x, j), dtype64));
            yi = mk(yi.v + term.v, dtype64); /* stored back into y::Vector{T} every iteration */
        }
        stv(dtype64, y, i - i0, yi.v);
    }
    free(Yd);
}

/*
 * mul!(Y::AbstractMatrix, G::Gramian, X::AbstractMatrix, alpha, beta)   reference src/gramian.jl:89-99
 *     threads over RHS columns j; for i; @simd for k: Y[i,j] += alpha * G[i,k] * X[k,j]
 * (the kernel entry is recomputed for every column, exactly as the reference does).
 */
void orc_gramian_mul_mat(const orc_knode_t* prog, int nnodes, int dtype64, int d, int64_t n, const void* X,
                         int64_t ldx, int64_t m, const void* Y, int64_t ldy, int64_t i0, int64_t i1,
                         void* B, int64_t ldb, const void* A, int64_t lda, int64_t nrhs, double alpha, double beta) {
    (void)n;
    double* Yd = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * (size_t)(d > 0 ? d : 1));
    for (int64_t j = 0; j < m; j++) load_point(dtype64, Y, ldy, j, d, Yd + j * d);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < nrhs; j++) {
        double xi[d > 0 ? d : 1];
        for (int64_t i = i0; i < i1; i++) {
            load_point(dtype64, X, ldx, i, d, xi);
            num yi = mk(beta == 0 ? 0.0 : beta * ldv(dtype64, B, (i - i0) + ldb * j), dtype64);
            num al = mk(alpha, dtype64);
            for (int64_t k = 0; k < m; k++) {
                num g = kernel_eval(prog, nnodes, d, xi, Yd + k * d, dtype64);
                num term = n_mul(n_mul(al, g), mk(ldv(dtype64, A, k + lda * j), dtype64));
                yi = mk(yi.v + term.v, dtype64);
            }
            stv(dtype64, B, (i - i0) + ldb * j, yi.v);
        }
    }
    free(Yd);
}

/* Same product but evaluating each entry once per row (loop order i, k, j): used only as the
 * *timed CPU baseline* variant "port-fused" in bench.py, never as the checker. */
void orc_gramian_mul_mat_fused(const orc_knode_t* prog, int nnodes, int dtype64, int d, int64_t n, const void* X,
                               int64_t ldx, int64_t m, const void* Y, int64_t ldy, int64_t i0, int64_t i1,
                               void* B, int64_t ldb, const void* A, int64_t lda, int64_t nrhs, double alpha, double beta) {
    (void)n;
    double* Yd = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1) * (size_t)(d > 0 ? d : 1));
    for (int64_t j = 0; j < m; j++) load_point(dtype64, Y, ldy, j, d, Yd + j * d);
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        double xi[d > 0 ? d : 1];
        double acc[nrhs > 0 ? nrhs : 1];
        load_point(dtype64, X, ldx, i, d, xi);
        for (int64_t j = 0; j < nrhs; j++) acc[j] = beta == 0 ? 0.0 : beta * ldv(dtype64, B, (i - i0) + ldb * j);
        for (int64_t k = 0; k < m; k++) {
            double g = alpha * kernel_eval(prog, nnodes, d, xi, Yd + k * d, dtype64).v;
            for (int64_t j = 0; j < nrhs; j++) acc[j] += g * ldv(dtype64, A, k + lda * j);
        }
        for (int64_t j = 0; j < nrhs; j++) stv(dtype64, B, (i - i0) + ldb * j, acc[j]);
    }
    free(Yd);
}

/* ------------------------------------------------------------------------------------------
 * Second-order jets in r2: value, d/dr2, d2/dr2^2 -- the restatement of the nested
 * ForwardDiff.derivative calls in derivative_laplacian / value_derivative
 * (reference src/gradient.jl:589-600).  ForwardDiff differentiates the code path actually taken
 * (incl. MaternP's Taylor branch, reference src/stationary.jl:136-146); so does this.
 * Float64 only.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double v, d1, d2; } jet;
static inline jet j_const(double c) { jet r = {c, 0, 0}; return r; }
static inline jet j_var(double x) { jet r = {x, 1, 0}; return r; }
static inline jet j_add(jet a, jet b) { jet r = {a.v + b.v, a.d1 + b.d1, a.d2 + b.d2}; return r; }
static inline jet j_scale(jet a, double c) { jet r = {a.v * c, a.d1 * c, a.d2 * c}; return r; }
static inline jet j_mul(jet a, jet b) {
    jet r = {a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + 2 * a.d1 * b.d1 + a.v * b.d2};
    return r;
}
/* f(a) with f', f'' given */
static inline jet j_chain(jet a, double f, double f1, double f2) {
    jet r = {f, f1 * a.d1, f2 * a.d1 * a.d1 + f1 * a.d2};
    return r;
}
static inline jet j_exp(jet a) { double e = exp(a.v); return j_chain(a, e, e, e); }
static inline jet j_sqrt(jet a) { double s = sqrt(a.v); return j_chain(a, s, 0.5 / s, -0.25 / (s * a.v)); }
static inline jet j_powr(jet a, double p) { /* a^p, real p */
    double f = pow(a.v, p);
    return j_chain(a, f, p * pow(a.v, p - 1), p * (p - 1) * pow(a.v, p - 2));
}
static inline jet j_powi(jet a, int p) {
    if (p == 0) return j_const(1);
    if (p == 1) return a;
    if (p == 2) return j_mul(a, a);
    if (p == 3) return j_mul(j_mul(a, a), a);
    return j_powr(a, (double)p);
}

static jet jet_iso_leaf(const orc_knode_t* nd, jet r2) {
    switch (nd->op) {
        case OP_EQ: return j_exp(j_scale(r2, -0.5));
        case OP_EXP: return j_exp(j_scale(j_sqrt(r2), -1));
        case OP_RQ: {
            jet base = j_add(j_scale(r2, 1.0 / (2 * nd->fparam)), j_const(1));
            return j_powr(base, -nd->fparam);
        }
        case OP_MATERNP: {
            int p = nd->iparam;
            double coef[ORC_MAXP], der[ORC_MAXP];
            double taylor_bound = (p == 0) ? 0.0 : pow(2.220446049250313e-16, 1.0 / p);
            if (r2.v < taylor_bound) {
                maternp_derivatives_at_zero(p, der);
                jet y = j_const(1), r2i = r2;
                for (int i = 1; i <= p; i++) {
                    y = j_add(y, j_scale(r2i, der[i - 1] / dfactorial(i)));
                    r2i = j_mul(r2i, r2);
                }
                return y;
            }
            maternp_coefficients(p, coef);
            jet y = j_const(0), r = j_sqrt(j_scale(r2, 2 * p + 1)), ri = j_const(1);
            for (int i = 1; i <= p; i++) {
                y = j_add(y, j_scale(ri, coef[i - 1]));
                ri = j_mul(ri, j_scale(r, 2));
            }
            y = j_add(y, ri);
            double norm = dfactorial(2 * p) / dfactorial(p);
            return j_mul(y, j_scale(j_exp(j_scale(r, -1)), 1.0 / norm));
        }
        default: return j_const(NAN);
    }
}

/* k(r2) for a program with the IsotropicInput trait (reference src/properties.jl:39-63):
 * (S::Sum)(tau), (P::Product)(tau), (P::Power)(tau) at src/algebra.jl:16,39,61 */
static jet kernel_jet(const orc_knode_t* prog, int nnodes, double r2v) {
    jet st[ORC_MAXSTACK];
    int sp = 0;
    for (int t = 0; t < nnodes; t++) {
        const orc_knode_t* nd = &prog[t];
        if (is_iso_leaf(nd->op)) {
            jet r2 = j_var(r2v);
            int u = t + 1, nls = 0;
            while (u < nnodes && prog[u].op == OP_LENGTHSCALE) { u++; nls++; }
            for (int w = t + nls; w > t; w--) r2 = j_scale(r2, 1.0 / (prog[w].fparam * prog[w].fparam));
            st[sp++] = jet_iso_leaf(nd, r2);
            t += nls;
        } else if (nd->op == OP_DOT) {
            st[sp++] = j_var(r2v); /* DotProductInput programs: the variable is x.y, (k::Dot)(d) = d (mercer.jl:9) */
        } else if (nd->op == OP_CONST) {
            st[sp++] = j_const(nd->fparam);
        } else if (nd->op == OP_SUM || nd->op == OP_PROD) {
            int k = nd->iparam;
            jet acc = st[sp - k];
            for (int q = 1; q < k; q++) acc = (nd->op == OP_SUM) ? j_add(acc, st[sp - k + q]) : j_mul(acc, st[sp - k + q]);
            sp -= k;
            st[sp++] = acc;
        } else if (nd->op == OP_POW) {
            st[sp - 1] = j_powi(st[sp - 1], nd->iparam);
        } else {
            return j_const(NAN);
        }
    }
    return st[0];
}

/* value, k', k'' at r2 (reference src/gradient.jl:584-592) */
void orc_value_derivative_laplacian(const orc_knode_t* prog, int nnodes, double r2, double* out3) {
    jet j = kernel_jet(prog, nnodes, r2);
    out3[0] = j.v; out3[1] = j.d1; out3[2] = j.d2;
}

/*
 * blockmul!(y, G::Gramian, x, alpha, beta) with the isotropic gradient element
 *   reference src/gramian.jl:241-253 and src/gradient.jl:86-92:
 *     y[i] = beta == 0 ? 0 : beta*y[i]
 *     for j:  r = x_i - y_j;  r2 = sum(abs2, r);  (k1, k2) = derivative_laplacian(k, r2)
 *             dot_r_a = r'a;  b .= alpha * -2(k1*a + 2*k2*r*dot_r_a) + 1*b
 * flat vectors: entry i*d + c (BlockFactorization, isstrided = true; reference src/gramian.jl:120-123).
 * Float64 only.  rows [i0, i1) of points.
 */
void orc_gradient_mul(const orc_knode_t* prog, int nnodes, int d, int64_t n, const double* X, int64_t ldx,
                      int64_t m, const double* Y, int64_t ldy, int64_t i0, int64_t i1, double* y,
                      const double* x, double alpha, double beta) {
    (void)n;
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        const double* xi = X + ldx * i;
        double* b = y + (i - i0) * d;
        double r[d > 0 ? d : 1];
        for (int c = 0; c < d; c++) b[c] = (beta == 0) ? 0.0 : beta * b[c];
        for (int64_t j = 0; j < m; j++) {
            const double* yj = Y + ldy * j;
            const double* a = x + j * d;
            double r2 = 0, dot_r_a = 0;
            for (int c = 0; c < d; c++) { r[c] = xi[c] - yj[c]; r2 += r[c] * r[c]; }
            jet kj = kernel_jet(prog, nnodes, r2);
            double k1 = kj.d1, k2 = kj.d2;
            for (int c = 0; c < d; c++) dot_r_a += r[c] * a[c];
            for (int c = 0; c < d; c++) b[c] = alpha * (-2 * (k1 * a[c] + 2 * k2 * r[c] * dot_r_a)) + 1.0 * b[c];
        }
    }
}

/*
 * Generalised block multiply for the derivative kernels on the path:
 *   trait 0 = IsotropicInput  (variable r2 = |x - y|^2):  gradient element  src/gradient.jl:86-92
 *   trait 1 = DotProductInput (variable t = x . y):       gradient element  src/gradient.jl:109-115
 *                                                          b = alpha (k1 a + k2 y (x . a)) + beta b
 *   vg = 0: GradientKernel, blocks d x d;  vg = 1: ValueGradientKernel, blocks (d+1) x (d+1), entry 0 = value
 *           (DerivativeKernelElement [vv vg; gv gg], src/gradient.jl:217-239; value_gradient_kernel! :442-463:
 *            isotropic: vv = k0, vg = -2 k1 r, gv = 2 k1 r;  dot product: vv = k0, vg = k1 x, gv = k1 y)
 * blockmul! loop structure as in src/gramian.jl:241-253.  Float64.
 */
void orc_derivative_mul(const orc_knode_t* prog, int nnodes, int d, int trait, int vg, int64_t n, const double* X, int64_t ldx,
                        int64_t m, const double* Y, int64_t ldy, int64_t i0, int64_t i1, double* y, const double* x,
                        double alpha, double beta) {
    (void)n;
    const int bs = d + (vg ? 1 : 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        const double* xi = X + ldx * i;
        double* b = y + (i - i0) * bs;
        double r[d > 0 ? d : 1], t[bs > 0 ? bs : 1];
        for (int c = 0; c < bs; c++) b[c] = (beta == 0) ? 0.0 : beta * b[c];
        for (int64_t j = 0; j < m; j++) {
            const double* yj = Y + ldy * j;
            const double* a = x + j * bs;
            const double* ag = a + (vg ? 1 : 0);
            double var = 0;
            if (trait == 0) { for (int c = 0; c < d; c++) { r[c] = xi[c] - yj[c]; var += r[c] * r[c]; } }
            else { for (int c = 0; c < d; c++) var += xi[c] * yj[c]; }
            jet kj = kernel_jet(prog, nnodes, var);
            const double k0 = kj.v, k1 = kj.d1, k2 = kj.d2;
            /* t = block * a */
            if (trait == 0) {
                double dra = 0;
                for (int c = 0; c < d; c++) dra += r[c] * ag[c];
                for (int c = 0; c < d; c++) t[c + (vg ? 1 : 0)] = -2 * (k1 * ag[c] + 2 * k2 * r[c] * dra);
                if (vg) {
                    double vgdot = 0;
                    for (int c = 0; c < d; c++) vgdot += (-2 * k1 * r[c]) * ag[c];
                    t[0] = k0 * a[0] + vgdot;
                    for (int c = 0; c < d; c++) t[c + 1] += (2 * k1 * r[c]) * a[0];
                }
            } else {
                double dxa = 0;
                for (int c = 0; c < d; c++) dxa += xi[c] * ag[c];
                for (int c = 0; c < d; c++) t[c + (vg ? 1 : 0)] = k1 * ag[c] + k2 * yj[c] * dxa;
                if (vg) {
                    double vgdot = 0;
                    for (int c = 0; c < d; c++) vgdot += (k1 * xi[c]) * ag[c];
                    t[0] = k0 * a[0] + vgdot;
                    for (int c = 0; c < d; c++) t[c + 1] += (k1 * yj[c]) * a[0];
                }
            }
            for (int c = 0; c < bs; c++) b[c] = alpha * t[c] + 1.0 * b[c];
        }
    }
}

/* dense instantiation of the same operators, column-major, ldm >= n*bs */
void orc_derivative_matrix(const orc_knode_t* prog, int nnodes, int d, int trait, int vg, int64_t n, const double* X, int64_t ldx,
                           int64_t m, const double* Y, int64_t ldy, double* M, int64_t ldm) {
    const int bs = d + (vg ? 1 : 0), o = vg ? 1 : 0;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            const double* xi = X + ldx * i;
            const double* yj = Y + ldy * j;
            double r[d > 0 ? d : 1], var = 0;
            if (trait == 0) { for (int c = 0; c < d; c++) { r[c] = xi[c] - yj[c]; var += r[c] * r[c]; } }
            else { for (int c = 0; c < d; c++) var += xi[c] * yj[c]; }
            jet kj = kernel_jet(prog, nnodes, var);
            double* B = M + (i * bs) + ldm * (j * bs);
            for (int c = 0; c < d; c++)
                for (int e = 0; e < d; e++)
                    B[(c + o) + ldm * (e + o)] = (trait == 0) ? -2 * (kj.d1 * (c == e) + 2 * kj.d2 * r[c] * r[e])
                                                              : kj.d1 * (c == e) + kj.d2 * yj[c] * xi[e];
            if (vg) {
                B[0] = kj.v;
                for (int c = 0; c < d; c++) {
                    B[0 + ldm * (c + 1)] = (trait == 0) ? -2 * kj.d1 * r[c] : kj.d1 * xi[c]; /* value_gradient (row) */
                    B[(c + 1) + 0] = (trait == 0) ? 2 * kj.d1 * r[c] : kj.d1 * yj[c];         /* gradient_value (column) */
                }
            }
        }
}

/* dense (n d) x (m d) instantiation of the gradient Gramian: block (i,j) = -2 (k1 I + 2 k2 r r^T)
 * (Matrix(GradientKernelElement) = K * I, reference src/gradient.jl:61).  Column-major, ldm >= n d. */
void orc_gradient_matrix(const orc_knode_t* prog, int nnodes, int d, int64_t n, const double* X, int64_t ldx,
                         int64_t m, const double* Y, int64_t ldy, double* M, int64_t ldm) {
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < m; j++) {
            double r[d > 0 ? d : 1], r2 = 0;
            for (int c = 0; c < d; c++) { r[c] = X[ldx * i + c] - Y[ldy * j + c]; r2 += r[c] * r[c]; }
            jet kj = kernel_jet(prog, nnodes, r2);
            for (int c = 0; c < d; c++)
                for (int e = 0; e < d; e++)
                    M[(i * d + c) + ldm * (j * d + e)] = -2 * (kj.d1 * (c == e ? 1.0 : 0.0) + 2 * kj.d2 * r[c] * r[e]);
        }
}

/*
 * x = (sigma2*I + K) \ b by conjugate gradients:
 *   LazyMatrixSum(Diagonal, Gramian) mul!  reference src/lazy_linear_algebra.jl:126-133
 *     (y = 0; y += 1*(D*x); y += 1*(G*x)), ldiv! -> cg!  src/lazy_linear_algebra.jl:142-144
 *   cg! restated from IterativeSolvers.jl 0.9.2 src/cg.jl [upstream, source absent, UNPINNED]:
 *     u = 0; r = b - A x; residual = ||r||; prev_residual = 1; tol = max(reltol*residual0, abstol)
 *     loop while residual > tol and it < maxiter:
 *        beta = residual^2 / prev_residual^2;  u = r + beta u;  c = A u
 *        alpha = residual^2 / dot(u, c);  x += alpha u;  r -= alpha c
 *        prev_residual = residual; residual = ||r||
 * gradient != 0 uses the gradient operator (size n d).  history (if non-NULL) receives the residual
 * norm after every iteration (length maxiter).  Returns the iteration count.
 */
int orc_cg_solve(const orc_knode_t* prog, int nnodes, int d, int64_t n, const double* X, int64_t ldx,
                 double sigma2, double* x, const double* b, double reltol, int maxiter, int gradient,
                 double* resnorm, double* history) {
    int64_t N = gradient ? n * d : n;
    double *u = calloc(N, sizeof(double)), *r = malloc(N * sizeof(double)), *c = malloc(N * sizeof(double));
    if (reltol <= 0) reltol = sqrt(2.220446049250313e-16);
    if (maxiter <= 0) maxiter = (int)N;
#define APPLY(out, in)                                                                                   \
    do {                                                                                                 \
        for (int64_t q = 0; q < N; q++) (out)[q] = 0.0;                                                  \
        for (int64_t q = 0; q < N; q++) (out)[q] += 1.0 * (sigma2 * (in)[q]);                            \
        if (gradient) orc_gradient_mul(prog, nnodes, d, n, X, ldx, n, X, ldx, 0, n, (out), (in), 1.0, 1.0); \
        else orc_gramian_mul_vec(prog, nnodes, 1, d, n, X, ldx, n, X, ldx, 0, n, (out), (in), 1.0, 1.0);   \
    } while (0)
    APPLY(c, x);
    for (int64_t q = 0; q < N; q++) r[q] = b[q] - c[q];
    double residual = 0, prev_residual = 1;
    for (int64_t q = 0; q < N; q++) residual += r[q] * r[q];
    residual = sqrt(residual);
    double tol = reltol * residual;
    int it = 0;
    while (residual > tol && it < maxiter) {
        double beta = (residual * residual) / (prev_residual * prev_residual);
        for (int64_t q = 0; q < N; q++) u[q] = r[q] + beta * u[q];
        APPLY(c, u);
        double uc = 0;
        for (int64_t q = 0; q < N; q++) uc += u[q] * c[q];
        double alpha = (residual * residual) / uc;
        for (int64_t q = 0; q < N; q++) x[q] += alpha * u[q];
        for (int64_t q = 0; q < N; q++) r[q] -= alpha * c[q];
        prev_residual = residual;
        residual = 0;
        for (int64_t q = 0; q < N; q++) residual += r[q] * r[q];
        residual = sqrt(residual);
        if (history) history[it] = residual;
        it++;
    }
#undef APPLY
    if (resnorm) *resnorm = residual;
    free(u); free(r); free(c);
    return it;
}

/* ------------------------------------------------------------------------------------------
 * long-double "truth" for the vector product (Float64 data, exact leaves in extended precision,
 * Kahan-free but 64-bit mantissa accumulation): used to show oracle and device are both within
 * tolerance of the exact result and differ by rounding / summation order only.
 * ------------------------------------------------------------------------------------------ */
static long double truth_iso(const orc_knode_t* nd, long double r2) {
    switch (nd->op) {
        case OP_EQ: return expl(-r2 / 2);
        case OP_EXP: return expl(-sqrtl(r2));
        case OP_RQ: return powl(1 + r2 / (2 * (long double)nd->fparam), -(long double)nd->fparam);
        case OP_MATERNP: {
            int p = nd->iparam;
            long double r = sqrtl((2 * p + 1) * r2), s = 0;
            for (int i = 0; i <= p; i++)
                s += (long double)(dfactorial(p + i) / (dfactorial(p - i) * dfactorial(i))) * powl(2 * r, p - i);
            return s * expl(-r) / (long double)(dfactorial(2 * p) / dfactorial(p));
        }
        default: return NAN;
    }
}
static long double truth_eval(const orc_knode_t* prog, int nnodes, int d, const double* x, const double* y) {
    long double st[ORC_MAXSTACK];
    int sp = 0;
    long double r2 = 0, dt = 0;
    for (int c = 0; c < d; c++) { long double t = (long double)x[c] - y[c]; r2 += t * t; dt += (long double)x[c] * y[c]; }
    for (int t = 0; t < nnodes; t++) {
        const orc_knode_t* nd = &prog[t];
        if (is_iso_leaf(nd->op)) {
            long double s = r2;
            const orc_knode_t* sc = ard_scales_for(prog, nnodes, t);
            if (sc) { s = 0; for (int c = 0; c < d; c++) { long double tt = (long double)x[c] - y[c]; s += tt * tt / (long double)sc[c].fparam; } }
            int u = t + 1, nls = 0;
            while (u < nnodes && prog[u].op == OP_LENGTHSCALE) { u++; nls++; }
            for (int w = t + nls; w > t; w--) s /= (long double)prog[w].fparam * prog[w].fparam;
            st[sp++] = truth_iso(nd, s);
            t += nls;
        } else if (nd->op == OP_DOT) st[sp++] = dt;
        else if (nd->op == OP_CONST) st[sp++] = nd->fparam;
        else if (nd->op == OP_SUM || nd->op == OP_PROD) {
            int k = nd->iparam;
            long double acc = st[sp - k];
            for (int q = 1; q < k; q++) acc = (nd->op == OP_SUM) ? acc + st[sp - k + q] : acc * st[sp - k + q];
            sp -= k; st[sp++] = acc;
        } else if (nd->op == OP_POW) st[sp - 1] = powl(st[sp - 1], nd->iparam);
        else if (nd->op == OP_ARDSCALE || nd->op == OP_ARD) { /* applied at the leaves */ }
        else return NAN;
    }
    return st[0];
}
void orc_truth_mul_vec(const orc_knode_t* prog, int nnodes, int d, int64_t n, const double* X, int64_t ldx,
                       int64_t m, const double* Y, int64_t ldy, int64_t i0, int64_t i1, double* y,
                       const double* x, double alpha, double beta) {
    (void)n;
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; i++) {
        long double acc = 0;
        for (int64_t j = 0; j < m; j++) acc += truth_eval(prog, nnodes, d, X + ldx * i, Y + ldy * j) * x[j];
        y[i - i0] = (double)((long double)alpha * acc + (beta == 0 ? 0.0L : (long double)beta * y[i - i0]));
    }
}
double orc_truth_getindex(const orc_knode_t* prog, int nnodes, int d, const double* x, const double* y) {
    return (double)truth_eval(prog, nnodes, d, x, y);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}
