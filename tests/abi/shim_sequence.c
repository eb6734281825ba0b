/*
 * shim_sequence.c -- replays, call for call, the C-ABI sequences that julia/CovarianceFunctionsB200.jl issues through `ccall`.
 *
 * Julia is not available in the build image or on the GPU box, so the shim itself cannot be executed; this harness is the
 * language-neutral stand-in: plain C, built with gcc against include/covfn_b200.h, the library bound at run time with
 * dlopen/dlsym exactly like `ccall((:sym, libcovfn), ...)`.  Each scenario S<n> carries the tag the shim's comments use.
 *
 *   shim_sequence <libcovfn_b200.so> --symbols            resolve every symbol the shim calls (no GPU needed)
 *   shim_sequence <libcovfn_b200.so> <in.bin> <out.bin>   run the scenarios on the inputs, write the named result arrays
 *
 * in.bin  : int64 n, d, p, then doubles: X[n*d] (point-major == d x n column-major), Y2[n*d] (a second point set), a[n], A[n*p],
 *           ag[n*d], avg[n*(d+1)], rhs[n], rhsg[n*d], l[d] (ARD length scales)
 * out.bin : records { char name[32]; int64 count; double data[count]; }
 * tests/test_gpu_abi_shim.py writes the inputs, runs this program and compares every record with the oracle.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "covfn_b200.h"

/* the symbols the Julia shim binds */
static int (*p_create)(cf_gramian_t*, const cf_knode_t*, int, int, int, int64_t, const void*, int64_t, int64_t, const void*, int64_t);
static int (*p_destroy)(cf_gramian_t);
static int (*p_mul)(cf_gramian_t, void*, int64_t, const void*, int64_t, int64_t, double, double);
static int (*p_gradient_mul)(cf_gramian_t, void*, int64_t, const void*, int64_t, int64_t, double, double);
static int (*p_value_gradient_mul)(cf_gramian_t, void*, int64_t, const void*, int64_t, int64_t, double, double);
static int (*p_cg_solve)(cf_gramian_t, double, void*, const void*, double, int, int, int*, double*);
static const char* (*p_last_error)(void);
static int (*p_init)(int, const int*);
static int (*p_device_count)(void);

static void* must_sym(void* lib, const char* name) {
    void* s = dlsym(lib, name);
    if (!s) { fprintf(stderr, "missing symbol %s\n", name); exit(3); }
    return s;
}

static FILE* g_out;
static void record(const char* name, const double* v, int64_t count) {
    char nm[32];
    memset(nm, 0, sizeof(nm));
    strncpy(nm, name, 31);
    fwrite(nm, 1, 32, g_out);
    fwrite(&count, 8, 1, g_out);
    fwrite(v, 8, (size_t)count, g_out);
}
#define CHECK(call)                                                                                           \
    do {                                                                                                      \
        int rc__ = (call);                                                                                    \
        if (rc__ != 0) { fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc__, p_last_error()); exit(4); } \
    } while (0)

static double* rd(FILE* f, int64_t count) {
    double* v = (double*)malloc(sizeof(double) * (size_t)(count > 0 ? count : 1));
    if (fread(v, 8, (size_t)count, f) != (size_t)count) { fprintf(stderr, "short input\n"); exit(5); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s lib.so (--symbols | in.bin out.bin)\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
    *(void**)&p_create = must_sym(lib, "cf_gramian_create");
    *(void**)&p_destroy = must_sym(lib, "cf_gramian_destroy");
    *(void**)&p_mul = must_sym(lib, "cf_gramian_mul");
    *(void**)&p_gradient_mul = must_sym(lib, "cf_gradient_mul");
    *(void**)&p_value_gradient_mul = must_sym(lib, "cf_value_gradient_mul");
    *(void**)&p_cg_solve = must_sym(lib, "cf_cg_solve");
    *(void**)&p_last_error = must_sym(lib, "cf_last_error");
    *(void**)&p_init = must_sym(lib, "cf_init");
    *(void**)&p_device_count = must_sym(lib, "cf_device_count");
    if (strcmp(argv[2], "--symbols") == 0) { printf("symbols ok\n"); return 0; }
    if (argc < 4) return 2;

    FILE* fin = fopen(argv[2], "rb");
    g_out = fopen(argv[3], "wb");
    if (!fin || !g_out) { fprintf(stderr, "cannot open files\n"); return 5; }
    int64_t hdr[3];
    if (fread(hdr, 8, 3, fin) != 3) return 5;
    const int64_t n = hdr[0], d = hdr[1], p = hdr[2];
    double *X = rd(fin, n * d), *Y2 = rd(fin, n * d), *a = rd(fin, n), *A = rd(fin, n * p), *ag = rd(fin, n * d),
           *avg = rd(fin, n * (d + 1)), *rhs = rd(fin, n), *rhsg = rd(fin, n * d), *l = rd(fin, d);
    fclose(fin);
    double* y = (double*)malloc(sizeof(double) * (size_t)(n * (d + 1) * (p > 1 ? p : 1) + 16));

    /* the programs program!(p, k) emits (postfix cf_knode_t) */
    const cf_knode_t prog_eq[] = {{CF_OP_EQ, 0, 0.0}};
    const cf_knode_t prog_m2[] = {{CF_OP_MATERNP, 2, 0.0}};
    const cf_knode_t prog_halfrq[] = {{CF_OP_CONST, 0, 0.5}, {CF_OP_RQ, 1, 2.0}, {CF_OP_PROD, 2, 0.0}}; /* 0.5 * RQ(2) */

    /* [S1] two Gramians on ONE point vector x: gramian(EQ(), x) then gramian(MaternP(2), x).  The shim's handle key contains the
     * program, so each gets its own cf_gramian_create; the results must be the two different kernels' products.  Then the same
     * x with another y (the key contains objectid(y)). */
    cf_gramian_t h1 = NULL, h2 = NULL, h3 = NULL;
    CHECK(p_create(&h1, prog_eq, 1, CF_F64, (int)d, n, X, d, n, NULL, d));
    CHECK(p_mul(h1, y, n, a, n, 1, 1.0, 0.0));
    record("S1_eq", y, n);
    CHECK(p_create(&h2, prog_m2, 1, CF_F64, (int)d, n, X, d, n, NULL, d));
    CHECK(p_mul(h2, y, n, a, n, 1, 1.0, 0.0));
    record("S1_matern2", y, n);
    CHECK(p_mul(h1, y, n, a, n, 1, 1.0, 0.0)); /* the first handle is still the EQ Gramian */
    record("S1_eq_again", y, n);
    const int64_t m2 = n / 2;
    CHECK(p_create(&h3, prog_eq, 1, CF_F64, (int)d, n, X, d, m2, Y2, d));
    CHECK(p_mul(h3, y, n, a, m2, 1, 1.0, 0.0));
    record("S1_eq_xy", y, n);

    /* [S2] Float32 points under a Float64 Gramian: gramian(0.5 * RQ(2), x32) has T = Float64 (Constant{Float64} promotes,
     * SURVEY.md Appendix A).  pack(T, x) converts the points to T while packing; dtype and buffer agree. */
    {
        float* X32 = (float*)malloc(sizeof(float) * (size_t)(n * d));
        double* Xc = (double*)malloc(sizeof(double) * (size_t)(n * d));
        for (int64_t q = 0; q < n * d; q++) { X32[q] = (float)X[q]; Xc[q] = (double)X32[q]; }
        cf_gramian_t h = NULL;
        CHECK(p_create(&h, prog_halfrq, 3, CF_F64, (int)d, n, Xc, d, n, NULL, d));
        CHECK(p_mul(h, y, n, a, n, 1, 1.0, 0.0));
        record("S2_f32pts_f64gram", y, n);
        CHECK(p_destroy(h));
        /* and the genuine Float32 Gramian (EQ on Float32 points: T = Float32, Float32 vectors) */
        float* a32 = (float*)malloc(sizeof(float) * (size_t)n);
        float* y32 = (float*)malloc(sizeof(float) * (size_t)n);
        for (int64_t q = 0; q < n; q++) a32[q] = (float)a[q];
        h = NULL;
        CHECK(p_create(&h, prog_eq, 1, CF_F32, (int)d, n, X32, d, n, NULL, d));
        CHECK(p_mul(h, y32, n, a32, n, 1, 1.0, 0.0));
        for (int64_t q = 0; q < n; q++) y[q] = (double)y32[q];
        record("S2_f32", y, n);
        CHECK(p_destroy(h));
        free(X32); free(Xc); free(a32); free(y32);
    }

    /* [S3] hyper-parameter loop: a new Gramian per length scale on the same x -> create, mul!, finalizer -> destroy.  Every
     * result must belong to ITS length scale (a cache keyed on x alone would return the first one every time). */
    {
        double* all = (double*)malloc(sizeof(double) * (size_t)(4 * n));
        for (int it = 0; it < 4; it++) {
            const cf_knode_t prog_ls[] = {{CF_OP_EQ, 0, 0.0}, {CF_OP_LENGTHSCALE, 0, 0.5 + 0.25 * it}};
            cf_gramian_t h = NULL;
            CHECK(p_create(&h, prog_ls, 2, CF_F64, (int)d, n, X, d, n, NULL, d));
            CHECK(p_mul(h, all + it * n, n, a, n, 1, 1.0, 0.0));
            CHECK(p_destroy(h));
        }
        record("S3_lengthscales", all, 4 * n);
        CHECK(p_destroy(NULL)); /* a finalizer on an already released handle passes C_NULL: must be harmless */
        free(all);
    }

    /* [S4] mul!(Y, G, X, alpha, beta) on matrices with a leading dimension larger than the row count (a view) and beta != 0 */
    {
        const int64_t ldY = n + 3;
        double* Yv = (double*)malloc(sizeof(double) * (size_t)(ldY * p));
        for (int64_t c = 0; c < p; c++)
            for (int64_t i = 0; i < ldY; i++) Yv[c * ldY + i] = (i < n) ? 0.25 * A[c * n + i] : -77.0;
        CHECK(p_mul(h2, Yv, ldY, A, n, p, -0.5, 2.0));
        record("S4_matrix", Yv, ldY * p);
        free(Yv);
    }

    /* [S5] blockmul!: flat (n d) vectors for GradientKernel, (n (d + 1)) for ValueGradientKernel, alpha / beta as given */
    for (int64_t q = 0; q < n * d; q++) y[q] = 0.5 * ag[q];
    CHECK(p_gradient_mul(h1, y, n * d, ag, n * d, 1, 1.5, -1.0));
    record("S5_gradient", y, n * d);
    CHECK(p_value_gradient_mul(h2, y, n * (d + 1), avg, n * (d + 1), 1, 1.0, 0.0));
    record("S5_value_gradient", y, n * (d + 1));

    /* [S6] (sigma^2 I + K) \ b: ldiv! on LazyMatrixSum(Diagonal, Gramian) -> cf_cg_solve with x = zeros (the reference's `\`
     * allocates zeros), reltol = 0 -> sqrt(eps), maxiter = 0 -> n */
    {
        int iters = -1;
        double res = -1;
        for (int64_t q = 0; q < n; q++) y[q] = 0.0;
        CHECK(p_cg_solve(h2, 1e-2, y, rhs, 0.0, 0, 0, &iters, &res));
        y[n] = (double)iters; y[n + 1] = res;
        record("S6_ldiv_lazysum", y, n + 2);
    }
    /* [S7] BlockGramian \ b: GradientKernel operator, sigma2 = 0, gradient = 1 (on a spread-out copy of the second point set: the
     * plain gradient Gramian of 384 standard-normal points is numerically singular) */
    {
        int iters = -1;
        double res = -1;
        double* Xw = (double*)malloc(sizeof(double) * (size_t)(n * d));
        for (int64_t q = 0; q < n * d; q++) Xw[q] = 4.0 * Y2[q];
        cf_gramian_t h = NULL;
        CHECK(p_create(&h, prog_eq, 1, CF_F64, (int)d, n, Xw, d, n, NULL, d));
        for (int64_t q = 0; q < n * d; q++) y[q] = 0.0;
        CHECK(p_cg_solve(h, 0.0, y, rhsg, 1e-10, 0, 1, &iters, &res));
        y[n * d] = (double)iters; y[n * d + 1] = res;
        record("S7_ldiv_blockgramian", y, n * d + 2);
        CHECK(p_destroy(h));
        free(Xw);
    }

    /* ARD(k, l): <k> ARDSCALE(l_1..l_d) ARD(d), as program!(p, ::Normed) emits it */
    {
        cf_knode_t* prog = (cf_knode_t*)malloc(sizeof(cf_knode_t) * (size_t)(d + 2));
        prog[0].op = CF_OP_MATERNP; prog[0].iparam = 2; prog[0].fparam = 0.0;
        for (int64_t c = 0; c < d; c++) { prog[1 + c].op = CF_OP_ARDSCALE; prog[1 + c].iparam = 0; prog[1 + c].fparam = l[c]; }
        prog[d + 1].op = CF_OP_ARD; prog[d + 1].iparam = (int32_t)d; prog[d + 1].fparam = 0.0;
        cf_gramian_t h = NULL;
        CHECK(p_create(&h, prog, (int)d + 2, CF_F64, (int)d, n, X, d, n, NULL, d));
        CHECK(p_mul(h, y, n, a, n, 1, 1.0, 0.0));
        record("S9_ard", y, n);
        CHECK(p_destroy(h));
        /* errors the shim maps to exceptions: wrong number of length scales -> DimensionMismatch (-2) with a message */
        prog[d].op = CF_OP_ARD; prog[d].iparam = (int32_t)d - 1; prog[d].fparam = 0.0; /* d - 1 length scales for d-dimensional points */
        h = NULL;
        double st[3];
        st[0] = (double)p_create(&h, prog, (int)d + 1, CF_F64, (int)d, n, X, d, n, NULL, d);
        st[1] = (double)p_create(&h, prog_eq, 1, CF_F64, (int)d, n, X, d - 1, n, NULL, d);           /* ldx < d */
        st[2] = (double)(strlen(p_last_error()) > 0);
        record("S9_errors", st, 3);
        free(prog);
    }

    /* [S10] Float32 Gramians through blockmul! and ldiv!: the shim's methods are generic in T = Float32 / Float64; inside the library the
     * derivative operators and the solves compute in Float64 on a copy of the points (vectors stay Float32).  reltol = 0 selects
     * sqrt(eps(Float32)), the cg! default for Float32 vectors. */
    {
        const int64_t nd = n * d;
        float* X32 = (float*)malloc(sizeof(float) * (size_t)nd);
        float* v32 = (float*)malloc(sizeof(float) * (size_t)nd);
        float* y32 = (float*)malloc(sizeof(float) * (size_t)nd);
        for (int64_t q = 0; q < nd; q++) { X32[q] = (float)(2.0 * Y2[q]); v32[q] = (float)ag[q]; y32[q] = 0.f; }
        cf_gramian_t h = NULL;
        CHECK(p_create(&h, prog_eq, 1, CF_F32, (int)d, n, X32, d, n, NULL, d));
        CHECK(p_gradient_mul(h, y32, nd, v32, nd, 1, 1.0, 0.0));
        for (int64_t q = 0; q < nd; q++) y[q] = (double)y32[q];
        record("S10_f32_gradient", y, nd);
        CHECK(p_destroy(h));
        h = NULL;
        CHECK(p_create(&h, prog_m2, 1, CF_F32, (int)d, n, X32, d, n, NULL, d));
        int iters = -1;
        double res = -1;
        for (int64_t q = 0; q < n; q++) { v32[q] = (float)rhs[q]; y32[q] = 0.f; }
        CHECK(p_cg_solve(h, 0.5, y32, v32, 0.0, 0, 0, &iters, &res));
        for (int64_t q = 0; q < n; q++) y[q] = (double)y32[q];
        y[n] = (double)iters; y[n + 1] = res;
        record("S10_f32_ldiv", y, n + 2);
        CHECK(p_destroy(h));
        free(X32); free(v32); free(y32);
    }

    CHECK(p_destroy(h1)); CHECK(p_destroy(h2)); CHECK(p_destroy(h3));

    /* [S8] init(devices): rows sharded inside the library over every visible device; handles created afterwards use all of them */
    {
        int ng = p_device_count();
        if (ng > 8) ng = 8;
        int devs[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        CHECK(p_init(ng, devs));
        cf_gramian_t h = NULL;
        CHECK(p_create(&h, prog_m2, 1, CF_F64, (int)d, n, X, d, n, NULL, d));
        CHECK(p_mul(h, y, n, a, n, 1, 1.0, 0.0));
        y[n] = (double)ng;
        record("S8_init_all_devices", y, n + 1);
        int iters = -1;
        double res = -1;
        for (int64_t q = 0; q < n; q++) y[q] = 0.0;
        CHECK(p_cg_solve(h, 1e-2, y, rhs, 0.0, 0, 0, &iters, &res));
        y[n] = (double)iters; y[n + 1] = res;
        record("S8_ldiv_all_devices", y, n + 2);
        CHECK(p_destroy(h));
    }
    fclose(g_out);
    printf("shim sequences ok\n");
    return 0;
}
