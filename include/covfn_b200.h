/*
 * covfn_b200.h -- C ABI of the B200-native lazy-Gramian MVM library (libcovfn_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of CovarianceFunctions.jl (reference v0.3.5):
 * the lazy `Gramian` matrix-vector / matrix-matrix multiply and the isotropic GradientKernel MVM.
 * The reference has no FFI for this path (it is Julia multiple dispatch on LinearAlgebra.mul!);
 * each entry point below names the reference method it replaces.  A Julia shim
 * (julia/CovarianceFunctionsB200.jl) adds more specific `mul!` methods that `ccall` these symbols;
 * the Python mirror (covariancefunctions.jl_b200/) binds the same symbols through ctypes.
 *
 * Conventions
 *   - every function returns an int status: 0 = CF_OK, negative = error (see cf_status); the message
 *     of the last error on the calling thread is available from cf_last_error().  No C++ exception
 *     crosses this boundary.
 *   - points are a contiguous d x n column-major array: X[c + ldx*i] is coordinate c of point i
 *     (ldx >= d), exactly the `Matrix` form accepted by gramian(k, X::AbstractMatrix)
 *     (reference src/gramian.jl:154).  Vectors are contiguous; multi-RHS matrices are column-major
 *     with a leading dimension; gradient vectors are flat with index i*d + c (reference
 *     src/gramian.jl:120-123, BlockFactorization(isstrided=true)).
 *   - dtype: CF_F32 / CF_F64 is the Gramian's eltype T (reference src/gramian.jl:30-33).
 *   - the caller owns all host buffers; the library never keeps a host pointer after a call returns.
 *   - there is NO CPU fallback: if no CUDA device is usable every compute entry point fails with
 *     CF_ERR_CUDA.
 */
#ifndef COVFN_B200_H
#define COVFN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_VERSION 100 /* 0.1.0 */

typedef enum {
    CF_OK = 0,
    CF_ERR_BAD_ARGUMENT = -1,  /* NULL pointer, negative size, bad dtype ...            */
    CF_ERR_DIMENSION = -2,     /* Julia DimensionMismatch (reference src/util.jl:9,41)   */
    CF_ERR_UNSUPPORTED = -3,   /* kernel program not lowerable to the device             */
    CF_ERR_DOMAIN = -4,        /* Julia DomainError (reference src/stationary.jl:19,47,124) */
    CF_ERR_CUDA = -5,          /* CUDA runtime / driver failure, or no device            */
    CF_ERR_NCCL = -6,
    CF_ERR_NONFINITE = -7,     /* NaN/Inf coordinate found at create time                */
    CF_ERR_INTERNAL = -8
} cf_status;

typedef enum { CF_F32 = 0, CF_F64 = 1 } cf_dtype;

/*
 * Kernel program: the reference's kernel object tree in postfix order.
 *   leaves push one value, SUM/PROD pop `iparam` values, POW / LENGTHSCALE pop one.
 * Reference definitions (all under /root/reference/src):
 *   EQ        exp(-r2/2)                          stationary.jl:42
 *   EXP       exp(-sqrt(r2))                      stationary.jl:60
 *   RQ        (1 + r2/(2a))^-a                    stationary.jl:53   fparam=a, iparam=1 if a is an Int
 *   MATERNP   nu = p + 1/2, closed form + Taylor  stationary.jl:134-158   iparam=p
 *   DOT       dot(x, y)                           mercer.jl:3,9
 *   CONST     c                                   stationary.jl:30-32     fparam=c
 *   SUM       sum of iparam children              algebra.jl:40
 *   PROD      product of iparam children          algebra.jl:17
 *   POW       child ^ iparam (Int power)          algebra.jl:62
 *   LENGTHSCALE  child(r2 / l^2), child isotropic leaf   transformation.jl:19   fparam=l
 *   ARDSCALE  pushes one length-scale entry l_c                                  fparam=l_c
 *   ARD       pops iparam = d ARDSCALE entries and one isotropic child k: k(sum_c (x_c - y_c)^2 / l_c), i.e.
 *             ARD(k, l::AbstractVector) = Normed(k, tau -> enorm2(Diagonal(inv.(l)), tau))   transformation.jl:25-45, util.jl:52
 *             (postfix: <child nodes> ARDSCALE(l_1) ... ARDSCALE(l_d) ARD(d)).  The library applies the metric by scaling the
 *             point coordinates with 1/sqrt(l_c) on the device while uploading them (as the reference pre-transforms the data of
 *             its input-scaling kernels, transformation.jl:83-95), so every r2-kernel of the program must sit under ONE ARD with
 *             ONE l and the program may not contain DOT; anything else is CF_ERR_UNSUPPORTED (the reference method runs).
 *             ARD programs have the StationaryInput trait: no derivative operators (the reference uses its generic fallback).
 */
typedef enum {
    CF_OP_EQ = 1,
    CF_OP_EXP = 2,
    CF_OP_RQ = 3,
    CF_OP_MATERNP = 4,
    CF_OP_DOT = 5,
    CF_OP_CONST = 6,
    CF_OP_SUM = 7,
    CF_OP_PROD = 8,
    CF_OP_POW = 9,
    CF_OP_LENGTHSCALE = 10,
    CF_OP_ARDSCALE = 11,
    CF_OP_ARD = 12
} cf_op;

typedef struct {
    int32_t op;     /* cf_op */
    int32_t iparam; /* arity / integer power / p / int-flag */
    double fparam;  /* alpha / c / l */
} cf_knode_t;

typedef struct cf_gramian_s* cf_gramian_t; /* opaque handle: device copies of X, Y + lowered program */

/* ---- library ---------------------------------------------------------------------------- */
int cf_version(void);
const char* cf_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing); never fails */
int cf_device_count(void);
/* Select the devices later handles shard their rows over (single-process multi-GPU mode).
 * devices == NULL -> devices 0..ngpus-1.  Optional: the default is {current device}. */
int cf_init(int ngpus, const int* devices);

/* ---- lazy Gramian ----------------------------------------------------------------------- */
/*
 * Replaces gramian(k, x, y) / Gramian(k, x, y)  (reference src/gramian.jl:18-21, 144-148).
 * O(n d) work: validates the program, copies X (n points) and Y (m points) to every device.
 * Y == NULL means y === x (symmetric case).
 */
int cf_gramian_create(cf_gramian_t* out, const cf_knode_t* prog, int nnodes, int dtype, int d,
                      int64_t n, const void* X, int64_t ldx, int64_t m, const void* Y, int64_t ldy);
int cf_gramian_destroy(cf_gramian_t g);
int cf_gramian_size(cf_gramian_t g, int64_t* n, int64_t* m, int* d, int* dtype);

/*
 * Restrict the rows this handle computes to [row_begin, row_end) (multi-process row sharding:
 * one process per GPU, each owning a contiguous row block).  Output pointers
 * passed to the *_mul functions then address the shard's first row.
 */
int cf_gramian_set_row_range(cf_gramian_t g, int64_t row_begin, int64_t row_end);

/*
 * Options.  CF_OPT_SYMMETRIC (default 1; the environment variable COVFN_SYMMETRIC=0 at create time turns it off): for y === x,
 * Float64, nrhs == 1 and n >= 32768, evaluate every unordered pair {i, j} once and use it for both b_i and b_j (csrc/
 * gram_mvm_sym.cuh).  Halves the kernel evaluations of a symmetric Gramian MVM.  Every partial sum has a single writer and the
 * partials are combined in a fixed order, so the result is bit-reproducible; it differs from the all-pairs path by summation
 * order only (<= 1e-13).  The reference always evaluates all n*m entries (src/gramian.jl:78-87); set the option to 0 for that.
 */
#define CF_OPT_SYMMETRIC 1
int cf_gramian_set_option(cf_gramian_t g, int option, int value);

/*
 * y <- alpha * K * x + beta * y, beta == 0 overwrites y (NaN-safe).
 * nrhs == 1 replaces mul!(y::AbstractVector, G::Gramian, x::AbstractVector, alpha, beta)
 *   (reference src/gramian.jl:78-87);
 * nrhs > 1 replaces mul!(Y::AbstractMatrix, G::Gramian, X::AbstractMatrix, alpha, beta)
 *   (reference src/gramian.jl:89-99): x is m x nrhs (leading dim ldx), y is n x nrhs (ldy).
 * HOST pointers; the call copies x in, y out (and y in when beta != 0) and blocks until done.
 */
int cf_gramian_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs,
                   double alpha, double beta);
/* Same contract with DEVICE pointers on the handle's (first) device; asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = the library's own stream, then the call blocks).  A handle supports ONE in-flight
 * caller stream: its scratch buffers are shared between calls, and cf_gramian_destroy waits for the stream of the last
 * asynchronous call before the memory returns to the pool. */
int cf_gramian_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx,
                          int64_t nrhs, double alpha, double beta, void* stream);

/*
 * Dense instantiation M[i + ldm*j] = k(x_i, y_j): replaces Matrix!(M, G)
 * (reference src/gramian.jl:107-114).  HOST pointer.
 */
int cf_gramian_matrix(cf_gramian_t g, void* M, int64_t ldm);
/* single entry, replaces getindex(G, i, j) (reference src/gramian.jl:37-40); 0-based i, j */
int cf_gramian_getindex(cf_gramian_t g, int64_t i, int64_t j, double* out);

/* ---- derivative kernels: GradientKernel and ValueGradientKernel -------------------------------------- */
/*
 * y <- alpha * G * x + beta * y for G = gramian(GradientKernel(k), X[, Y]), the (n d) x (m d) operator whose d x d
 * block (i, j) is
 *     -2 (k' I + 2 k'' r r^T),  r = x_i - y_j,  k', k'' derivatives of k in r^2   (IsotropicInput kernels)
 *     k' I + k'' y_j x_i^T,     k', k'' derivatives of k in t = x_i . y_j         (DotProductInput kernels)
 * Replaces blockmul!(y, G::Gramian, x, alpha, beta) (reference src/gramian.jl:241-253) with the lazy
 * IsotropicGradientKernelElement / DotProductGradientKernelElement mul! (reference src/gradient.jl:86-92, 109-115) and
 * derivative_laplacian (src/gradient.jl:589-600).  The handle's program must have the IsotropicInput or the
 * DotProductInput trait (reference src/properties.jl:39-63).
 * x: (m d) x nrhs, y: (n d) x nrhs, flat index i*d + c.  HOST pointers.  Float32 handles: vectors are Float32, the arithmetic
 * runs on a Float64 copy of the points built on first use (as do cf_cg_solve and every product with d > 32).
 */
int cf_gradient_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs,
                    double alpha, double beta);
int cf_gradient_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx,
                           int64_t nrhs, double alpha, double beta, void* stream);
/*
 * The same for G = gramian(ValueGradientKernel(k), X[, Y]): (d+1) x (d+1) blocks [vv vg; gv gg] with entry 0 the value
 * observation (reference src/gradient.jl:400-474, DerivativeKernelElement :217-239): flat index i*(d+1) + e.
 */
int cf_value_gradient_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs,
                          double alpha, double beta);
int cf_value_gradient_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx,
                                 int64_t nrhs, double alpha, double beta, void* stream);

/* ---- chained MVMs: conjugate gradients on (K + sigma2 I) x = b -------------------------- */
/*
 * Replaces  (sigma2*I + K) \ b : LazyMatrixSum(Diagonal, Gramian) -> ldiv! -> IterativeSolvers.cg!
 * (reference src/gramian.jl:55-60, src/lazy_linear_algebra.jl:126-144).  x holds the initial guess
 * on entry (an initial residual MVM is always done, as cg! does) and the solution on exit.
 * reltol <= 0 selects sqrt(eps(T)); maxiter <= 0 selects n.  Iterates stay on the device(s); with
 * several devices each owns a row block and the search direction is re-assembled once per iteration.
 * gradient = 1 solves with the GradientKernel operator instead, gradient = 2 with the ValueGradientKernel operator
 * (reference src/gramian.jl:229-238).
 */
int cf_cg_solve(cf_gramian_t g, double sigma2, void* x, const void* b, double reltol, int maxiter,
                int gradient, int* iters, double* resnorm);

/* device / NCCL time of the last cf_cg_solve on this handle: host wall clock of the call, device time of the operator products
 * and of the row-block all-gathers (multi-process mode only), number of operator products (iterations + 1) */
int cf_cg_timing(cf_gramian_t g, double* total_ms, double* product_ms, double* gather_ms, int* products);

/* ---- multi-process row sharding: one process per GPU, NCCL over NVLink --------------------- */
/*
 * The reference is single-process; rows of K are independent (src/gramian.jl:81,244), so across GPUs every rank owns the
 * contiguous row block [n r / world, n (r + 1) / world) (cf_gramian_set_row_range) and chained products (CG) re-assemble the
 * vector with ONE all-gather per product.  libnccl.so.2 is dlopen()ed on first use (COVFN_NCCL_LIB overrides the name).
 * Bootstrap: rank 0 calls cf_comm_unique_id, the host ships the 128 bytes to the other ranks, every rank calls cf_comm_init
 * with its CUDA device current.  Afterwards cf_cg_solve accepts a handle restricted to the rank's block: iterates are replicated
 * and updated redundantly (bit-identical on every rank), the product's row blocks are gathered in place.
 */
int cf_comm_unique_id(void* id128, int bytes);
int cf_comm_init(const void* id128, int rank, int world);
int cf_comm_destroy(void);
int cf_comm_info(int* rank, int* world, int* nccl_version);
/* y_full <- alpha K x + beta y_full computed collectively by all ranks (same x on every rank, complete y on every rank; device
 * pointers, one right-hand side).  Symmetric Float64 Gramians: each rank evaluates the unordered pairs of its row tiles, then
 * ncclAllReduce(sum); otherwise the rank's row block followed by the in-place all-gather. */
int cf_gramian_mul_collective_device(cf_gramian_t g, void* d_y_full, const void* d_x, double alpha, double beta, void* stream);
/* in-place all-gather of a device vector of n blocks of `block` elements whose rank-r part is rows [n r / world, n (r+1) / world) */
int cf_comm_allgather_rows(void* d_full, int64_t n, int64_t block, int dtype, void* stream);

/* ---- measurement helpers (used by bench.py; not part of the reference surface) ------------ */
/* timing of the last *_mul call on this handle: device ms of the dominant kernel and its launch count */
int cf_last_timing(cf_gramian_t g, float* kernel_ms, int* launches);
/* measured pipe peaks on the current device: issues `iters` dependent-chain-free FMAs per thread.
 * kind: 0 = FP64 DFMA, 1 = FP32 FFMA, 2 = MUFU.EX2.  Returns lane-instructions per second. */
int cf_peak_probe(int kind, int iters, double* lane_ops_per_s, float* ms);
/* Run-time specialisation of composite kernel programs (csrc/cf_jit.h): the reference gets a fused evaluation of every
 * kernel composition from Julia's compiler (src/algebra.jl:17,40,62 inline per concrete Sum/Product type); here large
 * multi-RHS products re-compile the same kernel source with the program STRUCTURE as compile-time constants (NVRTC,
 * once per structure; hyper-parameters stay run-time arguments).  Counters since process start; any pointer may be NULL.
 * Environment: COVFN_JIT=0 never, =1 always, unset = calls evaluating >= 2^33 entries. */
int cf_jit_stats(int* compiled, int* cache_hits, int* failures, double* compile_seconds);
/* Compile (only) the run-time specialisation of one kernel for a program, without a GPU: the "does the generated code build"
 * check.  which: 0 value MVM (K1), 1 Float64 tensor-core multi-RHS (K4d), 2 Float64 tensor-core MVM (K1d), 3 Float32 3xTF32
 * multi-RHS on tcgen05 (K4u), 4 Float32 3xTF32 MVM with mma.sync (K1t), 5 Float64 tensor-core gradient MVM with generated jets (K5d),
 * 6 Float32 3xTF32 multi-RHS with mma.sync (K4t), 7 Float32 3xTF32 MVM on tcgen05 (K1u); d = point dimension.  log (may be NULL) receives the NVRTC log and, on failure,
 * the generated evaluator.  CF_ERR_UNSUPPORTED if NVRTC is missing or the kernel does not exist for this d. */
int cf_jit_check(const cf_knode_t* prog, int nnodes, int d, int which, char* log, int loglen);

#ifdef __cplusplus
}
#endif
#endif /* COVFN_B200_H */
