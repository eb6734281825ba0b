// capi.cu -- C ABI of libcovfn_b200.so (include/covfn_b200.h): handles, validation, kernel-program lowering,
// launch planning, multi-device row sharding.  No C++ exception crosses the boundary; there is no CPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

// this translation unit holds the non-hot kernels (dense instantiation, d > 32): they use the two-step exp reduction
#define CF_EXP_ACCURATE 1
#include "../../include/covfn_b200.h"
#include "cf_lower.h"
#include "cf_registry.h"
#include "cf_jit.h"
#include "cf_comm.h"
#include "cf_extra.cuh"
#include "bigd.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CF_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) return fail(CF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// Handle-owned device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release threshold
// raised to "never": creating and destroying Gramian handles (one per mul! in the end-to-end path) then costs no
// cudaMalloc / cudaFree, whose implicit device synchronisation and unmapping were measured at up to 0.7 s per destroy.
// Allocation is followed by a synchronisation of the null stream's ordering point, so the pointer may be used on any stream.
inline cudaError_t dev_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes, (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize((cudaStream_t)0);
}
inline void dev_free(void* p) {
    if (p) cudaFreeAsync(p, (cudaStream_t)0);
}

// ---- devices --------------------------------------------------------------------------------------------
struct DeviceCtx {
    int dev = -1;
    double* exp2_tbl = nullptr;
    int sms = 0;
};
std::mutex g_mu;
std::vector<int> g_devices;          // devices new handles shard over (empty: current device)
std::vector<DeviceCtx> g_ctx;        // lazily created, indexed by position

int get_ctx(int dev, DeviceCtx** out) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& c : g_ctx)
        if (c.dev == dev) { *out = &c; return CF_OK; }
    DeviceCtx c;
    c.dev = dev;
    CF_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CF_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(CF_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", dev, prop.major, prop.minor);
    c.sms = prop.multiProcessorCount;
    {
        cudaMemPool_t pool;
        CF_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull; // keep freed blocks cached in the pool
        CF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    double tbl[CF_EXP_TBL];
    for (int j = 0; j < CF_EXP_TBL; j++) {
        // 2^(j/256) correctly rounded, with (j << 12) subtracted from its high word: cf_exp_cv adds kk << 12 = (k << 20) + (j << 12)
        // to the high word, which then holds the exponent of 2^k 2^(j/256) (one integer instruction instead of mask + add)
        union { double d; uint64_t u; } v;
        v.d = (double)exp2l((long double)j / CF_EXP_TBL);
        v.u -= (uint64_t)j << (32 + 20 - CF_EXP_TBL_BITS);
        tbl[j] = v.d;
    }
    CF_CUDA(cudaMalloc(&c.exp2_tbl, sizeof(tbl)));
    CF_CUDA(cudaMemcpy(c.exp2_tbl, tbl, sizeof(tbl), cudaMemcpyHostToDevice));
    g_ctx.reserve(64);
    g_ctx.push_back(c);
    *out = &g_ctx.back();
    return CF_OK;
}

const cf_kernel_entry* find_entry(int d) {
    const cf_kernel_entry* all[] = {cf_kernels_d1(), cf_kernels_d2(), cf_kernels_d3(),  cf_kernels_d4(),  cf_kernels_d6(),
                                    cf_kernels_d8(), cf_kernels_d12(), cf_kernels_d16(), cf_kernels_d24(), cf_kernels_d32()};
    for (auto* e : all)
        if (e->D >= d) return e;
    return nullptr;
}

// ---- handle ---------------------------------------------------------------------------------------------
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CF_OK;
        // growing a scratch buffer returns the old block to the pool: nothing -- on the library's streams or on a caller's stream of an
        // asynchronous *_mul_device call -- may still be using it.  Growth only happens on the first calls of a handle, so a
        // device-wide synchronisation here is cheap and makes the free safe on any stream.
        if (p) cudaDeviceSynchronize();
        dev_free(p);
        p = nullptr;
        cap = 0;
        CF_CUDA(dev_alloc(&p, bytes));
        cap = bytes;
        return CF_OK;
    }
    void release() { dev_free(p); p = nullptr; cap = 0; }
};

struct Shard {
    DeviceCtx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* X = nullptr;  // n x D padded points (all rows, replicated)
    void* Y = nullptr;  // m x D (== X when symmetric)
    void* xn = nullptr;  // squared norms of the padded points (multi-RHS kernel)
    void* yn = nullptr;
    Buf a, y, partial, apad, ypad, at, cg[6], sym_items, bsym, sym_col, bd_t, bd_s, xp, yp, ap, yc_hi, yc_lo, yc_n, ac_hi, ac_lo, yt, au, cv_in, cv_out;
    int64_t yt_ld = 0;      // > 0: yt holds the transposed Float32 column points with this leading dimension (gram_mvm_f32p.cuh)
    bool tc5_ready = false; // yc_* hold the canonical column-point images of the tcgen05 multi-RHS kernel
    bool mmd_ready = false; // xp / yp hold the padded point copies of the DMMA multi-RHS kernel
    int sym_nitems = -1;    // symmetric variant: work items (-1: not built), row tile, chunk length and device share they were built for
    int64_t sym_tr = 0, sym_ch = 0;
    int sym_part = 0, sym_parts = 1;
    int64_t r0 = 0, r1 = 0; // rows owned
};

}  // namespace

struct cf_gramian_s {
    int dtype = CF_F64, d = 0, D = 0;
    int64_t n = 0, m = 0;
    bool symmetric = false;
    bool use_norms = false; // multi-RHS kernel may use r2 = |x|^2 + |y|^2 - 2 x.y (well-scaled data, d >= 8)
    bool use_norms_grad = false; // ... and so may the isotropic gradient operator (stricter: needs k'')
    bool eq_fast = false;   // single EQ atom, any d: norm expansion error < 1e-13 and |c| (|x| + |y|)^2 < 600 (gram_mvm_eq.cuh)
    bool mat_fast = false;  // single MaternP atom, p >= 1, any d: norm expansion error < 1e-13 and |c| (|x| + |y|) < 600 (gram_mvm_eq.cuh, FAST = 2)
    int64_t row_begin = 0, row_end = 0;
    cf_program prog;
    cf_sop_val sop_val;    // parameter-resident program for the value kernels
    cf_sop_grad sop_grad;  // ... for the gradient kernel (valid when grad_ok)
    bool grad_ok = false;
    int kind = CF_ATOM_SOP; // kernel kind used for the value MVM
    double coef = 1.0;      // leading constant when prog.single (value kernels fold it into alpha)
    double coef_grad = 1.0; // the same constant for the derivative kernels (always the term's coefficient)
    const cf_kernel_entry* entry = nullptr;
    std::vector<Shard> shards;
    std::mutex mu;
    bool opt_symmetric = true;  // use the symmetric variant (each unordered pair evaluated once) when applicable
    float last_ms = 0;
    int last_launches = 0;
    void* user_stream = nullptr;  // stream of the last asynchronous *_mul_device call (one in-flight user stream per handle)
    // Float32 handles: a Float64 copy of the points (same devices, same row range), built on first use, on which the operators that
    // exist in Float64 only run -- derivative kernels, conjugate gradients, d > 32.  The reference is generic in T (src/gradient.jl:86-92);
    // here Float32 inputs and outputs are converted on the fly and the arithmetic in between is Float64 (at least as accurate).
    cf_gramian_s* shadow64 = nullptr;
    double max_sq = 0;            // largest squared point norm (scale checks)
    // timing of the last cf_cg_solve (cf_cg_timing): host wall clock of the whole solve, device time of the operator products and
    // of the NCCL row-block gathers (multi-process mode), number of operator products
    double cg_total_ms = 0, cg_mvm_ms = 0, cg_gather_ms = 0;
    int cg_products = 0;
};

namespace {

size_t esize(int dtype) { return dtype == CF_F64 ? 8 : 4; }

void split_rows(cf_gramian_s* g) {
    const int S = (int)g->shards.size();
    const int64_t rows = g->row_end - g->row_begin;
    for (int s = 0; s < S; s++) {
        g->shards[s].r0 = g->row_begin + rows * s / S;
        g->shards[s].r1 = g->row_begin + rows * (s + 1) / S;
    }
}

// choose the number of column chunks so that the grid fills the machine in (nearly) whole waves
struct Plan { int row_tiles; int chunks; int64_t cols_per_chunk; };
Plan make_plan(int64_t nrows, int64_t m, const cf_mvm_config& cfg, int sms) {
    Plan p;
    p.row_tiles = (int)((nrows + cfg.rows_per_cta - 1) / cfg.rows_per_cta);
    if (p.row_tiles < 1) p.row_tiles = 1;
    const int64_t col_tiles = std::max<int64_t>(1, (m + cfg.tj - 1) / cfg.tj);
    const double conc = (double)sms * cfg.min_blocks;
    int64_t smin = (int64_t)std::ceil(6.0 * conc / p.row_tiles);
    smin = std::max<int64_t>(1, std::min<int64_t>(smin, col_tiles));
    int64_t best = smin;
    double best_eff = -1;
    for (int64_t s = smin; s <= std::min<int64_t>(col_tiles, 2 * smin + 4); s++) {
        const int64_t cpc = ((col_tiles + s - 1) / s) * cfg.tj;
        const int64_t real_s = (m + cpc - 1) / cpc;
        const double w = (double)p.row_tiles * real_s / conc;
        const double eff = w / std::ceil(w);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    p.cols_per_chunk = ((col_tiles + best - 1) / best) * cfg.tj;
    p.chunks = (int)std::max<int64_t>(1, (m + p.cols_per_chunk - 1) / p.cols_per_chunk);
    return p;
}

int check_handle(cf_gramian_t g) {
    if (!g) return fail(CF_ERR_BAD_ARGUMENT, "NULL Gramian handle");
    return CF_OK;
}

int launch_scale(int dtype, void* y, const void* yin, int64_t n, double beta, cudaStream_t stream) {
    if (n <= 0) return CF_OK;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 4096);
    if (dtype == CF_F64) cf_scale_kernel<double><<<blocks, 256, 0, stream>>>((double*)y, (const double*)yin, n, beta);
    else cf_scale_kernel<float><<<blocks, 256, 0, stream>>>((float*)y, (const float*)yin, n, beta);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

// dense tile: rows of the shard x columns [j0, j0 + nj), column-major with leading dimension ld
int launch_dense(cf_gramian_s* g, Shard& sh, void* d_M, int64_t ld, int64_t j0, int64_t nj, cudaStream_t stream) {
    const int64_t nrows = sh.r1 - sh.r0;
    if (nrows <= 0 || nj <= 0) return CF_OK;
    const int blocks = (int)std::min<int64_t>((nrows * nj + 255) / 256, 148 * 16);
    if (g->dtype == CF_F64)
        gram_dense_kernel<double><<<blocks, 256, 0, stream>>>((const double*)sh.X, (const double*)sh.Y, g->D, g->sop_val,
                                                              sh.ctx->exp2_tbl, sh.r0, nrows, j0, nj, (double*)d_M, ld);
    else
        gram_dense_kernel<float><<<blocks, 256, 0, stream>>>((const float*)sh.X, (const float*)sh.Y, g->D, g->sop_val,
                                                             sh.ctx->exp2_tbl, sh.r0, nrows, j0, nj, (float*)d_M, ld);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

// point copies with the row stride (= 4 mod 8 doubles) the DMMA kernels read fragments from, built once per shard
int ensure_padded_points(cf_gramian_s* g, Shard& sh, cudaStream_t stream) {
    if (sh.mmd_ready) return CF_OK;
    const int sx = (g->D % 8 == 4) ? g->D + 8 : g->D + 4;  // cf_mmd_smem<D>::sx
    if (int rc = sh.xp.ensure((size_t)g->n * sx * 8)) return rc;
    cf_pad_rows_kernel<<<148 * 8, 256, 0, stream>>>((const double*)sh.X, g->D, sx, g->n, (double*)sh.xp.p);
    if (sh.Y != sh.X) {
        if (int rc = sh.yp.ensure((size_t)g->m * sx * 8)) return rc;
        cf_pad_rows_kernel<<<148 * 8, 256, 0, stream>>>((const double*)sh.Y, g->D, sx, g->m, (double*)sh.yp.p);
    }
    CF_CUDA(cudaGetLastError());
    sh.mmd_ready = true;
    return CF_OK;
}

int ensure_padded_points_f32(cf_gramian_s* g, Shard& sh, cudaStream_t stream);

static bool env_flag(const char* name) {
    const char* e = std::getenv(name);
    return e && std::atoi(e) != 0;
}

// canonical (tensor-core operand layout) hi / lo images of the column points and their zero-padded squared norms, built once per shard:
// the B operands of the distance GEMMs of the tcgen05 kernels (gram_mm_tc5.cuh, gram_mvm_tc5.cuh)
int ensure_canonical_points(cf_gramian_s* g, Shard& sh, cudaStream_t stream) {
    if (sh.tc5_ready) return CF_OK;
    const int dk = g->entry->mm_tc5_dk;
    const int64_t mpad = ((g->m + CF_MMU_TJ - 1) / CF_MMU_TJ) * CF_MMU_TJ;
    if (int rc = sh.yc_hi.ensure((size_t)mpad * dk * 4)) return rc;
    if (int rc = sh.yc_lo.ensure((size_t)mpad * dk * 4)) return rc;
    if (int rc = sh.yc_n.ensure((size_t)mpad * 4)) return rc;
    cf_canon_points_kernel<<<148 * 8, 256, 0, stream>>>((const float*)sh.Y, g->D, dk, g->m, mpad, (float*)sh.yc_hi.p, (float*)sh.yc_lo.p,
                                                        (const float*)sh.yn, (float*)sh.yc_n.p);
    CF_CUDA(cudaGetLastError());
    sh.tc5_ready = true;
    return CF_OK;
}

// B <- alpha K A + beta B, device pointers, column-major with leading dimensions (elements)
int launch_mm(cf_gramian_s* g, Shard& sh, void* d_B, int64_t ldb, const void* d_A, int64_t lda, int64_t nrhs, double alpha,
              double beta, cudaStream_t stream) {
    const int64_t nrows = sh.r1 - sh.r0;
    if (nrows <= 0) return CF_OK;
    const size_t es = esize(g->dtype);
    if (g->m == 0) {
        for (int64_t c = 0; c < nrhs; c++)
            if (int rc = launch_scale(g->dtype, (char*)d_B + c * ldb * es, (char*)d_B + c * ldb * es, nrows, beta, stream)) return rc;
        return CF_OK;
    }
    // Float64, well-scaled points, d >= 8: FP64 tensor-core kernel (gram_mm_dmma.cuh); COVFN_MM_SCALAR=1 forces the scalar one
    const bool dmma = g->dtype == CF_F64 && g->use_norms && g->entry->mm_dmma != nullptr && !env_flag("COVFN_MM_SCALAR");
    // Float32, well-scaled points, d >= 8: tensor cores in 3xTF32 split precision (gram_mm_tf32.cuh)
    const bool tf32 = g->dtype == CF_F32 && g->use_norms && g->entry->mm_tf32 != nullptr && !env_flag("COVFN_MM_SCALAR");
    // ... on the 5th-generation tensor cores (tcgen05 / TMEM, gram_mm_tc5.cuh); COVFN_MM_LEGACY=1 keeps the mma.sync kernel
    if (tf32 && g->entry->mm_tc5 != nullptr && !env_flag("COVFN_MM_LEGACY")) {
        const int dk = g->entry->mm_tc5_dk;
        const int64_t ntiles = (g->m + CF_MMU_TJ - 1) / CF_MMU_TJ, mpad = ntiles * CF_MMU_TJ;
        if (int rc = ensure_canonical_points(g, sh, stream)) return rc;
        if (int rc = sh.ac_hi.ensure((size_t)mpad * CF_MMU_PC * 4)) return rc;
        if (int rc = sh.ac_lo.ensure((size_t)mpad * CF_MMU_PC * 4)) return rc;
        cf_mmu_params PP;
        std::memset(&PP, 0, sizeof(PP));
        cf_mm_params& Q = PP.mm;
        Q.X = sh.X; Q.xn = sh.xn; Q.sop = g->sop_val;
        Q.row0 = sh.r0; Q.nrows = nrows; Q.m = g->m; Q.ldb = ldb; Q.alpha = alpha; Q.beta = beta; Q.use_norms = 1;
        PP.yhi = (const float*)sh.yc_hi.p; PP.ylo = (const float*)sh.yc_lo.p; PP.ynpad = (const float*)sh.yc_n.p;
        PP.ahi = (const float*)sh.ac_hi.p; PP.alo = (const float*)sh.ac_lo.p; PP.ntiles = ntiles;
        const int row_tiles5 = (int)((nrows + CF_MMU_TI - 1) / CF_MMU_TI);
        cfjit::Kernel* jit5 = nullptr;  // large products: the same kernel source with the program structure compiled in (cf_jit.h)
        if (cfjit::wanted((double)nrows * (double)g->m))
            jit5 = cfjit::get_kernel(g->sop_val, "gram_mm_tc5.cuh", "gram_mm_tc5_kernel<" + std::to_string(g->D) + ">");
        for (int64_t c0 = 0; c0 < nrhs; c0 += CF_MMU_PC) {
            Q.nrhs = (int)std::min<int64_t>(CF_MMU_PC, nrhs - c0);
            cf_canon_rhs_kernel<<<148 * 8, 256, 0, stream>>>((const float*)d_A + c0 * lda, lda, g->m, mpad, Q.nrhs, (float*)sh.ac_hi.p, (float*)sh.ac_lo.p);
            CF_CUDA(cudaGetLastError());
            Q.B = (char*)d_B + c0 * ldb * es;
            bool launched = false;
            if (jit5) {
                launched = cfjit::launch(jit5, &PP, (unsigned)row_tiles5, 1, CF_MMU_THREADS, (unsigned)g->entry->mm_tc5_smem, stream) == 0;
                if (!launched) jit5 = nullptr;
            }
            if (!launched) CF_CUDA(g->entry->mm_tc5(PP, row_tiles5, stream));
            g->last_launches += 2;
        }
        return CF_OK;
    }
    const int ldat = dmma ? CF_MMD_SA : (tf32 ? CF_MMT_SA : CF_MM_PC);
    if (int rc = sh.at.ensure((size_t)g->m * ldat * es)) return rc;
    cf_mm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.X; P.Y = sh.Y; P.xn = sh.xn; P.yn = sh.yn; P.At = sh.at.p;
    P.exp2_tbl = sh.ctx->exp2_tbl; P.sop = g->sop_val;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m; P.ldb = ldb;
    P.alpha = alpha; P.beta = beta;
    P.use_norms = g->use_norms ? 1 : 0;
    P.ldat = ldat;
    if (dmma) {
        if (int rc = ensure_padded_points(g, sh, stream)) return rc;
        P.X = sh.xp.p;
        P.Y = (sh.Y != sh.X) ? sh.yp.p : sh.xp.p;
    }
    if (tf32) {  // the column points with the padded row stride (B fragments); the row tile is transposed inside the kernel
        if (int rc = ensure_padded_points_f32(g, sh, stream)) return rc;
        P.X = sh.X;
        P.Y = sh.yp.p;
    }
    const int row_tiles = (int)((nrows + CF_MM_TI - 1) / CF_MM_TI);
    cfjit::Kernel* jit = nullptr;
    if (dmma && cfjit::wanted((double)nrows * (double)g->m))
        jit = cfjit::get_kernel(g->sop_val, "gram_mm_dmma.cuh", "gram_mm_dmma_kernel<" + std::to_string(g->D) + ">");
    if (tf32 && cfjit::wanted((double)nrows * (double)g->m))
        jit = cfjit::get_kernel(g->sop_val, "gram_mm_tf32.cuh", "gram_mm_tf32_kernel<" + std::to_string(g->D) + ">");
    const unsigned jit_smem = (unsigned)(tf32 ? g->entry->mm_tf32_smem : g->entry->mm_dmma_smem);
    for (int64_t c0 = 0; c0 < nrhs; c0 += CF_MM_PC) {
        P.nrhs = (int)std::min<int64_t>(CF_MM_PC, nrhs - c0);
        const dim3 tg((unsigned)((g->m + 31) / 32), CF_MM_PC / 32), tb(32, 8);
        if (g->dtype == CF_F64)
            cf_transpose_rhs<double><<<tg, tb, 0, stream>>>((const double*)d_A + c0 * lda, lda, g->m, P.nrhs, (double*)sh.at.p, ldat);
        else
            cf_transpose_rhs<float><<<tg, tb, 0, stream>>>((const float*)d_A + c0 * lda, lda, g->m, P.nrhs, (float*)sh.at.p, ldat);
        CF_CUDA(cudaGetLastError());
        P.B = (char*)d_B + c0 * ldb * es;
        bool launched = false;
        if (jit) {  // same kernel source, program structure compiled in (cf_jit.h); any failure falls back to the interpreter build
            launched = cfjit::launch(jit, &P, (unsigned)row_tiles, 1, 256, jit_smem, stream) == 0;
            if (!launched) jit = nullptr;
        }
        if (!launched) CF_CUDA((dmma ? g->entry->mm_dmma : (tf32 ? g->entry->mm_tf32 : g->entry->mm[g->dtype]))(P, row_tiles, stream));
        g->last_launches += 2;
    }
    return CF_OK;
}

int launch_mvm(cf_gramian_s* g, Shard& sh, void* d_y, const void* d_yin, const void* d_a, double alpha, double beta,
               cudaStream_t stream, const cf_peer_out* peers = nullptr);
int launch_grad(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                cudaStream_t stream, int vg, const cf_peer_out* peers = nullptr);
int launch_mvm_sym(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                   cudaStream_t stream, const cf_peer_out* peers, int part, int parts);
bool sym_applicable(cf_gramian_s* g, const void* d_a);

// peer access between all devices of a handle (single-process multi-GPU); handle memory comes from the stream-ordered pool, whose
// blocks must be made peer-accessible explicitly.  Returns 0 when every pair can access each other, 1 otherwise.
int enable_peers(cf_gramian_s* g) {
    const int S = (int)g->shards.size();
    if (S > 8) return 1;
    for (int s = 0; s < S; s++)
        for (int t = 0; t < S; t++) {
            if (s == t) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->shards[s].ctx->dev, g->shards[t].ctx->dev) != cudaSuccess || !can) return 1;
        }
    for (int s = 0; s < S; s++) {
        if (cudaSetDevice(g->shards[s].ctx->dev) != cudaSuccess) return 1;
        for (int t = 0; t < S; t++) {
            if (s == t) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(g->shards[t].ctx->dev, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); return 1; }
            cudaMemPool_t pool;
            cudaMemAccessDesc desc;
            std::memset(&desc, 0, sizeof(desc));
            desc.location.type = cudaMemLocationTypeDevice;
            desc.location.id = g->shards[t].ctx->dev;
            desc.flags = cudaMemAccessFlagsProtReadWrite;
            if (cudaDeviceGetDefaultMemPool(&pool, g->shards[s].ctx->dev) != cudaSuccess ||
                cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) { cudaGetLastError(); return 1; }
        }
    }
    return 0;
}

// Multi-GPU CG, SPMD inside one process: every device holds the full iterates and updates them redundantly (identical
// arithmetic, so identical decisions); each device computes its row block of A u and the kernel epilogue stores the block
// straight into every peer's copy over NVLink (cf_peer_out) -- the all-gather is fused into the producing kernel, there is no
// staging through device 0 and no separate collective.  Cross-device ordering uses two events per device: "block written"
// (peers wait for it before reading the assembled vector) and "vector consumed" (peers wait for it before overwriting it).
// Returns 1 if peer access is unavailable (caller falls back to the gather-through-device-0 path).
int cg_solve_spmd(cf_gramian_s* g, double sigma2, double* x, const double* b, double reltol, int maxiter, int deriv, int* iters,
                  double* resnorm) {
    const int S = (int)g->shards.size();
    const bool gradient = deriv != 0;
    const int vg = deriv == 2 ? 1 : 0;
    const int64_t blk = deriv == 0 ? 1 : g->d + vg;
    const int64_t N = g->n * blk;
    if (S > 8 || !g->entry) return 1;
    if (enable_peers(g)) return 1;
    if (reltol <= 0) reltol = std::sqrt(2.220446049250313e-16);
    if (maxiter <= 0) maxiter = (int)std::min<int64_t>(N, 2147483647);
    std::vector<cudaEvent_t> ev_written(S), ev_consumed(S);
    for (int s = 0; s < S; s++) {
        Shard& sh = g->shards[s];
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        for (int q = 0; q < 5; q++)
            if (int rc = sh.cg[q].ensure((size_t)N * 8)) return rc;
        if (int rc = sh.cg[5].ensure(64)) return rc;
        CF_CUDA(cudaEventCreateWithFlags(&ev_written[s], cudaEventDisableTiming));
        CF_CUDA(cudaEventCreateWithFlags(&ev_consumed[s], cudaEventDisableTiming));
        CF_CUDA(cudaMemcpyAsync(sh.cg[0].p, x, N * 8, cudaMemcpyHostToDevice, sh.stream));
        CF_CUDA(cudaMemcpyAsync(sh.cg[4].p, b, N * 8, cudaMemcpyHostToDevice, sh.stream));
        CF_CUDA(cudaMemsetAsync(sh.cg[2].p, 0, N * 8, sh.stream));
    }
    auto vec = [&](int s, int q) { return (double*)g->shards[s].cg[q].p; };
    const int vb = (int)std::min<int64_t>((N + 255) / 256, 4096);
    g->last_launches = 0;
    int rc_all = CF_OK;

    // symmetric Gramian: every device evaluates the unordered pairs of its row tiles (cyclic) into a partial vector over all n
    // elements; then every device sums the S partial vectors with peer loads in device order (an all-reduce over NVLink without a
    // staging copy) and adds sigma2 * in.  Bit-identical on every device and from run to run.
    const bool use_sym = !gradient && sym_applicable(g, vec(0, 0));
    auto apply_sym = [&](int out, int in) -> int {
        for (int s = 0; s < S; s++) {
            Shard& sh = g->shards[s];
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            if (int rc = sh.ypad.ensure((size_t)N * 8)) return rc;
            for (int t = 0; t < S; t++)
                if (t != s) CF_CUDA(cudaStreamWaitEvent(sh.stream, ev_consumed[t], 0)); // peers are done reading the old partial
            int rc = launch_mvm_sym(g, sh, (double*)sh.ypad.p, nullptr, vec(s, in), 1.0, 0.0, sh.stream, nullptr, s, S);
            if (rc) return rc == 1 ? fail(CF_ERR_INTERNAL, "symmetric variant unavailable inside the multi-device solve") : rc;
            CF_CUDA(cudaEventRecord(ev_written[s], sh.stream));
        }
        cf_parts_in parts;
        std::memset(&parts, 0, sizeof(parts));
        parts.n = S;
        for (int s = 0; s < S; s++) parts.ptr[s] = (const double*)g->shards[s].ypad.p;
        for (int s = 0; s < S; s++) {
            Shard& sh = g->shards[s];
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            for (int t = 0; t < S; t++)
                if (t != s) CF_CUDA(cudaStreamWaitEvent(sh.stream, ev_written[t], 0));
            cf_sum_parts_kernel<<<vb, 256, 0, sh.stream>>>(parts, 0, N, vec(s, out), vec(s, in), sigma2);
            CF_CUDA(cudaGetLastError());
            CF_CUDA(cudaEventRecord(ev_consumed[s], sh.stream));
        }
        return CF_OK;
    };
    // out = sigma2 * in + K in on every device (out, in: indices into cg[])
    auto apply = [&](int out, int in) -> int {
        if (use_sym) return apply_sym(out, in);
        for (int s = 0; s < S; s++) {
            Shard& sh = g->shards[s];
            const int64_t off = sh.r0 * blk, cnt = (sh.r1 - sh.r0) * blk;
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            for (int t = 0; t < S; t++)
                if (t != s) CF_CUDA(cudaStreamWaitEvent(sh.stream, ev_consumed[t], 0)); // peers are done reading the old `out`
            if (cnt == 0) { CF_CUDA(cudaEventRecord(ev_written[s], sh.stream)); continue; }
            const int sb = (int)std::min<int64_t>((cnt + 255) / 256, 4096);
            cf_axpby_kernel<<<sb, 256, 0, sh.stream>>>(vec(s, out) + off, sigma2, vec(s, in) + off, 0.0, vec(s, in) + off, cnt);
            CF_CUDA(cudaGetLastError());
            cf_peer_out peers;
            std::memset(&peers, 0, sizeof(peers));
            for (int t = 0; t < S; t++)
                if (t != s) peers.ptr[peers.n++] = vec(t, out) + off;
            int rc = gradient ? launch_grad(g, sh, vec(s, out) + off, vec(s, out) + off, vec(s, in), 1.0, 1.0, sh.stream, vg, &peers)
                              : launch_mvm(g, sh, vec(s, out) + off, vec(s, out) + off, vec(s, in), 1.0, 1.0, sh.stream, &peers);
            if (rc) return rc;
            CF_CUDA(cudaEventRecord(ev_written[s], sh.stream));
        }
        for (int s = 0; s < S; s++) {
            CF_CUDA(cudaSetDevice(g->shards[s].ctx->dev));
            for (int t = 0; t < S; t++)
                if (t != s) CF_CUDA(cudaStreamWaitEvent(g->shards[s].stream, ev_written[t], 0)); // all blocks have landed here
        }
        return CF_OK;
    };
    // z = a x + b y on every device
    auto axpby_all = [&](int z, double a_, int xq, double b_, int yq) -> int {
        for (int s = 0; s < S; s++) {
            CF_CUDA(cudaSetDevice(g->shards[s].ctx->dev));
            cf_axpby_kernel<<<vb, 256, 0, g->shards[s].stream>>>(vec(s, z), a_, vec(s, xq), b_, vec(s, yq), N);
            CF_CUDA(cudaGetLastError());
        }
        return CF_OK;
    };
    // dot product on every device (same value everywhere); the host reads device 0's copy
    auto dot_all = [&](int xq, int yq, double* host, bool consumed_after) -> int {
        for (int s = 0; s < S; s++) {
            Shard& sh = g->shards[s];
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            cf_dot_kernel<<<1, 1024, 0, sh.stream>>>(vec(s, xq), vec(s, yq), N, (double*)sh.cg[5].p);
            CF_CUDA(cudaGetLastError());
            if (consumed_after) CF_CUDA(cudaEventRecord(ev_consumed[s], sh.stream));
        }
        Shard& s0 = g->shards[0];
        CF_CUDA(cudaSetDevice(s0.ctx->dev));
        CF_CUDA(cudaMemcpyAsync(host, s0.cg[5].p, 8, cudaMemcpyDeviceToHost, s0.stream));
        CF_CUDA(cudaStreamSynchronize(s0.stream));
        return CF_OK;
    };
    // cg[]: 0 = x, 1 = r, 2 = u, 3 = c, 4 = b
    double rr = 0, residual = 0, prev_residual = 1.0;
    int it = 0;
    do {
        if ((rc_all = apply(3, 0))) break;                         // c = A x
        if ((rc_all = axpby_all(1, 1.0, 4, -1.0, 3))) break;       // r = b - c
        if ((rc_all = dot_all(1, 1, &rr, true))) break;            // c consumed
        residual = std::sqrt(rr);
        const double tol = reltol * residual;
        while (residual > tol && it < maxiter) {
            const double beta = (residual * residual) / (prev_residual * prev_residual);
            if ((rc_all = axpby_all(2, 1.0, 1, beta, 2))) break;   // u = r + beta u
            if ((rc_all = apply(3, 2))) break;                     // c = A u
            double uc = 0;
            if ((rc_all = dot_all(2, 3, &uc, false))) break;
            const double alpha = (residual * residual) / uc;
            if ((rc_all = axpby_all(0, 1.0, 0, alpha, 2))) break;  // x += alpha u
            if ((rc_all = axpby_all(1, 1.0, 1, -alpha, 3))) break; // r -= alpha c
            prev_residual = residual;
            if ((rc_all = dot_all(1, 1, &rr, true))) break;        // c consumed
            residual = std::sqrt(rr);
            it++;
        }
    } while (false);
    if (!rc_all) {
        Shard& s0 = g->shards[0];
        cudaSetDevice(s0.ctx->dev);
        if (cudaMemcpyAsync(x, s0.cg[0].p, N * 8, cudaMemcpyDeviceToHost, s0.stream) != cudaSuccess ||
            cudaStreamSynchronize(s0.stream) != cudaSuccess)
            rc_all = fail(CF_ERR_CUDA, "cg_solve: result copy failed");
    }
    for (int s = 0; s < S; s++) {
        cudaSetDevice(g->shards[s].ctx->dev);
        cudaStreamSynchronize(g->shards[s].stream);
        cudaEventDestroy(ev_written[s]);
        cudaEventDestroy(ev_consumed[s]);
    }
    if (rc_all) return rc_all;
    if (iters) *iters = it;
    if (resnorm) *resnorm = residual;
    return CF_OK;
}

// Multi-process CG (one process per GPU, cf_comm_init): every rank holds the full iterates and updates them redundantly -- the
// scalars come from identical vectors through the same deterministic reduction, so all ranks take bit-identical decisions and no
// all-reduce is needed -- each rank computes its row block of (sigma2 I + K) u in place and ONE NCCL all-gather per product
// re-assembles the vector.  The handle must be restricted to this rank's block of the standard split.
int cg_solve_comm(cf_gramian_s* g, double sigma2, double* x, const double* b, double reltol, int maxiter, int deriv, int* iters,
                  double* resnorm) {
    const bool gradient = deriv != 0;
    const int vg = deriv == 2 ? 1 : 0;
    const int64_t blk = deriv == 0 ? 1 : g->d + vg;
    const int64_t N = g->n * blk;
    cfcomm::Comm& cm = cfcomm::comm();
    Shard& sh = g->shards[0];
    if (sh.ctx->dev != cm.dev)
        return fail(CF_ERR_NCCL, "cf_cg_solve: the handle lives on device %d, the communicator on device %d", sh.ctx->dev, cm.dev);
    if (reltol <= 0) reltol = std::sqrt(2.220446049250313e-16);
    if (maxiter <= 0) maxiter = (int)std::min<int64_t>(N, 2147483647);
    CF_CUDA(cudaSetDevice(sh.ctx->dev));
    for (int q = 0; q < 5; q++)
        if (int rc = sh.cg[q].ensure((size_t)N * 8)) return rc;
    if (int rc = sh.cg[5].ensure(64)) return rc;
    double *dx = (double*)sh.cg[0].p, *dr = (double*)sh.cg[1].p, *du = (double*)sh.cg[2].p, *dc = (double*)sh.cg[3].p,
           *db = (double*)sh.cg[4].p, *dscal = (double*)sh.cg[5].p;
    cudaStream_t st = sh.stream;
    const int vb = (int)std::min<int64_t>((N + 255) / 256, 4096);
    const int64_t off = sh.r0 * blk, cnt = (sh.r1 - sh.r0) * blk;
    CF_CUDA(cudaMemcpyAsync(dx, x, N * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemcpyAsync(db, b, N * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemsetAsync(du, 0, N * 8, st));
    cudaEvent_t ea, eb, ec;
    CF_CUDA(cudaEventCreate(&ea)); CF_CUDA(cudaEventCreate(&eb)); CF_CUDA(cudaEventCreate(&ec));
    g->last_launches = 0;
    g->cg_mvm_ms = g->cg_gather_ms = 0; g->cg_products = 0;
    bool pending = false;  // events of the last product not yet read

    auto collect = [&]() {  // after a stream synchronisation: add the last product's device times
        if (!pending) return;
        float m = 0, ga = 0;
        if (cudaEventElapsedTime(&m, ea, eb) == cudaSuccess) g->cg_mvm_ms += m;
        if (cudaEventElapsedTime(&ga, eb, ec) == cudaSuccess) g->cg_gather_ms += ga;
        pending = false;
    };
    // symmetric Gramian: every rank evaluates the unordered pairs of its row tiles (cyclic) into a partial vector over all n elements
    // (rank 0's partial also carries sigma2 v), then ONE ncclAllReduce(sum) leaves the full product on every rank
    const bool use_sym = !gradient && sym_applicable(g, dx);
    // out = (sigma2 I + K) v on every rank: own row block in place, then the all-gather
    auto apply = [&](double* out, const double* v) -> int {
        CF_CUDA(cudaEventRecord(ea, st));
        if (use_sym) {
            if (cm.rank == 0) {
                cf_axpby_kernel<<<vb, 256, 0, st>>>(out, sigma2, v, 0.0, v, N);
                CF_CUDA(cudaGetLastError());
            }
            int rc = launch_mvm_sym(g, sh, out, out, v, 1.0, 1.0, st, nullptr, cm.rank, cm.world);  // beta y is added by part 0 only
            if (rc) return rc == 1 ? fail(CF_ERR_INTERNAL, "symmetric variant unavailable inside the multi-process solve") : rc;
            CF_CUDA(cudaEventRecord(eb, st));
            const int nrc = cfcomm::allreduce_sum_f64(out, N, st);
            if (nrc) return fail(CF_ERR_NCCL, "cf_cg_solve: NCCL all-reduce failed: %s", cfcomm::api().GetErrorString(nrc));
            CF_CUDA(cudaEventRecord(ec, st));
            pending = true;
            g->cg_products++;
            return CF_OK;
        }
        if (cnt > 0) {
            const int sb = (int)std::min<int64_t>((cnt + 255) / 256, 4096);
            cf_axpby_kernel<<<sb, 256, 0, st>>>(out + off, sigma2, v + off, 0.0, v + off, cnt);
            CF_CUDA(cudaGetLastError());
            int rc = gradient ? launch_grad(g, sh, out + off, out + off, v, 1.0, 1.0, st, vg) : launch_mvm(g, sh, out + off, out + off, v, 1.0, 1.0, st);
            if (rc) return rc;
        }
        CF_CUDA(cudaEventRecord(eb, st));
        const int nrc = cfcomm::allgather_rows(out, g->n, blk, 8, st);
        if (nrc) return fail(CF_ERR_NCCL, "cf_cg_solve: NCCL all-gather failed: %s", cfcomm::api().GetErrorString(nrc));
        CF_CUDA(cudaEventRecord(ec, st));
        pending = true;
        g->cg_products++;
        return CF_OK;
    };
    auto dot = [&](const double* u, const double* v, double* host) -> int {
        cf_dot_kernel<<<1, 1024, 0, st>>>(u, v, N, dscal);
        CF_CUDA(cudaGetLastError());
        CF_CUDA(cudaMemcpyAsync(host, dscal, 8, cudaMemcpyDeviceToHost, st));
        CF_CUDA(cudaStreamSynchronize(st));
        collect();
        return CF_OK;
    };
    int rc_all = CF_OK, it = 0;
    double rr = 0, residual = 0, prev_residual = 1.0;
    do {
        if ((rc_all = apply(dc, dx))) break;                       // r = b - A x
        cf_axpby_kernel<<<vb, 256, 0, st>>>(dr, 1.0, db, -1.0, dc, N);
        if ((rc_all = dot(dr, dr, &rr))) break;
        residual = std::sqrt(rr);
        const double tol = reltol * residual;
        while (residual > tol && it < maxiter) {
            const double beta = (residual * residual) / (prev_residual * prev_residual);
            cf_axpby_kernel<<<vb, 256, 0, st>>>(du, 1.0, dr, beta, du, N);   // u = r + beta u
            if ((rc_all = apply(dc, du))) break;                             // c = A u
            double uc = 0;
            if ((rc_all = dot(du, dc, &uc))) break;
            const double alpha = (residual * residual) / uc;
            cf_axpby_kernel<<<vb, 256, 0, st>>>(dx, 1.0, dx, alpha, du, N);   // x += alpha u
            cf_axpby_kernel<<<vb, 256, 0, st>>>(dr, 1.0, dr, -alpha, dc, N);  // r -= alpha c
            prev_residual = residual;
            if ((rc_all = dot(dr, dr, &rr))) break;
            residual = std::sqrt(rr);
            it++;
        }
    } while (false);
    if (!rc_all) {
        if (cudaMemcpyAsync(x, dx, N * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            rc_all = fail(CF_ERR_CUDA, "cg_solve: result copy failed");
    }
    cudaStreamSynchronize(st);
    cudaEventDestroy(ea); cudaEventDestroy(eb); cudaEventDestroy(ec);
    if (rc_all) return rc_all;
    if (iters) *iters = it;
    if (resnorm) *resnorm = residual;
    return CF_OK;
}

// conjugate gradients on (sigma2 I + K) x = b; restates IterativeSolvers.cg! 0.9.2 [upstream] behind
// ldiv!(x, ::LazyMatrixSum, b) (reference src/lazy_linear_algebra.jl:126-144).  State lives on shard 0; with several shards
// the search direction is broadcast to every device and the row blocks of K u are gathered back once per iteration
// (peer copies over NVLink).
int cg_solve_impl(cf_gramian_s* g, double sigma2, double* x, const double* b, double reltol, int maxiter, int deriv,
                  int* iters, double* resnorm) {
    const bool gradient = deriv != 0;
    const int vg = deriv == 2 ? 1 : 0;
    const int64_t blk = deriv == 0 ? 1 : g->d + vg;
    const int64_t N = g->n * blk;
    if (reltol <= 0) reltol = std::sqrt(2.220446049250313e-16);
    if (maxiter <= 0) maxiter = (int)std::min<int64_t>(N, 2147483647);
    if (N == 0) { if (iters) *iters = 0; if (resnorm) *resnorm = 0; return CF_OK; }
    g->cg_mvm_ms = g->cg_gather_ms = 0; g->cg_products = 0;
    if (cfcomm::active() && g->shards.size() == 1 && (g->row_begin != 0 || g->row_end != g->n))
        return cg_solve_comm(g, sigma2, x, b, reltol, maxiter, deriv, iters, resnorm);
    if (g->shards.size() > 1) {
        int rc = cg_solve_spmd(g, sigma2, x, b, reltol, maxiter, deriv, iters, resnorm);
        if (rc != 1) return rc; // 1: no peer access -> gather through device 0 below
    }
    Shard& s0 = g->shards[0];
    CF_CUDA(cudaSetDevice(s0.ctx->dev));
    for (int q = 0; q < 5; q++)
        if (int rc = s0.cg[q].ensure((size_t)N * 8)) return rc;
    if (int rc = s0.cg[5].ensure(64)) return rc;
    double *dx = (double*)s0.cg[0].p, *dr = (double*)s0.cg[1].p, *du = (double*)s0.cg[2].p, *dc = (double*)s0.cg[3].p,
           *db = (double*)s0.cg[4].p, *dscal = (double*)s0.cg[5].p;
    cudaStream_t st = s0.stream;
    const int vb = (int)std::min<int64_t>((N + 255) / 256, 4096);
    CF_CUDA(cudaMemcpyAsync(dx, x, N * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemcpyAsync(db, b, N * 8, cudaMemcpyHostToDevice, st));
    CF_CUDA(cudaMemsetAsync(du, 0, N * 8, st));
    g->last_launches = 0;

    // c = sigma2 * v + K v   (LazyMatrixSum mul!: y = 0; y += D v; y += G v)
    bool pending = false;
    auto apply = [&](double* out, const double* v) -> int {
        CF_CUDA(cudaEventRecord(s0.ev0, st));
        cf_axpby_kernel<<<vb, 256, 0, st>>>(out, sigma2, v, 0.0, v, N);
        CF_CUDA(cudaGetLastError());
        g->cg_products++;
        if (g->shards.size() == 1) {
            int rc = gradient ? launch_grad(g, s0, out, out, v, 1.0, 1.0, st, vg) : launch_mvm(g, s0, out, out, v, 1.0, 1.0, st);
            CF_CUDA(cudaEventRecord(s0.ev1, st));
            pending = true;
            return rc;
        }
        CF_CUDA(cudaStreamSynchronize(st));
        for (size_t q = 0; q < g->shards.size(); q++) {
            Shard& sh = g->shards[q];
            const int64_t srows = (sh.r1 - sh.r0) * blk;
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            if (int rc = sh.a.ensure((size_t)N * 8)) return rc;
            if (int rc = sh.y.ensure(std::max<size_t>(16, (size_t)srows * 8))) return rc;
            CF_CUDA(cudaMemcpyPeerAsync(sh.a.p, sh.ctx->dev, v, s0.ctx->dev, N * 8, sh.stream));
            if (srows == 0) continue;
            int rc = gradient ? launch_grad(g, sh, (double*)sh.y.p, nullptr, (const double*)sh.a.p, 1.0, 0.0, sh.stream, vg)
                              : launch_mvm(g, sh, sh.y.p, nullptr, sh.a.p, 1.0, 0.0, sh.stream);
            if (rc) return rc;
            // gather this row block next to the sigma2 term (scratch: cg[4] is b, so use shard-0 partial-free buffer)
        }
        // gather: copy every block into a scratch vector on device 0 and add
        CF_CUDA(cudaSetDevice(s0.ctx->dev));
        if (int rc = s0.ypad.ensure((size_t)N * 8)) return rc;
        for (size_t q = 0; q < g->shards.size(); q++) {
            Shard& sh = g->shards[q];
            const int64_t srows = (sh.r1 - sh.r0) * blk;
            if (srows == 0) continue;
            CF_CUDA(cudaSetDevice(sh.ctx->dev));
            CF_CUDA(cudaMemcpyPeerAsync((double*)s0.ypad.p + sh.r0 * blk, s0.ctx->dev, sh.y.p, sh.ctx->dev, srows * 8, sh.stream));
            CF_CUDA(cudaStreamSynchronize(sh.stream));
        }
        CF_CUDA(cudaSetDevice(s0.ctx->dev));
        cf_axpby_kernel<<<vb, 256, 0, st>>>(out, 1.0, out, 1.0, (const double*)s0.ypad.p, N);
        CF_CUDA(cudaGetLastError());
        return CF_OK;
    };
    auto dot = [&](const double* u, const double* v, double* host) -> int {
        cf_dot_kernel<<<1, 1024, 0, st>>>(u, v, N, dscal);
        CF_CUDA(cudaGetLastError());
        CF_CUDA(cudaMemcpyAsync(host, dscal, 8, cudaMemcpyDeviceToHost, st));
        CF_CUDA(cudaStreamSynchronize(st));
        if (pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s0.ev0, s0.ev1) == cudaSuccess) g->cg_mvm_ms += ms;
            pending = false;
        }
        return CF_OK;
    };

    if (int rc = apply(dc, dx)) return rc;                       // r = b - A x
    cf_axpby_kernel<<<vb, 256, 0, st>>>(dr, 1.0, db, -1.0, dc, N);
    CF_CUDA(cudaGetLastError());
    double rr = 0;
    if (int rc = dot(dr, dr, &rr)) return rc;
    double residual = std::sqrt(rr), prev_residual = 1.0;
    const double tol = reltol * residual;
    int it = 0;
    while (residual > tol && it < maxiter) {
        const double beta = (residual * residual) / (prev_residual * prev_residual);
        cf_axpby_kernel<<<vb, 256, 0, st>>>(du, 1.0, dr, beta, du, N);   // u = r + beta u
        CF_CUDA(cudaGetLastError());
        if (int rc = apply(dc, du)) return rc;                           // c = A u
        double uc = 0;
        if (int rc = dot(du, dc, &uc)) return rc;
        const double alpha = (residual * residual) / uc;
        cf_axpby_kernel<<<vb, 256, 0, st>>>(dx, 1.0, dx, alpha, du, N);   // x += alpha u
        cf_axpby_kernel<<<vb, 256, 0, st>>>(dr, 1.0, dr, -alpha, dc, N);  // r -= alpha c
        CF_CUDA(cudaGetLastError());
        prev_residual = residual;
        if (int rc = dot(dr, dr, &rr)) return rc;
        residual = std::sqrt(rr);
        it++;
    }
    CF_CUDA(cudaMemcpyAsync(x, dx, N * 8, cudaMemcpyDeviceToHost, st));
    CF_CUDA(cudaStreamSynchronize(st));
    if (iters) *iters = it;
    if (resnorm) *resnorm = residual;
    return CF_OK;
}

int peak_probe_impl(int kind, int iters, double* lane_ops_per_s, float* ms_out) {
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CF_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    void* out = nullptr;
    CF_CUDA(cudaMalloc(&out, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CF_CUDA(cudaEventCreate(&e0));
    CF_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CF_CUDA(cudaEventRecord(e0, 0));
        if (kind == 0) cf_peak_dfma_kernel<<<blocks, threads>>>((double*)out, iters, 0.999999, 1e-7);
        else if (kind == 1) cf_peak_ffma_kernel<<<blocks, threads>>>((float*)out, iters, 0.999999f, 1e-7f);
        else cf_peak_mufu_kernel<<<blocks, threads>>>((float*)out, iters, 0.5f);
        CF_CUDA(cudaGetLastError());
        CF_CUDA(cudaEventRecord(e1, 0));
        CF_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (ms_out) *ms_out = best;
    iters = ((iters + 7) / 8) * 8;
    if (lane_ops_per_s) *lane_ops_per_s = (double)blocks * threads * (double)iters * 16.0 / (best * 1e-3);
    return CF_OK;
}

// ---- d > 32 (bigd.cuh): row blocks sized so that the two [rows][m] scratch matrices stay <= 1 GiB each ------------------
int64_t bigd_row_block(int64_t nrows, int64_t m) {
    int64_t rb = ((int64_t(1) << 27) / std::max<int64_t>(m, 1) / CF_BD_T) * CF_BD_T;
    return std::max<int64_t>(CF_BD_T, std::min<int64_t>(rb, ((nrows + CF_BD_T - 1) / CF_BD_T) * CF_BD_T));
}

int launch_bigd_mvm(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                    cudaStream_t stream) {
    const int64_t nrows = sh.r1 - sh.r0, m = g->m;
    const int64_t rb = bigd_row_block(nrows, m);
    if (int rc = sh.bd_t.ensure((size_t)rb * m * 8)) return rc;
    if (int rc = sh.bd_s.ensure((size_t)rb * m * 8)) return rc;
    for (int64_t b0 = 0; b0 < nrows; b0 += rb) {
        const int64_t nb = std::min(rb, nrows - b0);
        const dim3 grid((unsigned)((m + CF_BD_T - 1) / CF_BD_T), (unsigned)((nb + CF_BD_T - 1) / CF_BD_T));
        bigd_pair_kernel<CF_GRAD_ISO, false><<<grid, 256, 0, stream>>>((const double*)sh.X, (const double*)sh.Y, nullptr, g->D,
                                                                       sh.r0 + b0, nb, m, (double*)sh.bd_t.p, (double*)sh.bd_s.p);
        CF_CUDA(cudaGetLastError());
        const int blocks = (int)std::min<int64_t>(nb, 148 * 8);
        bigd_value_kernel<<<blocks, 256, CF_EXP_TBL_DOUBLES * 8, stream>>>((const double*)sh.bd_t.p, (const double*)sh.bd_s.p, d_a, nb, m,
                                                                           g->sop_val, sh.ctx->exp2_tbl, d_y + b0,
                                                                           d_yin ? d_yin + b0 : nullptr, alpha, beta);
        CF_CUDA(cudaGetLastError());
        g->last_launches += 2;
    }
    return CF_OK;
}

int launch_bigd_grad(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                     cudaStream_t stream, int vg) {
    const int64_t nrows = sh.r1 - sh.r0, m = g->m;
    const int d = g->d, D = g->D, bs = d + vg;
    const int64_t rb = bigd_row_block(nrows, m);
    if (int rc = sh.bd_t.ensure((size_t)rb * m * 8)) return rc;
    if (int rc = sh.bd_s.ensure((size_t)rb * m * 8)) return rc;
    if (int rc = sh.apad.ensure(((size_t)m * D + (size_t)m) * 8)) return rc;  // padded gradient weights, then the value weights (ValueGradient)
    double* a0 = (double*)sh.apad.p + (size_t)m * D;
    cf_pad_points<double><<<(int)std::min<int64_t>((m * D + 255) / 256, 8192), 256, 0, stream>>>(d_a + vg, bs, d, (double*)sh.apad.p, D, m);
    if (vg) cf_pad_points<double><<<(int)std::min<int64_t>((m + 255) / 256, 8192), 256, 0, stream>>>(d_a, bs, 1, a0, 1, m);
    CF_CUDA(cudaGetLastError());
    const bool dot = g->prog.dotproduct != 0;
    for (int64_t b0 = 0; b0 < nrows; b0 += rb) {
        const int64_t nb = std::min(rb, nrows - b0);
        const dim3 grid((unsigned)((m + CF_BD_T - 1) / CF_BD_T), (unsigned)((nb + CF_BD_T - 1) / CF_BD_T));
        const dim3 ugrid((unsigned)((D + CF_BD_T - 1) / CF_BD_T), (unsigned)((nb + CF_BD_T - 1) / CF_BD_T));
        const int jb = (int)std::min<int64_t>((nb * m + 255) / 256, 148 * 16);
        const int vb = (int)std::min<int64_t>(nb, 148 * 8);
        double* T = (double*)sh.bd_t.p;
        double* S = (double*)sh.bd_s.p;
        const double* X = (const double*)sh.X;
        const double* Y = (const double*)sh.Y;
        const double* A = (const double*)sh.apad.p;
        double* out = d_y + b0 * bs;                             // entry 0 of row block b0 (the value entry when vg)
        const double* yin = d_yin ? d_yin + b0 * bs : nullptr;
        if (dot) {
            bigd_pair_kernel<CF_GRAD_DOT, true><<<grid, 256, 0, stream>>>(X, Y, A, D, sh.r0 + b0, nb, m, T, S);
            if (vg) bigd_jet_vg_kernel<CF_GRAD_DOT><<<vb, 256, CF_EXP_TBL_DOUBLES * 8, stream>>>(T, S, a0, nb, m, g->sop_grad, sh.ctx->exp2_tbl, out, yin, bs, alpha, beta);
            else bigd_jet_kernel<CF_GRAD_DOT><<<jb, 256, CF_EXP_TBL_DOUBLES * 8, stream>>>(T, S, nb * m, g->sop_grad, sh.ctx->exp2_tbl);
            bigd_update_kernel<CF_GRAD_DOT><<<ugrid, 256, 0, stream>>>(X, Y, A, D, d, sh.r0 + b0, nb, m, T, S, out + vg, yin ? yin + vg : nullptr, alpha, beta, bs);
        } else {
            bigd_pair_kernel<CF_GRAD_ISO, true><<<grid, 256, 0, stream>>>(X, Y, A, D, sh.r0 + b0, nb, m, T, S);
            if (vg) bigd_jet_vg_kernel<CF_GRAD_ISO><<<vb, 256, CF_EXP_TBL_DOUBLES * 8, stream>>>(T, S, a0, nb, m, g->sop_grad, sh.ctx->exp2_tbl, out, yin, bs, alpha, beta);
            else bigd_jet_kernel<CF_GRAD_ISO><<<jb, 256, CF_EXP_TBL_DOUBLES * 8, stream>>>(T, S, nb * m, g->sop_grad, sh.ctx->exp2_tbl);
            bigd_update_kernel<CF_GRAD_ISO><<<ugrid, 256, 0, stream>>>(X, Y, A, D, d, sh.r0 + b0, nb, m, T, S, out + vg, yin ? yin + vg : nullptr, alpha, beta, bs);
        }
        CF_CUDA(cudaGetLastError());
        g->last_launches += 3;
    }
    return CF_OK;
}

// symmetric variant (gram_mvm_sym.cuh): diagonal row blocks with the plain kernel, everything beyond them once; deterministic
// (single-writer partial sums, fixed-order combine).  Returns 1 when the variant does not apply (caller runs the plain path).
// part / parts: this device's share of the row tiles (cyclic); with parts > 1 d_y receives the device's PARTIAL vector over all n
// elements (see gram_sym_combine) and the caller sums the partial vectors.
int launch_mvm_sym(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                   cudaStream_t stream, const cf_peer_out* peers, int part = 0, int parts = 1) {
    cf_peer_out no_peers;
    std::memset(&no_peers, 0, sizeof(no_peers));
    if (!peers) peers = &no_peers;
    const int64_t n = g->n;
    const bool eqf = g->eq_fast && g->entry->sym_eq != nullptr && !env_flag("COVFN_MVM_SCALAR");
    const bool matf = g->mat_fast && g->entry->sym_mat != nullptr && g->entry->mvm_mat != nullptr && !env_flag("COVFN_MVM_SCALAR");
    const cf_mvm_config& cfg = eqf ? g->entry->mvm_eq_cfg : (matf ? g->entry->mvm_mat_cfg : g->entry->mvm_cfg[CF_F64]);
    const int64_t TR = cfg.rows_per_cta, TJ = cfg.tj;
    const int64_t T = (n + TR - 1) / TR;
    if (sh.sym_nitems < 0 || sh.sym_tr != TR || sh.sym_part != part || sh.sym_parts != parts) { // build the (row tile, column chunk) list once
        // column chunk: ~8192 work items PER DEVICE.  (With a chunk length independent of the device count a device of eight had ~1000
        // items for its 296 resident CTAs, 3.7 waves: the last, partly filled wave cost ~9 % at N = 8.)
        int64_t ch = ((T * n / 2 / (8192 * (int64_t)parts)) / TJ) * TJ;
        if (ch < 8 * TJ) ch = 8 * TJ;
        std::vector<cf_sym_item> items;
        int64_t colpart_elems = 0, maxchunks = 0;
        for (int64_t I = 0; I < T; I++) {
            const int64_t first = (I + 1) * TR;
            const int64_t off = I * n - TR * (I * (I + 1) / 2);  // triangular layout, see gram_sym_combine
            int64_t c = 0;
            for (int64_t c0 = first; c0 < n && I % parts == part; c0 += ch, c++) {
                cf_sym_item it;
                it.col0 = c0; it.col1 = std::min(n, c0 + ch); it.colpart_off = off + (c0 - first);
                it.row_tile = (int32_t)I; it.chunk = (int32_t)c;
                items.push_back(it);
            }
            maxchunks = std::max(maxchunks, c);
            if (first < n) colpart_elems = off + (n - first);
        }
        size_t free_b = 0, total_b = 0;
        CF_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if ((size_t)colpart_elems * 8 > total_b / 4) return 1;  // column partials would not fit comfortably: plain path
        if (int rc = sh.sym_items.ensure(std::max<size_t>(16, items.size() * sizeof(cf_sym_item)))) return rc;
        if (!items.empty())
            CF_CUDA(cudaMemcpyAsync(sh.sym_items.p, items.data(), items.size() * sizeof(cf_sym_item), cudaMemcpyHostToDevice, stream));
        CF_CUDA(cudaStreamSynchronize(stream)); // items is a host temporary
        if (int rc = sh.sym_col.ensure(std::max<size_t>(16, (size_t)colpart_elems * 8))) return rc;
        if (int rc = sh.bsym.ensure(std::max<size_t>(16, (size_t)maxchunks * n * 8))) return rc;
        sh.sym_nitems = (int)items.size();
        sh.sym_tr = TR; sh.sym_ch = ch; sh.sym_part = part; sh.sym_parts = parts;
    }
    if (int rc = sh.partial.ensure((size_t)n * 8)) return rc;
    // 1. diagonal blocks: plain kernel, each CTA sweeps only its own row block, raw sums into partial[0][*]
    cf_mvm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.X; P.Y = sh.X; P.a = d_a; P.xn = sh.xn; P.yn = sh.xn; P.out = sh.partial.p;
    P.exp2_tbl = sh.ctx->exp2_tbl; P.sop = g->sop_val;
    P.row0 = 0; P.nrows = n; P.m = n; P.cols_per_chunk = ((n + TJ - 1) / TJ) * TJ;
    P.alpha = 1.0; P.beta = 0.0; P.direct = 0; P.use_tma = 1; P.diag_block = TR;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    const long double lam = 0.693147180559945309417232121458176568L / 256.0L;
    P.eqc[0] = (double)lam; P.eqc[1] = (double)(lam * lam / 2); P.eqc[2] = (double)(lam * lam * lam / 6);
    P.eqc[3] = (double)(lam * lam * lam * lam / 24);
    CF_CUDA((eqf ? g->entry->mvm_eq : (matf ? g->entry->mvm_mat : g->entry->mvm[CF_F64][cf_kind_slot(g->kind)]))(P, dim3((unsigned)T, 1), stream));
    // 2. everything beyond the diagonal blocks, each unordered pair once
    if (sh.sym_nitems > 0) {
        cf_sym_params S;
        std::memset(&S, 0, sizeof(S));
        S.X = (const double*)sh.X; S.xn = (const double*)sh.xn; S.a = d_a;
        S.rowpart = (double*)sh.bsym.p; S.colpart = (double*)sh.sym_col.p; S.exp2_tbl = sh.ctx->exp2_tbl;
        S.items = (const cf_sym_item*)sh.sym_items.p; S.n = n; S.use_tma = 1;
        S.atom = P.atom; S.sop = g->sop_val;
        for (int q = 0; q < 4; q++) S.eqc[q] = P.eqc[q];
        CF_CUDA((eqf ? g->entry->sym_eq : (matf ? g->entry->sym_mat : g->entry->sym[cf_kind_slot(g->kind)]))(S, sh.sym_nitems, stream));
    }
    // 3. y = alpha (diag + row partials + column partials) + beta y, fixed summation order
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 8192);
    gram_sym_combine<<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, (const double*)sh.bsym.p, (const double*)sh.sym_col.p, n, TR,
                                                 sh.sym_ch, d_y, d_yin, alpha * g->coef, beta, *peers, part, parts);
    CF_CUDA(cudaGetLastError());
    g->last_launches += 3;
    return CF_OK;
}

// does the symmetric variant apply to a full-vector product of this handle (whatever row range it is restricted to)?
bool sym_applicable(cf_gramian_s* g, const void* d_a) {
    return g->opt_symmetric && g->symmetric && g->dtype == CF_F64 && g->entry && g->n >= 32768 && (((uintptr_t)d_a) % 16) == 0 &&
           !(g->kind == CF_ATOM_SOP && cfjit::wanted((double)g->n * (double)g->m));
}

// one column of  y <- alpha K a + beta y  on one shard; device pointers; asynchronous on sh.stream
int launch_mvm_dmma(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                    cudaStream_t stream, const cf_peer_out* peers) {
    const int64_t nrows = sh.r1 - sh.r0;
    const cf_mvm_config& cfg = g->entry->mvm_dmma_cfg;
    const int slot = cf_kind_slot(g->kind);
    if (int rc = ensure_padded_points(g, sh, stream)) return rc;
    Plan pl = make_plan(nrows, g->m, cfg, sh.ctx->sms);
    cf_mvm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.xp.p; P.Y = (sh.Y != sh.X) ? sh.yp.p : sh.xp.p; P.xn = sh.xn; P.yn = sh.yn; P.a = d_a;
    P.exp2_tbl = sh.ctx->exp2_tbl;
    P.sop = g->sop_val;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m;
    P.cols_per_chunk = pl.cols_per_chunk;
    P.alpha = alpha * g->coef; P.beta = beta;
    P.use_tma = (((uintptr_t)d_a) % 16 == 0) ? 1 : 0;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    P.direct = (pl.chunks == 1) ? 1 : 0;
    P.peers = *peers;
    if (P.direct) {
        P.out = d_y; P.yin = d_yin;
    } else {
        if (int rc = sh.partial.ensure((size_t)pl.chunks * nrows * sizeof(double))) return rc;
        P.out = sh.partial.p;
    }
    bool launched = false;
    if (g->kind == CF_ATOM_SOP && cfjit::wanted((double)nrows * (double)g->m)) {
        const std::string name = "gram_mvm_dmma_kernel<" + std::to_string(g->D) + ", " + std::to_string((int)CF_ATOM_SOP) + ">";
        if (cfjit::Kernel* jit = cfjit::get_kernel(g->sop_val, "gram_mvm_dmma.cuh", name))
            launched = cfjit::launch(jit, &P, (unsigned)pl.row_tiles, (unsigned)pl.chunks, 256, (unsigned)cfg.smem_bytes, stream) == 0;
    }
    if (!launched) CF_CUDA(g->entry->mvm_dmma[slot](P, dim3(pl.row_tiles, pl.chunks), stream));
    g->last_launches++;
    if (!P.direct) {
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, 4096);
        gram_reduce_partials<double><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, d_y, d_yin,
                                                                  alpha * g->coef, beta, *peers);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
    }
    return CF_OK;
}

// padded Float32 copy of the column points for the 3xTF32 kernels (B fragments), built once per shard
int ensure_padded_points_f32(cf_gramian_s* g, Shard& sh, cudaStream_t stream) {
    if (sh.mmd_ready) return CF_OK;
    const int sx = g->entry->mm_tf32_sx;
    if (int rc = sh.yp.ensure((size_t)g->m * sx * 4)) return rc;
    cf_pad_rows_f32_kernel<<<148 * 8, 256, 0, stream>>>((const float*)sh.Y, g->D, sx, g->m, (float*)sh.yp.p);
    CF_CUDA(cudaGetLastError());
    sh.mmd_ready = true;
    return CF_OK;
}

int launch_mvm_tf32(cf_gramian_s* g, Shard& sh, void* d_y, const void* d_yin, const void* d_a, double alpha, double beta,
                    cudaStream_t stream, const cf_peer_out* peers) {
    const int64_t nrows = sh.r1 - sh.r0;
    const cf_mvm_config& cfg = g->entry->mvm_tf32_cfg;
    const int slot = cf_kind_slot(g->kind);
    if (int rc = ensure_padded_points_f32(g, sh, stream)) return rc;
    Plan pl = make_plan(nrows, g->m, cfg, sh.ctx->sms);
    cf_mvm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.X; P.Y = sh.yp.p; P.xn = sh.xn; P.yn = sh.yn; P.a = d_a;
    P.sop = g->sop_val;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m;
    P.cols_per_chunk = pl.cols_per_chunk;
    P.alpha = alpha * g->coef; P.beta = beta;
    P.use_tma = (((uintptr_t)d_a) % 16 == 0) ? 1 : 0;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    P.direct = (pl.chunks == 1) ? 1 : 0;
    P.peers = *peers;
    if (P.direct) {
        P.out = d_y; P.yin = d_yin;
    } else {
        if (int rc = sh.partial.ensure((size_t)pl.chunks * nrows * sizeof(double))) return rc;
        P.out = sh.partial.p;
    }
    bool launched = false;
    if (g->kind == CF_ATOM_SOP && cfjit::wanted((double)nrows * (double)g->m)) {
        const std::string name = "gram_mvm_tf32_kernel<" + std::to_string(g->D) + ", " + std::to_string((int)CF_ATOM_SOP) + ">";
        if (cfjit::Kernel* jit = cfjit::get_kernel(g->sop_val, "gram_mvm_tf32.cuh", name))
            launched = cfjit::launch(jit, &P, (unsigned)pl.row_tiles, (unsigned)pl.chunks, 256, (unsigned)cfg.smem_bytes, stream) == 0;
    }
    if (!launched) CF_CUDA(g->entry->mvm_tf32[slot](P, dim3(pl.row_tiles, pl.chunks), stream));
    g->last_launches++;
    if (!P.direct) {
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, 4096);
        gram_reduce_partials<float><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, (float*)d_y, (const float*)d_yin,
                                                                 alpha * g->coef, beta, *peers);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
    }
    return CF_OK;
}

// Float32 value MVM, padded D >= 8, well-scaled points: dot products on tcgen05 / TMEM in 3xTF32 (gram_mvm_tc5.cuh)
int launch_mvm_tc5(cf_gramian_s* g, Shard& sh, void* d_y, const void* d_yin, const void* d_a, double alpha, double beta,
                   cudaStream_t stream, const cf_peer_out* peers) {
    const int64_t nrows = sh.r1 - sh.r0;
    const cf_mvm_config& cfg = g->entry->mvm_tc5_cfg;
    const int slot = cf_kind_slot(g->kind);
    if (int rc = ensure_canonical_points(g, sh, stream)) return rc;
    const int64_t mpad = ((g->m + CF_MVU_TJ - 1) / CF_MVU_TJ) * CF_MVU_TJ;
    const float* ap = (const float*)d_a;
    if (mpad != g->m || ((uintptr_t)d_a) % 16 != 0) {  // the weight tiles arrive by TMA: 16-byte aligned, zero beyond m
        if (int rc = sh.au.ensure((size_t)mpad * 4)) return rc;
        cf_pad_vec_f32_kernel<<<(int)std::min<int64_t>((mpad + 255) / 256, 2048), 256, 0, stream>>>((const float*)d_a, g->m, mpad, (float*)sh.au.p);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
        ap = (const float*)sh.au.p;
    }
    Plan pl = make_plan(nrows, g->m, cfg, sh.ctx->sms);
    cf_mvu_params PP;
    std::memset(&PP, 0, sizeof(PP));
    cf_mvm_params& P = PP.mv;
    P.X = sh.X; P.xn = sh.xn; P.sop = g->sop_val;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m;
    P.cols_per_chunk = pl.cols_per_chunk;
    P.alpha = alpha * g->coef; P.beta = beta;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    P.direct = (pl.chunks == 1) ? 1 : 0;
    P.peers = *peers;
    if (P.direct) {
        P.out = d_y; P.yin = d_yin;
    } else {
        if (int rc = sh.partial.ensure((size_t)pl.chunks * nrows * sizeof(double))) return rc;
        P.out = sh.partial.p;
    }
    PP.yhi = (const float*)sh.yc_hi.p; PP.ylo = (const float*)sh.yc_lo.p; PP.ynpad = (const float*)sh.yc_n.p; PP.apad = ap;
    bool launched = false;
    if (g->kind == CF_ATOM_SOP && cfjit::wanted((double)nrows * (double)g->m)) {
        const std::string name = "gram_mvm_tc5_kernel<" + std::to_string(g->D) + ", " + std::to_string((int)CF_ATOM_SOP) + ">";
        if (cfjit::Kernel* jit = cfjit::get_kernel(g->sop_val, "gram_mvm_tc5.cuh", name))
            launched = cfjit::launch(jit, &PP, (unsigned)pl.row_tiles, (unsigned)pl.chunks, CF_MVU_THREADS, (unsigned)cfg.smem_bytes, stream) == 0;
    }
    if (!launched) CF_CUDA(g->entry->mvm_tc5[slot](PP, dim3(pl.row_tiles, pl.chunks), stream));
    g->last_launches++;
    if (!P.direct) {
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, 4096);
        gram_reduce_partials<float><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, (float*)d_y, (const float*)d_yin,
                                                                 alpha * g->coef, beta, *peers);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
    }
    return CF_OK;
}

// Float32 value MVM of a single isotropic atom at small d in packed FP32 arithmetic (gram_mvm_f32p.cuh); the transposed copy of the
// column points it streams is built once per shard
int launch_mvm_f32p(cf_gramian_s* g, Shard& sh, void* d_y, const void* d_yin, const void* d_a, double alpha, double beta,
                    cudaStream_t stream, const cf_peer_out* peers) {
    const int64_t nrows = sh.r1 - sh.r0;
    const cf_mvm_config& cfg = g->entry->mvm_f32p_cfg;
    if (sh.yt_ld == 0) {
        const int64_t ldt = ((g->m + cfg.tj - 1) / cfg.tj) * cfg.tj;
        if (int rc = sh.yt.ensure((size_t)ldt * g->D * 4)) return rc;
        cf_transpose_points_f32_kernel<<<148 * 8, 256, 0, stream>>>((const float*)sh.Y, g->D, g->m, ldt, (float*)sh.yt.p);
        CF_CUDA(cudaGetLastError());
        sh.yt_ld = ldt;
    }
    Plan pl = make_plan(nrows, g->m, cfg, sh.ctx->sms);
    cf_mvm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.X; P.Y = sh.yt.p; P.diag_block = sh.yt_ld; P.a = d_a;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m;
    P.cols_per_chunk = pl.cols_per_chunk;
    P.alpha = alpha * g->coef; P.beta = beta;
    P.use_tma = (((uintptr_t)d_a) % 16 == 0) ? 1 : 0;
    P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    P.direct = (pl.chunks == 1) ? 1 : 0;
    P.peers = *peers;
    if (P.direct) {
        P.out = d_y; P.yin = d_yin;
    } else {
        if (int rc = sh.partial.ensure((size_t)pl.chunks * nrows * sizeof(double))) return rc;
        P.out = sh.partial.p;
    }
    CF_CUDA(g->entry->mvm_f32p[cf_kind_slot(g->kind)](P, dim3(pl.row_tiles, pl.chunks), stream));
    g->last_launches++;
    if (!P.direct) {
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, 4096);
        gram_reduce_partials<float><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, (float*)d_y, (const float*)d_yin,
                                                                 alpha * g->coef, beta, *peers);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
    }
    return CF_OK;
}

int launch_mvm(cf_gramian_s* g, Shard& sh, void* d_y, const void* d_yin, const void* d_a, double alpha, double beta,
               cudaStream_t stream, const cf_peer_out* peers) {
    cf_peer_out no_peers;
    std::memset(&no_peers, 0, sizeof(no_peers));
    if (!peers) peers = &no_peers;
    const int64_t nrows = sh.r1 - sh.r0;
    if (nrows <= 0) return CF_OK;
    const int dt = g->dtype;
    if (g->m == 0) { // empty sum: y = beta*y
        launch_scale(dt, d_y, d_yin, nrows, beta, stream);
        return CF_OK;
    }
    if (!g->entry) return launch_bigd_mvm(g, sh, (double*)d_y, (const double*)d_yin, (const double*)d_a, alpha, beta, stream);
    const cf_mvm_config& cfg = g->entry->mvm_cfg[dt];
    // y === x: every unordered pair once (gram_mvm_sym.cuh, deterministic); on by default, CF_OPT_SYMMETRIC / COVFN_SYMMETRIC=0 turn it off
    if (sh.r0 == 0 && sh.r1 == g->n && sym_applicable(g, d_a)) {
        const int rc = launch_mvm_sym(g, sh, (double*)d_y, (const double*)d_yin, (const double*)d_a, alpha, beta, stream, peers);
        if (rc != 1) return rc;
    }
    // high-dimensional, well-scaled Float64 points: pair distances on the FP64 tensor cores (gram_mvm_dmma.cuh);
    // COVFN_MVM_SCALAR=1 keeps the scalar kernel
    const int slot = cf_kind_slot(g->kind);
    const bool dmma = dt == CF_F64 && g->use_norms && g->entry->mvm_dmma[slot] != nullptr && !env_flag("COVFN_MVM_SCALAR");
    if (dmma) return launch_mvm_dmma(g, sh, (double*)d_y, (const double*)d_yin, (const double*)d_a, alpha, beta, stream, peers);
    // the Float32 counterpart: dot products on tcgen05 / TMEM in 3xTF32 (gram_mvm_tc5.cuh); COVFN_MVM_LEGACY=1 keeps the mma.sync kernel below
    if (dt == CF_F32 && g->use_norms && g->entry->mvm_tc5[slot] != nullptr && !env_flag("COVFN_MVM_SCALAR") && !env_flag("COVFN_MVM_LEGACY"))
        return launch_mvm_tc5(g, sh, d_y, d_yin, d_a, alpha, beta, stream, peers);
    // ... its predecessor: distance GEMM in 3xTF32 with mma.sync fragments (gram_mvm_tf32.cuh)
    if (dt == CF_F32 && g->use_norms && g->entry->mvm_tf32[slot] != nullptr && !env_flag("COVFN_MVM_SCALAR"))
        return launch_mvm_tf32(g, sh, d_y, d_yin, d_a, alpha, beta, stream, peers);
    // single isotropic atom in Float32 at small d (or ill-scaled points at d <= 8): packed FP32 arithmetic (gram_mvm_f32p.cuh)
    if (dt == CF_F32 && g->prog.single && slot < 3 && g->entry->mvm_f32p[slot] != nullptr && !env_flag("COVFN_MVM_SCALAR"))
        return launch_mvm_f32p(g, sh, d_y, d_yin, d_a, alpha, beta, stream, peers);
    // single EQ atom on well-scaled Float64 points, small d: exponent formed in the scaled domain (gram_mvm_eq.cuh)
    const bool eqf = g->eq_fast && g->entry->mvm_eq != nullptr && !env_flag("COVFN_MVM_SCALAR");
    // ... and its MaternP form (p >= 1): r2 from the norm expansion, 6-instruction square root, clamp-free exp
    const bool matf = g->mat_fast && g->entry->mvm_mat != nullptr && !env_flag("COVFN_MVM_SCALAR");
    Plan pl = make_plan(nrows, g->m, eqf ? g->entry->mvm_eq_cfg : (matf ? g->entry->mvm_mat_cfg : cfg), sh.ctx->sms);
    cf_mvm_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = sh.X; P.Y = sh.Y; P.a = d_a; P.xn = sh.xn; P.yn = sh.yn;
    P.exp2_tbl = sh.ctx->exp2_tbl;
    P.sop = g->sop_val;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m;
    P.cols_per_chunk = pl.cols_per_chunk;
    P.alpha = alpha * g->coef; P.beta = beta;
    P.use_tma = (((uintptr_t)d_a) % 16 == 0) ? 1 : 0;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom].v;
    if (eqf || matf) {
        const long double lam = 0.693147180559945309417232121458176568L / 256.0L;
        P.eqc[0] = (double)lam; P.eqc[1] = (double)(lam * lam / 2); P.eqc[2] = (double)(lam * lam * lam / 6);
        P.eqc[3] = (double)(lam * lam * lam * lam / 24);
    }
    P.direct = (pl.chunks == 1) ? 1 : 0;
    P.peers = *peers;
    if (P.direct) {
        P.out = d_y; P.yin = d_yin;
    } else {
        int rc = sh.partial.ensure((size_t)pl.chunks * nrows * sizeof(double));
        if (rc) return rc;
        P.out = sh.partial.p;
    }
    bool launched = false;
    if (dt == CF_F64 && g->kind == CF_ATOM_SOP && cfjit::wanted((double)nrows * (double)g->m)) {
        // composite program, large problem: the same kernel with the program structure compiled in (cf_jit.h)
        const int* tu = g->entry->tune;
        const std::string name = "gram_mvm_kernel<double, " + std::to_string(g->D) + ", " + std::to_string((int)CF_ATOM_SOP) + ", " +
                                 std::to_string(tu[0]) + ", " + std::to_string(tu[1]) + ", " + std::to_string(tu[2]) + ", " +
                                 std::to_string(tu[3]) + ", " + std::to_string(tu[4]) + ">";
        if (cfjit::Kernel* jit = cfjit::get_kernel(g->sop_val, "gram_mvm.cuh", name))
            launched = cfjit::launch(jit, &P, (unsigned)pl.row_tiles, (unsigned)pl.chunks, (unsigned)tu[1], (unsigned)cfg.smem_bytes, stream) == 0;
    }
    if (!launched) {
        cf_mvm_launch_fn fn = eqf ? g->entry->mvm_eq : (matf ? g->entry->mvm_mat : g->entry->mvm[dt][cf_kind_slot(g->kind)]);
        CF_CUDA(fn(P, dim3(pl.row_tiles, pl.chunks), stream));
    }
    g->last_launches++;
    if (!P.direct) {
        const int blocks = (int)std::min<int64_t>((nrows + 255) / 256, 4096);
        if (dt == CF_F64)
            gram_reduce_partials<double><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, (double*)d_y,
                                                                      (const double*)d_yin, alpha * g->coef, beta, *peers);
        else
            gram_reduce_partials<float><<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, pl.chunks, nrows, (float*)d_y,
                                                                     (const float*)d_yin, alpha * g->coef, beta, *peers);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
    }
    return CF_OK;
}

// derivative operators, one right-hand side, device pointers (unpadded flat vectors with blocks of d + vg entries):
// vg = 0 GradientKernel, vg = 1 ValueGradientKernel (entry 0 of every block is the value part)
int launch_grad(cf_gramian_s* g, Shard& sh, double* d_y, const double* d_yin, const double* d_a, double alpha, double beta,
                cudaStream_t stream, int vg, const cf_peer_out* peers) {
    cf_peer_out no_peers;
    std::memset(&no_peers, 0, sizeof(no_peers));
    if (!peers) peers = &no_peers;
    const int64_t nrows = sh.r1 - sh.r0;
    if (nrows <= 0) return CF_OK;
    const int d = g->d, D = g->D, bs = d + vg;
    if (g->m == 0) {
        launch_scale(CF_F64, d_y, d_yin, nrows * bs, beta, stream);
        return CF_OK;
    }
    if (!g->entry) return launch_bigd_grad(g, sh, d_y, d_yin, d_a, alpha, beta, stream, vg);
    {   // isotropic GradientKernel on well-scaled Float64 points: every d-dependent operation on the FP64 tensor cores
        // (grad_mvm_dmma.cuh); COVFN_GRAD_SCALAR=1 keeps the scalar kernel
        const bool eq = g->prog.single && g->prog.atoms[g->prog.terms[0].fac[0].atom].v.kind == CF_ATOM_EQ;
        const bool matern = g->prog.single && g->prog.atoms[g->prog.terms[0].fac[0].atom].v.kind == CF_ATOM_MATERN &&
                            g->prog.atoms[g->prog.terms[0].fac[0].atom].v.p >= 2;
        const bool dot = g->prog.dotproduct != 0;  // nothing cancels in the dot-product form: no scale check needed
        cf_gradd_launch_fn fn = g->entry->grad_dmma[vg][dot ? 3 : (eq ? 0 : (matern ? 2 : 1))];
        // (below ~2^22 blocks the extra preparation launches cost more than the tensor cores save)
        if ((dot || g->use_norms_grad) && fn &&
            ((double)nrows * (double)g->m >= 4194304.0 || env_flag("COVFN_GRAD_DMMA")) && !env_flag("COVFN_GRAD_SCALAR")) {
            if (int rc = ensure_padded_points(g, sh, stream)) return rc;
            const int sx = (D % 8 == 4) ? D + 8 : D + 4;  // cf_mmd_smem<D>::sx
            const cf_mvm_config& cfgd = g->entry->grad_dmma_cfg;
            Plan pl = make_plan(nrows, g->m, cfgd, sh.ctx->sms);
            // padded gradient weights, then q_j = y_j . a_j, then the value weights (ValueGradient): all 16-byte aligned TMA sources
            const size_t q_off = (((size_t)g->m * sx + 1) / 2) * 2;
            const size_t a0_off = q_off + (((size_t)g->m + 1) / 2) * 2;
            if (int rc = sh.ap.ensure((a0_off + (size_t)g->m + 2) * sizeof(double))) return rc;
            if (int rc = sh.partial.ensure((size_t)pl.chunks * nrows * (D + 1) * sizeof(double))) return rc;
            double* ap = (double*)sh.ap.p;
            const double* yp = (const double*)((sh.Y != sh.X) ? sh.yp.p : sh.xp.p);
            cf_pad_points<double><<<(int)std::min<int64_t>((g->m * sx + 255) / 256, 8192), 256, 0, stream>>>(d_a + vg, bs, d, ap, sx, g->m);
            cf_rowdot_kernel<<<(int)std::min<int64_t>((g->m + 255) / 256, 4096), 256, 0, stream>>>(yp, ap, sx, g->m, ap + q_off);
            if (vg) cf_pad_points<double><<<(int)std::min<int64_t>((g->m + 255) / 256, 8192), 256, 0, stream>>>(d_a, bs, 1, ap + a0_off, 1, g->m);
            CF_CUDA(cudaGetLastError());
            cf_gradd_params PP;
            std::memset(&PP, 0, sizeof(PP));
            cf_grad_params& P = PP.g;
            P.X = (const double*)sh.xp.p; P.Y = yp; P.a = ap; P.a0 = vg ? ap + a0_off : nullptr;
            P.partial = (double*)sh.partial.p;
            P.partial0 = (double*)sh.partial.p + (size_t)pl.chunks * nrows * D;
            P.exp2_tbl = sh.ctx->exp2_tbl;
            P.sop = g->sop_grad;
            P.row0 = sh.r0; P.nrows = nrows; P.m = g->m; P.cols_per_chunk = pl.cols_per_chunk;
            P.single = g->prog.single;
            P.coef = g->coef_grad;
            if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom];
            PP.xn = (const double*)sh.xn; PP.yn = (const double*)sh.yn; PP.q = ap + q_off;
            bool launched = false;
            if (!g->prog.single && cfjit::wanted((double)nrows * (double)g->m * (double)D)) {
                // composite derivative program: the same kernel with the jets' product rule generated for this structure (cf_jit.h)
                const std::string name = "grad_mvm_dmma_kernel<" + std::to_string(D) + ", " + std::to_string((int)CF_ATOM_SOP) + ", " +
                                         (vg ? "true" : "false") + ", " + std::to_string(dot ? CF_GRAD_DOT : CF_GRAD_ISO) + ">";
                if (cfjit::Kernel* jit = cfjit::get_kernel(g->sop_val, "grad_mvm_dmma.cuh", name, &g->sop_grad))
                    launched = cfjit::launch(jit, &PP, (unsigned)pl.row_tiles, (unsigned)pl.chunks, 256, (unsigned)cfgd.smem_bytes, stream) == 0;
            }
            if (!launched) CF_CUDA(fn(PP, dim3(pl.row_tiles, pl.chunks), stream));
            const int blocks = (int)std::min<int64_t>((nrows * bs + 255) / 256, 8192);
            grad_reduce_partials<<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, P.partial0, pl.chunks, nrows, D, d, vg, d_y, d_yin,
                                                             alpha, beta, *peers);
            CF_CUDA(cudaGetLastError());
            g->last_launches += 4 + vg;
            return CF_OK;
        }
    }
    const cf_mvm_config& cfg = g->entry->grad_cfg[vg];
    Plan pl = make_plan(nrows, g->m, cfg, sh.ctx->sms);
    const double* a_use = d_a;
    const double* a0_use = nullptr;
    if (vg || D != d || (((uintptr_t)d_a) % 16) != 0) {
        const size_t a0_off = (((size_t)g->m * D + 1) / 2) * 2; // the value weights follow the gradient weights, 16-byte aligned (TMA source)
        int rc = sh.apad.ensure((a0_off + (size_t)g->m + 2) * sizeof(double));
        if (rc) return rc;
        const int blocks = (int)std::min<int64_t>((g->m * D + 255) / 256, 8192);
        cf_pad_points<double><<<blocks, 256, 0, stream>>>(d_a + vg, bs, d, (double*)sh.apad.p, D, g->m);
        CF_CUDA(cudaGetLastError());
        g->last_launches++;
        a_use = (const double*)sh.apad.p;
        if (vg) {
            double* a0 = (double*)sh.apad.p + a0_off;
            cf_pad_points<double><<<(int)std::min<int64_t>((g->m + 255) / 256, 8192), 256, 0, stream>>>(d_a, bs, 1, a0, 1, g->m);
            CF_CUDA(cudaGetLastError());
            g->last_launches++;
            a0_use = a0;
        }
    }
    int rc = sh.partial.ensure((size_t)pl.chunks * nrows * (D + 1) * sizeof(double));
    if (rc) return rc;
    cf_grad_params P;
    std::memset(&P, 0, sizeof(P));
    P.X = (const double*)sh.X; P.Y = (const double*)sh.Y; P.a = a_use; P.a0 = a0_use;
    P.partial = (double*)sh.partial.p;
    P.partial0 = (double*)sh.partial.p + (size_t)pl.chunks * nrows * D;
    P.exp2_tbl = sh.ctx->exp2_tbl;
    P.sop = g->sop_grad;
    P.row0 = sh.r0; P.nrows = nrows; P.m = g->m; P.cols_per_chunk = pl.cols_per_chunk;
    P.single = g->prog.single;
    P.coef = g->coef_grad;
    if (g->prog.single) P.atom = g->prog.atoms[g->prog.terms[0].fac[0].atom];
    const int variant = g->prog.dotproduct ? 2 : ((g->prog.single && P.atom.v.kind == CF_ATOM_EQ) ? 0 : 1);
    CF_CUDA(g->entry->grad[vg][variant](P, dim3(pl.row_tiles, pl.chunks), stream));
    g->last_launches++;
    const int blocks = (int)std::min<int64_t>((nrows * bs + 255) / 256, 8192);
    grad_reduce_partials<<<blocks, 256, 0, stream>>>((const double*)sh.partial.p, P.partial0, pl.chunks, nrows, D, d, vg, d_y, d_yin,
                                                     alpha, beta, *peers);
    CF_CUDA(cudaGetLastError());
    g->last_launches++;
    return CF_OK;
}

// upload one point set: raw d x n (leading dimension ld) -> padded AoS on the device, squared norms, finiteness check.
// No host-side packing: one 2-D copy, one pad kernel (only when D != d) and one validation/norm kernel.
int upload_points(int dtype, const void* H, int64_t ld, int64_t n, int d, int D, void** dX, void** dN, Buf& scratch,
                  double* flags_dev, cudaStream_t st, const double* ard_scale_dev = nullptr) {
    const size_t es = esize(dtype);
    CF_CUDA(dev_alloc(dX, std::max<size_t>(16, (size_t)n * D * es)));
    CF_CUDA(dev_alloc(dN, std::max<size_t>(16, (size_t)n * es)));
    if (n == 0) return CF_OK;
    // a 2-D copy of n rows of d*es bytes is processed row by row by the driver (150 ms for 2^20 points): use the 1-D form
    // whenever the host buffer is dense (ld == d), which is the layout of a Julia Matrix / vecofvec column views
    auto h2d = [&](void* dst) -> cudaError_t {
        if (ld == d) return cudaMemcpyAsync(dst, H, (size_t)n * d * es, cudaMemcpyHostToDevice, st);
        return cudaMemcpy2DAsync(dst, (size_t)d * es, H, (size_t)ld * es, (size_t)d * es, n, cudaMemcpyHostToDevice, st);
    };
    if (D == d) {
        CF_CUDA(h2d(*dX));
    } else {
        if (int rc = scratch.ensure((size_t)n * d * es)) return rc;
        CF_CUDA(h2d(scratch.p));
        const int blocks = (int)std::min<int64_t>((n * D + 255) / 256, 8192);
        if (dtype == CF_F64) cf_pad_points<double><<<blocks, 256, 0, st>>>((const double*)scratch.p, d, d, (double*)*dX, D, n);
        else cf_pad_points<float><<<blocks, 256, 0, st>>>((const float*)scratch.p, d, d, (float*)*dX, D, n);
        CF_CUDA(cudaGetLastError());
    }
    if (ard_scale_dev) {  // ARD: the metric is applied to the device copy of the points (1/sqrt(l_c) per coordinate)
        const int sb = (int)std::min<int64_t>((n * D + 255) / 256, 8192);
        if (dtype == CF_F64) cf_scale_coords_kernel<double><<<sb, 256, 0, st>>>((double*)*dX, D, d, n, ard_scale_dev);
        else cf_scale_coords_kernel<float><<<sb, 256, 0, st>>>((float*)*dX, D, d, n, ard_scale_dev);
        CF_CUDA(cudaGetLastError());
    }
    const int nb = (int)std::min<int64_t>((n + 255) / 256, 4096);
    if (dtype == CF_F64) cf_sqnorm_validate_kernel<double><<<nb, 256, 0, st>>>((const double*)*dX, D, n, (double*)*dN, flags_dev);
    else cf_sqnorm_validate_kernel<float><<<nb, 256, 0, st>>>((const float*)*dX, D, n, (float*)*dN, flags_dev);
    CF_CUDA(cudaGetLastError());
    return CF_OK;
}

// which kernels may use r2 = |x|^2 + |y|^2 - 2 x.y for this handle (see the comments inside): called at create time and for the
// Float64 shadow of a Float32 handle
void set_norm_flags(cf_gramian_s* g, double max_sq) {
    const int dtype = g->dtype, d = g->d;
    // r2 from norms (multi-RHS kernel) only for larger d and well-scaled data:
    // |delta r2| <= (d + 2) eps (|x|^2 + |y|^2) must stay below 1e-13 (Float64) / 1e-5 (Float32)
    {
        const double eps = dtype == CF_F64 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
        const double bound = dtype == CF_F64 ? 1e-13 : 1e-5;
        // ... times the largest |dk/dr2| of the program's r2-atoms (short length scales amplify an absolute error in r2):
        // EQ exp(c r2): |c|;  MaternP(p >= 1): |d_1| / l^2 (its slope at 0, the maximum);  RQ (1 + w r2)^-a: a w.
        // exp(-sqrt(r2)) (Exp = MaternP(0)) is not differentiable in r2 at 0: an absolute error of 1e-16 in r2 of (nearly)
        // coincident points would become 1e-8 in k, so programs with that atom always keep direct differences.
        // LINE atoms use x.y, which the tensor-core chain reproduces exactly.
        // (atoms are bounded by 1, so the sum of the atoms' slopes bounds the slope of any product of them)
        // A product of powers prod_f atom_f^p_f (atoms bounded by 1) has slope <= sum_f p_f slope_f; the program's slope is bounded by
        // the largest such sum over its terms, weighted by the coefficients' share.
        bool sqrt_atom = false;
        std::vector<double> aslope(g->prog.natoms, 0.0);
        for (int i = 0; i < g->prog.natoms; i++) {
            const cf_atom& A = g->prog.atoms[i];
            if (A.v.kind == CF_ATOM_EQ) aslope[i] = std::fabs(A.v.e.c);
            else if (A.v.kind == CF_ATOM_MATERN && A.v.p == 0) sqrt_atom = true;
            else if (A.v.kind == CF_ATOM_MATERN) aslope[i] = std::fabs(A.tay[1]) * A.inv_l2;
            else if (A.v.kind == CF_ATOM_RQ_INT || A.v.kind == CF_ATOM_RQ_REAL) aslope[i] = A.v.alpha * A.v.w;
        }
        double slope = 0.0;
        for (int t = 0; t < g->prog.nterms; t++) {
            double ts = 0.0;
            for (int f = 0; f < g->prog.terms[t].nfac; f++) ts += g->prog.terms[t].fac[f].power * aslope[g->prog.terms[t].fac[f].atom];
            slope = std::max(slope, ts);
        }
        // accumulation factor: the random-walk value sqrt(d + 2).  The worst case (d + 2) is sqrt(d + 2) <= 5.9 times larger for
        // d <= 32, i.e. still below 6e-13 (Float64) / 6e-5 relative error of a kernel entry when this check passes at 1e-13 / 1e-5,
        // and unit-variance points (randn(d), the reference's README data) keep the tensor-core kernels at d = 32
        const double growth = std::sqrt((double)(d + 2));
        g->use_norms = (d >= 8) && !sqrt_atom && (growth * eps * 2.0 * max_sq * slope < bound);
        // the derivative operators also need k'' (one more factor of the slope) and MaternP(1) has a 1/sqrt(r2) term in k''
        bool smooth2 = true;
        for (int i = 0; i < g->prog.natoms; i++)
            if (g->prog.atoms[i].v.kind == CF_ATOM_MATERN && g->prog.atoms[i].v.p < 2) smooth2 = false;
        g->use_norms_grad = g->use_norms && smooth2 && (growth * eps * 2.0 * max_sq * slope * slope < bound);
        // the scaled-domain EQ kernel (gram_mvm_eq.cuh) has no clamp: besides the cancellation bound, the exponent of the
        // farthest pair, |c| (|x| + |y|)^2 <= 4 |c| max|x|^2, must stay clear of the 2^-1022 underflow (ln 2^-1022 = -708)
        g->eq_fast = dtype == CF_F64 && g->kind == CF_ATOM_EQ && (growth * eps * 2.0 * max_sq * slope < bound) &&
                     (4.0 * max_sq * slope < 600.0);
        // the MaternP form of that kernel (FAST = 2): p >= 1 (p = 0 is exp(-sqrt(r2)), not differentiable in r2 at 0: an error of 1e-16 in
        // r2 of coincident points would become 1e-8 in k), the same cancellation bound with the atom's largest slope |d_1| / l^2, and the
        // exponent of the farthest pair, |c| (|x| + |y|) <= 2 |c| max|x|, clear of the underflow
        g->mat_fast = false;
        if (dtype == CF_F64 && g->kind == CF_ATOM_MATERN && g->prog.single) {
            const cf_atom& A = g->prog.atoms[g->prog.terms[0].fac[0].atom];
            g->mat_fast = A.v.p >= 1 && (growth * eps * 2.0 * max_sq * slope < bound) && (2.0 * std::sqrt(max_sq) * std::fabs(A.v.e.c) < 600.0);
        }
    }
}

int destroy_impl(cf_gramian_s* g) {
    if (g->shadow64) { destroy_impl(g->shadow64); g->shadow64 = nullptr; }
    for (auto& sh : g->shards) {
        if (sh.ctx) cudaSetDevice(sh.ctx->dev);
        if (sh.stream) cudaStreamSynchronize(sh.stream); // nothing of this handle may still be running when its memory returns to the pool
        if (g->user_stream) {  // ... including work queued on the caller's stream by the last asynchronous *_mul_device call
            if (cudaStreamSynchronize((cudaStream_t)g->user_stream) != cudaSuccess) cudaGetLastError();  // (the caller may have destroyed it)
        }
        if (sh.Y && sh.Y != sh.X) dev_free(sh.Y);
        dev_free(sh.X);
        if (sh.yn && sh.yn != sh.xn) dev_free(sh.yn);
        dev_free(sh.xn);
        sh.a.release(); sh.y.release(); sh.partial.release(); sh.apad.release(); sh.ypad.release(); sh.at.release(); sh.sym_items.release(); sh.bsym.release(); sh.sym_col.release(); sh.bd_t.release(); sh.bd_s.release();
        sh.xp.release(); sh.yp.release(); sh.ap.release(); sh.yc_hi.release(); sh.yc_lo.release(); sh.yc_n.release(); sh.ac_hi.release(); sh.ac_lo.release(); sh.yt.release(); sh.au.release(); sh.cv_in.release(); sh.cv_out.release();
        for (auto& b : sh.cg) b.release();
        if (sh.ev0) cudaEventDestroy(sh.ev0);
        if (sh.ev1) cudaEventDestroy(sh.ev1);
        if (sh.stream) cudaStreamDestroy(sh.stream);
    }
    delete g;
    return CF_OK;
}

// The Float64 shadow of a Float32 handle (cf_gramian_s::shadow64): device-side conversion of the padded points, norms and scale flags
// recomputed in Float64, same devices and row range.  Built once, on the first call that needs it.
int ensure_shadow64(cf_gramian_s* g, cf_gramian_s** out) {
    if (g->dtype == CF_F64) { *out = g; return CF_OK; }
    if (!g->shadow64) {
        cf_gramian_s* h = new (std::nothrow) cf_gramian_s;
        if (!h) return fail(CF_ERR_INTERNAL, "out of host memory");
        h->dtype = CF_F64; h->d = g->d; h->D = g->D; h->n = g->n; h->m = g->m; h->symmetric = g->symmetric;
        h->prog = g->prog; h->sop_val = g->sop_val; h->sop_grad = g->sop_grad; h->grad_ok = g->grad_ok;
        h->kind = g->kind; h->coef = g->coef; h->coef_grad = g->coef_grad; h->entry = g->entry; h->opt_symmetric = g->opt_symmetric;
        h->shards.resize(g->shards.size());
        double max_sq = 0;
        int rc = CF_OK;
        for (size_t q = 0; q < g->shards.size() && !rc; q++) {
            Shard& src = g->shards[q];
            Shard& dst = h->shards[q];
            dst.ctx = src.ctx;
            auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) rc = fail(CF_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e)); };
            cu(cudaSetDevice(src.ctx->dev), "cudaSetDevice");
            cu(cudaStreamCreateWithFlags(&dst.stream, cudaStreamNonBlocking), "cudaStreamCreate");
            cu(cudaEventCreate(&dst.ev0), "cudaEventCreate");
            cu(cudaEventCreate(&dst.ev1), "cudaEventCreate");
            if (rc) break;
            cu(cudaStreamSynchronize(src.stream), "cudaStreamSynchronize");
            double* flags = nullptr;
            cu(dev_alloc((void**)&flags, 16), "cudaMallocAsync");
            if (rc) break;
            cu(cudaMemsetAsync(flags, 0, 16, dst.stream), "cudaMemsetAsync");
            auto conv = [&](const void* X32, int64_t cnt, void** X64, void** N64) {
                cu(dev_alloc(X64, std::max<size_t>(16, (size_t)cnt * g->D * 8)), "cudaMallocAsync");
                cu(dev_alloc(N64, std::max<size_t>(16, (size_t)cnt * 8)), "cudaMallocAsync");
                if (rc || cnt == 0) return;
                const int cb = (int)std::min<int64_t>((cnt * g->D + 255) / 256, 8192);
                cf_convert_kernel<float, double><<<cb, 256, 0, dst.stream>>>((const float*)X32, (double*)*X64, cnt * g->D);
                cf_sqnorm_validate_kernel<double><<<(int)std::min<int64_t>((cnt + 255) / 256, 4096), 256, 0, dst.stream>>>((const double*)*X64, g->D, cnt,
                                                                                                                      (double*)*N64, flags);
                cu(cudaGetLastError(), "point conversion");
            };
            conv(src.X, g->n, &dst.X, &dst.xn);
            if (src.Y != src.X) conv(src.Y, g->m, &dst.Y, &dst.yn);
            else { dst.Y = dst.X; dst.yn = dst.xn; }
            unsigned long long hf[2] = {0, 0};
            cu(cudaMemcpyAsync(hf, flags, 16, cudaMemcpyDeviceToHost, dst.stream), "cudaMemcpyAsync");
            cu(cudaStreamSynchronize(dst.stream), "cudaStreamSynchronize");
            dev_free(flags);
            double ms;
            std::memcpy(&ms, &hf[1], 8);
            max_sq = std::max(max_sq, ms);
        }
        if (rc) { destroy_impl(h); return rc; }
        set_norm_flags(h, max_sq);
        h->max_sq = max_sq;
        g->shadow64 = h;
    }
    cf_gramian_s* h = g->shadow64;
    h->row_begin = g->row_begin; h->row_end = g->row_end; h->opt_symmetric = g->opt_symmetric;
    split_rows(h);
    *out = h;
    return CF_OK;
}
// does this product of a Float32 handle have to run on the Float64 shadow?
inline bool needs_shadow(const cf_gramian_s* g, int deriv) { return g->dtype == CF_F32 && (deriv != 0 || !g->entry); }

}  // namespace

extern "C" {

int cf_version(void) { return CF_VERSION; }
const char* cf_last_error(void) { return g_err.c_str(); }

int cf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cf_init(int ngpus, const int* devices) {
    int have = cf_device_count();
    if (ngpus < 1) return fail(CF_ERR_BAD_ARGUMENT, "cf_init: ngpus must be >= 1");
    if (have < 1) return fail(CF_ERR_CUDA, "cf_init: no CUDA device available (this library has no CPU path)");
    std::vector<int> devs;
    for (int i = 0; i < ngpus; i++) {
        int dv = devices ? devices[i] : i;
        if (dv < 0 || dv >= have) return fail(CF_ERR_BAD_ARGUMENT, "cf_init: device %d out of range (have %d)", dv, have);
        devs.push_back(dv);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_devices = devs;
    return CF_OK;
}

int cf_gramian_create(cf_gramian_t* out, const cf_knode_t* prog, int nnodes, int dtype, int d, int64_t n, const void* X,
                      int64_t ldx, int64_t m, const void* Y, int64_t ldy) {
    if (!out) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_create: out is NULL");
    *out = nullptr;
    if (dtype != CF_F32 && dtype != CF_F64) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_create: dtype must be CF_F32 or CF_F64");
    if (d < 1) return fail(CF_ERR_DIMENSION, "cf_gramian_create: point dimension d = %d must be >= 1", d);
    if (n < 0 || m < 0) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_create: negative size");
    if ((n > 0 && !X) || (Y == nullptr && m != n)) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_create: X is NULL, or Y is NULL with m != n");
    if (ldx < d || (Y && ldy < d)) return fail(CF_ERR_DIMENSION, "cf_gramian_create: leading dimension smaller than d");
    cf_program lowered;
    cf_sop_val sop_val;
    cf_sop_grad sop_grad;
    bool grad_ok = false;
    std::vector<double> ard;  // ARD length scales l_c (empty: none)
    try {
        lowered = cf::lower(prog, nnodes, &ard);
        cf::to_sop_val(lowered, sop_val);
        grad_ok = cf::to_sop_grad(lowered, sop_grad);
    } catch (const cf::LowerError& e) {
        return fail(e.code, "cf_gramian_create: %s", e.msg.c_str());
    } catch (...) {
        return fail(CF_ERR_INTERNAL, "cf_gramian_create: unexpected failure while lowering the kernel program");
    }
    if (!ard.empty() && (int)ard.size() != d)
        return fail(CF_ERR_DIMENSION, "cf_gramian_create: ARD has %d length scales, the points have dimension %d", (int)ard.size(), d);
    for (double& l : ard) l = 1.0 / std::sqrt(l);  // coordinate scale
    const cf_kernel_entry* entry = find_entry(d);  // nullptr for d > 32: tiled contraction kernels (bigd.cuh; Float32 handles: on the Float64 shadow)
    if (d > (1 << 20)) return fail(CF_ERR_UNSUPPORTED, "cf_gramian_create: d = %d is too large", d);
    if (cf_device_count() < 1) return fail(CF_ERR_CUDA, "cf_gramian_create: no CUDA device available (this library has no CPU path)");

    cf_gramian_s* g = new (std::nothrow) cf_gramian_s;
    if (!g) return fail(CF_ERR_INTERNAL, "out of host memory");
    g->dtype = dtype; g->d = d; g->D = entry ? entry->D : ((d + CF_BD_K - 1) / CF_BD_K) * CF_BD_K; g->n = n; g->m = m;
    g->symmetric = (Y == nullptr);
    g->row_begin = 0; g->row_end = n;
    g->prog = lowered;
    g->sop_val = sop_val; g->sop_grad = sop_grad; g->grad_ok = grad_ok;
    g->opt_symmetric = true;  // y === x: each unordered pair evaluated once (deterministic, gram_mvm_sym.cuh)
    if (const char* e = std::getenv("COVFN_SYMMETRIC")) g->opt_symmetric = std::atoi(e) != 0;
    g->entry = entry;
    if (lowered.single) {
        const cf_atom& A = lowered.atoms[lowered.terms[0].fac[0].atom];
        g->coef = lowered.terms[0].coef;
        g->coef_grad = lowered.terms[0].coef;
        g->kind = (A.v.kind == CF_ATOM_EQ || A.v.kind == CF_ATOM_MATERN || A.v.kind == CF_ATOM_RQ_INT) ? A.v.kind : CF_ATOM_SOP;
        if (g->kind == CF_ATOM_SOP) g->coef = 1.0; // generic path applies the coefficient itself
    } else {
        g->kind = CF_ATOM_SOP;
        g->coef = 1.0;
    }

    std::vector<int> devs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        devs = g_devices;
    }
    if (devs.empty()) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) { delete g; return fail(CF_ERR_CUDA, "cudaGetDevice failed"); }
        devs.push_back(cur);
    }
    const size_t es = esize(dtype);
    double max_sq = 0;
    int rc = CF_OK;

    g->shards.resize(devs.size());
    for (size_t s = 0; s < devs.size(); s++) {
        Shard& sh = g->shards[s];
        rc = get_ctx(devs[s], &sh.ctx);
        if (rc) { destroy_impl(g); return rc; }
#define CF_CREATE_CUDA(call)                                                                                  \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            destroy_impl(g);                                                                                  \
            return fail(CF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));                        \
        }                                                                                                     \
    } while (0)
        CF_CREATE_CUDA(cudaSetDevice(devs[s]));
        CF_CREATE_CUDA(cudaStreamCreateWithFlags(&sh.stream, cudaStreamNonBlocking));
        CF_CREATE_CUDA(cudaEventCreate(&sh.ev0));
        CF_CREATE_CUDA(cudaEventCreate(&sh.ev1));
        {
            // flags[0] = number of non-finite coordinates, flags[1] = max squared norm (as ordered bits)
            double* flags = nullptr;
            CF_CREATE_CUDA(dev_alloc((void**)&flags, 16));
            CF_CREATE_CUDA(cudaMemsetAsync(flags, 0, 16, sh.stream));
            double* ard_dev = nullptr;
            if (!ard.empty()) {
                CF_CREATE_CUDA(dev_alloc((void**)&ard_dev, ard.size() * sizeof(double)));
                CF_CREATE_CUDA(cudaMemcpyAsync(ard_dev, ard.data(), ard.size() * sizeof(double), cudaMemcpyHostToDevice, sh.stream));
            }
            rc = upload_points(dtype, X, ldx, n, d, g->D, &sh.X, &sh.xn, sh.apad, flags, sh.stream, ard_dev);
            if (!rc) {
                if (Y) rc = upload_points(dtype, Y, ldy, m, d, g->D, &sh.Y, &sh.yn, sh.apad, flags, sh.stream, ard_dev);
                else { sh.Y = sh.X; sh.yn = sh.xn; }
            }
            unsigned long long hflags[2] = {0, 0};
            if (!rc && cudaMemcpyAsync(hflags, flags, 16, cudaMemcpyDeviceToHost, sh.stream) != cudaSuccess) rc = fail(CF_ERR_CUDA, "flag copy failed");
            if (!rc && cudaStreamSynchronize(sh.stream) != cudaSuccess) rc = fail(CF_ERR_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            dev_free(flags);
            dev_free(ard_dev);
            if (!rc && hflags[0] != 0) rc = fail(CF_ERR_NONFINITE, "%llu point coordinates are not finite (NaN or Inf)", hflags[0]);
            if (rc) { destroy_impl(g); return rc; }
            double ms;
            std::memcpy(&ms, &hflags[1], 8);
            max_sq = std::max(max_sq, ms);
        }
#undef CF_CREATE_CUDA
    }
    set_norm_flags(g, max_sq);
    g->max_sq = max_sq;
    (void)es;
    split_rows(g);
    *out = g;
    return CF_OK;
}

int cf_gramian_destroy(cf_gramian_t g) {
    if (!g) return CF_OK;
    return destroy_impl(g);
}

int cf_gramian_size(cf_gramian_t g, int64_t* n, int64_t* m, int* d, int* dtype) {
    if (int rc = check_handle(g)) return rc;
    if (n) *n = g->n;
    if (m) *m = g->m;
    if (d) *d = g->d;
    if (dtype) *dtype = g->dtype;
    return CF_OK;
}

int cf_gramian_set_option(cf_gramian_t g, int option, int value) {
    if (int rc = check_handle(g)) return rc;
    std::lock_guard<std::mutex> lk(g->mu);
    if (option == CF_OPT_SYMMETRIC) { g->opt_symmetric = value != 0; return CF_OK; }
    return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_set_option: unknown option %d", option);
}

int cf_gramian_set_row_range(cf_gramian_t g, int64_t row_begin, int64_t row_end) {
    if (int rc = check_handle(g)) return rc;
    if (row_begin < 0 || row_end < row_begin || row_end > g->n)
        return fail(CF_ERR_DIMENSION, "cf_gramian_set_row_range: [%lld, %lld) is not inside [0, %lld)", (long long)row_begin,
                    (long long)row_end, (long long)g->n);
    std::lock_guard<std::mutex> lk(g->mu);
    g->row_begin = row_begin; g->row_end = row_end;
    split_rows(g);
    return CF_OK;
}

static int check_derivative(cf_gramian_s* g) {
    if (!g->prog.isotropic && !g->prog.dotproduct)
        return fail(CF_ERR_UNSUPPORTED, "derivative operators: kernel has neither the IsotropicInput nor the DotProductInput trait");
    if (!g->grad_ok) return fail(CF_ERR_UNSUPPORTED, "derivative operators: kernel too complex (more than 4 terms or 3 base kernels)");
    return CF_OK;
}
static int check_vg_dim(cf_gramian_s*, int) { return CF_OK; }  // (the ValueGradientKernel covers every d since round 2: bigd_jet_vg_kernel)

static int mul_host_impl(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs, double alpha,
                         double beta, int deriv) {
    const bool gradient = deriv != 0;
    const int vg = deriv == 2 ? 1 : 0;
    if (int rc = check_handle(g)) return rc;
    if (nrhs < 0) return fail(CF_ERR_BAD_ARGUMENT, "nrhs is negative");
    const int64_t blk = deriv == 0 ? 1 : g->d + vg;
    const int64_t rows = (g->row_end - g->row_begin) * blk, cols = g->m * blk;
    if (nrhs == 0 || rows == 0) return CF_OK;
    if (!y || (!x && cols > 0)) return fail(CF_ERR_BAD_ARGUMENT, "NULL vector pointer");
    if (nrhs > 1 && (ldy < rows || ldx < cols))
        return fail(CF_ERR_DIMENSION, "leading dimension too small: ldy = %lld (rows %lld), ldx = %lld (cols %lld)", (long long)ldy,
                    (long long)rows, (long long)ldx, (long long)cols);
    if (gradient) {
        if (int rc = check_derivative(g)) return rc;
        if (int rc = check_vg_dim(g, deriv)) return rc;
    }
    if (nrhs == 1) { ldy = rows; ldx = cols; }
    if (needs_shadow(g, deriv)) {
        // Float32 handle, operator that exists in Float64 only: convert the vectors, run on the Float64 shadow, convert back
        cf_gramian_s* h = nullptr;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (int rc = ensure_shadow64(g, &h)) return rc;
        }
        std::vector<double> xd((size_t)cols * nrhs), yd((size_t)rows * nrhs, 0.0);
        const float* xf = (const float*)x;
        float* yf = (float*)y;
        for (int64_t c = 0; c < nrhs; c++) {
            for (int64_t q = 0; q < cols; q++) xd[(size_t)c * cols + q] = (double)xf[c * ldx + q];
            if (beta != 0.0)
                for (int64_t q = 0; q < rows; q++) yd[(size_t)c * rows + q] = (double)yf[c * ldy + q];
        }
        if (int rc = mul_host_impl(h, yd.data(), rows, xd.data(), cols, nrhs, alpha, beta, deriv)) return rc;
        for (int64_t c = 0; c < nrhs; c++)
            for (int64_t q = 0; q < rows; q++) yf[c * ldy + q] = (float)yd[(size_t)c * rows + q];
        g->last_ms = h->last_ms; g->last_launches = h->last_launches;
        return CF_OK;
    }
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t es = esize(g->dtype);
    g->last_launches = 0;
    // stage 1: copies in + launches on every shard
    for (auto& sh : g->shards) {
        const int64_t srows = (sh.r1 - sh.r0) * blk;
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        if (int rc = sh.a.ensure(std::max<size_t>(16, (size_t)cols * nrhs * es))) return rc;
        if (int rc = sh.y.ensure(std::max<size_t>(16, (size_t)srows * nrhs * es))) return rc;
        if (cols > 0)
            CF_CUDA(cudaMemcpy2DAsync(sh.a.p, cols * es, x, ldx * es, cols * es, nrhs, cudaMemcpyHostToDevice, sh.stream));
        if (srows == 0) continue;
        const int64_t off = (sh.r0 - g->row_begin) * blk;
        if (beta != 0.0)
            CF_CUDA(cudaMemcpy2DAsync(sh.y.p, srows * es, (const char*)y + off * es, ldy * es, srows * es, nrhs,
                                      cudaMemcpyHostToDevice, sh.stream));
        CF_CUDA(cudaEventRecord(sh.ev0, sh.stream));
        if (!gradient && nrhs > 1 && g->entry) {
            int rc = launch_mm(g, sh, sh.y.p, srows, sh.a.p, cols, nrhs, alpha, beta, sh.stream);
            if (rc) return rc;
        } else {
            for (int64_t c = 0; c < nrhs; c++) {
                void* yc = (char*)sh.y.p + (size_t)c * srows * es;
                const void* ac = (const char*)sh.a.p + (size_t)c * cols * es;
                int rc = gradient ? launch_grad(g, sh, (double*)yc, (const double*)yc, (const double*)ac, alpha, beta, sh.stream, vg)
                                  : launch_mvm(g, sh, yc, yc, ac, alpha, beta, sh.stream);
                if (rc) return rc;
            }
        }
        CF_CUDA(cudaEventRecord(sh.ev1, sh.stream));
    }
    // stage 2: results back (a device-to-host copy into pageable memory blocks the host, so it must not sit between
    // the launches of different shards), then wait
    for (auto& sh : g->shards) {
        const int64_t srows = (sh.r1 - sh.r0) * blk;
        if (srows == 0) continue;
        const int64_t off = (sh.r0 - g->row_begin) * blk;
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        CF_CUDA(cudaMemcpy2DAsync((char*)y + off * es, ldy * es, sh.y.p, srows * es, srows * es, nrhs, cudaMemcpyDeviceToHost,
                                  sh.stream));
    }
    float worst = 0;
    for (auto& sh : g->shards) {
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        CF_CUDA(cudaStreamSynchronize(sh.stream));
        if (sh.r1 > sh.r0) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, sh.ev0, sh.ev1) == cudaSuccess) worst = std::max(worst, ms);
        }
    }
    g->last_ms = worst;
    return CF_OK;
}

// single-process multi-GPU, symmetric Gramian, one right-hand side, full row range: every device evaluates the unordered pairs of
// its row tiles into a partial vector, device 0 sums the partial vectors with peer loads in device order.  Returns 1 if the variant
// does not apply (the caller runs the row-block path).
static int mul_host_sym_spmd(cf_gramian_t g, double* y, const double* x, double alpha, double beta) {
    const int S = (int)g->shards.size();
    const int64_t n = g->n;
    if (S < 2 || g->row_begin != 0 || g->row_end != n || !g->entry) return 1;
    for (auto& sh : g->shards) {
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        if (int rc = sh.a.ensure((size_t)n * 8)) return rc;
        if (int rc = sh.ypad.ensure((size_t)n * 8)) return rc;
    }
    if (!sym_applicable(g, g->shards[0].a.p) || enable_peers(g)) return 1;
    Shard& s0 = g->shards[0];
    CF_CUDA(cudaSetDevice(s0.ctx->dev));
    if (int rc = s0.y.ensure((size_t)n * 8)) return rc;
    if (beta != 0.0) CF_CUDA(cudaMemcpyAsync(s0.y.p, y, n * 8, cudaMemcpyHostToDevice, s0.stream));
    g->last_launches = 0;
    for (int s = 0; s < S; s++) {
        Shard& sh = g->shards[s];
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        CF_CUDA(cudaMemcpyAsync(sh.a.p, x, n * 8, cudaMemcpyHostToDevice, sh.stream));
        CF_CUDA(cudaEventRecord(sh.ev0, sh.stream));
        int rc = launch_mvm_sym(g, sh, (double*)sh.ypad.p, (const double*)s0.y.p, (const double*)sh.a.p, alpha, s == 0 ? beta : 0.0, sh.stream,
                                nullptr, s, S);
        if (rc) return rc;
        CF_CUDA(cudaEventRecord(sh.ev1, sh.stream));
    }
    cf_parts_in parts;
    std::memset(&parts, 0, sizeof(parts));
    parts.n = S;
    for (int s = 0; s < S; s++) parts.ptr[s] = (const double*)g->shards[s].ypad.p;
    CF_CUDA(cudaSetDevice(s0.ctx->dev));
    for (int s = 1; s < S; s++) CF_CUDA(cudaStreamWaitEvent(s0.stream, g->shards[s].ev1, 0));
    cf_sum_parts_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, s0.stream>>>(parts, 0, n, (double*)s0.y.p, nullptr, 0.0);
    CF_CUDA(cudaGetLastError());
    CF_CUDA(cudaMemcpyAsync(y, s0.y.p, n * 8, cudaMemcpyDeviceToHost, s0.stream));
    float worst = 0;
    for (auto& sh : g->shards) {
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        CF_CUDA(cudaStreamSynchronize(sh.stream));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sh.ev0, sh.ev1) == cudaSuccess) worst = std::max(worst, ms);
    }
    g->last_ms = worst;
    return CF_OK;
}

int cf_gramian_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs, double alpha, double beta) {
    if (g && nrhs == 1 && y && x && g->shards.size() > 1 && g->dtype == CF_F64) {
        std::unique_lock<std::mutex> lk(g->mu);
        const int rc = mul_host_sym_spmd(g, (double*)y, (const double*)x, alpha, beta);
        if (rc != 1) return rc;
    }
    return mul_host_impl(g, y, ldy, x, ldx, nrhs, alpha, beta, 0);
}

// y_full <- alpha K x + beta y_full computed COLLECTIVELY by the ranks of the communicator (cf_comm_init): every rank passes the
// same x and receives the complete y (the form chained products need).  Symmetric Float64 Gramians: unordered pairs of the rank's
// row tiles (cyclic), then ncclAllReduce(sum); otherwise the rank's row block, then the in-place all-gather.
int cf_gramian_mul_collective_device(cf_gramian_t g, void* d_y_full, const void* d_x, double alpha, double beta, void* stream) {
    if (int rc = check_handle(g)) return rc;
    if (!cfcomm::active()) return fail(CF_ERR_NCCL, "cf_gramian_mul_collective_device: no communicator (cf_comm_init)");
    if (g->shards.size() != 1) return fail(CF_ERR_UNSUPPORTED, "cf_gramian_mul_collective_device: needs a single-device handle");
    if (!d_y_full || !d_x) return fail(CF_ERR_BAD_ARGUMENT, "NULL device pointer");
    cfcomm::Comm& cm = cfcomm::comm();
    int64_t r0, r1;
    cfcomm::row_block(g->n, cm.rank, cm.world, &r0, &r1);
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->row_begin != r0 || g->row_end != r1)
        return fail(CF_ERR_DIMENSION, "cf_gramian_mul_collective_device: rank %d must own rows [%lld, %lld)", cm.rank, (long long)r0, (long long)r1);
    Shard& sh = g->shards[0];
    CF_CUDA(cudaSetDevice(sh.ctx->dev));
    cudaStream_t st = stream ? (cudaStream_t)stream : sh.stream;
    const size_t es = esize(g->dtype);
    g->last_launches = 0;
    CF_CUDA(cudaEventRecord(sh.ev0, st));
    bool done = false;
    if (g->dtype == CF_F64 && sym_applicable(g, d_x)) {
        const int rc = launch_mvm_sym(g, sh, (double*)d_y_full, (const double*)d_y_full, (const double*)d_x, alpha, beta, st, nullptr, cm.rank, cm.world);
        if (rc != 0 && rc != 1) return rc;
        if (rc == 0) {
            CF_CUDA(cudaEventRecord(sh.ev1, st));
            const int nrc = cfcomm::allreduce_sum_f64((double*)d_y_full, g->n, st);
            if (nrc) return fail(CF_ERR_NCCL, "NCCL all-reduce failed: %s", cfcomm::api().GetErrorString(nrc));
            done = true;
        }
    }
    if (!done) {
        void* yb = (char*)d_y_full + (size_t)r0 * es;
        if (int rc = launch_mvm(g, sh, yb, yb, d_x, alpha, beta, st)) return rc;
        CF_CUDA(cudaEventRecord(sh.ev1, st));
        const int nrc = cfcomm::allgather_rows(d_y_full, g->n, 1, es, st);
        if (nrc) return fail(CF_ERR_NCCL, "NCCL all-gather failed: %s", cfcomm::api().GetErrorString(nrc));
    }
    if (!stream) {
        CF_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sh.ev0, sh.ev1) == cudaSuccess) g->last_ms = ms;
    }
    return CF_OK;
}
int cf_gradient_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs, double alpha, double beta) {
    return mul_host_impl(g, y, ldy, x, ldx, nrhs, alpha, beta, 1);
}
int cf_value_gradient_mul(cf_gramian_t g, void* y, int64_t ldy, const void* x, int64_t ldx, int64_t nrhs, double alpha, double beta) {
    return mul_host_impl(g, y, ldy, x, ldx, nrhs, alpha, beta, 2);
}

static int mul_device_impl(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx, int64_t nrhs, double alpha,
                           double beta, void* stream, int deriv) {
    const bool gradient = deriv != 0;
    const int vg = deriv == 2 ? 1 : 0;
    if (int rc = check_handle(g)) return rc;
    if (g->shards.size() != 1) return fail(CF_ERR_UNSUPPORTED, "device-pointer multiply needs a single-device handle");
    const int64_t blk = deriv == 0 ? 1 : g->d + vg;
    const int64_t rows = (g->row_end - g->row_begin) * blk, cols = g->m * blk;
    if (nrhs <= 0 || rows == 0) return CF_OK;
    if (!d_y || (!d_x && cols > 0)) return fail(CF_ERR_BAD_ARGUMENT, "NULL device pointer");
    if (nrhs == 1) { ldy = rows; ldx = cols; }
    if (ldy < rows || ldx < cols) return fail(CF_ERR_DIMENSION, "leading dimension too small");
    if (gradient) {
        if (int rc = check_derivative(g)) return rc;
        if (int rc = check_vg_dim(g, deriv)) return rc;
    }
    if (needs_shadow(g, deriv)) {
        // Float32 device vectors, Float64-only operator: device-side conversion around the Float64 shadow's product (same stream)
        cf_gramian_s* h = nullptr;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (int rc = ensure_shadow64(g, &h)) return rc;
        }
        Shard& hs = h->shards[0];
        CF_CUDA(cudaSetDevice(hs.ctx->dev));
        cudaStream_t st = stream ? (cudaStream_t)stream : hs.stream;
        if (stream) { g->user_stream = stream; h->user_stream = stream; }
        // (buffers of their own: the operator launched below pads its input into apad / reduces through ypad)
        if (int rc = hs.cv_in.ensure((size_t)cols * nrhs * 8 + 16)) return rc;
        if (int rc = hs.cv_out.ensure((size_t)rows * nrhs * 8 + 16)) return rc;
        const int cb = (int)std::min<int64_t>((std::max(rows, cols) + 255) / 256, 8192);
        for (int64_t c = 0; c < nrhs; c++) {
            cf_convert_kernel<float, double><<<cb, 256, 0, st>>>((const float*)d_x + c * ldx, (double*)hs.cv_in.p + c * cols, cols);
            if (beta != 0.0) cf_convert_kernel<float, double><<<cb, 256, 0, st>>>((const float*)d_y + c * ldy, (double*)hs.cv_out.p + c * rows, rows);
        }
        CF_CUDA(cudaGetLastError());
        if (int rc = mul_device_impl(h, hs.cv_out.p, rows, hs.cv_in.p, cols, nrhs, alpha, beta, (void*)st, deriv)) return rc;
        for (int64_t c = 0; c < nrhs; c++)
            cf_convert_kernel<double, float><<<cb, 256, 0, st>>>((const double*)hs.cv_out.p + c * rows, (float*)d_y + c * ldy, rows);
        CF_CUDA(cudaGetLastError());
        if (!stream) CF_CUDA(cudaStreamSynchronize(st));
        g->last_launches = h->last_launches;
        return CF_OK;
    }
    std::lock_guard<std::mutex> lk(g->mu);
    Shard& sh = g->shards[0];
    CF_CUDA(cudaSetDevice(sh.ctx->dev));
    cudaStream_t st = stream ? (cudaStream_t)stream : sh.stream;
    if (stream) g->user_stream = stream;
    const size_t es = esize(g->dtype);
    g->last_launches = 0;
    CF_CUDA(cudaEventRecord(sh.ev0, st));
    if (!gradient && nrhs > 1 && g->entry) {
        int rc = launch_mm(g, sh, d_y, ldy, d_x, ldx, nrhs, alpha, beta, st);
        if (rc) return rc;
    } else {
        for (int64_t c = 0; c < nrhs; c++) {
            void* yc = (char*)d_y + (size_t)c * ldy * es;
            const void* ac = (const char*)d_x + (size_t)c * ldx * es;
            int rc = gradient ? launch_grad(g, sh, (double*)yc, (const double*)yc, (const double*)ac, alpha, beta, st, vg)
                              : launch_mvm(g, sh, yc, yc, ac, alpha, beta, st);
            if (rc) return rc;
        }
    }
    CF_CUDA(cudaEventRecord(sh.ev1, st));
    if (!stream) {
        CF_CUDA(cudaStreamSynchronize(st));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sh.ev0, sh.ev1) == cudaSuccess) g->last_ms = ms;
    }
    return CF_OK;
}

int cf_gramian_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx, int64_t nrhs, double alpha,
                          double beta, void* stream) {
    return mul_device_impl(g, d_y, ldy, d_x, ldx, nrhs, alpha, beta, stream, 0);
}
int cf_gradient_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx, int64_t nrhs, double alpha,
                           double beta, void* stream) {
    return mul_device_impl(g, d_y, ldy, d_x, ldx, nrhs, alpha, beta, stream, 1);
}
int cf_value_gradient_mul_device(cf_gramian_t g, void* d_y, int64_t ldy, const void* d_x, int64_t ldx, int64_t nrhs, double alpha,
                                 double beta, void* stream) {
    return mul_device_impl(g, d_y, ldy, d_x, ldx, nrhs, alpha, beta, stream, 2);
}

int cf_gramian_matrix(cf_gramian_t g, void* M, int64_t ldm) {
    if (int rc = check_handle(g)) return rc;
    const int64_t rows = g->row_end - g->row_begin;
    if (rows == 0 || g->m == 0) return CF_OK;
    if (!M) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_matrix: M is NULL");
    if (ldm < rows) return fail(CF_ERR_DIMENSION, "cf_gramian_matrix: ldm = %lld < rows = %lld", (long long)ldm, (long long)rows);
    std::lock_guard<std::mutex> lk(g->mu);
    const size_t es = esize(g->dtype);
    for (auto& sh : g->shards) {
        const int64_t srows = sh.r1 - sh.r0;
        if (srows == 0) continue;
        CF_CUDA(cudaSetDevice(sh.ctx->dev));
        // column panels bounded to 256 MiB of device scratch
        int64_t panel = std::max<int64_t>(1, std::min<int64_t>(g->m, (int64_t)((256u << 20) / (srows * es))));
        if (int rc = sh.ypad.ensure((size_t)srows * panel * es)) return rc;
        for (int64_t j0 = 0; j0 < g->m; j0 += panel) {
            const int64_t nj = std::min(panel, g->m - j0);
            int rc = launch_dense(g, sh, sh.ypad.p, srows, j0, nj, sh.stream);
            if (rc) return rc;
            CF_CUDA(cudaMemcpy2DAsync((char*)M + ((sh.r0 - g->row_begin) + j0 * ldm) * es, ldm * es, sh.ypad.p, srows * es,
                                      srows * es, nj, cudaMemcpyDeviceToHost, sh.stream));
            CF_CUDA(cudaStreamSynchronize(sh.stream));
        }
    }
    return CF_OK;
}

int cf_gramian_getindex(cf_gramian_t g, int64_t i, int64_t j, double* out) {
    if (int rc = check_handle(g)) return rc;
    if (!out) return fail(CF_ERR_BAD_ARGUMENT, "cf_gramian_getindex: out is NULL");
    if (i < 0 || i >= g->n || j < 0 || j >= g->m)
        return fail(CF_ERR_DIMENSION, "BoundsError: attempt to access %lld x %lld Gramian at index [%lld, %lld]", (long long)g->n,
                    (long long)g->m, (long long)i + 1, (long long)j + 1);
    std::lock_guard<std::mutex> lk(g->mu);
    Shard& sh = g->shards[0];
    CF_CUDA(cudaSetDevice(sh.ctx->dev));
    if (int rc = sh.ypad.ensure(16)) return rc;
    Shard tmp = sh; // view of a single row
    tmp.r0 = i; tmp.r1 = i + 1;
    int rc = launch_dense(g, tmp, sh.ypad.p, 1, j, 1, sh.stream);
    if (rc) return rc;
    double v64 = 0; float v32 = 0;
    if (g->dtype == CF_F64) CF_CUDA(cudaMemcpyAsync(&v64, sh.ypad.p, 8, cudaMemcpyDeviceToHost, sh.stream));
    else CF_CUDA(cudaMemcpyAsync(&v32, sh.ypad.p, 4, cudaMemcpyDeviceToHost, sh.stream));
    CF_CUDA(cudaStreamSynchronize(sh.stream));
    *out = g->dtype == CF_F64 ? v64 : (double)v32;
    return CF_OK;
}

int cf_cg_solve(cf_gramian_t g, double sigma2, void* x, const void* b, double reltol, int maxiter, int gradient, int* iters,
                double* resnorm) {
    if (int rc = check_handle(g)) return rc;
    if (!x || !b) return fail(CF_ERR_BAD_ARGUMENT, "cf_cg_solve: NULL vector");
    if (g->n != g->m) return fail(CF_ERR_DIMENSION, "cf_cg_solve: Gramian is %lld x %lld, not square", (long long)g->n, (long long)g->m);
    if (g->dtype == CF_F32) {
        // Float32 system: the reference's cg! iterates in Float32 with reltol = sqrt(eps(Float32)); here the vectors are converted and
        // the solve runs on the Float64 shadow with that tolerance (iterates at least as accurate), the solution is rounded back
        cf_gramian_s* h = nullptr;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (int rc = ensure_shadow64(g, &h)) return rc;
        }
        const int64_t N = g->n * (gradient == 0 ? 1 : g->d + (gradient == 2 ? 1 : 0));
        std::vector<double> xd((size_t)N), bd((size_t)N);
        for (int64_t q = 0; q < N; q++) { xd[q] = (double)((const float*)x)[q]; bd[q] = (double)((const float*)b)[q]; }
        if (reltol <= 0) reltol = std::sqrt(1.1920928955078125e-07);
        const int rc = cf_cg_solve(h, sigma2, xd.data(), bd.data(), reltol, maxiter, gradient, iters, resnorm);
        if (rc) return rc;
        for (int64_t q = 0; q < N; q++) ((float*)x)[q] = (float)xd[q];
        g->cg_total_ms = h->cg_total_ms; g->cg_mvm_ms = h->cg_mvm_ms; g->cg_gather_ms = h->cg_gather_ms; g->cg_products = h->cg_products;
        return CF_OK;
    }
    if (g->row_begin != 0 || g->row_end != g->n) {
        // multi-process mode: the handle of rank r must own exactly block r of the standard split
        if (!cfcomm::active() || g->shards.size() != 1)
            return fail(CF_ERR_UNSUPPORTED, "cf_cg_solve: handle is restricted to a row range (initialise cf_comm_init for the multi-process solve)");
        int64_t r0, r1;
        cfcomm::row_block(g->n, cfcomm::comm().rank, cfcomm::comm().world, &r0, &r1);
        if (g->row_begin != r0 || g->row_end != r1)
            return fail(CF_ERR_DIMENSION, "cf_cg_solve: rank %d must own rows [%lld, %lld), the handle has [%lld, %lld)", cfcomm::comm().rank,
                        (long long)r0, (long long)r1, (long long)g->row_begin, (long long)g->row_end);
    }
    if (gradient < 0 || gradient > 2) return fail(CF_ERR_BAD_ARGUMENT, "cf_cg_solve: gradient must be 0, 1 or 2");
    if (gradient) {
        if (int rc = check_derivative(g)) return rc;
        if (int rc = check_vg_dim(g, gradient)) return rc;
    }
    std::lock_guard<std::mutex> lk(g->mu);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = cg_solve_impl(g, sigma2, (double*)x, (const double*)b, reltol, maxiter, gradient, iters, resnorm);
    g->cg_total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int cf_cg_timing(cf_gramian_t g, double* total_ms, double* product_ms, double* gather_ms, int* products) {
    if (int rc = check_handle(g)) return rc;
    if (total_ms) *total_ms = g->cg_total_ms;
    if (product_ms) *product_ms = g->cg_mvm_ms;
    if (gather_ms) *gather_ms = g->cg_gather_ms;
    if (products) *products = g->cg_products;
    return CF_OK;
}

// ---- multi-process row sharding over NCCL (cf_comm.h) ---------------------------------------------------------------------
int cf_comm_unique_id(void* id, int bytes) {
    if (!id || bytes < 128) return fail(CF_ERR_BAD_ARGUMENT, "cf_comm_unique_id: need a 128-byte buffer");
    cfcomm::Api& a = cfcomm::api();
    if (!a.error.empty()) return fail(CF_ERR_NCCL, "cf_comm_unique_id: %s", a.error.c_str());
    cfcomm::ncclUniqueId u;
    const int rc = a.GetUniqueId(&u);
    if (rc) return fail(CF_ERR_NCCL, "ncclGetUniqueId failed: %s", a.GetErrorString(rc));
    std::memcpy(id, u.internal, 128);
    return CF_OK;
}

int cf_comm_init(const void* id, int rank, int world) {
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(CF_ERR_BAD_ARGUMENT, "cf_comm_init: bad arguments");
    if (cf_device_count() < 1) return fail(CF_ERR_CUDA, "cf_comm_init: no CUDA device available");
    cfcomm::Api& a = cfcomm::api();
    if (!a.error.empty()) return fail(CF_ERR_NCCL, "cf_comm_init: %s", a.error.c_str());
    cfcomm::Comm& c = cfcomm::comm();
    if (c.comm) return fail(CF_ERR_BAD_ARGUMENT, "cf_comm_init: communicator already initialised (cf_comm_destroy first)");
    int dev = 0;
    CF_CUDA(cudaGetDevice(&dev));
    cfcomm::ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    cfcomm::ncclComm_t comm = nullptr;
    const int rc = a.CommInitRank(&comm, world, u, rank);
    if (rc) return fail(CF_ERR_NCCL, "ncclCommInitRank failed: %s", a.GetErrorString(rc));
    c.comm = comm; c.rank = rank; c.world = world; c.dev = dev;
    return CF_OK;
}

int cf_comm_destroy(void) {
    cfcomm::Comm& c = cfcomm::comm();
    if (c.comm) cfcomm::api().CommDestroy(c.comm);
    c.comm = nullptr; c.rank = 0; c.world = 1; c.dev = -1;
    return CF_OK;
}

int cf_comm_info(int* rank, int* world, int* nccl_version) {
    cfcomm::Comm& c = cfcomm::comm();
    if (rank) *rank = c.rank;
    if (world) *world = c.comm ? c.world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        cfcomm::Api& a = cfcomm::api();
        if (a.error.empty() && a.GetVersion) a.GetVersion(nccl_version);
    }
    return CF_OK;
}

int cf_comm_allgather_rows(void* d_full, int64_t n, int64_t block, int dtype, void* stream) {
    if (!cfcomm::active()) return fail(CF_ERR_NCCL, "cf_comm_allgather_rows: no communicator (cf_comm_init)");
    if (!d_full || n < 0 || block < 1 || (dtype != CF_F32 && dtype != CF_F64)) return fail(CF_ERR_BAD_ARGUMENT, "cf_comm_allgather_rows: bad arguments");
    const int rc = cfcomm::allgather_rows(d_full, n, block, esize(dtype), (cudaStream_t)stream);
    if (rc) return fail(CF_ERR_NCCL, "NCCL all-gather failed: %s", cfcomm::api().GetErrorString(rc));
    return CF_OK;
}

int cf_last_timing(cf_gramian_t g, float* kernel_ms, int* launches) {
    if (int rc = check_handle(g)) return rc;
    if (kernel_ms) *kernel_ms = g->last_ms;
    if (launches) *launches = g->last_launches;
    return CF_OK;
}

int cf_peak_probe(int kind, int iters, double* lane_ops_per_s, float* ms) {
    if (cf_device_count() < 1) return fail(CF_ERR_CUDA, "cf_peak_probe: no CUDA device available");
    if (kind < 0 || kind > 2 || iters < 1) return fail(CF_ERR_BAD_ARGUMENT, "cf_peak_probe: bad arguments");
    return peak_probe_impl(kind, iters, lane_ops_per_s, ms);
}

int cf_jit_check(const cf_knode_t* prog, int nnodes, int d, int which, char* log, int loglen) {
    if (log && loglen > 0) log[0] = 0;
    cf_program lowered;
    cf_sop_val sop_val;
    try {
        std::vector<double> ard;
        lowered = cf::lower(prog, nnodes, &ard);
        cf::to_sop_val(lowered, sop_val);
    } catch (const cf::LowerError& e) {
        return fail(e.code, "cf_jit_check: %s", e.msg.c_str());
    } catch (...) {
        return fail(CF_ERR_INTERNAL, "cf_jit_check: unexpected failure while lowering the kernel program");
    }
    const cf_kernel_entry* entry = find_entry(d);
    if (!entry) return fail(CF_ERR_UNSUPPORTED, "cf_jit_check: d = %d > 32 has no specialised kernels", d);
    const int D = entry->D;
    const int* tu = entry->tune;
    const std::string sD = std::to_string(D), sop = std::to_string((int)CF_ATOM_SOP);
    std::string header, name;
    switch (which) {
        case 0:
            header = "gram_mvm.cuh";
            name = "gram_mvm_kernel<double, " + sD + ", " + sop + ", " + std::to_string(tu[0]) + ", " + std::to_string(tu[1]) + ", " +
                   std::to_string(tu[2]) + ", " + std::to_string(tu[3]) + ", " + std::to_string(tu[4]) + ">";
            break;
        case 1: header = "gram_mm_dmma.cuh"; name = "gram_mm_dmma_kernel<" + sD + ">"; break;
        case 2: header = "gram_mvm_dmma.cuh"; name = "gram_mvm_dmma_kernel<" + sD + ", " + sop + ">"; break;
        case 3: header = "gram_mm_tc5.cuh"; name = "gram_mm_tc5_kernel<" + sD + ">"; break;
        case 6: header = "gram_mm_tf32.cuh"; name = "gram_mm_tf32_kernel<" + sD + ">"; break;
        case 4: header = "gram_mvm_tf32.cuh"; name = "gram_mvm_tf32_kernel<" + sD + ", " + sop + ">"; break;
        case 5: header = "grad_mvm_dmma.cuh"; name = "grad_mvm_dmma_kernel<" + sD + ", " + sop + ", false, 0>"; break;
        case 7: header = "gram_mvm_tc5.cuh"; name = "gram_mvm_tc5_kernel<" + sD + ", " + sop + ">"; break;
        default: return fail(CF_ERR_BAD_ARGUMENT, "cf_jit_check: which must be 0..7");
    }
    cf_sop_grad sop_grad;
    const bool want_grad = which == 5;
    if (want_grad) {
        if (!cf::to_sop_grad(lowered, sop_grad) || !lowered.isotropic)
            return fail(CF_ERR_UNSUPPORTED, "cf_jit_check: the program has no isotropic derivative form within the device limits");
    }
    if (which >= 1 && (D < 8 || (which <= 2 && D % 4 != 0))) return fail(CF_ERR_UNSUPPORTED, "cf_jit_check: no tensor-core kernel for D = %d", D);
    std::string text;
    size_t nb = 0;
    const int rc = cfjit::compile_only(sop_val, header, name, text, &nb, want_grad ? &sop_grad : nullptr);
    if (log && loglen > 0) {
        std::snprintf(log, (size_t)loglen, "%s", text.c_str());
    }
    if (rc == 1) return fail(CF_ERR_UNSUPPORTED, "cf_jit_check: NVRTC is not available (%s)", text.c_str());
    if (rc != 0) return fail(CF_ERR_INTERNAL, "cf_jit_check: compilation of %s failed", name.c_str());
    return CF_OK;
}

int cf_jit_stats(int* compiled, int* cache_hits, int* failures, double* compile_seconds) {
    cfjit::State& st = cfjit::state();
    std::lock_guard<std::mutex> lk(st.mu);
    if (compiled) *compiled = st.stats.compiled;
    if (cache_hits) *cache_hits = st.stats.hits + st.stats.disk_hits;  // in-process and on-disk cache
    if (failures) *failures = st.stats.failures;
    if (compile_seconds) *compile_seconds = st.stats.compile_seconds;
    return CF_OK;
}

}  // extern "C"
