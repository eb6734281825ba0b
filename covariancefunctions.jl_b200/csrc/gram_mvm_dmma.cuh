// gram_mvm_dmma.cuh -- K1d: Float64 value MVM  b <- alpha K a + beta b  for high-dimensional points (padded D >= 12, multiple
// of 4, well-scaled) with the pair distances formed on the FP64 tensor cores.
//
// Replaces mul!(y::AbstractVector, G::Gramian, x::AbstractVector, alpha, beta) (reference src/gramian.jl:78-87) where the
// scalar kernel K1 is bound by the 2 D FP64 instructions of the direct differences (D subtractions + D FMAs per pair, at the
// FP64 pipe peak for D = 32: bench x2).  With r^2 = |x|^2 + |y|^2 - 2 x.y the D-dependent work is ONE FMA per coordinate and
// it is a GEMM, Dot (128 x 32) = Xs (128 x D) . Ys^T (D x 32), which DMMA m8n8k4 performs from two operand doubles per lane
// per 256 FMAs (see gram_mm_dmma.cuh for the operand-traffic argument and the fragment layout).  The kernel value is then
// evaluated on the C fragments, 8 entries at a time, multiplied by a_j and summed per row; the four lanes that share a row
// are reduced once at the end.  The host only selects this kernel under the same scale check as the multi-RHS kernel
// (capi.cu: use_norms) and never for programs with an exp(-sqrt(r2)) atom.
#pragma once
#include "gram_mm_dmma.cuh"

#define CF_MVD_TI 128
#define CF_MVD_TJ 32
#define CF_MVD_NS 3

template <int D>
struct cf_mvd_smem {
    static constexpr int sx = cf_mmd_smem<D>::sx;
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int xs_bytes = CF_MVD_TI * sx * 8;
    static constexpr int y_bytes = CF_MVD_TJ * sx * 8;
    static constexpr int n_bytes = CF_MVD_TJ * 8;
    static constexpr int a_bytes = CF_MVD_TJ * 8;
    static constexpr int stage_bytes = ((y_bytes + n_bytes + a_bytes + 127) / 128) * 128;
    static constexpr int total = tbl_bytes + bar_bytes + xs_bytes + CF_MVD_NS * stage_bytes;
};

// kernel values of N pairs from (r2, x.y): compile-time kind or the generic program
template <int KIND, int N>
__device__ __forceinline__ void cf_values_n(const double (&r2)[N], const double (&dt)[N], const cf_atom_val& atom, const cf_sop_val& sop,
                                            cf_tbl_t tbl_lane, double (&kv)[N]) {
    if constexpr (KIND == CF_ATOM_SOP) cf_sop_value_n<N>(r2, dt, sop, tbl_lane, kv);
    else if constexpr (KIND == CF_ATOM_MATERN) cf_atom_matern_n<N>(r2, atom, tbl_lane, kv);
    else if constexpr (KIND == CF_ATOM_RQ_INT) cf_atom_rq_int_n<N>(r2, atom, kv);
    else {
#pragma unroll
        for (int u = 0; u < N; u++) kv[u] = cf_atom_value<KIND>(r2[u], dt[u], atom, tbl_lane);
    }
}

template <int D, int KIND>
__global__ void __launch_bounds__(256, (KIND == CF_ATOM_SOP) ? 1 : 2) gram_mvm_dmma_kernel(const __grid_constant__ cf_mvm_params P) {
    using S = cf_mvd_smem<D>;
    constexpr int SX = S::sx, NTB = 256, TJ = CF_MVD_TJ, TI = CF_MVD_TI, NS = CF_MVD_NS;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    double* Xs = reinterpret_cast<double*>(smem + S::tbl_bytes + S::bar_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes + S::xs_bytes;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t4 = lane & 3;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const double* __restrict__ Xg = static_cast<const double*>(P.X);
    const double* __restrict__ Yg = static_cast<const double*>(P.Y);
    const double* __restrict__ yng = static_cast<const double*>(P.yn);
    const double* __restrict__ ag = static_cast<const double*>(P.a);

    const int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    const int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    const int nfull = P.use_tma ? (int)((c1 - c0) / TJ) : 0;
    const int64_t rem0 = c0 + (int64_t)nfull * TJ;

    cf_fill_exp_table(tbl, P.exp2_tbl, tid, NTB);
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
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + S::n_bytes + S::a_bytes));
        cf_tma_load_1d(st, Yg + j0 * SX, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, yng + j0, (uint32_t)S::n_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes + S::n_bytes, ag + j0, (uint32_t)S::a_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * TI;
    const int64_t rend = P.row0 + P.nrows;
    for (int q = tid; q < TI * SX; q += NTB) {  // the row tile's points (rows past the end: clamped, never stored)
        const int row = q / SX;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        Xs[q] = Xg[ir * SX + (q - row * SX)];
    }
    double xnorm[2], tot[2] = {0.0, 0.0};  // this lane's rows: 16 w + 8 rb + g
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        int64_t i = rbase + 16 * w + 8 * rb + g;
        if (i >= rend) i = rend - 1;
        xnorm[rb] = static_cast<const double*>(P.xn)[i];
    }
    __syncthreads();

    // columns past the end of a ragged tile carry a_j = 0 (and zero points): no contribution
    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ yns, const double* __restrict__ as) {
        double c[2][4][2];
#pragma unroll
        for (int rb = 0; rb < 2; rb++)
#pragma unroll
            for (int cb = 0; cb < 4; cb++) c[rb][cb][0] = c[rb][cb][1] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < D; k0 += 4) {
            double a[2], b[4];
#pragma unroll
            for (int rb = 0; rb < 2; rb++) a[rb] = Xs[(16 * w + 8 * rb + g) * SX + k0 + t4];
#pragma unroll
            for (int cb = 0; cb < 4; cb++) b[cb] = ys[(8 * cb + g) * SX + k0 + t4];
#pragma unroll
            for (int rb = 0; rb < 2; rb++)
#pragma unroll
                for (int cb = 0; cb < 4; cb++) cf_dmma884(c[rb][cb], a[rb], b[cb]);
        }
        double yn8[8], a8[8];  // this lane's columns: 8 cb + 2 t4 + e
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {
            const double2 v = *reinterpret_cast<const double2*>(&yns[8 * cb + 2 * t4]);
            const double2 z = *reinterpret_cast<const double2*>(&as[8 * cb + 2 * t4]);
            yn8[2 * cb] = v.x; yn8[2 * cb + 1] = v.y;
            a8[2 * cb] = z.x; a8[2 * cb + 1] = z.y;
        }
#pragma unroll
        for (int rb = 0; rb < 2; rb++) {
            double r2[8], dt[8], kv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                dt[u] = c[rb][u >> 1][u & 1];
                const double v = fma(-2.0, dt[u], xnorm[rb] + yn8[u]);
                r2[u] = (__double2hiint(v) < 0) ? 0.0 : v;  // rounding can leave a tiny negative value
            }
            cf_values_n<KIND, 8>(r2, dt, P.atom, P.sop, tbl_lane, kv);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                s0 = fma(kv[u], a8[u], s0);
                s1 = fma(kv[u + 1], a8[u + 1], s1);
            }
            tot[rb] += s0 + s1;
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                reinterpret_cast<const double*>(st + S::y_bytes + S::n_bytes));
        __syncthreads();  // every thread is done reading stage s
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {  // ragged tail (or everything when a is not TMA-aligned): cooperative loads
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* yns = reinterpret_cast<double*>(stages + S::y_bytes);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes + S::n_bytes);
        __syncthreads();
        for (int q = tid; q < TJ * SX; q += NTB) ys[q] = (q < cnt * SX) ? Yg[j0 * SX + q] : 0.0;
        for (int q = tid; q < TJ; q += NTB) {
            yns[q] = (q < cnt) ? yng[j0 + q] : 0.0;
            as[q] = (q < cnt) ? ag[j0 + q] : 0.0;
        }
        __syncthreads();
        compute(ys, yns, as);
    }

    // the four lanes t4 = 0..3 of a quad hold the partial sums of the same row
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        tot[rb] += cf_shfl_xor_f64(tot[rb], 1);
        tot[rb] += cf_shfl_xor_f64(tot[rb], 2);
    }
    if (t4 == 0) {
        double* out = static_cast<double*>(P.out);
        const double* yin = static_cast<const double*>(P.yin);
#pragma unroll
        for (int rb = 0; rb < 2; rb++) {
            const int64_t i = rbase + 16 * w + 8 * rb + g;
            if (i >= rend) continue;
            const int64_t o = i - P.row0;
            if (P.direct) {
                double v = P.alpha * tot[rb];
                if (P.beta != 0.0) v += P.beta * yin[o];
                out[o] = v;
                for (int p = 0; p < P.peers.n; p++) static_cast<double*>(P.peers.ptr[p])[o] = v;  // NVLink peer stores
            } else {
                out[(int64_t)blockIdx.y * P.nrows + o] = tot[rb];
            }
        }
    }
}

#ifndef __CUDACC_RTC__ // host side: not part of run-time specialised builds
template <int D, int KIND>
cudaError_t cf_mvd_launch(const cf_mvm_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvd_smem<D>;
    auto kern = gram_mvm_dmma_kernel<D, KIND>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<grid, 256, S::total, stream>>>(P);
    return cudaGetLastError();
}

// registry hook: available for padded dimensions that are multiples of 4 and at least 12
template <int D, bool OK = (D >= 8 && D % 4 == 0)>
struct cf_mvd_entry {
    static constexpr cf_mvm_launch_fn fn[4] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr cf_mvm_config cfg = {CF_MVD_TI, CF_MVD_TJ, 0, 1};
};
template <int D>
struct cf_mvd_entry<D, true> {
    static constexpr cf_mvm_launch_fn fn[4] = {&cf_mvd_launch<D, CF_ATOM_EQ>, &cf_mvd_launch<D, CF_ATOM_MATERN>,
                                               &cf_mvd_launch<D, CF_ATOM_RQ_INT>, &cf_mvd_launch<D, CF_ATOM_SOP>};
    static constexpr cf_mvm_config cfg = {CF_MVD_TI, CF_MVD_TJ, cf_mvd_smem<D>::total, 2};
};
#endif // !__CUDACC_RTC__
