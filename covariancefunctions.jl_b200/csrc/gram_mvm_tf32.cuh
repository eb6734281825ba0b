// gram_mvm_tf32.cuh -- K1t: Float32 value MVM  b <- alpha K a + beta b  for padded D >= 8 with the pair distances formed on the
// tensor cores in 3xTF32 split precision (phase A of gram_mm_tf32.cuh, see there for the split, the fragment relabelling and
// the accuracy argument), kernel values evaluated on the C fragments 8 at a time, multiplied by a_j and summed per row.
//
// Replaces mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta) (reference src/gramian.jl:78-87) where
// the scalar Float32 kernel spends 2 D FP32 instructions per pair on the direct differences (EQ, d = 32: 3.7e11 pairs/s, no
// faster than Float64).  Row sums are kept in Float64 across tiles (as in K1), per-tile sums in Float32.
#pragma once
#include "gram_mm_tf32.cuh"

#define CF_MVT_TI 128
#define CF_MVT_TJ 32  // columns per MMA pass
#define CF_MVT_NS 3
// 32-column passes per pipeline stage: at small D one pass is only a few hundred cycles of work per warp, less than the TMA round
// trip, so a stage carries several passes (one barrier per stage instead of one per pass)
template <int D> struct cf_mvt_nsub { static constexpr int v = (D <= 8) ? 8 : (D <= 16 ? 4 : 2); };

template <int D>
struct cf_mvt_layout {
    static constexpr int sx = cf_mmt_sx(D);
    static constexpr int dk = ((D + 7) / 8) * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int xs_bytes = dk * CF_MMT_SK * 4;
    static constexpr int tjs = CF_MVT_TJ * cf_mvt_nsub<D>::v;  // columns per stage
    static constexpr int y_bytes = tjs * sx * 4;
    static constexpr int n_bytes = tjs * 4;
    static constexpr int stage_bytes = ((y_bytes + 2 * n_bytes + 127) / 128) * 128;  // y | yn | a
    static constexpr int total = bar_bytes + xs_bytes + CF_MVT_NS * stage_bytes;
};

template <int KIND, int N>
__device__ __forceinline__ void cf_values_f32_n(const float (&r2)[N], const float (&dt)[N], const cf_atom_val& atom, const cf_sop_val& sop,
                                                float (&kv)[N]) {
    if constexpr (KIND == CF_ATOM_SOP) cf_sop_value_f32_n<N>(r2, dt, sop, kv);
    else if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
        for (int u = 0; u < N; u++) kv[u] = cf_ex2f(r2[u] * atom.f_clog2e);
    } else cf_atom_value_f32_n<N>(r2, dt, atom, kv);  // N-wide: the loop over the integer parameter runs once for the N values
}

// P.X: the points as uploaded (row stride D); P.Y: the padded Float32 copy (row stride sx); P.xn / P.yn: squared norms
template <int D, int KIND>
__global__ void __launch_bounds__(256, 2) gram_mvm_tf32_kernel(const __grid_constant__ cf_mvm_params P) {
    using S = cf_mvt_layout<D>;
    constexpr int SX = S::sx, SK = CF_MMT_SK, NTB = 256, DK = S::dk, TJ = CF_MVT_TJ, TJS = S::tjs, TI = CF_MVT_TI, NS = CF_MVT_NS;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    float* Xs = reinterpret_cast<float*>(smem + S::bar_bytes);
    unsigned char* stages = smem + S::bar_bytes + S::xs_bytes;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const float* __restrict__ Xg = static_cast<const float*>(P.X);
    const float* __restrict__ Yg = static_cast<const float*>(P.Y);
    const float* __restrict__ yng = static_cast<const float*>(P.yn);
    const float* __restrict__ ag = static_cast<const float*>(P.a);

    const int64_t c0 = (int64_t)blockIdx.y * P.cols_per_chunk;
    const int64_t c1 = (c0 + P.cols_per_chunk < P.m) ? c0 + P.cols_per_chunk : P.m;
    const int nfull = P.use_tma ? (int)((c1 - c0) / TJS) : 0;  // full stages streamed by TMA
    const int64_t rem0 = c0 + (int64_t)nfull * TJS;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) cf_mbar_init(&bars[s], 1);
        cf_fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int tile) {
        const int s = tile % NS;
        unsigned char* st = stages + (size_t)s * S::stage_bytes;
        const int64_t j0 = c0 + (int64_t)tile * TJS;
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + 2 * S::n_bytes));
        cf_tma_load_1d(st, Yg + j0 * SX, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, yng + j0, (uint32_t)S::n_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes + S::n_bytes, ag + j0, (uint32_t)S::n_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * TI;
    const int64_t rend = P.row0 + P.nrows;
    for (int q = tid; q < TI * DK; q += NTB) {  // XsT[k][row]: the row tile's points, transposed (rows past the end: clamped)
        const int row = q / DK, c = q - row * DK;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        Xs[c * SK + row] = (c < D) ? Xg[ir * D + c] : 0.f;
    }
    float xnorm[2];
    double tot[2] = {0.0, 0.0};  // this lane's rows: 16 w + 2 g + h
#pragma unroll
    for (int h = 0; h < 2; h++) {
        int64_t i = rbase + 16 * w + 2 * g + h;
        if (i >= rend) i = rend - 1;
        xnorm[h] = static_cast<const float*>(P.xn)[i];
    }
    __syncthreads();

    // columns past the end of a ragged tile carry a_j = 0 (and zero points): no contribution
    auto compute = [&](const float* __restrict__ ys, const float* __restrict__ yns, const float* __restrict__ as) {
        float c[4][4];
#pragma unroll
        for (int cb = 0; cb < 4; cb++)
#pragma unroll
            for (int e = 0; e < 4; e++) c[cb][e] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < DK; k0 += 8) {
            const float2 x0 = *reinterpret_cast<const float2*>(&Xs[(k0 + 2 * t4) * SK + 16 * w + 2 * g]);
            const float2 x1 = *reinterpret_cast<const float2*>(&Xs[(k0 + 2 * t4 + 1) * SK + 16 * w + 2 * g]);
            const uint32_t ah[4] = {__float_as_uint(x0.x), __float_as_uint(x0.y), __float_as_uint(x1.x), __float_as_uint(x1.y)};
            const uint32_t al[4] = {cf_tf32_lo(x0.x), cf_tf32_lo(x0.y), cf_tf32_lo(x1.x), cf_tf32_lo(x1.y)};
#pragma unroll
            for (int cb = 0; cb < 4; cb++) {
                const float2 yv = *reinterpret_cast<const float2*>(&ys[(8 * cb + g) * SX + k0 + 2 * t4]);
                const uint32_t bh[2] = {__float_as_uint(yv.x), __float_as_uint(yv.y)};
                const uint32_t bl[2] = {cf_tf32_lo(yv.x), cf_tf32_lo(yv.y)};
                cf_mma_3xtf32(c[cb], ah, al, bh, bl);
            }
        }
        float yn8[8], a8[8];  // this lane's columns: 8 cb + 2 t4 + e
#pragma unroll
        for (int cb = 0; cb < 4; cb++) {
            const float2 v = *reinterpret_cast<const float2*>(&yns[8 * cb + 2 * t4]);
            const float2 z = *reinterpret_cast<const float2*>(&as[8 * cb + 2 * t4]);
            yn8[2 * cb] = v.x; yn8[2 * cb + 1] = v.y;
            a8[2 * cb] = z.x; a8[2 * cb + 1] = z.y;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float r2[8], dt[8], kv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                dt[u] = c[u >> 1][2 * h + (u & 1)];
                r2[u] = fmaxf(fmaf(-2.f, dt[u], xnorm[h] + yn8[u]), 0.f);
            }
            cf_values_f32_n<KIND, 8>(r2, dt, P.atom, P.sop, kv);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                s0 = fmaf(kv[u], a8[u], s0);
                s1 = fmaf(kv[u + 1], a8[u + 1], s1);
            }
            tot[h] += (double)(s0 + s1);
        }
    };

    for (int t = 0; t < nfull; t++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
#pragma unroll 1
        for (int sub = 0; sub < TJS / TJ; sub++)
            compute(reinterpret_cast<const float*>(st) + sub * TJ * SX, reinterpret_cast<const float*>(st + S::y_bytes) + sub * TJ,
                    reinterpret_cast<const float*>(st + S::y_bytes + S::n_bytes) + sub * TJ);
        __syncthreads();  // every thread is done reading stage s
        if (tid == 0 && t + NS < nfull) issue(t + NS);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ) {  // ragged tail (or everything when a is not TMA-aligned): cooperative loads
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        float* ys = reinterpret_cast<float*>(stages);
        float* yns = reinterpret_cast<float*>(stages + S::y_bytes);
        float* as = reinterpret_cast<float*>(stages + S::y_bytes + S::n_bytes);
        __syncthreads();
        for (int q = tid; q < TJ * SX; q += NTB) ys[q] = (q < cnt * SX) ? Yg[j0 * SX + q] : 0.f;
        for (int q = tid; q < TJ; q += NTB) {
            yns[q] = (q < cnt) ? yng[j0 + q] : 0.f;
            as[q] = (q < cnt) ? ag[j0 + q] : 0.f;
        }
        __syncthreads();
        compute(ys, yns, as);
    }

#pragma unroll
    for (int h = 0; h < 2; h++) {  // the four lanes of a quad hold partial sums of the same rows
        tot[h] += cf_shfl_xor_f64(tot[h], 1);
        tot[h] += cf_shfl_xor_f64(tot[h], 2);
    }
    if (t4 == 0) {
        float* out = static_cast<float*>(P.out);
        const float* yin = static_cast<const float*>(P.yin);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int64_t i = rbase + 16 * w + 2 * g + h;
            if (i >= rend) continue;
            const int64_t o = i - P.row0;
            if (P.direct) {
                double v = P.alpha * tot[h];
                if (P.beta != 0.0) v += P.beta * (double)yin[o];
                out[o] = (float)v;
                for (int p = 0; p < P.peers.n; p++) static_cast<float*>(P.peers.ptr[p])[o] = (float)v;  // NVLink peer stores
            } else {
                reinterpret_cast<double*>(P.out)[(int64_t)blockIdx.y * P.nrows + o] = tot[h];
            }
        }
    }
}

#ifndef __CUDACC_RTC__ // host side
template <int D, int KIND>
cudaError_t cf_mvt_launch(const cf_mvm_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvt_layout<D>;
    auto kern = gram_mvm_tf32_kernel<D, KIND>;
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

template <int D, bool OK = (D >= 8)>
struct cf_mvt_entry {
    static constexpr cf_mvm_launch_fn fn[4] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr cf_mvm_config cfg = {CF_MVT_TI, CF_MVT_TJ, 0, 2};
};
template <int D>
struct cf_mvt_entry<D, true> {
    static constexpr cf_mvm_launch_fn fn[4] = {&cf_mvt_launch<D, CF_ATOM_EQ>, &cf_mvt_launch<D, CF_ATOM_MATERN>,
                                               &cf_mvt_launch<D, CF_ATOM_RQ_INT>, &cf_mvt_launch<D, CF_ATOM_SOP>};
    static constexpr cf_mvm_config cfg = {CF_MVT_TI, cf_mvt_layout<D>::tjs, cf_mvt_layout<D>::total, 2};
};
#endif // !__CUDACC_RTC__
