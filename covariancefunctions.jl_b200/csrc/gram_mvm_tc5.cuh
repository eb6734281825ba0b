// gram_mvm_tc5.cuh -- K1u: Float32 value MVM  b <- alpha K a + beta b  for padded D >= 8 with the pair dot products on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory), 3xTF32 split precision.
//
// Replaces mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta) (reference src/gramian.jl:78-87) for well-scaled
// points, like gram_mvm_tf32.cuh (K1t), whose legacy mma.sync distance GEMM it supersedes: there every warp loads fragments, splits
// them and issues 3 D / 8 mma.sync per 16 x 8 entries before it can evaluate anything (EQ, d = 32: 8.7e11 pairs/s); here one thread
// issues 3 D / 8 asynchronous MMAs per 128 x 64 tile -- A operand: the CTA's row tile, written ONCE to tensor memory; B operand: the
// canonical shared-memory images that gram_mm_tc5.cuh (K4u) already keeps per handle, delivered by TMA -- and the sixteen evaluation
// warps only evaluate: tcgen05.ld of 32 dot products,
// r2 = |x|^2 + |y|^2 - 2 x.y, the kernel, the weighted sum -- in packed FP32 instructions over column pairs (gram_mvm_f32p.cuh).
// For EQ that is 3 FFMA2 + 2 MUFU.EX2 + 2 LDS.64 per two pairs: the MUFU pipe (one ex2 per pair, 16 per clock and SM) is the bound.
//
//   per row tile of 128 rows and column chunk (one CTA), per column tile of TJ = 64 points:
//   TMA producer        Yhi | Ylo images, |y|^2, a  -> stage s                              4 bulk copies, full[s]
//   MMA issuer          Dot (128 x 64) = Xlo Yhi^T + Xhi Ylo^T + Xhi Yhi^T  -> TMEM dot[t % 4]    dotfull[t % 4]
//   evaluation group t & 1 (8 warps: TMEM lane quarter x column half)  tcgen05.ld -> dotfree, kernel, acc2 += k2 a2, every lane -> empty[s]
// Row sums: Float32 within a tile (16 terms per accumulator half), Float64 across tiles, as in K1 / K1t.
#pragma once
#include "gram_mm_tc5.cuh"
#include "gram_mvm_f32p.cuh"

#define CF_MVU_TI 128
#define CF_MVU_TJ 64
#define CF_MVU_NS 4
#define CF_MVU_NB 4                        // TMEM dot buffers (64 columns each)
#define CF_MVU_EW 16                       // evaluation warps: two groups of 8 that take alternate column tiles
#define CF_MVU_THREADS (32 * CF_MVU_EW + 64)   // + TMA producer warp, MMA issuer warp
// EQ: which of a thread's 16 column pairs per tile take their two exponentials from the polynomial cf_ex2_poly2 (FMA pipe) instead of
// MUFU.EX2.  With every exponential on the MUFU pipe that pipe is 84 % busy and the FMA pipe ~37 % (profiles/r2_ncu_x2_f32.md), and per
// column pair the MUFU path costs 16 MUFU cycles + 6 FMA-pipe cycles against 22 FMA-pipe cycles for the polynomial, so on paper the pipes
// balance at 5 of 16 pairs.  MEASURED (n = 131072, bench_aux/micro/mvu_variants.sh): d = 8: 4.73 ms with none, 4.64 with 1 pair, 4.53
// with 3, 4.79 with 5; d = 32: 5.10 / 5.29 / 5.39 / 5.42 ms -- the extra ~10 issue slots per pair cost more than the MUFU cycles they
// free once the tile hand-over (tcgen05.ld, mbarriers) shares the issue port.  Off by default; kept for kernels with a heavier MUFU load.
// (The exponent c log2(e) r2 stays above -126 here: these kernels run only on points that passed the scale check, capi.cu set_norm_flags.)
#ifndef CF_MVU_POLY_MASK
#define CF_MVU_POLY_MASK 0x0u  // e.g. 0x1084u: pairs 2, 7, 12
#endif

template <int D>
struct cf_mvu_layout {
    static constexpr int dk = ((D + 7) / 8) * 8;
    static constexpr int y_bytes = CF_MVU_TJ * dk * 4;            // each of hi, lo
    static constexpr int n_bytes = CF_MVU_TJ * 4;                 // |y|^2, and the weights
    static constexpr int stage_bytes = 2 * y_bytes + 2 * n_bytes;
    static constexpr int bar_bytes = 256;
    static constexpr int total = bar_bytes + CF_MVU_NS * stage_bytes + 1024;  // + alignment slack (the row tile lives in tensor memory)
};

struct cf_mvu_params {
    cf_mvm_params mv;      // X (rows as uploaded, stride D), xn, out / yin, row0, nrows, m, cols_per_chunk, alpha, beta, direct, atom, sop, peers
    const float* yhi;      // canonical column-point images (cf_canon_points_kernel)
    const float* ylo;
    const float* ynpad;    // squared norms of the columns, zero padded to a multiple of TJ
    const float* apad;     // weights, readable (and zero) up to a multiple of TJ
};

// weights copied to a 16-byte aligned buffer that is zero beyond m (tiles of the last chunk read up to mpad)
static __global__ void cf_pad_vec_f32_kernel(const float* __restrict__ a, int64_t m, int64_t mpad, float* __restrict__ out) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < mpad; q += (int64_t)gridDim.x * blockDim.x) out[q] = (q < m) ? a[q] : 0.f;
}

template <int D, int KIND>
__global__ void __launch_bounds__(CF_MVU_THREADS, 1) gram_mvm_tc5_kernel(const __grid_constant__ cf_mvu_params PP) {
    using S = cf_mvu_layout<D>;
    constexpr int DK = S::dk, TI = CF_MVU_TI, TJ = CF_MVU_TJ, NS = CF_MVU_NS, NB = CF_MVU_NB, GW = CF_MVU_EW / 2;
    const cf_mvm_params& P = PP.mv;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (cf_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t *full = bars, *empty = bars + NS, *dotfull = bars + 2 * NS, *dotfree = dotfull + NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 2 * NB);
    unsigned char* stages = smem + S::bar_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // this CTA's column tiles
    const int64_t tile0 = ((int64_t)blockIdx.y * P.cols_per_chunk) / TJ;
    const int64_t c1 = ((int64_t)(blockIdx.y + 1) * P.cols_per_chunk < P.m) ? (int64_t)(blockIdx.y + 1) * P.cols_per_chunk : P.m;
    const int ntiles = (int)((c1 + TJ - 1) / TJ - tile0);

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { cf_mbar_init(&full[s], 1); cf_mbar_init(&empty[s], 32 * GW); }  // every thread of the tile's evaluation group (see below)
        for (int b = 0; b < NB; b++) { cf_mbar_init(&dotfull[b], 1); cf_mbar_init(&dotfree[b], GW); }
        cf_fence_barrier_init();
    }
    if (warp == CF_MVU_EW + 1) {  // 512 TMEM columns (4 x 64 dot buffers + 2 DK for the row tile; the allocation must be a power of two)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(cf_smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    cf_tc_fence_before();
    __syncthreads();
    cf_tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_xhi = tmem + 64 * NB, tm_xlo = tm_xhi + DK;
    const int64_t rbase = P.row0 + (int64_t)blockIdx.x * TI;
    const int64_t rend = P.row0 + P.nrows;
    // The row tile is the A operand of every MMA of this CTA: it is written ONCE to tensor memory (lane = tile row, column = coordinate;
    // hi = the raw word, lo = v - trunc_tf32(v)), so that an MMA reads only its 64 x 8 B operand from shared memory.  (With A in shared
    // memory each M128 N64 K8 instruction re-read 4 KB of X: EQ, d = 32 ran at 2.2e12 pairs/s, bound by that operand traffic.)
    if (warp < 4) {
        const float* __restrict__ Xg = static_cast<const float*>(P.X);
        int64_t ir = rbase + 32 * warp + lane;
        if (ir >= rend) ir = rend - 1;  // rows past the end: clamped, never stored
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
#pragma unroll
        for (int k0 = 0; k0 < DK; k0 += 8) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const float v = (k0 + e < D) ? Xg[ir * D + k0 + e] : 0.f;
                h[e] = __float_as_uint(v);
                l[e] = cf_tf32_lo(v);
            }
            cf_tmem_st8(tm_xhi + lane_base + k0, h);
            cf_tmem_st8(tm_xlo + lane_base + k0, l);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    cf_tc_fence_before();
    __syncthreads();
    cf_tc_fence_after();

    if (warp == CF_MVU_EW) {
        // ---- TMA producer --------------------------------------------------------------------------------------------------------------
        if (lane == 0) {
            for (int t = 0; t < ntiles; t++) {
                const int s = t % NS;
                if (t >= NS) cf_mbar_wait(&empty[s], (uint32_t)(((t / NS) - 1) & 1));
                unsigned char* st = stages + (size_t)s * S::stage_bytes;
                const int64_t gt = tile0 + t;
                cf_mbar_expect_tx(&full[s], (uint32_t)S::stage_bytes);
                cf_tma_load_1d(st, PP.yhi + gt * TJ * DK, (uint32_t)S::y_bytes, &full[s]);
                cf_tma_load_1d(st + S::y_bytes, PP.ylo + gt * TJ * DK, (uint32_t)S::y_bytes, &full[s]);
                cf_tma_load_1d(st + 2 * S::y_bytes, PP.ynpad + gt * TJ, (uint32_t)S::n_bytes, &full[s]);
                cf_tma_load_1d(st + 2 * S::y_bytes + S::n_bytes, PP.apad + gt * TJ, (uint32_t)S::n_bytes, &full[s]);
            }
        }
    } else if (warp == CF_MVU_EW + 1) {
        // ---- MMA issuer: the dot products of tile t as soon as its stage has landed and TMEM dot[t % NB] has been read ---------------------------
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // D fp32, A / B tf32 K-major, N 64, M 128
        constexpr uint32_t LBO_Y = (TJ / 8) * 128;
        for (int t = 0; t < ntiles; t++) {
            const int s = t % NS, b = t % NB;
            cf_mbar_wait(&full[s], (uint32_t)((t / NS) & 1));
            if (t >= NB) cf_mbar_wait(&dotfree[b], (uint32_t)(((t / NB) - 1) & 1));
            cf_tc_fence_after();
            const uint64_t dyh = cf_umma_desc(cf_smem_u32(stages + (size_t)s * S::stage_bytes), LBO_Y, 128), dyl = dyh + (S::y_bytes >> 4);
            if (cf_elect_one()) {
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {  // small terms first: lo.hi, hi.lo, hi.hi
                    const uint32_t xa = pr == 0 ? tm_xlo : tm_xhi;
                    const uint64_t yb = pr == 1 ? dyl : dyh;
#pragma unroll
                    for (int ks = 0; ks < DK / 8; ks++)
                        cf_umma_tf32_ta(tmem + 64 * b, xa + 8 * ks, yb + (uint64_t)((ks * 2 * LBO_Y) >> 4), idesc, (pr > 0 || ks > 0) ? 1u : 0u);
                }
                cf_umma_commit(&dotfull[b]);  // (also what frees the stage's point images: the evaluation warps arrive on empty[s] only after they
                                              // have seen dotfull[b], i.e. after every MMA that reads the stage has completed)
            }
            __syncwarp();
        }
    } else {
        // ---- evaluation warps -------------------------------------------------------------------------------------------------------------
        const int grp = warp / GW, wg = warp % GW;
        const int q4 = wg & 3, cq = wg >> 2;  // TMEM lane quarter (tile rows 32 q4 ..), column half (32 cq ..)
        const int row = 32 * q4 + lane;
        int64_t ir = rbase + row;
        if (ir >= rend) ir = rend - 1;
        const float xnorm = static_cast<const float*>(P.xn)[ir];
        const uint32_t lane_base = (uint32_t)(32 * q4) << 16;
        // EQ: the exponent of ex2 directly, c log2(e) (|x|^2 + |y|^2) - 2 c log2(e) x.y
        const float cl = P.atom.f_clog2e;
        const uint64_t cl2 = cf_pk2(cl, cl), m2cl2 = cf_pk2(-2.f * cl, -2.f * cl), cxn2 = cf_pk2(cl * xnorm, cl * xnorm);
        const uint64_t xn2 = cf_pk2(xnorm, xnorm), m2 = cf_pk2(-2.f, -2.f);
        double tot = 0.0;
        for (int t = grp; t < ntiles; t += 2) {
            const int s = t % NS, b = t % NB;
            cf_mbar_wait(&full[s], (uint32_t)((t / NS) & 1));  // |y|^2 and the weights of the tile (landed long ago: the MMA waited for it too)
            cf_mbar_wait(&dotfull[b], (uint32_t)((t / NB) & 1));
            cf_tc_fence_after();
            uint32_t dv[32];
            cf_tmem_ld32(tmem + 64 * b + lane_base + 32 * cq, dv);
            cf_tc_fence_before();
            __syncwarp();
            if (lane == 0) cf_mbar_arrive(&dotfree[b]);
            const unsigned char* st = stages + (size_t)s * S::stage_bytes + 2 * S::y_bytes;
            const uint64_t* yn2 = reinterpret_cast<const uint64_t*>(st) + 16 * cq;
            const uint64_t* a2 = reinterpret_cast<const uint64_t*>(st + S::n_bytes) + 16 * cq;
            uint64_t acc = 0ull;
            if constexpr (KIND == CF_ATOM_EQ) {
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const uint64_t dot2 = cf_pk2(__uint_as_float(dv[2 * u]), __uint_as_float(dv[2 * u + 1]));
                    const uint64_t arg2 = cf_fma2(dot2, m2cl2, cf_fma2(yn2[u], cl2, cxn2));
                    if ((CF_MVU_POLY_MASK >> u) & 1) {  // this column pair's exponentials on the FMA pipe (see CF_MVU_POLY_MASK)
                        acc = cf_fma2(cf_ex2_poly2(arg2), a2[u], acc);
                    } else {
                        float lo, hi;
                        cf_upk2(arg2, lo, hi);
                        acc = cf_fma2(cf_pk2(cf_ex2f(lo), cf_ex2f(hi)), a2[u], acc);
                    }
                }
            } else if constexpr (KIND == CF_ATOM_SOP) {
#pragma unroll
                for (int g8 = 0; g8 < 4; g8++) {
                    float r2[8], dt[8], kv[8];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        dt[2 * u] = __uint_as_float(dv[8 * g8 + 2 * u]);
                        dt[2 * u + 1] = __uint_as_float(dv[8 * g8 + 2 * u + 1]);
                        const uint64_t dot2 = cf_pk2(dt[2 * u], dt[2 * u + 1]);
                        const uint64_t s2 = cf_fma2(dot2, m2, cf_add2(yn2[4 * g8 + u], xn2));
                        cf_upk2(s2, r2[2 * u], r2[2 * u + 1]);
                        r2[2 * u] = fmaxf(r2[2 * u], 0.f);
                        r2[2 * u + 1] = fmaxf(r2[2 * u + 1], 0.f);
                    }
                    cf_sop_value_f32_n<8>(r2, dt, P.sop, kv);
#pragma unroll
                    for (int u = 0; u < 4; u++) acc = cf_fma2(cf_pk2(kv[2 * u], kv[2 * u + 1]), a2[4 * g8 + u], acc);
                }
            } else {  // single MaternP / RQ atom: packed evaluation, 8 column pairs at a time (points are well scaled here: no clamp of sqrt(r2))
#pragma unroll
                for (int g8 = 0; g8 < 2; g8++) {
                    uint64_t r2[8], kv[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const uint64_t dot2 = cf_pk2(__uint_as_float(dv[16 * g8 + 2 * u]), __uint_as_float(dv[16 * g8 + 2 * u + 1]));
                        r2[u] = cf_fma2(dot2, m2, cf_add2(yn2[8 * g8 + u], xn2));
                    }
                    cf_atom_value_f32x2_n<KIND, 8, false, false>(r2, P.atom, kv);
#pragma unroll
                    for (int u = 0; u < 8; u++) acc = cf_fma2(kv[u], a2[8 * g8 + u], acc);
                }
            }
            // done with the stage's |y|^2 and weights.  Every lane arrives itself (one warp-wide instruction): lane 0 arriving after
            // __syncwarp orders the same accesses, but compute-sanitizer's racecheck only follows a thread's OWN arrival (profiles/r2_sanitizer.txt)
            cf_mbar_arrive(&empty[s]);
            float lo, hi;
            cf_upk2(acc, lo, hi);
            tot += (double)(lo + hi);
        }
        // four partial sums per row (two groups x two column halves), combined in a fixed order
        double* xch = reinterpret_cast<double*>(stages);  // every stage has been consumed once all evaluation warps are here
        asm volatile("bar.sync 1, %0;" ::"r"(CF_MVU_EW * 32) : "memory");
        if (grp + cq > 0) xch[(2 * grp + cq) * TI + row] = tot;
        asm volatile("bar.sync 1, %0;" ::"r"(CF_MVU_EW * 32) : "memory");
        const int64_t i = rbase + row;
        if (grp + cq == 0 && i < rend) {
            tot = (tot + xch[1 * TI + row]) + (xch[2 * TI + row] + xch[3 * TI + row]);
            const int64_t o = i - P.row0;
            if (P.direct) {
                float* out = static_cast<float*>(P.out);
                double v = P.alpha * tot;
                if (P.beta != 0.0) v += P.beta * (double)static_cast<const float*>(P.yin)[o];
                out[o] = (float)v;
                for (int p = 0; p < P.peers.n; p++) static_cast<float*>(P.peers.ptr[p])[o] = (float)v;  // NVLink peer stores
            } else {
                reinterpret_cast<double*>(P.out)[(int64_t)blockIdx.y * P.nrows + o] = tot;
            }
        }
    }
    cf_tc_fence_before();
    __syncthreads();
    if (warp == CF_MVU_EW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

#ifndef __CUDACC_RTC__
typedef cudaError_t (*cf_mvu_launch_fn)(const cf_mvu_params& P, dim3 grid, cudaStream_t stream);
template <int D, int KIND>
cudaError_t cf_mvu_launch(const cf_mvu_params& P, dim3 grid, cudaStream_t stream) {
    using S = cf_mvu_layout<D>;
    auto kern = gram_mvm_tc5_kernel<D, KIND>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<grid, CF_MVU_THREADS, S::total, stream>>>(P);
    return cudaGetLastError();
}
template <int D, bool OK = (D >= 8)>
struct cf_mvu_entry {
    static constexpr cf_mvu_launch_fn fn[4] = {nullptr, nullptr, nullptr, nullptr};
    static constexpr cf_mvm_config cfg = {CF_MVU_TI, CF_MVU_TJ, 0, 1};
};
template <int D>
struct cf_mvu_entry<D, true> {
    static constexpr cf_mvu_launch_fn fn[4] = {&cf_mvu_launch<D, CF_ATOM_EQ>, &cf_mvu_launch<D, CF_ATOM_MATERN>, &cf_mvu_launch<D, CF_ATOM_RQ_INT>,
                                               &cf_mvu_launch<D, CF_ATOM_SOP>};
    static constexpr cf_mvm_config cfg = {CF_MVU_TI, CF_MVU_TJ, cf_mvu_layout<D>::total, 1};
};
#endif
