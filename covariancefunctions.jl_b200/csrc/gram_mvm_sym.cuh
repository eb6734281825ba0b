// gram_mvm_sym.cuh -- symmetric variant of the value MVM for y === x (K = K^T), Float64, DETERMINISTIC.
//
// Every unordered pair {i, j} in different row blocks is evaluated ONCE and used twice:
//     b_i += k(x_i, x_j) a_j      (row sums: registers, exactly as in gram_mvm_kernel / gram_mvm_eq_kernel)
//     b_j += k(x_i, x_j) a_i      (column sums: reduced across the CTA's rows)
// which halves the exp / distance work of the symmetric Gramians (BASELINE configs 1, 2, 5).  The reference evaluates all n*m
// entries (src/gramian.jl:78-87); the result differs from it only by summation order (<= 1e-13), like every other kernel here.
//
// No floating-point atomics: every partial sum has exactly ONE writer and the partials are added in a fixed order, so the
// result is bit-reproducible run to run (tests/test_gpu_sym.py).
//   * work item = (row tile I of NT*R rows, column chunk c of the columns beyond the tile's own block); the host builds the list;
//   * row sums of item (I, c) go to rowpart[c][rows of I];
//   * column sums: 8 columns at a time a 5-stage transpose-reduce over the warp (9 shuffle exchanges) leaves one column total per
//     lane quad; the warps' totals meet in shared memory (fixed warp order) and the CTA writes ONE value per column to
//     colpart[I][j] (triangular layout: tile I only has columns j >= (I + 1) NT R);
//   * gram_sym_combine adds, per element, the diagonal-block product, the row partials (chunk order) and the column partials
//     (tile order), then applies alpha / beta.
// The column partials take n^2 / (2 NT R) doubles (4.3 GB at n = 2^20, written and read once per product: ~1.5 ms of HBM time
// against ~0.5 s of arithmetic); the host only selects this variant while that stays below a quarter of the device memory.
//
// EQF = 1 evaluates the EQ kernel in the scaled domain (gram_mvm_eq.cuh: exponent from |x|^2 + |y|^2 - 2 x.y, 12 FP64
// instructions per pair); the row's leftover factor exp(lo_i ln2 / 256) multiplies the row sum in the epilogue and the row's weight
// on the column side.  EQF = 2 evaluates a single MaternP atom (p >= 1) from the norm expansion with cf_matern_fast_n (gram_mvm_eq.cuh,
// FAST = 2).  EQF = 0 evaluates any kernel kind through cf_rows_value (direct differences).
#pragma once
#include "gram_mvm_eq.cuh"

struct cf_sym_item {
    int64_t col0, col1;  // column range (col0 is a multiple of TJ)
    int64_t colpart_off; // offset of colpart[I][col0] in the triangular column-partial buffer
    int32_t row_tile;
    int32_t chunk;       // ordinal of this chunk within its row tile: row sums go to rowpart[chunk][.]
};

struct cf_sym_params {
    const double* X;       // padded AoS, stride D (rows and columns: the same point set)
    const double* xn;      // squared norms (EQF)
    const double* a;       // weights, length n
    double* rowpart;       // [max chunks][n] row-side partial sums
    double* colpart;       // triangular column-side partial sums
    const double* exp2_tbl;
    const cf_sym_item* items;
    int64_t n;
    int use_tma;
    cf_atom_val atom;
    cf_sop_val sop;
    double eqc[4];         // EQF: polynomial constants (see gram_mvm_eq.cuh)
};

template <int D, int TJ, int NS, int NW, int EQF>
struct cf_sym_smem {
    static constexpr int tbl_bytes = CF_EXP_TBL_DOUBLES * 8;
    static constexpr int bar_bytes = 128;
    static constexpr int y_bytes = TJ * D * 8;
    static constexpr int v_bytes = TJ * 8;
    static constexpr int stage_bytes = ((y_bytes + (EQF ? 2 : 1) * v_bytes + 127) / 128) * 128;
    static constexpr int col_bytes = 2 * NW * TJ * 8;  // per-warp column totals, double buffered by tile parity
    static constexpr int total = tbl_bytes + bar_bytes + NS * stage_bytes + col_bytes;
};

template <int D, int KIND, int EQF, int R, int NT, int TJ, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) gram_mvm_sym_kernel(const __grid_constant__ cf_sym_params P) {
    constexpr int NW = NT / 32;
    using S = cf_sym_smem<D, TJ, NS, NW, EQF>;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tbl = reinterpret_cast<double*>(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::tbl_bytes);
    unsigned char* stages = smem + S::tbl_bytes + S::bar_bytes;
    double* colbuf = reinterpret_cast<double*>(stages + NS * S::stage_bytes);  // [2][NW][TJ]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cf_tbl_t tbl_lane = cf_tbl_lane(tbl, tid);
    const cf_sym_item it = P.items[blockIdx.x];
    const int64_t c0 = it.col0, c1 = it.col1;
    const int nfull = P.use_tma ? (int)((c1 - c0) / TJ) : 0;
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
        cf_mbar_expect_tx(&bars[s], (uint32_t)(S::y_bytes + (EQF ? 2 : 1) * S::v_bytes));
        cf_tma_load_1d(st, P.X + j0 * D, (uint32_t)S::y_bytes, &bars[s]);
        cf_tma_load_1d(st + S::y_bytes, P.a + j0, (uint32_t)S::v_bytes, &bars[s]);
        if (EQF) cf_tma_load_1d(st + S::y_bytes + S::v_bytes, P.xn + j0, (uint32_t)S::v_bytes, &bars[s]);
    };
    if (tid == 0)
        for (int t = 0; t < NS && t < nfull; t++) issue(t);

    // this thread's rows (rows past the end: clamped point, zero weight, never stored).  EQF: xs = -2 x, mrow as in gram_mvm_eq.cuh,
    // at = a_i times the row's leftover factor (the column side sees the complete kernel value)
    double x[R][D], mrow[R], at[R], ez[R], tot[R];
    const int64_t rbase = (int64_t)it.row_tile * (NT * R);
    const double c1s = P.atom.e.c1;
#pragma unroll
    for (int r = 0; r < R; r++) {
        int64_t i = rbase + (int64_t)r * NT + tid;
        const bool ok = i < P.n;
        if (!ok) i = P.n - 1;
        const double ai = ok ? P.a[i] : 0.0;
        if (EQF == 2) {
#pragma unroll
            for (int c = 0; c < D; c++) x[r][c] = -2.0 * P.X[i * D + c];
            mrow[r] = P.xn[i]; ez[r] = 1.0; at[r] = ai;
        } else if (EQF == 1) {
#pragma unroll
            for (int c = 0; c < D; c++) x[r][c] = -2.0 * P.X[i * D + c];
            const double nx = P.xn[i];
            const double w = c1s * nx;
            mrow[r] = rint(w) + CF_MAGIC;
            const double lo = (w - rint(w)) + fma(c1s, nx, -w);
            const double z = lo * (0.693147180559945309417232121458 / 256.0);
            ez[r] = 1.0 + z * (1.0 + z * (0.5 + z * (1.0 / 6.0 + z * (1.0 / 24.0 + z * (1.0 / 120.0)))));
            at[r] = ai * ez[r];
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) x[r][c] = P.X[i * D + c];
            mrow[r] = 0.0; ez[r] = 1.0; at[r] = ai;
        }
        tot[r] = 0.0;
    }
    const double g0 = P.eqc[0], g1 = P.eqc[1], g2 = P.eqc[2];
    constexpr double g3 = 0x1.3b2ab00000000p-39;
    double magic = CF_MAGIC, mtop = (EQF == 2) ? P.atom.mat[P.atom.p] : 0.0;  // EQF = 2: pinned in registers (see cf_matern_fast_n)
    if (EQF == 2) asm volatile("" : "+d"(magic), "+d"(mtop));

    // k (without the row factor in EQF mode) of one column against the R rows
    auto column = [&](const double* __restrict__ ys, const double* __restrict__ ns, int j, double (&kv)[R]) {
        double yj[D];
#pragma unroll
        for (int c = 0; c < D; c++) yj[c] = ys[j * D + c];
        if constexpr (EQF == 2) {
            const double nj = ns[j];
            double r2[R];
#pragma unroll
            for (int r = 0; r < R; r++) r2[r] = fma(x[r][D - 1], yj[D - 1], nj);
#pragma unroll
            for (int c = D - 2; c >= 0; c--) {
#pragma unroll
                for (int r = 0; r < R; r++) r2[r] = fma(x[r][c], yj[c], r2[r]);
            }
#pragma unroll
            for (int r = 0; r < R; r++) r2[r] += mrow[r];
            cf_matern_fast_n<R>(r2, P.atom, c1s, g0, g1, g2, magic, mtop, tbl_lane, kv);
        } else if constexpr (EQF == 1) {
            const double nj = ns[j];
            double i0[R], t[R], f[R], p[R], s[R];
#pragma unroll
            for (int r = 0; r < R; r++) i0[r] = fma(x[r][D - 1], yj[D - 1], nj);
#pragma unroll
            for (int c = D - 2; c >= 0; c--) {
#pragma unroll
                for (int r = 0; r < R; r++) i0[r] = fma(x[r][c], yj[c], i0[r]);
            }
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
                asm("mad.lo.s32 %0, %1, 4096, %2;" : "=r"(hi) : "r"(kk), "r"(__double2hiint(tj)));
                s[r] = __hiloint2double(hi, __double2loint(tj));
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                p[r] = fma(g3, f[r], g2);
                p[r] = fma(p[r], f[r], g1);
                p[r] = fma(p[r], f[r], g0);
                const double q = fma(f[r], p[r], 1.0);
                kv[r] = s[r] * q;
            }
        } else {
            cf_rows_value<double, D, KIND, R>(x, yj, P.atom, P.sop, tbl_lane, kv);
        }
    };

    auto compute = [&](const double* __restrict__ ys, const double* __restrict__ as, const double* __restrict__ ns, int cnt, double* cb) {
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0.0;
        for (int jb = 0; jb < TJ; jb += 8) {
            if (jb >= cnt) break;
            double cv[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int j = jb + jj;
                const double aj = as[j];
                double kv[R];
                column(ys, ns, j, kv);
                double cs = kv[0] * at[0];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    acc[r] = fma(kv[r], aj, acc[r]);
                    if (r > 0) cs = fma(kv[r], at[r], cs);
                }
                cv[jj] = cs;
            }
            // transpose-reduce 8 columns over the 32 lanes (fixed tree: deterministic)
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
            double c4[4], c2[2];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double send = b16 ? cv[q] : cv[q + 4];
                const double keep = b16 ? cv[q + 4] : cv[q];
                c4[q] = keep + cf_shfl_xor_f64(send, 16);
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const double send = b8 ? c4[q] : c4[q + 2];
                const double keep = b8 ? c4[q + 2] : c4[q];
                c2[q] = keep + cf_shfl_xor_f64(send, 8);
            }
            double c1v;
            {
                const double send = b4 ? c2[0] : c2[1];
                const double keep = b4 ? c2[1] : c2[0];
                c1v = keep + cf_shfl_xor_f64(send, 4);
            }
            c1v += cf_shfl_xor_f64(c1v, 2);
            c1v += cf_shfl_xor_f64(c1v, 1);
            const int col = jb + (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0);
            if ((lane & 3) == 0) cb[warp * TJ + col] = c1v;  // one writer per (warp, column)
        }
#pragma unroll
        for (int r = 0; r < R; r++) tot[r] += acc[r];
    };
    // after the barrier that ends a tile: the warps' column totals, added in warp order, one global store per column
    auto flush_cols = [&](const double* cb, int cnt, int64_t j0) {
        if (tid < cnt) {
            double s = cb[tid];
#pragma unroll
            for (int wq = 1; wq < NW; wq++) s += cb[wq * TJ + tid];
            P.colpart[it.colpart_off + (j0 - c0) + tid] = s;
        }
    };
    static_assert(TJ <= NT, "one thread per tile column in flush_cols");

    int tile_no = 0;
    for (int t = 0; t < nfull; t++, tile_no++) {
        const int s = t % NS;
        cf_mbar_wait(&bars[s], (uint32_t)((t / NS) & 1));
        const unsigned char* st = stages + (size_t)s * S::stage_bytes;
        double* cb = colbuf + (tile_no & 1) * (NW * TJ);
        compute(reinterpret_cast<const double*>(st), reinterpret_cast<const double*>(st + S::y_bytes),
                reinterpret_cast<const double*>(st + S::y_bytes + S::v_bytes), TJ, cb);
        __syncthreads();  // stage s is free, this tile's column totals are complete
        if (tid == 0 && t + NS < nfull) issue(t + NS);
        flush_cols(cb, TJ, c0 + (int64_t)t * TJ);
    }
    for (int64_t j0 = rem0; j0 < c1; j0 += TJ, tile_no++) {
        const int cnt = (int)((c1 - j0 < TJ) ? c1 - j0 : TJ);
        double* ys = reinterpret_cast<double*>(stages);
        double* as = reinterpret_cast<double*>(stages + S::y_bytes);
        double* ns = reinterpret_cast<double*>(stages + S::y_bytes + S::v_bytes);
        double* cb = colbuf + (tile_no & 1) * (NW * TJ);
        __syncthreads();
        for (int q = tid; q < TJ * D; q += NT) ys[q] = (q < cnt * D) ? P.X[j0 * D + q] : 0.0;
        for (int q = tid; q < TJ; q += NT) {
            as[q] = (q < cnt) ? P.a[j0 + q] : 0.0;
            if (EQF) ns[q] = (q < cnt) ? P.xn[j0 + q] : 0.0;
        }
        __syncthreads();
        compute(ys, as, ns, cnt, cb);
        __syncthreads();
        flush_cols(cb, cnt, j0);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t i = rbase + (int64_t)r * NT + tid;
        if (i < P.n) P.rowpart[(int64_t)it.chunk * P.n + i] = tot[r] * ez[r];
    }
}

// y[o] = alpha * (diag[o] + sum_c rowpart[c][o] + sum_I colpart[I][o]) + beta * y[o], every sum in a fixed order.
// tr = rows per tile, ch = columns per chunk; element o belongs to row tile I = o / tr, which has ceil((n - (I+1) tr) / ch) chunks,
// and receives column partials from the tiles I' < I (triangular layout: colpart[I'] starts at I' n - tr I' (I' + 1) / 2 and holds
// the columns >= (I' + 1) tr).
// Several devices (part of parts): a device owns the row tiles I with I % parts == part and produces a PARTIAL vector over all n
// elements -- the diagonal block and the row partials of its own tiles, the column partials its tiles contribute everywhere; the
// partial vectors are then summed across devices (ncclAllReduce, or peer loads in fixed device order: cf_sum_parts_kernel).  beta y
// is added by part 0 only.
static __global__ void gram_sym_combine(const double* __restrict__ diag, const double* __restrict__ rowpart, const double* __restrict__ colpart,
                                        int64_t n, int64_t tr, int64_t ch, double* __restrict__ y, const double* __restrict__ yin, double alpha,
                                        double beta, const cf_peer_out peers, int part, int parts) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t I = o / tr;
        double s = 0.0;
        if (I % parts == part) {
            s = diag[o];
            const int64_t beyond = n - (I + 1) * tr;
            const int64_t nch = beyond > 0 ? (beyond + ch - 1) / ch : 0;
            for (int64_t c = 0; c < nch; c++) s += rowpart[c * n + o];
        }
        double cs = 0.0;
        for (int64_t Ip = part; Ip < I; Ip += parts) cs += colpart[Ip * n - tr * (Ip * (Ip + 1) / 2) + (o - (Ip + 1) * tr)];
        double v = alpha * (s + cs);
        if (beta != 0.0 && part == 0) v += beta * yin[o];
        y[o] = v;
        for (int p = 0; p < peers.n; p++) static_cast<double*>(peers.ptr[p])[o] = v;
    }
}

// out[o] = sum_q parts[q][o] (+ addend[o] * scale), fixed order q = 0, 1, ...: the cross-device sum of the symmetric variant's partial
// vectors by peer loads (single-process multi-GPU); every device runs it on the elements [o0, o1) it needs
struct cf_parts_in {
    int32_t n;
    int32_t pad_;
    const double* ptr[8];
};
static __global__ void cf_sum_parts_kernel(const cf_parts_in in, int64_t o0, int64_t o1, double* __restrict__ out, const double* __restrict__ addend,
                                           double scale) {
    for (int64_t o = o0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < o1; o += (int64_t)gridDim.x * blockDim.x) {
        double s = in.ptr[0][o];
        for (int q = 1; q < in.n; q++) s += in.ptr[q][o];
        if (addend) s += scale * addend[o];
        out[o] = s;
    }
}

#ifndef __CUDACC_RTC__
typedef cudaError_t (*cf_sym_launch_fn)(const cf_sym_params& P, int nitems, cudaStream_t stream);
template <int D, int KIND, int EQF, int R, int NT, int TJ, int NS, int MINB>
cudaError_t cf_sym_launch(const cf_sym_params& P, int nitems, cudaStream_t stream) {
    using S = cf_sym_smem<D, TJ, NS, NT / 32, EQF>;
    auto kern = gram_mvm_sym_kernel<D, KIND, EQF, R, NT, TJ, NS, MINB>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    kern<<<nitems, NT, S::total, stream>>>(P);
    return cudaGetLastError();
}
// scaled-domain EQ form: same row tile as gram_mvm_eq_kernel (cf_mvme_entry)
template <int D, bool OK = (D <= 6)>
struct cf_syme_entry {
    using E = cf_mvme_entry<D>;
    static constexpr cf_sym_launch_fn fn = &cf_sym_launch<D, CF_ATOM_EQ, 1, E::R, E::NT, E::TJ, E::NS, E::MINB>;
    static constexpr int smem = cf_sym_smem<D, E::TJ, E::NS, E::NT / 32, 1>::total;
};
template <int D>
struct cf_syme_entry<D, false> {
    static constexpr cf_sym_launch_fn fn = nullptr;
    static constexpr int smem = 0;
};
// MaternP form: same row tile as the FAST = 2 instantiation of gram_mvm_eq_kernel (cf_mvmm_entry)
template <int D, bool OK = (D <= 8)>
struct cf_symm_entry {
    using E = cf_mvmm_entry<D>;
    static constexpr cf_sym_launch_fn fn = &cf_sym_launch<D, CF_ATOM_MATERN, 2, E::R, E::NT, E::TJ, E::NS, E::MINB>;
    static constexpr int smem = cf_sym_smem<D, E::TJ, E::NS, E::NT / 32, 2>::total;
};
template <int D>
struct cf_symm_entry<D, false> {
    static constexpr cf_sym_launch_fn fn = nullptr;
    static constexpr int smem = 0;
};
#endif
