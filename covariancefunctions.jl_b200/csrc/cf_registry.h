// cf_registry.h -- table of compiled kernel instantiations, one entry per padded point dimension D.
#pragma once
#include "gram_mvm.cuh"
#include "grad_mvm.cuh"
#include "cf_extra.cuh"
#include "gram_mvm_sym.cuh"
#include "gram_mm_dmma.cuh"
#include "gram_mvm_dmma.cuh"
#include "grad_mvm_dmma.cuh"
#include "gram_mm_tf32.cuh"
#include "gram_mvm_tf32.cuh"
#include "gram_mvm_eq.cuh"
#include "gram_mm_tc5.cuh"
#include "gram_mvm_f32p.cuh"
#include "gram_mvm_tc5.cuh"

#define CF_NKINDS 4 /* EQ, MATERN, RQ_INT, SOP */
inline int cf_kind_slot(int kind) {
    switch (kind) {
        case CF_ATOM_EQ: return 0;
        case CF_ATOM_MATERN: return 1;
        case CF_ATOM_RQ_INT: return 2;
        default: return 3;
    }
}

struct cf_kernel_entry {
    int D;
    cf_mvm_launch_fn mvm[2][CF_NKINDS]; // [dtype][kind slot]
    cf_mvm_config mvm_cfg[2];
    // gradient kernels: [value_gradient][variant]; variants: 0 = isotropic EQ specialised, 1 = isotropic generic, 2 = dot product
    cf_grad_launch_fn grad[2][3];
    cf_mvm_config grad_cfg[2];
    cf_mm_launch_fn mm[2]; // [dtype]
    cf_mm_launch_fn mm_dmma; // Float64 tensor-core (DMMA) variant, nullptr when D % 4 != 0 or D < 8
    int mm_dmma_smem;        // its dynamic shared memory (run-time specialised launches need it)
    cf_sym_launch_fn sym[CF_NKINDS]; // Float64 symmetric variant (gram_mvm_sym.cuh), [kind slot]; row tile = mvm_cfg[1].rows_per_cta
    int sym_smem;                    // its dynamic shared memory
    cf_sym_launch_fn sym_eq;         // ... with the scaled-domain EQ evaluation (row tile = mvm_eq_cfg.rows_per_cta), nullptr for D > 6
    int sym_eq_smem;
    cf_mvm_launch_fn mvm_dmma[CF_NKINDS]; // Float64 tensor-core value MVM (gram_mvm_dmma.cuh), nullptr when unavailable for D
    cf_mvm_config mvm_dmma_cfg;
    cf_gradd_launch_fn grad_dmma[2][4]; // Float64 tensor-core gradient MVM: [value_gradient][0 EQ, 1 generic isotropic, 2 MaternP(p>=2), 3 dot product]; nullptr when unavailable
    cf_mvm_config grad_dmma_cfg;
    cf_mm_launch_fn mm_tf32; // Float32 multi-RHS on the tensor cores in 3xTF32 (gram_mm_tf32.cuh), nullptr for D < 8
    int mm_tf32_sx;          // row stride (floats) of its padded point copies
    int mm_tf32_smem;        // its dynamic shared memory (run-time specialised launches)
    cf_mvm_launch_fn mvm_tf32[CF_NKINDS]; // Float32 value MVM with the distance GEMM in 3xTF32 (gram_mvm_tf32.cuh), nullptr for D < 8
    cf_mvm_config mvm_tf32_cfg;
    cf_mmu_launch_fn mm_tc5;  // Float32 multi-RHS on tcgen05 / TMEM in 3xTF32 (gram_mm_tc5.cuh), nullptr for D < 8
    int mm_tc5_dk;            // k extent of its distance GEMM (D rounded up to 8)
    int mm_tc5_smem;          // its dynamic shared memory (run-time specialised launches)
    cf_mvm_launch_fn mvm_eq;  // Float64 EQ value MVM with the exponent formed in the scaled domain (gram_mvm_eq.cuh), nullptr for D > 6
    cf_mvm_config mvm_eq_cfg;
    cf_mvm_launch_fn mvm_f32p[3]; // Float32 value MVM in packed FP32 arithmetic (gram_mvm_f32p.cuh): EQ, MaternP, RQ_INT; nullptr for D > 8
    cf_mvm_config mvm_f32p_cfg;
    cf_mvu_launch_fn mvm_tc5[CF_NKINDS]; // Float32 value MVM with the dot products on tcgen05 / TMEM in 3xTF32 (gram_mvm_tc5.cuh), nullptr for D < 8
    cf_mvm_config mvm_tc5_cfg;
    cf_mvm_launch_fn mvm_mat;  // Float64 MaternP (p >= 1) value MVM from the norm expansion (gram_mvm_eq.cuh, FAST = 2), nullptr for D > 8
    cf_mvm_config mvm_mat_cfg;
    cf_sym_launch_fn sym_mat;  // ... and the symmetric variant with that evaluation (row tile = mvm_mat_cfg.rows_per_cta)
    int sym_mat_smem;
    int tune[5];                     // R, NT, TJ, NS, MINB of the value MVM kernel (names the instantiation for cf_jit.h)
};

// tuning per D: rows per thread R, threads NT, tile TJ, stages NS, min CTAs/SM
template <int D> struct cf_tune {
    // D = 6, 8: 3 rows at <= 128 registers keeps 2 CTAs per SM and measured +7 % over R = 2 (R = 4 needs 156 registers: -10 %)
    static constexpr int R = (D <= 4) ? 4 : (D <= 8 ? 3 : (D <= 16 ? 2 : 1));
    static constexpr int NT = 256;
    static constexpr int TJ = (D <= 8) ? 128 : 64;
    static constexpr int NS = 3;
    static constexpr int MINB = (D <= 8) ? 2 : 1;
    // gradient kernel
    static constexpr int GR = (D <= 4) ? 2 : 1; // (D = 16 with 2 rows: +2.5 % at n = 65536 but -45 % at n = 1024: fewer, larger CTAs)
    static constexpr int GNT = (D <= 8) ? 256 : 128;
    static constexpr int GTJ = (D <= 8) ? 128 : (D <= 16 ? 64 : 32);
    static constexpr int GMINB = (D <= 6) ? 2 : 1;
};

const cf_kernel_entry* cf_kernels_d1();
const cf_kernel_entry* cf_kernels_d2();
const cf_kernel_entry* cf_kernels_d3();
const cf_kernel_entry* cf_kernels_d4();
const cf_kernel_entry* cf_kernels_d6();
const cf_kernel_entry* cf_kernels_d8();
const cf_kernel_entry* cf_kernels_d12();
const cf_kernel_entry* cf_kernels_d16();
const cf_kernel_entry* cf_kernels_d24();
const cf_kernel_entry* cf_kernels_d32();
