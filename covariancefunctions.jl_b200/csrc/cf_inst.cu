// cf_inst.cu -- compiled once per padded dimension: nvcc -DCF_D=<D> ... -o cf_inst_d<D>.o
#include "cf_registry.h"

#ifndef CF_D
#error "compile with -DCF_D=<dimension>"
#endif
#define CF_CAT2(a, b) a##b
#define CF_CAT(a, b) CF_CAT2(a, b)

namespace {
constexpr int D = CF_D;
using TU = cf_tune<D>;

template <typename T, int KIND>
constexpr cf_mvm_launch_fn mvm_fn() {
    return &cf_mvm_launch<T, D, KIND, TU::R, TU::NT, TU::TJ, TU::NS, TU::MINB>;
}

const cf_kernel_entry entry = {
    D,
    {{mvm_fn<float, CF_ATOM_EQ>(), mvm_fn<float, CF_ATOM_MATERN>(), mvm_fn<float, CF_ATOM_RQ_INT>(), mvm_fn<float, CF_ATOM_SOP>()},
     {mvm_fn<double, CF_ATOM_EQ>(), mvm_fn<double, CF_ATOM_MATERN>(), mvm_fn<double, CF_ATOM_RQ_INT>(), mvm_fn<double, CF_ATOM_SOP>()}},
    {{TU::NT * TU::R, TU::TJ, cf_mvm_smem<float, D, TU::TJ, TU::NS>::total, TU::MINB},
     {TU::NT * TU::R, TU::TJ, cf_mvm_smem<double, D, TU::TJ, TU::NS>::total, TU::MINB}},
    {{&cf_grad_launch<D, CF_ATOM_EQ, CF_GRAD_ISO, false, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>,
      &cf_grad_launch<D, CF_ATOM_SOP, CF_GRAD_ISO, false, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>,
      &cf_grad_launch<D, CF_ATOM_SOP, CF_GRAD_DOT, false, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>},
     {&cf_grad_launch<D, CF_ATOM_EQ, CF_GRAD_ISO, true, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>,
      &cf_grad_launch<D, CF_ATOM_SOP, CF_GRAD_ISO, true, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>,
      &cf_grad_launch<D, CF_ATOM_SOP, CF_GRAD_DOT, true, TU::GR, TU::GNT, TU::GTJ, TU::NS, TU::GMINB>}},
    {{TU::GNT * TU::GR, TU::GTJ, cf_grad_smem<D, TU::GTJ, TU::NS, false>::total, TU::GMINB},
     {TU::GNT * TU::GR, TU::GTJ, cf_grad_smem<D, TU::GTJ, TU::NS, true>::total, TU::GMINB}},
    // scalar multi-RHS kernels (d < 8, ill-scaled points, COVFN_MM_SCALAR): fp32 at 512 threads (+15 %), fp64 at 256 threads with the
    // x tile in shared memory and 8-wide program evaluation.  Both are bound by shared-memory operand delivery; well-scaled
    // points with d >= 8 go to the tensor-core kernels below (DMMA for fp64, 3xTF32 for fp32).
    {&cf_mm_launch<float, D, 512>, &cf_mm_launch<double, D, 256>},
    cf_mmd_entry<D>::fn,
    cf_mmd_entry<D>::smem,
    {&cf_sym_launch<D, CF_ATOM_EQ, 0, TU::R, TU::NT, TU::TJ, TU::NS, 1>, &cf_sym_launch<D, CF_ATOM_MATERN, 0, TU::R, TU::NT, TU::TJ, TU::NS, 1>,
     &cf_sym_launch<D, CF_ATOM_RQ_INT, 0, TU::R, TU::NT, TU::TJ, TU::NS, 1>, &cf_sym_launch<D, CF_ATOM_SOP, 0, TU::R, TU::NT, TU::TJ, TU::NS, 1>},
    cf_sym_smem<D, TU::TJ, TU::NS, TU::NT / 32, 0>::total,
    cf_syme_entry<D>::fn,
    cf_syme_entry<D>::smem,
    {cf_mvd_entry<D>::fn[0], cf_mvd_entry<D>::fn[1], cf_mvd_entry<D>::fn[2], cf_mvd_entry<D>::fn[3]},
    cf_mvd_entry<D>::cfg,
    {{cf_gradd_entry<D>::fn[0][0], cf_gradd_entry<D>::fn[0][1], cf_gradd_entry<D>::fn[0][2], cf_gradd_entry<D>::fn[0][3]},
     {cf_gradd_entry<D>::fn[1][0], cf_gradd_entry<D>::fn[1][1], cf_gradd_entry<D>::fn[1][2], cf_gradd_entry<D>::fn[1][3]}},
    cf_gradd_entry<D>::cfg,
    cf_mmt_entry<D>::fn,
    cf_mmt_entry<D>::sx,
    cf_mmt_entry<D>::smem,
    {cf_mvt_entry<D>::fn[0], cf_mvt_entry<D>::fn[1], cf_mvt_entry<D>::fn[2], cf_mvt_entry<D>::fn[3]},
    cf_mvt_entry<D>::cfg,
    cf_mmu_entry<D>::fn,
    cf_mmu_entry<D>::dk,
    cf_mmu_entry<D>::smem,
    cf_mvme_entry<D>::fn,
    cf_mvme_entry<D>::cfg,
    {cf_mvp_entry<D>::fn[0], cf_mvp_entry<D>::fn[1], cf_mvp_entry<D>::fn[2]},
    cf_mvp_entry<D>::cfg,
    {cf_mvu_entry<D>::fn[0], cf_mvu_entry<D>::fn[1], cf_mvu_entry<D>::fn[2], cf_mvu_entry<D>::fn[3]},
    cf_mvu_entry<D>::cfg,
    cf_mvmm_entry<D>::fn,
    cf_mvmm_entry<D>::cfg,
    cf_symm_entry<D>::fn,
    cf_symm_entry<D>::smem,
    {TU::R, TU::NT, TU::TJ, TU::NS, TU::MINB},
};
}  // namespace

const cf_kernel_entry* CF_CAT(cf_kernels_d, CF_D)() { return &entry; }
