// cf_program.h -- lowered kernel program shared by host (lowering) and device (evaluation).
//
// The reference evaluates a kernel object tree recursively per pair (src/algebra.jl:17,40,62), each
// leaf recomputing its own squared distance or dot product.  The device evaluates a canonical
// sum-of-products instead:   k(x,y) = sum_t coef_t * prod_f atom_f(r2 or x.y)^power_f
// with ONE r2 and ONE x.y per pair shared by every atom (kernel K3).
#pragma once
#ifdef __CUDACC_RTC__ // NVRTC has no system headers
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#define CF_MAX_MATERN_P 12
#define CF_MAX_TERMS 16
#define CF_MAX_FACTORS 6
#define CF_EXP_POLY 4 /* degree of the exp polynomial after the 256-entry table reduction */

enum cf_atom_kind {
    CF_ATOM_EQ = 0,      // exp(c r2), c = -1/(2 l^2)                      reference src/stationary.jl:42
    CF_ATOM_MATERN = 1,  // M(g) exp(c g), g = sqrt(r2); p = 0 is Exp       reference src/stationary.jl:60,134-158
    CF_ATOM_RQ_INT = 2,  // (1 + w r2)^-a, a positive Int                  reference src/stationary.jl:53
    CF_ATOM_RQ_REAL = 3, // (1 + w r2)^-a, a real
    CF_ATOM_LINE = 4,    // x.y + sigma  (Dot: sigma = 0)                  reference src/mercer.jl:9,12
    CF_ATOM_SOP = 5      // (kernel-kind tag only) generic sum of products
};

// Constants for exp(c*v) = 2^k * T[j] * P(u):  t = fma(v, c1, MAGIC); kk = t - MAGIC = 256 k + j;
// u = fma(kk, c2, v) (so c*u is the reduced argument, |c u| <= ln2/512); P(u) = 1 + u*(q0 + q1 u + ...)
struct cf_exp_consts {
    double c1;             // c * 256 / ln2
    double c2;             // -(ln2 / 256) / c
    double c2_hi, c2_lo;   // the same constant split for the two-step (accurate) reduction: c2_hi has 34 significant bits
    double q[CF_EXP_POLY]; // c^(i+1) / (i+1)!
    double c;              // the plain multiplier (fp32 path, derivative formulas)
    double vmax;           // clamp: c*vmax = -700 (result ~1e-304, i.e. 0)
    int32_t vmax_hi;       // high word of vmax
    int32_t pad_;
};

// value-path constants of one atom (kept small: programs travel in kernel parameters = constant bank, so that every
// coefficient is an instruction operand instead of a dependent global/shared load)
struct cf_atom_val {
    int32_t kind;
    int32_t p;       // Matern p / integer alpha / unused
    cf_exp_consts e; // EQ, MATERN
    // MATERN value polynomial in g: M(g) = sum_i mat[i] g^i (degree p), already divided by (2p)!/p!
    double mat[CF_MAX_MATERN_P + 1];
    double w, alpha; // RQ: w = 1 / (2 alpha l^2)
    double sigma;    // LINE
    // Float32 copies (no fp64 instruction in the fp32 inner loop)
    float f_clog2e;  // c * log2(e): exp(c v) = ex2(f_clog2e * v)
    float f_gmax;    // clamp of g = sqrt(r2) so that M(g) exp(c g) underflows cleanly
    float f_w, f_alpha, f_sigma, f_pad;
    float f_mat[CF_MAX_MATERN_P + 1];
    float f_pad2;
};

// full atom: value constants + derivative data for the gradient kernel
struct cf_atom {
    cf_atom_val v;
    // MATERN derivative polynomials: k1 = (A(g) + am1/g) e, k2 = (B(g) + bm[0]/g + bm[1]/g^2 + bm[2]/g^3) e
    double matA[CF_MAX_MATERN_P + 1];
    double matB[CF_MAX_MATERN_P + 1];
    double am1, bm[3];
    // MATERN Taylor branch (r2 / l^2 < taylor_bound): derivatives at zero d_i / i!  (reference src/stationary.jl:139-146)
    double taylor_bound;
    double tay[CF_MAX_MATERN_P + 1]; // tay[0] = 1
    double inv_l2;                   // plain r2 multiplier 1/l^2
};

struct cf_factor {
    int32_t atom;  // index into atoms[]
    int32_t power; // >= 1
};

struct cf_term {
    double coef;
    int32_t nfac;
    int32_t pad_;
    cf_factor fac[CF_MAX_FACTORS];
};

// host-side lowered program (generous limits)
struct cf_program {
    int32_t nterms;
    int32_t natoms;
    int32_t needs_r2, needs_dot;
    int32_t isotropic; // every atom is a function of r2 (IsotropicInput trait, reference src/properties.jl:39-63)
    int32_t dotproduct; // every atom is a function of x.y (DotProductInput trait)
    int32_t single;    // 1 if the program is coef * one atom ^ 1 -> specialised kernels
    cf_term terms[CF_MAX_TERMS];
    cf_atom atoms[CF_MAX_TERMS];
};

// device-side programs, passed BY VALUE in kernel parameters
#define CF_SOP_MAX_TERMS 8
#define CF_SOP_MAX_FACTORS 4
#define CF_SOP_MAX_ATOMS 6
#define CF_SOPG_MAX_TERMS 4
#define CF_SOPG_MAX_ATOMS 3
struct cf_sop_term {
    double coef;
    int32_t nfac;
    int32_t atom[CF_SOP_MAX_FACTORS];
    int32_t power[CF_SOP_MAX_FACTORS];
    int32_t pad_;
};
struct cf_sop_val { // value kernels (MVM, multi-RHS, dense)
    int32_t nterms, natoms;
    cf_sop_term terms[CF_SOP_MAX_TERMS];
    cf_atom_val atoms[CF_SOP_MAX_ATOMS];
};
struct cf_sop_grad { // gradient kernel (needs the derivative data)
    int32_t nterms, natoms;
    cf_sop_term terms[CF_SOPG_MAX_TERMS];
    cf_atom atoms[CF_SOPG_MAX_ATOMS];
};
