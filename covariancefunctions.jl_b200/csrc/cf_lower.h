// cf_lower.h -- host-side lowering of the reference's kernel object tree (postfix cf_knode_t program)
// into the canonical sum-of-products cf_program the device evaluates.
//
// Reference semantics being lowered (all /root/reference/src):
//   Sum      (S::Sum)(x,y) = sum(k->k(x,y), S.args)          algebra.jl:40
//   Product  (P::Product)(x,y) = prod(k->k(x,y), P.args)     algebra.jl:17
//   Power    (P::Power)(x,y) = P.k(x,y)^P.p                  algebra.jl:62
//   Constant (k::Constant)(x,y) = k.c                        stationary.jl:30-32
//   Lengthscale (k)(r2) = k.k(r2 / l^2)                      transformation.jl:19
//   traits   input_trait / sum_and_product_input_trait       properties.jl:39-63
// Products distribute over sums and integer powers of sums expand, except the pattern Dot()+sigma
// (Line, mercer.jl:12), which stays one atom so that Poly(p, sigma) is (x.y + sigma)^p as in the reference.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../../include/covfn_b200.h"
#include "cf_program.h"

namespace cf {

struct LowerError {
    int code;
    std::string msg;
};

namespace detail {

struct HTerm {
    double coef;
    std::vector<cf_factor> fac;
};
struct HExpr {
    std::vector<HTerm> terms;
    // bookkeeping for pattern matching
    bool is_plain_dot = false;   // exactly Dot()
    bool is_const = false;       // exactly Constant(c)
    double cval = 0;
    bool is_iso_leaf = false;    // isotropic leaf (possibly under Lengthscale): atom index below
    int atom = -1;
    bool is_ardscale = false;    // an ARDSCALE entry waiting for its ARD node (cval = l_c)
};

inline long double lfact(int n) { long double f = 1; for (int i = 2; i <= n; i++) f *= i; return f; }
inline long double lbinom(int n, int k) { return lfact(n) / (lfact(k) * lfact(n - k)); }

// exp(c * v) constants
inline void fill_exp(cf_exp_consts& E, long double c) {
    const long double ln2 = 0.693147180559945309417232121458176568L;
    E.c = (double)c;
    E.c1 = (double)(c * 256.0L / ln2);   // CF_EXP_TBL = 256 (cf_math.cuh)
    E.c2 = (double)(-(ln2 / 256.0L) / c);
    {   // split for CF_EXP_ACCURATE: kk (|kk| < 2^19) times c2_hi is exact in double
        const long double c2l = -(ln2 / 256.0L) / c;
        union { double d; uint64_t u; } hv;
        hv.d = (double)c2l;
        hv.u &= ~((uint64_t(1) << 19) - 1);
        E.c2_hi = hv.d;
        E.c2_lo = (double)(c2l - (long double)hv.d);
    }
    long double ci = c;
    for (int i = 0; i < CF_EXP_POLY; i++) {
        E.q[i] = (double)(ci / lfact(i + 1));
        ci *= c;
    }
    E.vmax = (double)(-700.0L / c);
    union { double d; uint64_t u; } cv;
    cv.d = E.vmax;
    E.vmax_hi = (int32_t)(cv.u >> 32);
    E.pad_ = 0;
}

// Matern nu = p + 1/2 with length scale l: value/derivative polynomials in g = sqrt(r2).
// unit scale: s = kappa g, kappa = sqrt(2p+1)/l, k = M_s(s) e^{-s},
//   M_s(s) = sum_{i=0..p} mu_i s^i,  mu_i = c_{i+1} 2^i / norm  (c = MaternP_coefficients, c_{p+1} = 1,
//   norm = (2p)!/p!; reference src/stationary.jl:148-157, 184-191)
// d/d(s^2) [F(s) e^{-s}] = ((F' - F) / (2 s)) e^{-s}  =: D(F) e^{-s}
inline void fill_matern(cf_atom& A, int p, long double l) {
    const long double nu2 = 2 * p + 1;
    const long double kappa = sqrtl(nu2) / l;
    const long double norm = lfact(2 * p) / lfact(p);
    A.v.kind = CF_ATOM_MATERN;
    A.v.p = p;
    fill_exp(A.v.e, -kappa);
    A.inv_l2 = (double)(1.0L / (l * l));
    // Laurent polynomials in s with exponents -3..p, index = exponent + 3
    const int OFF = 3, LEN = CF_MAX_MATERN_P + 1 + OFF;
    std::vector<long double> M(LEN, 0.0L), Ad(LEN, 0.0L), Bd(LEN, 0.0L);
    for (int i = 0; i <= p; i++) {
        // coefficient of (2r)^i in the reference loop: coefficients[i+1] for i < p (1-based), 1 for i = p;
        // coefficients = reverse(binomial(p,q) (p+q)!/p!, q = 1..p)  => entry for power i is q = p - i
        long double ci = (i == p) ? 1.0L : lbinom(p, p - i) * (lfact(2 * p - i) / lfact(p));
        M[i + OFF] = ci * powl(2.0L, i) / norm;
    }
    auto Dop = [&](const std::vector<long double>& F, std::vector<long double>& G) {
        for (auto& g : G) g = 0;
        for (int e = -3; e <= p; e++) {
            long double f = F[e + OFF];
            if (f == 0) continue;
            if (e != 0 && e - 2 >= -3) G[e - 2 + OFF] += e * f / 2;  // F'/(2s)
            if (e - 1 >= -3) G[e - 1 + OFF] += -f / 2;                // -F/(2s)
        }
    };
    Dop(M, Ad);
    if (p >= 1) { Ad[-1 + OFF] = 0; Ad[-2 + OFF] = 0; Ad[-3 + OFF] = 0; } // structural zeros (once differentiable in r2)
    Dop(Ad, Bd);
    if (p >= 2) { Bd[-1 + OFF] = 0; Bd[-2 + OFF] = 0; Bd[-3 + OFF] = 0; } // twice differentiable in r2
    // convert: s = kappa g;  dk/dr2 = (nu2/l^2) dk/ds2 ;  store polynomials in g
    const long double j1 = nu2 / (l * l), j2 = j1 * j1;
    for (int i = 0; i <= CF_MAX_MATERN_P; i++) {
        A.v.mat[i] = (i <= p) ? (double)(M[i + OFF] * powl(kappa, i)) : 0.0;
        A.matA[i] = (i <= p) ? (double)(j1 * Ad[i + OFF] * powl(kappa, i)) : 0.0;
        A.matB[i] = (i <= p) ? (double)(j2 * Bd[i + OFF] * powl(kappa, i)) : 0.0;
    }
    A.am1 = (double)(j1 * Ad[-1 + OFF] / kappa);
    A.bm[0] = (double)(j2 * Bd[-1 + OFF] / kappa);
    A.bm[1] = (double)(j2 * Bd[-2 + OFF] / (kappa * kappa));
    A.bm[2] = (double)(j2 * Bd[-3 + OFF] / (kappa * kappa * kappa));
    // Taylor branch: y = 1 + sum_i d_i s2^i / i!, d_i = (nu/2)^i / prod_{m<=i} (m - nu), bound eps^(1/p)
    // (reference src/stationary.jl:137-146, 172-182); s2 = r2 / l^2
    A.taylor_bound = (p == 0) ? 0.0 : std::pow(2.220446049250313e-16, 1.0 / p);
    long double nu = p + 0.5L, num = 1, den = 1;
    for (int i = 0; i <= CF_MAX_MATERN_P; i++) A.tay[i] = 0;
    A.tay[0] = 1.0;
    for (int i = 1; i <= p; i++) {
        num *= nu / 2;
        den *= (i - nu);
        A.tay[i] = (double)((num / den) / lfact(i));
    }
}

// combine terms with the same multiset of factors: (a + b)^3 has 4 distinct terms, not 8 (a Julia Sum / Power is evaluated as
// written, reference src/algebra.jl:40,62; the expansion is ours, so is the duty to keep it small)
inline void merge_like_terms(std::vector<HTerm>& terms) {
    auto canon = [](HTerm& t) {
        for (size_t i = 1; i < t.fac.size(); i++)
            for (size_t j = i; j > 0 && t.fac[j - 1].atom > t.fac[j].atom; j--) std::swap(t.fac[j - 1], t.fac[j]);
    };
    auto same = [](const HTerm& a, const HTerm& b) {
        if (a.fac.size() != b.fac.size()) return false;
        for (size_t i = 0; i < a.fac.size(); i++)
            if (a.fac[i].atom != b.fac[i].atom || a.fac[i].power != b.fac[i].power) return false;
        return true;
    };
    std::vector<HTerm> out;
    for (auto& t : terms) {
        canon(t);
        bool merged = false;
        for (auto& o : out)
            if (same(o, t)) { o.coef += t.coef; merged = true; break; }
        if (!merged) out.push_back(t);
    }
    terms.swap(out);
}

inline cf_atom zero_atom() {
    cf_atom A;
    std::memset(&A, 0, sizeof(A));
    A.inv_l2 = 1.0;
    return A;
}

}  // namespace detail

// Lower a postfix program.  Throws LowerError.  ard (may be NULL -> ARD programs are rejected) receives the per-dimension
// length scales l_c of the program's ARD node, empty if there is none; the caller applies them to the points (1/sqrt(l_c)).
inline cf_program lower(const cf_knode_t* prog, int nnodes, std::vector<double>* ard = nullptr) {
    using namespace detail;
    if (!prog || nnodes <= 0) throw LowerError{CF_ERR_BAD_ARGUMENT, "empty kernel program"};
    std::vector<cf_atom> atoms;
    std::vector<HExpr> st;
    std::vector<double> ard_l;       // the one ARD metric of the program
    std::vector<char> atom_in_ard;   // per atom: sits under the ARD node
    if (ard) ard->clear();
    auto add_atom = [&](const cf_atom& A) {
        if ((int)atoms.size() >= CF_MAX_TERMS) throw LowerError{CF_ERR_UNSUPPORTED, "too many distinct base kernels"};
        atoms.push_back(A);
        return (int)atoms.size() - 1;
    };
    auto leaf_expr = [&](int atom, bool iso) {
        HExpr e;
        e.terms.push_back(HTerm{1.0, {cf_factor{atom, 1}}});
        e.is_iso_leaf = iso;
        e.atom = atom;
        return e;
    };
    for (int t = 0; t < nnodes; t++) {
        const cf_knode_t& nd = prog[t];
        switch (nd.op) {
            case CF_OP_EQ:
            case CF_OP_EXP:
            case CF_OP_RQ:
            case CF_OP_MATERNP: {
                // collect Lengthscale wrappers directly above the leaf (reference src/transformation.jl:6-19)
                long double l = 1.0L;
                int u = t + 1;
                while (u < nnodes && prog[u].op == CF_OP_LENGTHSCALE) {
                    if (!(prog[u].fparam > 0) || !std::isfinite(prog[u].fparam))
                        throw LowerError{CF_ERR_DOMAIN, "Lengthscale: l is non-positive"};  // transformation.jl:10
                    l *= (long double)prog[u].fparam;
                    u++;
                }
                cf_atom A = zero_atom();
                A.inv_l2 = (double)(1.0L / (l * l));
                if (nd.op == CF_OP_EQ) {
                    A.v.kind = CF_ATOM_EQ;
                    fill_exp(A.v.e, -0.5L / (l * l));
                } else if (nd.op == CF_OP_EXP) {
                    fill_matern(A, 0, l);
                } else if (nd.op == CF_OP_MATERNP) {
                    if (nd.iparam < 0) throw LowerError{CF_ERR_DOMAIN, "MaternP: p is negative"};  // stationary.jl:124
                    if (nd.iparam > CF_MAX_MATERN_P) throw LowerError{CF_ERR_UNSUPPORTED, "MaternP: p > 12 not supported"};
                    fill_matern(A, nd.iparam, l);
                } else {
                    if (!(nd.fparam > 0) || !std::isfinite(nd.fparam))
                        throw LowerError{CF_ERR_DOMAIN, "RQ: alpha not positive"};  // stationary.jl:47
                    const bool is_int = nd.iparam != 0 && nd.fparam == std::floor(nd.fparam) && nd.fparam <= 64;
                    A.v.kind = is_int ? CF_ATOM_RQ_INT : CF_ATOM_RQ_REAL;
                    A.v.p = is_int ? (int)nd.fparam : 0;
                    A.v.alpha = nd.fparam;
                    A.v.w = (double)(1.0L / (2.0L * (long double)nd.fparam * l * l));
                }
                st.push_back(leaf_expr(add_atom(A), true));
                t = u - 1;
                break;
            }
            case CF_OP_LENGTHSCALE:
                // Lengthscale(Constant(c), l) is valid in the reference (Constant <: IsotropicKernel, src/stationary.jl:15): c(r2 / l^2) = c
                if (!st.empty() && st.back().is_const) {
                    if (!(nd.fparam > 0) || !std::isfinite(nd.fparam)) throw LowerError{CF_ERR_DOMAIN, "Lengthscale: l is non-positive"};
                    break;
                }
                throw LowerError{CF_ERR_UNSUPPORTED, "Lengthscale must wrap an isotropic base kernel"};
            case CF_OP_DOT: {
                cf_atom A = zero_atom();
                A.v.kind = CF_ATOM_LINE;
                A.v.sigma = 0.0;
                HExpr e = leaf_expr(add_atom(A), false);
                e.is_plain_dot = true;
                st.push_back(e);
                break;
            }
            case CF_OP_CONST: {
                if (!std::isfinite(nd.fparam)) throw LowerError{CF_ERR_DOMAIN, "Constant is not finite"};
                if (nd.fparam < 0) throw LowerError{CF_ERR_DOMAIN, "Constant is not positive semi-definite"};  // stationary.jl:17-21
                HExpr e;
                e.terms.push_back(HTerm{nd.fparam, {}});
                e.is_const = true;
                e.cval = nd.fparam;
                st.push_back(e);
                break;
            }
            case CF_OP_SUM:
            case CF_OP_PROD: {
                const int k = nd.iparam;
                if (k < 1 || k > (int)st.size()) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: Sum/Product arity"};
                std::vector<HExpr> args(st.end() - k, st.end());
                st.resize(st.size() - k);
                for (auto& a : args)
                    if (a.is_ardscale) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARDSCALE outside ARD"};
                HExpr out;
                if (nd.op == CF_OP_SUM) {
                    // Line: Dot() + sigma  (mercer.jl:12) -> one atom
                    if (k == 2 && ((args[0].is_plain_dot && args[1].is_const) || (args[1].is_plain_dot && args[0].is_const))) {
                        const HExpr& d = args[0].is_plain_dot ? args[0] : args[1];
                        const HExpr& c = args[0].is_plain_dot ? args[1] : args[0];
                        atoms[d.atom].v.sigma = c.cval;
                        out = leaf_expr(d.atom, false);
                    } else {
                        for (auto& a : args) out.terms.insert(out.terms.end(), a.terms.begin(), a.terms.end());
                    }
                } else {
                    out = args[0];
                    out.is_plain_dot = out.is_const = out.is_iso_leaf = false;
                    for (int q = 1; q < k; q++) {
                        std::vector<HTerm> prod;
                        for (auto& a : out.terms)
                            for (auto& b : args[q].terms) {
                                HTerm tt{a.coef * b.coef, a.fac};
                                for (auto& f : b.fac) {
                                    bool merged = false;
                                    for (auto& g : tt.fac)
                                        if (g.atom == f.atom) { g.power += f.power; merged = true; break; }
                                    if (!merged) tt.fac.push_back(f);
                                }
                                prod.push_back(tt);
                            }
                        merge_like_terms(prod);
                        out.terms.swap(prod);
                        if ((int)out.terms.size() > CF_MAX_TERMS) throw LowerError{CF_ERR_UNSUPPORTED, "kernel expands to too many terms"};
                    }
                    if (k == 1) { out.is_plain_dot = args[0].is_plain_dot; out.is_const = args[0].is_const; out.cval = args[0].cval; out.is_iso_leaf = args[0].is_iso_leaf; out.atom = args[0].atom; }
                }
                if ((int)out.terms.size() > CF_MAX_TERMS) throw LowerError{CF_ERR_UNSUPPORTED, "kernel expands to too many terms"};
                st.push_back(out);
                break;
            }
            case CF_OP_ARDSCALE: {
                if (!(nd.fparam > 0) || !std::isfinite(nd.fparam)) throw LowerError{CF_ERR_DOMAIN, "ARD: length scale is non-positive"};
                HExpr e;
                e.is_ardscale = true;
                e.cval = nd.fparam;
                st.push_back(e);
                break;
            }
            case CF_OP_ARD: {
                const int k = nd.iparam;
                if (k < 1 || k + 1 > (int)st.size()) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARD arity"};
                std::vector<double> l(k);
                for (int q = 0; q < k; q++) {
                    const HExpr& e = st[st.size() - k + q];
                    if (!e.is_ardscale) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARD expects ARDSCALE entries"};
                    l[q] = e.cval;
                }
                st.resize(st.size() - k);
                HExpr child = st.back();
                st.pop_back();
                if (child.is_ardscale) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARD without a kernel"};
                if (!ard) throw LowerError{CF_ERR_UNSUPPORTED, "ARD kernels are not supported by this entry point"};
                if (!ard_l.empty() && ard_l != l)
                    throw LowerError{CF_ERR_UNSUPPORTED, "two ARD kernels with different length scales in one program"};
                ard_l = l;
                atom_in_ard.resize(atoms.size(), 0);
                for (auto& tt : child.terms)
                    for (auto& f : tt.fac) {
                        if (atoms[f.atom].v.kind == CF_ATOM_LINE)   // Normed calls k(n2(tau)): k must be a function of r2
                            throw LowerError{CF_ERR_UNSUPPORTED, "ARD must wrap an isotropic kernel"};
                        atom_in_ard[f.atom] = 1;
                    }
                child.is_plain_dot = child.is_const = child.is_iso_leaf = false;
                st.push_back(child);
                break;
            }
            case CF_OP_POW: {
                if (st.empty()) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: Power without operand"};
                const int p = nd.iparam;
                if (p < 0) throw LowerError{CF_ERR_UNSUPPORTED, "negative kernel powers are not supported"};
                HExpr base = st.back();
                st.pop_back();
                if (base.is_ardscale) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARDSCALE outside ARD"};
                HExpr out;
                if (p == 0) {
                    out.terms.push_back(HTerm{1.0, {}});
                } else if (base.terms.size() == 1) {
                    HTerm tt = base.terms[0];
                    tt.coef = std::pow(tt.coef, p);
                    for (auto& f : tt.fac) f.power *= p;
                    out.terms.push_back(tt);
                } else {
                    out = base;
                    for (int q = 1; q < p; q++) {
                        std::vector<HTerm> prod;
                        for (auto& a : out.terms)
                            for (auto& b : base.terms) {
                                HTerm tt{a.coef * b.coef, a.fac};
                                for (auto& f : b.fac) {
                                    bool merged = false;
                                    for (auto& g : tt.fac)
                                        if (g.atom == f.atom) { g.power += f.power; merged = true; break; }
                                    if (!merged) tt.fac.push_back(f);
                                }
                                prod.push_back(tt);
                            }
                        merge_like_terms(prod);
                        out.terms.swap(prod);
                        if ((int)out.terms.size() > CF_MAX_TERMS) throw LowerError{CF_ERR_UNSUPPORTED, "kernel expands to too many terms"};
                    }
                }
                out.is_plain_dot = out.is_const = out.is_iso_leaf = false;
                st.push_back(out);
                break;
            }
            default:
                throw LowerError{CF_ERR_UNSUPPORTED, "unknown kernel op " + std::to_string(nd.op)};
        }
    }
    if (st.size() != 1) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: stack does not reduce to one kernel"};
    const HExpr& root = st[0];
    if (root.is_ardscale) throw LowerError{CF_ERR_BAD_ARGUMENT, "malformed program: ARDSCALE outside ARD"};
    cf_program P;
    std::memset(&P, 0, sizeof(P));
    if ((int)root.terms.size() > CF_MAX_TERMS) throw LowerError{CF_ERR_UNSUPPORTED, "kernel expands to too many terms"};
    P.nterms = (int)root.terms.size();
    P.natoms = (int)atoms.size();
    for (int a = 0; a < P.natoms; a++) {
        cf_atom& A = atoms[a];
        A.v.f_clog2e = (float)(A.v.e.c * 1.44269504088896340736);
        A.v.f_gmax = (A.v.e.c != 0.0) ? (float)(-87.0 / A.v.e.c) : 3.0e38f;  // exp(-87) ~ 1.6e-38
        A.v.f_w = (float)A.v.w; A.v.f_alpha = (float)A.v.alpha; A.v.f_sigma = (float)A.v.sigma; A.v.f_pad = 0; A.v.f_pad2 = 0;
        for (int i = 0; i <= CF_MAX_MATERN_P; i++) A.v.f_mat[i] = (float)A.v.mat[i];
        P.atoms[a] = A;
    }
    bool used[CF_MAX_TERMS] = {false};
    for (int t = 0; t < P.nterms; t++) {
        const HTerm& ht = root.terms[t];
        if ((int)ht.fac.size() > CF_MAX_FACTORS) throw LowerError{CF_ERR_UNSUPPORTED, "too many factors in one product"};
        P.terms[t].coef = ht.coef;
        P.terms[t].nfac = (int)ht.fac.size();
        for (size_t f = 0; f < ht.fac.size(); f++) { P.terms[t].fac[f] = ht.fac[f]; used[ht.fac[f].atom] = true; }
    }
    P.isotropic = 1;
    P.dotproduct = 1;
    bool any = false;
    for (int a = 0; a < P.natoms; a++) {
        if (!used[a]) continue;
        any = true;
        if (P.atoms[a].v.kind == CF_ATOM_LINE) { P.needs_dot = 1; P.isotropic = 0; }
        else { P.needs_r2 = 1; P.dotproduct = 0; }
    }
    if (!any) P.dotproduct = 0; // all-constant programs count as isotropic (reference src/properties.jl:49-50)
    if (!ard_l.empty()) {
        // the metric is applied to the points, so every atom that is used must see it: r2-atoms outside the ARD node or x.y atoms
        // would be evaluated on the wrong coordinates
        atom_in_ard.resize(atoms.size(), 0);
        for (int a = 0; a < P.natoms; a++)
            if (used[a] && !atom_in_ard[a])
                throw LowerError{CF_ERR_UNSUPPORTED, "ARD: every base kernel of the program must be wrapped by the same ARD"};
        P.isotropic = 0;  // Normed <: StationaryKernel (reference src/transformation.jl:25): not IsotropicInput, no derivative operators
        P.dotproduct = 0;
        *ard = ard_l;
    }
    P.single = (P.nterms == 1 && P.terms[0].nfac == 1 && P.terms[0].fac[0].power == 1) ? 1 : 0;
    return P;
}

// device program builders (limits of the parameter-resident forms)
inline void to_sop_val(const cf_program& P, cf_sop_val& out) {
    if (P.nterms > CF_SOP_MAX_TERMS || P.natoms > CF_SOP_MAX_ATOMS)
        throw LowerError{CF_ERR_UNSUPPORTED, "kernel is too complex for the device program (more than 8 terms or 6 base kernels)"};
    std::memset(&out, 0, sizeof(out));
    out.nterms = P.nterms; out.natoms = P.natoms;
    for (int t = 0; t < P.nterms; t++) {
        if (P.terms[t].nfac > CF_SOP_MAX_FACTORS) throw LowerError{CF_ERR_UNSUPPORTED, "more than 4 factors in one product"};
        out.terms[t].coef = P.terms[t].coef; out.terms[t].nfac = P.terms[t].nfac;
        for (int f = 0; f < P.terms[t].nfac; f++) { out.terms[t].atom[f] = P.terms[t].fac[f].atom; out.terms[t].power[f] = P.terms[t].fac[f].power; }
    }
    for (int a = 0; a < P.natoms; a++) out.atoms[a] = P.atoms[a].v;
}
inline bool to_sop_grad(const cf_program& P, cf_sop_grad& out) {
    std::memset(&out, 0, sizeof(out));
    if (P.nterms > CF_SOPG_MAX_TERMS || P.natoms > CF_SOPG_MAX_ATOMS) return false;
    out.nterms = P.nterms; out.natoms = P.natoms;
    for (int t = 0; t < P.nterms; t++) {
        if (P.terms[t].nfac > CF_SOP_MAX_FACTORS) return false;
        out.terms[t].coef = P.terms[t].coef; out.terms[t].nfac = P.terms[t].nfac;
        for (int f = 0; f < P.terms[t].nfac; f++) { out.terms[t].atom[f] = P.terms[t].fac[f].atom; out.terms[t].power[f] = P.terms[t].fac[f].power; }
    }
    for (int a = 0; a < P.natoms; a++) out.atoms[a] = P.atoms[a];
    return true;
}

}  // namespace cf
