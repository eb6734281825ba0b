// cf_jit.h -- run-time specialisation of composite kernel programs (host side, used by capi.cu only).
//
// The reference gets a fused, fully specialised evaluation of every kernel composition from Julia's compiler
// (src/algebra.jl:17,40,62 are inlined per concrete Sum/Product type).  The ahead-of-time kernels here interpret the
// lowered sum-of-products program instead, which costs ~45 integer/move instructions per pair for a two-term kernel
// (profiles/r1_ncu_gram_mm_dmma_c3.md).  For large problems the library therefore re-compiles THE SAME hand-written kernel
// source (embedded at build time, embed_sources.py) with NVRTC, replacing only cf_sop_value_n<N> by a generated body in
// which atom kinds, integer parameters, powers and the term list are compile-time constants.  Coefficients, length scales
// and every other real parameter stay in the kernel arguments, so one compilation serves a whole hyper-parameter search.
//
// No link-time dependency: libnvrtc and libcuda are dlopen()ed on first use; if either is missing, or compilation fails,
// the caller keeps using the ahead-of-time kernel (still CUDA -- there is no CPU path).  COVFN_JIT=0 disables, =1 forces
// specialisation for every composite program, unset = only when one call evaluates >= 2^33 kernel entries.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cf_program.h"
#include "cf_jit_sources.inc"

namespace cfjit {

struct Api {
    bool tried = false, ok = false, nvrtc_tried = false, nvrtc_ok = false;
    std::string why;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    CUresult (*LibraryLoadData)(CUlibrary*, const void*, CUjit_option*, void**, unsigned, CUlibraryOption*, void**, unsigned) = nullptr;
    CUresult (*LibraryGetKernel)(CUkernel*, CUlibrary, const char*) = nullptr;
    CUresult (*KernelSetAttribute)(CUfunction_attribute, int, CUkernel, CUdevice) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                             void**) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
};

struct Kernel {
    CUkernel k = nullptr;
    bool smem_set[64] = {false};
};

struct Stats {
    int compiled = 0, hits = 0, failures = 0, disk_hits = 0;
    double compile_seconds = 0.0;
};

struct State {
    std::mutex mu;
    Api api;
    std::map<std::string, Kernel*> cache;  // nullptr entry: compilation failed, do not retry
    Stats stats;
};
inline State& state() {
    static State s;
    return s;
}

template <typename F>
bool sym(void* h, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(h, name));
    return out != nullptr;
}

// NVRTC alone is enough to COMPILE a specialisation (cf_jit_check, also on a machine without a GPU); loading and launching
// additionally needs the driver
inline bool load_nvrtc(Api& a) {
    if (a.nvrtc_tried) return a.nvrtc_ok;
    a.nvrtc_tried = true;
    if (std::getenv("COVFN_JIT_TEST_NO_NVRTC")) { a.why = "disabled for testing"; return false; }  // exercises the fallback path
    void* hn = nullptr;
    std::vector<std::string> cands;
    if (const char* e = std::getenv("COVFN_NVRTC_LIB")) cands.push_back(e);
    cands.insert(cands.end(), {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"});
    for (const auto& c : cands)
        if ((hn = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!hn) { a.why = "libnvrtc not found"; return false; }
    a.nvrtc_ok = sym(hn, "nvrtcCreateProgram", a.CreateProgram) && sym(hn, "nvrtcDestroyProgram", a.DestroyProgram) &&
                 sym(hn, "nvrtcCompileProgram", a.CompileProgram) && sym(hn, "nvrtcGetProgramLogSize", a.GetProgramLogSize) &&
                 sym(hn, "nvrtcGetProgramLog", a.GetProgramLog) && sym(hn, "nvrtcGetCUBINSize", a.GetCUBINSize) &&
                 sym(hn, "nvrtcGetCUBIN", a.GetCUBIN) && sym(hn, "nvrtcAddNameExpression", a.AddNameExpression) &&
                 sym(hn, "nvrtcGetLoweredName", a.GetLoweredName);
    if (!a.nvrtc_ok) a.why = "missing NVRTC entry points";
    return a.nvrtc_ok;
}
inline bool load_api(Api& a) {
    if (a.tried) return a.ok;
    a.tried = true;
    if (!load_nvrtc(a)) return false;
    void* hc = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!hc) { a.why = "libcuda.so.1 not found"; return false; }
    a.ok = sym(hc, "cuLibraryLoadData", a.LibraryLoadData) && sym(hc, "cuLibraryGetKernel", a.LibraryGetKernel) &&
           sym(hc, "cuKernelSetAttribute", a.KernelSetAttribute) && sym(hc, "cuLaunchKernel", a.LaunchKernel) &&
           sym(hc, "cuDeviceGet", a.DeviceGet);
    if (!a.ok) a.why = "missing driver entry points";
    return a.ok;
}

inline const char* kind_name(int kind) {
    switch (kind) {
        case CF_ATOM_EQ: return "CF_ATOM_EQ";
        case CF_ATOM_MATERN: return "CF_ATOM_MATERN";
        case CF_ATOM_RQ_INT: return "CF_ATOM_RQ_INT";
        case CF_ATOM_RQ_REAL: return "CF_ATOM_RQ_REAL";
        default: return "CF_ATOM_LINE";
    }
}

// the structure of a program: everything cf_jit_shape.h bakes in (and nothing else)
inline std::string shape_key(const cf_sop_val& P) {
    std::string k = "A";
    for (int i = 0; i < P.natoms; i++) {
        const int kind = P.atoms[i].kind;
        const int ps = (kind == CF_ATOM_MATERN || kind == CF_ATOM_RQ_INT) ? P.atoms[i].p : 0;
        k += std::to_string(kind) + "." + std::to_string(ps) + ",";
    }
    k += "T";
    for (int t = 0; t < P.nterms; t++) {
        for (int f = 0; f < P.terms[t].nfac; f++) k += std::to_string(P.terms[t].atom[f]) + "^" + std::to_string(P.terms[t].power[f]) + "*";
        k += "+";
    }
    return k;
}

// generated cf_jit_shape.h: atoms first (each distinct atom evaluated once per group of N pairs), then the terms.  Part 1 is
// the Float64 evaluator, part 2 the Float32 one (cf_math.cuh includes the header twice, after the helpers each part needs).
inline std::string shape_source_part(const cf_sop_val& P, bool f32) {
    const std::string T = f32 ? "float" : "double";
    std::string s = "template <int N>\n__device__ __forceinline__ void ";
    s += f32 ? "cf_sop_value_f32_n(const float (&r2)[N], const float (&dt)[N], const cf_sop_val& P, float (&val)[N]) {\n"
             : "cf_sop_value_n(const double (&r2)[N], const double (&dt)[N], const cf_sop_val& P, cf_tbl_t tbl_lane, double (&val)[N]) {\n";
    for (int i = 0; i < P.natoms; i++) {
        const int kind = P.atoms[i].kind;
        const int ps = (kind == CF_ATOM_MATERN || kind == CF_ATOM_RQ_INT) ? P.atoms[i].p : 0;
        const std::string ai = "a" + std::to_string(i);
        s += "    " + T + " " + ai + "[N];\n";
        s += std::string("    ") + (f32 ? "cf_atom_pow_f32_s" : "cf_atom_pow_s") + "<N, " + kind_name(kind) + ", " + std::to_string(ps) +
             ", 1>(r2, dt, P.atoms[" + std::to_string(i) + "], " + (f32 ? "" : "tbl_lane, ") + ai + ");\n";
    }
    s += "#pragma unroll\n    for (int u = 0; u < N; u++) {\n        " + T + " v = 0, p;\n";
    for (int t = 0; t < P.nterms; t++) {
        const cf_sop_term& Tm = P.terms[t];
        const std::string coef = std::string(f32 ? "(float)" : "") + "P.terms[" + std::to_string(t) + "].coef";
        if (Tm.nfac == 0) {
            s += "        v += " + coef + ";\n";
            continue;
        }
        std::string prod;
        for (int f = 0; f < Tm.nfac; f++)
            for (int q = 0; q < Tm.power[f]; q++) prod += (prod.empty() ? "" : " * ") + ("a" + std::to_string(Tm.atom[f]) + "[u]");
        s += "        p = " + prod + ";\n";
        s += (t == 0) ? "        v = " + coef + " * p;\n" : std::string("        v = ") + (f32 ? "fmaf(" : "fma(") + coef + ", p, v);\n";
    }
    s += "        val[u] = v;\n    }\n}\n";
    return s;
}
// structure of a derivative program (same encoding as shape_key)
inline std::string shape_key_grad(const cf_sop_grad& P) {
    std::string k = "G";
    for (int i = 0; i < P.natoms; i++) {
        const int kind = P.atoms[i].v.kind;
        const int ps = (kind == CF_ATOM_MATERN || kind == CF_ATOM_RQ_INT) ? P.atoms[i].v.p : 0;
        k += std::to_string(kind) + "." + std::to_string(ps) + ",";
    }
    k += "T";
    for (int t = 0; t < P.nterms; t++) {
        for (int f = 0; f < P.terms[t].nfac; f++) k += std::to_string(P.terms[t].atom[f]) + "^" + std::to_string(P.terms[t].power[f]) + "*";
        k += "+";
    }
    return k;
}
// part 3: jets (k, k', k'') of the derivative program by the product rule, atoms first; G == nullptr: declaration only
inline std::string shape_source_jets(const cf_sop_grad* G) {
    const std::string sig = "template <int N>\n__device__ __forceinline__ void cf_sop_jet_n(const double (&r2)[N], const cf_sop_grad& P, "
                            "cf_tbl_t tbl_lane, double (&k)[N], double (&k1)[N], double (&k2)[N])";
    if (!G) return sig + ";\n";
    std::string s = sig + " {\n";
    for (int i = 0; i < G->natoms; i++) {
        const int kind = G->atoms[i].v.kind;
        const int ps = (kind == CF_ATOM_MATERN || kind == CF_ATOM_RQ_INT) ? G->atoms[i].v.p : 0;
        const std::string id = std::to_string(i);
        s += "    double v" + id + "[N], d" + id + "[N], e" + id + "[N];\n";
        s += "    cf_atom_jet_s<N, " + std::string(kind_name(kind)) + ", " + std::to_string(ps) + ">(r2, P.atoms[" + id + "], tbl_lane, v" + id +
             ", d" + id + ", e" + id + ");\n";
    }
    s += "#pragma unroll\n    for (int u = 0; u < N; u++) {\n        double sv = 0.0, s1 = 0.0, s2 = 0.0, pv, p1, p2, nv, n1, n2;\n";
    for (int t = 0; t < G->nterms; t++) {
        const cf_sop_term& T = G->terms[t];
        s += "        pv = P.terms[" + std::to_string(t) + "].coef; p1 = 0.0; p2 = 0.0;\n";
        for (int f = 0; f < T.nfac; f++)
            for (int q = 0; q < T.power[f]; q++) {
                const std::string a = std::to_string(T.atom[f]);
                s += "        nv = pv * v" + a + "[u]; n1 = fma(p1, v" + a + "[u], pv * d" + a + "[u]); n2 = fma(p2, v" + a + "[u], fma(2.0 * p1, d" + a +
                     "[u], pv * e" + a + "[u])); pv = nv; p1 = n1; p2 = n2;\n";
            }
        s += "        sv += pv; s1 += p1; s2 += p2;\n";
    }
    s += "        k[u] = sv; k1[u] = s1; k2[u] = s2;\n    }\n}\n";
    return s;
}
inline std::string shape_source(const cf_sop_val& P, const cf_sop_grad* G = nullptr) {
    return "// generated by cf_jit.h for program structure " + shape_key(P) + (G ? " " + shape_key_grad(*G) : std::string()) +
           "\n#if CF_JIT_PART == 1\n" + shape_source_part(P, false) + "#elif CF_JIT_PART == 2\n" + shape_source_part(P, true) +
           "#elif CF_JIT_PART == 3\n" + shape_source_jets(G) + "#endif\n";
}

// ---- on-disk cache of compiled cubins: $COVFN_JIT_CACHE or ~/.cache/covfn_b200 (COVFN_JIT_CACHE=off disables) ------------------
// The file name hashes the kernel name, the program structure and a stamp of this library build (the embedded headers change
// with it), so a stale cubin can never be picked up; files are written to a temporary name and renamed.
inline uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
inline std::string cache_path(const std::string& key) {
    const char* e = std::getenv("COVFN_JIT_CACHE");
    std::string dir;
    if (e && std::string(e) == "off") return "";
    if (e && *e) dir = e;
    else if (const char* h = std::getenv("HOME")) dir = std::string(h) + "/.cache/covfn_b200";
    else return "";
    ::mkdir((dir.substr(0, dir.rfind('/'))).c_str(), 0755);
    ::mkdir(dir.c_str(), 0755);
    uint64_t h = fnv1a(key);
    for (int i = 0; i < cf_jit_num_headers; i++) h = fnv1a(cf_jit_header_srcs[i], h);  // build stamp: the embedded sources
    char name[64];
    std::snprintf(name, sizeof(name), "/%016llx.cubin", (unsigned long long)h);
    return dir + name;
}
inline bool cache_read(const std::string& path, std::vector<char>& out, std::string& lowered) {
    if (path.empty()) return false;
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    uint32_t nl = 0;
    uint64_t nb = 0;
    f.read(reinterpret_cast<char*>(&nl), 4);
    f.read(reinterpret_cast<char*>(&nb), 8);
    if (!f || nl == 0 || nl > 4096 || nb == 0 || nb > (1ull << 30)) return false;
    lowered.resize(nl);
    out.resize(nb);
    f.read(&lowered[0], nl);
    f.read(out.data(), (std::streamsize)nb);
    return (bool)f;
}
inline void cache_write(const std::string& path, const std::vector<char>& cubin, const std::string& lowered) {
    if (path.empty()) return;
    const std::string tmp = path + "." + std::to_string((long)::getpid()) + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f) return;
        const uint32_t nl = (uint32_t)lowered.size();
        const uint64_t nb = cubin.size();
        f.write(reinterpret_cast<const char*>(&nl), 4);
        f.write(reinterpret_cast<const char*>(&nb), 8);
        f.write(lowered.data(), nl);
        f.write(cubin.data(), (std::streamsize)nb);
        if (!f) { ::unlink(tmp.c_str()); return; }
    }
    if (::rename(tmp.c_str(), path.c_str()) != 0) ::unlink(tmp.c_str());
}

// Compile (only) the specialisation of `name_expr` from `entry_header` for the structure of P; returns 0 on success, 1 if NVRTC is
// unavailable, 2 on a compilation error.  `log` receives the NVRTC log or the reason.
inline int compile_only(const cf_sop_val& P, const std::string& entry_header, const std::string& name_expr, std::string& log,
                        size_t* cubin_bytes, const cf_sop_grad* G = nullptr) {
    State& st = state();
    std::lock_guard<std::mutex> lk(st.mu);
    if (!load_nvrtc(st.api)) { log = st.api.why; return 1; }
    Api& a = st.api;
    const std::string shape = shape_source(P, G);
    const std::string main_src = "#include \"" + entry_header + "\"\n";
    std::vector<const char*> names(cf_jit_header_names, cf_jit_header_names + cf_jit_num_headers);
    std::vector<const char*> srcs(cf_jit_header_srcs, cf_jit_header_srcs + cf_jit_num_headers);
    names.push_back("cf_jit_shape.h");
    srcs.push_back(shape.c_str());
    nvrtcProgram prog = nullptr;
    if (a.CreateProgram(&prog, main_src.c_str(), "cf_jit.cu", (int)names.size(), srcs.data(), names.data()) != NVRTC_SUCCESS) {
        log = "nvrtcCreateProgram failed";
        return 2;
    }
    a.AddNameExpression(prog, name_expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DCF_JIT_SHAPE=1"};
    const nvrtcResult rc = a.CompileProgram(prog, 4, opts);
    size_t n = 0;
    a.GetProgramLogSize(prog, &n);
    log.assign(n, ' ');
    if (n) a.GetProgramLog(prog, &log[0]);
    size_t nb = 0;
    if (rc == NVRTC_SUCCESS) a.GetCUBINSize(prog, &nb);
    if (cubin_bytes) *cubin_bytes = nb;
    a.DestroyProgram(&prog);
    if (rc != NVRTC_SUCCESS) log += "\n--- generated cf_jit_shape.h ---\n" + shape;
    return rc == NVRTC_SUCCESS && nb > 0 ? 0 : 2;
}

// Returns the specialised kernel for (shape of P, name_expr), compiling it on first use; nullptr if unavailable.
inline Kernel* get_kernel(const cf_sop_val& P, const std::string& entry_header, const std::string& name_expr,
                          const cf_sop_grad* G = nullptr) {
    State& st = state();
    std::lock_guard<std::mutex> lk(st.mu);
    const std::string key = name_expr + "|" + shape_key(P) + (G ? "|" + shape_key_grad(*G) : std::string());
    auto it = st.cache.find(key);
    if (it != st.cache.end()) {
        if (it->second) st.stats.hits++;
        return it->second;
    }
    Kernel*& slot = st.cache[key];
    slot = nullptr;
    if (!load_api(st.api)) {
        st.stats.failures++;
        return nullptr;
    }
    Api& a = st.api;
    const bool verbose = std::getenv("COVFN_JIT_VERBOSE") != nullptr;
    const std::string disk = cache_path(key);
    {
        std::vector<char> cubin0;
        std::string lowered0;
        CUlibrary lib0 = nullptr;
        CUkernel kern0 = nullptr;
        if (cache_read(disk, cubin0, lowered0) &&
            a.LibraryLoadData(&lib0, cubin0.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == CUDA_SUCCESS &&
            a.LibraryGetKernel(&kern0, lib0, lowered0.c_str()) == CUDA_SUCCESS) {
            st.stats.disk_hits++;
            if (verbose) std::fprintf(stderr, "[covfn_b200] specialised %s loaded from %s\n", key.c_str(), disk.c_str());
            slot = new Kernel();
            slot->k = kern0;
            return slot;
        }
    }
    const std::string shape = shape_source(P, G);
    const std::string main_src = "#include \"" + entry_header + "\"\n";
    std::vector<const char*> names(cf_jit_header_names, cf_jit_header_names + cf_jit_num_headers);
    std::vector<const char*> srcs(cf_jit_header_srcs, cf_jit_header_srcs + cf_jit_num_headers);
    names.push_back("cf_jit_shape.h");
    srcs.push_back(shape.c_str());
    nvrtcProgram prog = nullptr;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (a.CreateProgram(&prog, main_src.c_str(), "cf_jit.cu", (int)names.size(), srcs.data(), names.data()) != NVRTC_SUCCESS) {
        st.stats.failures++;
        return nullptr;
    }
    a.AddNameExpression(prog, name_expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DCF_JIT_SHAPE=1"};
    const nvrtcResult rc = a.CompileProgram(prog, 4, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        a.GetProgramLogSize(prog, &n);
        std::string log(n, ' ');
        if (n) a.GetProgramLog(prog, &log[0]);
        std::fprintf(stderr, "[covfn_b200] run-time specialisation failed for %s (falling back to the interpreter kernel):\n%s\n%s\n",
                     key.c_str(), log.c_str(), verbose ? shape.c_str() : "");
        a.DestroyProgram(&prog);
        st.stats.failures++;
        return nullptr;
    }
    const char* lowered = nullptr;
    size_t nb = 0;
    std::vector<char> cubin;
    CUlibrary lib = nullptr;
    CUkernel kern = nullptr;
    bool ok = a.GetLoweredName(prog, name_expr.c_str(), &lowered) == NVRTC_SUCCESS && lowered &&
              a.GetCUBINSize(prog, &nb) == NVRTC_SUCCESS && nb > 0;
    if (ok) {
        cubin.resize(nb);
        ok = a.GetCUBIN(prog, cubin.data()) == NVRTC_SUCCESS &&
             a.LibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == CUDA_SUCCESS &&
             a.LibraryGetKernel(&kern, lib, lowered) == CUDA_SUCCESS;
    }
    const std::string lowered_name = (ok && lowered) ? lowered : "";  // owned by the program: copy before destroying it
    a.DestroyProgram(&prog);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (!ok) {
        std::fprintf(stderr, "[covfn_b200] run-time specialisation: loading the compiled kernel failed for %s\n", key.c_str());
        st.stats.failures++;
        return nullptr;
    }
    cache_write(disk, cubin, lowered_name);
    const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    st.stats.compiled++;
    st.stats.compile_seconds += secs;
    if (verbose) std::fprintf(stderr, "[covfn_b200] specialised %s in %.2f s\n%s", key.c_str(), secs, shape.c_str());
    slot = new Kernel();
    slot->k = kern;
    return slot;
}

// launch with one by-value parameter struct (all kernels here take `const __grid_constant__ params P`)
inline int launch(Kernel* K, const void* params, unsigned gx, unsigned gy, unsigned block, unsigned smem_bytes, cudaStream_t stream) {
    State& st = state();
    Api& a = st.api;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(st.mu);
        if (!K->smem_set[dev & 63]) {
            CUdevice cd;
            if (a.DeviceGet(&cd, dev) != CUDA_SUCCESS) return 1;
            if (a.KernelSetAttribute(CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem_bytes, K->k, cd) != CUDA_SUCCESS) return 1;
            K->smem_set[dev & 63] = true;
        }
    }
    void* args[] = {const_cast<void*>(params)};
    return a.LaunchKernel(reinterpret_cast<CUfunction>(K->k), gx, gy, 1, block, 1, 1, smem_bytes, reinterpret_cast<CUstream>(stream), args,
                          nullptr) == CUDA_SUCCESS
               ? 0
               : 1;
}

// policy: 0 = never, 1 = always for composite programs, -1 (unset) = by problem size
inline int mode() {
    const char* e = std::getenv("COVFN_JIT");
    if (!e) return -1;
    return std::atoi(e) != 0 ? 1 : 0;
}
inline bool wanted(double entries_per_call) {
    const int m = mode();
    if (m == 0) return false;
    if (m == 1) return true;
    return entries_per_call >= 8589934592.0;  // 2^33: ~0.1 s of kernel time, against ~1.5 s of compilation paid once per shape
}

}  // namespace cfjit
