// cf_comm.h -- multi-process row sharding: one process per GPU, the row blocks of a chained product re-assembled with NCCL.
//
// The reference is single-process (src/gramian.jl:81 gives each THREAD whole rows); across GPUs the same row partition needs
// one exchange per chained MVM: an all-gather of the product's row blocks.  The library talks to NCCL
// itself so that the exchange is reachable from the C ABI (a Julia / C host needs no torch): libnccl.so.2 is dlopen()ed on
// first use -- no link-time dependency, and inside a torch process this resolves to the copy torch already loaded.
// Bootstrap is the usual NCCL one: rank 0 calls cf_comm_unique_id, the host language ships the 128 bytes to the other ranks
// (torch.distributed / MPI / a file), every rank calls cf_comm_init.  Row blocks follow capi.cu's split (rows * r / world), so
// they may differ by one row: the gather is a group of per-rank broadcasts (in place), which NCCL fuses into one operation.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>

namespace cfcomm {

typedef struct { char internal[128]; } ncclUniqueId;  // NCCL_UNIQUE_ID_BYTES = 128 (nccl.h)
typedef void* ncclComm_t;
enum { ncclSuccess = 0, ncclUint8 = 1, ncclFloat64 = 8, ncclSum = 0 };

struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string error;
};

inline Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {std::getenv("COVFN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) { a.error = "libnccl.so.2 not found (set COVFN_NCCL_LIB)"; return; }
        auto sym = [&](const char* n) { void* s = dlsym(a.lib, n); if (!s && a.error.empty()) a.error = std::string("missing NCCL symbol ") + n; return s; };
        *(void**)&a.GetUniqueId = sym("ncclGetUniqueId");
        *(void**)&a.CommInitRank = sym("ncclCommInitRank");
        *(void**)&a.CommDestroy = sym("ncclCommDestroy");
        *(void**)&a.Broadcast = sym("ncclBroadcast");
        *(void**)&a.AllGather = sym("ncclAllGather");
        *(void**)&a.AllReduce = sym("ncclAllReduce");
        *(void**)&a.GroupStart = sym("ncclGroupStart");
        *(void**)&a.GroupEnd = sym("ncclGroupEnd");
        *(void**)&a.GetErrorString = sym("ncclGetErrorString");
        *(void**)&a.GetVersion = sym("ncclGetVersion");
    });
    return a;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, dev = -1;
};
inline Comm& comm() {
    static Comm c;
    return c;
}
inline bool active() { return comm().comm != nullptr && comm().world > 1; }

// rows [r0, r1) of rank r: the same split as capi.cu split_rows / covfn_b200.distributed.row_block
inline void row_block(int64_t n, int r, int world, int64_t* r0, int64_t* r1) {
    *r0 = n * r / world;
    *r1 = n * (r + 1) / world;
}

// In-place all-gather of a vector whose rank-r part is elements [r0_r * blk, r1_r * blk) of d_full (es bytes each).
// Returns the NCCL status (0 = ok).
inline int allgather_rows(void* d_full, int64_t n, int64_t blk, size_t es, cudaStream_t stream) {
    Api& a = api();
    Comm& c = comm();
    if (n % c.world == 0) {  // equal blocks: one in-place ncclAllGather
        const size_t bytes = (size_t)(n / c.world) * blk * es;
        return a.AllGather((const char*)d_full + (size_t)c.rank * bytes, d_full, bytes, ncclUint8, c.comm, stream);
    }
    int rc = a.GroupStart();
    if (rc) return rc;
    for (int r = 0; r < c.world; r++) {
        int64_t r0, r1;
        row_block(n, r, c.world, &r0, &r1);
        char* p = (char*)d_full + (size_t)r0 * blk * es;
        rc = a.Broadcast(p, p, (size_t)(r1 - r0) * blk * es, ncclUint8, r, c.comm, stream);
        if (rc) { a.GroupEnd(); return rc; }
    }
    return a.GroupEnd();
}

// In-place sum of a Float64 device vector over the ranks (the symmetric variant's partial vectors): ncclAllReduce(sum)
inline int allreduce_sum_f64(double* d_v, int64_t count, cudaStream_t stream) {
    return api().AllReduce(d_v, d_v, (size_t)count, ncclFloat64, ncclSum, comm().comm, stream);
}

}  // namespace cfcomm
