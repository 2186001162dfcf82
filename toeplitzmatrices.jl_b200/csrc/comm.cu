// Multi-GPU plumbing of the C ABI: the ONLY exchange on the hot path is the spectrum, broadcast once per
// factorization -- plain ncclBroadcast over NVLink.  NCCL is bound at run time (dlopen of libnccl.so.2:
// the copy already loaded by the host process -- torch's, NCCL.jl's -- or the system one), so libtoeplitz_b200.so has
// no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>          // types and enums only; every call goes through the table below

#include <cstring>
#include <mutex>
#include <vector>

#include "capi_internal.h"

namespace tb {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("TB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
#define TB_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
        TB_SYM(GetUniqueId, "ncclGetUniqueId");
        TB_SYM(CommInitRank, "ncclCommInitRank");
        TB_SYM(CommInitAll, "ncclCommInitAll");
        TB_SYM(CommDestroy, "ncclCommDestroy");
        TB_SYM(Broadcast, "ncclBroadcast");
        TB_SYM(GroupStart, "ncclGroupStart");
        TB_SYM(GroupEnd, "ncclGroupEnd");
        TB_SYM(GetErrorString, "ncclGetErrorString");
#undef TB_SYM
    });
    const bool ok = api.lib && api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Broadcast &&
                    api.GroupStart && api.GroupEnd && api.GetErrorString;
    return ok ? &api : nullptr;
}

static int nccl_status(NcclApi* a, ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return TB200_OK;
    return set_error(TB200_NCCL_ERR, "%s: %s", what, a->GetErrorString(r));
}

}  // namespace tb

using namespace tb;

struct tb200_comm {
    std::vector<int> devs;               // local devices, one communicator each
    std::vector<ncclComm_t> comms;
    int nranks = 0;                      // size of the communicator (== devs.size() in single-process mode)
    int first_rank = 0;                  // rank of comms[0]
};

#define TB_NCCL(a, call, what)                                \
    do {                                                      \
        int st__ = nccl_status((a), (call), what);            \
        if (st__ != TB200_OK) return st__;                    \
    } while (0)

extern "C" {

int tb200_comm_unique_id(void* id128) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(TB200_NCCL_ERR, "libnccl.so.2 could not be loaded (set TB200_NCCL_LIB)");
    if (!id128) return set_error(TB200_BAD_ARG, "id128 is NULL");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    TB_NCCL(a, a->GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return TB200_OK;
}

int tb200_comm_init(int ndev, const int* devs, tb200_comm_t* out) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(TB200_NCCL_ERR, "libnccl.so.2 could not be loaded (set TB200_NCCL_LIB)");
    if (!out || ndev < 1) return set_error(TB200_BAD_ARG, "bad arguments");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return set_error(TB200_CUDA_ERR, "cudaGetDeviceCount failed");
    tb200_comm* c = new tb200_comm();
    for (int i = 0; i < ndev; ++i) {
        const int d = devs ? devs[i] : i;
        if (d < 0 || d >= count) { delete c; return set_error(TB200_BAD_ARG, "device %d out of range (have %d)", d, count); }
        c->devs.push_back(d);
    }
    c->comms.resize((size_t)ndev);
    c->nranks = ndev;
    c->first_rank = 0;
    int st = nccl_status(a, a->CommInitAll(c->comms.data(), ndev, c->devs.data()), "ncclCommInitAll");
    if (st != TB200_OK) { delete c; return st; }
    *out = c;
    return TB200_OK;
}

int tb200_comm_init_rank(int nranks, int rank, const void* id128, tb200_comm_t* out) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(TB200_NCCL_ERR, "libnccl.so.2 could not be loaded (set TB200_NCCL_LIB)");
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return set_error(TB200_BAD_ARG, "bad arguments");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(TB200_CUDA_ERR, "cudaGetDevice failed");
    tb200_comm* c = new tb200_comm();
    c->devs.push_back(dev);
    c->comms.resize(1);
    c->nranks = nranks;
    c->first_rank = rank;
    int st = nccl_status(a, a->CommInitRank(&c->comms[0], nranks, id, rank), "ncclCommInitRank");
    if (st != TB200_OK) { delete c; return st; }
    *out = c;
    return TB200_OK;
}

int tb200_bcast_spectrum(tb200_comm_t c, tb200_handle_t* handles, int nhandles, int root, void** streams) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(TB200_NCCL_ERR, "libnccl.so.2 could not be loaded (set TB200_NCCL_LIB)");
    if (!c || !handles) return set_error(TB200_BAD_ARG, "NULL argument");
    if (nhandles != (int)c->comms.size())
        return set_error(TB200_BAD_ARG, "expected %d handle(s), one per local device of the communicator, got %d",
                         (int)c->comms.size(), nhandles);
    if (root < 0 || root >= c->nranks) return set_error(TB200_BAD_ARG, "root %d out of range", root);
    std::vector<void*> ptr((size_t)nhandles);
    int64_t bytes0 = -1;
    for (int i = 0; i < nhandles; ++i) {
        int64_t bytes = 0;
        int st = tb200_spectrum_buffer(handles[i], &ptr[(size_t)i], &bytes);
        if (st != TB200_OK) return st;
        if (bytes0 < 0) bytes0 = bytes;
        if (bytes != bytes0) return set_error(TB200_DIM_MISMATCH, "handles of different shapes in one broadcast");
    }
    int prev = 0;
    cudaGetDevice(&prev);
    TB_NCCL(a, a->GroupStart(), "ncclGroupStart");
    for (int i = 0; i < nhandles; ++i) {
        cudaSetDevice(c->devs[(size_t)i]);
        cudaStream_t s = streams ? (cudaStream_t)streams[i] : nullptr;
        ncclResult_t r = a->Broadcast(ptr[(size_t)i], ptr[(size_t)i], (size_t)bytes0, ncclChar, root, c->comms[(size_t)i], s);
        if (r != ncclSuccess) { a->GroupEnd(); cudaSetDevice(prev); return nccl_status(a, r, "ncclBroadcast"); }
    }
    ncclResult_t r = a->GroupEnd();
    cudaSetDevice(prev);
    TB_NCCL(a, r, "ncclGroupEnd");
    count_launch(nhandles);
    for (int i = 0; i < nhandles; ++i) {
        int st = tb200_spectrum_ready(handles[i], streams ? streams[i] : nullptr);
        if (st != TB200_OK) return st;
    }
    return TB200_OK;
}

int tb200_comm_size(tb200_comm_t c, int* nranks, int* nlocal) {
    if (!c) return set_error(TB200_BAD_ARG, "communicator is NULL");
    if (nranks) *nranks = c->nranks;
    if (nlocal) *nlocal = (int)c->comms.size();
    return TB200_OK;
}

int tb200_comm_destroy(tb200_comm_t c) {
    if (!c) return TB200_OK;
    NcclApi* a = nccl_api();
    if (a) for (ncclComm_t cm : c->comms) if (cm) a->CommDestroy(cm);
    delete c;
    return TB200_OK;
}

}  // extern "C"
