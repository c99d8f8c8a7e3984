// comm.cu -- the one collective of the deployment behind the C ABI: the evaluation key reaches every GPU by a broadcast
// over NCCL (NVLink / NVSwitch) at setup.  A Swift host has no torch.distributed: it hands a
// 128-byte NCCL unique id from rank 0 to the other ranks through whatever channel it already has (a file, its RPC
// layer) and calls these entry points.  NCCL is opened with dlopen so that libhecuda loads on hosts without it;
// inside a process that already loaded an NCCL (e.g. PyTorch's) the same soname resolves to that copy.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/hecuda.h"
#include "capi_internal.hpp"

using namespace hecuda;
using namespace hecuda::api;

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[HECUDA_COMM_UNIQUE_ID_BYTES];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint64 = 5 };  // ncclDataType_t, nccl.h

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = std::getenv("HECUDA_NCCL_LIBRARY");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            if (!name) continue;
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) {
            n.error = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found");
            return;
        }
        n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(n.lib, "ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))dlsym(n.lib, "ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))dlsym(n.lib, "ncclCommDestroy");
        n.Broadcast = (decltype(n.Broadcast))dlsym(n.lib, "ncclBroadcast");
        n.GetErrorString = (decltype(n.GetErrorString))dlsym(n.lib, "ncclGetErrorString");
        if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.Broadcast) n.error = "NCCL library lacks the expected symbols";
    });
    return n;
}

int32_t nccl_fail(int rc, const char *what) {
    Nccl &n = nccl();
    return fail(HECUDA_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}

}  // namespace

struct hecuda_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;
};

extern "C" {

int32_t hecuda_comm_unique_id(uint8_t *id) {
    if (!id) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null id");
    Nccl &n = nccl();
    if (!n.error.empty()) return fail(HECUDA_ERR_UNSUPPORTED, n.error);
    ncclUniqueId u;
    const int rc = n.GetUniqueId(&u);
    if (rc != ncclSuccess) return nccl_fail(rc, "ncclGetUniqueId");
    std::memcpy(id, u.internal, HECUDA_COMM_UNIQUE_ID_BYTES);
    return HECUDA_OK;
}

int32_t hecuda_comm_create(const uint8_t *id, int32_t rank, int32_t world_size, hecuda_comm **out) {
    if (!id || !out) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size) return fail(HECUDA_ERR_INVALID_ARGUMENT, "rank / world_size out of range");
    Nccl &n = nccl();
    if (!n.error.empty()) return fail(HECUDA_ERR_UNSUPPORTED, n.error);
    hecuda_comm *c = new (std::nothrow) hecuda_comm();
    if (!c) return fail(HECUDA_ERR_CUDA, "out of host memory");
    c->rank = rank;
    c->world = world_size;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return cuda_fail(e, "hecuda_comm_create");
    }
    ncclUniqueId u;
    std::memcpy(u.internal, id, HECUDA_COMM_UNIQUE_ID_BYTES);
    const int rc = n.CommInitRank(&c->comm, world_size, u, rank);  // collective: returns when every rank has joined
    if (rc != ncclSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return nccl_fail(rc, "ncclCommInitRank");
    }
    *out = c;
    return HECUDA_OK;
}

int32_t hecuda_comm_destroy(hecuda_comm *c) {
    if (!c) return HECUDA_OK;
    Nccl &n = nccl();
    if (c->comm && n.CommDestroy) n.CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return HECUDA_OK;
}

// Broadcasts the relinearization key (when has_relin) and the Galois keys of `elements` from rank `root` into the
// matching device buffers of every other rank's (empty) evaluation key.  Collective: every rank calls it with the same
// has_relin / elements / root.  On return the key is usable on this rank (the call waits for its stream).
int32_t hecuda_evk_broadcast(hecuda_evk *evk, hecuda_comm *comm, int32_t root, int32_t has_relin, const uint32_t *elements,
                             int32_t element_count) {
    if (!evk || !comm) return fail(HECUDA_ERR_INVALID_ARGUMENT, "null argument");
    if (root < 0 || root >= comm->world) return fail(HECUDA_ERR_INVALID_ARGUMENT, "root out of range");
    if (element_count < 0 || (element_count && !elements)) return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalid element list");
    int32_t rc = check_ctx(evk->owner);
    if (rc) return rc;
    if (evk->owner->ctx->device != comm->device)
        return fail(HECUDA_ERR_INVALID_ARGUMENT, "invalidContext: communicator and evaluation key live on different devices");
    Nccl &n = nccl();
    if (!n.error.empty()) return fail(HECUDA_ERR_UNSUPPORTED, n.error);
    if (has_relin) {
        void *p = nullptr;
        uint64_t bytes = 0;
        if (comm->rank == root && !evk->loaded) return fail(HECUDA_ERR_MISSING_KEY, "missingRelinearizationKey on the root rank");
        if ((rc = hecuda_evk_device_buffer(evk, &p, &bytes)) != HECUDA_OK) return rc;
        const int nrc = n.Broadcast(p, p, bytes / 8, ncclUint64, root, comm->comm, comm->stream);
        if (nrc != ncclSuccess) return nccl_fail(nrc, "ncclBroadcast(relinearization key)");
    }
    for (int32_t i = 0; i < element_count; ++i) {
        void *p = nullptr;
        uint64_t bytes = 0;
        if (comm->rank == root) {
            std::lock_guard<std::mutex> g(evk->mu);
            if (evk->galois.find(elements[i]) == evk->galois.end())
                return fail(HECUDA_ERR_MISSING_KEY, "missingGaloisKey on the root rank: element " + std::to_string(elements[i]));
        }
        if ((rc = hecuda_evk_galois_device_buffer(evk, elements[i], &p, &bytes)) != HECUDA_OK) return rc;
        const int nrc = n.Broadcast(p, p, bytes / 8, ncclUint64, root, comm->comm, comm->stream);
        if (nrc != ncclSuccess) return nccl_fail(nrc, "ncclBroadcast(Galois key)");
    }
    CK(cudaStreamSynchronize(comm->stream));
    return HECUDA_OK;
}

}  // extern "C"
