// NCCL through dlopen: the engine's only collective is the one-time weight-arena broadcast.
// Loading lazily keeps libsmelter_b200.so loadable on machines without NCCL or without a GPU.
// The reference has no distributed code at all; this is new.
#include <dlfcn.h>

#include <cstring>

#include "engine.h"

namespace smelter {

namespace {

struct ncclUniqueIdT { char internal[128]; };
typedef int (*fn_get_unique_id)(ncclUniqueIdT*);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, ncclUniqueIdT id, int rank);
typedef int (*fn_comm_destroy)(void* comm);
typedef int (*fn_broadcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t stream);
typedef const char* (*fn_error_string)(int);

struct Nccl {
    void* lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_broadcast broadcast = nullptr;
    fn_error_string error_string = nullptr;
};

Nccl* load() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n.lib ? &n : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib) break;
    }
    if (!n.lib) return nullptr;
    n.get_unique_id = reinterpret_cast<fn_get_unique_id>(dlsym(n.lib, "ncclGetUniqueId"));
    n.comm_init_rank = reinterpret_cast<fn_comm_init_rank>(dlsym(n.lib, "ncclCommInitRank"));
    n.comm_destroy = reinterpret_cast<fn_comm_destroy>(dlsym(n.lib, "ncclCommDestroy"));
    n.broadcast = reinterpret_cast<fn_broadcast>(dlsym(n.lib, "ncclBroadcast"));
    n.error_string = reinterpret_cast<fn_error_string>(dlsym(n.lib, "ncclGetErrorString"));
    if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.broadcast) {
        dlclose(n.lib);
        n.lib = nullptr;
        return nullptr;
    }
    return &n;
}

int nccl_fail(Nccl* n, const char* what, int rc) {
    return fail(SMELTER_ERR_NCCL, std::string(what) + ": " + (n && n->error_string ? n->error_string(rc) : "error") + " (" + std::to_string(rc) + ")");
}

}  // namespace

int nccl_unique_id(uint8_t id[128]) {
    Nccl* n = load();
    if (!n) return fail(SMELTER_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueIdT u;
    int rc = n->get_unique_id(&u);
    if (rc) return nccl_fail(n, "ncclGetUniqueId", rc);
    memcpy(id, u.internal, 128);
    return SMELTER_OK;
}

int nccl_init(Context* ctx, const uint8_t id[128], int rank, int world) {
    Nccl* n = load();
    if (!n) return fail(SMELTER_ERR_NCCL, "libnccl.so.2 could not be loaded");
    if (ctx->nccl_comm) return fail(SMELTER_ERR_INCONSISTENT_STATE, "NCCL already initialised on this context");
    SM_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueIdT u;
    memcpy(u.internal, id, 128);
    void* comm = nullptr;
    int rc = n->comm_init_rank(&comm, world, u, rank);
    if (rc) return nccl_fail(n, "ncclCommInitRank", rc);
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return SMELTER_OK;
}

int nccl_broadcast(Context* ctx, void* buf, size_t bytes, int root, cudaStream_t stream) {
    if (ctx->world <= 1 && !ctx->nccl_comm) return SMELTER_OK;  // single replica: nothing to do
    Nccl* n = load();
    if (!n || !ctx->nccl_comm) return fail(SMELTER_ERR_NCCL, "NCCL not initialised on this context");
    SM_CUDA(cudaSetDevice(ctx->device));
    int rc = n->broadcast(buf, buf, bytes, /*ncclInt8*/ 0, root, ctx->nccl_comm, stream);
    if (rc) return nccl_fail(n, "ncclBroadcast", rc);
    SM_CUDA(cudaStreamSynchronize(stream));
    return SMELTER_OK;
}

void nccl_destroy(Context* ctx) {
    Nccl* n = load();
    if (n && ctx->nccl_comm) n->comm_destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

}  // namespace smelter
