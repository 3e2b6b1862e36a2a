// Gradient all-reduce over NCCL (NVLink 5 / NVSwitch), one process per GPU.
// The reference has no collectives at all (cross-device operands are rejected, func.py:18-20);
// this is the exchange step of the data-parallel wrapper.  libnccl is dlopen()ed so that the
// library loads on hosts without NCCL; the symbols are resolved on first use.
#include "lg_common.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <string.h>

using namespace lg;

namespace {

struct Api {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    // optional (fail-fast on rank death): absent symbols only disable the check
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
} api;

ncclComm_t g_comm_nccl = nullptr;
int g_world = 1, g_rank = 0;
cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;

int load_api() {
    if (api.h) return 0;
    const char* names[] = {getenv("LG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return set_error("cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
    *(void**)(&api.field) = dlsym(api.h, name);                                  \
    if (!api.field) return set_error("libnccl is missing symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Broadcast, "ncclBroadcast")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    *(void**)(&api.CommGetAsyncError) = dlsym(api.h, "ncclCommGetAsyncError");
    *(void**)(&api.CommAbort) = dlsym(api.h, "ncclCommAbort");
    return 0;
}

#define LG_NCCL(expr)                                                                              \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) return set_error("%s failed: %s", #expr, api.GetErrorString(_r));   \
    } while (0)

}  // namespace

namespace lg {

bool nccl_active() { return g_comm_nccl != nullptr; }

void nccl_abort(const char* why) {
    if (!g_comm_nccl) return;
    // ncclCommAbort raises the communicator's abort flag: collective kernels still spinning for a peer leave the GPU
    if (api.CommAbort) api.CommAbort(g_comm_nccl);
    g_comm_nccl = nullptr;
    (void)why;
}

int nccl_async_check() {
    if (!g_comm_nccl || !api.CommGetAsyncError) return 0;
    ncclResult_t state = ncclSuccess;
    if (api.CommGetAsyncError(g_comm_nccl, &state) != ncclSuccess || state == ncclSuccess || state == ncclInProgress)
        return 0;
    set_error("NCCL reported an asynchronous error (%s): a peer rank died or a link failed; the communicator was aborted",
              api.GetErrorString ? api.GetErrorString(state) : "?");
    nccl_abort("async error");
    return 1;
}

}  // namespace lg

static int ensure_events() {
    if (!g_ev_fork) LG_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
    if (!g_ev_join) LG_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
    return 0;
}

extern "C" {

int lg_nccl_unique_id(void* id128) {
    if (load_api()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    LG_NCCL(api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int lg_nccl_init(const void* id128, int world, int rank) {
    LG_INIT();
    if (load_api()) return 1;
    LG_REQUIRE(!g_comm_nccl, "lg_nccl_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    LG_NCCL(api.CommInitRank(&g_comm_nccl, world, id, rank));
    g_world = world;
    g_rank = rank;
    return ensure_events();
}

int lg_nccl_fork(void) {
    LG_INIT();
    if (ensure_events()) return 1;
    LG_CUDA(cudaEventRecord(g_ev_fork, stream()));
    LG_CUDA(cudaStreamWaitEvent(comm_stream(), g_ev_fork, 0));
    // weight gradients are accumulated on the side stream: the collective must see those writes too
    if (side_order_before(comm_stream())) return 1;
    return 0;
}

int lg_nccl_wait(void) {
    LG_INIT();
    if (ensure_events()) return 1;
    LG_CUDA(cudaEventRecord(g_ev_join, comm_stream()));
    LG_CUDA(cudaStreamWaitEvent(stream(), g_ev_join, 0));
    comm_release_deferred();
    return 0;
}

int lg_nccl_allreduce_f32(void* buf, int64_t n, int op, int on_comm_stream) {
    LG_REQUIRE(g_comm_nccl, "lg_nccl_allreduce_f32: communicator not initialised");
    if (n == 0) return 0;
    LG_NCCL(api.AllReduce(buf, buf, (size_t)n, ncclFloat, op == 1 ? ncclAvg : (op == 2 ? ncclMax : ncclSum), g_comm_nccl,
                          on_comm_stream ? comm_stream() : stream()));
    count_launch();
    return 0;
}

int lg_nccl_broadcast(void* buf, int64_t nbytes, int root) {
    LG_REQUIRE(g_comm_nccl, "lg_nccl_broadcast: communicator not initialised");
    if (nbytes == 0) return 0;
    LG_NCCL(api.Broadcast(buf, buf, (size_t)nbytes, ncclChar, root, g_comm_nccl, stream()));
    count_launch();
    return 0;
}

int lg_nccl_destroy(void) {
    if (!g_comm_nccl) return 0;
    cudaStreamSynchronize(comm_stream());
    cudaStreamSynchronize(stream());
    LG_NCCL(api.CommDestroy(g_comm_nccl));
    g_comm_nccl = nullptr;
    return 0;
}

}  // extern "C"
