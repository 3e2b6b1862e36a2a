// Runtime: device binding, streams, caching allocator, copies, events.
// Replaces what the reference's OpenCL backend gets from pyopencl: one context + one in-order
// queue + a MemoryPool per device (opencl/device.py:51-115) and blocking enqueue_copy
// (opencl/tensor.py:74-93).  Here copies and kernels are stream-ordered and only D2H blocks.
#include <unordered_set>
#include <chrono>
#include <unistd.h>
#include "lg_common.cuh"
#include <mutex>
#include <unordered_map>
#include <map>
#include <vector>
#include <atomic>
#include <string.h>
#include <cuda_profiler_api.h>

namespace lg {

static thread_local char g_err[1024] = "";
static int g_device = -1;
static cudaStream_t g_stream = nullptr, g_comm = nullptr;
// side stream: work that nothing on the critical path of backward waits for (weight / bias gradients) is
// issued here between lg_side_begin / lg_side_end and runs concurrently with the main stream, backfilling
// the SMs a one-wave GEMM leaves idle.  g_cur is the stream every kernel wrapper launches on.
static cudaStream_t g_side = nullptr, g_cur = nullptr;
static cudaEvent_t g_ev_side_fork = nullptr, g_ev_side_join = nullptr;
static bool g_side_mode = false, g_side_pending = false;
static std::vector<void*> g_side_deferred;   // blocks freed in side mode: reusable only after the join
// lg_comm_compute_begin / end: kernels issued in between go to the collective stream, right behind the
// all-reduce they depend on (per-bucket optimizer updates of the data-parallel wrapper)
static bool g_comm_mode = false;
static std::vector<void*> g_comm_deferred;   // blocks freed in that mode: reusable after lg_nccl_wait
static int g_sms = 148;
static unsigned int* g_errflag_host = nullptr;   // pinned + mapped: kernels store LG_DEVERR_* codes here
static unsigned int* g_errflag_dev = nullptr;
static std::atomic<uint64_t> g_launches{0};

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
cudaStream_t stream() { return g_cur; }
void prefer_gemm_carveout(const void* kernel) {
    // opt-in (LG_GEMM_CARVEOUT=1): measured without effect, see do_init
    static const bool on = getenv("LG_GEMM_CARVEOUT") != nullptr;
    static std::unordered_set<const void*> done;
    if (!on || done.count(kernel)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    done.insert(kernel);
}
unsigned int* error_flag() { return g_errflag_dev; }
// after a synchronisation: report (once) what a kernel flagged since the last check
static int check_device_error() {
    if (!g_errflag_host) return 0;
    const unsigned int code = *(volatile unsigned int*)g_errflag_host;
    if (!code) return 0;
    *(volatile unsigned int*)g_errflag_host = 0;
    if (code == LG_DEVERR_LABEL)
        return set_error("IndexError: a cross-entropy label is out of range for the number of classes "
                         "(that row's loss is NaN and its gradient zero)");
    return set_error("IndexError: an index array holds an entry that is out of bounds for the indexed axis "
                     "(gathered rows were zero-filled, scattered rows skipped)");
}
bool on_side_stream() { return g_side_mode || g_comm_mode; }
int alt_stream_index() { return g_comm_mode ? 2 : (g_side_mode ? 1 : 0); }
cudaStream_t comm_stream() { return g_comm; }
int sm_count() { return g_sms; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- caching allocator ------------------------------------------------------------------------
// Size-class free lists; a freed block is immediately reusable because every consumer runs on the
// one compute stream (stream order == program order).  The comm stream only touches long-lived
// arenas and is joined by events before those could be released.
// Pools: pool 0 serves eager work.  While a step is being captured into a CUDA graph the allocator
// switches to a pool private to that graph; its blocks are baked into the graph as raw addresses, so
// they may only ever be recycled among that graph's own captures (never handed to eager tensors).
struct Cache {
    std::mutex mu;
    struct Block { size_t cls; int pool; };
    std::unordered_map<void*, Block> live;                              // ptr -> class size, pool
    std::map<int, std::map<size_t, std::vector<void*>>> pools;         // pool -> class size -> free blocks
    size_t in_use = 0, reserved = 0, peak = 0;
    int cur_pool = 0, next_pool = 0;

    static size_t size_class(size_t n) {
        if (n < 512) return 512;
        // 8 classes per power of two: waste <= 12.5 %
        size_t p = 512;
        while ((p << 1) <= n) p <<= 1;
        size_t step = p >> 3;
        if (step < 512) step = 512;
        return (n + step - 1) / step * step;
    }
    void release_all() {
        // only the eager pool is trimmed: graph pools hold addresses that instantiated graphs still use
        auto& free_lists = pools[0];
        for (auto& kv : free_lists) {
            for (void* p : kv.second) {
                cudaFree(p);
                reserved -= kv.first;
            }
        }
        free_lists.clear();
    }
    void* get(size_t n) {
        size_t cls = size_class(n);
        std::lock_guard<std::mutex> lk(mu);
        void* p = nullptr;
        auto& free_lists = pools[cur_pool];
        auto it = free_lists.find(cls);
        if (it != free_lists.end() && !it->second.empty()) {
            p = it->second.back();
            it->second.pop_back();
        } else {
            cudaError_t e = cudaMalloc(&p, cls);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cudaStreamSynchronize(g_stream);
                release_all();
                e = cudaMalloc(&p, cls);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    set_error("out of device memory allocating %zu bytes (in use %zu, reserved %zu)", cls, in_use,
                              reserved);
                    return nullptr;
                }
            }
            reserved += cls;
        }
        live[p] = Block{cls, cur_pool};
        in_use += cls;
        if (in_use > peak) peak = in_use;
        return p;
    }
    int put(void* p) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) return set_error("lg_free: unknown pointer %p", p);
        size_t cls = it->second.cls;
        int pool = it->second.pool;
        live.erase(it);
        in_use -= cls;
        pools[pool][cls].push_back(p);
        return 0;
    }
};
static Cache* g_cache = nullptr;
static bool g_capturing = false;
bool capturing() { return g_capturing; }

void* tmp_alloc(size_t nbytes) { return g_cache->get(nbytes ? nbytes : 1); }
void tmp_free(void* p) {
    if (!p) return;
    if (g_comm_mode) g_comm_deferred.push_back(p);
    else if (g_side_mode) g_side_deferred.push_back(p);
    else g_cache->put(p);
}

void comm_defer_free(void* p) {
    if (p) g_comm_deferred.push_back(p);
}

// called once the compute stream has been ordered after the collective stream (lg_nccl_wait)
void comm_release_deferred() {
    for (void* p : g_comm_deferred) g_cache->put(p);
    g_comm_deferred.clear();
}

// main stream waits for everything issued on the side stream so far; deferred blocks become reusable
int side_join() {
    if (!g_side_pending) return 0;
    LG_CUDA(cudaEventRecord(g_ev_side_join, g_side));
    LG_CUDA(cudaStreamWaitEvent(g_stream, g_ev_side_join, 0));
    for (void* p : g_side_deferred) g_cache->put(p);
    g_side_deferred.clear();
    g_side_pending = false;
    return 0;
}
// `other` (the collective stream) must also see side-stream writes issued so far
int side_order_before(cudaStream_t other) {
    if (!g_side_pending) return 0;
    LG_CUDA(cudaEventRecord(g_ev_side_join, g_side));
    LG_CUDA(cudaStreamWaitEvent(other, g_ev_side_join, 0));
    return 0;
}

static int do_init(int device) {
    if (g_device >= 0) {
        if (device >= 0 && device != g_device)
            return set_error("lg_init: process already bound to device %d (asked for %d)", g_device, device);
        return 0;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error("no CUDA device available (%s); lightgrad_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0) {
        const char* lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % n : 0;
    }
    if (device >= n) return set_error("lg_init: device %d out of range (%d visible)", device, n);
    LG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    LG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error("lightgrad_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                         prop.minor);
    g_sms = prop.multiProcessorCount;
    {
        // Hypothesis tested in r2: the GEMM / attention CTAs need the largest shared-memory carve-out, a kernel without
        // shared memory prefers the largest L1, and an SM can only change its split when it is empty -- so a small kernel
        // of another stream (column sums on the side stream, the exchange kernel) might keep the next GEMM's CTAs off its
        // SMs.  Measured (profiles/r2_bench_carveout_*.json): preferring shared memory DEVICE-WIDE (LG_PREFER_SHARED=1)
        // costs the L1-reliant kernels 0.2-0.3 ms per step (8.28 vs 8.02 ms on one GPU; at 2 GPUs 8.58-8.62 vs 8.49-8.51);
        // asking for the GEMM's carve-out only on the kernels that run beside GEMMs (LG_GEMM_CARVEOUT=1:
        // prefer_gemm_carveout) changes nothing (7.96 vs 7.99 ms; 2 GPUs 8.50-8.57 vs 8.48-8.50).  Both stay off.
        static const bool all = getenv("LG_PREFER_SHARED") != nullptr;
        if (all) LG_CUDA(cudaDeviceSetCacheConfig(cudaFuncCachePreferShared));
    }
    LG_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    {
        // the collective stream gets the highest priority: its few small CTAs (NCCL, or the multicast exchange kernel)
        // are placed ahead of the next compute kernel's CTAs whenever an SM has room
        int lo_prio = 0, hi_prio = 0;
        LG_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        static const bool flat = getenv("LG_COMM_NO_PRIORITY") != nullptr;
        LG_CUDA(cudaStreamCreateWithPriority(&g_comm, cudaStreamNonBlocking, flat ? lo_prio : hi_prio));
    }
    LG_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
    LG_CUDA(cudaEventCreateWithFlags(&g_ev_side_fork, cudaEventDisableTiming));
    LG_CUDA(cudaEventCreateWithFlags(&g_ev_side_join, cudaEventDisableTiming));
    g_cur = g_stream;
    LG_CUDA(cudaHostAlloc((void**)&g_errflag_host, sizeof(unsigned int), cudaHostAllocMapped));
    *g_errflag_host = 0;
    LG_CUDA(cudaHostGetDevicePointer((void**)&g_errflag_dev, g_errflag_host, 0));
    g_cache = new Cache();
    g_device = device;
    return 0;
}

int ensure_init() { return do_init(-1); }

}  // namespace lg

using namespace lg;

namespace {
// occupies the compute stream for `ns` nanoseconds: lets the host queue a whole step behind it, so that
// CUDA events recorded between the queued kernels time the kernels and not the host's dispatch latency
__global__ void spin_kernel(unsigned long long ns) {
    LG_PDL_TRIGGER();
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}
}  // namespace

extern "C" {

int lg_stream_delay_us(uint64_t us) {
    LG_INIT();
    LG_REQUIRE(us <= 2000000, "lg_stream_delay_us: at most 2 s");
    spin_kernel<<<1, 1, 0, g_stream>>>((unsigned long long)us * 1000ull);
    LG_CHECK_LAUNCH();
    return 0;
}

const char* lg_last_error(void) { return g_err; }

int lg_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

int lg_init(int device) { return do_init(device); }

int lg_device(int* device) {
    *device = g_device;
    return g_device >= 0 ? 0 : set_error("not initialised");
}

int lg_device_props(int* sm_count_, int* cc_major, int* cc_minor, size_t* total_mem) {
    LG_INIT();
    cudaDeviceProp prop;
    LG_CUDA(cudaGetDeviceProperties(&prop, g_device));
    *sm_count_ = prop.multiProcessorCount;
    *cc_major = prop.major;
    *cc_minor = prop.minor;
    *total_mem = prop.totalGlobalMem;
    return 0;
}

// Synchronise a stream.  With a NCCL communicator alive the wait is watched: a collective whose peer has died spins on
// the GPU for ever, so the stream is polled, NCCL's asynchronous error state is checked, and after LG_SYNC_TIMEOUT_S
// (default 120; 0 = wait for ever) the communicator is aborted -- its kernels leave the GPU -- and the call fails instead
// of hanging the rank (fail fast on rank death).  The multicast exchange kernels bound their own
// waits (20 s, then trap).  Single-GPU processes take the plain cudaStreamSynchronize.
static int watched_sync(cudaStream_t st) {
    if (!nccl_active()) {
        LG_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    static const double limit = getenv("LG_SYNC_TIMEOUT_S") ? atof(getenv("LG_SYNC_TIMEOUT_S")) : 120.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned long spins = 0;; ++spins) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return 0;
        if (q != cudaErrorNotReady) return set_error("cudaStreamQuery failed: %s", cudaGetErrorString(q));
        if ((spins & 1023) == 1023) {
            if (nccl_async_check()) return 1;
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (limit > 0 && waited > limit) {
                nccl_abort("timeout");
                cudaStreamSynchronize(st);       // the aborted collective's kernels exit
                return set_error("synchronisation did not finish within %.0f s (LG_SYNC_TIMEOUT_S): a peer rank is gone "
                                 "or stuck; the NCCL communicator was aborted", limit);
            }
            if (waited > 0.05) usleep(200);       // long waits: stop burning the core
        }
    }
}

int lg_sync(void) {
    LG_INIT();
    LG_REQUIRE(!g_capturing, "lg_sync: cannot synchronise while a step is being captured into a CUDA graph");
    if (side_join()) return 1;
    if (watched_sync(g_stream)) return 1;
    if (watched_sync(g_comm)) return 1;
    return check_device_error();
}

int lg_alloc(size_t nbytes, void** ptr) {
    LG_INIT();
    void* p = g_cache->get(nbytes ? nbytes : 1);
    if (!p) return 1;
    *ptr = p;
    return 0;
}

int lg_free(void* ptr) {
    if (!ptr || !g_cache) return 0;
    if (g_comm_mode) {
        g_comm_deferred.push_back(ptr);
        return 0;
    }
    if (g_side_mode) {
        g_side_deferred.push_back(ptr);
        return 0;
    }
    return g_cache->put(ptr);
}

int lg_comm_compute_begin(void) {
    LG_INIT();
    LG_REQUIRE(!g_side_mode && !g_comm_mode, "lg_comm_compute_begin: already on another stream");
    g_cur = g_comm;
    g_comm_mode = true;
    return 0;
}

int lg_comm_compute_end(void) {
    g_cur = g_stream;
    g_comm_mode = false;
    return 0;
}

// ---- side stream ---------------------------------------------------------------------------------
// lg_side_begin .. lg_side_end brackets launches that go to the side stream (ordered after everything issued
// on the main stream so far); lg_side_join makes the main stream wait for them.  The CALLER keeps every
// buffer those launches touch alive until the join.  LG_NO_SIDE_STREAM=1 turns the bracket into a no-op.
int lg_side_begin(void) {
    LG_INIT();
    static const bool off = getenv("LG_NO_SIDE_STREAM") != nullptr;
    if (off) return 0;
    LG_REQUIRE(!g_side_mode && !g_comm_mode, "lg_side_begin: already on another stream");
    LG_CUDA(cudaEventRecord(g_ev_side_fork, g_stream));
    LG_CUDA(cudaStreamWaitEvent(g_side, g_ev_side_fork, 0));
    g_cur = g_side;
    g_side_mode = true;
    g_side_pending = true;
    return 0;
}

int lg_side_end(void) {
    g_cur = g_stream;
    g_side_mode = false;
    return 0;
}

int lg_side_join(void) {
    LG_INIT();
    LG_REQUIRE(!g_side_mode, "lg_side_join: still inside lg_side_begin / lg_side_end");
    return side_join();
}

int lg_empty_cache(void) {
    LG_INIT();
    if (side_join()) return 1;
    LG_CUDA(cudaStreamSynchronize(g_stream));
    std::lock_guard<std::mutex> lk(g_cache->mu);
    g_cache->release_all();
    return 0;
}

int lg_mem_stats(size_t* in_use, size_t* reserved, size_t* peak_in_use) {
    LG_INIT();
    std::lock_guard<std::mutex> lk(g_cache->mu);
    *in_use = g_cache->in_use;
    *reserved = g_cache->reserved;
    *peak_in_use = g_cache->peak;
    return 0;
}

int lg_memcpy_h2d(void* dst, const void* src, size_t nbytes) {
    LG_INIT();
    if (nbytes == 0) return 0;
    LG_REQUIRE(!g_capturing, "lg_memcpy_h2d: host copies cannot be captured into a CUDA graph; stage inputs in "
                             "device tensors before capture and refresh them between replays");
    LG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, g_stream));
    return 0;
}

int lg_memcpy_d2h(void* dst, const void* src, size_t nbytes) {
    LG_INIT();
    LG_REQUIRE(!g_capturing, "lg_memcpy_d2h: reading a tensor back (numpy()/item()) is not possible while a step is "
                             "being captured into a CUDA graph");
    if (side_join()) return 1;
    if (nbytes) LG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, g_stream));
    if (watched_sync(g_stream)) return 1;
    return check_device_error();
}

int lg_memcpy_d2d(void* dst, const void* src, size_t nbytes) {
    LG_INIT();
    if (nbytes == 0) return 0;
    LG_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, g_cur));
    return 0;
}

int lg_memset(void* dst, int byte, size_t nbytes) {
    LG_INIT();
    if (nbytes == 0) return 0;
    LG_CUDA(cudaMemsetAsync(dst, byte, nbytes, g_cur));
    return 0;
}

int lg_host_alloc(size_t nbytes, void** ptr) {
    LG_INIT();
    LG_CUDA(cudaHostAlloc(ptr, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return 0;
}

int lg_host_free(void* ptr) {
    if (ptr) LG_CUDA(cudaFreeHost(ptr));
    return 0;
}

int lg_event_create(void** ev) {
    LG_INIT();
    cudaEvent_t e;
    LG_CUDA(cudaEventCreate(&e));
    *ev = (void*)e;
    return 0;
}
int lg_event_record(void* ev) {
    LG_CUDA(cudaEventRecord((cudaEvent_t)ev, g_stream));
    return 0;
}
int lg_event_sync(void* ev) {
    LG_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
    return 0;
}
int lg_event_elapsed_ms(void* start, void* stop, float* ms) {
    LG_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}
int lg_event_destroy(void* ev) {
    LG_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return 0;
}

int lg_launch_count(uint64_t* n) {
    *n = g_launches.load();
    return 0;
}

void* lg_stream_handle(void) { return (void*)g_stream; }

// ---- whole-step CUDA graphs ----------------------------------------------------------------------
int lg_graph_begin(int* pool_id) {
    LG_INIT();
    LG_REQUIRE(!g_capturing, "lg_graph_begin: a capture is already in progress");
    if (side_join()) return 1;
    LG_CUDA(cudaStreamSynchronize(g_stream));
    {
        std::lock_guard<std::mutex> lk(g_cache->mu);
        if (*pool_id <= 0) *pool_id = ++g_cache->next_pool;
        g_cache->cur_pool = *pool_id;
    }
    cudaError_t e = cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) {
        g_cache->cur_pool = 0;
        return set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(e));
    }
    g_capturing = true;
    return 0;
}

int lg_graph_end(void** graph_exec, uint64_t* n_nodes) {
    LG_REQUIRE(g_capturing, "lg_graph_end: no capture in progress");
    // a forked side stream must rejoin the capturing stream before the capture can end
    g_cur = g_stream;
    g_side_mode = false;
    g_comm_mode = false;
    if (side_join()) return 1;
    g_capturing = false;
    {
        std::lock_guard<std::mutex> lk(g_cache->mu);
        g_cache->cur_pool = 0;
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        return set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    }
    size_t nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &nodes);
    *n_nodes = nodes;
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    *graph_exec = (void*)exec;
    return 0;
}

int lg_graph_abort(void) {
    if (!g_capturing) return 0;
    g_cur = g_stream;
    g_side_mode = false;
    g_comm_mode = false;
    side_join();
    g_capturing = false;
    g_cache->cur_pool = 0;
    cudaGraph_t graph = nullptr;
    cudaStreamEndCapture(g_stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return 0;
}

int lg_graph_launch(void* graph_exec, uint64_t n_kernels) {
    LG_REQUIRE(!g_capturing, "lg_graph_launch: cannot replay during a capture");
    LG_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, g_stream));
    count_launch((int)n_kernels);
    return 0;
}

int lg_graph_destroy(void* graph_exec) {
    if (graph_exec) LG_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return 0;
}

int lg_profiler_range(int start) {
    LG_INIT();
    LG_CUDA(cudaStreamSynchronize(g_stream));
    if (start) LG_CUDA(cudaProfilerStart());
    else LG_CUDA(cudaProfilerStop());
    return 0;
}

}  // extern "C"
