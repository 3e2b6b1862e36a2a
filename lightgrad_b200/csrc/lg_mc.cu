// Gradient exchange fused with the optimizer over NVLink multicast (NVSwitch "NVLS"), one process per GPU.
//
// The reference has no multi-device path; its optimizer is the per-parameter python loop of lightgrad/optim.py:27-41,
// whose arithmetic (Adam with one step-counter increment per parameter, AdaBelief, SGD) is what the kernel below
// applies.  What it replaces on the data-parallel path is "ncclAllReduce(avg) over the gradient arena, then one Adam
// pass over all parameters" (parallel.py): every rank owns 1/world of each bucket and, in ONE kernel,
//
//     g   = multimem.ld_reduce.add(grad arena)      the switch sums that range over all GPUs and returns it (reduce-scatter)
//     p   = adam(p, g / world, m, v)                optimizer state exists only for the owned range
//     multimem.st(param arena) = p                  the switch writes the new values into every GPU's copy (all-gather)
//
// so each gradient element crosses NVLink once into the switch and each new parameter once out of it, the optimizer's
// HBM traffic per GPU drops to 1/world, and there is no separate all-reduce pass.  The kernel uses no shared memory and
// few registers, so its CTAs become resident NEXT TO the persistent one-CTA-per-SM tensor-core GEMMs of backward (which
// fill the shared memory, not the register file or the thread slots): the exchange of a bucket really overlaps the rest
// of backward, which NCCL's own kernels (large CTAs that need a free SM) could not do.
//
// Memory: both arenas and a page of flags live in ONE physical allocation per GPU (cuMemCreate), bound to a multicast
// object shared by all ranks (cuMulticastCreate on rank 0, POSIX file descriptor passed to the other ranks by the host
// layer over a unix socket) and mapped twice: a normal ("local") mapping and the multicast mapping.
// Cross-GPU ordering: per-CTA flags in that region; arrive = multimem.red.add (every GPU's copy is incremented), wait =
// acquire-load of the local copy until world x epoch arrivals; spins are bounded by the global timer and trap.
#include "lg_adam.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <unistd.h>

using namespace lg;
using namespace lg::adam;

namespace {

constexpr int MC_MAX_CTAS = 1024;
constexpr int MC_THREADS = 128;
constexpr size_t MC_FLAG_BYTES = 2 * MC_MAX_CTAS * sizeof(unsigned int);
constexpr int TRACE_MAX = 4096;

struct Drv {
    bool loaded = false;
    PFN_cuDeviceGet cuDeviceGet = nullptr;
    PFN_cuDeviceGetAttribute cuDeviceGetAttribute = nullptr;
    PFN_cuMulticastCreate cuMulticastCreate = nullptr;
    PFN_cuMulticastAddDevice cuMulticastAddDevice = nullptr;
    PFN_cuMulticastBindMem cuMulticastBindMem = nullptr;
    PFN_cuMulticastUnbind cuMulticastUnbind = nullptr;
    PFN_cuMulticastGetGranularity cuMulticastGetGranularity = nullptr;
    PFN_cuMemCreate cuMemCreate = nullptr;
    PFN_cuMemRelease cuMemRelease = nullptr;
    PFN_cuMemMap cuMemMap = nullptr;
    PFN_cuMemUnmap cuMemUnmap = nullptr;
    PFN_cuMemSetAccess cuMemSetAccess = nullptr;
    PFN_cuMemAddressReserve cuMemAddressReserve = nullptr;
    PFN_cuMemAddressFree cuMemAddressFree = nullptr;
    PFN_cuMemExportToShareableHandle cuMemExportToShareableHandle = nullptr;
    PFN_cuMemImportFromShareableHandle cuMemImportFromShareableHandle = nullptr;
    PFN_cuMemGetAllocationGranularity cuMemGetAllocationGranularity = nullptr;
} drv;

int load_driver() {
    if (drv.loaded) return 0;
#define SYM(name)                                                                                       \
    {                                                                                                   \
        void* fn = nullptr;                                                                             \
        cudaDriverEntryPointQueryResult q;                                                              \
        cudaError_t e = cudaGetDriverEntryPoint(#name, &fn, cudaEnableDefault, &q);                     \
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {                              \
            cudaGetLastError();                                                                         \
            return set_error("driver entry point %s is not available", #name);                          \
        }                                                                                               \
        drv.name = (PFN_##name)fn;                                                                      \
    }
    SYM(cuDeviceGet) SYM(cuDeviceGetAttribute) SYM(cuMulticastCreate) SYM(cuMulticastAddDevice) SYM(cuMulticastBindMem)
    SYM(cuMulticastUnbind) SYM(cuMulticastGetGranularity) SYM(cuMemCreate) SYM(cuMemRelease) SYM(cuMemMap)
    SYM(cuMemUnmap) SYM(cuMemSetAccess) SYM(cuMemAddressReserve) SYM(cuMemAddressFree)
    SYM(cuMemExportToShareableHandle) SYM(cuMemImportFromShareableHandle) SYM(cuMemGetAllocationGranularity)
#undef SYM
    drv.loaded = true;
    return 0;
}

#define LG_DRV(expr)                                                                                   \
    do {                                                                                               \
        CUresult _r = (expr);                                                                          \
        if (_r != CUDA_SUCCESS) return set_error("%s failed with CUresult %d (%s:%d)", #expr, (int)_r, __FILE__, __LINE__); \
    } while (0)

struct Region {
    bool have_mc = false, bound = false;
    int world = 0;
    CUdevice dev = 0;
    CUmemGenericAllocationHandle mc = 0, mem = 0;
    size_t bytes = 0;            // size of the multicast object = of the physical allocation
    CUdeviceptr local = 0, mcva = 0;
    unsigned int* epochs = nullptr;   // plain device memory: launches seen so far, per CTA index
    int grid = 0;                // CTAs of every exchange launch (identical on all ranks)
    unsigned long long* trace = nullptr;   // LG_MC_TRACE=1: 4 x u64 per record, TRACE_MAX records (device memory)
    int trace_n = 0;
} rg;

CUmulticastObjectProp mc_prop(size_t bytes, int world) {
    CUmulticastObjectProp p = {};
    p.numDevices = (unsigned)world;
    p.size = bytes;
    p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    p.flags = 0;
    return p;
}

int current_device(CUdevice* dev) {
    int ord = 0;
    LG_CUDA(cudaGetDevice(&ord));
    LG_DRV(drv.cuDeviceGet(dev, ord));
    return 0;
}

// ---- device side ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mc_ld_reduce(const float* mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc_addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc_addr, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void mc_arrive(unsigned int* mc_flag) {
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_flag), "r"(1u) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// thread 0 of the CTA: wait until every rank's CTA of this index has arrived `epoch` times (bounded: a peer that
// died must surface as a failed launch, not as a hung GPU)
__device__ __forceinline__ void mc_wait(const unsigned int* local_flag, unsigned int target) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(local_flag) - target) < 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 20000000000ull) __trap();     // 20 s
        __nanosleep(64);
    }
}

int stream_hint() {
    static const int on = getenv("LG_MC_NO_STREAM_HINT") ? 0 : 1;
    return on;
}

struct McArgs {
    const float* grad;         // LOCAL variant: this GPU's gradient arena (plain mapping)
    float* mc_grad;            // multicast mapping of the gradient arena
    float* mc_param;           // multicast mapping of the parameter arena
    float* param;              // local mapping of the parameter arena
    float* m;                  // optimizer state (local, full-size arrays; only the owned ranges are touched)
    float* v;
    unsigned int* mc_flags;    // [2][MC_MAX_CTAS]: start / done arrivals, multicast mapping
    unsigned int* flags;       // the same words through the local mapping
    unsigned int* epochs;      // [MC_MAX_CTAS] launches seen so far by the CTA of that index (plain device memory)
    int64_t lo, hi;            // element range of this bucket inside the arenas (multiples of 4)
    int rank, world;
    int kind;                  // 0 Adam, 1 AdaBelief, 2 SGD (m = previous delta when momentum != 0), 3 dry run (p unchanged)
    int n_seg;
    const int64_t* seg_end;    // arena offsets (exclusive ends) of this bucket's tensors
    const float* c1;           // per-tensor bias corrections (adam_prep_kernel): c1 | c2 | 1/c1 | 1/c2
    const float* c2;
    float neg_lr, b1, b2, omb1, omb2, eps, inv_world, momentum;
    unsigned long long* trace;  // LG_MC_TRACE: {entered, all ranks met, finished} of CTA 0 in globaltimer ns, or nullptr
};

// The optimizer operands are touched once per step.  Read and written with the streaming (evict-first) hint they do not
// push the operand tiles of the GEMMs that run beside this kernel out of L2 (LG_MC_NO_STREAM_HINT=1: plain accesses,
// for the A/B measurement).
template <bool STREAM>
__device__ __forceinline__ float4 ld_state(const float4* p) { return STREAM ? __ldcs(p) : *p; }
template <bool STREAM>
__device__ __forceinline__ void st_state(float4* p, const float4& v) {
    if (STREAM) __stcs(p, v);
    else *p = v;
}

// a time stamp on whatever stream it is launched on (LG_MC_TRACE: where the compute stream is while buckets run)
__global__ void mc_mark_kernel(unsigned long long* slot) {
    LG_PDL_TRIGGER();
    *slot = globaltimer_ns();
}

// LOCAL = true: the same small-footprint kernel for ONE GPU -- no switch, no flags: the gradients are read from the
// local arena and the parameters written back to it.  What it keeps is the reason to exist: 128 threads, no shared
// memory, <= 88 registers, so it runs beside the GEMMs of backward and the optimizer leaves the critical path.
template <int KIND, bool LOCAL, bool STREAM>
__global__ void __launch_bounds__(MC_THREADS) mc_exchange_kernel(const McArgs a) {
    LG_PDL_TRIGGER();
    // ---- every rank's gradients of this bucket are final once its kernel runs (stream order on that rank):
    //      meet the CTAs of this index on all ranks
    // (no shared memory at all: next to a tensor-core GEMM CTA an SM has room for the 1 KB every CTA reserves, not more)
    unsigned int epoch = 0;                                  // thread 0 only
    if (!LOCAL && threadIdx.x == 0) {
        epoch = a.epochs[blockIdx.x] + 1;
        a.epochs[blockIdx.x] = epoch;
        if (a.trace && blockIdx.x == 0) {
            a.trace[0] = globaltimer_ns();
            a.trace[3] = (unsigned long long)(a.hi - a.lo) * 4;
        }
        __threadfence_system();                              // this rank's earlier writes precede its arrival
        mc_arrive(a.mc_flags + blockIdx.x);
        mc_wait(a.flags + blockIdx.x, epoch * (unsigned)a.world);
        if (a.trace && blockIdx.x == 0) a.trace[1] = globaltimer_ns();
    }
    __syncthreads();
    // ---- this rank's share of the bucket: vectors [v0, v1)
    const int64_t nv = (a.hi - a.lo) >> 2;
    const int64_t per = (nv + a.world - 1) / a.world;
    const int64_t v0 = a.rank * per < nv ? a.rank * per : nv;
    const int64_t v1 = v0 + per < nv ? v0 + per : nv;
    const int64_t base = a.lo >> 2;
    const int64_t stride = (int64_t)gridDim.x * MC_THREADS;
    // independent 16-byte switch reductions in flight per thread; chosen so that the kernel stays under the ~88
    // registers per thread that are free on an SM beside a tensor-core GEMM CTA (128 threads x 88 = 11.3 K of 11.7 K)
    constexpr int U = KIND == 2 ? 4 : (LOCAL ? 6 : 8);
    int seg = 0;
    for (int64_t i0 = v0 + (int64_t)blockIdx.x * MC_THREADS + threadIdx.x; i0 < v1; i0 += stride * U) {
        // the switch reductions have the long latency (NVLink round trip): all U are issued first; the local
        // operands of a vector are fetched when its turn comes, which keeps the register footprint small enough for
        // this CTA to sit next to a tensor-core GEMM CTA (11.7 K registers are left per SM beside one)
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < v1) {
                if (LOCAL) g[u] = ld_state<STREAM>(reinterpret_cast<const float4*>(a.grad) + base + i);
                else g[u] = mc_ld_reduce(a.mc_grad + ((base + i) << 2));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= v1) break;
            float4 p = ld_state<STREAM>(reinterpret_cast<const float4*>(a.param) + base + i);
            float4 gi = g[u];
            gi.x *= a.inv_world; gi.y *= a.inv_world; gi.z *= a.inv_world; gi.w *= a.inv_world;
            if (KIND <= 1) {
                float4 mm = ld_state<STREAM>(reinterpret_cast<const float4*>(a.m) + base + i);
                float4 vv = ld_state<STREAM>(reinterpret_cast<const float4*>(a.v) + base + i);
                seg = adam_segment(seg, (base + i) << 2, a.n_seg, a.seg_end);
                adam_update4<KIND == 1>(p, gi, mm, vv, a.c1[seg], a.c2[seg], a.c1[2 * a.n_seg + seg],
                                        a.c2[2 * a.n_seg + seg], a.neg_lr, a.b1, a.b2, a.omb1, a.omb2, a.eps);
                st_state<STREAM>(reinterpret_cast<float4*>(a.m) + base + i, mm);
                st_state<STREAM>(reinterpret_cast<float4*>(a.v) + base + i, vv);
            } else if (KIND == 2) {
                float4 d;
                if (a.momentum != 0.f) {
                    const float4 mm = reinterpret_cast<const float4*>(a.m)[base + i];
                    d.x = a.neg_lr * gi.x + a.momentum * mm.x; d.y = a.neg_lr * gi.y + a.momentum * mm.y;
                    d.z = a.neg_lr * gi.z + a.momentum * mm.z; d.w = a.neg_lr * gi.w + a.momentum * mm.w;
                    reinterpret_cast<float4*>(a.m)[base + i] = d;
                } else {
                    d.x = a.neg_lr * gi.x; d.y = a.neg_lr * gi.y; d.z = a.neg_lr * gi.z; d.w = a.neg_lr * gi.w;
                }
                p.x += d.x; p.y += d.y; p.z += d.z; p.w += d.w;
            }
            if (LOCAL) reinterpret_cast<float4*>(a.param)[base + i] = p;
            else mc_st(a.mc_param + ((base + i) << 2), p);   // every GPU's copy of the parameters, this one included
        }
    }
    // ---- done: nobody may reuse the gradients (next step's zero fill) or read the parameters (next forward) of
    //      this bucket before all ranks have finished reading / writing them
    __syncthreads();
    if (!LOCAL && threadIdx.x == 0) {
        __threadfence_system();
        mc_arrive(a.mc_flags + MC_MAX_CTAS + blockIdx.x);
        mc_wait(a.flags + MC_MAX_CTAS + blockIdx.x, epoch * (unsigned)a.world);
        if (a.trace && blockIdx.x == 0) a.trace[2] = globaltimer_ns();
    }
}

}  // namespace

extern "C" {

int lg_mc_supported(int* yes) {
    *yes = 0;
    LG_INIT();
    if (load_driver()) return 0;      // not an error: the caller falls back to NCCL
    CUdevice dev;
    if (current_device(&dev)) return 0;
    int v = 0;
    if (drv.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
    *yes = v ? 1 : 0;
    return 0;
}

// bytes needed for two arenas of `arena_bytes` each plus the flag page, rounded to the multicast granularity
int lg_mc_region_bytes(size_t arena_bytes, int world, size_t* region_bytes, size_t* grad_offset, size_t* param_offset,
                       size_t* flag_offset) {
    LG_INIT();
    if (load_driver()) return 1;
    LG_REQUIRE(world >= 1, "lg_mc_region_bytes: world must be >= 1");
    CUmulticastObjectProp prop = mc_prop(0, world);
    size_t gran = 0;
    LG_DRV(drv.cuMulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    CUdevice dev;
    if (current_device(&dev)) return 1;
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = dev;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mgran = 0;
    LG_DRV(drv.cuMemGetAllocationGranularity(&mgran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (mgran > gran) gran = mgran;
    auto up = [&](size_t x) { return (x + gran - 1) / gran * gran; };
    const size_t a = up(arena_bytes);
    *grad_offset = 0;
    *param_offset = a;
    *flag_offset = 2 * a;
    *region_bytes = 2 * a + up(MC_FLAG_BYTES);
    return 0;
}

// rank 0: create the multicast object and hand out a POSIX file descriptor for the other ranks
int lg_mc_create(size_t region_bytes, int world, int* fd) {
    LG_INIT();
    if (load_driver()) return 1;
    LG_REQUIRE(!rg.have_mc, "lg_mc_create: a multicast region already exists in this process");
    CUmulticastObjectProp prop = mc_prop(region_bytes, world);
    LG_DRV(drv.cuMulticastCreate(&rg.mc, &prop));
    int h = -1;
    LG_DRV(drv.cuMemExportToShareableHandle(&h, rg.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd = h;
    rg.have_mc = true;
    rg.world = world;
    rg.bytes = region_bytes;
    return 0;
}

// other ranks: import the object from the descriptor received from rank 0 (the descriptor may be closed afterwards)
int lg_mc_import(int fd, size_t region_bytes, int world) {
    LG_INIT();
    if (load_driver()) return 1;
    LG_REQUIRE(!rg.have_mc, "lg_mc_import: a multicast region already exists in this process");
    LG_DRV(drv.cuMemImportFromShareableHandle(&rg.mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    rg.have_mc = true;
    rg.world = world;
    rg.bytes = region_bytes;
    return 0;
}

// every rank: join the team.  ALL ranks must have returned from this (host barrier) before any calls lg_mc_bind.
int lg_mc_add_device(void) {
    LG_REQUIRE(rg.have_mc, "lg_mc_add_device: no multicast object");
    if (current_device(&rg.dev)) return 1;
    LG_DRV(drv.cuMulticastAddDevice(rg.mc, rg.dev));
    return 0;
}

// every rank: allocate this GPU's physical memory, map it, bind it to the multicast object and map the object.
// local_ptr / mc_ptr address the same bytes on this GPU; a store through mc_ptr lands on every GPU of the team.
int lg_mc_bind(void** local_ptr, void** mc_ptr) {
    LG_REQUIRE(rg.have_mc && !rg.bound, "lg_mc_bind: no multicast object, or already bound");
    CUmemAllocationProp ap = {};
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = rg.dev;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    LG_DRV(drv.cuMemCreate(&rg.mem, rg.bytes, &ap, 0));
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = rg.dev;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    LG_DRV(drv.cuMemAddressReserve(&rg.local, rg.bytes, 0, 0, 0));
    LG_DRV(drv.cuMemMap(rg.local, rg.bytes, 0, rg.mem, 0));
    LG_DRV(drv.cuMemSetAccess(rg.local, rg.bytes, &acc, 1));
    LG_DRV(drv.cuMulticastBindMem(rg.mc, 0, rg.mem, 0, rg.bytes, 0));
    LG_DRV(drv.cuMemAddressReserve(&rg.mcva, rg.bytes, 0, 0, 0));
    LG_DRV(drv.cuMemMap(rg.mcva, rg.bytes, 0, rg.mc, 0));
    LG_DRV(drv.cuMemSetAccess(rg.mcva, rg.bytes, &acc, 1));
    LG_CUDA(cudaMemsetAsync((void*)rg.local, 0, rg.bytes, stream()));
    LG_CUDA(cudaMalloc((void**)&rg.epochs, MC_MAX_CTAS * sizeof(unsigned int)));
    LG_CUDA(cudaMemsetAsync(rg.epochs, 0, MC_MAX_CTAS * sizeof(unsigned int), stream()));
    LG_CUDA(cudaStreamSynchronize(stream()));
    // one CTA per SM by default: small enough to be resident next to whatever else runs, enough threads in flight
    // to cover the switch round trip (LG_MC_CTAS overrides; the value must be the same on every rank)
    const char* env = getenv("LG_MC_CTAS");
    rg.grid = env ? atoi(env) : sm_count();
    if (rg.grid < 1) rg.grid = 1;
    if (rg.grid > MC_MAX_CTAS) rg.grid = MC_MAX_CTAS;
    if (getenv("LG_MC_TRACE")) {
        LG_CUDA(cudaMalloc((void**)&rg.trace, (size_t)TRACE_MAX * 4 * sizeof(unsigned long long)));
        LG_CUDA(cudaMemset(rg.trace, 0, (size_t)TRACE_MAX * 4 * sizeof(unsigned long long)));
    }
    rg.bound = true;
    *local_ptr = (void*)rg.local;
    *mc_ptr = (void*)rg.mcva;
    return 0;
}

// LG_MC_TRACE=1: lg_mc_trace_mark stamps the CURRENT stream's position in time; lg_mc_trace_read drains the device and
// copies out up to max_records records of 4 x uint64 {entered, all ranks met, finished, bytes} (globaltimer ns; a mark
// is {t, 0, 0, 0}) in launch order.  Launches captured into a CUDA graph keep their slots, so after a replay the trace
// is that replay's timeline; reset != 0 clears the trace (only when no captured step will be replayed any more).
int lg_mc_trace_mark(void) {
    LG_INIT();
    if (!rg.trace || rg.trace_n >= TRACE_MAX) return 0;
    mc_mark_kernel<<<1, 1, 0, stream()>>>(rg.trace + 4 * (size_t)rg.trace_n);
    ++rg.trace_n;
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_mc_trace_read(uint64_t* out, int max_records, int* n_records, int reset) {
    LG_INIT();
    *n_records = 0;
    if (!rg.trace) return 0;
    LG_CUDA(cudaDeviceSynchronize());
    const int n = rg.trace_n < max_records ? rg.trace_n : max_records;
    if (n > 0) LG_CUDA(cudaMemcpy(out, rg.trace, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (reset) {
        // (not while a captured step that recorded slots is still going to be replayed)
        LG_CUDA(cudaMemset(rg.trace, 0, (size_t)TRACE_MAX * 4 * sizeof(unsigned long long)));
        rg.trace_n = 0;
    }
    *n_records = n;
    return 0;
}

int lg_mc_release(void) {
    if (!rg.have_mc) return 0;
    cudaStreamSynchronize(comm_stream());
    cudaStreamSynchronize(stream());
    if (rg.bound) {
        drv.cuMemUnmap(rg.mcva, rg.bytes);
        drv.cuMemAddressFree(rg.mcva, rg.bytes);
        drv.cuMulticastUnbind(rg.mc, rg.dev, 0, rg.bytes);
        drv.cuMemUnmap(rg.local, rg.bytes);
        drv.cuMemAddressFree(rg.local, rg.bytes);
        drv.cuMemRelease(rg.mem);
        cudaFree(rg.epochs);
        if (rg.trace) cudaFree(rg.trace);
    }
    drv.cuMemRelease(rg.mc);
    rg = Region();
    return 0;
}

// One bucket [lo, hi) of the arenas: reduce-scatter of the gradients through the switch, optimizer update of this
// rank's share, all-gather of the new parameters -- one kernel, on the collective stream (after lg_nccl_fork /
// lg_comm_fork ordered it behind the gradients).  Offsets are in elements from the start of the region's arenas
// (grad_offset / param_offset of lg_mc_region_bytes, in bytes); kind: 0 Adam, 1 AdaBelief, 2 SGD, 3 dry run (the
// exchange without an update: parameters are written back unchanged).  The Adam arguments are those of lg_adam_step.
int lg_mc_exchange_step(int kind, size_t grad_offset, size_t param_offset, size_t flag_offset, int64_t lo, int64_t hi,
                        int rank, int world, void* m, void* v, int n_seg, const int64_t* seg_end_dev, int64_t* t_dev,
                        double lr, double beta1, double beta2, double eps, double momentum, int seg_offset,
                        int t_advance) {
    LG_REQUIRE(rg.bound, "lg_mc_exchange_step: no bound multicast region");
    LG_REQUIRE(world == rg.world && rank >= 0 && rank < world, "lg_mc_exchange_step: rank %d / world %d do not match the region", rank, world);
    LG_REQUIRE(kind >= 0 && kind <= 3, "lg_mc_exchange_step: unknown kind %d", kind);
    LG_REQUIRE(lo % 4 == 0 && hi % 4 == 0 && lo <= hi, "lg_mc_exchange_step: range must be a multiple of 4 elements");
    LG_REQUIRE(flag_offset + MC_FLAG_BYTES <= rg.bytes, "lg_mc_exchange_step: flag page outside the region");
    if (hi == lo) return 0;
    cudaStream_t st = comm_stream();
    McArgs a = {};
    a.mc_grad = (float*)(rg.mcva + grad_offset);
    a.mc_param = (float*)(rg.mcva + param_offset);
    a.param = (float*)(rg.local + param_offset);
    a.m = (float*)m;
    a.v = (float*)v;
    a.mc_flags = (unsigned int*)(rg.mcva + flag_offset);
    a.flags = (unsigned int*)(rg.local + flag_offset);
    a.epochs = rg.epochs;
    a.lo = lo;
    a.hi = hi;
    a.rank = rank;
    a.world = world;
    a.kind = kind;
    a.n_seg = n_seg;
    a.seg_end = seg_end_dev;
    a.neg_lr = (float)(-lr);
    a.b1 = (float)beta1;
    a.b2 = (float)beta2;
    a.omb1 = (float)(1.0 - beta1);
    a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.inv_world = 1.0f / (float)world;
    a.momentum = (float)momentum;
    a.trace = nullptr;
    if (rg.trace && rg.trace_n < TRACE_MAX) {
        a.trace = rg.trace + 4 * (size_t)rg.trace_n;     // (a captured launch keeps its slot: every replay rewrites it)
        ++rg.trace_n;
    }
    float* corr = nullptr;
    if (kind <= 1) {
        LG_REQUIRE(n_seg >= 1 && seg_end_dev && t_dev && m && v, "lg_mc_exchange_step: Adam needs its state and segments");
        corr = (float*)tmp_alloc(4 * (size_t)n_seg * sizeof(float));
        if (!corr) return 1;
        a.c1 = corr;
        a.c2 = corr + n_seg;
        adam_prep_kernel<<<1, 256, 0, st>>>(n_seg, t_dev, beta1, beta2, corr, corr + n_seg, seg_offset, t_advance);
        count_launch();
    }
    // (tried: 2 or 4 CTAs per SM for the bucket that closes the step -- the word embeddings, fully exposed -- 8.22-8.24 /
    //  8.28 ms per step at 2 GPUs against 8.20-8.24 with one: its 218 us are the NVLink transfer, not latency)
    const int grid = rg.grid;
#define MC_LAUNCH(K_, S_)                                                                   \
    do {                                                                                    \
        prefer_gemm_carveout((const void*)mc_exchange_kernel<K_, false, S_>);               \
        mc_exchange_kernel<K_, false, S_><<<grid, MC_THREADS, 0, st>>>(a);                  \
    } while (0)
    if (stream_hint()) {
        switch (kind) {
            case 0: MC_LAUNCH(0, true); break;
            case 1: MC_LAUNCH(1, true); break;
            case 2: MC_LAUNCH(2, true); break;
            default: MC_LAUNCH(3, true); break;
        }
    } else {
        switch (kind) {
            case 0: MC_LAUNCH(0, false); break;
            case 1: MC_LAUNCH(1, false); break;
            case 2: MC_LAUNCH(2, false); break;
            default: MC_LAUNCH(3, false); break;
        }
    }
#undef MC_LAUNCH
    if (corr) {
        // the collective stream still reads it: hand it back once the compute stream has joined (lg_nccl_wait)
        comm_defer_free(corr);
    }
    LG_CHECK_LAUNCH();
    return 0;
}

// One GPU: the optimizer update of arena elements [lo, hi) by the same small-footprint kernel, on the collective
// stream (ordered by lg_nccl_fork / lg_nccl_wait), so that it runs beside the rest of backward instead of after it.
// param / grad / m / v address the START of the arenas; kind 0 Adam, 1 AdaBelief, 2 SGD (m = previous deltas when
// momentum != 0).  Arithmetic, segments and step counter as lg_adam_step / lg_sgd_step.
int lg_bucket_step(int kind, void* param, const void* grad, void* m, void* v, int64_t lo, int64_t hi, int n_seg,
                   const int64_t* seg_end_dev, int64_t* t_dev, double lr, double beta1, double beta2, double eps,
                   double momentum, int seg_offset, int t_advance) {
    LG_INIT();
    LG_REQUIRE(kind >= 0 && kind <= 2, "lg_bucket_step: unknown kind %d", kind);
    LG_REQUIRE(lo % 4 == 0 && hi % 4 == 0 && lo <= hi, "lg_bucket_step: range must be a multiple of 4 elements");
    LG_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
               "lg_bucket_step: arenas must be 16-byte aligned");
    if (hi == lo) return 0;
    cudaStream_t st = comm_stream();
    McArgs a = {};
    a.grad = (const float*)grad;
    a.param = (float*)param;
    a.m = (float*)m;
    a.v = (float*)v;
    a.lo = lo;
    a.hi = hi;
    a.rank = 0;
    a.world = 1;
    a.kind = kind;
    a.n_seg = n_seg;
    a.seg_end = seg_end_dev;
    a.neg_lr = (float)(-lr);
    a.b1 = (float)beta1;
    a.b2 = (float)beta2;
    a.omb1 = (float)(1.0 - beta1);
    a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps;
    a.inv_world = 1.0f;
    a.momentum = (float)momentum;
    a.trace = nullptr;
    float* corr = nullptr;
    if (kind <= 1) {
        LG_REQUIRE(n_seg >= 1 && seg_end_dev && t_dev && m && v, "lg_bucket_step: Adam needs its state and segments");
        corr = (float*)tmp_alloc(4 * (size_t)n_seg * sizeof(float));
        if (!corr) return 1;
        a.c1 = corr;
        a.c2 = corr + n_seg;
        adam_prep_kernel<<<1, 256, 0, st>>>(n_seg, t_dev, beta1, beta2, corr, corr + n_seg, seg_offset, t_advance);
        count_launch();
    }
    static const int grid_env = getenv("LG_MC_CTAS") ? atoi(getenv("LG_MC_CTAS")) : 0;
    const int grid = grid_env > 0 ? (grid_env < MC_MAX_CTAS ? grid_env : MC_MAX_CTAS) : sm_count();
#define MC_LAUNCH(K_, S_)                                                                   \
    do {                                                                                    \
        prefer_gemm_carveout((const void*)mc_exchange_kernel<K_, true, S_>);                \
        mc_exchange_kernel<K_, true, S_><<<grid, MC_THREADS, 0, st>>>(a);                   \
    } while (0)
    if (stream_hint()) {
        switch (kind) {
            case 0: MC_LAUNCH(0, true); break;
            case 1: MC_LAUNCH(1, true); break;
            default: MC_LAUNCH(2, true); break;
        }
    } else {
        switch (kind) {
            case 0: MC_LAUNCH(0, false); break;
            case 1: MC_LAUNCH(1, false); break;
            default: MC_LAUNCH(2, false); break;
        }
    }
#undef MC_LAUNCH
    if (corr) comm_defer_free(corr);
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
