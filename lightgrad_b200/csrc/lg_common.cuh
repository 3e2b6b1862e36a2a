// Internal helpers shared by all translation units of liblightgrad_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/lightgrad_b200.h"

// Programmatic dependent launch: every kernel signals at its very start that a dependent kernel launched
// with the programmatic-serialisation attribute (the tensor-core GEMMs) may begin its prologue; that kernel
// executes LG_PDL_WAIT() -- completion and memory flush of everything before it -- before touching memory.
#define LG_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define LG_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

namespace lg {

// Launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the previous kernel of the
// stream is still finishing (it has executed LG_PDL_TRIGGER in every CTA); the kernel must execute LG_PDL_WAIT before
// it touches anything the previous kernel wrote.  Used for the small kernels between the GEMMs of a step (LayerNorm,
// attention, cross entropy) as it is for the GEMMs; inside the replayed step graph the effect is within the noise
// (7.97 / 8.06 vs 8.04 / 8.01 ms per step with LG_NO_PDL_SMALL=1), eager streams gain the launch latency.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    static const bool off = getenv("LG_NO_PDL_SMALL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- error plumbing ---------------------------------------------------------------------------
int set_error(const char* fmt, ...);
void prefer_gemm_carveout(const void* kernel);   // this kernel runs beside GEMM CTAs: ask for their shared-memory split
cudaStream_t stream();        // stream kernels launch on: the compute stream, or the side stream inside lg_side_begin/end
bool on_side_stream();      // launching on the side stream or (lg_comm_compute_begin) on the collective stream
int alt_stream_index();     // 0 compute, 1 side, 2 collective stream
void comm_release_deferred();
void comm_defer_free(void* p);   // block from tmp_alloc that the collective stream still uses: freed by comm_release_deferred
int side_join();
int side_order_before(cudaStream_t other);
cudaStream_t comm_stream();   // collective stream
// lg_nccl.cu: 0 = no communicator or no error; 1 = NCCL reported an asynchronous error (message set, communicator
// aborted).  nccl_abort(): abort the communicator so that its kernels leave the GPU (a peer died / timed out).
int nccl_async_check();
void nccl_abort(const char* why);
bool nccl_active();
int sm_count();
void count_launch(int n = 1);
int ensure_init();
bool capturing();     // a whole-step CUDA graph capture is in progress on the compute stream
// Device-side error word (host-mapped, so stable under graph replay): kernels that meet an out-of-range index or
// label store a code there instead of reading out of bounds; the next lg_sync / lg_memcpy_d2h reports it once.
enum { LG_DEVERR_INDEX = 1, LG_DEVERR_LABEL = 2 };
unsigned int* error_flag();

#define LG_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return lg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                                 __FILE__, __LINE__);                                         \
    } while (0)

#define LG_CHECK_LAUNCH()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return lg::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                                 __FILE__, __LINE__);                                         \
        lg::count_launch();                                                                   \
    } while (0)

#define LG_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) return lg::set_error(__VA_ARGS__);                                       \
    } while (0)

#define LG_INIT()                                                                             \
    do {                                                                                      \
        int _rc = lg::ensure_init();                                                          \
        if (_rc) return _rc;                                                                  \
    } while (0)

inline size_t dtype_size(int dt) {
    switch (dt) {
        case LG_F32: case LG_I32: return 4;
        case LG_F64: case LG_I64: return 8;
        case LG_I16: case LG_BF16: return 2;
        case LG_U8: case LG_I8: return 1;
    }
    return 0;
}

// internal allocation helpers (same cache as lg_alloc / lg_free)
void* tmp_alloc(size_t nbytes);
void tmp_free(void* p);

// grid sizing: persistent-style grids are multiples of the SM count
inline int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 8) {
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace lg
