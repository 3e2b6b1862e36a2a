// Integer-array indexing: gather, scatter-add, scatter-assign, index linearisation.
// Replaces numpy fancy indexing in the reference's getitem/setitem (cpu/ops.py:234-255); the OpenCL
// backend has no such path (opencl/ops.py:292-331 supports int/slice only, which is why
// examples/bert.py:21 round-trips the embedding table through the CPU).
// Gather is bit-exact (pure data movement).  getitem backward is scatter-ADD (atomicAdd), see
// SURVEY.md F4c.  Algorithmic bytes: gather = 2 * n_idx * row_len * sizeof(T) (+ indices).
#include "lg_ew.cuh"

using namespace lg;

namespace {

template <typename I>
__device__ __forceinline__ int64_t load_idx(const void* idx, int64_t i) { return (int64_t)((const I*)idx)[i]; }

__device__ __forceinline__ int64_t fetch_index(const void* idx, int dt, int64_t i) {
    switch (dt) {
        case LG_I32: return load_idx<int32_t>(idx, i);
        case LG_I64: return load_idx<int64_t>(idx, i);
        case LG_I16: return load_idx<int16_t>(idx, i);
        case LG_U8: return load_idx<uint8_t>(idx, i);
        case LG_I8: return load_idx<int8_t>(idx, i);
    }
    return 0;
}

// W = machine word used for the copy (uint4 when rows are 16-byte multiples)
template <typename W>
__global__ void __launch_bounds__(256) gather_rows_kernel(const W* __restrict__ src, int64_t n_src_rows,
                                                          int64_t row_stride_w, const void* __restrict__ idx,
                                                          int idx_dt, int64_t n_idx, int64_t row_words,
                                                          W* __restrict__ out, unsigned int* __restrict__ errflag) {
    LG_PDL_TRIGGER();
    const int64_t total = n_idx * row_words;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = w / row_words, j = w - i * row_words;
        int64_t r = fetch_index(idx, idx_dt, i);
        if (r < 0) r += n_src_rows;
        W v;
        if (r >= 0 && r < n_src_rows) v = src[r * row_stride_w + j];
        else {
            // numpy raises IndexError here; the row is zero-filled and the error surfaces at the next sync
            memset(&v, 0, sizeof(W));
            if (j == 0) *errflag = LG_DEVERR_INDEX;
        }
        out[w] = v;
    }
}

template <typename W>
__global__ void __launch_bounds__(256) scatter_set_rows_kernel(W* __restrict__ dst, int64_t n_dst_rows,
                                                               int64_t row_stride_w, const void* __restrict__ idx,
                                                               int idx_dt, int64_t n_idx, int64_t row_words,
                                                               const W* __restrict__ src, W value,
                                                               unsigned int* __restrict__ errflag) {
    LG_PDL_TRIGGER();
    const int64_t total = n_idx * row_words;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = w / row_words, j = w - i * row_words;
        int64_t r = fetch_index(idx, idx_dt, i);
        if (r < 0) r += n_dst_rows;
        if (r >= 0 && r < n_dst_rows) dst[r * row_stride_w + j] = src ? src[w] : value;
        else if (j == 0) *errflag = LG_DEVERR_INDEX;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(T* __restrict__ dst, int64_t n_dst_rows,
                                                               int64_t row_stride, const void* __restrict__ idx,
                                                               int idx_dt, int64_t n_idx, int64_t row_len,
                                                               const T* __restrict__ src,
                                                               unsigned int* __restrict__ errflag) {
    LG_PDL_TRIGGER();
    const int64_t total = n_idx * row_len;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = w / row_len, j = w - i * row_len;
        int64_t r = fetch_index(idx, idx_dt, i);
        if (r < 0) r += n_dst_rows;
        if (r >= 0 && r < n_dst_rows) atomicAdd(dst + r * row_stride + j, src[w]);
        else if (j == 0) *errflag = LG_DEVERR_INDEX;
    }
}

struct LinArgs {
    int n_arrays;
    const void* idx[4];
    int dt[4];
    int64_t size[4], stride[4];
};

__global__ void __launch_bounds__(256) linearize_kernel(LinArgs a, int64_t n, int64_t* __restrict__ lin,
                                                        unsigned int* __restrict__ errflag) {
    LG_PDL_TRIGGER();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t off = 0;
        bool bad = false;
        for (int k = 0; k < a.n_arrays; ++k) {
            int64_t r = fetch_index(a.idx[k], a.dt[k], i);
            if (r < 0) r += a.size[k];
            if (r < 0 || r >= a.size[k]) bad = true;
            off += r * a.stride[k];
        }
        // an out-of-range entry raises IndexError at the next sync; INT64_MIN makes the consumers skip that row
        // (they are handed element offsets with a row count of 2^62, so it stays negative after their wrap test)
        if (bad) *errflag = LG_DEVERR_INDEX;
        lin[i] = bad ? INT64_MIN : off;
    }
}

bool idx_dtype_ok(int dt) { return dt == LG_I32 || dt == LG_I64 || dt == LG_I16 || dt == LG_U8 || dt == LG_I8; }

}  // namespace

extern "C" {

int lg_gather_rows(int dtype, int idx_dtype, const void* src, int64_t n_src_rows, int64_t row_stride,
                   const void* idx, int64_t n_idx, int64_t row_len, void* out) {
    LG_INIT();
    LG_REQUIRE(idx_dtype_ok(idx_dtype), "lg_gather_rows: index dtype %d is not an integer type", idx_dtype);
    size_t es = dtype_size(dtype);
    LG_REQUIRE(es > 0, "lg_gather_rows: bad dtype %d", dtype);
    if (n_idx * row_len == 0) return 0;
    int64_t row_bytes = row_len * (int64_t)es, stride_bytes = row_stride * (int64_t)es;
    int grid;
#define GO(W)                                                                                              \
    grid = grid_for(n_idx * (row_bytes / (int64_t)sizeof(W)), 256, 8);                                     \
    gather_rows_kernel<W><<<grid, 256, 0, stream()>>>((const W*)src, n_src_rows,                           \
                                                      stride_bytes / (int64_t)sizeof(W), idx, idx_dtype,  \
                                                      n_idx, row_bytes / (int64_t)sizeof(W), (W*)out, error_flag())
    if (row_bytes % 16 == 0 && stride_bytes % 16 == 0 && aligned16(src) && aligned16(out)) { GO(uint4); }
    else if (es == 8) { GO(uint64_t); }
    else if (es == 4) { GO(uint32_t); }
    else if (es == 2) { GO(uint16_t); }
    else { GO(uint8_t); }
#undef GO
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_scatter_set_rows(int dtype, int idx_dtype, void* dst, int64_t n_dst_rows, int64_t row_stride, const void* idx,
                        int64_t n_idx, int64_t row_len, const void* src, double value) {
    LG_INIT();
    LG_REQUIRE(idx_dtype_ok(idx_dtype), "lg_scatter_set_rows: index dtype %d is not an integer type", idx_dtype);
    if (n_idx * row_len == 0) return 0;
    int grid = grid_for(n_idx * row_len, 256, 8);
#define GO(W, VAL)                                                                                         \
    scatter_set_rows_kernel<W><<<grid, 256, 0, stream()>>>((W*)dst, n_dst_rows, row_stride, idx, idx_dtype, \
                                                           n_idx, row_len, (const W*)src, VAL, error_flag())
    switch (dtype) {
        case LG_F32: GO(float, (float)value); break;
        case LG_F64: GO(double, value); break;
        case LG_I32: GO(int32_t, (int32_t)value); break;
        case LG_I64: GO(int64_t, (int64_t)value); break;
        case LG_I16: GO(int16_t, (int16_t)value); break;
        case LG_U8: GO(uint8_t, (uint8_t)value); break;
        case LG_I8: GO(int8_t, (int8_t)value); break;
        default: return set_error("lg_scatter_set_rows: bad dtype %d", dtype);
    }
#undef GO
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_scatter_add_rows(int dtype, int idx_dtype, void* dst, int64_t n_dst_rows, int64_t row_stride, const void* idx,
                        int64_t n_idx, int64_t row_len, const void* src) {
    LG_INIT();
    LG_REQUIRE(idx_dtype_ok(idx_dtype), "lg_scatter_add_rows: index dtype %d is not an integer type", idx_dtype);
    if (n_idx * row_len == 0) return 0;
    int grid = grid_for(n_idx * row_len, 256, 8);
    if (dtype == LG_F32)
        scatter_add_rows_kernel<float><<<grid, 256, 0, stream()>>>((float*)dst, n_dst_rows, row_stride, idx,
                                                                   idx_dtype, n_idx, row_len, (const float*)src, error_flag());
    else if (dtype == LG_F64)
        scatter_add_rows_kernel<double><<<grid, 256, 0, stream()>>>((double*)dst, n_dst_rows, row_stride, idx,
                                                                    idx_dtype, n_idx, row_len, (const double*)src, error_flag());
    else
        return set_error("lg_scatter_add_rows: unsupported dtype %d", dtype);
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_index_linearize(int n_arrays, const void* const* idx, const int* idx_dtypes, const int64_t* dim_sizes,
                       const int64_t* dim_strides, int64_t n, int64_t* lin) {
    LG_INIT();
    LG_REQUIRE(n_arrays >= 1 && n_arrays <= 4, "lg_index_linearize: 1..4 index arrays supported, got %d", n_arrays);
    LinArgs a;
    a.n_arrays = n_arrays;
    for (int k = 0; k < n_arrays; ++k) {
        LG_REQUIRE(idx_dtype_ok(idx_dtypes[k]), "lg_index_linearize: index dtype %d is not an integer type",
                   idx_dtypes[k]);
        a.idx[k] = idx[k];
        a.dt[k] = idx_dtypes[k];
        a.size[k] = dim_sizes[k];
        a.stride[k] = dim_strides[k];
    }
    if (n == 0) return 0;
    linearize_kernel<<<grid_for(n, 256, 8), 256, 0, stream()>>>(a, n, lin, error_flag());
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
