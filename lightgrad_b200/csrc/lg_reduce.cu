// Reductions sum / max / min over a contiguous (outer, reduce, inner) factorisation.
// Replaces the reference's multi-pass tree reduction (opencl/kernels.py:344-501), whose non-last-axis
// case walks a transposed view with per-element index math (uncoalesced).  Three kernels:
//   rows_warp  : inner == 1, short rows   -> one warp per row, shuffle tree
//   rows_block : inner == 1, long rows    -> (row, chunk) per CTA, 128-bit loads, 4 accumulators,
//                                            shuffle + shared-memory tree; second pass over chunk partials
//   cols       : inner  > 1               -> threads along the contiguous inner dim (coalesced),
//                                            8 row-lanes per CTA combined through shared memory
// Deterministic (no atomics).  Algorithmic bytes: 4 per input element (f32), output negligible.
#include "lg_ew.cuh"
#include <math.h>

using namespace lg;

namespace {

template <typename T> struct Lim;
template <> struct Lim<float> { static __device__ float inf() { return INFINITY; } };
template <> struct Lim<double> { static __device__ double inf() { return (double)INFINITY; } };

struct RSum {
    template <typename T> static __device__ __forceinline__ T init() { return T(0); }
    template <typename T> static __device__ __forceinline__ T comb(T a, T b) { return a + b; }
};
struct RMax {  // NaN-propagating like np.max
    template <typename T> static __device__ __forceinline__ T init() { return -Lim<T>::inf(); }
    template <typename T> static __device__ __forceinline__ T comb(T a, T b) { return (b > a || b != b) ? b : a; }
};
struct RMin {
    template <typename T> static __device__ __forceinline__ T init() { return Lim<T>::inf(); }
    template <typename T> static __device__ __forceinline__ T comb(T a, T b) { return (b < a || b != b) ? b : a; }
};

template <class R, typename T>
__device__ __forceinline__ T warp_reduce(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = R::template comb<T>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- inner == 1, short rows ----------------------------------------------------------------------
template <class R, typename T, int V>
__global__ void __launch_bounds__(256) red_rows_warp(const T* __restrict__ x, T* __restrict__ out, int64_t rows,
                                                     int64_t len, T scale, int acc_out) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nwarps) {
        const T* p = x + row * len;
        T acc = R::template init<T>();
        const int64_t nv = len / V;
        for (int64_t i = lane; i < nv; i += 32) {
            VT v = reinterpret_cast<const VT*>(p)[i];
#pragma unroll
            for (int k = 0; k < V; ++k) acc = R::template comb<T>(acc, v.v[k]);
        }
        for (int64_t i = nv * V + lane; i < len; i += 32) acc = R::template comb<T>(acc, p[i]);
        acc = warp_reduce<R, T>(acc);
        if (lane == 0) out[row] = acc_out ? out[row] + acc * scale : acc * scale;
    }
}

// ---- inner == 1, long rows -----------------------------------------------------------------------
template <class R, typename T, int V>
__global__ void __launch_bounds__(256) red_rows_block(const T* __restrict__ x, T* __restrict__ out, int64_t len,
                                                      int64_t chunk, int S, T scale, int acc_out) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    __shared__ T sm[8];
    const int64_t row = blockIdx.x / S;
    const int part = (int)(blockIdx.x % S);
    const int64_t beg = (int64_t)part * chunk;
    int64_t end = beg + chunk;
    if (end > len) end = len;
    const T* p = x + row * len + beg;
    const int64_t n = end - beg;
    const int64_t nv = n / V;
    T a0 = R::template init<T>(), a1 = a0, a2 = a0, a3 = a0;
    int64_t i = threadIdx.x;
    for (; i + 3 * 256 < nv; i += 4 * 256) {
        VT v0 = reinterpret_cast<const VT*>(p)[i], v1 = reinterpret_cast<const VT*>(p)[i + 256],
           v2 = reinterpret_cast<const VT*>(p)[i + 512], v3 = reinterpret_cast<const VT*>(p)[i + 768];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            a0 = R::template comb<T>(a0, v0.v[k]);
            a1 = R::template comb<T>(a1, v1.v[k]);
            a2 = R::template comb<T>(a2, v2.v[k]);
            a3 = R::template comb<T>(a3, v3.v[k]);
        }
    }
    for (; i < nv; i += 256) {
        VT v0 = reinterpret_cast<const VT*>(p)[i];
#pragma unroll
        for (int k = 0; k < V; ++k) a0 = R::template comb<T>(a0, v0.v[k]);
    }
    for (int64_t j = nv * V + threadIdx.x; j < n; j += 256) a1 = R::template comb<T>(a1, p[j]);
    T acc = R::template comb<T>(R::template comb<T>(a0, a1), R::template comb<T>(a2, a3));
    acc = warp_reduce<R, T>(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < 8 ? sm[threadIdx.x] : R::template init<T>();
        v = warp_reduce<R, T>(v);
        if (threadIdx.x == 0) out[row * S + part] = acc_out ? out[row * S + part] + v * scale : v * scale;
    }
}

// ---- inner > 1 -----------------------------------------------------------------------------------
// L2-coherent (L1-bypassing) vector load of partials written by other CTAs of the same launch
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldcg_vec(const T* p) {
    Vec<T, V> r;
    if constexpr (sizeof(Vec<T, V>) == 16) {
        const float4 raw = __ldcg(reinterpret_cast<const float4*>(p));
        memcpy(&r, &raw, 16);
    } else if constexpr (sizeof(Vec<T, V>) == 8) {
        const float2 raw = __ldcg(reinterpret_cast<const float2*>(p));
        memcpy(&r, &raw, 8);
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) r.v[k] = __ldcg(p + k);
    }
    return r;
}
template <class R, typename T, int V>
__global__ void __launch_bounds__(256) red_cols(const T* __restrict__ x, T* __restrict__ out, int64_t rlen,
                                                int64_t inner, int64_t ld, int64_t chunk, int S, int64_t ntiles,
                                                T scale, int acc_out, T* __restrict__ final_out,
                                                unsigned int* __restrict__ tickets) {
    LG_PDL_TRIGGER();
    // final_out != nullptr (S > 1): `out` holds the per-chunk partials and the LAST CTA to finish a column
    // tile (ticket counter) sums them in chunk order into final_out -- one launch, still deterministic
    using VT = Vec<T, V>;
    __shared__ T sm[8][32 * V + 1];
    __shared__ int s_last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t tile = blockIdx.x % ntiles;
    const int64_t rest = blockIdx.x / ntiles;
    const int part = (int)(rest % S);
    const int64_t o = rest / S;
    const int64_t col = (tile * 32 + tx) * V;
    const int64_t beg = (int64_t)part * chunk;
    int64_t end = beg + chunk;
    if (end > rlen) end = rlen;
    T acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = R::template init<T>();
    if (col < inner) {
        const T* p = x + o * rlen * ld + col;   // ld = elements between consecutive rows (>= inner)
        int64_t r = beg + ty;
        for (; r + 56 < end; r += 64) {
            VT w0 = *reinterpret_cast<const VT*>(p + r * ld), w1 = *reinterpret_cast<const VT*>(p + (r + 8) * ld),
               w2 = *reinterpret_cast<const VT*>(p + (r + 16) * ld), w3 = *reinterpret_cast<const VT*>(p + (r + 24) * ld),
               w4 = *reinterpret_cast<const VT*>(p + (r + 32) * ld), w5 = *reinterpret_cast<const VT*>(p + (r + 40) * ld),
               w6 = *reinterpret_cast<const VT*>(p + (r + 48) * ld), w7 = *reinterpret_cast<const VT*>(p + (r + 56) * ld);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                T lo = R::template comb<T>(R::template comb<T>(w0.v[k], w1.v[k]), R::template comb<T>(w2.v[k], w3.v[k]));
                T hi = R::template comb<T>(R::template comb<T>(w4.v[k], w5.v[k]), R::template comb<T>(w6.v[k], w7.v[k]));
                acc[k] = R::template comb<T>(acc[k], R::template comb<T>(lo, hi));
            }
        }
        for (; r + 24 < end; r += 32) {
            VT v0 = *reinterpret_cast<const VT*>(p + r * ld), v1 = *reinterpret_cast<const VT*>(p + (r + 8) * ld),
               v2 = *reinterpret_cast<const VT*>(p + (r + 16) * ld),
               v3 = *reinterpret_cast<const VT*>(p + (r + 24) * ld);
#pragma unroll
            for (int k = 0; k < V; ++k)
                acc[k] = R::template comb<T>(R::template comb<T>(R::template comb<T>(acc[k], v0.v[k]), v1.v[k]),
                                             R::template comb<T>(v2.v[k], v3.v[k]));
        }
        for (; r < end; r += 8) {
            VT v0 = *reinterpret_cast<const VT*>(p + r * ld);
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = R::template comb<T>(acc[k], v0.v[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) sm[ty][tx * V + k] = acc[k];
    __syncthreads();
    // 32*V columns, 8 partials each; threads 0 .. 32*V-1 finish one column each
    for (int c = threadIdx.x; c < 32 * V; c += 256) {
        T v = sm[0][c];
#pragma unroll
        for (int j = 1; j < 8; ++j) v = R::template comb<T>(v, sm[j][c]);
        int64_t gc = tile * 32 * V + c;
        if (gc < inner) {
            T* dst = out + (o * S + part) * inner + gc;
            if (final_out) *dst = v;
            else *dst = acc_out ? *dst + v * scale : v * scale;
        }
    }
    if (final_out == nullptr) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&tickets[o * ntiles + tile], 1u);
        s_last = (t == (unsigned int)(S - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // final pass over the S partial rows of this column tile: warp ty takes partials ty, ty+8, ... (their loads
    // are independent, so the whole pass costs about one L2 round trip), then the 8 warp sums are combined in
    // a fixed order -- the result does not depend on which CTA happened to finish last
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = R::template init<T>();
    if (col < inner) {
        const T* q = out + o * S * inner + col;
        int pp = ty;
        for (; pp + 24 < S; pp += 32) {
            VT w0 = ldcg_vec<T, V>(q + (int64_t)pp * inner), w1 = ldcg_vec<T, V>(q + (int64_t)(pp + 8) * inner),
               w2 = ldcg_vec<T, V>(q + (int64_t)(pp + 16) * inner), w3 = ldcg_vec<T, V>(q + (int64_t)(pp + 24) * inner);
#pragma unroll
            for (int k = 0; k < V; ++k)
                acc[k] = R::template comb<T>(R::template comb<T>(R::template comb<T>(acc[k], w0.v[k]), w1.v[k]),
                                             R::template comb<T>(w2.v[k], w3.v[k]));
        }
        for (; pp < S; pp += 8) {
            VT w0 = ldcg_vec<T, V>(q + (int64_t)pp * inner);
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = R::template comb<T>(acc[k], w0.v[k]);
        }
    }
    __syncthreads();   // everyone is done reading sm from the first pass
#pragma unroll
    for (int k = 0; k < V; ++k) sm[ty][tx * V + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * V; c += 256) {
        int64_t gc = tile * 32 * V + c;
        if (gc >= inner) continue;
        T v = sm[0][c];
#pragma unroll
        for (int j = 1; j < 8; ++j) v = R::template comb<T>(v, sm[j][c]);
        T* dst = final_out + o * inner + gc;
        *dst = acc_out ? *dst + v * scale : v * scale;
    }
    if (threadIdx.x == 0) tickets[o * ntiles + tile] = 0;   // ready for the next launch
}

// ticket counters of the single-launch column reduction (zeroed once; every launch leaves them zero)
constexpr int64_t kTicketSlots = 16384;
// (one set per stream: reductions on the side stream may run concurrently with reductions on the main one)
unsigned int* ticket_buffer() {
    static unsigned int* bufs[3] = {nullptr, nullptr, nullptr};
    unsigned int*& buf = bufs[alt_stream_index()];
    if (!buf) {
        if (cudaMalloc(&buf, kTicketSlots * sizeof(unsigned int)) != cudaSuccess) {
            cudaGetLastError();
            buf = nullptr;
            return nullptr;
        }
        cudaMemsetAsync(buf, 0, kTicketSlots * sizeof(unsigned int), stream());
    }
    return buf;
}

template <class R, typename T>
int reduce_impl(const T* x, T* out, int64_t outer, int64_t rlen, int64_t inner, T scale, int64_t ld = 0,
                int acc_out = 0) {
    if (ld <= 0 || inner == 1) ld = inner;
    constexpr int VMAX = 16 / sizeof(T);
    const int sms = sm_count();
    if (outer * inner == 0) return 0;
    if (rlen == 0) return set_error("lg_reduce: zero-size reduction has no identity here");
    if (inner == 1) {
        const bool vec = (rlen % VMAX == 0) && aligned16(x);
        if (rlen <= 2048 && outer >= 64) {
            int64_t blocks = (outer + 7) / 8;
            int64_t cap = (int64_t)sms * 32;
            int grid = (int)(blocks < cap ? blocks : cap);
            if (vec) red_rows_warp<R, T, VMAX><<<grid, 256, 0, stream()>>>(x, out, outer, rlen, scale, acc_out);
            else red_rows_warp<R, T, 1><<<grid, 256, 0, stream()>>>(x, out, outer, rlen, scale, acc_out);
            LG_CHECK_LAUNCH();
            return 0;
        }
        // split long rows so that the grid covers the machine (~4 CTAs per SM)
        int64_t S = 1;
        if (outer < (int64_t)sms * 4) {
            S = ((int64_t)sms * 4 + outer - 1) / outer;
            int64_t maxS = (rlen + 4095) / 4096;  // at least 4096 elements per chunk
            if (S > maxS) S = maxS;
            if (S < 1) S = 1;
        }
        int64_t chunk = (rlen + S - 1) / S;
        chunk = (chunk + 1023) / 1024 * 1024;  // keeps every chunk start 16-byte aligned
        S = (rlen + chunk - 1) / chunk;
        LG_REQUIRE(outer * S < 0x7fffffff, "lg_reduce: grid too large");
        T* dst = out;
        T* partial = nullptr;
        if (S > 1) {
            partial = (T*)tmp_alloc((size_t)(outer * S) * sizeof(T));
            if (!partial) return 1;
            dst = partial;
        }
        T sc = (S > 1) ? T(1) : scale;
        int grid = (int)(outer * S);
        if (vec) red_rows_block<R, T, VMAX><<<grid, 256, 0, stream()>>>(x, dst, rlen, chunk, (int)S, sc, S > 1 ? 0 : acc_out);
        else red_rows_block<R, T, 1><<<grid, 256, 0, stream()>>>(x, dst, rlen, chunk, (int)S, sc, S > 1 ? 0 : acc_out);
        LG_CHECK_LAUNCH();
        if (S > 1) {
            int rc = reduce_impl<R, T>(partial, out, outer, S, 1, scale, 0, acc_out);
            tmp_free(partial);
            return rc;
        }
        return 0;
    }
    // column reduce
    const bool vec = (inner % VMAX == 0) && (ld % VMAX == 0) && aligned16(x);
    const int V = vec ? VMAX : 1;
    int64_t ntiles = (inner + 32 * V - 1) / (32 * V);
    int64_t S = 1;
    if (ntiles * outer < (int64_t)sms * 4) {
        S = ((int64_t)sms * 4 + ntiles * outer - 1) / (ntiles * outer);
        int64_t maxS = (rlen + 63) / 64;
        if (S > maxS) S = maxS;
        if (S < 1) S = 1;
    }
    int64_t chunk = (rlen + S - 1) / S;
    S = (rlen + chunk - 1) / chunk;
    LG_REQUIRE(ntiles * outer * S < 0x7fffffff, "lg_reduce: grid too large");
    T* dst = out;
    T* partial = nullptr;
    if (S > 1) {
        partial = (T*)tmp_alloc((size_t)(outer * S * inner) * sizeof(T));
        if (!partial) return 1;
        dst = partial;
    }
    T sc = (S > 1) ? T(1) : scale;
    int grid = (int)(ntiles * outer * S);
    unsigned int* tickets = (S > 1 && ntiles * outer <= kTicketSlots) ? ticket_buffer() : nullptr;
    if (tickets) {
        // single launch: partials + last-CTA final pass
        if (vec) { prefer_gemm_carveout((const void*)red_cols<R, T, VMAX>); red_cols<R, T, VMAX><<<grid, 256, 0, stream()>>>(x, dst, rlen, inner, ld, chunk, (int)S, ntiles, scale, acc_out, out, tickets); }
        else { prefer_gemm_carveout((const void*)red_cols<R, T, 1>); red_cols<R, T, 1><<<grid, 256, 0, stream()>>>(x, dst, rlen, inner, ld, chunk, (int)S, ntiles, scale, acc_out, out, tickets); }
        LG_CHECK_LAUNCH();
        tmp_free(partial);
        return 0;
    }
    if (vec) { prefer_gemm_carveout((const void*)red_cols<R, T, VMAX>); red_cols<R, T, VMAX><<<grid, 256, 0, stream()>>>(x, dst, rlen, inner, ld, chunk, (int)S, ntiles, sc, S > 1 ? 0 : acc_out, (T*)nullptr, nullptr); }
    else { prefer_gemm_carveout((const void*)red_cols<R, T, 1>); red_cols<R, T, 1><<<grid, 256, 0, stream()>>>(x, dst, rlen, inner, ld, chunk, (int)S, ntiles, sc, S > 1 ? 0 : acc_out, (T*)nullptr, nullptr); }
    LG_CHECK_LAUNCH();
    if (S > 1) {
        int rc = reduce_impl<R, T>(partial, out, outer, S, inner, scale, 0, acc_out);
        tmp_free(partial);
        return rc;
    }
    return 0;
}

// out[c] += scale * sum_r x[r, c] for a float32 (rows, cols) matrix with row pitch ld, SMALL FOOTPRINT: 128 threads, no
// shared memory, ~40 registers.  This is the bias-gradient column sum of the backward pass (added straight into the
// gradient arena), issued on the side stream while the compute stream runs one-CTA-per-SM tensor-core GEMMs: the general
// column reducer above needs 5 KB of shared memory per CTA and therefore waits for a free SM, this one fits into the
// ~1.9 KB / 11.7 K registers an SM has left beside a GEMM CTA and runs next to it.
// Every thread owns one 16-byte column group over a slice of the rows (a warp reads 512 contiguous bytes per row,
// U rows in flight) and adds its four partial sums to the result with atomics (red.global.add.f32).
__global__ void __launch_bounds__(128) colsum_acc_small_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                               int64_t rows, int nvec, int64_t ld, int64_t rows_per,
                                                               float scale) {
    LG_PDL_TRIGGER();
    constexpr int U = 8;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int v = (int)(g % nvec);
    const int64_t r0 = (g / nvec) * rows_per;
    int64_t r1 = r0 + rows_per;
    if (r1 > rows) r1 = rows;
    if (r0 >= rows) return;
    const float4* p = reinterpret_cast<const float4*>(x) + v;
    const int64_t ldv = ld >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = r0;
    for (; r + U <= r1; r += U) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = __ldg(p + (r + u) * ldv);
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w; }
    }
    for (; r < r1; ++r) {
        const float4 t = __ldg(p + r * ldv);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float* o = out + 4 * (int64_t)v;
    atomicAdd(o + 0, acc.x * scale);
    atomicAdd(o + 1, acc.y * scale);
    atomicAdd(o + 2, acc.z * scale);
    atomicAdd(o + 3, acc.w * scale);
}

// returns true when it took the problem
bool colsum_acc_small(const float* x, float* out, int64_t rows, int64_t cols, int64_t ld, float scale) {
    static const bool off = getenv("LG_NO_SMALL_COLSUM") != nullptr;
    if (off || cols % 4 || ld % 4 || !aligned16(x) || rows < 256 || cols > (1 << 20)) return false;
    const int nvec = (int)(cols / 4);
    const int64_t threads = (int64_t)lg::sm_count() * 128;      // one CTA per SM: resident beside a GEMM CTA
    int64_t slices = threads / nvec;
    if (slices < 1) slices = 1;
    if (slices > rows / 32) slices = rows / 32;                 // >= 32 rows per thread
    const int64_t rows_per = (rows + slices - 1) / slices;
    slices = (rows + rows_per - 1) / rows_per;
    const int64_t total = slices * nvec;
    prefer_gemm_carveout((const void*)colsum_acc_small_kernel);     // runs on the side stream beside GEMM CTAs
    colsum_acc_small_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream()>>>(x, out, rows, nvec, ld, rows_per, scale);
    return true;
}

template <typename T>
int reduce_op(int op, const void* x, void* out, int64_t outer, int64_t rlen, int64_t inner, double scale, int64_t ld,
              int acc_out = 0) {
    switch (op) {
        case LG_RED_SUM: return reduce_impl<RSum, T>((const T*)x, (T*)out, outer, rlen, inner, (T)scale, ld, acc_out);
        case LG_RED_MAX: return reduce_impl<RMax, T>((const T*)x, (T*)out, outer, rlen, inner, (T)1, ld);
        case LG_RED_MIN: return reduce_impl<RMin, T>((const T*)x, (T*)out, outer, rlen, inner, (T)1, ld);
    }
    return set_error("lg_reduce: unknown op %d", op);
}

}  // namespace

extern "C" int lg_reduce(int op, int dtype, const void* x, void* out, int64_t outer, int64_t rlen, int64_t inner,
                         double scale) {
    LG_INIT();
    if (dtype == LG_F32) return reduce_op<float>(op, x, out, outer, rlen, inner, scale, 0);
    if (dtype == LG_F64) return reduce_op<double>(op, x, out, outer, rlen, inner, scale, 0);
    return set_error("lg_reduce: unsupported dtype %d", dtype);
}

extern "C" int lg_reduce_pitched(int op, int dtype, const void* x, void* out, int64_t outer, int64_t rlen,
                                 int64_t inner, int64_t ld, double scale, int accumulate) {
    LG_INIT();
    LG_REQUIRE(ld >= inner && inner >= 1, "lg_reduce_pitched: need ld >= inner >= 1");
    LG_REQUIRE(inner > 1 || ld == 1, "lg_reduce_pitched: a pitch needs inner > 1");
    LG_REQUIRE(!accumulate || op == LG_RED_SUM, "lg_reduce_pitched: accumulate is defined for sums only");
    if (dtype == LG_F32 && op == LG_RED_SUM && accumulate && outer == 1 && inner > 1 &&
        colsum_acc_small((const float*)x, (float*)out, rlen, inner, ld, (float)scale)) {
        LG_CHECK_LAUNCH();
        return 0;
    }
    if (dtype == LG_F32) return reduce_op<float>(op, x, out, outer, rlen, inner, scale, ld, accumulate);
    if (dtype == LG_F64) return reduce_op<double>(op, x, out, outer, rlen, inner, scale, ld, accumulate);
    return set_error("lg_reduce_pitched: unsupported dtype %d", dtype);
}
