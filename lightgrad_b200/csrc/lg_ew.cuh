// Elementwise engine: one templated kernel family for every out = f(a[,b[,c]]; alpha).
//
// Replaces the reference's `atom` JIT generator (opencl/kernels.py:24-195), which emits one
// element per work-item and recomputes ndim div/mod chains for every operand.  Here:
//   * flat   -- all operands contiguous: 128-bit vector accesses, 8 independent vectors in flight per thread over
//               contiguous 32 KB chunks per CTA, grid sized from the SM count (HBM-streaming path, the roofline case);
//   * nd_vec -- broadcast / strided outer dims with a unit-or-zero inner stride: one 128-bit vector
//               per thread along the inner dim (bias adds, (R,1) statistics, head split/merge copies);
//   * nd_any -- arbitrary strides, one element per thread (transposed views, slices with steps).
// Algorithmic bytes per element (f32): 4 per operand read + 4 written.
#pragma once
#include "lg_common.cuh"

namespace lg {

template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
    T v[V];
};

struct EwShape {
    int ndim;
    int64_t shape[LG_MAX_DIMS];
    int64_t st[4][LG_MAX_DIMS];  // a, b, c, out
};

int ew_dispatch1(int opc, int dtype, const void* a, void* out, const EwShape& s, double alpha);
int ew_dispatch2(int opc, int dtype, const void* a, const void* b, void* out, const EwShape& s, double alpha);
int ew_dispatch3(int opc, int dtype, const void* a, const void* b, const void* c, void* out, const EwShape& s,
                 double alpha);

// (evict-first ld/st.cs variants of the flat kernel were measured: no faster, 3-8 % slower for add/gelu)
#define LG_EW_LD(p) (*(p))
#define LG_EW_ST(p, v) (*(p) = (v))

// ---- flat ---------------------------------------------------------------------------------------
template <class Op, typename T, int NIN, int V>
__global__ void __launch_bounds__(256) ew_flat_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                      const T* __restrict__ c, T* __restrict__ out, int64_t n,
                                                      T alpha) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    constexpr int U = 4;
    const int64_t nv = n / V;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const VT* av = reinterpret_cast<const VT*>(a);
    const VT* bv = reinterpret_cast<const VT*>(b);
    const VT* cv = reinterpret_cast<const VT*>(c);
    VT* ov = reinterpret_cast<VT*>(out);
    int64_t i = tid;
    for (; i + (U - 1) * nthreads < nv; i += U * nthreads) {
        VT ra[U], rb[U], rc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ra[u] = LG_EW_LD(av + i + u * nthreads);
            if (NIN > 1) rb[u] = LG_EW_LD(bv + i + u * nthreads);
            if (NIN > 2) rc[u] = LG_EW_LD(cv + i + u * nthreads);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            VT r;
#pragma unroll
            for (int k = 0; k < V; ++k)
                r.v[k] = Op::apply(ra[u].v[k], NIN > 1 ? rb[u].v[k] : T(0), NIN > 2 ? rc[u].v[k] : T(0), alpha);
            LG_EW_ST(ov + i + u * nthreads, r);
        }
    }
    for (; i < nv; i += nthreads) {
        VT ra = av[i], rb, rc, r;
        if (NIN > 1) rb = bv[i];
        if (NIN > 2) rc = cv[i];
#pragma unroll
        for (int k = 0; k < V; ++k)
            r.v[k] = Op::apply(ra.v[k], NIN > 1 ? rb.v[k] : T(0), NIN > 2 ? rc.v[k] : T(0), alpha);
        ov[i] = r;
    }
    for (int64_t j = nv * V + tid; j < n; j += nthreads)
        out[j] = Op::apply(a[j], NIN > 1 ? b[j] : T(0), NIN > 2 ? c[j] : T(0), alpha);
}

// flat, chunked: a CTA walks contiguous chunks of 256 x U vectors (U x 4 KB of every operand) instead of U
// grid-strided vectors per thread -- the 1-read / 1-write streams of the unary operators (relu, exp, gelu) then touch
// one contiguous read region and one contiguous write region per CTA at a time
template <class Op, typename T, int NIN, int V, int U>
__global__ void __launch_bounds__(256) ew_flat_chunk_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                            const T* __restrict__ c, T* __restrict__ out, int64_t n,
                                                            T alpha) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    const int64_t nv = n / V;
    const VT* av = reinterpret_cast<const VT*>(a);
    const VT* bv = reinterpret_cast<const VT*>(b);
    const VT* cv = reinterpret_cast<const VT*>(c);
    VT* ov = reinterpret_cast<VT*>(out);
    constexpr int64_t CH = 256 * U;
    for (int64_t base = (int64_t)blockIdx.x * CH; base < nv; base += (int64_t)gridDim.x * CH) {
        VT ra[U], rb[U], rc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * 256 + threadIdx.x;
            if (i < nv) {
                ra[u] = LG_EW_LD(av + i);
                if (NIN > 1) rb[u] = LG_EW_LD(bv + i);
                if (NIN > 2) rc[u] = LG_EW_LD(cv + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * 256 + threadIdx.x;
            if (i < nv) {
                VT r;
#pragma unroll
                for (int k = 0; k < V; ++k)
                    r.v[k] = Op::apply(ra[u].v[k], NIN > 1 ? rb[u].v[k] : T(0), NIN > 2 ? rc[u].v[k] : T(0), alpha);
                LG_EW_ST(ov + i, r);
            }
        }
    }
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t j = nv * V + tid; j < n; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = Op::apply(a[j], NIN > 1 ? b[j] : T(0), NIN > 2 ? c[j] : T(0), alpha);
}

// ---- nd_vec -------------------------------------------------------------------------------------
// inner (last) dim: every operand has stride 1 or 0 there, out has stride 1, inner % V == 0.
template <class Op, typename T, int NIN, int V, typename I>
__global__ void __launch_bounds__(256) ew_ndvec_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                       const T* __restrict__ c, T* __restrict__ out, EwShape s,
                                                       int64_t total_vecs, T alpha) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    constexpr int U = 4;   // independent vectors in flight per thread
    const int nd = s.ndim;
    const I inner_vecs = (I)(s.shape[nd - 1] / V);
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < total_vecs; w0 += U * nthreads) {
        VT ra[U], rb[U], rc[U];
        int64_t oo[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t w = w0 + u * nthreads;
            live[u] = w < total_vecs;
            if (!live[u]) continue;
            I rest = (I)w;
            I iv = rest % inner_vecs;
            rest /= inner_vecs;
            int64_t oa = (int64_t)iv * V * s.st[0][nd - 1];
            int64_t ob = NIN > 1 ? (int64_t)iv * V * s.st[1][nd - 1] : 0;
            int64_t oc = NIN > 2 ? (int64_t)iv * V * s.st[2][nd - 1] : 0;
            oo[u] = (int64_t)iv * V;
            for (int d = nd - 2; d >= 0; --d) {
                I dim = (I)s.shape[d];
                I q = rest / dim;
                I r = rest - q * dim;
                rest = q;
                oa += (int64_t)r * s.st[0][d];
                if (NIN > 1) ob += (int64_t)r * s.st[1][d];
                if (NIN > 2) oc += (int64_t)r * s.st[2][d];
                oo[u] += (int64_t)r * s.st[3][d];
            }
            if (s.st[0][nd - 1] == 1) ra[u] = *reinterpret_cast<const VT*>(a + oa);
            else {
                T x = a[oa];
#pragma unroll
                for (int k = 0; k < V; ++k) ra[u].v[k] = x;
            }
            if (NIN > 1) {
                if (s.st[1][nd - 1] == 1) rb[u] = *reinterpret_cast<const VT*>(b + ob);
                else {
                    T x = b[ob];
#pragma unroll
                    for (int k = 0; k < V; ++k) rb[u].v[k] = x;
                }
            }
            if (NIN > 2) {
                if (s.st[2][nd - 1] == 1) rc[u] = *reinterpret_cast<const VT*>(c + oc);
                else {
                    T x = c[oc];
#pragma unroll
                    for (int k = 0; k < V; ++k) rc[u].v[k] = x;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            VT r;
#pragma unroll
            for (int k = 0; k < V; ++k)
                r.v[k] = Op::apply(ra[u].v[k], NIN > 1 ? rb[u].v[k] : T(0), NIN > 2 ? rc[u].v[k] : T(0), alpha);
            *reinterpret_cast<VT*>(out + oo[u]) = r;
        }
    }
}

// ---- nd_any -------------------------------------------------------------------------------------
template <class Op, typename T, int NIN, typename I>
__global__ void __launch_bounds__(256) ew_ndany_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                       const T* __restrict__ c, T* __restrict__ out, EwShape s,
                                                       int64_t total, T alpha) {
    LG_PDL_TRIGGER();
    const int nd = s.ndim;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (int64_t)gridDim.x * blockDim.x) {
        I rest = (I)w;
        int64_t oa = 0, ob = 0, oc = 0, oo = 0;
        for (int d = nd - 1; d >= 0; --d) {
            I dim = (I)s.shape[d];
            I q = rest / dim;
            I r = rest - q * dim;
            rest = q;
            oa += (int64_t)r * s.st[0][d];
            if (NIN > 1) ob += (int64_t)r * s.st[1][d];
            if (NIN > 2) oc += (int64_t)r * s.st[2][d];
            oo += (int64_t)r * s.st[3][d];
        }
        out[oo] = Op::apply(a[oa], NIN > 1 ? b[ob] : T(0), NIN > 2 ? c[oc] : T(0), alpha);
    }
}

// ---- host-side shape canonicalisation -------------------------------------------------------------
// Drops size-1 dims and merges neighbours that are jointly contiguous for every operand.
inline void ew_collapse(int ndim, const int64_t* shape, const int64_t* const st_in[4], int nops_mask, EwShape& o) {
    int64_t shp[LG_MAX_DIMS], st[4][LG_MAX_DIMS];
    int n = 0;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] == 1) continue;
        shp[n] = shape[d];
        for (int k = 0; k < 4; ++k) st[k][n] = st_in[k] ? st_in[k][d] : 0;
        ++n;
    }
    if (n == 0) {
        o.ndim = 1;
        o.shape[0] = 1;
        for (int k = 0; k < 4; ++k) o.st[k][0] = 1;
        return;
    }
    int m = 0;
    o.shape[0] = shp[0];
    for (int k = 0; k < 4; ++k) o.st[k][0] = st[k][0];
    for (int d = 1; d < n; ++d) {
        bool merge = true;
        for (int k = 0; k < 4; ++k) {
            if (!((nops_mask >> k) & 1)) continue;
            if (o.st[k][m] != shp[d] * st[k][d]) {
                merge = false;
                break;
            }
        }
        if (merge) {
            o.shape[m] *= shp[d];
            for (int k = 0; k < 4; ++k) o.st[k][m] = st[k][d];
        } else {
            ++m;
            o.shape[m] = shp[d];
            for (int k = 0; k < 4; ++k) o.st[k][m] = st[k][d];
        }
    }
    o.ndim = m + 1;
}

inline void contiguous_strides(int ndim, const int64_t* shape, int64_t* st) {
    int64_t acc = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        st[d] = acc;
        acc *= shape[d];
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Launch one op over a canonicalised problem.
template <class Op, typename T, int NIN>
int ew_launch(const void* a_, const void* b_, const void* c_, void* out_, const EwShape& s, double alpha_) {
    constexpr int V = 16 / sizeof(T);
    const T* a = (const T*)a_;
    const T* b = (const T*)b_;
    const T* c = (const T*)c_;
    T* out = (T*)out_;
    T alpha = (T)alpha_;
    int64_t total = 1;
    for (int d = 0; d < s.ndim; ++d) total *= s.shape[d];
    if (total == 0) return 0;
    const int nd = s.ndim;
    const void* ptrs[4] = {a_, NIN > 1 ? b_ : nullptr, NIN > 2 ? c_ : nullptr, out_};
    // flat?
    bool flat = (nd == 1);
    for (int k = 0; k < 4 && flat; ++k)
        if (ptrs[k] && s.st[k][0] != 1) flat = false;
    if (flat) {
        bool al = true;
        for (int k = 0; k < 4; ++k)
            if (ptrs[k] && !aligned16(ptrs[k])) al = false;
        // LG_EW_FLAT_MODE: 2 (default) contiguous chunks of 8 vectors per thread, 1 chunks of 4, 0 grid-strided vectors;
        // LG_EW_BPS: CTAs per SM the grid is sized for.  Measured at 2^26-2^28 fp32 (profiles/r2_ew_flat_variants.txt):
        // grid-strided x4 at 8 CTAs/SM (round 1): relu 5.8-6.0, exp 5.7, gelu 5.6-5.7, add 6.3-6.45 TB/s;
        // chunks of 8 at 16 CTAs/SM: relu 6.3-6.4, exp 6.3-6.4, gelu 6.2-6.3, add 6.63 TB/s.
        static const int flat_mode = getenv("LG_EW_FLAT_MODE") ? atoi(getenv("LG_EW_FLAT_MODE")) : 2;
        static const int bps = getenv("LG_EW_BPS") ? atoi(getenv("LG_EW_BPS")) : 16;
        if (al && flat_mode == 1) {
            int grid = grid_for((total / V + 3) / 4 + 1, 256, bps);
            ew_flat_chunk_kernel<Op, T, NIN, V, 4><<<grid, 256, 0, stream()>>>(a, b, c, out, total, alpha);
        } else if (al && flat_mode == 2) {
            int grid = grid_for((total / V + 7) / 8 + 1, 256, bps);
            ew_flat_chunk_kernel<Op, T, NIN, V, 8><<<grid, 256, 0, stream()>>>(a, b, c, out, total, alpha);
        } else if (al) {
            int grid = grid_for((total / V + 3) / 4 + 1, 256, bps);
            ew_flat_kernel<Op, T, NIN, V><<<grid, 256, 0, stream()>>>(a, b, c, out, total, alpha);
        } else {
            int grid = grid_for((total + 3) / 4 + 1, 256, 8);
            ew_flat_kernel<Op, T, NIN, 1><<<grid, 256, 0, stream()>>>(a, b, c, out, total, alpha);
        }
        LG_CHECK_LAUNCH();
        return 0;
    }
    // largest offset any operand reaches decides the index width
    int64_t max_off = total;
    for (int k = 0; k < 4; ++k) {
        if (!ptrs[k]) continue;
        int64_t off = 0;
        for (int d = 0; d < nd; ++d) off += (s.shape[d] - 1) * (s.st[k][d] < 0 ? -s.st[k][d] : s.st[k][d]);
        if (off > max_off) max_off = off;
    }
    const bool small = max_off < (int64_t)0x7fffffff;
    // vectorisable inner dim?
    bool vec = (s.shape[nd - 1] % V == 0) && s.st[3][nd - 1] == 1;
    for (int k = 0; k < 4 && vec; ++k) {
        if (!ptrs[k]) continue;
        int64_t is = s.st[k][nd - 1];
        if (is != 0 && is != 1) vec = false;
        if (is == 1) {
            if (!aligned16(ptrs[k])) vec = false;
            for (int d = 0; d < nd - 1; ++d)
                if (s.st[k][d] % V != 0) vec = false;
        }
    }
    if (vec) {
        int64_t tv = total / V;
        int grid = grid_for((tv + 3) / 4, 256, 8);
        if (small)
            ew_ndvec_kernel<Op, T, NIN, V, uint32_t><<<grid, 256, 0, stream()>>>(a, b, c, out, s, tv, alpha);
        else
            ew_ndvec_kernel<Op, T, NIN, V, int64_t><<<grid, 256, 0, stream()>>>(a, b, c, out, s, tv, alpha);
    } else {
        int grid = grid_for(total, 256, 8);
        if (small)
            ew_ndany_kernel<Op, T, NIN, uint32_t><<<grid, 256, 0, stream()>>>(a, b, c, out, s, total, alpha);
        else
            ew_ndany_kernel<Op, T, NIN, int64_t><<<grid, 256, 0, stream()>>>(a, b, c, out, s, total, alpha);
    }
    LG_CHECK_LAUNCH();
    return 0;
}

}  // namespace lg
