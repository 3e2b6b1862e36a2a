// C-ABI entry points of the elementwise engine (see lg_ew.cuh for the kernels).
#include "lg_ew.cuh"
#include "lg_ew_ops.cuh"

using namespace lg;

namespace {

int nin_of(int op) { return op < 32 ? 1 : (op < 64 ? 2 : 3); }

int dispatch_op(int opc, int dtype, const void* a, const void* b, const void* c, void* out, const EwShape& s,
                double alpha) {
    switch (nin_of(opc)) {
        case 1: return lg::ew_dispatch1(opc, dtype, a, out, s, alpha);
        case 2: return lg::ew_dispatch2(opc, dtype, a, b, out, s, alpha);
        default: return lg::ew_dispatch3(opc, dtype, a, b, c, out, s, alpha);
    }
}

// fused two-output backward (mul / div), contiguous operands
template <typename T, int KIND, int V>
__global__ void __launch_bounds__(256) ew_bwd2_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                      const T* __restrict__ g, T* __restrict__ da,
                                                      T* __restrict__ db, int64_t n) {
    LG_PDL_TRIGGER();
    using VT = Vec<T, V>;
    const int64_t nv = n / V;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += nthreads) {
        VT ra = reinterpret_cast<const VT*>(a)[i], rb = reinterpret_cast<const VT*>(b)[i],
           rg = reinterpret_cast<const VT*>(g)[i], xa, xb;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (KIND == 0) {
                xa.v[k] = rg.v[k] * rb.v[k];
                xb.v[k] = ra.v[k] * rg.v[k];
            } else {
                xa.v[k] = rg.v[k] / rb.v[k];
                xb.v[k] = -ra.v[k] / (rb.v[k] * rb.v[k]) * rg.v[k];
            }
        }
        reinterpret_cast<VT*>(da)[i] = xa;
        reinterpret_cast<VT*>(db)[i] = xb;
    }
    for (int64_t j = nv * V + tid; j < n; j += nthreads) {
        if (KIND == 0) {
            da[j] = g[j] * b[j];
            db[j] = a[j] * g[j];
        } else {
            da[j] = g[j] / b[j];
            db[j] = -a[j] / (b[j] * b[j]) * g[j];
        }
    }
}

template <typename T>
int bwd2_launch(int kind, const void* a, const void* b, const void* g, void* da, void* db, int64_t n) {
    constexpr int V = 16 / sizeof(T);
    bool al = aligned16(a) && aligned16(b) && aligned16(g) && aligned16(da) && aligned16(db);
    int grid = grid_for(al ? n / V + 1 : n, 256, 8);
#define L(K, VV) ew_bwd2_kernel<T, K, VV><<<grid, 256, 0, stream()>>>((const T*)a, (const T*)b, (const T*)g, (T*)da, (T*)db, n)
    if (kind == 0) { if (al) L(0, V); else L(0, 1); }
    else { if (al) L(1, V); else L(1, 1); }
#undef L
    LG_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" {

int lg_ew_flat(int opc, int dtype, const void* a, const void* b, const void* c, void* out, int64_t n, double alpha) {
    LG_INIT();
    if (n == 0) return 0;
    EwShape s;
    s.ndim = 1;
    s.shape[0] = n;
    for (int k = 0; k < 4; ++k) s.st[k][0] = 1;
    return dispatch_op(opc, dtype, a, b, c, out, s, alpha);
}

int lg_ew(int opc, int dtype, int ndim, const int64_t* shape, const void* a, const int64_t* sa, const void* b,
          const int64_t* sb, const void* c, const int64_t* sc, void* out, const int64_t* so, double alpha) {
    LG_INIT();
    LG_REQUIRE(ndim >= 0 && ndim <= LG_MAX_DIMS, "lg_ew: ndim %d exceeds %d", ndim, LG_MAX_DIMS);
    int64_t contig[LG_MAX_DIMS];
    contiguous_strides(ndim, shape, contig);
    int nin = nin_of(opc);
    const int64_t* st[4] = {sa ? sa : contig, nin > 1 ? (sb ? sb : contig) : nullptr,
                            nin > 2 ? (sc ? sc : contig) : nullptr, so ? so : contig};
    int mask = 1 | (nin > 1 ? 2 : 0) | (nin > 2 ? 4 : 0) | 8;
    EwShape s;
    ew_collapse(ndim, shape, st, mask, s);
    return dispatch_op(opc, dtype, a, b, c, out, s, alpha);
}

int lg_ew_bwd2_flat(int kind, int dtype, const void* a, const void* b, const void* g, void* da, void* db, int64_t n) {
    LG_INIT();
    if (n == 0) return 0;
    LG_REQUIRE(kind == 0 || kind == 1, "lg_ew_bwd2_flat: kind must be 0 (mul) or 1 (div)");
    if (dtype == LG_F32) return bwd2_launch<float>(kind, a, b, g, da, db, n);
    if (dtype == LG_F64) return bwd2_launch<double>(kind, a, b, g, da, db, n);
    return set_error("lg_ew_bwd2_flat: unsupported dtype %d", dtype);
}

}  // extern "C"
