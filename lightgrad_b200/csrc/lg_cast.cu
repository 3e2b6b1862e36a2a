// dtype conversion and strided gather-copy.
// Replaces OpenCLTensor.contiguous (opencl/tensor.py:103-116, an `atom('o = a')` launch) and the
// implicit numpy astype of CpuTensor (cpu/tensor.py:8-14).  Same-type float copies go through the
// vectorised elementwise engine; everything else is one element per thread.
#include "lg_ew.cuh"

using namespace lg;

namespace {

template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_flat_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
    LG_PDL_TRIGGER();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (D)src[i];
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_nd_kernel(const S* __restrict__ src, D* __restrict__ dst, EwShape s,
                                                      int64_t total) {
    LG_PDL_TRIGGER();
    const int nd = s.ndim;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (int64_t)gridDim.x * blockDim.x) {
        int64_t rest = w, os = 0, od = 0;
        for (int d = nd - 1; d >= 0; --d) {
            int64_t q = rest / s.shape[d];
            int64_t r = rest - q * s.shape[d];
            rest = q;
            os += r * s.st[0][d];
            od += r * s.st[3][d];
        }
        dst[od] = (D)src[os];
    }
}

template <typename S, typename D>
int cast_launch(const void* src, void* dst, const EwShape& s) {
    int64_t total = 1;
    for (int d = 0; d < s.ndim; ++d) total *= s.shape[d];
    if (total == 0) return 0;
    int grid = grid_for(total, 256, 8);
    if (s.ndim == 1 && s.st[0][0] == 1 && s.st[3][0] == 1)
        cast_flat_kernel<S, D><<<grid, 256, 0, stream()>>>((const S*)src, (D*)dst, total);
    else
        cast_nd_kernel<S, D><<<grid, 256, 0, stream()>>>((const S*)src, (D*)dst, s, total);
    LG_CHECK_LAUNCH();
    return 0;
}

template <typename S>
int cast_from(int dd, const void* src, void* dst, const EwShape& s) {
    switch (dd) {
        case LG_F32: return cast_launch<S, float>(src, dst, s);
        case LG_F64: return cast_launch<S, double>(src, dst, s);
        case LG_I32: return cast_launch<S, int32_t>(src, dst, s);
        case LG_I64: return cast_launch<S, int64_t>(src, dst, s);
        case LG_I16: return cast_launch<S, int16_t>(src, dst, s);
        case LG_U8: return cast_launch<S, uint8_t>(src, dst, s);
        case LG_I8: return cast_launch<S, int8_t>(src, dst, s);
    }
    return set_error("lg_cast: unsupported destination dtype %d", dd);
}

}  // namespace

extern "C" int lg_cast(int sd, int dd, int ndim, const int64_t* shape, const void* src, const int64_t* ssrc,
                       void* dst, const int64_t* sdst) {
    LG_INIT();
    LG_REQUIRE(ndim >= 0 && ndim <= LG_MAX_DIMS, "lg_cast: ndim %d exceeds %d", ndim, LG_MAX_DIMS);
    int64_t contig[LG_MAX_DIMS];
    contiguous_strides(ndim, shape, contig);
    const int64_t* st[4] = {ssrc ? ssrc : contig, nullptr, nullptr, sdst ? sdst : contig};
    EwShape s;
    ew_collapse(ndim, shape, st, 1 | 8, s);
    if (sd == dd && (sd == LG_F32 || sd == LG_I32)) return ew_dispatch1(LG_EW_COPY, LG_F32, src, dst, s, 0.0);
    if (sd == dd && (sd == LG_F64 || sd == LG_I64)) return ew_dispatch1(LG_EW_COPY, LG_F64, src, dst, s, 0.0);
    switch (sd) {
        case LG_F32: return cast_from<float>(dd, src, dst, s);
        case LG_F64: return cast_from<double>(dd, src, dst, s);
        case LG_I32: return cast_from<int32_t>(dd, src, dst, s);
        case LG_I64: return cast_from<int64_t>(dd, src, dst, s);
        case LG_I16: return cast_from<int16_t>(dd, src, dst, s);
        case LG_U8: return cast_from<uint8_t>(dd, src, dst, s);
        case LG_I8: return cast_from<int8_t>(dd, src, dst, s);
    }
    return set_error("lg_cast: unsupported source dtype %d", sd);
}
