// dtype conversion and strided gather-copy.
// Replaces OpenCLTensor.contiguous (opencl/tensor.py:103-116, an `atom('o = a')` launch) and the
// implicit numpy astype of CpuTensor (cpu/tensor.py:8-14).  Same-type float copies go through the
// vectorised elementwise engine; everything else is one element per thread.
#include "lg_ew.cuh"
#include <cuda_bf16.h>

using namespace lg;

namespace {

// fp32 -> bf16 staging copy of a tensor-core GEMM operand (round to nearest even): 8 elements per thread and
// iteration, two 16-byte loads in flight per 16-byte store.  Algorithmic bytes: 6 per element.
__global__ void __launch_bounds__(256) f32_to_bf16_flat_kernel(const float* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dst, int64_t n) {
    LG_PDL_TRIGGER();
    const int64_t nv = n / 8;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += nt) {
        const float4 lo = reinterpret_cast<const float4*>(src)[2 * i];
        const float4 hi = reinterpret_cast<const float4*>(src)[2 * i + 1];
        __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x, lo.y), p1 = __floats2bfloat162_rn(lo.z, lo.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(hi.x, hi.y), p3 = __floats2bfloat162_rn(hi.z, hi.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0);
        o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2);
        o.w = *reinterpret_cast<uint32_t*>(&p3);
        reinterpret_cast<uint4*>(dst)[i] = o;
    }
    for (int64_t j = nv * 8 + tid; j < n; j += nt) dst[j] = __float2bfloat16_rn(src[j]);
}

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
    if constexpr (std::is_same<S, float>::value && std::is_same<D, __nv_bfloat16>::value) {
        if (s.ndim == 1 && s.st[0][0] == 1 && s.st[3][0] == 1 && aligned16(src) && aligned16(dst)) {
            f32_to_bf16_flat_kernel<<<grid_for(total / 8 + 1, 256, 8), 256, 0, stream()>>>(
                (const float*)src, (__nv_bfloat16*)dst, total);
            LG_CHECK_LAUNCH();
            return 0;
        }
    }
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
        case LG_BF16:
            if constexpr (std::is_same<S, float>::value) return cast_launch<S, __nv_bfloat16>(src, dst, s);
            break;
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
        case LG_BF16:
            // staging copies are only ever read back for inspection
            if (dd == LG_F32) return cast_launch<__nv_bfloat16, float>(src, dst, s);
            break;
    }
    return set_error("lg_cast: unsupported source dtype %d", sd);
}
