// Sliding-window unfold / fold for convolutions (n-dimensional, n <= 4 window dims).
// The reference lowers conv to a matmul over an unfolded copy of the input (cpu/ops.py:298-322: a strided window
// view, reshaped to (positions, kernel elements)) and, in backward, folds the matmul's result back by adding every
// kernel offset's plane into the input gradient from a python loop (cpu/ops.py:339-355).  Here:
//   lg_im2col  cols[l, p_0..p_{n-1}, k_0..k_{n-1}] = x[l, p_0 s_0 + k_0, ..., p_{n-1} s_{n-1} + k_{n-1}]
//   lg_col2im  dx[l, i_0..i_{n-1}] = sum over (p, k) with p_d s_d + k_d = i_d of cols[l, p, k]
// col2im is a GATHER (one thread per input element, looping over the <= prod ceil(k_d / s_d) windows that cover it):
// no atomics, deterministic summation order, one launch instead of prod(k_d) strided adds.
// Roofline: HBM; algorithmic bytes = 4 (read or written) per element of `cols` plus 4 per element of x / dx.
#include "lg_common.cuh"

using namespace lg;

namespace {

constexpr int MAXW = 4;
struct WinShape {
    int n;                       // window dims
    int64_t lead;                // product of the leading (batch-like) dims
    int64_t in[MAXW], k[MAXW], s[MAXW], pos[MAXW];    // input extent, kernel extent, stride, number of positions
    int64_t in_total, pos_total, k_total;             // products
};

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, T* __restrict__ cols, WinShape w,
                                                     int64_t total) {
    LG_PDL_TRIGGER();
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        int64_t rest = o;
        int64_t kk[MAXW], pp[MAXW];
#pragma unroll
        for (int d = MAXW - 1; d >= 0; --d)
            if (d < w.n) { kk[d] = rest % w.k[d]; rest /= w.k[d]; }
#pragma unroll
        for (int d = MAXW - 1; d >= 0; --d)
            if (d < w.n) { pp[d] = rest % w.pos[d]; rest /= w.pos[d]; }
        int64_t src = rest;                       // leading index
#pragma unroll
        for (int d = 0; d < MAXW; ++d)
            if (d < w.n) src = src * w.in[d] + pp[d] * w.s[d] + kk[d];
        cols[o] = x[src];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ cols, T* __restrict__ dx, WinShape w,
                                                     int64_t total) {
    LG_PDL_TRIGGER();
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t rest = e, idx[MAXW];
#pragma unroll
        for (int d = MAXW - 1; d >= 0; --d)
            if (d < w.n) { idx[d] = rest % w.in[d]; rest /= w.in[d]; }
        const int64_t l = rest;
        // windows covering this element along every dim: p in [lo_d, hi_d], k = i - p s
        int64_t lo[MAXW], hi[MAXW], p[MAXW];
        bool any = true;
#pragma unroll
        for (int d = 0; d < MAXW; ++d) {
            if (d < w.n) {
                const int64_t i = idx[d];
                int64_t a = i - w.k[d] + 1;
                a = a <= 0 ? 0 : (a + w.s[d] - 1) / w.s[d];
                int64_t b = i / w.s[d];
                if (b > w.pos[d] - 1) b = w.pos[d] - 1;
                lo[d] = a; hi[d] = b; p[d] = a;
                any = any && a <= b;
            } else { lo[d] = hi[d] = p[d] = 0; }
        }
        T acc = T(0);
        while (any) {
            int64_t pos_lin = 0, k_lin = 0;
#pragma unroll
            for (int d = 0; d < MAXW; ++d)
                if (d < w.n) {
                    pos_lin = pos_lin * w.pos[d] + p[d];
                    k_lin = k_lin * w.k[d] + (idx[d] - p[d] * w.s[d]);
                }
            acc += cols[(l * w.pos_total + pos_lin) * w.k_total + k_lin];
            // odometer over the position ranges (last dim fastest: a fixed, deterministic summation order)
            int d = w.n - 1;
            for (; d >= 0; --d) {
                if (++p[d] <= hi[d]) break;
                p[d] = lo[d];
            }
            if (d < 0) break;
        }
        dx[e] = acc;
    }
}

int fill_shape(const char* who, int n, int64_t lead, const int64_t* in_dims, const int64_t* k_dims, const int64_t* strides,
               WinShape& w) {
    LG_REQUIRE(n >= 1 && n <= MAXW, "%s: 1..%d window dims (got %d)", who, MAXW, n);
    LG_REQUIRE(lead >= 0, "%s: negative leading extent", who);
    w.n = n;
    w.lead = lead;
    w.in_total = w.pos_total = w.k_total = 1;
    for (int d = 0; d < MAXW; ++d) {
        if (d < n) {
            LG_REQUIRE(k_dims[d] >= 1 && strides[d] >= 1 && in_dims[d] >= k_dims[d],
                       "%s: dim %d: kernel %lld / stride %lld do not fit the input extent %lld", who, d,
                       (long long)k_dims[d], (long long)strides[d], (long long)in_dims[d]);
            w.in[d] = in_dims[d]; w.k[d] = k_dims[d]; w.s[d] = strides[d];
            w.pos[d] = (in_dims[d] - k_dims[d]) / strides[d] + 1;
        } else {
            w.in[d] = w.k[d] = w.s[d] = w.pos[d] = 1;
        }
        w.in_total *= w.in[d]; w.pos_total *= w.pos[d]; w.k_total *= w.k[d];
    }
    return 0;
}

}  // namespace

extern "C" {

int lg_im2col(int dtype, int n, int64_t lead, const int64_t* in_dims, const int64_t* k_dims, const int64_t* strides,
              const void* x, void* cols) {
    LG_INIT();
    WinShape w;
    if (fill_shape("lg_im2col", n, lead, in_dims, k_dims, strides, w)) return 1;
    const int64_t total = lead * w.pos_total * w.k_total;
    if (total == 0) return 0;
    const int grid = grid_for(total, 256, 8);
    if (dtype == LG_F32) im2col_kernel<float><<<grid, 256, 0, stream()>>>((const float*)x, (float*)cols, w, total);
    else if (dtype == LG_F64) im2col_kernel<double><<<grid, 256, 0, stream()>>>((const double*)x, (double*)cols, w, total);
    else return set_error("lg_im2col: unsupported dtype %d", dtype);
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_col2im(int dtype, int n, int64_t lead, const int64_t* in_dims, const int64_t* k_dims, const int64_t* strides,
              const void* cols, void* dx) {
    LG_INIT();
    WinShape w;
    if (fill_shape("lg_col2im", n, lead, in_dims, k_dims, strides, w)) return 1;
    const int64_t total = lead * w.in_total;
    if (total == 0) return 0;
    const int grid = grid_for(total, 256, 8);
    if (dtype == LG_F32) col2im_kernel<float><<<grid, 256, 0, stream()>>>((const float*)cols, (float*)dx, w, total);
    else if (dtype == LG_F64) col2im_kernel<double><<<grid, 256, 0, stream()>>>((const double*)cols, (double*)dx, w, total);
    else return set_error("lg_col2im: unsupported dtype %d", dtype);
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
