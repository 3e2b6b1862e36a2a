// On-device optimizer updates over flat fp32 arenas (one launch per step instead of the reference's
// per-parameter Python loop of 3-14 elementwise launches, lightgrad/optim.py:10-52).
// Arithmetic restates optim.py term by term (python-float coefficients become fp32 scalars, as numpy's
// weak-scalar promotion does):
//   SGD       delta = -lr*g + momentum*delta_prev ; p += delta                       (optim.py:23-25)
//   Adam      m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2
//             p += -lr * (m/(1-b1^t)) / ((v/(1-b2^t))^0.5 + eps)                      (optim.py:35-41)
//   AdaBelief as Adam with v = b2*v + (1-b2)*(g-m)^2                                  (optim.py:47-52)
// The reference increments t once per PARAMETER per step (optim.py:36-37), so the
// bias corrections differ per tensor: they arrive as per-segment scalars.
// Algorithmic bytes per parameter: SGD 12 (+8 with momentum), Adam 28.
#include "lg_ew.cuh"
#include "lg_adam.cuh"

using namespace lg;
using namespace lg::adam;

namespace {

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ delta, int64_t n, float neg_lr, float mom) {
    LG_PDL_TRIGGER();
    const int64_t nv = n / 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += nt) {
        float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<const float4*>(g)[i], dv;
        if (delta) {
            float4 pd = reinterpret_cast<float4*>(delta)[i];
            dv.x = neg_lr * gv.x + mom * pd.x; dv.y = neg_lr * gv.y + mom * pd.y;
            dv.z = neg_lr * gv.z + mom * pd.z; dv.w = neg_lr * gv.w + mom * pd.w;
            reinterpret_cast<float4*>(delta)[i] = dv;
        } else {
            dv.x = neg_lr * gv.x; dv.y = neg_lr * gv.y; dv.z = neg_lr * gv.z; dv.w = neg_lr * gv.w;
        }
        pv.x += dv.x; pv.y += dv.y; pv.z += dv.z; pv.w += dv.w;
        reinterpret_cast<float4*>(p)[i] = pv;
    }
    for (int64_t j = nv * 4 + tid; j < n; j += nt) {
        float d = neg_lr * g[j] + (delta ? mom * delta[j] : 0.0f);
        if (delta) delta[j] = d;
        p[j] += d;
    }
}

// U float4 per array and thread are loaded before any arithmetic (4*U independent 16-byte loads in flight)
template <bool BELIEF, int U>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, int n_seg,
                                                   const int64_t* __restrict__ seg_end, int64_t seg_base,
                                                   const float* __restrict__ c1,
                                                   const float* __restrict__ c2, float neg_lr, float b1, float b2,
                                                   float omb1, float omb2, float eps) {
    LG_PDL_TRIGGER();
    const int64_t nv = n / 4;
    // a CTA walks contiguous runs of U x 256 float4 (U x 4 KB) of every arena
    const int64_t nt = (int64_t)gridDim.x * blockDim.x;
    int seg = 0;
    for (int64_t i0 = (int64_t)blockIdx.x * (U * 256) + threadIdx.x; i0 < nv; i0 += nt * U) {
        float4 pv[U], gv[U], mv[U], vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < nv) {
                pv[u] = reinterpret_cast<float4*>(p)[i];
                gv[u] = reinterpret_cast<const float4*>(g)[i];
                mv[u] = reinterpret_cast<float4*>(m)[i];
                vv[u] = reinterpret_cast<float4*>(v)[i];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < nv) {
                seg = adam_segment(seg, i * 4 + seg_base, n_seg, seg_end);
                adam_update4<BELIEF>(pv[u], gv[u], mv[u], vv[u], c1[seg], c2[seg], c1[2 * n_seg + seg], c2[2 * n_seg + seg],
                                     neg_lr, b1, b2, omb1, omb2, eps);
                reinterpret_cast<float4*>(p)[i] = pv[u];
                reinterpret_cast<float4*>(m)[i] = mv[u];
                reinterpret_cast<float4*>(v)[i] = vv[u];
            }
        }
    }
}

}  // namespace

extern "C" {

int lg_sgd_step(void* param, const void* grad, void* delta, int64_t n, double lr, double momentum) {
    LG_INIT();
    if (n == 0) return 0;
    LG_REQUIRE(aligned16(param) && aligned16(grad) && (!delta || aligned16(delta)), "lg_sgd_step: arenas must be 16-byte aligned");
    sgd_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, stream()>>>((float*)param, (const float*)grad, (float*)delta, n,
                                                                  (float)(-lr), (float)momentum);
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_adam_step(int belief, void* param, const void* grad, void* m, void* v, int64_t n, int n_seg,
                 const int64_t* seg_end_dev, int64_t* t_dev, double lr, double beta1, double beta2, double eps,
                 int64_t seg_base, int seg_offset, int t_advance) {
    LG_INIT();
    if (n == 0) return 0;
    LG_REQUIRE(n_seg >= 1, "lg_adam_step: need at least one segment");
    LG_REQUIRE(n % 4 == 0 && aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v),
               "lg_adam_step: arenas must be 16-byte aligned and padded to a multiple of 4 elements");
    float* corr = (float*)tmp_alloc(4 * (size_t)n_seg * sizeof(float));   // c1 | c2 | 1/c1 | 1/c2
    if (!corr) return 1;
    float* seg_c1_dev = corr;
    float* seg_c2_dev = corr + n_seg;
    adam_prep_kernel<<<1, 256, 0, stream()>>>(n_seg, t_dev, beta1, beta2, seg_c1_dev, seg_c2_dev, seg_offset, t_advance);
    count_launch();
    static const int U = getenv("LG_ADAM_U") ? atoi(getenv("LG_ADAM_U")) : 1;
    static const int bps = getenv("LG_ADAM_BPS") ? atoi(getenv("LG_ADAM_BPS")) : 8;
    int grid = grid_for(n / 4 / (U > 1 ? U : 1) + 1, 256, bps);
    float b1 = (float)beta1, b2 = (float)beta2;
    // (1 - beta) is formed in double by python and then rounded to fp32, as numpy does with the scalar
    float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
#define LG_ADAM(B_, U_)                                                                                          \
    adam_kernel<B_, U_><<<grid, 256, 0, stream()>>>((float*)param, (const float*)grad, (float*)m, (float*)v, n, n_seg, \
                                                    seg_end_dev, seg_base, seg_c1_dev, seg_c2_dev, (float)(-lr), b1, b2,  \
                                                    omb1,                                                              \
                                                    omb2, (float)eps)
    if (belief) {
        if (U == 1) LG_ADAM(true, 1); else if (U == 2) LG_ADAM(true, 2); else LG_ADAM(true, 4);
    } else {
        if (U == 1) LG_ADAM(false, 1); else if (U == 2) LG_ADAM(false, 2); else LG_ADAM(false, 4);
    }
#undef LG_ADAM
    tmp_free(corr);
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
