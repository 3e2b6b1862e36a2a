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

using namespace lg;

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

// per-tensor bias corrections 1 - beta^t with t = *t_dev + i + 1, evaluated in double like python does;
// one CTA, so the counter can be advanced after every thread has read it
__global__ void adam_prep_kernel(int n_seg, int64_t* __restrict__ t_dev, double b1, double b2,
                                 float* __restrict__ c1, float* __restrict__ c2) {
    LG_PDL_TRIGGER();
    const int64_t t0 = *t_dev;
    for (int i = threadIdx.x; i < n_seg; i += blockDim.x) {
        const double t = (double)(t0 + i + 1);
        c1[i] = (float)(1.0 - pow(b1, t));
        c2[i] = (float)(1.0 - pow(b2, t));
    }
    __syncthreads();
    if (threadIdx.x == 0) *t_dev = t0 + n_seg;
}

// arenas are padded so that every tensor starts on a 64-element boundary: a float4 never straddles tensors
template <bool BELIEF>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, int n_seg,
                                                   const int64_t* __restrict__ seg_end, const float* __restrict__ c1,
                                                   const float* __restrict__ c2, float neg_lr, float b1, float b2,
                                                   float omb1, float omb2, float eps) {
    LG_PDL_TRIGGER();
    const int64_t nv = n / 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    int seg = 0;
    for (int64_t i = tid; i < nv; i += nt) {
        const int64_t e = i * 4;
        if (!(e < seg_end[seg] && (seg == 0 || e >= seg_end[seg - 1]))) {
            int lo = 0, hi = n_seg - 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (seg_end[mid] > e) hi = mid; else lo = mid + 1;
            }
            seg = lo;
        }
        const float d1 = c1[seg], d2 = c2[seg];
        float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<const float4*>(g)[i],
               mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gi = gp[k];
            const float mi = b1 * mp[k] + omb1 * gi;
            const float r = BELIEF ? (gi - mi) : gi;
            const float vi = b2 * vp[k] + omb2 * (r * r);
            mp[k] = mi;
            vp[k] = vi;
            pp[k] += neg_lr * (mi / d1) / (sqrtf(vi / d2) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
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
                 const int64_t* seg_end_dev, int64_t* t_dev, double lr, double beta1, double beta2, double eps) {
    LG_INIT();
    if (n == 0) return 0;
    LG_REQUIRE(n_seg >= 1, "lg_adam_step: need at least one segment");
    LG_REQUIRE(n % 4 == 0 && aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v),
               "lg_adam_step: arenas must be 16-byte aligned and padded to a multiple of 4 elements");
    float* corr = (float*)tmp_alloc(2 * (size_t)n_seg * sizeof(float));
    if (!corr) return 1;
    float* seg_c1_dev = corr;
    float* seg_c2_dev = corr + n_seg;
    adam_prep_kernel<<<1, 256, 0, stream()>>>(n_seg, t_dev, beta1, beta2, seg_c1_dev, seg_c2_dev);
    count_launch();
    int grid = grid_for(n / 4, 256, 8);
    float b1 = (float)beta1, b2 = (float)beta2;
    // (1 - beta) is formed in double by python and then rounded to fp32, as numpy does with the scalar
    float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    if (belief)
        adam_kernel<true><<<grid, 256, 0, stream()>>>((float*)param, (const float*)grad, (float*)m, (float*)v, n,
                                                      n_seg, seg_end_dev, seg_c1_dev, seg_c2_dev, (float)(-lr), b1,
                                                      b2, omb1, omb2, (float)eps);
    else
        adam_kernel<false><<<grid, 256, 0, stream()>>>((float*)param, (const float*)grad, (float*)m, (float*)v, n,
                                                       n_seg, seg_end_dev, seg_c1_dev, seg_c2_dev, (float)(-lr), b1,
                                                       b2, omb1, omb2, (float)eps);
    tmp_free(corr);
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
