// Exact-precision matmul on the FFMA/DFMA pipes (the "exact-fp32 SIMT mode").
// Replaces the reference's tiled OpenCL SGEMM (opencl/kernels.py:201-337), which zero-pads operands
// to multiples of 128 with extra kernels, materialises transposed operands and slices the result.
// Here operands are described by element strides: transposes, head-split views and broadcast batch
// dims are consumed in place, ragged edges are predicated, and bias / accumulate are fused.
//   tile 128x128x16, 256 threads, 8x8 register micro-tile (two 4-wide halves so that shared-memory
//   reads are conflict-free 128-bit), register-staged double buffering, one barrier per k-tile.
// Roofline: FP32 FMA pipe (148 SM x 128 lanes x 2 x clock).  Algorithmic flops 2*M*N*K per batch.
#include "lg_common.cuh"

using namespace lg;

namespace {

template <typename T, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C, const T* __restrict__ bias,
                 LgGemmDesc d, int accumulate, int a_m_fast, int b_n_fast) {
    LG_PDL_TRIGGER();
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int HM = TM / 2, HN = TN / 2;       // micro-tile halves
    constexpr int LA = BM * BK / NT, LB = BN * BK / NT;  // elements per thread per tile
    __shared__ T As[2][BK][BM + 4];
    __shared__ T Bs[2][BK][BN + 4];

    const int t = threadIdx.x;
    const int tx = t % (BN / TN), ty = t / (BN / TN);
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t bz = blockIdx.z;
    const int64_t b0 = bz / d.batch1, b1 = bz % d.batch1;
    A += b0 * d.sa_b0 + b1 * d.sa_b1;
    B += b0 * d.sb_b0 + b1 * d.sb_b1;
    C += b0 * d.sc_b0 + b1 * d.sc_b1;

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

    T ra[LA], rb[LB];
    auto load_tile = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int e = t + i * NT;
            int mm, kk;
            if (a_m_fast) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            int64_t gm = m0 + mm, gk = k0 + kk;
            ra[i] = (gm < d.M && gk < d.K) ? A[gm * d.sa_m + gk * d.sa_k] : T(0);
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int e = t + i * NT;
            int nn, kk;
            if (b_n_fast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
            int64_t gn = n0 + nn, gk = k0 + kk;
            rb[i] = (gn < d.N && gk < d.K) ? B[gk * d.sb_k + gn * d.sb_n] : T(0);
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int e = t + i * NT;
            int mm, kk;
            if (a_m_fast) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            As[buf][kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int e = t + i * NT;
            int nn, kk;
            if (b_n_fast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
            Bs[buf][kk][nn] = rb[i];
        }
    };

    const int64_t ktiles = (d.K + BK - 1) / BK;
    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) load_tile((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < HM; ++i) {
                a[i] = As[cur][kk][ty * HM + i];
                a[HM + i] = As[cur][kk][BM / 2 + ty * HM + i];
            }
#pragma unroll
            for (int j = 0; j < HN; ++j) {
                b[j] = Bs[cur][kk][tx * HN + j];
                b[HN + j] = Bs[cur][kk][BN / 2 + tx * HN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            store_tile(cur ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t gm = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
        if (gm >= d.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t gn = n0 + (j < HN ? tx * HN + j : BN / 2 + tx * HN + (j - HN));
            if (gn >= d.N) continue;
            T v = acc[i][j];
            if (bias) v += bias[gn];
            T* p = C + gm * d.sc_m + gn * d.sc_n;
            if (accumulate) v += *p;
            *p = v;
        }
    }
}

template <typename T, int BM, int BN, int BK, int TM, int TN>
int launch(const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int accumulate) {
    int64_t batch = d->batch0 * d->batch1;
    if (d->M == 0 || d->N == 0 || batch == 0) return 0;
    LG_REQUIRE(batch <= 65535, "lg_gemm: batch %lld exceeds 65535", (long long)batch);
    dim3 grid((unsigned)((d->N + BN - 1) / BN), (unsigned)((d->M + BM - 1) / BM), (unsigned)batch);
    LG_REQUIRE((d->M + BM - 1) / BM <= 65535, "lg_gemm: M too large for the SIMT grid");
    int a_m_fast = (d->sa_m == 1 && d->sa_k != 1) ? 1 : 0;
    int b_n_fast = (d->sb_n == 1) ? 1 : 0;
    gemm_simt_kernel<T, BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, stream()>>>(
        (const T*)a, (const T*)b, (T*)c, (const T*)bias, *d, accumulate, a_m_fast, b_n_fast);
    LG_CHECK_LAUNCH();
    return 0;
}

}  // namespace

namespace lg {
int gemm_simt(int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
              int accumulate) {
    if (dtype == LG_F32) {
        // small problems: 64x64 tiles keep more SMs busy
        if (d->M * d->N * d->batch0 * d->batch1 <= (int64_t)64 * 64 * 148 * 4)
            return launch<float, 64, 64, 16, 4, 4>(d, a, b, c, bias, accumulate);
        return launch<float, 128, 128, 16, 8, 8>(d, a, b, c, bias, accumulate);
    }
    if (dtype == LG_F64) return launch<double, 64, 64, 16, 4, 4>(d, a, b, c, bias, accumulate);
    return set_error("lg_gemm: unsupported dtype %d", dtype);
}
}  // namespace lg
