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


// ---- fast path (float32, whole 128x128x16 tiles, unit inner strides, 16-byte aligned rows) ------------------------------
// Same tile and micro-tile as above; what changes is how the operands reach shared memory: 128-bit global loads with
// affine addressing (no div / mod / bounds test per element in the k-loop) and a shared-memory layout swizzled at
// float4 granularity, m4' = m4 ^ (2 * (k >> 2)), under which the transposing stores of a K-contiguous operand (a thread
// holds 4 consecutive k of one row) hit 32 distinct banks and the 128-bit stores of an M-contiguous operand stay whole.
// KFAST: the operand is contiguous along K (x of x @ W^T, and W itself given as (N, K)); else contiguous along M / N.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256, 2)
gemm_simt_fast_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                      const float* __restrict__ bias, LgGemmDesc d, int accumulate) {
    LG_PDL_TRIGGER();
    constexpr int BM = 128, BN = 128, BK = 16;
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t bz = blockIdx.z;
    const int64_t b0 = bz / d.batch1, b1 = bz % d.batch1;
    A += b0 * d.sa_b0 + b1 * d.sa_b1;
    B += b0 * d.sb_b0 + b1 * d.sb_b1;
    C += b0 * d.sc_b0 + b1 * d.sc_b1;

    // per-thread source pointers of the two float4 it fetches per operand and k-tile, advanced by BK each tile
    const float* pa[2];
    const float* pb[2];
    int64_t a_step, b_step;
    if (A_KFAST) {          // row = t / 4 + 64 i, k = 4 (t % 4) .. + 3
        pa[0] = A + (m0 + (t >> 2)) * d.sa_m + 4 * (t & 3);
        pa[1] = pa[0] + 64 * d.sa_m;
        a_step = BK;
    } else {                // k = t / 32 + 8 i, m = 4 (t % 32) .. + 3
        pa[0] = A + (int64_t)(t >> 5) * d.sa_k + m0 + 4 * (t & 31);
        pa[1] = pa[0] + 8 * d.sa_k;
        a_step = BK * d.sa_k;
    }
    if (B_KFAST) {
        pb[0] = B + (n0 + (t >> 2)) * d.sb_n + 4 * (t & 3);
        pb[1] = pb[0] + 64 * d.sb_n;
        b_step = BK;
    } else {
        pb[0] = B + (int64_t)(t >> 5) * d.sb_k + n0 + 4 * (t & 31);
        pb[1] = pb[0] + 8 * d.sb_k;
        b_step = BK * d.sb_k;
    }
    float4 ra[2], rb[2];
    auto load_tile = [&]() {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ra[i] = __ldg(reinterpret_cast<const float4*>(pa[i]));
            rb[i] = __ldg(reinterpret_cast<const float4*>(pb[i]));
            pa[i] += a_step;
            pb[i] += b_step;
        }
    };
    auto store_one = [&](float (*S)[BM], bool kfast, int i, const float4& v) {
        if (kfast) {
            const int row = (t >> 2) + 64 * i, kq = t & 3;
            const int col = (((row >> 2) ^ (2 * kq)) << 2) | (row & 3);
            S[4 * kq + 0][col] = v.x;
            S[4 * kq + 1][col] = v.y;
            S[4 * kq + 2][col] = v.z;
            S[4 * kq + 3][col] = v.w;
        } else {
            const int k = (t >> 5) + 8 * i, m4 = t & 31;
            *reinterpret_cast<float4*>(&S[k][(m4 ^ (2 * (k >> 2))) << 2]) = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            store_one(As[buf], A_KFAST, i, ra[i]);
            store_one(Bs[buf], B_KFAST, i, rb[i]);
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int64_t ktiles = d.K / BK;
    load_tile();
    store_tile(0);
    __syncthreads();
    for (int64_t kt = 0; kt < ktiles; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < ktiles) load_tile();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const int sw = 2 * (kk >> 2);
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][(ty ^ sw) << 2]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][((16 + ty) ^ sw) << 2]);
            const float4 b0v = *reinterpret_cast<const float4*>(&Bs[cur][kk][(tx ^ sw) << 2]);
            const float4 b1v = *reinterpret_cast<const float4*>(&Bs[cur][kk][((16 + tx) ^ sw) << 2]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            store_tile(cur ^ 1);
            __syncthreads();
        }
    }
    // rows ty*4 .. +3 and 64 + ty*4 .. +3; columns tx*4 .. +3 and 64 + tx*4 .. +3 (C rows are contiguous)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t gn = n0 + 64 * h + tx * 4;
            float4 v = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            if (bias) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + gn));
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            }
            float4* p = reinterpret_cast<float4*>(C + gm * d.sc_m + gn);
            if (accumulate) {
                const float4 o = *p;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *p = v;
        }
    }
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// 0 = launched, -1 = not eligible
int try_fast(const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int accumulate) {
    static const bool off = getenv("LG_SIMT_NO_FAST") != nullptr;
    if (off) return -1;
    const int64_t batch = d->batch0 * d->batch1;
    if (d->M % 128 || d->N % 128 || d->K % 16 || d->K < 16 || batch < 1 || batch > 65535 || d->M / 128 > 65535) return -1;
    const bool a_k = d->sa_k == 1 && d->sa_m % 4 == 0, a_m = d->sa_m == 1 && d->sa_k % 4 == 0;
    const bool b_k = d->sb_k == 1 && d->sb_n % 4 == 0, b_n = d->sb_n == 1 && d->sb_k % 4 == 0;
    if (!(a_k || a_m) || !(b_k || b_n) || d->sc_n != 1 || d->sc_m % 4) return -1;
    if (!al16(a) || !al16(b) || !al16(c) || (bias && !al16(bias))) return -1;
    const int64_t bs[6] = {d->sa_b0, d->sa_b1, d->sb_b0, d->sb_b1, d->sc_b0, d->sc_b1};
    for (int64_t v : bs)
        if (v % 4) return -1;
    dim3 grid((unsigned)(d->N / 128), (unsigned)(d->M / 128), (unsigned)batch);
    const float *A = (const float*)a, *B = (const float*)b, *bi = (const float*)bias;
    float* C = (float*)c;
    if (a_k && b_k) gemm_simt_fast_kernel<true, true><<<grid, 256, 0, stream()>>>(A, B, C, bi, *d, accumulate);
    else if (a_k) gemm_simt_fast_kernel<true, false><<<grid, 256, 0, stream()>>>(A, B, C, bi, *d, accumulate);
    else if (b_k) gemm_simt_fast_kernel<false, true><<<grid, 256, 0, stream()>>>(A, B, C, bi, *d, accumulate);
    else gemm_simt_fast_kernel<false, false><<<grid, 256, 0, stream()>>>(A, B, C, bi, *d, accumulate);
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
        if (try_fast(d, a, b, c, bias, accumulate) == 0) {
            LG_CHECK_LAUNCH();
            return 0;
        }
        return launch<float, 128, 128, 16, 8, 8>(d, a, b, c, bias, accumulate);
    }
    if (dtype == LG_F64) return launch<double, 64, 64, 16, 4, 4>(d, a, b, c, bias, accumulate);
    return set_error("lg_gemm: unsupported dtype %d", dtype);
}
}  // namespace lg
