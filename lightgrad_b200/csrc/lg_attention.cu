// Fused multi-head self-attention for sm_100a: one CTA per (batch, head) tile, seq = 128, head_dim = 64.
//
// Replaces, for BertSelfAttention of the reference (examples/bert.py:68-88: scores = Q K^T / sqrt(d); softmax;
// context = P V), the six batched GEMM launches per layer of the composed path and their (b, h, s, s) round trips
// through HBM (scores / probabilities 25 MB per layer at batch 32).  Everything between the projections stays on
// chip:
//
//   forward   S = Q K^T (tcgen05 kind::tf32, TMEM) -> row softmax in registers -> P as a K-major smem operand
//             (over the Q / K tiles, which are dead by then) -> O = P V (TMEM) -> O / rowsum, LSE saved
//   backward  S = Q K^T and dP = dO V^T recomputed into TMEM from one residency of Q, K, V, dO;
//             P = exp(alpha S - LSE), dS = alpha P (dP - D), D = rowsum(dO o O);
//             dV = P^T dO, dQ = dS K, dK = dS^T Q -- P^T / dS / dS^T are written to shared memory by the threads that
//             own the rows (the transposes cost one conflict-free 4-byte store per element), the MN-major B operands
//             (dO, K, Q with the head dimension contiguous) are fetched by TMA in the 32-byte-base swizzle the tf32
//             tensor core needs.  The probabilities never exist in HBM.
//
// Operands are the fp32 tensors the projections wrote (the stacked (3, rows, H) Q / K / V buffer, per-head tiles
// addressed through TMA coordinates); products are tf32 x tf32 -> fp32, exactly like the GEMMs they replace.
// Roofline: HBM.  Algorithmic bytes per tile: forward 3 x 32 KB read + 32 KB written; backward 5 x 32 KB read
// (Q, K, V, dO, O) + 3 x 32 KB written.
#include "lg_tc.cuh"

using namespace lg;
using namespace lg::tc;

namespace {

constexpr int SEQ = 128, HD = 64;
constexpr int TILE_BYTES = SEQ * HD * 4;         // 32 KB: one (128 x 64) fp32 operand tile
constexpr int KB_BYTES = SEQ * 128;              // 16 KB: 128 rows x one 128-byte swizzle row (32 fp32 of K)
constexpr int ATT_THREADS = 320;                 // warp 0: TMA + MMA issue (one lane); warp 1: TMEM; warps 2..9: rows
constexpr int ROW_THREADS = 256;                 // two warps per TMEM lane quarter, each takes half of the columns: the
                                                 // row work is a chain of TMEM loads, ex2 and shared-memory stores whose
                                                 // latency one warp per scheduler cannot hide

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void row_threads_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory"); }

// A (128 x 128 of K, four 16 KB k-blocks, K-major SWIZZLE_128B) times a K-major B tile of N rows
template <int N, int KBLOCKS>
__device__ __forceinline__ void mma_kmajor(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr) {
    constexpr uint32_t idesc = umma_idesc<4>(128, N, false, false);
#pragma unroll
    for (int kb = 0; kb < KBLOCKS; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_issue<4, 1>(d_tmem, umma_desc(a_addr + kb * KB_BYTES + ks * 32, 16, 1024, 2),
                             umma_desc(b_addr + kb * (N * 128) + ks * 32, 16, 1024, 2), idesc, (kb | ks) ? 1u : 0u);
}
// A (128 x 128 of K, K-major) times an MN-major B tile (K = 128 rows of 64 contiguous values: two 32-value chunks,
// each 128 rows x 128 bytes in the 32-byte-base swizzle; a k-step is 8 rows = 1024 bytes)
__device__ __forceinline__ void mma_mn_b(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr) {
    constexpr uint32_t idesc = umma_idesc<4>(128, HD, false, true);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
            umma_issue<4, 1>(d_tmem, umma_desc(a_addr + kb * KB_BYTES + ks * 32, 16, 1024, 2),
                             umma_desc(b_addr + (kb * 4 + ks) * 1024, SEQ * 128, 512, 1), idesc, (kb | ks) ? 1u : 0u);
}

// row `row` of a K-major (128 rows x 32 fp32 per k-block) operand: this thread's 32 values of k-block `kb`
__device__ __forceinline__ void store_row_chunk(uint8_t* base, int kb, int row, const float (&x)[32]) {
    uint8_t* p = base + kb * KB_BYTES + row * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(p + ((j ^ (row & 7)) << 4)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
}
// the TRANSPOSE: this thread owns source row `row` (= column `row` of the operand); x[k] goes to operand row r0 + k.
// The 32 lanes of a warp own 32 consecutive source rows, i.e. 128 contiguous bytes of each operand row: no conflicts.
__device__ __forceinline__ void store_col_chunk(uint8_t* base, int row, int r0, const float (&x)[32]) {
    uint8_t* p = base + (row >> 5) * KB_BYTES + (row & 3) * 4;
    const int c = (row & 31) >> 2;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const int r = r0 + k;
        *reinterpret_cast<float*>(p + r * 128 + ((c ^ (r & 7)) << 4)) = x[k];
    }
}

struct AttnParams {
    int batch, heads, rows;          // rows = batch * SEQ
    float scale_log2e;               // alpha * log2(e)
    float scale;
    float* out;                      // forward: (rows, H) result
    float* lse;                      // (batch * heads * SEQ): alpha * rowmax + ln(rowsum) (natural log)
    const float* o;                  // backward: forward result
    const float* dout;
    float* dqkv;                     // backward: (3, rows, H)
    float* dbias[3];                 // backward: d(bias) of the Q / K / V projections (H each), accumulated; or nullptr
};

// column sums of a 32 x 32 chunk held one row per lane (see lg_gemm_tc.cu): lane j ends with the sum of column j
__device__ __forceinline__ float warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float give = upper ? x[i] : x[i + s];
            const float keep = upper ? x[i + s] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, give, s);
        }
    }
    return x[0];
}
// d(bias)[col0 + lane] += column sum of this warp's 32 rows of a gradient chunk
__device__ __forceinline__ void bias_grad_chunk(float* dbias, const uint32_t (&v)[32], int col0, int lane) {
    if (dbias == nullptr) return;
    float x[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
    const float sum = warp_transpose_sum(x, lane);
    atomicAdd(dbias + col0 + lane, sum);
}

struct AttnMaps {
    CUtensorMap qkv_k;    // stacked (3 * rows, H) Q/K/V buffer, box 32 x 128, SWIZZLE_128B       (K-major operand tiles)
    CUtensorMap qkv_mn;   // same memory, SWIZZLE_128B_ATOM_32B                                     (MN-major operand tiles)
    CUtensorMap do_k;     // (rows, H) gradient of the result, SWIZZLE_128B
    CUtensorMap do_mn;    // same, SWIZZLE_128B_ATOM_32B
    CUtensorMap dqkv;     // backward result (3 * rows, H), box 32 x 128, SWIZZLE_128B: written by TMA from staged row tiles
};

// ======================================= forward =========================================================
// One persistent CTA per SM with TWO tiles in flight: the eight row warps form two groups of four, group g owns the
// CTA's tiles 2n + g together with operand set g (Q | K | V, 96 KB) and its own S / O columns of tensor memory.  For one
// tile the work is a serial chain -- loads, S = Q K^T, softmax, O = P V, store: ~5.5 us, most of it latencies of barrier
// hand-offs, the tensor pipe and the loads (ncu: the row warps spent two thirds of their samples waiting for S or O) --
// so while one group computes its softmax the other tile's products and loads run.  A set can only be refilled when its
// P V product has finished (P lives over Q | K), hence no deeper pipeline at 4-byte operands.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    LG_PDL_TRIGGER();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // set g at smem + g * 96 KB: Q (32 KB; with K: the 64 KB probability operand once S is complete) | K | V (MN-major)
    uint64_t* bars = (uint64_t*)(smem + 6 * TILE_BYTES);
    uint64_t* b_qk = bars;        // [2] Q, K of a set landed
    uint64_t* b_v = bars + 2;     // [2] V of a set landed
    uint64_t* b_s = bars + 4;     // [2] S = Q K^T complete
    uint64_t* b_p = bars + 6;     // [2] probabilities staged (128 arrivals)
    uint64_t* b_o = bars + 8;     // [2] O = P V complete: the set may be refilled
    uint32_t* tmem_slot = (uint32_t*)(bars + 10);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.heads * HD;
    const int tiles = p.batch * p.heads;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; the i-th of them belongs to group / set i & 1
    const int my_tiles = (int)blockIdx.x < tiles ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_qk + i, 1); mbar_init(b_v + i, 1); mbar_init(b_s + i, 1); mbar_init(b_p + i, 128);
            mbar_init(b_o + i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.qkv_k));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.qkv_mn));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;          // group g: S at columns 256 g .. +127, O at 256 g + 128 .. +63
    LG_PDL_WAIT();

    if (warp == 0) {
        // ---- TMA producer: tile i goes to set i & 1 as soon as that set's previous P V product is done
        if (lane == 0) {
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                const int set = i & 1, n = i >> 1;
                if (n > 0) mbar_wait(b_o + set, (uint32_t)(n - 1) & 1);
                const int b = tile / p.heads, h = tile - b * p.heads;
                const int r0 = b * SEQ, c0 = h * HD;
                uint8_t* sQ = smem + set * 3 * TILE_BYTES;
                uint8_t* sK = sQ + TILE_BYTES;
                uint8_t* sV = sQ + 2 * TILE_BYTES;
                mbar_expect_tx(b_qk + set, 2 * TILE_BYTES);
                tma_load_2d(&maps.qkv_k, b_qk + set, sQ, c0, r0);
                tma_load_2d(&maps.qkv_k, b_qk + set, sQ + KB_BYTES, c0 + 32, r0);
                tma_load_2d(&maps.qkv_k, b_qk + set, sK, c0, p.rows + r0);
                tma_load_2d(&maps.qkv_k, b_qk + set, sK + KB_BYTES, c0 + 32, p.rows + r0);
                mbar_expect_tx(b_v + set, TILE_BYTES);
                tma_load_2d(&maps.qkv_mn, b_v + set, sV, c0, 2 * p.rows + r0);
                tma_load_2d(&maps.qkv_mn, b_v + set, sV + KB_BYTES, c0 + 32, 2 * p.rows + r0);
            }
        }
    } else if (warp == 1) {
        // ---- tensor-core issue: whichever of {S of the next tile, P V of the oldest open tile} has its operands first
        if (lane == 0) {
            int s_next = 0, pv_next = 0;
            while (pv_next < my_tiles) {
                bool progressed = false;
                // S(i) overwrites the S columns of tile i - 2: its softmax has read them once P(i - 2) was staged, i.e.
                // once P V (i - 2) has been issued
                if (s_next < my_tiles && s_next - pv_next < 2) {
                    const int set = s_next & 1;
                    if (mbar_test(b_qk + set, (uint32_t)(s_next >> 1) & 1)) {
                        tc_fence_after();
                        uint8_t* sQ = smem + set * 3 * TILE_BYTES;
                        mma_kmajor<SEQ, 2>(tmem + 256 * set, smem_u32(sQ), smem_u32(sQ + TILE_BYTES));
                        umma_commit(b_s + set);
                        ++s_next;
                        progressed = true;
                    }
                }
                if (pv_next < s_next) {
                    const int set = pv_next & 1;
                    const uint32_t ph = (uint32_t)(pv_next >> 1) & 1;
                    if (mbar_test(b_p + set, ph) && mbar_test(b_v + set, ph)) {
                        tc_fence_after();
                        uint8_t* sQ = smem + set * 3 * TILE_BYTES;
                        mma_mn_b(tmem + 256 * set + 128, smem_u32(sQ), smem_u32(sQ + 2 * TILE_BYTES));
                        umma_commit(b_o + set);
                        ++pv_next;
                        progressed = true;
                    }
                }
                if (!progressed) __nanosleep(32);
            }
        }
    } else {
        // ---- row warps: group g = tiles g, g + 2, ... of this CTA; a thread owns one row of the tile
        const int q = warp & 3;                      // TMEM lane quarter of this warp
        const int g = (warp - 2) >> 2;
        const int row = 32 * q + lane;
        const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16) + 256 * g;
        uint8_t* sP = smem + g * 3 * TILE_BYTES;
        for (int i = g; i < my_tiles; i += 2) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const uint32_t ph = (uint32_t)(i >> 1) & 1;
            const int b = tile / p.heads, h = tile - b * p.heads;
            mbar_wait(b_s + g, ph);
            tc_fence_after();
            // the whole row of S stays in registers between the maximum and the exponentials
            uint32_t v0[32], v1[32], v2[32], v3[32];
            tmem_ld32(v0, t_row);
            tmem_ld32(v1, t_row + 32);
            tmem_ld32(v2, t_row + 64);
            tmem_ld32(v3, t_row + 96);
            tmem_wait_ld();
            float m = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                m = fmaxf(m, fmaxf(fmaxf(__uint_as_float(v0[k]), __uint_as_float(v1[k])),
                                   fmaxf(__uint_as_float(v2[k]), __uint_as_float(v3[k]))));
            // exp(alpha (s - max)) as ex2 of a single fma; the unnormalised values are the A operand of P V
            const float mneg = -m * p.scale_log2e;
            float sum = 0.f;
            {
                float x[32];
#define LG_ATT_EXP_CHUNK(V_, C_)                                                       \
    _Pragma("unroll") for (int k = 0; k < 32; ++k) {                                    \
        x[k] = ex2(fmaf(__uint_as_float(V_[k]), p.scale_log2e, mneg));                 \
        sum += x[k];                                                                   \
    }                                                                                  \
    store_row_chunk(sP, C_, row, x);      /* Q and K tiles of this set are dead: S is complete */
                LG_ATT_EXP_CHUNK(v0, 0)
                LG_ATT_EXP_CHUNK(v1, 1)
                LG_ATT_EXP_CHUNK(v2, 2)
                LG_ATT_EXP_CHUNK(v3, 3)
#undef LG_ATT_EXP_CHUNK
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(b_p + g);
            p.lse[(size_t)tile * SEQ + row] = m * p.scale + __logf(sum);
            const float inv = 1.0f / sum;
            mbar_wait(b_o + g, ph);
            tc_fence_after();
            float* orow = p.out + (size_t)(b * SEQ + row) * H + h * HD;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(v, t_row + 128 + c * 32);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<float4*>(orow + c * 32)[j] =
                        make_float4(__uint_as_float(v[4 * j]) * inv, __uint_as_float(v[4 * j + 1]) * inv,
                                    __uint_as_float(v[4 * j + 2]) * inv, __uint_as_float(v[4 * j + 3]) * inv);
            }
            tc_fence_before();       // this group's next P V product overwrites these columns only after its next b_p
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// ======================================= backward ========================================================
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
    LG_PDL_TRIGGER();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sX = smem;                        // 64 KB: Q_k | K_k, then P^T, then dS^T
    uint8_t* sY = smem + 2 * TILE_BYTES;       // 64 KB: dO_k | V_k, then dS
    uint8_t* sZ1 = smem + 4 * TILE_BYTES;      // 32 KB: dO (MN-major)
    uint8_t* sZ2 = smem + 5 * TILE_BYTES;      // 32 KB: K (MN-major)
    uint8_t* sZ3 = smem + 6 * TILE_BYTES;      // 32 KB: Q (MN-major)
    uint64_t* bars = (uint64_t*)(smem + 7 * TILE_BYTES);
    uint64_t* b_ld1 = bars;         // Q_k, K_k, dO_k, V_k landed
    uint64_t* b_ld2 = bars + 1;     // dO_mn, K_mn, Q_mn landed
    uint64_t* b_sdp = bars + 2;     // S and dP complete
    uint64_t* b_op1 = bars + 3;     // P^T and dS staged (128 arrivals)
    uint64_t* b_dv = bars + 4;      // dV complete (P^T consumed: the row threads may overwrite it with dS^T)
    uint64_t* b_dq = bars + 5;      // dQ complete
    uint64_t* b_op2 = bars + 6;     // dS^T staged (128 arrivals)
    uint64_t* b_dk = bars + 7;      // dK complete
    uint64_t* b_out = bars + 8;     // the result tiles staged over dO / K / Q (MN-major) have left shared memory
    uint32_t* tmem_slot = (uint32_t*)(bars + 9);
    float* xch = (float*)(bars + 10);  // [128]: rowsum(dO o O)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.heads * HD;
    const int tiles = p.batch * p.heads;

    if (warp == 0 && lane == 0) {
        mbar_init(b_ld1, 1); mbar_init(b_ld2, 1); mbar_init(b_sdp, 1); mbar_init(b_op1, ROW_THREADS); mbar_init(b_dv, 1); mbar_init(b_dq, 1);
        mbar_init(b_op2, ROW_THREADS); mbar_init(b_dk, 1); mbar_init(b_out, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.qkv_k));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.qkv_mn));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.do_k));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.do_mn));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.dqkv));
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;   // S 0..127 | dP 128..255 | dV 256..319 | dQ 320..383 | dK 384..447
    LG_PDL_WAIT();

    if (warp == 0) {
        if (lane == 0) {
            // the K-major tiles (needed first) are requested as soon as this tile's last product has consumed the shared
            // memory; the MN-major ones once the results staged over them have been written out
            auto load_k = [&](int tile) {
                const int b = tile / p.heads, h = tile - b * p.heads;
                const int r0 = b * SEQ, c0 = h * HD;
                mbar_expect_tx(b_ld1, 4 * TILE_BYTES);
                tma_load_2d(&maps.qkv_k, b_ld1, sX, c0, r0);                                   // Q
                tma_load_2d(&maps.qkv_k, b_ld1, sX + KB_BYTES, c0 + 32, r0);
                tma_load_2d(&maps.qkv_k, b_ld1, sX + TILE_BYTES, c0, p.rows + r0);             // K
                tma_load_2d(&maps.qkv_k, b_ld1, sX + TILE_BYTES + KB_BYTES, c0 + 32, p.rows + r0);
                tma_load_2d(&maps.do_k, b_ld1, sY, c0, r0);                                    // dO
                tma_load_2d(&maps.do_k, b_ld1, sY + KB_BYTES, c0 + 32, r0);
                tma_load_2d(&maps.qkv_k, b_ld1, sY + TILE_BYTES, c0, 2 * p.rows + r0);         // V
                tma_load_2d(&maps.qkv_k, b_ld1, sY + TILE_BYTES + KB_BYTES, c0 + 32, 2 * p.rows + r0);
            };
            auto load_mn = [&](int tile) {
                const int b = tile / p.heads, h = tile - b * p.heads;
                const int r0 = b * SEQ, c0 = h * HD;
                mbar_expect_tx(b_ld2, 3 * TILE_BYTES);
                tma_load_2d(&maps.do_mn, b_ld2, sZ1, c0, r0);                                  // dO, MN-major
                tma_load_2d(&maps.do_mn, b_ld2, sZ1 + KB_BYTES, c0 + 32, r0);
                tma_load_2d(&maps.qkv_mn, b_ld2, sZ2, c0, p.rows + r0);                        // K, MN-major
                tma_load_2d(&maps.qkv_mn, b_ld2, sZ2 + KB_BYTES, c0 + 32, p.rows + r0);
                tma_load_2d(&maps.qkv_mn, b_ld2, sZ3, c0, r0);                                 // Q, MN-major
                tma_load_2d(&maps.qkv_mn, b_ld2, sZ3 + KB_BYTES, c0 + 32, r0);
            };
            auto load_tile = [&](int tile) { load_k(tile); load_mn(tile); };
            if ((int)blockIdx.x < tiles) load_tile(blockIdx.x);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(b_ld1, ph);
                tc_fence_after();
                mma_kmajor<SEQ, 2>(tmem, smem_u32(sX), smem_u32(sX + TILE_BYTES));             // S  = Q K^T
                mma_kmajor<SEQ, 2>(tmem + 128, smem_u32(sY), smem_u32(sY + TILE_BYTES));       // dP = dO V^T
                umma_commit(b_sdp);
                mbar_wait(b_op1, ph);
                mbar_wait(b_ld2, ph);
                tc_fence_after();
                mma_mn_b(tmem + 256, smem_u32(sX), smem_u32(sZ1));                             // dV = P^T dO
                umma_commit(b_dv);
                mma_mn_b(tmem + 320, smem_u32(sY), smem_u32(sZ2));                             // dQ = dS K
                umma_commit(b_dq);
                mbar_wait(b_op2, ph);
                tc_fence_after();
                mma_mn_b(tmem + 384, smem_u32(sX), smem_u32(sZ3));                             // dK = dS^T Q
                umma_commit(b_dk);
                if (tile + (int)gridDim.x < tiles) {
                    mbar_wait(b_dk, ph);             // every operand of this tile has been consumed
                    load_k(tile + gridDim.x);
                    mbar_wait(b_out, ph);            // dV / dQ / dK tiles have been read out of shared memory
                    load_mn(tile + gridDim.x);
                }
            }
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // which half of the columns (the other warp of the quarter: the rest)
        const int row = 32 * q + lane;
        const uint32_t t_row = tmem + ((uint32_t)(32 * q) << 16);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            const int b = tile / p.heads, h = tile - b * p.heads;
            // D = rowsum(dO o O), fetched while the loads and the first products are in flight.  Coalesced: this warp takes
            // 16 of its quarter's rows, two 256-byte rows per instruction (16 lanes each), and leaves the sums in shared
            // memory for the threads that own the rows
            {
                const int sub = lane >> 4, chunk = lane & 15;
                const size_t g0 = (size_t)(b * SEQ + 32 * q + 16 * half + sub) * H + h * HD + 4 * chunk;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(p.o + g0 + (size_t)(2 * i) * H));
                    const float4 g = __ldg(reinterpret_cast<const float4*>(p.dout + g0 + (size_t)(2 * i) * H));
                    float d = a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (chunk == 0) xch[32 * q + 16 * half + 2 * i + sub] = d;
                }
            }
            const float lneg = -p.lse[(size_t)tile * SEQ + row] * 1.4426950408889634f;
            row_threads_sync();
            const float dsum = xch[row];
            mbar_wait(b_sdp, ph);
            tc_fence_after();
            // pass A: P^T -> X (transposed), dS -> Y (row form); both chunks' TMEM loads in flight together.  dS stays in
            // registers: its transpose goes to X once dV has consumed P^T (no second TMEM pass, no second exponential)
            float ds0[32], ds1[32];
            {
                uint32_t s0[32], g0[32], s1[32], g1[32];
                float pv[32];
                tmem_ld32(s0, t_row + half * 64);
                tmem_ld32(g0, t_row + 128 + half * 64);
                tmem_ld32(s1, t_row + half * 64 + 32);
                tmem_ld32(g1, t_row + 128 + half * 64 + 32);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    pv[k] = ex2(fmaf(__uint_as_float(s0[k]), p.scale_log2e, lneg));
                    ds0[k] = p.scale * pv[k] * (__uint_as_float(g0[k]) - dsum);
                }
                store_col_chunk(sX, row, half * 64, pv);
                store_row_chunk(sY, 2 * half, row, ds0);
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    pv[k] = ex2(fmaf(__uint_as_float(s1[k]), p.scale_log2e, lneg));
                    ds1[k] = p.scale * pv[k] * (__uint_as_float(g1[k]) - dsum);
                }
                store_col_chunk(sX, row, half * 64 + 32, pv);
                store_row_chunk(sY, 2 * half + 1, row, ds1);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(b_op1);
            mbar_wait(b_dv, ph);
            tc_fence_after();
            // pass B: dS^T -> X
            store_col_chunk(sX, row, half * 64, ds0);
            store_col_chunk(sX, row, half * 64 + 32, ds1);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(b_op2);
            // dV (rows = keys) and dQ (rows = queries) while dK is being formed.  The rows are staged in shared memory in
            // the TMA box layout -- over dO / K / Q (MN-major), which dV / dQ / dK have consumed -- and written by TMA:
            // a thread owns one 128-byte piece of a row, and 32 such pieces per store instruction, each in its own
            // line, kept the load-store unit busy for longer than the rest of the tile took.
            // (the projections' bias gradients are the column sums of dQ / dK / dV: taken from the registers that hold the
            //  rows, not by three kernels that read the matrices back)
            const int r0 = b * SEQ, c0 = h * HD;
            {
                uint32_t v[32];
                float x[32];
                tmem_ld32(v, t_row + 256 + half * 32);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
                store_row_chunk(sZ1, half, row, x);
                bias_grad_chunk(p.dbias[2], v, c0 + half * 32, lane);
                mbar_wait(b_dq, ph);
                tc_fence_after();
                tmem_ld32(v, t_row + 320 + half * 32);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
                store_row_chunk(sZ2, half, row, x);
                bias_grad_chunk(p.dbias[0], v, c0 + half * 32, lane);
            }
            fence_async_smem();
            row_threads_sync();
            if (threadIdx.x == 64) {
                tma_store_2d(&maps.dqkv, sZ1, c0, 2 * p.rows + r0);
                tma_store_2d(&maps.dqkv, sZ1 + KB_BYTES, c0 + 32, 2 * p.rows + r0);
                tma_store_2d(&maps.dqkv, sZ2, c0, r0);
                tma_store_2d(&maps.dqkv, sZ2 + KB_BYTES, c0 + 32, r0);
                tma_store_commit();
            }
            mbar_wait(b_dk, ph);
            tc_fence_after();
            {
                uint32_t v[32];
                float x[32];
                tmem_ld32(v, t_row + 384 + half * 32);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
                store_row_chunk(sZ3, half, row, x);
                bias_grad_chunk(p.dbias[1], v, c0 + half * 32, lane);
            }
            fence_async_smem();
            row_threads_sync();
            if (threadIdx.x == 64) {
                tma_store_2d(&maps.dqkv, sZ3, c0, p.rows + r0);
                tma_store_2d(&maps.dqkv, sZ3 + KB_BYTES, c0 + 32, p.rows + r0);
                tma_store_commit();
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(b_out);
            }
            tc_fence_before();       // (the next tile's dV / dQ / dK products need this warp's next arrivals first)
        }
    }
    if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------
int make_map_2d(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, CUtensorMapSwizzle swz) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)SEQ};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r);
    return 0;
}

constexpr size_t FWD_SMEM = 6 * TILE_BYTES + 128 + 2048 + 1024;     // operands | barriers | row exchange | alignment
constexpr size_t BWD_SMEM = 7 * TILE_BYTES + 128 + 1024 + 1024;

int check_shape(const char* who, int dtype, int64_t batch, int64_t seq, int64_t heads, int64_t head_dim) {
    LG_REQUIRE(dtype == LG_F32 && seq == SEQ && head_dim == HD && batch >= 1 && heads >= 1 &&
                   batch * heads < (1 << 30) && batch * seq < (1ll << 30),
               "%s: the fused kernel takes float32, seq = %d, head_dim = %d (ask lg_attention_supported first)", who, SEQ, HD);
    return 0;
}

}  // namespace

extern "C" {

int lg_attention_supported(int dtype, int64_t seq, int64_t head_dim) {
    static const bool off = getenv("LG_NO_FUSED_ATTENTION") != nullptr;
    return (!off && dtype == LG_F32 && seq == SEQ && head_dim == HD) ? 1 : 0;
}

int lg_attention_fwd(int dtype, const void* qkv, int64_t batch, int64_t seq, int64_t heads, int64_t head_dim,
                     double scale, void* out, void* lse) {
    LG_INIT();
    if (check_shape("lg_attention_fwd", dtype, batch, seq, heads, head_dim)) return 1;
    LG_REQUIRE((((uintptr_t)qkv | (uintptr_t)out) & 15) == 0, "lg_attention_fwd: buffers must be 16-byte aligned");
    if (load_encode()) return 1;
    const int64_t rows = batch * seq, H = heads * head_dim;
    AttnMaps maps;
    if (make_map_2d(&maps.qkv_k, qkv, H, 3 * rows, H, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_2d(&maps.qkv_mn, qkv, H, 3 * rows, H, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    maps.do_k = maps.qkv_k;
    maps.do_mn = maps.qkv_mn;
    maps.dqkv = maps.qkv_k;
    AttnParams p = {};
    p.batch = (int)batch;
    p.heads = (int)heads;
    p.rows = (int)rows;
    p.scale = (float)scale;
    p.scale_log2e = (float)(scale * 1.4426950408889634);
    p.out = (float*)out;
    p.lse = (float*)lse;
    static bool attr_done = false;
    if (!attr_done) {
        LG_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
        attr_done = true;
    }
    const int tiles = (int)(batch * heads);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    LG_CUDA(launch_pdl(attention_fwd_kernel, dim3(grid), dim3(ATT_THREADS), FWD_SMEM, stream(), maps, p));
    LG_CHECK_LAUNCH();
    return 0;
}

int lg_attention_bwd(int dtype, const void* qkv, const void* out, const void* dout, const void* lse, int64_t batch,
                     int64_t seq, int64_t heads, int64_t head_dim, double scale, void* dqkv, void* dbq, void* dbk,
                     void* dbv) {
    LG_INIT();
    if (check_shape("lg_attention_bwd", dtype, batch, seq, heads, head_dim)) return 1;
    LG_REQUIRE((((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dqkv) & 15) == 0,
               "lg_attention_bwd: buffers must be 16-byte aligned");
    if (load_encode()) return 1;
    const int64_t rows = batch * seq, H = heads * head_dim;
    AttnMaps maps;
    if (make_map_2d(&maps.qkv_k, qkv, H, 3 * rows, H, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_2d(&maps.qkv_mn, qkv, H, 3 * rows, H, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    if (make_map_2d(&maps.do_k, dout, H, rows, H, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map_2d(&maps.do_mn, dout, H, rows, H, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 1;
    if (make_map_2d(&maps.dqkv, dqkv, H, 3 * rows, H, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    AttnParams p = {};
    p.batch = (int)batch;
    p.heads = (int)heads;
    p.rows = (int)rows;
    p.scale = (float)scale;
    p.scale_log2e = (float)(scale * 1.4426950408889634);
    p.lse = (float*)lse;
    p.o = (const float*)out;
    p.dout = (const float*)dout;
    p.dqkv = (float*)dqkv;
    p.dbias[0] = (float*)dbq;
    p.dbias[1] = (float*)dbk;
    p.dbias[2] = (float*)dbv;
    static bool attr_done = false;
    if (!attr_done) {
        LG_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        attr_done = true;
    }
    const int tiles = (int)(batch * heads);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    LG_CUDA(launch_pdl(attention_bwd_kernel, dim3(grid), dim3(ATT_THREADS), BWD_SMEM, stream(), maps, p));
    LG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
