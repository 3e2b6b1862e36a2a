// Tensor-core matmul for sm_100a: tcgen05.mma (kind::tf32 on fp32 operands, kind::f16 on bf16 operands) with
// TMEM accumulators, operands staged by TMA, persistent warp-specialised CTAs.
//
// Replaces the reference's OpenCL SGEMM (opencl/kernels.py:201-337) for the "TF32 tensor-core mode".
// fp32 operands are read straight from HBM by TMA (128-byte swizzled boxes) and consumed by
// tcgen05.mma.kind::tf32 (the tensor core drops the low mantissa bits; accumulation is fp32 in
// TMEM) -- there is no conversion pass and no operand copy: transposed operands (x @ W^T,
// dC @ B^T, A^T @ dC) are expressed through the UMMA "major" bits, row strides through the TMA
// descriptor, ragged edges through TMA out-of-bounds zero fill / clipped stores.
//
//   CTA tile 128 x BN (BN in {64,128,192,256}), K step 32 fp32 = one 128-byte swizzle row
//   warp 0        TMA producer          (ring of kStages {A,B} stages, full/empty mbarriers)
//   warp 1        TMEM allocator + MMA issuer (one elected lane; 4 x tcgen05.mma m128nBNk8 per stage;
//                                       tcgen05.commit releases the stage / publishes the accumulator)
//   warps 2..9    epilogue (two per TMEM lane quarter, alternate 32-column chunks): tcgen05.ld 32x32b.x32 -> (+bias) -> swizzled smem -> TMA store
//                 (cp.reduce.async.bulk.tensor .add when the K range is split across CTAs)
//   two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
// Most launches of a training step run the CTA-PAIR variant further down (cta_group::2, 256 x BN tile over two SMs, each
// staging half of B): at 4-byte operands it is the L2 -> SM operand traffic, not the tensor pipe, that bounds a CTA.
// BF16 mode (LG_GEMM_BF16_TC): the same kernels instantiated with 2-byte operands (ES = 2).  The byte geometry
// of a stage is identical -- a k-block is still one 128-byte swizzle row (64 bf16 instead of 32 fp32), a UMMA
// k-step still 32 bytes -- only the MN-major layout differs: 16-bit operands use the plain SWIZZLE_128B atoms
// (64-element chunks, 8-row groups) where tf32 needs the 32-byte-base variant.  Operands are bf16 copies staged
// by the host layer (lg_cast / fused producer epilogues); accumulation and the result stay fp32.
// Roofline: tensor pipe (TF32 dense = half the bf16 rate).  Algorithmic flops 2*M*N*K.
#include "lg_tc.cuh"

using namespace lg;
using namespace lg::tc;

namespace {

constexpr int LG_MAX_GROUPS = 4;
constexpr int BM = 128;
// ES = operand element size in bytes (4: fp32 read as tf32, 2: bf16).  Per-ES geometry of one k-block:
template <int ES> struct Geo {
    static constexpr int BKE = 128 / ES;            // elements per k-block = one 128-byte swizzle row
    static constexpr int CH = 128 / ES;             // MN-major: elements per contiguous chunk (one 128-byte row)
    static constexpr int CH_SHIFT = ES == 4 ? 5 : 6;
    static constexpr int CHUNK_BYTES = BKE * 128;   // MN-major: a chunk is BKE k-rows of 128 bytes
    static constexpr int KSTEP_MN_BYTES = (32 / ES) * 128;   // MN-major: k-rows of one UMMA k-step (8 / 16) x 128 B
    static constexpr int MN_SBO = ES == 4 ? 512 : 1024;      // 4-row (base32b) / 8-row groups
    static constexpr int MN_LAYOUT = ES == 4 ? 1 : 2;        // SWIZZLE_128B_BASE32B / SWIZZLE_128B
};
constexpr int KSTEPS = 4;             // UMMA k-steps (32 bytes of K each) per k-block
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int EPI_WARPS = 8;             // two warps per TMEM lane quarter, taking alternate 32-column chunks
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int SCHED_SLOTS = 4;           // work items in flight between the producer thread and the other roles
constexpr int EPI_BUF_BYTES = 32 * 128;   // 32 rows x 32 fp32, per warp, double buffered

// SMs the persistent GEMM grids may occupy (lg_gemm_sm_limit): a data-parallel step leaves a few SMs to the
// collective's CTAs, so that a one-CTA-per-SM grid is never forced into a second wave by them
static int g_gemm_sm_limit = 0;
inline int gemm_sms() {
    const int n = lg::sm_count();
    return (g_gemm_sm_limit > 0 && g_gemm_sm_limit < n) ? g_gemm_sm_limit : n;
}

struct TcParams {
    int M, N, K;
    int tiles_m, tiles_n, splits, kblocks_per_split, kblocks_total;
    int batch1, batches;      // batches = batch0 * batch1
    int reduce_out;           // 1: C += tile (TMA reduce-add): split-K partials and/or accumulate mode
    int groups;               // problems of identical shape served by this launch (1-CTA kernel only)
    int a_chunked, b_chunked; // MN-major operand addressed through a rank-5 chunked map (one box per stage)
    int epi_op;               // LgGemmEpilogue: 0 none, 1 aux = gelu(C) as a second result, 2 C = acc * gelu'(aux),
                              // 3 C = softmax_row(alpha * acc), 4 C = alpha * aux * (acc - sum_row(aux * acc))
    const float* aux;         // epi_op 2 / 4: saved matrix of C's shape, row pitch aux_ld elements
    long long aux_ld;
    long long aux_sb0, aux_sb1;   // its batch strides (epi_op 4; same batch dims as C)
    float epi_alpha;
    float* colsum;            // epi_op 2: colsum[n] += sum over rows of C (the bias gradient of the layer that produced
                              // the pre-activations), accumulated by the epilogue; nullptr: not wanted
    // tail-wave split (plain one-problem GEMMs, splits == 1): the last `tiles % clusters` tiles, which would run as
    // a mostly empty extra wave, are each cut into tail_splits K-ranges whose partials meet by reduce-add
    int items_per_batch;      // work items of one problem: tail_first whole tiles + (tiles - tail_first) * tail_splits
    int tail_first, tail_splits, tail_kper;
    int kcat;                 // operand pairs concatenated along K into ONE result: C = sum_g A_g B_g (1-CTA kernel)
    const float* bias[LG_MAX_GROUPS];
    // dynamic tile order (1-CTA kernel): sched[0] counts the work items claimed beyond the first one of every CTA,
    // sched[1] the CTAs that have made their last claim (the last of them zeroes both for the next launch on this
    // stream); nullptr: the static order w = cta, cta + grid, ...
    unsigned int* sched;
};

// work item -> (tile, K-range); `wi` counts inside one problem of one batch
struct TcItem {
    int tile, split, kb0, kb1, partial;   // partial: the result is one of several K-range partials of its tile
};
__device__ __forceinline__ TcItem decode_item(const TcParams& p, int wi) {
    TcItem it;
    int kper;
    if (wi >= p.tail_first && p.tail_splits > 1) {
        const int r = wi - p.tail_first;
        it.tile = p.tail_first + r / p.tail_splits;
        it.split = r - (r / p.tail_splits) * p.tail_splits;
        kper = p.tail_kper;
        it.partial = 1;
    } else {
        it.tile = wi / p.splits;
        it.split = wi - it.tile * p.splits;
        kper = p.kblocks_per_split;
        it.partial = p.splits > 1;
    }
    it.kb0 = it.split * kper;
    it.kb1 = it.kb0 + kper;
    if (it.kb1 > p.kblocks_total) it.kb1 = p.kblocks_total;
    return it;
}

struct TcMaps {
    CUtensorMap a[LG_MAX_GROUPS], b[LG_MAX_GROUPS], c[LG_MAX_GROUPS];
    CUtensorMap aux;          // epi_op 1: second result, same geometry as c[0]
};

// ---------------------------------------------------------------------------------------------------
// Epilogue of one 128 x BN accumulator tile, run by the four epilogue warps (warp q owns TMEM lanes 32q..32q+31,
// i.e. 32 output rows; thread = row).  Per 32-column chunk: tcgen05.ld (the NEXT chunk's load is in flight while
// this one is staged), + bias, 128B-swizzled staging buffer, TMA store / reduce-add.  The bias slice of a chunk
// is fetched one chunk ahead with one coalesced load per warp and broadcast by shuffles, so no global-memory
// latency sits between the TMEM read and the store.
__device__ __forceinline__ float bias_slice(const float* bias, int col, int N) {
    return (bias != nullptr && col < N) ? __ldg(bias + col) : 0.f;
}

// GELU in the epilogue.  With u = sqrt(2/pi) (x + 0.044715 x^3):  0.5 (1 + tanh u) = sigma(2u) = 1 / (1 + exp(-2u)), so
//   gelu(x)  = x s,                                   s = 1 / (1 + 2^(x (k1 + k2 x^2))),  k1 = -2 log2(e) sqrt(2/pi)
//   gelu'(x) = s + x s (1 - s) 2 u',                  2u' = 2 sqrt(2/pi) (1 + 3 * 0.044715 x^2)
// -- one ex2.approx and one rcp.approx per element (abs. error ~2e-7, well inside the tf32 result's own rounding) and
// 7 / 13 instructions instead of the 12 / 20 of the tanh form: the eight epilogue warps share their schedulers with
// the TMA and MMA threads, and the last tile's epilogue of a launch is not hidden behind a main loop.
__device__ __forceinline__ float epi_sigmoid2u(float x, float x2) {
    constexpr float k1 = -2.0f * 1.4426950408889634f * 0.7978845608f, k2 = k1 * 0.044715f;
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * fmaf(k2, x2, k1)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return r;
}
__device__ __forceinline__ float epi_gelu(float x) { return x * epi_sigmoid2u(x, x * x); }
__device__ __forceinline__ float epi_gelu_bwd(float x, float g) {
    constexpr float c1 = 0.7978845608f, c2 = 0.044715f;
    const float x2 = x * x;
    const float s = epi_sigmoid2u(x, x2);
    const float du2 = fmaf(6.0f * c1 * c2, x2, 2.0f * c1);
    return fmaf((x * s) * (1.0f - s), du2, s) * g;
}

struct EpiTile {
    const CUtensorMap* map_c;
    const CUtensorMap* map_aux;   // epi_op 1
    const float* bias;      // nullptr: none (also for split-K partials other than the first)
    const float* aux;       // epi_op 2: row `row0 + lane` of the pre-activation matrix, or nullptr past the last row
    int epi_op;
    int row0, n0, N;        // first output row of this warp, first column of the tile, columns of C
    int bc1, bc0;           // batch coordinates
    int reduce_out;
    bool rows_live;
    float* colsum;          // epi_op 2: column sums of the finished chunk are added here (or nullptr)
};

// Column sums of a 32 x 32 chunk held one row per lane: five exchange steps in which every lane keeps the half of the
// columns selected by one bit of its lane number and hands the other half to its partner -- 31 shuffles instead of
// 32 x 5 -- after which lane j holds the sum of column j over the warp's 32 rows.
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

// stage one 32 x 32 chunk (thread = row) in the 128B-swizzled buffer and hand it to the TMA unit
__device__ __forceinline__ void epi_store(const CUtensorMap* map, const uint32_t (&v)[32], uint8_t* buf, int col0,
                                          int row0, int bc1, int bc0, int reduce_out, int lane) {
    // the TMA store that last read this warp's staging buffer must have drained
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)   // 128-byte swizzle: 16-byte chunk j of row `lane` lives at chunk (j ^ (lane & 7))
        *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        if (reduce_out) tma_reduce_add_4d(map, buf, col0, row0, bc1, bc0);
        else tma_store_4d(map, buf, col0, row0, bc1, bc0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}

// d(bias) of the layer whose pre-activations an op-2 epilogue multiplies with = column sums of this very result: taken
// here, from registers, instead of by a kernel that reads the matrix back (50 MB per BERT layer for the FFN).
// Compiled into the wide-tile kernels only (BN > 128: the shapes a GELU-backward GEMM runs at); the narrow-tile kernels
// carry the row epilogues and have no registers to spare -- for them the host adds the column sums with lg_reduce.
__device__ __forceinline__ void epi_colsum(const uint32_t (&v)[32], float* colsum, int col0, int N, int lane) {
    float x[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = __uint_as_float(v[k]);
    const float sum = warp_transpose_sum(x, lane);
    if (col0 + lane < N) atomicAdd(colsum + col0 + lane, sum);
}

// v: this thread's 32 accumulator values of the chunk (as raw bits), finished and stored in place
// this thread's 32 pre-activation values of a chunk (op 2): one 128-byte line; columns past N are never stored
__device__ __forceinline__ void epi_load_h(const EpiTile& t, int col0, float4 (&h)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        h[j] = (t.epi_op == 2 && t.aux != nullptr && col0 + 4 * j + 3 < t.N)
                   ? __ldg(reinterpret_cast<const float4*>(t.aux + col0 + 4 * j))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool COLSUM>
__device__ __forceinline__ void epi_emit(const EpiTile& t, uint32_t (&v)[32], float bcur, const float4 (&h)[8], int col0,
                                         uint8_t* buf, int lane) {
    if (t.bias != nullptr) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __shfl_sync(0xffffffffu, bcur, k));
    }
    if (!t.rows_live) return;
    if (t.epi_op == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[4 * j + 0] = __float_as_uint(epi_gelu_bwd(h[j].x, __uint_as_float(v[4 * j + 0])));
            v[4 * j + 1] = __float_as_uint(epi_gelu_bwd(h[j].y, __uint_as_float(v[4 * j + 1])));
            v[4 * j + 2] = __float_as_uint(epi_gelu_bwd(h[j].z, __uint_as_float(v[4 * j + 2])));
            v[4 * j + 3] = __float_as_uint(epi_gelu_bwd(h[j].w, __uint_as_float(v[4 * j + 3])));
        }
    }
    epi_store(t.map_c, v, buf, col0, t.row0, t.bc1, t.bc0, t.reduce_out, lane);
    if constexpr (COLSUM) {
        if (t.epi_op == 2 && t.colsum != nullptr) epi_colsum(v, t.colsum, col0, t.N, lane);
    }
    if (t.epi_op == 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(epi_gelu(__uint_as_float(v[k])));
        epi_store(t.map_aux, v, buf, col0, t.row0, t.bc1, t.bc0, 0, lane);
    }
}

// This warp takes chunks half, half + 2, ... of the tile (its partner on the same TMEM lane quarter takes the
// others; eight epilogue warps keep enough TMEM loads / stores in flight without software pipelining).
// `bfirst` / `hfirst` = bias slice / saved pre-activations (op 2) of its first chunk, loaded by the caller BEFORE
// it waited for the accumulator; inside the loop both are fetched one chunk ahead, so no global-memory latency
// sits between the TMEM read and the store.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const EpiTile& t, float bfirst, const float (&hfirst)[32], uint32_t taddr0,
                                              uint8_t* buf, int half, int lane) {
    constexpr int NC = BN / 32;
    float bnext = bfirst;
    float4 hnext[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hnext[j] = make_float4(hfirst[4 * j], hfirst[4 * j + 1], hfirst[4 * j + 2], hfirst[4 * j + 3]);
#pragma unroll 1
    for (int c = half; c < NC; c += 2) {
        const int col0 = t.n0 + c * 32;
        if (col0 >= t.N) break;
        uint32_t v[32];
        tmem_ld32(v, taddr0 + (uint32_t)(c * 32));
        const float bcur = bnext;
        float4 hcur[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hcur[j] = hnext[j];
        if (c + 2 < NC) {
            bnext = bias_slice(t.bias, col0 + 64 + lane, t.N);
            if (t.epi_op == 2) epi_load_h(t, col0 + 64, hnext);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        epi_emit<(BN > 128)>(t, v, bcur, hcur, col0, buf, lane);
    }
}

// Row-wise epilogues of the batched attention GEMMs (one tile spans the whole row: N <= BN <= 128).  A row's
// BN/32 chunks are split between the two warps of a TMEM lane quarter, which exchange their partial row
// statistic through each other's staging buffer (named barrier 1 + q, 64 threads).
//   op 3 (scores = Q K^T):  C = softmax(alpha * acc) along the row            (softmax_fwd fused away)
//   op 4 (dP = dO V^T):     C = alpha * P * (acc - sum_row(P * acc)), P = aux  (softmax_bwd fused away)
__device__ __forceinline__ void pair_bar(int q) {
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
}
// this thread's slice of the saved matrix (op 4): issued BEFORE the wait for the accumulator, so the global-memory
// latency hides behind the tile's main loop
template <int BN>
__device__ __forceinline__ void epilogue_rows_prefetch(const EpiTile& t, int half, float (&aux)[(BN / 32 + 1) / 2][32]) {
    constexpr int NC = BN / 32;
    constexpr int PER = (NC + 1) / 2;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = half + 2 * k;
        const int col0 = t.n0 + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t.epi_op == 4 && t.aux != nullptr && c < NC && col0 + 4 * j + 3 < t.N)
                h = __ldg(reinterpret_cast<const float4*>(t.aux + col0 + 4 * j));
            aux[k][4 * j] = h.x; aux[k][4 * j + 1] = h.y; aux[k][4 * j + 2] = h.z; aux[k][4 * j + 3] = h.w;
        }
    }
}

template <int BN>
__device__ __forceinline__ void epilogue_tile_rows(const EpiTile& t, uint32_t taddr0, uint8_t* buf, uint8_t* partner_buf,
                                                   int half, int q, int lane, float alpha,
                                                   const float (&aux)[(BN / 32 + 1) / 2][32]) {
    constexpr int NC = BN / 32;
    constexpr int PER = (NC + 1) / 2;        // chunks per warp: half, half + 2, ...
    float x[PER][32];
    float* mine = reinterpret_cast<float*>(buf);
    const float* other = reinterpret_cast<const float*>(partner_buf);
    // previous tile's TMA stores must have drained before the staging buffers carry statistics
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = half + 2 * k;
        const int col0 = t.n0 + c * 32;
        const bool live = c < NC && col0 < t.N;
        uint32_t raw[32];
        if (c < NC) {
            tmem_ld32(raw, taddr0 + (uint32_t)(c * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const bool ok = live && col0 + i < t.N;
            x[k][i] = ok ? __uint_as_float(raw[i]) : 0.f;
        }
    }
    if (t.epi_op == 3) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int col0 = t.n0 + (half + 2 * k) * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                // scores in units of log2(e), so that exp() below is a bare ex2
                x[k][i] = (half + 2 * k < NC && col0 + i < t.N) ? x[k][i] * (alpha * 1.4426950408889634f) : -INFINITY;
                m = fmaxf(m, x[k][i]);
            }
        }
        mine[lane] = m;
        pair_bar(q);
        m = fmaxf(m, other[lane]);
        pair_bar(q);                      // both warps have read the maxima
        float ssum = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                // only 8 warps per SM do this: ex2.approx (2^-22 relative) instead of expf; 2^-inf = 0 masks columns
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[k][i]) : "f"(x[k][i] - m));
                ssum += x[k][i];
            }
        mine[lane] = ssum;
        pair_bar(q);
        ssum += other[lane];
        pair_bar(q);                      // statistics consumed: the buffers may be used for staging again
        const float inv = 1.0f / ssum;
#pragma unroll
        for (int k = 0; k < PER; ++k)
#pragma unroll
            for (int i = 0; i < 32; ++i) x[k][i] *= inv;
    } else {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k)
#pragma unroll
            for (int i = 0; i < 32; ++i) dot += aux[k][i] * x[k][i];
        mine[lane] = dot;
        pair_bar(q);
        dot += other[lane];
        pair_bar(q);
#pragma unroll
        for (int k = 0; k < PER; ++k)
#pragma unroll
            for (int i = 0; i < 32; ++i) x[k][i] = aux[k][i] * (x[k][i] - dot) * alpha;
    }
    if (!t.rows_live) return;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = half + 2 * k;
        const int col0 = t.n0 + c * 32;
        if (c < NC && col0 < t.N) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(x[k][i]);
            epi_store(t.map_c, v, buf, col0, t.row0, t.bc1, t.bc0, 0, lane);
        }
    }
}

// CL = CTAs per cluster.  With CL = 2 the two CTAs of a cluster work on vertically adjacent output tiles
// (same N block): each loads its own A tile and HALF of the shared B tile, multicast to both; a stage may be
// refilled only when BOTH consumers have released it, so the MMA warps commit to the `empty` barrier of
// both CTAs.  Measured on B200: no faster than CL = 1 (a 2-CTA multicast does not reduce L2 traffic, and
// every CTA still stages the whole B tile), so only CL = 1 is instantiated; CTA pairs use the
// cta_group::2 kernel below, which really halves the B bytes per SM.
template <int ES, int BN, bool A_MN, bool B_MN, int STAGES, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    LG_PDL_TRIGGER();
    using G = Geo<ES>;
    constexpr int B_STAGE_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atoms
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(epi_base + EPI_WARPS * EPI_BUF_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);
    // work items reach the three roles through a small ring: the producer thread claims them (statically, or from a
    // global counter so that CTAs which start late -- behind another stream's kernel -- or run beside one take less)
    static_assert(CL == 1, "the work-item ring is per CTA");
    uint64_t* sfull = bars + 2 * STAGES + 6;
    uint64_t* sempty = sfull + SCHED_SLOTS;
    volatile int* ring = (volatile int*)(sempty + SCHED_SLOTS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // p.tiles_m counts groups of CL vertically adjacent tiles; work items are distributed over clusters
    const int items_per_batch = p.items_per_batch;
    // grouped launch: `groups` problems of identical shape (own operand / result maps), enumerated group-major
    const int items_per_group = items_per_batch * p.batches;
    const int work_items = items_per_group * p.groups;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = (int)blockIdx.x / CL, n_clusters = (int)gridDim.x / CL;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], EPI_WARPS * 32);
        }
        for (int s = 0; s < SCHED_SLOTS; ++s) {
            mbar_init(&sfull[s], 1);
            mbar_init(&sempty[s], 1 + EPI_WARPS);        // the MMA thread and one lane of every epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int g = 0; g < p.groups * p.kcat; ++g) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[g]));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[g]));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.c[g]));
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before anyone multicasts into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the preceding kernel
    LG_PDL_WAIT();

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int sslot = 0;
            uint32_t sphase = 0;
            int w = cluster_id < work_items ? cluster_id : -1;
            while (true) {
                // publish the item (or the end mark) to the MMA thread and the epilogue warps
                mbar_wait(&sempty[sslot], sphase ^ 1);
                ring[sslot] = w;
                mbar_arrive(&sfull[sslot]);
                if (++sslot == SCHED_SLOTS) {
                    sslot = 0;
                    sphase ^= 1;
                }
                if (w < 0) break;
                // claim the next one now: the atomic's round trip hides behind this item's loads
                const int w_next = p.sched ? n_clusters + (int)atomicAdd(p.sched, 1u) : w + n_clusters;
                const int grp = w / items_per_group, wg = w - grp * items_per_group;
                const int bi = wg / items_per_batch, wi = wg - bi * items_per_batch;
                const int bc0 = bi / p.batch1, bc1 = bi - bc0 * p.batch1;
                const TcItem item = decode_item(p, wi);
                const int tile = item.tile, kb0 = item.kb0, kb1 = item.kb1;
                const int m0 = ((tile % p.tiles_m) * CL + crank) * BM, n0 = (tile / p.tiles_m) * BN;
                const int kspan = kb1 - kb0;
                for (int it = 0; it < kspan * p.kcat; ++it) {
                    // K-concatenated operands: chunk kc supplies k-blocks kb0..kb1 of its own (A, B) pair
                    const int kc = it / kspan, kb = kb0 + (it - kc * kspan);
                    const CUtensorMap* map_a = &maps.a[grp + kc];
                    const CUtensorMap* map_b = &maps.b[grp + kc];
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = stage_base + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    const int k0 = kb * G::BKE;
                    if (!A_MN) {
                        tma_load_4d(map_a, &full[stage], sa, k0, m0, bc1, bc0);
                    } else if (p.a_chunked) {
                        tma_load_5d(map_a, &full[stage], sa, 0, k0, m0 >> G::CH_SHIFT, bc1, bc0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / G::CH; ++j)
                            tma_load_4d(map_a, &full[stage], sa + j * G::CHUNK_BYTES, m0 + G::CH * j, k0, bc1, bc0);
                    }
                    if (CL == 1) {
                        if (!B_MN) {
                            tma_load_4d(map_b, &full[stage], sb, k0, n0, bc1, bc0);
                        } else if (p.b_chunked) {
                            tma_load_5d(map_b, &full[stage], sb, 0, k0, n0 >> G::CH_SHIFT, bc1, bc0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / G::CH; ++j)
                                tma_load_4d(map_b, &full[stage], sb + j * G::CHUNK_BYTES, n0 + G::CH * j, k0, bc1, bc0);
                        }
                    } else {
                        // this CTA fetches its half of the B tile for the whole cluster
                        constexpr uint16_t kAll = (uint16_t)((1u << CL) - 1);
                        if (!B_MN) {
                            constexpr int HALF = BN / CL;   // rows of B per CTA (the map's box height)
                            tma_load_4d_mc(map_b, &full[stage], sb + crank * HALF * 128, k0, n0 + crank * HALF, bc1,
                                           bc0, kAll);
                        } else {
                            constexpr int PER = (BN / G::CH) / CL;
#pragma unroll
                            for (int jj = 0; jj < PER; ++jj) {
                                const int j = crank * PER + jj;
                                tma_load_4d_mc(map_b, &full[stage], sb + j * G::CHUNK_BYTES, n0 + G::CH * j, k0, bc1, bc0,
                                               kAll);
                            }
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                w = w_next < work_items ? w_next : -1;
            }
            if (p.sched) {
                // this CTA has made its last claim; the last CTA to get here leaves the counters at zero
                if (atomicAdd(p.sched + 1, 1u) == (unsigned)n_clusters - 1) {
                    atomicExch(p.sched, 0u);
                    atomicExch(p.sched + 1, 0u);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ========================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc<ES>(BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int sslot = 0;
            uint32_t sphase = 0;
            while (true) {
                mbar_wait(&sfull[sslot], sphase);
                const int w = ring[sslot];
                mbar_arrive(&sempty[sslot]);
                if (++sslot == SCHED_SLOTS) {
                    sslot = 0;
                    sphase ^= 1;
                }
                if (w < 0) break;
                const TcItem item = decode_item(p, (w % items_per_group) % items_per_batch);
                const int kb0 = item.kb0, kb1 = item.kb1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                const int n_it = (kb1 - kb0) * p.kcat;
                for (int it = 0; it < n_it; ++it) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(stage_base + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        // K-major : SWIZZLE_128B; 32 bytes along the swizzled row per k-step, 8-row groups 1024 B apart
                        // MN-major: SWIZZLE_128B_BASE32B; 8 k-rows (1024 B) per k-step, 4-row groups 512 B apart,
                        //           32-element MN chunks BK*128 B apart
                        const uint64_t da = A_MN ? umma_desc(sa + ks * G::KSTEP_MN_BYTES, G::CHUNK_BYTES, G::MN_SBO, G::MN_LAYOUT)
                                                 : umma_desc(sa + ks * 32, 16, 1024, 2);
                        const uint64_t db = B_MN ? umma_desc(sb + ks * G::KSTEP_MN_BYTES, G::CHUNK_BYTES, G::MN_SBO, G::MN_LAYOUT)
                                                 : umma_desc(sb + ks * 32, 16, 1024, 2);
                        const uint32_t accum = (it > 0 || ks > 0) ? 1u : 0u;
                        umma_issue<ES, 1>(d_tmem, da, db, idesc, accum);
                    }
                    // frees the smem stage once the MMAs above have consumed it (in every CTA that fills it)
                    if (CL == 1) {
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         smem_u32(&empty[stage]))
                                     : "memory");
                    } else {
                        asm volatile(
                            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
                            "[%0], %1;" ::"r"(smem_u32(&empty[stage])),
                            "h"((uint16_t)((1u << CL) - 1))
                            : "memory");
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                // accumulator complete -> epilogue
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&tfull[acc]))
                             : "memory");
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================================== epilogue ==========================================
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int ew = warp - 2;                // staging slot
        const int half = ew >> 2;               // which of the two warps on this lane quarter
        uint8_t* buf0 = epi_base + ew * EPI_BUF_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        int sslot = 0;
        uint32_t sphase = 0;
        while (true) {
            mbar_wait(&sfull[sslot], sphase);
            const int w = ring[sslot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty[sslot]);
            if (++sslot == SCHED_SLOTS) {
                sslot = 0;
                sphase ^= 1;
            }
            if (w < 0) break;
            const int grp = w / items_per_group, wg = w - grp * items_per_group;
            const CUtensorMap* map_c = &maps.c[grp];
            const float* bias = p.bias[grp];
            const int bi = wg / items_per_batch, wi = wg - bi * items_per_batch;
            const int bc0 = bi / p.batch1, bc1 = bi - bc0 * p.batch1;
            const TcItem item = decode_item(p, wi);
            const int tile = item.tile, split = item.split;
            const int m0 = ((tile % p.tiles_m) * CL + crank) * BM, n0 = (tile / p.tiles_m) * BN;
            EpiTile t;
            t.map_c = map_c;
            t.map_aux = &maps.aux;
            t.epi_op = p.epi_op;
            t.bias = split == 0 ? bias : nullptr;
            t.row0 = m0 + 32 * q;
            t.aux = ((p.epi_op == 2 || p.epi_op == 4) && t.row0 + lane < p.M)
                        ? p.aux + bc0 * p.aux_sb0 + bc1 * p.aux_sb1 + (long long)(t.row0 + lane) * p.aux_ld
                        : nullptr;
            t.n0 = n0;
            t.N = p.N;
            t.bc1 = bc1;
            t.bc0 = bc0;
            t.reduce_out = p.reduce_out | item.partial;
            t.rows_live = t.row0 < p.M;
            t.colsum = p.colsum;
            const float bfirst = bias_slice(t.bias, n0 + half * 32 + lane, p.N);
            // operands the epilogue reads from global memory, fetched before the wait for the accumulator:
            // op 2: the first chunk's pre-activations (row_aux[0]); op 4: this thread's slice of the probabilities
            float row_aux[BN <= 128 ? (BN / 32 + 1) / 2 : 1][32];
            bool rows_op = false;
            if constexpr (BN <= 128) rows_op = p.epi_op == 4;
            if constexpr (BN <= 128) {
                if (rows_op) epilogue_rows_prefetch<BN>(t, half, row_aux);
            }
            if (!rows_op) {
                float4 h4[8];
                epi_load_h(t, n0 + half * 32, h4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    row_aux[0][4 * j] = h4[j].x; row_aux[0][4 * j + 1] = h4[j].y;
                    row_aux[0][4 * j + 2] = h4[j].z; row_aux[0][4 * j + 3] = h4[j].w;
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN);
            if constexpr (BN <= 128) {
                if (p.epi_op >= 3)
                    epilogue_tile_rows<BN>(t, taddr0, buf0, epi_base + (ew ^ 4) * EPI_BUF_BYTES, half, q, lane,
                                           p.epi_alpha, row_aux);
                else
                    epilogue_tile<BN>(t, bfirst, row_aux[0], taddr0, buf0, half, lane);
            } else {
                epilogue_tile<BN>(t, bfirst, row_aux[0], taddr0, buf0, half, lane);
            }
            // this warp is done reading the accumulator stage
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&tempty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================
// CTA-pair variant: two CTAs on neighbouring SMs compute one 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages only ITS 128 rows of A and ITS half (BN/2 rows) of B: shared-memory footprint and
// L2 -> SM traffic per k-step drop from 16 KB + BN*128 B to 16 KB + BN*64 B per CTA, so the ring holds
// 6 stages instead of 4 at BN = 256 -- what the fp32-operand (tf32) GEMM needs to cover L2 latency.
//   both CTAs : warp 0 = TMA producer (completion credited to the leader's `full` barrier)
//               warp 1 = collective TMEM allocation; in the leader also the single MMA-issuing thread
//               warps 2..9 = epilogue on the CTA's own 128 accumulator rows
//   leader    : waits `full`, issues m256 nBN k8 MMAs, tcgen05.commit(multicast) -> `empty` of both CTAs,
//               and `tfull` of both CTAs when a tile is complete; waits `tempty` (16 warp arrivals: 8 local,
//               4 remote) before overwriting an accumulator stage.
template <int ES, int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_2cta_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    LG_PDL_TRIGGER();
    using G = Geo<ES>;
    constexpr int HB = BN / 2;                          // B rows staged per CTA
    constexpr int B_STAGE_BYTES = HB * 128;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(epi_base + EPI_WARPS * EPI_BUF_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items_per_batch = p.items_per_batch;   // p.tiles_m counts 256-row tile pairs
    // grouped launch: `groups` problems of identical shape (own operand / result maps), enumerated group-major;
    // kcat > 1: operand pairs concatenated along K into one result (both as in the independent-CTA kernel)
    const int items_per_group = items_per_batch * p.batches;
    const int work_items = items_per_group * p.groups;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = (int)blockIdx.x / 2, n_clusters = (int)gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 2 * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int g = 0; g < p.groups * p.kcat; ++g) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[g]));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[g]));
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.c[g]));
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    LG_PDL_WAIT();

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) ===========================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = cluster_id; w < work_items; w += n_clusters) {
                const int grp = w / items_per_group, wg = w - grp * items_per_group;
                const int bi = wg / items_per_batch, wi = wg - bi * items_per_batch;
                const int bc0 = bi / p.batch1, bc1 = bi - bc0 * p.batch1;
                const TcItem item = decode_item(p, wi);
                const int tile = item.tile;
                const int m0 = ((tile % p.tiles_m) * 2 + crank) * BM;
                const int n0 = (tile / p.tiles_m) * BN + crank * HB;     // this CTA's half of the B tile
                const int kb0 = item.kb0, kb1 = item.kb1;
                const int kspan = kb1 - kb0;
                for (int it = 0; it < kspan * p.kcat; ++it) {
                    const int kc = it / kspan, kb = kb0 + (it - kc * kspan);
                    const CUtensorMap* map_a = &maps.a[grp + kc];
                    const CUtensorMap* map_b = &maps.b[grp + kc];
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = stage_base + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_STAGE_BYTES;
                    // the leader's barrier collects the bytes of both CTAs
                    if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
                    const int k0 = kb * G::BKE;
                    if (!A_MN) {
                        tma_load_4d_2sm(map_a, &full[stage], sa, k0, m0, bc1, bc0);
                    } else if (p.a_chunked) {
                        tma_load_5d_2sm(map_a, &full[stage], sa, 0, k0, m0 >> G::CH_SHIFT, bc1, bc0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / G::CH; ++j)
                            tma_load_4d_2sm(map_a, &full[stage], sa + j * G::CHUNK_BYTES, m0 + G::CH * j, k0, bc1, bc0);
                    }
                    if (!B_MN) {
                        tma_load_4d_2sm(map_b, &full[stage], sb, k0, n0, bc1, bc0);
                    } else if (p.b_chunked) {
                        tma_load_5d_2sm(map_b, &full[stage], sb, 0, k0, n0 >> G::CH_SHIFT, bc1, bc0);
                    } else {
#pragma unroll
                        for (int j = 0; j < HB / G::CH; ++j)
                            tma_load_4d_2sm(map_b, &full[stage], sb + j * G::CHUNK_BYTES, n0 + G::CH * j, k0, bc1, bc0);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) =======================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc<ES>(2 * BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int w = cluster_id; w < work_items; w += n_clusters) {
                const TcItem item = decode_item(p, (w % items_per_group) % items_per_batch);
                const int n_it = (item.kb1 - item.kb0) * p.kcat;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int it = 0; it < n_it; ++it) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(stage_base + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint64_t da = A_MN ? umma_desc(sa + ks * G::KSTEP_MN_BYTES, G::CHUNK_BYTES, G::MN_SBO, G::MN_LAYOUT)
                                                 : umma_desc(sa + ks * 32, 16, 1024, 2);
                        const uint64_t db = B_MN ? umma_desc(sb + ks * G::KSTEP_MN_BYTES, G::CHUNK_BYTES, G::MN_SBO, G::MN_LAYOUT)
                                                 : umma_desc(sb + ks * 32, 16, 1024, 2);
                        const uint32_t accum = (it > 0 || ks > 0) ? 1u : 0u;
                        umma_issue<ES, 2>(d_tmem, da, db, idesc, accum);
                    }
                    asm volatile(
                        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], "
                        "%1;" ::"r"(smem_u32(&empty[stage])),
                        "h"((uint16_t)3)
                        : "memory");
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                asm volatile(
                    "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                        smem_u32(&tfull[acc])),
                    "h"((uint16_t)3)
                    : "memory");
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================================== epilogue (both CTAs, own 128 rows) ==================
        const int q = warp & 3;
        const int ew = warp - 2;
        const int half = ew >> 2;
        uint8_t* buf0 = epi_base + ew * EPI_BUF_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = cluster_id; w < work_items; w += n_clusters) {
            const int grp = w / items_per_group, wg = w - grp * items_per_group;
            const int bi = wg / items_per_batch, wi = wg - bi * items_per_batch;
            const int bc0 = bi / p.batch1, bc1 = bi - bc0 * p.batch1;
            const TcItem item = decode_item(p, wi);
            const int tile = item.tile, split = item.split;
            const int m0 = ((tile % p.tiles_m) * 2 + crank) * BM, n0 = (tile / p.tiles_m) * BN;
            EpiTile t;
            t.map_c = &maps.c[grp];
            t.map_aux = &maps.aux;
            t.epi_op = p.epi_op;
            t.bias = split == 0 ? p.bias[grp] : nullptr;
            t.row0 = m0 + 32 * q;
            t.aux = (p.epi_op == 2 && t.row0 + lane < p.M) ? p.aux + (long long)(t.row0 + lane) * p.aux_ld : nullptr;
            t.n0 = n0;
            t.N = p.N;
            t.bc1 = bc1;
            t.bc0 = bc0;
            t.reduce_out = p.reduce_out | item.partial;
            t.rows_live = t.row0 < p.M;
            t.colsum = p.colsum;
            const float bfirst = bias_slice(t.bias, n0 + half * 32 + lane, p.N);
            float hfirst[32];
            {
                float4 h4[8];
                epi_load_h(t, n0 + half * 32, h4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    hfirst[4 * j] = h4[j].x; hfirst[4 * j + 1] = h4[j].y; hfirst[4 * j + 2] = h4[j].z; hfirst[4 * j + 3] = h4[j].w;
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            epilogue_tile<BN>(t, bfirst, hfirst, tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN), buf0, half,
                              lane);
            // one arrival per warp on the LEADER's barrier: the accumulator stage of this CTA is drained
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tempty[acc], 0);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------
// rank-4 tensor map (es = element size: 4 fp32, 2 bf16): dim0 = contiguous extent, dim1 = strided extent,
// dim2 = batch1, dim3 = batch0
struct BatchDims {
    int64_t n1, s1, n0, s0;   // extents and element strides of batch1 (inner) and batch0 (outer)
};
int make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t outer_stride_elems,
             const BatchDims& bd, int box_inner, int box_outer, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
             int es = 4) {
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)bd.n1, (cuuint64_t)bd.n0};
    // a size-1 dim never advances: give it any legal (16-byte multiple, non-zero) stride
    const int64_t q = 16 / es;
    const int64_t s_outer = outer > 1 ? outer_stride_elems : ((inner + q - 1) / q * q);
    const int64_t dflt = s_outer * (outer > 0 ? outer : 1);
    cuuint64_t strides[3] = {(cuuint64_t)s_outer * es, (cuuint64_t)(bd.n1 > 1 ? bd.s1 : dflt) * es,
                             (cuuint64_t)(bd.n0 > 1 ? bd.s0 : dflt * (bd.n1 > 0 ? bd.n1 : 1)) * es};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                          const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// rank-5 map of an MN-major operand (k rows of `mn` contiguous elements, row pitch `ld`), CH = 128 / es elements
// per chunk: (CH, k, ceil(mn/CH), batch1, batch0) with element strides (1, ld, CH, s1, s0);
// box = (CH, k-block, chunks, 1, 1)
int make_map_chunked(CUtensorMap* map, const void* base, int64_t mn, int64_t k, int64_t ld, const BatchDims& bd,
                     int box_chunks, int es = 4) {
    const int64_t ch = 128 / es;
    const int64_t chunks = (mn + ch - 1) / ch;
    cuuint64_t dims[5] = {(cuuint64_t)ch, (cuuint64_t)k, (cuuint64_t)chunks, (cuuint64_t)bd.n1, (cuuint64_t)bd.n0};
    const int64_t s_k = k > 1 ? ld : chunks * ch;
    const int64_t dflt = s_k * (k > 0 ? k : 1);
    cuuint64_t strides[4] = {(cuuint64_t)s_k * es, 128, (cuuint64_t)(bd.n1 > 1 ? bd.s1 : dflt) * es,
                             (cuuint64_t)(bd.n0 > 1 ? bd.s0 : dflt * (bd.n1 > 0 ? bd.n1 : 1)) * es};
    cuuint32_t box[5] = {(cuuint32_t)ch, (cuuint32_t)(128 / es), (cuuint32_t)box_chunks, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                          const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          es == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (chunked) failed with CUresult %d", (int)r);
    return 0;
}
// chunks must tile a row exactly: a partial last chunk would read past the end of the row (and, for the last
// row of a view, possibly past the end of the allocation), which the tensor map cannot clip
inline bool chunkable(int64_t mn, int es) {
    static const bool off = getenv("LG_GEMM_NO_CHUNKED") != nullptr;
    return !off && mn % (128 / es) == 0;
}

// LG_GEMM_SMEM_CUT_KB (compile-time experiment knob): shrink the operand ring by that many KB so that another
// kernel's CTAs (NCCL) can become resident next to a GEMM CTA.  Measured with 48 KB at 2 GPUs: the step without
// any exchange gets 5 % slower and the exposed communication does not shrink (1.45 ms) -- left at 0.
#ifndef LG_GEMM_SMEM_CUT_KB
#define LG_GEMM_SMEM_CUT_KB 0
#endif
template <int BN>
constexpr int stages_for() {
    constexpr int full = BN == 256 ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
    constexpr int cut = (LG_GEMM_SMEM_CUT_KB * 1024 + (A_STAGE_BYTES + BN * 128) - 1) / (A_STAGE_BYTES + BN * 128);
    return full - cut >= 2 ? full - cut : 2;
}

template <int BN>
constexpr size_t smem_for() {
    return (size_t)stages_for<BN>() * (A_STAGE_BYTES + BN * 128) + EPI_WARPS * EPI_BUF_BYTES +
           (2 * stages_for<BN>() + 4) * 8 + 16 + SCHED_SLOTS * (8 + 8 + 4) + 1024;
}

inline bool pdl_enabled() {
    static const bool on = getenv("LG_GEMM_NO_PDL") == nullptr;
    return on;
}

template <int BN>
constexpr int stages2_for() {
    // per-CTA stage = 16 KB of A + BN*64 B of B; keep ~192 KB of operands in flight
    return ((192 - LG_GEMM_SMEM_CUT_KB) * 1024) / (A_STAGE_BYTES + BN * 64) > 10
               ? 10
               : ((192 - LG_GEMM_SMEM_CUT_KB) * 1024) / (A_STAGE_BYTES + BN * 64);
}
template <int BN>
constexpr size_t smem2_for() {
    return (size_t)stages2_for<BN>() * (A_STAGE_BYTES + BN * 64) + EPI_WARPS * EPI_BUF_BYTES +
           (2 * stages2_for<BN>() + 4) * 8 + 16 + 1024;
}

template <int ES, int BN, bool A_MN, bool B_MN>
int launch_2cta(const TcMaps& maps, const TcParams& p, int grid) {
    constexpr int ST = stages2_for<BN>();
    constexpr size_t smem = smem2_for<BN>();
    auto kern = gemm_tc_2cta_kernel<ES, BN, A_MN, B_MN, ST>;
    static bool attr_done = false;
    if (!attr_done) {
        LG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    LG_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p));
    LG_CHECK_LAUNCH();
    return 0;
}

template <int ES, int BN>
int launch_2cta_bn(bool a_mn, bool b_mn, const TcMaps& maps,
                   const TcParams& p, int grid) {
    if (!a_mn && !b_mn) return launch_2cta<ES, BN, false, false>(maps, p, grid);
    if (!a_mn && b_mn) return launch_2cta<ES, BN, false, true>(maps, p, grid);
    if (a_mn && !b_mn) return launch_2cta<ES, BN, true, false>(maps, p, grid);
    return launch_2cta<ES, BN, true, true>(maps, p, grid);
}

template <int ES, int BN, bool A_MN, bool B_MN, int CL>
int launch_cfg(const TcMaps& maps, const TcParams& p, int grid) {
    constexpr int ST = stages_for<BN>();
    constexpr size_t smem = smem_for<BN>();
    auto kern = gemm_tc_kernel<ES, BN, A_MN, B_MN, ST, CL>;
    static bool attr_done = false;
    if (!attr_done) {
        LG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream();
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    LG_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p));
    LG_CHECK_LAUNCH();
    return 0;
}

template <int ES, int BN, int CL>
int launch_bn(bool a_mn, bool b_mn, const TcMaps& maps, const TcParams& p, int grid) {
    if (!a_mn && !b_mn) return launch_cfg<ES, BN, false, false, CL>(maps, p, grid);
    if (!a_mn && b_mn) return launch_cfg<ES, BN, false, true, CL>(maps, p, grid);
    if (a_mn && !b_mn) return launch_cfg<ES, BN, true, false, CL>(maps, p, grid);
    return launch_cfg<ES, BN, true, true, CL>(maps, p, grid);
}
template <int ES>
int launch_any(bool pair_mma, int bn, bool a_mn, bool b_mn, const TcMaps& maps, const TcParams& p, int grid) {
    if (pair_mma) {
        switch (bn) {
            case 256: return launch_2cta_bn<ES, 256>(a_mn, b_mn, maps, p, grid);
            case 192: return launch_2cta_bn<ES, 192>(a_mn, b_mn, maps, p, grid);
            case 128: return launch_2cta_bn<ES, 128>(a_mn, b_mn, maps, p, grid);
            default: return launch_2cta_bn<ES, 64>(a_mn, b_mn, maps, p, grid);
        }
    }
    switch (bn) {
        case 256: return launch_bn<ES, 256, 1>(a_mn, b_mn, maps, p, grid);
        case 192: return launch_bn<ES, 192, 1>(a_mn, b_mn, maps, p, grid);
        case 128: return launch_bn<ES, 128, 1>(a_mn, b_mn, maps, p, grid);
        default: return launch_bn<ES, 64, 1>(a_mn, b_mn, maps, p, grid);
    }
}

// strides must be 16-byte multiples for TMA: q = 16 / element size elements
bool k_major(int64_t s_mn, int64_t s_k, int64_t extent_mn, int q) { return s_k == 1 && (s_mn % q == 0 || extent_mn == 1); }
bool mn_major(int64_t s_mn, int64_t s_k, int64_t extent_k, int q) { return s_mn == 1 && (s_k % q == 0 || extent_k == 1); }

struct Plan {
    int bn, splits, tiles_m, tiles_n, kblocks, kper;
};

// Counters of the dynamic tile order, one pair per stream a GEMM can be launched on (compute, side, collective):
// launches on one stream are ordered, so a pair is always back at zero when the next launch starts claiming.
// Opt-in (LG_GEMM_DYNAMIC=1): measured on the BERT-base step, where weight-gradient GEMMs on the side stream run beside
// the activation-gradient GEMMs, 8.122 / 8.124 ms against 8.132 / 8.144 ms with the static order -- inside the noise,
// so the default stays the static order.
unsigned int* sched_counters() {
    static unsigned int* base = nullptr;
    static const bool on = getenv("LG_GEMM_DYNAMIC") != nullptr;
    if (!on) return nullptr;
    if (!base) {
        if (capturing()) return nullptr;          // (first GEMM ever inside a capture: static order for this graph)
        unsigned int* q = nullptr;
        if (cudaMalloc((void**)&q, 3 * 128) != cudaSuccess) return nullptr;
        cudaMemset(q, 0, 3 * 128);
        cudaDeviceSynchronize();
        base = q;
    }
    return base + 32 * alt_stream_index();
}

Plan choose_plan(int64_t M, int64_t N, int64_t K, int64_t batches, bool allow_split, bool tail_ok, int es) {
    const int sms = gemm_sms();
    const int bke = 128 / es;            // elements per k-block
    const int kblocks = (int)((K + bke - 1) / bke);
    Plan best{};
    double best_score = -1.0;
    const int cands[4] = {256, 192, 128, 64};
    static const int force_bn = getenv("LG_GEMM_BN") ? atoi(getenv("LG_GEMM_BN")) : 0;        // tuning knobs
    static const int force_splits = getenv("LG_GEMM_SPLITS") ? atoi(getenv("LG_GEMM_SPLITS")) : 0;
    for (int bn : cands) {
        if (force_bn && bn != force_bn) continue;
        if (bn > 64 && N <= bn / 2) continue;            // mostly padding
        const int tm = (int)((M + BM - 1) / BM), tn = (int)((N + bn - 1) / bn);
        const int tiles = tm * tn * (int)batches;
        int splits = 1;
        if (tiles < sms && allow_split) {
            splits = sms / tiles;
            const int max_splits = kblocks / 8 > 0 ? kblocks / 8 : 1;   // >= 8 k-blocks (256 of K) per split
            if (splits > max_splits) splits = max_splits;
            if (splits > 16) splits = 16;
            if (splits < 1) splits = 1;
        }
        if (force_splits && allow_split) splits = force_splits < kblocks ? force_splits : kblocks;
        // very long K (the decoder's dX, K = vocabulary): the reduce-add of the partials is negligible next to
        // the main loop, so also try the split counts that fill 2 and 3 whole waves (96 tiles x 3 = 288 items on
        // 148 SMs instead of 128 one-wave tiles)
        int cand_splits[3] = {splits, splits, splits};
        if (tiles < sms && allow_split && !force_splits && kblocks >= 512) {
            cand_splits[1] = (2 * sms) / tiles;
            cand_splits[2] = (3 * sms) / tiles;
        }
        const double useful = ((double)M * N) / ((double)tm * BM * tn * bn);   // padding waste
        const double shape = bn == 256 ? 1.0 : (bn == 192 ? 0.95 : (bn == 128 ? 0.85 : 0.6));
        for (int ci = 0; ci < 3; ++ci) {
            int sp = cand_splits[ci];
            if (sp < 1) sp = 1;
            if (sp > 16) sp = 16;
            if (ci > 0 && (sp == cand_splits[0] || kblocks / sp < 64)) continue;
            int kper = (kblocks + sp - 1) / sp;
            sp = (kblocks + kper - 1) / kper;
            const int items = tiles * sp;
            const int waves = (items + sms - 1) / sms;
            double util = (double)items / ((double)waves * sms);
            if (tail_ok && sp == 1 && tiles > sms) {
                // the launch code cuts the tiles of a mostly idle last wave into K-ranges (tail-wave split):
                // that wave then costs 1/ranges of a full one
                const bool pair = tiles >= 2 * sms && tm >= 2;
                const int clusters = pair ? (sms / 2 > 0 ? sms / 2 : 1) : sms;
                const int items_c = pair ? ((tm + 1) / 2) * tn : tiles;
                const int tail = items_c % clusters;
                int ts = tail > 0 ? clusters / tail : 1;
                if (ts > 4) ts = 4;
                while (ts > 1 && kblocks / ts < 16) --ts;
                // (0.94: the cut tiles pay a zero-fill and reduce-add stores)
                if (ts > 1) util = 0.94 * (double)items_c / (((double)(items_c / clusters) + 1.0 / ts) * clusters);
            }
            const double split_cost = sp > 1 ? 0.92 : 1.0;
            const double score = util * useful * shape * split_cost;
            if (score > best_score) {
                best_score = score;
                best = Plan{bn, sp, tm, tn, kblocks, kper};
            }
        }
    }
    return best;
}

}  // namespace

namespace lg {

// tf32 mode: fp32 operands.  bf16 mode: bf16 operands (dtype LG_BF16; strides in bf16 elements) -- asked with
// dtype LG_F32 it answers for the bf16 staging copy of an fp32 operand with the same element strides (pass no
// pointers then).  The result is fp32 in both modes.
int gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, const void* c) {
    if (!((mode == LG_GEMM_TF32_TC && dtype == LG_F32) || (mode == LG_GEMM_BF16_TC && (dtype == LG_BF16 || dtype == LG_F32))))
        return 0;
    const int q = mode == LG_GEMM_BF16_TC ? 8 : 4;     // operand elements per 16 bytes
    const int64_t batches = d->batch0 * d->batch1;
    if (batches < 1 || batches > 65535) return 0;
    if (d->M < 1 || d->N < 1 || d->K < 1) return 0;
    // batch strides go into the TMA descriptor: 16-byte multiples, no broadcast (stride 0) over a real batch dim
    const int64_t bs[6][2] = {{d->batch1, d->sa_b1}, {d->batch0, d->sa_b0}, {d->batch1, d->sb_b1},
                              {d->batch0, d->sb_b0}, {d->batch1, d->sc_b1}, {d->batch0, d->sc_b0}};
    for (int i = 0; i < 6; ++i)
        if (bs[i][0] > 1 && (bs[i][1] <= 0 || bs[i][1] % (i < 4 ? q : 4) != 0)) return 0;
    if (d->M > 0x7fffffff || d->N > 0x7fffffff || d->K > 0x7fffffff) return 0;
    // tiny problems are launch bound either way; keep them on the exact path
    if ((double)d->M * d->N * d->K * (double)batches < 64.0 * 64.0 * 64.0 * 8) return 0;
    const bool a_ok = k_major(d->sa_m, d->sa_k, d->M, q) || mn_major(d->sa_m, d->sa_k, d->K, q);
    const bool b_ok = k_major(d->sb_n, d->sb_k, d->N, q) || mn_major(d->sb_n, d->sb_k, d->K, q);
    const bool c_ok = d->sc_n == 1 && (d->sc_m % 4 == 0 || d->M == 1);
    if (!a_ok || !b_ok || !c_ok) return 0;
    if (a && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15)) return 0;
    return 1;
}

// `groups` problems sharing one LgGemmDesc (shape, strides) but with their own operand / result / bias
// pointers run as ONE launch of the 1-CTA kernel (e.g. the Q, K and V projections of an attention block:
// 3 x 128 tiles instead of three one-wave launches).  Several groups may name the same C: with
// accumulate they all reduce-add into it (dX = sum_g dY_g W_g).
int gemm_tc_grouped(int mode, const LgGemmDesc* d, int groups, const void* const* a, const void* const* b, void* const* c,
                    const void* const* bias, int accumulate, int epi_op, void* aux, int64_t aux_ld, double alpha) {
    if (load_encode()) return 1;
    const int es = mode == LG_GEMM_BF16_TC ? 2 : 4;    // operand element size
    const int q = 16 / es, bke = 128 / es, ch = 128 / es;
    const CUtensorMapSwizzle mn_swz = es == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (epi_op == 1 || epi_op == 2) {
        LG_REQUIRE(groups == 1 && d->batch0 * d->batch1 == 1 && !accumulate, "gemm_tc: epilogue ops need one plain GEMM");
        LG_REQUIRE(aux && (((uintptr_t)aux) & 15) == 0 && aux_ld % 4 == 0 && aux_ld >= d->N && d->N % 4 == 0,
                   "gemm_tc: epilogue operand must be 16-byte aligned with a row pitch that is a multiple of 4");
    }
    if (epi_op >= 3) {
        LG_REQUIRE(groups == 1 && !accumulate && !bias && d->N <= 128, "gemm_tc: row epilogues need N <= 128, no bias");
        LG_REQUIRE(epi_op == 3 || (aux && (((uintptr_t)aux) & 15) == 0 && aux_ld % 4 == 0 && d->N % 4 == 0),
                   "gemm_tc: row epilogue operand must be 16-byte aligned with a row pitch that is a multiple of 4");
    }
    LG_REQUIRE(groups >= 1 && groups <= LG_MAX_GROUPS, "gemm_tc: 1..%d groups per launch", LG_MAX_GROUPS);
    const int64_t M = d->M, N = d->N, K = d->K;
    const bool a_mn = !k_major(d->sa_m, d->sa_k, M, q);
    const bool b_mn = !k_major(d->sb_n, d->sb_k, N, q);
    const int64_t batches = d->batch0 * d->batch1;
    // every group naming the same C: one K-concatenated product C (+)= sum_g A_g B_g, accumulated in TMEM
    bool same_c = groups > 1;
    for (int g = 1; g < groups; ++g) same_c = same_c && c[g] == c[0];
    if (!same_c)
        for (int g = 0; g < groups; ++g)
            for (int h = 0; h < g; ++h) LG_REQUIRE(c[g] != c[h], "gemm_tc: groups must share ONE result or none");
    if (same_c)
        for (int g = 1; g < groups; ++g)
            LG_REQUIRE(!bias || bias[g] == bias[0], "gemm_tc: K-concatenated groups share one bias");
    const int n_problems = same_c ? 1 : groups;
    // split-K partials meet in C by reduce-add: C must be zeroed first (done below for one plain matrix) or
    // already hold the value being accumulated into
    Plan pl = choose_plan(M, N, K, batches * n_problems, !epi_op && (accumulate || batches * n_problems == 1),
                          !epi_op && n_problems == 1 && batches == 1 && !same_c, es);
    if (epi_op >= 3) {
        // the whole row must live in one tile
        const int bn = N <= 64 ? 64 : 128;
        pl = Plan{bn, 1, (int)((M + BM - 1) / BM), 1, pl.kblocks, pl.kblocks};
    }
    int rc;
    const BatchDims ba{d->batch1, d->sa_b1, d->batch0, d->sa_b0}, bb{d->batch1, d->sb_b1, d->batch0, d->sb_b0},
        bc{d->batch1, d->sc_b1, d->batch0, d->sc_b0};
    // CTA pairs (cta_group::2) take 256-row tiles and split the B tile between the two SMs; needs >= 2 tile rows.
    // Every SM then stages 16 KB of A + BN/2 rows of B per k-block instead of 16 KB + BN rows: at 4-byte operands the
    // independent-CTA kernel asks L2 for 125-140 GB/s per SM at full tensor rate, the pair kernel for 85-100.
    // r2, final kernels, every shape of the BERT step inside a graph (profiles/r2_gemm_in_graph{,_pairs}.jsonl): pairs are
    // never slower and 3-10 % faster on the N = 768 / K = 3072 shapes (40.9 -> 37.1 us), the weight gradients
    // (35.1 -> 31.9) and the decoder's dX (339 -> 305) -- so pairs are the default wherever the launch allows them
    // (LG_GEMM_PAIR=0: independent CTAs; r1 used pairs only on >= 2-wave problems).
    static const int force_pair = getenv("LG_GEMM_PAIR") ? atoi(getenv("LG_GEMM_PAIR")) : -1;
    // (an MN-major B tile is staged in chunks of 128 bytes of N: each CTA's half, BN / 2 columns, must be whole chunks --
    //  bf16 with BN = 192 or 64 is not, and keeps independent CTAs)
    const bool halves_ok = !b_mn || ((pl.bn / 2) % (128 / es) == 0);
    const bool pair_mma = batches == 1 && epi_op < 3 && pl.tiles_m >= 2 && halves_ok && force_pair != 0;
    const int cl = pair_mma ? 2 : 1;
    TcMaps maps;
    TcParams p;
    const bool a_ch = a_mn && chunkable(M, es), b_ch = b_mn && chunkable(N, es);
    p.a_chunked = a_ch ? 1 : 0;
    p.b_chunked = b_ch ? 1 : 0;
    for (int g = 0; g < groups; ++g) {
        // operand maps: K-major -> (inner = K, outer = rows); MN-major -> (inner = rows, outer = K)
        if (!a_mn) rc = make_map(&maps.a[g], a[g], K, M, d->sa_m, ba, bke, BM, CU_TENSOR_MAP_SWIZZLE_128B, es);
        else if (a_ch) rc = make_map_chunked(&maps.a[g], a[g], M, K, d->sa_k, ba, BM / ch, es);
        else rc = make_map(&maps.a[g], a[g], M, K, d->sa_k, ba, ch, bke, mn_swz, es);
        if (rc) return rc;
        if (!b_mn) rc = make_map(&maps.b[g], b[g], K, N, d->sb_n, bb, bke, pl.bn / cl, CU_TENSOR_MAP_SWIZZLE_128B, es);
        else if (b_ch) rc = make_map_chunked(&maps.b[g], b[g], N, K, d->sb_k, bb, pl.bn / cl / ch, es);
        else rc = make_map(&maps.b[g], b[g], N, K, d->sb_k, bb, ch, bke, mn_swz, es);
        if (rc) return rc;
        rc = make_map(&maps.c[g], c[g], N, M, d->sc_m, bc, 32, 32);
        if (rc) return rc;
        p.bias[g] = bias ? (const float*)bias[g] : nullptr;
    }
    for (int g = groups; g < LG_MAX_GROUPS; ++g) p.bias[g] = nullptr;
    p.M = (int)M;
    p.N = (int)N;
    p.K = (int)K;
    p.tiles_m = (pl.tiles_m + cl - 1) / cl;
    p.tiles_n = pl.tiles_n;
    p.splits = pl.splits;
    p.kblocks_per_split = pl.kper;
    p.kblocks_total = pl.kblocks;
    p.batch1 = (int)d->batch1;
    p.batches = (int)batches;
    p.groups = n_problems;
    p.epi_op = epi_op;
    p.aux = (const float*)aux;
    p.aux_ld = aux_ld;
    p.aux_sb0 = d->sc_b0;      // a saved matrix of a row epilogue has C's batch layout
    p.aux_sb1 = d->sc_b1;
    p.epi_alpha = (float)alpha;
    p.colsum = nullptr;
    p.sched = pair_mma ? nullptr : sched_counters();
    float* colsum_after = nullptr;
    if (epi_op == 2 && p.bias[0] != nullptr) {
        // LG_EPI_GELU_BWD has no bias to add: the pointer names where the column sums of the result are accumulated --
        // by the epilogue of the wide-tile kernels, by a column reduction after the launch for narrow tiles
        if (pl.bn > 128) p.colsum = const_cast<float*>(p.bias[0]);
        else colsum_after = const_cast<float*>(p.bias[0]);
        p.bias[0] = nullptr;
    }
    if (epi_op == 1) {
        // second result: same geometry as C, its own row pitch
        rc = make_map(&maps.aux, aux, N, M, aux_ld, BatchDims{1, 0, 1, 0}, 32, 32);
        if (rc) return rc;
    } else {
        maps.aux = maps.c[0];
    }
    p.kcat = same_c ? groups : 1;
    p.reduce_out = (pl.splits > 1 || accumulate) ? 1 : 0;
    if (pl.splits > 1 && !accumulate) {
        // split-K partials are summed by TMA reduce-add into a zeroed C (single group here)
        if (d->sc_m == N) {
            LG_CUDA(cudaMemsetAsync(c[0], 0, (size_t)M * N * 4, stream()));
        } else {
            LG_CUDA(cudaMemset2DAsync(c[0], (size_t)d->sc_m * 4, 0, (size_t)N * 4, (size_t)M, stream()));
        }
    }
    const int max_clusters = gemm_sms() / cl > 0 ? gemm_sms() / cl : 1;
    // tail-wave split: tiles % clusters tiles would occupy a mostly idle extra wave; when at least two K-ranges
    // of each fit into that wave, cut them up (partials meet by reduce-add in a zeroed / accumulating C)
    int64_t per_batch = (int64_t)p.tiles_m * pl.tiles_n * pl.splits;
    p.tail_first = (int)per_batch;
    p.tail_splits = 1;
    p.tail_kper = pl.kper;
    static const bool tail_off = getenv("LG_GEMM_NO_TAIL_SPLIT") != nullptr;
    if (!tail_off && pl.splits == 1 && n_problems == 1 && batches == 1 && !epi_op && p.kcat == 1 &&
        per_batch > max_clusters) {
        const int tail = (int)(per_batch % max_clusters);
        int ts = tail > 0 ? max_clusters / tail : 1;
        if (ts > 4) ts = 4;
        while (ts > 1 && pl.kblocks / ts < 16) --ts;          // >= 16 k-blocks (512 of K) per range
        if (ts > 1) {
            p.tail_first = (int)per_batch - tail;
            p.tail_kper = (pl.kblocks + ts - 1) / ts;
            p.tail_splits = (pl.kblocks + p.tail_kper - 1) / p.tail_kper;
            per_batch = p.tail_first + (int64_t)tail * p.tail_splits;
            if (!accumulate) {
                // zero the tile columns the cut tiles live in (whole tiles there overwrite the zeros)
                const int64_t col0 = (int64_t)(p.tail_first / p.tiles_m) * pl.bn;
                LG_CUDA(cudaMemset2DAsync((float*)c[0] + col0, (size_t)d->sc_m * 4, 0, (size_t)(N - col0) * 4, (size_t)M,
                                          stream()));
            }
        }
    }
    p.items_per_batch = (int)per_batch;
    const int64_t items64 = per_batch * batches * n_problems;   // cluster work items
    LG_REQUIRE(items64 < 0x7fffffff, "gemm_tc: too many tiles");
    const int items = (int)items64;
    const int grid = cl * (items < max_clusters ? items : max_clusters);
    rc = es == 4 ? launch_any<4>(pair_mma, pl.bn, a_mn, b_mn, maps, p, grid)
                 : launch_any<2>(pair_mma, pl.bn, a_mn, b_mn, maps, p, grid);
    if (!rc && colsum_after)
        rc = lg_reduce_pitched(LG_RED_SUM, LG_F32, c[0], colsum_after, 1, M, N, d->sc_m, 1.0, 1);
    return rc;
}

int gemm_tc(int mode, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int accumulate) {
    const void* av[1] = {a};
    const void* bv[1] = {b};
    void* cv[1] = {c};
    const void* biasv[1] = {bias};
    return gemm_tc_grouped(mode, d, 1, av, bv, cv, bias ? biasv : nullptr, accumulate, 0, nullptr, 0, 0.0);
}

int gemm_tc_epilogue(int mode, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int epi_op,
                     void* aux, int64_t aux_ld, double alpha) {
    const void* av[1] = {a};
    const void* bv[1] = {b};
    void* cv[1] = {c};
    const void* biasv[1] = {bias};
    return gemm_tc_grouped(mode, d, 1, av, bv, cv, bias ? biasv : nullptr, 0, epi_op, aux, aux_ld, alpha);
}

int gemm_set_sm_limit(int n) {
    g_gemm_sm_limit = n;
    return 0;
}

}  // namespace lg
