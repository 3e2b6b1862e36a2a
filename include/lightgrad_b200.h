/*
 * lightgrad_b200 -- C-ABI of the sm_100a tensor backend.
 *
 * This is the drop-in boundary: everything lightgrad's Python backend layer
 * needs from a device (what the reference's OpenCL backend gets from pyopencl
 * + its JIT-built kernels) is one of the plain-C entry points below.  Host code
 * (lightgrad_b200/autograd/cuda/*.py) reaches them with ctypes; nothing here
 * mentions torch, numpy or Python.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from lg_last_error() (thread-local, valid until the next call);
 *   - device pointers are void* obtained from lg_alloc (or any cudaMalloc);
 *   - sizes, shapes and strides are int64_t, strides are in ELEMENTS;
 *   - all work is enqueued on the library's compute stream and the call returns
 *     immediately; only lg_sync, lg_memcpy_d2h and lg_event_sync block;
 *   - not thread-safe by contract (the reference is single-threaded), one
 *     device per process (data parallelism = one process per GPU).
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   runtime      lightgrad/autograd/opencl/device.py:51-115 (context, queue, MemoryPool)
 *                lightgrad/autograd/opencl/tensor.py:58-93  (allocate, enqueue_copy)
 *   elementwise  lightgrad/autograd/opencl/kernels.py:24-195 (atom)  /  cpu/ops.py:52-229
 *   reductions   lightgrad/autograd/opencl/kernels.py:344-501 (reduce) / cpu/ops.py:260-293
 *   matmul       lightgrad/autograd/opencl/kernels.py:201-337 (matmul) / cpu/ops.py:107-116
 *   indexing     lightgrad/autograd/cpu/ops.py:234-255 (getitem / setitem)
 *   fused        lightgrad/autograd/ops.py:62-66 (softmax), lightgrad/nn.py:109-124 (LayerNorm),
 *                lightgrad/loss.py:14-24 (cross_entropy), examples/bert.py:12 (gelu)
 *   optimizers   lightgrad/optim.py:3-52
 *   collectives  (none in the reference -- new: NCCL gradient all-reduce)
 */
#ifndef LIGHTGRAD_B200_H
#define LIGHTGRAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_MAX_DIMS 8

/* element types */
enum LgDType {
    LG_F32 = 0,
    LG_F64 = 1,
    LG_I32 = 2,
    LG_I64 = 3,
    LG_I16 = 4,
    LG_U8  = 5,
    LG_I8  = 6,
    LG_BF16 = 7   /* internal: operand staging for the BF16 tensor-core GEMM */
};

/* elementwise operator codes (out = f(a[,b[,c]]; alpha)) */
enum LgEwOp {
    /* one input */
    LG_EW_COPY = 0, LG_EW_NEG = 1, LG_EW_SIN = 2, LG_EW_COS = 3, LG_EW_EXP = 4, LG_EW_LOG = 5,
    LG_EW_SIGMOID = 6, LG_EW_TANH = 7, LG_EW_RELU = 8, LG_EW_GELU = 9,
    LG_EW_ADD_S = 10,   /* a + alpha          */
    LG_EW_MUL_S = 11,   /* a * alpha          */
    LG_EW_RSUB_S = 12,  /* alpha - a          */
    LG_EW_RDIV_S = 13,  /* alpha / a          */
    LG_EW_POW_S = 14,   /* a ** alpha         */
    LG_EW_RPOW_S = 15,  /* alpha ** a         */
    LG_EW_SQRT = 16,
    LG_EW_DIV_S = 17,   /* a / alpha          */
    LG_EW_FILL = 18,    /* alpha (a ignored)  */
    /* two inputs */
    LG_EW_ADD = 32, LG_EW_SUB = 33, LG_EW_MUL = 34, LG_EW_DIV = 35, LG_EW_POW = 36,
    LG_EW_SIN_BWD = 40,      /* cos(a) * b                         (a = x, b = g)   */
    LG_EW_COS_BWD = 41,      /* -sin(a) * b                                         */
    LG_EW_LOG_BWD = 42,      /* (1 / a) * b                                         */
    LG_EW_SIGMOID_BWD = 43,  /* a * (1 - a) * b                    (a = y)          */
    LG_EW_TANH_BWD = 44,     /* (1 - a*a) * b                      (a = y)          */
    LG_EW_RELU_BWD = 45,     /* b * (a >= 0)                                        */
    LG_EW_GELU_BWD = 46,     /* gelu'(a) * b                                        */
    LG_EW_POW_S_BWD = 47,    /* alpha * a**(alpha-1) * b                            */
    LG_EW_RPOW_S_BWD = 48,   /* b * a * ln(alpha)                  (a = y)          */
    LG_EW_RDIV_S_BWD = 49,   /* -alpha / a**2 * b                                   */
    LG_EW_AXPY = 50,         /* a + alpha * b                                       */
    /* three inputs */
    LG_EW_DIV_BWD_B = 64,    /* -a / b**2 * c                      (c = g)          */
    LG_EW_POW_BWD_A = 65,    /* b * a**(b-1) * c                                    */
    LG_EW_POW_BWD_B = 66,    /* c * b * ln(a)                      (b = y)          */
    LG_EW_EQ_MASK_MUL = 67   /* c * (a == b)                       (max/min bwd)    */
};

enum LgReduceOp { LG_RED_SUM = 0, LG_RED_MAX = 1, LG_RED_MIN = 2 };

/* matmul arithmetic modes */
enum LgGemmMode {
    LG_GEMM_FP32_SIMT = 0,  /* exact fp32 FFMA (also the only mode for f64)       */
    LG_GEMM_TF32_TC = 1,    /* tcgen05 kind::tf32, fp32 operands read by TMA       */
    LG_GEMM_BF16_TC = 2     /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate  */
};

/* ---- runtime (replaces opencl/device.py + pyopencl.tools.MemoryPool) ---------------------- */
const char* lg_last_error(void);
int lg_device_count(int* count);
int lg_init(int device);                 /* idempotent; binds the process to one GPU       */
int lg_device(int* device);
int lg_device_props(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
int lg_sync(void);                        /* drain compute + comm streams                   */
/* (with a NCCL communicator alive, lg_sync and lg_memcpy_d2h watch their wait: an asynchronous NCCL error, or no
 *  completion within LG_SYNC_TIMEOUT_S seconds -- default 120, 0 = wait for ever -- aborts the communicator and fails
 *  the call instead of hanging the rank on a collective whose peer is gone) */
int lg_alloc(size_t nbytes, void** ptr);  /* caching, stream-ordered; 256-byte aligned      */
int lg_free(void* ptr);                   /* returns the block to the cache                 */
int lg_empty_cache(void);
int lg_mem_stats(size_t* in_use, size_t* reserved, size_t* peak_in_use);
int lg_memcpy_h2d(void* dst, const void* src, size_t nbytes);   /* async wrt host when src is pinned */
int lg_memcpy_d2h(void* dst, const void* src, size_t nbytes);   /* blocks until the bytes landed      */
int lg_memcpy_d2d(void* dst, const void* src, size_t nbytes);
int lg_memset(void* dst, int byte, size_t nbytes);
int lg_host_alloc(size_t nbytes, void** ptr);                   /* pinned host memory */
int lg_host_free(void* ptr);
int lg_event_create(void** ev);
int lg_event_record(void* ev);            /* on the compute stream */
int lg_event_sync(void* ev);
int lg_event_elapsed_ms(void* start, void* stop, float* ms);
int lg_event_destroy(void* ev);
int lg_launch_count(uint64_t* n);         /* kernels launched by this library so far */
void* lg_stream_handle(void);             /* cudaStream_t of the compute stream, for profilers */
int lg_profiler_range(int start);
/* keeps the compute stream busy for `us` microseconds (measurement aid: work queued behind it runs back
 * to back on the device, so events between kernels are free of host dispatch latency) */
int lg_stream_delay_us(uint64_t us);
/* Side stream.  Launches issued between lg_side_begin and lg_side_end go to a second stream that is ordered
 * after everything issued on the compute stream so far and runs concurrently with what follows (weight / bias
 * gradients, which nothing else in backward waits for, backfill the SMs a one-wave GEMM leaves idle; the
 * reference has a single in-order OpenCL queue, opencl/device.py:58-60).  lg_side_join makes the compute stream
 * wait for them; lg_sync, lg_memcpy_d2h, lg_graph_begin/end and lg_nccl_fork join / order implicitly.  The caller
 * keeps every buffer those launches touch alive until the join. */
/* lg_comm_compute_begin .. end: launches in between go to the COLLECTIVE stream, i.e. right behind the all-reduce
 * queued there (per-bucket optimizer updates that must not stall backward on the compute stream); lg_nccl_wait
 * orders the compute stream after them. */
int lg_comm_compute_begin(void);
int lg_comm_compute_end(void);
int lg_side_begin(void);
int lg_side_end(void);
int lg_side_join(void);
/* whole-step CUDA graphs: everything enqueued between begin and end (kernels, memsets, NCCL calls)
 * is recorded instead of executed; allocations made meanwhile come from a pool private to the
 * graph (*pool_id: 0 = create one, reused by later re-captures).  lg_graph_launch replays the step
 * with no host-side dispatch.  Host copies and synchronisation are rejected during capture. */
int lg_graph_begin(int* pool_id);
int lg_graph_end(void** graph_exec, uint64_t* n_nodes);
int lg_graph_abort(void);
int lg_graph_launch(void* graph_exec, uint64_t n_kernels /* added to lg_launch_count */);
int lg_graph_destroy(void* graph_exec);         /* cudaProfilerStart/Stop (ncu --profile-from-start off) */

/* ---- elementwise (replaces kernels.atom) --------------------------------------------------- */
/* all operands contiguous, n elements, same dtype; b/c may be NULL for 1/2-input ops */
int lg_ew_flat(int op, int dtype, const void* a, const void* b, const void* c, void* out,
               int64_t n, double alpha);
/* broadcast / strided form: every operand described over one common shape; a stride of 0
 * broadcasts; s? == NULL means "contiguous over shape"; out may alias an input (in-place ops) */
int lg_ew(int op, int dtype, int ndim, const int64_t* shape,
          const void* a, const int64_t* sa, const void* b, const int64_t* sb,
          const void* c, const int64_t* sc, void* out, const int64_t* so, double alpha);
/* fused two-output backward of mul/div (kernels.atom with two outputs, opencl/ops.py:78-98):
 * contiguous operands; kind 0: da = g*b, db = a*g;  kind 1: da = g/b, db = -a/b^2*g */
int lg_ew_bwd2_flat(int kind, int dtype, const void* a, const void* b, const void* g,
                    void* da, void* db, int64_t n);
/* dtype conversion / strided gather-copy (replaces OpenCLTensor.contiguous + astype) */
int lg_cast(int src_dtype, int dst_dtype, int ndim, const int64_t* shape,
            const void* src, const int64_t* ssrc, void* dst, const int64_t* sdst);

/* ---- reductions (replaces kernels.reduce) -------------------------------------------------- */
/* x is a contiguous (outer, reduce, inner) block; out is (outer, inner); out = scale * reduce(x) */
int lg_reduce(int op, int dtype, const void* x, void* out,
              int64_t outer, int64_t reduce, int64_t inner, double scale);
/* same with rows `ld` elements apart (ld >= inner > 1): element (o, r, c) at x[(o*reduce + r)*ld + c]
 * (lets a column reduction read a GEMM result whose leading dimension was padded for TMA);
 * accumulate != 0 (sums only): out += scale * sum, e.g. a bias gradient added straight into .grad */
int lg_reduce_pitched(int op, int dtype, const void* x, void* out,
                      int64_t outer, int64_t reduce, int64_t inner, int64_t ld, double scale, int accumulate);

/* ---- matmul (replaces kernels.dot) ---------------------------------------------------------- */
/* C[b0,b1] (M x N) = A[b0,b1] (M x K) * B[b0,b1] (K x N) (+ bias[N] if bias != NULL)
 * element strides: A(m,k) at a + b0*sa_b0 + b1*sa_b1 + m*sa_m + k*sa_k, likewise B(k,n), C(m,n).
 * Transposed operands are expressed through the strides -- no copies are made.
 * accumulate != 0: C += A*B. */
typedef struct LgGemmDesc {
    int64_t M, N, K;
    int64_t batch0, batch1;
    int64_t sa_b0, sa_b1, sa_m, sa_k;
    int64_t sb_b0, sb_b1, sb_k, sb_n;
    int64_t sc_b0, sc_b1, sc_m, sc_n;
} LgGemmDesc;
int lg_gemm(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c,
            const void* bias, int accumulate);
/* `groups` (1..4) problems that share one LgGemmDesc but have their own operand / result / bias pointers, as one
 * launch (the Q, K and V projections of an attention block; their three dW).  Several groups may name the same
 * C: with accumulate != 0 every group adds into it (dX = sum_g dY_g W_g). */
int lg_gemm_grouped(int mode, int dtype, const LgGemmDesc* d, int groups, const void* const* a, const void* const* b,
                    void* const* c, const void* const* bias, int accumulate);
/* A product with the operation that follows it fused into the GEMM epilogue.
 * One plain (unbatched) product -- the two halves of gelu(x W1^T + b1) in BertLayer (examples/bert.py:12,150-153
 * of the reference):
 *   LG_EPI_GELU_FWD: c = a b + bias (kept for backward) and aux = gelu(c), both written by the epilogue;
 *   LG_EPI_GELU_BWD: c = (a b) * gelu'(aux), aux = the pre-activation saved by the forward call; `bias`, if not
 *                    NULL, is not added: bias[n] += sum over rows of c[:, n] (c is the gradient of the pre-activation,
 *                    so this is the gradient of the first Linear's bias, accumulated by the epilogue instead of a
 *                    second pass over c).
 * Batched, N <= 128, tensor-core mode only (check lg_gemm_tc_supported) -- attention of BertSelfAttention
 * (examples/bert.py:78-88):
 *   LG_EPI_SOFTMAX_FWD: c = softmax(alpha * (a b)) along each row            (scores -> probabilities);
 *   LG_EPI_SOFTMAX_BWD: c = alpha * aux * (a b - sum_row(aux * (a b))), aux = the probabilities (c's layout).
 * aux has c's shape, row pitch aux_ld elements. */
typedef enum {
    LG_EPI_NONE = 0, LG_EPI_GELU_FWD = 1, LG_EPI_GELU_BWD = 2, LG_EPI_SOFTMAX_FWD = 3, LG_EPI_SOFTMAX_BWD = 4
} LgGemmEpilogue;
int lg_gemm_epilogue(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
                     int epi_op, void* aux, int64_t aux_ld, double alpha);
/* Persistent tensor-core GEMM grids use at most n_sms SMs (0 = all).  The data-parallel wrapper lowers it while
 * gradient all-reduces overlap backward, so the collective's CTAs do not push a one-CTA-per-SM grid into a
 * second wave. */
int lg_gemm_sm_limit(int n_sms);
/* measurement hooks for the matmul share of a step (off by default):
 * lg_prof_gemm(1): CUDA events on the compute stream around every lg_gemm launch;
 * lg_prof_gemm(2): lg_gemm launches NOTHING and only counts -- timing a captured step with and without its
 *                  matmuls gives their cost inside the replayed step (results are garbage in this mode);
 * lg_prof_gemm_read drains the stream and returns the summed event time (mode 1), the number of launches
 * and their algorithmic flops (2*M*N*K*batch) since the last read */
int lg_prof_gemm(int enable);
int lg_prof_gemm_read(double* total_ms, uint64_t* launches, double* total_flops);
/* 1 if the tensor-core kernel can take this problem in `mode` without a fallback */
int lg_gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d);

/* ---- convolution lowering (replaces the window view + python fold loop of cpu/ops.py:298-355) ------------------
 * n window dims (1..4) over the trailing dims of a contiguous (lead, in_0..in_{n-1}) block; kernel extents k_d,
 * strides s_d, positions pos_d = (in_d - k_d) / s_d + 1.  `cols` is contiguous (lead, pos_0..pos_{n-1}, k_0..k_{n-1}).
 *   lg_im2col: cols[l, p, k] = x[l, p*s + k]                     (unfold: the matmul operand of conv forward)
 *   lg_col2im: dx[l, i] = sum_{p*s + k = i} cols[l, p, k]        (fold: input gradient of conv backward; a gather) */
int lg_im2col(int dtype, int n, int64_t lead, const int64_t* in_dims, const int64_t* k_dims, const int64_t* strides,
              const void* x, void* cols);
int lg_col2im(int dtype, int n, int64_t lead, const int64_t* in_dims, const int64_t* k_dims, const int64_t* strides,
              const void* cols, void* dx);

/* ---- indexing (replaces cpu getitem/setitem with integer-array indices) -------------------- */
/* rows: out[i, :] = src[idx[i], :]; src rows `row_stride` elements apart, row_len contiguous */
int lg_gather_rows(int dtype, int idx_dtype, const void* src, int64_t n_src_rows, int64_t row_stride,
                   const void* idx, int64_t n_idx, int64_t row_len, void* out);
/* dst[idx[i], :] += src[i, :]  (scatter-add; duplicates accumulate) */
int lg_scatter_add_rows(int dtype, int idx_dtype, void* dst, int64_t n_dst_rows, int64_t row_stride,
                        const void* idx, int64_t n_idx, int64_t row_len, const void* src);
/* dst[idx[i], :] = src[i, :] or = value when src == NULL  (setitem; last writer wins) */
int lg_scatter_set_rows(int dtype, int idx_dtype, void* dst, int64_t n_dst_rows, int64_t row_stride,
                        const void* idx, int64_t n_idx, int64_t row_len, const void* src, double value);
/* lin[i] = sum_k wrap(idx_k[i]) * stride_k : folds up to 4 index arrays into row numbers */
int lg_index_linearize(int n_arrays, const void* const* idx, const int* idx_dtypes,
                       const int64_t* dim_sizes, const int64_t* dim_strides, int64_t n, int64_t* lin);

/* ---- fused layers ---------------------------------------------------------------------------- */
/* softmax over the last axis of a contiguous (rows, cols) block; x is pre-multiplied by `scale` */
int lg_softmax_fwd(int dtype, const void* x, void* y, int64_t rows, int64_t cols, double scale);
/* dx = scale * y * (g - sum(g*y)) */
int lg_softmax_bwd(int dtype, const void* y, const void* g, void* dx, int64_t rows, int64_t cols, double scale);
/* loss_rows[i] = logsumexp(x[i,:]) - x[i,label[i]];  lse[i] saved for backward.
 * `ld` = elements between consecutive rows of logits (>= cols; rows may be padded for TMA alignment) */
int lg_cross_entropy_fwd(int dtype, int idx_dtype, const void* logits, int64_t ld, const void* labels,
                         void* loss_rows, void* lse, int64_t rows, int64_t cols);
/* dlogits[i,j] = (exp(x[i,j]-lse[i]) - [j==label[i]]) / rows * gscale[0] ; may run in place on logits */
int lg_cross_entropy_bwd(int dtype, int idx_dtype, const void* logits, int64_t ld, const void* labels, const void* lse,
                         const void* gscale, void* dlogits, int64_t ld_out, int64_t rows, int64_t cols);
/* layer norm over the last axis (nn.py:109-124): y = (x-mean)/sqrt(var+eps)*gamma+beta */
int lg_layernorm_fwd(int dtype, const void* x, const void* gamma, const void* beta, void* y,
                     void* mean, void* rstd, int64_t rows, int64_t cols, double eps);
/* accumulate != 0: dgamma / dbeta are added to the buffers (existing .grad) instead of overwriting them */
/* y = layernorm(a + b): the residual add of BertAttention / BertLayer (examples/bert.py:113,158 of the reference)
 * folded into the normalisation; the sum is written to sum_out (it is the x that lg_layernorm_bwd needs). */
int lg_add_layernorm_fwd(int dtype, const void* a, const void* b, void* sum_out, const void* gamma, const void* beta,
                         void* y, void* mean, void* rstd, int64_t rows, int64_t cols, double eps);
/* dx_colsum (cols values, or NULL; float32 rows of <= 1024 elements): overwritten with the column sums of dx -- when x
 * is the output of a Linear layer (LayerNorm(dense(h) + skip)) they are that layer's bias gradient, formed here instead
 * of by a pass that reads dx back */
int lg_layernorm_bwd(int dtype, const void* x, const void* gamma, const void* mean, const void* rstd,
                     const void* g, void* dx, void* dgamma, void* dbeta, int64_t rows, int64_t cols,
                     int accumulate, void* dx_colsum);

/* Multi-head self-attention core of BertSelfAttention (examples/bert.py:68-88 of the reference) as ONE kernel per
 * direction: per (batch, head) tile  out = softmax(scale * Q K^T) V  with the scores and probabilities kept on chip
 * (tcgen05 tf32 products, TMEM accumulators, softmax in registers).
 *   qkv   (3, batch*seq, heads*head_dim) contiguous: the stacked Q, K, V projections; head h of a token is the
 *         head_dim values at column h*head_dim (the reference's reshape(b, s, h, d).transpose(0, 2, 1, 3) as a view)
 *   out   (batch*seq, heads*head_dim): heads merged back, the reference's context.transpose(0, 2, 1, 3).reshape
 *   lse   (batch*heads*seq): scale * rowmax + ln(rowsum) per query, saved for backward
 *   lg_attention_bwd recomputes the probabilities from qkv and lse and writes dQ, dK, dV in qkv's layout; dbq / dbk /
 *         dbv (heads*head_dim each, or NULL): the column sums of dQ / dK / dV are ADDED to them -- the bias gradients
 *         of the three projections, taken from the registers that hold the rows.
 * lg_attention_supported: 1 when the fused kernels take the shape (float32, seq 128, head_dim 64); other shapes use
 * the batched lg_gemm / lg_gemm_epilogue path. */
int lg_attention_supported(int dtype, int64_t seq, int64_t head_dim);
int lg_attention_fwd(int dtype, const void* qkv, int64_t batch, int64_t seq, int64_t heads, int64_t head_dim,
                     double scale, void* out, void* lse);
int lg_attention_bwd(int dtype, const void* qkv, const void* out, const void* dout, const void* lse, int64_t batch,
                     int64_t seq, int64_t heads, int64_t head_dim, double scale, void* dqkv, void* dbq, void* dbk,
                     void* dbv);

/* ---- optimizers (replaces the per-parameter python loops of optim.py) ----------------------- */
/* all tensors of one optimizer live in flat fp32 arenas; `seg_end_dev[i]` (device, int64) is the
 * exclusive end offset of tensor i.  Tensor i uses step count t = *t_dev + i + 1 in its bias
 * corrections 1 - beta^t (the reference advances t once per parameter, optim.py:36-37); the counter
 * lives on the device so that a captured step stays correct when replayed.
 * A call may cover a sub-range of the arena (per-bucket updates): the pointers address its first element,
 * seg_base is that element's offset in the arena (seg_end_dev entries are arena offsets), its tensors are
 * number seg_offset .. seg_offset + n_seg - 1, and the counter is advanced by t_advance (0 for all but one call
 * of a step; the whole-arena call passes 0, 0, n_seg). */
int lg_sgd_step(void* param, const void* grad, void* delta, int64_t n, double lr, double momentum);
int lg_adam_step(int belief, void* param, const void* grad, void* m, void* v, int64_t n,
                 int n_seg, const int64_t* seg_end_dev, int64_t* t_dev, double lr, double beta1, double beta2,
                 double eps, int64_t seg_base, int seg_offset, int t_advance);

/* ---- collectives (new; NCCL over NVLink, one process per GPU) ------------------------------- */
int lg_nccl_unique_id(void* id128);                       /* 128-byte ncclUniqueId */
int lg_nccl_init(const void* id128, int world, int rank);
int lg_nccl_allreduce_f32(void* buf, int64_t n, int op /* 0 sum, 1 avg, 2 max */, int on_comm_stream);
int lg_nccl_broadcast(void* buf, int64_t nbytes, int root);
/* stream ordering of the collective stream (usable without a communicator: the multicast exchange needs it too) */
int lg_nccl_wait(void);                                   /* compute stream waits for comm stream */
int lg_nccl_fork(void);                                   /* comm stream waits for compute stream */
int lg_nccl_destroy(void);

/* ---- gradient exchange fused with the optimizer over NVLink multicast (new; NVSwitch in-fabric reduction) --------
 * Replaces "lg_nccl_allreduce_f32 over the gradient arena + lg_adam_step over all parameters" of the data-parallel
 * step (the optimizer arithmetic is lightgrad/optim.py:27-52 of the reference, as in lg_adam_step / lg_sgd_step).
 * Both arenas live in one region per GPU that is bound to a multicast object shared by all ranks:
 *   lg_mc_supported      *yes = 1 when the device and driver offer multicast objects
 *   lg_mc_region_bytes   region size and the byte offsets of the gradient arena, the parameter arena and the flag
 *                        page inside it, for arenas of `arena_bytes` each
 *   lg_mc_create         rank 0: create the object; *fd is a POSIX descriptor to pass to the other ranks (SCM_RIGHTS)
 *   lg_mc_import         other ranks: open the object from the received descriptor
 *   lg_mc_add_device     every rank; ALL ranks must have returned (host barrier) before the first lg_mc_bind
 *   lg_mc_bind           every rank: allocate, map and bind this GPU's memory; *local_ptr / *mc_ptr are the plain and
 *                        the multicast mapping of the same region (zero-filled); host barrier before first use
 *   lg_mc_exchange_step  one bucket [lo, hi) (elements, multiples of 4) on the collective stream, ordered by
 *                        lg_nccl_fork / lg_nccl_wait like an all-reduce: reduce-scatter of the gradients through the
 *                        switch (multimem.ld_reduce), optimizer update of this rank's 1/world share (kind 0 Adam,
 *                        1 AdaBelief -- arguments as lg_adam_step, m / v full-size arrays of which only the owned
 *                        ranges are used; 2 SGD with m = previous deltas when momentum != 0; 3 no update), all-gather
 *                        of the new parameters (multimem.st) -- one kernel whose CTAs fit next to a GEMM CTA
 *   lg_mc_release        unmap / unbind / free (after a host barrier) */
int lg_mc_supported(int* yes);
int lg_mc_region_bytes(size_t arena_bytes, int world, size_t* region_bytes, size_t* grad_offset, size_t* param_offset,
                       size_t* flag_offset);
int lg_mc_create(size_t region_bytes, int world, int* fd);
int lg_mc_import(int fd, size_t region_bytes, int world);
int lg_mc_add_device(void);
int lg_mc_bind(void** local_ptr, void** mc_ptr);
int lg_mc_exchange_step(int kind, size_t grad_offset, size_t param_offset, size_t flag_offset, int64_t lo, int64_t hi,
                        int rank, int world, void* m, void* v, int n_seg, const int64_t* seg_end_dev, int64_t* t_dev,
                        double lr, double beta1, double beta2, double eps, double momentum, int seg_offset,
                        int t_advance);
int lg_mc_release(void);
/* One GPU: the optimizer update (lg_adam_step / lg_sgd_step arithmetic; kind 0 Adam, 1 AdaBelief, 2 SGD) of arena
 * elements [lo, hi) by the exchange kernel's single-GPU variant -- 128 threads, no shared memory, <= 88 registers per CTA --
 * on the collective stream (order it with lg_nccl_fork / lg_nccl_wait): it runs beside the GEMMs of the rest of backward
 * as soon as a bucket's gradients are final, so the optimizer leaves the critical path.  param / grad / m / v address the
 * start of the arenas. */
int lg_bucket_step(int kind, void* param, const void* grad, void* m, void* v, int64_t lo, int64_t hi, int n_seg,
                   const int64_t* seg_end_dev, int64_t* t_dev, double lr, double beta1, double beta2, double eps,
                   double momentum, int seg_offset, int t_advance);
/* measurement aid (LG_MC_TRACE=1 in the environment): lg_mc_trace_mark stamps the position of the current stream;
 * lg_mc_trace_read drains the device and returns, in launch order, records of 4 x uint64 {entered, all ranks met,
 * finished, bucket bytes} per exchange launch (GPU global timer, ns) and {t, 0, 0, 0} per mark.  Captured launches keep
 * their record, so after a graph replay the trace is that replay's timeline; reset != 0 clears it. */
int lg_mc_trace_mark(void);
int lg_mc_trace_read(uint64_t* out, int max_records, int* n_records, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTGRAD_B200_H */
