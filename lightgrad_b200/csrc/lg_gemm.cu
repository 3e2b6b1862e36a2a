// lg_gemm: mode dispatch between the exact SIMT kernel and the tcgen05 tensor-core kernel.
#include "lg_common.cuh"

namespace lg {
int gemm_simt(int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
              int accumulate);
int gemm_tc(int mode, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int accumulate);
int gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, const void* c);
int gemm_tc_grouped(int mode, const LgGemmDesc* d, int groups, const void* const* a, const void* const* b, void* const* c,
                    const void* const* bias, int accumulate, int epi_op, void* aux, int64_t aux_ld, double alpha);
int gemm_set_sm_limit(int n);
int gemm_tc_epilogue(int mode, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int epi_op, void* aux,
                     int64_t aux_ld, double alpha);
}  // namespace lg

using namespace lg;

#include <vector>
namespace {
// opt-in per-launch timing of the matmul kernels (bench.py's roofline leg): CUDA events on the
// compute stream around every lg_gemm launch while enabled
struct GemmProbe { cudaEvent_t e0, e1; double flops; };
bool g_prof_on = false;
std::vector<GemmProbe> g_probes;
// mode 2: lg_gemm only counts (launches, flops) and launches nothing -- bench.py times a captured step with
// and without its matmuls; the difference is the matmul time inside the replayed step, free of the
// serialisation that event pairs around every launch introduce
bool g_skip = false;
uint64_t g_skip_launches = 0;
double g_skip_flops = 0.0;

// LG_GEMM_BF16_TC multiplies bf16 operands (dtype LG_BF16: staging copies made with lg_cast or written by a producer
// kernel) into an fp32 result; fp32 operands in that mode are an error, never a silent change of arithmetic
int bf16_contract(int mode, int dtype, const char* who) {
    if (mode == LG_GEMM_BF16_TC && dtype != LG_BF16)
        return set_error("%s: LG_GEMM_BF16_TC takes bf16 operands (dtype LG_BF16); stage fp32 operands with lg_cast first", who);
    if (dtype == LG_BF16 && mode != LG_GEMM_BF16_TC)
        return set_error("%s: bf16 operands need mode LG_GEMM_BF16_TC", who);
    return 0;
}
}  // namespace

extern "C" {

int lg_prof_gemm(int enable) {
    LG_INIT();
    g_prof_on = enable == 1;
    g_skip = enable == 2;
    return 0;
}

int lg_prof_gemm_read(double* total_ms, uint64_t* launches, double* total_flops) {
    LG_INIT();
    LG_CUDA(cudaStreamSynchronize(stream()));
    double ms = 0.0, fl = 0.0;
    for (auto& p : g_probes) {
        float t = 0.f;
        LG_CUDA(cudaEventElapsedTime(&t, p.e0, p.e1));
        ms += t;
        fl += p.flops;
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    *total_ms = ms;
    *launches = g_probes.size() + g_skip_launches;
    *total_flops = fl + g_skip_flops;
    g_probes.clear();
    g_skip_launches = 0;
    g_skip_flops = 0.0;
    return 0;
}

int lg_gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d) {
    return gemm_tc_supported(mode, dtype, d, nullptr, nullptr, nullptr);
}

int lg_gemm(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
            int accumulate) {
    LG_INIT();
    LG_REQUIRE(d->M >= 0 && d->N >= 0 && d->K >= 0, "lg_gemm: negative dimension");
    LG_REQUIRE(d->batch0 >= 1 && d->batch1 >= 1, "lg_gemm: batch dims must be >= 1");
    if (g_skip) {
        g_skip_launches += 1;
        g_skip_flops += 2.0 * (double)d->M * (double)d->N * (double)d->K * (double)(d->batch0 * d->batch1);
        return 0;
    }
    GemmProbe pr;
    if (g_prof_on) {
        LG_CUDA(cudaEventCreate(&pr.e0));
        LG_CUDA(cudaEventCreate(&pr.e1));
        pr.flops = 2.0 * (double)d->M * (double)d->N * (double)d->K * (double)(d->batch0 * d->batch1);
        LG_CUDA(cudaEventRecord(pr.e0, stream()));
    }
    int rc;
    if (bf16_contract(mode, dtype, "lg_gemm")) return 1;
    if (mode != LG_GEMM_FP32_SIMT && !(accumulate && bias) && gemm_tc_supported(mode, dtype, d, a, b, c))
        rc = gemm_tc(mode, d, a, b, c, bias, accumulate);
    else if (dtype == LG_BF16)
        rc = set_error("lg_gemm: this problem cannot run on the bf16 tensor-core path (check lg_gemm_tc_supported "
                       "first and use the exact mode on the fp32 operands instead)");
    else
        rc = gemm_simt(dtype, d, a, b, c, bias, accumulate);
    if (g_prof_on) {
        LG_CUDA(cudaEventRecord(pr.e1, stream()));
        g_probes.push_back(pr);
    }
    return rc;
}

int lg_gemm_grouped(int mode, int dtype, const LgGemmDesc* d, int groups, const void* const* a, const void* const* b,
                    void* const* c, const void* const* bias, int accumulate) {
    LG_INIT();
    LG_REQUIRE(groups >= 1 && groups <= 4, "lg_gemm_grouped: 1..4 groups");
    if (g_skip) {
        g_skip_launches += 1;
        g_skip_flops += 2.0 * (double)d->M * (double)d->N * (double)d->K * (double)(d->batch0 * d->batch1) * groups;
        return 0;
    }
    if (bf16_contract(mode, dtype, "lg_gemm_grouped")) return 1;
    bool tc = mode != LG_GEMM_FP32_SIMT && !(accumulate && bias);
    for (int g = 0; g < groups && tc; ++g) tc = gemm_tc_supported(mode, dtype, d, a[g], b[g], c[g]) != 0;
    LG_REQUIRE(tc || dtype != LG_BF16, "lg_gemm_grouped: this problem cannot run on the bf16 tensor-core path");
    GemmProbe pr;
    if (g_prof_on) {
        LG_CUDA(cudaEventCreate(&pr.e0));
        LG_CUDA(cudaEventCreate(&pr.e1));
        pr.flops = 2.0 * (double)d->M * (double)d->N * (double)d->K * (double)(d->batch0 * d->batch1) * groups;
        LG_CUDA(cudaEventRecord(pr.e0, stream()));
    }
    int rc = 0;
    if (tc) {
        rc = gemm_tc_grouped(mode, d, groups, a, b, c, bias, accumulate, 0, nullptr, 0, 0.0);
    } else {
        // exact path: one launch per group (a repeated C accumulates from the second group on)
        for (int g = 0; g < groups && !rc; ++g) {
            bool seen = false;
            for (int h = 0; h < g; ++h) seen = seen || c[h] == c[g];
            rc = gemm_simt(dtype, d, a[g], b[g], c[g], (bias && !seen) ? bias[g] : nullptr, (accumulate || seen) ? 1 : 0);
        }
    }
    if (g_prof_on) {
        LG_CUDA(cudaEventRecord(pr.e1, stream()));
        g_probes.push_back(pr);
    }
    return rc;
}

int lg_gemm_epilogue(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
                     int epi_op, void* aux, int64_t aux_ld, double alpha) {
    LG_INIT();
    LG_REQUIRE(epi_op >= LG_EPI_GELU_FWD && epi_op <= LG_EPI_SOFTMAX_BWD, "lg_gemm_epilogue: unknown epilogue %d", epi_op);
    const bool rows = epi_op >= LG_EPI_SOFTMAX_FWD;
    LG_REQUIRE(d->sc_n == 1, "lg_gemm_epilogue: row-major result");
    LG_REQUIRE(rows || (d->batch0 == 1 && d->batch1 == 1), "lg_gemm_epilogue: the GELU epilogues take one plain product");
    LG_REQUIRE(epi_op == LG_EPI_SOFTMAX_FWD || (aux && aux_ld >= d->N), "lg_gemm_epilogue: aux operand missing");
    const double flops = 2.0 * (double)d->M * (double)d->N * (double)d->K * (double)(d->batch0 * d->batch1);
    if (g_skip) {
        g_skip_launches += 1;
        g_skip_flops += flops;
        return 0;
    }
    if (bf16_contract(mode, dtype, "lg_gemm_epilogue")) return 1;
    bool tc = mode != LG_GEMM_FP32_SIMT && gemm_tc_supported(mode, dtype, d, a, b, c) && d->N % 4 == 0 &&
              (!aux || ((((uintptr_t)aux) & 15) == 0 && aux_ld % 4 == 0));
    LG_REQUIRE(tc || dtype != LG_BF16, "lg_gemm_epilogue: this problem cannot run on the bf16 tensor-core path");
    if (rows) {
        // the row epilogues exist only on the tensor-core path (a row must fit one 128-column tile); callers
        // check lg_gemm_tc_supported first and otherwise run the product and the softmax kernel separately
        LG_REQUIRE(tc && d->N <= 128 && !bias, "lg_gemm_epilogue: row epilogue needs the tensor-core path, N <= 128, no bias");
    }
    GemmProbe pr;
    if (g_prof_on) {
        LG_CUDA(cudaEventCreate(&pr.e0));
        LG_CUDA(cudaEventCreate(&pr.e1));
        pr.flops = flops;
        LG_CUDA(cudaEventRecord(pr.e0, stream()));
    }
    int rc;
    if (tc) {
        rc = gemm_tc_epilogue(mode, d, a, b, c, bias, epi_op, aux, aux_ld, alpha);
    } else {
        // exact path: the product, then the activation as a strided elementwise pass over the same buffers
        rc = gemm_simt(dtype, d, a, b, c, epi_op == LG_EPI_GELU_BWD ? nullptr : bias, 0);
        const int64_t shape[2] = {d->M, d->N}, sc[2] = {d->sc_m, 1}, sx[2] = {aux_ld, 1};
        if (!rc && epi_op == LG_EPI_GELU_FWD)
            rc = lg_ew(LG_EW_GELU, dtype, 2, shape, c, sc, nullptr, nullptr, nullptr, nullptr, aux, sx, 0.0);
        else if (!rc) {
            rc = lg_ew(LG_EW_GELU_BWD, dtype, 2, shape, aux, sx, c, sc, nullptr, nullptr, c, sc, 0.0);
            // (bias names the accumulation target of the result's column sums for this epilogue, see the header)
            if (!rc && bias) rc = lg_reduce_pitched(LG_RED_SUM, dtype, c, const_cast<void*>(bias), 1, d->M, d->N, d->sc_m, 1.0, 1);
        }
    }
    if (g_prof_on) {
        LG_CUDA(cudaEventRecord(pr.e1, stream()));
        g_probes.push_back(pr);
    }
    return rc;
}

int lg_gemm_sm_limit(int n_sms) {
    LG_INIT();
    LG_REQUIRE(n_sms >= 0, "lg_gemm_sm_limit: negative SM count");
    return gemm_set_sm_limit(n_sms);
}

}  // extern "C"
