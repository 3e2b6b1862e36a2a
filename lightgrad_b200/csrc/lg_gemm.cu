// lg_gemm: mode dispatch between the exact SIMT kernel and the tcgen05 tensor-core kernel.
#include "lg_common.cuh"

namespace lg {
int gemm_simt(int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
              int accumulate);
int gemm_tc(int mode, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias, int accumulate);
int gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, const void* c);
}  // namespace lg

using namespace lg;

extern "C" {

int lg_gemm_tc_supported(int mode, int dtype, const LgGemmDesc* d) {
    return gemm_tc_supported(mode, dtype, d, nullptr, nullptr, nullptr);
}

int lg_gemm(int mode, int dtype, const LgGemmDesc* d, const void* a, const void* b, void* c, const void* bias,
            int accumulate) {
    LG_INIT();
    LG_REQUIRE(d->M >= 0 && d->N >= 0 && d->K >= 0, "lg_gemm: negative dimension");
    LG_REQUIRE(d->batch0 >= 1 && d->batch1 >= 1, "lg_gemm: batch dims must be >= 1");
    if (mode != LG_GEMM_FP32_SIMT && gemm_tc_supported(mode, dtype, d, a, b, c))
        return gemm_tc(mode, d, a, b, c, bias, accumulate);
    return gemm_simt(dtype, d, a, b, c, bias, accumulate);
}

}  // extern "C"
