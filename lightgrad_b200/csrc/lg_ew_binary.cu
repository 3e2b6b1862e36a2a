// Part of the elementwise engine: two-input operator instantiations (split out so the
// translation units build in parallel).
#include "lg_ew.cuh"
#include "lg_ew_ops.cuh"

namespace lg {
namespace {
template <class Op, int NIN>
int by_dtype(int dtype, const void* a, const void* b, const void* c, void* out, const EwShape& s, double alpha) {
    switch (dtype) {
        case LG_F32: return ew_launch<Op, float, NIN>(a, b, c, out, s, alpha);
        case LG_F64: return ew_launch<Op, double, NIN>(a, b, c, out, s, alpha);
    }
    return set_error("elementwise op: unsupported dtype %d (float32/float64 only)", dtype);
}
}  // namespace

int ew_dispatch2(int opc, int dtype, const void* a, const void* b, void* out, const EwShape& s, double alpha) {
    using namespace lg::op;
    switch (opc) {
#define C2(code, OP) case code: return by_dtype<OP, 2>(dtype, a, b, nullptr, out, s, alpha);
        C2(LG_EW_ADD, Add) C2(LG_EW_SUB, Sub) C2(LG_EW_MUL, Mul) C2(LG_EW_DIV, Div) C2(LG_EW_POW, Pow)
        C2(LG_EW_SIN_BWD, SinBwd) C2(LG_EW_COS_BWD, CosBwd) C2(LG_EW_LOG_BWD, LogBwd)
        C2(LG_EW_SIGMOID_BWD, SigmoidBwd) C2(LG_EW_TANH_BWD, TanhBwd) C2(LG_EW_RELU_BWD, ReluBwd)
        C2(LG_EW_GELU_BWD, GeluBwd) C2(LG_EW_POW_S_BWD, PowSBwd) C2(LG_EW_RPOW_S_BWD, RPowSBwd)
        C2(LG_EW_RDIV_S_BWD, RDivSBwd) C2(LG_EW_AXPY, Axpy)
#undef C2
    }
    return set_error("unknown two-input elementwise op code %d", opc);
}
}  // namespace lg
