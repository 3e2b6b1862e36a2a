// Part of the elementwise engine: one-input operator instantiations (split out so the
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

int ew_dispatch1(int opc, int dtype, const void* a, void* out, const EwShape& s, double alpha) {
    using namespace lg::op;
    switch (opc) {
#define C1(code, OP) case code: return by_dtype<OP, 1>(dtype, a, nullptr, nullptr, out, s, alpha);
        C1(LG_EW_COPY, Copy) C1(LG_EW_NEG, Neg) C1(LG_EW_SIN, Sin) C1(LG_EW_COS, Cos) C1(LG_EW_EXP, Exp)
        C1(LG_EW_LOG, Log) C1(LG_EW_SIGMOID, Sigmoid) C1(LG_EW_TANH, Tanh) C1(LG_EW_RELU, Relu)
        C1(LG_EW_GELU, Gelu) C1(LG_EW_ADD_S, AddS) C1(LG_EW_MUL_S, MulS) C1(LG_EW_RSUB_S, RSubS)
        C1(LG_EW_RDIV_S, RDivS) C1(LG_EW_POW_S, PowS) C1(LG_EW_RPOW_S, RPowS) C1(LG_EW_SQRT, Sqrt)
        C1(LG_EW_DIV_S, DivS) C1(LG_EW_FILL, Fill)
#undef C1
    }
    return set_error("unknown one-input elementwise op code %d", opc);
}
}  // namespace lg
