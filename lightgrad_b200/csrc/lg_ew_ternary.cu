// Part of the elementwise engine: three-input operator instantiations (split out so the
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

int ew_dispatch3(int opc, int dtype, const void* a, const void* b, const void* c, void* out, const EwShape& s,
                 double alpha) {
    using namespace lg::op;
    switch (opc) {
#define C3(code, OP) case code: return by_dtype<OP, 3>(dtype, a, b, c, out, s, alpha);
        C3(LG_EW_DIV_BWD_B, DivBwdB) C3(LG_EW_POW_BWD_A, PowBwdA) C3(LG_EW_POW_BWD_B, PowBwdB)
        C3(LG_EW_EQ_MASK_MUL, EqMaskMul)
#undef C3
    }
    return set_error("unknown three-input elementwise op code %d", opc);
}
}  // namespace lg
