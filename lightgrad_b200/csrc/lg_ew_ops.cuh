// Operator functors for the elementwise engine.  Arithmetic follows the reference CPU tensor
// (lightgrad/autograd/cpu/ops.py) operation by operation so results agree to <= 1e-6 relative:
// same association order, IEEE division, libm-grade transcendentals (no fast-math).
#pragma once
#include <math.h>

namespace lg {
namespace op {

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T t_sin(T x);
template <> __device__ __forceinline__ float t_sin(float x) { return sinf(x); }
template <> __device__ __forceinline__ double t_sin(double x) { return sin(x); }
template <typename T> __device__ __forceinline__ T t_cos(T x);
template <> __device__ __forceinline__ float t_cos(float x) { return cosf(x); }
template <> __device__ __forceinline__ double t_cos(double x) { return cos(x); }
template <typename T> __device__ __forceinline__ T t_tanh(T x);
template <> __device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double t_tanh(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ T t_pow(T x, T y);
template <> __device__ __forceinline__ float t_pow(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double t_pow(double x, double y) { return pow(x, y); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }

#define LG_OP(NAME, EXPR)                                                          \
    struct NAME {                                                                  \
        template <typename T>                                                      \
        static __device__ __forceinline__ T apply(T a, T b, T c, T alpha) {        \
            (void)b; (void)c; (void)alpha;                                         \
            return EXPR;                                                           \
        }                                                                          \
    };

// ---- one input (cpu/ops.py:52-58, 158-229) ----
LG_OP(Copy, a)
LG_OP(Neg, -a)
LG_OP(Sin, t_sin(a))
LG_OP(Cos, t_cos(a))
LG_OP(Exp, t_exp(a))
LG_OP(Log, t_log(a))
LG_OP(Sigmoid, T(1) / (T(1) + t_exp(-a)))
LG_OP(Tanh, t_tanh(a))
LG_OP(Relu, (a < T(0)) ? T(0) : a)  // NaN stays NaN like np.maximum
LG_OP(AddS, a + alpha)
LG_OP(MulS, a * alpha)
LG_OP(RSubS, alpha - a)
LG_OP(RDivS, alpha / a)
LG_OP(PowS, t_pow(a, alpha))
LG_OP(RPowS, t_pow(alpha, a))
LG_OP(Sqrt, t_sqrt(a))
LG_OP(DivS, a / alpha)
LG_OP(Fill, alpha)

// tanh-GELU exactly as the lambda of examples/bert.py:12 associates it:
//   0.5 * x * (1.0 + tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))
struct Gelu {
    template <typename T>
    static __device__ __forceinline__ T apply(T x, T, T, T) {
        T inner = (x * T(0.7978845608)) * (T(1.0) + (T(0.044715) * x) * x);
        return (T(0.5) * x) * (T(1.0) + t_tanh(inner));
    }
};
struct GeluBwd {  // a = x, b = g
    template <typename T>
    static __device__ __forceinline__ T apply(T x, T g, T, T) {
        const T c1 = T(0.7978845608), c2 = T(0.044715);
        T x2 = x * x;
        T u = (x * c1) * (T(1.0) + c2 * x2);
        T t = t_tanh(u);
        T du = c1 * (T(1.0) + T(3.0) * c2 * x2);
        T d = T(0.5) * (T(1.0) + t) + (T(0.5) * x) * (T(1.0) - t * t) * du;
        return d * g;
    }
};

// ---- two inputs (cpu/ops.py:60-105 forward; 158-229 backward with a = saved, b = out_grad) ----
LG_OP(Add, a + b)
LG_OP(Sub, a - b)
LG_OP(Mul, a * b)
LG_OP(Div, a / b)
LG_OP(Pow, t_pow(a, b))
LG_OP(SinBwd, t_cos(a) * b)
LG_OP(CosBwd, -t_sin(a) * b)
LG_OP(LogBwd, (T(1) / a) * b)
LG_OP(SigmoidBwd, a * (T(1) - a) * b)
LG_OP(TanhBwd, (T(1) - a * a) * b)
LG_OP(ReluBwd, (a >= T(0)) ? b : T(0))          // grad 1 at 0 (cpu/ops.py:227-229)
LG_OP(PowSBwd, alpha * t_pow(a, alpha - T(1)) * b)
LG_OP(RPowSBwd, b * a * t_log(alpha))
LG_OP(RDivSBwd, -alpha / (a * a) * b)
LG_OP(Axpy, a + alpha * b)

// ---- three inputs ----
LG_OP(DivBwdB, -a / (b * b) * c)                 // cpu/ops.py:92-94
LG_OP(PowBwdA, b * t_pow(a, b - T(1)) * c)       // cpu/ops.py:103-105
LG_OP(PowBwdB, c * b * t_log(a))                 // b = y
LG_OP(EqMaskMul, (a == b) ? c : T(0))            // cpu/ops.py:268-272

#undef LG_OP
}  // namespace op
}  // namespace lg
