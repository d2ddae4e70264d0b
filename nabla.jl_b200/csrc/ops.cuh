// Device-side scalar catalogue: values and DiffRules/ChainRules partial derivatives of every
// nb_op, plus the program interpreter used for composite scalar functions.
//
// Replaces `fmad(f, x, Val{N})` (src/core.jl:285-292): the reference evaluates f on a
// ForwardDiff.Dual per element; here single-primitive programs use the analytic partial directly
// (same formula DiffRules would apply), and composite programs are differentiated by an exact
// reverse accumulation over the program DAG (same derivative, one pass for all operands).
#pragma once
#include "common.cuh"
#include "special.cuh"

struct nb_prog {
  int32_t nargs;
  int32_t n_instr;
  int32_t n_consts;
  int32_t op[NB_MAX_INSTR];
  int32_t a[NB_MAX_INSTR];
  int32_t b[NB_MAX_INSTR];
  double consts[NB_MAX_CONSTS];
  // derived by nb_prog_create for the shared-memory interpreter (GenEvS in bcast.cu)
  uint8_t hop[NB_MAX_INSTR];     // dense dispatch index (NB_HOP_*): the switch becomes a jump table
  uint8_t rflags[NB_MAX_INSTR];  // NB_RF_* : what the reverse step of instruction i reads and how it writes
  uint32_t arg_live;             // bit k: operand k reaches the result (otherwise its sensitivity is zero)
  uint32_t fwd_skip;             // pullback mode: bit i = the VALUE of instruction i is read by no partial that is
                                 // computed (nor needed to compute one that is): the forward pass skips it
  // the program selects between branches (NB_OP_MASK / NB_OP_MASKN): a zero adjoint then contributes exactly zero
  // (0 * Inf from the branch NOT taken must not reach the result: ForwardDiff never evaluates its derivative)
  uint32_t guard;
};

// adjoint * partial, with the selection guard above
template <typename T>
__device__ __forceinline__ T nb_gmul(T g, T d, bool guard) { return (guard && g == T(0)) ? T(0) : g * d; }

enum {
  NB_HOP_IDENTITY = 0, NB_HOP_NEG, NB_HOP_ABS2, NB_HOP_SQRT, NB_HOP_INV, NB_HOP_EXP, NB_HOP_LOG, NB_HOP_TANH,
  NB_HOP_LOGISTIC, NB_HOP_ADD, NB_HOP_SUB, NB_HOP_MUL, NB_HOP_DIV, NB_HOP_COLD_UN, NB_HOP_COLD_BIN
};
enum {
  NB_RF_NEED_A = 1, NB_RF_NEED_B = 2, NB_RF_NEED_Y = 4,  // values the partials read
  NB_RF_FIRST_A = 8, NB_RF_FIRST_B = 16,                 // first contribution to that adjoint: store, do not add
  NB_RF_SLOT_A = 32, NB_RF_SLOT_B = 64,                  // operand is a slot (not a constant) -> gets an adjoint
  NB_RF_LIVE = 128                                       // the instruction reaches the result
};

#define NB_PI 3.14159265358979323846
#define NB_D2R (NB_PI / 180.0)
#define NB_R2D (180.0 / NB_PI)
#define NB_LN2 0.69314718055994530942
#define NB_LN10 2.30258509299404568402
#define NB_2_SQRTPI 1.12837916709551257390

// Internal (not ABI) fused primitives un(bin(a, b)) — `tanh.(A .+ b)`, `log.(f .+ eps)`, `exp.(a .* b)` … as ONE
// compile-time-specialised op of the lean kernels instead of a two-instruction program on the interpreter (113
// thread-instructions per element, 17-28 % of the HBM peak).  bit 12 marks the family, bit 11 "the derivative of `un`
// is taken from the saved forward value y" (only for un whose derivative is a function of y alone, and only when the
// Branch's saved value is passed), bits 4-10 the unary op, bits 0-3 the binary op - NB_OP_ADD.
__host__ __device__ constexpr int nb_fop(int un, int bin, int use_y) {
  return 0x1000 | (use_y ? 0x800 : 0) | (un << 4) | (bin - NB_OP_ADD);
}
__host__ __device__ constexpr bool nb_is_fop(int op) { return (op & 0x1000) != 0; }
__host__ __device__ constexpr int nb_fop_un(int op) { return (op >> 4) & 0x7F; }
__host__ __device__ constexpr int nb_fop_bin(int op) { return NB_OP_ADD + (op & 0xF); }
__host__ __device__ constexpr bool nb_fop_y(int op) { return (op & 0x800) != 0; }
__host__ __device__ constexpr bool nb_fop_supported(int un, int bin) {
  return (un == NB_OP_TANH || un == NB_OP_EXP || un == NB_OP_LOG || un == NB_OP_LOGISTIC || un == NB_OP_SQRT ||
          un == NB_OP_ABS2) && (bin == NB_OP_ADD || bin == NB_OP_SUB || bin == NB_OP_MUL);
}
__host__ __device__ constexpr bool nb_fop_un_y_only(int un) {
  return un == NB_OP_TANH || un == NB_OP_EXP || un == NB_OP_LOGISTIC || un == NB_OP_SQRT;
}

__host__ __device__ constexpr bool nb_op_is_unary(int op) {
  return (op >= 0 && op < NB_OP_UNARY_END) || (op >= NB_OP_SF_BEGIN && op < NB_OP_SF_END);
}
__host__ __device__ constexpr bool nb_op_is_binary(int op) {
  return (op >= NB_OP_ADD && op < NB_OP_BINARY_END) || nb_is_fop(op);
}

// Which inputs does the partial derivative need?  bit0: operand a, bit1: operand b, bit2: y.
// (for d/da of a unary op; binary ops report the union over both partials — refined on the host
//  per wanted operand by nb_partial_needs)
__host__ __device__ constexpr int nb_partial_needs(int op, int wrt /*0=a,1=b*/) {
  const int A = 1, B = 2, Y = 4;
  if (nb_is_fop(op))   // f'(bin(a, b)) from the saved y or from both operands, times the inner partial
    return (nb_fop_y(op) ? Y : (A | B)) | (nb_fop_bin(op) == NB_OP_MUL ? (wrt == 0 ? B : A) : 0);
  switch (op) {
    case NB_OP_IDENTITY: case NB_OP_NEG: case NB_OP_DEG2RAD: case NB_OP_RAD2DEG: return 0;
    case NB_OP_SIGN: case NB_OP_STEP: case NB_OP_FLOOR: case NB_OP_CEIL: case NB_OP_ROUND: case NB_OP_TRUNC: return 0;
    case NB_OP_GT: case NB_OP_LT: case NB_OP_GE: case NB_OP_LE: return 0;
    case NB_OP_MASK: case NB_OP_MASKN: return wrt == 0 ? 0 : A;   // d/dv is the indicator of the condition (operand a)
    case NB_OP_TANH: case NB_OP_EXP: case NB_OP_LOGISTIC: case NB_OP_SQRT: case NB_OP_INV:
    case NB_OP_TAN: case NB_OP_EXP2: case NB_OP_EXP10: case NB_OP_CBRT: case NB_OP_COT:
    case NB_OP_COTH: case NB_OP_TAND: case NB_OP_COTD: case NB_OP_EXPM1:
      return Y;
    case NB_OP_SEC: case NB_OP_CSC: case NB_OP_SECH: case NB_OP_CSCH: case NB_OP_SECD:
    case NB_OP_CSCD:
      return A | Y;
    case NB_OP_ADD: case NB_OP_SUB: return 0;
    case NB_OP_MUL: return wrt == 0 ? B : A;
    case NB_OP_DIV: return wrt == 0 ? B : (B | Y);   // 1/b ; -y/b
    case NB_OP_LDIV: return wrt == 0 ? (A | Y) : A;  // -y/a ; 1/a
    case NB_OP_POW: return wrt == 0 ? (A | B) : (A | Y);
    case NB_OP_HYPOT: return wrt == 0 ? (A | Y) : (B | Y);
    case NB_OP_MAX: case NB_OP_MIN: case NB_OP_ATAN2: return A | B;
    case NB_OP_BETA: return A | B | Y;
    case NB_OP_BESSELJ: case NB_OP_BESSELY: case NB_OP_BESSELI: case NB_OP_BESSELK: case NB_OP_POLYGAMMA:
      return wrt == 0 ? 0 : (A | B);   // d/d(order) is the constant NaN
    case NB_OP_ERFINV: case NB_OP_ERFCINV: case NB_OP_INVDIGAMMA: return Y;
    case NB_OP_GAMMA: case NB_OP_ERFCX: case NB_OP_DAWSON: return A | Y;
    default: return A;  // unary ops whose partial is a function of x
  }
}

template <typename T> __device__ __forceinline__ T nb_sinpi(T x);
template <> __device__ __forceinline__ float nb_sinpi<float>(float x) { return sinpif(x); }
template <> __device__ __forceinline__ double nb_sinpi<double>(double x) { return sinpi(x); }
template <typename T> __device__ __forceinline__ T nb_cospi(T x);
template <> __device__ __forceinline__ float nb_cospi<float>(float x) { return cospif(x); }
template <> __device__ __forceinline__ double nb_cospi<double>(double x) { return cospi(x); }
template <typename T> __device__ __forceinline__ T nb_exp10(T x);
template <> __device__ __forceinline__ float nb_exp10<float>(float x) { return exp10f(x); }
template <> __device__ __forceinline__ double nb_exp10<double>(double x) { return exp10(x); }
template <typename T> __device__ __forceinline__ T nb_rsqrt(T x);
template <> __device__ __forceinline__ float nb_rsqrt<float>(float x) { return 1.0f / sqrtf(x); }
template <> __device__ __forceinline__ double nb_rsqrt<double>(double x) { return 1.0 / sqrt(x); }


// ---- SpecialFunctions.jl subset (test/runtests.jl:34-36): values from the CUDA math library, polygammas by
// recurrence to x >= 10 + asymptotic series (|error| < 1e-15 relative there) and the reflection formulas ----
template <typename T> __device__ __forceinline__ T nb_j0(T x);
template <> __device__ __forceinline__ float nb_j0<float>(float x) { return j0f(x); }
template <> __device__ __forceinline__ double nb_j0<double>(double x) { return j0(x); }
template <typename T> __device__ __forceinline__ T nb_j1(T x);
template <> __device__ __forceinline__ float nb_j1<float>(float x) { return j1f(x); }
template <> __device__ __forceinline__ double nb_j1<double>(double x) { return j1(x); }
template <typename T> __device__ __forceinline__ T nb_y0(T x);
template <> __device__ __forceinline__ float nb_y0<float>(float x) { return y0f(x); }
template <> __device__ __forceinline__ double nb_y0<double>(double x) { return y0(x); }
template <typename T> __device__ __forceinline__ T nb_y1(T x);
template <> __device__ __forceinline__ float nb_y1<float>(float x) { return y1f(x); }
template <> __device__ __forceinline__ double nb_y1<double>(double x) { return y1(x); }
template <typename T> __device__ __forceinline__ T nb_erfinv(T x);
template <> __device__ __forceinline__ float nb_erfinv<float>(float x) { return erfinvf(x); }
template <> __device__ __forceinline__ double nb_erfinv<double>(double x) { return erfinv(x); }
template <typename T> __device__ __forceinline__ T nb_erfcinv(T x);
template <> __device__ __forceinline__ float nb_erfcinv<float>(float x) { return erfcinvf(x); }
template <> __device__ __forceinline__ double nb_erfcinv<double>(double x) { return erfcinv(x); }
template <typename T> __device__ __forceinline__ T nb_erfcx(T x);
template <> __device__ __forceinline__ float nb_erfcx<float>(float x) { return erfcxf(x); }
template <> __device__ __forceinline__ double nb_erfcx<double>(double x) { return erfcx(x); }

// polygamma(ORDER, x) for ORDER = 0 (digamma), 1 (trigamma), 2; evaluated in double for both element types
template <int ORDER>
__device__ __noinline__ double nb_polygamma(double x) {
  double refl = 0.0;
  bool reflected = false;
  if (x <= 0.0) {
    if (x == floor(x)) return nan("");
    const double s = sinpi(x), c = cospi(x);
    // psi(1-x) - psi(x) = pi cot(pi x);  psi1(1-x) + psi1(x) = pi^2 / sin^2;  psi2(1-x) - psi2(x) = 2 pi^3 cos / sin^3
    if (ORDER == 0) refl = -NB_PI * c / s;
    else if (ORDER == 1) refl = NB_PI * NB_PI / (s * s);
    else refl = -2.0 * NB_PI * NB_PI * NB_PI * c / (s * s * s);
    x = 1.0 - x;
    reflected = true;
  }
  double r = 0.0;
  while (x < 10.0) {
    if (ORDER == 0) r -= 1.0 / x;
    else if (ORDER == 1) r += 1.0 / (x * x);
    else r -= 2.0 / (x * x * x);
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  double v;
  if (ORDER == 0) {
    v = log(x) - 0.5 / x -
        f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760 - f / 12))))));
  } else if (ORDER == 1) {
    v = 1.0 / x + 0.5 * f +
        (f / x) * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66 - f * (691.0 / 2730 - f * 7.0 / 6))))));
  } else {
    v = -f - f / x -
        f * f * (0.5 - f * (1.0 / 6 - f * (1.0 / 6 - f * (0.3 - f * (5.0 / 6 - f * (691.0 / 210 - f * 35.0 / 2))))));
  }
  v += r;
  if (!reflected) return v;
  // values at the original point from the values at 1 - x
  if (ORDER == 0) return v + refl;       // psi(x)  = psi(1-x) - pi cot(pi x)
  if (ORDER == 1) return refl - v;       // psi1(x) = pi^2/sin^2 - psi1(1-x)
  return v + refl;                       // psi2(x) = psi2(1-x) - 2 pi^3 cos/sin^3
}

// value of a catalogue op: full table, inlined (folds to one expression when `op` is a compile-time
// constant, as in the lean kernels)
template <typename T>
__device__ __forceinline__ T nb_eval_inner(int bin, T a, T b) {   // the binary half of a fused primitive
  return bin == NB_OP_ADD ? a + b : bin == NB_OP_SUB ? a - b : a * b;
}
template <typename T>
__device__ __forceinline__ T nb_eval_outer(int un, T t) {         // the unary half
  switch (un) {
    case NB_OP_TANH: return tanh(t);
    case NB_OP_EXP: return exp(t);
    case NB_OP_LOG: return log(t);
    case NB_OP_LOGISTIC: return T(1) / (T(1) + exp(-t));
    case NB_OP_SQRT: return sqrt(t);
    default: return t * t;   // NB_OP_ABS2
  }
}

template <typename T>
__device__ __forceinline__ T nb_eval_full(int op, T a, T b) {
  const T one = T(1);
  const T d2r = T(NB_D2R), r2d = T(NB_R2D);
  if (nb_is_fop(op)) return nb_eval_outer<T>(nb_fop_un(op), nb_eval_inner<T>(nb_fop_bin(op), a, b));
  switch (op) {
    case NB_OP_IDENTITY: return a;
    case NB_OP_NEG: return -a;
    case NB_OP_ABS: return fabs(a);
    case NB_OP_ABS2: return a * a;
    case NB_OP_SQRT: return sqrt(a);
    case NB_OP_CBRT: return cbrt(a);
    case NB_OP_INV: return one / a;
    case NB_OP_EXP: return exp(a);
    case NB_OP_EXP2: return exp2(a);
    case NB_OP_EXP10: return nb_exp10<T>(a);
    case NB_OP_EXPM1: return expm1(a);
    case NB_OP_LOG: return log(a);
    case NB_OP_LOG2: return log2(a);
    case NB_OP_LOG10: return log10(a);
    case NB_OP_LOG1P: return log1p(a);
    case NB_OP_SIN: return sin(a);
    case NB_OP_COS: return cos(a);
    case NB_OP_TAN: return tan(a);
    case NB_OP_SEC: return one / cos(a);
    case NB_OP_CSC: return one / sin(a);
    case NB_OP_COT: return one / tan(a);
    case NB_OP_SINH: return sinh(a);
    case NB_OP_COSH: return cosh(a);
    case NB_OP_TANH: return tanh(a);
    case NB_OP_SECH: return one / cosh(a);
    case NB_OP_CSCH: return one / sinh(a);
    case NB_OP_COTH: return one / tanh(a);
    case NB_OP_ASIN: return asin(a);
    case NB_OP_ACOS: return acos(a);
    case NB_OP_ATAN: return atan(a);
    case NB_OP_ASEC: return acos(one / a);
    case NB_OP_ACSC: return asin(one / a);
    case NB_OP_ACOT: return atan(one / a);
    case NB_OP_ASINH: return asinh(a);
    case NB_OP_ACOSH: return acosh(a);
    case NB_OP_ATANH: return atanh(a);
    case NB_OP_ASECH: return acosh(one / a);
    case NB_OP_ACSCH: return asinh(one / a);
    case NB_OP_ACOTH: return atanh(one / a);
    case NB_OP_SIND: return sin(d2r * a);
    case NB_OP_COSD: return cos(d2r * a);
    case NB_OP_TAND: return tan(d2r * a);
    case NB_OP_SECD: return one / cos(d2r * a);
    case NB_OP_CSCD: return one / sin(d2r * a);
    case NB_OP_COTD: return one / tan(d2r * a);
    case NB_OP_ASIND: return r2d * asin(a);
    case NB_OP_ACOSD: return r2d * acos(a);
    case NB_OP_ATAND: return r2d * atan(a);
    case NB_OP_ASECD: return r2d * acos(one / a);
    case NB_OP_ACSCD: return r2d * asin(one / a);
    case NB_OP_ACOTD: return r2d * atan(one / a);
    case NB_OP_SINPI: return nb_sinpi<T>(a);
    case NB_OP_COSPI: return nb_cospi<T>(a);
    case NB_OP_DEG2RAD: return d2r * a;
    case NB_OP_RAD2DEG: return r2d * a;
    case NB_OP_LOGISTIC: return one / (one + exp(-a));
    case NB_OP_ERF: return erf(a);
    case NB_OP_ERFC: return erfc(a);
    case NB_OP_ADD: return a + b;
    case NB_OP_SUB: return a - b;
    case NB_OP_MUL: return a * b;
    case NB_OP_DIV: return a / b;
    case NB_OP_LDIV: return b / a;
    case NB_OP_POW: return pow(a, b);
    case NB_OP_HYPOT: return hypot(a, b);
    case NB_OP_MAX: return fmax(a, b);
    case NB_OP_MIN: return fmin(a, b);
    case NB_OP_ATAN2: return atan2(a, b);
    case NB_OP_SIGN: return a > T(0) ? one : (a < T(0) ? -one : a);   // sign(+-0) = +-0, sign(NaN) = NaN (Julia)
    case NB_OP_STEP: return a > T(0) ? one : T(0);
    case NB_OP_FLOOR: return floor(a);
    case NB_OP_CEIL: return ceil(a);
    case NB_OP_ROUND: return rint(a);   // Julia's round: ties to even
    case NB_OP_TRUNC: return trunc(a);
    case NB_OP_GT: return a > b ? one : T(0);
    case NB_OP_LT: return a < b ? one : T(0);
    case NB_OP_GE: return a >= b ? one : T(0);
    case NB_OP_LE: return a <= b ? one : T(0);
    case NB_OP_MASK: return a != T(0) ? b : T(0);
    case NB_OP_MASKN: return a == T(0) ? b : T(0);
    case NB_OP_BETA: return exp(lgamma(a) + lgamma(b) - lgamma(a + b));
    case NB_OP_BESSELJ: return T(nb_bessel(0, (double)a, (double)b));
    case NB_OP_BESSELY: return T(nb_bessel(1, (double)a, (double)b));
    case NB_OP_BESSELI: return T(nb_bessel(2, (double)a, (double)b));
    case NB_OP_BESSELK: return T(nb_bessel(3, (double)a, (double)b));
    case NB_OP_POLYGAMMA: return T(nb_polygamma_n((double)a, (double)b));
    case NB_OP_GAMMA: return tgamma(a);
    case NB_OP_DIGAMMA: return T(nb_polygamma<0>((double)a));
    case NB_OP_TRIGAMMA: return T(nb_polygamma<1>((double)a));
    case NB_OP_BESSELJ0: return nb_j0<T>(a);
    case NB_OP_BESSELJ1: return nb_j1<T>(a);
    case NB_OP_BESSELY0: return nb_y0<T>(a);
    case NB_OP_BESSELY1: return nb_y1<T>(a);
    case NB_OP_ERFINV: return nb_erfinv<T>(a);
    case NB_OP_ERFCINV: return nb_erfcinv<T>(a);
    case NB_OP_ERFCX: return nb_erfcx<T>(a);
    case NB_OP_AIRYAI: return T(nb_airy((double)a, 0));
    case NB_OP_AIRYAIPRIME: return T(nb_airy((double)a, 1));
    case NB_OP_AIRYBI: return T(nb_airy((double)a, 2));
    case NB_OP_AIRYBIPRIME: return T(nb_airy((double)a, 3));
    case NB_OP_DAWSON: return T(nb_dawson((double)a));
    case NB_OP_ERFI: return T(nb_erfi((double)a));
    case NB_OP_INVDIGAMMA: return T(nb_invdigamma((double)a));
    default: return a;
  }
}

// the same table out of line, so kernels with a RUNTIME opcode stay small
template <typename T>
__device__ __noinline__ T nb_eval_cold(int op, T a, T b) { return nb_eval_full<T>(op, a, b); }

// d/da of a unary op; y is f(a) (valid only if nb_partial_needs says Y)
template <typename T>
__device__ __forceinline__ T nb_dunary_full(int op, T a, T y) {
  const T one = T(1);
  const T d2r = T(NB_D2R), r2d = T(NB_R2D);
  switch (op) {
    case NB_OP_IDENTITY: return one;
    case NB_OP_NEG: return -one;
    case NB_OP_ABS: return a > T(0) ? one : (a < T(0) ? -one : T(0));
    case NB_OP_SIGN: case NB_OP_STEP: case NB_OP_FLOOR: case NB_OP_CEIL: case NB_OP_ROUND: case NB_OP_TRUNC: return T(0);
    case NB_OP_ABS2: return T(2) * a;
    case NB_OP_SQRT: return T(0.5) / y;
    case NB_OP_CBRT: return one / (T(3) * y * y);
    case NB_OP_INV: return -(y * y);
    case NB_OP_EXP: return y;
    case NB_OP_EXP2: return T(NB_LN2) * y;
    case NB_OP_EXP10: return T(NB_LN10) * y;
    case NB_OP_EXPM1: return y + one;
    case NB_OP_LOG: return one / a;
    case NB_OP_LOG2: return one / (a * T(NB_LN2));
    case NB_OP_LOG10: return one / (a * T(NB_LN10));
    case NB_OP_LOG1P: return one / (one + a);
    case NB_OP_SIN: return cos(a);
    case NB_OP_COS: return -sin(a);
    case NB_OP_TAN: return one + y * y;
    case NB_OP_SEC: return y * tan(a);
    case NB_OP_CSC: return -y / tan(a);
    case NB_OP_COT: return -(one + y * y);
    case NB_OP_SINH: return cosh(a);
    case NB_OP_COSH: return sinh(a);
    case NB_OP_TANH: return one - y * y;
    case NB_OP_SECH: return -tanh(a) * y;
    case NB_OP_CSCH: return -y / tanh(a);
    case NB_OP_COTH: return one - y * y;
    case NB_OP_ASIN: return nb_rsqrt<T>(one - a * a);
    case NB_OP_ACOS: return -nb_rsqrt<T>(one - a * a);
    case NB_OP_ATAN: return one / (one + a * a);
    case NB_OP_ASEC: return one / (fabs(a) * sqrt(a * a - one));
    case NB_OP_ACSC: return -one / (fabs(a) * sqrt(a * a - one));
    case NB_OP_ACOT: return -one / (one + a * a);
    case NB_OP_ASINH: return nb_rsqrt<T>(a * a + one);
    case NB_OP_ACOSH: return nb_rsqrt<T>(a * a - one);
    case NB_OP_ATANH: return one / (one - a * a);
    case NB_OP_ASECH: return -one / (a * sqrt(one - a * a));
    case NB_OP_ACSCH: return -one / (fabs(a) * sqrt(one + a * a));
    case NB_OP_ACOTH: return one / (one - a * a);
    case NB_OP_SIND: return d2r * cos(d2r * a);
    case NB_OP_COSD: return -d2r * sin(d2r * a);
    case NB_OP_TAND: return d2r * (one + y * y);
    case NB_OP_SECD: return d2r * y * tan(d2r * a);
    case NB_OP_CSCD: return -d2r * y / tan(d2r * a);
    case NB_OP_COTD: return -d2r * (one + y * y);
    case NB_OP_ASIND: return r2d * nb_rsqrt<T>(one - a * a);
    case NB_OP_ACOSD: return -r2d * nb_rsqrt<T>(one - a * a);
    case NB_OP_ATAND: return r2d / (one + a * a);
    case NB_OP_ASECD: return r2d / (fabs(a) * sqrt(a * a - one));
    case NB_OP_ACSCD: return -r2d / (fabs(a) * sqrt(a * a - one));
    case NB_OP_ACOTD: return -r2d / (one + a * a);
    case NB_OP_SINPI: return T(NB_PI) * nb_cospi<T>(a);
    case NB_OP_COSPI: return -T(NB_PI) * nb_sinpi<T>(a);
    case NB_OP_DEG2RAD: return d2r;
    case NB_OP_RAD2DEG: return r2d;
    case NB_OP_LOGISTIC: return y * (one - y);
    case NB_OP_ERF: return T(NB_2_SQRTPI) * exp(-a * a);
    case NB_OP_ERFC: return -T(NB_2_SQRTPI) * exp(-a * a);
    case NB_OP_GAMMA: return y * T(nb_polygamma<0>((double)a));
    case NB_OP_DIGAMMA: return T(nb_polygamma<1>((double)a));
    case NB_OP_TRIGAMMA: return T(nb_polygamma<2>((double)a));
    case NB_OP_BESSELJ0: return -nb_j1<T>(a);
    case NB_OP_BESSELJ1: return nb_j0<T>(a) - nb_j1<T>(a) / a;
    case NB_OP_BESSELY0: return -nb_y1<T>(a);
    case NB_OP_BESSELY1: return nb_y0<T>(a) - nb_y1<T>(a) / a;
    case NB_OP_ERFINV: return T(1.0 / NB_2_SQRTPI) * exp(y * y);
    case NB_OP_ERFCINV: return -T(1.0 / NB_2_SQRTPI) * exp(y * y);
    case NB_OP_ERFCX: return T(2) * a * y - T(NB_2_SQRTPI);
    case NB_OP_AIRYAI: return T(nb_airy((double)a, 1));          // Ai'
    case NB_OP_AIRYAIPRIME: return a * T(nb_airy((double)a, 0)); // Ai'' = x Ai
    case NB_OP_AIRYBI: return T(nb_airy((double)a, 3));
    case NB_OP_AIRYBIPRIME: return a * T(nb_airy((double)a, 2));
    case NB_OP_DAWSON: return T(1) - T(2) * a * y;
    case NB_OP_ERFI: return T(NB_2_SQRTPI * nb_exp_sq((double)a));
    case NB_OP_INVDIGAMMA: return T(1.0 / nb_sf_trigamma_pos((double)y));
    default: return T(0);
  }
}

template <typename T>
__device__ __noinline__ T nb_dunary_cold(int op, T a, T y) { return nb_dunary_full<T>(op, a, y); }

// Hot ops are inlined into the kernels; the long tail of the catalogue goes through one call.
template <typename T>
__device__ __forceinline__ T nb_eval(int op, T a, T b) {
  switch (op) {
    case NB_OP_IDENTITY: return a;
    case NB_OP_NEG: return -a;
    case NB_OP_ABS2: return a * a;
    case NB_OP_SQRT: return sqrt(a);
    case NB_OP_INV: return T(1) / a;
    case NB_OP_EXP: return exp(a);
    case NB_OP_LOG: return log(a);
    case NB_OP_TANH: return tanh(a);
    case NB_OP_LOGISTIC: return T(1) / (T(1) + exp(-a));
    case NB_OP_ADD: return a + b;
    case NB_OP_SUB: return a - b;
    case NB_OP_MUL: return a * b;
    case NB_OP_DIV: return a / b;
    default: return nb_eval_cold<T>(op, a, b);
  }
}

template <typename T>
__device__ __forceinline__ T nb_dunary(int op, T a, T y) {
  switch (op) {
    case NB_OP_IDENTITY: return T(1);
    case NB_OP_NEG: return T(-1);
    case NB_OP_ABS2: return T(2) * a;
    case NB_OP_SQRT: return T(0.5) / y;
    case NB_OP_INV: return -(y * y);
    case NB_OP_EXP: return y;
    case NB_OP_LOG: return T(1) / a;
    case NB_OP_TANH: return T(1) - y * y;
    case NB_OP_LOGISTIC: return y * (T(1) - y);
    default: return nb_dunary_cold<T>(op, a, y);
  }
}

// d/da (wrt == 0) or d/db (wrt == 1) of a binary op
template <typename T>
__device__ __forceinline__ T nb_dbinary(int op, int wrt, T a, T b, T y) {
  const T one = T(1);
  if (nb_is_fop(op)) {
    const int un = nb_fop_un(op), bin = nb_fop_bin(op);
    T fp;   // f'(t), t = bin(a, b)
    if (nb_fop_y(op)) {
      fp = un == NB_OP_TANH ? one - y * y : un == NB_OP_EXP ? y : un == NB_OP_LOGISTIC ? y * (one - y) : T(0.5) / y;
    } else {
      const T t = nb_eval_inner<T>(bin, a, b);
      switch (un) {
        case NB_OP_TANH: { const T w = tanh(t); fp = one - w * w; break; }
        case NB_OP_EXP: fp = exp(t); break;
        case NB_OP_LOG: fp = one / t; break;
        case NB_OP_LOGISTIC: { const T w = one / (one + exp(-t)); fp = w * (one - w); break; }
        case NB_OP_SQRT: fp = T(0.5) / sqrt(t); break;
        default: fp = T(2) * t; break;
      }
    }
    const T inner = bin == NB_OP_MUL ? (wrt == 0 ? b : a) : (bin == NB_OP_SUB && wrt == 1 ? -one : one);
    return fp * inner;
  }
  switch (op) {
    case NB_OP_ADD: return one;
    case NB_OP_SUB: return wrt == 0 ? one : -one;
    case NB_OP_MUL: return wrt == 0 ? b : a;
    case NB_OP_DIV: return wrt == 0 ? one / b : -y / b;
    case NB_OP_LDIV: return wrt == 0 ? -y / a : one / a;
    case NB_OP_POW: return wrt == 0 ? b * pow(a, b - one) : y * log(a);
    case NB_OP_HYPOT: return wrt == 0 ? a / y : b / y;
    case NB_OP_MAX: return wrt == 0 ? (a > b ? one : T(0)) : (a > b ? T(0) : one);
    case NB_OP_MIN: return wrt == 0 ? (a < b ? one : T(0)) : (a < b ? T(0) : one);
    case NB_OP_ATAN2: {
      T den = a * a + b * b;
      return wrt == 0 ? b / den : -a / den;
    }
    case NB_OP_GT: case NB_OP_LT: case NB_OP_GE: case NB_OP_LE: return T(0);
    case NB_OP_MASK: return wrt == 0 ? T(0) : (a != T(0) ? one : T(0));
    case NB_OP_MASKN: return wrt == 0 ? T(0) : (a == T(0) ? one : T(0));
    case NB_OP_BETA: {
      const double pab = nb_polygamma<0>((double)a + (double)b);
      return y * T(nb_polygamma<0>((double)(wrt == 0 ? a : b)) - pab);
    }
    case NB_OP_BESSELJ: case NB_OP_BESSELY: case NB_OP_BESSELI: case NB_OP_BESSELK:
      return wrt == 0 ? T(nan("")) : T(nb_bessel_dx(op - NB_OP_BESSELJ, (double)a, (double)b));
    case NB_OP_POLYGAMMA: return wrt == 0 ? T(nan("")) : T(nb_polygamma_n((double)a + 1.0, (double)b));
    default: return T(0);
  }
}

// ---- composite programs ---------------------------------------------------------------------
#define NB_MAX_SLOTS (NB_MAX_ARGS + NB_MAX_INSTR)

template <typename T>
__device__ __forceinline__ T nb_slot(const nb_prog& P, const T* vals, int idx) {
  return idx >= 0 ? vals[idx] : T(P.consts[-idx - 1]);
}

// forward evaluation: vals[0..nargs) must hold the operands; returns the program value
template <typename T>
__device__ __forceinline__ T nb_prog_eval(const nb_prog& P, T* vals) {
  for (int i = 0; i < P.n_instr; ++i) {
    T a = nb_slot<T>(P, vals, P.a[i]);
    T b = nb_op_is_binary(P.op[i]) ? nb_slot<T>(P, vals, P.b[i]) : T(0);
    vals[P.nargs + i] = nb_eval<T>(P.op[i], a, b);
  }
  return P.n_instr ? vals[P.nargs + P.n_instr - 1] : vals[0];
}

// reverse accumulation: adj[] receives d(program value)/d(slot); vals must be fully evaluated
template <typename T>
__device__ __forceinline__ void nb_prog_grad(const nb_prog& P, const T* vals, T* adj) {
  const int nslots = P.nargs + P.n_instr;
  for (int s = 0; s < nslots; ++s) adj[s] = T(0);
  if (P.n_instr == 0) {
    adj[0] = T(1);
    return;
  }
  adj[nslots - 1] = T(1);
  for (int i = P.n_instr - 1; i >= 0; --i) {
    T g = adj[P.nargs + i];
    if (g == T(0)) continue;
    int op = P.op[i];
    T y = vals[P.nargs + i];
    T a = nb_slot<T>(P, vals, P.a[i]);
    if (nb_op_is_binary(op)) {
      T b = nb_slot<T>(P, vals, P.b[i]);
      if (P.a[i] >= 0) adj[P.a[i]] += g * nb_dbinary<T>(op, 0, a, b, y);
      if (P.b[i] >= 0) adj[P.b[i]] += g * nb_dbinary<T>(op, 1, a, b, y);
    } else {
      if (P.a[i] >= 0) adj[P.a[i]] += g * nb_dunary<T>(op, a, y);
    }
  }
}

// Fill the derived fields of a validated program (host).  nb_prog_create derives them for "every operand wanted"; a
// pullback launch of a composite program re-derives them on a copy for the operands it actually differentiates
// (`wanted`, bit k = operand k): adjoints are propagated only along paths that reach a wanted operand, and forward
// values no remaining partial reads are not computed — in `x .* log.(f .+ eps)` with x a constant (the data of
// examples/vae.jl:79) the sensitivity of f needs x and f + eps, never the logarithm.
inline void nb_prog_derive(nb_prog& P, uint32_t wanted = 0xFFFFFFFFu) {
  auto hop_of = [](int op) -> uint8_t {
    switch (op) {
      case NB_OP_IDENTITY: return NB_HOP_IDENTITY; case NB_OP_NEG: return NB_HOP_NEG;
      case NB_OP_ABS2: return NB_HOP_ABS2; case NB_OP_SQRT: return NB_HOP_SQRT; case NB_OP_INV: return NB_HOP_INV;
      case NB_OP_EXP: return NB_HOP_EXP; case NB_OP_LOG: return NB_HOP_LOG; case NB_OP_TANH: return NB_HOP_TANH;
      case NB_OP_LOGISTIC: return NB_HOP_LOGISTIC; case NB_OP_ADD: return NB_HOP_ADD;
      case NB_OP_SUB: return NB_HOP_SUB; case NB_OP_MUL: return NB_HOP_MUL; case NB_OP_DIV: return NB_HOP_DIV;
      default: return nb_op_is_binary(op) ? NB_HOP_COLD_BIN : NB_HOP_COLD_UN;
    }
  };
  const int n = P.n_instr, na = P.nargs;
  bool live[NB_MAX_INSTR] = {false};
  bool written[NB_MAX_SLOTS] = {false};
  bool reach[NB_MAX_SLOTS] = {false};     // the slot is a wanted operand or is computed from one
  bool need_val[NB_MAX_SLOTS] = {false};  // the slot's forward value is read in pullback mode
  for (int k = 0; k < na; ++k) reach[k] = (wanted >> k & 1u) != 0;
  for (int i = 0; i < n; ++i) {
    const bool bin = nb_op_is_binary(P.op[i]);
    const int a = P.a[i], b = bin ? P.b[i] : -1;
    reach[na + i] = (a >= 0 && reach[a]) || (b >= 0 && reach[b]);
  }
  if (n > 0) live[n - 1] = true;
  for (int i = n - 1; i >= 0; --i) {
    P.hop[i] = hop_of(P.op[i]);
    P.rflags[i] = 0;
    if (!live[i] || !reach[na + i]) continue;
    const bool bin = nb_op_is_binary(P.op[i]);
    const int a = P.a[i], b = bin ? P.b[i] : -1;
    const bool ca = a >= 0 && reach[a], cb = bin && b >= 0 && reach[b];
    uint8_t f = NB_RF_LIVE;
    int needs = 0;
    if (ca) { f |= NB_RF_SLOT_A; needs |= nb_partial_needs(P.op[i], 0); }
    if (cb) { f |= NB_RF_SLOT_B; needs |= nb_partial_needs(P.op[i], 1); }
    if (needs & 1) { f |= NB_RF_NEED_A; if (a >= 0) need_val[a] = true; }
    if (needs & 2) { f |= NB_RF_NEED_B; if (b >= 0) need_val[b] = true; }
    if (needs & 4) { f |= NB_RF_NEED_Y; need_val[na + i] = true; }
    if (ca) {
      if (!written[a]) { f |= NB_RF_FIRST_A; written[a] = true; }
      if (a >= na) live[a - na] = true;
    }
    if (cb && b != a) {
      if (!written[b]) { f |= NB_RF_FIRST_B; written[b] = true; }
      if (b >= na) live[b - na] = true;
    }
    P.rflags[i] = f;
  }
  // a value that is read needs its own operands' values
  P.fwd_skip = 0;
  for (int i = n - 1; i >= 0; --i) {
    if (!need_val[na + i]) { P.fwd_skip |= 1u << i; continue; }
    const bool bin = nb_op_is_binary(P.op[i]);
    if (P.a[i] >= 0) need_val[P.a[i]] = true;
    if (bin && P.b[i] >= 0) need_val[P.b[i]] = true;
  }
  P.guard = 0;
  for (int i = 0; i < n; ++i)
    if (P.op[i] == NB_OP_MASK || P.op[i] == NB_OP_MASKN) P.guard = 1;
  P.arg_live = 0;
  for (int k = 0; k < na; ++k)
    if (written[k] || n == 0) P.arg_live |= 1u << k;
}

// value of instruction `hop` (dense index) on V elements: a jump table instead of a compare chain
template <typename T, int V>
__device__ __forceinline__ void nb_eval_hop(int hop, int op, const T (&a)[V], const T (&b)[V], T (&r)[V]) {
#define NB_HOP_CASE(H, EXPR) \
  case H:                    \
    _Pragma("unroll") for (int v = 0; v < V; ++v) { const T x = a[v], y = b[v]; (void)y; r[v] = (EXPR); } \
    break;
  switch (hop) {
    NB_HOP_CASE(NB_HOP_IDENTITY, x) NB_HOP_CASE(NB_HOP_NEG, -x) NB_HOP_CASE(NB_HOP_ABS2, x * x)
    NB_HOP_CASE(NB_HOP_SQRT, sqrt(x)) NB_HOP_CASE(NB_HOP_INV, T(1) / x) NB_HOP_CASE(NB_HOP_EXP, exp(x))
    NB_HOP_CASE(NB_HOP_LOG, log(x)) NB_HOP_CASE(NB_HOP_TANH, tanh(x))
    NB_HOP_CASE(NB_HOP_LOGISTIC, T(1) / (T(1) + exp(-x))) NB_HOP_CASE(NB_HOP_ADD, x + y)
    NB_HOP_CASE(NB_HOP_SUB, x - y) NB_HOP_CASE(NB_HOP_MUL, x * y) NB_HOP_CASE(NB_HOP_DIV, x / y)
    default:
#pragma unroll
      for (int v = 0; v < V; ++v) r[v] = nb_eval_cold<T>(op, a[v], b[v]);
  }
#undef NB_HOP_CASE
}

// partials of instruction `hop` at (a, b) with value y, V-wide
template <typename T, int V>
__device__ __forceinline__ void nb_partials_hop(int hop, int op, const T (&a)[V], const T (&b)[V], const T (&y)[V],
                                                T (&da)[V], T (&db)[V]) {
#define NB_HOP_D(H, EA, EB)  \
  case H:                    \
    _Pragma("unroll") for (int v = 0; v < V; ++v) { const T x = a[v], z = b[v], w = y[v]; (void)x; (void)z; (void)w; \
      da[v] = (EA); db[v] = (EB); } \
    break;
  switch (hop) {
    NB_HOP_D(NB_HOP_IDENTITY, T(1), T(0)) NB_HOP_D(NB_HOP_NEG, T(-1), T(0)) NB_HOP_D(NB_HOP_ABS2, T(2) * x, T(0))
    NB_HOP_D(NB_HOP_SQRT, T(0.5) / w, T(0)) NB_HOP_D(NB_HOP_INV, -(w * w), T(0)) NB_HOP_D(NB_HOP_EXP, w, T(0))
    NB_HOP_D(NB_HOP_LOG, T(1) / x, T(0)) NB_HOP_D(NB_HOP_TANH, T(1) - w * w, T(0))
    NB_HOP_D(NB_HOP_LOGISTIC, w * (T(1) - w), T(0)) NB_HOP_D(NB_HOP_ADD, T(1), T(1)) NB_HOP_D(NB_HOP_SUB, T(1), T(-1))
    NB_HOP_D(NB_HOP_MUL, z, x) NB_HOP_D(NB_HOP_DIV, T(1) / z, -w / z)
    case NB_HOP_COLD_BIN:
#pragma unroll
      for (int v = 0; v < V; ++v) {
        da[v] = nb_dbinary<T>(op, 0, a[v], b[v], y[v]);
        db[v] = nb_dbinary<T>(op, 1, a[v], b[v], y[v]);
      }
      break;
    default:
#pragma unroll
      for (int v = 0; v < V; ++v) { da[v] = nb_dunary_cold<T>(op, a[v], y[v]); db[v] = T(0); }
  }
#undef NB_HOP_D
}

// ---- V-wide program interpreter: one opcode decode per V elements --------------------------------
// (the scalar interpreter above decodes every instruction once per ELEMENT; profiles/r1: 5-9 % of HBM
//  peak for a two-instruction lambda.  Here slot values live as V-vectors in thread-local memory and
//  the opcode switch is hoisted out of the element loop for the hot primitives.)
template <typename T, int V>
__device__ __forceinline__ void nb_eval_vec(int op, const T (&a)[V], const T (&b)[V], T (&r)[V]) {
#define NB_VEC_CASE(OPC, EXPR) \
  case OPC:                    \
    _Pragma("unroll") for (int v = 0; v < V; ++v) { const T x = a[v], y = b[v]; (void)y; r[v] = (EXPR); } \
    break;
  switch (op) {
    NB_VEC_CASE(NB_OP_IDENTITY, x) NB_VEC_CASE(NB_OP_NEG, -x) NB_VEC_CASE(NB_OP_ABS2, x * x)
    NB_VEC_CASE(NB_OP_SQRT, sqrt(x)) NB_VEC_CASE(NB_OP_INV, T(1) / x) NB_VEC_CASE(NB_OP_EXP, exp(x))
    NB_VEC_CASE(NB_OP_LOG, log(x)) NB_VEC_CASE(NB_OP_TANH, tanh(x))
    NB_VEC_CASE(NB_OP_LOGISTIC, T(1) / (T(1) + exp(-x))) NB_VEC_CASE(NB_OP_ADD, x + y)
    NB_VEC_CASE(NB_OP_SUB, x - y) NB_VEC_CASE(NB_OP_MUL, x * y) NB_VEC_CASE(NB_OP_DIV, x / y)
    default:
#pragma unroll
      for (int v = 0; v < V; ++v) r[v] = nb_eval_cold<T>(op, a[v], b[v]);
  }
#undef NB_VEC_CASE
}

// partials of instruction `op` at (a, b) with value y: da (and db for binary ops), V-wide
template <typename T, int V>
__device__ __forceinline__ void nb_partials_vec(int op, const T (&a)[V], const T (&b)[V], const T (&y)[V],
                                                T (&da)[V], T (&db)[V]) {
#define NB_VEC_D(OPC, EA, EB)  \
  case OPC:                    \
    _Pragma("unroll") for (int v = 0; v < V; ++v) { const T x = a[v], z = b[v], w = y[v]; (void)x; (void)z; (void)w; \
      da[v] = (EA); db[v] = (EB); } \
    break;
  switch (op) {
    NB_VEC_D(NB_OP_IDENTITY, T(1), T(0)) NB_VEC_D(NB_OP_NEG, T(-1), T(0)) NB_VEC_D(NB_OP_ABS2, T(2) * x, T(0))
    NB_VEC_D(NB_OP_SQRT, T(0.5) / w, T(0)) NB_VEC_D(NB_OP_INV, -(w * w), T(0)) NB_VEC_D(NB_OP_EXP, w, T(0))
    NB_VEC_D(NB_OP_LOG, T(1) / x, T(0)) NB_VEC_D(NB_OP_TANH, T(1) - w * w, T(0))
    NB_VEC_D(NB_OP_LOGISTIC, w * (T(1) - w), T(0)) NB_VEC_D(NB_OP_ADD, T(1), T(1)) NB_VEC_D(NB_OP_SUB, T(1), T(-1))
    NB_VEC_D(NB_OP_MUL, z, x) NB_VEC_D(NB_OP_DIV, T(1) / z, -w / z)
    default:
      if (nb_op_is_binary(op)) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          da[v] = nb_dbinary<T>(op, 0, a[v], b[v], y[v]);
          db[v] = nb_dbinary<T>(op, 1, a[v], b[v], y[v]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) { da[v] = nb_dunary_cold<T>(op, a[v], y[v]); db[v] = T(0); }
      }
  }
#undef NB_VEC_D
}

template <typename T, int V>
__device__ __forceinline__ void nb_slot_vec(const nb_prog& P, T (*vals)[V], int idx, T (&out)[V]) {
  if (idx >= 0) {
#pragma unroll
    for (int v = 0; v < V; ++v) out[v] = vals[idx][v];
  } else {
    const T c = T(P.consts[-idx - 1]);
#pragma unroll
    for (int v = 0; v < V; ++v) out[v] = c;
  }
}

// forward: vals[0..nargs) hold the operands on entry; on exit every slot is evaluated
template <typename T, int V>
__device__ __forceinline__ void nb_prog_eval_vec(const nb_prog& P, T (*vals)[V]) {
  for (int i = 0; i < P.n_instr; ++i) {
    T a[V], b[V], r[V];
    nb_slot_vec<T, V>(P, vals, P.a[i], a);
    if (nb_op_is_binary(P.op[i])) nb_slot_vec<T, V>(P, vals, P.b[i], b);
    else {
#pragma unroll
      for (int v = 0; v < V; ++v) b[v] = T(0);
    }
    nb_eval_vec<T, V>(P.op[i], a, b, r);
#pragma unroll
    for (int v = 0; v < V; ++v) vals[P.nargs + i][v] = r[v];
  }
}

// reverse accumulation over the program DAG: adj[s] = d(program value)/d(slot s), V-wide
template <typename T, int V>
__device__ __forceinline__ void nb_prog_grad_vec(const nb_prog& P, T (*vals)[V], T (*adj)[V]) {
  const int nslots = P.nargs + P.n_instr;
  for (int s = 0; s < nslots; ++s)
#pragma unroll
    for (int v = 0; v < V; ++v) adj[s][v] = T(0);
  const int last = P.n_instr ? nslots - 1 : 0;
#pragma unroll
  for (int v = 0; v < V; ++v) adj[last][v] = T(1);
  for (int i = P.n_instr - 1; i >= 0; --i) {
    const int op = P.op[i], ia = P.a[i], ib = P.b[i];
    const bool bin = nb_op_is_binary(op);
    T g[V], a[V], b[V], y[V], da[V], db[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { g[v] = adj[P.nargs + i][v]; y[v] = vals[P.nargs + i][v]; }
    nb_slot_vec<T, V>(P, vals, ia, a);
    if (bin) nb_slot_vec<T, V>(P, vals, ib, b);
    else {
#pragma unroll
      for (int v = 0; v < V; ++v) b[v] = T(0);
    }
    nb_partials_vec<T, V>(op, a, b, y, da, db);
    const bool gd = P.guard != 0;
    if (ia >= 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) adj[ia][v] += nb_gmul(g[v], da[v], gd);
    }
    if (bin && ib >= 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) adj[ib][v] += nb_gmul(g[v], db[v], gd);
    }
  }
}
