// SpecialFunctions.jl functions the CUDA math library does not have (test/runtests.jl:34-36 of the reference):
// airyai, airyaiprime, airybi, airybiprime, dawson, erfi, invdigamma.  Evaluated in double for both element types.
//
// The file compiles as plain C++ too (NB_SF_FN / NB_SF_TAB): tests/test_host_logic.py builds it with g++ and checks
// every function against mpmath / scipy over the whole real line without a GPU; the device build is the same source.
//
//   Airy:    y'' = x y.  |x| <= 16: Taylor recurrence (n+2)(n+1) a_{n+2} = x0 a_n + a_{n-1} from the nearest anchor
//            x0 (spacing 1/2, |h| <= 1/4, 26 terms; anchor values correctly rounded from mpmath,
//            tools/gen_airy_table.py) — each of Ai, Bi is continued from its OWN exact anchor over at most 1/4, so the
//            recessive solution is not polluted by the dominant one.  |x| > 16: the expansions DLMF 9.7.5-9.7.12
//            (16 terms, last term < 2e-19 at zeta = 42.7).
//   dawson:  |x| < 1/2 Maclaurin series; up to 25 Rybicki's sampling formula (samples 0.4 apart, error e^{-(pi/0.4)^2} ~ 1e-27);
//            beyond, the asymptotic series.  d/dx = 1 - 2 x F(x).
//   erfi:    2/sqrt(pi) exp(x^2) dawson(x), Maclaurin below 1/2; x^2 is split by an fma so that exp sees its exact
//            argument.  d/dx = 2/sqrt(pi) exp(x^2).
//   invdigamma: Newton on digamma(x) = y from Minka's closed-form start (the scheme SpecialFunctions.jl uses),
//            run to the last bit.  d/dy = 1 / trigamma(invdigamma(y)).
#pragma once
#if defined(__CUDACC__)
#define NB_SF_FN static __device__ __noinline__
#define NB_SF_TAB static __device__ const double
#else
#include <math.h>
#define NB_SF_FN static inline
#define NB_SF_TAB static const double
#endif

#include "special_airy_table.inc"

#define NB_SF_SQRTPI 1.77245385090551602730
#define NB_SF_PI_4 0.78539816339744830962

// which: 0 Ai, 1 Ai', 2 Bi, 3 Bi'
NB_SF_FN double nb_airy(double x, int which) {
  if (x != x) return x;
  const bool bi = which >= 2, prime = which & 1;
  if (isinf(x)) {  // Ai, Ai' -> 0 and Bi, Bi' -> +inf on the right; Ai, Bi -> 0 on the left, the derivatives oscillate unboundedly
    if (x > 0.0) return bi ? INFINITY : 0.0;
    return prime ? nan("") : 0.0;
  }
  if (fabs(x) <= NB_AIRY_XA) {
    int j = (int)floor((x + NB_AIRY_XA) * NB_AIRY_INV_H + 0.5);
    j = j < 0 ? 0 : j > NB_AIRY_N - 1 ? NB_AIRY_N - 1 : j;
    const double x0 = -NB_AIRY_XA + (double)j / NB_AIRY_INV_H, h = x - x0;
    double am1 = 0.0, a0 = NB_AIRY_ANCHORS[4 * j + (bi ? 2 : 0)], a1 = NB_AIRY_ANCHORS[4 * j + (bi ? 3 : 1)];
    double y = a0 + a1 * h, yp = a1, hn = h;  // hn = h^(n+1)
    for (int n = 0; n < 26; ++n) {
      const double a2 = (x0 * a0 + am1) / (double)((n + 1) * (n + 2));
      yp += (double)(n + 2) * a2 * hn;
      hn *= h;
      y += a2 * hn;
      am1 = a0; a0 = a1; a1 = a2;
    }
    return prime ? yp : y;
  }
  const double ax = fabs(x), r = sqrt(ax), zeta = (2.0 / 3.0) * ax * r, q = sqrt(r);  // q = |x|^(1/4)
  const double iz = 1.0 / zeta;
  const double* c = prime ? NB_AIRY_V : NB_AIRY_U;
  if (x > 0.0) {
    // Ai ~ e^-z / (2 sqrt(pi) q) sum (-1)^k u_k z^-k;  Ai' ~ -q e^-z / (2 sqrt(pi)) sum (-1)^k v_k z^-k
    // Bi ~ e^z  / (sqrt(pi) q)   sum u_k z^-k;         Bi' ~  q e^z  / sqrt(pi)     sum v_k z^-k
    double s = 0.0, p = 1.0;
    for (int k = 0; k < NB_AIRY_NASY; ++k) {
      s += ((!bi && (k & 1)) ? -c[k] : c[k]) * p;
      p *= iz;
    }
    const double amp = prime ? q : 1.0 / q;
    if (bi) return exp(zeta) * amp * s / NB_SF_SQRTPI;
    return (prime ? -0.5 : 0.5) * exp(-zeta) * amp * s / NB_SF_SQRTPI;
  }
  // x < 0: even / odd parts P = sum (-1)^k c_2k z^-2k, Q = sum (-1)^k c_2k+1 z^-(2k+1)
  double P = 0.0, Q = 0.0, p = 1.0;
  for (int k = 0; k < NB_AIRY_NASY; k += 2) {
    const double sg = (k & 2) ? -1.0 : 1.0;
    P += sg * c[k] * p;
    p *= iz;
    Q += sg * c[k + 1] * p;
    p *= iz;
  }
  const double th = zeta - NB_SF_PI_4, sn = sin(th), cs = cos(th);
  switch (which) {
    case 0: return (cs * P + sn * Q) / (NB_SF_SQRTPI * q);
    case 1: return q * (sn * P - cs * Q) / NB_SF_SQRTPI;
    case 2: return (-sn * P + cs * Q) / (NB_SF_SQRTPI * q);
    default: return q * (cs * P + sn * Q) / NB_SF_SQRTPI;
  }
}

NB_SF_FN double nb_dawson(double x) {
  const double ax = fabs(x);
  if (x != x) return x;
  if (ax < 0.5) {  // F = x sum (-2 x^2)^k / (2k+1)!!
    const double t = -2.0 * x * x;
    double term = 1.0, s = 1.0;
    for (int k = 1; k < 15; ++k) {
      term *= t / (double)(2 * k + 1);
      s += term;
    }
    return x * s;
  }
  if (ax > 25.0) {  // F ~ 1/(2x) sum (2k-1)!! / (2 x^2)^k
    const double t = 0.5 / (ax * ax);
    double term = 1.0, s = 1.0;
    for (int k = 1; k < 8; ++k) {
      term *= (double)(2 * k - 1) * t;
      s += term;
    }
    return (x < 0.0 ? -0.5 : 0.5) * s / ax;
  }
  // Rybicki: F(x) = 1/sqrt(pi) sum_{n odd} e^{-(x - n h)^2} / n, centred on the even multiple of h nearest to x
  const double H = 0.2;  // samples 2H apart: error e^{-(pi / 2H)^2} = 2e-27
  const int n0 = 2 * (int)(0.5 * ax / H + 0.5);
  const double xp = ax - (double)n0 * H;
  double e1 = exp(2.0 * xp * H);
  const double e2 = e1 * e1;
  double d1 = (double)n0 + 1.0, d2 = d1 - 2.0, s = 0.0;
  for (int i = 0; i < 17; ++i) {
    const double w = (double)(2 * i + 1) * H;
    s += exp(-w * w) * (e1 / d1 + 1.0 / (d2 * e1));
    d1 += 2.0;
    d2 -= 2.0;
    e1 *= e2;
  }
  const double f = exp(-xp * xp) * s / NB_SF_SQRTPI;
  return x < 0.0 ? -f : f;
}

// exp(x^2) with the rounding error of x*x folded back in
NB_SF_FN double nb_exp_sq(double x) {
  if (isinf(x)) return INFINITY;
  const double s = x * x, e = fma(x, x, -s);
  return exp(s) * (1.0 + e);
}

NB_SF_FN double nb_erfi(double x) {
  if (isinf(x)) return x;
  if (fabs(x) < 0.5) {  // 2/sqrt(pi) sum x^(2k+1) / (k! (2k+1))
    const double t = x * x;
    double pw = 1.0, s = 1.0;
    for (int k = 1; k < 16; ++k) {
      pw *= t / (double)k;
      s += pw / (double)(2 * k + 1);
    }
    return (2.0 / NB_SF_SQRTPI) * x * s;
  }
  return (2.0 / NB_SF_SQRTPI) * nb_exp_sq(x) * nb_dawson(x);
}

// digamma / trigamma for x > 0: recurrence to x >= 10, then the asymptotic series
NB_SF_FN double nb_sf_digamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x -
         f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760 - f / 12))))));
}
NB_SF_FN double nb_sf_trigamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + 1.0 / x + 0.5 * f +
         (f / x) * (1.0 / 6 - f * (1.0 / 30 - f * (1.0 / 42 - f * (1.0 / 30 - f * (5.0 / 66 - f * (691.0 / 2730 - f * 7.0 / 6))))));
}

NB_SF_FN double nb_invdigamma(double y) {
  if (y != y) return y;
  double x = y >= -2.22 ? exp(y) + 0.5 : -1.0 / (y + 0.57721566490153286061);
  if (!(x < 1.0e300)) return x;  // exp overflow: digamma(x) ~ log x
  for (int it = 0; it < 40; ++it) {
    const double d = (nb_sf_digamma_pos(x) - y) / nb_sf_trigamma_pos(x);
    double xn = x - d;
    if (!(xn > 0.0)) xn = 0.5 * x;  // stay in the domain
    const bool done = fabs(xn - x) <= 4.0e-16 * xn;
    x = xn;
    if (done) break;
  }
  return x;
}

// ---- integer-order Bessel functions and polygamma(n, x): the second-argument-only binary functions of
// test/runtests.jl:43-47 (the reference tests them with integer first arguments 0..5; a non-integer order gives NaN).
//   J_n: power series below |x| = max(3, 2 sqrt(n+1)); periodic trapezoid rule on (1/2pi) int cos(n t - x sin t) dt with 128 points
//        (aliasing error J_{128-n}(x), < 1e-30 for |x| <= 50); Hankel's expansion beyond.
//   I_n: the same three regimes on (1/2pi) int e^{x cos t} cos(n t) dt.
//   K_n: trapezoid rule on int_0^inf e^{-x cosh t} cosh(n t) dt, step min(1/8, 1/(2 sqrt x)) (doubly exponential decay,
//        error e^{-2 pi d / h} with d ~ 1.4 the half-width of the strip where the integrand decays).
//   Y_n: y0 / y1 of the math library and the upward recurrence (stable for Y).
//   polygamma(n, x) = (-1)^(n+1) n! zeta(n+1, x): shift to x >= 20 by the defining sum, then Euler-Maclaurin.
#define NB_SF_2PI 6.28318530717958647692

// sum_k (-1)^k? a_k / (8x)^k pieces of Hankel's expansions: P (even k, alternating), Q (odd k, alternating), and the
// plain sums with (+) or (-1)^k signs for K_n / I_n.  mu = 4 n^2.
NB_SF_FN void nb_hankel_terms(int n, double ax, double* P, double* Q, double* Sp, double* Sm) {
  const double mu = 4.0 * (double)n * (double)n, i8x = 0.125 / ax;
  double a = 1.0, p = 1.0, q = 0.0, sp = 1.0, sm = 1.0;
  for (int k = 1; k <= 30; ++k) {
    const double od = (double)(2 * k - 1);
    const double an = a * (mu - od * od) / (double)k * i8x;
    if (fabs(an) > fabs(a) && k > n) break;  // the expansion has started to diverge
    a = an;
    sp += a;
    sm += (k & 1) ? -a : a;
    if (k & 1) q += (k & 2) ? -a : a;   // k = 1: +, 3: -, 5: + ...
    else p += (k & 2) ? -a : a;         // k = 2: -, 4: +, ...
    if (fabs(a) < 1e-18) break;
  }
  *P = p; *Q = q; *Sp = sp; *Sm = sm;
}

// kind 0: J_n(x), kind 1: I_n(x); any real x, any integer n
NB_SF_FN double nb_bessel_ji(int n, double x, int kind) {
  if (x != x) return x;
  double sign = 1.0;
  if (n < 0) { n = -n; if (kind == 0 && (n & 1)) sign = -sign; }
  if (x < 0.0) { x = -x; if (n & 1) sign = -sign; }
  if (x < 3.0 || x * x < 4.0 * (double)(n + 1)) {  // (x/2)^n / n! sum (-+ x^2/4)^k / (k! (n+1)_k): terms decrease, no cancellation
    const double t = (kind == 0 ? -0.25 : 0.25) * x * x;
    double term = 1.0, s = 1.0;
    for (int k = 1; k < 40; ++k) {
      term *= t / ((double)k * (double)(k + n));
      s += term;
      if (fabs(term) < 1e-18 * fabs(s)) break;
    }
    double lead = 1.0;
    for (int k = 1; k <= n; ++k) lead *= 0.5 * x / (double)k;
    return sign * lead * s;
  }
  if (isinf(x)) return kind == 0 ? 0.0 : sign * INFINITY;
  if (x > 50.0) {
    double P, Q, Sp, Sm;
    nb_hankel_terms(n, x, &P, &Q, &Sp, &Sm);
    if (kind == 1) return sign * exp(x) * Sm / sqrt(NB_SF_2PI * x);
    const double chi = x - (0.5 * (double)n + 0.25) * (0.5 * NB_SF_2PI);
    return sign * sqrt(4.0 / (NB_SF_2PI * x)) * (P * cos(chi) - Q * sin(chi));
  }
  // periodic trapezoid rule, N = 128; the integrand is even about pi: f(0) + f(pi) + 2 sum_{k=1}^{63} f(t_k)
  const int N = 128;
  const double dt = NB_SF_2PI / (double)N;
  double s;
  if (kind == 0) s = 1.0 + ((n & 1) ? -1.0 : 1.0);
  else s = 1.0 + ((n & 1) ? -1.0 : 1.0) * exp(-2.0 * x);   // terms scaled by e^-x
  for (int k = 1; k < N / 2; ++k) {
    const double t = dt * (double)k;
    if (kind == 0) s += 2.0 * cos((double)n * t - x * sin(t));
    else s += 2.0 * exp(x * (cos(t) - 1.0)) * cos((double)n * t);
  }
  s /= (double)N;
  return sign * (kind == 0 ? s : s * exp(x));
}

// K_n(x), x > 0
NB_SF_FN double nb_bessel_k(int n, double x) {
  if (x != x) return x;
  if (!(x > 0.0)) return x == 0.0 ? INFINITY : nan("");
  if (isinf(x)) return 0.0;
  if (n < 0) n = -n;
  const double nn = (double)n;
  if (x < 1.0e-8) {  // leading terms (relative error O(x^2 log x)): the integrand would need t up to log(1/x)
    if (n == 0) return -log(0.5 * x) - 0.57721566490153286061;
    double r = 0.5;
    for (int k = 1; k < n; ++k) r *= (double)k;
    for (int k = 0; k < n; ++k) r *= 2.0 / x;   // (n-1)!/2 (2/x)^n, overflowing to +inf when it must
    return r;
  }
  double h = 0.5 / sqrt(x);
  if (h > 0.125) h = 0.125;
  double s = 0.5;  // t = 0: cosh(0) = 1, half weight
  for (int k = 1; k < 4000; ++k) {
    const double t = h * (double)k, em = exp(-t);
    const double ch1 = 0.5 * (1.0 - em) * (1.0 - em) / em;  // cosh t - 1 without cancellation
    const double e = -x * ch1;
    const double term = 0.5 * (exp(e + nn * t) + exp(e - nn * t));
    s += term;
    if (term < 1e-18 * s && x * ch1 > nn * t) break;
  }
  return h * s * exp(-x);
}

// Y_n(x), x > 0
NB_SF_FN double nb_bessel_y(int n, double x) {
  if (x != x) return x;
  if (!(x > 0.0)) return x == 0.0 ? -INFINITY : nan("");
  double sign = 1.0;
  if (n < 0) { n = -n; if (n & 1) sign = -1.0; }
  double ym = y0(x);
  if (n == 0) return ym;
  double yc = y1(x);
  for (int k = 1; k < n; ++k) {
    const double yn_ = (2.0 * (double)k / x) * yc - ym;
    ym = yc;
    yc = yn_;
  }
  return sign * yc;
}

// first-argument dispatch: nu must be an integer.  which: 0 J, 1 Y, 2 I, 3 K
NB_SF_FN double nb_bessel(int which, double nu, double x) {
  if (nu != floor(nu) || fabs(nu) > 1.0e6) return nan("");
  const int n = (int)nu;
  switch (which) {
    case 0: return nb_bessel_ji(n, x, 0);
    case 1: return nb_bessel_y(n, x);
    case 2: return nb_bessel_ji(n, x, 1);
    default: return nb_bessel_k(n, x);
  }
}
// d/dx: (J_{n-1} - J_{n+1})/2, same for Y; (I_{n-1} + I_{n+1})/2; -(K_{n-1} + K_{n+1})/2
NB_SF_FN double nb_bessel_dx(int which, double nu, double x) {
  const double lo = nb_bessel(which, nu - 1.0, x), hi = nb_bessel(which, nu + 1.0, x);
  if (which <= 1) return 0.5 * (lo - hi);
  return which == 2 ? 0.5 * (lo + hi) : -0.5 * (lo + hi);
}

// polygamma(n, x), integer n >= 0, any real x that is not a pole
NB_SF_FN double nb_polygamma_n(double nu, double x) {
  if (x != x || nu != floor(nu) || nu < 0.0 || nu > 100.0) return nan("");
  if (x <= 0.0 && x == floor(x)) return nan("");
  if (x < -1.0e5) return nan("");  // (the shift below would not terminate in reasonable time)
  const int n = (int)nu;
  if (isinf(x)) return n == 0 ? x : 0.0;
  if (n == 0) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    return r + nb_sf_digamma_pos(x);
  }
  // zeta(s, x) = sum_k (x + k)^-s, s = n + 1
  const double s = (double)(n + 1);
  double z = 0.0;
  const double X0 = 20.0 + (double)n;
  while (x < X0) { z += pow(x, -s); x += 1.0; }
  // Euler-Maclaurin tail: x^(1-s)/(s-1) + x^-s/2 + sum_j B_2j/(2j)! (s)_(2j-1) x^(-s-2j+1)
  const double B2J_OVER_FACT[10] = {1.0 / 12, -1.0 / 720, 1.0 / 30240, -1.0 / 1209600, 1.0 / 47900160,
                                    -691.0 / 1307674368000.0, 1.0 / 74724249600.0, -3617.0 / 10670622842880000.0,
                                    43867.0 / 5109094217170944000.0, -174611.0 / 802857662698291200000.0};
  const double xs = pow(x, -s), ix2 = 1.0 / (x * x);
  double tail = x * xs / (s - 1.0) + 0.5 * xs;
  double poch = s, pw = xs / x;  // (s)_(2j-1) and x^(-s-2j+1) for j = 1
  for (int j = 1; j <= 10; ++j) {
    tail += B2J_OVER_FACT[j - 1] * poch * pw;
    poch *= (s + (double)(2 * j - 1)) * (s + (double)(2 * j));
    pw *= ix2;
  }
  z += tail;
  double fact = 1.0;
  for (int k = 2; k <= n; ++k) fact *= (double)k;
  return ((n & 1) ? 1.0 : -1.0) * fact * z;
}
