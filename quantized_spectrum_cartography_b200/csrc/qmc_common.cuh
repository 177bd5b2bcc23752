// Shared device helpers of libqmc_b200.so: error plumbing, the probit bin likelihood and its
// derivative in a tail-stable form, warp reductions.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qmc_b200.h"
#include "erfcx_coeffs.h"

namespace qmc {

// ---- host-side error plumbing (defined in qmc_abi.cu) ---------------------------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
#define QMC_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::qmc::set_error(QMC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
#define QMC_REQUIRE(cond, ...)                                       \
  do {                                                               \
    if (!(cond)) return ::qmc::set_error(QMC_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// The reference writes sqrt(2) as 1.414213 (quantization_model.py:61) and forms sigma*1.414213
// in Python double before it meets the fp32 tensor.
inline float probit_scale(float noise_std) { return (float)((double)noise_std * 1.414213); }

// ---- device math -----------------------------------------------------------------------------
constexpr float kInvSqrtPi = 0.56418958354775628695f;
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kLn2 = 0.69314718055994530942f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp / log / reciprocal in two grades: FAST uses one SFU op each (ex2/lg2/rcp.approx, ~2^-22
// relative), otherwise the CUDA library functions (~1 ulp).  The SFU grade is what the
// throughput kernels use; the library grade exists to measure what the SFU grade costs in parity.
template <bool FAST>
__device__ __forceinline__ float q_exp(float x) {
  return FAST ? ex2_approx(x * kLog2e) : expf(x);
}
template <bool FAST>
__device__ __forceinline__ float q_log(float x) {
  return FAST ? lg2_approx(x) * kLn2 : logf(x);
}
template <bool FAST>
__device__ __forceinline__ float q_rcp(float x) {
  return FAST ? rcp_approx(x) : 1.0f / x;
}

// erfcx(w) = exp(w^2) erfc(w) for w >= 0 (w = +inf gives 0).  One reciprocal and a low-degree
// Horner chain; see gen_erfcx_coeffs.py for the derivation and the measured error.
template <bool FAST>
__device__ __forceinline__ float erfcx_pos(float w) {
  constexpr float c[QMC_ERFCX_DEG + 1] = QMC_ERFCX_COEFFS;
  const float t = q_rcp<FAST>(w + QMC_ERFCX_C);
  const float q = fmaf(-2.0f * QMC_ERFCX_C, t, 1.0f);
  float p = c[QMC_ERFCX_DEG];
#pragma unroll
  for (int i = QMC_ERFCX_DEG - 1; i >= 0; --i) p = fmaf(p, q, c[i]);
  return p * t;
}

struct BinEval {
  float logp;  // log P(level | x)
  float gx;    // d(-log P)/dx
};

// Two-sided probit bin:  P = F(hi - x) - F(lo - x),  F(y) = 0.5*(1 + erf(y*inv_a))
// (quantization_model.py:38,61), so with zl = (lo-x)*inv_a <= zu = (hi-x)*inv_a
//     P = 0.5*(erfc(zl) - erfc(zu)),   dP/dx = -(inv_a/sqrt(pi)) * (exp(-zu^2) - exp(-zl^2)).
// Stable evaluation of log P and of (exp(-zu^2) - exp(-zl^2))/P:
//   both bounds on one side of x (a tail): factor exp(-n^2) of the nearer bound n out of both
//   terms, P = exp(-n^2) * 0.5*(erfcx(n) - D*erfcx(f)), D = exp(n^2 - f^2) <= 1, so that
//   log P = -n^2 + log(core) never underflows and the ratio is (1-D)/core up to sign;
//   bounds straddling x: P = 1 - 0.5*(erfc|zl| + erfc(zu)), no cancellation.
template <bool FAST>
__device__ __forceinline__ BinEval probit_bin_stable(float lo, float hi, float x, float inv_a) {
  const float zl = (lo - x) * inv_a;
  const float zu = (hi - x) * inv_a;
  const float wl = fabsf(zl), wu = fabsf(zu);
  const float rl = erfcx_pos<FAST>(wl), ru = erfcx_pos<FAST>(wu);
  const float gscale = inv_a * kInvSqrtPi;
  BinEval o;
  if (FAST) {
    // Branch-free form for the throughput kernels (neighbouring entries fall on different sides of their bins, so a
    // warp would walk both branches anyway): the tail and the straddling case share the reciprocals of erfcx, one
    // exponential (D or El), the logarithm and the final reciprocal; same operations, hence the same bits, as the
    // branches below.
    const bool right = zl >= 0.0f, tail = right || zu <= 0.0f;
    const float n = right ? wl : wu, f = right ? wu : wl;
    const float rn = right ? rl : ru, rf = right ? ru : rl;
    const float Ea = ex2_approx((tail ? (n - f) * (n + f) : -wl * wl) * kLog2e);   // D (tail) or El (straddle)
    const float Eu = ex2_approx((-wu * wu) * kLog2e);
    const float core = 0.5f * fmaf(-Ea, rf, rn);
    const float P = 1.0f - 0.5f * fmaf(Ea, rl, Eu * ru);
    const float arg = tail ? core : P;
    const float l = lg2_approx(arg) * kLn2;
    o.logp = tail ? fmaf(-n, n, l) : l;
    const float v = (tail ? 1.0f - Ea : Eu - Ea) * rcp_approx(arg);
    o.gx = ((tail && right) ? -v : v) * gscale;
    return o;
  }
  if (zl >= 0.0f || zu <= 0.0f) {
    const bool right = zl >= 0.0f;
    const float n = right ? wl : wu, f = right ? wu : wl;
    const float rn = right ? rl : ru, rf = right ? ru : rl;
    const float D = q_exp<FAST>((n - f) * (n + f));
    const float core = 0.5f * fmaf(-D, rf, rn);
    o.logp = fmaf(-n, n, q_log<FAST>(core));
    const float ratio = (1.0f - D) * q_rcp<FAST>(core);
    o.gx = (right ? -ratio : ratio) * gscale;
  } else {
    const float El = q_exp<FAST>(-wl * wl), Eu = q_exp<FAST>(-wu * wu);
    const float P = 1.0f - 0.5f * fmaf(El, rl, Eu * ru);
    o.logp = q_log<FAST>(P);
    o.gx = (Eu - El) * q_rcp<FAST>(P) * gscale;
  }
  return o;
}

// One-sided probit bin: the other bound is (numerically) infinite, which is every bin of the
// one-bit model with the reference's +-1e5 sentinels.  sgn = +1: P = F(x - thr) (upper bin),
// sgn = -1: P = F(thr - x) (lower bin).  P = 0.5*erfc(u), u = -sgn*(x-thr)*inv_a.
template <bool FAST>
__device__ __forceinline__ BinEval probit_one_sided(float thr, float sgn, float x, float inv_a) {
  const float u = -sgn * (x - thr) * inv_a;
  const float w = fabsf(u);
  const float rx = erfcx_pos<FAST>(w);
  BinEval o;
  float dlogp_du;
  if (u >= 0.0f) {
    o.logp = fmaf(-w, w, q_log<FAST>(0.5f * rx));
    dlogp_du = -2.0f * kInvSqrtPi * q_rcp<FAST>(rx);
  } else {
    const float E = q_exp<FAST>(-w * w);
    const float P = fmaf(-0.5f * E, rx, 1.0f);
    o.logp = q_log<FAST>(P);
    dlogp_du = -kInvSqrtPi * E * q_rcp<FAST>(P);
  }
  o.gx = dlogp_du * sgn * inv_a;  // d(-logP)/dx = -dlogp_du * du/dx, du/dx = -sgn*inv_a
  return o;
}

// Branch-free, SFU-grade form of probit_one_sided for the throughput kernels.  m = -sgn*inv_a
// (so u = (x - thr)*m); 4 SFU ops (rcp, ex2, lg2, rcp) and no divergence:
//   u >= 0:  log P = -u^2 + log(erfcx(u)/2),      d log P/du = -(1/sqrt(pi)) / (erfcx(u)/2)
//   u <  0:  log P = log(1 - e^{-u^2} erfcx|u|/2), d log P/du = -(1/sqrt(pi)) e^{-u^2} / P
__device__ __forceinline__ BinEval probit_one_sided_fast(float thr, float m, float x) {
  const float u = (x - thr) * m;
  const float w = fabsf(u);
  const float h = 0.5f * erfcx_pos<true>(w);
  const float w2 = w * w;
  const float E = ex2_approx(-kLog2e * w2);
  const bool pos = u >= 0.0f;
  const float arg = pos ? h : fmaf(-E, h, 1.0f);
  BinEval o;
  o.logp = fmaf(lg2_approx(arg), kLn2, pos ? -w2 : 0.0f);
  // d(-logP)/dx = -dlogp_du * m,  dlogp_du = -(1/sqrt(pi)) * (pos ? 1 : E) / arg
  o.gx = (kInvSqrtPi * m) * ((pos ? 1.0f : E) * rcp_approx(arg));
  return o;
}

// Logistic bin (F = F_sigmoid, quantization_model.py:43-47): with zl = (lo-x)/s < zu = (hi-x)/s,
//   P = F(zu) - F(zl) = (1 - e^{-(zu-zl)}) * F(zu) * F(-zl)
//   log P = log(-expm1(-(zu-zl))) - softplus(-zu) - softplus(zl)       (no cancellation, no overflow)
//   d(-log P)/dx = (F(-zu) - F(zl)) / s                                (the first term does not depend on x)
// SFU grade, like the probit epilogues: e^{-|t|} is one ex2 and is shared by softplus(t) = max(t, 0) + log1p(e^{-|t|})
// and F(t) = (t >= 0 ? 1 : e^{-|t|}) / (1 + e^{-|t|}); log1p(e) = ln2 * lg2(1 + e) (absolute error < 2^-23, the
// terms are O(1)); 3 ex2 + 3 lg2 + 2 rcp per bin.
__device__ __forceinline__ BinEval logistic_bin(float lo, float hi, float x, float inv_s) {
  const float zl = (lo - x) * inv_s, zu = (hi - x) * inv_s;
  const float width = (hi - lo) * inv_s;  // = zu - zl without the rounding of the two differences
  const float eu = ex2_approx(-kLog2e * fabsf(zu)), el = ex2_approx(-kLog2e * fabsf(zl));
  const float ru = rcp_approx(1.0f + eu), rl = rcp_approx(1.0f + el);
  const float sp_u = fmaxf(-zu, 0.0f) + kLn2 * lg2_approx(1.0f + eu);   // softplus(-zu)
  const float sp_l = fmaxf(zl, 0.0f) + kLn2 * lg2_approx(1.0f + el);    // softplus(zl)
  const float f_u = zu <= 0.0f ? ru : eu * ru;                           // F(-zu)
  const float f_l = zl >= 0.0f ? rl : el * rl;                           // F(zl)
  BinEval o;
  o.logp = kLn2 * lg2_approx(1.0f - ex2_approx(-kLog2e * width)) - sp_u - sp_l;
  o.gx = (f_u - f_l) * inv_s;
  return o;
}
// One-sided logistic bin (the one-bit model with the reference's +-1e5 sentinels): P = F(v), v = (x - thr) * m with
// m = +1/s for the upper level and -1/s for the lower one; log P = -softplus(-v), d(-log P)/dx = -F(-v) * m.
// 1 ex2 + 1 lg2 + 1 rcp.
__device__ __forceinline__ BinEval logistic_one_sided_fast(float thr, float m, float x) {
  const float v = (x - thr) * m;
  const float e = ex2_approx(-kLog2e * fabsf(v));
  const float r = rcp_approx(1.0f + e);
  BinEval o;
  o.logp = -(fmaxf(-v, 0.0f) + kLn2 * lg2_approx(1.0f + e));
  o.gx = -(v <= 0.0f ? r : e * r) * m;
  return o;
}

// The reference's own arithmetic, literally: F = 0.5*(1+erf(z)), P = F(zu) - F(zl), log P, and
// the gradient autograd derives from it.  Underflows to P == 0 (log -> -inf, gradient -> inf/NaN)
// where the reference does.
__device__ __forceinline__ BinEval probit_bin_reference(float lo, float hi, float x, float inv_a) {
  const float zl = (lo - x) * inv_a, zu = (hi - x) * inv_a;
  const float Fu = 0.5f * (1.0f + erff(zu));
  const float Fl = 0.5f * (1.0f + erff(zl));
  const float P = Fu - Fl;
  BinEval o;
  o.logp = logf(P);
  o.gx = (expf(-zu * zu) - expf(-zl * zl)) / P * (inv_a * kInvSqrtPi);
  return o;
}

// ---- one element of the fused factor update (qmc_solver.cu; also the tail of the fused S-step) ----------
// torch.optim.Adam (no amsgrad, no weight decay) in fp32, preceded by the gradient of lam*||p||_F
// (coef = lam/||p||) and followed by the projection onto p >= 0.  Same operations as torch; the square
// root and the quotient use the SFU (sqrt.approx, rcp.approx: ~2^-22 relative) and the bias correction is
// passed as its reciprocal, which keeps the update at a dozen instructions per element.
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, float coef, float one_m_b1, float b2,
                                          float one_m_b2, float step_size, float inv_bc2_sqrt, float eps, bool project) {
  const float gg = fmaf(coef, p, g);                         // d/dp (nll + lam*||p||_F) = g + lam*p/||p||
  m = fmaf(gg - m, one_m_b1, m);                             // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(one_m_b2 * gg, gg, v * b2);                       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = fmaf(sqrt_approx(v), inv_bc2_sqrt, eps);  // exp_avg_sq.sqrt() / bias_correction2_sqrt + eps
  float pn = fmaf(-step_size * m, rcp_approx(denom), p);     // param.addcdiv_(exp_avg, denom, value=-step_size)
  if (project && pn < 0.0f) pn = 0.0f;                       // X[X < 0] = 0
  return pn;
}

// ---- warp helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum R per-lane values over the warp with a transposing butterfly: R + log2(32/R)-ish shuffles
// instead of 5R.  After the call lane l (l < R... see below) holds a complete sum.
// Layout of the result: for R a power of two <= 32, the total of component r ends up in every
// lane whose (lane % R) == r ... restricted to lanes that took part in the final steps; callers
// use lane == r (r < R).
template <int R>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[R], int lane) {
  static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "R must be a power of two");
  // Stage j halves the number of live components: lanes with bit (R>>1 >> j)... we implement it
  // with a simple recursive halving on the register array.
  int live = R;
  int bit = 16;
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    if (live > 1) {
      const int half = live >> 1;
      const bool upper = (lane & bit) != 0;
#pragma unroll
      for (int i = 0; i < R / 2; ++i) {
        if (i < half) {
          // lanes with the bit clear keep components [0, half), send [half, live); others mirror
          const float send = upper ? v[i] : v[i + half];
          const float keep = upper ? v[i + half] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      live = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    }
    bit >>= 1;
  }
  return v[0];
}
// Which component the lane holds after warp_transpose_sum<R>: the bits of `lane` consumed by the
// halving steps (16, 8, ...), most significant first, select upper/lower halves.
template <int R>
__device__ __forceinline__ int warp_transpose_owner(int lane) {
  int r = 0, live = R, bit = 16;
  while (live > 1) {
    const int half = live >> 1;
    if (lane & bit) r += half;
    live = half;
    bit >>= 1;
  }
  return r;
}

}  // namespace qmc
