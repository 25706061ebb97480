// optmc_math.cuh -- scalar building blocks shared by every kernel (and host-compilable for unit tests).
//
// Everything here is per-path arithmetic: Philox4x32-10, uniform->normal transforms, the one-step
// updates of each path scheme, payoff, the Gram accumulators and the guarded LDL^T solve.
// Reference citations: om3 = options_model_3/options_model_3.py, om3gpu = option_model_3_gpu.py,
// hc = heston_calibration.py (all under the reference tree).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define OPTMC_HD __host__ __device__ __forceinline__
#else
#define OPTMC_HD inline
#endif

namespace optmc {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., Random123).  KATs: SURVEY.md App. A-7 / tests/test_philox.py.
// ------------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t v[4];
};

OPTMC_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

OPTMC_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

// Counter layout (results are independent of launch geometry and GPU count):
//   ctr = (pair_lo, pair_hi, step_block, stream), key = (seed_lo, seed_hi).
//   Heston: step_block b yields (z1,z2) for steps 2b+1 and 2b+2;  GBM: z for steps 4b+1..4b+4.
OPTMC_HD Philox4 philox_for(uint64_t pair, uint32_t step_block, uint32_t stream, uint64_t seed) {
  return philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), step_block, stream, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

// ------------------------------------------------------------------------------------------------
// small float helpers
// ------------------------------------------------------------------------------------------------
OPTMC_HD float u32_as_f32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(x);
#else
  union { uint32_t u; float f; } c; c.u = x; return c.f;
#endif
}

template <typename R> struct Real;  // per-precision math: device fast paths for float, IEEE for double

// Device fast-math primitives (MUFU, flush-to-zero: none of the arguments below can be denormal).
#if defined(__CUDACC__)
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

template <> struct Real<float> {
  static OPTMC_HD float exp_(float x) {
#if defined(__CUDA_ARCH__)
    return mufu_ex2(x * 1.4426950408889634f);
#else
    return expf(x);
#endif
  }
  static OPTMC_HD float sqrt_(float x) {
#if defined(__CUDA_ARCH__)
    return mufu_sqrt(x);
#else
    return sqrtf(x);
#endif
  }
  // Box-Muller on two 32-bit words -> two N(0,1).  Uniforms built by mantissa stuffing (no I2F):
  // u in (0,1] with 2^-23 resolution (|z| <= 5.65); the angle is 2 pi w with w in [1,2) (one period, so no
  // subtraction is needed).
  static OPTMC_HD void normal2(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u = 2.0f - u32_as_f32(0x3f800000u | (a >> 9));
    float ang = 6.283185307179586f * u32_as_f32(0x3f800000u | (b >> 9));
#if defined(__CUDA_ARCH__)
    float rad = mufu_sqrt(-1.3862943611198906f * mufu_lg2(u));  // sqrt(-2 ln u) = sqrt(-2 ln2 log2 u)
    float s = mufu_sin(ang), c = mufu_cos(ang);
#else
    float rad = sqrtf(-2.0f * logf(u));
    float s = sinf(ang), c = cosf(ang);
#endif
    n0 = rad * c;
    n1 = rad * s;
  }
};

template <> struct Real<double> {
  static OPTMC_HD double exp_(double x) { return exp(x); }
  static OPTMC_HD double sqrt_(double x) { return sqrt(x); }
  static OPTMC_HD void normal2(uint32_t a, uint32_t b, double& n0, double& n1) {
    double u = ((double)(~a) + 1.0) * 2.3283064365386963e-10;  // 1 - a/2^32 in (0,1]: same map as the fp32 path
    double v = (double)b * 2.3283064365386963e-10;          // [0,1)
    double rad = sqrt(-2.0 * log(u));
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(2.0 * v, &s, &c);
#else
    s = sin(6.283185307179586 * v); c = cos(6.283185307179586 * v);
#endif
    n0 = rad * c;
    n1 = rad * s;
  }
};

template <typename R> OPTMC_HD R rmax(R a, R b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------
// one-step path updates
// ------------------------------------------------------------------------------------------------
template <typename R> struct GbmConsts {
  R drift, diffusion;  // (r - sigma^2/2) dt, sigma sqrt(dt)   (om3:473-474)
};

// om3:480  S[t] = S[t-1] * exp(drift + diffusion * Z[t-1])
template <typename R> OPTMC_HD R gbm_step(R S, R z, const GbmConsts<R>& c) {
  return S * Real<R>::exp_(c.drift + c.diffusion * z);
}

template <typename R> struct HestonConsts {
  R dt, sqrt_dt, r, kappa, theta, xi, rho, rho_c;  // rho_c = sqrt(1 - rho^2)
};

// om3:228-233 (absorption Euler).  z1, z2 independent N(0,1); w2 = rho z1 + sqrt(1-rho^2) z2.
template <typename R> OPTMC_HD void heston_absorb_step(R& S, R& v, R z1, R z2, const HestonConsts<R>& c) {
  R w2 = c.rho * z1 + c.rho_c * z2;
  R vp = rmax(v, (R)0);
  R sq = Real<R>::sqrt_(vp * c.dt);
  R vn = vp + c.kappa * (c.theta - vp) * c.dt + c.xi * sq * w2;
  v = rmax(vn, (R)0);
  S = S * Real<R>::exp_((c.r - (R)0.5 * vp) * c.dt + sq * z1);
}

// Lord-Koekkoek-van Dijk full truncation: v may go negative, v+ enters drift and diffusion.
template <typename R> OPTMC_HD void heston_fulltrunc_step(R& S, R& v, R z1, R z2, const HestonConsts<R>& c) {
  R w2 = c.rho * z1 + c.rho_c * z2;
  R vp = rmax(v, (R)0);
  R sq = Real<R>::sqrt_(vp * c.dt);
  S = S * Real<R>::exp_((c.r - (R)0.5 * vp) * c.dt + sq * z1);
  v = v + c.kappa * (c.theta - vp) * c.dt + c.xi * sq * w2;
}

// hc:240-255 (calibrator): variance floored at 1e-8, arithmetic Euler on S.
template <typename R> OPTMC_HD void heston_calib_step(R& S, R& v, R z1, R z2, const HestonConsts<R>& c) {
  R w2 = c.rho * z1 + c.rho_c * z2;
  R vp = rmax(v, (R)1e-8);
  R sqv = Real<R>::sqrt_(vp);
  R dV = c.kappa * (c.theta - vp) * c.dt + c.xi * sqv * c.sqrt_dt * w2;
  v = rmax(vp + dV, (R)1e-8);
  R dS = c.r * S * c.dt + sqv * S * c.sqrt_dt * z1;
  S = S + dS;
}

// Andersen's quadratic-exponential scheme (L. Andersen, "Simple and efficient simulation of the Heston stochastic
// volatility model", J. Comp. Finance 11(3), 2008, sections 3.2.4 and 4.2; gamma1 = gamma2 = 1/2, no martingale
// correction).  Not in the reference: the variance is sampled from a distribution matched to the
// first two conditional moments of the CIR step -- a squared Gaussian when psi = s^2/m^2 <= 1.5, a mass at zero
// plus an exponential tail otherwise -- and stays non-negative.  z1 drives the asset, z2 the variance (the
// correlation enters through K0..K4); the uniform of the exponential branch is Phi(z2), so the antithetic
// partner (-z1, -z2) sees 1 - U.  Restated in oracle.lsm_oracle.heston_paths_qe.
template <typename R> struct QeConsts {
  R E, c1, c2, theta1mE;      // exp(-kappa dt); xi^2 E (1-E)/kappa; theta xi^2 (1-E)^2/(2 kappa); theta (1-E)
  R k0r, k1, k2, k3, k4;      // r dt + K0, K1..K4
};
template <typename R> OPTMC_HD QeConsts<R> qe_consts(const HestonConsts<R>& c) {
  QeConsts<R> q;
  const double dt = (double)c.dt, kappa = (double)c.kappa, theta = (double)c.theta, xi = (double)c.xi, rho = (double)c.rho;
  const double E = exp(-kappa * dt);
  q.E = (R)E;
  q.c1 = (R)(xi * xi * E * (1.0 - E) / kappa);
  q.c2 = (R)(theta * xi * xi * (1.0 - E) * (1.0 - E) / (2.0 * kappa));
  q.theta1mE = (R)(theta * (1.0 - E));
  const double a = kappa * rho / xi - 0.5;
  q.k0r = (R)((double)c.r * dt - rho * kappa * theta * dt / xi);
  q.k1 = (R)(0.5 * dt * a - rho / xi);
  q.k2 = (R)(0.5 * dt * a + rho / xi);
  q.k3 = (R)(0.5 * dt * (1.0 - rho * rho));
  q.k4 = q.k3;
  return q;
}
OPTMC_HD double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
OPTMC_HD float norm_cdf(float z) { return 0.5f * erfcf(-z * 0.70710678118654752440f); }
OPTMC_HD double log_any(double x) { return log(x); }
OPTMC_HD float log_any(float x) { return logf(x); }
template <typename R> OPTMC_HD void heston_qe_step(R& S, R& v, R z1, R z2, const QeConsts<R>& q) {
  const R m = q.theta1mE + v * q.E;
  const R s2 = v * q.c1 + q.c2;
  const R psi = s2 / (m * m);
  R vn;
  if (!(m > (R)0)) {
    vn = (R)0;
  } else if (psi <= (R)1.5) {
    const R ip = (R)2 / psi;                                        // >= 4/3
    const R b2 = ip - (R)1 + Real<R>::sqrt_(ip) * Real<R>::sqrt_(ip - (R)1);
    const R a = m / ((R)1 + b2);
    const R bz = Real<R>::sqrt_(b2) + z2;
    vn = a * bz * bz;
  } else {
    const R p = (psi - (R)1) / (psi + (R)1);
    const R beta = ((R)1 - p) / m;
    const R u = norm_cdf(z2);
    vn = u <= p ? (R)0 : log_any(((R)1 - p) / norm_cdf(-z2)) / beta;
  }
  const R var = q.k3 * v + q.k4 * vn;
  S = S * Real<R>::exp_(q.k0r + q.k1 * v + q.k2 * vn + Real<R>::sqrt_(var) * z1);
  v = vn;
}

#if defined(__CUDA_ARCH__)
// fp32 production form of the QE step: the same algebra with MUFU reciprocals / roots / log / exp (relative error
// ~1e-7 per operation, far below the fp32 storage step) and ONE erfc per exponential-branch draw
// (q = Phi(-|z|); U = q or 1 - q by the sign of z).
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <> __device__ __forceinline__ void heston_qe_step<float>(float& S, float& v, float z1, float z2, const QeConsts<float>& q) {
  const float m = fmaf(v, q.E, q.theta1mE);
  const float s2 = fmaf(v, q.c1, q.c2);
  float vn = 0.f;
  if (m > 0.f) {
    const float rm = mufu_rcp(m);
    const float psi = s2 * rm * rm;
    if (psi <= 1.5f) {
      const float ip = 2.0f * mufu_rcp(psi);  // >= 4/3
      const float b2 = ip - 1.0f + mufu_sqrt(ip * (ip - 1.0f));
      const float a = m * mufu_rcp(1.0f + b2);
      const float bz = mufu_sqrt(b2) + z2;
      vn = a * bz * bz;
    } else {
      const float p = (psi - 1.0f) * mufu_rcp(psi + 1.0f);
      const float tail = 0.5f * erfcf(fabsf(z2) * 0.70710678118654752440f);  // Phi(-|z2|)
      const float u = z2 < 0.f ? tail : 1.0f - tail;                        // Phi(z2)
      const float one_minus_u = z2 < 0.f ? 1.0f - tail : tail;
      // ln((1 - p) / (1 - U)) / beta,  beta = (1 - p) / m
      vn = u <= p ? 0.f : 0.6931471805599453f * (mufu_lg2(1.0f - p) - mufu_lg2(one_minus_u)) * m * mufu_rcp(1.0f - p);
    }
  }
  const float var = q.k3 * (v + vn);
  S *= mufu_ex2(1.4426950408889634f * (fmaf(q.k1, v, q.k0r) + fmaf(q.k2, vn, mufu_sqrt(var) * z1)));
  v = vn;
}
#endif

// Scheme dispatch shared by the path and the fused European kernels (scheme ids: include/optmc.h).
template <typename R, int SCHEME>
OPTMC_HD void heston_step_any(R& S, R& v, R z1, R z2, const HestonConsts<R>& c, const QeConsts<R>& q) {
  if (SCHEME == 2) heston_absorb_step<R>(S, v, z1, z2, c);
  else if (SCHEME == 3) heston_fulltrunc_step<R>(S, v, z1, z2, c);
  else if (SCHEME == 5) heston_qe_step<R>(S, v, z1, z2, q);
  else heston_calib_step<R>(S, v, z1, z2, c);
}

// fp32 production form of the two Euler schemes for one antithetic pair (+z, -z): the drift / diffusion
// constants are pre-folded (log2 e into the exponent so the exponential is one MUFU.EX2), and everything the
// two partners share (w2, z1 log2 e) is computed once.  Algebraically identical to heston_absorb_step /
// heston_fulltrunc_step; rounding differs in the last ulp.  10 instructions + 2 MUFU per path-step.
struct HestonPairF32 {
  float dt, a_v, b_v, xi, c_s, r_s, rho, rho_c, l2e;
};
OPTMC_HD HestonPairF32 heston_pair_consts(const HestonConsts<float>& c) {
  HestonPairF32 f;
  f.dt = c.dt;
  f.a_v = 1.0f - c.kappa * c.dt;        // v + kappa (theta - v) dt = v (1 - kappa dt) + kappa theta dt
  f.b_v = c.kappa * c.theta * c.dt;
  f.xi = c.xi;
  f.l2e = 1.4426950408889634f;
  f.c_s = -0.5f * c.dt * f.l2e;         // exp(x) = 2^(x log2 e)
  f.r_s = c.r * c.dt * f.l2e;
  f.rho = c.rho; f.rho_c = c.rho_c;
  return f;
}
OPTMC_HD float exp2_fast(float x) {
#if defined(__CUDA_ARCH__)
  return mufu_ex2(x);
#else
  return exp2f(x);
#endif
}
// ABSORB: requires vp, vm >= 0 on entry (the stored variance is already truncated, om3:233)
template <bool ABSORB>
OPTMC_HD void heston_pair_step_f32(float& sp, float& vp, float& sm, float& vm, float z1, float z2,
                                   const HestonPairF32& f) {
  const float w2 = fmaf(f.rho, z1, f.rho_c * z2);
  const float z1l = z1 * f.l2e;
  {
    const float v = ABSORB ? vp : rmax(vp, 0.0f);
    const float sq = Real<float>::sqrt_(v * f.dt);
    const float e = fmaf(sq, z1l, fmaf(v, f.c_s, f.r_s));
    const float vn = ABSORB ? fmaf(f.xi * sq, w2, fmaf(v, f.a_v, f.b_v))
                            : fmaf(f.xi * sq, w2, vp + fmaf(v, f.a_v - 1.0f, f.b_v));
    vp = ABSORB ? rmax(vn, 0.0f) : vn;
    sp *= exp2_fast(e);
  }
  {
    const float v = ABSORB ? vm : rmax(vm, 0.0f);
    const float sq = Real<float>::sqrt_(v * f.dt);
    const float e = fmaf(-sq, z1l, fmaf(v, f.c_s, f.r_s));
    const float vn = ABSORB ? fmaf(-(f.xi * sq), w2, fmaf(v, f.a_v, f.b_v))
                            : fmaf(-(f.xi * sq), w2, vm + fmaf(v, f.a_v - 1.0f, f.b_v));
    vm = ABSORB ? rmax(vn, 0.0f) : vn;
    sm *= exp2_fast(e);
  }
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// Packed fp32 (sm_100 f32x2: FFMA2 / FMUL2 / FADD2 -- two IEEE fp32 operations per issue slot).  The fp32 path
// kernel is bound by issue slots and the MUFU pipe, not by HBM, so its production step runs two
// antithetic pairs per instruction.  Each half of a packed operation rounds exactly like the scalar
// fmaf / __fmul_rn / __fadd_rn, so packed and scalar code that use the same operation order agree bit for bit.
// ------------------------------------------------------------------------------------------------
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t f2_splat(float x) { return f2_pack(x, x); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void f2_store4(float* p, f2_t a, f2_t b) {
  float x0, x1, x2, x3;
  f2_unpack(a, x0, x1);
  f2_unpack(b, x2, x3);
  *reinterpret_cast<float4*>(p) = make_float4(x0, x1, x2, x3);
}
// 1.mantissa in [1, 2) from the top 23 bits of a word: (w >> 9) | 0x3f800000 as ONE funnel shift of {0x7f : w}
__device__ __forceinline__ float mant12(uint32_t w) {
  uint32_t r;
  asm("shf.r.clamp.b32 %0, %1, %2, 9;" : "=r"(r) : "r"(w), "r"(0x7fu));
  return __uint_as_float(r);
}

// ---- fp32 Heston draws: one Philox4x32-10 block serves THREE steps ------------------------------------------
// The 128 bits of a block are cut into three 42-bit draws (126 bits used); draw s (s = 0, 1, 2) is
//   bits [42 s, 42 s + 22)        U: 22-bit radius uniform, u = 2 - 1.U in (0, 1], |z| <= 5.52
//   bits [42 s + 22, 42 s + 41)   A: 19-bit angle fraction,  y = 1.A - 3/2 in [-1/2, 1/2)
//   bit   42 s + 41               sign of the cosine
// Box-Muller on the half circle phi = pi y plus the sign bit (the same uniform law on the circle as 2 pi w),
// with sin / cos as degree-4 polynomials in y^2 on the FMA pipe (Chebyshev-node fits, absolute error 2e-7:
// tighter than MUFU.SIN / MUFU.COS) and the radius from MUFU.LG2 + MUFU.SQRT.  Why: measured on B200
// (tools/paths_bench.cu, profiles/paths_bench_r2.txt) the fp32 path step is bound by the FMA pipe -- the two
// IMAD.WIDE of a Philox round issue at quarter rate, 40 of the 93 cycles per warp pair-step -- and by the MUFU
// pipe; three steps per block cut the Philox share by a third and the polynomials take two of the eight MUFU
// operations of a pair-step.  The normals come out pre-scaled by log2(e) (the exponent of the step is evaluated
// as one MUFU.EX2): z1l = z1 log2 e, z2l = z2 log2 e.  The scalar and the packed form use the same operations
// in the same order and agree bit for bit (normals_kernel / the fused European kernel vs the path kernel).
constexpr int kHestonF32Spb = 3;
// 32-bit window of the block's 128-bit string starting at bit START (START < 0: shifted in from below)
template <int START> __device__ __forceinline__ uint32_t philox_window(const Philox4& p) {
  if constexpr (START < 0) {
    return p.v[0] << (-START);
  } else {
    constexpr int W = START >> 5, SH = START & 31;
    if constexpr (SH == 0) return p.v[W];
    const uint32_t lo = p.v[W], hi = W < 3 ? p.v[W + 1] : 0u;
    uint32_t v;
    asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(lo), "r"(hi), "r"(SH));
    return v;
  }
}
template <int S> __device__ __forceinline__ void heston_draw_fields(const Philox4& p, float& fu, float& fa, uint32_t& sgn) {
  const uint32_t wu = philox_window<42 * S - 1>(p);    // U on mantissa bits [22:1]
  const uint32_t wa = philox_window<42 * S + 18>(p);   // A on mantissa bits [22:4], sign on bit 23
  fu = __uint_as_float((wu & 0x007ffffeu) | 0x3f800000u);
  fa = __uint_as_float((wa & 0x007ffff0u) | 0x3f800000u);
  sgn = (wa << 8) & 0x80000000u;
}
// sin / cos polynomials pre-multiplied by sqrt(2 / ln 2): with rad = sqrt(-lg2 u),  rad * PC(t) = log2(e) sqrt(-2 ln u) cos
#define OPTMC_HALFCIRCLE_SIN                                                                 \
  constexpr float PS0 = 5.336446285e+00f, PS1 = -8.778098106e+00f, PS2 = 4.331672668e+00f,     \
                  PS3 = -1.016282082e+00f, PS4 = 1.319097131e-01f;
#define OPTMC_HALFCIRCLE_COS                                                                 \
  constexpr float PC0 = 1.698643446e+00f, PC1 = -8.382454872e+00f, PC2 = 6.893793106e+00f,     \
                  PC3 = -2.262377501e+00f, PC4 = 3.731621504e-01f;
__device__ __forceinline__ float mufu_sqrt_neg(float x) {  // sqrt(-x), x <= 0: the sign is cleared on the ALU pipe (LOP3)
  return mufu_sqrt(__uint_as_float(__float_as_uint(x) & 0x7fffffffu));
}
template <int S> __device__ __forceinline__ void heston_normals_f32_scaled(const Philox4& p, float& z1l, float& z2l) {
  OPTMC_HALFCIRCLE_SIN
  OPTMC_HALFCIRCLE_COS
  float fu, fa;
  uint32_t sgn;
  heston_draw_fields<S>(p, fu, fa, sgn);
  const float u = fmaf(fu, -1.0f, 2.0f);
  const float rad = mufu_sqrt_neg(mufu_lg2(u));
  const float y = __fadd_rn(fa, -1.5f);
  const float t = __fmul_rn(y, y);
  float ps = fmaf(PS4, t, PS3), pc = fmaf(PC4, t, PC3);
  ps = fmaf(ps, t, PS2); pc = fmaf(pc, t, PC2);
  ps = fmaf(ps, t, PS1); pc = fmaf(pc, t, PC1);
  ps = fmaf(ps, t, PS0); pc = fmaf(pc, t, PC0);
  const float sn = __fmul_rn(ps, y);
  const float cs = __uint_as_float(__float_as_uint(pc) ^ sgn);
  z1l = __fmul_rn(rad, cs);
  z2l = __fmul_rn(rad, sn);
}
// unscaled N(0,1) pair of step S of the block (kernels that step with the generic scalar schemes)
template <int S> __device__ __forceinline__ void heston_normals_f32(const Philox4& p, float& z1, float& z2) {
  float a, b;
  heston_normals_f32_scaled<S>(p, a, b);
  z1 = __fmul_rn(a, 0.6931471805599453f);
  z2 = __fmul_rn(b, 0.6931471805599453f);
}

// Production fp32 Euler step of TWO antithetic pairs (i, j): the "+" partners of both pairs are the two halves
// of (sP, uP), the "-" partners those of (sM, uM).  The variance is carried as u = v dt, so sqrt(v dt) is one
// MUFU.SQRT of the state, and with the log2(e)-scaled normals the step is four packed FMAs per partner-pair:
//     sq = +-sqrt(u+)     e  = sq * z1l + u+ * (-log2 e / 2) + r dt log2 e           S *= 2^e   (one MUFU.EX2)
//     xw = xi dt (rho z1 + sqrt(1 - rho^2) z2)          u' = u+ (1 - kappa dt) + kappa theta dt^2 + sq * xw
// (ABSORB: u' = max(u', 0), om3:228-233; otherwise full truncation: u keeps its sign, u+ = max(u, 0) enters).
// The draw feeds the step directly: z2 enters only through xw, so the sine polynomial carries the factor
// xb = sqrt(1 - rho^2) xi dt ln 2 in its coefficients and xw = rad * (xa * cos~ + sin~) costs two packed operations.
// Algebraically heston_absorb_step / heston_fulltrunc_step on the normals heston_normals_f32 reports; rounding
// differs in the last ulps.
struct HestonPairX2 {
  f2_t xa, sb0, sb1, sb2, sb3, sb4, cs, rs, av, bv;
  float inv_dt;
};
__device__ __forceinline__ HestonPairX2 heston_pair_x2_consts(const HestonConsts<float>& c, bool absorb) {
  OPTMC_HALFCIRCLE_SIN
  HestonPairX2 f;
  const float l2e = 1.4426950408889634f, ln2 = 0.6931471805599453f;
  const float xb = c.rho_c * c.xi * c.dt * ln2;
  f.xa = f2_splat(c.rho * c.xi * c.dt * ln2);     // xw from the scaled normals
  f.sb0 = f2_splat(xb * PS0); f.sb1 = f2_splat(xb * PS1); f.sb2 = f2_splat(xb * PS2);
  f.sb3 = f2_splat(xb * PS3); f.sb4 = f2_splat(xb * PS4);
  f.cs = f2_splat(-0.5f * l2e);
  f.rs = f2_splat(c.r * c.dt * l2e);
  f.av = f2_splat(absorb ? 1.0f - c.kappa * c.dt : -(c.kappa * c.dt));
  f.bv = f2_splat(c.kappa * c.theta * c.dt * c.dt);
  f.inv_dt = 1.0f / c.dt;
  return f;
}
// draw S of the blocks of pairs i and j -> z1l = (z1 log2 e) and xw of both pairs (halves = pair i, pair j)
template <int S>
__device__ __forceinline__ void heston_draw_x2(const Philox4& pi, const Philox4& pj, const HestonPairX2& f, f2_t& z1l,
                                               f2_t& xw) {
  OPTMC_HALFCIRCLE_COS
  float fui, fai, fuj, faj;
  uint32_t si, sj;
  heston_draw_fields<S>(pi, fui, fai, si);
  heston_draw_fields<S>(pj, fuj, faj, sj);
  float u0, u1;
  f2_unpack(f2_fma(f2_pack(fui, fuj), f2_splat(-1.0f), f2_splat(2.0f)), u0, u1);
  const f2_t rad = f2_pack(mufu_sqrt_neg(mufu_lg2(u0)), mufu_sqrt_neg(mufu_lg2(u1)));
  const f2_t y = f2_add(f2_pack(fai, faj), f2_splat(-1.5f));
  const f2_t t = f2_mul(y, y);
  f2_t ps = f2_fma(f.sb4, t, f.sb3), pc = f2_fma(f2_splat(PC4), t, f2_splat(PC3));
  ps = f2_fma(ps, t, f.sb2); pc = f2_fma(pc, t, f2_splat(PC2));
  ps = f2_fma(ps, t, f.sb1); pc = f2_fma(pc, t, f2_splat(PC1));
  ps = f2_fma(ps, t, f.sb0); pc = f2_fma(pc, t, f2_splat(PC0));
  const f2_t sb = f2_mul(ps, y);
  float c0, c1;
  f2_unpack(pc, c0, c1);
  const f2_t cs = f2_pack(__uint_as_float(__float_as_uint(c0) ^ si), __uint_as_float(__float_as_uint(c1) ^ sj));
  z1l = f2_mul(rad, cs);
  xw = f2_mul(rad, f2_fma(f.xa, cs, sb));
}
template <bool ABSORB, bool MINUS>
__device__ __forceinline__ void heston_half_step_x2(f2_t& s, f2_t& u, f2_t z1l, f2_t xw, const HestonPairX2& f) {
  float u0, u1;
  f2_unpack(u, u0, u1);
  f2_t up = u;
  if (!ABSORB) { u0 = fmaxf(u0, 0.0f); u1 = fmaxf(u1, 0.0f); up = f2_pack(u0, u1); }
  float q0 = mufu_sqrt(u0), q1 = mufu_sqrt(u1);
  if (MINUS) {  // the antithetic partner sees (-z1, -z2): fold the sign into the root (ALU pipe)
    q0 = __uint_as_float(__float_as_uint(q0) ^ 0x80000000u);
    q1 = __uint_as_float(__float_as_uint(q1) ^ 0x80000000u);
  }
  const f2_t sq = f2_pack(q0, q1);
  const f2_t e = f2_fma(sq, z1l, f2_fma(up, f.cs, f.rs));
  if (ABSORB) {
    float n0, n1;
    f2_unpack(f2_fma(sq, xw, f2_fma(up, f.av, f.bv)), n0, n1);
    u = f2_pack(fmaxf(n0, 0.0f), fmaxf(n1, 0.0f));
  } else {
    u = f2_fma(sq, xw, f2_add(u, f2_fma(up, f.av, f.bv)));
  }
  float e0, e1;
  f2_unpack(e, e0, e1);
  s = f2_mul(s, f2_pack(mufu_ex2(e0), mufu_ex2(e1)));
}
template <bool ABSORB>
__device__ __forceinline__ void heston_pair_step_x2(f2_t& sP, f2_t& uP, f2_t& sM, f2_t& uM, f2_t z1l, f2_t xw,
                                                    const HestonPairX2& f) {
  heston_half_step_x2<ABSORB, false>(sP, uP, z1l, xw, f);
  heston_half_step_x2<ABSORB, true>(sM, uM, z1l, xw, f);
}
#endif  // __CUDACC__

// om3:376-380
template <typename R> OPTMC_HD R payoff(R S, R K, bool is_put) {
  R d = is_put ? K - S : S - K;
  return d > (R)0 ? d : (R)0;
}

// ------------------------------------------------------------------------------------------------
// Polynomial-basis Gram moments and the guarded solve.
//   basis phi_i = x^i, i = 0..DEG, x = S/K.   G_ij = m[i+j], g_i = gy[i].
//   moment vector layout (length 3*DEG+2):  m[0..2DEG] = sum x^k,  then gy[0..DEG] = sum x^k y.
// ------------------------------------------------------------------------------------------------
template <int DEG> struct Moments {
  static constexpr int NM = 2 * DEG + 1;
  static constexpr int NG = DEG + 1;
  static constexpr int Q = NM + NG;
};

// One row (x, y) into the moment vector with the fewest fp64 operations: powers up to DEG explicitly,
// higher moments as one FMA each (x^k = x^DEG * x^(k-DEG)).
template <int DEG> OPTMC_HD void moments_accumulate(double (&acc)[Moments<DEG>::Q], double x, double y) {
  double p[DEG + 1];
  p[0] = 1.0;
#pragma unroll
  for (int k = 1; k <= DEG; ++k) p[k] = p[k - 1] * x;
  acc[0] += 1.0;
#pragma unroll
  for (int k = 1; k <= DEG; ++k) acc[k] += p[k];
#pragma unroll
  for (int k = DEG + 1; k <= 2 * DEG; ++k) acc[k] = fma(p[DEG], p[k - DEG], acc[k]);
  acc[Moments<DEG>::NM] += y;
#pragma unroll
  for (int k = 1; k <= DEG; ++k) acc[Moments<DEG>::NM + k] = fma(p[k], y, acc[Moments<DEG>::NM + k]);
}

// The same row without the count (the caller counts rows with an integer): acc[1..Q) only.
template <int DEG> OPTMC_HD void moments_accumulate_nocount(double (&acc)[Moments<DEG>::Q], double x, double y) {
  double p[DEG + 1];
  p[1] = x;
#pragma unroll
  for (int k = 2; k <= DEG; ++k) p[k] = p[k - 1] * x;
#pragma unroll
  for (int k = 1; k <= DEG; ++k) acc[k] += p[k];
#pragma unroll
  for (int k = DEG + 1; k <= 2 * DEG; ++k) acc[k] = fma(p[DEG], p[k - DEG], acc[k]);
  acc[Moments<DEG>::NM] += y;
#pragma unroll
  for (int k = 1; k <= DEG; ++k) acc[Moments<DEG>::NM + k] = fma(p[k], y, acc[Moments<DEG>::NM + k]);
}

// Power of the regressor carried by moment q: m[k] = sum x^k -> k;  gy[k] = sum x^k y -> k.
template <int DEG> OPTMC_HD int moment_power(int q) { return q < Moments<DEG>::NM ? q : q - Moments<DEG>::NM; }

// ------------------------------------------------------------------------------------------------
// Order-independent cross-CTA summation: a double is split into two non-negative 48-bit fixed-point
// chunks (resolution 2^-52, range |v| < 2^43) that are summed with INTEGER atomics, so the total does
// not depend on arrival order.  Each chunk travels in the low 56 bits of a 64-bit word whose top 8
// bits count arrivals (<= 255 contributors): sum and completion flag are one single-copy-atomic word.
// ------------------------------------------------------------------------------------------------
constexpr int kFxCountShift = 56;
constexpr unsigned long long kFxValueMask = (1ull << kFxCountShift) - 1ull;

OPTMC_HD bool fx_encode(double v, unsigned long long& hi, unsigned long long& lo) {
  const double s = v * 16.0;   // exact
  const double h = floor(s);   // integer part in units of 2^-4
  if (!(fabs(h) < 140737488355328.0)) {  // 2^47; also catches NaN/Inf
    hi = 1ull << 47;
    lo = 0ull;
    return false;
  }
  const double l = (s - h) * 281474976710656.0;  // exact, in [0, 2^48)
  hi = (unsigned long long)((long long)h + (1ll << 47));
  lo = (unsigned long long)l;  // truncation: error < 2^-52
  return true;
}

// sum_hi / sum_lo: the value fields after n contributors were added.
OPTMC_HD double fx_decode(unsigned long long sum_hi, unsigned long long sum_lo, int n) {
  const long long H = (long long)sum_hi - ((long long)n << 47);
  return (double)H * 0.0625 + (double)sum_lo * 2.220446049250313e-16;  // 2^-4, 2^-52
}

#define OPTMC_PIVOT_RTOL 1e-14

// Reciprocal of a positive, normal double.  Device: 20-bit hardware seed + two Newton steps (within an ulp
// of 1/x, a third of the latency of the IEEE division sequence -- the solve sits on the per-date critical path).
OPTMC_HD double pivot_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// LDL^T without pivoting; returns false (no exercise at this date) when n < p or a pivot
// d_k <= 1e-14 * trace(G).  Identical recurrence to oracle.lsm_oracle.cholesky_solve_guarded.
template <int DEG> OPTMC_HD bool solve_poly(const double* mom, double* beta) {
  constexpr int P = DEG + 1;
  constexpr int NM = 2 * DEG + 1;
  if (!(mom[0] >= (double)P)) return false;
  double L[P][P];
  double d[P], rd[P];
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < P; ++i) tr += mom[2 * i];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    double s = mom[2 * k];
#pragma unroll
    for (int j = 0; j < k; ++j) s -= L[k][j] * L[k][j] * d[j];
    if (!(s > OPTMC_PIVOT_RTOL * tr)) return false;
    d[k] = s;
    const double rs = pivot_rcp(s);
    rd[k] = rs;
#pragma unroll
    for (int i = k + 1; i < P; ++i) {
      double u = mom[i + k];
#pragma unroll
      for (int j = 0; j < k; ++j) u -= L[i][j] * L[k][j] * d[j];
      L[i][k] = u * rs;
    }
  }
  double z[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    double a = mom[NM + i];
#pragma unroll
    for (int j = 0; j < i; ++j) a -= L[i][j] * z[j];
    z[i] = a;
  }
#pragma unroll
  for (int i = 0; i < P; ++i) z[i] = z[i] * rd[i];
#pragma unroll
  for (int i = P - 1; i >= 0; --i) {
    double a = z[i];
#pragma unroll
    for (int j = i + 1; j < P; ++j) a -= L[j][i] * beta[j];
    beta[i] = a;
  }
  return true;
}

template <int DEG> OPTMC_HD double poly_eval(const double* beta, double x) {
  double c = beta[DEG];
#pragma unroll
  for (int i = DEG - 1; i >= 0; --i) c = c * x + beta[i];
  return c;
}

// om3:105-121 -- the seven reference features of one path.
template <typename R> OPTMC_HD void features_ref7(R S, R K, R tau_sqrt, R* f) {
  R x = S / K;
  f[0] = (R)1;
  f[1] = x;
  f[2] = x * x;
  f[3] = x * x * x;
  f[4] = rmax(x - (R)1, (R)0);
  f[5] = tau_sqrt;
  f[6] = x * tau_sqrt;
}

}  // namespace optmc
