// paths_bench_variants.cuh -- step / Box-Muller variants measured and REJECTED on the way to the production fp32
// Heston step of csrc/optmc_math.cuh (kept so that tools/paths_bench.cu can reproduce profiles/paths_bench_r2.txt):
// full-circle polynomial sin / cos, polynomial 2^x, two steps per Philox block with unscaled normals.
#pragma once
namespace oldv {
using namespace optmc;
// sin(2 pi x), cos(2 pi x) on x in [-1/2, 1/2) as polynomials in x^2 (Chebyshev-node fit; fp32 Horner error
// 8e-7 absolute, the accuracy class of MUFU.SIN / MUFU.COS) -- on the FMA pipe, two angles per instruction.
#define OPTMC_SINCOS_COEFFS                                                                                    \
  constexpr float S0 = 6.283185005e+00f, S1 = -4.134158707e+01f, S2 = 8.159991455e+01f, S3 = -7.661414337e+01f, \
                  S4 = 4.134136581e+01f, S5 = -1.246881866e+01f;                                               \
  constexpr float C0 = 1.0f, C1 = -1.973920441e+01f, C2 = 6.493911743e+01f, C3 = -8.545011139e+01f,            \
                  C4 = 6.016743088e+01f, C5 = -2.596688461e+01f, C6 = 6.527706146e+00f;
__device__ __forceinline__ void sincos2pi_x2(f2_t x, f2_t& s, f2_t& c) {
  OPTMC_SINCOS_COEFFS
  const f2_t t = f2_mul(x, x);
  f2_t ps = f2_fma(f2_splat(S5), t, f2_splat(S4));
  f2_t pc = f2_fma(f2_splat(C6), t, f2_splat(C5));
  ps = f2_fma(ps, t, f2_splat(S3));
  pc = f2_fma(pc, t, f2_splat(C4));
  ps = f2_fma(ps, t, f2_splat(S2));
  pc = f2_fma(pc, t, f2_splat(C3));
  ps = f2_fma(ps, t, f2_splat(S1));
  pc = f2_fma(pc, t, f2_splat(C2));
  ps = f2_fma(ps, t, f2_splat(S0));
  pc = f2_fma(pc, t, f2_splat(C1));
  s = f2_mul(ps, x);
  c = f2_fma(pc, t, f2_splat(C0));
}
// the same arithmetic, one angle (normals_kernel, the fused European kernel): bit-identical to a packed half
__device__ __forceinline__ void sincos2pi_x1(float x, float& s, float& c) {
  OPTMC_SINCOS_COEFFS
  const float t = __fmul_rn(x, x);
  float ps = fmaf(S5, t, S4), pc = fmaf(C6, t, C5);
  ps = fmaf(ps, t, S3); pc = fmaf(pc, t, C4);
  ps = fmaf(ps, t, S2); pc = fmaf(pc, t, C3);
  ps = fmaf(ps, t, S1); pc = fmaf(pc, t, C2);
  ps = fmaf(ps, t, S0); pc = fmaf(pc, t, C1);
  s = __fmul_rn(ps, x);
  c = fmaf(pc, t, C0);
}

// Box-Muller for two pairs at once: words (a, b) of pair i and pair j -> z1 = (n0_i, n0_j), z2 = (n1_i, n1_j).
// POLY = false reproduces Real<float>::normal2 bit for bit (MUFU.SIN / MUFU.COS); POLY = true takes the angle as
// 2 pi (w - 3/2), w in [1, 2) -- the same uniform on the circle, shifted by half a turn -- and evaluates
// sin / cos on the FMA pipe.
template <bool POLY>
__device__ __forceinline__ void normal2_x2(uint32_t ai, uint32_t bi, uint32_t aj, uint32_t bj, f2_t& z1, f2_t& z2) {
  const f2_t fa = f2_pack(mant12(ai), mant12(aj));
  const f2_t fb = f2_pack(mant12(bi), mant12(bj));
  float u0, u1;
  f2_unpack(f2_fma(fa, f2_splat(-1.0f), f2_splat(2.0f)), u0, u1);  // 2 - fa in (0, 1], exact
  float l0, l1;
  f2_unpack(f2_mul(f2_pack(mufu_lg2(u0), mufu_lg2(u1)), f2_splat(-1.3862943611198906f)), l0, l1);
  const f2_t rad = f2_pack(mufu_sqrt(l0), mufu_sqrt(l1));
  f2_t s, c;
  if (POLY) {
    sincos2pi_x2(f2_add(fb, f2_splat(-1.5f)), s, c);
  } else {
    float g0, g1;
    f2_unpack(f2_mul(fb, f2_splat(6.283185307179586f)), g0, g1);
    s = f2_pack(mufu_sin(g0), mufu_sin(g1));
    c = f2_pack(mufu_cos(g0), mufu_cos(g1));
  }
  z1 = f2_mul(rad, c);
  z2 = f2_mul(rad, s);
}

// 2^e for two exponents on the FMA / ALU pipes: n = rint(e) by the 1.5 * 2^23 trick, a degree-5 polynomial
// of the fraction in [-1/2, 1/2] (relative error 1e-7), exponent insertion by integer shift-add.  |e| < 120.
__device__ __forceinline__ f2_t ex2_poly_x2(f2_t e) {
  const f2_t magic = f2_splat(12582912.0f);
  const f2_t t = f2_add(e, magic);
  const f2_t nf = f2_add(t, f2_splat(-12582912.0f));
  const f2_t f = f2_fma(nf, f2_splat(-1.0f), e);
  f2_t p = f2_fma(f2_splat(1.3390863314e-03f), f, f2_splat(9.6760317683e-03f));  // Chebyshev-node fit of 2^f
  p = f2_fma(p, f, f2_splat(5.5503569543e-02f));
  p = f2_fma(p, f, f2_splat(2.4022106826e-01f));
  p = f2_fma(p, f, f2_splat(6.9314718246e-01f));
  p = f2_fma(p, f, f2_splat(1.0000001192e+00f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  const float r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  const float r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
  return f2_pack(r0, r1);
}

// Production fp32 Euler step of TWO antithetic pairs (i, j): the "+" partners of both pairs are the two halves
// of (sP, uP), the "-" partners those of (sM, uM).  The variance is carried as u = v dt, so sqrt(v dt) is one
// MUFU.SQRT of the state and the per-step constants fold into four packed FMAs per partner-pair:
//     sq = sqrt(u+)           e  = sq * (+-z1 log2 e) + u+ * (-log2 e / 2) + r dt log2 e      S *= 2^e
//     u' = u+ (1 - kappa dt) + kappa theta dt^2 + sq * (+-xi dt w2)      (ABSORB: u' = max(u', 0), om3:228-233)
// Algebraically heston_absorb_step / heston_fulltrunc_step; rounding differs in the last ulps.
struct HestonPairX2 {
  f2_t rho, rho_c, l2e, nl2e, xidt, nxidt, cs, rs, av, bv, inv_dt;
};
__device__ __forceinline__ HestonPairX2 heston_pair_x2_consts(const HestonConsts<float>& c) {
  HestonPairX2 f;
  const float l2e = 1.4426950408889634f;
  f.rho = f2_splat(c.rho); f.rho_c = f2_splat(c.rho_c);
  f.l2e = f2_splat(l2e); f.nl2e = f2_splat(-l2e);
  f.xidt = f2_splat(c.xi * c.dt); f.nxidt = f2_splat(-(c.xi * c.dt));
  f.cs = f2_splat(-0.5f * l2e);
  f.rs = f2_splat(c.r * c.dt * l2e);
  f.av = f2_splat(1.0f - c.kappa * c.dt);
  f.bv = f2_splat(c.kappa * c.theta * c.dt * c.dt);
  f.inv_dt = f2_splat(1.0f / c.dt);
  return f;
}
template <bool ABSORB>
__device__ __forceinline__ void heston_half_step_x2(f2_t& s, f2_t& u, f2_t zl, f2_t xw, const HestonPairX2& f, bool polyex) {
  float u0, u1;
  f2_unpack(u, u0, u1);
  f2_t up = u;
  if (!ABSORB) { u0 = fmaxf(u0, 0.0f); u1 = fmaxf(u1, 0.0f); up = f2_pack(u0, u1); }
  const f2_t sq = f2_pack(mufu_sqrt(u0), mufu_sqrt(u1));
  const f2_t e = f2_fma(sq, zl, f2_fma(up, f.cs, f.rs));
  f2_t un;
  if (ABSORB) {
    un = f2_fma(sq, xw, f2_fma(up, f.av, f.bv));
    float n0, n1;
    f2_unpack(un, n0, n1);
    u = f2_pack(fmaxf(n0, 0.0f), fmaxf(n1, 0.0f));
  } else {
    u = f2_fma(sq, xw, f2_add(u, f2_fma(up, f2_add(f.av, f2_splat(-1.0f)), f.bv)));
  }
  f2_t g;
  if (polyex) {
    g = ex2_poly_x2(e);
  } else {
    float e0, e1;
    f2_unpack(e, e0, e1);
    g = f2_pack(mufu_ex2(e0), mufu_ex2(e1));
  }
  s = f2_mul(s, g);
}
template <bool ABSORB, bool POLYEX_MINUS>
__device__ __forceinline__ void heston_pair_step_x2(f2_t& sP, f2_t& uP, f2_t& sM, f2_t& uM, f2_t z1, f2_t z2,
                                                    const HestonPairX2& f) {
  const f2_t w2 = f2_fma(f.rho, z1, f2_mul(f.rho_c, z2));
  heston_half_step_x2<ABSORB>(sP, uP, f2_mul(z1, f.l2e), f2_mul(w2, f.xidt), f, false);
  heston_half_step_x2<ABSORB>(sM, uM, f2_mul(z1, f.nl2e), f2_mul(w2, f.nxidt), f, POLYEX_MINUS);
}

}  // namespace oldv
