// paths_bench.cu -- development microbenchmark of the fp32 Heston path step (K1) in isolation: the production
// step of csrc/optmc_math.cuh against packed-fp32 (f32x2) / polynomial variants, on the bench workload
// (4 options x 1 M paths x 252 steps, step-major fp32 slabs).  Prints ms, TB/s written and the maximum
// relative difference of each variant's terminal row against variant 0.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/paths_bench tools/paths_bench.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "../options-model_b200/csrc/optmc_math.cuh"
#include "paths_bench_variants.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace optmc;

struct BArgs {
  float* S;
  long long ld, Mh, slab_stride;
  int N;
  unsigned long long seed;
  float S0, v0;
  HestonConsts<float> hc;
};

// ---- variant 0: the round-1 production loop (scalar ops, MUFU sin/cos) ------------------------------------
__global__ void __launch_bounds__(256, 4) k_base(const BArgs a) {
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= a.Mh) return;
  const HestonPairF32 hf = heston_pair_consts(a.hc);
  float sp[4], sm[4], vp[4], vm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { sp[i] = sm[i] = a.S0; vp[i] = vm[i] = fmaxf(a.v0, 0.f); }
  float* Srow = a.S + (size_t)blockIdx.y * a.slab_stride + c0;
  *reinterpret_cast<float4*>(Srow) = make_float4(sp[0], sp[1], sp[2], sp[3]);
  *reinterpret_cast<float4*>(Srow + a.Mh) = make_float4(sp[0], sp[1], sp[2], sp[3]);
  for (int t0 = 0; t0 < a.N; t0 += 2) {
    float nrm[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Philox4 p = philox_for((unsigned long long)(c0 + i), (unsigned)(t0 / 2), blockIdx.y, a.seed);
      Real<float>::normal2(p.v[0], p.v[1], nrm[i][0], nrm[i][1]);
      Real<float>::normal2(p.v[2], p.v[3], nrm[i][2], nrm[i][3]);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (t0 + s + 1 > a.N) break;
      Srow += a.ld;
#pragma unroll
      for (int i = 0; i < 4; ++i) heston_pair_step_f32<true>(sp[i], vp[i], sm[i], vm[i], nrm[i][2 * s], nrm[i][2 * s + 1], hf);
      *reinterpret_cast<float4*>(Srow) = make_float4(sp[0], sp[1], sp[2], sp[3]);
      *reinterpret_cast<float4*>(Srow + a.Mh) = make_float4(sm[0], sm[1], sm[2], sm[3]);
    }
  }
}

// ---- packed variants -----------------------------------------------------------------------------------
// MODE bit 0: polynomial sin/cos (FMA pipe) instead of MUFU.SIN/COS    bit 4: ... for packed group 0 only
// MODE bit 1: polynomial ex2 for the antithetic partner
// MODE bit 2: no stores (compute only);  bit 3: stores only (no compute)
// NPG: packed groups (2 antithetic pairs each) per thread;  MINB: CTAs per SM promised to the compiler
template <int MODE, int NPG, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_packed(const BArgs a) {
  const bool do_store = !(MODE & 4) || a.N < 0;
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (2 * NPG);
  if (c0 >= a.Mh) return;
  const oldv::HestonPairX2 hx = oldv::heston_pair_x2_consts(a.hc);
  f2_t sP[NPG], sM[NPG], uP[NPG], uM[NPG];
  {
    const float u0 = fmaxf(a.v0, 0.f) * a.hc.dt;
#pragma unroll
    for (int h = 0; h < NPG; ++h) { sP[h] = sM[h] = f2_pack(a.S0, a.S0); uP[h] = uM[h] = f2_pack(u0, u0); }
  }
  float* Srow = a.S + (size_t)blockIdx.y * a.slab_stride + c0;
  auto store = [&](float* p, const f2_t (&x)[NPG]) {
    if constexpr (NPG == 2) f2_store4(p, x[0], x[1]);
    else {
#pragma unroll
      for (int h = 0; h < NPG; ++h) *reinterpret_cast<f2_t*>(p + 2 * h) = x[h];
    }
  };
  store(Srow, sP);
  store(Srow + a.Mh, sM);
  for (int t0 = 0; t0 < a.N; t0 += 2) {
    Philox4 p[2 * NPG];
    if (!(MODE & 8))
#pragma unroll
    for (int i = 0; i < 2 * NPG; ++i) p[i] = philox_for((unsigned long long)(c0 + i), (unsigned)(t0 / 2), blockIdx.y, a.seed);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (t0 + s + 1 > a.N) break;
      Srow += a.ld;
      if (MODE & 8) {
        store(Srow, sP);
        store(Srow + a.Mh, sM);
        continue;
      }
#pragma unroll
      for (int h = 0; h < NPG; ++h) {
        f2_t z1, z2;
        if ((MODE & 1) || ((MODE & 16) && h == 0))
          oldv::normal2_x2<true>(p[2 * h].v[2 * s], p[2 * h].v[2 * s + 1], p[2 * h + 1].v[2 * s], p[2 * h + 1].v[2 * s + 1], z1, z2);
        else
          oldv::normal2_x2<false>(p[2 * h].v[2 * s], p[2 * h].v[2 * s + 1], p[2 * h + 1].v[2 * s], p[2 * h + 1].v[2 * s + 1], z1, z2);
        oldv::heston_pair_step_x2<true, (MODE & 2) != 0>(sP[h], uP[h], sM[h], uM[h], z1, z2, hx);
      }
      if (do_store) {
        store(Srow, sP);
        store(Srow + a.Mh, sM);
      }
    }
  }
  if (!do_store) {  // keep the computation alive
    store(a.S + (size_t)blockIdx.y * a.slab_stride + c0, sP);
    store(a.S + (size_t)blockIdx.y * a.slab_stride + c0 + a.Mh, sM);
  }
}

// ---- 3 steps per Philox block: six 21-bit uniforms from the 128 bits ---------------------------------------
__device__ __forceinline__ float mant21(const Philox4& p, int k) {  // field k = bits [21k, 21k + 21) -> [1, 2)
  const int lsb = 21 * k - 2;  // window start so that the field lands on mantissa bits [22:2]
  uint32_t v;
  if (lsb < 0) {
    v = p.v[0] << 2;
  } else {
    const int w = lsb >> 5, sh = lsb & 31;
    const uint32_t lo = p.v[w], hi = w < 3 ? p.v[w + 1] : 0u;
    asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(lo), "r"(hi), "r"(sh));
  }
  return __uint_as_float((v & 0x007ffffcu) | 0x3f800000u);
}
template <bool POLY>
__device__ __forceinline__ void normal2_x2_f(float fai, float fbi, float faj, float fbj, f2_t& z1, f2_t& z2) {
  const f2_t fa = f2_pack(fai, faj);
  const f2_t fb = f2_pack(fbi, fbj);
  float u0, u1;
  f2_unpack(f2_fma(fa, f2_splat(-1.0f), f2_splat(2.0f)), u0, u1);
  float l0, l1;
  f2_unpack(f2_mul(f2_pack(mufu_lg2(u0), mufu_lg2(u1)), f2_splat(-1.3862943611198906f)), l0, l1);
  const f2_t rad = f2_pack(mufu_sqrt(l0), mufu_sqrt(l1));
  f2_t s, c;
  if (POLY) {
    oldv::sincos2pi_x2(f2_add(fb, f2_splat(-1.5f)), s, c);
  } else {
    float g0, g1;
    f2_unpack(f2_mul(fb, f2_splat(6.283185307179586f)), g0, g1);
    s = f2_pack(mufu_sin(g0), mufu_sin(g1));
    c = f2_pack(mufu_cos(g0), mufu_cos(g1));
  }
  z1 = f2_mul(rad, c);
  z2 = f2_mul(rad, s);
}
// MODE bit 0: poly sincos for all groups; bit 4: for group 0 only; bit 5: for steps s == 0 only; ROUNDS: Philox rounds
template <int MODE, int NPG, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_spb3(const BArgs a) {
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (2 * NPG);
  if (c0 >= a.Mh) return;
  const oldv::HestonPairX2 hx = oldv::heston_pair_x2_consts(a.hc);
  f2_t sP[NPG], sM[NPG], uP[NPG], uM[NPG];
  {
    const float u0 = fmaxf(a.v0, 0.f) * a.hc.dt;
#pragma unroll
    for (int h = 0; h < NPG; ++h) { sP[h] = sM[h] = f2_pack(a.S0, a.S0); uP[h] = uM[h] = f2_pack(u0, u0); }
  }
  float* Srow = a.S + (size_t)blockIdx.y * a.slab_stride + c0;
  auto store = [&](float* p, const f2_t (&x)[NPG]) {
    if constexpr (NPG == 2) f2_store4(p, x[0], x[1]);
    else {
#pragma unroll
      for (int h = 0; h < NPG; ++h) *reinterpret_cast<f2_t*>(p + 2 * h) = x[h];
    }
  };
  store(Srow, sP);
  store(Srow + a.Mh, sM);
  for (int t0 = 0; t0 < a.N; t0 += 3) {
    Philox4 p[2 * NPG];
#pragma unroll
    for (int i = 0; i < 2 * NPG; ++i) p[i] = philox_for((unsigned long long)(c0 + i), (unsigned)(t0 / 3), blockIdx.y, a.seed);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (t0 + s + 1 > a.N) break;
      Srow += a.ld;
#pragma unroll
      for (int h = 0; h < NPG; ++h) {
        f2_t z1, z2;
        const float fai = mant21(p[2 * h], 2 * s), fbi = mant21(p[2 * h], 2 * s + 1);
        const float faj = mant21(p[2 * h + 1], 2 * s), fbj = mant21(p[2 * h + 1], 2 * s + 1);
        if ((MODE & 1) || ((MODE & 16) && h == 0) || ((MODE & 32) && s == 0) || ((MODE & 64) && s != 1))
          normal2_x2_f<true>(fai, fbi, faj, fbj, z1, z2);
        else
          normal2_x2_f<false>(fai, fbi, faj, fbj, z1, z2);
        oldv::heston_pair_step_x2<true, false>(sP[h], uP[h], sM[h], uM[h], z1, z2, hx);
      }
      store(Srow, sP);
      store(Srow + a.Mh, sM);
    }
  }
}

// ---- the production step (csrc/optmc_math.cuh): 3 steps per block, half-circle polynomials, scaled normals ----
template <int NPG, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_prod(const BArgs a) {
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (2 * NPG);
  if (c0 >= a.Mh) return;
  const HestonPairX2 hx = heston_pair_x2_consts(a.hc, true);
  f2_t sP[NPG], sM[NPG], uP[NPG], uM[NPG];
  {
    const float u0 = fmaxf(a.v0, 0.f) * a.hc.dt;
#pragma unroll
    for (int h = 0; h < NPG; ++h) { sP[h] = sM[h] = f2_pack(a.S0, a.S0); uP[h] = uM[h] = f2_pack(u0, u0); }
  }
  float* Srow = a.S + (size_t)blockIdx.y * a.slab_stride + c0;
  auto store = [&](float* p, const f2_t (&x)[NPG]) {
    if constexpr (NPG == 2) f2_store4(p, x[0], x[1]);
    else {
#pragma unroll
      for (int h = 0; h < NPG; ++h) *reinterpret_cast<f2_t*>(p + 2 * h) = x[h];
    }
  };
  store(Srow, sP);
  store(Srow + a.Mh, sM);
  auto step = [&](auto s_tag, const Philox4 (&p)[2 * NPG]) {
    constexpr int S = decltype(s_tag)::value;
    Srow += a.ld;
#pragma unroll
    for (int h = 0; h < NPG; ++h) {
      f2_t z1l, xw;
      heston_draw_x2<S>(p[2 * h], p[2 * h + 1], hx, z1l, xw);
      heston_pair_step_x2<true>(sP[h], uP[h], sM[h], uM[h], z1l, xw, hx);
    }
    store(Srow, sP);
    store(Srow + a.Mh, sM);
  };
  for (int t0 = 0; t0 < a.N; t0 += 3) {
    Philox4 p[2 * NPG];
#pragma unroll
    for (int i = 0; i < 2 * NPG; ++i) p[i] = philox_for((unsigned long long)(c0 + i), (unsigned)(t0 / 3), blockIdx.y, a.seed);
    step(std::integral_constant<int, 0>{}, p);
    if (t0 + 2 > a.N) break;
    step(std::integral_constant<int, 1>{}, p);
    if (t0 + 3 > a.N) break;
    step(std::integral_constant<int, 2>{}, p);
  }
}

// ---- Philox alone: 4 blocks per thread per 2 steps, as the path kernels consume them ---------------------
// PV 0: the production philox4x32_10 (compiles to IMAD.WIDE)   PV 1: mul.hi + mul.lo kept apart (asm volatile)
// PV 2: 7 rounds                                                 PV 3: no Philox (a counter hash of 4 ALU ops)
template <int PV>
__global__ void __launch_bounds__(256, 4) k_philox(const BArgs a) {
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= a.Mh) return;
  uint32_t acc[4] = {0, 0, 0, 0};
  for (int t0 = 0; t0 < a.N; t0 += 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t x0 = (uint32_t)(c0 + i), x1 = 0, x2 = (uint32_t)(t0 / 2), x3 = blockIdx.y;
      uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
      if (PV == 3) {
        x0 = (x0 ^ k0) * 0x9E3779B9u + x2; x1 = x0 ^ (x0 >> 15); x2 = x1 * 0x85EBCA6Bu; x3 = x2 ^ (x2 >> 13);
      } else {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
        for (int r = 0; r < (PV == 2 ? 7 : 10); ++r) {
          uint32_t hi0, lo0, hi1, lo1;
          if (PV == 1) {
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi0) : "r"(M0), "r"(x0));
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo0) : "r"(M0), "r"(x0));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi1) : "r"(M1), "r"(x2));
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(lo1) : "r"(M1), "r"(x2));
          } else {
            hi0 = __umulhi(M0, x0); lo0 = M0 * x0; hi1 = __umulhi(M1, x2); lo1 = M1 * x2;
          }
          const uint32_t n0 = hi1 ^ x1 ^ k0, n2 = hi0 ^ x3 ^ k1;
          x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
          k0 += W0; k1 += W1;
        }
      }
      acc[0] ^= x0; acc[1] ^= x1; acc[2] ^= x2; acc[3] ^= x3;
    }
  }
  float* Srow = a.S + (size_t)blockIdx.y * a.slab_stride + c0;
  *reinterpret_cast<uint4*>(Srow) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
}

typedef void (*kern_t)(const BArgs);

int main(int argc, char** argv) {
  const long long M = argc > 1 ? atoll(argv[1]) : 1000000;
  const int N = argc > 2 ? atoi(argv[2]) : 252;
  const int G = argc > 3 ? atoi(argv[3]) : 4;
  const int reps = argc > 4 ? atoi(argv[4]) : 10;   // timed launches per variant (profiling runs use 1)
  const int warm = argc > 5 ? atoi(argv[5]) : 3;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const long long ld = (M + 63) / 64 * 64, Mh = M / 2;
  const size_t slab = (size_t)(N + 1) * ld;
  float* S;
  CK(cudaMalloc(&S, slab * G * sizeof(float)));
  BArgs a{};
  a.S = S; a.ld = ld; a.Mh = Mh; a.slab_stride = (long long)slab; a.N = N; a.seed = 42; a.S0 = 100.f; a.v0 = 0.04f;
  const double dt = 1.0 / N;
  a.hc.dt = (float)dt; a.hc.sqrt_dt = (float)sqrt(dt); a.hc.r = 0.05f; a.hc.kappa = 2.f; a.hc.theta = 0.04f; a.hc.xi = 0.5f;
  a.hc.rho = -0.7f; a.hc.rho_c = (float)sqrt(1 - 0.49);
  printf("M %lld N %d G %d\n", M, N, G);
  struct V { const char* name; kern_t k; int npg, nt, minb; int tuned = 0; };
  const V vs[] = {{"base (round 1)", k_base, 2, 256, 4},
                  {"packed", k_packed<0, 2, 256, 4>, 2, 256, 4},
                  {"packed + poly sincos", k_packed<1, 2, 256, 4>, 2, 256, 4},
                  {"packed + poly sincos(half)", k_packed<16, 2, 256, 4>, 2, 256, 4},
                  {"packed + poly sincos(half), 5 CTA/SM", k_packed<16, 2, 256, 5>, 2, 256, 5},
                  {"packed + poly sincos, 5 CTA/SM", k_packed<1, 2, 256, 5>, 2, 256, 5},
                  {"packed, 5 CTA/SM", k_packed<0, 2, 256, 5>, 2, 256, 5},
                  {"packed NPG1 6 CTA/SM", k_packed<0, 1, 256, 6>, 1, 256, 6},
                  {"packed NPG1 + poly sincos 6 CTA/SM", k_packed<1, 1, 256, 6>, 1, 256, 6},
                  {"packed NPG1 8 CTA/SM", k_packed<0, 1, 256, 8>, 1, 256, 8},
                  {"packed NPG1 + poly sincos 8 CTA/SM", k_packed<1, 1, 256, 8>, 1, 256, 8},
                  {"packed NPG3 + poly sincos 3 CTA/SM", k_packed<1, 3, 256, 3>, 3, 256, 3},
                  {"packed NPG3 3 CTA/SM", k_packed<0, 3, 256, 3>, 3, 256, 3},
                  {"spb3 packed", k_spb3<0, 2, 256, 4>, 2, 256, 4},
                  {"spb3 packed + poly sincos", k_spb3<1, 2, 256, 4>, 2, 256, 4},
                  {"spb3 packed + poly sincos(group 0)", k_spb3<16, 2, 256, 4>, 2, 256, 4},
                  {"spb3 packed + poly sincos(1 step of 3)", k_spb3<32, 2, 256, 4>, 2, 256, 4},
                  {"spb3 packed + poly sincos(2 steps of 3)", k_spb3<64, 2, 256, 4>, 2, 256, 4},
                  {"spb3 packed NPG1 6 CTA/SM", k_spb3<0, 1, 256, 6>, 1, 256, 6},
                  {"spb3 packed NPG1 + poly(1 of 3) 6 CTA/SM", k_spb3<32, 1, 256, 6>, 1, 256, 6},
                  {"spb3 packed NPG3 + poly(group 0) 3 CTA/SM", k_spb3<16, 3, 256, 3>, 3, 256, 3},
                  {"spb3 packed NPG3 + poly(2 of 3) 3 CTA/SM", k_spb3<64, 3, 256, 3>, 3, 256, 3},
                  {"PRODUCTION step", k_prod<2, 256, 4>, 2, 256, 4},
                  {"PRODUCTION step NPG3 3 CTA/SM", k_prod<3, 256, 3>, 3, 256, 3},
                  {"PRODUCTION step NPG1 6 CTA/SM", k_prod<1, 256, 6>, 1, 256, 6},
                  {"PRODUCTION step NPG4 2 CTA/SM", k_prod<4, 256, 2>, 4, 256, 2},
                  {"PRODUCTION step NPG3 128thr 6 CTA/SM", k_prod<3, 128, 6>, 3, 128, 6},
                  {"PRODUCTION step NPG2 128thr 8 CTA/SM", k_prod<2, 128, 8>, 2, 128, 8},
                  {"PRODUCTION step NPG2 128thr 8 CTA/SM tuned grid", k_prod<2, 128, 8>, 2, 128, 8, 1},
                  {"PRODUCTION step NPG2 64thr 16 CTA/SM", k_prod<2, 64, 16>, 2, 64, 16},
                  {"PRODUCTION step NPG2 64thr 16 CTA/SM tuned grid", k_prod<2, 64, 16>, 2, 64, 16, 1},
                  {"PRODUCTION step NPG2 256thr 4 CTA/SM tuned grid", k_prod<2, 256, 4>, 2, 256, 4, 1},
                  {"packed, no stores", k_packed<4, 2, 256, 4>, 2, 256, 4},
                  {"philox only (IMAD.WIDE)", k_philox<0>, 2, 256, 4},
                  {"philox only (mul.hi + mul.lo)", k_philox<1>, 2, 256, 4},
                  {"philox only, 7 rounds", k_philox<2>, 2, 256, 4},
                  {"cheap hash only", k_philox<3>, 2, 256, 4},
                  {"stores only", k_packed<8, 2, 256, 4>, 2, 256, 4}};
  std::vector<float> ref(M), cur(M);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int v = 0; v < (int)(sizeof(vs) / sizeof(vs[0])); ++v) {
    const long long units = Mh / (2 * vs[v].npg);
    long long per_opt = ((long long)prop.multiProcessorCount * vs[v].minb + G - 1) / G;
    long long tpc = (units + per_opt - 1) / per_opt;
    if (tpc > vs[v].nt) tpc = vs[v].nt;
    if (tpc < 64) tpc = 64;
    if (vs[v].tuned) {  // threads per CTA minimising the busiest SM's lane count: ceil(CTAs / SMs) * roundup32(tpc)
      long long best = -1;
      for (long long c = vs[v].nt / 2 + 1; c <= vs[v].nt; ++c) {
        const long long ctas = (units + c - 1) / c * G;
        const long long cost = (ctas + prop.multiProcessorCount - 1) / prop.multiProcessorCount * ((c + 31) / 32 * 32);
        if (best < 0 || cost < best) { best = cost; tpc = c; }
      }
    }
    dim3 grid((unsigned)((units + tpc - 1) / tpc), (unsigned)G);
    if (vs[v].tuned) printf("   [tuned: %lld threads per CTA, %u CTAs]\n", tpc, grid.x * grid.y);
    for (int w = 0; w < warm; ++w) vs[v].k<<<grid, (unsigned)tpc>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) vs[v].k<<<grid, (unsigned)tpc>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    CK(cudaMemcpy(cur.data(), S + (size_t)N * ld, M * sizeof(float), cudaMemcpyDeviceToHost));
    double mx = 0, mean = 0, m2 = 0;
    if (v == 0) ref = cur;
    for (long long i = 0; i < M; ++i) {
      mx = fmax(mx, fabs((double)cur[i] - ref[i]) / ref[i]);
      mean += cur[i]; m2 += (double)cur[i] * cur[i];
    }
    mean /= M;
    printf("%-42s %.4f ms  %.2f TB/s written  mean S_T %.5f sd %.4f  max rel diff vs base %.3e\n", vs[v].name, ms,
           (double)G * (N + 1) * M * 4 / ms / 1e9, mean, sqrt(m2 / M - mean * mean), mx);
  }
  return 0;
}
