// paths.cu -- K1: path generation.  One thread owns VEC antithetic pairs for all N steps; the Philox
// counter, Box-Muller normals, the variance and the running price live in registers; the only HBM
// traffic is one coalesced (128-bit when aligned) store of S[t][j] per path per step.
//
// Replaces: om3:473-480 (GBM), om3:211-233 (Heston, absorption Euler), om3gpu:117-248 (fp32 variants),
// hc:204-257 (calibrator scheme, step-major here).
#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

#include <type_traits>

namespace optmc {

struct PathArgs {
  void* S;
  void* V;
  long long ld, Mh;
  int N, anti;
  const void* z1;
  const void* z2;
  int z_f64;
  unsigned long long seed;
  unsigned int stream;
  long long pair_offset;
  double S0, v0;
  double g_drift, g_diff;
  double dt, sqrt_dt, r, kappa, theta, xi, rho, rho_c;
};

// Normals of one Philox block.  GBM: four z (steps 4b+1 .. 4b+4).  Heston: (z1, z2) pairs, n[2s] / n[2s+1] --
// two steps per block in fp64 (Box-Muller on 32-bit words), THREE in fp32 (optmc_math.cuh: heston_normals_f32).
template <typename R> struct HestonDraws {
  static constexpr int SPB = 2;
  static __device__ __forceinline__ void normals(const Philox4& p, R (&n)[6]) {
    Real<R>::normal2(p.v[0], p.v[1], n[0], n[1]);
    Real<R>::normal2(p.v[2], p.v[3], n[2], n[3]);
    n[4] = n[5] = (R)0;
  }
};
template <> struct HestonDraws<float> {
  static constexpr int SPB = kHestonF32Spb;
  static __device__ __forceinline__ void normals(const Philox4& p, float (&n)[6]) {
    heston_normals_f32<0>(p, n[0], n[1]);
    heston_normals_f32<1>(p, n[2], n[3]);
    heston_normals_f32<2>(p, n[4], n[5]);
  }
};
template <typename R, bool HES>
__device__ __forceinline__ void philox_block_normals(unsigned long long pair, unsigned int blk, unsigned int stream,
                                                     unsigned long long seed, R (&n)[6]) {
  const Philox4 p = philox_for(pair, blk, stream, seed);
  if (HES) {
    HestonDraws<R>::normals(p, n);
  } else {
    Real<R>::normal2(p.v[0], p.v[1], n[0], n[1]);
    Real<R>::normal2(p.v[2], p.v[3], n[2], n[3]);
    n[4] = n[5] = (R)0;
  }
}

__device__ __forceinline__ double load_z(const void* z, int is_f64, long long idx) {
  return is_f64 ? static_cast<const double*>(z)[idx] : (double)static_cast<const float*>(z)[idx];
}

template <typename R, int SCHEME> struct FastPair { static constexpr bool value = false; };
template <> struct FastPair<float, OPTMC_SCHEME_HESTON_REF_ABSORB> { static constexpr bool value = true; };
template <> struct FastPair<float, OPTMC_SCHEME_HESTON_FULL_TRUNC> { static constexpr bool value = true; };

template <typename R, int SCHEME, int VEC, bool EXTZ>
__device__ __forceinline__ void paths_body(const PathArgs& a) {
  constexpr bool HES = (SCHEME >= OPTMC_SCHEME_HESTON_REF_ABSORB);
  constexpr bool LOGSPACE = (SCHEME == OPTMC_SCHEME_GBM_LOGSPACE);
  constexpr bool FAST = FastPair<R, SCHEME>::value;  // fp32 Heston: both partners of a pair in one fused step
  constexpr int SPB = HES ? HestonDraws<R>::SPB : 4;  // steps served by one Philox block
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (c0 >= a.Mh) return;
  const bool anti = a.anti != 0;

  GbmConsts<R> gc;
  gc.drift = (R)a.g_drift;
  gc.diffusion = (R)a.g_diff;
  HestonConsts<R> hc;
  hc.dt = (R)a.dt; hc.sqrt_dt = (R)a.sqrt_dt; hc.r = (R)a.r; hc.kappa = (R)a.kappa; hc.theta = (R)a.theta;
  hc.xi = (R)a.xi; hc.rho = (R)a.rho; hc.rho_c = (R)a.rho_c;
  HestonPairF32 hf{};
  if constexpr (FAST) hf = heston_pair_consts(hc);
  QeConsts<R> qe{};
  if constexpr (SCHEME == OPTMC_SCHEME_HESTON_QE) qe = qe_consts(hc);

  R sp[VEC], sm[VEC], vp[VEC], vm[VEC], out[VEC];
  const R s_init = LOGSPACE ? (R)log(fmax(a.S0, 1e-12)) : (R)a.S0;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    sp[i] = s_init; sm[i] = s_init; vp[i] = (R)a.v0; vm[i] = (R)a.v0;
  }
  R* Srow = static_cast<R*>(a.S) + c0;
  R* Vrow = a.V ? static_cast<R*>(a.V) + c0 : nullptr;
#pragma unroll
  for (int i = 0; i < VEC; ++i) out[i] = LOGSPACE ? Real<R>::exp_(sp[i]) : sp[i];
  VecIO<R, VEC>::store(Srow, out);
  if (anti) VecIO<R, VEC>::store(Srow + a.Mh, out);
  if (HES && Vrow) {
    VecIO<R, VEC>::store(Vrow, vp);
    if (anti) VecIO<R, VEC>::store(Vrow + a.Mh, vp);
  }
  if (FAST && SCHEME == OPTMC_SCHEME_HESTON_REF_ABSORB) {  // the fused step assumes the truncated state (om3:229)
#pragma unroll
    for (int i = 0; i < VEC; ++i) { vp[i] = rmax(vp[i], (R)0); vm[i] = rmax(vm[i], (R)0); }
  }

  for (int t0 = 0; t0 < a.N; t0 += SPB) {
    R nrm[VEC][6];
    if (!EXTZ) {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        philox_block_normals<R, HES>((unsigned long long)(a.pair_offset + c0 + i), (unsigned int)(t0 / SPB), a.stream,
                                     a.seed, nrm[i]);
    }
#pragma unroll
    for (int s = 0; s < SPB; ++s) {
      const int t = t0 + s + 1;
      if (t > a.N) break;
      R z1[VEC], z2[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (EXTZ) {
          const long long idx = (long long)(t - 1) * a.Mh + c0 + i;
          z1[i] = (R)load_z(a.z1, a.z_f64, idx);
          z2[i] = HES ? (R)load_z(a.z2, a.z_f64, idx) : (R)0;
        } else {
          z1[i] = HES ? nrm[i][2 * s] : nrm[i][s];
          z2[i] = HES ? nrm[i][2 * s + 1] : (R)0;
        }
      }
      Srow += a.ld;
      if (Vrow) Vrow += a.ld;
      if constexpr (FAST) {
        if (anti) {
#pragma unroll
          for (int i = 0; i < VEC; ++i)
            heston_pair_step_f32<SCHEME == OPTMC_SCHEME_HESTON_REF_ABSORB>(sp[i], vp[i], sm[i], vm[i], z1[i], z2[i], hf);
          VecIO<R, VEC>::store(Srow, sp);
          VecIO<R, VEC>::store(Srow + a.Mh, sm);
          if (Vrow) {
            VecIO<R, VEC>::store(Vrow, vp);
            VecIO<R, VEC>::store(Vrow + a.Mh, vm);
          }
          continue;
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        if (HES) heston_step_any<R, SCHEME>(sp[i], vp[i], z1[i], z2[i], hc, qe);
        else if (LOGSPACE) sp[i] = sp[i] + (gc.drift + gc.diffusion * z1[i]);
        else sp[i] = gbm_step<R>(sp[i], z1[i], gc);
        out[i] = LOGSPACE ? Real<R>::exp_(sp[i]) : sp[i];
      }
      VecIO<R, VEC>::store(Srow, out);
      if (HES && Vrow) VecIO<R, VEC>::store(Vrow, vp);
      if (anti) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          if (HES) heston_step_any<R, SCHEME>(sm[i], vm[i], -z1[i], -z2[i], hc, qe);
          else if (LOGSPACE) sm[i] = sm[i] + (gc.drift - gc.diffusion * z1[i]);
          else sm[i] = gbm_step<R>(sm[i], -z1[i], gc);
          out[i] = LOGSPACE ? Real<R>::exp_(sm[i]) : sm[i];
        }
        VecIO<R, VEC>::store(Srow + a.Mh, out);
        if (HES && Vrow) VecIO<R, VEC>::store(Vrow + a.Mh, vm);
      }
    }
  }
}

// ---- production fp32 Heston generation (schemes REF_ABSORB / FULL_TRUNC, antithetic, Philox draws) --------------
// One thread owns FOUR antithetic pairs as two packed groups (optmc_math.cuh: heston_draw_x2 / heston_pair_step_x2):
// every floating-point operation of the step is an f32x2 instruction on two pairs, the variance is carried as
// u = v dt, one Philox block feeds three steps.  128-thread CTAs, 8 per SM (64 registers).  Measured on B200
// (tools/paths_bench.cu): 1.02 ms for 4 x 1 M x 252 = 3.96 TB/s written, against 1.33 ms for the scalar step above;
// the kernel is bound by the FMA pipe (quarter-rate IMAD.WIDE of Philox + the packed FMAs) and the MUFU pipe.
constexpr int kFastThreads = 128;
// HASV: the variance slab is written too (callers that ask for V); without it the kernel carries no second row pointer
// and no per-step branch -- two registers less in a kernel that spills at its 64-register budget.
template <int SCHEME, bool HASV>
__device__ __forceinline__ void paths_body_x2(const PathArgs& a) {
  constexpr bool ABSORB = SCHEME == OPTMC_SCHEME_HESTON_REF_ABSORB;
  const long long c0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= a.Mh) return;
  HestonConsts<float> hc;
  hc.dt = (float)a.dt; hc.sqrt_dt = (float)a.sqrt_dt; hc.r = (float)a.r; hc.kappa = (float)a.kappa;
  hc.theta = (float)a.theta; hc.xi = (float)a.xi; hc.rho = (float)a.rho; hc.rho_c = (float)a.rho_c;
  const HestonPairX2 hx = heston_pair_x2_consts(hc, ABSORB);
  f2_t sP[2], sM[2], uP[2], uM[2];
  float* Srow = static_cast<float*>(a.S) + c0;
  float* Vrow = HASV ? static_cast<float*>(a.V) + c0 : nullptr;
  {
    const float s0 = (float)a.S0, v0 = (float)a.v0;
    const float u0 = (ABSORB ? fmaxf(v0, 0.0f) : v0) * hc.dt;  // the step assumes the truncated state (om3:229)
#pragma unroll
    for (int h = 0; h < 2; ++h) { sP[h] = sM[h] = f2_splat(s0); uP[h] = uM[h] = f2_splat(u0); }
    f2_store4(Srow, sP[0], sP[1]);
    f2_store4(Srow + a.Mh, sP[0], sP[1]);
    if (HASV) {
      f2_store4(Vrow, f2_splat(v0), f2_splat(v0));
      f2_store4(Vrow + a.Mh, f2_splat(v0), f2_splat(v0));
    }
  }
  const f2_t inv_dt = f2_splat(hx.inv_dt);
  auto step = [&](auto s_tag, const Philox4 (&p)[4]) {
    constexpr int S = decltype(s_tag)::value;
    Srow += a.ld;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      f2_t z1l, xw;
      heston_draw_x2<S>(p[2 * h], p[2 * h + 1], hx, z1l, xw);
      heston_pair_step_x2<ABSORB>(sP[h], uP[h], sM[h], uM[h], z1l, xw, hx);
    }
    f2_store4(Srow, sP[0], sP[1]);
    f2_store4(Srow + a.Mh, sM[0], sM[1]);
    if (HASV) {
      Vrow += a.ld;
      f2_store4(Vrow, f2_mul(uP[0], inv_dt), f2_mul(uP[1], inv_dt));
      f2_store4(Vrow + a.Mh, f2_mul(uM[0], inv_dt), f2_mul(uM[1], inv_dt));
    }
  };
  for (int t0 = 0; t0 < a.N; t0 += kHestonF32Spb) {
    Philox4 p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      p[i] = philox_for((unsigned long long)(a.pair_offset + c0 + i), (unsigned int)(t0 / kHestonF32Spb), a.stream, a.seed);
    step(std::integral_constant<int, 0>{}, p);
    if (t0 + 2 > a.N) break;
    step(std::integral_constant<int, 1>{}, p);
    if (t0 + 3 > a.N) break;
    step(std::integral_constant<int, 2>{}, p);
  }
}
template <int SCHEME, bool HASV> __global__ void __launch_bounds__(kFastThreads, 8) paths_x2_kernel(const PathArgs a) {
  paths_body_x2<SCHEME, HASV>(a);
}
template <int SCHEME> __global__ void __launch_bounds__(kFastThreads, 8) paths_x2_batch_kernel(const PathArgs* __restrict__ args) {
  const PathArgs a = args[blockIdx.y];
  paths_body_x2<SCHEME, false>(a);  // batches never store the variance
}
// Threads per CTA (<= 128) for `units` threads per option and G options: the count that minimises the busiest SM's
// lane count, ceil(CTAs / SMs) * roundup32(threads); ties go to the larger CTA.
static unsigned fast_block(optmc_ctx* ctx, long long units, int G) {
  long long best = -1;
  unsigned tpc = kFastThreads;
  for (long long c = kFastThreads; c >= 48; --c) {
    const long long ctas = (units + c - 1) / c * G;
    const long long cost = (ctas + ctx->sm_count - 1) / ctx->sm_count * ((c + 31) / 32 * 32);
    if (best < 0 || cost < best) { best = cost; tpc = (unsigned)c; }
  }
  return tpc;
}
static bool fast_eligible(int scheme, int dtype, const PathArgs& a, bool extz, bool vec4) {
  return dtype == OPTMC_F32 && !extz && vec4 && a.anti &&
         (scheme == OPTMC_SCHEME_HESTON_REF_ABSORB || scheme == OPTMC_SCHEME_HESTON_FULL_TRUNC);
}

template <typename R, int SCHEME, int VEC, bool EXTZ>
__global__ void __launch_bounds__(256) paths_kernel(const PathArgs a) {
  paths_body<R, SCHEME, VEC, EXTZ>(a);
}

// Batched form: blockIdx.y selects the option (its own slab, S0, dt, step count and Philox stream).
template <typename R, int SCHEME, int VEC>
__global__ void __launch_bounds__(256, 4) paths_batch_kernel(const PathArgs* __restrict__ args) {
  const PathArgs a = args[blockIdx.y];
  paths_body<R, SCHEME, VEC, false>(a);
}

// Grid shape: one wave of (4 CTAs per SM) with the work split evenly across the CTAs, so every SM carries the
// same number of threads (a plain ceil(units / 256) grid leaves some SMs with 4 CTAs and others with 3).
static void path_grid(optmc_ctx* ctx, long long units, unsigned* grid, unsigned* block) {
  const long long ctas = (long long)ctx->sm_count * 4;
  long long tpc = (units + ctas - 1) / ctas;
  if (tpc > 256) tpc = 256;  // several waves
  if (tpc < 64) tpc = 64;
  *block = (unsigned)tpc;
  *grid = (unsigned)((units + tpc - 1) / tpc);
}

template <typename R, int SCHEME, int VEC> static int launch_t3(optmc_ctx* ctx, const PathArgs& a, bool extz) {
  const long long threads = (a.Mh + VEC - 1) / VEC;
  unsigned grid, block;
  path_grid(ctx, threads, &grid, &block);
  if (extz) paths_kernel<R, SCHEME, VEC, true><<<grid, block, 0, ctx->stream>>>(a);
  else paths_kernel<R, SCHEME, VEC, false><<<grid, block, 0, ctx->stream>>>(a);
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

template <typename R, int SCHEME> static int launch_t2(optmc_ctx* ctx, const PathArgs& a, bool extz, bool vec4) {
  return vec4 ? launch_t3<R, SCHEME, 4>(ctx, a, extz) : launch_t3<R, SCHEME, 1>(ctx, a, extz);
}

template <typename R> static int launch_t1(optmc_ctx* ctx, int scheme, const PathArgs& a, bool extz, bool vec4) {
  switch (scheme) {
    case OPTMC_SCHEME_GBM_LOG_EULER: return launch_t2<R, OPTMC_SCHEME_GBM_LOG_EULER>(ctx, a, extz, vec4);
    case OPTMC_SCHEME_GBM_LOGSPACE: return launch_t2<R, OPTMC_SCHEME_GBM_LOGSPACE>(ctx, a, extz, vec4);
    case OPTMC_SCHEME_HESTON_REF_ABSORB: return launch_t2<R, OPTMC_SCHEME_HESTON_REF_ABSORB>(ctx, a, extz, vec4);
    case OPTMC_SCHEME_HESTON_FULL_TRUNC: return launch_t2<R, OPTMC_SCHEME_HESTON_FULL_TRUNC>(ctx, a, extz, vec4);
    case OPTMC_SCHEME_HESTON_REF_CALIB: return launch_t2<R, OPTMC_SCHEME_HESTON_REF_CALIB>(ctx, a, extz, vec4);
    case OPTMC_SCHEME_HESTON_QE: return launch_t2<R, OPTMC_SCHEME_HESTON_QE>(ctx, a, extz, vec4);
  }
  set_error("unknown scheme");
  return OPTMC_EINVAL;
}

static int fill_path_args(const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int32_t N, PathArgs* a,
                          bool* extz) {
  if (!mp || !rng) { set_error("null params"); return OPTMC_EINVAL; }
  if (!(mp->S0 > 0) || !(mp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  const bool hes = mp->scheme >= OPTMC_SCHEME_HESTON_REF_ABSORB;
  if (hes != (mp->model == OPTMC_MODEL_HESTON)) { set_error("scheme does not belong to model"); return OPTMC_EINVAL; }
  if (rng->antithetic && (M % 2)) { set_error("antithetic layout needs an even path count"); return OPTMC_EINVAL; }
  if (hes && !(mp->rho > -1.0 && mp->rho < 1.0)) { set_error("rho must be in (-1, 1)"); return OPTMC_EINVAL; }
  if (!hes && !(mp->sigma >= 0)) { set_error("sigma must be non-negative"); return OPTMC_EINVAL; }
  *extz = rng->z1_dev != nullptr;
  if (*extz && hes && !rng->z2_dev) { set_error("Heston external normals need z1 and z2"); return OPTMC_EINVAL; }
  const double dt = mp->T / N;
  a->N = N;
  a->anti = rng->antithetic ? 1 : 0;
  a->Mh = rng->antithetic ? M / 2 : M;
  a->z1 = rng->z1_dev; a->z2 = rng->z2_dev; a->z_f64 = rng->z_dtype == OPTMC_F64;
  a->seed = rng->seed; a->stream = (unsigned int)rng->stream; a->pair_offset = rng->pair_offset;
  a->S0 = mp->S0; a->v0 = mp->v0;
  a->g_drift = (mp->r - 0.5 * mp->sigma * mp->sigma) * dt;
  a->g_diff = mp->sigma * sqrt(dt);
  a->dt = dt; a->sqrt_dt = sqrt(dt); a->r = mp->r; a->kappa = mp->kappa; a->theta = mp->theta; a->xi = mp->xi;
  a->rho = mp->rho; a->rho_c = sqrt(1.0 - mp->rho * mp->rho);
  return OPTMC_OK;
}

int launch_paths(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int32_t N,
                 int32_t dtype, void* S, void* V, int64_t ld) {
  PathArgs a{};
  bool extz = false;
  int rc = fill_path_args(mp, rng, M, N, &a, &extz);
  if (rc) return rc;
  if (!S) { set_error("S_dev is null"); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  a.S = S; a.V = V; a.ld = ld;
  const size_t es = dtype == OPTMC_F64 ? 8 : 4;
  bool vec4 = (a.Mh % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)S % 16 == 0) && (!V || (uintptr_t)V % 16 == 0) &&
              ((a.Mh * es) % 16 == 0);
  if (fast_eligible(mp->scheme, dtype, a, extz, vec4)) {
    const long long units = a.Mh / 4;
    const unsigned block = fast_block(ctx, units, 1);
    const unsigned grid = (unsigned)((units + block - 1) / block);
    const bool absorb = mp->scheme == OPTMC_SCHEME_HESTON_REF_ABSORB;
    if (a.V) {
      if (absorb) paths_x2_kernel<OPTMC_SCHEME_HESTON_REF_ABSORB, true><<<grid, block, 0, ctx->stream>>>(a);
      else paths_x2_kernel<OPTMC_SCHEME_HESTON_FULL_TRUNC, true><<<grid, block, 0, ctx->stream>>>(a);
    } else {
      if (absorb) paths_x2_kernel<OPTMC_SCHEME_HESTON_REF_ABSORB, false><<<grid, block, 0, ctx->stream>>>(a);
      else paths_x2_kernel<OPTMC_SCHEME_HESTON_FULL_TRUNC, false><<<grid, block, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
    return OPTMC_OK;
  }
  if (dtype == OPTMC_F64) return launch_t1<double>(ctx, mp->scheme, a, extz, vec4);
  return launch_t1<float>(ctx, mp->scheme, a, extz, vec4);
}

// ---- the normals the Philox path kernels consume (test aid) ---------------------------------------------
template <typename R> __global__ void normals_kernel(PathArgs a, int hes, int which, R* Z) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.Mh) return;
  const int spb = hes ? HestonDraws<R>::SPB : 4;
  for (int t0 = 0; t0 < a.N; t0 += spb) {
    R n[6];
    if (hes) philox_block_normals<R, true>((unsigned long long)(a.pair_offset + col), (unsigned int)(t0 / spb), a.stream, a.seed, n);
    else philox_block_normals<R, false>((unsigned long long)(a.pair_offset + col), (unsigned int)(t0 / spb), a.stream, a.seed, n);
    for (int s = 0; s < spb; ++s) {
      const int t = t0 + s + 1;
      if (t > a.N) break;
      Z[(long long)(t - 1) * a.Mh + col] = hes ? n[2 * s + which] : n[s];
    }
  }
}

int launch_philox_normals(optmc_ctx* ctx, const optmc_rng_params* rng, int32_t model, int64_t M, int32_t N,
                          int32_t which, int32_t dtype, void* Z) {
  if (!rng || !Z || M <= 0 || N <= 0 || which < 0 || which > 1) { set_error("bad arguments"); return OPTMC_EINVAL; }
  const int hes = model == OPTMC_MODEL_HESTON;
  if (!hes && which != 0) { set_error("GBM has a single normal stream"); return OPTMC_EINVAL; }
  PathArgs a{};
  a.N = N; a.Mh = rng->antithetic ? M / 2 : M; a.seed = rng->seed; a.stream = (unsigned int)rng->stream;
  a.pair_offset = rng->pair_offset;
  const unsigned grid = (unsigned)((a.Mh + 255) / 256);
  if (dtype == OPTMC_F64) normals_kernel<double><<<grid, 256, 0, ctx->stream>>>(a, hes, which, static_cast<double*>(Z));
  else normals_kernel<float><<<grid, 256, 0, ctx->stream>>>(a, hes, which, static_cast<float*>(Z));
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

// ---- Philox KAT ------------------------------------------------------------------------------------------
__global__ void philox_kat_kernel(int n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 p = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1]);
  for (int k = 0; k < 4; ++k) out[4 * i + k] = p.v[k];
}

int launch_philox_kat(optmc_ctx* ctx, int n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  if (n <= 0 || !ctr || !key || !out) { set_error("bad arguments"); return OPTMC_EINVAL; }
  uint32_t* d = nullptr;
  OPTMC_CUDA(cudaMalloc(&d, sizeof(uint32_t) * 10 * n));
  cudaError_t e = cudaMemcpyAsync(d, ctr, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + 4 * n, key, sizeof(uint32_t) * 2 * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    philox_kat_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, d, d + 4 * n, d + 6 * n);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + 6 * n, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "philox_kat");
  return OPTMC_OK;
}

// ---- create_regression_features (om3:105-121) ----------------------------------------------------------------
template <typename R> __global__ void features_kernel(const R* S, long long n, R K, R tau_sqrt, R* F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  R f[7];
  features_ref7<R>(S[i], K, tau_sqrt, f);
#pragma unroll
  for (int k = 0; k < 7; ++k) F[i * 7 + k] = f[k];
}

int launch_features(optmc_ctx* ctx, const void* S, int64_t n, int32_t dtype, double K, double T, double t_current,
                    void* F) {
  if (!S || !F || n < 0) { set_error("bad arguments"); return OPTMC_EINVAL; }
  if (n == 0) return OPTMC_OK;
  const double tau = T - t_current;
  const double tau_sqrt = sqrt(tau > 1e-6 ? tau : 1e-6);  // om3:109
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == OPTMC_F64)
    features_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double*>(S), n, K, tau_sqrt,
                                                           static_cast<double*>(F));
  else
    features_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(S), n, (float)K, (float)tau_sqrt,
                                                          static_cast<float*>(F));
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}


// ---- batched generation: G options, one slab each (optmc_price_american_batch) ---------------------------
template <typename R, int VEC> static void launch_batch_scheme(int scheme, dim3 grid, unsigned block, cudaStream_t st,
                                                               const PathArgs* d) {
  switch (scheme) {
    case OPTMC_SCHEME_GBM_LOG_EULER: paths_batch_kernel<R, OPTMC_SCHEME_GBM_LOG_EULER, VEC><<<grid, block, 0, st>>>(d); break;
    case OPTMC_SCHEME_GBM_LOGSPACE: paths_batch_kernel<R, OPTMC_SCHEME_GBM_LOGSPACE, VEC><<<grid, block, 0, st>>>(d); break;
    case OPTMC_SCHEME_HESTON_REF_ABSORB: paths_batch_kernel<R, OPTMC_SCHEME_HESTON_REF_ABSORB, VEC><<<grid, block, 0, st>>>(d); break;
    case OPTMC_SCHEME_HESTON_FULL_TRUNC: paths_batch_kernel<R, OPTMC_SCHEME_HESTON_FULL_TRUNC, VEC><<<grid, block, 0, st>>>(d); break;
    case OPTMC_SCHEME_HESTON_QE: paths_batch_kernel<R, OPTMC_SCHEME_HESTON_QE, VEC><<<grid, block, 0, st>>>(d); break;
    default: paths_batch_kernel<R, OPTMC_SCHEME_HESTON_REF_CALIB, VEC><<<grid, block, 0, st>>>(d); break;
  }
}

size_t path_args_bytes() { return sizeof(PathArgs); }

// Two halves so that the caller can time the kernel alone: prepare = fill the per-option descriptors and copy them
// to the device, launch = the batched generation kernel.
int prepare_paths_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int G,
                        const optmc_american_option* opts, void* slab, size_t slab_stride_bytes, int64_t ld, void* d_args,
                        void* h_args) {
  PathArgs* h = static_cast<PathArgs*>(h_args);
  for (int g = 0; g < G; ++g) {
    optmc_model_params m = *mp;
    m.S0 = opts[g].S0; m.T = opts[g].T;
    optmc_rng_params r = *rng;
    r.stream = rng->stream + opts[g].stream;
    r.z1_dev = nullptr; r.z2_dev = nullptr;
    bool extz = false;
    int rc = fill_path_args(&m, &r, M, opts[g].N, &h[g], &extz);
    if (rc) return rc;
    h[g].S = static_cast<char*>(slab) + (size_t)g * slab_stride_bytes;
    h[g].V = nullptr;
    h[g].ld = ld;
  }
  OPTMC_CUDA(cudaMemcpyAsync(d_args, h, sizeof(PathArgs) * G, cudaMemcpyHostToDevice, ctx->stream));
  return OPTMC_OK;
}

int launch_paths_batch(optmc_ctx* ctx, const optmc_model_params* mp, int32_t dtype, int G, void* slab,
                       size_t slab_stride_bytes, int64_t ld, void* d_args, void* h_args) {
  const PathArgs* h = static_cast<const PathArgs*>(h_args);
  const long long Mh = h[0].Mh;
  const bool vec4 = (Mh % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)slab % 16 == 0) && (slab_stride_bytes % 16 == 0);
  const long long units = vec4 ? Mh / 4 : Mh;
  const PathArgs* d = static_cast<const PathArgs*>(d_args);
  if (fast_eligible(mp->scheme, dtype, h[0], false, vec4)) {
    const unsigned block = fast_block(ctx, units, G);
    dim3 grid((unsigned)((units + block - 1) / block), (unsigned)G);
    if (mp->scheme == OPTMC_SCHEME_HESTON_REF_ABSORB) paths_x2_batch_kernel<OPTMC_SCHEME_HESTON_REF_ABSORB><<<grid, block, 0, ctx->stream>>>(d);
    else paths_x2_batch_kernel<OPTMC_SCHEME_HESTON_FULL_TRUNC><<<grid, block, 0, ctx->stream>>>(d);
    ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
    return OPTMC_OK;
  }
  // G options share the machine: split each option's work over ~ (4 CTAs per SM) / G blocks
  long long per_opt = ((long long)ctx->sm_count * 4 + G - 1) / G;
  if (per_opt < 1) per_opt = 1;
  long long tpc = (units + per_opt - 1) / per_opt;
  if (tpc > 256) tpc = 256;
  if (tpc < 64) tpc = 64;
  dim3 grid((unsigned)((units + tpc - 1) / tpc), (unsigned)G);
  if (dtype == OPTMC_F64) {
    if (vec4) launch_batch_scheme<double, 4>(mp->scheme, grid, (unsigned)tpc, ctx->stream, d);
    else launch_batch_scheme<double, 1>(mp->scheme, grid, (unsigned)tpc, ctx->stream, d);
  } else {
    if (vec4) launch_batch_scheme<float, 4>(mp->scheme, grid, (unsigned)tpc, ctx->stream, d);
    else launch_batch_scheme<float, 1>(mp->scheme, grid, (unsigned)tpc, ctx->stream, d);
  }
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

}  // namespace optmc
