// european.cu -- K5: European payoff reductions.
//
//  * european_fused_kernel: paths are generated, stepped and reduced entirely in registers; nothing is
//    written to HBM except one (sum, sumsq) pair per block.  Replaces price_european_streaming
//    (om3:382-437: builds and discards a full path array per 500-path chunk), price_european_gpu
//    (om3gpu:605-653) and HestonPricer.price_european_option (hc:259-281, one full simulation per option row).
//    blockIdx.y = option index; option i draws from Philox stream (stream + i).
//  * european_slab_kernel: the same reduction over a stored terminal row.
#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

constexpr int kEuThreads = 256;
constexpr int kEuWarps = kEuThreads / 32;

struct EuArgs {
  long long Mh;  // pairs (or paths when not antithetic)
  int N, anti;
  unsigned long long seed;
  unsigned int stream;
  long long pair_offset;
  double S0, v0, r, sigma, kappa, theta, xi, rho, rho_c;
  const double* par;  // [n_options][6]: K, T, is_put, stream id, N (0: a.N), S0 (0: a.S0)
  double* partials;   // [n_options][gridDim.x][2]
  unsigned int* tickets;  // [n_options]
  double* out;        // [n_options][3]: mean, stderr, n
};

template <typename R, int SCHEME>
__global__ void __launch_bounds__(kEuThreads) european_fused_kernel(const EuArgs a) {
  constexpr bool HES = (SCHEME >= OPTMC_SCHEME_HESTON_REF_ABSORB);
  constexpr bool F32H = HES && sizeof(R) == 4;  // fp32 Heston draws: three steps per Philox block (optmc_math.cuh)
  constexpr int SPB = HES ? (F32H ? kHestonF32Spb : 2) : 4;
  __shared__ double red[kEuWarps * 2];
  __shared__ bool is_last;
  const int opt = blockIdx.y;
  const double K = a.par[opt * 6 + 0], T = a.par[opt * 6 + 1];
  const bool is_put = a.par[opt * 6 + 2] != 0.0;
  const int N = a.par[opt * 6 + 4] > 0.0 ? (int)a.par[opt * 6 + 4] : a.N;  // per-option steps (curve drivers, om3:709)
  const double S0 = a.par[opt * 6 + 5] > 0.0 ? a.par[opt * 6 + 5] : a.S0;
  const double dt = T / N;
  GbmConsts<R> gc;
  gc.drift = (R)((a.r - 0.5 * a.sigma * a.sigma) * dt);
  gc.diffusion = (R)(a.sigma * sqrt(dt));
  HestonConsts<R> hc;
  hc.dt = (R)dt; hc.sqrt_dt = (R)sqrt(dt); hc.r = (R)a.r; hc.kappa = (R)a.kappa; hc.theta = (R)a.theta;
  hc.xi = (R)a.xi; hc.rho = (R)a.rho; hc.rho_c = (R)a.rho_c;
  QeConsts<R> qe{};
  if (SCHEME == OPTMC_SCHEME_HESTON_QE) qe = qe_consts(hc);
  const double df = exp(-a.r * T);
  const unsigned int stream = a.stream + (unsigned int)a.par[opt * 6 + 3];
  const bool anti = a.anti != 0;

  double acc[2] = {0.0, 0.0};
  for (long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x; col < a.Mh;
       col += (long long)gridDim.x * blockDim.x) {
    R sp = (R)S0, sm = (R)S0, vp = (R)a.v0, vm = (R)a.v0;
    for (int t0 = 0; t0 < N; t0 += SPB) {
      R n[6];
      Philox4 p = philox_for((unsigned long long)(a.pair_offset + col), (unsigned int)(t0 / SPB), stream, a.seed);
      if constexpr (F32H) {
        heston_normals_f32<0>(p, n[0], n[1]);
        heston_normals_f32<1>(p, n[2], n[3]);
        heston_normals_f32<2>(p, n[4], n[5]);
      } else {
        Real<R>::normal2(p.v[0], p.v[1], n[0], n[1]);
        Real<R>::normal2(p.v[2], p.v[3], n[2], n[3]);
      }
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        if (t0 + s + 1 > N) break;
        const R z1 = HES ? n[2 * s] : n[s];
        const R z2 = HES ? n[2 * s + 1] : (R)0;
        if (HES) {
          heston_step_any<R, SCHEME>(sp, vp, z1, z2, hc, qe);
          if (anti) heston_step_any<R, SCHEME>(sm, vm, -z1, -z2, hc, qe);
        } else {
          sp = gbm_step<R>(sp, z1, gc);
          if (anti) sm = gbm_step<R>(sm, -z1, gc);
        }
      }
    }
    const double pp = payoff<double>((double)sp, K, is_put) * df;
    acc[0] += pp;
    acc[1] += pp * pp;
    if (anti) {
      const double pm = payoff<double>((double)sm, K, is_put) * df;
      acc[0] += pm;
      acc[1] += pm * pm;
    }
  }
  block_reduce_sum<2, kEuWarps>(acc, red);
  double* part = a.partials + (size_t)opt * gridDim.x * 2;
  if (threadIdx.x == 0) {
    part[blockIdx.x * 2 + 0] = acc[0];
    part[blockIdx.x * 2 + 1] = acc[1];
    __threadfence();
    is_last = (atomicAdd(a.tickets + opt, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) { s[0] += part[b * 2]; s[1] += part[b * 2 + 1]; }
    warp_allreduce_sum<2>(s);
    if (threadIdx.x == 0) {
      const double n = (double)(anti ? 2 * a.Mh : a.Mh);
      const double mean = s[0] / n;
      double var = n > 1.0 ? (s[1] - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      a.out[opt * 3 + 0] = mean;
      a.out[opt * 3 + 1] = sqrt(var / n);
      a.out[opt * 3 + 2] = n;
      a.tickets[opt] = 0u;
    }
  }
}

template <typename R> static void launch_eu_scheme(int scheme, dim3 grid, cudaStream_t st, const EuArgs& a) {
  switch (scheme) {
    case OPTMC_SCHEME_GBM_LOG_EULER:
    case OPTMC_SCHEME_GBM_LOGSPACE:
      european_fused_kernel<R, OPTMC_SCHEME_GBM_LOG_EULER><<<grid, kEuThreads, 0, st>>>(a); break;
    case OPTMC_SCHEME_HESTON_REF_ABSORB:
      european_fused_kernel<R, OPTMC_SCHEME_HESTON_REF_ABSORB><<<grid, kEuThreads, 0, st>>>(a); break;
    case OPTMC_SCHEME_HESTON_FULL_TRUNC:
      european_fused_kernel<R, OPTMC_SCHEME_HESTON_FULL_TRUNC><<<grid, kEuThreads, 0, st>>>(a); break;
    case OPTMC_SCHEME_HESTON_QE:
      european_fused_kernel<R, OPTMC_SCHEME_HESTON_QE><<<grid, kEuThreads, 0, st>>>(a); break;
    default:
      european_fused_kernel<R, OPTMC_SCHEME_HESTON_REF_CALIB><<<grid, kEuThreads, 0, st>>>(a); break;
  }
}

int launch_european_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                          int32_t N, int32_t dtype, int32_t n_options, const double* K, const double* T,
                          const int32_t* is_put, const int32_t* stream_id, optmc_european_result* results,
                          const int32_t* N_opt, const double* S0_opt) {
  if (!mp || !rng || !K || !T || !is_put || !results) { set_error("null argument"); return OPTMC_EINVAL; }
  if (n_options <= 0 || n_options > 65535) { set_error("n_options must be in [1, 65535]"); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (!(mp->S0 > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (rng->z1_dev) { set_error("the fused European kernel generates its own normals"); return OPTMC_EUNSUPPORTED; }
  if (rng->antithetic && (M % 2)) { set_error("antithetic layout needs an even path count"); return OPTMC_EINVAL; }
  if (mp->scheme < 0 || mp->scheme > OPTMC_SCHEME_HESTON_QE) { set_error("unknown scheme"); return OPTMC_EINVAL; }
  for (int i = 0; i < n_options; ++i)
    if (!(K[i] > 0) || !(T[i] > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }

  EuArgs a{};
  a.Mh = rng->antithetic ? M / 2 : M;
  a.N = N; a.anti = rng->antithetic ? 1 : 0; a.seed = rng->seed; a.stream = (unsigned int)rng->stream;
  a.pair_offset = rng->pair_offset;
  a.S0 = mp->S0; a.v0 = mp->v0; a.r = mp->r; a.sigma = mp->sigma; a.kappa = mp->kappa; a.theta = mp->theta;
  a.xi = mp->xi; a.rho = mp->rho; a.rho_c = sqrt(1.0 - mp->rho * mp->rho);

  long long gx = (a.Mh + kEuThreads - 1) / kEuThreads;
  // enough blocks to fill the machine across all options, but no more than ~4 waves
  long long cap = ((long long)ctx->sm_count * 8 * 4 + n_options - 1) / n_options;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  int rc = ensure_bytes((void**)&ctx->eu_par, &ctx->eu_par_cap, (size_t)n_options * 6 * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes((void**)&ctx->eu_out, &ctx->eu_out_cap, (size_t)n_options * 3 * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes((void**)&ctx->partials, &ctx->partials_bytes, (size_t)n_options * gx * 2 * sizeof(double));
  if (rc) return rc;
  if ((size_t)n_options * sizeof(unsigned int) > ctx->eu_tickets_cap) {  // grow-only, zeroed once: kernels reset their ticket
    rc = ensure_bytes((void**)&ctx->eu_tickets, &ctx->eu_tickets_cap, (size_t)n_options * sizeof(unsigned int));
    if (rc) return rc;
    OPTMC_CUDA(cudaMemsetAsync(ctx->eu_tickets, 0, ctx->eu_tickets_cap, ctx->stream));
  }
  unsigned int* tickets = ctx->eu_tickets;
  std::string par(sizeof(double) * 6 * n_options, '\0');
  double* hp = reinterpret_cast<double*>(&par[0]);
  for (int i = 0; i < n_options; ++i) {
    if (N_opt && N_opt[i] <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
    if (S0_opt && !(S0_opt[i] > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
    hp[6 * i] = K[i]; hp[6 * i + 1] = T[i]; hp[6 * i + 2] = is_put[i] ? 1.0 : 0.0;
    hp[6 * i + 3] = (double)(stream_id ? stream_id[i] : i);
    hp[6 * i + 4] = N_opt ? (double)N_opt[i] : 0.0;
    hp[6 * i + 5] = S0_opt ? S0_opt[i] : 0.0;
  }
  OPTMC_CUDA(cudaMemcpyAsync(ctx->eu_par, hp, par.size(), cudaMemcpyHostToDevice, ctx->stream));
  a.par = ctx->eu_par; a.partials = ctx->partials; a.tickets = tickets; a.out = ctx->eu_out;
  dim3 grid((unsigned)gx, (unsigned)n_options);
  if (dtype == OPTMC_F64) launch_eu_scheme<double>(mp->scheme, grid, ctx->stream, a);
  else launch_eu_scheme<float>(mp->scheme, grid, ctx->stream, a);
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  std::string outb(sizeof(double) * 3 * n_options, '\0');
  OPTMC_CUDA(cudaMemcpyAsync(&outb[0], ctx->eu_out, outb.size(), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  const double* ho = reinterpret_cast<const double*>(outb.data());
  for (int i = 0; i < n_options; ++i) {
    results[i].mean = ho[3 * i];
    results[i].stderr_ = ho[3 * i + 1];
    results[i].n_paths = (int64_t)(ho[3 * i + 2] + 0.5);
  }
  return OPTMC_OK;
}

template <typename R>
__global__ void __launch_bounds__(kEuThreads)
european_slab_kernel(const R* __restrict__ ST, long long M, double K, int is_put, double df, double* partials,
                     unsigned int* ticket, double* out) {
  __shared__ double red[kEuWarps * 2];
  __shared__ bool is_last;
  double acc[2] = {0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const double p = payoff<double>((double)ST[j], K, is_put != 0) * df;
    acc[0] += p;
    acc[1] += p * p;
  }
  block_reduce_sum<2, kEuWarps>(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = acc[0];
    partials[blockIdx.x * 2 + 1] = acc[1];
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) { s[0] += partials[b * 2]; s[1] += partials[b * 2 + 1]; }
    warp_allreduce_sum<2>(s);
    if (threadIdx.x == 0) {
      const double n = (double)M;
      const double mean = s[0] / n;
      double var = n > 1.0 ? (s[1] - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      out[0] = mean; out[1] = sqrt(var / n); out[2] = n;
      *ticket = 0u;
    }
  }
}

int launch_european_slab(optmc_ctx* ctx, const void* ST, int64_t M, int32_t dtype, double K, double r, double T,
                         int32_t is_put, optmc_european_result* out) {
  if (!ST || !out || M <= 0) { set_error("bad arguments"); return OPTMC_EINVAL; }
  long long g = (M + kEuThreads * 4 - 1) / (kEuThreads * 4);
  if (g > ctx->sm_count * 4) g = ctx->sm_count * 4;
  int rc = ensure_bytes((void**)&ctx->partials, &ctx->partials_bytes, (size_t)g * 2 * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes((void**)&ctx->eu_out, &ctx->eu_out_cap, 3 * sizeof(double));
  if (rc) return rc;
  const double df = exp(-r * T);
  if (dtype == OPTMC_F64)
    european_slab_kernel<double><<<(unsigned)g, kEuThreads, 0, ctx->stream>>>(static_cast<const double*>(ST), M, K, is_put,
                                                                              df, ctx->partials, ctx->tickets, ctx->eu_out);
  else
    european_slab_kernel<float><<<(unsigned)g, kEuThreads, 0, ctx->stream>>>(static_cast<const float*>(ST), M, K, is_put, df,
                                                                             ctx->partials, ctx->tickets, ctx->eu_out);
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  double h[3];
  OPTMC_CUDA(cudaMemcpyAsync(h, ctx->eu_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  out->mean = h[0]; out->stderr_ = h[1]; out->n_paths = (int64_t)(h[2] + 0.5);
  return OPTMC_OK;
}

}  // namespace optmc
