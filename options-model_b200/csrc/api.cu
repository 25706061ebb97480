// api.cu -- the extern "C" boundary declared in include/optmc.h: context, workspaces, argument
// validation (the reference's ValueError conditions, om3:447-452), dispatch and result read-back.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "optmc_internal.h"

namespace optmc {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return e == cudaErrorMemoryAllocation ? OPTMC_ENOMEM : OPTMC_ECUDA;
}

int ensure_bytes(void** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return OPTMC_OK;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  if (need < 256) need = 256;
  OPTMC_CUDA(cudaMalloc(p, need));
  *cap = need;
  return OPTMC_OK;
}

int ensure_per_date(optmc_ctx* ctx, int N) {
  const size_t need = (size_t)N + 1;
  if (need <= ctx->per_date_cap) return OPTMC_OK;
  cudaFree(ctx->d_betas); cudaFree(ctx->d_bnd); cudaFree(ctx->d_exc); cudaFree(ctx->d_nitm); cudaFree(ctx->d_valid);
  ctx->d_betas = nullptr; ctx->d_bnd = nullptr; ctx->d_exc = nullptr; ctx->d_nitm = nullptr; ctx->d_valid = nullptr;
  ctx->per_date_cap = 0;
  const size_t cap = need < 512 ? 512 : need;
  OPTMC_CUDA(cudaMalloc((void**)&ctx->d_betas, cap * kMaxBeta * sizeof(double)));
  OPTMC_CUDA(cudaMalloc((void**)&ctx->d_bnd, cap * sizeof(unsigned long long)));
  OPTMC_CUDA(cudaMalloc((void**)&ctx->d_exc, cap * sizeof(unsigned long long)));
  OPTMC_CUDA(cudaMalloc((void**)&ctx->d_nitm, cap * sizeof(long long)));
  OPTMC_CUDA(cudaMalloc((void**)&ctx->d_valid, cap * sizeof(int)));
  ctx->per_date_cap = cap;
  return OPTMC_OK;
}

static int validate_lsm(const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype, const optmc_lsm_params* lp) {
  if (!S || !lp) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(lp->K > 0) || !(lp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (lp->r < 0) { set_error("r must be non-negative."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  if (lp->basis != OPTMC_BASIS_POLY2 && lp->basis != OPTMC_BASIS_POLY3) { set_error("basis must be POLY2 or POLY3"); return OPTMC_EINVAL; }
  if (lp->impl < OPTMC_SWEEP_AUTO || lp->impl > OPTMC_SWEEP_SPLIT) { set_error("bad sweep impl"); return OPTMC_EINVAL; }
  return OPTMC_OK;
}

static int bind_sweep(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                      const optmc_lsm_params* lp) {
  int rc = validate_lsm(S, ld, M, N, dtype, lp);
  if (rc) return rc;
  rc = ensure_per_date(ctx, N);
  if (rc) return rc;
  SweepDesc& sw = ctx->sw;
  sw = SweepDesc{};
  sw.S = S; sw.ld = ld; sw.M = M; sw.N = N; sw.dtype = dtype; sw.lp = *lp;
  sw.deg = lp->basis == OPTMC_BASIS_POLY3 ? 3 : 2;
  const double dt = lp->T / N;
  sw.disc = exp(-lp->r * dt);
  sw.final_scale = (lp->semantics & OPTMC_SEM_REF_DISCOUNT) ? 1.0 : sw.disc;
  sw.Dt.assign((size_t)N + 1, 1.0);
  sw.Dinv.assign((size_t)N + 1, 1.0);
  {
    const double inv_disc = 1.0 / sw.disc;
    double d = 1.0, di = 1.0;
    for (int t = N; t >= 0; --t) { sw.Dt[t] = d; sw.Dinv[t] = di; d *= sw.disc; di *= inv_disc; }
  }
  sw.Kh = lp->K; sw.Kl = 0.0;
  if (dtype == OPTMC_F32) { sw.Kh = (double)(float)lp->K; sw.Kl = (double)(float)(lp->K - sw.Kh); }
  return OPTMC_OK;
}

static int run_sweep(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  std::string why;
  const bool can_res = resident_eligible(ctx, sw, &why);
  int impl = sw.lp.impl;
  if (impl == OPTMC_SWEEP_AUTO) impl = can_res ? OPTMC_SWEEP_RESIDENT : OPTMC_SWEEP_SPLIT;
  if (impl == OPTMC_SWEEP_RESIDENT) {
    if (!can_res) { set_error("resident sweep unavailable: " + why); return OPTMC_EUNSUPPORTED; }
    int rc = sweep_resident(ctx);
    if (rc) return rc;
  } else {
    int rc = sweep_begin(ctx);
    if (rc) return rc;
    for (int t = sw.N - 1; t >= 1; --t) {
      rc = sweep_gram_date(ctx, t, ctx->gram);
      if (rc) return rc;
      rc = sweep_update_date(ctx, t, ctx->gram);
      if (rc) return rc;
    }
    rc = sweep_finish(ctx, ctx->gram);
    if (rc) return rc;
    rc = sweep_finalize_price(ctx, ctx->gram);
    if (rc) return rc;
  }
  sw.impl_used = impl;
  sw.have_results = true;
  return OPTMC_OK;
}

static int fetch_results(optmc_ctx* ctx, optmc_lsm_result* out) {
  SweepDesc& sw = ctx->sw;
  if (!out) { set_error("null result"); return OPTMC_EINVAL; }
  if (!sw.have_results) { set_error("no sweep results to fetch"); return OPTMC_EINVAL; }
  const int n1 = sw.N + 1;
  const int p = sw.deg + 1;
  double fin[4];
  std::vector<double> hb;
  std::vector<unsigned long long> hbnd, hexc;
  std::vector<long long> hn;
  if (sw.impl_used == OPTMC_SWEEP_RESIDENT) {
    // The persistent sweep sums Gram moments in fixed point (|moment| < 2^43).  On overflow (or NaN/Inf
    // prices) AUTO repeats the sweep with the split kernels; an explicit RESIDENT request fails loudly.
    int flags[4] = {0, 0, 0, 0};
    OPTMC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flags[0]) {
      if (sw.lp.impl != OPTMC_SWEEP_AUTO) {
        set_error("resident sweep: a Gram moment left the fixed-point exchange range (|m| < 2^43) or is not finite; use impl = SPLIT");
        return OPTMC_EUNSUPPORTED;
      }
      const int launches = sw.n_launches;
      sw.lp.impl = OPTMC_SWEEP_SPLIT;
      int rc = run_sweep(ctx);
      sw.lp.impl = OPTMC_SWEEP_AUTO;
      if (rc) return rc;
      sw.n_launches += launches;
    }
  }
  OPTMC_CUDA(cudaMemcpyAsync(fin, ctx->d_final, sizeof(fin), cudaMemcpyDeviceToHost, ctx->stream));
  if (out->betas) { hb.resize((size_t)n1 * kMaxBeta); OPTMC_CUDA(cudaMemcpyAsync(hb.data(), ctx->d_betas, hb.size() * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->boundary) { hbnd.resize(n1); OPTMC_CUDA(cudaMemcpyAsync(hbnd.data(), ctx->d_bnd, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->ex_count) { hexc.resize(n1); OPTMC_CUDA(cudaMemcpyAsync(hexc.data(), ctx->d_exc, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->n_itm) { hn.resize(n1); OPTMC_CUDA(cudaMemcpyAsync(hn.data(), ctx->d_nitm, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  out->price = fin[0];
  out->stderr_ = fin[1];
  out->n_paths = sw.M;
  out->impl_used = sw.impl_used;
  out->n_launches = sw.n_launches;
  if (out->betas)
    for (int t = 0; t < n1; ++t)
      for (int i = 0; i < p; ++i) out->betas[(size_t)t * p + i] = hb[(size_t)t * kMaxBeta + i];
  if (out->boundary) {
    const unsigned long long none = sw.lp.is_put ? 0ull : ~0ull;
    for (int t = 0; t < n1; ++t) {
      if (hbnd[t] == none) out->boundary[t] = nan("");
      else memcpy(&out->boundary[t], &hbnd[t], 8);
    }
  }
  if (out->ex_count) for (int t = 0; t < n1; ++t) out->ex_count[t] = (int64_t)hexc[t];
  if (out->n_itm) for (int t = 0; t < n1; ++t) out->n_itm[t] = (int64_t)hn[t];
  return OPTMC_OK;
}

}  // namespace optmc

using namespace optmc;

#define OPTMC_TRY_BEGIN try {
#define OPTMC_TRY_END                                          \
  }                                                            \
  catch (const std::bad_alloc&) { set_error("host allocation failed"); return OPTMC_ENOMEM; } \
  catch (const std::exception& e) { set_error(e.what()); return OPTMC_ECUDA; }

static int use_device(optmc_ctx* ctx) {
  if (!ctx) { set_error("null context"); return OPTMC_EINVAL; }
  OPTMC_CUDA(cudaSetDevice(ctx->device));
  return OPTMC_OK;
}
#define OPTMC_ENTER(ctx)            \
  do {                              \
    int _rc = use_device(ctx);      \
    if (_rc) return _rc;            \
  } while (0)

extern "C" {

int optmc_abi_version(void) { return OPTMC_ABI_VERSION; }

const char* optmc_last_error(void) { return g_last_error.c_str(); }

int optmc_ctx_create(int device, optmc_ctx** out) {
  OPTMC_TRY_BEGIN
  if (!out) { set_error("null out pointer"); return OPTMC_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device available (there is no CPU fallback): ") +
              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    cudaGetLastError();
    return OPTMC_ECUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return OPTMC_EINVAL; }
  OPTMC_CUDA(cudaSetDevice(device));
  optmc_ctx* c = new optmc_ctx();
  c->device = device;
  cudaDeviceProp prop;
  OPTMC_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = prop.l2CacheSize;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->cc = prop.major * 10 + prop.minor;
  OPTMC_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  OPTMC_CUDA(cudaMalloc((void**)&c->tickets, 1024 * sizeof(unsigned int)));
  OPTMC_CUDA(cudaMemset(c->tickets, 0, 1024 * sizeof(unsigned int)));
  OPTMC_CUDA(cudaMalloc((void**)&c->gram, 16 * sizeof(double)));
  OPTMC_CUDA(cudaMalloc((void**)&c->d_final, 4 * sizeof(double)));
  OPTMC_CUDA(cudaMalloc(&c->xchg, xchg_bytes()));
  OPTMC_CUDA(cudaMemset(c->xchg, 0, xchg_bytes()));
  for (int i = 0; i < 3; ++i) OPTMC_CUDA(cudaEventCreate(&c->ev[i]));
  OPTMC_CUDA(cudaMalloc((void**)&c->d_flags, 4 * sizeof(int)));
  OPTMC_CUDA(cudaMemset(c->d_flags, 0, 4 * sizeof(int)));
  *out = c;
  return OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_ctx_destroy(optmc_ctx* ctx) {
  if (!ctx) return OPTMC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->slab); cudaFree(ctx->cf); cudaFree(ctx->partials); cudaFree(ctx->tickets); cudaFree(ctx->gram);
  cudaFree(ctx->d_betas); cudaFree(ctx->d_bnd); cudaFree(ctx->d_exc); cudaFree(ctx->d_nitm); cudaFree(ctx->d_valid);
  cudaFree(ctx->d_final); cudaFree(ctx->xchg); cudaFree(ctx->d_flags); cudaFree(ctx->batch_dev); cudaFree(ctx->eu_out); cudaFree(ctx->eu_par); cudaFree(ctx->eu_tickets);
  for (int i = 0; i < 3; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return OPTMC_OK;
}

int optmc_ctx_set_stream(optmc_ctx* ctx, void* cuda_stream) {
  if (!ctx) { set_error("null context"); return OPTMC_EINVAL; }
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return OPTMC_OK;
}

int optmc_ctx_synchronize(optmc_ctx* ctx) {
  OPTMC_ENTER(ctx);
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPTMC_OK;
}

int64_t optmc_ctx_launch_count(optmc_ctx* ctx) { return ctx ? ctx->launches : -1; }

int optmc_ctx_kernel_times(optmc_ctx* ctx, double* paths_ms, double* sweep_ms) {
  if (!ctx) { set_error("null context"); return OPTMC_EINVAL; }
  if (paths_ms) *paths_ms = ctx->last_paths_ms;
  if (sweep_ms) *sweep_ms = ctx->last_sweep_ms;
  return OPTMC_OK;
}

int optmc_ctx_device_info(optmc_ctx* ctx, int64_t out[4]) {
  if (!ctx || !out) { set_error("null argument"); return OPTMC_EINVAL; }
  out[0] = ctx->sm_count; out[1] = ctx->l2_bytes; out[2] = ctx->max_smem_optin; out[3] = ctx->cc;
  return OPTMC_OK;
}

int optmc_paths_gbm(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int32_t N,
                    int32_t dtype, void* S_dev, int64_t ld) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!mp || mp->model != OPTMC_MODEL_GBM) { set_error("optmc_paths_gbm needs a GBM model"); return OPTMC_EINVAL; }
  return launch_paths(ctx, mp, rng, M, N, dtype, S_dev, nullptr, ld);
  OPTMC_TRY_END
}

int optmc_paths_heston(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                       int32_t N, int32_t dtype, void* S_dev, void* V_dev, int64_t ld) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!mp || mp->model != OPTMC_MODEL_HESTON) { set_error("optmc_paths_heston needs a Heston model"); return OPTMC_EINVAL; }
  return launch_paths(ctx, mp, rng, M, N, dtype, S_dev, V_dev, ld);
  OPTMC_TRY_END
}

int optmc_philox_normals(optmc_ctx* ctx, const optmc_rng_params* rng, int32_t model, int64_t M, int32_t N,
                         int32_t which, int32_t dtype, void* Z_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_philox_normals(ctx, rng, model, M, N, which, dtype, Z_dev);
  OPTMC_TRY_END
}

int optmc_philox_kat(optmc_ctx* ctx, int32_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_philox_kat(ctx, n, ctr, key, out);
  OPTMC_TRY_END
}

int optmc_lsm_poly(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                   const optmc_lsm_params* lp, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  int rc = bind_sweep(ctx, S_dev, ld, M, N, dtype, lp);
  if (rc) return rc;
  rc = run_sweep(ctx);
  if (rc) return rc;
  return out ? fetch_results(ctx, out) : OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_lsm_fetch(optmc_ctx* ctx, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return fetch_results(ctx, out);
  OPTMC_TRY_END
}

int optmc_lsm_global(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                     const optmc_lsm_params* lp, optmc_global_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return lsm_global(ctx, S_dev, ld, M, N, dtype, lp, out);
  OPTMC_TRY_END
}

int optmc_lsm_mlp(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                  const optmc_lsm_params* lp, const optmc_mlp_params* np, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!lp) { set_error("null argument"); return OPTMC_EINVAL; }
  optmc_lsm_params l2 = *lp;
  l2.basis = OPTMC_BASIS_POLY2;  // unused by the network; keeps the shared validation
  l2.impl = OPTMC_SWEEP_SPLIT;
  int rc = bind_sweep(ctx, S_dev, ld, M, N, dtype, &l2);
  if (rc) return rc;
  rc = lsm_mlp(ctx, np, out);
  if (rc) return rc;
  return out ? fetch_results(ctx, out) : OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_mlp_init_params(int32_t hidden, uint64_t seed, int32_t date, float* out) {
  if ((hidden != 32 && hidden != 128) || !out) { set_error("hidden width must be 32 or 128"); return OPTMC_EINVAL; }
  return mlp_init_params_host(hidden, seed, date, out);
}

int optmc_mlp_grad_debug(optmc_ctx* ctx, int32_t hidden, int64_t n, const float* xs, const float* ys, const float* params,
                         float* grads, float* cont) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return mlp_grad_debug(ctx, hidden, n, xs, ys, params, grads, cont);
  OPTMC_TRY_END
}

int optmc_lsm_gram_len(int32_t basis) {
  if (basis == OPTMC_BASIS_POLY2) return 8;
  if (basis == OPTMC_BASIS_POLY3) return 11;
  return OPTMC_EINVAL;
}

int optmc_lsm_begin(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                    const optmc_lsm_params* lp) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  int rc = bind_sweep(ctx, S_dev, ld, M, N, dtype, lp);
  if (rc) return rc;
  ctx->sw.impl_used = OPTMC_SWEEP_SPLIT;
  return sweep_begin(ctx);
  OPTMC_TRY_END
}

int optmc_lsm_gram_date(optmc_ctx* ctx, int32_t t, double* gram_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!ctx->sw.S || t < 1 || t >= ctx->sw.N || !gram_dev) { set_error("bad date or no sweep bound"); return OPTMC_EINVAL; }
  return sweep_gram_date(ctx, t, gram_dev);
  OPTMC_TRY_END
}

int optmc_lsm_update_date(optmc_ctx* ctx, int32_t t, const double* gram_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!ctx->sw.S || t < 1 || t >= ctx->sw.N || !gram_dev) { set_error("bad date or no sweep bound"); return OPTMC_EINVAL; }
  return sweep_update_date(ctx, t, gram_dev);
  OPTMC_TRY_END
}

int optmc_lsm_finish(optmc_ctx* ctx, double* sums_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!ctx->sw.S || !sums_dev) { set_error("no sweep bound"); return OPTMC_EINVAL; }
  int rc = sweep_finish(ctx, sums_dev);
  if (rc) return rc;
  rc = sweep_finalize_price(ctx, sums_dev);  // local price; multi-GPU callers recompute from reduced sums
  if (rc) return rc;
  ctx->sw.have_results = true;
  return OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_price_american(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                         int32_t N, int32_t dtype, const optmc_lsm_params* lp, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!mp || !rng || !lp) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(mp->S0 > 0) || !(lp->K > 0) || !(mp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (mp->r < 0) { set_error("r must be non-negative."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  const size_t es = dtype == OPTMC_F64 ? 8 : 4;
  const int64_t ld = (M + 63) / 64 * 64;  // rows start on 256-byte boundaries: 128-bit stores and bulk copies
  int rc = ensure_bytes(&ctx->slab, &ctx->slab_bytes, (size_t)(N + 1) * ld * es);
  if (rc) return rc;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  rc = launch_paths(ctx, mp, rng, M, N, dtype, ctx->slab, nullptr, ld);
  if (rc) return rc;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  rc = bind_sweep(ctx, ctx->slab, ld, M, N, dtype, lp);
  if (rc) return rc;
  rc = run_sweep(ctx);
  if (rc) return rc;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->sw.n_launches += 1;  // the path kernel
  if (!out) return OPTMC_OK;
  rc = fetch_results(ctx, out);
  if (rc) return rc;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last_paths_ms = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) ctx->last_sweep_ms = ms;
  return OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_price_american_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                               int32_t dtype, int32_t basis, uint32_t semantics, int32_t n_options,
                               const optmc_american_option* opts, optmc_price_result* results) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return price_american_batch(ctx, mp, rng, M, dtype, basis, semantics, n_options, opts, results);
  OPTMC_TRY_END
}

int optmc_price_european_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                               int32_t N, int32_t dtype, int32_t n_options, const double* K, const double* T,
                               const int32_t* is_put, const int32_t* stream_id, optmc_european_result* results) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_european_batch(ctx, mp, rng, M, N, dtype, n_options, K, T, is_put, stream_id, results);
  OPTMC_TRY_END
}

int optmc_european_from_slab(optmc_ctx* ctx, const void* ST_dev, int64_t M, int32_t dtype, double K, double r,
                             double T, int32_t is_put, optmc_european_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_european_slab(ctx, ST_dev, M, dtype, K, r, T, is_put, out);
  OPTMC_TRY_END
}

int optmc_features_ref7(optmc_ctx* ctx, const void* S_dev, int64_t n, int32_t dtype, double K, double r, double T,
                        double t_current, void* F_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  (void)r;  // unused by the reference as well (om3:105)
  return launch_features(ctx, S_dev, n, dtype, K, T, t_current, F_dev);
  OPTMC_TRY_END
}

}  // extern "C"
