// api.cu -- the extern "C" boundary declared in include/optmc.h: context, workspaces, argument
// validation (the reference's ValueError conditions, om3:447-452), dispatch and result read-back.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include <stdlib.h>

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

static size_t per_date_row_bytes() { return kMaxBeta * 8 + 8 + 8 + 8 + 4; }

int ensure_per_date(optmc_ctx* ctx, int N) {
  const size_t need = (size_t)N + 1;
  if (need <= ctx->per_date_cap) return OPTMC_OK;
  // ONE block [betas | boundary | exercise counts | ITM counts | valid]: fetch_results brings the per-date outputs back
  // with a single copy (five pageable copies + two stream synchronisations were ~10 % of a 1 ms pricing call)
  cudaFree(ctx->d_betas);
  ctx->d_betas = nullptr; ctx->d_bnd = nullptr; ctx->d_exc = nullptr; ctx->d_nitm = nullptr; ctx->d_valid = nullptr;
  ctx->per_date_cap = 0;
  const size_t cap = need < 512 ? 512 : need;
  char* block = nullptr;
  OPTMC_CUDA(cudaMalloc((void**)&block, cap * per_date_row_bytes()));
  ctx->d_betas = reinterpret_cast<double*>(block);
  ctx->d_bnd = reinterpret_cast<unsigned long long*>(block + cap * kMaxBeta * 8);
  ctx->d_exc = ctx->d_bnd + cap;
  ctx->d_nitm = reinterpret_cast<long long*>(ctx->d_exc + cap);
  ctx->d_valid = reinterpret_cast<int*>(ctx->d_nitm + cap);
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
  // REF7 per date: within one date sqrt(tau) is a constant, x sqrt(tau) is a multiple of x and the hinge max(x-1, 0) is
  // 0 (puts) or x - 1 (calls) on the in-the-money rows, so the seven reference features span exactly [1, x, x^2, x^3]:
  // the fitted continuation values are those of POLY3 (betas are reported in that 4-term form).
  if (lp->basis != OPTMC_BASIS_POLY2 && lp->basis != OPTMC_BASIS_POLY3 && lp->basis != OPTMC_BASIS_REF7) { set_error("basis must be POLY2, POLY3 or REF7"); return OPTMC_EINVAL; }
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
  sw.deg = lp->basis == OPTMC_BASIS_POLY2 ? 2 : 3;
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
  { const StrikeConsts kc = strike_consts(lp->K, lp->is_put != 0, dtype == OPTMC_F32); sw.Kh = kc.Kh; sw.Kl = kc.Kl; }
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
    int rc = sweep_split_fused(ctx);
    if (rc) return rc;
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
  // one pinned staging buffer: [final (4 doubles) | flags (4 ints) | per-date block up to the requested arrays]
  const bool want_arrays = out->betas || out->boundary || out->ex_count || out->n_itm;
  const size_t cap = ctx->per_date_cap;
  const size_t arr_bytes = want_arrays ? cap * (kMaxBeta * 8 + 3 * 8) : 0;  // betas, boundary, exercise counts, ITM counts
  const size_t need = 64 + arr_bytes;
  if (need > ctx->h_fetch_cap) {
    cudaFreeHost(ctx->h_fetch);
    ctx->h_fetch = nullptr; ctx->h_fetch_cap = 0;
    OPTMC_CUDA(cudaMallocHost(&ctx->h_fetch, need));
    ctx->h_fetch_cap = need;
  }
  char* hbuf = static_cast<char*>(ctx->h_fetch);
  auto fetch = [&]() -> int {
    OPTMC_CUDA(cudaMemcpyAsync(hbuf, ctx->d_final, 4 * sizeof(double) + 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (want_arrays) OPTMC_CUDA(cudaMemcpyAsync(hbuf + 64, ctx->d_betas, arr_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
    return OPTMC_OK;
  };
  int rcf = fetch();
  if (rcf) return rcf;
  if (sw.impl_used == OPTMC_SWEEP_RESIDENT) {
    // The persistent sweep sums Gram moments in fixed point (|moment| < 2^43).  On overflow (or NaN/Inf
    // prices) AUTO repeats the sweep with the split kernels; an explicit RESIDENT request fails loudly.
    const int* flags = reinterpret_cast<const int*>(hbuf + 4 * sizeof(double));
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
      rcf = fetch();
      if (rcf) return rcf;
    }
  }
  double fin[4];
  memcpy(fin, hbuf, sizeof(fin));
  const double* hb = reinterpret_cast<const double*>(hbuf + 64);
  const unsigned long long* hbnd = reinterpret_cast<const unsigned long long*>(hbuf + 64 + cap * kMaxBeta * 8);
  const unsigned long long* hexc = hbnd + cap;
  const long long* hn = reinterpret_cast<const long long*>(hexc + cap);
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
  OPTMC_CUDA(cudaMalloc((void**)&c->d_final, 4 * sizeof(double) + 4 * sizeof(int)));  // [price, stderr, sum, sum^2 | flags]
  c->d_flags = reinterpret_cast<int*>(c->d_final + 4);
  OPTMC_CUDA(cudaMalloc(&c->xchg, xchg_bytes()));
  OPTMC_CUDA(cudaMemset(c->xchg, 0, xchg_bytes()));
  for (int i = 0; i < 3; ++i) OPTMC_CUDA(cudaEventCreate(&c->ev[i]));
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
  cudaFree(ctx->d_betas);  // one block with d_bnd / d_exc / d_nitm / d_valid
  cudaFreeHost(ctx->h_fetch);
  cudaFree(ctx->d_final);  // one block with d_flags
  cudaFree(ctx->xchg); cudaFree(ctx->batch_dev); cudaFree(ctx->spill); cudaFree(ctx->qmc_dev); cudaFree(ctx->gnet_rows); cudaFree(ctx->eu_out); cudaFree(ctx->eu_par); cudaFree(ctx->eu_tickets);
  for (int r = 0; r < 8; ++r) if (ctx->comm.opened[r]) cudaIpcCloseMemHandle(ctx->comm.peers[r]);
  cudaFree(ctx->comm.local);
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

int optmc_workspace_bytes(int64_t M, int32_t N, int32_t dtype, int32_t n_options, int64_t* bytes) {
  if (!bytes) { set_error("null argument"); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0 || n_options <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  const int64_t es = dtype == OPTMC_F64 ? 8 : 4;
  const int64_t ld = (M + 63) / 64 * 64;
  const int64_t slab = (int64_t)(N + 1) * ld * es;
  // a batch is priced in waves of at most one option per SM (lsm_resident.cu): <= 160 slabs alive at once
  const int64_t wave = n_options < 160 ? n_options : 160;
  const int64_t per_date = (int64_t)((N + 1) < 512 ? 512 : (N + 1)) * (kMaxBeta * 8 + 8 + 8 + 8 + 4);
  *bytes = wave * slab + ld * es /* split-sweep cash-flows */ + wave * (per_date + 4096) + (1 << 20);
  return OPTMC_OK;
}

int optmc_ctx_workspace_bytes(optmc_ctx* ctx, int64_t* bytes) {
  if (!ctx || !bytes) { set_error("null argument"); return OPTMC_EINVAL; }
  *bytes = (int64_t)(ctx->slab_bytes + ctx->cf_bytes + ctx->partials_bytes + ctx->batch_dev_cap + ctx->spill_bytes + ctx->gnet_rows_cap + ctx->eu_out_cap +
                     ctx->eu_par_cap + ctx->eu_tickets_cap +
                     ctx->per_date_cap * (kMaxBeta * sizeof(double) + 2 * sizeof(unsigned long long) + sizeof(long long) + sizeof(int)) +
                     xchg_bytes() + 1024 * sizeof(unsigned int) + 20 * sizeof(double) + 4 * sizeof(int));
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

int optmc_paths_localvol(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, const optmc_ivnet* net,
                         int64_t M, int32_t N, int32_t dtype, void* S_dev, int64_t ld) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_paths_localvol(ctx, mp, rng, net, M, N, dtype, S_dev, ld);
  OPTMC_TRY_END
}

int optmc_ivnet_sigma(optmc_ctx* ctx, const optmc_ivnet* net, double tau, const double* S_dev, int64_t n, double* sigma_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return ivnet_sigma_batch(ctx, net, tau, S_dev, n, sigma_dev);
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

// ---- path-sharded sweep over several GPUs: exchange slots in peer-mapped memory ----------------
static size_t comm_slot_bytes() {  // one slot block [2 parities][ranks][words] per option of a grouped launch
  return (optmc::kCommSweepWords + optmc::kCommGnetWords) * sizeof(unsigned long long);  // + the network LSM's region
}

int optmc_comm_export(optmc_ctx* ctx, void* handle_out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!handle_out) { set_error("null argument"); return OPTMC_EINVAL; }
  static_assert(sizeof(cudaIpcMemHandle_t) == OPTMC_COMM_HANDLE_BYTES, "handle size");
  if (!ctx->comm.local) {
    OPTMC_CUDA(cudaMalloc((void**)&ctx->comm.local, comm_slot_bytes()));
  }
  OPTMC_CUDA(cudaMemset(ctx->comm.local, 0, comm_slot_bytes()));
  OPTMC_CUDA(cudaDeviceSynchronize());
  ctx->comm.g = 2;
  ctx->comm.gn_step = 1; ctx->comm.gn_meta = 1;
  cudaIpcMemHandle_t h;
  OPTMC_CUDA(cudaIpcGetMemHandle(&h, ctx->comm.local));
  memcpy(handle_out, &h, sizeof(h));
  return OPTMC_OK;
  OPTMC_TRY_END
}

static void comm_close(optmc_ctx* ctx) {
  for (int r = 0; r < 8; ++r) {
    if (ctx->comm.opened[r]) cudaIpcCloseMemHandle(ctx->comm.peers[r]);
    ctx->comm.opened[r] = false;
    ctx->comm.peers[r] = nullptr;
  }
  ctx->comm.nranks = 0;
}

int optmc_comm_init(optmc_ctx* ctx, int32_t rank, int32_t nranks, const void* handles) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!handles) { set_error("null argument"); return OPTMC_EINVAL; }
  if (nranks < 1 || nranks > optmc::kCommMaxRanks || rank < 0 || rank >= nranks) { set_error("bad rank / nranks (at most 8 ranks)"); return OPTMC_EINVAL; }
  if (!ctx->comm.local) { set_error("optmc_comm_export must be called first"); return OPTMC_EINVAL; }
  comm_close(ctx);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { ctx->comm.peers[r] = ctx->comm.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { comm_close(ctx); return cuda_fail(e, "cudaIpcOpenMemHandle (peer exchange slots)"); }
    ctx->comm.peers[r] = static_cast<unsigned long long*>(p);
    ctx->comm.opened[r] = true;
  }
  ctx->comm.rank = rank;
  ctx->comm.nranks = nranks;
  return OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_comm_finalize(optmc_ctx* ctx) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  comm_close(ctx);
  return OPTMC_OK;
  OPTMC_TRY_END
}

int optmc_lsm_poly_sharded(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M_local, int64_t M_total, int32_t N,
                           int32_t dtype, const optmc_lsm_params* lp, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (ctx->comm.nranks < 1) { set_error("optmc_comm_init must be called first"); return OPTMC_EINVAL; }
  if (M_total < M_local) { set_error("M_total must be >= M_local"); return OPTMC_EINVAL; }
  int rc = bind_sweep(ctx, S_dev, ld, M_local, N, dtype, lp);
  if (rc) return rc;
  std::string why;
  if (!resident_eligible(ctx, ctx->sw, &why)) {
    set_error("sharded sweep needs the persistent kernel on every rank: " + why);
    return OPTMC_EUNSUPPORTED;
  }
  ctx->sharded_M_total = M_total;
  rc = sweep_resident(ctx);
  ctx->sharded_M_total = 0;
  if (rc) return rc;
  ctx->sw.impl_used = OPTMC_SWEEP_RESIDENT;
  ctx->sw.have_results = true;
  int flags[4] = {0, 0, 0, 0};
  OPTMC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flags[1]) { set_error("sharded sweep: a peer rank did not answer (exchange timed out); re-run optmc_comm_export / optmc_comm_init on every rank"); return OPTMC_ECUDA; }
  if (flags[0]) { set_error("sharded sweep: a Gram moment left the fixed-point exchange range (|m| < 2^43) or is not finite"); return OPTMC_EUNSUPPORTED; }
  if (!out) return OPTMC_OK;
  rc = fetch_results(ctx, out);
  out->n_paths = M_total;
  return rc;
  OPTMC_TRY_END
}

int optmc_lsm_zero_cashflows(optmc_ctx* ctx, int64_t* count) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return sweep_zero_count(ctx, count);
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

int optmc_lsm_apply_policy(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                           const optmc_lsm_params* lp, const double* betas, optmc_lsm_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return lsm_apply_policy(ctx, S_dev, ld, M, N, dtype, lp, betas, out);
  OPTMC_TRY_END
}

int optmc_lsm_gnet(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                   const optmc_lsm_params* lp, const optmc_gnet_params* gp, optmc_gnet_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (gp && gp->per_date) {  // a fresh network per exercise date (the loop of om2:277-310 with om3's regressor)
    if (!lp || !out) { set_error("null argument"); return OPTMC_EINVAL; }
    optmc_lsm_params l2 = *lp;
    l2.basis = OPTMC_BASIS_POLY2;  // unused by the network; keeps the shared validation
    l2.impl = OPTMC_SWEEP_SPLIT;
    int rc = bind_sweep(ctx, S_dev, ld, M, N, dtype, &l2);
    if (rc) return rc;
    rc = gnet_validate(ctx, gp);
    if (rc) return rc;
    return lsm_gnet_per_date(ctx, gp, out);
  }
  return lsm_gnet(ctx, S_dev, ld, M, 0, N, dtype, lp, gp, out);
  OPTMC_TRY_END
}

int optmc_lsm_gnet_sharded(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M_local, int64_t M_total, int32_t N,
                           int32_t dtype, const optmc_lsm_params* lp, const optmc_gnet_params* gp, optmc_gnet_result* out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (ctx->comm.nranks < 1) { set_error("optmc_comm_init must be called first"); return OPTMC_EINVAL; }
  if (M_total < M_local || M_total <= 0) { set_error("M_total must be >= M_local"); return OPTMC_EINVAL; }
  if (gp && gp->per_date) { set_error("sharded network LSM: the per-date fit is single-GPU"); return OPTMC_EUNSUPPORTED; }
  return lsm_gnet(ctx, S_dev, ld, M_local, M_total, N, dtype, lp, gp, out);
  OPTMC_TRY_END
}

int optmc_gnet_shard_plan(const int64_t* n_rows, int32_t nranks, int32_t batch, int64_t b, int32_t rank, int64_t* lo,
                          int64_t* hi, int64_t* global_rows) {
  OPTMC_TRY_BEGIN
  return gnet_shard_plan(n_rows, nranks, batch, b, rank, lo, hi, global_rows);
  OPTMC_TRY_END
}

int optmc_gnet_grad_debug(optmc_ctx* ctx, int64_t n, const float* feat, const float* ys, const float* params, float* grads,
                          float* loss) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return gnet_grad_debug(ctx, n, feat, ys, params, grads, loss);
  OPTMC_TRY_END
}

int optmc_gnet_streams_debug(optmc_ctx* ctx, uint64_t seed, int32_t epoch, int32_t step, double dropout, int64_t n_rows,
                             int64_t* perm_out, const uint32_t* row_ids, int64_t n_ids, uint32_t* keep_out) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return gnet_streams_debug(ctx, seed, epoch, step, dropout, n_rows, reinterpret_cast<long long*>(perm_out), row_ids, n_ids, keep_out);
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
  // the caller wants the price only (the reference's price_american_enhanced_lsm returns a float): no per-date outputs
  ctx->sw.no_arrays = out && !out->betas && !out->boundary && !out->ex_count && !out->n_itm;
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
  return price_american_batch(ctx, mp, rng, M, dtype, basis, semantics, n_options, opts, results, nullptr);
  OPTMC_TRY_END
}

int optmc_price_american_batch_ex(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                                  int32_t dtype, int32_t basis, uint32_t semantics, int32_t n_options,
                                  const optmc_american_option* opts, optmc_price_result* results,
                                  optmc_batch_extras* extras) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return price_american_batch(ctx, mp, rng, M, dtype, basis, semantics, n_options, opts, results, extras);
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

int optmc_qmc_normals(optmc_ctx* ctx, int64_t M, int32_t N, int32_t factors, int32_t brownian_bridge, int64_t pair_offset,
                      const uint32_t* digital_shift, int32_t dtype, void* Z1_dev, void* Z2_dev) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  return launch_qmc_normals(ctx, M, N, factors, brownian_bridge, pair_offset, digital_shift, dtype, Z1_dev, Z2_dev);
  OPTMC_TRY_END
}

int optmc_qmc_bridge_schedule(int32_t N, int32_t* idx, int32_t* left, int32_t* right, double* wl, double* wr, double* sd) {
  OPTMC_TRY_BEGIN
  if (!idx || !left || !right || !wl || !wr || !sd) { set_error("null argument"); return OPTMC_EINVAL; }
  return bridge_schedule_host(N, idx, left, right, wl, wr, sd);
  OPTMC_TRY_END
}

int optmc_price_european_grid(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                              int32_t dtype, int32_t n_options, const double* S0, const double* K, const double* T,
                              const int32_t* N, const int32_t* is_put, const int32_t* stream_id,
                              optmc_european_result* results) {
  OPTMC_TRY_BEGIN
  OPTMC_ENTER(ctx);
  if (!N) { set_error("null argument"); return OPTMC_EINVAL; }
  return launch_european_batch(ctx, mp, rng, M, N[0], dtype, n_options, K, T, is_put, stream_id, results, N, S0);
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
