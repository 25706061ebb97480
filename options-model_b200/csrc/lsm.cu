// lsm.cu -- K2/K3: the Longstaff-Schwartz backward sweep (loop of om3:615-651 with the polynomial
// regressor of SURVEY.md 8(c)).  Two implementations with identical results:
//
//  * RESIDENT (default): ONE cooperative launch walks all exercise dates.  Each CTA owns a contiguous
//    slice of paths; the slice's cash-flows live in REGISTERS for the whole sweep (the exercised flag is
//    the sign bit), each date's price slab slice is staged into shared memory by the TMA engine
//    (cp.async.bulk + mbarrier, 2-3 stages ahead), and the per-date regression is one block reduction
//    plus one flag-based all-gather of 128-byte Gram slots through L2 -- no grid-wide barrier, no
//    atomics on the data path, fixed summation order (bit-reproducible).  HBM traffic: S is read once.
//  * SPLIT: three small launches per date (Gram / solve / update) with cash-flows in HBM.  Used when the
//    slice does not fit on chip or the slab is not 16-byte aligned, and as the per-date building blocks
//    of the path-sharded multi-GPU sweep (the host all-reduces the Gram vector between the calls).
#include <math.h>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

// =================================================================================================
// SPLIT implementation
// =================================================================================================
constexpr int kSplitThreads = 256;
constexpr int kSplitWarps = kSplitThreads / 32;

template <typename R> __global__ void lsm_init_kernel(const R* __restrict__ S_N, R* __restrict__ cf, long long M,
                                                      double K, int is_put) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x)
    cf[j] = (R)payoff<double>((double)S_N[j], K, is_put != 0);
}

// Sum the per-block partials in a fixed order (lane-strided, then butterfly): deterministic.
template <int Q>
__device__ __forceinline__ void reduce_partials_last_block(const double* partials, int nblocks, double* out,
                                                           int nwarps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = warp; q < Q; q += nwarps) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * Q + q];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += shfl_xor_f64(s, m);
    if (lane == 0) out[q] = s;
  }
}

template <typename R, int DEG>
__global__ void __launch_bounds__(kSplitThreads)
lsm_gram_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, R disc, double K, double invK,
                int is_put, int sticky, double* partials, unsigned int* ticket, double* gram_out) {
  constexpr int Q = Moments<DEG>::Q;
  __shared__ double red[kSplitWarps * Q];
  __shared__ bool is_last;
  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const R c = cf[j];
    const bool ex = sticky && signbit(c);
    const double s = (double)S_t[j];
    const double pay = payoff<double>(s, K, is_put != 0);
    if (pay > 0.0 && !ex) {
      const R y = fabs(c) * disc;  // the discounted cash-flow exactly as the update kernel will store it
      moments_accumulate<DEG>(acc, s * invK, (double)y);
    }
  }
  block_reduce_sum<Q, kSplitWarps>(acc, red);
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < Q; ++q) partials[(size_t)blockIdx.x * Q + q] = acc[q];
      __threadfence();
      const unsigned int prev = atomicAdd(ticket, 1u);
      is_last = (prev == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    reduce_partials_last_block<Q>(partials, gridDim.x, gram_out, kSplitWarps);
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

template <int DEG>
__global__ void lsm_solve_kernel(const double* __restrict__ gram, double* beta_t, long long* nitm_t, int* valid_t) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mom[Moments<DEG>::Q];
  for (int q = 0; q < Moments<DEG>::Q; ++q) mom[q] = gram[q];
  double beta[DEG + 1];
  const bool ok = solve_poly<DEG>(mom, beta);
  for (int i = 0; i <= DEG; ++i) beta_t[i] = ok ? beta[i] : nan("");
  for (int i = DEG + 1; i < kMaxBeta; ++i) beta_t[i] = nan("");
  *nitm_t = (long long)(mom[0] + 0.5);
  *valid_t = ok ? 1 : 0;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o > v ? o : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o < v ? o : v;
  }
  return v;
}

// Boundary bookkeeping: put -> max exercised S (bits of a positive double order like integers),
// call -> min exercised S.  "none" is 0 for max and ~0 for min.
__device__ __forceinline__ unsigned long long bnd_none(int is_put) { return is_put ? 0ull : ~0ull; }

template <typename R, int DEG>
__global__ void __launch_bounds__(kSplitThreads)
lsm_update_kernel(const R* __restrict__ S_t, R* __restrict__ cf, long long M, R disc, double K, double invK,
                  int is_put, int sticky, const double* __restrict__ beta_t, const int* __restrict__ valid_t,
                  unsigned long long* bnd_t, unsigned long long* exc_t) {
  double beta[DEG + 1];
#pragma unroll
  for (int i = 0; i <= DEG; ++i) beta[i] = beta_t[i];
  const bool valid = *valid_t != 0;
  unsigned long long bnd = bnd_none(is_put);
  unsigned int cnt = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const R c = cf[j];
    bool ex = sticky && signbit(c);
    R a = fabs(c) * disc;
    if (valid && !ex) {
      const double s = (double)S_t[j];
      const double pay = payoff<double>(s, K, is_put != 0);
      if (pay > 0.0 && pay > poly_eval<DEG>(beta, s * invK)) {  // strict '>' (om3:644)
        a = (R)pay;
        ex = true;
        cnt++;
        const unsigned long long b = (unsigned long long)__double_as_longlong(s);
        bnd = is_put ? (b > bnd ? b : bnd) : (b < bnd ? b : bnd);
      }
    }
    cf[j] = (sticky && ex) ? -a : a;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  bnd = is_put ? warp_max_u64(bnd) : warp_min_u64(bnd);
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(exc_t, (unsigned long long)cnt);
    if (is_put) atomicMax(bnd_t, bnd); else atomicMin(bnd_t, bnd);
  }
}

template <typename R>
__global__ void __launch_bounds__(kSplitThreads)
lsm_final_kernel(const R* __restrict__ cf, long long M, double* partials, unsigned int* ticket, double* sums_out) {
  __shared__ double red[kSplitWarps * 2];
  __shared__ bool is_last;
  double acc[2] = {0.0, 0.0};
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const double c = fabs((double)cf[j]);
    acc[0] += c;
    acc[1] += c * c;
  }
  block_reduce_sum<2, kSplitWarps>(acc, red);
  if (threadIdx.x == 0) {
    partials[(size_t)blockIdx.x * 2 + 0] = acc[0];
    partials[(size_t)blockIdx.x * 2 + 1] = acc[1];
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    reduce_partials_last_block<2>(partials, gridDim.x, sums_out, kSplitWarps);
    if (threadIdx.x == 0) { sums_out[2] = (double)M; *ticket = 0u; }
  }
}

__global__ void lsm_price_from_sums_kernel(const double* sums, double scale, double* final_out) {
  const double n = sums[2], s = sums[0], ss = sums[1];
  const double mean = s / n;
  double var = n > 1.0 ? (ss - n * mean * mean) / (n - 1.0) : 0.0;
  if (var < 0.0) var = 0.0;
  final_out[0] = mean * scale;
  final_out[1] = sqrt(var / n) * scale;
  final_out[2] = s;
  final_out[3] = ss;
}

__global__ void lsm_reset_stats_kernel(int n1, int is_put, double* betas, unsigned long long* bnd,
                                       unsigned long long* exc, long long* nitm, int* valid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1) return;
  for (int i = 0; i < kMaxBeta; ++i) betas[t * kMaxBeta + i] = nan("");
  bnd[t] = bnd_none(is_put);
  exc[t] = 0ull;
  nitm[t] = 0;
  valid[t] = 0;
}

static int split_grid(optmc_ctx* ctx, long long M) {
  long long g = (M + kSplitThreads * 4 - 1) / (kSplitThreads * 4);
  long long cap = (long long)ctx->sm_count * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int reset_stats(optmc_ctx* ctx) {
  const SweepDesc& sw = ctx->sw;
  const int n1 = sw.N + 1;
  lsm_reset_stats_kernel<<<(n1 + 127) / 128, 128, 0, ctx->stream>>>(n1, sw.lp.is_put, ctx->d_betas, ctx->d_bnd,
                                                                    ctx->d_exc, ctx->d_nitm, ctx->d_valid);
  ctx->launches++; ctx->sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

template <typename R> static int sweep_begin_t(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  int rc = ensure_bytes(&ctx->cf, &ctx->cf_bytes, (size_t)sw.M * sizeof(R));
  if (rc) return rc;
  const int grid = split_grid(ctx, sw.M);
  rc = ensure_bytes((void**)&ctx->partials, &ctx->partials_bytes, (size_t)grid * 16 * sizeof(double));
  if (rc) return rc;
  rc = reset_stats(ctx);
  if (rc) return rc;
  const R* S_N = static_cast<const R*>(sw.S) + (size_t)sw.N * sw.ld;
  lsm_init_kernel<R><<<grid, kSplitThreads, 0, ctx->stream>>>(S_N, static_cast<R*>(ctx->cf), sw.M, sw.lp.K,
                                                              sw.lp.is_put);
  ctx->launches++; sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

int sweep_begin(optmc_ctx* ctx) {
  return ctx->sw.dtype == OPTMC_F64 ? sweep_begin_t<double>(ctx) : sweep_begin_t<float>(ctx);
}

template <typename R, int DEG> static int gram_date_t(optmc_ctx* ctx, int t, double* gram_out) {
  SweepDesc& sw = ctx->sw;
  const int grid = split_grid(ctx, sw.M);
  const R* S_t = static_cast<const R*>(sw.S) + (size_t)t * sw.ld;
  const bool sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) != 0;
  lsm_gram_kernel<R, DEG><<<grid, kSplitThreads, 0, ctx->stream>>>(S_t, static_cast<const R*>(ctx->cf), sw.M,
                                                                   (R)sw.disc, sw.lp.K, 1.0 / sw.lp.K, sw.lp.is_put,
                                                                   sticky, ctx->partials, ctx->tickets, gram_out);
  ctx->launches++; sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

int sweep_gram_date(optmc_ctx* ctx, int t, double* gram_out) {
  const SweepDesc& sw = ctx->sw;
  if (sw.dtype == OPTMC_F64) return sw.deg == 2 ? gram_date_t<double, 2>(ctx, t, gram_out) : gram_date_t<double, 3>(ctx, t, gram_out);
  return sw.deg == 2 ? gram_date_t<float, 2>(ctx, t, gram_out) : gram_date_t<float, 3>(ctx, t, gram_out);
}

template <typename R, int DEG> static int update_date_t(optmc_ctx* ctx, int t, const double* gram) {
  SweepDesc& sw = ctx->sw;
  const int grid = split_grid(ctx, sw.M);
  const R* S_t = static_cast<const R*>(sw.S) + (size_t)t * sw.ld;
  const bool sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) != 0;
  lsm_solve_kernel<DEG><<<1, 32, 0, ctx->stream>>>(gram, ctx->d_betas + (size_t)t * kMaxBeta, ctx->d_nitm + t,
                                                   ctx->d_valid + t);
  lsm_update_kernel<R, DEG><<<grid, kSplitThreads, 0, ctx->stream>>>(
      S_t, static_cast<R*>(ctx->cf), sw.M, (R)sw.disc, sw.lp.K, 1.0 / sw.lp.K, sw.lp.is_put, sticky,
      ctx->d_betas + (size_t)t * kMaxBeta, ctx->d_valid + t, ctx->d_bnd + t, ctx->d_exc + t);
  ctx->launches += 2; sw.n_launches += 2;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

int sweep_update_date(optmc_ctx* ctx, int t, const double* gram) {
  const SweepDesc& sw = ctx->sw;
  if (sw.dtype == OPTMC_F64) return sw.deg == 2 ? update_date_t<double, 2>(ctx, t, gram) : update_date_t<double, 3>(ctx, t, gram);
  return sw.deg == 2 ? update_date_t<float, 2>(ctx, t, gram) : update_date_t<float, 3>(ctx, t, gram);
}

int sweep_finish(optmc_ctx* ctx, double* sums_out) {
  SweepDesc& sw = ctx->sw;
  const int grid = split_grid(ctx, sw.M);
  if (sw.dtype == OPTMC_F64)
    lsm_final_kernel<double><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const double*>(ctx->cf), sw.M,
                                                                      ctx->partials, ctx->tickets, sums_out);
  else
    lsm_final_kernel<float><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const float*>(ctx->cf), sw.M,
                                                                     ctx->partials, ctx->tickets, sums_out);
  ctx->launches++; sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

int sweep_finalize_price(optmc_ctx* ctx, const double* sums) {
  lsm_price_from_sums_kernel<<<1, 1, 0, ctx->stream>>>(sums, ctx->sw.final_scale, ctx->d_final);
  ctx->launches++; ctx->sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

// =================================================================================================
// RESIDENT implementation
// =================================================================================================
constexpr int kResWarps = kResThreads / 32;

struct ResArgs {
  const void* S;
  long long ld, M, chunk;
  int N, nstage;
  unsigned int stage_stride;  // bytes between stages in shared memory
  double K, invK, disc, final_scale;
  int is_put, sticky;
  double* xchg;               // [2][ncta][16]
  unsigned long long epoch_base;
  double* betas;              // [(N+1)][kMaxBeta]
  unsigned long long* bnd;    // [(N+1)]
  unsigned long long* exc;    // [(N+1)]
  long long* nitm;            // [(N+1)]
  double* final_out;          // [4]
};

// All-gather + fixed-order sum of QN doubles per CTA through 128-byte slots in L2.
// Called by warp 0 only; `tot` holds this CTA's block total in every lane on entry and the grid total in
// every lane on exit.  Double-buffered by `par`: a slot for epoch e can only be rewritten for e+2, which a
// CTA publishes only after it has gathered e+1 from everyone, i.e. after everyone has finished reading e.
template <int QN>
__device__ __forceinline__ void grid_allreduce_warp0(double (&tot)[QN], double* xchg, int par, int cta, int ncta,
                                                     unsigned long long epoch) {
  const int lane = threadIdx.x & 31;
  double* my = xchg + ((size_t)par * kMaxResidentCtas + cta) * kXchgSlotDoubles;
#pragma unroll
  for (int q = 0; q < QN; ++q)
    if (lane == q) st_relaxed_f64(my + q, tot[q]);
  __threadfence();
  __syncwarp();
  if (lane == 0) st_release_u64(reinterpret_cast<unsigned long long*>(my + (kXchgSlotDoubles - 1)), epoch);

  constexpr int MAXC = kMaxResidentCtas / 32;  // 5 slots per lane
  const double* base = xchg + (size_t)par * kMaxResidentCtas * kXchgSlotDoubles;
  bool ready[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) ready[i] = (lane + 32 * i) >= ncta;
  bool all;
  do {
    all = true;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      if (!ready[i]) {
        const unsigned long long f = ld_acquire_u64(reinterpret_cast<const unsigned long long*>(
            base + (size_t)(lane + 32 * i) * kXchgSlotDoubles + (kXchgSlotDoubles - 1)));
        ready[i] = (f == epoch);
        all = all && ready[i];
      }
    }
  } while (!all);
  double acc[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) acc[q] = 0.0;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    const int c = lane + 32 * i;
    if (c < ncta) {
      const double* slot = base + (size_t)c * kXchgSlotDoubles;
#pragma unroll
      for (int q = 0; q < QN; ++q) acc[q] += ld_relaxed_f64(slot + q);
    }
  }
  warp_allreduce_sum<QN>(acc);
#pragma unroll
  for (int q = 0; q < QN; ++q) tot[q] = acc[q];
}

template <typename R, int DEG, int PPT>
__global__ void __launch_bounds__(kResThreads, 1) lsm_resident_kernel(const ResArgs a) {
  constexpr int Q = Moments<DEG>::Q;
  static_assert(Q <= kXchgSlotDoubles - 1, "Gram vector must fit one exchange slot");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3];
  __shared__ double red[kResWarps * Q];
  __shared__ double s_beta[DEG + 1];
  __shared__ int s_valid;
  __shared__ unsigned long long s_bnd[2];
  __shared__ unsigned int s_cnt[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const long long base = (long long)cta * a.chunk;
  const long long rem = a.M - base;
  const int n_local = (int)(rem < a.chunk ? rem : a.chunk);
  const unsigned int bytes = (unsigned int)(((size_t)n_local * sizeof(R) + 15) / 16 * 16);
  const bool is_put = a.is_put != 0;
  const bool sticky = a.sticky != 0;
  const R disc = (R)a.disc;
  const int N = a.N, nstage = a.nstage;
  const R* Sbase = static_cast<const R*>(a.S) + base;

  auto stage_ptr = [&](int t) -> const R* {
    return reinterpret_cast<const R*>(smem_raw + (size_t)(t % nstage) * a.stage_stride);
  };
  auto issue_load = [&](int t) {  // thread 0 only
    uint64_t* bar = &mbar[t % nstage];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_load_1d(smem_raw + (size_t)(t % nstage) * a.stage_stride, Sbase + (size_t)t * a.ld, bytes, bar);
  };
  auto wait_stage = [&](int t) { mbar_wait(&mbar[t % nstage], (unsigned)(((N - t) / nstage) & 1)); };

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&mbar[s], 1);
    mbar_fence_init();
    s_bnd[0] = s_bnd[1] = bnd_none(a.is_put);
    s_cnt[0] = s_cnt[1] = 0u;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i)
      if (N - i >= 1) issue_load(N - i);
  }

  // ---- date N: cash-flows = payoff(S[N]) (om3:616) ----
  R cf[PPT];
  wait_stage(N);
  {
    const R* st = stage_ptr(N);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int j = tid + k * kResThreads;
      cf[k] = (j < n_local) ? (R)payoff<double>((double)st[j], a.K, is_put) : (R)0;
    }
  }

  for (int t = N - 1; t >= 1; --t) {
    wait_stage(t);
    const R* st = stage_ptr(t);
    // -- discount every path (om3:620), then the ITM-masked Gram moments (fp64) --
    double acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0.0;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int j = tid + k * kResThreads;
      if (j < n_local) {
        const R c = cf[k];
        const bool ex = sticky && signbit(c);
        const R y = fabs(c) * disc;
        cf[k] = ex ? -y : y;
        const double s = (double)st[j];
        if (!ex && payoff<double>(s, a.K, is_put) > 0.0) moments_accumulate<DEG>(acc, s * a.invK, (double)y);
      }
    }
    warp_allreduce_sum<Q>(acc);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < Q; ++q) red[warp * Q + q] = acc[q];
    }
    __syncthreads();  // #1: block partials visible; every thread is done with stage (t+1)
    if (tid == 0) {
      if (t + 1 <= N - 1) {  // flush the exercise statistics of date t+1
        const int p = (t + 1) & 1;
        if (s_cnt[p]) {
          atomicAdd(a.exc + (t + 1), (unsigned long long)s_cnt[p]);
          if (is_put) atomicMax(a.bnd + (t + 1), s_bnd[p]); else atomicMin(a.bnd + (t + 1), s_bnd[p]);
        }
        s_cnt[p] = 0u;
        s_bnd[p] = bnd_none(a.is_put);
      }
      if (t + 1 - nstage >= 1) issue_load(t + 1 - nstage);  // refill the stage date t+1 just vacated
    }
    if (warp == 0) {
      double tot[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) tot[q] = (lane < kResWarps) ? red[lane * Q + q] : 0.0;
      warp_allreduce_sum<Q>(tot);
      grid_allreduce_warp0<Q>(tot, a.xchg, t & 1, cta, ncta, a.epoch_base + (unsigned long long)(N - t));
      if (lane == 0) {
        double beta[DEG + 1];
        const bool ok = solve_poly<DEG>(tot, beta);
        s_valid = ok ? 1 : 0;
#pragma unroll
        for (int i = 0; i <= DEG; ++i) s_beta[i] = ok ? beta[i] : 0.0;
        if (cta == 0) {
#pragma unroll
          for (int i = 0; i <= DEG; ++i) a.betas[(size_t)t * kMaxBeta + i] = ok ? beta[i] : nan("");
          a.nitm[t] = (long long)(tot[0] + 0.5);
        }
      }
    }
    __syncthreads();  // #2: beta visible
    if (s_valid) {
      double beta[DEG + 1];
#pragma unroll
      for (int i = 0; i <= DEG; ++i) beta[i] = s_beta[i];
      unsigned int cnt = 0;
      unsigned long long bnd = bnd_none(a.is_put);
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int j = tid + k * kResThreads;
        if (j < n_local) {
          const R c = cf[k];
          if (!(sticky && signbit(c))) {
            const double s = (double)st[j];
            const double pay = payoff<double>(s, a.K, is_put);
            if (pay > 0.0 && pay > poly_eval<DEG>(beta, s * a.invK)) {  // strict '>' (om3:644)
              cf[k] = sticky ? -(R)pay : (R)pay;                         // sticky flag = sign bit (om3:649)
              cnt++;
              const unsigned long long b = (unsigned long long)__double_as_longlong(s);
              bnd = is_put ? (b > bnd ? b : bnd) : (b < bnd ? b : bnd);
            }
          }
        }
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt) {  // warp-uniform
        bnd = is_put ? warp_max_u64(bnd) : warp_min_u64(bnd);
        if (lane == 0) {
          atomicAdd(&s_cnt[t & 1], cnt);
          if (is_put) atomicMax(&s_bnd[t & 1], bnd); else atomicMin(&s_bnd[t & 1], bnd);
        }
      }
    }
  }

  // ---- final reduction: mean and standard error of the cash-flows (om3:651) ----
  double fin[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int j = tid + k * kResThreads;
    if (j < n_local) {
      const double c = fabs((double)cf[k]);
      fin[0] += c;
      fin[1] += c * c;
    }
  }
  warp_allreduce_sum<2>(fin);
  if (lane == 0) { red[warp * 2] = fin[0]; red[warp * 2 + 1] = fin[1]; }
  __syncthreads();
  if (tid == 0 && N - 1 >= 1) {  // statistics of date 1 (or of the last processed date)
    const int p = 1 & 1;
    if (s_cnt[p]) {
      atomicAdd(a.exc + 1, (unsigned long long)s_cnt[p]);
      if (is_put) atomicMax(a.bnd + 1, s_bnd[p]); else atomicMin(a.bnd + 1, s_bnd[p]);
    }
  }
  if (warp == 0) {
    double tot[2];
    tot[0] = (lane < kResWarps) ? red[lane * 2] : 0.0;
    tot[1] = (lane < kResWarps) ? red[lane * 2 + 1] : 0.0;
    warp_allreduce_sum<2>(tot);
    grid_allreduce_warp0<2>(tot, a.xchg, 0, cta, ncta, a.epoch_base + (unsigned long long)N);
    if (cta == 0 && lane == 0) {
      const double n = (double)a.M;
      const double mean = tot[0] / n;
      double var = n > 1.0 ? (tot[1] - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      a.final_out[0] = mean * a.final_scale;
      a.final_out[1] = sqrt(var / n) * a.final_scale;
      a.final_out[2] = tot[0];
      a.final_out[3] = tot[1];
    }
  }
}

struct ResPlan {
  int ncta = 0, ppt = 0, nstage = 0;
  long long chunk = 0;
  unsigned int stage_stride = 0;
  size_t smem = 0;
};

static const int kPptChoices[] = {1, 2, 4, 8, 16, 32, 56};

static bool plan_resident(optmc_ctx* ctx, const SweepDesc& sw, ResPlan* p, std::string* why) {
  const size_t es = sw.dtype == OPTMC_F64 ? 8 : 4;
  if ((uintptr_t)sw.S % 16 != 0 || (sw.ld * es) % 16 != 0) { *why = "slab not 16-byte aligned"; return false; }
  if (ctx->cc < 90) { *why = "bulk async copy needs sm_90+"; return false; }
  int ncta_cap = ctx->sm_count < kMaxResidentCtas ? ctx->sm_count : kMaxResidentCtas;
  long long ncta = (sw.M + kResThreads - 1) / kResThreads;  // at least one path per thread before adding CTAs
  if (ncta > ncta_cap) ncta = ncta_cap;
  if (ncta < 1) ncta = 1;
  long long chunk = (sw.M + ncta - 1) / ncta;
  chunk = (chunk + 3) / 4 * 4;
  ncta = (sw.M + chunk - 1) / chunk;
  const int max_ppt = sw.dtype == OPTMC_F64 ? 32 : 56;
  int ppt = 0;
  for (int c : kPptChoices)
    if ((long long)c * kResThreads >= chunk) { ppt = c; break; }
  if (ppt == 0 || ppt > max_ppt) { *why = "slice exceeds the register-resident capacity"; return false; }
  const size_t stride = (chunk * es + 127) / 128 * 128;
  const size_t avail = (size_t)ctx->max_smem_optin - 4096;  // static shared + slack
  int nstage = 3;
  if (stride * 3 > avail) nstage = 2;
  if (stride * 2 > avail) { *why = "slice exceeds shared memory"; return false; }
  p->ncta = (int)ncta; p->ppt = ppt; p->nstage = nstage; p->chunk = chunk;
  p->stage_stride = (unsigned int)stride; p->smem = stride * nstage;
  return true;
}

bool resident_eligible(optmc_ctx* ctx, const SweepDesc& sw, std::string* why) {
  ResPlan p;
  return plan_resident(ctx, sw, &p, why);
}

template <typename R, int DEG, int PPT> static int launch_resident_t(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  auto kern = lsm_resident_kernel<R, DEG, PPT>;
  OPTMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  void* args[] = {(void*)&a};
  OPTMC_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(p.ncta), dim3(kResThreads), args, p.smem, ctx->stream));
  ctx->launches++; ctx->sw.n_launches++;
  return OPTMC_OK;
}

template <typename R, int DEG> static int launch_resident_ppt(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  switch (p.ppt) {
    case 1: return launch_resident_t<R, DEG, 1>(ctx, p, a);
    case 2: return launch_resident_t<R, DEG, 2>(ctx, p, a);
    case 4: return launch_resident_t<R, DEG, 4>(ctx, p, a);
    case 8: return launch_resident_t<R, DEG, 8>(ctx, p, a);
    case 16: return launch_resident_t<R, DEG, 16>(ctx, p, a);
    case 32: return launch_resident_t<R, DEG, 32>(ctx, p, a);
    case 56:
      if (sizeof(R) == 4) return launch_resident_t<float, DEG, 56>(ctx, p, a);
  }
  set_error("no resident instantiation for this slice size");
  return OPTMC_EUNSUPPORTED;
}

int sweep_resident(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  ResPlan p;
  std::string why;
  if (!plan_resident(ctx, sw, &p, &why)) { set_error("resident sweep unavailable: " + why); return OPTMC_EUNSUPPORTED; }
  int rc = reset_stats(ctx);
  if (rc) return rc;
  ResArgs a{};
  a.S = sw.S; a.ld = sw.ld; a.M = sw.M; a.chunk = p.chunk; a.N = sw.N; a.nstage = p.nstage;
  a.stage_stride = p.stage_stride;
  a.K = sw.lp.K; a.invK = 1.0 / sw.lp.K; a.disc = sw.disc; a.final_scale = sw.final_scale;
  a.is_put = sw.lp.is_put; a.sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) ? 1 : 0;
  a.xchg = ctx->xchg; a.epoch_base = ctx->epoch;
  a.betas = ctx->d_betas; a.bnd = ctx->d_bnd; a.exc = ctx->d_exc; a.nitm = ctx->d_nitm; a.final_out = ctx->d_final;
  ctx->epoch += (unsigned long long)sw.N + 2ull;
  if (sw.dtype == OPTMC_F64)
    return sw.deg == 2 ? launch_resident_ppt<double, 2>(ctx, p, a) : launch_resident_ppt<double, 3>(ctx, p, a);
  return sw.deg == 2 ? launch_resident_ppt<float, 2>(ctx, p, a) : launch_resident_ppt<float, 3>(ctx, p, a);
}

}  // namespace optmc
