// lsm.cu -- K2/K3: the Longstaff-Schwartz backward sweep (loop of om3:615-651 with the polynomial
// regressor of SURVEY.md 8(c)).  Two implementations with identical results:
//
//  * RESIDENT (default, lsm_resident*.cu): ONE cooperative launch walks all exercise dates.  Each CTA owns a
//    contiguous slice of paths; the slice's cash-flows live in REGISTERS for the whole sweep (the exercised flag is
//    the sign bit), each date's price slab slice is staged into shared memory by the TMA engine (cp.async.bulk +
//    mbarrier, 2-3 stages ahead), and the per-date regression is one block reduction plus an order-independent
//    fixed-point sum of the CTA totals through L2 (integer `red`, sum and arrival count in one word) -- no grid-wide
//    barrier, bit-reproducible.  HBM traffic: S is read once.
//  * SPLIT (this file): cash-flows in HBM.  The internal sweep is one fused streaming launch per date
//    (lsm_stream_kernel: decision of t+1, moments of t, last-CTA solve); the three per-date kernels (Gram / solve /
//    update) are the building blocks of the host-looped path-sharded multi-GPU sweep, where the host all-reduces the
//    Gram vector between the calls.  Used when a slice does not fit on chip or the slab is not 16-byte aligned.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

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
                                                      double K, double Kh, double Kl, int is_put) {
  const R sgn = is_put ? (R)-1 : (R)1;
  const R c1 = (R)(is_put ? Kh : -Kh), c2 = (R)(is_put ? Kl : -Kl);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const R sr = S_N[j];
    cf[j] = payoff<double>((double)sr, K, is_put != 0) > 0.0 ? fma(sgn, sr, c1) + c2 : (R)0;
  }
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

// Cash-flows are stored in "date-N money" (c~ = c_t / D_t, D_t = disc^(N - t); see lsm_resident_kernel.cuh):
// the value of path j at date t is cf[j] * D_t (dg), so the per-date discount (om3:620) never touches HBM.
template <typename R, int DEG>
__global__ void __launch_bounds__(kSplitThreads)
lsm_gram_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, R dg, double K, double invK,
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
      const R y = fabs(c) * dg;  // the cash-flow at date t exactly as the persistent sweep forms it
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

// ---- fused streaming pass of the internal split sweep (options whose cash-flows do not fit on chip) --------------
// One launch per exercise date does what update(t+1) + Gram(t) + solve(t) did in three: every path is touched once
// (128-bit loads of S[t+1], S[t] and the cash-flows: 12 bytes per path instead of 16), the exercise decision of date
// t+1 is applied first -- with the coefficients the previous launch solved -- then the path enters the moments of date
// t exactly as the separate kernels would see it; the last CTA to finish reduces the partials in a fixed order and
// solves the normal equations.  Same arithmetic per path as lsm_gram_kernel / lsm_update_kernel (which remain the
// per-date building blocks of the host-looped multi-GPU sweep).
template <typename R, int VEC> struct VecMem;
template <> struct VecMem<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) { const float4 x = *reinterpret_cast<const float4*>(p); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct VecMem<double, 4> {
  static __device__ __forceinline__ void load(const double* p, double (&v)[4]) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]); reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
  }
};
template <typename R> struct VecMem<R, 1> {
  static __device__ __forceinline__ void load(const R* p, R (&v)[1]) { v[0] = p[0]; }
  static __device__ __forceinline__ void store(R* p, const R (&v)[1]) { p[0] = v[0]; }
};

struct StreamArgs {
  const void* S_dec;   // price row of the decision date t+1, or NULL (first launch: nothing to decide yet)
  const void* S_gram;  // price row of the regression date t, or NULL (last launch: only the decision of date 1)
  void* cf;
  long long M;
  double dinv, dg;     // 1 / D_(t+1), D_t
  double K, Kh, Kl, invK;
  double kk;           // sgn * Kcmp: (sgn * s > kk) in the storage type <=> payoff(s) > 0 in fp64 (see fill_group)
  int is_put, sticky;
  const double* beta_dec; const int* valid_dec;          // coefficients of date t+1 (solved by the previous launch)
  unsigned long long* bnd_dec; unsigned long long* exc_dec;
  double* partials; unsigned int* ticket;
  double* beta_out; long long* nitm_out; int* valid_out; // solve of date t
};

template <typename R, int DEG, int VEC>
__global__ void __launch_bounds__(kSplitThreads) lsm_stream_kernel(const StreamArgs a) {
  constexpr int Q = Moments<DEG>::Q;
  __shared__ double red[kSplitWarps * Q];
  __shared__ bool is_last;
  const bool have_dec = a.S_dec != nullptr && *a.valid_dec != 0;
  const bool have_gram = a.S_gram != nullptr;
  const bool is_put = a.is_put != 0, sticky = a.sticky != 0;
  double dec[DEG + 1];
#pragma unroll
  for (int i = 0; i <= DEG; ++i) dec[i] = 0.0;
  if (have_dec) {  // decision polynomial in the raw price, exactly as lsm_update_kernel forms it
    double sc = 1.0;
#pragma unroll
    for (int i = 0; i <= DEG; ++i) {
      double d = -a.beta_dec[i] * sc;
      if (i == 0) d += is_put ? a.K : -a.K;
      if (i == 1) d += is_put ? -1.0 : 1.0;
      dec[i] = d;
      sc *= a.invK;
    }
  }
  const R sgn = is_put ? (R)-1 : (R)1;
  const R c1 = (R)(is_put ? a.Kh : -a.Kh), c2 = (R)(is_put ? a.Kl : -a.Kl);
  const R dinv = (R)a.dinv, dg = (R)a.dg, kk = (R)a.kk;
  const R* Sd = static_cast<const R*>(a.S_dec);
  const R* Sg = static_cast<const R*>(a.S_gram);
  R* cfp = static_cast<R*>(a.cf);
  double acc[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) acc[q] = 0.0;
  unsigned long long bnd = bnd_none(a.is_put);
  unsigned int cnt = 0;
  const long long units = a.M / VEC;  // M % VEC == 0 (launcher)
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
    const long long j = u * VEC;
    R c[VEC], sd[VEC], sg[VEC];
    VecMem<R, VEC>::load(cfp + j, c);
    if (have_dec) VecMem<R, VEC>::load(Sd + j, sd);
    if (have_gram) VecMem<R, VEC>::load(Sg + j, sg);
    bool changed = false;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      if (have_dec && !(sticky && signbit(c[i])) && sgn * sd[i] > kk) {  // live and in the money
        const double s = (double)sd[i];
        if (poly_eval<DEG>(dec, s) > 0.0) {  // strict '>' (om3:644)
          const R v = (fma(sgn, sd[i], c1) + c2) * dinv;                                // payoff in date-N money
          c[i] = sticky ? -v : v;
          changed = true;
          cnt++;
          const unsigned long long b = (unsigned long long)__double_as_longlong(s);
          bnd = is_put ? (b > bnd ? b : bnd) : (b < bnd ? b : bnd);
        }
      }
      if (have_gram && !(sticky && signbit(c[i])) && sgn * sg[i] > kk) {
        const R y = fabs(c[i]) * dg;
        moments_accumulate<DEG>(acc, (double)sg[i] * a.invK, (double)y);
      }
    }
    if (changed) VecMem<R, VEC>::store(cfp + j, c);
  }
  if (a.S_dec != nullptr) {
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    bnd = is_put ? warp_max_u64(bnd) : warp_min_u64(bnd);
    if ((threadIdx.x & 31) == 0 && cnt) {
      atomicAdd(a.exc_dec, (unsigned long long)cnt);
      if (is_put) atomicMax(a.bnd_dec, bnd); else atomicMin(a.bnd_dec, bnd);
    }
  }
  if (!have_gram) return;
  block_reduce_sum<Q, kSplitWarps>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) a.partials[(size_t)blockIdx.x * Q + q] = acc[q];
    __threadfence();
    is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    __shared__ double tot[Q];
    reduce_partials_last_block<Q>(a.partials, gridDim.x, tot, kSplitWarps);
    __syncthreads();
    if (threadIdx.x == 0) {
      double mom[Q], beta[DEG + 1];
      for (int q = 0; q < Q; ++q) mom[q] = tot[q];
      const bool ok = solve_poly<DEG>(mom, beta);
      for (int i = 0; i <= DEG; ++i) a.beta_out[i] = ok ? beta[i] : nan("");
      for (int i = DEG + 1; i < kMaxBeta; ++i) a.beta_out[i] = nan("");
      *a.nitm_out = (long long)(mom[0] + 0.5);
      *a.valid_out = ok ? 1 : 0;
      *a.ticket = 0u;
    }
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

template <typename R, int DEG>
__global__ void __launch_bounds__(kSplitThreads)
lsm_update_kernel(const R* __restrict__ S_t, R* __restrict__ cf, long long M, R dinv, double K, double Kh, double Kl,
                  double invK, int is_put, int sticky, const double* __restrict__ beta_t,
                  const int* __restrict__ valid_t, unsigned long long* bnd_t, unsigned long long* exc_t) {
  if (*valid_t == 0) return;  // no regression at this date: nothing changes (cash-flows are in date-N money)
  double beta[DEG + 1];
#pragma unroll
  for (int i = 0; i <= DEG; ++i) beta[i] = beta_t[i];
  // decision polynomial in the raw price, exactly as the persistent sweep forms it
  double dec[DEG + 1];
  {
    double sc = 1.0;
#pragma unroll
    for (int i = 0; i <= DEG; ++i) {
      double d = -beta[i] * sc;
      if (i == 0) d += is_put ? K : -K;
      if (i == 1) d += is_put ? -1.0 : 1.0;
      dec[i] = d;
      sc *= invK;
    }
  }
  const R sgn = is_put ? (R)-1 : (R)1;
  const R c1 = (R)(is_put ? Kh : -Kh), c2 = (R)(is_put ? Kl : -Kl);
  unsigned long long bnd = bnd_none(is_put);
  unsigned int cnt = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x) {
    const R c = cf[j];
    if (sticky && signbit(c)) continue;
    const R sr = S_t[j];
    const double s = (double)sr;
    const double pay = payoff<double>(s, K, is_put != 0);
    if (pay > 0.0 && poly_eval<DEG>(dec, s) > 0.0) {  // strict '>' (om3:644)
      const R a = (fma(sgn, sr, c1) + c2) * dinv;     // payoff in date-N money
      cf[j] = sticky ? -a : a;
      cnt++;
      const unsigned long long b = (unsigned long long)__double_as_longlong(s);
      bnd = is_put ? (b > bnd ? b : bnd) : (b < bnd ? b : bnd);
    }
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
lsm_final_kernel(const R* __restrict__ cf, long long M, double d1, double* partials, unsigned int* ticket,
                 double* sums_out) {
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
    __syncthreads();
    if (threadIdx.x == 0) {  // date-N money -> value after N - 1 discounts (om3:651)
      sums_out[0] *= d1;
      sums_out[1] *= d1 * d1;
      sums_out[2] = (double)M;
      *ticket = 0u;
    }
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
                                       unsigned long long* exc, long long* nitm, int* valid,
                                       unsigned long long* xchg, int xchg_words, int* flags) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  // the persistent sweep's exchange accumulators and overflow flag start every sweep at zero
  for (int i = t; i < xchg_words; i += gridDim.x * blockDim.x) xchg[i] = 0ull;
  if (t < 4) flags[t] = 0;
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

int sweep_reset_stats(optmc_ctx* ctx) {
  const SweepDesc& sw = ctx->sw;
  const int n1 = sw.N + 1;
  lsm_reset_stats_kernel<<<(n1 + 127) / 128, 128, 0, ctx->stream>>>(n1, sw.lp.is_put, ctx->d_betas, ctx->d_bnd,
                                                                    ctx->d_exc, ctx->d_nitm, ctx->d_valid,
                                                                    reinterpret_cast<unsigned long long*>(ctx->xchg),
                                                                    (int)(xchg_bytes() / 8), ctx->d_flags);
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
  rc = sweep_reset_stats(ctx);
  if (rc) return rc;
  const R* S_N = static_cast<const R*>(sw.S) + (size_t)sw.N * sw.ld;
  lsm_init_kernel<R><<<grid, kSplitThreads, 0, ctx->stream>>>(S_N, static_cast<R*>(ctx->cf), sw.M, sw.lp.K,
                                                              sw.Kh, sw.Kl, sw.lp.is_put);
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
                                                                   (R)sw.Dt[t], sw.lp.K, 1.0 / sw.lp.K, sw.lp.is_put,
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
      S_t, static_cast<R*>(ctx->cf), sw.M, (R)sw.Dinv[t], sw.lp.K, sw.Kh, sw.Kl, 1.0 / sw.lp.K, sw.lp.is_put, sticky,
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

// The whole split sweep of the bound option with the fused streaming kernel: N launches instead of 3 (N - 1).
template <typename R, int DEG> static int sweep_split_fused_t(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  int rc = sweep_begin(ctx);
  if (rc) return rc;
  const R* S = static_cast<const R*>(sw.S);
  const bool vec4 = (sw.M % 4 == 0) && (sw.ld % 4 == 0) && ((uintptr_t)sw.S % 16 == 0) && ((uintptr_t)ctx->cf % 16 == 0);
  const long long units = vec4 ? sw.M / 4 : sw.M;
  long long g = (units + kSplitThreads - 1) / kSplitThreads;
  const long long cap = (long long)ctx->sm_count * 8;
  const int grid = (int)(g < 1 ? 1 : (g > cap ? cap : g));
  rc = ensure_bytes((void**)&ctx->partials, &ctx->partials_bytes, (size_t)grid * 16 * sizeof(double));
  if (rc) return rc;
  StreamArgs a{};
  a.cf = ctx->cf; a.M = sw.M; a.K = sw.lp.K; a.Kh = sw.Kh; a.Kl = sw.Kl; a.invK = 1.0 / sw.lp.K;
  a.is_put = sw.lp.is_put; a.sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) ? 1 : 0;
  a.kk = (sw.lp.is_put ? -1.0 : 1.0) * strike_consts(sw.lp.K, sw.lp.is_put != 0, sizeof(R) == 4).Kcmp;
  a.partials = ctx->partials; a.ticket = ctx->tickets;
  for (int t = sw.N - 1; t >= 0; --t) {  // launch t: decision of date t+1 (if any) + regression of date t (if t >= 1)
    const int td = t + 1;
    const bool dec = td <= sw.N - 1, gram = t >= 1;
    if (!dec && !gram) continue;
    a.S_dec = dec ? S + (size_t)td * sw.ld : nullptr;
    a.S_gram = gram ? S + (size_t)t * sw.ld : nullptr;
    a.dinv = dec ? sw.Dinv[td] : 1.0;
    a.dg = gram ? sw.Dt[t] : 1.0;
    a.beta_dec = ctx->d_betas + (size_t)td * kMaxBeta; a.valid_dec = ctx->d_valid + td;
    a.bnd_dec = ctx->d_bnd + td; a.exc_dec = ctx->d_exc + td;
    a.beta_out = ctx->d_betas + (size_t)t * kMaxBeta; a.nitm_out = ctx->d_nitm + t; a.valid_out = ctx->d_valid + t;
    if (vec4) lsm_stream_kernel<R, DEG, 4><<<grid, kSplitThreads, 0, ctx->stream>>>(a);
    else lsm_stream_kernel<R, DEG, 1><<<grid, kSplitThreads, 0, ctx->stream>>>(a);
    ctx->launches++; sw.n_launches++;
  }
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

int sweep_split_fused(optmc_ctx* ctx) {
  const SweepDesc& sw = ctx->sw;
  if (sw.dtype == OPTMC_F64) return sw.deg == 2 ? sweep_split_fused_t<double, 2>(ctx) : sweep_split_fused_t<double, 3>(ctx);
  return sw.deg == 2 ? sweep_split_fused_t<float, 2>(ctx) : sweep_split_fused_t<float, 3>(ctx);
}

int sweep_finish(optmc_ctx* ctx, double* sums_out) {
  SweepDesc& sw = ctx->sw;
  const int grid = split_grid(ctx, sw.M);
  if (sw.dtype == OPTMC_F64)
    lsm_final_kernel<double><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const double*>(ctx->cf), sw.M,
                                                                      sw.Dt[sw.N >= 1 ? 1 : 0], ctx->partials, ctx->tickets,
                                                                      sums_out);
  else
    lsm_final_kernel<float><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const float*>(ctx->cf), sw.M,
                                                                     sw.Dt[sw.N >= 1 ? 1 : 0], ctx->partials, ctx->tickets,
                                                                     sums_out);
  ctx->launches++; sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

// Paths whose final cash-flow is exactly zero ("expires worthless", om1:168) -- valid after a sweep that keeps the
// cash-flows in HBM (split / network sweeps).
template <typename R> __global__ void lsm_zero_count_kernel(const R* __restrict__ cf, long long M, unsigned long long* out) {
  unsigned int c = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (long long)gridDim.x * blockDim.x)
    c += cf[j] == (R)0 ? 1u : 0u;  // -0.0 (an exercised zero payoff cannot occur: exercise needs payoff > continuation)
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

int sweep_zero_count(optmc_ctx* ctx, int64_t* count) {
  SweepDesc& sw = ctx->sw;
  if (!count) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!sw.have_results || sw.impl_used != OPTMC_SWEEP_SPLIT || !ctx->cf) {
    set_error("zero-cash-flow count needs a sweep that keeps its cash-flows in device memory (impl SPLIT, optmc_lsm_mlp)");
    return OPTMC_EINVAL;
  }
  unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->gram);  // scratch, 16 doubles
  OPTMC_CUDA(cudaMemsetAsync(d, 0, 8, ctx->stream));
  const int grid = split_grid(ctx, sw.M);
  if (sw.dtype == OPTMC_F64) lsm_zero_count_kernel<double><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const double*>(ctx->cf), sw.M, d);
  else lsm_zero_count_kernel<float><<<grid, kSplitThreads, 0, ctx->stream>>>(static_cast<const float*>(ctx->cf), sw.M, d);
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  OPTMC_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  *count = (int64_t)h;
  return OPTMC_OK;
}

int sweep_finalize_price(optmc_ctx* ctx, const double* sums) {
  lsm_price_from_sums_kernel<<<1, 1, 0, ctx->stream>>>(sums, ctx->sw.final_scale, ctx->d_final);
  ctx->launches++; ctx->sw.n_launches++;
  OPTMC_CUDA(cudaGetLastError());
  return OPTMC_OK;
}

}  // namespace optmc
