// lsm_resident.cu -- the persistent LSM sweep: ONE cooperative launch walks every exercise date.
//
// Data layout and flow (per CTA, one CTA per SM, 512 threads):
//   * the CTA owns a contiguous slice of paths; their cash-flows live in REGISTERS for the whole sweep
//     (PPT values per thread; the reference's `exercised` flag, om3:617/649, is the sign bit);
//   * each date's slice of the step-major price slab is staged into shared memory by the TMA engine
//     (1-D cp.async.bulk + mbarrier), 2-3 dates ahead of use, so HBM sees S exactly once;
//   * per date: masked Gram moments in fp64 (thread) -> recursive-halving reduce-scatter (warp) -> shared
//     memory (block) -> all-gather of the CTA totals through L2 -> guarded LDL^T solve -> fused
//     exercise decision / cash-flow update.
//   * the cross-CTA exchange is an LL-style protocol: every double travels as two 8-byte words
//     {payload32, epoch32}; a word is valid when its epoch matches, so there are no fences, no atomics and
//     no grid-wide barrier on the data path, and all 16 warps gather in parallel.  Summation order is fixed
//     => bit-reproducible betas on every CTA and from run to run.
//
// Replaces the Python loop of om3:615-651 (= om3:485-500, om2:278-310) for the polynomial regressor of
// SURVEY.md 8(c).  Semantics flags: sticky mask (om3:621,649), N-1 discounts (om3:619-620,651).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

constexpr int kResWarps = kResThreads / 32;
constexpr int kGatherChunk = 96;                                  // CTAs per gather item: 3 per lane
constexpr int kGatherParts = (kMaxResidentCtas + kGatherChunk - 1) / kGatherChunk;  // 2

struct ResArgs {
  const void* S;
  long long ld, M, chunk;
  int N, nstage;
  unsigned int stage_stride;  // bytes between stages in shared memory
  double K, invK, disc, final_scale;
  double Kcmp;                // float-exact threshold equivalent to K for the fp32 ITM test
  int is_put, sticky;
  unsigned long long* xw;     // exchange words [2][kXchgMaxQ][kMaxResidentCtas][2]
  unsigned int epoch_base;
  double* betas;              // [(N+1)][kMaxBeta]
  unsigned long long* bnd;    // [(N+1)]
  unsigned long long* exc;    // [(N+1)]
  long long* nitm;            // [(N+1)]
  double* final_out;          // [4]
  long long* trace;           // optional [2][(N+1)][8] phase clocks of the first and last CTA (OPTMC_TRACE)
};

#define OPTMC_TRACE_AT(ph)                                                                         \
  do {                                                                                             \
    if (a.trace && tid == 0 && (cta == 0 || cta == ncta - 1))                                      \
      a.trace[((size_t)(cta == 0 ? 0 : 1) * (N + 1) + t) * 8 + (ph)] = clock64();                  \
  } while (0)

template <int QN> struct Pow2 { static constexpr int v = QN <= 2 ? 2 : QN <= 4 ? 4 : QN <= 8 ? 8 : 16; };

__device__ __forceinline__ size_t xw_index(int par, int q, int cta) {
  return (((size_t)par * kXchgMaxQ + q) * kMaxResidentCtas + cta) * 2;
}

// Block-wide sum of QN per-thread doubles followed by the grid-wide all-gather.  Every thread calls it.
// On return s_tot[0..QN) holds the GRID totals (visible to every thread).  Shared scratch: s_red
// [kResWarps][QP], s_part [QN][kGatherParts].
template <int QN>
__device__ __forceinline__ void grid_allreduce(double (&acc)[Pow2<QN>::v], double* s_red, double* s_part,
                                               double* s_tot, unsigned long long* xw, int par, int cta, int ncta,
                                               unsigned int epoch, long long* tr = nullptr) {
  constexpr int QP = Pow2<QN>::v;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1. warp: recursive-halving reduce-scatter (QP + 3 adds per lane instead of 5 QP)
  warp_reduce_scatter<QP>(acc, lane);
  if (reduce_scatter_owner<QP>(lane)) s_red[warp * QP + reduce_scatter_index<QP>(lane)] = acc[0];
  __syncthreads();
  if (tr) tr[2] = clock64();
  // 2. warp 0: block total and publish.  lane -> quantity lane % QP, group lane / QP sums warps g, g+G, ...
  if (warp == 0) {
    constexpr int G = 32 / QP;
    const int q = lane % QP, g = lane / QP;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kResWarps; w += G)
      if (w + g < kResWarps) v += s_red[(w + g) * QP + q];
#pragma unroll
    for (int m = QP; m <= 16; m <<= 1) v += shfl_xor_f64(v, m);
    const double t = __shfl_sync(0xffffffffu, v, (lane >> 1) % QP);  // lane l publishes half (l & 1) of quantity l >> 1
    if (lane < 2 * QN) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
      const unsigned int payload = (lane & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
      st_relaxed_u64(xw + xw_index(par, lane >> 1, cta) + (lane & 1), ((unsigned long long)epoch << 32) | payload);
    }
  }
  if (tr) tr[3] = clock64();
  int spins = 0;
  // 3. every warp gathers: item = (quantity, part); lane handles CTAs part*96 + lane + 32 i
  for (int item = warp; item < QN * kGatherParts; item += kResWarps) {
    const int q = item / kGatherParts, part = item % kGatherParts;
    const unsigned long long* base = xw + xw_index(par, q, 0);
    double val[3];
    bool ready[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = part * kGatherChunk + lane + 32 * i;
      ready[i] = c >= ncta;
      val[i] = 0.0;
    }
    bool all;
    do {
      all = true;
      ++spins;
      unsigned long long w0[3], w1[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (!ready[i]) ld_relaxed_v2u64(base + (size_t)(part * kGatherChunk + lane + 32 * i) * 2, w0[i], w1[i]);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (!ready[i]) {
          if ((unsigned int)(w0[i] >> 32) == epoch && (unsigned int)(w1[i] >> 32) == epoch) {
            val[i] = __longlong_as_double((long long)((w1[i] << 32) | (w0[i] & 0xffffffffull)));
            ready[i] = true;
          } else {
            all = false;
          }
        }
      }
    } while (!all);
    double s = (val[0] + val[1]) + val[2];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += shfl_xor_f64(s, m);
    if (lane == 0) s_part[q * kGatherParts + part] = s;
  }
  if (tr) { tr[4] = clock64(); tr[7] = spins; }
  __syncthreads();
  if (tid < QN) {
    double s = s_part[tid * kGatherParts];
#pragma unroll
    for (int p = 1; p < kGatherParts; ++p) s += s_part[tid * kGatherParts + p];
    s_tot[tid] = s;
  }
}

template <typename R> __device__ __forceinline__ bool itm_test(R s, double K, double Kcmp, bool is_put);
template <> __device__ __forceinline__ bool itm_test<float>(float s, double, double Kcmp, bool is_put) {
  return is_put ? s < (float)Kcmp : s > (float)Kcmp;  // exact: Kcmp is the float bracket of K on the right side
}
template <> __device__ __forceinline__ bool itm_test<double>(double s, double K, double, bool is_put) {
  return is_put ? s < K : s > K;
}

template <typename R, int DEG, int PPT>
__global__ void __launch_bounds__(kResThreads, 1) lsm_resident_kernel(const ResArgs a) {
  constexpr int Q = Moments<DEG>::Q;
  constexpr int QP = Pow2<Q>::v;
  static_assert(Q <= kXchgMaxQ, "Gram vector must fit the exchange buffer");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3];
  __shared__ double s_red[kResWarps * QP];
  __shared__ double s_part[Q * kGatherParts];
  __shared__ double s_tot[Q];
  __shared__ double s_beta[DEG + 1];
  __shared__ int s_valid;
  __shared__ unsigned long long s_bnd[2];
  __shared__ unsigned int s_cnt[2];

  const int tid = threadIdx.x, lane = tid & 31;
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
    OPTMC_TRACE_AT(0);
    const R* st = stage_ptr(t);
    // -- discount every path (om3:620), then the ITM-masked Gram moments in fp64 (om3:621 mask) --
    double acc[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) acc[q] = 0.0;
    {
      double mom[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) mom[q] = 0.0;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int j = tid + k * kResThreads;
        if (j < n_local) {
          const R c = cf[k];
          const bool ex = sticky && signbit(c);
          const R y = fabs(c) * disc;
          cf[k] = ex ? -y : y;
          const R s = st[j];
          if (!ex && itm_test<R>(s, a.K, a.Kcmp, is_put)) moments_accumulate<DEG>(mom, (double)s * a.invK, (double)y);
        }
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[q] = mom[q];
    }
    OPTMC_TRACE_AT(1);
    // the first __syncthreads inside grid_allreduce also proves every thread is done with stage (t+1)
    long long* tr = (a.trace && tid == 0 && (cta == 0 || cta == ncta - 1))
                        ? a.trace + ((size_t)(cta == 0 ? 0 : 1) * (N + 1) + t) * 8 : nullptr;
    grid_allreduce<Q>(acc, s_red, s_part, s_tot, a.xw, t & 1, cta, ncta, a.epoch_base + (unsigned int)(N - t), tr);
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
      if (t + 1 - nstage >= 1) issue_load(t + 1 - nstage);  // refill the stage date t+1 vacated
    }
    __syncthreads();  // s_tot visible
    OPTMC_TRACE_AT(5);
    if (tid == 0) {
      double tot[Q], beta[DEG + 1];
#pragma unroll
      for (int q = 0; q < Q; ++q) tot[q] = s_tot[q];
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
    __syncthreads();  // beta visible
    OPTMC_TRACE_AT(6);
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
          const R sr = st[j];
          if (!(sticky && signbit(c)) && itm_test<R>(sr, a.K, a.Kcmp, is_put)) {
            const double s = (double)sr;
            const double pay = is_put ? a.K - s : s - a.K;
            if (pay > poly_eval<DEG>(beta, s * a.invK)) {   // strict '>' (om3:644)
              cf[k] = sticky ? -(R)pay : (R)pay;            // sticky flag = sign bit (om3:649)
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
  grid_allreduce<2>(fin, s_red, s_part, s_tot, a.xw, 0, cta, ncta, a.epoch_base + (unsigned int)N);
  if (tid == 0 && N - 1 >= 1) {  // statistics of date 1 (all update loops are behind the barriers above)
    if (s_cnt[1]) {
      atomicAdd(a.exc + 1, (unsigned long long)s_cnt[1]);
      if (is_put) atomicMax(a.bnd + 1, s_bnd[1]); else atomicMin(a.bnd + 1, s_bnd[1]);
    }
  }
  __syncthreads();
  if (cta == 0 && tid == 0) {
    const double n = (double)a.M;
    const double mean = s_tot[0] / n;
    double var = n > 1.0 ? (s_tot[1] - n * mean * mean) / (n - 1.0) : 0.0;
    if (var < 0.0) var = 0.0;
    a.final_out[0] = mean * a.final_scale;
    a.final_out[1] = sqrt(var / n) * a.final_scale;
    a.final_out[2] = s_tot[0];
    a.final_out[3] = s_tot[1];
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
  const size_t avail = (size_t)ctx->max_smem_optin - 6144;  // static shared + slack
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
      break;
  }
  set_error("no resident instantiation for this slice size");
  return OPTMC_EUNSUPPORTED;
}

int sweep_resident(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  ResPlan p;
  std::string why;
  if (!plan_resident(ctx, sw, &p, &why)) { set_error("resident sweep unavailable: " + why); return OPTMC_EUNSUPPORTED; }
  int rc = sweep_reset_stats(ctx);
  if (rc) return rc;
  if (ctx->epoch > 0xFFFF0000u - (unsigned)sw.N) {  // 32-bit epoch about to wrap: start a fresh era
    OPTMC_CUDA(cudaMemsetAsync(ctx->xchg, 0, xchg_bytes(), ctx->stream));
    ctx->epoch = 0;
  }
  ResArgs a{};
  a.S = sw.S; a.ld = sw.ld; a.M = sw.M; a.chunk = p.chunk; a.N = sw.N; a.nstage = p.nstage;
  a.stage_stride = p.stage_stride;
  a.K = sw.lp.K; a.invK = 1.0 / sw.lp.K; a.disc = sw.disc; a.final_scale = sw.final_scale;
  {  // float threshold with (s < K) <=> (s < Kcmp) for every float s (puts); mirrored for calls
    float kf = (float)sw.lp.K;
    if (sw.lp.is_put) { if ((double)kf < sw.lp.K) kf = nextafterf(kf, INFINITY); }
    else { if ((double)kf > sw.lp.K) kf = nextafterf(kf, -INFINITY); }
    a.Kcmp = (double)kf;
  }
  a.is_put = sw.lp.is_put; a.sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) ? 1 : 0;
  a.xw = reinterpret_cast<unsigned long long*>(ctx->xchg); a.epoch_base = (unsigned int)ctx->epoch;
  a.betas = ctx->d_betas; a.bnd = ctx->d_bnd; a.exc = ctx->d_exc; a.nitm = ctx->d_nitm; a.final_out = ctx->d_final;
  ctx->epoch += (unsigned long long)sw.N + 2ull;
  // Debug aid: OPTMC_TRACE=<file> dumps per-date phase clocks (SM cycles) of the first and last CTA.
  const char* trace_path = getenv("OPTMC_TRACE");
  long long* d_trace = nullptr;
  const size_t trace_n = (size_t)2 * (sw.N + 1) * 8;
  if (trace_path && *trace_path) {
    OPTMC_CUDA(cudaMalloc((void**)&d_trace, trace_n * sizeof(long long)));
    OPTMC_CUDA(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), ctx->stream));
    a.trace = d_trace;
  }
  if (sw.dtype == OPTMC_F64)
    rc = sw.deg == 2 ? launch_resident_ppt<double, 2>(ctx, p, a) : launch_resident_ppt<double, 3>(ctx, p, a);
  else
    rc = sw.deg == 2 ? launch_resident_ppt<float, 2>(ctx, p, a) : launch_resident_ppt<float, 3>(ctx, p, a);
  if (d_trace) {
    std::vector<long long> h(trace_n);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_trace);
    if (e == cudaSuccess) {
      if (FILE* f = fopen(trace_path, "w")) {
        fprintf(f, "# ncta=%d chunk=%lld ppt=%d nstage=%d N=%d ; columns: cta t ph0..ph7 (clock64)\n", p.ncta, p.chunk,
                p.ppt, p.nstage, sw.N);
        for (int c = 0; c < 2; ++c)
          for (int t = sw.N - 1; t >= 1; --t) {
            fprintf(f, "%d %d", c, t);
            for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[((size_t)c * (sw.N + 1) + t) * 8 + k]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
  }
  return rc;
}

}  // namespace optmc
