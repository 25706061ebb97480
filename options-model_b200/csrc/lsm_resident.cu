// lsm_resident.cu -- the persistent LSM sweep: ONE cooperative launch walks every exercise date.
//
// Data layout and flow (per CTA, one CTA per SM, 512 threads):
//   * the CTA owns a contiguous slice of paths; their cash-flows live in REGISTERS for the whole sweep
//     (PPT values per thread; the reference's `exercised` flag, om3:617/649, is the sign bit);
//   * each date's slice of the step-major price slab is staged into shared memory by the TMA engine
//     (1-D cp.async.bulk + mbarrier), 2-3 dates ahead of use, so HBM sees S exactly once;
//   * per date: masked Gram moments of the raw price in fp64 (thread) -> recursive-halving reduce-scatter
//     (warp) -> shared memory (block) -> grid-wide sum -> guarded LDL^T solve (warp 0) -> fused exercise
//     decision / cash-flow update.
//   * the grid-wide sum is ORDER-INDEPENDENT: every block total is split into two 48-bit fixed-point
//     chunks (optmc_math.cuh: fx_encode) and added with one integer `red` per chunk into a per-parity
//     accumulator word whose top 8 bits count arrivals.  Sum and completion flag are the same word, so
//     there is no fence, no second flag and no grid-wide barrier on the data path; one lane polls each
//     word.  Integer addition commutes => betas are bit-reproducible on every CTA and from run to run.
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

struct ResArgs {
  const void* S;
  long long ld, M, chunk;
  int N, nstage;
  unsigned int stage_stride;  // bytes between stages in shared memory
  double K, invK, disc, final_scale;
  double sgn, kk, c1, c2;     // storage-precision pass constants (see Store<>); exact in the storage type
  int is_put, sticky;
  unsigned long long* xw;     // exchange accumulators [2][kXchgWords][kXchgStride]
  int* flags;                 // [0] = exchange overflow
  double* betas;              // [(N+1)][kMaxBeta]
  unsigned long long* bnd;    // [(N+1)]
  unsigned long long* exc;    // [(N+1)]
  long long* nitm;            // [(N+1)]
  double* final_out;          // [4]
  long long* trace;           // optional [2][(N+1)][8] phase clocks of the first and last CTA (OPTMC_TRACE)
};

template <int QN> struct Pow2 { static constexpr int v = QN <= 2 ? 2 : QN <= 4 ? 4 : QN <= 8 ? 8 : 16; };

// ---- storage-type helpers ------------------------------------------------------------------------------
// Put and call share one instruction stream: with sgn = -1 (put) / +1 (call)
//   in the money   <=>  sgn * s > kk            (kk = sgn * Kcmp; products with +-1 are exact)
//   payoff          =   fma(sgn, s, c1) + c2    (c1 = -sgn * Kh, c2 = -sgn * Kl, K = Kh + Kl)
// For fp64 storage Kh = K, Kl = 0 and the payoff is the correctly rounded K - s / s - K of the reference
// (om3:376-380); for fp32 storage it is that whenever K is a float, and within one ulp otherwise.
// The reference's `exercised` flag (om3:617/649) is the sign bit of the stored cash-flow; `flag` is the
// sign-bit mask under the sticky semantics and 0 otherwise.
template <typename R> struct Store;
template <> struct Store<float> {
  static __device__ __forceinline__ bool flagged(float c, unsigned int flag) { return (__float_as_uint(c) & flag) != 0u; }
  static __device__ __forceinline__ float with_flag(float p, unsigned int flag) {
    return __uint_as_float(__float_as_uint(p) | flag);
  }
};
template <> struct Store<double> {
  static __device__ __forceinline__ bool flagged(double c, unsigned int flag) {
    return ((unsigned int)__double2hiint(c) & flag) != 0u;
  }
  static __device__ __forceinline__ double with_flag(double p, unsigned int flag) {
    return __hiloint2double((int)((unsigned int)__double2hiint(p) | flag), __double2loint(p));
  }
};

// Block-wide sum of QP (power of two) per-thread doubles.  Every thread calls it; contains one
// __syncthreads.  On return, in warp 0, lane l < 2*QP holds the CTA total of quantity l >> 1.
template <int QP>
__device__ __forceinline__ double block_totals(double (&acc)[QP], double* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  warp_reduce_scatter<QP>(acc, lane);
  if (reduce_scatter_owner<QP>(lane)) s_red[warp * QP + reduce_scatter_index<QP>(lane)] = acc[0];
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {  // lane -> quantity lane % QP; group lane / QP sums warps g, g+G, ...
    constexpr int G = 32 / QP;
    const int q = lane % QP, g = lane / QP;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kResWarps; w += G)
      if (w + g < kResWarps) v += s_red[(w + g) * QP + q];
#pragma unroll
    for (int m = QP; m <= 16; m <<= 1) v += shfl_xor_f64(v, m);
    t = __shfl_sync(0xffffffffu, v, (lane >> 1) % QP);
  }
  return t;
}

// Warp 0 only.  `mine`: this CTA's total of quantity lane >> 1 (lanes < 2*QN), already in its final
// units.  Adds it into the parity's accumulators, waits until all `ncta` CTAs have arrived and returns
// the grid total of quantity `lane` in lanes < QN.  prev = this lane's accumulator value after the last
// completed exchange of the same parity (0 at launch).
template <int QN>
__device__ __forceinline__ double warp0_grid_sum(double mine, unsigned long long* xw, int par, int ncta,
                                                 unsigned long long& prev, int* flags, int* spins_out) {
  const int lane = threadIdx.x & 31;
  unsigned long long sum = 0ull;
  int spins = 0;
  if (lane < 2 * QN) {
    unsigned long long hi, lo;
    if (!fx_encode(mine, hi, lo)) atomicExch(flags, 1);
    unsigned long long* w = xw + ((size_t)par * kXchgWords + lane) * kXchgStride;
    red_relaxed_add_u64(w, (1ull << kFxCountShift) | ((lane & 1) ? lo : hi));
    unsigned long long d;
    do {
      d = ld_relaxed_u64(w) - prev;
      ++spins;
    } while ((d >> kFxCountShift) != (unsigned long long)ncta);
    prev += d;
    sum = d & kFxValueMask;
  }
  __syncwarp();
  const unsigned long long other = __shfl_down_sync(0xffffffffu, sum, 1);
  double tot = fx_decode(sum, other, ncta);                 // meaningful on even lanes < 2*QN
  tot = __shfl_sync(0xffffffffu, tot, (2 * lane) & 31);     // quantity q: lane 2q -> lane q
  if (spins_out) *spins_out = spins;
  return tot;
}

#define OPTMC_TRACE_AT(ph)            \
  do {                                \
    if (tr) tr[(ph)] = clock64();     \
  } while (0)

template <typename R> struct PassConsts {
  R sgn, kk, c1, c2, disc;
  unsigned int flag;
  int n_local;
};

// One branch-free pass over the thread's PPT paths (path j = tid + k * kResThreads):
//   DECIDE: exercise decision of date t -- exercise iff dec(S_t) > 0 (payoff - continuation, strict, om3:644);
//   GRAM:   discount (om3:620), then the ITM-masked (om3:621) raw-price moments of date t-1.
// Dead lanes contribute zeros, so the fp64 pipeline sees one straight-line stream per path.
template <typename R, int DEG, int PPT, bool DECIDE, bool GRAM>
__device__ __forceinline__ void fused_pass(R (&cf)[PPT], const R* __restrict__ st_t, const R* __restrict__ st_g,
                                           const double (&dec)[DEG + 1], const PassConsts<R>& pc,
                                           double (&mom)[Moments<DEG>::Q], unsigned int& rows, unsigned int& cnt,
                                           R& em) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int j = tid + k * kResThreads;
    const bool in = j < pc.n_local;
    R c = cf[k];
    if (DECIDE) {
      const R sr = in ? st_t[j] : (R)0;
      const R u = pc.sgn * sr;
      const bool live = in & !Store<R>::flagged(c, pc.flag) & (u > pc.kk);
      const bool exer = live & (poly_eval<DEG>(dec, (double)sr) > 0.0);
      const R pay = Store<R>::with_flag(fma(pc.sgn, sr, pc.c1) + pc.c2, pc.flag);  // sticky flag = sign bit (om3:649)
      c = exer ? pay : c;
      cnt += exer ? 1u : 0u;
      em = fmax(em, exer ? -u : (R)-INFINITY);
    }
    if (GRAM) {
      const bool ex = Store<R>::flagged(c, pc.flag);
      const R y = fabs(c) * pc.disc;
      c = ex ? -y : y;
      const R s = in ? st_g[j] : (R)0;
      const bool live = in & !ex & (pc.sgn * s > pc.kk);
      rows += live ? 1u : 0u;
      moments_accumulate_nocount<DEG>(mom, (double)(live ? s : (R)0), (double)(live ? y : (R)0));
    }
    cf[k] = c;
  }
}

template <typename R, int DEG, int PPT>
__global__ void __launch_bounds__(kResThreads, 1) lsm_resident_kernel(const ResArgs a) {
  constexpr int Q = Moments<DEG>::Q;
  constexpr int QP = Pow2<Q>::v;
  static_assert(Q <= kXchgMaxQ, "Gram vector must fit the exchange buffer");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3];
  __shared__ double s_red[kResWarps * 16];
  __shared__ double s_dec[DEG + 1];   // exercise iff s_dec(s) > 0  (payoff - continuation as a polynomial in S)
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
  const R sgn = (R)a.sgn, kk = (R)a.kk, c1 = (R)a.c1, c2 = (R)a.c2;
  const unsigned int flag = sticky ? 0x80000000u : 0u;
  const R disc = (R)a.disc;
  const int N = a.N, nstage = a.nstage;
  const R* Sbase = static_cast<const R*>(a.S) + base;

  auto stage_ptr = [&](int t) -> const R* {
    return reinterpret_cast<const R*>(smem_raw + (size_t)(t % nstage) * a.stage_stride);
  };
  auto issue_load = [&](int t) {  // one thread
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
    s_valid = 0;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i)
      if (N - i >= 1) issue_load(N - i);
  }

  // warp-0 lane constants: lane l serves quantity l >> 1; raw-price moments are rescaled to x = S/K
  // (the regressor of SURVEY.md 8(c)) by invK^power before they enter the exchange.
  double qscale = 1.0;
  {
    const int pw = moment_power<DEG>(lane >> 1);
    for (int i = 0; i < pw; ++i) qscale *= a.invK;
  }
  unsigned long long prev0 = 0ull, prev1 = 0ull;  // warp 0: accumulator baselines of the two parities
  long long* const tr_base = (a.trace && tid == 0 && (cta == 0 || cta == ncta - 1))
                                 ? a.trace + (size_t)(cta == 0 ? 0 : 1) * (N + 1) * 8 : nullptr;

  // ---- date N: cash-flows = payoff(S[N]) (om3:616) ----
  R cf[PPT];
  wait_stage(N);
  {
    const R* st = stage_ptr(N);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int j = tid + k * kResThreads;
      const bool in = j < n_local;
      const R s = in ? st[j] : (R)0;
      const R p = fma(sgn, s, c1) + c2;
      cf[k] = (in & (sgn * s > kk)) ? p : (R)0;
    }
  }

  // Iteration t (t = N .. 1) makes ONE branch-free pass over the CTA's paths:
  //   (1) exercise decision of date t with the polynomial solved at the end of iteration t+1 (none at t = N),
  //   (2) discount (om3:620) and the ITM-masked raw-price moments of date t-1 (om3:621 mask) -- skipped at t = 1,
  // followed by the block reduction, the grid sum and the solve for date t-1.
  int seq = 0;
  for (int t = N; t >= 1; --t) {
    long long* tr = tr_base ? tr_base + (size_t)t * 8 : nullptr;
    const bool gram = t >= 2;
    const bool decide = t <= N - 1 && s_valid != 0;
    if (gram) wait_stage(t - 1);
    OPTMC_TRACE_AT(0);
    const R* st_t = stage_ptr(t);
    const R* st_g = stage_ptr(gram ? t - 1 : t);
    double dec[DEG + 1];
#pragma unroll
    for (int i = 0; i <= DEG; ++i) dec[i] = decide ? s_dec[i] : 0.0;
    double mom[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) mom[q] = 0.0;
    unsigned int rows = 0, cnt = 0;
    R em = (R)-INFINITY;  // max over exercised paths of -sgn * S: put -> max S, call -> -(min S)
    const PassConsts<R> pc{sgn, kk, c1, c2, disc, flag, n_local};
    if (decide && gram) fused_pass<R, DEG, PPT, true, true>(cf, st_t, st_g, dec, pc, mom, rows, cnt, em);
    else if (gram) fused_pass<R, DEG, PPT, false, true>(cf, st_t, st_g, dec, pc, mom, rows, cnt, em);
    else if (decide) fused_pass<R, DEG, PPT, true, false>(cf, st_t, st_g, dec, pc, mom, rows, cnt, em);
    if (decide) {
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt) {  // warp-uniform
        const double ext = -(double)sgn * (double)em;
        unsigned long long b = (unsigned long long)__double_as_longlong(ext);
        if (!isfinite(ext)) b = bnd_none(a.is_put);
        b = is_put ? warp_max_u64(b) : warp_min_u64(b);
        if (lane == 0) {
          atomicAdd(&s_cnt[0], cnt);
          if (is_put) atomicMax(&s_bnd[0], b); else atomicMin(&s_bnd[0], b);
        }
      }
    }
    OPTMC_TRACE_AT(1);
    if (!gram) break;

    double acc[QP];
    mom[0] = (double)rows;
#pragma unroll
    for (int q = 0; q < QP; ++q) acc[q] = q < Q ? mom[q] : 0.0;
    // the __syncthreads inside block_totals also proves every thread is done with stage t
    const double mine = block_totals<QP>(acc, s_red) * qscale;
    OPTMC_TRACE_AT(2);
    if (tid == 32) {  // bookkeeping off the critical path (warp 1)
      if (s_cnt[0]) {  // exercise statistics of date t
        atomicAdd(a.exc + t, (unsigned long long)s_cnt[0]);
        if (is_put) atomicMax(a.bnd + t, s_bnd[0]); else atomicMin(a.bnd + t, s_bnd[0]);
        s_cnt[0] = 0u;
        s_bnd[0] = bnd_none(a.is_put);
      }
      if (t - nstage >= 1) issue_load(t - nstage);  // refill the stage date t vacated
    }
    if (warp == 0) {
      int spins = 0;
      unsigned long long pv = (seq & 1) ? prev1 : prev0;
      const double tot_l = warp0_grid_sum<Q>(mine, a.xw, seq & 1, ncta, pv, a.flags, &spins);
      if (seq & 1) prev1 = pv; else prev0 = pv;
      OPTMC_TRACE_AT(4);
      if (tr) tr[7] = spins;
      double tot[Q], beta[DEG + 1];
#pragma unroll
      for (int q = 0; q < Q; ++q) tot[q] = __shfl_sync(0xffffffffu, tot_l, q);
      const bool ok = solve_poly<DEG>(tot, beta);  // every lane, identical inputs: no divergence
      if (lane == 0) {
        s_valid = ok ? 1 : 0;
        if (ok) {
          // payoff - continuation = (+-K - b0) + (-+1 - b1/K) S - (b2/K^2) S^2 ...  (x = S/K)
          double sc = 1.0;
#pragma unroll
          for (int i = 0; i <= DEG; ++i) {
            double d = -beta[i] * sc;
            if (i == 0) d += is_put ? a.K : -a.K;
            if (i == 1) d += is_put ? -1.0 : 1.0;
            s_dec[i] = d;
            sc *= a.invK;
          }
        }
        if (cta == 0) {
#pragma unroll
          for (int i = 0; i <= DEG; ++i) a.betas[(size_t)(t - 1) * kMaxBeta + i] = ok ? beta[i] : nan("");
          a.nitm[t - 1] = (long long)(tot[0] + 0.5);
        }
      }
      OPTMC_TRACE_AT(5);
    }
    ++seq;
    __syncthreads();  // decision polynomial of date t-1 visible
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
  const double mine = block_totals<2>(fin, s_red);
  if (tid == 32 && s_cnt[0]) {  // statistics of date 1 (its update pass is behind the barrier above)
    atomicAdd(a.exc + 1, (unsigned long long)s_cnt[0]);
    if (is_put) atomicMax(a.bnd + 1, s_bnd[0]); else atomicMin(a.bnd + 1, s_bnd[0]);
  }
  if (warp == 0) {
    unsigned long long pv = (seq & 1) ? prev1 : prev0;
    const double tot_l = warp0_grid_sum<2>(mine, a.xw, seq & 1, ncta, pv, a.flags, nullptr);
    const double s1 = __shfl_sync(0xffffffffu, tot_l, 0), s2 = __shfl_sync(0xffffffffu, tot_l, 1);
    if (cta == 0 && lane == 0) {
      const double n = (double)a.M;
      const double mean = s1 / n;
      double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      a.final_out[0] = mean * a.final_scale;
      a.final_out[1] = sqrt(var / n) * a.final_scale;
      a.final_out[2] = s1;
      a.final_out[3] = s2;
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
  int rc = sweep_reset_stats(ctx);  // also zeroes the exchange accumulators and the overflow flag
  if (rc) return rc;
  ResArgs a{};
  a.S = sw.S; a.ld = sw.ld; a.M = sw.M; a.chunk = p.chunk; a.N = sw.N; a.nstage = p.nstage;
  a.stage_stride = p.stage_stride;
  a.K = sw.lp.K; a.invK = 1.0 / sw.lp.K; a.disc = sw.disc; a.final_scale = sw.final_scale;
  {  // pass constants, exact in the storage type (see Store<>)
    const double sg = sw.lp.is_put ? -1.0 : 1.0;
    double Kcmp = sw.lp.K, Kh = sw.lp.K, Kl = 0.0;
    if (sw.dtype == OPTMC_F32) {
      // float threshold with (s < K) <=> (s < Kcmp) for every float s (puts); mirrored for calls
      float kf = (float)sw.lp.K;
      Kh = (double)kf;
      Kl = (double)(float)(sw.lp.K - Kh);
      if (sw.lp.is_put) { if ((double)kf < sw.lp.K) kf = nextafterf(kf, INFINITY); }
      else { if ((double)kf > sw.lp.K) kf = nextafterf(kf, -INFINITY); }
      Kcmp = (double)kf;
    }
    a.sgn = sg; a.kk = sg * Kcmp; a.c1 = -sg * Kh; a.c2 = -sg * Kl;
  }
  a.is_put = sw.lp.is_put; a.sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) ? 1 : 0;
  a.xw = reinterpret_cast<unsigned long long*>(ctx->xchg); a.flags = ctx->d_flags;
  a.betas = ctx->d_betas; a.bnd = ctx->d_bnd; a.exc = ctx->d_exc; a.nitm = ctx->d_nitm; a.final_out = ctx->d_final;
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
          for (int t = sw.N; t >= 1; --t) {
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
