// lsm_global.cu -- LSM with ONE global regression over all (date, path) rows: the structure of the reference's
// v3 pricer (om3:482-651, om3gpu:695-833) with its network replaced by linear least squares on the seven
// reference features [1, x, x^2, x^3, max(x-1,0), sqrt(tau), x sqrt(tau)] (om3:105-121).
//
//   pass 1  global_gram_kernel   one CTA per (date, path tile): ITM-masked moments of the rows of that date
//                                (om3:485-516: targets = discounted TERMINAL payoffs; `exercised` is never set),
//                                block-reduced in fp64 and added to per-date fixed-point accumulators (integer
//                                atomics: order-independent, bit-reproducible);
//           global_solve_kernel  combines the per-date moments with the date weights sqrt(tau), tau, solves the
//                                normal equations (guarded LDL^T that drops redundant columns) and tabulates the
//                                per-date decision polynomial payoff - continuation in the raw price;
//   pass 2  global_walk_kernel   every thread walks its paths back in time (om3:615-651: sticky mask, strict '>',
//                                N-1 discounts) -- no regression, no grid synchronisation.
// Both passes stream the slab exactly once (pass 1 also re-reads the terminal row, which stays in L2) and are
// independent across paths: HBM-bound by construction.
//
// Numerics.  Internally the regressors are powers of u = x - 1 (x = S/K) instead of x: the span is the same
// ([1,x,x^2,x^3] <-> [1,u,u^2,u^3]; x sqrt(tau) = u sqrt(tau) + sqrt(tau)), the normal equations are orders of
// magnitude better conditioned, and the coefficients are converted back to the reference features for output.
// Within the ITM rows the hinge feature max(x-1,0) is identically 0 (puts) or x-1 (calls), i.e. redundant; it is
// dropped (beta = 0), which is also what a minimum-norm least-squares solution predicts with.  The z-scoring of
// om3:550-563 is an affine change of variables and does not change a linear model's predictions.
#include <math.h>
#include <string.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

constexpr int kGQ = 12;          // per-date moments: A0..A6 = sum u^a, Y0..Y3 = sum y u^a, YY = sum y^2  (y in units of K)
constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;
constexpr int kGCols = 6;        // [1, u, u^2, u^3, s, u s]

struct DecEntry {  // per-date decision polynomial in the raw price: exercise iff d(S) > 0
  double d[4];
  float f[4], b[4];  // fp32 coefficients and error-bound coefficients (see Decider<float, DEG>)
  double dinv;       // 1 / D_t: payoff -> date-N money
  double pad;
};

struct GlobalModel {
  double beta_ref[7];
  double n_rows;
  int rank, valid;
};

template <typename R> struct Vec4IO;
template <> struct Vec4IO<float> {
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
};
template <> struct Vec4IO<double> {
  static __device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};
template <typename R, int VEC> __device__ __forceinline__ void load_vec(const R* p, R (&v)[VEC]) {
  if constexpr (VEC == 4) Vec4IO<R>::load4(p, v);
  else v[0] = p[0];
}

// One (date, path) row of pass 1, branch-free: rows that are out of the money contribute zeros.
template <typename R>
__device__ __forceinline__ void gram_row(double (&acc)[kGQ], unsigned int& rows, R s, R sn, R sgn, R kk, double K,
                                         double invK, double dsc, bool is_put) {
  const bool itm = sgn * s > kk;  // om3:492 (no `exercised` in pass 1)
  rows += itm ? 1u : 0u;
  const double u = itm ? fma((double)s, invK, -1.0) : 0.0;
  const double y = itm ? payoff<double>((double)sn, K, is_put) * dsc : 0.0;
  const double u2 = u * u, u3 = u2 * u;
  acc[1] += u; acc[2] += u2; acc[3] += u3;
  acc[4] = fma(u2, u2, acc[4]); acc[5] = fma(u2, u3, acc[5]); acc[6] = fma(u3, u3, acc[6]);
  acc[7] += y; acc[8] = fma(y, u, acc[8]); acc[9] = fma(y, u2, acc[9]); acc[10] = fma(y, u3, acc[10]);
  acc[11] = fma(y, y, acc[11]);
}

template <typename R, int VEC>
__global__ void __launch_bounds__(kGThreads)
global_gram_kernel(const R* __restrict__ S, long long ld, long long M, int N, double K, double invK, int is_put,
                   double sgn_d, double kk_d, const double* __restrict__ Dt, unsigned long long* part, int* flags) {
  __shared__ double red[kGWarps * 16];
  __shared__ double tot[kGQ];
  const int t = N - 1 - (int)blockIdx.y;
  const R* __restrict__ St = S + (size_t)t * ld;
  const R* __restrict__ SN = S + (size_t)N * ld;
  const double dsc = Dt[t] * invK;  // discounted terminal payoff in units of K
  const R sgn = (R)sgn_d, kk = (R)kk_d;
  const bool put = is_put != 0;
  double acc[kGQ];
#pragma unroll
  for (int q = 0; q < kGQ; ++q) acc[q] = 0.0;
  unsigned int rows = 0;
  const long long units = M / VEC;  // M % VEC == 0 (launcher)
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; j + stride < units; j += 2 * stride) {  // two independent pairs of vector loads in flight
    R a0[VEC], b0[VEC], a1[VEC], b1[VEC];
    load_vec<R, VEC>(St + j * VEC, a0); load_vec<R, VEC>(SN + j * VEC, b0);
    load_vec<R, VEC>(St + (j + stride) * VEC, a1); load_vec<R, VEC>(SN + (j + stride) * VEC, b1);
#pragma unroll
    for (int i = 0; i < VEC; ++i) gram_row<R>(acc, rows, a0[i], b0[i], sgn, kk, K, invK, dsc, put);
#pragma unroll
    for (int i = 0; i < VEC; ++i) gram_row<R>(acc, rows, a1[i], b1[i], sgn, kk, K, invK, dsc, put);
  }
  for (; j < units; j += stride) {
    R a0[VEC], b0[VEC];
    load_vec<R, VEC>(St + j * VEC, a0); load_vec<R, VEC>(SN + j * VEC, b0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) gram_row<R>(acc, rows, a0[i], b0[i], sgn, kk, K, invK, dsc, put);
  }
  acc[0] = (double)rows;
  {  // block totals: recursive-halving warp reduce-scatter (16 + 3 adds instead of 5 x 12), then shared memory
    double a16[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a16[q] = q < kGQ ? acc[q] : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_scatter<16>(a16, lane);
    if (reduce_scatter_owner<16>(lane)) red[warp * 16 + reduce_scatter_index<16>(lane)] = a16[0];
    __syncthreads();
    if (threadIdx.x < kGQ) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kGWarps; ++w) v += red[w * 16 + threadIdx.x];
      tot[threadIdx.x] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kGQ) {
    unsigned long long hi, lo;
    if (!fx_encode(tot[threadIdx.x >> 1], hi, lo)) atomicExch(flags, 1);
    atomicAdd(part + (size_t)t * 2 * kGQ + threadIdx.x, (threadIdx.x & 1) ? lo : hi);
  }
}

// LDL^T on a p x p SPD system that DROPS a column whose pivot is <= rtol * its own diagonal (the column is a
// linear combination of the kept ones): beta = 0 there.  numpy's minimum-norm lstsq agrees on the predictions.
// Returns the number of kept columns.
template <int P> __device__ int solve_drop(const double (&G)[P][P], const double (&g)[P], double (&beta)[P], double rtol) {
  double L[P][P], d[P], z[P];
  bool keep[P];
  int rank = 0;
  for (int k = 0; k < P; ++k) {
    double s = G[k][k];
    for (int j = 0; j < k; ++j)
      if (keep[j]) s -= L[k][j] * L[k][j] * d[j];
    keep[k] = s > rtol * G[k][k] && G[k][k] > 0.0;
    if (!keep[k]) { d[k] = 0.0; continue; }
    ++rank;
    d[k] = s;
    for (int i = k + 1; i < P; ++i) {
      double v = G[i][k];
      for (int j = 0; j < k; ++j)
        if (keep[j]) v -= L[i][j] * L[k][j] * d[j];
      L[i][k] = v / s;
    }
  }
  for (int i = 0; i < P; ++i) {
    double a = g[i];
    for (int j = 0; j < i; ++j)
      if (keep[j]) a -= L[i][j] * z[j];
    z[i] = a;
  }
  for (int i = P - 1; i >= 0; --i) {
    if (!keep[i]) { beta[i] = 0.0; continue; }
    double a = z[i] / d[i];
    for (int j = i + 1; j < P; ++j)
      if (keep[j]) a -= L[j][i] * beta[j];
    beta[i] = a;
  }
  return rank;
}

__global__ void __launch_bounds__(256)
global_solve_kernel(const unsigned long long* __restrict__ part, int N, int n_contrib,
                    const double* __restrict__ sqrt_tau, const double* __restrict__ Dinv, double K, double invK,
                    int is_put, GlobalModel* model, DecEntry* table) {
  __shared__ double b_u[kGCols];
  __shared__ int s_valid;
  __shared__ double sred[8 * 32];
  __shared__ double stot[21];
  {  // thread-strided dates, then a fixed-order block reduction: deterministic
    double w[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) w[q] = 0.0;
    for (int t = N - 1 - (int)threadIdx.x; t >= 1; t -= (int)blockDim.x) {
      double m[kGQ];
#pragma unroll
      for (int q = 0; q < kGQ; ++q)
        m[q] = fx_decode(part[(size_t)t * 2 * kGQ + 2 * q], part[(size_t)t * 2 * kGQ + 2 * q + 1], n_contrib);
      const double s = sqrt_tau[t], tau = s * s;
#pragma unroll
      for (int a = 0; a < 7; ++a) w[a] += m[a];                            // A
#pragma unroll
      for (int a = 0; a < 5; ++a) w[7 + a] = fma(s, m[a], w[7 + a]);        // SA
#pragma unroll
      for (int a = 0; a < 3; ++a) w[12 + a] = fma(tau, m[a], w[12 + a]);    // TA
#pragma unroll
      for (int a = 0; a < 4; ++a) w[15 + a] += m[7 + a];                    // Y
#pragma unroll
      for (int a = 0; a < 2; ++a) w[19 + a] = fma(s, m[7 + a], w[19 + a]);  // SY
    }
    block_reduce_sum<32, 8>(w, sred);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < 21; ++q) stot[q] = w[q];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double A[7], SA[5], TA[3], Y[4], SY[2];
    for (int a = 0; a < 7; ++a) A[a] = stot[a];
    for (int a = 0; a < 5; ++a) SA[a] = stot[7 + a];
    for (int a = 0; a < 3; ++a) TA[a] = stot[12 + a];
    for (int a = 0; a < 4; ++a) Y[a] = stot[15 + a];
    for (int a = 0; a < 2; ++a) SY[a] = stot[19 + a];
    double G[kGCols][kGCols], g[kGCols], beta[kGCols];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) G[i][j] = A[i + j];
    for (int i = 0; i < 4; ++i) { G[i][4] = G[4][i] = SA[i]; G[i][5] = G[5][i] = SA[i + 1]; }
    G[4][4] = TA[0]; G[4][5] = G[5][4] = TA[1]; G[5][5] = TA[2];
    for (int i = 0; i < 4; ++i) g[i] = Y[i];
    g[4] = SY[0]; g[5] = SY[1];
    int rank = 0;
    const bool any = A[0] >= 1.0;
    if (any) rank = solve_drop<kGCols>(G, g, beta, 1e-11);
    for (int i = 0; i < kGCols; ++i) b_u[i] = (any && rank > 0) ? beta[i] * K : 0.0;  // y was in units of K
    s_valid = (any && rank > 0) ? 1 : 0;
    // reference-feature coefficients: cont = b0 + b1 u + b2 u^2 + b3 u^3 + b4 s + b5 u s, u = x - 1
    const double b0 = b_u[0], b1 = b_u[1], b2 = b_u[2], b3 = b_u[3], b4 = b_u[4], b5 = b_u[5];
    model->beta_ref[0] = b0 - b1 + b2 - b3;
    model->beta_ref[1] = b1 - 2.0 * b2 + 3.0 * b3;
    model->beta_ref[2] = b2 - 3.0 * b3;
    model->beta_ref[3] = b3;
    model->beta_ref[4] = 0.0;  // hinge: redundant within the ITM rows
    model->beta_ref[5] = b4 - b5;
    model->beta_ref[6] = b5;
    model->n_rows = A[0];
    model->rank = rank;
    model->valid = s_valid;
  }
  __syncthreads();
  // per-date decision polynomial in the raw price S: payoff - continuation, u = S invK - 1
  for (int t = threadIdx.x; t <= N; t += blockDim.x) {
    DecEntry e;
    if (t >= 1 && t <= N - 1 && s_valid) {
      const double s = sqrt_tau[t];
      const double c0 = b_u[0] + b_u[4] * s, c1 = b_u[1] + b_u[5] * s, c2 = b_u[2], c3 = b_u[3];
      const double a = invK;
      double d0 = -(c0 - c1 + c2 - c3), d1 = -(c1 - 2.0 * c2 + 3.0 * c3) * a, d2 = -(c2 - 3.0 * c3) * a * a,
             d3 = -c3 * a * a * a;
      d0 += is_put ? K : -K;
      d1 += is_put ? -1.0 : 1.0;
      e.d[0] = d0; e.d[1] = d1; e.d[2] = d2; e.d[3] = d3;
    } else {  // no model: never exercise
      e.d[0] = -1.0; e.d[1] = 0.0; e.d[2] = 0.0; e.d[3] = 0.0;
    }
    for (int i = 0; i < 4; ++i) { e.f[i] = (float)e.d[i]; e.b[i] = fabsf(e.f[i]) * 4.76837158203125e-7f; }
    e.dinv = Dinv[t];
    e.pad = 0.0;
    table[t] = e;
  }
}

template <typename R> __device__ __forceinline__ bool walk_exercise(const DecEntry& e, R s, bool live);
template <> __device__ __forceinline__ bool walk_exercise<float>(const DecEntry& e, float s, bool live) {
  float p = e.f[3], er = e.b[3];
  const float as = fabsf(s);
#pragma unroll
  for (int i = 2; i >= 0; --i) { p = fmaf(p, s, e.f[i]); er = fmaf(er, as, e.b[i]); }
  bool pos = p > 0.0f;
  if (live && !(fabsf(p) > er)) pos = poly_eval<3>(e.d, (double)s) > 0.0;  // rare: inside the fp32 error bound
  return live && pos;
}
template <> __device__ __forceinline__ bool walk_exercise<double>(const DecEntry& e, double s, bool live) {
  return live && poly_eval<3>(e.d, s) > 0.0;
}

struct WalkArgs {
  const void* S;
  long long ld, M;
  int N, sticky, is_put, tab_in_smem;
  double sgn, kk, c1, c2;        // storage-precision constants (see lsm_resident_kernel.cuh: Store<>)
  const DecEntry* table;         // [N+1]
  double d1_scale;               // D_1 * final_scale
  double* partials;              // [grid][2]
  unsigned int* ticket;
  double* final_out;             // [4]
  unsigned long long* exc;       // [N+1]
  unsigned long long* bnd;       // [N+1] double bits
};

// VEC consecutive paths per thread; cash-flows in date-N money; exercised flag = sign bit (sticky semantics).
template <typename R, int VEC, bool STATS>
__global__ void __launch_bounds__(kGThreads, 4) global_walk_kernel(const WalkArgs a) {
  // dynamic shared memory: [N+1] decision entries (when they fit), then STATS: [N+1] counts, [N+1] boundary keys
  extern __shared__ unsigned long long s_dyn[];
  __shared__ double red[kGWarps * 2];
  __shared__ bool is_last;
  const int N = a.N;
  DecEntry* s_tab = reinterpret_cast<DecEntry*>(s_dyn);
  unsigned long long* s_stats = s_dyn + (a.tab_in_smem ? (size_t)(N + 1) * (sizeof(DecEntry) / 8) : 0);
  const long long j0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const bool act = j0 < a.M;  // M % VEC == 0 (launcher)
  const R sgn = (R)a.sgn, kk = (R)a.kk, c1 = (R)a.c1, c2 = (R)a.c2;
  const bool sticky = a.sticky != 0, is_put = a.is_put != 0;
  if (STATS) {
    for (int i = threadIdx.x; i <= N; i += blockDim.x) { s_stats[i] = 0ull; s_stats[N + 1 + i] = bnd_none(a.is_put); }
  }
  if (a.tab_in_smem) {
    for (int i = threadIdx.x; i < (N + 1) * (int)(sizeof(DecEntry) / 8); i += blockDim.x)
      s_dyn[i] = reinterpret_cast<const unsigned long long*>(a.table)[i];
  }
  __syncthreads();
  const DecEntry* __restrict__ tab = a.tab_in_smem ? s_tab : a.table;
  const R* __restrict__ Sp = static_cast<const R*>(a.S) + j0;
  R cf[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) cf[i] = (R)0;
  if (act) {
    R v[VEC];
    load_vec<R, VEC>(Sp + (size_t)N * a.ld, v);
#pragma unroll
    for (int i = 0; i < VEC; ++i) cf[i] = (sgn * v[i] > kk) ? fma(sgn, v[i], c1) + c2 : (R)0;
  }
  constexpr int CH = 8;  // dates per chunk: the chunk's rows are loaded up front (8 x 16 bytes in flight per thread)
  for (int t0 = N - 1; t0 >= 1; t0 -= CH) {
    R rows[CH][VEC];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int t = t0 - c;
      if (act && t >= 1) {
        load_vec<R, VEC>(Sp + (size_t)t * a.ld, rows[c]);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) rows[c][i] = (R)0;
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int t = t0 - c;
      if (t < 1) break;
      const DecEntry e = tab[t];
      const R dinv = (R)e.dinv;
      unsigned int cnt = 0;
      R em = (R)-INFINITY;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const R v = rows[c][i];
        const R u = sgn * v;
        const bool live = act && !(sticky && cf[i] < (R)0) && (u > kk);
        const bool ex = walk_exercise<R>(e, v, live);
        const R pay = (fma(sgn, v, c1) + c2) * dinv;
        cf[i] = ex ? (sticky ? -pay : pay) : cf[i];
        if (STATS) { cnt += ex ? 1u : 0u; em = fmax(em, ex ? -u : (R)-INFINITY); }
      }
      if (STATS) {
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt) {  // warp-uniform
          const double ext = -(double)sgn * (double)em;
          unsigned long long b = (unsigned long long)__double_as_longlong(ext);
          if (!isfinite(ext)) b = bnd_none(a.is_put);
          b = is_put ? warp_max_u64(b) : warp_min_u64(b);
          if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_stats[t], (unsigned long long)cnt);
            if (is_put) atomicMax(&s_stats[N + 1 + t], b); else atomicMin(&s_stats[N + 1 + t], b);
          }
        }
      }
    }
  }
  double fin[2] = {0.0, 0.0};
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const double c = fabs((double)cf[i]);
    fin[0] += c;
    fin[1] += c * c;
  }
  block_reduce_sum<2, kGWarps>(fin, red);
  if (STATS) {
    __syncthreads();
    for (int t = threadIdx.x; t <= N; t += blockDim.x) {
      if (s_stats[t]) {
        atomicAdd(a.exc + t, s_stats[t]);
        if (is_put) atomicMax(a.bnd + t, s_stats[N + 1 + t]); else atomicMin(a.bnd + t, s_stats[N + 1 + t]);
      }
    }
  }
  if (threadIdx.x == 0) {
    a.partials[(size_t)blockIdx.x * 2] = fin[0];
    a.partials[(size_t)blockIdx.x * 2 + 1] = fin[1];
    __threadfence();
    is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) { s[0] += a.partials[(size_t)b * 2]; s[1] += a.partials[(size_t)b * 2 + 1]; }
    warp_allreduce_sum<2>(s);
    if (threadIdx.x == 0) {
      const double n = (double)a.M;
      const double mean = s[0] / n;
      double var = n > 1.0 ? (s[1] - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      a.final_out[0] = mean * a.d1_scale;
      a.final_out[1] = sqrt(var / n) * a.d1_scale;
      a.final_out[2] = s[0];
      a.final_out[3] = s[1];
      *a.ticket = 0u;
    }
  }
}

template <typename R> static int lsm_global_t(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N,
                                              const optmc_lsm_params* lp, optmc_global_result* out) {
  const bool f32 = sizeof(R) == 4;
  const bool sticky = (lp->semantics & OPTMC_SEM_STICKY_MASK) != 0;
  const double dt = lp->T / N, disc = exp(-lp->r * dt);
  const double final_scale = (lp->semantics & OPTMC_SEM_REF_DISCOUNT) ? 1.0 : disc;
  int rc = ensure_per_date(ctx, N);
  if (rc) return rc;
  // host tables: D_t = disc^(N-t), 1/D_t (running products, as in the persistent sweep), sqrt(tau_t) (om3:109)
  std::vector<double> tab((size_t)3 * (N + 1));
  double* Dt = tab.data(); double* Dinv = Dt + (N + 1); double* st = Dinv + (N + 1);
  {
    const double inv_disc = 1.0 / disc;
    double d = 1.0, di = 1.0;
    for (int t = N; t >= 0; --t) { Dt[t] = d; Dinv[t] = di; d *= disc; di *= inv_disc; }
    for (int t = 0; t <= N; ++t) { const double tau = lp->T - t * dt; st[t] = sqrt(tau > 1e-6 ? tau : 1e-6); }
  }
  // pass constants, exact in the storage type (see lsm_resident_kernel.cuh: Store<>)
  const double sg = lp->is_put ? -1.0 : 1.0;
  const StrikeConsts kc = strike_consts(lp->K, lp->is_put != 0, f32);
  const double Kcmp = kc.Kcmp, Kh = kc.Kh, Kl = kc.Kl;
  const bool vec4 = (M % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)S % 16 == 0);
  const long long units = vec4 ? M / 4 : M;
  const unsigned wg = (unsigned)((units + kGThreads - 1) / kGThreads);
  // device workspace: tables | per-date fixed-point moments | decision table | model | partials
  const size_t off_tab = 0;
  const size_t off_part = off_tab + ((tab.size() * 8 + 255) / 256 * 256);
  const size_t off_dec = off_part + (((size_t)(N + 1) * 2 * kGQ * 8 + 255) / 256 * 256);
  const size_t off_model = off_dec + (((size_t)(N + 1) * sizeof(DecEntry) + 255) / 256 * 256);
  const size_t off_partials = off_model + 256;
  const size_t total = off_partials + (size_t)wg * 2 * sizeof(double);
  rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, total);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  double* d_tab = reinterpret_cast<double*>(dev + off_tab);
  unsigned long long* d_part = reinterpret_cast<unsigned long long*>(dev + off_part);
  DecEntry* d_dec = reinterpret_cast<DecEntry*>(dev + off_dec);
  GlobalModel* d_model = reinterpret_cast<GlobalModel*>(dev + off_model);
  double* d_partials = reinterpret_cast<double*>(dev + off_partials);
  OPTMC_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemsetAsync(d_part, 0, (size_t)(N + 1) * 2 * kGQ * 8, ctx->stream));
  ctx->sw = SweepDesc{};
  ctx->sw.N = N; ctx->sw.lp = *lp;
  rc = sweep_reset_stats(ctx);  // per-date statistics + flags
  if (rc) return rc;
  int n_launches = 1;

  const R* Sr = static_cast<const R*>(S);
  int gx = 1;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (N >= 2) {
    long long g = (units + kGThreads * 8 - 1) / (kGThreads * 8);  // ~8 vector rows per thread: amortises the block reduce
    if (g < 1) g = 1;
    if (g > 4096) g = 4096;
    gx = (int)g;
    if (vec4)
      global_gram_kernel<R, 4><<<dim3(gx, N - 1), kGThreads, 0, ctx->stream>>>(Sr, ld, M, N, lp->K, 1.0 / lp->K, lp->is_put,
                                                                              sg, sg * Kcmp, d_tab, d_part, ctx->d_flags);
    else
      global_gram_kernel<R, 1><<<dim3(gx, N - 1), kGThreads, 0, ctx->stream>>>(Sr, ld, M, N, lp->K, 1.0 / lp->K, lp->is_put,
                                                                              sg, sg * Kcmp, d_tab, d_part, ctx->d_flags);
    ctx->launches++; ++n_launches;
    OPTMC_CUDA(cudaGetLastError());
  }
  global_solve_kernel<<<1, 256, 0, ctx->stream>>>(d_part, N, gx, d_tab + 2 * (N + 1), d_tab + (N + 1), lp->K, 1.0 / lp->K,
                                                  lp->is_put, d_model, d_dec);
  ctx->launches++; ++n_launches;
  OPTMC_CUDA(cudaGetLastError());
  OPTMC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));

  WalkArgs a{};
  a.S = S; a.ld = ld; a.M = M; a.N = N; a.sticky = sticky ? 1 : 0; a.is_put = lp->is_put;
  a.sgn = sg; a.kk = sg * Kcmp; a.c1 = -sg * Kh; a.c2 = -sg * Kl;
  a.table = d_dec;
  const size_t tab_bytes = (size_t)(N + 1) * sizeof(DecEntry);
  a.tab_in_smem = tab_bytes <= 40 * 1024 ? 1 : 0;
  a.d1_scale = Dt[N >= 1 ? 1 : 0] * final_scale;
  a.partials = d_partials; a.ticket = ctx->tickets; a.final_out = ctx->d_final;
  a.exc = ctx->d_exc; a.bnd = ctx->d_bnd;
  const bool stats = out->ex_count != nullptr || out->boundary != nullptr;
  const size_t smem = (a.tab_in_smem ? tab_bytes : 0) + (stats ? (size_t)2 * (N + 1) * sizeof(unsigned long long) : 0);
  if (smem > 46 * 1024) { set_error("global LSM: too many exercise dates for the per-date statistics"); return OPTMC_EUNSUPPORTED; }
  if (vec4) {
    if (stats) global_walk_kernel<R, 4, true><<<wg, kGThreads, smem, ctx->stream>>>(a);
    else global_walk_kernel<R, 4, false><<<wg, kGThreads, smem, ctx->stream>>>(a);
  } else {
    if (stats) global_walk_kernel<R, 1, true><<<wg, kGThreads, smem, ctx->stream>>>(a);
    else global_walk_kernel<R, 1, false><<<wg, kGThreads, smem, ctx->stream>>>(a);
  }
  ctx->launches++; ++n_launches;
  OPTMC_CUDA(cudaGetLastError());
  OPTMC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));

  // results
  double fin[4];
  GlobalModel hm;
  int flags[4];
  std::vector<unsigned long long> hexc, hbnd;
  OPTMC_CUDA(cudaMemcpyAsync(fin, ctx->d_final, sizeof(fin), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(&hm, d_model, sizeof(hm), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  if (out->ex_count) { hexc.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hexc.data(), ctx->d_exc, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->boundary) { hbnd.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hbnd.data(), ctx->d_bnd, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  {  // optmc_ctx_kernel_times: (pass 1 + solve, pass 2)
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last_paths_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) ctx->last_sweep_ms = ms;
  }
  if (flags[0]) { set_error("global LSM: a moment left the fixed-point range or is not finite"); return OPTMC_EUNSUPPORTED; }
  out->price = fin[0]; out->stderr_ = fin[1]; out->n_paths = M; out->n_rows = (int64_t)(hm.n_rows + 0.5);
  out->n_launches = n_launches; out->rank = hm.rank;
  for (int i = 0; i < 7; ++i) out->beta[i] = hm.valid ? hm.beta_ref[i] : nan("");
  if (out->ex_count) for (int t = 0; t <= N; ++t) out->ex_count[t] = (int64_t)hexc[t];
  if (out->boundary) {
    const unsigned long long none = lp->is_put ? 0ull : ~0ull;
    for (int t = 0; t <= N; ++t) {
      if (hbnd[t] == none) out->boundary[t] = nan("");
      else memcpy(&out->boundary[t], &hbnd[t], 8);
    }
  }
  return OPTMC_OK;
}

// ---- out-of-sample exercise: apply per-date regression coefficients fitted on ONE set of paths to
// ANOTHER slab.  The policy "exercise at date t iff payoff(S) > sum_i beta[t][i] (S/K)^i" is tabulated as the decision
// polynomial of global_walk_kernel, which then prices it in one streaming pass (no regression on these paths, hence
// no in-sample look-ahead: the estimate is biased low, the usual companion of the in-sample LSM value).
template <typename R>
static int lsm_apply_policy_t(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, const optmc_lsm_params* lp,
                              const double* betas, optmc_lsm_result* out) {
  const bool f32 = sizeof(R) == 4;
  const bool sticky = (lp->semantics & OPTMC_SEM_STICKY_MASK) != 0;
  const int p = lp->basis == OPTMC_BASIS_POLY3 ? 4 : 3;
  const double dt = lp->T / N, disc = exp(-lp->r * dt), inv_disc = 1.0 / disc;
  const double final_scale = (lp->semantics & OPTMC_SEM_REF_DISCOUNT) ? 1.0 : disc;
  int rc = ensure_per_date(ctx, N);
  if (rc) return rc;
  std::vector<DecEntry> tab(N + 1);
  double d1 = 1.0;
  {
    double d = 1.0, di = 1.0;
    for (int t = N; t >= 0; --t) {
      DecEntry e{};
      const double* b = betas + (size_t)t * p;
      bool ok = t >= 1 && t <= N - 1;
      for (int i = 0; i < p && ok; ++i) ok = b[i] == b[i];  // NaN row: the regression of that date was skipped
      if (ok) {
        double sc = 1.0;
        for (int i = 0; i < 4; ++i) {
          double c = i < p ? -b[i] * sc : 0.0;
          if (i == 0) c += lp->is_put ? lp->K : -lp->K;
          if (i == 1) c += lp->is_put ? -1.0 : 1.0;
          e.d[i] = c;
          sc /= lp->K;
        }
      } else {
        e.d[0] = -1.0;
      }
      for (int i = 0; i < 4; ++i) { e.f[i] = (float)e.d[i]; e.b[i] = fabsf(e.f[i]) * 4.76837158203125e-7f; }
      e.dinv = di;
      tab[t] = e;
      if (t == 1) d1 = d;
      d *= disc; di *= inv_disc;
    }
  }
  const double sg = lp->is_put ? -1.0 : 1.0;
  const StrikeConsts kc = strike_consts(lp->K, lp->is_put != 0, f32);
  const double Kcmp = kc.Kcmp, Kh = kc.Kh, Kl = kc.Kl;
  const bool vec4 = (M % 4 == 0) && (ld % 4 == 0) && ((uintptr_t)S % 16 == 0);
  const long long units = vec4 ? M / 4 : M;
  const unsigned wg = (unsigned)((units + kGThreads - 1) / kGThreads);
  const size_t off_part = ((size_t)(N + 1) * sizeof(DecEntry) + 255) / 256 * 256;
  rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off_part + (size_t)wg * 2 * sizeof(double));
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  OPTMC_CUDA(cudaMemcpyAsync(dev, tab.data(), tab.size() * sizeof(DecEntry), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sw = SweepDesc{};
  ctx->sw.N = N; ctx->sw.lp = *lp;
  rc = sweep_reset_stats(ctx);
  if (rc) return rc;
  WalkArgs a{};
  a.S = S; a.ld = ld; a.M = M; a.N = N; a.sticky = sticky ? 1 : 0; a.is_put = lp->is_put;
  a.sgn = sg; a.kk = sg * Kcmp; a.c1 = -sg * Kh; a.c2 = -sg * Kl;
  a.table = reinterpret_cast<const DecEntry*>(dev);
  const size_t tab_bytes = (size_t)(N + 1) * sizeof(DecEntry);
  a.tab_in_smem = tab_bytes <= 40 * 1024 ? 1 : 0;
  a.d1_scale = d1 * final_scale;
  a.partials = reinterpret_cast<double*>(dev + off_part); a.ticket = ctx->tickets; a.final_out = ctx->d_final;
  a.exc = ctx->d_exc; a.bnd = ctx->d_bnd;
  const bool stats = out->ex_count != nullptr || out->boundary != nullptr;
  const size_t smem = (a.tab_in_smem ? tab_bytes : 0) + (stats ? (size_t)2 * (N + 1) * sizeof(unsigned long long) : 0);
  if (smem > 46 * 1024) { set_error("policy application: too many exercise dates for the per-date statistics"); return OPTMC_EUNSUPPORTED; }
  if (vec4) {
    if (stats) global_walk_kernel<R, 4, true><<<wg, kGThreads, smem, ctx->stream>>>(a);
    else global_walk_kernel<R, 4, false><<<wg, kGThreads, smem, ctx->stream>>>(a);
  } else {
    if (stats) global_walk_kernel<R, 1, true><<<wg, kGThreads, smem, ctx->stream>>>(a);
    else global_walk_kernel<R, 1, false><<<wg, kGThreads, smem, ctx->stream>>>(a);
  }
  ctx->launches += 2;
  OPTMC_CUDA(cudaGetLastError());
  double fin[4];
  std::vector<unsigned long long> hexc, hbnd;
  OPTMC_CUDA(cudaMemcpyAsync(fin, ctx->d_final, sizeof(fin), cudaMemcpyDeviceToHost, ctx->stream));
  if (out->ex_count) { hexc.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hexc.data(), ctx->d_exc, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->boundary) { hbnd.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hbnd.data(), ctx->d_bnd, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  out->price = fin[0]; out->stderr_ = fin[1]; out->n_paths = M; out->impl_used = OPTMC_SWEEP_SPLIT; out->n_launches = 2;
  if (out->betas) memcpy(out->betas, betas, (size_t)(N + 1) * p * sizeof(double));
  if (out->n_itm) for (int t = 0; t <= N; ++t) out->n_itm[t] = 0;
  if (out->ex_count) for (int t = 0; t <= N; ++t) out->ex_count[t] = (int64_t)hexc[t];
  if (out->boundary) {
    const unsigned long long none = lp->is_put ? 0ull : ~0ull;
    for (int t = 0; t <= N; ++t) {
      if (hbnd[t] == none) out->boundary[t] = nan("");
      else memcpy(&out->boundary[t], &hbnd[t], 8);
    }
  }
  return OPTMC_OK;
}

int lsm_apply_policy(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype, const optmc_lsm_params* lp,
                     const double* betas, optmc_lsm_result* out) {
  if (!S || !lp || !betas || !out) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(lp->K > 0) || !(lp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (lp->r < 0) { set_error("r must be non-negative."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  if (lp->basis != OPTMC_BASIS_POLY2 && lp->basis != OPTMC_BASIS_POLY3) { set_error("basis must be POLY2 or POLY3"); return OPTMC_EINVAL; }
  if (dtype == OPTMC_F64) return lsm_apply_policy_t<double>(ctx, S, ld, M, N, lp, betas, out);
  return lsm_apply_policy_t<float>(ctx, S, ld, M, N, lp, betas, out);
}

int lsm_global(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype,
               const optmc_lsm_params* lp, optmc_global_result* out) {
  if (!S || !lp || !out) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(lp->K > 0) || !(lp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (lp->r < 0) { set_error("r must be non-negative."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  if (dtype == OPTMC_F64) return lsm_global_t<double>(ctx, S, ld, M, N, lp, out);
  return lsm_global_t<float>(ctx, S, ld, M, N, lp, out);
}

}  // namespace optmc
