// lsm_mlp.cu -- per-date neural-network LSM: the reference's v1/v2 loop (om2:277-310 = om15:145-186 =
// om1:107-150).  At every exercise date a FRESH ContNet (1 -> H -> H -> 1, ReLU; om2:114-126) is trained for
// `epochs` full-batch Adam steps (lr, betas 0.9/0.999, eps 1e-8; om2:293-300) on the standardised prices of the
// in-the-money, not-yet-exercised paths against their discounted cash-flows, and its in-sample prediction is
// the continuation value (om2:301-306).  Loop semantics (discount order, sticky mask, strict '>') as in lsm.cu.
//
// Per date (host-launched, like the split sweep; cash-flows live in HBM in date-N money):
//   mlp_count_kernel    live-row count per block + sum(S), sum(S^2) (fp64, fixed-point atomics)
//   mlp_scan_kernel     exclusive scan of the block counts (deterministic compaction offsets), mean / std,
//                       fresh parameters (Philox uniform in torch's default Linear range), Adam state = 0
//   mlp_compact_kernel  dense arrays of the live rows: standardised x (fp32, om2:289-291), target y, path index
//   epochs x { mlp_grad_kernel  forward + backward over 256-row tiles; the four H x 256 activation tiles stay
//                               in shared memory, parameter gradients are tile GEMMs (dW2 = dZ2^T H1, ...)
//                               accumulated per CTA and added to fixed-point accumulators (order-independent);
//              mlp_adam_kernel  one CTA: decode the gradient sums, Adam step, clear the accumulators }
//   mlp_decide_kernel   continuation = net(x); exercise iff payoff > continuation; scatter to the cash-flows
// The network arithmetic is fp32, as in the reference (`.float()`, om2:296-297).
#include <math.h>
#include <string.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

constexpr int kMH = 32;                                   // hidden width (om2 default nn_hidden = 32)
constexpr int kMP = 3 * kMH + kMH * kMH + kMH + 1;         // parameters: w1[H] b1[H] W2[H][H] b2[H] w3[H] b3 = 1153
constexpr int kMThreads = 256;
constexpr int kMWarps = kMThreads / 32;
constexpr int kMTile = 256;                                // rows per tile = threads
constexpr int kMPad = kMH + 1;                             // padded row length of the shared-memory tiles
// parameter offsets
constexpr int oW1 = 0, oB1 = kMH, oW2 = 2 * kMH, oB2 = 2 * kMH + kMH * kMH, oW3 = oB2 + kMH, oB3 = oW3 + kMH;

struct MlpState {          // device scalars of the current date
  long long n_live;        // rows
  double mean, inv_std;    // standardisation (population std; inv_std = 1 when std == 0, om2:289)
  double loss;             // last epoch's mean squared error (diagnostic)
  int step;                // Adam step count
};

template <typename R>
__device__ __forceinline__ bool mlp_live(R s, R c, double K, int is_put, int sticky) {
  const bool ex = sticky && signbit(c);
  return !ex && payoff<double>((double)s, K, is_put != 0) > 0.0;
}

// ---- 1. count + moments --------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kMThreads)
mlp_count_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, double K, int is_put, int sticky,
                 unsigned int* block_count, unsigned long long* mom_fx) {
  __shared__ double red[kMWarps * 2];
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0u;
  __syncthreads();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double acc[2] = {0.0, 0.0};
  bool live = false;
  if (j < M) {
    const R s = S_t[j];
    live = mlp_live<R>(s, cf[j], K, is_put, sticky);
    if (live) { acc[0] = (double)s; acc[1] = (double)s * (double)s; }
  }
  const unsigned int w = __popc(__ballot_sync(0xffffffffu, live));
  if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cnt, w);
  block_reduce_sum<2, kMWarps>(acc, red);
  __syncthreads();
  if (threadIdx.x == 0) block_count[blockIdx.x] = s_cnt;
  if (threadIdx.x < 4 && s_cnt) {  // two quantities x (hi, lo) fixed-point chunks; every contributing block adds a bias
    unsigned long long hi, lo;
    const double v = (threadIdx.x >> 1) ? acc[1] : acc[0];  // warp 0 holds the block totals in every lane
    fx_encode(v * 9.5367431640625e-7, hi, lo);  // scaled by 2^-20: sum(S^2) of 4M paths stays inside the 2^43 range
    atomicAdd(mom_fx + threadIdx.x, (threadIdx.x & 1) ? lo : hi);
  }
  if (threadIdx.x == 0 && s_cnt) atomicAdd(mom_fx + 4, 1ull);  // contributing blocks (bias count)
}

// ---- 2. scan + standardisation + fresh network ------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mlp_scan_kernel(unsigned int* block_count, int nblocks, unsigned long long* mom_fx, MlpState* st, float* params,
                float* adam_m, float* adam_v, unsigned long long* grad_fx, unsigned long long seed, int date) {
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x;
  // exclusive scan of block_count (in place), chunked by 1024
  if (tid == 0) s_base = 0ull;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + tid;
    const unsigned long long v = b < nblocks ? block_count[b] : 0u;
    s_part[tid] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      const unsigned long long add = tid >= off ? s_part[tid - off] : 0ull;
      __syncthreads();
      s_part[tid] += add;
      __syncthreads();
    }
    if (b < nblocks) block_count[b] = (unsigned int)(s_base + s_part[tid] - v);
    __syncthreads();
    if (tid == 0) s_base += s_part[1023];
    __syncthreads();
  }
  if (tid == 0) {
    const long long n = (long long)s_base;
    const int nb = (int)mom_fx[4];
    const double s1 = fx_decode(mom_fx[0], mom_fx[1], nb) * 1048576.0;
    const double s2 = fx_decode(mom_fx[2], mom_fx[3], nb) * 1048576.0;
    double mean = 0.0, inv_std = 1.0;
    if (n > 0) {
      mean = s1 / (double)n;
      double var = s2 / (double)n - mean * mean;
      if (var < 0.0) var = 0.0;
      const double sd = sqrt(var);
      inv_std = sd > 0.0 ? 1.0 / sd : 1.0;
    }
    st->n_live = n; st->mean = mean; st->inv_std = inv_std; st->loss = 0.0; st->step = 0;
    for (int i = 0; i < 5; ++i) mom_fx[i] = 0ull;
  }
  // fresh parameters: torch.nn.Linear default init = U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases
  for (int i = tid; i < kMP; i += blockDim.x) {
    const Philox4 p = philox_for((unsigned long long)i, (unsigned int)date, 0x4D4C50u, seed);
    const float u = ((float)(p.v[0] >> 8) + 0.5f) * (2.0f / 16777216.0f) - 1.0f;  // (-1, 1)
    const float bound = (i < oW2) ? 1.0f : 0.17677669529663687f;                   // fan_in 1 | fan_in H = 32
    params[i] = u * bound;
    adam_m[i] = 0.f; adam_v[i] = 0.f;
    grad_fx[2 * i] = 0ull; grad_fx[2 * i + 1] = 0ull;
  }
  if (tid == 0) { grad_fx[2 * kMP] = 0ull; grad_fx[2 * kMP + 1] = 0ull; grad_fx[2 * kMP + 2] = 0ull; }
}

// ---- 3. compaction ----------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kMThreads)
mlp_compact_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, double K, int is_put, int sticky,
                   double dg, const unsigned int* __restrict__ block_off, const MlpState* __restrict__ st,
                   float* xs, float* ys, unsigned int* idx) {
  __shared__ unsigned int s_woff[kMWarps];
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool live = false;
  R s = (R)0, c = (R)0;
  if (j < M) { s = S_t[j]; c = cf[j]; live = mlp_live<R>(s, c, K, is_put, sticky); }
  const unsigned int mask = __ballot_sync(0xffffffffu, live);
  if (lane == 0) s_woff[warp] = __popc(mask);
  __syncthreads();
  unsigned int woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_woff[w];
  if (live) {
    const unsigned int pos = block_off[blockIdx.x] + woff + __popc(mask & ((1u << lane) - 1u));
    xs[pos] = (float)(((double)s - st->mean) * st->inv_std);   // om2:289-291, then .float() (om2:296)
    ys[pos] = (float)((double)fabs(c) * dg);                    // cash-flow at date t (date-N money x D_t), .float()
    idx[pos] = (unsigned int)j;
  }
}

// ---- 4. forward + backward --------------------------------------------------------------------------------
struct MlpSmem {
  float W2[kMH * kMPad];
  float w1[kMH], b1[kMH], b2[kMH], w3[kMH];
  float b3;
  float H1[kMTile * kMPad], H2[kMTile * kMPad], DZ2[kMTile * kMPad], DH1[kMTile * kMPad];
  float xs[kMTile], dout[kMTile];
};

__device__ __forceinline__ void mlp_load_params(MlpSmem& sm, const float* __restrict__ params) {
  for (int i = threadIdx.x; i < kMH * kMH; i += blockDim.x) sm.W2[(i / kMH) * kMPad + (i % kMH)] = params[oW2 + i];
  for (int i = threadIdx.x; i < kMH; i += blockDim.x) {
    sm.w1[i] = params[oW1 + i]; sm.b1[i] = params[oB1 + i]; sm.b2[i] = params[oB2 + i]; sm.w3[i] = params[oW3 + i];
  }
  if (threadIdx.x == 0) sm.b3 = params[oB3];
}

// forward of one row; h1 / h2 are written to the caller's arrays
__device__ __forceinline__ float mlp_forward_row(const MlpSmem& sm, float x, float (&h1)[kMH], float (&h2)[kMH]) {
#pragma unroll
  for (int i = 0; i < kMH; ++i) h1[i] = fmaxf(fmaf(sm.w1[i], x, sm.b1[i]), 0.f);
  float out = sm.b3;
#pragma unroll
  for (int j = 0; j < kMH; ++j) {
    float z = sm.b2[j];
#pragma unroll
    for (int i = 0; i < kMH; ++i) z = fmaf(sm.W2[j * kMPad + i], h1[i], z);
    h2[j] = fmaxf(z, 0.f);
    out = fmaf(sm.w3[j], h2[j], out);
  }
  return out;
}

__global__ void __launch_bounds__(kMThreads, 1)
mlp_grad_kernel(const float* __restrict__ params, const float* __restrict__ xs, const float* __restrict__ ys,
                const MlpState* __restrict__ st, unsigned long long* grad_fx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
  const long long n = st->n_live;
  if ((long long)blockIdx.x * kMTile >= n) return;
  mlp_load_params(sm, params);
  __syncthreads();
  const int tid = threadIdx.x;
  const float inv_n2 = (float)(2.0 / (double)n);  // d(mean squared error)/d(out) = 2 (out - y) / n
  // per-thread gradient accumulators: dW2 elements (j, i0..i0+3) with j = tid / 8, i0 = 4 * (tid % 8);
  // threads 0..31 also own the vector gradients of unit i = tid
  float gW2[4] = {0.f, 0.f, 0.f, 0.f};
  float gw1 = 0.f, gb1 = 0.f, gb2 = 0.f, gw3 = 0.f, gb3 = 0.f, loss = 0.f;
  const int gj = tid >> 3, gi0 = (tid & 7) * 4;
  for (long long r0 = (long long)blockIdx.x * kMTile; r0 < n; r0 += (long long)gridDim.x * kMTile) {
    const long long r = r0 + tid;
    const bool act = r < n;
    {  // forward + output-side backward of this thread's row
      float h1[kMH], h2[kMH];
      const float x = act ? xs[r] : 0.f;
      const float out = mlp_forward_row(sm, x, h1, h2);
      const float err = act ? out - ys[r] : 0.f;
      const float dout = err * inv_n2;
      loss = fmaf(err, err, loss);
      sm.xs[tid] = x; sm.dout[tid] = dout;
#pragma unroll
      for (int i = 0; i < kMH; ++i) { sm.H1[tid * kMPad + i] = h1[i]; sm.H2[tid * kMPad + i] = h2[i]; }
      float dz2[kMH];
#pragma unroll
      for (int j = 0; j < kMH; ++j) {
        dz2[j] = h2[j] > 0.f ? dout * sm.w3[j] : 0.f;
        sm.DZ2[tid * kMPad + j] = dz2[j];
      }
#pragma unroll
      for (int i = 0; i < kMH; ++i) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < kMH; ++j) d = fmaf(sm.W2[j * kMPad + i], dz2[j], d);
        sm.DH1[tid * kMPad + i] = h1[i] > 0.f ? d : 0.f;
      }
    }
    __syncthreads();
    // parameter gradients of the tile: small GEMMs over the 256 rows held in shared memory
    for (int rr = 0; rr < kMTile; ++rr) {
      const float dz = sm.DZ2[rr * kMPad + gj];
#pragma unroll
      for (int k = 0; k < 4; ++k) gW2[k] = fmaf(dz, sm.H1[rr * kMPad + gi0 + k], gW2[k]);
    }
    // vector gradients: warp 1 -> (dw1, db1), warp 2 -> db2, warp 3 -> (dw3, db3); unit = lane
    if (tid >= 32 && tid < 64) {
      const int u = tid - 32;
      for (int rr = 0; rr < kMTile; ++rr) {
        const float dh = sm.DH1[rr * kMPad + u];
        gw1 = fmaf(dh, sm.xs[rr], gw1);
        gb1 += dh;
      }
    } else if (tid >= 64 && tid < 96) {
      const int u = tid - 64;
      for (int rr = 0; rr < kMTile; ++rr) gb2 += sm.DZ2[rr * kMPad + u];
    } else if (tid >= 96 && tid < 128) {
      const int u = tid - 96;
      for (int rr = 0; rr < kMTile; ++rr) {
        const float d = sm.dout[rr];
        gw3 = fmaf(d, sm.H2[rr * kMPad + u], gw3);
        if (u == 0) gb3 += d;
      }
    }
    __syncthreads();
  }
  // CTA totals -> fixed-point accumulators (order-independent across CTAs)
  auto add_fx = [&](int p, float v) {
    unsigned long long hi, lo;
    fx_encode((double)v, hi, lo);
    atomicAdd(grad_fx + 2 * p, hi);
    atomicAdd(grad_fx + 2 * p + 1, lo);
  };
#pragma unroll
  for (int k = 0; k < 4; ++k) add_fx(oW2 + gj * kMH + gi0 + k, gW2[k]);
  if (tid >= 32 && tid < 64) { add_fx(oW1 + tid - 32, gw1); add_fx(oB1 + tid - 32, gb1); }
  else if (tid >= 64 && tid < 96) add_fx(oB2 + tid - 64, gb2);
  else if (tid >= 96 && tid < 128) { add_fx(oW3 + tid - 96, gw3); if (tid == 96) add_fx(oB3, gb3); }
  {  // loss (diagnostic) and the number of contributing CTAs (fixed-point bias count)
    __shared__ double lred[kMWarps];
    double l = (double)loss;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) l += shfl_xor_f64(l, m);
    if ((tid & 31) == 0) lred[tid >> 5] = l;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kMWarps; ++w) t += lred[w];
      unsigned long long hi, lo;
      fx_encode(t * 9.5367431640625e-7, hi, lo);
      atomicAdd(grad_fx + 2 * kMP, hi);
      atomicAdd(grad_fx + 2 * kMP + 1, lo);
      atomicAdd(grad_fx + 2 * kMP + 2, 1ull);
    }
  }
}

// ---- 5. Adam ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mlp_adam_kernel(float* params, float* adam_m, float* adam_v, unsigned long long* grad_fx, MlpState* st, float lr) {
  __shared__ int s_nb, s_step;
  if (threadIdx.x == 0) { s_nb = (int)grad_fx[2 * kMP + 2]; s_step = st->step + 1; }
  __syncthreads();
  const int nb = s_nb, step = s_step;
  if (nb == 0) return;  // no live rows at this date
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  for (int i = threadIdx.x; i < kMP; i += blockDim.x) {
    const float g = (float)fx_decode(grad_fx[2 * i], grad_fx[2 * i + 1], nb);
    grad_fx[2 * i] = 0ull; grad_fx[2 * i + 1] = 0ull;
    const float m = b1 * adam_m[i] + (1.0f - b1) * g;
    const float v = b2 * adam_v[i] + (1.0f - b2) * g * g;
    adam_m[i] = m; adam_v[i] = v;
    // torch.optim.Adam: param -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
    const float denom = sqrtf(v) / sqrtf(bc2) + eps;
    params[i] -= (lr / bc1) * (m / denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->loss = fx_decode(grad_fx[2 * kMP], grad_fx[2 * kMP + 1], nb) * 1048576.0 / (double)st->n_live;
    st->step = step;
    grad_fx[2 * kMP] = 0ull; grad_fx[2 * kMP + 1] = 0ull; grad_fx[2 * kMP + 2] = 0ull;
  }
}

// ---- 6. decision ------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kMThreads, 1)
mlp_decide_kernel(const float* __restrict__ params, const float* __restrict__ xs, const unsigned int* __restrict__ idx,
                  const MlpState* __restrict__ st, const R* __restrict__ S_t, R* cf, double K, double Kh, double Kl,
                  int is_put, int sticky, R dinv, float* cont_out, unsigned long long* exc_t,
                  unsigned long long* bnd_t) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
  const long long n = st->n_live;
  if ((long long)blockIdx.x * kMTile >= n) return;
  mlp_load_params(sm, params);
  __syncthreads();
  const R sgn = is_put ? (R)-1 : (R)1;
  const R c1 = (R)(is_put ? Kh : -Kh), c2 = (R)(is_put ? Kl : -Kl);
  unsigned int cnt = 0;
  unsigned long long bnd = bnd_none(is_put);
  for (long long r = (long long)blockIdx.x * kMTile + threadIdx.x; r < n; r += (long long)gridDim.x * kMTile) {
    float h1[kMH], h2[kMH];
    const float cont = mlp_forward_row(sm, xs[r], h1, h2);
    if (cont_out) cont_out[r] = cont;
    const unsigned int j = idx[r];
    const R sr = S_t[j];
    const double s = (double)sr;
    const double pay = payoff<double>(s, K, is_put != 0);
    if (pay > (double)cont) {  // strict '>' (om2:304); float32 continuation promoted as in numpy
      const R a = (fma(sgn, sr, c1) + c2) * dinv;  // payoff in date-N money
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

__global__ void mlp_nitm_kernel(const MlpState* st, long long* nitm_t, double* loss_t) {
  *nitm_t = st->n_live;
  if (loss_t) *loss_t = st->loss;
}

// ---- host driver --------------------------------------------------------------------------------------------
template <typename R> static int lsm_mlp_t(optmc_ctx* ctx, const optmc_mlp_params* np_, optmc_lsm_result* out) {
  SweepDesc& sw = ctx->sw;
  const long long M = sw.M;
  const int N = sw.N;
  const int nblocks = (int)((M + kMThreads - 1) / kMThreads);
  const bool sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) != 0;
  // workspace: block counts | xs | ys | idx | params | adam m, v | grad accumulators | moments | state | cont
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_cnt = take((size_t)nblocks * 4), o_xs = take((size_t)M * 4), o_ys = take((size_t)M * 4),
               o_idx = take((size_t)M * 4), o_par = take(kMP * 4), o_m = take(kMP * 4), o_v = take(kMP * 4),
               o_g = take((2 * kMP + 3) * 8), o_mom = take(5 * 8), o_st = take(sizeof(MlpState)),
               o_loss = take((size_t)(N + 1) * 8);
  int rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  unsigned int* d_cnt = reinterpret_cast<unsigned int*>(dev + o_cnt);
  float* d_xs = reinterpret_cast<float*>(dev + o_xs);
  float* d_ys = reinterpret_cast<float*>(dev + o_ys);
  unsigned int* d_idx = reinterpret_cast<unsigned int*>(dev + o_idx);
  float* d_par = reinterpret_cast<float*>(dev + o_par);
  float* d_m = reinterpret_cast<float*>(dev + o_m);
  float* d_v = reinterpret_cast<float*>(dev + o_v);
  unsigned long long* d_g = reinterpret_cast<unsigned long long*>(dev + o_g);
  unsigned long long* d_mom = reinterpret_cast<unsigned long long*>(dev + o_mom);
  MlpState* d_st = reinterpret_cast<MlpState*>(dev + o_st);
  double* d_loss = reinterpret_cast<double*>(dev + o_loss);
  OPTMC_CUDA(cudaMemsetAsync(d_mom, 0, 5 * 8, ctx->stream));
  OPTMC_CUDA(cudaMemsetAsync(d_loss, 0, (size_t)(N + 1) * 8, ctx->stream));

  rc = sweep_begin(ctx);  // cf = payoff(S[N]) (date-N money), per-date statistics reset
  if (rc) return rc;
  const R* Sr = static_cast<const R*>(sw.S);
  R* cf = static_cast<R*>(ctx->cf);
  const size_t smem = sizeof(MlpSmem);
  OPTMC_CUDA(cudaFuncSetAttribute(mlp_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OPTMC_CUDA(cudaFuncSetAttribute(mlp_decide_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int tiles = (int)((M + kMTile - 1) / kMTile);
  const int grid_rows = tiles < ctx->sm_count ? tiles : ctx->sm_count;  // one CTA per SM (135 KB of tiles each)
  for (int t = N - 1; t >= 1; --t) {
    const R* S_t = Sr + (size_t)t * sw.ld;
    mlp_count_kernel<R><<<nblocks, kMThreads, 0, ctx->stream>>>(S_t, cf, M, sw.lp.K, sw.lp.is_put, sticky, d_cnt, d_mom);
    mlp_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, nblocks, d_mom, d_st, d_par, d_m, d_v, d_g, np_->seed, t);
    mlp_compact_kernel<R><<<nblocks, kMThreads, 0, ctx->stream>>>(S_t, cf, M, sw.lp.K, sw.lp.is_put, sticky, sw.Dt[t], d_cnt,
                                                                  d_st, d_xs, d_ys, d_idx);
    for (int e = 0; e < np_->epochs; ++e) {
      mlp_grad_kernel<<<grid_rows, kMThreads, smem, ctx->stream>>>(d_par, d_xs, d_ys, d_st, d_g);
      mlp_adam_kernel<<<1, 256, 0, ctx->stream>>>(d_par, d_m, d_v, d_g, d_st, (float)np_->lr);
    }
    mlp_decide_kernel<R><<<grid_rows, kMThreads, smem, ctx->stream>>>(d_par, d_xs, d_idx, d_st, S_t, cf, sw.lp.K, sw.Kh, sw.Kl,
                                                                      sw.lp.is_put, sticky, (R)sw.Dinv[t], nullptr,
                                                                      ctx->d_exc + t, ctx->d_bnd + t);
    mlp_nitm_kernel<<<1, 1, 0, ctx->stream>>>(d_st, ctx->d_nitm + t, d_loss + t);
    ctx->launches += 5 + 2 * np_->epochs; sw.n_launches += 5 + 2 * np_->epochs;
  }
  OPTMC_CUDA(cudaGetLastError());
  rc = sweep_finish(ctx, ctx->gram);
  if (rc) return rc;
  rc = sweep_finalize_price(ctx, ctx->gram);
  if (rc) return rc;
  sw.impl_used = OPTMC_SWEEP_SPLIT;
  sw.have_results = true;
  (void)out;
  return OPTMC_OK;
}

int lsm_mlp(optmc_ctx* ctx, const optmc_mlp_params* np_, optmc_lsm_result* out) {
  if (!np_) { set_error("null argument"); return OPTMC_EINVAL; }
  if (np_->hidden != kMH) { set_error("per-date MLP: hidden width must be 32 (om2 default nn_hidden)"); return OPTMC_EUNSUPPORTED; }
  if (np_->epochs < 0 || np_->epochs > 10000 || !(np_->lr > 0)) { set_error("per-date MLP: bad epochs / lr"); return OPTMC_EINVAL; }
  if (ctx->sw.M >= (1ll << 32)) { set_error("per-date MLP: too many paths"); return OPTMC_EUNSUPPORTED; }
  return ctx->sw.dtype == OPTMC_F64 ? lsm_mlp_t<double>(ctx, np_, out) : lsm_mlp_t<float>(ctx, np_, out);
}

// The fresh parameters of date `date` (test aid: lets the oracle start from the same network).
int mlp_init_params_host(unsigned long long seed, int date, float* out) {
  for (int i = 0; i < kMP; ++i) {
    const Philox4 p = philox_for((unsigned long long)i, (unsigned int)date, 0x4D4C50u, seed);
    const float u = ((float)(p.v[0] >> 8) + 0.5f) * (2.0f / 16777216.0f) - 1.0f;
    const float bound = (i < oW2) ? 1.0f : 0.17677669529663687f;
    out[i] = u * bound;
  }
  return kMP;
}

}  // namespace optmc
