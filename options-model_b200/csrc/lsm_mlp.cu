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
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"
#include "optmc_tc.cuh"

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
  // A CTA walks several 256-path compaction blocks (grid-stride): the per-block live counts come from one counting
  // barrier each, the moments are reduced and added to the fixed-point accumulators ONCE per CTA (a launch of one CTA
  // per block spent its time serialising 4 atomics per block on the same four words).
  __shared__ double red[kMWarps * 2];
  const int nblocks = (int)((M + kMThreads - 1) / kMThreads);
  double acc[2] = {0.0, 0.0};
  unsigned int total = 0u;
  for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
    const long long j = (long long)b * kMThreads + threadIdx.x;
    bool live = false;
    if (j < M) {
      const R s = S_t[j];
      live = mlp_live<R>(s, cf[j], K, is_put, sticky);
      if (live) { acc[0] += (double)s; acc[1] += (double)s * (double)s; }
    }
    const unsigned int c = (unsigned int)__syncthreads_count(live);
    if (threadIdx.x == 0) block_count[b] = c;
    total += c;
  }
  block_reduce_sum<2, kMWarps>(acc, red);
  if (threadIdx.x < 4 && total) {  // two quantities x (hi, lo) fixed-point chunks; every contributing CTA adds a bias
    unsigned long long hi, lo;
    const double v = (threadIdx.x >> 1) ? acc[1] : acc[0];  // warp 0 holds the block totals in every lane
    fx_encode(v * 9.5367431640625e-7, hi, lo);  // scaled by 2^-20: sum(S^2) of 4M paths stays inside the 2^43 range
    atomicAdd(mom_fx + threadIdx.x, (threadIdx.x & 1) ? lo : hi);
  }
  if (threadIdx.x == 0 && total) atomicAdd(mom_fx + 4, 1ull);  // contributing CTAs (bias count)
}

// ---- 2. scan + standardisation + fresh network ------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mlp_scan_kernel(unsigned int* block_count, int nblocks, unsigned long long* mom_fx, MlpState* st, float* params,
                float* adam_m, float* adam_v, unsigned long long* grad_fx, unsigned long long seed, int date, int H) {  // (the tcgen05 path packs W2 afterwards: mlp_tc_pack_kernel)
  const int P = 3 * H + H * H + H + 1;
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x;
  // exclusive scan of block_count (in place): every thread scans a contiguous run, thread 0 scans the 1024 run totals
  {
    const int per = (nblocks + 1023) / 1024;
    const int lo = tid * per, hi = lo + per < nblocks ? lo + per : nblocks;
    unsigned long long sum = 0ull;
    for (int i = lo; i < hi; ++i) sum += block_count[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
      unsigned long long run = 0ull;
      for (int i = 0; i < 1024; ++i) { const unsigned long long x = s_part[i]; s_part[i] = run; run += x; }
      s_base = run;
    }
    __syncthreads();
    unsigned int run = (unsigned int)s_part[tid];
    for (int i = lo; i < hi; ++i) { const unsigned int x = block_count[i]; block_count[i] = run; run += x; }
  }
  __syncthreads();
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
  const float hb = rsqrtf((float)H);
  for (int i = tid; i < P; i += blockDim.x) {
    const Philox4 p = philox_for((unsigned long long)i, (unsigned int)date, 0x4D4C50u, seed);
    const float u = ((float)(p.v[0] >> 8) + 0.5f) * (2.0f / 16777216.0f) - 1.0f;  // (-1, 1)
    const float bound = (i < 2 * H) ? 1.0f : hb;                                   // fan_in 1 | fan_in H
    params[i] = u * bound;
    adam_m[i] = 0.f; adam_v[i] = 0.f;
    if (grad_fx) { grad_fx[2 * i] = 0ull; grad_fx[2 * i + 1] = 0ull; }
  }
  if (tid == 0 && grad_fx) { grad_fx[2 * P] = 0ull; grad_fx[2 * P + 1] = 0ull; grad_fx[2 * P + 2] = 0ull; }
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
#pragma unroll 2
  for (int j = 0; j < kMH; ++j) {
    float z = sm.b2[j];
#pragma unroll
    for (int i = 0; i < kMH; ++i) z = fmaf(sm.W2[j * kMPad + i], h1[i], z);
    const float h = fmaxf(z, 0.f);
    out = fmaf(sm.w3[j], h, out);
  }
  (void)h2;
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
    {  // forward + backward of this thread's row; only h1 (then dh1) lives in registers, the rest goes to the tiles
      const float x = act ? xs[r] : 0.f;
      float h1[kMH];
      unsigned int m1 = 0u, m2 = 0u;  // h1 > 0, h2 > 0
#pragma unroll
      for (int i = 0; i < kMH; ++i) {
        h1[i] = fmaxf(fmaf(sm.w1[i], x, sm.b1[i]), 0.f);
        m1 |= (h1[i] > 0.f ? 1u : 0u) << i;
        sm.H1[tid * kMPad + i] = h1[i];
      }
      float out = sm.b3;
#pragma unroll 2
      for (int j = 0; j < kMH; ++j) {  // partial unroll: a full 32 x 32 unroll front-loads 1024 shared loads and spills
        float z = sm.b2[j];
#pragma unroll
        for (int i = 0; i < kMH; ++i) z = fmaf(sm.W2[j * kMPad + i], h1[i], z);
        const float h2 = fmaxf(z, 0.f);
        m2 |= (h2 > 0.f ? 1u : 0u) << j;
        sm.H2[tid * kMPad + j] = h2;
        out = fmaf(sm.w3[j], h2, out);
      }
      const float err = act ? out - ys[r] : 0.f;
      const float dout = err * inv_n2;
      loss = fmaf(err, err, loss);
      sm.xs[tid] = x; sm.dout[tid] = dout;
      float dh[kMH];
#pragma unroll
      for (int i = 0; i < kMH; ++i) dh[i] = 0.f;
#pragma unroll 2
      for (int j = 0; j < kMH; ++j) {
        const float dz = ((m2 >> j) & 1u) ? dout * sm.w3[j] : 0.f;
        sm.DZ2[tid * kMPad + j] = dz;
#pragma unroll
        for (int i = 0; i < kMH; ++i) dh[i] = fmaf(sm.W2[j * kMPad + i], dz, dh[i]);
      }
#pragma unroll
      for (int i = 0; i < kMH; ++i) sm.DH1[tid * kMPad + i] = ((m1 >> i) & 1u) ? dh[i] : 0.f;
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

__global__ void __launch_bounds__(kMThreads, 1)
mlp_forward_kernel(const float* __restrict__ params, const float* __restrict__ xs, const MlpState* __restrict__ st, float* cont) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
  const long long n = st->n_live;
  if ((long long)blockIdx.x * kMTile >= n) return;
  mlp_load_params(sm, params);
  __syncthreads();
  for (long long r = (long long)blockIdx.x * kMTile + threadIdx.x; r < n; r += (long long)gridDim.x * kMTile) {
    float h1[kMH], h2[kMH];
    cont[r] = mlp_forward_row(sm, xs[r], h1, h2);
  }
}

__global__ void mlp_nitm_kernel(const MlpState* st, long long* nitm_t, double* loss_t) {
  *nitm_t = st->n_live;
  if (loss_t) *loss_t = st->loss;
}


// =================================================================================================================
// Tensor-core path: hidden = 128 (the reference's default nn_hidden in om3:340-358).  The three 128 x 128 x 128
// contractions of a 128-row tile run on tcgen05 (bf16 operands staged in shared memory, fp32 accumulators in
// tensor memory); the vector gradients are three more N = 16 MMAs against a small [1, x, dout] panel:
//   Z2   = H1 W2^T          (A = H1 tile K-major,  B = W2 tile K-major)           -> TMEM cols [0,128)
//   dH1  = dZ2 W2           (A = dZ2 tile K-major, B = W2 tile MN-major view)     -> TMEM cols [0,128) (Z2 is consumed)
//   dW2 += dZ2^T H1         (A = dZ2 MN-major view, B = H1 MN-major view)         -> TMEM cols [128,256), across tiles
//   [db2 .]       += dZ2^T P,  [. . dw3] += H2^T P,  [db1 dw1 .] += dH1'^T P      -> TMEM cols [256,304)
// Every tile is stored once in the NO-SWIZZLE core-matrix layout (8 x 8 cores of 128 contiguous bytes, k-cores 128 B
// apart, 8-row groups 2048 B apart), which serves as a K-major operand [row][col] and, with LBO / SBO exchanged, as
// an MN-major operand [col][row] -- no transposed copies (validated by tools/umma_test.cu).  One thread per row
// (TMEM lane), one elected thread issues the MMAs, completion through tcgen05.commit -> mbarrier.
// =================================================================================================================
constexpr int kTH = 128;
constexpr int kTP = 3 * kTH + kTH * kTH + kTH + 1;  // 16897
constexpr int tW1 = 0, tB1 = kTH, tW2 = 2 * kTH, tB2 = 2 * kTH + kTH * kTH, tW3 = tB2 + kTH, tB3 = tW3 + kTH;
constexpr int kTcThreads = 512;            // thread = (row = tid & 127, column part = tid >> 7): 16 warps hide the TMEM / shared
                                           // latencies of the epilogues (8 warps: issue slots 30 % busy, profiles/ncu_r1_summary)
constexpr int kTcParts = kTcThreads / 128; // column parts per row (2 or 4)
constexpr int kTcG8 = 16 / kTcParts;       // 8-column groups per part
constexpr int kTcG32 = 4 / kTcParts;       // 32-column TMEM chunks per part
constexpr int kTileBytes = kTcTileBytes;   // 32 KB
constexpr int kAuxBytes = kTcPanelBytes;   // 4 KB panel [row][16]

struct TcSmem {
  unsigned char T1[kTileBytes], T2[kTileBytes], T3[kTileBytes], T4[kTileBytes], W2[kTileBytes];
  unsigned char aux[2][kAuxBytes];
  float w1[kTH], b1[kTH], b2[kTH], w3[kTH];
  float dot[kTcParts][128];       // partial output-layer dot products of the column parts
  unsigned int mask[2][4][128];   // per layer: 128 "unit is active" bits of each row (word-major: conflict-free)
  float b3;
  float red[8];
  unsigned long long bar, wbar;
  unsigned int tmem;
};

// bf16 core-layout image of W2, kept in step with the fp32 parameters (scan / Adam kernels): a CTA stages it with one
// bulk async copy instead of re-packing 16384 floats (the per-date fits run one tile per CTA: set-up time matters).
__device__ __forceinline__ int tc_pack_index(int i) { const int k = i - tW2; return core_off(k >> 7, k & 127) >> 1; }

// setup shared by the training and the inference kernel: W2 image by bulk copy, vectors, panels, TMEM, mbarriers
__device__ __forceinline__ unsigned int tc_setup(TcSmem& sm, const float* __restrict__ params,
                                                 const __nv_bfloat16* __restrict__ wpack, bool big) {
  const int tid = threadIdx.x, row = tid & 127;
  if (tid == 0) {
    mbar_init(reinterpret_cast<uint64_t*>(&sm.bar), 1);
    mbar_init(reinterpret_cast<uint64_t*>(&sm.wbar), 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.wbar), kTileBytes);
    bulk_load_1d(sm.W2, wpack, kTileBytes, reinterpret_cast<uint64_t*>(&sm.wbar));
  }
  if (tid < 128) { sm.w1[row] = params[tW1 + row]; sm.b1[row] = params[tB1 + row]; }
  else if (tid < 256) { sm.b2[row] = params[tB2 + row]; sm.w3[row] = params[tW3 + row]; }
  if (tid == 0) sm.b3 = params[tB3];
  if (tid < 256) {  // panels: [1, x, dout, 0 ...]; the constant and zero columns are written once (one panel per 128 threads)
    unsigned char* ax = sm.aux[tid >> 7];
    *reinterpret_cast<uint4*>(ax + aux_off(row, 0)) = make_uint4(0x00003f80u, 0u, 0u, 0u);  // bf16 1.0 in column 0
    *reinterpret_cast<uint4*>(ax + aux_off(row, 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
  if ((tid >> 5) == 0) {
    if (big) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem)), "n"(512));
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_publish();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return sm.tmem;
}

// The stage loops below are NOT unrolled: with one tile per CTA (the per-date fits of the sticky sweep) straight-line
// code would be fetched cold from the instruction cache end to end.

// S1: h1 = relu(w1 x + b1) -> bf16 H1 tile (this thread's 64 columns); mask bit = h1 > 0
__device__ __forceinline__ void tc_layer1(TcSmem& sm, float x, int row, int half) {
  unsigned int word = 0u;
#pragma unroll 1
  for (int c = half * kTcG8; c < half * kTcG8 + kTcG8; ++c) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float h = fmaf(sm.w1[c * 8 + k], x, sm.b1[c * 8 + k]);
      v[k] = fmaxf(h, 0.f);
      word |= (h > 0.f ? 1u : 0u) << ((c & 3) * 8 + k);
    }
    *reinterpret_cast<uint4*>(sm.T1 + core_off(row, c * 8)) = pack8_bf16(v);
    if ((c & 3) == 3) { sm.mask[0][c >> 2][row] = word; word = 0u; }
  }
}

// S2a: h2 = relu(z2 + b2) for this thread's 64 columns; returns the partial dot with w3; STORE: H2 tile + mask
template <bool STORE>
__device__ __forceinline__ float tc_layer2(TcSmem& sm, unsigned int taddr, int row, int half) {
  float out = 0.f;
#pragma unroll 1
  for (int c0 = half * kTcG32; c0 < half * kTcG32 + kTcG32; ++c0) {
    float z[32];
    tmem_ld32(taddr + c0 * 32, z);
    unsigned int word = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = c0 * 32 + q * 8 + k;
        const float h = z[q * 8 + k] + sm.b2[j];
        v[k] = fmaxf(h, 0.f);
        word |= (h > 0.f ? 1u : 0u) << (q * 8 + k);
        out = fmaf(sm.w3[j], v[k], out);
      }
      if (STORE) *reinterpret_cast<uint4*>(sm.T3 + core_off(row, c0 * 32 + q * 8)) = pack8_bf16(v);
    }
    if (STORE) sm.mask[1][c0][row] = word;
  }
  return out;
}

__device__ __forceinline__ float tc_join_dot(TcSmem& sm, float part, int row, int half) {
  sm.dot[half][row] = part;
  __syncthreads();
  if (kTcParts == 4) return ((sm.dot[0][row] + sm.dot[1][row]) + (sm.dot[2 % kTcParts][row] + sm.dot[3 % kTcParts][row])) + sm.b3;
  return (sm.dot[0][row] + sm.dot[1][row]) + sm.b3;
}

__device__ __forceinline__ void tc_issue_layer2(TcSmem& sm, unsigned int tmem) {  // Z2 = H1 W2^T
  const unsigned int a0 = smem_u32(sm.T1), b0 = smem_u32(sm.W2);
  const unsigned int id = umma_idesc(128, 128, 0, 0);
#pragma unroll 1
  for (int k = 0; k < 8; ++k) umma_f16(tmem, umma_desc(a0 + k * 256, 128, 2048), umma_desc(b0 + k * 256, 128, 2048), id, k > 0);
}

__global__ void __launch_bounds__(kTcThreads, 1)
mlp_tc_grad_kernel(const float* __restrict__ params, const __nv_bfloat16* __restrict__ wpack, const float* __restrict__ xs,
                   const float* __restrict__ ys, const MlpState* __restrict__ st, float* __restrict__ gpart) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_tc);
  const long long n = st->n_live;
  const long long ntiles = (n + 127) / 128;
  if ((long long)blockIdx.x >= ntiles) return;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const unsigned int tmem = tc_setup(sm, params, wpack, true);
  const unsigned int lane_base = (unsigned int)((warp & 3) * 32) << 16;
  const unsigned int cZ = 0, cW = 128, cV1 = 256, cV2 = 272, cV3 = 288;
  const float inv_n2 = (float)(2.0 / (double)n);
  unsigned int phase = 0;
  float loss = 0.f, gb3 = 0.f;
  int local = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++local) {
    const long long r = tile * 128 + row;
    const bool act = r < n;
    const float x = act ? xs[r] : 0.f, y = act ? ys[r] : 0.f;
    unsigned char* aux = sm.aux[local & 1];
    tc_layer1(sm, x, row, half);
    if (half == 0) *reinterpret_cast<unsigned short*>(aux + aux_off(row, 1)) = __bfloat16_as_ushort(__float2bfloat16_rn(x));
    tc_publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (local == 0) mbar_wait(reinterpret_cast<uint64_t*>(&sm.wbar), 0u);  // the W2 image has landed
      if (local > 0) {  // vector gradients of the previous tile's first layer: [db1 dw1 .] += dH1'^T P(prev)
        const unsigned int a0 = smem_u32(sm.T4), p0 = smem_u32(sm.aux[(local - 1) & 1]);
        const unsigned int idv = umma_idesc(128, 16, 1, 1);
#pragma unroll 1
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem + cV1, umma_desc(a0 + k * 4096, 2048, 128), umma_desc(p0 + k * 512, 256, 128), idv, (local > 1) || k > 0);
      }
      tc_issue_layer2(sm, tmem + cZ);
      umma_commit(&sm.bar);
    }
    tc_bar_wait(&sm.bar, phase); phase ^= 1u;
    // S2a: output and H2 tile
    const float out = tc_join_dot(sm, tc_layer2<true>(sm, tmem + lane_base + cZ, row, half), row, half);
    const float err = act ? out - y : 0.f;
    const float dout = err * inv_n2;
    if (half == 0) { loss = fmaf(err, err, loss); gb3 += dout; }
    // S2b: dZ2 tile (this thread's 64 columns) and the dout column of the panel
#pragma unroll 1
    for (int c = half * kTcG8; c < half * kTcG8 + kTcG8; ++c) {
      const unsigned int word = sm.mask[1][c >> 2][row] >> ((c & 3) * 8);
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ((word >> k) & 1u) ? dout * sm.w3[c * 8 + k] : 0.f;
      *reinterpret_cast<uint4*>(sm.T2 + core_off(row, c * 8)) = pack8_bf16(v);
    }
    if (half == 0) *reinterpret_cast<unsigned short*>(aux + aux_off(row, 2)) = __bfloat16_as_ushort(__float2bfloat16_rn(dout));
    tc_publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned int t1 = smem_u32(sm.T1), t2 = smem_u32(sm.T2), t3 = smem_u32(sm.T3), w2 = smem_u32(sm.W2), p0 = smem_u32(aux);
      const unsigned int id_kmn = umma_idesc(128, 128, 0, 1), id_mm = umma_idesc(128, 128, 1, 1), idv = umma_idesc(128, 16, 1, 1);
      const unsigned int accT = local > 0;
#pragma unroll 1
      for (int k = 0; k < 8; ++k)  // dH1 = dZ2 W2
        umma_f16(tmem + cZ, umma_desc(t2 + k * 256, 128, 2048), umma_desc(w2 + k * 4096, 2048, 128), id_kmn, k > 0);
#pragma unroll 1
      for (int k = 0; k < 8; ++k)  // dW2^T += H1^T dZ2  (lane = input unit: the read-out is coalesced)
        umma_f16(tmem + cW, umma_desc(t1 + k * 4096, 2048, 128), umma_desc(t2 + k * 4096, 2048, 128), id_mm, accT || k > 0);
#pragma unroll 1
      for (int k = 0; k < 8; ++k)  // [db2 . .] += dZ2^T P
        umma_f16(tmem + cV2, umma_desc(t2 + k * 4096, 2048, 128), umma_desc(p0 + k * 512, 256, 128), idv, accT || k > 0);
#pragma unroll 1
      for (int k = 0; k < 8; ++k)  // [. . dw3] += H2^T P
        umma_f16(tmem + cV3, umma_desc(t3 + k * 4096, 2048, 128), umma_desc(p0 + k * 512, 256, 128), idv, accT || k > 0);
      umma_commit(&sm.bar);
    }
    tc_bar_wait(&sm.bar, phase); phase ^= 1u;
    // S3: dH1' = dH1 (h1 > 0) tile
#pragma unroll 1
    for (int c0 = half * kTcG32; c0 < half * kTcG32 + kTcG32; ++c0) {
      float d[32];
      tmem_ld32(tmem + lane_base + cZ + c0 * 32, d);
      const unsigned int word = sm.mask[0][c0][row];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ((word >> (q * 8 + k)) & 1u) ? d[q * 8 + k] : 0.f;
        *reinterpret_cast<uint4*>(sm.T4 + core_off(row, c0 * 32 + q * 8)) = pack8_bf16(v);
      }
    }
  }
  // tail: first-layer vector gradients of the last tile
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int a0 = smem_u32(sm.T4), p0 = smem_u32(sm.aux[(local - 1) & 1]);
    const unsigned int idv = umma_idesc(128, 16, 1, 1);
#pragma unroll 1
    for (int k = 0; k < 8; ++k)
      umma_f16(tmem + cV1, umma_desc(a0 + k * 4096, 2048, 128), umma_desc(p0 + k * 512, 256, 128), idv, (local > 1) || k > 0);
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  // read-out: dW2 sits transposed in TMEM (lane = input unit i, column = output unit j): for a fixed j the threads of a
  // warp write consecutive addresses of row j of W2's gradient.  The vector gradients have lane = unit.
  float* gp = gpart + (size_t)blockIdx.x * (kTP + 1);
#pragma unroll 1
  for (int c0 = half * kTcG32; c0 < half * kTcG32 + kTcG32; ++c0) {
    float w[32];
    tmem_ld32(tmem + lane_base + cW + c0 * 32, w);
#pragma unroll
    for (int i = 0; i < 32; ++i) gp[tW2 + (c0 * 32 + i) * kTH + row] = w[i];
  }
  if (half == 0) {
    float v1[16];
    tmem_ld16(tmem + lane_base + cV1, v1);
    gp[tB1 + row] = v1[0]; gp[tW1 + row] = v1[1];
  } else if (half == 1) {
    float v2[16], v3[16];
    tmem_ld16(tmem + lane_base + cV2, v2);
    tmem_ld16(tmem + lane_base + cV3, v3);
    gp[tB2 + row] = v2[0]; gp[tW3 + row] = v3[2];
  }
  {  // scalars: db3 and the loss (column half 0 carries them)
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { gb3 += __shfl_xor_sync(0xffffffffu, gb3, m); loss += __shfl_xor_sync(0xffffffffu, loss, m); }
    if (half == 0 && (tid & 31) == 0) { sm.red[warp] = gb3; sm.red[4 + warp] = loss; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    gp[tB3] = (sm.red[0] + sm.red[1]) + (sm.red[2] + sm.red[3]);
    gp[kTP] = (sm.red[4] + sm.red[5]) + (sm.red[6] + sm.red[7]);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// Adam for the tensor-core path: deterministic fixed-order sum of the per-CTA partial gradients; keeps the bf16 image
// of W2 in step; the extra last block reduces the loss (st->loss).  `step` = 1-based optimiser step of this date.
__global__ void __launch_bounds__(256)
mlp_tc_adam_kernel(float* params, __nv_bfloat16* wpack, float* adam_m, float* adam_v, const float* __restrict__ gpart,
                   MlpState* st, float lr, int grid_rows, int step) {
  const long long n = st->n_live;
  if (n <= 0) return;
  const long long ntiles = (n + 127) / 128;
  const int nb = (int)(ntiles < grid_rows ? ntiles : grid_rows);
  if (blockIdx.x == gridDim.x - 1) {  // loss block
    if (threadIdx.x < 32) {
      double l = 0.0;
      for (int b = threadIdx.x; b < nb; b += 32) l += (double)gpart[(size_t)b * (kTP + 1) + kTP];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) l += __shfl_xor_sync(0xffffffffu, l, m);
      if (threadIdx.x == 0) { st->loss = l / (double)n; st->step = step; }
    }
    return;
  }
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kTP) {
    float g = 0.f;
    for (int b = 0; b < nb; ++b) g += gpart[(size_t)b * (kTP + 1) + i];
    const float m = b1 * adam_m[i] + (1.0f - b1) * g;
    const float v = b2 * adam_v[i] + (1.0f - b2) * g * g;
    adam_m[i] = m; adam_v[i] = v;
    const float p = params[i] - (lr / bc1) * (m / (sqrtf(v) / sqrtf(bc2) + eps));
    params[i] = p;
    if (i >= tW2 && i < tB2) wpack[tc_pack_index(i)] = __float2bfloat16_rn(p);
  }
}
__global__ void mlp_tc_pack_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ wpack) {
  const int i = tW2 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tB2) wpack[tc_pack_index(i)] = __float2bfloat16_rn(params[i]);
}

// continuation = net(x) on tensor cores, then the exercise decision (strict '>', om2:304)
template <typename R>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_tc_decide_kernel(const float* __restrict__ params, const __nv_bfloat16* __restrict__ wpack, const float* __restrict__ xs,
                     const unsigned int* __restrict__ idx, const MlpState* __restrict__ st, const R* __restrict__ S_t, R* cf,
                     double K, double Kh, double Kl, int is_put, int sticky, R dinv, float* cont_out,
                     unsigned long long* exc_t, unsigned long long* bnd_t) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_tc);
  const long long n = st->n_live;
  const long long ntiles = (n + 127) / 128;
  if ((long long)blockIdx.x >= ntiles) return;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const unsigned int tmem = tc_setup(sm, params, wpack, false);
  if (tid == 0) mbar_wait(reinterpret_cast<uint64_t*>(&sm.wbar), 0u);
  const unsigned int lane_base = (unsigned int)((warp & 3) * 32) << 16;
  const R sgn = is_put ? (R)-1 : (R)1;
  const R c1 = (R)(is_put ? Kh : -Kh), c2 = (R)(is_put ? Kl : -Kl);
  unsigned int phase = 0, cnt = 0;
  unsigned long long bnd = bnd_none(is_put);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r = tile * 128 + row;
    const bool act = r < n;
    tc_layer1(sm, act ? xs[r] : 0.f, row, half);
    tc_publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tc_issue_layer2(sm, tmem);
      umma_commit(&sm.bar);
    }
    tc_bar_wait(&sm.bar, phase); phase ^= 1u;
    const float out = tc_join_dot(sm, tc_layer2<false>(sm, tmem + lane_base, row, half), row, half);
    if (act && half == 0) {
      if (cont_out) cont_out[r] = out;
      if (idx) {
        const unsigned int j = idx[r];
        const R sr = S_t[j];
        const double s = (double)sr;
        const double pay = payoff<double>(s, K, is_put != 0);
        if (pay > (double)out) {
          const R a = (fma(sgn, sr, c1) + c2) * dinv;  // payoff in date-N money
          cf[j] = sticky ? -a : a;
          cnt++;
          const unsigned long long b = (unsigned long long)__double_as_longlong(s);
          bnd = is_put ? (b > bnd ? b : bnd) : (b < bnd ? b : bnd);
        }
      }
    }
    __syncthreads();  // sm.dot is rewritten by the next tile
  }
  if (half == 0) {
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    bnd = is_put ? warp_max_u64(bnd) : warp_min_u64(bnd);
    if ((tid & 31) == 0 && cnt && exc_t) {
      atomicAdd(exc_t, (unsigned long long)cnt);
      if (is_put) atomicMax(bnd_t, bnd); else atomicMin(bnd_t, bnd);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

// ---- host driver --------------------------------------------------------------------------------------------
template <typename R> static int lsm_mlp_t(optmc_ctx* ctx, const optmc_mlp_params* np_, optmc_lsm_result* out) {
  SweepDesc& sw = ctx->sw;
  const long long M = sw.M;
  const int N = sw.N;
  const int nblocks = (int)((M + kMThreads - 1) / kMThreads);
  const int count_grid = nblocks < 8 * ctx->sm_count ? nblocks : 8 * ctx->sm_count;  // grid-stride over the compaction blocks
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
    mlp_count_kernel<R><<<count_grid, kMThreads, 0, ctx->stream>>>(S_t, cf, M, sw.lp.K, sw.lp.is_put, sticky, d_cnt, d_mom);
    mlp_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, nblocks, d_mom, d_st, d_par, d_m, d_v, d_g, np_->seed, t, kMH);
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


template <typename R> static int lsm_mlp_tc_t(optmc_ctx* ctx, const optmc_mlp_params* np_, optmc_lsm_result* out) {
  SweepDesc& sw = ctx->sw;
  const long long M = sw.M;
  const int N = sw.N;
  const int nblocks = (int)((M + kMThreads - 1) / kMThreads);
  const int count_grid = nblocks < 8 * ctx->sm_count ? nblocks : 8 * ctx->sm_count;  // grid-stride over the compaction blocks
  const bool sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) != 0;
  long long tiles = (M + 127) / 128;
  const int grid_rows = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);  // one CTA per SM (tiles + TMEM)
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_cnt = take((size_t)nblocks * 4), o_xs = take((size_t)M * 4), o_ys = take((size_t)M * 4),
               o_idx = take((size_t)M * 4), o_par = take(kTP * 4), o_m = take(kTP * 4), o_v = take(kTP * 4),
               o_g = take((size_t)grid_rows * (kTP + 1) * 4), o_mom = take(5 * 8), o_st = take(sizeof(MlpState)),
               o_loss = take((size_t)(N + 1) * 8), o_pack = take(kTileBytes);
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
  float* d_g = reinterpret_cast<float*>(dev + o_g);
  unsigned long long* d_mom = reinterpret_cast<unsigned long long*>(dev + o_mom);
  MlpState* d_st = reinterpret_cast<MlpState*>(dev + o_st);
  double* d_loss = reinterpret_cast<double*>(dev + o_loss);
  __nv_bfloat16* d_pack = reinterpret_cast<__nv_bfloat16*>(dev + o_pack);
  OPTMC_CUDA(cudaMemsetAsync(d_mom, 0, 5 * 8, ctx->stream));
  OPTMC_CUDA(cudaMemsetAsync(d_loss, 0, (size_t)(N + 1) * 8, ctx->stream));
  rc = sweep_begin(ctx);
  if (rc) return rc;
  const R* Sr = static_cast<const R*>(sw.S);
  R* cf = static_cast<R*>(ctx->cf);
  const size_t smem = sizeof(TcSmem) + 1024;
  OPTMC_CUDA(cudaFuncSetAttribute(mlp_tc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OPTMC_CUDA(cudaFuncSetAttribute(mlp_tc_decide_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int adam_grid = (kTP + 255) / 256;
  for (int t = N - 1; t >= 1; --t) {
    const R* S_t = Sr + (size_t)t * sw.ld;
    mlp_count_kernel<R><<<count_grid, kMThreads, 0, ctx->stream>>>(S_t, cf, M, sw.lp.K, sw.lp.is_put, sticky, d_cnt, d_mom);
    mlp_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cnt, nblocks, d_mom, d_st, d_par, d_m, d_v, nullptr, np_->seed, t, kTH);
    mlp_compact_kernel<R><<<nblocks, kMThreads, 0, ctx->stream>>>(S_t, cf, M, sw.lp.K, sw.lp.is_put, sticky, sw.Dt[t], d_cnt,
                                                                  d_st, d_xs, d_ys, d_idx);
    mlp_tc_pack_kernel<<<kTH * kTH / 256, 256, 0, ctx->stream>>>(d_par, d_pack);
    for (int e = 0; e < np_->epochs; ++e) {
      mlp_tc_grad_kernel<<<grid_rows, kTcThreads, smem, ctx->stream>>>(d_par, d_pack, d_xs, d_ys, d_st, d_g);
      mlp_tc_adam_kernel<<<adam_grid + 1, 256, 0, ctx->stream>>>(d_par, d_pack, d_m, d_v, d_g, d_st, (float)np_->lr, grid_rows, e + 1);
    }
    mlp_tc_decide_kernel<R><<<grid_rows, kTcThreads, smem, ctx->stream>>>(d_par, d_pack, d_xs, d_idx, d_st, S_t, cf, sw.lp.K, sw.Kh,
                                                                         sw.Kl, sw.lp.is_put, sticky, (R)sw.Dinv[t], nullptr,
                                                                         ctx->d_exc + t, ctx->d_bnd + t);
    mlp_nitm_kernel<<<1, 1, 0, ctx->stream>>>(d_st, ctx->d_nitm + t, d_loss + t);
    ctx->launches += 6 + 2 * np_->epochs; sw.n_launches += 6 + 2 * np_->epochs;
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

__global__ void mlp_set_state_kernel(MlpState* st, long long n) { st->n_live = n; st->mean = 0.0; st->inv_std = 1.0; st->loss = 0.0; st->step = 0; }
__global__ void mlp_fx_to_float_kernel(const unsigned long long* grad_fx, float* out, int P) {
  const int nb = (int)grad_fx[2 * P + 2];
  for (int i = threadIdx.x; i < P; i += blockDim.x) out[i] = nb ? (float)fx_decode(grad_fx[2 * i], grad_fx[2 * i + 1], nb) : 0.f;
}
__global__ void mlp_sum_partials_kernel(const float* gpart, int nb, float* out, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float g = 0.f;
  for (int b = 0; b < nb; ++b) g += gpart[(size_t)b * (P + 1) + i];
  out[i] = g;
}

// Test aid (optmc_mlp_grad_debug): one gradient evaluation + forward on n host rows.
int mlp_grad_debug(optmc_ctx* ctx, int H, long long n, const float* xs, const float* ys, const float* params, float* grads,
                   float* cont) {
  if ((H != kMH && H != kTH) || n <= 0 || !xs || !ys || !params) { set_error("bad arguments"); return OPTMC_EINVAL; }
  const int P = 3 * H + H * H + H + 1;
  const int tile = H == kTH ? 128 : kMTile;
  long long tiles = (n + tile - 1) / tile;
  const int grid_rows = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_xs = take((size_t)n * 4), o_ys = take((size_t)n * 4), o_cont = take((size_t)n * 4), o_par = take((size_t)P * 4),
               o_out = take((size_t)P * 4), o_g = take(H == kTH ? (size_t)grid_rows * (P + 1) * 4 : (size_t)(2 * P + 3) * 8),
               o_st = take(sizeof(MlpState)), o_pack = take(kTileBytes);
  int rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  float* d_xs = reinterpret_cast<float*>(dev + o_xs);
  float* d_ys = reinterpret_cast<float*>(dev + o_ys);
  float* d_cont = reinterpret_cast<float*>(dev + o_cont);
  float* d_par = reinterpret_cast<float*>(dev + o_par);
  float* d_out = reinterpret_cast<float*>(dev + o_out);
  MlpState* d_st = reinterpret_cast<MlpState*>(dev + o_st);
  OPTMC_CUDA(cudaMemcpyAsync(d_xs, xs, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(d_ys, ys, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(d_par, params, (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
  mlp_set_state_kernel<<<1, 1, 0, ctx->stream>>>(d_st, n);
  if (H == kTH) {
    float* d_g = reinterpret_cast<float*>(dev + o_g);
    const size_t smem = sizeof(TcSmem) + 1024;
    OPTMC_CUDA(cudaFuncSetAttribute(mlp_tc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OPTMC_CUDA(cudaFuncSetAttribute(mlp_tc_decide_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    __nv_bfloat16* d_pack = reinterpret_cast<__nv_bfloat16*>(dev + o_pack);
    mlp_tc_pack_kernel<<<kTH * kTH / 256, 256, 0, ctx->stream>>>(d_par, d_pack);
    mlp_tc_grad_kernel<<<grid_rows, kTcThreads, smem, ctx->stream>>>(d_par, d_pack, d_xs, d_ys, d_st, d_g);
    mlp_sum_partials_kernel<<<(P + 255) / 256, 256, 0, ctx->stream>>>(d_g, grid_rows, d_out, P);
    mlp_tc_decide_kernel<float><<<grid_rows, kTcThreads, smem, ctx->stream>>>(d_par, d_pack, d_xs, nullptr, d_st, nullptr, nullptr, 1.0,
                                                                             1.0, 0.0, 1, 0, 1.0f, d_cont, nullptr, nullptr);
    ctx->launches += 5;
  } else {
    unsigned long long* d_g = reinterpret_cast<unsigned long long*>(dev + o_g);
    OPTMC_CUDA(cudaMemsetAsync(d_g, 0, (size_t)(2 * P + 3) * 8, ctx->stream));
    const size_t smem = sizeof(MlpSmem);
    OPTMC_CUDA(cudaFuncSetAttribute(mlp_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OPTMC_CUDA(cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_grad_kernel<<<grid_rows, kMThreads, smem, ctx->stream>>>(d_par, d_xs, d_ys, d_st, d_g);
    mlp_fx_to_float_kernel<<<1, 256, 0, ctx->stream>>>(d_g, d_out, P);
    mlp_forward_kernel<<<grid_rows, kMThreads, smem, ctx->stream>>>(d_par, d_xs, d_st, d_cont);
    ctx->launches += 4;
  }
  OPTMC_CUDA(cudaGetLastError());
  if (grads) OPTMC_CUDA(cudaMemcpyAsync(grads, d_out, (size_t)P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (cont) OPTMC_CUDA(cudaMemcpyAsync(cont, d_cont, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPTMC_OK;
}

int lsm_mlp(optmc_ctx* ctx, const optmc_mlp_params* np_, optmc_lsm_result* out) {
  if (!np_) { set_error("null argument"); return OPTMC_EINVAL; }
  if (np_->hidden != kMH && np_->hidden != kTH) {
    set_error("per-date MLP: hidden width must be 32 (CUDA cores; om2 default nn_hidden) or 128 (tcgen05 tensor cores)");
    return OPTMC_EUNSUPPORTED;
  }
  if (np_->epochs < 0 || np_->epochs > 10000 || !(np_->lr > 0)) { set_error("per-date MLP: bad epochs / lr"); return OPTMC_EINVAL; }
  if (ctx->sw.M >= (1ll << 32)) { set_error("per-date MLP: too many paths"); return OPTMC_EUNSUPPORTED; }
  if (np_->hidden == kTH) return ctx->sw.dtype == OPTMC_F64 ? lsm_mlp_tc_t<double>(ctx, np_, out) : lsm_mlp_tc_t<float>(ctx, np_, out);
  return ctx->sw.dtype == OPTMC_F64 ? lsm_mlp_t<double>(ctx, np_, out) : lsm_mlp_t<float>(ctx, np_, out);
}

// The fresh parameters of date `date` (test aid: lets the oracle start from the same network).
int mlp_init_params_host(int H, unsigned long long seed, int date, float* out) {
  const int P = 3 * H + H * H + H + 1;
  const float hb = 1.0f / sqrtf((float)H);
  for (int i = 0; i < P; ++i) {
    const Philox4 p = philox_for((unsigned long long)i, (unsigned int)date, 0x4D4C50u, seed);
    const float u = ((float)(p.v[0] >> 8) + 0.5f) * (2.0f / 16777216.0f) - 1.0f;
    const float bound = (i < 2 * H) ? 1.0f : hb;
    out[i] = u * bound;
  }
  return P;
}

}  // namespace optmc
