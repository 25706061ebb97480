// paths_localvol.cu -- local-volatility paths with the implied-volatility network evaluated INSIDE the step
// (SURVEY 8f n3).  Replaces om3:263-333 (`IVModel.get_volatility_batch` + `simulate_local_vol_paths_antithetic`) and
// om3gpu:250-298: per step the reference ships all paths through a torch MLP (numpy -> torch -> numpy); here one thread
// owns one path for all N steps, the network weights sit in shared memory and the activations in registers.
//
//   sigma_t = max(net([ln(K / S_{t-1}) / m_scale, tau_t / tau_scale]), epsilon, 1e-6),  tau_t = max(T - (t-1) dt, 1e-6)
//   S_t     = S_{t-1} exp((r - sigma_t^2 / 2) dt + sigma_t sqrt(dt) z_t),  z = [Z_half, -Z_half]      (om3:311-318)
//   net (nniv:109-155, `ImprovedIVNetwork`): h = GELU(W_in x + b_in); L times h += GELU(LayerNorm(W_l h + b_l)); W_out h + b_out
//
// The moneyness is scaled but NOT centred (om3:285-288 divide by the scaler's std only, although training subtracts
// the mean, nniv:97-98 -- SURVEY App. B-7); this kernel reproduces the pricer, not the training convention.
// Arithmetic: the network in fp32 like the reference's `.float()` tensors (exact GELU through erff, LayerNorm with
// the biased variance and eps = 1e-5), the log-moneyness and the price update in the storage type.  The work is
// 2 H + L H^2 + H fused multiply-adds per path-step (16.6 k for the default H = 64, L = 4): CUDA-core bound by
// design -- bf16 tensor-core products would move sigma in the third digit, outside the parity tolerance.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

constexpr int kLvThreads = 256;   // sigma kernel; the path kernel runs 512 (fp32) / 384 (fp64) threads: one CTA per SM,
                                  // weights + one pre-activation column per thread in shared memory

struct LvArgs {
  void* S;
  long long ld, M, Mh, per_cta;
  int N, anti;
  const void* z1;
  int z_f64;
  unsigned long long seed;
  unsigned int stream;
  long long pair_offset;
  double S0, r, T, dt, sqrt_dt, K;
  const float* weights;  // device: state_dict order (see optmc_ivnet in include/optmc.h)
  int layers;
  double m_scale, tau_scale;
  float epsilon;
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// sigma = max(net([x0, x1]), epsilon, 1e-6); w = weights in shared memory, zc = this thread's pre-activation column
template <int H, int NT>
__device__ __forceinline__ float ivnet_sigma(const float* w, int layers, float x0, float x1, float* zc, float epsilon) {
  const float* w_in = w;            // [H][2]
  const float* b_in = w + 2 * H;    // [H]
  float h[H];
#pragma unroll
  for (int i = 0; i < H; ++i) h[i] = gelu_exact(fmaf(w_in[2 * i], x0, fmaf(w_in[2 * i + 1], x1, b_in[i])));
  const float* wl = w + 3 * H;
#pragma unroll 1
  for (int l = 0; l < layers; ++l) {
    const float* W = wl;               // [H][H] (out, in)
    const float* b = wl + H * H;
    const float* g = b + H;
    const float* be = g + H;
    float sum = 0.f;
#pragma unroll 2
    for (int j = 0; j < H; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(W + j * H);
      float acc0 = b[j], acc1 = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < H / 4; ++i4) {
        const float4 q = wr[i4];
        acc0 = fmaf(q.x, h[4 * i4], acc0);
        acc1 = fmaf(q.y, h[4 * i4 + 1], acc1);
        acc0 = fmaf(q.z, h[4 * i4 + 2], acc0);
        acc1 = fmaf(q.w, h[4 * i4 + 3], acc1);
      }
      const float zj = acc0 + acc1;
      zc[j * NT] = zj;
      sum += zj;
    }
    const float mean = sum * (1.0f / H);
    float var = 0.f;
#pragma unroll 8
    for (int j = 0; j < H; ++j) { const float d = zc[j * NT] - mean; var = fmaf(d, d, var); }
    const float rstd = rsqrtf(var * (1.0f / H) + 1e-5f);
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] += gelu_exact(fmaf((zc[j * NT] - mean) * rstd, g[j], be[j]));
    wl += H * H + 3 * H;
  }
  const float* w_out = wl;
  float o = w_out[H];
#pragma unroll
  for (int j = 0; j < H; ++j) o = fmaf(w_out[j], h[j], o);
  return fmaxf(fmaxf(o, epsilon), 1e-6f);  // nniv:155 clamp(min = epsilon), om3:293 clamp_min(1e-6)
}

// shared memory: weights (state_dict order) followed by the per-thread pre-activation columns z[H][threads]
template <typename R, int H, int NT>
__global__ void __launch_bounds__(NT, 1) paths_localvol_kernel(const LvArgs a) {
  extern __shared__ __align__(16) float smem_lv[];
  const int n_w = 3 * H + a.layers * (H * H + 3 * H) + H + 1;
  const int n_w4 = (n_w + 3) & ~3;
  float* w = smem_lv;
  float* zcol = smem_lv + n_w4;  // [H][NT]
  for (int i = threadIdx.x; i < n_w; i += NT) w[i] = a.weights[i];
  __syncthreads();
  // a CTA owns the contiguous paths [blockIdx.x per_cta, ...): one balanced wave of CTAs whatever the path count
  const long long p_end = (long long)(blockIdx.x + 1) * a.per_cta < a.M ? (long long)(blockIdx.x + 1) * a.per_cta : a.M;
  for (long long p = (long long)blockIdx.x * a.per_cta + threadIdx.x; p < p_end; p += NT) {
  const bool minus = a.anti && p >= a.Mh;
  const long long col = minus ? p - a.Mh : p;   // the pair this path belongs to (om3:308-309 column layout)
  const float sign = minus ? -1.f : 1.f;
  R* out = static_cast<R*>(a.S) + p;
  R s = (R)a.S0;
  out[0] = s;
  float* zc = zcol + threadIdx.x;
  for (int t0 = 0; t0 < a.N; t0 += 4) {
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (!a.z1) {
      Philox4 ph = philox_for((unsigned long long)(a.pair_offset + col), (unsigned int)(t0 >> 2), a.stream, a.seed);
      R n0, n1, n2, n3;
      Real<R>::normal2(ph.v[0], ph.v[1], n0, n1);
      Real<R>::normal2(ph.v[2], ph.v[3], n2, n3);
      nrm[0] = (float)n0; nrm[1] = (float)n1; nrm[2] = (float)n2; nrm[3] = (float)n3;
    }
#pragma unroll 1
    for (int sidx = 0; sidx < 4; ++sidx) {
      const int t = t0 + sidx + 1;
      if (t > a.N) break;
      double z;
      if (a.z1) {
        const long long idx = (long long)(t - 1) * a.Mh + col;
        z = a.z_f64 ? static_cast<const double*>(a.z1)[idx] : (double)static_cast<const float*>(a.z1)[idx];
      } else {
        z = (double)nrm[sidx];
      }
      z *= (double)sign;
      // network inputs (om3:281-289)
      double tau = a.T - (t - 1) * a.dt;
      if (tau < 1e-6) tau = 1e-6;
      const double sp = (double)s > 1e-8 ? (double)s : 1e-8;
      const double kp = a.K > 1e-8 ? a.K : 1e-8;
      const float x0 = (float)(log(kp / sp) / a.m_scale);   // float64 quotient, then .float() (om3:286-291)
      const float x1 = (float)(tau / a.tau_scale);
      const float sig = ivnet_sigma<H, NT>(w, a.layers, x0, x1, zc, a.epsilon);
      const double sg = (double)sig;
      s = (R)((double)s * exp((a.r - 0.5 * sg * sg) * a.dt + sg * a.sqrt_dt * z));
      out[(size_t)t * a.ld] = s;
    }
  }
  }
}

// IVModel.get_volatility_batch (om3:277-298): sigma for n spots at one time to expiry
template <int H>
__global__ void __launch_bounds__(kLvThreads, 1) ivnet_sigma_kernel(const float* __restrict__ weights, int layers, double K,
                                                                    double m_scale, double tau_scale, float epsilon, double tau,
                                                                    const double* __restrict__ S, long long n, double* __restrict__ out) {
  extern __shared__ __align__(16) float smem_lv[];
  const int n_w = 3 * H + layers * (H * H + 3 * H) + H + 1;
  const int n_w4 = (n_w + 3) & ~3;
  for (int i = threadIdx.x; i < n_w; i += kLvThreads) smem_lv[i] = weights[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * kLvThreads + threadIdx.x;
  if (p >= n) return;
  if (tau < 1e-6) tau = 1e-6;
  const double sp = S[p] > 1e-8 ? S[p] : 1e-8, kp = K > 1e-8 ? K : 1e-8;
  out[p] = (double)ivnet_sigma<H, kLvThreads>(smem_lv, layers, (float)(log(kp / sp) / m_scale), (float)(tau / tau_scale),
                                  smem_lv + n_w4 + threadIdx.x, epsilon);
}

static int ivnet_count(int H, int L) { return 3 * H + L * (H * H + 3 * H) + H + 1; }

static int ivnet_check(optmc_ctx* ctx, const optmc_ivnet* net, size_t* smem);

int launch_paths_localvol(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, const optmc_ivnet* net,
                          int64_t M, int32_t N, int32_t dtype, void* S, int64_t ld) {
  if (!mp || !rng || !net || !S) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(mp->S0 > 0) || !(mp->T > 0) || !(net->K > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  if (rng->antithetic && (M % 2)) { set_error("antithetic layout needs an even path count"); return OPTMC_EINVAL; }
  size_t smem = 0;
  int rc = ivnet_check(ctx, net, &smem);
  if (rc) return rc;
  const int H = net->hidden, n_w = ivnet_count(H, net->layers);
  rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, (size_t)n_w * 4);
  if (rc) return rc;
  OPTMC_CUDA(cudaMemcpyAsync(ctx->batch_dev, net->weights, (size_t)n_w * 4, cudaMemcpyHostToDevice, ctx->stream));
  LvArgs a{};
  a.S = S; a.ld = ld; a.M = M; a.Mh = rng->antithetic ? M / 2 : M; a.N = N; a.anti = rng->antithetic ? 1 : 0;
  a.z1 = rng->z1_dev; a.z_f64 = rng->z_dtype == OPTMC_F64;
  a.seed = rng->seed; a.stream = (unsigned int)rng->stream; a.pair_offset = rng->pair_offset;
  a.S0 = mp->S0; a.r = mp->r; a.T = mp->T; a.dt = mp->T / N; a.sqrt_dt = sqrt(a.dt); a.K = net->K;
  a.weights = static_cast<const float*>(ctx->batch_dev); a.layers = net->layers;
  a.m_scale = net->m_scale; a.tau_scale = net->tau_scale; a.epsilon = net->epsilon;
  // one wave: per_cta = ceil(M / SMs) paths per CTA, and the thread count that wastes the fewest thread slots
  const long long per_cta = (M + ctx->sm_count - 1) / ctx->sm_count;
  a.per_cta = per_cta;
  const unsigned grid = (unsigned)((M + per_cta - 1) / per_cta);
  auto waste = [&](int nt) { const long long it = (per_cta + nt - 1) / nt; return (double)(it * nt) / (double)per_cta; };
#define OPTMC_LV_LAUNCH(R_, H_, NT_)                                                                                    \
  do {                                                                                                                  \
    const size_t sm_ = (size_t)((n_w + 3) & ~3) * 4 + (size_t)H_ * NT_ * 4;                                             \
    if (sm_ > (size_t)ctx->max_smem_optin) { set_error("IV network does not fit shared memory"); return OPTMC_EUNSUPPORTED; } \
    OPTMC_CUDA(cudaFuncSetAttribute(paths_localvol_kernel<R_, H_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_)); \
    paths_localvol_kernel<R_, H_, NT_><<<grid, NT_, sm_, ctx->stream>>>(a);                                             \
  } while (0)
#define OPTMC_LV_PICK(R_, H_, BIG_)                                                                  \
  do {                                                                                               \
    int nt = 256;                                                                                    \
    if (waste(384) <= waste(nt) + 0.02) nt = 384;                                                    \
    if (BIG_ && waste(512) <= waste(nt) + 0.02) nt = 512;                                            \
    if (const char* e_ = getenv("OPTMC_LV_NT")) { const int v_ = atoi(e_); if (v_ == 256 || v_ == 384 || (BIG_ && v_ == 512)) nt = v_; } \
    if (nt == 512) OPTMC_LV_LAUNCH(R_, H_, 512);                                                     \
    else if (nt == 384) OPTMC_LV_LAUNCH(R_, H_, 384);                                                \
    else OPTMC_LV_LAUNCH(R_, H_, 256);                                                               \
  } while (0)
  if (dtype == OPTMC_F64) { if (H == 64) OPTMC_LV_PICK(double, 64, false); else OPTMC_LV_PICK(double, 32, false); }
  else { if (H == 64) OPTMC_LV_PICK(float, 64, true); else OPTMC_LV_PICK(float, 32, true); }
#undef OPTMC_LV_PICK
#undef OPTMC_LV_LAUNCH
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  // the weights were staged from host memory: do not let the caller free them under the copy
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPTMC_OK;
}

static int ivnet_check(optmc_ctx* ctx, const optmc_ivnet* net, size_t* smem) {
  if (!net || !net->weights || net->layers < 0 || net->layers > 16) { set_error("bad IV network"); return OPTMC_EINVAL; }
  if (net->hidden != 32 && net->hidden != 64) { set_error("IV network: hidden_dim must be 32 or 64 (nniv default 64)"); return OPTMC_EUNSUPPORTED; }
  if (!(net->m_scale > 0) || !(net->tau_scale > 0)) { set_error("IV network: the scaler is not fitted (om3:271-272)"); return OPTMC_EINVAL; }
  const int n_w = ivnet_count(net->hidden, net->layers);
  if (net->n_weights != n_w) { set_error("IV network: weight count does not match hidden_dim / num_hidden_layers"); return OPTMC_EINVAL; }
  *smem = (size_t)((n_w + 3) & ~3) * 4 + (size_t)net->hidden * kLvThreads * 4;
  if (*smem > (size_t)ctx->max_smem_optin) { set_error("IV network does not fit shared memory"); return OPTMC_EUNSUPPORTED; }
  return OPTMC_OK;
}

int ivnet_sigma_batch(optmc_ctx* ctx, const optmc_ivnet* net, double tau, const double* S_dev, int64_t n, double* sigma_dev) {
  size_t smem = 0;
  int rc = ivnet_check(ctx, net, &smem);
  if (rc) return rc;
  if (!S_dev || !sigma_dev || n <= 0) { set_error("bad argument"); return OPTMC_EINVAL; }
  if (!(net->K > 0)) { set_error("K must be positive"); return OPTMC_EINVAL; }
  const int n_w = ivnet_count(net->hidden, net->layers);
  rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, (size_t)n_w * 4);
  if (rc) return rc;
  OPTMC_CUDA(cudaMemcpyAsync(ctx->batch_dev, net->weights, (size_t)n_w * 4, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = (unsigned)((n + kLvThreads - 1) / kLvThreads);
  const float* w = static_cast<const float*>(ctx->batch_dev);
  if (net->hidden == 64) {
    OPTMC_CUDA(cudaFuncSetAttribute(ivnet_sigma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ivnet_sigma_kernel<64><<<grid, kLvThreads, smem, ctx->stream>>>(w, net->layers, net->K, net->m_scale, net->tau_scale, net->epsilon, tau, S_dev, n, sigma_dev);
  } else {
    OPTMC_CUDA(cudaFuncSetAttribute(ivnet_sigma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ivnet_sigma_kernel<32><<<grid, kLvThreads, smem, ctx->stream>>>(w, net->layers, net->K, net->m_scale, net->tau_scale, net->epsilon, tau, S_dev, n, sigma_dev);
  }
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPTMC_OK;
}

}  // namespace optmc
