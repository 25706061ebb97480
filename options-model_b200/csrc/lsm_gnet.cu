// lsm_gnet.cu -- the reference's v3 algorithm with its own regressor: ONE network for all exercise dates
// (`SingleLSMNet`, om3:85-103: 7 -> 128 -> 128 -> 128 -> 1, ReLU + dropout after every hidden layer), trained by
// mini-batch Adam on the rows of every date at once and then used for the exercise decisions of a second backward
// pass (om3:482-651; torch variant om3gpu:695-833).  SURVEY 8(a) rows a5-a8.
//
//   pass 1  (om3:485-500)   gnet_count_kernel / gnet_scan_kernel / gnet_compact_kernel: the in-the-money rows of every
//                           date t = N-1..1 as a dense table (x = S/K, t, y = payoff(S_N) disc^(N-t)) in a
//                           deterministic order, plus the moments of the seven reference features and of the target
//                           (fp64, order-independent fixed-point atomics)
//   a6      (om3:550-563)   gnet_norm_kernel: feature z-scores (std == 0 -> 1), target z-score
//   a7      (om3:565-613)   per optimiser step gnet_grad_kernel (tcgen05: forward + backward of a 128-row tile per CTA,
//                           the two 128 x 128 layers and all parameter-gradient contractions on the tensor cores,
//                           accumulators in TMEM) + gnet_adam_kernel (fixed-order sum of the per-CTA partial gradients,
//                           Adam / AdamW, batch loss).  Shuffling = a keyed Feistel bijection of the row index per epoch
//                           (no permutation array); dropout = counter-based hash per (step, row, layer, unit).
//                           Host: ReduceLROnPlateau, best-weights snapshot, early stopping -- one scalar read per epoch.
//   pass 2  (om3:615-651)   gnet_walk_kernel: a CTA owns 128 paths and walks the dates backwards, evaluating the network
//                           on the tensor cores only at dates where the tile has a live row; strict '>', sticky mask,
//                           N - 1 discounts (reference semantics) or their textbook counterparts.
//
// Arithmetic: bf16 operands, fp32 accumulation for the 128 x 128 layers (the torch variant runs them in TF32,
// om3gpu:573-574); the input layer (7 -> 128), biases, activations, loss and Adam in fp32.  The reference's
// initialisation / shuffling / dropout streams come from torch's global RNG and are not reproducible bit for bit
// (SURVEY 8c); parity is (i) gradients against torch on identical inputs without dropout, (ii) prices against the
// torch restatement of the same loop within Monte-Carlo / training noise (tests/test_gpu_network.py).
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"
#include "optmc_tc.cuh"

namespace optmc {

constexpr int kGH = 128, kGIn = 7;
constexpr int gW1 = 0;                      // [128][7]   torch Linear layout [out][in]
constexpr int gB1 = gW1 + kGH * kGIn;       // 896
constexpr int gW2 = gB1 + kGH;              // 1024  [128][128]
constexpr int gB2 = gW2 + kGH * kGH;
constexpr int gW3 = gB2 + kGH;              // [128][128]
constexpr int gB3 = gW3 + kGH * kGH;
constexpr int gW4 = gB3 + kGH;              // [128]
constexpr int gB4 = gW4 + kGH;
constexpr int kGP = gB4 + 1;                // 34177 parameters
constexpr int kGThreads = 256;              // tile kernels: thread = (row = tid & 127, column half = tid >> 7)
constexpr int kGChunk = 1024;               // paths per compaction block (256 threads x 4)
constexpr int kGQ = 16;                     // moment sums: 7 x (sum f, sum f^2), sum y, sum y^2

struct GnetNorm {
  float fmean[8], finv[8];
  float ymean, ystd, yinv, pad;
  long long n_rows;
};

__device__ __forceinline__ unsigned int mix32(unsigned int h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}

// ---- pass 1: rows of every date ------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) gnet_count_kernel(const R* __restrict__ S, long long ld, long long M, double K, int is_put,
                                                         unsigned long long* __restrict__ counts, int nchunks) {
  const int t = blockIdx.y + 1;
  const long long base = (long long)blockIdx.x * kGChunk;
  const R* row = S + (size_t)t * ld;
  unsigned int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long j = base + k * 256 + threadIdx.x;
    if (j < M) c += payoff<double>((double)row[j], K, is_put != 0) > 0.0 ? 1u : 0u;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ unsigned int s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < 8; ++w) tot += s[w];
    counts[(size_t)(t - 1) * nchunks + blockIdx.x] = tot;
  }
}

// exclusive scan in place (one CTA): counts -> offsets; total -> *n_rows
__global__ void __launch_bounds__(1024) gnet_scan_kernel(unsigned long long* v, long long n, long long* n_rows) {
  __shared__ unsigned long long part[1024];
  const long long per = (n + 1023) / 1024;
  const long long lo = (long long)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  unsigned long long s = 0;
  for (long long i = lo; i < hi; ++i) s += v[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; ++i) { const unsigned long long x = part[i]; part[i] = run; run += x; }
    *n_rows = (long long)run;
  }
  __syncthreads();
  unsigned long long run = part[threadIdx.x];
  for (long long i = lo; i < hi; ++i) { const unsigned long long x = v[i]; v[i] = run; run += x; }
}

template <typename R>
__global__ void __launch_bounds__(256) gnet_compact_kernel(const R* __restrict__ S, long long ld, long long M, int N, double K,
                                                           int is_put, const unsigned long long* __restrict__ offsets, int nchunks,
                                                           const double* __restrict__ Dt, const double* __restrict__ sqrt_tau,
                                                           float* __restrict__ xs, int* __restrict__ ts, float* __restrict__ ys,
                                                           unsigned long long* __restrict__ sums, int* flags) {
  const int t = blockIdx.y + 1;
  const long long base = (long long)blockIdx.x * kGChunk;
  const R* row = S + (size_t)t * ld;
  const R* last = S + (size_t)N * ld;
  const double invK = 1.0 / K, dg = Dt[t], stau = sqrt_tau[t];
  __shared__ unsigned int wsum[4][8];
  __shared__ double red[8][kGQ];
  double acc[kGQ];
#pragma unroll
  for (int q = 0; q < kGQ; ++q) acc[q] = 0.0;
  unsigned long long off = offsets[(size_t)(t - 1) * nchunks + blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // order inside a chunk: k-major (sub-row k = 256 consecutive paths), then path -- deterministic
  bool live[4];
  double sv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long j = base + k * 256 + threadIdx.x;
    sv[k] = j < M ? (double)row[j] : 0.0;
    live[k] = j < M && payoff<double>(sv[k], K, is_put != 0) > 0.0;
    const unsigned int b = __ballot_sync(0xffffffffu, live[k]);
    if (lane == 0) wsum[k][warp] = __popc(b);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int before = 0;
    for (int kk = 0; kk < k; ++kk)
      for (int w = 0; w < 8; ++w) before += wsum[kk][w];
    for (int w = 0; w < warp; ++w) before += wsum[k][w];
    const unsigned int b = __ballot_sync(0xffffffffu, live[k]);
    if (live[k]) {
      const long long j = base + k * 256 + threadIdx.x;
      const unsigned long long r = off + before + __popc(b & ((1u << lane) - 1u));
      const double x = sv[k] * invK;
      const double y = payoff<double>((double)last[j], K, is_put != 0) * dg;
      xs[r] = (float)x; ts[r] = t; ys[r] = (float)y;
      double f[7];
      f[0] = 1.0; f[1] = x; f[2] = x * x; f[3] = x * x * x; f[4] = x > 1.0 ? x - 1.0 : 0.0; f[5] = stau; f[6] = x * stau;
#pragma unroll
      for (int q = 0; q < 7; ++q) { acc[2 * q] += f[q]; acc[2 * q + 1] += f[q] * f[q]; }
      acc[14] += y; acc[15] += y * y;
    }
  }
#pragma unroll
  for (int q = 0; q < kGQ; ++q) {
    double v = acc[q];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < kGQ) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    if (v != 0.0) {
      unsigned long long hi, lo;
      if (!fx_encode(v, hi, lo)) atomicExch(flags, 1);
      // signed accumulation: subtract the bias so that no contributor count is needed
      atomicAdd(sums + 2 * threadIdx.x, hi - (1ull << 47));
      atomicAdd(sums + 2 * threadIdx.x + 1, lo);
    }
  }
}

__global__ void gnet_norm_kernel(const unsigned long long* __restrict__ sums, const long long* __restrict__ n_rows, int ddof,
                                 GnetNorm* nm) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n = (double)*n_rows;
  nm->n_rows = *n_rows;
  auto val = [&](int q) { return (double)(long long)sums[2 * q] * 0.0625 + (double)sums[2 * q + 1] * 2.220446049250313e-16; };
  for (int k = 0; k < 7; ++k) {
    const double mean = n > 0 ? val(2 * k) / n : 0.0;
    double var = n > 0 ? val(2 * k + 1) / n - mean * mean : 0.0;
    // a constant feature (the intercept; sqrt(tau) when there is one date) has std 0 -> 1 (om3:561)
    const double tol = 1e-13 * (mean * mean > 1.0 ? mean * mean : 1.0);
    double sd = var > tol ? sqrt(var) : 1.0;
    nm->fmean[k] = (float)mean; nm->finv[k] = (float)(1.0 / sd);
  }
  nm->fmean[7] = 0.f; nm->finv[7] = 0.f;
  const double ym = n > 0 ? val(14) / n : 0.0;
  double yv = n > ddof ? (val(15) - n * ym * ym) / (n - ddof) : 0.0;
  double ys = yv > 0.0 ? sqrt(yv) : 0.0;
  if (!(ys > 0.0)) ys = 1.0;  // om3:552-556
  nm->ymean = (float)ym; nm->ystd = (float)ys; nm->yinv = (float)(1.0 / ys); nm->pad = 0.f;
}

// torch.nn.Linear default initialisation: weights and biases U(-1/sqrt(fan_in), 1/sqrt(fan_in))
__global__ void gnet_init_kernel(float* params, float* adam_m, float* adam_v, unsigned long long seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kGP) return;
  const float bound = i < gW2 ? 0.3779644730092272f : 0.08838834764831845f;  // 1/sqrt(7), 1/sqrt(128)
  Philox4 p = philox_for((unsigned long long)i, 0u, 0x474e4554u, seed);
  const float u = (float)((p.v[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  params[i] = (2.0f * u - 1.0f) * bound;
  adam_m[i] = 0.f; adam_v[i] = 0.f;
}

// ---- tile machinery ---------------------------------------------------------------------------------------------
struct GnetSmem {
  unsigned char W2[kTcTileBytes], W3[kTcTileBytes];
  unsigned char A1[kTcTileBytes], A2[kTcTileBytes], A3[kTcTileBytes], A4[kTcTileBytes];  // H1, H2, H3 / D2, D3 / D1
  unsigned char panel[kTcPanelBytes];  // [row][16] = fn0..fn6, 1, dout, 0...
  float W1[kGIn][kGH];                 // transposed: feature-major
  float b1[kGH], b2[kGH], b3[kGH], w4[kGH];
  float b4;
  float red[8];
  float dot[2][128];                   // partial output-layer dot products of the two column halves
  unsigned int mask[3][4][128];  // per layer: 128 "unit is active" bits of each row (word-major: conflict-free)
  unsigned long long bar, wbar;
  unsigned int tmem;
  int any;
};

struct Drop {
  unsigned int thr4;   // threshold byte replicated; 0 = dropout off
  float scale;         // 1 / keep probability
  unsigned int key;    // seed ^ step stream
};

// keep mask of the 8 units [8 c, 8 c + 8) of `layer` for row `r`: bit k set = keep
__device__ __forceinline__ unsigned int drop_keep8(const Drop& d, unsigned int r, int layer, int c) {
  if (d.thr4 == 0u) return 0xffu;
  const unsigned int base = d.key + r * 0x9e3779b1u + (unsigned int)(layer * 16 + c) * 0x7feb352du;
  const unsigned int h0 = mix32(base), h1 = mix32(base ^ 0x68e31da4u);
  const unsigned int k0 = __vcmpgeu4(h0, d.thr4), k1 = __vcmpgeu4(h1, d.thr4);  // 0xff per byte kept
  return ((k0 & 1u) | ((k0 >> 7) & 2u) | ((k0 >> 14) & 4u) | ((k0 >> 21) & 8u)) |
         (((k1 & 1u) | ((k1 >> 7) & 2u) | ((k1 >> 14) & 4u) | ((k1 >> 21) & 8u)) << 4);
}

// bf16 core-layout images of W2 and W3 (2 x 32 KB), kept in step with the fp32 parameters by the optimiser kernel:
// a CTA stages them with two bulk async copies instead of re-packing 32768 floats.
__device__ __forceinline__ int gnet_pack_index(int i) {  // parameter index -> bf16 element index in the images, or -1
  if (i >= gW2 && i < gB2) { const int k = i - gW2; return core_off(k >> 7, k & 127) >> 1; }
  if (i >= gW3 && i < gB3) { const int k = i - gW3; return (kTcTileBytes >> 1) + (core_off(k >> 7, k & 127) >> 1); }
  return -1;
}
__global__ void gnet_pack_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ wpack) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kGP) return;
  const int k = gnet_pack_index(i);
  if (k >= 0) wpack[k] = __float2bfloat16_rn(params[i]);
}

__device__ __forceinline__ unsigned int gnet_setup(GnetSmem& sm, const float* __restrict__ params,
                                                   const __nv_bfloat16* __restrict__ wpack) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(reinterpret_cast<uint64_t*>(&sm.bar), 1);
    mbar_init(reinterpret_cast<uint64_t*>(&sm.wbar), 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(reinterpret_cast<uint64_t*>(&sm.wbar), 2 * kTcTileBytes);
    bulk_load_1d(sm.W2, wpack, kTcTileBytes, reinterpret_cast<uint64_t*>(&sm.wbar));
    bulk_load_1d(sm.W3, wpack + (kTcTileBytes >> 1), kTcTileBytes, reinterpret_cast<uint64_t*>(&sm.wbar));
  }
  if (tid < 128) {
#pragma unroll
    for (int k = 0; k < kGIn; ++k) sm.W1[k][tid] = params[gW1 + tid * kGIn + k];
    sm.b1[tid] = params[gB1 + tid]; sm.b2[tid] = params[gB2 + tid];
    *reinterpret_cast<uint4*>(sm.panel + aux_off(tid, 0)) = make_uint4(0u, 0u, 0u, 0u);
  } else {
    const int u = tid - 128;
    sm.b3[u] = params[gB3 + u]; sm.w4[u] = params[gW4 + u];
    *reinterpret_cast<uint4*>(sm.panel + aux_off(u, 8)) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) sm.b4 = params[gB4];
  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_publish();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return sm.tmem;
}
// the MMA-issuing thread calls this once before its first MMA: the weight images have landed
__device__ __forceinline__ void gnet_weights_ready(GnetSmem& sm) { mbar_wait(reinterpret_cast<uint64_t*>(&sm.wbar), 0u); }

// normalised features of (x, sqrt(tau)) -- om3:105-121 then om3:559-563
__device__ __forceinline__ void gnet_features(float x, float stau, const GnetNorm& nm, float (&fn)[kGIn]) {
  float f[kGIn];
  f[0] = 1.f; f[1] = x; f[2] = x * x; f[3] = x * x * x; f[4] = fmaxf(x - 1.f, 0.f); f[5] = stau; f[6] = x * stau;
#pragma unroll
  for (int k = 0; k < kGIn; ++k) fn[k] = (f[k] - nm.fmean[k]) * nm.finv[k];
}

// The tile kernels run each stage ONCE per CTA, so straight-line code would be fetched cold from the instruction
// cache end to end (ncu: "no instruction" was the top stall); the stage loops below are deliberately NOT unrolled
// and the per-row activity masks live in shared memory (dynamic word index) instead of registers.

// layer 1 on the CUDA cores (fp32): h1 = drop(relu(W1 fn + b1)) -> bf16 tile A1; mask bit = h1 > 0
__device__ __forceinline__ void gnet_layer1(GnetSmem& sm, const float (&fn)[kGIn], int row, int half, unsigned int r, const Drop& d) {
  unsigned int word = 0u;
#pragma unroll 1
  for (int c = half * 8; c < half * 8 + 8; ++c) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sm.b1[c * 8 + k];
#pragma unroll
    for (int i = 0; i < kGIn; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaf(sm.W1[i][c * 8 + k], fn[i], v[k]);
    }
    const unsigned int keep = drop_keep8(d, r, 0, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool on = v[k] > 0.f && ((keep >> k) & 1u);
      v[k] = on ? v[k] * d.scale : 0.f;
      word |= (on ? 1u : 0u) << ((c & 3) * 8 + k);
    }
    *reinterpret_cast<uint4*>(sm.A1 + core_off(row, c * 8)) = pack8_bf16(v);
    if ((c & 3) == 3) { sm.mask[0][c >> 2][row] = word; word = 0u; }
  }
}

// hidden epilogue: h = drop(relu(z + b)) from TMEM columns [col0, col0 + 128) -> bf16 tile; optional dot with w4
template <bool DOT, bool STORE = true>
__device__ __forceinline__ float gnet_hidden(GnetSmem& sm, unsigned int taddr, const float* __restrict__ bias, unsigned char* tile,
                                             int row, int half, unsigned int r, int layer, const Drop& d) {
  float out = 0.f;
#pragma unroll 1
  for (int c0 = half * 2; c0 < half * 2 + 2; ++c0) {
    float z[32];
    tmem_ld32(taddr + c0 * 32, z);
    unsigned int word = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned int keep = drop_keep8(d, r, layer, c0 * 4 + q);
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = c0 * 32 + q * 8 + k;
        const float h = z[q * 8 + k] + bias[j];
        const bool on = h > 0.f && ((keep >> k) & 1u);
        v[k] = on ? h * d.scale : 0.f;
        word |= (on ? 1u : 0u) << (q * 8 + k);
        if (DOT) out = fmaf(sm.w4[j], v[k], out);
      }
      if (STORE) *reinterpret_cast<uint4*>(tile + core_off(row, c0 * 32 + q * 8)) = pack8_bf16(v);
    }
    if (STORE) sm.mask[layer][c0][row] = word;
  }
  return out;
}

// masked copy of a TMEM accumulator: tile = mask ? acc * scale : 0
__device__ __forceinline__ void gnet_masked(GnetSmem& sm, unsigned int taddr, unsigned char* tile, int row, int half, int layer, float scale) {
#pragma unroll 1
  for (int c0 = half * 2; c0 < half * 2 + 2; ++c0) {
    float z[32];
    tmem_ld32(taddr + c0 * 32, z);
    const unsigned int word = sm.mask[layer][c0][row];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ((word >> (q * 8 + k)) & 1u) ? z[q * 8 + k] * scale : 0.f;
      *reinterpret_cast<uint4*>(tile + core_off(row, c0 * 32 + q * 8)) = pack8_bf16(v);
    }
  }
}

// output layer: the two column halves of a row add their partial dot products through shared memory
__device__ __forceinline__ float gnet_join_dot(GnetSmem& sm, float part, int row, int half) {
  sm.dot[half][row] = part;
  __syncthreads();
  return (sm.dot[0][row] + sm.dot[1][row]) + sm.b4;
}

// D[row][n] = sum_k A[row][k] B[n][k]   (A, B K-major tiles)
__device__ __forceinline__ void mma_ab_t(unsigned int d, const unsigned char* A, const unsigned char* B) {
  const unsigned int a0 = smem_u32(A), b0 = smem_u32(B), id = umma_idesc(128, 128, 0, 0);
#pragma unroll 1
  for (int k = 0; k < 8; ++k) umma_f16(d, umma_desc(a0 + k * 256, 128, 2048), umma_desc(b0 + k * 256, 128, 2048), id, k > 0);
}
// D[row][n] = sum_k A[row][k] B[k][n]   (A K-major, B stored [k][n] -> MN-major view)
__device__ __forceinline__ void mma_ab(unsigned int d, const unsigned char* A, const unsigned char* B) {
  const unsigned int a0 = smem_u32(A), b0 = smem_u32(B), id = umma_idesc(128, 128, 0, 1);
#pragma unroll 1
  for (int k = 0; k < 8; ++k) umma_f16(d, umma_desc(a0 + k * 256, 128, 2048), umma_desc(b0 + k * 4096, 2048, 128), id, k > 0);
}
// D[m][n] = sum_row A[row][m] B[row][n]  (both stored [row][.] -> MN-major views)
__device__ __forceinline__ void mma_at_b(unsigned int d, const unsigned char* A, const unsigned char* B) {
  const unsigned int a0 = smem_u32(A), b0 = smem_u32(B), id = umma_idesc(128, 128, 1, 1);
#pragma unroll 1
  for (int k = 0; k < 8; ++k) umma_f16(d, umma_desc(a0 + k * 4096, 2048, 128), umma_desc(b0 + k * 4096, 2048, 128), id, k > 0);
}
// D[m][c] = sum_row A[row][m] P[row][c]  (panel, 16 columns)
__device__ __forceinline__ void mma_at_panel(unsigned int d, const unsigned char* A, const unsigned char* P) {
  const unsigned int a0 = smem_u32(A), p0 = smem_u32(P), id = umma_idesc(128, 16, 1, 1);
#pragma unroll 1
  for (int k = 0; k < 8; ++k) umma_f16(d, umma_desc(a0 + k * 4096, 2048, 128), umma_desc(p0 + k * 512, 256, 128), id, k > 0);
}

// Feistel bijection on [0, 2^(2 half)) with cycle walking onto [0, n): the epoch's shuffle without a permutation array
struct Perm { unsigned long long n; unsigned int half, key; };
__device__ __forceinline__ unsigned long long perm_apply(const Perm& p, unsigned long long i) {
  if (p.half == 0u) return i;
  const unsigned int mask = (1u << p.half) - 1u;
  do {
    unsigned int L = (unsigned int)(i >> p.half) & mask, Rr = (unsigned int)i & mask;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const unsigned int f = mix32(Rr + p.key * (2u * rd + 1u) + 0x9e3779b9u * (rd + 1u)) & mask;
      const unsigned int nl = Rr;
      Rr = L ^ f;
      L = nl;
    }
    i = ((unsigned long long)L << p.half) | Rr;
  } while (i >= p.n);
  return i;
}

struct GradArgs {
  const float* params;
  const __nv_bfloat16* wpack;
  const float* xs; const int* ts; const float* ys;   // row table
  const float* feat;                                  // debug: normalised features [n][7] (then xs/ts unused, ys = targets)
  const double* sqrt_tau;
  const GnetNorm* nm;
  long long start, end;                               // rows [start, end) of the (permuted) order form this batch
  float inv_b2;                                       // 2 / rows of the optimiser step (all ranks' rows when path-sharded)
  Perm perm;
  Drop drop;
  float* gpart;                                       // [tiles][kGP + 1]
};

__global__ void __launch_bounds__(kGThreads, 1) gnet_grad_kernel(const GradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_g[];
  GnetSmem& sm = *reinterpret_cast<GnetSmem*>(smem_g);
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const unsigned int tmem = gnet_setup(sm, a.params, a.wpack);
  const unsigned int lane_base = (unsigned int)((warp & 3) * 32) << 16;  // a warp reads TMEM lanes 32 (warp % 4) ..
  const unsigned int cA = 0, cB = 128, cC = 256, cV1 = 384, cV2 = 400, cV3 = 416, cV4 = 432;
  const long long r = a.start + (long long)blockIdx.x * 128 + row;
  const bool act = r < a.end;
  const float inv_b2 = a.inv_b2;
  unsigned int phase = 0;
  float fn[kGIn], y = 0.f;
#pragma unroll
  for (int k = 0; k < kGIn; ++k) fn[k] = 0.f;
  if (act) {
    if (a.feat) {
#pragma unroll
      for (int k = 0; k < kGIn; ++k) fn[k] = a.feat[(size_t)r * kGIn + k];
      y = a.ys[r];
    } else {
      const unsigned long long src = perm_apply(a.perm, (unsigned long long)r);
      const GnetNorm nm = *a.nm;
      gnet_features(a.xs[src], (float)a.sqrt_tau[a.ts[src]], nm, fn);
      y = (a.ys[src] - nm.ymean) * nm.yinv;
    }
  }
  const unsigned int rr = (unsigned int)r;
  gnet_layer1(sm, fn, row, half, rr, a.drop);
  if (half == 0) {  // panel: fn0..fn6, 1 (inactive rows: all zero, so they add nothing to the gradients)
    float v[8];
#pragma unroll
    for (int k = 0; k < kGIn; ++k) v[k] = fn[k];
    v[7] = act ? 1.f : 0.f;
    *reinterpret_cast<uint4*>(sm.panel + aux_off(row, 0)) = pack8_bf16(v);
  }
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    gnet_weights_ready(sm);
    mma_ab_t(tmem + cA, sm.A1, sm.W2);  // Z2 = H1 W2^T
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  gnet_hidden<false>(sm, tmem + lane_base + cA, sm.b2, sm.A2, row, half, rr, 1, a.drop);
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ab_t(tmem + cB, sm.A2, sm.W3);  // Z3 = H2 W3^T
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  const float part = gnet_hidden<true>(sm, tmem + lane_base + cB, sm.b3, sm.A3, row, half, rr, 2, a.drop);
  const float out = gnet_join_dot(sm, part, row, half);
  const float err = act ? out - y : 0.f;
  const float dout = err * inv_b2;
  // D3 = dout w4 (h3 > 0) scale -> A4; dout -> panel column 8
  {
    const float ds = dout * a.drop.scale;
#pragma unroll 1
    for (int c = half * 8; c < half * 8 + 8; ++c) {
      const unsigned int word = sm.mask[2][c >> 2][row] >> ((c & 3) * 8);
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ((word >> k) & 1u) ? ds * sm.w4[c * 8 + k] : 0.f;
      *reinterpret_cast<uint4*>(sm.A4 + core_off(row, c * 8)) = pack8_bf16(v);
    }
  }
  if (half == 0) *reinterpret_cast<unsigned short*>(sm.panel + aux_off(row, 8)) = __bfloat16_as_ushort(__float2bfloat16_rn(dout));
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ab(tmem + cA, sm.A4, sm.W3);           // dH2 = D3 W3
    mma_at_b(tmem + cC, sm.A2, sm.A4);         // dW3^T = H2^T D3 (lane = input unit: coalesced read-out)
    mma_at_panel(tmem + cV3, sm.A4, sm.panel); // column 7: db3
    mma_at_panel(tmem + cV4, sm.A3, sm.panel); // column 8: dw4 = H3^T dout
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  gnet_masked(sm, tmem + lane_base + cA, sm.A3, row, half, 1, a.drop.scale);  // D2 -> A3 (H3 is done)
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ab(tmem + cA, sm.A3, sm.W2);           // dH1 = D2 W2
    mma_at_b(tmem + cB, sm.A1, sm.A3);         // dW2^T = H1^T D2
    mma_at_panel(tmem + cV2, sm.A3, sm.panel); // column 7: db2
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  gnet_masked(sm, tmem + lane_base + cA, sm.A4, row, half, 0, a.drop.scale);  // D1 -> A4 (D3 is done)
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_at_panel(tmem + cV1, sm.A4, sm.panel); // columns 0..6: dW1, column 7: db1
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  // read-out.  dW2 / dW3 sit transposed in TMEM (lane = input unit i, column = output unit j): for a fixed j the
  // threads of a warp write consecutive addresses of row j.  The vector gradients have lane = unit.
  float* gp = a.gpart + (size_t)blockIdx.x * (kGP + 1);
#pragma unroll 1
  for (int c0 = half * 2; c0 < half * 2 + 2; ++c0) {
    float w[32];
    tmem_ld32(tmem + lane_base + cB + c0 * 32, w);
#pragma unroll
    for (int i = 0; i < 32; ++i) gp[gW2 + (c0 * 32 + i) * kGH + row] = w[i];
    tmem_ld32(tmem + lane_base + cC + c0 * 32, w);
#pragma unroll
    for (int i = 0; i < 32; ++i) gp[gW3 + (c0 * 32 + i) * kGH + row] = w[i];
  }
  if (half == 0) {
    float v1[16], v2[16];
    tmem_ld16(tmem + lane_base + cV1, v1);
    tmem_ld16(tmem + lane_base + cV2, v2);
#pragma unroll
    for (int k = 0; k < kGIn; ++k) gp[gW1 + row * kGIn + k] = v1[k];
    gp[gB1 + row] = v1[7]; gp[gB2 + row] = v2[7];
  } else {
    float v3[16], v4[16];
    tmem_ld16(tmem + lane_base + cV3, v3);
    tmem_ld16(tmem + lane_base + cV4, v4);
    gp[gB3 + row] = v3[7]; gp[gW4 + row] = v4[8];
  }
  float gb4 = half == 0 ? dout : 0.f, loss = half == 0 ? err * err : 0.f;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { gb4 += __shfl_xor_sync(0xffffffffu, gb4, m); loss += __shfl_xor_sync(0xffffffffu, loss, m); }
  if (half == 0 && (tid & 31) == 0) { sm.red[warp] = gb4; sm.red[4 + warp] = loss; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    gp[gB4] = (sm.red[0] + sm.red[1]) + (sm.red[2] + sm.red[3]);
    gp[kGP] = (sm.red[4] + sm.red[5]) + (sm.red[6] + sm.red[7]);  // sum of squared errors of the tile
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ---- path-sharded training: gradients and small vectors travel through peer-mapped memory (optmc_comm_*) -----------
// Every 64-bit word is {tag32, payload32} (single-copy atomic, self-validating: no flag, no fence).  A rank PUSHES its
// words into slot [tag & 1][rank] of every rank's region with st.relaxed.sys; readers poll their OWN memory until the
// tags match.  A slot is rewritten two exchanges later, which stream order puts behind every reader: rank r pushes
// exchange s + 2 after its kernel of exchange s + 1 has seen rank q's words of s + 1, which q pushed after its reader of
// exchange s had finished.
struct GnetPeers {
  unsigned long long* base[kCommMaxRanks];  // each rank's region (behind the sweep's exchange slots)
  int rank, nranks;
};
constexpr unsigned int kGnetSpinLimit = 1u << 23;  // ~seconds: a peer that never launched must not hang the GPU
__device__ __forceinline__ unsigned long long* gpeer_grad(unsigned long long* base, unsigned int par, int src) {
  return base + ((size_t)par * kCommMaxRanks + src) * kGnetPad;
}
__device__ __forceinline__ unsigned long long* gpeer_meta(unsigned long long* base, unsigned int par, int src) {
  return base + (size_t)2 * kCommMaxRanks * kGnetPad + ((size_t)par * kCommMaxRanks + src) * kGnetMetaWords;
}
// polls slot [par][r][i] of this rank's region for every rank r; returns false after a time-out (flags[1] set: every
// later poll of this and the following kernels gives up at once, the host reports OPTMC_ECUDA)
__device__ __forceinline__ bool gpeer_poll(const unsigned long long* mine, size_t stride, int nranks, unsigned int tag,
                                           unsigned int (&pay)[kCommMaxRanks], int* flags) {
  bool dead = *reinterpret_cast<volatile int*>(flags + 1) != 0;
  unsigned long long v[kCommMaxRanks];
  unsigned int n = 0;
  for (;;) {
#pragma unroll
    for (int r = 0; r < kCommMaxRanks; ++r) v[r] = r < nranks ? ld_relaxed_sys_u64(mine + (size_t)r * stride) : (unsigned long long)tag << 32;
    bool all = true;
#pragma unroll
    for (int r = 0; r < kCommMaxRanks; ++r) all &= (unsigned int)(v[r] >> 32) == tag;
    if (all || dead) break;
    if (++n >= kGnetSpinLimit) { dead = true; atomicExch(flags + 1, 1); }
  }
#pragma unroll
  for (int r = 0; r < kCommMaxRanks; ++r) pay[r] = (unsigned int)v[r];
  return !dead;
}

// all-gather + integer sum of n64 <= 64 64-bit words (row counts, fixed-point moments, the final value sums):
// gath[r][j] = rank r's word j, tot[j] = their wrap-around sum.  One CTA.
__global__ void __launch_bounds__(128) gnet_gather_kernel(const unsigned long long* __restrict__ src, int n64, GnetPeers pe, unsigned int tag,
                                                          unsigned long long* __restrict__ gath, unsigned long long* __restrict__ tot, int* flags) {
  __shared__ unsigned int half[kCommMaxRanks][128];
  const int i = threadIdx.x;
  if (i < 2 * n64) {
    const unsigned int mine = (unsigned int)(src[i >> 1] >> ((i & 1) * 32));
    const unsigned long long w = ((unsigned long long)tag << 32) | mine;
    for (int p = 0; p < pe.nranks; ++p) st_relaxed_sys_u64(gpeer_meta(pe.base[p], tag & 1u, pe.rank) + i, w);
    unsigned int pay[kCommMaxRanks];
    gpeer_poll(gpeer_meta(pe.base[pe.rank], tag & 1u, 0) + i, kGnetMetaWords, pe.nranks, tag, pay, flags);
#pragma unroll
    for (int r = 0; r < kCommMaxRanks; ++r) half[r][i] = pay[r];
  }
  __syncthreads();
  if (i < n64) {
    unsigned long long t = 0ull;
    for (int r = 0; r < pe.nranks; ++r) {
      const unsigned long long v = (unsigned long long)half[r][2 * i] | ((unsigned long long)half[r][2 * i + 1] << 32);
      gath[(size_t)r * n64 + i] = v;
      t += v;
    }
    tot[i] = t;
  }
}

// Adam (torch.optim.Adam with L2 weight decay, om3:579) or AdamW (decoupled, om3gpu:753): fixed-order sum of the
// per-tile partial gradients (PEER: this rank's sum is pushed to every rank, then the ranks' vectors are added in rank
// order -- identical on every rank); element kGP carries the step's squared error into the epoch accumulator.
template <bool PEER>
__global__ void __launch_bounds__(256) gnet_adam_kernel(float* params, __nv_bfloat16* wpack, float* adam_m, float* adam_v, const float* __restrict__ gpart,
                                                        int ntiles, float lr, float wd, int decoupled, int step, float inv_batch,
                                                        double* epoch_loss, GnetPeers pe, unsigned int tag, int* flags) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= kGP) {
    float g = 0.f;
    if (PEER) {
      // this rank's vector: fixed-order sum of its tiles, pushed into slot [tag & 1][rank] of EVERY rank (its own
      // included) before anything is polled -- the peers' optimiser kernels do the same concurrently
      for (int b = 0; b < ntiles; ++b) g += gpart[(size_t)b * (kGP + 1) + i];
      const unsigned long long w = ((unsigned long long)tag << 32) | __float_as_uint(g);
      for (int p = 0; p < pe.nranks; ++p) st_relaxed_sys_u64(gpeer_grad(pe.base[p], tag & 1u, pe.rank) + i, w);
      unsigned int pay[kCommMaxRanks];
      gpeer_poll(gpeer_grad(pe.base[pe.rank], tag & 1u, 0) + i, kGnetPad, pe.nranks, tag, pay, flags);
      g = __uint_as_float(pay[0]);
#pragma unroll
      for (int r = 1; r < kCommMaxRanks; ++r) if (r < pe.nranks) g += __uint_as_float(pay[r]);
    } else if (i < kGP || epoch_loss) {
      for (int b = 0; b < ntiles; ++b) g += gpart[(size_t)b * (kGP + 1) + i];
    }
    if (i < kGP) {
      float p = params[i];
      if (decoupled) p -= lr * wd * p; else g = fmaf(wd, p, g);
      const float m = b1 * adam_m[i] + (1.0f - b1) * g;
      const float v = b2 * adam_v[i] + (1.0f - b2) * g * g;
      adam_m[i] = m; adam_v[i] = v;
      p -= (lr / bc1) * (m / (sqrtf(v) / sqrtf(bc2) + eps));
      params[i] = p;
      const int k = gnet_pack_index(i);
      if (k >= 0) wpack[k] = __float2bfloat16_rn(p);
    } else if (epoch_loss) {
      *epoch_loss += (double)(g * inv_batch);  // element kGP: the step's sum of squared errors
    }
  }
}

__global__ void gnet_sum_partials_kernel(const float* __restrict__ gpart, int ntiles, float* out, float inv_batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > kGP) return;
  float g = 0.f;
  for (int b = 0; b < ntiles; ++b) g += gpart[(size_t)b * (kGP + 1) + i];
  out[i] = i == kGP ? g * inv_batch : g;
}

// ---- pass 2 -------------------------------------------------------------------------------------------------------
struct WalkGArgs {
  const void* S; long long ld, M; int N, is_put, sticky;
  double K, invK;
  const float* params; const __nv_bfloat16* wpack; const GnetNorm* nm;
  const double* sqrt_tau; const double* Dm;     // Dm[t] = disc^(t-1): value today of a payoff taken at date t (N-1 convention)
  Drop drop;                                     // inference dropout (reference: the net is never put in eval mode)
  unsigned long long* fin;                       // fixed-point sums: value, value^2 (signed, bias-free)
  unsigned long long* exc; unsigned long long* bnd;
  int* flags;
};

template <typename R>
__global__ void __launch_bounds__(kGThreads, 1) gnet_walk_kernel(const WalkGArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_g[];
  GnetSmem& sm = *reinterpret_cast<GnetSmem*>(smem_g);
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const unsigned int tmem = gnet_setup(sm, a.params, a.wpack);
  if (tid == 0) gnet_weights_ready(sm);
  const unsigned int lane_base = (unsigned int)((warp & 3) * 32) << 16;
  const long long j = (long long)blockIdx.x * 128 + row;
  const bool act = j < a.M;
  const R* Sp = static_cast<const R*>(a.S);
  const GnetNorm nm = *a.nm;
  const bool put = a.is_put != 0;
  double value = act ? payoff<double>((double)Sp[(size_t)a.N * a.ld + j], a.K, put) * a.Dm[a.N] : 0.0;
  bool exercised = false;
  unsigned int phase = 0;
  for (int t = a.N - 1; t >= 1; --t) {
    const double s = act ? (double)Sp[(size_t)t * a.ld + j] : 0.0;
    const double pay = act ? payoff<double>(s, a.K, put) : 0.0;
    const bool live = act && pay > 0.0 && !(a.sticky && exercised);
    if (!__syncthreads_or(live)) continue;  // nothing to decide for these 128 paths at this date
    float fn[kGIn];
    gnet_features((float)(s * a.invK), (float)a.sqrt_tau[t], nm, fn);
    const unsigned int rr = (unsigned int)j * 0x01000193u + (unsigned int)t;
    gnet_layer1(sm, fn, row, half, rr, a.drop);
    tc_publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mma_ab_t(tmem, sm.A1, sm.W2);
      umma_commit(&sm.bar);
    }
    tc_bar_wait(&sm.bar, phase); phase ^= 1u;
    gnet_hidden<false>(sm, tmem + lane_base, sm.b2, sm.A2, row, half, rr, 1, a.drop);
    tc_publish();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mma_ab_t(tmem + 128, sm.A2, sm.W3);
      umma_commit(&sm.bar);
    }
    tc_bar_wait(&sm.bar, phase); phase ^= 1u;
    const float out = gnet_join_dot(sm, gnet_hidden<true, false>(sm, tmem + lane_base + 128, sm.b3, sm.A3, row, half, rr, 2, a.drop), row, half);
    const double cont = (double)(out * nm.ystd + nm.ymean);  // om3:640
    const bool ex = live && pay > cont;                       // strict, om3:644
    if (ex) { value = pay * a.Dm[t]; exercised = true; }
    if (a.exc && half == 0) {
      const unsigned int cnt = __reduce_add_sync(0xffffffffu, ex ? 1u : 0u);
      if (cnt) {
        unsigned long long b = ex ? (unsigned long long)__double_as_longlong(s) : bnd_none(a.is_put);
        b = put ? warp_max_u64(b) : warp_min_u64(b);
        if ((tid & 31) == 0) {
          atomicAdd(a.exc + t, (unsigned long long)cnt);
          if (put) atomicMax(a.bnd + t, b); else atomicMin(a.bnd + t, b);
        }
      }
    }
  }
  double f0 = value, f1 = value * value;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { f0 += __shfl_xor_sync(0xffffffffu, f0, m); f1 += __shfl_xor_sync(0xffffffffu, f1, m); }
  __shared__ double fred[2][4];
  if (half == 0 && (tid & 31) == 0) { fred[0][warp] = f0; fred[1][warp] = f1; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 2) {
    const double v = (fred[tid][0] + fred[tid][1]) + (fred[tid][2] + fred[tid][3]);
    unsigned long long hi, lo;
    if (!fx_encode(v, hi, lo)) atomicExch(a.flags, 1);
    atomicAdd(a.fin + 2 * tid, hi - (1ull << 47));
    atomicAdd(a.fin + 2 * tid + 1, lo);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

__global__ void gnet_final_kernel(const unsigned long long* fin, long long M, double scale, double* out) {
  auto val = [&](int q) { return (double)(long long)fin[2 * q] * 0.0625 + (double)fin[2 * q + 1] * 2.220446049250313e-16; };
  const double n = (double)M, s1 = val(0), s2 = val(1);
  const double mean = s1 / n;
  double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
  if (var < 0.0) var = 0.0;
  out[0] = mean * scale; out[1] = sqrt(var / n) * scale; out[2] = s1; out[3] = s2;
}

// ---- host driver --------------------------------------------------------------------------------------------------
static Drop make_drop(double p, unsigned int key) {
  Drop d{};
  int thr = (int)lrint(p * 256.0);
  if (thr < 0) thr = 0;
  if (thr > 255) thr = 255;
  d.thr4 = (unsigned int)thr * 0x01010101u;
  d.scale = thr ? 256.0f / (float)(256 - thr) : 1.0f;  // realised keep probability (256 - thr) / 256
  d.key = key;
  return d;
}

static Perm make_perm(unsigned long long n, unsigned int key) {
  Perm p{};
  p.n = n; p.key = key;
  unsigned int bits = 1;
  while ((1ull << bits) < n) ++bits;
  p.half = (bits + 1) / 2;
  if (n <= 1) p.half = 0;
  return p;
}

static size_t gnet_smem_bytes() { return sizeof(GnetSmem) + 1024; }

// Step b of an epoch covers positions [b batch, (b+1) batch) of the global order of all ranks' rows; a rank's share is the
// proportional slice of its own order (floor arithmetic: the slices of consecutive steps tile [0, n_rank) exactly).
static long long shard_pos(long long pos, long long n_rank, long long n_total) {
  return n_total > 0 ? (long long)(((unsigned __int128)pos * (unsigned __int128)n_rank) / (unsigned __int128)n_total) : 0;
}
int gnet_shard_plan(const int64_t* n_rows, int32_t nranks, int32_t batch, int64_t b, int32_t rank, int64_t* lo, int64_t* hi,
                    int64_t* global_rows) {
  if (!n_rows || nranks < 1 || nranks > kCommMaxRanks || batch < 1 || b < 0 || rank < 0 || rank >= nranks) { set_error("bad argument"); return OPTMC_EINVAL; }
  long long n_total = 0;
  for (int r = 0; r < nranks; ++r) { if (n_rows[r] < 0) { set_error("bad argument"); return OPTMC_EINVAL; } n_total += n_rows[r]; }
  const long long p0 = b * (long long)batch < n_total ? b * (long long)batch : n_total;
  const long long p1 = p0 + batch < n_total ? p0 + batch : n_total;
  long long g = 0;
  for (int r = 0; r < nranks; ++r) {
    const long long l = shard_pos(p0, n_rows[r], n_total), h = shard_pos(p1, n_rows[r], n_total);
    g += h - l;
    if (r == rank) { if (lo) *lo = l; if (hi) *hi = h; }
  }
  if (global_rows) *global_rows = g;
  return OPTMC_OK;
}

// End of an epoch without a host round trip (no scheduler, no early stopping: nothing the host must decide): the
// best-weights snapshot of om3:599-603 on the device.  st[2][4] = {best, have_best, bad (non-finite loss seen), -},
// double-buffered by epoch parity so that every thread tests the same `best`.
__global__ void __launch_bounds__(256) gnet_epoch_end_kernel(const double* __restrict__ loss, double nb, double min_delta, double* st, int ep,
                                                             const float* __restrict__ params, float* __restrict__ best) {
  const double* cur = st + (ep & 1) * 4;
  double* nxt = st + ((ep + 1) & 1) * 4;
  const double avg = *loss / nb;
  const bool better = avg < cur[0] - min_delta;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (better && i < kGP) best[i] = params[i];
  if (i == 0) {
    nxt[0] = better ? avg : cur[0];
    nxt[1] = better ? 1.0 : cur[1];
    nxt[2] = (avg == avg) ? cur[2] : 1.0;
  }
}

// a7: mini-batch training (om3:565-613) of the network in io.params on the rows (io.xs, io.ts, io.ys); the caller has
// initialised io.params / io.m / io.v / io.pack and io.norm.  Used by the global fit (all dates' rows, optionally the
// ranks' shares of a path-sharded fit) and by the per-date fit (one date's rows).
struct GnetTrainIo {
  float *params, *best, *m, *v, *gpart;
  double* loss;                    // [1] epoch accumulator + [8] device epoch state (gnet_epoch_end_kernel)
  __nv_bfloat16* pack;
  const GnetNorm* norm;
  const double* sqrt_tau;
  const float* xs; const int* ts; const float* ys;
  long long n_rows, n_total;       // this rank's rows, all ranks' rows
  const long long* n_rank;         // [nranks]
  bool sharded;
  GnetPeers pe;
  int batch;
  unsigned long long seed;         // shuffle / dropout streams
  // results
  double lr, best_loss;
  int epochs_run, n_launches;
  bool have_best;
};

static int gnet_train(optmc_ctx* ctx, const optmc_gnet_params* gp, GnetTrainIo& io) {
  const int batch = io.batch;
  double lr = gp->lr, best = INFINITY;
  int step = 0, since_best = 0, sched_bad = 0, epochs_run = 0;
  double sched_best = INFINITY;
  bool have_best = false;
  if (io.n_total > 0) {
    const long long nb = (io.n_total + batch - 1) / batch;
    const unsigned int rank_key = io.sharded ? (unsigned int)io.pe.rank * 0x3c6ef372u : 0u;  // the ranks shuffle / mask independently
    // the host looks at every epoch's loss only when it has something to decide (scheduler, early stopping)
    const bool host_epochs = gp->sched_patience > 0 || gp->stop_patience > 0;
    double* d_state = io.loss + 1;
    if (!host_epochs) {
      const double st0[8] = {INFINITY, 0.0, 0.0, 0.0, INFINITY, 0.0, 0.0, 0.0};
      OPTMC_CUDA(cudaMemcpyAsync(d_state, st0, sizeof(st0), cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int ep = 0; ep < gp->epochs; ++ep) {
      OPTMC_CUDA(cudaMemsetAsync(io.loss, 0, 8, ctx->stream));
      GradArgs ga{};
      ga.params = io.params; ga.wpack = io.pack; ga.xs = io.xs; ga.ts = io.ts; ga.ys = io.ys; ga.feat = nullptr; ga.sqrt_tau = io.sqrt_tau; ga.nm = io.norm;
      ga.perm = make_perm((unsigned long long)io.n_rows, (unsigned int)(io.seed * 0x9e3779b97f4a7c15ull >> 32) + 0x632be5abu * (unsigned int)(ep + 1) + rank_key);
      ga.gpart = io.gpart;
      for (long long b = 0; b < nb; ++b) {
        long long step_rows;  // rows of all ranks in this step: the gradient's normaliser
        if (io.sharded) {
          const long long p0 = b * batch, p1 = p0 + batch < io.n_total ? p0 + batch : io.n_total;
          ga.start = shard_pos(p0, io.n_rows, io.n_total); ga.end = shard_pos(p1, io.n_rows, io.n_total);
          step_rows = 0;
          for (int r = 0; r < io.pe.nranks; ++r) step_rows += shard_pos(p1, io.n_rank[r], io.n_total) - shard_pos(p0, io.n_rank[r], io.n_total);
        } else {
          ga.start = b * batch;
          ga.end = ga.start + batch < io.n_rows ? ga.start + batch : io.n_rows;
          step_rows = ga.end - ga.start;
        }
        ga.inv_b2 = 2.0f / (float)step_rows;
        ++step;
        ga.drop = make_drop(gp->dropout, (unsigned int)io.seed * 0x2545f491u + (unsigned int)step * 0x9e3779b1u + rank_key);
        const int tiles = (int)((ga.end - ga.start + 127) / 128);
        if (tiles > 0) { gnet_grad_kernel<<<tiles, kGThreads, gnet_smem_bytes(), ctx->stream>>>(ga); ++io.n_launches; ctx->launches++; }
        if (io.sharded) {
          const unsigned int tag = ctx->comm.gn_step++;
          gnet_adam_kernel<true><<<(kGP + 256) / 256, 256, 0, ctx->stream>>>(io.params, io.pack, io.m, io.v, io.gpart, tiles, (float)lr, (float)gp->weight_decay,
                                                                             gp->decoupled_wd, step, 1.0f / (float)step_rows, io.loss, io.pe, tag, ctx->d_flags);
          ++io.n_launches; ctx->launches++;
        } else {
          gnet_adam_kernel<false><<<(kGP + 256) / 256, 256, 0, ctx->stream>>>(io.params, io.pack, io.m, io.v, io.gpart, tiles, (float)lr, (float)gp->weight_decay,
                                                                              gp->decoupled_wd, step, 1.0f / (float)step_rows, io.loss, io.pe, 0u, ctx->d_flags);
          ++io.n_launches; ctx->launches++;
        }
      }
      OPTMC_CUDA(cudaGetLastError());
      if (!host_epochs) {
        gnet_epoch_end_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(io.loss, (double)nb, gp->min_delta, d_state, ep, io.params, io.best);
        ++io.n_launches; ctx->launches++;
        ++epochs_run;
        continue;
      }
      double sum_loss = 0.0;
      int hf[4] = {0, 0, 0, 0};
      OPTMC_CUDA(cudaMemcpyAsync(&sum_loss, io.loss, 8, cudaMemcpyDeviceToHost, ctx->stream));
      if (io.sharded) OPTMC_CUDA(cudaMemcpyAsync(hf, ctx->d_flags, sizeof(hf), cudaMemcpyDeviceToHost, ctx->stream));
      OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
      if (hf[1]) { set_error("sharded network LSM: a peer rank did not answer (gradient exchange timed out); re-run optmc_comm_export / optmc_comm_init on every rank"); return OPTMC_ECUDA; }
      const double avg = sum_loss / (double)nb;
      ++epochs_run;
      if (!(avg == avg)) { set_error("network LSM: the training loss is not finite"); return OPTMC_ECUDA; }
      // ReduceLROnPlateau (mode min, rel threshold 1e-4, om3:580): patience epochs without improvement -> lr *= factor
      if (gp->sched_patience > 0) {
        if (avg < sched_best * (1.0 - 1e-4)) { sched_best = avg; sched_bad = 0; }
        else if (++sched_bad > gp->sched_patience) {
          const double nl = lr * gp->sched_factor > gp->min_lr ? lr * gp->sched_factor : gp->min_lr;
          lr = nl; sched_bad = 0;
        }
      }
      if (avg < best - gp->min_delta) {  // om3:599-603
        best = avg; since_best = 0; have_best = true;
        OPTMC_CUDA(cudaMemcpyAsync(io.best, io.params, (size_t)kGP * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      } else if (gp->stop_patience > 0 && ++since_best >= gp->stop_patience) {
        break;
      }
    }
    if (!host_epochs && gp->epochs > 0) {
      double hs[4];
      int hf[4] = {0, 0, 0, 0};
      OPTMC_CUDA(cudaMemcpyAsync(hs, d_state + (gp->epochs & 1) * 4, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
      if (io.sharded) OPTMC_CUDA(cudaMemcpyAsync(hf, ctx->d_flags, sizeof(hf), cudaMemcpyDeviceToHost, ctx->stream));
      OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
      if (hf[1]) { set_error("sharded network LSM: a peer rank did not answer (gradient exchange timed out); re-run optmc_comm_export / optmc_comm_init on every rank"); return OPTMC_ECUDA; }
      if (hs[2] != 0.0) { set_error("network LSM: the training loss is not finite"); return OPTMC_ECUDA; }
      best = hs[0]; have_best = hs[1] != 0.0;
    }
    if (have_best) {  // om3:611-613
      OPTMC_CUDA(cudaMemcpyAsync(io.params, io.best, (size_t)kGP * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      gnet_pack_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(io.params, io.pack);
      ++io.n_launches; ctx->launches++;
    }
  }
  io.lr = lr; io.best_loss = best; io.epochs_run = epochs_run; io.have_best = have_best;
  return OPTMC_OK;
}

template <typename R>
static int lsm_gnet_t(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int64_t M_total, int32_t N, const optmc_lsm_params* lp,
                      const optmc_gnet_params* gp, optmc_gnet_result* out) {
  const bool sharded = M_total > 0;
  GnetPeers pe{};
  if (sharded) {
    pe.rank = ctx->comm.rank; pe.nranks = ctx->comm.nranks;
    for (int r = 0; r < pe.nranks; ++r) pe.base[r] = ctx->comm.peers[r] + kCommSweepWords;
  }
  const bool sticky = (lp->semantics & OPTMC_SEM_STICKY_MASK) != 0;
  const bool refdisc = (lp->semantics & OPTMC_SEM_REF_DISCOUNT) != 0;
  const double dt = lp->T / N, disc = exp(-lp->r * dt);
  int rc = ensure_per_date(ctx, N);
  if (rc) return rc;
  ctx->sw = SweepDesc{};
  ctx->sw.N = N; ctx->sw.lp = *lp;
  rc = sweep_reset_stats(ctx);
  if (rc) return rc;
  int n_launches = 1;
  // host tables: Dt[t] = disc^(N-t) (pass-1 targets), sqrt(tau_t), Dm[t] = disc^(t-1) (pass-2 values)
  std::vector<double> tab((size_t)3 * (N + 1));
  double* Dt = tab.data(); double* st = Dt + (N + 1); double* Dm = st + (N + 1);
  {
    double d = 1.0;
    for (int t = N; t >= 0; --t) { Dt[t] = d; d *= disc; }
    for (int t = 0; t <= N; ++t) { const double tau = lp->T - t * dt; st[t] = sqrt(tau > 1e-6 ? tau : 1e-6); }
    d = 1.0;
    Dm[0] = 1.0 / disc;
    for (int t = 1; t <= N; ++t) { Dm[t] = d; d *= disc; }
  }
  const int nchunks = (int)((M + kGChunk - 1) / kGChunk);
  const long long ncounts = (long long)(N - 1) * nchunks;
  const int batch = gp->batch > 0 ? gp->batch : 256;
  const int max_tiles = (batch + 127) / 128;
  // device workspace (everything but the row table)
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  // meta block: [0] row count, [1, 33) fixed-point moments, [33, 37) fixed-point value sums -- contiguous, so that a
  // path-sharded run gathers them as one vector; tot / gath: the ranks' sums and the ranks' own words
  constexpr int kMetaN = 1 + 2 * kGQ + 4;
  const size_t o_tab = take(tab.size() * 8), o_counts = take((size_t)(ncounts > 0 ? ncounts : 1) * 8), o_meta = take(kMetaN * 8),
               o_tot = take(kMetaN * 8), o_gath = take((size_t)kCommMaxRanks * kMetaN * 8), o_norm = take(sizeof(GnetNorm)), o_params = take((size_t)kGP * 4), o_best = take((size_t)kGP * 4),
               o_m = take((size_t)kGP * 4), o_v = take((size_t)kGP * 4), o_gpart = take((size_t)max_tiles * (kGP + 1) * 4),
               o_loss = take(9 * 8), o_pack = take(2 * kTcTileBytes);
  rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  double* d_tab = reinterpret_cast<double*>(dev + o_tab);
  unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(dev + o_counts);
  unsigned long long* d_meta = reinterpret_cast<unsigned long long*>(dev + o_meta);
  unsigned long long* d_tot = reinterpret_cast<unsigned long long*>(dev + o_tot);
  unsigned long long* d_gath = reinterpret_cast<unsigned long long*>(dev + o_gath);
  long long* d_nrows = reinterpret_cast<long long*>(d_meta);
  unsigned long long* d_sums = d_meta + 1;
  unsigned long long* d_fin = d_sums + 2 * kGQ;
  GnetNorm* d_norm = reinterpret_cast<GnetNorm*>(dev + o_norm);
  float* d_params = reinterpret_cast<float*>(dev + o_params);
  float* d_best = reinterpret_cast<float*>(dev + o_best);
  float* d_m = reinterpret_cast<float*>(dev + o_m);
  float* d_v = reinterpret_cast<float*>(dev + o_v);
  float* d_gpart = reinterpret_cast<float*>(dev + o_gpart);
  double* d_loss = reinterpret_cast<double*>(dev + o_loss);
  __nv_bfloat16* d_pack = reinterpret_cast<__nv_bfloat16*>(dev + o_pack);
  OPTMC_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemsetAsync(d_meta, 0, kMetaN * 8, ctx->stream));
  const R* Sr = static_cast<const R*>(S);
  OPTMC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));

  long long n_rows = 0;
  if (N >= 2) {
    gnet_count_kernel<R><<<dim3(nchunks, N - 1), 256, 0, ctx->stream>>>(Sr, ld, M, lp->K, lp->is_put, d_counts, nchunks);
    gnet_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, ncounts, d_nrows);
    n_launches += 2; ctx->launches += 2;
    OPTMC_CUDA(cudaGetLastError());
    OPTMC_CUDA(cudaMemcpyAsync(&n_rows, d_nrows, 8, cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  out->n_rows = n_rows; out->epochs_run = 0; out->best_loss = nan(""); out->final_lr = gp->lr;
  const double final_scale = refdisc ? 1.0 : disc;
  // n_rows == 0 (om3:517-518: no regression rows -> mean of the discounted terminal payoffs): pass 2 runs with a
  // network that is never consulted
  float *d_xs = nullptr, *d_ys = nullptr;
  int* d_ts = nullptr;
  if (n_rows > 0) {
    rc = ensure_bytes(&ctx->gnet_rows, &ctx->gnet_rows_cap, (size_t)n_rows * 12 + 768);
    if (rc) return rc;
    char* rows = static_cast<char*>(ctx->gnet_rows);
    const size_t seg = ((size_t)n_rows * 4 + 255) / 256 * 256;
    d_xs = reinterpret_cast<float*>(rows); d_ts = reinterpret_cast<int*>(rows + seg); d_ys = reinterpret_cast<float*>(rows + 2 * seg);
    gnet_compact_kernel<R><<<dim3(nchunks, N - 1), 256, 0, ctx->stream>>>(Sr, ld, M, N, lp->K, lp->is_put, d_counts, nchunks, d_tab,
                                                                         d_tab + (N + 1), d_xs, d_ts, d_ys, d_sums, ctx->d_flags);
    ++n_launches; ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
  }
  // rows of every rank (sharded: gathered with the moments through peer memory; every rank launches this, rows or not)
  long long n_rank[kCommMaxRanks] = {n_rows};
  long long n_total = n_rows;
  const unsigned long long* d_sums_all = d_sums;
  const long long* d_nrows_all = d_nrows;
  if (sharded) {
    gnet_gather_kernel<<<1, 128, 0, ctx->stream>>>(d_meta, 1 + 2 * kGQ, pe, ctx->comm.gn_meta++, d_gath, d_tot, ctx->d_flags);
    ++n_launches; ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
    unsigned long long hg[kCommMaxRanks * (1 + 2 * kGQ)];
    int hf[4];
    OPTMC_CUDA(cudaMemcpyAsync(hg, d_gath, sizeof(hg), cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaMemcpyAsync(hf, ctx->d_flags, sizeof(hf), cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hf[1]) { set_error("sharded network LSM: a peer rank did not answer (exchange timed out); re-run optmc_comm_export / optmc_comm_init on every rank"); return OPTMC_ECUDA; }
    n_total = 0;
    for (int r = 0; r < pe.nranks; ++r) { n_rank[r] = (long long)hg[(size_t)r * (1 + 2 * kGQ)]; n_total += n_rank[r]; }
    d_sums_all = d_tot + 1; d_nrows_all = reinterpret_cast<const long long*>(d_tot);
    out->n_rows = n_total;
  }
  if (n_total > 0) {
    gnet_norm_kernel<<<1, 32, 0, ctx->stream>>>(d_sums_all, d_nrows_all, gp->target_ddof ? 1 : 0, d_norm);
    ++n_launches; ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
  } else {
    GnetNorm z{};
    z.ystd = 1.f; z.yinv = 1.f;
    for (int k = 0; k < 8; ++k) z.finv[k] = 1.f;
    OPTMC_CUDA(cudaMemcpyAsync(d_norm, &z, sizeof(z), cudaMemcpyHostToDevice, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  gnet_init_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(d_params, d_m, d_v, gp->seed);
  if (gp->init_params)  // warm start (om3gpu:741-748); the optimiser state is fresh either way (om3gpu:753)
    OPTMC_CUDA(cudaMemcpyAsync(d_params, gp->init_params, (size_t)kGP * 4, cudaMemcpyHostToDevice, ctx->stream));
  gnet_pack_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(d_params, d_pack);
  n_launches += 2; ctx->launches += 2;
  OPTMC_CUDA(cudaFuncSetAttribute(gnet_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnet_smem_bytes()));
  OPTMC_CUDA(cudaFuncSetAttribute(gnet_walk_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnet_smem_bytes()));

  // ---- a7: mini-batch training (om3:565-613) ----
  GnetTrainIo io{};
  io.params = d_params; io.best = d_best; io.m = d_m; io.v = d_v; io.gpart = d_gpart; io.loss = d_loss; io.pack = d_pack; io.norm = d_norm;
  io.sqrt_tau = d_tab + (N + 1); io.xs = d_xs; io.ts = d_ts; io.ys = d_ys; io.n_rows = n_rows; io.n_total = n_total; io.n_rank = n_rank;
  io.sharded = sharded; io.pe = pe; io.batch = batch; io.seed = gp->seed;
  rc = gnet_train(ctx, gp, io);
  n_launches += io.n_launches;
  if (rc) return rc;
  const double lr = io.lr, best = io.best_loss;
  const int epochs_run = io.epochs_run;
  const bool have_best = io.have_best;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  out->epochs_run = epochs_run; out->best_loss = have_best ? best : nan(""); out->final_lr = lr;

  // ---- pass 2 ----
  WalkGArgs wa{};
  wa.S = S; wa.ld = ld; wa.M = M; wa.N = N; wa.is_put = lp->is_put; wa.sticky = sticky ? 1 : 0;
  wa.K = lp->K; wa.invK = 1.0 / lp->K; wa.params = d_params; wa.wpack = d_pack; wa.nm = d_norm;
  wa.sqrt_tau = d_tab + (N + 1); wa.Dm = d_tab + 2 * (N + 1);
  const int inf_drop = gp->inference_dropout < 0 ? (refdisc && sticky ? 1 : 0) : gp->inference_dropout;
  wa.drop = make_drop(inf_drop ? gp->dropout : 0.0, (unsigned int)gp->seed * 0x2545f491u + 0x51ed270bu + (sharded ? (unsigned int)pe.rank * 0x3c6ef372u : 0u));
  wa.fin = d_fin; wa.flags = ctx->d_flags;
  const bool stats = out->ex_count != nullptr || out->boundary != nullptr;
  wa.exc = stats ? ctx->d_exc : nullptr; wa.bnd = stats ? ctx->d_bnd : nullptr;
  const unsigned wg = (unsigned)((M + 127) / 128);
  gnet_walk_kernel<R><<<wg, kGThreads, gnet_smem_bytes(), ctx->stream>>>(wa);
  if (sharded) {  // the ranks' fixed-point value sums: integer addition, the same price on every rank
    gnet_gather_kernel<<<1, 128, 0, ctx->stream>>>(d_fin, 4, pe, ctx->comm.gn_meta++, d_gath, d_tot, ctx->d_flags);
    ++n_launches; ctx->launches++;
  }
  gnet_final_kernel<<<1, 1, 0, ctx->stream>>>(sharded ? d_tot : d_fin, sharded ? M_total : M, final_scale, ctx->d_final);
  n_launches += 2; ctx->launches += 2;
  OPTMC_CUDA(cudaGetLastError());
  OPTMC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));

  double fin[4];
  int flags[4];
  OPTMC_CUDA(cudaMemcpyAsync(fin, ctx->d_final, sizeof(fin), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  if (gp->final_params) OPTMC_CUDA(cudaMemcpyAsync(gp->final_params, d_params, (size_t)kGP * 4, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<unsigned long long> hb, he;
  if (out->boundary) { hb.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hb.data(), ctx->d_bnd, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->ex_count) { he.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(he.data(), ctx->d_exc, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (sharded && flags[1]) { set_error("sharded network LSM: a peer rank did not answer (exchange timed out); re-run optmc_comm_export / optmc_comm_init on every rank"); return OPTMC_ECUDA; }
  if (flags[0]) { set_error("network LSM: a sum left the fixed-point range or is not finite"); return OPTMC_EUNSUPPORTED; }
  float t01 = 0.f, t12 = 0.f;
  cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
  ctx->last_paths_ms = t01; ctx->last_sweep_ms = t12;  // here: fit (pass 1 + training) and pass 2
  out->price = fin[0]; out->stderr_ = fin[1]; out->n_paths = sharded ? M_total : M; out->n_launches = n_launches;
  if (out->boundary) {
    const unsigned long long none = lp->is_put ? 0ull : ~0ull;
    for (int t = 0; t <= N; ++t) {
      if (hb[t] == none) out->boundary[t] = nan("");
      else memcpy(&out->boundary[t], &hb[t], 8);
    }
  }
  if (out->ex_count) for (int t = 0; t <= N; ++t) out->ex_count[t] = (int64_t)he[t];
  return OPTMC_OK;
}

// ---- PER-DATE fit of the same network (gp->per_date): the loop of om2:277-310 / om15:145-186 with om3's regressor ----
// At every exercise date a fresh SingleLSMNet(7, 128, 3) is trained on the date's live rows (seven reference features of
// (S/K, tau), z-scored with the date's own moments -- constant features get std 1, om3:561 -- against the z-scored
// current cash-flows) by the mini-batch loop of gnet_train, and its in-sample prediction decides (strict '>').  The
// cash-flows live in HBM in date-N money with the exercised flag in the sign bit, exactly as in the other split sweeps.
template <typename R>
__device__ __forceinline__ bool gpd_live(R s, R c, double K, int is_put, int sticky) {
  return !(sticky && signbit(c)) && payoff<double>((double)s, K, is_put != 0) > 0.0;
}

template <typename R>
__global__ void __launch_bounds__(256) gpd_count_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, double K, int is_put,
                                                        int sticky, unsigned long long* __restrict__ counts) {
  const long long base = (long long)blockIdx.x * kGChunk;
  unsigned int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long j = base + k * 256 + threadIdx.x;
    if (j < M) c += gpd_live<R>(S_t[j], cf[j], K, is_put, sticky) ? 1u : 0u;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ unsigned int s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tot = 0;
    for (int w = 0; w < 8; ++w) tot += s[w];
    counts[blockIdx.x] = tot;
  }
}

// rows of one date in the order of gnet_compact_kernel (k-major inside a 1024-path chunk); y = value of the cash-flow at
// date t (date-N money x D_t); idx = the path of the row
template <typename R>
__global__ void __launch_bounds__(256) gpd_compact_kernel(const R* __restrict__ S_t, const R* __restrict__ cf, long long M, int t, double K,
                                                          int is_put, int sticky, double dg, double stau,
                                                          const unsigned long long* __restrict__ offsets, float* __restrict__ xs,
                                                          int* __restrict__ ts, float* __restrict__ ys, unsigned int* __restrict__ idx,
                                                          unsigned long long* __restrict__ sums, int* flags) {
  const long long base = (long long)blockIdx.x * kGChunk;
  const double invK = 1.0 / K;
  __shared__ unsigned int wsum[4][8];
  __shared__ double red[8][kGQ];
  double acc[kGQ];
#pragma unroll
  for (int q = 0; q < kGQ; ++q) acc[q] = 0.0;
  const unsigned long long off = offsets[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool live[4];
  double sv[4], cv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long j = base + k * 256 + threadIdx.x;
    sv[k] = 0.0; cv[k] = 0.0; live[k] = false;
    if (j < M) {
      const R s = S_t[j], c = cf[j];
      sv[k] = (double)s; cv[k] = fabs((double)c);
      live[k] = gpd_live<R>(s, c, K, is_put, sticky);
    }
    const unsigned int b = __ballot_sync(0xffffffffu, live[k]);
    if (lane == 0) wsum[k][warp] = __popc(b);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned int before = 0;
    for (int kk = 0; kk < k; ++kk)
      for (int w = 0; w < 8; ++w) before += wsum[kk][w];
    for (int w = 0; w < warp; ++w) before += wsum[k][w];
    const unsigned int b = __ballot_sync(0xffffffffu, live[k]);
    if (live[k]) {
      const long long j = base + k * 256 + threadIdx.x;
      const unsigned long long r = off + before + __popc(b & ((1u << lane) - 1u));
      const double x = sv[k] * invK;
      const double y = cv[k] * dg;
      xs[r] = (float)x; ts[r] = t; ys[r] = (float)y; idx[r] = (unsigned int)j;
      double f[7];
      f[0] = 1.0; f[1] = x; f[2] = x * x; f[3] = x * x * x; f[4] = x > 1.0 ? x - 1.0 : 0.0; f[5] = stau; f[6] = x * stau;
#pragma unroll
      for (int q = 0; q < 7; ++q) { acc[2 * q] += f[q]; acc[2 * q + 1] += f[q] * f[q]; }
      acc[14] += y; acc[15] += y * y;
    }
  }
#pragma unroll
  for (int q = 0; q < kGQ; ++q) {
    double v = acc[q];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < kGQ) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    if (v != 0.0) {
      unsigned long long hi, lo;
      if (!fx_encode(v, hi, lo)) atomicExch(flags, 1);
      atomicAdd(sums + 2 * threadIdx.x, hi - (1ull << 47));
      atomicAdd(sums + 2 * threadIdx.x + 1, lo);
    }
  }
}

// in-sample continuation of the date's rows on the tensor cores + exercise decision (strict '>', om2:304 / om3:644)
template <typename R>
__global__ void __launch_bounds__(kGThreads, 1) gpd_decide_kernel(const float* __restrict__ params, const __nv_bfloat16* __restrict__ wpack,
                                                                  const GnetNorm* __restrict__ nmp, const float* __restrict__ xs,
                                                                  const unsigned int* __restrict__ idx, long long n, int t, float stau,
                                                                  const R* __restrict__ S_t, R* cf, double K, double Kh, double Kl, int is_put,
                                                                  int sticky, R dinv, Drop drop, unsigned long long* exc_t,
                                                                  unsigned long long* bnd_t) {
  extern __shared__ __align__(1024) unsigned char smem_g[];
  GnetSmem& sm = *reinterpret_cast<GnetSmem*>(smem_g);
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const unsigned int tmem = gnet_setup(sm, params, wpack);
  if (tid == 0) gnet_weights_ready(sm);
  const unsigned int lane_base = (unsigned int)((warp & 3) * 32) << 16;
  const long long r = (long long)blockIdx.x * 128 + row;
  const bool act = r < n;
  const GnetNorm nm = *nmp;
  const unsigned int j = act ? idx[r] : 0u;
  float fn[kGIn];
  gnet_features(act ? xs[r] : 0.f, stau, nm, fn);
  const unsigned int rr = j * 0x01000193u + (unsigned int)t;  // the decision pass's stream (gnet_walk_kernel)
  unsigned int phase = 0;
  gnet_layer1(sm, fn, row, half, rr, drop);
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ab_t(tmem, sm.A1, sm.W2);
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  gnet_hidden<false>(sm, tmem + lane_base, sm.b2, sm.A2, row, half, rr, 1, drop);
  tc_publish();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ab_t(tmem + 128, sm.A2, sm.W3);
    umma_commit(&sm.bar);
  }
  tc_bar_wait(&sm.bar, phase); phase ^= 1u;
  const float out = gnet_join_dot(sm, gnet_hidden<true, false>(sm, tmem + lane_base + 128, sm.b3, sm.A3, row, half, rr, 2, drop), row, half);
  unsigned int cnt = 0;
  unsigned long long bnd = bnd_none(is_put);
  if (act && half == 0) {
    const R sr = S_t[j];
    const double s = (double)sr;
    const double pay = payoff<double>(s, K, is_put != 0);
    const double cont = (double)(out * nm.ystd + nm.ymean);  // om3:640
    if (pay > cont) {
      const R sgn = is_put ? (R)-1 : (R)1;
      const R a = (fma(sgn, sr, (R)(is_put ? Kh : -Kh)) + (R)(is_put ? Kl : -Kl)) * dinv;  // payoff in date-N money
      cf[j] = sticky ? -a : a;
      cnt = 1;
      bnd = (unsigned long long)__double_as_longlong(s);
    }
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// Host loop; the sweep is bound to the context (api.cu: bind_sweep) like the other split sweeps.  The first fitted date
// (t = N - 1) uses gp->seed itself, so that a slab with ONE exercise date (N = 2) reproduces optmc_lsm_gnet bit for bit.
template <typename R>
static int lsm_gnet_per_date_t(optmc_ctx* ctx, const optmc_gnet_params* gp, optmc_gnet_result* out) {
  SweepDesc& sw = ctx->sw;
  const long long M = sw.M;
  const int N = sw.N;
  const optmc_lsm_params* lp = &sw.lp;
  const bool sticky = (lp->semantics & OPTMC_SEM_STICKY_MASK) != 0;
  const bool refdisc = (lp->semantics & OPTMC_SEM_REF_DISCOUNT) != 0;
  const double dt = lp->T / N;
  const int nchunks = (int)((M + kGChunk - 1) / kGChunk);
  const int batch = gp->batch > 0 ? gp->batch : 256;
  const int max_tiles = (batch + 127) / 128;
  std::vector<double> st((size_t)N + 1);
  for (int t = 0; t <= N; ++t) { const double tau = lp->T - t * dt; st[t] = sqrt(tau > 1e-6 ? tau : 1e-6); }
  constexpr int kMetaN = 1 + 2 * kGQ;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t seg = ((size_t)M * 4 + 255) / 256 * 256;
  const size_t o_tab = take(st.size() * 8), o_counts = take((size_t)nchunks * 8), o_meta = take(kMetaN * 8), o_norm = take(sizeof(GnetNorm)),
               o_params = take((size_t)kGP * 4), o_best = take((size_t)kGP * 4), o_m = take((size_t)kGP * 4), o_v = take((size_t)kGP * 4),
               o_gpart = take((size_t)max_tiles * (kGP + 1) * 4), o_loss = take(9 * 8), o_pack = take(2 * kTcTileBytes),
               o_rows = take(4 * seg);
  int rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  double* d_tab = reinterpret_cast<double*>(dev + o_tab);
  unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(dev + o_counts);
  unsigned long long* d_meta = reinterpret_cast<unsigned long long*>(dev + o_meta);
  long long* d_nrows = reinterpret_cast<long long*>(d_meta);
  unsigned long long* d_sums = d_meta + 1;
  GnetNorm* d_norm = reinterpret_cast<GnetNorm*>(dev + o_norm);
  float* d_xs = reinterpret_cast<float*>(dev + o_rows);
  int* d_ts = reinterpret_cast<int*>(dev + o_rows + seg);
  float* d_ys = reinterpret_cast<float*>(dev + o_rows + 2 * seg);
  unsigned int* d_idx = reinterpret_cast<unsigned int*>(dev + o_rows + 3 * seg);
  OPTMC_CUDA(cudaMemcpyAsync(d_tab, st.data(), st.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaFuncSetAttribute(gnet_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnet_smem_bytes()));
  OPTMC_CUDA(cudaFuncSetAttribute(gpd_decide_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnet_smem_bytes()));
  rc = sweep_begin(ctx);  // cf = payoff(S[N]) (date-N money), per-date statistics reset
  if (rc) return rc;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  const R* Sr = static_cast<const R*>(sw.S);
  R* cf = static_cast<R*>(ctx->cf);
  const int inf_drop = gp->inference_dropout < 0 ? (refdisc && sticky ? 1 : 0) : gp->inference_dropout;
  long long rows_all = 0;
  int epochs_all = 0, n_launches = 0, fitted = 0;
  double loss_sum = 0.0, lr_last = gp->lr;
  GnetTrainIo io{};
  io.params = reinterpret_cast<float*>(dev + o_params); io.best = reinterpret_cast<float*>(dev + o_best);
  io.m = reinterpret_cast<float*>(dev + o_m); io.v = reinterpret_cast<float*>(dev + o_v);
  io.gpart = reinterpret_cast<float*>(dev + o_gpart); io.loss = reinterpret_cast<double*>(dev + o_loss);
  io.pack = reinterpret_cast<__nv_bfloat16*>(dev + o_pack); io.norm = d_norm; io.sqrt_tau = d_tab;
  io.xs = d_xs; io.ts = d_ts; io.ys = d_ys; io.sharded = false; io.batch = batch;
  for (int t = N - 1; t >= 1; --t) {
    const R* S_t = Sr + (size_t)t * sw.ld;
    const unsigned long long seed_t = gp->seed + (unsigned long long)(N - 1 - t) * 0x9e3779b97f4a7c15ull;
    OPTMC_CUDA(cudaMemsetAsync(d_meta, 0, kMetaN * 8, ctx->stream));
    gpd_count_kernel<R><<<nchunks, 256, 0, ctx->stream>>>(S_t, cf, M, lp->K, lp->is_put, sticky ? 1 : 0, d_counts);
    gnet_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nchunks, d_nrows);
    n_launches += 2; ctx->launches += 2;
    long long n_rows = 0;
    OPTMC_CUDA(cudaMemcpyAsync(&n_rows, d_nrows, 8, cudaMemcpyDeviceToHost, ctx->stream));
    OPTMC_CUDA(cudaMemcpyAsync(ctx->d_nitm + t, d_nrows, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_rows == 0) continue;  // om2:284-285
    gpd_compact_kernel<R><<<nchunks, 256, 0, ctx->stream>>>(S_t, cf, M, t, lp->K, lp->is_put, sticky ? 1 : 0, sw.Dt[t], st[t], d_counts, d_xs,
                                                            d_ts, d_ys, d_idx, d_sums, ctx->d_flags);
    gnet_norm_kernel<<<1, 32, 0, ctx->stream>>>(d_sums, d_nrows, gp->target_ddof ? 1 : 0, d_norm);
    gnet_init_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(io.params, io.m, io.v, seed_t);
    gnet_pack_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(io.params, io.pack);
    n_launches += 4; ctx->launches += 4;
    const long long n_rank[1] = {n_rows};
    io.n_rows = n_rows; io.n_total = n_rows; io.n_rank = n_rank; io.seed = seed_t; io.n_launches = 0;
    rc = gnet_train(ctx, gp, io);
    n_launches += io.n_launches;
    if (rc) return rc;
    const Drop drop = make_drop(inf_drop ? gp->dropout : 0.0, (unsigned int)seed_t * 0x2545f491u + 0x51ed270bu);
    const bool stats = out->ex_count != nullptr || out->boundary != nullptr;
    gpd_decide_kernel<R><<<(unsigned)((n_rows + 127) / 128), kGThreads, gnet_smem_bytes(), ctx->stream>>>(
        io.params, io.pack, d_norm, d_xs, d_idx, n_rows, t, (float)st[t], S_t, cf, lp->K, sw.Kh, sw.Kl, lp->is_put, sticky ? 1 : 0, (R)sw.Dinv[t],
        drop, stats ? ctx->d_exc + t : nullptr, stats ? ctx->d_bnd + t : nullptr);
    ++n_launches; ctx->launches++;
    OPTMC_CUDA(cudaGetLastError());
    rows_all += n_rows; epochs_all += io.epochs_run; lr_last = io.lr; ++fitted;
    if (io.have_best) loss_sum += io.best_loss;
  }
  OPTMC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  rc = sweep_finish(ctx, ctx->gram);
  if (rc) return rc;
  rc = sweep_finalize_price(ctx, ctx->gram);
  if (rc) return rc;
  OPTMC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  sw.impl_used = OPTMC_SWEEP_SPLIT;
  sw.have_results = true;
  double fin[4];
  int flags[4];
  OPTMC_CUDA(cudaMemcpyAsync(fin, ctx->d_final, sizeof(fin), cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
  if (gp->final_params) OPTMC_CUDA(cudaMemcpyAsync(gp->final_params, io.params, (size_t)kGP * 4, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<unsigned long long> hb, he;
  if (out->boundary) { hb.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(hb.data(), ctx->d_bnd, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (out->ex_count) { he.resize(N + 1); OPTMC_CUDA(cudaMemcpyAsync(he.data(), ctx->d_exc, (size_t)(N + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); }
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flags[0]) { set_error("network LSM: a sum left the fixed-point range or is not finite"); return OPTMC_EUNSUPPORTED; }
  float t01 = 0.f, t12 = 0.f;
  cudaEventElapsedTime(&t01, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&t12, ctx->ev[1], ctx->ev[2]);
  ctx->last_paths_ms = 0.0; ctx->last_sweep_ms = t01 + t12;
  out->price = fin[0]; out->stderr_ = fin[1]; out->n_paths = M; out->n_rows = rows_all; out->epochs_run = epochs_all;
  out->n_launches = n_launches + sw.n_launches;
  out->best_loss = fitted ? loss_sum / fitted : nan("");  // mean over the fitted dates
  out->final_lr = lr_last;
  if (out->boundary) {
    const unsigned long long none = lp->is_put ? 0ull : ~0ull;
    for (int t = 0; t <= N; ++t) {
      if (hb[t] == none) out->boundary[t] = nan("");
      else memcpy(&out->boundary[t], &hb[t], 8);
    }
  }
  if (out->ex_count) for (int t = 0; t <= N; ++t) out->ex_count[t] = (int64_t)he[t];
  return OPTMC_OK;
}

int lsm_gnet_per_date(optmc_ctx* ctx, const optmc_gnet_params* gp, optmc_gnet_result* out) {
  return ctx->sw.dtype == OPTMC_F64 ? lsm_gnet_per_date_t<double>(ctx, gp, out) : lsm_gnet_per_date_t<float>(ctx, gp, out);
}

int gnet_validate(optmc_ctx* ctx, const optmc_gnet_params* gp) {
  if (gp->hidden != kGH || gp->layers != 3) { set_error("network LSM: the tensor-core kernels are built for SingleLSMNet(7, 128, 3)"); return OPTMC_EUNSUPPORTED; }
  if (gp->epochs < 0 || gp->batch < 0 || gp->batch > 128 * 1024 || !(gp->lr > 0) || gp->dropout < 0 || gp->dropout >= 1) {
    set_error("network LSM: bad training parameters"); return OPTMC_EINVAL;
  }
  if (ctx->cc < 100) { set_error("network LSM needs tcgen05 (sm_100)"); return OPTMC_EUNSUPPORTED; }
  return OPTMC_OK;
}

int lsm_gnet(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int64_t M_total, int32_t N, int32_t dtype, const optmc_lsm_params* lp,
             const optmc_gnet_params* gp, optmc_gnet_result* out) {
  if (!S || !lp || !gp || !out) { set_error("null argument"); return OPTMC_EINVAL; }
  if (!(lp->K > 0) || !(lp->T > 0)) { set_error("S0, K, T must be positive."); return OPTMC_EINVAL; }
  if (lp->r < 0) { set_error("r must be non-negative."); return OPTMC_EINVAL; }
  if (M <= 0 || N <= 0) { set_error("num_simulations and num_time_steps must be positive integers."); return OPTMC_EINVAL; }
  if (ld < M) { set_error("ld must be >= M"); return OPTMC_EINVAL; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  const int rcv = gnet_validate(ctx, gp);
  if (rcv) return rcv;
  return dtype == OPTMC_F64 ? lsm_gnet_t<double>(ctx, S, ld, M, M_total, N, lp, gp, out)
                            : lsm_gnet_t<float>(ctx, S, ld, M, M_total, N, lp, gp, out);
}

// Test aid: loss and parameter gradients of one batch of n <= 16384 rows given as normalised features [n][7] and
// targets [n], without dropout -- compared with torch autograd in tests/test_gpu_network.py.  Host pointers.
int gnet_grad_debug(optmc_ctx* ctx, long long n, const float* feat, const float* ys, const float* params, float* grads, float* loss) {
  if (!feat || !ys || !params || !grads || !loss || n <= 0 || n > 16384) { set_error("bad argument"); return OPTMC_EINVAL; }
  if (ctx->cc < 100) { set_error("network LSM needs tcgen05 (sm_100)"); return OPTMC_EUNSUPPORTED; }
  const int tiles = (int)((n + 127) / 128);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_feat = take((size_t)n * kGIn * 4), o_ys = take((size_t)n * 4), o_par = take((size_t)kGP * 4),
               o_gpart = take((size_t)tiles * (kGP + 1) * 4), o_out = take((size_t)(kGP + 1) * 4), o_pack = take(2 * kTcTileBytes);
  int rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  OPTMC_CUDA(cudaMemcpyAsync(dev + o_feat, feat, (size_t)n * kGIn * 4, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(dev + o_ys, ys, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaMemcpyAsync(dev + o_par, params, (size_t)kGP * 4, cudaMemcpyHostToDevice, ctx->stream));
  OPTMC_CUDA(cudaFuncSetAttribute(gnet_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnet_smem_bytes()));
  GradArgs ga{};
  ga.params = reinterpret_cast<float*>(dev + o_par); ga.feat = reinterpret_cast<float*>(dev + o_feat);
  ga.wpack = reinterpret_cast<__nv_bfloat16*>(dev + o_pack);
  gnet_pack_kernel<<<(kGP + 255) / 256, 256, 0, ctx->stream>>>(ga.params, reinterpret_cast<__nv_bfloat16*>(dev + o_pack));
  ga.ys = reinterpret_cast<float*>(dev + o_ys); ga.start = 0; ga.end = n; ga.inv_b2 = 2.0f / (float)n; ga.perm = make_perm(1, 0); ga.drop = make_drop(0.0, 0);
  ga.gpart = reinterpret_cast<float*>(dev + o_gpart);
  gnet_grad_kernel<<<tiles, kGThreads, gnet_smem_bytes(), ctx->stream>>>(ga);
  gnet_sum_partials_kernel<<<(kGP + 256) / 256, 256, 0, ctx->stream>>>(ga.gpart, tiles, reinterpret_cast<float*>(dev + o_out), 1.0f / (float)n);
  ctx->launches += 2;
  OPTMC_CUDA(cudaGetLastError());
  std::vector<float> h(kGP + 1);
  OPTMC_CUDA(cudaMemcpyAsync(h.data(), dev + o_out, (size_t)(kGP + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(grads, h.data(), (size_t)kGP * 4);
  *loss = h[kGP];
  return OPTMC_OK;
}

// Test aid: the engine's shuffle and dropout streams as the training / decision kernels evaluate them.  perm_out[i] = the
// row position i of epoch `epoch` reads (n_rows rows); keep_out[(i * 3 + layer) * 4 + w] = 32 keep bits (units 32 w ..)
// of row id row_ids[i] under the stream of optimiser step `step` (step > 0) or of the decision pass (step == 0).
__global__ void gnet_streams_kernel(Perm perm, Drop drop, long long n_perm, const unsigned int* __restrict__ row_ids, long long n_ids,
                                    long long* __restrict__ perm_out, unsigned int* __restrict__ keep_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_perm) perm_out[i] = (long long)perm_apply(perm, (unsigned long long)i);
  if (i < n_ids) {
    const unsigned int r = row_ids[i];
    for (int layer = 0; layer < 3; ++layer)
      for (int w = 0; w < 4; ++w) {
        unsigned int word = 0u;
        for (int q = 0; q < 4; ++q) word |= drop_keep8(drop, r, layer, w * 4 + q) << (q * 8);
        keep_out[(i * 3 + layer) * 4 + w] = word;
      }
  }
}

int gnet_streams_debug(optmc_ctx* ctx, unsigned long long seed, int epoch, int step, double dropout, long long n_rows,
                       long long* perm_out, const unsigned int* row_ids, long long n_ids, unsigned int* keep_out) {
  if (n_rows < 0 || n_ids < 0 || epoch < 0 || step < 0 || (n_rows > 0 && !perm_out) || (n_ids > 0 && (!row_ids || !keep_out))) {
    set_error("bad argument"); return OPTMC_EINVAL;
  }
  const long long n = n_rows > n_ids ? n_rows : n_ids;
  if (n == 0) return OPTMC_OK;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_perm = take((size_t)n_rows * 8), o_ids = take((size_t)n_ids * 4), o_keep = take((size_t)n_ids * 48);
  int rc = ensure_bytes(&ctx->batch_dev, &ctx->batch_dev_cap, off);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->batch_dev);
  if (n_ids) OPTMC_CUDA(cudaMemcpyAsync(dev + o_ids, row_ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, ctx->stream));
  const Perm perm = make_perm((unsigned long long)n_rows, (unsigned int)(seed * 0x9e3779b97f4a7c15ull >> 32) + 0x632be5abu * (unsigned int)(epoch + 1));
  const Drop drop = make_drop(dropout, (unsigned int)seed * 0x2545f491u + (step > 0 ? (unsigned int)step * 0x9e3779b1u : 0x51ed270bu));
  gnet_streams_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(perm, drop, n_rows, reinterpret_cast<unsigned int*>(dev + o_ids), n_ids,
                                                                           reinterpret_cast<long long*>(dev + o_perm),
                                                                           reinterpret_cast<unsigned int*>(dev + o_keep));
  ctx->launches += 1;
  OPTMC_CUDA(cudaGetLastError());
  if (n_rows) OPTMC_CUDA(cudaMemcpyAsync(perm_out, dev + o_perm, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_ids) OPTMC_CUDA(cudaMemcpyAsync(keep_out, dev + o_keep, (size_t)n_ids * 48, cudaMemcpyDeviceToHost, ctx->stream));
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPTMC_OK;
}

}  // namespace optmc
