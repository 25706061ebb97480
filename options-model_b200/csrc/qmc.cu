// qmc.cu -- Sobol' + Brownian-bridge normals (no counterpart in the reference, which draws PCG64
// pseudo-random normals, om3:223-224,475).  A separate generator kernel writes the step-major normals Z[N][M/2] that
// every path kernel accepts as external draws (optmc_rng_params.z1_dev / z2_dev), so all schemes can run on them.
//
//   point i (= antithetic pair index, counters stay global through pair_offset) of the Sobol' sequence in
//   D = factors * N dimensions (Joe-Kuo direction numbers, sobol_table.h; Gray-code formula, so any block of indices
//   can be generated independently); a random digital shift per dimension (host-supplied 32-bit words) randomises the
//   point set without touching its net structure; u = (x + 1/2) 2^-32, z = Phi^-1(u).
//   Brownian bridge: dimension 0 fixes W(N), the following ones the midpoints of ever finer intervals (schedule built
//   on the host for any N), so the variance of the path sits in the leading -- best distributed -- dimensions; the
//   normals handed to the path kernel are the increments W(t) - W(t-1) (unit time grid: standard normals).
#include <algorithm>
#include <vector>

#include "optmc_internal.h"
#include "sobol_table.h"

namespace optmc {

// path row p holds W(p + 1).  W[idx] = wl * W[left - 1] (0 if left == 0) + wr * W[right] + sd * z; step 0 sets the last row.
struct BridgeStep { int idx, left, right, pad; double wl, wr, sd; };

// Standard construction (Jaeckel, "Monte Carlo Methods in Finance", 2002, sec. 10.8.3): the terminal point first, then
// always the middle of the next unpopulated gap, cycling through the grid.
static std::vector<BridgeStep> bridge_schedule(int N) {
  std::vector<BridgeStep> s(N);
  std::vector<int> map(N, 0);
  map[N - 1] = 1;
  s[0] = {N - 1, 0, N - 1, 0, 0.0, 0.0, sqrt((double)N)};
  int j = 0;
  for (int i = 1; i < N; ++i) {
    while (map[j]) ++j;                 // next unpopulated entry
    int k = j;
    while (!map[k]) ++k;                // next populated entry behind it
    const int l = j + ((k - 1 - j) >> 1);
    map[l] = i + 1;
    s[i] = {l, j, k, 0, (double)(k - l) / (double)(k + 1 - j), (double)(l + 1 - j) / (double)(k + 1 - j),
            sqrt((double)(l + 1 - j) * (double)(k - l) / (double)(k + 1 - j))};
    j = k + 1;
    if (j >= N) j = 0;
  }
  return s;
}

struct QmcArgs {
  long long Mh, pair_offset;
  int N, factors, bridge, f64;
  const uint32_t* V;         // [D][32]
  const uint32_t* shift;     // [D] or NULL
  const BridgeStep* sched;   // [N]
  void* Z1;
  void* Z2;
};

__device__ __forceinline__ double sobol_normal(const uint32_t* __restrict__ V, const uint32_t* __restrict__ shift, int d,
                                               unsigned long long gray) {
  uint32_t x = shift ? shift[d] : 0u;
  const uint32_t* v = V + (size_t)d * 32;
  unsigned long long g = gray;
  int b = 0;
  while (g) {
    if (g & 1ull) x ^= v[b];
    g >>= 1;
    ++b;
  }
  const double u = ((double)x + 0.5) * 2.3283064365386963e-10;  // (0, 1)
  return normcdfinv(u);
}

template <typename R> __global__ void __launch_bounds__(256) qmc_normals_kernel(const QmcArgs a) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.Mh) return;
  const unsigned long long i = (unsigned long long)(a.pair_offset + j) + 1ull;  // skip the all-zero point
  const unsigned long long gray = i ^ (i >> 1);
  for (int f = 0; f < a.factors; ++f) {
    R* Z = static_cast<R*>(f == 0 ? a.Z1 : a.Z2) + j;  // column j, row stride Mh
    const int d0 = f * a.N;
    if (!a.bridge) {
      for (int t = 0; t < a.N; ++t) Z[(size_t)t * a.Mh] = (R)sobol_normal(a.V, a.shift, d0 + t, gray);
      continue;
    }
    // W(p + 1) is built in place in row p, then differenced from the back
    for (int s = 0; s < a.N; ++s) {
      const BridgeStep st = a.sched[s];
      const double z = sobol_normal(a.V, a.shift, d0 + s, gray);
      const double wl = (s > 0 && st.left > 0) ? (double)Z[(size_t)(st.left - 1) * a.Mh] : 0.0;
      const double wr = s > 0 ? (double)Z[(size_t)st.right * a.Mh] : 0.0;
      Z[(size_t)st.idx * a.Mh] = (R)(st.wl * wl + st.wr * wr + st.sd * z);
    }
    for (int p = a.N - 1; p >= 1; --p)
      Z[(size_t)p * a.Mh] = (R)((double)Z[(size_t)p * a.Mh] - (double)Z[(size_t)(p - 1) * a.Mh]);
  }
}

int launch_qmc_normals(optmc_ctx* ctx, int64_t M, int32_t N, int32_t factors, int32_t bridge, int64_t pair_offset,
                       const uint32_t* shift_host, int32_t dtype, void* Z1, void* Z2) {
  if (M <= 0 || (M & 1) || N <= 0) { set_error("QMC normals need an even, positive path count and N > 0"); return OPTMC_EINVAL; }
  if (factors < 1 || factors > 2 || !Z1 || (factors == 2 && !Z2)) { set_error("bad arguments"); return OPTMC_EINVAL; }
  const int D = factors * N;
  if (D > kSobolDims) { set_error("Sobol table holds 512 dimensions (factors x steps)"); return OPTMC_EUNSUPPORTED; }
  if (dtype != OPTMC_F32 && dtype != OPTMC_F64) { set_error("bad dtype"); return OPTMC_EINVAL; }
  const size_t bytesV = sizeof(kSobolV), bytesS = sizeof(uint32_t) * D, bytesB = sizeof(BridgeStep) * N;
  int rc = ensure_bytes(&ctx->qmc_dev, &ctx->qmc_dev_cap, bytesV + ((bytesS + 255) / 256 * 256) + bytesB + 512);
  if (rc) return rc;
  char* dev = static_cast<char*>(ctx->qmc_dev);
  if (!ctx->qmc_table_ready) {
    OPTMC_CUDA(cudaMemcpyAsync(dev, kSobolV, bytesV, cudaMemcpyHostToDevice, ctx->stream));
    ctx->qmc_table_ready = true;
  }
  QmcArgs a{};
  a.Mh = M / 2; a.pair_offset = pair_offset; a.N = N; a.factors = factors; a.bridge = bridge ? 1 : 0; a.f64 = dtype == OPTMC_F64;
  a.V = reinterpret_cast<const uint32_t*>(dev);
  char* p = dev + bytesV;
  if (shift_host) {
    OPTMC_CUDA(cudaMemcpyAsync(p, shift_host, bytesS, cudaMemcpyHostToDevice, ctx->stream));
    a.shift = reinterpret_cast<const uint32_t*>(p);
  }
  p += (bytesS + 255) / 256 * 256;
  std::vector<BridgeStep> sched;
  if (bridge) {
    sched = bridge_schedule(N);
    OPTMC_CUDA(cudaMemcpyAsync(p, sched.data(), bytesB, cudaMemcpyHostToDevice, ctx->stream));
    a.sched = reinterpret_cast<const BridgeStep*>(p);
  }
  a.Z1 = Z1; a.Z2 = Z2;
  const unsigned grid = (unsigned)((a.Mh + 255) / 256);
  if (dtype == OPTMC_F64) qmc_normals_kernel<double><<<grid, 256, 0, ctx->stream>>>(a);
  else qmc_normals_kernel<float><<<grid, 256, 0, ctx->stream>>>(a);
  ctx->launches++;
  OPTMC_CUDA(cudaGetLastError());
  OPTMC_CUDA(cudaStreamSynchronize(ctx->stream));  // the host-side schedule / shift buffers go out of scope
  return OPTMC_OK;
}

// host copy of the schedule (test aid: the oracle restates the construction from it)
int bridge_schedule_host(int32_t N, int32_t* idx, int32_t* left, int32_t* right, double* wl, double* wr, double* sd) {
  if (N <= 0) { set_error("N must be positive"); return OPTMC_EINVAL; }
  const std::vector<BridgeStep> s = bridge_schedule(N);
  for (int i = 0; i < N; ++i) { idx[i] = s[i].idx; left[i] = s[i].left; right[i] = s[i].right; wl[i] = s[i].wl; wr[i] = s[i].wr; sd[i] = s[i].sd; }
  return OPTMC_OK;
}

}  // namespace optmc
