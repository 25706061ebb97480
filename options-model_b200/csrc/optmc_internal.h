// optmc_internal.h -- context layout and launcher prototypes shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/optmc.h"

namespace optmc {

constexpr int kXchgMaxQ = 16;         // quantities per exchange (poly3 needs 11)
constexpr int kMaxResidentCtas = 160; // CTAs of the persistent sweep (B200: 148 SMs); < 255 (8-bit arrival count)
// Exchange buffer of the persistent sweep: [2 parities][kXchgWords] 64-bit accumulator words, packed
// (kXchgStride = 1): one warp's reds / polls of a parity then coalesce into one or two L2 requests per CTA,
// which measured 25% faster than one 128-byte line per word (tools/xchg_bench.cu) -- the exchange is bound
// by L2 request count on the hot lines, not by atomic throughput.
constexpr int kXchgWords = 2 * kXchgMaxQ;
constexpr int kCommMaxRanks = 8;   // ranks of a path-sharded sweep (optmc_comm_*)
constexpr int kCommMaxGroups = 16; // options of one path-sharded grouped launch (one slot block each)
constexpr int kXchgStride = 1;
// Peer-mapped region of the path-sharded network LSM (lsm_gnet.cu), behind the sweep's exchange slots in the same
// allocation: grad[2 parities][ranks][kGnetPad] and meta[2][ranks][kGnetMetaWords] 64-bit words {tag32, payload32}.
constexpr int kGnetPad = 34304;        // 34177 parameters + the batch loss, rounded up to 128 words
constexpr int kGnetMetaWords = 256;
constexpr size_t kCommSweepWords = (size_t)kCommMaxGroups * 2 * kCommMaxRanks * kXchgWords;
constexpr size_t kCommGnetWords = (size_t)2 * kCommMaxRanks * (kGnetPad + kGnetMetaWords);
inline size_t xchg_bytes() { return (size_t)2 * kXchgWords * kXchgStride * sizeof(unsigned long long); }
constexpr int kMaxBeta = 4;

struct SweepDesc {  // the sweep currently bound to the context (begin/gram/update/finish/fetch)
  const void* S = nullptr;
  int64_t ld = 0, M = 0;
  int32_t N = 0, dtype = 0;
  optmc_lsm_params lp{};
  int deg = 2;
  double disc = 1.0, final_scale = 1.0;
  // cash-flows live in "date-N money": Dt[t] = disc^(N - t) and Dinv[t] = disc^-(N - t), built by the same
  // running fp64 products the persistent kernel forms on the device (bit-identical factors)
  std::vector<double> Dt, Dinv;
  double Kh = 0.0, Kl = 0.0;  // K = Kh + Kl with Kh exact in the storage type (Kl = 0 for fp64 storage)
  int impl_used = 0;
  int n_launches = 0;
  bool have_results = false;
  bool finals_on_device = false;
  bool rerun_done = false;  // AUTO: the split sweep already replaced an overflowed resident sweep
  bool no_arrays = false;   // one-shot pricing whose caller asked for the price only: the persistent sweep skips the
                            // per-date outputs (betas, boundary, exercise / ITM counts stay NaN / none / 0)
};

// Strike constants of a sweep in the storage type (fp32 slabs: K = Kh + Kl with Kh a float; Kcmp = the float
// threshold for which (s < Kcmp) <=> (s < K) for every float s (puts), mirrored for calls; fp64 slabs: K itself).
// With these the in-the-money test and the payoff run in storage precision yet agree with the fp64 formulas.
struct StrikeConsts { double Kcmp, Kh, Kl; };
inline StrikeConsts strike_consts(double K, bool is_put, bool f32) {
  StrikeConsts c{K, K, 0.0};
  if (f32) {
    float kf = (float)K;
    c.Kh = (double)kf;
    c.Kl = (double)(float)(K - c.Kh);
    if (is_put) { if ((double)kf < K) kf = nextafterf(kf, INFINITY); }
    else { if ((double)kf > K) kf = nextafterf(kf, -INFINITY); }
    c.Kcmp = (double)kf;
  }
  return c;
}

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define OPTMC_CUDA(call)                                        \
  do {                                                          \
    cudaError_t _e = (call);                                    \
    if (_e != cudaSuccess) return ::optmc::cuda_fail(_e, #call); \
  } while (0)

}  // namespace optmc

struct optmc_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int64_t l2_bytes = 0;
  int max_smem_optin = 0;
  int cc = 0;
  int64_t launches = 0;
  std::vector<std::pair<unsigned long long, int>> sweep_choice;  // batch shape -> timed kernel choice (lsm_resident.cu)

  // grow-only device workspaces
  void* slab = nullptr;      size_t slab_bytes = 0;      // price_american path slab
  void* cf = nullptr;        size_t cf_bytes = 0;        // split sweep cash-flows
  double* partials = nullptr; size_t partials_bytes = 0; // [grid][Q] block partials (gram / final / european)
  unsigned int* tickets = nullptr;                        // last-block tickets [1024]
  double* gram = nullptr;                                 // [16] reduced moments of the current date
  double* d_betas = nullptr;   // [(N+1)][kMaxBeta]
  unsigned long long* d_bnd = nullptr;  // [(N+1)] boundary as double bits
  unsigned long long* d_exc = nullptr;  // [(N+1)]
  long long* d_nitm = nullptr;          // [(N+1)]
  int* d_valid = nullptr;               // [(N+1)]
  size_t per_date_cap = 0;              // N+1 capacity of the per-date arrays
  double* d_final = nullptr;            // [4] price, stderr, sum, sumsq (d_flags follows in the same allocation)
  void* h_fetch = nullptr; size_t h_fetch_cap = 0;  // pinned staging of fetch_results
  void* xchg = nullptr;                 // exchange accumulators of the persistent sweep, xchg_bytes()
  int* d_flags = nullptr;               // [4]: [0] = fixed-point exchange overflow
  void* batch_dev = nullptr; size_t batch_dev_cap = 0;  // per-wave descriptors / accumulators / results
  void* spill = nullptr;     size_t spill_bytes = 0;     // speculative sweep: candidate-list overflow
  void* qmc_dev = nullptr;   size_t qmc_dev_cap = 0;     // Sobol direction numbers, digital shifts, bridge schedule
  bool qmc_table_ready = false;
  void* gnet_rows = nullptr; size_t gnet_rows_cap = 0;  // global network LSM: regression rows of all dates (x, t, y)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};      // kernel timing of the fused calls (optmc_ctx_kernel_times)
  double last_paths_ms = 0.0, last_sweep_ms = 0.0;
  double* eu_out = nullptr; size_t eu_out_cap = 0;  // [n_options][3]
  double* eu_par = nullptr; size_t eu_par_cap = 0;  // [n_options][4] K, T, is_put, pad
  unsigned int* eu_tickets = nullptr; size_t eu_tickets_cap = 0;

  // path-sharded sweeps: peer-mapped exchange slots (optmc_comm_*; lsm_resident_kernel.cuh: ResComm)
  struct Comm {
    int nranks = 0, rank = 0;
    unsigned int g = 2;                      // running exchange counter (tags 2, 3 differ from the zeroed slots)
    unsigned int gn_step = 1, gn_meta = 1;   // network LSM: running tags of the gradient / small-vector exchanges (0 = empty)
    unsigned long long* local = nullptr;     // this rank's slot array
    unsigned long long* peers[optmc::kCommMaxRanks] = {};       // peers[rank] == local; the others are cudaIpcOpenMemHandle mappings
    bool opened[optmc::kCommMaxRanks] = {};
  } comm;
  int64_t sharded_M_total = 0;               // > 0 while optmc_lsm_poly_sharded runs its sweep

  optmc::SweepDesc sw;
};

namespace optmc {

int ensure_bytes(void** p, size_t* cap, size_t need);
int ensure_per_date(optmc_ctx* ctx, int N);

// paths.cu
int launch_paths(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int32_t N,
                 int32_t dtype, void* S, void* V, int64_t ld);
size_t path_args_bytes();
// paths_localvol.cu
int ivnet_sigma_batch(optmc_ctx* ctx, const optmc_ivnet* net, double tau, const double* S_dev, int64_t n, double* sigma_dev);
int launch_paths_localvol(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, const optmc_ivnet* net,
                          int64_t M, int32_t N, int32_t dtype, void* S, int64_t ld);
int prepare_paths_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int G,
                        const optmc_american_option* opts, void* slab, size_t slab_stride_bytes, int64_t ld, void* d_args,
                        void* h_args);
int launch_paths_batch(optmc_ctx* ctx, const optmc_model_params* mp, int32_t dtype, int G, void* slab,
                       size_t slab_stride_bytes, int64_t ld, void* d_args, void* h_args);
int launch_philox_normals(optmc_ctx* ctx, const optmc_rng_params* rng, int32_t model, int64_t M, int32_t N,
                          int32_t which, int32_t dtype, void* Z);
int launch_philox_kat(optmc_ctx* ctx, int n, const uint32_t* ctr, const uint32_t* key, uint32_t* out);
int launch_features(optmc_ctx* ctx, const void* S, int64_t n, int32_t dtype, double K, double T, double t_current,
                    void* F);

// qmc.cu
int launch_qmc_normals(optmc_ctx* ctx, int64_t M, int32_t N, int32_t factors, int32_t bridge, int64_t pair_offset,
                       const uint32_t* shift_host, int32_t dtype, void* Z1, void* Z2);
int bridge_schedule_host(int32_t N, int32_t* idx, int32_t* left, int32_t* right, double* wl, double* wr, double* sd);

// lsm.cu
int sweep_begin(optmc_ctx* ctx);
int sweep_zero_count(optmc_ctx* ctx, int64_t* count);
int sweep_split_fused(optmc_ctx* ctx);  // begin + one fused launch per date (internal split sweep)                                    // split: cf = payoff(S[N])
int sweep_gram_date(optmc_ctx* ctx, int t, double* gram_out);       // split: local moments of date t
int sweep_update_date(optmc_ctx* ctx, int t, const double* gram);   // split: solve + decide + discount
int sweep_finish(optmc_ctx* ctx, double* sums_out);                 // split: sum(cf), sum(cf^2), n
int sweep_finalize_price(optmc_ctx* ctx, const double* sums);       // split: d_final from sums
int sweep_reset_stats(optmc_ctx* ctx);                              // NaN/none-initialise the per-date outputs
bool resident_eligible(optmc_ctx* ctx, const SweepDesc& sw, std::string* why);
int sweep_resident(optmc_ctx* ctx);                                 // one cooperative launch, all dates
// grouped persistent sweep of a batch (lsm_resident.cu)
int price_american_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                         int32_t dtype, int32_t basis, uint32_t semantics, int32_t n_options,
                         const optmc_american_option* opts, optmc_price_result* results, optmc_batch_extras* ex);

// lsm_global.cu
int lsm_global(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype,
               const optmc_lsm_params* lp, optmc_global_result* out);

// lsm_mlp.cu
int lsm_mlp(optmc_ctx* ctx, const optmc_mlp_params* np, optmc_lsm_result* out);
int lsm_apply_policy(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int32_t N, int32_t dtype, const optmc_lsm_params* lp,
                     const double* betas, optmc_lsm_result* out);
// lsm_gnet.cu
int lsm_gnet(optmc_ctx* ctx, const void* S, int64_t ld, int64_t M, int64_t M_total /* > 0: path-sharded */, int32_t N, int32_t dtype,
             const optmc_lsm_params* lp, const optmc_gnet_params* gp, optmc_gnet_result* out);
int gnet_validate(optmc_ctx* ctx, const optmc_gnet_params* gp);
int lsm_gnet_per_date(optmc_ctx* ctx, const optmc_gnet_params* gp, optmc_gnet_result* out);  // sweep bound to the context
int gnet_shard_plan(const int64_t* n_rows, int32_t nranks, int32_t batch, int64_t b, int32_t rank, int64_t* lo, int64_t* hi,
                    int64_t* global_rows);
int gnet_grad_debug(optmc_ctx* ctx, long long n, const float* feat, const float* ys, const float* params, float* grads, float* loss);
int gnet_streams_debug(optmc_ctx* ctx, unsigned long long seed, int epoch, int step, double dropout, long long n_rows,
                       long long* perm_out, const unsigned int* row_ids, long long n_ids, unsigned int* keep_out);
int mlp_init_params_host(int H, unsigned long long seed, int date, float* out);
int mlp_grad_debug(optmc_ctx* ctx, int H, long long n, const float* xs, const float* ys, const float* params, float* grads,
                   float* cont);

// european.cu
int launch_european_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                          int32_t N, int32_t dtype, int32_t n_options, const double* K, const double* T,
                          const int32_t* is_put, const int32_t* stream_id, optmc_european_result* results,
                          const int32_t* N_opt = nullptr, const double* S0_opt = nullptr);
int launch_european_slab(optmc_ctx* ctx, const void* ST, int64_t M, int32_t dtype, double K, double r, double T,
                         int32_t is_put, optmc_european_result* out);

}  // namespace optmc
