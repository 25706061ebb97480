/*
 * optmc.h -- C ABI of the B200-native American-option Monte Carlo engine (liboptmc.so).
 *
 * Drop-in boundary for the data-parallel hot path of Levicoz/Options-model:
 * path simulation + Longstaff-Schwartz backward induction (+ the European payoff reduction).
 * The reference has NO FFI of its own (it is pure Python); each entry point below names the
 * reference function it replaces (file:line relative to the reference tree;
 * om3 = options_model_3/options_model_3.py, om3gpu = options_model_3/option_model_3_gpu.py,
 * om2 = options_model_2.py, hc = options_model_3/heston_calibration.py).
 *
 * Conventions
 *  - Every function returns OPTMC_OK (0) or a negative error class; the message is available from
 *    optmc_last_error() (thread-local).  OPTMC_EINVAL maps to the reference's ValueError
 *    (om3:447-452), everything else to RuntimeError.
 *  - "dev" pointers are CUDA device pointers owned by the CALLER (e.g. torch tensor.data_ptr()).
 *    "host" pointers are ordinary host memory owned by the caller.  The library owns only the
 *    opaque context (stream, workspace, exchange slots).
 *  - Path slabs are STEP-MAJOR: S[(N+1)][ld], row t = exercise date t, S[0][*] = S0 -- the layout of
 *    the reference's S array (om3:477, om3:217).  Columns [0, M/2) are the +Z paths, [M/2, M) the
 *    antithetic -Z paths (om3:476, om3:225-226).
 *  - A context is bound to one device and is not thread-safe.  All work is issued on the context's
 *    stream; only the *_fetch / host-result calls synchronise.
 *  - There is no CPU fallback: without a CUDA device optmc_ctx_create fails with OPTMC_ECUDA.
 */
#ifndef OPTMC_H
#define OPTMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPTMC_ABI_VERSION 1

enum optmc_status {
  OPTMC_OK = 0,
  OPTMC_EINVAL = -1,       /* bad argument: ValueError in the compat layer */
  OPTMC_ECUDA = -2,        /* CUDA runtime / launch failure */
  OPTMC_ENOMEM = -3,       /* workspace allocation failed */
  OPTMC_EUNSUPPORTED = -4  /* valid request this build cannot serve */
};

enum optmc_dtype { OPTMC_F32 = 0, OPTMC_F64 = 1 };

enum optmc_model { OPTMC_MODEL_GBM = 0, OPTMC_MODEL_HESTON = 1 };

enum optmc_scheme {
  OPTMC_SCHEME_GBM_LOG_EULER = 0,        /* om3:473-480  S *= exp(drift + diffusion Z), per step */
  OPTMC_SCHEME_GBM_LOGSPACE = 1,         /* om3gpu:150-185 cumulative log-space sum, exp at the end */
  OPTMC_SCHEME_HESTON_REF_ABSORB = 2,    /* om3:228-233 / om3gpu:218-225 absorption Euler (the reference's scheme) */
  OPTMC_SCHEME_HESTON_FULL_TRUNC = 3,    /* Lord et al. full truncation (north-star scheme; not in the reference) */
  OPTMC_SCHEME_HESTON_REF_CALIB = 4,     /* hc:240-255 arithmetic Euler on S, variance floored at 1e-8 */
  OPTMC_SCHEME_HESTON_QE = 5             /* Andersen (2008) quadratic-exponential (north-star scheme; not in the reference) */
};

enum optmc_basis {
  OPTMC_BASIS_POLY2 = 2, /* [1, x, x^2], x = S/K : first three reference features (om3:112-115) */
  OPTMC_BASIS_POLY3 = 3, /* [1, x, x^2, x^3]     : first four reference features */
  OPTMC_BASIS_REF7 = 7   /* all seven reference features (om3:105-121): the basis of the global fits (optmc_lsm_global,
                            optmc_lsm_gnet); in a per-date fit their span is that of POLY3, which is what runs */
};

/* LSM loop semantics.  REFERENCE = STICKY | REF_DISCOUNT reproduces om3:616-651. */
#define OPTMC_SEM_STICKY_MASK 1u  /* itm = payoff>0 && !exercised (om3:621); exercised is sticky (om3:649) */
#define OPTMC_SEM_REF_DISCOUNT 2u /* N-1 discounts, value at time dt (om3:619-620,651) instead of N */
#define OPTMC_SEM_REFERENCE (OPTMC_SEM_STICKY_MASK | OPTMC_SEM_REF_DISCOUNT)
#define OPTMC_SEM_TEXTBOOK 0u

/* Which sweep implementation to run (results are identical; this exists for tests and profiling). */
enum optmc_sweep_impl {
  OPTMC_SWEEP_AUTO = 0,     /* resident persistent kernel when the slab slice fits on chip, else split */
  OPTMC_SWEEP_RESIDENT = 1, /* one cooperative launch for all dates; cash-flows live in registers */
  OPTMC_SWEEP_SPLIT = 2     /* three small launches per date; cash-flows in HBM */
};

typedef struct optmc_ctx optmc_ctx;

typedef struct optmc_model_params {
  int32_t model;  /* optmc_model */
  int32_t scheme; /* optmc_scheme */
  double S0, r, T;
  double sigma;                     /* GBM volatility */
  double v0, kappa, theta, xi, rho; /* Heston (xi = vol-of-vol; hc.HestonParams calls it sigma) */
} optmc_model_params;

typedef struct optmc_rng_params {
  uint64_t seed;   /* Philox4x32-10 key */
  uint64_t stream; /* Philox counter word 3 (low 32 bits) -- one independent stream per option */
  /* Optional external normals (device pointers), step-major [N][M/2] (or [N][M] when antithetic = 0),
   * row t-1 drives step t (om3:475-480; om3:223-224).  NULL => Philox in-register generation.
   * GBM uses z1 only.  z_dtype is the element type of the z buffers. */
  const void* z1_dev;
  const void* z2_dev;
  int32_t z_dtype;     /* optmc_dtype */
  int32_t antithetic;  /* 1: [Z, -Z] column layout (om3:476); 0: independent columns (om3gpu:173) */
  int64_t pair_offset; /* global index of local pair 0: path-sharding across GPUs keeps Philox counters global */
} optmc_rng_params;

typedef struct optmc_lsm_params {
  double K, r, T;
  int32_t is_put;      /* 1 put, 0 call (om3:376-380) */
  int32_t basis;       /* optmc_basis */
  uint32_t semantics;  /* OPTMC_SEM_* */
  int32_t impl;        /* optmc_sweep_impl */
} optmc_lsm_params;

/* Host-side result block.  Array members may be NULL (not copied back). */
typedef struct optmc_lsm_result {
  double price;      /* mean cash-flow (om3:651), times one more discount under TEXTBOOK */
  double stderr_;    /* sample std / sqrt(M) */
  int64_t n_paths;
  int32_t impl_used; /* optmc_sweep_impl actually run */
  int32_t n_launches;/* kernels launched by this call */
  double* betas;     /* host [N+1][p], NaN rows where no regression was solved */
  double* boundary;  /* host [N+1]: put -> max exercised S, call -> min exercised S, NaN if none */
  int64_t* ex_count; /* host [N+1]: paths newly exercised at date t */
  int64_t* n_itm;    /* host [N+1]: regression rows at date t */
} optmc_lsm_result;

typedef struct optmc_european_result {
  double mean;    /* discounted payoff mean: exp(-rT) * mean(payoff(S_T))  (om3:425, hc:274-275) */
  double stderr_; /* sample std / sqrt(n) (om3:61-62) */
  int64_t n_paths;
} optmc_european_result;

/* ---- context ------------------------------------------------------------------------------- */
int optmc_abi_version(void);
const char* optmc_last_error(void);
int optmc_ctx_create(int device, optmc_ctx** out);
int optmc_ctx_destroy(optmc_ctx* ctx);
/* Use a caller-provided cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0/NULL => the
 * context's own stream. */
int optmc_ctx_set_stream(optmc_ctx* ctx, void* cuda_stream);
int optmc_ctx_synchronize(optmc_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t optmc_ctx_launch_count(optmc_ctx* ctx);
/* Device time of the path kernel(s) and the sweep kernel(s) of the last optmc_price_american /
 * optmc_price_american_batch call on this context, measured with CUDA events on the context's stream
 * (bench.py's per-kernel roofline). */
int optmc_ctx_kernel_times(optmc_ctx* ctx, double* paths_ms, double* sweep_ms);
/* Device properties the host layer needs for grid sizing / reporting: out[0]=SM count,
 * out[1]=L2 bytes, out[2]=max opt-in shared memory per block, out[3]=compute capability major*10+minor. */
int optmc_ctx_device_info(optmc_ctx* ctx, int64_t out[4]);
/* Device memory the library itself allocates (SURVEY 8b "workspace size is queryable"; the caller owns every
 * other buffer).  optmc_workspace_bytes: upper estimate, before the call, of what optmc_price_american_batch
 * (n_options >= 1; optmc_price_american is n_options = 1) will hold for M paths x N dates of `dtype` -- the
 * step-major slabs of one wave (at most one option per SM) plus cash-flows and per-date arrays; no context needed, no CUDA call.
 * optmc_ctx_workspace_bytes: what the context holds right now (workspaces only grow). */
int optmc_workspace_bytes(int64_t M, int32_t N, int32_t dtype, int32_t n_options, int64_t* bytes);
int optmc_ctx_workspace_bytes(optmc_ctx* ctx, int64_t* bytes);

/* ---- path-sharded sweep over the GPUs of one box (no counterpart in the reference) ----
 * One process per GPU.  The per-date exchange of the Gram totals happens INSIDE the persistent sweep kernel
 * through peer-mapped memory over NVLink (no host-launched collective on the data path):
 *   1. every rank: optmc_comm_export -> a 64-byte CUDA-IPC handle of its exchange slots;
 *   2. the host plumbing all-gathers the handles (torch.distributed / MPI / a pipe -- not this library's job);
 *   3. every rank: optmc_comm_init(rank, nranks, handles[nranks][64]) maps the peers' slots;
 *   4. every rank, in the same order: optmc_lsm_poly_sharded on its block of paths.  Every rank returns the
 *      same price / stderr / betas / n_itm (bit-identical: the totals are integer sums); ex_count and boundary
 *      cover the rank's own paths (sum / max them on the host if wanted).
 * At most 8 ranks; every rank's block must fit the persistent kernel (OPTMC_EUNSUPPORTED otherwise); a peer that
 * never launches makes the others give up after a few seconds with OPTMC_ECUDA instead of hanging the GPU. */
#define OPTMC_COMM_HANDLE_BYTES 64
int optmc_comm_export(optmc_ctx* ctx, void* handle_out);
int optmc_comm_init(optmc_ctx* ctx, int32_t rank, int32_t nranks, const void* handles);
int optmc_comm_finalize(optmc_ctx* ctx);
int optmc_lsm_poly_sharded(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M_local, int64_t M_total, int32_t N,
                           int32_t dtype, const optmc_lsm_params* lp, optmc_lsm_result* out);

/* ---- path simulation (replaces om3:473-480, om3:211-251, om3gpu:117-248, hc:204-257) ---------- */
/* S_dev: [(N+1)][ld] of `dtype`; ld >= M.  V_dev (Heston only, may be NULL): same shape, the variance
 * slab the reference also builds (om3:218, hc:217).  M must be even when rng->antithetic. */
int optmc_paths_gbm(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M, int32_t N,
                    int32_t dtype, void* S_dev, int64_t ld);
int optmc_paths_heston(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                       int32_t N, int32_t dtype, void* S_dev, void* V_dev, int64_t ld);
/* Local-volatility paths (om3:263-333, om3gpu:250-298): the implied-volatility network is evaluated inside the step.
 * `weights` = the ImprovedIVNetwork (nniv:109-155) state_dict flattened in its own order, host fp32:
 *   input_proj.weight [H][2], input_proj.bias [H], then per hidden layer l: Linear weight [H][H], bias [H],
 *   LayerNorm weight [H], bias [H]; output.weight [H], output.bias [1]   (3 H + L (H^2 + 3 H) + H + 1 values).
 * m_scale / tau_scale: the fitted DataScaler's stds (the pricer divides by them without centring, om3:285-288);
 * epsilon: the network's output clamp (TrainingConfig.epsilon, 1e-4); K: the strike entering ln(K / S).
 * Uses mp->S0, mp->r, mp->T; rng as for optmc_paths_gbm (external normals: z1_dev [N][M/2]). */
typedef struct optmc_ivnet {
  int32_t hidden;      /* H: 32 or 64 (TrainingConfig.hidden_dim default 64) */
  int32_t layers;      /* L: TrainingConfig.num_hidden_layers, default 4 */
  int32_t n_weights;   /* length of weights, checked against H and L */
  float epsilon;
  const float* weights;
  double m_scale, tau_scale, K;
} optmc_ivnet;
int optmc_paths_localvol(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, const optmc_ivnet* net,
                         int64_t M, int32_t N, int32_t dtype, void* S_dev, int64_t ld);
/* IVModel.get_volatility_batch (om3:277-298): sigma_dev[i] = max(net(ln(K / S_dev[i]) / m_scale, tau / tau_scale), epsilon,
 * 1e-6) for n device spots (fp64 in and out, like the reference's numpy arrays). */
int optmc_ivnet_sigma(optmc_ctx* ctx, const optmc_ivnet* net, double tau, const double* S_dev, int64_t n, double* sigma_dev);
/* The standard normals the Philox path kernels consume, written step-major [N][M/2] (test aid: feed
 * them to the oracle).  which = 0 -> z1 (asset), 1 -> z2 (variance, Heston only). */
int optmc_philox_normals(optmc_ctx* ctx, const optmc_rng_params* rng, int32_t model, int64_t M, int32_t N,
                         int32_t which, int32_t dtype, void* Z_dev);
/* Philox4x32-10 known-answer hook: runs n blocks on the device.  ctr[n][4], key[n][2] -> out[n][4] (host). */
int optmc_philox_kat(optmc_ctx* ctx, int32_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out);

/* ---- LSM backward induction (replaces the loop of om3:615-651 / om3:485-500 / om2:278-310 with the
 *      polynomial regressor of SURVEY.md 8(c)) ------------------------------------------------------- */
int optmc_lsm_poly(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                   const optmc_lsm_params* lp, optmc_lsm_result* out /* NULL: asynchronous, fetch later */);
/* Synchronise and copy the last sweep's results to the host. */
int optmc_lsm_fetch(optmc_ctx* ctx, optmc_lsm_result* out);
/* Number of paths whose final cash-flow is exactly zero (om1:168 `zero_prob = mean(cashflows == 0)`).  Valid after a
 * sweep that keeps its cash-flows in device memory: optmc_lsm_poly with impl = SPLIT, optmc_lsm_mlp. */
int optmc_lsm_zero_cashflows(optmc_ctx* ctx, int64_t* count);

/* ---- LSM with ONE global regression over all (date, path) rows: the structure of the reference's v3 pricer
 *      (om3:482-651 / om3gpu:695-833) with its network replaced by linear least squares on the seven reference
 *      features.  Pass 1 (om3:485-516): targets are the discounted TERMINAL payoffs of every in-the-money path at
 *      every date (`exercised` is never set in pass 1).  The target / feature z-scoring of om3:550-563 is an affine
 *      change of variables, under which a linear model with intercept is invariant, so it is not materialised.
 *      Pass 2 (om3:615-651): the usual loop (sticky mask, strict '>', N-1 discounts) with the global model.
 *      Both passes stream the slab once and are independent across paths: no per-date regression, no grid
 *      synchronisation. */
typedef struct optmc_global_result {
  double price, stderr_;
  int64_t n_paths;
  int64_t n_rows;    /* regression rows = sum over dates of in-the-money paths */
  int32_t n_launches;
  int32_t rank;      /* columns kept by the guarded solve (redundant columns get beta = 0) */
  double beta[7];    /* coefficients of [1, x, x^2, x^3, max(x-1,0), sqrt(tau), x sqrt(tau)] (om3:112-121) */
  double* boundary;  /* host [N+1] or NULL: put -> max exercised S, call -> min exercised S, NaN if none */
  int64_t* ex_count; /* host [N+1] or NULL */
} optmc_global_result;

int optmc_lsm_global(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                     const optmc_lsm_params* lp /* basis ignored (REF7) */, optmc_global_result* out);

/* ---- per-date neural-network LSM: the reference's v1/v2 loop (om2:277-310, om15:145-186, om1:107-150).  At every
 *      exercise date a fresh ContNet (1 -> hidden -> hidden -> 1, ReLU; om2:114-126) is trained for `epochs`
 *      full-batch Adam steps on the standardised prices of the live paths and its in-sample prediction is the
 *      continuation value.  Initial weights come from Philox (torch's default Linear range), so prices agree with
 *      the reference statistically; fed the same initial weights (optmc_mlp_init_params) the fit is reproducible
 *      to fp32 rounding.  Results through optmc_lsm_result (betas are NaN). */
typedef struct optmc_mlp_params {
  int32_t hidden;  /* 32 (om2 default nn_hidden; fp32 CUDA cores) or 128 (om3 default nn_hidden; bf16 tcgen05) */
  int32_t epochs;  /* om2 default nn_epochs = 10 */
  double lr;       /* om2 default nn_lr = 1e-3 */
  uint64_t seed;   /* Philox key of the per-date initial weights */
} optmc_mlp_params;

int optmc_lsm_mlp(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                  const optmc_lsm_params* lp /* basis ignored */, const optmc_mlp_params* np, optmc_lsm_result* out);
/* The initial parameters of date `date`: host out[3 H + H^2 + H + 1] in the order w1[H] b1[H] W2[H][H] b2[H] w3[H] b3.
 * Returns the parameter count, or a negative status. */
int optmc_mlp_init_params(int32_t hidden, uint64_t seed, int32_t date, float* out);
/* Test aid: one full-batch gradient evaluation of the ContNet on n host rows (xs, ys) at host `params`:
 * grads[P] = d(mean squared error)/d(params), cont[n] = forward output.  hidden = 32 runs the fp32 CUDA-core
 * kernels, hidden = 128 the bf16 tcgen05 kernels (fp32 accumulation in tensor memory). */
int optmc_mlp_grad_debug(optmc_ctx* ctx, int32_t hidden, int64_t n, const float* xs, const float* ys, const float* params,
                         float* grads, float* cont);

/* Out-of-sample exercise (not in the reference, whose continuation is in-sample): price the policy
 * "exercise at date t iff payoff(S) > sum_i betas[t][i] (S/K)^i" -- betas = host [(N+1)][p] exactly as optmc_lsm_poly
 * returns them (p = 3 / 4 for POLY2 / POLY3; NaN rows = no exercise at that date) -- on a DIFFERENT slab in one
 * streaming pass.  Semantics flags as for the sweep (sticky mask, N-1 discounts).  out->betas (if given) echoes the
 * input, out->n_itm is zero. */
int optmc_lsm_apply_policy(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                           const optmc_lsm_params* lp, const double* betas, optmc_lsm_result* out);

/* ---- global network LSM: the reference's v3 algorithm with its own regressor (om3:482-651, om3gpu:695-833) ----
 * Pass 1 collects the in-the-money rows of every date (features om3:105-121, targets = discounted terminal
 * payoffs), the features and the target are z-scored (om3:550-563), ONE SingleLSMNet(7, 128, 3) (om3:85-103) is
 * trained on all rows by mini-batch Adam (om3:565-613), pass 2 decides with the network (om3:615-651).  The 128x128
 * layers, forward and backward, run on the tcgen05 tensor cores (bf16 operands, fp32 accumulation).  Defaults of
 * the two reference variants:
 *   CPU  (om3:565-613):     batch 256, Adam + L2 weight decay 1e-5, ReduceLROnPlateau(patience 5, factor 0.5,
 *                           min_lr 1e-6), <= 25 epochs, early stop after 8 epochs without a 1e-6 improvement of
 *                           the mean training loss, best weights restored, population std of the target;
 *   GPU  (om3gpu:740-798):  batch <= 8192, AdamW (decoupled) 1e-4, no scheduler, patience 3, sample std.
 * Initial weights (torch's default Linear range), the per-epoch shuffle and the dropout masks come from counter-based
 * generators keyed by `seed`: runs are reproducible, and agree with the reference statistically (its streams are
 * torch's global RNG). */
typedef struct optmc_gnet_params {
  int32_t hidden;             /* 128 (nn_hidden default, om3:347) */
  int32_t layers;             /* 3 hidden layers (SingleLSMNet num_layers) */
  int32_t epochs;             /* nn_epochs, default 25 */
  int32_t batch;              /* 256 (CPU variant) .. 8192 (GPU variant); <= 131072 */
  double lr;                  /* nn_lr, default 1e-3 */
  double weight_decay;        /* 1e-5 (Adam L2) / 1e-4 (AdamW) */
  int32_t decoupled_wd;       /* 0 = torch.optim.Adam(weight_decay), 1 = AdamW */
  int32_t sched_patience;     /* ReduceLROnPlateau patience; 0 = no scheduler */
  double sched_factor;        /* 0.5 */
  double min_lr;              /* 1e-6 */
  int32_t stop_patience;      /* early stopping patience (8 / 3); 0 = never */
  int32_t target_ddof;        /* 0 = population std (np.std, om3:552), 1 = sample std (torch.std, om3gpu:731) */
  double min_delta;           /* improvement threshold of the best-weights snapshot, 1e-6 (om3:599) */
  double dropout;             /* 0.1; realised as the nearest multiple of 1/256 */
  int32_t inference_dropout;  /* pass 2 with dropout active: 1 / 0; -1 = as the reference under reference semantics
                                 (the net is never switched to eval mode), off under textbook */
  int32_t per_date;           /* 0 = ONE network for all dates (om3:482-651).  1 = a FRESH SingleLSMNet(7, 128, 3) per exercise date,
                                 trained on that date's live rows against the current cash-flows and used in-sample -- the loop of
                                 om2:277-310 / om15:145-186 with om3's regressor (BASELINE config 3: "tensor-core fit per date").
                                 Features and target are z-scored with the date's own moments; `epochs` x ceil(rows / batch)
                                 optimiser steps per date (batch >= the date's rows = om2's full-batch steps); per-date initial
                                 weights / shuffles / masks derive from seed and the date; init_params is ignored; n_rows /
                                 epochs_run = totals over the dates, best_loss = mean over the fitted dates, final_params = the
                                 network of date 1.  A slab with one exercise date (N = 2) reproduces per_date = 0 bit for bit. */
  uint64_t seed;
  const float* init_params;   /* host [34177] (state_dict order) or NULL = fresh torch-default initialisation.  The torch-GPU
                                 file keeps ONE network across pricing calls (om3gpu:741-748): pass the previous call's
                                 final_params here to reproduce that warm start */
  float* final_params;        /* host [34177] or NULL: the weights used by the decision pass */
} optmc_gnet_params;

typedef struct optmc_gnet_result {
  double price, stderr_;
  int64_t n_paths, n_rows;    /* n_rows = regression rows collected in pass 1 */
  int32_t epochs_run, n_launches;
  double best_loss, final_lr; /* mean batch MSE (normalised target) of the restored weights; learning rate at the end */
  double* boundary;           /* host [(N+1)] or NULL */
  int64_t* ex_count;          /* host [(N+1)] or NULL */
} optmc_gnet_result;

int optmc_lsm_gnet(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                   const optmc_lsm_params* lp /* basis ignored */, const optmc_gnet_params* gp, optmc_gnet_result* out);
/* PATH-SHARDED global network LSM ("all-reduce of the 34 177 fp32 gradients per optimiser step"; the loop
 * being sharded is om3:565-613).  Every rank of an optmc_comm_init group calls this with its own block of paths
 * (M_local columns of S_dev) and the same parameters:
 *   - pass 1 runs on the rank's paths; the row counts and the fixed-point feature / target moments are gathered through
 *     peer memory, so every rank z-scores with the moments of ALL rows;
 *   - optimiser step b of an epoch covers positions [b batch, (b+1) batch) of the global order; rank r contributes the
 *     proportional slice of its own (shuffled) rows (optmc_gnet_shard_plan).  Each rank sums its tiles' partial
 *     gradients in a fixed order and PUSHES the 34 178 words (gradient + batch loss, {tag, fp32} pairs) into every
 *     peer's memory over NVLink; the optimiser kernel polls its own memory and adds the ranks' vectors in rank order:
 *     weights, losses, scheduler and early-stopping decisions are bit-identical on every rank, with no host-launched
 *     collective on the data path;
 *   - pass 2 runs on the rank's paths; the fixed-point (sum, sum^2) of the values is gathered the same way.
 * Every rank returns the same price / stderr / best_loss; n_rows = rows of all ranks; n_paths = M_total; ex_count and
 * boundary cover the rank's own paths.  With a one-rank group the result is bit-identical to optmc_lsm_gnet.  A peer
 * that never launches makes the others give up after a few seconds with OPTMC_ECUDA. */
int optmc_lsm_gnet_sharded(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M_local, int64_t M_total, int32_t N,
                           int32_t dtype, const optmc_lsm_params* lp, const optmc_gnet_params* gp, optmc_gnet_result* out);
/* Host arithmetic of that split (no device needed): given every rank's row count, [*lo, *hi) = the positions of rank
 * `rank`'s own shuffled order that belong to optimiser step b (0-based) of an epoch, *global_rows = rows of all ranks in
 * that step (the gradient's normaliser).  Steps per epoch = ceil(sum(n_rows) / batch); over them every row of every rank
 * is used exactly once; with nranks == 1 the steps are optmc_lsm_gnet's batches. */
int optmc_gnet_shard_plan(const int64_t* n_rows, int32_t nranks, int32_t batch, int64_t b, int32_t rank, int64_t* lo,
                          int64_t* hi, int64_t* global_rows);
/* Test aid: mean-squared-error loss and its gradient for one batch of n <= 16384 host rows given as NORMALISED
 * features feat[n][7] and targets ys[n], at host params[34177] in the order W1[128][7] b1 W2[128][128] b2
 * W3[128][128] b3 w4[128] b4 (torch state_dict order), without dropout. */
int optmc_gnet_grad_debug(optmc_ctx* ctx, int64_t n, const float* feat, const float* ys, const float* params, float* grads,
                          float* loss);
/* Test aid: the counter-based shuffle and dropout streams of optmc_lsm_gnet, evaluated by the device code the training
 * and decision kernels use (the reference draws both from torch's global generator, om3:577 / om3:96-101; a paired
 * check against the torch restatement needs the engine's).  perm_out[i], i < n_rows: the row that position i of epoch
 * `epoch` (0-based) reads.  keep_out[(i * 3 + layer) * 4 + w]: bit b set = hidden unit 32 w + b of `layer` is kept for
 * row id row_ids[i], under optimiser step `step` (1-based, counted across epochs; the row id is the position in the
 * epoch's order) or, with step == 0, under the decision pass (row id = (uint32) path * 0x01000193 + date).  Host pointers. */
int optmc_gnet_streams_debug(optmc_ctx* ctx, uint64_t seed, int32_t epoch, int32_t step, double dropout, int64_t n_rows,
                             int64_t* perm_out, const uint32_t* row_ids, int64_t n_ids, uint32_t* keep_out);

/* Per-date building blocks for path-sharded multi-GPU sweeps: the host all-reduces `gram_dev`
 * (optmc_lsm_gram_len doubles) between the two calls. */
int optmc_lsm_gram_len(int32_t basis);
int optmc_lsm_begin(optmc_ctx* ctx, const void* S_dev, int64_t ld, int64_t M, int32_t N, int32_t dtype,
                    const optmc_lsm_params* lp);
int optmc_lsm_gram_date(optmc_ctx* ctx, int32_t t, double* gram_dev);
int optmc_lsm_update_date(optmc_ctx* ctx, int32_t t, const double* gram_dev);
/* sums_dev[0..2] = sum(cf), sum(cf^2), n over the local paths (all-reduce, then price = s0/n). */
int optmc_lsm_finish(optmc_ctx* ctx, double* sums_dev);

/* ---- fused host-facing calls (what the compat layer's pricer methods call) ------------------- */
/* price_american_enhanced_lsm (om3:439-651): paths into a context-owned slab + sweep.  Host scalars in,
 * host results out.  When `out` carries no per-date arrays (the reference's function returns the price alone) the
 * persistent sweep does not produce them either: a later optmc_lsm_fetch then reads NaN / none / 0 for every date.  Pass
 * out = NULL (fetch later) or the arrays to have them. */
int optmc_price_american(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                         int32_t N, int32_t dtype, const optmc_lsm_params* lp, optmc_lsm_result* out);
/* Batched American pricing: the S0 x maturity curve drivers (om3:697-713 compute_curve_for_S0, om3gpu:934-956
 * compute_multiple_S0_gpu_batch, om2:336-457) and strike x maturity grids call price_american_option once per
 * grid point; this entry point prices the whole list in a few launches.  Every option simulates its own
 * M paths (as the reference does) with Philox stream rng->stream + opts[i].stream; mp supplies r and the
 * model parameters (mp->S0 / mp->T are ignored).  Options are swept G at a time by one grouped persistent
 * kernel: the SMs are split into G groups, each sweeping one option with no cross-group synchronisation. */
typedef struct optmc_american_option {
  double S0, K, T;
  int32_t N;       /* exercise dates / time steps (may differ per option: om3:709) */
  int32_t is_put;
  uint64_t stream; /* added to rng->stream */
} optmc_american_option;

typedef struct optmc_price_result {
  double price, stderr_;
} optmc_price_result;

int optmc_price_american_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                               int32_t dtype, int32_t basis, uint32_t semantics, int32_t n_options,
                               const optmc_american_option* opts, optmc_price_result* results);

/* The same batch with optional extras (every pointer may be NULL):
 *   - per-date outputs of every option, as optmc_lsm_result holds them for one option (rows of ld_dates >= max N + 1
 *     entries; betas rows of ld_dates * 4 doubles, NaN where no regression was solved);
 *   - european[i][0..1]: mean / stderr of exp(-r T) payoff(S_T) over option i's OWN paths -- the European leg of
 *     price_american_with_control_variate (om3:653-677) with control_variate_same_paths=True, reduced inside the
 *     grouped sweep launch from the terminal row it stages anyway (no second pass over the slab);
 *   - shape[0..3] (out): compute threads per CTA, paths per thread, CTAs per option, options per grouped launch of
 *     the persistent sweep (all 0 when the batch ran through the split sweep);
 *   - M_total > 0: PATH-SHARDED batch.  Every rank of an optmc_comm_init group calls this with its own
 *     block of M paths of EACH option (rng->pair_offset = the block's first antithetic pair, M_total = paths per option
 *     over all ranks, the same options in the same order on every rank, one N for the whole batch); the per-date totals
 *     of every option are exchanged inside the grouped sweep kernel over NVLink peer memory, prices / betas are those
 *     of the whole option and identical on every rank.  A failure (fixed-point range, NaN) is propagated through the
 *     exchange itself, so every rank returns the same status. */
typedef struct optmc_batch_extras {
  int64_t* ex_count;  /* host [n_options][ld_dates] */
  double* boundary;   /* host [n_options][ld_dates] */
  double* betas;      /* host [n_options][ld_dates][4] */
  int64_t* n_itm;     /* host [n_options][ld_dates] */
  int32_t ld_dates;
  int32_t reserved;
  double* european;   /* host [n_options][2] */
  int32_t shape[4];
  int64_t M_total;
} optmc_batch_extras;

int optmc_price_american_batch_ex(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                                  int32_t dtype, int32_t basis, uint32_t semantics, int32_t n_options,
                                  const optmc_american_option* opts, optmc_price_result* results,
                                  optmc_batch_extras* extras);

/* ---- quasi-Monte-Carlo draws (no counterpart in the reference, which draws PCG64 pseudo-random normals,
 * om3:223-224,475) ----
 * Sobol' points (Joe-Kuo 2008 direction numbers, first 512 dimensions = factors x N) with an optional random digital
 * shift (digital_shift: host [factors * N] 32-bit words, NULL = the plain sequence) and an optional Brownian-bridge
 * ordering of the dimensions.  Writes step-major normals Z1 (and Z2 for factors = 2) [N][M/2] of `dtype` that the path
 * entry points accept as external draws (optmc_rng_params.z1_dev / z2_dev, antithetic layout); point index = antithetic
 * pair index + pair_offset + 1, so shards of a path set generate their own blocks. */
int optmc_qmc_normals(optmc_ctx* ctx, int64_t M, int32_t N, int32_t factors, int32_t brownian_bridge, int64_t pair_offset,
                      const uint32_t* digital_shift, int32_t dtype, void* Z1_dev, void* Z2_dev);
/* The bridge construction order the kernel uses (test aid): step s sets row idx[s] (holding W(idx+1)) to
 * wl W(left) + wr W(right+1) + sd z_s, W(0) = 0.  Host arrays of length N. */
int optmc_qmc_bridge_schedule(int32_t N, int32_t* idx, int32_t* left, int32_t* right, double* wl, double* wr, double* sd);

/* The same fused kernel with per-option spot and step count: the INDEPENDENT European leg of a whole S0 x maturity
 * curve with the control variate on (om3:653-677 inside om3:697-713: one price_european_streaming per grid point,
 * steps = max(10, min(130, ceil(days)))) as one launch.  S0 may be NULL (mp->S0 for every option). */
int optmc_price_european_grid(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                              int32_t dtype, int32_t n_options, const double* S0, const double* K, const double* T,
                              const int32_t* N, const int32_t* is_put, const int32_t* stream_id,
                              optmc_european_result* results);

/* price_european_streaming / price_european_gpu / HestonPricer.price_european_option
 * (om3:382-437, om3gpu:605-653, hc:259-281): paths are generated and reduced in registers, nothing is
 * stored.  n_options options share mp/rng except K[i], T[i], is_put[i]; option i uses Philox stream
 * rng->stream + stream_id[i] (options with equal ids see the same paths: "simulate once per unique T",
 * hc:289-306).  results: host [n_options]. */
int optmc_price_european_batch(optmc_ctx* ctx, const optmc_model_params* mp, const optmc_rng_params* rng, int64_t M,
                               int32_t N, int32_t dtype, int32_t n_options, const double* K, const double* T,
                               const int32_t* is_put, const int32_t* stream_id /* NULL: option i -> i */,
                               optmc_european_result* results);
/* European reduction over an existing terminal slab row (a12 on stored paths). */
int optmc_european_from_slab(optmc_ctx* ctx, const void* ST_dev, int64_t M, int32_t dtype, double K, double r,
                             double T, int32_t is_put, optmc_european_result* out);

/* ---- parity aid: create_regression_features (om3:105-121, om3gpu:342-377) ---------------------- */
/* F_dev: [n][7] row-major of `dtype`. */
int optmc_features_ref7(optmc_ctx* ctx, const void* S_dev, int64_t n, int32_t dtype, double K, double r, double T,
                        double t_current, void* F_dev);

#ifdef __cplusplus
}
#endif
#endif /* OPTMC_H */
