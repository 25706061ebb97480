/* Minimal C client of liboptmc.so: prices BASELINE config 1 (American put, GBM, 100 k paths x 50 exercise dates, quadratic
 * LSM) and config 2 (Heston, 1 M x 252) through the C ABI alone -- no Python, no torch.
 *
 *   gcc -I include examples/price_american.c -o price_american -L options-model_b200 -loptmc \
 *       -Wl,-rpath,$PWD/options-model_b200 && ./price_american
 */
#include <stdio.h>
#include <string.h>

#include "optmc.h"

int main(void) {
  optmc_ctx* ctx = NULL;
  if (optmc_ctx_create(0, &ctx) != OPTMC_OK) {
    fprintf(stderr, "no CUDA device: %s\n", optmc_last_error());
    return 2;
  }
  optmc_rng_params rng;
  memset(&rng, 0, sizeof(rng));
  rng.seed = 42;
  rng.antithetic = 1;

  optmc_model_params gbm;
  memset(&gbm, 0, sizeof(gbm));
  gbm.model = OPTMC_MODEL_GBM; gbm.scheme = OPTMC_SCHEME_GBM_LOG_EULER;
  gbm.S0 = 100.0; gbm.r = 0.05; gbm.T = 1.0; gbm.sigma = 0.2;
  optmc_lsm_params lp;
  memset(&lp, 0, sizeof(lp));
  lp.K = 100.0; lp.r = 0.05; lp.T = 1.0; lp.is_put = 1; lp.basis = OPTMC_BASIS_POLY2;
  lp.semantics = OPTMC_SEM_REFERENCE; lp.impl = OPTMC_SWEEP_AUTO;
  optmc_lsm_result res;
  memset(&res, 0, sizeof(res));
  if (optmc_price_american(ctx, &gbm, &rng, 100000, 50, OPTMC_F32, &lp, &res) != OPTMC_OK) {
    fprintf(stderr, "config 1 failed: %s\n", optmc_last_error());
    return 1;
  }
  printf("config1 price %.6f stderr %.6f launches %d\n", res.price, res.stderr_, res.n_launches);

  optmc_model_params hes = gbm;
  hes.model = OPTMC_MODEL_HESTON; hes.scheme = OPTMC_SCHEME_HESTON_REF_ABSORB;
  hes.v0 = 0.04; hes.kappa = 2.0; hes.theta = 0.04; hes.xi = 0.5; hes.rho = -0.7;
  if (optmc_price_american(ctx, &hes, &rng, 1000000, 252, OPTMC_F32, &lp, &res) != OPTMC_OK) {
    fprintf(stderr, "config 2 failed: %s\n", optmc_last_error());
    return 1;
  }
  double paths_ms = 0.0, sweep_ms = 0.0;
  optmc_ctx_kernel_times(ctx, &paths_ms, &sweep_ms);
  printf("config2 price %.6f stderr %.6f paths %.3f ms sweep %.3f ms\n", res.price, res.stderr_, paths_ms, sweep_ms);
  /* the reference's ValueError texts travel through optmc_last_error */
  if (optmc_price_american(ctx, &hes, &rng, 0, 252, OPTMC_F32, &lp, &res) != OPTMC_EINVAL) return 1;
  printf("error path: %s\n", optmc_last_error());
  optmc_ctx_destroy(ctx);
  return 0;
}
