// TEST INFRASTRUCTURE ONLY -- never linked into liboptmc.so.
// Compiles the per-path arithmetic of csrc/optmc_math.cuh for the host (the header is __host__ __device__)
// so the CPU test-suite can check Philox, the step functions, the Gram moments and the guarded solve against
// the numpy oracle without a GPU.  Kernel scaffolding (indexing, synchronisation) is only tested on the GPU.
#include "../../options-model_b200/csrc/optmc_math.cuh"

using namespace optmc;

extern "C" {

void hs_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  Philox4 p = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = p.v[i];
}

// step-major paths [(N+1)][M] from external normals [N][M/2], antithetic layout; scheme: 2 absorb, 3 full trunc, 4 calib, 5 QE
void hs_heston_paths(int scheme, double S0, double v0, double r, double T, double kappa, double theta, double xi,
                     double rho, long M, int N, const double* z1, const double* z2, double* S, double* V) {
  HestonConsts<double> c;
  c.dt = T / N; c.sqrt_dt = sqrt(c.dt); c.r = r; c.kappa = kappa; c.theta = theta; c.xi = xi; c.rho = rho;
  c.rho_c = sqrt(1.0 - rho * rho);
  long Mh = M / 2;
  const QeConsts<double> q = qe_consts(c);
  for (long j = 0; j < M; ++j) { S[j] = S0; if (V) V[j] = v0; }
  for (long j = 0; j < Mh; ++j) {
    double sp = S0, sm = S0, vp = v0, vm = v0;
    for (int t = 1; t <= N; ++t) {
      double a = z1[(long)(t - 1) * Mh + j], b = z2[(long)(t - 1) * Mh + j];
      if (scheme == 2) { heston_absorb_step<double>(sp, vp, a, b, c); heston_absorb_step<double>(sm, vm, -a, -b, c); }
      else if (scheme == 3) { heston_fulltrunc_step<double>(sp, vp, a, b, c); heston_fulltrunc_step<double>(sm, vm, -a, -b, c); }
      else if (scheme == 5) { heston_qe_step<double>(sp, vp, a, b, q); heston_qe_step<double>(sm, vm, -a, -b, q); }
      else { heston_calib_step<double>(sp, vp, a, b, c); heston_calib_step<double>(sm, vm, -a, -b, c); }
      S[(long)t * M + j] = sp; S[(long)t * M + Mh + j] = sm;
      if (V) { V[(long)t * M + j] = vp; V[(long)t * M + Mh + j] = vm; }
    }
  }
}

void hs_gbm_paths(double S0, double r, double sigma, double T, long M, int N, const double* z, double* S) {
  GbmConsts<double> c;
  double dt = T / N;
  c.drift = (r - 0.5 * sigma * sigma) * dt; c.diffusion = sigma * sqrt(dt);
  long Mh = M / 2;
  for (long j = 0; j < M; ++j) S[j] = S0;
  for (long j = 0; j < Mh; ++j) {
    double sp = S0, sm = S0;
    for (int t = 1; t <= N; ++t) {
      double a = z[(long)(t - 1) * Mh + j];
      sp = gbm_step<double>(sp, a, c); sm = gbm_step<double>(sm, -a, c);
      S[(long)t * M + j] = sp; S[(long)t * M + Mh + j] = sm;
    }
  }
}

// moments over (x, y) rows then the guarded solve; returns 1 when a beta was produced
int hs_fit(int deg, long n, const double* x, const double* y, double* mom_out, double* beta) {
  if (deg == 2) {
    double acc[Moments<2>::Q] = {0};
    for (long i = 0; i < n; ++i) moments_accumulate<2>(acc, x[i], y[i]);
    for (int q = 0; q < Moments<2>::Q; ++q) mom_out[q] = acc[q];
    return solve_poly<2>(acc, beta) ? 1 : 0;
  }
  double acc[Moments<3>::Q] = {0};
  for (long i = 0; i < n; ++i) moments_accumulate<3>(acc, x[i], y[i]);
  for (int q = 0; q < Moments<3>::Q; ++q) mom_out[q] = acc[q];
  return solve_poly<3>(acc, beta) ? 1 : 0;
}

double hs_poly_eval(int deg, const double* beta, double x) { return deg == 2 ? poly_eval<2>(beta, x) : poly_eval<3>(beta, x); }

void hs_normal2_f32(uint32_t a, uint32_t b, float* out) { Real<float>::normal2(a, b, out[0], out[1]); }
void hs_normal2_f64(uint32_t a, uint32_t b, double* out) { Real<double>::normal2(a, b, out[0], out[1]); }

void hs_features(double S, double K, double tau_sqrt, double* f) { features_ref7<double>(S, K, tau_sqrt, f); }

// fixed-point exchange: sum n doubles through fx_encode + 64-bit integer accumulation (arrival counts in the
// top 8 bits, exactly as the persistent sweep does) and decode.  Returns 0 on overflow, else 1.
int hs_fx_sum(long n, const double* v, double* out, unsigned long long* words) {
  unsigned long long whi = 0, wlo = 0;
  int ok = 1;
  for (long i = 0; i < n; ++i) {
    unsigned long long hi, lo;
    if (!fx_encode(v[i], hi, lo)) ok = 0;
    whi += (1ull << kFxCountShift) | hi;
    wlo += (1ull << kFxCountShift) | lo;
  }
  words[0] = whi; words[1] = wlo;
  *out = fx_decode(whi & kFxValueMask, wlo & kFxValueMask, (int)n);
  return ok;
}
}
