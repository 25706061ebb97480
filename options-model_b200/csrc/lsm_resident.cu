// lsm_resident.cu -- the persistent LSM sweep: ONE cooperative launch walks every exercise date.
//
// Data layout and flow (per CTA, one CTA per SM, 512 threads):
//   * the CTA owns a contiguous slice of paths; their cash-flows live in REGISTERS for the whole sweep
//     (PPT values per thread; the reference's `exercised` flag, om3:617/649, is the sign bit);
//   * each date's slice of the step-major price slab is staged into shared memory by the TMA engine
//     (1-D cp.async.bulk + mbarrier), 2-3 dates ahead of use, so HBM sees S exactly once;
//   * per date: masked Gram moments of the raw price in fp64 (thread) -> recursive-halving reduce-scatter
//     (warp) -> shared memory (block) -> grid-wide sum -> guarded LDL^T solve (warp 0) -> fused exercise
//     decision / cash-flow update.
//   * the grid-wide sum is ORDER-INDEPENDENT: every block total is split into two 48-bit fixed-point
//     chunks (optmc_math.cuh: fx_encode) and added with one integer `red` per chunk into a per-parity
//     accumulator word whose top 8 bits count arrivals.  Sum and completion flag are the same word, so
//     there is no fence, no second flag and no grid-wide barrier on the data path; one lane polls each
//     word.  Integer addition commutes => betas are bit-reproducible on every CTA and from run to run.
//
// Replaces the Python loop of om3:615-651 (= om3:485-500, om2:278-310) for the polynomial regressor of
// SURVEY.md 8(c).  Semantics flags: sticky mask (om3:621,649), N-1 discounts (om3:619-620,651).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "lsm_resident_kernel.cuh"

namespace optmc {

struct ResShape { int nt, ppt; };
#define X(NT_, PPT_) {NT_, PPT_},
static const ResShape kShapes[] = {OPTMC_RES_SHAPES(X)};
#undef X

// Slice the M paths over at most one CTA per SM and pick the smallest compiled (threads, paths-per-thread)
// shape that holds a slice.
static bool plan_resident(optmc_ctx* ctx, const SweepDesc& sw, ResPlan* p, std::string* why) {
  const size_t es = sw.dtype == OPTMC_F64 ? 8 : 4;
  if ((uintptr_t)sw.S % 16 != 0 || (sw.ld * es) % 16 != 0) { *why = "slab not 16-byte aligned"; return false; }
  if ((sw.M * es) % 16 != 0) { *why = "path count is not a multiple of 16 bytes (bulk copies move whole 16-byte units)"; return false; }
  if (ctx->cc < 90) { *why = "bulk async copy needs sm_90+"; return false; }
  const int ncta_cap = ctx->sm_count < kMaxResidentCtas ? ctx->sm_count : kMaxResidentCtas;
  long long ncta = (sw.M + 511) / 512;  // at least one path per thread before adding CTAs
  if (ncta > ncta_cap) ncta = ncta_cap;
  if (ncta < 1) ncta = 1;
  long long chunk = (sw.M + ncta - 1) / ncta;
  chunk = (chunk + 3) / 4 * 4;
  ncta = (sw.M + chunk - 1) / chunk;
  const ResShape* shape = nullptr;
  for (const ResShape& c : kShapes)
    if ((long long)c.nt * c.ppt >= chunk) { shape = &c; break; }
  if (!shape) { *why = "slice exceeds the register-resident capacity"; return false; }
  // every stage holds the full NT x PPT slot grid: the tail behind the slice is an out-of-the-money sentinel
  const size_t stride = ((size_t)shape->nt * shape->ppt * es + 127) / 128 * 128;
  const size_t avail = (size_t)ctx->max_smem_optin - 8192;  // static shared + slack
  int nstage = 3;
  if (stride * 3 > avail) nstage = 2;
  if (stride * 2 > avail) { *why = "slice exceeds shared memory"; return false; }
  p->ncta = (int)ncta; p->ppt = shape->ppt; p->nt = shape->nt; p->nstage = nstage; p->chunk = chunk;
  p->stage_stride = (unsigned int)stride; p->smem = stride * nstage;
  // sticky mask => only a small fraction of the paths is in the regression at any date: vote-skip passes.
  // Tuning aid: OPTMC_RES_SPARSE=0|1 overrides.
  p->sparse = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) != 0;
  if (const char* e = getenv("OPTMC_RES_SPARSE")) p->sparse = atoi(e) != 0;
  return true;
}

bool resident_eligible(optmc_ctx* ctx, const SweepDesc& sw, std::string* why) {
  ResPlan p;
  return plan_resident(ctx, sw, &p, why);
}

int sweep_resident(optmc_ctx* ctx) {
  SweepDesc& sw = ctx->sw;
  ResPlan p;
  std::string why;
  if (!plan_resident(ctx, sw, &p, &why)) { set_error("resident sweep unavailable: " + why); return OPTMC_EUNSUPPORTED; }
  int rc = sweep_reset_stats(ctx);  // also zeroes the exchange accumulators and the overflow flag
  if (rc) return rc;
  ResArgs a{};
  a.S = sw.S; a.ld = sw.ld; a.M = sw.M; a.chunk = p.chunk; a.N = sw.N; a.nstage = p.nstage;
  a.stage_stride = p.stage_stride;
  a.K = sw.lp.K; a.invK = 1.0 / sw.lp.K; a.disc = sw.disc; a.inv_disc = 1.0 / sw.disc; a.final_scale = sw.final_scale;
  {  // pass constants, exact in the storage type (see Store<>)
    const double sg = sw.lp.is_put ? -1.0 : 1.0;
    double Kcmp = sw.lp.K, Kh = sw.lp.K, Kl = 0.0;
    if (sw.dtype == OPTMC_F32) {
      // float threshold with (s < K) <=> (s < Kcmp) for every float s (puts); mirrored for calls
      float kf = (float)sw.lp.K;
      Kh = (double)kf;
      Kl = (double)(float)(sw.lp.K - Kh);
      if (sw.lp.is_put) { if ((double)kf < sw.lp.K) kf = nextafterf(kf, INFINITY); }
      else { if ((double)kf > sw.lp.K) kf = nextafterf(kf, -INFINITY); }
      Kcmp = (double)kf;
    }
    a.sgn = sg; a.kk = sg * Kcmp; a.c1 = -sg * Kh; a.c2 = -sg * Kl;
  }
  a.is_put = sw.lp.is_put; a.sticky = (sw.lp.semantics & OPTMC_SEM_STICKY_MASK) ? 1 : 0;
  a.xw = reinterpret_cast<unsigned long long*>(ctx->xchg); a.flags = ctx->d_flags;
  a.betas = ctx->d_betas; a.bnd = ctx->d_bnd; a.exc = ctx->d_exc; a.nitm = ctx->d_nitm; a.final_out = ctx->d_final;
  // Debug aid: OPTMC_TRACE=<file> dumps per-date phase clocks (SM cycles) of the first and last CTA.
  const char* trace_path = getenv("OPTMC_TRACE");
  long long* d_trace = nullptr;
  const size_t trace_n = (size_t)2 * (sw.N + 1) * 8;
  if (trace_path && *trace_path) {
    OPTMC_CUDA(cudaMalloc((void**)&d_trace, trace_n * sizeof(long long)));
    OPTMC_CUDA(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), ctx->stream));
    a.trace = d_trace;
  }
  if (sw.dtype == OPTMC_F64) rc = sw.deg == 2 ? launch_resident_f64_deg2(ctx, p, a) : launch_resident_f64_deg3(ctx, p, a);
  else rc = sw.deg == 2 ? launch_resident_f32_deg2(ctx, p, a) : launch_resident_f32_deg3(ctx, p, a);
  if (d_trace) {
    std::vector<long long> h(trace_n);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_trace);
    if (e == cudaSuccess) {
      if (FILE* f = fopen(trace_path, "w")) {
        fprintf(f, "# ncta=%d chunk=%lld nt=%d ppt=%d nstage=%d N=%d ; columns: cta t ph0..ph7 (clock64)\n", p.ncta,
                p.chunk, p.nt, p.ppt, p.nstage, sw.N);
        for (int c = 0; c < 2; ++c)
          for (int t = sw.N; t >= 1; --t) {
            fprintf(f, "%d %d", c, t);
            for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h[((size_t)c * (sw.N + 1) + t) * 8 + k]);
            fprintf(f, "\n");
          }
        fclose(f);
      }
    }
  }
  return rc;
}

}  // namespace optmc
