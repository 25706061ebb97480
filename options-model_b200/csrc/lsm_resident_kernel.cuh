// lsm_resident_kernel.cuh -- device code of the persistent LSM sweep (see lsm_resident.cu for the design).
// Included by the per-precision translation units lsm_resident_f32.cu / lsm_resident_f64.cu.
#pragma once
#include <math.h>

#include "optmc_device.cuh"
#include "optmc_internal.h"
#include "optmc_math.cuh"

namespace optmc {

// One option of a (possibly grouped) sweep.  A launch sweeps G independent options at once: CTAs
// [g * cpg, (g + 1) * cpg) form group g, which owns its slab, its exchange accumulators and its outputs;
// groups never synchronise with each other.  All groups of a launch share the storage type, the basis, the
// semantics, the path count and hence the CTA shape; the number of exercise dates may differ.
struct ResGroup {
  const void* S;
  long long ld, M, chunk;
  long long M_total;          // paths of the option over all ranks (= M unless the sweep is path-sharded)
  int N, is_put;
  double K, invK, disc, inv_disc, final_scale;
  double sgn, kk, c1, c2;     // storage-precision pass constants (see Store<>); exact in the storage type
  unsigned long long* xw;     // exchange accumulators [2][kXchgWords][kXchgStride]
  int* flags;                 // [0] = exchange overflow
  double* betas;              // [(N+1)][kMaxBeta] or NULL
  unsigned long long* bnd;    // [(N+1)] or NULL
  unsigned long long* exc;    // [(N+1)] or NULL
  long long* nitm;            // [(N+1)] or NULL
  double* final_out;          // [8]: price, stderr, sum, sum^2, European mean, European stderr (want_eu), -, -
};

// Path-sharded sweep over several GPUs: every rank runs this kernel on its block of paths and the
// per-date totals are exchanged INSIDE the kernel through peer-mapped memory (NVLink), not by a host-launched
// collective.  slots[p] is rank p's slot array [2 parities][kCommMaxRanks][kXchgWords] (slots[rank] is local,
// the others are CUDA-IPC mappings).  Per exchange g (a counter that keeps running across launches, identical
// on every rank): CTA 0 of rank r, once its local accumulators are complete, stores its 16 words -- the
// 56-bit fixed-point payload with the tag g & 255 in the top byte -- into slot [g & 1][r] of EVERY rank; all
// CTAs of a rank then poll their own rank's slots until the nranks tags match and add the payloads as
// integers, so every CTA of every rank obtains bit-identical totals.  A slot is rewritten two exchanges later,
// which needs all ranks' pushes of exchange g + 1, which every CTA issues only after it read exchange g.
constexpr int kCommSlotWords = 2 * kCommMaxRanks * kXchgWords;  // slot block of one option (group) on one rank
constexpr unsigned int kCommSpinLimit = 1u << 22;  // ~seconds: a peer that never launched must not hang the GPU
struct ResComm {
  int nranks = 1, rank = 0;
  unsigned int g0 = 0;        // running exchange counter at launch
  long long M_total = 0;      // paths of the option over all ranks
  unsigned long long* slots[kCommMaxRanks] = {};
};

struct ResArgs {
  ResComm comm;
  ResGroup one;               // the group of a single-option launch (groups == NULL)
  const ResGroup* groups;     // device array [G] for grouped launches
  int cpg;                    // CTAs per group
  int nstage, sticky;
  int want_eu;                // speculative kernel: also reduce the European payoff of the terminal row (control variate)
  unsigned int stage_stride;  // bytes between stages in shared memory
  long long* trace;           // optional [2][(N+1)][kTraceCols] phase clocks of the first and last CTA (OPTMC_TRACE)
  void* spill;                // speculative kernel: candidate-list overflow, [CTAs][warps][PPT * 32] entries of 3 values
};

template <int QN> struct Pow2 { static constexpr int v = QN <= 2 ? 2 : QN <= 4 ? 4 : QN <= 8 ? 8 : 16; };

// ---- storage-type helpers ------------------------------------------------------------------------------
// Put and call share one instruction stream: with sgn = -1 (put) / +1 (call)
//   in the money   <=>  sgn * s > kk            (kk = sgn * Kcmp; products with +-1 are exact)
//   payoff          =   fma(sgn, s, c1) + c2    (c1 = -sgn * Kh, c2 = -sgn * Kl, K = Kh + Kl)
// For fp64 storage Kh = K, Kl = 0 and the payoff is the correctly rounded K - s / s - K of the reference
// (om3:376-380); for fp32 storage it is that whenever K is a float, and within one ulp otherwise.
// The reference's `exercised` flag (om3:617/649) is the sign bit of the stored cash-flow; `flag` is the
// sign-bit mask under the sticky semantics and 0 otherwise.
template <typename R> struct Store;
template <> struct Store<float> {
  // exercised payoffs are > 0, so "sign bit set" is simply c < 0 (never true without the sticky mask)
  static __device__ __forceinline__ bool flagged(float c, unsigned int) { return c < 0.0f; }
  static __device__ __forceinline__ float with_flag(float p, unsigned int flag) {
    return __uint_as_float(__float_as_uint(p) | flag);
  }
};
template <> struct Store<double> {
  static __device__ __forceinline__ bool flagged(double c, unsigned int) { return c < 0.0; }
  static __device__ __forceinline__ double with_flag(double p, unsigned int flag) {
    return __hiloint2double((int)((unsigned int)__double2hiint(p) | flag), __double2loint(p));
  }
};

// Block-wide sum of QP (power of two) per-thread doubles.  Every thread calls it; contains one
// __syncthreads.  On return, in warp 0, lane l < 2*QP holds the CTA total of quantity l >> 1.
template <int QP, int NW>
__device__ __forceinline__ double block_totals(double (&acc)[QP], double* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  warp_reduce_scatter<QP>(acc, lane);
  if (reduce_scatter_owner<QP>(lane)) s_red[warp * QP + reduce_scatter_index<QP>(lane)] = acc[0];
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {  // lane -> quantity lane % QP; group lane / QP sums warps g, g+G, ...
    constexpr int G = 32 / QP;
    const int q = lane % QP, g = lane / QP;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w += G)
      if (w + g < NW) v += s_red[(w + g) * QP + q];
#pragma unroll
    for (int m = QP; m <= 16; m <<= 1) v += shfl_xor_f64(v, m);
    t = __shfl_sync(0xffffffffu, v, (lane >> 1) % QP);
  }
  return t;
}

constexpr unsigned long long kFxPoisonUnits = 1ull << 40;  // in units of 2^-4: 2^36 rows
constexpr double kFxPoisonCount = 68719476736.0;            // 2^36

// Warp 0 only.  `mine`: this CTA's total of quantity lane >> 1 (lanes < 2*QN), already in its final
// units.  Adds it into the parity's accumulators, waits until all `ncta` CTAs have arrived and returns
// the grid total of quantity `lane` in lanes < QN.  prev = this lane's accumulator value after the last
// completed exchange of the same parity (0 at launch).
template <int QN>
__device__ __forceinline__ double warp0_grid_sum(double mine, unsigned long long* xw, int par, int ncta,
                                                 unsigned long long& prev, int* flags, int* spins_out,
                                                 const ResComm& cm, unsigned int seq, int cta, bool& dead,
                                                 size_t slot_off = 0) {
  const int lane = threadIdx.x & 31;
  if (ncta == 1 && cm.nranks == 1) {  // the option fits one CTA: its totals are the grid totals
    if (spins_out) *spins_out = 0;
    return __shfl_sync(0xffffffffu, mine, (2 * lane) & 31);
  }
  unsigned long long sum = 0ull;
  int spins = 0;
  const bool single = cm.nranks == 1;
  // A total that does not fit the fixed-point format (or is NaN) POISONS the exchange: 2^36 is added to the row
  // count (quantity 0), which flows through the local and the cross-GPU sums like any other contribution, so every
  // CTA of every rank sees a count >= kFxPoisonCount at this date and fails alike (no rank returns a price that
  // silently misses another rank's moments).
  unsigned long long hi, lo;
  const bool enc_ok = fx_encode(mine, hi, lo) || lane >= 2 * QN;
  const bool poison = !__all_sync(0xffffffffu, enc_ok);
  if (poison && lane == 0) {
    atomicExch(flags, 1);
    hi += kFxPoisonUnits;
  }
  // Every lane runs the polling loops and leaves them on a warp-wide vote: with per-lane exits (lanes see their words
  // complete in different iterations) the warp stayed split into convergence groups after the loop, and every later
  // __shfl_sync / vote of the solve took the compiler's divergent-warp path (WARPSYNC.COLLECTIVE per instruction):
  // ~4 000 extra cycles per date whenever the poll needed more than one or two rounds ("box variance").
  const bool act = lane < 2 * QN;
  const int wl = act ? lane : 0;  // idle lanes re-read word 0 and ignore it
  unsigned long long* w = xw + ((size_t)par * kXchgWords + wl) * kXchgStride;
  if (act) red_relaxed_add_u64(w, (1ull << kFxCountShift) | ((lane & 1) ? lo : hi));
  if (single || cta == 0) {  // warp-uniform
    unsigned long long d;
    bool done;
    do {
      d = ld_relaxed_u64(w) - prev;
      ++spins;
      done = !act || (d >> kFxCountShift) == (unsigned long long)ncta;
    } while (!__all_sync(0xffffffffu, done));
    // a word that is complete cannot change before this CTA has contributed to the next exchange of the same parity
    if (act) {
      prev += d;
      sum = d & kFxValueMask;
    }
  }
  if (!single) {
    const unsigned int g = cm.g0 + seq;
    const unsigned long long tag = (unsigned long long)(g & 0xffu) << kFxCountShift;
    const size_t row = (size_t)(g & 1u) * kCommMaxRanks;
    if (cta == 0 && act) {
      // re-centre the biased high chunk so that the reader needs no CTA counts: |sum_hi - ncta 2^47| < 2^55
      const unsigned long long pay =
          (lane & 1) ? sum : (sum - ((unsigned long long)ncta << 47) + (1ull << 55)) & kFxValueMask;
      for (int p = 0; p < cm.nranks; ++p)
        st_relaxed_sys_u64(cm.slots[p] + slot_off + (row + cm.rank) * kXchgWords + lane, tag | pay);
    }
    // all ranks' words are polled together (independent loads in flight: one L2 round trip per sweep of the slots)
    const unsigned long long* mine_slots = cm.slots[cm.rank] + slot_off + row * kXchgWords + wl;
    unsigned long long v[kCommMaxRanks];
    unsigned int n = 0;
    for (;;) {
#pragma unroll
      for (int r = 0; r < kCommMaxRanks; ++r)
        v[r] = r < cm.nranks ? ld_relaxed_sys_u64(mine_slots + (size_t)r * kXchgWords) : tag;
      bool all = true;
#pragma unroll
      for (int r = 0; r < kCommMaxRanks; ++r) all &= (v[r] & ~kFxValueMask) == tag;
      if (__all_sync(0xffffffffu, all || !act) || dead) break;  // `dead` and n are warp-uniform
      if (++n >= kCommSpinLimit) dead = true;
    }
    if (act) {
      sum = 0ull;
#pragma unroll
      for (int r = 0; r < kCommMaxRanks; ++r) sum += v[r] & kFxValueMask;
      if (!(lane & 1)) sum -= (unsigned long long)cm.nranks << 55;  // two's complement: signed total of the high chunks
    }
    if (dead && lane == 0) atomicExch(flags + 1, 1);
  }
  __syncwarp();
  dead = __any_sync(0xffffffffu, dead);
  const unsigned long long other = __shfl_down_sync(0xffffffffu, sum, 1);
  double tot = single ? fx_decode(sum, other, ncta)  // meaningful on even lanes < 2*QN
                      : (double)(long long)sum * 0.0625 + (double)other * 2.220446049250313e-16;
  tot = __shfl_sync(0xffffffffu, tot, (2 * lane) & 31);     // quantity q: lane 2q -> lane q
  if (spins_out) *spins_out = spins;
  return tot;
}

constexpr int kTraceCols = 16;
// A stamp that cannot be issued before `dep` (a value loaded after a barrier) is available: BAR.SYNC defers its
// blocking, so a bare clock read placed after it measures the issue of the barrier, not its completion.
__device__ __forceinline__ long long clock_after(int dep) {
  long long c;
  asm volatile("{\n\t.reg .b32 t;\n\tmov.b32 t, %1;\n\tmov.u64 %0, %%clock64;\n\t}" : "=l"(c) : "r"(dep) : "memory");
  return c;
}
#define OPTMC_TRACE_AT(ph)            \
  do {                                \
    if (tr) tr[(ph)] = clock64();     \
  } while (0)

// Cash-flows are kept in "date-N money": the stored value is c~ = c_t / D_t with D_t = disc^(N - t), so
// the reference's per-date `cashflows *= discount` (om3:620) costs nothing for paths that are not in the
// regression: the value at date t is c~ * D_t (dg below), an exercise at date t stores payoff / D_t
// (payoff * dinv), and the price is mean(c~) * D_1.
template <typename R> struct PassConsts {
  R sgn, kk, c1, c2;
  R dinv;  // 1 / D_t of the decision date
  R dg;    // D_(t-1) of the Gram date
  unsigned int flag;
  int n_local;
};

// Exercise test  dec(S) > 0  (payoff - continuation as a polynomial in the raw price; strict, om3:644).
// fp64 storage evaluates it in fp64.  fp32 storage first evaluates it in fp32 together with a rigorous
// bound on the fp32 evaluation error (coefficient rounding + Horner: <= 4 u * sum |d_i| S^i, u = 2^-24;
// the filter uses 8 u); only paths whose fp32 value is inside the bound are re-decided in fp64, so the
// decisions are identical to the fp64 evaluation while the fp64 pipe sees a handful of paths per date.
template <bool WIDE> struct MaskOf { typedef unsigned int type; };
template <> struct MaskOf<true> { typedef unsigned long long type; };

template <typename R, int DEG> struct Decider;
template <int DEG> struct Decider<double, DEG> {
  double d[DEG + 1];
  __device__ __forceinline__ void load(const double* src, bool on) {
#pragma unroll
    for (int i = 0; i <= DEG; ++i) d[i] = on ? src[i] : 0.0;
  }
  __device__ __forceinline__ bool exact(double s) const { return poly_eval<DEG>(d, s) > 0.0; }
  __device__ __forceinline__ bool fast(double s, bool& sure) const { sure = true; return exact(s); }
};
template <int DEG> struct Decider<float, DEG> {
  double d[DEG + 1];
  float f[DEG + 1], b[DEG + 1];
  __device__ __forceinline__ void load(const double* src, bool on) {
#pragma unroll
    for (int i = 0; i <= DEG; ++i) {
      d[i] = on ? src[i] : 0.0;
      f[i] = (float)d[i];
      b[i] = fabsf(f[i]) * 4.76837158203125e-7f;  // 8 * 2^-24
    }
  }
  __device__ __forceinline__ bool exact(float s) const { return poly_eval<DEG>(d, (double)s) > 0.0; }
  __device__ __forceinline__ bool fast(float s, bool& sure) const {
    float p = f[DEG], e = b[DEG];
    const float as = fabsf(s);
#pragma unroll
    for (int i = DEG - 1; i >= 0; --i) { p = fmaf(p, s, f[i]); e = fmaf(e, as, b[i]); }
    sure = fabsf(p) > e;
    return p > 0.0f;
  }
};

// The sparse pass of the fp32 sweep reads the per-date constants it needs on a hit -- fp32 coefficients and error
// bounds of the decision polynomial, 1 / D_t, D_(t-1) -- from shared memory (written once per date by the solving
// lane).  Kept as fp64 uniform registers they were re-converted (F2F.F32.F64, a quarter-rate pipe) on EVERY hit: eight
// conversions per hit were a measurable part of the pass in grouped launches (27 k paths per CTA).
//   layout of the float block: f[0..DEG], b[0..DEG], dinv, dg
template <typename R, int DEG> struct HitConsts;
template <int DEG> struct HitConsts<double, DEG> {
  Decider<double, DEG> dec;
  double dinv_, dg_;
  __device__ __forceinline__ bool fast(double s, bool& sure) const { return dec.fast(s, sure); }
  __device__ __forceinline__ bool exact(double s) const { return dec.exact(s); }
  __device__ __forceinline__ double dinv() const { return dinv_; }
  __device__ __forceinline__ double dg() const { return dg_; }
};
template <int DEG> struct HitConsts<float, DEG> {
  const double* d;  // shared: decision polynomial in fp64 (rare exact path)
  const float* c;   // shared float block
  __device__ __forceinline__ bool fast(float s, bool& sure) const {
    float p = c[DEG], e = c[2 * DEG + 1];
    const float as = fabsf(s);
#pragma unroll
    for (int i = DEG - 1; i >= 0; --i) { p = fmaf(p, s, c[i]); e = fmaf(e, as, c[DEG + 1 + i]); }
    sure = fabsf(p) > e;
    return p > 0.0f;
  }
  __device__ __forceinline__ bool exact(float s) const {
    double dd[DEG + 1];
#pragma unroll
    for (int i = 0; i <= DEG; ++i) dd[i] = d[i];
    return poly_eval<DEG>(dd, (double)s) > 0.0;
  }
  __device__ __forceinline__ float dinv() const { return c[2 * DEG + 2]; }
  __device__ __forceinline__ float dg() const { return c[2 * DEG + 3]; }
};

// Passes over the thread's PPT paths (path j = tid + k * NT).  Slots past the end of the CTA's slice read an
// out-of-the-money sentinel from shared memory and hold a zero cash-flow, so no bounds predicate is needed.
//   decision of date t: exercise iff dec(S_t) > 0; the sticky flag is the sign bit of the stored cash-flow
//                       (om3:649);
//   Gram of date t-1:   ITM-masked (om3:621) raw-price moments, after the decision of date t.
//
// DENSE (textbook semantics, ~40% of the paths live at every date): two straight-line loops, dead lanes
// contribute zeros, paths the fp32 filter cannot decide are collected in a mask and re-decided afterwards.
template <typename R, int DEG, int PPT, int NT>
__device__ __forceinline__ void decide_pass(R (&cf)[PPT], const R* __restrict__ st_t, const Decider<R, DEG>& dec,
                                            const PassConsts<R>& pc, unsigned int& cnt, R& em) {
  const int tid = threadIdx.x;
  typedef typename MaskOf<(PPT > 32)>::type Mask;
  Mask amb = 0;  // paths the fp32 filter could not decide
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const R c = cf[k];
    const R sr = st_t[tid + k * NT];
    const R u = pc.sgn * sr;
    const bool live = !Store<R>::flagged(c, pc.flag) & (u > pc.kk);
    bool sure;
    const bool pos = dec.fast(sr, sure);
    const bool exer = live & sure & pos;
    amb |= (live & !sure) ? ((Mask)1 << k) : (Mask)0;
    const R pay = Store<R>::with_flag((fma(pc.sgn, sr, pc.c1) + pc.c2) * pc.dinv, pc.flag);
    cf[k] = exer ? pay : c;
    cnt += exer ? 1u : 0u;
    em = fmax(em, exer ? -u : (R)-INFINITY);
  }
  if (amb) {  // rare: a handful of paths per date sit within fp32 rounding of the exercise boundary
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      if (amb & ((Mask)1 << k)) {
        const R sr = st_t[tid + k * NT];
        if (dec.exact(sr)) {
          cf[k] = Store<R>::with_flag((fma(pc.sgn, sr, pc.c1) + pc.c2) * pc.dinv, pc.flag);
          cnt += 1u;
          em = fmax(em, -(pc.sgn * sr));
        }
      }
    }
  }
}

template <typename R, int DEG, int PPT, int NT>
__device__ __forceinline__ void gram_pass(const R (&cf)[PPT], const R* __restrict__ st_g, const PassConsts<R>& pc,
                                          double (&mom)[Moments<DEG>::Q], unsigned int& rows) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const R c = cf[k];
    const R s = st_g[tid + k * NT];
    const bool live = !Store<R>::flagged(c, pc.flag) & (pc.sgn * s > pc.kk);
    const R y = c * pc.dg;  // live lanes are unflagged: c >= 0
    rows += live ? 1u : 0u;
    moments_accumulate_nocount<DEG>(mom, (double)(live ? s : (R)0), (double)(live ? y : (R)0));
  }
}

// SPARSE (the reference's sticky mask: most in-the-money paths are already flagged, ~0.5% of the paths are in
// the regression per date): one loop; a warp skips step k when none of its 32 paths is live at date t or t-1.
template <typename R, int DEG, int PPT, int NT, bool DECIDE, bool GRAM>
__device__ __forceinline__ void sparse_pass(R (&cf)[PPT], const R* __restrict__ st_t, const R* __restrict__ st_g,
                                            const HitConsts<R, DEG>& dec, const PassConsts<R>& pc,
                                            double (&mom)[Moments<DEG>::Q], unsigned int& rows, unsigned int& cnt,
                                            R& em) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int j = tid + k * NT;
    R c = cf[k];
    const bool open = !Store<R>::flagged(c, pc.flag);
    const R st = DECIDE ? st_t[j] : (R)0;
    const R sg = GRAM ? st_g[j] : (R)0;
    const R ut = pc.sgn * st;
    const bool lt = DECIDE & open & (ut > pc.kk);
    const bool lg = GRAM & open & (pc.sgn * sg > pc.kk);
    if (__any_sync(0xffffffffu, lt | lg)) {
      if (DECIDE) {
        bool sure;
        bool pos = dec.fast(st, sure);
        if (lt & !sure) pos = dec.exact(st);  // rare: within fp32 rounding of the exercise boundary
        const bool exer = lt & pos;
        const R pay = Store<R>::with_flag((fma(pc.sgn, st, pc.c1) + pc.c2) * dec.dinv(), pc.flag);
        c = exer ? pay : c;
        cf[k] = c;
        cnt += exer ? 1u : 0u;
        em = fmax(em, exer ? -ut : (R)-INFINITY);
      }
      if (GRAM) {
        const bool live = lg & !Store<R>::flagged(c, pc.flag);  // not exercised just now (sticky mask)
        const R y = c * dec.dg();                                // live lanes are unflagged: c >= 0
        rows += live ? 1u : 0u;
        moments_accumulate_nocount<DEG>(mom, (double)(live ? sg : (R)0), (double)(live ? y : (R)0));
      }
    }
  }
}

template <typename R, int DEG, int PPT, int NT, bool SPARSE>
__global__ void __launch_bounds__(NT, 1) lsm_resident_kernel(const ResArgs ga) {
  constexpr int Q = Moments<DEG>::Q;
  constexpr int QP = Pow2<Q>::v;
  constexpr int NW = NT / 32;
  static_assert(Q <= kXchgMaxQ, "Gram vector must fit the exchange buffer");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3];
  __shared__ double s_red[NW * 16];
  __shared__ double s_dec[DEG + 1];   // exercise iff s_dec(s) > 0  (payoff - continuation as a polynomial in S)
  __shared__ float s_hit[2 * DEG + 4];  // fp32 sweep: f[], b[] of s_dec, 1 / D_t, D_(t-1) of the coming pass (HitConsts)
  __shared__ int s_valid;
  __shared__ unsigned long long s_bnd[2];
  __shared__ unsigned int s_cnt[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x / ga.cpg;
  const ResGroup a = ga.groups ? ga.groups[grp] : ga.one;
  const int cta = blockIdx.x - grp * ga.cpg, ncta = ga.cpg;
  const long long base = (long long)cta * a.chunk;
  const long long rem = a.M - base;
  const int n_local = (int)(rem < a.chunk ? rem : a.chunk);
  const unsigned int bytes = (unsigned int)((size_t)n_local * sizeof(R));  // multiple of 16 (planner)
  const bool is_put = a.is_put != 0;
  const bool sticky = ga.sticky != 0;
  const R sgn = (R)a.sgn, kk = (R)a.kk, c1 = (R)a.c1, c2 = (R)a.c2;
  const unsigned int flag = sticky ? 0x80000000u : 0u;
  const int N = a.N, nstage = ga.nstage;
  const R* Sbase = static_cast<const R*>(a.S) + base;

  // Stage ring: date t lives in slot (N - t) % nstage; the slots of the dates in use are tracked
  // incrementally (no integer division on the per-date path).
  auto slot_ptr = [&](int slot) -> R* { return reinterpret_cast<R*>(smem_raw + (size_t)slot * ga.stage_stride); };
  auto issue_load = [&](int t, int slot) {  // one thread
    uint64_t* bar = &mbar[slot];
    mbar_arrive_expect_tx(bar, bytes);
    bulk_load_1d(smem_raw + (size_t)slot * ga.stage_stride, Sbase + (size_t)t * a.ld, bytes, bar);
  };

  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) mbar_init(&mbar[s], 1);
    mbar_fence_init();
    s_bnd[0] = s_bnd[1] = bnd_none(a.is_put);
    s_cnt[0] = s_cnt[1] = 0u;
    s_valid = 0;
#pragma unroll
    for (int i = 0; i < 2 * DEG + 2; ++i) s_hit[i] = 0.f;
    s_hit[2 * DEG + 2] = 1.0f;              // 1 / D_N
    s_hit[2 * DEG + 3] = (float)a.disc;     // D_(N-1)
  }
  // out-of-the-money sentinel behind the slice in every stage (the bulk copies never touch it)
  {
    const R sentinel = is_put ? (R)INFINITY : (R)-INFINITY;
    for (int s = 0; s < nstage; ++s)
      for (int i = n_local + tid; i < PPT * NT; i += NT) slot_ptr(s)[i] = sentinel;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i)
      if (N - i >= 1) issue_load(N - i, i);
  }

  // warp-0 lane constants: lane l serves quantity l >> 1; raw-price moments are rescaled to x = S/K
  // (the regressor of SURVEY.md 8(c)) by invK^power before they enter the exchange.
  double qscale = 1.0;
  {
    const int pw = moment_power<DEG>(lane >> 1);
    for (int i = 0; i < pw; ++i) qscale *= a.invK;
  }
  unsigned long long prev0 = 0ull, prev1 = 0ull;  // warp 0: accumulator baselines of the two parities
  bool comm_dead = false;                         // warp 0: a peer rank stopped answering (sharded sweeps)
  // grouped + path-sharded: every group owns its own slot block on every rank
  const size_t slot_off = ga.groups ? (size_t)grp * kCommSlotWords : 0;
  long long* const tr_base = (ga.trace && grp == 0 && tid == 0 && (cta == 0 || cta == ncta - 1))
                                 ? ga.trace + (size_t)(cta == 0 ? 0 : 1) * (N + 1) * kTraceCols : nullptr;

  // ---- date N: cash-flows = payoff(S[N]) (om3:616) ----
  R cf[PPT];
  int slot_t = 0;           // ring slot of the iteration's decision date
  unsigned int phase = 0u;  // bit s: mbarrier phase parity the next wait on slot s expects
  auto wait_slot = [&](int slot) {
    mbar_wait_warp(&mbar[slot], (phase >> slot) & 1u);  // every thread of the CTA waits: whole warps
    phase ^= 1u << slot;
  };
  wait_slot(0);
  {
    const R* st = slot_ptr(0);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const R s = st[tid + k * NT];  // sentinel past the slice: out of the money -> 0
      const R p = fma(sgn, s, c1) + c2;
      cf[k] = (sgn * s > kk) ? p : (R)0;
    }
  }

  int seq = 0;
  if (ga.want_eu) {  // European leg on the option's own paths (om3:653-677): exp(-rT) payoff(S_N), summed like the price
    double eu[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const double c = (double)cf[k];
      eu[0] += c;
      eu[1] += c * c;
    }
    const double mine_eu = block_totals<2, NW>(eu, s_red);
    if (warp == 0) {
      unsigned long long pv = prev0;
      const double tot_l = warp0_grid_sum<2>(mine_eu, a.xw, 0, ncta, pv, a.flags, nullptr, ga.comm, 0u, cta, comm_dead, slot_off);
      prev0 = pv;
      const double s1 = __shfl_sync(0xffffffffu, tot_l, 0), s2 = __shfl_sync(0xffffffffu, tot_l, 1);
      if (cta == 0 && lane == 0) {
        const double n = (double)a.M_total;
        const double mean = s1 / n;
        double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
        if (var < 0.0) var = 0.0;
        double df = 1.0;
        for (int i = 0; i < N; ++i) df *= a.disc;
        a.final_out[4] = mean * df;
        a.final_out[5] = sqrt(var / n) * df;
      }
    }
    seq = 1;
  }

  // Iteration t (t = N .. 1) makes ONE branch-free pass over the CTA's paths:
  //   (1) exercise decision of date t with the polynomial solved at the end of iteration t+1 (none at t = N),
  //   (2) discount (om3:620) and the ITM-masked raw-price moments of date t-1 (om3:621 mask) -- skipped at t = 1,
  // followed by the block reduction, the grid sum and the solve for date t-1.
  double d_t = 1.0, dinv_t = 1.0;  // D_t and 1 / D_t of the iteration's decision date
  for (int t = N; t >= 1; --t) {
    long long* tr = tr_base ? tr_base + (size_t)t * kTraceCols : nullptr;
    const bool gram = t >= 2;
    const bool decide = t <= N - 1 && s_valid != 0;
    const int slot_g = slot_t + 1 == nstage ? 0 : slot_t + 1;  // ring slot of the Gram date t-1
    if (gram) wait_slot(slot_g);
    OPTMC_TRACE_AT(0);
    const R* st_t = slot_ptr(slot_t);
    const R* st_g = slot_ptr(gram ? slot_g : slot_t);
    double mom[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) mom[q] = 0.0;
    unsigned int rows = 0, cnt = 0;
    R em = (R)-INFINITY;  // max over exercised paths of -sgn * S: put -> max S, call -> -(min S)
    // D_t = disc^(N - t): every thread advances the same fp64 products, so they agree bit for bit
    const PassConsts<R> pc{sgn, kk, c1, c2, (R)dinv_t, (R)(d_t * a.disc), flag, n_local};
    Decider<R, DEG> dec;
    dec.load(s_dec, decide);
    if (SPARSE) {
      HitConsts<R, DEG> hc;
      if constexpr (sizeof(R) == 4) { hc.d = s_dec; hc.c = s_hit; }
      else { hc.dec = dec; hc.dinv_ = pc.dinv; hc.dg_ = pc.dg; }
      if (decide && gram) sparse_pass<R, DEG, PPT, NT, true, true>(cf, st_t, st_g, hc, pc, mom, rows, cnt, em);
      else if (gram) sparse_pass<R, DEG, PPT, NT, false, true>(cf, st_t, st_g, hc, pc, mom, rows, cnt, em);
      else if (decide) sparse_pass<R, DEG, PPT, NT, true, false>(cf, st_t, st_g, hc, pc, mom, rows, cnt, em);
    } else {
      if (decide) decide_pass<R, DEG, PPT, NT>(cf, st_t, dec, pc, cnt, em);
      if (gram) gram_pass<R, DEG, PPT, NT>(cf, st_g, pc, mom, rows);
    }
    d_t *= a.disc;
    dinv_t *= a.inv_disc;
    if (decide) {
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt) {  // warp-uniform
        const double ext = -(double)sgn * (double)em;
        unsigned long long b = (unsigned long long)__double_as_longlong(ext);
        if (!isfinite(ext)) b = bnd_none(a.is_put);
        b = is_put ? warp_max_u64(b) : warp_min_u64(b);
        if (lane == 0) {
          atomicAdd(&s_cnt[0], cnt);
          if (is_put) atomicMax(&s_bnd[0], b); else atomicMin(&s_bnd[0], b);
        }
      }
    }
    OPTMC_TRACE_AT(1);
    if (!gram) break;

    double acc[QP];
    mom[0] = (double)rows;
#pragma unroll
    for (int q = 0; q < QP; ++q) acc[q] = q < Q ? mom[q] : 0.0;
    // the __syncthreads inside block_totals also proves every thread is done with stage t
    const double mine = block_totals<QP, NW>(acc, s_red) * qscale;
    OPTMC_TRACE_AT(2);
    if (tid == 32) {  // bookkeeping off the critical path (warp 1)
      if (s_cnt[0]) {  // exercise statistics of date t
        if (a.exc) atomicAdd(a.exc + t, (unsigned long long)s_cnt[0]);
        if (a.bnd) {
          if (is_put) atomicMax(a.bnd + t, s_bnd[0]); else atomicMin(a.bnd + t, s_bnd[0]);
        }
        s_cnt[0] = 0u;
        s_bnd[0] = bnd_none(a.is_put);
      }
      if (t - nstage >= 1) issue_load(t - nstage, slot_t);  // refill the slot date t vacated
    }
    if (warp == 0) {
      int spins = 0;
      unsigned long long pv = (seq & 1) ? prev1 : prev0;
      const double tot_l = warp0_grid_sum<Q>(mine, a.xw, seq & 1, ncta, pv, a.flags, &spins, ga.comm, (unsigned int)seq, cta, comm_dead, slot_off);
      if (seq & 1) prev1 = pv; else prev0 = pv;
      OPTMC_TRACE_AT(4);
      if (tr) tr[7] = spins;
      double tot[Q], beta[DEG + 1];
#pragma unroll
      for (int q = 0; q < Q; ++q) tot[q] = __shfl_sync(0xffffffffu, tot_l, q);
      const bool poisoned = !(tot[0] < kFxPoisonCount);  // see warp0_grid_sum
      const bool ok = !poisoned && solve_poly<DEG>(tot, beta);  // every lane, identical inputs: no divergence
      if (lane == 0) {
        if (poisoned) atomicExch(a.flags, 1);
        s_valid = ok ? 1 : 0;
        if (ok) {
          // payoff - continuation = (+-K - b0) + (-+1 - b1/K) S - (b2/K^2) S^2 ...  (x = S/K)
          double sc = 1.0;
#pragma unroll
          for (int i = 0; i <= DEG; ++i) {
            double d = -beta[i] * sc;
            if (i == 0) d += is_put ? a.K : -a.K;
            if (i == 1) d += is_put ? -1.0 : 1.0;
            s_dec[i] = d;
            const float f = (float)d;
            s_hit[i] = f;
            s_hit[DEG + 1 + i] = fabsf(f) * 4.76837158203125e-7f;  // 8 * 2^-24, as Decider<float>::load
            sc *= a.invK;
          }
        }
        // d_t / dinv_t were advanced right after the pass: they are D_(t-1), 1 / D_(t-1) -- the next pass's constants
        s_hit[2 * DEG + 2] = (float)dinv_t;
        s_hit[2 * DEG + 3] = (float)(d_t * a.disc);
        if (cta == 0 && a.betas) {
#pragma unroll
          for (int i = 0; i <= DEG; ++i) a.betas[(size_t)(t - 1) * kMaxBeta + i] = ok ? beta[i] : nan("");
          a.nitm[t - 1] = poisoned ? 0ll : (long long)(tot[0] + 0.5);
        }
      }
      OPTMC_TRACE_AT(5);
    }
    ++seq;
    slot_t = slot_g;
    __syncthreads();  // decision polynomial of date t-1 visible
  }

  // ---- final reduction: mean and standard error of the cash-flows (om3:651) ----
  double fin[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const double c = fabs((double)cf[k]);  // slots past the slice hold 0
    fin[0] += c;
    fin[1] += c * c;
  }
  const double mine = block_totals<2, NW>(fin, s_red);
  if (tid == 32 && s_cnt[0]) {  // statistics of date 1 (its update pass is behind the barrier above)
    if (a.exc) atomicAdd(a.exc + 1, (unsigned long long)s_cnt[0]);
    if (a.bnd) {
      if (is_put) atomicMax(a.bnd + 1, s_bnd[0]); else atomicMin(a.bnd + 1, s_bnd[0]);
    }
  }
  if (warp == 0) {
    unsigned long long pv = (seq & 1) ? prev1 : prev0;
    const double tot_l = warp0_grid_sum<2>(mine, a.xw, seq & 1, ncta, pv, a.flags, nullptr, ga.comm, (unsigned int)seq, cta, comm_dead, slot_off);
    const double s1 = __shfl_sync(0xffffffffu, tot_l, 0), s2 = __shfl_sync(0xffffffffu, tot_l, 1);
    if (cta == 0 && lane == 0) {
      const double n = (double)a.M_total;
      const double mean = s1 / n;
      double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      const double scale = d_t * a.inv_disc * a.final_scale;  // D_1 (N - 1 discounts, om3:651), one more under TEXTBOOK
      a.final_out[0] = mean * scale;
      a.final_out[1] = sqrt(var / n) * scale;
      a.final_out[2] = s1;
      a.final_out[3] = s2;
    }
  }
}


// ================================================================================================================
// Speculative, warp-specialised form of the sparse (sticky-mask) sweep.
//
// Under the reference's sticky mask a path can change at date t only if it is still open AND in the money at t
// (~0.5% of the paths): for every other path the contribution to the moments of date t-1 does not depend on
// beta_t.  The CTA therefore runs NT compute threads plus ONE communication warp:
//   compute warps, iteration t:   S(t)  speculative pass -- moments of date t-1 of the paths that are not
//                                        candidates at t; candidates (open & in the money at t) go to a bit mask
//                                 wait  beta_t                                   (named barrier BETA)
//                                 C(t)  candidates only: decision of date t, their share of the moments of t-1
//                                 warp reduce-scatter -> shared memory, arrive   (named barrier TOT)
//   communication warp, iteration t:  arrive BETA (beta_t is in shared memory) ... wait TOT: CTA totals,
//                                 exchange statistics, refill the stage ring (TMA), grid-wide sum (and the NVLink
//                                 exchange of a path-sharded sweep), solve beta_(t-1) -> shared memory.
// S(t-1) of the compute warps runs while the communication warp is in the exchange + solve of beta_(t-1): the
// per-date critical path is  C + warp reduce + exchange + solve  instead of  pass + reduce + exchange + solve +
// barrier (profiles/trace_summary_r2.txt).  Decisions, moments and prices are those of lsm_resident_kernel
// (same arithmetic per path; the grid sum is order-independent), so the parity tests cover both.
// ================================================================================================================
// Barrier ids and thread counts are IMMEDIATES: ptxas sizes the CTA's barrier allocation from the ids it can see
// (with register operands it reports "used 1 barriers" and barriers 1, 2 do not block).
template <int ID, int NTHREADS> __device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
}
template <int ID, int NTHREADS> __device__ __forceinline__ void named_bar_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
}
constexpr int kBarBeta = 1, kBarTot = 2;

// One listed path of a warp at the current date: what the decision and its row of the regression need (price at the
// decision date, price at the Gram date, cash-flow), so that neither touches the price rows again.  The owning warp
// fills its own segment in S(t) (order: step k, then lane -- fixed), processes it with one lane per entry once beta_t
// is known, and the owners read the resulting cash-flow back.  Three planes (st, sg, c) of kCandCap entries per warp
// in shared memory; the rest spill to the warp's block in global memory (dates close to maturity only).
constexpr int kCandCap = 32;  // one full round of the warp; steady-state lists (2-18 entries) stay on chip
template <typename R> struct CandList {
  R* smem;         // this warp's segment: planes [3][kCandCap]
  R* spill;        // this warp's block in global memory: planes [3][spill_cap]
  int spill_cap;   // PPT * 32
  unsigned char* step_of;   // shared [kCandCap]: scan step k of entry e (entries in shared memory only)
  unsigned char* n_before;  // shared [PPT]: entries listed before step k (written at the steps that listed something)
  __device__ __forceinline__ void store(int e, R st, R sg, R c) const {
    if (e < kCandCap) { smem[e] = st; smem[kCandCap + e] = sg; smem[2 * kCandCap + e] = c; }
    else { R* p = spill + (e - kCandCap); p[0] = st; p[spill_cap] = sg; p[2 * spill_cap] = c; }
  }
  __device__ __forceinline__ void load(int e, R& st, R& sg, R& c) const {
    if (e < kCandCap) { st = smem[e]; sg = smem[kCandCap + e]; c = smem[2 * kCandCap + e]; }
    else { const R* p = spill + (e - kCandCap); st = p[0]; sg = p[spill_cap]; c = p[2 * spill_cap]; }
  }
  __device__ __forceinline__ void store_c(int e, R c) const {
    if (e < kCandCap) smem[2 * kCandCap + e] = c; else spill[2 * spill_cap + (e - kCandCap)] = c;
  }
  __device__ __forceinline__ R load_c(int e) const {
    return e < kCandCap ? smem[2 * kCandCap + e] : spill[2 * spill_cap + (e - kCandCap)];
  }
};

// S(t), steps [K0, K1): no beta needed, no arithmetic beyond the in-the-money tests.  Every path that is open and in
// the money at date t (a CANDIDATE: its cash-flow may change at t) or at date t-1 (a row of the regression of t-1)
// goes to the warp's list -- the price of a date it is out of the money at is replaced by an out-of-the-money
// sentinel -- and to the bit mask `listed`.  All floating-point work on these paths happens in cand_round, in one
// rolled loop with one lane per entry: the unrolled scan stays small (instruction cache) and holds no fp64 state.
// shared-memory accesses of the scan by 32-bit shared address + immediate offset: one instruction per access and two
// live address registers for the whole pass (generic pointers cost 64-bit address arithmetic per step, which at the
// 80-register budget of the wide shape the compiler re-materialises instead of keeping live)
template <int OFF> __device__ __forceinline__ float lds_imm(uint32_t addr, float) {
  float v; asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF)); return v;
}
template <int OFF> __device__ __forceinline__ double lds_imm(uint32_t addr, double) {
  double v; asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF)); return v;
}
__device__ __forceinline__ void sts_at(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_at(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

template <typename R, int PPT, int NT, int K, int K1> struct ScanSteps {
  typedef typename MaskOf<(PPT > 32)>::type Mask;
  // (st, sg) of step K arrive preloaded: the loads of step K + 1 are issued before the vote of step K, so their
  // latency overlaps it (the per-step chain load -> compare -> vote -> branch bounds the scan of a warp)
  static __device__ __forceinline__ void run(const R (&cf)[PPT], uint32_t at, uint32_t ag, R sgn, R kk_t, R kk_g, Mask& listed,
                                             const CandList<R>& list, uint32_t seg, int& n_w, R st, R sg) {
    const R c = cf[K];
    constexpr int KN = K + 1 < K1 ? K + 1 : K;
    const R st_n = lds_imm<KN * NT * (int)sizeof(R)>(at, R());
    const R sg_n = lds_imm<KN * NT * (int)sizeof(R)>(ag, R());
    const bool in = !(c < (R)0) & ((sgn * st > kk_t) | (sgn * sg > kk_g));  // open (sign bit clear) and in the money
    const unsigned int bal = __ballot_sync(0xffffffffu, in);
    if (bal) {  // warp-uniform; ~15% of the steps
      if (in) {
        const R otm = (R)(-sgn * INFINITY);
        const int e = n_w + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
        const R vt = (sgn * st > kk_t) ? st : otm, vg = (sgn * sg > kk_g) ? sg : otm;
        if (e < kCandCap) {
          const uint32_t q = seg + (uint32_t)e * (uint32_t)sizeof(R);
          sts_at(q, vt);
          sts_at(q + kCandCap * (uint32_t)sizeof(R), vg);
          sts_at(q + 2 * kCandCap * (uint32_t)sizeof(R), c);
          list.step_of[e] = (unsigned char)K;
        } else {
          R* p = list.spill + (e - kCandCap);
          p[0] = vt; p[list.spill_cap] = vg; p[2 * list.spill_cap] = c;
        }
        listed |= (Mask)1 << K;
      }
      list.n_before[K] = (unsigned char)(n_w < 255 ? n_w : 255);
      n_w += __popc(bal);
    }
    ScanSteps<R, PPT, NT, K + 1, K1>::run(cf, at, ag, sgn, kk_t, kk_g, listed, list, seg, n_w, st_n, sg_n);
  }
};
template <typename R, int PPT, int NT, int K1> struct ScanSteps<R, PPT, NT, K1, K1> {
  typedef typename MaskOf<(PPT > 32)>::type Mask;
  static __device__ __forceinline__ void run(const R (&)[PPT], uint32_t, uint32_t, R, R, R, Mask&, const CandList<R>&, uint32_t, int&, R, R) {}
};

template <typename R, int PPT, int NT, int K0, int K1>
__device__ __forceinline__ void scan_pass(const R (&cf)[PPT], const R* __restrict__ st_t, const R* __restrict__ st_g,
                                          const PassConsts<R>& pc, bool decide, bool gram,
                                          typename MaskOf<(PPT > 32)>::type& listed, const CandList<R>& list, int& n_w) {
  // decide / gram are off at the first / last date only: an "in the money" threshold nothing passes switches the test off
  const R kk_t = decide ? pc.kk : (R)INFINITY, kk_g = gram ? pc.kk : (R)INFINITY;
  const uint32_t at = smem_u32(st_t) + threadIdx.x * (uint32_t)sizeof(R), ag = smem_u32(st_g) + threadIdx.x * (uint32_t)sizeof(R);
  if (K0 >= K1) return;
  ScanSteps<R, PPT, NT, K0, K1>::run(cf, at, ag, pc.sgn, kk_t, kk_g, listed, list, smem_u32(list.smem), n_w,
                                     lds_imm<K0 * NT * (int)sizeof(R)>(at, R()), lds_imm<K0 * NT * (int)sizeof(R)>(ag, R()));
}

// The listed paths of the warp, one lane per entry, once beta_t is known (valid = the regression of date t
// succeeded): decision of date t for the candidates, the cash-flow after it (written back into the entry), and the
// row of the regression of date t-1 if the path is in the money there and was not exercised just now (sticky mask).
template <typename R, int DEG>
__device__ __forceinline__ void cand_round(const CandList<R>& list, int n_w, const HitConsts<R, DEG>& dec,
                                           const PassConsts<R>& pc, bool valid, double (&mom)[Moments<DEG>::Q],
                                           unsigned int& rows, unsigned int& cnt, R& em, unsigned long long& exer_steps) {
  const int lane = threadIdx.x & 31;
  for (int e0 = 0; e0 < n_w; e0 += 32) {  // warp-uniform trip count
    const int e = e0 + lane;
    const bool have = e < n_w;
    R st, sg, c;
    list.load(have ? e : 0, st, sg, c);
    const bool lt = valid & have & (pc.sgn * st > pc.kk);
    bool sure;
    bool pos = dec.fast(lt ? st : (R)0, sure);
    if (lt & !sure) pos = dec.exact(st);  // rare: within fp32 rounding of the exercise boundary
    const bool exer = lt & pos;
    const R pay = Store<R>::with_flag((fma(pc.sgn, st, pc.c1) + pc.c2) * pc.dinv, pc.flag);
    c = exer ? pay : c;
    if (exer) {
      list.store_c(e, c);
      exer_steps |= 1ull << (e < kCandCap ? list.step_of[e] : 63);  // bit 63: an exercised entry lives in the spill block
    }
    cnt += exer ? 1u : 0u;
    em = fmax(em, exer ? -(pc.sgn * st) : (R)-INFINITY);
    const bool live = have & !exer & (pc.sgn * sg > pc.kk);
    const R y = c * pc.dg;
    rows += live ? 1u : 0u;
    moments_accumulate_nocount<DEG>(mom, (double)(live ? sg : (R)0), (double)(live ? y : (R)0));
  }
}

// The owners read their candidates' cash-flows back (same ballots, hence the same entry indices, as in S(t)).  The
// steps with a candidate anywhere in the warp come from one warp-wide OR of the masks, so the walk over the PPT steps
// is a chain of warp-uniform tests (four steps per test), not of votes.
__device__ __forceinline__ unsigned int warp_or(unsigned int m) { return __reduce_or_sync(0xffffffffu, m); }
__device__ __forceinline__ unsigned long long warp_or(unsigned long long m) {
  return ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned int)(m >> 32)) << 32) |
         __reduce_or_sync(0xffffffffu, (unsigned int)m);
}
template <typename R, int PPT>
__device__ __forceinline__ void cand_apply(R (&cf)[PPT], const CandList<R>& list, int n_w,
                                           typename MaskOf<(PPT > 32)>::type cand, unsigned long long exer_steps) {
  typedef typename MaskOf<(PPT > 32)>::type Mask;
  const unsigned int lt_mask = (1u << (threadIdx.x & 31)) - 1u;
  // Only an EXERCISED entry changes its path's cash-flow, and few do (none at most dates): the steps that need the
  // read-back come from the warp-wide OR of the lanes' exercised-step masks, the entry index of (step, lane) from the
  // count the scan recorded at that step.  The walk over the PPT steps is a chain of warp-uniform tests.
  const unsigned long long ex = warp_or(exer_steps);
  if (ex == 0ull) return;
  if (!(ex >> 63)) {
    const R* cplane = list.smem + 2 * kCandCap;
#pragma unroll
    for (int k0 = 0; k0 < PPT; k0 += 4) {
      if ((ex >> k0) & 0xfull) {
#pragma unroll
        for (int k = k0; k < k0 + 4 && k < PPT; ++k) {
          if ((ex >> k) & 1ull) {
            const bool lt = (cand >> k) & (Mask)1;
            const unsigned int bal = __ballot_sync(0xffffffffu, lt);
            const int e = (int)list.n_before[k] + __popc(bal & lt_mask);
            if (lt && e < kCandCap) cf[k] = cplane[e];
          }
        }
      }
    }
    return;
  }
  const Mask any = warp_or(cand);  // dates close to maturity: entries spilled to global memory -- read everything back
  int n = 0;
#pragma unroll
  for (int k0 = 0; k0 < PPT; k0 += 4) {
    if ((any >> k0) & (Mask)0xf) {
#pragma unroll
      for (int k = k0; k < k0 + 4 && k < PPT; ++k) {
        if ((any >> k) & (Mask)1) {
          const bool lt = (cand >> k) & (Mask)1;
          const unsigned int bal = __ballot_sync(0xffffffffu, lt);
          if (lt) cf[k] = list.load_c(n + __popc(bal & lt_mask));
          n += __popc(bal);
        }
      }
    }
  }
}

template <typename R, int DEG, int PPT, int NT>
__global__ void __launch_bounds__(NT + 32, 1) lsm_resident_spec_kernel(const ResArgs ga) {
  constexpr int Q = Moments<DEG>::Q;
  constexpr int QP = Pow2<Q>::v;
  constexpr int NW = NT / 32;
  constexpr int NALL = NT + 32;
  constexpr int KH = (PPT + 1) / 2;  // steps [0, KH) = half A of a row, [KH, PPT) = half B
  typedef typename MaskOf<(PPT > 32)>::type Mask;
  static_assert(Q <= kXchgMaxQ, "Gram vector must fit the exchange buffer");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3][2];
  __shared__ double s_red[NW * 16];
  __shared__ double s_dec[DEG + 1];     // exercise iff s_dec(s) > 0  (payoff - continuation as a polynomial in S)
  __shared__ float s_hit[2 * DEG + 4];  // fp32 sweep: f[], b[] of s_dec (HitConsts); the discount slots are unused here
  __shared__ R s_seg[NW][3 * kCandCap];
  __shared__ unsigned char s_step[NW][kCandCap];
  __shared__ unsigned char s_nbefore[NW][(PPT + 3) / 4 * 4];
  __shared__ int s_valid;
  __shared__ unsigned int s_done[2];    // compute warps that finished half A / half B of the current date's rows
  __shared__ unsigned long long s_bnd[2];
  __shared__ unsigned int s_cnt[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_comm = warp == NW;
  const int grp = blockIdx.x / ga.cpg;
  const ResGroup a = ga.groups ? ga.groups[grp] : ga.one;
  const int cta = blockIdx.x - grp * ga.cpg, ncta = ga.cpg;
  const long long base = (long long)cta * a.chunk;
  const long long rem = a.M - base;
  const int n_local = (int)(rem < a.chunk ? rem : a.chunk);
  const int n_a = n_local < KH * NT ? n_local : KH * NT;  // paths of the slice in half A (a multiple of 16 bytes)
  const unsigned int bytes_a = (unsigned int)((size_t)n_a * sizeof(R));
  const unsigned int bytes_b = (unsigned int)((size_t)(n_local - n_a) * sizeof(R));
  const bool is_put = a.is_put != 0;
  const R sgn = (R)a.sgn, kk = (R)a.kk, c1 = (R)a.c1, c2 = (R)a.c2;
  const unsigned int flag = 0x80000000u;  // sticky semantics only
  const int N = a.N, nrow = ga.nstage;
  const R* Sbase = static_cast<const R*>(a.S) + base;
  const ResComm& cm = ga.comm;
  // grouped + path-sharded: every group owns its own slot block on every rank
  const size_t slot_off = ga.groups ? (size_t)grp * kCommSlotWords : 0;

  // Stage ring: date t lives in row (N - t) % nrow; every row is staged as two halves with their own mbarriers, so
  // that half A of row t can be refilled (with date t - nrow) while the pass is still in half B: with two rows on
  // chip (grouped launches: 27 k paths per CTA) the refill otherwise lands on the critical path of the next pass.
  auto row_ptr = [&](int row) -> R* { return reinterpret_cast<R*>(smem_raw + (size_t)row * ga.stage_stride); };
  auto issue_half = [&](int t, int row, int half) {  // one thread
    const unsigned int nb = half ? bytes_b : bytes_a;
    if (nb == 0) return;
    uint64_t* bar = &mbar[row][half];
    mbar_arrive_expect_tx(bar, nb);
    bulk_load_1d(smem_raw + (size_t)row * ga.stage_stride + (half ? (size_t)KH * NT * sizeof(R) : 0),
                 Sbase + (size_t)t * a.ld + (half ? KH * NT : 0), nb, bar);
  };

  if (tid == 0) {
    for (int s = 0; s < nrow; ++s) { mbar_init(&mbar[s][0], 1); mbar_init(&mbar[s][1], 1); }
    mbar_fence_init();
    s_bnd[0] = s_bnd[1] = bnd_none(a.is_put);
    s_cnt[0] = s_cnt[1] = 0u;
    s_done[0] = s_done[1] = 0u;
    s_valid = 0;
#pragma unroll
    for (int i = 0; i < 2 * DEG + 4; ++i) s_hit[i] = 0.f;
#pragma unroll
    for (int i = 0; i <= DEG; ++i) s_dec[i] = 0.0;
  }
  {  // out-of-the-money sentinel behind the slice in every row (the bulk copies never touch it)
    const R sentinel = is_put ? (R)INFINITY : (R)-INFINITY;
    for (int s = 0; s < nrow; ++s)
      for (int i = n_local + tid; i < PPT * NT; i += NALL) row_ptr(s)[i] = sentinel;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nrow; ++i)
      if (N - i >= 1) { issue_half(N - i, i, 0); issue_half(N - i, i, 1); }
  }
  long long* const tr_base = (ga.trace && grp == 0 && (cta == 0 || cta == ncta - 1))
                                 ? ga.trace + (size_t)(cta == 0 ? 0 : 1) * (N + 1) * kTraceCols : nullptr;

  if (is_comm) {
    // ---------------------------------------------------------------------------------------------------------
    // communication warp: CTA totals -> grid-wide sum -> solve; never touches a path
    // ---------------------------------------------------------------------------------------------------------
    double qscale = 1.0;  // lane l serves quantity l >> 1; raw-price moments are rescaled to x = S/K
    {
      const int pw = moment_power<DEG>(lane >> 1);
      for (int i = 0; i < pw; ++i) qscale *= a.invK;
    }
    unsigned long long prev0 = 0ull, prev1 = 0ull;
    bool comm_dead = false;
    int seq = 0;
    double d_t = 1.0;
    auto cta_totals = [&](int qp) -> double {  // lane -> quantity lane % qp; sums the warps' partials in a fixed order
      const int G = 32 / qp;
      const int q = lane % qp, g = lane / qp;
      double v = 0.0;
      for (int w = g; w < NW; w += G) v += s_red[w * qp + q];
      for (int m = qp; m <= 16; m <<= 1) v += shfl_xor_f64(v, m);
      return __shfl_sync(0xffffffffu, v, (lane >> 1) % qp);
    };
    // lane 0: exercise statistics of date t.  The compute warps add them (slot t & 1) AFTER they arrived at TOT(t),
    // off the critical path, so date t is flushed one iteration later: once every compute warp has passed BETA(t-1).
    auto flush_stats = [&](int t) {
      const int p = t & 1;
      if (s_cnt[p]) {
        if (a.exc) atomicAdd(a.exc + t, (unsigned long long)s_cnt[p]);
        if (a.bnd) {
          if (is_put) atomicMax(a.bnd + t, s_bnd[p]); else atomicMin(a.bnd + t, s_bnd[p]);
        }
        s_cnt[p] = 0u;
        s_bnd[p] = bnd_none(a.is_put);
      }
    };
    if (ga.want_eu) {  // European leg on the option's own paths (om3:653-677): exp(-rT) payoff(S_N), summed like the price
      named_bar_sync<kBarTot, NALL>();
      const double mine = cta_totals(2);
      unsigned long long pv = prev0;
      const double tot_l = warp0_grid_sum<2>(mine, a.xw, 0, ncta, pv, a.flags, nullptr, cm, 0u, cta, comm_dead, slot_off);
      prev0 = pv;
      const double s1 = __shfl_sync(0xffffffffu, tot_l, 0), s2 = __shfl_sync(0xffffffffu, tot_l, 1);
      if (cta == 0 && lane == 0) {
        const double n = (double)a.M_total;
        const double mean = s1 / n;
        double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
        if (var < 0.0) var = 0.0;
        double df = 1.0;
        for (int i = 0; i < N; ++i) df *= a.disc;
        a.final_out[4] = mean * df;
        a.final_out[5] = sqrt(var / n) * df;
      }
      seq = 1;
    }
    for (int t = N; t >= 1; --t) {
      long long* tr = (tr_base && lane == 0) ? tr_base + (size_t)t * kTraceCols : nullptr;
      if (tr) tr[10] = clock64();
      named_bar_arrive<kBarBeta, NALL>();  // beta_t (or "none" at t = N) is in shared memory
      d_t *= a.disc;
      if (t == 1) break;
      named_bar_sync<kBarTot, NALL>();     // every compute warp is done with date t; partials are in s_red
      if (tr) tr[0] = clock_after(*(volatile int*)&s_cnt[1]);
      const double mine = cta_totals(QP) * qscale;
      if (tr) tr[9] = clock_after(__double2loint(mine));
      if (lane == 0 && t + 1 < N) flush_stats(t + 1);
      int spins = 0;
      unsigned long long pv = (seq & 1) ? prev1 : prev0;
      const double tot_l = warp0_grid_sum<Q>(mine, a.xw, seq & 1, ncta, pv, a.flags, &spins, cm, (unsigned int)seq, cta, comm_dead, slot_off);
      if (seq & 1) prev1 = pv; else prev0 = pv;
      OPTMC_TRACE_AT(4);
      if (tr) tr[7] = spins;
      double tot[Q], beta[DEG + 1];
#pragma unroll
      for (int q = 0; q < Q; ++q) tot[q] = __shfl_sync(0xffffffffu, tot_l, q);
      const bool poisoned = !(tot[0] < kFxPoisonCount);  // some CTA / rank could not encode its totals (or NaN)
      const bool ok = !poisoned && solve_poly<DEG>(tot, beta);  // every lane, identical inputs: no divergence
      if (lane == 0) {
        if (poisoned) atomicExch(a.flags, 1);
        s_valid = ok ? 1 : 0;
        if (ok) {
          // payoff - continuation = (+-K - b0) + (-+1 - b1/K) S - (b2/K^2) S^2 ...  (x = S/K)
          double sc = 1.0;
#pragma unroll
          for (int i = 0; i <= DEG; ++i) {
            double d = -beta[i] * sc;
            if (i == 0) d += is_put ? a.K : -a.K;
            if (i == 1) d += is_put ? -1.0 : 1.0;
            s_dec[i] = d;
            const float f = (float)d;
            s_hit[i] = f;
            s_hit[DEG + 1 + i] = fabsf(f) * 4.76837158203125e-7f;  // 8 * 2^-24, as Decider<float>::load
            sc *= a.invK;
          }
        }
        if (cta == 0 && a.betas) {
#pragma unroll
          for (int i = 0; i <= DEG; ++i) a.betas[(size_t)(t - 1) * kMaxBeta + i] = ok ? beta[i] : nan("");
          a.nitm[t - 1] = poisoned ? 0ll : (long long)(tot[0] + 0.5);
        }
      }
      OPTMC_TRACE_AT(5);
      ++seq;
    }
    // ---- final reduction: mean and standard error of the cash-flows (om3:651) ----
    named_bar_sync<kBarTot, NALL>();
    const double mine = cta_totals(2);
    if (lane == 0) {
      if (N >= 3) flush_stats(2);
      if (N >= 2) flush_stats(1);
    }
    unsigned long long pv = (seq & 1) ? prev1 : prev0;
    const double tot_l = warp0_grid_sum<2>(mine, a.xw, seq & 1, ncta, pv, a.flags, nullptr, cm, (unsigned int)seq, cta, comm_dead, slot_off);
    const double s1 = __shfl_sync(0xffffffffu, tot_l, 0), s2 = __shfl_sync(0xffffffffu, tot_l, 1);
    if (cta == 0 && lane == 0) {
      const double n = (double)a.M_total;
      const double mean = s1 / n;
      double var = n > 1.0 ? (s2 - n * mean * mean) / (n - 1.0) : 0.0;
      if (var < 0.0) var = 0.0;
      const double scale = d_t * a.inv_disc * a.final_scale;  // D_1 (N - 1 discounts, om3:651), one more under TEXTBOOK
      a.final_out[0] = mean * scale;
      a.final_out[1] = sqrt(var / n) * scale;
      a.final_out[2] = s1;
      a.final_out[3] = s2;
    }
    return;
  }

  // -----------------------------------------------------------------------------------------------------------
  // compute warps
  // -----------------------------------------------------------------------------------------------------------
  R cf[PPT];
  unsigned int phase = 0u;  // bit (2 row + half): mbarrier phase parity the next wait on that half expects
  auto wait_half = [&](int row, int half) {
    if ((half ? bytes_b : bytes_a) == 0) return;
    const int b = 2 * row + half;
    mbar_wait_warp(&mbar[row][half], (phase >> b) & 1u);  // whole compute warps wait
    phase ^= 1u << b;
  };
  // the last compute warp to finish a half of row `row` (date t) refills it with date t - nrow
  auto release_half = [&](int t, int row, int half) {
    if (t - nrow < 1) return;
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&s_done[half], 1u) == (unsigned int)(NW - 1)) {
        s_done[half] = 0u;
        __threadfence_block();
        issue_half(t - nrow, row, half);
      }
    }
  };
  const CandList<R> list{&s_seg[warp][0],
                         static_cast<R*>(ga.spill) + ((size_t)blockIdx.x * NW + (warp < NW ? warp : 0)) * (size_t)(3 * PPT * 32), PPT * 32,
                         &s_step[warp < NW ? warp : 0][0], &s_nbefore[warp < NW ? warp : 0][0]};
  wait_half(0, 0);
  wait_half(0, 1);
  {  // date N: cash-flows = payoff(S[N]) (om3:616)
    const R* st = row_ptr(0);
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const R s = st[tid + k * NT];  // sentinel past the slice: out of the money -> 0
      const R p = fma(sgn, s, c1) + c2;
      cf[k] = (sgn * s > kk) ? p : (R)0;
    }
  }
  if (ga.want_eu) {
    double eu[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const double c = (double)cf[k];
      eu[0] += c;
      eu[1] += c * c;
    }
    warp_reduce_scatter<2>(eu, lane);
    if (reduce_scatter_owner<2>(lane)) s_red[warp * 2 + reduce_scatter_index<2>(lane)] = eu[0];
    named_bar_arrive<kBarTot, NALL>();
  }
  int row_t = 0;                   // ring row of the iteration's decision date
  double d_t = 1.0, dinv_t = 1.0;  // D_t and 1 / D_t of the iteration's decision date
  for (int t = N; t >= 1; --t) {
    // stamps from a warp on the communication warp's scheduler (NW % 4 == 3)
    long long* tr = (tr_base && tid == 96) ? tr_base + (size_t)t * kTraceCols : nullptr;
    const bool gram = t >= 2;
    const int row_g = row_t + 1 == nrow ? 0 : row_t + 1;  // ring row of the Gram date t-1
    const R* st_t = row_ptr(row_t);
    const R* st_g = row_ptr(gram ? row_g : row_t);
    double mom[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) mom[q] = 0.0;
    unsigned int rows = 0, cnt = 0;
    R em = (R)-INFINITY;  // max over exercised paths of -sgn * S: put -> max S, call -> -(min S)
    const PassConsts<R> pc{sgn, kk, c1, c2, (R)dinv_t, (R)(d_t * a.disc), flag, n_local};
    Mask listed = 0;
    unsigned long long exer_steps = 0ull;  // steps k whose entry of this lane was exercised at date t
    int n_w = 0;  // listed paths of this warp at date t (warp-uniform)
    if (gram) wait_half(row_g, 0);
    if (tr) tr[6] = clock64();  // S(t) starts
    scan_pass<R, PPT, NT, 0, KH>(cf, st_t, st_g, pc, t < N, gram, listed, list, n_w);
    release_half(t, row_t, 0);
    if (gram) wait_half(row_g, 1);
    scan_pass<R, PPT, NT, KH, PPT>(cf, st_t, st_g, pc, t < N, gram, listed, list, n_w);
    release_half(t, row_t, 1);
    __syncwarp();  // the warp's entries are visible to all of its lanes
    if (tr) tr[1] = clock_after(n_w);
    named_bar_sync<kBarBeta, NALL>();  // beta_t is in shared memory
    if (tr) tr[2] = clock_after(*(volatile int*)&s_valid);
    {
      HitConsts<R, DEG> hc;
      if constexpr (sizeof(R) == 4) { hc.d = s_dec; hc.c = s_hit; }
      else { hc.dec.load(s_dec, true); hc.dinv_ = pc.dinv; hc.dg_ = pc.dg; }
      cand_round<R, DEG>(list, n_w, hc, pc, t < N && s_valid != 0, mom, rows, cnt, em, exer_steps);
      if (tr) tr[8] = clock_after((int)rows + (int)cnt);
    }
    d_t *= a.disc;
    dinv_t *= a.inv_disc;
    if (gram) {
      double acc[QP];
      mom[0] = (double)rows;
#pragma unroll
      for (int q = 0; q < QP; ++q) acc[q] = q < Q ? mom[q] : 0.0;
      warp_reduce_scatter<QP>(acc, lane);
      if (reduce_scatter_owner<QP>(lane)) s_red[warp * QP + reduce_scatter_index<QP>(lane)] = acc[0];
      if (tr) tr[3] = clock_after(__double2loint(acc[0]));
      named_bar_arrive<kBarTot, NALL>();
    }
    // off the critical path (the communication warp is in the exchange): statistics, cash-flows of the candidates
    if (t < N) {
      if (a.exc != nullptr || a.bnd != nullptr) {  // per-date exercise statistics were asked for
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (cnt) {  // warp-uniform
          const double ext = -(double)sgn * (double)em;
          unsigned long long b = (unsigned long long)__double_as_longlong(ext);
          if (!isfinite(ext)) b = bnd_none(a.is_put);
          b = is_put ? warp_max_u64(b) : warp_min_u64(b);
          if (lane == 0) {
            atomicAdd(&s_cnt[t & 1], cnt);
            if (is_put) atomicMax(&s_bnd[t & 1], b); else atomicMin(&s_bnd[t & 1], b);
          }
        }
      }
      __syncwarp();
      if (tr) tr[12] = clock64();
      cand_apply<R, PPT>(cf, list, n_w, listed, exer_steps);
      __syncwarp();  // before the next pass overwrites the segment
      if (tr) tr[11] = clock_after(__float_as_int((float)cf[0]));
    }
    if (!gram) break;
    row_t = row_g;
  }
  // ---- final reduction ----
  double fin[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const double c = fabs((double)cf[k]);  // slots past the slice hold 0
    fin[0] += c;
    fin[1] += c * c;
  }
  warp_reduce_scatter<2>(fin, lane);
  if (reduce_scatter_owner<2>(lane)) s_red[warp * 2 + reduce_scatter_index<2>(lane)] = fin[0];
  named_bar_arrive<kBarTot, NALL>();
}

// ---- launch table (one translation unit per storage precision) -------------------------------------------
struct ResPlan {
  int ncta = 0;     // CTAs per option (group)
  int ngroups = 1;  // options swept by one launch
  int ppt = 0, nstage = 0, nt = 0;
  bool sparse = false;  // vote-skip passes (sticky semantics: few live paths per date)
  bool spec = false;    // sparse only: speculative, warp-specialised kernel (NT compute threads + 1 communication warp)
  long long chunk = 0;
  unsigned int stage_stride = 0;
  size_t smem = 0;
};

template <typename R, int DEG, int PPT, int NT, bool SPARSE>
int launch_resident_t(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  auto kern = lsm_resident_kernel<R, DEG, PPT, NT, SPARSE>;
  OPTMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  void* args[] = {(void*)&a};
  OPTMC_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(p.ncta * p.ngroups), dim3(NT), args, p.smem, ctx->stream));
  ctx->launches++; ctx->sw.n_launches++;
  return OPTMC_OK;
}

// (threads, paths-per-thread) shapes compiled for every precision / degree / sparsity, in order of capacity.
// 512-thread CTAs (<= 128 registers per thread); 1024-thread CTAs measured slower (barrier cost) and are not
// built; the 768 x 36 shape (80 registers) hides the scan latency of the largest fp32 slices 8% better than
// 512 x 54 and is used for the sparse fp32 sweep only (the other variants spill at 80 registers).
#define OPTMC_RES_SHAPES(X) \
  X(512, 1) X(512, 2) X(512, 4) X(512, 8) X(512, 12) X(512, 14) X(512, 16) X(512, 20) X(512, 24) X(512, 28) \
  X(512, 32) X(512, 40) X(512, 48) X(768, 36) X(512, 54)

template <typename R, int DEG, int PPT, int NT>
int launch_resident_spec_t(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  auto kern = lsm_resident_spec_kernel<R, DEG, PPT, NT>;
  OPTMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  {  // candidate-list overflow (dates close to maturity, where a third of the paths are candidates)
    const size_t need = (size_t)p.ncta * p.ngroups * (NT / 32) * (size_t)(3 * PPT * 32) * sizeof(R);
    int rc = ensure_bytes(&ctx->spill, &ctx->spill_bytes, need);
    if (rc) return rc;
    a.spill = ctx->spill;
  }
  void* args[] = {(void*)&a};
  OPTMC_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(p.ncta * p.ngroups), dim3(NT + 32), args, p.smem, ctx->stream));
  ctx->launches++; ctx->sw.n_launches++;
  return OPTMC_OK;
}

// Shapes of the speculative kernel: (compute threads, paths per thread); the CTA has 32 more threads (the
// communication warp), so that the warp count stays a multiple of four -- registers are partitioned per scheduler,
// and a 17th / 25th warp would cost every thread of the CTA a fifth of its registers.
#define OPTMC_RES_SPEC_SHAPES(X) \
  X(480, 1) X(480, 2) X(480, 4) X(480, 8) X(480, 12) X(480, 15) X(480, 16) X(480, 20) X(480, 24) X(480, 28) \
  X(480, 32) X(480, 40) X(480, 48) X(736, 37) X(480, 56)

template <typename R, int DEG> int launch_resident_shape(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  if (p.spec) {
#define X(NT_, PPT_) \
    if (p.nt == NT_ && p.ppt == PPT_) return launch_resident_spec_t<R, DEG, PPT_, NT_>(ctx, p, a);
    OPTMC_RES_SPEC_SHAPES(X)
#undef X
    set_error("no speculative resident instantiation for this slice size");
    return OPTMC_EUNSUPPORTED;
  }
#define X(NT_, PPT_)                                                                       \
  if (p.nt == NT_ && p.ppt == PPT_)                                                        \
    return p.sparse ? launch_resident_t<R, DEG, PPT_, NT_, true>(ctx, p, a)                \
                    : launch_resident_t<R, DEG, PPT_, NT_, false>(ctx, p, a);
  OPTMC_RES_SHAPES(X)
#undef X
  set_error("no resident instantiation for this slice size");
  return OPTMC_EUNSUPPORTED;
}

int launch_resident_f32_deg2(optmc_ctx* ctx, const ResPlan& p, ResArgs& a);
int launch_resident_f32_deg3(optmc_ctx* ctx, const ResPlan& p, ResArgs& a);
int launch_resident_f64_deg2(optmc_ctx* ctx, const ResPlan& p, ResArgs& a);
int launch_resident_f64_deg3(optmc_ctx* ctx, const ResPlan& p, ResArgs& a);

}  // namespace optmc
