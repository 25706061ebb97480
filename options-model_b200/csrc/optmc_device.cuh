// optmc_device.cuh -- device-only helpers: deterministic reductions, vector I/O, mbarrier + bulk-copy
// (TMA) PTX, and acquire/release flag accessors for the cross-CTA exchange.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace optmc {

__device__ __forceinline__ double shfl_xor_f64(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Fixed-order butterfly: every lane ends with the same total; order is independent of data => deterministic.
template <int Q> __device__ __forceinline__ void warp_allreduce_sum(double (&a)[Q]) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
    for (int q = 0; q < Q; ++q) a[q] += shfl_xor_f64(a[q], m);
  }
}

// Block reduction of Q doubles.  smem: [NWARPS][Q].  Result valid in every lane of warp 0.
template <int Q, int NWARPS> __device__ __forceinline__ void block_reduce_sum(double (&a)[Q], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_allreduce_sum<Q>(a);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) smem[warp * Q + q] = a[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) a[q] = (lane < NWARPS) ? smem[lane * Q + q] : 0.0;
    warp_allreduce_sum<Q>(a);
  }
}

// Recursive-halving warp reduce-scatter of QP (power of two) doubles: QP + 3 adds instead of 5*QP.
// On exit a[0] holds the warp total of quantity `reduce_scatter_index<QP>(lane)`; lanes that differ only
// in the low log2(32/QP) lane bits hold identical copies.  Fixed pattern => deterministic.
template <int QP> __device__ __forceinline__ void warp_reduce_scatter(double (&a)[QP], int lane) {
  int width = QP;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    if (width > 1) {
      const int half = width / 2;
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < QP / 2; ++i) {
        if (i < half) {
          const double send = up ? a[i] : a[i + half];
          const double keep = up ? a[i + half] : a[i];
          a[i] = keep + shfl_xor_f64(send, m);
        }
      }
      width = half;
    } else {
      a[0] += shfl_xor_f64(a[0], m);
    }
  }
}
template <int QP> __device__ __forceinline__ int reduce_scatter_index(int lane) {
  int q = 0, width = QP;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    if (width > 1) { q = q * 2 + ((lane & m) ? 1 : 0); width /= 2; }
  }
  return q;
}
template <int QP> __device__ __forceinline__ bool reduce_scatter_owner(int lane) {
  return (lane & ((32 / QP) - 1)) == 0;  // one writer per quantity (QP <= 32)
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o > v ? o : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
    v = o < v ? o : v;
  }
  return v;
}
// Exercise-boundary bookkeeping: put -> max exercised S (bits of a positive double order like integers),
// call -> min exercised S.  "none" is 0 for max and ~0 for min.
__device__ __forceinline__ unsigned long long bnd_none(int is_put) { return is_put ? 0ull : ~0ull; }

// ---- vector store / load of VEC consecutive elements -------------------------------------------------
template <typename R, int VEC> struct VecIO;
template <> struct VecIO<float, 4> {
  static __device__ __forceinline__ void store(float* p, const float (&x)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  }
};
template <> struct VecIO<double, 4> {
  static __device__ __forceinline__ void store(double* p, const double (&x)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(x[0], x[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(x[2], x[3]);
  }
};
template <typename R> struct VecIO<R, 1> {
  static __device__ __forceinline__ void store(R* p, const R (&x)[1]) { *p = x[0]; }
};
template <> struct VecIO<float, 2> {
  static __device__ __forceinline__ void store(float* p, const float (&x)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]);
  }
};
template <> struct VecIO<double, 2> {
  static __device__ __forceinline__ void store(double* p, const double (&x)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
  }
};

// ---- shared-memory barrier + 1-D bulk async copy (TMA engine, no tensor map needed for 1-D) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// The same wait for a WHOLE warp (all 32 lanes call it): the loop is left on a warp-wide vote, so the warp is converged
// afterwards.  With per-lane exits (a lane whose try_wait timed out goes round again) the warp can stay split, and the
// compiler's divergent-warp path then runs every later vote / shuffle through WARPSYNC.COLLECTIVE.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!__all_sync(0xffffffffu, ok != 0u));
}
// global -> shared bulk copy, completion signalled on the mbarrier (bytes % 16 == 0, 16-byte aligned).
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- cross-CTA flags ----------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// LL-style exchange words: {payload32, epoch32} in one 8-byte scalar (single-copy atomic), two words per double.
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v2u64(const unsigned long long* p, unsigned long long& a,
                                                 unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void red_relaxed_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// system scope: the word lives in (or is read back from) another GPU's memory over NVLink
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace optmc
