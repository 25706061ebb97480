// xchg_bench.cu -- microbenchmark of the persistent sweep's grid-wide sum (development tool, not product).
// One warp per CTA, one CTA per SM, cooperative launch; every CTA runs ITER back-to-back exchanges of NW
// 64-bit words and reports cycles per exchange.  Variants:
//   mode 0: one accumulator word per quantity-chunk; every CTA reds into it and polls it (the kernel's scheme)
//   mode 1: R replicas: every CTA reds into all R copies, polls copy (cta % R)
//   mode 2: returning atom; the last arriver broadcasts the totals to per-CTA mailboxes; CTAs poll their own
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench tools/xchg_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void red_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add(unsigned long long* p, unsigned long long v) {
  unsigned long long o;
  asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory");
  return o;
}
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rlx(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct Args {
  unsigned long long* acc;   // [2][R][NW] * stride
  unsigned long long* mbox;  // [2][ncta][32]
  long long* out;            // [ncta]
  unsigned long long* chk;   // [ncta]
  int mode, nw, R, stride, iters, pollkind, sleep_ns, work;
};

__global__ void __launch_bounds__(32, 1) xchg_kernel(const Args a) {
  const int lane = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
  unsigned long long prev[2] = {0, 0};
  unsigned long long check = 0;
  // warm-up barrier so that every CTA starts roughly together
  long long t0 = 0;
  for (int it = -8; it < a.iters; ++it) {
    if (it == 0) t0 = clock64();
    const int par = it & 1;
    const unsigned long long add = (1ull << 56) | (unsigned long long)(cta + lane + 1);
    unsigned long long sum = 0;
    if (a.work) {  // emulate compute between exchanges (skew source)
      long long w0 = clock64();
      while (clock64() - w0 < a.work) {}
    }
    if (a.mode == 0 || a.mode == 1) {
      if (lane < a.nw) {
        const int myrep = cta % a.R;
        for (int r = 0; r < a.R; ++r)
          red_add(a.acc + (((size_t)par * a.R + r) * 32 + lane) * a.stride, add);
        unsigned long long* w = a.acc + (((size_t)par * a.R + myrep) * 32 + lane) * a.stride;
        unsigned long long d;
        unsigned long long& pv = par ? prev[1] : prev[0];
        for (;;) {
          d = (a.pollkind == 0 ? ld_rlx(w) : ld_vol(w)) - pv;
          if ((d >> 56) == (unsigned long long)ncta) break;
          if (a.sleep_ns) __nanosleep(a.sleep_ns);
        }
        pv += d;
        sum = d & ((1ull << 56) - 1);
      }
    } else if (a.mode == 2) {
      if (lane < a.nw) {
        unsigned long long* w = a.acc + ((size_t)par * 32 + lane) * a.stride;
        unsigned long long& pv = par ? prev[1] : prev[0];
        const unsigned long long old = atom_add(w, add);
        const unsigned long long d = old + add - pv;
        const unsigned long long tag = (unsigned long long)((it + 16) & 0xff) << 56;
        if ((d >> 56) == (unsigned long long)ncta) {  // last arriver for this word: broadcast
          const unsigned long long msg = (d & ((1ull << 56) - 1)) | tag;
          for (int c = 0; c < ncta; ++c) st_rlx(a.mbox + ((size_t)par * ncta + c) * 32 + lane, msg);
        }
        const unsigned long long* mb = a.mbox + ((size_t)par * ncta + cta) * 32 + lane;
        unsigned long long m;
        for (;;) {
          m = a.pollkind == 0 ? ld_rlx(mb) : ld_vol(mb);
          if ((m >> 56) == (tag >> 56)) break;
          if (a.sleep_ns) __nanosleep(a.sleep_ns);
        }
        sum = m & ((1ull << 56) - 1);
        pv += ((unsigned long long)ncta << 56) + sum;
      }
    }
    __syncwarp();
    check += sum;
  }
  const long long t1 = clock64();
  if (lane == 0) a.out[cta] = t1 - t0;
  if (lane == 0) a.chk[cta] = check;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int ncta = prop.multiProcessorCount;
  const int iters = 2000;
  unsigned long long *acc, *mbox, *chk;
  long long* out;
  const size_t acc_words = (size_t)2 * 16 * 32 * 512;
  CK(cudaMalloc(&acc, acc_words * 8));
  CK(cudaMalloc(&mbox, (size_t)2 * ncta * 32 * 8));
  CK(cudaMalloc(&out, ncta * 8));
  CK(cudaMalloc(&chk, ncta * 8));
  struct Cfg { int mode, nw, R, stride, pollkind, sleep_ns, work; const char* name; };
  std::vector<Cfg> cfgs = {
      {0, 16, 1, 1, 0, 0, 0, "mode0 nw16 stride8B (one line)"},
      {0, 16, 1, 2, 0, 0, 0, "mode0 nw16 stride16B (two lines)"},
      {0, 16, 1, 16, 0, 0, 0, "mode0 nw16 stride128B"},
      {0, 16, 1, 32, 0, 0, 0, "mode0 nw16 stride256B"},
      {0, 16, 1, 64, 0, 0, 0, "mode0 nw16 stride512B"},
      {0, 16, 1, 128, 0, 0, 0, "mode0 nw16 stride1KB"},
      {0, 16, 1, 256, 0, 0, 0, "mode0 nw16 stride2KB"},
      {0, 16, 1, 512, 0, 0, 0, "mode0 nw16 stride4KB"},
      {0, 1, 1, 1, 0, 0, 0, "mode0 nw1"},
      {0, 2, 1, 1, 0, 0, 0, "mode0 nw2 one line"},
      {0, 4, 1, 1, 0, 0, 0, "mode0 nw4 one line"},
      {0, 8, 1, 1, 0, 0, 0, "mode0 nw8 one line"},
      {0, 22, 1, 1, 0, 0, 0, "mode0 nw22 packed"},
      {0, 4, 1, 512, 0, 0, 0, "mode0 nw4 stride4KB"},
      {0, 8, 1, 512, 0, 0, 0, "mode0 nw8 stride4KB"},
      {1, 16, 2, 1, 0, 0, 0, "mode1 R2 packed"},
      {1, 16, 4, 1, 0, 0, 0, "mode1 R4 packed"},
      {0, 16, 1, 1, 0, 0, 3000, "mode0 packed + 3000-cycle work"},
  };
  for (const Cfg& c : cfgs) {
    CK(cudaMemset(acc, 0, acc_words * 8));
    CK(cudaMemset(mbox, 0, (size_t)2 * ncta * 32 * 8));
    Args a{acc, mbox, out, chk, c.mode, c.nw, c.R, c.stride, iters, c.pollkind, c.sleep_ns, c.work};
    void* args[] = {(void*)&a};
    CK(cudaLaunchCooperativeKernel((const void*)xchg_kernel, dim3(ncta), dim3(32), args, 0, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(ncta);
    std::vector<unsigned long long> hc(ncta);
    CK(cudaMemcpy(h.data(), out, ncta * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.data(), chk, ncta * 8, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    bool same = true;
    for (int i = 1; i < ncta; ++i) same = same && hc[i] == hc[0];
    printf("%-44s cycles/exchange: min %7.0f med %7.0f max %7.0f   (net of work: %7.0f) check %s\n", c.name,
           (double)h[0] / iters, (double)h[ncta / 2] / iters, (double)h[ncta - 1] / iters,
           (double)h[ncta / 2] / iters - c.work, same ? "ok" : "MISMATCH");
  }
  return 0;
}
