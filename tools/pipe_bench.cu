// pipe_bench.cu -- instruction-throughput microbenchmark for the fp64-pipe operations the LSM sweep uses
// (development tool).  One 512-thread CTA per SM (4 warps per scheduler, like the persistent sweep); each
// thread runs ITER x 8 independent operations of one kind; reports cycles per warp-instruction per scheduler.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench tools/pipe_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITER = 2000;

template <int OP> __global__ void __launch_bounds__(1024, 1) k(double* out, long long* cyc, float seed, int nthreads) {
  double a[8];
  float f[8];
  int pc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; f[i] = seed * (i + 1) + threadIdx.x; }
  const double m = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma(a[i], m, c);                       // DFMA
      if (OP == 1) a[i] = a[i] + c;                               // DADD
      if (OP == 2) a[i] = a[i] * m;                               // DMUL
      if (OP == 3) { f[i] = f[i] * 1.0000001f; a[i] = (double)f[i]; }            // FMUL + F2F.F64.F32
      if (OP == 4) { pc += (a[i] > (double)it) ? 1 : 0; a[i] = a[i] + c; }       // DSETP + DADD
      if (OP == 5) f[i] = fmaf(f[i], 1.0000001f, 1e-9f);          // FFMA (reference)
      if (OP == 6) { f[i] = f[i] * 1.0000001f; a[i] += (double)f[i]; }           // FMUL + F2F + DADD
      if (OP == 7) { unsigned u = __float_as_uint(f[i]); f[i] = f[i] * 1.0000001f;   // integer widening f32 -> f64
                     int hi = (int)(((u >> 3) & 0x0fffffffu) + 0x38000000u) | (int)(u & 0x80000000u);
                     a[i] = __hiloint2double(hi, (int)(u << 29)); }
      if (OP == 8) { f[i] = (float)a[i]; a[i] = a[i] + (double)i; }              // F2F.F32.F64 + DADD
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + pc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char* name, int nthreads, double* out, long long* cyc, int nsm) {
  k<OP><<<nsm, nthreads>>>(out, cyc, 1.5f, nthreads);
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double warps_per_sched = nthreads / 32 / 4.0;
  printf("%-34s threads %4d: %6.2f cycles per warp-instruction-group per scheduler (8 ops: %.2f each)\n", name, nthreads,
         (double)h / ITER / 8 / warps_per_sched, (double)h / ITER / 8 / warps_per_sched);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  double* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 1024 * 8));
  CK(cudaMalloc(&cyc, prop.multiProcessorCount * 8));
  for (int nt : {512, 1024}) {
    run<0>("DFMA", nt, out, cyc, prop.multiProcessorCount);
    run<1>("DADD", nt, out, cyc, prop.multiProcessorCount);
    run<2>("DMUL", nt, out, cyc, prop.multiProcessorCount);
    run<3>("FMUL + F2F.F64.F32", nt, out, cyc, prop.multiProcessorCount);
    run<4>("DSETP + DADD", nt, out, cyc, prop.multiProcessorCount);
    run<5>("FFMA", nt, out, cyc, prop.multiProcessorCount);
    run<6>("FMUL + F2F.F64.F32 + DADD", nt, out, cyc, prop.multiProcessorCount);
    run<7>("FMUL + integer widening (5 ALU)", nt, out, cyc, prop.multiProcessorCount);
    run<8>("F2F.F32.F64 + DADD", nt, out, cyc, prop.multiProcessorCount);
  }
  return 0;
}
