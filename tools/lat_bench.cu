// lat_bench.cu -- dependent-chain latency (SM cycles per operation, one warp) of the fp64 / conversion instructions on
// the per-date critical path of the persistent sweep (solve, fixed-point encode / decode).  Development tool: the pool's
// B200 boxes differ by 6x on some of them.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/lat_bench tools/lat_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int OP> __global__ void k(double x, long long* out) {
  double a = x;
  float f = (float)x;
  long long acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    if (OP == 0) a = fma(a, 1.0000001, 1e-9);
    if (OP == 1) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 1.0; }
    if (OP == 2) { long long v = (long long)a; a = (double)v + 1.5; }
    if (OP == 3) a = floor(a) + 1.25;
    if (OP == 4) { f = (float)a; a = (double)f + 1e-9; }
    if (OP == 5) { unsigned long long v = (unsigned long long)a; a = (double)v + 1.5; }
    if (OP == 6) a = 1.0 / a + 1.0;
    if (OP == 7) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(f)); f += 1.0f; }
    if (OP == 8) a = sqrt(a) + 1.0;
  }
  const long long t1 = clock64();
  out[0] = t1 - t0;
  out[1] = (long long)a + (long long)f + acc;
}
template <int OP> void run(const char* name, long long* d) {
  long long h[2];
  k<OP><<<1, 32>>>(3.7, d); k<OP><<<1, 32>>>(3.7, d);
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s %7.1f cycles per iteration\n", name, h[0] / 256.0);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0>("DFMA", d);
  run<1>("rcp.approx.ftz.f64 (MUFU.RCP64H) + DADD", d);
  run<2>("F2I.S64.F64 + I2F.F64.S64 + DADD", d);
  run<3>("floor (FRND.F64) + DADD", d);
  run<4>("F2F.F32.F64 + F2F.F64.F32 + DADD", d);
  run<5>("F2I.U64.F64 + I2F.F64.U64 + DADD", d);
  run<6>("IEEE 1.0 / x + DADD", d);
  run<7>("rcp.approx.ftz.f32 + FADD", d);
  run<8>("sqrt (f64) + DADD", d);
  return 0;
}
