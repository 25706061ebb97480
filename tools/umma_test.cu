// umma_test.cu -- validates the hand-built tcgen05 descriptors used by the tensor-core MLP kernels (development
// tool).  One CTA of 128 threads: two bf16 128x128 tiles are stored in shared memory in the NO-SWIZZLE core-matrix
// layout (8x8 cores of 128 contiguous bytes; k-cores 128 B apart, 8-row groups 2048 B apart), then
//   D1[m][n] = sum_k A[m][k] B[n][k]     both operands K-major
//   D2[m][n] = sum_k A[k][m] B[k][n]     both operands MN-major views of the SAME bytes
// are computed with tcgen05.mma (M = N = 128, K = 16 x 8), read back with tcgen05.ld and compared with a host
// reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test tools/umma_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (row, col) of a [128][128] bf16 tile in the core-matrix layout
__host__ __device__ inline int core_off(int row, int col) { return (row >> 3) * 2048 + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                     // D format: f32
  d |= 1u << 7;                     // A format: bf16
  d |= 1u << 10;                    // B format: bf16
  d |= (uint32_t)(a_mn & 1) << 15;  // A major: 0 = K, 1 = MN
  d |= (uint32_t)(b_mn & 1) << 16;  // B major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_test_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D1, float* D2) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 32768;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill the tiles: thread r writes row r (16 chunks of 16 bytes)
  for (int c = 0; c < 16; ++c) {
    *reinterpret_cast<uint4*>(sA + core_off(tid, c * 8)) = *reinterpret_cast<const uint4*>(A + tid * 128 + c * 8);
    *reinterpret_cast<uint4*>(sB + core_off(tid, c * 8)) = *reinterpret_cast<const uint4*>(B + tid * 128 + c * 8);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy tile writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    // D1: K-major.  k-step = 16 elements = 2 cores along K = 256 bytes; LBO = 128 (k-cores), SBO = 2048 (8-row groups)
    const uint32_t id1 = make_idesc(128, 128, 0, 0);
    for (int k = 0; k < 8; ++k) umma(tmem + 0, make_desc(a0 + k * 256, 128, 2048), make_desc(b0 + k * 256, 128, 2048), id1, k > 0);
    // D2: MN-major views of the same bytes.  MN groups of 8 are 128 B apart (SBO), K groups of 8 rows 2048 B apart (LBO);
    // k-step = 16 rows = 4096 bytes
    const uint32_t id2 = make_idesc(128, 128, 1, 1);
    for (int k = 0; k < 8; ++k) umma(tmem + 128, make_desc(a0 + k * 4096, 2048, 128), make_desc(b0 + k * 4096, 2048, 128), id2, k > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // read back: thread r owns TMEM lane r
  for (int which = 0; which < 2; ++which) {
    float* D = which ? D2 : D1;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + which * 128 + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) D[tid * 128 + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}

int main() {
  std::vector<__nv_bfloat16> hA(128 * 128), hB(128 * 128);
  std::vector<float> fA(128 * 128), fB(128 * 128);
  srand(1);
  for (int i = 0; i < 128 * 128; ++i) {
    hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    fA[i] = __bfloat162float(hA[i]);
    fB[i] = __bfloat162float(hB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float *dD1, *dD2;
  CK(cudaMalloc(&dA, 128 * 128 * 2)); CK(cudaMalloc(&dB, 128 * 128 * 2));
  CK(cudaMalloc(&dD1, 128 * 128 * 4)); CK(cudaMalloc(&dD2, 128 * 128 * 4));
  CK(cudaMemcpy(dA, hA.data(), 128 * 128 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), 128 * 128 * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
  umma_test_kernel<<<1, 128, 65536 + 1024>>>(dA, dB, dD1, dD2);
  CK(cudaDeviceSynchronize());
  std::vector<float> D1(128 * 128), D2(128 * 128);
  CK(cudaMemcpy(D1.data(), dD1, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D2.data(), dD2, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0, mx = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double r1 = 0, r2 = 0;
      for (int k = 0; k < 128; ++k) { r1 += (double)fA[m * 128 + k] * fB[n * 128 + k]; r2 += (double)fA[k * 128 + m] * fB[k * 128 + n]; }
      e1 = fmax(e1, fabs(r1 - D1[m * 128 + n]));
      e2 = fmax(e2, fabs(r2 - D2[m * 128 + n]));
      mx = fmax(mx, fabs(r1));
    }
  printf("K-major  max abs err %.3e ; MN-major max abs err %.3e ; max |ref| %.3f -> %s\n", e1, e2, mx,
         (e1 < 1e-3 && e2 < 1e-3) ? "PASS" : "FAIL");
  printf("D1[0][0..3] = %f %f %f %f ; D2[0][0..3] = %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3], D2[0], D2[1], D2[2], D2[3]);
  return 0;
}
