// optmc_tc.cuh -- tcgen05 (5th-generation tensor core) building blocks shared by the network kernels
// (lsm_mlp.cu, lsm_gnet.cu): NO-SWIZZLE core-matrix shared-memory layout, UMMA shared-memory / instruction
// descriptors, single-CTA bf16 x bf16 -> fp32 MMA issue + commit, TMEM loads, bf16 packing, and the two
// synchronisation idioms (publish tile writes to the async proxy; wait for a commit).  Descriptor encodings are
// validated against a host reference by tools/umma_test.cu.
//
// Tile layout: a [128][128] bf16 tile is stored as 8x8 "core matrices" of 128 contiguous bytes (8 rows x 16 B);
// cores along the column index are 128 B apart, 8-row groups 2048 B apart.  The same bytes serve as a K-major
// operand [row][col] (LBO 128, SBO 2048, 256 B per K = 16 step) and, with LBO / SBO exchanged, as an MN-major
// operand [col][row] (LBO 2048, SBO 128, 4096 B per step) -- no transposed copies.  A [128][16] "panel" uses the
// same core layout with 256 B per 8-row group.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "optmc_device.cuh"

namespace optmc {

constexpr int kTcTileBytes = 128 * 128 * 2;  // 32 KB
constexpr int kTcPanelBytes = 128 * 16 * 2;  // 4 KB

__device__ __forceinline__ int core_off(int row, int col) {  // bytes; [128][128] bf16 tile
  return (row >> 3) * 2048 + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2;
}
__device__ __forceinline__ int aux_off(int row, int col) {   // bytes; [128][16] bf16 panel
  return (row >> 3) * 256 + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2;
}
__device__ __forceinline__ unsigned long long umma_desc(unsigned int saddr, unsigned int lbo, unsigned int sbo) {
  return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)((lbo >> 4) & 0x3fff) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);  // version 1, no swizzle
}
__device__ __forceinline__ unsigned int umma_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(a_mn & 1) << 15) | ((unsigned)(b_mn & 1) << 16) |
         ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);  // f32 accumulate, bf16 x bf16
}
__device__ __forceinline__ void umma_f16(unsigned int d_tmem, unsigned long long a, unsigned long long b, unsigned int idesc,
                                         unsigned int acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned int taddr, float (&v)[32]) {
  unsigned int r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(unsigned int taddr, float (&v)[16]) {
  unsigned int r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ unsigned int pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // one packed conversion (low half = first argument)
  return *reinterpret_cast<const unsigned int*>(&h);
}
__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
  return make_uint4(pack2_bf16(v[0], v[1]), pack2_bf16(v[2], v[3]), pack2_bf16(v[4], v[5]), pack2_bf16(v[6], v[7]));
}
// every thread of the CTA calls this (whole warps): the wait is left on a warp-wide vote, so the warp is converged for
// the .sync.aligned tensor-memory loads that follow
__device__ __forceinline__ void tc_bar_wait(unsigned long long* bar, unsigned int parity) {
  mbar_wait_warp(reinterpret_cast<uint64_t*>(bar), parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// order this thread's shared-memory tile writes (generic proxy) and TMEM reads before the barrier that precedes the
// next MMA issue (async proxy)
__device__ __forceinline__ void tc_publish() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
}


}  // namespace optmc
