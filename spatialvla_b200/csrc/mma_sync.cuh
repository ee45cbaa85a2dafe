// Warp-level tensor-core helpers (mma.sync.m16n8k16 bf16, ldmatrix, cp.async) shared by the backward kernels of the fine-tune step.
// The forward hot path runs on tcgen05 (gemm_tcgen05.cu, attention_tc.cu); the backward attention and the rank-r LoRA gradient
// reductions are small next to the dX GEMMs (which reuse the tcgen05 GEMM) and are written on the warp-level MMA first.
#pragma once
#include "svla_common.cuh"

namespace svla_mma {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Fragment addressing on a row-major bf16 tile T[rows][ld] in shared memory (lane = threadIdx.x & 31):
//   A operand (16 x 16 block at row r0, col c0, rows = M, cols = K):            ldsm_x4  (a_addr)
//   B operand from a tile stored [n][k] (K contiguous; e.g. K / V rows for Q K^T): ldsm_x4  (bnk_addr) -> {b0,b1} n-tile 0, {b2,b3} n-tile 1
//   B operand from a tile stored [k][n] (N contiguous; e.g. V rows for P V):       ldsm_x4_t(bkn_addr) -> {b0,b1} n-tile 0, {b2,b3} n-tile 1
//   A operand from a tile stored [k][m] (M contiguous, i.e. A^T in memory):         ldsm_x4_t(akm_addr)
__device__ __forceinline__ const __nv_bfloat16* a_addr(const __nv_bfloat16* t, int ld, int r0, int c0, int lane) {
  return t + (r0 + (lane & 15)) * ld + c0 + (lane >> 4) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* bnk_addr(const __nv_bfloat16* t, int ld, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return t + (n0 + (mi >> 1) * 8 + (lane & 7)) * ld + k0 + (mi & 1) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* bkn_addr(const __nv_bfloat16* t, int ld, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  return t + (k0 + (mi & 1) * 8 + (lane & 7)) * ld + n0 + (mi >> 1) * 8;
}
__device__ __forceinline__ const __nv_bfloat16* akm_addr(const __nv_bfloat16* t, int ld, int k0, int m0, int lane) {
  const int mi = lane >> 3;
  return t + (k0 + (mi >> 1) * 8 + (lane & 7)) * ld + m0 + (mi & 1) * 8;
}

}  // namespace svla_mma
