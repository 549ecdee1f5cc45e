// Tensor Memory (tcgen05.alloc / ld / st) used as a software-managed register
// file by the fused solver kernels (rpie_fast.cu, multislice_fused.cu).
#pragma once

#include "common.cuh"

namespace tb {

// ---- Tensor Memory as a software-managed register file ----------------------
// The object-gradient accumulator and the interpolated patch of the current
// position (KMAX complex per thread each) live in TMEM (256 KiB per SM,
// private to the CTA) instead of registers: each warp owns the 32 TMEM lanes
// of its quadrant (warp % 4) and a private range of columns.  That frees
// 4*KMAX registers per thread: no spills at 128 registers, no patch re-read in
// the gradient sweep, and room to batch global loads.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_slot);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// issue only: the values are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
        "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]),
        "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
// The wait names the loaded registers as in/out operands so that the compiler
// cannot move their first use above it.
__device__ __forceinline__ void tmem_wait_ld(float (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]),
                 "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]),
                 "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  tmem_ld16_issue(taddr, v);
  tmem_wait_ld(v);
}
// 8-column variants (4 complex values per thread): half the transient registers
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
        "=f"(v[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]),
                 "+f"(v[6]), "+f"(v[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}

}  // namespace tb
