// Shared device helpers for the tcn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tcn_b200.h"

namespace tcn {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> TCN_ERR_CUDA
int num_sms();

#define TCN_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::tcn::set_error(__VA_ARGS__);           \
      return TCN_ERR_INVALID_ARG;              \
    }                                          \
  } while (0)

// Per-128-row block metadata of the packed, time-major activation layout.
//   lo, hi : valid row range [lo, hi) of the sequence that owns this block (padded-row coordinates,
//            lo is a multiple of 128)
//   in_delta: add to a padded row index to address caller-owned, unpadded per-frame inputs
//            (feature rows X[b*T + t], label rows)
//   seq    : sequence index inside the batch
struct __align__(16) BlkMeta {
  int lo, hi, in_delta, seq;
};
constexpr int kBlkRows = 128;

// ----------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared (bypasses L1); !valid writes 16 zero bytes and reads nothing.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// fp32 -> (big, small) tf32 pair, big + small == x to ~2^-21 relative (the 3xTF32 split).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lo) : "f"(r));
}

// D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs, fp32 accumulate.
// Fragment layout (g = lane >> 2, t = lane & 3):
//   a0:(g, t) a1:(g+8, t) a2:(g, t+4) a3:(g+8, t+4);  b0:(k=t, n=g) b1:(k=t+4, n=g)
//   d0:(g, 2t) d1:(g, 2t+1) d2:(g+8, 2t) d3:(g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 3-term split product: acc += a * b with ~fp32 accuracy (small cross terms first).
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                           uint32_t b0hi, uint32_t b1hi, uint32_t b0lo, uint32_t b1lo) {
  mma_tf32(d, alo, b0hi, b1hi);
  mma_tf32(d, ahi, b0lo, b1lo);
  mma_tf32(d, ahi, b0hi, b1hi);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Counter-based dropout RNG: one 32-bit hash per element, keyed by (seed, stream, row, col).
// Forward and backward regenerate the same keep-mask from the same key; no mask tensor is stored.
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t stream, uint32_t row, uint32_t col) {
  uint32_t h = seed ^ (stream * 0xC2B2AE3Du) ^ (row * 0x9E3779B1u) ^ (col * 0x85EBCA77u);
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  h += row;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}
// keep iff hash >= thresh, thresh = p * 2^32 (p = 0.5 -> 0x80000000)
__host__ __device__ __forceinline__ uint32_t drop_thresh(float p) {
  double v = (double)p * 4294967296.0;
  if (v < 0.0) v = 0.0;
  if (v > 4294967295.0) v = 4294967295.0;
  return (uint32_t)v;
}

}  // namespace tcn
