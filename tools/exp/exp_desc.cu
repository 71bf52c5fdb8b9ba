// Experiment: can a K-major SWIZZLE_128B UMMA descriptor start at a row that is not a multiple of 8
// (needed to read dilation taps as row-offset views of ONE staged slab)?  Tests base_offset = (row & 7) vs 0.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../computervision_codes_b200/csrc/gemm_tc.cuh"
using namespace tcn;
namespace tcn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } bool pdl_enabled() { return false; } }

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ CUtensorMap map_a,
                                                     const __grid_constant__ CUtensorMap map_b, float* out, int r0,
                                                     int use_bo) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + 32768 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 32768 + 8192);
    tma_load_2d(tiles, &map_a, &bars[0], 0, 0);            // 256 rows x 32 cols
    tma_load_2d(tiles + 32768, &map_b, &bars[0], 0, 0);    // 64 rows x 32 cols
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = base + r0 * 128 + k * 32;
      const uint64_t da = desc_sw128_bo(a_addr, use_bo ? (uint32_t)(r0 & 7) : 0u);
      const uint64_t db = umma_desc_sw128(base + 32768 + k * 32);
      umma_tf32(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(64u));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int mkmap(CUtensorMap* m, float* p, long rows, long cols, int box_rows) {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  return ((EncodeTiledFn)sym)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  const int RA = 256, K = 32, N = 64;
  std::vector<float> A(RA * K), B(N * K);
  for (int i = 0; i < RA * K; ++i) A[i] = (float)((i * 7 + (i / K) * 3) % 17 - 8);
  for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (mkmap(&ma, dA, RA, K, 256) || mkmap(&mb, dB, N, K, 64)) { printf("map failed\n"); return 1; }
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 2048);
  std::vector<float> O(128 * 64);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int r0 : {0, 8, 1, 3, 4, 13, 64, 77}) {
      cudaMemset(dO, 0, O.size() * 4);
      exp_kernel<<<1, 128, 49152 + 2048>>>(ma, mb, dO, r0, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("r0=%d bo=%d CUDA error %s\n", r0, use_bo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[(r0 + m) * K + k] * B[n * K + k];
          maxerr = fmax(maxerr, fabs(ref - O[m * 64 + n]));
        }
      printf("r0=%3d base_offset=%s maxerr=%g\n", r0, use_bo ? "row&7" : "0", maxerr);
    }
  return 0;
}
