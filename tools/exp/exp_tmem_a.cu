// Experiment: tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st from registers), B in shared
// memory (K-major SWIZZLE_128B, TMA).  Needed to chain conv3 -> relu -> conv1x1 of a residual layer in one kernel
// without staging h through shared memory.  Checks D[m][n] = sum_k A[m][k] B[n][k] for A rows held one per TMEM lane,
// K along TMEM columns.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../computervision_codes_b200/csrc/gemm_tc.cuh"
using namespace tcn;
namespace tcn { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } bool pdl_enabled() { return false; } }

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// A: (128, 64) row-major global; B: (64 n, 64 k) row-major global (two 32-column k-blocks by TMA)
__global__ void __launch_bounds__(128, 1) exp_kernel(const float* __restrict__ A, const __grid_constant__ CUtensorMap map_b,
                                                     float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;      // columns [0, 64): D, [64, 128): A
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384);
    tma_load_2d(tiles, &map_b, &bars[0], 0, 0);          // k-block 0: 64 rows x 32 cols
    tma_load_2d(tiles + 8192, &map_b, &bars[0], 32, 0);   // k-block 1
  }
  // every thread stores its own row of A (lane = row within the warp's 32-lane quadrant)
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float v[32];
    for (int j = 0; j < 32; ++j) v[j] = A[row * 64 + c0 + j];
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    for (int k = 0; k < 8; ++k) {
      const uint64_t db = umma_desc_sw128(base + (k >> 2) * 8192 + (k & 3) * 32);
      umma_tf32_ts(tmem, tmem + 64 + k * 8, db, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j) out[row * 64 + c0 + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128u));
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
  const int M = 128, K = 64, N = 64;
  std::vector<float> A(M * K), B(N * K);
  for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7 + (i / K) * 3) % 17 - 8);
  for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + (i / K)) % 13 - 6);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap mb;
  if (mkmap(&mb, dB, N, K, 64)) { printf("map failed\n"); return 1; }
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 2048);
  std::vector<float> O(M * N);
  cudaMemset(dO, 0, O.size() * 4);
  exp_kernel<<<1, 128, 16384 + 2048>>>(dA, mb, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
    }
  printf("tmem-A mma: maxerr=%g  (O[0][0]=%g O[5][7]=%g)\n", maxerr, O[0], O[5 * 64 + 7]);
  return 0;
}
