// Fixed cost of one launch in a chain of dependent kernels (CUDA graph, programmatic dependent launch), by ingredient:
//   0: empty kernel        1: + 225 KB dynamic smem, mbarrier init, __syncthreads
//   2: + tcgen05.alloc / dealloc of 512 TMEM columns       3: + 128 KB read from L2 into smem (cp.async.bulk)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/launch_floor tools/exp/launch_floor.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ w, float* out) {
  extern __shared__ uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if (MODE >= 1) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (MODE >= 2 && threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&slot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    __syncthreads();
  }
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  if (MODE >= 3) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(131072) : "memory");
      for (int i = 0; i < 16; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 8192, [%2];\n" ::"r"(
                         smem_u32(smem + i * 8192)), "l"(w + i * 2048), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    if (threadIdx.x == 0) out[blockIdx.x] = reinterpret_cast<float*>(smem)[blockIdx.x];
  }
  if (MODE >= 2) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(slot));
  }
}

template <int MODE>
float run(int grid, bool pdl, const float* w, float* out) {
  const int smem = MODE >= 1 ? 225 * 1024 : 0;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  const int N = 200;
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < N; ++i) cudaLaunchKernelEx(&cfg, k<MODE>, w, out);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / (5 * N);
}

int main() {
  float *w, *out;
  cudaMalloc(&w, 1 << 20); cudaMalloc(&out, 4096);
  cudaMemset(w, 0, 1 << 20);
  for (int grid : {15, 148}) {
    for (int pdl = 0; pdl < 2; ++pdl) {
      printf("grid %3d pdl %d: empty %.2f us | +smem/barrier %.2f | +TMEM alloc %.2f | +128 KB weights %.2f\n", grid, pdl,
             run<0>(grid, pdl, w, out), run<1>(grid, pdl, w, out), run<2>(grid, pdl, w, out), run<3>(grid, pdl, w, out));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
