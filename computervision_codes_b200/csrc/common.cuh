// Shared device helpers for the tcn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tcn_b200.h"

namespace tcn {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> TCN_ERR_CUDA; counts the launch on success
long long launch_count();
int num_sms();
bool pdl_enabled();  // TCN_NO_PDL=1 turns programmatic dependent launch off
bool pdl_allowed_on(cudaStream_t stream);  // ... and it is used on capturing streams only (runtime.cu)

// Launch with (optionally) the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_allowed_on(stream)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define TCN_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::tcn::set_error(__VA_ARGS__);           \
      return TCN_ERR_INVALID_ARG;              \
    }                                          \
  } while (0)

#define TCN_CHECK(expr)            \
  do {                             \
    const int rc__ = (expr);       \
    if (rc__ != TCN_OK) return rc__; \
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

// Batch descriptor living in device memory so that a captured CUDA graph can be replayed on batches
// of a different shape: kernels launched with a non-null `dyn` read nblk / num_seqs from here.
struct __align__(16) BatchDesc {
  int nblk, rows, num_seqs, frames;
  unsigned seed;  // dropout seed of this step
  int norm_seqs;  // > 0: the loss / its gradient are normalised by this many sequences instead of num_seqs (data parallel
                  // with unequal shares: every rank divides by the GLOBAL number of videos, the all-reduce then sums)
  int pad[2];
};

// One (weight tensor, orientation) pair of the batched weight preparation.
struct PrepJob {
  long first;    // exclusive prefix of float4 outputs
  long src_off;  // float offset of the torch-layout weight in the flat parameter buffer
  int n_out, c_in, ntaps, transpose, NT8, kpt;
};
int launch_prep_batched(const PrepJob* jobs_dev, int njobs, const float* params, float* wf_base, long total_f4,
                        cudaStream_t stream);

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

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start
// while its stream predecessor is still running; it must execute pdl_wait() before it touches anything the
// predecessor wrote.  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as SMs free up.
// No kernel of this library triggers its dependents early (TCN_PDL_TRIGGER 0): a programmatically launched kernel is
// scheduled when the last CTA of its predecessor has exited -- without the full inter-kernel drain -- runs its prologue
// (barriers, TMEM allocation, step-constant weights) and then waits in griddepcontrol.wait for the predecessor's memory.
// Round 2 first triggered at the very top of every kernel (TCN_PDL_TRIGGER 2: up to nine 15-CTA grids of the 41-layer
// chain resident at once, each spinning in its wait), then right after the wait (1).  With either, a layer occasionally
// computed 1-40 frames from input its predecessor had not stored yet: tools/exp/pdl_graph_check.py (same weights, same
// batches, graph replay) diverges from a serialised run after 10-1000 steps, tools/exp/test_hunt.py shows it on 3-30 %
// of eager first forwards; without the trigger both behave like a run without programmatic launch.  Fences did not help
// (TCN_PDL_FIX bit 0: proxy fence after the wait, bit 1: __threadfence() before a producer CTA exits, bit 2: acquire
// fence after the wait, bit 4: proxy fence on the writer side before a producer CTA exits -- all still diverged and cost up
// to 6 %).  The trigger was worth 2.6 % of the step
// (1.772 -> 1.819 ms in the probe; no programmatic launch at all: 1.883 ms).  DESIGN.md section 3.
#ifndef TCN_PDL_FIX
#define TCN_PDL_FIX 0
#endif
#ifndef TCN_PDL_TRIGGER
#define TCN_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
#if TCN_PDL_FIX & 4
  asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
#endif
#if TCN_PDL_FIX & 5
  asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
#if TCN_PDL_TRIGGER == 1
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_exit_fence() {
#if TCN_PDL_FIX & 2
  __threadfence();
#endif
#if TCN_PDL_FIX & 16
  asm volatile("fence.proxy.async;\n" ::: "memory");   // writer side: generic stores -> a successor's TMA (async proxy) loads
#endif
}
// Call sites sit at the top of the kernels, before pdl_wait() (the original order); a no-op unless TCN_PDL_TRIGGER is 2.
__device__ __forceinline__ void pdl_launch_dependents() {
#if TCN_PDL_TRIGGER == 2
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
#endif
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
// Not volatile: independent accumulator chains must be free to interleave.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[MT][NT] += A * B with the 3-term split (small cross terms first), issued as three sweeps over
// the MT*NT independent accumulators so that consecutive MMAs never depend on each other.
template <int MT, int NT>
__device__ __forceinline__ void mma_block_3xtf32(float (&acc)[MT][NT][4], const uint32_t (&ahi)[MT][4],
                                                 const uint32_t (&alo)[MT][4], const uint32_t (&bhi)[NT][2],
                                                 const uint32_t (&blo)[NT][2], int ntc = NT) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
    if (nt < ntc) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], alo[mt], bhi[nt][0], bhi[nt][1]);
    }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
    if (nt < ntc) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], ahi[mt], blo[nt][0], blo[nt][1]);
    }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
    if (nt < ntc) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], ahi[mt], bhi[nt][0], bhi[nt][1]);
    }
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

// Counter-based dropout RNG keyed by (seed, stream, row, col).  ONE 32-bit hash serves a group of four consecutive
// columns (col >> 2); column k of the group compares the hash rotated left by 8 k bits against the threshold, so each
// column's decision is led by its own byte of the hash while the full 32-bit threshold resolution is kept.  The
// epilogues and operand-split loops, which own four consecutive columns per thread, pay one hash per float4.
// Forward and backward regenerate the same keep-mask from the same key; no mask tensor is stored.
__host__ __device__ __forceinline__ uint32_t drop_hash4(uint32_t seed, uint32_t stream, uint32_t row, uint32_t col4) {
  uint32_t h = seed ^ (stream * 0xC2B2AE3Du) ^ (row * 0x9E3779B1u) ^ (col4 * 0x85EBCA77u);
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
__host__ __device__ __forceinline__ uint32_t drop_rotl(uint32_t h, uint32_t k) {  // rotate left by 8 k bits, k = 0..3
  const uint32_t s = 8u * k;
  return s ? ((h << s) | (h >> (32u - s))) : h;
}
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t stream, uint32_t row, uint32_t col) {
  return drop_rotl(drop_hash4(seed, stream, row, col >> 2), col & 3u);
}
// keep iff hash >= thresh, thresh = p * 2^32 (p = 0.5 -> 0x80000000)
__host__ __device__ __forceinline__ uint32_t drop_thresh(float p) {
  double v = (double)p * 4294967296.0;
  if (v < 0.0) v = 0.0;
  if (v > 4294967295.0) v = 4294967295.0;
  return (uint32_t)v;
}
// scale (1/(1-p)) if kept, 0 if dropped
__device__ __forceinline__ float drop_factor(uint32_t seed, uint32_t stream, uint32_t thresh, float scale, int row,
                                             int col) {
  return (drop_hash(seed, stream, (uint32_t)row, (uint32_t)col) >= thresh) ? scale : 0.f;
}

// the factors of four consecutive columns col .. col + 3, col a multiple of 4: one hash
__device__ __forceinline__ void drop_factor4(uint32_t seed, uint32_t stream, uint32_t thresh, float scale, int row,
                                             int col, float (&f)[4]) {
  const uint32_t h = drop_hash4(seed, stream, (uint32_t)row, (uint32_t)col >> 2);
  f[0] = h >= thresh ? scale : 0.f;
  f[1] = drop_rotl(h, 1) >= thresh ? scale : 0.f;
  f[2] = drop_rotl(h, 2) >= thresh ? scale : 0.f;
  f[3] = drop_rotl(h, 3) >= thresh ? scale : 0.f;
}

// ----------------------------------------------------------------------------------------------
// internal launchers shared by the C ABI wrappers and the whole-model executor (model.cu)
struct TapGemmDev {
  const float* X;
  int ldx;
  int x_unpadded;
  const float* colscale;
  int colscale_ld;
  const float4* Wf;
  const float* bias;
  float* Y;
  int ldy;
  const float* R;
  int ldr;
  const float* M;
  int ldm;
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;  // nullable: overrides nblk at run time
  int kpt;               // padded K per tap (multiple of 8)
  int c_in;
  int n_out;
  int NT8;
  int ntaps;
  int shift[3];
  int relu;
  // dropout on the output (forward) ...
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t drop_seed, drop_stream;
  // ... or on the A operand as it is loaded (backward: gv = keep * gy / (1 - p))
  uint32_t in_drop_thresh;
  float in_drop_scale;
  uint32_t in_drop_seed, in_drop_stream;
};
int launch_tapgemm(TapGemmDev& p, int grid_cap_blocks, cudaStream_t stream);

struct WgradDev {
  const float* G;
  int ldg;
  int g_cols;  // readable columns of G (>= n_out, multiple of 4, pad columns must be zero)
  const float* X;
  int ldx;
  int x_unpadded;
  const float* colscale;
  int colscale_ld;
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int n_out, c_in, ntaps;
  int shift[3];
  float* dW;
  float* db;
  int n_tiles, c_tiles, row_splits;
  uint32_t g_drop_thresh;  // dropout applied to G as it is loaded
  float g_drop_scale;
  uint32_t g_drop_seed, g_drop_stream;
  uint32_t x_drop_thresh;  // keep-mask applied to X as it is loaded (input masking of the projection)
  float x_drop_scale;
  uint32_t x_drop_seed, x_drop_stream;
};
int launch_wgrad(WgradDev& p, int cap_nblk, cudaStream_t stream);

struct BceDev {
  const float* logits;
  int ldl;
  const uint8_t* labels;
  int ldlab;
  int lab_unpadded;
  const BlkMeta* meta;  // nullptr: plain (nrows x K) problem
  int nrows;            // padded rows (meta) or plain rows
  const BatchDesc* dyn; // nullable: rows / num_seqs read at run time (row_scale_const = 1/num_seqs)
  int ncols;
  int zero_cols;
  const float* pos_w;
  const float* col_scale;
  const float* col_unit;
  const int* col_head;
  float row_scale_const;
  float* loss;
  float* dL;
  int lddl;
  float grad_scale;
  // several problems that differ only in the logits / gradient buffers (the four FPN levels) in ONE launch:
  // blockIdx.y selects the level when nlev > 0
  int nlev;
  const float* logits_lv[4];
  float* dL_lv[4];
  // cum_lv != null (with nlev > 0): ONE warp walks the levels of a row (labels read once) and also writes the running sum
  // cum_lv[l] = dL_lv[0] + ... + dL_lv[l]: with shared head weights the gradient w.r.t. FPN level l is cum_lv[l] W (network.py:98-106)
  float* cum_lv[4];
};
int launch_bce(const BceDev& p, int cap_rows, cudaStream_t stream);
int launch_chan_scale(float* out, int ld, int max_seqs, const BatchDesc* dyn, float p, uint32_t seed,
                      uint32_t stream_id, cudaStream_t stream);

struct LayerFwdDev {
  const float* X;   // (rows, 64)
  float* Y;         // (rows, 64)
  float* H;         // (rows, 64) relu(u), saved for backward; nullable (inference)
  const float4* W1f;  // fragment-ordered, KS = 24, NT8 = 8
  const float4* W2f;  // KS = 8, NT8 = 8
  const float* b1;
  const float* b2;
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int shift[3];
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t drop_seed, drop_stream;
};
int launch_layer_fwd64(const LayerFwdDev& p, int cap_nblk, cudaStream_t stream);

}  // namespace tcn
