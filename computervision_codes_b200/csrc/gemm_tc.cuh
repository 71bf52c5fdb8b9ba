// Declarations shared between gemm_tc.cu and the whole-model executor.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tcn {

constexpr int TC_BM = 128;       // frames per tile
constexpr int TC_BK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;  // 6 warps: TMA producer, MMA issuer, 4 x operand-split / epilogue

// ---- raw PTX wrappers (mbarrier, TMA, tcgen05) ------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)0 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a warp, chosen by the hardware (elect.sync; every lane of the warp must reach it).  ptxas then KNOWS that
// exactly one thread runs the guarded region and issues tcgen05.mma / tcgen05.commit / TMA from uniform registers back to
// back.  With `if (lane == 0)` it cannot, and wraps every such instruction in an ELECT / BRA.U.ANY loop over the
// "possibly several" active threads: ~10 SASS instructions per MMA, which made the issuing warp the bottleneck of the
// fused layer kernels (profiles/r1_fused_stress_source_stalls.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM (32 lanes x 32 columns per warp); the caller issues tmem_st_wait() once after its last store
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
}
// 32 lanes x 16 columns per warp
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// kind::tf32 with the A operand in TENSOR MEMORY (row m of A in lane m, K along 32-bit columns; verified on B200 with
// tools/exp/exp_tmem_a.cu), B K-major in shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct GemmTcDev {
  float* Y;
  int ldy;
  int N;                    // valid output columns (tiles of 64 / 128; the tail is masked)
  const float* bias;
  const float* R;           // residual added after dropout (rows, >= N) or nullptr
  int ldr;
  const float* M;           // relu mask: output zeroed where M <= 0, or nullptr
  int ldm;
  int relu;
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int x_unpadded;
  int ntaps;                // 1..3
  int shift[3];             // row offset of each tap
  int kbp;                  // 32-wide k-blocks per tap = ceil(c_in / 32); weights hold ntaps * kbp * 32 columns
  int c_in;
  const float* colscale;    // optional per-(sequence, k) scale of X
  int colscale_ld;
  uint32_t in_drop_thresh;  // optional keep-mask on X elements (row = padded source row, col = channel)
  float in_drop_scale;
  uint32_t in_drop_seed, in_drop_stream;
  uint32_t drop_thresh;     // optional dropout on the output
  float drop_scale;
  uint32_t drop_seed, drop_stream;
};

// (n_out, c_in, ntaps) torch weight -> hi / lo halves in the layout the TMA maps read:
// rows = output column n (padded to a multiple of 64 with zero rows), columns k = tap * kbp*32 + c.
// transpose = 1 swaps the roles for the input-gradient pass (rows = c_in, k = tap * kbp*32 + o).
struct SplitJob {
  long first;    // exclusive prefix of output elements
  long src_off;  // float offset of the torch weight in the parameter buffer
  long dst_off;  // float offset in the hi (and lo) buffer
  int n_out, c_in, ntaps, transpose, rows_pad, kcols;
};
inline int tc_kbp(int k) { return (k + TC_BK - 1) / TC_BK; }
inline long tc_weight_rows(int n_out, int c_in, int transpose) { return ((transpose ? c_in : n_out) + 63) / 64 * 64; }
inline long tc_weight_cols(int n_out, int c_in, int ntaps, int transpose) {
  return (long)ntaps * tc_kbp(transpose ? n_out : c_in) * TC_BK;
}
int make_tensor_map_2d(CUtensorMap* map, const float* ptr, long rows, long cols, long ld, int box_rows,
                       bool atom32 = false);
int launch_split_weight(const float* w, float* whi, float* wlo, int n_out, int c_in, int ntaps, int transpose,
                        cudaStream_t stream);
int launch_split_batched(const SplitJob* jobs_dev, int njobs, const float* params, float* whi, float* wlo, long total,
                         cudaStream_t stream);
// mx32: optional map of the same tensor with 32-row boxes; when given and the problem is a small-dilation k = 3
// convolution over 64 channels the slab kernel is used (each frame staged once per tile instead of once per tap)
// Fused residual layer forward on tcgen05 (gemm_tc.cu: layer_fwd_tc_kernel).  h: epilogue of the dilated conv
// (Y = h buffer or nullptr, bias b1, relu); y: epilogue of the 1x1 conv (Y, bias b2, dropout, residual = layer input);
// y.shift = the three taps, y.meta / nblk / dyn the batch.
struct LayerTcDev {
  GemmTcDev h;
  GemmTcDev y;
  uint32_t* masks;   // nullable: (rows, 4) bit words per frame {h > 0 [0..31], [32..63], dropout keep [0..31], [32..63]}
};
int launch_layer_fwd_tc(const CUtensorMap& mx, const CUtensorMap& w1hi, const CUtensorMap& w1lo, const CUtensorMap& w2hi,
                        const CUtensorMap& w2lo, const LayerTcDev& p, int cap_nblk, cudaStream_t stream, bool pdl);

// Fused residual layer input gradient on tcgen05 (gemm_tc.cu: layer_bwd_tc_kernel):
//   gu = ((keep * gy / (1 - p)) W2) * [h > 0];   gx[t] = gy[t] + sum_k W1_k^T gu[t - s_k]
// gu: store of gu (Y, ldy; meta / nblk / dyn; shift[] = the taps of the transposed conv, -s_k); gx: Y = gx, R = gy.
// masks: the (rows, 4) bit words written by layer_fwd_tc_kernel; drop_scale = 1 / (1 - p) (1 in eval mode).
struct LayerBwdTcDev {
  GemmTcDev gu;
  GemmTcDev gx;
  const uint32_t* masks;
  float drop_scale;
  int use_drop;
};
int launch_layer_bwd_tc(const CUtensorMap& mgy, const CUtensorMap& w2thi, const CUtensorMap& w2tlo,
                        const CUtensorMap& w1thi, const CUtensorMap& w1tlo, const LayerBwdTcDev& p, int cap_nblk,
                        cudaStream_t stream, bool pdl);

int launch_gemm_tc(const CUtensorMap& mx, const CUtensorMap& mwhi, const CUtensorMap& mwlo, const GemmTcDev& p,
                   int cap_nblk, cudaStream_t stream, const CUtensorMap* mx32 = nullptr);
bool gemm_tc_wants_slab(const GemmTcDev& p);

// tcgen05 weight-gradient kernel (wgrad_tc.cu)
constexpr int WG_BOX_ROWS = 32;
struct WgradTcDev {
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int x_unpadded;
  int n_out, c_in, ntaps;
  int shift[3];
  int cbn;          // 32-column blocks per tap = ceil(c_in / 32)   (set by the launcher)
  int row_splits;   // (set by the launcher)
  float* dW;
  float* db;        // nullable
  const float* colscale;
  int colscale_ld;
  uint32_t g_drop_thresh;
  float g_drop_scale;
  uint32_t g_drop_seed, g_drop_stream;
  uint32_t x_drop_thresh;
  float x_drop_scale;
  uint32_t x_drop_seed, x_drop_stream;
  // deterministic mode: every row split stores its partial dW / db to its own slab (pre-zeroed by the caller) instead of
  // adding it to dW / db with fp32 atomics; slab_reduce() then adds the slabs in split order
  float* slab;          // [slab_splits][slab_stride] or nullptr
  long slab_stride;     // floats per split: n_out * c_in * ntaps (+ n_out for db, stored behind dW)
  int slab_splits;      // capacity: the launcher never uses more row splits than this
};
// out[i] += sum over s < nsplit of slab[s * stride + i] in split order (fixed summation tree: bit-identical results)
int launch_slab_reduce(float* out, const float* slab, long n, int nsplit, long stride, cudaStream_t stream);
int wgrad_tc_splits_cap(int n_out, int c_in, int ntaps);   // upper bound of the row splits launch_wgrad_tc chooses
// mx / mg: maps built with make_tensor_map_2d(..., WG_BOX_ROWS, /*atom32=*/true)
int launch_wgrad_tc(const CUtensorMap& mx, const CUtensorMap& mg, WgradTcDev& p, int cap_nblk, cudaStream_t stream);
// Many independent weight-gradient problems in ONE launch (all residual layers of a stage): descriptors (tensor maps +
// parameters, row_splits / cbn filled in by the caller) and the (problem, m-tile) table live in device memory.
struct alignas(64) WgradMultiDesc {
  CUtensorMap mx, mg;
  WgradTcDev p;
};
int launch_wgrad_tc_multi(const WgradMultiDesc* descs_dev, const int2* tiles_dev, int ntiles, int row_splits,
                          cudaStream_t stream);
int launch_wgrad_tc_pair(const CUtensorMap& mx0, const CUtensorMap& mg0, WgradTcDev& p0, const CUtensorMap& mx1,
                         const CUtensorMap& mg1, WgradTcDev& p1, int cap_nblk, cudaStream_t stream);

}  // namespace tcn
