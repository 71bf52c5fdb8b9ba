// Fused dilated residual layer, forward, 64 channels:
//
//     y = x + Dropout_p( W2 . relu( W1 (*)_d x + b1 ) + b2 )          one launch per layer
//
// DilatedResidualLayer.forward (MT4MTLKD/Temporal_tenco/network.py:193-198, taps t-d, t, t+d) and
// DilatedResidualCausalLayer.forward (network.py:178-183, taps t-2d, t-d, t); TERL duplicates at
// TERL/0_5fold_TCN_black/network.py:199-234.
//
// One CTA owns a 64-frame time tile.  The time-slab x[t0+s0 .. t0+64+s2) (tile plus its dilation
// halo; three disjoint 64-frame slabs once the dilation exceeds the tile) is staged in shared memory
// with 128-bit cp.async loads, zero-filled outside the sequence.  Each warp then runs both
// contractions for its 16 frames on the tensor cores (mma.sync TF32, 3-term split => fp32-level
// accuracy): u (16x64, K = 3x64) -> bias, ReLU in registers -> the accumulator fragments are
// re-laid into A fragments with quad shuffles -> v (16x64, K = 64) -> bias, dropout, + x (exact fp32
// from the slab) -> 64-bit stores of y.  h = relu(u) is also written once, for the backward pass.
// Weights come fragment-ordered and pre-split from global memory (L1/L2 resident, read-only path).
//
// Algorithmic HBM bytes per frame: read x (256 B) + write y (256 B) [+ write h (256 B) in training].
#include "common.cuh"

namespace tcn {

constexpr int LF_C = 64;
constexpr int LF_TM = 64;
constexpr int LF_LD = 68;           // smem row stride (floats): bank = (4g + t) -> conflict-free fragments
constexpr int LF_THREADS = 128;     // 4 warps x 16 rows
constexpr int LF_SLAB_ROWS = 3 * LF_TM;
constexpr int LF_SMEM = LF_SLAB_ROWS * LF_LD * 4;

__global__ void __launch_bounds__(LF_THREADS, 3) layer_fwd64_kernel(const LayerFwdDev p) {
  extern __shared__ __align__(16) float slab[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const int total_tiles = nblk * 2;
  const int s0 = p.shift[0], s1 = p.shift[1], s2 = p.shift[2];
  const bool halo = (s2 - s0) <= 2 * LF_TM;  // tile + halo fits the 3*TM-row slab
  const int slab_rows = halo ? (LF_TM + (s2 - s0)) : LF_SLAB_ROWS;
  const int tb1 = halo ? (s1 - s0) : LF_TM, tb2 = halo ? (s2 - s0) : 2 * LF_TM;  // slab row of tap k, frame 0
  const int zbase = (s1 == 0) ? tb1 : ((s2 == 0) ? tb2 : 0);                      // the tap with shift 0
  const uint32_t seed = p.drop_seed ^ (p.dyn ? p.dyn->seed : 0u);
  pdl_launch_dependents();
  pdl_wait();

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int blk = tile >> 1;
    const BlkMeta m = p.meta[blk];
    const int row0 = blk * kBlkRows + (tile & 1) * LF_TM;
    if (row0 >= m.hi) continue;  // CTA-uniform

    // ---- stage the slab
    for (int piece = tid; piece < slab_rows * (LF_C / 4); piece += LF_THREADS) {
      const int j = piece >> 4, cq = (piece & 15) * 4;
      int src;
      if (halo) {
        src = row0 + s0 + j;
      } else {
        const int k = j >> 6;  // LF_TM == 64
        src = row0 + (k == 0 ? s0 : (k == 1 ? s1 : s2)) + (j & 63);
      }
      const bool valid = (src >= m.lo) && (src < m.hi);
      const float* gp = valid ? (p.X + (size_t)src * LF_C + cq) : p.X;
      cp_async16(&slab[j * LF_LD + cq], gp, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- GEMM 1: u = sum_taps x[t + s_k] W1_k^T      (K = 3 x 64)
    float acc[1][8][4];
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[0][b][c] = 0.f;

#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      const float* Ab = slab + ((k == 0 ? 0 : (k == 1 ? tb1 : tb2)) + warp * 16 + g) * LF_LD + t;
#pragma unroll 2
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t ahi[1][4], alo[1][4], bhi[8][2], blo[8][2];
        const float4* wp = p.W1f + ((size_t)(k * 8 + kk) * 8) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float4 w = __ldg(wp + nt * 32);
          bhi[nt][0] = __float_as_uint(w.x); bhi[nt][1] = __float_as_uint(w.y);
          blo[nt][0] = __float_as_uint(w.z); blo[nt][1] = __float_as_uint(w.w);
        }
        const float* ap = Ab + kk * 8;
        split_tf32(ap[0], ahi[0][0], alo[0][0]);
        split_tf32(ap[8 * LF_LD], ahi[0][1], alo[0][1]);
        split_tf32(ap[4], ahi[0][2], alo[0][2]);
        split_tf32(ap[8 * LF_LD + 4], ahi[0][3], alo[0][3]);
        mma_block_3xtf32<1, 8>(acc, ahi, alo, bhi, blo);
      }
    }

    // ---- bias + ReLU (registers); write h for the backward pass
    const int r_lo = row0 + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      const float2 b = __ldg(reinterpret_cast<const float2*>(p.b1 + col));
      acc[0][nt][0] = fmaxf(acc[0][nt][0] + b.x, 0.f);
      acc[0][nt][1] = fmaxf(acc[0][nt][1] + b.y, 0.f);
      acc[0][nt][2] = fmaxf(acc[0][nt][2] + b.x, 0.f);
      acc[0][nt][3] = fmaxf(acc[0][nt][3] + b.y, 0.f);
      if (p.H != nullptr) {
        if (r_lo < m.hi)
          *reinterpret_cast<float2*>(p.H + (size_t)r_lo * LF_C + col) = make_float2(acc[0][nt][0], acc[0][nt][1]);
        if (r_hi < m.hi)
          *reinterpret_cast<float2*>(p.H + (size_t)r_hi * LF_C + col) = make_float2(acc[0][nt][2], acc[0][nt][3]);
      }
    }

    // ---- GEMM 2: v = h W2^T   (K = 64); A fragments rebuilt from the accumulator layout by quad shuffles
    float acc2[1][8][4];
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc2[0][b][c] = 0.f;
    const int src_a = (lane & ~3) | (t >> 1), src_b = src_a + 2;
    const bool odd = (t & 1) != 0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t ahi[1][4], alo[1][4], bhi[8][2], blo[8][2];
      const float4* wp = p.W2f + ((size_t)kk * 8) * 32 + lane;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float4 w = __ldg(wp + nt * 32);
        bhi[nt][0] = __float_as_uint(w.x); bhi[nt][1] = __float_as_uint(w.y);
        blo[nt][0] = __float_as_uint(w.z); blo[nt][1] = __float_as_uint(w.w);
      }
      float a[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float v0 = acc[0][kk][half * 2], v1 = acc[0][kk][half * 2 + 1];
        const float x0 = __shfl_sync(0xffffffffu, v0, src_a), x1 = __shfl_sync(0xffffffffu, v1, src_a);
        const float y0 = __shfl_sync(0xffffffffu, v0, src_b), y1 = __shfl_sync(0xffffffffu, v1, src_b);
        a[half] = odd ? x1 : x0;      // (row g + 8*half, col t)
        a[2 + half] = odd ? y1 : y0;  // (row g + 8*half, col t + 4)
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(a[e], ahi[0][e], alo[0][e]);
      mma_block_3xtf32<1, 8>(acc2, ahi, alo, bhi, blo);
    }

    // ---- bias, dropout, residual (exact fp32 x from the slab), store
    const float* xz = slab + (zbase + warp * 16 + g) * LF_LD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      const float2 b = __ldg(reinterpret_cast<const float2*>(p.b2 + col));
      float v00 = acc2[0][nt][0] + b.x, v01 = acc2[0][nt][1] + b.y;
      float v10 = acc2[0][nt][2] + b.x, v11 = acc2[0][nt][3] + b.y;
      if (p.drop_thresh != 0u) {
        v00 *= drop_factor(seed, p.drop_stream, p.drop_thresh, p.drop_scale, r_lo, col);
        v01 *= drop_factor(seed, p.drop_stream, p.drop_thresh, p.drop_scale, r_lo, col + 1);
        v10 *= drop_factor(seed, p.drop_stream, p.drop_thresh, p.drop_scale, r_hi, col);
        v11 *= drop_factor(seed, p.drop_stream, p.drop_thresh, p.drop_scale, r_hi, col + 1);
      }
      const float2 xa = *reinterpret_cast<const float2*>(xz + col);
      const float2 xb = *reinterpret_cast<const float2*>(xz + 8 * LF_LD + col);
      if (r_lo < m.hi)
        *reinterpret_cast<float2*>(p.Y + (size_t)r_lo * LF_C + col) = make_float2(xa.x + v00, xa.y + v01);
      if (r_hi < m.hi)
        *reinterpret_cast<float2*>(p.Y + (size_t)r_hi * LF_C + col) = make_float2(xb.x + v10, xb.y + v11);
    }
    __syncthreads();  // the slab is reused by the next tile
  }
}

int launch_layer_fwd64(const LayerFwdDev& p, int cap_nblk, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(layer_fwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM);
    if (e != cudaSuccess) {
      set_error("layer_fwd64: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  const int nb = cap_nblk > 0 ? cap_nblk : p.nblk;
  long tiles = (long)nb * 2;
  const long cap = (long)num_sms() * 3;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  launch_kernel(layer_fwd64_kernel, dim3((int)tiles), dim3(LF_THREADS), LF_SMEM, stream, true, p);
  return check_launch("layer_fwd64_kernel");
}

}  // namespace tcn

using namespace tcn;

extern "C" int tcn_layer_fwd(const tcn_layer_fwd_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->x && a->y && a->w1f && a->w2f && a->b1 && a->b2 && a->meta, "tcn_layer_fwd: null pointer");
  if (a->channels != LF_C) {
    set_error("tcn_layer_fwd: the fused kernel is built for 64 channels (got %d); use tcn_tapgemm", a->channels);
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->nblk > 0, "tcn_layer_fwd: empty problem");
  TCN_REQUIRE(a->shift[0] < a->shift[1] && a->shift[1] < a->shift[2] &&
                  (a->shift[0] == 0 || a->shift[1] == 0 || a->shift[2] == 0),
              "tcn_layer_fwd: shifts must be increasing and contain 0");
  TCN_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "tcn_layer_fwd: drop_p must be in [0, 1)");
  LayerFwdDev p;
  p.X = a->x; p.Y = a->y; p.H = a->h;
  p.W1f = reinterpret_cast<const float4*>(a->w1f); p.W2f = reinterpret_cast<const float4*>(a->w2f);
  p.b1 = a->b1; p.b2 = a->b2;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  for (int i = 0; i < 3; ++i) p.shift[i] = a->shift[i];
  p.drop_thresh = a->drop_p > 0.f ? drop_thresh(a->drop_p) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = a->drop_seed; p.drop_stream = a->drop_stream;
  return launch_layer_fwd64(p, 0, (cudaStream_t)stream);
}
