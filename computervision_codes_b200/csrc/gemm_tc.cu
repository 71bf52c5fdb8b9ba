// tcgen05 + TMA "tap GEMM": per-frame dense contractions on the 5th-generation tensor cores
//
//     Y[r, n] = epi( sum_tap sum_c X[r + shift[tap], c] * W[n, c, tap] + bias[n] )      (fp32 in, fp32 out)
//     epi = relu -> (* [M > 0]) -> dropout -> (+ R)
//
//  - the stage-input projection Conv1d(dim -> num_f_maps, 1) of BaseCausalTCN
//    (MT4MTLKD/Temporal_tenco/network.py:113,129; x arrives as (B, T, D) = K-major rows, :42), with
//    Dropout2d's channel scale (:125-127) and the 25 % input mask (:43-50) folded into the operand load;
//  - the k=3 dilated convolution and the 1x1 convolution of the residual layers and their input-gradient
//    passes (network.py:178-198): each tap is the same TMA box fetched at a shifted row coordinate;
//  - the FPN lateral (network.py:98-106), the four heads (:63-67) and their gradients;
//  - nn.Linear / Conv1d(k=1,3) of the MS-TCT blocks (Temporal_mstct/MSTCT/Temporal_Encoder.py:12,15,57-59,139).
//
// Kernels in this file (all: TMA producer warp, one MMA-issuing lane, operand-split warps, epilogue warps, mbarrier
// rings, accumulators in TMEM, 3 x TF32 products per k-slice: lo*hi + hi*lo + hi*hi):
//   gemm_tc_kernel          streaming: long contractions (projection, K = D); raw X ring, A operand through TMEM
//   gemm_tc_persist_kernel  persistent, <= 6 k-blocks of pre-split weights resident in shared memory (1x1 / k=3 convs,
//                           their input gradients, FPN lateral, heads); double-buffered TMEM accumulator
//   gemm_tc_slab_kernel     persistent, the three dilation taps read as row-offset views of ONE staged time slab
//   gemm_tc_wide_kernel     persistent over (frame tile, 128 / 256-column tile): the MS-TCT Linear layers
//   layer_fwd_tc_kernel     fused residual layer forward: both GEMMs take A from TMEM, h never leaves the SM
// Operand split: hi = x & 0xffffe000 is exact in TF32, lo = x - hi; rows whose tap leaves the sequence are zeroed
// there (Conv1d zero padding / causal F.pad).  Epilogues: tcgen05.ld (32 lanes x 32 columns) -> XOR-swizzled
// shared-memory transpose -> bias / ReLU / ReLU-mask / dropout / residual -> 128-bit coalesced stores.
// Accuracy: 3xTF32 == fp32 to ~1e-6 relative (see tests), the bar is 1e-3 on logits.
#include <cstdlib>

#include "gemm_tc.cuh"

namespace tcn {


// Operand split of one 128 x 32 fp32 tile (TMA layout, 128B swizzle) by NT = 128 / 256 threads: X -> (X_hi in place, X_lo).
// All eight 16-byte chunks of a thread are loaded before any is processed; rows whose tap source leaves the
// sequence are zeroed with selects (no divergent control flow on the fast path).
template <int NT>
__device__ __forceinline__ void split_tile(float4* __restrict__ xa, float4* __restrict__ xl, int ct, int row0, int sh,
                                           const BlkMeta& m, int kc, const GemmTcDev& p, uint32_t in_seed) {
  constexpr int NV = 1024 / NT;   // 16-byte chunks per thread (128 rows x 8 chunks, NT threads)
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xa[ct + i * NT];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int src = row0 + ((ct + i * NT) >> 3) + sh;
    // select, not multiply: a pad row may hold anything (0 * NaN would poison the accumulator)
    if (!(src >= m.lo && src < m.hi)) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.colscale != nullptr || p.in_drop_thresh != 0u) {  // CTA-uniform, rare (projection in training)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = ct + i * NT, r = c >> 3;
      const int src = row0 + r + sh;
      const int col = kc * TC_BK + (((c & 7) ^ (r & 7)) << 2);  // undo the 128B swizzle
      if (p.colscale != nullptr && col < p.c_in) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.colscale + (size_t)m.seq * p.colscale_ld + col));
        v[i].x *= sc.x; v[i].y *= sc.y; v[i].z *= sc.z; v[i].w *= sc.w;
      }
      if (p.in_drop_thresh != 0u) {
        float f[4];
        drop_factor4(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, src, col, f);
        v[i].x *= f[0]; v[i].y *= f[1]; v[i].z *= f[2]; v[i].w *= f[3];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u); l.x = v[i].x - h.x;
    h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u); l.y = v[i].y - h.y;
    h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u); l.z = v[i].z - h.z;
    h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u); l.w = v[i].w - h.w;
    xa[ct + i * NT] = h;
    xl[ct + i * NT] = l;
  }
}

// Epilogue of four consecutive output columns of one row: bias, relu, relu-mask, dropout, residual, store.
__device__ __forceinline__ void epilogue_vec4(float (&o)[4], int row, int n, const GemmTcDev& p, uint32_t out_seed,
                                              bool vec_ok) {
  if (n >= p.N) return;
  const bool full4 = (n + 3 < p.N) && vec_ok;
  if (full4) {
    if (p.bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
    }
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    if (p.M != nullptr) {
      const float4 mk = *reinterpret_cast<const float4*>(p.M + (size_t)row * p.ldm + n);
      if (!(mk.x > 0.f)) o[0] = 0.f;
      if (!(mk.y > 0.f)) o[1] = 0.f;
      if (!(mk.z > 0.f)) o[2] = 0.f;
      if (!(mk.w > 0.f)) o[3] = 0.f;
    }
    if (p.drop_thresh != 0u) {
      float f[4];
      drop_factor4(out_seed, p.drop_stream, p.drop_thresh, p.drop_scale, row, n, f);
      o[0] *= f[0]; o[1] *= f[1]; o[2] *= f[2]; o[3] *= f[3];
    }
    if (p.R != nullptr) {
      const float4 rr = *reinterpret_cast<const float4*>(p.R + (size_t)row * p.ldr + n);
      o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
    }
    *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (n + e < p.N) {
        float x = o[e];
        if (p.bias != nullptr) x += __ldg(p.bias + n + e);
        if (p.relu) x = fmaxf(x, 0.f);
        if (p.M != nullptr && !(p.M[(size_t)row * p.ldm + n + e] > 0.f)) x = 0.f;
        if (p.drop_thresh != 0u) x *= drop_factor(out_seed, p.drop_stream, p.drop_thresh, p.drop_scale, row, n + e);
        if (p.R != nullptr) x += p.R[(size_t)row * p.ldr + n + e];
        p.Y[(size_t)row * p.ldy + n + e] = x;
      }
    }
  }
}

// Row-per-thread epilogue of 32 accumulator columns (streaming kernel: one tile per CTA).
__device__ __forceinline__ void epilogue_cols(const float (&v)[32], int row, int n_base, const GemmTcDev& p,
                                              uint32_t out_seed, bool vec_ok) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    float o[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
    epilogue_vec4(o, row, n_base + j, p, out_seed, vec_ok);
  }
}

// Coalesced epilogue of a warp's 32 x 64 accumulator block: the TMEM layout (one row per lane) is
// transposed through a private, XOR-swizzled 8 KB staging buffer so that every global access of the
// epilogue (mask / residual loads, output stores) covers two full 256-byte rows per warp instruction.
__device__ __forceinline__ void epilogue_block_coalesced(const float (&v0)[32], const float (&v1)[32], float4* st,
                                                         int lane, int row_base, int row_hi, int n_base,
                                                         const GemmTcDev& p, uint32_t out_seed, bool vec_ok) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int cs = (c & 8) | ((c & 7) ^ (lane & 7));
    st[lane * 16 + cs] = (c < 8) ? make_float4(v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3])
                                 : make_float4(v1[4 * c - 32], v1[4 * c - 31], v1[4 * c - 30], v1[4 * c - 29]);
  }
  __syncwarp();
  if (vec_ok && n_base + 64 <= p.N) {
    // fast path (full 64-column block): all mask / residual loads of 8 float4 are in flight before they are used
    const uint32_t th = p.drop_thresh;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 a[8], mk[8], rr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = (half * 8 + i) * 32 + lane;
        const int r = f >> 4, c = f & 15;
        const int rowc = min(row_base + r, row_hi - 1);  // clamp: loads stay in bounds, the store is predicated
        a[i] = st[r * 16 + ((c & 8) | ((c & 7) ^ (r & 7)))];
        if (p.M != nullptr) mk[i] = *reinterpret_cast<const float4*>(p.M + (size_t)rowc * p.ldm + n_base + c * 4);
        if (p.R != nullptr) rr[i] = *reinterpret_cast<const float4*>(p.R + (size_t)rowc * p.ldr + n_base + c * 4);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = (half * 8 + i) * 32 + lane;
        const int r = f >> 4, c = f & 15;
        const int row = row_base + r, n = n_base + c * 4;
        float o[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
        if (p.bias != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        if (p.M != nullptr) {
          if (!(mk[i].x > 0.f)) o[0] = 0.f;
          if (!(mk[i].y > 0.f)) o[1] = 0.f;
          if (!(mk[i].z > 0.f)) o[2] = 0.f;
          if (!(mk[i].w > 0.f)) o[3] = 0.f;
        }
        if (th != 0u) {
          float f[4];
          drop_factor4(out_seed, p.drop_stream, th, p.drop_scale, row, n, f);
          o[0] *= f[0]; o[1] *= f[1]; o[2] *= f[2]; o[3] *= f[3];
        }
        if (p.R != nullptr) { o[0] += rr[i].x; o[1] += rr[i].y; o[2] += rr[i].z; o[3] += rr[i].w; }
        if (row < row_hi)
          *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < 16; ++i) {
      const int f = i * 32 + lane;
      const int r = f >> 4, c = f & 15;
      const float4 a = st[r * 16 + ((c & 8) | ((c & 7) ^ (r & 7)))];
      const int row = row_base + r;
      if (row < row_hi) {
        float o[4] = {a.x, a.y, a.z, a.w};
        epilogue_vec4(o, row, n_base + c * 4, p, out_seed, vec_ok);
      }
    }
  }
  __syncwarp();
}

// Coalesced epilogue of a warp's 32 x 32 accumulator block (one row per lane in `v`) through a private 4 KB staging
// buffer (XOR-swizzled 16-byte chunks: conflict-free both ways); every global access covers four full 128-byte row
// segments per warp instruction.  Compile-time feature set, bias always on.  Used by the fused layer kernel, whose
// epilogue warps each own 32 columns.
// residual rows of a 32 x 32 block, loaded ahead of the accumulator (the L2 round trip overlaps the MMAs)
__device__ __forceinline__ void epi32_load_residual(float4 (&rr)[8], int lane, int row_base, int row_hi, int n_base,
                                                    const GemmTcDev& p) {
  const int n = n_base + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rowc = min(row_base + i * 4 + (lane >> 3), row_hi - 1);  // clamp: loads stay in bounds
    rr[i] = *reinterpret_cast<const float4*>(p.R + (size_t)rowc * p.ldr + n);
  }
}

template <bool RELU, bool RES, bool DROP, bool BIAS = true>
__device__ __forceinline__ void epilogue_block32(const float (&v)[32], float4* st, int lane, int row_base, int row_hi,
                                                 int n_base, const GemmTcDev& p, uint32_t out_seed,
                                                 const float4 (&rr)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  __syncwarp();
  const int c = lane & 7, n = n_base + c * 4;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (BIAS) b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), row = row_base + r;
    const float4 a = st[r * 8 + (c ^ (r & 7))];
    float o[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    if (RELU) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    if (DROP) {
      if (p.drop_thresh != 0u) {
        float f[4];
        drop_factor4(out_seed, p.drop_stream, p.drop_thresh, p.drop_scale, row, n, f);
        o[0] *= f[0]; o[1] *= f[1]; o[2] *= f[2]; o[3] *= f[3];
      }
    }
    if (RES) { o[0] += rr[i].x; o[1] += rr[i].y; o[2] += rr[i].z; o[3] += rr[i].w; }
    if (row < row_hi) *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncwarp();
}

// The same block epilogue with the dropout decision taken from a keep word per row instead of the hash: `keepw` is the
// word of the row this lane owns in the accumulator layout (bit j: column n_base + j is kept); after the transpose a lane
// writes rows i * 4 + lane / 8, so it fetches their words with a shuffle.  y = R + keep * scale * (acc + bias).
__device__ __forceinline__ void epilogue_block32_keepbits(const float (&v)[32], float4* st, int lane, int row_base,
                                                          int row_hi, int n_base, const GemmTcDev& p, uint32_t keepw,
                                                          float scale, const float4 (&rr)[8]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  __syncwarp();
  const int c = lane & 7, n = n_base + c * 4;
  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + (lane >> 3), row = row_base + r;
    const uint32_t bits = __shfl_sync(0xffffffffu, keepw, r) >> (c * 4);
    const float4 a = st[r * 8 + (c ^ (r & 7))];
    float4 o;
    o.x = (bits & 1u) ? (a.x + b.x) * scale : 0.f;
    o.y = (bits & 2u) ? (a.y + b.y) * scale : 0.f;
    o.z = (bits & 4u) ? (a.z + b.z) * scale : 0.f;
    o.w = (bits & 8u) ? (a.w + b.w) * scale : 0.f;
    o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w;
    if (row < row_hi) *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + n) = o;
  }
  __syncwarp();
}

// Same 32 x 32 block epilogue with the feature set read from `p` at run time (bias / ReLU / ReLU-mask / dropout /
// residual), plus the scalar path for ragged column counts.  Used by the persistent kernels, whose eight epilogue warps
// each own one TMEM lane quadrant x 32 columns.
struct Epi32Pre {
  float4 mk[8], rr[8];
  bool fast;
};
// ReLU-mask / residual rows of the block, loaded BEFORE the accumulator is waited for
__device__ __forceinline__ void epi32_prefetch(Epi32Pre& pre, int lane, int row_base, int row_hi, int n_base,
                                               const GemmTcDev& p, bool vec_ok) {
  pre.fast = vec_ok && n_base + 32 <= p.N;
  if (!pre.fast) return;
  const int n = n_base + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rowc = min(row_base + i * 4 + (lane >> 3), row_hi - 1);  // clamp: loads stay in bounds
    if (p.M != nullptr) pre.mk[i] = *reinterpret_cast<const float4*>(p.M + (size_t)rowc * p.ldm + n);
    if (p.R != nullptr) pre.rr[i] = *reinterpret_cast<const float4*>(p.R + (size_t)rowc * p.ldr + n);
  }
}

__device__ __forceinline__ void epilogue_block32_rt(const float (&v)[32], float4* st, int lane, int row_base, int row_hi,
                                                    int n_base, const GemmTcDev& p, uint32_t out_seed, bool vec_ok,
                                                    const Epi32Pre& pre) {
  if (n_base >= p.N) return;  // warp-uniform
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  __syncwarp();
  const int c = lane & 7, n = n_base + c * 4;
  if (pre.fast) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const uint32_t th = p.drop_thresh;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3), row = row_base + r;
      const float4 a = st[r * 8 + (c ^ (r & 7))];
      float o[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
      if (p.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
      }
      if (p.M != nullptr) {
        if (!(pre.mk[i].x > 0.f)) o[0] = 0.f;
        if (!(pre.mk[i].y > 0.f)) o[1] = 0.f;
        if (!(pre.mk[i].z > 0.f)) o[2] = 0.f;
        if (!(pre.mk[i].w > 0.f)) o[3] = 0.f;
      }
      if (th != 0u) {
        float f[4];
        drop_factor4(out_seed, p.drop_stream, th, p.drop_scale, row, n, f);
        o[0] *= f[0]; o[1] *= f[1]; o[2] *= f[2]; o[3] *= f[3];
      }
      if (p.R != nullptr) { o[0] += pre.rr[i].x; o[1] += pre.rr[i].y; o[2] += pre.rr[i].z; o[3] += pre.rr[i].w; }
      if (row < row_hi) *reinterpret_cast<float4*>(p.Y + (size_t)row * p.ldy + n) = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3), row = row_base + r;
      const float4 a = st[r * 8 + (c ^ (r & 7))];
      if (row < row_hi) {
        float o[4] = {a.x, a.y, a.z, a.w};
        epilogue_vec4(o, row, n, p, out_seed, vec_ok);
      }
    }
  }
  __syncwarp();
}

// one frame (= one thread = one TMEM lane) of a raw 128 x 32 TMA tile -> TF32 hi / lo halves in a TMEM operand slot
// (hi: 32 columns at taddr, lo: the next 32).  Same masking / scaling rules as split_tile.
__device__ __forceinline__ void split_row_to_tmem(const float4* __restrict__ tile, int r, int src, const BlkMeta& m,
                                                  int kc, const GemmTcDev& p, uint32_t in_seed, uint32_t taddr) {
  const float4* row = tile + r * 8;
  const bool inside = src >= m.lo && src < m.hi;
  float hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 v = row[c ^ (r & 7)];                   // undo the 128B swizzle: logical 16-byte chunk c
    if (!inside) v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.colscale != nullptr || p.in_drop_thresh != 0u) {  // CTA-uniform, rare (projection in training)
      const int col = kc * TC_BK + c * 4;
      if (p.colscale != nullptr && col < p.c_in) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(p.colscale + (size_t)m.seq * p.colscale_ld + col));
        v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
      }
      if (p.in_drop_thresh != 0u) {
        float f[4];
        drop_factor4(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, src, col, f);
        v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
      }
    }
    hi[4 * c + 0] = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo[4 * c + 0] = v.x - hi[4 * c + 0];
    hi[4 * c + 1] = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo[4 * c + 1] = v.y - hi[4 * c + 1];
    hi[4 * c + 2] = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo[4 * c + 2] = v.z - hi[4 * c + 2];
    hi[4 * c + 3] = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo[4 * c + 3] = v.w - hi[4 * c + 3];
  }
  tmem_st32(taddr, hi);
  tmem_st32(taddr + 32, lo);
}

// The same for HALF a row (16 of the 32 columns of the k-block): two threads per frame, twice the split warps.  `cs`: the
// Dropout2d scale of this tile's sequence, staged in shared memory by the caller (nullptr: none).
__device__ __forceinline__ void split_half_row_to_tmem(const float4* __restrict__ tile, int r, int half, int src,
                                                       const BlkMeta& m, int kc, const GemmTcDev& p, const float* cs,
                                                       uint32_t in_seed, uint32_t taddr) {
  const float4* row = tile + r * 8;
  const bool inside = src >= m.lo && src < m.hi;
  float hi[16], lo[16];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c = half * 4 + cc;
    float4 v = row[c ^ (r & 7)];                   // undo the 128B swizzle: logical 16-byte chunk c
    if (!inside) v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int col = kc * TC_BK + c * 4;
    if (cs != nullptr && col < p.c_in) {
      const float4 sc = *reinterpret_cast<const float4*>(cs + col);   // same address for the whole warp: one broadcast read
      v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
    }
    if (p.in_drop_thresh != 0u) {
      float f[4];
      drop_factor4(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, src, col, f);
      v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
    }
    hi[4 * cc + 0] = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo[4 * cc + 0] = v.x - hi[4 * cc + 0];
    hi[4 * cc + 1] = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo[4 * cc + 1] = v.y - hi[4 * cc + 1];
    hi[4 * cc + 2] = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo[4 * cc + 2] = v.z - hi[4 * cc + 2];
    hi[4 * cc + 3] = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo[4 * cc + 3] = v.w - hi[4 * cc + 3];
  }
  tmem_st16(taddr + half * 16, hi);
  tmem_st16(taddr + 32 + half * 16, lo);
}

// Streaming kernel (long contractions: the stage-input projection, K = D): one (frame tile, 64-column tile) per CTA.
// The raw X tile and the pre-split weight tiles of a k-block travel through a 6-deep TMA ring (32 KB per stage: the
// ring only holds RAW activations, so 96 KB of X are in flight per SM -- the projection is bound by HBM latency x
// bandwidth); one split thread per frame moves the hi / lo halves into a double-buffered TMEM operand slot and the
// MMAs take A from tensor memory.
template <int BN>
struct TcSmem {
  static constexpr int kStages = 6;
  static constexpr int kA = TC_BM * TC_BK * 4;  // 16384: raw X tile
  static constexpr int kB = BN * TC_BK * 4;     // one weight half
  static constexpr int kStage = kA + 2 * kB;
  static constexpr int kScale = 16384;          // Dropout2d scale of the tile's sequence (up to 4096 input channels)
  static constexpr int kBytes = kStages * kStage + kScale + 1024 /*align slack*/ + 256 /*barriers*/;
};

constexpr int TS_THREADS = 320;   // TMA, MMA, 8 x operand split (two threads per frame) + epilogue

template <int BN>
__global__ void __launch_bounds__(TS_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
               const __grid_constant__ CUtensorMap map_wlo, const GemmTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  using S = TcSmem<BN>;
  constexpr int TC_STAGES = S::kStages;
  static_assert(BN == 64, "TMEM layout below: D @ 0..63, operand slots @ 64 / 128");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = blockIdx.x, ntile = blockIdx.y;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;

  // carve shared memory (tiles need 1024-byte alignment for the 128B swizzle)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  float* cscale = reinterpret_cast<float*>(tiles + TC_STAGES * S::kStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * S::kStage + S::kScale);
  uint64_t* full_bar = bars;                    // TMA bytes landed          (count 1 + tx)
  uint64_t* empty_bar = bars + TC_STAGES;       // MMAs that read the stage retired (count 1, tcgen05.commit)
  uint64_t* aready = bars + 2 * TC_STAGES;      // [2] X halves written to the TMEM operand slot (256 split threads)
  uint64_t* aempty = aready + 2;                // [2] ... consumed (tcgen05.commit)
  uint64_t* accum_bar = aempty + 2;             // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aready[s], 256);
      mbar_init(&aempty[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 64 accumulator columns + two (hi | lo) operand slots of 64 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kAslot = 64;
  pdl_launch_dependents();
  pdl_wait();

  const bool active = blk < nblk;
  BlkMeta m = {0, 0, 0, 0};
  if (active) m = p.meta[blk];
  const int row0 = blk * kBlkRows;
  const bool has_rows = active && row0 < m.hi;
  const int kblocks = p.ntaps * p.kbp;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  // Dropout2d: the scale vector of this tile's sequence goes to shared memory once (the split threads then read it as a
  // broadcast); read through __ldg inside the k loop it cost 70 us of the 150 us projection of a training step
  const bool cs_smem = p.colscale != nullptr && p.c_in * 4 <= S::kScale;
  if (has_rows && cs_smem && warp >= 2) {
    const float* src = p.colscale + (size_t)m.seq * p.colscale_ld;
    for (int i = threadIdx.x - 64; i < p.c_in; i += TS_THREADS - 64) cscale[i] = __ldg(src + i);
    asm volatile("bar.sync 1, 256;\n" ::: "memory");   // warps 2..9
  }
  const float* cs = p.colscale == nullptr ? nullptr : (cs_smem ? cscale : p.colscale + (size_t)m.seq * p.colscale_ld);

  if (has_rows) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        const int xrow = row0 + (p.x_unpadded ? m.in_delta : 0);
        int tap = 0, kc = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % TC_STAGES;
          const uint32_t ph = (kb / TC_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = tiles + s * S::kStage;
          mbar_arrive_expect_tx(&full_bar[s], S::kA + 2 * S::kB);
          const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
          tma_load_2d(st, &map_x, &full_bar[s], kc * TC_BK, xrow + sh);  // rows < 0 or past the end: zero fill
          tma_load_2d(st + S::kA, &map_whi, &full_bar[s], kb * TC_BK, ntile * BN);
          tma_load_2d(st + S::kA + S::kB, &map_wlo, &full_bar[s], kb * TC_BK, ntile * BN);
          if (++kc == p.kbp) { kc = 0; ++tap; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % TC_STAGES, ta = kb & 1;
        mbar_wait(&aready[ta], ((uint32_t)kb >> 1) & 1);   // implies full_bar[s]: the split threads waited for it
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tmem_base + kAslot + ta * 64, a_lo = a_hi + 32;
          const uint32_t b_hi = base + s * S::kStage + S::kA, b_lo = b_hi + S::kB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = umma_desc_sw128(b_hi + k * 32), dbl = umma_desc_sw128(b_lo + k * 32);
            umma_tf32_ts(tmem_base, a_lo + k * 8, dbh, idesc, (kb | k) != 0);
            umma_tf32_ts(tmem_base, a_hi + k * 8, dbl, idesc, 1u);
            umma_tf32_ts(tmem_base, a_hi + k * 8, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs retire
          umma_commit(&aempty[ta]);
          if (kb == kblocks - 1) umma_commit(accum_bar);    // accumulator complete
        }
        __syncwarp();
      }
    } else {
      {
        // ===================== operand split (warps 2..9): two threads per frame (= TMEM lane), 16 columns each =======
        // (warps w and w + 4 share the lane quadrant w % 4; with one thread per frame the four split warps -- hash of the
        // input mask, scale, split, tcgen05.st -- were the bottleneck of the projection: 2.4 us per k-block)
        const int r = (warp & 3) * 32 + lane, half = (warp - 2) >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t in_seed = p.in_drop_seed ^ dseed;
        int tap = 0, kc = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % TC_STAGES, ta = kb & 1;
          const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
          mbar_wait(&full_bar[s], (kb / TC_STAGES) & 1);
          mbar_wait(&aempty[ta], (((uint32_t)kb >> 1) & 1) ^ 1);
          tc_fence_after();
          split_half_row_to_tmem(reinterpret_cast<const float4*>(tiles + s * S::kStage), r, half, row0 + r + sh, m, kc, p,
                                 cs, in_seed, lane_base + kAslot + ta * 64);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&aready[ta]);
          if (++kc == p.kbp) { kc = 0; ++tap; }
        }
      }
      // ===================== epilogue (warps 2..9; TMEM lane quadrant = warp % 4, two warps share the columns) ======
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int row = row0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const uint32_t out_seed = p.drop_seed ^ dseed;
      const bool vec_ok = ((p.ldy & 3) == 0) && (p.R == nullptr || (p.ldr & 3) == 0) && (p.M == nullptr || (p.ldm & 3) == 0);
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 64) {
        float v[32];
        tmem_ld32(taddr + c0, v);  // warp-collective: every lane takes part, stores are predicated below
        if (row < m.hi) epilogue_cols(v, row, ntile * BN + c0, p, out_seed, vec_ok);
      }
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(256u));
}


// ------------------------------------------------------------------------------------------------
// Persistent variant for short contractions (ntaps * ceil(c_in/32) <= 6 k-blocks: every convolution of the
// residual layers, the FPN lateral, the heads and their input gradients).  One CTA per SM walks over the
// 128-frame tiles; the split weights stay resident in shared memory, the accumulator is double-buffered in
// TMEM, and four dedicated epilogue warps drain tile i while the producer / split / MMA warps are already
// working on tile i + 1.
//   warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2-5: operand split | warps 6-13: epilogue
constexpr int TP_THREADS = 448;   // TMA, MMA, 4 x operand split, 8 x epilogue (lane quadrant x 32-column half)
constexpr int TP_STAGES = 3;
constexpr int TP_MAX_KB = 6;
constexpr int TP_KA = TC_BM * TC_BK * 4;  // 16384: one A tile
constexpr int TP_KB = 64 * TC_BK * 4;     //  8192: one B tile (64 output columns)
constexpr int TP_EPI = 4 * 32 * 64 * 4;   // 32768: per-warp 32 x 64 fp32 epilogue staging
static inline int tp_smem_bytes(int kblocks) { return kblocks * 2 * TP_KB + TP_STAGES * 2 * TP_KA + TP_EPI + 1024 + 256; }

__global__ void __launch_bounds__(TP_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                       const __grid_constant__ CUtensorMap map_wlo, const GemmTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = blockIdx.y;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const int kblocks = p.ntaps * p.kbp;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* wt = smem_raw + (base - smem_u32(smem_raw));   // resident weights: kb -> [hi 8 KB | lo 8 KB]
  uint8_t* at = wt + kblocks * 2 * TP_KB;                 // A ring: stage -> [X / X_hi 16 KB | X_lo 16 KB]
  const uint32_t wt_addr = base, at_addr = base + kblocks * 2 * TP_KB;
  float4* epi = reinterpret_cast<float4*>(at + TP_STAGES * 2 * TP_KA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(at + TP_STAGES * 2 * TP_KA + TP_EPI);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* ready_bar = full_bar + TP_STAGES;
  uint64_t* empty_bar = ready_bar + TP_STAGES;
  uint64_t* tfull = empty_bar + TP_STAGES;   // [2] accumulator complete (tcgen05.commit)
  uint64_t* tempty = tfull + 2;              // [2] accumulator drained (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < TP_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  // everything above (barriers, TMEM) overlapped the tail of the previous kernel; its results are needed from here on
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, kblocks * 2 * TP_KB);
      for (int kb = 0; kb < kblocks; ++kb) {
        tma_load_2d(wt + kb * 2 * TP_KB, &map_whi, wfull, kb * TC_BK, ntile * 64);
        tma_load_2d(wt + kb * 2 * TP_KB + TP_KB, &map_wlo, wfull, kb * TC_BK, ntile * 64);
      }
      int it = 0;
      for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const BlkMeta m = p.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
        const int xrow = row0 + (p.x_unpadded ? m.in_delta : 0);
        int tap = 0, kc = 0;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % TP_STAGES;
          const uint32_t ph = (it / TP_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], TP_KA);
          const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
          tma_load_2d(at + s * 2 * TP_KA, &map_x, &full_bar[s], kc * TC_BK, xrow + sh);
          if (++kc == p.kbp) { kc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, 64);
    mbar_wait(wfull, 0);
    int it = 0, tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      if (blk * kBlkRows >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      mbar_wait(&tempty[a], tph ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * 64;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % TP_STAGES;
        const uint32_t ph = (it / TP_STAGES) & 1;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = at_addr + s * 2 * TP_KA, a_lo = a_hi + TP_KA;
          const uint32_t b_hi = wt_addr + kb * 2 * TP_KB, b_lo = b_hi + TP_KB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint32_t ko = k * 32;
            const uint64_t dah = umma_desc_sw128(a_hi + ko), dal = umma_desc_sw128(a_lo + ko);
            const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
            umma_tf32(tacc, dal, dbh, idesc, (kb | k) != 0);
            umma_tf32(tacc, dah, dbl, idesc, 1u);
            umma_tf32(tacc, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == kblocks - 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp < 6) {
    // ===================== operand split (warps 2..5) =====================
    const int ct = threadIdx.x - 64;
    const uint32_t in_seed = p.in_drop_seed ^ dseed;
    int it = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      int tap = 0, kc = 0;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % TP_STAGES;
        const uint32_t ph = (it / TP_STAGES) & 1;
        const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
        mbar_wait(&full_bar[s], ph);
        split_tile<128>(reinterpret_cast<float4*>(at + s * 2 * TP_KA), reinterpret_cast<float4*>(at + s * 2 * TP_KA + TP_KA),
                   ct, row0, sh, m, kc, p, in_seed);
        fence_proxy_async();
        mbar_arrive(&ready_bar[s]);
        if (++kc == p.kbp) { kc = 0; ++tap; }
      }
    }
  } else {
    // ===================== epilogue (warps 6..13; TMEM lane quadrant = warp % 4, 32-column half) ================
    const int q = warp & 3, half = (warp - 6) >> 2;
    const uint32_t out_seed = p.drop_seed ^ dseed;
    const bool vec_ok = ((p.ldy & 3) == 0) && (p.R == nullptr || (p.ldr & 3) == 0) && (p.M == nullptr || (p.ldm & 3) == 0);
    int tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      Epi32Pre pre;
      epi32_prefetch(pre, lane, row0 + q * 32, m.hi, ntile * 64 + half * 32, p, vec_ok);
      mbar_wait(&tfull[a], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * 64 + half * 32;
      float v[32];
      tmem_ld32(taddr, v);
      tc_fence_before();
      mbar_arrive(&tempty[a]);  // the MMA warp may overwrite this accumulator
      epilogue_block32_rt(v, epi + (warp - 6) * 256, lane, row0 + q * 32, m.hi, ntile * 64 + half * 32, p, out_seed,
                          vec_ok, pre);
      ++tcount;
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(128u));
}


// ------------------------------------------------------------------------------------------------
// Wide persistent variant for the dense per-frame contractions with many output columns (the MS-TCT Linear layers:
// qkv / proj / fc1 / fc2 and their input gradients).  One CTA per SM walks over (frame tile, BN-column tile) pairs;
// the 128 x 32 activation tile is split ONCE per BN = 128 / 256 output columns (vs. once per 64 in the streaming
// kernel, which made the split warps the bottleneck), X and the pre-split weights stream through one ring, the
// accumulator is double-buffered in TMEM (2 x BN columns) and four epilogue warps drain tile i while tile i + 1 is
// being multiplied.  Tile order: frame tile fastest, so that the CTAs running at the same time share a weight tile
// through L2.
template <int BN>
struct TwSmem {
  static constexpr int kStages = BN > 128 ? 2 : 3;
  static constexpr int kA = TC_BM * TC_BK * 4;  // 16384
  static constexpr int kB = BN * TC_BK * 4;
  static constexpr int kStage = 2 * kA + 2 * kB;
  static constexpr int kBytes = kStages * kStage + TP_EPI + 1024 + 256;
};

constexpr int TW_THREADS = 448;   // TMA, MMA, 4 x operand split, 8 x epilogue

template <int BN>
__global__ void __launch_bounds__(TW_THREADS, 1)
gemm_tc_wide_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                    const __grid_constant__ CUtensorMap map_wlo, const GemmTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  using S = TwSmem<BN>;
  constexpr int NST = S::kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const int kblocks = p.ntaps * p.kbp;
  const int nty = (p.N + BN - 1) / BN;
  const int ntiles = nblk * nty;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  float4* epi = reinterpret_cast<float4*>(tiles + NST * S::kStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + NST * S::kStage + TP_EPI);
  uint64_t* full_bar = bars;
  uint64_t* ready_bar = full_bar + NST;
  uint64_t* empty_bar = ready_bar + NST;
  uint64_t* tfull = empty_bar + NST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int blk = tile % nblk, ntile = tile / nblk;
        const BlkMeta m = p.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
        const int xrow = row0 + (p.x_unpadded ? m.in_delta : 0);
        int tap = 0, kc = 0;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % NST;
          const uint32_t ph = (it / NST) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = tiles + s * S::kStage;
          mbar_arrive_expect_tx(&full_bar[s], S::kA + 2 * S::kB);
          const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
          tma_load_2d(st, &map_x, &full_bar[s], kc * TC_BK, xrow + sh);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {  // weight maps have 64-row boxes; rows past the end are zero filled
            tma_load_2d(st + 2 * S::kA + j * TP_KB, &map_whi, &full_bar[s], kb * TC_BK, ntile * BN + j * 64);
            tma_load_2d(st + 2 * S::kA + S::kB + j * TP_KB, &map_wlo, &full_bar[s], kb * TC_BK, ntile * BN + j * 64);
          }
          if (++kc == p.kbp) { kc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int blk = tile % nblk;
      const BlkMeta m = p.meta[blk];
      if (blk * kBlkRows >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      mbar_wait(&tempty[a], tph ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * BN;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = base + s * S::kStage, a_lo = a_hi + S::kA;
          const uint32_t b_hi = a_hi + 2 * S::kA, b_lo = b_hi + S::kB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint32_t ko = k * 32;
            const uint64_t dah = umma_desc_sw128(a_hi + ko), dal = umma_desc_sw128(a_lo + ko);
            const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
            umma_tf32(tacc, dal, dbh, idesc, (kb | k) != 0);
            umma_tf32(tacc, dah, dbl, idesc, 1u);
            umma_tf32(tacc, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == kblocks - 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp < 6) {
    // ===================== operand split (warps 2..5) =====================
    const int ct = threadIdx.x - 64;
    const uint32_t in_seed = p.in_drop_seed ^ dseed;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int blk = tile % nblk;
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      int tap = 0, kc = 0;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
        mbar_wait(&full_bar[s], ph);
        split_tile<128>(reinterpret_cast<float4*>(tiles + s * S::kStage),
                        reinterpret_cast<float4*>(tiles + s * S::kStage + S::kA), ct, row0, sh, m, kc, p, in_seed);
        fence_proxy_async();
        mbar_arrive(&ready_bar[s]);
        if (++kc == p.kbp) { kc = 0; ++tap; }
      }
    }
  } else {
    // ===================== epilogue (warps 6..13: TMEM lane quadrant x half of the BN columns) =====================
    const int q = warp & 3, half = (warp - 6) >> 2;
    const uint32_t out_seed = p.drop_seed ^ dseed;
    const bool vec_ok = ((p.ldy & 3) == 0) && (p.R == nullptr || (p.ldr & 3) == 0) && (p.M == nullptr || (p.ldm & 3) == 0);
    float4* st = epi + (warp - 6) * 256;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int blk = tile % nblk, ntile = tile / nblk;
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      const int cbase = half * (BN / 2);
      Epi32Pre pre;
      epi32_prefetch(pre, lane, row0 + q * 32, m.hi, ntile * BN + cbase, p, vec_ok);
      mbar_wait(&tfull[a], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + cbase;
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (c0 + 32 >= BN / 2) {  // last read of this accumulator: hand it back before the stores
          tc_fence_before();
          mbar_arrive(&tempty[a]);
        }
        epilogue_block32_rt(v, st, lane, row0 + q * 32, m.hi, ntile * BN + cbase + c0, p, out_seed, vec_ok, pre);
        if (c0 + 32 < BN / 2) epi32_prefetch(pre, lane, row0 + q * 32, m.hi, ntile * BN + cbase + c0 + 32, p, vec_ok);
      }
      ++tcount;
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
}

template <int BN>
static int launch_wide(const CUtensorMap& mx, const CUtensorMap& mwhi, const CUtensorMap& mwlo, const GemmTcDev& p, int nb,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(gemm_tc_wide_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               TwSmem<BN>::kBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tc_wide: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  const long tiles = (long)nb * ((p.N + BN - 1) / BN);
  const int gx = (int)(tiles < num_sms() ? tiles : num_sms());
  launch_kernel(gemm_tc_wide_kernel<BN>, dim3(gx), dim3(TW_THREADS), TwSmem<BN>::kBytes, stream, true, mx, mwhi, mwlo, p);
  return check_launch("gemm_tc_wide_kernel");
}

// 0: not the wide kernel; else the column-tile width.  Cost model: waves of tiles x (MMA time ~ BN, plus a fixed
// split / pipeline-fill share).
static int wide_bn(const GemmTcDev& p, int nb) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("TCN_WIDE_BN");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0) return (forced == 128 || forced == 256) && p.N > 64 ? forced : 0;
  // short contractions stay on the persistent kernel (weights resident) unless they need three or more 64-column tiles:
  // the heads (N = 131, K = 64) split and load every frame tile once per 64 columns there -- one 256-column tile is faster
  if (p.N < 128 || (p.ntaps * p.kbp < 4 && p.N <= 128)) return 0;
  const int sms = num_sms();
  long best = -1;
  int best_bn = 0;
  for (int bn = 128; bn <= 256; bn += 128) {
    const long tiles = (long)nb * ((p.N + bn - 1) / bn);
    const long cost = ((tiles + sms - 1) / sms) * (bn + 48);
    if (best < 0 || cost < best) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

// ------------------------------------------------------------------------------------------------
// Fused residual layer forward (network.py:186-198), 64 channels, one kernel:
//     u = b1 + sum_k W1_k x[t + s_k];  h = relu(u);  v = b2 + W2 h;  y = x + dropout(v)
// Both GEMMs take their A operand from TENSOR MEMORY (tcgen05.mma with A in TMEM, tools/exp/exp_tmem_a.cu):
//   * GEMM 1 (dilated conv): TMA drops the raw 128 x 32 fp32 tiles of x (one per tap and 32-channel block) into a
//     4-deep shared-memory ring; each operand-split thread owns one frame (= one TMEM lane), un-swizzles its 128-byte
//     row, zeroes it when the tap leaves the sequence, and stores the TF32 hi / lo halves straight into a
//     double-buffered TMEM operand slot.  Shared memory holds only raw tiles (16 KB per stage instead of 32 KB for
//     hi + lo), is written once and read once, and the tensor pipe reads A from TMEM instead of shared memory.
//   * GEMM 2 (1x1 conv): the epilogue warps read u from TMEM, apply bias + ReLU, store the hi / lo halves of h back to
//     TMEM (and h once to HBM for the backward pass); h never touches shared memory.
// W1 / W2 (hi and lo) stay resident in shared memory; the MMA warp issues GEMM 1 of tile i + 1 before GEMM 2 of tile i.
//   TMEM (512 columns): u[2] @ 0 / 64, v[2] @ 128 / 192, h_hi @ 256, h_lo @ 320, x operand slots @ 384 / 448 (hi | lo)
//   warp 0: TMA | warp 1: MMA + TMEM owner | warps 2-5: operand split of x | warps 6-9: u -> h (TMEM + HBM) |
//   warps 10-13: v -> y | warps 14-15: dropout keep words.  The two epilogue groups work on different tiles at the same time.
constexpr int LFT_STAGES = 4;
constexpr int LFT_THREADS = 448;
constexpr int LFT_FWD_THREADS = 512;   // forward kernel: + two warps that precompute the dropout keep words
constexpr int LFT_EPI = 8 * 4096;   // eight epilogue warps x (32 x 32 fp32) staging
static inline int lft_smem_bytes() { return 8 * 2 * TP_KB + LFT_STAGES * TP_KA + LFT_EPI + 1024 + 256; }

__global__ void __launch_bounds__(LFT_FWD_THREADS, 1)
layer_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1hi,
                    const __grid_constant__ CUtensorMap map_w1lo, const __grid_constant__ CUtensorMap map_w2hi,
                    const __grid_constant__ CUtensorMap map_w2lo, const LayerTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.y.dyn ? p.y.dyn->nblk : p.y.nblk;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* wt = smem_raw + (base - smem_u32(smem_raw));   // slots 0..5: W1 (tap * 2 + kc), 6..7: W2 (kc); [hi | lo]
  uint8_t* at = wt + 8 * 2 * TP_KB;                       // raw x tiles, LFT_STAGES x 16 KB
  const uint32_t wt_addr = base;
  float4* epi = reinterpret_cast<float4*>(at + LFT_STAGES * TP_KA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(at + LFT_STAGES * TP_KA + LFT_EPI);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;                  // [4] TMA bytes landed
  uint64_t* empty_bar = full_bar + LFT_STAGES;    // [4] raw tile read by the 128 split threads
  uint64_t* aready = empty_bar + LFT_STAGES;      // [2] x operand slot written to TMEM (128 split threads)
  uint64_t* aempty = aready + 2;                  // [2] ... consumed by the tensor pipe (tcgen05.commit)
  uint64_t* ufull = aempty + 2;                   // [2] GEMM 1 accumulator complete
  uint64_t* uempty = ufull + 2;                   // [2] ... drained by the epilogue warps
  uint64_t* vfull = uempty + 2;                   // [2] GEMM 2 accumulator complete (also: h in TMEM is free again)
  uint64_t* vempty = vfull + 2;                   // [2]
  uint64_t* hready = vempty + 2;                  // h_hi / h_lo written to TMEM (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready + 1);
  uint32_t* kcount = tmem_slot + 1;               // keep-word threads done, 64 per tile (warps 14-15 -> v -> y warps)

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < LFT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&aready[s], 128);
      mbar_init(&aempty[s], 1);
      mbar_init(&ufull[s], 1);
      mbar_init(&uempty[s], 128);
      mbar_init(&vfull[s], 1);
      mbar_init(&vempty[s], 128);
    }
    mbar_init(hready, 128);
    *kcount = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dseed = p.y.dyn ? p.y.dyn->seed : 0u;
  constexpr uint32_t kU = 0, kV = 128, kHhi = 256, kHlo = 320, kA = 384;
  pdl_launch_dependents();
  // The split weights are constants of the step (written by split_weight_batched, behind a full stream dependency: see
  // launch_layer_fwd_tc): their 128 KB are requested BEFORE griddepcontrol.wait, i.e. while the predecessor's last tiles
  // still run -- tools/exp/launch_floor.cu measures 1.4 us for this load when it sits behind the wait.
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, 8 * 2 * TP_KB);
      for (int kb = 0; kb < 6; ++kb) {
        tma_load_2d(wt + kb * 2 * TP_KB, &map_w1hi, wfull, kb * TC_BK, 0);
        tma_load_2d(wt + kb * 2 * TP_KB + TP_KB, &map_w1lo, wfull, kb * TC_BK, 0);
      }
      for (int kc = 0; kc < 2; ++kc) {
        tma_load_2d(wt + (6 + kc) * 2 * TP_KB, &map_w2hi, wfull, kc * TC_BK, 0);
        tma_load_2d(wt + (6 + kc) * 2 * TP_KB + TP_KB, &map_w2lo, wfull, kc * TC_BK, 0);
      }
    }
    __syncwarp();
  }
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const BlkMeta m = p.y.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
        for (int j = 0; j < 6; ++j, ++it) {
          const int tap = j >> 1, kc = j & 1;
          const int s = it % LFT_STAGES;
          const uint32_t ph = (it / LFT_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], TP_KA);
          const int sh = tap == 0 ? p.y.shift[0] : (tap == 1 ? p.y.shift[1] : p.y.shift[2]);
          tma_load_2d(at + s * TP_KA, &map_x, &full_bar[s], kc * TC_BK, row0 + sh);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, 64);
    mbar_wait(wfull, 0);
    auto gemm2 = [&](int t) {  // v[t & 1] = h W2^T, h taken from TMEM
      const int a = t & 1;
      mbar_wait(hready, (uint32_t)(t & 1));
      mbar_wait(&vempty[a], (((uint32_t)t >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tacc = tmem_base + kV + a * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t b_hi = wt_addr + (6 + (k >> 2)) * 2 * TP_KB + (k & 3) * 32, b_lo = b_hi + TP_KB;
          const uint64_t dbh = umma_desc_sw128(b_hi), dbl = umma_desc_sw128(b_lo);
          umma_tf32_ts(tacc, tmem_base + kHlo + k * 8, dbh, idesc, k != 0);
          umma_tf32_ts(tacc, tmem_base + kHhi + k * 8, dbl, idesc, 1u);
          umma_tf32_ts(tacc, tmem_base + kHhi + k * 8, dbh, idesc, 1u);
        }
        umma_commit(&vfull[a]);
      }
      __syncwarp();
    };
    int it = 0, tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.y.meta[blk];
      if (blk * kBlkRows >= m.hi) continue;
      const int a = tcount & 1;
      mbar_wait(&uempty[a], (((uint32_t)tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + kU + a * 64;
      for (int j = 0; j < 6; ++j, ++it) {
        const int ta = it & 1;
        mbar_wait(&aready[ta], ((uint32_t)it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tmem_base + kA + ta * 64, a_lo = a_hi + 32;
          const uint32_t b_hi = wt_addr + j * 2 * TP_KB, b_lo = b_hi + TP_KB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = umma_desc_sw128(b_hi + k * 32), dbl = umma_desc_sw128(b_lo + k * 32);
            umma_tf32_ts(tacc, a_lo + k * 8, dbh, idesc, (j | k) != 0);
            umma_tf32_ts(tacc, a_hi + k * 8, dbl, idesc, 1u);
            umma_tf32_ts(tacc, a_hi + k * 8, dbh, idesc, 1u);
          }
          umma_commit(&aempty[ta]);
          if (j == 5) umma_commit(&ufull[a]);
        }
        __syncwarp();
      }
      if (tcount > 0) gemm2(tcount - 1);
      ++tcount;
    }
    if (tcount > 0) gemm2(tcount - 1);
  } else if (warp < 6) {
    // ===================== operand split of x (warps 2..5): one frame = one thread = one TMEM lane ================
    const int r = (warp & 3) * 32 + lane;                      // frame inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int it = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.y.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      for (int j = 0; j < 6; ++j, ++it) {
        const int tap = j >> 1;
        const int s = it % LFT_STAGES, ta = it & 1;
        const int sh = tap == 0 ? p.y.shift[0] : (tap == 1 ? p.y.shift[1] : p.y.shift[2]);
        mbar_wait(&full_bar[s], (it / LFT_STAGES) & 1);
        const float4* row = reinterpret_cast<const float4*>(at + s * TP_KA) + r * 8;
        const int src = row0 + r + sh;
        const bool keep = src >= m.lo && src < m.hi;           // select, not multiply: pad rows may hold anything
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = row[c ^ (r & 7)];                   // undo the 128B swizzle: logical 16-byte chunk c
          const float x0 = keep ? v.x : 0.f, x1 = keep ? v.y : 0.f, x2 = keep ? v.z : 0.f, x3 = keep ? v.w : 0.f;
          hi[4 * c + 0] = __uint_as_float(__float_as_uint(x0) & 0xffffe000u); lo[4 * c + 0] = x0 - hi[4 * c + 0];
          hi[4 * c + 1] = __uint_as_float(__float_as_uint(x1) & 0xffffe000u); lo[4 * c + 1] = x1 - hi[4 * c + 1];
          hi[4 * c + 2] = __uint_as_float(__float_as_uint(x2) & 0xffffe000u); lo[4 * c + 2] = x2 - hi[4 * c + 2];
          hi[4 * c + 3] = __uint_as_float(__float_as_uint(x3) & 0xffffe000u); lo[4 * c + 3] = x3 - hi[4 * c + 3];
        }
        mbar_arrive(&empty_bar[s]);                            // the raw tile is in registers: TMA may refill the slot
        mbar_wait(&aempty[ta], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        tmem_st32(lane_base + kA + ta * 64, hi);
        tmem_st32(lane_base + kA + ta * 64 + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&aready[ta]);
      }
    }
  } else if (warp < 10) {
    // ===================== u -> h (warps 6..9; TMEM lane quadrant = warp % 4) =====================
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* st = epi + (warp - 6) * 256;
    int tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.y.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int a = tcount & 1;
      mbar_wait(&ufull[a], ((uint32_t)tcount >> 1) & 1);
      tc_fence_after();
      if (tcount > 0) mbar_wait(&vfull[(tcount - 1) & 1], ((uint32_t)(tcount - 1) >> 1) & 1);  // h of the previous tile consumed
      tc_fence_after();
      // bias + ReLU, hi / lo halves of h into TMEM (32 columns at a time)
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float u[32], lo[32];
        uint32_t pos = 0u;   // bit j: h[row, c0 + j] > 0 (the ReLU mask of the backward pass)
        tmem_ld32(lane_base + kU + a * 64 + c0, u);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.h.bias + c0 + j));
          const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hv = fmaxf(u[j + e] + bb[e], 0.f);
            pos |= (hv > 0.f ? 1u : 0u) << (j + e);
            u[j + e] = __uint_as_float(__float_as_uint(hv) & 0xffffe000u);
            lo[j + e] = hv - u[j + e];
          }
        }
        tmem_st32(lane_base + kHhi + c0, u);
        tmem_st32(lane_base + kHlo + c0, lo);
        const int row = row0 + q * 32 + lane;
        if (p.masks != nullptr && row < m.hi) p.masks[(size_t)row * 4 + (c0 >> 5)] = pos;
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(hready);   // GEMM 2 can start; the copy of h to HBM below overlaps it
      if (p.h.Y != nullptr) {
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float u[32];
          tmem_ld32(lane_base + kU + a * 64 + c0, u);
          if (c0 == 32) {
            tc_fence_before();
            mbar_arrive(&uempty[a]);
          }
          float4 none[8];
          epilogue_block32<true, false, false>(u, st, lane, row0 + q * 32, m.hi, c0, p.h, 0u, none);
        }
      } else {
        mbar_arrive(&uempty[a]);
      }
      ++tcount;
    }
  } else if (warp >= 14) {
    // ===================== dropout keep words (warps 14..15) =====================
    // bit c of word c / 32 of a frame = column c is kept.  Two frames per thread and tile; the words go to the mask array
    // (the fused backward kernel reads them as its dropout mask) and v -> y reads them back instead of hashing: the hash
    // costs ~25 instructions per four columns, which made the v -> y warps (later the split warps) the bottleneck.
    // Nothing here depends on the pipeline, so these warps run ahead of it.
    if (p.masks != nullptr) {
      const uint32_t keep_seed = p.y.drop_seed ^ dseed;
      const int hrow = threadIdx.x - 14 * 32;   // 0..63
      for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const BlkMeta m = p.y.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int row = row0 + hrow + half * 64;
          uint32_t w[2] = {0xffffffffu, 0xffffffffu};
          if (p.y.drop_thresh != 0u) {
#pragma unroll
            for (int wi = 0; wi < 2; ++wi) {
              uint32_t bits = 0u;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const uint32_t h = drop_hash4(keep_seed, p.y.drop_stream, (uint32_t)row, (uint32_t)(wi * 8 + g));
#pragma unroll
                for (int e = 0; e < 4; ++e) bits |= (drop_rotl(h, e) >= p.y.drop_thresh ? 1u : 0u) << (g * 4 + e);
              }
              w[wi] = bits;
            }
          }
          if (row < m.hi) *reinterpret_cast<uint2*>(p.masks + (size_t)row * 4 + 2) = make_uint2(w[0], w[1]);
        }
        __threadfence();   // device scope: the v -> y warps read the words back with ld.global.cg (L2), not through L1
        atomicAdd(kcount, 1u);
      }
    }
  } else {
    // ===================== v -> y: bias, dropout, + x (warps 10..13) =====================
    const int q = warp & 3;
    const uint32_t out_seed = p.y.drop_seed ^ dseed;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* st = epi + (warp - 6) * 256;
    int tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.y.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int a = tcount & 1;
      mbar_wait(&vfull[a], ((uint32_t)tcount >> 1) & 1);
      tc_fence_after();
      uint2 kw = make_uint2(0u, 0u);   // keep words of this lane's frame, written by warps 14-15 (normally long ago)
      if (p.masks != nullptr) {
        while (*reinterpret_cast<volatile uint32_t*>(kcount) < 64u * (uint32_t)(tcount + 1)) {}
        __threadfence_block();
        const int row = min(row0 + q * 32 + lane, m.hi - 1);
        kw = __ldcg(reinterpret_cast<const uint2*>(p.masks + (size_t)row * 4 + 2));
      }
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        // (fetching the residual before the wait, as the persistent kernels do, measured slower here: the same rows
        // are in flight through the TMA ring at that time)
        float4 rr[8];
        epi32_load_residual(rr, lane, row0 + q * 32, m.hi, c0, p.y);
        float v[32];
        tmem_ld32(lane_base + kV + a * 64 + c0, v);
        if (c0 == 32) {
          tc_fence_before();
          mbar_arrive(&vempty[a]);
        }
        if (p.masks == nullptr) {
          epilogue_block32<false, true, true>(v, st, lane, row0 + q * 32, m.hi, c0, p.y, out_seed, rr);
        } else {
          epilogue_block32_keepbits(v, st, lane, row0 + q * 32, m.hi, c0, p.y, c0 == 0 ? kw.x : kw.y, p.y.drop_scale, rr);
        }
      }
      ++tcount;
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
}

// pdl: programmatic dependent launch.  The kernel reads its weights before griddepcontrol.wait, so `pdl` may only be set when
// the weights were final before the PREDECESSOR started -- the executor launches the first layer of a pass without it
// (a full dependency: everything enqueued earlier, the weight split included, has completed), the C ABI never sets it.
int launch_layer_fwd_tc(const CUtensorMap& mx, const CUtensorMap& w1hi, const CUtensorMap& w1lo, const CUtensorMap& w2hi,
                        const CUtensorMap& w2lo, const LayerTcDev& p, int cap_nblk, cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(layer_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lft_smem_bytes());
    if (e != cudaSuccess) {
      set_error("layer_fwd_tc: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  const int nb = cap_nblk > 0 ? cap_nblk : p.y.nblk;
  int gx = num_sms();
  if (gx > nb) gx = nb;
  launch_kernel(layer_fwd_tc_kernel, dim3(gx), dim3(LFT_FWD_THREADS), lft_smem_bytes(), stream, pdl, mx, w1hi, w1lo, w2hi, w2lo,
                p);
  return check_launch("layer_fwd_tc_kernel");
}

// ------------------------------------------------------------------------------------------------
// Fused residual layer INPUT GRADIENT (network.py:186-198 backward), 64 channels, one kernel:
//     gv = keep * gy / (1 - p);  gu = (gv W2) * [h > 0];  gx[t] = gy[t] + sum_k W1_k^T gu[t - s_k]
// For each tap k the kernel recomputes gu on the 128 frames t0 - s_k .. (1x1 GEMM, A = gv from a TMEM operand slot
// fed by the split threads), masks it with the ReLU bit words the forward kernel saved, stores the hi / lo halves back
// to TMEM and feeds them to the tap's GEMM against W1_k^T -- gu never returns from HBM.  The centre tap (s = 0)
// also writes gu once for the weight-gradient kernel.  Dropout / ReLU masks come as 32-bit words per (frame, half)
// written by layer_fwd_tc_kernel, so the backward pass hashes nothing.
//   TMEM (512 columns): gu accumulators U[2] @ 0 / 64 -- the gu warps write the hi half of the masked gu back IN PLACE, so
//   U[b] is also the hi operand of the tap GEMM --, gx accumulators ACC[2] @ 128 / 192, gv operand slots @ 256 / 320
//   (hi | lo, one 32-channel block each), lo halves of gu @ 384 / 448.  Both gu operand buffers alternate per tap: the gu
//   warps convert tap n + 1 while the tensor pipe runs the tap GEMM of tap n.  The MMA warp issues one flat, software-
//   pipelined sequence over (tile, tap) steps: U(0), then { U(n + 1), G2(n) } -- tcgen05.mma executes in issue order, so
//   U(n + 2) overwriting buffer n & 1 after G2(n) read it needs no barrier.
//   warp 0: TMA | warp 1: MMA + TMEM owner | warps 2-5: gy -> gv split | warps 6-9: U -> gu (TMEM + HBM) |
//   warps 10-13: ACC + gy -> gx
__global__ void __launch_bounds__(LFT_THREADS, 1)
layer_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_w2thi,
                    const __grid_constant__ CUtensorMap map_w2tlo, const __grid_constant__ CUtensorMap map_w1thi,
                    const __grid_constant__ CUtensorMap map_w1tlo, const LayerBwdTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.gx.dyn ? p.gx.dyn->nblk : p.gx.nblk;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* wt = smem_raw + (base - smem_u32(smem_raw));   // slots 0..1: W2^T (kc), 2..7: W1^T (tap * 2 + kc); [hi | lo]
  uint8_t* at = wt + 8 * 2 * TP_KB;                       // raw gy tiles, LFT_STAGES x 16 KB
  const uint32_t wt_addr = base;
  float4* epi = reinterpret_cast<float4*>(at + LFT_STAGES * TP_KA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(at + LFT_STAGES * TP_KA + LFT_EPI);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;                  // [4] TMA bytes landed
  uint64_t* empty_bar = full_bar + LFT_STAGES;    // [4] raw tile read by the 128 split threads
  uint64_t* a1ready = empty_bar + LFT_STAGES;     // [2] gv operand slot written (128 split threads)
  uint64_t* a1empty = a1ready + 2;                // [2] ... consumed (tcgen05.commit)
  uint64_t* ufull = a1empty + 2;                  // [2] gu accumulator complete
  uint64_t* a2ready = ufull + 2;                  // [2] gu operand (hi in place of U, lo) written to TMEM (128 threads)
  uint64_t* accfull = a2ready + 2;                // [2] gx accumulator complete
  uint64_t* accempty = accfull + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < LFT_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 128);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a1ready[s], 128);
      mbar_init(&a1empty[s], 1);
      mbar_init(&ufull[s], 1);
      mbar_init(&a2ready[s], 128);
      mbar_init(&accfull[s], 1);
      mbar_init(&accempty[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kU = 0, kAcc = 128, kA1 = 256, kA2lo = 384;
  pdl_launch_dependents();
  if (warp == 0) {   // step-constant weights: requested before griddepcontrol.wait (see layer_fwd_tc_kernel)
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, 8 * 2 * TP_KB);
      for (int kc = 0; kc < 2; ++kc) {
        tma_load_2d(wt + kc * 2 * TP_KB, &map_w2thi, wfull, kc * TC_BK, 0);
        tma_load_2d(wt + kc * 2 * TP_KB + TP_KB, &map_w2tlo, wfull, kc * TC_BK, 0);
      }
      for (int kb = 0; kb < 6; ++kb) {
        tma_load_2d(wt + (2 + kb) * 2 * TP_KB, &map_w1thi, wfull, kb * TC_BK, 0);
        tma_load_2d(wt + (2 + kb) * 2 * TP_KB + TP_KB, &map_w1tlo, wfull, kb * TC_BK, 0);
      }
    }
    __syncwarp();
  }
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const BlkMeta m = p.gx.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
        for (int j = 0; j < 6; ++j, ++it) {
          const int tap = j >> 1, kc = j & 1;
          const int s = it % LFT_STAGES;
          const uint32_t ph = (it / LFT_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], TP_KA);
          const int sh = tap == 0 ? p.gx.shift[0] : (tap == 1 ? p.gx.shift[1] : p.gx.shift[2]);
          tma_load_2d(at + s * TP_KA, &map_gy, &full_bar[s], kc * TC_BK, row0 + sh);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, 64);
    mbar_wait(wfull, 0);
    int ntiles = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.gx.meta[blk];
      if (blk * kBlkRows < m.hi) ++ntiles;
    }
    const int nsteps = ntiles * 3;   // (tile, tap) steps of this CTA
    int it = 0;
    auto gemm1 = [&](int n) {  // U[n & 1] = gv W2 for step n: two 32-channel operand slots from the split warps
      const uint32_t tacc = tmem_base + kU + (n & 1) * 64;
      for (int kc = 0; kc < 2; ++kc, ++it) {
        const int ta = it & 1;
        mbar_wait(&a1ready[ta], ((uint32_t)it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = tmem_base + kA1 + ta * 64, a_lo = a_hi + 32;
          const uint32_t b_hi = wt_addr + kc * 2 * TP_KB, b_lo = b_hi + TP_KB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t dbh = umma_desc_sw128(b_hi + k * 32), dbl = umma_desc_sw128(b_lo + k * 32);
            umma_tf32_ts(tacc, a_lo + k * 8, dbh, idesc, (kc | k) != 0);
            umma_tf32_ts(tacc, a_hi + k * 8, dbl, idesc, 1u);
            umma_tf32_ts(tacc, a_hi + k * 8, dbh, idesc, 1u);
          }
          umma_commit(&a1empty[ta]);
          if (kc == 1) umma_commit(&ufull[n & 1]);
        }
        __syncwarp();
      }
    };
    auto gemm2 = [&](int n, int tap, int tile) {  // ACC[tile & 1] (+)= gu_tap W1_tap^T, gu (hi in place of U, lo) from TMEM
      const int ub = n & 1, ab = tile & 1;
      mbar_wait(&a2ready[ub], ((uint32_t)n >> 1) & 1);
      if (tap == 0) mbar_wait(&accempty[ab], (((uint32_t)tile >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tacc = tmem_base + kAcc + ab * 64;
        const uint32_t a_hi = tmem_base + kU + ub * 64, a_lo = tmem_base + kA2lo + ub * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t b_hi = wt_addr + (2 + tap * 2 + (k >> 2)) * 2 * TP_KB + (k & 3) * 32, b_lo = b_hi + TP_KB;
          const uint64_t dbh = umma_desc_sw128(b_hi), dbl = umma_desc_sw128(b_lo);
          umma_tf32_ts(tacc, a_lo + k * 8, dbh, idesc, (tap | k) != 0);
          umma_tf32_ts(tacc, a_hi + k * 8, dbl, idesc, 1u);
          umma_tf32_ts(tacc, a_hi + k * 8, dbh, idesc, 1u);
        }
        if (tap == 2) umma_commit(&accfull[ab]);
      }
      __syncwarp();
    };
    if (nsteps > 0) gemm1(0);
    for (int n = 0, tap = 0, tile = 0; n < nsteps; ++n) {
      if (n + 1 < nsteps) gemm1(n + 1);
      gemm2(n, tap, tile);
      if (++tap == 3) { tap = 0; ++tile; }
    }
  } else if (warp < 6) {
    // ===================== gy -> gv (warps 2..5): one frame = one thread = one TMEM lane =====================
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int it = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.gx.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      for (int j = 0; j < 6; ++j, ++it) {
        const int tap = j >> 1, kc = j & 1;
        const int s = it % LFT_STAGES, ta = it & 1;
        const int sh = tap == 0 ? p.gx.shift[0] : (tap == 1 ? p.gx.shift[1] : p.gx.shift[2]);
        const int src = row0 + r + sh;
        const bool inside = src >= m.lo && src < m.hi;
        uint32_t keep = 0u;
        if (inside) keep = p.use_drop ? __ldg(p.masks + (size_t)src * 4 + 2 + kc) : 0xffffffffu;
        mbar_wait(&full_bar[s], (it / LFT_STAGES) & 1);
        const float4* row = reinterpret_cast<const float4*>(at + s * TP_KA) + r * 8;
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = row[c ^ (r & 7)];
          const float x0 = (keep >> (4 * c + 0)) & 1u ? v.x * p.drop_scale : 0.f;
          const float x1 = (keep >> (4 * c + 1)) & 1u ? v.y * p.drop_scale : 0.f;
          const float x2 = (keep >> (4 * c + 2)) & 1u ? v.z * p.drop_scale : 0.f;
          const float x3 = (keep >> (4 * c + 3)) & 1u ? v.w * p.drop_scale : 0.f;
          hi[4 * c + 0] = __uint_as_float(__float_as_uint(x0) & 0xffffe000u); lo[4 * c + 0] = x0 - hi[4 * c + 0];
          hi[4 * c + 1] = __uint_as_float(__float_as_uint(x1) & 0xffffe000u); lo[4 * c + 1] = x1 - hi[4 * c + 1];
          hi[4 * c + 2] = __uint_as_float(__float_as_uint(x2) & 0xffffe000u); lo[4 * c + 2] = x2 - hi[4 * c + 2];
          hi[4 * c + 3] = __uint_as_float(__float_as_uint(x3) & 0xffffe000u); lo[4 * c + 3] = x3 - hi[4 * c + 3];
        }
        mbar_arrive(&empty_bar[s]);
        mbar_wait(&a1empty[ta], (((uint32_t)it >> 1) & 1) ^ 1);
        tc_fence_after();
        tmem_st32(lane_base + kA1 + ta * 64, hi);
        tmem_st32(lane_base + kA1 + ta * 64 + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a1ready[ta]);
      }
    }
  } else if (warp < 10) {
    // ===================== U -> gu (warps 6..9): ReLU mask, hi / lo halves back to TMEM, centre tap to HBM ==========
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* st = epi + (warp - 6) * 256;
    int uc = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.gx.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      for (int tap = 0; tap < 3; ++tap, ++uc) {
        const int ub = uc & 1;
        const int sh = tap == 0 ? p.gx.shift[0] : (tap == 1 ? p.gx.shift[1] : p.gx.shift[2]);
        const int src = row0 + q * 32 + lane + sh;
        const bool inside = src >= m.lo && src < m.hi;
        uint32_t pos0 = 0u, pos1 = 0u;
        if (inside) {
          const uint2 w = __ldg(reinterpret_cast<const uint2*>(p.masks + (size_t)src * 4));
          pos0 = w.x; pos1 = w.y;
        }
        // U[ub] complete; the tap GEMM that read this operand buffer two steps ago was issued before it: in-order pipe
        mbar_wait(&ufull[ub], ((uint32_t)uc >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          const uint32_t pos = c0 == 0 ? pos0 : pos1;
          float u[32], hi[32], lo[32];
          tmem_ld32(lane_base + kU + ub * 64 + c0, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            u[j] = (pos >> j) & 1u ? u[j] : 0.f;
            hi[j] = __uint_as_float(__float_as_uint(u[j]) & 0xffffe000u);
            lo[j] = u[j] - hi[j];
          }
          tmem_st32(lane_base + kU + ub * 64 + c0, hi);      // in place: U[ub] becomes the hi operand
          tmem_st32(lane_base + kA2lo + ub * 64 + c0, lo);
          if (sh == 0 && p.gu.Y != nullptr) {   // the centre tap is gu of this tile: keep it for the weight gradients
            float4 none[8];
            epilogue_block32<false, false, false, false>(u, st, lane, row0 + q * 32, m.hi, c0, p.gu, 0u, none);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a2ready[ub]);
      }
    }
  } else {
    // ===================== ACC + gy -> gx (warps 10..13) =====================
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float4* st = epi + (warp - 6) * 256;
    int tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.gx.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int ab = tcount & 1;
      float4 rr0[8], rr1[8];   // gy rows of the tile (residual), fetched while the MMAs run
      epi32_load_residual(rr0, lane, row0 + q * 32, m.hi, 0, p.gx);
      epi32_load_residual(rr1, lane, row0 + q * 32, m.hi, 32, p.gx);
      mbar_wait(&accfull[ab], ((uint32_t)tcount >> 1) & 1);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(lane_base + kAcc + ab * 64, v);
        epilogue_block32<false, true, false, false>(v, st, lane, row0 + q * 32, m.hi, 0, p.gx, 0u, rr0);
        tmem_ld32(lane_base + kAcc + ab * 64 + 32, v);
        tc_fence_before();
        mbar_arrive(&accempty[ab]);
        epilogue_block32<false, true, false, false>(v, st, lane, row0 + q * 32, m.hi, 32, p.gx, 0u, rr1);
      }
      ++tcount;
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u));
}

int launch_layer_bwd_tc(const CUtensorMap& mgy, const CUtensorMap& w2thi, const CUtensorMap& w2tlo,
                        const CUtensorMap& w1thi, const CUtensorMap& w1tlo, const LayerBwdTcDev& p, int cap_nblk,
                        cudaStream_t stream, bool pdl) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(layer_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lft_smem_bytes());
    if (e != cudaSuccess) {
      set_error("layer_bwd_tc: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  const int nb = cap_nblk > 0 ? cap_nblk : p.gx.nblk;
  int gx = num_sms();
  if (gx > nb) gx = nb;
  launch_kernel(layer_bwd_tc_kernel, dim3(gx), dim3(LFT_THREADS), lft_smem_bytes(), stream, pdl, mgy, w2thi, w2tlo, w1thi,
                w1tlo, p);
  return check_launch("layer_bwd_tc_kernel");
}

// ------------------------------------------------------------------------------------------------
// Slab variant of the persistent kernel for the k = 3 dilated convolutions with a small dilation
// (max shift - min shift <= 64 frames, 64 input channels).  tcgen05 shared-memory descriptors may start at ANY row
// of a TMA-swizzled tile (the 128B swizzle is a function of the absolute shared-memory address; verified on
// B200, tools/exp/exp_desc.cu), so the three taps are read as row-offset views of ONE staged time slab
// x[t0 + s_min .. t0 + 128 + s_max): every frame is loaded and split once per tile instead of once per tap.
//   ring: 2 slots (one per 32-channel block), each [hi 192 x 128 B | lo 192 x 128 B]
constexpr int SL_ROWS = 192;
constexpr int SL_HALF = SL_ROWS * 128;      // 24576
constexpr int SL_SLOT = 2 * SL_HALF;        // 49152
static inline int sl_smem_bytes() { return 6 * 2 * TP_KB + 2 * SL_SLOT + TP_EPI + 1024 + 256; }

__global__ void __launch_bounds__(TP_THREADS, 1)
gemm_tc_slab_kernel(const __grid_constant__ CUtensorMap map_x32, const __grid_constant__ CUtensorMap map_whi,
                    const __grid_constant__ CUtensorMap map_wlo, const GemmTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = blockIdx.y;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const int smin = min(p.shift[0], min(p.shift[1], p.shift[2]));
  const int smax = max(p.shift[0], max(p.shift[1], p.shift[2]));
  const int nrows = ((kBlkRows + (smax - smin)) + 31) & ~31;  // slab rows staged per tile (<= 192)

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* wt = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sl = wt + 6 * 2 * TP_KB;
  const uint32_t wt_addr = base, sl_addr = base + 6 * 2 * TP_KB;
  float4* epi = reinterpret_cast<float4*>(sl + 2 * SL_SLOT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sl + 2 * SL_SLOT + TP_EPI);
  uint64_t* wfull = bars;
  uint64_t* full_bar = bars + 1;    // [2]
  uint64_t* ready_bar = bars + 3;   // [2]
  uint64_t* empty_bar = bars + 5;   // [2]
  uint64_t* tfull = bars + 7;       // [2]
  uint64_t* tempty = bars + 9;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, 6 * 2 * TP_KB);
      for (int kb = 0; kb < 6; ++kb) {
        tma_load_2d(wt + kb * 2 * TP_KB, &map_whi, wfull, kb * TC_BK, ntile * 64);
        tma_load_2d(wt + kb * 2 * TP_KB + TP_KB, &map_wlo, wfull, kb * TC_BK, ntile * 64);
      }
      int it = 0;
      for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const BlkMeta m = p.meta[blk];
        const int row0 = blk * kBlkRows;
        if (row0 >= m.hi) continue;
        for (int kc = 0; kc < 2; ++kc, ++it) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], nrows * 128);
          for (int i = 0; i < nrows / 32; ++i)
            tma_load_2d(sl + s * SL_SLOT + i * 4096, &map_x32, &full_bar[s], kc * TC_BK, row0 + smin + 32 * i);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, 64);
    mbar_wait(wfull, 0);
    int it = 0, tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      if (blk * kBlkRows >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      mbar_wait(&tempty[a], tph ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + a * 64;
      for (int kc = 0; kc < 2; ++kc, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
            const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
            const uint32_t a_hi = sl_addr + s * SL_SLOT + (uint32_t)(sh - smin) * 128, a_lo = a_hi + SL_HALF;
            const uint32_t b_hi = wt_addr + (tap * 2 + kc) * 2 * TP_KB, b_lo = b_hi + TP_KB;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint32_t ko = k * 32;
              const uint64_t dah = umma_desc_sw128(a_hi + ko), dal = umma_desc_sw128(a_lo + ko);
              const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
              umma_tf32(tacc, dal, dbh, idesc, (kc | tap | k) != 0);
              umma_tf32(tacc, dah, dbl, idesc, 1u);
              umma_tf32(tacc, dah, dbh, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (kc == 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp < 6) {
    // ===================== operand split of the slab (warps 2..5) =====================
    const int ct = threadIdx.x - 64;
    const int nchunks = nrows * 8;
    int it = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      for (int kc = 0; kc < 2; ++kc, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&full_bar[s], ph);
        float4* xa = reinterpret_cast<float4*>(sl + s * SL_SLOT);
        float4* xl = reinterpret_cast<float4*>(sl + s * SL_SLOT + SL_HALF);
        for (int c0 = ct; c0 < nchunks; c0 += 4 * 128) {  // nchunks is a multiple of 256
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = (c0 + i * 128 < nchunks) ? xa[c0 + i * 128] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = c0 + i * 128;
            if (c < nchunks) {
              const int src = row0 + smin + (c >> 3);  // a slab row is valid iff its source frame is inside the sequence
              const bool keep = src >= m.lo && src < m.hi;   // select, not multiply: pad rows may hold anything
              float4 h, l;
              const float x0 = keep ? v[i].x : 0.f, x1 = keep ? v[i].y : 0.f, x2 = keep ? v[i].z : 0.f, x3 = keep ? v[i].w : 0.f;
              h.x = __uint_as_float(__float_as_uint(x0) & 0xffffe000u); l.x = x0 - h.x;
              h.y = __uint_as_float(__float_as_uint(x1) & 0xffffe000u); l.y = x1 - h.y;
              h.z = __uint_as_float(__float_as_uint(x2) & 0xffffe000u); l.z = x2 - h.z;
              h.w = __uint_as_float(__float_as_uint(x3) & 0xffffe000u); l.w = x3 - h.w;
              xa[c] = h;
              xl[c] = l;
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&ready_bar[s]);
      }
    }
  } else {
    // ===================== epilogue (warps 6..13: TMEM lane quadrant x 32-column half) =====================
    const int q = warp & 3, half = (warp - 6) >> 2;
    const uint32_t out_seed = p.drop_seed ^ dseed;
    const bool vec_ok = ((p.ldy & 3) == 0) && (p.R == nullptr || (p.ldr & 3) == 0) && (p.M == nullptr || (p.ldm & 3) == 0);
    int tcount = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
      const BlkMeta m = p.meta[blk];
      const int row0 = blk * kBlkRows;
      if (row0 >= m.hi) continue;
      const int a = tcount & 1;
      const uint32_t tph = (tcount >> 1) & 1;
      Epi32Pre pre;
      epi32_prefetch(pre, lane, row0 + q * 32, m.hi, ntile * 64 + half * 32, p, vec_ok);
      mbar_wait(&tfull[a], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * 64 + half * 32;
      float v[32];
      tmem_ld32(taddr, v);
      tc_fence_before();
      mbar_arrive(&tempty[a]);
      epilogue_block32_rt(v, epi + (warp - 6) * 256, lane, row0 + q * 32, m.hi, ntile * 64 + half * 32, p, out_seed,
                          vec_ok, pre);
      ++tcount;
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(128u));
}

static bool slab_eligible(const GemmTcDev& p) {
  if (p.ntaps != 3 || p.kbp != 2 || p.x_unpadded || p.colscale != nullptr || p.in_drop_thresh != 0u) return false;
  const int smin = p.shift[0] < p.shift[1] ? (p.shift[0] < p.shift[2] ? p.shift[0] : p.shift[2])
                                           : (p.shift[1] < p.shift[2] ? p.shift[1] : p.shift[2]);
  const int smax = p.shift[0] > p.shift[1] ? (p.shift[0] > p.shift[2] ? p.shift[0] : p.shift[2])
                                           : (p.shift[1] > p.shift[2] ? p.shift[1] : p.shift[2]);
  return (smax - smin) <= 64;
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      cudaGetLastError();
  }
  return fn;
}

// fp32 row-major (rows x cols, leading dimension ld floats) -> 2D map with a (32 x box_rows) box, 128B swizzle
// atom32: 128-byte swizzle with 32-byte atoms (the only layout tcgen05 accepts for MN-major TF32 operands)
int make_tensor_map_2d(CUtensorMap* map, const float* ptr, long rows, long cols, long ld, int box_rows, bool atom32) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TCN_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr %p rows %ld cols %ld ld %ld", (int)r, (const void*)ptr, rows,
              cols, ld);
    return TCN_ERR_CUDA;
  }
  return TCN_OK;
}

__device__ __forceinline__ void split_one(const float* __restrict__ w, float* __restrict__ whi, float* __restrict__ wlo,
                                          int n_out, int c_in, int ntaps, int transpose, int kcols, long i) {
  const int r = (int)(i / kcols), k = (int)(i - (long)r * kcols);
  const int kp = kcols / ntaps;  // padded K per tap
  const int tap = k / kp, kk = k - tap * kp;
  float v = 0.f;
  if (!transpose) {
    if (r < n_out && kk < c_in) v = w[((long)r * c_in + kk) * ntaps + tap];
  } else {
    if (r < c_in && kk < n_out) v = w[((long)kk * c_in + r) * ntaps + tap];
  }
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  whi[i] = h;
  wlo[i] = v - h;
}

__global__ void split_weight_kernel(const float* __restrict__ w, float* __restrict__ whi, float* __restrict__ wlo,
                                    int n_out, int c_in, int ntaps, int transpose, int rows_pad, int kcols) {
  const long total = (long)rows_pad * kcols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    split_one(w, whi, wlo, n_out, c_in, ntaps, transpose, kcols, i);
}

__global__ void split_weight_batched_kernel(const SplitJob* __restrict__ jobs, int njobs,
                                            const float* __restrict__ params, float* __restrict__ whi,
                                            float* __restrict__ wlo, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const SplitJob jb = jobs[lo];
    split_one(params + jb.src_off, whi + jb.dst_off, wlo + jb.dst_off, jb.n_out, jb.c_in, jb.ntaps, jb.transpose,
              jb.kcols, i - jb.first);
  }
}

int launch_split_weight(const float* w, float* whi, float* wlo, int n_out, int c_in, int ntaps, int transpose,
                        cudaStream_t stream) {
  const int rows_pad = (int)tc_weight_rows(n_out, c_in, transpose);
  const int kcols = (int)tc_weight_cols(n_out, c_in, ntaps, transpose);
  long b = ((long)rows_pad * kcols + 255) / 256;
  if (b > 1024) b = 1024;
  split_weight_kernel<<<(int)b, 256, 0, stream>>>(w, whi, wlo, n_out, c_in, ntaps, transpose, rows_pad, kcols);
  return check_launch("split_weight_kernel");
}

int launch_split_batched(const SplitJob* jobs_dev, int njobs, const float* params, float* whi, float* wlo, long total,
                         cudaStream_t stream) {
  long b = (total + 255) / 256;
  if (b > 2048) b = 2048;
  split_weight_batched_kernel<<<(int)b, 256, 0, stream>>>(jobs_dev, njobs, params, whi, wlo, total);
  return check_launch("split_weight_batched_kernel");
}

bool gemm_tc_wants_slab(const GemmTcDev& p) {
  static int on = -1;
  if (on < 0) on = getenv("TCN_NO_SLAB") == nullptr ? 1 : 0;
  return on == 1 && slab_eligible(p);
}

int launch_gemm_tc(const CUtensorMap& mx, const CUtensorMap& mwhi, const CUtensorMap& mwlo, const GemmTcDev& p,
                   int cap_nblk, cudaStream_t stream, const CUtensorMap* mx32) {
  const int nb = cap_nblk > 0 ? cap_nblk : p.nblk;
  const int kblocks = p.ntaps * p.kbp;
  const int nty = (p.N + 63) / 64;
  if (mx32 != nullptr && gemm_tc_wants_slab(p)) {
    static bool slab_set = false;
    if (!slab_set) {
      const cudaError_t e =
          cudaFuncSetAttribute(gemm_tc_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sl_smem_bytes());
      if (e != cudaSuccess) {
        set_error("gemm_tc_slab: smem attribute: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return TCN_ERR_CUDA;
      }
      slab_set = true;
    }
    int gx = num_sms() / nty;
    if (gx < 1) gx = 1;
    if (gx > nb) gx = nb;
    launch_kernel(gemm_tc_slab_kernel, dim3(gx, nty), dim3(TP_THREADS), sl_smem_bytes(), stream, true, *mx32, mwhi, mwlo, p);
    return check_launch("gemm_tc_slab_kernel");
  }
  if (const int bn = wide_bn(p, nb)) {
    return bn == 256 ? launch_wide<256>(mx, mwhi, mwlo, p, nb, stream) : launch_wide<128>(mx, mwhi, mwlo, p, nb, stream);
  }
  if (kblocks <= TP_MAX_KB) {  // persistent, weights resident in shared memory
    static int max_set = 0;
    const int smem = tp_smem_bytes(kblocks);
    if (smem > max_set) {
      const cudaError_t e =
          cudaFuncSetAttribute(gemm_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tp_smem_bytes(TP_MAX_KB));
      if (e != cudaSuccess) {
        set_error("gemm_tc_persist: smem attribute: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return TCN_ERR_CUDA;
      }
      max_set = tp_smem_bytes(TP_MAX_KB);
    }
    int gx = num_sms() / nty;
    if (gx < 1) gx = 1;
    if (gx > nb) gx = nb;
    launch_kernel(gemm_tc_persist_kernel, dim3(gx, nty), dim3(TP_THREADS), smem, stream, true, mx, mwhi, mwlo, p);
    return check_launch("gemm_tc_persist_kernel");
  }
  dim3 grid(nb, nty);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64>::kBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tc: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_kernel(gemm_tc_kernel<64>, grid, dim3(TS_THREADS), TcSmem<64>::kBytes, stream, true, mx, mwhi, mwlo, p);
  return check_launch("gemm_tc_kernel");
}

}  // namespace tcn

using namespace tcn;

extern "C" int tcn_gemm_tc_supported(int c_in, int n_out) {
  // x rows must be 16-byte multiples for the TMA map; everything else is padded / masked
  return (c_in > 0 && c_in % 4 == 0 && n_out > 0) ? 1 : 0;
}

extern "C" long long tcn_split_weight_floats(int n_out, int c_in, int ntaps, int transpose) {
  return (long long)tc_weight_rows(n_out, c_in, transpose) * tc_weight_cols(n_out, c_in, ntaps, transpose);
}

extern "C" int tcn_split_weight(const float* w, int n_out, int c_in, int ntaps, int transpose, float* w_hi, float* w_lo,
                                tcn_stream_t stream) {
  TCN_REQUIRE(w && w_hi && w_lo && n_out > 0 && c_in > 0 && ntaps >= 1 && ntaps <= 3, "tcn_split_weight: bad arguments");
  return launch_split_weight(w, w_hi, w_lo, n_out, c_in, ntaps, transpose, (cudaStream_t)stream);
}

extern "C" int tcn_gemm_tc(const tcn_gemm_tc_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->x && a->w_hi && a->w_lo && a->y && a->meta, "tcn_gemm_tc: null pointer");
  if (!tcn_gemm_tc_supported(a->c_in, a->n_out)) {
    set_error("tcn_gemm_tc: needs c_in %% 4 == 0 (got c_in=%d n_out=%d); use tcn_tapgemm", a->c_in, a->n_out);
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->nblk > 0 && a->x_rows > 0 && a->ldx >= a->c_in && a->ldx % 4 == 0 && a->ldy >= a->n_out,
              "tcn_gemm_tc: bad shape");
  TCN_REQUIRE(a->ntaps >= 1 && a->ntaps <= 3, "tcn_gemm_tc: ntaps must be 1..3");
  TCN_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w_hi) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->w_lo) & 15) == 0,
              "tcn_gemm_tc: x, w_hi, w_lo must be 16-byte aligned");
  TCN_REQUIRE(a->in_drop_p >= 0.f && a->in_drop_p < 1.f && a->drop_p >= 0.f && a->drop_p < 1.f,
              "tcn_gemm_tc: dropout probabilities must be in [0, 1)");
  const long wrows = tc_weight_rows(a->n_out, a->c_in, 0);  // caller passes the (already oriented) logical shape
  const long wcols = (long)a->ntaps * tc_kbp(a->c_in) * TC_BK;
  CUtensorMap mx, mh, ml;
  TCN_CHECK(make_tensor_map_2d(&mx, a->x, a->x_rows, a->c_in, a->ldx, TC_BM));
  TCN_CHECK(make_tensor_map_2d(&mh, a->w_hi, wrows, wcols, wcols, 64));
  TCN_CHECK(make_tensor_map_2d(&ml, a->w_lo, wrows, wcols, wcols, 64));
  GemmTcDev p;
  memset(&p, 0, sizeof(p));
  p.Y = a->y; p.ldy = a->ldy; p.N = a->n_out; p.bias = a->bias;
  p.R = a->residual; p.ldr = a->ldr; p.M = a->relu_mask; p.ldm = a->ldm; p.relu = a->relu;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  p.x_unpadded = a->x_unpadded; p.ntaps = a->ntaps;
  for (int i = 0; i < 3; ++i) p.shift[i] = a->shift[i];
  p.kbp = tc_kbp(a->c_in); p.c_in = a->c_in;
  p.colscale = a->colscale; p.colscale_ld = a->colscale_ld;
  p.in_drop_thresh = a->in_drop_p > 0.f ? drop_thresh(a->in_drop_p) : 0u;
  p.in_drop_scale = (a->in_drop_p > 0.f && a->in_drop_rescale) ? 1.f / (1.f - a->in_drop_p) : 1.f;
  p.in_drop_seed = a->drop_seed; p.in_drop_stream = a->drop_stream;
  p.drop_thresh = a->drop_p > 0.f ? drop_thresh(a->drop_p) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = a->drop_seed; p.drop_stream = a->drop_stream;
  CUtensorMap mx32;
  const bool slab = gemm_tc_wants_slab(p);
  if (slab) TCN_CHECK(make_tensor_map_2d(&mx32, a->x, a->x_rows, a->c_in, a->ldx, 32));
  return launch_gemm_tc(mx, mh, ml, p, 0, (cudaStream_t)stream, slab ? &mx32 : nullptr);
}

extern "C" int tcn_layer_fwd_tc(const tcn_layer_fwd_tc_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->x && a->y && a->w1_hi && a->w1_lo && a->w2_hi && a->w2_lo && a->b1 && a->b2 && a->meta,
              "tcn_layer_fwd_tc: null pointer");
  if (a->channels != 64) {
    set_error("tcn_layer_fwd_tc: the fused kernel is built for 64 channels (got %d); use tcn_gemm_tc", a->channels);
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->nblk > 0 && a->x_rows > 0, "tcn_layer_fwd_tc: empty problem");
  TCN_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "tcn_layer_fwd_tc: drop_p must be in [0, 1)");
  CUtensorMap mx, w1h, w1l, w2h, w2l;
  TCN_CHECK(make_tensor_map_2d(&mx, a->x, a->x_rows, 64, 64, TC_BM));
  TCN_CHECK(make_tensor_map_2d(&w1h, a->w1_hi, 64, 192, 192, 64));
  TCN_CHECK(make_tensor_map_2d(&w1l, a->w1_lo, 64, 192, 192, 64));
  TCN_CHECK(make_tensor_map_2d(&w2h, a->w2_hi, 64, 64, 64, 64));
  TCN_CHECK(make_tensor_map_2d(&w2l, a->w2_lo, 64, 64, 64, 64));
  LayerTcDev p;
  memset(&p, 0, sizeof(p));
  p.h.Y = a->h; p.h.ldy = 64; p.h.N = 64; p.h.bias = a->b1; p.h.relu = 1; p.h.drop_scale = 1.f; p.h.in_drop_scale = 1.f;
  p.y.Y = a->y; p.y.ldy = 64; p.y.N = 64; p.y.bias = a->b2; p.y.R = a->x; p.y.ldr = 64;
  p.y.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.y.nblk = a->nblk; p.y.ntaps = 3; p.y.kbp = 2; p.y.c_in = 64;
  for (int i = 0; i < 3; ++i) p.y.shift[i] = a->shift[i];
  p.y.in_drop_scale = 1.f;
  p.y.drop_thresh = a->drop_p > 0.f ? drop_thresh(a->drop_p) : 0u;
  p.y.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.y.drop_seed = a->drop_seed; p.y.drop_stream = a->drop_stream;
  p.masks = a->masks;
  return launch_layer_fwd_tc(mx, w1h, w1l, w2h, w2l, p, 0, (cudaStream_t)stream, false);
}

extern "C" int tcn_layer_bwd_tc(const tcn_layer_bwd_tc_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->gy && a->gx && a->masks && a->w2t_hi && a->w2t_lo && a->w1t_hi && a->w1t_lo && a->meta,
              "tcn_layer_bwd_tc: null pointer");
  if (a->channels != 64) {
    set_error("tcn_layer_bwd_tc: the fused kernel is built for 64 channels (got %d); use tcn_gemm_tc", a->channels);
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->nblk > 0 && a->g_rows > 0, "tcn_layer_bwd_tc: empty problem");
  TCN_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "tcn_layer_bwd_tc: drop_p must be in [0, 1)");
  TCN_REQUIRE(a->shift[0] == 0 || a->shift[1] == 0 || a->shift[2] == 0, "tcn_layer_bwd_tc: one tap must have shift 0");
  CUtensorMap mg, w2h, w2l, w1h, w1l;
  TCN_CHECK(make_tensor_map_2d(&mg, a->gy, a->g_rows, 64, 64, TC_BM));
  TCN_CHECK(make_tensor_map_2d(&w2h, a->w2t_hi, 64, 64, 64, 64));
  TCN_CHECK(make_tensor_map_2d(&w2l, a->w2t_lo, 64, 64, 64, 64));
  TCN_CHECK(make_tensor_map_2d(&w1h, a->w1t_hi, 64, 192, 192, 64));
  TCN_CHECK(make_tensor_map_2d(&w1l, a->w1t_lo, 64, 192, 192, 64));
  LayerBwdTcDev p;
  memset(&p, 0, sizeof(p));
  p.gu.Y = a->gu; p.gu.ldy = 64; p.gu.N = 64; p.gu.drop_scale = 1.f; p.gu.in_drop_scale = 1.f;
  p.gx.Y = a->gx; p.gx.ldy = 64; p.gx.N = 64; p.gx.R = a->gy; p.gx.ldr = 64;
  p.gx.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.gx.nblk = a->nblk; p.gx.ntaps = 3; p.gx.kbp = 2; p.gx.c_in = 64;
  for (int i = 0; i < 3; ++i) p.gx.shift[i] = -a->shift[i];
  p.gx.drop_scale = 1.f; p.gx.in_drop_scale = 1.f;
  p.masks = a->masks;
  p.use_drop = a->drop_p > 0.f ? 1 : 0;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  return launch_layer_bwd_tc(mg, w2h, w2l, w1h, w1l, p, 0, (cudaStream_t)stream, false);
}
