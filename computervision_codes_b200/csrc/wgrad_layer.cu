// All weight / bias gradients of a residual layer (network.py:186-198 backward) from ONE pass over the
// frames, deterministic:
//     gW1[:, :, k] = sum_t gu[t] x[t + s_k]^T (k = 0..2)     gb1 = sum_t gu[t]
//     gW2          = sum_t gv[t] h[t]^T,  gv = keep * gy / (1 - p)     gb2 = sum_t gv[t]
// One CTA owns a contiguous range of frames of one layer and computes all four 64 x 64 products for it, so every
// operand row is staged and split once (the pair kernel of round 1 ran three CTAs per range, each re-loading and
// re-splitting gu / gy: its shared-memory pipe was 94 % busy).
//
// Tensor-core mapping (contraction over frames => both operands "MN-major": 16-frame x 32-channel TMA boxes, 128-byte
// rows, 32-byte swizzle atoms):
//   stage = 12 boxes [x(t+s0) | x(t+s1) | x(t+s2) | h | gu | gy], two 32-channel boxes each, + the same 12 boxes of lo halves
//   D_a (TMEM 128 x 64)  += [x_s0 | x_s1]^T gu          lanes 0-63: gW1[:, :, 0]^T, lanes 64-127: gW1[:, :, 1]^T
//   D_b (TMEM 128 x 128) += [x_s2 | h]^T [gu | gv]      lanes 0-63 x cols 0-63: gW1[:, :, 2]^T; lanes 64-127 x cols 64-127: gW2^T
// 3xTF32: tcgen05.mma kind::tf32 ignores the low 13 mantissa bits of its operands, so the RAW tile is the hi operand
// as it lies; the split threads only write lo = x - trunc(x) (and gv / zeroed boundary rows in place).
// Reduction over CTAs: every CTA stores its partial (4 x 64 x 64 + 2 x 64 floats) to its own slab; a second kernel adds
// the slabs in a fixed order -- no atomics, bit-identical from run to run.  Several layers share one launch
// (blockIdx.y) so that a launch fills the GPU even when a layer only has work for a few dozen CTAs.
//   warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2-9: operand split, bias sums, epilogue
#include <cstdlib>
#include <cstring>

#include "wgrad_layer.cuh"

namespace tcn {

constexpr int WL_THREADS = 320;
constexpr int WL_SPLIT = 256;
constexpr int WL_ATOM = WL_RC * 128;          // 2048 B: 16 frames x 32 fp32 channels
constexpr int WL_NATOM = 12;
constexpr int WL_HALF = WL_NATOM * WL_ATOM;   // 24576 B raw (= hi) region; the lo region follows
constexpr int WL_STAGE = 2 * WL_HALF;         // 49152 B
constexpr int WL_STAGES = 4;
constexpr int WL_SMEM = WL_STAGES * WL_STAGE + 1024 + 256;
constexpr int WL_TMEM_COLS = 256;             // D_a @ 0 (64 columns), D_b @ 64 (128 columns)

__device__ __forceinline__ uint64_t wl_desc(uint32_t smem_addr) {   // MN-major, SWIZZLE_128B_BASE32B, LBO = one box
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((WL_ATOM >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t wl_idesc(int M, int N) {   // kind::tf32, fp32 accumulate, A and B MN-major
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void wgrad_layer_body(const WgLayerDev* __restrict__ d, const WgLayersLaunch& q, const int split) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = q.dyn ? q.dyn->nblk : q.nblk;
  // 16-frame slots of this CTA: an even share of the nblk * 8 slots (slots past the end of a sequence are skipped)
  const long nslots = (long)nblk * 8;
  const int s_lo = (int)(nslots * split / q.splits), s_hi = (int)(nslots * (split + 1) / q.splits);

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + WL_STAGES * WL_STAGE);
  uint64_t* full_bar = bars;
  uint64_t* ready_bar = bars + WL_STAGES;
  uint64_t* empty_bar = bars + 2 * WL_STAGES;
  uint64_t* accum_bar = bars + 3 * WL_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WL_STAGES + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WL_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], WL_SPLIT / 32);   // one arrival per split warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)WL_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  // every role walks the same sequence of valid slots
  int total = 0;
  for (int slot = s_lo; slot < s_hi; ++slot) {
    const int blk = slot >> 3;
    if (blk * kBlkRows + (slot & 7) * WL_RC < q.meta[blk].hi) ++total;
  }
  float* part = d->part + (size_t)split * WL_PART_FLOATS;

  if (total == 0) {   // nothing to add: the slab must still read as zero
    for (int i = threadIdx.x; i < WL_PART_FLOATS / 4; i += blockDim.x)
      reinterpret_cast<float4*>(part)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int it = 0;
      for (int slot = s_lo; slot < s_hi; ++slot) {
        const int blk = slot >> 3;
        const int r0 = blk * kBlkRows + (slot & 7) * WL_RC;
        if (r0 >= q.meta[blk].hi) continue;
        const int s = it % WL_STAGES;
        mbar_wait(&empty_bar[s], ((it / WL_STAGES) & 1) ^ 1);
        uint8_t* st = tiles + s * WL_STAGE;
        if (q.debug & 4) {
          mbar_arrive_expect_tx(&full_bar[s], 4 * WL_ATOM);
          tma_load_2d(st + 0 * WL_ATOM, &d->mx, &full_bar[s], 0, r0);
          tma_load_2d(st + 1 * WL_ATOM, &d->mx, &full_bar[s], 32, r0);
          tma_load_2d(st + 8 * WL_ATOM, &d->mgu, &full_bar[s], 0, r0);
          tma_load_2d(st + 9 * WL_ATOM, &d->mgu, &full_bar[s], 32, r0);
          ++it;
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[s], WL_HALF);
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
          const int sh = tap == 0 ? d->shift[0] : (tap == 1 ? d->shift[1] : d->shift[2]);
          tma_load_2d(st + (tap * 2 + 0) * WL_ATOM, &d->mx, &full_bar[s], 0, r0 + sh);
          tma_load_2d(st + (tap * 2 + 1) * WL_ATOM, &d->mx, &full_bar[s], 32, r0 + sh);
        }
        tma_load_2d(st + 6 * WL_ATOM, &d->mh, &full_bar[s], 0, r0);
        tma_load_2d(st + 7 * WL_ATOM, &d->mh, &full_bar[s], 32, r0);
        tma_load_2d(st + 8 * WL_ATOM, &d->mgu, &full_bar[s], 0, r0);
        tma_load_2d(st + 9 * WL_ATOM, &d->mgu, &full_bar[s], 32, r0);
        tma_load_2d(st + 10 * WL_ATOM, &d->mgy, &full_bar[s], 0, r0);
        tma_load_2d(st + 11 * WL_ATOM, &d->mgy, &full_bar[s], 32, r0);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_a = wl_idesc(128, 64), idesc_b = wl_idesc(128, 128);
    const uint32_t t_da = tmem_base, t_db = tmem_base + 64;
    for (int it = 0; it < total; ++it) {
      const int s = it % WL_STAGES;
      mbar_wait(&ready_bar[s], (it / WL_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t hi0 = base + s * WL_STAGE, lo0 = hi0 + WL_HALF;
#pragma unroll
        for (int k = 0; k < WL_RC / 8; ++k) {
          if (q.debug & 1) break;
          const uint32_t ko = k * 1024;   // 8 frames
          const uint64_t a1h = wl_desc(hi0 + ko), a1l = wl_desc(lo0 + ko);
          const uint64_t a2h = wl_desc(hi0 + 4 * WL_ATOM + ko), a2l = wl_desc(lo0 + 4 * WL_ATOM + ko);
          const uint64_t bh = wl_desc(hi0 + 8 * WL_ATOM + ko), bl = wl_desc(lo0 + 8 * WL_ATOM + ko);
          const uint32_t acc = (it | k) != 0;
          umma_tf32(t_da, a1l, bh, idesc_a, acc);
          umma_tf32(t_da, a1h, bl, idesc_a, 1u);
          umma_tf32(t_da, a1h, bh, idesc_a, 1u);
          umma_tf32(t_db, a2l, bh, idesc_b, acc);
          umma_tf32(t_db, a2h, bl, idesc_b, 1u);
          umma_tf32(t_db, a2h, bh, idesc_b, 1u);
        }
        umma_commit(&empty_bar[s]);
        if (it == total - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== operand split + bias sums (warps 2..9) =====================
    const int ct = threadIdx.x - 64;   // 0..255
    const int ph = ct >> 7;            // channel half: boxes 2 i + ph
    const int idx = ct & 127;          // 16-byte chunk inside a 16 x 128 B box
    const int r = idx >> 3;            // frame inside the slot (the same in every slot)
    const int lc = ((((idx & 7) >> 1) ^ (r & 3)) << 1) | (idx & 1);   // logical 16-byte column chunk (swizzle undone)
    const int col = ph * 32 + lc * 4;
    const uint32_t g_seed = d->drop_seed ^ (q.dyn ? q.dyn->seed : 0u);
    float4 bs_gu = make_float4(0.f, 0.f, 0.f, 0.f), bs_gv = bs_gu;
    int it = 0;
    int cur_blk = -1;
    BlkMeta m = {0, 0, 0, 0}, m_next = {0, 0, 0, 0};
    const bool use_bits = d->use_drop && d->masks != nullptr;
    const int total_rows = nblk * kBlkRows;
    // the keep word of the NEXT slot and the block table entry of the NEXT block are fetched one iteration ahead: a global
    // load issued and consumed inside one iteration put its full latency on every slot (ncu: 22 % of all stall samples
    // sat on the first use of the keep word)
    uint32_t keep_pref = 0u;
    int pref_slot = -1;
    for (int slot = s_lo; slot < s_hi; ++slot) {
      const int blk = slot >> 3;
      if (blk != cur_blk) {
        m = (blk == cur_blk + 1 && cur_blk >= 0) ? m_next : q.meta[blk];
        cur_blk = blk;
        if (blk + 1 < nblk) m_next = q.meta[blk + 1];
      }
      const int r0 = blk * kBlkRows + (slot & 7) * WL_RC;
      if (r0 >= m.hi) continue;
      const int row = r0 + r;
      const bool row_ok = row < m.hi;
      // dropout keep bits of (row, 32 channels of this half)
      uint32_t keepw = 0xffffffffu;
      if (use_bits) {
        if (pref_slot == slot) keepw = keep_pref;
        else if (row_ok) keepw = __ldg(d->masks + (size_t)row * 4 + 2 + ph);
        const int nrow = row + WL_RC;   // the same frame position in slot + 1 (blocks are contiguous)
        if (slot + 1 < s_hi && nrow < total_rows) {
          keep_pref = __ldg(d->masks + (size_t)nrow * 4 + 2 + ph);
          pref_slot = slot + 1;
        }
      }
      const int s = it % WL_STAGES;
      // ONE lane polls, the warp follows through __syncwarp: ncu counts 32 shared-memory wavefronts for a try_wait executed
      // by a full warp (9.6 M of the kernel's 15.7 M shared-load wavefronts were such replays, on a 90 % busy pipe)
      if (lane == 0) mbar_wait(&full_bar[s], (it / WL_STAGES) & 1);
      __syncwarp();
      if (q.debug & 2) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[s]);
        ++it;
        continue;
      }
      float4* raw = reinterpret_cast<float4*>(tiles + s * WL_STAGE) + ph * 128 + idx;
      float4* lo = raw + WL_HALF / 16;
      float4 v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = raw[i * 256];
      bool dirty[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) {   // x at the three taps: rows outside the sequence contribute zero
        const int src = row + (i == 0 ? d->shift[0] : (i == 1 ? d->shift[1] : d->shift[2]));
        dirty[i] = !(row_ok && src >= m.lo && src < m.hi);
      }
      dirty[3] = dirty[4] = !row_ok;
      dirty[5] = true;                // gv is always rewritten
#pragma unroll
      for (int i = 0; i < 5; ++i)
        if (dirty[i]) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!row_ok) {
        v[5] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (d->use_drop) {
        if (d->masks != nullptr) {
          const uint32_t b = keepw >> (lc * 4);
          v[5].x = (b & 1u) ? v[5].x * d->drop_scale : 0.f;
          v[5].y = (b & 2u) ? v[5].y * d->drop_scale : 0.f;
          v[5].z = (b & 4u) ? v[5].z * d->drop_scale : 0.f;
          v[5].w = (b & 8u) ? v[5].w * d->drop_scale : 0.f;
        } else {
          float f[4];
          drop_factor4(g_seed, d->drop_stream, d->drop_thresh, d->drop_scale, row, col, f);
          v[5].x *= f[0]; v[5].y *= f[1]; v[5].z *= f[2]; v[5].w *= f[3];
        }
      }
      bs_gu.x += v[4].x; bs_gu.y += v[4].y; bs_gu.z += v[4].z; bs_gu.w += v[4].w;
      bs_gv.x += v[5].x; bs_gv.y += v[5].y; bs_gv.z += v[5].z; bs_gv.w += v[5].w;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u); l.x = v[i].x - h.x;
        h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u); l.y = v[i].y - h.y;
        h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u); l.z = v[i].z - h.z;
        h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u); l.w = v[i].w - h.w;
        if (dirty[i]) raw[i * 256] = v[i];   // otherwise the raw tile stays: the tensor core truncates it to hi itself
        lo[i * 256] = l;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready_bar[s]);
      ++it;
    }
    // ===================== epilogue: the CTA's partial to its slab =====================
    if (lane == 0) mbar_wait(accum_bar, 0);
    __syncwarp();
    tc_fence_after();
    // bias sums: fixed-order reduction over the 16 frame positions through shared memory (the stages are free now)
    float* red = reinterpret_cast<float*>(tiles);   // [2][16][64]
    *reinterpret_cast<float4*>(red + (0 * 16 + r) * 64 + col) = bs_gu;
    *reinterpret_cast<float4*>(red + (1 * 16 + r) * 64 + col) = bs_gv;
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (ct < 128) {
      const int g = ct >> 6, n = ct & 63;
      float sacc = 0.f;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) sacc += red[(g * 16 + rr) * 64 + n];
      part[4 * 64 * 64 + g * 64 + n] = sacc;
    }
    const int qd = warp & 3;            // TMEM lane quadrant
    const int half = (warp - 2) >> 2;   // 32-column half
    const int mrow = qd * 32 + lane;    // accumulator lane
    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
    float v[32];
    {   // D_a: lanes 0-63 = tap 0, lanes 64-127 = tap 1
      tmem_ld32(taddr + half * 32, v);
      float4* dst = reinterpret_cast<float4*>(part + ((size_t)(mrow >> 6) * 64 + (mrow & 63)) * 64 + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    {   // D_b: lanes 0-63 x columns 0-63 = tap 2; lanes 64-127 x columns 64-127 = W2
      const int which = 2 + (qd >> 1);
      tmem_ld32(taddr + 64 + (qd >> 1) * 64 + half * 32, v);
      float4* dst = reinterpret_cast<float4*>(part + ((size_t)which * 64 + (mrow & 63)) * 64 + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)WL_TMEM_COLS));
}

// blockIdx.x = frame split, blockIdx.y = layer of this launch (descriptors in device memory)
__global__ void __launch_bounds__(WL_THREADS, 1)
wgrad_layers_kernel(const WgLayerDev* __restrict__ descs, const WgLayersLaunch q) {
  wgrad_layer_body(descs + blockIdx.y, q, blockIdx.x);
}

// one layer, descriptor passed by value (C ABI entry point)
__global__ void __launch_bounds__(WL_THREADS, 1)
wgrad_layer_kernel(const __grid_constant__ WgLayerDev d, const WgLayersLaunch q) {
  wgrad_layer_body(&d, q, blockIdx.x);
}

// dW / db += sum over the slabs in slab order (fixed summation tree for a given split count => bit-identical results).
// GROUPS = 8 (one layer, up to 148 slabs, few blocks): a block is 32 float4 columns x 8 slab groups, group g adds slabs
// g, g + 8, ... and the group sums are added in group order.  GROUPS = 1 (all layers of a model in one launch, plenty of
// blocks): a block reads 4 KB of every slab in turn, eight loads in flight.
template <int GROUPS>
__device__ __forceinline__ void wl_reduce_body(const WgLayerOut& o, int splits) {
  __shared__ float4 red[GROUPS > 1 ? GROUPS : 1][32];
  pdl_wait();   // launched with programmatic serialization: the slabs are final only when the producer grid has completed
  constexpr int COLS = 256 / GROUPS;
  const int lane = threadIdx.x % COLS, g = threadIdx.x / COLS;
  const int e4 = blockIdx.x * COLS + lane;
  constexpr size_t kStride = WL_PART_FLOATS / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e4 < WL_PART_FLOATS / 4) {
    const float4* p = reinterpret_cast<const float4*>(o.part) + e4;
    int s = g;
    for (; s + 7 * GROUPS < splits; s += 8 * GROUPS) {   // eight loads in flight, additions in slab order
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(s + u * GROUPS) * kStride);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; s < splits; s += GROUPS) {
      const float4 a = __ldg(p + (size_t)s * kStride);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  if (GROUPS > 1) {
    red[g][lane] = acc;
    __syncthreads();
    if (g != 0) return;
#pragma unroll
    for (int k = 1; k < GROUPS; ++k) {
      const float4 a = red[k][lane];
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  if (e4 >= WL_PART_FLOATS / 4) return;
  const float av[4] = {acc.x, acc.y, acc.z, acc.w};
  const int e0 = e4 * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = e0 + j;
    if (e < 4 * 64 * 64) {
      const int which = e >> 12, c = (e >> 6) & 63, n = e & 63;
      if (which < 3) o.dw1[((size_t)n * 64 + c) * 3 + which] += av[j];
      else o.dw2[(size_t)n * 64 + c] += av[j];
    } else {
      const int b = e - 4 * 64 * 64, n = b & 63;
      if (b < 64) { if (o.db1) o.db1[n] += av[j]; }
      else { if (o.db2) o.db2[n] += av[j]; }
    }
  }
}
__global__ void __launch_bounds__(256)
wgrad_layers_reduce_kernel(const WgLayerOut* __restrict__ outs, int splits) {
  wl_reduce_body<1>(outs[blockIdx.y], splits);
}
__global__ void __launch_bounds__(256)
wgrad_layer_reduce_kernel(const WgLayerOut o, int splits) {
  wl_reduce_body<8>(o, splits);
}

static int wl_set_attr() {
  static bool done = false;
  if (done) return TCN_OK;
  cudaError_t e = cudaFuncSetAttribute(wgrad_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WL_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WL_SMEM);
  if (e != cudaSuccess) {
    set_error("wgrad_layer: smem attribute: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  done = true;
  return TCN_OK;
}

void wgrad_layers_plan(int cap_nblk, int* layers_per_launch, int* splits) {
  // aim at ~80 slots of 16 frames per CTA at capacity (measured on the bench step, 233-block capacity: 2 / 4 / 6 / 8 layers
  // per launch -> 1.871 / 1.841 / 1.794 / 1.798 ms per step: fewer, longer CTAs amortise the prologue and the slab store)
  const int sms = num_sms();
  const long slots = (long)cap_nblk * 8;
  long lg = (2L * sms * 80 + slots) / (2 * slots);   // round(sms * 80 / slots)
  if (lg < 1) lg = 1;
  if (lg > 8) lg = 8;
  int sp = sms / (int)lg;
  if (sp > cap_nblk * 8) sp = cap_nblk * 8;
  if (sp < 1) sp = 1;
  *layers_per_launch = (int)lg;
  *splits = sp;
}

int launch_wgrad_layers(const WgLayerDev* descs_dev, int nlayers, const WgLayersLaunch& q, cudaStream_t stream) {
  TCN_CHECK(wl_set_attr());
  launch_kernel(wgrad_layers_kernel, dim3(q.splits, nlayers, 1), dim3(WL_THREADS), WL_SMEM, stream, true, descs_dev, q);
  return check_launch("wgrad_layers_kernel");
}

int launch_wgrad_layers_reduce(const WgLayerOut* outs_dev, int nlayers, int splits, cudaStream_t stream) {
  const int nb = (WL_PART_FLOATS / 4 + 255) / 256;
  launch_kernel(wgrad_layers_reduce_kernel, dim3(nb, nlayers, 1), dim3(256), 0, stream, true, outs_dev, splits);
  return check_launch("wgrad_layers_reduce_kernel");
}

int make_wgrad_layer_maps(WgLayerDev* d, const float* x, const float* h, const float* gu, const float* gy, long rows) {
  TCN_CHECK(make_tensor_map_2d(&d->mx, x, rows, 64, 64, WL_RC, true));
  TCN_CHECK(make_tensor_map_2d(&d->mh, h, rows, 64, 64, WL_RC, true));
  TCN_CHECK(make_tensor_map_2d(&d->mgu, gu, rows, 64, 64, WL_RC, true));
  TCN_CHECK(make_tensor_map_2d(&d->mgy, gy, rows, 64, 64, WL_RC, true));
  return TCN_OK;
}

}  // namespace tcn

using namespace tcn;

extern "C" long long tcn_wgrad_layer_workspace_bytes(int nblk) {
  (void)nblk;
  return (long long)num_sms() * WL_PART_FLOATS * 4;   // one slab per CTA, at most one CTA per SM
}

extern "C" int tcn_wgrad_layer(const tcn_wgrad_layer_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->gu && a->x && a->gy && a->h && a->dw1 && a->dw2 && a->meta && a->workspace,
              "tcn_wgrad_layer: null pointer");
  TCN_REQUIRE(a->channels == 64, "tcn_wgrad_layer: 64 channels only (use tcn_wgrad_tc for other widths)");
  TCN_REQUIRE(a->nblk > 0 && a->rows >= (long long)a->nblk * kBlkRows, "tcn_wgrad_layer: bad shape");
  TCN_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f, "tcn_wgrad_layer: drop_p must be in [0, 1)");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->gu) | reinterpret_cast<uintptr_t>(a->x) |
                       reinterpret_cast<uintptr_t>(a->gy) | reinterpret_cast<uintptr_t>(a->h) |
                       reinterpret_cast<uintptr_t>(a->workspace);
  TCN_REQUIRE((al & 15) == 0, "tcn_wgrad_layer: operands must be 16-byte aligned");
  TCN_CHECK(wl_set_attr());
  int splits = num_sms();
  if (splits > a->nblk * 8) splits = a->nblk * 8;
  TCN_REQUIRE(a->workspace_bytes >= (long long)splits * WL_PART_FLOATS * 4,
              "tcn_wgrad_layer: workspace too small (tcn_wgrad_layer_workspace_bytes)");
  WgLayerDev d;
  memset(&d, 0, sizeof(d));
  TCN_CHECK(make_wgrad_layer_maps(&d, a->x, a->h, a->gu, a->gy, (long)a->rows));
  d.masks = a->masks;
  for (int i = 0; i < 3; ++i) d.shift[i] = a->shift[i];
  d.use_drop = a->drop_p > 0.f ? 1 : 0;
  d.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  d.drop_thresh = a->drop_p > 0.f ? drop_thresh(a->drop_p) : 0u;
  d.drop_seed = a->drop_seed; d.drop_stream = a->drop_stream;
  d.part = reinterpret_cast<float*>(a->workspace);
  WgLayersLaunch q;
  q.meta = reinterpret_cast<const BlkMeta*>(a->meta); q.nblk = a->nblk; q.dyn = nullptr; q.splits = splits;
  {
    const char* e = getenv("TCN_WL_DEBUG");
    q.debug = e ? atoi(e) : 0;
  }
  launch_kernel(wgrad_layer_kernel, dim3(splits, 1, 1), dim3(WL_THREADS), WL_SMEM, (cudaStream_t)stream, true, d, q);
  TCN_CHECK(check_launch("wgrad_layer_kernel"));
  if (a->flags & 1) return TCN_OK;
  WgLayerOut o;
  o.part = d.part; o.dw1 = a->dw1; o.db1 = a->db1; o.dw2 = a->dw2; o.db2 = a->db2;
  const int nb = (WL_PART_FLOATS / 4 + 31) / 32;
  launch_kernel(wgrad_layer_reduce_kernel, dim3(nb, 1, 1), dim3(256), 0, (cudaStream_t)stream, true, o, splits);
  return check_launch("wgrad_layer_reduce_kernel");
}
