// tcgen05 + TMA weight-gradient kernel: contraction over frames (time)
//
//     dW[n, c, tap] += sum_r G[r, n] * X[r + shift[tap], c]          db[n] += sum_r G[r, n]
//
// the weight / bias gradients of every Conv1d / Linear on the path (closed forms of SURVEY.md 8a:
// gW2 = sum_t gv[t] h[t]^T, gW1[:, :, k] = sum_t gu[t] x[t + s_k]^T, the projection, lateral and head weights).
//
// Tensor-core mapping.  The reduction runs over rows, so both operands are "MN-major": a TMA box of
// 32 columns x 32 frames (128-byte swizzled rows) is read by tcgen05.mma with K = frames.
//   D[m, n'] (TMEM, 128 x 64 fp32):  m  = 4 blocks of 32 input channels, each block = one (tap, column block) of X
//                                    n' = 64 output channels of G
//   A = X blocks (hi / lo), B = G blocks (hi / lo), both split in the operand-split warps (the kernel's
//   operands are activations, so neither can be pre-split); 3 MMAs per 8-frame slice (lo*hi + hi*lo + hi*hi).
// One CTA owns one (m-tile, n-tile) of dW and a slice of the rows; it accumulates over all its rows in TMEM
// and adds its partial to dW once, at the end (fp32 atomics).  The dropout mask of the layer (gv = keep * gy /
// (1 - p)), Dropout2d's channel scale and the input mask of the projection are folded into the operand split.
//   warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2-9: operand split, bias sums, epilogue
//   (eight split warps: the per-chunk split is instruction-bound, 4 + NGA 16-byte chunks per thread)
#include <cstdlib>

#include "gemm_tc.cuh"

namespace tcn {

constexpr int WG_THREADS = 320;                // TMA, MMA, 8 x operand split / epilogue
constexpr int WG_SPLIT = 256;                  // operand-split threads
constexpr int WG_RC = 32;                      // frames per pipeline stage
constexpr int WG_ATOM = WG_RC * 128;           // 4096 B: 32 frames x 32 fp32 columns
// NGA = 32-column G blocks per tile (output-channel tile = 32 NGA): 2 for the 64-channel layers of the TCN, 4 for the
// wide Linear layers of MS-TCT, where a 128 x 128 tile cuts the shared-memory operand traffic per MAC by a third and
// splits every X block once per 128 output channels instead of once per 64 (8 = 128 x 256 is kept behind
// TCN_WGRAD_NGA=8: with only two stages fitting it measured slower).
template <int NGA>
struct WgSmem {
  static constexpr int kRaw = (4 + NGA) * WG_ATOM;   // 4 X blocks + NGA G blocks
  static constexpr int kStage = 2 * kRaw;            // raw / hi + lo
  static constexpr int kStages = NGA <= 2 ? 4 : (NGA <= 4 ? 3 : 2);
  static constexpr int kBytes = kStages * kStage + 1024 + 512 + NGA * 32 * 4;
};

// MN-major TF32 operand.  tcgen05 accepts exactly one shared-memory layout for it: 128-byte rows (one frame each,
// 32 fp32 columns) swizzled with 32-byte atoms (byte-address bits [5,7) ^= bits [7,9); TMA mode 128B_ATOM_32B),
// K-atoms of 4 frames 512 bytes apart (SBO), 32-column blocks `lbo` bytes apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int NGA>
__device__ __forceinline__ void wgrad_tc_body(const CUtensorMap* map_xp, const CUtensorMap* map_gp, const WgradTcDev& p,
                                              const int split, const int mtile, const int ntile) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int WG_RAW = WgSmem<NGA>::kRaw, WG_STAGE = WgSmem<NGA>::kStage, WG_STAGES = WgSmem<NGA>::kStages;
  constexpr int NCH = 4 + NGA;         // 16-byte chunks per split thread and stage (one per 32 x 32 block)
  constexpr int NCOL = NGA * 32;       // output channels per tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  const int blk_begin = (int)((long)split * nblk / p.row_splits);
  const int blk_end = (int)((long)(split + 1) * nblk / p.row_splits);
  const int nv = p.ntaps * p.cbn;  // virtual 32-column blocks of X

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + WG_STAGES * WG_STAGE);
  uint64_t* full_bar = bars;
  uint64_t* ready_bar = bars + WG_STAGES;
  uint64_t* empty_bar = bars + 2 * WG_STAGES;
  uint64_t* accum_bar = bars + 3 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * WG_STAGES + 1);
  float* bias_red = reinterpret_cast<float*>(bars + 3 * WG_STAGES + 2);  // NCOL floats

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], WG_SPLIT);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = threadIdx.x; i < NCOL; i += blockDim.x) bias_red[i] = 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)NCOL));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  // the (tap, column block) of each of the four X blocks of this m-tile (blocks past the end repeat the last one)
  int a_tap[4], a_cb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int v = min(mtile * 4 + j, nv - 1);
    a_tap[j] = v / p.cbn;
    a_cb[j] = v - a_tap[j] * p.cbn;
  }

  // every role walks the same chunk sequence: 32-frame chunks of the CTA's blocks that contain valid frames
  int total_chunks = 0;
  for (int blk = blk_begin; blk < blk_end; ++blk) {
    const BlkMeta m = p.meta[blk];
    const int valid = m.hi - blk * kBlkRows;
    if (valid > 0) total_chunks += min(4, (valid + WG_RC - 1) / WG_RC);
  }

  if (total_chunks > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        int it = 0;
        for (int blk = blk_begin; blk < blk_end; ++blk) {
          const BlkMeta m = p.meta[blk];
          const int dz = p.x_unpadded ? m.in_delta : 0;
          for (int ch = 0; ch < 4; ++ch) {
            const int r0 = blk * kBlkRows + ch * WG_RC;
            if (r0 >= m.hi) break;
            const int s = it % WG_STAGES;
            const uint32_t ph = (it / WG_STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = tiles + s * WG_STAGE;
            mbar_arrive_expect_tx(&full_bar[s], WG_RAW);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int sh = a_tap[j] == 0 ? p.shift[0] : (a_tap[j] == 1 ? p.shift[1] : p.shift[2]);
              tma_load_2d(st + j * WG_ATOM, map_xp, &full_bar[s], a_cb[j] * 32, r0 + dz + sh);
            }
#pragma unroll
            for (int j = 0; j < NGA; ++j)  // column blocks past the end of G are zero filled
              tma_load_2d(st + (4 + j) * WG_ATOM, map_gp, &full_bar[s], (ntile * NGA + j) * 32, r0);
            ++it;
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_tf32_mn(128, NCOL);
      for (int it = 0; it < total_chunks; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = base + s * WG_STAGE, a_lo = a_hi + WG_RAW;
          const uint32_t b_hi = a_hi + 4 * WG_ATOM, b_lo = b_hi + WG_RAW;
#pragma unroll
          for (int k = 0; k < WG_RC / 8; ++k) {
            const uint32_t ko = k * 1024;  // 8 frames = one 1024-byte swizzle atom
            const uint64_t dah = umma_desc_mn_sw128(a_hi + ko, WG_ATOM), dal = umma_desc_mn_sw128(a_lo + ko, WG_ATOM);
            const uint64_t dbh = umma_desc_mn_sw128(b_hi + ko, WG_ATOM), dbl = umma_desc_mn_sw128(b_lo + ko, WG_ATOM);
            umma_tf32(tmem_base, dal, dbh, idesc, (it | k) != 0);
            umma_tf32(tmem_base, dah, dbl, idesc, 1u);
            umma_tf32(tmem_base, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (it == total_chunks - 1) umma_commit(accum_bar);
        }
        __syncwarp();
      }
    } else {
      // ===================== operand split + bias sums (warps 2..9) =====================
      const int ct = threadIdx.x - 64;  // 0..255: chunk ct of every 32 x 32 block (frame ct / 8, 16-byte chunk ct % 8)
      const uint32_t g_seed = p.g_drop_seed ^ dseed, x_seed = p.x_drop_seed ^ dseed;
      // every 16-byte chunk this thread touches has the same position inside its 32 x 32 block up to a row offset:
      const int lc = ((((ct & 7) >> 1) ^ ((ct >> 3) & 3)) << 1) | (ct & 1);  // logical 16-byte column chunk (swizzle undone)
      const bool extras = p.colscale != nullptr || p.x_drop_thresh != 0u || p.g_drop_thresh != 0u;
      const bool raw_is_hi = NGA >= 4;   // measured: pays off for the 128 x 128 tiles of MS-TCT, not for the 128 x 64 ones
      float4 bsum[NGA];
#pragma unroll
      for (int a = 0; a < NGA; ++a) bsum[a] = make_float4(0.f, 0.f, 0.f, 0.f);
      int it = 0;
      // Dropout2d scale of the four X blocks: a thread always owns the same four columns of a block, and the scale only
      // changes with the sequence -- fetched once per sequence instead of once per 32-frame chunk
      float4 csc[4];
      int csc_seq = -1;
      for (int blk = blk_begin; blk < blk_end; ++blk) {
        const BlkMeta m = p.meta[blk];
        if (p.colscale != nullptr && m.seq != csc_seq && blk * kBlkRows < m.hi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = a_cb[j] * 32 + lc * 4;
            csc[j] = col < p.c_in ? __ldg(reinterpret_cast<const float4*>(p.colscale + (size_t)m.seq * p.colscale_ld + col))
                                  : make_float4(1.f, 1.f, 1.f, 1.f);
          }
          csc_seq = m.seq;
        }
        for (int ch = 0; ch < 4; ++ch) {
          const int r0 = blk * kBlkRows + ch * WG_RC;
          if (r0 >= m.hi) break;
          const int s = it % WG_STAGES;
          const uint32_t ph = (it / WG_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          float4* raw = reinterpret_cast<float4*>(tiles + s * WG_STAGE);
          float4* lo = reinterpret_cast<float4*>(tiles + s * WG_STAGE + WG_RAW);
          float4 v[NCH];
          bool dirty[NCH];   // the tile entry differs from what TMA delivered: the hi half must be stored back
#pragma unroll
          for (int i = 0; i < NCH; ++i) v[i] = raw[ct + i * WG_SPLIT];
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            const int atom = i;                         // chunk ct + 256 i lies in block i ...
            const int r = ct >> 3;                      // ... at frame r of the chunk
            int src = r0 + r;
            if (atom < 4) src += a_tap[atom] == 0 ? p.shift[0] : (a_tap[atom] == 1 ? p.shift[1] : p.shift[2]);
            // select, not multiply: a pad row may hold anything (0 * NaN would poison the accumulator)
            dirty[i] = extras;
            if (!(src >= m.lo && src < m.hi)) { v[i] = make_float4(0.f, 0.f, 0.f, 0.f); dirty[i] = true; }
            if (extras) {
              if (atom < 4) {
                const int col = a_cb[atom] * 32 + lc * 4;
                if (p.colscale != nullptr) {
                  const float4 sc = csc[atom < 4 ? atom : 0];
                  v[i].x *= sc.x; v[i].y *= sc.y; v[i].z *= sc.z; v[i].w *= sc.w;
                }
                if (p.x_drop_thresh != 0u) {
                  float f[4];
                  drop_factor4(x_seed, p.x_drop_stream, p.x_drop_thresh, p.x_drop_scale, src, col, f);
                  v[i].x *= f[0]; v[i].y *= f[1]; v[i].z *= f[2]; v[i].w *= f[3];
                }
              } else if (p.g_drop_thresh != 0u) {
                const int col = (ntile * NGA + (atom - 4)) * 32 + lc * 4;
                float f[4];
                drop_factor4(g_seed, p.g_drop_stream, p.g_drop_thresh, p.g_drop_scale, src, col, f);
                v[i].x *= f[0]; v[i].y *= f[1]; v[i].z *= f[2]; v[i].w *= f[3];
              }
            }
          }
#pragma unroll
          for (int i = 4; i < NCH; ++i) {  // G blocks: column sums for the bias gradient
            bsum[i - 4].x += v[i].x; bsum[i - 4].y += v[i].y;
            bsum[i - 4].z += v[i].z; bsum[i - 4].w += v[i].w;
          }
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u); l.x = v[i].x - h.x;
            h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u); l.y = v[i].y - h.y;
            h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u); l.z = v[i].z - h.z;
            h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u); l.w = v[i].w - h.w;
            // tcgen05.mma kind::tf32 ignores the low 13 mantissa bits: an untouched raw tile IS the hi operand
            if (dirty[i] || !raw_is_hi) raw[ct + i * WG_SPLIT] = h;
            lo[ct + i * WG_SPLIT] = l;
          }
          fence_proxy_async();
          mbar_arrive(&ready_bar[s]);
          ++it;
        }
      }
      // ===================== epilogue =====================
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      // ---- bias gradient: the per-thread column sums are added over the 32 frame positions in a fixed order through
      // shared memory (the stages are free once the accumulator is complete) -- no shared atomics: bit-reproducible
      if (p.db != nullptr && mtile == 0) {
        float* red = reinterpret_cast<float*>(tiles);   // [NGA][32 frame positions][32 columns]
#pragma unroll
        for (int a = 0; a < NGA; ++a)
          *reinterpret_cast<float4*>(red + ((a * 32 + (ct >> 3)) * 32) + lc * 4) = bsum[a];
        asm volatile("bar.sync 1, 256;\n" ::: "memory");  // the eight split warps only
        for (int i = ct; i < NCOL; i += WG_SPLIT) {
          const int a = i >> 5, col = i & 31;
          float sacc = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) sacc += red[(a * 32 + rr) * 32 + col];
          const int n = ntile * NCOL + i;
          if (n < p.n_out) {
            if (p.slab != nullptr) p.slab[(size_t)split * p.slab_stride + (size_t)p.n_out * p.c_in * p.ntaps + n] = sacc;
            else atomicAdd(p.db + n, sacc);
          }
        }
      }
      const int q = warp & 3;  // TMEM lane quadrant == X block of this m-tile; two warps per quadrant split the columns
      const int half = (warp - 2) >> 2;
      const int vblk = mtile * 4 + q;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int c = a_cb[q] * 32 + lane;
      const bool row_ok = (vblk < nv) && (c < p.c_in);
#pragma unroll 1
      for (int c0 = half * (NCOL / 2); c0 < (half + 1) * (NCOL / 2); c0 += 32) {
        if (ntile * NCOL + c0 >= p.n_out) break;  // warp-uniform
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = ntile * NCOL + c0 + j;
            if (n < p.n_out) {
              const size_t idx = ((size_t)n * p.c_in + c) * p.ntaps + a_tap[q];
              if (p.slab != nullptr) p.slab[(size_t)split * p.slab_stride + idx] = v[j];
              else atomicAdd(p.dW + idx, v[j]);
            }
          }
        }
      }
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)NCOL));
}

template <int NGA>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                const WgradTcDev p) {
  wgrad_tc_body<NGA>(&map_x, &map_g, p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// two independent problems in one launch (the two weight gradients of a residual layer): blockIdx.y < mt0 -> problem 0
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_g0,
                     const WgradTcDev p0, const __grid_constant__ CUtensorMap map_x1,
                     const __grid_constant__ CUtensorMap map_g1, const WgradTcDev p1, int mt0) {
  if ((int)blockIdx.y < mt0)
    wgrad_tc_body<2>(&map_x0, &map_g0, p0, blockIdx.x, blockIdx.y, 0);
  else
    wgrad_tc_body<2>(&map_x1, &map_g1, p1, blockIdx.x, blockIdx.y - mt0, 0);
}

template <int NGA>
static int launch_wgrad_tc_n(const CUtensorMap& mx, const CUtensorMap& mg, WgradTcDev& p, int nb, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(wgrad_tc_kernel<NGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<NGA>::kBytes);
    if (e != cudaSuccess) {
      set_error("wgrad_tc: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  const int mt = (p.ntaps * p.cbn + 3) / 4, nt = (p.n_out + NGA * 32 - 1) / (NGA * 32);
  const int sms = num_sms();
  int rs;
  if (mt * nt <= sms) {
    rs = sms / (mt * nt);
  } else {  // more tiles than SMs: split the rows so that the last wave is (nearly) full
    rs = 1;
    double best = 0.0;
    for (int r = 1; r <= 6; ++r) {
      const long units = (long)mt * nt * r;
      const double eff = (double)units / (double)(((units + sms - 1) / sms) * sms);
      if (eff > best + 0.03) { best = eff; rs = r; }
    }
  }
  if (rs < 1) rs = 1;
  if (rs > nb) rs = nb;
  if (p.slab != nullptr && rs > p.slab_splits) rs = p.slab_splits;
  p.row_splits = rs;
  launch_kernel(wgrad_tc_kernel<NGA>, dim3(rs, mt, nt), dim3(WG_THREADS), WgSmem<NGA>::kBytes, stream, true, mx, mg, p);
  return check_launch("wgrad_tc_kernel");
}

int wgrad_tc_splits_cap(int n_out, int c_in, int ntaps) {
  const int cbn = (c_in + 31) / 32;
  const int mt = (ntaps * cbn + 3) / 4;
  int best = 1;
  for (int nga : {2, 4, 8}) {   // whichever tile width launch_wgrad_tc picks
    const int nt = (n_out + nga * 32 - 1) / (nga * 32);
    const int rs = num_sms() / (mt * nt);
    if (rs > best) best = rs;
  }
  return best > 6 ? best : 6;
}

constexpr int SR_GROUPS = 8;
__global__ void __launch_bounds__(256)
slab_reduce_kernel(float* __restrict__ out, const float* __restrict__ slab, long n, int nsplit, long stride) {
  __shared__ float red[SR_GROUPS][32];
  pdl_wait();
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long i = (long)blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (i < n)
    for (int s = g; s < nsplit; s += SR_GROUPS) acc += __ldg(slab + (size_t)s * stride + i);
  red[g][lane] = acc;
  __syncthreads();
  if (g != 0 || i >= n) return;
#pragma unroll
  for (int k = 1; k < SR_GROUPS; ++k) acc += red[k][lane];
  out[i] += acc;
}
int launch_slab_reduce(float* out, const float* slab, long n, int nsplit, long stride, cudaStream_t stream) {
  launch_kernel(slab_reduce_kernel, dim3((unsigned)((n + 31) / 32)), dim3(256), 0, stream, true, out, slab, n, nsplit, stride);
  return check_launch("slab_reduce_kernel");
}

int launch_wgrad_tc(const CUtensorMap& mx, const CUtensorMap& mg, WgradTcDev& p, int cap_nblk, cudaStream_t stream) {
  p.cbn = (p.c_in + 31) / 32;
  const int nb = cap_nblk > 0 ? cap_nblk : p.nblk;
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("TCN_WGRAD_NGA");
    forced = e ? atoi(e) : -1;
  }
  int nga = 2;
  if (forced == 2 || forced == 4 || forced == 8) {
    nga = forced;
  } else if (p.n_out >= 128 && p.ntaps * p.cbn >= 4) {
    nga = 4;  // measured (tools/gemm_shapes_bench.py): 128 x 128 tiles with 3 stages beat 128 x 256 with 2
  }
  if (nga == 8) return launch_wgrad_tc_n<8>(mx, mg, p, nb, stream);
  if (nga == 4) return launch_wgrad_tc_n<4>(mx, mg, p, nb, stream);
  return launch_wgrad_tc_n<2>(mx, mg, p, nb, stream);
}

// blockIdx.y -> (problem, m-tile) through a device table; tensor maps are read from global memory
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_multi_kernel(const WgradMultiDesc* __restrict__ descs, const int2* __restrict__ tiles) {
  const int2 t = tiles[blockIdx.y];
  const WgradMultiDesc& d = descs[t.x];
  wgrad_tc_body<2>(&d.mx, &d.mg, d.p, blockIdx.x, t.y, 0);
}

int launch_wgrad_tc_multi(const WgradMultiDesc* descs_dev, const int2* tiles_dev, int ntiles, int row_splits,
                          cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(wgrad_tc_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<2>::kBytes);
    if (e != cudaSuccess) {
      set_error("wgrad_tc_multi: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_kernel(wgrad_tc_multi_kernel, dim3(row_splits, ntiles, 1), dim3(WG_THREADS), WgSmem<2>::kBytes, stream, true,
                descs_dev, tiles_dev);
  return check_launch("wgrad_tc_multi_kernel");
}

int launch_wgrad_tc_pair(const CUtensorMap& mx0, const CUtensorMap& mg0, WgradTcDev& p0, const CUtensorMap& mx1,
                         const CUtensorMap& mg1, WgradTcDev& p1, int cap_nblk, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(wgrad_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<2>::kBytes);
    if (e != cudaSuccess) {
      set_error("wgrad_tc_pair: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  p0.cbn = (p0.c_in + 31) / 32;
  p1.cbn = (p1.c_in + 31) / 32;
  const int mt0 = (p0.ntaps * p0.cbn + 3) / 4, mt1 = (p1.ntaps * p1.cbn + 3) / 4;
  if (p0.n_out > 64 || p1.n_out > 64) {
    set_error("wgrad_tc_pair: both problems must have n_out <= 64");
    return TCN_ERR_INVALID_ARG;
  }
  const int nb = cap_nblk > 0 ? cap_nblk : p0.nblk;
  int rs = num_sms() / (mt0 + mt1);
  if (rs < 1) rs = 1;
  if (rs > nb) rs = nb;
  p0.row_splits = rs;
  p1.row_splits = rs;
  launch_kernel(wgrad_tc_pair_kernel, dim3(rs, mt0 + mt1, 1), dim3(WG_THREADS), WgSmem<2>::kBytes, stream, true, mx0, mg0, p0, mx1,
                mg1, p1, mt0);
  return check_launch("wgrad_tc_pair_kernel");
}

}  // namespace tcn

using namespace tcn;

// argument check + tensor maps + device descriptor of one tcn_wgrad_tc_args problem
static int wgrad_tc_from_args(const tcn_wgrad_tc_args* a, CUtensorMap* mx, CUtensorMap* mg, WgradTcDev* out) {
  TCN_REQUIRE(a && a->g && a->x && a->dw && a->meta, "tcn_wgrad_tc: null pointer");
  TCN_REQUIRE(a->nblk > 0 && a->c_in > 0 && a->n_out > 0 && a->ntaps >= 1 && a->ntaps <= 3, "tcn_wgrad_tc: bad shape");
  if (a->ldx % 4 != 0 || a->ldg % 4 != 0 || a->c_in % 4 != 0) {
    set_error("tcn_wgrad_tc: ldx, ldg and c_in must be multiples of 4 (TMA row pitch); use tcn_wgrad");
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->g_cols >= a->n_out && a->g_cols <= a->ldg && a->x_rows > 0 && a->g_rows > 0, "tcn_wgrad_tc: bad shape");
  TCN_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->g) & 15) == 0,
              "tcn_wgrad_tc: x and g must be 16-byte aligned");
  TCN_REQUIRE(a->g_drop_p >= 0.f && a->g_drop_p < 1.f, "tcn_wgrad_tc: g_drop_p must be in [0, 1)");
  TCN_CHECK(make_tensor_map_2d(mx, a->x, a->x_rows, a->c_in, a->ldx, WG_RC, true));
  TCN_CHECK(make_tensor_map_2d(mg, a->g, a->g_rows, a->g_cols, a->ldg, WG_RC, true));
  WgradTcDev p;
  memset(&p, 0, sizeof(p));
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  p.x_unpadded = a->x_unpadded; p.n_out = a->n_out; p.c_in = a->c_in; p.ntaps = a->ntaps;
  for (int i = 0; i < 3; ++i) p.shift[i] = a->shift[i];
  p.dW = a->dw; p.db = a->db;
  p.colscale = a->colscale; p.colscale_ld = a->colscale_ld;
  p.g_drop_thresh = a->g_drop_p > 0.f ? drop_thresh(a->g_drop_p) : 0u;
  p.g_drop_scale = a->g_drop_p > 0.f ? 1.f / (1.f - a->g_drop_p) : 1.f;
  p.g_drop_seed = a->drop_seed; p.g_drop_stream = a->drop_stream;
  p.x_drop_scale = 1.f;
  *out = p;
  return TCN_OK;
}

extern "C" int tcn_wgrad_tc(const tcn_wgrad_tc_args* a, tcn_stream_t stream) {
  CUtensorMap mx, mg;
  WgradTcDev p;
  TCN_CHECK(wgrad_tc_from_args(a, &mx, &mg, &p));
  return launch_wgrad_tc(mx, mg, p, 0, (cudaStream_t)stream);
}

extern "C" int tcn_wgrad_tc_pair(const tcn_wgrad_tc_args* a0, const tcn_wgrad_tc_args* a1, tcn_stream_t stream) {
  CUtensorMap mx0, mg0, mx1, mg1;
  WgradTcDev p0, p1;
  TCN_CHECK(wgrad_tc_from_args(a0, &mx0, &mg0, &p0));
  TCN_CHECK(wgrad_tc_from_args(a1, &mx1, &mg1, &p1));
  TCN_REQUIRE(a0->meta == a1->meta && a0->nblk == a1->nblk, "tcn_wgrad_tc_pair: both problems must share the block table");
  TCN_REQUIRE(!a0->x_unpadded && !a1->x_unpadded, "tcn_wgrad_tc_pair: padded operands only");
  return launch_wgrad_tc_pair(mx0, mg0, p0, mx1, mg1, p1, 0, (cudaStream_t)stream);
}
