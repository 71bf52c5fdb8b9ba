// Global_Relational_Block attention (MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py:76-88) on tcgen05:
//     attn = softmax(scale q k^T);  o = attn v                      (per window and head, windows of <= 256 frames)
// and its backward  dV = P^T dO,  dP = dO V^T,  dS = P (dP - rowsum(dP P)),  dQ = scale dS K,  dK = scale dS^T Q.
// The six products are ONE batched tensor-core kernel (`bgemm_tc_kernel`), instantiated by operand majorness:
//   shape 0 "NT"  C[m, n] = sum_k A[m, k] B[n, k]     A, B window tensors, K-major        S = Q K^T, dP = dO V^T   (K = head dim)
//   shape 1 "NN"  C[m, n] = sum_k A[m, k] B[k, n]     A = score buffer (K-major), B window tensor (MN-major)   O = P V, dQ = dS K
//   shape 2 "TN"  C[m, n] = sum_k A[k, m] B[k, n]     A = score buffer (MN-major), B window tensor (MN-major)  dV = P^T dO, dK = dS^T Q
// One CTA = (window, head, 128-row tile of C); the 128 x N accumulator (N <= 256) lives in TMEM; operands arrive by TMA
// in 32-deep k chunks (K-major: 128B-swizzled boxes; MN-major: 32-column boxes with 32-byte swizzle atoms, as in
// wgrad_tc.cu).  3xTF32 as everywhere: tcgen05.mma kind::tf32 ignores the low 13 mantissa bits, so the raw TMA tile
// is the hi operand; the split warps write lo = x - trunc(x) and zero whatever lies outside the window / the head
// (rows of the next window, columns of the next head).  The scores of a window (256 x 256 per head) are materialised
// once in HBM (P is what the backward pass needs anyway); softmax and its backward are row kernels.
//   warp 0: TMA producer | warp 1: MMA issuer + TMEM owner | warps 2-9: operand split, then epilogue
#include <cstring>

#include "gemm_tc.cuh"

namespace tcn {

constexpr int BG_THREADS = 320;
constexpr int BG_KC = 32;                       // k elements per stage
constexpr int BG_A = 128 * BG_KC * 4;           // 16384 B: A tile (either majorness)
constexpr int BG_BMAX = 256 * BG_KC * 4;        // 32768 B: largest B tile
constexpr int BG_STAGES = 3;                    // barrier slots; a launch uses 2 stages of 96 KB (256-row B tiles) or 3 of 64 KB
constexpr int BG_TILES = 2 * 2 * (BG_A + BG_BMAX);   // 196608 B of operand tiles either way
constexpr int BG_SMEM = BG_TILES + 1024 + 256;
constexpr int BG_BOX = 32 * 128;                // 4096 B: one MN-major box (32 k rows x 32 columns)

struct BgDev {
  int shape;              // 0 NT, 1 NN, 2 TN
  int heads, hd, tmax;    // tmax: rows / columns of one (window, head) block of the score buffer
  const int* seq_lo;      // [nseq] first row of each window in the window tensors
  const int* seq_len;     // [nseq]
  int colA0, colB0;       // column of head 0 in the A / B window tensors (score-buffer operands: unused)
  float* C;               // output: score buffer (shape 0) or window tensor (shapes 1, 2)
  int ldc, colC0;
  float alpha;
};

__device__ __forceinline__ uint64_t bg_desc_mn(uint32_t smem_addr) {   // MN-major, SWIZZLE_128B_BASE32B, LBO = one box
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((BG_BOX >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ uint32_t bg_idesc(int N, bool a_mn, bool b_mn) {   // kind::tf32, fp32 accumulate, M = 128
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(BG_THREADS, 1)
bgemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const BgDev p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mtile = blockIdx.x, prob = blockIdx.y;
  const int w = prob / p.heads, h = prob - w * p.heads;
  const int T = p.seq_len[w], lo = p.seq_lo[w];
  const int m0 = mtile * 128;
  const bool a_mn = p.shape == 2, b_mn = p.shape != 0;
  // contraction length, its valid part, and the (padded) width of the accumulator
  const int kvalid = p.shape == 0 ? p.hd : T;
  const int nk = (kvalid + BG_KC - 1) / BG_KC;
  const int nvalid = p.shape == 0 ? T : p.hd;
  const int npad = (nvalid + 15) & ~15;
  const int nboxb = (npad + 31) >> 5;                       // MN-major B: 32-column boxes
  const int bbytes = b_mn ? nboxb * BG_BOX : BG_BMAX;
  // stage = [A raw | B raw | A lo | B lo]; MN-major B tiles are at most 16 KB (head dim <= 128): three 64 KB stages
  const int bcap = (b_mn && nboxb <= 4) ? BG_A : BG_BMAX;
  const int BG_HALF = BG_A + bcap, BG_STAGE = 2 * BG_HALF;
  const int nst = bcap == BG_A ? 3 : 2;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + BG_TILES);
  uint64_t* full_bar = bars;
  uint64_t* ready_bar = bars + BG_STAGES;
  uint64_t* empty_bar = bars + 2 * BG_STAGES;
  uint64_t* accum_bar = bars + 3 * BG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * BG_STAGES + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < BG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 8);    // one arrival per split warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const bool active = m0 < T && nk > 0 && nvalid > 0;   // CTA-uniform
  if (active) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        const int srow = prob * p.tmax;   // first row of this (window, head) block in the score buffer
        for (int kc = 0; kc < nk; ++kc) {
          const int s = kc % nst;
          mbar_wait(&empty_bar[s], ((kc / nst) & 1) ^ 1);
          uint8_t* st = tiles + s * BG_STAGE;
          mbar_arrive_expect_tx(&full_bar[s], BG_A + bbytes);
          if (p.shape == 0) {
            tma_load_2d(st, &map_a, &full_bar[s], p.colA0 + h * p.hd + kc * BG_KC, lo + m0);          // 128 rows x 32 k
            tma_load_2d(st + BG_A, &map_b, &full_bar[s], p.colB0 + h * p.hd + kc * BG_KC, lo);        // 256 rows x 32 k
          } else {
            if (p.shape == 1) {
              tma_load_2d(st, &map_a, &full_bar[s], kc * BG_KC, srow + m0);                           // 128 rows x 32 k
            } else {
#pragma unroll
              for (int mb = 0; mb < 4; ++mb)                                                          // 32 k rows x 32 m columns
                tma_load_2d(st + mb * BG_BOX, &map_a, &full_bar[s], m0 + mb * 32, srow + kc * BG_KC);
            }
            for (int nb = 0; nb < nboxb; ++nb)                                                        // 32 k rows x 32 n columns
              tma_load_2d(st + BG_A + nb * BG_BOX, &map_b, &full_bar[s], p.colB0 + h * p.hd + nb * 32, lo + kc * BG_KC);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = bg_idesc(npad, a_mn, b_mn);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % nst;
        mbar_wait(&ready_bar[s], (kc / nst) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = base + s * BG_STAGE, a_lo = a_hi + BG_HALF;
          const uint32_t b_hi = a_hi + BG_A, b_lo = b_hi + BG_HALF;
          const int k8n = min(4, (kvalid - kc * BG_KC + 7) >> 3);
          for (int k = 0; k < k8n; ++k) {
            const uint32_t ao = a_mn ? k * 1024 : k * 32, bo = b_mn ? k * 1024 : k * 32;
            const uint64_t dah = a_mn ? bg_desc_mn(a_hi + ao) : umma_desc_sw128(a_hi + ao);
            const uint64_t dal = a_mn ? bg_desc_mn(a_lo + ao) : umma_desc_sw128(a_lo + ao);
            const uint64_t dbh = b_mn ? bg_desc_mn(b_hi + bo) : umma_desc_sw128(b_hi + bo);
            const uint64_t dbl = b_mn ? bg_desc_mn(b_lo + bo) : umma_desc_sw128(b_lo + bo);
            umma_tf32(tmem_base, dal, dbh, idesc, (kc | k) != 0);
            umma_tf32(tmem_base, dah, dbl, idesc, 1u);
            umma_tf32(tmem_base, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (kc == nk - 1) umma_commit(accum_bar);
        }
        __syncwarp();
      }
    } else {
      // ===================== operand split (warps 2..9): lo halves, zeros outside the window / head =====================
      const int ct = threadIdx.x - 64;   // 0..255
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % nst;
        if (lane == 0) mbar_wait(&full_bar[s], (kc / nst) & 1);
        __syncwarp();
        float4* raw = reinterpret_cast<float4*>(tiles + s * BG_STAGE);
        float4* lov = raw + BG_HALF / 16;
        const int k0 = kc * BG_KC;
        // ---- A tile: 1024 float4
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = ct + i * 256;
          bool ok;
          if (!a_mn) {   // [128 rows m][8 chunks of 4 k], 128B swizzle
            const int r = idx >> 3, c = ((idx & 7) ^ (r & 7)) << 2;
            ok = (m0 + r < T) && (k0 + c < kvalid);   // (kvalid is a multiple of 4: a float4 is inside or outside as a whole)
          } else {       // 4 boxes [32 rows k][8 chunks of 4 m], 32-byte swizzle atoms: only the k row matters
            const int r = (idx & 255) >> 3;
            ok = k0 + r < kvalid;
          }
          float4 v = raw[idx];
          if (!ok) {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            raw[idx] = v;
          }
          float4 l;
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          lov[idx] = l;
        }
        // ---- B tile: bbytes / 16 float4 behind the A tile
        const int nb4 = bbytes >> 4;
        for (int idx = ct; idx < nb4; idx += 256) {
          bool ok;
          if (!b_mn) {   // [256 rows n][8 chunks of 4 k]
            const int r = idx >> 3, c = ((idx & 7) ^ (r & 7)) << 2;
            ok = (r < T) && (k0 + c < kvalid);
          } else {       // boxes [32 rows k][32 columns n]
            const int r = (idx & 255) >> 3;
            ok = k0 + r < kvalid;
          }
          float4 v = raw[BG_A / 16 + idx];
          if (!ok) {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            raw[BG_A / 16 + idx] = v;
          }
          float4 l;
          l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
          l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
          l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          lov[BG_A / 16 + idx] = l;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[s]);
      }
      // ===================== epilogue: C tile = alpha * accumulator =====================
      if (lane == 0) mbar_wait(accum_bar, 0);
      __syncwarp();
      tc_fence_after();
      const int q = warp & 3, half = (warp - 2) >> 2;
      const int m = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      const size_t crow = p.shape == 0 ? (size_t)prob * p.tmax + m : (size_t)lo + m;
      const int ccol = p.shape == 0 ? 0 : p.colC0 + h * p.hd;
      float* crp = p.C + crow * p.ldc + ccol;
      const bool vec = ((p.ldc & 3) == 0) && ((ccol & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
      for (int c0 = half * 32; c0 < npad; c0 += 64) {
        float v[32];
        tmem_ld32(taddr + c0, v);   // warp-collective; stores are predicated
        if (m < T) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = c0 + j;
            if (vec && n + 3 < nvalid) {
              *reinterpret_cast<float4*>(crp + n) =
                  make_float4(v[j] * p.alpha, v[j + 1] * p.alpha, v[j + 2] * p.alpha, v[j + 3] * p.alpha);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < nvalid) crp[n + e] = v[j + e] * p.alpha;
            }
          }
        }
      }
    }
  }
  pdl_exit_fence();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(256u));
}

// P = softmax over the valid keys of every valid query row, in place in the score buffer; everything outside the window
// (columns >= T, rows >= T of a block) is set to zero so that later operand loads of the block are fully defined.
__global__ void __launch_bounds__(256) attn_softmax_rows_kernel(float* __restrict__ s, const int* __restrict__ seq_len,
                                                                 int heads, int tmax, int nprob) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= (long)nprob * tmax) return;
  const int prob = (int)(row / tmax), qi = (int)(row - (long)prob * tmax);
  const int T = seq_len[prob / heads];
  float* r = s + row * tmax;
  if (qi >= T) {
    for (int c = lane; c < tmax; c += 32) r[c] = 0.f;
    return;
  }
  float mx = -INFINITY;
  for (int c = lane; c < T; c += 32) mx = fmaxf(mx, r[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < T; c += 32) sum += expf(r[c] - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int c = lane; c < tmax; c += 32) r[c] = c < T ? expf(r[c] - mx) * inv : 0.f;
}

// dS = P (dP - sum_k dP P) in place of dP; zero outside the window
__global__ void __launch_bounds__(256) attn_softmax_bwd_rows_kernel(const float* __restrict__ pbuf, float* __restrict__ dp,
                                                                     const int* __restrict__ seq_len, int heads, int tmax,
                                                                     int nprob) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= (long)nprob * tmax) return;
  const int prob = (int)(row / tmax), qi = (int)(row - (long)prob * tmax);
  const int T = seq_len[prob / heads];
  const float* pr = pbuf + row * tmax;
  float* dr = dp + row * tmax;
  if (qi >= T) {
    for (int c = lane; c < tmax; c += 32) dr[c] = 0.f;
    return;
  }
  float dot = 0.f;
  for (int c = lane; c < T; c += 32) dot += dr[c] * pr[c];
  dot = warp_sum(dot);
  for (int c = lane; c < tmax; c += 32) dr[c] = c < T ? pr[c] * (dr[c] - dot) : 0.f;
}

}  // namespace tcn

using namespace tcn;

static int bg_launch(const CUtensorMap& ma, const CUtensorMap& mb, const BgDev& p, int nseq, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(bgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM);
    if (e != cudaSuccess) {
      set_error("attn_tc: smem attribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TCN_ERR_CUDA;
    }
    attr_set = true;
  }
  launch_kernel(bgemm_tc_kernel, dim3((p.tmax + 127) / 128, nseq * p.heads, 1), dim3(BG_THREADS), BG_SMEM, stream, true, ma,
                mb, p);
  return check_launch("bgemm_tc_kernel");
}

extern "C" int tcn_attn_tc_supported(int max_len, int heads, int head_dim, int ldq, int ldkv) {
  return (max_len > 0 && max_len <= 256 && heads > 0 && head_dim >= 4 && head_dim <= 256 && head_dim % 4 == 0 && ldq % 4 == 0 &&
          ldkv % 4 == 0) ? 1 : 0;
}

static int attn_tc_check(const tcn_attn_tc_args* a) {
  TCN_REQUIRE(a && a->q && a->k && a->v && a->p && a->seq_lo && a->seq_len, "tcn_attn_tc: null pointer");
  TCN_REQUIRE(a->nseq > 0 && a->rows > 0 && a->tmax > 0 && a->tmax <= 256 && a->tmax % 32 == 0,
              "tcn_attn_tc: windows of at most 256 frames (tmax a multiple of 32)");
  if (!tcn_attn_tc_supported(a->tmax, a->heads, a->head_dim, a->ldq, a->ldk) || a->ldv % 4 != 0) {
    set_error("tcn_attn_tc: head_dim and the leading dimensions must be multiples of 4 (TMA); use tcn_attn_fwd / _bwd");
    return TCN_ERR_UNSUPPORTED;
  }
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
                       reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->p);
  TCN_REQUIRE((al & 15) == 0, "tcn_attn_tc: operands must be 16-byte aligned");
  return TCN_OK;
}

static BgDev bg_base(const tcn_attn_tc_args* a, int shape) {
  BgDev p;
  memset(&p, 0, sizeof(p));
  p.shape = shape; p.heads = a->heads; p.hd = a->head_dim; p.tmax = a->tmax;
  p.seq_lo = a->seq_lo; p.seq_len = a->seq_len; p.alpha = 1.f;
  return p;
}

extern "C" int tcn_attn_fwd_tc(const tcn_attn_tc_args* a, tcn_stream_t stream) {
  TCN_CHECK(attn_tc_check(a));
  TCN_REQUIRE(a->o != nullptr, "tcn_attn_fwd_tc: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = a->heads * a->head_dim, nprob = a->nseq * a->heads;
  const long srows = (long)nprob * a->tmax;
  CUtensorMap mq, mk, mv_mn, mp;
  TCN_CHECK(make_tensor_map_2d(&mq, a->q, a->rows, d, a->ldq, 128));
  TCN_CHECK(make_tensor_map_2d(&mk, a->k, a->rows, d, a->ldk, 256));
  TCN_CHECK(make_tensor_map_2d(&mv_mn, a->v, a->rows, d, a->ldv, 32, true));
  TCN_CHECK(make_tensor_map_2d(&mp, a->p, srows, a->tmax, a->tmax, 128));
  {   // S = scale Q K^T
    BgDev p = bg_base(a, 0);
    p.C = a->p; p.ldc = a->tmax; p.alpha = a->scale;
    TCN_CHECK(bg_launch(mq, mk, p, a->nseq, st));
  }
  attn_softmax_rows_kernel<<<(unsigned)((srows + 7) / 8), 256, 0, st>>>(a->p, a->seq_len, a->heads, a->tmax, nprob);
  TCN_CHECK(check_launch("attn_softmax_rows_kernel"));
  {   // O = P V
    BgDev p = bg_base(a, 1);
    p.C = a->o; p.ldc = a->ldo;
    TCN_CHECK(bg_launch(mp, mv_mn, p, a->nseq, st));
  }
  return TCN_OK;
}

extern "C" int tcn_attn_bwd_tc(const tcn_attn_tc_args* a, tcn_stream_t stream) {
  TCN_CHECK(attn_tc_check(a));
  TCN_REQUIRE(a->dout && a->dq && a->dk && a->dv && a->dp, "tcn_attn_bwd_tc: null gradient / scratch pointer");
  TCN_REQUIRE(a->lddo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->dout) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->dp) & 15) == 0,
              "tcn_attn_bwd_tc: dout / dp must be 16-byte aligned with a leading dimension that is a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int d = a->heads * a->head_dim, nprob = a->nseq * a->heads;
  const long srows = (long)nprob * a->tmax;
  CUtensorMap mdo, mdo_mn, mv, mk_mn, mq_mn, mp_mn, mds, mds_mn;
  TCN_CHECK(make_tensor_map_2d(&mdo, a->dout, a->rows, d, a->lddo, 128));
  TCN_CHECK(make_tensor_map_2d(&mdo_mn, a->dout, a->rows, d, a->lddo, 32, true));
  TCN_CHECK(make_tensor_map_2d(&mv, a->v, a->rows, d, a->ldv, 256));
  TCN_CHECK(make_tensor_map_2d(&mk_mn, a->k, a->rows, d, a->ldk, 32, true));
  TCN_CHECK(make_tensor_map_2d(&mq_mn, a->q, a->rows, d, a->ldq, 32, true));
  TCN_CHECK(make_tensor_map_2d(&mp_mn, a->p, srows, a->tmax, a->tmax, 32, true));
  TCN_CHECK(make_tensor_map_2d(&mds, a->dp, srows, a->tmax, a->tmax, 128));
  TCN_CHECK(make_tensor_map_2d(&mds_mn, a->dp, srows, a->tmax, a->tmax, 32, true));
  {   // dV = P^T dO
    BgDev p = bg_base(a, 2);
    p.C = a->dv; p.ldc = a->lddv;
    TCN_CHECK(bg_launch(mp_mn, mdo_mn, p, a->nseq, st));
  }
  {   // dP = dO V^T
    BgDev p = bg_base(a, 0);
    p.C = a->dp; p.ldc = a->tmax;
    TCN_CHECK(bg_launch(mdo, mv, p, a->nseq, st));
  }
  attn_softmax_bwd_rows_kernel<<<(unsigned)((srows + 7) / 8), 256, 0, st>>>(a->p, a->dp, a->seq_len, a->heads, a->tmax, nprob);
  TCN_CHECK(check_launch("attn_softmax_bwd_rows_kernel"));
  {   // dQ = scale dS K
    BgDev p = bg_base(a, 1);
    p.C = a->dq; p.ldc = a->lddq; p.alpha = a->scale;
    TCN_CHECK(bg_launch(mds, mk_mn, p, a->nseq, st));
  }
  {   // dK = scale dS^T Q
    BgDev p = bg_base(a, 2);
    p.C = a->dk; p.ldc = a->lddk; p.alpha = a->scale;
    TCN_CHECK(bg_launch(mds_mn, mq_mn, p, a->nseq, st));
  }
  return TCN_OK;
}
