// Generic time-major "tap GEMM" family (tensor cores via mma.sync m16n8k8 TF32, 3-term split for
// fp32-level accuracy):
//
//   forward / dgrad : Y[r, n] = epi( sum_tap sum_c X[r + shift[tap], c] * W[n, c, tap] + bias[n] )
//   wgrad           : dW[n, c, tap] += sum_r G[r, n] * X[r + shift[tap], c],  db[n] += sum_r G[r, n]
//
// Rows are frames of a packed batch of sequences (BlkMeta, common.cuh); a tap that leaves its own
// sequence contributes zero, exactly like Conv1d zero padding / F.pad in the reference
// (MT4MTLKD/Temporal_tenco/network.py:178-198).  Everything with a 1x1 or k=3 convolution or a
// Linear layer on the path maps onto this family: the stage-input projection (network.py:113,129),
// the backward passes of the residual layers, the FPN lateral (network.py:98-106), the four heads
// (network.py:63-67) and the MS-TCT Linear / merge-conv layers (Temporal_mstct/MSTCT/*.py).
//
// Weights are consumed as "fragment-ordered, pre-split" buffers written once per optimizer step by
// prep_weight_kernel: for k-step ks (8 K-values) and n8-tile nt, lane l holds the float4
//   { B[ks*8 + (l&3)][nt*8 + (l>>2)], B[ks*8 + (l&3) + 4][...] } as (hi.x, hi.y, lo.x, lo.y),
// so a warp fetches its B fragments for one MMA with a single coalesced 512-byte read-only load.
#include <cstring>
#include "common.cuh"

namespace tcn {

// ------------------------------------------------------------------------------------ weight prep
__device__ __forceinline__ void prep_one(const float* __restrict__ w, int n_out, int c_in, int ntaps, int transpose,
                                         float4* __restrict__ wf, int NT8, int kpt, long i) {
  const int lane = (int)(i & 31);
  const long q = i >> 5;
  const int nt = (int)(q % NT8);
  const int ks = (int)(q / NT8);
  const int n = nt * 8 + (lane >> 2);
  float v[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = ks * 8 + (lane & 3) + 4 * j;
    const int tap = k / kpt;
    const int kc = k - tap * kpt;
    float x = 0.f;
    if (!transpose) {
      // logical B[k = tap*kpt + c][n] = W[n][c][tap]
      if (n < n_out && kc < c_in) x = w[((long)n * c_in + kc) * ntaps + tap];
    } else {
      // logical B[k = tap*kpt + o][c] = W[o][c][tap]   (dgrad: contraction over output channels)
      if (n < c_in && kc < n_out) x = w[((long)kc * c_in + n) * ntaps + tap];
    }
    v[j] = x;
  }
  uint32_t h0, l0, h1, l1;
  split_tf32(v[0], h0, l0);
  split_tf32(v[1], h1, l1);
  wf[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
}

__global__ void prep_weight_kernel(const float* __restrict__ w, int n_out, int c_in, int ntaps, int transpose,
                                   float4* __restrict__ wf, int KS, int NT8, int kpt) {
  const long total = (long)KS * NT8 * 32;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    prep_one(w, n_out, c_in, ntaps, transpose, wf, NT8, kpt, i);
}

// All weights of a model in one launch: jobs[j] describes one (weight, orientation) pair; `first`
// holds the exclusive prefix of float4 counts.
__global__ void prep_weight_batched_kernel(const PrepJob* __restrict__ jobs, int njobs, const float* __restrict__ params,
                                           float4* __restrict__ wf_base, long total) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {  // last job with first <= i
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const PrepJob jb = jobs[lo];
    prep_one(params + jb.src_off, jb.n_out, jb.c_in, jb.ntaps, jb.transpose, wf_base + jb.first, jb.NT8, jb.kpt,
             i - jb.first);
  }
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------ forward / dgrad
constexpr int G1_TM = 64;   // rows per CTA tile (half a BlkMeta block)
constexpr int G1_KC = 32;   // K columns staged per pipeline stage
constexpr int G1_LD = 36;   // padded smem row stride (floats): bank = (4g + t) -> conflict-free A fragments
constexpr int G1_THREADS = 128;

__global__ void __launch_bounds__(G1_THREADS) tapgemm_kernel(const TapGemmDev p) {
  __shared__ __align__(16) float As[2][G1_TM * G1_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;  // per-step seed when replayed from a graph
  const uint32_t out_seed = p.drop_seed ^ dseed, in_seed = p.in_drop_seed ^ dseed;
  const int n64 = (p.NT8 + 7) >> 3;
  const int total_tiles = nblk * 2 * n64;
  const int kchunks = (p.kpt + G1_KC - 1) / G1_KC;
  const int nchunks = p.ntaps * kchunks;

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int rt = tile / n64, ntile = tile - rt * n64;
    const int blk = rt >> 1;
    const BlkMeta m = p.meta[blk];
    const int row0 = blk * kBlkRows + (rt & 1) * G1_TM;
    if (row0 >= m.hi) continue;  // CTA-uniform
    const int nt0 = ntile * 8 + wn * 4;          // first n8-tile of this warp
    const int ntc = max(0, min(4, p.NT8 - nt0));  // valid n8-tiles of this warp
    const int dz = p.x_unpadded ? m.in_delta : 0;

    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

    auto issue = [&](int chunk, int buf) {
      const int tap = chunk / kchunks;
      const int c0 = (chunk - tap * kchunks) * G1_KC;
      const int sh = p.shift[tap];
#pragma unroll
      for (int i = 0; i < (G1_TM * G1_KC / 4) / G1_THREADS; ++i) {
        const int piece = tid + i * G1_THREADS;
        const int r = piece >> 3, cq = (piece & 7) * 4;
        const int row = row0 + r, src = row + sh;
        const bool valid = (row < m.hi) && (src >= m.lo) && (src < m.hi) && (c0 + cq < p.c_in);
        const float* gp = valid ? (p.X + (size_t)(src + dz) * p.ldx + c0 + cq) : p.X;
        cp_async16(&As[buf][r * G1_LD + cq], gp, valid);
      }
      cp_async_commit();
    };

    issue(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      if (chunk + 1 < nchunks) {
        issue(chunk + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();

      const int tap = chunk / kchunks;
      const int c0 = (chunk - tap * kchunks) * G1_KC;
      const int ksn = min(G1_KC, p.kpt - c0) >> 3;
      const int ks_base = (tap * p.kpt + c0) >> 3;
      const int sh = p.shift[tap];
      const float* Ab = &As[buf][(wm * 32) * G1_LD];
      if (ntc > 0) {
        for (int kk = 0; kk < ksn; ++kk) {
          uint32_t ahi[2][4], alo[2][4], bhi[4][2], blo[4][2];
          const float4* wp = p.Wf + ((size_t)(ks_base + kk) * p.NT8 + nt0) * 32 + lane;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (nt < ntc) {
              const float4 w = __ldg(wp + nt * 32);
              bhi[nt][0] = __float_as_uint(w.x); bhi[nt][1] = __float_as_uint(w.y);
              blo[nt][0] = __float_as_uint(w.z); blo[nt][1] = __float_as_uint(w.w);
            }
          }
          const int c = c0 + kk * 8 + t;
          float s0 = 1.f, s1 = 1.f;
          if (p.colscale != nullptr) {
            s0 = (c < p.c_in) ? __ldg(p.colscale + (size_t)m.seq * p.colscale_ld + c) : 0.f;
            s1 = (c + 4 < p.c_in) ? __ldg(p.colscale + (size_t)m.seq * p.colscale_ld + c + 4) : 0.f;
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float* ap = Ab + (mt * 16 + g) * G1_LD + kk * 8 + t;
            float a0 = ap[0] * s0, a1 = ap[8 * G1_LD] * s0, a2 = ap[4] * s1, a3 = ap[8 * G1_LD + 4] * s1;
            if (p.in_drop_thresh != 0u) {
              const int r_lo = row0 + wm * 32 + mt * 16 + g + sh;
              a0 *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, r_lo, c);
              a1 *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, r_lo + 8, c);
              a2 *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, r_lo, c + 4);
              a3 *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, r_lo + 8, c + 4);
            }
            split_tf32(a0, ahi[mt][0], alo[mt][0]);
            split_tf32(a1, ahi[mt][1], alo[mt][1]);
            split_tf32(a2, ahi[mt][2], alo[mt][2]);
            split_tf32(a3, ahi[mt][3], alo[mt][3]);
          }
          mma_block_3xtf32<2, 4>(acc, ahi, alo, bhi, blo, ntc);
        }
      }
      __syncthreads();
    }

    // ---- epilogue: bias, relu, relu-mask, dropout, residual, store
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = row0 + wm * 32 + mt * 16 + g + half * 8;
        if (row >= m.hi) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (nt >= ntc) continue;
          const int col = (nt0 + nt) * 8 + 2 * t;
          if (col >= p.n_out) continue;
          const bool has1 = (col + 1 < p.n_out);
          float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
          if (p.bias != nullptr) {
            v0 += __ldg(p.bias + col);
            if (has1) v1 += __ldg(p.bias + col + 1);
          }
          if (p.relu) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          if (p.M != nullptr) {
            const float* mp = p.M + (size_t)row * p.ldm + col;
            if (!(mp[0] > 0.f)) v0 = 0.f;
            if (has1 && !(mp[1] > 0.f)) v1 = 0.f;
          }
          if (p.drop_thresh != 0u) {
            v0 *= drop_factor(out_seed, p.drop_stream, p.drop_thresh, p.drop_scale, row, col);
            v1 *= drop_factor(out_seed, p.drop_stream, p.drop_thresh, p.drop_scale, row, col + 1);
          }
          if (p.R != nullptr) {
            const float* rp = p.R + (size_t)row * p.ldr + col;
            v0 += rp[0];
            if (has1) v1 += rp[1];
          }
          float* yp = p.Y + (size_t)row * p.ldy + col;
          if (has1 && ((p.ldy & 1) == 0)) {
            *reinterpret_cast<float2*>(yp) = make_float2(v0, v1);
          } else {
            yp[0] = v0;
            if (has1) yp[1] = v1;
          }
        }
      }
    }
  }
}

int launch_tapgemm(TapGemmDev& p, int grid_cap_blocks, cudaStream_t stream) {
  const int n64 = (p.NT8 + 7) / 8;
  long tiles = (long)p.nblk * 2 * n64;
  const long cap = grid_cap_blocks > 0 ? grid_cap_blocks : (long)num_sms() * 6;
  if (tiles > cap) tiles = cap;
  if (tiles < 1) tiles = 1;
  tapgemm_kernel<<<(int)tiles, G1_THREADS, 0, stream>>>(p);
  return check_launch("tapgemm_kernel");
}

// ------------------------------------------------------------------------------------ wgrad
constexpr int G2_RC = 32;  // rows (K) per pipeline stage
constexpr int G2_LD = 72;  // smem row stride: bank = (8t + g) -> conflict-free transposed fragments
constexpr int G2_THREADS = 128;

__global__ void __launch_bounds__(G2_THREADS) wgrad_kernel(const WgradDev p) {
  __shared__ __align__(16) float Gs[2][G2_RC * G2_LD];
  __shared__ __align__(16) float Xs[2][G2_RC * G2_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;
  const uint32_t dseed = p.dyn ? p.dyn->seed : 0u;
  const uint32_t g_seed = p.g_drop_seed ^ dseed, x_seed = p.x_drop_seed ^ dseed;
  int bid = blockIdx.x;
  const int split = bid % p.row_splits;
  bid /= p.row_splits;
  const int tap = bid % p.ntaps;
  bid /= p.ntaps;
  const int ct = bid % p.c_tiles;
  const int ntile = bid / p.c_tiles;
  const int n0 = ntile * 64, c0 = ct * 64;
  const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
  const int blk_begin = (int)((long)split * nblk / p.row_splits);
  const int blk_end = (int)((long)(split + 1) * nblk / p.row_splits);
  const int wn = (warp & 1) * 32, wc = (warp >> 1) * 32;
  const bool do_bias = (p.db != nullptr) && ct == 0 && tap == 0;

  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  float bias_acc = 0.f;

  constexpr int CPB = kBlkRows / G2_RC;  // chunks per block
  const int nchunks = (blk_end - blk_begin) * CPB;

  auto issue = [&](int chunk, int buf) {
    const int blk = blk_begin + chunk / CPB;
    const int r0 = blk * kBlkRows + (chunk % CPB) * G2_RC;
    const BlkMeta m = p.meta[blk];
    const int dz = p.x_unpadded ? m.in_delta : 0;
#pragma unroll
    for (int i = 0; i < (G2_RC * 64 / 4) / G2_THREADS; ++i) {
      const int piece = tid + i * G2_THREADS;
      const int r = piece >> 4, cq = (piece & 15) * 4;
      const int row = r0 + r, src = row + sh;
      const bool vg = (row < m.hi) && (n0 + cq < p.g_cols);
      const float* gp = vg ? (p.G + (size_t)row * p.ldg + n0 + cq) : p.G;
      cp_async16(&Gs[buf][r * G2_LD + cq], gp, vg);
      const bool vx = (row < m.hi) && (src >= m.lo) && (src < m.hi) && (c0 + cq < p.c_in);
      const float* xp = vx ? (p.X + (size_t)(src + dz) * p.ldx + c0 + cq) : p.X;
      cp_async16(&Xs[buf][r * G2_LD + cq], xp, vx);
    }
    cp_async_commit();
  };

  if (nchunks > 0) issue(0, 0);
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) {
      issue(chunk + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int blk = blk_begin + chunk / CPB;
    const int r0 = blk * kBlkRows + (chunk % CPB) * G2_RC;
    const BlkMeta m = p.meta[blk];
    if (r0 < m.hi) {  // CTA-uniform: chunks past the end of the sequence hold only zeros
      float* gs = Gs[buf];
      const float* xs = Xs[buf];
      if (p.g_drop_thresh != 0u) {  // gv = keep * gy / (1 - p), applied in place once per chunk
        for (int i = tid; i < G2_RC * 64; i += G2_THREADS) {
          const int r = i >> 6, cc = i & 63;
          gs[r * G2_LD + cc] *= drop_factor(g_seed, p.g_drop_stream, p.g_drop_thresh, p.g_drop_scale, r0 + r,
                                            n0 + cc);
        }
        __syncthreads();
      }
      if (p.x_drop_thresh != 0u) {  // input keep-mask on X (same key as the forward projection)
        float* xw = Xs[buf];
        for (int i = tid; i < G2_RC * 64; i += G2_THREADS) {
          const int r = i >> 6, cc = i & 63;
          xw[r * G2_LD + cc] *= drop_factor(x_seed, p.x_drop_stream, p.x_drop_thresh, p.x_drop_scale, r0 + r + sh,
                                            c0 + cc);
        }
        __syncthreads();
      }
#pragma unroll
      for (int kk = 0; kk < G2_RC / 8; ++kk) {
        const int k0 = kk * 8;
        uint32_t ahi[2][4], alo[2][4], bhi[4][2], blo[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int mrow = wn + mt * 16 + g;
          split_tf32(gs[(k0 + t) * G2_LD + mrow], ahi[mt][0], alo[mt][0]);
          split_tf32(gs[(k0 + t) * G2_LD + mrow + 8], ahi[mt][1], alo[mt][1]);
          split_tf32(gs[(k0 + t + 4) * G2_LD + mrow], ahi[mt][2], alo[mt][2]);
          split_tf32(gs[(k0 + t + 4) * G2_LD + mrow + 8], ahi[mt][3], alo[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int ncol = wc + nt * 8 + g;
          float x0 = xs[(k0 + t) * G2_LD + ncol], x1 = xs[(k0 + t + 4) * G2_LD + ncol];
          if (p.colscale != nullptr) {
            const int c = c0 + ncol;
            const float s = (c < p.c_in) ? __ldg(p.colscale + (size_t)m.seq * p.colscale_ld + c) : 0.f;
            x0 *= s;
            x1 *= s;
          }
          split_tf32(x0, bhi[nt][0], blo[nt][0]);
          split_tf32(x1, bhi[nt][1], blo[nt][1]);
        }
        mma_block_3xtf32<2, 4>(acc, ahi, alo, bhi, blo);
      }
      if (do_bias && tid < 64) {
#pragma unroll 8
        for (int r = 0; r < G2_RC; ++r) bias_acc += gs[r * G2_LD + tid];
      }
    }
    __syncthreads();
  }

  // ---- epilogue: accumulate the partial into dW (torch layout [n][c][tap]) and db
  if (nchunks > 0) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = n0 + wn + mt * 16 + g + (e >> 1) * 8;
          const int c = c0 + wc + nt * 8 + 2 * t + (e & 1);
          if (n < p.n_out && c < p.c_in) atomicAdd(p.dW + ((size_t)n * p.c_in + c) * p.ntaps + tap, acc[mt][nt][e]);
        }
    if (do_bias && tid < 64 && n0 + tid < p.n_out) atomicAdd(p.db + n0 + tid, bias_acc);
  }
}

int launch_wgrad(WgradDev& p, int cap_nblk, cudaStream_t stream) {
  p.n_tiles = (p.n_out + 63) / 64;
  p.c_tiles = (p.c_in + 63) / 64;
  const long base = (long)p.n_tiles * p.c_tiles * p.ntaps;
  const int nb = cap_nblk > 0 ? cap_nblk : p.nblk;
  long want = (2L * num_sms() + base - 1) / base;
  if (want < 1) want = 1;
  if (want > nb) want = nb;
  p.row_splits = (int)want;
  const long grid = base * p.row_splits;
  if (grid >= (1L << 31)) {
    set_error("wgrad: grid too large");
    return TCN_ERR_INVALID_ARG;
  }
  wgrad_kernel<<<(int)grid, G2_THREADS, 0, stream>>>(p);
  return check_launch("wgrad_kernel");
}

}  // namespace tcn

// ================================================================================================ C ABI
using namespace tcn;

extern "C" long long tcn_prep_weight_floats(int n_out, int c_in, int ntaps, int transpose) {
  const int kdim = transpose ? n_out : c_in;
  const int ncols = transpose ? c_in : n_out;
  const int kpt = round_up(kdim, 8);
  const long long KS = (long long)ntaps * kpt / 8, NT8 = (ncols + 7) / 8;
  return KS * NT8 * 32 * 4;
}

extern "C" int tcn_prep_weight(const float* w, int n_out, int c_in, int ntaps, int transpose, float* wf,
                               tcn_stream_t stream) {
  TCN_REQUIRE(w && wf && n_out > 0 && c_in > 0 && ntaps >= 1 && ntaps <= 3, "tcn_prep_weight: bad arguments");
  const int kdim = transpose ? n_out : c_in;
  const int ncols = transpose ? c_in : n_out;
  const int kpt = round_up(kdim, 8);
  const int KS = ntaps * kpt / 8, NT8 = (ncols + 7) / 8;
  const long total = (long)KS * NT8 * 32;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads);
  prep_weight_kernel<<<blocks > 4096 ? 4096 : blocks, threads, 0, (cudaStream_t)stream>>>(
      w, n_out, c_in, ntaps, transpose, reinterpret_cast<float4*>(wf), KS, NT8, kpt);
  return check_launch("prep_weight_kernel");
}

namespace tcn {
int launch_prep_batched(const PrepJob* jobs_dev, int njobs, const float* params, float* wf_base, long total_f4,
                        cudaStream_t stream) {
  const int threads = 256;
  long blocks = (total_f4 + threads - 1) / threads;
  if (blocks > 2048) blocks = 2048;
  prep_weight_batched_kernel<<<(int)blocks, threads, 0, stream>>>(jobs_dev, njobs, params,
                                                                 reinterpret_cast<float4*>(wf_base), total_f4);
  return check_launch("prep_weight_batched_kernel");
}
}  // namespace tcn

extern "C" int tcn_tapgemm(const tcn_tapgemm_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->x && a->wf && a->y && a->meta, "tcn_tapgemm: null pointer");
  TCN_REQUIRE(a->nblk > 0 && a->c_in > 0 && a->n_out > 0, "tcn_tapgemm: empty problem");
  TCN_REQUIRE(a->ntaps >= 1 && a->ntaps <= 3, "tcn_tapgemm: ntaps must be 1..3");
  TCN_REQUIRE((a->ldx % 4) == 0 && (a->c_in % 4) == 0, "tcn_tapgemm: ldx and c_in must be multiples of 4 (16-byte rows)");
  TCN_REQUIRE(a->ldx >= a->c_in && a->ldy >= a->n_out, "tcn_tapgemm: leading dimension smaller than the row");
  TCN_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->wf) & 15) == 0,
              "tcn_tapgemm: x and wf must be 16-byte aligned");
  TCN_REQUIRE(a->drop_p >= 0.f && a->drop_p < 1.f && a->in_drop_p >= 0.f && a->in_drop_p < 1.f,
              "tcn_tapgemm: drop_p must be in [0, 1)");
  TapGemmDev p;
  p.X = a->x; p.ldx = a->ldx; p.x_unpadded = a->x_unpadded;
  p.colscale = a->colscale; p.colscale_ld = a->colscale_ld;
  p.Wf = reinterpret_cast<const float4*>(a->wf); p.bias = a->bias;
  p.Y = a->y; p.ldy = a->ldy; p.R = a->residual; p.ldr = a->ldr; p.M = a->relu_mask; p.ldm = a->ldm;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  p.kpt = round_up(a->c_in, 8); p.c_in = a->c_in; p.n_out = a->n_out; p.NT8 = (a->n_out + 7) / 8;
  p.ntaps = a->ntaps;
  for (int i = 0; i < 3; ++i) p.shift[i] = a->shift[i];
  p.relu = a->relu;
  p.drop_thresh = a->drop_p > 0.f ? drop_thresh(a->drop_p) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.drop_seed = a->drop_seed; p.drop_stream = a->drop_stream;
  p.in_drop_thresh = a->in_drop_p > 0.f ? drop_thresh(a->in_drop_p) : 0u;
  p.in_drop_scale = a->in_drop_p > 0.f ? 1.f / (1.f - a->in_drop_p) : 1.f;
  p.in_drop_seed = a->drop_seed; p.in_drop_stream = a->drop_stream;
  return launch_tapgemm(p, 0, (cudaStream_t)stream);
}

extern "C" int tcn_wgrad(const tcn_wgrad_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->g && a->x && a->dw && a->meta, "tcn_wgrad: null pointer");
  TCN_REQUIRE(a->nblk > 0 && a->c_in > 0 && a->n_out > 0, "tcn_wgrad: empty problem");
  TCN_REQUIRE(a->ntaps >= 1 && a->ntaps <= 3, "tcn_wgrad: ntaps must be 1..3");
  TCN_REQUIRE((a->ldx % 4) == 0 && (a->c_in % 4) == 0 && (a->ldg % 4) == 0,
              "tcn_wgrad: ldx, ldg and c_in must be multiples of 4");
  TCN_REQUIRE(a->g_cols >= a->n_out && a->g_cols <= a->ldg && (a->g_cols % 4) == 0,
              "tcn_wgrad: g_cols must be a multiple of 4 in [n_out, ldg]");
  TCN_REQUIRE(a->g_drop_p >= 0.f && a->g_drop_p < 1.f, "tcn_wgrad: g_drop_p must be in [0, 1)");
  WgradDev p;
  memset(&p, 0, sizeof(p));
  p.G = a->g; p.ldg = a->ldg; p.g_cols = a->g_cols;
  p.X = a->x; p.ldx = a->ldx; p.x_unpadded = a->x_unpadded;
  p.colscale = a->colscale; p.colscale_ld = a->colscale_ld;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  p.n_out = a->n_out; p.c_in = a->c_in; p.ntaps = a->ntaps;
  for (int i = 0; i < 3; ++i) p.shift[i] = a->shift[i];
  p.dW = a->dw; p.db = a->db;
  p.g_drop_thresh = a->g_drop_p > 0.f ? drop_thresh(a->g_drop_p) : 0u;
  p.g_drop_scale = a->g_drop_p > 0.f ? 1.f / (1.f - a->g_drop_p) : 1.f;
  p.g_drop_seed = a->drop_seed; p.g_drop_stream = a->drop_stream;
  p.x_drop_thresh = 0u; p.x_drop_scale = 1.f; p.x_drop_seed = 0u; p.x_drop_stream = 0u;
  return launch_wgrad(p, 0, (cudaStream_t)stream);
}
