// MS-TCT temporal blocks (MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py): the pieces that are not
// per-frame dense contractions (those run on gemm_tc / wgrad_tc):
//   layernorm_{fwd,bwd}      nn.LayerNorm over channels           (Temporal_Encoder.py:97,103,140,160,175-199)
//   attn_fwd / attn_bwd_*    Global_Relational_Block attention     (:76-88) softmax(q k^T * hd^-0.5) v per (window, head)
//   dwconv_gelu_{fwd,bwd}    Local_Relational_Block depthwise conv k=3 + GELU over time (:13-14,36-39)
//   axpby                    y = a x + b y (residual bookkeeping of the Temporal_Mixer, TS_Mixer.py:66-76)
// All tensors time-major fp32 (rows = frames of the packed windows, columns = channels), fp32 arithmetic.
#include "common.cuh"

namespace tcn {

// ------------------------------------------------------------------------------------------------ LayerNorm
// one warp per row; mean / rstd saved for the backward pass
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                            int ldy, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ mean,
                                                            float* __restrict__ rstd, const BlkMeta* meta, int nrows,
                                                            int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    if (meta != nullptr && row >= meta[row / kBlkRows].hi) continue;
    const float* xr = x + (size_t)row * ldx;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mu = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = xr[c] - mu;
      v += d * d;
    }
    const float rs = rsqrtf(warp_sum(v) / (float)C + eps);
    float* yr = y + (size_t)row * ldy;
    for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// dx = rstd * (g*dy - mean_c(g*dy) - xhat * mean_c(g*dy*xhat));  dgamma += sum_r dy*xhat, dbeta += sum_r dy
// (column sums: per-CTA partials in shared memory, then one atomic per column and CTA)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, int ldx,
                                                            const float* __restrict__ dy, int lddy,
                                                            float* __restrict__ dx, int lddx,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            const BlkMeta* meta, int nrows, int C) {
  extern __shared__ float sm[];  // [2][C] column partials
  float* sg = sm;
  float* sb = sm + C;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    if (meta != nullptr && row >= meta[row / kBlkRows].hi) continue;
    const float* xr = x + (size_t)row * ldx;
    const float* gr = dy + (size_t)row * lddy;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs, gg = gr[c] * __ldg(gamma + c);
      s1 += gg;
      s2 += gg * xh;
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    float* dr = dx + (size_t)row * lddx;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs, g = gr[c];
      dr[c] = rs * (g * __ldg(gamma + c) - s1 - xh * s2);
      atomicAdd(&sg[c], g * xh);
      atomicAdd(&sb[c], g);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (sg[c] != 0.f) atomicAdd(dgamma + c, sg[c]);
    if (sb[c] != 0.f) atomicAdd(dbeta + c, sb[c]);
  }
}

// ------------------------------------------------------------------------------------------------ attention
// q: (rows, ldq) with head h at columns [h*hd, (h+1)*hd); k and v likewise inside their buffers (column offsets
// given by the caller).  One CTA = (sequence, head, 16 queries); 4 warps x 4 queries; keys / values streamed through
// shared memory in chunks of 64 with an online softmax.  lse[row * heads + h] = log sum exp of the scaled scores.
constexpr int AT_Q = 16, AT_KC = 64, AT_THREADS = 128, AT_MAXHD = 128;

struct AttnDev {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
  float* o; int ldo;
  float* lse;
  const float* dout; int lddo;
  float* dq; int lddq;
  float* dk; int lddk;
  float* dv; int lddv;
  const int* seq_lo;  // [nseq] first row of each sequence
  const int* seq_len; // [nseq]
  int heads, hd;
  float scale;
};

__global__ void __launch_bounds__(AT_THREADS) attn_fwd_kernel(const AttnDev p, int qtiles) {
  extern __shared__ float sm[];
  const int hd = p.hd, ldh = hd + 1;
  float* ks = sm;                       // [AT_KC][ldh]
  float* vs = ks + AT_KC * ldh;         // [AT_KC][ldh]
  float* qs = vs + AT_KC * ldh;         // [AT_Q][ldh]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bid = blockIdx.x;
  const int qt = bid % qtiles; bid /= qtiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int q0 = qt * AT_Q;
  if (q0 >= T) return;
  const int col = h * hd;
  for (int i = threadIdx.x; i < AT_Q * hd; i += AT_THREADS) {
    const int r = i / hd, d = i - r * hd;
    qs[r * ldh + d] = (q0 + r < T) ? p.q[(size_t)(lo + q0 + r) * p.ldq + col + d] * p.scale : 0.f;
  }
  float m_run[4], l_run[4], acc[4][4];  // 4 queries per warp, up to 128 value columns = 4 per lane
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY; l_run[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  }
  for (int k0 = 0; k0 < T; k0 += AT_KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_KC * hd; i += AT_THREADS) {
      const int r = i / hd, d = i - r * hd;
      const bool ok = k0 + r < T;
      ks[r * ldh + d] = ok ? p.k[(size_t)(lo + k0 + r) * p.ldk + col + d] : 0.f;
      vs[r * ldh + d] = ok ? p.v[(size_t)(lo + k0 + r) * p.ldv + col + d] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int qi = warp * 4 + a;
      if (q0 + qi >= T) continue;  // warp-uniform
      const float* qr = qs + qi * ldh;
      float s[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = lane + u * 32;
        float d = 0.f;
        const float* kr = ks + j * ldh;
        for (int e = 0; e < hd; ++e) d += qr[e] * kr[e];
        s[u] = (k0 + j < T) ? d : -INFINITY;
      }
      const float mx = fmaxf(m_run[a], warp_max(fmaxf(s[0], s[1])));
      const float corr = expf(m_run[a] - mx);
      const float p0 = expf(s[0] - mx), p1 = expf(s[1] - mx);
      l_run[a] = l_run[a] * corr + warp_sum(p0 + p1);
      m_run[a] = mx;
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] *= corr;
      for (int j = 0; j < AT_KC; ++j) {
        const float pj = __shfl_sync(0xffffffffu, (j < 32) ? p0 : p1, j & 31);
        const float* vr = vs + j * ldh;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int d = lane + b * 32;
          if (d < hd) acc[a][b] += pj * vr[d];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int qi = q0 + warp * 4 + a;
    if (qi >= T) continue;
    const float inv = 1.f / l_run[a];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int d = lane + b * 32;
      if (d < hd) p.o[(size_t)(lo + qi) * p.ldo + col + d] = acc[a][b] * inv;
    }
    if (lane == 0) p.lse[(size_t)(lo + qi) * p.heads + h] = m_run[a] + logf(l_run[a]);
  }
}

// backward, pass 1: one CTA = (sequence, head, 64 keys) accumulates dK, dV for its keys over all queries.
// pass 2: one CTA = (sequence, head, 16 queries) accumulates dQ over all keys.  No atomics.
// P_ij = exp(scale q_i.k_j - lse_i);  D_i = dO_i . O_i;  dV_j = sum_i P_ij dO_i;  dS_ij = P_ij (dO_i.v_j - D_i);
// dQ_i = scale sum_j dS_ij k_j;  dK_j = scale sum_i dS_ij q_i
__global__ void __launch_bounds__(AT_THREADS) attn_bwd_kv_kernel(const AttnDev p, int ktiles) {
  extern __shared__ float sm[];
  const int hd = p.hd, ldh = hd + 1;
  float* ks = sm;                 // [64][ldh]
  float* vs = ks + AT_KC * ldh;   // [64][ldh]
  float* qs = vs + AT_KC * ldh;   // [AT_Q][ldh]  (scaled q)
  float* gs = qs + AT_Q * ldh;    // [AT_Q][ldh]  dO
  float* ps = gs + AT_Q * ldh;    // [AT_Q][64]   P
  float* ds = ps + AT_Q * AT_KC;  // [AT_Q][64]   dS
  float* Di = ds + AT_Q * AT_KC;  // [AT_Q]
  int bid = blockIdx.x;
  const int kt = bid % ktiles; bid /= ktiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int k0 = kt * AT_KC;
  if (k0 >= T) return;
  const int col = h * hd;
  for (int i = threadIdx.x; i < AT_KC * hd; i += AT_THREADS) {
    const int r = i / hd, d = i - r * hd;
    const bool ok = k0 + r < T;
    ks[r * ldh + d] = ok ? p.k[(size_t)(lo + k0 + r) * p.ldk + col + d] : 0.f;
    vs[r * ldh + d] = ok ? p.v[(size_t)(lo + k0 + r) * p.ldv + col + d] : 0.f;
  }
  // thread t owns key j = t & 63 and half of the head columns: accumulators in registers
  // thread t owns key j = t & 63 and the head columns d = half + 2 e
  const int j = threadIdx.x & 63, half = threadIdx.x >> 6;
  float dk_acc[AT_MAXHD / 2], dv_acc[AT_MAXHD / 2];
#pragma unroll
  for (int e = 0; e < AT_MAXHD / 2; ++e) { dk_acc[e] = 0.f; dv_acc[e] = 0.f; }

  for (int q0 = 0; q0 < T; q0 += AT_Q) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_Q * hd; i += AT_THREADS) {
      const int r = i / hd, d = i - r * hd;
      const bool ok = q0 + r < T;
      qs[r * ldh + d] = ok ? p.q[(size_t)(lo + q0 + r) * p.ldq + col + d] * p.scale : 0.f;
      gs[r * ldh + d] = ok ? p.dout[(size_t)(lo + q0 + r) * p.lddo + col + d] : 0.f;
    }
    if (threadIdx.x < AT_Q) {
      const int r = threadIdx.x;
      float dsum = 0.f;
      if (q0 + r < T)
        for (int d = 0; d < hd; ++d)
          dsum += p.dout[(size_t)(lo + q0 + r) * p.lddo + col + d] * p.o[(size_t)(lo + q0 + r) * p.ldo + col + d];
      Di[r] = dsum;
    }
    __syncthreads();
    // P and dS for (16 queries x 64 keys): 1024 entries, 8 per thread
    for (int e = threadIdx.x; e < AT_Q * AT_KC; e += AT_THREADS) {
      const int r = e >> 6, jj = e & 63;
      float pv = 0.f, dsv = 0.f;
      if (q0 + r < T && k0 + jj < T) {
        float s = 0.f, dp = 0.f;
        for (int d = 0; d < hd; ++d) {
          s += qs[r * ldh + d] * ks[jj * ldh + d];
          dp += gs[r * ldh + d] * vs[jj * ldh + d];
        }
        pv = expf(s - p.lse[(size_t)(lo + q0 + r) * p.heads + h]);
        dsv = pv * (dp - Di[r]);
      }
      ps[e] = pv;
      ds[e] = dsv;
    }
    __syncthreads();
    for (int r = 0; r < AT_Q; ++r) {
      const float pv = ps[r * AT_KC + j], dsv = ds[r * AT_KC + j];
#pragma unroll
      for (int e = 0; e < AT_MAXHD / 2; ++e) {
        const int d = half + 2 * e;
        if (d < hd) {
          dv_acc[e] += pv * gs[r * ldh + d];
          dk_acc[e] += dsv * qs[r * ldh + d];  // qs already carries the scale
        }
      }
    }
  }
  if (k0 + j < T) {
#pragma unroll
    for (int e = 0; e < AT_MAXHD / 2; ++e) {
      const int d = half + 2 * e;
      if (d < hd) {
        p.dk[(size_t)(lo + k0 + j) * p.lddk + col + d] = dk_acc[e];
        p.dv[(size_t)(lo + k0 + j) * p.lddv + col + d] = dv_acc[e];
      }
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS) attn_bwd_q_kernel(const AttnDev p, int qtiles) {
  extern __shared__ float sm[];
  const int hd = p.hd, ldh = hd + 1;
  float* ks = sm;
  float* vs = ks + AT_KC * ldh;
  float* qs = vs + AT_KC * ldh;
  float* gs = qs + AT_Q * ldh;
  float* ds = gs + AT_Q * ldh;    // [AT_Q][64]
  float* Di = ds + AT_Q * AT_KC;  // [AT_Q]
  int bid = blockIdx.x;
  const int qt = bid % qtiles; bid /= qtiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int q0 = qt * AT_Q;
  if (q0 >= T) return;
  const int col = h * hd;
  for (int i = threadIdx.x; i < AT_Q * hd; i += AT_THREADS) {
    const int r = i / hd, d = i - r * hd;
    const bool ok = q0 + r < T;
    qs[r * ldh + d] = ok ? p.q[(size_t)(lo + q0 + r) * p.ldq + col + d] * p.scale : 0.f;
    gs[r * ldh + d] = ok ? p.dout[(size_t)(lo + q0 + r) * p.lddo + col + d] : 0.f;
  }
  if (threadIdx.x < AT_Q) {
    const int r = threadIdx.x;
    float dsum = 0.f;
    if (q0 + r < T)
      for (int d = 0; d < hd; ++d)
        dsum += p.dout[(size_t)(lo + q0 + r) * p.lddo + col + d] * p.o[(size_t)(lo + q0 + r) * p.ldo + col + d];
    Di[r] = dsum;
  }
  // thread t owns query r = t & 15 and an eighth of the head columns
  const int r_own = threadIdx.x & 15, part = threadIdx.x >> 4;  // columns d = part + 8 e
  float dq_acc[AT_MAXHD / 8];
#pragma unroll
  for (int e = 0; e < AT_MAXHD / 8; ++e) dq_acc[e] = 0.f;
  for (int k0 = 0; k0 < T; k0 += AT_KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_KC * hd; i += AT_THREADS) {
      const int r = i / hd, d = i - r * hd;
      const bool ok = k0 + r < T;
      ks[r * ldh + d] = ok ? p.k[(size_t)(lo + k0 + r) * p.ldk + col + d] : 0.f;
      vs[r * ldh + d] = ok ? p.v[(size_t)(lo + k0 + r) * p.ldv + col + d] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < AT_Q * AT_KC; e += AT_THREADS) {
      const int r = e >> 6, jj = e & 63;
      float dsv = 0.f;
      if (q0 + r < T && k0 + jj < T) {
        float s = 0.f, dp = 0.f;
        for (int d = 0; d < hd; ++d) {
          s += qs[r * ldh + d] * ks[jj * ldh + d];
          dp += gs[r * ldh + d] * vs[jj * ldh + d];
        }
        dsv = expf(s - p.lse[(size_t)(lo + q0 + r) * p.heads + h]) * (dp - Di[r]);
      }
      ds[e] = dsv;
    }
    __syncthreads();
    for (int jj = 0; jj < AT_KC; ++jj) {
      const float dsv = ds[r_own * AT_KC + jj];
#pragma unroll
      for (int e = 0; e < AT_MAXHD / 8; ++e) {
        const int d = part + 8 * e;
        if (d < hd) dq_acc[e] += dsv * ks[jj * ldh + d];
      }
    }
  }
  if (q0 + r_own < T) {
#pragma unroll
    for (int e = 0; e < AT_MAXHD / 8; ++e) {
      const int d = part + 8 * e;
      if (d < hd) p.dq[(size_t)(lo + q0 + r_own) * p.lddq + col + d] = dq_acc[e] * p.scale;
    }
  }
}

// ------------------------------------------------------------------------------------------------ dwconv3 + GELU
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * expf(-0.5f * u * u);
}

// y[t, c] = gelu( w[c,0] x[t-1,c] + w[c,1] x[t,c] + w[c,2] x[t+1,c] + b[c] ), zero outside the sequence
__global__ void dwconv_gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                                       const float* __restrict__ b, const BlkMeta* meta, int nrows, int C) {
  const long total = (long)nrows * (C / 4);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int row = (int)(i / (C / 4)), c = (int)(i - (long)row * (C / 4)) * 4;
    const BlkMeta m = meta[row / kBlkRows];
    if (row >= m.hi) continue;
    const float4 xc = *reinterpret_cast<const float4*>(x + (size_t)row * C + c);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 xm = (row - 1 >= m.lo) ? *reinterpret_cast<const float4*>(x + (size_t)(row - 1) * C + c) : z;
    const float4 xp = (row + 1 < m.hi) ? *reinterpret_cast<const float4*>(x + (size_t)(row + 1) * C + c) : z;
    const float xcv[4] = {xc.x, xc.y, xc.z, xc.w}, xmv[4] = {xm.x, xm.y, xm.z, xm.w}, xpv[4] = {xp.x, xp.y, xp.z, xp.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* wc = w + (size_t)(c + e) * 3;
      o[e] = gelu_f(wc[0] * xmv[e] + wc[1] * xcv[e] + wc[2] * xpv[e] + b[c + e]);
    }
    *reinterpret_cast<float4*>(y + (size_t)row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// du = dy * gelu'(u) (u recomputed);  dx[t] = w0 du[t+1] + w1 du[t] + w2 du[t-1];  dw[c,k] += sum_t du[t] x[t+k-1]
// pass A writes du, pass B computes dx; dw / db partial sums per CTA then atomics.
__global__ void dwconv_gelu_bwd_du_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                          float* __restrict__ du, const float* __restrict__ w,
                                          const float* __restrict__ b, float* __restrict__ dw, float* __restrict__ db,
                                          const BlkMeta* meta, int nrows, int C, int rows_per_cta) {
  // each CTA owns a range of rows and loops over channels with its threads (thread = channel): coalesced, and the
  // per-channel weight-gradient partial sums stay in registers
  const int r_begin = blockIdx.x * rows_per_cta, r_end = min(nrows, r_begin + rows_per_cta);
  for (int c = threadIdx.x + blockIdx.y * blockDim.x; c < C; c += blockDim.x * gridDim.y) {
    const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2], bb = b[c];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gb = 0.f;
    for (int row = r_begin; row < r_end; ++row) {
      const BlkMeta m = meta[row / kBlkRows];
      if (row >= m.hi) continue;
      const float xc = x[(size_t)row * C + c];
      const float xm = (row - 1 >= m.lo) ? x[(size_t)(row - 1) * C + c] : 0.f;
      const float xp = (row + 1 < m.hi) ? x[(size_t)(row + 1) * C + c] : 0.f;
      const float u = w0 * xm + w1 * xc + w2 * xp + bb;
      const float d = dy[(size_t)row * C + c] * gelu_grad_f(u);
      du[(size_t)row * C + c] = d;
      g0 += d * xm; g1 += d * xc; g2 += d * xp; gb += d;
    }
    atomicAdd(dw + c * 3, g0);
    atomicAdd(dw + c * 3 + 1, g1);
    atomicAdd(dw + c * 3 + 2, g2);
    atomicAdd(db + c, gb);
  }
}

__global__ void dwconv_bwd_dx_kernel(const float* __restrict__ du, float* __restrict__ dx, const float* __restrict__ w,
                                     const BlkMeta* meta, int nrows, int C) {
  const long total = (long)nrows * (C / 4);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int row = (int)(i / (C / 4)), c = (int)(i - (long)row * (C / 4)) * 4;
    const BlkMeta m = meta[row / kBlkRows];
    if (row >= m.hi) continue;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 dc = *reinterpret_cast<const float4*>(du + (size_t)row * C + c);
    const float4 dm = (row - 1 >= m.lo) ? *reinterpret_cast<const float4*>(du + (size_t)(row - 1) * C + c) : z;
    const float4 dp = (row + 1 < m.hi) ? *reinterpret_cast<const float4*>(du + (size_t)(row + 1) * C + c) : z;
    const float dcv[4] = {dc.x, dc.y, dc.z, dc.w}, dmv[4] = {dm.x, dm.y, dm.z, dm.w}, dpv[4] = {dp.x, dp.y, dp.z, dp.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float* wc = w + (size_t)(c + e) * 3;
      // u[t] uses x[t-1] w0, x[t] w1, x[t+1] w2  =>  dx[t] = du[t+1] w0 + du[t] w1 + du[t-1] w2
      o[e] = wc[0] * dpv[e] + wc[1] * dcv[e] + wc[2] * dmv[e];
    }
    *reinterpret_cast<float4*>(dx + (size_t)row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + b * y[i];
}

}  // namespace tcn

using namespace tcn;

static inline int cap_grid(long n, int per, int cap) {
  long b = (n + per - 1) / per;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

extern "C" int tcn_layernorm_fwd(const float* x, int ldx, float* y, int ldy, const float* gamma, const float* beta,
                                 float* mean, float* rstd, const int* meta, int nrows, int channels, float eps,
                                 tcn_stream_t stream) {
  TCN_REQUIRE(x && y && gamma && beta && mean && rstd && nrows > 0 && channels > 0, "tcn_layernorm_fwd: bad arguments");
  layernorm_fwd_kernel<<<cap_grid(nrows, 8, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, gamma, beta, mean, rstd, reinterpret_cast<const BlkMeta*>(meta), nrows, channels, eps);
  return check_launch("layernorm_fwd_kernel");
}

extern "C" int tcn_layernorm_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx,
                                 const float* gamma, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                                 const int* meta, int nrows, int channels, tcn_stream_t stream) {
  TCN_REQUIRE(x && dy && dx && gamma && mean && rstd && dgamma && dbeta && nrows > 0 && channels > 0,
              "tcn_layernorm_bwd: bad arguments");
  TCN_REQUIRE(channels <= 8192, "tcn_layernorm_bwd: too many channels");
  layernorm_bwd_kernel<<<cap_grid(nrows, 64, num_sms() * 2), 256, 2 * channels * sizeof(float), (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, dx, lddx, gamma, mean, rstd, dgamma, dbeta, reinterpret_cast<const BlkMeta*>(meta), nrows,
      channels);
  return check_launch("layernorm_bwd_kernel");
}

static int attn_common(const tcn_attn_args* a, AttnDev* p) {
  TCN_REQUIRE(a && a->q && a->k && a->v && a->o && a->lse && a->seq_lo && a->seq_len, "tcn_attn: null pointer");
  TCN_REQUIRE(a->nseq > 0 && a->heads > 0 && a->head_dim > 0 && a->max_len > 0, "tcn_attn: bad shape");
  if (a->head_dim > AT_MAXHD) {
    set_error("tcn_attn: head_dim %d > %d is not supported", a->head_dim, AT_MAXHD);
    return TCN_ERR_UNSUPPORTED;
  }
  p->q = a->q; p->ldq = a->ldq; p->k = a->k; p->ldk = a->ldk; p->v = a->v; p->ldv = a->ldv; p->o = a->o; p->ldo = a->ldo;
  p->lse = a->lse; p->dout = a->dout; p->lddo = a->lddo; p->dq = a->dq; p->lddq = a->lddq; p->dk = a->dk;
  p->lddk = a->lddk; p->dv = a->dv; p->lddv = a->lddv; p->seq_lo = a->seq_lo; p->seq_len = a->seq_len;
  p->heads = a->heads; p->hd = a->head_dim; p->scale = a->scale;
  return TCN_OK;
}

extern "C" int tcn_attn_fwd(const tcn_attn_args* a, tcn_stream_t stream) {
  AttnDev p;
  TCN_CHECK(attn_common(a, &p));
  const int qtiles = (a->max_len + AT_Q - 1) / AT_Q;
  const int ldh = a->head_dim + 1;
  const size_t smem = (size_t)(2 * AT_KC + AT_Q) * ldh * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_fwd_kernel<<<a->nseq * a->heads * qtiles, AT_THREADS, smem, (cudaStream_t)stream>>>(p, qtiles);
  return check_launch("attn_fwd_kernel");
}

extern "C" int tcn_attn_bwd(const tcn_attn_args* a, tcn_stream_t stream) {
  AttnDev p;
  TCN_CHECK(attn_common(a, &p));
  TCN_REQUIRE(a->dout && a->dq && a->dk && a->dv, "tcn_attn_bwd: null gradient pointer");
  const int qtiles = (a->max_len + AT_Q - 1) / AT_Q, ktiles = (a->max_len + AT_KC - 1) / AT_KC;
  const int ldh = a->head_dim + 1;
  const size_t smem_kv = (size_t)((2 * AT_KC + 2 * AT_Q) * ldh + 2 * AT_Q * AT_KC + AT_Q) * sizeof(float);
  const size_t smem_q = (size_t)((2 * AT_KC + 2 * AT_Q) * ldh + AT_Q * AT_KC + AT_Q) * sizeof(float);
  if (smem_kv > 48 * 1024) {
    cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv);
    cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
  }
  attn_bwd_kv_kernel<<<a->nseq * a->heads * ktiles, AT_THREADS, smem_kv, (cudaStream_t)stream>>>(p, ktiles);
  TCN_CHECK(check_launch("attn_bwd_kv_kernel"));
  attn_bwd_q_kernel<<<a->nseq * a->heads * qtiles, AT_THREADS, smem_q, (cudaStream_t)stream>>>(p, qtiles);
  return check_launch("attn_bwd_q_kernel");
}

extern "C" int tcn_dwconv_gelu_fwd(const float* x, float* y, const float* w, const float* b, const int* meta, int nrows,
                                   int channels, tcn_stream_t stream) {
  TCN_REQUIRE(x && y && w && b && meta && nrows > 0 && channels > 0 && channels % 4 == 0,
              "tcn_dwconv_gelu_fwd: bad arguments (channels must be a multiple of 4)");
  dwconv_gelu_fwd_kernel<<<cap_grid((long)nrows * channels / 4, 256, num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
      x, y, w, b, reinterpret_cast<const BlkMeta*>(meta), nrows, channels);
  return check_launch("dwconv_gelu_fwd_kernel");
}

extern "C" int tcn_dwconv_gelu_bwd(const float* x, const float* dy, float* du, float* dx, const float* w, const float* b,
                                   float* dw, float* db, const int* meta, int nrows, int channels,
                                   tcn_stream_t stream) {
  TCN_REQUIRE(x && dy && du && dx && w && b && dw && db && meta && nrows > 0 && channels > 0 && channels % 4 == 0,
              "tcn_dwconv_gelu_bwd: bad arguments");
  const int rows_per_cta = 64;
  dim3 grid((nrows + rows_per_cta - 1) / rows_per_cta, (channels + 1023) / 1024);
  dwconv_gelu_bwd_du_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, du, w, b, dw, db,
                                                                   reinterpret_cast<const BlkMeta*>(meta), nrows,
                                                                   channels, rows_per_cta);
  TCN_CHECK(check_launch("dwconv_gelu_bwd_du_kernel"));
  dwconv_bwd_dx_kernel<<<cap_grid((long)nrows * channels / 4, 256, num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
      du, dx, w, reinterpret_cast<const BlkMeta*>(meta), nrows, channels);
  return check_launch("dwconv_bwd_dx_kernel");
}

extern "C" int tcn_axpby(float* y, const float* x, float a, float b, long long n, tcn_stream_t stream) {
  TCN_REQUIRE(y && x && n > 0, "tcn_axpby: bad arguments");
  axpby_kernel<<<cap_grid(n, 1024, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(y, x, a, b, (long)n);
  return check_launch("axpby_kernel");
}
