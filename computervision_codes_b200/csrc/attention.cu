// Global_Relational_Block attention (MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py:76-88) on the tensor
// cores: per (window, head) o = softmax(scale q k^T) v and its backward, all four / five products as
// mma.sync m16n8k8 TF32 with the 3-term hi/lo split (fp32-grade results, see common.cuh).
//
// One CTA = (window, head, 64 rows); warp w owns rows [16 w, 16 w + 16) of the tile and streams the other
// side of the product through shared memory in chunks.  Every tile is row-major [rows][ld], ld = hdp + 4 with
// hdp = head_dim rounded up to 8 (zero filled), which makes both access patterns bank-conflict free:
//   * "nk"  C += A_rows * B^T with B stored [n][k]  (scores: q k^T, dO v^T)         bank = 4 g' + t
//   * "pk"  C += P * B with B stored [k][n] and P taken straight from a C fragment   (P v, dS k, P^T dO, dS^T q):
//     the k index of one k-step is permuted (slot t <-> row 2t, slot t+4 <-> row 2t+1) so that the C fragment
//     {d0,d2,d1,d3} IS the A fragment -- no shuffles -- and the B rows 2t / 2t+1 land in distinct banks.
// Backward is two passes (dK,dV per key tile; dQ per query tile), no atomics; D_i = dO_i . O_i is computed once
// into the caller's `delta` scratch.
#include "common.cuh"

namespace tcn {

constexpr int ATT_ROWS = 64, ATT_THREADS = 128, ATT_MAXHD = 128;

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
  float* delta;
  const int* seq_lo;  // [nseq] first row of each window
  const int* seq_len; // [nseq]
  int heads, hd, max_len;
  float scale;
  int vec;  // all pointers 16 B aligned, all leading dimensions and hd multiples of 4
};

// dst[r][d] = mul * src[r * lds + d] for r < nvalid, d < hd; zero elsewhere (rows up to nrows, columns up to hdp)
__device__ __forceinline__ void load_tile(float* dst, int ld, const float* __restrict__ src, int lds, int nrows,
                                          int nvalid, int hd, int hdp, float mul, bool vec) {
  if (vec) {
    const int n4 = hdp >> 2;
    for (int i = threadIdx.x; i < nrows * n4; i += ATT_THREADS) {
      const int r = i / n4, d = (i - r * n4) << 2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nvalid && d < hd) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * lds + d));
      v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
      *reinterpret_cast<float4*>(dst + r * ld + d) = v;
    }
  } else {
    for (int i = threadIdx.x; i < nrows * hdp; i += ATT_THREADS) {
      const int r = i / hdp, d = i - r * hdp;
      dst[r * ld + d] = (r < nvalid && d < hd) ? __ldg(src + (size_t)r * lds + d) * mul : 0.f;
    }
  }
}

// acc[nt] (16 x 8 each) += A(16 rows at As, k = 8 ksteps columns) * B^T, B stored [n][k] at Bs
template <int NT>
__device__ __forceinline__ void gemm_nk(float (&acc)[1][NT][4], const float* __restrict__ As,
                                        const float* __restrict__ Bs, int ld, int ksteps, int g, int t) {
  const float* a0p = As + g * ld + t;
  const float* b0p = Bs + g * ld + t;
#pragma unroll 2
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k0 = ks * 8;
    uint32_t ahi[1][4], alo[1][4], bhi[NT][2], blo[NT][2];
    split_tf32(a0p[k0], ahi[0][0], alo[0][0]);
    split_tf32(a0p[8 * ld + k0], ahi[0][1], alo[0][1]);
    split_tf32(a0p[k0 + 4], ahi[0][2], alo[0][2]);
    split_tf32(a0p[8 * ld + k0 + 4], ahi[0][3], alo[0][3]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      split_tf32(b0p[nt * 8 * ld + k0], bhi[nt][0], blo[nt][0]);
      split_tf32(b0p[nt * 8 * ld + k0 + 4], bhi[nt][1], blo[nt][1]);
    }
    mma_block_3xtf32<1, NT>(acc, ahi, alo, bhi, blo);
  }
}

// acc[nt] += P(16 x 8 NTC, a C fragment) * B, B stored [k][n] at Bs (k = the P columns), n-tiles < ntd only
template <int NTC, int NTD>
__device__ __forceinline__ void gemm_pk(float (&acc)[NTD][4], const float (&P)[1][NTC][4],
                                        const float* __restrict__ Bs, int ld, int ntd, int g, int t) {
  constexpr int GN = NTD <= 8 ? NTD : NTD / 2;
#pragma unroll
  for (int j = 0; j < NTC; ++j) {
    uint32_t ahi[4], alo[4];
    split_tf32(P[0][j][0], ahi[0], alo[0]);
    split_tf32(P[0][j][2], ahi[1], alo[1]);
    split_tf32(P[0][j][1], ahi[2], alo[2]);
    split_tf32(P[0][j][3], ahi[3], alo[3]);
    const float* b0p = Bs + (j * 8 + 2 * t) * ld + g;
#pragma unroll
    for (int n0 = 0; n0 < NTD; n0 += GN) {
      uint32_t bhi[GN][2], blo[GN][2];
#pragma unroll
      for (int n = 0; n < GN; ++n)
        if (n0 + n < ntd) {
          split_tf32(b0p[(n0 + n) * 8], bhi[n][0], blo[n][0]);
          split_tf32(b0p[ld + (n0 + n) * 8], bhi[n][1], blo[n][1]);
        }
#pragma unroll
      for (int n = 0; n < GN; ++n)
        if (n0 + n < ntd) mma_tf32(acc[n0 + n], alo, bhi[n][0], bhi[n][1]);
#pragma unroll
      for (int n = 0; n < GN; ++n)
        if (n0 + n < ntd) mma_tf32(acc[n0 + n], ahi, blo[n][0], blo[n][1]);
#pragma unroll
      for (int n = 0; n < GN; ++n)
        if (n0 + n < ntd) mma_tf32(acc[n0 + n], ahi, bhi[n][0], bhi[n][1]);
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// rows g / g+8 of a 16 x (8 NTD) accumulator -> global, columns < hd
template <int NTD>
__device__ __forceinline__ void store_rows(const float (&acc)[NTD][4], float* dst, int ldd, int row0, int T, int hd,
                                           float mul0, float mul1, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) {
    const int c = nt * 8 + 2 * t;
    if (c < hd) {
      if (row0 + g < T) dst[(size_t)(row0 + g) * ldd + c] = acc[nt][0] * mul0;
      if (row0 + g + 8 < T) dst[(size_t)(row0 + g + 8) * ldd + c] = acc[nt][2] * mul1;
    }
    if (c + 1 < hd) {
      if (row0 + g < T) dst[(size_t)(row0 + g) * ldd + c + 1] = acc[nt][1] * mul0;
      if (row0 + g + 8 < T) dst[(size_t)(row0 + g + 8) * ldd + c + 1] = acc[nt][3] * mul1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// CTA = (window, head, 64 queries); keys / values in chunks of 64 with an online softmax.
// lse[row * heads + h] = log sum_j exp(scale q_i.k_j)
template <int NTD>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnDev p, int qtiles) {
  extern __shared__ __align__(16) float sm[];
  const int hd = p.hd, hdp = (hd + 7) & ~7, ld = hdp + 4, ntd = hdp >> 3;
  float* qs = sm;
  float* ks = qs + ATT_ROWS * ld;
  float* vs = ks + ATT_ROWS * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int bid = blockIdx.x;
  const int qt = bid % qtiles; bid /= qtiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int q0 = qt * ATT_ROWS;
  if (q0 >= T) return;
  const int col = h * hd;
  load_tile(qs, ld, p.q + (size_t)(lo + q0) * p.ldq + col, p.ldq, ATT_ROWS, T - q0, hd, hdp, p.scale, p.vec);
  float O[NTD][4];
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) O[nt][0] = O[nt][1] = O[nt][2] = O[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const float* qw = qs + warp * 16 * ld;
  const bool active = q0 + warp * 16 < T;
  for (int kc0 = 0; kc0 < T; kc0 += ATT_ROWS) {
    __syncthreads();
    load_tile(ks, ld, p.k + (size_t)(lo + kc0) * p.ldk + col, p.ldk, ATT_ROWS, T - kc0, hd, hdp, 1.f, p.vec);
    load_tile(vs, ld, p.v + (size_t)(lo + kc0) * p.ldv + col, p.ldv, ATT_ROWS, T - kc0, hd, hdp, 1.f, p.vec);
    __syncthreads();
    if (!active) continue;
    float S[1][8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) S[0][nt][0] = S[0][nt][1] = S[0][nt][2] = S[0][nt][3] = 0.f;
    gemm_nk<8>(S, qw, ks, ld, ntd, g, t);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = kc0 + nt * 8 + 2 * t;
      if (c >= T) S[0][nt][0] = S[0][nt][2] = -INFINITY;
      if (c + 1 >= T) S[0][nt][1] = S[0][nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(S[0][nt][0], S[0][nt][1]));
      mx1 = fmaxf(mx1, fmaxf(S[0][nt][2], S[0][nt][3]));
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));  // finite: key kc0 is valid
    const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      S[0][nt][0] = expf(S[0][nt][0] - mn0);
      S[0][nt][1] = expf(S[0][nt][1] - mn0);
      S[0][nt][2] = expf(S[0][nt][2] - mn1);
      S[0][nt][3] = expf(S[0][nt][3] - mn1);
      s0 += S[0][nt][0] + S[0][nt][1];
      s1 += S[0][nt][2] + S[0][nt][3];
    }
    l0 = l0 * c0 + quad_sum(s0);
    l1 = l1 * c1 + quad_sum(s1);
    m0 = mn0; m1 = mn1;
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) {
      O[nt][0] *= c0; O[nt][1] *= c0; O[nt][2] *= c1; O[nt][3] *= c1;
    }
    gemm_pk<8, NTD>(O, S, vs, ld, ntd, g, t);
  }
  if (!active) return;
  const int r0 = q0 + warp * 16;
  store_rows<NTD>(O, p.o + (size_t)lo * p.ldo + col, p.ldo, r0, T, hd, 1.f / l0, 1.f / l1, g, t);
  if (t == 0) {
    if (r0 + g < T) p.lse[(size_t)(lo + r0 + g) * p.heads + h] = m0 + logf(l0);
    if (r0 + g + 8 < T) p.lse[(size_t)(lo + r0 + g + 8) * p.heads + h] = m1 + logf(l1);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// P_ij = exp(scale q_i.k_j - lse_i);  D_i = dO_i . O_i;  dV_j = sum_i P_ij dO_i;  dS_ij = P_ij (dO_i.v_j - D_i);
// dQ_i = scale sum_j dS_ij k_j;  dK_j = scale sum_i dS_ij q_i
__global__ void __launch_bounds__(128) attn_delta_kernel(const AttnDev p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.y;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int r = blockIdx.x * 4 + warp;
  if (r >= T) return;
  const float* go = p.dout + (size_t)(lo + r) * p.lddo;
  const float* o = p.o + (size_t)(lo + r) * p.ldo;
  for (int h = 0; h < p.heads; ++h) {
    float s = 0.f;
    for (int d = lane; d < p.hd; d += 32) s += go[h * p.hd + d] * o[h * p.hd + d];
    s = warp_sum(s);
    if (lane == 0) p.delta[(size_t)(lo + r) * p.heads + h] = s;
  }
}

// pass 1: CTA = (window, head, 64 keys), warp = 16 keys; queries in chunks of 8 NTC.  Works on the transposed
// score tile S^T = k q^T so that the keys are the accumulator rows.
template <int NTD, int NTC>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kv_kernel(const AttnDev p, int ktiles) {
  constexpr int CQ = NTC * 8;
  extern __shared__ __align__(16) float sm[];
  const int hd = p.hd, hdp = (hd + 7) & ~7, ld = hdp + 4, ntd = hdp >> 3;
  float* ks = sm;
  float* vs = ks + ATT_ROWS * ld;
  float* qs = vs + ATT_ROWS * ld;   // [CQ][ld] scaled q
  float* gs = qs + CQ * ld;         // [CQ][ld] dO
  float* lse_s = gs + CQ * ld;      // [CQ]
  float* d_s = lse_s + CQ;          // [CQ]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int bid = blockIdx.x;
  const int kt = bid % ktiles; bid /= ktiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int k0 = kt * ATT_ROWS;
  if (k0 >= T) return;
  const int col = h * hd;
  load_tile(ks, ld, p.k + (size_t)(lo + k0) * p.ldk + col, p.ldk, ATT_ROWS, T - k0, hd, hdp, 1.f, p.vec);
  load_tile(vs, ld, p.v + (size_t)(lo + k0) * p.ldv + col, p.ldv, ATT_ROWS, T - k0, hd, hdp, 1.f, p.vec);
  float dK[NTD][4], dV[NTD][4];
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) {
    dK[nt][0] = dK[nt][1] = dK[nt][2] = dK[nt][3] = 0.f;
    dV[nt][0] = dV[nt][1] = dV[nt][2] = dV[nt][3] = 0.f;
  }
  const int r0 = k0 + warp * 16;
  const bool active = r0 < T;
  const bool ok0 = r0 + g < T, ok1 = r0 + g + 8 < T;
  const float* kw = ks + warp * 16 * ld;
  const float* vw = vs + warp * 16 * ld;
  for (int q0 = 0; q0 < T; q0 += CQ) {
    __syncthreads();
    load_tile(qs, ld, p.q + (size_t)(lo + q0) * p.ldq + col, p.ldq, CQ, T - q0, hd, hdp, p.scale, p.vec);
    load_tile(gs, ld, p.dout + (size_t)(lo + q0) * p.lddo + col, p.lddo, CQ, T - q0, hd, hdp, 1.f, p.vec);
    if (threadIdx.x < CQ) {
      const int r = threadIdx.x;
      const bool ok = q0 + r < T;
      lse_s[r] = ok ? p.lse[(size_t)(lo + q0 + r) * p.heads + h] : INFINITY;   // exp(s - inf) = 0
      d_s[r] = ok ? p.delta[(size_t)(lo + q0 + r) * p.heads + h] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    float S[1][NTC][4], dP[1][NTC][4];
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) S[0][nt][0] = S[0][nt][1] = S[0][nt][2] = S[0][nt][3] = 0.f;
    gemm_nk<NTC>(S, kw, qs, ld, ntd, g, t);
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const float2 l2 = *reinterpret_cast<const float2*>(lse_s + nt * 8 + 2 * t);
      S[0][nt][0] = ok0 ? expf(S[0][nt][0] - l2.x) : 0.f;
      S[0][nt][1] = ok0 ? expf(S[0][nt][1] - l2.y) : 0.f;
      S[0][nt][2] = ok1 ? expf(S[0][nt][2] - l2.x) : 0.f;
      S[0][nt][3] = ok1 ? expf(S[0][nt][3] - l2.y) : 0.f;
    }
    gemm_pk<NTC, NTD>(dV, S, gs, ld, ntd, g, t);
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) dP[0][nt][0] = dP[0][nt][1] = dP[0][nt][2] = dP[0][nt][3] = 0.f;
    gemm_nk<NTC>(dP, vw, gs, ld, ntd, g, t);
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const float2 d2 = *reinterpret_cast<const float2*>(d_s + nt * 8 + 2 * t);
      dP[0][nt][0] = S[0][nt][0] * (dP[0][nt][0] - d2.x);
      dP[0][nt][1] = S[0][nt][1] * (dP[0][nt][1] - d2.y);
      dP[0][nt][2] = S[0][nt][2] * (dP[0][nt][2] - d2.x);
      dP[0][nt][3] = S[0][nt][3] * (dP[0][nt][3] - d2.y);
    }
    gemm_pk<NTC, NTD>(dK, dP, qs, ld, ntd, g, t);   // qs carries the scale
  }
  if (!active) return;
  store_rows<NTD>(dK, p.dk + (size_t)lo * p.lddk + col, p.lddk, r0, T, hd, 1.f, 1.f, g, t);
  store_rows<NTD>(dV, p.dv + (size_t)lo * p.lddv + col, p.lddv, r0, T, hd, 1.f, 1.f, g, t);
}

// pass 2: CTA = (window, head, 64 queries), warp = 16 queries; keys in chunks of 8 NTC
template <int NTD, int NTC>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_q_kernel(const AttnDev p, int qtiles) {
  constexpr int CK = NTC * 8;
  extern __shared__ __align__(16) float sm[];
  const int hd = p.hd, hdp = (hd + 7) & ~7, ld = hdp + 4, ntd = hdp >> 3;
  float* qs = sm;                    // [64][ld] scaled q
  float* gs = qs + ATT_ROWS * ld;    // [64][ld] dO
  float* ks = gs + ATT_ROWS * ld;    // [CK][ld]
  float* vs = ks + CK * ld;          // [CK][ld]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int bid = blockIdx.x;
  const int qt = bid % qtiles; bid /= qtiles;
  const int h = bid % p.heads;
  const int seq = bid / p.heads;
  const int lo = p.seq_lo[seq], T = p.seq_len[seq];
  const int q0 = qt * ATT_ROWS;
  if (q0 >= T) return;
  const int col = h * hd;
  load_tile(qs, ld, p.q + (size_t)(lo + q0) * p.ldq + col, p.ldq, ATT_ROWS, T - q0, hd, hdp, p.scale, p.vec);
  load_tile(gs, ld, p.dout + (size_t)(lo + q0) * p.lddo + col, p.lddo, ATT_ROWS, T - q0, hd, hdp, 1.f, p.vec);
  const int r0 = q0 + warp * 16;
  const bool active = r0 < T;
  const bool ok0 = r0 + g < T, ok1 = r0 + g + 8 < T;
  const float lse0 = ok0 ? p.lse[(size_t)(lo + r0 + g) * p.heads + h] : INFINITY;
  const float lse1 = ok1 ? p.lse[(size_t)(lo + r0 + g + 8) * p.heads + h] : INFINITY;
  const float D0 = ok0 ? p.delta[(size_t)(lo + r0 + g) * p.heads + h] : 0.f;
  const float D1 = ok1 ? p.delta[(size_t)(lo + r0 + g + 8) * p.heads + h] : 0.f;
  float dQ[NTD][4];
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) dQ[nt][0] = dQ[nt][1] = dQ[nt][2] = dQ[nt][3] = 0.f;
  const float* qw = qs + warp * 16 * ld;
  const float* gw = gs + warp * 16 * ld;
  for (int kc0 = 0; kc0 < T; kc0 += CK) {
    __syncthreads();
    load_tile(ks, ld, p.k + (size_t)(lo + kc0) * p.ldk + col, p.ldk, CK, T - kc0, hd, hdp, 1.f, p.vec);
    load_tile(vs, ld, p.v + (size_t)(lo + kc0) * p.ldv + col, p.ldv, CK, T - kc0, hd, hdp, 1.f, p.vec);
    __syncthreads();
    if (!active) continue;
    float S[1][NTC][4], dP[1][NTC][4];
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      S[0][nt][0] = S[0][nt][1] = S[0][nt][2] = S[0][nt][3] = 0.f;
      dP[0][nt][0] = dP[0][nt][1] = dP[0][nt][2] = dP[0][nt][3] = 0.f;
    }
    gemm_nk<NTC>(S, qw, ks, ld, ntd, g, t);
    gemm_nk<NTC>(dP, gw, vs, ld, ntd, g, t);
#pragma unroll
    for (int nt = 0; nt < NTC; ++nt) {
      const int c = kc0 + nt * 8 + 2 * t;
      const bool c0 = c < T, c1 = c + 1 < T;
      S[0][nt][0] = c0 ? expf(S[0][nt][0] - lse0) * (dP[0][nt][0] - D0) : 0.f;
      S[0][nt][1] = c1 ? expf(S[0][nt][1] - lse0) * (dP[0][nt][1] - D0) : 0.f;
      S[0][nt][2] = c0 ? expf(S[0][nt][2] - lse1) * (dP[0][nt][2] - D1) : 0.f;
      S[0][nt][3] = c1 ? expf(S[0][nt][3] - lse1) * (dP[0][nt][3] - D1) : 0.f;
    }
    gemm_pk<NTC, NTD>(dQ, S, ks, ld, ntd, g, t);
  }
  if (!active) return;
  store_rows<NTD>(dQ, p.dq + (size_t)lo * p.lddq + col, p.lddq, r0, T, hd, p.scale, p.scale, g, t);
}

}  // namespace tcn

using namespace tcn;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int attn_common(const tcn_attn_args* a, AttnDev* p) {
  TCN_REQUIRE(a && a->q && a->k && a->v && a->o && a->lse && a->seq_lo && a->seq_len, "tcn_attn: null pointer");
  TCN_REQUIRE(a->nseq > 0 && a->heads > 0 && a->head_dim > 0 && a->max_len > 0, "tcn_attn: bad shape");
  if (a->head_dim > ATT_MAXHD) {
    set_error("tcn_attn: head_dim %d > %d is not supported", a->head_dim, ATT_MAXHD);
    return TCN_ERR_UNSUPPORTED;
  }
  p->q = a->q; p->ldq = a->ldq; p->k = a->k; p->ldk = a->ldk; p->v = a->v; p->ldv = a->ldv; p->o = a->o; p->ldo = a->ldo;
  p->lse = a->lse; p->dout = a->dout; p->lddo = a->lddo; p->dq = a->dq; p->lddq = a->lddq; p->dk = a->dk;
  p->lddk = a->lddk; p->dv = a->dv; p->lddv = a->lddv; p->delta = a->delta; p->seq_lo = a->seq_lo;
  p->seq_len = a->seq_len; p->heads = a->heads; p->hd = a->head_dim; p->max_len = a->max_len; p->scale = a->scale;
  bool vec = a->head_dim % 4 == 0 && a->ldq % 4 == 0 && a->ldk % 4 == 0 && a->ldv % 4 == 0 && aligned16(a->q) &&
             aligned16(a->k) && aligned16(a->v);
  if (a->dout) vec = vec && a->lddo % 4 == 0 && aligned16(a->dout);
  p->vec = vec ? 1 : 0;
  return TCN_OK;
}

template <typename K>
static void set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// head-dim buckets (in 8-column tiles); the backward kernels take 64-row chunks while the accumulators are small
#define ATT_DISPATCH(ntd, X) \
  do {                       \
    if (ntd <= 4) { X(4, 8); } else if (ntd <= 8) { X(8, 8); } else if (ntd <= 10) { X(10, 4); } \
    else if (ntd <= 14) { X(14, 4); } else { X(16, 4); }                                          \
  } while (0)

extern "C" int tcn_attn_fwd(const tcn_attn_args* a, tcn_stream_t stream) {
  AttnDev p;
  TCN_CHECK(attn_common(a, &p));
  const int qtiles = (a->max_len + ATT_ROWS - 1) / ATT_ROWS;
  const int hdp = (a->head_dim + 7) & ~7, ld = hdp + 4, ntd = hdp / 8;
  const size_t smem = (size_t)3 * ATT_ROWS * ld * sizeof(float);
  const int grid = a->nseq * a->heads * qtiles;
#define X(NTD, NTC)                                                                        \
  set_smem(attn_fwd_kernel<NTD>, smem);                                                    \
  attn_fwd_kernel<NTD><<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(p, qtiles)
  ATT_DISPATCH(ntd, X);
#undef X
  return check_launch("attn_fwd_kernel");
}

extern "C" int tcn_attn_bwd(const tcn_attn_args* a, tcn_stream_t stream) {
  AttnDev p;
  TCN_CHECK(attn_common(a, &p));
  TCN_REQUIRE(a->dout && a->dq && a->dk && a->dv && a->delta, "tcn_attn_bwd: null gradient / scratch pointer");
  const int tiles = (a->max_len + ATT_ROWS - 1) / ATT_ROWS;
  const int hdp = (a->head_dim + 7) & ~7, ld = hdp + 4, ntd = hdp / 8;
  const int grid = a->nseq * a->heads * tiles;
  attn_delta_kernel<<<dim3((a->max_len + 3) / 4, a->nseq), 128, 0, (cudaStream_t)stream>>>(p);
  TCN_CHECK(check_launch("attn_delta_kernel"));
#define X(NTD, NTC)                                                                                      \
  {                                                                                                      \
    const size_t smem_kv = ((size_t)(2 * ATT_ROWS + 2 * NTC * 8) * ld + 2 * NTC * 8) * sizeof(float);    \
    const size_t smem_q = (size_t)(2 * ATT_ROWS + 2 * NTC * 8) * ld * sizeof(float);                     \
    set_smem(attn_bwd_kv_kernel<NTD, NTC>, smem_kv);                                                     \
    set_smem(attn_bwd_q_kernel<NTD, NTC>, smem_q);                                                       \
    attn_bwd_kv_kernel<NTD, NTC><<<grid, ATT_THREADS, smem_kv, (cudaStream_t)stream>>>(p, tiles);        \
    attn_bwd_q_kernel<NTD, NTC><<<grid, ATT_THREADS, smem_q, (cudaStream_t)stream>>>(p, tiles);          \
  }
  ATT_DISPATCH(ntd, X);
#undef X
  return check_launch("attn_bwd kernels");
}
