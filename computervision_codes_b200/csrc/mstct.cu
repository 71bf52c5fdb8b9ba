// MS-TCT temporal blocks (MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py): the pieces that are not
// per-frame dense contractions (those run on gemm_tc / wgrad_tc):
//   layernorm_{fwd,bwd}      nn.LayerNorm over channels           (Temporal_Encoder.py:97,103,140,160,175-199)
//   (the Global_Relational_Block attention, :76-88, lives in attention.cu)
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
// Column sums: KMAX > 0 (C <= 32 KMAX): every lane keeps the partial sums of its columns c = lane + 32 k in registers over
// all rows of its warp, the eight warps of the CTA are added through shared memory once, then one atomic per column and
// CTA -- the first version added every element to shared memory with atomics (8 warps on the same addresses: 73 us per
// call at cfg3).  KMAX == 0: that version, for very wide rows.
template <int KMAX>
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
  constexpr int KR = KMAX > 0 ? KMAX : 1;
  float ag[KR], ab[KR];
#pragma unroll
  for (int k = 0; k < KR; ++k) { ag[k] = 0.f; ab[k] = 0.f; }
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    if (meta != nullptr && row >= meta[row / kBlkRows].hi) continue;
    const float* xr = x + (size_t)row * ldx;
    const float* gr = dy + (size_t)row * lddy;
    const float mu = mean[row], rs = rstd[row];
    float* dr = dx + (size_t)row * lddx;
    if (KMAX > 0) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
      for (int k = 0; k < KR; ++k) {
        const int c = lane + 32 * k;
        if (c < C) {
          const float xh = (xr[c] - mu) * rs, gg = gr[c] * __ldg(gamma + c);
          s1 += gg;
          s2 += gg * xh;
        }
      }
      s1 = warp_sum(s1) / (float)C;
      s2 = warp_sum(s2) / (float)C;
#pragma unroll
      for (int k = 0; k < KR; ++k) {   // second pass over the row (L1-resident): dx and the register column sums
        const int c = lane + 32 * k;
        if (c < C) {
          const float xh = (xr[c] - mu) * rs, g = gr[c];
          dr[c] = rs * (g * __ldg(gamma + c) - s1 - xh * s2);
          ag[k] += g * xh;
          ab[k] += g;
        }
      }
      continue;
    }
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs, gg = gr[c] * __ldg(gamma + c);
      s1 += gg;
      s2 += gg * xh;
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mu) * rs, g = gr[c];
      dr[c] = rs * (g * __ldg(gamma + c) - s1 - xh * s2);
      atomicAdd(&sg[c], g * xh);
      atomicAdd(&sb[c], g);
    }
  }
  if (KMAX > 0) {
#pragma unroll
    for (int k = 0; k < KR; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        atomicAdd(&sg[c], ag[k]);   // eight warps, once per CTA
        atomicAdd(&sb[c], ab[k]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (sg[c] != 0.f) atomicAdd(dgamma + c, sg[c]);
    if (sb[c] != 0.f) atomicAdd(dbeta + c, sb[c]);
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
    // walk the rows block by block (one block table entry per 128 rows); x[t-1], x[t], x[t+1] rotate through registers so
    // that a row costs one load of x and one of dy, issued four rows ahead of their use
    for (int blk0 = r_begin; blk0 < r_end; blk0 = (blk0 / kBlkRows + 1) * kBlkRows) {
      const BlkMeta m = meta[blk0 / kBlkRows];
      const int lo_r = blk0, hi_r = min(min(r_end, (blk0 / kBlkRows + 1) * kBlkRows), m.hi);
      if (lo_r >= hi_r) continue;
      float xm = (lo_r - 1 >= m.lo) ? x[(size_t)(lo_r - 1) * C + c] : 0.f;
      float xc = x[(size_t)lo_r * C + c];
      int row = lo_r;
      for (; row + 4 <= hi_r; row += 4) {
        float xn[4], dyv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xn[e] = (row + e + 1 < m.hi) ? x[(size_t)(row + e + 1) * C + c] : 0.f;
          dyv[e] = dy[(size_t)(row + e) * C + c];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xp = xn[e];
          const float u = w0 * xm + w1 * xc + w2 * xp + bb;
          const float d = dyv[e] * gelu_grad_f(u);
          du[(size_t)(row + e) * C + c] = d;
          g0 += d * xm; g1 += d * xc; g2 += d * xp; gb += d;
          xm = xc; xc = xp;
        }
      }
      for (; row < hi_r; ++row) {
        const float xp = (row + 1 < m.hi) ? x[(size_t)(row + 1) * C + c] : 0.f;
        const float u = w0 * xm + w1 * xc + w2 * xp + bb;
        const float d = dy[(size_t)row * C + c] * gelu_grad_f(u);
        du[(size_t)row * C + c] = d;
        g0 += d * xm; g1 += d * xc; g2 += d * xp; gb += d;
        xm = xc; xc = xp;
      }
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
  const dim3 lg(cap_grid(nrows, 64, num_sms() * 2));
  const size_t lsm = 2 * channels * sizeof(float);
  auto kern = channels <= 256 ? layernorm_bwd_kernel<8> : (channels <= 512 ? layernorm_bwd_kernel<16>
                              : (channels <= 1024 ? layernorm_bwd_kernel<32> : layernorm_bwd_kernel<0>));
  kern<<<lg, 256, lsm, (cudaStream_t)stream>>>(
      x, ldx, dy, lddy, dx, lddx, gamma, mean, rstd, dgamma, dbeta, reinterpret_cast<const BlkMeta*>(meta), nrows,
      channels);
  return check_launch("layernorm_bwd_kernel");
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
