// Per-frame heads' losses and their gradients: warp-per-row kernels with shuffle reductions.
//
//   bce_rows_kernel   : sigmoid-BCE with optional pos_weight over (rows x K) logits, several heads
//                       concatenated along K, each column carrying its own 1/K_head * head_weight;
//                       restates nn.BCEWithLogitsLoss as composed by
//                       MT4MTLKD/Temporal_tenco/run.py:190-212 and TERL/0_5fold_TCN_black/run.py:307-343.
//   kd_kl_rows_kernel : DistillKL (MT4MTLKD/Spatial_cnn/run.py:284-295) against sigmoid(teacher logits)
//                       (run.py:180-182): loss and closed-form gradient T*(softmax(s/T) - p_t)/N.
//   mse_kernel        : nn.MSELoss feature-KD (Spatial_cnn/run.py:187-191,328).
//   ce_rows_kernel    : softmax cross-entropy for the 7-way phase head (no reference counterpart).
#include <cstring>

#include "common.cuh"

namespace tcn {

constexpr int kMaxHeads = 8;

// loss[h] accumulates  sum_{r, c in head h} row_scale(r) * col_unit(c) * bce(r, c),  col_unit = 1 / K_head
// (so that loss[h] is the reference's mean-BCE of head h, averaged over the sequences of the batch);
// dL carries the full chain factor col_scale(c) (= head_weight / K_head) * row_scale(r) * grad_scale.
constexpr int kBceMaxIt = 8;   // column c = lane + 32 k, k < kBceMaxIt: up to 256 logits per row on the fast path
__global__ void __launch_bounds__(256) bce_rows_kernel(const BceDev p) {
  __shared__ float red[8][kMaxHeads];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) part[h] = 0.f;

  const int nrows = p.dyn ? p.dyn->rows : p.nrows;
  const float rsc = p.dyn ? 1.f / (float)(p.dyn->norm_seqs > 0 ? p.dyn->norm_seqs : p.dyn->num_seqs) : p.row_scale_const;
  const bool walk = p.nlev > 0 && p.cum_lv[0] != nullptr;   // one warp per row over all levels (gridDim.y == 1)
  const float* logits = p.nlev > 0 ? p.logits_lv[walk ? 0 : blockIdx.y] : p.logits;
  float* dL = p.nlev > 0 ? p.dL_lv[walk ? 0 : blockIdx.y] : p.dL;
  const bool fast = p.zero_cols <= 32 * kBceMaxIt;
  // fast path: a lane owns the same columns in every row, so the per-column constants live in registers and the loss is
  // summed per column; the head of each column is looked at once, after the row loop
  float acc[kBceMaxIt], cpw[kBceMaxIt], cscale[kBceMaxIt];
#pragma unroll
  for (int k = 0; k < kBceMaxIt; ++k) {
    const int c = lane + 32 * k;
    acc[k] = 0.f;
    cpw[k] = (fast && c < p.ncols && p.pos_w) ? __ldg(p.pos_w + c) : 1.f;
    cscale[k] = (fast && c < p.ncols) ? __ldg(p.col_scale + c) * p.grad_scale : 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    float rs = rsc;
    int lrow = row;
    if (p.meta != nullptr) {
      const BlkMeta m = p.meta[row / kBlkRows];
      if (row >= m.hi) continue;
      rs = rsc / (float)(m.hi - m.lo);
      if (p.lab_unpadded) lrow = row + m.in_delta;
    }
    const float* x = logits + (size_t)row * p.ldl;
    const uint8_t* y = p.labels + (size_t)lrow * p.ldlab;
    if (walk) {   // fast-path arithmetic for every level of this row; labels and per-column constants are read once
      float yv[kBceMaxIt], run[kBceMaxIt], xv[4][kBceMaxIt];
#pragma unroll
      for (int k = 0; k < kBceMaxIt; ++k) {
        const int c = lane + 32 * k;
        yv[k] = (c < p.ncols && y[c]) ? 1.f : 0.f;
        run[k] = 0.f;
      }
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {   // all loads of the row in flight before the arithmetic
        if (lv < p.nlev) {
          const float* xl = p.logits_lv[lv] + (size_t)row * p.ldl;
#pragma unroll
          for (int k = 0; k < kBceMaxIt; ++k) {
            const int c = lane + 32 * k;
            xv[lv][k] = c < p.ncols ? xl[c] : 0.f;
          }
        }
      }
#pragma unroll
      for (int lv = 0; lv < 4; ++lv) {
        if (lv < p.nlev) {
          float* dl = p.dL_lv[lv] + (size_t)row * p.lddl;
          float* cu = p.cum_lv[lv] + (size_t)row * p.lddl;
#pragma unroll
          for (int k = 0; k < kBceMaxIt; ++k) {
            const int c = lane + 32 * k;
            if (c < p.ncols) {
              const float xx = xv[lv][k], pw = cpw[k];
              const float e = __expf(-fabsf(xx));
              const float lw = 1.f + (pw - 1.f) * yv[k];
              const float sp = __logf(1.f + e) + fmaxf(-xx, 0.f);
              acc[k] += rs * ((1.f - yv[k]) * xx + lw * sp);
              const float inv = __fdividef(1.f, 1.f + e);
              const float sg = xx >= 0.f ? inv : e * inv;
              const float gr = (sg * (pw * yv[k] + 1.f - yv[k]) - pw * yv[k]) * rs * cscale[k];
              run[k] += gr;
              dl[c] = gr;
              cu[c] = run[k];
            } else if (c < p.zero_cols) {
              dl[c] = 0.f;
              cu[c] = 0.f;
            }
          }
        }
      }
      continue;
    }
    if (fast) {
#pragma unroll
      for (int k = 0; k < kBceMaxIt; ++k) {
        const int c = lane + 32 * k;
        if (c < p.ncols) {
          const float xv = x[c];
          const float yv = y[c] ? 1.f : 0.f;
          const float pw = cpw[k];
          // (1 - y) x + (1 + (pw - 1) y) (log1p(exp(-|x|)) + max(-x, 0))      [torch's stable form]
          // exp / log / divide on the SFU fast paths (ex2.approx, lg2.approx, rcp.approx: <= 2 ulp each); log(1 + e)
          // instead of log1p(e) costs <= 6e-8 absolute per term.  The libm versions made this kernel issue-bound (71 us).
          const float e = __expf(-fabsf(xv));
          const float lw = 1.f + (pw - 1.f) * yv;
          const float sp = __logf(1.f + e) + fmaxf(-xv, 0.f);
          acc[k] += rs * ((1.f - yv) * xv + lw * sp);
          if (dL != nullptr) {
            const float inv = __fdividef(1.f, 1.f + e);           // sigmoid(x) from the same exp(-|x|)
            const float sg = xv >= 0.f ? inv : e * inv;
            const float gr = sg * (pw * yv + 1.f - yv) - pw * yv;
            dL[(size_t)row * p.lddl + c] = gr * rs * cscale[k];
          }
        } else if (c < p.zero_cols && dL != nullptr) {
          dL[(size_t)row * p.lddl + c] = 0.f;
        }
      }
      continue;
    }
    for (int c = lane; c < p.zero_cols; c += 32) {
      if (c < p.ncols) {
        const float xv = x[c];
        const float yv = y[c] ? 1.f : 0.f;
        const float pw = p.pos_w ? __ldg(p.pos_w + c) : 1.f;
        const float lw = 1.f + (pw - 1.f) * yv;
        const float sp = log1pf(expf(-fabsf(xv))) + fmaxf(-xv, 0.f);
        const float l = (1.f - yv) * xv + lw * sp;
        const int h = __ldg(p.col_head + c);
        const float lu = rs * __ldg(p.col_unit + c) * l;
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) part[k] += (k == h) ? lu : 0.f;
        if (dL != nullptr) {
          const float sg = 1.f / (1.f + expf(-xv));
          const float gr = sg * (pw * yv + 1.f - yv) - pw * yv;
          dL[(size_t)row * p.lddl + c] = gr * rs * __ldg(p.col_scale + c) * p.grad_scale;
        }
      } else if (dL != nullptr) {
        dL[(size_t)row * p.lddl + c] = 0.f;
      }
    }
  }
  if (fast) {
#pragma unroll
    for (int k = 0; k < kBceMaxIt; ++k) {
      const int c = lane + 32 * k;
      if (c < p.ncols) {
        const int h = __ldg(p.col_head + c);
        const float lu = acc[k] * __ldg(p.col_unit + c);
#pragma unroll
        for (int j = 0; j < kMaxHeads; ++j) part[j] += (j == h) ? lu : 0.f;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
    const float v = warp_sum(part[h]);
    if (lane == 0) red[warp][h] = v;
  }
  __syncthreads();
  if (threadIdx.x < kMaxHeads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    if (v != 0.f) atomicAdd(p.loss + threadIdx.x, v);
  }
}

// ---------------------------------------------------------------------------------------- DistillKL
__global__ void __launch_bounds__(256) kd_kl_rows_kernel(const float* __restrict__ ys, int lds,
                                                         const float* __restrict__ yt, int ldt, int teacher_sigmoid,
                                                         int nrows, int K, float T, float* loss, float loss_scale,
                                                         float* gys, int ldg, float grad_scale) {
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invT = 1.f / T;
  float part = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    const float* s = ys + (size_t)row * lds;
    const float* tp = yt + (size_t)row * ldt;
    float ms = -INFINITY, mt = -INFINITY;
    for (int c = lane; c < K; c += 32) {
      const float tv = teacher_sigmoid ? 1.f / (1.f + expf(-tp[c])) : tp[c];
      ms = fmaxf(ms, s[c] * invT);
      mt = fmaxf(mt, tv * invT);
    }
    ms = warp_max(ms);
    mt = warp_max(mt);
    float zs = 0.f, zt = 0.f;
    for (int c = lane; c < K; c += 32) {
      const float tv = teacher_sigmoid ? 1.f / (1.f + expf(-tp[c])) : tp[c];
      zs += expf(s[c] * invT - ms);
      zt += expf(tv * invT - mt);
    }
    zs = warp_sum(zs);
    zt = warp_sum(zt);
    const float lzs = logf(zs), lzt = logf(zt);
    for (int c = lane; c < K; c += 32) {
      const float tv = teacher_sigmoid ? 1.f / (1.f + expf(-tp[c])) : tp[c];
      const float lps = s[c] * invT - ms - lzs;
      const float lpt = tv * invT - mt - lzt;
      const float pt = expf(lpt);
      part += pt * (lpt - lps);
      if (gys != nullptr) gys[(size_t)row * ldg + c] = T * (expf(lps) - pt) / (float)nrows * grad_scale;
    }
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    atomicAdd(loss, v * T * T / (float)nrows * loss_scale);
  }
}

// ---------------------------------------------------------------------------------------- MSE
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long n,
                                                  float* loss, float loss_scale, float* ga, float grad_scale) {
  __shared__ float red[8];
  float part = 0.f;
  const float inv = 1.f / (float)n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    part += d * d;
    if (ga != nullptr) ga[i] = 2.f * d * inv * grad_scale;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    atomicAdd(loss, v * inv * loss_scale);
  }
}

// ---------------------------------------------------------------------------------------- softmax CE
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ tgt,
                                                      const BlkMeta* meta, int tgt_unpadded, int nrows, int K,
                                                      float row_scale_const, float* loss, float* gx, int ldg,
                                                      float grad_scale) {
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    float rs = row_scale_const;
    int trow = row;
    if (meta != nullptr) {
      const BlkMeta m = meta[row / kBlkRows];
      if (row >= m.hi) continue;
      rs = row_scale_const / (float)(m.hi - m.lo);
      if (tgt_unpadded) trow = row + m.in_delta;
    }
    const float* xr = x + (size_t)row * ldx;
    float mx = -INFINITY;
    for (int c = lane; c < K; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float z = 0.f;
    for (int c = lane; c < K; c += 32) z += expf(xr[c] - mx);
    z = warp_sum(z);
    const float lz = logf(z);
    const int tg = tgt[trow];
    for (int c = lane; c < K; c += 32) {
      const float lp = xr[c] - mx - lz;
      if (c == tg) part += -lp * rs;
      if (gx != nullptr) gx[(size_t)row * ldg + c] = (expf(lp) - (c == tg ? 1.f : 0.f)) * rs * grad_scale;
    }
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    atomicAdd(loss, v);
  }
}

// ---------------------------------------------------------------------------------------- dropout helpers / SGD
__global__ void dropout_apply_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int nrows,
                                     int ncols, uint32_t thresh, float scale, uint32_t seed, uint32_t stream) {
  const long total = (long)nrows * ncols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ncols), col = (int)(i - (long)row * ncols);
    const float v = x[(size_t)row * ldx + col];
    y[(size_t)row * ldy + col] = (drop_hash(seed, stream, (uint32_t)row, (uint32_t)col) >= thresh) ? v * scale : 0.f;
  }
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, int nrows, int ncols, uint32_t thresh, uint32_t seed,
                                    uint32_t stream) {
  const long total = (long)nrows * ncols;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ncols), col = (int)(i - (long)row * ncols);
    keep[i] = (drop_hash(seed, stream, (uint32_t)row, (uint32_t)col) >= thresh) ? 1 : 0;
  }
}

// torch.optim.SGD(lr, weight_decay) without momentum (Temporal_tenco/run.py:345-346): p -= lr * (g + wd * p)
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long n, float lr, float wd,
                           float grad_scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float w = p[i];
    p[i] = w - lr * (g[i] * grad_scale + wd * w);
  }
}

// same update with the hyper-parameters in device memory {lr, weight_decay, grad_scale}: a captured CUDA graph then
// follows the learning-rate schedule (LinearLR warm-up + ExponentialLR, run.py:345-350) without re-capture
__global__ void sgd_dev_kernel(float* __restrict__ p, const float* __restrict__ g, long n,
                               const float* __restrict__ hyper) {
  const float lr = hyper[0], wd = hyper[1], grad_scale = hyper[2];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float w = p[i];
    p[i] = w - lr * (g[i] * grad_scale + wd * w);
  }
}


// ------------------------------------------------------------------------------------------------
// Multi-teacher attention re-weighting of the student feature (MT4MTLKD/Spatial_cnn/network.py:47-71).
// The reference's einsum over F stacked copies of s collapses to  logit[b, c, n] = s[b, c] * S[b, n] / sqrt(F)  with
// S[b, n] = sum_d m_n(t_n)[b, d]; attn = softmax over the three teachers; z_n = s * attn_n feeds w_n.
// One warp per frame.  tsum (N, 3) keeps S for the backward pass.
struct KdAttnDev {
  const float* s; int lds;
  const float* tea[3]; int ldt;
  float* z[3]; int ldz;
  float* tsum;
  const float* gz[3];
  float* gs; int ldgs;
  float* gtea[3];
  int N, F;
};

__device__ __forceinline__ void softmax3(float l0, float l1, float l2, float& a0, float& a1, float& a2) {
  const float m = fmaxf(l0, fmaxf(l1, l2));
  a0 = expf(l0 - m); a1 = expf(l1 - m); a2 = expf(l2 - m);
  const float inv = 1.f / (a0 + a1 + a2);
  a0 *= inv; a1 *= inv; a2 *= inv;
}

__global__ void __launch_bounds__(128) kd_attn_fwd_kernel(const KdAttnDev p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= p.N) return;
  float S[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    float acc = 0.f;
    for (int d = lane; d < p.F; d += 32) acc += p.tea[n][(size_t)b * p.ldt + d];
    S[n] = warp_sum(acc);
    if (lane == 0) p.tsum[b * 3 + n] = S[n];
  }
  const float scale = rsqrtf((float)p.F);
  for (int c = lane; c < p.F; c += 32) {
    const float sv = p.s[(size_t)b * p.lds + c];
    float a0, a1, a2;
    softmax3(sv * scale * S[0], sv * scale * S[1], sv * scale * S[2], a0, a1, a2);
    p.z[0][(size_t)b * p.ldz + c] = sv * a0;
    p.z[1][(size_t)b * p.ldz + c] = sv * a1;
    p.z[2][(size_t)b * p.ldz + c] = sv * a2;
  }
}

// gl_k = s a_k (gz_k - sum_n gz_n a_n);  gs = sum_n gz_n a_n + sum_k gl_k S_k / sqrt(F);
// gS_k = sum_c gl_k s / sqrt(F), broadcast to every column of the projected teacher (S is a plain row sum)
__global__ void __launch_bounds__(128) kd_attn_bwd_kernel(const KdAttnDev p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= p.N) return;
  const float S0 = p.tsum[b * 3], S1 = p.tsum[b * 3 + 1], S2 = p.tsum[b * 3 + 2];
  const float scale = rsqrtf((float)p.F);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int c = lane; c < p.F; c += 32) {
    const float sv = p.s[(size_t)b * p.lds + c];
    float a0, a1, a2;
    softmax3(sv * scale * S0, sv * scale * S1, sv * scale * S2, a0, a1, a2);
    const float z0 = p.gz[0][(size_t)b * p.ldz + c], z1 = p.gz[1][(size_t)b * p.ldz + c],
                z2 = p.gz[2][(size_t)b * p.ldz + c];
    const float dot = z0 * a0 + z1 * a1 + z2 * a2;
    const float l0 = sv * a0 * (z0 - dot), l1 = sv * a1 * (z1 - dot), l2 = sv * a2 * (z2 - dot);
    p.gs[(size_t)b * p.ldgs + c] = dot + (l0 * S0 + l1 * S1 + l2 * S2) * scale;
    g0 += l0 * sv; g1 += l1 * sv; g2 += l2 * sv;
  }
  g0 = warp_sum(g0) * scale; g1 = warp_sum(g1) * scale; g2 = warp_sum(g2) * scale;
  for (int d = lane; d < p.F; d += 32) {
    p.gtea[0][(size_t)b * p.ldt + d] = g0;
    p.gtea[1][(size_t)b * p.ldt + d] = g1;
    p.gtea[2][(size_t)b * p.ldt + d] = g2;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-class average precision of one video (the quantity ivtmetrics.Recognition.compute_video_AP takes from
// sklearn.metrics.average_precision_score; call sites Temporal_tenco/run.py:257-269,428-450):
//     AP_c = (1 / P_c) * sum over positives i of  #{j : y_j = 1, s_j >= s_i} / #{j : s_j >= s_i}
// which is sklearn's step-wise sum over distinct thresholds, ties included.  s = sigmoid(logit) in fp32 when
// apply_sigmoid (the reference ranks the fp32 sigmoid outputs, saturation ties included).  One CTA per class; scores
// and labels of the class staged in shared memory in chunks; classes without a positive frame give NaN.
constexpr int AP_THREADS = 256, AP_CHUNK = 4096;

__global__ void __launch_bounds__(AP_THREADS) ap_rows_kernel(const float* __restrict__ logits, int ldl,
                                                             const unsigned char* __restrict__ labels, int ldlab,
                                                             int nrows, int apply_sigmoid, float* __restrict__ ap) {
  __shared__ float ss[AP_CHUNK];
  __shared__ unsigned char sy[AP_CHUNK];
  __shared__ float red[AP_THREADS / 32];
  __shared__ int redp[AP_THREADS / 32];
  const int c = blockIdx.x;
  auto score = [&](int r) {
    const float x = logits[(size_t)r * ldl + c];
    return apply_sigmoid ? 1.f / (1.f + expf(-x)) : x;
  };
  float acc = 0.f;   // sum of the precisions at this thread's positives
  int npos = 0;
  for (int i0 = 0; i0 < nrows; i0 += AP_THREADS) {     // every thread owns one candidate frame per sweep
    const int i = i0 + threadIdx.x;
    const bool mine = i < nrows && labels[(size_t)i * ldlab + c] != 0;
    const float si = mine ? score(i) : 0.f;
    int cnt = 0, pos = 0;
    for (int j0 = 0; j0 < nrows; j0 += AP_CHUNK) {
      __syncthreads();
      for (int j = threadIdx.x; j < AP_CHUNK && j0 + j < nrows; j += AP_THREADS) {
        ss[j] = score(j0 + j);
        sy[j] = labels[(size_t)(j0 + j) * ldlab + c];
      }
      __syncthreads();
      if (mine) {
        const int n = min(AP_CHUNK, nrows - j0);
        for (int j = 0; j < n; ++j) {
          const bool ge = ss[j] >= si;
          cnt += ge;
          pos += ge && sy[j] != 0;
        }
      }
    }
    if (mine) {
      acc += (float)pos / (float)cnt;
      ++npos;
    }
  }
  acc = warp_sum(acc);
  for (int o = 16; o > 0; o >>= 1) npos += __shfl_xor_sync(0xffffffffu, npos, o);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redp[threadIdx.x >> 5] = npos; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    int n = 0;
    for (int w = 0; w < AP_THREADS / 32; ++w) { a += red[w]; n += redp[w]; }
    ap[c] = n > 0 ? a / (float)n : __int_as_float(0x7fc00000);
  }
}

int launch_bce(const BceDev& p, int cap_rows, cudaStream_t stream) {
  const long rows = cap_rows > 0 ? cap_rows : p.nrows;
  long b = (rows + 7) / 8;
  const long cap = (long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  int nlev = p.nlev > 0 ? p.nlev : 1;
  if (p.nlev > 0 && p.cum_lv[0] != nullptr) {
    if (p.zero_cols > 32 * kBceMaxIt || p.dL_lv[0] == nullptr) {
      set_error("bce: the level-walking mode needs <= %d columns and gradient buffers", 32 * kBceMaxIt);
      return TCN_ERR_INVALID_ARG;
    }
    nlev = 1;
  }
  if (nlev > 1 && b > cap / nlev) b = cap / nlev;
  bce_rows_kernel<<<dim3((int)b, nlev), 256, 0, stream>>>(p);
  return check_launch("bce_rows_kernel");
}

// keep / (1 - p) per (sequence, input channel): Dropout2d over whole input channels
// (MT4MTLKD/Temporal_tenco/network.py:117,125-127)
__global__ void chan_scale_kernel(float* __restrict__ out, int ld, int max_seqs, const BatchDesc* dyn, uint32_t thresh,
                                  float scale, uint32_t seed0, uint32_t stream_id) {
  const int nseq = dyn ? dyn->num_seqs : max_seqs;
  const uint32_t seed = seed0 ^ (dyn ? dyn->seed : 0u);
  const long total = (long)nseq * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int s = (int)(i / ld), d = (int)(i - (long)s * ld);
    out[i] = drop_factor(seed, stream_id, thresh, scale, s, d);
  }
}
int launch_chan_scale(float* out, int ld, int max_seqs, const BatchDesc* dyn, float p, uint32_t seed,
                      uint32_t stream_id, cudaStream_t stream) {
  long b = ((long)max_seqs * ld + 255) / 256;
  if (b > 1024) b = 1024;
  chan_scale_kernel<<<(int)b, 256, 0, stream>>>(out, ld, max_seqs, dyn, drop_thresh(p), 1.f / (1.f - p), seed,
                                               stream_id);
  return check_launch("chan_scale_kernel");
}

}  // namespace tcn

using namespace tcn;

static inline int grid_for(long n, int per_block, int cap) {
  long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

extern "C" int tcn_bce_rows(const tcn_bce_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->logits && a->labels && a->loss && a->col_scale && a->col_head && a->col_unit,
              "tcn_bce_rows: null pointer");
  TCN_REQUIRE(a->nrows > 0 && a->ncols > 0 && a->ldl >= a->ncols, "tcn_bce_rows: bad shape");
  TCN_REQUIRE(a->zero_cols >= a->ncols && (a->dl == nullptr || a->lddl >= a->zero_cols),
              "tcn_bce_rows: zero_cols must be in [ncols, lddl]");
  BceDev p;
  memset(&p, 0, sizeof(p));
  p.logits = a->logits; p.ldl = a->ldl; p.labels = a->labels; p.ldlab = a->ldlab; p.lab_unpadded = a->lab_unpadded;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nrows = a->nrows; p.ncols = a->ncols;
  p.zero_cols = a->zero_cols; p.pos_w = a->pos_w; p.col_scale = a->col_scale; p.col_head = a->col_head;
  p.row_scale_const = a->row_scale; p.loss = a->loss; p.dL = a->dl; p.lddl = a->lddl; p.grad_scale = a->grad_scale;
  p.col_unit = a->col_unit; p.dyn = nullptr;
  return launch_bce(p, 0, (cudaStream_t)stream);
}

extern "C" int tcn_kd_kl_rows(const float* ys, int lds, const float* yt, int ldt, int teacher_sigmoid, int nrows,
                              int K, float T, float* loss, float loss_scale, float* gys, int ldg, float grad_scale,
                              tcn_stream_t stream) {
  TCN_REQUIRE(ys && yt && loss && nrows > 0 && K > 0 && T > 0.f, "tcn_kd_kl_rows: bad arguments");
  kd_kl_rows_kernel<<<grid_for(nrows, 8, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      ys, lds, yt, ldt, teacher_sigmoid, nrows, K, T, loss, loss_scale, gys, ldg, grad_scale);
  return check_launch("kd_kl_rows_kernel");
}

extern "C" int tcn_mse(const float* a, const float* b, long long n, float* loss, float loss_scale, float* ga,
                       float grad_scale, tcn_stream_t stream) {
  TCN_REQUIRE(a && b && loss && n > 0, "tcn_mse: bad arguments");
  mse_kernel<<<grid_for(n, 1024, num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(a, b, (long)n, loss, loss_scale, ga,
                                                                                grad_scale);
  return check_launch("mse_kernel");
}

extern "C" int tcn_ce_rows(const float* x, int ldx, const int* target, const int* meta, int tgt_unpadded, int nrows,
                           int K, float row_scale, float* loss, float* gx, int ldg, float grad_scale,
                           tcn_stream_t stream) {
  TCN_REQUIRE(x && target && loss && nrows > 0 && K > 0, "tcn_ce_rows: bad arguments");
  ce_rows_kernel<<<grid_for(nrows, 8, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, target, reinterpret_cast<const BlkMeta*>(meta), tgt_unpadded, nrows, K, row_scale, loss, gx, ldg,
      grad_scale);
  return check_launch("ce_rows_kernel");
}

extern "C" int tcn_dropout_apply(const float* x, int ldx, float* y, int ldy, int nrows, int ncols, float p,
                                 unsigned seed, unsigned stream_id, tcn_stream_t stream) {
  TCN_REQUIRE(x && y && nrows > 0 && ncols > 0 && p > 0.f && p < 1.f, "tcn_dropout_apply: bad arguments");
  dropout_apply_kernel<<<grid_for((long)nrows * ncols, 1024, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, nrows, ncols, drop_thresh(p), 1.f / (1.f - p), seed, stream_id);
  return check_launch("dropout_apply_kernel");
}

extern "C" int tcn_dropout_mask(unsigned char* keep, int nrows, int ncols, float p, unsigned seed, unsigned stream_id,
                                tcn_stream_t stream) {
  TCN_REQUIRE(keep && nrows > 0 && ncols > 0 && p > 0.f && p < 1.f, "tcn_dropout_mask: bad arguments");
  dropout_mask_kernel<<<grid_for((long)nrows * ncols, 1024, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      keep, nrows, ncols, drop_thresh(p), seed, stream_id);
  return check_launch("dropout_mask_kernel");
}

extern "C" int tcn_sgd_step(float* params, const float* grads, long long n, float lr, float weight_decay,
                            float grad_scale, tcn_stream_t stream) {
  TCN_REQUIRE(params && grads && n > 0, "tcn_sgd_step: bad arguments");
  sgd_kernel<<<grid_for(n, 1024, num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(params, grads, (long)n, lr,
                                                                                weight_decay, grad_scale);
  return check_launch("sgd_kernel");
}

extern "C" int tcn_sgd_step_dev(float* params, const float* grads, long long n, const float* hyper,
                                tcn_stream_t stream) {
  TCN_REQUIRE(params && grads && hyper && n > 0, "tcn_sgd_step_dev: bad arguments");
  sgd_dev_kernel<<<grid_for(n, 1024, num_sms() * 4), 256, 0, (cudaStream_t)stream>>>(params, grads, (long)n, hyper);
  return check_launch("sgd_dev_kernel");
}

static int kd_attn_common(const tcn_kd_attn_args* a, KdAttnDev* p) {
  TCN_REQUIRE(a && a->s && a->tsum && a->n_rows > 0 && a->feat_dim > 0, "tcn_kd_attn: bad arguments");
  for (int n = 0; n < 3; ++n) TCN_REQUIRE(a->tea[n] && a->z[n], "tcn_kd_attn: null teacher / output pointer");
  TCN_REQUIRE(a->lds >= a->feat_dim && a->ldt >= a->feat_dim && a->ldz >= a->feat_dim, "tcn_kd_attn: bad pitch");
  p->s = a->s; p->lds = a->lds; p->ldt = a->ldt; p->ldz = a->ldz; p->tsum = a->tsum; p->gs = a->gs; p->ldgs = a->ldgs;
  for (int n = 0; n < 3; ++n) { p->tea[n] = a->tea[n]; p->z[n] = a->z[n]; p->gz[n] = a->gz[n]; p->gtea[n] = a->gtea[n]; }
  p->N = a->n_rows; p->F = a->feat_dim;
  return TCN_OK;
}

extern "C" int tcn_kd_attn_fwd(const tcn_kd_attn_args* a, tcn_stream_t stream) {
  KdAttnDev p;
  TCN_CHECK(kd_attn_common(a, &p));
  kd_attn_fwd_kernel<<<(a->n_rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("kd_attn_fwd_kernel");
}

extern "C" int tcn_kd_attn_bwd(const tcn_kd_attn_args* a, tcn_stream_t stream) {
  KdAttnDev p;
  TCN_CHECK(kd_attn_common(a, &p));
  TCN_REQUIRE(a->gs && a->ldgs >= a->feat_dim, "tcn_kd_attn_bwd: null gradient pointer");
  for (int n = 0; n < 3; ++n) TCN_REQUIRE(a->gz[n] && a->gtea[n], "tcn_kd_attn_bwd: null gradient pointer");
  kd_attn_bwd_kernel<<<(a->n_rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("kd_attn_bwd_kernel");
}

extern "C" int tcn_ap_rows(const float* logits, int ldl, const unsigned char* labels, int ldlab, int nrows, int ncols,
                           int apply_sigmoid, float* ap, tcn_stream_t stream) {
  TCN_REQUIRE(logits && labels && ap && nrows > 0 && ncols > 0 && ldl >= ncols && ldlab >= ncols,
              "tcn_ap_rows: bad arguments");
  ap_rows_kernel<<<ncols, AP_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, labels, ldlab, nrows, apply_sigmoid, ap);
  return check_launch("ap_rows_kernel");
}
