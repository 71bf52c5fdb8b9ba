/* tcn_b200 -- C ABI of the B200-native temporal-head library (libtcn_b200.so).
 *
 * The reference (CIAM-Group/ComputerVision_Codes) has no FFI: its interface for this path is the
 * nn.Module constructor / forward / state_dict of
 *   MT4MTLKD/Temporal_tenco/network.py   (== TERL/0_5fold_TCN_black/network.py)
 *   MT4MTLKD/Temporal_mstct/MSTCT/*.py, MT4MTLKD/Temporal_mstct/network.py
 * and the loss arithmetic inlined in the run scripts.  Each entry point below names the reference
 * symbol (file:line, relative to the reference root) whose arithmetic it replaces; the Python
 * mirror in computervision_codes_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; nothing here allocates;
 *   - all functions return 0 on success and a negative TCN_ERR_* code otherwise, never throw, and
 *     leave a message for tcn_last_error() (thread local);
 *   - `stream` is a cudaStream_t; calls only enqueue work (no host synchronisation);
 *   - activations are TIME-MAJOR fp32: row = frame, columns = channels, row stride `ld*`;
 *     frames of a batch of sequences are packed along the row axis, each sequence starting at a
 *     multiple of 128 rows ("padded rows"); `meta` is an array of `nblk` int4 records, one per
 *     128-row block: {lo, hi, in_delta, seq} = valid row range of the owning sequence, the offset
 *     to add to a padded row to index caller-owned unpadded per-frame arrays, sequence index.
 */
#ifndef TCN_B200_H_
#define TCN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define TCN_OK 0
#define TCN_ERR_INVALID_ARG (-1)
#define TCN_ERR_UNSUPPORTED (-2)
#define TCN_ERR_CUDA (-3)

#define TCN_VERSION 100

typedef struct CUstream_st* tcn_stream_t;

/* ---- library ---------------------------------------------------------------------------------- */
int tcn_version(void);
const char* tcn_last_error(void);
/* Kernels this library has enqueued on a stream (or captured into a CUDA graph) since it was loaded; process wide.
 * bench.py's "gpu_launches" is the difference over one step x the steps timed. */
long long tcn_launch_count(void);
/* compute capability of the current device and the architecture the kernels were built for (100) */
int tcn_device_info(int* cc_major, int* cc_minor, int* num_sms, int* built_for_sm);

/* ---- weights ---------------------------------------------------------------------------------- */
/* Re-orders a torch Conv1d / Linear weight (n_out, c_in, ntaps) into the fragment-ordered, hi/lo
 * split buffer the tensor-core kernels read.  transpose=1 builds the operand of the input-gradient
 * pass (contraction over n_out).  Call once per optimizer step per weight. */
long long tcn_prep_weight_floats(int n_out, int c_in, int ntaps, int transpose);
int tcn_prep_weight(const float* w, int n_out, int c_in, int ntaps, int transpose, float* wf, tcn_stream_t stream);

/* ---- generic tap GEMM ---------------------------------------------------------------------------
 * y[r, n] = epi( sum_tap sum_c x[r + shift[tap], c] * W[n, c, tap] + bias[n] ), taps that leave the
 * sequence contribute zero.  epi = relu -> (* [relu_mask > 0]) -> dropout(p, seed, stream id) ->
 * (+ residual).  Replaces nn.Conv1d (k = 1 or 3, any dilation, zero padding / causal front padding)
 * and nn.Linear wherever the reference applies them per frame:
 *   network.py:113,129 (stage projection), :114-115,189-190 (layer convs), :74,98-106 (FPN lateral),
 *   :21-24,63-67 (heads); Temporal_mstct/MSTCT/Temporal_Encoder.py:12,15,57-59,139 ; TS_Mixer.py:10,39-48. */
typedef struct {
  const float* x; int ldx; int x_unpadded;      /* x_unpadded: rows of x are addressed with meta.in_delta */
  const float* colscale; int colscale_ld;       /* optional per-(sequence, input column) scale of x */
  const float* wf;                              /* from tcn_prep_weight */
  const float* bias;                            /* [n_out] or NULL */
  float* y; int ldy;
  const float* residual; int ldr;               /* NULL or (rows, n_out) added after dropout */
  const float* relu_mask; int ldm;              /* NULL or (rows, n_out): output zeroed where mask <= 0 */
  const int* meta; int nblk;
  int c_in; int n_out; int ntaps; int shift[3];
  int relu;
  float drop_p; unsigned drop_seed; unsigned drop_stream;   /* dropout on the output */
  float in_drop_p;                              /* dropout (same key) on x as it is loaded: backward of a dropout */
} tcn_tapgemm_args;
int tcn_tapgemm(const tcn_tapgemm_args* args, tcn_stream_t stream);

/* dw[n, c, tap] += sum_r g[r, n] * x[r + shift[tap], c] ;  db[n] += sum_r g[r, n]   (fp32 atomics).
 * The weight / bias gradients of the same layers. */
typedef struct {
  const float* g; int ldg; int g_cols;          /* g_cols: readable columns (multiple of 4, pads are zero) */
  const float* x; int ldx; int x_unpadded;
  const float* colscale; int colscale_ld;
  const int* meta; int nblk;
  int n_out; int c_in; int ntaps; int shift[3];
  float* dw; float* db;                         /* db may be NULL */
  float g_drop_p; unsigned drop_seed; unsigned drop_stream;   /* dropout on g as it is loaded */
} tcn_wgrad_args;
int tcn_wgrad(const tcn_wgrad_args* args, tcn_stream_t stream);

/* ---- tcgen05 + TMA tap GEMM ------------------------------------------------------------------------
 * Same contract as tcn_tapgemm, executed on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with a
 * 3-term operand split (fp32-level accuracy), accumulator in TMEM, operands fed by TMA through a 4-stage
 * shared-memory ring; a tap is the same TMA box fetched at a shifted row coordinate.
 *   y[r, n] = epi( sum_tap sum_c x[r + shift[tap], c] * W[n, c, tap] + bias[n] ),
 *   epi = relu -> (* [relu_mask > 0]) -> dropout(drop_p) -> (+ residual)
 * Used for the stage-input projection Conv1d(dim -> C, 1) of BaseCausalTCN (network.py:113,129) with
 * Dropout2d's channel scale (colscale, :125-127) and the 25 % input mask (in_drop_p, in_drop_rescale = 0,
 * :43-50) folded into the operand load; for the convolutions of the residual layers and their
 * input-gradient passes (network.py:178-198; in_drop_p with in_drop_rescale = 1 regenerates the
 * dropout mask on the incoming gradient); for the FPN lateral, the heads, and the MS-TCT Linear layers.
 * w_hi / w_lo come from tcn_split_weight (tcn_split_weight_floats each).  For an input-gradient pass
 * split with transpose = 1 and call with (c_in, n_out) exchanged.  x_rows = rows addressable behind x
 * (rows outside [0, x_rows) read as zero).  Needs c_in % 4 == 0 (tcn_gemm_tc_supported). */
typedef struct {
  const float* x; int ldx; long long x_rows; int x_unpadded;
  const float* w_hi; const float* w_lo; const float* bias;
  float* y; int ldy;
  const float* residual; int ldr;
  const float* relu_mask; int ldm;
  const int* meta; int nblk;
  int c_in; int n_out; int ntaps; int shift[3];
  int relu;
  const float* colscale; int colscale_ld;
  float in_drop_p; int in_drop_rescale;
  float drop_p; unsigned drop_seed; unsigned drop_stream;
} tcn_gemm_tc_args;
int tcn_gemm_tc_supported(int c_in, int n_out);
int tcn_gemm_tc(const tcn_gemm_tc_args* args, tcn_stream_t stream);
long long tcn_split_weight_floats(int n_out, int c_in, int ntaps, int transpose);
int tcn_split_weight(const float* w, int n_out, int c_in, int ntaps, int transpose, float* w_hi, float* w_lo,
                     tcn_stream_t stream);

/* tcgen05 + TMA weight gradient (same contract as tcn_wgrad): dw[n, c, tap] += sum_r g[r, n] x[r + shift[tap], c],
 * db[n] += sum_r g[r, n]; the reduction over frames runs on tcgen05.mma kind::tf32 with both operands MN-major
 * (TMA boxes of 32 frames x 32 columns), 3-term split of both operands in-kernel, accumulator in TMEM,
 * one fp32 atomic add per output element and CTA.  g_rows / x_rows = rows addressable behind g / x. */
typedef struct {
  const float* g; int ldg; int g_cols; long long g_rows;
  const float* x; int ldx; long long x_rows; int x_unpadded;
  const float* colscale; int colscale_ld;
  const int* meta; int nblk;
  int n_out; int c_in; int ntaps; int shift[3];
  float* dw; float* db;
  float g_drop_p; unsigned drop_seed; unsigned drop_stream;
} tcn_wgrad_tc_args;
int tcn_wgrad_tc(const tcn_wgrad_tc_args* args, tcn_stream_t stream);
/* The two weight gradients of one residual layer (gW1 / gb1 from (gu, x), gW2 / gb2 from (gy, h); network.py:186-198
 * backward) in ONE launch: the grid is split between the two problems.  Both must share meta / nblk, have
 * n_out <= 64 and padded operands.  This is the launch the executor issues per layer. */
int tcn_wgrad_tc_pair(const tcn_wgrad_tc_args* w1, const tcn_wgrad_tc_args* w2, tcn_stream_t stream);
/* ALL weight / bias gradients of one 64-channel residual layer in one pass over the frames, deterministic (what autograd
 * accumulates into conv_dilated.{weight,bias}.grad and conv_1x1.{weight,bias}.grad of network.py:186-198 / :165-183):
 *   dw1[n, c, k] += sum_t gu[t, n] x[t + shift[k], c]     db1[n] += sum_t gu[t, n]
 *   dw2[n, c]    += sum_t gv[t, n] h[t, c]                db2[n] += sum_t gv[t, n],   gv = keep * gy / (1 - drop_p)
 * (csrc/wgrad_layer.cu).  One CTA per frame range computes the four products, stores its partial to a slab of
 * `workspace` and a second launch adds the slabs in a fixed order: no atomics, bit-identical from run to run.
 * gu / x / gy / h: packed (rows, 64) fp32, 16-byte aligned.  masks: the (rows, 4) bit words of tcn_layer_fwd_tc (dropout
 * keep bits in words 2 and 3) or NULL, in which case the keep mask is regenerated from (drop_seed, drop_stream).
 * workspace: at least tcn_wgrad_layer_workspace_bytes(nblk) bytes of device memory, 16-byte aligned. */
typedef struct {
  const float* gu; const float* x; const float* gy; const float* h; long long rows;
  const unsigned* masks;
  const int* meta; int nblk; int channels;
  int shift[3];
  float drop_p; unsigned drop_seed; unsigned drop_stream;
  float* dw1; float* db1; float* dw2; float* db2;   /* db1 / db2 may be NULL */
  void* workspace; long long workspace_bytes;
  int flags;                                        /* bit 0: leave the slabs unreduced (timing the main kernel alone) */
} tcn_wgrad_layer_args;
long long tcn_wgrad_layer_workspace_bytes(int nblk);
int tcn_wgrad_layer(const tcn_wgrad_layer_args* args, tcn_stream_t stream);

/* ---- fused residual layer, forward ---------------------------------------------------------------
 * y = x + Dropout_p(W2 relu(W1 (*)_d x + b1) + b2), one launch: DilatedResidualLayer.forward
 * (network.py:193-198; shift = {-d, 0, +d}) and DilatedResidualCausalLayer.forward (network.py:178-183;
 * shift = {-2d, -d, 0}).  x, y, h are (rows, 64) time-major; h = relu(u) is written when non-NULL
 * (saved for the backward pass, which runs on tcn_tapgemm / tcn_wgrad).  channels must be 64
 * (TCN_ERR_UNSUPPORTED otherwise: other widths run as two tcn_tapgemm calls). */
typedef struct {
  const float* x; float* y; float* h;
  const float* w1f; const float* w2f;           /* tcn_prep_weight of (64,64,3) and (64,64,1) */
  const float* b1; const float* b2;
  const int* meta; int nblk; int channels;
  int shift[3];
  float drop_p; unsigned drop_seed; unsigned drop_stream;
} tcn_layer_fwd_args;
int tcn_layer_fwd(const tcn_layer_fwd_args* args, tcn_stream_t stream);
/* The same layer on tcgen05 / TMA / TMEM (csrc/gemm_tc.cu: layer_fwd_tc_kernel): the dilated conv accumulates in
 * tensor memory, bias + ReLU run on the accumulator, and h re-enters the 1x1 conv as a TMEM A operand.  Weights as
 * written by tcn_split_weight (transpose = 0) for (64, 64, 3) and (64, 64, 1); x_rows = rows of the x buffer. */
typedef struct {
  const float* x; long long x_rows; float* y; float* h;
  const float* w1_hi; const float* w1_lo; const float* w2_hi; const float* w2_lo;
  const float* b1; const float* b2;
  const int* meta; int nblk; int channels;
  int shift[3];
  float drop_p; unsigned drop_seed; unsigned drop_stream;
  unsigned* masks;   /* optional (rows, 4) bit words per frame for tcn_layer_bwd_tc: h > 0 [0..63], dropout keep [0..63] */
} tcn_layer_fwd_tc_args;
int tcn_layer_fwd_tc(const tcn_layer_fwd_tc_args* args, tcn_stream_t stream);
/* Input gradient of the same layer in one launch (what autograd computes for network.py:193-198 / :178-183):
 *   gu = ((keep * gy / (1 - p)) W2) * [h > 0]   (written for the weight gradients),
 *   gx[t] = gy[t] + sum_k W1[:, :, k]^T gu[t - shift[k]].
 * masks: the bit words written by tcn_layer_fwd_tc; w2t_* / w1t_*: tcn_split_weight with transpose = 1;
 * shift = the forward taps; drop_p = 0 in eval mode.  The weight gradients run on tcn_wgrad_tc (gy, h) / (gu, x). */
typedef struct {
  const float* gy; long long g_rows; float* gu; float* gx;
  const unsigned* masks;
  const float* w2t_hi; const float* w2t_lo; const float* w1t_hi; const float* w1t_lo;
  const int* meta; int nblk; int channels;
  int shift[3];
  float drop_p;
} tcn_layer_bwd_tc_args;
int tcn_layer_bwd_tc(const tcn_layer_bwd_tc_args* args, tcn_stream_t stream);

/* ---- whole-model executor -------------------------------------------------------------------------
 * VideoNas(fpn) of network.py:14-68 (BaseCausalTCN -> num_r x Refinement -> FPN -> 4 heads x 4 levels)
 * plus the train_loop loss of Temporal_tenco/run.py:190-212 (TERL run.py:307-343 with pos_weight),
 * forward + backward in one call, every launch on `stream`, graph-capturable: batch shape, block
 * table and dropout seed are read from device memory written by tcn_model_set_batch.
 * Parameters / gradients are two caller-owned flat fp32 buffers; tcn_model_param_layout gives the
 * float offset of each tensor in canonical order:
 *   PG.conv_1x1.{weight,bias}; for stage in PG, Rs.0.. : for layer: conv_dilated.{weight,bias},
 *   conv_1x1.{weight,bias}; fpn.latlayer1.{weight,bias}; conv_out, conv_out_i, conv_out_v,
 *   conv_out_t weights (contiguous), then their biases (contiguous). */
typedef struct {
  int layers_pg, layers_r, num_r, channels, in_dim;
  int head_sizes[4];                            /* ivt, i, v, t */
  int causal;
  int max_rows;                                 /* capacity in padded rows (multiple of 128) */
  int max_seqs;
} tcn_model_config;
typedef struct tcn_model tcn_model;
int tcn_model_create(const tcn_model_config* cfg, tcn_model** out);
void tcn_model_destroy(tcn_model* m);
long long tcn_model_num_params(const tcn_model* m);
int tcn_model_num_tensors(const tcn_model* m);
int tcn_model_param_layout(const tcn_model* m, long long* offsets, long long* sizes, int n);
/* Diagnostics: device pointer of a saved activation of the last forward, (max_rows, channels) fp32.
 * kind 0 = input of residual layer idx (idx = number of layers: the last output), kind 1 = relu output h of layer idx.
 * (No reference counterpart: the reference keeps these inside autograd.) */
int tcn_model_debug_ptr(tcn_model* m, int kind, int idx, void** out);
int tcn_model_bind(tcn_model* m, float* params, float* grads);
/* head_weights[4] for (ivt, i, v, t); pos_w: HOST array of sum(head_sizes) floats or NULL */
int tcn_model_set_loss(tcn_model* m, const float* head_weights, const float* pos_w_host);
/* input_mask_p: probability of zeroing an input element (network.py:43-48 uses 0.25), 0 = off;
 * chan_drop_p: Dropout2d over input channels (network.py:117), layer_drop_p: nn.Dropout of the layers */
int tcn_model_set_dropout(tcn_model* m, float input_mask_p, float chan_drop_p, float layer_drop_p);
/* Data parallel with unequal shares: average the loss (run.py:190-212 takes the mean over the frames of ONE video per step;
 * a batch averages over its videos) over `norm_seqs` sequences -- the GLOBAL number of videos of the step -- instead of
 * the sequences of this rank's batch, so that a plain sum all-reduce of the gradients gives the global mean.  0 restores
 * the default.  Takes effect with the next tcn_model_set_batch. */
int tcn_model_set_loss_norm(tcn_model* m, int norm_seqs);
/* meta_host: HOST int4[nblk]; copied (async) with the batch descriptor into device memory */
int tcn_model_set_batch(tcn_model* m, const int* meta_host, int nblk, int rows, int num_seqs, int frames,
                        unsigned seed, tcn_stream_t stream);
/* x: (x_rows >= frames, in_dim) fp32 unpadded, x_rows = rows addressable behind x; labels: (frames, ldlab) uint8 unpadded, columns ivt|i|v|t.
 * training != 0: dropout on, gradients accumulated into the bound grad buffer (zeroed first).
 * loss_out (device, 8 floats): [0..3] = per-head mean BCE summed over levels (ivt, i, v, t), [4] = total. */
int tcn_model_train_step(tcn_model* m, const float* x, long long x_rows, const unsigned char* labels, int ldlab,
                         int training, float* loss_out, tcn_stream_t stream);
/* forward only: fills pointers to the 4 FPN feature maps (rows, C) and the 4 logit maps (rows, ld_logits) owned
 * by the model (valid until the next call).  training: 0 = inference, 1 = train mode (dropout on, activations kept
 * for tcn_model_backward), 2 = eval-mode arithmetic but activations kept for tcn_model_backward */
int tcn_model_forward(tcn_model* m, const float* x, long long x_rows, int training, const float** feats,
                      const float** logits, int* ld_logits, tcn_stream_t stream);
/* backward from externally supplied gradients w.r.t. the 4 logit maps (rows, ld_logits; may be NULL)
 * and the 4 feature maps (rows, C; may be NULL), after tcn_model_forward(training=1) */
int tcn_model_backward(tcn_model* m, const float* x, long long x_rows, const float* const* glogits,
                       const float* const* gfeats, tcn_stream_t stream);

/* ---- losses ------------------------------------------------------------------------------------
 * Sigmoid-BCE over concatenated heads: nn.BCEWithLogitsLoss(pos_weight) as composed by
 * MT4MTLKD/Temporal_tenco/run.py:190-212, TERL/0_5fold_TCN_black/run.py:307-343,
 * MT4MTLKD/Spatial_cnn/run.py:159-162 and Temporal_mstct/run.py:185-196.
 * loss[h] += sum over rows and over the columns of head h of row_scale(r) * col_unit[c] * bce;
 * dl = d(sum_h head_weight_h * loss[h]) / d logits * grad_scale, with col_scale = head_weight * col_unit.
 * With meta: row_scale(r) = row_scale / T_seq (per-video mean, then mean over videos);
 * without: row_scale(r) = row_scale. */
typedef struct {
  const float* logits; int ldl;
  const unsigned char* labels; int ldlab; int lab_unpadded;
  const int* meta; int nrows; int ncols; int zero_cols;
  const float* pos_w; const float* col_scale; const float* col_unit; const int* col_head;
  float row_scale;
  float* loss;                                  /* [8] accumulators */
  float* dl; int lddl; float grad_scale;
} tcn_bce_args;
int tcn_bce_rows(const tcn_bce_args* args, tcn_stream_t stream);

/* DistillKL.forward (MT4MTLKD/Spatial_cnn/run.py:284-295; Spatial_transformer/run.py:290-300):
 * *loss += loss_scale * T^2/N * KL(softmax(t/T) || softmax(ys/T)),  t = sigmoid(yt) if teacher_sigmoid
 * (run.py:180-182);  gys = grad_scale * T * (softmax(ys/T) - softmax(t/T)) / N. */
int tcn_kd_kl_rows(const float* ys, int lds, const float* yt, int ldt, int teacher_sigmoid, int nrows, int K, float T,
                   float* loss, float loss_scale, float* gys, int ldg, float grad_scale, tcn_stream_t stream);

/* nn.MSELoss feature-KD (Spatial_cnn/run.py:187-191,328): *loss += loss_scale * mean((a-b)^2). */
int tcn_mse(const float* a, const float* b, long long n, float* loss, float loss_scale, float* ga, float grad_scale,
            tcn_stream_t stream);

/* Multi-teacher attention re-weighting of the student feature (MT4MTLKD/Spatial_cnn/network.py:47-71; same block
 * Spatial_transformer/network.py:102-124): with tea[n] = m_n(teacher_n) (N, F), S[b, n] = sum_d tea[n][b, d],
 * attn[b, c, :] = softmax_n(s[b, c] * S[b, n] / sqrt(F)) and z[n] = s * attn[:, :, n], the inputs of w_i / w_v / w_t.
 * Backward: gs (N, F) and gtea[n] (N, F; every column of row b carries dL/dS[b, n]).  tsum: (N, 3) scratch saved by
 * the forward. */
typedef struct {
  const float* s; int lds;
  const float* tea[3]; int ldt;
  float* z[3]; int ldz;
  float* tsum;
  const float* gz[3];
  float* gs; int ldgs;
  float* gtea[3];
  int n_rows; int feat_dim;
} tcn_kd_attn_args;
int tcn_kd_attn_fwd(const tcn_kd_attn_args* args, tcn_stream_t stream);
int tcn_kd_attn_bwd(const tcn_kd_attn_args* args, tcn_stream_t stream);

/* Softmax cross-entropy (7-way phase head; no reference counterpart). */
int tcn_ce_rows(const float* x, int ldx, const int* target, const int* meta, int tgt_unpadded, int nrows, int K,
                float row_scale, float* loss, float* gx, int ldg, float grad_scale, tcn_stream_t stream);

/* ---- MS-TCT blocks (MT4MTLKD/Temporal_mstct/MSTCT/Temporal_Encoder.py) -----------------------------
 * nn.LayerNorm over channels (:97,103,140,175-199): y = (x - mean) * rstd * gamma + beta, one warp per row;
 * mean / rstd (nrows floats each) are saved for the backward, which also accumulates dgamma / dbeta. */
int tcn_layernorm_fwd(const float* x, int ldx, float* y, int ldy, const float* gamma, const float* beta, float* mean,
                      float* rstd, const int* meta, int nrows, int channels, float eps, tcn_stream_t stream);
int tcn_layernorm_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, const float* gamma,
                      const float* mean, const float* rstd, float* dgamma, float* dbeta, const int* meta, int nrows,
                      int channels, tcn_stream_t stream);
/* Global_Relational_Block attention (:76-88): per (sequence, head) o = softmax(scale q k^T) v over the frames of
 * the sequence; head h lives at columns [h*head_dim, (h+1)*head_dim) of q / k / v / o (pass k and v already offset
 * into the kv projection).  lse: (rows, heads) log-sum-exp saved for the backward (two passes, no atomics).  All products run as
 * 3xTF32 mma.sync tiles (csrc/attention.cu). */
typedef struct {
  const float* q; int ldq; const float* k; int ldk; const float* v; int ldv;
  float* o; int ldo; float* lse;
  const float* dout; int lddo; float* dq; int lddq; float* dk; int lddk; float* dv; int lddv;
  const int* seq_lo; const int* seq_len; int nseq; int max_len;
  int heads; int head_dim; float scale;
  float* delta; /* backward scratch, (rows, heads): dO_i . O_i */
} tcn_attn_args;
int tcn_attn_fwd(const tcn_attn_args* args, tcn_stream_t stream);
int tcn_attn_bwd(const tcn_attn_args* args, tcn_stream_t stream);
/* The same attention on tcgen05 / TMEM / TMA for windows of at most 256 frames (csrc/attention_tc.cu): the six products
 * (S = q k^T, o = P v; dV = P^T dO, dP = dO v^T, dQ = dS k, dK = dS^T q) run on one batched tensor-core kernel, the scores of
 * a (window, head) are materialised once: p (nseq * heads * tmax, tmax) receives P = softmax(scale q k^T) in the forward
 * pass and is read by the backward pass; dp (same shape) is backward scratch.  tmax: a multiple of 32, >= the longest
 * window, <= 256; rows: rows of the q / k / v / o buffers.  head_dim and all leading dimensions multiples of 4. */
typedef struct {
  const float* q; int ldq; const float* k; int ldk; const float* v; int ldv;
  float* o; int ldo; float* p;
  const float* dout; int lddo; float* dq; int lddq; float* dk; int lddk; float* dv; int lddv; float* dp;
  const int* seq_lo; const int* seq_len; int nseq; long long rows;
  int heads; int head_dim; int tmax; float scale;
} tcn_attn_tc_args;
int tcn_attn_tc_supported(int max_len, int heads, int head_dim, int ldq, int ldkv);
int tcn_attn_fwd_tc(const tcn_attn_tc_args* args, tcn_stream_t stream);
int tcn_attn_bwd_tc(const tcn_attn_tc_args* args, tcn_stream_t stream);
/* Local_Relational_Block depthwise Conv1d(k=3, pad=1, groups=C) over time + GELU (:13-14,36-39), x / y (rows, C);
 * w: (C, 1, 3) torch layout.  Backward: du (scratch, rows x C), dx, dw += , db += . */
int tcn_dwconv_gelu_fwd(const float* x, float* y, const float* w, const float* b, const int* meta, int nrows,
                        int channels, tcn_stream_t stream);
int tcn_dwconv_gelu_bwd(const float* x, const float* dy, float* du, float* dx, const float* w, const float* b, float* dw,
                        float* db, const int* meta, int nrows, int channels, tcn_stream_t stream);
/* y = a * x + b * y (residual sums of the Temporal_Mixer, TS_Mixer.py:66-76) */
int tcn_axpby(float* y, const float* x, float a, float b, long long n, tcn_stream_t stream);

/* ---- evaluation ------------------------------------------------------------------------------------
 * Per-class average precision of ONE video on the device: what ivtmetrics.Recognition.compute_video_AP obtains from
 * sklearn.metrics.average_precision_score for the frames logged by mAP.update / video_end
 * (MT4MTLKD/Temporal_tenco/run.py:257-269, read at :428-450).  logits (nrows, ldl) fp32, labels (nrows, ldlab) uint8;
 * apply_sigmoid != 0 ranks sigmoid(logit) in fp32 like the reference's `activation`; ap: ncols floats, NaN for a
 * class without a positive frame.  ivtmetrics == 0.0.6 is not available here: parity with it is unpinned, the
 * arithmetic is checked against sklearn's average_precision_score. */
int tcn_ap_rows(const float* logits, int ldl, const unsigned char* labels, int ldlab, int nrows, int ncols,
                int apply_sigmoid, float* ap, tcn_stream_t stream);

/* ---- dropout / optimizer -----------------------------------------------------------------------
 * Counter-based dropout keyed by (seed, stream id, row, column): nn.Dropout() of the residual layers
 * (network.py:175,191) and Dropout2d over input channels (network.py:117,125-127). */
int tcn_dropout_apply(const float* x, int ldx, float* y, int ldy, int nrows, int ncols, float p, unsigned seed,
                      unsigned stream_id, tcn_stream_t stream);
int tcn_dropout_mask(unsigned char* keep, int nrows, int ncols, float p, unsigned seed, unsigned stream_id,
                     tcn_stream_t stream);
/* torch.optim.SGD(lr, weight_decay), no momentum (Temporal_tenco/run.py:345-346), over a flat buffer:
 * p -= lr * (grad_scale * g + weight_decay * p). */
int tcn_sgd_step(float* params, const float* grads, long long n, float lr, float weight_decay, float grad_scale,
                 tcn_stream_t stream);
/* Same update with hyper = {lr, weight_decay, grad_scale} read from DEVICE memory, so that a captured CUDA graph
 * follows the per-epoch schedule LinearLR(start_factor=power, warmups) -> ExponentialLR(decay_rate)
 * (Temporal_tenco/run.py:345-350, stepped at :235-236) without re-capture. */
int tcn_sgd_step_dev(float* params, const float* grads, long long n, const float* hyper, tcn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TCN_B200_H_ */
