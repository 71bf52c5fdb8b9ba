// Whole-model executor: VideoNas(fpn) forward + multi-head BCE loss + backward as one sequence of
// kernel launches on one stream, no host synchronisation, CUDA-graph capturable.
//
// Mirrors VideoNas.forward (MT4MTLKD/Temporal_tenco/network.py:36-68): BaseCausalTCN (:120-135) ->
// num_r x Refinement (:149-162, use_output = hier = False as in every reference script) -> FPN
// (:98-106, latlayer1 three times, interpolate == identity) -> conv_out / _i / _v / _t on the four
// levels (:63-67), and the loss of train_loop (Temporal_tenco/run.py:190-212; TERL variant
// TERL/0_5fold_TCN_black/run.py:307-343).  Backward follows the closed forms of SURVEY.md section 8a.
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <utility>
#include <vector>

#include "gemm_tc.cuh"
#include "wgrad_layer.cuh"

namespace tcn {

struct TensorSlot {
  long off, size;
};

static inline int rup(int a, int b) { return (a + b - 1) / b * b; }
static inline long rupl(long a, long b) { return (a + b - 1) / b * b; }

}  // namespace tcn

using namespace tcn;

struct tcn_model {
  tcn_model_config cfg;
  int L = 0;       // residual layers in total
  int C = 0, D = 0, NH = 0, LDH = 0, NLVL = 4;
  std::vector<int> stage_first;  // first global layer index of each stage (+ sentinel)
  std::vector<int> dilation;     // per global layer
  // flat parameter layout
  std::vector<TensorSlot> slots;
  long off_proj_w = 0, off_proj_b = 0, off_lat_w = 0, off_lat_b = 0, off_head_w = 0, off_head_b = 0;
  std::vector<long> off_w1, off_b1, off_w2, off_b2;
  long n_params = 0;
  float* params = nullptr;
  float* grads = nullptr;
  // device workspace
  char* ws = nullptr;
  size_t ws_bytes = 0;
  // prepared weights (float offsets into wf)
  float* wf = nullptr;
  long wf_floats = 0;
  long wf_proj = 0, wf_lat = 0, wf_latT = 0, wf_head = 0, wf_headT = 0;
  std::vector<long> wf_w1, wf_w2, wf_w1T, wf_w2T;
  PrepJob* jobs_dev = nullptr;
  int njobs = 0;
  long prep_total_f4 = 0;
  // activations
  std::vector<float*> act, H;
  // FPN level stacks: [f0 | f1 | f2] and [p1 | p2 | p3 | p4 = f3] are contiguous (max_rows each), as are the four logit /
  // dLogits / cumulative-dLogits / level-gradient maps, so that the heads, their weight and input gradients and the
  // lateral weight gradient are ONE launch each over a 4x (3x) block table (meta4, desc4 / desc3)
  float* cum[4] = {nullptr, nullptr, nullptr, nullptr};
  BatchDesc *desc4 = nullptr, *desc3 = nullptr;
  BlkMeta* meta4 = nullptr;
  int norm_seqs = 0;          // tcn_model_set_loss_norm: sequences the loss is averaged over (0 = those of the batch)
  bool stack_levels = true;   // TCN_NO_STACK=1: one launch per level (A/B)
  // slabs of the deterministic weight-gradient reduction of the heads, the lateral and the projection (wgrad_tc.cu)
  float* slab[3] = {nullptr, nullptr, nullptr};
  int slab_cap[3] = {0, 0, 0};
  bool det_wgrad = true;      // TCN_WGRAD_ATOMIC=1: fp32 atomics instead (A/B)
  float* P[3] = {nullptr, nullptr, nullptr};
  float* logits[4] = {nullptr, nullptr, nullptr, nullptr};
  float* dL[4] = {nullptr, nullptr, nullptr, nullptr};
  float* Gp[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<float*> gpool;  // one gradient buffer per residual layer (+ 3 lateral adds): no reuse inside a step, so
  std::vector<float*> gus;    // the weight-gradient kernels can run on a second stream behind the input-gradient chain
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> evs;
  // batched weight gradients: one launch per stage over all its residual layers (built lazily, per training flag)
  struct WgMulti {
    bool ready = false;
    WgradMultiDesc* descs = nullptr;
    int2* tiles = nullptr;
    std::vector<int> tile_first, tile_count, row_splits;   // per stage
  } wgm[2];
  // per-layer weight gradients (wgrad_layer.cu): descriptor tables in backward order, built lazily per training flag
  struct WlTable {
    bool ready = false;
    WgLayerDev* descs = nullptr;
    WgLayerOut* outs = nullptr;
  } wl[2];
  float* wl_part = nullptr;   // [L][wl_splits][WL_PART_FLOATS] partial slabs
  int wl_lg = 1, wl_splits = 1;
  bool use_wl = true;         // TCN_WGRAD_PAIR=1: round 1's pair kernel (fp32 atomics) instead (A/B)
  bool wg_multi = false;  // TCN_WGRAD_MULTI=1: one launch per stage (measured slower than per-layer pair launches:
                          // 3.12 vs 3.09 ms / step -- the long per-CTA frame loops lose more than the launches save)
  bool overlap_wgrad = true;
  float* colscale = nullptr;
  // tcgen05 path: split (hi / lo) copies of every weight + their TMA maps, keyed by the float offset of the
  // fragment-ordered copy the mma.sync kernels use; TMA maps of the activation buffers keyed by (pointer, columns)
  struct TcW {
    long off = 0, rows = 0, cols = 0;
    CUtensorMap mh, ml;
  };
  bool use_tc = false;
  bool fused_tc = true;   // TCN_NO_FUSED_TC=1: previous forward schedule (A/B)
  bool fused_bwd = true;  // TCN_NO_FUSED_BWD=1: two input-gradient launches per layer (A/B)
  std::vector<uint32_t*> masks;  // per layer (rows, 4) bit words: ReLU / dropout masks saved by the fused forward
  std::map<long, TcW> tcw;
  std::map<std::pair<const float*, int>, CUtensorMap> xmaps;
  std::map<std::pair<const float*, int>, CUtensorMap> xmaps32;  // same tensors, 32-row boxes (slab kernel)
  std::map<std::pair<const float*, long>, CUtensorMap> wgmaps;  // 32-frame boxes, 32-byte swizzle atoms (wgrad_tc)
  float *tc_whi = nullptr, *tc_wlo = nullptr;
  long tc_wfloats = 0;
  SplitJob* sjobs_dev = nullptr;
  int nsjobs = 0;
  long split_total = 0;
  // tcgen05 projection: split weight halves + their TMA maps (fixed addresses), X map cached per pointer
  bool proj_tc = false;
  float *proj_whi = nullptr, *proj_wlo = nullptr;
  CUtensorMap map_whi, map_wlo, map_x;
  const float* map_x_ptr = nullptr;
  long map_x_rows = 0;
  BatchDesc* desc = nullptr;
  BlkMeta* meta = nullptr;
  static constexpr int kSlots = 4;  // ring of pinned staging slots: desc followed by the block table
  char* desc_host = nullptr;
  cudaEvent_t slot_done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  int slot = 0;
  size_t slot_bytes = 0;
  int max_blk = 0;
  // loss tables
  float *col_unit = nullptr, *col_scale = nullptr, *pos_w = nullptr, *loss8 = nullptr;
  int* col_head = nullptr;
  bool has_pos_w = false;
  float head_w[4] = {1.f, 0.1f, 0.1f, 0.1f};
  float input_mask_p = 0.f, chan_drop_p = 0.5f, layer_drop_p = 0.5f;
  bool fwd_training = false;

  const float* p_(long off) const { return params + off; }
  float* g_(long off) const { return grads + off; }
  const float4* wf_(long off) const { return reinterpret_cast<const float4*>(wf + off); }
};

namespace {

constexpr uint32_t kStreamChan = 0x7fff0001u;  // dropout stream ids that cannot collide with layer indices
constexpr uint32_t kStreamMask = 0x7fff0002u;

long prep_floats(int n_out, int c_in, int ntaps, int transpose) {
  return (long)tcn_prep_weight_floats(n_out, c_in, ntaps, transpose);
}

void add_job(std::vector<PrepJob>& jobs, long& first_f4, long src_off, int n_out, int c_in, int ntaps, int transpose) {
  PrepJob j;
  const int kdim = transpose ? n_out : c_in, ncols = transpose ? c_in : n_out;
  j.first = first_f4;
  j.src_off = src_off;
  j.n_out = n_out; j.c_in = c_in; j.ntaps = ntaps; j.transpose = transpose;
  j.kpt = rup(kdim, 8);
  j.NT8 = (ncols + 7) / 8;
  first_f4 += (long)ntaps * j.kpt / 8 * j.NT8 * 32;
  jobs.push_back(j);
}

void layer_shifts(const tcn_model* m, int l, int* s) {
  const int d = m->dilation[l];
  if (m->cfg.causal) { s[0] = -2 * d; s[1] = -d; s[2] = 0; } else { s[0] = -d; s[1] = 0; s[2] = d; }
}

TapGemmDev base_tapgemm(const tcn_model* m) {
  TapGemmDev p;
  memset(&p, 0, sizeof(p));
  p.meta = m->meta; p.nblk = m->max_blk; p.dyn = m->desc;
  p.ntaps = 1; p.drop_scale = 1.f; p.in_drop_scale = 1.f;
  return p;
}

WgradDev base_wgrad(const tcn_model* m) {
  WgradDev p;
  memset(&p, 0, sizeof(p));
  p.meta = m->meta; p.nblk = m->max_blk; p.dyn = m->desc;
  p.ntaps = 1; p.g_drop_scale = 1.f; p.x_drop_scale = 1.f;
  return p;
}

// levels > 1: X / Y / R are level stacks (levels * max_rows rows) and p.meta / p.dyn the stacked block table
int gemm(tcn_model* m, TapGemmDev& p, int c_in, int n_out, cudaStream_t st, int levels = 1) {
  const int cap_blk = levels * m->max_blk;
  const long map_rows = (long)levels * m->cfg.max_rows;
  if (m->use_tc && !p.x_unpadded) {
    const long key = reinterpret_cast<const float*>(p.Wf) - m->wf;
    auto it = m->tcw.find(key);
    if (it != m->tcw.end()) {
      const auto xkey = std::make_pair(p.X, p.ldx + (levels << 20));
      auto xm = m->xmaps.find(xkey);
      if (xm == m->xmaps.end()) {
        CUtensorMap map;
        TCN_CHECK(make_tensor_map_2d(&map, p.X, map_rows, p.ldx, p.ldx, TC_BM));
        xm = m->xmaps.emplace(xkey, map).first;
      }
      GemmTcDev q;
      memset(&q, 0, sizeof(q));
      q.Y = p.Y; q.ldy = p.ldy; q.N = n_out; q.bias = p.bias; q.R = p.R; q.ldr = p.ldr; q.M = p.M; q.ldm = p.ldm;
      q.relu = p.relu; q.meta = p.meta; q.nblk = p.nblk; q.dyn = p.dyn; q.x_unpadded = 0;
      q.ntaps = p.ntaps;
      for (int i = 0; i < 3; ++i) q.shift[i] = p.shift[i];
      q.kbp = tc_kbp(c_in); q.c_in = c_in;
      q.colscale = p.colscale; q.colscale_ld = p.colscale_ld;
      q.in_drop_thresh = p.in_drop_thresh; q.in_drop_scale = p.in_drop_scale;
      q.in_drop_seed = p.in_drop_seed; q.in_drop_stream = p.in_drop_stream;
      q.drop_thresh = p.drop_thresh; q.drop_scale = p.drop_scale; q.drop_seed = p.drop_seed; q.drop_stream = p.drop_stream;
      const CUtensorMap* mx32 = nullptr;
      if (gemm_tc_wants_slab(q)) {
        auto x32 = m->xmaps32.find(xkey);
        if (x32 == m->xmaps32.end()) {
          CUtensorMap map;
          TCN_CHECK(make_tensor_map_2d(&map, p.X, map_rows, p.ldx, p.ldx, 32));
          x32 = m->xmaps32.emplace(xkey, map).first;
        }
        mx32 = &x32->second;
      }
      return launch_gemm_tc(xm->second, it->second.mh, it->second.ml, q, cap_blk, st, mx32);
    }
  }
  p.c_in = c_in; p.kpt = rup(c_in, 8); p.n_out = n_out; p.NT8 = (n_out + 7) / 8;
  return launch_tapgemm(p, levels > 1 ? cap_blk : 0, st);
}

int get_wgmap(tcn_model* m, const float* ptr, long rows, int cols, const CUtensorMap** out) {
  const auto key = std::make_pair(ptr, rows * 65536 + cols);
  auto it = m->wgmaps.find(key);
  if (it == m->wgmaps.end()) {
    CUtensorMap map;
    TCN_CHECK(make_tensor_map_2d(&map, ptr, rows, cols, cols, WG_BOX_ROWS, true));
    it = m->wgmaps.emplace(key, map).first;
  }
  *out = &it->second;
  return TCN_OK;
}

bool wgrad_tc_ok(const tcn_model* m, const WgradDev& w) {
  return m->use_tc && (w.ldx % 4 == 0) && (w.ldg % 4 == 0) && (w.c_in % 4 == 0);
}

WgradTcDev to_tc(const WgradDev& w) {
  WgradTcDev q;
  memset(&q, 0, sizeof(q));
  q.meta = w.meta; q.nblk = w.nblk; q.dyn = w.dyn; q.x_unpadded = w.x_unpadded;
  q.n_out = w.n_out; q.c_in = w.c_in; q.ntaps = w.ntaps;
  for (int i = 0; i < 3; ++i) q.shift[i] = w.shift[i];
  q.dW = w.dW; q.db = w.db; q.colscale = w.colscale; q.colscale_ld = w.colscale_ld;
  q.g_drop_thresh = w.g_drop_thresh; q.g_drop_scale = w.g_drop_scale;
  q.g_drop_seed = w.g_drop_seed; q.g_drop_stream = w.g_drop_stream;
  q.x_drop_thresh = w.x_drop_thresh; q.x_drop_scale = w.x_drop_scale;
  q.x_drop_seed = w.x_drop_seed; q.x_drop_stream = w.x_drop_stream;
  return q;
}

// weight-gradient dispatcher: tcgen05 kernel when available, mma.sync kernel otherwise
// slab_id >= 0: deterministic reduction through m->slab[slab_id] (pre-zeroed here, added to dW / db in split order)
int wgrad(tcn_model* m, WgradDev& w, long x_rows, cudaStream_t st, int levels = 1, int slab_id = -1) {
  const long map_rows = (long)levels * m->cfg.max_rows;
  if (wgrad_tc_ok(m, w)) {
    const CUtensorMap *mx, *mg;
    TCN_CHECK(get_wgmap(m, w.X, w.x_unpadded ? x_rows : map_rows, w.ldx, &mx));
    TCN_CHECK(get_wgmap(m, w.G, map_rows, w.ldg, &mg));
    WgradTcDev q = to_tc(w);
    if (slab_id >= 0 && m->det_wgrad) {
      const long nw = (long)w.n_out * w.c_in * w.ntaps;
      q.slab = m->slab[slab_id]; q.slab_stride = nw + w.n_out; q.slab_splits = m->slab_cap[slab_id];
      if (cudaMemsetAsync(q.slab, 0, (size_t)q.slab_splits * q.slab_stride * 4, st) != cudaSuccess) {
        set_error("tcn_model backward: slab memset failed");
        cudaGetLastError();
        return TCN_ERR_CUDA;
      }
      TCN_CHECK(launch_wgrad_tc(*mx, *mg, q, levels * m->max_blk, st));
      if (w.db == w.dW + nw) {   // the bias follows its weight in the flat gradient buffer: one reduction for both
        TCN_CHECK(launch_slab_reduce(w.dW, q.slab, nw + w.n_out, q.row_splits, q.slab_stride, st));
      } else {
        TCN_CHECK(launch_slab_reduce(w.dW, q.slab, nw, q.row_splits, q.slab_stride, st));
        if (w.db != nullptr) TCN_CHECK(launch_slab_reduce(w.db, q.slab + nw, w.n_out, q.row_splits, q.slab_stride, st));
      }
      return TCN_OK;
    }
    return launch_wgrad_tc(*mx, *mg, q, levels * m->max_blk, st);
  }
  return launch_wgrad(w, levels * m->max_blk, st);
}

// the two weight gradients of one residual layer in a single launch
int wgrad_pair(tcn_model* m, WgradDev& w0, WgradDev& w1, long x_rows, cudaStream_t st) {
  static const bool pair_on = std::getenv("TCN_NO_WGRAD_PAIR") == nullptr;
  // (a single launch wins while the step is latency-bound; with several waves of rows two full-width launches do)
  if (pair_on && m->max_blk <= 2 * num_sms() && wgrad_tc_ok(m, w0) && wgrad_tc_ok(m, w1) && w0.n_out <= 64 && w1.n_out <= 64 && !w0.x_unpadded && !w1.x_unpadded) {
    const CUtensorMap *mx0, *mg0, *mx1, *mg1;
    TCN_CHECK(get_wgmap(m, w0.X, m->cfg.max_rows, w0.ldx, &mx0));
    TCN_CHECK(get_wgmap(m, w0.G, m->cfg.max_rows, w0.ldg, &mg0));
    TCN_CHECK(get_wgmap(m, w1.X, m->cfg.max_rows, w1.ldx, &mx1));
    TCN_CHECK(get_wgmap(m, w1.G, m->cfg.max_rows, w1.ldg, &mg1));
    WgradTcDev q0 = to_tc(w0), q1 = to_tc(w1);
    return launch_wgrad_tc_pair(*mx0, *mg0, q0, *mx1, *mg1, q1, m->max_blk, st);
  }
  TCN_CHECK(wgrad(m, w0, x_rows, st));
  return wgrad(m, w1, x_rows, st);
}

// Descriptor tables of the batched weight-gradient launches.  Every pointer is fixed for the life of the model (the
// activations and gradients of all layers keep their own buffers through a step), so the tables are built once.
int build_wg_multi(tcn_model* m, int training) {
  tcn_model::WgMulti& t = m->wgm[training];
  const int C = m->C;
  const float pl = training ? m->layer_drop_p : 0.f;
  std::vector<WgradMultiDesc> descs;
  std::vector<int2> tiles;
  // replay the pointer walk of model_backward
  const float* g = m->Gp[3];
  int gi = 0;
  for (int s = 3; s >= 0; --s) {
    t.tile_first.push_back((int)tiles.size());
    for (int l = m->stage_first[s + 1] - 1; l >= m->stage_first[s]; --l) {
      int sh[3];
      layer_shifts(m, l, sh);
      for (int which = 0; which < 2; ++which) {   // 0: W1 (G = gu, X = layer input, 3 taps); 1: W2 (G = gy, X = h)
        WgradMultiDesc d;
        memset(&d, 0, sizeof(d));
        const float* G = which == 0 ? m->gus[l] : g;
        const float* X = which == 0 ? m->act[l] : m->H[l];
        TCN_CHECK(make_tensor_map_2d(&d.mx, X, m->cfg.max_rows, C, C, WG_BOX_ROWS, true));
        TCN_CHECK(make_tensor_map_2d(&d.mg, G, m->cfg.max_rows, C, C, WG_BOX_ROWS, true));
        WgradTcDev& q = d.p;
        q.meta = m->meta; q.nblk = m->max_blk; q.dyn = m->desc;
        q.n_out = C; q.c_in = C; q.ntaps = which == 0 ? 3 : 1;
        for (int i = 0; i < 3; ++i) q.shift[i] = which == 0 ? sh[i] : 0;
        q.cbn = (C + 31) / 32;
        q.dW = which == 0 ? m->g_(m->off_w1[l]) : m->g_(m->off_w2[l]);
        q.db = which == 0 ? m->g_(m->off_b1[l]) : m->g_(m->off_b2[l]);
        q.g_drop_scale = 1.f; q.x_drop_scale = 1.f;
        if (which == 1 && pl > 0.f) {
          q.g_drop_thresh = drop_thresh(pl); q.g_drop_scale = 1.f / (1.f - pl); q.g_drop_stream = (uint32_t)l;
        }
        const int mt = (q.ntaps * q.cbn + 3) / 4;
        for (int mtile = 0; mtile < mt; ++mtile) tiles.push_back(make_int2((int)descs.size(), mtile));
        descs.push_back(d);
      }
      g = m->gpool[gi++];
    }
    if (s > 0) g = m->gpool[gi++];
    t.tile_count.push_back((int)tiles.size() - t.tile_first.back());
  }
  // row splits per stage: fill the SMs once
  for (size_t i = 0; i < t.tile_count.size(); ++i) {
    int rs = t.tile_count[i] > 0 ? num_sms() / t.tile_count[i] : 1;
    if (rs < 1) rs = 1;
    if (rs > m->max_blk) rs = m->max_blk;
    t.row_splits.push_back(rs);
    for (int k = t.tile_first[i]; k < t.tile_first[i] + t.tile_count[i]; ++k) descs[tiles[k].x].p.row_splits = rs;
  }
  if (cudaMalloc(&t.descs, descs.size() * sizeof(WgradMultiDesc)) != cudaSuccess ||
      cudaMalloc(&t.tiles, tiles.size() * sizeof(int2)) != cudaSuccess ||
      cudaMemcpy(t.descs, descs.data(), descs.size() * sizeof(WgradMultiDesc), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tcn_model backward: descriptor upload failed: %s (the first backward of a model must run outside "
              "stream capture)", cudaGetErrorString(cudaGetLastError()));
    return TCN_ERR_CUDA;
  }
  t.ready = true;
  return TCN_OK;
}

// Descriptor tables of the per-layer weight-gradient kernel, in the order model_backward visits the layers.
int build_wl_table(tcn_model* m, int training) {
  tcn_model::WlTable& t = m->wl[training];
  const float pl = training ? m->layer_drop_p : 0.f;
  std::vector<WgLayerDev> descs;
  std::vector<WgLayerOut> outs;
  const float* g = m->Gp[3];
  int gi = 0;
  for (int s = 3; s >= 0; --s) {
    for (int l = m->stage_first[s + 1] - 1; l >= m->stage_first[s]; --l) {
      WgLayerDev d;
      memset(&d, 0, sizeof(d));
      TCN_CHECK(make_wgrad_layer_maps(&d, m->act[l], m->H[l], m->gus[l], g, m->cfg.max_rows));
      d.masks = m->masks[l];
      layer_shifts(m, l, d.shift);
      d.use_drop = pl > 0.f ? 1 : 0;
      d.drop_scale = pl > 0.f ? 1.f / (1.f - pl) : 1.f;
      d.drop_thresh = pl > 0.f ? drop_thresh(pl) : 0u;
      d.drop_seed = 0u; d.drop_stream = (uint32_t)l;
      d.part = m->wl_part + (size_t)descs.size() * m->wl_splits * WL_PART_FLOATS;
      WgLayerOut o;
      o.part = d.part;
      o.dw1 = m->g_(m->off_w1[l]); o.db1 = m->g_(m->off_b1[l]);
      o.dw2 = m->g_(m->off_w2[l]); o.db2 = m->g_(m->off_b2[l]);
      descs.push_back(d);
      outs.push_back(o);
      g = m->gpool[gi++];
    }
    if (s > 0) g = m->gpool[gi++];
  }
  if (cudaMalloc(&t.descs, descs.size() * sizeof(WgLayerDev)) != cudaSuccess ||
      cudaMalloc(&t.outs, outs.size() * sizeof(WgLayerOut)) != cudaSuccess ||
      cudaMemcpy(t.descs, descs.data(), descs.size() * sizeof(WgLayerDev), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(t.outs, outs.data(), outs.size() * sizeof(WgLayerOut), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tcn_model backward: descriptor upload failed: %s (the first backward of a model must run outside "
              "stream capture)", cudaGetErrorString(cudaGetLastError()));
    return TCN_ERR_CUDA;
  }
  t.ready = true;
  return TCN_OK;
}

}  // namespace

// ================================================================================================ create
extern "C" int tcn_model_create(const tcn_model_config* cfg, tcn_model** out) {
  TCN_REQUIRE(cfg && out, "tcn_model_create: null pointer");
  TCN_REQUIRE(cfg->layers_pg > 0 && cfg->layers_r >= 0 && cfg->num_r == 3,
              "tcn_model_create: the FPN path needs exactly 3 refinement stages (network.py:98-106)");
  TCN_REQUIRE(cfg->channels > 0 && cfg->channels % 4 == 0 && cfg->in_dim > 0 && cfg->in_dim % 4 == 0,
              "tcn_model_create: channels and in_dim must be multiples of 4");
  TCN_REQUIRE(cfg->max_rows > 0 && cfg->max_rows % kBlkRows == 0 && cfg->max_seqs > 0,
              "tcn_model_create: max_rows must be a positive multiple of 128");
  TCN_REQUIRE(cfg->layers_pg <= 20 && cfg->layers_r <= 20, "tcn_model_create: dilation 2^i overflows");
  tcn_model* m = new (std::nothrow) tcn_model();
  TCN_REQUIRE(m != nullptr, "tcn_model_create: out of host memory");
  m->cfg = *cfg;
  m->C = cfg->channels; m->D = cfg->in_dim;
  m->NH = cfg->head_sizes[0] + cfg->head_sizes[1] + cfg->head_sizes[2] + cfg->head_sizes[3];
  m->LDH = rup(m->NH, 4);
  const int C = m->C, D = m->D, NH = m->NH;
  // ---- stages / layers
  m->stage_first.push_back(0);
  for (int i = 0; i < cfg->layers_pg; ++i) m->dilation.push_back(1 << i);
  m->stage_first.push_back((int)m->dilation.size());
  for (int s = 0; s < cfg->num_r; ++s) {
    for (int i = 0; i < cfg->layers_r; ++i) m->dilation.push_back(1 << i);
    m->stage_first.push_back((int)m->dilation.size());
  }
  m->L = (int)m->dilation.size();
  // ---- flat parameter layout (every tensor starts on a 16-byte boundary except inside the head block)
  long off = 0;
  auto slot = [&](long size, bool align) {
    if (align) off = rupl(off, 4);
    const long o = off;
    m->slots.push_back({o, size});
    off += size;
    return o;
  };
  m->off_proj_w = slot((long)C * D, true);
  m->off_proj_b = slot(C, true);
  for (int l = 0; l < m->L; ++l) {
    m->off_w1.push_back(slot((long)C * C * 3, true));
    m->off_b1.push_back(slot(C, true));
    m->off_w2.push_back(slot((long)C * C, true));
    m->off_b2.push_back(slot(C, true));
  }
  m->off_lat_w = slot((long)C * C, true);
  m->off_lat_b = slot(C, true);
  m->off_head_w = slot((long)cfg->head_sizes[0] * C, true);
  for (int h = 1; h < 4; ++h) slot((long)cfg->head_sizes[h] * C, false);
  m->off_head_b = slot(cfg->head_sizes[0], true);
  for (int h = 1; h < 4; ++h) slot(cfg->head_sizes[h], false);
  m->n_params = rupl(off, 4);

  // ---- prepared-weight buffer and the batched prep job table
  std::vector<PrepJob> jobs;
  std::vector<SplitJob> sjobs;
  long f4 = 0, sfirst = 0, sdst = 0;
  auto reg = [&](long src, int n_out, int c_in, int ntaps, int tr) {
    const long at = f4 * 4;
    add_job(jobs, f4, src, n_out, c_in, ntaps, tr);
    // the same weight for the tcgen05 kernels
    SplitJob sj;
    sj.first = sfirst; sj.src_off = src; sj.dst_off = sdst;
    sj.n_out = n_out; sj.c_in = c_in; sj.ntaps = ntaps; sj.transpose = tr;
    sj.rows_pad = (int)tc_weight_rows(n_out, c_in, tr);
    sj.kcols = (int)tc_weight_cols(n_out, c_in, ntaps, tr);
    tcn_model::TcW w;
    w.off = sdst; w.rows = sj.rows_pad; w.cols = sj.kcols;
    m->tcw[at] = w;
    const long n = (long)sj.rows_pad * sj.kcols;
    sfirst += n;
    sdst += (n + 255) / 256 * 256;  // keep every matrix 1 KB aligned
    sjobs.push_back(sj);
    return at;
  };
  m->wf_proj = reg(m->off_proj_w, C, D, 1, 0);
  for (int l = 0; l < m->L; ++l) {
    m->wf_w1.push_back(reg(m->off_w1[l], C, C, 3, 0));
    m->wf_w2.push_back(reg(m->off_w2[l], C, C, 1, 0));
    m->wf_w1T.push_back(reg(m->off_w1[l], C, C, 3, 1));
    m->wf_w2T.push_back(reg(m->off_w2[l], C, C, 1, 1));
  }
  m->wf_lat = reg(m->off_lat_w, C, C, 1, 0);
  m->wf_latT = reg(m->off_lat_w, C, C, 1, 1);
  m->wf_head = reg(m->off_head_w, NH, C, 1, 0);
  m->wf_headT = reg(m->off_head_w, NH, C, 1, 1);
  m->njobs = (int)jobs.size();
  m->prep_total_f4 = f4;
  m->wf_floats = f4 * 4;
  m->nsjobs = (int)sjobs.size();
  m->split_total = sfirst;
  m->tc_wfloats = sdst;

  // ---- workspace carve-up
  const long rows = cfg->max_rows;
  m->max_blk = cfg->max_rows / kBlkRows;
  size_t bytes = 0;
  auto carve = [&](size_t n) {
    const size_t at = bytes;
    bytes += (n + 255) / 256 * 256;
    return at;
  };
  const size_t o_wf = carve((size_t)m->wf_floats * 4);
  const size_t o_jobs = carve(jobs.size() * sizeof(PrepJob));
  const size_t o_sjobs = carve(sjobs.size() * sizeof(SplitJob));
  const size_t o_tcwhi = carve((size_t)m->tc_wfloats * 4), o_tcwlo = carve((size_t)m->tc_wfloats * 4);
  std::vector<size_t> o_act(m->L + 1), o_H(m->L);
  // stage outputs f0..f2 contiguous, then p1..p3 and f3 contiguous (rows * C * 4 is a multiple of 256: carve() adds no gap)
  const size_t o_fstack = carve((size_t)3 * rows * C * 4);
  const size_t o_pstack = carve((size_t)4 * rows * C * 4);
  for (int i = 0; i <= m->L; ++i) {
    int stage_out = -1;
    for (int st = 0; st < 4; ++st)
      if (i == m->stage_first[st + 1]) stage_out = st;
    if (stage_out >= 0 && stage_out < 3) o_act[i] = o_fstack + (size_t)stage_out * rows * C * 4;
    else if (stage_out == 3) o_act[i] = o_pstack + (size_t)3 * rows * C * 4;
    else o_act[i] = carve((size_t)rows * C * 4);
  }
  for (int i = 0; i < m->L; ++i) o_H[i] = carve((size_t)rows * C * 4);
  std::vector<size_t> o_masks(m->L);
  for (int i = 0; i < m->L; ++i) o_masks[i] = carve((size_t)rows * 16);
  size_t o_P[3], o_log[4], o_dL[4], o_Gp[4], o_cum[4];
  std::vector<size_t> o_gpool(m->L + 4), o_gus(m->L);
  for (int i = 0; i < 3; ++i) o_P[i] = o_pstack + (size_t)i * rows * C * 4;
  for (int i = 0; i < 4; ++i) o_log[i] = carve((size_t)rows * m->LDH * 4);
  for (int i = 0; i < 4; ++i) o_dL[i] = carve((size_t)rows * m->LDH * 4);
  for (int i = 0; i < 4; ++i) o_Gp[i] = carve((size_t)rows * C * 4);
  for (int i = 0; i < 4; ++i) o_cum[i] = carve((size_t)rows * m->LDH * 4);
  for (auto& o : o_gpool) o = carve((size_t)rows * C * 4);
  for (auto& o : o_gus) o = carve((size_t)rows * C * 4);
  m->slab_cap[0] = wgrad_tc_splits_cap(NH, C, 1);
  m->slab_cap[1] = wgrad_tc_splits_cap(C, C, 1);
  m->slab_cap[2] = wgrad_tc_splits_cap(C, D, 1);
  const size_t o_slab0 = carve((size_t)m->slab_cap[0] * ((size_t)NH * C + NH) * 4);
  const size_t o_slab1 = carve((size_t)m->slab_cap[1] * ((size_t)C * C + C) * 4);
  const size_t o_slab2 = carve((size_t)m->slab_cap[2] * ((size_t)C * D + C) * 4);
  wgrad_layers_plan(m->max_blk, &m->wl_lg, &m->wl_splits);
  if (const char* e = std::getenv("TCN_WL_LG")) {   // experiments: layers per weight-gradient launch
    const int lg = atoi(e);
    if (lg >= 1 && lg <= 16) {
      m->wl_lg = lg;
      m->wl_splits = num_sms() / lg;
      if (m->wl_splits > m->max_blk * 8) m->wl_splits = m->max_blk * 8;
      if (m->wl_splits < 1) m->wl_splits = 1;
    }
  }
  const size_t o_wlpart = carve((size_t)m->L * m->wl_splits * WL_PART_FLOATS * 4);
  const size_t o_cs = carve((size_t)cfg->max_seqs * D * 4);
  m->proj_tc = tcn_gemm_tc_supported(D, C) != 0;
  const size_t proj_wf = (size_t)tcn_split_weight_floats(C, D, 1, 0);
  const size_t o_whi = carve(proj_wf * 4), o_wlo = carve(proj_wf * 4);
  // [desc | desc4 | desc3 | meta (max_blk) | meta4 (4 max_blk)]: one upload per batch
  const size_t o_desc = carve(3 * sizeof(BatchDesc) + (size_t)5 * m->max_blk * sizeof(BlkMeta));
  const size_t o_cu = carve((size_t)m->LDH * 4), o_cscale = carve((size_t)m->LDH * 4), o_pw = carve((size_t)m->LDH * 4);
  const size_t o_ch = carve((size_t)m->LDH * 4), o_loss = carve(64);
  m->ws_bytes = bytes;
  cudaError_t e = cudaMalloc(&m->ws, bytes);
  if (e == cudaSuccess) e = cudaMemset(m->ws, 0, bytes);
  m->slot_bytes = 3 * sizeof(BatchDesc) + (size_t)5 * m->max_blk * sizeof(BlkMeta);
  if (e == cudaSuccess) e = cudaMallocHost(&m->desc_host, m->slot_bytes * tcn_model::kSlots);
  for (int i = 0; i < tcn_model::kSlots && e == cudaSuccess; ++i)
    e = cudaEventCreateWithFlags(&m->slot_done[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("tcn_model_create: allocating %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    if (m->ws) cudaFree(m->ws);
    delete m;
    return TCN_ERR_CUDA;
  }
  m->wf = reinterpret_cast<float*>(m->ws + o_wf);
  m->jobs_dev = reinterpret_cast<PrepJob*>(m->ws + o_jobs);
  m->sjobs_dev = reinterpret_cast<SplitJob*>(m->ws + o_sjobs);
  m->tc_whi = reinterpret_cast<float*>(m->ws + o_tcwhi);
  m->tc_wlo = reinterpret_cast<float*>(m->ws + o_tcwlo);
  for (int i = 0; i <= m->L; ++i) m->act.push_back(reinterpret_cast<float*>(m->ws + o_act[i]));
  for (int i = 0; i < m->L; ++i) m->H.push_back(reinterpret_cast<float*>(m->ws + o_H[i]));
  for (int i = 0; i < m->L; ++i) m->masks.push_back(reinterpret_cast<uint32_t*>(m->ws + o_masks[i]));
  for (int i = 0; i < 3; ++i) m->P[i] = reinterpret_cast<float*>(m->ws + o_P[i]);
  for (int i = 0; i < 4; ++i) {
    m->logits[i] = reinterpret_cast<float*>(m->ws + o_log[i]);
    m->dL[i] = reinterpret_cast<float*>(m->ws + o_dL[i]);
    m->Gp[i] = reinterpret_cast<float*>(m->ws + o_Gp[i]);
    m->cum[i] = reinterpret_cast<float*>(m->ws + o_cum[i]);
  }
  m->stack_levels = std::getenv("TCN_NO_STACK") == nullptr;
  for (auto o : o_gpool) m->gpool.push_back(reinterpret_cast<float*>(m->ws + o));
  for (auto o : o_gus) m->gus.push_back(reinterpret_cast<float*>(m->ws + o));
  m->wl_part = reinterpret_cast<float*>(m->ws + o_wlpart);
  m->slab[0] = reinterpret_cast<float*>(m->ws + o_slab0);
  m->slab[1] = reinterpret_cast<float*>(m->ws + o_slab1);
  m->slab[2] = reinterpret_cast<float*>(m->ws + o_slab2);
  m->det_wgrad = std::getenv("TCN_WGRAD_ATOMIC") == nullptr;
  m->use_wl = std::getenv("TCN_WGRAD_PAIR") == nullptr;
  m->overlap_wgrad = std::getenv("TCN_NO_WGRAD_STREAM") == nullptr;
  if (cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) != cudaSuccess) m->overlap_wgrad = false;
  m->evs.resize(m->L + 8, nullptr);
  for (auto& ev : m->evs)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) m->overlap_wgrad = false;
  cudaGetLastError();
  m->colscale = reinterpret_cast<float*>(m->ws + o_cs);
  m->proj_whi = reinterpret_cast<float*>(m->ws + o_whi);
  m->proj_wlo = reinterpret_cast<float*>(m->ws + o_wlo);
  if (m->proj_tc) {
    const long wr = tc_weight_rows(C, D, 0), wc = tc_weight_cols(C, D, 1, 0);
    if (make_tensor_map_2d(&m->map_whi, m->proj_whi, wr, wc, wc, 64) != TCN_OK ||
        make_tensor_map_2d(&m->map_wlo, m->proj_wlo, wr, wc, wc, 64) != TCN_OK)
      m->proj_tc = false;  // driver without tensor-map support: stay on the mma.sync path
  }
  m->desc = reinterpret_cast<BatchDesc*>(m->ws + o_desc);
  m->desc4 = m->desc + 1;
  m->desc3 = m->desc + 2;
  m->meta = reinterpret_cast<BlkMeta*>(m->ws + o_desc + 3 * sizeof(BatchDesc));
  m->meta4 = m->meta + m->max_blk;
  m->col_unit = reinterpret_cast<float*>(m->ws + o_cu);
  m->col_scale = reinterpret_cast<float*>(m->ws + o_cscale);
  m->pos_w = reinterpret_cast<float*>(m->ws + o_pw);
  m->col_head = reinterpret_cast<int*>(m->ws + o_ch);
  m->loss8 = reinterpret_cast<float*>(m->ws + o_loss);
  e = cudaMemcpy(m->jobs_dev, jobs.data(), jobs.size() * sizeof(PrepJob), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(m->sjobs_dev, sjobs.data(), sjobs.size() * sizeof(SplitJob), cudaMemcpyHostToDevice);
  // tcgen05 path available? (needs the driver's tensor-map encoder; TCN_NO_TCGEN05=1 forces the mma.sync kernels)
  m->use_tc = (std::getenv("TCN_NO_TCGEN05") == nullptr) && (C % 4 == 0);
  m->fused_tc = std::getenv("TCN_NO_FUSED_TC") == nullptr;
  m->fused_bwd = m->fused_tc && std::getenv("TCN_NO_FUSED_BWD") == nullptr;
  m->wg_multi = std::getenv("TCN_WGRAD_MULTI") != nullptr;
  if (m->use_tc) {
    for (auto& kv : m->tcw) {
      tcn_model::TcW& w = kv.second;
      if (make_tensor_map_2d(&w.mh, m->tc_whi + w.off, w.rows, w.cols, w.cols, 64) != TCN_OK ||
          make_tensor_map_2d(&w.ml, m->tc_wlo + w.off, w.rows, w.cols, w.cols, 64) != TCN_OK) {
        m->use_tc = false;
        break;
      }
    }
  }
  if (!m->use_tc) m->proj_tc = false;
  if (e != cudaSuccess) {
    set_error("tcn_model_create: job table upload failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    tcn_model_destroy(m);
    return TCN_ERR_CUDA;
  }
  *out = m;
  const float w[4] = {1.f, 0.1f, 0.1f, 0.1f};
  return tcn_model_set_loss(m, w, nullptr);
}

extern "C" void tcn_model_destroy(tcn_model* m) {
  if (!m) return;
  if (m->ws) cudaFree(m->ws);
  if (m->desc_host) cudaFreeHost(m->desc_host);
  for (int i = 0; i < tcn_model::kSlots; ++i)
    if (m->slot_done[i]) cudaEventDestroy(m->slot_done[i]);
  for (auto& t : m->wgm) {
    if (t.descs) cudaFree(t.descs);
    if (t.tiles) cudaFree(t.tiles);
  }
  for (auto& t : m->wl) {
    if (t.descs) cudaFree(t.descs);
    if (t.outs) cudaFree(t.outs);
  }
  for (auto ev : m->evs)
    if (ev) cudaEventDestroy(ev);
  if (m->side) cudaStreamDestroy(m->side);
  delete m;
}

extern "C" long long tcn_model_num_params(const tcn_model* m) { return m ? m->n_params : 0; }
// diagnostics (include/tcn_b200.h): kind 0 = input of layer idx (idx = L: last output), 1 = h of layer idx
extern "C" int tcn_model_debug_ptr(tcn_model* m, int kind, int idx, void** out) {
  if (!m || !out || idx < 0 || idx > m->L || (kind == 1 && idx >= m->L)) return TCN_ERR_INVALID_ARG;
  *out = kind == 0 ? (void*)m->act[idx] : (void*)m->H[idx];
  return TCN_OK;
}

extern "C" int tcn_model_num_tensors(const tcn_model* m) { return m ? (int)m->slots.size() : 0; }

extern "C" int tcn_model_param_layout(const tcn_model* m, long long* offsets, long long* sizes, int n) {
  TCN_REQUIRE(m && offsets && sizes && n == (int)m->slots.size(), "tcn_model_param_layout: bad arguments");
  for (int i = 0; i < n; ++i) {
    offsets[i] = m->slots[i].off;
    sizes[i] = m->slots[i].size;
  }
  return TCN_OK;
}

extern "C" int tcn_model_bind(tcn_model* m, float* params, float* grads) {
  TCN_REQUIRE(m && params, "tcn_model_bind: null pointer");
  TCN_REQUIRE((reinterpret_cast<uintptr_t>(params) & 15) == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0,
              "tcn_model_bind: buffers must be 16-byte aligned");
  m->params = params;
  m->grads = grads;
  return TCN_OK;
}

extern "C" int tcn_model_set_loss(tcn_model* m, const float* head_weights, const float* pos_w_host) {
  TCN_REQUIRE(m && head_weights, "tcn_model_set_loss: null pointer");
  std::vector<float> unit(m->LDH, 0.f), scale(m->LDH, 0.f), pw(m->LDH, 1.f);
  std::vector<int> head(m->LDH, 0);
  int c = 0;
  for (int h = 0; h < 4; ++h) {
    m->head_w[h] = head_weights[h];
    for (int k = 0; k < m->cfg.head_sizes[h]; ++k, ++c) {
      unit[c] = 1.f / (float)m->cfg.head_sizes[h];
      scale[c] = head_weights[h] / (float)m->cfg.head_sizes[h];
      head[c] = h;
      if (pos_w_host) pw[c] = pos_w_host[c];
    }
  }
  m->has_pos_w = pos_w_host != nullptr;
  cudaError_t e = cudaMemcpy(m->col_unit, unit.data(), unit.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->col_scale, scale.data(), scale.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->pos_w, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->col_head, head.data(), head.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("tcn_model_set_loss: upload failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  return TCN_OK;
}

extern "C" int tcn_model_set_loss_norm(tcn_model* m, int norm_seqs) {
  TCN_REQUIRE(m && norm_seqs >= 0, "tcn_model_set_loss_norm: bad arguments");
  m->norm_seqs = norm_seqs;   // takes effect with the next tcn_model_set_batch
  return TCN_OK;
}

extern "C" int tcn_model_set_dropout(tcn_model* m, float input_mask_p, float chan_drop_p, float layer_drop_p) {
  TCN_REQUIRE(m, "tcn_model_set_dropout: null pointer");
  TCN_REQUIRE(input_mask_p >= 0.f && input_mask_p < 1.f && chan_drop_p >= 0.f && chan_drop_p < 1.f &&
                  layer_drop_p >= 0.f && layer_drop_p < 1.f,
              "tcn_model_set_dropout: probabilities must be in [0, 1)");
  m->input_mask_p = input_mask_p;
  m->chan_drop_p = chan_drop_p;
  m->layer_drop_p = layer_drop_p;
  return TCN_OK;
}

extern "C" int tcn_model_set_batch(tcn_model* m, const int* meta_host, int nblk, int rows, int num_seqs, int frames,
                                   unsigned seed, tcn_stream_t stream) {
  TCN_REQUIRE(m && meta_host, "tcn_model_set_batch: null pointer");
  TCN_REQUIRE(nblk > 0 && nblk <= m->max_blk && rows == nblk * kBlkRows,
              "tcn_model_set_batch: %d blocks exceed the capacity of %d (max_rows)", nblk, m->max_blk);
  TCN_REQUIRE(num_seqs > 0 && num_seqs <= m->cfg.max_seqs, "tcn_model_set_batch: too many sequences");
  // pinned staging slots are reused round-robin: wait until this slot's previous upload has been consumed
  const int sl = m->slot;
  m->slot = (m->slot + 1) % tcn_model::kSlots;
  cudaEventSynchronize(m->slot_done[sl]);
  BatchDesc* d = reinterpret_cast<BatchDesc*>(m->desc_host + (size_t)sl * m->slot_bytes);
  d->nblk = nblk; d->rows = rows; d->num_seqs = num_seqs; d->frames = frames; d->seed = seed;
  d->norm_seqs = m->norm_seqs; d->pad[0] = d->pad[1] = 0;
  // level-stacked views: level l occupies rows [l * max_rows, (l + 1) * max_rows) of a stack and blocks
  // [l * max_blk, l * max_blk + nblk) of meta4; the blocks in between are empty (hi = 0: every kernel skips them)
  const int MB = m->max_blk, MR = m->cfg.max_rows;
  d[1] = d[0]; d[1].nblk = 3 * MB + nblk; d[1].rows = 3 * MR + rows;
  d[2] = d[0]; d[2].nblk = 2 * MB + nblk; d[2].rows = 2 * MR + rows;
  BlkMeta* mh = reinterpret_cast<BlkMeta*>(d + 3);
  memcpy(mh, meta_host, (size_t)nblk * sizeof(BlkMeta));
  if (nblk < MB) memset(mh + nblk, 0, (size_t)(MB - nblk) * sizeof(BlkMeta));
  BlkMeta* m4 = mh + MB;
  memset(m4, 0, (size_t)4 * MB * sizeof(BlkMeta));
  for (int lv = 0; lv < 4; ++lv)
    for (int b = 0; b < nblk; ++b) {
      BlkMeta e4 = mh[b];
      e4.lo += lv * MR; e4.hi += lv * MR; e4.in_delta -= lv * MR;
      m4[(size_t)lv * MB + b] = e4;
    }
  cudaError_t e = cudaMemcpyAsync(m->desc, d, m->slot_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaEventRecord(m->slot_done[sl], (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error("tcn_model_set_batch: upload failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  return TCN_OK;
}

// ================================================================================================ forward
static int model_forward(tcn_model* m, const float* x, long x_rows, int training, bool save_h, cudaStream_t st) {
  const int C = m->C, D = m->D, L = m->L;
  m->fwd_training = training != 0;
  const float pl = training ? m->layer_drop_p : 0.f;
  // 0. weights -> fragment order (one launch) for the mma.sync kernels; nothing reads them when every contraction of the
  // model runs on the tcgen05 kernels (split hi / lo weights below)
  const bool all_tc = m->use_tc && m->proj_tc && !(C == 64 && !m->fused_tc);
  if (!all_tc) TCN_CHECK(launch_prep_batched(m->jobs_dev, m->njobs, m->params, m->wf, m->prep_total_f4, st));
  if (m->use_tc)
    TCN_CHECK(launch_split_batched(m->sjobs_dev, m->nsjobs, m->params, m->tc_whi, m->tc_wlo, m->split_total, st));
  // 1. stage-input projection (network.py:113,122-129), input mask + channel dropout folded into the load
  const bool chan = training && m->chan_drop_p > 0.f;
  if (chan) TCN_CHECK(launch_chan_scale(m->colscale, D, m->cfg.max_seqs, m->desc, m->chan_drop_p, 0u, kStreamChan, st));
  if (m->proj_tc) {  // tcgen05 + TMA (gemm_tc.cu)
    TCN_CHECK(launch_split_weight(m->p_(m->off_proj_w), m->proj_whi, m->proj_wlo, C, D, 1, 0, st));
    if (m->map_x_ptr != x || m->map_x_rows != x_rows) {
      TCN_CHECK(make_tensor_map_2d(&m->map_x, x, x_rows, D, D, 128));
      m->map_x_ptr = x; m->map_x_rows = x_rows;
    }
    GemmTcDev p;
    memset(&p, 0, sizeof(p));
    p.Y = m->act[0]; p.ldy = C; p.N = C; p.bias = m->p_(m->off_proj_b);
    p.meta = m->meta; p.nblk = m->max_blk; p.dyn = m->desc; p.x_unpadded = 1;
    p.ntaps = 1; p.kbp = tc_kbp(D); p.c_in = D;
    if (chan) { p.colscale = m->colscale; p.colscale_ld = D; }
    p.in_drop_scale = 1.f; p.drop_scale = 1.f;
    if (training && m->input_mask_p > 0.f) { p.in_drop_thresh = drop_thresh(m->input_mask_p); p.in_drop_stream = kStreamMask; }
    TCN_CHECK(launch_gemm_tc(m->map_x, m->map_whi, m->map_wlo, p, m->max_blk, st));
  } else {
    TapGemmDev p = base_tapgemm(m);
    p.X = x; p.ldx = D; p.x_unpadded = 1;
    if (chan) { p.colscale = m->colscale; p.colscale_ld = D; }
    if (training && m->input_mask_p > 0.f) {
      p.in_drop_thresh = drop_thresh(m->input_mask_p); p.in_drop_scale = 1.f; p.in_drop_stream = kStreamMask;
    }
    p.Wf = m->wf_(m->wf_proj); p.bias = m->p_(m->off_proj_b);
    p.Y = m->act[0]; p.ldy = C;
    TCN_CHECK(gemm(m, p, D, C, st));
  }
  // 2. residual layers
  for (int l = 0; l < L; ++l) {
    int s[3];
    layer_shifts(m, l, s);
    // one wave of tiles or less: the fused mma.sync kernel has the shorter latency; beyond that two tcgen05 launches
    // (conv3 + ReLU -> h, 1x1 + dropout + residual -> y) move more frames per second
    if (C == 64 && m->use_tc && m->fused_tc) {
      // one tcgen05 launch: conv3 accumulates in TMEM, bias + ReLU on the accumulator, h re-enters the 1x1 conv as a
      // TMEM A operand (gemm_tc.cu: layer_fwd_tc_kernel)
      const long k1 = reinterpret_cast<const float*>(m->wf_(m->wf_w1[l])) - m->wf;
      const long k2 = reinterpret_cast<const float*>(m->wf_(m->wf_w2[l])) - m->wf;
      auto w1 = m->tcw.find(k1), w2 = m->tcw.find(k2);
      TCN_REQUIRE(w1 != m->tcw.end() && w2 != m->tcw.end(), "tcn_model forward: missing split weights for the fused layer");
      const auto xkey = std::make_pair((const float*)m->act[l], C);
      auto xm = m->xmaps.find(xkey);
      if (xm == m->xmaps.end()) {
        CUtensorMap map;
        TCN_CHECK(make_tensor_map_2d(&map, m->act[l], m->cfg.max_rows, C, C, TC_BM));
        xm = m->xmaps.emplace(xkey, map).first;
      }
      LayerTcDev p;
      memset(&p, 0, sizeof(p));
      p.h.Y = save_h ? m->H[l] : nullptr; p.h.ldy = C; p.h.N = C; p.h.bias = m->p_(m->off_b1[l]); p.h.relu = 1;
      p.h.drop_scale = 1.f; p.h.in_drop_scale = 1.f;
      p.y.Y = m->act[l + 1]; p.y.ldy = C; p.y.N = C; p.y.bias = m->p_(m->off_b2[l]); p.y.R = m->act[l]; p.y.ldr = C;
      p.y.meta = m->meta; p.y.nblk = m->max_blk; p.y.dyn = m->desc; p.y.ntaps = 3; p.y.kbp = 2; p.y.c_in = C;
      for (int i = 0; i < 3; ++i) p.y.shift[i] = s[i];
      p.y.in_drop_scale = 1.f;
      p.y.drop_thresh = pl > 0.f ? drop_thresh(pl) : 0u;
      p.y.drop_scale = pl > 0.f ? 1.f / (1.f - pl) : 1.f;
      p.y.drop_seed = 0u; p.y.drop_stream = (uint32_t)l;
      p.masks = (save_h && m->fused_bwd) ? m->masks[l] : nullptr;
      // layer 0 without programmatic launch: a full dependency behind the weight split, which lets every later layer
      // kernel of the step (forward and backward) fetch its weights before its griddepcontrol.wait
      TCN_CHECK(launch_layer_fwd_tc(xm->second, w1->second.mh, w1->second.ml, w2->second.mh, w2->second.ml, p, m->max_blk,
                                    st, l != 0));
    } else if (C == 64 && !(m->use_tc && m->max_blk > 2 * num_sms())) {
      LayerFwdDev p;
      p.X = m->act[l]; p.Y = m->act[l + 1]; p.H = save_h ? m->H[l] : nullptr;
      p.W1f = m->wf_(m->wf_w1[l]); p.W2f = m->wf_(m->wf_w2[l]);
      p.b1 = m->p_(m->off_b1[l]); p.b2 = m->p_(m->off_b2[l]);
      p.meta = m->meta; p.nblk = m->max_blk; p.dyn = m->desc;
      for (int i = 0; i < 3; ++i) p.shift[i] = s[i];
      p.drop_thresh = pl > 0.f ? drop_thresh(pl) : 0u;
      p.drop_scale = pl > 0.f ? 1.f / (1.f - pl) : 1.f;
      p.drop_seed = 0u; p.drop_stream = (uint32_t)l;
      TCN_CHECK(launch_layer_fwd64(p, m->max_blk, st));
    } else {
      TapGemmDev p = base_tapgemm(m);
      p.X = m->act[l]; p.ldx = C; p.Wf = m->wf_(m->wf_w1[l]); p.bias = m->p_(m->off_b1[l]);
      p.Y = m->H[l]; p.ldy = C; p.ntaps = 3; p.relu = 1;
      for (int i = 0; i < 3; ++i) p.shift[i] = s[i];
      TCN_CHECK(gemm(m, p, C, C, st));
      TapGemmDev q = base_tapgemm(m);
      q.X = m->H[l]; q.ldx = C; q.Wf = m->wf_(m->wf_w2[l]); q.bias = m->p_(m->off_b2[l]);
      q.Y = m->act[l + 1]; q.ldy = C; q.R = m->act[l]; q.ldr = C;
      if (pl > 0.f) { q.drop_thresh = drop_thresh(pl); q.drop_scale = 1.f / (1.f - pl); q.drop_stream = (uint32_t)l; }
      TCN_CHECK(gemm(m, q, C, C, st));
    }
  }
  // 3. FPN (network.py:98-106): p4 = f3, p3 = p4 + lat(f2), p2 = p3 + lat(f1), p1 = p2 + lat(f0)
  const float* f[4];
  for (int s = 0; s < 4; ++s) f[s] = m->act[m->stage_first[s + 1]];
  const float* top = f[3];
  for (int lv = 2; lv >= 0; --lv) {
    TapGemmDev p = base_tapgemm(m);
    p.X = f[lv]; p.ldx = C; p.Wf = m->wf_(m->wf_lat); p.bias = m->p_(m->off_lat_b);
    p.Y = m->P[lv]; p.ldy = C; p.R = top; p.ldr = C;
    TCN_CHECK(gemm(m, p, C, C, st));
    top = m->P[lv];
  }
  // 4. heads on the four levels (network.py:63-67), all 131 classes in one GEMM per level
  if (m->stack_levels) {   // [p1 | p2 | p3 | p4] -> [logits of the four levels]: one launch over the 4x block table
    TapGemmDev p = base_tapgemm(m);
    p.meta = m->meta4; p.nblk = 4 * m->max_blk; p.dyn = m->desc4;
    p.X = m->P[0]; p.ldx = C; p.Wf = m->wf_(m->wf_head); p.bias = m->p_(m->off_head_b);
    p.Y = m->logits[0]; p.ldy = m->LDH;
    TCN_CHECK(gemm(m, p, C, m->NH, st, 4));
    return TCN_OK;
  }
  for (int lv = 0; lv < 4; ++lv) {
    TapGemmDev p = base_tapgemm(m);
    p.X = lv < 3 ? m->P[lv] : f[3]; p.ldx = C; p.Wf = m->wf_(m->wf_head); p.bias = m->p_(m->off_head_b);
    p.Y = m->logits[lv]; p.ldy = m->LDH;
    TCN_CHECK(gemm(m, p, C, m->NH, st));
  }
  return TCN_OK;
}

// ================================================================================================ backward
// gl[lv]: gradient w.r.t. the logits of level lv (rows, LDH; pad columns zero) or nullptr;
// gf[lv]: extra gradient w.r.t. feature level lv (rows, C) or nullptr.
static int model_backward(tcn_model* m, const float* x, long x_rows, const float* const* gl,
                          const float* const* gf, cudaStream_t st, bool cum_ready = false) {
  const int C = m->C, D = m->D, NH = m->NH, LDH = m->LDH;
  TCN_REQUIRE(m->grads != nullptr, "tcn_model backward: no gradient buffer bound");
  const float pl = m->fwd_training ? m->layer_drop_p : 0.f;
  if (cudaMemsetAsync(m->grads, 0, (size_t)m->n_params * 4, st) != cudaSuccess) {
    set_error("tcn_model backward: memset failed");
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  // The input-gradient chain (critical path) stays on `st`; weight gradients only feed the optimizer, so they run
  // on a second stream, ordered by events (inside a graph capture these become plain graph edges).
  cudaStream_t ws = m->overlap_wgrad ? m->side : st;
  int nev = 0;
  auto hand_over = [&]() -> int {  // everything enqueued on st so far is visible to ws
    if (ws == st) return TCN_OK;
    cudaEvent_t ev = m->evs[nev++];
    if (cudaEventRecord(ev, st) != cudaSuccess || cudaStreamWaitEvent(ws, ev, 0) != cudaSuccess) {
      set_error("tcn_model backward: event hand-over failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TCN_ERR_CUDA;
    }
    return TCN_OK;
  };
  const float* f[4];
  for (int s = 0; s < 4; ++s) f[s] = m->act[m->stage_first[s + 1]];
  const float* plv[4] = {m->P[0], m->P[1], m->P[2], f[3]};
  for (int lv = 0; lv < 4; ++lv) {
    if (gf && gf[lv]) {
      set_error("tcn_model backward: gradients w.r.t. the feature maps are not supported yet");
      return TCN_ERR_UNSUPPORTED;
    }
    if (!(gl && gl[lv])) {
      set_error("tcn_model backward: every level needs a logits gradient");
      return TCN_ERR_UNSUPPORTED;
    }
  }
  TCN_CHECK(hand_over());  // gradient buffer zeroed, logits gradients ready
  // heads: weight grads (side stream), then the gradient of each FPN level, accumulated top-down (p_l feeds p_{l-1})
  // stacked: the logit gradients are the executor's own dL stack and the loss kernel left their running sums in cum
  const bool stacked = m->stack_levels && cum_ready && gl[0] == m->dL[0] && gl[1] == m->dL[1] && gl[2] == m->dL[2] &&
                       gl[3] == m->dL[3];
  if (stacked) {
    {   // head weights: sum over the four levels of dL_l^T p_l
      WgradDev w = base_wgrad(m);
      w.meta = m->meta4; w.nblk = 4 * m->max_blk; w.dyn = m->desc4;
      w.G = m->dL[0]; w.ldg = LDH; w.g_cols = LDH; w.X = m->P[0]; w.ldx = C;
      w.n_out = NH; w.c_in = C; w.dW = m->g_(m->off_head_w); w.db = m->g_(m->off_head_b);
      TCN_CHECK(wgrad(m, w, x_rows, ws, 4, 0));
    }
    {   // gradient of every level: (dL_0 + ... + dL_l) W_head -- the heads share their weights (network.py:63-67)
      TapGemmDev p = base_tapgemm(m);
      p.meta = m->meta4; p.nblk = 4 * m->max_blk; p.dyn = m->desc4;
      p.X = m->cum[0]; p.ldx = LDH; p.Wf = m->wf_(m->wf_headT); p.Y = m->Gp[0]; p.ldy = C;
      TCN_CHECK(gemm(m, p, LDH, C, st, 4));
    }
    TCN_CHECK(hand_over());  // Gp[0..3] ready
    {   // lateral weights: sum over l = 0..2 of Gp_l^T f_l  (p_l = p_{l+1} + lat(f_l))
      WgradDev w = base_wgrad(m);
      w.meta = m->meta4; w.nblk = 3 * m->max_blk; w.dyn = m->desc3;
      w.G = m->Gp[0]; w.ldg = C; w.g_cols = C; w.X = f[0]; w.ldx = C;
      w.n_out = C; w.c_in = C; w.dW = m->g_(m->off_lat_w); w.db = m->g_(m->off_lat_b);
      TCN_CHECK(wgrad(m, w, x_rows, ws, 3, 1));
    }
  } else {
  const float* prev = nullptr;
  for (int lv = 0; lv < 4; ++lv) {
    WgradDev w = base_wgrad(m);
    w.G = gl[lv]; w.ldg = LDH; w.g_cols = LDH; w.X = plv[lv]; w.ldx = C;
    w.n_out = NH; w.c_in = C; w.dW = m->g_(m->off_head_w); w.db = m->g_(m->off_head_b);
    TCN_CHECK(wgrad(m, w, x_rows, ws));
    TapGemmDev p = base_tapgemm(m);
    p.X = gl[lv]; p.ldx = LDH; p.Wf = m->wf_(m->wf_headT); p.Y = m->Gp[lv]; p.ldy = C;
    p.R = prev; p.ldr = C;
    TCN_CHECK(gemm(m, p, LDH, C, st));
    prev = m->Gp[lv];
  }
  TCN_CHECK(hand_over());  // Gp[0..3] ready
  // lateral weight grads: p_l = p_{l+1} + lat(f_l) for l = 0..2
  for (int lv = 0; lv < 3; ++lv) {
    WgradDev w = base_wgrad(m);
    w.G = m->Gp[lv]; w.ldg = C; w.g_cols = C; w.X = f[lv]; w.ldx = C;
    w.n_out = C; w.c_in = C; w.dW = m->g_(m->off_lat_w); w.db = m->g_(m->off_lat_b);
    TCN_CHECK(wgrad(m, w, x_rows, ws));
  }
  }
  // stages, last to first
  const int tr = m->fwd_training ? 1 : 0;
  const bool multi = m->wg_multi && m->use_tc && C == 64 && m->max_blk <= 2 * num_sms();
  if (multi && !m->wgm[tr].ready) TCN_CHECK(build_wg_multi(m, tr));
  // per-layer weight gradients from one pass (wgrad_layer.cu): wl_lg layers per launch behind their input gradients,
  // slabs added in fixed order by ONE reduction launch at the end
  const bool wl = m->use_wl && !multi && m->fused_bwd && m->use_tc && C == 64;
  if (wl && !m->wl[tr].ready) TCN_CHECK(build_wl_table(m, tr));
  WgLayersLaunch wq;
  wq.meta = m->meta; wq.nblk = m->max_blk; wq.dyn = m->desc; wq.splits = m->wl_splits; wq.debug = 0;
  int wl_first = 0, wl_pending = 0;
  auto wl_flush = [&]() -> int {
    if (wl_pending == 0) return TCN_OK;
    TCN_CHECK(hand_over());  // gy and gu of the pending layers are final
    TCN_CHECK(launch_wgrad_layers(m->wl[tr].descs + wl_first, wl_pending, wq, ws));
    wl_first += wl_pending;
    wl_pending = 0;
    return TCN_OK;
  };
  const float* g = m->Gp[3];
  int gi = 0;
  for (int s = 3; s >= 0; --s) {
    for (int l = m->stage_first[s + 1] - 1; l >= m->stage_first[s]; --l) {
      int sh[3];
      layer_shifts(m, l, sh);
      float* gu = m->gus[l];
      const bool fused = m->fused_bwd && m->use_tc && C == 64;
      if (fused) {
        // one launch: gu = (gv W2) * [h > 0] recomputed per tap in tensor memory, gx = gy + sum_k W1_k^T gu[t - s_k];
        // gu of the tile itself is written once for the weight gradients (gemm_tc.cu: layer_bwd_tc_kernel)
        const long k2 = reinterpret_cast<const float*>(m->wf_(m->wf_w2T[l])) - m->wf;
        const long k1 = reinterpret_cast<const float*>(m->wf_(m->wf_w1T[l])) - m->wf;
        auto w2 = m->tcw.find(k2), w1 = m->tcw.find(k1);
        TCN_REQUIRE(w1 != m->tcw.end() && w2 != m->tcw.end(), "tcn_model backward: missing split weights for the fused layer");
        const auto gkey = std::make_pair(g, C);
        auto gm = m->xmaps.find(gkey);
        if (gm == m->xmaps.end()) {
          CUtensorMap map;
          TCN_CHECK(make_tensor_map_2d(&map, g, m->cfg.max_rows, C, C, TC_BM));
          gm = m->xmaps.emplace(gkey, map).first;
        }
        LayerBwdTcDev p;
        memset(&p, 0, sizeof(p));
        p.gu.Y = gu; p.gu.ldy = C; p.gu.N = C; p.gu.drop_scale = 1.f; p.gu.in_drop_scale = 1.f;
        p.gx.Y = m->gpool[gi]; p.gx.ldy = C; p.gx.N = C; p.gx.R = g; p.gx.ldr = C;
        p.gx.meta = m->meta; p.gx.nblk = m->max_blk; p.gx.dyn = m->desc; p.gx.ntaps = 3; p.gx.kbp = 2; p.gx.c_in = C;
        for (int i = 0; i < 3; ++i) p.gx.shift[i] = -sh[i];
        p.gx.drop_scale = 1.f; p.gx.in_drop_scale = 1.f;
        p.masks = m->masks[l];
        p.use_drop = pl > 0.f ? 1 : 0;
        p.drop_scale = pl > 0.f ? 1.f / (1.f - pl) : 1.f;
        TCN_CHECK(launch_layer_bwd_tc(gm->second, w2->second.mh, w2->second.ml, w1->second.mh, w1->second.ml, p, m->max_blk,
                                      st, true));
      } else {
        // gu = (gv W2) * [h > 0],   gv = keep * gy / (1 - p) applied as gy is loaded
        TapGemmDev p = base_tapgemm(m);
        p.X = g; p.ldx = C; p.Wf = m->wf_(m->wf_w2T[l]); p.Y = gu; p.ldy = C; p.M = m->H[l]; p.ldm = C;
        if (pl > 0.f) { p.in_drop_thresh = drop_thresh(pl); p.in_drop_scale = 1.f / (1.f - pl); p.in_drop_stream = (uint32_t)l; }
        TCN_CHECK(gemm(m, p, C, C, st));
      }
      if (wl) {
        if (++wl_pending == m->wl_lg) TCN_CHECK(wl_flush());
      } else if (!multi) {
      TCN_CHECK(hand_over());  // gy (= g) and gu of this layer are final
        WgradDev w2 = base_wgrad(m);
        w2.G = g; w2.ldg = C; w2.g_cols = C; w2.X = m->H[l]; w2.ldx = C; w2.n_out = C; w2.c_in = C;
        w2.dW = m->g_(m->off_w2[l]); w2.db = m->g_(m->off_b2[l]);
        if (pl > 0.f) { w2.g_drop_thresh = drop_thresh(pl); w2.g_drop_scale = 1.f / (1.f - pl); w2.g_drop_stream = (uint32_t)l; }
        WgradDev w1 = base_wgrad(m);
        w1.G = gu; w1.ldg = C; w1.g_cols = C; w1.X = m->act[l]; w1.ldx = C; w1.n_out = C; w1.c_in = C; w1.ntaps = 3;
        for (int i = 0; i < 3; ++i) w1.shift[i] = sh[i];
        w1.dW = m->g_(m->off_w1[l]); w1.db = m->g_(m->off_b1[l]);
        TCN_CHECK(wgrad_pair(m, w1, w2, x_rows, ws));
      }
      if (!fused) {  // gx = gy + sum_k W1_k^T gu[t - s_k]
        TapGemmDev p = base_tapgemm(m);
        p.X = gu; p.ldx = C; p.Wf = m->wf_(m->wf_w1T[l]); p.Y = m->gpool[gi]; p.ldy = C; p.R = g; p.ldr = C;
        p.ntaps = 3;
        for (int i = 0; i < 3; ++i) p.shift[i] = -sh[i];
        TCN_CHECK(gemm(m, p, C, C, st));
      }
      g = m->gpool[gi++];
    }
    if (multi) {  // every weight gradient of this stage in one launch, behind the stage's input-gradient chain
      const tcn_model::WgMulti& t = m->wgm[tr];
      const int si = 3 - s;
      TCN_CHECK(hand_over());
      if (t.tile_count[si] > 0)
        TCN_CHECK(launch_wgrad_tc_multi(t.descs, t.tiles + t.tile_first[si], t.tile_count[si], t.row_splits[si], ws));
    }
    if (s > 0) {  // f_{s-1} also feeds the lateral of level s-1
      TapGemmDev p = base_tapgemm(m);
      p.X = m->Gp[s - 1]; p.ldx = C; p.Wf = m->wf_(m->wf_latT); p.Y = m->gpool[gi]; p.ldy = C; p.R = g; p.ldr = C;
      TCN_CHECK(gemm(m, p, C, C, st));
      g = m->gpool[gi++];
    }
  }
  if (wl) {
    TCN_CHECK(wl_flush());
    TCN_CHECK(launch_wgrad_layers_reduce(m->wl[tr].outs, m->L, m->wl_splits, ws));
  }
  // projection weight grads (x is a leaf: no input gradient); same mask / channel scale as the forward
  {
    WgradDev w = base_wgrad(m);
    w.G = g; w.ldg = C; w.g_cols = C; w.X = x; w.ldx = D; w.x_unpadded = 1; w.n_out = C; w.c_in = D;
    w.dW = m->g_(m->off_proj_w); w.db = m->g_(m->off_proj_b);
    if (m->fwd_training && m->chan_drop_p > 0.f) { w.colscale = m->colscale; w.colscale_ld = D; }
    if (m->fwd_training && m->input_mask_p > 0.f) {
      w.x_drop_thresh = drop_thresh(m->input_mask_p); w.x_drop_scale = 1.f; w.x_drop_stream = kStreamMask;
    }
    TCN_CHECK(wgrad(m, w, x_rows, st, 1, 2));
  }
  if (ws != st) {  // join: the caller's stream owns every gradient again
    cudaEvent_t ev = m->evs[nev++];
    if (cudaEventRecord(ev, ws) != cudaSuccess || cudaStreamWaitEvent(st, ev, 0) != cudaSuccess) {
      set_error("tcn_model backward: join failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TCN_ERR_CUDA;
    }
  }
  return TCN_OK;
}

extern "C" int tcn_model_forward(tcn_model* m, const float* x, long long x_rows, int training, const float** feats,
                                 const float** logits, int* ld_logits, tcn_stream_t stream) {
  TCN_REQUIRE(m && x && m->params && x_rows > 0, "tcn_model_forward: null pointer / parameters not bound");
  // training: 0 = inference, 1 = train (dropout on, activations kept for backward), 2 = keep activations, no dropout
  TCN_CHECK(model_forward(m, x, x_rows, training == 1, training != 0, (cudaStream_t)stream));
  if (feats) {
    for (int lv = 0; lv < 3; ++lv) feats[lv] = m->P[lv];
    feats[3] = m->act[m->stage_first[4]];
  }
  if (logits)
    for (int lv = 0; lv < 4; ++lv) logits[lv] = m->logits[lv];
  if (ld_logits) *ld_logits = m->LDH;
  return TCN_OK;
}

extern "C" int tcn_model_backward(tcn_model* m, const float* x, long long x_rows, const float* const* glogits,
                                  const float* const* gfeats, tcn_stream_t stream) {
  TCN_REQUIRE(m && x && m->params, "tcn_model_backward: null pointer / parameters not bound");
  return model_backward(m, x, x_rows, glogits, gfeats, (cudaStream_t)stream);
}

__global__ void finish_loss_kernel(float* loss8, float* out, float w0, float w1, float w2, float w3) {
  if (threadIdx.x == 0) {
    const float a = loss8[0], b = loss8[1], c = loss8[2], d = loss8[3];
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
    out[4] = w0 * a + w1 * b + w2 * c + w3 * d;
    out[5] = out[6] = out[7] = 0.f;
  }
}

extern "C" int tcn_model_train_step(tcn_model* m, const float* x, long long x_rows, const unsigned char* labels, int ldlab,
                                    int training, float* loss_out, tcn_stream_t stream) {
  TCN_REQUIRE(m && x && labels && loss_out && m->params, "tcn_model_train_step: null pointer / parameters not bound");
  TCN_REQUIRE(ldlab >= m->NH, "tcn_model_train_step: labels need %d columns", m->NH);
  cudaStream_t st = (cudaStream_t)stream;
  TCN_CHECK(model_forward(m, x, x_rows, training, true, st));
  if (cudaMemsetAsync(m->loss8, 0, 32, st) != cudaSuccess) {
    set_error("tcn_model_train_step: memset failed");
    cudaGetLastError();
    return TCN_ERR_CUDA;
  }
  {  // the four FPN levels in one launch (same labels, weights and layout; blockIdx.y = level)
    BceDev b;
    memset(&b, 0, sizeof(b));
    b.ldl = m->LDH; b.labels = labels; b.ldlab = ldlab; b.lab_unpadded = 1;
    b.meta = m->meta; b.nrows = m->cfg.max_rows; b.dyn = m->desc; b.ncols = m->NH; b.zero_cols = m->LDH;
    b.pos_w = m->has_pos_w ? m->pos_w : nullptr; b.col_scale = m->col_scale; b.col_unit = m->col_unit;
    b.col_head = m->col_head; b.row_scale_const = 1.f; b.loss = m->loss8; b.lddl = m->LDH;
    b.grad_scale = 1.f;
    b.nlev = 4;
    for (int lv = 0; lv < 4; ++lv) {
      b.logits_lv[lv] = m->logits[lv]; b.dL_lv[lv] = m->dL[lv];
      b.cum_lv[lv] = m->stack_levels ? m->cum[lv] : nullptr;
    }
    TCN_CHECK(launch_bce(b, m->cfg.max_rows, st));
  }
  finish_loss_kernel<<<1, 32, 0, st>>>(m->loss8, loss_out, m->head_w[0], m->head_w[1], m->head_w[2], m->head_w[3]);
  TCN_CHECK(check_launch("finish_loss_kernel"));
  const float* gl[4] = {m->dL[0], m->dL[1], m->dL[2], m->dL[3]};
  return model_backward(m, x, x_rows, gl, nullptr, st, m->stack_levels);
}
