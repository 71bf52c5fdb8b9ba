// Declarations of the per-layer weight-gradient kernel (wgrad_layer.cu) shared with the executor (model.cu).
#pragma once
#include "gemm_tc.cuh"

namespace tcn {

constexpr int WL_RC = 16;                              // frames per pipeline stage (TMA box rows)
constexpr int WL_PART_FLOATS = 4 * 64 * 64 + 2 * 64;   // one CTA's partial: gW1 taps 0..2, gW2 as [c][n]; gb1, gb2

// one residual layer: operands (TMA maps with WL_RC-row boxes and 32-byte swizzle atoms) and its partial slabs
struct alignas(64) WgLayerDev {
  CUtensorMap mx, mh, mgu, mgy;   // layer input x, h = relu(u), gu, gy
  const uint32_t* masks;          // (rows, 4) bit words of layer_fwd_tc_kernel (dropout keep bits in words 2, 3) or nullptr
  float* part;                    // [splits][WL_PART_FLOATS]
  int shift[3];                   // forward taps s_k
  int use_drop;                   // 0: gv = gy
  float drop_scale;               // 1 / (1 - p)
  uint32_t drop_thresh, drop_seed, drop_stream;   // used when masks == nullptr: the mask is regenerated from the key
};
struct WgLayersLaunch {
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int splits;
  int debug;   // experiments only (TCN_WL_DEBUG): 1 = no MMAs, 2 = no operand split, 4 = load 4 of the 12 boxes
};
struct WgLayerOut {   // where the reduction adds a layer's partials
  const float* part;
  float *dw1, *db1, *dw2, *db2;
};

void wgrad_layers_plan(int cap_nblk, int* layers_per_launch, int* splits);
int make_wgrad_layer_maps(WgLayerDev* d, const float* x, const float* h, const float* gu, const float* gy, long rows);
int launch_wgrad_layers(const WgLayerDev* descs_dev, int nlayers, const WgLayersLaunch& q, cudaStream_t stream);
int launch_wgrad_layers_reduce(const WgLayerOut* outs_dev, int nlayers, int splits, cudaStream_t stream);

}  // namespace tcn
