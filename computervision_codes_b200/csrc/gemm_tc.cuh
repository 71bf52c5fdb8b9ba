// Declarations shared between gemm_tc.cu and the whole-model executor.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tcn {

struct GemmTcDev {
  float* Y;
  int ldy;
  const float* bias;
  const BlkMeta* meta;
  int nblk;
  const BatchDesc* dyn;
  int x_unpadded;
  int K;                    // multiple of 32
  int N;                    // multiple of 64
  const float* colscale;    // optional per-(sequence, k) scale of X
  int colscale_ld;
  uint32_t in_drop_thresh;  // optional keep-mask on X elements (row = padded row, col = k)
  float in_drop_scale;
  uint32_t in_drop_seed, in_drop_stream;
};

int make_tensor_map_2d(CUtensorMap* map, const float* ptr, long rows, long cols, long ld, int box_rows);
int launch_split_weight(const float* w, float* whi, float* wlo, long n, cudaStream_t stream);
int launch_gemm_tc(const CUtensorMap& mx, const CUtensorMap& mwhi, const CUtensorMap& mwlo, const GemmTcDev& p,
                   int cap_nblk, cudaStream_t stream);
int gemm_tc_box_rows_for_n(int n);

}  // namespace tcn
