// tcgen05 + TMA GEMM for per-frame dense contractions with a long K:
//
//     Y[r, n] = sum_k X[r, k] * W[n, k] + bias[n]                (fp32 in, fp32 out)
//
// the stage-input projection Conv1d(dim -> num_f_maps, 1) of BaseCausalTCN
// (MT4MTLKD/Temporal_tenco/network.py:113,129; x arrives as (B, T, D) = K-major rows, :42) and every
// nn.Linear / 1x1 Conv1d of the MS-TCT blocks (Temporal_mstct/MSTCT/Temporal_Encoder.py:12,15,57-59).
//
// Blackwell structure (one CTA = one 128-frame x BN tile, 6 warps):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D loads of the X tile (128 x 32 fp32, 128B swizzle)
//              and of the pre-split weight tiles W_hi / W_lo (BN x 32) into a 4-stage smem ring
//   warps 2-5: operand split -- X -> (X_hi in place, X_lo) : hi = x & 0xffffe000 is exact in TF32, lo = x - hi;
//              also folds Dropout2d's per-(sequence, channel) scale and the 25 % input mask into the load
//   warp 1   : MMA issuer    -- one elected lane issues tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8), three
//              products per k-slice (lo*hi + hi*lo + hi*hi) accumulating in TMEM (fp32); tcgen05.commit
//              releases the smem stage back to the producer
//   warps 2-5: epilogue      -- tcgen05.ld (32 lanes x 32 columns) -> + bias -> 128-bit stores of Y
// Accuracy: 3xTF32 == fp32 to ~1e-6 relative (see tests), the bar is 1e-3 on logits.
// HBM traffic: X is read exactly once (the algorithmic 4*(D + C) bytes per frame); W tiles come from L2.
#include "gemm_tc.cuh"

namespace tcn {

constexpr int TC_BM = 128;       // frames per tile
constexpr int TC_BK = 32;        // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;  // 6 warps

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)0 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN>
struct TcSmem {
  static constexpr int kStages = BN > 64 ? 3 : 4;
  static constexpr int kA = TC_BM * TC_BK * 4;  // 16384
  static constexpr int kB = BN * TC_BK * 4;
  static constexpr int kStage = 2 * kA + 2 * kB;
  static constexpr int kBytes = kStages * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
               const __grid_constant__ CUtensorMap map_wlo, const GemmTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  using S = TcSmem<BN>;
  constexpr int TC_STAGES = S::kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = blockIdx.x, ntile = blockIdx.y;
  const int nblk = p.dyn ? p.dyn->nblk : p.nblk;

  // carve shared memory (tiles need 1024-byte alignment for the 128B swizzle)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_STAGES * S::kStage);
  uint64_t* full_bar = bars;                    // TMA bytes landed          (count 1 + tx)
  uint64_t* ready_bar = bars + TC_STAGES;       // operands split            (count 128)
  uint64_t* empty_bar = bars + 2 * TC_STAGES;   // MMAs that read the stage retired (count 1, tcgen05.commit)
  uint64_t* accum_bar = bars + 3 * TC_STAGES;   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {  // TMEM: BN fp32 accumulator columns (power of two >= 32)
    constexpr uint32_t kCols = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool active = blk < nblk;
  BlkMeta m = {0, 0, 0, 0};
  if (active) m = p.meta[blk];
  const int row0 = blk * kBlkRows;
  const bool has_rows = active && row0 < m.hi;
  const int kblocks = p.K / TC_BK;

  if (has_rows) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const int xrow = row0 + (p.x_unpadded ? m.in_delta : 0);
        for (int kb = 0; kb < kblocks; ++kb) {
          const int s = kb % TC_STAGES;
          const uint32_t ph = (kb / TC_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = tiles + s * S::kStage;
          mbar_arrive_expect_tx(&full_bar[s], S::kA + 2 * S::kB);
          tma_load_2d(st, &map_x, &full_bar[s], kb * TC_BK, xrow);
          tma_load_2d(st + 2 * S::kA, &map_whi, &full_bar[s], kb * TC_BK, ntile * BN);
          tma_load_2d(st + 2 * S::kA + S::kB, &map_wlo, &full_bar[s], kb * TC_BK, ntile * BN);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&ready_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = base + s * S::kStage, a_lo = a_hi + S::kA;
          const uint32_t b_hi = a_hi + 2 * S::kA, b_lo = b_hi + S::kB;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint32_t ko = k * 32;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t dah = umma_desc_sw128(a_hi + ko), dal = umma_desc_sw128(a_lo + ko);
            const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
            umma_tf32(tmem_base, dal, dbh, idesc, (kb | k) != 0);
            umma_tf32(tmem_base, dah, dbl, idesc, 1u);
            umma_tf32(tmem_base, dah, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs retire
          if (kb == kblocks - 1) umma_commit(accum_bar);    // accumulator complete
        }
        __syncwarp();
      }
    } else {
      // ===================== operand split (warps 2..5) =====================
      const int ct = threadIdx.x - 64;  // 0..127
      const uint32_t in_seed = p.in_drop_seed ^ (p.dyn ? p.dyn->seed : 0u);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        float4* xa = reinterpret_cast<float4*>(tiles + s * S::kStage);
        float4* xl = reinterpret_cast<float4*>(tiles + s * S::kStage + S::kA);
#pragma unroll
        for (int i = 0; i < (TC_BM * TC_BK / 4) / 128; ++i) {
          const int c = ct + i * 128;  // physical 16-byte chunk inside the tile
          float4 v = xa[c];
          if (p.colscale != nullptr || p.in_drop_thresh != 0u) {
            const int r = c >> 3;
            const int col = kb * TC_BK + (((c & 7) ^ (r & 7)) << 2);  // undo the 128B swizzle
            if (p.colscale != nullptr) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(p.colscale + (size_t)m.seq * p.colscale_ld + col));
              v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
            }
            if (p.in_drop_thresh != 0u) {
              const int row = row0 + r;
              v.x *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, row, col);
              v.y *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, row, col + 1);
              v.z *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, row, col + 2);
              v.w *= drop_factor(in_seed, p.in_drop_stream, p.in_drop_thresh, p.in_drop_scale, row, col + 3);
            }
          }
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
          xa[c] = h;
          xl[c] = l;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready_bar[s]);
      }
      // ===================== epilogue (same warps; TMEM lane quadrant = warp % 4) =====================
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int row = row0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (row < m.hi) {
          const int n0 = ntile * BN + c0;
          float* yp = p.Y + (size_t)row * p.ldy + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (p.bias != nullptr) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            *reinterpret_cast<float4*>(yp + j) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    constexpr uint32_t kCols = BN < 32 ? 32 : BN;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kCols));
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      cudaGetLastError();
  }
  return fn;
}

// fp32 row-major (rows x cols, leading dimension ld floats) -> 2D map with a (32 x box_rows) box, 128B swizzle
int make_tensor_map_2d(CUtensorMap* map, const float* ptr, long rows, long cols, long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return TCN_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr %p rows %ld cols %ld ld %ld", (int)r, (const void*)ptr, rows,
              cols, ld);
    return TCN_ERR_CUDA;
  }
  return TCN_OK;
}

__global__ void split_weight_kernel(const float* __restrict__ w, float* __restrict__ whi, float* __restrict__ wlo,
                                    long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    whi[i] = h;
    wlo[i] = v - h;
  }
}

int launch_split_weight(const float* w, float* whi, float* wlo, long n, cudaStream_t stream) {
  long b = (n + 255) / 256;
  if (b > 1024) b = 1024;
  split_weight_kernel<<<(int)b, 256, 0, stream>>>(w, whi, wlo, n);
  return check_launch("split_weight_kernel");
}

int gemm_tc_box_rows_for_n(int n) { return (n % 128 == 0) ? 128 : 64; }

int launch_gemm_tc(const CUtensorMap& mx, const CUtensorMap& mwhi, const CUtensorMap& mwlo, const GemmTcDev& p,
                   int cap_nblk, cudaStream_t stream) {
  const int nb = cap_nblk > 0 ? cap_nblk : p.nblk;
  const int bn = (p.N % 128 == 0) ? 128 : 64;
  dim3 grid(nb, p.N / bn);
  cudaError_t e;
  if (bn == 128) {
    static bool set128 = false;
    if (!set128) {
      e = cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::kBytes);
      if (e != cudaSuccess) { set_error("gemm_tc<128>: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return TCN_ERR_CUDA; }
      set128 = true;
    }
    gemm_tc_kernel<128><<<grid, TC_THREADS, TcSmem<128>::kBytes, stream>>>(mx, mwhi, mwlo, p);
  } else {
    static bool set64 = false;
    if (!set64) {
      e = cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64>::kBytes);
      if (e != cudaSuccess) { set_error("gemm_tc<64>: smem attribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return TCN_ERR_CUDA; }
      set64 = true;
    }
    gemm_tc_kernel<64><<<grid, TC_THREADS, TcSmem<64>::kBytes, stream>>>(mx, mwhi, mwlo, p);
  }
  return check_launch("gemm_tc_kernel");
}

}  // namespace tcn

using namespace tcn;

extern "C" int tcn_gemm_tc_supported(int k, int n) { return (k > 0 && k % TC_BK == 0 && n > 0 && n % 64 == 0) ? 1 : 0; }

extern "C" int tcn_gemm_tc(const tcn_gemm_tc_args* a, tcn_stream_t stream) {
  TCN_REQUIRE(a && a->x && a->w_hi && a->w_lo && a->y && a->meta, "tcn_gemm_tc: null pointer");
  if (!tcn_gemm_tc_supported(a->k, a->n)) {
    set_error("tcn_gemm_tc: needs k %% 32 == 0 and n %% 64 == 0 (got k=%d n=%d); use tcn_tapgemm", a->k, a->n);
    return TCN_ERR_UNSUPPORTED;
  }
  TCN_REQUIRE(a->nblk > 0 && a->x_rows > 0 && a->ldx >= a->k && a->ldx % 4 == 0 && a->ldy >= a->n && a->ldy % 4 == 0,
              "tcn_gemm_tc: bad shape");
  TCN_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->w_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w_lo) & 15) == 0,
              "tcn_gemm_tc: pointers must be 16-byte aligned");
  TCN_REQUIRE(a->in_drop_p >= 0.f && a->in_drop_p < 1.f, "tcn_gemm_tc: in_drop_p must be in [0, 1)");
  const int bn = (a->n % 128 == 0) ? 128 : 64;
  CUtensorMap mx, mh, ml;
  TCN_CHECK(make_tensor_map_2d(&mx, a->x, a->x_rows, a->k, a->ldx, TC_BM));
  TCN_CHECK(make_tensor_map_2d(&mh, a->w_hi, a->n, a->k, a->k, bn));
  TCN_CHECK(make_tensor_map_2d(&ml, a->w_lo, a->n, a->k, a->k, bn));
  GemmTcDev p;
  p.Y = a->y; p.ldy = a->ldy; p.bias = a->bias;
  p.meta = reinterpret_cast<const BlkMeta*>(a->meta); p.nblk = a->nblk; p.dyn = nullptr;
  p.x_unpadded = a->x_unpadded; p.K = a->k; p.N = a->n;
  p.colscale = a->colscale; p.colscale_ld = a->colscale_ld;
  p.in_drop_thresh = a->in_drop_p > 0.f ? drop_thresh(a->in_drop_p) : 0u;
  p.in_drop_scale = a->in_drop_rescale ? 1.f / (1.f - a->in_drop_p) : 1.f;
  p.in_drop_seed = a->drop_seed; p.in_drop_stream = a->drop_stream;
  return launch_gemm_tc(mx, mh, ml, p, 0, (cudaStream_t)stream);
}

extern "C" int tcn_split_weight(const float* w, float* w_hi, float* w_lo, long long n, tcn_stream_t stream) {
  TCN_REQUIRE(w && w_hi && w_lo && n > 0, "tcn_split_weight: bad arguments");
  return launch_split_weight(w, w_hi, w_lo, (long)n, (cudaStream_t)stream);
}
