// bf16 x bf16 -> fp32 GEMM on tcgen05 tensor cores with TMA-staged operands (sm_100a).
//
//   C[M, N] (fp32, row-major, ldc) = A[M, K] (bf16, K contiguous) * B[N, K]^T (bf16, K contiguous) + bias[N]
//
// Used for to_out (K = 1024), the front MLP and the data-gradient GEMMs of the backward, i.e. the nn.Linear layers of
// diffab_pytorch.py:391-403,464,572-574.  Persistent CTAs (one per SM) walk over 128 x BN output tiles:
//   warp 0    TMA producer: 128B-swizzled 64-wide K chunks through a ring of STAGES slots that runs across tiles
//   warp 1    tcgen05.mma issuer (the whole warp walks the loop, one elected lane issues; M = 128, N = BN, K = 16 per
//             instruction) into one of TWO TMEM accumulators, so the next tile's MMAs run under this tile's epilogue
//   warps 2-5 epilogue: tcgen05.ld (thread = tile row), bias, store (fp32 or bf16)
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;   // bf16 elements per K chunk = one 128-byte swizzle atom row
constexpr int kGemmStages = 0;   // 0 = pick the ring depth from BN (about 190 KB of shared memory)
constexpr int kGemmThreads = 192;

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = STAGES > 0 ? STAGES : (BN >= 128 ? 6 : 8);
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

template <int BN, int STAGES = kGemmStages>
__global__ void __launch_bounds__(kGemmThreads) gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b,
                                                                 float* __restrict__ C, int64_t ldc,
                                                                 const float* __restrict__ bias, int K, int n_nt, int n_items,
                                                                 __nv_bfloat16* __restrict__ C16 = nullptr, int relu = 0) {
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "BN");
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = GemmSmem<BN, STAGES>;
  constexpr int kStages = S::kStages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2], 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kAccCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  constexpr uint32_t kTmemCols = 2 * kAccCols;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int nk = K / kGemmBK;
  const int n_mine = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items blockIdx.x + q * gridDim.x
  if (warp == 0) {
    // ---- TMA producer: keeps the whole ring in flight, refilling a slot as soon as its MMAs have read it
    for (int q = 0, g = 0; q < n_mine; ++q) {
      const int item = (int)blockIdx.x + q * (int)gridDim.x;
      const int m0 = (item / n_nt) * kGemmBM, n0 = (item % n_nt) * BN;
      for (int kc = 0; kc < nk; ++kc, ++g) {
        const int s = g % kStages;
        if (g >= kStages) mbar_wait(&empty[s], ((g / kStages) - 1) & 1);
        if (elect_one()) {
          uint8_t* a = smem + s * S::kStageBytes;
          mbar_arrive_expect_tx(&full[s], S::kStageBytes);
          tma_load_2d(a, &map_a, &full[s], kc * kGemmBK, m0);
          tma_load_2d(a + S::kABytes, &map_b, &full[s], kc * kGemmBK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- tcgen05.mma issuer
    constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN, 0, 0);
    const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024, kSwizzle128B);
    for (int q = 0, g = 0; q < n_mine; ++q) {
      const int acc = q & 1;
      if (q >= 2) {                                   // the epilogue has drained this accumulator (tile q - 2)
        mbar_wait(&acc_empty[acc], ((q >> 1) - 1) & 1);
        tcgen05_fence_after_sync();
      }
      for (int kc = 0; kc < nk; ++kc, ++g) {
        const int s = g % kStages;
        mbar_wait(&full[s], (g / kStages) & 1);
        tcgen05_fence_after_sync();
        if (elect_one()) {
          // K-major, 128B swizzle: 8-row groups are 1024 B apart (SBO); a K step of 16 bf16 is +32 B
          const uint64_t da = d0 + (uint32_t)((s * S::kStageBytes) >> 4);
          const uint64_t db = da + (uint32_t)(S::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_bf16(tmem_base + acc * kAccCols, da + (uint32_t)((k * 32) >> 4), db + (uint32_t)((k * 32) >> 4), idesc,
                      (kc | k) != 0);
          umma_commit(&empty[s]);  // slot reusable once these MMAs have read it
          if (kc == nk - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 = tile rows; thread = row
    const int quarter = warp & 3;
    for (int q = 0; q < n_mine; ++q) {
      const int item = (int)blockIdx.x + q * (int)gridDim.x;
      const int m0 = (item / n_nt) * kGemmBM, n0 = (item % n_nt) * BN;
      const int acc = q & 1;
      mbar_wait(&acc_full[acc], (q >> 1) & 1);
      tcgen05_fence_after_sync();
      const int row = m0 + quarter * 32 + lane;
      float* crow = C + (int64_t)row * ldc + n0;
      __nv_bfloat16* crow16 = C16 ? C16 + (int64_t)row * ldc + n0 : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kAccCols;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(taddr + c0, v);
        tmem_wait_ld();
        if (crow16) {      // bf16 output (round-to-nearest-even, what the consumer's own fp32 -> bf16 conversion would give)
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            o[i] = v[i] + (bias ? __ldg(bias + n0 + c0 + i) : 0.f);
            if (relu) o[i] = fmaxf(o[i], 0.f);
          }
          uint4* dst = reinterpret_cast<uint4*>(crow16 + c0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[8 * h], o[8 * h + 1]), p1 = __floats2bfloat162_rn(o[8 * h + 2], o[8 * h + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[8 * h + 4], o[8 * h + 5]), p3 = __floats2bfloat162_rn(o[8 * h + 6], o[8 * h + 7]);
            dst[h] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
          }
          continue;
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float4 o;
          o.x = v[4 * qq] + (bias ? __ldg(bias + n0 + c0 + 4 * qq) : 0.f);
          o.y = v[4 * qq + 1] + (bias ? __ldg(bias + n0 + c0 + 4 * qq + 1) : 0.f);
          o.z = v[4 * qq + 2] + (bias ? __ldg(bias + n0 + c0 + 4 * qq + 2) : 0.f);
          o.w = v[4 * qq + 3] + (bias ? __ldg(bias + n0 + c0 + 4 * qq + 3) : 0.f);
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(crow + c0 + 4 * qq) = o;
        }
      }
      tcgen05_fence_before_sync();
      mbar_arrive(&acc_empty[acc]);
    }
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem_base, kTmemCols);
}

// Host launcher.  A: [M, K] bf16 row-major (lda elements), B: [N, K] bf16 row-major (ldb), C fp32 - or, when C_bf16 is
// given, bf16 with the same leading dimension - (+ bias, ReLU when `relu`), M % 128 == 0,
// N % BN == 0, K % 64 == 0.
template <int BN, int STAGES = kGemmStages>
int launch_gemm_bf16(const void* A, int64_t lda, const void* Bm, int64_t ldb, float* C, int64_t ldc, const float* bias,
                     int M, int N, int K, cudaStream_t stream, void* C_bf16 = nullptr, int relu = 0) {
  DAB_REQUIRE(M % kGemmBM == 0 && N % BN == 0 && K % kGemmBK == 0 && K > 0, DAB_EUNSUPPORTED,
              "gemm_bf16: M %% 128, N %% %d, K %% 64 required (M=%d N=%d K=%d)", BN, M, N, K);
  if (M == 0 || N == 0) return DAB_OK;
  CUtensorMap ma, mb;
  uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M}, str_a[1] = {(uint64_t)lda * 2};
  uint64_t dims_b[2] = {(uint64_t)K, (uint64_t)N}, str_b[1] = {(uint64_t)ldb * 2};
  uint32_t box_a[2] = {kGemmBK, kGemmBM}, box_b[2] = {kGemmBK, (uint32_t)BN};
  if (int rc = make_tensor_map_bf16(&ma, A, 2, dims_a, str_a, box_a, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_tensor_map_bf16(&mb, Bm, 2, dims_b, str_b, box_b, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  DAB_ENSURE_SMEM((gemm_bf16_kernel<BN, STAGES>), (GemmSmem<BN, STAGES>::kTotal));
  const int n_nt = N / BN, n_items = (M / kGemmBM) * n_nt;
  int n_sm = 148;
  {
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev_id);
  }
  const int grid = n_items < n_sm ? n_items : n_sm;
  gemm_bf16_kernel<BN, STAGES><<<grid, kGemmThreads, GemmSmem<BN, STAGES>::kTotal, stream>>>(
      ma, mb, C, ldc, bias, K, n_nt, n_items, reinterpret_cast<__nv_bfloat16*>(C_bf16), relu);
  count_launch();
  return check_launch("gemm_bf16");
}

}  // namespace sm100
}  // namespace dab
