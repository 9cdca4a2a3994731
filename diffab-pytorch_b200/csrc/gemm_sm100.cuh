// bf16 x bf16 -> fp32 GEMM on tcgen05 tensor cores with TMA-staged operands (sm_100a).
//
//   C[M, N] (fp32, row-major, ldc) = A[M, K] (bf16, K contiguous) * B[N, K]^T (bf16, K contiguous) + bias[N]
//
// Used for the IPA projections (K = 128) and to_out (K = 1024), i.e. the nn.Linear layers of
// diffab_pytorch.py:391-403,464.  One CTA computes a 128 x BN tile: one lane issues the TMA loads
// (128B-swizzled 64-wide K chunks, ring of kStages), a lane of another warp the tcgen05.mma chain (M = 128,
// N = BN, K = 16 per instruction, fp32 accumulator in TMEM); all four warps then read the accumulator with tcgen05.ld
// (warp w owns TMEM lanes 32w..32w+31 = tile rows), add the bias and store.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;   // bf16 elements per K chunk = one 128-byte swizzle atom row
constexpr int kGemmStages = 3;   // default ring depth (template parameter STAGES)

template <int BN, int STAGES = kGemmStages>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int STAGES = kGemmStages>
__global__ void __launch_bounds__(128) gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                        const __grid_constant__ CUtensorMap map_b,
                                                        float* __restrict__ C, int64_t ldc,
                                                        const float* __restrict__ bias, int K,
                                                        __nv_bfloat16* __restrict__ C16 = nullptr) {
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "BN");
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = GemmSmem<BN, STAGES>;
  constexpr int kStages = STAGES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* done = empty + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kGemmBM, n0 = blockIdx.x * BN;
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
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
  if (warp == 0) {
    // ---- TMA producer (whole warp walks the loop, one elected lane issues): keeps the whole ring in flight, refilling a
    // slot as soon as its MMAs have read it
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kStages;
      if (kc >= kStages) mbar_wait(&empty[s], ((kc / kStages) - 1) & 1);
      if (elect_one()) {
        uint8_t* a = smem + s * S::kStageBytes;
        mbar_arrive_expect_tx(&full[s], S::kStageBytes);
        tma_load_2d(a, &map_a, &full[s], kc * kGemmBK, m0);
        tma_load_2d(a + S::kABytes, &map_b, &full[s], kc * kGemmBK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- tcgen05.mma issuer (a different warp, so loads and MMAs never wait for each other's bookkeeping)
    constexpr uint32_t idesc = make_idesc_bf16(kGemmBM, BN, 0, 0);
    const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024, kSwizzle128B);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kStages;
      mbar_wait(&full[s], (kc / kStages) & 1);
      tcgen05_fence_after_sync();
      if (elect_one()) {
        // K-major, 128B swizzle: 8-row groups are 1024 B apart (SBO); a K step of 16 bf16 is +32 B
        const uint64_t da = d0 + (uint32_t)((s * S::kStageBytes) >> 4);
        const uint64_t db = da + (uint32_t)(S::kABytes >> 4);
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k)
          umma_bf16(tmem_base, da + (uint32_t)((k * 32) >> 4), db + (uint32_t)((k * 32) >> 4), idesc, (kc | k) != 0);
        umma_commit(&empty[s]);  // slot reusable once these MMAs have read it
        if (kc == nk - 1) umma_commit(done);
      }
      __syncwarp();
    }
  }
  __syncwarp();
  mbar_wait(done, 0);
  tcgen05_fence_after_sync();

  // epilogue: thread (warp, lane) owns tile row 32*warp + lane
  const int row = m0 + warp * 32 + lane;
  float* crow = C + (int64_t)row * ldc + n0;
  __nv_bfloat16* crow16 = C16 ? C16 + (int64_t)row * ldc + n0 : nullptr;
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 16) {
    float v[16];
    tmem_ld_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    if (crow16) {      // bf16 output (round-to-nearest-even, what the consumer's own fp32 -> bf16 conversion would give)
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = v[i] + (bias ? __ldg(bias + n0 + c0 + i) : 0.f);
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
    for (int q = 0; q < 4; ++q) {
      float4 o;
      o.x = v[4 * q] + (bias ? __ldg(bias + n0 + c0 + 4 * q) : 0.f);
      o.y = v[4 * q + 1] + (bias ? __ldg(bias + n0 + c0 + 4 * q + 1) : 0.f);
      o.z = v[4 * q + 2] + (bias ? __ldg(bias + n0 + c0 + 4 * q + 2) : 0.f);
      o.w = v[4 * q + 3] + (bias ? __ldg(bias + n0 + c0 + 4 * q + 3) : 0.f);
      *reinterpret_cast<float4*>(crow + c0 + 4 * q) = o;
    }
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem_base, kTmemCols);
}

// Host launcher.  A: [M, K] bf16 row-major (lda elements), B: [N, K] bf16 row-major (ldb), C fp32 - or, when C_bf16 is
// given, bf16 with the same leading dimension -, M % 128 == 0,
// N % BN == 0, K % 64 == 0.
template <int BN, int STAGES = kGemmStages>
int launch_gemm_bf16(const void* A, int64_t lda, const void* Bm, int64_t ldb, float* C, int64_t ldc, const float* bias,
                     int M, int N, int K, cudaStream_t stream, void* C_bf16 = nullptr) {
  DAB_REQUIRE(M % kGemmBM == 0 && N % BN == 0 && K % kGemmBK == 0 && K > 0, DAB_EUNSUPPORTED,
              "gemm_bf16: M %% 128, N %% %d, K %% 64 required (M=%d N=%d K=%d)", BN, M, N, K);
  CUtensorMap ma, mb;
  uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M}, str_a[1] = {(uint64_t)lda * 2};
  uint64_t dims_b[2] = {(uint64_t)K, (uint64_t)N}, str_b[1] = {(uint64_t)ldb * 2};
  uint32_t box_a[2] = {kGemmBK, kGemmBM}, box_b[2] = {kGemmBK, (uint32_t)BN};
  if (int rc = make_tensor_map_bf16(&ma, A, 2, dims_a, str_a, box_a, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_tensor_map_bf16(&mb, Bm, 2, dims_b, str_b, box_b, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  DAB_ENSURE_SMEM((gemm_bf16_kernel<BN, STAGES>), (GemmSmem<BN, STAGES>::kTotal));
  dim3 grid(N / BN, M / kGemmBM);
  gemm_bf16_kernel<BN, STAGES><<<grid, 128, GemmSmem<BN, STAGES>::kTotal, stream>>>(ma, mb, C, ldc, bias, K,
                                                                     reinterpret_cast<__nv_bfloat16*>(C_bf16));
  count_launch();
  return check_launch("gemm_bf16");
}

}  // namespace sm100
}  // namespace dab
