// One layer of PairEmbedding's MLPs backward, in ONE pass over the B*L*L pairs (tcgen05, sm_100a).
//
// The layer is y = a W^T + b on 64 channels (diffab_pytorch.py:214-223 distance_embedding, :303-311 mlp of the reference),
// a = ReLU output of the layer before it.  Given g = dL/dy (bf16, [P, 64]) this kernel produces everything autograd
// would, reading g and a once and writing the next gradient once (3 x 128 B per pair instead of ~9 x 128 B for
// ReLU-backward + data-gradient GEMM + weight-gradient GEMM + bias reduction as separate passes):
//
//   dW[out, in]   += sum_p g[p, out] a[p, in]          wgrad : M = 64 out, N = 64 in (+8: see db), K = 128 pairs per tile,
//                                                              both operands read MN-major from the tiles as TMA wrote them
//   db[out]       += sum_p g[p, out]                   = 8 extra accumulator columns of the same MMAs: the B operand's second
//                                                        64-wide atom is a constant tile of ones (N = 72)
//   g_prev[p, in]  = (sum_out g[p, out] W[out, in]) * (a[p, in] > 0)     dgrad : M = 128 pairs, N = 64 in, K = 64 out
//   db_prev[in]   += sum_p g_prev[p, in]               (optional: the bias gradient of the layer before, in registers)
//
// `valid` (optional, residue mask [B, L]): pairs with an invalid residue do not count towards db (their rows of g are
// subtracted again; such rows are rare) - their rows of a are zero by construction (dab_pair_zero_masked), which already
// removes them from dW and g_prev.
//
// Persistent CTAs (one per SM) walk over 128-pair tiles:
//   warp 0   TMA producer: g tile + a tile (16 KB each, 128B-swizzled), ring of 4
//   warp 1   tcgen05.mma issuer (warp-convergent, one elected lane): dgrad into one of two TMEM accumulators, wgrad into a
//            third that lives for the whole kernel
//   warps 2-5  epilogue, thread = pair: sign mask of its a row (from shared memory, so the stage can be released before
//            the MMAs retire), accumulator -> mask -> bf16 -> swizzled staging tile -> TMA store
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int kMbStages = 4;
constexpr int kMbThreads = 192;

struct MbSmem {
  static constexpr int kTile = 16384;                       // [128 pairs][64 channels] bf16, 128B swizzle
  static constexpr int kStage = 2 * kTile;                  // g, a
  static constexpr int kStaging = kMbStages * kStage;       // 2 output tiles
  static constexpr int kWt = kStaging + 2 * kTile;          // W^T [64 in][64 out] bf16, K-major 128B-swizzled
  static constexpr int kOnes = kWt + 8192;                  // constant tile of ones (second N atom of the wgrad B operand)
  static constexpr int kCorr = kOnes + kTile;               // 64 floats: rows of g that must not count towards db
  static constexpr int kBars = kCorr + 256;
  static constexpr int kTmemSlot = kBars + 128;
  static constexpr int kTotal = kTmemSlot + 16 + 1024 /* alignment slack */;
};
enum MbBar { MB_FULL = 0, MB_EMPTY = 4, MB_ACC_FULL = 8, MB_ACC_EMPTY = 10, MB_FINAL = 12, MB_N_BARS = 13 };

__device__ __forceinline__ void mb_bar_epilogue() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <bool OUT_COLSUM>
__global__ void __launch_bounds__(kMbThreads, 1)
pair_mlp_bwd_layer_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_a,
                          const __grid_constant__ CUtensorMap map_out, const __nv_bfloat16* __restrict__ W,
                          const uint8_t* __restrict__ valid, int L, int has_out, float* __restrict__ dW,
                          float* __restrict__ db, float* __restrict__ db_prev, int n_tiles, int64_t n_pairs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = MbSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float* s_corr = reinterpret_cast<float*>(smem + S::kCorr);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const int n_local = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < kMbStages; ++s) { mbar_init(&bars[MB_FULL + s], 1); mbar_init(&bars[MB_EMPTY + s], 129); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars[MB_ACC_FULL + a], 1); mbar_init(&bars[MB_ACC_EMPTY + a], 128); }
    mbar_init(&bars[MB_FINAL], 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_g); tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_out);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // W^T into its K-major swizzled tile (row = in, 64 out values = 128 B), the tile of ones, the correction sums
  for (int idx = tid; idx < 64 * 64; idx += kMbThreads) {
    const int out = idx >> 6, in = idx & 63;
    *reinterpret_cast<__nv_bfloat16*>(smem + S::kWt + swz128_offset(in, out >> 3) + (out & 7) * 2) = W[idx];
  }
  for (int idx = tid; idx < S::kTile / 16; idx += kMbThreads)
    reinterpret_cast<uint4*>(smem + S::kOnes)[idx] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid < 64) s_corr[tid] = 0.f;
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t kColW = 128;          // wgrad accumulator: [64 out] x 72 columns (64 in + 8 copies of the column sum)

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();       // both streams are read exactly once
      for (int it = 0; it < n_local; ++it) {
        const int s = it % kMbStages, row0 = ((int)blockIdx.x + it * (int)gridDim.x) * 128;
        if (it >= kMbStages) mbar_wait(&bars[MB_EMPTY + s], (it / kMbStages - 1) & 1);
        mbar_arrive_expect_tx(&bars[MB_FULL + s], S::kStage);
        tma_load_2d_hint(smem + s * S::kStage, &map_g, &bars[MB_FULL + s], 0, row0, pol);
        tma_load_2d_hint(smem + s * S::kStage + S::kTile, &map_a, &bars[MB_FULL + s], 0, row0, pol);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_d = make_idesc_bf16(128, 64, 0, 0);    // dgrad: A = g tile K-major, B = W^T K-major
    constexpr uint32_t idesc_w = make_idesc_bf16(64, 72, 1, 1);     // wgrad: A = g tile MN-major, B = [a tile | ones] MN-major
    for (int it = 0; it < n_local; ++it) {
      const int s = it % kMbStages, acc = it & 1;
      mbar_wait(&bars[MB_FULL + s], (it / kMbStages) & 1);
      if (has_out && it >= 2) mbar_wait(&bars[MB_ACC_EMPTY + acc], ((it >> 1) - 1) & 1);
      tcgen05_fence_after_sync();
      if (elect_one()) {
        const uint32_t ga = smem_base + s * S::kStage, aa = ga + S::kTile;
        if (has_out) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t da = make_smem_desc(ga + k * 32, 16, 1024, kSwizzle128B);
            uint64_t dbw = make_smem_desc(smem_base + S::kWt + k * 32, 16, 1024, kSwizzle128B);
            umma_bf16(tmem + acc * 64, da, dbw, idesc_d, k != 0);
          }
          umma_commit(&bars[MB_ACC_FULL + acc]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // 16 pairs (K) per instruction = 2,048 B of either tile; the second 64-wide N atom of B is the tile of ones
          uint64_t da = make_smem_desc(ga + k * 2048, 16384, 1024, kSwizzle128B);
          uint64_t dba = make_smem_desc(aa + k * 2048, (smem_base + S::kOnes) - aa, 1024, kSwizzle128B);
          umma_bf16(tmem + kColW, da, dba, idesc_w, (it | k) != 0);
        }
        umma_commit(&bars[MB_EMPTY + s]);
        if (it == n_local - 1) umma_commit(&bars[MB_FINAL]);
      }
      __syncwarp();
    }
  } else {
    // ============================ epilogue: thread = pair ============================
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;
    const int et = (warp - 2) * 32 + lane;         // 0..127 among the epilogue threads
    const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
    float osum[OUT_COLSUM ? 64 : 1];
    if (OUT_COLSUM) {
#pragma unroll
      for (int c = 0; c < 64; ++c) osum[c] = 0.f;
    }
    for (int it = 0; it < n_local; ++it) {
      const int s = it % kMbStages, acc = it & 1;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      mbar_wait(&bars[MB_FULL + s], (it / kMbStages) & 1);
      // sign mask of this pair's row of a (bf16 > 0: sign clear and not zero)
      uint32_t m_lo = 0, m_hi = 0;
      const uint8_t* arow = smem + s * S::kStage + S::kTile;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(arow + swz128_offset(row, c));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t bits = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
          bits |= (uint32_t)(lo - 1u < 0x7FFFu) << (2 * e);        // 0x0001..0x7FFF: positive (NaN payloads count as > 0: never produced)
          bits |= (uint32_t)(hi - 1u < 0x7FFFu) << (2 * e + 1);
        }
        if (c < 4) m_lo |= bits << (8 * c); else m_hi |= bits << (8 * (c - 4));
      }
      if (valid != nullptr) {
        const int64_t p = (int64_t)tile * 128 + row;
        const int64_t bi = p / L;
        const int j = (int)(p - bi * L);
        const int64_t b = bi / L;
        // rare: take this pair's row of g out of the bias gradient again (rows past the end of a ragged last tile are zero)
        if (p < n_pairs && !(valid[bi] && valid[b * L + j])) {
          const uint8_t* grow = smem + s * S::kStage;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(grow + swz128_offset(row, c));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              atomicAdd(&s_corr[c * 8 + 2 * e], __uint_as_float(w[e] << 16));
              atomicAdd(&s_corr[c * 8 + 2 * e + 1], __uint_as_float(w[e] & 0xFFFF0000u));
            }
          }
        }
      }
      mbar_arrive(&bars[MB_EMPTY + s]);
      if (!has_out) continue;                      // (first layer of a chain: no gradient to pass on)
      mbar_wait(&bars[MB_ACC_FULL + acc], (it >> 1) & 1);
      tcgen05_fence_after_sync();
      float v[64];
      {
        float t0[32], t1[32];
        tmem_ld_x32(tmem_lane + acc * 64, t0);
        tmem_ld_x32(tmem_lane + acc * 64 + 32, t1);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) { v[c] = t0[c]; v[32 + c] = t1[c]; }
      }
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[MB_ACC_EMPTY + acc]);
      // staging tile it & 1: the TMA store that read it two tiles ago must have finished reading
      if (et == 0) tma_store_wait_read_1();
      mb_bar_epilogue();
      uint8_t* stage = smem + S::kStaging + (it & 1) * S::kTile;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = c * 8 + 2 * e;
          const uint32_t mm = c0 < 32 ? m_lo >> c0 : m_hi >> (c0 - 32);
          const float x0 = (mm & 1u) ? v[c0] : 0.f, x1 = (mm & 2u) ? v[c0 + 1] : 0.f;
          const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&h);
          if (OUT_COLSUM) { osum[c0] += __low2float(h); osum[c0 + 1] += __high2float(h); }
        }
        *reinterpret_cast<uint4*>(stage + swz128_offset(row, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      mb_bar_epilogue();
      if (et == 0) {
        tma_store_2d(&map_out, stage, 0, tile * 128);
        tma_store_commit();
      }
    }
    if (et == 0) tma_store_wait_all();
    // ---- the kernel-long accumulators: dW / db from TMEM (M = 64: row out = 16 q + lane on lanes 0-15 of each quadrant)
    mbar_wait(&bars[MB_FINAL], 0);
    tcgen05_fence_after_sync();
    mb_bar_epilogue();                             // s_corr complete, staging tiles free
    {
      float w0[32], w1[32], w2[8];
      tmem_ld_x32(tmem_lane + kColW, w0);
      tmem_ld_x32(tmem_lane + kColW + 32, w1);
      tmem_ld_x8(tmem_lane + kColW + 64, w2);
      tmem_wait_ld();
      if (lane < 16) {
        const int out = q * 16 + lane;
        float* drow = dW + out * 64;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c), "f"(w0[c]), "f"(w0[c + 1]),
                       "f"(w0[c + 2]), "f"(w0[c + 3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + 32 + c), "f"(w1[c]), "f"(w1[c + 1]),
                       "f"(w1[c + 2]), "f"(w1[c + 3]) : "memory");
        }
        atomicAdd(db + out, w2[0] - s_corr[out]);
      }
    }
    if (OUT_COLSUM) {
      // column sums of the rows this CTA wrote: 128 threads x 64 partial sums through the (free) staging tiles
      float* red = reinterpret_cast<float*>(smem + S::kStaging);
#pragma unroll
      for (int c = 0; c < 64; ++c) red[c * 128 + et] = osum[c];
      mb_bar_epilogue();
      if (et < 64) {
        float sum = 0.f;
        for (int r = 0; r < 128; ++r) sum += red[et * 128 + ((r + et) & 127)];
        atomicAdd(db_prev + et, sum);
      }
    }
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 256);
}

}  // namespace sm100
}  // namespace dab

using namespace dab;
using namespace dab::sm100;

extern "C" {

/* One 64-channel layer y = a W^T + b of PairEmbedding's MLPs backward in one pass (see the header of this file):
 * dW[64][64] += g^T a, db[64] += column sums of g over the valid pairs, g_prev = (g W) * (a > 0), db_prev[64] += column
 * sums of g_prev.  g_bf16, a_bf16, g_prev_bf16: [P, 64] bf16, P = B*L*L; W_bf16: [64 out][64 in];
 * res_mask (optional): [B, L], pairs with a masked residue are left out of db (their rows of a must be zero);
 * g_prev_bf16 / db_prev may be null.  dW, db, db_prev are ACCUMULATED into (fp32; the caller zeroes them). */
int dab_pair_mlp_bwd_layer_sm100(const void* g_bf16, const void* a_bf16, const void* W_bf16, const uint8_t* res_mask, int B,
                                 int L, void* g_prev_bf16, float* dW, float* db, float* db_prev, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0, DAB_EINVAL, "dab_pair_mlp_bwd_layer_sm100: negative size");
  const int64_t P = (int64_t)B * L * L;
  if (P == 0) return DAB_OK;
  DAB_REQUIRE(P < ((int64_t)1 << 31) - 128, DAB_EUNSUPPORTED, "dab_pair_mlp_bwd_layer_sm100: B*L*L must be below 2^31");
  DAB_REQUIRE(g_bf16 && a_bf16 && W_bf16 && dW && db, DAB_EINVAL, "dab_pair_mlp_bwd_layer_sm100: null pointer");
  DAB_REQUIRE(db_prev == nullptr || g_prev_bf16 != nullptr, DAB_EINVAL,
              "dab_pair_mlp_bwd_layer_sm100: db_prev needs g_prev");
  DAB_REQUIRE(aligned16(g_bf16) && aligned16(a_bf16) && aligned16(g_prev_bf16) && aligned16(dW), DAB_EINVAL,
              "dab_pair_mlp_bwd_layer_sm100: pointers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap mg, ma, mo;
  uint64_t dims[2] = {64, (uint64_t)P}, strides[1] = {128};
  uint32_t box[2] = {64, 128};
  if (int rc = make_tensor_map_bf16(&mg, g_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_tensor_map_bf16(&ma, a_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  mo = mg;
  if (g_prev_bf16)
    if (int rc = make_tensor_map_bf16(&mo, g_prev_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  const int n_tiles = (int)((P + 127) / 128);      // a ragged last tile is zero-filled on load and clipped on store
  int n_sm = 148;
  {
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev_id);
  }
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(W_bf16);
  if (db_prev) {
    DAB_ENSURE_SMEM(pair_mlp_bwd_layer_kernel<true>, MbSmem::kTotal);
    pair_mlp_bwd_layer_kernel<true><<<grid, kMbThreads, MbSmem::kTotal, s>>>(mg, ma, mo, W, res_mask, L, 1, dW, db, db_prev,
                                                                            n_tiles, P);
  } else {
    DAB_ENSURE_SMEM(pair_mlp_bwd_layer_kernel<false>, MbSmem::kTotal);
    pair_mlp_bwd_layer_kernel<false><<<grid, kMbThreads, MbSmem::kTotal, s>>>(mg, ma, mo, W, res_mask, L,
                                                                             g_prev_bf16 ? 1 : 0, dW, db, nullptr, n_tiles, P);
  }
  count_launch();
  return check_launch("dab_pair_mlp_bwd_layer_sm100");
}

}  // extern "C"
