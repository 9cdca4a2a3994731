// Gradients of PairEmbedding's two embedding tables (diffab_pytorch.py:262-285 of the reference) as class sums of the per-pair
// gradient g1 [B, L, L, 64] (bf16) on the tensor cores (sm_100a, L = 128):
//
//   S_type[s_i*21 + s_j, :] += g1[b, i, j, :]                                  (441 classes)
//   S_rel[clamp(r_i - r_j, -D, D) + D, :] += chain_i chain_j g1[b, i, j, :]    (2 D + 1 classes)
//
// A class sum is a GEMM against a one-hot matrix: for a query row (b, i) - one 16 KB tile of g1, read once -
//   U[s', :]   = sum_j [s_j = s'] g1[i, j, :]          M = 64 (21 used), N = 64, K = 128 keys;  A = one-hot of s_j, constant per patch
//   S_rel     += sum_j [rel(i, j) = r] c_i c_j g1[...]  M = 128 (2 D + 1 used), N = 64, K = 128;  A = weighted one-hot of the row:
//                                                        every key thread moves ONE entry per row
// Both A operands are read MN-major (class index contiguous) from 128B-swizzled shared-memory tiles that the key threads
// edit in place; B is the g1 tile as TMA wrote it.  U goes to S_type[s_i*21 + s'] with vector red.global (21 x 16 per row),
// S_rel accumulates in TMEM for the whole CTA and is added once.  The kernel is bound by the 128 B per pair it reads
// (the previous CUDA-core version walked sorted key runs per class: 162 us for the 134 MB of a 64-patch batch).
//   warp 0     TMA producer (ring of 3 tiles)
//   warp 1     tcgen05.mma issuer (warp-convergent, one elected lane)
//   warps 2-5  thread = key j: one-hot upkeep; thread = class: accumulator read-out
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int TG_L = 128, TG_C = 64, TG_V = 21, TG_STAGES = 3, kTgThreads = 192;

struct TgSmem {
  static constexpr int kTile = 16384;
  static constexpr int kAType = TG_STAGES * kTile;           // [128 keys][64 classes] bf16
  static constexpr int kARel = kAType + kTile;               // 2 x [128 keys][128 classes] bf16 (two 64-wide atoms each)
  static constexpr int kBars = kARel + 4 * kTile;
  static constexpr int kTmemSlot = kBars + 160;
  static constexpr int kTotal = kTmemSlot + 16 + 1024 /* alignment slack */;
};
enum TgBar { TG_FULL = 0, TG_EMPTY = 3, TG_READY = 6, TG_REL_FREE = 8, TG_TYPE_DONE = 10, TG_TYPE_FREE = 12, TG_FINAL = 14,
             TG_N_BARS = 15 };

__global__ void __launch_bounds__(kTgThreads, 1)
pair_table_grad_mma_kernel(const __grid_constant__ CUtensorMap map_g, const int64_t* __restrict__ seq,
                           const int64_t* __restrict__ residue_idx, const int64_t* __restrict__ chain_idx, int n_rows,
                           int max_dist, float* __restrict__ s_type, float* __restrict__ s_rel) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = TgSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const int per = (n_rows + (int)gridDim.x - 1) / (int)gridDim.x;      // contiguous query rows: a CTA sees at most a few patches
  const int row_lo = (int)blockIdx.x * per, row_hi = min(n_rows, row_lo + per);
  const int n_local = max(row_hi - row_lo, 0);

  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&bars[TG_FULL + s], 1); mbar_init(&bars[TG_EMPTY + s], 1); }
    for (int k = 0; k < 2; ++k) {
      mbar_init(&bars[TG_READY + k], 128); mbar_init(&bars[TG_REL_FREE + k], 1);
      mbar_init(&bars[TG_TYPE_DONE + k], 1); mbar_init(&bars[TG_TYPE_FREE + k], 128);
    }
    mbar_init(&bars[TG_FINAL], 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_g);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int idx = tid; idx < 5 * S::kTile / 16; idx += kTgThreads)      // the one-hot tiles start as zeros
    reinterpret_cast<uint4*>(smem + S::kAType)[idx] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t kColRel = 128;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int it = 0; it < n_local; ++it) {
        const int s = it % TG_STAGES;
        if (it >= TG_STAGES) mbar_wait(&bars[TG_EMPTY + s], (it / TG_STAGES - 1) & 1);
        mbar_arrive_expect_tx(&bars[TG_FULL + s], S::kTile);
        tma_load_2d_hint(smem + s * S::kTile, &map_g, &bars[TG_FULL + s], 0, (row_lo + it) * TG_L, pol);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(64, 64, 1, 1);      // A = one-hot of s_j (MN-major), B = g1 tile (MN-major)
    constexpr uint32_t idesc_r = make_idesc_bf16(128, 64, 1, 1);     // A = weighted one-hot of the relative position
    for (int it = 0; it < n_local; ++it) {
      const int s = it % TG_STAGES, k2 = it & 1;
      mbar_wait(&bars[TG_FULL + s], (it / TG_STAGES) & 1);
      mbar_wait(&bars[TG_READY + k2], (it >> 1) & 1);
      if (it >= 2) mbar_wait(&bars[TG_TYPE_FREE + k2], ((it >> 1) - 1) & 1);
      tcgen05_fence_after_sync();
      if (elect_one()) {
        const uint32_t ga = smem_base + s * S::kTile;
        const uint32_t ra = smem_base + S::kARel + k2 * 2 * S::kTile;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                  // 16 keys (K) per instruction = 2,048 B of every tile
          uint64_t da = make_smem_desc(smem_base + S::kAType + k * 2048, 16384, 1024, kSwizzle128B);
          uint64_t db = make_smem_desc(ga + k * 2048, 1024, 1024, kSwizzle128B);
          umma_bf16(tmem + k2 * 64, da, db, idesc_t, k != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t da = make_smem_desc(ra + k * 2048, 16384, 1024, kSwizzle128B);   // two 64-class atoms 16 KB apart
          uint64_t db = make_smem_desc(ga + k * 2048, 1024, 1024, kSwizzle128B);
          umma_bf16(tmem + kColRel, da, db, idesc_r, (it | k) != 0);
        }
        umma_commit(&bars[TG_TYPE_DONE + k2]);      // (after the second chain too: both one-hot tiles have been read)
        umma_commit(&bars[TG_EMPTY + s]);
        umma_commit(&bars[TG_REL_FREE + k2]);
        if (it == n_local - 1) umma_commit(&bars[TG_FINAL]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int j = (warp - 2) * 32 + lane;              // the key this thread keeps the one-hot entries of
    const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16);
    auto entry = [&](uint8_t* tile, int cls) {        // element (key j, class cls) of a [128][64-class atoms] MN-major tile
      return reinterpret_cast<__nv_bfloat16*>(tile + (cls >> 6) * S::kTile + swz128_offset(j, (cls & 63) >> 3) + (cls & 7) * 2);
    };
    auto read_out_type = [&](int it) {                 // U of row `it` -> S_type[s_i*21 + class]
      const int k2 = it & 1;
      mbar_wait(&bars[TG_TYPE_DONE + k2], (it >> 1) & 1);
      tcgen05_fence_after_sync();
      if (q < 2) {                                     // M = 64: class 16 q + lane sits on lanes 0-15 of quadrant q
        float t0[32], t1[32];
        tmem_ld_x32(tmem_lane + k2 * 64, t0);
        tmem_ld_x32(tmem_lane + k2 * 64 + 32, t1);
        tmem_wait_ld();
        const int cls = q * 16 + lane;
        if (lane < 16 && cls < TG_V) {
          const int si = (int)__ldg(seq + row_lo + it);
          float* dst = s_type + ((size_t)si * TG_V + cls) * TG_C;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(t0[c]), "f"(t0[c + 1]),
                         "f"(t0[c + 2]), "f"(t0[c + 3]) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 + c), "f"(t1[c]), "f"(t1[c + 1]),
                         "f"(t1[c + 2]), "f"(t1[c + 3]) : "memory");
          }
        }
      }
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[TG_TYPE_FREE + k2]);
    };
    int cur_b = -1, sj = -1, ridx_j = 0, prev_rel[2] = {-1, -1};
    float chain_j = 0.f;
    for (int it = 0; it < n_local; ++it) {
      const int row = row_lo + it, b = row / TG_L, k2 = it & 1;
      const bool new_patch = b != cur_b;
      int old_sj = sj;
      if (new_patch) {
        cur_b = b;
        const int64_t rj = (int64_t)b * TG_L + j;
        sj = (int)__ldg(seq + rj);
        ridx_j = (int)__ldg(residue_idx + rj);
        chain_j = (float)__ldg(chain_idx + rj);
      }
      // ---- weighted one-hot of the relative position: this key's single entry of the row moves
      if (it >= 2) mbar_wait(&bars[TG_REL_FREE + k2], ((it >> 1) - 1) & 1);
      {
        uint8_t* tile = smem + S::kARel + k2 * 2 * S::kTile;
        int rel = (int)__ldg(residue_idx + row) - ridx_j;
        rel = max(-max_dist, min(max_dist, rel)) + max_dist;
        if (prev_rel[k2] >= 0) *entry(tile, prev_rel[k2]) = __float2bfloat16_rn(0.f);
        *entry(tile, rel) = __float2bfloat16_rn((float)__ldg(chain_idx + row) * chain_j);
        prev_rel[k2] = rel;
      }
      if (new_patch) {
        // the one-hot of the residue types changes with the patch: the chains of the previous row must have read it
        if (it > 0) read_out_type(it - 1);
        if (old_sj >= 0) *entry(smem + S::kAType, old_sj) = __float2bfloat16_rn(0.f);
        *entry(smem + S::kAType, sj) = __float2bfloat16_rn(1.f);
      }
      fence_proxy_async_smem();
      mbar_arrive(&bars[TG_READY + k2]);
      if (!new_patch && it > 0) read_out_type(it - 1);
    }
    if (n_local > 0) {
      read_out_type(n_local - 1);
      mbar_wait(&bars[TG_FINAL], 0);
      tcgen05_fence_after_sync();
      float t0[32], t1[32];
      tmem_ld_x32(tmem_lane + kColRel, t0);
      tmem_ld_x32(tmem_lane + kColRel + 32, t1);
      tmem_wait_ld();
      const int cls = q * 32 + lane;                   // M = 128: class = TMEM lane
      if (cls < 2 * max_dist + 1) {
        float* dst = s_rel + (size_t)cls * TG_C;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(t0[c]), "f"(t0[c + 1]),
                       "f"(t0[c + 2]), "f"(t0[c + 3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 + c), "f"(t1[c]), "f"(t1[c + 1]),
                       "f"(t1[c + 2]), "f"(t1[c + 3]) : "memory");
        }
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

/* s_type[441,64] += class sums of g1_bf16[B,L,L,64] over the pair types s_i*21+s_j; s_rel[2*max_dist+1,64] += class sums over
 * the clamped residue-index offsets, weighted by chain_i*chain_j (header of this file).  Tensor-core form of
 * dab_pair_table_grad for L = 128, max_dist <= 63; the tables are ACCUMULATED into with red.global (fp32; the caller zeroes
 * them; summation order not fixed). */
int dab_pair_table_grad_sm100(const void* g1_bf16, const int64_t* seq_masked, const int64_t* residue_idx,
                              const int64_t* chain_idx, int B, int L, int max_dist, float* s_type, float* s_rel, void* stream) {
  DAB_REQUIRE(B >= 0, DAB_EINVAL, "dab_pair_table_grad_sm100: negative size");
  DAB_REQUIRE(L == TG_L && max_dist >= 1 && max_dist <= 63, DAB_EUNSUPPORTED,
              "dab_pair_table_grad_sm100: L = 128 and 1 <= max_dist <= 63 only (got L = %d, max_dist = %d)", L, max_dist);
  if (B == 0) return DAB_OK;
  DAB_REQUIRE((int64_t)B * L < (1 << 24), DAB_EUNSUPPORTED, "dab_pair_table_grad_sm100: batch too large");
  DAB_REQUIRE(g1_bf16 && seq_masked && residue_idx && chain_idx && s_type && s_rel, DAB_EINVAL,
              "dab_pair_table_grad_sm100: null pointer");
  DAB_REQUIRE(aligned16(g1_bf16) && aligned16(s_type) && aligned16(s_rel), DAB_EINVAL,
              "dab_pair_table_grad_sm100: pointers must be 16-byte aligned");
  const int n_rows = B * L;
  CUtensorMap mg;
  uint64_t dims[2] = {64, (uint64_t)n_rows * L}, strides[1] = {128};
  uint32_t box[2] = {64, 128};
  if (int rc = make_tensor_map_bf16(&mg, g1_bf16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  int n_sm = 148;
  {
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev_id);
  }
  const int grid = n_rows < n_sm ? n_rows : n_sm;
  DAB_ENSURE_SMEM(pair_table_grad_mma_kernel, TgSmem::kTotal);
  pair_table_grad_mma_kernel<<<grid, kTgThreads, TgSmem::kTotal, (cudaStream_t)stream>>>(mg, seq_masked, residue_idx, chain_idx,
                                                                                      n_rows, max_dist, s_type, s_rel);
  count_launch();
  return check_launch("dab_pair_table_grad_sm100");
}

}  // extern "C"
