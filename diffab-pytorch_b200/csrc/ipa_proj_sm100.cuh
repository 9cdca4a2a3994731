// Fused projection stage of the sm_100a IPA path: x (fp32) -> bf16 -> six projections on tcgen05 ->
// frame transform, logit scales, split-bf16 -> packed attention operands Qp / Kp / Vp (+ centred t).
//
// Replaces diffab_pytorch.py:391-413 (to_{q,k,v}_{scalar,point} + euclidean_transform) in one launch:
// the fp32 projection tensor never exists in HBM.  One CTA per patch (128 residues = the M tile of the MMA).
//   warps 0-7: convert the x tile to bf16 in shared memory (128B-swizzled, K-major), then run the epilogue as
//              two groups of 128 threads, one per TMEM accumulator (group g packs tiles t = g, g+2, ...):
//              thread r of a group owns residue r = TMEM lane r, so the frame (R_r, t_r) lives in its registers
//   warp 8   : streams the 24 weight tiles by TMA (ring of 3) and issues the tcgen05.mma chains
//              (M=128, N=64 for two heads of scalars, N=48 for two heads of points, K=128) into the two
//              alternating accumulators, so two tiles are packed while the next ones are computed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

struct ProjSmem {
  static constexpr int kA = 0;                        // [kb(2)][128 rows][128 B] bf16
  static constexpr int kABytes = 2 * 128 * 128;       // 32,768
  static constexpr int kWStage = 2 * 64 * 128;        // 16,384: [kb(2)][64 rows][128 B]
  static constexpr int kWStages = 3;
  static constexpr int kW = kA + kABytes;
  static constexpr int kStage = kW + kWStages * kWStage;   // 81,920: per-warp staging, two [32 rows][64 B] tiles, 64B-swizzled
  static constexpr int kStageGroup = 2 * 8192;        // (512-byte aligned), the source of the TMA stores; 4 warps per group
  static constexpr int kMisc = kStage + 2 * kStageGroup;   // 114,688
  static constexpr int kBars = kMisc;                 // 24 mbarriers
  static constexpr int kTmemSlot = kBars + 24 * 8;
  static constexpr int kCen = kTmemSlot + 16;         // centroid partials [4][3] + result [3]
  static constexpr int kTotal = kMisc + 512;          // 115,200: two CTAs per SM, exactly
  // fused to_out phase (the previous layer's y = cat Wout^T + b, computed straight into the A tile): ring of three 32 KB
  // stages [128 rows of cat | 128 rows of Wout] x 64 K columns, in regions that are idle until the projections start
  static constexpr int kGStage0 = kA, kGStage1 = kW, kGStage2 = kStage;
  static constexpr int kGStageBytes = 32768;
};
static_assert(ProjSmem::kW + ProjSmem::kGStageBytes <= ProjSmem::kStage && ProjSmem::kStage % 1024 == 0, "to_out ring layout");
enum ProjBar { W_FULL = 0, W_EMPTY = 3, ACC_FULL = 6, ACC_EMPTY = 8, G_FULL = 10, G_EMPTY = 13, G_DONE = 16, A_READY = 17,
               PROJ_N_BARS = 18 };

constexpr int kProjTiles = 24;   // 12 scalar tiles (q,k,v x 4 head pairs) + 12 point tiles

__device__ __forceinline__ uint32_t pk_bf(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ uint32_t pk_h(float a, float b) {
  __half2 p = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// map_w64 / map_w48: the packed bf16 weight matrix [1344, 128] with boxes {64, 64} / {64, 48}
__global__ void __launch_bounds__(288, 2)
ipa_proj_kernel(const __grid_constant__ CUtensorMap map_w64, const __grid_constant__ CUtensorMap map_w48,
                const __grid_constant__ CUtensorMap map_sq, const __grid_constant__ CUtensorMap map_sk,
                const __grid_constant__ CUtensorMap map_sv, const float* __restrict__ x, const float* __restrict__ R, const float* __restrict__ t,
                const float* __restrict__ gamma, __nv_bfloat16* __restrict__ Qp, __nv_bfloat16* __restrict__ Kp,
                __nv_bfloat16* __restrict__ Vp, float* __restrict__ tc, long long* __restrict__ dbg,
                const __nv_bfloat16* __restrict__ x16, const __grid_constant__ CUtensorMap map_cat,
                const __grid_constant__ CUtensorMap map_wout, const float* __restrict__ b_out, int fuse_out,
                const float* __restrict__ cen_ext, const float4* __restrict__ front_c, const float4* __restrict__ front_t1,
                const int64_t* __restrict__ front_seq) {
  // fuse_out: 0 = x / x16 is the layer input; 1 = the input is the PREVIOUS layer's to_out, y = cat Wout^T + b, computed
  // here (map_cat, map_wout, b_out); 2 = the input is the epsilon network's front MLP (Denoiser.to_res_emb during sampling,
  // diffab_pytorch.py:572-574): y = relu(c[row] + t1[seq[row]]) W2^T + b2 with the first layer regrouped into the per-run
  // constant c and the 25-row table t1 (map_wout = W2 [128][128] bf16, b_out = b2) - neither ever exists in HBM.
  long long* dbg_cta = dbg ? dbg + (size_t)blockIdx.y * 64 : nullptr;   // (split 0 of) one patch per record
#define PROJ_STAMP(k) do { if (dbg_cta && threadIdx.x == 0) dbg_cta[(k)] = clock64(); } while (0)
  PROJ_STAMP(0);
  constexpr int L = 128, D = 128, H = 8, DS = 32, P = 8, QK_W = 96, V_W = 64;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(1024) uint8_t smem[];
  using S = ProjSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  float* s_cen = reinterpret_cast<float*>(smem + S::kCen);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // grid (n_split, B): CTA (s, b) computes tiles s, s + n_split, ... of patch b (small batches: more CTAs per patch)
  const int b = blockIdx.y, split = blockIdx.x, n_split = gridDim.x, n_local = kProjTiles / n_split;
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) asm volatile("trap;");

  if (tid == 0) {
    for (int i = 0; i < PROJ_N_BARS; ++i)
      mbar_init(&bars[i], (i == ACC_EMPTY || i == ACC_EMPTY + 1) ? 128u : (i == A_READY ? 256u : 1u));
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 128);

  if (warp < 8) {
    // ---- x tile: fp32 global (coalesced float4) -> bf16, K-major 128B-swizzled A operand (16 rows per warp)
    const uint32_t kb = lane >> 4, chunk = (lane & 15) >> 1, half = (lane & 1) * 8;
    if (fuse_out == 2) {
      // first layer of the front MLP -> the A parts of the first two ring stages (K columns 0..63 | 64..127), bf16
      const int g_off2[2] = {S::kGStage0, S::kGStage1};
#pragma unroll 4
      for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const int64_t row = (int64_t)b * L + r;
        const float4 cv = __ldg(front_c + row * (D / 4) + lane);
        const float4 tv = __ldg(front_t1 + __ldg(front_seq + row) * (D / 4) + lane);
        const uint2 o = make_uint2(pk_bf(fmaxf(cv.x + tv.x, 0.f), fmaxf(cv.y + tv.y, 0.f)),
                                   pk_bf(fmaxf(cv.z + tv.z, 0.f), fmaxf(cv.w + tv.w, 0.f)));
        *reinterpret_cast<uint2*>(smem + g_off2[kb] + swz128_offset(r, chunk) + half) = o;
      }
    } else if (fuse_out) {
      // the A tile is produced in place by the fused to_out phase below
    } else if (x16 != nullptr) {   // the previous layer's to_out GEMM already rounded its output to bf16: straight copy
      const __nv_bfloat16* xb = x16 + (int64_t)b * L * D;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const uint2 o = __ldg(reinterpret_cast<const uint2*>(xb + r * D) + lane);
        *reinterpret_cast<uint2*>(smem + S::kA + kb * 16384 + swz128_offset(r, chunk) + half) = o;
      }
    } else {
      const float* xb = x + (int64_t)b * L * D;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {   // all 16 loads of the warp in flight before the first conversion
        const int r = warp * 16 + rr;
        float4 v = __ldg(reinterpret_cast<const float4*>(xb + r * D) + lane);
        uint2 o = make_uint2(pk_bf(v.x, v.y), pk_bf(v.z, v.w));
        *reinterpret_cast<uint2*>(smem + S::kA + kb * 16384 + swz128_offset(r, chunk) + half) = o;
      }
    }
    // ---- patch centroid of the translations (see ipa_sm100.cu: keeps the expanded distance well conditioned).  A patch
    //      of 256 residues is two blocks of this kernel: both must centre on the SAME point, handed in as cen_ext
    if (cen_ext != nullptr) {
      if (tid < 12) s_cen[tid] = tid < 3 ? __ldg(cen_ext + (int64_t)b * 3 + tid) * (float)L : 0.f;
    } else if (warp < 4) {
      const float* tp = t + ((int64_t)b * L + tid) * 3;
      float cx = warp_sum(tp[0]), cy = warp_sum(tp[1]), cz = warp_sum(tp[2]);
      if (lane == 0) { s_cen[warp * 3] = cx; s_cen[warp * 3 + 1] = cy; s_cen[warp * 3 + 2] = cz; }
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  PROJ_STAMP(1);

  if (warp == 8) {
    // The whole warp walks the tile loop (warp-uniform values stay in uniform registers); one elected lane issues the TMA
    // loads and the tcgen05.mma chains.
    auto tile_rows = [](int tt) { return tt < 12 ? 64 : 48; };
    auto tile_row0 = [](int tt) { return tt < 12 ? tt * 64 : 768 + (tt - 12) * 48; };
    auto load_w = [&](int tt, int k) {      // called by ONE lane
      const int s = k % S::kWStages;
      uint8_t* dst = smem + S::kW + s * S::kWStage;
      const int rows = tile_rows(tt);
      mbar_arrive_expect_tx(&bars[W_FULL + s], 2 * rows * 128);
      const CUtensorMap* m = tt < 12 ? &map_w64 : &map_w48;
      tma_load_2d(dst, m, &bars[W_FULL + s], 0, tile_row0(tt));            // K 0..63
      tma_load_2d(dst + rows * 128, m, &bars[W_FULL + s], 64, tile_row0(tt));  // K 64..127
    };
    if (fuse_out) {
      // ---- fused to_out of the previous layer: acc[128 residues x 128] = cat[128 x 1024] Wout^T, 16 K chunks of 64
      constexpr uint32_t idesc_g = make_idesc_bf16(128, 128, 0, 0);
      const int g_off[3] = {S::kGStage0, S::kGStage1, S::kGStage2};
      const bool front = fuse_out == 2;             // A parts already in place (front MLP): only the weight chunk is loaded
      auto load_g = [&](int kc) {     // ONE lane
        const int st = kc % 3;
        uint8_t* dst = smem + g_off[st];
        mbar_arrive_expect_tx(&bars[G_FULL + st], front ? 16384 : S::kGStageBytes);
        if (!front) tma_load_2d(dst, &map_cat, &bars[G_FULL + st], kc * 64, b * L);
        tma_load_2d(dst + 16384, &map_wout, &bars[G_FULL + st], kc * 64, 0);
      };
      const int kChunks = front ? D / 64 : 1024 / 64;
      if (elect_one()) {
        if (!front) tma_prefetch_desc(&map_cat);
        tma_prefetch_desc(&map_wout);
        tma_prefetch_desc(&map_w64); tma_prefetch_desc(&map_w48);
        for (int kc = 0; kc < 3 && kc < kChunks; ++kc) load_g(kc);
      }
      __syncwarp();
      for (int kc = 0; kc < kChunks; ++kc) {
        const int st = kc % 3;
        mbar_wait(&bars[G_FULL + st], (kc / 3) & 1);
        tcgen05_fence_after_sync();
        if (elect_one()) {
          const uint64_t da = make_smem_desc(smem_base + g_off[st], 16, 1024, kSwizzle128B);
          const uint64_t db = make_smem_desc(smem_base + g_off[st] + 16384, 16, 1024, kSwizzle128B);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem, da + (uint32_t)((kk * 32) >> 4), db + (uint32_t)((kk * 32) >> 4), idesc_g, (kc | kk) != 0);
          umma_commit(&bars[G_EMPTY + st]);
          if (kc == kChunks - 1) umma_commit(&bars[G_DONE]);
        }
        __syncwarp();
        if (kc + 3 < kChunks) {
          mbar_wait(&bars[G_EMPTY + st], (kc / 3) & 1);
          if (elect_one()) load_g(kc + 3);
          __syncwarp();
        }
      }
      // the projections' weight ring and the A tile share the to_out ring's memory: every to_out MMA must have completed
      mbar_wait(&bars[G_DONE], 0);
    }
    if (elect_one()) {
      if (!fuse_out) { tma_prefetch_desc(&map_w64); tma_prefetch_desc(&map_w48); }
      for (int k = 0; k < S::kWStages && k < n_local; ++k) load_w(split + k * n_split, k);
    }
    __syncwarp();
    if (fuse_out) {                  // the epilogue warps have written y (bf16) into the A tile and left the accumulator
      mbar_wait(&bars[A_READY], 0);
      tcgen05_fence_after_sync();
    }
    const uint64_t dA0 = make_smem_desc(smem_base + S::kA, 16, 1024, kSwizzle128B);
    const uint64_t dW0 = make_smem_desc(smem_base + S::kW, 16, 1024, kSwizzle128B);
    for (int k = 0; k < n_local; ++k) {       // k-th tile of this CTA = global tile tt
      const int tt = split + k * n_split;
      const int s = k % S::kWStages, acc = k & 1;
      const int rows = tile_rows(tt);
      mbar_wait(&bars[W_FULL + s], (k / S::kWStages) & 1);
      if (k >= 2) mbar_wait(&bars[ACC_EMPTY + acc], ((k >> 1) - 1) & 1);   // epilogue drained this accumulator
      tcgen05_fence_after_sync();
      if (elect_one()) {
        const uint64_t dw = dW0 + (uint32_t)((s * S::kWStage) >> 4);
        const uint32_t idesc = make_idesc_bf16(128, rows, 0, 0);
        const uint32_t kb1 = (uint32_t)((rows * 128) >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16(tmem + acc * 64, dA0 + (uint32_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                    dw + (kk >> 2) * kb1 + (uint32_t)(((kk & 3) * 32) >> 4), idesc, kk != 0);
        umma_commit(&bars[ACC_FULL + acc]);
        umma_commit(&bars[W_EMPTY + s]);
      }
      __syncwarp();
      if (k >= 1 && k - 1 + S::kWStages < n_local) {
        const int sp = (k - 1) % S::kWStages;
        mbar_wait(&bars[W_EMPTY + sp], ((k - 1) / S::kWStages) & 1);
        if (elect_one()) load_w(split + (k - 1 + S::kWStages) * n_split, k - 1 + S::kWStages);
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: thread = residue; group g = warp / 4 owns accumulator g
    const int g = warp >> 2, gt = tid & 127;
    const int64_t row = (int64_t)b * L + gt;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float cenx = (s_cen[0] + s_cen[3] + s_cen[6] + s_cen[9]) * (1.0f / L);
    const float ceny = (s_cen[1] + s_cen[4] + s_cen[7] + s_cen[10]) * (1.0f / L);
    const float cenz = (s_cen[2] + s_cen[5] + s_cen[8] + s_cen[11]) * (1.0f / L);
    float Rm[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) Rm[c] = __ldg(R + row * 9 + c);
    const float tcx = __ldg(t + row * 3) - cenx, tcy = __ldg(t + row * 3 + 1) - ceny, tcz = __ldg(t + row * 3 + 2) - cenz;
    if (g == 0 && split == 0) { tc[row * 3] = tcx; tc[row * 3 + 1] = tcy; tc[row * 3 + 2] = tcz; }
    const float ss = rsqrtf((float)DS), sp = rsqrtf(4.5f * P), st = rsqrtf(3.0f);
    if (fuse_out) {
      // y = acc + b_out of the previous layer, rounded to bf16 exactly as its to_out GEMM would have written it, goes
      // straight into the A tile (K-major, 128B swizzle): group g converts columns 64 g .. 64 g + 63 = K block g
      mbar_wait(&bars[G_DONE], 0);
      tcgen05_fence_after_sync();
      float a[32], c[32];
      tmem_ld_x32(tmem_lane + g * 64, a);
      tmem_ld_x32(tmem_lane + g * 64 + 32, c);
      tmem_wait_ld();
      uint8_t* arow = smem + S::kA + g * 16384;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float* v8 = q < 4 ? a + 8 * q : c + 8 * (q - 4);
        const float* b8 = b_out + g * 64 + 8 * q;
        *reinterpret_cast<uint4*>(arow + swz128_offset(gt, q)) =
            make_uint4(pk_bf(v8[0] + __ldg(b8), v8[1] + __ldg(b8 + 1)), pk_bf(v8[2] + __ldg(b8 + 2), v8[3] + __ldg(b8 + 3)),
                       pk_bf(v8[4] + __ldg(b8 + 4), v8[5] + __ldg(b8 + 5)), pk_bf(v8[6] + __ldg(b8 + 6), v8[7] + __ldg(b8 + 7)));
      }
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[A_READY]);
    }
    // Every output segment is 64 bytes per residue (a head's scalars, point-hi or point-lo columns).  A thread
    // drops its segment into a warp-private [32 rows][64 B] shared-memory tile (64B swizzle: conflict-free 16-byte
    // stores) and lane 0 hands the tile to the TMA as a 2-D store (row pitch = the packed row).  Two tiles per warp
    // and round; a round = fill, fence, __syncwarp, issue - no block-level barrier anywhere in the epilogue.
    // (Measured at B = 256 with the fused to_out phase: 48.1 us this way, 52.5 us with two 256-bit stores per thread and
    // segment instead, 33.9 us without any output store: the packed operands' write-back costs ~14 us of the kernel.)
    uint8_t* stg = smem + S::kStage + warp * 4096;            // 2 x 2 KB
    const int lrow = gt & 31, grow0 = b * L + (gt & ~31);     // row inside the warp's tile; first row of the tile
    // The two staging tiles of a warp alternate: a tile is refilled as soon as every bulk group but the most recent one
    // (the OTHER tile's store) has finished reading shared memory, so one store is always in flight behind the packing.
    int n_put = 0;
    auto put = [&](const uint4 (&seg)[4], const CUtensorMap* m, int col) {
      const int slot = n_put & 1;
      if (n_put >= 2) {
        if (lane == 0) tma_store_wait_read_1();
        __syncwarp();
      }
      ++n_put;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stg + slot * 2048 + swz64_offset(lrow, q)) = seg[q];
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(m, stg + slot * 2048, col, grow0);
        tma_store_commit();
      }
    };

    for (int k = g; k < n_local; k += 2) {
      const int tt = split + k * n_split;
      const int acc = g;
      mbar_wait(&bars[ACC_FULL + acc], (k >> 1) & 1);
      tcgen05_fence_after_sync();
      if (tt == 4 || tt == 16) PROJ_STAMP(tt == 4 ? 2 : 5);
      float v[64];
      {
        float a[32], c[32];
        tmem_ld_x32(tmem_lane + acc * 64, a);
        tmem_ld_x32(tmem_lane + acc * 64 + 32, c);   // columns 48..63 are stale for point tiles (unused)
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = a[i]; v[32 + i] = c[i]; }
      }
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[ACC_EMPTY + acc]);
      if (tt == 4 || tt == 16) PROJ_STAMP(tt == 4 ? 3 : 6);
      if (tt < 12) {
        // ---------------- two heads of scalars: 32 features each
        const int seg = tt >> 2, h0 = (tt & 3) * 2;
        const float sc = seg == 0 ? st * ss * kLog2e : 1.0f;   // scale_total * scale_scalar * log2(e) folded into q
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int h = h0 + hh;
          uint4 o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float* s8 = v + hh * 32 + q * 8;
            if (seg == 2) {
              o[q] = make_uint4(pk_h(s8[0], s8[1]), pk_h(s8[2], s8[3]), pk_h(s8[4], s8[5]), pk_h(s8[6], s8[7]));
            } else {
              o[q] = make_uint4(pk_bf(s8[0] * sc, s8[1] * sc), pk_bf(s8[2] * sc, s8[3] * sc),
                                pk_bf(s8[4] * sc, s8[5] * sc), pk_bf(s8[6] * sc, s8[7] * sc));
            }
          }
          const CUtensorMap* m = seg == 0 ? &map_sq : (seg == 1 ? &map_sk : &map_sv);
          put(o, m, h * (seg == 2 ? V_W : QK_W));
        }
      } else {
        // ---------------- two heads of points: 8 points x 3 each; euclidean_transform (diffab_pytorch.py:315-324)
        const int seg = (tt - 12) >> 2, h0 = ((tt - 12) & 3) * 2;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int h = h0 + hh;
          float gl[24];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const float px = v[hh * 24 + 3 * p], py = v[hh * 24 + 3 * p + 1], pz = v[hh * 24 + 3 * p + 2];
            gl[3 * p] = px * Rm[0] + py * Rm[3] + pz * Rm[6] + tcx;
            gl[3 * p + 1] = px * Rm[1] + py * Rm[4] + pz * Rm[7] + tcy;
            gl[3 * p + 2] = px * Rm[2] + py * Rm[5] + pz * Rm[8] + tcz;
          }
          if (seg == 2) {   // values: fp16 [32..55]; column 56 = 1 (the O^T MMA then returns sum_j p), zeros after
            uint4 o[4];
#pragma unroll
            for (int q = 0; q < 3; ++q)
              o[q] = make_uint4(pk_h(gl[8 * q], gl[8 * q + 1]), pk_h(gl[8 * q + 2], gl[8 * q + 3]),
                                pk_h(gl[8 * q + 4], gl[8 * q + 5]), pk_h(gl[8 * q + 6], gl[8 * q + 7]));
            o[3] = make_uint4(pk_h(1.0f, 0.0f), 0, 0, 0);
            put(o, &map_sv, h * V_W + 32);
          } else {
            const float ch = st * sp * __ldg(gamma + h) * kLog2e;
            const float sc = seg == 0 ? ch : 1.0f;
            float hi[24], lo[24], n2 = 0.f;
#pragma unroll
            for (int c = 0; c < 24; ++c) {
              const float val = gl[c] * sc;
              hi[c] = __bfloat162float(__float2bfloat16_rn(val));
              lo[c] = val - hi[c];
              n2 = fmaf(gl[c], gl[c], n2);
            }
            uint4 oh[4], ol[4];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              oh[q] = make_uint4(pk_bf(hi[8 * q], hi[8 * q + 1]), pk_bf(hi[8 * q + 2], hi[8 * q + 3]),
                                 pk_bf(hi[8 * q + 4], hi[8 * q + 5]), pk_bf(hi[8 * q + 6], hi[8 * q + 7]));
              ol[q] = make_uint4(pk_bf(lo[8 * q], lo[8 * q + 1]), pk_bf(lo[8 * q + 2], lo[8 * q + 3]),
                                 pk_bf(lo[8 * q + 4], lo[8 * q + 5]), pk_bf(lo[8 * q + 6], lo[8 * q + 7]));
            }
            uint4 tail = make_uint4(0, 0, 0, 0);
            if (seg == 0) {                       // columns that pick up the key-side norm term: (1, 1, 1)
              tail.x = pk_bf(1.0f, 1.0f); tail.y = pk_bf(1.0f, 0.0f);
            } else {                              // -0.5 c_h |k|^2 split three ways (hi + mid + lo)
              const float nk = -0.5f * ch * n2;
              const float a = __bfloat162float(__float2bfloat16_rn(nk));
              const float m = __bfloat162float(__float2bfloat16_rn(nk - a));
              // column 59 = 1 (its Q counterpart is 0): the backward's dQ^T MMA then also returns sum_j dlogit
              tail.x = pk_bf(a, m); tail.y = pk_bf(nk - a - m, 1.0f);
            }
            oh[3] = tail;
            ol[3] = make_uint4(0, 0, 0, 0);
            const CUtensorMap* m = seg == 0 ? &map_sq : &map_sk;
            put(oh, m, h * QK_W + 32);
            put(ol, m, h * QK_W + 64);
          }
        }
      }
      if (tt == 4 || tt == 16) PROJ_STAMP(tt == 4 ? 4 : 7);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  PROJ_STAMP(8);
#undef PROJ_STAMP
  if (warp == 0) tmem_free(tmem, 128);
}

}  // namespace sm100
}  // namespace dab
