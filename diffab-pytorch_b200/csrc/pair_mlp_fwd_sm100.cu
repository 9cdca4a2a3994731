// PairEmbedding's MLPs behind the first distance layer, forward for TRAINING, as one kernel (tcgen05, sm_100a).
//
// Reference: diffab_pytorch.py:214-223 (distance_embedding layer 2), :262-285 (pair-type / relative-position / dihedral
// features), :303-311 (mlp + pair mask).  With a1 = relu(Wd1 rbf + bd1) given ([P, 64] bf16, P = B*L*L, L = 128):
//
//   fd  = relu(a1 Wd2^T + bd2)
//   h1  = relu(T_type[s_i*21 + s_j] + c_i c_j T_rel[clamp(r_i - r_j)] + fd W1d^T + xh W1h^T)     T_* = W1 applied to the tables
//   h2  = relu(h1 W2^T + b2) * m_i m_j ;   out = (h2 W3^T + b3) * m_i m_j
//
// and every activation the backward pass needs (fd, h1, h2, the angular features xh) is written exactly once, in bf16:
// per query row (b, i) the kernel reads the 16 KB a1 tile and 1 KB of dihedrals and writes 4 x 16 KB + 8 KB.  The per-pair
// base row of h1 (two table rows) is formed in registers - neither it nor a concat tensor exists in HBM.
//
// A persistent CTA (one per SM) works on TWO query rows at a time (two contexts in anti-phase: the four dependent
// GEMM -> epilogue stages of a row are a serial chain):
//   warps 0 / 1     tcgen05.mma issuer of context 0 / 1 (warp-convergent, one elected lane)
//   warps 2-5 / 6-9 context 0 / 1: thread = key j.  Thread 0 of a context is also its TMA producer (a1 tile + the 21 pair-type
//                   table rows this query row can hit) and issues the TMA stores of the activation tiles, which double as the
//                   next stage's A operand (two tiles alternate: fd, h1, h2, out).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int MF_L = 128, MF_C = 64, MF_V = 21, MF_MAXD = 32, MF_NREL = 2 * MF_MAXD + 1, MF_XW = 32;
constexpr int kMfThreads = 320;

struct MfSmem {
  static constexpr int kTile = 16384;                        // [128 keys][64 channels] bf16, 128B swizzle
  static constexpr int kA1 = 0, kT0 = kTile, kT1 = 2 * kTile, kXh = 3 * kTile;
  static constexpr int kType = 4 * kTile;                    // 2 x 21 table rows of 128 B (double-buffered over rows)
  static constexpr int kTypeBytes = MF_V * MF_C * 2;         // 2,688
  static constexpr int kCtxBytes = 4 * kTile + 6144;
  static constexpr int kW = 2 * kCtxBytes;                   // 5 x [64 out][64 in] bf16 K-major swizzled: Wd2, W1d, W1h, W2, W3
  static constexpr int kRel = kW + 5 * 8192;                 // [65][64] bf16
  static constexpr int kBias = kRel + 8448;                  // [3][64] fp32: bd2, b2, b3
  static constexpr int kBars = kBias + 768;
  static constexpr int kTmemSlot = kBars + 128;
  static constexpr int kTotal = kTmemSlot + 16 + 1024 /* alignment slack */;
};
static_assert(MfSmem::kTotal <= 227 * 1024, "shared memory");
enum MfBar { MF_A_FULL = 0, MF_XH_READY = 1, MF_ACC = 2, MF_ACT = 3, MF_CTX_BARS = 4 };

__device__ __forceinline__ uint32_t mf_pk(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(kMfThreads, 1)
pair_mlp_fwd_train_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_fd,
                          const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2,
                          const __grid_constant__ CUtensorMap map_out, const __nv_bfloat16* __restrict__ w5,
                          const float* __restrict__ bias3, const __nv_bfloat16* __restrict__ t_type,
                          const __nv_bfloat16* __restrict__ t_rel, const int64_t* __restrict__ seq,
                          const int64_t* __restrict__ residue_idx, const int64_t* __restrict__ chain_idx,
                          const uint8_t* __restrict__ res_mask, const float* __restrict__ dihedrals,
                          __nv_bfloat16* __restrict__ xh_out, int n_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = MfSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  const float* s_bias = reinterpret_cast<const float*>(smem + S::kBias);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem);
  const int n_local = (n_rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // query rows of this CTA
  auto row_of = [&](int c, int n) { return (int)blockIdx.x + (2 * n + c) * (int)gridDim.x; };
  auto n_ctx_rows = [&](int c) { return (n_local - c + 1) / 2; };

  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      mbar_init(&bars[c * MF_CTX_BARS + MF_A_FULL], 1);
      mbar_init(&bars[c * MF_CTX_BARS + MF_XH_READY], 128);
      mbar_init(&bars[c * MF_CTX_BARS + MF_ACC], 1);
      mbar_init(&bars[c * MF_CTX_BARS + MF_ACT], 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_fd); tma_prefetch_desc(&map_h1);
    tma_prefetch_desc(&map_h2); tma_prefetch_desc(&map_out);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  // weights [5][64 out][64 in] -> K-major swizzled tiles; relative-position table; biases; the feature tiles start as zeros
  for (int idx = tid; idx < 5 * 64 * 8; idx += kMfThreads) {
    const int m = idx >> 9, r = (idx >> 3) & 63, c = idx & 7;
    *reinterpret_cast<uint4*>(smem + S::kW + m * 8192 + swz128_offset(r, c)) = __ldg(reinterpret_cast<const uint4*>(w5) + idx);
  }
  for (int idx = tid; idx < MF_NREL * 8; idx += kMfThreads)
    reinterpret_cast<uint4*>(smem + S::kRel)[idx] = __ldg(reinterpret_cast<const uint4*>(t_rel) + idx);
  for (int idx = tid; idx < 3 * 64; idx += kMfThreads) reinterpret_cast<float*>(smem + S::kBias)[idx] = bias3[idx];
  for (int c = 0; c < 2; ++c)
    for (int idx = tid; idx < S::kTile / 16; idx += kMfThreads)
      reinterpret_cast<uint4*>(smem + c * S::kCtxBytes + S::kXh)[idx] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 2) {
    // ======================================= MMA issuer of context `warp` =======================================
    const int c = warp;
    uint64_t* cb = bars + c * MF_CTX_BARS;
    const uint32_t cs = smem_base + c * S::kCtxBytes;
    const uint32_t acc = tmem + c * 64;
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    auto chain = [&](uint32_t a_addr, int w_slot, bool first) {     // D (+)= A[128 x 64] W[w_slot]^T
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024, kSwizzle128B);
        uint64_t db = make_smem_desc(smem_base + S::kW + w_slot * 8192 + k * 32, 16, 1024, kSwizzle128B);
        umma_bf16(acc, da, db, idesc, !(first && k == 0));
      }
    };
    const int nr = n_ctx_rows(c);
    for (int n = 0; n < nr; ++n) {
      const uint32_t ph = n & 1;
      mbar_wait(&cb[MF_A_FULL], ph);
      if (n > 0) mbar_wait(&cb[MF_ACT], 1);            // the previous row's last epilogue has drained the accumulator
      tcgen05_fence_after_sync();
      if (elect_one()) { chain(cs + S::kA1, 0, true); umma_commit(&cb[MF_ACC]); }            // fd
      __syncwarp();
      mbar_wait(&cb[MF_ACT], 0);
      mbar_wait(&cb[MF_XH_READY], ph);
      tcgen05_fence_after_sync();
      if (elect_one()) { chain(cs + S::kT0, 1, true); chain(cs + S::kXh, 2, false); umma_commit(&cb[MF_ACC]); }   // h1
      __syncwarp();
      mbar_wait(&cb[MF_ACT], 1);
      tcgen05_fence_after_sync();
      if (elect_one()) { chain(cs + S::kT1, 3, true); umma_commit(&cb[MF_ACC]); }            // h2
      __syncwarp();
      mbar_wait(&cb[MF_ACT], 0);
      tcgen05_fence_after_sync();
      if (elect_one()) { chain(cs + S::kT0, 4, true); umma_commit(&cb[MF_ACC]); }            // out
      __syncwarp();
    }
  } else {
    // ======================================= context: thread = key j =======================================
    const int c = (warp - 2) >> 2;
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int j = q * 32 + lane;
    const int et = ((warp - 2) & 3) * 32 + lane;   // 0..127 inside the context
    uint64_t* cb = bars + c * MF_CTX_BARS;
    uint8_t* cs = smem + c * S::kCtxBytes;
    const uint32_t tmem_lane = tmem + ((uint32_t)(q * 32) << 16) + c * 64;
    auto bar_ctx = [&] { asm volatile("bar.sync %0, 128;" ::"r"(c + 1) : "memory"); };
    const int nr = n_ctx_rows(c);
    auto issue_loads = [&](int n) {                // a1 tile and the 21 pair-type table rows of query row n (thread 0 only)
      const int row = row_of(c, n);
      const int si = (int)__ldg(seq + row);
      mbar_arrive_expect_tx(&cb[MF_A_FULL], S::kTile + S::kTypeBytes);
      tma_load_2d_hint(cs + S::kA1, &map_a1, &cb[MF_A_FULL], 0, row * MF_L, policy_evict_first());
      bulk_load_1d(cs + S::kType + (n & 1) * 3072, t_type + (size_t)si * MF_V * MF_C, S::kTypeBytes, &cb[MF_A_FULL]);
    };
    if (et == 0 && nr > 0) issue_loads(0);
    // accumulator row -> f(v, column) -> bf16 -> swizzled tile `dst` -> TMA store to `map`; `dst` doubles as the next A operand
    auto stage_out = [&](uint8_t* dst, const CUtensorMap* map, int row, auto&& f) {
      float v[64];
      {
        float t0[32], t1[32];
        tmem_ld_x32(tmem_lane, t0);
        tmem_ld_x32(tmem_lane + 32, t1);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 32; ++k) { v[k] = t0[k]; v[32 + k] = t1[k]; }
      }
      if (et == 0) tma_store_wait_read_1();         // the store that read `dst` two stages ago has finished reading
      bar_ctx();
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        float o[8];
        f(o, v + k8 * 8, k8);                        // eight channels k8 * 8 .. k8 * 8 + 7
        *reinterpret_cast<uint4*>(dst + swz128_offset(j, k8)) =
            make_uint4(mf_pk(o[0], o[1]), mf_pk(o[2], o[3]), mf_pk(o[4], o[5]), mf_pk(o[6], o[7]));
      }
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&cb[MF_ACT]);
      bar_ctx();
      if (et == 0) {
        tma_store_2d(map, dst, 0, row * MF_L);
        tma_store_commit();
      }
    };
    int cur_b = -1, sj = 0, ridx_j = 0;
    float chain_j = 0.f, mask_j = 0.f;
    for (int n = 0; n < nr; ++n) {
      const uint32_t ph = n & 1;
      const int row = row_of(c, n);
      const int b = row / MF_L;
      if (b != cur_b) {                            // per-patch data of key j
        cur_b = b;
        const int64_t rj = (int64_t)b * MF_L + j;
        sj = (int)__ldg(seq + rj);
        ridx_j = (int)__ldg(residue_idx + rj);
        chain_j = (float)__ldg(chain_idx + rj);
        mask_j = __ldg(res_mask + rj) ? 1.f : 0.f;
      }
      mbar_wait(&cb[MF_A_FULL], ph);               // (the pair-type rows of this query row come with its a1 tile)
      const int ridx_i = (int)__ldg(residue_idx + row);
      const float cp = (float)__ldg(chain_idx + row) * chain_j;
      const float pm = __ldg(res_mask + row) ? mask_j : 0.f;
      // ---- angular encoding of the two pairwise dihedrals (:20-54): per angle [x, sin(f x) x4, cos(f x) x4], f = 1, 2, 1, 1/2
      {
        const float2 d = __ldg(reinterpret_cast<const float2*>(dihedrals) + (int64_t)row * MF_L + j);
        float f[24];
        const float ang[2] = {d.x, d.y};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const float x = ang[a];
          float s1, c1, s2, c2, sh, ch;
          sincosf(x, &s1, &c1);
          sincosf(2.f * x, &s2, &c2);
          sincosf(0.5f * x, &sh, &ch);
          f[a * 9] = x;
          f[a * 9 + 1] = s1; f[a * 9 + 2] = s2; f[a * 9 + 3] = s1; f[a * 9 + 4] = sh;
          f[a * 9 + 5] = c1; f[a * 9 + 6] = c2; f[a * 9 + 7] = c1; f[a * 9 + 8] = ch;
        }
#pragma unroll
        for (int k = 18; k < 24; ++k) f[k] = 0.f;
        uint4* dst = reinterpret_cast<uint4*>(xh_out + ((int64_t)row * MF_L + j) * MF_XW);
        uint4 u[4];
#pragma unroll
        for (int k8 = 0; k8 < 3; ++k8) {
          u[k8] = make_uint4(mf_pk(f[8 * k8], f[8 * k8 + 1]), mf_pk(f[8 * k8 + 2], f[8 * k8 + 3]),
                             mf_pk(f[8 * k8 + 4], f[8 * k8 + 5]), mf_pk(f[8 * k8 + 6], f[8 * k8 + 7]));
          // (the previous row's h1 chain has finished reading the tile: its accumulator was waited for below)
          *reinterpret_cast<uint4*>(cs + S::kXh + swz128_offset(j, k8)) = u[k8];
        }
        u[3] = make_uint4(0, 0, 0, 0);
        // the pair's 64-byte row as two 256-bit stores (whole sectors; 16-byte stores at a 64-byte lane stride are half sectors)
        st_global_v8(dst, u[0].x, u[0].y, u[0].z, u[0].w, u[1].x, u[1].y, u[1].z, u[1].w);
        st_global_v8(dst + 2, u[2].x, u[2].y, u[2].z, u[2].w, u[3].x, u[3].y, u[3].z, u[3].w);
        fence_proxy_async_smem();
        mbar_arrive(&cb[MF_XH_READY]);
      }
      // ---- fd = relu(a1 Wd2^T + bd2) -> T0
      mbar_wait(&cb[MF_ACC], 0);
      tcgen05_fence_after_sync();
      stage_out(cs + S::kT0, &map_fd, row, [&](float* o, const float* x, int k8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaxf(x[e] + s_bias[k8 * 8 + e], 0.f);
      });
      // the a1 tile has been consumed and (two barriers inside stage_out) every thread of the context is past its wait for
      // this row's loads: the next row's may start (a waiter must never fall two phases behind its barrier)
      if (et == 0 && n + 1 < nr) issue_loads(n + 1);
      // ---- h1 = relu(base + fd W1d^T + xh W1h^T) -> T1; base = pair-type row + c_i c_j relative-position row (registers)
      mbar_wait(&cb[MF_ACC], 1);
      tcgen05_fence_after_sync();
      {
        int rel = ridx_i - ridx_j;
        rel = max(-MF_MAXD, min(MF_MAXD, rel)) + MF_MAXD;
        const uint8_t* ty = cs + S::kType + (n & 1) * 3072 + sj * 128;
        const uint8_t* rl = smem + S::kRel + rel * 128;
        stage_out(cs + S::kT1, &map_h1, row, [&](float* o, const float* x, int k8) {
          const uint4 t4 = *reinterpret_cast<const uint4*>(ty + k8 * 16), r4 = *reinterpret_cast<const uint4*>(rl + k8 * 16);
          const __nv_bfloat162* tp = reinterpret_cast<const __nv_bfloat162*>(&t4);
          const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __bfloat1622float2(tp[e]), r = __bfloat1622float2(rp[e]);
            o[2 * e] = fmaxf(x[2 * e] + fmaf(cp, r.x, t.x), 0.f);
            o[2 * e + 1] = fmaxf(x[2 * e + 1] + fmaf(cp, r.y, t.y), 0.f);
          }
        });
      }
      // ---- h2 = relu(h1 W2^T + b2) * pair mask -> T0
      mbar_wait(&cb[MF_ACC], 0);
      tcgen05_fence_after_sync();
      stage_out(cs + S::kT0, &map_h2, row, [&](float* o, const float* x, int k8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaxf(x[e] + s_bias[64 + k8 * 8 + e], 0.f) * pm;
      });
      // ---- out = (h2 W3^T + b3) * pair mask -> T1
      mbar_wait(&cb[MF_ACC], 1);
      tcgen05_fence_after_sync();
      stage_out(cs + S::kT1, &map_out, row, [&](float* o, const float* x, int k8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (x[e] + s_bias[128 + k8 * 8 + e]) * pm;
      });
    }
    if (et == 0) tma_store_wait_all();
  }
  tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_free(tmem, 128);
}

}  // namespace sm100
}  // namespace dab

using namespace dab;
using namespace dab::sm100;

extern "C" {

/* PairEmbedding's MLPs behind the first distance layer, training forward, in one kernel (header of this file).
 * a1_bf16 [B,L,L,64] = relu(Wd1 rbf + bd1); w5_bf16 [5][64 out][64 in] = Wd2, W1[:, 128:192], W1[:, 192:] zero-padded to 64
 * columns, W2, W3; bias3 [3][64] fp32 = bd2, b2, b3; t_type_bf16 [441,64] = E_type W1[:, :64]^T + b1; t_rel_bf16 [65,64] =
 * E_rel W1[:, 64:128]^T.  Outputs (bf16): fd, h1, h2 (masked pairs zero), out (masked pairs zero) [B,L,L,64] and the angular
 * features xh [B,L,L,32] (columns 18..31 zero).  L = 128 and max_dist = 32 only. */
int dab_pair_mlp_fwd_train_sm100(const void* a1_bf16, const float* pairwise_dihedrals, const int64_t* seq_masked,
                                 const int64_t* residue_idx, const int64_t* chain_idx, const uint8_t* res_mask,
                                 const void* t_type_bf16, const void* t_rel_bf16, const void* w5_bf16, const float* bias3, int B,
                                 int L, int max_dist, void* fd_bf16, void* h1_bf16, void* h2_bf16, void* out_bf16,
                                 void* xh_bf16, void* stream) {
  DAB_REQUIRE(B >= 0, DAB_EINVAL, "dab_pair_mlp_fwd_train_sm100: negative size");
  DAB_REQUIRE(L == MF_L && max_dist == MF_MAXD, DAB_EUNSUPPORTED,
              "dab_pair_mlp_fwd_train_sm100: L = 128 and max_dist = 32 only (got L = %d, max_dist = %d)", L, max_dist);
  if (B == 0) return DAB_OK;
  DAB_REQUIRE((int64_t)B * L < (1 << 24), DAB_EUNSUPPORTED, "dab_pair_mlp_fwd_train_sm100: batch too large");
  DAB_REQUIRE(a1_bf16 && pairwise_dihedrals && seq_masked && residue_idx && chain_idx && res_mask && t_type_bf16 && t_rel_bf16 &&
                  w5_bf16 && bias3 && fd_bf16 && h1_bf16 && h2_bf16 && out_bf16 && xh_bf16,
              DAB_EINVAL, "dab_pair_mlp_fwd_train_sm100: null pointer");
  DAB_REQUIRE(aligned16(a1_bf16) && aligned16(t_type_bf16) && aligned16(t_rel_bf16) && aligned16(w5_bf16) && aligned16(fd_bf16) &&
                  aligned16(h1_bf16) && aligned16(h2_bf16) && aligned16(out_bf16) && aligned32(xh_bf16) &&
                  (reinterpret_cast<uintptr_t>(pairwise_dihedrals) & 7) == 0,
              DAB_EINVAL, "dab_pair_mlp_fwd_train_sm100: pointers must be 16-byte aligned (xh: 32)");
  const int n_rows = B * L;
  const uint64_t P = (uint64_t)n_rows * L;
  CUtensorMap maps[5];
  const void* bases[5] = {a1_bf16, fd_bf16, h1_bf16, h2_bf16, out_bf16};
  uint64_t dims[2] = {64, P}, strides[1] = {128};
  uint32_t box[2] = {64, 128};
  for (int k = 0; k < 5; ++k)
    if (int rc = make_tensor_map_bf16(&maps[k], bases[k], 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  int n_sm = 148;
  {
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev_id);
  }
  const int grid = n_rows < n_sm ? n_rows : n_sm;
  DAB_ENSURE_SMEM(pair_mlp_fwd_train_kernel, MfSmem::kTotal);
  pair_mlp_fwd_train_kernel<<<grid, kMfThreads, MfSmem::kTotal, (cudaStream_t)stream>>>(
      maps[0], maps[1], maps[2], maps[3], maps[4], reinterpret_cast<const __nv_bfloat16*>(w5_bf16), bias3,
      reinterpret_cast<const __nv_bfloat16*>(t_type_bf16), reinterpret_cast<const __nv_bfloat16*>(t_rel_bf16), seq_masked,
      residue_idx, chain_idx, res_mask, pairwise_dihedrals, reinterpret_cast<__nv_bfloat16*>(xh_bf16), n_rows);
  count_launch();
  return check_launch("dab_pair_mlp_fwd_train_sm100");
}

}  // extern "C"
