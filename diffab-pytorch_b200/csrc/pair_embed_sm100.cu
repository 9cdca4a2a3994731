// Fused forward of PairEmbedding (diffab_pytorch.py:186-312 of the reference) on sm_100a, emitting the pair tensor
// directly in bf16 - the layout the IPA kernels stream (inference / sampling only).
//
//   pair_type = s_i * 21 + s_j
//   f_type = E_type[pair_type]                                              (64)
//   f_rel  = E_rel[clamp(r_i - r_j, -32, 32) + 32] * (chain_i * chain_j)    (64)
//   f_dist = relu(Wd2 relu(Wd1 rbf + bd1) + bd2),  rbf[a,a'] = exp(-softplus(C[pair_type][a,a']) |x_i,a - x_j,a'|^2) m_i,a m_j,a'
//   f_dih  = [x, sin(f x), cos(f x)], f in (1, 2, 1, 1/2), for the two pairwise dihedrals        (18)
//   e_ij   = (W3 relu(W2 relu(W1 [f_type | f_rel | f_dist | f_dih] + b1) + b2) + b3) * m_i,CA m_j,CA
//
// In PyTorch this moves ~4.5 GB of (B, L, L, 225)-sized fp32 intermediates per 32 patches through HBM.  Here a
// persistent CTA walks over query rows (b, i); thread = key j (two threads per j) builds the 128 x 256 bf16 RBF tile in
// shared memory straight from the coordinates, and the five linear layers run as tcgen05.mma chains with the
// activations going TMEM -> bias/relu -> bf16 -> shared memory -> next chain.  HBM traffic per row: the xyz / index
// vectors of the patch (L1/L2 resident), 1 KB of dihedrals and the 16 KB output row.
// Per-row constants come in by 1-D bulk copies: for a fixed i only the 21 table rows s_i*21 .. s_i*21+20 of the
// coefficient and pair-type tables can be hit, and they are contiguous.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

constexpr int PE_L = 128, PE_A = 15, PE_C = 64, PE_V = 21, PE_MAXD = 32;
constexpr int PE_COEF_STRIDE = 244;   // floats per (pair type) row: 15 x 16 (a' padded) + 4: bank-conflict-free stride
constexpr float kPeLog2e = 1.4426950408889634f;

struct PePacked {   // byte offsets inside the packed blob (1024-byte aligned base)
  // bf16 [320][256]: rows 0..63 Wd1 (K = a*16 + a'), 64..127 Wd2, 128..191 W1 (K = [type|rel|dist|dih,0]), 192..255 W2, 256..319 W3
  static constexpr size_t kW = 0;
  static constexpr size_t kBias = (size_t)320 * 256 * 2;                 // fp32 [5][64]
  static constexpr size_t kCoef = kBias + 5 * 64 * 4;                    // fp32 [441][244]: softplus(C) * log2(e)
  static constexpr size_t kType = kCoef + (size_t)441 * PE_COEF_STRIDE * 4;   // bf16 [441][64]
  static constexpr size_t kRel = kType + (size_t)441 * 64 * 2;           // bf16 [65][64]
  static constexpr size_t kTotal = (kRel + 65 * 64 * 2 + 1023) / 1024 * 1024;
};
static_assert(PePacked::kCoef % 16 == 0 && PePacked::kType % 16 == 0 && PePacked::kRel % 16 == 0, "bulk-copy alignment");

struct PeSmem {
  static constexpr int kA = 0;                         // [4 kb][128 rows][128 B] bf16: RBF tile, then the concat features
  static constexpr int kA2 = 65536;                    // [128 rows][128 B] bf16: 64-wide activations
  static constexpr int kW = kA2 + 16384;               // 11 x [64 rows][128 B]: Wd1 (4 K blocks), Wd2, W1 (4), W2, W3
  static constexpr int kCoef = kW + 11 * 8192;         // 21 x 244 fp32 (this row's coefficient rows)
  static constexpr int kCoefBytes = PE_V * PE_COEF_STRIDE * 4;   // 20,496
  static constexpr int kType = kCoef + 20736;          // 21 x 64 bf16
  static constexpr int kRel = kType + 2816;            // 65 x 64 bf16
  static constexpr int kBias = kRel + 8320;            // 5 x 64 fp32
  static constexpr int kRow = kBias + 1280;            // xyz_i (45 f), atom mask bits, s_i, chain_i, ridx_i, resmask_i
  static constexpr int kBars = kRow + 256;
  static constexpr int kTmemSlot = kBars + 16 * 8;
  static constexpr int kTotal = kTmemSlot + 16;
};
static_assert(PeSmem::kTotal <= 227 * 1024, "shared memory");
enum PeBar { PE_W_FULL = 0, PE_ROW_FULL = 1, PE_A_READY = 2 /* 256 */, PE_ACC = 3 /* 5 */, PE_ACT = 8 /* 4 x 256 */, PE_N_BARS = 12 };

__device__ __forceinline__ uint32_t pe_pk(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float pe_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void pack_pair_embed_kernel(DabPairEmbedWeights w, uint8_t* packed) {
  __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(packed + PePacked::kW);
  float* bias = reinterpret_cast<float*>(packed + PePacked::kBias);
  float* coef = reinterpret_cast<float*>(packed + PePacked::kCoef);
  __nv_bfloat16* ty = reinterpret_cast<__nv_bfloat16*>(packed + PePacked::kType);
  __nv_bfloat16* rl = reinterpret_cast<__nv_bfloat16*>(packed + PePacked::kRel);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 320 * 256; i += nth) {
    const int r = i >> 8, k = i & 255, m = r >> 6, n = r & 63;
    float v = 0.f;
    if (m == 0) {                       // Wd1 (64, 225): K index a*16 + a'  <-  a*15 + a'
      const int a = k >> 4, ap = k & 15;
      if (a < PE_A && ap < PE_A) v = w.d_w1[n * 225 + a * PE_A + ap];
    } else if (m == 1) { if (k < 64) v = w.d_w2[n * 64 + k]; }
    else if (m == 2) { if (k < 210) v = w.m_w1[n * 210 + k]; }      // [type 64 | rel 64 | dist 64 | dih 18] as in the cat (:307)
    else if (m == 3) { if (k < 64) v = w.m_w2[n * 64 + k]; }
    else { if (k < 64) v = w.m_w3[n * 64 + k]; }
    wb[i] = __float2bfloat16_rn(v);
  }
  const float* bsrc[5] = {w.d_b1, w.d_b2, w.m_b1, w.m_b2, w.m_b3};
  for (int i = tid; i < 5 * 64; i += nth) bias[i] = bsrc[i >> 6][i & 63];
  for (int i = tid; i < 441 * PE_COEF_STRIDE; i += nth) {
    const int t = i / PE_COEF_STRIDE, k = i % PE_COEF_STRIDE, a = k >> 4, ap = k & 15;
    float v = 0.f;
    if (a < PE_A && ap < PE_A) {
      const float c = w.pair2distcoef[t * 225 + a * PE_A + ap];
      v = (c > 20.f ? c : log1pf(expf(c))) * kPeLog2e;          // F.softplus (threshold 20), log2(e) folded in
    }
    coef[i] = v;
  }
  for (int i = tid; i < 441 * 64; i += nth) ty[i] = __float2bfloat16_rn(w.type_emb[i]);
  for (int i = tid; i < 65 * 64; i += nth) rl[i] = __float2bfloat16_rn(w.relpos_emb[i]);
}

// 288 threads: warps 0-7 = compute (thread = key j = tid & 127, half = tid >> 7), warp 8 lane 0 = TMA + tcgen05.mma issuer.
__global__ void __launch_bounds__(288, 1)
pair_embed_kernel(const __grid_constant__ CUtensorMap map_w, const uint8_t* __restrict__ packed,
                  const int64_t* __restrict__ seq, const float* __restrict__ xyz, const float* __restrict__ dihedrals,
                  const int64_t* __restrict__ residue_idx, const int64_t* __restrict__ chain_idx,
                  const uint8_t* __restrict__ atom_mask, int n_rows, __nv_bfloat16* __restrict__ e_out, int Lp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using S = PeSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);
  const float* s_coef = reinterpret_cast<const float*>(smem + S::kCoef);
  const float* s_bias = reinterpret_cast<const float*>(smem + S::kBias);
  float* s_row = reinterpret_cast<float*>(smem + S::kRow);     // [0..44] xyz_i, [48] mask bits, [49] s_i, [50] chain_i, [51] ridx_i, [52] resmask_i
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) asm volatile("trap;");

  // Work item = (query row, block of 128 keys): `n_rows` items.  Patches of Lp = 256 residues have two key blocks per
  // query row; the items of a patch are ordered key block first, so consecutive items share the per-key registers.
  const int rows_per_cta = (n_rows + gridDim.x - 1) / gridDim.x;
  const int row_lo = blockIdx.x * rows_per_cta, row_hi = min(n_rows, row_lo + rows_per_cta);
  const int nkb = Lp / PE_L, per_patch = Lp * nkb;
  auto qrow_of = [&](int item) { const int b = item / per_patch; return b * Lp + (item - b * per_patch) % Lp; };
  auto kblk_of = [&](int item) { const int b = item / per_patch; return b * nkb + (item - b * per_patch) / Lp; };   // (patch, key block)

  if (tid == 0) {
    for (int i = 0; i < PE_N_BARS; ++i) mbar_init(&bars[i], (i == PE_A_READY || i >= PE_ACT) ? 256u : 1u);
    fence_barrier_init();
  }
  // constant tables: relative-position embedding and the five bias vectors
  for (int i = tid; i < 65 * 64 / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + S::kRel)[i] = __ldg(reinterpret_cast<const uint4*>(packed + PePacked::kRel) + i);
  for (int i = tid; i < 5 * 64; i += blockDim.x)
    reinterpret_cast<float*>(smem + S::kBias)[i] = __ldg(reinterpret_cast<const float*>(packed + PePacked::kBias) + i);
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0 && row_lo < row_hi) {
      tma_prefetch_desc(&map_w);
      // weights: 11 [64 rows][64 K] boxes of the packed [320][256] matrix
      mbar_arrive_expect_tx(&bars[PE_W_FULL], 11 * 8192);
      {
        int slot = 0;
        const int kblocks[5] = {4, 1, 4, 1, 1};
        for (int m = 0; m < 5; ++m)
          for (int kb = 0; kb < kblocks[m]; ++kb, ++slot)
            tma_load_2d(smem + S::kW + slot * 8192, &map_w, &bars[PE_W_FULL], kb * 64, m * 64);
      }
      auto load_row_tables = [&](int item) {    // the 21 coefficient rows and 21 pair-type rows this query row can hit
        const int si = (int)__ldg(seq + qrow_of(item));
        mbar_arrive_expect_tx(&bars[PE_ROW_FULL], S::kCoefBytes + PE_V * 64 * 2);
        bulk_load_1d(smem + S::kCoef, packed + PePacked::kCoef + (size_t)si * PE_V * PE_COEF_STRIDE * 4, S::kCoefBytes,
                     &bars[PE_ROW_FULL]);
        bulk_load_1d(smem + S::kType, packed + PePacked::kType + (size_t)si * PE_V * 64 * 2, PE_V * 64 * 2,
                     &bars[PE_ROW_FULL]);
      };
      load_row_tables(row_lo);
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      auto mma = [&](uint32_t a_addr, int w_slot, int kblocks, uint32_t dcol) {
        for (int k = 0; k < kblocks * 4; ++k) {
          uint64_t da = make_smem_desc(a_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, kSwizzle128B);
          uint64_t db = make_smem_desc(smem_base + S::kW + (w_slot + (k >> 2)) * 8192 + (k & 3) * 32, 16, 1024, kSwizzle128B);
          umma_bf16(tmem + dcol, da, db, idesc, k != 0);
        }
      };
      mbar_wait(&bars[PE_W_FULL], 0);
      for (int row = row_lo, it = 0; row < row_hi; ++row, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(&bars[PE_A_READY], ph);            // RBF tile complete (and the row tables no longer needed)
        tcgen05_fence_after_sync();
        if (row + 1 < row_hi) load_row_tables(row + 1);
        mma(smem_base + S::kA, 0, 4, 0);             // distance_embedding layer 1
        umma_commit(&bars[PE_ACC + 0]);
        mbar_wait(&bars[PE_ACT + 0], ph);
        tcgen05_fence_after_sync();
        mma(smem_base + S::kA2, 4, 1, 64);           // distance_embedding layer 2
        umma_commit(&bars[PE_ACC + 1]);
        mbar_wait(&bars[PE_ACT + 1], ph);
        tcgen05_fence_after_sync();
        mma(smem_base + S::kA, 5, 4, 0);             // mlp layer 1 on [type | rel | dist | dih]
        umma_commit(&bars[PE_ACC + 2]);
        mbar_wait(&bars[PE_ACT + 2], ph);
        tcgen05_fence_after_sync();
        mma(smem_base + S::kA2, 9, 1, 64);           // mlp layer 2
        umma_commit(&bars[PE_ACC + 3]);
        mbar_wait(&bars[PE_ACT + 3], ph);
        tcgen05_fence_after_sync();
        mma(smem_base + S::kA2, 10, 1, 0);           // mlp layer 3
        umma_commit(&bars[PE_ACC + 4]);
      }
    }
  } else {
    const int j = tid & 127, half = tid >> 7;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    auto bar_compute = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // 32 accumulator columns (this thread's half) + bias, relu -> bf16 -> chunks 4 half .. 4 half + 3 of row j of `dst`
    auto act_to_smem = [&](uint32_t col0, const float* bias, uint8_t* dst) {
      float v[32];
      tmem_ld_x32(tmem_lane + col0 + half * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[q * 8 + e] + bias[half * 32 + q * 8 + e], 0.f);
        *reinterpret_cast<uint4*>(dst + swz128_offset(j, half * 4 + q)) =
            make_uint4(pe_pk(o[0], o[1]), pe_pk(o[2], o[3]), pe_pk(o[4], o[5]), pe_pk(o[6], o[7]));
      }
    };
    int cur_b = -1;
    float xj[PE_A * 3];
    uint32_t mj = 0;
    int sj = 0;
    float chain_j = 0.f, resmask_j = 0.f;
    int ridx_j = 0;
    for (int item = row_lo, it = 0; item < row_hi; ++item, ++it) {
      const uint32_t ph = it & 1;
      const int row = qrow_of(item);                  // query row (b, i)
      const int b = kblk_of(item);                    // (patch, key block): the keys are residues 128 kb + j of the patch
      const int64_t joff = (int64_t)(b % nkb) * PE_L + j;      // key index inside the patch
      if (b != cur_b) {       // per-(patch, key block) data of key j (registers)
        cur_b = b;
        const int64_t rj = (int64_t)(b / nkb) * Lp + joff;
#pragma unroll
        for (int c = 0; c < PE_A * 3; ++c) xj[c] = __ldg(xyz + rj * (PE_A * 3) + c);
        mj = 0;
#pragma unroll
        for (int a = 0; a < PE_A; ++a) mj |= (__ldg(atom_mask + rj * PE_A + a) ? 1u : 0u) << a;
        sj = (int)__ldg(seq + rj);
        chain_j = (float)__ldg(chain_idx + rj);
        ridx_j = (int)__ldg(residue_idx + rj);
        resmask_j = (mj >> 1) & 1u ? 1.f : 0.f;      // CA_IDX = 1
      }
      // ---- data of query row i, shared by the CTA
      bar_compute();                                  // previous iteration's readers of s_row are done
      if (tid < PE_A * 3) s_row[tid] = __ldg(xyz + (int64_t)row * (PE_A * 3) + tid);
      if (tid == 64) {
        uint32_t mi = 0;
        for (int a = 0; a < PE_A; ++a) mi |= (__ldg(atom_mask + (int64_t)row * PE_A + a) ? 1u : 0u) << a;
        s_row[48] = __uint_as_float(mi);
        s_row[49] = __int_as_float((int)__ldg(seq + row));
        s_row[50] = (float)__ldg(chain_idx + row);
        s_row[51] = __int_as_float((int)__ldg(residue_idx + row));
        s_row[52] = (mi >> 1) & 1u ? 1.f : 0.f;
      }
      bar_compute();
      const uint32_t mi = __float_as_uint(s_row[48]);
      const float2 dih = __ldg(reinterpret_cast<const float2*>(dihedrals) + (int64_t)row * Lp + joff);
      // ---- RBF tile: thread (j, half) fills atoms a = 8 half .. 8 half + 7 of its row (a = 15 is zero padding)
      mbar_wait(&bars[PE_ROW_FULL], ph);
      if (it > 0) {                                   // the previous row's last chain has finished reading tile A
        mbar_wait(&bars[PE_ACC + 2], (it - 1) & 1);
      }
      const float* crow = s_coef + sj * PE_COEF_STRIDE;
#pragma unroll 1
      for (int aa = 0; aa < 8; ++aa) {
        const int a = half * 8 + aa;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
        if (a < PE_A) {
          const float xi = s_row[a * 3], yi = s_row[a * 3 + 1], zi = s_row[a * 3 + 2];
          const bool ai = (mi >> a) & 1u;
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 c = *reinterpret_cast<const float4*>(crow + a * 16 + q * 4);
            const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ap = q * 4 + e;
              if (ap < PE_A) {
                const float dx = xi - xj[ap * 3], dy = yi - xj[ap * 3 + 1], dz = zi - xj[ap * 3 + 2];
                const float d2 = dx * dx + dy * dy + dz * dz;
                v[ap] = (ai && ((mj >> ap) & 1u)) ? pe_ex2(-cc[e] * d2) : 0.f;
              } else {
                v[ap] = 0.f;
              }
            }
          }
          lo = make_uint4(pe_pk(v[0], v[1]), pe_pk(v[2], v[3]), pe_pk(v[4], v[5]), pe_pk(v[6], v[7]));
          hi = make_uint4(pe_pk(v[8], v[9]), pe_pk(v[10], v[11]), pe_pk(v[12], v[13]), pe_pk(v[14], v[15]));
        }
        uint8_t* blk = smem + S::kA + (a >> 2) * 16384;
        *reinterpret_cast<uint4*>(blk + swz128_offset(j, (a & 3) * 2)) = lo;
        *reinterpret_cast<uint4*>(blk + swz128_offset(j, (a & 3) * 2 + 1)) = hi;
      }
      // pair-type feature of this key, fetched before the row tables are released
      uint4 ftype[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) ftype[q] = *reinterpret_cast<const uint4*>(smem + S::kType + sj * 128 + (half * 4 + q) * 16);
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[PE_A_READY]);
      // ---- distance_embedding layer 1 -> A2; then tile A is free: [type | rel | . | dih] go in while layer 2 runs
      mbar_wait(&bars[PE_ACC + 0], ph);
      tcgen05_fence_after_sync();
      act_to_smem(0, s_bias, smem + S::kA2);
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[PE_ACT + 0]);
      {
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(smem + S::kA + swz128_offset(j, half * 4 + q)) = ftype[q];
        // relative position x product of chain indices (:279-285)
        const int ri = __float_as_int(s_row[51]);
        int rel = ri - ridx_j;
        rel = max(-PE_MAXD, min(PE_MAXD, rel)) + PE_MAXD;
        const float cp = s_row[50] * chain_j;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 r = *reinterpret_cast<const uint4*>(smem + S::kRel + rel * 128 + (half * 4 + q) * 16);
          const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&r);
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(rp[e]);
            o[e] = pe_pk(f.x * cp, f.y * cp);
          }
          *reinterpret_cast<uint4*>(smem + S::kA + 16384 + swz128_offset(j, half * 4 + q)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        // angular encoding of the two pairwise dihedrals (:20-54): per angle [x, sin(f x) x4, cos(f x) x4], f = 1, 2, 1, 1/2
        uint8_t* blk3 = smem + S::kA + 3 * 16384;
        if (half == 0) {
          float f[24];
          const float ang[2] = {dih.x, dih.y};
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const float x = ang[d];
            const float fr[4] = {1.f, 2.f, 1.f, 0.5f};
            f[d * 9] = x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float s, c;
              sincosf(fr[k] * x, &s, &c);
              f[d * 9 + 1 + k] = s;
              f[d * 9 + 5 + k] = c;
            }
          }
#pragma unroll
          for (int k = 18; k < 24; ++k) f[k] = 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            *reinterpret_cast<uint4*>(blk3 + swz128_offset(j, q)) =
                make_uint4(pe_pk(f[8 * q], f[8 * q + 1]), pe_pk(f[8 * q + 2], f[8 * q + 3]), pe_pk(f[8 * q + 4], f[8 * q + 5]),
                           pe_pk(f[8 * q + 6], f[8 * q + 7]));
          *reinterpret_cast<uint4*>(blk3 + swz128_offset(j, 3)) = make_uint4(0, 0, 0, 0);
        } else {
#pragma unroll
          for (int q = 4; q < 8; ++q) *reinterpret_cast<uint4*>(blk3 + swz128_offset(j, q)) = make_uint4(0, 0, 0, 0);
        }
      }
      // ---- distance_embedding layer 2 -> f_dist = K block 2 of tile A
      mbar_wait(&bars[PE_ACC + 1], ph);
      tcgen05_fence_after_sync();
      act_to_smem(64, s_bias + 64, smem + S::kA + 2 * 16384);
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[PE_ACT + 1]);
      // ---- mlp layer 1 -> A2
      mbar_wait(&bars[PE_ACC + 2], ph);
      tcgen05_fence_after_sync();
      act_to_smem(0, s_bias + 128, smem + S::kA2);
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[PE_ACT + 2]);
      // ---- mlp layer 2 -> A2 (its chain has finished reading A2)
      mbar_wait(&bars[PE_ACC + 3], ph);
      tcgen05_fence_after_sync();
      act_to_smem(64, s_bias + 192, smem + S::kA2);
      fence_proxy_async_smem();
      tcgen05_fence_before_sync();
      mbar_arrive(&bars[PE_ACT + 3]);
      // ---- mlp layer 3 (no relu) x residue-pair mask -> bf16 row of the pair tensor
      mbar_wait(&bars[PE_ACC + 4], ph);
      tcgen05_fence_after_sync();
      {
        float v[32];
        tmem_ld_x32(tmem_lane + half * 32, v);
        tmem_wait_ld();
        const float rm = s_row[52] * resmask_j;
        uint4* dst = reinterpret_cast<uint4*>(e_out + ((int64_t)row * Lp + joff) * PE_C + half * 32);
#pragma unroll
        for (int q = 0; q < 2; ++q) {     // 256-bit stores: one whole 32-byte sector per request
          float o[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = (v[q * 16 + e] + s_bias[256 + half * 32 + q * 16 + e]) * rm;
          st_global_v8(dst + 2 * q, pe_pk(o[0], o[1]), pe_pk(o[2], o[3]), pe_pk(o[4], o[5]), pe_pk(o[6], o[7]),
                       pe_pk(o[8], o[9]), pe_pk(o[10], o[11]), pe_pk(o[12], o[13]), pe_pk(o[14], o[15]));
        }
      }
      tcgen05_fence_before_sync();
    }
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

size_t dab_pair_embed_packed_bytes(void) { return PePacked::kTotal; }

int dab_pair_embed_pack_weights(const DabPairEmbedWeights* w, void* packed, void* stream) {
  DAB_REQUIRE(w && packed, DAB_EINVAL, "dab_pair_embed_pack_weights: null pointer");
  const float* const* p = reinterpret_cast<const float* const*>(w);
  for (int i = 0; i < 13; ++i) DAB_REQUIRE(p[i] != nullptr, DAB_EINVAL, "dab_pair_embed_pack_weights: null weight pointer %d", i);
  DAB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 1023) == 0, DAB_EINVAL, "dab_pair_embed_pack_weights: packed buffer must be 1024-byte aligned");
  pack_pair_embed_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(*w, reinterpret_cast<uint8_t*>(packed));
  count_launch();
  return check_launch("dab_pair_embed_pack_weights");
}

int dab_pair_embed_fwd_sm100(const void* packed, const int64_t* seq_masked, const float* xyz, const float* pairwise_dihedrals,
                             const int64_t* residue_idx, const int64_t* chain_idx, const uint8_t* atom_mask, int B, int L_,
                             int A, void* e_bf16, void* stream) {
  DAB_REQUIRE((L_ == PE_L || L_ == 2 * PE_L) && A == PE_A, DAB_EUNSUPPORTED,
              "dab_pair_embed_fwd_sm100: needs L = 128 or 256 residues and 15 atoms per residue");
  DAB_REQUIRE(B >= 0, DAB_EINVAL, "dab_pair_embed_fwd_sm100: negative batch");
  if (B == 0) return DAB_OK;
  DAB_REQUIRE(packed && seq_masked && xyz && pairwise_dihedrals && residue_idx && chain_idx && atom_mask && e_bf16, DAB_EINVAL,
              "dab_pair_embed_fwd_sm100: null pointer");
  DAB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 1023) == 0 && aligned32(e_bf16) &&
                  (reinterpret_cast<uintptr_t>(pairwise_dihedrals) & 7) == 0,
              DAB_EINVAL, "dab_pair_embed_fwd_sm100: misaligned pointer (packed 1024 B, e 32 B, dihedrals 8 B)");
  CUtensorMap mw;
  uint64_t dims[2] = {256, 320}, strides[1] = {512};
  uint32_t box[2] = {64, 64};
  if (int rc = make_tensor_map_bf16(&mw, reinterpret_cast<const uint8_t*>(packed) + PePacked::kW, 2, dims, strides, box,
                                    CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  DAB_ENSURE_SMEM(pair_embed_kernel, PeSmem::kTotal);
  const int n_rows = B * L_ * (L_ / PE_L);      // work items: (query row, block of 128 keys)
  const int grid = n_rows < 148 ? n_rows : 148;
  pair_embed_kernel<<<grid, 288, PeSmem::kTotal, (cudaStream_t)stream>>>(
      mw, reinterpret_cast<const uint8_t*>(packed), seq_masked, xyz, pairwise_dihedrals, residue_idx, chain_idx, atom_mask,
      n_rows, reinterpret_cast<__nv_bfloat16*>(e_bf16), L_);
  count_launch();
  return check_launch("dab_pair_embed_fwd_sm100");
}

}  // extern "C"
