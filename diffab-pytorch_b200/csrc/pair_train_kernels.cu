// Per-pair glue of PairEmbedding for TRAINING (diffab_pytorch.py:262-311 of the reference), mixed-precision path.
//
// The first mlp layer acts on cat[f_type | f_rel | f_dist | f_dih] (:307).  Its three input-only blocks are folded into a
// per-pair "base" row computed straight from the residue-level index vectors:
//   base[b,i,j,:] = T_type[s_i*21 + s_j] + (chain_i*chain_j) * T_rel[clamp(r_i - r_j) + maxd]
// with T_type = E_type W1_type^T + b1 and T_rel = E_rel W1_rel^T precomputed on the 441 / 65 table rows, and emitted
// together with the angular encoding of the pairwise dihedrals (bf16, padded to 32 columns for the GEMM that applies
// W1_dih), so neither the (B, L, L) int64 index tensors nor the two 10^6-row embedding gathers nor the
// angular-encoding passes exist.
// Backward: the per-pair gradient g1 (B, L, L, 64) is summed into the class tables S_type (441 x 64) and S_rel (65 x 64,
// weighted by the chain product) in ONE pass with shared-memory accumulators - PyTorch's embedding backward sorts 10^6
// indices per table.  All three kernels are HBM-bound streams over (B, L, L, 64) bf16 tensors.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {

constexpr int PT_C = 64, PT_V = 21, PT_DIHP = 32;

__device__ __forceinline__ uint32_t pt_pk(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void pt_unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __bfloat1622float2(p[e]);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}

// Lane q of an 8-lane group owns channels 4q..4q+3 and 32+4q..32+4q+3 of a 64-channel row: the group's 8-byte loads from
// a bf16 row and its 16-byte read-modify-writes of an fp32 accumulator row each cover contiguous memory (no bank conflicts).
__device__ __forceinline__ void pt_load_row8(const __nv_bfloat16* row, int q, float* f) {
  const uint2 lo = *reinterpret_cast<const uint2*>(row + 4 * q), hi = *reinterpret_cast<const uint2*>(row + 32 + 4 * q);
  const __nv_bfloat162* pl = reinterpret_cast<const __nv_bfloat162*>(&lo);
  const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&hi);
  const float2 a = __bfloat1622float2(pl[0]), b = __bfloat1622float2(pl[1]), c = __bfloat1622float2(ph[0]), d = __bfloat1622float2(ph[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void pt_add_row8(float* acc_row, int q, const float* f, float scale) {
  float4* lo = reinterpret_cast<float4*>(acc_row + 4 * q);
  float4* hi = reinterpret_cast<float4*>(acc_row + 32 + 4 * q);
  float4 a = *lo, b = *hi;
  a.x = fmaf(scale, f[0], a.x); a.y = fmaf(scale, f[1], a.y); a.z = fmaf(scale, f[2], a.z); a.w = fmaf(scale, f[3], a.w);
  b.x = fmaf(scale, f[4], b.x); b.y = fmaf(scale, f[5], b.y); b.z = fmaf(scale, f[6], b.z); b.w = fmaf(scale, f[7], b.w);
  *lo = a;
  *hi = b;
}

// One block = `rows_per_block` consecutive query rows (b, i) of one patch; 256 threads.
// dynamic smem: seq [L] int | ridx [L] int | chain [L] float
__global__ void __launch_bounds__(256) pair_base_fwd_kernel(
    const int64_t* __restrict__ seq, const int64_t* __restrict__ residue_idx, const int64_t* __restrict__ chain_idx,
    const float* __restrict__ dihedrals, const __nv_bfloat16* __restrict__ t_type, const __nv_bfloat16* __restrict__ t_rel,
    int L, int max_dist, int rows_per_block, __nv_bfloat16* __restrict__ base, __nv_bfloat16* __restrict__ xh) {
  extern __shared__ int s_i[];
  int* s_seq = s_i;
  int* s_ridx = s_seq + L;
  float* s_chain = reinterpret_cast<float*>(s_ridx + L);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t b = row0 / L;
  for (int j = tid; j < L; j += blockDim.x) {
    s_seq[j] = (int)seq[b * L + j];
    s_ridx[j] = (int)residue_idx[b * L + j];
    s_chain[j] = (float)chain_idx[b * L + j];
  }
  __syncthreads();
  for (int r = 0; r < rows_per_block; ++r) {
    const int64_t row = row0 + r;
    const int i = (int)(row - b * L);
    const int si = s_seq[i], ri = s_ridx[i];
    const float ci = s_chain[i];
    // ---- base rows: 8 lanes per pair (16 bytes each), a warp writes 512 contiguous bytes
    for (int idx = tid; idx < L * 8; idx += blockDim.x) {
      const int j = idx >> 3, q = idx & 7;
      const int pt = si * PT_V + s_seq[j];
      int rel = ri - s_ridx[j];
      rel = max(-max_dist, min(max_dist, rel)) + max_dist;
      const float cp = ci * s_chain[j];
      float ty[8], rl[8];
      pt_unpack8(__ldg(reinterpret_cast<const uint4*>(t_type + pt * PT_C) + q), ty);
      pt_unpack8(__ldg(reinterpret_cast<const uint4*>(t_rel + rel * PT_C) + q), rl);
#pragma unroll
      for (int e = 0; e < 8; ++e) ty[e] = fmaf(cp, rl[e], ty[e]);
      *(reinterpret_cast<uint4*>(base + (row * L + j) * PT_C) + q) =
          make_uint4(pt_pk(ty[0], ty[1]), pt_pk(ty[2], ty[3]), pt_pk(ty[4], ty[5]), pt_pk(ty[6], ty[7]));
    }
    // ---- angular encoding of the two pairwise dihedrals (:20-54): per angle [x, sin(f x) x4, cos(f x) x4], f = 1, 2, 1, 1/2;
    //      one thread per pair writes the 64-byte padded feature row (18 features, zeros after)
    for (int j = tid; j < L; j += blockDim.x) {
      const float2 d = __ldg(reinterpret_cast<const float2*>(dihedrals) + row * L + j);
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
      uint4* dst = reinterpret_cast<uint4*>(xh + (row * L + j) * PT_DIHP);
#pragma unroll
      for (int q = 0; q < 3; ++q)
        dst[q] = make_uint4(pt_pk(f[8 * q], f[8 * q + 1]), pt_pk(f[8 * q + 2], f[8 * q + 3]), pt_pk(f[8 * q + 4], f[8 * q + 5]),
                            pt_pk(f[8 * q + 6], f[8 * q + 7]));
      dst[3] = make_uint4(0, 0, 0, 0);
    }
  }
}

// Class sums of the per-pair gradient without per-element atomics.  A persistent block owns a contiguous range of query
// rows (b, i); per patch it sorts the keys j by residue type (21 classes) and by residue index.  For a row, thread group
// (class s', 8 channels) walks the keys of its type class and adds into S_type[s_i*21 + s'] - an address only it ever
// touches; the clamped residue-index offset is monotone in the sorted residue index, so every offset class is a
// contiguous range of the sorted keys found by binary search: one thread group per interior offset (usually a single
// key), four per saturated end class and two per type class (partial sums meet through warp shuffles).
// The gradient rows (L x 128 bytes, contiguous) arrive by 1-D bulk copies, a ring of up to four rows ahead.
// Each block leaves its tables in `partials`; a second small kernel sums them (no global atomics).
// dynamic smem: n_stages row tiles [L][64] bf16 | S_type [441][64] | S_rel [n_rel][64] fp32 | seq, ridx, type-sorted j,
//               ridx-sorted r, ridx-sorted j [L] int | chain [L] float | type offsets [24] int | mbarriers
// warps 0-10: residue-type classes (16 lanes each); warps 11-26: interior offsets (8 lanes each); warps 27-28: the two
// saturated offset classes (a warp each, four parts)
constexpr int PT_TG_THREADS = 928, PT_TG_TYPE_THREADS = 352, PT_TG_INNER_THREADS = 512;

__global__ void __launch_bounds__(PT_TG_THREADS) pair_table_grad_kernel(
    const __nv_bfloat16* __restrict__ g1, const int64_t* __restrict__ seq, const int64_t* __restrict__ residue_idx,
    const int64_t* __restrict__ chain_idx, int64_t n_rows, int L, int max_dist, int n_stages, float* __restrict__ partials) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int n_rel = 2 * max_dist + 1;
  const int n_type = PT_V * PT_V * PT_C, n_all = n_type + n_rel * PT_C;
  const uint32_t tile_bytes = (uint32_t)L * PT_C * 2;
  float* s_acc = reinterpret_cast<float*>(s_raw + n_stages * tile_bytes);
  int* s_seq = reinterpret_cast<int*>(s_acc + n_all);
  int* s_ridx = s_seq + L;
  int* s_tj = s_ridx + L;       // keys sorted by residue type
  int* s_sr = s_tj + L;         // residue indices, ascending
  int* s_sj = s_sr + L;         // the key of each sorted residue index
  float* s_chain = reinterpret_cast<float*>(s_sj + L);
  int* s_off = reinterpret_cast<int*>(s_chain + L);   // [22] start of each type class in s_tj
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_off + 24);
  const int tid = threadIdx.x;
  for (int i = tid; i < n_all; i += blockDim.x) s_acc[i] = 0.f;
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t row_lo = blockIdx.x * per, row_hi = min(n_rows, row_lo + per);
  auto fetch = [&](int64_t row, int buf) {
    sm100::mbar_arrive_expect_tx(&bars[buf], tile_bytes);
    sm100::bulk_load_1d(s_raw + buf * tile_bytes, g1 + row * L * PT_C, tile_bytes, &bars[buf]);
  };
  if (tid == 0) {
    for (int k = 0; k < n_stages; ++k) sm100::mbar_init(&bars[k], 1);
    sm100::fence_barrier_init();
    for (int k = 0; k < n_stages - 1; ++k)
      if (row_lo + k < row_hi) fetch(row_lo + k, k);
  }
  int64_t cur_b = -1;
  int it = 0;
  for (int64_t row = row_lo; row < row_hi; ++row, ++it) {
    const int64_t b = row / L;
    const int buf = it % n_stages;
    __syncthreads();                         // the previous row's readers are done with their tile; barrier init visible
    if (tid == 0 && row + n_stages - 1 < row_hi) fetch(row + n_stages - 1, (it + n_stages - 1) % n_stages);
    if (b != cur_b) {
      cur_b = b;
      __syncthreads();                       // every reader of the previous patch's lists is done
      for (int j = tid; j < L; j += blockDim.x) {
        s_seq[j] = (int)seq[b * L + j];
        s_ridx[j] = (int)residue_idx[b * L + j];
        s_chain[j] = (float)chain_idx[b * L + j];
      }
      __syncthreads();
      for (int j = tid; j < L; j += blockDim.x) {        // ranks by comparison: L <= 512 keys, once per patch
        const int sj = s_seq[j], rj = s_ridx[j];
        int rank_t = 0, rank_r = 0;
        for (int k = 0; k < L; ++k) {
          const int sk = s_seq[k], rk = s_ridx[k];
          rank_t += (sk < sj) || (sk == sj && k < j);
          rank_r += (rk < rj) || (rk == rj && k < j);
        }
        s_tj[rank_t] = j;
        s_sr[rank_r] = rj;
        s_sj[rank_r] = j;
      }
      if (tid <= PT_V) {
        int c = 0;
        for (int k = 0; k < L; ++k) c += s_seq[k] < tid;
        s_off[tid] = c;
      }
      __syncthreads();
    }
    const int i = (int)(row - b * L);
    const int si = s_seq[i], ri = s_ridx[i];
    sm100::mbar_wait(&bars[buf], (it / n_stages) & 1);
    const __nv_bfloat16* grow = reinterpret_cast<const __nv_bfloat16*>(s_raw + buf * tile_bytes);
    const int q = tid & 7;
    auto lower = [&](int v) {            // first position of the ascending residue indices with s_sr >= v
      int lo_ = 0, n = L;
      while (n > 0) { const int h = n >> 1; if (s_sr[lo_ + h] < v) { lo_ += h + 1; n -= h + 1; } else n = h; }
      return lo_;
    };
    if (tid < PT_TG_TYPE_THREADS) {
      // residue-type classes: two 8-lane parts per class interleaved over its keys, combined by a shuffle (shared-memory
      // float atomics are compare-and-swap loops: ~500 cycles each under contention)
      const int cls = tid >> 4, part = (tid >> 3) & 1;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (cls < PT_V) {
        const int k1 = s_off[cls + 1];
#pragma unroll 4
        for (int k = s_off[cls] + part; k < k1; k += 2) {
          float g[8];
          pt_load_row8(grow + s_tj[k] * PT_C, q, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += g[e];
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      if (cls < PT_V && part == 0) {
        pt_add_row8(s_acc + (si * PT_V + cls) * PT_C, q, acc, 1.f);      // only this thread group ever touches this row
      }
    } else if (tid < PT_TG_TYPE_THREADS + PT_TG_INNER_THREADS) {
      // interior offsets o = cls - max_dist: the keys with r_j == r_i - o, a contiguous run of the sorted residue indices
      const float ci = s_chain[i];
      for (int cls = 1 + ((tid - PT_TG_TYPE_THREADS) >> 3); cls < n_rel - 1; cls += PT_TG_INNER_THREADS >> 3) {
        const int target = ri - (cls - max_dist);
        const int k1 = lower(target + 1);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = lower(target); k < k1; ++k) {
          const int j = s_sj[k];
          const float cj = s_chain[j];
          float g[8];
          pt_load_row8(grow + j * PT_C, q, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(cj, g[e], acc[e]);
        }
        pt_add_row8(s_acc + n_type + cls * PT_C, q, acc, ci);
      }
    } else {
      // the two saturated classes: one warp each, four 8-lane parts interleaved over the run, combined by shuffles
      const float ci = s_chain[i];
      const int w = (tid - PT_TG_TYPE_THREADS - PT_TG_INNER_THREADS) >> 5, part = (tid & 31) >> 3;
      int k0, k1;
      if (w == 0) { k0 = lower(ri + max_dist); k1 = L; }          // class 0: r_i - r_j <= -max_dist
      else { k0 = 0; k1 = lower(ri - max_dist + 1); }             // class n_rel - 1: r_i - r_j >= max_dist
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int k = k0 + part; k < k1; k += 4) {
        const int j = s_sj[k];
        const float cj = s_chain[j];
        float g[8];
        pt_load_row8(grow + j * PT_C, q, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(cj, g[e], acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
      }
      if (part == 0) {
        pt_add_row8(s_acc + n_type + (w == 0 ? 0 : n_rel - 1) * PT_C, q, acc, ci);
      }
    }
  }
  __syncthreads();
  float* dst = partials + (int64_t)blockIdx.x * n_all;
  for (int i = tid; i < n_all; i += blockDim.x) dst[i] = s_acc[i];
}

// out[i] += sum over blocks of partials[blk][i]  (out = s_type followed by s_rel)
__global__ void __launch_bounds__(256) pair_table_reduce_kernel(const float* __restrict__ partials, int n_blocks, int n_type, int n_all,
                                                                float* __restrict__ s_type, float* __restrict__ s_rel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_all) return;
  float t = 0.f;
  for (int k = 0; k < n_blocks; ++k) t += partials[(int64_t)k * n_all + i];
  if (i < n_type) s_type[i] += t; else s_rel[i - n_type] += t;
}

// g_out = g_in where y > 0 else 0 (ReLU backward on bf16 rows of 64 channels), with the column sums of g_out - the bias
// gradient of the layer - accumulated into colsum[64] in the same pass.
__global__ void __launch_bounds__(256) relu_bwd_colsum_kernel(const __nv_bfloat16* g_in /* may alias g_out (in-place) */, const __nv_bfloat16* __restrict__ y,
                                                              int64_t n_chunks, __nv_bfloat16* g_out,
                                                              float* __restrict__ colsum) {
  __shared__ float s_part[256][9];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;       // a multiple of 8: a thread keeps its 8 channels
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chunks; idx += stride) {
    float g[8], a[8];
    pt_unpack8(reinterpret_cast<const uint4*>(g_in)[idx], g);   // plain load: g_in may be g_out
    pt_unpack8(__ldg(reinterpret_cast<const uint4*>(y) + idx), a);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      g[e] = a[e] > 0.f ? g[e] : 0.f;
      acc[e] += g[e];
    }
    reinterpret_cast<uint4*>(g_out)[idx] = make_uint4(pt_pk(g[0], g[1]), pt_pk(g[2], g[3]), pt_pk(g[4], g[5]), pt_pk(g[6], g[7]));
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_part[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int q = threadIdx.x >> 3, e = threadIdx.x & 7;
    float t = 0.f;
    for (int k = q; k < 256; k += 8) t += s_part[k][e];
    atomicAdd(colsum + q * 8 + e, t);
  }
}

// x[b,i,j,:] = 0 wherever residue i or residue j is masked out (the reference's `* pair_mask`, :309-311, for 0/1 masks).
// One thread per pair: with every residue valid (the usual case) the pass only reads the two mask bytes.
__global__ void __launch_bounds__(256) pair_zero_masked_kernel(__nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ res_mask,
                                                               int64_t n_pairs, int L) {
  const int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  const int64_t row = pair / L;       // b * L + i
  const int64_t b = row / L;
  const int j = (int)(pair - row * L);
  if (!(__ldg(res_mask + row) && __ldg(res_mask + b * L + j))) {
    uint4* dst = reinterpret_cast<uint4*>(x) + pair * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_uint4(0, 0, 0, 0);
  }
}

// out = sum of up to 8 bf16 tensors, accumulated in fp32, one pass (the pair-tensor gradients of the IPA layers)
struct SumPtrs { const uint4* p[8]; };
__global__ void __launch_bounds__(256) sum_bf16_kernel(SumPtrs src, int n_src, int64_t n_chunks, uint4* out /* may be one of the sources */) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_chunks; idx += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < n_src) {
        float v[8];
        pt_unpack8(src.p[k][idx], v);   // plain load: `out` may be one of the sources
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    }
    out[idx] = make_uint4(pt_pk(acc[0], acc[1]), pt_pk(acc[2], acc[3]), pt_pk(acc[4], acc[5]), pt_pk(acc[6], acc[7]));
  }
}

}  // namespace dab

using namespace dab;

extern "C" {

/* base_bf16[B,L,L,64] and the padded dihedral features xh_bf16[B,L,L,32] (columns 18..31 zero) of the first mlp layer of
 * PairEmbedding (reference diffab_pytorch.py:262-285,303-307); t_type_bf16[441,64] = E_type W1[:, :64]^T + b1,
 * t_rel_bf16[2*max_dist+1,64] = E_rel W1[:, 64:128]^T. */
int dab_pair_base_fwd(const int64_t* seq_masked, const int64_t* residue_idx, const int64_t* chain_idx,
                      const float* pairwise_dihedrals, const void* t_type_bf16, const void* t_rel_bf16, int B, int L,
                      int max_dist, void* base_bf16, void* xh_bf16, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 512 && max_dist >= 0 && max_dist <= 64, DAB_EUNSUPPORTED,
              "dab_pair_base_fwd: 0 <= L <= 512 and 0 <= max_dist <= 64 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(seq_masked && residue_idx && chain_idx && pairwise_dihedrals && t_type_bf16 && t_rel_bf16 && base_bf16 && xh_bf16,
              DAB_EINVAL, "dab_pair_base_fwd: null pointer");
  DAB_REQUIRE(aligned16(t_type_bf16) && aligned16(t_rel_bf16) && aligned16(base_bf16) && aligned16(xh_bf16) &&
                  (reinterpret_cast<uintptr_t>(pairwise_dihedrals) & 7) == 0,
              DAB_EINVAL, "dab_pair_base_fwd: misaligned pointer (tables / outputs 16 B, dihedrals 8 B)");
  int rows = 1;
  for (int r = 8; r > 1; r >>= 1)
    if (L % r == 0) { rows = r; break; }
  const size_t smem = (size_t)3 * L * 4;
  pair_base_fwd_kernel<<<(unsigned)((int64_t)B * L / rows), 256, smem, (cudaStream_t)stream>>>(
      seq_masked, residue_idx, chain_idx, pairwise_dihedrals, reinterpret_cast<const __nv_bfloat16*>(t_type_bf16),
      reinterpret_cast<const __nv_bfloat16*>(t_rel_bf16), L, max_dist, rows, reinterpret_cast<__nv_bfloat16*>(base_bf16),
      reinterpret_cast<__nv_bfloat16*>(xh_bf16));
  count_launch();
  return check_launch("dab_pair_base_fwd");
}

static int table_grad_grid(int B, int L) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_rows = (int64_t)B * L;
  return (int)(n_rows < n_sm ? n_rows : n_sm);
}

size_t dab_pair_table_grad_workspace_bytes(int B, int L, int max_dist) {
  if (B <= 0 || L <= 0 || max_dist < 0) return 0;
  return (size_t)table_grad_grid(B, L) * (PT_V * PT_V + 2 * max_dist + 1) * PT_C * 4;
}

/* s_type[441,64] += sum over pairs with pair type s_i*21+s_j of g1;  s_rel[2*max_dist+1,64] += sum over pairs with that
 * clamped residue-index offset of chain_i*chain_j*g1.  workspace: dab_pair_table_grad_workspace_bytes(B, L, max_dist). */
int dab_pair_table_grad(const void* g1_bf16, const int64_t* seq_masked, const int64_t* residue_idx, const int64_t* chain_idx,
                        int B, int L, int max_dist, float* s_type, float* s_rel, void* workspace, size_t workspace_bytes,
                        void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 256 && max_dist >= 1 && max_dist <= 64, DAB_EUNSUPPORTED,
              "dab_pair_table_grad: 0 <= L <= 256 and 1 <= max_dist <= 64 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(g1_bf16 && seq_masked && residue_idx && chain_idx && s_type && s_rel && workspace, DAB_EINVAL,
              "dab_pair_table_grad: null pointer");
  DAB_REQUIRE(aligned16(g1_bf16) && aligned16(workspace), DAB_EINVAL, "dab_pair_table_grad: g1 and workspace must be 16-byte aligned");
  DAB_REQUIRE(workspace_bytes >= dab_pair_table_grad_workspace_bytes(B, L, max_dist), DAB_EINVAL,
              "dab_pair_table_grad: workspace too small (%zu bytes)", workspace_bytes);
  const int n_type = PT_V * PT_V * PT_C, n_all = n_type + (2 * max_dist + 1) * PT_C;
  const size_t tile = (size_t)L * PT_C * 2, fixed = (size_t)n_all * 4 + (size_t)(6 * L + 24) * 4 + 64;
  int n_stages = (int)((220 * 1024 - fixed) / tile);
  n_stages = n_stages > 4 ? 4 : n_stages;
  DAB_REQUIRE(n_stages >= 2, DAB_EUNSUPPORTED, "dab_pair_table_grad: L too large for the shared-memory row ring");
  const size_t smem = n_stages * tile + fixed;
  DAB_ENSURE_SMEM(pair_table_grad_kernel, smem);
  const int64_t n_rows = (int64_t)B * L;
  const int grid = table_grad_grid(B, L);
  float* partials = reinterpret_cast<float*>(workspace);
  pair_table_grad_kernel<<<grid, PT_TG_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(g1_bf16), seq_masked, residue_idx, chain_idx, n_rows, L, max_dist, n_stages, partials);
  count_launch();
  if (int rc = check_launch("dab_pair_table_grad")) return rc;
  pair_table_reduce_kernel<<<(n_all + 255) / 256, 256, 0, (cudaStream_t)stream>>>(partials, grid, n_type, n_all, s_type, s_rel);
  count_launch();
  return check_launch("dab_pair_table_grad (reduce)");
}

/* g_out_bf16[n,64] = g_in_bf16 where y_bf16 > 0 else 0; colsum[64] += column sums of g_out (fp32 atomics).  g_out may alias g_in. */
int dab_relu_bwd_colsum(const void* g_in_bf16, const void* y_bf16, int64_t n, void* g_out_bf16, float* colsum, void* stream) {
  DAB_REQUIRE(n >= 0, DAB_EINVAL, "dab_relu_bwd_colsum: negative size");
  if (n == 0) return DAB_OK;
  DAB_REQUIRE(g_in_bf16 && y_bf16 && g_out_bf16 && colsum, DAB_EINVAL, "dab_relu_bwd_colsum: null pointer");
  DAB_REQUIRE(aligned16(g_in_bf16) && aligned16(y_bf16) && aligned16(g_out_bf16), DAB_EINVAL, "dab_relu_bwd_colsum: tensors must be 16-byte aligned");
  const int64_t n_chunks = n * 8;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (n_chunks + 255) / 256;
  if (grid > (int64_t)n_sm * 8) grid = (int64_t)n_sm * 8;
  relu_bwd_colsum_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(g_in_bf16), reinterpret_cast<const __nv_bfloat16*>(y_bf16), n_chunks,
      reinterpret_cast<__nv_bfloat16*>(g_out_bf16), colsum);
  count_launch();
  return check_launch("dab_relu_bwd_colsum");
}

/* out_bf16[n] = sum of the n_src (1..8) bf16 tensors src[k][n], accumulated in fp32; n a multiple of 8, 16-byte aligned. */
int dab_sum_bf16(const void* const* src, int n_src, int64_t n, void* out_bf16, void* stream) {
  DAB_REQUIRE(src && out_bf16 && n_src >= 1 && n_src <= 8 && n >= 0 && n % 8 == 0, DAB_EINVAL,
              "dab_sum_bf16: 1 <= n_src <= 8 tensors of n %% 8 == 0 elements required");
  if (n == 0) return DAB_OK;
  SumPtrs sp;
  for (int k = 0; k < 8; ++k) {
    sp.p[k] = reinterpret_cast<const uint4*>(src[k < n_src ? k : 0]);
    DAB_REQUIRE(sp.p[k] && aligned16(sp.p[k]), DAB_EINVAL, "dab_sum_bf16: null or misaligned source %d", k);
  }
  DAB_REQUIRE(aligned16(out_bf16), DAB_EINVAL, "dab_sum_bf16: output must be 16-byte aligned");
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_chunks = n / 8;
  int64_t grid = (n_chunks + 255) / 256;
  if (grid > (int64_t)n_sm * 16) grid = (int64_t)n_sm * 16;
  sum_bf16_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(sp, n_src, n_chunks, reinterpret_cast<uint4*>(out_bf16));
  count_launch();
  return check_launch("dab_sum_bf16");
}

/* x_bf16[B,L,L,64]: rows (b,i,j) with res_mask[b,i] == 0 or res_mask[b,j] == 0 are set to zero, in place. */
int dab_pair_zero_masked(void* x_bf16, const uint8_t* res_mask, int B, int L, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0, DAB_EINVAL, "dab_pair_zero_masked: negative size");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(x_bf16 && res_mask, DAB_EINVAL, "dab_pair_zero_masked: null pointer");
  DAB_REQUIRE(aligned16(x_bf16), DAB_EINVAL, "dab_pair_zero_masked: x must be 16-byte aligned");
  const int64_t n_pairs = (int64_t)B * L * L;
  pair_zero_masked_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(x_bf16), res_mask, n_pairs, L);
  count_launch();
  return check_launch("dab_pair_zero_masked");
}

}  // extern "C"
