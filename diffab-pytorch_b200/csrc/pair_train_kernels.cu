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
    //      4 lanes per pair, lane q writes 16 bytes of the 64-byte padded feature row
    for (int idx = tid; idx < L * 4; idx += blockDim.x) {
      const int j = idx >> 2, q = idx & 3;
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
      uint4 o = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int qq = 0; qq < 3; ++qq)
        if (q == qq)
          o = make_uint4(pt_pk(f[8 * qq], f[8 * qq + 1]), pt_pk(f[8 * qq + 2], f[8 * qq + 3]), pt_pk(f[8 * qq + 4], f[8 * qq + 5]),
                         pt_pk(f[8 * qq + 6], f[8 * qq + 7]));
      *(reinterpret_cast<uint4*>(xh + (row * L + j) * PT_DIHP) + q) = o;
    }
  }
}

// Persistent blocks; dynamic smem: S_type [441][64] fp32 | S_rel [n_rel][64] fp32 accumulators (shared-memory atomics),
// flushed once per block with global atomics.
__global__ void __launch_bounds__(512) pair_table_grad_kernel(
    const __nv_bfloat16* __restrict__ g1, const int64_t* __restrict__ seq, const int64_t* __restrict__ residue_idx,
    const int64_t* __restrict__ chain_idx, int64_t n_rows, int L, int max_dist, float* __restrict__ s_type,
    float* __restrict__ s_rel) {
  extern __shared__ float s_acc[];
  const int n_type = PT_V * PT_V * PT_C, n_all = n_type + (2 * max_dist + 1) * PT_C;
  const int tid = threadIdx.x;
  for (int i = tid; i < n_all; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int64_t b = row / L;
    const int si = (int)__ldg(seq + row), ri = (int)__ldg(residue_idx + row);
    const float ci = (float)__ldg(chain_idx + row);
    for (int idx = tid; idx < L * 8; idx += blockDim.x) {
      const int j = idx >> 3, q = idx & 7;
      const int64_t rj = b * L + j;
      const int pt = si * PT_V + (int)__ldg(seq + rj);
      int rel = ri - (int)__ldg(residue_idx + rj);
      rel = max(-max_dist, min(max_dist, rel)) + max_dist;
      const float cp = ci * (float)__ldg(chain_idx + rj);
      float g[8];
      pt_unpack8(__ldg(reinterpret_cast<const uint4*>(g1 + (row * L + j) * PT_C) + q), g);
      float* at = s_acc + pt * PT_C + q * 8;
      float* ar = s_acc + n_type + rel * PT_C + q * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (g[e] != 0.f) {
          atomicAdd(at + e, g[e]);
          atomicAdd(ar + e, cp * g[e]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n_all; i += blockDim.x) {
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(i < n_type ? s_type + i : s_rel + (i - n_type), v);
  }
}

// x[b,i,j,:] = 0 wherever residue i or residue j is masked out (the reference's `* pair_mask`, :309-311, for 0/1 masks).
__global__ void __launch_bounds__(256) pair_zero_masked_kernel(__nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ res_mask,
                                                               int64_t n_chunks, int L) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks) return;
  const int64_t pair = idx >> 3;
  const int64_t row = pair / L;       // b * L + i
  const int64_t b = row / L;
  const int j = (int)(pair - row * L);
  if (!(__ldg(res_mask + row) && __ldg(res_mask + b * L + j))) reinterpret_cast<uint4*>(x)[idx] = make_uint4(0, 0, 0, 0);
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

/* s_type[441,64] += sum over pairs with pair type s_i*21+s_j of g1;  s_rel[2*max_dist+1,64] += sum over pairs with that
 * clamped residue-index offset of chain_i*chain_j*g1 (fp32 atomics; the caller zeroes the outputs). */
int dab_pair_table_grad(const void* g1_bf16, const int64_t* seq_masked, const int64_t* residue_idx, const int64_t* chain_idx,
                        int B, int L, int max_dist, float* s_type, float* s_rel, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && max_dist >= 0 && max_dist <= 64, DAB_EUNSUPPORTED, "dab_pair_table_grad: 0 <= max_dist <= 64 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(g1_bf16 && seq_masked && residue_idx && chain_idx && s_type && s_rel, DAB_EINVAL, "dab_pair_table_grad: null pointer");
  DAB_REQUIRE(aligned16(g1_bf16), DAB_EINVAL, "dab_pair_table_grad: g1 must be 16-byte aligned");
  const size_t smem = (size_t)(PT_V * PT_V + 2 * max_dist + 1) * PT_C * 4;
  DAB_ENSURE_SMEM(pair_table_grad_kernel, smem);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_rows = (int64_t)B * L;
  const int grid = (int)(n_rows < n_sm ? n_rows : n_sm);
  pair_table_grad_kernel<<<grid, 512, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(g1_bf16), seq_masked,
                                                                     residue_idx, chain_idx, n_rows, L, max_dist, s_type, s_rel);
  count_launch();
  return check_launch("dab_pair_table_grad");
}

/* x_bf16[B,L,L,64]: rows (b,i,j) with res_mask[b,i] == 0 or res_mask[b,j] == 0 are set to zero, in place. */
int dab_pair_zero_masked(void* x_bf16, const uint8_t* res_mask, int B, int L, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0, DAB_EINVAL, "dab_pair_zero_masked: negative size");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(x_bf16 && res_mask, DAB_EINVAL, "dab_pair_zero_masked: null pointer");
  DAB_REQUIRE(aligned16(x_bf16), DAB_EINVAL, "dab_pair_zero_masked: x must be 16-byte aligned");
  const int64_t n_chunks = (int64_t)B * L * L * 8;
  pair_zero_masked_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(x_bf16), res_mask, n_chunks, L);
  count_launch();
  return check_launch("dab_pair_zero_masked");
}

}  // extern "C"
