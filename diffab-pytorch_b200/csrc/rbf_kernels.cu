// Distance radial-basis features of PairEmbedding for TRAINING (diffab_pytorch.py:287-294 of the reference):
//   rbf[b,i,j,k] = exp(-softplus(C[s_i * 21 + s_j, k]) * d[b,i,j,k]^2) * m[b,i,a] * m[b,j,a'],   k = a * 15 + a'
// In PyTorch this is an embedding gather of C, softplus, square, multiply, exp and a mask multiply - six passes over
// (B, L, L, 225) fp32 tensors (0.94 GB each at B = 64) forward and as many backward, plus a sort-based embedding
// backward.  Here: one pass forward (reads d, writes rbf as bf16, padded to 232 columns so that the following
// nn.Linear runs as an aligned tensor-core GEMM) and one pass backward (reads d and the upstream gradient,
// recomputes rbf, accumulates dC).  One block per query row (b, i): only the 21 table rows s_i*21 .. s_i*21+20 can
// be hit, so their softplus values (forward) / gradient accumulators (backward) live in shared memory, thread = k.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace dab {

constexpr int RBF_V = 21, RBF_K = 225, RBF_KP = 232, RBF_A = 15;

__device__ __forceinline__ float softplus_f(float c) { return c > 20.f ? c : log1pf(expf(c)); }

template <bool BWD>
__global__ void __launch_bounds__(256) rbf_kernel(const float* __restrict__ dist, const int64_t* __restrict__ seq,
                                                  const uint8_t* __restrict__ atom_mask, const float* __restrict__ coef,
                                                  int L, int squared, __nv_bfloat16* __restrict__ rbf,
                                                  const __nv_bfloat16* __restrict__ grad, float* __restrict__ dcoef) {
  extern __shared__ float s_tab[];                 // [21][225] softplus(C); backward: [21][225] gradient accumulators first
  __shared__ int s_seq[512];
  __shared__ unsigned s_mask[512];
  const int64_t row = blockIdx.x;                  // (b, i)
  const int64_t b = row / L;
  const int tid = threadIdx.x;
  for (int j = tid; j < L; j += blockDim.x) {
    s_seq[j] = (int)seq[b * L + j];
    unsigned m = 0;
    for (int a = 0; a < RBF_A; ++a) m |= (atom_mask[(b * L + j) * RBF_A + a] ? 1u : 0u) << a;
    s_mask[j] = m;
  }
  const int si = (int)seq[row];
  unsigned mi = 0;
  for (int a = 0; a < RBF_A; ++a) mi |= (atom_mask[row * RBF_A + a] ? 1u : 0u) << a;
  const float* crow = coef + (int64_t)si * RBF_V * RBF_K;
  float sp[RBF_V];                                  // softplus of this thread's column for the 21 possible s_j
  if (tid < RBF_K) {
#pragma unroll
    for (int s = 0; s < RBF_V; ++s) sp[s] = softplus_f(__ldg(crow + s * RBF_K + tid));
  }
  float* s_sp = BWD ? s_tab + RBF_V * RBF_K : s_tab;   // backward: accumulators first, softplus table second
  if (BWD) {
    for (int i = tid; i < RBF_V * RBF_K; i += blockDim.x) s_tab[i] = 0.f;
  }
  if (tid < RBF_K) {
#pragma unroll
    for (int s = 0; s < RBF_V; ++s) s_sp[s * RBF_K + tid] = sp[s];
  }
  __syncthreads();
  const int a = tid / RBF_A, ap = tid % RBF_A;
  const bool ai = tid < RBF_K && ((mi >> a) & 1u);
  constexpr int U = 8;                               // keys per iteration: U independent loads in flight per thread
  for (int j0 = 0; j0 < L; j0 += U) {
    const int64_t p0 = row * L + j0;
    if (tid < RBF_K) {
      float d[U], g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = j0 + u < L;
        d[u] = in ? __ldg(dist + (p0 + u) * RBF_K + tid) : 0.f;
        if (BWD) g[u] = in ? __bfloat162float(grad[(p0 + u) * RBF_KP + tid]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u >= L) break;
        const int sj = s_seq[j0 + u];
        const float d2 = squared ? d[u] : d[u] * d[u];
        const bool on = ai && ((s_mask[j0 + u] >> ap) & 1u);
        if (!BWD) {
          const float v = on ? __expf(-s_sp[sj * RBF_K + tid] * d2) : 0.f;
          rbf[(p0 + u) * RBF_KP + tid] = __float2bfloat16_rn(v);
        } else if (on) {
          // d rbf / d C = rbf * (-d2) * sigmoid(C); the sigmoid factor is applied once at the end
          s_tab[sj * RBF_K + tid] += g[u] * __expf(-s_sp[sj * RBF_K + tid] * d2) * (-d2);   // own column: no race
        }
      }
    } else if (!BWD && tid < RBF_KP) {
      for (int u = 0; u < U && j0 + u < L; ++u) rbf[(p0 + u) * RBF_KP + tid] = __float2bfloat16_rn(0.f);
    }
  }
  if (BWD && tid < RBF_K) {
#pragma unroll 1
    for (int s = 0; s < RBF_V; ++s) {
      const float g = s_tab[s * RBF_K + tid];
      if (g != 0.f) {
        const float c = __ldg(crow + s * RBF_K + tid);
        const float sig = c > 20.f ? 1.f : 1.f / (1.f + __expf(-c));   // d softplus / d c (threshold as F.softplus)
        atomicAdd(dcoef + ((int64_t)si * RBF_V + s) * RBF_K + tid, g * sig);
      }
    }
  }
}

}  // namespace dab

using namespace dab;

extern "C" {

/* rbf_bf16[B,L,L,232] (columns 225..231 zero) from distmat[B,L,L,225] fp32 (distances, or squared distances if
 * `squared`), seq_masked[B,L] int64 in [0, 21), atom_mask[B,L,15] uint8, coef[441,225] fp32 (pair2distcoef.weight). */
int dab_rbf_fwd(const float* distmat, const int64_t* seq_masked, const uint8_t* atom_mask, const float* coef, int B, int L,
                int squared, void* rbf_bf16, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 512, DAB_EUNSUPPORTED, "dab_rbf_fwd: 0 <= L <= 512 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(distmat && seq_masked && atom_mask && coef && rbf_bf16, DAB_EINVAL, "dab_rbf_fwd: null pointer");
  rbf_kernel<false><<<B * L, 256, RBF_V * RBF_K * 4, (cudaStream_t)stream>>>(
      distmat, seq_masked, atom_mask, coef, L, squared, reinterpret_cast<__nv_bfloat16*>(rbf_bf16), nullptr, nullptr);
  count_launch();
  return check_launch("dab_rbf_fwd");
}

/* d_coef[441,225] += d rbf / d coef contracted with grad_bf16[B,L,L,232] (accumulated into; fp32 atomics). */
int dab_rbf_bwd(const void* grad_bf16, const float* distmat, const int64_t* seq_masked, const uint8_t* atom_mask,
                const float* coef, int B, int L, int squared, float* d_coef, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 512, DAB_EUNSUPPORTED, "dab_rbf_bwd: 0 <= L <= 512 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(grad_bf16 && distmat && seq_masked && atom_mask && coef && d_coef, DAB_EINVAL, "dab_rbf_bwd: null pointer");
  rbf_kernel<true><<<B * L, 256, 2 * RBF_V * RBF_K * 4, (cudaStream_t)stream>>>(
      distmat, seq_masked, atom_mask, coef, L, squared, nullptr, reinterpret_cast<const __nv_bfloat16*>(grad_bf16), d_coef);
  count_launch();
  return check_launch("dab_rbf_bwd");
}

}  // extern "C"
