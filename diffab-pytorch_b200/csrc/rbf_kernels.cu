// Distance radial-basis features of PairEmbedding for TRAINING (diffab_pytorch.py:287-294 of the reference):
//   rbf[b,i,j,k] = exp(-softplus(C[s_i * 21 + s_j, k]) * d[b,i,j,k]^2) * m[b,i,a] * m[b,j,a'],   k = a * 15 + a'
// In PyTorch this is an embedding gather of C, softplus, square, multiply, exp and a mask multiply - six passes over
// (B, L, L, 225) fp32 tensors (0.94 GB each at B = 64) forward and as many backward, plus a sort-based embedding
// backward.  Here: one pass forward (reads d, writes rbf as bf16, padded to 232 columns so that the following
// nn.Linear runs as an aligned tensor-core GEMM) and one pass backward (reads d and the upstream gradient,
// recomputes rbf, accumulates dC).  For a query row (b, i) only the 21 table rows s_i*21 .. s_i*21+20 can be hit, so
// their softplus values (forward; the 441 x 225 table is evaluated once per call) / gradient accumulators (backward)
// live in shared memory, thread = k.  Backward: persistent blocks visit the rows grouped by s_i (counting sort), so the
// accumulators go to d_coef once per block and residue type instead of once per row (38.7 M -> ~3 M atomics at B = 64).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace dab {

constexpr int RBF_V = 21, RBF_K = 225, RBF_KP = 232, RBF_A = 15, RBF_T = RBF_V * RBF_V * RBF_K;

// softplus(C) * log2(e) (threshold 20 as F.softplus) for the whole 441 x 225 table, once per call; `sig` = its derivative
__global__ void rbf_softplus_kernel(const float* __restrict__ coef, float* __restrict__ sp, float* __restrict__ sig) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= RBF_T) return;
  const float c = coef[i];
  sp[i] = (c > 20.f ? c : log1pf(expf(c))) * 1.4426950408889634f;     // log2(e) folded in: the kernels use ex2
  if (sig) sig[i] = c > 20.f ? 1.f : 1.f / (1.f + __expf(-c));
}

// rows (b, i) grouped by residue type s_i (counting sort, one block): consecutive rows of `order` share their 21 table rows
__global__ void __launch_bounds__(1024) rbf_sort_rows_kernel(const int64_t* __restrict__ seq, int n_rows, int* __restrict__ order) {
  __shared__ int s_cnt[RBF_V], s_pos[RBF_V];
  if (threadIdx.x < RBF_V) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) atomicAdd(&s_cnt[(int)seq[r]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < RBF_V; ++s) { s_pos[s] = acc; acc += s_cnt[s]; }
  }
  __syncthreads();
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) order[atomicAdd(&s_pos[(int)seq[r]], 1)] = r;
}

__device__ __forceinline__ float rbf_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Persistent blocks over contiguous ranges of `order` (forward: order == nullptr, rows in natural order); thread = column
// k = a*15 + a'.  Per patch the keys j are sorted by residue type into s_key (packed: j | type << 9 | atom mask << 14) and
// visited in that order, so the coefficient of (s_i, s_j, k) sits in a register for a whole run of keys and the backward
// pass accumulates in a register, touching shared memory once per run - the inner loop is one global load (+ one store /
// one more load) per element.
// dynamic smem: [21][225] softplus * log2(e) rows of the current s_i; backward: + [21][225] gradient accumulators
// Run switch of the backward pass (the keys are walked in residue-type order): flush the finished run's accumulator into
// its shared-memory slot and fetch the next run's coefficient.  Deliberately not inlined: inlined, its five instructions
// are predicated into every element of the loop; as a call they cost only at the ~21 run boundaries of a row.
__device__ __noinline__ float rbf_switch_run(float* s_g_col, const float* s_sp_col, int cur_t, int t, float acc) {
  if (cur_t >= 0) s_g_col[cur_t * RBF_K] += acc;
  return s_sp_col[t * RBF_K];
}

template <bool BWD, bool FULL>   // FULL: L is a multiple of the keys per iteration - no bounds checks in the inner loop
__global__ void __launch_bounds__(256) rbf_kernel(const float* __restrict__ dist, const int64_t* __restrict__ seq,
                                                  const uint8_t* __restrict__ atom_mask, const float* __restrict__ sp_table,
                                                  const float* __restrict__ sig_table, const int* __restrict__ order,
                                                  int n_rows, int L, int squared, __nv_bfloat16* __restrict__ rbf,
                                                  const __nv_bfloat16* __restrict__ grad, float* __restrict__ dcoef) {
  extern __shared__ float s_tab[];                 // softplus rows; backward: accumulators behind them
  float* s_sp = s_tab;
  float* s_g = s_tab + RBF_V * RBF_K;
  __shared__ __align__(16) unsigned s_key[512];    // keys sorted by residue type
  __shared__ int s_seq[512];
  __shared__ unsigned s_mask[512];
  const int tid = threadIdx.x;
  const int per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n_rows, lo + per);
  const int a = tid / RBF_A, ap = tid % RBF_A;
  const bool col = tid < RBF_K;
  int cur_s = -1;
  int64_t cur_b = -1;
  auto flush = [&]() {                             // accumulators of residue type cur_s -> d_coef (x d softplus / d c)
    if (BWD && cur_s >= 0 && col) {
#pragma unroll 1
      for (int s = 0; s < RBF_V; ++s) {
        const float g = s_g[s * RBF_K + tid];
        const int64_t at = ((int64_t)cur_s * RBF_V + s) * RBF_K + tid;
        if (g != 0.f) atomicAdd(dcoef + at, g * __ldg(sig_table + at));
      }
    }
  };
  for (int it = lo; it < hi; ++it) {
    const int64_t row = order ? order[it] : it;
    const int64_t b = row / L;
    const int si = (int)seq[row];
    if (si != cur_s || b != cur_b) __syncthreads();            // the previous row's readers of the tables are done
    if (si != cur_s) {
      flush();
      if (BWD) __syncthreads();                                 // flushed columns are re-zeroed by other threads below
      cur_s = si;
      const float* src = sp_table + (int64_t)si * RBF_V * RBF_K;
      for (int i = tid; i < RBF_V * RBF_K; i += blockDim.x) {
        s_sp[i] = __ldg(src + i);
        if (BWD) s_g[i] = 0.f;
      }
    }
    if (b != cur_b) {
      cur_b = b;
      for (int j = tid; j < L; j += blockDim.x) {
        s_seq[j] = (int)seq[b * L + j];
        unsigned m = 0;
        for (int c = 0; c < RBF_A; ++c) m |= (atom_mask[(b * L + j) * RBF_A + c] ? 1u : 0u) << c;
        s_mask[j] = m;
      }
      __syncthreads();
      for (int j = tid; j < L; j += blockDim.x) {   // rank by comparison: L <= 512 keys, once per patch
        const int sj = s_seq[j];
        int rank = 0;
        for (int k = 0; k < L; ++k) rank += (s_seq[k] < sj) || (s_seq[k] == sj && k < j);
        s_key[rank] = (unsigned)j | ((unsigned)sj << 9) | (s_mask[j] << 14);
      }
    }
    __syncthreads();
    const unsigned mi = s_mask[row - b * L];
    const bool ai = col && ((mi >> a) & 1u);
    const unsigned key_bit = ai ? 1u << (14 + ap) : 0u;       // this column's atom of key j present (and atom a of row i)
    const int c_col = col ? tid : 0;
    const float* dp = dist + row * L * RBF_K + c_col;
    const __nv_bfloat16* gp = BWD ? grad + row * L * RBF_KP + c_col : nullptr;
    __nv_bfloat16* op = BWD ? nullptr : rbf + row * L * RBF_KP + tid;
    constexpr int U = BWD ? 8 : 16;                    // keys per iteration: U independent loads in flight per thread
    int cur_t = -1;
    float c = 0.f, acc = 0.f;
    for (int j0 = 0; j0 < L; j0 += U) {
      unsigned kw[U];
      if (FULL || j0 + U <= L) {
#pragma unroll
        for (int v = 0; v < U / 4; ++v) {
          const uint4 k4 = *reinterpret_cast<const uint4*>(s_key + j0 + 4 * v);
          kw[4 * v] = k4.x; kw[4 * v + 1] = k4.y; kw[4 * v + 2] = k4.z; kw[4 * v + 3] = k4.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) kw[u] = j0 + u < L ? s_key[j0 + u] : 0xffffffffu;   // past the end: never valid
      }
      if (col) {
        float d[U], g[BWD ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool in = FULL || kw[u] != 0xffffffffu;
          const int j = kw[u] & 511;
          d[u] = in ? __ldg(dp + j * RBF_K) : 0.f;
          if (BWD) g[u] = in ? __bfloat162float(gp[j * RBF_KP]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!FULL && kw[u] == 0xffffffffu) break;
          const int t = (kw[u] >> 9) & 31;
          if (t != cur_t) {                         // uniform over the block: every thread walks the same keys
            if (BWD) {
              c = rbf_switch_run(s_g + tid, s_sp + tid, cur_t, t, acc);   // own column: no race
            } else {
              c = s_sp[t * RBF_K + tid];
            }
            acc = 0.f;
            cur_t = t;
          }
          const float d2 = squared ? d[u] : d[u] * d[u];
          const bool on = (kw[u] & key_bit) != 0u;
          if (!BWD) {
            op[(kw[u] & 511) * RBF_KP] = __float2bfloat16_rn(on ? rbf_ex2(-c * d2) : 0.f);
          } else if (on) {
            // d rbf / d C = rbf * (-d2) * sigmoid(C); the sigmoid factor is applied once, when the accumulators are flushed
            acc = fmaf(g[BWD ? u : 0] * rbf_ex2(-c * d2), -d2, acc);
          }
        }
      } else if (!BWD && tid < RBF_KP) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (FULL || kw[u] != 0xffffffffu) op[(kw[u] & 511) * RBF_KP] = __float2bfloat16_rn(0.f);
      }
    }
    if (BWD && col && cur_t >= 0) s_g[cur_t * RBF_K + tid] += acc;
  }
  __syncthreads();
  flush();
}

}  // namespace dab

using namespace dab;

static int rbf_grid(int n_rows, int blocks_per_sm) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const int g = n_sm * blocks_per_sm;
  return n_rows < g ? n_rows : g;
}

extern "C" {

/* Workspace of dab_rbf_fwd / dab_rbf_bwd: the softplus table and its derivative (2 x 441 x 225 fp32) and the row order
 * of the backward pass (B*L int32). */
size_t dab_rbf_workspace_bytes(int B, int L) {
  if (B < 0 || L < 0) return 0;
  return (size_t)2 * RBF_T * 4 + (size_t)B * L * 4;
}

/* rbf_bf16[B,L,L,232] (columns 225..231 zero) from distmat[B,L,L,225] fp32 (distances, or squared distances if
 * `squared`), seq_masked[B,L] int64 in [0, 21), atom_mask[B,L,15] uint8, coef[441,225] fp32 (pair2distcoef.weight). */
int dab_rbf_fwd(const float* distmat, const int64_t* seq_masked, const uint8_t* atom_mask, const float* coef, int B, int L,
                int squared, void* rbf_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 512, DAB_EUNSUPPORTED, "dab_rbf_fwd: 0 <= L <= 512 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(distmat && seq_masked && atom_mask && coef && rbf_bf16 && workspace, DAB_EINVAL, "dab_rbf_fwd: null pointer");
  DAB_REQUIRE(workspace_bytes >= dab_rbf_workspace_bytes(B, L) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, DAB_EINVAL,
              "dab_rbf_fwd: workspace too small or misaligned");
  float* sp = reinterpret_cast<float*>(workspace);
  rbf_softplus_kernel<<<(RBF_T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(coef, sp, nullptr);
  count_launch();
  const int n_rows = B * L;
  // (the check-free variant of the forward kernel compiles to 32 registers with its loads serialised and is slower)
  rbf_kernel<false, false><<<rbf_grid(n_rows, 8), 256, RBF_V * RBF_K * 4, (cudaStream_t)stream>>>(
      distmat, seq_masked, atom_mask, sp, nullptr, nullptr, n_rows, L, squared, reinterpret_cast<__nv_bfloat16*>(rbf_bf16),
      nullptr, nullptr);
  count_launch();
  return check_launch("dab_rbf_fwd");
}

/* d_coef[441,225] += d rbf / d coef contracted with grad_bf16[B,L,L,232] (accumulated into; fp32 atomics, one flush per
 * block and residue type: rows are visited grouped by s_i). */
int dab_rbf_bwd(const void* grad_bf16, const float* distmat, const int64_t* seq_masked, const uint8_t* atom_mask,
                const float* coef, int B, int L, int squared, float* d_coef, void* workspace, size_t workspace_bytes,
                void* stream) {
  DAB_REQUIRE(B >= 0 && L >= 0 && L <= 512, DAB_EUNSUPPORTED, "dab_rbf_bwd: 0 <= L <= 512 required");
  if ((int64_t)B * L == 0) return DAB_OK;
  DAB_REQUIRE(grad_bf16 && distmat && seq_masked && atom_mask && coef && d_coef && workspace, DAB_EINVAL, "dab_rbf_bwd: null pointer");
  DAB_REQUIRE(workspace_bytes >= dab_rbf_workspace_bytes(B, L) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, DAB_EINVAL,
              "dab_rbf_bwd: workspace too small or misaligned");
  float* sp = reinterpret_cast<float*>(workspace);
  float* sig = sp + RBF_T;
  int* order = reinterpret_cast<int*>(sig + RBF_T);
  const int n_rows = B * L;
  rbf_softplus_kernel<<<(RBF_T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(coef, sp, sig);
  count_launch();
  rbf_sort_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(seq_masked, n_rows, order);
  count_launch();
  auto kernel = L % 8 == 0 ? rbf_kernel<true, true> : rbf_kernel<true, false>;
  kernel<<<rbf_grid(n_rows, 4), 256, 2 * RBF_V * RBF_K * 4, (cudaStream_t)stream>>>(
      distmat, seq_masked, atom_mask, sp, sig, order, n_rows, L, squared, nullptr,
      reinterpret_cast<const __nv_bfloat16*>(grad_bf16), d_coef);
  count_launch();
  return check_launch("dab_rbf_bwd");
}

}  // extern "C"
