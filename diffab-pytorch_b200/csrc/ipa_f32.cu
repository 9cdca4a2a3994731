// Shape-generic fp32 invariant point attention, forward and backward (sm_100a, CUDA cores).
//
// Replaces InvariantPointAttentionLayer.forward and its autograd backward
// (/root/reference/diffab_pytorch/diffab_pytorch.py:315-465).  This is the "<= 1e-4" fp32 path of
// north_star, the path for shapes other than the train.py configuration, the training path, and
// the on-GPU cross-check of the tcgen05 kernel (ipa_sm100.cu).  It evaluates the point term in the
// reference's direct form sum (q - k)^2, so it carries no cancellation error.
//
// Forward = 4 launches:  projections (6 GEMM segments) -> frame transform of the points ->
// attention core (one CTA per (patch, block of IB query rows); the pair row e[b,i,:,:] is read once
// for the bias and once more, from L1/L2, for the pair aggregation; logits never leave shared
// memory) -> to_out GEMM.
// Backward = to_out grads -> query-side core (recomputes the softmax, writes de, dq, attention and
// dlogit tiles) -> key-side core (dk, dv from the tiles) -> inverse frame -> projection grads.
#include "common.cuh"

namespace dab {

struct IpaD {
  int B, L, D, C, H, ds, Pq, Pv;
  int NS, NQ, NV, NPROJ, NCAT;
  int o_qs, o_ks, o_vs, o_qp, o_kp, o_vp;   // offsets inside a projection row
  int c_pair, c_point, c_norm;              // offsets inside a concat row
  float ss, sp, st;
};

static IpaD make_dims(const DabIpaDims* d) {
  IpaD r;
  r.B = d->B; r.L = d->L; r.D = d->D; r.C = d->C; r.H = d->H; r.ds = d->ds; r.Pq = d->Pq; r.Pv = d->Pv;
  r.NS = r.H * r.ds; r.NQ = r.H * r.Pq * 3; r.NV = r.H * r.Pv * 3;
  r.NPROJ = 3 * r.NS + 2 * r.NQ + r.NV;
  r.NCAT = r.NS + r.H * r.C + r.NV + r.H * r.Pv;
  r.o_qs = 0; r.o_ks = r.NS; r.o_vs = 2 * r.NS; r.o_qp = 3 * r.NS; r.o_kp = 3 * r.NS + r.NQ; r.o_vp = 3 * r.NS + 2 * r.NQ;
  r.c_pair = r.NS; r.c_point = r.NS + r.H * r.C; r.c_norm = r.c_point + r.NV;
  r.ss = 1.0f / sqrtf((float)r.ds);            // diffab_pytorch.py:359
  r.sp = 1.0f / sqrtf(4.5f * (float)r.Pq);     // :372
  r.st = 1.0f / sqrtf(3.0f);                   // :385-387 (use_pair_bias=True)
  return r;
}

constexpr int MAXH = 8;
constexpr int IBF = 8;   // query rows per CTA, forward
constexpr int IBB = 4;   // query rows per CTA, backward (more live state per row)
constexpr int JBB = 8;   // key columns per CTA, key-side backward
constexpr int NT = 128;  // threads per attention CTA

// ------------------------------------------------------------------------------------------
// Generic tiled SGEMM: C[m,n] (+)= sum_k A(m,k) B(k,n) (+ bias[n]);  A(m,k) = A[m*sam + k*sak],
// B(k,n) = Bp[k*sbk + n*sbn].  64x64x16 tiles, 256 threads, 4x4 register tile.  Split-K over
// gridDim.z accumulates with atomicAdd (used for the weight gradients, K = B*L).
// mode: 0 overwrite, 1 add to existing C, 2 atomicAdd.
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                    const float* __restrict__ Bp, int64_t sbk, int64_t sbn,
                                                    float* __restrict__ Cp, int64_t ldc, const float* __restrict__ bias,
                                                    int M, int N, int K, int kchunk, int mode) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool a_kfast = (sak == 1), b_nfast = (sbn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = threadIdx.x + i * 256;
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? __ldg(A + gm * sam + gk * sak) : 0.f;
      int nn, k2;
      if (b_nfast) { nn = e & 63; k2 = e >> 6; } else { k2 = e & 15; nn = e >> 4; }
      int gn = n0 + nn, gk2 = k0 + k2;
      Bs[k2][nn] = (gn < N && gk2 < kend) ? __ldg(Bp + gk2 * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += __ldg(bias + gn);
      float* dst = Cp + gm * ldc + gn;
      if (mode == 0) *dst = v;
      else if (mode == 1) *dst += v;
      else atomicAdd(dst, v);
    }
  }
}

static void sgemm(cudaStream_t s, const float* A, int64_t sam, int64_t sak, const float* Bp, int64_t sbk, int64_t sbn,
                  float* Cp, int64_t ldc, const float* bias, int M, int N, int K, int mode, int splitk = 1) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  int kchunk = ((K + splitk - 1) / splitk + 15) / 16 * 16;
  int nz = (K + kchunk - 1) / kchunk;
  dim3 grid((N + 63) / 64, (M + 63) / 64, nz);
  sgemm_kernel<<<grid, 256, 0, s>>>(A, sam, sak, Bp, sbk, sbn, Cp, ldc, bias, M, N, K, kchunk, nz > 1 ? 2 : mode);
  count_launch();
}

// column sums: out[n] += sum_m A[m, n]   (d bias of to_out)
__global__ void colsum_kernel(const float* __restrict__ A, int M, int N, float* __restrict__ out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int rows_per = (M + gridDim.y - 1) / gridDim.y;
  int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc = 0.f;
  for (int m = m0; m < m1; ++m) acc += __ldg(A + (int64_t)m * N + n);
  atomicAdd(out + n, acc);
}

// ------------------------------------------------------------------------------------------
// Frame transform of the projected points, in place on the projection rows
// (euclidean_transform, diffab_pytorch.py:315-324: p_glob = p_loc @ R + t, row vectors).
// inverse_grad: d_loc = d_glob @ R^T (no translation) for the backward.
__global__ void __launch_bounds__(256) frame_kernel(float* __restrict__ proj, const float* __restrict__ R,
                                                    const float* __restrict__ t, IpaD d, int inverse_grad) {
  int npts = (2 * d.NQ + d.NV) / 3;
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)d.B * d.L * npts;
  if (idx >= total) return;
  int64_t row = idx / npts;
  int p = (int)(idx - row * npts);
  float* v = proj + row * d.NPROJ + d.o_qp + p * 3;
  const float* Rr = R + row * 9;
  float x = v[0], y = v[1], z = v[2];
  if (!inverse_grad) {
    const float* tr = t + row * 3;
    v[0] = x * Rr[0] + y * Rr[3] + z * Rr[6] + tr[0];
    v[1] = x * Rr[1] + y * Rr[4] + z * Rr[7] + tr[1];
    v[2] = x * Rr[2] + y * Rr[5] + z * Rr[8] + tr[2];
  } else {
    v[0] = x * Rr[0] + y * Rr[1] + z * Rr[2];
    v[1] = x * Rr[3] + y * Rr[4] + z * Rr[5];
    v[2] = x * Rr[6] + y * Rr[7] + z * Rr[8];
  }
}

// ------------------------------------------------------------------------------------------
// Logits of IB query rows against key column j (thread-private), reference direct form
// (diffab_pytorch.py:416-441).  qk[r][h] = <qs, ks>, bias[r][h] = e . Wpb, d2[r][h] = sum (qp - kp)^2.
template <int IB>
__device__ __forceinline__ void logits_for_column(const IpaD& d, const float* __restrict__ proj,
                                                  const float* __restrict__ e, const float* s_q, const float* s_wpb,
                                                  int b, int i0, int nrows, int j, float (&qk)[IB][MAXH],
                                                  float (&d2)[IB][MAXH]) {
  const int NQK = d.NS + d.NQ;
  const float* krow = proj + ((int64_t)b * d.L + j) * d.NPROJ;
#pragma unroll
  for (int r = 0; r < IB; ++r)
#pragma unroll
    for (int h = 0; h < MAXH; ++h) { qk[r][h] = 0.f; d2[r][h] = 0.f; }
  const bool vec = ((d.ds & 3) == 0) && ((d.NPROJ & 3) == 0) && ((d.NS & 3) == 0) && ((NQK & 3) == 0);
#pragma unroll
  for (int h = 0; h < MAXH; ++h) {
    if (h >= d.H) break;
    const float* kp = krow + d.o_ks + h * d.ds;
    if (vec) {
      for (int dd = 0; dd < d.ds; dd += 4) {
        float4 k4 = __ldg(reinterpret_cast<const float4*>(kp + dd));
#pragma unroll
        for (int r = 0; r < IB; ++r) {
          float4 q4 = *reinterpret_cast<const float4*>(s_q + r * NQK + h * d.ds + dd);
          qk[r][h] += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
        }
      }
    } else {
      for (int dd = 0; dd < d.ds; ++dd) {
        float kv = __ldg(kp + dd);
#pragma unroll
        for (int r = 0; r < IB; ++r) qk[r][h] += s_q[r * NQK + h * d.ds + dd] * kv;
      }
    }
    const float* kpp = krow + d.o_kp + h * d.Pq * 3;
    for (int pc = 0; pc < d.Pq * 3; ++pc) {
      float kv = __ldg(kpp + pc);
#pragma unroll
      for (int r = 0; r < IB; ++r) {
        float df = s_q[r * NQK + d.NS + h * d.Pq * 3 + pc] - kv;
        d2[r][h] += df * df;
      }
    }
  }
  // scale scalar part, then add the pair bias into qk
#pragma unroll
  for (int r = 0; r < IB; ++r)
#pragma unroll
    for (int h = 0; h < MAXH; ++h) qk[r][h] *= d.ss;
  const bool vec_c = (d.C & 3) == 0;
#pragma unroll
  for (int r = 0; r < IB; ++r) {
    if (r >= nrows) break;
    const float* er = e + (((int64_t)b * d.L + i0 + r) * d.L + j) * d.C;
    if (vec_c) {
      for (int c = 0; c < d.C; c += 4) {
        float4 ev = __ldg(reinterpret_cast<const float4*>(er + c));
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
          if (h >= d.H) break;
          float4 w4 = *reinterpret_cast<const float4*>(s_wpb + h * d.C + c);
          qk[r][h] += ev.x * w4.x + ev.y * w4.y + ev.z * w4.z + ev.w * w4.w;
        }
      }
    } else {
      for (int c = 0; c < d.C; ++c) {
        float ev = __ldg(er + c);
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
          if (h >= d.H) break;
          qk[r][h] += ev * s_wpb[h * d.C + c];
        }
      }
    }
  }
}

// softmax over j of s_logit rows [nrow][L]; one warp per row
__device__ __forceinline__ void softmax_rows(float* s_logit, int nrow, int L) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int row = warp; row < nrow; row += nwarp) {
    float* p = s_logit + row * L;
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, p[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < L; j += 32) { float v = expf(p[j] - m); p[j] = v; s += v; }
    s = warp_sum(s);
    float inv = 1.0f / s;
    for (int j = lane; j < L; j += 32) p[j] *= inv;
  }
}

// out[r][f] = sum_j s_w[r][h(f)][j] * V[j][f] for f in [0, nfeat), V rows inside projection rows
// at offset voff; fdiv = features per head.  Result handed to `sink(r, f, value)`.
template <int IB, typename Sink>
__device__ __forceinline__ void aggregate_rows(const IpaD& d, const float* __restrict__ proj, const float* s_w, int b,
                                               int voff, int nfeat, int fdiv, Sink sink) {
  for (int f = threadIdx.x; f < nfeat; f += blockDim.x) {
    int h = f / fdiv;
    float acc[IB];
#pragma unroll
    for (int r = 0; r < IB; ++r) acc[r] = 0.f;
    const float* vcol = proj + (int64_t)b * d.L * d.NPROJ + voff + f;
    for (int j = 0; j < d.L; ++j) {
      float v = __ldg(vcol + (int64_t)j * d.NPROJ);
#pragma unroll
      for (int r = 0; r < IB; ++r) acc[r] = fmaf(s_w[(r * d.H + h) * d.L + j], v, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < IB; ++r) sink(r, f, acc[r]);
  }
}

// ---- forward attention core ---------------------------------------------------------------
__global__ void __launch_bounds__(NT) ipa_attn_fwd_kernel(IpaD d, const float* __restrict__ proj,
                                                          const float* __restrict__ e, const float* __restrict__ R,
                                                          const float* __restrict__ t, const float* __restrict__ wpb,
                                                          const float* __restrict__ gamma, float* __restrict__ cat) {
  extern __shared__ __align__(16) float smem[];
  const int NQK = d.NS + d.NQ;
  float* s_q = smem;                           // [IBF][NQK]
  float* s_wpb = s_q + IBF * NQK;              // [H][C]
  float* s_attn = s_wpb + d.H * d.C;           // [IBF][H][L]
  float* s_og = s_attn + IBF * d.H * d.L;      // [IBF][NV]
  int b = blockIdx.y, i0 = blockIdx.x * IBF;
  int nrows = min(IBF, d.L - i0);
  for (int idx = threadIdx.x; idx < IBF * NQK; idx += blockDim.x) {
    int r = idx / NQK, c = idx - r * NQK;
    float v = 0.f;
    if (r < nrows) {
      const float* row = proj + ((int64_t)b * d.L + i0 + r) * d.NPROJ;
      v = (c < d.NS) ? row[d.o_qs + c] : row[d.o_qp + (c - d.NS)];
    }
    s_q[idx] = v;
  }
  for (int idx = threadIdx.x; idx < d.H * d.C; idx += blockDim.x) s_wpb[idx] = wpb[idx];
  __syncthreads();
  for (int j = threadIdx.x; j < d.L; j += blockDim.x) {
    float qk[IBF][MAXH], d2[IBF][MAXH];
    logits_for_column<IBF>(d, proj, e, s_q, s_wpb, b, i0, nrows, j, qk, d2);
#pragma unroll
    for (int r = 0; r < IBF; ++r)
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h >= d.H) break;
        s_attn[(r * d.H + h) * d.L + j] = d.st * (qk[r][h] - 0.5f * d.sp * __ldg(gamma + h) * d2[r][h]);
      }
  }
  __syncthreads();
  softmax_rows(s_attn, IBF * d.H, d.L);
  __syncthreads();
  // scalar values -> concat[0:NS]
  aggregate_rows<IBF>(d, proj, s_attn, b, d.o_vs, d.NS, d.ds, [&](int r, int f, float v) {
    if (r < nrows) cat[((int64_t)b * d.L + i0 + r) * d.NCAT + f] = v;
  });
  // point values (global frame) -> shared
  aggregate_rows<IBF>(d, proj, s_attn, b, d.o_vp, d.NV, d.Pv * 3, [&](int r, int f, float v) { s_og[r * d.NV + f] = v; });
  // pair values: out_pair[r][h][c] = sum_j attn[r][h][j] e[i0+r][j][c]   (:449-450)
  for (int task = threadIdx.x; task < IBF * d.C; task += blockDim.x) {
    int r = task / d.C, c = task - r * d.C;
    if (r >= nrows) continue;
    float acc[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) acc[h] = 0.f;
    const float* ecol = e + ((int64_t)b * d.L + i0 + r) * d.L * d.C + c;
    for (int j = 0; j < d.L; ++j) {
      float ev = __ldg(ecol + (int64_t)j * d.C);
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h >= d.H) break;
        acc[h] = fmaf(s_attn[(r * d.H + h) * d.L + j], ev, acc[h]);
      }
    }
    float* dst = cat + ((int64_t)b * d.L + i0 + r) * d.NCAT + d.c_pair;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= d.H) break;
      dst[h * d.C + c] = acc[h];
    }
  }
  __syncthreads();
  // inverse frame + norms (:327-336, :453-457): ol[c'] = sum_k (og[k] - t[k]) R[c'][k]
  for (int task = threadIdx.x; task < IBF * d.H * d.Pv; task += blockDim.x) {
    int r = task / (d.H * d.Pv), hp = task - r * d.H * d.Pv;
    if (r >= nrows) continue;
    int64_t row = (int64_t)b * d.L + i0 + r;
    const float* Rr = R + row * 9;
    const float* tr = t + row * 3;
    float gx = s_og[r * d.NV + hp * 3] - tr[0], gy = s_og[r * d.NV + hp * 3 + 1] - tr[1], gz = s_og[r * d.NV + hp * 3 + 2] - tr[2];
    float lx = gx * Rr[0] + gy * Rr[1] + gz * Rr[2];
    float ly = gx * Rr[3] + gy * Rr[4] + gz * Rr[5];
    float lz = gx * Rr[6] + gy * Rr[7] + gz * Rr[8];
    float* dst = cat + row * d.NCAT;
    dst[d.c_point + hp * 3] = lx; dst[d.c_point + hp * 3 + 1] = ly; dst[d.c_point + hp * 3 + 2] = lz;
    dst[d.c_norm + hp] = sqrtf(lx * lx + ly * ly + lz * lz);
  }
}

// ---- backward, query side -----------------------------------------------------------------
// Per (patch, IBB query rows): recompute attn; dattn from dcat; dlogit; writes
//   attnbuf[b,h,i,j], dsbuf[b,h,i,j] (= scale_total * dlogit), de[b,i,j,:], dproj q-part rows,
//   dog[b,i,NV] (global-frame grads of the aggregated points), atomics into dWpb and dgamma.
__global__ void __launch_bounds__(NT) ipa_attn_bwd_q_kernel(
    IpaD d, const float* __restrict__ proj, const float* __restrict__ e, const float* __restrict__ R,
    const float* __restrict__ t, const float* __restrict__ wpb, const float* __restrict__ gamma,
    const float* __restrict__ cat, const float* __restrict__ dcat, float* __restrict__ attnbuf,
    float* __restrict__ dsbuf, float* __restrict__ de, float* __restrict__ dproj, float* __restrict__ dog,
    float* __restrict__ g_wpb, float* __restrict__ g_gamma) {
  extern __shared__ __align__(16) float smem[];
  const int NQK = d.NS + d.NQ;
  const int HC = d.H * d.C;
  float* s_q = smem;                          // [IBB][NQK]
  float* s_wpb = s_q + IBB * NQK;             // [H][C]
  float* s_attn = s_wpb + HC;                 // [IBB][H][L]
  float* s_ds = s_attn + IBB * d.H * d.L;     // [IBB][H][L]  dattn, then dS
  float* s_dcat = s_ds + IBB * d.H * d.L;     // [IBB][NS + HC + NV]: d out_scalar | d out_pair | d og (global)
  float* s_red = s_dcat + IBB * (d.NS + HC + d.NV);  // [IBB*H] delta, then [H] gamma partials
  const int DC = d.NS + HC + d.NV;
  int b = blockIdx.y, i0 = blockIdx.x * IBB;
  int nrows = min(IBB, d.L - i0);
  for (int idx = threadIdx.x; idx < IBB * NQK; idx += blockDim.x) {
    int r = idx / NQK, c = idx - r * NQK;
    float v = 0.f;
    if (r < nrows) {
      const float* row = proj + ((int64_t)b * d.L + i0 + r) * d.NPROJ;
      v = (c < d.NS) ? row[d.o_qs + c] : row[d.o_qp + (c - d.NS)];
    }
    s_q[idx] = v;
  }
  for (int idx = threadIdx.x; idx < HC; idx += blockDim.x) s_wpb[idx] = wpb[idx];
  // upstream grads of scalar and pair features
  for (int idx = threadIdx.x; idx < IBB * (d.NS + HC); idx += blockDim.x) {
    int r = idx / (d.NS + HC), c = idx - r * (d.NS + HC);
    s_dcat[r * DC + c] = (r < nrows) ? dcat[((int64_t)b * d.L + i0 + r) * d.NCAT + c] : 0.f;
  }
  // upstream grads of local points and norms -> grads of global aggregated points
  // ol = (og - t) R^T ; nrm = |ol| ;  d ol_tot = d ol + d nrm * ol / nrm ;  d og[k] = sum_c' d ol_tot[c'] R[c'][k]
  for (int task = threadIdx.x; task < IBB * d.H * d.Pv; task += blockDim.x) {
    int r = task / (d.H * d.Pv), hp = task - r * d.H * d.Pv;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (r < nrows) {
      int64_t row = (int64_t)b * d.L + i0 + r;
      const float* crow = cat + row * d.NCAT;
      const float* drow = dcat + row * d.NCAT;
      const float* Rr = R + row * 9;
      float lx = crow[d.c_point + hp * 3], ly = crow[d.c_point + hp * 3 + 1], lz = crow[d.c_point + hp * 3 + 2];
      float nrm = crow[d.c_norm + hp];
      float dn = drow[d.c_norm + hp];
      float k = nrm > 0.f ? dn / nrm : 0.f;
      float dx = drow[d.c_point + hp * 3] + k * lx, dy = drow[d.c_point + hp * 3 + 1] + k * ly,
            dz = drow[d.c_point + hp * 3 + 2] + k * lz;
      gx = dx * Rr[0] + dy * Rr[3] + dz * Rr[6];
      gy = dx * Rr[1] + dy * Rr[4] + dz * Rr[7];
      gz = dx * Rr[2] + dy * Rr[5] + dz * Rr[8];
      float* o = dog + row * d.NV + hp * 3;
      o[0] = gx; o[1] = gy; o[2] = gz;
    }
    float* sd = s_dcat + r * DC + d.NS + HC + hp * 3;
    sd[0] = gx; sd[1] = gy; sd[2] = gz;
  }
  __syncthreads();
  // pass 1: logits -> s_attn
  for (int j = threadIdx.x; j < d.L; j += blockDim.x) {
    float qk[IBB][MAXH], d2[IBB][MAXH];
    logits_for_column<IBB>(d, proj, e, s_q, s_wpb, b, i0, nrows, j, qk, d2);
#pragma unroll
    for (int r = 0; r < IBB; ++r)
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h >= d.H) break;
        s_attn[(r * d.H + h) * d.L + j] = d.st * (qk[r][h] - 0.5f * d.sp * __ldg(gamma + h) * d2[r][h]);
      }
  }
  __syncthreads();
  softmax_rows(s_attn, IBB * d.H, d.L);
  __syncthreads();
  // pass 2: dattn[r][h][j] = <d os, vs_j> + <d op, e_ij> + <d og, vp_j>
  float gpart[MAXH];
#pragma unroll
  for (int h = 0; h < MAXH; ++h) gpart[h] = 0.f;
  for (int j = threadIdx.x; j < d.L; j += blockDim.x) {
    const float* vrow = proj + ((int64_t)b * d.L + j) * d.NPROJ;
    float da[IBB][MAXH];
#pragma unroll
    for (int r = 0; r < IBB; ++r)
#pragma unroll
      for (int h = 0; h < MAXH; ++h) da[r][h] = 0.f;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= d.H) break;
      for (int dd = 0; dd < d.ds; ++dd) {
        float v = __ldg(vrow + d.o_vs + h * d.ds + dd);
#pragma unroll
        for (int r = 0; r < IBB; ++r) da[r][h] = fmaf(s_dcat[r * DC + h * d.ds + dd], v, da[r][h]);
      }
      for (int pc = 0; pc < d.Pv * 3; ++pc) {
        float v = __ldg(vrow + d.o_vp + h * d.Pv * 3 + pc);
#pragma unroll
        for (int r = 0; r < IBB; ++r) da[r][h] = fmaf(s_dcat[r * DC + d.NS + HC + h * d.Pv * 3 + pc], v, da[r][h]);
      }
    }
#pragma unroll
    for (int r = 0; r < IBB; ++r) {
      if (r >= nrows) break;
      const float* er = e + (((int64_t)b * d.L + i0 + r) * d.L + j) * d.C;
      for (int c = 0; c < d.C; ++c) {
        float ev = __ldg(er + c);
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
          if (h >= d.H) break;
          da[r][h] = fmaf(s_dcat[r * DC + d.NS + h * d.C + c], ev, da[r][h]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < IBB; ++r)
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h >= d.H) break;
        s_ds[(r * d.H + h) * d.L + j] = da[r][h];
      }
  }
  __syncthreads();
  // delta[r][h] = sum_j attn * dattn  (one warp per row)
  {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int row = warp; row < IBB * d.H; row += nwarp) {
      float s = 0.f;
      for (int j = lane; j < d.L; j += 32) s += s_attn[row * d.L + j] * s_ds[row * d.L + j];
      s = warp_sum(s);
      if (lane == 0) s_red[row] = s;
    }
  }
  __syncthreads();
  // pass 3: dS = st * attn * (dattn - delta); write tiles, de row, gamma partials
  for (int j = threadIdx.x; j < d.L; j += blockDim.x) {
    // recompute d2 for the gamma gradient (cheap: 3*Pq*H flops per row)
    const float* krow = proj + ((int64_t)b * d.L + j) * d.NPROJ;
    float dsv[IBB][MAXH], av[IBB][MAXH];
#pragma unroll
    for (int r = 0; r < IBB; ++r)
#pragma unroll
      for (int h = 0; h < MAXH; ++h) {
        if (h >= d.H) { dsv[r][h] = 0.f; av[r][h] = 0.f; continue; }
        int o = (r * d.H + h) * d.L + j;
        float a = s_attn[o];
        float v = d.st * a * (s_ds[o] - s_red[r * d.H + h]);
        if (r >= nrows) { v = 0.f; a = 0.f; }
        dsv[r][h] = v; av[r][h] = a;
        s_ds[o] = v;
        if (r < nrows) {
          int64_t g = (((int64_t)b * d.H + h) * d.L + i0 + r) * d.L + j;
          attnbuf[g] = a;
          dsbuf[g] = v;
        }
      }
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= d.H) break;
      const float* kpp = krow + d.o_kp + h * d.Pq * 3;
      float d2r[IBB];
#pragma unroll
      for (int r = 0; r < IBB; ++r) d2r[r] = 0.f;
      for (int pc = 0; pc < d.Pq * 3; ++pc) {
        float kv = __ldg(kpp + pc);
#pragma unroll
        for (int r = 0; r < IBB; ++r) {
          float df = s_q[r * NQK + d.NS + h * d.Pq * 3 + pc] - kv;
          d2r[r] = fmaf(df, df, d2r[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < IBB; ++r) gpart[h] += dsv[r][h] * (-0.5f * d.sp) * d2r[r];
    }
    // de[i,j,c] = sum_h dS[h] Wpb[h,c] + attn[h] d_op[h,c]
#pragma unroll
    for (int r = 0; r < IBB; ++r) {
      if (r >= nrows) break;
      float* der = de + (((int64_t)b * d.L + i0 + r) * d.L + j) * d.C;
      for (int c = 0; c < d.C; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
          if (h >= d.H) break;
          acc = fmaf(dsv[r][h], s_wpb[h * d.C + c], acc);
          acc = fmaf(av[r][h], s_dcat[r * DC + d.NS + h * d.C + c], acc);
        }
        der[c] = acc;
      }
    }
  }
  __syncthreads();
  // gamma gradient: block reduce then one atomic per head
  {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      float v = warp_sum(gpart[h]);
      if (lane == 0) s_red[IBB * d.H + warp * MAXH + h] = v;
    }
    __syncthreads();
    if (threadIdx.x < d.H) {
      float v = 0.f;
      for (int w = 0; w < nwarp; ++w) v += s_red[IBB * d.H + w * MAXH + threadIdx.x];
      atomicAdd(g_gamma + threadIdx.x, v);
    }
  }
  // dq scalar: dqs[r][f] = ss * sum_j dS[r][h][j] ks[j][f]
  aggregate_rows<IBB>(d, proj, s_ds, b, d.o_ks, d.NS, d.ds, [&](int r, int f, float v) {
    if (r < nrows) dproj[((int64_t)b * d.L + i0 + r) * d.NPROJ + d.o_qs + f] = d.ss * v;
  });
  // dq point: dqp[r][h][pc] = -sp gamma_h sum_j dS (qp - kp_j) = -sp gamma_h (qp sum_j dS - sum_j dS kp_j)
  // sum_j dS over a softmax row is 0 analytically, but we keep the exact expression.
  for (int f = threadIdx.x; f < d.NQ; f += blockDim.x) {
    int h = f / (d.Pq * 3);
    float acc[IBB], tot[IBB];
#pragma unroll
    for (int r = 0; r < IBB; ++r) { acc[r] = 0.f; tot[r] = 0.f; }
    const float* kcol = proj + (int64_t)b * d.L * d.NPROJ + d.o_kp + f;
    for (int j = 0; j < d.L; ++j) {
      float kv = __ldg(kcol + (int64_t)j * d.NPROJ);
#pragma unroll
      for (int r = 0; r < IBB; ++r) {
        float w = s_ds[(r * d.H + h) * d.L + j];
        acc[r] = fmaf(w, kv, acc[r]);
        tot[r] += w;
      }
    }
    float gh = -d.sp * __ldg(gamma + h);
#pragma unroll
    for (int r = 0; r < IBB; ++r)
      if (r < nrows)
        dproj[((int64_t)b * d.L + i0 + r) * d.NPROJ + d.o_qp + f] = gh * (s_q[r * NQK + d.NS + f] * tot[r] - acc[r]);
  }
  // dWpb[h][c] += sum_{r,j} dS[r][h][j] e[i0+r][j][c]
  for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
    float acc[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) acc[h] = 0.f;
    for (int r = 0; r < nrows; ++r) {
      const float* ecol = e + ((int64_t)b * d.L + i0 + r) * d.L * d.C + c;
      for (int j = 0; j < d.L; ++j) {
        float ev = __ldg(ecol + (int64_t)j * d.C);
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
          if (h >= d.H) break;
          acc[h] = fmaf(s_ds[(r * d.H + h) * d.L + j], ev, acc[h]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h >= d.H) break;
      atomicAdd(g_wpb + h * d.C + c, acc[h]);
    }
  }
}

// ---- backward, key side -------------------------------------------------------------------
// Per (patch, JBB key columns): dvs, dvp, dks, dkp rows of dproj from the attention / dS tiles.
__global__ void __launch_bounds__(NT) ipa_attn_bwd_k_kernel(IpaD d, const float* __restrict__ proj,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ dcat,
                                                            const float* __restrict__ dog,
                                                            const float* __restrict__ attnbuf,
                                                            const float* __restrict__ dsbuf,
                                                            float* __restrict__ dproj) {
  extern __shared__ __align__(16) float smem[];
  float* s_a = smem;                        // [JBB][H][L]  attn[h][i][j0+c]
  float* s_g = s_a + JBB * d.H * d.L;       // [JBB][H][L]  dS
  int b = blockIdx.y, j0 = blockIdx.x * JBB;
  int ncols = min(JBB, d.L - j0);
  for (int idx = threadIdx.x; idx < d.H * d.L * JBB; idx += blockDim.x) {
    int c = idx % JBB;
    int hi = idx / JBB;  // h * L + i
    float a = 0.f, g = 0.f;
    if (c < ncols) {
      int64_t src = ((int64_t)b * d.H * d.L + hi) * d.L + j0 + c;
      a = __ldg(attnbuf + src);
      g = __ldg(dsbuf + src);
    }
    s_a[c * d.H * d.L + hi] = a;
    s_g[c * d.H * d.L + hi] = g;
  }
  __syncthreads();
  auto reduce_i = [&](const float* s_w, const float* __restrict__ src, int64_t src_stride, int f, int h,
                      float (&acc)[JBB]) {
#pragma unroll
    for (int c = 0; c < JBB; ++c) acc[c] = 0.f;
    for (int i = 0; i < d.L; ++i) {
      float v = __ldg(src + (int64_t)i * src_stride + f);
#pragma unroll
      for (int c = 0; c < JBB; ++c) acc[c] = fmaf(s_w[(c * d.H + h) * d.L + i], v, acc[c]);
    }
  };
  float acc[JBB];
  // dvs[j][f] = sum_i attn[h][i][j] d_os[i][f]
  for (int f = threadIdx.x; f < d.NS; f += blockDim.x) {
    int h = f / d.ds;
    reduce_i(s_a, dcat + (int64_t)b * d.L * d.NCAT, d.NCAT, f, h, acc);
#pragma unroll
    for (int c = 0; c < JBB; ++c)
      if (c < ncols) dproj[((int64_t)b * d.L + j0 + c) * d.NPROJ + d.o_vs + f] = acc[c];
  }
  // dvp[j][f] = sum_i attn[h][i][j] d_og[i][f]
  for (int f = threadIdx.x; f < d.NV; f += blockDim.x) {
    int h = f / (d.Pv * 3);
    reduce_i(s_a, dog + (int64_t)b * d.L * d.NV, d.NV, f, h, acc);
#pragma unroll
    for (int c = 0; c < JBB; ++c)
      if (c < ncols) dproj[((int64_t)b * d.L + j0 + c) * d.NPROJ + d.o_vp + f] = acc[c];
  }
  // dks[j][f] = ss sum_i dS[h][i][j] qs[i][f]
  for (int f = threadIdx.x; f < d.NS; f += blockDim.x) {
    int h = f / d.ds;
    reduce_i(s_g, proj + (int64_t)b * d.L * d.NPROJ + d.o_qs, d.NPROJ, f, h, acc);
#pragma unroll
    for (int c = 0; c < JBB; ++c)
      if (c < ncols) dproj[((int64_t)b * d.L + j0 + c) * d.NPROJ + d.o_ks + f] = d.ss * acc[c];
  }
  // dkp[j][f] = sp gamma_h (sum_i dS qp[i][f] - kp[j][f] sum_i dS)
  for (int f = threadIdx.x; f < d.NQ; f += blockDim.x) {
    int h = f / (d.Pq * 3);
    reduce_i(s_g, proj + (int64_t)b * d.L * d.NPROJ + d.o_qp, d.NPROJ, f, h, acc);
    float gh = d.sp * __ldg(gamma + h);
#pragma unroll
    for (int c = 0; c < JBB; ++c) {
      if (c >= ncols) break;
      float tot = 0.f;
      for (int i = 0; i < d.L; ++i) tot += s_g[(c * d.H + h) * d.L + i];
      int64_t row = (int64_t)b * d.L + j0 + c;
      dproj[row * d.NPROJ + d.o_kp + f] = gh * (acc[c] - __ldg(proj + row * d.NPROJ + d.o_kp + f) * tot);
    }
  }
}

// ------------------------------------------------------------------------------------------
static size_t fwd_smem_bytes(const IpaD& d) {
  return sizeof(float) * ((size_t)IBF * (d.NS + d.NQ) + d.H * d.C + (size_t)IBF * d.H * d.L + (size_t)IBF * d.NV);
}
static size_t bwdq_smem_bytes(const IpaD& d) {
  return sizeof(float) * ((size_t)IBB * (d.NS + d.NQ) + d.H * d.C + 2 * (size_t)IBB * d.H * d.L +
                          (size_t)IBB * (d.NS + d.H * d.C + d.NV) + IBB * d.H + 8 * MAXH + 16);
}
static size_t bwdk_smem_bytes(const IpaD& d) { return sizeof(float) * 2 * (size_t)JBB * d.H * d.L; }

static int validate(const DabIpaDims* dims, const char* name) {
  DAB_REQUIRE(dims, DAB_EINVAL, "%s: null dims", name);
  DAB_REQUIRE(dims->B >= 0 && dims->L > 0 && dims->D > 0 && dims->C > 0 && dims->H > 0 && dims->ds > 0 &&
                  dims->Pq > 0 && dims->Pv > 0,
              DAB_EINVAL, "%s: dimensions must be positive", name);
  DAB_REQUIRE(dims->H <= MAXH, DAB_EUNSUPPORTED, "%s: n_head > %d is not supported", name, MAXH);
  IpaD d = make_dims(dims);
  size_t need = fwd_smem_bytes(d);
  if (bwdq_smem_bytes(d) > need) need = bwdq_smem_bytes(d);
  if (bwdk_smem_bytes(d) > need) need = bwdk_smem_bytes(d);
  DAB_REQUIRE(need <= 227 * 1024, DAB_EUNSUPPORTED, "%s: shape needs %zu B of shared memory (> 227 KB)", name, need);
  return DAB_OK;
}

struct Workspace {
  float *proj, *cat, *dcat, *dproj, *attn, *ds, *dog;
  size_t bytes_fwd, bytes_bwd;
};

static Workspace carve(const IpaD& d, void* base) {
  auto al = [](size_t n) { return (n + 63) / 64 * 64; };  // 256-byte granules (in floats)
  size_t rows = (size_t)d.B * d.L;
  size_t n_proj = al(rows * d.NPROJ), n_cat = al(rows * d.NCAT);
  size_t n_tile = al((size_t)d.B * d.H * d.L * d.L), n_dog = al(rows * d.NV);
  Workspace w;
  float* p = reinterpret_cast<float*>(base);
  w.proj = p; p += n_proj;
  w.cat = p; p += n_cat;
  w.bytes_fwd = (size_t)(p - reinterpret_cast<float*>(base)) * sizeof(float);
  w.dcat = p; p += n_cat;
  w.dproj = p; p += n_proj;
  w.attn = p; p += n_tile;
  w.ds = p; p += n_tile;
  w.dog = p; p += n_dog;
  w.bytes_bwd = (size_t)(p - reinterpret_cast<float*>(base)) * sizeof(float);
  return w;
}

template <typename K>
static void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace dab

using namespace dab;

extern "C" {

size_t dab_ipa_f32_workspace_bytes(const DabIpaDims* dims, int for_backward) {
  if (!dims) return 0;
  IpaD d = make_dims(dims);
  Workspace w = carve(d, nullptr);
  return for_backward ? w.bytes_bwd : w.bytes_fwd;
}

int dab_ipa_fwd_f32(const DabIpaDims* dims, const DabIpaWeights* w, const float* x, const float* e, const float* R,
                    const float* t, float* y, void* workspace, size_t workspace_bytes, int save_for_bwd, void* stream) {
  (void)save_for_bwd;  // forward products always live in the workspace; the flag documents intent
  if (int rc = validate(dims, "dab_ipa_fwd_f32")) return rc;
  if (dims->B == 0) return DAB_OK;
  DAB_REQUIRE(w && x && e && R && t && y && workspace, DAB_EINVAL, "dab_ipa_fwd_f32: null pointer");
  DAB_REQUIRE(w->w_q_scalar && w->w_k_scalar && w->w_v_scalar && w->w_q_point && w->w_k_point && w->w_v_point &&
                  w->w_pair_bias && w->gamma && w->w_out && w->b_out,
              DAB_EINVAL, "dab_ipa_fwd_f32: null weight pointer");
  DAB_REQUIRE(aligned16(x) && aligned16(e) && aligned16(y) && aligned16(workspace), DAB_EINVAL,
              "dab_ipa_fwd_f32: pointers must be 16-byte aligned");
  IpaD d = make_dims(dims);
  Workspace ws = carve(d, workspace);
  DAB_REQUIRE(workspace_bytes >= ws.bytes_fwd, DAB_EWORKSPACE, "dab_ipa_fwd_f32: workspace %zu < %zu", workspace_bytes,
              ws.bytes_fwd);
  cudaStream_t s = (cudaStream_t)stream;
  int M = d.B * d.L;
  // projections (bias-free nn.Linear: y = x W^T), diffab_pytorch.py:391-403
  const float* Ws[6] = {w->w_q_scalar, w->w_k_scalar, w->w_v_scalar, w->w_q_point, w->w_k_point, w->w_v_point};
  int offs[6] = {d.o_qs, d.o_ks, d.o_vs, d.o_qp, d.o_kp, d.o_vp};
  int ns[6] = {d.NS, d.NS, d.NS, d.NQ, d.NQ, d.NV};
  const int phases = 7;   // all three stages (projections, attention core, to_out)
  if (phases & 1) {
    for (int k = 0; k < 6; ++k) sgemm(s, x, d.D, 1, Ws[k], 1, d.D, ws.proj + offs[k], d.NPROJ, nullptr, M, ns[k], d.D, 0);
    int64_t npts = (int64_t)M * ((2 * d.NQ + d.NV) / 3);
    frame_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(ws.proj, R, t, d, 0);
    count_launch();
  }
  if (phases & 2) {
    size_t smem = fwd_smem_bytes(d);
    allow_smem(ipa_attn_fwd_kernel, smem);
    dim3 grid((d.L + IBF - 1) / IBF, d.B);
    ipa_attn_fwd_kernel<<<grid, NT, smem, s>>>(d, ws.proj, e, R, t, w->w_pair_bias, w->gamma, ws.cat);
    count_launch();
  }
  // to_out, diffab_pytorch.py:464
  if (phases & 4) sgemm(s, ws.cat, d.NCAT, 1, w->w_out, 1, d.NCAT, y, d.D, w->b_out, M, d.D, d.NCAT, 0);
  return check_launch("dab_ipa_fwd_f32");
}

int dab_ipa_bwd_f32(const DabIpaDims* dims, const DabIpaWeights* w, const float* x, const float* e, const float* R,
                    const float* t, const float* dy, float* dx, float* de, const DabIpaGrads* g, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (int rc = validate(dims, "dab_ipa_bwd_f32")) return rc;
  if (dims->B == 0) return DAB_OK;
  DAB_REQUIRE(w && x && e && R && t && dy && dx && de && g && workspace, DAB_EINVAL, "dab_ipa_bwd_f32: null pointer");
  DAB_REQUIRE(g->w_q_scalar && g->w_k_scalar && g->w_v_scalar && g->w_q_point && g->w_k_point && g->w_v_point &&
                  g->w_pair_bias && g->gamma && g->w_out && g->b_out,
              DAB_EINVAL, "dab_ipa_bwd_f32: null gradient pointer");
  IpaD d = make_dims(dims);
  Workspace ws = carve(d, workspace);
  DAB_REQUIRE(workspace_bytes >= ws.bytes_bwd, DAB_EWORKSPACE, "dab_ipa_bwd_f32: workspace %zu < %zu", workspace_bytes,
              ws.bytes_bwd);
  cudaStream_t s = (cudaStream_t)stream;
  int M = d.B * d.L;
  int splitk = M >= 4096 ? 32 : (M >= 512 ? 8 : 1);
  // to_out backward: dcat = dy @ Wout ; dWout += dy^T @ cat ; dbout += colsum(dy)
  sgemm(s, dy, d.D, 1, w->w_out, d.NCAT, 1, ws.dcat, d.NCAT, nullptr, M, d.NCAT, d.D, 0);
  sgemm(s, dy, 1, d.D, ws.cat, d.NCAT, 1, g->w_out, d.NCAT, nullptr, d.D, d.NCAT, M, 1, splitk);
  {
    dim3 grid((d.D + 127) / 128, M >= 1024 ? 32 : 1);
    colsum_kernel<<<grid, 128, 0, s>>>(dy, M, d.D, g->b_out);
    count_launch();
  }
  // attention core
  size_t smq = bwdq_smem_bytes(d), smk = bwdk_smem_bytes(d);
  allow_smem(ipa_attn_bwd_q_kernel, smq);
  allow_smem(ipa_attn_bwd_k_kernel, smk);
  dim3 gq((d.L + IBB - 1) / IBB, d.B);
  ipa_attn_bwd_q_kernel<<<gq, NT, smq, s>>>(d, ws.proj, e, R, t, w->w_pair_bias, w->gamma, ws.cat, ws.dcat, ws.attn,
                                            ws.ds, de, ws.dproj, ws.dog, g->w_pair_bias, g->gamma);
  dim3 gk((d.L + JBB - 1) / JBB, d.B);
  ipa_attn_bwd_k_kernel<<<gk, NT, smk, s>>>(d, ws.proj, w->gamma, ws.dcat, ws.dog, ws.attn, ws.ds, ws.dproj);
  count_launch(3);  // q-side, k-side, inverse frame
  // frames: d_local = d_global @ R^T
  int64_t npts = (int64_t)M * ((2 * d.NQ + d.NV) / 3);
  frame_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(ws.dproj, R, t, d, 1);
  // projections backward: dx = sum_k dproj_k @ W_k ; dW_k += dproj_k^T @ x
  const float* Ws[6] = {w->w_q_scalar, w->w_k_scalar, w->w_v_scalar, w->w_q_point, w->w_k_point, w->w_v_point};
  float* Gs[6] = {g->w_q_scalar, g->w_k_scalar, g->w_v_scalar, g->w_q_point, g->w_k_point, g->w_v_point};
  int offs[6] = {d.o_qs, d.o_ks, d.o_vs, d.o_qp, d.o_kp, d.o_vp};
  int ns[6] = {d.NS, d.NS, d.NS, d.NQ, d.NQ, d.NV};
  for (int k = 0; k < 6; ++k) {
    sgemm(s, ws.dproj + offs[k], d.NPROJ, 1, Ws[k], d.D, 1, dx, d.D, nullptr, M, d.D, ns[k], k == 0 ? 0 : 1);
    sgemm(s, ws.dproj + offs[k], 1, d.NPROJ, x, d.D, 1, Gs[k], d.D, nullptr, ns[k], d.D, M, 1, splitk);
  }
  return check_launch("dab_ipa_bwd_f32");
}

}  // extern "C"
