// SO(3) maps and IGSO(3) table/sampler for sm_100a.
//
// Replaces the arithmetic of /root/reference/diffab_pytorch/so3.py (file:line cited per kernel).
// All of these are HBM-bound elementwise kernels (12-76 B per rotation): one thread per rotation,
// global traffic staged through shared memory so every warp-level access is a contiguous
// 128-bit-vectorised stream, grid sized in whole waves of the 148 SMs by a grid-stride loop.
#include <math_constants.h>

#include "common.cuh"

namespace dab {

constexpr int kTile = 256;  // rotations per block iteration (= threads per block)

// Cooperative contiguous copy global->shared / shared->global of `count` floats starting at a
// 16-byte aligned global offset (tiles of 256 rotations are 3072 / 9216 B, so always aligned).
__device__ __forceinline__ void tile_load(const float* __restrict__ g, float* s, int count) {
  int nvec = count >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* s4 = reinterpret_cast<float4*>(s);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) s4[i] = __ldg(g4 + i);
  for (int i = (nvec << 2) + threadIdx.x; i < count; i += blockDim.x) s[i] = __ldg(g + i);
}
__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* s, int count) {
  int nvec = count >> 2;
  float4* g4 = reinterpret_cast<float4*>(g);
  const float4* s4 = reinterpret_cast<const float4*>(s);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) g4[i] = s4[i];
  for (int i = (nvec << 2) + threadIdx.x; i < count; i += blockDim.x) g[i] = s[i];
}

enum MapKind { kExpVec = 0, kLogVec = 1, kLogSkew = 2, kExpSkew = 3, kScaleRot = 4 };

template <int KIND>
__global__ void __launch_bounds__(kTile) so3_map_kernel(const float* __restrict__ in, const float* __restrict__ kscale,
                                                        float* __restrict__ out, int64_t n, int64_t group) {
  constexpr int IN_W = (KIND == kExpVec) ? 3 : 9;
  constexpr int OUT_W = (KIND == kLogVec) ? 3 : 9;
  __shared__ __align__(16) float s_in[kTile * IN_W];
  __shared__ __align__(16) float s_out[kTile * OUT_W];
  int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t base = tile * kTile;
    int cnt = (int)min((int64_t)kTile, n - base);
    tile_load(in + base * IN_W, s_in, cnt * IN_W);
    __syncthreads();
    if (threadIdx.x < cnt) {
      // stride-3 / stride-9 shared accesses are conflict-free (3 and 9 are odd)
      const float* a = s_in + threadIdx.x * IN_W;
      float* o = s_out + threadIdx.x * OUT_W;
      if (KIND == kExpVec) {
        so3_exp(a[0], a[1], a[2], o);
      } else if (KIND == kLogVec) {
        so3_log(a, o[0], o[1], o[2]);
      } else if (KIND == kLogSkew) {
        float x, y, z;
        so3_log(a, x, y, z);
        o[0] = 0.f; o[1] = -z; o[2] = y; o[3] = z; o[4] = 0.f; o[5] = -x; o[6] = -y; o[7] = x; o[8] = 0.f;
      } else if (KIND == kExpSkew) {
        so3_exp(a[7], a[2], a[3], o);  // vee(S) = (S21, S02, S10), so3.py:165-170
      } else {  // scale_rot, so3.py:240-259
        float x, y, z;
        so3_log(a, x, y, z);
        float k = __ldg(kscale + (base + threadIdx.x) / group);
        so3_exp(k * x, k * y, k * z, o);
      }
    }
    __syncthreads();
    tile_store(out + base * OUT_W, s_out, cnt * OUT_W);
    __syncthreads();
  }
}

static int grid_for(int64_t n_tiles) {
  int64_t cap = 148 * 8;  // 8 resident 256-thread blocks per SM
  return (int)(n_tiles < cap ? (n_tiles > 0 ? n_tiles : 1) : cap);
}

template <int KIND>
static int launch_map(const float* in, const float* k, float* out, int64_t n, int64_t group, void* stream,
                      const char* name) {
  DAB_REQUIRE(n >= 0, DAB_EINVAL, "%s: negative n", name);
  if (n == 0) return DAB_OK;
  DAB_REQUIRE(in && out, DAB_EINVAL, "%s: null pointer", name);
  DAB_REQUIRE(aligned16(in) && aligned16(out), DAB_EINVAL, "%s: pointers must be 16-byte aligned", name);
  int64_t n_tiles = (n + kTile - 1) / kTile;
  so3_map_kernel<KIND><<<grid_for(n_tiles), kTile, 0, (cudaStream_t)stream>>>(in, k, out, n, group);
  count_launch();
  return check_launch(name);
}

// ------------------------------------------------------------------------------------------
// IGSO(3) angular pdf table, so3.py:52-72.
//   out[s][k] = clamp0(nan_to_num( sum_{l<n_terms} fl(fl(a_k * b_l) * fl(sin(fl((l+.5) th_k)) / sin(th_k/2))) ))
//   a_k = (1 - cos th_k)/pi, b_l = (2l+1) exp(fl(-l(l+1)) * fl(sigma^2)), th_k = fl(k*binsize) + fl(binsize/2)
// Compute-bound (n_sigma*n_bins*n_terms sinf + divide).  One block per (sigma, 256-bin tile); b_l is
// computed once per block into shared memory; each thread owns one bin and runs the l-series
// sequentially in the reference's fp32 operation order (no FMA contraction across the product).
__global__ void __launch_bounds__(256) igso3_table_kernel(const float* __restrict__ sigma, int n_bins, int n_terms,
                                                          double binsize, float* __restrict__ out) {
  extern __shared__ float s_b[];  // n_terms
  int s = blockIdx.y;
  float sg = __ldg(sigma + s);
  float sg2 = __fmul_rn(sg, sg);
  for (int l = threadIdx.x; l < n_terms; l += blockDim.x) {
    float ll = -(float)((long long)l * (l + 1));
    s_b[l] = __fmul_rn((float)(2 * l + 1), expf(__fmul_rn(ll, sg2)));
  }
  __syncthreads();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_bins) return;
  float th = __fadd_rn((float)((double)k * binsize), (float)(binsize * 0.5));
  float a = __fdiv_rn(__fsub_rn(1.0f, cosf(th)), (float)CUDART_PI);
  float d = sinf(th * 0.5f);
  float acc = 0.f;
  for (int l = 0; l < n_terms; ++l) {
    float c = __fdiv_rn(sinf(__fmul_rn((float)l + 0.5f, th)), d);
    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(a, s_b[l]), c));
  }
  // nan_to_num (nan -> 0, +-inf -> +-FLT_MAX) then clamp_min(0)
  if (isnan(acc)) acc = 0.f;
  if (isinf(acc)) acc = acc > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  out[(int64_t)s * n_bins + k] = fmaxf(acc, 0.f);
}

// ------------------------------------------------------------------------------------------
// IGSO(3) sampler, so3.py:74-126, noise injected.  One block per batch row.
//  histogram branch: torch.multinomial(p, L) without replacement == indices of the L largest
//  p/q (q ~ Exp(1)) in descending order -> full bitonic sort of (key, index) pairs in shared memory
//  (n_bins <= 8192 -> 64 KB), ties broken by lower index.
//  gaussian branch: (2 sigma + sigma * g) mod pi  (torch.remainder semantics).
constexpr int kSortThreads = 1024;

// The draw of one batch row b by the whole block; rv[j * 3 + c] receives residue j's rotation vector (global or shared).
__device__ __forceinline__ void igso3_block(
    const float* __restrict__ hist, const float* __restrict__ sigmas, int n_bins, int n_sigma, int n_pow2,
    const int64_t* __restrict__ sigma_idx, int b, int L, const float* __restrict__ axis_noise,
    const float* __restrict__ exp_noise, const float* __restrict__ jitter, const float* __restrict__ gauss,
    float thr, double binsize, float* rv, int64_t* __restrict__ bins, unsigned char* smem_raw) {
  float* s_key = reinterpret_cast<float*>(smem_raw);
  int* s_idx = reinterpret_cast<int*>(smem_raw + sizeof(float) * n_pow2);
  int64_t row = sigma_idx[b];
  if (row < 0 || row >= n_sigma) asm volatile("trap;");   // the reference raises IndexError; here the launch fails loudly
  float sg = __ldg(sigmas + row);
  bool use_hist = sg < thr;
  bool need_sort = use_hist || bins != nullptr;
  if (need_sort) {
    const float* p = hist + row * (int64_t)n_bins;
    const float* q = exp_noise + (int64_t)b * n_bins;
    for (int k = threadIdx.x; k < n_pow2; k += blockDim.x) {
      s_key[k] = (k < n_bins) ? __fdiv_rn(__ldg(p + k), __ldg(q + k)) : -CUDART_INF_F;
      s_idx[k] = k;
    }
    __syncthreads();
    // bitonic sort, descending by key then ascending by index
    for (int size = 2; size <= n_pow2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
          int lo = 2 * i - (i & (stride - 1));
          int hi = lo + stride;
          bool desc = ((lo & size) == 0);
          float ka = s_key[lo], kb = s_key[hi];
          int ia = s_idx[lo], ib = s_idx[hi];
          bool a_first = (ka > kb) || (ka == kb && ia < ib);  // a should precede b in final order
          if (a_first != desc) {
            s_key[lo] = kb; s_key[hi] = ka; s_idx[lo] = ib; s_idx[hi] = ia;
          }
        }
        __syncthreads();
      }
    }
  }
  float binsize_f = (float)binsize;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    int64_t o = (int64_t)b * L + j;
    float ax = __ldg(axis_noise + o * 3), ay = __ldg(axis_noise + o * 3 + 1), az = __ldg(axis_noise + o * 3 + 2);
    float nrm = fmaxf(sqrtf(ax * ax + ay * ay + az * az), 1e-12f);  // F.normalize eps, so3.py:114
    float theta;
    if (need_sort) {
      int bin = s_idx[j];
      if (bins) bins[o] = bin;
      if (use_hist) theta = __fadd_rn((float)((double)bin * binsize), __fmul_rn(binsize_f, __ldg(jitter + o)));
    }
    if (!use_hist) {
      float v = __fadd_rn(__fmul_rn(sg, 2.0f), __fmul_rn(sg, __ldg(gauss + o)));
      float m = fmodf(v, (float)CUDART_PI);
      if (m != 0.f && m < 0.f) m += (float)CUDART_PI;  // torch.remainder: result takes the divisor's sign
      theta = m;
    }
    rv[j * 3] = (ax / nrm) * theta;
    rv[j * 3 + 1] = (ay / nrm) * theta;
    rv[j * 3 + 2] = (az / nrm) * theta;
  }
}

__global__ void __launch_bounds__(kSortThreads) igso3_sample_kernel(
    const float* __restrict__ hist, const float* __restrict__ sigmas, int n_bins, int n_sigma, int n_pow2,
    const int64_t* __restrict__ sigma_idx, int L, const float* __restrict__ axis_noise,
    const float* __restrict__ exp_noise, const float* __restrict__ jitter, const float* __restrict__ gauss,
    float thr, double binsize, float* __restrict__ rotvec, int64_t* __restrict__ bins) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  igso3_block(hist, sigmas, n_bins, n_sigma, n_pow2, sigma_idx, b, L, axis_noise, exp_noise, jitter, gauss, thr, binsize,
              rotvec + (int64_t)b * L * 3, bins, smem_raw);
}

}  // namespace dab

using namespace dab;

extern "C" {

int dab_so3_exp(const float* v, float* R, int64_t n, void* stream) {
  return launch_map<kExpVec>(v, nullptr, R, n, 1, stream, "dab_so3_exp");
}
int dab_so3_log(const float* R, float* v, int64_t n, void* stream) {
  return launch_map<kLogVec>(R, nullptr, v, n, 1, stream, "dab_so3_log");
}
int dab_so3_log_skew(const float* R, float* S, int64_t n, void* stream) {
  return launch_map<kLogSkew>(R, nullptr, S, n, 1, stream, "dab_so3_log_skew");
}
int dab_so3_exp_skew(const float* S, float* R, int64_t n, void* stream) {
  return launch_map<kExpSkew>(S, nullptr, R, n, 1, stream, "dab_so3_exp_skew");
}
int dab_so3_scale_rot(const float* R, const float* k, int64_t n, int64_t group, float* out, void* stream) {
  DAB_REQUIRE(k != nullptr && group > 0, DAB_EINVAL, "dab_so3_scale_rot: k null or group <= 0");
  return launch_map<kScaleRot>(R, k, out, n, group, stream, "dab_so3_scale_rot");
}

int dab_igso3_table(const float* sigma, int n_sigma, int n_bins, int n_terms, float* out, void* stream) {
  DAB_REQUIRE(sigma && out, DAB_EINVAL, "dab_igso3_table: null pointer");
  DAB_REQUIRE(n_sigma > 0 && n_bins > 0 && n_terms > 0, DAB_EINVAL, "dab_igso3_table: sizes must be positive");
  DAB_REQUIRE(n_terms <= 8192, DAB_EUNSUPPORTED, "dab_igso3_table: n_terms > 8192");
  dim3 grid((n_bins + 255) / 256, n_sigma);
  igso3_table_kernel<<<grid, 256, n_terms * sizeof(float), (cudaStream_t)stream>>>(
      sigma, n_bins, n_terms, 3.14159265358979323846 / (double)n_bins, out);
  count_launch();
  return check_launch("dab_igso3_table");
}

int dab_igso3_sample(const float* hist, const float* sigmas, int n_sigma, int n_bins, const int64_t* sigma_idx,
                     int B, int L, const float* axis_noise, const float* exp_noise, const float* jitter,
                     const float* gauss, float sigma_threshold, float* rotvec, int64_t* bins, void* stream) {
  DAB_REQUIRE(hist && sigmas && sigma_idx && axis_noise && exp_noise && jitter && gauss && rotvec, DAB_EINVAL,
              "dab_igso3_sample: null pointer");
  DAB_REQUIRE(B >= 0 && L >= 0 && n_bins > 0 && n_sigma > 0, DAB_EINVAL, "dab_igso3_sample: bad sizes");
  if (B == 0 || L == 0) return DAB_OK;
  DAB_REQUIRE(L <= n_bins, DAB_EINVAL, "dab_igso3_sample: cannot draw %d bins without replacement from %d", L, n_bins);
  int n_pow2 = 1;
  while (n_pow2 < n_bins) n_pow2 <<= 1;
  DAB_REQUIRE(n_pow2 <= 16384, DAB_EUNSUPPORTED, "dab_igso3_sample: n_bins > 16384 does not fit shared memory");
  size_t smem = (size_t)n_pow2 * 8;
  DAB_ENSURE_SMEM(igso3_sample_kernel, 16384 * 8);      // per device (the attribute is a per-device property)
  igso3_sample_kernel<<<B, kSortThreads, smem, (cudaStream_t)stream>>>(
      hist, sigmas, n_bins, n_sigma, n_pow2, sigma_idx, L, axis_noise, exp_noise, jitter, gauss, sigma_threshold,
      3.14159265358979323846 / (double)n_bins, rotvec, bins);
  count_launch();
  return check_launch("dab_igso3_sample");
}


}  // extern "C"
