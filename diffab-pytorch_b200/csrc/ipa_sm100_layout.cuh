// Shared layout of the sm_100a IPA path (forward ipa_sm100.cu, backward ipa_bwd_sm100.cu): the fixed
// train.py configuration, the packed-weight blob, the per-call workspace and small packing helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100_prims.cuh"

namespace dab {
namespace sm100 {

// fixed configuration of the fast path
constexpr int L = 128, D = 128, C = 64, H = 8, DS = 32, P = 8;
constexpr int NS = H * DS;            // 256
constexpr int NPT = H * P * 3;        // 192
constexpr int NPROJ = 3 * NS + 3 * NPT;  // 1344
constexpr int NCAT = NS + H * C + NPT + H * P;  // 1024
constexpr int QK_W = 96;              // packed q/k row per head: [scalar 32 | point hi 24 + 3 + pad 5 | point lo 24 + pad 8]
constexpr int V_W = 64;               // packed v row per head:   [scalar 32 | point 24 | pad 8]
constexpr int IB = 16;                // query rows per CTA
constexpr float kLog2e = 1.4426950408889634f;

struct PackedOffsets {
  size_t wcat, wout, wpb, bout, gamma, wcat_t, wout_t, total;
};
__host__ __device__ inline PackedOffsets packed_offsets() {
  PackedOffsets o;
  o.wcat = 0;
  o.wout = o.wcat + (size_t)NPROJ * D * 2;          // 344,064
  o.wpb = o.wout + (size_t)D * NCAT * 2;            // +262,144
  o.bout = o.wpb + 2048;
  o.gamma = o.bout + 512;
  // transposed copies for the backward data-gradient GEMMs (dx = dproj Wcat, dcat = dy Wout)
  o.wcat_t = (o.gamma + 64 + 1023) / 1024 * 1024;
  o.wout_t = o.wcat_t + (size_t)D * NPROJ * 2;      // [D][NPROJ] bf16
  o.total = o.wout_t + (size_t)NCAT * D * 2;        // [NCAT][D] bf16
  return o;
}

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 p = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- workspace --------------------------------------------------------------------------------------
struct Ws {
  __nv_bfloat16 *Qp, *Kp, *Vp, *cat;
  float* tc;
  uint4* bias;   // fallback plane when the caller did not precompute the layer's pair bias
  float* stats;  // [rows][16]: row maxima (log2 units) [8 heads] | 1 / sum_j p [8 heads]; written for the backward
  uint4* pu;     // [rows][128 j][8 h] bf16: un-normalised probabilities 2^(l - max), written for the backward
  size_t bytes;
};
inline Ws carve_ws(int B, void* base) {
  auto al = [](size_t n) { return (n + 1023) / 1024 * 1024; };
  size_t rows = (size_t)B * L;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  Ws w;
  w.Qp = reinterpret_cast<__nv_bfloat16*>(p); p += al(rows * H * QK_W * 2);
  w.Kp = reinterpret_cast<__nv_bfloat16*>(p); p += al(rows * H * QK_W * 2);
  w.Vp = reinterpret_cast<__nv_bfloat16*>(p); p += al(rows * H * V_W * 2);
  w.tc = reinterpret_cast<float*>(p); p += al(rows * 3 * 4);
  w.cat = reinterpret_cast<__nv_bfloat16*>(p); p += al(rows * NCAT * 2);
  w.bias = reinterpret_cast<uint4*>(p); p += al(rows * L * 16);
  w.stats = reinterpret_cast<float*>(p); p += al(rows * 16 * 4);
  w.pu = reinterpret_cast<uint4*>(p); p += al(rows * L * 16);
  w.bytes = (size_t)(p - reinterpret_cast<uint8_t*>(base));
  return w;
}

inline bool shape_ok(const DabIpaDims* d) {
  return d && d->L == L && d->D == D && d->C == C && d->H == H && d->ds == DS && d->Pq == P && d->Pv == P && d->B >= 0;
}
// inference forward: patches of 256 residues too (two blocks of 128: block-wise projections on a common centroid, the
// attention core once per (query block, key block) pair, the two key blocks' results merged by their softmax statistics)
inline bool shape_ok_fwd(const DabIpaDims* d) {
  return d && (d->L == L || d->L == 2 * L) && d->D == D && d->C == C && d->H == H && d->ds == DS && d->Pq == P && d->Pv == P &&
         d->B >= 0;
}
// extra workspace sections of a 256-residue forward, behind the sections of carve_ws(2 B blocks)
struct Ws2 {
  __nv_bfloat16* cat2;   // [2 key blocks][rows][NCAT] bf16: per-key-block concat features (normalised inside the block)
  float* stats2;         // [2][rows][16]: their softmax statistics
  float* cen;            // [blocks][3]: the patch centroid, repeated for the patch's two blocks
  uint4* bias2;          // fallback bias plane [rows][256] (caller without precomputed planes)
  size_t bytes;          // total, including the base sections
};
inline Ws2 carve_ws2(int B, void* base) {
  auto al = [](size_t n) { return (n + 1023) / 1024 * 1024; };
  const size_t rows = (size_t)B * 2 * L;
  uint8_t* p = reinterpret_cast<uint8_t*>(base) + carve_ws(2 * B, base).bytes;
  Ws2 w;
  w.cat2 = reinterpret_cast<__nv_bfloat16*>(p); p += al(2 * rows * NCAT * 2);
  w.stats2 = reinterpret_cast<float*>(p); p += al(2 * rows * 16 * 4);
  w.cen = reinterpret_cast<float*>(p); p += al((size_t)2 * B * 3 * 4);
  w.bias2 = reinterpret_cast<uint4*>(p); p += al(rows * 2 * L * 16);
  w.bytes = (size_t)(p - reinterpret_cast<uint8_t*>(base));
  return w;
}


}  // namespace sm100
}  // namespace dab
