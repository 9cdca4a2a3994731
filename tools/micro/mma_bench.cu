// Microbenchmark: cost of small tcgen05.mma instructions issued by one thread (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I diffab-pytorch_b200/csrc -o /tmp/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sm100_prims.cuh"
using namespace dab::sm100;

template <int M, int N, int AMN, int NACC = 1, int BMN = 0>
__global__ void __launch_bounds__(128) k_mma(long long* out, int n_mma, int smem_pad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncwarp();
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  for (int i = threadIdx.x; i < 8192; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tcgen05_fence_before_sync();
  __syncthreads();
  tcgen05_fence_after_sync();
  uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N, AMN, BMN);
    uint32_t a = smem_u32(smem), b = a + 16384;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      uint64_t da = AMN ? make_smem_desc(a + (i & 7) * 2048, 16384, 1024, kSwizzle128B)
                        : (smem_pad == 64 ? make_smem_desc(a + (i & 1) * 32, 16, 512, kSwizzle64B)
                                          : make_smem_desc(a + (i & 3) * 32, 16, 1024, kSwizzle128B));
      uint64_t db = BMN ? make_smem_desc(b + (i & 7) * 2048, 1024, 1024, kSwizzle128B)
                        : (smem_pad == 64 ? make_smem_desc(b + (i & 1) * 32, 16, 512, kSwizzle64B)
                                          : make_smem_desc(b + (i & 3) * 32, 16, 1024, kSwizzle128B));
      umma_bf16(tmem + (i % NACC) * 64, da, db, idesc, i >= NACC);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_free(tmem, 256);
}

static int g_pad = 0;
template <int M, int N, int AMN, int NACC = 1, int BMN = 0>
void run(const char* name, int grid, int smem_bytes, int n_mma) {
  long long* d;
  cudaMalloc(&d, grid * 16);
  cudaFuncSetAttribute(k_mma<M, N, AMN, NACC, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; ++rep) k_mma<M, N, AMN, NACC, BMN><<<grid, 128, smem_bytes>>>(d, n_mma, g_pad);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  printf("%-28s grid %4d smem %6d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", name, grid, smem_bytes,
         issue / grid / n_mma, total / grid / n_mma, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int n = 256;
  // one CTA per SM (big smem) vs two CTAs per SM
  run<128, 16, 0>("M128 N16  K-major", 148, 200 * 1024, n);
  run<128, 16, 0>("M128 N16  K-major", 296, 100 * 1024, n);
  run<128, 64, 0>("M128 N64  K-major", 148, 200 * 1024, n);
  run<128, 128, 0>("M128 N128 K-major", 148, 200 * 1024, n);
  run<128, 256, 0>("M128 N256 K-major", 148, 200 * 1024, n);
  run<64, 8, 1>("M64  N8   A MN-major", 148, 200 * 1024, n);
  run<64, 8, 1>("M64  N8   A MN-major", 296, 100 * 1024, n);
  run<64, 16, 1>("M64  N16  A MN-major", 148, 200 * 1024, n);
  run<64, 64, 1>("M64  N64  A MN-major", 148, 200 * 1024, n);
  run<64, 8, 0>("M64  N8   K-major", 148, 200 * 1024, n);
  run<128, 16, 0>("M128 N16  K-major 1 CTA", 1, 200 * 1024, n);
  run<128, 16, 0, 2>("M128 N16 K-major 2 acc", 148, 200 * 1024, n);
  run<128, 16, 0, 4>("M128 N16 K-major 4 acc", 148, 200 * 1024, n);
  run<128, 16, 0, 4>("M128 N16 K-major 4 acc", 296, 100 * 1024, n);
  run<64, 8, 1, 2>("M64 N8 A-MN 2 acc", 148, 200 * 1024, n);
  run<64, 8, 1, 4>("M64 N8 A-MN 4 acc", 148, 200 * 1024, n);
  run<64, 8, 1, 4>("M64 N8 A-MN 4 acc", 296, 100 * 1024, n);
  run<64, 64, 0, 1, 1>("M64 N64 A-K B-MN", 148, 200 * 1024, n);
  run<64, 64, 0, 4, 1>("M64 N64 A-K B-MN 4 acc", 148, 200 * 1024, n);
  run<128, 64, 0, 1, 1>("M128 N64 A-K B-MN", 148, 200 * 1024, n);
  run<128, 64, 0, 2, 1>("M128 N64 A-K B-MN 2acc", 296, 100 * 1024, n);
  printf("---- shapes of the attention core\n");
  run<128, 16, 0>("M128 N16 K-major SW128", 296, 100 * 1024, n);
  g_pad = 64;
  run<128, 16, 0>("M128 N16 K-major SW64", 296, 100 * 1024, n);
  run<128, 16, 0>("M128 N16 K-major SW64", 148, 200 * 1024, n);
  g_pad = 0;
  run<128, 16, 1>("M128 N16 A-MN (pair x2)", 296, 100 * 1024, n);
  run<128, 32, 1>("M128 N32 A-MN (O^T x2)", 296, 100 * 1024, n);
  run<128, 32, 0>("M128 N32 K-major", 296, 100 * 1024, n);
  run<128, 64, 0>("M128 N64 K-major", 296, 100 * 1024, n);
  run<64, 16, 0>("M64 N16 K-major", 296, 100 * 1024, n);
  return 0;
}
