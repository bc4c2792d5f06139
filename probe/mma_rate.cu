// mma_rate.cu — measures tcgen05.mma kind::i8 issue/execute rate from shared-memory operands (SS mode):
// cycles per MMA (M=128, K=32) for N in {64,128,256}, SW128 K-major operands, 1 CTA per SM on all SMs.
// Also kind::f8f6f4 for comparison.  Build: see probe/Makefile.  Run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dlq_b200/csrc/sm100_ptx.cuh"
using namespace dlq;

__global__ void __launch_bounds__(128, 1) mma_rate(int n, int iters, int kind, int distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    // whole warp walks the loop (uniform datapath), one elected lane issues; 4 MMAs unrolled per iteration
    const bool leader = elect_one();
    const uint32_t idesc = kind == 0 ? umma_idesc_s8(128, n) : umma_idesc_e4m3(128, n);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_SWZ_128B);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B);
    const int nacc = distinct ? 2 : 1;    // distinct=1: alternate two accumulators (independent chains)
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = tmem + ((nacc == 2 && (j & 1)) ? 256u : 0u);
          if (kind == 0) umma_i8(d, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
          else umma_f8(d, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 8192;
  for (int kind = 0; kind < 2; ++kind)
    for (int distinct = 0; distinct < 2; ++distinct)
      for (int n : {16, 32, 64, 96, 128, 160, 192, 224, 256}) {
        for (int grid : {148}) {
          mma_rate<<<grid, 128, 200 * 1024>>>(n, iters, kind, distinct, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2];
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          const double cyc = (double)h[1] / iters;
          printf("kind=%s distinct=%d N=%3d grid=%3d : issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f MAC/cyc/SM (ideal %d cyc)\n",
                 kind ? "f8" : "i8", distinct, n, grid, (double)h[0] / iters, cyc, 128.0 * n * 32 / cyc, n / 2);
        }
      }
  return 0;
}
