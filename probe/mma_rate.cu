// mma_rate.cu — measures tcgen05.mma kind::i8 issue/execute rate from shared-memory operands (SS mode):
// cycles per MMA (M=128, K=32) for N in {64,128,256}, SW128 K-major operands, 1 CTA per SM on all SMs.
// Also kind::f8f6f4 for comparison.  Build: see probe/Makefile.  Run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../dlq_b200/csrc/sm100_ptx.cuh"
using namespace dlq;

__global__ void __launch_bounds__(128, 1) mma_rate(int n, int iters, int kind, int distinct, long long* out, int layout) {
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
    // layout 0: SW128 rows; 1: SW32 (32-byte rows); 2: no swizzle, 16-byte pixels, K chunk 1 = next pixel (LBO 16 B)
    const uint64_t ad0 = layout == 0   ? umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_SWZ_128B)
                         : layout == 1 ? umma_smem_desc(smem_u32(smem), 0, 256, UMMA_SWZ_32B)
                                       : umma_smem_desc(smem_u32(smem), 16, 128, UMMA_SWZ_NONE);
    const uint64_t bd0 = layout == 0   ? umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B)
                         : layout == 1 ? umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 256, UMMA_SWZ_32B)
                                       : umma_smem_desc(smem_u32(smem + 64 * 1024), n * 16, 128, UMMA_SWZ_NONE);
    const uint32_t kadv = layout == 0 ? 2u : 0u;
    const int nacc = distinct ? 2 : 1;    // distinct=1: alternate two accumulators (independent chains)
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = tmem + ((nacc == 2 && (j & 1)) ? 256u : 0u);
          if (kind == 0) umma_i8(d, ad0 + kadv * j, bd0 + kadv * j, idesc, 1u);
          else umma_f8(d, ad0 + kadv * j, bd0 + kadv * j, idesc, 1u);
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

// Pattern test: groups of G MMAs (N=64) separated by the per-step synchronisation the conv kernel performs.
//   mode bit0: tcgen05.commit to a scratch barrier after each group
//   mode bit1: mbarrier try_wait on an already-completed barrier before each group
//   mode bit2: tcgen05.fence::after_thread_sync before each group
//   mode bit3: __syncwarp after each group
__global__ void __launch_bounds__(128, 1) mma_pattern(int n, int groups, int G, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar, done_bar, scratch;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done_bar, 1); mbar_init(&scratch, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&done_bar);      // phase 0 of done_bar completes: waits on parity 0 succeed at once
  __syncthreads();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_s8(128, n);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_SWZ_128B);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode & 2) mbar_wait(&done_bar, 0);
      if (mode & 4) tc_fence_after();
      if (leader) {
        for (int j = 0; j < G; ++j) umma_i8(tmem + (j & 3) * 64, ad0 + 2u * (j & 3), bd0 + 2u * (j & 3), idesc, 1u);
        if (mode & 1) umma_commit(&scratch);
      }
      if (mode & 8) __syncwarp();
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && leader) { out[0] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// Multi-issuer test: W warps, each with its own elected lane issuing MMAs into its own accumulator.
// If the ~42-cycle fixed cost per tcgen05.mma is issue-side latency, several issuers overlap it.
__global__ void __launch_bounds__(128, 1) mma_multi(int n, int iters, int W, long long* out, int var, int kind) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp < W) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_s8(128, n);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem) + warp * 16384u, 0, 1024, UMMA_SWZ_128B);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B);
    long long tl0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = tmem + warp * 128 + (((var & 8) && (j & 1)) ? 256u : 0u);
          if (var & 2) {
            if (kind == 0) umma_i8(d, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
            else umma_f8(d, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
          } else {
            umma_i8(d, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (leader) umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    long long tl2 = clock64();
    if ((var & 1) && blockIdx.x == 0 && leader && warp == 0) out[1] = tl2 - tl0;
  }
  __syncthreads();
  long long t2 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t2 - t0;
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// TMEM read-out test: 4 warps (one per lane quarter) each issue `iters` tcgen05.ld.32x32b.x32 (4 KB per warp
// instruction); optionally one more warp issues MMAs (N=64) at the same time into other TMEM columns.
__global__ void __launch_bounds__(160, 1) tmem_ld_rate(int iters, int with_mma, int mma_iters, long long* out) {
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
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  uint32_t sink = 0;
  if (warp < 4) {
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + ((i & 3) * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) sink ^= v[j];
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  } else if (with_mma) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_s8(128, 64);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_SWZ_128B);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B);
    for (int i = 0; i < mma_iters; i += 4) {
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_i8(tmem + 256, ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
      }
      __syncwarp();
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && leader) out[1] = t1 - t0;
  }
  if (sink == 0x12345678u) out[2] = sink;
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// CTA-pair test: cluster of two CTAs, the rank-0 CTA issues tcgen05.mma.cta_group::2 (M = 256, N = n) in a tight loop.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_pair_rate(int n, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc_pair(&slot, 512); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int rank = cluster_ctarank();
  if (threadIdx.x < 32 && rank == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_s8(256, n);
    const uint64_t ad0 = umma_smem_desc(smem_u32(smem), 0, 1024, UMMA_SWZ_128B);
    const uint64_t bd0 = umma_smem_desc(smem_u32(smem + 64 * 1024), 0, 1024, UMMA_SWZ_128B);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_i8_pair(tmem + ((j & 1) ? 256u : 0u), ad0 + 2u * j, bd0 + 2u * j, idesc, 1u);
      }
      __syncwarp();
    }
    if (leader) umma_commit_pair(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && leader) out[0] = t2 - t0;
  } else if (threadIdx.x < 32) {
    mbar_wait(&bar, 0);     // the multicast commit also lands here
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 8192;
  cudaFuncSetAttribute(mma_pair_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n : {64, 128, 256}) {
    cudaMemset(d, 0, 32);
    mma_pair_rate<<<148, 128, 200 * 1024>>>(n, 4096, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair error %s\n", cudaGetErrorString(e)); return 1; }
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("pair (cta_group::2, M=256) N=%3d: %.1f cyc/MMA\n", n, (double)h / 4096);
  }
  cudaFuncSetAttribute(tmem_ld_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int with_mma : {0, 1}) {
    cudaMemset(d, 0, 32);
    tmem_ld_rate<<<148, 160, 200 * 1024>>>(2048, with_mma, 4096, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("tmem_ld 32x32b.x32 x 4 warps: %.1f cyc per 16 KB (all 4 warps) -> %.1f B/cyc/SM; concurrent MMA(N=64): %s %.1f cyc/MMA\n", (double)h[0] / 2048, 16384.0 * 2048 / h[0], with_mma ? "yes" : "no", with_mma ? (double)h[1] / 4096 : 0.0);
  }
  cudaFuncSetAttribute(mma_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n : {64, 128})
    for (int W : {1, 2, 4}) {
      mma_multi<<<148, 128, 200 * 1024>>>(n, 4096, W, d, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("multi-issuer N=%d W=%d: %.1f cyc per MMA (all warps together)\n", n, W, (double)h[0] / (4096.0 * W));
    }
  for (int var : {0, 1, 2, 3, 8, 11})
    for (int it : {4096, 8192}) {
      mma_multi<<<148, 128, 200 * 1024>>>(64, it, 1, d, var, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("morph N=64 var=%2d iters=%d: block-timed %.1f, leader-timed %.1f cyc/MMA\n", var, it, (double)h[0] / it, (double)h[1] / it);
    }
  cudaFuncSetAttribute(mma_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int G : {4, 8})
    for (int mode : {0, 1, 2, 4, 8, 3, 7, 15}) {
      const int groups = 2048;
      mma_pattern<<<148, 128, 200 * 1024>>>(64, groups, G, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("pattern N=64 G=%d mode=%2d (commit=%d wait=%d fence=%d syncwarp=%d): %.1f cyc/MMA\n", G, mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, (double)h[0] / (groups * G));
    }
  for (int layout = 0; layout < 3; ++layout)
    for (int n : {64, 128, 256}) {
      mma_rate<<<148, 128, 200 * 1024>>>(n, iters, 0, 0, d, layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("layout=%s N=%3d : %.1f cyc/MMA\n", layout == 0 ? "SW128" : layout == 1 ? "SW32" : "NOSWZ-LBO16", n, (double)h[1] / iters);
    }
  for (int kind = 0; kind < 2; ++kind)
    for (int distinct = 0; distinct < 2; ++distinct)
      for (int n : {16, 32, 64, 96, 128, 160, 192, 224, 256}) {
        for (int grid : {148}) {
          mma_rate<<<grid, 128, 200 * 1024>>>(n, iters, kind, distinct, d, 0);
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
