// ldc_rate.cu — how should the conv epilogue fetch its per-channel (alpha, beta)?  Compares, for 8 warps / SM,
//   A: broadcast LDS.128 from shared memory (4 wavefronts per instruction on the smem data pipe)
//   B: LDC.64 from the kernel-parameter constant bank with a warp-uniform runtime offset
// Prints cycles per 32-channel block (32 x fmaf(acc, alpha, beta) per lane).  Run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct P { float2 ab[512]; int n0; int iters; float* out; const int* in; long long* cyc; };

__global__ void __launch_bounds__(256, 1) k_const(const __grid_constant__ P p) {
  float accv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) accv[j] = (float)p.in[(threadIdx.x + j * 256) & 1023];
  float s = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < p.iters; ++it) {
    const int c0 = (p.n0 + it * 32) & 511 & ~31;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float2 ab = p.ab[c0 + j];
      s += __fmaf_rn(accv[j], ab.x, ab.y);
    }
  }
  long long t1 = clock64();
  p.out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) p.cyc[0] = t1 - t0;
}

__global__ void __launch_bounds__(256, 1) k_smem(const float* alpha, const float* beta, int n0, int iters, float* out,
                                                 const int* in, long long* cyc) {
  __shared__ __align__(16) float sa[512], sb[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) { sa[i] = alpha[i]; sb[i] = beta[i]; }
  float accv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) accv[j] = (float)in[(threadIdx.x + j * 256) & 1023];
  float s = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int c0 = (n0 + it * 32) & 511 & ~31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 a = reinterpret_cast<const float4*>(sa + c0)[j];
      const float4 b = reinterpret_cast<const float4*>(sb + c0)[j];
      s += __fmaf_rn(accv[4 * j], a.x, b.x);
      s += __fmaf_rn(accv[4 * j + 1], a.y, b.y);
      s += __fmaf_rn(accv[4 * j + 2], a.z, b.z);
      s += __fmaf_rn(accv[4 * j + 3], a.w, b.w);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  P p;
  for (int i = 0; i < 512; ++i) p.ab[i] = make_float2(1.f + i * 1e-3f, 0.5f * i);
  float *da, *db, *out; int* in; long long* cyc;
  cudaMalloc(&da, 2048); cudaMalloc(&db, 2048); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 8);
  cudaMemset(da, 0, 2048); cudaMemset(db, 0, 2048); cudaMemset(in, 1, 4096);
  p.n0 = 0; p.iters = 4096; p.out = out; p.in = in; p.cyc = cyc;
  long long h;
  for (int rep = 0; rep < 2; ++rep) {
    k_const<<<148, 256>>>(p);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("B  LDC.64 param bank : %.1f cycles per 32-channel block per warp-set (8 warps/SM)\n", (double)h / p.iters);
    k_smem<<<148, 256>>>(da, db, 0, p.iters, out, in, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("A  LDS.128 broadcast : %.1f cycles per 32-channel block per warp-set (8 warps/SM)\n", (double)h / p.iters);
  }
  return 0;
}
