// cvt_rate.cu — issue rate of the conversion instructions the conv epilogue leans on, per SM sub-partition:
//   I2FP.F32.S32 (cvt.rn.f32.s32), F2IP.S8.F32 (cvt.rni.s32.f32 x2 + cvt.pack.sat.s8.s32), against FFMA and IADD3,
//   for 1, 2 and 4 warps per scheduler (the conv kernel runs 2 epilogue warps per scheduler).
// Prints cycles per warp-instruction per scheduler.  Run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(int iters, const int* in, int* out, long long* cyc) {
  int v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = in[(threadIdx.x * 16 + j) & 1023];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (OP == 0) {                 // int -> float -> back through the bit pattern (keeps a dependency per register)
        v[j] = __float_as_int(static_cast<float>(v[j])) & 0x00ffffff;
      } else if (OP == 1) {          // FFMA
        v[j] = __float_as_int(__fmaf_rn(__int_as_float(v[j]), 1.0001f, 0.5f));
      } else if (OP == 2) {          // IADD3 / LOP3
        v[j] = (v[j] + it) ^ 0x5555;
      }
    }
    if (OP == 3) {                   // two floats -> packed saturated int8 pair (F2IP.S8.F32)
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const int a = __float2int_rn(__int_as_float(v[j])), b = __float2int_rn(__int_as_float(v[j + 1]));
        uint32_t r;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(0));
        v[j] = static_cast<int>(r) + 0x3f800000;
      }
    }
  }
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  int *in, *out;
  long long* cyc;
  cudaMalloc(&in, 4096);
  cudaMemset(in, 1, 4096);
  cudaMalloc(&out, 4 * 1024 * 512);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const char* names[4] = {"I2FP.F32.S32 (+LOP3)", "FFMA", "IADD3+LOP3", "F2I x2 + cvt.pack (F2IP.S8)"};
  const double per_iter[4] = {16, 16, 16, 8};
  for (int op = 0; op < 4; ++op)
    for (int wps = 1; wps <= 4; wps *= 2) {
      const int threads = wps * 4 * 32;
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, threads>>>(iters, in, out, cyc);
        if (op == 1) k<1><<<148, threads>>>(iters, in, out, cyc);
        if (op == 2) k<2><<<148, threads>>>(iters, in, out, cyc);
        if (op == 3) k<3><<<148, threads>>>(iters, in, out, cyc);
        cudaDeviceSynchronize();
      }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-30s warps/scheduler %d: %.2f cycles per op-group per scheduler (%.2f per warp)\n", names[op], wps,
             (double)c / (iters * per_iter[op] * wps), (double)c / (iters * per_iter[op]));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
