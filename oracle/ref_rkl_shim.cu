// ref_rkl_shim.cu — host-callable launchers around the REFERENCE's own FP32 kernels
// (/root/reference/CUDA/resnet18-kernel-lab/cpp/fp32/kernels/*.cu, compiled unmodified by
// oracle/Makefile `make ref`).  TEST INFRASTRUCTURE ONLY: gives the GPU tests a live run of the
// reference implementation to pin the CPU oracle against.  Launch geometry is the reference's:
//   gemm  (32,32) blocks                         runtime/infer_e2e.cu:64-70
//   im2col (16,16) blocks over (OW,OH)           runtime/infer_e2e.cu:72-81
//   bn / relu / add  256 threads                 runtime/infer_e2e.cu:95-96, :169, :199
//   maxpool  grid (OH*OW, C, N), block (1,1,1)   runtime/infer_e2e.cu:289-291
//   gap  grid C, block 256                       runtime/infer_head.cu:68-69
//   softmax grid 1, block 256                    runtime/infer_head.cu:98-99
// All pointers are DEVICE pointers; every wrapper synchronises and returns the cudaError_t as int.
#include <cuda_runtime.h>

extern "C" {
// prototypes as the reference re-declares them (runtime/utils.hpp:75-82, runtime/infer_e2e.cu:14-32)
__global__ void im2col_nchw(const float*, int, int, int, int, int, int, int, int, int, int, float*);
__global__ void sgemm_tiled(const float*, const float*, float*, int, int, int);
__global__ void bn_inference(float*, const float*, const float*, const float*, const float*, float, int, int, int);
__global__ void relu_forward(float*, int);
__global__ void add_inplace(float*, const float*, int);
__global__ void maxpool2d_3x3_s2p1_nchw(const float*, int, int, int, int, float*);
__global__ void gap_global(const float*, int, int, int, float*);
__global__ void softmax_1d(const float*, int, float*);
}

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline int done() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}

extern "C" {
int ref_im2col_nchw(const float* x, int N, int C, int H, int W, int kH, int kW, int sH, int sW, int pH, int pW,
                    float* col) {
  const int OH = (H + 2 * pH - kH) / sH + 1, OW = (W + 2 * pW - kW) / sW + 1;
  dim3 blk(16, 16), grd(div_up(OW, 16), div_up(OH, 16));
  im2col_nchw<<<grd, blk>>>(x, N, C, H, W, kH, kW, sH, sW, pH, pW, col);
  return done();
}
int ref_sgemm_tiled(const float* A, const float* B, float* C, int M, int N, int K) {
  dim3 blk(32, 32), grd(div_up(N, 32), div_up(M, 32));
  sgemm_tiled<<<grd, blk>>>(A, B, C, M, N, K);
  return done();
}
int ref_bn_inference(float* x, const float* g, const float* b, const float* m, const float* v, float eps, int C,
                     int OH, int OW) {
  const int total = C * OH * OW;
  bn_inference<<<div_up(total, 256), 256>>>(x, g, b, m, v, eps, C, OH, OW);
  return done();
}
int ref_relu_forward(float* x, int n) {
  relu_forward<<<div_up(n, 256), 256>>>(x, n);
  return done();
}
int ref_add_inplace(float* y, const float* x, int n) {
  add_inplace<<<div_up(n, 256), 256>>>(y, x, n);
  return done();
}
int ref_maxpool2d_3x3_s2p1_nchw(const float* x, int N, int C, int H, int W, float* y) {
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  dim3 blk(1, 1, 1), grd(OH * OW, C, N);
  maxpool2d_3x3_s2p1_nchw<<<grd, blk>>>(x, N, C, H, W, y);
  return done();
}
int ref_gap_global(const float* x, int C, int H, int W, float* y) {
  gap_global<<<C, 256>>>(x, C, H, W, y);
  return done();
}
int ref_softmax_1d(const float* x, int K, float* y) {
  softmax_1d<<<1, 256>>>(x, K, y);
  return done();
}
}
