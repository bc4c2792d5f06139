// f32_ref.cu — the reference's FP32 network on the GPU, operator for operator:
//   * PTQ calibrator: per-tensor absmax of every activation the INT8 / FP8 path quantises, recorded while a
//     calibration batch runs through the reference's FP32 arithmetic;
//   * accuracy baseline: FP32 checkpoints and logits to compare the quantised network against, in-process (the
//     reference's harness spawns one process per image, T/bench_fp32_vs_torch_e2e.py:90-132);
//   * weight-directory loader / writer for the reference's export format (T/export_resnet18.py:85-92).
// The arithmetic follows the reference bit for bit so that calibrated scales equal the CPU oracle's:
//   conv  = sequential FMA chain over (c, kh, kw), out-of-bounds taps contribute exactly 0
//           (K/im2col.cu:32-57 + K/sgemm_tiled.cu:22-40, wrapped by R/infer_e2e.cu:102-136)
//   BN    = (x - mean) / sqrtf(var + eps), then fma(gamma, y, beta)     (K/bn_inference.cu:22-27, eps R/infer_e2e.cu:89)
//   ReLU  = x < 0 ? 0 : x   (K/relu.cu:9)     add = y + x   (K/add.cu:7)
//   pool  = max over in-bounds taps, init -FLT_MAX   (K/maxpool2d.cu:14-40)
//   GAP   = 256-lane strided partial sums + shared-memory tree, / (float)HW   (K/gap_global.cu:10-32)
//   FC    = sequential FMA over 512 inputs, bias added afterwards   (R/infer_e2e.cu:206-219)
// This is not the hot path (one thread per output, CUDA cores): it runs once per calibration / evaluation batch.
#include <cfloat>
#include <cmath>
#include <fstream>
#include <memory>

#include "dlq_internal.h"

namespace dlq {
namespace {

constexpr int kOCB = 8;   // output channels per thread (independent FMA chains)

struct F32ConvParams {
  const float* x;      // [N,C,H,W]
  const float* wt;     // transposed weights [K = C*kH*kW][OC]
  const float* g; const float* b; const float* m; const float* v;   // BN (null g: no BN)
  const float* res;    // [N,OC,OH,OW] added after BN, or null
  float* y;            // [N,OC,OH,OW]
  unsigned* absmax;    // running max |y| as float bits, or null
  int N, C, H, W, OC, kH, kW, s, p, OH, OW, relu;
  float eps;
};

__global__ void __launch_bounds__(128) conv_bn_f32_kernel(const F32ConvParams q) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int oc0 = blockIdx.y * kOCB;
  const int n = blockIdx.z;
  const bool live = pix < q.OH * q.OW;
  const int oh = live ? pix / q.OW : 0, ow = live ? pix - (pix / q.OW) * q.OW : 0;
  float acc[kOCB];
#pragma unroll
  for (int j = 0; j < kOCB; ++j) acc[j] = 0.f;
  const float* xn = q.x + static_cast<size_t>(n) * q.C * q.H * q.W;
  const int ih0 = oh * q.s - q.p, iw0 = ow * q.s - q.p;
  int k = 0;
  for (int c = 0; c < q.C; ++c) {
    const float* xc = xn + static_cast<size_t>(c) * q.H * q.W;
    for (int kh = 0; kh < q.kH; ++kh) {
      const int ih = ih0 + kh;
      for (int kw = 0; kw < q.kW; ++kw, ++k) {
        const int iw = iw0 + kw;
        // an out-of-bounds tap multiplies the weight by 0.0f in the reference's col buffer: fma(w, 0, acc) == acc
        // for finite w (acc is never -0.0f: it starts at +0 and x + (-x) rounds to +0), so the tap is skipped
        if (ih < 0 || iw < 0 || ih >= q.H || iw >= q.W) continue;
        const float xv = live ? xc[ih * q.W + iw] : 0.f;
        const float4 w0 = *reinterpret_cast<const float4*>(q.wt + static_cast<size_t>(k) * q.OC + oc0);
        const float4 w1 = *reinterpret_cast<const float4*>(q.wt + static_cast<size_t>(k) * q.OC + oc0 + 4);
        acc[0] = fmaf(w0.x, xv, acc[0]); acc[1] = fmaf(w0.y, xv, acc[1]);
        acc[2] = fmaf(w0.z, xv, acc[2]); acc[3] = fmaf(w0.w, xv, acc[3]);
        acc[4] = fmaf(w1.x, xv, acc[4]); acc[5] = fmaf(w1.y, xv, acc[5]);
        acc[6] = fmaf(w1.z, xv, acc[6]); acc[7] = fmaf(w1.w, xv, acc[7]);
      }
    }
  }
  float mx = 0.f;
  if (live) {
#pragma unroll
    for (int j = 0; j < kOCB; ++j) {
      const int oc = oc0 + j;
      float val = acc[j];
      if (q.g) {
        const float yv = __fdiv_rn(__fsub_rn(val, q.m[oc]), __fsqrt_rn(__fadd_rn(q.v[oc], q.eps)));
        val = fmaf(q.g[oc], yv, q.b[oc]);
      }
      const size_t o = ((static_cast<size_t>(n) * q.OC + oc) * q.OH + oh) * q.OW + ow;
      if (q.res) val = __fadd_rn(val, q.res[o]);
      if (q.relu && val < 0.f) val = 0.f;
      q.y[o] = val;
      mx = fmaxf(mx, fabsf(val));
    }
  }
  if (q.absmax) {
#pragma unroll
    for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(q.absmax, __float_as_uint(mx));   // non-negative floats order as uints
  }
}

__global__ void absmax_f32_kernel(const float* x, size_t n, unsigned* out) {
  float mx = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    mx = fmaxf(mx, fabsf(x[i]));
#pragma unroll
  for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out, __float_as_uint(mx));
}

__global__ void maxpool_f32_kernel(const float* x, int N, int C, int H, int W, float* y) {
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const size_t total = static_cast<size_t>(N) * C * OH * OW;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % OW), oh = static_cast<int>((i / OW) % OH);
    const size_t nc = i / (static_cast<size_t>(OW) * OH);
    const float* xp = x + nc * H * W;
    float m = -FLT_MAX;
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        const float v = xp[ih * W + iw];
        if (v > m) m = v;
      }
    }
    y[i] = m;
  }
}

// one block of 256 threads per (n, c): the reference's partial sums and tree, K/gap_global.cu:10-32
__global__ void __launch_bounds__(256) gap_f32_kernel(const float* x, int HW, float* y, unsigned* absmax) {
  __shared__ float sm[256];
  const float* xp = x + static_cast<size_t>(blockIdx.x) * HW;
  float sum = 0.f;
  for (int i = threadIdx.x; i < HW; i += 256) sum = __fadd_rn(sum, xp[i]);
  sm[threadIdx.x] = sum;
  __syncthreads();
  for (int s = 128; s > 1; s >>= 1) {
    if (static_cast<int>(threadIdx.x) < s) sm[threadIdx.x] = __fadd_rn(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float g = __fdiv_rn(__fadd_rn(sm[0], sm[1]), static_cast<float>(HW));
    y[blockIdx.x] = g;
    if (absmax && fabsf(g) > 0.f) atomicMax(absmax, __float_as_uint(fabsf(g)));
  }
}

__global__ void fc_f32_kernel(const float* gap, const float* w, const float* b, int N, int O, int I, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * O) return;
  const int n = i / O, o = i - n * O;
  float acc = 0.f;
  for (int k = 0; k < I; ++k) acc = fmaf(w[static_cast<size_t>(o) * I + k], gap[static_cast<size_t>(n) * I + k], acc);
  out[i] = __fadd_rn(acc, b[o]);
}

// top-k by repeated arg-max (k is small); ties resolve to the lowest index, like the reference's strict '>' scan
// (R/infer_e2e.cu:436-438)
__global__ void __launch_bounds__(256) topk_f32_kernel(const float* x, int K, int k, int* idx, float* val) {
  __shared__ float sv[256];
  __shared__ int si[256];
  __shared__ int taken[32];
  const float* row = x + static_cast<size_t>(blockIdx.x) * K;
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < K; i += 256) {
      bool skip = false;
      for (int t = 0; t < r; ++t) skip |= taken[t] == i;
      const float v = row[i];
      if (!skip && (bi == 0x7fffffff || v > best)) { best = v; bi = i; }   // (i ascends: ties keep the lower index)
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
      if (static_cast<int>(threadIdx.x) < s) {
        const float v2 = sv[threadIdx.x + s]; const int i2 = si[threadIdx.x + s];
        if (i2 != 0x7fffffff && (si[threadIdx.x] == 0x7fffffff || v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x]))) {
          sv[threadIdx.x] = v2; si[threadIdx.x] = i2;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      taken[r] = si[0];
      idx[static_cast<size_t>(blockIdx.x) * k + r] = si[0];
      if (val) val[static_cast<size_t>(blockIdx.x) * k + r] = sv[0];
    }
    __syncthreads();
  }
}

// max |a-b|, sum |a-b|, dot(a,b), |a|^2, |b|^2 in double (R/utils.hpp:163-177 diff_max_mean; T/diag_e2e_compare.py:12-24)
__global__ void __launch_bounds__(256) compare_f32_kernel(const float* a, const float* b, size_t n, double* out /*[5]*/) {
  double mx = 0, sa = 0, dot = 0, na = 0, nb = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double x = a[i], y = b[i], d = fabs(x - y);
    mx = fmax(mx, d); sa += d; dot += x * y; na += x * x; nb += y * y;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    sa += __shfl_xor_sync(0xffffffffu, sa, d); dot += __shfl_xor_sync(0xffffffffu, dot, d);
    na += __shfl_xor_sync(0xffffffffu, na, d); nb += __shfl_xor_sync(0xffffffffu, nb, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(out), static_cast<unsigned long long>(__double_as_longlong(mx)));
    atomicAdd(out + 1, sa); atomicAdd(out + 2, dot); atomicAdd(out + 3, na); atomicAdd(out + 4, nb);
  }
}

struct ConvSpec { int ic, oc, k, s, p; };
const ConvSpec* conv_specs() {
  static ConvSpec S[DLQ_NUM_CONVS];
  static bool init = false;
  if (!init) {
    S[0] = {3, 64, 7, 2, 3};
    const int chans[4] = {64, 128, 256, 512};
    int ic = 64;
    for (int b = 0; b < 8; ++b) {
      const int oc = chans[b / 2];
      const int s = (b % 2 == 0 && b >= 2) ? 2 : 1;
      S[1 + 3 * b] = {ic, oc, 3, s, 1};
      S[2 + 3 * b] = {oc, oc, 3, 1, 1};
      S[3 + 3 * b] = (s == 2) ? ConvSpec{ic, oc, 1, 2, 0} : ConvSpec{0, 0, 0, 0, 0};
      ic = oc;
    }
    init = true;
  }
  return S;
}

}  // namespace
}  // namespace dlq

using namespace dlq;

// ------------------------------------------------------------------ FP32 network object
struct dlq_resnet18_f32 {
  dlq_ctx* ctx = nullptr;
  int max_batch = 0;
  float* wt[DLQ_NUM_CONVS] = {nullptr};     // transposed [K][OC]
  float* bn[DLQ_NUM_CONVS][4] = {{nullptr}};
  float* fc_w = nullptr;
  float* fc_b = nullptr;
  float* buf[4] = {nullptr, nullptr, nullptr, nullptr};   // activation ping-pong (largest: N x 64 x 112 x 112)
  float* ck[6] = {nullptr};                 // stem_pool, layer1..4, gap (copies of the last forward)
  unsigned* absmax = nullptr;               // [DLQ_NUM_ACTS] float bits
  int last_n = 0;
};

static const char* kCkNames[6] = {"stem_pool", "layer1", "layer2", "layer3", "layer4", "gap"};
static const size_t kCkElems[6] = {64 * 56 * 56, 64 * 56 * 56, 128 * 28 * 28, 256 * 14 * 14, 512 * 7 * 7, 512};

extern "C" {

void dlq_resnet18_f32_destroy(dlq_resnet18_f32* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    cudaFree(m->wt[i]);
    for (int j = 0; j < 4; ++j) cudaFree(m->bn[i][j]);
  }
  cudaFree(m->fc_w); cudaFree(m->fc_b);
  for (float* b : m->buf) cudaFree(b);
  for (float* c : m->ck) cudaFree(c);
  cudaFree(m->absmax);
  delete m;
}

int dlq_resnet18_f32_create(dlq_ctx* ctx, const dlq_resnet18_weights* w, int max_batch, dlq_resnet18_f32** out) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, w && out && max_batch >= 1, "null weights / output, or max_batch < 1");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<dlq_resnet18_f32, void (*)(dlq_resnet18_f32*)> m(new dlq_resnet18_f32, dlq_resnet18_f32_destroy);
  m->ctx = ctx;
  m->max_batch = max_batch;
  const ConvSpec* S = conv_specs();
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    if (!S[i].oc) continue;
    DLQ_ARG(ctx, w->conv_w[i] && w->bn_gamma[i] && w->bn_beta[i] && w->bn_mean[i] && w->bn_var[i], "missing conv / BN tensor");
    const int K = S[i].ic * S[i].k * S[i].k, OC = S[i].oc;
    std::vector<float> t(static_cast<size_t>(K) * OC);
    for (int o = 0; o < OC; ++o)
      for (int k = 0; k < K; ++k) t[static_cast<size_t>(k) * OC + o] = w->conv_w[i][static_cast<size_t>(o) * K + k];
    DLQ_CUDA(ctx, cudaMalloc(&m->wt[i], t.size() * sizeof(float)));
    DLQ_CUDA(ctx, cudaMemcpy(m->wt[i], t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    const float* src[4] = {w->bn_gamma[i], w->bn_beta[i], w->bn_mean[i], w->bn_var[i]};
    for (int j = 0; j < 4; ++j) {
      DLQ_CUDA(ctx, cudaMalloc(&m->bn[i][j], OC * sizeof(float)));
      DLQ_CUDA(ctx, cudaMemcpy(m->bn[i][j], src[j], OC * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  DLQ_ARG(ctx, w->fc_w && w->fc_b, "missing FC tensor");
  DLQ_CUDA(ctx, cudaMalloc(&m->fc_w, 1000 * 512 * sizeof(float)));
  DLQ_CUDA(ctx, cudaMemcpy(m->fc_w, w->fc_w, 1000 * 512 * sizeof(float), cudaMemcpyHostToDevice));
  DLQ_CUDA(ctx, cudaMalloc(&m->fc_b, 1000 * sizeof(float)));
  DLQ_CUDA(ctx, cudaMemcpy(m->fc_b, w->fc_b, 1000 * sizeof(float), cudaMemcpyHostToDevice));
  const size_t big = static_cast<size_t>(max_batch) * 64 * 112 * 112, mid = static_cast<size_t>(max_batch) * 64 * 56 * 56;
  DLQ_CUDA(ctx, cudaMalloc(&m->buf[0], big * sizeof(float)));
  for (int i = 1; i < 4; ++i) DLQ_CUDA(ctx, cudaMalloc(&m->buf[i], mid * sizeof(float)));
  for (int i = 0; i < 6; ++i) DLQ_CUDA(ctx, cudaMalloc(&m->ck[i], static_cast<size_t>(max_batch) * kCkElems[i] * sizeof(float)));
  DLQ_CUDA(ctx, cudaMalloc(&m->absmax, DLQ_NUM_ACTS * sizeof(unsigned)));
  DLQ_CUDA(ctx, cudaMemset(m->absmax, 0, DLQ_NUM_ACTS * sizeof(unsigned)));
  *out = m.release();
  return DLQ_OK;
}

static int f32_conv(dlq_resnet18_f32* m, int ci, const float* x, int N, int H, int W, const float* res, int relu,
                    float* y, int track, int* OH, int* OW) {
  dlq_ctx* ctx = m->ctx;
  const ConvSpec& s = conv_specs()[ci];
  F32ConvParams q;
  q.x = x; q.wt = m->wt[ci];
  q.g = m->bn[ci][0]; q.b = m->bn[ci][1]; q.m = m->bn[ci][2]; q.v = m->bn[ci][3];
  q.res = res; q.y = y; q.absmax = track >= 0 ? m->absmax + track : nullptr;
  q.N = N; q.C = s.ic; q.H = H; q.W = W; q.OC = s.oc; q.kH = s.k; q.kW = s.k; q.s = s.s; q.p = s.p;
  q.OH = (H + 2 * s.p - s.k) / s.s + 1; q.OW = (W + 2 * s.p - s.k) / s.s + 1;
  q.relu = relu; q.eps = 1e-5f;   // R/infer_e2e.cu:89
  *OH = q.OH; *OW = q.OW;
  const dim3 grid((q.OH * q.OW + 127) / 128, s.oc / kOCB, N);
  conv_bn_f32_kernel<<<grid, 128, 0, ctx->stream>>>(q);
  DLQ_CUDA(ctx, cudaGetLastError());
  return DLQ_OK;
}

int dlq_resnet18_f32_forward(dlq_resnet18_f32* m, const float* x, int N, float* logits) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, x && logits && N >= 1 && N <= m->max_batch, "null pointer or batch outside [1, max_batch]");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int rc, H = 224, W = 224, OH, OW;
  absmax_f32_kernel<<<296, 256, 0, st>>>(x, static_cast<size_t>(N) * 3 * H * W, m->absmax + 0);
  if ((rc = f32_conv(m, 0, x, N, H, W, nullptr, 1, m->buf[0], 1, &OH, &OW))) return rc;      // stem: conv + BN + ReLU
  float* cur = m->buf[1];
  maxpool_f32_kernel<<<1184, 256, 0, st>>>(m->buf[0], N, 64, OH, OW, cur);
  H = (OH + 2 - 3) / 2 + 1; W = (OW + 2 - 3) / 2 + 1;
  DLQ_CUDA(ctx, cudaMemcpyAsync(m->ck[0], cur, static_cast<size_t>(N) * kCkElems[0] * sizeof(float), cudaMemcpyDeviceToDevice, st));
  float* t1 = m->buf[2];
  float* t2 = m->buf[3];
  float* sk = m->buf[0];
  int C = 64;
  for (int b = 0; b < 8; ++b) {
    const int i1 = 1 + 3 * b, i2 = 2 + 3 * b, id = 3 + 3 * b;
    int H1, W1, H2, W2;
    if ((rc = f32_conv(m, i1, cur, N, H, W, nullptr, 1, t1, 2 + 3 * b, &H1, &W1))) return rc;
    const float* res = cur;
    if (conv_specs()[id].oc) {
      int Hd, Wd;
      if ((rc = f32_conv(m, id, cur, N, H, W, nullptr, 0, sk, 3 + 3 * b, &Hd, &Wd))) return rc;
      res = sk;
    }
    if ((rc = f32_conv(m, i2, t1, N, H1, W1, res, 1, t2, 4 + 3 * b, &H2, &W2))) return rc;     // BN, + skip, ReLU
    std::swap(cur, t2);
    H = H2; W = W2; C = conv_specs()[i2].oc;
    if (b & 1)
      DLQ_CUDA(ctx, cudaMemcpyAsync(m->ck[1 + b / 2], cur, static_cast<size_t>(N) * C * H * W * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  gap_f32_kernel<<<N * C, 256, 0, st>>>(cur, H * W, m->ck[5], m->absmax + DLQ_NUM_ACTS - 1);
  fc_f32_kernel<<<(N * 1000 + 127) / 128, 128, 0, st>>>(m->ck[5], m->fc_w, m->fc_b, N, 1000, 512, logits);
  DLQ_CUDA(ctx, cudaGetLastError());
  m->last_n = N;
  return DLQ_OK;
}

int dlq_resnet18_f32_checkpoint(dlq_resnet18_f32* m, const char* name, float* out) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, name && out && m->last_n > 0, "no forward has run, or null argument");
  for (int i = 0; i < 6; ++i)
    if (!strcmp(name, kCkNames[i])) {
      DLQ_CUDA(ctx, cudaMemcpyAsync(out, m->ck[i], static_cast<size_t>(m->last_n) * kCkElems[i] * sizeof(float),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
      return DLQ_OK;
    }
  ctx->err = std::string("bad argument: unknown checkpoint '") + name + "'";
  return DLQ_ERR_ARG;
}

int dlq_resnet18_f32_absmax(dlq_resnet18_f32* m, float* absmax /*[DLQ_NUM_ACTS]*/) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, absmax, "null output");
  DLQ_CUDA(ctx, cudaMemcpyAsync(absmax, m->absmax, DLQ_NUM_ACTS * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DLQ_OK;
}

int dlq_resnet18_f32_reset_absmax(dlq_resnet18_f32* m) {
  if (!m) return DLQ_ERR_ARG;
  DLQ_CUDA(m->ctx, cudaMemsetAsync(m->absmax, 0, DLQ_NUM_ACTS * sizeof(unsigned), m->ctx->stream));
  return DLQ_OK;
}

// QUANT_SPEC 2 / 6: scale = absmax / qmax, computed in double and rounded once; a tensor the network never produces
// (blocks without a downsample branch) or that is identically zero calibrates as absmax 1
void dlq_act_scales_from_absmax(const float* absmax, int fp8, float* act_scale) {
  const double qmax = fp8 ? 448.0 : 127.0;
  for (int i = 0; i < DLQ_NUM_ACTS; ++i)
    act_scale[i] = static_cast<float>(static_cast<double>(absmax[i] > 0.f ? absmax[i] : 1.f) / qmax);
}

int dlq_topk_f32(dlq_ctx* ctx, const float* x, int N, int K, int k, int* idx, float* val) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, x && idx && N >= 1 && K >= 1 && k >= 1 && k <= 32 && k <= K, "null pointer, or k outside [1, min(32, K)]");
  topk_f32_kernel<<<N, 256, 0, ctx->stream>>>(x, K, k, idx, val);
  DLQ_CUDA(ctx, cudaGetLastError());
  return DLQ_OK;
}

int dlq_compare_f32(dlq_ctx* ctx, const float* a, const float* b, size_t n, double* out3) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, a && b && out3 && n > 0, "null pointer or empty tensor");
  double* d = static_cast<double*>(ctx->small);
  DLQ_CUDA(ctx, cudaMemsetAsync(d, 0, 5 * sizeof(double), ctx->stream));
  compare_f32_kernel<<<296, 256, 0, ctx->stream>>>(a, b, n, d);
  double h[5];
  DLQ_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out3[0] = h[0];
  out3[1] = h[1] / static_cast<double>(n);
  out3[2] = (h[3] > 0 && h[4] > 0) ? h[2] / (std::sqrt(h[3]) * std::sqrt(h[4])) : 0.0;   // T/diag_e2e_compare.py:12-17
  return DLQ_OK;
}

// ------------------------------------------------------------------ weight directory (reference export format)
struct dlq_weight_dir {
  std::vector<std::vector<float>> store;
  dlq_resnet18_weights w;
  std::string err;
};

static std::string conv_key(int i) {
  if (i == 0) return "conv1";
  const int b = (i - 1) / 3, r = (i - 1) % 3;
  const std::string base = "layer" + std::to_string(b / 2 + 1) + "." + std::to_string(b % 2);
  return r == 0 ? base + ".conv1" : r == 1 ? base + ".conv2" : base + ".downsample.0";
}
static std::string bn_key(int i) {
  if (i == 0) return "bn1";
  const int b = (i - 1) / 3, r = (i - 1) % 3;
  const std::string base = "layer" + std::to_string(b / 2 + 1) + "." + std::to_string(b % 2);
  return r == 0 ? base + ".bn1" : r == 1 ? base + ".bn2" : base + ".downsample.1";
}

// R/utils.hpp:48-60 load_bin_f32: raw little-endian fp32, size checked; returns an error instead of exiting
static bool read_bin(const std::string& path, size_t expected, std::vector<float>& v, std::string& err) {
  std::ifstream ifs(path, std::ios::binary);
  if (!ifs) { err = "open fail: " + path; return false; }
  ifs.seekg(0, std::ios::end);
  const size_t bytes = static_cast<size_t>(ifs.tellg());
  ifs.seekg(0);
  if (bytes % 4) { err = "size not float-aligned: " + path; return false; }
  if (bytes / 4 != expected) {
    err = "unexpected size: " + path + " got " + std::to_string(bytes / 4) + " expected " + std::to_string(expected);
    return false;
  }
  v.resize(expected);
  if (expected) ifs.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(bytes));
  return static_cast<bool>(ifs);
}
static bool write_bin(const std::string& path, const float* v, size_t n) {
  std::ofstream ofs(path, std::ios::binary);
  if (!ofs) return false;
  ofs.write(reinterpret_cast<const char*>(v), static_cast<std::streamsize>(n * sizeof(float)));
  return static_cast<bool>(ofs);
}

static void set_err(char* err, size_t err_len, const std::string& s) {
  if (err && err_len) { strncpy(err, s.c_str(), err_len - 1); err[err_len - 1] = 0; }
}

int dlq_weight_dir_load(const char* dir, dlq_weight_dir** out, char* err, size_t err_len) {
  if (!dir || !out) { set_err(err, err_len, "bad argument: null directory / output"); return DLQ_ERR_ARG; }
  std::unique_ptr<dlq_weight_dir> d(new dlq_weight_dir);
  memset(&d->w, 0, sizeof(d->w));
  const std::string root(dir);
  const ConvSpec* S = conv_specs();
  auto load = [&](const std::string& key, size_t n, const float** dst) -> bool {
    d->store.emplace_back();
    if (!read_bin(root + "/" + key + ".bin", n, d->store.back(), d->err)) return false;
    *dst = d->store.back().data();
    return true;
  };
  d->store.reserve(DLQ_NUM_CONVS * 5 + 4);
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    if (!S[i].oc) continue;
    const size_t nw = static_cast<size_t>(S[i].oc) * S[i].ic * S[i].k * S[i].k, oc = S[i].oc;
    if (!load(conv_key(i) + ".weight", nw, &d->w.conv_w[i]) || !load(bn_key(i) + ".weight", oc, &d->w.bn_gamma[i]) ||
        !load(bn_key(i) + ".bias", oc, &d->w.bn_beta[i]) || !load(bn_key(i) + ".running_mean", oc, &d->w.bn_mean[i]) ||
        !load(bn_key(i) + ".running_var", oc, &d->w.bn_var[i])) {
      set_err(err, err_len, d->err);
      return DLQ_ERR_ARG;
    }
  }
  if (!load("fc.weight", 1000 * 512, &d->w.fc_w) || !load("fc.bias", 1000, &d->w.fc_b)) {
    set_err(err, err_len, d->err);
    return DLQ_ERR_ARG;
  }
  // optional: activation scales written by dlq_weight_dir_save_scales (the "quant" block RKL/reports/Step1.md:92 plans)
  {
    std::vector<float> s;
    std::string e2;
    if (read_bin(root + "/quant.act_scale.int8.bin", DLQ_NUM_ACTS, s, e2))
      for (int i = 0; i < DLQ_NUM_ACTS; ++i) d->w.act_scale[i] = s[i];
  }
  *out = d.release();
  return DLQ_OK;
}

const dlq_resnet18_weights* dlq_weight_dir_weights(const dlq_weight_dir* d) { return d ? &d->w : nullptr; }
void dlq_weight_dir_free(dlq_weight_dir* d) { delete d; }

int dlq_weight_dir_save(const char* dir, const dlq_resnet18_weights* w, int with_scales) {
  if (!dir || !w) return DLQ_ERR_ARG;
  const std::string root(dir);
  const ConvSpec* S = conv_specs();
  std::string manifest = "{\n  \"model\": \"resnet18\",\n  \"dtype\": \"fp32\",\n  \"layout\": \"NCHW\",\n  \"version\": 1,\n"
                         "  \"preprocess\": {\"resize\": 256, \"center_crop\": 224, \"mean\": [0.485, 0.456, 0.406], "
                         "\"std\": [0.229, 0.224, 0.225]},\n  \"tensors\": {\n";
  bool first = true;
  auto put = [&](const std::string& key, const float* v, const std::string& shape, const char* layout, const char* kind,
                 size_t n) -> bool {
    if (!v || !write_bin(root + "/" + key + ".bin", v, n)) return false;
    manifest += std::string(first ? "" : ",\n") + "    \"" + key + "\": {\"shape\": " + shape + ", \"layout\": \"" + layout +
                "\", \"kind\": \"" + kind + "\", \"path\": \"" + key + ".bin\"}";
    first = false;
    return true;
  };
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    if (!S[i].oc) continue;
    const std::string oc = std::to_string(S[i].oc);
    const std::string shp = "[" + oc + ", " + std::to_string(S[i].ic) + ", " + std::to_string(S[i].k) + ", " + std::to_string(S[i].k) + "]";
    const size_t nw = static_cast<size_t>(S[i].oc) * S[i].ic * S[i].k * S[i].k;
    if (!put(conv_key(i) + ".weight", w->conv_w[i], shp, "OIHW", "conv_weight", nw) ||
        !put(bn_key(i) + ".weight", w->bn_gamma[i], "[" + oc + "]", "O", "bn_param", S[i].oc) ||
        !put(bn_key(i) + ".bias", w->bn_beta[i], "[" + oc + "]", "O", "bn_param", S[i].oc) ||
        !put(bn_key(i) + ".running_mean", w->bn_mean[i], "[" + oc + "]", "O", "bn_buffer", S[i].oc) ||
        !put(bn_key(i) + ".running_var", w->bn_var[i], "[" + oc + "]", "O", "bn_buffer", S[i].oc))
      return DLQ_ERR_ARG;
  }
  if (!put("fc.weight", w->fc_w, "[1000, 512]", "OI", "fc_weight", 1000 * 512) ||
      !put("fc.bias", w->fc_b, "[1000]", "O", "fc_bias", 1000))
    return DLQ_ERR_ARG;
  manifest += "\n  }";
  if (with_scales) {
    if (!write_bin(root + "/quant.act_scale.int8.bin", w->act_scale, DLQ_NUM_ACTS)) return DLQ_ERR_ARG;
    manifest += ",\n  \"quant\": {\"scheme\": \"int8 per-channel weights / per-tensor activations (spec/QUANT_SPEC.md)\", "
                "\"act_scale\": \"quant.act_scale.int8.bin\", \"num_act_scales\": " + std::to_string(DLQ_NUM_ACTS) + "}";
  }
  manifest += "\n}\n";
  std::ofstream mf(root + "/manifest.json");
  if (!mf) return DLQ_ERR_ARG;
  mf << manifest;
  return mf ? DLQ_OK : DLQ_ERR_ARG;
}

}  // extern "C"
