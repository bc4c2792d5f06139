// fc_mlp.cu — fully-connected layers on the tcgen05 conv core, and the MNIST MLP forward behind the C ABI.
//
// A fully-connected layer  y[n, o] = sum_i x[n, i] * W[o, i]  is the 1x1 convolution of an [N, 1, 1, I] tensor
// (GEMM view M = N, K = I, N = O), so it runs on conv_i8_kernel unchanged: TMA-staged activation rows and weight
// steps, tcgen05.mma kind::i8 / kind::f8f6f4 into TMEM, epilogue either requantising to int8 (hidden layers) or
// writing fp32  fmaf(acc, scale[o], bias[o])  (logits; bias after the product, as the reference adds it on the host).
//
// Reference anchors: runtime/infer_e2e.cu:206-219 (fc_forward = sgemm_tiled with N = 1 + host bias),
// CUDA/MNIST_on_GPU/v4.cu:255-302 (forward_timed: matmul -> bias -> relu -> matmul -> bias -> softmax),
// v5.cu:127-157 (forward_pass_only, the cuBLAS twin), v3.c:125-215 (the CPU form BASELINE config 0 times).
#include "dlq_internal.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>

using namespace dlq;

namespace {

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int fc_pad_in(int I) { return I <= 64 ? 64 : round_up(I, 128); }      // the conv core takes IC = 64 or k * 128
inline int fc_pad_out(int O) { return round_up(O, 64); }
inline size_t al1k(size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); }

// [N, I] bytes -> [N, Ip] bytes, zero fill (int8 0 and E4M3 +0 are both 0x00)
__global__ void pad_rows_u8_kernel(const uint8_t* __restrict__ x, int N, int I, int Ip, uint8_t* __restrict__ y) {
  const size_t total = static_cast<size_t>(N) * Ip;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % Ip);
    const size_t n = i / Ip;
    y[i] = c < I ? x[n * I + c] : static_cast<uint8_t>(0);
  }
}
__global__ void pad_vec_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, int O, int Op,
                                   float* __restrict__ ap, float* __restrict__ bp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Op) {
    ap[i] = i < O ? a[i] : 0.f;
    bp[i] = i < O ? b[i] : 0.f;
  }
}
__global__ void unpad_cols_f32_kernel(const float* __restrict__ z, int N, int O, int Op, float* __restrict__ y) {
  const size_t total = static_cast<size_t>(N) * O;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % O);
    const size_t n = i / O;
    y[i] = z[n * Op + c];
  }
}

// fp32 [B, K] -> quantised [B, Kp] bytes (columns >= K stay zero: the buffer is zeroed once at creation).
// QUANT_SPEC 2 (int8: clamp(rne(x * inv_s), -128, 127)) or 6 (E4M3 round-to-nearest-even, saturate to +-448)
template <bool FP8>
__global__ void quantize_rows_kernel(const float* __restrict__ x, int B, int K, int Kp, float inv_s, uint8_t* __restrict__ q) {
  const size_t total = static_cast<size_t>(B) * K;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % K);
    const size_t n = i / K;
    const float t = __fmul_rn(__ldg(x + i), inv_s);
    uint8_t v;
    if (FP8) {
      v = static_cast<uint8_t>(__nv_cvt_float_to_fp8(t, __NV_SATFINITE, __NV_E4M3));
    } else {
      const int r = max(-128, min(127, __float2int_rn(t)));
      v = static_cast<uint8_t>(static_cast<int8_t>(r));
    }
    q[n * Kp + c] = v;
  }
}

// logits [B, Op] (padded) -> logits [B, O] and, optionally, row softmax probabilities [B, O] (K/softmax.cu:6-47 form:
// exp(x - max) / sum, expf).  One thread per row (O is small: 10 for MNIST).
__global__ void mlp_head_kernel(const float* __restrict__ z, int B, int O, int Op, float* __restrict__ logits,
                                float* __restrict__ probs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B) return;
  const float* zr = z + static_cast<size_t>(n) * Op;
  float m = -INFINITY;
  for (int o = 0; o < O; ++o) {
    const float v = zr[o];
    if (logits) logits[static_cast<size_t>(n) * O + o] = v;
    m = fmaxf(m, v);
  }
  if (probs) {
    float s = 0.f;
    for (int o = 0; o < O; ++o) s += expf(zr[o] - m);
    for (int o = 0; o < O; ++o) probs[static_cast<size_t>(n) * O + o] = expf(zr[o] - m) / s;
  }
}

inline int blocks_for(dlq_ctx* ctx, size_t work, int threads) {
  const size_t b = (work + threads - 1) / threads;
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>(b, static_cast<size_t>(ctx->num_sms) * 8)));
}

// pack a [O, I] byte matrix as the 1x1 conv [Op, Ip, 1, 1]
int pack_fc_bytes(dlq_ctx* ctx, const int8_t* q, int O, int I, int fp8, dlq_conv_weights** out) {
  *out = nullptr;
  DLQ_ARG(ctx, q && O > 0 && I > 0, "null weights or bad dims");
  const int Op = fc_pad_out(O), Ip = fc_pad_in(I);
  DLQ_ARG(ctx, Ip / 128 <= kMaxSteps && Ip <= 16 * 128, "fully-connected input too wide for one pass (max 2048)");
  std::vector<int8_t> qp(static_cast<size_t>(Op) * Ip, 0);
  for (int o = 0; o < O; ++o) std::copy(q + static_cast<size_t>(o) * I, q + static_cast<size_t>(o + 1) * I, qp.begin() + static_cast<size_t>(o) * Ip);
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<dlq_conv_weights> w(new dlq_conv_weights());
  const int rc = pack_conv_weights(ctx, qp.data(), Op, Ip, 1, 1, 1, 1, 0, 0, w.get());
  if (rc != DLQ_OK) {
    if (w->d_img) cudaFree(w->d_img);
    return rc;
  }
  w->fp8 = fp8;
  w->fc_O = O;
  w->fc_I = I;
  w->scale.assign(Op, 1.0f);
  *out = w.release();
  return DLQ_OK;
}

// plan the FC as a 1x1 conv over [N,1,1,Ip]: y (int8 / E4M3, may be null) and / or z (fp32 [N, Op], may be null)
int plan_fc(dlq_ctx* ctx, const dlq_conv_weights* w, const int8_t* x, int N, const float* alpha, const float* beta, int relu,
            int8_t* y, float* z, ConvLaunch* L) {
  Act in, out;
  in.ptr = const_cast<int8_t*>(x); in.N = N; in.H = 1; in.W = 1; in.C = w->IC; in.PR = 0;
  out.ptr = y; out.N = N; out.H = 1; out.W = 1; out.C = w->OC; out.PR = 0;
  const int rc = plan_conv(ctx, w, in, out, alpha, beta, nullptr, 0.f, relu, reinterpret_cast<int32_t*>(z), L);
  if (rc != DLQ_OK) return rc;
  L->p.acc_f32 = z ? 1 : 0;
  return DLQ_OK;
}

struct FcWs { size_t off_x = 0, off_ab = 0, off_z = 0, total = 0; };
FcWs fc_ws_layout(const dlq_conv_weights* w, int N) {
  FcWs L;
  size_t off = 0;
  L.off_x = off; if (w->fc_I != w->IC) off += al1k(static_cast<size_t>(N) * w->IC + 1024);
  L.off_ab = off; if (w->fc_O != w->OC) off += al1k(2 * sizeof(float) * w->OC);
  L.off_z = off; if (w->fc_O != w->OC) off += al1k(static_cast<size_t>(N) * w->OC * sizeof(float));
  L.total = off;
  return L;
}

int fc_forward_bytes(dlq_ctx* ctx, const int8_t* g, const dlq_conv_weights* w, const float* scale, const float* bias, int N,
                     float* logits) {
  DLQ_ARG(ctx, g && w && scale && bias && logits && N >= 0, "null pointer or negative batch");
  DLQ_ARG(ctx, w->fc_O > 0, "weights were not packed by dlq_fc_weights_pack*");
  if (N == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const FcWs ws = fc_ws_layout(w, N);
  int rc = ctx_workspace(ctx, ws.total);
  if (rc != DLQ_OK) return rc;
  uint8_t* base = static_cast<uint8_t*>(ctx->ws);
  const int8_t* x = g;
  if (w->fc_I != w->IC) {
    uint8_t* xp = base + ws.off_x;
    pad_rows_u8_kernel<<<blocks_for(ctx, static_cast<size_t>(N) * w->IC, 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const uint8_t*>(g), N, w->fc_I, w->IC, xp);
    DLQ_CUDA(ctx, cudaGetLastError());
    x = reinterpret_cast<const int8_t*>(xp);
  }
  const float *al = scale, *be = bias;
  float* z = logits;
  if (w->fc_O != w->OC) {
    float* ab = reinterpret_cast<float*>(base + ws.off_ab);
    pad_vec_f32_kernel<<<(w->OC + 255) / 256, 256, 0, ctx->stream>>>(scale, bias, w->fc_O, w->OC, ab, ab + w->OC);
    DLQ_CUDA(ctx, cudaGetLastError());
    al = ab; be = ab + w->OC;
    z = reinterpret_cast<float*>(base + ws.off_z);
  }
  ConvLaunch L;
  rc = plan_fc(ctx, w, x, N, al, be, 0, nullptr, z, &L);
  if (rc != DLQ_OK) return rc;
  rc = launch_conv(ctx, L);
  if (rc != DLQ_OK) return rc;
  if (z != logits) {
    unpad_cols_f32_kernel<<<blocks_for(ctx, static_cast<size_t>(N) * w->fc_O, 256), 256, 0, ctx->stream>>>(z, N, w->fc_O, w->OC, logits);
    DLQ_CUDA(ctx, cudaGetLastError());
  }
  return DLQ_OK;
}

}  // namespace

// ================================================================================================
// MNIST-style two-layer MLP
// ================================================================================================
struct dlq_mlp {
  dlq_ctx* ctx = nullptr;
  int in = 0, hid = 0, out = 0, in_p = 0, hid_p = 0, out_p = 0, max_batch = 0, fp8 = 0;
  float s_x = 1.f, s_h = 1.f;
  dlq_conv_weights* w1 = nullptr;
  dlq_conv_weights* w2 = nullptr;
  float *d_alpha1 = nullptr, *d_beta1 = nullptr, *d_sc2 = nullptr, *d_b2 = nullptr;
  uint8_t* d_xq = nullptr;      // [max_batch, in_p]
  int8_t* d_h = nullptr;        // [max_batch, hid_p]
  float* d_z = nullptr;         // [max_batch, out_p]
  struct Plan { ConvLaunch l1, l2; };
  std::map<int, Plan> plans;
  int last_B = 0;
};

extern "C" {

/* ---- FC weights: [O, I] row-major HOST matrix (the reference's fc.weight layout, R/infer_e2e.cu:206-219) */
int dlq_fc_weights_pack_i8(dlq_ctx* ctx, const int8_t* wq_host, int O, int I, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  return pack_fc_bytes(ctx, wq_host, O, I, 0, out);
}
int dlq_fc_weights_pack_e4m3(dlq_ctx* ctx, const uint8_t* wq_host, int O, int I, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  return pack_fc_bytes(ctx, reinterpret_cast<const int8_t*>(wq_host), O, I, 1, out);
}
int dlq_fc_weights_pack(dlq_ctx* ctx, const float* w_host, int O, int I, int fp8, float* w_scale_host, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, w_host && O > 0 && I > 0, "null weights or bad dims");
  std::vector<int8_t> q;
  std::vector<float> s;
  if (fp8) quantize_rows_e4m3(w_host, O, I, q, s); else quantize_rows(w_host, O, I, q, s);
  const int rc = pack_fc_bytes(ctx, q.data(), O, I, fp8 ? 1 : 0, out);
  if (rc != DLQ_OK) return rc;
  std::copy(s.begin(), s.end(), (*out)->scale.begin());
  if (w_scale_host) std::copy(s.begin(), s.end(), w_scale_host);
  return DLQ_OK;
}
size_t dlq_fc_workspace_bytes(const dlq_conv_weights* w, int N) {
  if (!w || w->fc_O <= 0 || N <= 0) return 0;
  return fc_ws_layout(w, N).total;
}

/* tensor-core FC, the arithmetic of dlq_fc_forward_i8 bit for bit: acc = exact int32 dot product,
 * logits[n, o] = fmaf((float)acc, scale[o], bias[o]).  g: int8 [N, I] device. */
int dlq_fc_forward_i8_tc(dlq_ctx* ctx, const int8_t* g, const dlq_conv_weights* w, const float* scale, const float* bias,
                         int N, float* logits) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, !w || !w->fp8, "weights were packed as E4M3: use dlq_fc_forward_fp8");
  return fc_forward_bytes(ctx, g, w, scale, bias, N, logits);
}
/* E4M3 x E4M3 -> FP32 accumulation on kind::f8f6f4 (order unspecified: tolerance parity, QUANT_SPEC 6-7) */
int dlq_fc_forward_fp8(dlq_ctx* ctx, const uint8_t* g, const dlq_conv_weights* w, const float* scale, const float* bias, int N,
                       float* logits) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, w && w->fp8, "weights were not packed as E4M3");
  return fc_forward_bytes(ctx, reinterpret_cast<const int8_t*>(g), w, scale, bias, N, logits);
}

/* ---- MLP */
void dlq_mlp_destroy(dlq_mlp* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  dlq_conv_weights_free(m->w1);
  dlq_conv_weights_free(m->w2);
  for (void* p : {static_cast<void*>(m->d_alpha1), static_cast<void*>(m->d_beta1), static_cast<void*>(m->d_sc2),
                  static_cast<void*>(m->d_b2), static_cast<void*>(m->d_xq), static_cast<void*>(m->d_h), static_cast<void*>(m->d_z)})
    if (p) cudaFree(p);
  delete m;
}

int dlq_mlp_create(dlq_ctx* ctx, const float* w1_in_hid, const float* b1, const float* w2_hid_out, const float* b2, int in,
                   int hid, int out, float s_x, float s_h, int max_batch, int fp8, dlq_mlp** out_m) {
  if (!ctx || !out_m) return DLQ_ERR_ARG;
  *out_m = nullptr;
  DLQ_ARG(ctx, w1_in_hid && b1 && w2_hid_out && b2 && in > 0 && hid > 0 && out > 0 && max_batch > 0, "null pointer or bad dims");
  DLQ_ARG(ctx, s_x > 0.f && s_h > 0.f && std::isfinite(s_x) && std::isfinite(s_h), "activation scales must be positive");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<dlq_mlp, void (*)(dlq_mlp*)> m(new dlq_mlp(), dlq_mlp_destroy);
  m->ctx = ctx; m->in = in; m->hid = hid; m->out = out; m->max_batch = max_batch; m->fp8 = fp8 ? 1 : 0;
  m->s_x = s_x; m->s_h = s_h;
  // the reference stores W as [in, out] (matmul_a_b(A[m,n], B[n,k]), MN/v3.c:125-134): rows of the FC are its columns
  std::vector<float> w1t(static_cast<size_t>(hid) * in), w2t(static_cast<size_t>(out) * hid);
  for (int i = 0; i < in; ++i)
    for (int h = 0; h < hid; ++h) w1t[static_cast<size_t>(h) * in + i] = w1_in_hid[static_cast<size_t>(i) * hid + h];
  for (int h = 0; h < hid; ++h)
    for (int o = 0; o < out; ++o) w2t[static_cast<size_t>(o) * hid + h] = w2_hid_out[static_cast<size_t>(h) * out + o];
  std::vector<float> sw1(hid), sw2(out);
  int rc = dlq_fc_weights_pack(ctx, w1t.data(), hid, in, m->fp8, sw1.data(), &m->w1);
  if (rc != DLQ_OK) return rc;
  rc = dlq_fc_weights_pack(ctx, w2t.data(), out, hid, m->fp8, sw2.data(), &m->w2);
  if (rc != DLQ_OK) return rc;
  m->in_p = m->w1->IC; m->hid_p = m->w1->OC; m->out_p = m->w2->OC;
  DLQ_ARG(ctx, m->w2->IC == m->hid_p, "hidden width must be 64 or a multiple of 128");
  // layer 1 epilogue, QUANT_SPEC 3 form: h_q = clamp(rne(fmaf(acc, s_x s_w1 / s_h, b1 / s_h)), 0, 127)  (bias + ReLU + requant)
  // layer 2: logit = fmaf(acc, s_h s_w2, b2)                                                    (QUANT_SPEC 5, FC)
  std::vector<float> a1(m->hid_p, 0.f), be1(m->hid_p, 0.f), sc2(m->out_p, 0.f), bb2(m->out_p, 0.f);
  for (int h = 0; h < hid; ++h) {
    a1[h] = static_cast<float>(static_cast<double>(s_x) * static_cast<double>(sw1[h]) / static_cast<double>(s_h));
    be1[h] = static_cast<float>(static_cast<double>(b1[h]) / static_cast<double>(s_h));
  }
  for (int o = 0; o < out; ++o) {
    sc2[o] = static_cast<float>(static_cast<double>(s_h) * static_cast<double>(sw2[o]));
    bb2[o] = b2[o];
  }
  auto up = [&](const std::vector<float>& h, float** d) -> int {
    DLQ_CUDA(ctx, cudaMalloc(d, h.size() * sizeof(float)));
    DLQ_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DLQ_OK;
  };
  if ((rc = up(a1, &m->d_alpha1)) != DLQ_OK || (rc = up(be1, &m->d_beta1)) != DLQ_OK || (rc = up(sc2, &m->d_sc2)) != DLQ_OK ||
      (rc = up(bb2, &m->d_b2)) != DLQ_OK)
    return rc;
  const size_t xb = static_cast<size_t>(max_batch) * m->in_p + 1024, hb = static_cast<size_t>(max_batch) * m->hid_p + 1024;
  DLQ_CUDA(ctx, cudaMalloc(&m->d_xq, xb));
  DLQ_CUDA(ctx, cudaMalloc(&m->d_h, hb));
  DLQ_CUDA(ctx, cudaMalloc(&m->d_z, static_cast<size_t>(max_batch) * m->out_p * sizeof(float)));
  DLQ_CUDA(ctx, cudaMemsetAsync(m->d_xq, 0, xb, ctx->stream));      // the pad columns stay zero for good
  DLQ_CUDA(ctx, cudaMemsetAsync(m->d_h, 0, hb, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_m = m.release();
  return DLQ_OK;
}

/* x: fp32 [B, in] device; logits / probs: fp32 [B, out] device (either may be NULL).  Four launches: quantise, FC1
 * (tcgen05, fused bias + ReLU + requantisation), FC2 (tcgen05, fp32 logits), head (un-pad + softmax). */
int dlq_mlp_forward(dlq_mlp* m, const float* x, int B, float* logits, float* probs) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, x && (logits || probs) && B >= 0 && B <= m->max_batch, "null pointer or batch larger than max_batch");
  if (B == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  auto it = m->plans.find(B);
  if (it == m->plans.end()) {
    dlq_mlp::Plan P;
    int rc = plan_fc(ctx, m->w1, reinterpret_cast<const int8_t*>(m->d_xq), B, m->d_alpha1, m->d_beta1, 1, m->d_h, nullptr, &P.l1);
    if (rc != DLQ_OK) return rc;
    rc = plan_fc(ctx, m->w2, m->d_h, B, m->d_sc2, m->d_b2, 0, nullptr, m->d_z, &P.l2);
    if (rc != DLQ_OK) return rc;
    it = m->plans.emplace(B, P).first;
  }
  const int qb = blocks_for(ctx, static_cast<size_t>(B) * m->in, 256);
  if (m->fp8) quantize_rows_kernel<true><<<qb, 256, 0, ctx->stream>>>(x, B, m->in, m->in_p, inv_scale(m->s_x), m->d_xq);
  else quantize_rows_kernel<false><<<qb, 256, 0, ctx->stream>>>(x, B, m->in, m->in_p, inv_scale(m->s_x), m->d_xq);
  DLQ_CUDA(ctx, cudaGetLastError());
  int rc = launch_conv(ctx, it->second.l1);
  if (rc != DLQ_OK) return rc;
  rc = launch_conv(ctx, it->second.l2);
  if (rc != DLQ_OK) return rc;
  mlp_head_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(m->d_z, B, m->out, m->out_p, logits, probs);
  DLQ_CUDA(ctx, cudaGetLastError());
  m->last_B = B;
  return DLQ_OK;
}

/* checkpoints of the LAST forward: the quantised input [B, in] and hidden activations [B, hid] (bytes: int8 or E4M3) */
int dlq_mlp_checkpoint(dlq_mlp* m, const char* name, uint8_t* out) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, name && out && m->last_B > 0, "null pointer or no forward has run yet");
  const std::string s(name);
  const bool is_x = s == "input", is_h = s == "hidden";
  DLQ_ARG(ctx, is_x || is_h, "unknown checkpoint name (input | hidden)");
  const uint8_t* src = is_x ? m->d_xq : reinterpret_cast<const uint8_t*>(m->d_h);
  const int w = is_x ? m->in : m->hid, wp = is_x ? m->in_p : m->hid_p;
  DLQ_CUDA(ctx, cudaMemcpy2DAsync(out, w, src, wp, w, m->last_B, cudaMemcpyDeviceToDevice, ctx->stream));
  return DLQ_OK;
}
/* HOST: the per-row weight scales the MLP quantised its layers with (layer 1: hid values, layer 2: out values) */
int dlq_mlp_weight_scales(const dlq_mlp* m, int layer, float* scale_host) {
  if (!m || !scale_host || (layer != 1 && layer != 2)) return DLQ_ERR_ARG;
  const dlq_conv_weights* w = layer == 1 ? m->w1 : m->w2;
  std::copy(w->scale.begin(), w->scale.begin() + w->fc_O, scale_host);
  return DLQ_OK;
}

}  // extern "C"
