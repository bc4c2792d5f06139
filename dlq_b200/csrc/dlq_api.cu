// dlq_api.cu — C ABI glue: context, per-layer convolution entry point, the ResNet-18 runner
// (static buffer plan, one launch sequence per batch size) and the batch-sharded multi-GPU driver.
//
// Reference anchors: runtime/infer_e2e.cu:102-136 (conv2d_nchw_im2col_gemm), :156-203
// (basic_block_forward), :206-219 (fc_forward), :254-433 (network wiring, checkpoints).
#include "dlq_internal.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <thread>

using namespace dlq;

// ================================================================================================
// context
// ================================================================================================
extern "C" {

const char* dlq_version(void) { return "dlq_b200 0.1 (sm_100a)"; }

int dlq_create(int device, dlq_ctx** out) {
  if (!out) return DLQ_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return DLQ_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DLQ_ERR_CUDA;
  if (prop.major != 10) {
    fprintf(stderr, "dlq_b200: device %d is sm_%d%d; this library contains sm_100a code only\n", device, prop.major,
            prop.minor);
    return DLQ_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) return DLQ_ERR_CUDA;
  dlq_ctx* c = new dlq_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->no_pdl = dlq_dbg_env("DLQ_DBG_NO_PDL") != nullptr;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return DLQ_ERR_CUDA;
  }
  c->own_stream = true;
  // kernel attributes once per context; an sm_100a image that does not load on this device (sm_101 / sm_103 report
  // major 10 too) fails here, loudly, instead of at the first launch
  if (configure_conv_kernels(c) != DLQ_OK || configure_elementwise_kernels(c) != DLQ_OK ||
      cudaMalloc(&c->small, 256) != cudaSuccess) {
    fprintf(stderr, "dlq_b200: device %d (sm_%d%d) cannot run the sm_100a kernels of this library: %s\n", device, prop.major,
            prop.minor, c->err.c_str());
    cudaStreamDestroy(c->stream);
    delete c;
    return DLQ_ERR_CUDA;
  }
  *out = c;
  return DLQ_OK;
}

void dlq_destroy(dlq_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->small) cudaFree(ctx->small);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

/* workspace of the per-layer NCHW entry points (SURVEY 8b "Ownership") */
size_t dlq_workspace_bytes(const dlq_ctx* ctx) { return ctx ? ctx->ws_bytes : 0; }
int dlq_workspace_reserve(dlq_ctx* ctx, size_t bytes) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->ws_reserved = false;
  const int rc = ctx_workspace(ctx, bytes);
  ctx->ws_reserved = (rc == DLQ_OK);
  return rc;
}

const char* dlq_last_error_string(const dlq_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dlq_sync(dlq_ctx* ctx) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return DLQ_OK;
}
void* dlq_stream(dlq_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
int dlq_set_stream(dlq_ctx* ctx, void* stream) {
  if (!ctx) return DLQ_ERR_ARG;
  if (ctx->own_stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  ctx->stream = static_cast<cudaStream_t>(stream);
  return DLQ_OK;
}

// deterministic synthetic data: SplitMix64 keyed by (seed, FNV-1a(name)); SURVEY §8d
void dlq_synth_fill_f32(float* v, size_t n, uint64_t seed, const char* name, int lo, int hi, int shift) {
  uint64_t h = 0xCBF29CE484222325ULL;
  for (const unsigned char* p = reinterpret_cast<const unsigned char*>(name); *p; ++p) {
    h ^= *p;
    h *= 0x100000001B3ULL;
  }
  uint64_t s = seed * 0xD1342543DE82EF95ULL + h;
  auto next = [&s]() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  };
  (void)next();
  const uint64_t span = static_cast<uint64_t>(hi - lo + 1);
  const float scale = std::ldexp(1.0f, -shift);
  for (size_t i = 0; i < n; ++i) v[i] = static_cast<float>(lo + static_cast<int>(next() % span)) * scale;
}

}  // extern "C"

// ================================================================================================
// weight quantisation helpers (host; QUANT_SPEC §1-§3)
// ================================================================================================
namespace {

inline int8_t quant_host(float t, int lo, int hi) {
  float r = std::nearbyintf(t);   // round-half-to-even under the default rounding mode
  if (!(r >= static_cast<float>(lo))) r = static_cast<float>(lo);
  if (r > static_cast<float>(hi)) r = static_cast<float>(hi);
  return static_cast<int8_t>(static_cast<int>(r));
}

// float -> E4M3 (1-4-3, bias 7, max 448, no infinities), round-to-nearest-even, saturate-to-finite (QUANT_SPEC 6)
uint8_t f32_to_e4m3_host(float f) {
  const uint8_t sign = std::signbit(f) ? 0x80 : 0x00;
  if (std::isnan(f)) return static_cast<uint8_t>(sign | 0x7F);
  const float a = std::fabs(f);
  if (a >= 448.0f) return static_cast<uint8_t>(sign | 0x7E);
  if (a < 0.015625f) {                                  // below the smallest normal 2^-6: multiples of 2^-9
    const int q = static_cast<int>(std::nearbyintf(a * 512.0f));    // 0..8 (8 == smallest normal, same encoding)
    return static_cast<uint8_t>(sign | q);
  }
  int e;
  const float fr = std::frexp(a, &e);                   // a = fr * 2^e, fr in [0.5, 1)
  int ex = e - 1;                                       // a = (2 fr) * 2^ex, 2 fr in [1, 2)
  int m = static_cast<int>(std::nearbyintf((2.0f * fr - 1.0f) * 8.0f));   // 0..8
  if (m == 8) { m = 0; ++ex; }
  if (ex > 8 || (ex == 8 && m == 7)) return static_cast<uint8_t>(sign | 0x7E);
  return static_cast<uint8_t>(sign | ((ex + 7) << 3) | m);
}

}  // namespace

namespace dlq {
void quantize_rows_e4m3(const float* w, int rows, int K, std::vector<int8_t>& q, std::vector<float>& s) {
  q.resize(static_cast<size_t>(rows) * K);
  s.resize(rows);
  for (int r = 0; r < rows; ++r) {
    float am = 0.f;
    for (int k = 0; k < K; ++k) am = std::max(am, std::fabs(w[static_cast<size_t>(r) * K + k]));
    const float sc = am > 0.f ? static_cast<float>(static_cast<double>(am) / 448.0) : 1.0f;
    s[r] = sc;
    const float inv = inv_scale(sc);
    for (int k = 0; k < K; ++k)
      q[static_cast<size_t>(r) * K + k] = static_cast<int8_t>(f32_to_e4m3_host(w[static_cast<size_t>(r) * K + k] * inv));
  }
}

void quantize_rows(const float* w, int rows, int K, std::vector<int8_t>& q, std::vector<float>& s) {
  q.resize(static_cast<size_t>(rows) * K);
  s.resize(rows);
  for (int r = 0; r < rows; ++r) {
    float am = 0.f;
    for (int k = 0; k < K; ++k) am = std::max(am, std::fabs(w[static_cast<size_t>(r) * K + k]));
    const float sc = am > 0.f ? static_cast<float>(static_cast<double>(am) / 127.0) : 1.0f;
    s[r] = sc;
    const float inv = inv_scale(sc);
    for (int k = 0; k < K; ++k) q[static_cast<size_t>(r) * K + k] = quant_host(w[static_cast<size_t>(r) * K + k] * inv, -127, 127);
  }
}

}  // namespace dlq

extern "C" {
void dlq_fold_bn(const float* g, const float* b, const float* m, const float* v, float eps, const float* s_w,
                 float s_x, float s_y, int OC, float* alpha, float* beta) {
  for (int oc = 0; oc < OC; ++oc) {
    const double a = static_cast<double>(g[oc]) / std::sqrt(static_cast<double>(v[oc]) + static_cast<double>(eps));
    alpha[oc] = static_cast<float>(static_cast<double>(s_x) * static_cast<double>(s_w[oc]) * a / static_cast<double>(s_y));
    beta[oc] = static_cast<float>((static_cast<double>(b[oc]) - static_cast<double>(m[oc]) * a) / static_cast<double>(s_y));
  }
}
float dlq_res_mul(float s_r, float s_y) { return static_cast<float>(static_cast<double>(s_r) / static_cast<double>(s_y)); }
}

// ================================================================================================
// per-layer convolution
// ================================================================================================
extern "C" {

int dlq_conv_weights_pack_i8(dlq_ctx* ctx, const int8_t* wq, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                             int pW, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, wq != nullptr, "null weights");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<dlq_conv_weights> w(new dlq_conv_weights());
  const int rc = pack_conv_weights(ctx, wq, OC, IC, kH, kW, sH, sW, pH, pW, w.get());
  if (rc != DLQ_OK) {
    if (w->d_img) cudaFree(w->d_img);
    return rc;
  }
  w->scale.assign(OC, 1.0f);
  *out = w.release();
  return DLQ_OK;
}

int dlq_conv_weights_pack(dlq_ctx* ctx, const float* w_host, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                          int pW, float* w_scale_host, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, w_host != nullptr && OC > 0 && IC > 0 && kH > 0 && kW > 0, "null weights or bad dims");
  std::vector<int8_t> q;
  std::vector<float> s;
  quantize_rows(w_host, OC, IC * kH * kW, q, s);
  const int rc = dlq_conv_weights_pack_i8(ctx, q.data(), OC, IC, kH, kW, sH, sW, pH, pW, out);
  if (rc != DLQ_OK) return rc;
  (*out)->scale = s;
  if (w_scale_host) std::copy(s.begin(), s.end(), w_scale_host);
  return DLQ_OK;
}

/* E4M3 variants: the packed bytes are E4M3, the conv accumulates in FP32 on kind::f8f6f4 */
int dlq_conv_weights_pack_e4m3(dlq_ctx* ctx, const uint8_t* wq, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                               int pW, dlq_conv_weights** out) {
  const int rc = dlq_conv_weights_pack_i8(ctx, reinterpret_cast<const int8_t*>(wq), OC, IC, kH, kW, sH, sW, pH, pW, out);
  if (rc == DLQ_OK) (*out)->fp8 = 1;
  return rc;
}

int dlq_conv_weights_pack_fp8(dlq_ctx* ctx, const float* w_host, int OC, int IC, int kH, int kW, int sH, int sW, int pH,
                              int pW, float* w_scale_host, dlq_conv_weights** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, w_host != nullptr && OC > 0 && IC > 0 && kH > 0 && kW > 0, "null weights or bad dims");
  std::vector<int8_t> q;
  std::vector<float> s;
  quantize_rows_e4m3(w_host, OC, IC * kH * kW, q, s);
  const int rc = dlq_conv_weights_pack_e4m3(ctx, reinterpret_cast<const uint8_t*>(q.data()), OC, IC, kH, kW, sH, sW, pH, pW, out);
  if (rc != DLQ_OK) return rc;
  (*out)->scale = s;
  if (w_scale_host) std::copy(s.begin(), s.end(), w_scale_host);
  return DLQ_OK;
}

uint8_t dlq_f32_to_e4m3(float f) { return f32_to_e4m3_host(f); }

void dlq_conv_weights_free(dlq_conv_weights* w) {
  if (!w) return;
  cudaSetDevice(w->device);
  if (w->d_img) cudaFree(w->d_img);
  delete w;
}

static int conv2d_bytes(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                       const dlq_epilogue* ep, int8_t* y, int32_t* acc_out, int* OH, int* OW);

int dlq_conv2d_i8(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                  const dlq_epilogue* ep, int8_t* y, int32_t* acc_out, int* OH, int* OW) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, !w || !w->fp8, "weights were packed as E4M3: use dlq_conv2d_fp8");
  return conv2d_bytes(ctx, x, N, C, H, W, w, ep, y, acc_out, OH, OW);
}

/* E4M3 activations and weights, FP32 accumulators (acc_out: raw fp32 accumulators, NCHW) */
int dlq_conv2d_fp8(dlq_ctx* ctx, const uint8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                   const dlq_epilogue* ep, uint8_t* y, float* acc_out, int* OH, int* OW) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, w && w->fp8, "weights were not packed as E4M3");
  return conv2d_bytes(ctx, reinterpret_cast<const int8_t*>(x), N, C, H, W, w, ep, reinterpret_cast<int8_t*>(y),
                      reinterpret_cast<int32_t*>(acc_out), OH, OW);
}

// geometry of the four workspace regions dlq_conv2d_* carves (input in the library's layout, output, residual, raw
// accumulators), each rounded up to 1 KB
namespace {
struct ConvWs {
  Act in, out;
  size_t off_in = 0, off_out = 0, off_res = 0, off_acc = 0, total = 0;
};
inline size_t ws_align(size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); }
ConvWs conv_ws_layout(const dlq_conv_weights* w, int N, int H, int W, bool want_y, bool has_res, bool want_acc) {
  ConvWs L;
  int oh, ow;
  conv_out_dims(w, H, W, &oh, &ow);
  L.in.N = N; L.in.PR = conv_required_in_pr(w);
  if (w->kind == CONV_STEM) { L.in.H = H / 2; L.in.W = W / 2 + 3; L.in.C = 32; }
  else { L.in.H = H; L.in.W = W; L.in.C = w->IC; }
  L.out.N = N; L.out.H = oh; L.out.W = ow; L.out.C = w->OC; L.out.PR = 0;
  size_t off = 0;
  L.off_in = off; off += ws_align(L.in.bytes() + 1024);
  L.off_out = off; if (want_y) off += ws_align(L.out.bytes());
  L.off_res = off; if (has_res) off += ws_align(L.out.bytes());
  L.off_acc = off; if (want_acc) off += ws_align(static_cast<size_t>(N) * oh * ow * w->OC * 4);
  L.total = off;
  return L;
}
}  // namespace

/* bytes of workspace dlq_conv2d_i8 / dlq_conv2d_fp8 need for this call shape: pass the maximum over the calls you will
 * make to dlq_workspace_reserve() once and the per-layer entry points never allocate (they grow the workspace on demand,
 * with a synchronisation, only when nothing was reserved) */
size_t dlq_conv2d_workspace_bytes(const dlq_conv_weights* w, int N, int H, int W, int has_residual, int want_acc) {
  if (!w || N <= 0 || H <= 0 || W <= 0) return 0;
  return conv_ws_layout(w, N, H, W, true, has_residual != 0, want_acc != 0).total;
}

static int conv2d_bytes(dlq_ctx* ctx, const int8_t* x, int N, int C, int H, int W, const dlq_conv_weights* w,
                       const dlq_epilogue* ep, int8_t* y, int32_t* acc_out, int* OH, int* OW) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, x && w && N >= 0 && C == w->IC && H > 0 && W > 0, "null pointer or dims do not match the packed weights");
  DLQ_ARG(ctx, (y == nullptr) || (ep != nullptr && ep->alpha && ep->beta),
          "an int8 output needs an epilogue with alpha and beta");
  DLQ_ARG(ctx, y || acc_out, "no output requested");
  int oh, ow;
  conv_out_dims(w, H, W, &oh, &ow);
  DLQ_ARG(ctx, oh > 0 && ow > 0, "empty output");
  if (OH) *OH = oh;
  if (OW) *OW = ow;
  if (N == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  if (w->kind == CONV_STEM) DLQ_ARG(ctx, H % 2 == 0 && W % 2 == 0, "stem needs even H, W");

  const bool has_res = ep && ep->residual;
  const ConvWs ws = conv_ws_layout(w, N, H, W, y != nullptr, has_res, acc_out != nullptr);
  int rc = ctx_workspace(ctx, ws.total);        // no allocation once the workspace is large enough / reserved
  if (rc != DLQ_OK) return rc;
  uint8_t* base = static_cast<uint8_t*>(ctx->ws);
  Act in = ws.in, out = ws.out, res = ws.out;
  in.ptr = reinterpret_cast<int8_t*>(base + ws.off_in);
  DLQ_CUDA(ctx, cudaMemsetAsync(in.ptr, 0, in.bytes() + 1024, ctx->stream));
  rc = (w->kind == CONV_STEM) ? nchw_i8_to_stem_s2d(ctx, x, N, H, W, in) : nchw_to_act_i8(ctx, x, in);
  if (rc != DLQ_OK) return rc;
  if (y) out.ptr = reinterpret_cast<int8_t*>(base + ws.off_out);
  if (has_res) {
    res.ptr = reinterpret_cast<int8_t*>(base + ws.off_res);
    rc = nchw_to_act_i8(ctx, ep->residual, res);
    if (rc != DLQ_OK) return rc;
  }
  int32_t* acc_nhwc = acc_out ? reinterpret_cast<int32_t*>(base + ws.off_acc) : nullptr;
  ConvLaunch L;
  rc = plan_conv(ctx, w, in, out, ep ? ep->alpha : nullptr, ep ? ep->beta : nullptr, has_res ? &res : nullptr,
                 ep ? ep->res_mul : 0.f, ep ? ep->relu : 0, acc_nhwc, &L);
  if (rc != DLQ_OK) return rc;
  rc = launch_conv(ctx, L);
  if (rc != DLQ_OK) return rc;
  if (y) {
    rc = act_to_nchw_i8(ctx, out, y);
    if (rc != DLQ_OK) return rc;
  }
  if (acc_out) {
    rc = nhwc_to_nchw_i32(ctx, acc_nhwc, N, w->OC, oh * ow, acc_out);
    if (rc != DLQ_OK) return rc;
  }
  return DLQ_OK;
}

}  // extern "C"

struct dlq_conv_plan {
  ConvLaunch L;
};

extern "C" {

size_t dlq_act_bytes(int N, int H, int W, int C, int PR) {
  return static_cast<size_t>(PR + static_cast<size_t>(N) * (H + PR)) * W * C;
}
int dlq_conv_required_pad_rows(const dlq_conv_weights* w) { return w ? conv_required_in_pr(w) : 0; }

static Act to_act(const dlq_act* a) {
  Act r;
  r.ptr = a->ptr; r.N = a->N; r.H = a->H; r.W = a->W; r.C = a->C; r.PR = a->PR;
  return r;
}

static int plan_act(dlq_ctx* ctx, const dlq_act* x, const dlq_conv_weights* w, const dlq_epilogue* ep,
                    const dlq_act* residual, const dlq_act* y, int32_t* acc_out, ConvLaunch* L) {
  DLQ_ARG(ctx, x && x->ptr && w && (y || acc_out), "null pointer");
  DLQ_ARG(ctx, !(y && y->ptr) || (ep && ep->alpha && ep->beta), "an int8 output needs an epilogue with alpha and beta");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  Act in = to_act(x), out, res;
  if (y) {
    out = to_act(y);
  } else {
    out.N = in.N; out.C = w->OC; out.PR = 0;
    if (w->kind == CONV_STEM) { out.H = in.H; out.W = in.W - 3; }
    else conv_out_dims(w, in.H, in.W, &out.H, &out.W);
  }
  if (residual) res = to_act(residual);
  return plan_conv(ctx, w, in, out, ep ? ep->alpha : nullptr, ep ? ep->beta : nullptr, residual ? &res : nullptr,
                   ep ? ep->res_mul : 0.f, ep ? ep->relu : 0, acc_out, L);
}

int dlq_conv2d_i8_act(dlq_ctx* ctx, const dlq_act* x, const dlq_conv_weights* w, const dlq_epilogue* ep,
                      const dlq_act* residual, const dlq_act* y, int32_t* acc_out) {
  if (!ctx) return DLQ_ERR_ARG;
  if (x && x->N == 0) return DLQ_OK;
  ConvLaunch L;
  const int rc = plan_act(ctx, x, w, ep, residual, y, acc_out, &L);
  if (rc != DLQ_OK) return rc;
  return launch_conv(ctx, L);
}

int dlq_conv_plan_create(dlq_ctx* ctx, const dlq_act* x, const dlq_conv_weights* w, const dlq_epilogue* ep,
                         const dlq_act* residual, const dlq_act* y, int32_t* acc_out, dlq_conv_plan** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, x && x->N > 0, "empty batch");
  std::unique_ptr<dlq_conv_plan> P(new dlq_conv_plan());
  const int rc = plan_act(ctx, x, w, ep, residual, y, acc_out, &P->L);
  if (rc != DLQ_OK) return rc;
  *out = P.release();
  return DLQ_OK;
}
int dlq_conv_plan_launch(dlq_ctx* ctx, const dlq_conv_plan* plan) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, plan != nullptr, "null plan");
  return launch_conv(ctx, plan->L);
}
void dlq_conv_plan_destroy(dlq_conv_plan* plan) { delete plan; }

int dlq_act_from_nchw_i8(dlq_ctx* ctx, const int8_t* x, const dlq_act* a) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, x && a && a->ptr, "null pointer");
  return nchw_to_act_i8(ctx, x, to_act(a));
}
int dlq_act_to_nchw_i8(dlq_ctx* ctx, const dlq_act* a, int8_t* y) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, y && a && a->ptr, "null pointer");
  return act_to_nchw_i8(ctx, to_act(a), y);
}
int dlq_stem_pack_input_i8(dlq_ctx* ctx, const int8_t* x, int N, int H, int W, const dlq_act* a) {
  if (!ctx) return DLQ_ERR_ARG;
  DLQ_ARG(ctx, x && a && a->ptr, "null pointer");
  return nchw_i8_to_stem_s2d(ctx, x, N, H, W, to_act(a));
}

}  // extern "C"

// ================================================================================================
// ResNet-18 runner
// ================================================================================================
struct dlq_resnet18 {
  dlq_ctx* ctx = nullptr;
  int max_batch = 0;
  dlq_conv_weights* conv[DLQ_NUM_CONVS] = {nullptr};
  float* d_alpha[DLQ_NUM_CONVS] = {nullptr};
  float* d_beta[DLQ_NUM_CONVS] = {nullptr};
  float act_scale[DLQ_NUM_ACTS];
  int fp8 = 0;                // E4M3 activations / weights, FP32 accumulation (else int8 / int32)
  bool has_lut = false;       // dlq_resnet18_set_preprocess was called
  bool fuse_ds = true;        // downsample blocks: the 1x1/s2 shortcut conv rides on conv1's launch (one patch load)
  int8_t* d_fc_w = nullptr;
  float* d_fc_scale = nullptr;
  float* d_fc_bias = nullptr;
  int8_t* d_gap_q = nullptr;
  // activation buffers (row-padded NHWC), geometry for max_batch
  Act a_in, a_stem, a_pool;
  Act a_t1[8], a_ds[8], a_out[8];
  std::vector<void*> allocs;
  dlq_conv_weights* conv_fused[8] = {nullptr};   // downsample blocks: conv1 + shortcut in one weight image (small batches)
  uint8_t* d_lut = nullptr;   // [3][256] uint8 pixel value -> quantised stem input (dlq_resnet18_set_preprocess)
  // Host-buffer entry points: two staging slots (input + logits), allocated on first use, so that the H2D copy of
  // call k+1 overlaps the forward of call k and the D2H copy of call k-1 (dlq_resnet18_submit_host* / _wait).
  struct HostSlot {
    void* d_in = nullptr;       // fp32 NCHW or uint8 HWC batch
    size_t in_bytes = 0;
    float* d_logits = nullptr;
    cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
    bool busy = false;
  };
  HostSlot slot[2];
  int next_slot = 0;
  int fifo[2] = {0, 0};       // outstanding submits, oldest first
  int n_outstanding = 0;
  cudaStream_t d2h_stream = nullptr;
  // dependency flags between the conv launches of a forward (conv_kernel.cuh "dependency flags"): per-unit completion
  // counters of every conv, one buffer sized for max_batch; [n_flags] is the timeout counter
  bool tile_flags = false;    // (measured, batch 256: 0.752 ms/step with flags between separate launches vs 0.740 with grid-level
                              // dependencies - the per-SM hand-over from one launch's CTA to the next one's is what costs, and
                              // flags do not remove it; they are what lets ONE kernel run several layers, see the conv chain)
  unsigned int* d_flags = nullptr;
  int n_flags = 0;
  bool flags_dirty = false;   // a forward failed half-way: clear the counters before the next one
  bool conv_chain = true;     // batches above kFuseMaxBatch: the blocks from chain_first_block on as one persistent launch
  int chain_min_batch = 40;   // smallest batch planned with chains (option "chain_min_batch"); below it: one launch per conv
                              // with tile shapes picked for the problem size (measured per forward, chains vs separate launches:
                              // batch 17: 208 vs 182 us, 32: 235 vs 229, 48: 260 vs 276, 64: 280 vs 299, 96: 335 vs 382)
  int chain_mode0 = 0;        // first launch mode new plans try for the chain (launch_chain; option "chain_launch_mode")
  bool chain_l1 = true;       // layer1's four convs as a chain of their own (option "chain_layer1")
  int chain_start = 7;        // conv index of the main chain's first member: conv1 (1 + 3b) or conv2 (2 + 3b) of a block;
                              // everything behind it belongs to the chain.  7 = layer2.0.conv1: the fifteen convs of layer2..4
  // span stamps: ring of [forward][launch][2] globaltimer values (dlq_resnet18_enable_stamps)
  unsigned long long* d_stamps = nullptr;
  int stamp_ring = 0;
  unsigned long long fwd_count = 0;
  struct Plan {
    int N = 0;
    ConvLaunch L[DLQ_NUM_CONVS];
    bool fused[8] = {false};    // block b's shortcut conv runs inside L[conv1 of b]
    int flag_units = 0;         // dependency counters this plan uses
    // runs of convs that go out as ONE persistent cooperative launch each (conv_chain.cuh): layer1 (single-CTA tile
    // shape) and everything from layer2.0 on (CTA-pair shape)
    struct Chain {
      ChainLaunch launch;
      int n = 0;
      int conv[kMaxChainLayers] = {0};   // conv index of layer l of the chain
      int first_conv = 0;                // the member the forward's launch order reaches first
    };
    std::vector<Chain> chains;
    int chained_convs() const { int c = 0; for (const Chain& ch : chains) c += ch.n; return c; }
  };
  std::map<int, std::unique_ptr<Plan>> plans;
  int last_N = 0;
  cudaGraph_t graph = nullptr;          // dlq_resnet18_graph_capture / _launch
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t copy_stream = nullptr;   // forward_host pipelining
  cudaEvent_t copy_done[64] = {nullptr};
};

namespace {

struct BlockCfg { int ic, oc, stride; bool down; };
const BlockCfg kBlocks[8] = {{64, 64, 1, false},  {64, 64, 1, false},   {64, 128, 2, true},  {128, 128, 1, false},
                             {128, 256, 2, true}, {256, 256, 1, false}, {256, 512, 2, true}, {512, 512, 1, false}};

Act with_n(const Act& a, int N) {
  Act b = a;
  b.N = N;
  return b;
}

int alloc_act(dlq_resnet18* m, Act& a, int N, int H, int W, int C, int PR) {
  a.N = N; a.H = H; a.W = W; a.C = C; a.PR = PR;
  void* p = nullptr;
  const size_t bytes = a.bytes() + 1024;
  DLQ_CUDA(m->ctx, cudaMalloc(&p, bytes));
  DLQ_CUDA(m->ctx, cudaMemsetAsync(p, 0, bytes, m->ctx->stream));
  a.ptr = static_cast<int8_t*>(p);
  m->allocs.push_back(p);
  return DLQ_OK;
}

template <typename T>
int upload(dlq_resnet18* m, const std::vector<T>& h, T** d) {
  void* p = nullptr;
  DLQ_CUDA(m->ctx, cudaMalloc(&p, h.size() * sizeof(T)));
  DLQ_CUDA(m->ctx, cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, m->ctx->stream));
  DLQ_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
  *d = static_cast<T*>(p);
  m->allocs.push_back(p);
  return DLQ_OK;
}

// activation-scale indices (must match the oracle's enum)
inline int act_c1(int b) { return 2 + 3 * b; }
inline int act_ds(int b) { return 3 + 3 * b; }
inline int act_out(int b) { return 4 + 3 * b; }
constexpr int kActInput = 0, kActStem = 1, kActGap = 26;
constexpr int kFuseMaxBatch = 39;      // shortcut convs ride on conv1's launch up to this batch (measured per forward, fused vs
                                       // separate: batch 17: 175 vs 182 us, 32: 219 vs 229; equal at batch 256)

int build_plan_without_chain(dlq_resnet18* m, int N, dlq_resnet18::Plan* P);

// conv index i belongs to the chain that starts at conv `start` (conv1 = 1 + 3b or conv2 = 2 + 3b of a block b): the convs of
// the later blocks, conv2 of the start block, and - when the chain starts with conv1 - the start block's conv1 and shortcut
bool chain_member(int start, int i) {
  if (start >= DLQ_NUM_CONVS || i < 1) return false;
  const int sb = (start - 1) / 3, b = (i - 1) / 3;
  if (b != sb) return b > sb;
  return (start - 1) % 3 == 0 || i == start;
}

int build_plan(dlq_resnet18* m, int N, dlq_resnet18::Plan* P) {
  dlq_ctx* ctx = m->ctx;
  P->N = N;
  const float* S = m->act_scale;
  int rc = plan_conv(ctx, m->conv[0], with_n(m->a_in, N), with_n(m->a_stem, N), m->d_alpha[0], m->d_beta[0], nullptr, 0.f, 1,
                     nullptr, &P->L[0]);
  if (rc != DLQ_OK) return rc;
  Act cur = with_n(m->a_pool, N);
  float s_cur = S[kActStem];
  // the layers of the conv chain stream their weights (one static ring layout for every layer, conv_chain.cuh)
  const bool chains_on = m->conv_chain && N >= m->chain_min_batch;
  const int chain_start = chains_on ? m->chain_start : DLQ_NUM_CONVS;
  const bool l1_chain = chains_on && m->chain_l1 && chain_start > 5;      // (layer1 = convs 1, 2, 4, 5)
  auto in_l1 = [&](int i) { return l1_chain && (i == 1 || i == 2 || i == 4 || i == 5); };
  auto in_chain = [&](int i) { return chain_member(chain_start, i) || in_l1(i); };
  for (int b = 0; b < 8; ++b) {
    const int i1 = 1 + 3 * b, i2 = 2 + 3 * b, id = 3 + 3 * b;
    Act t1 = with_n(m->a_t1[b], N), o = with_n(m->a_out[b], N);
    // Fusing the shortcut into conv1 saves a launch and a second pass over the input, but forces one tile per
    // item (two accumulator blocks per tile in TMEM): it wins at small batches (latency), loses at large ones.
    const int fuse_max = dlq_dbg_env("DLQ_DBG_FUSE_MAX") ? atoi(dlq_dbg_env("DLQ_DBG_FUSE_MAX")) : kFuseMaxBatch;   // (tuning)
    P->fused[b] = kBlocks[b].down && m->fuse_ds && m->conv_fused[b] && N <= fuse_max && !chains_on;
    if (P->fused[b]) {
      SecondConv sc;
      sc.alpha = m->d_alpha[id]; sc.beta = m->d_beta[id]; sc.relu = 0; sc.out = with_n(m->a_ds[b], N);
      rc = plan_conv(ctx, m->conv_fused[b], cur, t1, m->d_alpha[i1], m->d_beta[i1], nullptr, 0.f, 1, nullptr, &P->L[i1], &sc);
    } else {
      rc = plan_conv(ctx, m->conv[i1], cur, t1, m->d_alpha[i1], m->d_beta[i1], nullptr, 0.f, 1, nullptr, &P->L[i1], nullptr, !in_chain(i1));
    }
    if (rc != DLQ_OK) return rc;
    if (kBlocks[b].down) {
      Act ds = with_n(m->a_ds[b], N);
      if (!P->fused[b]) {
        rc = plan_conv(ctx, m->conv[id], cur, ds, m->d_alpha[id], m->d_beta[id], nullptr, 0.f, 0, nullptr, &P->L[id], nullptr, !in_chain(id));
        if (rc != DLQ_OK) return rc;
      }
      rc = plan_conv(ctx, m->conv[i2], t1, o, m->d_alpha[i2], m->d_beta[i2], &ds, dlq_res_mul(S[act_ds(b)], S[act_out(b)]), 1,
                     nullptr, &P->L[i2], nullptr, !in_chain(i2));
    } else {
      rc = plan_conv(ctx, m->conv[i2], t1, o, m->d_alpha[i2], m->d_beta[i2], &cur, dlq_res_mul(s_cur, S[act_out(b)]), 1,
                     nullptr, &P->L[i2], nullptr, !in_chain(i2));
    }
    if (rc != DLQ_OK) return rc;
    cur = o;
    s_cur = S[act_out(b)];
  }
  // ---- dependency flags: every block conv keeps per-unit completion counters; every conv behind the first one waits
  // for the units of its producers instead of for their whole grids.  (The stem, the max-pool behind it and the first
  // block's conv1 keep the grid-level dependency: their producers are not conv launches.)
  int units = 0;
  auto keep = [&](int i) {
    ConvLaunch& L = P->L[i];
    if (!(m->tile_flags || in_chain(i))) return;      // (no consumer waits on this conv tile by tile)
    const int n = conv_flag_units(L);
    if (m->d_flags && units + n <= m->n_flags) conv_set_flags(&L, m->d_flags + units, m->d_flags + m->n_flags);
    units += n;
  };
  for (int b = 0; b < 8; ++b) {
    keep(1 + 3 * b);
    if (kBlocks[b].down && !P->fused[b]) keep(3 + 3 * b);
    keep(2 + 3 * b);
  }
  P->flag_units = units;
  P->chains.clear();
  // (conv_add_dep ignores producers that keep no counters: without "tile_flags" only the convs of the chains depend on
  // each other through flags, and whatever a chain reads from before it is covered by its one grid-level wait)
  if (m->d_flags && units <= m->n_flags) {
    const ConvLaunch* prev_out = nullptr;      // the conv that produced this block's input (null: the max-pool)
    for (int b = 0; b < 8; ++b) {
      ConvLaunch& c1 = P->L[1 + 3 * b];
      ConvLaunch& c2 = P->L[2 + 3 * b];
      const int s = kBlocks[b].stride;
      if (prev_out) conv_add_dep(&c1, *prev_out, s, 1, 1);                   // 3x3: input rows r*s - 1 .. r*s + 1
      const ConvLaunch* ds = nullptr;
      if (kBlocks[b].down) {
        if (P->fused[b]) {
          ds = &c1;                                                          // the shortcut's rows are conv1's units
        } else {
          ConvLaunch& d = P->L[3 + 3 * b];
          if (prev_out) conv_add_dep(&d, *prev_out, s, 0, 0);               // 1x1/s2: input row 2r
          ds = &d;
        }
      }
      conv_add_dep(&c2, c1, 1, 1, 1);
      if (ds) conv_add_dep(&c2, *ds, 1, 0, 0);                              // residual rows r
      else if (prev_out) conv_add_dep(&c2, *prev_out, 1, 0, 0);            // identity skip = the block input
      prev_out = &c2;
    }
    // ---- chains: members in execution order.  The 1x1 shortcut conv goes FIRST in its block: conv2 needs its rows as
    // residual, and behind conv1 they would be the last thing every CTA produces before conv2's first items ask for
    // them (measured: 25 us instead of 17 for conv2)
    auto add_chain = [&](auto member, int first_conv) -> bool {
      dlq_resnet18::Plan::Chain ch;
      const ConvLaunch* layers[kMaxChainLayers];
      for (int b = 0; b < 8; ++b) {
        const int order[3] = {3 + 3 * b, 1 + 3 * b, 2 + 3 * b};
        for (int i : order)
          if (member(i) && (i != 3 + 3 * b || kBlocks[b].down)) {
            if (ch.n == kMaxChainLayers) return false;
            ch.conv[ch.n] = i;
            layers[ch.n++] = &P->L[i];
          }
      }
      ch.first_conv = first_conv;
      // (too many layers, or one without the chain's static configuration: they stay separate launches)
      if (ch.n < 2 || plan_chain(ctx, layers, ch.n, &ch.launch) != DLQ_OK) return false;
      ch.launch.mode = m->chain_mode0;
      P->chains.push_back(ch);
      return true;
    };
    bool ok = true;
    if (l1_chain) ok = add_chain(in_l1, 1) && ok;
    if (chain_start < DLQ_NUM_CONVS) ok = add_chain([&](int i) { return chain_member(chain_start, i); }, chain_start) && ok;
    // a planned member that ended up outside a chain streams its weights for nothing and waits on flags: plan again plainly
    if (!ok && !m->tile_flags) return build_plan_without_chain(m, N, P);
  }
  return DLQ_OK;
}

int build_plan_without_chain(dlq_resnet18* m, int N, dlq_resnet18::Plan* P) {
  const bool keep = m->conv_chain;
  m->conv_chain = false;
  const int rc = build_plan(m, N, P);
  m->conv_chain = keep;
  return rc;
}

}  // namespace

extern "C" {

void dlq_resnet18_destroy(dlq_resnet18* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (auto*& cw : m->conv_fused) { if (cw) { dlq_conv_weights_free(cw); cw = nullptr; } }
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  if (m->graph) cudaGraphDestroy(m->graph);
  for (void* p : m->allocs) cudaFree(p);
  for (auto& c : m->conv) dlq_conv_weights_free(c);
  if (m->copy_stream) {
    cudaStreamSynchronize(m->copy_stream);
    cudaStreamDestroy(m->copy_stream);
    for (auto& e : m->copy_done) if (e) cudaEventDestroy(e);
  }
  if (m->d2h_stream) {
    cudaStreamSynchronize(m->d2h_stream);
    cudaStreamDestroy(m->d2h_stream);
  }
  for (auto& sl : m->slot) {
    if (sl.d_in) cudaFree(sl.d_in);
    if (sl.d_logits) cudaFree(sl.d_logits);
    if (sl.h2d_done) cudaEventDestroy(sl.h2d_done);
    if (sl.compute_done) cudaEventDestroy(sl.compute_done);
    if (sl.d2h_done) cudaEventDestroy(sl.d2h_done);
  }
  if (m->d_stamps) cudaFree(m->d_stamps);
  delete m;
}

int dlq_resnet18_create(dlq_ctx* ctx, const dlq_resnet18_weights* w, int max_batch, dlq_resnet18** out) {
  if (!ctx || !out) return DLQ_ERR_ARG;
  *out = nullptr;
  DLQ_ARG(ctx, w && max_batch > 0, "null weights or non-positive batch");
  for (int i = 0; i < DLQ_NUM_ACTS; ++i) {
    const bool used = !(i >= 2 && i < 26 && (i - 2) % 3 == 1 && !kBlocks[(i - 2) / 3].down);
    DLQ_ARG(ctx, !used || (w->act_scale[i] > 0.f && std::isfinite(w->act_scale[i])), "activation scales must be positive");
  }
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<dlq_resnet18, void (*)(dlq_resnet18*)> m(new dlq_resnet18(), dlq_resnet18_destroy);
  m->ctx = ctx;
  m->max_batch = max_batch;
  std::copy(w->act_scale, w->act_scale + DLQ_NUM_ACTS, m->act_scale);
  m->fp8 = w->fp8 ? 1 : 0;
  m->fuse_ds = dlq_dbg_env("DLQ_NO_FUSE_DS") == nullptr;

  // ---- convs: geometry as wired by runtime/infer_e2e.cu:258-407
  struct CG { int ic, oc, k, s, p; float s_in, s_out; };
  CG geo[DLQ_NUM_CONVS];
  geo[0] = {3, 64, 7, 2, 3, w->act_scale[kActInput], w->act_scale[kActStem]};
  {
    float s_cur = w->act_scale[kActStem];
    for (int b = 0; b < 8; ++b) {
      const BlockCfg& B = kBlocks[b];
      geo[1 + 3 * b] = {B.ic, B.oc, 3, B.stride, 1, s_cur, w->act_scale[act_c1(b)]};
      geo[2 + 3 * b] = {B.oc, B.oc, 3, 1, 1, w->act_scale[act_c1(b)], w->act_scale[act_out(b)]};
      geo[3 + 3 * b] = {B.ic, B.oc, 1, B.stride, 0, s_cur, w->act_scale[act_ds(b)]};
      s_cur = w->act_scale[act_out(b)];
    }
  }
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    const bool present = (i == 0) || ((i - 1) % 3 != 2) || kBlocks[(i - 1) / 3].down;
    if (!present) continue;
    DLQ_ARG(ctx, w->conv_w[i] && w->bn_gamma[i] && w->bn_beta[i] && w->bn_mean[i] && w->bn_var[i], "missing conv/bn weights");
    const CG& g = geo[i];
    std::vector<float> s_w(g.oc);
    int rc = m->fp8 ? dlq_conv_weights_pack_fp8(ctx, w->conv_w[i], g.oc, g.ic, g.k, g.k, g.s, g.s, g.p, g.p, s_w.data(), &m->conv[i])
                    : dlq_conv_weights_pack(ctx, w->conv_w[i], g.oc, g.ic, g.k, g.k, g.s, g.s, g.p, g.p, s_w.data(), &m->conv[i]);
    if (rc != DLQ_OK) return rc;
    const bool is_c1_of_down = m->fuse_ds && i >= 1 && (i - 1) % 3 == 0 && kBlocks[(i - 1) / 3].down;
    if (is_c1_of_down) {
      // small batches additionally get ONE launch for conv1 + the 1x1/s2 shortcut: both weight sets in one image
      // (conv1's K steps + the shortcut's, which reads the centre tap's view of the same patch)
      std::vector<int8_t> q1, q2;
      std::vector<float> s1, s2;
      const CG& gd = geo[i + 2];
      if (m->fp8) { quantize_rows_e4m3(w->conv_w[i], g.oc, g.ic * 9, q1, s1); quantize_rows_e4m3(w->conv_w[i + 2], gd.oc, gd.ic, q2, s2); }
      else { quantize_rows(w->conv_w[i], g.oc, g.ic * 9, q1, s1); quantize_rows(w->conv_w[i + 2], gd.oc, gd.ic, q2, s2); }
      std::unique_ptr<dlq_conv_weights> cw(new dlq_conv_weights());
      rc = pack_conv_weights(ctx, q1.data(), g.oc, g.ic, 3, 3, 2, 2, 1, 1, cw.get(), q2.data());
      if (rc != DLQ_OK) { if (cw->d_img) cudaFree(cw->d_img); return rc; }
      cw->fp8 = m->fp8;
      cw->scale = s1;
      m->conv_fused[(i - 1) / 3] = cw.release();
    }
    // folded constants, QUANT_SPEC §3 (double arithmetic, rounded once)
    std::vector<float> alpha(g.oc), beta(g.oc);
    dlq_fold_bn(w->bn_gamma[i], w->bn_beta[i], w->bn_mean[i], w->bn_var[i], 1e-5f, s_w.data(), g.s_in, g.s_out, g.oc,
                alpha.data(), beta.data());
    rc = upload(m.get(), alpha, &m->d_alpha[i]);
    if (rc != DLQ_OK) return rc;
    rc = upload(m.get(), beta, &m->d_beta[i]);
    if (rc != DLQ_OK) return rc;
  }
  // ---- FC (R/infer_e2e.cu:206-219): per-row int8, scale folded with the GAP activation scale
  DLQ_ARG(ctx, w->fc_w && w->fc_b, "missing fc weights");
  {
    std::vector<int8_t> q;
    std::vector<float> s;
    if (m->fp8) quantize_rows_e4m3(w->fc_w, 1000, 512, q, s); else quantize_rows(w->fc_w, 1000, 512, q, s);
    std::vector<float> sc(1000), bias(w->fc_b, w->fc_b + 1000);
    for (int o = 0; o < 1000; ++o) sc[o] = static_cast<float>(static_cast<double>(w->act_scale[kActGap]) * static_cast<double>(s[o]));
    // transposed image for the fused GAP+FC kernel: [512/16][1024][16 B]
    std::vector<int8_t> qT(static_cast<size_t>(32) * 1024 * 16, 0);
    for (int o = 0; o < 1000; ++o)
      for (int k = 0; k < 512; ++k) qT[(static_cast<size_t>(k / 16) * 1024 + o) * 16 + (k % 16)] = q[static_cast<size_t>(o) * 512 + k];
    int rc = upload(m.get(), qT, &m->d_fc_w);
    if (rc != DLQ_OK) return rc;
    rc = upload(m.get(), sc, &m->d_fc_scale);
    if (rc != DLQ_OK) return rc;
    rc = upload(m.get(), bias, &m->d_fc_bias);
    if (rc != DLQ_OK) return rc;
  }
  // ---- activation buffers
  const int N = max_batch;
  int rc = alloc_act(m.get(), m->a_in, N, 112, 115, 32, 2);
  if (rc == DLQ_OK) rc = alloc_act(m.get(), m->a_stem, N, 112, 112, 64, 0);
  if (rc == DLQ_OK) rc = alloc_act(m.get(), m->a_pool, N, 56, 56, 64, 1);
  int hw = 56;
  for (int b = 0; b < 8 && rc == DLQ_OK; ++b) {
    const BlockCfg& B = kBlocks[b];
    if (B.stride == 2) hw /= 2;
    const bool feeds_s2 = (b + 1 < 8) && kBlocks[b + 1].stride == 2;
    rc = alloc_act(m.get(), m->a_t1[b], N, hw, hw, B.oc, 1);
    if (rc == DLQ_OK && B.down) rc = alloc_act(m.get(), m->a_ds[b], N, hw, hw, B.oc, 0);
    if (rc == DLQ_OK) rc = alloc_act(m.get(), m->a_out[b], N, hw, hw, B.oc, feeds_s2 ? 2 : 1);
    // a block output that only the next block's stride-2 convs read is kept as four parity planes
    if (rc == DLQ_OK && feeds_s2 && !dlq_dbg_env("DLQ_NO_PLANES")) {
      m->a_out[b].planes = 1;
      m->a_out[b].plane_rows = m->a_out[b].rows() / 2;
    }
  }
  if (rc != DLQ_OK) return rc;
  {
    void* p = nullptr;
    DLQ_CUDA(ctx, cudaMalloc(&p, static_cast<size_t>(N) * 512));
    m->allocs.push_back(p);
    m->d_gap_q = static_cast<int8_t*>(p);
    // (the staging buffers of the host-buffer entry points are allocated on their first use)
    DLQ_CUDA(ctx, cudaMalloc(&p, 768));
    m->allocs.push_back(p);
    m->d_lut = static_cast<uint8_t*>(p);
  }
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // plan the full batch now so the first forward does no host planning; a first pass sizes the dependency counters
  std::unique_ptr<dlq_resnet18::Plan> P(new dlq_resnet18::Plan());
  rc = build_plan(m.get(), N, P.get());
  if (rc != DLQ_OK) return rc;
  {
    // (room for every option setting at every batch <= max_batch: with "tile_flags" all twenty convs keep counters, and
    // small batches are planned with down to a quarter of the positions per unit)
    std::unique_ptr<dlq_resnet18::Plan> Pall(new dlq_resnet18::Plan());
    const bool tf = m->tile_flags;
    m->tile_flags = true;
    rc = build_plan(m.get(), N, Pall.get());
    m->tile_flags = tf;
    if (rc != DLQ_OK) return rc;
    m->n_flags = 4 * std::max(P->flag_units, Pall->flag_units) + 1024;
    void* p = nullptr;
    DLQ_CUDA(ctx, cudaMalloc(&p, (static_cast<size_t>(m->n_flags) + 1) * sizeof(unsigned int)));
    m->allocs.push_back(p);
    m->d_flags = static_cast<unsigned int*>(p);
    DLQ_CUDA(ctx, cudaMemsetAsync(p, 0, (static_cast<size_t>(m->n_flags) + 1) * sizeof(unsigned int), ctx->stream));
    DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rc = build_plan(m.get(), N, P.get());
    if (rc != DLQ_OK) return rc;
  }
  m->plans[N] = std::move(P);
  *out = m.release();
  return DLQ_OK;
}

/* profile slots of one forward (an upper bound of the launches: at batches <= 16 the three shortcut convs run inside
 * conv1's launch and their slots stay 0) */
int dlq_resnet18_launches(const dlq_resnet18* m) {
  (void)m;
  return 1 /*quantise+s2d*/ + 20 /*convs*/ + 1 /*max-pool*/ + 1 /*GAP+FC*/;
}
/* kernels one forward of batch N really launches: 6 with the conv chains (N >= 40), 20 with the shortcut convs fused into conv1 (N < 40), else 23 */
int dlq_resnet18_launches_for_batch(const dlq_resnet18* m, int N) {
  if (!m || N <= 0) return 0;
  int fused = 0;
  int chained = 0;       // convs that share the one chain launch (conv_chain.cuh)
  if (N <= m->max_batch) {
    dlq_resnet18* mm = const_cast<dlq_resnet18*>(m);      // (the plan cache is logically mutable)
    auto it = mm->plans.find(N);
    if (it == mm->plans.end()) {
      std::unique_ptr<dlq_resnet18::Plan> P(new dlq_resnet18::Plan());
      if (cudaSetDevice(m->ctx->device) == cudaSuccess && build_plan(mm, N, P.get()) == DLQ_OK) it = mm->plans.emplace(N, std::move(P)).first;
    }
    if (it != mm->plans.end()) {
      chained = it->second->chained_convs() - static_cast<int>(it->second->chains.size());
      for (int b = 0; b < 8; ++b) fused += it->second->fused[b] ? 1 : 0;
    }
  }
  return 23 - fused - chained;
}

// x: fp32 NCHW input, or (x == nullptr) x_u8: uint8 HWC images mapped through m->d_lut
static int forward_impl(dlq_resnet18* m, const float* x, int N, float* logits, cudaEvent_t* ev,
                        const uint8_t* x_u8 = nullptr) {
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, (x || x_u8) && logits && N >= 0 && N <= m->max_batch, "null pointer or batch larger than max_batch");
  if (N == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  auto it = m->plans.find(N);
  if (it == m->plans.end()) {
    std::unique_ptr<dlq_resnet18::Plan> P(new dlq_resnet18::Plan());
    const int rc = build_plan(m, N, P.get());
    if (rc != DLQ_OK) return rc;
    it = m->plans.emplace(N, std::move(P)).first;
  }
  dlq_resnet18::Plan& P = *it->second;
  const float* S = m->act_scale;
  // dependency counters: zero at the start of every forward - the last kernel of the previous one cleared them; only a
  // forward that failed half-way leaves them dirty
  const bool flags_on = (m->tile_flags || !P.chains.empty()) && m->d_flags && P.flag_units <= m->n_flags;
  if (m->flags_dirty && m->d_flags)
    DLQ_CUDA(ctx, cudaMemsetAsync(m->d_flags, 0, static_cast<size_t>(m->n_flags) * sizeof(unsigned int), ctx->stream));
  m->flags_dirty = flags_on;
  int e = 0;
  auto mark = [&]() -> int {
    if (ev) DLQ_CUDA(ctx, cudaEventRecord(ev[e++], ctx->stream));
    return DLQ_OK;
  };
  // span stamps (off unless dlq_resnet18_enable_stamps was called): slot i of this forward's ring entry
  unsigned long long* st_base =
      m->d_stamps ? m->d_stamps + static_cast<size_t>(m->fwd_count % static_cast<unsigned long long>(m->stamp_ring)) * 23 * 2 : nullptr;
  ++m->fwd_count;
  int slot = 0;
  auto stamp = [&]() -> unsigned long long* { unsigned long long* r = st_base ? st_base + 2 * slot : nullptr; ++slot; return r; };
  // conv index i of the plan: its own launch, or - for the layers of the chain - the chain's one launch at its first layer
  auto conv = [&](int i) -> int {
    unsigned long long* sp = stamp();
    for (dlq_resnet18::Plan::Chain& ch : P.chains)
      for (int l = 0; l < ch.n; ++l)
        if (ch.conv[l] == i) {
          if (i != ch.first_conv) return DLQ_OK;      // (launched once, where the forward reaches its first member)
          if (!sp) return launch_chain(ctx, ch.launch);
          std::unique_ptr<ChainLaunch> C(new ChainLaunch(ch.launch));      // (measurement only: per-layer stamp slots)
          for (int k = 0; k < ch.n; ++k) {
            // slot of conv index j in launch order: 1 stem, then per block conv1, [downsample], conv2 behind the max-pool
            const int j = ch.conv[k], b = (j - 1) / 3, r = (j - 1) % 3;
            int pos = 3;
            for (int bb = 0; bb < b; ++bb) pos += kBlocks[bb].down ? 3 : 2;
            pos += r == 0 ? 0 : r == 2 ? 1 : (kBlocks[b].down ? 2 : 1);
            C->cp.layer[k].p.stamps = st_base + 2 * pos;
          }
          const int rc = launch_chain(ctx, *C);
          ch.launch.mode = C->mode;
          ch.launch.warned = C->warned;
          return rc;
        }
    const ConvLaunch& L0 = P.L[i];
    if (!sp) return launch_conv(ctx, L0);
    ConvLaunch L = L0;
    L.p.stamps = sp;
    return launch_conv(ctx, L);
  };
  int rc = mark();
  if (rc != DLQ_OK) return rc;
  rc = x ? quantize_input_s2d(ctx, x, N, 224, 224, inv_scale(S[kActInput]), with_n(m->a_in, N), m->fp8, stamp())
         : preprocess_u8_s2d(ctx, x_u8, N, 224, 224, m->d_lut, with_n(m->a_in, N), stamp());
  if (rc != DLQ_OK) return rc;
  if ((rc = mark()) != DLQ_OK) return rc;
  rc = conv(0);
  if (rc != DLQ_OK) return rc;
  if ((rc = mark()) != DLQ_OK) return rc;
  rc = maxpool_act(ctx, with_n(m->a_stem, N), with_n(m->a_pool, N), stamp());
  if (rc != DLQ_OK) return rc;
  if ((rc = mark()) != DLQ_OK) return rc;
  for (int b = 0; b < 8; ++b) {
    rc = conv(1 + 3 * b);
    if (rc != DLQ_OK) return rc;
    if ((rc = mark()) != DLQ_OK) return rc;
    if (kBlocks[b].down) {
      if (!P.fused[b]) {          // (fused: the shortcut conv ran inside conv1's launch; its profile entry stays 0)
        rc = conv(3 + 3 * b);
        if (rc != DLQ_OK) return rc;
      } else {
        (void)stamp();
      }
      if ((rc = mark()) != DLQ_OK) return rc;
    }
    rc = conv(2 + 3 * b);
    if (rc != DLQ_OK) return rc;
    if ((rc = mark()) != DLQ_OK) return rc;
  }
  const Act last = with_n(m->a_out[7], N);
  const float s_over_hw = static_cast<float>(static_cast<double>(S[act_out(7)]) / static_cast<double>(last.H * last.W));
  unsigned int* zero = flags_on ? m->d_flags : nullptr;
  const int n_zero = flags_on ? P.flag_units : 0;
  unsigned long long* tail_stamp = stamp();
  rc = m->fp8 ? gap_fc_act_e4m3(ctx, last, s_over_hw, inv_scale(S[kActGap]), m->d_fc_w, m->d_fc_scale, m->d_fc_bias, 1000,
                                m->d_gap_q, logits, tail_stamp, zero, n_zero)
              : gap_fc_act(ctx, last, s_over_hw, inv_scale(S[kActGap]), m->d_fc_w, m->d_fc_scale, m->d_fc_bias, 1000, m->d_gap_q,
                           logits, tail_stamp, zero, n_zero);
  if (rc != DLQ_OK) return rc;
  if ((rc = mark()) != DLQ_OK) return rc;
  m->flags_dirty = false;
  m->last_N = N;
  return DLQ_OK;
}

int dlq_resnet18_forward(dlq_resnet18* m, const float* x, int N, float* logits) {
  if (!m) return DLQ_ERR_ARG;
  return forward_impl(m, x, N, logits, nullptr);
}

/* Options (measurement / A-B comparisons).  "tile_flags" = 1 (default): consecutive conv launches depend on each other tile
 * by tile through completion counters, so the tail of one overlaps the head of the next; 0: every launch waits for the
 * whole previous grid (griddepcontrol.wait).  Both give bit-identical results.  Synchronises and drops the cached plans. */
int dlq_resnet18_set_option(dlq_resnet18* m, const char* key, int value) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, key != nullptr, "null key");
  const std::string k(key);
  DLQ_ARG(ctx, k == "tile_flags" || k == "conv_chain" || k == "chain_first_block" || k == "chain_start" || k == "chain_launch_mode" ||
                   k == "chain_layer1" || k == "chain_min_batch",
          "unknown option (tile_flags | conv_chain | chain_first_block | chain_start | chain_layer1 | chain_launch_mode | chain_min_batch)");
  DLQ_ARG(ctx, k != "chain_min_batch" || value >= 1, "chain_min_batch must be >= 1");
  DLQ_ARG(ctx, k != "chain_launch_mode" || (value >= 0 && value <= 2), "chain_launch_mode outside 0..2");
  DLQ_ARG(ctx, k != "chain_first_block" || (value >= 1 && value <= 7), "chain_first_block outside 1..7");
  DLQ_ARG(ctx, k != "chain_start" || (value >= 1 && value < DLQ_NUM_CONVS && (value - 1) % 3 != 2),
          "chain_start must be the index of a block's conv1 (1 + 3b) or conv2 (2 + 3b)");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (k == "tile_flags") m->tile_flags = value != 0;
  else if (k == "conv_chain") m->conv_chain = value != 0;
  else if (k == "chain_launch_mode") m->chain_mode0 = value;
  else if (k == "chain_layer1") m->chain_l1 = value != 0;
  else if (k == "chain_min_batch") m->chain_min_batch = value;
  else if (k == "chain_first_block") m->chain_start = 1 + 3 * value;
  else m->chain_start = value;
  m->plans.clear();
  if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
  if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
  if (m->d_flags) {
    DLQ_CUDA(ctx, cudaMemsetAsync(m->d_flags, 0, (static_cast<size_t>(m->n_flags) + 1) * sizeof(unsigned int), ctx->stream));
    DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  m->flags_dirty = false;
  return DLQ_OK;
}
/* read-only facts about the plan of batch N: "chain_layers" (convs inside chains; 0: none), "chains" (chain launches), "chain_launch_mode" (0: cooperative launch with
 * programmatic stream serialization, 1: cooperative, 2: neither - only when the driver refuses cooperative launches), "chain_cta_pairs", "chain_a_stages", "chain_b_stages", "flag_units" */
int dlq_resnet18_plan_info(dlq_resnet18* m, int N, const char* key, int* value) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, key && value && N > 0 && N <= m->max_batch, "null pointer or batch outside 1..max_batch");
  (void)dlq_resnet18_launches_for_batch(m, N);      // plans the batch if it is not planned yet
  auto it = m->plans.find(N);
  DLQ_ARG(ctx, it != m->plans.end(), "batch could not be planned");
  const dlq_resnet18::Plan& P = *it->second;
  const std::string k(key);
  const dlq_resnet18::Plan::Chain* main_chain = P.chains.empty() ? nullptr : &P.chains.back();
  if (k == "chain_layers") *value = P.chained_convs();
  else if (k == "chains") *value = static_cast<int>(P.chains.size());
  else if (k == "chain_launch_mode") *value = main_chain ? main_chain->launch.mode : -1;
  else if (k == "chain_cta_pairs") *value = main_chain ? static_cast<int>(main_chain->launch.grid.x / 2) : 0;
  else if (k == "chain_a_stages") *value = main_chain ? main_chain->launch.cp.a_stages : 0;
  else if (k == "chain_b_stages") *value = main_chain ? main_chain->launch.cp.b_stages : 0;
  else if (k == "flag_units") *value = P.flag_units;
  else DLQ_ARG(ctx, false, "unknown key");
  return DLQ_OK;
}
/* dependency waits that timed out since creation (a lost producer; must stay 0).  Synchronises. */
int dlq_resnet18_dep_timeouts(dlq_resnet18* m, unsigned int* count) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, count != nullptr, "null pointer");
  *count = 0;
  if (!m->d_flags) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  DLQ_CUDA(ctx, cudaMemcpy(count, m->d_flags + m->n_flags, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  return DLQ_OK;
}

/* Span stamps: with a ring of `ring_forwards` entries enabled, every kernel of a forward records the globaltimer (ns) of
 * its first block entry and last block exit into entry (forward index mod ring) - the launch's span INSIDE a running
 * sequence of forwards, programmatic-dependent-launch overlap included (what a CUDA event between launches cannot
 * see).  One atomic per block at each end; ring_forwards = 0 switches them off.  dlq_resnet18_read_stamps synchronises,
 * copies the ring to HOST memory [ring][dlq_resnet18_launches()][2] (entry, exit; entry == UINT64_MAX: slot not
 * written) and resets it.  Measurement only (bench.py roofline); the reference brackets each launch with its cudaEvent
 * Timer instead (R/utils.hpp:85-92). */
static int reset_stamps(dlq_resnet18* m) {
  dlq_ctx* ctx = m->ctx;
  std::vector<unsigned long long> init(static_cast<size_t>(m->stamp_ring) * 23 * 2);
  for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ULL; init[i + 1] = 0ULL; }
  DLQ_CUDA(ctx, cudaMemcpyAsync(m->d_stamps, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice,
                                ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->fwd_count = 0;
  return DLQ_OK;
}
int dlq_resnet18_enable_stamps(dlq_resnet18* m, int ring_forwards) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, ring_forwards >= 0 && ring_forwards <= 4096, "ring size outside [0, 4096]");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (m->d_stamps) { cudaFree(m->d_stamps); m->d_stamps = nullptr; }
  m->stamp_ring = ring_forwards;
  if (ring_forwards == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaMalloc(&m->d_stamps, static_cast<size_t>(ring_forwards) * 23 * 2 * sizeof(unsigned long long)));
  return reset_stamps(m);
}
int dlq_resnet18_read_stamps(dlq_resnet18* m, unsigned long long* out_host, int* forwards_recorded) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, m->d_stamps && out_host, "stamps are not enabled (dlq_resnet18_enable_stamps) or null pointer");
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  DLQ_CUDA(ctx, cudaMemcpy(out_host, m->d_stamps, static_cast<size_t>(m->stamp_ring) * 23 * 2 * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost));
  if (forwards_recorded)
    *forwards_recorded = static_cast<int>(std::min<unsigned long long>(m->fwd_count, static_cast<unsigned long long>(m->stamp_ring)));
  return reset_stamps(m);
}

// One forward with a CUDA event between consecutive launches (on the context's stream).
// ms[i] = device time of launch i, in the order: quantise+s2d, stem conv, max-pool, then per block
// conv1, [downsample], conv2, and finally GAP+FC  (dlq_resnet18_launches() entries).  Synchronises.
int dlq_resnet18_profile(dlq_resnet18* m, const float* x, int N, float* logits, float* ms) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, ms != nullptr, "null pointer");
  const int L = dlq_resnet18_launches(m);
  std::vector<cudaEvent_t> ev(L + 1);
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  for (auto& e : ev) DLQ_CUDA(ctx, cudaEventCreate(&e));
  int rc = forward_impl(m, x, N, logits, ev.data());
  if (rc == DLQ_OK) {
    DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < L; ++i) DLQ_CUDA(ctx, cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int dlq_resnet18_graph_capture(dlq_resnet18* m, const float* x, int N, float* logits) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  // a plain forward first: builds the launch plans outside the capture
  int rc = forward_impl(m, x, N, logits, nullptr);
  if (rc != DLQ_OK) return rc;
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
  if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
  DLQ_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  rc = forward_impl(m, x, N, logits, nullptr);
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &m->graph);
  if (rc != DLQ_OK || e != cudaSuccess) {          // a failed capture leaves no half-built graph behind
    if (m->graph) { cudaGraphDestroy(m->graph); m->graph = nullptr; }
    cudaGetLastError();
    if (rc != DLQ_OK) return rc;
    DLQ_CUDA(ctx, e);
  }
  const cudaError_t ei = cudaGraphInstantiate(&m->graph_exec, m->graph, 0);
  if (ei != cudaSuccess) {
    cudaGraphDestroy(m->graph);
    m->graph = nullptr;
    m->graph_exec = nullptr;
    DLQ_CUDA(ctx, ei);
  }
  return DLQ_OK;
}

int dlq_resnet18_graph_launch(dlq_resnet18* m) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, m->graph_exec != nullptr, "no graph captured (call dlq_resnet18_graph_capture first)");
  DLQ_CUDA(ctx, cudaGraphLaunch(m->graph_exec, ctx->stream));
  return DLQ_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------------
namespace {
constexpr size_t kImgElems = static_cast<size_t>(3) * 224 * 224;

int host_streams(dlq_resnet18* m) {
  dlq_ctx* ctx = m->ctx;
  if (!m->copy_stream) {
    DLQ_CUDA(ctx, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    DLQ_CUDA(ctx, cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
    for (auto& e : m->copy_done) DLQ_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& sl : m->slot) {
      DLQ_CUDA(ctx, cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming));
      DLQ_CUDA(ctx, cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
      DLQ_CUDA(ctx, cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming));
    }
  }
  return DLQ_OK;
}
// staging of one slot, grown to `in_bytes` (first use; never on the steady-state path)
int slot_buffers(dlq_resnet18* m, dlq_resnet18::HostSlot& sl, size_t in_bytes) {
  dlq_ctx* ctx = m->ctx;
  if (sl.in_bytes < in_bytes) {
    if (sl.d_in) { DLQ_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(sl.d_in); sl.d_in = nullptr; sl.in_bytes = 0; }
    DLQ_CUDA(ctx, cudaMalloc(&sl.d_in, in_bytes));
    sl.in_bytes = in_bytes;
  }
  if (!sl.d_logits) DLQ_CUDA(ctx, cudaMalloc(&sl.d_logits, static_cast<size_t>(m->max_batch) * 1000 * sizeof(float)));
  return DLQ_OK;
}
int wait_oldest(dlq_resnet18* m) {
  dlq_ctx* ctx = m->ctx;
  if (m->n_outstanding == 0) return DLQ_OK;
  dlq_resnet18::HostSlot& sl = m->slot[m->fifo[0]];
  DLQ_CUDA(ctx, cudaEventSynchronize(sl.d2h_done));
  sl.busy = false;
  m->fifo[0] = m->fifo[1];
  --m->n_outstanding;
  return DLQ_OK;
}

// Enqueue one whole-batch forward from host memory into slot staging: H2D on the copy stream (after the slot's previous
// forward has consumed its input), forward on the compute stream, D2H of the logits on a third stream.  Returns
// without synchronising unless both slots are in flight (then it first waits for the older one).
int submit_host(dlq_resnet18* m, const void* x_host, int N, float* logits_host, bool u8) {
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, x_host && logits_host && N >= 0 && N <= m->max_batch, "null pointer or batch larger than max_batch");
  DLQ_ARG(ctx, !u8 || m->has_lut, "call dlq_resnet18_set_preprocess first");
  if (N == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = host_streams(m);
  if (rc != DLQ_OK) return rc;
  if (m->n_outstanding == 2 && (rc = wait_oldest(m)) != DLQ_OK) return rc;
  const int si = m->next_slot;
  dlq_resnet18::HostSlot& sl = m->slot[si];
  const size_t elem = u8 ? 1 : sizeof(float);
  rc = slot_buffers(m, sl, static_cast<size_t>(m->max_batch) * kImgElems * elem);
  if (rc != DLQ_OK) return rc;
  // H2D: the slot's previous forward (two submits ago) must have read its input
  DLQ_CUDA(ctx, cudaStreamWaitEvent(m->copy_stream, sl.compute_done, 0));
  DLQ_CUDA(ctx, cudaMemcpyAsync(sl.d_in, x_host, static_cast<size_t>(N) * kImgElems * elem, cudaMemcpyHostToDevice, m->copy_stream));
  DLQ_CUDA(ctx, cudaEventRecord(sl.h2d_done, m->copy_stream));
  // forward: after its input has landed and the slot's previous logits have left the device
  DLQ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, sl.h2d_done, 0));
  DLQ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, sl.d2h_done, 0));
  rc = u8 ? forward_impl(m, nullptr, N, sl.d_logits, nullptr, static_cast<const uint8_t*>(sl.d_in))
          : forward_impl(m, static_cast<const float*>(sl.d_in), N, sl.d_logits, nullptr);
  if (rc != DLQ_OK) return rc;
  DLQ_CUDA(ctx, cudaEventRecord(sl.compute_done, ctx->stream));
  // D2H of the logits
  DLQ_CUDA(ctx, cudaStreamWaitEvent(m->d2h_stream, sl.compute_done, 0));
  DLQ_CUDA(ctx, cudaMemcpyAsync(logits_host, sl.d_logits, static_cast<size_t>(N) * 1000 * sizeof(float), cudaMemcpyDeviceToHost,
                                m->d2h_stream));
  DLQ_CUDA(ctx, cudaEventRecord(sl.d2h_done, m->d2h_stream));
  sl.busy = true;
  m->fifo[m->n_outstanding++] = si;
  m->next_slot ^= 1;
  return DLQ_OK;
}

// Synchronous form: within ONE call, large batches are pipelined in chunks (the H2D copy of chunk k+1 overlaps the forward
// of chunk k); the fp32 input is 602 KB/image, so the host link, not the GPU, bounds this entry point.
int forward_host_sync(dlq_resnet18* m, const void* x_host, int N, float* logits_host, bool u8) {
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, x_host && logits_host && N >= 0 && N <= m->max_batch, "null pointer or batch larger than max_batch");
  DLQ_ARG(ctx, !u8 || m->has_lut, "call dlq_resnet18_set_preprocess first");
  if (N == 0) return DLQ_OK;
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = host_streams(m);
  if (rc != DLQ_OK) return rc;
  while (m->n_outstanding) if ((rc = wait_oldest(m)) != DLQ_OK) return rc;     // drain submitted work first
  const size_t elem = u8 ? 1 : sizeof(float);
  dlq_resnet18::HostSlot& sl = m->slot[0];
  rc = slot_buffers(m, sl, static_cast<size_t>(m->max_batch) * kImgElems * elem);
  if (rc != DLQ_OK) return rc;
  // chunk size: 64 (fp32) / 128 (uint8: a chunk copies 4x faster) images for N >= 128, never more than 64 chunks
  const int n_ev = static_cast<int>(sizeof(m->copy_done) / sizeof(m->copy_done[0]));
  int chunk = N;
  if (N >= 128) chunk = std::max((u8 && N >= 256) ? 128 : 64, (N + n_ev - 1) / n_ev);
  // the copy stream must not overwrite the staging buffer while a previous call's compute still reads it
  DLQ_CUDA(ctx, cudaEventRecord(sl.compute_done, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamWaitEvent(m->copy_stream, sl.compute_done, 0));
  const uint8_t* src = static_cast<const uint8_t*>(x_host);
  uint8_t* dst = static_cast<uint8_t*>(sl.d_in);
  int k = 0;
  for (int n0 = 0; n0 < N; n0 += chunk, ++k) {
    const int n = std::min(chunk, N - n0);
    DLQ_CUDA(ctx, cudaMemcpyAsync(dst + n0 * kImgElems * elem, src + n0 * kImgElems * elem, n * kImgElems * elem,
                                  cudaMemcpyHostToDevice, m->copy_stream));
    DLQ_CUDA(ctx, cudaEventRecord(m->copy_done[k], m->copy_stream));
  }
  k = 0;
  int last_n = N;
  for (int n0 = 0; n0 < N; n0 += chunk, ++k) {
    const int n = std::min(chunk, N - n0);
    DLQ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->copy_done[k], 0));
    float* lg = sl.d_logits + static_cast<size_t>(n0) * 1000;
    rc = u8 ? forward_impl(m, nullptr, n, lg, nullptr, dst + n0 * kImgElems)
            : forward_impl(m, reinterpret_cast<const float*>(dst) + n0 * kImgElems, n, lg, nullptr);
    if (rc != DLQ_OK) return rc;
    last_n = n;
  }
  DLQ_CUDA(ctx, cudaMemcpyAsync(logits_host, sl.d_logits, static_cast<size_t>(N) * 1000 * sizeof(float),
                                cudaMemcpyDeviceToHost, ctx->stream));
  DLQ_CUDA(ctx, cudaEventRecord(sl.compute_done, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->last_N = last_n;   // checkpoints refer to the last chunk
  return DLQ_OK;
}
}  // namespace

int dlq_resnet18_forward_host(dlq_resnet18* m, const float* x_host, int N, float* logits_host) {
  if (!m) return DLQ_ERR_ARG;
  return forward_host_sync(m, x_host, N, logits_host, false);
}

/* Pipelined form of the host-buffer entry points: submit enqueues H2D -> forward (the WHOLE batch in one pass) -> D2H and
 * returns; up to two submits may be in flight, so that the copy of batch k+1 overlaps the forward of batch k and the
 * logits copy of batch k-1.  dlq_resnet18_wait blocks until the OLDEST outstanding submit's logits are in host memory
 * (returns immediately when nothing is outstanding).  x_host must stay valid until the matching wait; pinned host
 * memory gives full speed. */
int dlq_resnet18_submit_host(dlq_resnet18* m, const float* x_host, int N, float* logits_host) {
  if (!m) return DLQ_ERR_ARG;
  return submit_host(m, x_host, N, logits_host, false);
}
int dlq_resnet18_submit_host_u8(dlq_resnet18* m, const uint8_t* x_hwc_host, int N, float* logits_host) {
  if (!m) return DLQ_ERR_ARG;
  return submit_host(m, x_hwc_host, N, logits_host, true);
}
int dlq_resnet18_wait(dlq_resnet18* m) {
  if (!m) return DLQ_ERR_ARG;
  DLQ_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
  return wait_oldest(m);
}

/* uint8 input path.  The reference normalises on the host in Python (tools/preprocess_to_bin.py:24-33:
 * x = u8 / 255, (x - mean) / std in float32) and feeds fp32; here the normalisation and the input quantisation are a
 * 3 x 256 byte table applied on the device, built with exactly that fp32 arithmetic, so the logits are bit-identical
 * to dlq_resnet18_forward on the fp32 tensor the reference's preprocessing would have produced. */
int dlq_resnet18_set_preprocess(dlq_resnet18* m, const float* mean3, const float* std3) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, mean3 && std3 && std3[0] > 0.f && std3[1] > 0.f && std3[2] > 0.f, "null pointer or non-positive std");
  std::vector<uint8_t> lut(768);
  const float inv = inv_scale(m->act_scale[kActInput]);
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) {
      volatile float v = static_cast<float>(u) / 255.0f;      // (volatile: every step rounded to binary32, as numpy does)
      v = v - mean3[c];
      v = v / std3[c];
      v = v * inv;
      lut[c * 256 + u] = m->fp8 ? f32_to_e4m3_host(v) : static_cast<uint8_t>(quant_host(v, -128, 127));
    }
  DLQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DLQ_CUDA(ctx, cudaMemcpyAsync(m->d_lut, lut.data(), 768, cudaMemcpyHostToDevice, ctx->stream));
  DLQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  m->has_lut = true;
  return DLQ_OK;
}

/* x_hwc: uint8 [N,224,224,3] device; logits: fp32 [N,1000] device */
int dlq_resnet18_forward_u8(dlq_resnet18* m, const uint8_t* x_hwc, int N, float* logits) {
  if (!m) return DLQ_ERR_ARG;
  DLQ_ARG(m->ctx, m->has_lut, "call dlq_resnet18_set_preprocess first");
  return forward_impl(m, nullptr, N, logits, nullptr, x_hwc);
}

/* same with HOST buffers: H2D of the uint8 images (150 KB/image instead of 602 KB), forward, D2H of the logits;
 * large batches are pipelined in chunks like dlq_resnet18_forward_host */
int dlq_resnet18_forward_host_u8(dlq_resnet18* m, const uint8_t* x_hwc_host, int N, float* logits_host) {
  if (!m) return DLQ_ERR_ARG;
  return forward_host_sync(m, x_hwc_host, N, logits_host, true);
}

int dlq_resnet18_checkpoint(dlq_resnet18* m, const char* name, int8_t* out) {
  if (!m) return DLQ_ERR_ARG;
  dlq_ctx* ctx = m->ctx;
  DLQ_ARG(ctx, name && out, "null pointer");
  DLQ_ARG(ctx, m->last_N > 0, "no forward has run yet");
  const int N = m->last_N;
  const std::string s(name);
  if (s == "gap") {
    DLQ_CUDA(ctx, cudaMemcpyAsync(out, m->d_gap_q, static_cast<size_t>(N) * 512, cudaMemcpyDeviceToDevice, ctx->stream));
    return DLQ_OK;
  }
  const Act* a = nullptr;
  if (s == "stem_pool") a = &m->a_pool;
  else if (s == "stem") a = &m->a_stem;
  else if (s == "layer1") a = &m->a_out[1];
  else if (s == "layer2") a = &m->a_out[3];
  else if (s == "layer3") a = &m->a_out[5];
  else if (s == "layer4") a = &m->a_out[7];
  DLQ_ARG(ctx, a != nullptr, "unknown checkpoint name");
  return act_to_nchw_i8(ctx, with_n(*a, N), out);
}

}  // extern "C"

// ================================================================================================
// batch-sharded multi-GPU driver: one context + model replica + PERSISTENT host worker thread per device.  Images are
// split contiguously; every worker drives its device through the pipelined host entry points and writes its logits
// straight into the caller's host array - no collective.  The same device may be listed more than once
// (independent replicas on separate streams).
// ================================================================================================
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>

struct dlq_multi {
  struct Worker {
    dlq_ctx* ctx = nullptr;
    dlq_resnet18* model = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<std::function<int()>> jobs;
    int pending = 0;        // queued + running jobs
    int rc = DLQ_OK;        // first failure since the last wait
    bool stop = false;
  };
  std::vector<std::unique_ptr<Worker>> w;
  int max_per_dev = 0;
  std::string err;
};

namespace {
void worker_loop(dlq_multi::Worker* w) {
  cudaSetDevice(w->ctx->device);
  for (;;) {
    std::function<int()> job;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv_job.wait(lk, [&] { return w->stop || !w->jobs.empty(); });
      if (w->jobs.empty()) return;      // stop requested and nothing left to run
      job = std::move(w->jobs.front());
      w->jobs.pop_front();
    }
    const int rc = job();
    {
      std::lock_guard<std::mutex> lk(w->mu);
      if (rc != DLQ_OK && w->rc == DLQ_OK) w->rc = rc;
      --w->pending;
    }
    w->cv_done.notify_all();
  }
}
void post(dlq_multi::Worker* w, std::function<int()> job) {
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->jobs.push_back(std::move(job));
    ++w->pending;
  }
  w->cv_job.notify_one();
}
// wait until every worker is idle; returns the first failure (and clears it)
int drain(dlq_multi* m) {
  int first = DLQ_OK;
  for (size_t g = 0; g < m->w.size(); ++g) {
    dlq_multi::Worker* w = m->w[g].get();
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv_done.wait(lk, [&] { return w->pending == 0; });
    if (w->rc != DLQ_OK && first == DLQ_OK) {
      first = w->rc;
      m->err = std::string("device slot ") + std::to_string(g) + ": " + dlq_last_error_string(w->ctx);
    }
    w->rc = DLQ_OK;
  }
  return first;
}
// contiguous split of N images over G workers: worker g takes [g * ceil(N/G), ...)
inline void shard_range(int N, int G, int g, int* n0, int* n1) {
  const int per = (N + G - 1) / G;
  *n0 = std::min(N, g * per);
  *n1 = std::min(N, (g + 1) * per);
}
int multi_submit(dlq_multi* m, const void* x_host, int N, float* logits_host, bool u8) {
  if (!m || !x_host || !logits_host || N < 0) return DLQ_ERR_ARG;
  const int G = static_cast<int>(m->w.size());
  if ((N + G - 1) / G > m->max_per_dev) {
    m->err = "batch exceeds n_devices * max_batch_per_device";
    return DLQ_ERR_ARG;
  }
  const size_t elem = u8 ? 1 : sizeof(float);
  for (int g = 0; g < G; ++g) {
    int n0, n1;
    shard_range(N, G, g, &n0, &n1);
    if (n1 <= n0) continue;
    dlq_multi::Worker* w = m->w[g].get();
    const uint8_t* xs = static_cast<const uint8_t*>(x_host) + static_cast<size_t>(n0) * 3 * 224 * 224 * elem;
    float* ls = logits_host + static_cast<size_t>(n0) * 1000;
    const int n = n1 - n0;
    post(w, [w, xs, ls, n, u8]() {
      return u8 ? dlq_resnet18_submit_host_u8(w->model, xs, n, ls)
                : dlq_resnet18_submit_host(w->model, reinterpret_cast<const float*>(xs), n, ls);
    });
  }
  return DLQ_OK;
}
}  // namespace

extern "C" {

void dlq_multi_destroy(dlq_multi* m) {
  if (!m) return;
  for (auto& up : m->w) {
    dlq_multi::Worker* w = up.get();
    if (w->th.joinable()) {
      { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
      w->cv_job.notify_all();
      w->th.join();
    }
    if (w->model) dlq_resnet18_destroy(w->model);
    if (w->ctx) dlq_destroy(w->ctx);
  }
  delete m;
}

int dlq_multi_create(const int* devices, int n_devices, const dlq_resnet18_weights* w, int max_batch_per_device,
                     dlq_multi** out) {
  if (!out || !devices || n_devices <= 0 || !w || max_batch_per_device <= 0) return DLQ_ERR_ARG;
  *out = nullptr;
  dlq_multi* m = new dlq_multi();
  m->max_per_dev = max_batch_per_device;
  for (int i = 0; i < n_devices; ++i) {
    m->w.emplace_back(new dlq_multi::Worker());
    dlq_multi::Worker* wk = m->w.back().get();
    int rc = dlq_create(devices[i], &wk->ctx);
    if (rc == DLQ_OK) {
      rc = dlq_resnet18_create(wk->ctx, w, max_batch_per_device, &wk->model);
      if (rc != DLQ_OK) fprintf(stderr, "dlq_multi_create: device %d: %s\n", devices[i], dlq_last_error_string(wk->ctx));
    }
    if (rc != DLQ_OK) {
      dlq_multi_destroy(m);
      return rc;
    }
    wk->th = std::thread(worker_loop, wk);
  }
  *out = m;
  return DLQ_OK;
}

const char* dlq_multi_last_error_string(const dlq_multi* m) { return m ? m->err.c_str() : "null"; }
int dlq_multi_n_devices(const dlq_multi* m) { return m ? static_cast<int>(m->w.size()) : 0; }

/* uint8 preprocessing table on every replica (dlq_resnet18_set_preprocess) */
int dlq_multi_set_preprocess(dlq_multi* m, const float* mean3, const float* std3) {
  if (!m || !mean3 || !std3) return DLQ_ERR_ARG;
  for (auto& up : m->w) {
    dlq_multi::Worker* w = up.get();
    const float mean[3] = {mean3[0], mean3[1], mean3[2]}, sd[3] = {std3[0], std3[1], std3[2]};
    post(w, [w, mean, sd]() { return dlq_resnet18_set_preprocess(w->model, mean, sd); });
  }
  return drain(m);
}

/* asynchronous: the batch is split, every worker enqueues H2D -> forward -> D2H for its share and returns to its queue;
 * up to two batches per device are in flight (dlq_resnet18_submit_host).  dlq_multi_wait blocks until everything
 * submitted so far has its logits in host memory. */
int dlq_multi_submit_host(dlq_multi* m, const float* x_host, int N, float* logits_host) {
  return multi_submit(m, x_host, N, logits_host, false);
}
int dlq_multi_submit_host_u8(dlq_multi* m, const uint8_t* x_hwc_host, int N, float* logits_host) {
  return multi_submit(m, x_hwc_host, N, logits_host, true);
}
int dlq_multi_wait(dlq_multi* m) {
  if (!m) return DLQ_ERR_ARG;
  for (auto& up : m->w) {
    dlq_multi::Worker* w = up.get();
    post(w, [w]() {
      int rc = DLQ_OK;
      while (rc == DLQ_OK && w->model->n_outstanding) rc = dlq_resnet18_wait(w->model);
      return rc;
    });
  }
  return drain(m);
}
int dlq_multi_forward_host(dlq_multi* m, const float* x_host, int N, float* logits_host) {
  const int rc = multi_submit(m, x_host, N, logits_host, false);
  return rc != DLQ_OK ? rc : dlq_multi_wait(m);
}
int dlq_multi_forward_host_u8(dlq_multi* m, const uint8_t* x_hwc_host, int N, float* logits_host) {
  const int rc = multi_submit(m, x_hwc_host, N, logits_host, true);
  return rc != DLQ_OK ? rc : dlq_multi_wait(m);
}

/* device-resident form: x_dev[g] / logits_dev[g] are DEVICE pointers on the g-th listed device holding n_per_dev[g] images
 * (fp32 NCHW) and room for n_per_dev[g] x 1000 logits; every replica runs its forward and is synchronised.  This is the
 * driver's compute-only number (no host link in the way). */
int dlq_multi_forward_device(dlq_multi* m, const float* const* x_dev, const int* n_per_dev, float* const* logits_dev) {
  if (!m || !x_dev || !n_per_dev || !logits_dev) return DLQ_ERR_ARG;
  for (size_t g = 0; g < m->w.size(); ++g) {
    if (n_per_dev[g] <= 0) continue;
    dlq_multi::Worker* w = m->w[g].get();
    const float* x = x_dev[g];
    float* l = logits_dev[g];
    const int n = n_per_dev[g];
    post(w, [w, x, l, n]() {
      const int rc = dlq_resnet18_forward(w->model, x, n, l);
      return rc != DLQ_OK ? rc : dlq_sync(w->ctx);
    });
  }
  return drain(m);
}

}  // extern "C"
