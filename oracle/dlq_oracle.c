/* dlq_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see dlq_oracle.h).
 *
 * Reference citations are relative to /root/reference/CUDA/resnet18-kernel-lab/cpp/fp32
 * ("K/" = kernels/, "R/" = runtime/) and /root/reference/CUDA/MNIST_on_GPU ("MN/").
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (see oracle/Makefile)
 * -ffp-contract=off: every rounding below is explicit (fmaf where the GPU fuses, separate
 * mul/add where it does not).
 *
 * INT8 half: PARITY UNPINNED by the reference (it has no quantised code); follows spec/QUANT_SPEC.md.
 */
#include "dlq_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ============================================================ synthetic data */
uint64_t orc_splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
uint64_t orc_name_hash(const char* name) {
  uint64_t h = 0xCBF29CE484222325ULL;
  for (const unsigned char* p = (const unsigned char*)name; *p; ++p) {
    h ^= *p;
    h *= 0x100000001B3ULL;
  }
  return h;
}
static uint64_t stream_state(uint64_t seed, const char* name) {
  uint64_t s = seed * 0xD1342543DE82EF95ULL + orc_name_hash(name);
  (void)orc_splitmix64(&s);
  return s;
}
void orc_fill_lattice_f32(float* v, size_t n, uint64_t seed, const char* name, int lo, int hi, int shift) {
  uint64_t s = stream_state(seed, name);
  const uint64_t span = (uint64_t)(hi - lo + 1);
  const float scale = ldexpf(1.0f, -shift);
  for (size_t i = 0; i < n; ++i) {
    int q = lo + (int)(orc_splitmix64(&s) % span);
    v[i] = (float)q * scale;
  }
}
void orc_fill_lattice_i8(int8_t* v, size_t n, uint64_t seed, const char* name, int lo, int hi) {
  uint64_t s = stream_state(seed, name);
  const uint64_t span = (uint64_t)(hi - lo + 1);
  for (size_t i = 0; i < n; ++i) v[i] = (int8_t)(lo + (int)(orc_splitmix64(&s) % span));
}

/* ============================================================ FP32 reference semantics */

/* im2col + tiled SGEMM fused into one loop nest.
 * Geometry: K/im2col.cu:15-16 (OH/OW), :37-38 (ih = oh*s - p + kh), :41-46 (OOB -> 0).
 * K order r = c*kH*kW + kh*kW + kw: K/im2col.cu:50, weight order R/infer_e2e.cu:116-126.
 * Accumulation: sequential `acc += a*b` over K (K/sgemm_tiled.cu:22-40), which nvcc contracts
 * to an FMA chain; padded taps are multiplied in as zeros exactly as the col buffer holds them.
 * Batch: the reference kernel ignores N (K/im2col.cu:12); images are independent, so n loops. */
void orc_conv2d_f32(const float* x, int N, int C, int H, int W, const float* w, int OC, int kH, int kW,
                    int sH, int sW, int pH, int pW, float* y) {
  const int OH = (H + 2 * pH - kH) / sH + 1, OW = (W + 2 * pW - kW) / sW + 1;
  const int K = C * kH * kW;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int oc = 0; oc < OC; ++oc) {
      const float* xn = x + (size_t)n * C * H * W;
      const float* wr = w + (size_t)oc * K;
      float* yo = y + ((size_t)n * OC + oc) * OH * OW;
      for (int oh = 0; oh < OH; ++oh)
        for (int ow = 0; ow < OW; ++ow) {
          float acc = 0.f;
          for (int c = 0; c < C; ++c)
            for (int kh = 0; kh < kH; ++kh) {
              const int ih = oh * sH - pH + kh;
              for (int kw = 0; kw < kW; ++kw) {
                const int iw = ow * sW - pW + kw;
                float v = 0.f;
                if (ih >= 0 && iw >= 0 && ih < H && iw < W) v = xn[(size_t)c * H * W + (size_t)ih * W + iw];
                acc = fmaf(wr[c * kH * kW + kh * kW + kw], v, acc);
              }
            }
          yo[oh * OW + ow] = acc;
        }
    }
}

/* K/sgemm_tiled.cu:6-46 — C[M,N] = A[M,K] * B[K,N], row-major, sequential-K FMA chain. */
void orc_sgemm_f32(const float* A, const float* B, float* C, int M, int N, int K) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(A[(size_t)i * K + k], B[(size_t)k * N + j], acc);
      C[(size_t)i * N + j] = acc;
    }
}

/* K/bn_inference.cu:22-27: y = (x - m[c]) / sqrtf(v[c] + eps); x = g[c]*y + b[c].
 * true division, then multiply-add (contracted to FMA by nvcc); c = (idx/(OH*OW)) % C. */
void orc_bn_inference_f32(float* x, const float* g, const float* b, const float* m, const float* v, float eps,
                          int N, int C, int OH, int OW) {
  const size_t total = (size_t)N * C * OH * OW;
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < total; ++idx) {
    const int c = (int)((idx / ((size_t)OH * OW)) % C);
    const float yv = (x[idx] - m[c]) / sqrtf(v[c] + eps);
    x[idx] = fmaf(g[c], yv, b[c]);
  }
}

/* K/relu.cu:9 — `if (x < 0) x = 0` (NaN and -0.0 pass through). */
void orc_relu_f32(float* x, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (x[i] < 0.f) x[i] = 0.f;
}
/* K/add.cu:7 */
void orc_add_f32(float* y, const float* x, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] += x[i];
}

/* K/maxpool2d.cu:14-40: OH=(H+2-3)/2+1, window origin oh*2-1, OOB taps skipped, init -FLT_MAX. */
void orc_maxpool3x3s2p1_f32(const float* x, int N, int C, int H, int W, float* y) {
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const float* xp = x + (size_t)nc * H * W;
    float* yp = y + (size_t)nc * OH * OW;
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) {
        float vmax = -FLT_MAX;
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = oh * 2 - 1 + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < 3; ++kw) {
            const int iw = ow * 2 - 1 + kw;
            if (iw < 0 || iw >= W) continue;
            const float v = xp[ih * W + iw];
            vmax = v > vmax ? v : vmax;
          }
        }
        yp[oh * OW + ow] = vmax;
      }
  }
}

/* K/gap_global.cu:10-32 (== R/infer_e2e.cu:37-61): 256 threads, thread t sums x[t], x[t+256], ...;
 * shared-memory tree 128 -> 2, then smem[0]+smem[1], divided by (float)HW. */
void orc_gap_f32(const float* x, int N, int C, int H, int W, float* y) {
  const int HW = H * W;
  for (int nc = 0; nc < N * C; ++nc) {
    float smem[256];
    for (int t = 0; t < 256; ++t) {
      float sum = 0.f;
      for (int i = t; i < HW; i += 256) sum += x[(size_t)nc * HW + i];
      smem[t] = sum;
    }
    for (int s = 128; s > 1; s >>= 1)
      for (int t = 0; t < s; ++t) smem[t] += smem[t + s];
    y[nc] = (smem[0] + smem[1]) / (float)HW;
  }
}

/* R/infer_e2e.cu:206-219: sgemm_tiled(M=O, N=1, K=I) then host `out[o] += B[o]`. */
void orc_fc_f32(const float* gap, const float* W, const float* B, float* out, int N, int O, int I) {
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < O; ++o) {
      float acc = 0.f;
      for (int k = 0; k < I; ++k) acc = fmaf(W[(size_t)o * I + k], gap[(size_t)n * I + k], acc);
      out[(size_t)n * O + o] = acc + B[o];
    }
}

/* K/softmax.cu:6-47.  The reference uses __expf (fast-math, not CPU reproducible) and a 256-thread
 * tree; this restatement uses expf and the same strided partial sums -> compare with tolerance. */
void orc_softmax_f32(const float* x, int K, float* y) {
  float tmp[256];
  for (int t = 0; t < 256; ++t) {
    float m = -INFINITY;
    for (int i = t; i < K; i += 256) m = fmaxf(m, x[i]);
    tmp[t] = m;
  }
  for (int s = 128; s > 0; s >>= 1)
    for (int t = 0; t < s; ++t) tmp[t] = fmaxf(tmp[t], tmp[t + s]);
  const float smax = tmp[0];
  for (int t = 0; t < 256; ++t) {
    float sum = 0.f;
    for (int i = t; i < K; i += 256) sum += expf(x[i] - smax);
    tmp[t] = sum;
  }
  for (int s = 128; s > 0; s >>= 1)
    for (int t = 0; t < s; ++t) tmp[t] += tmp[t + s];
  const float ssum = tmp[0];
  for (int i = 0; i < K; ++i) y[i] = expf(x[i] - smax) / ssum;
}

static float absmax_f32(const float* x, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) {
    const float a = fabsf(x[i]);
    if (a > m) m = a;
  }
  return m;
}
static void track(float* absmax, int idx, const float* x, size_t n) {
  if (!absmax) return;
  const float m = absmax_f32(x, n);
  if (m > absmax[idx]) absmax[idx] = m;
}

static float* conv_bn_f32(const orc_convbn* p, const float* x, int N, int H, int W, int* OH, int* OW) {
  *OH = (H + 2 * p->pad - p->k) / p->stride + 1;
  *OW = (W + 2 * p->pad - p->k) / p->stride + 1;
  float* y = (float*)malloc((size_t)N * p->oc * (*OH) * (*OW) * sizeof(float));
  orc_conv2d_f32(x, N, p->ic, H, W, p->w, p->oc, p->k, p->k, p->stride, p->stride, p->pad, p->pad, y);
  orc_bn_inference_f32(y, p->gamma, p->beta, p->mean, p->var, 1e-5f, N, p->oc, *OH, *OW); /* eps R/infer_e2e.cu:89 */
  return y;
}

/* Network wiring: R/infer_e2e.cu:254-433; BasicBlock: R/infer_e2e.cu:156-203. */
void orc_resnet18_f32_forward(const orc_resnet18_f32* m, const float* x, int N, float* logits,
                              orc_checkpoints_f32* ck) {
  float* am = ck ? ck->absmax : NULL;
  int H = 224, W = 224, OH, OW;
  track(am, ORC_ACT_INPUT, x, (size_t)N * 3 * H * W);
  float* y0 = conv_bn_f32(&m->convs[0], x, N, H, W, &OH, &OW); /* stem :259-279 */
  orc_relu_f32(y0, (size_t)N * 64 * OH * OW);
  track(am, ORC_ACT_STEM, y0, (size_t)N * 64 * OH * OW);
  const int PH = (OH + 2 - 3) / 2 + 1, PW = (OW + 2 - 3) / 2 + 1; /* :283-293 */
  float* cur = (float*)malloc((size_t)N * 64 * PH * PW * sizeof(float));
  orc_maxpool3x3s2p1_f32(y0, N, 64, OH, OW, cur);
  free(y0);
  H = PH;
  W = PW;
  if (ck && ck->stem_pool) memcpy(ck->stem_pool, cur, (size_t)N * 64 * H * W * sizeof(float));
  int C = 64;
  for (int b = 0; b < ORC_NUM_BLOCKS; ++b) {
    const orc_convbn* c1 = &m->convs[1 + 3 * b];
    const orc_convbn* c2 = &m->convs[2 + 3 * b];
    const orc_convbn* ds = &m->convs[3 + 3 * b];
    int H1, W1, H2, W2;
    float* t1 = conv_bn_f32(c1, cur, N, H, W, &H1, &W1); /* :164-169 */
    orc_relu_f32(t1, (size_t)N * c1->oc * H1 * W1);
    track(am, ORC_ACT_BLOCK0 + 3 * b + 0, t1, (size_t)N * c1->oc * H1 * W1);
    float* t2 = conv_bn_f32(c2, t1, N, H1, W1, &H2, &W2); /* :173-177 */
    free(t1);
    const size_t on = (size_t)N * c2->oc * H2 * W2;
    if (ds->w) { /* :187-196 */
      int Hd, Wd;
      float* sk = conv_bn_f32(ds, cur, N, H, W, &Hd, &Wd);
      track(am, ORC_ACT_BLOCK0 + 3 * b + 1, sk, on);
      orc_add_f32(t2, sk, on);
      free(sk);
    } else { /* identity :181-186 */
      orc_add_f32(t2, cur, on);
    }
    orc_relu_f32(t2, on); /* :199-200 */
    track(am, ORC_ACT_BLOCK0 + 3 * b + 2, t2, on);
    free(cur);
    cur = t2;
    H = H2;
    W = W2;
    C = c2->oc;
    if (ck) {
      float* dst = (b == 1) ? ck->layer1 : (b == 3) ? ck->layer2 : (b == 5) ? ck->layer3 : (b == 7) ? ck->layer4 : NULL;
      if (dst) memcpy(dst, cur, on * sizeof(float));
    }
  }
  float* gap = (float*)malloc((size_t)N * C * sizeof(float));
  orc_gap_f32(cur, N, C, H, W, gap); /* :418-424 */
  track(am, ORC_ACT_GAP, gap, (size_t)N * C);
  if (ck && ck->gap) memcpy(ck->gap, gap, (size_t)N * C * sizeof(float));
  orc_fc_f32(gap, m->fc_w, m->fc_b, logits, N, 1000, C); /* :432 */
  free(gap);
  free(cur);
}

/* ============================================================ INT8 spec operators (QUANT_SPEC.md) */
float orc_inv_scale(float s) { return (float)(1.0 / (double)s); }

static inline int8_t quant1(float t, int lo, int hi) {
  /* rne: rintf under the default FE_TONEAREST mode == __float2int_rn on the GPU.
   * clamp in float first so the int conversion cannot overflow (NaN -> lo). */
  float r = rintf(t);
  if (!(r >= (float)lo)) r = (float)lo;
  if (r > (float)hi) r = (float)hi;
  return (int8_t)(int)r;
}

void orc_quantize_f32_i8(const float* x, size_t n, float inv_s, int lo, int hi, int8_t* q) {
  for (size_t i = 0; i < n; ++i) q[i] = quant1(x[i] * inv_s, lo, hi);
}
void orc_dequantize_i8_f32(const int8_t* q, size_t n, float s, float* x) {
  for (size_t i = 0; i < n; ++i) x[i] = (float)q[i] * s;
}
void orc_dequantize_i8_f32_per_channel(const int8_t* q, int N, int C, int HW, const float* s, float* x) {
  for (size_t i = 0; i < (size_t)N * C * HW; ++i) x[i] = (float)q[i] * s[(i / HW) % C];
}
void orc_quantize_weights_per_channel(const float* w, int OC, int K, int8_t* q, float* s) {
  for (int oc = 0; oc < OC; ++oc) {
    const float am = absmax_f32(w + (size_t)oc * K, (size_t)K);
    const float sc = am > 0.f ? (float)((double)am / 127.0) : 1.0f;
    s[oc] = sc;
    const float inv = orc_inv_scale(sc);
    for (int k = 0; k < K; ++k) q[(size_t)oc * K + k] = quant1(w[(size_t)oc * K + k] * inv, -127, 127);
  }
}
void orc_fold_bn(const float* g, const float* b, const float* m, const float* v, float eps, const float* s_w,
                 float s_x, float s_y, int OC, float* alpha, float* beta) {
  for (int oc = 0; oc < OC; ++oc) {
    const double a = (double)g[oc] / sqrt((double)v[oc] + (double)eps);
    alpha[oc] = (float)((double)s_x * (double)s_w[oc] * a / (double)s_y);
    beta[oc] = (float)(((double)b[oc] - (double)m[oc] * a) / (double)s_y);
  }
}
float orc_res_mul(float s_r, float s_y) { return (float)((double)s_r / (double)s_y); }

/* epilogue, QUANT_SPEC §3 — each line is one binary32 rounding */
static inline int8_t epilogue1(int32_t acc, float alpha, float beta, int has_res, int8_t r, float res_mul,
                               int relu) {
  float t = fmaf((float)acc, alpha, beta);
  if (has_res) t = fmaf((float)r, res_mul, t);
  /* ReLU (reference form `if (x < 0) x = 0`, K/relu.cu:9) commutes with the positive rescale and the
   * rounding, so it is exactly the lower clamp at 0 */
  return quant1(t, relu ? 0 : -128, 127);
}

void orc_conv2d_i8(const int8_t* x, int N, int C, int H, int W, const int8_t* w, int OC, int kH, int kW, int sH,
                   int sW, int pH, int pW, const orc_epilogue* ep, int32_t* acc_out, int8_t* y) {
  const int OH = (H + 2 * pH - kH) / sH + 1, OW = (W + 2 * pW - kW) / sW + 1;
  const int K = C * kH * kW;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int oc = 0; oc < OC; ++oc) {
      const int8_t* xn = x + (size_t)n * C * H * W;
      const int8_t* wr = w + (size_t)oc * K;
      const size_t obase = ((size_t)n * OC + oc) * OH * OW;
      for (int oh = 0; oh < OH; ++oh)
        for (int ow = 0; ow < OW; ++ow) {
          int32_t acc = 0;
          for (int kh = 0; kh < kH; ++kh) {
            const int ih = oh * sH - pH + kh;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < kW; ++kw) {
              const int iw = ow * sW - pW + kw;
              if (iw < 0 || iw >= W) continue;
              const int8_t* xp = xn + (size_t)ih * W + iw;
              const int8_t* wp = wr + kh * kW + kw;
              for (int c = 0; c < C; ++c)
                acc += (int32_t)xp[(size_t)c * H * W] * (int32_t)wp[c * kH * kW];
            }
          }
          const size_t o = obase + (size_t)oh * OW + ow;
          if (acc_out) acc_out[o] = acc;
          if (y && ep)
            y[o] = epilogue1(acc, ep->alpha[oc], ep->beta[oc], ep->residual != NULL,
                             ep->residual ? ep->residual[o] : 0, ep->res_mul, ep->relu);
        }
    }
}

/* QUANT_SPEC §4: standalone residual add of two int8 tensors with different scales */
void orc_add_requant_i8(int8_t* y, float y_scale, const int8_t* x, float x_scale, size_t n, int relu, float out_scale) {
  const float inv = orc_inv_scale(out_scale);
  for (size_t i = 0; i < n; ++i) {
    float t = (float)y[i] * y_scale;
    t = fmaf((float)x[i], x_scale, t);
    if (relu && t < 0.f) t = 0.f;
    y[i] = quant1(t * inv, relu ? 0 : -128, 127);
  }
}

/* max-pool on int8 (monotone => commutes with quantisation); geometry as K/maxpool2d.cu:14-40 */
void orc_maxpool3x3s2p1_i8(const int8_t* x, int N, int C, int H, int W, int8_t* y) {
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const int8_t* xp = x + (size_t)nc * H * W;
    int8_t* yp = y + (size_t)nc * OH * OW;
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) {
        int vmax = -128;
        for (int kh = 0; kh < 3; ++kh) {
          const int ih = oh * 2 - 1 + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < 3; ++kw) {
            const int iw = ow * 2 - 1 + kw;
            if (iw < 0 || iw >= W) continue;
            const int v = xp[ih * W + iw];
            vmax = v > vmax ? v : vmax;
          }
        }
        yp[oh * OW + ow] = (int8_t)vmax;
      }
  }
}

void orc_gap_i8(const int8_t* x, int N, int C, int H, int W, float scale_over_hw, float inv_out_scale,
                float* y_f32, int8_t* y_i8) {
  const int HW = H * W;
  for (int nc = 0; nc < N * C; ++nc) {
    int32_t sum = 0;
    for (int i = 0; i < HW; ++i) sum += x[(size_t)nc * HW + i];
    const float g = (float)sum * scale_over_hw;
    if (y_f32) y_f32[nc] = g;
    if (y_i8) y_i8[nc] = quant1(g * inv_out_scale, -128, 127);
  }
}

void orc_fc_i8(const int8_t* g, const int8_t* w, const float* w_scale_times_g, const float* bias, int N, int O,
               int I, int32_t* acc_out, float* logits) {
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < O; ++o) {
      int32_t acc = 0;
      for (int k = 0; k < I; ++k) acc += (int32_t)w[(size_t)o * I + k] * (int32_t)g[(size_t)n * I + k];
      if (acc_out) acc_out[(size_t)n * O + o] = acc;
      if (logits) logits[(size_t)n * O + o] = fmaf((float)acc, w_scale_times_g[o], bias[o]);
    }
}

static int8_t* convq(const orc_convq* p, const int8_t* x, int N, int H, int W, const int8_t* res, float res_mul,
                     int relu, int* OH, int* OW) {
  *OH = (H + 2 * p->pad - p->k) / p->stride + 1;
  *OW = (W + 2 * p->pad - p->k) / p->stride + 1;
  int8_t* y = (int8_t*)malloc((size_t)N * p->oc * (*OH) * (*OW));
  orc_epilogue ep = {p->alpha, p->beta, res, res_mul, relu};
  orc_conv2d_i8(x, N, p->ic, H, W, p->w, p->oc, p->k, p->k, p->stride, p->stride, p->pad, p->pad, &ep, NULL, y);
  return y;
}

/* Whole network in INT8; wiring identical to orc_resnet18_f32_forward, arithmetic per QUANT_SPEC.md §5. */
void orc_resnet18_i8_forward(const orc_resnet18_i8* m, const float* x, int N, float* logits,
                             orc_checkpoints_i8* ck) {
  const float* S = m->act_scale;
  int H = 224, W = 224, OH, OW;
  int8_t* qx = (int8_t*)malloc((size_t)N * 3 * H * W);
  orc_quantize_f32_i8(x, (size_t)N * 3 * H * W, orc_inv_scale(S[ORC_ACT_INPUT]), -128, 127, qx);
  int8_t* y0 = convq(&m->convs[0], qx, N, H, W, NULL, 0.f, 1, &OH, &OW);
  free(qx);
  const int PH = (OH + 2 - 3) / 2 + 1, PW = (OW + 2 - 3) / 2 + 1;
  int8_t* cur = (int8_t*)malloc((size_t)N * 64 * PH * PW);
  orc_maxpool3x3s2p1_i8(y0, N, 64, OH, OW, cur);
  free(y0);
  H = PH;
  W = PW;
  if (ck && ck->stem_pool) memcpy(ck->stem_pool, cur, (size_t)N * 64 * H * W);
  float s_cur = S[ORC_ACT_STEM];
  int C = 64;
  for (int b = 0; b < ORC_NUM_BLOCKS; ++b) {
    const orc_convq* c1 = &m->convs[1 + 3 * b];
    const orc_convq* c2 = &m->convs[2 + 3 * b];
    const orc_convq* ds = &m->convs[3 + 3 * b];
    const float s_ds = S[ORC_ACT_BLOCK0 + 3 * b + 1], s_out = S[ORC_ACT_BLOCK0 + 3 * b + 2];
    int H1, W1, H2, W2;
    int8_t* t1 = convq(c1, cur, N, H, W, NULL, 0.f, 1, &H1, &W1);
    int8_t* t2;
    if (ds->w) {
      int Hd, Wd;
      int8_t* sk = convq(ds, cur, N, H, W, NULL, 0.f, 0, &Hd, &Wd);
      t2 = convq(c2, t1, N, H1, W1, sk, orc_res_mul(s_ds, s_out), 1, &H2, &W2);
      free(sk);
    } else {
      t2 = convq(c2, t1, N, H1, W1, cur, orc_res_mul(s_cur, s_out), 1, &H2, &W2);
    }
    free(t1);
    free(cur);
    cur = t2;
    s_cur = s_out;
    H = H2;
    W = W2;
    C = c2->oc;
    if (ck) {
      int8_t* dst = (b == 1) ? ck->layer1 : (b == 3) ? ck->layer2 : (b == 5) ? ck->layer3 : (b == 7) ? ck->layer4 : NULL;
      if (dst) memcpy(dst, cur, (size_t)N * C * H * W);
    }
  }
  int8_t* g = (int8_t*)malloc((size_t)N * C);
  const float s_over_hw = (float)((double)s_cur / (double)(H * W));
  orc_gap_i8(cur, N, C, H, W, s_over_hw, orc_inv_scale(S[ORC_ACT_GAP]), NULL, g);
  if (ck && ck->gap) memcpy(ck->gap, g, (size_t)N * C);
  orc_fc_i8(g, m->fc_w, m->fc_scale, m->fc_b, N, 1000, C, NULL, logits);
  free(g);
  free(cur);
}

/* ============================================================ FP8 (E4M3) path — spec/QUANT_SPEC.md section 6 */
uint8_t orc_f32_to_e4m3(float f) {
  const uint8_t sign = signbit(f) ? 0x80 : 0x00;
  if (isnan(f)) return (uint8_t)(sign | 0x7F);
  const float a = fabsf(f);
  if (a >= 448.0f) return (uint8_t)(sign | 0x7E);            /* saturate to the largest finite value */
  if (a < 0.015625f) {                                       /* below 2^-6: subnormals, multiples of 2^-9 */
    const int q = (int)nearbyintf(a * 512.0f);               /* 0..8; 8 is the smallest normal (same encoding) */
    return (uint8_t)(sign | q);
  }
  int e;
  const float fr = frexpf(a, &e);                            /* a = fr * 2^e, fr in [0.5,1) */
  int ex = e - 1;
  int m = (int)nearbyintf((2.0f * fr - 1.0f) * 8.0f);        /* 3 mantissa bits, round half to even */
  if (m == 8) { m = 0; ++ex; }
  if (ex > 8 || (ex == 8 && m == 7)) return (uint8_t)(sign | 0x7E);
  return (uint8_t)(sign | ((ex + 7) << 3) | m);
}

float orc_e4m3_to_f32(uint8_t b) {
  const int sgn = b >> 7, ex = (b >> 3) & 0xF, m = b & 7;
  float v;
  if (ex == 0xF && m == 7) v = NAN;
  else if (ex == 0) v = ldexpf((float)m, -9);
  else v = ldexpf(1.0f + (float)m / 8.0f, ex - 7);
  return sgn ? -v : v;
}

void orc_quantize_f32_e4m3(const float* x, size_t n, float inv_s, uint8_t* q) {
  for (size_t i = 0; i < n; ++i) q[i] = orc_f32_to_e4m3(x[i] * inv_s);
}
void orc_dequantize_e4m3_f32(const uint8_t* q, size_t n, float s, float* x) {
  for (size_t i = 0; i < n; ++i) x[i] = orc_e4m3_to_f32(q[i]) * s;
}
void orc_quantize_weights_per_channel_e4m3(const float* w, int OC, int K, uint8_t* q, float* s) {
  for (int o = 0; o < OC; ++o) {
    float am = 0.f;
    for (int k = 0; k < K; ++k) am = fmaxf(am, fabsf(w[(size_t)o * K + k]));
    const float sc = am > 0.f ? (float)((double)am / 448.0) : 1.0f;
    s[o] = sc;
    const float inv = orc_inv_scale(sc);
    for (int k = 0; k < K; ++k) q[(size_t)o * K + k] = orc_f32_to_e4m3(w[(size_t)o * K + k] * inv);
  }
}

static uint8_t epilogue_e4m3(float acc, float alpha, float beta, int has_res, uint8_t r, float res_mul, int relu) {
  float t = fmaf(acc, alpha, beta);
  if (has_res) t = fmaf(orc_e4m3_to_f32(r), res_mul, t);
  if (relu && t < 0.f) t = 0.f;
  return orc_f32_to_e4m3(t);
}

void orc_conv2d_e4m3(const uint8_t* x, int N, int C, int H, int W, const uint8_t* w, int OC, int kH, int kW, int sH,
                     int sW, int pH, int pW, const orc_epilogue* ep, float* acc_out, uint8_t* y) {
  const int OH = (H + 2 * pH - kH) / sH + 1, OW = (W + 2 * pW - kW) / sW + 1;
  const int K = C * kH * kW;
  float lut[256];
  for (int i = 0; i < 256; ++i) lut[i] = orc_e4m3_to_f32((uint8_t)i);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int oc = 0; oc < OC; ++oc) {
      const uint8_t* xn = x + (size_t)n * C * H * W;
      const uint8_t* wr = w + (size_t)oc * K;
      const size_t obase = ((size_t)n * OC + oc) * OH * OW;
      for (int oh = 0; oh < OH; ++oh)
        for (int ow = 0; ow < OW; ++ow) {
          double acc = 0.0;
          for (int kh = 0; kh < kH; ++kh) {
            const int ih = oh * sH - pH + kh;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < kW; ++kw) {
              const int iw = ow * sW - pW + kw;
              if (iw < 0 || iw >= W) continue;
              const uint8_t* xp = xn + (size_t)ih * W + iw;
              const uint8_t* wp = wr + kh * kW + kw;
              for (int c = 0; c < C; ++c)
                acc += (double)lut[xp[(size_t)c * H * W]] * (double)lut[wp[c * kH * kW]];
            }
          }
          const size_t o = obase + (size_t)oh * OW + ow;
          if (acc_out) acc_out[o] = (float)acc;
          if (y && ep)
            y[o] = epilogue_e4m3((float)acc, ep->alpha[oc], ep->beta[oc], ep->residual != NULL,
                                 ep->residual ? (uint8_t)ep->residual[o] : 0, ep->res_mul, ep->relu);
        }
    }
}

void orc_gap_e4m3(const uint8_t* x, int N, int C, int H, int W, float scale_over_hw, float inv_out_scale, uint8_t* y) {
  const int HW = H * W;
  for (int nc = 0; nc < N * C; ++nc) {
    double sum = 0.0;
    for (int i = 0; i < HW; ++i) sum += (double)orc_e4m3_to_f32(x[(size_t)nc * HW + i]);
    y[nc] = orc_f32_to_e4m3(((float)sum * scale_over_hw) * inv_out_scale);
  }
}

void orc_fc_e4m3(const uint8_t* g, const uint8_t* w, const float* w_scale_times_g, const float* bias, int N, int O,
                 int I, float* logits) {
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < O; ++o) {
      double acc = 0.0;
      for (int k = 0; k < I; ++k) acc += (double)orc_e4m3_to_f32(w[(size_t)o * I + k]) * (double)orc_e4m3_to_f32(g[(size_t)n * I + k]);
      logits[(size_t)n * O + o] = fmaf((float)acc, w_scale_times_g[o], bias[o]);
    }
}

static uint8_t* convq8(const orc_convq* p, const uint8_t* x, int N, int H, int W, const uint8_t* res, float res_mul,
                       int relu, int* OH, int* OW) {
  *OH = (H + 2 * p->pad - p->k) / p->stride + 1;
  *OW = (W + 2 * p->pad - p->k) / p->stride + 1;
  uint8_t* y = (uint8_t*)malloc((size_t)N * p->oc * (*OH) * (*OW));
  orc_epilogue ep = {p->alpha, p->beta, (const int8_t*)res, res_mul, relu};
  orc_conv2d_e4m3(x, N, p->ic, H, W, (const uint8_t*)p->w, p->oc, p->k, p->k, p->stride, p->stride, p->pad, p->pad, &ep,
                  NULL, y);
  return y;
}

/* Whole network in E4M3; wiring identical to orc_resnet18_i8_forward.  The max-pool runs on the bytes as signed
 * int8: its input is a ReLU output (codes 0x00..0x7E), on which the E4M3 order and the byte order agree. */
void orc_resnet18_fp8_forward(const orc_resnet18_i8* m, const float* x, int N, float* logits, orc_checkpoints_i8* ck) {
  const float* S = m->act_scale;
  int H = 224, W = 224, OH, OW;
  uint8_t* qx = (uint8_t*)malloc((size_t)N * 3 * H * W);
  orc_quantize_f32_e4m3(x, (size_t)N * 3 * H * W, orc_inv_scale(S[ORC_ACT_INPUT]), qx);
  uint8_t* y0 = convq8(&m->convs[0], qx, N, H, W, NULL, 0.f, 1, &OH, &OW);
  free(qx);
  const int PH = (OH + 2 - 3) / 2 + 1, PW = (OW + 2 - 3) / 2 + 1;
  uint8_t* cur = (uint8_t*)malloc((size_t)N * 64 * PH * PW);
  orc_maxpool3x3s2p1_i8((const int8_t*)y0, N, 64, OH, OW, (int8_t*)cur);
  free(y0);
  H = PH;
  W = PW;
  if (ck && ck->stem_pool) memcpy(ck->stem_pool, cur, (size_t)N * 64 * H * W);
  float s_cur = S[ORC_ACT_STEM];
  int C = 64;
  for (int b = 0; b < ORC_NUM_BLOCKS; ++b) {
    const orc_convq* c1 = &m->convs[1 + 3 * b];
    const orc_convq* c2 = &m->convs[2 + 3 * b];
    const orc_convq* ds = &m->convs[3 + 3 * b];
    const float s_ds = S[ORC_ACT_BLOCK0 + 3 * b + 1], s_out = S[ORC_ACT_BLOCK0 + 3 * b + 2];
    int H1, W1, H2, W2;
    uint8_t* t1 = convq8(c1, cur, N, H, W, NULL, 0.f, 1, &H1, &W1);
    uint8_t* t2;
    if (ds->w) {
      int Hd, Wd;
      uint8_t* sk = convq8(ds, cur, N, H, W, NULL, 0.f, 0, &Hd, &Wd);
      t2 = convq8(c2, t1, N, H1, W1, sk, orc_res_mul(s_ds, s_out), 1, &H2, &W2);
      free(sk);
    } else {
      t2 = convq8(c2, t1, N, H1, W1, cur, orc_res_mul(s_cur, s_out), 1, &H2, &W2);
    }
    free(t1);
    free(cur);
    cur = t2;
    s_cur = s_out;
    H = H2;
    W = W2;
    C = c2->oc;
    if (ck) {
      int8_t* dst = (b == 1) ? ck->layer1 : (b == 3) ? ck->layer2 : (b == 5) ? ck->layer3 : (b == 7) ? ck->layer4 : NULL;
      if (dst) memcpy(dst, cur, (size_t)N * C * H * W);
    }
  }
  uint8_t* g = (uint8_t*)malloc((size_t)N * C);
  const float s_over_hw = (float)((double)s_cur / (double)(H * W));
  orc_gap_e4m3(cur, N, C, H, W, s_over_hw, orc_inv_scale(S[ORC_ACT_GAP]), g);
  if (ck && ck->gap) memcpy(ck->gap, g, (size_t)N * C);
  orc_fc_e4m3(g, (const uint8_t*)m->fc_w, m->fc_scale, m->fc_b, N, 1000, C, logits);
  free(g);
  free(cur);
}

/* ============================================================ MNIST MLP forward (config #1) */
/* MN/v3.c:125-134 matmul_a_b (accumulates through memory in l order; plain mul+add, no FMA since
 * this file is built with -ffp-contract=off like `gcc -O2` on x86-64 without -mfma),
 * :168-174 bias_forward, :161-165 relu_forward (fmaxf), :108-123 softmax (expf, clamp >= 1e-7),
 * sequence :177-215 forward_timed. */
void orc_mnist_mlp_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                           int batch, int in_dim, int hid, int out_dim, float* hidden, float* out) {
  for (int i = 0; i < batch; ++i)
    for (int j = 0; j < hid; ++j) {
      float c = 0.0f;
      for (int l = 0; l < in_dim; ++l) c += x[(size_t)i * in_dim + l] * w1[(size_t)l * hid + j];
      hidden[(size_t)i * hid + j] = c;
    }
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < hid; ++i) hidden[(size_t)b * hid + i] += b1[i];
  for (size_t i = 0; i < (size_t)batch * hid; ++i) hidden[i] = fmaxf(0.0f, hidden[i]);
  for (int i = 0; i < batch; ++i)
    for (int j = 0; j < out_dim; ++j) {
      float c = 0.0f;
      for (int l = 0; l < hid; ++l) c += hidden[(size_t)i * hid + l] * w2[(size_t)l * out_dim + j];
      out[(size_t)i * out_dim + j] = c;
    }
  for (int b = 0; b < batch; ++b)
    for (int i = 0; i < out_dim; ++i) out[(size_t)b * out_dim + i] += b2[i];
  for (int b = 0; b < batch; ++b) {
    float* o = out + (size_t)b * out_dim;
    float mx = o[0];
    for (int i = 1; i < out_dim; ++i)
      if (o[i] > mx) mx = o[i];
    float sum = 0.0f;
    for (int i = 0; i < out_dim; ++i) {
      o[i] = expf(o[i] - mx);
      sum += o[i];
    }
    for (int i = 0; i < out_dim; ++i) o[i] = fmaxf(o[i] / sum, 1e-7f);
  }
}
