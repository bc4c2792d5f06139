// dlq_bench — throughput / latency of the INT8 (or FP8) ResNet-18 path measured from C++ through the C ABI only
// (no Python, no torch): the host-language counterpart of bench.py for the reference's C++ users.
//
//   dlq_bench [--weights DIR] [--batch 256] [--iters 50] [--warmup 10] [--fp8] [--gpus G | --devices 0,1,..]
//
// --weights DIR : weight directory in the reference's export format (tools/export_resnet18.py:85-92), with or without
//                 the quant block; without it (and without --weights at all: synthetic weights, the recipe of
//                 dlq_b200/synth.py) the activation scales are calibrated here, on 8 synthetic images, through the
//                 reference's FP32 arithmetic on the GPU.
// one GPU  : device-resident images/s (cudaEvents on the library's stream), host-buffer images/s
//            (dlq_resnet18_forward_host from pinned memory, H2D + D2H inside the timed region), batch-1 CUDA-graph latency.
// --gpus G : BASELINE config 4 - `batch` images split over G devices by the batch-sharded driver (dlq_multi_*): host
//            buffers in, logits gathered on the host, wall-clock (synchronous, pipelined fp32, pipelined uint8) and the
//            device-resident number.  --devices lists them explicitly (a device may repeat).
// Prints one JSON line.  Exit codes: 1 usage / IO, 3 CUDA or library error.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/dlq.h"

#define CK_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s @ %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 3;                                                                                \
    }                                                                                          \
  } while (0)
#define CK_DLQ(ctx, call)                                                                                 \
  do {                                                                                                    \
    int rc_ = (call);                                                                                     \
    if (rc_ != DLQ_OK) {                                                                                  \
      fprintf(stderr, "dlq error %d: %s @ %s:%d\n", rc_, dlq_last_error_string(ctx), __FILE__, __LINE__); \
      return 3;                                                                                           \
    }                                                                                                     \
  } while (0)

static const size_t kImg = 3 * 224 * 224;

// synthetic weights: dlq_b200/synth.py make_weights(seed) restated (same names, ranges and shifts)
struct SynthWeights {
  std::vector<std::vector<float>> store;
  dlq_resnet18_weights w;
};
static std::string conv_key(int i) {
  if (i == 0) return "conv1.weight";
  const int b = (i - 1) / 3, r = (i - 1) % 3;
  const std::string base = "layer" + std::to_string(b / 2 + 1) + "." + std::to_string(b % 2);
  return base + (r == 0 ? ".conv1.weight" : r == 1 ? ".conv2.weight" : ".downsample.0.weight");
}
static std::string bn_key(int i) {
  if (i == 0) return "bn1";
  const int b = (i - 1) / 3, r = (i - 1) % 3;
  const std::string base = "layer" + std::to_string(b / 2 + 1) + "." + std::to_string(b % 2);
  return base + (r == 0 ? ".bn1" : r == 1 ? ".bn2" : ".downsample.1");
}
static int weight_shift(int fan_in) { return static_cast<int>(std::lround(std::log2(73.3 * std::sqrt(fan_in / 2.0)))); }
static void make_synth(uint64_t seed, SynthWeights* s) {
  memset(&s->w, 0, sizeof(s->w));
  s->store.reserve(128);
  auto fill = [&](size_t n, const std::string& name, int lo, int hi, int shift) -> const float* {
    s->store.emplace_back(n);
    dlq_synth_fill_f32(s->store.back().data(), n, seed, name.c_str(), lo, hi, shift);
    return s->store.back().data();
  };
  const int chans[4] = {64, 128, 256, 512};
  struct G { int ic, oc, k; };
  G g[DLQ_NUM_CONVS] = {};
  g[0] = {3, 64, 7};
  int ic = 64;
  for (int b = 0; b < 8; ++b) {
    const int oc = chans[b / 2];
    const bool down = (b % 2 == 0 && b >= 2);
    g[1 + 3 * b] = {ic, oc, 3};
    g[2 + 3 * b] = {oc, oc, 3};
    g[3 + 3 * b] = down ? G{ic, oc, 1} : G{0, 0, 0};
    ic = oc;
  }
  for (int i = 0; i < DLQ_NUM_CONVS; ++i) {
    if (!g[i].oc) continue;
    const int fan = g[i].ic * g[i].k * g[i].k;
    s->w.conv_w[i] = fill(static_cast<size_t>(g[i].oc) * fan, conv_key(i), -127, 127, weight_shift(fan));
    s->w.bn_gamma[i] = fill(g[i].oc, bn_key(i) + ".weight", 128, 384, 8);
    s->w.bn_beta[i] = fill(g[i].oc, bn_key(i) + ".bias", -64, 64, 8);
    s->w.bn_mean[i] = fill(g[i].oc, bn_key(i) + ".running_mean", -64, 64, 8);
    s->w.bn_var[i] = fill(g[i].oc, bn_key(i) + ".running_var", 128, 384, 8);
  }
  s->w.fc_w = fill(1000 * 512, "fc.weight", -127, 127, weight_shift(512));
  s->w.fc_b = fill(1000, "fc.bias", -64, 64, 8);
}

int main(int argc, char** argv) {
  std::string wdir;
  int batch = 256, iters = 50, warmup = 10, gpus = 1;
  bool fp8 = false;
  std::vector<int> devlist;      // --devices 0,0 : explicit device list (a device may repeat: independent replicas)
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--weights" && i + 1 < argc) wdir = argv[++i];
    else if (a == "--batch" && i + 1 < argc) batch = atoi(argv[++i]);
    else if (a == "--iters" && i + 1 < argc) iters = atoi(argv[++i]);
    else if (a == "--warmup" && i + 1 < argc) warmup = atoi(argv[++i]);
    else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (a == "--devices" && i + 1 < argc) {
      for (const char* p = argv[++i]; *p;) { devlist.push_back(atoi(p)); while (*p && *p != ',') ++p; if (*p) ++p; }
      gpus = static_cast<int>(devlist.size());
    }
    else if (a == "--fp8") fp8 = true;
    else { fprintf(stderr, "usage: dlq_bench [--weights DIR] [--batch B] [--iters K] [--warmup W] [--fp8] [--gpus G | --devices a,b,..]\n"); return 1; }
  }
  if (batch < 1 || iters < 1 || warmup < 0 || gpus < 1 || batch % gpus) { fprintf(stderr, "bad batch / iters / gpus\n"); return 1; }

  // ---- weights
  SynthWeights synth;
  dlq_weight_dir* wd = nullptr;
  dlq_resnet18_weights W;
  if (!wdir.empty()) {
    char err[512];
    if (dlq_weight_dir_load(wdir.c_str(), &wd, err, sizeof err) != DLQ_OK) { fprintf(stderr, "%s\n", err); return 1; }
    W = *dlq_weight_dir_weights(wd);
  } else {
    make_synth(0, &synth);
    W = synth.w;
  }
  dlq_ctx* ctx = nullptr;
  if (dlq_create(0, &ctx) != DLQ_OK) { fprintf(stderr, "dlq_create failed (an sm_100 GPU is required; there is no CPU path)\n"); return 3; }

  // ---- synthetic images (SURVEY 8d lattice), 8 distinct ones tiled over the batch
  std::vector<float> img8(8 * kImg);
  dlq_synth_fill_f32(img8.data(), img8.size(), 0, "input", -192, 192, 6);

  // ---- PTQ calibration when the weights carry no (INT8) scales, or for E4M3
  bool have = false;
  for (int i = 0; i < DLQ_NUM_ACTS; ++i) have |= W.act_scale[i] != 0.f;
  if (!have || fp8) {
    dlq_resnet18_f32* f = nullptr;
    float *dx = nullptr, *dl = nullptr;
    CK_DLQ(ctx, dlq_resnet18_f32_create(ctx, &W, 8, &f));
    CK_CUDA(cudaMalloc(&dx, img8.size() * 4));
    CK_CUDA(cudaMalloc(&dl, 8 * 1000 * 4));
    CK_CUDA(cudaMemcpy(dx, img8.data(), img8.size() * 4, cudaMemcpyHostToDevice));
    CK_DLQ(ctx, dlq_resnet18_f32_forward(f, dx, 8, dl));
    float am[DLQ_NUM_ACTS];
    CK_DLQ(ctx, dlq_resnet18_f32_absmax(f, am));
    dlq_act_scales_from_absmax(am, fp8 ? 1 : 0, W.act_scale);
    dlq_resnet18_f32_destroy(f);
    cudaFree(dx); cudaFree(dl);
  }
  W.fp8 = fp8 ? 1 : 0;

  float* hx = nullptr;   // pinned host batch + logits
  float* hl = nullptr;
  CK_CUDA(cudaMallocHost(&hx, static_cast<size_t>(batch) * kImg * 4));
  CK_CUDA(cudaMallocHost(&hl, static_cast<size_t>(batch) * 1000 * 4));
  for (int n = 0; n < batch; ++n) memcpy(hx + static_cast<size_t>(n) * kImg, img8.data() + static_cast<size_t>(n % 8) * kImg, kImg * 4);

  if (gpus > 1 || !devlist.empty()) {
    // ---- BASELINE config 4: `batch` images split over G devices by the batch-sharded driver.  Four numbers: host fp32
    // (synchronous call per batch), host fp32 / uint8 pipelined (submit k+1 while k runs; "all devices start -> logits
    // gathered on host", wall clock), and device-resident (compute only).
    std::vector<int> dev(gpus);
    for (int g = 0; g < gpus; ++g) dev[g] = devlist.empty() ? g : devlist[g];
    dlq_multi* mm = nullptr;
    if (dlq_multi_create(dev.data(), gpus, &W, batch / gpus, &mm) != DLQ_OK) { fprintf(stderr, "dlq_multi_create failed\n"); return 3; }
#define CK_MULTI(call) do { if ((call) != DLQ_OK) { fprintf(stderr, "%s\n", dlq_multi_last_error_string(mm)); return 3; } } while (0)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    // (1) synchronous fp32 host batches
    for (int i = 0; i < std::max(1, warmup); ++i) CK_MULTI(dlq_multi_forward_host(mm, hx, batch, hl));
    auto t0 = now();
    for (int i = 0; i < iters; ++i) CK_MULTI(dlq_multi_forward_host(mm, hx, batch, hl));
    const double sec_sync = secs(t0);
    const int top0 = static_cast<int>(std::max_element(hl, hl + 1000) - hl);
    // (2) pipelined fp32 host batches
    t0 = now();
    for (int i = 0; i < iters; ++i) CK_MULTI(dlq_multi_submit_host(mm, hx, batch, hl));
    CK_MULTI(dlq_multi_wait(mm));
    const double sec_pipe = secs(t0);
    // (3) pipelined uint8 host batches (ImageNet mean / std table on the device)
    uint8_t* hu = nullptr;
    CK_CUDA(cudaMallocHost(&hu, static_cast<size_t>(batch) * kImg));
    for (size_t i = 0; i < static_cast<size_t>(batch) * kImg; ++i) hu[i] = static_cast<uint8_t>((i * 2654435761u) >> 24);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
    CK_MULTI(dlq_multi_set_preprocess(mm, mean, sd));
    CK_MULTI(dlq_multi_forward_host_u8(mm, hu, batch, hl));
    t0 = now();
    for (int i = 0; i < iters; ++i) CK_MULTI(dlq_multi_submit_host_u8(mm, hu, batch, hl));
    CK_MULTI(dlq_multi_wait(mm));
    const double sec_u8 = secs(t0);
    // (4) device-resident shards
    std::vector<const float*> dxs(gpus);
    std::vector<float*> dls(gpus);
    std::vector<int> per(gpus, batch / gpus);
    for (int g = 0; g < gpus; ++g) {
      CK_CUDA(cudaSetDevice(dev[g]));
      float* p = nullptr;
      CK_CUDA(cudaMalloc(&p, static_cast<size_t>(per[g]) * kImg * 4));
      CK_CUDA(cudaMemcpy(p, hx + static_cast<size_t>(g) * per[g] * kImg, static_cast<size_t>(per[g]) * kImg * 4, cudaMemcpyHostToDevice));
      dxs[g] = p;
      CK_CUDA(cudaMalloc(&dls[g], static_cast<size_t>(per[g]) * 1000 * 4));
    }
    for (int i = 0; i < 3; ++i) CK_MULTI(dlq_multi_forward_device(mm, dxs.data(), per.data(), dls.data()));
    t0 = now();
    for (int i = 0; i < iters; ++i) CK_MULTI(dlq_multi_forward_device(mm, dxs.data(), per.data(), dls.data()));
    const double sec_dev = secs(t0);
    for (int g = 0; g < gpus; ++g) { cudaSetDevice(dev[g]); cudaFree(const_cast<float*>(dxs[g])); cudaFree(dls[g]); }
    std::string devs;
    for (int g = 0; g < gpus; ++g) devs += (g ? "," : "") + std::to_string(dev[g]);
    const double B = static_cast<double>(batch) * iters;
    printf("{\"driver\": \"dlq_bench (C++)\", \"dtype\": \"%s\", \"n_gpus\": %d, \"devices\": [%s], \"global_batch\": %d, "
           "\"batch_per_gpu\": %d, \"iters\": %d, \"host_images_per_s\": %.1f, \"host_pipelined_images_per_s\": %.1f, "
           "\"host_u8_pipelined_images_per_s\": %.1f, \"device_resident_images_per_s\": %.1f, \"ms_per_batch\": %.4f, "
           "\"top1_image0\": %d, "
           "\"how\": \"dlq_multi_*: persistent worker thread per device, pinned host batch split contiguously, logits written "
           "straight into one host array, wall clock; pipelined = submit k+1 while k runs; device-resident = one synchronised "
           "forward per replica and iteration\"}\n",
           fp8 ? "e4m3" : "s8", gpus, devs.c_str(), batch, batch / gpus, iters, B / sec_sync, B / sec_pipe, B / sec_u8, B / sec_dev,
           1e3 * sec_sync / iters, top0);
    cudaFreeHost(hu);
    dlq_multi_destroy(mm);
  } else {
    dlq_resnet18* net = nullptr;
    CK_DLQ(ctx, dlq_resnet18_create(ctx, &W, batch, &net));
    float *dx = nullptr, *dl = nullptr;
    CK_CUDA(cudaMalloc(&dx, static_cast<size_t>(batch) * kImg * 4));
    CK_CUDA(cudaMalloc(&dl, static_cast<size_t>(batch) * 1000 * 4));
    CK_CUDA(cudaMemcpy(dx, hx, static_cast<size_t>(batch) * kImg * 4, cudaMemcpyHostToDevice));
    cudaStream_t st = static_cast<cudaStream_t>(dlq_stream(ctx));
    cudaEvent_t e0, e1;
    CK_CUDA(cudaEventCreate(&e0));
    CK_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < std::max(3, warmup); ++i) CK_DLQ(ctx, dlq_resnet18_forward(net, dx, batch, dl));
    CK_DLQ(ctx, dlq_sync(ctx));
    CK_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) CK_DLQ(ctx, dlq_resnet18_forward(net, dx, batch, dl));
    CK_CUDA(cudaEventRecord(e1, st));
    CK_DLQ(ctx, dlq_sync(ctx));
    float ms = 0.f;
    CK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // host buffers
    const int hiters = std::max(3, std::min(iters, 20));
    for (int i = 0; i < 2; ++i) CK_DLQ(ctx, dlq_resnet18_forward_host(net, hx, batch, hl));
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < hiters; ++i) CK_DLQ(ctx, dlq_resnet18_forward_host(net, hx, batch, hl));
    const double hsec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int top0 = static_cast<int>(std::max_element(hl, hl + 1000) - hl);
    // host buffers, pipelined uint8 images: submit k+1 while k runs
    uint8_t* hu = nullptr;
    CK_CUDA(cudaMallocHost(&hu, static_cast<size_t>(batch) * kImg));
    for (size_t i = 0; i < static_cast<size_t>(batch) * kImg; ++i) hu[i] = static_cast<uint8_t>((i * 2654435761u) >> 24);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
    CK_DLQ(ctx, dlq_resnet18_set_preprocess(net, mean, sd));
    CK_DLQ(ctx, dlq_resnet18_forward_host_u8(net, hu, batch, hl));
    const auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < hiters; ++i) {
      CK_DLQ(ctx, dlq_resnet18_submit_host_u8(net, hu, batch, hl));
      if (i > 0) CK_DLQ(ctx, dlq_resnet18_wait(net));
    }
    CK_DLQ(ctx, dlq_resnet18_wait(net));
    const double usec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    cudaFreeHost(hu);
    // batch-1 latency, CUDA graph
    dlq_resnet18* net1 = nullptr;
    CK_DLQ(ctx, dlq_resnet18_create(ctx, &W, 1, &net1));
    CK_DLQ(ctx, dlq_resnet18_graph_capture(net1, dx, 1, dl));
    for (int i = 0; i < 20; ++i) CK_DLQ(ctx, dlq_resnet18_graph_launch(net1));
    CK_DLQ(ctx, dlq_sync(ctx));
    CK_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < 200; ++i) CK_DLQ(ctx, dlq_resnet18_graph_launch(net1));
    CK_CUDA(cudaEventRecord(e1, st));
    CK_DLQ(ctx, dlq_sync(ctx));
    float lat_ms = 0.f;
    CK_CUDA(cudaEventElapsedTime(&lat_ms, e0, e1));
    printf("{\"driver\": \"dlq_bench (C++)\", \"dtype\": \"%s\", \"n_gpus\": 1, \"batch\": %d, \"iters\": %d, "
           "\"images_per_s\": %.1f, \"ms_per_step\": %.4f, \"host_images_per_s\": %.1f, \"host_u8_pipelined_images_per_s\": %.1f, "
           "\"latency_b1_us\": %.1f, "
           "\"launches_per_forward\": %d, \"top1_image0\": %d, "
           "\"how\": \"device-resident: cudaEvents on the library stream; host: dlq_resnet18_forward_host from pinned memory, "
           "wall clock; latency: 200 back-to-back CUDA-graph replays at batch 1\"}\n",
           fp8 ? "e4m3" : "s8", batch, iters, batch * static_cast<double>(iters) / (ms * 1e-3), ms / iters,
           batch * static_cast<double>(hiters) / hsec, batch * static_cast<double>(hiters) / usec, 1e3 * lat_ms / 200,
           dlq_resnet18_launches_for_batch(net, batch), top0);
    dlq_resnet18_destroy(net1);
    dlq_resnet18_destroy(net);
    cudaFree(dx); cudaFree(dl);
  }
  cudaFreeHost(hx); cudaFreeHost(hl);
  dlq_destroy(ctx);
  if (wd) dlq_weight_dir_free(wd);
  return 0;
}
