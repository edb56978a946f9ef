// jxlb200 — the C ABI (include/jxlb200.h): context, device arenas, pipeline driver, taps.
// Replaces the encoder seam of the thesis harness, DockerManager::execute_cjxl
// (benchmark-jpegxl/src/docker_manager.rs:100-137).  No CPU fallback: every stage of the
// hot path is a kernel launch on the context's stream.
#include "../../include/jxlb200.h"
#include "encoder.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>
#include <algorithm>

using namespace jxlb;

// One context = a set of independent pipelines (each an Encoder with its own CUDA stream and device
// arenas).  Single encodes use pipeline 0; the batch entry points keep several images in flight so
// that the serial stages of one image (the per-group rANS chains) overlap the other images' work.
struct jxlb200_ctx {
  Encoder enc;                        // pipeline 0
  std::vector<Encoder*> extra;        // pipelines 1..P-1, created on first batch use
  int device = 0;
  int num_pipelines = 4;
  cudaStream_t copy_stream = nullptr; // batch mode: one H2D stream for all pipelines
  bool use_copy_stream = true;
  int ans_warps_single = 8;           // warps per rANS CTA for a single image (4: 2.51 ms, 8: 2.45 ms, 16: 2.65 ms per 4K frame) ($JXLB200_ANS_WARPS_SINGLE)
  int ans_warps = 16;                 // warps per rANS CTA in batch mode (32 measured no faster: the chains slow down)
  bool fork_single = true;            // a lone frame's independent stages on auxiliary streams ($JXLB200_FORK=0 switches it off)
  int ans_gpw = 2;                    // AC groups per warp of the rANS kernel in batch mode (fewer, longer-lived CTAs)
  std::vector<uint8_t> forced_acs;    // jxlb200_debug_set_strategy_map: handed to every pipeline
  int forced_bxs = 0, forced_bys = 0;
  std::string err;
  Encoder* pipe(int i) { return i == 0 ? &enc : extra[i - 1]; }
};

static bool ensure_pipelines(jxlb200_ctx* ctx, int n) {
  while ((int)ctx->extra.size() + 1 < n) {
    Encoder* e = new Encoder();
    std::string err;
    if (!e->Init(ctx->device, &err)) { delete e; ctx->err = err; return false; }
    if (!ctx->forced_acs.empty() && !e->SetForcedAcs(ctx->forced_acs.data(), ctx->forced_bxs, ctx->forced_bys, &err)) { delete e; ctx->err = err; return false; }
    ctx->extra.push_back(e);
  }
  return true;
}

static int fail(jxlb200_ctx* ctx, const std::string& msg, int code = -1) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" {

int jxlb200_abi_version(void) { return JXLB200_ABI_VERSION; }

jxlb200_ctx* jxlb200_create(int device) {
  // one hardware work queue per pipeline stream; only effective when the CUDA context does not exist yet
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);   // (32 is the maximum the driver accepts)
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
  if (prop.major != 10) return nullptr;  // sm_100a only; there is no fallback path
  jxlb200_ctx* ctx = new jxlb200_ctx();
  ctx->device = device;
  if (const char* env = getenv("JXLB200_PIPELINES")) { const int v = atoi(env); if (v >= 1 && v <= 64) ctx->num_pipelines = v; }
  if (const char* env = getenv("JXLB200_COPY_STREAM")) ctx->use_copy_stream = atoi(env) != 0;
  if (const char* env = getenv("JXLB200_FORK")) ctx->fork_single = atoi(env) != 0;
  if (const char* env = getenv("JXLB200_ANS_WARPS_SINGLE")) { const int v = atoi(env); if (v == 4 || v == 8 || v == 16) ctx->ans_warps_single = v; }
  if (const char* env = getenv("JXLB200_ANS_WARPS")) { const int v = atoi(env); if (v == 4 || v == 8 || v == 16) ctx->ans_warps = v; }
  if (const char* env = getenv("JXLB200_ANS_GPW")) { const int v = atoi(env); if (v >= 1 && v <= 64) ctx->ans_gpw = v; }
  std::string e;
  if (!ctx->enc.Init(device, &e)) { delete ctx; return nullptr; }
  return ctx;
}

void jxlb200_destroy(jxlb200_ctx* ctx) {
  if (!ctx) return;
  ctx->enc.Destroy();
  for (Encoder* e : ctx->extra) { e->Destroy(); delete e; }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* jxlb200_last_error(const jxlb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int check_params(jxlb200_ctx* ctx, uint32_t w, uint32_t h, const jxlb200_params* p) {
  if (!p) return fail(ctx, "null params");
  if (w == 0 || h == 0) return fail(ctx, "invalid image");
  if ((uint64_t)w * h > (1ull << 28)) return fail(ctx, "image too large");
  // several launches put image rows in grid.y (limit 65535): refuse taller images up front instead of failing a launch
  if (h > 65528u) return fail(ctx, "image taller than 65528 rows is not supported");
  if (!(p->distance >= 0.01f && p->distance <= 25.0f)) return fail(ctx, "distance out of range [0.01, 25]");
  if (p->effort < 1 || p->effort > 9) return fail(ctx, "effort out of range [1, 9]");
  if (p->proposal > 3) return fail(ctx, "unknown proposal");
  return 0;
}

int jxlb200_encode_device(jxlb200_ctx* ctx, const uint8_t* d_pixels, uint32_t width, uint32_t height, size_t stride,
                          const jxlb200_params* params, jxlb200_stats* stats) {
  if (!ctx) return -1;
  if (!d_pixels) return fail(ctx, "invalid image");
  if (int rc = check_params(ctx, width, height, params)) return rc;
  if (stride < (size_t)3 * width) return fail(ctx, "stride smaller than 3*width");
  std::string e;
  EncodeParams ep{params->distance, params->effort, params->proposal, params->flags};
  ctx->enc.set_ans_groups_per_warp(1);
  ctx->enc.set_ans_warps(ctx->ans_warps_single);
  ctx->enc.set_fork(ctx->fork_single);
  if (!ctx->enc.EncodeDevice(d_pixels, (int)width, (int)height, stride, ep, stats, &e)) return fail(ctx, e);
  return 0;
}

int jxlb200_fetch(jxlb200_ctx* ctx, uint8_t** out, size_t* out_len) {
  if (!ctx) return -1;
  if (!out || !out_len) return fail(ctx, "null output");
  std::string e;
  if (!ctx->enc.Fetch(out, out_len, &e)) return fail(ctx, e);
  return 0;
}

int jxlb200_encode(jxlb200_ctx* ctx, const jxlb200_image* image, const jxlb200_params* params, uint8_t** out,
                   size_t* out_len, jxlb200_stats* stats) {
  if (!ctx) return -1;
  if (!image || !image->pixels) return fail(ctx, "invalid image");
  if (!out || !out_len) return fail(ctx, "null output");
  if (int rc = check_params(ctx, image->width, image->height, params)) return rc;
  if (image->stride < (size_t)3 * image->width) return fail(ctx, "stride smaller than 3*width");
  std::string e;
  EncodeParams ep{params->distance, params->effort, params->proposal, params->flags};
  ctx->enc.set_ans_groups_per_warp(1);
  ctx->enc.set_ans_warps(ctx->ans_warps_single);
  ctx->enc.set_copy_stream(nullptr);
  ctx->enc.set_fork(ctx->fork_single);
  if (!ctx->enc.EncodeHost(image->pixels, (int)image->width, (int)image->height, image->stride, ep, stats, &e))
    return fail(ctx, e);
  if (!ctx->enc.Fetch(out, out_len, &e)) return fail(ctx, e);
  return 0;
}

int jxlb200_encode_batch(jxlb200_ctx* ctx, const jxlb200_image* images, const jxlb200_params* params, size_t n,
                         uint8_t** outs, size_t* out_lens, jxlb200_stats* stats) {
  if (!ctx) return -1;
  if (!images || !params || !outs || !out_lens) return fail(ctx, "null argument");
  for (size_t i = 0; i < n; ++i) {
    outs[i] = nullptr; out_lens[i] = 0;
    if (!images[i].pixels) return fail(ctx, "invalid image");
    if (int rc = check_params(ctx, images[i].width, images[i].height, &params[i])) return rc;
    if (images[i].stride < (size_t)3 * images[i].width) return fail(ctx, "stride smaller than 3*width");
  }
  const int P = (int)std::min<size_t>((size_t)ctx->num_pipelines, n ? n : 1);
  if (!ensure_pipelines(ctx, P)) return -1;
  if (ctx->use_copy_stream && P > 1 && !ctx->copy_stream) {
    cudaSetDevice(ctx->device);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) ctx->copy_stream = nullptr;
  }
  std::vector<long> owner(P, -1);   // image currently in flight on each pipeline
  std::string e;
  int rc = 0;
  auto retire = [&](int p) {
    const long i = owner[p];
    if (i < 0) return;
    owner[p] = -1;
    if (!ctx->pipe(p)->Finish(stats ? &stats[i] : nullptr, &e) || !ctx->pipe(p)->Fetch(&outs[i], &out_lens[i], &e)) rc = fail(ctx, e);
  };
  for (size_t i = 0; i < n && rc == 0; ++i) {
    const int p = (int)(i % P);
    retire(p);
    if (rc) break;
    EncodeParams ep{params[i].distance, params[i].effort, params[i].proposal, params[i].flags};
    ctx->pipe(p)->set_ans_groups_per_warp(P > 1 ? ctx->ans_gpw : 1);
    ctx->pipe(p)->set_ans_warps(P > 1 ? ctx->ans_warps : ctx->ans_warps_single);
    ctx->pipe(p)->set_copy_stream(P > 1 && ctx->use_copy_stream ? ctx->copy_stream : nullptr);
    ctx->pipe(p)->set_fork(P == 1 && ctx->fork_single);
    if (!ctx->pipe(p)->EnqueueHost(images[i].pixels, (int)images[i].width, (int)images[i].height, images[i].stride, ep, &e)) {
      rc = fail(ctx, e);
      break;
    }
    owner[p] = (long)i;
  }
  for (int k = 0; k < P; ++k) retire((int)((n + k) % P));   // oldest first
  if (rc) for (size_t i = 0; i < n; ++i) { free(outs[i]); outs[i] = nullptr; out_lens[i] = 0; }
  return rc;
}

int jxlb200_encode_batch_device(jxlb200_ctx* ctx, const uint8_t* const* d_pixels, const uint32_t* widths,
                                const uint32_t* heights, const size_t* strides, const jxlb200_params* params, size_t n,
                                jxlb200_stats* stats, float* device_ms) {
  if (!ctx) return -1;
  if (!d_pixels || !widths || !heights || !strides || !params) return fail(ctx, "null argument");
  for (size_t i = 0; i < n; ++i) {
    if (!d_pixels[i]) return fail(ctx, "invalid image");
    if (int rc = check_params(ctx, widths[i], heights[i], &params[i])) return rc;
    if (strides[i] < (size_t)3 * widths[i]) return fail(ctx, "stride smaller than 3*width");
  }
  const int P = (int)std::min<size_t>((size_t)ctx->num_pipelines, n ? n : 1);
  if (!ensure_pipelines(ctx, P)) return -1;
  std::vector<long> owner(P, -1);
  std::string e;
  int rc = 0;
  double enqueue_us = 0.0;
  // device time of the whole batch: every pipeline starts after `t0` and `t1` waits for all of them
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (device_ms) {
    cudaSetDevice(ctx->device);
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, ctx->pipe(0)->stream());
    for (int p = 1; p < P; ++p) cudaStreamWaitEvent(ctx->pipe(p)->stream(), t0, 0);
  }
  auto retire = [&](int p) {
    const long i = owner[p];
    if (i < 0) return;
    owner[p] = -1;
    if (!ctx->pipe(p)->Finish(stats ? &stats[i] : nullptr, &e)) rc = fail(ctx, e);
  };
  for (size_t i = 0; i < n && rc == 0; ++i) {
    const int p = (int)(i % P);
    retire(p);
    if (rc) break;
    EncodeParams ep{params[i].distance, params[i].effort, params[i].proposal, params[i].flags};
    ctx->pipe(p)->set_ans_groups_per_warp(P > 1 ? ctx->ans_gpw : 1);
    ctx->pipe(p)->set_ans_warps(P > 1 ? ctx->ans_warps : ctx->ans_warps_single);
    ctx->pipe(p)->set_fork(P == 1 && ctx->fork_single);
    const auto tq0 = std::chrono::steady_clock::now();
    if (!ctx->pipe(p)->EnqueueDevice(d_pixels[i], (int)widths[i], (int)heights[i], strides[i], ep, &e)) { rc = fail(ctx, e); break; }
    enqueue_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tq0).count();
    owner[p] = (long)i;
  }
  if (device_ms) {
    std::vector<cudaEvent_t> done(P, nullptr);
    for (int p = 1; p < P; ++p) {
      cudaEventCreateWithFlags(&done[p], cudaEventDisableTiming);
      cudaEventRecord(done[p], ctx->pipe(p)->stream());
      cudaStreamWaitEvent(ctx->pipe(0)->stream(), done[p], 0);
    }
    cudaEventRecord(t1, ctx->pipe(0)->stream());
    cudaEventSynchronize(t1);
    *device_ms = 0.0f;
    cudaEventElapsedTime(device_ms, t0, t1);
    for (int p = 1; p < P; ++p) cudaEventDestroy(done[p]);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
  }
  for (int k = 0; k < P; ++k) retire((int)((n + k) % P));
  if (getenv("JXLB200_DEBUG")) fprintf(stderr, "[jxlb200] batch of %zu on %d pipelines: host enqueue %.1f us per image\n", n, P, enqueue_us / (double)(n ? n : 1));
  return rc;
}

int jxlb200_set_pipelines(jxlb200_ctx* ctx, int n) {
  if (!ctx) return -1;
  if (n < 1 || n > 64) return fail(ctx, "pipelines out of range [1, 64]");
  ctx->num_pipelines = n;
  return 0;
}

void jxlb200_free(void* buf) { free(buf); }

int64_t jxlb200_dump(jxlb200_ctx* ctx, int stage, void* dst, size_t cap) {
  if (!ctx) return -1;
  std::string e;
  const int64_t r = ctx->enc.Dump(stage, dst, cap, &e);
  if (r < 0) ctx->err = e;
  return r;
}

int jxlb200_debug_set_strategy_map(jxlb200_ctx* ctx, const uint8_t* acs, uint32_t bxs, uint32_t bys) {
  if (!ctx) return -1;
  if (!acs || bxs == 0 || bys == 0 || (uint64_t)bxs * bys > (1ull << 22)) return fail(ctx, "invalid strategy map");
  std::string e;
  ctx->forced_acs.assign(acs, acs + (size_t)bxs * bys);
  ctx->forced_bxs = (int)bxs; ctx->forced_bys = (int)bys;
  for (int p = 0; p <= (int)ctx->extra.size(); ++p)
    if (!ctx->pipe(p)->SetForcedAcs(acs, (int)bxs, (int)bys, &e)) return fail(ctx, e);
  return 0;
}

int jxlb200_debug_homogeneity(jxlb200_ctx* ctx, const float* x, const float* y, const float* b, uint32_t stride, uint32_t ysize,
                              float distance, float* out) {
  if (!ctx) return -1;
  if (!x || !y || !b || !out) return fail(ctx, "null plane");
  if (stride < 8 || ysize < 8 || stride > 16384 || ysize > 16384) return fail(ctx, "plane size out of range [8, 16384]");
  std::string e;
  if (!ctx->enc.DebugHomogeneity(x, y, b, (int)stride, (int)ysize, distance, out, &e)) return fail(ctx, e);
  return 0;
}

void jxlb200_dims(uint32_t width, uint32_t height, int32_t* dims) {
  FrameDim fd; fd.Set((int)width, (int)height);
  const int32_t v[16] = {fd.xsize, fd.ysize, fd.xs_pad, fd.ys_pad, fd.pitch, fd.bxs, fd.bys, fd.gxs, fd.gys,
                         fd.num_groups, fd.dgxs, fd.dgys, fd.num_dc_groups, fd.txs, fd.tys, 0};
  memcpy(dims, v, sizeof(v));
}

}  // extern "C"
