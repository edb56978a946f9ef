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

using namespace jxlb;

struct jxlb200_ctx {
  Encoder enc;
  std::string err;
};

static int fail(jxlb200_ctx* ctx, const std::string& msg, int code = -1) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" {

int jxlb200_abi_version(void) { return JXLB200_ABI_VERSION; }

jxlb200_ctx* jxlb200_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return nullptr;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
  if (prop.major != 10) return nullptr;  // sm_100a only; there is no fallback path
  jxlb200_ctx* ctx = new jxlb200_ctx();
  std::string e;
  if (!ctx->enc.Init(device, &e)) { delete ctx; return nullptr; }
  return ctx;
}

void jxlb200_destroy(jxlb200_ctx* ctx) {
  if (!ctx) return;
  ctx->enc.Destroy();
  delete ctx;
}

const char* jxlb200_last_error(const jxlb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int check_params(jxlb200_ctx* ctx, uint32_t w, uint32_t h, const jxlb200_params* p) {
  if (!p) return fail(ctx, "null params");
  if (w == 0 || h == 0) return fail(ctx, "invalid image");
  if ((uint64_t)w * h > (1ull << 28)) return fail(ctx, "image too large");
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
  if (!ctx->enc.EncodeHost(image->pixels, (int)image->width, (int)image->height, image->stride, ep, stats, &e))
    return fail(ctx, e);
  if (!ctx->enc.Fetch(out, out_len, &e)) return fail(ctx, e);
  return 0;
}

int jxlb200_encode_batch(jxlb200_ctx* ctx, const jxlb200_image* images, const jxlb200_params* params, size_t n,
                         uint8_t** outs, size_t* out_lens, jxlb200_stats* stats) {
  if (!ctx) return -1;
  for (size_t i = 0; i < n; ++i) {
    const int rc = jxlb200_encode(ctx, &images[i], &params[i], &outs[i], &out_lens[i], stats ? &stats[i] : nullptr);
    if (rc) return rc;
  }
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

void jxlb200_dims(uint32_t width, uint32_t height, int32_t* dims) {
  FrameDim fd; fd.Set((int)width, (int)height);
  const int32_t v[16] = {fd.xsize, fd.ysize, fd.xs_pad, fd.ys_pad, fd.pitch, fd.bxs, fd.bys, fd.gxs, fd.gys,
                         fd.num_groups, fd.dgxs, fd.dgys, fd.num_dc_groups, fd.txs, fd.tys, 0};
  memcpy(dims, v, sizeof(v));
}

}  // extern "C"
