/* A plain C99 client of include/jxlb200.h: what a cgo / Rust -sys / JNI binding sees.  Built and run by
 * tests/test_cabi_symbols.py: without a GPU jxlb200_create returns NULL (there is no CPU fallback) and the program
 * reports that; with one it encodes a small synthetic image through jxlb200_encode and checks the contract
 * (0 / negative return codes, library-owned output released with jxlb200_free, error text, stats). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jxlb200.h"

int main(void) {
  if (jxlb200_abi_version() != JXLB200_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
  int32_t dims[16];
  jxlb200_dims(100, 60, dims);
  if (dims[0] != 100 || dims[1] != 60 || dims[2] != 104 || dims[3] != 64) { fprintf(stderr, "jxlb200_dims\n"); return 2; }
  jxlb200_ctx* ctx = jxlb200_create(0);
  if (!ctx) { printf("no sm_100 device: jxlb200_create returned NULL\n"); return 0; }
  const uint32_t w = 100, h = 60;
  uint8_t* px = (uint8_t*)malloc((size_t)3 * w * h);
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      px[3 * (y * w + x) + 0] = (uint8_t)(x * 2);
      px[3 * (y * w + x) + 1] = (uint8_t)(y * 4);
      px[3 * (y * w + x) + 2] = (uint8_t)(128 + (((x / 8) ^ (y / 8)) & 1 ? 60 : -60));
    }
  jxlb200_image img = {px, w, h, (size_t)3 * w};
  jxlb200_params par = {1.0f, 7, JXLB200_PROPOSAL_COMBINED, JXLB200_FLAG_QUALITY};
  uint8_t* out = NULL;
  size_t out_len = 0;
  jxlb200_stats st;
  int rc = jxlb200_encode(ctx, &img, &par, &out, &out_len, &st);
  if (rc != 0) { fprintf(stderr, "encode failed: %s\n", jxlb200_last_error(ctx)); return 3; }
  if (out_len < 16 || out[0] != 0xFF || out[1] != 0x0A) { fprintf(stderr, "not a JPEG XL codestream\n"); return 3; }
  if (st.codestream_bytes != out_len || st.width != w || st.height != h || !st.quality_valid || !(st.psnr > 20.0)) {
    fprintf(stderr, "stats: bytes %lu / %lu, %ux%u, quality_valid %u, psnr %.2f\n", (unsigned long)st.codestream_bytes,
            (unsigned long)out_len, st.width, st.height, st.quality_valid, st.psnr);
    return 3;
  }
  jxlb200_free(out);
  /* the reference's skip-and-continue contract: a bad request fails with a message, the context stays usable */
  par.distance = 100.0f;
  rc = jxlb200_encode(ctx, &img, &par, &out, &out_len, &st);
  if (rc >= 0 || strlen(jxlb200_last_error(ctx)) == 0) { fprintf(stderr, "bad distance accepted\n"); return 3; }
  par.distance = 2.0f;
  rc = jxlb200_encode(ctx, &img, &par, &out, &out_len, &st);
  if (rc != 0) { fprintf(stderr, "context unusable after an error\n"); return 3; }
  jxlb200_free(out);
  printf("ok: %zu bytes, %.3f bpp, %.2f dB\n", (size_t)st.codestream_bytes, st.bpp, st.psnr);
  jxlb200_destroy(ctx);
  free(px);
  return 0;
}
