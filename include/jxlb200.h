/* jxlb200 — C ABI of the B200-native JPEG XL lossy (VarDCT) encode hot path.
 *
 * Drop-in boundary for the encoder call of the thesis harness
 * pscoro/JPEG-XL-Lossy-Image-Compression-Thesis.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * The reference reaches the encoder through exactly one call:
 *   DockerManager::execute_cjxl(input_file, output_file, distance: f64, effort: u32)
 *     benchmark-jpegxl/src/docker_manager.rs:100-137
 *   -> `docker exec ... /libjxl/build/tools/cjxl <in> <out> --distance=<d> --effort=<e>`
 * called once per (image, distance, effort) from
 *   JXLCompressionBenchmark::run   benchmark-jpegxl/src/benchmark.rs:654-660
 * with failures mapped to "skip" (benchmark.rs:661-677).  Which proposal is active
 * is not an argument there: it is "which proposals/NAME.diff was applied before libjxl
 * was rebuilt" (benchmark.rs:460-484, docker_manager.rs:303-368).  Here it is a
 * runtime enum.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; 0 = success,
 * negative = error (text via jxlb200_last_error); nothing throws or aborts across
 * the boundary.  A context is owned by one thread (one CUDA stream + device arenas);
 * distinct contexts may be used concurrently (the reference runs up to
 * num_workers = 6 threads, benchmark-jpegxl/src/config.rs:22).
 */
#ifndef JXLB200_H_
#define JXLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JXLB200_ABI_VERSION 3

typedef struct jxlb200_ctx jxlb200_ctx;

/* Which libjxl patch of the thesis is emulated (proposals/NAME.diff). */
enum {
  JXLB200_PROPOSAL_NONE = 0,             /* unpatched libjxl ("main", context.rs:17)           */
  JXLB200_PROPOSAL_PARTITIONING = 1,     /* proposals/homogeneity-partitioning.diff:272-276     */
  JXLB200_PROPOSAL_FACTORED_ENTROPY = 2, /* proposals/homogeneity-factored-entropy.diff:248-253 */
  JXLB200_PROPOSAL_COMBINED = 3          /* proposals/combined.diff:248-253,270-274             */
};

/* flags */
#define JXLB200_FLAG_FIXED_DCT8 1u /* skip the AC-strategy search: DCT8 everywhere (BASELINE config 2) */
#define JXLB200_FLAG_UNIFORM_QF 2u /* skip the adaptive quant field: qf = 0.841/distance everywhere   */
#define JXLB200_FLAG_FORCED_ACS 8u /* use the map of jxlb200_debug_set_strategy_map instead of the search (jxlb200_encode only)  */
#define JXLB200_FLAG_QUALITY 4u    /* also reconstruct the coded frame on the device and fill stats.sse / stats.psnr  */
#define JXLB200_FLAG_GABORISH 16u  /* Gaborish loop filter: signal the decoder's default 3x3 blur and sharpen the XYB planes with its
                                      5x5 least-squares inverse before the search (libjxl's default is on, with its own tuned
                                      kernel; off here unless asked for: SURVEY 8a row U3)                              */
#define JXLB200_FLAG_CFL 32u       /* chroma from luma: fit ytox / ytob per 64x64 tile (ridge least squares on the weighted DCT8
                                      coefficients; libjxl's own fit is not available offline) instead of the all-zero map  */

/* Input image: 8-bit sRGB, interleaved RGB, row-major (what the harness hands to cjxl
 * as a PNG; image_reader.rs ColorType::Rgb8).  `stride` is bytes per row (>= 3*width). */
typedef struct {
  const uint8_t* pixels;
  uint32_t width;
  uint32_t height;
  size_t stride;
} jxlb200_image;

/* cjxl's `--distance` / `--effort` (docker_manager.rs:125-127) plus the proposal toggle. */
typedef struct {
  float distance;    /* Butteraugli distance, 0.01 .. 25                                   */
  uint32_t effort;   /* 1 .. 9; the harness sweeps 5..=9 (benchmark.rs:638)                */
  uint32_t proposal; /* JXLB200_PROPOSAL_*                                                 */
  uint32_t flags;    /* JXLB200_FLAG_*                                                     */
} jxlb200_params;

/* What the device already knows after an encode.  bpp = 8*bytes/(w*h), the quantity the
 * harness derives as 24 / raw_file_size_ratio (benchmark.rs:921). */
typedef struct {
  uint64_t codestream_bytes;
  double bpp;
  uint32_t width, height;
  uint32_t num_groups, num_dc_groups;
  uint32_t global_scale, quant_dc;
  uint64_t num_tokens;
  uint32_t num_clusters;
  uint32_t acs_histogram[27];   /* first blocks per AcStrategy code               */
  float stage_ms[16];           /* CUDA-event time per pipeline stage (JXLB200_T_*) */
  float total_ms;               /* device time H2D .. D2H                          */
  uint32_t kernel_launches;     /* CUDA kernels launched by this encode            */
  /* JXLB200_FLAG_QUALITY: squared error of the decoded 8-bit sRGB image against the input, summed per channel
   * (R, G, B), and the PSNR over all samples — calculate_mse / calculate_psnr of the harness
   * (benchmark-jpegxl/src/image_reader.rs:555-606) without the djxl round trip. */
  uint32_t quality_valid;
  uint64_t sse[3];
  double psnr;
} jxlb200_stats;

enum {
  JXLB200_T_H2D = 0, JXLB200_T_XYB = 1, JXLB200_T_AQ = 2, JXLB200_T_HOMOG = 3, JXLB200_T_ACS = 4,
  JXLB200_T_COEFF = 5, JXLB200_T_TOKENIZE = 6, JXLB200_T_HISTO = 7, JXLB200_T_ANS = 8,
  JXLB200_T_DC = 9, JXLB200_T_ASSEMBLE = 10, JXLB200_T_D2H = 11, JXLB200_T_QUALITY = 12
};

/* Intermediate taps for parity tests (same ids in oracle/jxo_frame.h). */
enum {
  JXLB200_STAGE_XYB = 1,          /* f32 [3][ys_pad][pitch]                          */
  JXLB200_STAGE_QF_FLOAT = 2,     /* f32 [bys][bxs]                                  */
  JXLB200_STAGE_MASK1X1 = 3,      /* f32 [ys_pad][pitch]                             */
  JXLB200_STAGE_HOMOG = 4,        /* f32 [bys][bxs][3] = r_h, r_v, r_d               */
  JXLB200_STAGE_ACS = 5,          /* u8  [bys][bxs] raw strategy | 0x80 first block  */
  JXLB200_STAGE_RAW_QF = 6,       /* i32 [bys][bxs]                                  */
  JXLB200_STAGE_QUANT_PARAMS = 7, /* i32 [4] global_scale, quant_dc, x_qm, b_qm      */
  JXLB200_STAGE_COEFFS = 8,       /* i16 [groups][1024][3 (Y,X,B)][64] scan order    */
  JXLB200_STAGE_DC_QUANT = 9,     /* i16 [3 (X,Y,B)][bys][bxs]                       */
  JXLB200_STAGE_NZEROS = 10,      /* u8  [3 (X,Y,B)][bys][bxs]                       */
  JXLB200_STAGE_TOKENS = 11,      /* u32 (ctx << 16 | value), all groups             */
  JXLB200_STAGE_HISTOGRAMS = 12,  /* u32 [num_ctx][alphabet]                         */
  JXLB200_STAGE_CONTEXT_MAP = 13, /* u8  [num_ctx]                                   */
  JXLB200_STAGE_GROUP_STREAMS = 14,/* u8 concatenated AC group sections              */
  JXLB200_STAGE_CODESTREAM = 15,  /* u8  the .jxl codestream                         */
  JXLB200_STAGE_MASK = 16,        /* f32 [bys][bxs]                                  */
  JXLB200_STAGE_CMAP = 17,        /* i8  [2][tys][txs]                               */
  JXLB200_STAGE_TOKEN_OFFSETS = 18,/* u32 [groups+1]                                 */
  JXLB200_STAGE_GROUP_OFFSETS = 19,/* u32 [groups+1] byte offsets                    */
  JXLB200_STAGE_ACS_ENTROPY = 20  /* f32 [bys][bxs]                                  */
};

/* Lifecycle.  Replaces DockerManager::new/setup/teardown (docker_manager.rs:184-216, :261):
 * no container, a CUDA device ordinal instead.  Returns NULL when no usable sm_100 device
 * exists — there is no CPU fallback. */
jxlb200_ctx* jxlb200_create(int device);
void jxlb200_destroy(jxlb200_ctx* ctx);
const char* jxlb200_last_error(const jxlb200_ctx* ctx);
int jxlb200_abi_version(void);

/* Replaces DockerManager::execute_cjxl (docker_manager.rs:100-137) + retrieve_file (:72-87):
 * host image in, library-owned codestream out (release with jxlb200_free). */
int jxlb200_encode(jxlb200_ctx* ctx, const jxlb200_image* image, const jxlb200_params* params,
                   uint8_t** out, size_t* out_len, jxlb200_stats* stats);
void jxlb200_free(void* buf);

/* Same encode with the RGB8 image already resident in device memory (`d_pixels` is a
 * device pointer); the codestream stays on the device until jxlb200_fetch().  Used for the
 * HBM-resident `value` measurement and by callers that produce pixels on the GPU. */
int jxlb200_encode_device(jxlb200_ctx* ctx, const uint8_t* d_pixels, uint32_t width, uint32_t height,
                          size_t stride, const jxlb200_params* params, jxlb200_stats* stats);
int jxlb200_fetch(jxlb200_ctx* ctx, uint8_t** out, size_t* out_len);

/* Batch form of the per-image loop of JXLCompressionBenchmark::run (benchmark.rs:637-660):
 * n images, n parameter sets, n outputs; up to jxlb200_set_pipelines() images are in flight at once,
 * each on its own CUDA stream.  Page-locked (pinned) input buffers are copied without staging. */
int jxlb200_encode_batch(jxlb200_ctx* ctx, const jxlb200_image* images, const jxlb200_params* params,
                         size_t n, uint8_t** outs, size_t* out_lens, jxlb200_stats* stats);

/* Batch with the RGB8 images already resident in device memory; the codestreams stay on the device
 * (their sizes are reported in stats[i].codestream_bytes).  Used for the HBM-resident throughput
 * measurement. */
int jxlb200_encode_batch_device(jxlb200_ctx* ctx, const uint8_t* const* d_pixels, const uint32_t* widths,
                                const uint32_t* heights, const size_t* strides, const jxlb200_params* params, size_t n,
                                jxlb200_stats* stats, float* device_ms /* may be NULL: CUDA-event time of the batch */);

/* Number of images the batch entry points keep in flight (default 4, or $JXLB200_PIPELINES): each
 * pipeline owns a CUDA stream and its device arenas, like each of the reference's num_workers = 6
 * worker threads owns a container (benchmark-jpegxl/src/config.rs:22, benchmark.rs:173-198). */
int jxlb200_set_pipelines(jxlb200_ctx* ctx, int n);

/* Copies the intermediate `stage` of the LAST encode on this context to host memory.
 * Returns the stage size in bytes (copying only if cap is large enough), or < 0. */
int64_t jxlb200_dump(jxlb200_ctx* ctx, int stage, void* dst, size_t cap);

/* Parity tap for row U5: the AC-strategy map (bys x bxs bytes, raw strategy | 0x80 on the first block of every transform:
 * the layout of JXLB200_STAGE_ACS) that the next jxlb200_encode / jxlb200_encode_device calls with JXLB200_FLAG_FORCED_ACS
 * code with, instead of searching.  It must be a partition into DCT, IDENTITY, DCT2X2, DCT4X4, DCT4X8, DCT8X4, 16X8, 8X16,
 * 16X16, 32X8, 8X32, 32X16, 16X32, 32X32, 64X32, 32X64, 64X64 transforms inside their 64x64 tiles.  It reaches the transforms
 * the search seldom or never picks (libjxl's merge table does not propose DCT32X8 / DCT8X32). */
int jxlb200_debug_set_strategy_map(jxlb200_ctx* ctx, const uint8_t* acs, uint32_t bxs, uint32_t bys);

/* Parity tap for the reference-pinned rows H1-H7 (proposals/homogeneity-partitioning.diff:17-211): runs the
 * homogeneity kernel on caller-supplied planar XYB (host memory, `stride` floats per row, `ysize` rows — the
 * `config.src_stride` / `config.src_ysize` of the diff, :402-421) and returns r_h, r_v, r_d of every 8x8 block,
 * out[(by * (stride / 8) + bx) * 3 + k].  No counterpart in the reference; it lets the hand-computed vectors of
 * tests/golden/ reach the CUDA kernel with exactly their inputs. */
int jxlb200_debug_homogeneity(jxlb200_ctx* ctx, const float* x, const float* y, const float* b, uint32_t stride,
                              uint32_t ysize, float distance, float* out);

/* Frame geometry helper: dims[16] = xsize ysize xs_pad ys_pad pitch bxs bys gxs gys
 * num_groups dgxs dgys num_dc_groups txs tys 0 */
void jxlb200_dims(uint32_t width, uint32_t height, int32_t* dims);

#ifdef __cplusplus
}
#endif
#endif /* JXLB200_H_ */
