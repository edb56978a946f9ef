// jxlb200 — pipeline driver (libjxl enc_frame.cc / enc_heuristics.cc order of operations,
// SURVEY.md section 3.2 [UPSTREAM]): XYB -> quant field -> homogeneity map -> AC strategy ->
// transform + quantise -> tokens -> histograms -> ANS -> frame assembly.  All stages are
// kernel launches on one stream; the host only sequences them.
#include "encoder.h"
#include "host_tables.h"
#include "kernels.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace jxlb {

thread_local unsigned g_kernel_launches = 0;

#define CUDA_OK(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) { *err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } \
  } while (0)

bool Encoder::Init(int device, std::string* err) {
  device_ = device;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  for (auto& e : ev_) CUDA_OK(cudaEventCreate(&e));
  float lut[256];
  host_srgb_lut(lut);
  if (!d_lut_.Reserve(256)) { *err = "alloc"; return false; }
  CUDA_OK(cudaMemcpy(d_lut_.p, lut, sizeof(lut), cudaMemcpyHostToDevice));
  for (int k = 0; k < 17; ++k) {
    std::vector<float> w;
    host_quant_weights(k, &w);
    if (w.empty()) continue;
    std::vector<float> dq(w.size());
    for (size_t i = 0; i < w.size(); ++i) dq[i] = 1.0f / w[i];
    if (!d_weights_[k].Reserve(w.size()) || !d_dequant_[k].Reserve(w.size())) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_weights_[k].p, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dequant_[k].p, dq.data(), dq.size() * 4, cudaMemcpyHostToDevice));
  }
  {
    std::vector<uint16_t> order;
    host_natural_order(0, &order);
    uint8_t izz[64];
    for (int k = 0; k < 64; ++k) izz[order[k]] = (uint8_t)k;
    if (!d_izz8_.Reserve(64)) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_izz8_.p, izz, 64, cudaMemcpyHostToDevice));
  }
  if (!d_cvx_.Reserve(27) || !d_cvy_.Reserve(27) || !d_q_.Reserve(1)) { *err = "alloc"; return false; }
  CUDA_OK(cudaMemcpy(d_cvx_.p, kCoveredX, 27, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_cvy_.p, kCoveredY, 27, cudaMemcpyHostToDevice));
  return true;
}

void Encoder::Destroy() {
  if (device_ < 0) return;
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  d_lut_.Release();
  for (int k = 0; k < 17; ++k) { d_weights_[k].Release(); d_dequant_[k].Release(); }
  d_izz8_.Release(); d_cvx_.Release(); d_cvy_.Release();
  d_rgb_.Release(); d_xyb_.Release(); d_mask1x1_.Release(); d_pre_.Release(); d_qf_.Release(); d_mask_.Release();
  d_homog_.Release(); d_acs_entropy_.Release(); d_acs_.Release(); d_raw_qf_.Release(); d_cmap_.Release();
  d_coeffs_.Release(); d_dc_quant_.Release(); d_nzeros_.Release(); d_lastpos_.Release(); d_q_.Release();
  if (h_pinned_) cudaFreeHost(h_pinned_);
  h_pinned_ = nullptr;
  for (auto& e : ev_) if (e) cudaEventDestroy(e);
  if (stream_) cudaStreamDestroy(stream_);
  stream_ = nullptr; device_ = -1;
}

bool Encoder::Reserve(const FrameDim& fd, std::string* err) {
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  bool ok = d_xyb_.Reserve(3 * plane) && d_mask1x1_.Reserve(plane) && d_pre_.Reserve(plane / 16 + 1) &&
            d_qf_.Reserve(nblk) && d_mask_.Reserve(nblk) && d_homog_.Reserve(3 * nblk) &&
            d_acs_entropy_.Reserve(nblk) && d_acs_.Reserve(nblk) && d_raw_qf_.Reserve(nblk) &&
            d_cmap_.Reserve((size_t)2 * fd.txs * fd.tys) &&
            d_coeffs_.Reserve((size_t)fd.num_groups * kGroupBlocks * 192) && d_dc_quant_.Reserve(3 * nblk) &&
            d_nzeros_.Reserve(3 * nblk) && d_lastpos_.Reserve(3 * nblk);
  if (!ok) { *err = "device allocation failed"; return false; }
  return true;
}

bool Encoder::EncodeHost(const uint8_t* pixels, int w, int h, size_t stride, const EncodeParams& p, jxlb200_stats* stats,
                         std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  const size_t row = (size_t)3 * w;
  const size_t bytes = row * h;
  if (!d_rgb_.Reserve(bytes + 16)) { *err = "device allocation failed"; return false; }
  if (bytes > h_pinned_cap_) {
    if (h_pinned_) cudaFreeHost(h_pinned_);
    h_pinned_ = nullptr; h_pinned_cap_ = 0;
    CUDA_OK(cudaMallocHost(&h_pinned_, bytes));
    h_pinned_cap_ = bytes;
  }
  CUDA_OK(cudaEventRecord(ev_[0], stream_));
  // pack rows into the pinned staging buffer (drops any row padding), then one async copy
  if (stride == row) memcpy(h_pinned_, pixels, bytes);
  else for (int y = 0; y < h; ++y) memcpy(h_pinned_ + (size_t)y * row, pixels + (size_t)y * stride, row);
  CUDA_OK(cudaMemcpyAsync(d_rgb_.p, h_pinned_, bytes, cudaMemcpyHostToDevice, stream_));
  fd_.Set(w, h);
  return Run(d_rgb_.p, row, p, stats, true, err);
}

bool Encoder::EncodeDevice(const uint8_t* d_pixels, int w, int h, size_t stride, const EncodeParams& p,
                           jxlb200_stats* stats, std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaEventRecord(ev_[0], stream_));
  fd_.Set(w, h);
  return Run(d_pixels, stride, p, stats, false, err);
}

bool Encoder::Run(const uint8_t* d_rgb, size_t stride, const EncodeParams& p, jxlb200_stats* stats, bool h2d_timed,
                  std::string* err) {
  (void)h2d_timed;
  const FrameDim& fd = fd_;
  if (!Reserve(fd, err)) return false;
  params_ = p;
  have_frame_ = false;
  g_kernel_launches = 0;
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  float* X = d_xyb_.p; float* Y = X + plane; float* B = Y + plane;
  {
    int xq = 2;
    if (p.distance > 1.25f) xq++;
    if (p.distance > 9.0f) xq++;
    if (p.distance < 0.299f) xq++;
    x_qm_scale_ = xq; b_qm_scale_ = 2;
    x_qm_mul_ = powf(1.25f, (float)(xq - 2));
    b_qm_mul_ = 1.0f;
  }
  CUDA_OK(cudaEventRecord(ev_[1], stream_));
  // K1: XYB
  launch_rgb8_to_xyb(d_rgb, stride, fd.xsize, fd.ysize, fd, d_lut_.p, X, Y, B, stream_);
  CUDA_OK(cudaEventRecord(ev_[2], stream_));
  // K2: quant field
  if (p.flags & JXLB200_FLAG_UNIFORM_QF) {
    launch_fill(d_qf_.p, nblk, 0.841f / p.distance, stream_);
    CUDA_OK(cudaMemsetAsync(d_mask_.p, 0, nblk * 4, stream_));
    CUDA_OK(cudaMemsetAsync(d_mask1x1_.p, 0, plane * 4, stream_));
  } else {
    launch_aq(X, Y, B, fd, p.distance, d_mask1x1_.p, d_pre_.p, d_qf_.p, d_mask_.p, stream_);
  }
  launch_quant_params(d_qf_.p, nblk, host_initial_quant_dc(p.distance), d_q_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[3], stream_));
  // K4: homogeneity map (the thesis' proposals)
  launch_homogeneity(X, Y, B, fd, p.distance, d_homog_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[4], stream_));
  // K6: AC strategy (fixed DCT8 until the search kernel lands)
  CUDA_OK(cudaMemsetAsync(d_acs_.p, 0x80, nblk, stream_));
  CUDA_OK(cudaMemsetAsync(d_acs_entropy_.p, 0, nblk * 4, stream_));
  CUDA_OK(cudaMemsetAsync(d_cmap_.p, 0, (size_t)2 * fd.txs * fd.tys, stream_));
  launch_raw_qf(d_qf_.p, d_acs_.p, fd, d_q_.p, d_cvx_.p, d_cvy_.p, d_raw_qf_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[5], stream_));
  // K7: transform + quantise
  launch_dct8_quant(X, Y, B, fd, d_q_.p, d_weights_[0].p, d_dequant_[0].p + 64, d_izz8_.p, d_cmap_.p, x_qm_mul_,
                    b_qm_mul_, p.effort >= 5 ? 1 : 0, d_raw_qf_.p, d_coeffs_.p, d_dc_quant_.p, d_nzeros_.p,
                    d_lastpos_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[6], stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  CUDA_OK(cudaGetLastError());
  have_frame_ = true;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->width = fd.xsize; stats->height = fd.ysize;
    stats->num_groups = fd.num_groups; stats->num_dc_groups = fd.num_dc_groups;
    QuantDev q;
    CUDA_OK(cudaMemcpy(&q, d_q_.p, sizeof(q), cudaMemcpyDeviceToHost));
    stats->kernel_launches = g_kernel_launches;
    stats->global_scale = q.global_scale; stats->quant_dc = q.quant_dc;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_[0], ev_[1]); stats->stage_ms[JXLB200_T_H2D] = ms;
    cudaEventElapsedTime(&ms, ev_[1], ev_[2]); stats->stage_ms[JXLB200_T_XYB] = ms;
    cudaEventElapsedTime(&ms, ev_[2], ev_[3]); stats->stage_ms[JXLB200_T_AQ] = ms;
    cudaEventElapsedTime(&ms, ev_[3], ev_[4]); stats->stage_ms[JXLB200_T_HOMOG] = ms;
    cudaEventElapsedTime(&ms, ev_[4], ev_[5]); stats->stage_ms[JXLB200_T_ACS] = ms;
    cudaEventElapsedTime(&ms, ev_[5], ev_[6]); stats->stage_ms[JXLB200_T_COEFF] = ms;
    cudaEventElapsedTime(&ms, ev_[0], ev_[6]); stats->total_ms = ms;
  }
  return true;
}

bool Encoder::Fetch(uint8_t** out, size_t* out_len, std::string* err) {
  if (!have_frame_) { *err = "no encoded frame"; return false; }
  *out = (uint8_t*)malloc(1);
  *out_len = 0;
  return true;
}

int64_t Encoder::Dump(int stage, void* dst, size_t cap, std::string* err) {
  if (!have_frame_) { *err = "no encoded frame"; return -1; }
  cudaSetDevice(device_);
  const FrameDim& fd = fd_;
  const size_t plane = (size_t)fd.ys_pad * fd.pitch;
  const size_t nblk = (size_t)fd.bxs * fd.bys;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (stage) {
    case JXLB200_STAGE_XYB: src = d_xyb_.p; bytes = 3 * plane * 4; break;
    case JXLB200_STAGE_QF_FLOAT: src = d_qf_.p; bytes = nblk * 4; break;
    case JXLB200_STAGE_MASK1X1: src = d_mask1x1_.p; bytes = plane * 4; break;
    case JXLB200_STAGE_MASK: src = d_mask_.p; bytes = nblk * 4; break;
    case JXLB200_STAGE_HOMOG: src = d_homog_.p; bytes = 3 * nblk * 4; break;
    case JXLB200_STAGE_ACS: src = d_acs_.p; bytes = nblk; break;
    case JXLB200_STAGE_ACS_ENTROPY: src = d_acs_entropy_.p; bytes = nblk * 4; break;
    case JXLB200_STAGE_RAW_QF: src = d_raw_qf_.p; bytes = nblk * 4; break;
    case JXLB200_STAGE_CMAP: src = d_cmap_.p; bytes = (size_t)2 * fd.txs * fd.tys; break;
    case JXLB200_STAGE_COEFFS: src = d_coeffs_.p; bytes = (size_t)fd.num_groups * kGroupBlocks * 192 * 2; break;
    case JXLB200_STAGE_DC_QUANT: src = d_dc_quant_.p; bytes = 3 * nblk * 2; break;
    case JXLB200_STAGE_NZEROS: src = d_nzeros_.p; bytes = 3 * nblk; break;
    case JXLB200_STAGE_QUANT_PARAMS: {
      QuantDev q;
      if (cudaMemcpy(&q, d_q_.p, sizeof(q), cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
      const int32_t v[4] = {q.global_scale, q.quant_dc, x_qm_scale_, b_qm_scale_};
      if (dst && cap >= sizeof(v)) memcpy(dst, v, sizeof(v));
      return sizeof(v);
    }
    default: *err = "unknown stage"; return -1;
  }
  if (dst && cap >= bytes && bytes) {
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
  }
  return (int64_t)bytes;
}

}  // namespace jxlb
