// jxlb200 — pipeline driver: device arenas + the ordered kernel launches of one encode.
#pragma once
#include "../../include/jxlb200.h"
#include "jxl_common.cuh"
#include <string>
#include <vector>

namespace jxlb {

struct EncodeParams { float distance; uint32_t effort; uint32_t proposal; uint32_t flags; };

// grow-only device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool Reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return false;
    cap = n;
    return true;
  }
  void Release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

class Encoder {
 public:
  bool Init(int device, std::string* err);
  void Destroy();
  bool EncodeHost(const uint8_t* pixels, int w, int h, size_t stride, const EncodeParams& p, jxlb200_stats* stats,
                  std::string* err);
  bool EncodeDevice(const uint8_t* d_pixels, int w, int h, size_t stride, const EncodeParams& p, jxlb200_stats* stats,
                    std::string* err);
  bool Fetch(uint8_t** out, size_t* out_len, std::string* err);
  int64_t Dump(int stage, void* dst, size_t cap, std::string* err);

 private:
  bool Reserve(const FrameDim& fd, std::string* err);
  bool Run(const uint8_t* d_rgb, size_t stride, const EncodeParams& p, jxlb200_stats* stats, bool h2d_timed,
           std::string* err);

  int device_ = -1;
  cudaStream_t stream_ = nullptr;
  cudaEvent_t ev_[16] = {};
  FrameDim fd_{};
  EncodeParams params_{};
  bool have_frame_ = false;
  float x_qm_mul_ = 1.0f, b_qm_mul_ = 1.0f;
  int x_qm_scale_ = 2, b_qm_scale_ = 2;

  // constant tables
  DevBuf<float> d_lut_;
  DevBuf<float> d_weights_[17];
  DevBuf<float> d_dequant_[17];
  DevBuf<uint8_t> d_izz8_;        // DCT8: position -> scan index
  DevBuf<uint8_t> d_cvx_, d_cvy_;
  // per-frame arenas
  DevBuf<uint8_t> d_rgb_;
  uint8_t* h_pinned_ = nullptr; size_t h_pinned_cap_ = 0;
  DevBuf<float> d_xyb_;           // 3 planes
  DevBuf<float> d_mask1x1_, d_pre_, d_qf_, d_mask_, d_homog_, d_acs_entropy_;
  DevBuf<uint8_t> d_acs_;
  DevBuf<int32_t> d_raw_qf_;
  DevBuf<int8_t> d_cmap_;
  DevBuf<int16_t> d_coeffs_, d_dc_quant_;
  DevBuf<uint8_t> d_nzeros_, d_lastpos_;
  DevBuf<QuantDev> d_q_;
};

}  // namespace jxlb
