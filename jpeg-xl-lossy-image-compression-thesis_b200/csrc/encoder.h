// jxlb200 — pipeline driver: device arenas + the ordered kernel launches of one encode.
#pragma once
#include "../../include/jxlb200.h"
#include "jxl_common.cuh"
#include "kernels.h"
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
  // asynchronous halves: Enqueue* only launches (H2D copy + kernels) on this encoder's stream, Finish waits
  bool EnqueueHost(const uint8_t* pixels, int w, int h, size_t stride, const EncodeParams& p, std::string* err);
  bool EnqueueDevice(const uint8_t* d_pixels, int w, int h, size_t stride, const EncodeParams& p, std::string* err);
  bool Finish(jxlb200_stats* stats, std::string* err);
  bool Fetch(uint8_t** out, size_t* out_len, std::string* err);
  int64_t Dump(int stage, void* dst, size_t cap, std::string* err);
  // parity tap: homogeneity map of caller-supplied XYB planes (see jxlb200_debug_homogeneity)
  // parity tap: the strategy map the next encodes with JXLB200_FLAG_FORCED_ACS use instead of the search
  bool SetForcedAcs(const uint8_t* acs, int bxs, int bys, std::string* err);
  bool DebugHomogeneity(const float* x, const float* y, const float* b, int stride, int ysize, float distance, float* out, std::string* err);
  cudaStream_t stream() const { return stream_; }
  // 1 = lowest latency (one group per warp); > 1 packs the rANS chains onto fewer SMs (batch throughput)
  void set_ans_groups_per_warp(int n) { ans_groups_per_warp_ = n; }
  void set_ans_warps(int n) { ans_warps_ = n; }
  // batch mode: all pipelines send their input over ONE copy stream (whole images back to back on the H2D engine
  // instead of 32 streams' copies time-sliced against each other); nullptr = copy on the encoder's own stream
  void set_copy_stream(cudaStream_t s) { copy_stream_ = s; }
  // single-image calls: independent stages of one frame run side by side on two auxiliary streams (batch mode keeps one
  // stream per image: the other images fill the GPU and the hardware has 32 queues)
  void set_fork(bool on) { fork_ = on; }

 private:
  bool Reserve(const FrameDim& fd, std::string* err);
  bool Run(const uint8_t* d_rgb, size_t stride, const EncodeParams& p, std::string* err);
  bool in_flight_ = false;
  int dct8_rows_ = 2, dct8_tps_ = 512;   // k_dct8_quant_v4 launch shape
  int ans_groups_per_warp_ = 1, ans_warps_ = 8;
  unsigned launches_ = 0;

  int device_ = -1;
  cudaStream_t stream_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_copy_ = nullptr;
  DevBuf<float2> d_cfl_;              // JXLB200_FLAG_CFL: per-tile (x, b) factors the search reads
  DevBuf<float> d_xyb_gab_;           // JXLB200_FLAG_GABORISH: the sharpened planes every stage after the quant field reads
  const float* xyb_cur_ = nullptr;    // the planes the last encode's search saw (JXLB200_STAGE_XYB tap)
  bool fork_ = false;
  cudaStream_t aux_[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork_ = nullptr, ev_join_[2] = {nullptr, nullptr};
  bool EnsureFork();
  cudaEvent_t ev_[16] = {};   // ev_[12] closes the optional quality stage
  FrameDim fd_{};
  EncodeParams params_{};
  bool have_frame_ = false;
  float x_qm_mul_ = 1.0f, b_qm_mul_ = 1.0f;
  int x_qm_scale_ = 2, b_qm_scale_ = 2;

  // constant tables
  DevBuf<float> d_lut_, d_recon_tab_;
  DevBuf<float> d_weights_[17];
  DevBuf<float> d_dequant_[17];
  DevBuf<float> d_weights_j_[11], d_dequant_j_[11];  // coefficient-stage tables in 16-byte chunks (AcsTables::wJ)
  DevBuf<uint16_t> d_inv_j_[11];
  DevBuf<float> d_weights_c_[6], d_dequant_c_[6];   // 32 / 64-sized search tables in 16-byte chunks (AcsTables::wC)
  DevBuf<float> d_w8_[4], d_dq8_[4];   // 8x8 tables of DCT, DCT4X4, DCT4X8, DCT8X4 in the lane order of the search kernel
  DevBuf<float> d_weights_t_[17], d_dequant_t_[17];   // transposed ([hf][vf] of the wide strategy) for kinds 6, 8, 12
  DevBuf<float> d_bias8_;         // k_dct8_v4: Y dequantisation bias per |q|
  DevBuf<uint8_t> d_lastlut8_;    // k_dct8_v4: last scan index per (lane half, mask byte, byte value)
  DevBuf<uint16_t> d_inv_order_[17];  // per order class: coefficient position -> scan index; [13..16]: classes 4 / 6 / 8 / 5 transposed (wide)
  DevBuf<uint8_t> d_cvx_, d_cvy_;
  // per-frame arenas
  DevBuf<uint8_t> d_rgb_;
  uint8_t* h_pinned_ = nullptr; size_t h_pinned_cap_ = 0;
  std::vector<uint8_t> forced_acs_; int forced_bxs_ = 0, forced_bys_ = 0;
  DevBuf<float> d_xyb_;           // 3 planes
  DevBuf<float> d_mask1x1_, d_pre_, d_qf_, d_mask_, d_homog_, d_acs_entropy_;
  DevBuf<uint8_t> d_acs_;
  DevBuf<int32_t> d_raw_qf_;
  DevBuf<int8_t> d_cmap_;
  DevBuf<int16_t> d_coeffs_, d_dc_quant_;
  DevBuf<uint8_t> d_nzeros_;
  DevBuf<uint16_t> d_nzcount_, d_lastk_;
  DevBuf<QuantDev> d_q_;
  DevBuf<float> d_acs_work_;      // candidate-value tables of the AC-strategy search
  DevBuf<uint32_t> d_acs_jobs_;   // its non-aligned work lists
  DevBuf<uint32_t> d_coeff_lists_;  // first blocks per strategy (k_coeff / k_recon)
  DevBuf<float> d_recon_xyb_;       // reconstructed XYB planes (JXLB200_FLAG_QUALITY only)
  // entropy stage: AC tokens / histograms / clusters / ANS tables / group streams
  DevBuf<int> d_log2lut_;
  DevBuf<uint32_t> d_tokens_, d_token_counts_, d_hist_, d_cluster_hist_, d_hdr_bits_, d_hdr_len_, d_group_arena_;
  DevBuf<uint8_t> d_cluster_state_, d_ctx_map_, d_info_;
  DevBuf<uint16_t> d_norm_, d_rmap_;
  DevBuf<unsigned long long> d_group_start_;
  // modular streams (DC + AC metadata)
  std::vector<DcGroupInfo> h_dgs_;
  uint32_t total_elems_ = 0;
  int tree_ndc_ = -1;
  int dgs_w_ = -1, dgs_h_ = -1;
  DevBuf<DcGroupInfo> d_dgs_;
  DevBuf<int32_t> d_strat_c_, d_qf_c_;
  DevBuf<uint32_t> d_first_count_, d_mod_tokens_, d_mod_hist_, d_lf_words_, d_small_, d_tile_sums_, d_mod_words_, d_dg_start_,
      d_tree_words_;
  DevBuf<uint8_t> d_code_len_;
  DevBuf<uint16_t> d_code_bits_;
  // frame assembly
  DevBuf<uint32_t> d_cm_back_, d_hf_words_, d_hdr_stage_, d_out_;
  DevBuf<Section> d_sections_;
  DevBuf<unsigned long long> d_out_info_;
  unsigned long long* h_out_info_ = nullptr;   // pinned
  size_t codestream_bytes_ = 0;
  int num_clusters_ = 0;
  uint64_t num_tokens_ = 0;
};

}  // namespace jxlb
