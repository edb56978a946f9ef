// jxlb200 — pipeline driver (libjxl enc_frame.cc / enc_heuristics.cc order of operations,
// SURVEY.md section 3.2 [UPSTREAM]): XYB -> quant field -> homogeneity map -> AC strategy ->
// transform + quantise -> tokens -> histograms -> ANS -> frame assembly.  All stages are
// kernel launches on one stream; the host only sequences them.
#include "encoder.h"
#include "host_tables.h"
#include "kernels.h"
#include "entropy.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>

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
  {
    float tab[264];
    host_recon_tables(tab);
    if (!d_recon_tab_.Reserve(264)) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_recon_tab_.p, tab, sizeof(tab), cudaMemcpyHostToDevice));
  }
  for (int k = 0; k < 17; ++k) {
    std::vector<float> w;
    host_quant_weights(k, &w);
    if (w.empty()) continue;
    std::vector<float> dq(w.size());
    for (size_t i = 0; i < w.size(); ++i) dq[i] = 1.0f / w[i];
    if (!d_weights_[k].Reserve(w.size()) || !d_dequant_[k].Reserve(w.size())) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_weights_[k].p, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dequant_[k].p, dq.data(), dq.size() * 4, cudaMemcpyHostToDevice));
    if (k == 6 || k == 7 || k == 8 || k == 12) {
      // lane order of the WIDE strategy of the kind: [c][hf][vf] (the stored table is [c][vf][hf], H rows of W)
      const size_t W = k == 6 ? 16 : (k == 12 ? 64 : 32), H = k == 7 ? 8 : W / 2;
      std::vector<float> wt(w.size()), dt(w.size());
      for (size_t c = 0; c < 3; ++c) for (size_t y = 0; y < H; ++y) for (size_t x = 0; x < W; ++x) {
        wt[c * W * H + x * H + y] = w[c * W * H + y * W + x];
        dt[c * W * H + x * H + y] = dq[c * W * H + y * W + x];
      }
      if (!d_weights_t_[k].Reserve(wt.size()) || !d_dequant_t_[k].Reserve(dt.size())) { *err = "alloc"; return false; }
      CUDA_OK(cudaMemcpy(d_weights_t_[k].p, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(d_dequant_t_[k].p, dt.data(), dt.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  {
    // the tables k_acs_evalsq<32 / 64> reads from global memory, in 16-byte chunks: [c][chunk][lane row][4] — the lanes of
    // a warp then fetch one contiguous run per load instead of one cache line each
    struct Src { int slot, kind; bool transposed; size_t chan, row; };
    const Src srcs[6] = {{0, 8, false, 512, 32}, {1, 8, true, 512, 16}, {2, 5, false, 1024, 32},
                         {3, 12, false, 2048, 64}, {4, 12, true, 2048, 32}, {5, 11, false, 4096, 64}};
    for (const Src& sc : srcs) {
      std::vector<float> w;
      host_quant_weights(sc.kind, &w);
      if (sc.transposed) {
        const size_t W = sc.kind == 12 ? 64 : 32, H = W / 2;
        std::vector<float> wt(w.size());
        for (size_t c = 0; c < 3; ++c) for (size_t y = 0; y < H; ++y) for (size_t x = 0; x < W; ++x) wt[c * W * H + x * H + y] = w[c * W * H + y * W + x];
        w.swap(wt);
      }
      const size_t rows = sc.chan / sc.row;
      std::vector<float> wc(w.size()), dc(w.size());
      for (size_t c = 0; c < 3; ++c) for (size_t r = 0; r < rows; ++r) for (size_t j = 0; j < sc.row; ++j) {
        const float v = w[c * sc.chan + r * sc.row + j];
        const size_t o = c * sc.chan + (j / 4) * rows * 4 + r * 4 + (j % 4);
        wc[o] = v; dc[o] = 1.0f / v;
      }
      if (!d_weights_c_[sc.slot].Reserve(wc.size()) || !d_dequant_c_[sc.slot].Reserve(dc.size())) { *err = "alloc"; return false; }
      CUDA_OK(cudaMemcpy(d_weights_c_[sc.slot].p, wc.data(), wc.size() * 4, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(d_dequant_c_[sc.slot].p, dc.data(), dc.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  {
    if (const char* e = getenv("JXLB200_DCT8_ROWS")) dct8_rows_ = atoi(e);
    if (const char* e = getenv("JXLB200_DCT8_TPS")) dct8_tps_ = atoi(e);
  }
  for (int k = 0; k < 4; ++k) {
    // coefficient position of (lane, j) in k_acs_evalsq<8> (oracle/jxo_acs.cc LanePosition): DCT, DCT4X4, DCT4X8, DCT8X4
    std::vector<float> w;
    host_quant_weights(k == 0 ? 0 : (k == 1 ? 3 : 9), &w);
    std::vector<float> wl(192), dl(192);
    for (int c = 0; c < 3; ++c) for (int l = 0; l < 8; ++l) for (int j = 0; j < 8; ++j) {
      int pos;
      if (k == 0) pos = l * 8 + j;
      else if (k == 1) pos = ((j >> 2) + (l & 3) * 2) * 8 + (l >> 2) + (j & 3) * 2;
      else if (k == 2) pos = ((l >> 2) + (l & 3) * 2) * 8 + j;
      else pos = ((j >> 2) + (j & 3) * 2) * 8 + l;
      wl[c * 64 + l * 8 + j] = w[c * 64 + pos];
      dl[c * 64 + l * 8 + j] = 1.0f / w[c * 64 + pos];
    }
    if (!d_w8_[k].Reserve(192) || !d_dq8_[k].Reserve(192)) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_w8_[k].p, wl.data(), 192 * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_dq8_[k].p, dl.data(), 192 * 4, cudaMemcpyHostToDevice));
  }
  {
    std::vector<uint16_t> order;
    host_natural_order(0, &order);
    uint8_t izz[64];
    for (int k = 0; k < 64; ++k) izz[order[k]] = (uint8_t)k;
    std::vector<float> bias(dct8_v4_bias_entries());
    std::vector<uint8_t> lut(2 * 4 * 256);
    dct8_v4_host_tables(izz, bias.data(), lut.data());
    if (!d_bias8_.Reserve(bias.size()) || !d_lastlut8_.Reserve(lut.size())) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_bias8_.p, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_lastlut8_.p, lut.data(), lut.size(), cudaMemcpyHostToDevice));
  }
  {
    // Q20 log2 table of the clustering cost (same expression as the oracle's Log2Q20)
    std::vector<int> lut(1025);
    for (int i = 0; i <= 1024; ++i) lut[i] = (int)lrint(ldexp(log2(1.0 + (double)i / 1024.0), 20));
    if (!d_log2lut_.Reserve(1025)) { *err = "alloc"; return false; }
    CUDA_OK(cudaMemcpy(d_log2lut_.p, lut.data(), 1025 * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (!(d_hist_.Reserve((size_t)kNumAcContexts * kAcAlphabet) && d_cluster_hist_.Reserve(kMaxClusters * kAcAlphabet) &&
        d_hdr_bits_.Reserve(kMaxClusters * 64) && d_hdr_len_.Reserve(kMaxClusters) &&
        d_cluster_state_.Reserve(cluster_state_bytes()) && d_ctx_map_.Reserve(kNumAcContexts + 64) &&
        d_info_.Reserve((size_t)kMaxClusters * kAcAlphabet * 8) && d_norm_.Reserve(kMaxClusters * kAcAlphabet) &&
        d_rmap_.Reserve((size_t)kMaxClusters * kAnsTabSize) && d_mod_hist_.Reserve(kNumModularCtx * kModAlphabet) &&
        d_lf_words_.Reserve(4096) && d_small_.Reserve(16) && d_tree_words_.Reserve(256) &&
        d_code_len_.Reserve(kNumModularCtx * kModAlphabet) && d_code_bits_.Reserve(kNumModularCtx * kModAlphabet) &&
        d_cm_back_.Reserve(8192) && d_hf_words_.Reserve(8192 + kMaxClusters * 64 + 256) && d_out_info_.Reserve(40))) {
    *err = "alloc"; return false;
  }
  CUDA_OK(cudaMallocHost(&h_out_info_, 40 * sizeof(unsigned long long)));
  {
    // inverse natural coefficient orders of the order classes the search can emit
    static const int rep[13] = {0, 3, 4, 5, 6, 8, 10, 18, 19, -1, -1, -1, -1};
    for (int o = 0; o < 13; ++o) {
      if (rep[o] < 0) continue;
      std::vector<uint16_t> order;
      host_natural_order(rep[o], &order);
      std::vector<uint16_t> inv(order.size());
      for (size_t k = 0; k < order.size(); ++k) inv[order[k]] = (uint16_t)k;
      if (!d_inv_order_[o].Reserve(inv.size())) { *err = "alloc"; return false; }
      CUDA_OK(cudaMemcpy(d_inv_order_[o].p, inv.data(), inv.size() * 2, cudaMemcpyHostToDevice));
      if (o == 4 || o == 5 || o == 6 || o == 8) {   // wide strategy of the class: positions in [hf][vf] lane order
        const size_t W = o == 4 ? 16 : (o == 8 ? 64 : 32), H = o == 5 ? 8 : W / 2;
        std::vector<uint16_t> invt(inv.size());
        for (size_t y = 0; y < H; ++y) for (size_t x = 0; x < W; ++x) invt[x * H + y] = inv[y * W + x];
        const int slot = o == 4 ? 13 : (o == 6 ? 14 : (o == 8 ? 15 : 16));
        if (!d_inv_order_[slot].Reserve(invt.size())) { *err = "alloc"; return false; }
        CUDA_OK(cudaMemcpy(d_inv_order_[slot].p, invt.data(), invt.size() * 2, cudaMemcpyHostToDevice));
      }
    }
  }
  {
    // coefficient-stage tables of the 16 / 32 / 64-sized strategies (list order of k_coeff.cu: tall, wide, square per level,
    // then DCT32X8, DCT8X32): lane order, 16-byte chunks: [c][chunk][lane][4].  Tall and square strategies keep the stored
    // orientation (lane = storage row); wide ones are transposed (lane = storage column).
    struct Src { int kind, order_class; bool wide; size_t lanes, vals; };
    const Src srcs[11] = {{6, 4, false, 8, 16},  {6, 4, true, 16, 8},   {4, 2, false, 16, 16}, {8, 6, false, 16, 32},
                          {8, 6, true, 32, 16},  {5, 3, false, 32, 32}, {12, 8, false, 32, 64}, {12, 8, true, 64, 32},
                          {11, 7, false, 64, 64}, {7, 5, false, 8, 32},  {7, 5, true, 32, 8}};
    static const int rep[13] = {0, 3, 4, 5, 6, 8, 10, 18, 19, -1, -1, -1, -1};
    for (int i = 0; i < 11; ++i) {
      const Src& sc = srcs[i];
      std::vector<float> w;
      host_quant_weights(sc.kind, &w);
      std::vector<uint16_t> order;
      host_natural_order(rep[sc.order_class], &order);
      std::vector<uint16_t> inv(order.size());
      for (size_t k = 0; k < order.size(); ++k) inv[order[k]] = (uint16_t)k;
      const size_t L = sc.lanes, V = sc.vals, n = L * V;
      std::vector<float> wj(3 * n), dj(3 * n);
      std::vector<uint16_t> ij(n);
      for (size_t l = 0; l < L; ++l) for (size_t j = 0; j < V; ++j) {
        // stored position: tall / square [lane][j]; wide: the stored table is [j][lane] (V rows of L)
        const size_t pos = sc.wide ? j * L + l : l * V + j;
        const size_t o = (j / 4) * L * 4 + l * 4 + (j % 4);
        for (size_t c = 0; c < 3; ++c) { wj[c * n + o] = w[c * n + pos]; dj[c * n + o] = 1.0f / w[c * n + pos]; }
        ij[o] = inv[pos];
      }
      if (!d_weights_j_[i].Reserve(wj.size()) || !d_dequant_j_[i].Reserve(dj.size()) || !d_inv_j_[i].Reserve(ij.size())) { *err = "alloc"; return false; }
      CUDA_OK(cudaMemcpy(d_weights_j_[i].p, wj.data(), wj.size() * 4, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(d_dequant_j_[i].p, dj.data(), dj.size() * 4, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(d_inv_j_[i].p, ij.data(), ij.size() * 2, cudaMemcpyHostToDevice));
    }
  }
  if (!d_cvx_.Reserve(27) || !d_cvy_.Reserve(27) || !d_q_.Reserve(1)) { *err = "alloc"; return false; }
  CUDA_OK(cudaMemcpy(d_cvx_.p, kCoveredX, 27, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(d_cvy_.p, kCoveredY, 27, cudaMemcpyHostToDevice));
  // the uploads above come from pageable memory on the legacy stream, the pipeline runs on a non-blocking stream that is
  // not ordered against it: make sure every table has landed before the first kernel can read it
  CUDA_OK(cudaDeviceSynchronize());
  return true;
}

bool Encoder::EnsureFork() {
  if (aux_[0]) return true;
  for (int k = 0; k < 2; ++k) {
    if (cudaStreamCreateWithFlags(&aux_[k], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&ev_join_[k], cudaEventDisableTiming) != cudaSuccess) return false;
  }
  return cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming) == cudaSuccess;
}

void Encoder::Destroy() {
  if (device_ < 0) return;
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  for (int k = 0; k < 2; ++k) {
    if (aux_[k]) { cudaStreamSynchronize(aux_[k]); cudaStreamDestroy(aux_[k]); aux_[k] = nullptr; }
    if (ev_join_[k]) { cudaEventDestroy(ev_join_[k]); ev_join_[k] = nullptr; }
  }
  if (ev_fork_) { cudaEventDestroy(ev_fork_); ev_fork_ = nullptr; }
  if (ev_copy_) { cudaEventDestroy(ev_copy_); ev_copy_ = nullptr; }
  d_lut_.Release();
  for (int k = 0; k < 17; ++k) { d_weights_[k].Release(); d_dequant_[k].Release(); d_weights_t_[k].Release(); d_dequant_t_[k].Release(); }
  for (int k = 0; k < 4; ++k) { d_w8_[k].Release(); d_dq8_[k].Release(); }
  for (int k = 0; k < 6; ++k) { d_weights_c_[k].Release(); d_dequant_c_[k].Release(); }
  for (int k = 0; k < 11; ++k) { d_weights_j_[k].Release(); d_dequant_j_[k].Release(); d_inv_j_[k].Release(); }
  d_xyb_gab_.Release(); d_cfl_.Release();
  d_acs_work_.Release(); d_acs_jobs_.Release(); d_coeff_lists_.Release(); d_recon_xyb_.Release();
  d_bias8_.Release(); d_lastlut8_.Release(); d_cvx_.Release(); d_cvy_.Release();
  for (int o = 0; o < 17; ++o) d_inv_order_[o].Release();
  d_rgb_.Release(); d_xyb_.Release(); d_mask1x1_.Release(); d_pre_.Release(); d_qf_.Release(); d_mask_.Release();
  d_homog_.Release(); d_acs_entropy_.Release(); d_acs_.Release(); d_raw_qf_.Release(); d_cmap_.Release();
  d_coeffs_.Release(); d_dc_quant_.Release(); d_nzeros_.Release(); d_nzcount_.Release(); d_lastk_.Release(); d_q_.Release();
  d_log2lut_.Release(); d_tokens_.Release(); d_token_counts_.Release(); d_hist_.Release(); d_cluster_hist_.Release();
  d_hdr_bits_.Release(); d_hdr_len_.Release(); d_group_arena_.Release(); d_cluster_state_.Release(); d_ctx_map_.Release();
  d_info_.Release(); d_norm_.Release(); d_rmap_.Release(); d_group_start_.Release(); d_dgs_.Release(); d_strat_c_.Release();
  d_qf_c_.Release(); d_first_count_.Release(); d_mod_tokens_.Release(); d_mod_hist_.Release(); d_lf_words_.Release();
  d_small_.Release(); d_tile_sums_.Release(); d_mod_words_.Release(); d_dg_start_.Release(); d_tree_words_.Release();
  d_code_len_.Release(); d_code_bits_.Release(); d_cm_back_.Release(); d_hf_words_.Release(); d_hdr_stage_.Release();
  d_out_.Release(); d_sections_.Release(); d_out_info_.Release();
  if (h_out_info_) cudaFreeHost(h_out_info_);
  h_out_info_ = nullptr;
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
            d_nzeros_.Reserve(3 * nblk) && d_nzcount_.Reserve(3 * nblk) && d_lastk_.Reserve(3 * nblk) &&
            d_acs_work_.Reserve(acs_work_floats(fd)) && d_acs_jobs_.Reserve(acs_work_jobs(fd)) &&
            d_coeff_lists_.Reserve(coeff_list_words(fd));
  // modular element space: one fixed-capacity run per DC group (k_modular.cu)
  h_dgs_.clear();
  uint32_t elem = 0, blocks = 0;
  for (int dg = 0; dg < fd.num_dc_groups; ++dg) {
    DcGroupInfo d;
    d.x0 = (dg % fd.dgxs) * 256; d.y0 = (dg / fd.dgxs) * 256;
    d.w = std::min(256, fd.bxs - d.x0); d.h = std::min(256, fd.bys - d.y0);
    d.tw = (d.w + 7) >> 3; d.th = (d.h + 7) >> 3;
    d.elem_base = elem; d.block_base = blocks;
    elem += 2 + 6 * (uint32_t)(d.w * d.h) + 2 * (uint32_t)(d.tw * d.th);
    blocks += (uint32_t)(d.w * d.h);
    h_dgs_.push_back(d);
  }
  total_elems_ = elem;
  const bool new_geometry = fd.xsize != dgs_w_ || fd.ysize != dgs_h_;
  const size_t nsec = (size_t)2 + fd.num_dc_groups + fd.num_groups;
  const size_t out_words = ((size_t)fd.xsize * fd.ysize * 4 + (1u << 20) + nsec * 8) / 4;
  ok = ok && d_tokens_.Reserve((size_t)fd.num_groups * kTokensPerGroupMax) && d_token_counts_.Reserve(fd.num_groups) &&
       d_group_arena_.Reserve((size_t)fd.num_groups * kTokensPerGroupMax) && d_group_start_.Reserve(fd.num_groups) &&
       d_dgs_.Reserve(fd.num_dc_groups) && d_strat_c_.Reserve(nblk) && d_qf_c_.Reserve(nblk) &&
       d_first_count_.Reserve(fd.num_dc_groups) && d_mod_tokens_.Reserve(total_elems_) &&
       d_tile_sums_.Reserve(total_elems_ / 2048 + 2) && d_mod_words_.Reserve((size_t)total_elems_ + 2) &&
       d_dg_start_.Reserve(fd.num_dc_groups) && d_hdr_stage_.Reserve(nsec + 64) && d_sections_.Reserve(nsec) &&
       d_out_.Reserve(out_words);
  if (!ok) { *err = "device allocation failed"; return false; }
  if (new_geometry) {
    // (a pageable-memory copy inside the pipeline would make the host wait for the stream: do it here, once per geometry)
    if (cudaMemcpy(d_dgs_.p, h_dgs_.data(), h_dgs_.size() * sizeof(DcGroupInfo), cudaMemcpyHostToDevice) != cudaSuccess) {
      *err = "memcpy"; return false;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { *err = "memcpy"; return false; }   // (same ordering argument as in Init)
    dgs_w_ = fd.xsize; dgs_h_ = fd.ysize;
  }
  return true;
}

// true when `p` is page-locked host memory (cudaMallocHost / cudaHostRegister): it can be the source of a
// truly asynchronous copy, so the staging memcpy is skipped
static bool IsPinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

bool Encoder::EnqueueHost(const uint8_t* pixels, int w, int h, size_t stride, const EncodeParams& p, std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  const size_t row = (size_t)3 * w;
  const size_t bytes = row * h;
  if (!d_rgb_.Reserve(bytes + 16)) { *err = "device allocation failed"; return false; }
  CUDA_OK(cudaEventRecord(ev_[0], stream_));
  cudaStream_t cs = copy_stream_ ? copy_stream_ : stream_;
  if (stride == row && IsPinned(pixels)) {
    CUDA_OK(cudaMemcpyAsync(d_rgb_.p, pixels, bytes, cudaMemcpyHostToDevice, cs));
  } else {
    if (bytes > h_pinned_cap_) {
      if (h_pinned_) cudaFreeHost(h_pinned_);
      h_pinned_ = nullptr; h_pinned_cap_ = 0;
      CUDA_OK(cudaMallocHost(&h_pinned_, bytes));
      h_pinned_cap_ = bytes;
    }
    // pack rows into the pinned staging buffer (drops any row padding), then one async copy
    if (stride == row) memcpy(h_pinned_, pixels, bytes);
    else for (int y = 0; y < h; ++y) memcpy(h_pinned_ + (size_t)y * row, pixels + (size_t)y * stride, row);
    CUDA_OK(cudaMemcpyAsync(d_rgb_.p, h_pinned_, bytes, cudaMemcpyHostToDevice, cs));
  }
  if (cs != stream_) {
    if (!ev_copy_) CUDA_OK(cudaEventCreateWithFlags(&ev_copy_, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(ev_copy_, cs));
    CUDA_OK(cudaStreamWaitEvent(stream_, ev_copy_, 0));
  }
  fd_.Set(w, h);
  return Run(d_rgb_.p, row, p, err);
}

bool Encoder::EnqueueDevice(const uint8_t* d_pixels, int w, int h, size_t stride, const EncodeParams& p, std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaEventRecord(ev_[0], stream_));
  fd_.Set(w, h);
  return Run(d_pixels, stride, p, err);
}

bool Encoder::EncodeHost(const uint8_t* pixels, int w, int h, size_t stride, const EncodeParams& p, jxlb200_stats* stats,
                         std::string* err) {
  return EnqueueHost(pixels, w, h, stride, p, err) && Finish(stats, err);
}

bool Encoder::EncodeDevice(const uint8_t* d_pixels, int w, int h, size_t stride, const EncodeParams& p,
                           jxlb200_stats* stats, std::string* err) {
  return EnqueueDevice(d_pixels, w, h, stride, p, err) && Finish(stats, err);
}

// Enqueues every kernel of one encode on the stream; nothing here waits for the device.
bool Encoder::Run(const uint8_t* d_rgb, size_t stride, const EncodeParams& p, std::string* err) {
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
  const bool fork = fork_ && EnsureFork();
  StreamFork sf{{aux_[0], aux_[1]}, ev_fork_, {ev_join_[0], ev_join_[1]}};
  // K4: homogeneity map (the thesis' proposals) — only the proposals read it (H8 / H9; the unpatched encoder has no use for
  // the map); it depends on the XYB planes alone, so a lone frame computes it beside the quant field
  const bool gab = (p.flags & JXLB200_FLAG_GABORISH) != 0;
  const bool homog_aside = fork && !gab && p.proposal != JXLB200_PROPOSAL_NONE;   // (with Gaborish the map reads the sharpened planes)
  if (homog_aside) {
    CUDA_OK(cudaEventRecord(ev_fork_, stream_));
    CUDA_OK(cudaStreamWaitEvent(aux_[0], ev_fork_, 0));
    launch_homogeneity(X, Y, B, fd, p.distance, d_homog_.p, aux_[0]);
    CUDA_OK(cudaEventRecord(ev_join_[0], aux_[0]));
  }
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
  // Gaborish (opt-in): the quant field relies on the pre-sharpening values; the homogeneity map, the search and the
  // coefficients see the sharpened planes (oracle EncodeFrame)
  if (gab) {
    if (!d_xyb_gab_.Reserve(3 * plane)) { *err = "device allocation failed"; return false; }
    launch_gab_inverse(d_xyb_.p, d_xyb_gab_.p, fd, stream_);
    X = d_xyb_gab_.p; Y = X + plane; B = Y + plane;
  }
  xyb_cur_ = X;
  if (homog_aside) CUDA_OK(cudaStreamWaitEvent(stream_, ev_join_[0], 0));
  else if (p.proposal != JXLB200_PROPOSAL_NONE) launch_homogeneity(X, Y, B, fd, p.distance, d_homog_.p, stream_);
  else CUDA_OK(cudaMemsetAsync(d_homog_.p, 0, 3 * nblk * 4, stream_));
  CUDA_OK(cudaEventRecord(ev_[4], stream_));
  // K6: AC strategy search (+ the proposals' hooks); DCT8 everywhere when fixed or below effort 5
  const bool forced = (p.flags & JXLB200_FLAG_FORCED_ACS) != 0;
  if (forced && (forced_bxs_ != fd.bxs || forced_bys_ != fd.bys)) { *err = "JXLB200_FLAG_FORCED_ACS without a strategy map of this frame's size"; return false; }
  // `search` selects the general coefficient path (every strategy); the forced map takes it too
  const bool search = forced || (!(p.flags & JXLB200_FLAG_FIXED_DCT8) && p.effort >= 5);
  CUDA_OK(cudaMemsetAsync(d_cmap_.p, 0, (size_t)2 * fd.txs * fd.tys, stream_));
  // chroma from luma (opt-in): per-tile factors fitted on the planes the search sees; the default map is all zero
  const bool cfl = (p.flags & JXLB200_FLAG_CFL) != 0;
  if (cfl) {
    if (!d_cfl_.Reserve((size_t)fd.txs * fd.tys)) { *err = "device allocation failed"; return false; }
    launch_cfl_fit(X, Y, B, fd, d_weights_[0].p, d_cmap_.p, d_cfl_.p, stream_);
  }
  AcsTables tables;
  for (int k = 0; k < 17; ++k) {
    tables.w[k] = d_weights_[k].p; tables.dq[k] = d_dequant_[k].p; tables.wT[k] = d_weights_t_[k].p; tables.dqT[k] = d_dequant_t_[k].p;
  }
  for (int k = 0; k < 4; ++k) { tables.w8[k] = d_w8_[k].p; tables.dq8[k] = d_dq8_[k].p; }
  for (int k = 0; k < 6; ++k) { tables.wC[k] = d_weights_c_[k].p; tables.dqC[k] = d_dequant_c_[k].p; }
  for (int k = 0; k < 11; ++k) { tables.wJ[k] = d_weights_j_[k].p; tables.dqJ[k] = d_dequant_j_[k].p; tables.invJ[k] = d_inv_j_[k].p; }
  if (forced) {
    CUDA_OK(cudaMemcpyAsync(d_acs_.p, forced_acs_.data(), nblk, cudaMemcpyHostToDevice, stream_));
    CUDA_OK(cudaMemsetAsync(d_acs_entropy_.p, 0, nblk * 4, stream_));
  } else if (search) {
    AcsParams ap;
    const float ratio = (p.distance + 0.1373f) / 1.1373f;
    ap.info_loss_multiplier = 1.2f * powf(ratio, 0.33677806662454718f);
    ap.zeros_mul = 9.3089171683409026f * powf(ratio, 0.50990926717963703f);
    ap.cost_delta = 10.833273317067883f * powf(ratio, 0.36702940662370243f);
    ap.distance = p.distance;
    ap.mul8x8 = 1.0f - 0.4f / (p.distance + 1.4f);
    ap.partitioning = p.proposal == JXLB200_PROPOSAL_PARTITIONING || p.proposal == JXLB200_PROPOSAL_COMBINED;
    ap.factored_entropy = p.proposal == JXLB200_PROPOSAL_FACTORED_ENTROPY || p.proposal == JXLB200_PROPOSAL_COMBINED;
    ap.speed_tier = 10 - (int)p.effort;
    launch_acs(X, Y, B, d_mask1x1_.p, d_qf_.p, d_homog_.p, cfl ? d_cfl_.p : nullptr, fd, ap, tables, d_acs_work_.p, d_acs_jobs_.p, d_acs_.p, d_acs_entropy_.p,
               stream_);
  } else {
    CUDA_OK(cudaMemsetAsync(d_acs_.p, 0x80, nblk, stream_));
    CUDA_OK(cudaMemsetAsync(d_acs_entropy_.p, 0, nblk * 4, stream_));
  }
  launch_raw_qf(d_qf_.p, d_acs_.p, fd, d_q_.p, d_cvx_.p, d_cvy_.p, d_raw_qf_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[5], stream_));
  // K7: transform + quantise
  if (search) {
    const uint16_t* inv_order[17];
    for (int o = 0; o < 17; ++o) inv_order[o] = d_inv_order_[o].p;
    launch_coeff_general(X, Y, B, d_acs_.p, fd, d_q_.p, tables, inv_order, d_cmap_.p, x_qm_mul_, b_qm_mul_,
                         p.effort >= 5 ? 1 : 0, d_raw_qf_.p, d_coeffs_.p, d_dc_quant_.p, d_nzeros_.p, d_nzcount_.p, d_lastk_.p,
                         d_coeff_lists_.p, stream_, fork ? &sf : nullptr);
  } else {
      launch_dct8_quant_v4(X, Y, B, fd, d_q_.p, d_weights_[0].p, d_dequant_[0].p + 64, d_bias8_.p, d_lastlut8_.p, d_cmap_.p,
                           x_qm_mul_, b_qm_mul_, p.effort >= 5 ? 1 : 0, dct8_rows_, dct8_tps_, d_raw_qf_.p, d_coeffs_.p, d_dc_quant_.p,
                           d_nzeros_.p, d_nzcount_.p, d_lastk_.p, stream_);
  }
  CUDA_OK(cudaEventRecord(ev_[6], stream_));
  // K11 (moved up for a lone frame): the modular DC + AC metadata streams and LfGlobal need the coefficient stage's outputs
  // only, so they run on an auxiliary stream beside tokenisation, clustering and the rANS chains
  uint32_t* lf_bits = d_small_.p + 0; uint32_t* mod_total_bits = d_small_.p + 1; uint32_t* hf_bits = d_small_.p + 2;
  uint32_t* tree_bits = d_small_.p + 3;
  auto modular_streams = [&](cudaStream_t ms) -> bool {
    if (tree_ndc_ != fd.num_dc_groups) {
      launch_tree_blob(fd.num_dc_groups, d_tree_words_.p, tree_bits, ms);
      tree_ndc_ = fd.num_dc_groups;
    }
    CUDA_OK(cudaMemsetAsync(d_mod_hist_.p, 0, kNumModularCtx * kModAlphabet * 4, ms));
    CUDA_OK(cudaMemsetAsync(d_mod_words_.p, 0, ((size_t)total_elems_ + 2) * 4, ms));
    CUDA_OK(cudaMemsetAsync(d_out_info_.p + 8, 0, 32 * sizeof(unsigned long long), ms));   // acs histogram slots
    launch_mod_ranks(d_acs_.p, d_raw_qf_.p, fd, d_dgs_.p, fd.num_dc_groups, d_strat_c_.p, d_qf_c_.p, d_first_count_.p,
                     d_out_info_.p + 8, ms);
    launch_mod_tokens(d_dc_quant_.p, d_cmap_.p, d_strat_c_.p, d_qf_c_.p, d_first_count_.p, fd, d_dgs_.p, fd.num_dc_groups,
                      total_elems_, d_mod_tokens_.p, d_mod_hist_.p, ms);
    launch_mod_codes(d_mod_hist_.p, d_q_.p, d_tree_words_.p, tree_bits, d_code_len_.p, d_code_bits_.p, d_lf_words_.p, lf_bits,
                     ms);
    launch_mod_write(d_mod_tokens_.p, d_code_len_.p, d_code_bits_.p, d_dgs_.p, fd.num_dc_groups, d_first_count_.p, total_elems_,
                     d_tile_sums_.p, mod_total_bits, d_mod_words_.p, d_dg_start_.p, ms);
    return true;
  };
  if (fork) {
    CUDA_OK(cudaEventRecord(ev_fork_, stream_));
    CUDA_OK(cudaStreamWaitEvent(aux_[0], ev_fork_, 0));
    if (!modular_streams(aux_[0])) return false;
    CUDA_OK(cudaEventRecord(ev_join_[0], aux_[0]));
  }
  // K8: tokens + per-context histograms
  CUDA_OK(cudaMemsetAsync(d_hist_.p, 0, (size_t)kNumAcContexts * kAcAlphabet * 4, stream_));
  CUDA_OK(cudaMemsetAsync(d_cluster_hist_.p, 0, kMaxClusters * kAcAlphabet * 4, stream_));
  launch_tokenize(d_acs_.p, d_nzeros_.p, d_nzcount_.p, d_lastk_.p, d_coeffs_.p, fd, d_tokens_.p, d_token_counts_.p,
                  d_hist_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[7], stream_));
  // K9: clustering, per-cluster ANS tables
  launch_cluster(d_hist_.p, d_log2lut_.p, d_cluster_state_.p, d_ctx_map_.p, d_cluster_hist_.p, stream_);
  launch_ans_tables(d_cluster_hist_.p, d_cluster_state_.p, d_norm_.p, d_rmap_.p, d_info_.p, d_hdr_bits_.p, d_hdr_len_.p,
                    stream_);
  CUDA_OK(cudaEventRecord(ev_[8], stream_));
  const int* d_num_clusters = reinterpret_cast<const int*>(d_cluster_state_.p + cluster_num_clusters_offset());
  if (fork) {
    // HfGlobal (context map + histogram headers) only needs the tables: beside the rANS chains
    CUDA_OK(cudaEventRecord(ev_fork_, stream_));
    CUDA_OK(cudaStreamWaitEvent(aux_[1], ev_fork_, 0));
    launch_hf_global(d_ctx_map_.p, d_num_clusters, d_hdr_bits_.p, d_hdr_len_.p, fd.num_groups, d_cm_back_.p, d_hf_words_.p,
                     hf_bits, aux_[1]);
    CUDA_OK(cudaEventRecord(ev_join_[1], aux_[1]));
  }
  // K10: one rANS stream per AC group
  launch_ans_groups(d_tokens_.p, d_token_counts_.p, d_ctx_map_.p, d_info_.p, d_rmap_.p, d_num_clusters, d_small_.p + 4,
                    ans_groups_per_warp_, ans_warps_, d_group_arena_.p, d_group_start_.p, fd.num_groups, stream_);
  CUDA_OK(cudaEventRecord(ev_[9], stream_));
  // K11: modular DC + AC metadata streams, LfGlobal
  if (fork) CUDA_OK(cudaStreamWaitEvent(stream_, ev_join_[0], 0));
  else if (!modular_streams(stream_)) return false;
  CUDA_OK(cudaEventRecord(ev_[10], stream_));
  // K12: HfGlobal, headers + TOC, concatenation
  if (fork) CUDA_OK(cudaStreamWaitEvent(stream_, ev_join_[1], 0));
  else launch_hf_global(d_ctx_map_.p, d_num_clusters, d_hdr_bits_.p, d_hdr_len_.p, fd.num_groups, d_cm_back_.p, d_hf_words_.p,
                        hf_bits, stream_);
  launch_finalize(fd, x_qm_scale_, b_qm_scale_, gab ? 1 : 0, lf_bits, d_dg_start_.p, mod_total_bits, hf_bits, d_group_start_.p,
                  d_sections_.p, d_hdr_stage_.p, d_out_.p, (unsigned long long)d_out_.cap * 32, d_out_info_.p, d_q_.p,
                  d_token_counts_.p, d_num_clusters, stream_);
  launch_assemble(d_sections_.p, 2 + fd.num_dc_groups + fd.num_groups, d_lf_words_.p, d_mod_words_.p, d_hf_words_.p,
                  d_group_arena_.p, d_out_.p, d_out_info_.p, stream_);
  CUDA_OK(cudaEventRecord(ev_[11], stream_));
  // K13 (optional): reconstruction error of the coded frame against the input, for stats.sse / stats.psnr
  if (p.flags & JXLB200_FLAG_QUALITY) {
    const uint16_t* inv_order[17];
    for (int o = 0; o < 17; ++o) inv_order[o] = d_inv_order_[o].p;
    if (!d_recon_xyb_.Reserve(3 * plane)) { *err = "device allocation failed"; return false; }
    if (!search) launch_coeff_lists(d_acs_.p, fd, d_coeff_lists_.p, stream_);   // (the search path binned the map already)
    launch_recon_sse(fd, d_q_.p, tables, inv_order, d_cmap_.p, 1.0f / powf(1.25f, (float)(x_qm_scale_ - 2)),
                     1.0f / powf(1.25f, (float)(b_qm_scale_ - 2)), d_acs_.p, d_raw_qf_.p, d_coeffs_.p, d_dc_quant_.p, d_rgb, stride,
                     d_recon_tab_.p, d_coeff_lists_.p, d_recon_xyb_.p, d_out_info_.p + 36, gab ? 1 : 0, stream_);
  }
  CUDA_OK(cudaEventRecord(ev_[12], stream_));
  CUDA_OK(cudaMemcpyAsync(h_out_info_, d_out_info_.p, 40 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
  launches_ = g_kernel_launches;
  in_flight_ = true;
  return true;
}

// Waits for the encode enqueued last and collects its statistics.
bool Encoder::Finish(jxlb200_stats* stats, std::string* err) {
  if (!in_flight_) { *err = "no encode in flight"; return false; }
  in_flight_ = false;
  const FrameDim& fd = fd_;
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  CUDA_OK(cudaGetLastError());
  if (h_out_info_[1]) { *err = "codestream larger than the output arena"; return false; }
  codestream_bytes_ = (size_t)h_out_info_[0];
  have_frame_ = true;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->width = fd.xsize; stats->height = fd.ysize;
    stats->num_groups = fd.num_groups; stats->num_dc_groups = fd.num_dc_groups;
    stats->kernel_launches = launches_;
    stats->global_scale = (uint32_t)h_out_info_[2]; stats->quant_dc = (uint32_t)h_out_info_[3];
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_[0], ev_[1]); stats->stage_ms[JXLB200_T_H2D] = ms;
    cudaEventElapsedTime(&ms, ev_[1], ev_[2]); stats->stage_ms[JXLB200_T_XYB] = ms;
    cudaEventElapsedTime(&ms, ev_[2], ev_[3]); stats->stage_ms[JXLB200_T_AQ] = ms;
    cudaEventElapsedTime(&ms, ev_[3], ev_[4]); stats->stage_ms[JXLB200_T_HOMOG] = ms;
    cudaEventElapsedTime(&ms, ev_[4], ev_[5]); stats->stage_ms[JXLB200_T_ACS] = ms;
    cudaEventElapsedTime(&ms, ev_[5], ev_[6]); stats->stage_ms[JXLB200_T_COEFF] = ms;
    cudaEventElapsedTime(&ms, ev_[6], ev_[7]); stats->stage_ms[JXLB200_T_TOKENIZE] = ms;
    cudaEventElapsedTime(&ms, ev_[7], ev_[8]); stats->stage_ms[JXLB200_T_HISTO] = ms;
    cudaEventElapsedTime(&ms, ev_[8], ev_[9]); stats->stage_ms[JXLB200_T_ANS] = ms;
    cudaEventElapsedTime(&ms, ev_[9], ev_[10]); stats->stage_ms[JXLB200_T_DC] = ms;
    cudaEventElapsedTime(&ms, ev_[10], ev_[11]); stats->stage_ms[JXLB200_T_ASSEMBLE] = ms;
    cudaEventElapsedTime(&ms, ev_[11], ev_[12]); stats->stage_ms[JXLB200_T_QUALITY] = ms;
    cudaEventElapsedTime(&ms, ev_[0], ev_[12]); stats->total_ms = ms;
    if (params_.flags & JXLB200_FLAG_QUALITY) {
      stats->quality_valid = 1;
      double sum = 0.0;
      for (int k = 0; k < 3; ++k) { stats->sse[k] = h_out_info_[36 + k]; sum += (double)stats->sse[k]; }
      const double mse = sum / (3.0 * (double)fd.xsize * (double)fd.ysize);
      stats->psnr = mse > 0.0 ? 10.0 * log10(255.0 * 255.0 / mse) : INFINITY;
    }
    stats->codestream_bytes = codestream_bytes_;
    stats->bpp = 8.0 * (double)codestream_bytes_ / ((double)fd.xsize * fd.ysize);
    for (int i = 0; i < 27; ++i) stats->acs_histogram[i] = (uint32_t)h_out_info_[8 + i];
    stats->num_tokens = h_out_info_[4];
    stats->num_clusters = (uint32_t)h_out_info_[5];
  }
  return true;
}

bool Encoder::Fetch(uint8_t** out, size_t* out_len, std::string* err) {
  if (!have_frame_) { *err = "no encoded frame"; return false; }
  cudaSetDevice(device_);
  uint8_t* buf = (uint8_t*)malloc(codestream_bytes_ ? codestream_bytes_ : 1);
  if (!buf) { *err = "out of host memory"; return false; }
  if (cudaMemcpyAsync(buf, d_out_.p, codestream_bytes_, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
      cudaStreamSynchronize(stream_) != cudaSuccess) { free(buf); *err = "memcpy"; return false; }
  *out = buf;
  *out_len = codestream_bytes_;
  return true;
}

bool Encoder::SetForcedAcs(const uint8_t* acs, int bxs, int bys, std::string* err) {
  // must be a partition into transforms this path codes, none leaving its 64x64 tile (libjxl's AcStrategyImage invariants)
  std::vector<uint8_t> seen((size_t)bxs * bys, 0);
  for (int by = 0; by < bys; ++by) for (int bx = 0; bx < bxs; ++bx) {
    const uint8_t a = acs[(size_t)by * bxs + bx];
    if (!(a & 0x80)) continue;
    const int s = a & 0x7f;
    const bool ok = s <= 13 || (s >= 18 && s <= 20);
    if (!ok) { *err = "strategy map: unsupported strategy"; return false; }
    const int cx = kCoveredX[s], cy = kCoveredY[s];
    if (bx + cx > bxs || by + cy > bys || (bx & 7) + cx > 8 || (by & 7) + cy > 8) { *err = "strategy map: transform leaves the frame or its tile"; return false; }
    for (int iy = 0; iy < cy; ++iy) for (int ix = 0; ix < cx; ++ix) {
      const size_t i = (size_t)(by + iy) * bxs + bx + ix;
      if ((acs[i] & 0x7f) != s || ((acs[i] & 0x80) != 0) != (ix == 0 && iy == 0) || seen[i]) { *err = "strategy map: inconsistent transform"; return false; }
      seen[i] = 1;
    }
  }
  for (uint8_t v : seen) if (!v) { *err = "strategy map: uncovered block"; return false; }
  forced_acs_.assign(acs, acs + (size_t)bxs * bys);
  forced_bxs_ = bxs; forced_bys_ = bys;
  return true;
}

bool Encoder::DebugHomogeneity(const float* x, const float* y, const float* b, int stride, int ysize, float distance, float* out,
                               std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  // geometry exactly as given: the row pitch IS the diff's src_stride bound, ysize its src_ysize (H1)
  FrameDim fd{};
  fd.xsize = fd.xs_pad = stride; fd.ysize = fd.ys_pad = ysize; fd.pitch = stride; fd.bxs = stride / 8; fd.bys = ysize / 8;
  const size_t plane = (size_t)stride * ysize, nblk = (size_t)fd.bxs * fd.bys;
  DevBuf<float> planes, res;
  if (!planes.Reserve(3 * plane) || !res.Reserve(3 * nblk + 1)) { *err = "device allocation failed"; return false; }
  const float* src[3] = {x, y, b};
  for (int c = 0; c < 3; ++c) CUDA_OK(cudaMemcpyAsync(planes.p + c * plane, src[c], plane * 4, cudaMemcpyHostToDevice, stream_));
  launch_homogeneity(planes.p, planes.p + plane, planes.p + 2 * plane, fd, distance, res.p, stream_);
  CUDA_OK(cudaMemcpyAsync(out, res.p, 3 * nblk * 4, cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  planes.Release(); res.Release();
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
    case JXLB200_STAGE_XYB: src = xyb_cur_ ? xyb_cur_ : d_xyb_.p; bytes = 3 * plane * 4; break;
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
    case JXLB200_STAGE_HISTOGRAMS: src = d_hist_.p; bytes = (size_t)kNumAcContexts * kAcAlphabet * 4; break;
    case JXLB200_STAGE_CONTEXT_MAP: src = d_ctx_map_.p; bytes = kNumAcContexts; break;
    case JXLB200_STAGE_CODESTREAM: src = d_out_.p; bytes = codestream_bytes_; break;
    case 21: src = d_cluster_state_.p + cluster_num_clusters_offset(); bytes = 4; break;
    case JXLB200_STAGE_TOKEN_OFFSETS:
    case JXLB200_STAGE_TOKENS: {
      std::vector<uint32_t> counts(fd.num_groups), offs(fd.num_groups + 1, 0);
      if (cudaMemcpy(counts.data(), d_token_counts_.p, counts.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
      for (int g = 0; g < fd.num_groups; ++g) offs[g + 1] = offs[g] + counts[g];
      if (stage == JXLB200_STAGE_TOKEN_OFFSETS) {
        bytes = offs.size() * 4;
        if (dst && cap >= bytes) memcpy(dst, offs.data(), bytes);
        return (int64_t)bytes;
      }
      bytes = (size_t)offs.back() * 4;
      if (dst && cap >= bytes) {
        for (int g = 0; g < fd.num_groups; ++g)
          if (counts[g] && cudaMemcpy((uint8_t*)dst + (size_t)offs[g] * 4, d_tokens_.p + (size_t)g * kTokensPerGroupMax,
                                      (size_t)counts[g] * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
      }
      return (int64_t)bytes;
    }
    case JXLB200_STAGE_GROUP_OFFSETS:
    case JXLB200_STAGE_GROUP_STREAMS: {
      // every AC group's rANS stream sits at the END of its arena slot (it is written back to front): bits
      // [start_bit, slot end).  The tap returns each stream from its first bit, LSB-first, zero-padded to a whole byte
      // (what the oracle's per-group BitWriter holds), and the byte offsets of the groups.
      std::vector<unsigned long long> start(fd.num_groups);
      if (cudaMemcpy(start.data(), d_group_start_.p, start.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
      const unsigned long long slot_bits = (unsigned long long)kTokensPerGroupMax * 32;
      std::vector<uint32_t> offs(fd.num_groups + 1, 0);
      for (int g = 0; g < fd.num_groups; ++g) offs[g + 1] = offs[g] + (uint32_t)((slot_bits - start[g] + 7) / 8);
      if (stage == JXLB200_STAGE_GROUP_OFFSETS) {
        bytes = offs.size() * 4;
        if (dst && cap >= bytes) memcpy(dst, offs.data(), bytes);
        return (int64_t)bytes;
      }
      bytes = offs.back();
      if (dst && cap >= bytes) {
        std::vector<uint32_t> words;
        for (int g = 0; g < fd.num_groups; ++g) {
          const size_t w0 = (size_t)(start[g] / 32), nw = (size_t)kTokensPerGroupMax - w0;
          words.assign(nw + 1, 0u);
          if (nw && cudaMemcpy(words.data(), d_group_arena_.p + (size_t)g * kTokensPerGroupMax + w0, nw * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "memcpy"; return -1; }
          const unsigned sh = (unsigned)(start[g] % 32);
          const unsigned long long nbits = slot_bits - start[g];
          uint8_t* o = (uint8_t*)dst + offs[g];
          for (unsigned long long b = 0; b < nbits; b += 8) {
            const unsigned long long pos = sh + b;                   // bit position inside `words`
            const size_t wi = (size_t)(pos / 32);
            const unsigned bo = (unsigned)(pos % 32);
            unsigned long long two = (unsigned long long)words[wi] | ((unsigned long long)words[wi + 1] << 32);
            unsigned v = (unsigned)((two >> bo) & 0xFFu);
            const unsigned long long left = nbits - b;
            if (left < 8) v &= (1u << left) - 1u;
            o[b / 8] = (uint8_t)v;
          }
        }
      }
      return (int64_t)bytes;
    }
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
