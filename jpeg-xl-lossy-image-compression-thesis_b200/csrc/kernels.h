// jxlb200 — kernel launchers (one .cu per pipeline stage)
#pragma once
#include "jxl_common.cuh"

namespace jxlb {
// number of kernels launched by this thread (reported as jxlb200_stats.kernel_launches)
extern thread_local unsigned g_kernel_launches;
// K1 (k_xyb.cu)
void launch_rgb8_to_xyb(const uint8_t* d_rgb, size_t stride, int w, int h, const FrameDim& fd, const float* d_lut,
                        float* x, float* y, float* b, cudaStream_t s);
// K2 (k_aq.cu)
void launch_aq(const float* x, const float* y, const float* b, const FrameDim& fd, float distance, float* mask1x1,
               float* pre, float* qf, float* mask, cudaStream_t s);
void launch_fill(float* p, size_t n, float v, cudaStream_t s);
void launch_quant_params(const float* qf, size_t n, float quant_dc, QuantDev* q, cudaStream_t s);
void launch_raw_qf(const float* qf, const uint8_t* acs, const FrameDim& fd, const QuantDev* q, const uint8_t* cvx,
                   const uint8_t* cvy, int32_t* raw, cudaStream_t s);
// K4 (k_homog.cu)
void launch_homogeneity(const float* x, const float* y, const float* b, const FrameDim& fd, float distance,
                        float* out, cudaStream_t s);
// K7, DCT8 frames, two threads per block with cp.async-staged tiles (k_dct8_v4.cu)
int dct8_v4_bias_entries();
void dct8_v4_host_tables(const uint8_t* izz64, float* bias, uint8_t* last_lut /* [2][4][256] */);
void launch_dct8_quant_v4(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd,
                          const float* weights, const float* dequant_y, const float* bias_tab, const uint8_t* last_lut,
                          const int8_t* cmap, float x_qm_mul, float b_qm_mul, int adjust, int rows_per_cta, int threads_per_sm,
                          int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant, uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk,
                          cudaStream_t s);
// K6 (k_acs.cu) / K7 general (k_coeff.cu)
struct AcsParams {
  float info_loss_multiplier, zeros_mul, cost_delta, distance, mul8x8;
  int factored_entropy, partitioning;   // H9 / H8 hooks of the thesis' proposals
  int speed_tier;                       // 10 - effort (libjxl SpeedTier: hare = 5 ... tortoise = 1)
};
// quantisation weights / their inverses per table kind; wT / dqT: the same tables transposed ([hf][vf] of the WIDE
// strategy of kinds 6, 8, 12: the order a lane that owns one horizontal frequency reads them in)
// w8 / dq8: the 8x8 tables of DCT, DCT4X4, DCT4X8, DCT8X4 in the lane order of k_acs_evalsq<8> ([lane][j])
// wC / dqC: the tables of k_acs_evalsq<32> (slots 0..2 = tall halves, wide halves, square) and <64> (3..5) in 16-byte
// chunks, [c][chunk][lane row][4]
struct AcsTables {
  const float* w[17]; const float* dq[17]; const float* wT[17]; const float* dqT[17]; const float* w8[4]; const float* dq8[4];
  const float* wC[6]; const float* dqC[6];
  // the coefficient stage's 16 / 32 / 64-sized strategies (k_coeff.cu list order from kList16Tall on): quantisation
  // tables and inverse scan orders in lane order and 16-byte chunks, [c][chunk][lane][4]
  const float* wJ[11]; const float* dqJ[11]; const uint16_t* invJ[11];
};
size_t acs_work_floats(const FrameDim& fd);   // candidate-value tables of the search
size_t acs_work_jobs(const FrameDim& fd);     // counters + lists of the non-aligned squares (uint32 words)
// chroma from luma (opt-in JXLB200_FLAG_CFL): per-tile ytox / ytob from the planes the search sees; w8 = DCT8 quantisation weights
void launch_cfl_fit(const float* x, const float* y, const float* b, const FrameDim& fd, const float* w8, int8_t* cmap, float2* factors,
                    cudaStream_t s);
void launch_acs(const float* x, const float* y, const float* b, const float* mask1x1, const float* qf, const float* homog,
                const float2* cfl /* per-tile factors of launch_cfl_fit, or nullptr for the default map */, const FrameDim& fd, const AcsParams& P, const AcsTables& T, float* work, uint32_t* jobs, uint8_t* acs, float* est,
                cudaStream_t s);
// two auxiliary streams + the events that order them against the main stream (nullptr: everything on the main stream)
struct StreamFork { cudaStream_t aux[2]; cudaEvent_t fork, join[2]; };
void launch_coeff_general(const float* x, const float* y, const float* b, const uint8_t* acs, const FrameDim& fd,
                          const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order, const int8_t* cmap,
                          float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant,
                          uint8_t* nzeros, uint16_t* nzcount, uint16_t* lastk, uint32_t* lists, cudaStream_t s,
                          const StreamFork* fork = nullptr);
void launch_coeff_lists(const uint8_t* acs, const FrameDim& fd, uint32_t* lists, cudaStream_t s);   // bins the first blocks by strategy
size_t coeff_list_words(const FrameDim& fd);  // per-class transform lists of k_coeff / k_recon (uint32 words)

// Gaborish, encoder side (k_gab.cu; opt-in JXLB200_FLAG_GABORISH): 5x5 sharpening of the padded XYB planes, src -> dst
void launch_gab_inverse(const float* src, float* dst, const FrameDim& fd, cudaStream_t s);

// K13 (k_recon.cu): per-channel sum of squared errors of the coded frame's reconstruction against the input pixels
void launch_recon_sse(const FrameDim& fd, const QuantDev* qd, const AcsTables& T, const uint16_t* const* inv_order,
                      const int8_t* cmap, float inv_qm_x, float inv_qm_b, const uint8_t* acs, const int32_t* raw_qf,
                      const int16_t* coeffs, const int16_t* dc_quant, const uint8_t* rgb, size_t stride, const float* tables,
                      const uint32_t* lists, float* scratch_xyb, unsigned long long* sse3, int gab, cudaStream_t s);

// ---- entropy stage -------------------------------------------------------------------------
// one 2048x2048 DC group: rectangle in blocks, its 64x64-tile rectangle, first element / first block slot
struct DcGroupInfo { int x0, y0, w, h, tw, th; uint32_t elem_base; uint32_t block_base; };
// one TOC section: where its bits live (kind 0 LfGlobal, 1 modular stream words, 2 HfGlobal, 3 AC group arena)
struct Section { int kind; int index; unsigned long long src_bit, nbits, dst_bit; };

// K8-K10 (k_entropy.cu)
size_t cluster_state_bytes();
int cluster_num_clusters_offset();
void launch_tokenize(const uint8_t* acs, const uint8_t* nzeros, const uint16_t* nzcount, const uint16_t* lastk,
                     const int16_t* coeffs, const FrameDim& fd, uint32_t* tokens, uint32_t* token_counts, uint32_t* hist,
                     cudaStream_t s);
void launch_cluster(const uint32_t* hist, const int* lut, void* state, uint8_t* cmap, uint32_t* cluster_hist,
                    cudaStream_t s);
void launch_ans_tables(const uint32_t* cluster_hist, const void* state, uint16_t* norm, uint16_t* rmap, void* info,
                       uint32_t* hdr_bits, uint32_t* hdr_len, cudaStream_t s);
void launch_ans_groups(const uint32_t* tokens, const uint32_t* token_counts, const uint8_t* cmap, const void* info,
                       const uint16_t* rmap, const int* num_clusters, uint32_t* work_counter, int groups_per_warp,
                       int warps_per_cta, uint32_t* out_arena, unsigned long long* start_bit, int num_groups, cudaStream_t s);
// K11 (k_modular.cu)
void launch_tree_blob(int num_dc_groups, uint32_t* tree_words, uint32_t* tree_bits, cudaStream_t s);
void launch_mod_ranks(const uint8_t* acs, const int32_t* raw_qf, const FrameDim& fd, const DcGroupInfo* dgs, int num_dg,
                      int32_t* strat_c, int32_t* qf_c, uint32_t* first_count, unsigned long long* acs_hist, cudaStream_t s);
void launch_mod_tokens(const int16_t* dc_quant, const int8_t* cmap, const int32_t* strat_c, const int32_t* qf_c,
                       const uint32_t* first_count, const FrameDim& fd, const DcGroupInfo* dgs, int num_dg,
                       uint32_t total_elems, uint32_t* tokens, uint32_t* mod_hist, cudaStream_t s);
void launch_mod_codes(const uint32_t* mod_hist, const QuantDev* qd, const uint32_t* tree_words, const uint32_t* tree_bits,
                      uint8_t* code_len, uint16_t* code_bits, uint32_t* lf_words, uint32_t* lf_bits, cudaStream_t s);
void launch_mod_write(const uint32_t* tokens, const uint8_t* code_len, const uint16_t* code_bits, const DcGroupInfo* dgs,
                      int num_dg, const uint32_t* first_count, uint32_t total_elems, uint32_t* tile_sums, uint32_t* total_bits,
                      uint32_t* words, uint32_t* dg_start_bit, cudaStream_t s);
// K12 (k_assemble.cu)
void launch_hf_global(const uint8_t* cmap, const int* num_clusters, const uint32_t* hdr_bits, const uint32_t* hdr_len,
                      int num_groups, uint32_t* cm_back, uint32_t* hf_words, uint32_t* hf_bits, cudaStream_t s);
void launch_finalize(const FrameDim& fd, int x_qm_scale, int b_qm_scale, int gab, const uint32_t* lf_bits, const uint32_t* dg_start_bit,
                     const uint32_t* mod_total_bits, const uint32_t* hf_bits, const unsigned long long* group_start_bit,
                     Section* sections, uint32_t* hdr_stage, uint32_t* out_words, unsigned long long out_capacity_bits,
                     unsigned long long* out_info, const QuantDev* qd, const uint32_t* token_counts, const int* num_clusters,
                     cudaStream_t s);
void launch_assemble(const Section* sections, int num_sections, const uint32_t* lf_words, const uint32_t* mod_words,
                     const uint32_t* hf_words, const uint32_t* group_arena, uint32_t* out_words,
                     const unsigned long long* out_info, cudaStream_t s);
}  // namespace jxlb
