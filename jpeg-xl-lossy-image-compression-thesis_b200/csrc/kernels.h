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
// K7 (k_dct_quant.cu)
void launch_dct8_quant(const float* x, const float* y, const float* b, const FrameDim& fd, const QuantDev* qd,
                       const float* weights, const float* dequant_y, const uint8_t* izz, const int8_t* cmap,
                       float x_qm_mul, float b_qm_mul, int adjust, int32_t* raw_qf, int16_t* coeffs, int16_t* dc_quant,
                       uint8_t* nzeros, uint8_t* lastpos, cudaStream_t s);
}  // namespace jxlb
