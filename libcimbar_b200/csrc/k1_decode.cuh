// k1_decode.cuh -- host-side entry points of the fused K1 kernel (k1_decode.cu)
#pragma once
#include "cb200_common.cuh"
#include "ccm.cuh"

namespace cb200 {
size_t k1_smem_bytes(bool sharpen);
// resident CTAs per SM the grid is sized for: `plain` (the context's tuning value) without sharpen, three with it
int k1_ctas_per_sm(bool sharpen, int plain);
cudaError_t k1_init_tables(const float* adjust256, const unsigned long long* tiles_L16);
cudaError_t k1_launch(const Mode& m, const uint8_t* d_rgb, int n_frames, int bands, int grid, int l2_ahead, bool sharpen,
                      uint8_t* d_cellvals, uint32_t* d_dirty, const CcmArg& cc, cudaStream_t stream);
cudaError_t k1_symbols_launch(const uint16_t* d_windows, const uint8_t* d_cooldown, int n, uint8_t* d_sym, uint8_t* d_off, uint8_t* d_dist, cudaStream_t st);
cudaError_t k1_colors_launch(const Mode& m, const uint8_t* d_rgb, int n, uint8_t* d_color, const CcmArg& cc, cudaStream_t st);
}  // namespace cb200
