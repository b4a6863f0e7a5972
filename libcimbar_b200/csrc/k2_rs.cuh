// k2_rs.cuh -- host-side entry points of the pack / Reed-Solomon / chunk-mask kernels (k2_rs.cu)
#pragma once
#include "cb200_common.cuh"

namespace cb200 {
cudaError_t k2_init_tables(const uint8_t* exp512, const uint8_t* log256);
cudaError_t k2_pack_launch(const Mode& m, const uint8_t* d_cellvals, const uint16_t* d_idx, int n_frames, uint8_t* d_raw, cudaStream_t st);
// d_rho: 4 x 64 bytes from k2_remainder_basis (the four x^(D+j) mod x^pad g vectors the remainder tables are built from)
cudaError_t k2_rs_launch(const Mode& m, const uint8_t* d_raw, int n_frames, uint8_t* d_data, uint8_t* d_ok, const uint8_t* d_rho, int sm_count, cudaStream_t st);
void k2_remainder_basis(const uint8_t* gen, int parity, const uint8_t* gexp512, const uint8_t* glog256, uint8_t* rho_out);
// same, with the block bytes gathered from K1's per-cell bytes through the interleave map (no raw stream materialised);
// b_begin / b_count restrict the launch to a block range of every frame (b_count < 0: all blocks)
cudaError_t k2_rs_fused_launch(const Mode& m, const uint8_t* d_cellvals, const uint16_t* d_idx, int n_frames, uint8_t* d_data,
                               uint8_t* d_ok, const uint8_t* d_rho, int sm_count, cudaStream_t st, int b_begin = 0, int b_count = -1);
cudaError_t k2_mask_launch(const Mode& m, const uint8_t* d_ok, int n_frames, uint32_t* d_mask, cudaStream_t st);
}  // namespace cb200
