// encode.cuh -- device-side input generator (encode.cu)
#pragma once
#include "cb200_common.cuh"

namespace cb200 {
cudaError_t encode_init_tables(const uint8_t* exp512, const uint8_t* log256);
cudaError_t encode_launch(const Mode& m, const uint8_t* d_gen, const uint16_t* d_inv, const uint8_t* d_payload, int n_frames,
                          uint8_t* d_raw, uint8_t* d_cellvals, cudaStream_t st);
}  // namespace cb200
