// render.cuh -- synthetic frame renderer (render.cu)
#pragma once
#include "cb200_common.cuh"

namespace cb200 {
cudaError_t render_init_tables(const unsigned long long* tiles_H16);
cudaError_t render_launch(const Mode& m, const uint8_t* d_cellvals, int n_frames, uint8_t* d_rgb, cudaStream_t st);
}  // namespace cb200
