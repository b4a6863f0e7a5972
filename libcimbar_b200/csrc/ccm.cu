// ccm.cu -- color_correction == 1: one von Kries adaptation matrix per frame from the anchors' white (see ccm.cuh)
#include "ccm.cuh"

namespace cb200 {

__device__ __forceinline__ void matx33_mul_rn(const float* a, const float* b, float* out)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) out[3 * i + j] = dot3_rn(a[3 * i], a[3 * i + 1], a[3 * i + 2], b[j], b[3 + j], b[6 + j]);
}

// cv::Matx<float,3,3>::inv(): closed-form 3x3 (opencv2/core/operations.hpp): determinant in float, d = 1 / d, cofactors * d
__device__ __forceinline__ bool matx33_inv_rn(const float* a, float* b)
{
    auto mm = [](float x, float y, float z, float w) { return __fsub_rn(__fmul_rn(x, y), __fmul_rn(z, w)); };   // x*y - z*w
    float d = __fadd_rn(__fsub_rn(__fmul_rn(a[0], mm(a[4], a[8], a[7], a[5])), __fmul_rn(a[1], mm(a[3], a[8], a[6], a[5]))),
                        __fmul_rn(a[2], mm(a[3], a[7], a[6], a[4])));
    if (d == 0.0f) return false;
    d = __fdiv_rn(1.0f, d);
    b[0] = __fmul_rn(mm(a[4], a[8], a[5], a[7]), d); b[1] = __fmul_rn(mm(a[2], a[7], a[1], a[8]), d); b[2] = __fmul_rn(mm(a[1], a[5], a[2], a[4]), d);
    b[3] = __fmul_rn(mm(a[5], a[6], a[3], a[8]), d); b[4] = __fmul_rn(mm(a[0], a[8], a[2], a[6]), d); b[5] = __fmul_rn(mm(a[2], a[3], a[0], a[5]), d);
    b[6] = __fmul_rn(mm(a[3], a[7], a[4], a[6]), d); b[7] = __fmul_rn(mm(a[1], a[6], a[0], a[7]), d); b[8] = __fmul_rn(mm(a[0], a[4], a[1], a[3]), d);
    return true;
}

// one thread per frame: calculateWhite (dark layout: anchors top-left, bottom-left, top-right; 4x4 means; floor 1) and
// get_adaptation_matrix<von_kries>(white, (255,255,255)) = T.inv() * diag((T*desired) / (T*white)) * T
__global__ void k_ccm_simple(const Mode m, const uint8_t* __restrict__ rgb, int n_frames, float* __restrict__ ccm)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const int W = m.width, H = m.height;
    const uint8_t* frame = rgb + (size_t)f * W * H * 3;
    const int tl = 30 - 2, right = W - 30 - 2, bottom = H - 30 - 2;       // Config::anchor_size() == 30, padding 0
    const int ax[3] = {tl, tl, right}, ay[3] = {tl, bottom, tl};
    float white[3] = {1.0f, 1.0f, 1.0f};
    for (int a = 0; a < 3; ++a) {
        uint32_t sum[3] = {0, 0, 0};
        for (int y = 0; y < 4; ++y)
            for (int x = 0; x < 4; ++x) {
                const uint8_t* p = frame + ((size_t)(ay[a] + y) * W + (ax[a] + x)) * 3;
                sum[0] += p[0]; sum[1] += p[1]; sum[2] += p[2];
            }
        for (int c = 0; c < 3; ++c) { const float v = (float)((double)sum[c] / 16.0); if (v > white[c]) white[c] = v; }   // cv::mean is double
    }
    const float T[9] = {0.4002400f, 0.7076000f, -0.0808100f, -0.2263000f, 1.1653200f, 0.0457000f, 0.0000000f, 0.0000000f, 0.9182200f};
    float d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ti[9], tmp[9], out[9];
    for (int i = 0; i < 3; ++i) {
        const float m1 = dot3_rn(T[3 * i], T[3 * i + 1], T[3 * i + 2], white[0], white[1], white[2]);
        const float m2 = dot3_rn(T[3 * i], T[3 * i + 1], T[3 * i + 2], 255.0f, 255.0f, 255.0f);
        d[4 * i] = __fdiv_rn(m2, m1);
    }
    if (!matx33_inv_rn(T, ti)) for (int i = 0; i < 9; ++i) ti[i] = 0.0f;
    matx33_mul_rn(ti, d, tmp);
    matx33_mul_rn(tmp, T, out);
    for (int i = 0; i < 9; ++i) ccm[(size_t)f * 9 + i] = out[i];
}

cudaError_t ccm_simple_launch(const Mode& m, const uint8_t* d_rgb, int n_frames, float* d_ccm, cudaStream_t st)
{
    k_ccm_simple<<<(n_frames + 63) / 64, 64, 0, st>>>(m, d_rgb, n_frames, d_ccm);
    return cudaGetLastError();
}

}  // namespace cb200
