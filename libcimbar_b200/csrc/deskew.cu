// deskew.cu -- the extractor's deskew in front of the decode path, sm_100a.
//
// Replaces (reference file:line relative to /root/reference/):
//   Deskewer::deskew        src/lib/extractor/Deskewer.h:25-40: cv::getPerspectiveTransform(corners, outputPoints) +
//                           cv::warpPerspective(img, output, transform, size, cv::INTER_LINEAR) to the mode's image size
//   (Extractor::extract     src/lib/extractor/Extractor.h:30-46 calls it with the four anchor centres Scanner found: scan.cu, or
//                           corners the caller supplies)
// OpenCV is a third-party dependency of the reference; its arithmetic is restated here and pinned against cv2 itself
// (tests/test_deskew.py): getPerspectiveTransform = an 8x8 LU solve in double (hal LUImpl: partial pivoting, d = -1/pivot,
// row updates, back substitution), cv::invert(3x3) in closed form, warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0) =
//   per destination pixel, in double:  W = 32 / (M6 x + M7 y + M8);  X = cvRound((M0 x + M1 y + M2) W), Y likewise,
//   evaluated block-wise exactly as OpenCV does (X0 = M0 bx + M1 y + M2 for the 64-pixel block origin bx, then + M0 x1);
//   source pixel (X >> 5, Y >> 5) with 5-bit fractions ax, ay and the fixed-point bilinear weights
//   (32-ax)(32-ay), ax(32-ay), (32-ax)ay, ax ay (x 32 = OpenCV's 15-bit table, which is exact for these fractions),
//   result (sum + 512) >> 10, taps outside the source count as 0.
// The output frames land in device memory in the layout cb200_decode_*_dev take: a camera frame goes H2D once and never
// returns to the host before its chunks do.
#include "ctx.cuh"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace cb200 {

struct DeskewState {
    double* d_minv = nullptr;      // n x 9 inverse maps (destination -> source)
    int cap = 0;
    uint8_t* d_src = nullptr;      // staging for the host-pointer entry point
    size_t src_bytes = 0;
    uint8_t* d_dst = nullptr;      // deskewed frames of the host-pointer entry points
    size_t dst_bytes = 0;
};

void deskew_destroy(DeskewState* d)
{
    if (!d) return;
    cudaFree(d->d_minv); cudaFree(d->d_src); cudaFree(d->d_dst);
    delete d;
}

// hal::LU64f (modules/core/src/matrix_decomp.cpp, LUImpl<double>) for an m x m system with one right-hand side
static bool lu_solve(double* A, int m, double* b)
{
    const double eps = DBL_EPSILON * 100;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j) if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
        if (std::fabs(A[k * m + i]) < eps) return false;
        if (k != i) {
            for (int j = i; j < m; ++j) { const double t = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; ++j) {
            const double alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; ++c) A[j * m + c] += alpha * A[i * m + c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < m; ++c) s -= A[i * m + c] * b[c];
        b[i] = s / A[i * m + i];
    }
    return true;
}

// cv::invert for a 3x3 double matrix (modules/core/src/lapack.cpp: det3 + adjugate)
static bool invert3(const double* S, double* t)
{
#define SD(r, c) S[(r) * 3 + (c)]
    double d = SD(0, 0) * (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) - SD(0, 1) * (SD(1, 0) * SD(2, 2) - SD(1, 2) * SD(2, 0)) +
               SD(0, 2) * (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0));
    if (d == 0.) return false;
    d = 1. / d;
    t[0] = (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) * d;
    t[1] = (SD(0, 2) * SD(2, 1) - SD(0, 1) * SD(2, 2)) * d;
    t[2] = (SD(0, 1) * SD(1, 2) - SD(0, 2) * SD(1, 1)) * d;
    t[3] = (SD(1, 2) * SD(2, 0) - SD(1, 0) * SD(2, 2)) * d;
    t[4] = (SD(0, 0) * SD(2, 2) - SD(0, 2) * SD(2, 0)) * d;
    t[5] = (SD(0, 2) * SD(1, 0) - SD(0, 0) * SD(1, 2)) * d;
    t[6] = (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0)) * d;
    t[7] = (SD(0, 1) * SD(2, 0) - SD(0, 0) * SD(2, 1)) * d;
    t[8] = (SD(0, 0) * SD(1, 1) - SD(0, 1) * SD(1, 0)) * d;
#undef SD
    return true;
}

// one thread = four consecutive destination pixels (12 bytes = three aligned words); blockIdx.y = frame (grid-strided), so the
// index arithmetic is 32-bit.  No branch sits between the address arithmetic and the loads, so the loads of a thread's four pixels
// are all in flight together, and the two horizontal taps of a source row -- six consecutive bytes -- are fetched as aligned 32-bit
// words and realigned by funnel shifts: a source that lies rotated in the photograph makes every lane hit its own sector, and the
// kernel is then bound by L1 sector look-ups (ncu: 27 sectors per request), so fewer, wider requests are what counts.
// A tap outside the source gets weight 0 (== BORDER_CONSTANT 0); its bytes come from a clamped, valid address.
__device__ __forceinline__ void load_row_pair(const uint8_t* __restrict__ p, const uint8_t* __restrict__ end, uint32_t& lo, uint32_t& hi)
{   // bytes p[0..5] -> lo = p[0..3], hi = p[4..5] (upper half unspecified); `end` = one past the last readable byte
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t a = (uint32_t)(addr & 3u);
    const uint8_t* base = p - a;
    if (base + 12 <= end) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(base);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
        const uint32_t w2 = (a == 3u) ? __ldg(w + 2) : 0u;           // six bytes from offset 3 reach into the third word
        lo = __funnelshift_r(w0, w1, 8u * a);
        hi = __funnelshift_r(w1, w2, 8u * a);
    } else {                                                        // the last few bytes of the buffer: byte loads
        lo = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
        hi = (uint32_t)__ldg(p + 4) | ((uint32_t)__ldg(p + 5) << 8);
    }
}

__global__ void __launch_bounds__(256)
k_deskew(const uint8_t* __restrict__ src, int src_w, int src_h, size_t src_frame_bytes, const double* __restrict__ minv, int n,
         int dst_w, int dst_h, int bw0, uint8_t* __restrict__ dst)
{
    const int quads = dst_w >> 2;
    const int per_frame = dst_h * quads;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= per_frame) return;
    const int y = i / quads, q = i - y * quads;
    const int x = 4 * q, bx = (x / bw0) * bw0;
    const uint8_t* const end = src + (size_t)n * src_frame_bytes;
    for (int f = (int)blockIdx.y; f < n; f += (int)gridDim.y) {
        const double* M = minv + (size_t)f * 9;
        const uint8_t* s = src + (size_t)f * src_frame_bytes;
        const double M0 = __ldg(M), M1 = __ldg(M + 1), M2 = __ldg(M + 2), M3 = __ldg(M + 3), M4 = __ldg(M + 4), M5 = __ldg(M + 5),
                     M6 = __ldg(M + 6), M7 = __ldg(M + 7), M8 = __ldg(M + 8);
        // no FMA contraction: OpenCV's x86-64 code rounds after every multiply and add
        const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M0, (double)bx), __dmul_rn(M1, (double)y)), M2);
        const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M3, (double)bx), __dmul_rn(M4, (double)y)), M5);
        const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M6, (double)bx), __dmul_rn(M7, (double)y)), M8);
        uint32_t wgt[4][4];
        uint32_t off[4][2];                      // byte offsets of the two row pairs inside the source picture (< 4 GB)
        bool left[4], right[4];                  // the pair was clamped: tap x1 is the pair's first pixel / tap x0 its second
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double x1 = (double)(x + u - bx);
            double W = __dadd_rn(W0, __dmul_rn(M6, x1));
            W = W != 0. ? __ddiv_rn(32.0, W) : 0.;
            double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M0, x1)), W), fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M3, x1)), W);
            fX = fmax(-2147483648.0, fmin(2147483647.0, fX)); fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
            const int X = __double2int_rn(fX), Y = __double2int_rn(fY);        // cvRound: to nearest, ties to even
            int sx = X >> 5, sy = Y >> 5;
            sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx); sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);   // saturate_cast<short>
            const uint32_t ax = (uint32_t)(X & 31), ay = (uint32_t)(Y & 31);
            const bool x0in = sx >= 0 && sx < src_w, x1in = sx + 1 >= 0 && sx + 1 < src_w;
            const bool y0in = sy >= 0 && sy < src_h, y1in = sy + 1 >= 0 && sy + 1 < src_h;
            wgt[u][0] = (x0in && y0in) ? (32u - ax) * (32u - ay) : 0u;
            wgt[u][1] = (x1in && y0in) ? ax * (32u - ay) : 0u;
            wgt[u][2] = (x0in && y1in) ? (32u - ax) * ay : 0u;
            wgt[u][3] = (x1in && y1in) ? ax * ay : 0u;
            // the pair of source pixels (c, c + 1) that holds whichever of the taps sx, sx + 1 are inside the row
            const int c = sx < 0 ? 0 : (sx > src_w - 2 ? src_w - 2 : sx);
            left[u] = sx < c; right[u] = sx > c;
            const int cy0 = sy < 0 ? 0 : (sy >= src_h ? src_h - 1 : sy), cy1 = sy + 1 < 0 ? 0 : (sy + 1 >= src_h ? src_h - 1 : sy + 1);
            off[u][0] = 3u * ((uint32_t)cy0 * (uint32_t)src_w + (uint32_t)c);
            off[u][1] = 3u * ((uint32_t)cy1 * (uint32_t)src_w + (uint32_t)c);
        }
        uint32_t lo[4][2], hi[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) { load_row_pair(s + off[u][0], end, lo[u][0], hi[u][0]); load_row_pair(s + off[u][1], end, lo[u][1], hi[u][1]); }
        uint32_t px[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                uint32_t acc = 512u;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint32_t pa = (lo[u][r] >> (8 * ch)) & 0xFFu;                                   // pixel c
                    const uint32_t pb = ch == 0 ? lo[u][r] >> 24 : (hi[u][r] >> (8 * (ch - 1))) & 0xFFu;  // pixel c + 1
                    const uint32_t v0 = right[u] ? pb : pa, v1 = left[u] ? pa : pb;                      // taps sx, sx + 1
                    acc += wgt[u][2 * r] * v0 + wgt[u][2 * r + 1] * v1;
                }
                px[u][ch] = acc >> 10;
            }
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(dst + ((size_t)f * dst_h + y) * (size_t)dst_w * 3 + (size_t)x * 3);
        out[0] = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
        out[1] = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
        out[2] = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
    }
}

static DeskewState* dstate(cb200_ctx* c)
{
    if (!c->deskew) c->deskew = new DeskewState();
    return c->deskew;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_perspective_transform(const float* src_xy, const float* dst_xy, double* m9_out)
{
    if (!src_xy || !dst_xy || !m9_out) return fail(CB200_ERR_ARG, "null argument");
    // cv::getPerspectiveTransform (modules/imgproc/src/imgwarp.cpp): c00*xi + c01*yi + c02 - ui*(c20*xi + c21*yi) = ui, ...
    double a[8][8], b[8];
    for (int i = 0; i < 4; ++i) {
        // the points are cv::Point2f: the four products are single-precision products, as in OpenCV's source
        const float sx = src_xy[2 * i], sy = src_xy[2 * i + 1], dx = dst_xy[2 * i], dy = dst_xy[2 * i + 1];
        a[i][0] = a[i + 4][3] = sx;
        a[i][1] = a[i + 4][4] = sy;
        a[i][2] = a[i + 4][5] = 1;
        a[i][3] = a[i][4] = a[i][5] = a[i + 4][0] = a[i + 4][1] = a[i + 4][2] = 0;
        a[i][6] = -sx * dx;
        a[i][7] = -sy * dx;
        a[i + 4][6] = -sx * dy;
        a[i + 4][7] = -sy * dy;
        b[i] = dx;
        b[i + 4] = dy;
    }
    if (!lu_solve(&a[0][0], 8, b)) {      // cv::solve returns false and leaves zeros: a degenerate quadrilateral
        for (int i = 0; i < 8; ++i) m9_out[i] = 0;
        m9_out[8] = 1;
        return fail(CB200_ERR_ARG, "degenerate corner quadrilateral");
    }
    for (int i = 0; i < 8; ++i) m9_out[i] = b[i];
    m9_out[8] = 1.;
    return CB200_OK;
}

int cb200_deskew_dev(cb200_ctx* c, const uint8_t* d_src, int src_w, int src_h, int n, const double* m9, uint8_t* d_dst)
{
    if (!c || !d_src || !m9 || !d_dst || n < 0 || src_w < 2 || src_h < 2) return fail(CB200_ERR_ARG, "bad arguments");
    if ((size_t)src_w * (size_t)src_h * 3 >= ((size_t)1 << 32)) return fail(CB200_ERR_ARG, "source picture of 4 GB or more");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    const Mode& m = c->mode;
    DeskewState* d = dstate(c);
    if (n > d->cap) {
        cudaFree(d->d_minv); d->d_minv = nullptr; d->cap = 0;
        CK(cudaMalloc(&d->d_minv, sizeof(double) * 9 * (size_t)n), "cudaMalloc transforms");
        d->cap = n;
    }
    // warpPerspective inverts the transform it is given (no WARP_INVERSE_MAP): cv::invert, DECOMP_LU; a singular matrix maps
    // everything to (0, 0) there (invert leaves zeros) -- reported here instead
    static thread_local std::vector<double> inv;
    inv.resize((size_t)n * 9);
    for (int f = 0; f < n; ++f)
        if (!invert3(m9 + (size_t)f * 9, inv.data() + (size_t)f * 9)) return fail(CB200_ERR_ARG, "singular perspective transform");
    CK(cudaMemcpyAsync(d->d_minv, inv.data(), sizeof(double) * 9 * (size_t)n, cudaMemcpyHostToDevice, c->stream), "H2D transforms");
    CK(cudaStreamSynchronize(c->stream), "sync (transforms staged from a host vector)");
    // OpenCV's block geometry (WarpPerspectiveInvoker): bh0 = min(16, H); bw0 = min(1024 / bh0, W)
    const int bh0 = m.height < 16 ? m.height : 16;
    int bw0 = 1024 / bh0; if (bw0 > m.width) bw0 = m.width;
    const int per_frame = m.height * (m.width / 4);
    const dim3 grid((unsigned)((per_frame + 255) / 256), (unsigned)(n < 32768 ? n : 32768));
    k_deskew<<<grid, 256, 0, c->stream>>>(d_src, src_w, src_h, (size_t)src_w * src_h * 3, d->d_minv, n, m.width, m.height, bw0, d_dst);
    count_launch();
    CK(cudaGetLastError(), "deskew launch");
    return CB200_OK;
}

int cb200_deskew(cb200_ctx* c, const uint8_t* src, int src_w, int src_h, int n, const double* m9, uint8_t* dst)
{
    if (!c || !src || !dst || n < 0 || n > c->max_frames) return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    DeskewState* d = dstate(c);
    const size_t sb = (size_t)n * src_w * src_h * 3, db = (size_t)n * c->mode.width * c->mode.height * 3;
    if (sb > d->src_bytes) { cudaFree(d->d_src); d->d_src = nullptr; d->src_bytes = 0; CK(cudaMalloc(&d->d_src, sb), "cudaMalloc deskew source"); d->src_bytes = sb; }
    if (db > d->dst_bytes) { cudaFree(d->d_dst); d->d_dst = nullptr; d->dst_bytes = 0; CK(cudaMalloc(&d->d_dst, db), "cudaMalloc deskew output"); d->dst_bytes = db; }
    CK(cudaMemcpyAsync(d->d_src, src, sb, cudaMemcpyHostToDevice, c->stream), "H2D camera frames");
    int rc = cb200_deskew_dev(c, d->d_src, src_w, src_h, n, m9, d->d_dst); if (rc) return rc;
    CK(cudaMemcpyAsync(dst, d->d_dst, db, cudaMemcpyDeviceToHost, c->stream), "D2H frames");
    CK(cudaStreamSynchronize(c->stream), "sync");
    return CB200_OK;
}

// the pictures are already in device memory (the scan entry points stage them once for scan + deskew)
int cb200_extract_decode_fountain_dev(cb200_ctx* c, const uint8_t* d_src, int src_w, int src_h, int n, const float* corners, uint32_t flags,
                                      uint8_t* chunks_out, uint32_t* chunk_count, uint32_t* chunk_mask, uint8_t* frame_flags)
{
    if (!c || !d_src || !corners || !chunks_out || !chunk_count || n < 0 || n > c->max_frames) return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    const Mode& m = c->mode;
    DeskewState* d = dstate(c);
    // Deskewer::deskew (Deskewer.h:27-39) with padding 0: the anchor centres go to (anchor, anchor) ... (W - anchor, H - anchor)
    const float an = 30.0f;                                     // Config::anchor_size() (Config.h)
    const float outp[8] = {an, an, (float)m.width - an, an, an, (float)m.height - an, (float)m.width - an, (float)m.height - an};
    std::vector<double> m9((size_t)n * 9);
    for (int f = 0; f < n; ++f) {
        int rc = cb200_perspective_transform(corners + (size_t)f * 8, outp, m9.data() + (size_t)f * 9); if (rc) return rc;
    }
    const size_t db = (size_t)n * m.width * m.height * 3;
    if (db > d->dst_bytes) { cudaFree(d->d_dst); d->d_dst = nullptr; d->dst_bytes = 0; CK(cudaMalloc(&d->d_dst, db), "cudaMalloc deskew output"); d->dst_bytes = db; }
    int rc = cb200_deskew_dev(c, d_src, src_w, src_h, n, m9.data(), d->d_dst); if (rc) return rc;
    // the deskewed frames never leave the device: straight into the decode
    return cb200_decode_fountain_from_dev(c, d->d_dst, n, flags, chunks_out, chunk_count, chunk_mask, frame_flags);
}

int cb200_extract_decode_fountain(cb200_ctx* c, const uint8_t* src, int src_w, int src_h, int n, const float* corners, uint32_t flags,
                                  uint8_t* chunks_out, uint32_t* chunk_count, uint32_t* chunk_mask, uint8_t* frame_flags)
{
    if (!c || !src || !corners || !chunks_out || !chunk_count || n < 0 || n > c->max_frames) return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    DeskewState* d = dstate(c);
    const size_t sb = (size_t)n * src_w * src_h * 3;
    if (sb > d->src_bytes) { cudaFree(d->d_src); d->d_src = nullptr; d->src_bytes = 0; CK(cudaMalloc(&d->d_src, sb), "cudaMalloc deskew source"); d->src_bytes = sb; }
    CK(cudaMemcpyAsync(d->d_src, src, sb, cudaMemcpyHostToDevice, c->stream), "H2D camera frames");
    return cb200_extract_decode_fountain_dev(c, d->d_src, src_w, src_h, n, corners, flags, chunks_out, chunk_count, chunk_mask, frame_flags);
}

}  // extern "C"
