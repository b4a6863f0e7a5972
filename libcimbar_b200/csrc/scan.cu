// scan.cu -- the extractor's anchor scan on the device (the step in front of the deskew), sm_100a.
//
// Replaces, for a batch of camera pictures resident in HBM (reference file:line relative to /root/reference/src/lib/extractor/):
//   Scanner::Scanner / preprocess_image(fast)   Scanner.h:146-174   cvtColor(RGB2GRAY) + GaussianBlur(unit x unit, sigma 0) + Otsu
//   Scanner::scan                               Scanner.cpp:182-199 scan_primary (t1 rows -> t2 column -> t3 diagonal -> t4 confirm,
//                                                                   filter_candidates, sort_top_to_bottom) + add_bottom_right_corner
//   Extractor::extract                          Extractor.h:30-46   scan -> Corners -> Deskewer (deskew.cu) -> NEEDS_SHARPEN test
// Three kernels:
//   k_scan_blur4<R>  (k_scan_blur<R>: the one-pixel-per-element first version, CB200_SCAN_BLUR=0)
//                    gray + separable fixed-point Gaussian (OpenCV's 8-bit path: coefficients / 256 from its small-kernel
//                    table, BORDER_REFLECT_101, (sum + 2^15) >> 16) in 128 x 32 tiles staged in shared memory, plus the
//                    picture's 256-bin histogram (shared-memory atomics, one global atomic per bin and tile)
//   k_scan_otsu      getThreshVal_Otsu_8u in double precision, one thread per picture (no FMA contraction)
//   k_scan_anchors   one CTA per picture runs scan_core.cuh's scan_picture: the rows of a t1 pass and the confirmation chains
//                    of its hits are spread over the threads, the order-dependent tail (on_t1_scan's shadow test, libstdc++'s
//                    std::sort, the bottom-right window) runs on thread 0
// The pixel work (gray/blur/histogram) moves 3 bytes in and 1 out per pixel (measured: 0.36 of the HBM copy peak, bound by
// instruction issue); the scan itself touches ~60 rows and a few hundred short lines of the blurred picture.
#include "ctx.cuh"
#include "scan_core.cuh"

#include <cstdlib>
#include <vector>

namespace cb200 {

using namespace scan;

struct ScanScratch {                 // per-context scratch of the scan entry points
    uint8_t* d_pics = nullptr; size_t pics_bytes = 0;       // staging of host pictures
    uint8_t* d_blur = nullptr; size_t blur_bytes = 0;       // blurred gray, n x h x w
    unsigned* d_hist = nullptr; int* d_thr = nullptr; int thr_cap = 0;
    Anchor* d_rowbuf = nullptr; int* d_rowcnt = nullptr; Anchor* d_pts = nullptr; Anchor* d_res = nullptr; int* d_nres = nullptr;
    int ws_pics = 0, ws_rows_cap = 0;
    int4* d_anchors = nullptr; int* d_count = nullptr; unsigned* d_cutoff = nullptr; int* d_status = nullptr;
    int4* h_anchors = nullptr; int* h_count = nullptr;      // pinned results: [n][4] anchors, then count / cutoff / status per picture
};

void scan_destroy(ScanScratch* s)
{
    if (!s) return;
    cudaFree(s->d_pics); cudaFree(s->d_blur); cudaFree(s->d_hist); cudaFree(s->d_thr);
    cudaFree(s->d_rowbuf); cudaFree(s->d_rowcnt); cudaFree(s->d_pts); cudaFree(s->d_res); cudaFree(s->d_nres);
    cudaFree(s->d_anchors); cudaFree(s->d_count); cudaFree(s->d_cutoff); cudaFree(s->d_status);
    cudaFreeHost(s->h_anchors); cudaFreeHost(s->h_count);
    delete s;
}

// ---------------------------------------------------------------------------------------------- gray + blur + histogram
template <int R> struct BlurK;
template <> struct BlurK<1> { __device__ static constexpr unsigned c(int i) { return i == 1 ? 128u : 64u; } };
template <> struct BlurK<2> { __device__ static constexpr unsigned c(int i) { return i == 2 ? 96u : ((i == 1 || i == 3) ? 64u : 16u); } };
template <> struct BlurK<3> { __device__ static constexpr unsigned c(int i) { return i == 3 ? 72u : ((i == 2 || i == 4) ? 56u : ((i == 1 || i == 5) ? 28u : 8u)); } };
template <> struct BlurK<4> {
    __device__ static constexpr unsigned c(int i) { return i == 4 ? 60u : ((i == 3 || i == 5) ? 51u : ((i == 2 || i == 6) ? 30u : ((i == 1 || i == 7) ? 13u : 4u))); }
};

__device__ __forceinline__ int reflect101_clamped(int p, int n)
{   // BORDER_REFLECT_101 for the positions a filter tap can reach (|overshoot| < n); far-outside tile padding is clamped (unused)
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

constexpr int kBlurTW = 128, kBlurTH = 32, kBlurThreads = 256;

template <int R>
__global__ void __launch_bounds__(kBlurThreads)
k_scan_blur(const uint8_t* __restrict__ rgb, int w, int h, uint8_t* __restrict__ out, unsigned* __restrict__ hist)
{
    constexpr int GW = kBlurTW + 2 * R, GH = kBlurTH + 2 * R, KS = 2 * R + 1;
    __shared__ uint8_t g[GH][GW];
    __shared__ uint16_t hs[GH][kBlurTW];
    __shared__ unsigned lh[256];
    const int tid = threadIdx.x, pic = blockIdx.z, tx0 = blockIdx.x * kBlurTW, ty0 = blockIdx.y * kBlurTH;
    const size_t npx = (size_t)w * (size_t)h;
    const uint8_t* src = rgb + (size_t)pic * npx * 3;
    lh[tid] = 0;
    for (int i = tid; i < GH * GW; i += kBlurThreads) {
        const int r = i / GW, c = i - r * GW;
        const int y = reflect101_clamped(ty0 - R + r, h), x = reflect101_clamped(tx0 - R + c, w);
        const uint8_t* p = src + ((size_t)y * w + x) * 3;
        g[r][c] = (uint8_t)((9798u * p[0] + 19235u * p[1] + 3735u * p[2] + 16384u) >> 15);      // cvtColor(RGB2GRAY), 8 bit
    }
    __syncthreads();
    for (int i = tid; i < GH * kBlurTW; i += kBlurThreads) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        unsigned s = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k) s += BlurK<R>::c(k) * g[r][c + k];
        hs[r][c] = (uint16_t)s;                                   // <= 255 * 256
    }
    __syncthreads();
    uint8_t* dst = out + (size_t)pic * npx;
    for (int i = tid; i < kBlurTH * kBlurTW; i += kBlurThreads) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        const int y = ty0 + r, x = tx0 + c;
        if (y < h && x < w) {
            unsigned s = 0;
#pragma unroll
            for (int k = 0; k < KS; ++k) s += BlurK<R>::c(k) * hs[r + k][c];
            const unsigned v = (s + 32768u) >> 16;
            dst[(size_t)y * w + x] = (uint8_t)v;
            atomicAdd(&lh[v], 1u);
        }
    }
    __syncthreads();
    if (lh[tid]) atomicAdd(&hist[(size_t)pic * 256 + tid], lh[tid]);
}

// ---------------------------------------------------------------------------------------------- the same, four pixels per thread
// k_scan_blur is bound by instruction issue (ncu: 85 % of the issue slots, 14 % of the DRAM bandwidth): a division per element for the
// tile indexing, three byte loads and three multiplies per pixel of gray, byte-wise filter taps.  This version gives every thread four
// consecutive pixels: the twelve RGB bytes come as three aligned words and are converted with K1's IDP.2A form (8 instructions per four
// pixels), the horizontal taps are IDP.4A dot products of realigned gray words with the packed coefficients, the vertical pass reads
// four 16-bit sums per 64-bit load, and a warp owns a tile row (no divisions).  Same arithmetic, same results.
// The word path needs 4-byte aligned rows: a picture width that is a multiple of four and an aligned base (checked by the caller);
// otherwise, and in tiles that cross the right edge, pixels are fetched one by one.
template <int R> struct BlurKW {     // the 2R+1 coefficients as bytes of up to three words (IDP.4A operands)
    __device__ static constexpr uint32_t w(int i)
    {
        return (4 * i + 0 <= 2 * R ? BlurK<R>::c(4 * i + 0) : 0u) | ((4 * i + 1 <= 2 * R ? BlurK<R>::c(4 * i + 1) : 0u) << 8) |
               ((4 * i + 2 <= 2 * R ? BlurK<R>::c(4 * i + 2) : 0u) << 16) | ((4 * i + 3 <= 2 * R ? BlurK<R>::c(4 * i + 3) : 0u) << 24);
    }
};

__device__ __forceinline__ uint32_t gray_scalar(const uint8_t* p)
{
    return (9798u * p[0] + 19235u * p[1] + 3735u * p[2] + 16384u) >> 15;
}

template <int R>
__global__ void __launch_bounds__(kBlurThreads)
k_scan_blur4(const uint8_t* __restrict__ rgb, int w, int h, int words_ok, uint8_t* __restrict__ out, unsigned* __restrict__ hist)
{
    constexpr int GH = kBlurTH + 2 * R, GP = kBlurTW + 8, NW = (2 * R + 1 + 3) / 4;      // gray pitch: interior at byte 4
    __shared__ __align__(16) uint8_t g[GH][GP];
    __shared__ __align__(16) uint16_t hs[GH][kBlurTW];
    __shared__ unsigned lh[256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pic = blockIdx.z, tx0 = blockIdx.x * kBlurTW, ty0 = blockIdx.y * kBlurTH;
    const size_t npx = (size_t)w * (size_t)h;
    const uint8_t* src = rgb + (size_t)pic * npx * 3;
    lh[tid] = 0;
    // ---- gray of the tile and its halo
    const int x = tx0 + 4 * lane;
    for (int r = warp; r < GH; r += kBlurThreads / 32) {
        const int y = reflect101_clamped(ty0 - R + r, h);
        const uint8_t* row = src + (size_t)y * (size_t)w * 3;
        uint32_t g4;
        if (words_ok && x + 3 < w) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(row + 3 * (size_t)x);
            const uint32_t q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);          // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
            const uint32_t cRG = 19596u | (38470u << 16), cB0 = 7470u, c0R = 19596u << 16, cGB = 38470u | (7470u << 16);
            uint32_t n0, n1, n2, n3;                     // (19596 R + 38470 G + 7470 B + 2^15): gray is byte 2 (== (9798 R + 19235 G + 3735 B + 2^14) >> 15)
            n0 = __dp2a_lo(cRG, q0, 32768u); n0 = __dp2a_hi(cB0, q0, n0);
            n1 = __dp2a_hi(c0R, q0, 32768u); n1 = __dp2a_lo(cGB, q1, n1);
            n2 = __dp2a_hi(cRG, q1, 32768u); n2 = __dp2a_lo(cB0, q2, n2);
            n3 = __dp2a_lo(c0R, q2, 32768u); n3 = __dp2a_hi(cGB, q2, n3);
            g4 = __byte_perm(__byte_perm(n0, n1, 0x0062), __byte_perm(n2, n3, 0x6200), 0x7610);
        } else {
            g4 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) g4 |= gray_scalar(row + 3 * (size_t)reflect101_clamped(x + k, w)) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(&g[r][4 + 4 * lane]) = g4;
        if (lane < 2 * R) {                               // halo columns: R to the left of the tile, R to the right
            const int xx = lane < R ? tx0 - R + lane : tx0 + kBlurTW + (lane - R);
            g[r][lane < R ? 4 - R + lane : 4 + kBlurTW + (lane - R)] = (uint8_t)gray_scalar(row + 3 * (size_t)reflect101_clamped(xx, w));
        }
    }
    __syncthreads();
    // ---- horizontal pass: output j of the thread is pixel 4 lane + j = gray byte 8 + 4 lane + j - 4; its taps start at byte 4 + j - R of the
    // 16-byte window (wA, wB, wC, 0) that begins at byte 4 lane of the row
    for (int r = warp; r < GH; r += kBlurThreads / 32) {
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g[r][4 * lane]);
        const uint32_t W4[5] = {gw[0], gw[1], gw[2], 0u, 0u};
        uint32_t sums[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = 4 + j - R, wi = t >> 2, sh = 8 * (t & 3);
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const uint32_t d = sh ? __funnelshift_r(W4[wi + k], W4[wi + k + 1 < 5 ? wi + k + 1 : 4], sh) : W4[wi + k];
                acc = __dp4a(d, BlurKW<R>::w(k), acc);
            }
            sums[j] = acc;                                // <= 255 * 256
        }
        *reinterpret_cast<uint2*>(&hs[r][4 * lane]) = make_uint2(sums[0] | (sums[1] << 16), sums[2] | (sums[3] << 16));
    }
    __syncthreads();
    // ---- vertical pass, histogram, store
    uint8_t* dst = out + (size_t)pic * npx;
    for (int r = warp; r < kBlurTH; r += kBlurThreads / 32) {
        const int y = ty0 + r;
        if (y >= h) break;
        uint32_t a0 = 32768u, a1 = 32768u, a2 = 32768u, a3 = 32768u;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) {
            const uint2 v = *reinterpret_cast<const uint2*>(&hs[r + k][4 * lane]);
            const uint32_t c = BlurK<R>::c(k);
            a0 += c * (v.x & 0xFFFFu); a1 += c * (v.x >> 16); a2 += c * (v.y & 0xFFFFu); a3 += c * (v.y >> 16);
        }
        const uint32_t v0 = a0 >> 16, v1 = a1 >> 16, v2 = a2 >> 16, v3 = a3 >> 16;
        uint8_t* o = dst + (size_t)y * w + x;
        if (x + 3 < w) {
            if (words_ok) *reinterpret_cast<uint32_t*>(o) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
            else { o[0] = (uint8_t)v0; o[1] = (uint8_t)v1; o[2] = (uint8_t)v2; o[3] = (uint8_t)v3; }
            atomicAdd(&lh[v0], 1u); atomicAdd(&lh[v1], 1u); atomicAdd(&lh[v2], 1u); atomicAdd(&lh[v3], 1u);
        } else {
            if (x < w) { o[0] = (uint8_t)v0; atomicAdd(&lh[v0], 1u); }
            if (x + 1 < w) { o[1] = (uint8_t)v1; atomicAdd(&lh[v1], 1u); }
            if (x + 2 < w) { o[2] = (uint8_t)v2; atomicAdd(&lh[v2], 1u); }
        }
    }
    __syncthreads();
    if (lh[tid]) atomicAdd(&hist[(size_t)pic * 256 + tid], lh[tid]);
}

// ---------------------------------------------------------------------------------------------- Otsu
// cv::threshold(THRESH_OTSU) -> getThreshVal_Otsu_8u (modules/imgproc/src/thresh.cpp), double precision, operation for operation
__global__ void k_scan_otsu(const unsigned* __restrict__ hist, int n, double npx, int* __restrict__ thr)
{
    const int pic = blockIdx.x * blockDim.x + threadIdx.x;
    if (pic >= n) return;
    const unsigned* hh = hist + (size_t)pic * 256;
    const double scale = ddiv(1., npx);
    double mu = 0;
    for (int i = 0; i < 256; ++i) mu = dadd(mu, dmul((double)i, (double)hh[i]));
    mu = dmul(mu, scale);
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    const double eps = 1.1920928955078125e-07;        // FLT_EPSILON
    for (int i = 0; i < 256; ++i) {
        const double p_i = dmul((double)hh[i], scale);
        mu1 = dmul(mu1, q1);
        q1 = dadd(q1, p_i);
        const double q2 = dsub(1., q1);
        const double lo = q1 < q2 ? q1 : q2, hi = q1 > q2 ? q1 : q2;
        if (lo < eps || hi > dsub(1., eps)) continue;
        mu1 = ddiv(dadd(mu1, dmul((double)i, p_i)), q1);
        const double mu2 = ddiv(dsub(mu, dmul(q1, mu1)), q2);
        const double d = dsub(mu1, mu2);
        const double sigma = dmul(dmul(dmul(q1, q2), d), d);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    thr[pic] = max_val;
}

// ---------------------------------------------------------------------------------------------- the scan
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads)
k_scan_anchors(const uint8_t* __restrict__ blurred, const int* __restrict__ thr, int w, int h, int rows_cap,
               Anchor* rowbuf, int* rowcnt, Anchor* pts, Anchor* res, int* nres,
               int4* __restrict__ anchors_out, int* __restrict__ count_out, unsigned* __restrict__ cutoff_out, int* __restrict__ status_out)
{
    __shared__ PicShared sh;
    const int pic = blockIdx.x;
    Img im;
    im.px = blurred + (size_t)pic * (size_t)w * (size_t)h; im.w = w; im.h = h; im.thr = thr[pic];
    PicWs ws;
    ws.rowbuf = rowbuf + (size_t)pic * rows_cap * kRowCap; ws.rowcnt = rowcnt + (size_t)pic * rows_cap;
    ws.pts = pts + (size_t)pic * kPtsCap; ws.res = res + (size_t)pic * kPtsCap * kResCap; ws.nres = nres + (size_t)pic * kPtsCap;
    ws.rows_cap = rows_cap;
    Exec ex;
    ex.tid = threadIdx.x; ex.nthreads = blockDim.x;
    Anchor* out = reinterpret_cast<Anchor*>(anchors_out + (size_t)pic * 4);
    if (threadIdx.x < 4) out[threadIdx.x] = mk(0, 0, 0, 0);
    __syncthreads();
    const int count = scan_picture(ex, im, ws, sh, out, cutoff_out + pic, status_out + pic);
    if (threadIdx.x == 0) count_out[pic] = count;
}

static ScanScratch* sstate(cb200_ctx* c)
{
    if (!c->scan) c->scan = new ScanScratch();
    return c->scan;
}

static int scan_blur_radius(int w, int h)               // Scanner.h:93-103, :155-157
{
    unsigned v = (unsigned)((w < h ? w : h) * 0.002);
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    unsigned unit = v + 2;
    if (unit < 3u) unit = 3u;
    return (int)(unit / 2);
}

// blurred pictures + thresholds + anchors for n pictures in device memory; results land in the pinned host arrays of the state
static int scan_run(cb200_ctx* c, const uint8_t* d_pics, int w, int h, int n)
{
    ScanScratch* s = sstate(c);
    cudaStream_t st = c->stream;
    const int R = scan_blur_radius(w, h);
    if (R > 4) return fail(CB200_ERR_ARG, "pictures with a short side of 4500 pixels or more need a Gaussian kernel beyond 9 taps, which is not restated");
    const int skip = (w < h ? w : h) / 60;
    if (skip < 1) return fail(CB200_ERR_ARG, "picture smaller than 60 pixels on its short side (Scanner's row step would be 0)");
    const size_t npx = (size_t)w * (size_t)h;
    if (npx * (size_t)n > s->blur_bytes) {
        cudaFree(s->d_blur); s->d_blur = nullptr; s->blur_bytes = 0;
        CK(cudaMalloc(&s->d_blur, npx * (size_t)n), "cudaMalloc blurred pictures");
        s->blur_bytes = npx * (size_t)n;
    }
    if (n > s->thr_cap) {
        cudaFree(s->d_hist); cudaFree(s->d_thr); cudaFree(s->d_anchors); cudaFree(s->d_count); cudaFree(s->d_cutoff); cudaFree(s->d_status);
        cudaFreeHost(s->h_anchors); cudaFreeHost(s->h_count);
        s->d_hist = nullptr; s->d_thr = nullptr; s->d_anchors = nullptr; s->d_count = nullptr; s->d_cutoff = nullptr; s->d_status = nullptr;
        s->h_anchors = nullptr; s->h_count = nullptr; s->thr_cap = 0;
        CK(cudaMalloc(&s->d_hist, sizeof(unsigned) * 256 * (size_t)n), "cudaMalloc histograms");
        CK(cudaMalloc(&s->d_thr, sizeof(int) * (size_t)n), "cudaMalloc thresholds");
        CK(cudaMalloc(&s->d_anchors, sizeof(int4) * 4 * (size_t)n), "cudaMalloc anchors");
        CK(cudaMalloc(&s->d_count, sizeof(int) * (size_t)n), "cudaMalloc counts");
        CK(cudaMalloc(&s->d_cutoff, sizeof(unsigned) * (size_t)n), "cudaMalloc cutoffs");
        CK(cudaMalloc(&s->d_status, sizeof(int) * (size_t)n), "cudaMalloc status");
        CK(cudaMallocHost(&s->h_anchors, sizeof(int4) * 4 * (size_t)n), "cudaMallocHost anchors");
        CK(cudaMallocHost(&s->h_count, sizeof(int) * 3 * (size_t)n), "cudaMallocHost counts");
        s->thr_cap = n;
    }
    // rows of a pass: the primary pass scans h / skip rows, the bottom-right window at most 2 h / skip (half the step)
    const int rows_cap = 2 * ((h + skip - 1) / skip) + 4;
    if (n > s->ws_pics || rows_cap > s->ws_rows_cap) {
        cudaFree(s->d_rowbuf); cudaFree(s->d_rowcnt); cudaFree(s->d_pts); cudaFree(s->d_res); cudaFree(s->d_nres);
        s->d_rowbuf = nullptr; s->d_rowcnt = nullptr; s->d_pts = nullptr; s->d_res = nullptr; s->d_nres = nullptr; s->ws_pics = 0; s->ws_rows_cap = 0;
        const int np = n > s->ws_pics ? n : s->ws_pics, rc = rows_cap > s->ws_rows_cap ? rows_cap : s->ws_rows_cap;
        CK(cudaMalloc(&s->d_rowbuf, sizeof(Anchor) * (size_t)np * rc * kRowCap), "cudaMalloc scan rows");
        CK(cudaMalloc(&s->d_rowcnt, sizeof(int) * (size_t)np * rc), "cudaMalloc scan row counts");
        CK(cudaMalloc(&s->d_pts, sizeof(Anchor) * (size_t)np * kPtsCap), "cudaMalloc scan points");
        CK(cudaMalloc(&s->d_res, sizeof(Anchor) * (size_t)np * kPtsCap * kResCap), "cudaMalloc scan results");
        CK(cudaMalloc(&s->d_nres, sizeof(int) * (size_t)np * kPtsCap), "cudaMalloc scan result counts");
        s->ws_pics = np; s->ws_rows_cap = rc;
    }
    CK(cudaMemsetAsync(s->d_hist, 0, sizeof(unsigned) * 256 * (size_t)n, st), "memset histograms");
    // cb200_set_timing: one event set per scan -- [blur + histogram, Otsu, anchors] through cb200_get_timing
    auto mark = [&]() { if (c->timing && c->ev_count[c->cur] < 8) cudaEventRecord(c->ev[c->cur][c->ev_count[c->cur]++], st); };
    if (c->timing) { c->cur = (int)(c->calls % cb200_ctx::kEvSets); c->calls++; c->ev_count[c->cur] = 0; }
    mark();
    const dim3 bgrid((unsigned)((w + kBlurTW - 1) / kBlurTW), (unsigned)((h + kBlurTH - 1) / kBlurTH), (unsigned)n);
    // CB200_SCAN_BLUR=0 (tests, A/B): the one-pixel-per-element kernel
    const bool blur4 = !(getenv("CB200_SCAN_BLUR") && atoi(getenv("CB200_SCAN_BLUR")) == 0);
    // aligned 32-bit accesses need rows that start on a word: width a multiple of four, base pointers aligned
    const int words_ok = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(d_pics) % 4 == 0) && (reinterpret_cast<uintptr_t>(s->d_blur) % 4 == 0);
    if (blur4) {
        switch (R) {
        case 1: k_scan_blur4<1><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, words_ok, s->d_blur, s->d_hist); break;
        case 2: k_scan_blur4<2><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, words_ok, s->d_blur, s->d_hist); break;
        case 3: k_scan_blur4<3><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, words_ok, s->d_blur, s->d_hist); break;
        default: k_scan_blur4<4><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, words_ok, s->d_blur, s->d_hist); break;
        }
    } else {
        switch (R) {
        case 1: k_scan_blur<1><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, s->d_blur, s->d_hist); break;
        case 2: k_scan_blur<2><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, s->d_blur, s->d_hist); break;
        case 3: k_scan_blur<3><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, s->d_blur, s->d_hist); break;
        default: k_scan_blur<4><<<bgrid, kBlurThreads, 0, st>>>(d_pics, w, h, s->d_blur, s->d_hist); break;
        }
    }
    count_launch();
    mark();
    k_scan_otsu<<<(n + 63) / 64, 64, 0, st>>>(s->d_hist, n, (double)npx, s->d_thr); count_launch();
    mark();
    k_scan_anchors<<<n, kScanThreads, 0, st>>>(s->d_blur, s->d_thr, w, h, s->ws_rows_cap, s->d_rowbuf, s->d_rowcnt, s->d_pts, s->d_res, s->d_nres,
                                               s->d_anchors, s->d_count, s->d_cutoff, s->d_status); count_launch();
    mark();
    CK(cudaGetLastError(), "scan launch");
    CK(cudaMemcpyAsync(s->h_anchors, s->d_anchors, sizeof(int4) * 4 * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H anchors");
    CK(cudaMemcpyAsync(s->h_count, s->d_count, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H counts");
    CK(cudaMemcpyAsync(s->h_count + n, s->d_cutoff, sizeof(unsigned) * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H cutoffs");
    CK(cudaMemcpyAsync(s->h_count + 2 * (size_t)n, s->d_status, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H status");
    CK(cudaStreamSynchronize(st), "sync (scan)");
    return CB200_OK;
}

static int scan_results(cb200_ctx* c, int n, int32_t* anchors, int32_t* count, uint32_t* cutoff)
{
    const ScanScratch* s = c->scan;
    for (int i = 0; i < n; ++i) {
        const bool overflow = (s->h_count[2 * (size_t)n + i] & kScanOverflow) != 0;
        count[i] = overflow ? -1 : s->h_count[i];
        if (cutoff) cutoff[i] = (uint32_t)s->h_count[(size_t)n + i];
    }
    if (anchors) memcpy(anchors, s->h_anchors, sizeof(int32_t) * 16 * (size_t)n);
    return CB200_OK;
}

static int stage_pictures(cb200_ctx* c, const uint8_t* pics, int w, int h, int n, const uint8_t** d_out)
{
    ScanScratch* s = sstate(c);
    const size_t bytes = (size_t)w * h * 3 * (size_t)n;
    if (bytes > s->pics_bytes) {
        cudaFree(s->d_pics); s->d_pics = nullptr; s->pics_bytes = 0;
        CK(cudaMalloc(&s->d_pics, bytes), "cudaMalloc picture staging");
        s->pics_bytes = bytes;
    }
    CK(cudaMemcpyAsync(s->d_pics, pics, bytes, cudaMemcpyHostToDevice, c->stream), "H2D pictures");
    *d_out = s->d_pics;
    return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_scan_dev(cb200_ctx* c, const uint8_t* d_pictures, int w, int h, int n, int32_t* anchors, int32_t* count, uint32_t* cutoff)
{
    if (!c || !d_pictures || !count || n < 0 || w < 1 || h < 1) return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    int rc = scan_run(c, d_pictures, w, h, n); if (rc) return rc;
    return scan_results(c, n, anchors, count, cutoff);
}

int cb200_scan(cb200_ctx* c, const uint8_t* pictures, int w, int h, int n, int32_t* anchors, int32_t* count, uint32_t* cutoff)
{
    if (!c || !pictures || !count || n < 0 || w < 1 || h < 1) return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    const uint8_t* d = nullptr;
    int rc = stage_pictures(c, pictures, w, h, n, &d); if (rc) return rc;
    rc = scan_run(c, d, w, h, n); if (rc) return rc;
    return scan_results(c, n, anchors, count, cutoff);
}

int cb200_scan_blurred(cb200_ctx* c, uint8_t* blurred_out, int32_t* thresholds_out, int w, int h, int n)
{
    if (!c || !c->scan || n < 0 || (size_t)w * h * (size_t)n > c->scan->blur_bytes || n > c->scan->thr_cap) return fail(CB200_ERR_ARG, "no scan of that size to read back");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    if (blurred_out) CK(cudaMemcpyAsync(blurred_out, c->scan->d_blur, (size_t)w * h * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "D2H blurred");
    if (thresholds_out) CK(cudaMemcpyAsync(thresholds_out, c->scan->d_thr, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "D2H thresholds");
    CK(cudaStreamSynchronize(c->stream), "sync");
    return CB200_OK;
}

int cb200_scan_extract_decode_fountain(cb200_ctx* c, const uint8_t* pictures, int w, int h, int n, uint32_t flags,
                                       uint8_t* chunks_out, uint32_t* chunk_count, uint32_t* chunk_mask, uint8_t* frame_flags,
                                       int32_t* extract_status)
{
    if (!c || !pictures || !chunks_out || !chunk_count || !extract_status || n < 0 || n > c->max_frames || w < 2 || h < 2)
        return fail(CB200_ERR_ARG, "bad arguments");
    if (n == 0) return CB200_OK;
    CK(cudaSetDevice(c->device), "cudaSetDevice");
    const Mode& m = c->mode;
    const uint8_t* d = nullptr;
    int rc = stage_pictures(c, pictures, w, h, n, &d); if (rc) return rc;
    rc = scan_run(c, d, w, h, n); if (rc) return rc;
    const ScanScratch* s = c->scan;
    // Extractor::extract (Extractor.h:30-46): fewer than four anchors -> FAILURE; Corners = the anchors' centres
    // (Corners.h:13-16); NEEDS_SHARPEN unless every side of the quadrilateral is longer than the frame (is_granular_scale, :57-75).
    // Failed pictures are deskewed with the identity (their output is discarded): one batch, no compaction.
    std::vector<float> corners((size_t)n * 8);
    const float an = 30.0f;
    const float ident[8] = {an, an, (float)m.width - an, an, an, (float)m.height - an, (float)m.width - an, (float)m.height - an};
    for (int i = 0; i < n; ++i) {
        const bool overflow = (s->h_count[2 * (size_t)n + i] & kScanOverflow) != 0;
        const Anchor* a = reinterpret_cast<const Anchor*>(s->h_anchors + 4 * (size_t)i);
        float* cr = corners.data() + 8 * (size_t)i;
        if (overflow || s->h_count[i] < 4) { extract_status[i] = overflow ? -1 : 0; memcpy(cr, ident, sizeof(ident)); continue; }
        int xy[8];
        for (int k = 0; k < 4; ++k) { xy[2 * k] = xavg(a[k]); xy[2 * k + 1] = yavg(a[k]); cr[2 * k] = (float)xy[2 * k]; cr[2 * k + 1] = (float)xy[2 * k + 1]; }
        const int pairs[4][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}};              // tl-tr, tr-br, br-bl, bl-tl
        bool granular = true;
        for (int k = 0; k < 4; ++k) {
            const int p = pairs[k][0], q = pairs[k][1];
            granular = granular && (iabs(xy[2 * p] - xy[2 * q]) > m.width || iabs(xy[2 * p + 1] - xy[2 * q + 1]) > m.height);
        }
        extract_status[i] = granular ? 1 : 2;
        double m9[9];
        if (cb200_perspective_transform(cr, ident, m9) != CB200_OK) { extract_status[i] = 0; memcpy(cr, ident, sizeof(ident)); }   // collinear anchors
    }
    rc = cb200_extract_decode_fountain_dev(c, d, w, h, n, corners.data(), flags, chunks_out, chunk_count, chunk_mask, frame_flags);
    if (rc) return rc;
    for (int i = 0; i < n; ++i)
        if (extract_status[i] <= 0) { chunk_count[i] = 0; if (chunk_mask) chunk_mask[i] = 0; }
    return CB200_OK;
}

}  // extern "C"
