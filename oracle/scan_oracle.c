/* scan_oracle.c -- TEST INFRASTRUCTURE (see cimbar_oracle.c): CPU restatement of the reference's anchor scan, the step of
 * the extractor that runs before the deskew.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may use it; nothing under libcimbar_b200/ does.
 *
 * Restates, in plain C99 (reference file:line relative to /root/reference/src/lib/extractor/):
 *   Scanner::preprocess_image (fast)   Scanner.h:146-166, :124-128   cvtColor(RGB2GRAY) + GaussianBlur(unit, sigma 0) + Otsu
 *   ... (fast = false)                 Scanner.h:130-136              adaptiveThreshold(MEAN_C, unit' x unit', -10) on the blurred picture
 *   ScanState / _114 / _122            ScanState.h:9-122
 *   Anchor                             Anchor.h:8-112
 *   scan_horizontal/vertical/diagonal  Scanner.h:176-276
 *   t1..t4, on_t1_scan                 Scanner.h:278-406
 *   filter_candidates, sort_top_to_bottom, add_bottom_right_corner, scan_primary, scan     Scanner.cpp:83-199
 *   deduplicate_candidates             Scanner.cpp:61-81 (only used by the reference's piecemeal test)
 * OpenCV (a third-party dependency of the reference, 4.x) arithmetic restated and pinned against cv2 4.13 in
 * tests/test_scan_oracle.py: 8-bit GaussianBlur = fixed-point separable filter with the small-kernel table
 * (3: 64 128 64, 5: 16 64 96 64 16, 7: 8 28 56 72 56 28 8, 9: 4 13 30 51 60 51 30 13 4, all /256), BORDER_REFLECT_101,
 * (sum + 2^15) >> 16; THRESH_OTSU = getThreshVal_Otsu_8u in double precision.  Kernels beyond 9 taps (images whose short
 * side is >= 4500 pixels) are not restated: cbo_scan_preprocess returns -1.
 * libstdc++'s std::sort (introsort + final insertion sort) is restated literally: candidates of equal size keep the
 * order that algorithm gives them.
 * Pinned by the reference's own golden strings: extractor/test/ScannerTest.cpp:16-251 (tests/test_scan_oracle.py).
 */
#include "scan_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ preprocessing */
static unsigned next_pow2_plus_one(unsigned v)          /* Scanner.h:93-103 */
{
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    unsigned r = v + 2;
    return r < 3u ? 3u : r;
}

int cbo_scan_blur_size(int w, int h)                    /* Scanner.h:155-157 */
{
    unsigned unit = (unsigned)(w < h ? w : h);
    unit = next_pow2_plus_one((unsigned)(unit * 0.002));
    return (int)(unit < 3u ? 3u : unit);
}

static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

static const int* blur_kernel(int ksize)
{
    static const int k3[3] = {64, 128, 64}, k5[5] = {16, 64, 96, 64, 16}, k7[7] = {8, 28, 56, 72, 56, 28, 8},
                     k9[9] = {4, 13, 30, 51, 60, 51, 30, 13, 4};
    switch (ksize) { case 3: return k3; case 5: return k5; case 7: return k7; case 9: return k9; default: return NULL; }
}

int cbo_scan_gaussian_blur(const uint8_t* src, int w, int h, int ksize, uint8_t* dst)
{
    const int* k = blur_kernel(ksize);
    if (!k) return -1;
    const int r = ksize / 2;
    uint16_t* hs = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned s = 0;
            for (int i = 0; i < ksize; ++i) s += (unsigned)k[i] * src[(size_t)y * w + reflect101(x + i - r, w)];
            hs[(size_t)y * w + x] = (uint16_t)s;            /* <= 255 * 256: no saturation */
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned s = 0;
            for (int i = 0; i < ksize; ++i) s += (unsigned)k[i] * hs[(size_t)reflect101(y + i - r, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)((s + 32768u) >> 16);
        }
    free(hs);
    return 0;
}

int cbo_scan_otsu(const uint8_t* img, size_t n)          /* cv::threshold(THRESH_OTSU): getThreshVal_Otsu_8u */
{
    int hist[256];
    memset(hist, 0, sizeof(hist));
    for (size_t i = 0; i < n; ++i) hist[img[i]]++;
    double mu = 0, scale = 1. / (double)n;
    for (int i = 0; i < 256; ++i) mu += i * (double)hist[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; ++i) {
        double p_i = hist[i] * scale, q2, mu2, sigma;
        mu1 *= q1;
        q1 += p_i;
        q2 = 1. - q1;
        if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        mu2 = (mu - q1 * mu1) / q2;
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return (int)max_val;
}

/* Scanner::preprocess_image(img, fast=true): bin[i] = 255 where blurred gray > otsu.  Returns the threshold, -1 if the
 * blur kernel for this image size is not restated.  blurred (optional) receives the blurred gray image. */
int cbo_scan_preprocess(const uint8_t* rgb, int w, int h, uint8_t* bin, uint8_t* blurred)
{
    const int ksize = cbo_scan_blur_size(w, h);
    if (!blur_kernel(ksize)) return -1;
    const size_t n = (size_t)w * (size_t)h;
    uint8_t* gray = (uint8_t*)malloc(n);
    uint8_t* bl = blurred ? blurred : (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; ++i)                          /* cvtColor(RGB2GRAY), 8 bit (cimbar_oracle.c cbo_gray) */
        gray[i] = (uint8_t)((9798u * rgb[3 * i] + 19235u * rgb[3 * i + 1] + 3735u * rgb[3 * i + 2] + 16384u) >> 15);
    cbo_scan_gaussian_blur(gray, w, h, ksize, bl);
    const int t = cbo_scan_otsu(bl, n);
    for (size_t i = 0; i < n; ++i) bin[i] = bl[i] > t ? 255 : 0;
    free(gray);
    if (!blurred) free(bl);
    return t;
}

/* Scanner::preprocess_image(img, fast=false) (Scanner.h:130-136, :146-166): the blurred gray image through
 * adaptiveThreshold(255, ADAPTIVE_THRESH_MEAN_C, THRESH_BINARY, unit, -10) with unit = nextPowerOfTwoPlusOne(unsigned(min(cols, rows) * 0.05)).
 * OpenCV: mean = boxFilter(src, unit x unit, normalised, BORDER_REPLICATE) rounded to nearest (an odd area has no ties),
 * dst = 255 where src - mean > 10 (delta = -10: tab[src - mean + 255] with idelta = cvCeil(delta)).  Pinned against cv2 in
 * tests/test_scan_oracle.py.  Not on Extractor::extract's path (it constructs Scanner(img), fast = true); the device scan rejects it. */
int cbo_scan_preprocess_adaptive(const uint8_t* rgb, int w, int h, uint8_t* bin)
{
    const int ksize = cbo_scan_blur_size(w, h);
    if (!blur_kernel(ksize)) return -1;
    const size_t n = (size_t)w * (size_t)h;
    uint8_t* gray = (uint8_t*)malloc(n);
    uint8_t* bl = (uint8_t*)malloc(n);
    for (size_t i = 0; i < n; ++i)
        gray[i] = (uint8_t)((9798u * rgb[3 * i] + 19235u * rgb[3 * i + 1] + 3735u * rgb[3 * i + 2] + 16384u) >> 15);
    cbo_scan_gaussian_blur(gray, w, h, ksize, bl);
    unsigned unit = (unsigned)(w < h ? w : h);
    const int bs = (int)next_pow2_plus_one((unsigned)(unit * 0.05));
    const int r = bs / 2;
    const long long area = (long long)bs * bs;
    /* integral image of the replicate-padded picture */
    const int pw = w + 2 * r, ph = h + 2 * r;
    long long* ii = (long long*)calloc((size_t)(pw + 1) * (size_t)(ph + 1), sizeof(long long));
    for (int y = 0; y < ph; ++y) {
        int sy = y - r; if (sy < 0) sy = 0; if (sy > h - 1) sy = h - 1;
        long long rowsum = 0;
        for (int x = 0; x < pw; ++x) {
            int sx = x - r; if (sx < 0) sx = 0; if (sx > w - 1) sx = w - 1;
            rowsum += bl[(size_t)sy * w + sx];
            ii[(size_t)(y + 1) * (pw + 1) + (x + 1)] = ii[(size_t)y * (pw + 1) + (x + 1)] + rowsum;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const long long S = ii[(size_t)(y + bs) * (pw + 1) + (x + bs)] - ii[(size_t)y * (pw + 1) + (x + bs)]
                              - ii[(size_t)(y + bs) * (pw + 1) + x] + ii[(size_t)y * (pw + 1) + x];
            const int mean = (int)((2 * S + area) / (2 * area));
            bin[(size_t)y * w + x] = ((int)bl[(size_t)y * w + x] - mean > 10) ? 255 : 0;
        }
    free(ii); free(gray); free(bl);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ Anchor (Anchor.h) */
static int a_xavg(const cbo_anchor* a) { return (a->x + a->xmax) / 2; }
static int a_yavg(const cbo_anchor* a) { return (a->y + a->ymax) / 2; }
static int a_xrange(const cbo_anchor* a) { return abs(a->x - a->xmax) / 2; }
static int a_yrange(const cbo_anchor* a) { return abs(a->y - a->ymax) / 2; }
static unsigned long long a_size(const cbo_anchor* a)
{
    return (unsigned long long)(pow((double)(a->x - a->xmax), 2) + pow((double)(a->y - a->ymax), 2));
}
static int a_max_range(const cbo_anchor* a)
{
    int dx = abs(a->x - a->xmax), dy = abs(a->y - a->ymax);
    return dx > dy ? dx : dy;
}
static void a_merge(cbo_anchor* a, const cbo_anchor* o)
{
    if (o->x < a->x) a->x = o->x;
    if (o->xmax > a->xmax) a->xmax = o->xmax;
    if (o->y < a->y) a->y = o->y;
    if (o->ymax > a->ymax) a->ymax = o->ymax;
}
static int a_is_mergeable(const cbo_anchor* a, const cbo_anchor* rhs, int max_distance)     /* Anchor.h:84-91 */
{
    if (abs(a_xavg(a) - a_xavg(rhs)) > max_distance || abs(a_yavg(a) - a_yavg(rhs)) > max_distance) return 0;
    int ratio = a_max_range(rhs) * 10 / a_max_range(a);
    return ratio > 6 && ratio < 17;
}
static cbo_anchor mk(int x, int xmax, int y, int ymax) { cbo_anchor a = {x, xmax, y, ymax}; return a; }

/* ------------------------------------------------------------------------------------------------ ScanState (ScanState.h) */
typedef struct { int state; int tally[8]; int n; const float (*limits)[2]; } scan_state;
static const float LIM_114[6][2] = {{0, 0}, {3.0f, 6.0f}, {3.0f, 6.0f}, {0, 0}, {3.0f, 6.0f}, {3.0f, 6.0f}};
static const float LIM_122[6][2] = {{0, 0}, {1.0f, 3.0f}, {0.5f, 1.5f}, {0, 0}, {0.5f, 1.5f}, {1.0f, 3.0f}};

static void ss_init(scan_state* s, int kind) { s->state = 0; s->tally[0] = 0; s->n = 1; s->limits = kind == 122 ? LIM_122 : LIM_114; }

static int ss_evaluate(const scan_state* s)
{
    if (s->state != 6) return -1;
    for (int i = 1; i <= 5; ++i) if (s->tally[i] == 0) return -1;
    float center = (float)s->tally[3];
    for (int i = 1; i <= 5; ++i) {
        if (i == 3) continue;
        float ratio_min = center / (float)(s->tally[i] + 1);
        int d = s->tally[i] - 1; if (d < 1) d = 1;
        float ratio_max = center / (float)d;
        if (ratio_max < s->limits[i][0] || ratio_min > s->limits[i][1]) return -1;
    }
    int size = 0;
    for (int i = 1; i <= 5; ++i) size += s->tally[i];
    return size;
}

static int ss_process(scan_state* s, int active)
{
    const int even = s->state == 0 || s->state == 2 || s->state == 4;
    const int odd = s->state == 1 || s->state == 3 || s->state == 5;
    if ((even && active) || (odd && !active)) {
        s->state += 1;
        s->tally[s->n++] = 1;
        if (s->state == 6) {
            int res = ss_evaluate(s);
            s->state -= 2;                                  /* pop_state: two pop_front */
            memmove(s->tally, s->tally + 2, sizeof(int) * (size_t)(s->n - 2));
            s->n -= 2;
            return res;
        }
        return -1;
    }
    if (odd && active) s->tally[s->n - 1] += 1;
    if (!active && (s->state == 2 || s->state == 4)) s->tally[s->n - 1] += 1;
    return -1;
}

/* ------------------------------------------------------------------------------------------------ Scanner */
typedef struct { cbo_anchor* v; int n, cap; } alist;
static void al_push(alist* l, cbo_anchor a)
{
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 64; l->v = (cbo_anchor*)realloc(l->v, sizeof(cbo_anchor) * (size_t)l->cap); }
    l->v[l->n++] = a;
}

void cbo_scanner_init(cbo_scanner* s, const uint8_t* bin, int w, int h, int skip)      /* Scanner.h:168-174 (dark = true) */
{
    s->bin = bin; s->w = w; s->h = h;
    s->skip = skip ? skip : (h < w ? h : w) / 60;
    s->merge_cutoff = w / 30;
    s->anchor_size = 30;
}

/* Scanner::test_pixel (dark): pixel > 127.  The reference reads outside the image when a confirm line of t4 lies one past an
 * image edge (undefined behaviour there); here such pixels count as inactive. */
static int test_pixel(const cbo_scanner* s, int x, int y)
{
    if (x < 0 || y < 0 || x >= s->w || y >= s->h) return 0;
    return s->bin[(size_t)y * s->w + x] > 127;
}

static int scan_horizontal(const cbo_scanner* s, int kind, alist* points, int y, int xstart, int xend)
{
    if (xstart < 0) xstart = 0;
    if (xend < 0 || xend > s->w) xend = s->w;
    const int init = points->n;
    scan_state st; ss_init(&st, kind);
    for (int x = xstart; x < xend; ++x) {
        int res = ss_process(&st, test_pixel(s, x, y));
        if (res > 0) al_push(points, mk(x - res, x - 1, y, y));
    }
    int res = ss_process(&st, 0);
    if (res > 0) { int x = xend; al_push(points, mk(x - res, x - 1, y, y)); }
    return init != points->n;
}

static int scan_vertical(const cbo_scanner* s, int kind, alist* points, int x, int xmax, int ystart, int yend)
{
    if (xmax < 0) xmax = x;
    const int xavg = (x + xmax) / 2;
    if (ystart < 0) ystart = 0;
    if (yend < 0 || yend > s->h) yend = s->h;
    const int init = points->n;
    scan_state st; ss_init(&st, kind);
    for (int y = ystart; y < yend; ++y) {
        int res = ss_process(&st, test_pixel(s, xavg, y));
        if (res > 0) al_push(points, mk(xavg, xavg, y - res, y - 1));
    }
    int res = ss_process(&st, 0);
    if (res > 0) { int y = yend; al_push(points, mk(xavg, xavg, y - res, y - 1)); }
    return init != points->n;
}

static int scan_diagonal(const cbo_scanner* s, int kind, alist* points, int xstart, int xend, int ystart, int yend)
{
    if (xend > s->w) xend = s->w;
    if (yend > s->h) yend = s->h;
    if (xstart < 0) { int off = -xstart; xstart += off; ystart += off; }
    if (ystart < 0) { int off = -ystart; xstart += off; ystart += off; }
    const int init = points->n;
    scan_state st; ss_init(&st, kind);
    int x = xstart, y = ystart;
    for (; x < xend && y < yend; ++x, ++y) {
        int res = ss_process(&st, test_pixel(s, x, y));
        if (res > 0) al_push(points, mk(x - res, x - 1, y - res, y - 1));
    }
    int res = ss_process(&st, 0);
    if (res > 0) al_push(points, mk(x - res, x - 1, y - res, y - 1));
    return init != points->n;
}

static void t1_scan_rows(const cbo_scanner* s, int kind, alist* points, int skip, int y, int yend, int xstart, int xend)
{
    if (skip <= 0) skip = s->skip;
    if (y < 0) y = skip;
    if (yend < 0 || yend > s->h) yend = s->h;
    for (; y < yend; y += skip) scan_horizontal(s, kind, points, y, xstart, xend);
}

static void t2_scan_column(const cbo_scanner* s, int kind, const cbo_anchor* hint, alist* out)
{
    int ystart = hint->y - 3 * a_xrange(hint), yend = hint->ymax + 3 * a_xrange(hint);
    scan_vertical(s, kind, out, hint->x, hint->xmax, ystart, yend);
}

static int t3_scan_diagonal(const cbo_scanner* s, int kind, const cbo_anchor* hint, cbo_anchor* out)
{
    alist confirms = {0};
    int xstart = a_xavg(hint) - 2 * a_yrange(hint), xend = a_xavg(hint) + 2 * a_yrange(hint);
    int ystart = hint->y - a_yrange(hint), yend = hint->ymax + a_yrange(hint);
    int found = 0;
    if (scan_diagonal(s, kind, &confirms, xstart, xend, ystart, yend)) {
        cbo_anchor merged = *hint;
        for (int i = 0; i < confirms.n; ++i)
            if (a_is_mergeable(&confirms.v[i], hint, s->merge_cutoff)) { found = 1; a_merge(&merged, &confirms.v[i]); }
        if (found) *out = merged;
    }
    free(confirms.v);
    return found;
}

static int t4_confirm_scan(const cbo_scanner* s, int kind, cbo_anchor hint, int merge_confirms, cbo_anchor* out)
{
    {
        alist confirms = {0};
        int xstart = hint.x - a_xrange(&hint), xend = hint.xmax + a_xrange(&hint), yavg = a_yavg(&hint);
        for (int d = -1; d <= 1; ++d)
            if (!scan_horizontal(s, kind, &confirms, yavg + d, xstart, xend)) { free(confirms.v); return 0; }
        int confirm = 0;
        for (int i = 0; i < confirms.n; ++i)
            if (a_is_mergeable(&confirms.v[i], &hint, s->merge_cutoff)) {
                confirm = 1;
                if (!merge_confirms) break;
                a_merge(&hint, &confirms.v[i]);
            }
        free(confirms.v);
        if (!confirm) return 0;
    }
    {
        alist confirms = {0};
        int ystart = hint.y - a_yrange(&hint), yend = hint.ymax + a_yrange(&hint), xavg = a_xavg(&hint);
        for (int d = -1; d <= 1; ++d)
            if (!scan_vertical(s, kind, &confirms, xavg + d, xavg + d, ystart, yend)) { free(confirms.v); return 0; }
        int confirm = 0;
        for (int i = 0; i < confirms.n; ++i)
            if (a_is_mergeable(&confirms.v[i], &hint, s->merge_cutoff)) {
                confirm = 1;
                if (!merge_confirms) break;
                a_merge(&hint, &confirms.v[i]);
            }
        free(confirms.v);
        if (!confirm) return 0;
    }
    *out = hint;
    return 1;
}

static void on_t1_scan(const cbo_scanner* s, int kind, const cbo_anchor* found, alist* candidates, int merge_confirms)
{
    for (int i = 0; i < candidates->n; ++i)
        if (a_is_mergeable(&candidates->v[i], found, s->merge_cutoff)) return;
    alist p2 = {0};
    t2_scan_column(s, kind, found, &p2);
    for (int i = 0; i < p2.n; ++i) {
        cbo_anchor p3, p4;
        if (t3_scan_diagonal(s, kind, &p2.v[i], &p3) && t4_confirm_scan(s, kind, p3, merge_confirms, &p4)) al_push(candidates, p4);
    }
    free(p2.v);
}

/* ---- std::sort(first, last, size_sort()) of libstdc++ (bits/stl_algo.h), restated literally */
static int size_gt(const cbo_anchor* a, const cbo_anchor* b) { return a_size(a) > a_size(b); }
static void swp(cbo_anchor* a, cbo_anchor* b) { cbo_anchor t = *a; *a = *b; *b = t; }

static void adjust_heap(cbo_anchor* first, long hole, long len, cbo_anchor value)
{
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (size_gt(first + second, first + (second - 1))) second--;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    long parent = (hole - 1) / 2;                           /* __push_heap */
    while (hole > top && size_gt(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void heap_sort_all(cbo_anchor* first, cbo_anchor* last)     /* __partial_sort(first, last, last) */
{
    const long len = last - first;
    if (len >= 2)
        for (long parent = (len - 2) / 2;; --parent) { adjust_heap(first, parent, len, first[parent]); if (parent == 0) break; }
    while (last - first > 1) { --last; cbo_anchor v = *last; *last = *first; adjust_heap(first, 0, last - first, v); }
}
static void move_median_to_first(cbo_anchor* result, cbo_anchor* a, cbo_anchor* b, cbo_anchor* c)
{
    if (size_gt(a, b)) { if (size_gt(b, c)) swp(result, b); else if (size_gt(a, c)) swp(result, c); else swp(result, a); }
    else if (size_gt(a, c)) swp(result, a);
    else if (size_gt(b, c)) swp(result, c);
    else swp(result, b);
}
static cbo_anchor* unguarded_partition(cbo_anchor* first, cbo_anchor* last, cbo_anchor* pivot)
{
    for (;;) {
        while (size_gt(first, pivot)) ++first;
        --last;
        while (size_gt(pivot, last)) --last;
        if (!(first < last)) return first;
        swp(first, last);
        ++first;
    }
}
static void introsort_loop(cbo_anchor* first, cbo_anchor* last, long depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) { heap_sort_all(first, last); return; }
        --depth_limit;
        cbo_anchor* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        cbo_anchor* cut = unguarded_partition(first + 1, last, first);
        introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static void unguarded_linear_insert(cbo_anchor* last)
{
    cbo_anchor val = *last;
    cbo_anchor* next = last - 1;
    while (size_gt(&val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort(cbo_anchor* first, cbo_anchor* last)
{
    if (first == last) return;
    for (cbo_anchor* i = first + 1; i != last; ++i) {
        if (size_gt(i, first)) { cbo_anchor val = *i; memmove(first + 1, first, sizeof(cbo_anchor) * (size_t)(i - first)); *first = val; }
        else unguarded_linear_insert(i);
    }
}
static void std_sort_by_size(cbo_anchor* first, cbo_anchor* last)
{
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) != 0) ++lg;                      /* std::__lg */
    introsort_loop(first, last, lg * 2);
    if (last - first > 16) {
        insertion_sort(first, first + 16);
        for (cbo_anchor* i = first + 16; i != last; ++i) unguarded_linear_insert(i);
    } else insertion_sort(first, last);
}

static unsigned filter_candidates(alist* c)                /* Scanner.cpp:83-105 */
{
    if (c->n < 3) return 0;
    std_sort_by_size(c->v, c->v + c->n);
    unsigned cutoff = 0;
    for (int i = 0; i < 3; ++i) cutoff += (unsigned)a_size(&c->v[i]);
    cutoff /= 8;
    int i = 0;
    for (; i < c->n; ++i) if (a_size(&c->v[i]) < cutoff) break;
    if (i > 3) i = 3;
    if (i < c->n) c->n = i;
    return cutoff;
}

static int fix3(int i) { if (i < 0) i = 2; else if (i >= 3) i = 0; return i; }

static int sort_top_to_bottom(alist* c)                    /* Scanner.cpp:107-137 */
{
    if (c->n < 3) return 0;
    int cx[3], cy[3];
    for (int i = 0; i < 3; ++i) { cx[i] = a_xavg(&c->v[i]); cy[i] = a_yavg(&c->v[i]); }
    int ex[3] = {cx[1] - cx[2], cx[2] - cx[0], cx[0] - cx[1]}, ey[3] = {cy[1] - cy[2], cy[2] - cy[0], cy[0] - cy[1]};
    int top_left = 0, max_d = 0;
    for (int i = 0; i < 3; ++i) { int d = ex[i] * ex[i] + ey[i] * ey[i]; if (d > max_d) { top_left = i; max_d = d; } }
    const int dep = fix3(top_left - 1), inc = fix3(top_left + 1);
    const int dx = ex[dep], dy = ey[dep];
    const int ix = -ey[inc], iy = ex[inc];                  /* rotated incoming edge */
    const int ox = dx - ix, oy = dy - iy;
    int top_right, bottom_left;
    if (ox * ox + oy * oy < dx * dx + dy * dy) { top_right = fix3(top_left + 1); bottom_left = fix3(top_left - 1); }
    else { top_right = fix3(top_left - 1); bottom_left = fix3(top_left + 1); }
    cbo_anchor a = c->v[top_left], b = c->v[top_right], d = c->v[bottom_left];
    c->v[0] = a; c->v[1] = b; c->v[2] = d; c->n = 3;
    return 1;
}

static int imax(int a, int b) { return a > b ? a : b; }

static int add_bottom_right_corner(const cbo_scanner* s, alist* anchors, unsigned cutoff)      /* Scanner.cpp:139-180 */
{
    const cbo_anchor* A = anchors->v;
    double topScalar = a_max_range(&A[2]) / fmax((double)a_max_range(&A[1]), (double)a_max_range(&A[0]));
    int tex = (int)((a_xavg(&A[1]) - a_xavg(&A[0])) * topScalar), tey = (int)((a_yavg(&A[1]) - a_yavg(&A[0])) * topScalar);
    int g1x = a_xavg(&A[2]) + tex, g1y = a_yavg(&A[2]) + tey;
    double leftScalar = a_max_range(&A[1]) / fmax((double)a_max_range(&A[2]), (double)a_max_range(&A[0]));
    int lex = (int)((a_xavg(&A[2]) - a_xavg(&A[0])) * leftScalar), ley = (int)((a_yavg(&A[2]) - a_yavg(&A[0])) * leftScalar);
    int g2x = a_xavg(&A[1]) + lex, g2y = a_yavg(&A[1]) + ley;
    int cx = (g1x + g2x) / 2, cy = (g1y + g2y) / 2;
    float uncertainty = 2;
    int range = (int)(imax(imax(a_max_range(&A[0]), a_max_range(&A[1])), a_max_range(&A[2])) * uncertainty);
    int skip = s->skip / 2;
    int ystart = cy - range, yend = cy + range, xstart = cx - range, xend = cx + range;
    alist points = {0}, candidates = {0};
    t1_scan_rows(s, 122, &points, skip, ystart, yend, xstart, xend);
    for (int i = 0; i < points.n; ++i) on_t1_scan(s, 122, &points.v[i], &candidates, 0);
    int ok = 0;
    for (int i = 0; i < candidates.n; ++i)
        if (a_size(&candidates.v[i]) > cutoff) { al_push(anchors, candidates.v[i]); ok = 1; break; }
    free(points.v); free(candidates.v);
    return ok;
}

static unsigned scan_primary(const cbo_scanner* s, alist* candidates)      /* Scanner.cpp:182-191 */
{
    alist points = {0};
    t1_scan_rows(s, 114, &points, -1, -1, -1, -1, -1);
    for (int i = 0; i < points.n; ++i) on_t1_scan(s, 114, &points.v[i], candidates, 1);
    free(points.v);
    unsigned cutoff = filter_candidates(candidates);
    sort_top_to_bottom(candidates);
    return cutoff;
}

/* ------------------------------------------------------------------------------------------------ C API for the tests */
static int copy_out(alist* l, cbo_anchor* out, int cap)
{
    int n = l->n < cap ? l->n : cap;
    if (n > 0) memcpy(out, l->v, sizeof(cbo_anchor) * (size_t)n);
    int total = l->n;
    free(l->v);
    return total;
}

int cbo_scan_t1(const cbo_scanner* s, int kind, int skip, int y, int yend, int xstart, int xend, cbo_anchor* out, int cap)
{
    alist l = {0};
    t1_scan_rows(s, kind, &l, skip, y, yend, xstart, xend);
    return copy_out(&l, out, cap);
}
int cbo_scan_t2(const cbo_scanner* s, int kind, const cbo_anchor* hint, cbo_anchor* out, int cap)
{
    alist l = {0};
    t2_scan_column(s, kind, hint, &l);
    return copy_out(&l, out, cap);
}
int cbo_scan_t3(const cbo_scanner* s, int kind, const cbo_anchor* hint, cbo_anchor* out) { return t3_scan_diagonal(s, kind, hint, out); }
int cbo_scan_t4(const cbo_scanner* s, int kind, const cbo_anchor* hint, int merge_confirms, cbo_anchor* out)
{
    return t4_confirm_scan(s, kind, *hint, merge_confirms, out);
}
int cbo_scan_deduplicate(const cbo_scanner* s, const cbo_anchor* in, int n, cbo_anchor* out)      /* Scanner.cpp:61-81 */
{
    int m = 0;
    for (int i = 0; i < n; ++i) {
        int found = 0;
        for (int j = 0; j < m; ++j)
            if (a_is_mergeable(&out[j], &in[i], s->merge_cutoff)) { found = 1; a_merge(&out[j], &in[i]); break; }
        if (!found) out[m++] = in[i];
    }
    return m;
}
int cbo_scan_filter(cbo_anchor* v, int n, unsigned* cutoff)
{
    alist l = {v, n, n};
    *cutoff = filter_candidates(&l);
    return l.n;
}
int cbo_scan_sort_top_to_bottom(cbo_anchor* v, int n)
{
    alist l = {v, n, n};
    return sort_top_to_bottom(&l) ? l.n : -1;
}
int cbo_scan_primary(const cbo_scanner* s, cbo_anchor* out, int cap, unsigned* cutoff)
{
    alist c = {0};
    *cutoff = scan_primary(s, &c);
    return copy_out(&c, out, cap);
}
int cbo_scan_bottom_right(const cbo_scanner* s, cbo_anchor* anchors /* [4], 3 valid */, unsigned cutoff)
{
    alist a = {0};
    for (int i = 0; i < 3; ++i) al_push(&a, anchors[i]);
    int ok = add_bottom_right_corner(s, &a, cutoff);
    if (ok) anchors[3] = a.v[3];
    free(a.v);
    return ok;
}
/* Scanner::scan (Scanner.cpp:193-199) on a binarised image; returns the number of anchors (out: up to 4) */
int cbo_scan_bin(const uint8_t* bin, int w, int h, cbo_anchor* out, unsigned* cutoff_out)
{
    cbo_scanner s;
    cbo_scanner_init(&s, bin, w, h, 0);
    alist c = {0};
    unsigned cutoff = scan_primary(&s, &c);
    if (c.n == 3 && cutoff != 0) add_bottom_right_corner(&s, &c, cutoff);
    if (cutoff_out) *cutoff_out = cutoff;
    return copy_out(&c, out, 4);
}
/* Scanner(img).scan() from RGB pixels; -1 if the image size is not supported by the restated blur */
int cbo_scan(const uint8_t* rgb, int w, int h, cbo_anchor* out, unsigned* cutoff_out)
{
    uint8_t* bin = (uint8_t*)malloc((size_t)w * (size_t)h);
    int n = -1;
    if (cbo_scan_preprocess(rgb, w, h, bin, NULL) >= 0) n = cbo_scan_bin(bin, w, h, out, cutoff_out);
    free(bin);
    return n;
}
/* Scanner(img, fast).scan(): fast = 0 is the adaptive-threshold variant (ScannerTest/testExampleScan.Adaptive) */
int cbo_scan2(const uint8_t* rgb, int w, int h, int fast, cbo_anchor* out, unsigned* cutoff_out)
{
    if (fast) return cbo_scan(rgb, w, h, out, cutoff_out);
    uint8_t* bin = (uint8_t*)malloc((size_t)w * (size_t)h);
    int n = -1;
    if (cbo_scan_preprocess_adaptive(rgb, w, h, bin) >= 0) n = cbo_scan_bin(bin, w, h, out, cutoff_out);
    free(bin);
    return n;
}
/* Anchor::center() of the four anchors (Corners(anchors), Corners.h:13-16) and Corners::is_granular_scale (:57-66) */
void cbo_scan_corners(const cbo_anchor* a4, int* xy8)
{
    for (int i = 0; i < 4; ++i) { xy8[2 * i] = a_xavg(&a4[i]); xy8[2 * i + 1] = a_yavg(&a4[i]); }
}
int cbo_scan_is_granular_scale(const int* xy8, int min_w, int min_h)
{
    /* order of xy8: top_left, top_right, bottom_left, bottom_right */
    const int tl = 0, tr = 1, bl = 2, br = 3;
    const int pairs[4][2] = {{tl, tr}, {tr, br}, {br, bl}, {bl, tl}};
    for (int i = 0; i < 4; ++i) {
        int a = pairs[i][0], b = pairs[i][1];
        if (!(abs(xy8[2 * a] - xy8[2 * b]) > min_w || abs(xy8[2 * a + 1] - xy8[2 * b + 1]) > min_h)) return 0;
    }
    return 1;
}
