// scan_core.cuh -- the per-thread pieces of the device anchor scan (scan.cu): the ScanState machine, the three line scans,
// the t2 -> t3 -> t4 confirmation chain of one t1 hit, and the sequential tail (candidate acceptance, libstdc++'s std::sort,
// corner ordering, the bottom-right search window).  Everything here is a plain function of (image, arguments), marked
// __host__ __device__ so that tests/cpp/scan_core_host_test.cpp can run the very same code on the CPU, thread by thread
// and phase by phase, against the oracle (there is no GPU where the code is written); the product only ever calls it from
// the kernels in scan.cu.
//
// Replaces (reference file:line relative to /root/reference/src/lib/extractor/):
//   ScanState, ScanState_114, ScanState_122     ScanState.h:9-122
//   Anchor                                      Anchor.h:8-112
//   Scanner::test_pixel                         Scanner.cpp:51-58 (dark)
//   scan_horizontal / scan_vertical / scan_diagonal     Scanner.h:176-276
//   t2_scan_column, t3_scan_diagonal, t4_confirm_scan, on_t1_scan     Scanner.h:297-406
//   filter_candidates, sort_top_to_bottom, add_bottom_right_corner (its arithmetic)     Scanner.cpp:83-180
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb200 {
namespace scan {

// IEEE single / double operations that must not be contracted or approximated on the device
#if defined(__CUDA_ARCH__)
CB_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
CB_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
CB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
CB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
CB_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
CB_HD float fdiv(float a, float b) { return a / b; }
CB_HD double ddiv(double a, double b) { return a / b; }
CB_HD double dmul(double a, double b) { return a * b; }
CB_HD double dadd(double a, double b) { return a + b; }
CB_HD double dsub(double a, double b) { return a - b; }
#endif

CB_HD int iabs(int v) { return v < 0 ? -v : v; }
CB_HD int imax(int a, int b) { return a > b ? a : b; }
CB_HD int imin(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------- Anchor (Anchor.h)
struct Anchor { int x, xmax, y, ymax; };
CB_HD Anchor mk(int x, int xmax, int y, int ymax) { Anchor a; a.x = x; a.xmax = xmax; a.y = y; a.ymax = ymax; return a; }
CB_HD int xavg(const Anchor& a) { return (a.x + a.xmax) / 2; }
CB_HD int yavg(const Anchor& a) { return (a.y + a.ymax) / 2; }
CB_HD int xrange(const Anchor& a) { return iabs(a.x - a.xmax) / 2; }
CB_HD int yrange(const Anchor& a) { return iabs(a.y - a.ymax) / 2; }
// std::pow(dx, 2) + std::pow(dy, 2) as unsigned long long: exact in double for these magnitudes, so integer arithmetic
CB_HD unsigned long long asize(const Anchor& a)
{
    const long long dx = a.x - a.xmax, dy = a.y - a.ymax;
    return (unsigned long long)(dx * dx + dy * dy);
}
CB_HD int max_range(const Anchor& a) { return imax(iabs(a.x - a.xmax), iabs(a.y - a.ymax)); }
CB_HD void merge(Anchor& a, const Anchor& o)
{
    a.x = imin(a.x, o.x); a.xmax = imax(a.xmax, o.xmax); a.y = imin(a.y, o.y); a.ymax = imax(a.ymax, o.ymax);
}
CB_HD bool is_mergeable(const Anchor& a, const Anchor& rhs, int max_distance)        // Anchor.h:84-91
{
    if (iabs(xavg(a) - xavg(rhs)) > max_distance || iabs(yavg(a) - yavg(rhs)) > max_distance) return false;
    const int mr = max_range(a);
    if (mr == 0) return false;                 // (the reference divides by zero here; no anchor the scan produces has range 0)
    const int ratio = max_range(rhs) * 10 / mr;
    return ratio > 6 && ratio < 17;
}

// fixed-capacity list; `overflow` is raised instead of writing past the end (the caller reports it, nothing is dropped silently)
struct AList {
    Anchor* v; int n, cap; bool overflow;
    CB_HD void push(const Anchor& a) { if (n < cap) v[n++] = a; else overflow = true; }
};
CB_HD AList alist(Anchor* v, int cap) { AList l; l.v = v; l.n = 0; l.cap = cap; l.overflow = false; return l; }

// ---------------------------------------------------------------------------------------------- image
// the blurred gray picture and its Otsu threshold: Scanner::test_pixel on the binarised image (pixel > 127, dark) is
// blurred > threshold.  Pixels outside the image (a t4 confirm line one past an edge: out-of-bounds read in the reference)
// count as inactive.
struct Img { const uint8_t* px; int w, h, thr; };
CB_HD bool test_pixel(const Img& im, int x, int y)
{
    if (x < 0 || y < 0 || x >= im.w || y >= im.h) return false;
    return (int)im.px[(size_t)y * (size_t)im.w + (size_t)x] > im.thr;
}

// ---------------------------------------------------------------------------------------------- ScanState (ScanState.h)
// The tallies are run lengths: [1..5] = active, inactive, active, inactive, active; at the end of every third-or-later active
// run the last five runs are evaluated and the window slides on by two runs.
struct ScanState {
    int state, n; int tally[8]; bool k122;
    CB_HD void init(bool is122) { state = 0; n = 1; tally[0] = 0; k122 = is122; }
    CB_HD int evaluate() const
    {
        for (int i = 1; i <= 5; ++i) if (tally[i] == 0) return -1;
        const float center = (float)tally[3];
        for (int i = 1; i <= 5; ++i) {
            if (i == 3) continue;
            // ScanState_114: {3, 6} for all four; ScanState_122: {1, 3} outer, {0.5, 1.5} inner
            const float lo = k122 ? ((i == 1 || i == 5) ? 1.0f : 0.5f) : 3.0f;
            const float hi = k122 ? ((i == 1 || i == 5) ? 3.0f : 1.5f) : 6.0f;
            const float ratio_min = fdiv(center, (float)(tally[i] + 1));
            const float ratio_max = fdiv(center, (float)imax(1, tally[i] - 1));
            if (ratio_max < lo || ratio_min > hi) return -1;
        }
        return tally[1] + tally[2] + tally[3] + tally[4] + tally[5];
    }
    CB_HD int process(bool active)
    {
        const bool even = state == 0 || state == 2 || state == 4;
        const bool odd = state == 1 || state == 3 || state == 5;
        if ((even && active) || (odd && !active)) {
            state += 1;
            tally[n++] = 1;
            if (state == 6) {
                const int res = evaluate();
                state -= 2;
                for (int i = 0; i < n - 2; ++i) tally[i] = tally[i + 2];
                n -= 2;
                return res;
            }
            return -1;
        }
        if (odd && active) tally[n - 1] += 1;
        if (!active && (state == 2 || state == 4)) tally[n - 1] += 1;
        return -1;
    }
};

// ---------------------------------------------------------------------------------------------- line scans (Scanner.h:176-276)
CB_HD bool scan_horizontal(const Img& im, bool k122, AList& points, int y, int xstart, int xend)
{
    if (xstart < 0) xstart = 0;
    if (xend < 0 || xend > im.w) xend = im.w;
    const int init = points.n;
    ScanState st; st.init(k122);
    for (int x = xstart; x < xend; ++x) {
        const int res = st.process(test_pixel(im, x, y));
        if (res > 0) points.push(mk(x - res, x - 1, y, y));
    }
    const int res = st.process(false);
    if (res > 0) points.push(mk(xend - res, xend - 1, y, y));
    return init != points.n || points.overflow;
}

CB_HD bool scan_vertical(const Img& im, bool k122, AList& points, int x, int xmax, int ystart, int yend)
{
    if (xmax < 0) xmax = x;
    const int xa = (x + xmax) / 2;
    if (ystart < 0) ystart = 0;
    if (yend < 0 || yend > im.h) yend = im.h;
    const int init = points.n;
    ScanState st; st.init(k122);
    for (int y = ystart; y < yend; ++y) {
        const int res = st.process(test_pixel(im, xa, y));
        if (res > 0) points.push(mk(xa, xa, y - res, y - 1));
    }
    const int res = st.process(false);
    if (res > 0) points.push(mk(xa, xa, yend - res, yend - 1));
    return init != points.n || points.overflow;
}

CB_HD bool scan_diagonal(const Img& im, bool k122, AList& points, int xstart, int xend, int ystart, int yend)
{
    xend = imin(xend, im.w);
    yend = imin(yend, im.h);
    if (xstart < 0) { const int off = -xstart; xstart += off; ystart += off; }
    if (ystart < 0) { const int off = -ystart; xstart += off; ystart += off; }
    const int init = points.n;
    ScanState st; st.init(k122);
    int x = xstart, y = ystart;
    for (; x < xend && y < yend; ++x, ++y) {
        const int res = st.process(test_pixel(im, x, y));
        if (res > 0) points.push(mk(x - res, x - 1, y - res, y - 1));
    }
    const int res = st.process(false);
    if (res > 0) points.push(mk(x - res, x - 1, y - res, y - 1));
    return init != points.n || points.overflow;
}

// ---------------------------------------------------------------------------------------------- the chain of one t1 hit
constexpr int kChainT2 = 12;        // column hits per t1 point (seen: <= 3)
constexpr int kChainConfirms = 24;  // confirm hits of one t3 / t4 stage (seen: <= 6)

CB_HD bool t3_scan_diagonal(const Img& im, bool k122, int merge_cutoff, const Anchor& hint, Anchor& out, bool& overflow)
{
    Anchor buf[kChainConfirms];
    AList confirms = alist(buf, kChainConfirms);
    const int xs = xavg(hint) - 2 * yrange(hint), xe = xavg(hint) + 2 * yrange(hint);
    const int ys = hint.y - yrange(hint), ye = hint.ymax + yrange(hint);
    if (!scan_diagonal(im, k122, confirms, xs, xe, ys, ye)) return false;
    overflow |= confirms.overflow;
    bool confirm = false;
    Anchor merged = hint;
    for (int i = 0; i < confirms.n; ++i)
        if (is_mergeable(confirms.v[i], hint, merge_cutoff)) { confirm = true; merge(merged, confirms.v[i]); }
    if (confirm) out = merged;
    return confirm;
}

CB_HD bool t4_confirm_scan(const Img& im, bool k122, int merge_cutoff, Anchor hint, bool merge_confirms, Anchor& out, bool& overflow)
{
    {
        Anchor buf[kChainConfirms];
        AList confirms = alist(buf, kChainConfirms);
        const int xs = hint.x - xrange(hint), xe = hint.xmax + xrange(hint), ya = yavg(hint);
        for (int d = -1; d <= 1; ++d)
            if (!scan_horizontal(im, k122, confirms, ya + d, xs, xe)) return false;
        overflow |= confirms.overflow;
        bool confirm = false;
        for (int i = 0; i < confirms.n; ++i)
            if (is_mergeable(confirms.v[i], hint, merge_cutoff)) {
                confirm = true;
                if (!merge_confirms) break;
                merge(hint, confirms.v[i]);
            }
        if (!confirm) return false;
    }
    {
        Anchor buf[kChainConfirms];
        AList confirms = alist(buf, kChainConfirms);
        const int ys = hint.y - yrange(hint), ye = hint.ymax + yrange(hint), xa = xavg(hint);
        for (int d = -1; d <= 1; ++d)
            if (!scan_vertical(im, k122, confirms, xa + d, xa + d, ys, ye)) return false;
        overflow |= confirms.overflow;
        bool confirm = false;
        for (int i = 0; i < confirms.n; ++i)
            if (is_mergeable(confirms.v[i], hint, merge_cutoff)) {
                confirm = true;
                if (!merge_confirms) break;
                merge(hint, confirms.v[i]);
            }
        if (!confirm) return false;
    }
    out = hint;
    return true;
}

// what on_t1_scan would append for `found` if no earlier candidate shadowed it (that test is applied afterwards, in order)
CB_HD void chain(const Img& im, bool k122, int merge_cutoff, const Anchor& found, bool merge_confirms, AList& results)
{
    Anchor buf[kChainT2];
    AList p2 = alist(buf, kChainT2);
    const int ys = found.y - 3 * xrange(found), ye = found.ymax + 3 * xrange(found);
    scan_vertical(im, k122, p2, found.x, found.xmax, ys, ye);
    bool overflow = p2.overflow;
    for (int i = 0; i < p2.n; ++i) {
        Anchor p3, p4;
        if (t3_scan_diagonal(im, k122, merge_cutoff, p2.v[i], p3, overflow) && t4_confirm_scan(im, k122, merge_cutoff, p3, merge_confirms, p4, overflow))
            results.push(p4);
    }
    results.overflow |= overflow;
}

// ---------------------------------------------------------------------------------------------- std::sort(size_sort) of libstdc++
// bits/stl_algo.h: introsort (median of three, unguarded partition, depth limit 2 lg n, heap sort beyond it) followed by the
// final insertion sort; restated literally because anchors of equal size come out in the order this algorithm leaves them.
// The two halves of a partition are independent, so the recursion is replaced by an explicit stack.
CB_HD bool size_gt(const Anchor& a, const Anchor& b) { return asize(a) > asize(b); }
CB_HD void aswap(Anchor& a, Anchor& b) { const Anchor t = a; a = b; b = t; }

CB_HD void adjust_heap(Anchor* first, long hole, long len, Anchor value)
{
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (size_gt(first[second], first[second - 1])) second--;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && size_gt(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
    first[hole] = value;
}
CB_HD void heap_sort_range(Anchor* first, long len)                  // __partial_sort(first, last, last)
{
    if (len >= 2)
        for (long parent = (len - 2) / 2;; --parent) { adjust_heap(first, parent, len, first[parent]); if (parent == 0) break; }
    long last = len;
    while (last > 1) { --last; const Anchor v = first[last]; first[last] = first[0]; adjust_heap(first, 0, last, v); }
}
CB_HD void unguarded_linear_insert(Anchor* v, long last)
{
    const Anchor val = v[last];
    long next = last - 1;
    while (size_gt(val, v[next])) { v[last] = v[next]; last = next; --next; }
    v[last] = val;
}
CB_HD void insertion_sort(Anchor* v, long first, long last)
{
    if (first == last) return;
    for (long i = first + 1; i != last; ++i) {
        if (size_gt(v[i], v[first])) { const Anchor val = v[i]; for (long k = i; k > first; --k) v[k] = v[k - 1]; v[first] = val; }
        else unguarded_linear_insert(v, i);
    }
}
CB_HD void std_sort_by_size(Anchor* v, long n)
{
    if (n == 0) return;
    long lg = 0;
    while ((n >> (lg + 1)) != 0) ++lg;
    struct Range { long first, last, depth; };
    Range stack[48];
    int sp = 0;
    stack[sp].first = 0; stack[sp].last = n; stack[sp].depth = 2 * lg; ++sp;
    while (sp > 0) {
        --sp;
        long first = stack[sp].first, last = stack[sp].last, depth = stack[sp].depth;
        while (last - first > 16) {
            if (depth == 0) { heap_sort_range(v + first, last - first); break; }
            --depth;
            const long mid = first + (last - first) / 2;
            {   // __move_median_to_first(first, first + 1, mid, last - 1)
                Anchor &r = v[first], &a = v[first + 1], &b = v[mid], &c = v[last - 1];
                if (size_gt(a, b)) { if (size_gt(b, c)) aswap(r, b); else if (size_gt(a, c)) aswap(r, c); else aswap(r, a); }
                else if (size_gt(a, c)) aswap(r, a);
                else if (size_gt(b, c)) aswap(r, c);
                else aswap(r, b);
            }
            long lo = first + 1, hi = last;                      // __unguarded_partition(first + 1, last, first)
            for (;;) {
                while (size_gt(v[lo], v[first])) ++lo;
                --hi;
                while (size_gt(v[first], v[hi])) --hi;
                if (!(lo < hi)) break;
                aswap(v[lo], v[hi]);
                ++lo;
            }
            if (sp < 48) { stack[sp].first = lo; stack[sp].last = last; stack[sp].depth = depth; ++sp; }   // (cut, last)
            last = lo;
        }
    }
    if (n > 16) { insertion_sort(v, 0, 16); for (long i = 16; i != n; ++i) unguarded_linear_insert(v, i); }
    else insertion_sort(v, 0, n);
}

// ---------------------------------------------------------------------------------------------- sequential tail
// on_t1_scan's shadow test applied in t1 order: point p contributes its chain results unless an earlier candidate is mergeable with it
CB_HD void accept_in_order(const Anchor* pts, int npts, const Anchor* res, const int* nres, int res_stride, int merge_cutoff, AList& candidates)
{
    for (int p = 0; p < npts; ++p) {
        bool shadowed = false;
        for (int c = 0; c < candidates.n && !shadowed; ++c) shadowed = is_mergeable(candidates.v[c], pts[p], merge_cutoff);
        if (shadowed) continue;
        for (int r = 0; r < nres[p]; ++r) candidates.push(res[(size_t)p * res_stride + r]);
    }
}

CB_HD unsigned filter_candidates(Anchor* c, int& n)           // Scanner.cpp:83-105
{
    if (n < 3) return 0;
    std_sort_by_size(c, n);
    unsigned cutoff = 0;
    for (int i = 0; i < 3; ++i) cutoff += (unsigned)asize(c[i]);
    cutoff /= 8;
    int i = 0;
    for (; i < n; ++i) if (asize(c[i]) < cutoff) break;
    if (i > 3) i = 3;
    if (i < n) n = i;
    return cutoff;
}

CB_HD int fix3(int i) { return i < 0 ? 2 : (i >= 3 ? 0 : i); }
CB_HD bool sort_top_to_bottom(Anchor* c, int& n)              // Scanner.cpp:107-137
{
    if (n < 3) return false;
    int cx[3], cy[3];
    for (int i = 0; i < 3; ++i) { cx[i] = xavg(c[i]); cy[i] = yavg(c[i]); }
    const int ex[3] = {cx[1] - cx[2], cx[2] - cx[0], cx[0] - cx[1]}, ey[3] = {cy[1] - cy[2], cy[2] - cy[0], cy[0] - cy[1]};
    int top_left = 0, max_d = 0;
    for (int i = 0; i < 3; ++i) { const int d = ex[i] * ex[i] + ey[i] * ey[i]; if (d > max_d) { top_left = i; max_d = d; } }
    const int dep = fix3(top_left - 1), inc = fix3(top_left + 1);
    const int dx = ex[dep], dy = ey[dep], ix = -ey[inc], iy = ex[inc];
    const int ox = dx - ix, oy = dy - iy;
    int top_right, bottom_left;
    if (ox * ox + oy * oy < dx * dx + dy * dy) { top_right = fix3(top_left + 1); bottom_left = fix3(top_left - 1); }
    else { top_right = fix3(top_left - 1); bottom_left = fix3(top_left + 1); }
    const Anchor a = c[top_left], b = c[top_right], d = c[bottom_left];
    c[0] = a; c[1] = b; c[2] = d; n = 3;
    return true;
}

// the window add_bottom_right_corner searches (Scanner.cpp:139-165): t1 rows ystart, ystart + skip, ... < yend over [xstart, xend)
struct Window { int skip, y, yend, xstart, xend; };
CB_HD Window bottom_right_window(const Anchor* A, int scanner_skip, int img_h)
{
    const double topScalar = ddiv((double)max_range(A[2]), (double)imax(max_range(A[1]), max_range(A[0])));
    const int tex = (int)dmul((double)(xavg(A[1]) - xavg(A[0])), topScalar), tey = (int)dmul((double)(yavg(A[1]) - yavg(A[0])), topScalar);
    const int g1x = xavg(A[2]) + tex, g1y = yavg(A[2]) + tey;
    const double leftScalar = ddiv((double)max_range(A[1]), (double)imax(max_range(A[2]), max_range(A[0])));
    const int lex = (int)dmul((double)(xavg(A[2]) - xavg(A[0])), leftScalar), ley = (int)dmul((double)(yavg(A[2]) - yavg(A[0])), leftScalar);
    const int g2x = xavg(A[1]) + lex, g2y = yavg(A[1]) + ley;
    const int cx = (g1x + g2x) / 2, cy = (g1y + g2y) / 2;
    const int range = (int)((float)imax(imax(max_range(A[0]), max_range(A[1])), max_range(A[2])) * 2.0f);
    Window w;
    // t1_scan_rows' argument defaults (Scanner.h:279-286): skip <= 0 -> the scanner's, y < 0 -> skip, yend clipped to the image
    w.skip = scanner_skip / 2; if (w.skip <= 0) w.skip = scanner_skip;
    w.y = cy - range; if (w.y < 0) w.y = w.skip;
    w.yend = cy + range; if (w.yend < 0 || w.yend > img_h) w.yend = img_h;
    w.xstart = cx - range; w.xend = cx + range;
    return w;
}

}  // namespace scan
}  // namespace cb200

// ---------------------------------------------------------------------------------------------- one picture, one thread group
// Scanner::scan() (Scanner.cpp:182-199) for one picture by a group of cooperating threads (a CTA in scan.cu; a single thread
// in the host test).  Parallel phases hand out rows / t1 points round-robin; the order-dependent parts run on thread 0.
namespace cb200 {
namespace scan {

constexpr int kRowCap = 64;        // t1 hits per scanned row (seen: <= 8 for ScanState_114, <= 32 inside a ScanState_122 window)
constexpr int kPtsCap = 2048;      // t1 hits per picture and pass (seen: <= 160)
constexpr int kResCap = 4;         // confirmed anchors per t1 hit (seen: <= 1)
constexpr int kCandCap = 64;       // accepted candidates per pass (seen: <= 8)

enum : int { kScanOk = 0, kScanOverflow = 1 };   // status bit: a fixed-capacity list overflowed -> the result is not trusted

struct PicWs {                     // per-picture scratch in global memory
    Anchor* rowbuf;                // [rows_cap][kRowCap]
    int* rowcnt;                   // [rows_cap]
    Anchor* pts;                   // [kPtsCap]
    Anchor* res;                   // [kPtsCap][kResCap]
    int* nres;                     // [kPtsCap]
    int rows_cap;
};
struct PicShared {                 // state every thread of the group reads (shared memory on the device)
    Window win; int nrows, npts, ncand, status; unsigned cutoff; int want_fourth;
    Anchor cand[kCandCap];
};
struct Exec {
    int tid, nthreads;
    CB_HD void barrier() const
    {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
};

// t1 over the window's rows -> on_t1_scan chains -> accepted candidates in sh.cand / sh.ncand
CB_HD void scan_pass(const Exec& ex, const Img& im, const PicWs& ws, PicShared& sh, bool k122, int merge_cutoff, bool merge_confirms)
{
    if (ex.tid == 0) {
        const Window& w = sh.win;
        int nrows = (w.yend > w.y) ? (w.yend - w.y + w.skip - 1) / w.skip : 0;
        if (nrows > ws.rows_cap) { nrows = ws.rows_cap; sh.status |= kScanOverflow; }
        sh.nrows = nrows;
    }
    ex.barrier();
    for (int r = ex.tid; r < sh.nrows; r += ex.nthreads) {                 // t1_scan_rows: one row per thread
        AList row = alist(ws.rowbuf + (size_t)r * kRowCap, kRowCap);
        scan_horizontal(im, k122, row, sh.win.y + r * sh.win.skip, sh.win.xstart, sh.win.xend);
        ws.rowcnt[r] = row.overflow ? -row.n : row.n;
    }
    ex.barrier();
    if (ex.tid == 0) {                                                     // row lists -> one list in t1 order
        int np = 0;
        for (int r = 0; r < sh.nrows; ++r) {
            int c = ws.rowcnt[r];
            if (c < 0) { c = -c; sh.status |= kScanOverflow; }
            for (int i = 0; i < c; ++i) {
                if (np < kPtsCap) ws.pts[np++] = ws.rowbuf[(size_t)r * kRowCap + i]; else sh.status |= kScanOverflow;
            }
        }
        sh.npts = np;
    }
    ex.barrier();
    for (int p = ex.tid; p < sh.npts; p += ex.nthreads) {                  // t2 -> t3 -> t4 of every hit, independent of each other
        AList results = alist(ws.res + (size_t)p * kResCap, kResCap);
        chain(im, k122, merge_cutoff, ws.pts[p], merge_confirms, results);
        ws.nres[p] = results.overflow ? -1 - results.n : results.n;
    }
    ex.barrier();
    if (ex.tid == 0) {
        for (int p = 0; p < sh.npts; ++p) if (ws.nres[p] < 0) { ws.nres[p] = -1 - ws.nres[p]; sh.status |= kScanOverflow; }
        AList cand = alist(sh.cand, kCandCap);
        accept_in_order(ws.pts, sh.npts, ws.res, ws.nres, kResCap, merge_cutoff, cand);
        if (cand.overflow) sh.status |= kScanOverflow;
        sh.ncand = cand.n;
    }
    ex.barrier();
}

// out: up to four anchors in the reference's order (top-left, top-right, bottom-left, bottom-right); returns their number
CB_HD int scan_picture(const Exec& ex, const Img& im, const PicWs& ws, PicShared& sh, Anchor* out, unsigned* cutoff_out, int* status_out)
{
    const int scanner_skip = imin(im.h, im.w) / 60, merge_cutoff = im.w / 30;       // Scanner.h:168-174
    if (ex.tid == 0) {
        sh.status = kScanOk;
        // scan_primary: t1_scan_rows<ScanState_114> with its defaults: rows skip, 2 skip, ... over the full width
        sh.win.skip = scanner_skip; sh.win.y = scanner_skip; sh.win.yend = im.h; sh.win.xstart = -1; sh.win.xend = -1;
    }
    ex.barrier();
    scan_pass(ex, im, ws, sh, false, merge_cutoff, true);
    if (ex.tid == 0) {
        int n = sh.ncand;
        sh.cutoff = filter_candidates(sh.cand, n);
        sort_top_to_bottom(sh.cand, n);
        sh.ncand = n;
        sh.want_fourth = (n == 3 && sh.cutoff != 0) ? 1 : 0;
        for (int i = 0; i < n && i < 4; ++i) out[i] = sh.cand[i];
        if (sh.want_fourth) sh.win = bottom_right_window(sh.cand, scanner_skip, im.h);
    }
    ex.barrier();
    int count = sh.ncand < 4 ? sh.ncand : 4;
    if (sh.want_fourth) {                                                  // add_bottom_right_corner, Scanner.cpp:139-180
        const unsigned cutoff = sh.cutoff;
        scan_pass(ex, im, ws, sh, true, merge_cutoff, false);
        if (ex.tid == 0) {
            sh.want_fourth = 0;
            for (int c = 0; c < sh.ncand; ++c)
                if (asize(sh.cand[c]) > cutoff) { out[3] = sh.cand[c]; sh.want_fourth = 1; break; }
        }
        ex.barrier();
        count = 3 + sh.want_fourth;
    }
    if (ex.tid == 0) { *cutoff_out = sh.cutoff; *status_out = sh.status; }
    return count;
}

}  // namespace scan
}  // namespace cb200
