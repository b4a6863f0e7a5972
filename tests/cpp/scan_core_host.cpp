// scan_core_host.cpp -- TEST HARNESS: compiles the device anchor-scan code (libcimbar_b200/csrc/scan_core.cuh, the functions the
// kernels of scan.cu are made of) for the host and runs Scanner::scan() with a one-thread "CTA", so that its logic can be
// compared with the oracle where there is no GPU.  Not part of the product; the product runs the same functions in CUDA kernels.
#include "../../libcimbar_b200/csrc/scan_core.cuh"

#include <cstdlib>
#include <vector>

using namespace cb200::scan;

extern "C" int sc_scan(const uint8_t* blurred, int w, int h, int thr, int* anchors16, unsigned* cutoff, int* status)
{
    Img im{blurred, w, h, thr};
    const int skip = (h < w ? h : w) / 60;
    const int rows_cap = 2 * ((h + skip - 1) / skip) + 4;
    std::vector<Anchor> rowbuf((size_t)rows_cap * kRowCap), pts(kPtsCap), res((size_t)kPtsCap * kResCap);
    std::vector<int> rowcnt(rows_cap), nres(kPtsCap);
    PicWs ws{rowbuf.data(), rowcnt.data(), pts.data(), res.data(), nres.data(), rows_cap};
    PicShared sh;
    Exec ex{0, 1};
    Anchor out[4] = {};
    const int n = scan_picture(ex, im, ws, sh, out, cutoff, status);
    for (int i = 0; i < 4; ++i) { anchors16[4 * i] = out[i].x; anchors16[4 * i + 1] = out[i].xmax; anchors16[4 * i + 2] = out[i].y; anchors16[4 * i + 3] = out[i].ymax; }
    return n;
}

// std::sort(size_sort()) restated for the device against the real std::sort of this toolchain's libstdc++ (the one the
// reference links): sizes with many ties, lengths around the 16-element insertion-sort threshold and beyond
#include <algorithm>
extern "C" int sc_sort_check(const int* xywh, int n, int* out_mine, int* out_std)
{
    std::vector<Anchor> a(n), b(n);
    for (int i = 0; i < n; ++i) a[i] = b[i] = mk(xywh[4 * i], xywh[4 * i + 1], xywh[4 * i + 2], xywh[4 * i + 3]);
    std_sort_by_size(a.data(), n);
    std::sort(b.begin(), b.end(), [](const Anchor& p, const Anchor& q) { return asize(p) > asize(q); });
    int same = 1;
    for (int i = 0; i < n; ++i) {
        const Anchor *p = &a[i], *q = &b[i];
        out_mine[4 * i] = p->x; out_mine[4 * i + 1] = p->xmax; out_mine[4 * i + 2] = p->y; out_mine[4 * i + 3] = p->ymax;
        out_std[4 * i] = q->x; out_std[4 * i + 1] = q->xmax; out_std[4 * i + 2] = q->y; out_std[4 * i + 3] = q->ymax;
        if (p->x != q->x || p->xmax != q->xmax || p->y != q->y || p->ymax != q->ymax) same = 0;
    }
    return same;
}
