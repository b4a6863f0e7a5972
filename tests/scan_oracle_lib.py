"""ctypes binding of oracle/scan_oracle.c (the anchor-scan restatement).  Test infrastructure only."""
import ctypes as C

import numpy as np

import oracle_lib as ol


class Anchor(C.Structure):
    _fields_ = [("x", C.c_int), ("xmax", C.c_int), ("y", C.c_int), ("ymax", C.c_int)]

    def tup(self):
        return (self.x, self.xmax, self.y, self.ymax)

    def __str__(self):                                   # Anchor.h:107-111
        return "%d+-%d,%d+-%d" % ((self.x + self.xmax) // 2, abs(self.x - self.xmax) // 2,
                                  (self.y + self.ymax) // 2, abs(self.y - self.ymax) // 2)


def anchor_str(t):
    return str(Anchor(*t))


def join(anchors):
    return " ".join(anchor_str(a) for a in anchors)


class Scanner(C.Structure):
    _fields_ = [("bin", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("skip", C.c_int), ("merge_cutoff", C.c_int), ("anchor_size", C.c_int)]


class ScanOracle:
    def __init__(self):
        self.lib = L = ol.Oracle().lib
        u8p = C.POINTER(C.c_uint8)
        ap = C.POINTER(Anchor)
        L.cbo_scan_preprocess.argtypes = [u8p, C.c_int, C.c_int, u8p, u8p]
        L.cbo_scan_gaussian_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.cbo_scan_otsu.argtypes = [u8p, C.c_size_t]
        L.cbo_scanner_init.argtypes = [C.POINTER(Scanner), u8p, C.c_int, C.c_int, C.c_int]
        L.cbo_scan_t1.argtypes = [C.POINTER(Scanner), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ap, C.c_int]
        L.cbo_scan_t2.argtypes = [C.POINTER(Scanner), C.c_int, ap, ap, C.c_int]
        L.cbo_scan_t3.argtypes = [C.POINTER(Scanner), C.c_int, ap, ap]
        L.cbo_scan_t4.argtypes = [C.POINTER(Scanner), C.c_int, ap, C.c_int, ap]
        L.cbo_scan_deduplicate.argtypes = [C.POINTER(Scanner), ap, C.c_int, ap]
        L.cbo_scan_filter.argtypes = [ap, C.c_int, C.POINTER(C.c_uint)]
        L.cbo_scan_sort_top_to_bottom.argtypes = [ap, C.c_int]
        L.cbo_scan_primary.argtypes = [C.POINTER(Scanner), ap, C.c_int, C.POINTER(C.c_uint)]
        L.cbo_scan_bottom_right.argtypes = [C.POINTER(Scanner), ap, C.c_uint]
        L.cbo_scan_bin.argtypes = [u8p, C.c_int, C.c_int, ap, C.POINTER(C.c_uint)]
        L.cbo_scan.argtypes = [u8p, C.c_int, C.c_int, ap, C.POINTER(C.c_uint)]
        L.cbo_scan_corners.argtypes = [ap, C.POINTER(C.c_int)]
        L.cbo_scan_is_granular_scale.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]

    def preprocess(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        bin_ = np.zeros((h, w), np.uint8); blurred = np.zeros((h, w), np.uint8)
        t = self.lib.cbo_scan_preprocess(ol._ptr(rgb), w, h, ol._ptr(bin_), ol._ptr(blurred))
        return t, bin_, blurred

    def scanner(self, bin_):
        s = Scanner()
        self._keep = np.ascontiguousarray(bin_)
        self.lib.cbo_scanner_init(C.byref(s), ol._ptr(self._keep), bin_.shape[1], bin_.shape[0], 0)
        return s

    def _list(self, fn, *args, cap=4096):
        out = (Anchor * cap)()
        n = fn(*args, out, cap)
        assert n <= cap
        return [out[i].tup() for i in range(n)]

    def t1(self, s, kind=114, skip=-1, y=-1, yend=-1, xstart=-1, xend=-1):
        return self._list(self.lib.cbo_scan_t1, C.byref(s), kind, skip, y, yend, xstart, xend)

    def t2(self, s, hint, kind=114):
        return self._list(self.lib.cbo_scan_t2, C.byref(s), kind, C.byref(Anchor(*hint)))

    def t3(self, s, hint, kind=114):
        o = Anchor()
        return [o.tup()] if self.lib.cbo_scan_t3(C.byref(s), kind, C.byref(Anchor(*hint)), C.byref(o)) else []

    def t4(self, s, hint, merge_confirms=True, kind=114):
        o = Anchor()
        return [o.tup()] if self.lib.cbo_scan_t4(C.byref(s), kind, C.byref(Anchor(*hint)), int(merge_confirms), C.byref(o)) else []

    def deduplicate(self, s, anchors):
        n = len(anchors)
        a = (Anchor * max(n, 1))(*[Anchor(*t) for t in anchors]); o = (Anchor * max(n, 1))()
        m = self.lib.cbo_scan_deduplicate(C.byref(s), a, n, o)
        return [o[i].tup() for i in range(m)]

    def filter(self, anchors):
        n = len(anchors)
        a = (Anchor * max(n, 1))(*[Anchor(*t) for t in anchors])
        cutoff = C.c_uint(0)
        m = self.lib.cbo_scan_filter(a, n, C.byref(cutoff))
        return [a[i].tup() for i in range(m)], cutoff.value

    def sort_top_to_bottom(self, anchors):
        n = len(anchors)
        a = (Anchor * max(n, 1))(*[Anchor(*t) for t in anchors])
        m = self.lib.cbo_scan_sort_top_to_bottom(a, n)
        return None if m < 0 else [a[i].tup() for i in range(m)]

    def primary(self, s):
        cutoff = C.c_uint(0)
        out = (Anchor * 64)()
        n = self.lib.cbo_scan_primary(C.byref(s), out, 64, C.byref(cutoff))
        return [out[i].tup() for i in range(min(n, 64))], cutoff.value

    def bottom_right(self, s, anchors3, cutoff):
        a = (Anchor * 4)(*[Anchor(*t) for t in anchors3])
        ok = self.lib.cbo_scan_bottom_right(C.byref(s), a, cutoff)
        return [a[i].tup() for i in range(4 if ok else 3)], bool(ok)

    def scan(self, rgb):
        """Scanner(img).scan(): (anchors, cutoff); anchors is None when the image size is not supported"""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        out = (Anchor * 4)(); cutoff = C.c_uint(0)
        n = self.lib.cbo_scan(ol._ptr(rgb), rgb.shape[1], rgb.shape[0], out, C.byref(cutoff))
        return (None if n < 0 else [out[i].tup() for i in range(n)]), cutoff.value

    def corners(self, anchors4):
        a = (Anchor * 4)(*[Anchor(*t) for t in anchors4]); xy = (C.c_int * 8)()
        self.lib.cbo_scan_corners(a, xy)
        return [xy[i] for i in range(8)]

    def is_granular_scale(self, xy8, w, h):
        return bool(self.lib.cbo_scan_is_granular_scale((C.c_int * 8)(*xy8), w, h))
