"""The cimbard_* receive facade (include/cimbard_b200.h, libcimbar_b200/host/cimbard_b200.cpp): the reference's C entry points
src/lib/cimbar_js/cimbar_recv_js.h:11-39 over libcb200.  CPU part: symbols, argument errors (the reference's return values,
cimbar_recv_js.cpp:152-165, :192-200) and the host-only fountain leg against wirehair; GPU part: frames in, file out."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def facade():
    from libcimbar_b200 import build
    build.build()
    build.build_wirehair()
    L = C.CDLL(build.build_facade())
    L.cimbard_get_report.restype = C.c_uint
    L.cimbard_get_report.argtypes = [C.c_void_p, C.c_uint]
    L.cimbard_scan_extract_decode.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_uint]
    L.cimbard_b200_extract_decode.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    L.cimbard_fountain_decode.restype = C.c_int64
    L.cimbard_fountain_decode.argtypes = [C.c_void_p, C.c_uint]
    L.cimbard_get_filesize.restype = C.c_uint
    L.cimbard_get_filesize.argtypes = [C.c_uint32]
    L.cimbard_b200_file_read.restype = C.c_int64
    L.cimbard_b200_file_read.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64]
    L.cimbard_b200_reset.restype = None
    return L


def test_facade_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cimbard_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(cimbard_\w+)\s*\(", hdr)))
    assert len(declared) == 9
    L = facade()
    for name in declared:
        assert hasattr(L, name), name


def test_facade_sizes_and_argument_errors():
    L = facade()
    L.cimbard_b200_reset()
    assert L.cimbard_configure_decode(0) == 0
    assert L.cimbard_get_bufsize() == 12 * 625                       # mode B
    assert L.cimbard_configure_decode(4) == 0
    assert L.cimbard_get_bufsize() == 10 * 750                       # 4C: Config.h:26-28 fountain_chunks_scalar -10
    assert L.cimbard_configure_decode(12345) == 0                    # unknown -> B (Config.h:41-43)
    assert L.cimbard_get_bufsize() == 7500
    img = np.zeros((64, 64, 3), np.uint8)
    buf = np.zeros(7500, np.uint8)
    assert L.cimbard_scan_extract_decode(img.ctypes.data, 0, 64, 3, buf.ctypes.data, 7500) == -1
    assert L.cimbard_scan_extract_decode(img.ctypes.data, 64, 64, 3, buf.ctypes.data, 7499) == -2
    # a picture without anchors: -3 (Extractor::FAILURE) from the GPU scan; -6 where there is no GPU (the library has no CPU path)
    assert L.cimbard_scan_extract_decode(img.ctypes.data, 64, 64, 3, buf.ctypes.data, 7500) in (-3, -6)
    assert L.cimbard_b200_extract_decode(img.ctypes.data, 64, 64, 3, None, buf.ctypes.data, 7500) == -3   # not an extracted frame
    big = np.zeros((1024, 1024, 3), np.uint8)
    assert L.cimbard_scan_extract_decode(big.ctypes.data, 1024, 1024, 12, buf.ctypes.data, 7500) == -4   # NV12 is not restated
    assert L.cimbard_fountain_decode(buf.ctypes.data, 0) == -5
    assert L.cimbard_fountain_decode(buf.ctypes.data, 626) == -5
    # FountainMetadata id -> size (FountainMetadata.h:29-45), including the 25th size bit in the top bit of byte 0
    for size in (1, 23586, (1 << 24) + 77):
        d = bytes([(0x55 & 0x7F) | ((size >> 17) & 0x80), (size >> 16) & 0xFF, (size >> 8) & 0xFF, size & 0xFF])
        assert L.cimbard_get_filesize(int.from_bytes(d, "little")) == size


def test_facade_fountain_leg_recovers_the_file_on_the_host():
    import libcimbar_b200 as cb
    from libcimbar_b200.fountain_bench import make_stream
    L = facade()
    L.cimbard_b200_reset()
    L.cimbard_configure_decode(68)
    info = cb.mode_info(68)
    data, chunks = make_stream(info, 40000, 8, seed=5)                # 96 chunks of 619 payload bytes for a 40,000-byte file
    res, fed = 0, 0
    for f in range(chunks.shape[0]):
        if f == 1:
            continue                                                   # a lost frame
        frame = np.ascontiguousarray(chunks[f])
        res = L.cimbard_fountain_decode(frame.ctypes.data, frame.size)
        fed += 1
        if res != 0:
            break
    assert res > 0 and fed < chunks.shape[0]
    assert L.cimbard_get_filesize(res & 0xFFFFFFFF) == 40000
    out = np.zeros(40000, np.uint8)
    assert L.cimbard_b200_file_read(res & 0xFFFFFFFF, out.ctypes.data, 39999) == -2
    assert L.cimbard_b200_file_read(res & 0xFFFFFFFF, out.ctypes.data, 40000) == 40000
    assert hashlib.sha256(out.tobytes()).digest() == hashlib.sha256(data.tobytes()).digest()
    # more chunks of a finished file: -1 (fountain_decoder_sink.h:142-143)
    last = np.ascontiguousarray(chunks[-1])
    assert L.cimbard_fountain_decode(last.ctypes.data, last.size) == -1
    L.cimbard_b200_reset()


@pytest.mark.gpu
def test_facade_frames_in_file_out():
    """cimbar_recv_js's loop: every camera frame -> cimbard_scan_extract_decode -> cimbard_fountain_decode, until an id comes back"""
    import torch
    import libcimbar_b200 as cb
    from libcimbar_b200.fountain_bench import make_stream
    L = facade()
    L.cimbard_b200_reset()
    L.cimbard_configure_decode(68)
    info = cb.mode_info(68)
    n = 10
    data, chunks = make_stream(info, 50000, n, seed=9)
    ctx = cb.Context(68, max_frames=n)
    d_payload = torch.from_numpy(chunks.reshape(n, -1)).cuda()
    d_cells = torch.empty((n, info.total_cells), dtype=torch.uint8, device="cuda")
    d_rgb = torch.empty((n, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
    ctx.encode_cells_dev(d_payload.data_ptr(), n, d_cells.data_ptr())
    ctx.render_frames_dev(d_cells.data_ptr(), n, d_rgb.data_ptr())
    ctx.sync()
    frames = d_rgb.cpu().numpy()
    del ctx
    # the synthetic renderer draws no anchors; the reference scans every image it is given (cimbar_recv_js.cpp:171-177), so the
    # four anchor squares of a real encoder frame are pasted in (they lie outside every cell's threshold support)
    import oracle_lib as ol
    tr, c = ol.load_sample("b/tr_0.png"), 62
    for ys in (slice(0, c), slice(1024 - c, 1024)):
        for xs in (slice(0, c), slice(1024 - c, 1024)):
            frames[:, ys, xs] = tr[ys, xs]
    blank = np.zeros((480, 640, 3), np.uint8)
    assert L.cimbard_scan_extract_decode(blank.ctypes.data, 640, 480, 3, np.zeros(7500, np.uint8).ctypes.data, 7500) == -3
    buf = np.zeros(L.cimbard_get_bufsize(), np.uint8)
    res = 0
    for f in range(n):
        if f == 2:
            continue
        img = np.ascontiguousarray(frames[f])
        if f == 3:                                                              # RGBA input (format 4)
            img = np.ascontiguousarray(np.concatenate([img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2))
            got = L.cimbard_scan_extract_decode(img.ctypes.data, 1024, 1024, 4, buf.ctypes.data, buf.size)
        else:
            got = L.cimbard_scan_extract_decode(img.ctypes.data, 1024, 1024, 3, buf.ctypes.data, buf.size)
        assert got == buf.size, (f, got)
        assert np.array_equal(buf.reshape(12, 625), chunks[f])
        res = L.cimbard_fountain_decode(buf.ctypes.data, got)
        assert res >= 0
        if res > 0:
            break
    assert res > 0
    out = np.zeros(50000, np.uint8)
    assert L.cimbard_b200_file_read(res & 0xFFFFFFFF, out.ctypes.data, out.size) == 50000
    assert np.array_equal(out, data)
    # an identity "camera": the anchor centres where Deskewer puts them (Deskewer.h:27-39) -> the same chunks through the
    # deskew path
    c = np.array([30, 30, 994, 30, 30, 994, 994, 994], np.float32)
    img = np.ascontiguousarray(frames[0])
    got = L.cimbard_b200_extract_decode(img.ctypes.data, 1024, 1024, 3, c.ctypes.data, buf.ctypes.data, buf.size)
    assert got == buf.size
    assert np.array_equal(buf.reshape(12, 625), chunks[0])
    rep = C.create_string_buffer(128)
    assert L.cimbard_get_report(rep, 128) > 0
    L.cimbard_b200_reset()
