"""Pins the CPU oracle (oracle/cimbar_oracle.c) to the reference's own golden vectors.
Sources: src/lib/encoder/test/DecoderTest.cpp:26-106 (SHA-256 of decoded bytes),
cimb_translator/test/CimbReaderTest.cpp:37-163 (first 22 cells in flood order),
cimb_translator/test/CimbDecoderTest.cpp:77-131 (colour known answers)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle_lib import Oracle, load_sample, manifest

ORA = Oracle()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_sample_pixels_match_manifest():
    # the decoded pixels (esp. of the JPEGs) must be the ones the goldens were pinned on
    for name, ent in manifest()["samples"].items():
        if "rgb_sha256" in ent:
            assert sha(load_sample(name)) == ent["rgb_sha256"], name


@pytest.mark.parametrize("g", manifest()["goldens"], ids=lambda g: f"{g['sample']}-m{g['mode']}-ecc{int(g['ecc'])}")
def test_decoder_sha256_goldens(g):
    m = ORA.mode(g["mode"])
    rgb = load_sample(g["sample"])
    if g["ecc"]:
        out, ok = ORA.decode(m, rgb, use_ecc=True)
    else:
        out = ORA.decode_raw(m, rgb)
    assert out.size == g["bytes"]
    assert sha(out) == g["sha256"], g["source"]


def _first22(sample, color_mode_override=None):
    m = ORA.mode(68)
    if color_mode_override is not None:
        m.color_mode = color_mode_override
    raw, cells = ORA.decode_raw(m, load_sample(sample), want_cells=True)
    idx = np.argsort(cells["order"])[:22]
    res = {int(i): int(cells["symbol"][i]) | (int(cells["color"][i]) << 4) for i in idx}
    return " ".join(f"{k}={v}" for k, v in sorted(res.items())), cells


def test_reader_first_read():
    # CimbReaderTest.cpp:37-58: first read is cell 0 at (62,8), bits 0, colour 1
    s, cells = _first22("6bit/4color_ecc30_fountain_0.png", 1)
    first = int(np.argmin(cells["order"]))
    assert first == 0 and cells["x"][0] == 62 and cells["y"][0] == 8
    assert cells["symbol"][0] == 0 and cells["color"][0] == 1


def test_reader_sample_colormode0():
    # CimbReaderTest.cpp:60-94
    s, _ = _first22("6bit/4color_ecc30_fountain_0.png", 0)
    assert s == ("0=0 99=8 11680=3 11681=32 11900=28 11901=25 11904=12 11995=2 11996=8 11998=6 "
                 "11999=54 12001=29 12004=6 12099=2 12195=57 12196=1 12200=5 12201=0 12298=32 "
                 "12299=34 12300=30 12399=15")


def test_reader_sample_colormode1():
    # CimbReaderTest.cpp:96-131
    s, _ = _first22("6bit/4color_ecc30_fountain_0.png", 1)
    assert s == ("0=16 99=24 11680=19 11681=48 11900=44 11901=41 11904=28 11995=18 11996=24 "
                 "11998=22 11999=6 12001=45 12004=22 12099=18 12195=9 12196=17 12200=21 12201=16 "
                 "12298=48 12299=50 12300=46 12399=31")


def test_reader_sample_messy():
    # CimbReaderTest.cpp:133-163 (camera JPEG: drift, cooldown, heap tie-breaks)
    s, _ = _first22("6bit/4_30_f0_627_extract.jpg", 1)
    assert s == ("0=16 1=44 99=24 100=44 600=49 601=54 711=46 712=9 11464=5 11576=48 11577=60 "
                 "11687=57 11688=7 11689=48 11690=0 11798=31 11799=41 12297=62 12298=48 12299=50 "
                 "12300=46 12399=31")


COLOR_CASES_MODE0 = [((255, 0, 255), 2), ((255, 255, 0), 1), ((0, 255, 255), 0), ((0, 255, 0), 3), ((0, 0, 0), 0),
                     ((70, 70, 70), 0), ((20, 200, 20), 3), ((50, 155, 50), 3), ((200, 30, 200), 2), ((155, 50, 155), 2),
                     ((200, 155, 20), 1), ((155, 155, 50), 1), ((50, 155, 200), 0), ((50, 155, 155), 0)]
COLOR_CASES_MODE1 = [((255, 0, 255), 3), ((255, 255, 0), 2), ((0, 255, 255), 1), ((0, 255, 0), 0), ((0, 0, 0), 0),
                     ((70, 70, 70), 0), ((20, 200, 20), 0), ((50, 155, 50), 0), ((200, 30, 200), 3), ((155, 50, 155), 3),
                     ((200, 155, 20), 2), ((155, 155, 50), 2), ((50, 155, 200), 1), ((50, 155, 155), 1)]


def test_best_color_known_answers():
    # CimbDecoderTest.cpp:77-131
    for mode, cases in ((0, COLOR_CASES_MODE0), (1, COLOR_CASES_MODE1)):
        for (r, g, b), want in cases:
            assert ORA.lib.cbo_best_color(r, g, b, 4, mode, None) == want, (mode, r, g, b)


def test_all_tile_colour_combos_roundtrip():
    # CimbDecoderTest.cpp:146-165 (all 64 tile x colour combos) via render -> decode of a whole frame
    m = ORA.mode(68)
    cells = (np.arange(m.total_cells) % 64).astype(np.uint8)
    rgb = ORA.render_frame(m, cells)
    raw, info = ORA.decode_raw(m, rgb, want_cells=True)
    got = info["symbol"].astype(np.uint8) | (info["color"].astype(np.uint8) << 4)
    assert np.array_equal(got, cells)
    assert np.all(info["drift_offset"] == 4) and np.all(info["distance"] == 0)


def test_preprocess_matches_cv2():
    # OpenCV is a third-party dependency of the reference (CimbReader.cpp:35,41,26); pin the restated
    # arithmetic to cv2 itself on camera frames, and to the SHA recorded when the goldens were made
    import cv2
    for name, pins in manifest()["cv_pins"].items():
        rgb = load_sample(name)
        h, w = rgb.shape[:2]
        gray = cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY)
        assert sha(gray) == pins["gray"]
        mine = np.zeros((h, w), np.uint8)
        ORA.lib.cbo_rgb_to_gray(rgb.ctypes.data_as(ORA.lib.cbo_rgb_to_gray.argtypes[0]), w, h, mine.ctypes.data_as(ORA.lib.cbo_rgb_to_gray.argtypes[3]))
        assert np.array_equal(mine, gray)
        thr = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, 5, 0)
        assert sha(thr) == pins["thr5"]
        assert np.array_equal(np.unpackbits(ORA.preprocess(rgb)).reshape(h, w) * 255, thr)
        k = np.array([[-0, -1, -0], [-1, 4.5, -1], [-0, -1, -0]], dtype=np.float32)
        sharp = cv2.filter2D(gray, -1, k)
        thr7 = cv2.adaptiveThreshold(sharp, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, 7, 0)
        assert sha(thr7) == pins["thr7_sharp"]
        assert np.array_equal(np.unpackbits(ORA.preprocess(rgb, sharpen=True)).reshape(h, w) * 255, thr7)


def test_fused_preprocess_equals_pass_by_pass_restatement():
    # cbo_preprocess is the fused, vectorisable form the timing legs run; cbo_preprocess_unfused is cvtColor -> [filter2D] ->
    # adaptiveThreshold -> mat_to_bitbuffer pass by pass.  Same bits on noise, synthetic frames and every mode's geometry.
    import ctypes as C
    from oracle_lib import _ptr, u8p
    ORA.lib.cbo_preprocess_unfused.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
    rng = np.random.default_rng(12)
    m = ORA.mode(68)
    cases = [rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8), rng.integers(0, 256, (637, 736, 3), dtype=np.uint8),
             rng.integers(0, 256, (720, 1024, 3), dtype=np.uint8), np.full((64, 64, 3), 255, np.uint8),
             ORA.render_frame(m, rng.integers(0, 64, m.total_cells, dtype=np.uint8))]
    for rgb in cases:
        h, w = rgb.shape[:2]
        for sharpen in (False, True):
            want = np.zeros(w * h // 8 + 16, np.uint8)
            ORA.lib.cbo_preprocess_unfused(_ptr(rgb), w, h, int(sharpen), _ptr(want))
            assert np.array_equal(ORA.preprocess(rgb, sharpen), want[: w * h // 8]), (rgb.shape, sharpen)


# ---------------------------------------------------------------------------------------------- colour correction
def _fmt_matx(mat):
    """operator<<(std::ostream&, cv::Matx<float,3,3>): "%.8g" elements, ", " / ";\n " separators"""
    return "[" + ";\n ".join(", ".join("%.8g" % float(x) for x in row) for row in mat) + "]"


def test_adaptation_matrix_matches_reference_string():
    g = manifest()["adaptation_golden"]          # color_correctionTest.cpp:14-30
    mat = ORA.adaptation_matrix(g["actual"], g["desired"])
    assert _fmt_matx(mat) == g["matrix_str"]
    r, gg, b = (np.float32(x) for x in g["transform_in"])
    out = [np.float32(np.float32(np.float32(mat[i, 0] * r) + np.float32(mat[i, 1] * gg)) + np.float32(mat[i, 2] * b)) for i in range(3)]
    assert np.allclose(out, g["transform_out"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("idx", range(3))
def test_ccm_goldens_first_colors(idx):
    g = manifest()["ccm_goldens"][idx]           # CimbReaderTest.cpp:181-272: colours of the first six reads with that CCM
    m = ORA.mode(g["mode"])
    rgb = load_sample(g["sample"])
    ORA.set_ccm(g["matrix"])
    try:
        _, cells = ORA.decode_raw(m, rgb, want_cells=True)
    finally:
        ORA.set_ccm(None)
    order = np.argsort(cells["order"])
    assert cells["color"][order[:6]].tolist() == g["first_colors"]


def test_ccm_is_very_necessary_for_ex380():
    """the reference names the ex380 case "VeryNecessary": without the CCM a visible share of the colours changes"""
    g = manifest()["ccm_goldens"][2]
    m = ORA.mode(68)
    rgb = load_sample(g["sample"])
    _, plain = ORA.decode_raw(m, rgb, want_cells=True)
    ORA.set_ccm(g["matrix"])
    try:
        _, fixed = ORA.decode_raw(m, rgb, want_cells=True)
    finally:
        ORA.set_ccm(None)
    assert np.array_equal(plain["symbol"], fixed["symbol"])
    assert (plain["color"] != fixed["color"]).mean() > 0.02


@pytest.mark.parametrize("idx", range(2))
def test_moore_penrose_matches_reference_strings(idx):
    # color_correctionTest.cpp:32-84: the float Jacobi-SVD pseudo-inverse, down to its asymmetric last digits
    g = manifest()["moore_penrose_goldens"][idx]
    assert _fmt_matx(ORA.moore_penrose_lsm(g["actual"], g["desired"])) == g["matrix_str"]


@pytest.mark.parametrize("idx", range(2))
def test_init_ccm_matches_reference_strings(idx):
    # CimbReaderTest.cpp:181-199 / :239-257: update_metadata(FountainMetadata(0, 23586, 7)) then init_ccm on a camera frame
    g = manifest()["init_ccm_goldens"][idx]
    m = ORA.mode(68)
    hdr = np.zeros(6, np.uint8)
    ORA.lib.cbo_md_pack(g["md"][0], g["md"][1], g["md"][2], hdr.ctypes.data_as(C.POINTER(C.c_uint8)))
    size = g["md"][1]
    radioactive = 0xFFFFFFFF if size % 625 == 0 else size // 625       # computeRadioactiveBlockId, CimbReader.cpp:99-104
    nxt = g["md"][2] + 1                                               # update_metadata: "we always want to be +1"
    if nxt == radioactive:
        nxt += 1
    hdr[4], hdr[5] = (nxt >> 8) & 0xFF, nxt & 0xFF
    mat = ORA.init_ccm(m, load_sample(g["sample"]), hdr, radioactive)
    assert mat is not None and _fmt_matx(mat) == g["matrix_str"]


def test_decode_fountain_cc2_fits_from_the_frames_own_header():
    # tr_0.png is a real fountain frame: with color_correction == 2 the decode fits a CCM from its header (a gain of ~1.5:
    # the cell means include the tiles' black pixels) and keeps it; the chunks are the same as without
    m = ORA.mode(68)
    rgb = load_sample("b/tr_0.png")
    ORA.set_ccm(None)
    good0, chunks0, mask0 = ORA.decode_fountain(m, rgb, color_correction=0)
    assert ORA.get_ccm() is None
    good2, chunks2, mask2 = ORA.decode_fountain(m, rgb, color_correction=2)
    mat = ORA.get_ccm()
    ORA.set_ccm(None)
    assert mat is not None and (np.diag(mat) > 1.3).all() and (np.abs(mat - np.diag(np.diag(mat))) < 0.4).all()
    assert (good0, mask0) == (good2, mask2) and np.array_equal(chunks0, chunks2)


def _header_rule_closed_form(data, ok, blocks_per_chunk, msg, chunk_size):
    """the rule k_ccm_fit uses on the device (libcimbar_b200/csrc/ccm.cu): with whole RS blocks per chunk a chunk raises an
    event only when its last block is good -- a flush if all its blocks are good and no bad flag is pending, else the
    "bad chunk" callback; a bad last block passes the flag on"""
    hdr = [0] * 6
    radioactive, carry = 0, False
    for q in range(len(ok) // blocks_per_chunk):
        blk = ok[q * blocks_per_chunk:(q + 1) * blocks_per_chunk]
        if not blk[-1]:
            carry = True
            continue
        good = all(blk) and not carry
        carry = False
        id_zero = not any(hdr[:4])
        if not good and id_zero:
            continue
        if id_zero:
            hdr = list(data[q * blocks_per_chunk * msg:q * blocks_per_chunk * msg + 6])
        if radioactive == 0:
            fs = hdr[3] | (hdr[2] << 8) | (hdr[1] << 16) | ((hdr[0] & 0x80) << 17)
            radioactive = 0xFFFFFFFF if fs % chunk_size == 0 else fs // chunk_size
        nxt = ((hdr[4] << 8) | hdr[5]) + 1
        if nxt == radioactive:
            nxt += 1
        hdr[4], hdr[5] = (nxt >> 8) & 0xFF, nxt & 0xFF
    return hdr, radioactive, any(hdr[:4])


def test_header_rule_closed_form_equals_stream_replay():
    """the device derives the fountain header from per-chunk rules; the oracle replays aligned_stream byte by byte
    (aligned_stream.h:39-116 + CimbReader::update_metadata): both must agree for every pattern of good / bad RS blocks"""
    rng = np.random.default_rng(77)
    L = ORA.lib
    L.cbo_header_after_symbols.argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_uint, C.c_uint, C.c_uint,
                                           C.POINTER(C.c_uint8), C.POINTER(C.c_uint)]
    nblocks, msg, chunk = 40, 125, 625
    for trial in range(3000):
        p_bad = rng.choice([0.0, 0.05, 0.3, 0.7])
        ok = (rng.random(nblocks) >= p_bad).astype(np.uint8)
        data = rng.integers(0, 256, nblocks * msg, dtype=np.uint8)
        for q in range(8):                                   # headers: sometimes an all-zero id, small / aligned file sizes
            kind = rng.integers(0, 4)
            h = data[q * chunk:q * chunk + 6]
            if kind == 0:
                h[:4] = 0
            elif kind == 1:
                h[0] &= 0x7F; h[1] = 0; h[2] = rng.integers(0, 8); h[4] = 0; h[5] = rng.integers(0, 6)   # radioactive id nearby
            elif kind == 2:
                size = 625 * int(rng.integers(1, 3000)); h[0] = 1 | ((size >> 17) & 0x80); h[1] = (size >> 16) & 0xFF; h[2] = (size >> 8) & 0xFF; h[3] = size & 0xFF
        hdr = np.zeros(6, np.uint8)
        rad = C.c_uint(0)
        has = L.cbo_header_after_symbols(data.ctypes.data_as(C.POINTER(C.c_uint8)), ok.ctypes.data_as(C.POINTER(C.c_uint8)), nblocks, msg, chunk,
                                         hdr.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(rad))
        whdr, wrad, whas = _header_rule_closed_form(data.tolist(), ok.tolist(), 5, msg, chunk)
        assert bool(has) == bool(whas), trial
        if has:
            assert hdr.tolist() == whdr and rad.value == wrad, trial
