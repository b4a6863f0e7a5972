#!/usr/bin/env python3
"""Tiny decode for compute-sanitizer (memcheck / racecheck / synccheck): 3 clean frames + 1 camera frame, all entry points."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import libcimbar_b200 as cb
from oracle_lib import Oracle, load_sample

ora = Oracle()
m = ora.mode(68)
rng = np.random.default_rng(1)
payloads = rng.integers(0, 256, (3, 7500), dtype=np.uint8)
cam = [load_sample("b/ex2434.jpg")] if os.environ.get("SANITIZE_CAMERA", "1") == "1" else [ora.render_frame(m, ora.payload_to_cells(m, payloads[0]))]
frames = np.stack([ora.render_frame(m, ora.payload_to_cells(m, p)) for p in payloads] + cam)
ctx = cb.Context(68, max_frames=4)
raw, ff = ctx.decode_raw(frames)
data, ok, _ = ctx.decode(frames)
assert np.array_equal(data[:3], payloads) and ff.tolist()[:3] == [0, 0, 0]
for f in range(4):
    assert np.array_equal(raw[f], ora.decode_raw(m, frames[f]))
# colour correction 1 (per-frame von Kries) and 2 (header fit: means pass, symbol RS, fit, carry, apply, colour RS)
raw1, _ = ctx.decode_raw(frames, flags=cb.FLAG_CC_SIMPLE)
for f in range(4):
    assert np.array_equal(raw1[f], ora.decode_raw(m, frames[f], color_correction=1))
ora.set_ccm(None)
ctx.set_ccm(None)
chunks, counts, masks, _ = ctx.decode_fountain(frames, flags=cb.FLAG_CC_FIT)
for f in range(4):
    good, wchunks, wmask = ora.decode_fountain(m, frames[f], color_correction=2)
    assert masks[f] == wmask and np.array_equal(chunks[f][:counts[f]], wchunks[:counts[f]])
assert np.array_equal(ctx.get_ccm(), ora.get_ccm())
ora.set_ccm(None)
# deskew (csrc/deskew.cu): a mildly skewed "camera" of frame 0 back to the frame, then the decode on the device
# (SANITIZE_CAMERA=0: the anchor centres where Deskewer puts them = an identity warp, so that the frame stays clean and the run
#  never enters the exact flood walk -- racecheck of everything but K1x)
corners = (np.array([[34, 28, 990, 33, 29, 996, 997, 991]], np.float32) if os.environ.get("SANITIZE_CAMERA", "1") == "1"
           else np.array([[30, 30, 994, 30, 30, 994, 994, 994]], np.float32))
out = ctx.extract_decode_fountain(frames[:1], corners)
assert out[0].shape[0] == 1
# five frames: an odd count for the frame-pair RS kernel
ctx5 = cb.Context(68, max_frames=5)
five = np.concatenate([frames[:3], frames[:2]])
d5, ok5, _ = ctx5.decode(five)
assert np.array_equal(d5[:3], payloads) and np.array_equal(d5[3:], payloads[:2]) and ok5.all()
# second half of round 2: sharpen inside K1 (two barriers per stage, extra halo array), the walk's streaming sharpen raster, the anchor
# scan (both blur kernels, Otsu, one CTA per picture) and scan + deskew (aligned word loads) + decode in one call
ctx.set_ccm(None)                                   # the CC_FIT batch above may have left a fitted matrix in the context
raw_s, ff_s = ctx.decode_raw(frames, flags=cb.FLAG_SHARPEN)
for f in range(4):
    assert np.array_equal(raw_s[f], ora.decode_raw(m, frames[f], sharpen=True))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scan_oracle_lib import ScanOracle
so = ScanOracle()
pic = load_sample("6bit/4_30_f2_734.jpg")[:, 1:958]              # 1280 x 957: rows that are not word aligned
pic2 = load_sample("6bit/4color_ecc30_fountain_0.png")
for p_ in (pic, pic2):
    for blur in ("1", "0"):
        os.environ["CB200_SCAN_BLUR"] = blur
        anchors, count, cutoff = ctx.scan(p_)
        want, want_cutoff = so.scan(p_)
        assert count[0] == len(want) and [tuple(int(v) for v in a) for a in anchors[0][:count[0]]] == want and cutoff[0] == want_cutoff
    del os.environ["CB200_SCAN_BLUR"]
ctx4 = cb.Context(4, max_frames=2)
chunks_c, count_c, mask_c, ff_c, status_c = ctx4.scan_extract_decode_fountain(np.stack([pic2, pic2[::-1].copy()]), flags=cb.FLAG_SHARPEN)
assert status_c[0] > 0 and count_c[0] > 0
print("sanitize_small: ok")
