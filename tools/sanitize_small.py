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
print("sanitize_small: ok")
