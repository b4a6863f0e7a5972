#!/usr/bin/env python3
"""Diagnostic: the sharpen decode of four clean frames against the oracle, frame by frame (which frames / how many bytes differ, frame flags)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import libcimbar_b200 as cb
from oracle_lib import Oracle

ora = Oracle()
m = ora.mode(68)
rng = np.random.default_rng(1)
payloads = rng.integers(0, 256, (3, 7500), dtype=np.uint8)
frames = np.stack([ora.render_frame(m, ora.payload_to_cells(m, p)) for p in payloads] + [ora.render_frame(m, ora.payload_to_cells(m, payloads[0]))])
want = [ora.decode_raw(m, fr, sharpen=True) for fr in frames]
ctx = cb.Context(68, max_frames=4)
for rep in range(3):
    raw, ff = ctx.decode_raw(frames, flags=cb.FLAG_SHARPEN)
    print("rep", rep, "flags", ff.tolist(), "differing bytes per frame", [int((raw[f] != want[f]).sum()) for f in range(4)])
    for f in range(4):
        d = np.nonzero(raw[f] != want[f])[0]
        if d.size:
            print("   frame", f, "first differing byte offsets", d[:12].tolist(), "got", raw[f][d[:6]].tolist(), "want", want[f][d[:6]].tolist())
raw1, ff1 = ctx.decode_raw(frames[:1], flags=cb.FLAG_SHARPEN)
print("one frame: flags", ff1.tolist(), "differing", int((raw1[0] != want[0]).sum()))
print("diag done")
