#!/usr/bin/env python3
"""Tuning helper: K1 kernel time vs the CB200_K1_* knobs (read at context creation). Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libcimbar_b200 as cb

B = int(os.environ.get("SWEEP_FRAMES", "4000"))
knob = sys.argv[1] if len(sys.argv) > 1 else "CB200_K1_L2_AHEAD"
values = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2", "4", "8", "16"]

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
frames = payload = None
for v in values:
    os.environ[knob] = v
    ctx = cb.Context(68, max_frames=B)
    ctx.set_stream(stream.cuda_stream)
    info = ctx.info
    if frames is None:
        payload = torch.randint(0, 256, (B, info.data_bytes), dtype=torch.uint8, device=dev)
        cells = torch.empty((B, info.total_cells), dtype=torch.uint8, device=dev)
        ctx.encode_cells_dev(payload.data_ptr(), B, cells.data_ptr())
        frames = torch.empty((B, 1024, 1024, 3), dtype=torch.uint8, device=dev)
        ctx.render_frames_dev(cells.data_ptr(), B, frames.data_ptr())
        chunks = torch.empty((B, info.data_bytes), dtype=torch.uint8, device=dev)
        mask = torch.empty(B, dtype=torch.int32, device=dev)
    for _ in range(2):
        ctx.decode_chunks_dev(frames.data_ptr(), B, chunks.data_ptr(), mask.data_ptr())
    ctx.set_timing(True)
    for _ in range(5):
        ctx.decode_chunks_dev(frames.data_ptr(), B, chunks.data_ptr(), mask.data_ptr())
    torch.cuda.synchronize()
    ms = [ctx.get_timing(i) for i in range(5)]
    k1 = sorted(r[0] for r in ms)[2]
    tot = sorted(sum(r) for r in ms)[2]
    ok = bool(torch.equal(chunks, payload))
    print(f"{knob}={v}: K1 {k1:.3f} ms -> {B / k1 / 1e3:.3f} Mfps ({B * 3158128 / k1 / 1e6:.0f} GB/s); pipeline {tot:.3f} ms -> {B / tot / 1e3:.3f} Mfps; "
          f"stages {[round(sorted(r[i] for r in ms)[2], 3) for i in range(5)]} parity={'ok' if ok else 'FAIL'}", flush=True)
    ctx.close()
