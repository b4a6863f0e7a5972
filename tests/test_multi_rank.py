"""The N>1 path on CPU (gloo, world_size 2): frames shard one-per-rank, each rank produces fixed-slot chunk records,
ONE gather brings them to rank 0, rank 0 runs the fountain ingest.  On the GPU box the records come from
cb200_decode_chunks_dev and the backend is NCCL (bench.py); here the per-rank decode is stood in for by the oracle so the
sharding / gather / ingest logic is exercised without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import libcimbar_b200 as cb
    from libcimbar_b200 import dist as cbdist
    from oracle_lib import Oracle, Ref, load_sample
    ora = Oracle()
    m = ora.mode(68)
    mine = cbdist.shard_frames(4, rank, world)                     # 4 sample frames: rank r takes r, r+2
    chunks = np.zeros((len(mine), m.chunks_per_frame * m.chunk_size), np.uint8)
    masks = np.zeros(len(mine), np.int32)
    for j, f in enumerate(mine):
        good, ch, mask = ora.decode_fountain(m, load_sample(f"b/tr_{f}.png"))
        chunks[j] = ch.reshape(-1)
        masks[j] = mask
    if rank == 1:                                                  # a rank that lost one chunk (RS failure) still contributes
        masks[0] &= ~1
    gc, gm = cbdist.gather_records(torch.from_numpy(chunks), torch.from_numpy(masks), dst=0)
    if rank == 0:
        sink = cb.FountainSink(m.chunk_size, Ref().lib)
        fid = 0
        for c, k in zip(gc, gm):
            r = sink.ingest(c.numpy(), k.numpy().astype(np.uint32))
            fid = r or fid
        out = sink.file(fid) if fid else None
        np.save(result_path, out if out is not None else np.zeros(0, np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_ingest(tmp_path):
    from oracle_lib import build_ref
    if build_ref() is None:
        pytest.skip("oracle/_ref not available")
    from libcimbar_b200 import build as cbbuild
    cbbuild.build()
    result = str(tmp_path / "file.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
    out = np.load(result)
    assert out.size == 23586          # the 4-frame sample stream reassembles although one chunk was dropped
    # identical to the single-process reassembly
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import libcimbar_b200 as cb
    from oracle_lib import Oracle, Ref, load_sample
    ora = Oracle()
    m = ora.mode(68)
    sink = cb.FountainSink(m.chunk_size, Ref().lib)
    fid = 0
    for k in range(4):
        _, ch, mask = ora.decode_fountain(m, load_sample(f"b/tr_{k}.png"))
        fid = sink.ingest(ch, np.array([mask], np.uint32)) or fid
    assert np.array_equal(out, sink.file(fid))
