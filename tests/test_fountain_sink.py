"""Rank-0 fountain ingest (cb200_sink_*, host only): header parse / de-dup / stream bookkeeping restated from
src/lib/fountain/fountain_decoder_sink.h:133-166, with the reference's own wirehair (oracle/_ref) as the codec.
Expectations follow src/lib/fountain/test/fountain_sinkTest.cpp and FountainEncodingTest.cpp:49-98."""
import ctypes as C

import numpy as np
import pytest

import libcimbar_b200 as cb
from libcimbar_b200 import build as cbbuild
from oracle_lib import Oracle, Ref, load_sample, _ptr

ORA = Oracle()
try:
    REF = Ref()
except (FileNotFoundError, OSError) as e:  # pragma: no cover
    REF = None
    pytestmark = pytest.mark.skip(reason=f"oracle/_ref not available: {e}")


@pytest.fixture(scope="module", autouse=True)
def built():
    cbbuild.build()


def make_stream(size, chunk_size, encode_id, n_blocks, seed):
    """wirehair-encode a random file into n_blocks chunks with FountainMetadata headers (fountain_encoder_stream.h:87)"""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, size, dtype=np.uint8)
    enc = REF.lib.wirehair_encoder_create(None, data.ctypes.data, size, chunk_size - 6)
    chunks = np.zeros((n_blocks, chunk_size), np.uint8)
    for b in range(n_blocks):
        ORA.lib.cbo_md_pack(encode_id, size, b, _ptr(chunks[b]))
        wrote = C.c_uint32(0)
        assert REF.lib.wirehair_encode(enc, b, chunks[b, 6:].ctypes.data, chunk_size - 6, C.byref(wrote)) == 0
    REF.lib.wirehair_free(enc)
    return data, chunks


def test_sample_stream_reassembles():
    m = ORA.mode(68)
    sink = cb.FountainSink(m.chunk_size, REF.lib)
    fid = 0
    for k in range(4):
        good, chunks, mask = ORA.decode_fountain(m, load_sample(f"b/tr_{k}.png"))
        r = sink.ingest(chunks, np.array([mask], np.uint32))
        fid = r or fid
    assert fid > 0
    out = sink.file(fid)
    assert out is not None and out.size == 23586
    # same bytes as driving wirehair directly
    codec = REF.lib.wirehair_decoder_create(None, 23586, m.chunk_size - 6)
    seen, done = set(), False
    for k in range(4):
        _, chunks, _ = ORA.decode_fountain(m, load_sample(f"b/tr_{k}.png"))
        for ch in chunks:
            bid = (int(ch[4]) << 8) | int(ch[5])
            if bid in seen or done:
                continue
            seen.add(bid)
            payload = ch[6:].copy()
            done = REF.lib.wirehair_decode(codec, bid, payload.ctypes.data, payload.size) == 0
    want = np.zeros(23586, np.uint8)
    assert REF.lib.wirehair_recover(codec, want.ctypes.data, 23586) == 0
    REF.lib.wirehair_free(codec)
    assert np.array_equal(out, want)
    sink.close()


def test_loss_reorder_duplicates_and_error_codes():
    size, cs = 10000, 626
    data, chunks = make_stream(size, cs, encode_id=5, n_blocks=40, seed=1)
    sink = cb.FountainSink(cs, REF.lib)
    assert sink.decode_frame(chunks[0][:5]) == -10                     # shorter than a header
    zero = chunks[0].copy(); zero[:4] = 0
    assert sink.decode_frame(zero) == -11                              # size 0
    order = np.random.default_rng(2).permutation(40)
    order = [b for b in order if b % 3 != 0]                           # drop a third of the blocks, shuffled order
    fid, fed = 0, 0
    for b in order:
        r = sink.decode_frame(chunks[b])
        assert r >= 0
        fed += 1
        if r > 0:
            fid = r
            break
        assert sink.decode_frame(chunks[b]) == 0                      # duplicate block id: ignored, never fed twice
    assert fid > 0 and fed >= size // (cs - 6) + 1
    assert np.array_equal(sink.file(fid), data)
    assert sink.decode_frame(chunks[1]) == -1                          # already done
    # a second stream with the same encode_id slot but another size is rejected while the first is live
    data2, chunks2 = make_stream(5000, cs, encode_id=6, n_blocks=12, seed=3)
    data3, chunks3 = make_stream(7000, cs, encode_id=6 + 8, n_blocks=12, seed=4)   # same slot (id & 7), other size
    assert sink.decode_frame(chunks2[0]) == 0
    assert sink.decode_frame(chunks3[0]) == -12
    sink.close()


def test_eight_concurrent_streams():
    cs = 626
    sink = cb.FountainSink(cs, REF.lib)
    streams = [make_stream(3000 + 100 * i, cs, encode_id=i, n_blocks=10, seed=10 + i) for i in range(8)]
    done = {}
    for b in range(10):
        for i, (data, chunks) in enumerate(streams):
            r = sink.decode_frame(chunks[b])
            if r > 0:
                done[i] = r
    assert len(done) == 8
    for i, (data, _) in enumerate(streams):
        assert np.array_equal(sink.file(done[i]), data)
    sink.close()
