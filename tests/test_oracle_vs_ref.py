"""Checks the restatement (oracle/cimbar_oracle.c) against the reference's own OpenCV-free code compiled
unmodified into oracle/_ref/libcimbar_ref.so (oracle/Makefile `ref`, oracle/ref_shim.cpp):
libcorrect RS incl. its failure/miscorrection behaviour, the flood walk on the real std::priority_queue,
cell geometry, interleave, the RS->aligned_stream->escrow stack, FountainMetadata, wirehair."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import Oracle, Ref, _ptr

ORA = Oracle()
try:
    REF = Ref()
except (FileNotFoundError, OSError) as e:  # pragma: no cover
    REF = None
    pytestmark = pytest.mark.skip(reason=f"oracle/_ref not available: {e}")

MODES = [68, 4, 8, 67, 66]


@pytest.mark.parametrize("mode", MODES)
def test_cell_positions_and_adjacency(mode):
    m = ORA.mode(mode)
    n = m.total_cells
    xs, ys = np.zeros(n, np.int32), np.zeros(n, np.int32)
    rxs, rys = np.zeros(n, np.int32), np.zeros(n, np.int32)
    assert ORA.lib.cbo_cell_positions(C.byref(m), 0, _ptr(xs, C.c_int), _ptr(ys, C.c_int)) == n
    assert REF.lib.ref_cell_positions(*REF.grid_args(m), _ptr(rxs, C.c_int), _ptr(rys, C.c_int)) == n
    assert np.array_equal(xs, rxs) and np.array_equal(ys, rys)
    if mode == 68:  # CellPositionsTest.cpp:11-27
        assert (xs[0], ys[0]) == (62, 8) and n == 12400
    a, ra = np.zeros(4, np.int32), np.zeros(4, np.int32)
    rng = np.random.default_rng(1)
    for idx in list(rng.integers(0, n, 300)) + [0, n - 1, 99, 100, 599, 600, 711, 712, n - 600, n - 601]:
        ORA.lib.cbo_adjacent(C.byref(m), _ptr(xs, C.c_int), n, int(idx), _ptr(a, C.c_int))
        REF.lib.ref_adjacent(*REF.grid_args(m), int(idx), _ptr(ra, C.c_int))
        assert np.array_equal(a, ra), idx


@pytest.mark.parametrize("args", [(12400, 155, 2), (12400, 155, 1), (12400, 0, 2), (8592, 179, 2), (5376, 168, 2)])
def test_interleave_reverse(args):
    size = args[0]
    a, b = np.zeros(size, np.uint32), np.zeros(size, np.uint32)
    ORA.lib.cbo_interleave_reverse(*args, _ptr(a, C.c_uint))
    REF.lib.ref_interleave_reverse(*args, _ptr(b, C.c_uint))
    assert np.array_equal(a, b)
    if args == (12400, 155, 1):  # InterleaveTest.cpp:14-58
        assert a[1] == 80


@pytest.mark.parametrize("mode", [68, 67, 66])
@pytest.mark.parametrize("noise", [0, 3, 30, 100])
def test_flood_walk_matches_std_priority_queue(mode, noise):
    m = ORA.mode(mode)
    n = m.total_cells
    for seed in range(4):
        o, d, c = np.zeros(n, np.uint16), np.zeros(2 * n, np.int8), np.zeros(n, np.uint8)
        ro, rd, rc = np.zeros(n, np.uint16), np.zeros(2 * n, np.int8), np.zeros(n, np.uint8)
        assert ORA.lib.cbo_flood_walk_synthetic(C.byref(m), seed, noise, _ptr(o, C.c_uint16), _ptr(d, C.c_int8), _ptr(c)) == n
        assert REF.lib.ref_flood_walk_synthetic(*REF.grid_args(m), seed, noise, _ptr(ro, C.c_uint16), _ptr(rd, C.c_int8), _ptr(rc)) == n
        assert np.array_equal(o, ro) and np.array_equal(d, rd) and np.array_equal(c, rc)
        assert sorted(o.tolist()) == list(range(n))
    if mode == 68 and noise == 0:  # FloodDecodePositionsTest.cpp:11-129 seeds; SURVEY appendix A.6 clean walk order
        assert o[:4].tolist() == [0, 12399, 99, 12300]


def test_fuzzy_ahash_vs_reference_extractors():
    rng = np.random.default_rng(7)
    w, h = 64, 40
    bits = rng.integers(0, 256, w * h // 8, dtype=np.uint8)
    mine, ref = (C.c_uint64 * 9)(), (C.c_uint64 * 9)()
    for _ in range(200):
        wx, wy, all_ = int(rng.integers(0, w - 10)), int(rng.integers(0, h - 10)), int(rng.integers(0, 2))
        ORA.lib.cbo_fuzzy_ahash(_ptr(bits), w, wx, wy, all_, mine)
        REF.lib.ref_fuzzy_ahash(_ptr(bits), bits.size, w, wx, wy, all_, ref)
        assert list(mine) == list(ref)


def _rs_pair(parity):
    o = ORA.lib.cbo_rs_create(parity)
    r = REF.lib.ref_rs_create(parity)
    return o, r


def test_rs_known_answer():
    # reed_solomon_streamTest.cpp:23: RS(155,140) parity of the ASCII digits message
    msg = np.frombuffer((b"0123456789" * 14), dtype=np.uint8).copy()
    o, r = _rs_pair(15)
    enc, renc = np.zeros(255, np.uint8), np.zeros(255, np.uint8)
    assert ORA.lib.cbo_rs_encode(o, _ptr(msg), 140, _ptr(enc)) == 255
    assert REF.lib.ref_rs_encode2(r, _ptr(msg), 140, _ptr(renc)) == 255
    assert np.array_equal(enc[:155], renc[:155])
    assert enc[140:155].tobytes().hex() == "a4740203 72c3adf2 60c5b69e 267873".replace(" ", "")


@pytest.mark.parametrize("parity,block", [(30, 155), (33, 168), (36, 179), (40, 216), (15, 155)])
def test_rs_decode_matches_libcorrect_including_failures(parity, block):
    rng = np.random.default_rng(parity * 1000 + block)
    o, r = _rs_pair(parity)
    msg_len = block - parity
    n_fail = n_miscorrect = 0
    # one persistent decoder object per side: libcorrect keeps state across calls
    for trial in range(1500):
        msg = rng.integers(0, 256, msg_len, dtype=np.uint8)
        enc = np.zeros(255, np.uint8)
        ORA.lib.cbo_rs_encode(o, _ptr(msg), msg_len, _ptr(enc))
        blk = enc[:block].copy()
        kind = trial % 6
        t = parity // 2
        nerr = [0, 1, t, t + 1, int(rng.integers(1, parity + 8)), int(rng.integers(t - 2, t + 3))][kind]
        nerr = max(0, min(nerr, block))
        pos = rng.choice(block, nerr, replace=False)
        blk[pos] ^= rng.integers(1, 256, nerr, dtype=np.uint8)
        if trial % 97 == 0:
            blk = rng.integers(0, 256, block, dtype=np.uint8)  # pure noise
        a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
        na = ORA.lib.cbo_rs_decode(o, _ptr(blk), block, _ptr(a))
        nb = REF.lib.ref_rs_decode2(r, _ptr(blk), block, _ptr(b))
        assert na == nb, (trial, nerr)
        if nb > 0:
            assert np.array_equal(a[:msg_len], b[:msg_len]), (trial, nerr)
            if nerr <= t:
                assert np.array_equal(a[:msg_len], msg)
            elif not np.array_equal(b[:msg_len], msg):
                n_miscorrect += 1
        else:
            n_fail += 1
    assert n_fail > 0  # the sweep really exercised libcorrect's failure path
    ORA.lib.cbo_rs_destroy(o)
    REF.lib.ref_rs_destroy(r)


@pytest.mark.parametrize("mode", [68, 4, 67, 66])
def test_rs_stream_and_aligned_chunks_match_reference_stack(mode):
    m = ORA.mode(mode)
    rng = np.random.default_rng(mode)
    cap = ORA.capacity(m)
    block, parity = m.ecc_block_size, m.ecc_bytes
    msg = block - parity
    nblocks = cap // block
    if m.legacy_mode:
        sym_len, col_len = cap, 0
    else:
        sym_len, col_len = ORA.capacity(m, m.symbol_bits), ORA.capacity(m, m.color_bits)
    rs = ORA.lib.cbo_rs_create(parity)
    for trial in range(40):
        payload = rng.integers(0, 256, nblocks * msg, dtype=np.uint8)
        raw = np.zeros(cap, np.uint8)
        enc = np.zeros(255, np.uint8)
        for b in range(nblocks):
            ORA.lib.cbo_rs_encode(rs, _ptr(payload[b * msg:(b + 1) * msg].copy()), msg, _ptr(enc))
            raw[b * block:(b + 1) * block] = enc[:block]
        # break some blocks beyond repair; always include the last block of a chunk / of the symbol stream
        bad = set(rng.choice(nblocks, int(rng.integers(0, 6)), replace=False).tolist())
        per_chunk = m.chunk_size // msg
        if trial % 4 == 1:
            bad.add(per_chunk - 1)
        if trial % 4 == 2 and not m.legacy_mode:
            bad.add(sym_len // block - 1)
        if trial % 4 == 3:
            bad.add(nblocks - 1)
        for b in bad:
            raw[b * block:b * block + parity] = rng.integers(0, 256, parity, dtype=np.uint8)
        # (a) Decoder::decode semantics
        out, ok = ORA.rs_stream(parity, block, raw[:sym_len])
        ref_out = np.zeros(nblocks * msg + 16, np.uint8)
        n = REF.lib.ref_rs_stream(parity, block, _ptr(raw), sym_len, _ptr(ref_out))
        assert n == out.size and np.array_equal(out, ref_out[:n])
        # (b) Decoder::decode_fountain semantics: oracle = rs_stream + align_chunks; reference = its stream stack
        data = np.zeros(nblocks * msg, np.uint8)
        okall = np.zeros(nblocks, np.uint8)
        o1, k1 = ORA.rs_stream(parity, block, raw[:sym_len])
        data[:o1.size], okall[:k1.size] = o1, k1
        if col_len:
            o2, k2 = ORA.rs_stream(parity, block, raw[sym_len:])
            data[o1.size:], okall[k1.size:] = o2, k2
        chunks = np.zeros((m.chunks_per_frame, m.chunk_size), np.uint8)
        mask = C.c_uint32(0)
        good = ORA.lib.cbo_align_chunks(_ptr(data), _ptr(okall), nblocks, msg, m.chunk_size, _ptr(chunks), C.byref(mask))
        rchunks = np.zeros((m.chunks_per_frame, m.chunk_size), np.uint8)
        used = C.c_uint(0)
        rgood = REF.lib.ref_rs_align_escrow(parity, block, _ptr(raw), sym_len, col_len, m.chunk_size, m.chunks_per_frame, _ptr(rchunks), C.byref(used))
        assert good == rgood
        assert used.value == bin(mask.value).count("1") and good == used.value * m.chunk_size
        assert np.array_equal(chunks, rchunks)
    ORA.lib.cbo_rs_destroy(rs)


def test_fountain_metadata():
    rng = np.random.default_rng(3)
    a, b = np.zeros(6, np.uint8), np.zeros(6, np.uint8)
    for _ in range(200):
        eid, size, bid = int(rng.integers(0, 256)), int(rng.integers(0, 1 << 25)), int(rng.integers(0, 65536))
        ORA.lib.cbo_md_pack(eid, size, bid, _ptr(a))
        REF.lib.ref_md_pack(eid, size, bid, _ptr(b))
        assert np.array_equal(a, b)
        assert ORA.lib.cbo_md_file_size(_ptr(a)) == REF.lib.ref_md_file_size(_ptr(a)) == size
        assert ORA.lib.cbo_md_block_id(_ptr(a)) == REF.lib.ref_md_block_id(_ptr(a)) == bid
        assert ORA.lib.cbo_md_encode_id(_ptr(a)) == REF.lib.ref_md_encode_id(_ptr(a)) == eid & 0x7F


def test_sample_stream_reassembles_with_reference_wirehair():
    # samples/b/tr_0..3.png are a complete 4-frame fountain stream (.github/workflows/wasm.yml:48-49):
    # oracle decode_fountain -> chunk headers -> the reference's wirehair -> recovered (zstd) file
    from oracle_lib import load_sample
    m = ORA.mode(68)
    codec, size = None, None
    seen = set()
    done = False
    for k in range(4):
        good, chunks, mask = ORA.decode_fountain(m, load_sample(f"b/tr_{k}.png"))
        assert good == 7500 and mask == 0xFFF
        for ch in chunks:
            fsize = ORA.lib.cbo_md_file_size(_ptr(ch))
            bid = ORA.lib.cbo_md_block_id(_ptr(ch))
            assert ORA.lib.cbo_md_encode_id(_ptr(ch)) == 0
            if codec is None:
                size = fsize
                codec = REF.lib.wirehair_decoder_create(None, size, m.chunk_size - 6)
            assert fsize == size == 23586
            if bid in seen or done:
                continue
            seen.add(bid)
            payload = ch[6:].copy()
            res = REF.lib.wirehair_decode(codec, bid, _ptr(payload, C.c_uint8), payload.size)
            done = res == 0
    assert done
    out = np.zeros(size, np.uint8)
    assert REF.lib.wirehair_recover(codec, _ptr(out), size) == 0
    assert out[:4].tobytes() == b"\x28\xb5\x2f\xfd" or out[:4].tobytes()[0] in (0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x5b, 0x5c, 0x5d, 0x5e, 0x5f)  # zstd or skippable-frame magic
    REF.lib.wirehair_free(codec)
