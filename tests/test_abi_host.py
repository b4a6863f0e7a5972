"""CPU-side checks of the product's host surface: the C-ABI library loads and exports every symbol include/cb200.h
declares, its mode table and interleave map agree with the oracle, and it refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import libcimbar_b200 as cb
from libcimbar_b200 import build as cbbuild
from oracle_lib import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORA = Oracle()


@pytest.fixture(scope="module", autouse=True)
def built():
    cbbuild.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cb200.h")).read()
    declared = sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(cb.EXPORTS)
    lib = cb.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cb200_version() >= 1


@pytest.mark.parametrize("mode_val", [68, 4, 8, 67, 66])
def test_mode_table_matches_oracle(mode_val):
    i, m = cb.mode_info(mode_val), ORA.mode(mode_val)
    assert (i.image_size_x, i.image_size_y, i.total_cells) == (m.image_size_x, m.image_size_y, m.total_cells)
    assert (i.symbol_bits, i.color_bits, i.ecc_bytes, i.ecc_block_size) == (m.symbol_bits, m.color_bits, m.ecc_bytes, m.ecc_block_size)
    assert i.raw_bytes == ORA.capacity(m) and i.chunk_size == m.chunk_size and i.chunks_per_frame == m.chunks_per_frame
    assert i.legacy_mode == m.legacy_mode
    assert i.data_bytes == (i.raw_bytes // i.ecc_block_size) * (i.ecc_block_size - i.ecc_bytes)
    idx = cb.interleave_indices(mode_val)
    want = np.zeros(m.total_cells, np.uint32)
    ORA.lib.cbo_interleave_indices(m.total_cells, m.interleave_blocks, m.interleave_partitions, want.ctypes.data_as(C.POINTER(C.c_uint)))
    assert np.array_equal(idx.astype(np.uint32), want)


@pytest.mark.parametrize("mode_val", [68, 4, 8, 67, 66])
def test_closed_form_adjacency_matches_literal_finder(mode_val):
    # the exact-walk kernel uses (row, column) arithmetic for AdjacentCellFinder; it must equal the literal evaluation,
    # which in turn equals the reference's own class (tests/test_oracle_vs_ref.py::test_cell_positions_and_adjacency)
    assert cb.load_library().cb200_selfcheck(mode_val) == 0


def test_unknown_mode_is_an_error():
    with pytest.raises(cb.Cb200Error):
        cb.mode_info(999)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.Cb200Error) as e:
        cb.Context(68, max_frames=1)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_mirrors_compile_against_the_references_own_stream_classes():
    """tests/cpp/shim_test.cpp drives the Decoder / CimbReader / CimbDecoder mirrors with the reference's call shapes; here it is
    compiled (syntax only, no GPU needed) once with the test-local chunk collector and, where the reference checkout exists,
    once against the reference's unmodified escrow_buffer_writer / aligned_stream headers"""
    import subprocess
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", src])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", os.path.join(ROOT, "tests", "cpp", "extractor_test.cpp")])   # Scanner / Extractor mirrors
    ref = "/root/reference/src/lib"
    if os.path.isdir(ref):
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-DCB200_WITH_REFERENCE_STREAMS", "-I" + ref, src])
    # nothing of the reference's stream classes is restated in the product headers
    hosts = os.path.join(ROOT, "libcimbar_b200", "host")
    text = "".join(open(os.path.join(hosts, f)).read() for f in os.listdir(hosts))
    assert "class aligned_stream" not in text and "class escrow_buffer_writer" not in text


def test_scan_entry_points_reject_bad_arguments_without_touching_a_gpu():
    # argument checks come before any CUDA call: a null context / null pictures is CB200_ERR_ARG (-1) on any machine
    lib = cb.load_library()
    cnt = (C.c_int32 * 1)()
    buf = np.zeros(64 * 64 * 3, np.uint8)
    assert lib.cb200_scan(None, buf.ctypes.data, 64, 64, 1, None, cnt, None) == -1
    assert lib.cb200_scan_dev(None, None, 64, 64, 1, None, cnt, None) == -1
    st = (C.c_int32 * 1)()
    cc = (C.c_uint32 * 1)()
    assert lib.cb200_scan_extract_decode_fountain(None, buf.ctypes.data, 64, 64, 1, 0, buf.ctypes.data, cc, None, None, st) == -1
    assert lib.cb200_scan_blurred(None, None, None, 64, 64, 1) == -1
    assert b"bad arguments" in lib.cb200_last_error() or b"scan" in lib.cb200_last_error()
