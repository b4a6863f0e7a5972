/*
 * ref_shim.cpp -- thin extern "C" wrappers around the REFERENCE's own OpenCV-free sources,
 * compiled from where they lie under /root/reference (see oracle/Makefile, target _ref).
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no algorithm of its own: every function
 * forwards to reference code (#included / linked unmodified) so that tests can compare the
 * restatement in cimbar_oracle.c, and the CUDA product, with the real thing:
 *   - libcorrect Reed-Solomon           (src/third_party_lib/libcorrect/src/reed-solomon/)
 *   - wirehair fountain codec           (src/third_party_lib/wirehair/)   [linked; C API used directly]
 *   - FloodDecodePositions + std::priority_queue, CellPositions, AdjacentCellFinder, CellDrift
 *   - Interleave, bitbuffer, ahash_result/bit_extractor, reed_solomon_stream, aligned_stream,
 *     escrow_buffer_writer, FountainMetadata
 *   - the extractor's OpenCV-free pieces: ScanState_114 / ScanState_122, Anchor
 * The OpenCV-dependent files (CimbReader.cpp, CimbDecoder.cpp, Cell.h, bitmatrix.h, average_hash.h,
 * Decoder.h) cannot be compiled here (no C++ OpenCV in the image); those are pinned by the
 * SHA-256 goldens instead (tests/test_oracle_goldens.py).
 */
#include "cimb_translator/AdjacentCellFinder.h"
#include "cimb_translator/CellDrift.h"
#include "cimb_translator/CellPositions.h"
#include "cimb_translator/FloodDecodePositions.h"
#include "cimb_translator/Interleave.h"
#include "bit_file/bitbuffer.h"
#include "image_hash/ahash_result.h"
#include "encoder/ReedSolomon.h"
#include "encoder/reed_solomon_stream.h"
#include "encoder/aligned_stream.h"
#include "encoder/escrow_buffer_writer.h"
#include "fountain/FountainMetadata.h"
#include "extractor/Anchor.h"
#include "extractor/ScanState.h"

#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {
// same synthetic per-cell result function as the oracle (oracle/cimbar_oracle.c), linked in
void cbo_synth_result(unsigned seed, unsigned i, int dx, int dy, unsigned cooldown, unsigned noise, unsigned* drift_offset, unsigned* dist);

int ref_rs_decode(unsigned parity, const uint8_t* enc, unsigned len, uint8_t* msg)
{
	ReedSolomon rs(parity);
	return (int)rs.decode(reinterpret_cast<const char*>(enc), len, reinterpret_cast<char*>(msg));
}

// persistent decoder object: exposes libcorrect's cross-call state exactly as reed_solomon_stream uses it
void* ref_rs_create(unsigned parity) { return new ReedSolomon(parity); }
void ref_rs_destroy(void* p) { delete static_cast<ReedSolomon*>(p); }
int ref_rs_decode2(void* p, const uint8_t* enc, unsigned len, uint8_t* msg)
{
	return (int)static_cast<ReedSolomon*>(p)->decode(reinterpret_cast<const char*>(enc), len, reinterpret_cast<char*>(msg));
}
int ref_rs_encode2(void* p, const uint8_t* msg, unsigned len, uint8_t* enc)
{
	return (int)static_cast<ReedSolomon*>(p)->encode(reinterpret_cast<const char*>(msg), len, reinterpret_cast<char*>(enc));
}

int ref_cell_positions(unsigned sx, unsigned sy, unsigned dx, unsigned dy, int offset, unsigned mx, unsigned my, int* xs, int* ys)
{
	CellPositions::positions_list pos = CellPositions::compute_linear({sx, sy}, {dx, dy}, offset, {mx, my});
	for (size_t i = 0; i < pos.size(); ++i) { xs[i] = pos[i].first; ys[i] = pos[i].second; }
	return (int)pos.size();
}

void ref_adjacent(unsigned sx, unsigned sy, unsigned dx, unsigned dy, int offset, unsigned mx, unsigned my, int index, int* adj)
{
	CellPositions::positions_list pos = CellPositions::compute_linear({sx, sy}, {dx, dy}, offset, {mx, my});
	AdjacentCellFinder f(pos, {dx, dy}, {mx, my});
	std::array<int,4> a = f.find(index);
	for (int k = 0; k < 4; ++k) adj[k] = a[k];
}

void ref_interleave_reverse(unsigned size, unsigned chunks, unsigned partitions, unsigned* inv)
{
	std::vector<unsigned> v = Interleave::interleave_reverse(size, chunks, partitions);
	std::memcpy(inv, v.data(), v.size() * sizeof(unsigned));
}

// the reference's flood walk (real std::priority_queue) driven by the shared synthetic result function
int ref_flood_walk_synthetic(unsigned sx, unsigned sy, unsigned dx, unsigned dy, int offset, unsigned mx, unsigned my,
                             unsigned seed, unsigned noise, uint16_t* order_out, int8_t* drift_out, uint8_t* cooldown_out)
{
	FloodDecodePositions fp({sx, sy}, {dx, dy}, offset, {mx, my});
	int n = 0;
	while (!fp.done())
	{
		auto [i, xy, drift, cooldown] = fp.next();
		unsigned off, dist;
		cbo_synth_result(seed, i, drift.x(), drift.y(), cooldown, noise, &off, &dist);
		order_out[n] = (uint16_t)i; drift_out[2*n] = (int8_t)drift.x(); drift_out[2*n+1] = (int8_t)drift.y(); cooldown_out[n] = cooldown;
		std::pair<int,int> best = CellDrift::driftPairs[off];
		drift.updateDrift(best.first, best.second);
		fp.update(i, drift, dist, CellDrift::calculate_cooldown(cooldown, off));
		++n;
	}
	return n;
}

// fuzzy_ahash<8>(bitmatrix) body (average_hash.h:63-75) re-using the reference bitbuffer + ahash_result
void ref_fuzzy_ahash(const uint8_t* bits, unsigned nbytes, unsigned width, unsigned wx, unsigned wy, int mode_all, uint64_t* out9)
{
	bitbuffer bb(nbytes);
	bb.copy_to_buffer(reinterpret_cast<const char*>(bits), nbytes);
	const unsigned readlen = 10;
	intx::uint128 res(0);
	int bitpos = readlen*readlen - readlen;
	for (unsigned i = 0; i < readlen; ++i, bitpos -= readlen)
	{
		intx::uint128 r = bb.read(wx + (wy + i) * width, readlen);
		res |= r << bitpos;
	}
	image_hash::ahash_result<8> hr(res, mode_all ? image_hash::ahash_result<8>::ALL : image_hash::ahash_result<8>::FAST);
	for (unsigned k = 0; k < 9; ++k) out9[k] = hr[k];
}

// bitbuffer::write round trip (MSB-first packing used for the output streams)
void ref_bitbuffer_write(uint8_t* buf, unsigned nbytes, const unsigned* values, const unsigned* positions, unsigned n, int length)
{
	bitbuffer bb(nbytes);
	for (unsigned k = 0; k < n; ++k) bb.write(values[k], positions[k], length);
	std::memcpy(buf, bb.buffer().data(), nbytes);
}

// Decoder::decode_fountain's stream stack: reed_solomon_stream -> aligned_stream -> escrow_buffer_writer
// raw = symbol stream then colour stream (or one coupled stream when sym_len == total)
unsigned ref_rs_align_escrow(unsigned parity, unsigned block, const uint8_t* raw, unsigned sym_len, unsigned col_len,
                             unsigned chunk_size, unsigned max_chunks, uint8_t* chunks_out, unsigned* buffers_in_use)
{
	escrow_buffer_writer ebw(chunks_out, max_chunks, chunk_size);
	aligned_stream<escrow_buffer_writer> aligner(ebw, chunk_size, 0, nullptr);
	{
		reed_solomon_stream<aligned_stream<escrow_buffer_writer>> rss(aligner, parity, block);
		rss.write(reinterpret_cast<const char*>(raw), sym_len);
	}
	if (col_len)
	{
		reed_solomon_stream<aligned_stream<escrow_buffer_writer>> rss(aligner, parity, block);
		rss.write(reinterpret_cast<const char*>(raw + sym_len), col_len);
	}
	*buffers_in_use = ebw.buffers_in_use();
	return (unsigned)aligner.tellp();
}

// reed_solomon_stream into a std::stringstream (Decoder::decode semantics: failed block -> zeros)
unsigned ref_rs_stream(unsigned parity, unsigned block, const uint8_t* raw, unsigned len, uint8_t* out)
{
	std::stringstream ss;
	reed_solomon_stream<std::stringstream> rss(ss, parity, block);
	rss.write(reinterpret_cast<const char*>(raw), len);
	std::string s = ss.str();
	std::memcpy(out, s.data(), s.size());
	return (unsigned)s.size();
}

void ref_md_pack(uint8_t encode_id, unsigned size, uint16_t block_id, uint8_t* out6)
{
	FountainMetadata md(encode_id, size, block_id);
	std::memcpy(out6, md.data(), 6);
}
unsigned ref_md_file_size(const uint8_t* md6) { FountainMetadata md(reinterpret_cast<const char*>(md6), 6); return md.file_size(); }
unsigned ref_md_block_id(const uint8_t* md6) { FountainMetadata md(reinterpret_cast<const char*>(md6), 6); return md.block_id(); }
unsigned ref_md_encode_id(const uint8_t* md6) { FountainMetadata md(reinterpret_cast<const char*>(md6), 6); return md.encode_id(); }

// extractor/ScanState.h: the state machine fed with a pixel sequence; res[i] = process(active[i]), res[n] = the closing process(false)
void ref_scanstate_run(int kind, const uint8_t* active, int n, int* res)
{
	if (kind == 122) { ScanState_122 s; for (int i = 0; i < n; ++i) res[i] = s.process(active[i] != 0); res[n] = s.process(false); }
	else { ScanState_114 s; for (int i = 0; i < n; ++i) res[i] = s.process(active[i] != 0); res[n] = s.process(false); }
}

// extractor/Anchor.h: out = {xavg, yavg, xrange, yrange, max_range, size, is_mergeable(b, max_distance)}, then a.merge(b) -> a
void ref_anchor_ops(int* a, const int* b, int max_distance, long long* out)
{
	Anchor A(a[0], a[1], a[2], a[3]), B(b[0], b[1], b[2], b[3]);
	out[0] = A.xavg(); out[1] = A.yavg(); out[2] = A.xrange(); out[3] = A.yrange(); out[4] = A.max_range();
	out[5] = (long long)A.size(); out[6] = A.max_range() != 0 ? A.is_mergeable(B, max_distance) : -1;
	A.merge(B);
	a[0] = A.x(); a[1] = A.xmax(); a[2] = A.y(); a[3] = A.ymax();
}

} // extern "C"
