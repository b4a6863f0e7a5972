// shim_test.cpp -- the reference's own unit tests for the decode path, re-run against the header-compatible shims
// (libcimbar_b200/host) which route to the CUDA kernels through the C ABI.  Sources of the expectations:
//   src/lib/encoder/test/DecoderTest.cpp:26-106            (SHA-256 of decoded bytes; checked by the python driver)
//   src/lib/cimb_translator/test/CimbReaderTest.cpp:37-163  (first read; first 22 cells in flood order; 12400 reads)
//   src/lib/cimb_translator/test/CimbDecoderTest.cpp:49-131 (prethreshold decode of all 16 tiles; colour known answers)
// Usage: shim_test <mode> <frame.rgb> <out_prefix>   (frame.rgb = raw RGB8 of the mode's image size)
#include "../../libcimbar_b200/host/CimbDecoder.h"
#include "../../libcimbar_b200/host/CimbReader.h"
#include "../../libcimbar_b200/host/Decoder.h"
#include "../../libcimbar_b200/host/cb200_fountain.h"

// The stream classes are the caller's.  Where the reference checkout is on the include path (-I<ref>/src/lib, the CPU
// compile check in tests/test_abi_host.py) the reference's own escrow_buffer_writer and aligned_stream are used UNMODIFIED,
// which is the drop-in claim; on the GPU box (no reference) a local collector with the same concept stands in.
#if defined(CB200_WITH_REFERENCE_STREAMS)
#include "encoder/aligned_stream.h"
#include "encoder/escrow_buffer_writer.h"
#else
class escrow_buffer_writer   // test-local stand-in: one buffer slot per write of exactly chunk_size bytes
{
public:
	escrow_buffer_writer(unsigned char* space, unsigned slots, unsigned slot_bytes) : _space(space), _slots(slots), _slotBytes(slot_bytes) {}
	bool good() const { return _ok; }
	unsigned chunk_size() const { return _slotBytes; }
	long tellp() const { return (long)_used * _slotBytes; }
	unsigned buffers_in_use() const { return _used; }
	escrow_buffer_writer& write(const char* data, unsigned length)
	{
		_ok = _ok and length == _slotBytes and _used < _slots;
		if (_ok) { std::memcpy(_space + (size_t)_used * _slotBytes, data, length); ++_used; }
		return *this;
	}
private:
	unsigned char* _space; unsigned _slots, _slotBytes, _used = 0; bool _ok = true;
};
#endif

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

using namespace cb200;

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

static const unsigned long long TILES[16] = {
    0xfffefcf8f0e0c080ULL, 0x80c0e0f0f8fcfeffULL, 0xff7f3f1f0f070301ULL, 0x0103070f1f3f7fffULL, 0x181818ffff181818ULL, 0x66e7e70000e7e766ULL,
    0x3c7ee7c3c3e77e3cULL, 0x18183c3c7e7effffULL, 0xc0f0fcfffffcf0c0ULL, 0xfffcf00000f0fcffULL, 0xff3f0f00000f3fffULL, 0xe7e7e7e7c3c38181ULL,
    0x8181c3c3e7e7e7e7ULL, 0x0000c3e77e3c1800ULL, 0x0c1c387070381c0cULL, 0x1e1e38381c1c7878ULL};

int main(int argc, char** argv)
{
	if (argc < 4) { std::printf("usage: shim_test <mode> <frame.rgb> <out_prefix>\n"); return 2; }
	int mode = std::atoi(argv[1]);
	cimbar::Config::update(mode);
	std::ifstream f(argv[2], std::ios::binary);
	std::vector<unsigned char> pix((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	Image img;
	img.rows = cimbar::Config::image_size_y(); img.cols = cimbar::Config::image_size_x(); img.data = pix.data();
	CHECK(pix.size() == (size_t)img.rows * img.cols * 3);
	std::string prefix = argv[3];

	{   // DecoderTest/testDecode + testDecodeEcc: write the decoded bytes; the driver compares SHA-256 with the goldens
		Decoder dec(false);
		std::stringstream ss;
		unsigned n = dec.decode(img, ss, false, 0);
		CHECK(n == cimbar::Config::capacity());
		std::ofstream(prefix + ".raw", std::ios::binary) << ss.str();
		Decoder dece;
		std::stringstream sse;
		unsigned ne = dece.decode(img, sse, false, 0);
		CHECK(ne == cimbar::Config::capacity() * (cimbar::Config::ecc_block_size() - cimbar::Config::ecc_bytes()) / cimbar::Config::ecc_block_size());
		std::ofstream(prefix + ".ecc", std::ios::binary) << sse.str();
		// decode_fountain into an escrow writer, as cimbard_scan_extract_decode does (cimbar_recv_js.cpp:148-189)
		std::vector<unsigned char> bufspace(cimbar::Config::fountain_chunks_per_frame() * cimbar::Config::fountain_chunk_size());
		escrow_buffer_writer ebw(bufspace.data(), cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
		unsigned good = dece.decode_fountain(img, ebw, false, 0);
		CHECK(good == ebw.buffers_in_use() * cimbar::Config::fountain_chunk_size());
		std::ofstream(prefix + ".chunks", std::ios::binary).write(reinterpret_cast<const char*>(bufspace.data()), good);
	}
	{   // colour correction through the mirrors: decode_fountain's default color_correction = 2 (header fit), then 1; the
		// fitted matrix goes through DecoderPlus::save_ccm / load_ccm's file format (9 float32)
		std::vector<unsigned char> bufspace(cimbar::Config::fountain_chunks_per_frame() * cimbar::Config::fountain_chunk_size());
		{
			Decoder d2;
			d2.clear_color_correction();          // one fresh reference decoder (the tests before this one may have left a CCM)
			escrow_buffer_writer ebw(bufspace.data(), cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
			unsigned good = d2.decode_fountain(img, ebw);
			CHECK(d2.last_warnings() == 0);
			std::ofstream(prefix + ".chunks_cc2", std::ios::binary).write(reinterpret_cast<const char*>(bufspace.data()), good);
			float m9[9];
			if (d2.get_ccm(m9))
			{
				CHECK(d2.save_ccm(prefix + ".ccm"));
				Decoder d3;
				float back[9];
				// the CCM is the thread's (CimbDecoder.cpp:69-73): a fresh Decoder still sees what d2 fitted, until it is cleared
				CHECK(d3.get_ccm(back) && std::memcmp(back, m9, sizeof(m9)) == 0);
				d3.clear_color_correction();
				CHECK(!d3.get_ccm(back));
				CHECK(d3.load_ccm(prefix + ".ccm"));
				CHECK(d3.get_ccm(back) && std::memcmp(back, m9, sizeof(m9)) == 0);
				// Decoder::decode on a plain stream never fits (no header callbacks): it decodes with the loaded matrix
				std::stringstream plain;
				d3.decode(img, plain, false, 2);
				CHECK(d3.get_ccm(back) && std::memcmp(back, m9, sizeof(m9)) == 0);
			}
			else std::remove((prefix + ".ccm").c_str());
		}
		{
			Decoder d1;
			d1.clear_color_correction();
			escrow_buffer_writer ebw(bufspace.data(), cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
			unsigned good = d1.decode_fountain(img, ebw, false, 1);
			std::ofstream(prefix + ".chunks_cc1", std::ios::binary).write(reinterpret_cast<const char*>(bufspace.data()), good);
		}
	}
	{   // CimbReaderTest: first 22 cells in flood order as "index=value" pairs, then the reader runs to exactly 12400 reads
		CimbDecoder rdec(cimbar::Config::symbol_bits(), cimbar::Config::color_bits(), cimbar::Config::dark(), 0xFF);
		rdec.clear_color_correction();
		CimbReader cr(img, rdec, cimbar::Config::color_mode(), false, 0);     // CimbReaderTest.cpp:44 call shape
		CHECK(!cr.done());
		std::map<unsigned, unsigned> res;
		int count = 0;
		PositionData first;
		for (int c = 0; c < 22; ++c)
		{
			PositionData pos;
			unsigned bits = cr.read(pos);
			if (c == 0) first = pos;
			res[pos.i] = bits | (cr.read_color(pos) << 4);
			++count;
		}
		std::ostringstream os;
		bool sep = false;
		for (auto& kv : res) { os << (sep ? " " : "") << kv.first << "=" << kv.second; sep = true; }
		std::ofstream(prefix + ".first22") << os.str() << "\n" << first.i << " " << first.x << " " << first.y << "\n";
		PositionData pos;
		while (!cr.done()) { cr.read(pos); ++count; }
		CHECK(cr.done());
		CHECK(count == (int)cimbar::Config::total_cells());
	}
	{   // CimbDecoderTest/testPrethresholdDecode: every tile, centred in a 10x10 window -> (symbol i, drift 4, distance 0)
		CimbDecoder cd(4, 2, true, 0xFF);
		for (unsigned i = 0; i < 16; ++i)
		{
			CellWindow w;
			for (int r = 0; r < 10; ++r) w.rows[r] = 0;
			for (int r = 0; r < 8; ++r) w.rows[r + 1] = (uint16_t)(((TILES[i] >> (8 * (7 - r))) & 0xFF) << 1);
			unsigned off = 9, dist = 99;
			unsigned res = cd.decode_symbol(w, off, dist);
			CHECK(res == i); CHECK(off == 4); CHECK(dist == 0);
		}
		// CimbDecoderTest/test_get_best_color_mode1 (CimbDecoderTest.cpp:104-131), when the active palette is mode B's
		if (cimbar::Config::color_mode() == 1)
		{
			CHECK(cd.get_best_color(255, 0, 255, 1) == 3); CHECK(cd.get_best_color(255, 255, 0, 1) == 2);
			CHECK(cd.get_best_color(0, 255, 255, 1) == 1); CHECK(cd.get_best_color(0, 255, 0, 1) == 0);
			CHECK(cd.get_best_color(0, 0, 0, 1) == 0); CHECK(cd.get_best_color(70, 70, 70, 1) == 0);
			CHECK(cd.get_best_color(20, 200, 20, 1) == 0); CHECK(cd.get_best_color(200, 30, 200, 1) == 3);
			CHECK(cd.get_best_color(200, 155, 20, 1) == 2); CHECK(cd.get_best_color(50, 155, 200, 1) == 1);
		}
		else
		{   // test_get_best_color_mode0 (CimbDecoderTest.cpp:77-102)
			CHECK(cd.get_best_color(255, 0, 255, 0) == 2); CHECK(cd.get_best_color(255, 255, 0, 0) == 1);
			CHECK(cd.get_best_color(0, 255, 255, 0) == 0); CHECK(cd.get_best_color(0, 255, 0, 0) == 3);
			CHECK(cd.get_best_color(20, 200, 20, 0) == 3); CHECK(cd.get_best_color(155, 50, 155, 0) == 2);
		}
	}
	{   // Decoder::do_decode restated call by call on the CimbReader / CimbDecoder mirrors with color_correction 2
		// (Decoder.h:60-117 + :171-189): read() every cell, track the fountain header through update_metadata as the
		// aligned_stream callbacks would, init_ccm, then read_color() -- must give the colours of Decoder::decode_fountain
		if (!cimbar::Config::legacy_mode())
		{
			CimbDecoder cdec(cimbar::Config::symbol_bits(), cimbar::Config::color_bits(), cimbar::Config::dark(), 0xFF);
			cdec.clear_color_correction();
			CimbReader reader(img, cdec, cimbar::Config::color_mode(), false, 2);
			std::vector<unsigned> idx(cimbar::Config::total_cells());
			{
				std::vector<uint16_t> m(idx.size());
				cb200_interleave_indices(cimbar::Config::mode_val(), m.data());
				for (size_t slot = 0; slot < m.size(); ++slot) idx[m[slot]] = (unsigned)slot;       // Interleave::interleave_reverse
			}
			std::vector<PositionData> colorPositions(reader.num_reads());
			std::vector<uint8_t> symbols(reader.num_reads());
			while (!reader.done())
			{
				PositionData pos;
				unsigned bits = reader.read(pos);
				symbols[pos.i] = (uint8_t)bits;
				colorPositions[pos.i] = pos;
			}
			// the header events of the symbol stream: decode the frame once more through the Decoder mirror with ECC to get the
			// symbol-stream chunks (the RS stage is not part of the reader) and replay them as update_metadata calls
			Decoder plain;
			plain.clear_color_correction();
			std::vector<unsigned char> space(cimbar::Config::fountain_chunks_per_frame() * cimbar::Config::fountain_chunk_size());
			std::vector<uint8_t> data(cimbar::Config::capacity() * (cimbar::Config::ecc_block_size() - cimbar::Config::ecc_bytes()) / cimbar::Config::ecc_block_size());
			std::stringstream blocks;
			plain.decode(img, blocks, false, 0);
			std::string all = blocks.str();
			const unsigned cs = cimbar::Config::fountain_chunk_size();
			const unsigned sym_chunks = cimbar::Config::capacity(cimbar::Config::symbol_bits()) * (cimbar::Config::ecc_block_size() - cimbar::Config::ecc_bytes()) / cimbar::Config::ecc_block_size() / cs;
			bool clean = true;     // (only frames whose symbol stream decodes completely are replayed here: zero blocks would need the bad-chunk events)
			for (unsigned q = 0; q < sym_chunks; ++q)
			{
				bool zero = true;
				for (unsigned k = 0; k < cs; ++k) zero = zero and all[(size_t)q * cs + k] == 0;
				clean = clean and !zero;
			}
			if (clean)
			{
				for (unsigned q = 0; q < sym_chunks; ++q) reader.update_metadata(&all[(size_t)q * cs], cs, cs);
				reader.init_ccm(cimbar::Config::color_bits(), cimbar::Config::interleave_blocks(), cimbar::Config::interleave_partitions(),
				                cimbar::Config::fountain_chunks_per_frame());
				std::vector<uint8_t> colors(reader.num_reads());
				for (const PositionData& p : colorPositions) colors[p.i] = (uint8_t)reader.read_color(p);
				std::ofstream(prefix + ".cells_cc2", std::ios::binary).write(reinterpret_cast<const char*>(colors.data()), colors.size());
				float m9[9];
				if (cdec.get_ccm(m9)) std::ofstream(prefix + ".ccm_reader", std::ios::binary).write(reinterpret_cast<const char*>(m9), sizeof(m9));
			}
			cdec.clear_color_correction();
		}
	}
	{   // the thread's CCM survives the Decoder object (CimbDecoder.cpp:69-73): a fresh Decoder per frame, as in
		// cimbar_recv_js.cpp:164, still decodes the second frame with the matrix the first one fitted; and the batched
		// decode_fountain(span) gives the same chunks as frame-by-frame calls
		std::vector<unsigned char> a(cimbar::Config::fountain_chunks_per_frame() * cimbar::Config::fountain_chunk_size() * 2), b(a.size());
		unsigned good_a = 0, good_b = 0;
		float ma[9] = {0}, mb[9] = {0};
		bool has_a, has_b;
		{
			Decoder().clear_color_correction();
			escrow_buffer_writer w(a.data(), 2 * cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
			{ Decoder d; good_a += d.decode_fountain(img, w); }
			{ Decoder d; good_a += d.decode_fountain(img, w); has_a = d.get_ccm(ma); }
		}
		{
			Decoder().clear_color_correction();
			escrow_buffer_writer w(b.data(), 2 * cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
			Decoder d;
			Image two[2] = {img, img};
			good_b = d.decode_fountain(two, 2, w);
			has_b = d.get_ccm(mb);
		}
		CHECK(good_a == good_b);
		CHECK(std::memcmp(a.data(), b.data(), good_a) == 0);
		CHECK(has_a == has_b and (!has_a or std::memcmp(ma, mb, sizeof(ma)) == 0));
		Decoder().clear_color_correction();
	}
	{   // Decoder(use_ecc, interleave = false): cells in linear order (Interleave.h:10-16); written for the driver to compare
		Decoder lin(false, false);
		lin.clear_color_correction();
		std::stringstream ss;
		unsigned n = lin.decode(img, ss, false, 0);
		CHECK(n == cimbar::Config::capacity());
		std::ofstream(prefix + ".raw_nointerleave", std::ios::binary) << ss.str();
	}
	{   // a short write into the fountain sink is buffered until the rest of the chunk arrives (fountain_decoder_stream.h:52-74)
		struct Fake { static void* create(void*, uint64_t, uint32_t) { return new int(0); }
		              static int decode(void* c, unsigned, const void*, uint32_t) { return ++*static_cast<int*>(c) >= 2 ? 0 : 1; }
		              static int recover(void*, void* out, uint64_t n) { std::memset(out, 7, n); return 0; }
		              static void free_codec(void* c) { delete static_cast<int*>(c); } };
		FountainCodecApi api{Fake::create, Fake::decode, Fake::recover, Fake::free_codec};
		fountain_sink sink(625, api);
		std::vector<char> chunk(625, 1);
		FountainMetadata md(3, 1000, 0);
		std::memcpy(chunk.data(), md.d, 6);
		std::memcpy(chunk.data() + 300, md.d, 6);                          // decode_frame reads the stream id from the head of every call
		CHECK(sink.decode_frame(chunk.data(), 5) == -10);
		CHECK(sink.decode_frame(chunk.data(), 300) == 0);                 // kept, nothing decoded yet
		CHECK(sink.decode_frame(chunk.data() + 300, 325) == 0);           // (header of the continuation bytes must parse: same id) -> block 0 fed
		FountainMetadata md1(3, 1000, 1);
		std::memcpy(chunk.data(), md1.d, 6);
		int64_t done = sink.decode_frame(chunk.data(), 625);
		CHECK(done == (int64_t)md1.id());
		CHECK(sink.is_done(md1.id()) and sink.file(md1.id()) and sink.file(md1.id())->size() == 1000);
	}
	std::printf(fails ? "shim_test: %d failure(s)\n" : "shim_test: ok\n", fails);
	return fails ? 1 : 0;
}
