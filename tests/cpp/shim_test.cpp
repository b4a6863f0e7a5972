// shim_test.cpp -- the reference's own unit tests for the decode path, re-run against the header-compatible shims
// (libcimbar_b200/host) which route to the CUDA kernels through the C ABI.  Sources of the expectations:
//   src/lib/encoder/test/DecoderTest.cpp:26-106            (SHA-256 of decoded bytes; checked by the python driver)
//   src/lib/cimb_translator/test/CimbReaderTest.cpp:37-163  (first read; first 22 cells in flood order; 12400 reads)
//   src/lib/cimb_translator/test/CimbDecoderTest.cpp:49-131 (prethreshold decode of all 16 tiles; colour known answers)
// Usage: shim_test <mode> <frame.rgb> <out_prefix>   (frame.rgb = raw RGB8 of the mode's image size)
#include "../../libcimbar_b200/host/CimbDecoder.h"
#include "../../libcimbar_b200/host/CimbReader.h"
#include "../../libcimbar_b200/host/Decoder.h"

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
			escrow_buffer_writer ebw(bufspace.data(), cimbar::Config::fountain_chunks_per_frame(), cimbar::Config::fountain_chunk_size());
			unsigned good = d1.decode_fountain(img, ebw, false, 1);
			std::ofstream(prefix + ".chunks_cc1", std::ios::binary).write(reinterpret_cast<const char*>(bufspace.data()), good);
		}
	}
	{   // CimbReaderTest: first 22 cells in flood order as "index=value" pairs, then the reader runs to exactly 12400 reads
		CimbReader cr(img, cimbar::Config::color_mode());
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
	std::printf(fails ? "shim_test: %d failure(s)\n" : "shim_test: ok\n", fails);
	return fails ? 1 : 0;
}
