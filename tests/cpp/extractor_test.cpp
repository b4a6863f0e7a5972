// extractor_test.cpp -- the reference's ScannerTest/testExampleScan and ExtractorTest call shapes through the mirrors
// (libcimbar_b200/host/Extractor.h): Scanner sc(img); sc.scan();  Extractor ext; ext.extract(img, out);  then the facade's
// next step, Decoder::decode_fountain(out, ebw, should_preprocess = true) (cimbar_recv_js.cpp:164-186).
//   extractor_test <mode> <w> <h> <picture.rgb> <out_prefix>
// writes <prefix>.anchors (operator<< of the anchors, space separated), <prefix>.status, <prefix>.frame (the extracted RGB8 frame)
// and <prefix>.ecc (Decoder::decode of that frame with should_preprocess = true).
#include "../../libcimbar_b200/host/Decoder.h"
#include "../../libcimbar_b200/host/Extractor.h"

#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

using namespace cb200;

int main(int argc, char** argv)
{
	if (argc < 6) { std::printf("usage: extractor_test <mode> <w> <h> <picture.rgb> <out_prefix>\n"); return 2; }
	cimbar::Config::update(std::atoi(argv[1]));
	std::ifstream f(argv[4], std::ios::binary);
	std::vector<unsigned char> pix((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	Image img;
	img.cols = std::atoi(argv[2]); img.rows = std::atoi(argv[3]); img.data = pix.data();
	if (pix.size() != (size_t)img.rows * img.cols * 3) { std::printf("bad picture size\n"); return 2; }
	const std::string prefix = argv[5];

	Scanner sc(img);
	std::vector<Anchor> candidates = sc.scan();
	std::ostringstream os;
	for (size_t i = 0; i < candidates.size(); ++i) os << (i ? " " : "") << candidates[i];
	std::ofstream(prefix + ".anchors") << os.str();

	Extractor ext;
	OwnedImage out;
	const int res = ext.extract(img, out);
	std::ofstream(prefix + ".status") << res;
	if (res != Extractor::FAILURE)
	{
		std::ofstream(prefix + ".frame", std::ios::binary).write(reinterpret_cast<const char*>(out.data), (std::streamsize)out.rows * out.cols * 3);
		Decoder dec;
		dec.clear_color_correction();
		std::stringstream ss;
		dec.decode(out, ss, true, 0);
		std::ofstream(prefix + ".ecc", std::ios::binary) << ss.str();
	}
	return 0;
}
