// Extractor.h -- mirrors of libcimbar's Scanner (its scan() result), Anchor and Extractor
// (reference: src/lib/extractor/Scanner.h:17-43, Scanner.cpp:182-199, Anchor.h:8-112, Extractor.h:11-46).
//   cb200::Scanner sc(img);  std::vector<cb200::Anchor> a = sc.scan();     // anchors in the reference's order, found on the GPU
//   cb200::Extractor ext;    int res = ext.extract(img, out);              // FAILURE / SUCCESS / NEEDS_SHARPEN, out = deskewed frame
// The scan (gray + Gaussian blur + Otsu + the t1..t4 line scans + corner ordering + the bottom-right search) and the deskew run
// on the device (csrc/scan.cu, csrc/deskew.cu); only `fast = true, dark = true, skip = 0` -- what Extractor::extract constructs
// (Extractor.h:33) -- is supported.  For camera pictures in / fountain chunks out without the frame ever leaving the device use
// cb200_scan_extract_decode_fountain.
#pragma once
#include "../../include/cb200.h"
#include "Config.h"
#include "Deskewer.h"
#include "detail.h"

#include <cstdlib>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cb200 {

class Anchor   // src/lib/extractor/Anchor.h:8-112 (the accessors the extractor's callers use)
{
public:
	Anchor() : Anchor(0, 0, 0, 0) {}
	Anchor(int x, int xmax, int y, int ymax) : _x(x), _xmax(xmax), _y(y), _ymax(ymax) {}
	Point center() const { Point p; p.x = xavg(); p.y = yavg(); return p; }
	int x() const { return _x; }
	int xmax() const { return _xmax; }
	int xavg() const { return (_x + _xmax) / 2; }
	int xrange() const { return std::abs(_x - _xmax) / 2; }
	int y() const { return _y; }
	int ymax() const { return _ymax; }
	int yavg() const { return (_y + _ymax) / 2; }
	int yrange() const { return std::abs(_y - _ymax) / 2; }
	int max_range() const { const int dx = std::abs(_x - _xmax), dy = std::abs(_y - _ymax); return dx > dy ? dx : dy; }
	unsigned long long size() const { const long long dx = _x - _xmax, dy = _y - _ymax; return (unsigned long long)(dx * dx + dy * dy); }
protected:
	int _x, _xmax, _y, _ymax;
};
inline std::ostream& operator<<(std::ostream& os, const Anchor& a)   // Anchor.h:107-111
{
	return os << a.xavg() << "+-" << a.xrange() << "," << a.yavg() << "+-" << a.yrange();
}

class Scanner
{
public:
	template <typename MAT>
	Scanner(const MAT& img, bool fast = true, bool dark = true, int skip = 0, int device = -1)
		: _data(img.data), _cols(img.cols), _rows(img.rows), _device(device)
	{
		if (!fast or !dark or skip != 0) throw std::invalid_argument("cb200::Scanner: only fast = true, dark = true, skip = 0 are supported");
		if (img.channels() != 3 or !img.isContinuous()) throw std::invalid_argument("cb200::Scanner: continuous RGB8 input expected");
	}

	// order: top-left, top-right, bottom-left, bottom-right; fewer than four when the picture does not show a code
	std::vector<Anchor> scan()
	{
		int32_t a[16], count = 0;
		cb200_ctx* c = detail::thread_context(_device, cimbar::Config::mode_val(), 1);
		if (cb200_scan(c, _data, _cols, _rows, 1, a, &count, &_cutoff) != CB200_OK)
			throw std::runtime_error(std::string("cb200_scan: ") + cb200_last_error());
		if (count < 0) throw std::runtime_error("cb200::Scanner: the picture produced more pattern hits than the device scan holds");
		std::vector<Anchor> out;
		for (int i = 0; i < count; ++i) out.emplace_back(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
		return out;
	}
	int anchor_size() const { return 30; }
	unsigned cutoff() const { return _cutoff; }          // filter_candidates' cutoff of the last scan (Scanner.cpp:83-105)

protected:
	const unsigned char* _data;
	int _cols, _rows, _device;
	uint32_t _cutoff = 0;
};

class Extractor
{
public:
	static constexpr int FAILURE = 0;
	static constexpr int SUCCESS = 1;
	static constexpr int NEEDS_SHARPEN = 2;

	Extractor(unsigned padding = 0, unsigned image_w = 0, unsigned image_h = 0, unsigned anchor_size = 0, int device = -1)
		: _deskewer(padding, image_w, image_h, anchor_size, device), _device(device) {}

	// Extractor::extract (Extractor.h:30-46); `out` may be a cb200::OwnedImage or anything assignable from one
	template <typename MAT, typename OUT>
	int extract(const MAT& img, OUT& out)
	{
		Scanner scanner(img, true, true, 0, _device);
		std::vector<Anchor> points = scanner.scan();
		if (points.size() < 4) return FAILURE;
		Corners corners(points[0].center(), points[1].center(), points[2].center(), points[3].center());
		out = _deskewer.deskew(img, corners);
		// Corners::is_granular_scale (Corners.h:57-75)
		const int W = (int)cimbar::Config::image_size_x(), H = (int)cimbar::Config::image_size_y();
		auto far_apart = [&](const Point& a, const Point& b) { return std::abs(a.x - b.x) > W or std::abs(a.y - b.y) > H; };
		if (!(far_apart(corners.top_left(), corners.top_right()) and far_apart(corners.top_right(), corners.bottom_right()) and
		      far_apart(corners.bottom_right(), corners.bottom_left()) and far_apart(corners.bottom_left(), corners.top_left())))
			return NEEDS_SHARPEN;
		return SUCCESS;
	}

protected:
	Deskewer _deskewer;
	int _device;
};

}  // namespace cb200
