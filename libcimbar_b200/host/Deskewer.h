// Deskewer.h -- mirror of libcimbar's Deskewer and Corners (reference: src/lib/extractor/Deskewer.h:11-40, Corners.h:9-60).
// Same constructor, same `deskew(img, corners)`: cv::getPerspectiveTransform(corners.all(), outputPoints) followed by
// cv::warpPerspective(..., cv::INTER_LINEAR) to the mode's image size -- both restated bit for bit and run on the GPU
// (cb200_perspective_transform / cb200_deskew, csrc/deskew.cu).  The anchor scan that produces the corners runs on the GPU too: host/Extractor.h
// (Scanner / Extractor mirrors over cb200_scan).  For the whole camera path in one call -- the picture goes to the device once, the
// deskewed frame never leaves it -- use cb200_scan_extract_decode_fountain (or cb200_extract_decode_fountain with corners of your own).
#pragma once
#include "../../include/cb200.h"
#include "Config.h"
#include "detail.h"

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cb200 {

struct Point { int x = 0, y = 0; };

class Corners   // src/lib/extractor/Corners.h:9-60: the four anchor centres, integer points
{
public:
	Corners(Point top_left, Point top_right, Point bottom_left, Point bottom_right)
		: _tl(top_left), _tr(top_right), _bl(bottom_left), _br(bottom_right) {}
	const Point& top_left() const { return _tl; }
	const Point& top_right() const { return _tr; }
	const Point& bottom_left() const { return _bl; }
	const Point& bottom_right() const { return _br; }
	// Corners::all(): top-left, top-right, bottom-left, bottom-right as cv::Point2f
	std::vector<float> all() const
	{
		return {(float)_tl.x, (float)_tl.y, (float)_tr.x, (float)_tr.y, (float)_bl.x, (float)_bl.y, (float)_br.x, (float)_br.y};
	}
protected:
	Point _tl, _tr, _bl, _br;
};

struct OwnedImage   // what deskew returns: a cv::Mat-shaped RGB8 frame that shares ownership of its pixels (copies stay valid)
{
	int rows = 0, cols = 0;
	std::shared_ptr<std::vector<unsigned char>> pixels;
	const unsigned char* data = nullptr;
	int channels() const { return 3; }
	bool isContinuous() const { return true; }
};

class Deskewer
{
public:
	// padding and a custom image / anchor size are not supported on the device path (the reference's callers on the decode path
	// pass the defaults: Extractor.h:33-41 with padding 0)
	Deskewer(unsigned padding = 0, unsigned image_w = 0, unsigned image_h = 0, unsigned anchor_size = 0, int device = -1)
		: _device(device)
	{
		if (padding != 0 or (image_w and image_w != cimbar::Config::image_size_x()) or (image_h and image_h != cimbar::Config::image_size_y()) or
		    (anchor_size and anchor_size != cimbar::Config::anchor_size()))
			throw std::invalid_argument("cb200::Deskewer: only padding 0 and the active Config's image / anchor size are supported");
	}

	template <typename MAT>
	OwnedImage deskew(const MAT& img, const Corners& corners)
	{
		if (img.channels() != 3 or !img.isContinuous()) throw std::invalid_argument("cb200::Deskewer: continuous RGB8 input expected");
		const float an = (float)cimbar::Config::anchor_size();
		const float W = (float)cimbar::Config::image_size_x(), H = (float)cimbar::Config::image_size_y();
		const float out_pts[8] = {an, an, W - an, an, an, H - an, W - an, H - an};       // Deskewer.h:28-32 with padding 0
		double m9[9];
		std::vector<float> in_pts = corners.all();
		if (cb200_perspective_transform(in_pts.data(), out_pts, m9) != CB200_OK)
			throw std::runtime_error(std::string("cb200_perspective_transform: ") + cb200_last_error());
		OwnedImage out;
		out.rows = (int)cimbar::Config::image_size_y(); out.cols = (int)cimbar::Config::image_size_x();
		out.pixels = std::make_shared<std::vector<unsigned char>>((size_t)out.rows * out.cols * 3);
		cb200_ctx* c = detail::thread_context(_device, cimbar::Config::mode_val(), 1);
		if (cb200_deskew(c, img.data, img.cols, img.rows, 1, m9, out.pixels->data()) != CB200_OK)
			throw std::runtime_error(std::string("cb200_deskew: ") + cb200_last_error());
		out.data = out.pixels->data();
		return out;
	}

protected:
	int _device;
};

}  // namespace cb200
