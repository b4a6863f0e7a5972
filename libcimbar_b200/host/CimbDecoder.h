// CimbDecoder.h -- mirror of libcimbar's CimbDecoder (reference: src/lib/cimb_translator/CimbDecoder.h:14-50,
// CimbDecoder.cpp:57-217).  Same method names, argument order and meaning; the work is done by the device kernels through
// cb200_decode_symbols / cb200_best_colors.  Per-call launches are for API parity and for re-running the reference's unit
// tests -- the fast path is Decoder (whole frames, batched).  The colour correction matrix is the calling THREAD's, as in the
// reference (CimbDecoder.cpp:69-85): every CimbDecoder / Decoder / CimbReader mirror of a thread sees the same one.
#pragma once
#include "../../include/cb200.h"
#include "Config.h"
#include "detail.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>

namespace cb200 {

// image_hash::ahash_result<8> holds the 10x10 pre-thresholded window a cell is decoded from and hands out the 8x8 hashes at
// the nine drift offsets (src/lib/image_hash/ahash_result.h:17-106); bitmatrix(grayscale, w, h, x-1, y-1) is the same window
// seen as a matrix (bit_file/bitmatrix.h:48-64).  Here both are the ten row words: rows[i] = the 10 bits of row i, bit 9 = the
// leftmost column.
struct CellWindow
{
	uint16_t rows[10];
};
using ahash_result = CellWindow;
using bitmatrix = CellWindow;

// chromatic_adaptation/color_correction.h:8-68 as far as the decode path uses it: an optional 3x3 float matrix
struct color_correction
{
	bool active = false;
	float mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	explicit operator bool() const { return active; }
};

class CimbDecoder
{
public:
	typedef unsigned char uchar;

	CimbDecoder(unsigned symbol_bits, unsigned color_bits, bool dark = true, uchar ahashThreshold = 0, int device = -1)
		: _symbolBits(symbol_bits), _colorBits(color_bits), _ahashThreshold(ahashThreshold), _device(device)
	{
		(void)dark;
		// the palette depends on (num_colors, color_mode): pick the context mode that carries it
		_modeVal = (color_bits == 3) ? 8 : cimbar::Config::mode_val();
		ctx();
	}

	// CimbDecoder::get_ccm / update_color_correction (CimbDecoder.cpp:76-85): the thread's matrix
	const color_correction& get_ccm() const
	{
		const detail::ThreadCcm& t = detail::thread_ccm();
		_ccmView.active = t.active;
		std::memcpy(_ccmView.mat, t.m, sizeof(t.m));
		return _ccmView;
	}
	bool get_ccm(float m9[9]) const
	{
		const detail::ThreadCcm& t = detail::thread_ccm();
		if (t.active) std::memcpy(m9, t.m, sizeof(t.m));
		return t.active;
	}
	void update_color_correction(const float m9[9])
	{
		detail::ThreadCcm& t = detail::thread_ccm();
		t.active = true;
		std::memcpy(t.m, m9, sizeof(t.m));
		++t.version;
	}
	template <typename MATX, typename = decltype(std::declval<MATX>().val)>     // cv::Matx<float,3,3>&& (row-major .val[9])
	void update_color_correction(MATX&& ccm) { update_color_correction(static_cast<const float*>(ccm.val)); }
	void clear_color_correction() { detail::thread_ccm().active = false; ++detail::thread_ccm().version; }       // TestableCimbDecoder: internal_ccm() = color_correction()

	// CimbDecoder::get_best_symbol(ahash_result&, drift_offset, best_distance, cooldown) -- CimbDecoder.cpp:101-132
	unsigned get_best_symbol(const ahash_result& results, unsigned& drift_offset, unsigned& best_distance, unsigned cooldown = 0xFF) const
	{
		uint8_t cd = (uint8_t)cooldown, sym = 0, off = 0, dist = 0;
		if (cb200_decode_symbols(ctx(), results.rows, &cd, 1, &sym, &off, &dist) != CB200_OK)
			throw std::runtime_error(std::string("cb200_decode_symbols: ") + cb200_last_error());
		drift_offset = off;
		best_distance = dist;
		return sym;
	}

	// CimbDecoder::decode_symbol(const bitmatrix&, drift_offset, best_distance, cooldown) -- CimbDecoder.cpp:142-147
	unsigned decode_symbol(const bitmatrix& cell, unsigned& drift_offset, unsigned& best_distance, unsigned cooldown = 0xFF) const
	{
		return get_best_symbol(cell, drift_offset, best_distance, cooldown);
	}

	// CimbDecoder::get_color(i, color_mode) -- CimbDecoder.cpp:149-152
	std::tuple<uchar, uchar, uchar> get_color(int i, unsigned color_mode) const
	{
		uint8_t rgb[3];
		if (cb200_palette_color((int)_colorBits, color_mode, i, rgb) != CB200_OK)
			throw std::invalid_argument(std::string("cb200_palette_color: ") + cb200_last_error());
		return {rgb[0], rgb[1], rgb[2]};
	}

	// CimbDecoder::get_best_color(float r, float g, float b, unsigned color_mode) -- CimbDecoder.cpp:168-200.
	// Inputs are the integer channel means the reference passes (avg_color returns uchar); color_mode must be the
	// context's (Config::color_mode()).
	unsigned get_best_color(float r, float g, float b, unsigned color_mode) const
	{
		if (color_mode != cimbar::Config::color_mode() and _colorBits != 3)
			throw std::invalid_argument("cb200::CimbDecoder: color_mode must match the active Config");
		uint8_t rgb[3] = {(uint8_t)r, (uint8_t)g, (uint8_t)b}, out = 0;
		cb200_ctx* c = ctx();
		detail::push_ccm(c);
		if (cb200_best_colors(c, rgb, 1, &out) != CB200_OK)
			throw std::runtime_error(std::string("cb200_best_colors: ") + cb200_last_error());
		return out;
	}

	// Cell::mean_rgb_continuous over the inner (n-2)x(n-2) of an n x n RGB8 cell at (x, y) -- Cell.h:30-62, CimbDecoder.cpp:202-209
	static std::tuple<uchar, uchar, uchar> avg_color(const unsigned char* rgb, int img_cols, int x, int y, int cell_size = 8)
	{
		uint16_t r = 0, g = 0, b = 0, count = 0;
		for (int i = 1; i < cell_size - 1; ++i)
		{
			const unsigned char* p = rgb + ((size_t)(y + i) * img_cols + (x + 1)) * 3;
			for (int j = 1; j < cell_size - 1; ++j, ++count, p += 3) { r += p[0]; g += p[1]; b += p[2]; }
		}
		if (!count) return {0, 0, 0};
		return {(uchar)(r / count), (uchar)(g / count), (uchar)(b / count)};
	}

	unsigned decode_color(const unsigned char* rgb, int img_cols, int x, int y, unsigned color_mode) const
	{
		if ((1u << _colorBits) <= 1) return 0;
		auto [r, g, b] = avg_color(rgb, img_cols, x, y);
		return get_best_color(r, g, b, color_mode);
	}

	bool expects_binary_threshold() const { return _ahashThreshold >= 0xFE; }
	unsigned symbol_bits() const { return _symbolBits; }
	unsigned color_bits() const { return _colorBits; }
	int device() const { return _device; }

protected:
	cb200_ctx* ctx() const { return detail::thread_context(_device, _modeVal, 1); }

	unsigned _symbolBits, _colorBits;
	uchar _ahashThreshold;
	int _device, _modeVal;
	mutable color_correction _ccmView;
};

}  // namespace cb200
