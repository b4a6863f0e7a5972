// CimbDecoder.h -- mirror of libcimbar's CimbDecoder (reference: src/lib/cimb_translator/CimbDecoder.h:14-50,
// CimbDecoder.cpp:57-217).  Same method names and argument meaning; the work is done by the device kernels through
// cb200_decode_symbols / cb200_best_colors.  Per-call launches are for API parity and for re-running the reference's
// unit tests -- the fast path is Decoder (whole frames, batched).
#pragma once
#include "../../include/cb200.h"
#include "Config.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>

namespace cb200 {

// the 10x10 pre-thresholded window a cell is decoded from == bitmatrix(grayscale, w, h, x-1, y-1) in the reference
// (bit_file/bitmatrix.h:48-64): rows[i] holds the 10 bits of row i, bit 9 = leftmost column
struct CellWindow
{
	uint16_t rows[10];
};

class CimbDecoder
{
public:
	CimbDecoder(unsigned symbol_bits, unsigned color_bits, bool dark = true, unsigned char ahashThreshold = 0, int device = -1)
		: _symbolBits(symbol_bits), _colorBits(color_bits), _ahashThreshold(ahashThreshold)
	{
		(void)dark;
		// the palette depends on (num_colors, color_mode): pick the context mode that carries it
		int mode_val = (color_bits == 3) ? 8 : cimbar::Config::mode_val();
		if (cb200_create(&_ctx, device, mode_val, 1) != CB200_OK)
			throw std::runtime_error(std::string("cb200_create: ") + cb200_last_error());
	}
	~CimbDecoder() { cb200_destroy(_ctx); }
	CimbDecoder(const CimbDecoder&) = delete;
	CimbDecoder& operator=(const CimbDecoder&) = delete;

	// CimbDecoder::decode_symbol(const bitmatrix&, drift_offset, best_distance, cooldown) -- CimbDecoder.cpp:142-147
	unsigned decode_symbol(const CellWindow& cell, unsigned& drift_offset, unsigned& best_distance, unsigned cooldown = 0xFF) const
	{
		uint8_t cd = (uint8_t)cooldown, sym = 0, off = 0, dist = 0;
		if (cb200_decode_symbols(_ctx, cell.rows, &cd, 1, &sym, &off, &dist) != CB200_OK)
			throw std::runtime_error(std::string("cb200_decode_symbols: ") + cb200_last_error());
		drift_offset = off;
		best_distance = dist;
		return sym;
	}

	// CimbDecoder::get_best_color(float r, float g, float b, unsigned color_mode) -- CimbDecoder.cpp:168-200.
	// Inputs are the integer channel means the reference passes (avg_color returns uchar); color_mode must be the
	// context's (Config::color_mode()).
	unsigned get_best_color(float r, float g, float b, unsigned color_mode) const
	{
		if (color_mode != cimbar::Config::color_mode() and _colorBits != 3)
			throw std::invalid_argument("cb200::CimbDecoder: color_mode must match the active Config");
		uint8_t rgb[3] = {(uint8_t)r, (uint8_t)g, (uint8_t)b}, out = 0;
		if (cb200_best_colors(_ctx, rgb, 1, &out) != CB200_OK)
			throw std::runtime_error(std::string("cb200_best_colors: ") + cb200_last_error());
		return out;
	}

	// Cell::mean_rgb_continuous over the inner (n-2)x(n-2) of an n x n RGB8 cell at (x, y) -- Cell.h:30-62, CimbDecoder.cpp:202-209
	static std::tuple<unsigned char, unsigned char, unsigned char> avg_color(const unsigned char* rgb, int img_cols, int x, int y, int cell_size = 8)
	{
		uint16_t r = 0, g = 0, b = 0, count = 0;
		for (int i = 1; i < cell_size - 1; ++i)
		{
			const unsigned char* p = rgb + ((size_t)(y + i) * img_cols + (x + 1)) * 3;
			for (int j = 1; j < cell_size - 1; ++j, ++count, p += 3) { r += p[0]; g += p[1]; b += p[2]; }
		}
		if (!count) return {0, 0, 0};
		return {(unsigned char)(r / count), (unsigned char)(g / count), (unsigned char)(b / count)};
	}

	unsigned decode_color(const unsigned char* rgb, int img_cols, int x, int y, unsigned color_mode) const
	{
		if ((1u << _colorBits) <= 1) return 0;
		auto [r, g, b] = avg_color(rgb, img_cols, x, y);
		return get_best_color(r, g, b, color_mode);
	}

	// CimbDecoder::update_color_correction / get_ccm (CimbDecoder.cpp:76-85): get_best_color and decode_color use it while set
	void update_color_correction(const float m9[9])
	{
		if (cb200_set_ccm(_ctx, m9) != CB200_OK) throw std::runtime_error(std::string("cb200_set_ccm: ") + cb200_last_error());
	}
	void clear_color_correction() { cb200_set_ccm(_ctx, nullptr); }
	bool get_ccm(float m9[9]) const { return cb200_get_ccm(_ctx, m9) == 1; }

	bool expects_binary_threshold() const { return _ahashThreshold >= 0xFE; }
	unsigned symbol_bits() const { return _symbolBits; }

protected:
	unsigned _symbolBits, _colorBits;
	unsigned char _ahashThreshold;
	cb200_ctx* _ctx = nullptr;
};

}  // namespace cb200
