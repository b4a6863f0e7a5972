// CimbReader.h -- mirror of libcimbar's CimbReader (reference: src/lib/cimb_translator/CimbReader.h:13-41,
// CimbReader.cpp:107-162).  The reference preprocesses in the constructor and then walks cells one read() at a time;
// here the constructor runs the whole exact flood walk on the GPU (cb200_decode_cells) and read()/read_color() replay
// its per-cell trace in the same order, with the same PositionData values.
#pragma once
#include "../../include/cb200.h"
#include "Config.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace cb200 {

struct PositionData   // src/lib/cimb_translator/PositionData.h:4-9
{
	unsigned i = 0;
	int x = 0;
	int y = 0;
};

class CimbReader
{
public:
	// img: rows x cols RGB8, continuous.  color_mode must be Config::color_mode(); color_correction must be 0.
	template <typename MAT>
	CimbReader(const MAT& img, unsigned color_mode, bool needs_sharpen = false, int color_correction = 0, int device = -1)
		: _good(img.cols >= (int)cimbar::Config::image_size_x() and img.rows >= (int)cimbar::Config::image_size_y())
	{
		(void)color_mode; (void)color_correction;
		cb200_ctx* ctx = nullptr;
		if (cb200_create(&ctx, device, cimbar::Config::mode_val(), 1) != CB200_OK)
			throw std::runtime_error(std::string("cb200_create: ") + cb200_last_error());
		cb200_info info;
		cb200_get_info(ctx, &info);
		_symbolBits = info.symbol_bits;
		_cells.resize(info.total_cells);
		_trace.resize(info.total_cells);
		if (_good)
		{
			if (img.cols != info.image_size_x or img.rows != info.image_size_y)
			{
				cb200_destroy(ctx);
				throw std::invalid_argument("cb200::CimbReader: frame must be exactly image_size_x x image_size_y");
			}
			int rc = cb200_decode_cells(ctx, img.data, 1, needs_sharpen ? CB200_FLAG_SHARPEN : 0, _cells.data(), _trace.data());
			if (rc != CB200_OK)
			{
				std::string err = cb200_last_error();
				cb200_destroy(ctx);
				throw std::runtime_error("cb200_decode_cells: " + err);
			}
			_order.resize(info.total_cells);
			for (unsigned i = 0; i < _trace.size(); ++i)
			{
				_order[_trace[i].order] = i;
				_byPos[key(_trace[i].x, _trace[i].y)] = i;
			}
		}
		cb200_destroy(ctx);
	}

	unsigned read(PositionData& pos)
	{
		if (done()) return 0;
		unsigned i = _order[_next++];
		pos.i = i;
		pos.x = _trace[i].x;
		pos.y = _trace[i].y;
		return _cells[i] & ((1u << _symbolBits) - 1u);
	}

	// colour of the cell whose (drift-adjusted) position read() reported
	unsigned read_color(const PositionData& pos) const
	{
		auto it = _byPos.find(key(pos.x, pos.y));
		if (it == _byPos.end())
			throw std::invalid_argument("cb200::CimbReader::read_color: position was not produced by read()");
		return (_cells[it->second] & 0x7Fu) >> _symbolBits;
	}

	// walk details of the cell returned by the last read(): drift_offset (4 = centre) and Hamming distance
	const cb200_cell_trace& trace(unsigned cell) const { return _trace[cell]; }

	bool done() const { return !_good or _next >= _order.size(); }
	unsigned num_reads() const { return (unsigned)_cells.size(); }

protected:
	static uint32_t key(int x, int y) { return ((uint32_t)(uint16_t)x << 16) | (uint16_t)y; }

	bool _good;
	unsigned _symbolBits = 4;
	size_t _next = 0;
	std::vector<uint8_t> _cells;
	std::vector<cb200_cell_trace> _trace;
	std::vector<unsigned> _order;
	std::unordered_map<uint32_t, unsigned> _byPos;
};

}  // namespace cb200
