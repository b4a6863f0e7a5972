// CimbReader.h -- mirror of libcimbar's CimbReader (reference: src/lib/cimb_translator/CimbReader.h:13-41,
// CimbReader.cpp:107-280).  Same constructor, same methods:
//     CimbReader(img, CimbDecoder& decoder, color_mode, needs_sharpen = false, color_correction = 2)
//     unsigned read(PositionData&)   unsigned read_color(const PositionData&) const   bool done() const
//     void init_ccm(color_bits, interleave_blocks, interleave_partitions, fountain_blocks)
//     void update_metadata(char* buff, unsigned len, unsigned chunk_size)             unsigned num_reads() const
// The reference preprocesses in the constructor and then walks cells one read() at a time; here the constructor runs the
// whole exact flood walk on the GPU (cb200_decode_cells_means) and read() replays its per-cell trace in the same order with the
// same PositionData values.  Colours are decided when they are first asked for, from the cells' mean colours and the CCM the
// decoder holds AT THAT MOMENT -- i.e. after init_ccm, as in Decoder::do_decode (Decoder.h:104-117).
#pragma once
#include "../../include/cb200.h"
#include "CimbDecoder.h"
#include "Config.h"
#include "detail.h"

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
	// img: rows x cols RGB8, continuous (cv::Mat works); the caller keeps it alive while the reader is used (the reference holds
	// a cv::Mat header, no deep copy: CimbReader.cpp:108).  color_mode must be Config::color_mode().
	template <typename MAT>
	CimbReader(const MAT& img, CimbDecoder& decoder, unsigned color_mode, bool needs_sharpen = false, int color_correction = 2)
		: _decoder(decoder)
		, _good(img.cols >= (int)cimbar::Config::image_size_x() and img.rows >= (int)cimbar::Config::image_size_y())
		, _colorCorrection(color_correction)
		, _colorMode(color_mode)
		, _pixels(img.data)
		, _modeVal(cimbar::Config::mode_val())
	{
		cb200_info info;
		cb200_mode_info(_modeVal, &info);
		_symbolBits = info.symbol_bits;
		_colorBits = info.color_bits;
		_cells.resize(info.total_cells);
		_trace.resize(info.total_cells);
		_means.resize(info.total_cells);
		if (!_good) return;
		if (img.cols != info.image_size_x or img.rows != info.image_size_y)
			throw std::invalid_argument("cb200::CimbReader: frame must be exactly image_size_x x image_size_y");
		cb200_ctx* c = ctx();
		detail::push_ccm(c);
		// color_correction == 1: simpleColorCorrection replaces the decoder's CCM before anything is read (CimbReader.cpp:124-125)
		uint32_t flags = (needs_sharpen ? CB200_FLAG_SHARPEN : 0u) | (color_correction == 1 ? CB200_FLAG_CC_SIMPLE : 0u);
		if (cb200_decode_cells_means(c, img.data, 1, flags, _cells.data(), _trace.data(), _means.data()) != CB200_OK)
			throw std::runtime_error(std::string("cb200_decode_cells_means: ") + cb200_last_error());
		detail::pull_ccm(c);
		_order.resize(info.total_cells);
		for (unsigned i = 0; i < _trace.size(); ++i)
		{
			_order[_trace[i].order] = i;
			_byPos[key(_trace[i].x, _trace[i].y)] = i;
		}
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

	// colour of the cell whose (drift-adjusted) position read() reported: CimbDecoder::decode_color of the 8x8 at (x, y)
	unsigned read_color(const PositionData& pos) const
	{
		auto it = _byPos.find(key(pos.x, pos.y));
		if (it == _byPos.end())
			throw std::invalid_argument("cb200::CimbReader::read_color: position was not produced by read()");
		classify();
		return _colors[it->second];
	}

	bool done() const { return !_good or _next >= _order.size(); }
	unsigned num_reads() const { return (unsigned)_cells.size(); }

	// CimbReader::update_metadata (CimbReader.cpp:269-280): the aligned_stream callback of Decoder::decode_fountain hands
	// every flushed chunk (or nullptr, 0 for a dropped one) to the reader, which keeps the fountain header "always +1"
	void update_metadata(char* buff, unsigned len, unsigned chunk_size)
	{
		if (len == 0 and header_id() == 0) return;
		if (header_id() == 0)
			for (unsigned k = 0; k < 6; ++k) _header[k] = k < len ? (uint8_t)buff[k] : 0;
		if (_radioactiveBlockId == 0)
		{   // computeRadioactiveBlockId (CimbReader.cpp:99-104)
			unsigned fs = header_file_size();
			_radioactiveBlockId = (fs % chunk_size == 0) ? 0xFFFFFFFFu : fs / chunk_size;
		}
		unsigned next = (((unsigned)_header[4] << 8) | _header[5]) + 1;      // FountainMetadata::increment_block_id
		if (next == _radioactiveBlockId) next += 1;
		_header[4] = (uint8_t)((next >> 8) & 0xFF);
		_header[5] = (uint8_t)(next & 0xFF);
	}

	// CimbReader::init_ccm (CimbReader.cpp:169-267): with color_correction == 2 and a header seen, fit the CCM on the device
	// and hand it to the decoder (== the calling thread's CCM)
	void init_ccm(unsigned color_bits, unsigned interleave_blocks, unsigned interleave_partitions, unsigned fountain_blocks)
	{
		if (_colorCorrection != 2) return;
		if (header_id() == 0) return;
		if (!_good) return;
		if (color_bits != _colorBits or interleave_partitions != cimbar::Config::interleave_partitions() or
		    fountain_blocks != cimbar::Config::fountain_chunks_per_frame() or
		    (interleave_blocks != 0 and interleave_blocks != cimbar::Config::interleave_blocks()))
			throw std::invalid_argument("cb200::CimbReader::init_ccm: arguments must be the active Config's");
		cb200_ctx* c = ctx();
		float m9[9];
		int rc = cb200_fit_ccm(c, _pixels, _header, _radioactiveBlockId, interleave_blocks == 0 ? CB200_FLAG_NO_INTERLEAVE : 0u, m9);
		if (rc < 0) throw std::runtime_error(std::string("cb200_fit_ccm: ") + cb200_last_error());
		// the reference advances the header once per colour-stream chunk while sampling (CimbReader.cpp:212-235)
		const unsigned col_chunks = colour_stream_chunks();
		for (unsigned k = 0; k < col_chunks; ++k)
		{
			unsigned next = (((unsigned)_header[4] << 8) | _header[5]) + 1;
			if (next == _radioactiveBlockId) next += 1;
			_header[4] = (uint8_t)((next >> 8) & 0xFF);
			_header[5] = (uint8_t)(next & 0xFF);
		}
		if (rc == 1) _decoder.update_color_correction(m9);    // colours asked for from now on see the new matrix
	}

	// walk details of a cell: drift_offset (4 = centre) and Hamming distance of its decode
	const cb200_cell_trace& trace(unsigned cell) const { return _trace[cell]; }

protected:
	static uint32_t key(int x, int y) { return ((uint32_t)(uint16_t)x << 16) | (uint16_t)y; }
	cb200_ctx* ctx() const { return detail::thread_context(_decoder.device(), _modeVal, 1); }
	uint32_t header_id() const { return (uint32_t)_header[0] | ((uint32_t)_header[1] << 8) | ((uint32_t)_header[2] << 16) | ((uint32_t)_header[3] << 24); }
	unsigned header_file_size() const
	{   // FountainMetadata::file_size (FountainMetadata.h:74-82)
		return (unsigned)_header[3] | ((unsigned)_header[2] << 8) | ((unsigned)_header[1] << 16) | (((unsigned)_header[0] & 0x80u) << 17);
	}
	unsigned colour_stream_chunks() const
	{
		const unsigned cb = _colorBits;
		if (!cb) return 0;
		const unsigned end = cimbar::Config::capacity(cb) * 8 / cb;
		const unsigned interval = cimbar::Config::capacity(_symbolBits + cb) * 8 / cimbar::Config::fountain_chunks_per_frame() / cb;
		return (end + interval - 1) / interval;
	}

	// CimbDecoder::decode_color for every cell, once per CCM: get_best_color of the stored inner-6x6 means
	void classify() const
	{
		if (_classified and _ccmVersion == detail::thread_ccm().version) return;
		_ccmVersion = detail::thread_ccm().version;
		_colors.assign(_cells.size(), 0);
		if (_good and _colorBits > 0)
		{
			std::vector<uint8_t> rgb(_means.size() * 3);
			for (size_t i = 0; i < _means.size(); ++i)
			{
				rgb[3 * i] = (uint8_t)(_means[i] & 0xFF); rgb[3 * i + 1] = (uint8_t)((_means[i] >> 8) & 0xFF); rgb[3 * i + 2] = (uint8_t)((_means[i] >> 16) & 0xFF);
			}
			cb200_ctx* c = ctx();
			detail::push_ccm(c);
			if (cb200_best_colors(c, rgb.data(), (int)_means.size(), _colors.data()) != CB200_OK)
				throw std::runtime_error(std::string("cb200_best_colors: ") + cb200_last_error());
		}
		_classified = true;
	}

	CimbDecoder& _decoder;
	bool _good;
	int _colorCorrection;
	unsigned _colorMode;
	const unsigned char* _pixels;
	int _modeVal;
	unsigned _symbolBits = 4, _colorBits = 2;
	size_t _next = 0;
	uint8_t _header[6] = {0, 0, 0, 0, 0, 0};       // FountainMetadata _fountainColorHeader
	unsigned _radioactiveBlockId = 0;
	std::vector<uint8_t> _cells;
	std::vector<cb200_cell_trace> _trace;
	std::vector<uint32_t> _means;
	std::vector<unsigned> _order;
	std::unordered_map<uint32_t, unsigned> _byPos;
	mutable std::vector<uint8_t> _colors;
	mutable bool _classified = false;
	mutable unsigned _ccmVersion = 0;
};

}  // namespace cb200
