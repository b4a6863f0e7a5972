// Decoder.h -- header-compatible mirror of libcimbar's Decoder (reference: src/lib/encoder/Decoder.h:16-189).
// Same constructor and the same two templates
//     unsigned decode(const MAT& img, STREAM& ostream, bool should_preprocess=false, int color_correction=2)
//     unsigned decode_fountain(const MAT& img, FOUNTAINSTREAM& ostream, bool should_preprocess=false, int color_correction=2)
// with the same stream protocol (write() of good RS blocks, `stream << BadChunk(n)` for failed ones) and the same
// return value (good bytes).  The frame is decoded on the GPU through the C ABI (cb200_decode / cb200_decode_raw);
// only the replay of the block results into the caller's stream runs on the host.
//
// MAT: anything with .rows, .cols, .channels(), .data, .isContinuous() -- a cv::Mat works, so does cb200::Image.
// Differences from the reference, by design of this round:
//   * frames must be exactly Config::image_size_x() x image_size_y(), RGB8, continuous (the Extractor's output);
//     a smaller image reproduces the reference's degenerate "reader not good" result (zero-filled streams);
//   * color_correction: 0 = off, 1 = simpleColorCorrection (per-frame von Kries matrix), 2 = the header fit of
//     CimbReader::init_ccm -- all on the device and bit-exact.  As in the reference the fit only happens in
//     decode_fountain (its aligned_stream hands the decoder the fountain headers); decode() with 2 uses whatever CCM the
//     decoder holds.  Any other value is decoded as 0 and reported through last_warnings().
#pragma once
#include "../../include/cb200.h"
#include "Config.h"
#include "streams.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace cb200 {

struct Image   // minimal cv::Mat-shaped view over caller-owned RGB8 pixels
{
	int rows = 0, cols = 0;
	const unsigned char* data = nullptr;
	int nchannels = 3;
	int channels() const { return nchannels; }
	bool isContinuous() const { return true; }
};

class Decoder
{
public:
	static const unsigned WARN_COLOR_CORRECTION_IGNORED = 1;

	Decoder(bool use_ecc = true, bool interleave = true, int device = -1)
		: _useEcc(use_ecc), _interleave(interleave), _modeVal(cimbar::Config::mode_val())
	{
		if (!interleave) throw std::invalid_argument("cb200::Decoder: interleave=false is not supported on the device path");
		if (cb200_create(&_ctx, device, _modeVal, 1) != CB200_OK)
			throw std::runtime_error(std::string("cb200_create: ") + cb200_last_error());
		cb200_get_info(_ctx, &_info);
	}
	~Decoder() { cb200_destroy(_ctx); }
	Decoder(const Decoder&) = delete;
	Decoder& operator=(const Decoder&) = delete;

	unsigned last_warnings() const { return _warnings; }
	unsigned last_frame_flags() const { return _frameFlags; }   // CB200_FRAME_FALLBACK: the exact flood walk was needed

	// The decoder's colour correction matrix: CimbDecoder::update_color_correction / get_ccm (CimbDecoder.cpp:76-85), row-major.
	void update_color_correction(const float m9[9])
	{
		if (cb200_set_ccm(_ctx, m9) != CB200_OK) throw std::runtime_error(std::string("cb200_set_ccm: ") + cb200_last_error());
	}
	void clear_color_correction() { cb200_set_ccm(_ctx, nullptr); }   // TestableCimbDecoder: internal_ccm() = color_correction()
	bool get_ccm(float m9[9]) const { return cb200_get_ccm(_ctx, m9) == 1; }
	// DecoderPlus::load_ccm / save_ccm (src/lib/encoder/DecoderPlus.h:31-55): the file is the 9 float32 of the matrix
	bool load_ccm(const std::string& filename)
	{
		FILE* f = fopen(filename.c_str(), "rb");
		if (!f) return false;
		float m[9];
		size_t got = fread(m, 1, sizeof(m), f);
		fclose(f);
		if (got < sizeof(m)) return false;
		update_color_correction(m);
		return true;
	}
	bool save_ccm(const std::string& filename) const
	{
		float m[9];
		if (!get_ccm(m)) return false;
		FILE* f = fopen(filename.c_str(), "wb");
		if (!f) return false;
		size_t put = fwrite(m, 1, sizeof(m), f);
		fclose(f);
		return put == sizeof(m);
	}

	template <typename MAT, typename STREAM>
	unsigned decode(const MAT& img, STREAM& ostream, bool should_preprocess = false, int color_correction = 2)
	{
		std::vector<uint8_t> data, ok;
		if (!run(img, should_preprocess, color_correction, data, ok))
			return 0;
		if (!_useEcc)   // reed_solomon_stream pass-through (reed_solomon_stream.h:58-62): the raw cell bits
		{
			ostream.write(reinterpret_cast<const char*>(data.data()), (unsigned)data.size());
			return (unsigned)ostream.tellp();
		}
		// two reed_solomon_streams (symbols, colours) flushing into the same ostream (Decoder.h:100-101, :115-117)
		const unsigned msg = _info.ecc_block_size - _info.ecc_bytes;
		for (int b = 0; b < _info.rs_blocks; ++b)
		{
			if (ok[b]) ostream.write(reinterpret_cast<const char*>(data.data()) + (size_t)b * msg, msg);
			else ostream << BadChunk(msg);
		}
		return (unsigned)ostream.tellp();
	}

	template <typename MAT, typename FOUNTAINSTREAM>
	unsigned decode_fountain(const MAT& img, FOUNTAINSTREAM& ostream, bool should_preprocess = false, int color_correction = 2)
	{
		unsigned chunk_size = _info.chunk_size;
		if (ostream.chunk_size() != chunk_size)   // Decoder.h:180-185: eat the decode on a chunk-size mismatch
		{
			null_stream devnull;
			aligned_stream<null_stream> aligner(devnull, chunk_size, 0);
			HeaderCallbacks on(*this);
			return decode(img, aligner, should_preprocess, color_correction);
		}
		aligned_stream<FOUNTAINSTREAM> aligner(ostream, ostream.chunk_size(), 0);
		HeaderCallbacks on(*this);   // this stream feeds CimbReader::update_metadata in the reference: color_correction 2 fits
		return decode(img, aligner, should_preprocess, color_correction);
	}

protected:
	struct HeaderCallbacks   // scope guard: the decode in progress is decode_fountain's (Decoder.h:171-189)
	{
		Decoder& d;
		explicit HeaderCallbacks(Decoder& dec) : d(dec) { d._headerCallbacks = true; }
		~HeaderCallbacks() { d._headerCallbacks = false; }
	};

	template <typename MAT>
	bool run(const MAT& img, bool should_preprocess, int color_correction, std::vector<uint8_t>& data, std::vector<uint8_t>& ok)
	{
		_warnings = (color_correction < 0 or color_correction > 2) ? WARN_COLOR_CORRECTION_IGNORED : 0;
		_frameFlags = 0;
		if (cimbar::Config::mode_val() != _modeVal)
			throw std::runtime_error("cb200::Decoder: Config mode changed after construction (one Decoder per mode)");
		const bool good = img.cols >= _info.image_size_x and img.rows >= _info.image_size_y;   // CimbReader.cpp:119
		const size_t out_bytes = _useEcc ? (size_t)_info.data_bytes : (size_t)_info.raw_bytes;
		data.assign(out_bytes, 0);
		ok.assign(_info.rs_blocks, 1);
		if (!good or img.channels() != 3)
		{
			// reader not good: zero-filled symbol/colour buffers are still flushed (CimbReader.cpp:141-142, Decoder.h:100-117).
			// Every colorPositions entry is still the default {0,0,0}, so the colour pass ORs the colour decoded at pixel
			// (0,0) into bit position 0 (Decoder.h:107-114 / :153-158); after RS that single byte is corrected back to 0.
			if (!_useEcc and img.channels() == 3 and img.cols >= 8 and img.rows >= 8 and img.isContinuous())
			{
				unsigned r = 0, g = 0, b = 0;
				for (int i = 1; i <= 6; ++i)
					for (int j = 1; j <= 6; ++j)
					{
						const unsigned char* p = img.data + ((size_t)i * img.cols + j) * 3;
						r += p[0]; g += p[1]; b += p[2];
					}
				uint8_t mean[3] = {(uint8_t)(r / 36), (uint8_t)(g / 36), (uint8_t)(b / 36)}, col = 0;
				if (cb200_best_colors(_ctx, mean, 1, &col) != CB200_OK)
					throw std::runtime_error(std::string("cb200_best_colors: ") + cb200_last_error());
				size_t at = _info.legacy_mode ? 0 : (size_t)_info.raw_symbol_bytes;
				data[at] |= (uint8_t)(col << (8 - _info.color_bits));
			}
			return true;
		}
		if (img.cols != _info.image_size_x or img.rows != _info.image_size_y or !img.isContinuous())
			throw std::invalid_argument("cb200::Decoder: frame must be exactly image_size_x x image_size_y, continuous RGB8");
		uint32_t flags = (should_preprocess ? CB200_FLAG_SHARPEN : 0) | (color_correction == 1 ? CB200_FLAG_CC_SIMPLE : 0) |
		                 ((color_correction == 2 and _headerCallbacks) ? CB200_FLAG_CC_FIT : 0);
		uint8_t ff = 0;
		int rc = _useEcc ? cb200_decode(_ctx, img.data, 1, flags, data.data(), ok.data(), &ff)
		                 : cb200_decode_raw(_ctx, img.data, 1, flags, data.data(), &ff);
		if (rc != CB200_OK)
			throw std::runtime_error(std::string("cb200 decode: ") + cb200_last_error());
		_frameFlags = ff;
		return true;
	}

protected:
	bool _useEcc;
	bool _interleave;
	int _modeVal;
	cb200_ctx* _ctx = nullptr;
	bool _headerCallbacks = false;
	cb200_info _info;
	unsigned _warnings = 0;
	unsigned _frameFlags = 0;
};

}  // namespace cb200
