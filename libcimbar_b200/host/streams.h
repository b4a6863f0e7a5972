// streams.h -- the STREAM concepts Decoder writes into, restated so the shims are self-contained.
// In an integrated build these are the reference's own headers (they are OpenCV-free and can be used as they are):
//   ReedSolomon::BadChunk            src/lib/encoder/ReedSolomon.h:12-20
//   aligned_stream<STREAM>           src/lib/encoder/aligned_stream.h:8-131
//   escrow_buffer_writer             src/lib/encoder/escrow_buffer_writer.h:9-69
//   null_stream                      src/lib/util/null_stream.h
#pragma once
#include <algorithm>
#include <cstddef>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

namespace cb200 {

struct BadChunk
{
	unsigned size;
	explicit BadChunk(unsigned size) : size(size) {}
};

// failed RS block into a plain stream: zeros (reed_solomon_stream.h:96-107)
inline std::stringstream& operator<<(std::stringstream& s, const BadChunk& chunk)
{
	std::string temp(chunk.size, '\0');
	s.write(temp.data(), temp.size());
	return s;
}
inline std::ofstream& operator<<(std::ofstream& os, const BadChunk& chunk)
{
	std::string temp(chunk.size, '\0');
	os.write(temp.data(), temp.size());
	return os;
}

class null_stream
{
public:
	bool good() const { return true; }
	long tellp() const { return _count; }
	null_stream& write(const char*, unsigned length) { _count += length; return *this; }
protected:
	long _count = 0;
};

template <typename STREAM>
class aligned_stream
{
public:
	aligned_stream(STREAM& stream, unsigned align_increment, unsigned align_offset = 0,
	               const std::function<void(char*, size_t)>& on_flush = nullptr)
		: _stream(stream), _buffer(align_increment, 0), _offset(0), _alignOffset(align_offset)
		, _alignIncrement(align_increment), _onFlush(on_flush), _badChunk(false), _good(true)
	{}

	bool good() const { return _good and _stream.good(); }
	long tellp() const { return _totalCount; }

	aligned_stream& write(const char* data, unsigned length)
	{
		if (!good())
			return *this;
		while (length > 0)
		{
			if (_offset < _alignOffset)
			{
				unsigned writeLen = std::min(_alignOffset - _offset, length);
				_stream.write(data, writeLen);
				_totalCount += writeLen;
				length -= writeLen; data += writeLen; _alignOffset -= writeLen;
				continue;
			}
			unsigned work = length + _offset;
			if (work >= _alignIncrement)
			{
				unsigned writeLen = _alignIncrement - _offset;
				if (_badChunk)
				{
					_badChunk = false;
					_offset = 0;
					if (_onFlush) _onFlush(nullptr, 0);
				}
				else
				{
					std::copy(data, data + writeLen, _buffer.data() + _offset);
					_offset += writeLen;
					flush();
				}
				length -= writeLen; data += writeLen;
				continue;
			}
			std::copy(data, data + length, _buffer.data() + _offset);
			_offset += length;
			length = 0;
		}
		return *this;
	}

	void mark_bad_chunk(unsigned len)
	{
		_badChunk = true;
		if (_offset < _alignOffset) _good = false;
		_offset += len;
		_offset = _offset % _alignIncrement;
	}

	void flush()
	{
		if (_offset > 0)
		{
			_stream.write(_buffer.data(), _offset);
			if (_onFlush) _onFlush(_buffer.data(), _offset);
		}
		_totalCount += _offset;
		_offset = 0;
	}

protected:
	STREAM& _stream;
	std::vector<char> _buffer;
	unsigned _offset, _alignOffset, _alignIncrement;
	std::function<void(char*, size_t)> _onFlush;
	bool _badChunk, _good;
	size_t _totalCount = 0;
};

template <typename STREAM>
inline aligned_stream<STREAM>& operator<<(aligned_stream<STREAM>& s, const BadChunk& chunk)
{
	s.mark_bad_chunk(chunk.size);
	return s;
}

class escrow_buffer_writer
{
public:
	escrow_buffer_writer(unsigned char* bufspace, unsigned bufcount, unsigned bufsize)
		: _bufspace(bufspace), _bufcount(bufcount), _bufsize(bufsize) {}

	bool good() const { return _good; }
	unsigned chunk_size() const { return _bufsize; }
	long tellp() const { return _totalCount; }
	unsigned buffers_in_use() const { return _bufIdx; }

	escrow_buffer_writer& write(const char* data, unsigned length)
	{
		if (length != _bufsize or _bufIdx >= _bufcount) _good = false;
		if (!good()) return *this;
		std::copy(data, data + length, _bufspace + (_bufIdx * _bufsize));
		_totalCount += length;
		++_bufIdx;
		return *this;
	}

protected:
	unsigned char* _bufspace;
	unsigned _bufcount, _bufsize;
	unsigned _bufIdx = 0;
	long _totalCount = 0;
	bool _good = true;
};

}  // namespace cb200
