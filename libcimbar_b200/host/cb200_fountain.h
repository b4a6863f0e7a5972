// cb200_fountain.h -- host-side fountain ingest (P14): FountainMetadata + the sink's accept/de-dup logic, restated
// (reference: src/lib/fountain/FountainMetadata.h:16-90, fountain_decoder_sink.h:133-166, fountain_decoder_stream.h:45-79,
// FountainDecoder.h:48-60).  The fountain codec itself (wirehair, third-party, O(N) on the CPU) is NOT reimplemented:
// it is plugged in through the three C function pointers below, which have exactly wirehair's C API signatures
// (src/third_party_lib/wirehair/include/wirehair/wirehair.h), so an integrated build passes wirehair_decoder_create /
// wirehair_decode / wirehair_recover / wirehair_free and gets the reference's behaviour chunk for chunk.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <vector>

namespace cb200 {

struct FountainMetadata
{
	static const unsigned md_size = 6;
	uint8_t d[6] = {0, 0, 0, 0, 0, 0};

	FountainMetadata() {}
	FountainMetadata(const uint8_t* buff, unsigned len) { std::memcpy(d, buff, len > md_size ? md_size : len); }
	FountainMetadata(uint8_t encode_id, unsigned size, uint16_t block_id)
	{
		d[0] = (encode_id & 0x7F) | ((size >> 17) & 0x80);
		d[1] = (size >> 16) & 0xFF; d[2] = (size >> 8) & 0xFF; d[3] = size & 0xFF;
		d[4] = (block_id >> 8) & 0xFF; d[5] = block_id & 0xFF;
	}
	uint32_t id() const { uint32_t v; std::memcpy(&v, d, 4); return v; }
	uint8_t encode_id() const { return d[0] & 0x7F; }
	uint16_t block_id() const { return (uint16_t)((d[4] << 8) | d[5]); }
	unsigned file_size() const { return (unsigned)d[3] | ((unsigned)d[2] << 8) | ((unsigned)d[1] << 16) | (((unsigned)d[0] & 0x80u) << 17); }
};

struct FountainCodecApi   // == wirehair's C API
{
	void* (*decoder_create)(void* reuse, uint64_t message_bytes, uint32_t block_bytes);
	int (*decode)(void* codec, unsigned block_id, const void* block_data, uint32_t data_bytes);   // 0 = success (complete), 1 = need more
	int (*recover)(void* codec, void* message_out, uint64_t message_bytes);
	void (*free_codec)(void* codec);
};

// fountain_decoder_sink::decode_frame semantics: >0 file id when complete, 0 progress, -1 already done,
// -10/-11/-12 malformed (fountain_decoder_sink.h:133-166).  Up to 8 concurrent streams keyed by encode_id & 7 (:198-201).
class fountain_sink
{
public:
	fountain_sink(unsigned chunk_size, const FountainCodecApi& api) : _chunkSize(chunk_size), _api(api) {}
	~fountain_sink() { for (auto& kv : _streams) _api.free_codec(kv.second.codec); }

	bool good() const { return true; }
	unsigned chunk_size() const { return _chunkSize; }

	int64_t decode_frame(const char* data, unsigned size)
	{
		if (size < FountainMetadata::md_size) return -10;
		FountainMetadata md(reinterpret_cast<const uint8_t*>(data), size);
		if (!md.file_size()) return -11;
		if (_done.count(md.id())) return -1;
		uint8_t slot = md.encode_id() & 0x7;
		auto it = _streams.find(slot);
		if (it == _streams.end())
		{
			Stream st;
			st.size = md.file_size();
			st.codec = _api.decoder_create(nullptr, st.size, _chunkSize - FountainMetadata::md_size);
			st.buffer.resize(_chunkSize);
			it = _streams.emplace(slot, std::move(st)).first;
		}
		Stream& s = it->second;
		if (s.size != md.file_size()) return -12;
		// fountain_decoder_stream::write (fountain_decoder_stream.h:52-74): bytes accumulate in a chunk-sized buffer -- a short
		// write is kept until the rest arrives, a long one spans chunks -- and every full buffer is one block for the codec
		bool finished = false;
		while (size > 0 and s.codec != nullptr)
		{
			unsigned n = std::min<unsigned>(size, _chunkSize - s.fill);
			std::memcpy(s.buffer.data() + s.fill, data, n);
			s.fill += n; data += n; size -= n;
			if (s.fill == _chunkSize)
			{
				s.fill = 0;
				unsigned block_id = ((unsigned)s.buffer[4] << 8) | s.buffer[5];
				if (!s.seen.insert(block_id).second) continue;            // FountainDecoder.h:48-52: never feed a block twice
				if (_api.decode(s.codec, block_id, s.buffer.data() + FountainMetadata::md_size, _chunkSize - FountainMetadata::md_size) == 0)
				{
					finished = true;
					break;
				}
			}
		}
		if (!finished) return 0;
		// complete: recover now (the reference's store() path with a write callback) and drop the stream.  Like the reference
		// the return value is the file id even when recover fails (fountain_decoder_sink.h:158-166 ignores store()'s result);
		// the file then simply is not available from file()
		std::vector<uint8_t> bytes(s.size);
		if (_api.recover(s.codec, bytes.data(), bytes.size()) == 0)
		{
			_done[md.id()] = std::move(bytes);
			_api.free_codec(s.codec);
			_streams.erase(it);
		}
		return (int64_t)md.id();
	}

	bool write(const char* data, unsigned length) { return decode_frame(data, length) > 0; }

	// feed the fixed-slot output of cb200_decode_chunks_dev / the gathered records of all ranks
	int64_t ingest(const uint8_t* chunks, const uint32_t* masks, int n_frames, int chunks_per_frame)
	{
		int64_t last = 0;
		for (int f = 0; f < n_frames; ++f)
			for (int q = 0; q < chunks_per_frame; ++q)
				if (masks[f] & (1u << q))
				{
					int64_t r = decode_frame(reinterpret_cast<const char*>(chunks) + ((size_t)f * chunks_per_frame + q) * _chunkSize, _chunkSize);
					if (r > 0) last = r;
				}
		return last;
	}

	bool is_done(uint32_t id) const { return _done.count(id) != 0; }
	const std::vector<uint8_t>* file(uint32_t id) const { auto it = _done.find(id); return it == _done.end() ? nullptr : &it->second; }
	unsigned num_streams() const { return (unsigned)_streams.size(); }
	unsigned num_done() const { return (unsigned)_done.size(); }

protected:
	struct Stream { void* codec = nullptr; unsigned size = 0; std::set<unsigned> seen; std::vector<uint8_t> buffer; unsigned fill = 0; };
	unsigned _chunkSize;
	FountainCodecApi _api;
	std::map<uint8_t, Stream> _streams;
	std::map<uint32_t, std::vector<uint8_t>> _done;
};

}  // namespace cb200
