// streams.h -- what the Decoder mirror needs to talk to the caller's output stream: the BadChunk tag and how a failed RS block
// lands in a plain std stream.  (reference: ReedSolomon::BadChunk, src/lib/encoder/ReedSolomon.h:12-20; the operator<< overloads of
// src/lib/encoder/reed_solomon_stream.h:96-114.)  The stream classes themselves -- aligned_stream, escrow_buffer_writer,
// fountain_decoder_sink, std::ofstream ... -- are the CALLER's: the mirror only uses the concept they share
// (write(const char*, unsigned), good(), tellp(), and chunk_size() for fountain streams).
#pragma once
#include <ostream>
#include <string>

namespace cb200 {

struct BadChunk
{
	unsigned size;
	explicit BadChunk(unsigned size) : size(size) {}
};

// a failed RS block in a plain byte stream is `size` zero bytes (reed_solomon_stream.h:96-107)
inline std::ostream& operator<<(std::ostream& os, const BadChunk& chunk)
{
	std::string zeros(chunk.size, '\0');
	os.write(zeros.data(), (std::streamsize)zeros.size());
	return os;
}

}  // namespace cb200
