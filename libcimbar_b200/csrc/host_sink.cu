// host_sink.cu -- C ABI around the host-side fountain sink (libcimbar_b200/host/cb200_fountain.h).  Pure host code
// (compiled by nvcc only so that it lands in the same shared library); needs no GPU.
#include "../../include/cb200.h"
#include "../host/cb200_fountain.h"

#include <dlfcn.h>
#include <cstring>
#include <new>
#include <string>

using cb200::fountain_sink;
using cb200::FountainCodecApi;

struct cb200_sink { fountain_sink* impl; };

extern "C" {

cb200_sink* cb200_sink_create(unsigned chunk_size, cb200_codec_create_fn create_fn, cb200_codec_decode_fn decode_fn,
                              cb200_codec_recover_fn recover_fn, cb200_codec_free_fn free_fn)
{
    if (chunk_size <= 6 || !create_fn || !decode_fn || !recover_fn || !free_fn) return nullptr;
    FountainCodecApi api{create_fn, decode_fn, recover_fn, free_fn};
    cb200_sink* s = new (std::nothrow) cb200_sink;
    if (!s) return nullptr;
    s->impl = new fountain_sink(chunk_size, api);
    return s;
}

// the reference's own fountain codec, bound at run time: wirehair's C API (wirehair/wirehair.h) from a shared library
cb200_sink* cb200_sink_create_wirehair(unsigned chunk_size, const char* wirehair_library)
{
    if (chunk_size <= 6) return nullptr;
    void* lib = dlopen(wirehair_library && *wirehair_library ? wirehair_library : "libwirehair.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) return nullptr;
    typedef int (*init_fn)(int);
    init_fn init = reinterpret_cast<init_fn>(dlsym(lib, "wirehair_init_"));
    cb200_codec_create_fn c = reinterpret_cast<cb200_codec_create_fn>(dlsym(lib, "wirehair_decoder_create"));
    cb200_codec_decode_fn d = reinterpret_cast<cb200_codec_decode_fn>(dlsym(lib, "wirehair_decode"));
    cb200_codec_recover_fn r = reinterpret_cast<cb200_codec_recover_fn>(dlsym(lib, "wirehair_recover"));
    cb200_codec_free_fn f = reinterpret_cast<cb200_codec_free_fn>(dlsym(lib, "wirehair_free"));
    if (!init || !c || !d || !r || !f || init(2 /* WIREHAIR_VERSION */) != 0) return nullptr;
    return cb200_sink_create(chunk_size, c, d, r, f);
}

void cb200_sink_destroy(cb200_sink* s)
{
    if (!s) return;
    delete s->impl;
    delete s;
}

int64_t cb200_sink_decode_frame(cb200_sink* s, const uint8_t* chunk, unsigned size)
{
    if (!s || !chunk) return -10;
    return s->impl->decode_frame(reinterpret_cast<const char*>(chunk), size);
}

int64_t cb200_sink_ingest(cb200_sink* s, const uint8_t* chunks, const uint32_t* masks, int n_frames, int chunks_per_frame)
{
    if (!s || !chunks || !masks) return -10;
    return s->impl->ingest(chunks, masks, n_frames, chunks_per_frame);
}

int64_t cb200_sink_file_size(const cb200_sink* s, uint32_t id)
{
    if (!s) return -1;
    const std::vector<uint8_t>* f = s->impl->file(id);
    return f ? (int64_t)f->size() : -1;
}

int cb200_sink_file_read(const cb200_sink* s, uint32_t id, uint8_t* out, uint64_t size)
{
    if (!s || !out) return CB200_ERR_ARG;
    const std::vector<uint8_t>* f = s->impl->file(id);
    if (!f || f->size() != size) return CB200_ERR_ARG;
    std::memcpy(out, f->data(), size);
    return CB200_OK;
}

}  // extern "C"
