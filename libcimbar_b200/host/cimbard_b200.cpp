// cimbard_b200.cpp -- the reference's receive-side C facade over libcb200 (lib/libcimbard_b200.so).
//
// Replaces, name for name: src/lib/cimbar_js/cimbar_recv_js.h:11-39 / cimbar_recv_js.cpp:122-300 -- the C entry points the
// reference's wasm receiver and its Android/JNI callers bind -- for the decode path:
//   cimbard_configure_decode      cimbar_recv_js.cpp:274-290   (Config::update + a fresh sink when the mode changes)
//   cimbard_get_bufsize           :147-150                     (fountain_chunks_per_frame * fountain_chunk_size)
//   cimbard_scan_extract_decode   :152-189                     (get_rgb + Extractor::extract + Decoder::decode_fountain into an
//                                                               escrow_buffer_writer; returns buffers_in_use * chunk size)
//   cimbard_fountain_decode       :192-216                     (fountain_decoder_sink::decode_frame per chunk; stops at the first
//                                                               nonzero result; -5 on a size that is not whole chunks)
//   cimbard_get_filesize          :219-223                     (FountainMetadata(id).file_size())
//   cimbard_get_report            :124-131
// cimbard_scan_extract_decode runs the whole of it on the GPU: cb200_scan_extract_decode_fountain = Scanner::scan + Corners +
// Deskewer::deskew + Decoder::decode_fountain with should_preprocess = true, as the reference does for EVERY image it is given
// (an image that already is an extracted frame is scanned and warped again there too); fewer than four anchors -> -3.
// What stays outside: the zstd read-back (cimbard_get_filename / cimbard_decompress_read), a consumer of the finished file.
// Additions: cimbard_b200_extract_decode takes the anchor centres from a caller that has its own scanner (or NULL for
// `./cimbar --no-deskew` input: an image of the mode's size, decoded as it is); the recovered file is read back, undecompressed,
// with cimbard_b200_file_read (what cimbard_get_reassembled_file_buff exposes for the tests).
//
// State is per process, like the reference's (file-static sink and mode); calls are not thread-safe, like the reference's.
#include "../../include/cb200.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

namespace {

cb200_ctx* g_ctx = nullptr;
cb200_sink* g_sink = nullptr;
int g_mode = 68;
std::string g_report;
std::vector<uint8_t> g_rgb, g_chunks;

int known_mode(int mode_val)
{   // Config::temp_conf: every unknown value is mode B (Config.h:20-44)
    return (mode_val == 4 || mode_val == 8 || mode_val == 66 || mode_val == 67) ? mode_val : 68;
}

bool mode_info(cb200_info& info) { return cb200_mode_info(g_mode, &info) == CB200_OK; }

int ensure_ctx()
{
    if (g_ctx) return 0;
    const char* dev = getenv("CB200_DEVICE");
    int rc = cb200_create(&g_ctx, dev ? atoi(dev) : 0, g_mode, 1);
    if (rc != CB200_OK) { g_ctx = nullptr; g_report = std::string("cb200_create: ") + cb200_last_error(); return rc; }
    return 0;
}

std::string wirehair_path()
{   // the codec library sits next to this one (libcimbar_b200/lib/); CB200_WIREHAIR_LIB overrides
    if (const char* e = getenv("CB200_WIREHAIR_LIB")) return e;
    Dl_info di;
    if (dladdr(reinterpret_cast<void*>(&known_mode), &di) && di.dli_fname) {
        std::string p = di.dli_fname;
        const size_t slash = p.rfind('/');
        return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/libwirehair.so";
    }
    return "libwirehair.so";
}

// get_rgb (cimbar_recv_js.cpp:92-119): format 3 = RGB, 4 = RGBA (alpha dropped).  The two YUV layouts (12, 420) need
// OpenCV's colour conversion, which is not restated: nullptr.
const uint8_t* to_rgb(const unsigned char* imgdata, unsigned w, unsigned h, int format)
{
    if (format == 3) return imgdata;
    if (format != 4) return nullptr;
    g_rgb.resize((size_t)w * h * 3);
    for (size_t i = 0, n = (size_t)w * h; i < n; ++i) {
        g_rgb[3 * i] = imgdata[4 * i]; g_rgb[3 * i + 1] = imgdata[4 * i + 1]; g_rgb[3 * i + 2] = imgdata[4 * i + 2];
    }
    return g_rgb.data();
}

// the good chunks of one frame, packed from the front of bufspace (what escrow_buffer_writer leaves there)
int finish(const cb200_info& info, uint32_t count, unsigned char* bufspace)
{
    std::memcpy(bufspace, g_chunks.data(), (size_t)count * info.chunk_size);
    char text[96];
    snprintf(text, sizeof text, "decoded %u bytes", count * (unsigned)info.chunk_size);
    g_report = text;
    return (int)(count * (unsigned)info.chunk_size);
}

}  // namespace

extern "C" {

unsigned cimbard_get_report(unsigned char* buff, unsigned maxlen)
{
    const unsigned len = g_report.size() < maxlen ? (unsigned)g_report.size() : maxlen;
    if (len) std::memcpy(buff, g_report.data(), len);
    return len;
}

int cimbard_configure_decode(int mode_val)
{
    mode_val = known_mode(mode_val <= 0 ? 68 : mode_val);
    if (mode_val != g_mode) {
        g_mode = mode_val;
        if (g_ctx) { cb200_destroy(g_ctx); g_ctx = nullptr; }
        if (g_sink) { cb200_sink_destroy(g_sink); g_sink = nullptr; }
    }
    return 0;
}

int cimbard_get_bufsize()
{
    cb200_info info;
    if (!mode_info(info)) return 0;
    return info.chunks_per_frame * info.chunk_size;
}

// corners: the four anchor centres (x, y pairs; top-left, top-right, bottom-left, bottom-right -- Corners::all()) or NULL for
// an image that already is the mode's frame
int cimbard_b200_extract_decode(const unsigned char* imgdata, unsigned imgw, unsigned imgh, int format, const float* corners,
                                unsigned char* bufspace, unsigned bufsize)
{
    if (format <= 0) format = 3;
    if (imgw == 0 || imgh == 0 || !imgdata || !bufspace) return -1;
    cb200_info info;
    if (!mode_info(info)) return -1;
    if (bufsize < (unsigned)(info.chunk_size * info.chunks_per_frame)) return -2;
    if (!corners && ((int)imgw != info.image_size_x || (int)imgh != info.image_size_y)) return -3;
    const uint8_t* rgb = to_rgb(imgdata, imgw, imgh, format);
    if (!rgb) return -4;
    if (ensure_ctx() != 0) return -6;
    g_chunks.resize((size_t)info.chunk_size * info.chunks_per_frame);
    uint32_t count = 0;
    // shouldPreprocess = true in the reference (cimbar_recv_js.cpp:171): the sharpening preprocess
    const uint32_t flags = CB200_FLAG_SHARPEN | CB200_FLAG_CC_FIT;
    const int rc = corners ? cb200_extract_decode_fountain(g_ctx, rgb, (int)imgw, (int)imgh, 1, corners, flags, g_chunks.data(), &count, nullptr, nullptr)
                           : cb200_decode_fountain(g_ctx, rgb, 1, flags, g_chunks.data(), &count, nullptr, nullptr);
    if (rc != CB200_OK) { g_report = std::string("decode: ") + cb200_last_error(); return -6; }
    return finish(info, count, bufspace);
}

int cimbard_scan_extract_decode(const unsigned char* imgdata, unsigned imgw, unsigned imgh, int format, unsigned char* bufspace, unsigned bufsize)
{
    if (format <= 0) format = 3;
    if (imgw == 0 || imgh == 0 || !imgdata || !bufspace) return -1;
    cb200_info info;
    if (!mode_info(info)) return -1;
    if (bufsize < (unsigned)(info.chunk_size * info.chunks_per_frame)) return -2;
    const uint8_t* rgb = to_rgb(imgdata, imgw, imgh, format);
    if (!rgb) return -4;
    if (ensure_ctx() != 0) return -6;
    g_chunks.resize((size_t)info.chunk_size * info.chunks_per_frame);
    uint32_t count = 0;
    int32_t status = 0;
    // Extractor::extract, then decode_fountain(img, ebw, shouldPreprocess = true) (cimbar_recv_js.cpp:171-186; color_correction
    // defaults to 2, Decoder.h:31)
    const uint32_t flags = CB200_FLAG_SHARPEN | CB200_FLAG_CC_FIT;
    const int rc = cb200_scan_extract_decode_fountain(g_ctx, rgb, (int)imgw, (int)imgh, 1, flags, g_chunks.data(), &count, nullptr, nullptr, &status);
    if (rc != CB200_OK) { g_report = std::string("scan/extract/decode: ") + cb200_last_error(); return -6; }
    if (status <= 0) return -3;                                   // Extractor::FAILURE (cimbar_recv_js.cpp:175-176)
    return finish(info, count, bufspace);
}

int64_t cimbard_fountain_decode(const unsigned char* buffer, unsigned size)
{
    cb200_info info;
    if (!mode_info(info)) return -5;
    const unsigned chunk = (unsigned)info.chunk_size;
    if (!g_sink) {
        g_sink = cb200_sink_create_wirehair(chunk, wirehair_path().c_str());
        if (!g_sink) { g_report = std::string("sink: ") + cb200_last_error(); return -6; }
    }
    if (size == 0 || size % chunk != 0) return -5;
    int64_t res = 0;
    for (unsigned i = 0; i < size && res == 0; i += chunk) res = cb200_sink_decode_frame(g_sink, buffer + i, chunk);
    return res;
}

unsigned cimbard_get_filesize(uint32_t id)
{   // FountainMetadata(id).file_size(): the id is the first four header bytes as they lie in memory (FountainMetadata.h:29-45)
    uint8_t d[4];
    std::memcpy(d, &id, 4);
    return (unsigned)d[3] | ((unsigned)d[2] << 8) | ((unsigned)d[1] << 16) | (((unsigned)d[0] & 0x80u) << 17);
}

// the reassembled (still compressed) file of a finished id: bytes copied, -1 when the sink does not hold it, -2 when `size`
// is smaller than cimbard_get_filesize(id)
int64_t cimbard_b200_file_read(uint32_t id, unsigned char* out, uint64_t size)
{
    if (!g_sink) return -1;
    const int64_t have = cb200_sink_file_size(g_sink, id);
    if (have < 0) return -1;
    if (size < (uint64_t)have) return -2;
    if (cb200_sink_file_read(g_sink, id, out, (uint64_t)have) != CB200_OK) return -1;
    return have;
}

void cimbard_b200_reset(void)
{
    if (g_ctx) { cb200_destroy(g_ctx); g_ctx = nullptr; }
    if (g_sink) { cb200_sink_destroy(g_sink); g_sink = nullptr; }
    g_report.clear();
}

}  // extern "C"
