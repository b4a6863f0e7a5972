/* cimbard_b200.h -- the reference's receive-side C facade, served by libcimbar_b200/lib/libcimbard_b200.so (over libcb200.so).

   The first six declarations are the reference's own, name and signature (src/lib/cimbar_js/cimbar_recv_js.h:11-39): a
   caller that binds cimbard_* from the reference's library binds the same symbols here.  Not provided: the zstd read-back
   (cimbard_get_filename, cimbard_get_decompress_bufsize, cimbard_decompress_read -- consumers of the finished file) and
   cimbard_get_debug.  cimbard_scan_extract_decode scans, extracts and decodes on the GPU exactly as the reference does on
   the CPU (Scanner + Extractor + Decoder with should_preprocess = true), -3 when fewer than four anchors are found;
   cimbard_b200_extract_decode is for a caller with a scanner of its own, or with `cimbar --no-deskew` input. */
#ifndef CIMBARD_B200_H
#define CIMBARD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cimbar_recv_js.h:11 */
unsigned cimbard_get_report(unsigned char* buff, unsigned maxlen);
/* cimbar_recv_js.h:16: fountain_chunks_per_frame * fountain_chunk_size of the configured mode */
int cimbard_get_bufsize();
/* cimbar_recv_js.h:17: format 3 = RGB, 4 = RGBA (<= 0: 3).  Returns the good bytes written to bufspace (whole chunks, packed
   from the front), -1 bad image size, -2 bufsize too small, -3 extract failed (fewer than four anchors), -4 unsupported format
   (the YUV layouts 12 / 420), -6 GPU error (text through cimbard_get_report) */
int cimbard_scan_extract_decode(const unsigned char* imgdata, unsigned imgw, unsigned imgh, int format, unsigned char* bufspace, unsigned bufsize);
/* cimbar_recv_js.h:21: chunks from cimbard_scan_extract_decode; > 0 = id of a completed file, 0 = progress, negative = error
   (-5: size is not whole chunks; the sink's own -1 / -10 / -11 / -12) */
int64_t cimbard_fountain_decode(const unsigned char* buffer, unsigned size);
/* cimbar_recv_js.h:25 */
unsigned cimbard_get_filesize(uint32_t id);
/* cimbar_recv_js.h:36: 68 (B, the default for <= 0 and unknown values), 67, 66, 4, 8; a change drops the sink */
int cimbard_configure_decode(int mode_val);

/* ---- additions ---- */
/* as cimbard_scan_extract_decode, with the four anchor centres of the camera image (x, y pairs: top-left, top-right,
   bottom-left, bottom-right) found by the caller's scanner; NULL = the image is an extracted frame of the mode's size and is
   decoded as it is (no scan, no warp; -3 for any other size) */
int cimbard_b200_extract_decode(const unsigned char* imgdata, unsigned imgw, unsigned imgh, int format, const float* corners,
                                unsigned char* bufspace, unsigned bufsize);
/* the reassembled file of a completed id (size >= cimbard_get_filesize(id)): bytes copied, -1 unknown id, -2 size too small */
int64_t cimbard_b200_file_read(uint32_t id, unsigned char* out, uint64_t size);
/* drops the GPU context and the sink */
void cimbard_b200_reset(void);

#ifdef __cplusplus
}
#endif

#endif /* CIMBARD_B200_H */
