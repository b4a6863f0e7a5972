/* scan_oracle.h -- TEST INFRASTRUCTURE: C API of the anchor-scan restatement (scan_oracle.c).  Not linked into the product. */
#ifndef CB200_SCAN_ORACLE_H
#define CB200_SCAN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cbo_anchor { int x, xmax, y, ymax; } cbo_anchor;                 /* extractor/Anchor.h:13-18 */
typedef struct cbo_scanner { const uint8_t* bin; int w, h, skip, merge_cutoff, anchor_size; } cbo_scanner;   /* Scanner.h:87-91 (dark) */

int cbo_scan_blur_size(int w, int h);
int cbo_scan_gaussian_blur(const uint8_t* src, int w, int h, int ksize, uint8_t* dst);
int cbo_scan_otsu(const uint8_t* img, size_t n);
int cbo_scan_preprocess(const uint8_t* rgb, int w, int h, uint8_t* bin, uint8_t* blurred);
int cbo_scan_preprocess_adaptive(const uint8_t* rgb, int w, int h, uint8_t* bin);
int cbo_scan2(const uint8_t* rgb, int w, int h, int fast, cbo_anchor* out, unsigned* cutoff_out);
void cbo_scanner_init(cbo_scanner* s, const uint8_t* bin, int w, int h, int skip);
/* kind: 114 = ScanState_114, 122 = ScanState_122.  The list functions return the total count (out holds min(count, cap)). */
int cbo_scan_t1(const cbo_scanner* s, int kind, int skip, int y, int yend, int xstart, int xend, cbo_anchor* out, int cap);
int cbo_scan_t2(const cbo_scanner* s, int kind, const cbo_anchor* hint, cbo_anchor* out, int cap);
int cbo_scan_t3(const cbo_scanner* s, int kind, const cbo_anchor* hint, cbo_anchor* out);
int cbo_scan_t4(const cbo_scanner* s, int kind, const cbo_anchor* hint, int merge_confirms, cbo_anchor* out);
int cbo_scan_deduplicate(const cbo_scanner* s, const cbo_anchor* in, int n, cbo_anchor* out);
int cbo_scan_filter(cbo_anchor* v, int n, unsigned* cutoff);
int cbo_scan_sort_top_to_bottom(cbo_anchor* v, int n);
int cbo_scan_primary(const cbo_scanner* s, cbo_anchor* out, int cap, unsigned* cutoff);
int cbo_scan_bottom_right(const cbo_scanner* s, cbo_anchor* anchors, unsigned cutoff);
int cbo_scan_bin(const uint8_t* bin, int w, int h, cbo_anchor* out, unsigned* cutoff_out);
int cbo_scan(const uint8_t* rgb, int w, int h, cbo_anchor* out, unsigned* cutoff_out);
void cbo_scan_corners(const cbo_anchor* a4, int* xy8);
int cbo_scan_is_granular_scale(const int* xy8, int min_w, int min_h);

#ifdef __cplusplus
}
#endif
#endif
