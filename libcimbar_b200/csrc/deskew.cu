// deskew.cu -- extractor stage in front of the decode path (filled in below)
#include "ctx.cuh"

namespace cb200 {
struct DeskewState { int unused = 0; };
void deskew_destroy(DeskewState* d) { delete d; }
}  // namespace cb200
