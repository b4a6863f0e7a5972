// k1x_flood.cuh -- host-side entry points of the exact flood-walk kernel (k1x_flood.cu)
#pragma once
#include "cb200_common.cuh"
#include "ccm.cuh"
#include <cstring>
#include <cstdlib>
#include <vector>

namespace cb200 {

// per-cell record of the exact walk (== what CimbReader::read returns step by step, CimbReader.cpp:139-162)
struct CellTrace {
    uint16_t order;        // position in the flood-walk order
    int16_t x, y;          // drift-adjusted cell position (PositionData x, y)
    uint8_t drift_offset;  // winning hash id 0..8 (4 = centre)
    uint8_t distance;      // best Hamming distance
};

struct FloodWorkspace {
    int sm_count;
    int slots;                 // walking warps resident at once (one frame each)
    int heap_smem;             // heap entries per walk kept in shared memory (odd)
    size_t walk_smem;          // dynamic shared memory of one walking warp
    int heap_smem_few; size_t walk_smem_few; int few_frames;   // batches of at most few_frames listed frames: the whole heap in shared memory
    size_t spill_cap;          // heap spill entries per slot
    int serial_above;          // heap sizes above this use the one-level-per-step pop (65536; tests lower it)
    uint32_t* spill;           // [slots][spill_cap]
    uint8_t* prio;             // [slots][kMaxCells] per-cell priority bytes of the walk in that slot
    uint16_t* cinfo;           // [num_cells][16] update candidates in push order (0xFFFF = none)
    int list_cap; uint32_t* list; uint32_t* counters;   // work list; counters[0] = listed frames, [1 + c] = chunk c's work counter
    int max_entries;           // upper bound of entry_cap (one chunk); larger work lists are processed chunk by chunk
    int entry_cap; uint16_t* raster; uint32_t* result;  // per listed frame of a chunk: 1-bit raster in 16x16 tiles, per-cell x | y<<11 | sym<<22
};

cudaError_t flood_init_tables(const float* adjust256, const unsigned long long* tiles_L16, uint32_t hash_mul);
// adj_host: [num_cells][4] = AdjacentCellFinder::find for every cell (built by the caller from the cell geometry)
cudaError_t flood_workspace_create(const Mode& m, int sm_count, const uint16_t* adj_host, FloodWorkspace* ws);
void flood_workspace_destroy(FloodWorkspace* ws);
// writes d_flags[f] for every frame: 0 = K1 result stands, CB200_FRAME_FALLBACK = re-decoded here,
// CB200_FRAME_INEXACT = needed but skipped (no_fallback)
cudaError_t flood_launch(const Mode& m, FloodWorkspace& ws, const uint8_t* d_rgb, int n_frames, bool no_fallback,
                         bool force_all, bool sharpen, uint8_t* d_cellvals, const uint32_t* d_dirty, uint8_t* d_flags, CellTrace* d_trace,
                         const CcmArg& cc, cudaStream_t st);

}  // namespace cb200
