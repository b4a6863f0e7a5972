// cb200_common.cuh -- shared definitions for the sm_100a decode kernels.
// Geometry mirrors cimbar::conf (reference: src/lib/cimb_translator/GridConf.h:8-190, Config.h:20-175).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cb200 {

constexpr int kMaxCells = 12544;   // 112*112
constexpr int kCellSize = 8;       // Config::cell_size() is constexpr 8 (Config.h:106-110)
constexpr int kSpacing = 9;        // every 8x8 mode uses cell_size+1 (GridConf.h:121-189)

// POD mode description, passed to kernels by value (lives in the kernel parameter bank = constant memory).
struct Mode {
    int mode_val;
    int color_bits, symbol_bits;
    int ecc_bytes, ecc_block;          // RS(ecc_block, ecc_block - ecc_bytes), GridConf.h:128-129
    int width, height;                 // image_size_x/y
    int cell_offset;                   // first cell coordinate (8 or 9)
    int cells_x, cells_y;              // cells_per_col_x/y
    int corner;                        // corner_padding (6): anchor exclusion in cells
    int legacy;                        // coupled 6-bit layout (Decoder.h:121-161)
    int color_mode;                    // 0 = legacy palette, 1 = mode-B palette (Config.h:61-64)
    int num_cells;                     // total_cells
    int cap_sym, cap_col, cap_all;     // capacity(symbol_bits), capacity(color_bits), capacity(all) in bytes
    int msg_len;                       // ecc_block - ecc_bytes
    int nblocks;                       // RS blocks per frame
    int nblocks_sym;                   // RS blocks in the symbol stream (== nblocks when legacy)
    int chunk_size, chunks_per_frame;  // fountain chunk geometry (GridConf.h:54-72)
    int blocks_per_chunk;
    int data_bytes;                    // nblocks * msg_len (7500 for mode B)
    int top_cells;                     // cells in the top (and bottom) marker rows: (cells_x - 2*corner) * corner
    int mid_cells;                     // cells_x * (cells_y - 2*corner)
    uint32_t hash_mul;                 // perfect hash over the tile dictionary: slot = (L_lo * hash_mul) >> 28
    uint8_t palette[8][4];             // decode palette for (num_colors, color_mode): Common.cpp:21-85, :122-139
};

// cell row k (0..cells_y-1): first cell index, number of cells, x of first cell
__host__ __device__ inline void cell_row_geom(const Mode& m, int k, int& base, int& ncols, int& x0)
{
    int narrow = m.cells_x - 2 * m.corner;
    if (k < m.corner) { base = k * narrow; ncols = narrow; x0 = m.cell_offset + kSpacing * m.corner; }
    else if (k < m.cells_y - m.corner) { base = m.top_cells + (k - m.corner) * m.cells_x; ncols = m.cells_x; x0 = m.cell_offset; }
    else { base = m.top_cells + m.mid_cells + (k - (m.cells_y - m.corner)) * narrow; ncols = narrow; x0 = m.cell_offset + kSpacing * m.corner; }
}

// per-cell result byte written by K1/K1x
constexpr uint8_t kCellDirty = 0x80;   // centre hash did not win -> frame needs the exact flood walk

// frame flags
constexpr uint8_t kFrameFallback = 0x01;   // exact flood-walk kernel was used for this frame
constexpr uint8_t kFrameDirtyK1 = 0x02;    // (internal) K1 saw a cell whose centre hash did not win

}  // namespace cb200
