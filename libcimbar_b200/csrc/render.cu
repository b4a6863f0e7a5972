// render.cu -- synthetic-frame generator on the device (input side of the benchmark, not part of the decode path).
// Restates what a frame looks like after Encoder::encode_next + CimbWriter::write
// (reference: src/lib/encoder/Encoder.h:69-129, src/lib/cimb_translator/CimbWriter.cpp:84-95, CimbEncoder.cpp:20-42,
//  Common.cpp:141-171): black canvas, cell i shows tile (value % 16) in palette colour (value / 16) at its linear position.
// Anchors and guides are not drawn (they lie outside every cell's 5x5 threshold support).
#include "cb200_common.cuh"
#include "render.cuh"

namespace cb200 {

__constant__ unsigned long long c_tiles_H[16];   // reference orientation: bit 63 = top-left

__global__ void __launch_bounds__(256)
k_render(const Mode m, const uint8_t* __restrict__ cellvals, int n_frames, uint8_t* __restrict__ rgb)
{
    // one thread per (cell, tile row)
    long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long per_frame = (long)m.num_cells * 8;
    if (gid >= per_frame * n_frames) return;
    int f = (int)(gid / per_frame);
    int rem = (int)(gid - (long)f * per_frame);
    int cell = rem >> 3, r = rem & 7;
    // linear cell index -> (row k, column) -> pixel position (CellPositions.cpp:5-50)
    int k, c;
    int narrow = m.cells_x - 2 * m.corner;
    if (cell < m.top_cells) { k = cell / narrow; c = cell - k * narrow; }
    else if (cell < m.top_cells + m.mid_cells) { int q = cell - m.top_cells; k = m.corner + q / m.cells_x; c = q % m.cells_x; }
    else { int q = cell - m.top_cells - m.mid_cells; k = m.cells_y - m.corner + q / narrow; c = q % narrow; }
    int base, ncols, x0;
    cell_row_geom(m, k, base, ncols, x0);
    int x = x0 + kSpacing * c, y = m.cell_offset + kSpacing * k + r;
    int nsym = 1 << m.symbol_bits, ncol = 1 << m.color_bits;
    int v = cellvals[(size_t)f * m.num_cells + cell] % (nsym * ncol);
    int sym = v % nsym, col = v / nsym;
    unsigned bits = (unsigned)(c_tiles_H[sym] >> (8 * (7 - r))) & 0xFFu;   // bit 7 = leftmost pixel
    uint8_t cr = m.palette[col][0], cg = m.palette[col][1], cb = m.palette[col][2];
    uint8_t* p = rgb + ((size_t)f * m.height + y) * (size_t)m.width * 3 + (size_t)x * 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bool on = (bits >> (7 - i)) & 1u;
        p[3 * i + 0] = on ? cr : 0; p[3 * i + 1] = on ? cg : 0; p[3 * i + 2] = on ? cb : 0;
    }
}

cudaError_t render_init_tables(const unsigned long long* tiles_H16)
{
    return cudaMemcpyToSymbol(c_tiles_H, tiles_H16, sizeof(unsigned long long) * 16);
}

cudaError_t render_launch(const Mode& m, const uint8_t* d_cellvals, int n_frames, uint8_t* d_rgb, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_rgb, 0, (size_t)n_frames * m.width * m.height * 3, st);
    if (e != cudaSuccess) return e;
    long total = (long)n_frames * m.num_cells * 8;
    k_render<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m, d_cellvals, n_frames, d_rgb); count_launch();
    return cudaGetLastError();
}

}  // namespace cb200
