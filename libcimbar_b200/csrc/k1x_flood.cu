// k1x_flood.cu -- K1x: the exact flood-walk decode for frames the drift-0 pass (K1) cannot prove exact, sm_100a.
//
// Restates the reference's serial semantics on the device (reference file:line relative to /root/reference/):
//   P1  preprocessSymbolGrid (+ sharpen)      src/lib/cimb_translator/CimbReader.cpp:17-46 (full frame, OpenCV borders)
//   P3  FloodDecodePositions                  src/lib/cimb_translator/FloodDecodePositions.cpp:17-134
//       std::priority_queue tie-breaking      libstdc++ bits/stl_heap.h (__push_heap / __adjust_heap), restated exactly:
//                                             equal priorities pop in the order that sift-up/sift-down produces
//   AdjacentCellFinder                        src/lib/cimb_translator/AdjacentCellFinder.cpp:54-105
//   P4  CimbReader::read, CellDrift           src/lib/cimb_translator/CimbReader.cpp:139-162, CellDrift.cpp:23-43
//   P5/P6 fuzzy_ahash + get_best_symbol       src/lib/image_hash/ahash_result.h:70-106, CimbDecoder.cpp:101-132
//   P8/P9 colour at the drift-adjusted cell   src/lib/cimb_translator/Cell.h:30-62, CimbDecoder.cpp:168-217
// One CTA per frame that needs it (clean frames cost one flag test).  The threshold raster of the whole frame is
// built by all 256 threads in global scratch; warp 0 then runs the 12 400-step walk: lane 0 owns the heap (global
// scratch) and the per-cell inherit table (shared memory), all 32 lanes score the 5x16 / 9x16 (hash, tile) candidates
// of each step with popcounts and a warp min-reduction; finally all threads classify colours at the recorded positions.
#include "cb200_common.cuh"
#include "k1x_flood.cuh"

namespace cb200 {

constexpr int kFloodThreads = 256;

__constant__ float cx_adjust[256];
__constant__ unsigned long long cx_tiles_L[16];

constexpr int kHeapSmem = 14336;        // heap entries kept in shared memory (observed maximum ~9.1k); the rest spills to global

struct FloodSmem {
    uint32_t instr[kMaxCells];          // before decode: dx(8) | dy(8) | prio(8) | cooldown(8); after: x(11) | y(11) | sym(4) | done
    uint32_t heap[kHeapSmem];
    uint32_t remaining[(kMaxCells + 31) / 32];
    float adjust[256];
    unsigned long long tiles[16];       // copy of cx_tiles_L: indexed per lane in the scoring loop
};

// ---------------------------------------------------------------------------------------------- geometry
__device__ __forceinline__ void cell_xy(const Mode& m, int index, int& x, int& y)
{   // CellPositions::compute_linear, CellPositions.cpp:5-50
    int narrow = m.cells_x - 2 * m.corner;
    if (index < m.top_cells) {
        int k = index / narrow, c = index - k * narrow;
        x = m.cell_offset + kSpacing * (m.corner + c); y = m.cell_offset + kSpacing * k;
    } else if (index < m.top_cells + m.mid_cells) {
        int q = index - m.top_cells; int k = q / m.cells_x, c = q - k * m.cells_x;
        x = m.cell_offset + kSpacing * c; y = m.cell_offset + kSpacing * (m.corner + k);
    } else {
        int q = index - m.top_cells - m.mid_cells; int k = q / narrow, c = q - k * narrow;
        x = m.cell_offset + kSpacing * (m.corner + c); y = m.cell_offset + kSpacing * (m.cells_y - m.corner + k);
    }
}
__device__ __forceinline__ int cell_x(const Mode& m, int index) { int x, y; cell_xy(m, index, x, y); return x; }

// ---------------------------------------------------------------------------------------------- heap (lane 0 only)
// entries: prio << 16 | index.  std::priority_queue<decode_prio, vector, PrioCompare> with comp(a,b) = a.prio > b.prio
struct Heap {
    uint32_t* sm; uint32_t* spill; int n;
    __device__ __forceinline__ uint32_t get(int i) const { return i < kHeapSmem ? sm[i] : spill[i - kHeapSmem]; }
    __device__ __forceinline__ void set(int i, uint32_t v) { if (i < kHeapSmem) sm[i] = v; else spill[i - kHeapSmem] = v; }
};
__device__ __forceinline__ uint32_t hprio(uint32_t e) { return e >> 16; }
__device__ void heap_push(Heap& h, uint32_t idx, uint32_t prio)
{
    int hole = h.n++;
    if (hole < kHeapSmem) {          // common case: the whole sift-up path lives in shared memory
        uint32_t* v = h.sm;
        int parent = (hole - 1) / 2;
        while (hole > 0) {
            uint32_t pe = v[parent];
            if (hprio(pe) <= prio) break;
            v[hole] = pe; hole = parent; parent = (hole - 1) / 2;
        }
        v[hole] = (prio << 16) | idx;
        return;
    }
    int parent = (hole - 1) / 2;
    while (hole > 0 && hprio(h.get(parent)) > prio) { h.set(hole, h.get(parent)); hole = parent; parent = (hole - 1) / 2; }
    h.set(hole, (prio << 16) | idx);
}
__device__ uint32_t heap_pop(Heap& h)
{
    uint32_t top = h.get(0);
    uint32_t value = h.get(h.n - 1);
    int len = --h.n;
    if (len == 0) return top;
    int hole = 0, second = 0;
    if (len <= kHeapSmem) {          // common case: shared memory only
        uint32_t* v = h.sm;
        const int lim = (len - 1) / 2;
        while (second < lim) {
            second = 2 * (second + 1);
            uint32_t a = v[second], b = v[second - 1];
            if (hprio(a) > hprio(b)) { second--; a = b; }
            v[hole] = a;
            hole = second;
        }
        if ((len & 1) == 0 && second == (len - 2) / 2) {
            second = 2 * (second + 1);
            v[hole] = v[second - 1];
            hole = second - 1;
        }
        const uint32_t vp = hprio(value);
        int parent = (hole - 1) / 2;
        while (hole > 0) {
            uint32_t pe = v[parent];
            if (hprio(pe) <= vp) break;
            v[hole] = pe; hole = parent; parent = (hole - 1) / 2;
        }
        v[hole] = value;
        return top;
    }
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        uint32_t a = h.get(second), b = h.get(second - 1);
        if (hprio(a) > hprio(b)) { second--; a = b; }
        h.set(hole, a);
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        h.set(hole, h.get(second - 1));
        hole = second - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > 0 && hprio(h.get(parent)) > hprio(value)) { h.set(hole, h.get(parent)); hole = parent; parent = (hole - 1) / 2; }
    h.set(hole, value);
    return top;
}

__device__ __forceinline__ uint32_t pack_instr(int dx, int dy, uint32_t prio, uint32_t cooldown)
{
    return ((uint32_t)(dx & 0xFF)) | ((uint32_t)(dy & 0xFF) << 8) | (prio << 16) | (cooldown << 24);
}
__device__ __forceinline__ bool is_remaining(const FloodSmem& s, int i) { return (s.remaining[i >> 5] >> (i & 31)) & 1u; }

// FloodDecodePositions::update (FloodDecodePositions.cpp:86-129) with update_adjacents (:69-83), done by the warp:
// lanes 0..11 each resolve one candidate neighbour -- lanes 0-3 the direct neighbours (right, left, bottom, top), lanes
// 4-7 the horizontal horizon (right of right, its right, left of left, its left), lanes 8-11 the vertical horizon -- and
// test it (still remaining? stored priority > err?); lane 0 then rewrites the inherit entries and pushes the survivors
// in the reference's order (adjacents, horizon, vert).  The candidates are distinct cells, so the tests are independent.
// adj: per-cell neighbours as AdjacentCellFinder::find computes them (AdjacentCellFinder.cpp:54-105), 0xFFFF = none.
__device__ __forceinline__ int adj_dir(const ushort4* __restrict__ adj, int cell, int dir)
{
    if (cell < 0) return -1;
    ushort4 a = __ldg(&adj[cell]);
    unsigned v = dir == 0 ? a.x : dir == 1 ? a.y : dir == 2 ? a.z : a.w;
    return v == 0xFFFFu ? -1 : (int)v;
}
__device__ void flood_update_warp(const ushort4* __restrict__ adj, FloodSmem& s, Heap& h, int lane, int index, int dx, int dy,
                                  uint32_t err, uint32_t cooldown, uint32_t self)
{
    const uint32_t prev_err = (self >> 16) & 0xFFu, prev_cd = self >> 24;
    const bool horizon = prev_err < 3 && err < 3 && prev_cd == 4 && cooldown == 4;
    int cand = -1;
    if (lane < 4) cand = adj_dir(adj, index, lane);
    else if (lane < 12 && horizon) {
        const int grp = (lane - 4) >> 1;               // 0: right chain, 1: left chain, 2: top chain, 3: bottom chain
        const int dir = grp == 0 ? 0 : grp == 1 ? 1 : grp == 2 ? 3 : 2;
        // horizontal horizon needs BOTH right and left neighbours, vertical BOTH top and bottom (FloodDecodePositions.cpp:102, :116)
        const int a0 = adj_dir(adj, index, grp < 2 ? 0 : 3), a1 = adj_dir(adj, index, grp < 2 ? 1 : 2);
        if (a0 >= 0 && a1 >= 0) {
            const int first = adj_dir(adj, dir == (grp < 2 ? 0 : 3) ? a0 : a1, dir);   // neighbour of the direct neighbour
            cand = ((lane - 4) & 1) ? adj_dir(adj, first, dir) : first;
        }
    }
    bool push = false;
    if (cand >= 0 && is_remaining(s, cand)) push = ((s.instr[cand] >> 16) & 0xFFu) > err;
    // reference order: adj = right, left, bottom, top; horizon = right+1, right+2, left+1, left+2; vert = top+1, top+2, bottom+1, bottom+2
    // lanes: 0..3 direct; 4,5 right chain; 6,7 left chain; 8,9 top chain; 10,11 bottom chain  -> already in reference order
    uint32_t todo = __ballot_sync(0xffffffffu, push) & 0xFFFu;
    const uint32_t entry = pack_instr(dx, dy, err, cooldown);
    while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int c = __shfl_sync(0xffffffffu, cand, l);
        if (lane == 0) { s.instr[c] = entry; heap_push(h, (uint32_t)c, err); }
    }
}

// ---------------------------------------------------------------------------------------------- preprocessing (P1)
__device__ __forceinline__ int reflect101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }
__device__ __forceinline__ int clampi(int p, int lo, int hi) { return p < lo ? lo : (p > hi ? hi : p); }

// gray -> (optionally sharpened) gray -> horizontal box sums -> threshold bits; all 256 threads, global scratch
__device__ void build_raster(const Mode& m, const uint8_t* __restrict__ frame, bool sharpen,
                             uint8_t* gray, uint8_t* gray2, uint16_t* hsum, uint32_t* raster)
{
    const int W = m.width, H = m.height, npx = W * H;
    for (int i = threadIdx.x; i < npx; i += kFloodThreads) {
        uint32_t r = frame[3 * i], g = frame[3 * i + 1], b = frame[3 * i + 2];
        gray[i] = (uint8_t)((9798u * r + 19235u * g + 3735u * b + 16384u) >> 15);   // cvtColor(RGB2GRAY)
    }
    __syncthreads();
    const uint8_t* src = gray;
    int radius = 2;
    if (sharpen) {
        // filter2D with [0 -1 0; -1 4.5 -1; 0 -1 0] (CimbReader.cpp:17-27): exact in float, cvRound = round-half-even,
        // saturate to 8 bits, BORDER_REFLECT_101; then block size 7 (CimbReader.cpp:37-40)
        for (int i = threadIdx.x; i < npx; i += kFloodThreads) {
            int y = i / W, x = i - y * W;
            int c = gray[i];
            int nb = gray[reflect101(y - 1, H) * W + x] + gray[reflect101(y + 1, H) * W + x] +
                     gray[y * W + reflect101(x - 1, W)] + gray[y * W + reflect101(x + 1, W)];
            int twice = 9 * c - 2 * nb, v;
            if (twice & 1) { int k = (twice - 1) / 2; v = (k & 1) ? k + 1 : k; } else v = twice / 2;
            gray2[i] = (uint8_t)clampi(v, 0, 255);
        }
        __syncthreads();
        src = gray2;
        radius = 3;
    }
    for (int i = threadIdx.x; i < npx; i += kFloodThreads) {
        int y = i / W, x = i - y * W;
        uint32_t sum = 0;
        for (int d = -radius; d <= radius; ++d) sum += src[y * W + clampi(x + d, 0, W - 1)];   // BORDER_REPLICATE
        hsum[i] = (uint16_t)sum;
    }
    __syncthreads();
    // adaptiveThreshold(MEAN_C, BINARY, bs, C=0): src > round(sum / bs^2)  <=>  bs^2 * src > sum + (bs^2 - 1) / 2
    const uint32_t area = (uint32_t)((2 * radius + 1) * (2 * radius + 1)), half = (area - 1) / 2;
    const int words = npx / 32;
    for (int wi = threadIdx.x; wi < words; wi += kFloodThreads) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            int i = wi * 32 + b;
            int y = i / W, x = i - y * W;
            uint32_t sum = 0;
            for (int d = -radius; d <= radius; ++d) sum += hsum[clampi(y + d, 0, H - 1) * W + x];
            if (area * (uint32_t)src[i] > sum + half) bits |= 1u << b;
        }
        raster[wi] = bits;   // little-endian: bit b of word wi = pixel wi*32+b
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- colour (P8/P9)
__device__ uint32_t flood_best_color(const float* adjust_tab, const Mode& m, uint32_t ri, uint32_t gi, uint32_t bi)
{   // CimbDecoder.cpp:168-200, float32 op for op (see k1_decode.cu best_color)
    float r = (float)ri, g = (float)gi, b = (float)bi;
    float mx = fmaxf(fmaxf(r, g), fmaxf(b, 1.0f));
    float mn = fminf(fminf(r, g), fminf(b, 48.0f));
    if (mn >= mx) mn = 0.0f;
    float adjust = adjust_tab[(int)(mx - mn)];
    int c[3];
    float in[3] = {r, g, b};
    for (int k = 0; k < 3; ++k) {
        float v = __fmul_rn(__fsub_rn(in[k], mn), adjust);
        if (v > __fsub_rn(245.0f, mn)) v = 255.0f;
        if (v < 0.0f) v = 0.0f;
        c[k] = (int)__float2uint_rz(v);
    }
    int a0 = c[0] - c[1], a1 = c[1] - c[2], a2 = c[2] - c[0];
    uint32_t best = 0, best_d = 0x7fffffffu;
    int num_colors = 1 << m.color_bits;
    for (int i = 0; i < num_colors; ++i) {
        int pr = m.palette[i][0], pg = m.palette[i][1], pb = m.palette[i][2];
        int d0 = a0 - (pr - pg), d1 = a1 - (pg - pb), d2 = a2 - (pb - pr);
        uint32_t d = (uint32_t)(d0 * d0 + d1 * d1 + d2 * d2);
        if (d < best_d) { best_d = d; best = (uint32_t)i; }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kFloodThreads, 2)
k_flood(const Mode m, const uint8_t* __restrict__ rgb, int n_frames, int no_fallback, int force_all, int sharpen,
        uint8_t* __restrict__ cellvals, const uint32_t* __restrict__ dirty, uint8_t* __restrict__ frame_flags,
        uint8_t* ws_gray, uint8_t* ws_gray2, uint16_t* ws_hsum, uint32_t* ws_raster, uint32_t* ws_heap, size_t heap_cap,
        const ushort4* __restrict__ adj, CellTrace* __restrict__ trace)
{
    extern __shared__ __align__(16) uint8_t flood_smem_raw[];
    FloodSmem& s = *reinterpret_cast<FloodSmem*>(flood_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = m.width, H = m.height, npx = W * H, ncells = m.num_cells;
    const size_t frame_bytes = (size_t)npx * 3;
    uint8_t* gray = ws_gray + (size_t)blockIdx.x * npx;
    uint8_t* gray2 = ws_gray2 + (size_t)blockIdx.x * npx;
    uint16_t* hsum = ws_hsum + (size_t)blockIdx.x * npx;
    uint32_t* raster = ws_raster + (size_t)blockIdx.x * (npx / 32 + 4);
    Heap heap; heap.sm = s.heap; heap.spill = ws_heap + (size_t)blockIdx.x * heap_cap; heap.n = 0;
    for (int i = tid; i < 256; i += kFloodThreads) s.adjust[i] = cx_adjust[i];
    if (tid < 16) s.tiles[tid] = cx_tiles_L[tid];

    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const bool need = force_all || (dirty[f] & kFrameDirtyK1);
        if (!need) { if (tid == 0) frame_flags[f] = 0; continue; }
        if (no_fallback) { if (tid == 0) frame_flags[f] = 0x2; continue; }   // CB200_FRAME_INEXACT
        const uint8_t* frame = rgb + (size_t)f * frame_bytes;
        __syncthreads();
        build_raster(m, frame, sharpen != 0, gray, gray2, hsum, raster);

        // ---- reset walk state (FloodDecodePositions::reset, FloodDecodePositions.cpp:17-42)
        for (int i = tid; i < ncells; i += kFloodThreads) s.instr[i] = pack_instr(0, 0, 0xFE, 0xFE);
        for (int i = tid; i < (ncells + 31) / 32; i += kFloodThreads) {
            int rem = ncells - i * 32;
            s.remaining[i] = rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u);
        }
        __syncthreads();

        if (warp == 0) {
            if (lane == 0) {
                heap.n = 0;
                int small_row = m.cells_x - 2 * m.corner, last = ncells - 1, bmb = m.top_cells;
                heap_push(heap, 0, 0); heap_push(heap, (uint32_t)(small_row - 1), 0);
                heap_push(heap, (uint32_t)last, 0); heap_push(heap, (uint32_t)(last - (small_row - 1)), 0);
                heap_push(heap, (uint32_t)bmb, 1); heap_push(heap, (uint32_t)(bmb + m.cells_x - 1), 1);
                heap_push(heap, (uint32_t)(last - bmb), 1); heap_push(heap, (uint32_t)(last - (bmb + m.cells_x - 1)), 1);
            }
            int count = 0;
            while (count < ncells) {
                // ---- FloodDecodePositions::next (FloodDecodePositions.cpp:49-67): lane 0 pops
                int ci = -1; uint32_t ins = 0;
                if (lane == 0) {
                    while (heap.n > 0) {
                        uint32_t e = heap_pop(heap);
                        int i = (int)(e & 0xFFFFu);
                        if (!is_remaining(s, i)) continue;
                        s.remaining[i >> 5] &= ~(1u << (i & 31));
                        ci = i; ins = s.instr[i];
                        break;
                    }
                }
                __syncwarp();   // lane 0's writes to remaining[] / the heap are ordered before the other lanes' reads below
                ci = __shfl_sync(0xffffffffu, ci, 0);
                ins = __shfl_sync(0xffffffffu, ins, 0);
                if (ci < 0) break;   // heap exhausted (cannot happen on a connected grid)
                ++count;
                const int ddx = (int)(int8_t)(ins & 0xFF), ddy = (int)(int8_t)((ins >> 8) & 0xFF);
                const uint32_t cooldown = ins >> 24;
                int px, py; cell_xy(m, ci, px, py);
                const int x = px + ddx, y = py + ddy;                 // CimbReader.cpp:146-148
                // ---- 10x10 window at (x-1, y-1): lane r < 10 fetches row r
                uint32_t myrow = 0;
                if (lane < 10) {
                    uint32_t bit = (uint32_t)(y - 1 + lane) * (uint32_t)W + (uint32_t)(x - 1);
                    uint32_t wi = bit >> 5;
                    myrow = __funnelshift_r(raster[wi], raster[wi + 1], bit & 31u) & 0x3FFu;   // bit i = window col i
                }
                uint32_t win[10];
#pragma unroll
                for (int r = 0; r < 10; ++r) win[r] = __shfl_sync(0xffffffffu, myrow, r);
                // ---- candidates (id order 4,5,7,3,1,8,0,2,6; tiles 0..15), key = dist<<8 | order<<4 | tile
                const bool all = (cooldown == 0xFEu);                 // CimbDecoder.cpp:144
                const int ncand = (all ? 9 : 5) * 16;
                uint32_t best_key = 0xFFFFFFFFu;
                for (int p = lane; p < ncand; p += 32) {
                    int q = p >> 4, t = p & 15;
                    int id = (0x620813754ULL >> (4 * q)) & 0xF;          // packed order table, nibble q
                    if ((uint32_t)id == cooldown && id != 4) continue; // CimbDecoder.cpp:116
                    int r0 = id / 3, c0 = id % 3;
                    uint32_t lo = 0, hi = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        lo |= ((win[r0 + k] >> c0) & 0xFFu) << (8 * k);
                        hi |= ((win[r0 + 4 + k] >> c0) & 0xFFu) << (8 * k);
                    }
                    unsigned long long L = ((unsigned long long)hi << 32) | lo;
                    uint32_t d = (uint32_t)__popcll(L ^ s.tiles[t]);
                    uint32_t key = (d << 8) | ((uint32_t)q << 4) | (uint32_t)t;
                    best_key = key < best_key ? key : best_key;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { uint32_t v = __shfl_xor_sync(0xffffffffu, best_key, o); best_key = v < best_key ? v : best_key; }
                {   // every lane derives the (warp-uniform) decision from the reduced key
                    const uint32_t dist = best_key >> 8, q = (best_key >> 4) & 0xFu, sym = best_key & 0xFu;
                    const int id = (int)((0x620813754ULL >> (4 * q)) & 0xF);
                    const int bx = id % 3 - 1, by = id / 3 - 1;             // CellDrift::driftPairs, CellDrift.h:13-15
                    const int ndx = clampi(ddx + bx, -7, 7), ndy = clampi(ddy + by, -7, 7);   // CellDrift.cpp:23-31
                    uint32_t ncd;                                      // CellDrift::calculate_cooldown, CellDrift.cpp:34-43
                    if (id == 4) ncd = 4; else if ((id & 1) == 0) ncd = 0xFF; else if (((cooldown ^ (uint32_t)id) & 0xFFu) == 6) ncd = 0xFF; else ncd = (uint32_t)id;
                    flood_update_warp(adj, s, heap, lane, ci, ndx, ndy, dist, ncd, ins);
                    if (lane == 0) {
                        s.instr[ci] = ((uint32_t)(x + bx) & 0x7FFu) | (((uint32_t)(y + by) & 0x7FFu) << 11) | (sym << 22);
                        if (trace) {
                            CellTrace tr;
                            tr.order = (uint16_t)(count - 1); tr.x = (int16_t)(x + bx); tr.y = (int16_t)(y + by);
                            tr.drift_offset = (uint8_t)id; tr.distance = (uint8_t)dist;
                            trace[(size_t)f * ncells + ci] = tr;
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- colours at the drift-adjusted positions (CimbReader::read_color, CimbReader.cpp:133-137)
        uint8_t* out = cellvals + (size_t)f * ncells;
        const int num_colors = 1 << m.color_bits;
        for (int i = tid; i < ncells; i += kFloodThreads) {
            uint32_t rec = s.instr[i];
            int x = (int)(rec & 0x7FFu), y = (int)((rec >> 11) & 0x7FFu);
            uint32_t sym = (rec >> 22) & 0xFu;
            uint32_t col = 0;
            if (num_colors > 1) {
                uint32_t R = 0, G = 0, B = 0;
                for (int r = 1; r <= 6; ++r) {
                    const uint8_t* p = frame + ((size_t)(y + r) * W + (size_t)(x + 1)) * 3;
                    for (int c = 0; c < 6; ++c) { R += p[3 * c]; G += p[3 * c + 1]; B += p[3 * c + 2]; }
                }
                col = flood_best_color(s.adjust, m, R / 36u, G / 36u, B / 36u);
            }
            out[i] = (uint8_t)(sym | (col << m.symbol_bits));
        }
        if (tid == 0) frame_flags[f] = 0x1;   // CB200_FRAME_FALLBACK
        __syncthreads();
    }
    (void)H;
}

// ---------------------------------------------------------------------------------------------- host side
cudaError_t flood_init_tables(const float* adjust256, const unsigned long long* tiles_L16)
{
    cudaError_t e = cudaMemcpyToSymbol(cx_adjust, adjust256, sizeof(float) * 256);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(cx_tiles_L, tiles_L16, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_flood, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FloodSmem));
}

cudaError_t flood_workspace_create(const Mode& m, int sm_count, const uint16_t* adj_host, FloodWorkspace* ws)
{
    memset(ws, 0, sizeof(*ws));
    ws->slots = 2 * sm_count;   // two resident CTAs per SM (108 KB shared memory each)
    size_t npx = (size_t)m.width * m.height;
    ws->heap_cap = 16 + 12 * (size_t)m.num_cells;   // spill area: every decoded cell pushes at most 12 entries (4 + 8 horizon)
    cudaError_t e;
    if ((e = cudaMalloc(&ws->gray, npx * ws->slots)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ws->gray2, npx * ws->slots)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ws->hsum, npx * ws->slots * sizeof(uint16_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ws->raster, (npx / 32 + 4) * ws->slots * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMemset(ws->raster, 0, (npx / 32 + 4) * ws->slots * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ws->heap, ws->heap_cap * ws->slots * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&ws->adj, (size_t)m.num_cells * 4 * sizeof(uint16_t))) != cudaSuccess) return e;
    if ((e = cudaMemcpy(ws->adj, adj_host, (size_t)m.num_cells * 4 * sizeof(uint16_t), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    return cudaSuccess;
}

void flood_workspace_destroy(FloodWorkspace* ws)
{
    cudaFree(ws->gray); cudaFree(ws->gray2); cudaFree(ws->hsum); cudaFree(ws->raster); cudaFree(ws->heap); cudaFree(ws->adj);
    memset(ws, 0, sizeof(*ws));
}

cudaError_t flood_launch(const Mode& m, const FloodWorkspace& ws, const uint8_t* d_rgb, int n_frames, bool no_fallback,
                         bool force_all, bool sharpen, uint8_t* d_cellvals, const uint32_t* d_dirty, uint8_t* d_flags, CellTrace* d_trace,
                         cudaStream_t st)
{
    int grid = n_frames < ws.slots ? n_frames : ws.slots;
    k_flood<<<grid, kFloodThreads, sizeof(FloodSmem), st>>>(m, d_rgb, n_frames, no_fallback ? 1 : 0, force_all ? 1 : 0, sharpen ? 1 : 0,
                                                            d_cellvals, d_dirty, d_flags, ws.gray, ws.gray2, ws.hsum, ws.raster, ws.heap, ws.heap_cap,
                                                            reinterpret_cast<const ushort4*>(ws.adj), d_trace);
    return cudaGetLastError();
}

}  // namespace cb200
