// encode.cu -- device-side input generator: payload bytes -> RS-encoded, interleaved cell values (benchmark support,
// the inverse of the decode path; not on the hot path).  Restates Encoder::encode_next / encode_next_coupled
// (reference: src/lib/encoder/Encoder.h:69-129, :131-165), reed_solomon_stream::readsome (reed_solomon_stream.h:32-52)
// -> correct_reed_solomon_encode (src/third_party_lib/libcorrect/src/reed-solomon/encode.c:3-35) and the interleaved
// CimbWriter position order (src/lib/cimb_translator/CimbWriter.cpp:40-44, Interleave.h:8-24).
#include "cb200_common.cuh"
#include "encode.cuh"

namespace cb200 {

__constant__ uint8_t ce_gf_exp[512];
__constant__ uint8_t ce_gf_log[256];

// systematic RS encode of one block per thread: remainder of msg(x) * x^parity mod g(x), message high-order first
__global__ void __launch_bounds__(128)
k_rs_encode(const Mode m, const uint8_t* __restrict__ gen /* parity+1 coefficients, low -> high */,
            const uint8_t* __restrict__ payload, int n_frames, uint8_t* __restrict__ raw)
{
    long gb = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long total = (long)n_frames * m.nblocks;
    if (gb >= total) return;
    int f = (int)(gb / m.nblocks), b = (int)(gb - (long)f * m.nblocks);
    const uint8_t* msg = payload + (size_t)f * m.data_bytes + (size_t)b * m.msg_len;
    uint8_t* out = raw + (size_t)f * m.cap_all + (size_t)b * m.ecc_block;
    const int md = m.ecc_bytes;
    uint8_t rem[64];
    for (int j = 0; j < md; ++j) rem[j] = 0;
    for (int i = 0; i < m.msg_len; ++i) {
        uint8_t v = msg[i];
        out[i] = v;
        uint32_t fb = v ^ rem[md - 1];
        uint32_t lfb = ce_gf_log[fb];
        for (int j = md - 1; j > 0; --j) {
            uint32_t g = gen[j];
            uint32_t p = (fb && g) ? (uint32_t)ce_gf_exp[lfb + ce_gf_log[g]] : 0u;
            rem[j] = (uint8_t)(rem[j - 1] ^ p);
        }
        uint32_t g0 = gen[0];
        rem[0] = (uint8_t)((fb && g0) ? (uint32_t)ce_gf_exp[lfb + ce_gf_log[g0]] : 0u);
    }
    for (int j = 0; j < md; ++j) out[m.msg_len + j] = rem[md - 1 - j];
}

// MSB-first read of `len` bits at bit offset `pos` (bitbuffer::read, src/lib/bit_file/bitbuffer.h:86-107)
__device__ __forceinline__ uint32_t read_bits(const uint8_t* __restrict__ buf, uint32_t pos, int len)
{
    uint32_t byte = pos >> 3, off = pos & 7u;
    uint32_t w = ((uint32_t)buf[byte] << 16) | ((uint32_t)buf[byte + 1] << 8) | (uint32_t)buf[byte + 2];
    return (w >> (24 - off - len)) & ((1u << len) - 1u);
}

// one thread per cell: value = (colour << symbol_bits) | symbol of the slot this cell occupies in the interleaved order
__global__ void __launch_bounds__(256)
k_unpack_cells(const Mode m, const uint8_t* __restrict__ raw, const uint16_t* __restrict__ inv, int n_frames,
               uint8_t* __restrict__ cellvals)
{
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    int f = blockIdx.y;
    if (cell >= m.num_cells || f >= n_frames) return;
    const uint8_t* r = raw + (size_t)f * m.cap_all;   // cap_all + 2 bytes readable: workspace is padded
    uint32_t s = inv[cell];
    uint32_t v;
    if (m.legacy) {
        int bpc = m.symbol_bits + m.color_bits;
        v = read_bits(r, s * (uint32_t)bpc, bpc);
    } else {
        uint32_t sym = read_bits(r, s * (uint32_t)m.symbol_bits, m.symbol_bits);
        uint32_t col = read_bits(r + m.cap_sym, s * (uint32_t)m.color_bits, m.color_bits);
        v = (col << m.symbol_bits) | sym;
    }
    cellvals[(size_t)f * m.num_cells + cell] = (uint8_t)v;
}

cudaError_t encode_init_tables(const uint8_t* exp512, const uint8_t* log256)
{
    cudaError_t e = cudaMemcpyToSymbol(ce_gf_exp, exp512, 512);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(ce_gf_log, log256, 256);
}

cudaError_t encode_launch(const Mode& m, const uint8_t* d_gen, const uint16_t* d_inv, const uint8_t* d_payload, int n_frames,
                          uint8_t* d_raw, uint8_t* d_cellvals, cudaStream_t st)
{
    long total = (long)n_frames * m.nblocks;
    k_rs_encode<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(m, d_gen, d_payload, n_frames, d_raw); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid((m.num_cells + 255) / 256, n_frames);
    k_unpack_cells<<<grid, 256, 0, st>>>(m, d_raw, d_inv, n_frames, d_cellvals); count_launch();
    return cudaGetLastError();
}

}  // namespace cb200
