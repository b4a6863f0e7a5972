// ctx.cuh -- the context object behind the C ABI (include/cb200.h) and the error helpers, shared by api.cu / gather.cu / deskew.cu
#pragma once
#include "../../include/cb200.h"
#include "cb200_common.cuh"
#include "k1x_flood.cuh"
#include "ccm.cuh"

#include <string>

namespace cb200 {
// thread-local error text behind cb200_last_error(); both return `code`
int fail(int code, const std::string& msg);
int fail_cuda(cudaError_t e, const char* what);
struct GatherState;      // gather.cu
struct DeskewState;      // deskew.cu
struct ScanScratch;      // scan.cu
void gather_destroy(GatherState* g);
void deskew_destroy(DeskewState* d);
void scan_destroy(ScanScratch* s);
}  // namespace cb200
#define CK(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cb200::fail_cuda(e__, what); } while (0)

struct cb200_ctx {
    cb200::Mode mode;
    int device = 0, max_frames = 0, sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device workspaces
    uint8_t* d_rgb = nullptr;        // host-pointer entry points only: max_frames frames
    uint8_t* d_cellvals = nullptr;   // max_frames * num_cells
    uint32_t* d_dirty = nullptr;     // max_frames
    uint8_t* d_raw = nullptr;        // max_frames * cap_all
    uint8_t* d_data = nullptr;       // max_frames * data_bytes
    uint8_t* d_ok = nullptr;         // max_frames * nblocks
    uint32_t* d_mask = nullptr;      // max_frames
    uint8_t* d_flags = nullptr;      // max_frames
    uint16_t* d_idx = nullptr;       // num_cells: slot -> cell (Interleave::interleave_indices)
    uint16_t* d_inv = nullptr;       // num_cells: cell -> slot (Interleave::interleave_reverse)
    uint16_t* d_idx_ident = nullptr; // identity map for CB200_FLAG_NO_INTERLEAVE (created on first use)
    uint8_t* d_gen = nullptr;        // RS generator polynomial, ecc_bytes+1 coefficients
    uint8_t* d_rho = nullptr;        // 4 x 64 bytes: x^(D+j) mod x^pad g, the basis of K2's remainder tables (k2_remainder_basis)
    // per-kernel timing (cb200_set_timing): events around every launch of the last pipeline call
    int l2_ahead = 0;                // K1: TMA L2-prefetch distance in stages (CB200_K1_L2_AHEAD overrides, tuning only)
    int k1_ctas_per_sm = 4;          // K1: resident CTAs per SM the grid is sized for (CB200_K1_CTAS_PER_SM, tuning only)
    bool timing = false;
    static constexpr int kEvSets = 64;
    cudaEvent_t ev[kEvSets][8] = {};
    int ev_count[kEvSets] = {};
    long calls = 0;                  // pipeline calls since timing was enabled
    int cur = 0;                     // event set of the call in progress
    cb200::FloodWorkspace flood;            // exact-walk fallback scratch
    // small scratch for the single-cell entry points
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    // pinned host staging for results of the host-pointer entry points
    uint8_t* h_pinned = nullptr; size_t h_pinned_bytes = 0;
    // colour correction (the reference's thread-local CimbDecoder CCM, CimbDecoder.cpp:69-85)
    float ccm[9] = {};               // active matrix, row-major
    bool ccm_active = false;
    bool ccm_pending = false;        // the last CC_SIMPLE batch's final matrix is still on its way to h_ccm
    bool ccm_pending_flag = false;   // ... and so is whether that frame had a CCM at all (CC_FIT batches)
    float* d_ccm = nullptr;          // per-frame matrices of a CC_SIMPLE / CC_FIT batch (the ones used): max_frames x 9
    float* h_ccm = nullptr;          // pinned: 9 floats + 1 activity byte (at float index 9)
    // CC_FIT scratch: per-cell mean colours of the first pass, per-frame fits
    uint32_t* d_means = nullptr;     // max_frames x num_cells
    float* d_fit = nullptr;          // max_frames x 9
    uint8_t* d_fit_valid = nullptr;  // max_frames
    uint8_t* d_ccm_active = nullptr; // max_frames: the frame is decoded with d_ccm[f]
    cudaEvent_t ccm_ev = nullptr;    // recorded after the D2H copies of the last batch's CCM (ccm_resolve waits on it)
    cb200::GatherState* gather = nullptr;   // multi-GPU chunk-record window (gather.cu)
    cb200::DeskewState* deskew = nullptr;   // extractor scratch (deskew.cu)
    cb200::ScanScratch* scan = nullptr;     // anchor-scan scratch (scan.cu)
};
