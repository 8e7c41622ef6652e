// orb.cuh -- internal definition of the extractor handle (shared by orb.cu and stereo.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace adb {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;        // EDGE_THRESHOLD   src/ORBextractor.cc:75
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD-3 src/ORBextractor.cc:775
constexpr int kCellBoxWMax = 96; // FAST cell tile: 15 B alignment slack + cell + 6 px, rounded up to 16 B
                                 // (TMA needs the box origin 16-byte aligned in x: measured on B200, tools/probe)
constexpr int kCellBoxHMax = 66;
constexpr int kPatchBoxW = 64;   // descriptor boxes: 64 B rows = 43 px (key-point column at byte 21 + alignment slack <= 15) ...
constexpr int kPatchR = 21;      // ... fetched from the 16-B aligned column at or left of x - 21

// Per-level constants, one array per handle in device memory.
struct LevelDev {
    int w, h, pitch;
    unsigned frame_stride;   // bytes between frames of this level
    int mpitch; unsigned mframe_stride;   // mask pyramid layout (always the handle's own buffers)
    int ncols, nrows, wcell, hcell;
    int cell_base, ncells;   // position in the flattened cell table
    int slotcap;             // candidate slots per cell = ceil(wcell/2) * ceil(hcell/2) (3x3 strict-NMS bound)
    int cand_base, cand_cap; // u32 offset of this level's slots inside one frame's candidate area
    int quota;               // mnFeaturesPerLevel
    int list_base, list_cap; // position of this level's kept list inside one frame's list area
    int n_ini;               // quad-tree roots
    float hx;                // root width
    float scale;             // mvScaleFactor
    float inv_scale;
    int patch_size;          // (int)(31 * scale)
    int box_w, box_h;        // FAST cell TMA box
};

struct BlurLevels {                       // all levels in ONE launch: warp tasks are numbered level by level
    const uint8_t* src[kMaxLevels];
    uint8_t* dst[kMaxLevels];
    int w[kMaxLevels], h[kMaxLevels], spitch[kMaxLevels], dpitch[kMaxLevels], ncg[kMaxLevels], task_end[kMaxLevels];
    unsigned sfstride[kMaxLevels], dfstride[kMaxLevels];
    int nlevels, ntasks;
};

struct LevelHost {
    LevelDev d;
    uint8_t* img = nullptr;    // [max_batch][h][pitch]   (level 0: internal staging copy)
    uint8_t* mask = nullptr;   // same layout, allocated on first masked call
    uint8_t* blur = nullptr;   // GaussianBlur 7x7 of the level, [max_batch][h][(w + 15) & ~15] (always the handle's own buffer)
    int2* xtab = nullptr;      // resize tables for producing this level from level-1: {src index, a0 | a1 << 16}
    int2* ytab = nullptr;
    bool strip_ok = false;     // pyr_resize_strip_kernel applies (scale factor <= 2)
};

}  // namespace adb

struct adb_orb {
    adb_orb_config cfg;
    int nlevels = 0;
    int capacity = 0;               // key-points per frame
    int ncells_total = 0;
    int cell_box_w = 64;                // shared-memory pitch of the FAST cell box (64 or kCellBoxWMax), one per handle
    int fast_tile_bytes = 0;            // warp-per-cell FAST: bytes of one warp's TMA box (largest level), a multiple of 128
    bool fast_warp_ok = false;
    int n_narrow_cells = 0;             // d_cell_table[ncells_total ..]: cell order of the warp kernel, this many single-tile cells first
    int cand_total = 0;             // u32 entries per frame
    int list_total = 0;             // entries per frame
    int qt_maxa = 0;                // quad-tree node capacity
    size_t qt_smem = 0;
    std::vector<adb::LevelHost> lv;
    uint8_t* mask_stage = nullptr;  // [max_batch][h][pitch0] caller's masks of a host-buffer call, before the erosion
    std::vector<float> sigma2, inv_sigma2;
    bool lazy = false, provisioned = false;   // created with width = height = 0: provisioned (and re-provisioned) from the image of each call
    std::vector<float> lazy_scale;            // mvScaleFactor before the first frame
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;   // chunked host-buffer calls: upload / download streams
    cudaEvent_t cev[65] = {};                                  // [0..31] chunk uploaded, [32..63] chunk computed, [64] entry fence
    cudaEvent_t sev[32] = {};                                  // stereo chunk pipeline: chunk matched (left handle)
    cudaStream_t stereo_stream = nullptr;                      // ... and the stream its matcher runs on
    adb::LevelDev* d_levels = nullptr;
    uint32_t* d_cell_table = nullptr;   // [ncells_total] level << 24 | row << 12 | col, then [ncells_total] the warp kernel's cell order
    float4* d_pattern = nullptr;        // [8][32] rBRIEF tests as floats {x0, y0, x1, y1}, lane-major
    uint32_t* d_cand = nullptr;         // [max_batch][cand_total] per-cell slots
    uint16_t* d_cellcnt = nullptr;      // [max_batch][ncells_total]
    uint32_t* d_qkeys = nullptr;        // [max_batch][cand_total] gathered candidates (reference order)
    uint32_t* d_qstate = nullptr;       // [max_batch][cand_total] node slot << 20 | node seq
    int32_t* d_qcount = nullptr;        // [max_batch][nlevels] candidates per (frame, level)
    uint32_t* d_list = nullptr;         // [max_batch][list_total] kept key-points per level
    int32_t* d_listcnt = nullptr;       // [max_batch][nlevels]
    int32_t* d_status = nullptr;        // device-side error flag
    adb_keypoint* d_kps = nullptr;      // [max_batch][capacity]
    uint8_t* d_desc = nullptr;          // [max_batch][capacity][32]
    int32_t* d_counts = nullptr;        // [max_batch]
    // stereo outputs (owned by the left handle)
    float* d_uright = nullptr;
    float* d_depth = nullptr;
    int32_t* d_best_idx = nullptr;
    int32_t* d_best_dist = nullptr;
    int32_t* d_sad = nullptr;
    int32_t* d_row_ptr = nullptr;     // stereo row buckets of the right key-points (CSR over image rows)
    uint16_t* d_row_items = nullptr;
    void* d_rinfo = nullptr;          // float2 {x, octave} per right key-point
    // level-0 view of the last call (caller's buffer when it is TMA-addressable, else lv[0].img)
    const uint8_t* l0_base = nullptr;
    int l0_pitch = 0;
    size_t l0_fstride = 0;
    bool have_mask = false;
    int last_frames = 0;
    adb::TmaMaps16 cell_maps;    // FAST cell boxes
    adb::TmaMaps16 patch_maps;   // descriptor patches
    adb::TmaMaps16 blur_maps;    // the same boxes of the blurred levels
    adb::BlurLevels blur_levels = {};   // geometry of the one-launch level blur
    // profiling: CUDA events on the handle's stream around each stage of the last call
    adb_gather_targets gather = {};   // peer buffers the descriptor kernel also writes to (n = 0: off)
    bool profiling = false;
    cudaEvent_t pev[8] = {};
    int pev_n = 0;
    long long launches = 0;         // kernels launched by this handle since creation
    int32_t* h_counts = nullptr; // pinned
    int32_t* h_status = nullptr; // pinned
};

// internal (C++ linkage), match.cu: stereo matching / result download of a frame range on a given stream
adb_status adb_stereo_match_range(adb_orb* L, adb_orb* R, int f0, int n, float mb, float mbf, cudaStream_t st);
adb_status adb_stereo_download_range(adb_orb* L, int f0, int n, float* ur, float* dp, int32_t* bi, int32_t* bd, int cap, cudaStream_t st);
