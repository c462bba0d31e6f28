// Internal launch interface between the C-ABI layer (fd_api.cu) and the kernels.
#pragma once

#include <algorithm>

#include "fd_common.cuh"

namespace fdb {

// ---- mask of pre-existing features (feature_point_detector.cpp:76-98), fd_mask.cu ----------------
// One bit per pixel, words_per_row = ceil(cols / 32) + 1 words per row (the spare word lets a kernel read the
// word pair that straddles any 4-pixel group without a bounds test).  bits == nullptr means "no mask".
struct MaskView {
    const uint32_t *bits;         // n_frames * rows * words_per_row
    const uint32_t *word_prefix;  // FAST only: masked-in interior pixels of the row before each word
    const uint32_t *row_base;     // FAST only: n_frames * (rows + 1); masked-in interior pixels in rows < r
    int words_per_row;
};
struct MaskArgs {
    int rows, cols, n_frames;
    int min_distance;
    const float *xy;              // n_frames slots of `capacity` (x, y) pairs
    const int32_t *counts;
    int capacity;
    uint32_t *bits, *word_prefix, *row_base;  // word_prefix / row_base may be null (not FAST)
    int words_per_row;
};
cudaError_t launch_mask(const MaskArgs &args, cudaStream_t stream);

// ---- kernel 2: FAST --------------------------------------------------------------------------
constexpr int FAST_THREADS = 256;
constexpr int FAST_CTAS_PER_SM = 2;
constexpr int FAST_MAX_SEGS = 96;     // linear pieces of the offset table kept in shared memory (incl. sentinel)
constexpr int FAST_STAGE_KEYS = 256;  // per-warp candidate staging buffer

// One linear piece of the reference's running float offset (fast.cpp:85,93): for masked-in pixel
// index k in [k_start, next.k_start) the offset's BIT PATTERN is bits_start + (k - k_start) * step.
struct OffsetSeg {
    uint32_t k_start, bits_start, step;
};

struct FastArgs {
    FrameView fv;
    int diff;                 // kMinPixelDiffValue
    float thr;                // kMinValidResponse
    const uint8_t *lut;       // 65536-entry longest-circular-run table (device)
    const OffsetSeg *segs;    // n_seg pieces + one sentinel (k_start = 0xFFFFFFFF)
    int n_seg;
    uint32_t kmin[17];        // first pixel index at which score s yields response > thr (0xFFFFFFFF = never)
    uint64_t *cand_keys;      // n_frames slots of cand_capacity keys
    uint32_t *cand_counts;    // n_frames
    uint32_t cand_capacity;
    uint8_t *score_map;       // optional dense score map (n_frames * rows * cols), may be null
    int score_aligned;        // score_map rows can be written as aligned words
    int n_strips, n_bands, band_rows;
    int64_t n_items;
    MaskView mask;            // pre-existing features (bits == nullptr: every pixel masked in)
    TileView tile;            // row tile of a larger frame (untiled: {0, 0, rows, rows})
    int proc_lo, proc_hi;     // local rows that get scored: [max(3, own_lo), min(rows - 3, own_hi, full_rows - 3 - row_offset))
    uint32_t *work_counter;   // next work item to hand out, zero on entry
    uint32_t absdiff_mask;    // sparse kernel: per byte, the bits at or above 2^absdiff_shift
    int absdiff_shift;        // 2^absdiff_shift = largest power of two <= diff + 1
};
size_t fast_smem_bytes(int n_seg);
cudaError_t launch_fast(const FastArgs &args, bool precheck, int grid, cudaStream_t stream);
// Sparse form (fd_fast_sparse.cu): needs a 3-D TMA map of the frames (a CUtensorMap, 128 bytes, passed opaquely),
// no mask, no score map, and a threshold that leaves s_min >= 4 (>= 1 with the pre-check) over the whole frame.
constexpr int FAST_SPARSE_THREADS = 768;
constexpr int FAST_SPARSE_GROUP_ROWS = 8;   // rows per TMA box (the host builds the tensor map with this box height)
size_t fast_sparse_smem_bytes();
cudaError_t launch_fast_sparse(const FastArgs &args, const void *tensor_map, bool precheck, int grid, cudaStream_t stream);

// ---- kernel 1: Harris / Shi-Tomasi -------------------------------------------------------------
constexpr int CORNER_THREADS = 256;
constexpr int CORNER_STRIP_OUT = 126;  // NMS-complete columns per 128-column warp strip

struct CornerArgs {
    FrameView fv;
    int kind;                 // 0 Harris, 1 Shi-Tomasi (reference formula: larger eigenvalue)
    float thr;                // kMinValidResponse
    float alpha;              // Harris kAlpha
    float inv_cnt, inv_cnt2;  // 1/9 and its square, rounded as the reference rounds them (harris.cpp:71-72)
    float harris_trace_min;   // Harris: the smallest trace that passes the pre-test of harris.cpp:98 (NaN: none does); see fd_api.cu
    uint64_t *cand_keys;
    uint32_t *cand_counts;
    uint32_t cand_capacity;
    float *response_map;      // optional dense thresholded response map, may be null
    int n_strips, n_bands, band_rows;
    int64_t n_items;
    MaskView mask;            // pre-existing features (bits == nullptr: every pixel masked in)
    TileView tile;            // row tile of a larger frame (untiled: {0, 0, rows, rows})
    uint32_t *work_counter;   // next work item to hand out, zero on entry
    int resp_lo, resp_hi;     // local rows with a defined response, inclusive (harris.cpp:90-92 on the FULL frame)
    int cand_lo, cand_hi;     // local rows that may emit candidates: [cand_lo, cand_hi)
};
cudaError_t launch_corner(const CornerArgs &args, int grid, cudaStream_t stream);
// TMA form (fd_corner_tma.cu): needs a 3-D TMA map of the frames with a 160 x CORNER_TMA_GROUP_ROWS x 1 box, and no mask.
#ifndef FD_CORNER_TMA_THREADS
#define FD_CORNER_TMA_THREADS 512   // 640 / 768 (102 / 85 registers per thread) were measured on B200: see DESIGN.md section 8
#endif
constexpr int CORNER_TMA_THREADS = FD_CORNER_TMA_THREADS;
constexpr int CORNER_TMA_GROUP_ROWS = 12;
size_t corner_tma_smem_bytes();
cudaError_t launch_corner_tma(const CornerArgs &args, const void *tensor_map, int grid, cudaStream_t stream);

// ---- kernel 3: per-frame sort + greedy min-distance selection ----------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_SMEM_MAX_KEYS = 8192;  // 64 KiB of keys per CTA -> three sorting CTAs per SM
#ifndef FD_SELECT_THREADS
#define FD_SELECT_THREADS 512
#endif
constexpr int SELECT_THREADS = FD_SELECT_THREADS;        // per-frame CTA when the cell grid fits shared memory (256 and 1024 threads measured: slower overall)
constexpr int SELECT_MAX_THREADS = 1024;   // ... and when it lives in global memory (very fine grids: thousands of cells per round)
#ifndef FD_SELECT_RANK_COUNT_MAX
#define FD_SELECT_RANK_COUNT_MAX 256
#endif
constexpr int SELECT_RANK_COUNT_MAX = FD_SELECT_RANK_COUNT_MAX;   // kept points ordered by counting smaller keys up to this many, by a bitonic network beyond
constexpr int SELECT_SORT_SMEM = 1024;    // kept points sorted in shared memory up to this many
constexpr int SELECT_CELLS_MIN = 65536;   // frames with more candidates than this group them by cell and run the rounds per cell
constexpr int SELECT_PREFIX_MIN = 8192;   // frames with more candidates than this run the rounds on a rank prefix first
#ifndef FD_SELECT_PREFIX_FIRST
#define FD_SELECT_PREFIX_FIRST 2048
#endif
constexpr int SELECT_PREFIX_FIRST = FD_SELECT_PREFIX_FIRST;   // ... of at least this many candidates (or 8 per wanted point); measured on B200: 2048 beats 4096 by 20 % on Harris at 752x480

struct SelectArgs {
    int rows, cols, n_frames;
    uint64_t *cand_keys;            // in: unsorted; out: sorted ascending (= response descending, raster ties)
    const uint32_t *cand_counts;
    uint32_t cand_capacity;
    int min_distance;
    uint32_t needed;
    uint32_t cells_min;             // frames with more candidates than this run the rounds per cell (SELECT_CELLS_MIN)
    const int32_t *existing_counts; // per frame, may be null (= 0)
    float4 *keypoints;              // n_frames slots of kp_capacity (x, y, response, 0)
    int32_t *kp_counts;
    int kp_capacity;
    uint64_t *live_scratch;         // n_frames * 2 * cand_capacity: the live-candidate key lists of two consecutive rounds
    uint64_t *kept_keys;            // n_frames slots of kept_capacity keys (the kept set before the cut)
    int kept_capacity;              // = number of grid cells (at most one kept point per cell)
    uint32_t *cell_scratch;         // global fallback for the per-cell state (select_cell_bytes() per frame), may be null
    int64_t cell_stride;            // bytes between the per-cell states of consecutive frames in cell_scratch
    int cells_x, cells_y;           // grid of (min_distance+1)-sided cells
    uint32_t cell_magic;            // ceil(2^32 / (min_distance+1))
    int cells_in_smem;
    int few_frames;                 // no more frames than SMs: the launch is latency-bound (one CTA or cluster per SM group), not throughput-bound
    uint32_t *overflow_flag;        // set to 1 if any frame's candidate count exceeded cand_capacity
    MaskView mask;                  // candidates on masked-out pixels are never accepted (feature_point_detector.cpp:66)
    uint32_t *pre_hist;             // optional (few frames, many candidates): n_frames * 2048 rank-histogram bins, zero on entry ...
    uint64_t *pre_keys;             // ... n_frames slots of cand_capacity keys for the first rank range ...
    uint32_t *pre_counts;           // ... and their fill, zero on entry (select_hist_kernel / select_admit_kernel)
    uint32_t pre_capacity;          // slots per frame of pre_keys
    // row tiles (fd_tiled.cu): the histogram is built per tile whatever its count, the first range's limit is handed to the tiles ...
    int hist_always;                // select_hist_kernel: also frames of SELECT_PREFIX_MIN candidates or fewer
    const uint64_t *ext_limits;     // select_admit_kernel: per frame, admit keys below this (0: leave the frame alone) instead of deriving it
    // ... and the selection runs on the gathered first ranges alone; a frame that needs more than that is flagged, not finished
    uint8_t *need_more;             // select_kernel, first-range mode: set to 1 for frames it could not finish (untouched otherwise)
    const uint8_t *only_flagged;    // select_kernel: skip frames whose flag is 0
    uint32_t xy_xor;                // 0, or 0xFFFFFFFF when the keys carry the complemented position (NN heat maps: among equal responses the later pixel first)
};
size_t select_cell_bytes(int cells_x, int cells_y);
cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream);
// The two preparation kernels on their own (fd_tiled.cu runs them on the tiles' devices): grid = chunks x n_frames.
cudaError_t launch_select_hist(const SelectArgs &args, cudaStream_t stream);
cudaError_t launch_select_admit(const SelectArgs &args, cudaStream_t stream);

// ---- kernel 4: steered BRIEF --------------------------------------------------------------------
struct BriefArgs {
    FrameView fv;
    const float4 *keypoints;   // slots of kp_capacity per frame: (x, y, *, *)
    const int32_t *kp_counts;
    int kp_capacity;
    int length, half_patch, sampling;
    int integral_keypoints;    // 1: the keypoints are detector output (integer coordinates): the leaner instantiation applies
    uint8_t *desc;             // 32 bytes per keypoint slot
};
cudaError_t launch_brief(const BriefArgs &args, cudaStream_t stream);
// The std::vector<Vec> overload of Descriptor::Compute (descriptor.h:43-62): packed bits -> +1 / -1 floats, `length` per keypoint slot.
cudaError_t launch_brief_to_float(const uint8_t *desc, const int32_t *kp_counts, int kp_capacity, int n_frames, int length, float *out, cudaStream_t stream);

// ---- Hamming matching of packed descriptors (no reference counterpart), fd_match.cu -------------------------
struct MatchArgs {
    const uint8_t *desc_a, *desc_b;       // descriptor sets: slots of capacity x 32 bytes
    const int32_t *counts_a, *counts_b;   // valid descriptors per set
    int capacity_a, capacity_b;
    int n_pairs;                          // pair i matches set i * stride_a_sets of A against set i * stride_b_sets + offset_b_sets of B
    int stride_a_sets, stride_b_sets, offset_b_sets;
    int4 *out;                            // n_pairs x capacity_a: (train index or -1, distance, second distance or -1, 0)
};
cudaError_t launch_match(const MatchArgs &args, cudaStream_t stream);

// ---- kernel 5: LSD gradient / level-line field --------------------------------------------------
constexpr int LSD_THREADS = 256;
constexpr int LSD_MAX_M = 2 * 255 * 255;   // largest ad^2 + bc^2 (feature_line_detector.cpp:76-82 on 8-bit pixels)
constexpr int LSD_BINS = LSD_MAX_M + 1;    // one bin per attainable gradient norm
struct LsdArgs {
    FrameView fv;
    float min_norm;
    float *norm;               // n_frames * (rows-1) * (cols-1)
    float *angle;
    uint64_t *seed_keys;       // optional: one region of band_rows * 128 keys per work item: (m << 32) | (col << 16) | row, m = ad^2 + bc^2
    uint32_t *item_counts;     // with seed_keys: keys in each work item's region
    uint32_t *seed_counts;     // with seed_keys: valid pixels per frame, zero on entry
    uint32_t *seed_hist;       // with seed_keys: n_frames * LSD_BINS counters, zero on entry (the scatter returns them to zero)
    int n_bands, band_rows;
    int64_t n_items;
    uint32_t *work_counter;    // next work item to hand out, zero on entry
};
cudaError_t launch_lsd(const LsdArgs &args, int grid, cudaStream_t stream);
// Seed order by exact magnitude binning: scan of the histogram, scatter into buckets, order inside buckets; writes the
// valid pixels of every frame as (row * cols + col) indices, norm descending, ties in the reference's push order.
constexpr int LSD_SEED_ORDER_LAUNCHES = 4;
size_t lsd_chunk_sum_bytes(int n_frames);
cudaError_t launch_seed_order(const LsdArgs &args, uint64_t *bucketed, uint32_t *start, uint32_t *chunk_sum, int32_t *sorted_idx, cudaStream_t stream);

// ---- NN detector post-processing (nn_feature_point_detector.cpp:59-72, 128-155, 163-193), fd_nn.cu --------------------
struct NnHeatmapArgs {
    const float *heatmap;      // n_frames * rows * cols, row-major, no pitch
    int rows, cols, n_frames;
    float min_response;        // kMinResponse
    int invalid_boundary;      // kInvalidBoundary
    uint64_t *cand_keys;       // high word ~ordered(response), low word ~((row << 16) | col): equal responses rank the later pixel first
    uint32_t *cand_counts;
    uint32_t cand_capacity;
    uint32_t *work_counter;    // next run of rows to hand out, zero on entry
};
cudaError_t launch_nn_heatmap(const NnHeatmapArgs &args, int sm_count, cudaStream_t stream);
struct NnDescriptorArgs {
    const float *maps;         // n_frames * channels planes of map_rows x map_cols floats
    int channels, map_rows, map_cols, n_frames;
    const float4 *keypoints;   // slots of kp_capacity per frame: (x, y, *, *)
    const int32_t *kp_counts;
    int kp_capacity;
    float *out;                // channels floats per keypoint slot
};
cudaError_t launch_nn_descriptors(const NnDescriptorArgs &args, cudaStream_t stream);

// ---- shared: segmented key sort (one CTA per segment) -------------------------------------------
cudaError_t launch_segment_sort(uint64_t *keys, const uint32_t *counts, int64_t slot, int n_segments, uint32_t capacity, uint32_t *overflow_flag,
                                cudaStream_t stream);

}  // namespace fdb
