// Entry points shared between fd_api.cu and fd_tiled.cu inside libfd_b200.so (not part of the C ABI).
#pragma once

#include <cstdint>

#include "../../include/fd_b200.h"

// Row tiles: selection over the tiles' FIRST rank ranges only.  All pointers are device memory on the context's device.
struct fd_select_prefilter {
    const uint32_t *total_counts;   // per frame: candidates of all tiles together
    uint32_t total_capacity;        // slots per frame of the buffer the full gather would fill
    uint32_t *hist;                 // n_frames x 2048: the tiles' rank histograms summed
    uint64_t *pre_keys;             // n_frames slots of pre_capacity keys: the tiles' first ranges, gathered
    uint32_t *pre_counts;
    uint32_t pre_capacity;
    uint8_t *need_more;             // out: 1 for frames that need their full key slot (zero on entry)
};
// Frames whose first range yields enough keypoints are finished; the others are flagged in need_more and left untouched.
fd_status fd_internal_select_first_range(fd_context *ctx, const fd_detect_params *params, const fd_select_prefilter *pf, int rows, int cols, int n_frames);
// The ordinary selection over full key slots, for the frames whose flag is set only (results of the other frames stay as they are).
fd_status fd_internal_select_flagged(fd_context *ctx, const fd_detect_params *params, uint64_t *dev_keys, const uint32_t *dev_counts, uint32_t capacity,
                                     const uint8_t *flags, int rows, int cols, int n_frames);
