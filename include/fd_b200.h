/* fd_b200.h -- C ABI of the B200-native dense feature-detection path.
 *
 * The reference (Horizon1026/Feature_Detector) has no plugin / FFI layer: its boundary is the public
 * C++ class API, all entry points non-virtual.  "Drop-in" therefore means link-time
 * substitution: the C++ classes under feature_detector_b200/cpp/ keep the reference's names,
 * namespaces, Options structs and accessors, and their bodies call THIS C ABI.  Each group of entry
 * points below cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every call returns an fd_status (0 = OK); fd_last_error(ctx) gives the text of the last failure;
 *   - there is NO CPU fallback anywhere behind this header: without a CUDA device fd_create fails;
 *   - one context = one device + one stream; calls are asynchronous on that stream unless they copy to
 *     pageable host memory or are documented as synchronising; a context is not re-entrant;
 *   - "frames" are 8-bit grayscale images.  Host frames are pitch-less and contiguous like the
 *     reference's GrayImage (feature_point_harris_detector.cpp:31 `data + r * cols`); device frames may
 *     carry a row pitch and a frame stride;
 *   - per-frame outputs live in fixed-capacity slots so that a batch needs no host round trip:
 *     slot f of an array with capacity C starts at element f*C, and counts[f] says how many are valid.
 */
#ifndef FD_B200_H_
#define FD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fd_context fd_context;

typedef enum fd_status {
    FD_OK = 0,
    FD_ERR_INVALID_ARGUMENT = 1,
    FD_ERR_NO_DEVICE = 2,       /* no usable CUDA device: there is no CPU path to fall back to */
    FD_ERR_CUDA = 3,            /* a CUDA runtime call or kernel failed; see fd_last_error */
    FD_ERR_OUT_OF_MEMORY = 4,
    FD_ERR_CAPACITY = 5,        /* a per-frame slot overflowed (candidates); raise the capacity and retry */
    FD_ERR_NOT_READY = 6        /* a stage was asked for before the stage that feeds it */
} fd_status;

/* Detector kinds.  SHI_TOMAS reproduces the reference, which returns the LARGER eigenvalue
 * (feature_point_shi_tomas_detector.cpp:94-100). */
typedef enum fd_detector_kind { FD_HARRIS = 0, FD_SHI_TOMAS = 1, FD_FAST = 2 } fd_detector_kind;

/* FeaturePointDetector::Options (feature_point_detector.h:15-20) plus the three detectors' private
 * SubOptions (harris_detector.h:12-15, shi_tomas_detector.h:12-14, fast_detector.h:12-15), which the
 * reference gives no accessor for; defaults below are the reference's values. */
typedef struct fd_detect_params {
    int32_t kind;                 /* fd_detector_kind */
    float min_valid_response;     /* kMinValidResponse, default 0.1f */
    int32_t min_feature_distance; /* kMinFeatureDistance, default 15 */
    uint32_t needed_feature_num;  /* DetectGoodFeatures argument */
    float harris_alpha;           /* SubOptions::kAlpha, default 0.04f (Harris only) */
    int32_t fast_n;               /* SubOptions::kN, default 12: only toggles the pre-check (>= 12) */
    int32_t fast_min_pixel_diff;  /* SubOptions::kMinPixelDiffValue, default 15 */
    int32_t reserved;             /* must be 0 */
} fd_detect_params;

/* BriefDescriptor::Options (descriptor_brief.h:16-19). */
typedef struct fd_brief_params {
    int32_t length;               /* kLength, 1..256, default 256 */
    int32_t half_patch_size;      /* kHalfPatchSize, default 8 (orientation patch) */
    int32_t sampling;             /* fd_brief_sampling */
    int32_t reserved;             /* must be 0 */
} fd_brief_params;

/* Float-coordinate pixel fetch of the absent upstream Image class (guess G1). */
typedef enum fd_brief_sampling {
    FD_SAMPLE_BILINEAR = 0,       /* four-tap bilinear around the truncated coordinate (the parity mode) */
    FD_SAMPLE_TRUNCATE = 1        /* nearest-below pixel; no reference parity available */
} fd_brief_sampling;

/* FeatureLineDetector::Options (feature_line_detector.h:40-45): only the field the map stage reads. */
typedef struct fd_lsd_params {
    float min_valid_gradient_norm; /* kMinValidGradientNorm, default 20.0f */
    int32_t want_sorted;           /* 1: also produce the seed order (norm descending) */
} fd_lsd_params;

/* One selected keypoint: the reference returns Vec2(x = col, y = row) as floats
 * (feature_point_detector.cpp:67); response is the candidate's score. */
typedef struct fd_keypoint {
    float x, y;
    float response;
    int32_t reserved;
} fd_keypoint;

/* One candidate as the reference's std::pair<float, Pixel> (feature_point_detector.h:36,51). */
typedef struct fd_candidate {
    float response;
    int32_t x, y;
} fd_candidate;

/* ---- lifecycle --------------------------------------------------------------------------------- */
fd_status fd_create(int device_ordinal, fd_context **out_ctx);
fd_status fd_destroy(fd_context *ctx);
const char *fd_last_error(const fd_context *ctx);
const char *fd_version(void);
/* Run on an externally owned cudaStream_t (e.g. PyTorch's current stream).  The handle is used as given, so
 * NULL is the legacy default stream; pass fd_own_stream(ctx) to return to the context's own stream. */
fd_status fd_set_stream(fd_context *ctx, void *cuda_stream);
void *fd_own_stream(const fd_context *ctx);
fd_status fd_sync(fd_context *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t fd_launch_count(const fd_context *ctx);

/* ---- frames (replaces the GrayImage argument of every reference entry point) ------------------- */
/* Copy n_frames contiguous host frames into context-owned device memory (pitched for aligned loads). */
fd_status fd_upload_frames(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames);
/* Use frames that already live on the device.  pitch / frame_stride in bytes. */
fd_status fd_bind_device_frames(fd_context *ctx, const uint8_t *dev_frames, int rows, int cols, int64_t pitch, int64_t frame_stride,
                                int n_frames);

/* ---- pre-existing features (the in/out `features` argument, feature_point_detector.cpp:12-16,90-98)
 * counts[f] features for frame f at xy[(f*capacity + i)*2 + {0,1}] (host memory, floats as in Vec2).
 * They mask their (2d+1)^2 neighbourhood out of candidate generation and count toward needed_feature_num.
 * Pass n_frames = 0 to clear. */
fd_status fd_set_existing_features(fd_context *ctx, const float *host_xy, const int32_t *host_counts, int capacity, int n_frames);

/* ---- kernels 1-3: FeaturePointDetector::DetectGoodFeatures (feature_point_detector.cpp:7-25) ---- */
/* Candidate generation (ComputeCandidates: harris.cpp:5-15, shi_tomas.cpp:5-15, fast.cpp:83-98) followed
 * by sort + greedy min-distance selection (SelectGoodFeatures, feature_point_detector.cpp:54-74) for every
 * bound frame.  cand_capacity bounds candidates per frame (0 = rows*cols, always sufficient);
 * FD_ERR_CAPACITY is reported by fd_sync / the download calls if a frame overflowed. */
fd_status fd_detect(fd_context *ctx, const fd_detect_params *params, int cand_capacity);
/* Candidate generation only (no selection); candidates stay unsorted in the context. */
fd_status fd_compute_candidates(fd_context *ctx, const fd_detect_params *params, int cand_capacity);
/* Optional dense outputs written during the NEXT fd_detect / fd_compute_candidates (device pointers,
 * n_frames*rows*cols elements, row-major, no pitch); NULL disables.  response: Harris / Shi-Tomasi
 * responses_ (thresholded, harris.cpp:74-103); score: raw FAST score 0..16 (fast.cpp:11-81). */
fd_status fd_set_dense_outputs(fd_context *ctx, float *dev_response_map, uint8_t *dev_fast_score_map);

/* Results.  Keypoints: only the NEW features, in selection order (append them after the pre-existing
 * ones to get the reference's vector).  Candidates: sorted by response descending, ties in raster order.
 * Host variants synchronise; the device variant returns context-owned pointers valid until the next
 * detect call (kp slots have capacity *kp_capacity). */
fd_status fd_download_keypoints(fd_context *ctx, fd_keypoint *host_kp, int32_t *host_counts, int kp_capacity);
fd_status fd_download_candidates(fd_context *ctx, int frame, fd_candidate *host_cand, int64_t capacity, int64_t *n_out);
fd_status fd_candidate_counts(fd_context *ctx, int32_t *host_counts);
fd_status fd_device_keypoints(fd_context *ctx, const fd_keypoint **dev_kp, const int32_t **dev_counts, int *kp_capacity);

/* The reference's call pattern -- one frame (or a few) per call, DetectGoodFeatures (feature_point_detector.cpp:7-25) followed by
 * Descriptor::Compute (descriptor.h:28-40) on the same image -- as ONE host-to-host call: upload, candidates, selection, BRIEF and the
 * downloads of counts, keypoints and descriptors with a single synchronisation (the separate calls synchronise three times, which is a
 * fifth of a one-frame call).  brief and host_desc are both null to skip the descriptors.  host_kp / host_desc are n_frames slots of
 * kp_capacity entries; pre-existing features (fd_set_existing_features) apply as in fd_detect. */
fd_status fd_detect_describe_host(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames, const fd_detect_params *det,
                                  const fd_brief_params *brief, int cand_capacity, fd_keypoint *host_kp, int32_t *host_counts, uint8_t *host_desc,
                                  int kp_capacity);

/* ---- one large frame, row-tiled across GPUs -----------------------------------------
 * fd_set_tile declares the bound frames to be rows [row_offset, row_offset + rows) of an image full_rows tall, of which
 * the local rows [own_first_row, own_first_row + own_row_count) are this tile's own and the rest is halo (3 rows each side
 * cover Harris / Shi-Tomasi / FAST).  fd_compute_candidates then emits candidates for the own rows only, with ABSOLUTE row
 * numbers, and FAST's running offset (fast.cpp:85-93) is indexed by the absolute pixel position, so tiles are seam-free.
 * full_rows <= 0 clears the tile.  The candidate keys of all tiles are gathered by the caller (NCCL / peer copies; device
 * pointers from fd_device_candidates) and handed to fd_select_candidates on one GPU: the greedy selection of
 * feature_point_detector.cpp:54-74 is global per frame.  A key is 64 bits: high word = ~ordered(response) (ascending key
 * = descending response), low word = (row << 16) | col. */
fd_status fd_set_tile(fd_context *ctx, int row_offset, int own_first_row, int own_row_count, int full_rows);
fd_status fd_device_candidates(fd_context *ctx, const uint64_t **dev_keys, const uint32_t **dev_counts, uint32_t *capacity);
/* Pack the candidate keys of all bound frames back to back into caller-owned device memory (device-to-device);
 * host_counts[f] receives each frame's count.  Synchronises. */
fd_status fd_export_candidates(fd_context *ctx, uint64_t *dev_dst, int64_t dst_capacity, int64_t *host_counts);
/* Selection over caller-supplied candidate keys: n_frames slots of `capacity` keys (dev_counts[f] valid) for frames of
 * rows x cols pixels.  dev_keys is scratch-free but must be writable device memory.  Results: fd_download_keypoints. */
fd_status fd_select_candidates(fd_context *ctx, const fd_detect_params *params, uint64_t *dev_keys, const uint32_t *dev_counts, uint32_t capacity,
                               int rows, int cols, int n_frames);

/* ---- the same, as one call sequence over the GPUs of a box (fd_tiled.cu) ------------------------------
 * Replaces, for frames too large (or too urgent) for one GPU, the dense stages of feature_point_harris_detector.cpp:17-137 /
 * feature_point_shi_tomas_detector.cpp:17-137 / feature_point_fast_detector.cpp:83-98 and the global selection of
 * feature_point_detector.cpp:54-74.  One process; tile k of every frame lives on device_ordinals[k] (ordinals may repeat: all tiles
 * on one GPU is a valid, slower, configuration).  Rows are split into n_tiles contiguous blocks; each tile keeps its own rows plus a
 * 3-row halo.  Own rows come from the host (fd_tiled_upload_frames: each device receives only its own rows) or from frames resident
 * on one device (fd_tiled_scatter_device_frames: peer copies); the halo rows then travel tile to tile as device-to-device peer copies
 * (NVLink when peer access is available), 2 x 3 x cols bytes per interior seam and frame, ordered by events.  fd_tiled_detect runs
 * fd_compute_candidates per tile and selects on the first tile's device; only the best-ranked few thousand keys of a frame travel
 * there (tile rank histograms summed on that device, the first rank limit handed back, the keys below it compacted per tile and read
 * through peer pointers -- no count ever visits the host), and a frame that needs more is finished from a full gather of its keys, so
 * the result is exactly the untiled one.  fd_tiled_compute_candidates packs EVERY key of every tile on the first tile's device (the
 * candidate list as a result).  Nothing synchronises with the host until a download / fd_tiled_sync.  Pre-existing features are not supported on tiles.  Results equal the untiled run
 * (tests/test_tiled_abi.py). */
typedef struct fd_tiled fd_tiled;
fd_status fd_tiled_create(const int *device_ordinals, int n_tiles, fd_tiled **out);   /* n_tiles <= 16 */
fd_status fd_tiled_destroy(fd_tiled *t);
const char *fd_tiled_last_error(const fd_tiled *t);
fd_status fd_tiled_upload_frames(fd_tiled *t, const uint8_t *host_frames, int rows, int cols, int n_frames);   /* synchronises (pageable source) */
fd_status fd_tiled_scatter_device_frames(fd_tiled *t, const uint8_t *dev_frames, int rows, int cols, int64_t pitch, int64_t frame_stride, int n_frames);
/* Where tile `tile` keeps its OWN rows (absolute rows [own_first_row, own_first_row + own_row_count) of every frame; row pitch and
 * frame stride in bytes) and the stream its work is ordered on: a producer may rewrite the own rows in place on that stream and
 * call fd_tiled_exchange_halos, which moves only the halo rows. */
fd_status fd_tiled_tile_info(fd_tiled *t, int tile, int *device, int *own_first_row, int *own_row_count, uint8_t **dev_own_rows, int64_t *pitch,
                             int64_t *frame_stride, void **cuda_stream);
fd_status fd_tiled_exchange_halos(fd_tiled *t);
uint64_t fd_tiled_halo_bytes(const fd_tiled *t);   /* bytes the last halo exchange moved between tiles */
fd_status fd_tiled_compute_candidates(fd_tiled *t, const fd_detect_params *params, int cand_capacity_per_tile);   /* candidates + gather */
fd_status fd_tiled_detect(fd_tiled *t, const fd_detect_params *params, int cand_capacity_per_tile);               /* ... + selection */
fd_status fd_tiled_sync(fd_tiled *t);   /* FD_ERR_CAPACITY if a tile's candidate slot overflowed */
fd_status fd_tiled_candidate_counts(fd_tiled *t, int32_t *host_counts);   /* per frame, all tiles together */
fd_status fd_tiled_device_candidates(fd_tiled *t, const uint64_t **dev_keys, const uint32_t **dev_counts, uint32_t *capacity, int *device);
fd_status fd_tiled_download_candidates(fd_tiled *t, int frame, fd_candidate *host_cand, int64_t capacity, int64_t *n_out);
fd_status fd_tiled_download_keypoints(fd_tiled *t, fd_keypoint *host_kp, int32_t *host_counts, int kp_capacity);
fd_context *fd_tiled_root_context(fd_tiled *t);   /* the first tile's context (holds the selected keypoints) */

/* FeaturePointDetector::SparsifyFeatures (feature_point_detector.cpp:27-52): first-come grid filter over an
 * existing feature list.  Pure host-side integer logic over at most a few thousand points; kept in the
 * library so the drop-in class has one implementation.  status is in/out (n entries). */
fd_status fd_sparsify(const float *host_xy, int n, int image_rows, int image_cols, int grid_rows, int grid_cols,
                      uint8_t status_need_filter, uint8_t status_after_filter, uint8_t *status);

/* ---- kernel 4: Descriptor<BriefType>::Compute (descriptor.h:28-40, descriptor_brief.cpp:8-50) --- */
/* Describe the keypoints selected by the last fd_detect (device-resident, no host round trip). */
fd_status fd_describe_selected(fd_context *ctx, const fd_brief_params *params);
/* Describe caller-supplied keypoints: counts[f] points for frame f at xy[(f*capacity+i)*2] (host memory). */
fd_status fd_describe_points(fd_context *ctx, const fd_brief_params *params, const float *host_xy, const int32_t *host_counts,
                             int capacity, int n_frames);
/* Descriptors: 32 bytes per keypoint, bit i at byte i/8, bit position i%8 (LSB first); bits >= length are 0.
 * Slot layout follows the keypoints that were described. */
fd_status fd_download_descriptors(fd_context *ctx, uint8_t *host_desc, int kp_capacity);
fd_status fd_device_descriptors(fd_context *ctx, const uint8_t **dev_desc, int *kp_capacity);
/* The std::vector<Vec> overload of Descriptor::Compute (descriptor.h:43-62) for the descriptors just computed: bit set -> +1.0f,
 * clear -> -1.0f, kLength floats per keypoint slot, into dev_out (n_frames * kp_capacity * kLength floats, device memory) or,
 * if NULL, into a context-owned buffer that fd_download_descriptors_float copies out (slots past a frame's count: unspecified). */
fd_status fd_descriptors_as_float(fd_context *ctx, float *dev_out);
fd_status fd_download_descriptors_float(fd_context *ctx, float *host_desc, int kp_capacity);

/* ---- page-locked host memory (no reference counterpart) ------------------------------------------ */
/* Buffers that cross the bus on every call -- the LSD maps a drop-in class downloads (16 bytes per pixel), frames a caller
 * uploads -- move at the link's speed only from page-locked memory; from pageable memory the driver stages them at a fifth of it.
 * fd_host_alloc returns cudaMallocHost memory (FD_ERR_OUT_OF_MEMORY when the host refuses), fd_host_free releases it. */
fd_status fd_host_alloc(void **ptr, size_t bytes);
fd_status fd_host_free(void *ptr);

/* ---- kernel 5: FeatureLineDetector::ComputeLineLevelAngleMap (feature_line_detector.cpp:56-97) -- */
/* For every bound frame: gradient norm and level-line angle maps, written as rows x cols floats, row-major
 * (the reference's pixels_ is (rows-1) x (cols-1); the extra last row / column is zero, which keeps every
 * map row 16-byte aligned for vector stores).  angle is 0 where the pixel is not valid.  If want_sorted:
 * the valid pixels ordered by norm descending, ties in the reference's column-major push order, as
 * (row * cols + col) indices.  Outputs are device pointers, 16-byte aligned: norm / angle
 * n_frames*rows*cols floats; sorted_idx n_frames*rows*cols int32 (one slot per frame); n_valid n_frames
 * int32.  Any may be context-owned: pass NULL and fetch with fd_lsd_device_outputs. */
fd_status fd_lsd_field(fd_context *ctx, const fd_lsd_params *params, float *dev_norm, float *dev_angle, int32_t *dev_sorted_idx,
                       int32_t *dev_n_valid);
fd_status fd_lsd_device_outputs(fd_context *ctx, const float **dev_norm, const float **dev_angle, const int32_t **dev_sorted_idx,
                                const int32_t **dev_n_valid);
fd_status fd_lsd_download(fd_context *ctx, int frame, float *host_norm, float *host_angle, int32_t *host_sorted_idx, int64_t sorted_capacity,
                          int32_t *host_n_valid);

/* ---- post-processing of a learned detector's outputs ------------------------------------------------
 * The part of NNFeaturePointDetector that is not the ONNX session (nn_feature_point_detector.cpp): the model runs wherever
 * the caller runs it and leaves its outputs in device memory.
 * fd_nn_select_from_heatmap = CreateMask (:59-72) + SelectKeypointCandidatesFromHeatMap (:128-139) +
 * SelectGoodFeaturesFromCandidates (:141-155) on n_frames row-major rows x cols float heat maps: pixels with response >
 * min_response, invalid_boundary pixels inside the map and outside the squares of the pre-existing features
 * (fd_set_existing_features), are walked by response descending -- equal responses: the later raster position first, as
 * the reference's multimap does -- and kept at min_feature_distance until max_features (pre-existing ones included) is
 * reached.  Results: fd_download_keypoints / fd_device_keypoints (new features only).
 * fd_nn_sample_descriptors = ExtractDescriptorsForSelectedFeatures (:163-193) for the keypoints just selected: dev_maps
 * holds n_frames x channels planes of map_rows x map_cols floats (the model's 1/8-resolution descriptor volume); the
 * output is `channels` floats per keypoint slot, in dev_out (n_frames * kp_capacity * channels floats) or, if NULL, in a
 * context-owned buffer that fd_nn_download_descriptors copies out. */
/* ---- Hamming matching of the packed descriptors ------------------------------------------------------
 * The step after Descriptor<BriefType>::Compute for any caller.  The reference has NO matcher (descriptor.h:43-62, the +1 / -1
 * float form, is its only consumer-facing transform), so this entry point has no reference parity; its checker is a numpy
 * XOR / popcount restatement.  For every query descriptor: the nearest train descriptor by Hamming distance over the 256 bits (lowest
 * index among equal distances), that distance and the second-smallest distance; index -1 where there is no query or no train set. */
typedef struct fd_match {
    int32_t train_index;      /* -1: no match */
    int32_t distance;         /* 0 .. 256, -1 without a match */
    int32_t second_distance;  /* -1 when the train set has fewer than two descriptors */
    int32_t reserved;
} fd_match;
/* Frame f of the last described set (fd_describe_selected / fd_describe_points) against frame f + 1: n_frames - 1 pairs, results in a
 * context-owned device buffer of (n_frames - 1) x kp_capacity fd_match (fd_download_matches copies it out; synchronises). */
fd_status fd_match_consecutive(fd_context *ctx);
/* Caller-supplied device sets: n_pairs query sets of capacity_a x 32 bytes against n_pairs train sets of capacity_b x 32 bytes;
 * dev_out receives n_pairs x capacity_a fd_match. */
fd_status fd_match_descriptors(fd_context *ctx, const uint8_t *dev_desc_a, const int32_t *dev_counts_a, int capacity_a, const uint8_t *dev_desc_b,
                               const int32_t *dev_counts_b, int capacity_b, int n_pairs, fd_match *dev_out);
fd_status fd_download_matches(fd_context *ctx, fd_match *host_matches, int kp_capacity);

typedef struct fd_nn_params {
    float min_response;           /* Options::kMinResponse, default 0.1f (nn_feature_point_detector.h:28) */
    int32_t invalid_boundary;     /* kInvalidBoundary, default 3 */
    int32_t min_feature_distance; /* kMinFeatureDistance, default 15 */
    uint32_t max_features;        /* kMaxNumberOfDetectedFeatures, default 240 */
    int32_t reserved;             /* must be 0 */
} fd_nn_params;
/* Model outputs that live in HOST memory (a CPU execution provider): copy `count` floats into one of two context-owned device
 * buffers (slot 0 / 1, e.g. heat map / descriptor volume) and get the device pointer to hand to the calls below.  The copy is
 * asynchronous on the context's stream when the source is pinned. */
fd_status fd_upload_floats(fd_context *ctx, int slot, const float *host, size_t count, const float **dev);
fd_status fd_nn_select_from_heatmap(fd_context *ctx, const float *dev_heatmap, int rows, int cols, int n_frames, const fd_nn_params *params,
                                    int cand_capacity);
fd_status fd_nn_sample_descriptors(fd_context *ctx, const float *dev_maps, int channels, int map_rows, int map_cols, float *dev_out);
fd_status fd_nn_download_descriptors(fd_context *ctx, float *host_desc, int kp_capacity);
/* The same sampling at caller-supplied points (the reference describes the pre-existing features too, :166): counts[f]
 * points for frame f at xy[(f*capacity + i)*2] (host memory); host_out receives n_frames*capacity*channels floats
 * (unused slots zero).  Synchronises. */
fd_status fd_nn_sample_descriptors_at(fd_context *ctx, const float *dev_maps, int channels, int map_rows, int map_cols, const float *host_xy,
                                      const int32_t *host_counts, int capacity, int n_frames, float *host_out);

/* ---- diagnostics of the host-built tables (no GPU needed; used by the CPU test-suite) ------------- */
/* Bit patterns of the FAST running offset (fast.cpp:85,93) for masked-in pixel index k = 0..count-1, as
 * evaluated from the piecewise-linear table the kernel uses; *n_segments = pieces in that table. */
fd_status fd_debug_fast_offset_bits(uint32_t count, uint32_t *out_bits, int32_t *n_segments);
/* The 65536-entry longest-circular-run table the FAST kernel looks scores up in. */
fd_status fd_debug_run_length_lut(uint8_t *out_65536);
/* The smallest trace (sum of the two windowed squared-gradient sums) that passes the Harris pre-test of harris.cpp:98 at this threshold --
 * the kernel replaces the test's three multiplications by one compare against it; NaN when no trace passes. */
fd_status fd_debug_harris_trace_min(float min_valid_response, float *out_trace_min);

/* Memory-safety check that needs no external tool: a context created while the environment holds FD_B200_GUARD=1 allocates every
 * context-owned device buffer at exactly the size a call asks for, between two 256-byte red zones.  This call synchronises and
 * verifies all of them (*n_checked = buffers looked at); a kernel that wrote outside its buffer turns up as FD_ERR_CUDA with the
 * buffer's name in fd_last_error.  FD_ERR_NOT_READY if the context was created without the variable.  tools/sanitize_paths.py
 * drives every kernel instantiation under it. */
fd_status fd_debug_check_guards(fd_context *ctx, int32_t *n_checked);

#ifdef __cplusplus
}
#endif

#endif /* FD_B200_H_ */
