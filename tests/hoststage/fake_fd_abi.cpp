// TEST INFRASTRUCTURE (CPU only, never shipped, never linked into libfd_b200.so or libfeature_detector_b200.so).
//
// A stand-in for the C ABI of include/fd_b200.h whose "kernels" are the CPU oracle (oracle/libfd_oracle.so).  It exists so
// that `-m "not gpu"` can run the HOST glue of the drop-in C++ classes (feature_detector_b200/cpp/*.cpp: marshalling, the
// in/out `features` contract, the lazily rebuilt mask() / candidates(), bit unpacking, PixelParam filling, the LSD host stage)
// end to end through fd_dropin_check and hold its output against the same golden answers as the GPU run.  It proves nothing
// about the CUDA kernels -- the gpu-marked tests do that against the real library -- and the product has no such path:
// libfd_b200.so fails with FD_ERR_NO_DEVICE without a GPU (tests/test_abi.py::test_no_cpu_fallback_without_device).
// Only the entry points the drop-in classes call are provided; one frame per call is enough for them.
#include <cstdlib>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fd_b200.h"
#include "../../oracle/fd_oracle.h"

struct fd_context {
    std::string err;
    std::vector<uint8_t> frame;
    int rows = 0, cols = 0;
    std::vector<float> pre_xy;                 // pre-existing features of frame 0
    std::vector<fd_keypoint> kp;               // new keypoints of the last selection
    std::vector<fd_candidate> cand;            // sorted: response descending, raster ties
    std::vector<uint8_t> desc;                 // 32 bytes per described point
    std::vector<float> norm, angle;            // rows x cols, zero last row / column
    std::vector<int32_t> seeds;
    std::vector<float> slot[2];
};

extern "C" {

fd_status fd_create(int, fd_context **out_ctx) {
    if (!out_ctx) return FD_ERR_INVALID_ARGUMENT;
    *out_ctx = new fd_context();
    return FD_OK;
}
fd_status fd_destroy(fd_context *ctx) {
    delete ctx;
    return FD_OK;
}
const char *fd_last_error(const fd_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

fd_status fd_upload_frames(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames) {
    if (!ctx || !host_frames || rows <= 0 || cols <= 0 || n_frames != 1) return FD_ERR_INVALID_ARGUMENT;
    ctx->frame.assign(host_frames, host_frames + size_t(rows) * cols);
    ctx->rows = rows;
    ctx->cols = cols;
    return FD_OK;
}

fd_status fd_set_existing_features(fd_context *ctx, const float *host_xy, const int32_t *host_counts, int, int n_frames) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    ctx->pre_xy.clear();
    if (n_frames > 0 && host_xy && host_counts) ctx->pre_xy.assign(host_xy, host_xy + size_t(host_counts[0]) * 2);
    return FD_OK;
}

fd_status fd_detect(fd_context *ctx, const fd_detect_params *p, int) {
    if (!ctx || !p || ctx->frame.empty()) return FD_ERR_INVALID_ARGUMENT;
    const int n_pre = int(ctx->pre_xy.size() / 2);
    const int max_feats = int(p->needed_feature_num) + n_pre + 8;
    std::vector<float> feats(size_t(max_feats) * 2, 0.0f);
    std::copy(ctx->pre_xy.begin(), ctx->pre_xy.end(), feats.begin());
    const int64_t max_cand = int64_t(ctx->rows) * ctx->cols;
    std::vector<float> resp(static_cast<size_t>(max_cand));
    std::vector<int32_t> xy(static_cast<size_t>(max_cand) * 2);
    int n_out = 0;
    int64_t n_cand = 0;
    const int rc = orc_detect(p->kind, ctx->frame.data(), ctx->rows, ctx->cols, p->min_valid_response, p->min_feature_distance, p->needed_feature_num,
                              p->fast_n, feats.data(), n_pre, max_feats, &n_out, resp.data(), xy.data(), max_cand, &n_cand, nullptr, nullptr);
    if (rc < 0) return FD_ERR_CAPACITY;
    ctx->cand.resize(size_t(n_cand));
    for (int64_t i = 0; i < n_cand; ++i) ctx->cand[size_t(i)] = {resp[size_t(i)], xy[2 * size_t(i)], xy[2 * size_t(i) + 1]};
    std::stable_sort(ctx->cand.begin(), ctx->cand.end(), [](const fd_candidate &a, const fd_candidate &b) {
        if (a.response != b.response) return a.response > b.response;
        if (a.y != b.y) return a.y < b.y;
        return a.x < b.x;
    });
    ctx->kp.clear();
    for (int i = n_pre; i < n_out; ++i) {
        const float x = feats[2 * size_t(i)], y = feats[2 * size_t(i) + 1];
        float r = 0.0f;
        for (const fd_candidate &c : ctx->cand)
            if (c.x == int32_t(x) && c.y == int32_t(y)) {
                r = c.response;
                break;
            }
        ctx->kp.push_back({x, y, r, 0});
    }
    return FD_OK;
}

fd_status fd_download_keypoints(fd_context *ctx, fd_keypoint *host_kp, int32_t *host_counts, int kp_capacity) {
    if (!ctx || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    host_counts[0] = int32_t(ctx->kp.size());
    if (host_kp) {
        if (int(ctx->kp.size()) > kp_capacity) return FD_ERR_CAPACITY;
        std::copy(ctx->kp.begin(), ctx->kp.end(), host_kp);
    }
    return FD_OK;
}

// the one-call form the drop-in detector uses: the same three steps (descriptors are not asked for by the classes under test)
fd_status fd_detect_describe_host(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames, const fd_detect_params *det,
                                  const fd_brief_params *brief, int cand_capacity, fd_keypoint *host_kp, int32_t *host_counts, uint8_t *host_desc,
                                  int kp_capacity) {
    if (brief != nullptr || host_desc != nullptr) return FD_ERR_INVALID_ARGUMENT;
    fd_status st = fd_upload_frames(ctx, host_frames, rows, cols, n_frames);
    if (st == FD_OK) st = fd_detect(ctx, det, cand_capacity);
    if (st == FD_OK) st = fd_download_keypoints(ctx, host_kp, host_counts, kp_capacity);
    return st;
}

fd_status fd_download_candidates(fd_context *ctx, int frame, fd_candidate *host_cand, int64_t capacity, int64_t *n_out) {
    if (!ctx || frame != 0 || !n_out) return FD_ERR_INVALID_ARGUMENT;
    *n_out = int64_t(ctx->cand.size());
    if (host_cand) {
        if (capacity < *n_out) return FD_ERR_CAPACITY;
        std::copy(ctx->cand.begin(), ctx->cand.end(), host_cand);
    }
    return FD_OK;
}

fd_status fd_sparsify(const float *host_xy, int n, int image_rows, int image_cols, int grid_rows, int grid_cols, uint8_t status_need_filter,
                      uint8_t status_after_filter, uint8_t *status) {
    orc_sparsify(host_xy, n, image_rows, image_cols, grid_rows, grid_cols, status_need_filter, status_after_filter, status, n);
    return FD_OK;
}

fd_status fd_describe_points(fd_context *ctx, const fd_brief_params *p, const float *host_xy, const int32_t *host_counts, int capacity, int n_frames) {
    if (!ctx || !p || !host_xy || !host_counts || n_frames != 1 || ctx->frame.empty() || p->sampling != FD_SAMPLE_BILINEAR) return FD_ERR_INVALID_ARGUMENT;
    const int n = host_counts[0];
    std::vector<uint8_t> bits(size_t(n) * p->length);
    orc_brief(ctx->frame.data(), ctx->rows, ctx->cols, host_xy, n, p->length, p->half_patch_size, bits.data());
    ctx->desc.assign(size_t(capacity) * 32, 0);
    for (int i = 0; i < n; ++i)
        for (int b = 0; b < p->length; ++b)
            if (bits[size_t(i) * p->length + b]) ctx->desc[size_t(i) * 32 + (b >> 3)] |= uint8_t(1u << (b & 7));
    return FD_OK;
}

fd_status fd_download_descriptors(fd_context *ctx, uint8_t *host_desc, int kp_capacity) {
    if (!ctx || !host_desc) return FD_ERR_INVALID_ARGUMENT;
    std::memcpy(host_desc, ctx->desc.data(), std::min(ctx->desc.size(), size_t(kp_capacity) * 32));
    return FD_OK;
}

fd_status fd_lsd_field(fd_context *ctx, const fd_lsd_params *p, float *, float *, int32_t *, int32_t *) {
    if (!ctx || !p || ctx->frame.empty() || ctx->rows < 2 || ctx->cols < 2) return FD_ERR_INVALID_ARGUMENT;
    const int pr = ctx->rows - 1, pc = ctx->cols - 1;
    std::vector<float> norm(size_t(pr) * pc), angle(size_t(pr) * pc);
    std::vector<uint8_t> valid(size_t(pr) * pc);
    std::vector<int32_t> sorted_rc(size_t(pr) * pc * 2);
    int64_t n_sorted = 0;
    if (orc_lsd_map(ctx->frame.data(), ctx->rows, ctx->cols, p->min_valid_gradient_norm, norm.data(), angle.data(), valid.data(), sorted_rc.data(),
                    int64_t(pr) * pc, &n_sorted) != 1)
        return FD_ERR_INVALID_ARGUMENT;
    ctx->norm.assign(size_t(ctx->rows) * ctx->cols, 0.0f);
    ctx->angle.assign(size_t(ctx->rows) * ctx->cols, 0.0f);
    for (int r = 0; r < pr; ++r)
        for (int c = 0; c < pc; ++c) {
            ctx->norm[size_t(r) * ctx->cols + c] = norm[size_t(r) * pc + c];
            ctx->angle[size_t(r) * ctx->cols + c] = valid[size_t(r) * pc + c] ? angle[size_t(r) * pc + c] : 0.0f;
        }
    ctx->seeds.resize(size_t(n_sorted));
    for (int64_t i = 0; i < n_sorted; ++i) ctx->seeds[size_t(i)] = sorted_rc[2 * i] * ctx->cols + sorted_rc[2 * i + 1];
    return FD_OK;
}

fd_status fd_lsd_download(fd_context *ctx, int frame, float *host_norm, float *host_angle, int32_t *host_sorted_idx, int64_t sorted_capacity,
                          int32_t *host_n_valid) {
    if (!ctx || frame != 0) return FD_ERR_INVALID_ARGUMENT;
    if (host_norm) std::copy(ctx->norm.begin(), ctx->norm.end(), host_norm);
    if (host_angle) std::copy(ctx->angle.begin(), ctx->angle.end(), host_angle);
    if (host_sorted_idx) {
        if (sorted_capacity < int64_t(ctx->seeds.size())) return FD_ERR_CAPACITY;
        std::copy(ctx->seeds.begin(), ctx->seeds.end(), host_sorted_idx);
    }
    if (host_n_valid) *host_n_valid = int32_t(ctx->seeds.size());
    return FD_OK;
}

fd_status fd_host_alloc(void **ptr, size_t bytes) {
    *ptr = std::malloc(bytes ? bytes : 16);
    return *ptr ? FD_OK : FD_ERR_OUT_OF_MEMORY;
}
fd_status fd_host_free(void *ptr) {
    std::free(ptr);
    return FD_OK;
}

fd_status fd_upload_floats(fd_context *ctx, int slot, const float *host, size_t count, const float **dev) {
    if (!ctx || slot < 0 || slot > 1 || !host || !dev) return FD_ERR_INVALID_ARGUMENT;
    ctx->slot[slot].assign(host, host + count);
    *dev = ctx->slot[slot].data();
    return FD_OK;
}

fd_status fd_nn_select_from_heatmap(fd_context *ctx, const float *heatmap, int rows, int cols, int n_frames, const fd_nn_params *p, int) {
    if (!ctx || !heatmap || !p || n_frames != 1) return FD_ERR_INVALID_ARGUMENT;
    const int n_pre = int(ctx->pre_xy.size() / 2);
    const int max_feats = int(p->max_features) + n_pre + 8;
    std::vector<float> feats(size_t(max_feats) * 2, 0.0f);
    std::copy(ctx->pre_xy.begin(), ctx->pre_xy.end(), feats.begin());
    int n_out = 0;
    int64_t n_cand = 0;
    orc_nn_select(heatmap, rows, cols, p->min_response, p->invalid_boundary, p->min_feature_distance, int(p->max_features), feats.data(), n_pre, max_feats,
                  &n_out, &n_cand);
    ctx->kp.clear();
    for (int i = n_pre; i < n_out; ++i) {
        const float x = feats[2 * size_t(i)], y = feats[2 * size_t(i) + 1];
        ctx->kp.push_back({x, y, heatmap[size_t(y) * cols + size_t(x)], 0});
    }
    return FD_OK;
}

fd_status fd_nn_sample_descriptors_at(fd_context *ctx, const float *maps, int channels, int map_rows, int map_cols, const float *host_xy,
                                      const int32_t *host_counts, int, int n_frames, float *host_out) {
    if (!ctx || !maps || !host_xy || !host_counts || !host_out || n_frames != 1) return FD_ERR_INVALID_ARGUMENT;
    orc_nn_descriptors(host_xy, host_counts[0], maps, channels, map_rows, map_cols, host_out);
    return FD_OK;
}

}  // extern "C"
