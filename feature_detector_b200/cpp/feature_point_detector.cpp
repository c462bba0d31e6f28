// Host glue of the drop-in point detectors: argument marshalling around the C ABI.  No pixel is touched here.
#include "feature_point_detector.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "fd_b200.h"

namespace feature_detector {

FeaturePointDetector::~FeaturePointDetector() {
    if (ctx_ != nullptr) fd_destroy(ctx_);
}

bool FeaturePointDetector::EnsureContext() const {
    if (ctx_ != nullptr) return true;
    int device = device_;
    if (const char *env = std::getenv("FD_B200_DEVICE")) device = std::atoi(env);
    const fd_status st = fd_create(device, &ctx_);
    if (st != FD_OK) {
        ctx_ = nullptr;
        last_error_ = (st == FD_ERR_NO_DEVICE) ? "no CUDA device (this library has no CPU path)" : "fd_create failed";
        return false;
    }
    return true;
}

bool FeaturePointDetector::Fail(const char *what) const {
    last_error_ = std::string(what) + ": " + (ctx_ ? fd_last_error(ctx_) : "no context");
    return false;
}

bool FeaturePointDetector::DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec2> &features) {
    if (image.data() == nullptr) return false;                    // feature_point_detector.cpp:9
    std::vector<std::vector<Vec2>> batch(1);
    batch[0].swap(features);
    const bool ok = DetectGoodFeaturesBatch(image.data(), image.rows(), image.cols(), 1, needed_feature_num, batch);
    features.swap(batch[0]);
    return ok;
}

bool FeaturePointDetector::DetectGoodFeaturesBatch(const uint8_t *frames, int32_t rows, int32_t cols, int32_t n_frames,
                                                   const uint32_t needed_feature_num, std::vector<std::vector<Vec2>> &features) {
    if (frames == nullptr || rows <= 0 || cols <= 0 || n_frames <= 0) return false;
    if (DetectorKind() < 0) return false;  // the abstract base has no ComputeCandidates
    if (!EnsureContext()) return false;
    features.resize(size_t(n_frames));


    // existing features -> device mask (feature_point_detector.cpp:12-16)
    size_t max_pre = 0;
    for (const auto &f : features) max_pre = std::max(max_pre, f.size());
    std::vector<int32_t> pre_counts(size_t(n_frames), 0);
    if (max_pre > 0) {
        std::vector<float> xy(size_t(n_frames) * max_pre * 2, 0.0f);
        for (int32_t f = 0; f < n_frames; ++f) {
            pre_counts[f] = int32_t(features[f].size());
            for (size_t i = 0; i < features[f].size(); ++i) {
                xy[(size_t(f) * max_pre + i) * 2 + 0] = features[f][i].x();
                xy[(size_t(f) * max_pre + i) * 2 + 1] = features[f][i].y();
            }
        }
        if (fd_set_existing_features(ctx_, xy.data(), pre_counts.data(), int(max_pre), n_frames) != FD_OK) return Fail("fd_set_existing_features");
    } else {
        fd_set_existing_features(ctx_, nullptr, nullptr, 0, 0);
    }

    fd_detect_params p = {};
    p.kind = DetectorKind();
    p.min_valid_response = options_.kMinValidResponse;
    p.min_feature_distance = options_.kMinFeatureDistance;
    p.needed_feature_num = needed_feature_num;
    p.harris_alpha = HarrisAlpha();
    p.fast_n = FastN();
    p.fast_min_pixel_diff = FastMinPixelDiff();
    // At most max(needed - existing, 1) new features per frame (push first, test afterwards: :67-68), and never more than the
    // frame has pixels or the library's keypoint slots hold (needed = UINT32_MAX is a legitimate "all of them").
    const int cap = int(std::min<uint64_t>({uint64_t(std::max<uint32_t>(needed_feature_num, 1u)), uint64_t(rows) * uint64_t(cols), uint64_t(1) << 20}));
    std::vector<fd_keypoint> kp(size_t(n_frames) * cap);
    std::vector<int32_t> counts(size_t(n_frames), 0);
    // Candidate slots: one per pixel can never overflow, but costs 24 B per pixel per frame of scratch on the device.  A batch
    // therefore starts with a quarter of that (a corner detector fills ~7 % on textured frames) and runs again with full slots if a
    // frame overflows (FAST at the reference's default threshold makes nearly every pixel a candidate).
    const int64_t px = int64_t(rows) * cols;
    int cand_capacity = (n_frames > 1 && px / 4 >= (1 << 16)) ? int(px / 4) : 0;
    for (;;) {
        // upload, candidates, selection and the download of counts and keypoints as one call with one synchronisation: the reference's
        // pattern is one frame per call, where every extra round trip to the device shows
        const fd_status st = fd_detect_describe_host(ctx_, frames, rows, cols, n_frames, &p, nullptr, cand_capacity, kp.data(), counts.data(), nullptr, cap);
        if (st == FD_OK) break;
        if (st != FD_ERR_CAPACITY || cand_capacity == 0) return Fail("fd_detect_describe_host");
        cand_capacity = 0;
    }

    // bookkeeping for the lazily rebuilt mask() of frame 0 (the single-frame call's frame)
    mask_rows_ = rows;
    mask_cols_ = cols;
    mask_distance_ = options_.kMinFeatureDistance;
    mask_squares_.clear();
    for (const Vec2 &f : features[0]) mask_squares_.emplace_back(int32_t(f.y()), int32_t(f.x()));  // :94-95 truncation

    for (int32_t f = 0; f < n_frames; ++f) {
        std::vector<Vec2> &out = features[f];
        for (int32_t i = 0; i < counts[f]; ++i) {
            const fd_keypoint &k = kp[size_t(f) * cap + i];
            out.emplace_back(Vec2(k.x, k.y));
            if (f == 0) {
                // the reference returns before clearing the square of the feature that reaches the count (:68-69)
                const bool last_and_full = (out.size() >= needed_feature_num);
                if (!last_and_full) mask_squares_.emplace_back(int32_t(k.y), int32_t(k.x));
            }
        }
    }
    candidates_stale_ = true;
    mask_stale_ = true;
    return true;
}

void FeaturePointDetector::MaterialiseCandidates() const {
    if (!candidates_stale_ || ctx_ == nullptr) return;
    candidates_stale_ = false;
    candidates_.clear();
    int64_t n = 0;
    if (fd_download_candidates(ctx_, 0, nullptr, 0, &n) != FD_OK || n == 0) return;
    std::vector<fd_candidate> c(static_cast<size_t>(n));
    if (fd_download_candidates(ctx_, 0, c.data(), n, &n) != FD_OK) return;
    candidates_.reserve(c.size());
    for (const fd_candidate &e : c) candidates_.emplace_back(e.response, Pixel(e.x, e.y));  // sorted like :58-60
}

void FeaturePointDetector::MaterialiseMask() const {
    if (!mask_stale_) return;
    mask_stale_ = false;
    mask_.setConstant(mask_rows_, mask_cols_, 1);                                            // :13 / :91
    const int32_t d = mask_distance_;
    for (const auto &rc : mask_squares_) {                                                    // :76-88
        const int32_t r0 = std::max(rc.first - d, 0), r1 = std::min(rc.first + d, mask_rows_ - 1);
        const int32_t c0 = std::max(rc.second - d, 0), c1 = std::min(rc.second + d, mask_cols_ - 1);
        for (int32_t r = r0; r <= r1; ++r)
            for (int32_t c = c0; c <= c1; ++c) mask_(r, c) = 0;
    }
}

void FeaturePointDetector::SparsifyFeatures(const std::vector<Vec2> &features, const int32_t image_rows, const int32_t image_cols,
                                            const uint8_t status_need_filter, const uint8_t status_after_filter, std::vector<uint8_t> &status) {
    if (features.size() != status.size()) status.assign(features.size(), 1);                 // :29-31
    const int32_t grid_rows = options_.kGridFilterRowDivideNumber, grid_cols = options_.kGridFilterColDivideNumber;
    if (grid_rows < 2 || grid_cols < 2) return;  // the reference divides by (n - 1) here (:34-35)
    std::vector<float> xy(features.size() * 2);
    for (size_t i = 0; i < features.size(); ++i) {
        xy[2 * i] = features[i].x();
        xy[2 * i + 1] = features[i].y();
    }
    const std::vector<uint8_t> before(status);
    fd_sparsify(xy.data(), int(features.size()), image_rows, image_cols, grid_rows, grid_cols, status_need_filter, status_after_filter, status.data());

    // Side effect kept for callers that look: the reference leaves its occupancy grid in mask_ (:36,46-47).  A cell is
    // taken iff a feature whose status asked for filtering fell into it.
    mask_stale_ = false;
    mask_.setConstant(grid_rows, grid_cols, 1);
    const float row_step = float(image_rows / (grid_rows - 1)), col_step = float(image_cols / (grid_cols - 1));
    for (size_t i = 0; i < features.size(); ++i) {
        if (before[i] != status_need_filter) continue;
        const float qr = features[i].y() / row_step, qc = features[i].x() / col_step;   // a step of 0 (image smaller than the grid): outside
        if (!(std::fabs(qr) < 2.0e9f && std::fabs(qc) < 2.0e9f)) continue;
        const int32_t row = int32_t(qr), col = int32_t(qc);
        if (row >= 0 && row < grid_rows && col >= 0 && col < grid_cols) mask_(row, col) = 0;
    }
}

}  // namespace feature_detector
