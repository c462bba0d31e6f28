// Host glue of the NN post-processing class: marshalling around fd_nn_select_from_heatmap / fd_nn_sample_descriptors_at.
#include "nn_feature_point_postprocess.h"

#include <algorithm>
#include <cstdlib>

#include "fd_b200.h"

namespace feature_detector {

NNFeaturePointPostProcessor::~NNFeaturePointPostProcessor() {
    if (ctx_ != nullptr) fd_destroy(ctx_);
}

bool NNFeaturePointPostProcessor::EnsureContext() {
    if (ctx_ != nullptr) return true;
    int device = device_;
    if (const char *env = std::getenv("FD_B200_DEVICE")) device = std::atoi(env);
    if (fd_create(device, &ctx_) != FD_OK) {
        ctx_ = nullptr;
        last_error_ = "fd_create failed (no CUDA device? this library has no CPU path)";
        return false;
    }
    return true;
}

bool NNFeaturePointPostProcessor::SelectGoodFeaturesFromHeatMap(const float *heatmap, int32_t rows, int32_t cols, std::vector<Vec2> &features,
                                                                bool on_device) {
    if (heatmap == nullptr || rows <= 0 || cols <= 0) return false;
    if (!EnsureContext()) return false;
    auto failed = [&]() {
        last_error_ = fd_last_error(ctx_);
        return false;
    };
    const float *dev = heatmap;
    if (!on_device && fd_upload_floats(ctx_, 0, heatmap, size_t(rows) * cols, &dev) != FD_OK) return failed();
    const int32_t n_pre = int32_t(features.size());            // UpdateMaskByFeatures (.cpp:85-91) + the count toward the maximum (:151)
    std::vector<float> pre(size_t(n_pre) * 2);
    for (int32_t i = 0; i < n_pre; ++i) {
        pre[2 * i] = features[i].x();
        pre[2 * i + 1] = features[i].y();
    }
    if (fd_set_existing_features(ctx_, n_pre ? pre.data() : nullptr, &n_pre, std::max(n_pre, 1), n_pre ? 1 : 0) != FD_OK) return failed();
    fd_nn_params p = {};
    p.min_response = options_.kMinResponse;
    p.invalid_boundary = options_.kInvalidBoundary;
    p.min_feature_distance = options_.kMinFeatureDistance;
    p.max_features = uint32_t(std::max(options_.kMaxNumberOfDetectedFeatures, 0));
    if (fd_nn_select_from_heatmap(ctx_, dev, rows, cols, 1, &p, 0) != FD_OK) return failed();
    std::vector<fd_keypoint> kp(std::max<uint32_t>(p.max_features, 1u));
    int32_t n = 0;
    if (fd_download_keypoints(ctx_, kp.data(), &n, int(kp.size())) != FD_OK) return failed();
    if (features.empty()) features.reserve(kp.size());           // .cpp:142-144
    for (int32_t i = 0; i < n; ++i) features.emplace_back(Vec2(kp[i].x, kp[i].y));
    return true;
}

bool NNFeaturePointPostProcessor::ExtractDescriptors(const std::vector<Vec2> &features, const float *maps, int32_t channels, int32_t map_rows,
                                                     int32_t map_cols, std::vector<float> &descriptors, bool on_device) {
    descriptors.clear();
    if (maps == nullptr || channels <= 0 || map_rows <= 0 || map_cols <= 0) return false;
    if (features.empty()) return true;                           // .cpp:166: an empty list yields an empty list
    if (!EnsureContext()) return false;
    auto failed = [&]() {
        last_error_ = fd_last_error(ctx_);
        return false;
    };
    const float *dev = maps;
    if (!on_device && fd_upload_floats(ctx_, 1, maps, size_t(channels) * map_rows * map_cols, &dev) != FD_OK) return failed();
    const int32_t n = int32_t(features.size());
    std::vector<float> xy(size_t(n) * 2);
    for (int32_t i = 0; i < n; ++i) {
        xy[2 * i] = features[i].x();
        xy[2 * i + 1] = features[i].y();
    }
    descriptors.assign(size_t(n) * channels, 0.0f);
    if (fd_nn_sample_descriptors_at(ctx_, dev, channels, map_rows, map_cols, xy.data(), &n, n, 1, descriptors.data()) != FD_OK) return failed();
    return true;
}

}  // namespace feature_detector
