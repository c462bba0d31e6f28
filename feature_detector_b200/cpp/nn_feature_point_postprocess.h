// The post-processing half of the reference's NNFeaturePointDetector (src/nn_feature_point_detector/nn_feature_point_detector.h:
// Options :22-31; CreateMask, SelectKeypointCandidatesFromHeatMap, SelectGoodFeaturesFromCandidates,
// ExtractDescriptorsForSelectedFeatures :48-61) on the GPU.  The ONNX session is not part of this library: the caller runs the
// model and hands over its outputs -- a heat map of the image's size and a channel-major descriptor volume at 1/8 resolution --
// from host memory (uploaded here) or, with on_device = true, from device memory (a CUDA execution provider).  In a reference
// checkout, DetectGoodFeaturesWithDescriptorBySuperpoint / ...ByDisk call the two methods below instead of the four private ones
// (INTEGRATION.md B.4).
#ifndef FD_B200_NN_FEATURE_POINT_POSTPROCESS_H_
#define FD_B200_NN_FEATURE_POINT_POSTPROCESS_H_

#include <cstdint>
#include <string>
#include <vector>

#include "basic_type.h"

struct fd_context;

namespace feature_detector {

class NNFeaturePointPostProcessor {
public:
    struct Options {   // the reference's names and defaults (nn_feature_point_detector.h:22-31)
        int32_t kInvalidBoundary = 3;
        int32_t kMinFeatureDistance = 15;
        int32_t kMaxNumberOfDetectedFeatures = 240;
        float kMinResponse = 0.1f;
    };

    explicit NNFeaturePointPostProcessor(int device = 0): device_(device) {}
    ~NNFeaturePointPostProcessor();
    NNFeaturePointPostProcessor(const NNFeaturePointPostProcessor &) = delete;
    NNFeaturePointPostProcessor &operator=(const NNFeaturePointPostProcessor &) = delete;

    Options &options() { return options_; }
    const Options &options() const { return options_; }
    const std::string &last_error() const { return last_error_; }

    // CreateMask + SelectKeypointCandidatesFromHeatMap + SelectGoodFeaturesFromCandidates (.cpp:59-72, 128-155).  `features` is
    // in / out like the reference's all_pixel_uv: existing entries are kept, avoided and counted toward the maximum.
    bool SelectGoodFeaturesFromHeatMap(const float *heatmap, int32_t rows, int32_t cols, std::vector<Vec2> &features, bool on_device = false);

    // ExtractDescriptorsForSelectedFeatures (.cpp:163-193) for every entry of `features`: `maps` holds `channels` planes of
    // map_rows x map_cols floats.  One `channels`-long row per feature in `descriptors` (row-major).
    bool ExtractDescriptors(const std::vector<Vec2> &features, const float *maps, int32_t channels, int32_t map_rows, int32_t map_cols,
                            std::vector<float> &descriptors, bool on_device = false);

    // Convenience for the reference's descriptor types (Eigen::Matrix<float, 256 / 128, 1>, or the dynamic Vec).
    template <typename NNFeatureDescriptorType>
    bool ExtractDescriptorsForSelectedFeatures(const std::vector<Vec2> &features, const float *maps, int32_t channels, int32_t map_rows, int32_t map_cols,
                                               std::vector<NNFeatureDescriptorType> &descriptors, bool on_device = false) {
        std::vector<float> flat;
        if (!ExtractDescriptors(features, maps, channels, map_rows, map_cols, flat, on_device)) return false;
        descriptors.resize(features.size());
        for (size_t i = 0; i < features.size(); ++i) {
            if (int(descriptors[i].size()) != channels) descriptors[i].resize(channels);
            for (int32_t j = 0; j < channels; ++j) descriptors[i](j) = flat[i * size_t(channels) + j];
        }
        return true;
    }

private:
    bool EnsureContext();

    Options options_;
    int device_ = 0;
    fd_context *ctx_ = nullptr;
    std::string last_error_;
};

}  // namespace feature_detector

#endif  // FD_B200_NN_FEATURE_POINT_POSTPROCESS_H_
