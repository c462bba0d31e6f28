// Drop-in point detectors: same namespace, class names, Options and accessors as the reference's
//   src/feature_point_detector/feature_point_detector.h:9-54
//   src/feature_point_detector/feature_point_harris_detector.h, ..._shi_tomas_detector.h, ..._fast_detector.h
// but every dense stage runs on the GPU through the C ABI of include/fd_b200.h (libfd_b200.so).  There is no CPU
// path behind these classes: without a CUDA device DetectGoodFeatures returns false (last_error() says why).
//
// Differences a caller can observe:
//   * the three detectors' SubOptions, private and unreachable in the reference, get accessors (sub_options());
//     their defaults are the reference's values, so untouched objects behave identically;
//   * candidates() / mask() are materialised from the device on first access after a call instead of eagerly
//     (the detection itself never needs them on the host);
//   * ties between equal responses are resolved in raster order (the reference's std::sort leaves them open);
//   * DetectGoodFeaturesBatch() is an addition: many frames per call, which is what the GPU is for.
// "basic_type.h" / "datatype_image.h" are the caller's Slam_Utility headers (compat/slam_utility/ has stand-ins).
#ifndef FD_B200_FEATURE_POINT_DETECTOR_H_
#define FD_B200_FEATURE_POINT_DETECTOR_H_

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "basic_type.h"
#include "datatype_image.h"

struct fd_context;  // include/fd_b200.h

namespace feature_detector {

class FeaturePointDetector {
public:
    struct Options {
        int32_t kMinFeatureDistance = 15;
        int32_t kGridFilterRowDivideNumber = 12;
        int32_t kGridFilterColDivideNumber = 12;
        float kMinValidResponse = 0.1f;
    };

    FeaturePointDetector() = default;
    virtual ~FeaturePointDetector();
    FeaturePointDetector(const FeaturePointDetector &) = delete;
    FeaturePointDetector &operator=(const FeaturePointDetector &) = delete;

    virtual std::string DetectorTypeName() const { return "None"; }

    // `features` is in/out exactly as in the reference (feature_point_detector.cpp:7-25): a non-empty vector holds
    // features to keep away from (and to count toward needed_feature_num); new ones are appended.
    bool DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec2> &features);
    // feature_point_detector.cpp:27-52
    void SparsifyFeatures(const std::vector<Vec2> &features, const int32_t image_rows, const int32_t image_cols, const uint8_t status_need_filter,
                          const uint8_t status_after_filter, std::vector<uint8_t> &status);

    // n_frames contiguous rows x cols frames; features[f] is in/out per frame like the single-frame call.
    bool DetectGoodFeaturesBatch(const uint8_t *frames, int32_t rows, int32_t cols, int32_t n_frames, const uint32_t needed_feature_num,
                                 std::vector<std::vector<Vec2>> &features);

    Options &options() { return options_; }
    const Options &options() const { return options_; }
    std::vector<std::pair<float, Pixel>> &candidates() { MaterialiseCandidates(); return candidates_; }
    const std::vector<std::pair<float, Pixel>> &candidates() const { MaterialiseCandidates(); return candidates_; }
    MatInt &mask() { MaterialiseMask(); return mask_; }
    const MatInt &mask() const { MaterialiseMask(); return mask_; }

    // GPU plumbing (additions)
    void set_device(int ordinal) { device_ = ordinal; }
    const std::string &last_error() const { return last_error_; }
    fd_context *context() { return EnsureContext() ? ctx_ : nullptr; }  // e.g. to hand the same frames to a BriefDescriptor

protected:
    // 0 Harris, 1 Shi-Tomasi, 2 FAST (fd_detector_kind); < 0 = no detector (the abstract base)
    virtual int32_t DetectorKind() const { return -1; }
    virtual float HarrisAlpha() const { return 0.04f; }
    virtual int32_t FastN() const { return 12; }
    virtual int32_t FastMinPixelDiff() const { return 15; }

private:
    bool EnsureContext() const;
    bool Fail(const char *what) const;
    void MaterialiseCandidates() const;
    void MaterialiseMask() const;

    Options options_;
    mutable std::vector<std::pair<float, Pixel>> candidates_;
    mutable MatInt mask_;

    int device_ = 0;
    mutable fd_context *ctx_ = nullptr;
    mutable std::string last_error_;
    // what mask_ / candidates_ have to be rebuilt from
    mutable bool candidates_stale_ = false, mask_stale_ = false;
    mutable int32_t mask_rows_ = 0, mask_cols_ = 0, mask_distance_ = 0;
    mutable std::vector<std::pair<int32_t, int32_t>> mask_squares_;  // (row, col) centres cleared in the mask
};

class FeaturePointHarrisDetector : public FeaturePointDetector {
public:
    struct SubOptions {
        float kAlpha = 0.04f;
        int32_t kHalfPatchSize = 1;  // the kernels implement the reference's fixed 3x3 window
    };
    std::string DetectorTypeName() const override { return "Harris"; }
    SubOptions &sub_options() { return sub_options_; }
    const SubOptions &sub_options() const { return sub_options_; }

protected:
    int32_t DetectorKind() const override { return 0; }
    float HarrisAlpha() const override { return sub_options_.kAlpha; }

private:
    SubOptions sub_options_;
};

class FeaturePointShiTomasDetector : public FeaturePointDetector {
public:
    struct SubOptions {
        int32_t kHalfPatchSize = 1;
    };
    std::string DetectorTypeName() const override { return "Shi-Tomas"; }
    SubOptions &sub_options() { return sub_options_; }
    const SubOptions &sub_options() const { return sub_options_; }

protected:
    int32_t DetectorKind() const override { return 1; }

private:
    SubOptions sub_options_;
};

class FeaturePointFastDetector : public FeaturePointDetector {
public:
    struct SubOptions {
        int32_t kN = 12;
        uint8_t kMinPixelDiffValue = 15;
    };
    std::string DetectorTypeName() const override { return "Fast"; }
    SubOptions &sub_options() { return sub_options_; }
    const SubOptions &sub_options() const { return sub_options_; }

protected:
    int32_t DetectorKind() const override { return 2; }
    int32_t FastN() const override { return sub_options_.kN; }
    int32_t FastMinPixelDiff() const override { return sub_options_.kMinPixelDiffValue; }

private:
    SubOptions sub_options_;
};

}  // namespace feature_detector

#endif  // FD_B200_FEATURE_POINT_DETECTOR_H_
