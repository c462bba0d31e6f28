// Drop-in for the reference's src/feature_line_detector/feature_line_detector.h:12-79: same class name, nested types,
// Options and accessors.  The dense stage (ComputeLineLevelAngleMap, .cpp:56-97) runs on the GPU through
// LineLevelAngleField; region growing, rectangle fitting and validation (.cpp:99-228) are host code here as they are in the
// reference -- the north star leaves them on the host -- written against the same PixelParam array the accessors expose.
// "circular_buffer.h" and "slam_basic_math.h" are the caller's Slam_Utility headers (compat/slam_utility/ has stand-ins), so
// the two upstream semantics the host stage depends on are whatever the caller's headers define.
#ifndef FD_B200_FEATURE_LINE_DETECTOR_H_
#define FD_B200_FEATURE_LINE_DETECTOR_H_

#include <vector>

#include "basic_type.h"
#include "circular_buffer.h"
#include "datatype_image.h"
#include "feature_line_field.h"
#include "slam_basic_math.h"

namespace feature_detector {

class FeatureLineDetector {
public:
    struct PixelParam {
        int32_t row = 0;
        int32_t col = 0;
        float line_level_angle = 0.0f;
        float gradient_norm = 0.0f;
        bool is_valid = false;     // gradient norm above Options::kMinValidGradientNorm
        bool is_used = false;      // already part of an accepted region
        bool is_occupied = false;  // queued or visited while the current region grows
    };
    struct RegionParam {
        std::vector<PixelParam *> pixels;
        float angle = 0.0f;
    };
    struct RectangleParam {
        Vec2 start_point = Vec2::Zero();
        Vec2 end_point = Vec2::Zero();
        Vec2 center_point = Vec2::Zero();
        float length = 0.0f;
        float width = 0.0f;
        float angle = 0.0f;
        Vec2 dir_vector = Vec2::Identity();
        float inlier_ratio = 0.0f;
    };
    struct Options {
        float kMinValidGradientNorm = 20.0f;
        float kMinToleranceAngleResidualInRad = 22.5f * kDegToRad;
        float kMinValidLineLengthInPixel = 20.0f;
        float kMaxToleranceInlierRation = 0.6f;
        // Addition: take the level-line angle of the valid pixels from the host libm (std::atan2, as the reference does)
        // instead of the GPU's atan2f.  The two agree within 1e-5; with this set the host stage sees bit-identical input.
        bool kHostLibmAngles = true;
    };

    FeatureLineDetector();
    virtual ~FeatureLineDetector() = default;

    bool DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec4> &features);

    Options &options() { return options_; }
    Eigen::Matrix<PixelParam, Eigen::Dynamic, Eigen::Dynamic> &pixels() { return pixels_; }
    std::vector<PixelParam *> &sorted_pixels() { return sorted_pixels_; }
    std::vector<RectangleParam> &rectangles() { return rectangles_; }
    const Options &options() const { return options_; }
    const Eigen::Matrix<PixelParam, Eigen::Dynamic, Eigen::Dynamic> &pixels() const { return pixels_; }
    const std::vector<PixelParam *> &sorted_pixels() const { return sorted_pixels_; }
    const std::vector<RectangleParam> &rectangles() const { return rectangles_; }

    LineLevelAngleField &field() { return field_; }   // GPU plumbing (device selection, last error)

private:
    bool ComputeLineLevelAngleMap(const GrayImage &image);
    void GrowRegion(PixelParam &seed, RegionParam &region);
    void Enqueue(PixelParam &neighbour);
    RectangleParam FitRectangle(const RegionParam &region);

    Options options_;
    LineLevelAngleField field_;
    Eigen::Matrix<PixelParam, Eigen::Dynamic, Eigen::Dynamic> pixels_;
    std::vector<PixelParam *> sorted_pixels_;
    CircularBuffer<PixelParam *, 1000> frontier_;   // the reference's candidates_
    CircularBuffer<PixelParam *, 1000> touched_;    // the reference's visited_pixels_
    std::vector<RectangleParam> rectangles_;
};

}  // namespace feature_detector

#endif  // FD_B200_FEATURE_LINE_DETECTOR_H_
