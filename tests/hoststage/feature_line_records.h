// Record types of the line detector's public surface, kept layout- and name-compatible with what callers of the reference
// read through FeatureLineDetector::pixels() / rectangles() (reference feature_line_detector.h:14-45): the line demo walks
// pixels() field by field and draws rectangles()[i].start_point / end_point (test/test_feature_line_detector.cpp:16,77).
// They live at namespace scope here so that the GPU field stage (feature_line_field.h) and the host stage share one
// definition; FeatureLineDetector re-exports them under the reference's nested names.
#ifndef FD_B200_FEATURE_LINE_RECORDS_H_
#define FD_B200_FEATURE_LINE_RECORDS_H_

#include <cstdint>
#include <vector>

#include "basic_type.h"
#include "slam_basic_math.h"

namespace feature_detector {
namespace line_records {

// One pixel of the level-line field: 20 bytes, column-major in pixels().  The three flags are the region-growing state:
// valid = gradient norm above Options::kMinValidGradientNorm; used = member of an accepted region; occupied = queued or
// visited while the current region grows (cleared when the region is abandoned or accepted).
struct Pixel {
    int32_t row = 0, col = 0;
    float line_level_angle = 0.0f, gradient_norm = 0.0f;
    bool is_valid = false, is_used = false, is_occupied = false;
};
static_assert(sizeof(Pixel) == 20, "PixelParam must keep the reference's 20-byte layout");

struct Region {
    std::vector<Pixel *> pixels;
    float angle = 0.0f;
};

// A fitted segment.  Member order follows the reference so that aggregate users see the same layout.
struct Rectangle {
    Vec2 start_point = Vec2::Zero(), end_point = Vec2::Zero(), center_point = Vec2::Zero();
    float length = 0.0f, width = 0.0f, angle = 0.0f;
    Vec2 dir_vector = Vec2::Identity();
    float inlier_ratio = 0.0f;
};

// Thresholds of the detector, reference names and defaults (feature_line_detector.h:40-45), plus one addition.
struct Thresholds {
    float kMinValidGradientNorm = 20.0f;
    float kMinToleranceAngleResidualInRad = 22.5f * kDegToRad;
    float kMinValidLineLengthInPixel = 20.0f;
    float kMaxToleranceInlierRation = 0.6f;
    // Addition: take the level-line angle of the valid pixels from the host libm (std::atan2, as the reference does)
    // instead of the GPU's arctangent.  The two agree within 1e-5; with this set the host stage sees bit-identical input.
    bool kHostLibmAngles = true;
};

}  // namespace line_records
}  // namespace feature_detector

#endif  // FD_B200_FEATURE_LINE_RECORDS_H_
