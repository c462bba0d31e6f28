// TEST SCAFFOLDING (not shipped): stand-in for the reference's src/feature_line_detector/feature_line_detector.h:12-79 -- same class name, nested type names,
// Options and accessors, so a caller recompiles unchanged.  The dense stage (ComputeLineLevelAngleMap, .cpp:56-97) runs on
// the GPU through LineLevelAngleField; region growing, rectangle fitting and validation (.cpp:99-228) are host code here as
// they are in the reference -- the north star leaves them on the host -- written against the same PixelParam array the
// accessors expose.  "circular_buffer.h" and "slam_basic_math.h" are the caller's Slam_Utility headers (compat/slam_utility/
// has stand-ins), so the two upstream semantics the host stage depends on are whatever the caller's
// headers define.
#ifndef FD_B200_FEATURE_LINE_DETECTOR_H_
#define FD_B200_FEATURE_LINE_DETECTOR_H_

#include <vector>

#include "circular_buffer.h"
#include "datatype_image.h"
#include "feature_line_field.h"
#include "feature_line_records.h"

namespace feature_detector {

class FeatureLineDetector {
public:
    // the reference's nested names
    using PixelParam = line_records::Pixel;
    using RegionParam = line_records::Region;
    using RectangleParam = line_records::Rectangle;
    using Options = line_records::Thresholds;
    using PixelGrid = Eigen::Matrix<PixelParam, Eigen::Dynamic, Eigen::Dynamic>;   // column-major, (rows - 1) x (cols - 1)
    using PixelList = std::vector<PixelParam *>;
    using RectangleList = std::vector<RectangleParam>;

    FeatureLineDetector();
    virtual ~FeatureLineDetector() = default;

    // false for a null image or one with fewer than two rows / columns (.cpp:14-15), and on any CUDA failure
    bool DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec4> &features);

    // state the reference exposes after a call, mutable and const
    Options &options() { return options_; }
    const Options &options() const { return options_; }
    PixelGrid &pixels() { return host_.grid; }
    const PixelGrid &pixels() const { return host_.grid; }
    PixelList &sorted_pixels() { return host_.seeds; }
    const PixelList &sorted_pixels() const { return host_.seeds; }
    RectangleList &rectangles() { return host_.segments; }
    const RectangleList &rectangles() const { return host_.segments; }

    LineLevelAngleField &field() { return field_; }   // GPU plumbing (device selection, last error)

private:
    // what the host stage works on: the field as the reference lays it out, the seed order, two bounded work queues
    struct HostState {
        PixelGrid grid;
        PixelList seeds;
        CircularBuffer<PixelParam *, 1000> frontier;   // pixels waiting to be examined (the reference's candidates_)
        CircularBuffer<PixelParam *, 1000> touched;    // pixels whose occupied flag must be cleared (visited_pixels_)
        RectangleList segments;
    };

    bool ComputeLineLevelAngleMap(const GrayImage &image);
    void GrowRegion(PixelParam &seed, RegionParam &region);
    void Enqueue(PixelParam &neighbour);
    RectangleParam FitRectangle(const RegionParam &region);

    Options options_;
    LineLevelAngleField field_;
    HostState host_;
};

}  // namespace feature_detector

#endif  // FD_B200_FEATURE_LINE_DETECTOR_H_
