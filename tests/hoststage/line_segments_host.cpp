// TEST SCAFFOLDING (not shipped).  Host stage behind the line detector's dense stage: seeds in, line segments out.  The dense stage that feeds it runs on the
// GPU (feature_line_field.cpp); everything here is pointer chasing over a few thousand pixels and stays on the host, as in
// the reference (src/feature_line_detector/feature_line_detector.cpp:12-54, 99-228), whose arithmetic order is kept so that
// the same seeds give the same segments.
#include <cmath>

#include "feature_line_detector.h"

namespace feature_detector {

namespace {
// 8-neighbourhood in the order the reference visits it (.cpp:112-119): row above left to right, same row, row below.
constexpr int kNeighbourRow[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
constexpr int kNeighbourCol[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
}  // namespace

FeatureLineDetector::FeatureLineDetector() { host_.seeds.reserve(10000); }

bool FeatureLineDetector::ComputeLineLevelAngleMap(const GrayImage &image) {
    field_.options().kMinValidGradientNorm = options_.kMinValidGradientNorm;
    if (!field_.Compute(image)) return false;                       // kernel 5 + seed order on the GPU
    field_.FillPixelParams(host_.grid, host_.seeds);                // what .cpp:56-97 leaves behind
    if (options_.kHostLibmAngles) {
        // same expression as .cpp:76-85 for the valid pixels only (a few percent of the frame)
        const int32_t cols = image.cols();
        const uint8_t *data = image.data();
        const PixelParam *first = host_.grid.data(), *last = host_.grid.data() + host_.grid.size();
        for (PixelParam *px : host_.seeds) {
            if (px < first || px >= last) continue;   // an entry left over from a call on a differently sized frame (.cpp:7-10 never clears)
            const int32_t r = px->row, c = px->col;
            const int32_t ad = int32_t(data[(r + 1) * cols + c + 1]) - int32_t(data[r * cols + c]);
            const int32_t bc = int32_t(data[r * cols + c + 1]) - int32_t(data[(r + 1) * cols + c]);
            const float gx = static_cast<float>(ad + bc) / 2.0f;
            const float gy = static_cast<float>(ad - bc) / 2.0f;
            px->line_level_angle = std::atan2(gx, -gy);
        }
    }
    return true;
}

bool FeatureLineDetector::DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec4> &features) {
    if (image.data() == nullptr || image.rows() < 2 || image.cols() < 2) return false;   // .cpp:14
    if (needed_feature_num == 0) return true;                                           // .cpp:15

    // Smallest region that counts (.cpp:18-20), with the reference's mix of float and double intermediates.
    const float p = options_.kMinToleranceAngleResidualInRad / kPai;
    const float log_nt = static_cast<float>(5.0f * (std::log10(double(image.cols())) + std::log10(double(image.rows()))) / 2.0f + std::log10(11.0f));
    const uint32_t min_region_size = static_cast<uint32_t>(-log_nt / std::log10(p));

    if (!ComputeLineLevelAngleMap(image)) return false;

    RegionParam region;
    host_.segments.clear();
    for (PixelParam *seed : host_.seeds) {                                           // .cpp:27-46
        if (!seed->is_valid || seed->is_used) continue;
        GrowRegion(*seed, region);
        if (region.pixels.size() < min_region_size) {
            for (PixelParam *member : region.pixels) member->is_used = false;
            continue;
        }
        RectangleParam rect = FitRectangle(region);
        if (rect.length < options_.kMinValidLineLengthInPixel || rect.inlier_ratio < options_.kMaxToleranceInlierRation) continue;
        rect.start_point += Vec2::Constant(0.5f);                                       // .cpp:43-44
        rect.end_point += Vec2::Constant(0.5f);
        host_.segments.emplace_back(rect);
    }

    features.clear();
    for (const RectangleParam &rect : host_.segments)
        features.emplace_back(Vec4(rect.start_point.x(), rect.start_point.y(), rect.end_point.x(), rect.end_point.y()));
    return true;
}

void FeatureLineDetector::Enqueue(PixelParam &neighbour) {                              // .cpp:156-161
    if (neighbour.is_occupied || neighbour.is_used || !neighbour.is_valid) return;
    neighbour.is_occupied = true;
    host_.frontier.PushBack(&neighbour);
}

void FeatureLineDetector::GrowRegion(PixelParam &seed, RegionParam &region) {           // .cpp:99-154
    host_.frontier.Clear();
    host_.touched.Clear();
    host_.touched.PushBack(&seed);
    seed.is_occupied = true;

    region.pixels.clear();
    region.angle = seed.line_level_angle;
    float sum_dx = std::cos(seed.line_level_angle);
    float sum_dy = std::sin(seed.line_level_angle);
    for (int k = 0; k < 8; ++k) Enqueue(host_.grid(seed.row + kNeighbourRow[k], seed.col + kNeighbourCol[k]));

    while (!host_.frontier.Empty()) {
        PixelParam *px = host_.frontier.Front();
        host_.frontier.PopFront();
        host_.touched.PushBack(px);
        const float residual = Utility::AngleDiffInRad(region.angle, px->line_level_angle);
        if (std::fabs(residual) > options_.kMinToleranceAngleResidualInRad) continue;
        sum_dx += std::cos(px->line_level_angle);
        sum_dy += std::sin(px->line_level_angle);
        region.angle = std::atan2(sum_dy, sum_dx);
        region.pixels.emplace_back(px);
        px->is_used = true;
        for (int k = 0; k < 8; ++k) Enqueue(host_.grid(px->row + kNeighbourRow[k], px->col + kNeighbourCol[k]));
    }
    while (!host_.touched.Empty()) {
        host_.touched.Front()->is_occupied = false;
        host_.touched.PopFront();
    }
}

FeatureLineDetector::RectangleParam FeatureLineDetector::FitRectangle(const RegionParam &region) {   // .cpp:163-228
    RectangleParam rect;
    float weight = 0.0f;
    for (const PixelParam *px : region.pixels) {                                        // norm-weighted centroid
        rect.center_point.x() += static_cast<float>(px->col) * px->gradient_norm;
        rect.center_point.y() += static_cast<float>(px->row) * px->gradient_norm;
        weight += px->gradient_norm;
    }
    if (weight == 0) return rect;
    rect.center_point /= weight;

    float ixx = 0.0f, iyy = 0.0f, ixy = 0.0f;                                           // weighted second moments
    for (const PixelParam *px : region.pixels) {
        const float dx = px->col - rect.center_point.x();
        const float dy = px->row - rect.center_point.y();
        ixx += dy * dy * px->gradient_norm;
        iyy += dx * dx * px->gradient_norm;
        ixy -= dx * dy * px->gradient_norm;
    }
    if (ixx == 0 || iyy == 0 || ixy == 0) return rect;
    const float lambda = 0.5f * (ixx + iyy - std::sqrt((ixx - iyy) * (ixx - iyy) + 4.0f * ixy * ixy));   // smaller eigenvalue
    rect.angle = std::fabs(ixx) > std::fabs(iyy) ? std::atan2(lambda - ixx, ixy) : std::atan2(ixy, lambda - iyy);
    if (std::fabs(Utility::AngleDiffInRad(rect.angle, region.angle)) > options_.kMinToleranceAngleResidualInRad) {
        rect.angle += kPai;
        if (rect.angle >= kPai) rect.angle -= k2Pai;
    }
    rect.dir_vector.x() = std::cos(rect.angle);
    rect.dir_vector.y() = std::sin(rect.angle);

    Vec2 along = Vec2::Zero(), across = Vec2::Zero();                                   // (min, max) extents, both start at 0
    for (const PixelParam *px : region.pixels) {
        const float dx = px->col - rect.center_point.x();
        const float dy = px->row - rect.center_point.y();
        const float l = dx * rect.dir_vector.x() + dy * rect.dir_vector.y();
        const float w = -dx * rect.dir_vector.y() + dy * rect.dir_vector.x();
        along(0) = std::min(along(0), l);
        along(1) = std::max(along(1), l);
        across(0) = std::min(across(0), w);
        across(1) = std::max(across(1), w);
    }
    rect.start_point = rect.center_point + along(0) * rect.dir_vector;
    rect.end_point = rect.center_point + along(1) * rect.dir_vector;
    rect.length = std::max(along(1) - along(0), 1.0f);                                  // at least one pixel each way
    rect.width = std::max(across(1) - across(0), 1.0f);
    const float area = (along(1) - along(0)) * rect.width;
    rect.inlier_ratio = static_cast<float>(region.pixels.size()) / area;
    return rect;
}

}  // namespace feature_detector
