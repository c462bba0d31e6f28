// Drop-in for the reference's src/feature_descriptor/descriptor_brief.h:10-37: steered BRIEF, up to 256 bits,
// std::vector<bool> per keypoint.  All keypoints of an image are described by one GPU call (kernel 4,
// csrc/fd_brief.cu) through the C ABI; there is no CPU path.
#ifndef FD_B200_DESCRIPTOR_BRIEF_H_
#define FD_B200_DESCRIPTOR_BRIEF_H_

#include <string>

#include "descriptor.h"

struct fd_context;

namespace feature_detector {

using BriefType = std::vector<bool>;

class BriefDescriptor : public Descriptor<BriefType> {
public:
    struct Options {
        int32_t kLength = 256;
        int32_t kHalfPatchSize = 8;
        // Addition: how the absent upstream Image::GetPixelValueNoCheck(float, float) samples.
        // 0 = four-tap bilinear (the mode parity is defined against), 1 = truncate to the pixel.
        int32_t kSampling = 0;
    };

    BriefDescriptor() = default;
    ~BriefDescriptor() override;
    BriefDescriptor(const BriefDescriptor &) = delete;
    BriefDescriptor &operator=(const BriefDescriptor &) = delete;

    Options &options() { return options_; }
    const Options &options() const { return options_; }

    // Packed form (32 bytes per keypoint, bit i at byte i/8, bit i%8) for callers that match with popcounts.
    bool ComputePacked(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<uint8_t> &packed) const;

    void set_device(int ordinal) { device_ = ordinal; }
    const std::string &last_error() const { return last_error_; }

protected:
    bool ComputeForOneFeature(const GrayImage &image, const Vec2 &pixel_uv, BriefType &descriptor) const override;
    bool ComputeAll(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<BriefType> &descriptors) const override;

private:
    bool EnsureContext() const;

    Options options_;
    int device_ = 0;
    mutable fd_context *ctx_ = nullptr;
    mutable std::string last_error_;
};

}  // namespace feature_detector

#endif  // FD_B200_DESCRIPTOR_BRIEF_H_
