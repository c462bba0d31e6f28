// Drop-in for the reference's src/feature_descriptor/descriptor.h:13-62: abstract Descriptor<T> with the two
// Compute overloads.  The reference loops over keypoints calling a virtual per-feature function; here the
// per-batch step is the virtual one (ComputeAll), so that BriefDescriptor can hand the whole list to the GPU in
// one call, while a user-defined descriptor that only overrides ComputeForOneFeature still works through the
// default ComputeAll loop.
#ifndef FD_B200_DESCRIPTOR_H_
#define FD_B200_DESCRIPTOR_H_

#include <cstdint>
#include <type_traits>
#include <vector>

#include "basic_type.h"
#include "datatype_image.h"

namespace feature_detector {

template <typename DescriptorType>
class Descriptor {
public:
    Descriptor() = default;
    virtual ~Descriptor() = default;

    // descriptor.h:28-40: false only for an empty keypoint list or a null image.
    bool Compute(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<DescriptorType> &descriptors) const {
        if (pixel_uv.empty() || image.data() == nullptr) return false;
        if (descriptors.size() != pixel_uv.size()) descriptors.resize(pixel_uv.size());
        return ComputeAll(image, pixel_uv, descriptors);
    }

    // descriptor.h:43-62: bits become +1 / -1, other element types are cast to float.
    bool Compute(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<Vec> &descriptors) const {
        std::vector<DescriptorType> typed;
        if (!Compute(image, pixel_uv, typed)) return false;
        descriptors.resize(typed.size());
        for (size_t i = 0; i < typed.size(); ++i) {
            const auto &src = typed[i];
            Vec &dst = descriptors[i];
            dst.setZero(int(src.size()), 1);
            for (size_t j = 0; j < src.size(); ++j) {
                if constexpr (std::is_same_v<DescriptorType, std::vector<bool>>) dst[int(j)] = src[j] ? 1.0f : -1.0f;
                else dst[int(j)] = static_cast<float>(src[j]);
            }
        }
        return true;
    }

protected:
    // One keypoint.  Return value ignored by the loop, as in the reference (descriptor.h:36).
    virtual bool ComputeForOneFeature(const GrayImage &image, const Vec2 &pixel_uv, DescriptorType &descriptor) const = 0;
    // All keypoints of one image; `descriptors` is already sized.
    virtual bool ComputeAll(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<DescriptorType> &descriptors) const {
        for (size_t i = 0; i < pixel_uv.size(); ++i) ComputeForOneFeature(image, pixel_uv[i], descriptors[i]);
        return true;
    }
};

}  // namespace feature_detector

#endif  // FD_B200_DESCRIPTOR_H_
