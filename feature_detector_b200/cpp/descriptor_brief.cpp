// Host glue of the drop-in BRIEF descriptor: marshalling around fd_describe_points.  The pattern table and all
// sampling live in the CUDA library.
#include "descriptor_brief.h"

#include <cstdlib>

#include "fd_b200.h"

namespace feature_detector {

BriefDescriptor::~BriefDescriptor() {
    if (ctx_ != nullptr) fd_destroy(ctx_);
}

bool BriefDescriptor::EnsureContext() const {
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

bool BriefDescriptor::ComputePacked(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<uint8_t> &packed) const {
    if (pixel_uv.empty() || image.data() == nullptr) return false;
    if (!EnsureContext()) return false;
    const int n = int(pixel_uv.size());
    std::vector<float> xy(size_t(n) * 2);
    for (int i = 0; i < n; ++i) {
        xy[2 * i] = pixel_uv[i].x();
        xy[2 * i + 1] = pixel_uv[i].y();
    }
    fd_brief_params p = {};
    p.length = options_.kLength;
    p.half_patch_size = options_.kHalfPatchSize;
    p.sampling = options_.kSampling;
    const int32_t count = n;
    packed.assign(size_t(n) * 32, 0);
    if (fd_upload_frames(ctx_, image.data(), image.rows(), image.cols(), 1) != FD_OK ||
        fd_describe_points(ctx_, &p, xy.data(), &count, n, 1) != FD_OK || fd_download_descriptors(ctx_, packed.data(), n) != FD_OK) {
        last_error_ = fd_last_error(ctx_);
        return false;
    }
    return true;
}

bool BriefDescriptor::ComputeAll(const GrayImage &image, const std::vector<Vec2> &pixel_uv, std::vector<BriefType> &descriptors) const {
    std::vector<uint8_t> packed;
    if (!ComputePacked(image, pixel_uv, packed)) return false;
    const int32_t length = options_.kLength;
    for (size_t i = 0; i < pixel_uv.size(); ++i) {
        BriefType &d = descriptors[i];
        d.assign(size_t(length), false);                                          // descriptor_brief.cpp:10
        const uint8_t *bytes = packed.data() + i * 32;
        for (int32_t b = 0; b < length; ++b) d[size_t(b)] = (bytes[b >> 3] >> (b & 7)) & 1u;
    }
    return true;
}

bool BriefDescriptor::ComputeForOneFeature(const GrayImage &image, const Vec2 &pixel_uv, BriefType &descriptor) const {
    std::vector<BriefType> one(1);
    if (!ComputeAll(image, std::vector<Vec2>{pixel_uv}, one)) return false;
    descriptor.swap(one[0]);
    return true;
}

}  // namespace feature_detector
