#include "feature_line_field.h"

#include <cstdlib>

#include "fd_b200.h"

namespace feature_detector {

void *LineLevelAngleField::AllocHost(size_t bytes) {
    void *p = nullptr;
    return fd_host_alloc(&p, bytes) == FD_OK ? p : nullptr;
}

void LineLevelAngleField::FreeHost(void *ptr) { fd_host_free(ptr); }

LineLevelAngleField::~LineLevelAngleField() {
    if (ctx_ != nullptr) fd_destroy(ctx_);
}

bool LineLevelAngleField::Compute(const GrayImage &image) {
    if (image.data() == nullptr || image.rows() < 2 || image.cols() < 2) return false;      // feature_line_detector.cpp:14
    if (ctx_ == nullptr) {
        int device = device_;
        if (const char *env = std::getenv("FD_B200_DEVICE")) device = std::atoi(env);
        if (fd_create(device, &ctx_) != FD_OK) {
            ctx_ = nullptr;
            last_error_ = "fd_create failed (no CUDA device? this library has no CPU path)";
            return false;
        }
    }
    rows_ = image.rows();
    cols_ = image.cols();
    const size_t px = size_t(rows_) * size_t(cols_);
    if (!norm_.reserve(px) || !angle_.reserve(px) || !seeds_.reserve(px)) {
        last_error_ = "fd_host_alloc failed";
        return false;
    }
    norm_.set_size(px);
    angle_.set_size(px);
    seeds_.set_size(0);
    fd_lsd_params p = {};
    p.min_valid_gradient_norm = options_.kMinValidGradientNorm;
    p.want_sorted = 1;
    int32_t n_valid = 0;
    if (fd_upload_frames(ctx_, image.data(), rows_, cols_, 1) != FD_OK || fd_lsd_field(ctx_, &p, nullptr, nullptr, nullptr, nullptr) != FD_OK ||
        fd_lsd_download(ctx_, 0, norm_.data(), angle_.data(), seeds_.data(), int64_t(px), &n_valid) != FD_OK) {
        last_error_ = fd_last_error(ctx_);
        return false;
    }
    seeds_.set_size(size_t(n_valid));
    return true;
}

}  // namespace feature_detector
