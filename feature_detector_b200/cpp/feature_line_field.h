// GPU replacement for FeatureLineDetector::ComputeLineLevelAngleMap
// (reference src/feature_line_detector/feature_line_detector.cpp:56-97).
//
// The reference's line detector keeps region growing, rectangle fitting and validation (:99-228) in host code, and the
// north star leaves them there.  What moves to the GPU is the dense stage: the 2x2 gradient, its norm, the validity
// test, the level-line angle and the ordering of the valid pixels by norm (kernel 5, csrc/fd_lsd.cu).  This class runs
// that stage and then leaves the host stage's input exactly as ComputeLineLevelAngleMap would have left it:
// FillPixelParams() writes the reference's column-major PixelParam matrix and its sorted pointer list, so the body of
// ComputeLineLevelAngleMap in a reference checkout becomes the two calls shown in INTEGRATION.md.
#ifndef FD_B200_FEATURE_LINE_FIELD_H_
#define FD_B200_FEATURE_LINE_FIELD_H_

#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "basic_type.h"
#include "datatype_image.h"

struct fd_context;

namespace feature_detector {

class LineLevelAngleField {
public:
    // A page-locked array that only ever grows and is never initialised: the maps come back from the GPU 16 bytes per pixel on
    // every call, which from pageable memory costs five times the copy itself (fd_host_alloc, include/fd_b200.h).
    template <typename T>
    class HostArray {
    public:
        HostArray() = default;
        ~HostArray() { LineLevelAngleField::FreeHost(data_); }
        HostArray(const HostArray &) = delete;
        HostArray &operator=(const HostArray &) = delete;
        bool reserve(size_t n) {   // contents are lost when the array has to grow
            if (n <= capacity_) return true;
            LineLevelAngleField::FreeHost(data_);
            data_ = static_cast<T *>(LineLevelAngleField::AllocHost(n * sizeof(T)));
            capacity_ = data_ != nullptr ? n : 0;
            size_ = 0;
            return data_ != nullptr;
        }
        void set_size(size_t n) { size_ = n; }
        size_t size() const { return size_; }
        bool empty() const { return size_ == 0; }
        const T *data() const { return data_; }
        T *data() { return data_; }
        const T &operator[](size_t i) const { return data_[i]; }
        const T *begin() const { return data_; }
        const T *end() const { return data_ + size_; }

    private:
        T *data_ = nullptr;
        size_t size_ = 0, capacity_ = 0;
    };

    struct Options {
        float kMinValidGradientNorm = 20.0f;  // feature_line_detector.h:41
    };

    LineLevelAngleField() = default;
    ~LineLevelAngleField();
    LineLevelAngleField(const LineLevelAngleField &) = delete;
    LineLevelAngleField &operator=(const LineLevelAngleField &) = delete;

    Options &options() { return options_; }
    const Options &options() const { return options_; }

    // false for a null image or rows / cols < 2 (feature_line_detector.cpp:14), or when the GPU call fails.
    bool Compute(const GrayImage &image);

    // Results of the last Compute: maps are image_rows x image_cols floats, row-major (last row / column zero);
    // seeds are row * image_cols + col of the valid pixels, norm descending, ties in the reference's push order.
    int32_t image_rows() const { return rows_; }
    int32_t image_cols() const { return cols_; }
    const HostArray<float> &gradient_norm() const { return norm_; }
    const HostArray<float> &line_level_angle() const { return angle_; }
    const HostArray<int32_t> &sorted_seeds() const { return seeds_; }

    // Leave `pixels` (Eigen::Matrix<PixelParam, Dynamic, Dynamic>) and `sorted` (std::vector<PixelParam *>) as the
    // reference's ComputeLineLevelAngleMap leaves pixels_ and sorted_pixels_, including what it does NOT touch:
    // a resize to an unchanged size keeps old flags, invalid pixels keep their old angle, and `sorted` is appended to.
    template <typename PixelMatrix, typename PixelPtrVector>
    void FillPixelParams(PixelMatrix &pixels, PixelPtrVector &sorted) const {
        const int32_t pr = rows_ - 1, pc = cols_ - 1;
        // Deliberate deviation: a frame of another size reallocates the matrix, and the pointers the reference would keep in
        // sorted_pixels_ dangle from then on (undefined behaviour there, so no parity to keep) -- they are dropped here.
        if (pixels.rows() != pr || pixels.cols() != pc) sorted.clear();
        pixels.resize(pr, pc);                                                    // .cpp:58
        for (int32_t i = 0; i < pr; ++i) {                                        // .cpp:59-63
            pixels(i, 0).row = i;
            pixels(i, pc - 1).row = i;
            pixels(i, pc - 1).col = pc - 1;
        }
        for (int32_t i = 0; i < pc; ++i) {                                        // .cpp:64-68 (the (0, 1) index is the reference's)
            if (pc > 1) pixels(0, 1).col = i;                                     // two-column images: the reference writes out of bounds here
            pixels(pr - 1, i).col = i;
            pixels(pr - 1, i).row = pr - 1;
        }
        // .cpp:71-89.  The reference walks columns outer / rows inner over a row-major image and a column-major matrix; every
        // pixel's record is written independently of the others, so the same records are produced here tile by tile (the rows of a
        // tile stay in cache while its columns are walked) and, for large frames, by a few threads that split the columns.
        const int32_t c_lo = 1, c_hi = cols_ - 2, r_lo = 1, r_hi = rows_ - 2;
        const float min_norm = options_.kMinValidGradientNorm;
        auto fill_columns = [&](int32_t col_begin, int32_t col_end) {
            constexpr int32_t kTileCols = 16, kTileRows = 256;
            for (int32_t col0 = col_begin; col0 < col_end; col0 += kTileCols) {
                const int32_t col1 = std::min(col0 + kTileCols, col_end);
                for (int32_t row0 = r_lo; row0 < r_hi; row0 += kTileRows) {
                    const int32_t row1 = std::min(row0 + kTileRows, r_hi);
                    for (int32_t col = col0; col < col1; ++col) {
                        for (int32_t row = row0; row < row1; ++row) {
                            auto &px = pixels(row, col);
                            const size_t at = size_t(row) * size_t(cols_) + size_t(col);
                            px.row = row;
                            px.col = col;
                            px.gradient_norm = norm_[at];
                            px.is_valid = norm_[at] > min_norm;
                            if (px.is_valid) px.line_level_angle = angle_[at];
                        }
                    }
                }
            }
        };
        const int64_t interior = int64_t(std::max(c_hi - c_lo, 0)) * std::max(r_hi - r_lo, 0);
        const unsigned n_threads = interior < (int64_t(1) << 18) ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        if (n_threads <= 1 || c_hi <= c_lo) {
            if (c_hi > c_lo) fill_columns(c_lo, c_hi);
        } else {
            std::vector<std::thread> workers;
            const int32_t per = (c_hi - c_lo + int32_t(n_threads) - 1) / int32_t(n_threads);
            for (unsigned t = 0; t < n_threads; ++t) {
                const int32_t b = c_lo + int32_t(t) * per, e = std::min(b + per, c_hi);
                if (b < e) workers.emplace_back(fill_columns, b, e);
            }
            for (std::thread &w : workers) w.join();
        }
        const size_t before = sorted.size();
        for (const int32_t s : seeds_) sorted.emplace_back(&pixels(s / cols_, s % cols_));   // .cpp:86, already in :92-94 order
        if (before != 0) {
            // the reference never clears sorted_pixels_ between calls and sorts the whole list again
            std::stable_sort(sorted.begin(), sorted.end(), [](const auto *a, const auto *b) { return a->gradient_norm > b->gradient_norm; });
        }
    }

    void set_device(int ordinal) { device_ = ordinal; }
    const std::string &last_error() const { return last_error_; }

private:
    Options options_;
    int32_t rows_ = 0, cols_ = 0;
    static void *AllocHost(size_t bytes);   // fd_host_alloc / fd_host_free (feature_line_field.cpp)
    static void FreeHost(void *ptr);
    HostArray<float> norm_, angle_;
    HostArray<int32_t> seeds_;
    int device_ = 0;
    fd_context *ctx_ = nullptr;
    std::string last_error_;
};

}  // namespace feature_detector

#endif  // FD_B200_FEATURE_LINE_FIELD_H_
