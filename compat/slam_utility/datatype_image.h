// Stand-in for Slam_Utility/src/data_type/image/datatype_image.h (absent here).
// Image<T> is a view of a contiguous row-major buffer with no pitch.
#ifndef FD_COMPAT_DATATYPE_IMAGE_H_
#define FD_COMPAT_DATATYPE_IMAGE_H_

#include <cstdlib>

#include "basic_type.h"

template <typename Scalar>
class Image {
public:
    Image() = default;
    Image(Scalar *data, int32_t rows, int32_t cols, bool is_owner = false): data_(data), rows_(rows), cols_(cols), owner_(is_owner) {}
    Image(const Image &) = delete;
    Image &operator=(const Image &) = delete;
    ~Image() {
        if (owner_ && data_ != nullptr) std::free(data_);
    }

    void SetImage(Scalar *data, int32_t rows, int32_t cols, bool is_owner = false) {
        if (owner_ && data_ != nullptr) std::free(data_);
        data_ = data;
        rows_ = rows;
        cols_ = cols;
        owner_ = is_owner;
    }

    Scalar *data() const { return data_; }
    int32_t rows() const { return rows_; }
    int32_t cols() const { return cols_; }

    // Integer coordinates: plain fetch, converted to T (fast detector .cpp:12, line detector .cpp:76-79).
    template <typename T = Scalar>
    T GetPixelValueNoCheck(int32_t row, int32_t col) const {
        return static_cast<T>(data_[row * cols_ + col]);
    }

    // Float coordinates.  GUESS G1: four-tap bilinear interpolation around the
    // truncated coordinate, in the author's own operand order
    // (reference src/nn_feature_point_detector/nn_feature_point_detector.cpp:169-189 uses the same idiom).
    // BRIEF parity is defined against THIS; the real upstream overload is unverified.
    float GetPixelValueNoCheck(float row, float col) const {
        const Scalar *p = data_ + static_cast<int32_t>(row) * cols_ + static_cast<int32_t>(col);
        const float sub_row = row - std::floor(row);
        const float sub_col = col - std::floor(col);
        const float inv_sub_row = 1.0f - sub_row;
        const float inv_sub_col = 1.0f - sub_col;
        return static_cast<float>(inv_sub_col * inv_sub_row * p[0] + sub_col * inv_sub_row * p[1] +
                                  inv_sub_col * sub_row * p[cols_] + sub_col * sub_row * p[cols_ + 1]);
    }

    void SetPixelValueNoCheck(int32_t row, int32_t col, Scalar v) { data_[row * cols_ + col] = v; }

private:
    Scalar *data_ = nullptr;
    int32_t rows_ = 0;
    int32_t cols_ = 0;
    bool owner_ = false;
};

using GrayImage = Image<uint8_t>;

#endif  // FD_COMPAT_DATATYPE_IMAGE_H_
