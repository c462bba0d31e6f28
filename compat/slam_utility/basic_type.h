// Stand-in for Slam_Utility/src/basic_type/basic_type.h plus the small slice of Eigen the
// Feature_Detector sources touch.  Horizon1026/Slam_Utility and Eigen3 are NOT in this image
// (SURVEY.md section 8c), so this header exists for two builds only:
//   * oracle/_ref  : compiling the reference's own .cpp files in place (test infrastructure)
//   * tests/cpp    : compiling this repo's drop-in C++ classes without the real dependency
// A downstream user who has Slam_Utility + Eigen puts THOSE on the include path instead.
// Type definitions only; nothing here is on the GPU path.
#ifndef FD_COMPAT_BASIC_TYPE_H_
#define FD_COMPAT_BASIC_TYPE_H_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

enum : int { Dynamic = -1 };
enum : int { ColMajor = 0, RowMajor = 1 };

namespace detail {
// Fixed-size payload.
template <typename T, int N>
struct FixedBuf {
    T v[N > 0 ? N : 1] = {};
    void Reshape(int, int) {}
};
// Heap payload.  Like Eigen, a resize to a different element count reallocates (elements are
// value-initialised here; Eigen leaves PODs uninitialised, the reference never reads those),
// and a resize to the same element count keeps the old contents.
template <typename T>
struct HeapBuf {
    std::unique_ptr<T[]> v;
    int64_t n = 0;
    void Reshape(int r, int c) {
        const int64_t want = int64_t(r) * c;
        if (want != n) {
            v.reset(want > 0 ? new T[want]() : nullptr);
            n = want;
        }
    }
};
}  // namespace detail

template <typename T, int R, int C, int Order = ColMajor>
class Matrix {
    static constexpr bool kDyn = (R == Dynamic) || (C == Dynamic);
    using Buf = std::conditional_t<kDyn, detail::HeapBuf<T>, detail::FixedBuf<T, R * C>>;

public:
    Matrix() = default;
    Matrix(const Matrix &o) { *this = o; }
    Matrix(Matrix &&) = default;
    Matrix &operator=(Matrix &&) = default;
    Matrix &operator=(const Matrix &o) {
        if (this == &o) return *this;
        if constexpr (kDyn) {
            resize(o.rows(), o.cols());
        }
        std::copy(o.data(), o.data() + o.size(), data());
        return *this;
    }

    // Vec2-style and Vec4-style element constructors (any arithmetic argument type).
    template <typename A, typename B, int RR = R, int CC = C,
              typename = std::enable_if_t<RR == 2 && CC == 1 && std::is_arithmetic_v<A> && std::is_arithmetic_v<B>>>
    Matrix(A a, B b) {
        buf_.v[0] = static_cast<T>(a);
        buf_.v[1] = static_cast<T>(b);
    }
    template <typename A, int RR = R, int CC = C, typename = std::enable_if_t<RR == 4 && CC == 1 && std::is_arithmetic_v<A>>>
    Matrix(A a, A b, A c, A d) {
        buf_.v[0] = static_cast<T>(a);
        buf_.v[1] = static_cast<T>(b);
        buf_.v[2] = static_cast<T>(c);
        buf_.v[3] = static_cast<T>(d);
    }

    int rows() const { return R == Dynamic ? rows_ : R; }
    int cols() const { return C == Dynamic ? cols_ : C; }
    int64_t size() const { return int64_t(rows()) * cols(); }

    T *data() {
        if constexpr (kDyn) return buf_.v.get(); else return buf_.v;
    }
    const T *data() const {
        if constexpr (kDyn) return buf_.v.get(); else return buf_.v;
    }

    void resize(int r, int c) {
        buf_.Reshape(r, c);
        rows_ = r;
        cols_ = c;
    }
    void resize(int n) { resize(n, 1); }

    T &operator()(int r, int c) { return data()[Order == RowMajor ? int64_t(r) * cols() + c : int64_t(c) * rows() + r]; }
    const T &operator()(int r, int c) const { return data()[Order == RowMajor ? int64_t(r) * cols() + c : int64_t(c) * rows() + r]; }
    T &operator()(int i) { return data()[i]; }
    const T &operator()(int i) const { return data()[i]; }
    T &operator[](int i) { return data()[i]; }
    const T &operator[](int i) const { return data()[i]; }
    T &x() { return data()[0]; }
    T &y() { return data()[1]; }
    const T &x() const { return data()[0]; }
    const T &y() const { return data()[1]; }

    void setZero() { std::fill(data(), data() + size(), T(0)); }
    void setConstant(const T &v) { std::fill(data(), data() + size(), v); }
    static Matrix Ones(int r, int c) {
        Matrix m;
        m.setConstant(r, c, T(1));
        return m;
    }
    // Rectangular views that can only be cleared: all the reference does with block / topRows / ... (nn_feature_point_detector.cpp:64-83).
    class BlockView {
    public:
        BlockView(Matrix &m, int r0, int c0, int h, int w): m_(m), r0_(r0), c0_(c0), h_(h), w_(w) {}
        void setZero() {
            for (int r = r0_; r < r0_ + h_; ++r)
                for (int c = c0_; c < c0_ + w_; ++c) m_(r, c) = T(0);
        }
    private:
        Matrix &m_;
        int r0_, c0_, h_, w_;
    };
    BlockView block(int r0, int c0, int h, int w) { return BlockView(*this, r0, c0, h, w); }
    BlockView topRows(int n) { return BlockView(*this, 0, 0, n, cols()); }
    BlockView bottomRows(int n) { return BlockView(*this, rows() - n, 0, n, cols()); }
    BlockView leftCols(int n) { return BlockView(*this, 0, 0, rows(), n); }
    BlockView rightCols(int n) { return BlockView(*this, 0, cols() - n, rows(), n); }
    template <typename U>
    Matrix<U, R, C, Order> cast() const {
        Matrix<U, R, C, Order> out;
        if constexpr (kDyn) out.resize(rows(), cols());
        for (int64_t i = 0; i < size(); ++i) out.data()[i] = static_cast<U>(data()[i]);
        return out;
    }
    void setZero(int r, int c) {
        resize(r, c);
        setZero();
    }
    void setConstant(int r, int c, const T &v) {
        resize(r, c);
        std::fill(data(), data() + size(), v);
    }

    static Matrix Zero() {
        Matrix m;
        m.setZero();
        return m;
    }
    static Matrix Constant(const T &v) {
        Matrix m;
        std::fill(m.data(), m.data() + m.size(), v);
        return m;
    }
    static Matrix Identity() {
        Matrix m = Zero();
        for (int i = 0; i < std::min(m.rows(), m.cols()); ++i) m(i, i) = T(1);
        return m;
    }

    // `m << a, b, c, d;` fills in row-major reading order, as Eigen does.
    class CommaFill {
    public:
        CommaFill(Matrix &m, const T &first): m_(m) { Put(first); }
        CommaFill &operator,(const T &v) {
            Put(v);
            return *this;
        }
    private:
        void Put(const T &v) {
            m_(k_ / m_.cols(), k_ % m_.cols()) = v;
            ++k_;
        }
        Matrix &m_;
        int k_ = 0;
    };
    CommaFill operator<<(const T &v) { return CommaFill(*this, v); }

    Matrix &operator+=(const Matrix &o) {
        for (int64_t i = 0; i < size(); ++i) data()[i] += o.data()[i];
        return *this;
    }
    Matrix &operator/=(const T &k) {
        for (int64_t i = 0; i < size(); ++i) data()[i] /= k;
        return *this;
    }
    friend Matrix operator+(const Matrix &a, const Matrix &b) {
        Matrix out = a;
        out += b;
        return out;
    }
    friend Matrix operator*(const T &k, const Matrix &a) {
        Matrix out = a;
        for (int64_t i = 0; i < out.size(); ++i) out.data()[i] = k * a.data()[i];
        return out;
    }
    friend Matrix operator*(const Matrix &a, const T &k) { return k * a; }

private:
    Buf buf_;
    int rows_ = (R == Dynamic ? 0 : R);
    int cols_ = (C == Dynamic ? (R == Dynamic ? 0 : 1) : C);
};

// One row of a mapped matrix, as values: .cast<U>() and .transpose() keep the values, and the result converts to / assigns
// into a column vector of the same length (nn_feature_point_detector.cpp:214,225).
template <typename T>
class RowValues {
public:
    explicit RowValues(std::vector<T> v): v_(std::move(v)) {}
    template <typename U>
    RowValues<U> cast() const {
        std::vector<U> out(v_.size());
        for (size_t i = 0; i < v_.size(); ++i) out[i] = static_cast<U>(v_[i]);
        return RowValues<U>(std::move(out));
    }
    const RowValues &transpose() const { return *this; }
    template <int N, int O>
    operator Matrix<T, N, 1, O>() const {
        Matrix<T, N, 1, O> out;
        if constexpr (N == Dynamic) out.resize(int(v_.size()), 1);
        for (int i = 0; i < int(v_.size()) && i < int(out.size()); ++i) out[i] = v_[i];
        return out;
    }
private:
    std::vector<T> v_;
};

// Read-only view of a row-major buffer with the shape of PlainType.
template <typename PlainType>
class Map;
template <typename T, int R, int C, int Order>
class Map<const Matrix<T, R, C, Order>> {
    static_assert(Order == RowMajor, "only row-major maps are used (MatImgF, TMatImg<int64_t>)");
public:
    Map(const T *data, int rows, int cols): data_(data), rows_(rows), cols_(cols) {}
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    const T *data() const { return data_; }
    const T &operator()(int r, int c) const { return data_[int64_t(r) * cols_ + c]; }
    RowValues<T> row(int r) const { return RowValues<T>(std::vector<T>(data_ + int64_t(r) * cols_, data_ + int64_t(r + 1) * cols_)); }
private:
    const T *data_;
    int rows_, cols_;
};

// 2x2 * 2x1: coefficient i = m(i,0)*v0 + m(i,1)*v1 (Eigen's lazy coefficient product order).
template <typename T, int O1, int O2>
Matrix<T, 2, 1, O2> operator*(const Matrix<T, 2, 2, O1> &m, const Matrix<T, 2, 1, O2> &v) {
    Matrix<T, 2, 1, O2> out;
    out[0] = m(0, 0) * v[0] + m(0, 1) * v[1];
    out[1] = m(1, 0) * v[0] + m(1, 1) * v[1];
    return out;
}

}  // namespace Eigen

template <typename Scalar> using TVec2 = Eigen::Matrix<Scalar, 2, 1>;
template <typename Scalar> using TMatImg = Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
using Vec2 = Eigen::Matrix<float, 2, 1>;
using Vec4 = Eigen::Matrix<float, 4, 1>;
using Vec = Eigen::Matrix<float, Eigen::Dynamic, 1>;
using Mat2 = Eigen::Matrix<float, 2, 2>;
using MatInt = Eigen::Matrix<int32_t, Eigen::Dynamic, Eigen::Dynamic>;
using MatImg = TMatImg<uint8_t>;
using MatImgF = TMatImg<float>;
using Pixel = TVec2<int32_t>;  // (x = col, y = row), see feature_point_harris_detector.cpp:133

constexpr float kPai = 3.14159265358979323846f;
constexpr float k2Pai = 2.0f * kPai;
constexpr float kDegToRad = kPai / 180.0f;
constexpr float kZeroFloat = 1e-6f;  // GUESS G2; any value in (0, 1] behaves identically (fact B2)

#endif  // FD_COMPAT_BASIC_TYPE_H_
