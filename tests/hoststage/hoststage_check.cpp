// TEST INFRASTRUCTURE (CPU only, never shipped): runs the drop-in FeatureLineDetector's HOST stage
// (feature_detector_b200/cpp/line_segments_host.cpp: region growing, rectangle fit, validation) without a GPU.
// The dense stage it consumes normally comes from kernel 5 through LineLevelAngleField::Compute; here that one
// member function is replaced by a stand-in that takes the same maps from the CPU oracle (oracle/libfd_oracle.so,
// orc_lsd_map), so `-m "not gpu"` covers the host logic with the reference's own segments as the expectation.
// Usage: hoststage_check image.u8 rows cols needed [min_norm [host_libm_angles [seed_order.i32]]]  ->  one line per segment, four
// floats as hex.  seed_order.i32 (row, col pairs) replaces the oracle's seed order: the reference orders seeds with an unstable
// std::sort (feature_line_detector.cpp:92-94), so which of two equal-norm seeds grows first -- and with it the segments -- depends on
// the C++ library's sort; the port, like the GPU kernel, keeps ties in push order.  With the reference's own order fed in, the host
// stage has to return the reference's segments bit for bit.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define private public   // the stand-in below fills LineLevelAngleField's result members
#include "feature_line_detector.h"
#undef private

#include "../../oracle/fd_oracle.h"

static std::vector<int32_t> g_seed_override;   // (row, col) pairs, empty = keep the oracle's order

namespace feature_detector {

LineLevelAngleField::~LineLevelAngleField() {}
void *LineLevelAngleField::AllocHost(size_t bytes) { return std::malloc(bytes ? bytes : 1); }   // no GPU here: plain memory
void LineLevelAngleField::FreeHost(void *ptr) { std::free(ptr); }

bool LineLevelAngleField::Compute(const GrayImage &image) {
    if (image.data() == nullptr || image.rows() < 2 || image.cols() < 2) return false;
    rows_ = image.rows();
    cols_ = image.cols();
    const int32_t pr = rows_ - 1, pc = cols_ - 1;
    std::vector<float> norm(size_t(pr) * pc), angle(size_t(pr) * pc);
    std::vector<uint8_t> valid(size_t(pr) * pc);
    std::vector<int32_t> sorted_rc(size_t(pr) * pc * 2);
    int64_t n_sorted = 0;
    if (orc_lsd_map(image.data(), rows_, cols_, options_.kMinValidGradientNorm, norm.data(), angle.data(), valid.data(), sorted_rc.data(),
                    int64_t(pr) * pc, &n_sorted) != 1)
        return false;
    // the library's layout: rows x cols maps with a zero last row / column, seeds as row * cols + col
    const size_t px = size_t(rows_) * cols_;
    if (!norm_.reserve(px) || !angle_.reserve(px) || !seeds_.reserve(px)) return false;
    norm_.set_size(px);
    angle_.set_size(px);
    std::fill(norm_.data(), norm_.data() + px, 0.0f);
    std::fill(angle_.data(), angle_.data() + px, 0.0f);
    for (int32_t r = 0; r < pr; ++r)
        for (int32_t c = 0; c < pc; ++c) {
            norm_.data()[size_t(r) * cols_ + c] = norm[size_t(r) * pc + c];
            angle_.data()[size_t(r) * cols_ + c] = valid[size_t(r) * pc + c] ? angle[size_t(r) * pc + c] : 0.0f;
        }
    if (!g_seed_override.empty()) {
        if (int64_t(g_seed_override.size()) != 2 * n_sorted) return false;   // must be a permutation of the same seeds
        sorted_rc.assign(g_seed_override.begin(), g_seed_override.end());
    }
    seeds_.set_size(size_t(n_sorted));
    for (int64_t i = 0; i < n_sorted; ++i) seeds_.data()[size_t(i)] = sorted_rc[2 * i] * cols_ + sorted_rc[2 * i + 1];
    return true;
}

}  // namespace feature_detector

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    const uint32_t needed = uint32_t(std::atoi(argv[4]));
    std::vector<uint8_t> buf(size_t(rows) * cols);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3;
    std::fclose(f);
    if (argc > 7) {
        FILE *sf = std::fopen(argv[7], "rb");
        if (!sf) return 4;
        int32_t rc[2];
        while (std::fread(rc, 4, 2, sf) == 2) {
            g_seed_override.push_back(rc[0]);
            g_seed_override.push_back(rc[1]);
        }
        std::fclose(sf);
    }
    GrayImage image(buf.data(), rows, cols, false);
    feature_detector::FeatureLineDetector det;
    if (argc > 5) det.options().kMinValidGradientNorm = float(std::atof(argv[5]));
    if (argc > 6) det.options().kHostLibmAngles = std::atoi(argv[6]) != 0;
    std::vector<Vec4> lines;
    const bool ok = det.DetectGoodFeatures(image, needed, lines);
    std::printf("ok %d lines %zu rectangles %zu seeds %zu\n", ok ? 1 : 0, lines.size(), det.rectangles().size(), det.sorted_pixels().size());
    for (const Vec4 &l : lines) {
        uint32_t b[4];
        for (int k = 0; k < 4; ++k) {
            const float v = l(k);
            std::memcpy(&b[k], &v, 4);
        }
        std::printf("%08x %08x %08x %08x\n", b[0], b[1], b[2], b[3]);
    }
    return ok ? 0 : 1;
}
