// Replays the call sequences of the reference's demo programs through the drop-in classes and prints what they
// return as one JSON object (the reference's demos assert nothing and need a display; tests/test_dropin_cpp.py
// compares this output with the golden answers).
//   test/test_feature_point_detector.cpp:27-97   TestHarris / TestShiTomas / TestFast / pre-seeded mask
//   test/test_feature_descriptor.cpp:14-58       Harris (thr 20, d 20, N 10) -> BRIEF kLength 128, kHalfPatchSize 8
//   test/test_feature_line_detector.cpp:99-106   LSD (only the level-line field stage is on the GPU)
// usage: fd_dropin_check <raw u8 image> <rows> <cols>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "descriptor_brief.h"
#include "feature_line_detector.h"
#include "nn_feature_point_postprocess.h"
#include "feature_line_field.h"
#include "feature_point_fast_detector.h"
#include "feature_point_harris_detector.h"
#include "feature_point_shi_tomas_detector.h"

using namespace feature_detector;

namespace {

struct Fnv {  // 64-bit FNV-1a with the (truncated) basis the golden file uses
    uint64_t h = 1469598103934665603ull;
    void Bytes(const void *p, size_t n) {
        const uint8_t *b = static_cast<const uint8_t *>(p);
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    }
    void I32(int32_t v) { Bytes(&v, 4); }
    void F32(float v) { Bytes(&v, 4); }
};

std::string Hex(uint64_t h) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "%016llx", static_cast<unsigned long long>(h));
    return buf;
}

std::string FeatureHash(const std::vector<Vec2> &f) {
    Fnv h;
    for (const Vec2 &v : f) {
        h.I32(int32_t(v.x()));
        h.I32(int32_t(v.y()));
    }
    return Hex(h.h);
}

template <typename Detector>
void RunDetector(const char *label, Detector &det, const GrayImage &image, float thr, int32_t dist, uint32_t needed, std::vector<Vec2> &features,
                 bool first) {
    det.options().kMinFeatureDistance = dist;
    det.options().kMinValidResponse = thr;
    const size_t n_pre = features.size();
    const bool ok = det.DetectGoodFeatures(image, needed, features);
    int64_t mask_zeros = 0;
    const MatInt &mask = det.mask();
    for (int64_t i = 0; i < mask.size(); ++i) mask_zeros += (mask.data()[i] == 0);
    Fnv mh;   // row-major walk, whatever the matrix's storage order
    for (int32_t r = 0; r < mask.rows(); ++r)
        for (int32_t c = 0; c < mask.cols(); ++c) mh.I32(int32_t(mask(r, c)));
    std::printf("%s\"%s\": {\"ok\": %s, \"name\": \"%s\", \"n_pre\": %zu, \"n_feat\": %zu, \"feat_hash\": \"%s\", \"n_cand\": %zu, "
                "\"first\": [%d, %d], \"mask_zeros\": %lld, \"mask_rows\": %d, \"mask_cols\": %d, \"mask_hash\": \"%s\", \"top_response\": %.9g}",
                first ? "" : ",\n ", label, ok ? "true" : "false", det.DetectorTypeName().c_str(), n_pre, features.size(), FeatureHash(features).c_str(),
                det.candidates().size(), features.size() > n_pre ? int(features[n_pre].x()) : -1, features.size() > n_pre ? int(features[n_pre].y()) : -1,
                static_cast<long long>(mask_zeros), int(mask.rows()), int(mask.cols()), Hex(mh.h).c_str(),
                det.candidates().empty() ? 0.0 : double(det.candidates()[0].first));
}

struct PixelParam {  // field-for-field the reference's FeatureLineDetector::PixelParam (feature_line_detector.h:14-22)
    int32_t row = 0;
    int32_t col = 0;
    float line_level_angle = 0.0f;
    float gradient_norm = 0.0f;
    bool is_valid = false;
    bool is_used = false;
    bool is_occupied = false;
};

}  // namespace

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <raw u8 image> <rows> <cols>\n", argv[0]);
        return 2;
    }
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    std::vector<uint8_t> buf(size_t(rows) * cols);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(buf.data(), 1, buf.size(), f) != buf.size()) {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    std::fclose(f);
    GrayImage image(buf.data(), rows, cols, false);

    std::printf("{");
    {   // test_feature_point_detector.cpp: the four demo calls, N = 200
        std::vector<Vec2> features;
        FeaturePointFastDetector fast;
        RunDetector("fast_demo", fast, image, 10.0f, 20, 200, features, true);
        features.clear();
        FeaturePointHarrisDetector harris;
        RunDetector("harris_demo", harris, image, 30.0f, 20, 200, features, false);
        features.clear();
        FeaturePointShiTomasDetector shi;
        RunDetector("shi_demo", shi, image, 40.0f, 20, 200, features, false);
        features.clear();
        for (int i = 1; i < 10; ++i)
            for (int j = 1; j < 10; ++j) features.emplace_back(Vec2(15 * i, 15 * j));   // :52-56
        FeaturePointHarrisDetector harris2;
        RunDetector("harris_preseeded81", harris2, image, 30.0f, 20, 200, features, false);
        // defaults (feature_point_detector.h:16-19), same object reused: scratch must not leak between calls
        features.clear();
        RunDetector("harris_default_reused", harris2, image, 0.1f, 15, 200, features, false);
        features.clear();
        FeaturePointFastDetector fast9;
        fast9.sub_options().kN = 9;
        RunDetector("fast9_demo", fast9, image, 10.0f, 20, 200, features, false);

        // null image -> false, nothing touched (feature_point_detector.cpp:9)
        GrayImage null_image(nullptr, rows, cols, false);
        std::vector<Vec2> untouched;
        std::printf(",\n \"null_image_returns\": %s", harris.DetectGoodFeatures(null_image, 10, untouched) ? "true" : "false");
    }
    {   // test_feature_descriptor.cpp: Harris thr 20, d 20, N 10 -> BRIEF 128 / 8
        FeaturePointHarrisDetector detector;
        detector.options().kMinFeatureDistance = 20;
        detector.options().kMinValidResponse = 20.0f;
        std::vector<Vec2> features;
        detector.DetectGoodFeatures(image, 10, features);
        BriefDescriptor descriptor;
        descriptor.options().kLength = 128;
        descriptor.options().kHalfPatchSize = 8;
        std::vector<BriefType> desc;
        const bool ok = descriptor.Compute(image, features, desc);
        Fnv h;
        int ones = 0, all_zero = 0;
        for (const BriefType &d : desc) {
            int mine = 0;
            for (size_t b = 0; b < d.size(); b += 8) {   // pack LSB first, like the golden file
                uint8_t byte = 0;
                for (size_t k = 0; k < 8 && b + k < d.size(); ++k) byte |= uint8_t(d[b + k]) << k;
                h.Bytes(&byte, 1);
            }
            for (bool b : d) mine += b;
            ones += mine;
            all_zero += (mine == 0);
        }
        std::vector<Vec> as_float;
        descriptor.Compute(image, features, as_float);
        double float_sum = 0.0;
        for (const Vec &v : as_float)
            for (int i = 0; i < v.rows(); ++i) float_sum += v[i];
        std::vector<BriefType> none;
        std::printf(",\n \"brief_harris10\": {\"ok\": %s, \"n\": %zu, \"length\": %zu, \"ones\": %d, \"all_zero\": %d, \"hash\": \"%s\", \"float_sum\": %.1f, "
                    "\"empty_returns\": %s}",
                    ok ? "true" : "false", desc.size(), desc.empty() ? size_t(0) : desc[0].size(), ones, all_zero, Hex(h.h).c_str(), float_sum,
                    descriptor.Compute(image, std::vector<Vec2>(), none) ? "true" : "false");
    }
    {   // test_feature_line_detector.cpp: the dense stage, then the host stage's input as the reference lays it out
        LineLevelAngleField field;
        const bool ok = field.Compute(image);
        Eigen::Matrix<PixelParam, Eigen::Dynamic, Eigen::Dynamic> pixels;
        std::vector<PixelParam *> sorted;
        if (ok) field.FillPixelParams(pixels, sorted);
        Fnv hn, hs;
        int64_t n_valid = 0;
        double norm_sum = 0.0, angle_sum = 0.0;
        for (int r = 0; r < pixels.rows(); ++r)
            for (int c = 0; c < pixels.cols(); ++c) {   // row-major dump, like the golden file
                const PixelParam &p = pixels(r, c);
                hn.F32(p.gradient_norm);
                n_valid += p.is_valid;
                norm_sum += p.gradient_norm;
                if (p.is_valid) angle_sum += p.line_level_angle;
            }
        bool descending = true, positions_ok = true;
        for (size_t i = 0; i < sorted.size(); ++i) {
            hs.F32(sorted[i]->gradient_norm);
            if (i > 0 && sorted[i]->gradient_norm > sorted[i - 1]->gradient_norm) descending = false;
            positions_ok = positions_ok && (&pixels(sorted[i]->row, sorted[i]->col) == sorted[i]);
        }
        std::printf(",\n \"lsd_field\": {\"ok\": %s, \"rows\": %d, \"cols\": %d, \"n_valid\": %lld, \"n_sorted\": %zu, \"norm_hash\": \"%s\", \"norm_sum\": %.6f, "
                    "\"angle_sum\": %.6f, \"sorted_norm_hash\": \"%s\", \"descending\": %s, \"positions_ok\": %s}",
                    ok ? "true" : "false", pixels.rows(), pixels.cols(), static_cast<long long>(n_valid), sorted.size(), Hex(hn.h).c_str(), norm_sum, angle_sum,
                    Hex(hs.h).c_str(), descending ? "true" : "false", positions_ok ? "true" : "false");
    }
    {   // test_feature_line_detector.cpp:99-106: the whole detector, N = 200, default options
        FeatureLineDetector detector;
        std::vector<Vec4> lines;
        const bool ok = detector.DetectGoodFeatures(image, 200, lines);
        std::printf(",\n \"lsd_detect\": {\"ok\": %s, \"n_lines\": %zu, \"n_rectangles\": %zu, \"n_seeds\": %zu, \"lines\": [", ok ? "true" : "false",
                    lines.size(), detector.rectangles().size(), detector.sorted_pixels().size());
        for (size_t i = 0; i < lines.size(); ++i)
            std::printf("%s[%.9g, %.9g, %.9g, %.9g]", i ? ", " : "", double(lines[i][0]), double(lines[i][1]), double(lines[i][2]), double(lines[i][3]));
        std::printf("]}");
    }
    {   // DetectGoodFeaturesBatch (an addition: many frames per call): three frames, each with its own pre-existing features, must give what
        // three single-frame calls give -- Harris at the demo settings, and FAST at the reference's default threshold, whose ~343 000
        // candidates per frame overflow the batch's first, bounded candidate slots and make the call run again with full ones
        std::vector<uint8_t> frames(size_t(3) * rows * cols);
        for (int f = 0; f < 3; ++f)
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c < cols; ++c) frames[(size_t(f) * rows + r) * cols + c] = buf[size_t((r + 7 * f) % rows) * cols + (c + 13 * f) % cols];
        bool all_ok = true, all_equal = true;
        size_t n_total = 0;
        for (int which = 0; which < 2; ++which) {
            FeaturePointHarrisDetector harris;
            FeaturePointFastDetector fast;
            FeaturePointDetector &det = which == 0 ? static_cast<FeaturePointDetector &>(harris) : static_cast<FeaturePointDetector &>(fast);
            det.options().kMinFeatureDistance = which == 0 ? 20 : 15;
            det.options().kMinValidResponse = which == 0 ? 30.0f : 0.1f;
            std::vector<std::vector<Vec2>> batch(3);
            batch[1].emplace_back(Vec2(100.0f, 100.0f));
            batch[2].emplace_back(Vec2(300.0f, 200.0f));
            batch[2].emplace_back(Vec2(50.5f, 400.25f));
            std::vector<std::vector<Vec2>> single = batch;
            all_ok = det.DetectGoodFeaturesBatch(frames.data(), rows, cols, 3, 200, batch) && all_ok;
            for (int f = 0; f < 3; ++f) {
                GrayImage one(frames.data() + size_t(f) * rows * cols, rows, cols, false);
                all_ok = det.DetectGoodFeatures(one, 200, single[f]) && all_ok;
                all_equal = all_equal && single[f].size() == batch[f].size();
                for (size_t i = 0; all_equal && i < single[f].size(); ++i) all_equal = single[f][i].x() == batch[f][i].x() && single[f][i].y() == batch[f][i].y();
                n_total += batch[f].size();
            }
        }
        std::printf(",\n \"batch_of_three\": {\"ok\": %s, \"equals_single_frame_calls\": %s, \"n_features\": %zu}", all_ok ? "true" : "false",
                    all_equal ? "true" : "false", n_total);
    }
    if (argc >= 8) {   // NN post-processing: <heat map f32 file> <descriptor volume f32 file> <channels> <pre-existing features>
        const int channels = std::atoi(argv[6]), n_pre = std::atoi(argv[7]);
        auto read_floats = [](const char *path, size_t n, std::vector<float> &v) {
            v.resize(n);
            FILE *g = std::fopen(path, "rb");
            const bool ok = g != nullptr && std::fread(v.data(), 4, n, g) == n;
            if (g) std::fclose(g);
            return ok;
        };
        std::vector<float> heat, vol;
        const int mr = rows / 8, mc = cols / 8;
        if (!read_floats(argv[4], size_t(rows) * cols, heat) || !read_floats(argv[5], size_t(channels) * mr * mc, vol)) {
            std::fprintf(stderr, "cannot read the NN inputs\n");
            return 2;
        }
        NNFeaturePointPostProcessor post;
        std::vector<Vec2> features;
        for (int i = 0; i < n_pre; ++i) features.emplace_back(Vec2(float(20 + 31 * i % (cols - 40)), float(20 + 17 * i % (rows - 40))));
        const bool ok = post.SelectGoodFeaturesFromHeatMap(heat.data(), rows, cols, features);
        std::vector<Vec> desc;
        const bool ok2 = post.ExtractDescriptorsForSelectedFeatures(features, vol.data(), channels, mr, mc, desc);
        Fnv dh;
        for (const Vec &d : desc)
            for (int j = 0; j < int(d.size()); ++j) dh.F32(d(j));
        std::printf(",\n \"nn\": {\"ok\": %s, \"ok_desc\": %s, \"n_feat\": %zu, \"feat_hash\": \"%s\", \"n_desc\": %zu, \"desc_hash\": \"%s\", \"features\": [",
                    ok ? "true" : "false", ok2 ? "true" : "false", features.size(), FeatureHash(features).c_str(), desc.size(), Hex(dh.h).c_str());
        for (size_t i = 0; i < features.size(); ++i) std::printf("%s[%d, %d]", i ? ", " : "", int(features[i].x()), int(features[i].y()));
        std::printf("]}");
    }
    std::printf("}\n");
    return 0;
}
