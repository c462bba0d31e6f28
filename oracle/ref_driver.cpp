// TEST INFRASTRUCTURE — not product code.
//
// C entry points around the UNMODIFIED reference classes, compiled in place from
// /root/reference/src/**.cpp by oracle/Makefile into oracle/_ref/libfd_ref.so.
// Nothing from the reference is copied: this file only includes its headers and calls its
// public (and, through the usual test-only `#define private public`, its private) members so
// that dense intermediate maps can be dumped.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load the resulting library.
//
// Build flags follow the reference's own CMakeLists.txt:6 (-std=c++17 -O3 -g -Wall -pthread,
// no -march => no FMA contraction), plus -fPIC -shared.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "basic_type.h"
#include "circular_buffer.h"
#include "datatype_image.h"
#include "slam_basic_math.h"
#include "slam_log_reporter.h"
#include "slam_operations.h"

#define private public
#define protected public
#include "descriptor_brief.h"
#include "feature_line_detector.h"
#include "feature_point_detector.h"
#include "feature_point_fast_detector.h"
#include "feature_point_harris_detector.h"
#include "feature_point_shi_tomas_detector.h"
#include "nn_feature_point_detector.h"
#undef private
#undef protected

using namespace feature_detector;

namespace {

enum Kind { kHarris = 0, kShiTomas = 1, kFast = 2 };

std::unique_ptr<FeaturePointDetector> MakeDetector(int kind, float min_response, int min_distance, int fast_n) {
    std::unique_ptr<FeaturePointDetector> det;
    if (kind == kHarris) {
        det.reset(new FeaturePointHarrisDetector());
    } else if (kind == kShiTomas) {
        det.reset(new FeaturePointShiTomasDetector());
    } else {
        auto *fast = new FeaturePointFastDetector();
        if (fast_n > 0) fast->sub_options_.kN = fast_n;  // private member, reachable only here
        det.reset(fast);
    }
    det->options().kMinValidResponse = min_response;
    det->options().kMinFeatureDistance = min_distance;
    return det;
}

double NowSec() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// Runs FeaturePointDetector::DetectGoodFeatures (feature_point_detector.cpp:7-25) once.
//  feats_xy     in/out, (x, y) float pairs; the first n_feats_in are the pre-existing features.
//  cand_*       the detector's candidates_ as the reference leaves them (sorted by std::sort).
//  response_map Harris / Shi-Tomasi responses_ (rows*cols, row-major) or NULL.
//  mask_out     mask_ after the call as int32 rows*cols ROW-major, or NULL.
// Returns 1/0 = the reference's bool, negative on a buffer that is too small.
int ref_detect(int kind, const uint8_t *img, int rows, int cols, float min_response, int min_distance, uint32_t needed,
               int fast_n, float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, float *cand_resp,
               int32_t *cand_xy, int64_t max_cand, int64_t *n_cand, float *response_map, int32_t *mask_out) {
    auto det = MakeDetector(kind, min_response, min_distance, fast_n);
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    std::vector<Vec2> features;
    for (int i = 0; i < n_feats_in; ++i) features.emplace_back(Vec2(feats_xy[2 * i], feats_xy[2 * i + 1]));
    const bool ok = det->DetectGoodFeatures(image, needed, features);
    if (static_cast<int>(features.size()) > max_feats) return -1;
    for (size_t i = 0; i < features.size(); ++i) {
        feats_xy[2 * i] = features[i].x();
        feats_xy[2 * i + 1] = features[i].y();
    }
    *n_feats_out = static_cast<int>(features.size());
    const auto &cands = det->candidates();
    if (n_cand != nullptr) *n_cand = static_cast<int64_t>(cands.size());
    if (cand_resp != nullptr && cand_xy != nullptr) {
        if (static_cast<int64_t>(cands.size()) > max_cand) return -2;
        for (size_t i = 0; i < cands.size(); ++i) {
            cand_resp[i] = cands[i].first;
            cand_xy[2 * i] = cands[i].second.x();
            cand_xy[2 * i + 1] = cands[i].second.y();
        }
    }
    if (response_map != nullptr) {
        const MatImgF *resp = nullptr;
        if (kind == kHarris) resp = &static_cast<FeaturePointHarrisDetector *>(det.get())->responses_;
        if (kind == kShiTomas) resp = &static_cast<FeaturePointShiTomasDetector *>(det.get())->responses_;
        if (resp != nullptr && resp->rows() == rows && resp->cols() == cols) {
            std::memcpy(response_map, resp->data(), sizeof(float) * size_t(rows) * cols);
        } else {
            std::memset(response_map, 0, sizeof(float) * size_t(rows) * cols);
        }
    }
    if (mask_out != nullptr) {
        const MatInt &m = det->mask();
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) mask_out[size_t(r) * cols + c] = m(r, c);
    }
    return ok ? 1 : 0;
}

// Raw FAST score of every interior pixel via FeaturePointFastDetector::ComputeResponseOfPixel
// (feature_point_fast_detector.cpp:11-81); border pixels (3 px) are written as 0.
void ref_fast_score_map(const uint8_t *img, int rows, int cols, int fast_n, int diff, uint8_t *score_out) {
    FeaturePointFastDetector det;
    if (fast_n > 0) det.sub_options_.kN = fast_n;
    if (diff >= 0) det.sub_options_.kMinPixelDiffValue = static_cast<uint8_t>(diff);
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    std::memset(score_out, 0, size_t(rows) * cols);
    for (int r = 3; r < rows - 3; ++r)
        for (int c = 3; c < cols - 3; ++c) score_out[size_t(r) * cols + c] = static_cast<uint8_t>(det.ComputeResponseOfPixel(image, r, c));
}

// Descriptor<BriefType>::Compute (descriptor.h:28-40) -> one byte (0/1) per bit, n * length bytes.
int ref_brief(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, uint8_t *bits_out) {
    BriefDescriptor desc;
    desc.options().kLength = length;
    desc.options().kHalfPatchSize = half_patch;
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    std::vector<Vec2> uv;
    for (int i = 0; i < n; ++i) uv.emplace_back(Vec2(kp_xy[2 * i], kp_xy[2 * i + 1]));
    std::vector<BriefType> out;
    const bool ok = desc.Compute(image, uv, out);
    if (!ok) return 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < length; ++j) bits_out[size_t(i) * length + j] = out[i][j] ? 1 : 0;
    return 1;
}

// Descriptor<BriefType>::Compute, std::vector<Vec> overload (descriptor.h:43-62) -> n * length floats (+1 / -1).
int ref_brief_vec(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, float *out) {
    BriefDescriptor desc;
    desc.options().kLength = length;
    desc.options().kHalfPatchSize = half_patch;
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    std::vector<Vec2> uv;
    for (int i = 0; i < n; ++i) uv.emplace_back(Vec2(kp_xy[2 * i], kp_xy[2 * i + 1]));
    std::vector<Vec> vecs;
    if (!desc.Compute(image, uv, vecs)) return 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < length; ++j) out[size_t(i) * length + j] = vecs[i](j);
    return 1;
}

// The 256x4 pattern table (descriptor_brief.cpp:52-309) as the reference holds it in memory.
void ref_brief_pattern(int16_t *out_1024) {
    for (int i = 0; i < 1024; ++i) out_1024[i] = BriefDescriptor::pattern_idx_[i];
}

// FeatureLineDetector::ComputeLineLevelAngleMap (feature_line_detector.cpp:56-97).
// norm / angle / valid are (rows-1) x (cols-1) ROW-major dumps of pixels_ (angle 0 where invalid);
// sorted_rc holds (row, col) of sorted_pixels_ in the reference's std::sort order.
int ref_lsd_map(const uint8_t *img, int rows, int cols, float min_norm, float *norm, float *angle, uint8_t *valid,
                int32_t *sorted_rc, int64_t max_sorted, int64_t *n_sorted) {
    FeatureLineDetector det;
    det.options().kMinValidGradientNorm = min_norm;
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    if (!det.ComputeLineLevelAngleMap(image)) return 0;
    const auto &px = det.pixels();
    for (int r = 0; r < rows - 1; ++r) {
        for (int c = 0; c < cols - 1; ++c) {
            const auto &p = px(r, c);
            const size_t i = size_t(r) * (cols - 1) + c;
            norm[i] = p.gradient_norm;
            angle[i] = p.is_valid ? p.line_level_angle : 0.0f;
            valid[i] = p.is_valid ? 1 : 0;
        }
    }
    const auto &sp = det.sorted_pixels();
    *n_sorted = static_cast<int64_t>(sp.size());
    if (sorted_rc != nullptr) {
        if (static_cast<int64_t>(sp.size()) > max_sorted) return -1;
        for (size_t i = 0; i < sp.size(); ++i) {
            sorted_rc[2 * i] = sp[i]->row;
            sorted_rc[2 * i + 1] = sp[i]->col;
        }
    }
    return 1;
}

// Full FeatureLineDetector::DetectGoodFeatures (feature_line_detector.cpp:12-54); lines as x0,y0,x1,y1.
int ref_lsd_detect(const uint8_t *img, int rows, int cols, uint32_t needed, float min_norm, float *lines, int max_lines, int *n_lines) {
    FeatureLineDetector det;
    det.options().kMinValidGradientNorm = min_norm;
    GrayImage image(const_cast<uint8_t *>(img), rows, cols, false);
    std::vector<Vec4> out;
    const bool ok = det.DetectGoodFeatures(image, needed, out);
    *n_lines = static_cast<int>(out.size());
    if (static_cast<int>(out.size()) > max_lines) return -1;
    for (size_t i = 0; i < out.size(); ++i)
        for (int k = 0; k < 4; ++k) lines[4 * i + k] = out[i][k];
    return ok ? 1 : 0;
}

// FeaturePointDetector::SparsifyFeatures (feature_point_detector.cpp:27-52).
void ref_sparsify(const float *feats_xy, int n, int rows, int cols, int grid_rows, int grid_cols, uint8_t need, uint8_t after,
                  uint8_t *status, int n_status) {
    FeaturePointHarrisDetector det;
    det.options().kGridFilterRowDivideNumber = grid_rows;
    det.options().kGridFilterColDivideNumber = grid_cols;
    std::vector<Vec2> f;
    for (int i = 0; i < n; ++i) f.emplace_back(Vec2(feats_xy[2 * i], feats_xy[2 * i + 1]));
    std::vector<uint8_t> st(status, status + n_status);
    det.SparsifyFeatures(f, rows, cols, need, after, st);
    for (int i = 0; i < n; ++i) status[i] = st[i];
}

// CPU baseline: DetectGoodFeatures (+ optional BRIEF) over a batch of frames, one detector object per
// thread (frames are independent; the reference itself never threads).  Returns wall seconds.
//  brief_length <= 0 skips the descriptor stage.  totals[0] = keypoints, totals[1] = candidates,
//  totals[2] = descriptor bits set (so the work cannot be optimised away).
double ref_bench_points(int kind, const uint8_t *frames, int n_frames, int rows, int cols, float min_response, int min_distance,
                        uint32_t needed, int fast_n, int brief_length, int brief_half_patch, int n_threads, int64_t *totals) {
    std::atomic<int> next(0);
    std::atomic<int64_t> kp(0), cand(0), ones(0);
    auto work = [&]() {
        auto det = MakeDetector(kind, min_response, min_distance, fast_n);
        BriefDescriptor desc;
        if (brief_length > 0) {
            desc.options().kLength = brief_length;
            desc.options().kHalfPatchSize = brief_half_patch;
        }
        std::vector<Vec2> features;
        std::vector<BriefType> descriptors;
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_frames) break;
            GrayImage image(const_cast<uint8_t *>(frames) + size_t(f) * rows * cols, rows, cols, false);
            features.clear();
            det->DetectGoodFeatures(image, needed, features);
            kp += static_cast<int64_t>(features.size());
            cand += static_cast<int64_t>(det->candidates().size());
            if (brief_length > 0 && !features.empty()) {
                desc.Compute(image, features, descriptors);
                int64_t s = 0;
                for (const auto &d : descriptors)
                    for (size_t j = 0; j < d.size(); ++j) s += d[j] ? 1 : 0;
                ones += s;
            }
        }
    };
    const double t0 = NowSec();
    if (n_threads <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
        for (auto &th : pool) th.join();
    }
    const double t1 = NowSec();
    if (totals != nullptr) {
        totals[0] = kp.load();
        totals[1] = cand.load();
        totals[2] = ones.load();
    }
    return t1 - t0;
}

// CPU baseline for the LSD map stage (ComputeLineLevelAngleMap only) or the full detector.
double ref_bench_lsd(const uint8_t *frames, int n_frames, int rows, int cols, float min_norm, int full_detect, int n_threads, int64_t *totals) {
    std::atomic<int> next(0);
    std::atomic<int64_t> valid(0), lines(0);
    auto work = [&]() {
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_frames) break;
            FeatureLineDetector det;  // fresh object per frame: sorted_pixels_ is never cleared otherwise
            det.options().kMinValidGradientNorm = min_norm;
            GrayImage image(const_cast<uint8_t *>(frames) + size_t(f) * rows * cols, rows, cols, false);
            if (full_detect) {
                std::vector<Vec4> out;
                det.DetectGoodFeatures(image, 200, out);
                lines += static_cast<int64_t>(out.size());
            } else {
                det.ComputeLineLevelAngleMap(image);
            }
            valid += static_cast<int64_t>(det.sorted_pixels().size());
        }
    };
    const double t0 = NowSec();
    if (n_threads <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
        for (auto &th : pool) th.join();
    }
    const double t1 = NowSec();
    if (totals != nullptr) {
        totals[0] = valid.load();
        totals[1] = lines.load();
    }
    return t1 - t0;
}

// NN post-processing: NNFeaturePointDetector::CreateMask + SelectKeypointCandidatesFromHeatMap + SelectGoodFeaturesFromCandidates
// (nn_feature_point_detector.cpp:59-72, 128-155) on a heat map, without any model behind it (the ONNX session is a stand-in).
//  feats_xy  in/out (x, y) float pairs; the first n_feats_in are pre-existing features.  Returns 1, or 0 on failure.
int ref_nn_select(const float *heatmap, int rows, int cols, float min_response, int invalid_boundary, int min_distance, int max_features,
                  float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, int64_t *n_candidates) {
    NNFeaturePointDetector det;
    det.options().kMinResponse = min_response;
    det.options().kInvalidBoundary = invalid_boundary;
    det.options().kMinFeatureDistance = min_distance;
    det.options().kMaxNumberOfDetectedFeatures = max_features;
    std::vector<Vec2> features;
    for (int i = 0; i < n_feats_in; ++i) features.emplace_back(Vec2(feats_xy[2 * i], feats_xy[2 * i + 1]));
    const GrayImage shape_only(nullptr, rows, cols, false);   // CreateMask reads rows() and cols() only
    if (!det.CreateMask(shape_only, features)) return 0;
    MatImgF map;
    map.resize(rows, cols);
    std::memcpy(map.data(), heatmap, sizeof(float) * size_t(rows) * cols);
    if (!det.SelectKeypointCandidatesFromHeatMap(map)) return 0;
    if (n_candidates != nullptr) *n_candidates = int64_t(det.candidates_.size());
    if (!det.SelectGoodFeaturesFromCandidates(features)) return 0;
    const int n = std::min<int>(int(features.size()), max_feats);
    for (int i = 0; i < n; ++i) {
        feats_xy[2 * i] = features[i].x();
        feats_xy[2 * i + 1] = features[i].y();
    }
    *n_feats_out = n;
    return 1;
}

// NNFeaturePointDetector::ExtractDescriptorsForSelectedFeatures (nn_feature_point_detector.cpp:163-193): `maps` holds
// `channels` (256 = SuperPoint, 128 = DISK) row-major map_rows x map_cols planes; out is n_feats x channels.
int ref_nn_descriptors(const float *feats_xy, int n_feats, const float *maps, int channels, int map_rows, int map_cols, float *out) {
    NNFeaturePointDetector det;
    std::vector<Vec2> features;
    for (int i = 0; i < n_feats; ++i) features.emplace_back(Vec2(feats_xy[2 * i], feats_xy[2 * i + 1]));
    std::vector<Eigen::Map<const MatImgF>> planes;
    for (int c = 0; c < channels; ++c) planes.emplace_back(maps + size_t(c) * map_rows * map_cols, map_rows, map_cols);
    auto run = [&](auto tag) {
        using D = decltype(tag);
        std::vector<D> desc;
        if (!det.ExtractDescriptorsForSelectedFeatures(features, planes, desc)) return 0;
        for (int i = 0; i < n_feats; ++i)
            for (int c = 0; c < channels; ++c) out[size_t(i) * channels + c] = desc[i](c);
        return 1;
    };
    if (channels == 256) return run(SuperpointDescriptorType());
    if (channels == 128) return run(DiskDescriptorType());
    return 0;
}

const char *ref_build_info() { return "reference .cpp compiled in place; g++ -std=c++17 -O3 -g -pthread (CMakeLists.txt:6) + -fPIC -shared"; }

}  // extern "C"
