// Stand-in for Slam_Utility's onnx_run_time.h and for the ONNX Runtime C++ API it wraps -- neither is in this image.
// It exists for ONE build: oracle/_ref compiling the reference's src/nn_feature_point_detector/nn_feature_point_detector.cpp
// unmodified, so that its post-processing functions (mask, heat-map candidates, greedy selection, descriptor sampling,
// .cpp:59-230) can serve as the checker for the NN post-processing kernels.  Nothing here can run a model: a Session is
// always empty, so Initialize() / InferenceSession() compile but do nothing useful.  Test infrastructure only.
#ifndef FD_COMPAT_ONNX_RUN_TIME_H_
#define FD_COMPAT_ONNX_RUN_TIME_H_

#include <cstddef>
#include <string>
#include <vector>

#include "basic_type.h"
#include "datatype_image.h"

enum OrtLoggingLevel { ORT_LOGGING_LEVEL_WARNING = 2 };
enum GraphOptimizationLevel { ORT_ENABLE_EXTENDED = 2 };
enum ExecutionMode { ORT_SEQUENTIAL = 0, ORT_PARALLEL = 1 };
enum OrtAllocatorType { OrtDeviceAllocator = 0 };
enum OrtMemType { OrtMemTypeDefault = 0 };

namespace Ort {
struct Exception {};
struct Env {
    Env(OrtLoggingLevel, const char *) {}
};
struct SessionOptions {
    void SetGraphOptimizationLevel(GraphOptimizationLevel) {}
    void SetExecutionMode(ExecutionMode) {}
};
struct MemoryInfo {
    MemoryInfo(std::nullptr_t) {}
    static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return MemoryInfo(nullptr); }
};
struct RunOptions {
    void SetRunLogVerbosityLevel(int) {}
};
struct Value {};
struct Session {
    Session(std::nullptr_t) {}
    Session(Env &, const char *, SessionOptions &) {}
    explicit operator bool() const { return false; }
    bool operator!() const { return true; }
    std::vector<Value> Run(RunOptions &, const char *const *, const Value *, size_t, const char *const *, size_t) { return {}; }
};
}  // namespace Ort

// Descriptor widths: SuperPoint 256 (stated in nn_feature_point_detector.cpp:179), DISK 128 (the published model's width; GUESS).
using SuperpointDescriptorType = Eigen::Matrix<float, 256, 1>;
using DiskDescriptorType = Eigen::Matrix<float, 128, 1>;

class OnnxRuntime {
public:
    struct MatrixTensor {
        Ort::Value value;
    };
    static void TryToEnableCuda(Ort::SessionOptions &) {}
    static void ReportInformationOfSession(Ort::Session &) {}
    static void GetSessionIO(Ort::Session &, std::vector<std::string> &, std::vector<std::string> &) {}
    static void ConvertImageToTensor(const GrayImage &, Ort::MemoryInfo &, MatrixTensor &) {}
    static void ConvertGrayImageToRgbTensor(const GrayImage &, Ort::MemoryInfo &, MatrixTensor &) {}
    template <typename MapType>
    static bool ConvertTensorToImageMatrice(const Ort::Value &, std::vector<MapType> &) { return false; }
};

namespace SlamOperation {
// Indices that sort `values` ascending (the callers walk the result backwards).
template <typename T>
void ArgSort(const T *values, int32_t n, std::vector<int32_t> &indices) {
    indices.resize(n);
    for (int32_t i = 0; i < n; ++i) indices[i] = i;
    std::stable_sort(indices.begin(), indices.end(), [values](int32_t a, int32_t b) { return values[a] < values[b]; });
}
}  // namespace SlamOperation

#endif  // FD_COMPAT_ONNX_RUN_TIME_H_
