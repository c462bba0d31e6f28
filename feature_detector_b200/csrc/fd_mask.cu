// Mask of pre-existing features (feature_point_detector.cpp:76-98) -- placeholder translation unit,
// filled in with the mask rasteriser and FAST prefix counts.
#include "fd_kernels.cuh"
namespace fdb {}
