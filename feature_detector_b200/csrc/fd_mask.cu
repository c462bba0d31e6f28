// Mask of pre-existing features -- FeaturePointDetector::UpdateMaskByFeatures + DrawRectangleInMask
// (reference src/feature_point_detector/feature_point_detector.cpp:76-98).
//
// The reference keeps an int32 per pixel (1.4 MB at 752x480) and, when DetectGoodFeatures is handed a non-empty
// `features` vector, clears the clipped (2d+1)^2 square around every existing feature (coordinates truncated
// float -> int, :94-95) before candidates are computed.  Here the mask is one BIT per pixel (45 KB per frame),
// rasterised on the device from the feature list, one CTA per frame:
//   1. every word = all ones;
//   2. one warp per feature clears its square with atomicAnd on the words it touches;
//   3. (FAST only) the running offset of fast.cpp:85-93 advances once per MASKED-IN interior pixel, so the
//      kernel needs k(r, c) = number of masked-in interior pixels before (r, c) in raster order:
//      word_prefix[r][w] = masked-in interior pixels of row r in words < w, row_base[r] = those in rows < r
//      (row_base[rows-3] = total).  k = row_base[r] + word_prefix[r][w] + popc(bits below c in word w).
#include "fd_kernels.cuh"

namespace fdb {

namespace {

__global__ void __launch_bounds__(256) mask_kernel(const MaskArgs p) {
    const int frame = blockIdx.x;
    const int wpr = p.words_per_row;
    uint32_t *bits = p.bits + int64_t(frame) * p.rows * wpr;
    const int64_t n_words = int64_t(p.rows) * wpr;
    for (int64_t i = threadIdx.x; i < n_words; i += blockDim.x) bits[i] = 0xFFFFFFFFu;
    __syncthreads();

    const int n = min(p.counts[frame], p.capacity);
    const float *xy = p.xy + int64_t(frame) * p.capacity * 2;
    const int d = p.min_distance;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int i = warp; i < n; i += n_warps) {
        const float fx = xy[2 * i], fy = xy[2 * i + 1];
        // `const int32_t row = feature.y()` (:94-95): truncation toward zero.  Anything that cannot be represented
        // (the reference's behaviour is undefined there) is treated as far outside the frame.
        if (!(fabsf(fx) < 1.0e9f) || !(fabsf(fy) < 1.0e9f)) continue;
        const int64_t row = int64_t(int(fy)), col = int64_t(int(fx));
        const int64_t r0 = max(row - d, int64_t(0)), r1 = min(row + d, int64_t(p.rows - 1));   // :77-86
        const int64_t c0 = max(col - d, int64_t(0)), c1 = min(col + d, int64_t(p.cols - 1));
        if (r0 > r1 || c0 > c1) continue;
        const int w0 = int(c0 >> 5), w1 = int(c1 >> 5), nw = w1 - w0 + 1;
        const int64_t cells = (r1 - r0 + 1) * nw;
        for (int64_t t = lane; t < cells; t += 32) {
            const int r = int(r0 + t / nw), w = w0 + int(t % nw);
            const int lo = max(int(c0) - 32 * w, 0), hi = min(int(c1) - 32 * w, 31);
            const uint32_t upto_hi = (hi == 31) ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u);
            atomicAnd(bits + int64_t(r) * wpr + w, ~(upto_hi & ~((1u << lo) - 1u)));
        }
    }
    __syncthreads();

    if (p.word_prefix == nullptr) return;
    // ---- FAST prefix counts over the interior [3, rows-4] x [3, cols-4] ----
    uint32_t *prefix = p.word_prefix + int64_t(frame) * p.rows * wpr;
    uint32_t *row_base = p.row_base + int64_t(frame) * (p.rows + 1);
    for (int r = threadIdx.x; r < p.rows; r += blockDim.x) {
        uint32_t run = 0u;
        const bool interior_row = (r >= 3 && r <= p.rows - 4);
        for (int w = 0; w < wpr; ++w) {
            prefix[int64_t(r) * wpr + w] = run;
            if (interior_row) run += __popc(bits[int64_t(r) * wpr + w] & fast_interior_bits(w, p.cols));
        }
        row_base[r + 1] = run;  // row count for now; scanned below
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the row counts, 32 rows per step
        uint32_t carry = 0u;
        for (int base = 0; base < p.rows; base += 32) {
            const int r = base + lane;
            const uint32_t v = (r < p.rows) ? row_base[r + 1] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += u;
            }
            __syncwarp();
            if (r < p.rows) row_base[r + 1] = carry + incl;  // inclusive at r  ==  exclusive at r+1
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) row_base[0] = 0u;
    }
}

}  // namespace

cudaError_t launch_mask(const MaskArgs &args, cudaStream_t stream) {
    mask_kernel<<<args.n_frames, 256, 0, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
