// Post-processing of a learned detector's outputs -- the part of the reference's NNFeaturePointDetector
// that is not the ONNX session: heat map -> candidates -> greedy minimum-distance selection -> descriptor sampling
// (reference src/nn_feature_point_detector/nn_feature_point_detector.cpp:59-72, 128-155, 163-193).  The model itself
// stays wherever the caller runs it; its outputs are expected in device memory.
//
// nn_heatmap_kernel: SelectKeypointCandidatesFromHeatMap (:128-139) fused with the invalid-boundary part of CreateMask
// (:62-67).  A pixel is a candidate iff response > kMinResponse and it lies kInvalidBoundary pixels inside the map.  The
// reference keeps candidates in a std::multimap<float, Pixel> and walks it backwards, so among equal responses the LATER
// raster position wins; the candidate key therefore carries the complemented position (selection runs with xy_xor set).
// 4 B/px read, HBM-bound.
//
// nn_descriptor_kernel: ExtractDescriptorsForSelectedFeatures (:163-193).  One warp per keypoint slot, lanes across the
// channel planes; the four bilinear weights and the four products / three sums in the reference's order, unfused.
#include "fd_kernels.cuh"

namespace fdb {

namespace {

constexpr int NN_STEPS = 2;                 // 128-column steps per pass over a row: 256 columns, one float4 per lane per step
constexpr int NN_CHUNK_ROWS = 16;            // rows per work item
constexpr int NN_LANE_SLOTS = 16;           // candidate keys each lane can stage between flushes (a pass adds at most 4 * NN_STEPS)
constexpr int NN_STAGE = 32 * NN_LANE_SLOTS; // ... per warp: entry k of lane l sits at stage[32 k + l]

// One warp per run of consecutive valid rows (the invalid boundary rows are never read).  Per pass the warp issues all its
// float4 loads, reduces every element to one bit (above the threshold, inside the column bounds), and only then do the lanes
// that own candidates -- a few per cent of the pixels -- take stage slots (one shared-memory add per lane per pass) and
// write their keys.  The stage leaves for the frame's candidate slot in blocks: one reservation, one coalesced copy.
__global__ void __launch_bounds__(256) nn_heatmap_kernel(const NnHeatmapArgs p) {
    __shared__ uint64_t stage_all[8][NN_STAGE];
    uint64_t *stage = stage_all[threadIdx.x >> 5];
    const int lane = lane_id();
    const int b = p.invalid_boundary;
    const int valid_rows = p.rows - 2 * b;
    if (valid_rows <= 0 || p.cols - 2 * b <= 0) return;
    const int64_t total_rows = int64_t(p.n_frames) * valid_rows;
    const int64_t map_px = int64_t(p.rows) * p.cols;
    const bool vec = (p.cols % 4 == 0) && (reinterpret_cast<uintptr_t>(p.heatmap) % 16 == 0);
    const int col_lo = b, col_hi = p.cols - b;   // valid columns: col_lo <= col < col_hi
    int frame = 0, row = 0;
    // Candidates are staged per LANE (as in fd_corner_tma.cu): a lane that finds one stores it and bumps its own count; the warp-wide
    // scan and the reservation in the frame's slot happen once per flush.
    uint32_t my_staged = 0u;   // this lane's staged keys
    auto flush = [&]() {
        __syncwarp();
        uint32_t incl = my_staged;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total != 0u) {
            uint32_t g = 0u;
            if (lane == 0) g = atomicAdd(p.cand_counts + frame, total);
            g = __shfl_sync(0xffffffffu, g, 0) + incl - my_staged;
            uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;
            for (uint32_t k = 0; k < my_staged; ++k)
                if (g + k < p.cand_capacity) slot[g + k] = stage[32u * k + uint32_t(lane)];
        }
        __syncwarp();
        my_staged = 0u;
    };
    // runs of NN_CHUNK_ROWS consecutive valid rows are handed out by a global counter (zeroed by the host): the cost of a run
    // follows its candidate count, so a fixed partition leaves a tail
    for (;;) {
    uint32_t chunk = 0u;
    if (lane == 0) chunk = atomicAdd(p.work_counter, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    const int64_t first = int64_t(chunk) * NN_CHUNK_ROWS;
    if (first >= total_rows) break;
    const int64_t last = min(first + NN_CHUNK_ROWS, total_rows);
    frame = int(first / valid_rows);
    row = b + int(first - int64_t(frame) * valid_rows);
    for (int64_t it = first; it < last; ++it) {
        const float *rp = p.heatmap + int64_t(frame) * map_px + int64_t(row) * p.cols;
        for (int base = 0; base < p.cols; base += 128 * NN_STEPS) {
            float4 q[NN_STEPS];
#pragma unroll
            for (int k = 0; k < NN_STEPS; ++k) {
                const int c0 = base + 128 * k + 4 * lane;
                q[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (vec) {
                    if (c0 < p.cols) q[k] = __ldcs(reinterpret_cast<const float4 *>(rp + c0));
                } else {
                    if (c0 < p.cols) q[k].x = __ldcs(rp + c0);
                    if (c0 + 1 < p.cols) q[k].y = __ldcs(rp + c0 + 1);
                    if (c0 + 2 < p.cols) q[k].z = __ldcs(rp + c0 + 2);
                    if (c0 + 3 < p.cols) q[k].w = __ldcs(rp + c0 + 3);
                }
            }
            uint32_t mine = 0u;   // bit 4 k + j: element j of step k is a candidate
#pragma unroll
            for (int k = 0; k < NN_STEPS; ++k) {
                uint32_t m = (q[k].x > p.min_response ? 1u : 0u) | (q[k].y > p.min_response ? 2u : 0u) | (q[k].z > p.min_response ? 4u : 0u) |
                             (q[k].w > p.min_response ? 8u : 0u);                                   // .cpp:133
                const int c0 = base + 128 * k + 4 * lane;
                if (c0 < col_lo || c0 + 4 > col_hi) {   // the few words that straddle the invalid boundary (.cpp:62-67) or the row's end
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + j < col_lo || c0 + j >= col_hi) m &= ~(1u << j);
                }
                mine |= m << (4 * k);
            }
            if (__any_sync(0xffffffffu, mine != 0u)) {
                if (mine != 0u) {
                    do {
                        const int bit = __ffs(mine) - 1;
                        mine &= mine - 1u;
                        const int c = base + 128 * (bit >> 2) + 4 * lane + (bit & 3);
                        const int k = bit >> 2;   // pick the element out of the registers (a select tree; no second trip to memory)
                        static_assert(NN_STEPS == 2, "the select tree below is written for two steps");
                        const float4 w = k ? q[1] : q[0];
                        const float v = (bit & 2) ? ((bit & 1) ? w.w : w.z) : ((bit & 1) ? w.y : w.x);
                        stage[32u * my_staged + uint32_t(lane)] = (uint64_t(~float_to_ordered(v)) << 32) | uint32_t(~((uint32_t(row) << 16) | uint32_t(c)));
                        ++my_staged;
                    } while (mine != 0u);
                }
                if (__any_sync(0xffffffffu, my_staged > uint32_t(NN_LANE_SLOTS - 4 * NN_STEPS))) flush();
            }
        }
        if (++row == p.rows - b) {   // the run continues in the next frame
            flush();
            row = b;
            ++frame;
        }
    }
    flush();   // the stage never carries keys across runs (the next run may belong to another frame)
    }
}

__global__ void __launch_bounds__(256) nn_descriptor_kernel(const NnDescriptorArgs p) {
    const int lane = lane_id();
    const int64_t slot = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int frame = int(slot / p.kp_capacity);
    const int idx = int(slot % p.kp_capacity);
    if (frame >= p.n_frames || idx >= p.kp_counts[frame]) return;
    const float4 kp = p.keypoints[slot];
    const float row = __fdiv_rn(kp.y, 8.0f), col = __fdiv_rn(kp.x, 8.0f);                      // .cpp:169-170
    const int int_row = int(row), int_col = int(col);                                            // .cpp:171-172
    const float sub_row = __fsub_rn(row, floorf(row)), sub_col = __fsub_rn(col, floorf(col));    // .cpp:173-174
    const float inv_sub_row = __fsub_rn(1.0f, sub_row), inv_sub_col = __fsub_rn(1.0f, sub_col);  // .cpp:175-176
    const float w0 = __fmul_rn(inv_sub_col, inv_sub_row), w1 = __fmul_rn(sub_col, inv_sub_row);  // .cpp:177
    const float w2 = __fmul_rn(inv_sub_col, sub_row), w3 = __fmul_rn(sub_col, sub_row);
    const bool outside = int_row < 0 || int_row >= p.map_rows - 1 || int_col < 0 || int_col >= p.map_cols - 1;   // .cpp:183
    const int64_t plane = int64_t(p.map_rows) * p.map_cols;
    const float *base = p.maps + int64_t(frame) * p.channels * plane + int64_t(int_row) * p.map_cols + int_col;
    float *out = p.out + slot * p.channels;
    for (int c = lane; c < p.channels; c += 32) {
        float v = 0.0f;
        if (!outside) {
            const float *q = base + int64_t(c) * plane;
            v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, __ldg(q)), __fmul_rn(w1, __ldg(q + 1))), __fmul_rn(w2, __ldg(q + p.map_cols))),
                          __fmul_rn(w3, __ldg(q + p.map_cols + 1)));                             // .cpp:188-189
        }
        out[c] = v;
    }
}

}  // namespace

cudaError_t launch_nn_heatmap(const NnHeatmapArgs &args, int sm_count, cudaStream_t stream) {
    const int64_t rows = int64_t(args.n_frames) * std::max(args.rows - 2 * args.invalid_boundary, 0);
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>((rows + 127) / 128, int64_t(8) * sm_count)));   // 8 CTAs of 8 warps per SM, at most one 16-row run per warp
    nn_heatmap_kernel<<<grid, 256, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_nn_descriptors(const NnDescriptorArgs &args, cudaStream_t stream) {
    const int64_t slots = int64_t(args.n_frames) * args.kp_capacity;
    if (slots == 0) return cudaSuccess;
    nn_descriptor_kernel<<<unsigned((slots + 7) / 8), 256, 0, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
