// Kernel 3 -- per-frame candidate sort and greedy minimum-distance selection.
// Replaces FeaturePointDetector::SelectGoodFeatures + DrawRectangleInMask
// (reference src/feature_point_detector/feature_point_detector.cpp:54-74, 76-88).
//
// The reference sorts all candidates by response (descending) and walks them once: a candidate is kept
// iff its mask pixel is still set, and each kept candidate clears a (2d+1)^2 square of the mask.  That is
// equivalent to: "kept iff no higher-ranked KEPT candidate lies within Chebyshev distance d".  Two kept
// points can therefore never share a cell of a grid with (d+1)-pixel cells, so the mask is replaced by a
// grid holding at most one kept point per cell and a candidate is tested against the 3x3 cells around it:
// O(1) per candidate, independent of d, and small enough for shared memory.
//
// One CTA per frame:
//   1. sort the frame's 64-bit keys ascending (= response descending, ties in raster order -- the rule
//      this framework fixes where the reference's unstable std::sort leaves ties open).  Bitonic network
//      with ascending-only comparators so that the tail beyond n acts as +inf padding; runs in shared
//      memory when the frame's candidates fit, else in place in global memory;
//   2. warp 0 walks the sorted keys 32 at a time: every lane tests its candidate against the grid in
//      parallel, then the survivors of the chunk are resolved in rank order with ballots / shuffles (an
//      accepted survivor kills later survivors of the same chunk within distance d).  The walk stops as
//      soon as existing + accepted >= needed, tested AFTER each push like the reference (:67-68), so
//      needed = 0 still yields one feature.
// Pre-existing features need no handling here: their squares were already masked out of candidate
// generation with the same d (feature_point_detector.cpp:12-16), they only count toward `needed`.
#include "fd_kernels.cuh"

namespace fdb {

namespace {

constexpr uint64_t kPadKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kEmptyCell = 0xFFFFFFFFu;

// Ascending bitonic network over keys[0, n) (n need not be a power of two).
__device__ void block_bitonic_sort(uint64_t *keys, uint32_t n) {
    if (n < 2) return;
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        // first stage of the merge: mirror partner, so every comparator sorts ascending
        for (uint32_t i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
            const uint32_t blk = i / (k >> 1), off = i % (k >> 1);
            const uint32_t lo = blk * k + off;
            const uint32_t hi = blk * k + (k - 1 - off);
            if (hi < n) {
                const uint64_t a = keys[lo], b = keys[hi];
                if (a > b) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
                const uint32_t lo = ((i / j) * (j << 1)) + (i % j);
                const uint32_t hi = lo + j;
                if (hi < n) {
                    const uint64_t a = keys[lo], b = keys[hi];
                    if (a > b) {
                        keys[lo] = b;
                        keys[hi] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SELECT_THREADS) select_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(smem);
    uint32_t *scells = reinterpret_cast<uint32_t *>(smem + size_t(p.smem_sort_capacity) * 8);

    const int frame = blockIdx.x;
    const uint32_t count = p.cand_counts[frame];
    if (count > p.cand_capacity && threadIdx.x == 0) atomicExch(p.overflow_flag, 1u);
    const uint32_t n = min(count, p.cand_capacity);
    uint64_t *gkeys = p.cand_keys + int64_t(frame) * p.cand_capacity;

    // ---- 1. sort ----
    const bool in_smem = n <= uint32_t(p.smem_sort_capacity);
    if (in_smem) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) skeys[i] = gkeys[i];
        __syncthreads();
        block_bitonic_sort(skeys, n);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) gkeys[i] = skeys[i];
    } else {
        __syncthreads();
        block_bitonic_sort(gkeys, n);
    }

    // ---- 2. greedy selection ----
    const int d = p.min_distance;
    const int cell = d + 1;
    const int n_cells = p.cells_x * p.cells_y;
    uint32_t *cells = p.cells_in_smem ? scells : (p.cell_scratch + int64_t(frame) * n_cells);
    if (d >= 0)
        for (int i = threadIdx.x; i < n_cells; i += blockDim.x) cells[i] = kEmptyCell;
    __syncthreads();
    if (threadIdx.x >= 32) return;

    const int lane = lane_id();
    const uint64_t *keys = in_smem ? skeys : gkeys;
    float4 *kp_out = p.keypoints + int64_t(frame) * p.kp_capacity;
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    uint32_t accepted = 0;
    bool done = (n == 0);
    for (uint32_t base = 0; base < n && !done; base += 32) {
        const uint32_t i = base + lane;
        bool live = i < n;
        int x = 0, y = 0;
        float resp = 0.0f;
        if (live) {
            const uint64_t key = keys[i];
            const uint32_t raster = cand_key_raster(key);
            y = int(raster / uint32_t(p.cols));
            x = int(raster - uint32_t(y) * uint32_t(p.cols));
            resp = cand_key_response(key);
            if (d >= 0) {
                const int cx = x / cell, cy = y / cell;
                for (int yy = max(cy - 1, 0); yy <= min(cy + 1, p.cells_y - 1) && live; ++yy) {
                    for (int xx = max(cx - 1, 0); xx <= min(cx + 1, p.cells_x - 1); ++xx) {
                        const uint32_t q = cells[yy * p.cells_x + xx];
                        if (q != kEmptyCell) {
                            const int qx = int(q & 0xFFFFu), qy = int(q >> 16);
                            if (abs(qx - x) <= d && abs(qy - y) <= d) live = false;
                        }
                    }
                }
            }
        }
        uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        while (live_mask != 0u) {
            const int leader = __ffs(live_mask) - 1;
            const int lx = __shfl_sync(0xffffffffu, x, leader);
            const int ly = __shfl_sync(0xffffffffu, y, leader);
            if (lane == leader) {
                if (accepted < uint32_t(p.kp_capacity)) kp_out[accepted] = make_float4(float(x), float(y), resp, 0.0f);  // Vec2(pixel.x(), pixel.y()), :67
                if (d >= 0) cells[(y / cell) * p.cells_x + (x / cell)] = uint32_t(x) | (uint32_t(y) << 16);               // DrawRectangleInMask, :69
                live = false;
            } else if (live && d >= 0 && abs(lx - x) <= d && abs(ly - y) <= d) {
                live = false;
            }
            ++accepted;
            if (n_pre + accepted >= p.needed || accepted >= uint32_t(p.kp_capacity)) {  // :68, tested after the push
                done = true;
                break;
            }
            live_mask = __ballot_sync(0xffffffffu, live);
        }
        __syncwarp();
    }
    if (lane == 0) p.kp_counts[frame] = int32_t(accepted);
}

}  // namespace

size_t select_smem_bytes(const SelectArgs &a) {
    return size_t(a.smem_sort_capacity) * 8 + (a.cells_in_smem ? size_t(a.cells_x) * a.cells_y * 4 : 0);
}

cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream) {
    const size_t smem = select_smem_bytes(args);
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    select_kernel<<<args.n_frames, SELECT_THREADS, smem, stream>>>(args);
    return cudaGetLastError();
}

// ---- generic segmented sort (used for the LSD seed order) ------------------------------------------
namespace {
__global__ void __launch_bounds__(SELECT_THREADS) segment_sort_kernel(uint64_t *keys, const uint32_t *counts, int64_t slot, uint32_t capacity,
                                                                       int smem_capacity) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(smem);
    const uint32_t n = min(counts[blockIdx.x], capacity);
    uint64_t *g = keys + int64_t(blockIdx.x) * slot;
    if (n <= uint32_t(smem_capacity)) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) skeys[i] = g[i];
        __syncthreads();
        block_bitonic_sort(skeys, n);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) g[i] = skeys[i];
    } else {
        block_bitonic_sort(g, n);
    }
}
}  // namespace

cudaError_t launch_segment_sort(uint64_t *keys, uint64_t *scratch, const uint32_t *counts, int64_t slot, int n_segments, uint32_t capacity,
                                cudaStream_t stream) {
    (void)scratch;
    int smem_cap = 16384;
    while (uint32_t(smem_cap) / 2 >= capacity && smem_cap > 1024) smem_cap >>= 1;
    const size_t smem = size_t(smem_cap) * 8;
    cudaError_t e = cudaFuncSetAttribute(segment_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    segment_sort_kernel<<<n_segments, SELECT_THREADS, smem, stream>>>(keys, counts, slot, capacity, smem_cap);
    return cudaGetLastError();
}

}  // namespace fdb
