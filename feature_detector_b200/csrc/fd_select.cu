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
// Two launches per batch:
//   sort_kernel    one CTA per frame sorts the frame's 64-bit keys ascending (= response descending, ties
//                  in raster order -- the rule this framework fixes where the reference's unstable
//                  std::sort leaves ties open).  Bitonic network with ascending-only comparators, so the
//                  tail beyond n acts as +inf padding; in shared memory when the frame's candidates fit,
//                  else in place in global memory (L2 resident).
//   greedy_kernel  one WARP per frame walks the sorted keys 32 at a time: every lane tests its candidate
//                  against the grid in parallel (nine independent cell reads), then the survivors of the
//                  chunk are resolved in rank order with ballots / shuffles (an accepted survivor kills later
//                  survivors of the same chunk within distance d).  The walk stops as soon as existing +
//                  accepted >= needed, tested AFTER each push like the reference (:67-68), so needed = 0
//                  still yields one feature.  The walk is a latency chain, so the parallelism is across
//                  frames: a whole batch is resident at once.
// Pre-existing features need no handling here: their squares were already masked out of candidate
// generation with the same d (feature_point_detector.cpp:12-16), they only count toward `needed`.
#include "fd_kernels.cuh"

namespace fdb {

namespace {

constexpr uint32_t kEmptyCell = 0xFFFFFFFFu;

// Ascending bitonic network over keys[0, n) (n need not be a power of two).
__device__ void block_bitonic_sort(uint64_t *keys, uint32_t n) {
    if (n < 2) return;
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        // first stage of the merge: mirror partner, so every comparator sorts ascending
        const uint32_t hk = k >> 1;
        for (uint32_t i = threadIdx.x; i < np2 / 2; i += blockDim.x) {
            const uint32_t off = i & (hk - 1);
            const uint32_t lo = ((i - off) << 1) + off;
            const uint32_t hi = ((i - off) << 1) + (k - 1 - off);
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
                const uint32_t off = i & (j - 1);
                const uint32_t lo = ((i - off) << 1) + off;
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

__global__ void __launch_bounds__(SORT_THREADS) sort_kernel(uint64_t *keys, const uint32_t *counts, int64_t slot, uint32_t capacity,
                                                            int smem_capacity, uint32_t *overflow_flag) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(smem);
    const uint32_t count = counts[blockIdx.x];
    if (overflow_flag != nullptr && count > capacity && threadIdx.x == 0) atomicExch(overflow_flag, 1u);
    const uint32_t n = min(count, capacity);
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

__global__ void __launch_bounds__(GREEDY_WARPS * 32) greedy_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const int frame = blockIdx.x * GREEDY_WARPS + wib;
    if (frame >= p.n_frames) return;

    const int d = p.min_distance;
    const int n_cells = p.cells_x * p.cells_y;
    uint32_t *cells = p.cells_in_smem ? reinterpret_cast<uint32_t *>(smem) + size_t(wib) * n_cells : (p.cell_scratch + int64_t(frame) * n_cells);
    if (d >= 0)
        for (int i = lane; i < n_cells; i += 32) cells[i] = kEmptyCell;
    __syncwarp();

    const uint32_t n = min(p.cand_counts[frame], p.cand_capacity);
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    float4 *kp_out = p.keypoints + int64_t(frame) * p.kp_capacity;
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    const uint32_t cell_magic = p.cell_magic;  // ceil(2^32 / (d+1)): exact quotient for coordinates < 65536
    uint32_t accepted = 0;
    bool done = (n == 0);
    uint64_t key_next = (uint32_t(lane) < n) ? __ldg(keys + lane) : 0ull;
    for (uint32_t base = 0; base < n && !done; base += 32) {
        const uint32_t i = base + lane;
        const uint64_t key = key_next;
        if (i + 32 < n) key_next = __ldg(keys + i + 32);  // prefetch the next chunk
        bool live = i < n;
        const uint32_t xy = cand_key_xy(key);
        const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
        int cx = 0, cy = 0;
        if (d >= 0) {
            cx = int(__umulhi(uint32_t(x), cell_magic));
            cy = int(__umulhi(uint32_t(y), cell_magic));
            // nine independent reads (clamped duplicates at the borders are harmless)
            const int x0 = max(cx - 1, 0), x2 = min(cx + 1, p.cells_x - 1);
            const int y0 = max(cy - 1, 0) * p.cells_x, y1 = cy * p.cells_x, y2 = min(cy + 1, p.cells_y - 1) * p.cells_x;
            uint32_t q[9];
            q[0] = cells[y0 + x0]; q[1] = cells[y0 + cx]; q[2] = cells[y0 + x2];
            q[3] = cells[y1 + x0]; q[4] = cells[y1 + cx]; q[5] = cells[y1 + x2];
            q[6] = cells[y2 + x0]; q[7] = cells[y2 + cx]; q[8] = cells[y2 + x2];
            bool hit = false;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int qx = int(q[t] & 0xFFFFu), qy = int(q[t] >> 16);
                // an empty cell decodes to (65535, 65535): never within d of a real pixel unless d is absurd, so test it explicitly
                hit |= (q[t] != kEmptyCell) && (abs(qx - x) <= d) && (abs(qy - y) <= d);
            }
            live = live && !hit;
        }
        uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        while (live_mask != 0u) {
            const int leader = __ffs(live_mask) - 1;
            const int lx = __shfl_sync(0xffffffffu, x, leader);
            const int ly = __shfl_sync(0xffffffffu, y, leader);
            if (lane == leader) {
                if (accepted < uint32_t(p.kp_capacity)) kp_out[accepted] = make_float4(float(x), float(y), cand_key_response(key), 0.0f);  // :67
                if (d >= 0) cells[cy * p.cells_x + cx] = xy;                                                                            // :69
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

cudaError_t launch_segment_sort(uint64_t *keys, const uint32_t *counts, int64_t slot, int n_segments, uint32_t capacity, uint32_t *overflow_flag,
                                cudaStream_t stream) {
    int smem_cap = 1024;
    while (uint32_t(smem_cap) < capacity && smem_cap < SORT_SMEM_MAX_KEYS) smem_cap <<= 1;
    const size_t smem = size_t(smem_cap) * 8;
    cudaError_t e = cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    sort_kernel<<<n_segments, SORT_THREADS, smem, stream>>>(keys, counts, slot, capacity, smem_cap, overflow_flag);
    return cudaGetLastError();
}

size_t greedy_smem_bytes(const SelectArgs &a) { return a.cells_in_smem ? size_t(GREEDY_WARPS) * a.cells_x * a.cells_y * 4 : 0; }

cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream) {
    cudaError_t e = launch_segment_sort(args.cand_keys, args.cand_counts, int64_t(args.cand_capacity), args.n_frames, args.cand_capacity,
                                        args.overflow_flag, stream);
    if (e != cudaSuccess) return e;
    const size_t smem = greedy_smem_bytes(args);
    e = cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    greedy_kernel<<<(args.n_frames + GREEDY_WARPS - 1) / GREEDY_WARPS, GREEDY_WARPS * 32, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
