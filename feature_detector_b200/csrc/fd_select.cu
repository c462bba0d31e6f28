// Kernel 3 -- greedy minimum-distance selection of the best candidates, per frame.
// Replaces FeaturePointDetector::SelectGoodFeatures + DrawRectangleInMask
// (reference src/feature_point_detector/feature_point_detector.cpp:54-74, 76-88).
//
// The reference sorts all candidates by response (descending) and walks them once: a candidate is kept
// iff its mask pixel is still set, and each kept candidate clears a (2d+1)^2 square of the mask.  That is
// exactly: "kept iff no higher-ranked KEPT candidate lies within Chebyshev distance d", and stopping after
// N features keeps the N best-ranked members of that set (a candidate's fate depends only on higher-ranked
// ones).  Rank = response descending, ties in raster order -- the tie rule this framework fixes where the
// reference's unstable std::sort leaves ties open.  One 64-bit key per candidate encodes the rank, smaller =
// better.
//
// select_kernel: one CTA per frame, no global sort.  Frame pixels are binned into a grid of (d+1)-sided
// cells; two kept points can never share a cell, and anything within d of a pixel lies in the 3x3 cells
// around it.  Rounds until no candidate is alive:
//   A  every live candidate within d of a point kept in the previous round dies; the others post their key
//      to their cell with a shared-memory atomicMin (two 32-bit phases: response word, then position word);
//   B  a live candidate that holds its cell's minimum and beats the minima of the 8 neighbouring cells has no
//      live better-ranked candidate within d, so the sequential walk would keep it: it is kept now, written
//      to the frame's kept list and to the cell grid.  All such candidates of a round are independent.
// The kept list is then sorted by key (a few hundred entries, bitonic in shared memory) and cut at
// max(needed - existing, 1) -- the reference tests the count AFTER each push (:67-68), so needed = 0 still
// yields one feature.  Work per round is proportional to the candidates still alive, and the first round kills
// most of them.  Pre-existing features need no handling here: their squares were already masked out of
// candidate generation with the same d (feature_point_detector.cpp:12-16); they only count toward `needed`.
//
// sort_kernel: bitonic sort of each frame's keys (shared memory when they fit).  Not on the detection path any
// more; used for the LSD seed order and when the caller asks for the sorted candidate list on the device.
#include "fd_kernels.cuh"

namespace fdb {

namespace {

constexpr uint32_t kEmptyCell = 0xFFFFFFFFu;
constexpr uint64_t kDeadKey = 0xFFFFFFFFFFFFFFFFull;

// Ascending bitonic network over keys[0, n) (n need not be a power of two): ascending-only comparators, so the
// tail beyond n acts as +inf padding.
__device__ void block_bitonic_sort(uint64_t *keys, uint32_t n) {
    if (n < 2) return;
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (uint32_t k = 2; k <= np2; k <<= 1) {
        const uint32_t hk = k >> 1;
        for (uint32_t i = threadIdx.x; i < np2 / 2; i += blockDim.x) {  // mirror stage
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

// Is (x, y) within Chebyshev distance d of a point stored in the 3x3 cells around (cx, cy)?
__device__ __forceinline__ bool near_kept(const uint32_t *cells, int cells_x, int cells_y, int cx, int cy, int x, int y, int d) {
    const int x0 = max(cx - 1, 0), x2 = min(cx + 1, cells_x - 1);
    const int y0 = max(cy - 1, 0) * cells_x, y1 = cy * cells_x, y2 = min(cy + 1, cells_y - 1) * cells_x;
    uint32_t q[9];
    q[0] = cells[y0 + x0]; q[1] = cells[y0 + cx]; q[2] = cells[y0 + x2];
    q[3] = cells[y1 + x0]; q[4] = cells[y1 + cx]; q[5] = cells[y1 + x2];
    q[6] = cells[y2 + x0]; q[7] = cells[y2 + cx]; q[8] = cells[y2 + x2];
    bool hit = false;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int qx = int(q[t] & 0xFFFFu), qy = int(q[t] >> 16);
        hit |= (q[t] != kEmptyCell) && (abs(qx - x) <= d) && (abs(qy - y) <= d);
    }
    return hit;
}

// Best (smallest) key posted to the 3x3 cells around (cx, cy), the centre cell excluded.
__device__ __forceinline__ uint64_t neighbour_min(const uint32_t *min_hi, const uint32_t *min_lo, int cells_x, int cells_y, int cx, int cy) {
    uint64_t best = kDeadKey;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            const int xx = cx + dx, yy = cy + dy;
            if (xx < 0 || yy < 0 || xx >= cells_x || yy >= cells_y) continue;
            const int c = yy * cells_x + xx;
            const uint64_t k = (uint64_t(min_hi[c]) << 32) | min_lo[c];
            best = min(best, k);
        }
    }
    return best;
}

__global__ void __launch_bounds__(SELECT_THREADS) select_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int frame = blockIdx.x;
    const int d = p.min_distance;
    const int n_cells = p.cells_x * p.cells_y;
    // shared (or, for very fine grids, global) per-cell state
    uint32_t *cells, *min_hi, *min_lo;
    if (p.cells_in_smem) {
        cells = reinterpret_cast<uint32_t *>(smem);
        min_hi = cells + n_cells;
        min_lo = min_hi + n_cells;
    } else {
        cells = p.cell_scratch + int64_t(frame) * n_cells * 3;
        min_hi = cells + n_cells;
        min_lo = min_hi + n_cells;
    }
    __shared__ uint32_t s_kept, s_alive;
    __shared__ uint64_t s_sort[SELECT_SORT_SMEM];

    const uint32_t count = p.cand_counts[frame];
    if (count > p.cand_capacity && threadIdx.x == 0) atomicExch(p.overflow_flag, 1u);
    const uint32_t n = min(count, p.cand_capacity);
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    uint8_t *alive = p.alive_scratch + int64_t(frame) * p.cand_capacity;
    uint64_t *kept = p.kept_keys + int64_t(frame) * p.kept_capacity;
    const uint32_t cell_magic = p.cell_magic;  // ceil(2^32 / (d+1)): exact quotient for coordinates < 65536
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;

    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
        cells[i] = kEmptyCell;
        min_hi[i] = 0xFFFFFFFFu;
        min_lo[i] = 0xFFFFFFFFu;
    }
    if (p.mask.bits != nullptr) {
        // a candidate on a masked-out pixel is never accepted (feature_point_detector.cpp:66).  Candidate generation
        // already honours the mask; this only matters where a zero response can pass a negative threshold.
        const uint32_t *mb = p.mask.bits + int64_t(frame) * p.rows * p.mask.words_per_row;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t xy = cand_key_xy(__ldg(keys + i));
            const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
            alive[i] = uint8_t((mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) alive[i] = 1;
    }
    if (threadIdx.x == 0) s_kept = 0u;
    __syncthreads();

    if (d < 0) {
        // no suppression at all (DrawRectangleInMask clears nothing for a negative distance): everything is kept
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            if (i < uint32_t(p.kept_capacity)) kept[i] = keys[i];
        if (threadIdx.x == 0) s_kept = min(n, uint32_t(p.kept_capacity));
        __syncthreads();
    } else {
        for (int round = 0;; ++round) {
            // ---- A: kill what the previous round's kept points cover; post to cells (response word) ----
            if (threadIdx.x == 0) s_alive = 0u;
            __syncthreads();
            uint32_t mine_alive = 0u;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                if (!alive[i]) continue;
                const uint64_t key = __ldg(keys + i);
                const uint32_t xy = cand_key_xy(key);
                const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
                const int cx = int(__umulhi(uint32_t(x), cell_magic)), cy = int(__umulhi(uint32_t(y), cell_magic));
                if (round > 0 && near_kept(cells, p.cells_x, p.cells_y, cx, cy, x, y, d)) {
                    alive[i] = 0;
                } else {
                    atomicMin(min_hi + cy * p.cells_x + cx, uint32_t(key >> 32));
                    ++mine_alive;
                }
            }
            if (mine_alive) atomicAdd(&s_alive, mine_alive);
            __syncthreads();
            if (s_alive == 0u) break;
            // ---- A2: among the candidates that share the cell's best response word, post the position word ----
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                if (!alive[i]) continue;
                const uint64_t key = __ldg(keys + i);
                const uint32_t xy = cand_key_xy(key);
                const int cx = int(__umulhi(xy & 0xFFFFu, cell_magic)), cy = int(__umulhi(xy >> 16, cell_magic));
                const int c = cy * p.cells_x + cx;
                if (min_hi[c] == uint32_t(key >> 32)) atomicMin(min_lo + c, xy);
            }
            __syncthreads();
            // ---- B: cell winners that also beat the 8 neighbouring cells are kept ----
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                if (!alive[i]) continue;
                const uint64_t key = __ldg(keys + i);
                const uint32_t xy = cand_key_xy(key);
                const int cx = int(__umulhi(xy & 0xFFFFu, cell_magic)), cy = int(__umulhi(xy >> 16, cell_magic));
                const int c = cy * p.cells_x + cx;
                if (min_hi[c] != uint32_t(key >> 32) || min_lo[c] != xy) continue;
                if (key < neighbour_min(min_hi, min_lo, p.cells_x, p.cells_y, cx, cy)) {
                    alive[i] = 0;
                    const uint32_t slot = atomicAdd(&s_kept, 1u);
                    if (slot < uint32_t(p.kept_capacity)) kept[slot] = key;
                    cells[c] = xy;  // read by the next round, after the barriers below
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
                min_hi[i] = 0xFFFFFFFFu;
                min_lo[i] = 0xFFFFFFFFu;
            }
            // (the barrier at the top of the next round orders these resets and the new `cells` entries)
        }
    }

    // ---- order the kept points by rank and cut at the number the reference would have pushed ----
    __syncthreads();
    const uint32_t n_kept = min(s_kept, uint32_t(p.kept_capacity));
    if (n_kept <= uint32_t(SELECT_SORT_SMEM)) {
        for (uint32_t i = threadIdx.x; i < n_kept; i += blockDim.x) s_sort[i] = kept[i];
        __syncthreads();
        block_bitonic_sort(s_sort, n_kept);
        kept = s_sort;
    } else {
        block_bitonic_sort(kept, n_kept);
    }
    __syncthreads();
    uint32_t want = (p.needed > n_pre) ? (p.needed - n_pre) : 1u;  // pushed, then tested: at least one
    want = min(want, uint32_t(p.kp_capacity));
    const uint32_t n_out = min(n_kept, want);
    float4 *kp_out = p.keypoints + int64_t(frame) * p.kp_capacity;
    for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
        const uint64_t key = kept[i];
        const uint32_t xy = cand_key_xy(key);
        kp_out[i] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);  // Vec2(pixel.x(), pixel.y()), :67
    }
    if (threadIdx.x == 0) p.kp_counts[frame] = int32_t(n_out);
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

size_t select_smem_bytes(const SelectArgs &a) { return a.cells_in_smem ? size_t(a.cells_x) * a.cells_y * 12 : 0; }

cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream) {
    const size_t smem = select_smem_bytes(args);
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    select_kernel<<<args.n_frames, SELECT_THREADS, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
