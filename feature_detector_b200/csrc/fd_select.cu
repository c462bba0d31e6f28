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

// The cell grid carries a one-cell border (pitch = cells_x + 2, cell (cx, cy) at (cy + 1) * pitch + cx + 1) that stays
// empty, so the 3x3 neighbourhood of any cell can be read without bounds tests.
//
// Is (x, y) within Chebyshev distance d of a point kept in the 3x3 cells around cell index c?  A kept point is stored as
// (y << 16) | x; "both coordinates inside [x-d, x+d] x [y-d, y+d]" is one packed clamp: clamp(q, lo, hi) == q on 16-bit
// halves (VIMNMX.U16x2).  The upper bounds stop at 65534, which no coordinate of a frame of at most 65535 columns / rows
// exceeds, so the empty marker 0xFFFFFFFF is never inside the box.
__device__ __forceinline__ bool near_kept(const uint32_t *cells, int pitch, int c, int x, int y, int d) {
    const uint32_t lo = (uint32_t(max(y - d, 0)) << 16) | uint32_t(max(x - d, 0));
    const uint32_t hi = (uint32_t(min(y + d, 65534)) << 16) | uint32_t(min(x + d, 65534));
    bool hit = false;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const uint32_t q = cells[c + dy * pitch + dx];
            hit |= (__vminu2(__vmaxu2(q, lo), hi) == q);
        }
    }
    return hit;
}

// Best (smallest) key posted to the 8 cells around cell index c.
__device__ __forceinline__ uint64_t neighbour_min(const unsigned long long *cmin, int pitch, int c) {
    uint64_t best = kDeadKey;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            best = min(best, uint64_t(cmin[c + dy * pitch + dx]));
        }
    }
    return best;
}

// Unordered append of this thread's surviving candidate key to a list (warp-aggregated counter bump).
__device__ __forceinline__ void list_push(bool keep, uint64_t value, uint64_t *list, uint32_t *counter) {
    const uint32_t m = __ballot_sync(__activemask(), keep);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0u;
    if (lane_id() == leader) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(__activemask(), base, leader);
    if (keep) list[base + __popc(m & ((1u << lane_id()) - 1u))] = value;
}

__global__ void __launch_bounds__(SELECT_THREADS) select_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int frame = blockIdx.x;
    const int d = p.min_distance;
    const int pitch = p.cells_x + 2;                  // one-cell empty border all round
    const int n_cells = pitch * (p.cells_y + 2);
    // per-cell state, shared (or, for very fine grids, global): best live key posted this round, and the kept point
    unsigned long long *cmin;
    uint32_t *cells;
    if (p.cells_in_smem) {
        cmin = reinterpret_cast<unsigned long long *>(smem);
        cells = reinterpret_cast<uint32_t *>(cmin + n_cells);
    } else {
        cmin = reinterpret_cast<unsigned long long *>(p.cell_scratch + int64_t(frame) * ((n_cells * 3 + 1) & ~1));  // keeps 8-byte alignment
        cells = reinterpret_cast<uint32_t *>(cmin + n_cells);
    }
    __shared__ uint32_t s_kept, s_count[2], s_admit;
    __shared__ uint64_t s_limit;
    // s_sort doubles as the 4096-bin (16 KiB) histogram of the rank-prefix search below
    __shared__ uint64_t s_sort[SELECT_SORT_SMEM > 2048 ? SELECT_SORT_SMEM : 2048];
    uint32_t *hist = reinterpret_cast<uint32_t *>(s_sort);

    const uint32_t count = p.cand_counts[frame];
    if (count > p.cand_capacity && threadIdx.x == 0) atomicExch(p.overflow_flag, 1u);
    const uint32_t n = min(count, p.cand_capacity);
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    // two key lists (live candidates of the current / next round), ping-pong: rounds read their candidates with one
    // coalesced load instead of an index and a gather
    uint64_t *list_a = p.live_scratch + int64_t(frame) * p.cand_capacity * 2;
    uint64_t *list_b = list_a + p.cand_capacity;
    uint64_t *kept = p.kept_keys + int64_t(frame) * p.kept_capacity;
    const uint32_t cell_magic = p.cell_magic;  // ceil(2^32 / (d+1)): exact quotient for coordinates < 65536
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    const uint32_t *mb = p.mask.bits ? p.mask.bits + int64_t(frame) * p.rows * p.mask.words_per_row : nullptr;

    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
        cells[i] = kEmptyCell;
        cmin[i] = kDeadKey;
    }
    if (threadIdx.x == 0) {
        s_kept = 0u;
        s_count[0] = 0u;
        s_count[1] = 0u;
    }
    __syncthreads();

    if (d < 0) {
        // no suppression at all (DrawRectangleInMask clears nothing for a negative distance): everything is kept
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            if (i < uint32_t(p.kept_capacity)) kept[i] = keys[i];
        if (threadIdx.x == 0) s_kept = min(n, uint32_t(p.kept_capacity));
        __syncthreads();
    } else {
        // Each round: (1) every live candidate that a point kept in the previous round covers dies (a kept point covers
        // itself); the others post their key to their cell (64-bit atomicMin) and move to the next round's list;
        // (2) a candidate that holds its cell's minimum and beats the minima of the 8 neighbouring cells has no live
        // better-ranked candidate within d, so the sequential walk would keep it: it is kept now.
        // Only the best-ranked `want` kept points are returned, and a candidate's fate depends on better-ranked candidates
        // only, so the walk may stop at any rank prefix that already yields `want` kept points.  With many candidates the
        // rounds therefore run on rank ranges: first the keys up to the histogram bin (top 12 key bits) that holds the K-th
        // best key; if that keeps too few, the next range (K fourfold) is admitted against the points kept so far.  Exact,
        // and a 4K Harris frame with 5 x 10^5 candidates and needed = 200 touches a few thousand of them.
        uint32_t want_kept = (p.needed > n_pre) ? (p.needed - n_pre) : 1u;
        want_kept = min(want_kept, uint32_t(p.kp_capacity));
        const bool by_prefix = n > SELECT_PREFIX_MIN;
        uint32_t prefix_k = by_prefix ? max(uint32_t(SELECT_PREFIX_MIN / 2), 8u * want_kept) : n;
        if (by_prefix) {   // histogram of the top 12 key bits, once
            for (int i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            // neighbouring candidates mostly share a bin (FAST at a low threshold: 3 x 10^5 keys in a dozen bins), so each warp
            // first groups its lanes by bin and the group leader adds the group's size
            const uint32_t n_warp_rounded = (n + 31u) & ~31u;
            for (uint32_t i = threadIdx.x; i < n_warp_rounded; i += blockDim.x) {
                const uint32_t bin = (i < n) ? uint32_t(__ldg(keys + i) >> 52) : 0xFFFFFFFFu;
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                if (bin != 0xFFFFFFFFu && lane_id() == __ffs(peers) - 1) atomicAdd(hist + bin, uint32_t(__popc(peers)));
            }
            __syncthreads();
        }
        uint64_t lower = 0ull;   // keys below this were admitted by earlier batches
        for (int batch = 0;; ++batch) {
            uint64_t limit = kDeadKey;   // this batch admits lower <= key < limit
            uint32_t admitted = n;       // candidates with key < limit
            if (by_prefix && prefix_k < n) {
                if (threadIdx.x < 32) {   // first bin at which the running count reaches prefix_k (one warp, 128 bins per lane)
                    uint32_t mine = 0u;
#pragma unroll 4
                    for (int b = 0; b < 128; ++b) mine += hist[threadIdx.x * 128 + b];
                    uint32_t incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane_id() >= o) incl += v;
                    }
                    const uint32_t before = incl - mine;
                    if (before < prefix_k && incl >= prefix_k) {   // exactly one lane
                        uint32_t run = before;
                        int b = 0;
#pragma unroll 1
                        for (; b < 128; ++b) {
                            run += hist[threadIdx.x * 128 + b];
                            if (run >= prefix_k) break;
                        }
                        const uint32_t bin = threadIdx.x * 128 + b;
                        s_limit = (bin >= 4095u) ? kDeadKey : (uint64_t(bin + 1u) << 52);
                        s_admit = (bin >= 4095u) ? n : run;
                    }
                }
                __syncthreads();
                limit = s_limit;
                admitted = s_admit;
                __syncthreads();
            }

            if (threadIdx.x == 0) {
                s_count[0] = 0u;
                s_count[1] = 0u;
            }
            __syncthreads();
            uint32_t m = n;          // live candidates entering the round
            const uint64_t *cur = keys;      // round 0 walks the candidate slot itself
            for (int round = 0;; ++round) {
                uint64_t *nxt = (round & 1) ? list_b : list_a;
                uint32_t *nxt_count = &s_count[round & 1];
                const uint32_t rounded = (m + 31u) & ~31u;   // whole warps enter list_push together
                for (uint32_t i = threadIdx.x; i < rounded; i += blockDim.x) {
                    bool live = i < m;
                    uint64_t key = 0ull;
                    int c = 0;
                    if (live) {
                        key = cur[i];
                        const uint32_t xy = cand_key_xy(key);
                        const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
                        const int cx = int(__umulhi(uint32_t(x), cell_magic)), cy = int(__umulhi(uint32_t(y), cell_magic));
                        c = (cy + 1) * pitch + cx + 1;
                        if (round == 0) {
                            live = key >= lower && key < limit;
                            // a candidate on a masked-out pixel is never accepted (feature_point_detector.cpp:66)
                            if (live && mb != nullptr) live = (mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u;
                            // later batches start against everything the better-ranked batches kept
                            if (live && batch > 0) live = !near_kept(cells, pitch, c, x, y, d);
                        } else {
                            live = !near_kept(cells, pitch, c, x, y, d);
                        }
                        if (live) atomicMin(cmin + c, static_cast<unsigned long long>(key));
                    }
                    list_push(live, key, nxt, nxt_count);
                }
                __syncthreads();
                m = *nxt_count;
                if (m == 0u) break;
                for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                    const uint64_t key = nxt[i];
                    const uint32_t xy = cand_key_xy(key);
                    const int cx = int(__umulhi(xy & 0xFFFFu, cell_magic)), cy = int(__umulhi(xy >> 16, cell_magic));
                    const int c = (cy + 1) * pitch + cx + 1;
                    if (uint64_t(cmin[c]) != key) continue;
                    if (key < neighbour_min(cmin, pitch, c)) {
                        const uint32_t slot = atomicAdd(&s_kept, 1u);
                        if (slot < uint32_t(p.kept_capacity)) kept[slot] = key;
                        cells[c] = xy;   // read by the next round (after the barrier below): covers the winner itself too
                    }
                }
                __syncthreads();
                for (int i = threadIdx.x; i < n_cells; i += blockDim.x) cmin[i] = kDeadKey;
                if (threadIdx.x == 0) s_count[(round + 1) & 1] = 0u;
                cur = nxt;
                __syncthreads();
            }
            // enough kept points (or every candidate admitted): done.  Otherwise admit the next, four times larger, rank range;
            // what has been kept so far stays kept (those decisions never depend on worse-ranked candidates).
            if (admitted >= n || s_kept >= want_kept) break;
            lower = limit;
            prefix_k = (prefix_k > n / 4u) ? n : prefix_k * 4u;
            __syncthreads();
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

size_t select_smem_bytes(const SelectArgs &a) { return a.cells_in_smem ? size_t(a.cells_x + 2) * (a.cells_y + 2) * 12 : 0; }

cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream) {
    const size_t smem = select_smem_bytes(args);
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    select_kernel<<<args.n_frames, SELECT_THREADS, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
