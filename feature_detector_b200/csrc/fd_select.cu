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
// around it.  Rounds until no candidate is alive, in one of two forms with identical results:
//   per candidate (select_kernel<false, .>, frames of up to SELECT_CELLS_MIN candidates)
//     A  every live candidate within d of a point kept in the previous round dies; the others post their key
//        to their cell with a shared-memory 64-bit atomicMin and move to the next round's key list;
//     B  a live candidate that holds its cell's minimum and beats the minima of the 8 neighbouring cells has no
//        live better-ranked candidate within d, so the sequential walk would keep it: it is kept now, written
//        to the frame's kept list and to the cell grid.  All such candidates of a round are independent.
//     Work per round is proportional to the candidates still alive, and the first round kills most of them.  While live
//     candidates outnumber cells (the first rounds) step B is one pass over the cell grid in shared memory -- the cells'
//     minima ARE the candidates that can win -- instead of a second pass over the live list in global memory.
//   per cell (select_kernel<true, .>, frames with more candidates: FAST at the reference's default threshold, 4K Harris)
//     the admitted candidates are grouped by cell once; a cell whose best live candidate beats the best of the 8 cells
//     around it keeps it; cells next to a fresh point drop the candidates it covers and recompute their best.  The rounds
//     walk the list of cells that hold candidates, not the grid.
// With more than SELECT_PREFIX_MIN candidates the rounds run on rank ranges (a histogram of the top key bits picks them).
// select_hist_kernel / select_admit_kernel: few frames with many candidates each (a batch of 3840x2160 frames, the gathered
// tiles of one) would leave most SMs idle while each frame's one CTA streams over all of its keys twice; these two kernels
// build the histogram and compact the first rank range with many CTAs per frame, and select_kernel starts from their output.
// The kept list is then ordered by key (a few hundred entries: bitonic in shared memory, or -- few frames, where latency counts -- by
// counting smaller keys, one barrier instead of dozens) and cut at
// max(needed - existing, 1) -- the reference tests the count AFTER each push (:67-68), so needed = 0 still
// yields one feature.  Pre-existing features need no handling here: their squares were already masked out of
// candidate generation with the same d (feature_point_detector.cpp:12-16); they only count toward `needed`.
//
// sort_kernel: bitonic sort of each frame's keys (shared memory when they fit).  Not on the detection path;
// used when the caller asks for the sorted candidate list on the device.
#include "fd_select_common.cuh"

#ifdef FD_SELECT_TRACE
// Debug build only (make EXTRA=-DFD_SELECT_TRACE): clock64 stamps of frame 0's CTA at the phase boundaries of select_kernel,
// read back by fd_debug_select_trace (tools/exp_select_trace.py).  Not part of the shipped library.
__device__ long long g_select_trace[256];
__device__ int g_select_trace_n;
#define SELECT_STAMP(tag)                                                                     \
    do {                                                                                      \
        if (blockIdx.x == 0 && threadIdx.x == 0 && g_select_trace_n < 127) {                  \
            g_select_trace[2 * g_select_trace_n] = clock64();                                 \
            g_select_trace[2 * g_select_trace_n + 1] = (tag);                                 \
            ++g_select_trace_n;                                                               \
        }                                                                                     \
    } while (0)
extern "C" int fd_debug_select_trace(long long *out, int max_pairs) {
    int n = 0;
    cudaMemcpyFromSymbol(&n, g_select_trace_n, sizeof(int));
    n = n < max_pairs ? n : max_pairs;
    cudaMemcpyFromSymbol(out, g_select_trace, size_t(n) * 16);
    int zero = 0;
    cudaMemcpyToSymbol(g_select_trace_n, &zero, sizeof(int));
    return n;
}
#else
#define SELECT_STAMP(tag) do { } while (0)
#endif

namespace fdb {

namespace {

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


// ---- few frames, many candidates (one 3840x2160 frame holds 5 x 10^5): the two passes that stream over ALL of a frame's keys -- the
// rank histogram and the admission of the first rank range -- are what one CTA per frame spends its time on while most SMs idle.
// These two kernels run them with many CTAs per frame; select_kernel then starts from the histogram and the admitted list.
__global__ void __launch_bounds__(256) select_hist_kernel(const SelectArgs p) {
    constexpr int BINS = 1 << SELECT_HIST_BITS;
    __shared__ uint32_t hist[BINS];
    const int frame = blockIdx.y;
    const uint32_t n = min(p.cand_counts[frame], p.cand_capacity);
    if (n <= uint32_t(SELECT_PREFIX_MIN) && !p.hist_always) return;
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    const uint32_t per = ((n + gridDim.x - 1) / gridDim.x + 1023u) & ~1023u;   // whole trips of 4 x 256 keys
    const uint32_t lo = blockIdx.x * per, hi = min(lo + per, n);
    if (lo >= hi) return;
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (uint32_t i0 = lo; i0 < hi; i0 += 4u * blockDim.x) {
        uint64_t k4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
            k4[u] = (i < hi) ? __ldg(keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
            const uint32_t bin = (i < hi) ? uint32_t(k4[u] >> (64 - SELECT_HIST_BITS)) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xffffffffu, bin);
            if (bin != 0xFFFFFFFFu && lane_id() == __ffs(peers) - 1) atomicAdd(hist + bin, uint32_t(__popc(peers)));
        }
    }
    __syncthreads();
    uint32_t *out = p.pre_hist + int64_t(frame) * BINS;
    for (int i = threadIdx.x; i < BINS; i += blockDim.x)
        if (hist[i] != 0u) atomicAdd(out + i, hist[i]);
}

__global__ void __launch_bounds__(256) select_admit_kernel(const SelectArgs p) {
    __shared__ uint64_t s_limit;
    __shared__ uint32_t s_admit;
    const int frame = blockIdx.y;
    const uint32_t n = min(p.cand_counts[frame], p.cand_capacity);
    uint64_t limit;
    if (p.ext_limits != nullptr) {   // a row tile: the limit comes from all tiles' histograms together
        limit = p.ext_limits[frame];
        if (limit == 0ull) return;
    } else {
        if (n <= uint32_t(SELECT_PREFIX_MIN)) return;
        const uint32_t prefix_k = select_first_range(select_want(p, frame));
        if (prefix_k >= n) return;   // the first range is the whole frame: select_kernel walks the candidate slot itself
        if (threadIdx.x < 32) warp_prefix_limit(p.pre_hist + int64_t(frame) * (1 << SELECT_HIST_BITS), prefix_k, n, &s_limit, &s_admit);
        __syncthreads();
        limit = s_limit;
    }
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    uint64_t *out = p.pre_keys + int64_t(frame) * p.pre_capacity;
    const uint32_t per = ((n + gridDim.x - 1) / gridDim.x + 1023u) & ~1023u;
    const uint32_t lo = blockIdx.x * per, hi = min(lo + per, n);
    for (uint32_t i0 = lo; i0 < hi; i0 += 4u * blockDim.x) {
        uint64_t k4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
            k4[u] = (i < hi) ? __ldg(keys + i) : kDeadKey;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) list_push(k4[u] < limit, k4[u], out, p.pre_counts + frame);   // (the padding key is never below a limit)
    }
}

// BY_CELLS picks the form of the rounds; a launch of one form leaves the frames of the other alone (the candidate counts live
// on the device, so the host launches both forms whenever the capacity admits the per-cell one).
// CELLS_SMEM: the per-cell state lives in shared memory (the usual case) -- a template parameter rather than a run-time choice so that
// its accesses compile to LDS / STS / ATOMS with 32-bit addresses instead of generic loads behind 64-bit address arithmetic.
template <bool BY_CELLS, bool CELLS_SMEM>
#ifndef FD_SELECT_MINBLOCKS
#define FD_SELECT_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(SELECT_MAX_THREADS, BY_CELLS ? 1 : FD_SELECT_MINBLOCKS) select_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int frame = blockIdx.x;
    const int d = p.min_distance;
    const int pitch = p.cells_x + 2;                  // one-cell empty border all round
    const int n_cells = pitch * (p.cells_y + 2);
    // per-cell state, shared (or, for very fine grids, global): the best live key of the cell's candidates, the point kept in
    // the cell, where the cell's candidates start in the binned list, and the round in which the cell's point was kept
    uint8_t *cell_base = CELLS_SMEM ? smem : reinterpret_cast<uint8_t *>(p.cell_scratch) + int64_t(frame) * p.cell_stride;
    unsigned long long *cmin = reinterpret_cast<unsigned long long *>(cell_base);
    uint32_t *cells = reinterpret_cast<uint32_t *>(cmin + n_cells);
    uint32_t *cstart = cells + n_cells;                                 // n_cells + 1 entries
    uint16_t *knew = reinterpret_cast<uint16_t *>(cstart + n_cells + 1);
    __shared__ uint32_t s_kept, s_count, s_count2[2], s_admit, s_work, s_active, s_warp[32];
    __shared__ uint64_t s_limit;
    // s_sort doubles as the 2048-bin (8 KiB) histogram of the rank-prefix search below
    __shared__ uint64_t s_sort[SELECT_SORT_SMEM];
    uint32_t *hist = reinterpret_cast<uint32_t *>(s_sort);
    static_assert(SELECT_SORT_SMEM * 8 >= (4 << SELECT_HIST_BITS), "histogram must fit the sort buffer");

    const uint32_t count = p.cand_counts[frame];
    if (count > p.cand_capacity && threadIdx.x == 0) atomicExch(p.overflow_flag, 1u);
    const uint32_t n = min(count, p.cand_capacity);
    if ((n > p.cells_min) != BY_CELLS) return;
    if (p.only_flagged != nullptr && p.only_flagged[frame] == 0) return;   // finished by the first-range launch
    const bool first_range_only = p.need_more != nullptr;   // row tiles: only the gathered first rank range is at hand
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    // binned: the admitted candidates grouped by cell; admitted: the same keys in arrival order, before grouping
    uint64_t *binned = p.live_scratch + int64_t(frame) * p.cand_capacity * 2;
    uint64_t *admitted_keys = binned + p.cand_capacity;
    // during the rounds the same storage lists the cells that must rescan their candidates (never more cells than candidates)
    uint32_t *work = reinterpret_cast<uint32_t *>(admitted_keys);
    uint64_t *kept = p.kept_keys + int64_t(frame) * p.kept_capacity;
    const uint32_t cell_magic = p.cell_magic;  // ceil(2^32 / (d+1)): exact quotient for coordinates < 65536; 0 = cells of one pixel (d = 0)
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    const uint32_t *mb = p.mask.bits ? p.mask.bits + int64_t(frame) * p.rows * p.mask.words_per_row : nullptr;
    const uint32_t xy_xor = p.xy_xor;
    auto key_xy = [&](uint64_t key) { return uint32_t(key) ^ xy_xor; };   // (row << 16) | col
    auto cell_of = [&](uint32_t xy) {
        const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
        const int cx = int(cell_magic ? __umulhi(x, cell_magic) : x), cy = int(cell_magic ? __umulhi(y, cell_magic) : y);
        return (cy + 1) * pitch + cx + 1;
    };

    SELECT_STAMP(1);
    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
        cells[i] = kEmptyCell;
        cmin[i] = kDeadKey;
        knew[i] = 0;
    }
    if (threadIdx.x == 0) s_kept = 0u;
    __syncthreads();
    SELECT_STAMP(2);

    if (first_range_only && (d < 0 || n <= uint32_t(SELECT_PREFIX_MIN))) {   // needs the frame's full key slot: leave it to the flagged launch
        if (threadIdx.x == 0) p.need_more[frame] = 1;
        return;
    }
    if (d < 0) {
        // no suppression at all (DrawRectangleInMask clears nothing for a negative distance): everything is kept
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            if (i < uint32_t(p.kept_capacity)) kept[i] = keys[i];
        if (threadIdx.x == 0) s_kept = min(n, uint32_t(p.kept_capacity));
        __syncthreads();
    } else {
        // Admitted candidates are grouped by cell once (count, scan, scatter); after that a round costs work per CELL, not
        // per live candidate:
        //   (1) a cell whose best live candidate beats the best live candidates of the 8 cells around it keeps that candidate:
        //       no live better-ranked candidate lies within d of it, so the sequential walk would keep it too;
        //   (2) every cell next to a cell that kept a point this round drops its candidates within d of that point (a kept
        //       point covers itself) and recomputes its best live candidate.
        // Only the best-ranked `want` kept points are returned, and a candidate's fate depends on better-ranked candidates
        // only, so the walk may stop at any rank prefix that already yields `want` kept points.  With many candidates the
        // rounds therefore run on rank ranges: first the keys up to the histogram bin (top 11 key bits) that holds the K-th
        // best key; if that keeps too few, the next range (K fourfold) is admitted against the points kept so far.  Exact,
        // and a 4K Harris frame with 5 x 10^5 candidates and needed = 200 touches a few thousand of them.
        const uint32_t want_kept = select_want(p, frame);
        const bool by_prefix = n > SELECT_PREFIX_MIN;
        const bool prepared = by_prefix && p.pre_hist != nullptr;   // select_hist_kernel / select_admit_kernel ran for this launch
        // Two forms of the rounds, same result.  Up to SELECT_CELLS_MIN candidates: work per round proportional to the
        // candidates still alive (the first round kills most of them).  Beyond that -- FAST at the reference's default
        // threshold makes every pixel a candidate, and its raster-ordered ranking needs hundreds of rounds -- the admitted
        // candidates are grouped by cell once and a round costs work per CELL.
        uint32_t prefix_k = by_prefix ? select_first_range(want_kept) : n;
        constexpr int BINS = 1 << SELECT_HIST_BITS;
        if (prepared) {
            for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = p.pre_hist[int64_t(frame) * BINS + i];
            __syncthreads();
        } else if (by_prefix) {   // histogram of the top key bits, once
            for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            // neighbouring candidates mostly share a bin (FAST at a low threshold: 3 x 10^5 keys in a dozen bins), so each warp
            // first groups its lanes by bin and the group leader adds the group's size
            // (four keys per thread per trip: the loads of a trip are in flight together -- this pass is pure streaming)
            for (uint32_t i0 = 0; i0 < n; i0 += 4u * blockDim.x) {
                uint64_t k4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
                    k4[u] = (i < n) ? __ldg(keys + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
                    const uint32_t bin = (i < n) ? uint32_t(k4[u] >> (64 - SELECT_HIST_BITS)) : 0xFFFFFFFFu;
                    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                    if (bin != 0xFFFFFFFFu && lane_id() == __ffs(peers) - 1) atomicAdd(hist + bin, uint32_t(__popc(peers)));
                }
            }
            __syncthreads();
        }
        uint64_t lower = 0ull;   // keys below this were admitted by earlier batches
        uint32_t stamp = 0u;     // round counter behind knew[]
        for (int batch = 0;; ++batch) {
            uint64_t limit = kDeadKey;   // this batch admits lower <= key < limit
            uint32_t admitted = n;       // candidates with key < limit
            if (by_prefix && prefix_k < n) {
                if (threadIdx.x < 32) warp_prefix_limit(hist, prefix_k, n, &s_limit, &s_admit);
                __syncthreads();
                limit = s_limit;
                admitted = s_admit;
                __syncthreads();
            }

            // the first rank range of a prepared frame was compacted by select_admit_kernel: walk that list instead of the whole slot
            const bool from_list = prepared && batch == 0 && limit != kDeadKey;
            if (first_range_only && !from_list) {   // (uniform over the CTA) no compacted range, or it did not yield enough points
                if (threadIdx.x == 0) p.need_more[frame] = 1;
                return;
            }
            const uint64_t *src = from_list ? p.pre_keys + int64_t(frame) * p.pre_capacity : keys;
            const uint32_t src_n = from_list ? min(p.pre_counts[frame], p.pre_capacity) : n;
            if constexpr (BY_CELLS) {
                // ---- admit: keys of this rank range that are not masked out and not covered by what is already kept ----
                for (int i = threadIdx.x; i <= n_cells; i += blockDim.x) cstart[i] = 0u;
                if (threadIdx.x == 0) s_count = 0u;
                __syncthreads();
                const bool everything = !by_prefix && mb == nullptr;   // one batch, nothing to filter: the candidate slot itself is the list
                if (everything) {
                    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(cstart + cell_of(key_xy(__ldg(keys + i))), 1u);
                } else {
                    // four keys per thread per trip (loads in flight together); whole warps enter list_push together
                    for (uint32_t i0 = 0; i0 < src_n; i0 += 4u * blockDim.x) {
                        uint64_t k4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
                            k4[u] = (i < src_n) ? __ldg(src + i) : kDeadKey;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint64_t key = k4[u];
                            bool live = key >= lower && key < limit;   // (the padding key is never below a limit)
                            if (live) {
                                const uint32_t xy = key_xy(key);
                                const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
                                const int c = cell_of(xy);
                                // a candidate on a masked-out pixel is never accepted (feature_point_detector.cpp:66)
                                if (mb != nullptr) live = (mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u;
                                // later batches start against everything the better-ranked batches kept
                                if (live && batch > 0) live = !near_kept(cells, pitch, c, x, y, d);
                                if (live) atomicAdd(cstart + c, 1u);
                            }
                            list_push(live, key, admitted_keys, &s_count);
                        }
                    }
                }
                __syncthreads();
                const uint32_t m = everything ? n : s_count;
                const uint64_t *source = everything ? keys : admitted_keys;
                if (m != 0u) {
                    // ---- group by cell: exclusive scan of the per-cell counts, then scatter ----
                    {
                        const int n_tot = n_cells + 1;
                        const int per = (n_tot + int(blockDim.x) - 1) / int(blockDim.x);
                        const int b0 = min(int(threadIdx.x) * per, n_tot), b1 = min(b0 + per, n_tot);
                        uint32_t sum = 0u;
                        for (int b = b0; b < b1; ++b) sum += cstart[b];
                        uint32_t run = block_exclusive_scan(sum, s_warp);
                        for (int b = b0; b < b1; ++b) {
                            const uint32_t v = cstart[b];
                            cstart[b] = run;
                            run += v;
                        }
                    }
                    __syncthreads();
                    // scatter: cstart[c] runs from the start of cell c to its end, i.e. to the start of cell c + 1, so afterwards
                    // the candidates of cell c are binned[cstart[c - 1] .. cstart[c]) (cell 0 is a border cell and stays empty)
                    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                        const uint64_t key = source[i];
                        const int c = cell_of(key_xy(key));
                        binned[atomicAdd(cstart + c, 1u)] = key;
                        atomicMin(cmin + c, static_cast<unsigned long long>(key));
                    }
                    __syncthreads();
                    // the cells that hold candidates: a fine grid (19 000 cells of a 3840x2160 frame) is mostly empty once the rounds run on
                    // a rank range of a few thousand keys, and the rounds below only walk this list
                    uint32_t *active = work + p.cand_capacity;   // second half of the list storage (never more cells than candidates)
                    if (threadIdx.x == 0) s_active = 0u;
                    __syncthreads();
                    for (int c = 1 + threadIdx.x; c < n_cells; c += blockDim.x)
                        if (cstart[c] != cstart[c - 1]) active[atomicAdd(&s_active, 1u)] = uint32_t(c);
                    __syncthreads();
                    const int n_active = int(s_active);

                    // ---- rounds ----
                    const int sub = lane_id() & 7, group = int(threadIdx.x >> 3), n_groups = int(blockDim.x >> 3);
                    const uint32_t group_mask = 0xFFu << (lane_id() & 24);
                    for (;;) {
                        if (++stamp == 0xFFFFu) {   // the 16-bit round stamps are about to wrap: forget the old ones
                            for (int i = threadIdx.x; i < n_cells; i += blockDim.x) knew[i] = 0;
                            stamp = 1u;
                            __syncthreads();
                        }
                        if (threadIdx.x == 0) s_work = 0u;
                        for (int a = threadIdx.x; a < n_active; a += blockDim.x) {
                            const int c = int(active[a]);
                            const uint64_t mine = cmin[c];
                            if (mine == kDeadKey) continue;   // no live candidate left in the cell
                            if (mine < neighbour_min(cmin, pitch, c)) {
                                const uint32_t slot = atomicAdd(&s_kept, 1u);
                                if (slot < uint32_t(p.kept_capacity)) kept[slot] = mine;
                                cells[c] = key_xy(mine);
                                knew[c] = uint16_t(stamp);
                            }
                        }
                        __syncthreads();
                        // cells with live candidates next to a point kept this round go on the work list
                        bool alive = false;
                        for (int a = threadIdx.x; a < n_active; a += blockDim.x) {
                            const int c = int(active[a]);
                            if (cmin[c] == kDeadKey) continue;
                            bool fresh = false;
#pragma unroll
                            for (int k = 0; k < 9; ++k) fresh |= knew[c + (k / 3 - 1) * pitch + (k % 3 - 1)] == uint16_t(stamp);
                            if (fresh) work[atomicAdd(&s_work, 1u)] = uint32_t(c);
                            else alive = true;
                        }
                        __syncthreads();
                        // eight lanes per listed cell: drop the candidates the fresh points cover, recompute the cell's best
                        const int n_work = int(s_work);
                        for (int w = group; w < n_work; w += n_groups) {
                            const int c = int(work[w]);
                            // the boxes of pixels the fresh points cover, as packed 16-bit bounds (0xFFFFFFFF: matches no pixel); up to
                            // three fresh points -- nearly always one -- take the short path
                            uint32_t lo0 = 0xFFFFFFFFu, hi0 = 0xFFFFFFFFu, lo1 = 0xFFFFFFFFu, hi1 = 0xFFFFFFFFu, lo2 = 0xFFFFFFFFu, hi2 = 0xFFFFFFFFu;
                            int n_fresh = 0;
#pragma unroll
                            for (int k = 0; k < 9; ++k) {
                                const int nb = c + (k / 3 - 1) * pitch + (k % 3 - 1);
                                if (knew[nb] == uint16_t(stamp)) {
                                    const uint32_t q = cells[nb];
                                    const int qx = int(q & 0xFFFFu), qy = int(q >> 16);
                                    const uint32_t lo = (uint32_t(max(qy - d, 0)) << 16) | uint32_t(max(qx - d, 0));
                                    const uint32_t hi = (uint32_t(min(qy + d, 65534)) << 16) | uint32_t(min(qx + d, 65534));
                                    if (n_fresh == 0) lo0 = lo, hi0 = hi;
                                    else if (n_fresh == 1) lo1 = lo, hi1 = hi;
                                    else if (n_fresh == 2) lo2 = lo, hi2 = hi;
                                    ++n_fresh;
                                }
                            }
                            const uint32_t js = cstart[c - 1] + sub, je = cstart[c];
                            uint64_t best = kDeadKey;
                            if (n_fresh <= 3) {
                                for (uint32_t j = js; j < je; j += 32) {   // four candidates per lane per trip: their loads are in flight together
                                    uint64_t key[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) key[u] = (j + 8 * u < je) ? binned[j + 8 * u] : kDeadKey;
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const uint32_t xy = key_xy(key[u]);
                                        const bool hit = key[u] != kDeadKey && ((__vminu2(__vmaxu2(xy, lo0), hi0) == xy) | (__vminu2(__vmaxu2(xy, lo1), hi1) == xy) |
                                                                                (__vminu2(__vmaxu2(xy, lo2), hi2) == xy));
                                        if (hit) binned[j + 8 * u] = kDeadKey;
                                        else best = min(best, key[u]);
                                    }
                                }
                            } else {
                                for (uint32_t j = js; j < je; j += 8) {
                                    const uint64_t key = binned[j];
                                    const uint32_t xy = key_xy(key);
                                    const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
                                    if (key != kDeadKey && near_kept(cells, pitch, c, x, y, d)) binned[j] = kDeadKey;   // every kept point, fresh or not
                                    else best = min(best, key);
                                }
                            }
#pragma unroll
                            for (int o = 1; o < 8; o <<= 1) best = min(best, __shfl_xor_sync(group_mask, best, o));
                            if (sub == 0) cmin[c] = best;
                            alive |= best != kDeadKey;
                        }
                        if (!__syncthreads_or(alive)) break;
                    }
                }
            } else {
                // ---- candidate-centric rounds ----
                // (1) every live candidate that a point kept in the previous round covers dies (a kept point covers itself);
                //     the others post their key to their cell (64-bit atomicMin) and move to the next round's list;
                // (2) a candidate that holds its cell's minimum and beats the minima of the 8 neighbouring cells is kept.
                uint64_t *list_a = binned, *list_b = admitted_keys;
                if (threadIdx.x == 0) {
                    s_count2[0] = 0u;
                    s_count2[1] = 0u;
                }
                __syncthreads();
                uint32_t m = src_n;           // live candidates entering the round
                const uint64_t *cur = src;    // round 0 walks the candidate slot itself (or the prepared first range)
                for (int round = 0;; ++round) {
                    uint64_t *nxt = (round & 1) ? list_b : list_a;
                    uint32_t *nxt_count = &s_count2[round & 1];
                    const uint32_t rounded = (m + 31u) & ~31u;   // whole warps enter list_push together
                    // the first round of an unfiltered frame (no rank range, no mask) keeps every candidate alive: its output list would be
                    // a copy of its input, so nothing is pushed and the next round walks the candidate slot again
                    const bool all_live = round == 0 && batch == 0 && mb == nullptr && limit == kDeadKey;
                    for (uint32_t i = threadIdx.x; i < rounded; i += blockDim.x) {
                        bool live = i < m;
                        uint64_t key = 0ull;
                        SELECT_STAMP(40);
                        if (live) {
                            key = cur[i];
#ifdef FD_SELECT_TRACE
                            key = __shfl_sync(__activemask(), key, lane_id());
                            SELECT_STAMP(41);
#endif
                            const uint32_t xy = key_xy(key);
                            const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
                            const int c = cell_of(xy);
                            if (round == 0) {
                                live = key >= lower && key < limit;
                                // a candidate on a masked-out pixel is never accepted (feature_point_detector.cpp:66)
                                if (live && mb != nullptr) live = (mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u;
                                // later batches start against everything the better-ranked batches kept
                                if (live && batch > 0) live = !near_kept(cells, pitch, c, x, y, d);
                            } else {
                                live = !near_kept(cells, pitch, c, x, y, d);
                            }
#ifdef FD_SELECT_TRACE
                            live = __shfl_sync(__activemask(), int(live), lane_id()) != 0;
                            SELECT_STAMP(42);
#endif
                            if (live) atomicMin(cmin + c, static_cast<unsigned long long>(key));
                        }
                        SELECT_STAMP(43);
                        if (!all_live) list_push(live, key, nxt, nxt_count);
                        SELECT_STAMP(44);
                    }
                    SELECT_STAMP(45);
                    __syncthreads();
                    SELECT_STAMP(10);
                    const uint64_t *alive = all_live ? cur : nxt;   // the candidates entering the next round
                    if (!all_live) m = *nxt_count;
                    if (m == 0u) break;
                    // (a round's winners -- dozens in the first rounds -- take their kept-list slots a warp at a time: one shared-memory
                    // atomic per warp, not per winner; at most one point is ever kept per cell, so the list cannot overflow)
                    if (m > uint32_t(n_cells)) {
                        // more live candidates than cells (the first rounds): the cells' minima ARE the candidates that can win, so the
                        // winners come from one pass over the cell grid in shared memory instead of a second pass over the live list
                        for (int c0 = 0; c0 < n_cells; c0 += blockDim.x) {
                            const int c = c0 + int(threadIdx.x);
                            uint64_t key = kDeadKey;
                            if (c < n_cells) key = cmin[c];
                            const bool win = key != kDeadKey && key < neighbour_min(cmin, pitch, c);
                            if (win) cells[c] = key_xy(key);
                            list_push(win, key, kept, &s_kept);
                        }
                    } else {
                        for (uint32_t i0 = 0; i0 < m; i0 += blockDim.x) {
                            const uint32_t i = i0 + threadIdx.x;
                            uint64_t key = kDeadKey;
                            bool win = false;
                            if (i < m) {
                                key = alive[i];
                                const int c = cell_of(key_xy(key));
                                win = uint64_t(cmin[c]) == key && key < neighbour_min(cmin, pitch, c);
                                if (win) cells[c] = key_xy(key);   // read by the next round (after the barrier below): covers the winner itself too
                            }
                            list_push(win, key, kept, &s_kept);
                        }
                    }
                    __syncthreads();
                    SELECT_STAMP(11);
                    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) cmin[i] = kDeadKey;
                    if (threadIdx.x == 0) s_count2[(round + 1) & 1] = 0u;
                    cur = alive;
                    __syncthreads();
                    SELECT_STAMP(12);
                }
            }
            // enough kept points (or every candidate admitted): done.  Otherwise admit the next, four times larger, rank range;
            // what has been kept so far stays kept (those decisions never depend on worse-ranked candidates).
            if (admitted >= n || s_kept >= want_kept) break;
            lower = limit;
            prefix_k = (prefix_k > n / 4u) ? n : prefix_k * 4u;
            __syncthreads();
        }
    }

    SELECT_STAMP(30);
    // ---- order the kept points by rank and cut at the number the reference would have pushed ----
    __syncthreads();
    const uint32_t n_kept = min(s_kept, uint32_t(p.kept_capacity));
    uint32_t want = (p.needed > n_pre) ? (p.needed - n_pre) : 1u;  // pushed, then tested: at least one
    want = min(want, uint32_t(p.kp_capacity));
    const uint32_t n_out = min(n_kept, want);
    float4 *kp_out = p.keypoints + int64_t(frame) * p.kp_capacity;
    if (p.few_frames && n_kept <= uint32_t(SELECT_RANK_COUNT_MAX)) {
        // Few frames (a CTA has its SM to itself, what counts is latency) and up to a few hundred distinct keys: a key's place in
        // the order is the number of smaller keys, counted by 2 to 32 lanes per key, and the points that make the cut go straight to
        // their output slots.  One barrier where the bitonic network takes dozens (5 us for 128 keys); its n^2 comparisons are more
        // instructions than the network's, which is what counts when the SMs are full of other frames' CTAs.
        for (uint32_t i = threadIdx.x; i < n_kept; i += blockDim.x) s_sort[i] = kept[i];
        __syncthreads();
        SELECT_STAMP(31);
        uint32_t team = 1u;   // lanes per key
        while (team < 32u && n_kept * team * 2u <= blockDim.x) team <<= 1;
        const uint32_t rounded = (n_kept * team + 31u) & ~31u;   // whole warps take part in the shuffles
        for (uint32_t t = threadIdx.x; t < rounded; t += blockDim.x) {
            const uint32_t i = t / team, part = t % team;
            const uint64_t key = i < n_kept ? s_sort[i] : 0ull;
            uint32_t place = 0u;
            if (i < n_kept)
                for (uint32_t j = part; j < n_kept; j += team) place += s_sort[j] < key;
            for (uint32_t o = 1u; o < team; o <<= 1) place += __shfl_xor_sync(0xffffffffu, place, o);
            if (i < n_kept && part == 0u && place < n_out) {
                const uint32_t xy = key_xy(key);
                kp_out[place] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);  // Vec2(pixel.x(), pixel.y()), :67
            }
        }
    } else {
        if (n_kept <= uint32_t(SELECT_SORT_SMEM)) {
            for (uint32_t i = threadIdx.x; i < n_kept; i += blockDim.x) s_sort[i] = kept[i];
            __syncthreads();
            block_bitonic_sort(s_sort, n_kept);
            kept = s_sort;
        } else {
            block_bitonic_sort(kept, n_kept);
        }
        __syncthreads();
        SELECT_STAMP(31);
        for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
            const uint64_t key = kept[i];
            const uint32_t xy = key_xy(key);
            kp_out[i] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);
        }
    }
    if (threadIdx.x == 0) p.kp_counts[frame] = int32_t(n_out);
    SELECT_STAMP(32);
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

size_t select_cell_bytes(int cells_x, int cells_y) {   // cmin (8) + kept point (4) + list start (4, one extra entry) + round stamp (2) per cell
    const size_t n = size_t(cells_x + 2) * (cells_y + 2);
    return (n * 18 + 4 + 15) & ~size_t(15);
}

size_t select_smem_bytes(const SelectArgs &a) { return a.cells_in_smem ? select_cell_bytes(a.cells_x, a.cells_y) : 0; }

namespace {
dim3 prepare_grid(int n_frames) { return dim3(unsigned(std::max(1, 4 * 148 / std::max(n_frames, 1))), unsigned(n_frames)); }
}  // namespace

cudaError_t launch_select_hist(const SelectArgs &args, cudaStream_t stream) {
    select_hist_kernel<<<prepare_grid(args.n_frames), 256, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_select_admit(const SelectArgs &args, cudaStream_t stream) {
    select_admit_kernel<<<prepare_grid(args.n_frames), 256, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_select(const SelectArgs &args, cudaStream_t stream) {
    const size_t smem = select_smem_bytes(args);
    if (args.pre_hist != nullptr && args.need_more == nullptr) {   // few frames with room for many candidates each: histogram and first range by many CTAs per frame
        select_hist_kernel<<<prepare_grid(args.n_frames), 256, 0, stream>>>(args);
        select_admit_kernel<<<prepare_grid(args.n_frames), 256, 0, stream>>>(args);
    }
    // one CTA per frame: with fewer frames than SMs (the drop-in classes' one frame per call) a CTA has its SM to itself, and the
    // passes that stream over a frame's candidates are what its latency is made of
    const int threads = (args.cells_in_smem && args.n_frames > 148) ? SELECT_THREADS : SELECT_MAX_THREADS;
    auto launch = [&](auto kernel) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        kernel<<<args.n_frames, threads, smem, stream>>>(args);
        return cudaSuccess;
    };
    cudaError_t e = args.cells_in_smem ? launch(select_kernel<false, true>) : launch(select_kernel<false, false>);
    if (e != cudaSuccess) return e;
    if (args.cand_capacity > args.cells_min) {   // some frame may hold enough candidates for the per-cell form
        e = args.cells_in_smem ? launch(select_kernel<true, true>) : launch(select_kernel<true, false>);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace fdb
