// Kernel 3, winner-driven form -- greedy minimum-distance selection (reference
// src/feature_point_detector/feature_point_detector.cpp:54-74, 76-88), same results as fd_select.cu key for key.
//
// Same reformulation as fd_select.cu: cells of side d+1 hold at most one kept point, everything within d of a pixel lies in the
// 3x3 cells around it, and a candidate that outranks every live candidate of those 3x3 cells is kept by the sequential walk too.
// What differs is what a round costs.  The forms in fd_select.cu touch every live candidate (or every cell) in every round; here
// the candidates are grouped by cell once (count, scan, scatter -- the scatter also posts each key to its cell's best), and a
// round only touches what the points kept in it can change:
//   K  the cells on the evaluation list compare their best live candidate with the best of the 8 cells around them; winners are kept;
//   Z  per winner, the bests of its 3x3 cells are cleared and the cells of the 5x5 block around it that hold candidates (those whose
//      3x3 neighbourhood is about to change) go on the next evaluation list;
//   R  one warp per winner walks the candidates of its 3x3 cells -- three contiguous runs of the binned list, a lane per
//      candidate, the loads of a run in flight together: a candidate within d of a kept point is marked dead, the others post
//      their key to their cell's best again.
// Three barriers per round, 7 to 12 rounds on the synthetic frames, work proportional to kept points x candidates per 3x3 cells
// (about twice the candidate count in total) on full warps, where the per-cell form of fd_select.cu rescans whole cells with
// eight lanes and the per-candidate form revisits every live candidate every round.  With more than SELECT_PREFIX_MIN candidates
// the rounds run on rank ranges exactly as in fd_select.cu.  One CTA per frame; frames whose cell grid does not fit shared
// memory stay with fd_select.cu.
#include "fd_select_common.cuh"

namespace fdb {

namespace {

constexpr int LEAN_BORDER = 2;   // empty cells round the grid: the 5x5 block around any cell can be addressed without bounds tests

__global__ void __launch_bounds__(SELECT_LEAN_MAX_THREADS, SELECT_LEAN_MIN_CTAS) select_lean_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int frame = blockIdx.x;
    const int d = p.min_distance;
    const int pitch = p.cells_x + 2 * LEAN_BORDER;
    const int n_cells = pitch * (p.cells_y + 2 * LEAN_BORDER);
    // per cell: best live key (8 B), list start (4 B, one extra entry), kept point (4 B), list stamp (4 B), two list entries (2 x 2 B)
    unsigned long long *best = reinterpret_cast<unsigned long long *>(smem);
    uint32_t *cstart = reinterpret_cast<uint32_t *>(best + n_cells);   // after the scatter the candidates of cell c are binned[cstart[c - 1] .. cstart[c])
    uint32_t *kept_xy = cstart + n_cells + 1;                         // (row << 16) | col of the point kept in the cell, or kEmptyCell
    uint32_t *mark = kept_xy + n_cells;                               // last evaluation-list stamp of the cell
    uint16_t *klist = reinterpret_cast<uint16_t *>(mark + n_cells);   // cells to evaluate
    uint16_t *wlist = klist + n_cells;                                // cells that kept a point this round
    __shared__ uint32_t s_kept, s_nk, s_nw, s_admit, s_warp[32];
    __shared__ uint64_t s_limit;
    __shared__ uint64_t s_sort[SELECT_SORT_SMEM];   // doubles as the rank-prefix histogram
    uint32_t *hist = reinterpret_cast<uint32_t *>(s_sort);
    static_assert(SELECT_SORT_SMEM * 8 >= (4 << SELECT_HIST_BITS), "histogram must fit the sort buffer");

    const uint32_t count = p.cand_counts[frame];
    if (count > p.cand_capacity && threadIdx.x == 0) atomicExch(p.overflow_flag, 1u);
    const uint32_t n = min(count, p.cand_capacity);
    if (n >= p.lean_limit) return;   // fd_select.cu's forms take the frame
    const uint64_t *keys = p.cand_keys + int64_t(frame) * p.cand_capacity;
    uint64_t *binned = p.live_scratch + int64_t(frame) * p.cand_capacity * 2;
    uint64_t *kept = p.kept_keys + int64_t(frame) * p.kept_capacity;
    const uint32_t cell_magic = p.cell_magic;
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    const uint32_t *mb = p.mask.bits ? p.mask.bits + int64_t(frame) * p.rows * p.mask.words_per_row : nullptr;
    const uint32_t xy_xor = p.xy_xor;
    auto cell_of = [&](uint32_t xy) {
        const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
        const int cx = int(cell_magic ? __umulhi(x, cell_magic) : x), cy = int(cell_magic ? __umulhi(y, cell_magic) : y);
        return (cy + LEAN_BORDER) * pitch + cx + LEAN_BORDER;
    };

    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
        kept_xy[i] = kEmptyCell;
        mark[i] = 0u;
        best[i] = kDeadKey;
    }
    if (threadIdx.x == 0) s_kept = 0u;
    uint32_t want_kept = (p.needed > n_pre) ? (p.needed - n_pre) : 1u;   // pushed, then tested: at least one (feature_point_detector.cpp:67-68)
    want_kept = min(want_kept, uint32_t(p.kp_capacity));
    const bool by_prefix = n > SELECT_PREFIX_MIN;
    uint32_t prefix_k = by_prefix ? max(uint32_t(SELECT_PREFIX_MIN / 2), 8u * want_kept) : n;
    constexpr int BINS = 1 << SELECT_HIST_BITS;
    __syncthreads();
    if (by_prefix) {   // histogram of the top key bits, once (as in fd_select.cu)
        for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
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

    const int lane = lane_id(), warp = int(threadIdx.x >> 5), n_warps = int(blockDim.x >> 5);
    uint64_t lower = 0ull;   // keys below this were admitted by earlier batches
    uint32_t round = 0u;     // evaluation-list stamp
    for (int batch = 0;; ++batch) {
        uint64_t limit = kDeadKey;   // this batch admits lower <= key < limit
        uint32_t admitted = n;       // candidates with key < limit
        if (by_prefix && prefix_k < n) {
            if (threadIdx.x < 32) {   // first bin at which the running count reaches prefix_k
                constexpr int PER = BINS / 32;
                uint32_t mine = 0u;
#pragma unroll 4
                for (int b = 0; b < PER; ++b) mine += hist[threadIdx.x * PER + b];
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t before = incl - mine;
                if (before < prefix_k && incl >= prefix_k) {   // exactly one lane
                    uint32_t run = before;
                    int b = 0;
#pragma unroll 1
                    for (; b < PER; ++b) {
                        run += hist[threadIdx.x * PER + b];
                        if (run >= prefix_k) break;
                    }
                    const uint32_t bin = threadIdx.x * PER + b;
                    s_limit = (bin >= uint32_t(BINS - 1)) ? kDeadKey : (uint64_t(bin + 1u) << (64 - SELECT_HIST_BITS));
                    s_admit = (bin >= uint32_t(BINS - 1)) ? n : run;
                }
            }
            __syncthreads();
            limit = s_limit;
            admitted = s_admit;
        }

        // ---- admit and group by cell: count, scan, scatter (which also posts every key to its cell's best) ----
        ++round;
        for (int i = threadIdx.x; i <= n_cells; i += blockDim.x) cstart[i] = 0u;
        if (threadIdx.x == 0) s_nk = s_nw = 0u;
        __syncthreads();
        // a candidate of this rank range is admitted unless it sits on a masked-out pixel (feature_point_detector.cpp:66) or a point
        // kept by a better-ranked batch covers it
        auto admit = [&](uint64_t key, int &c) {
            if (key < lower || key >= limit) return false;
            const uint32_t xy = uint32_t(key) ^ xy_xor;
            const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
            c = cell_of(xy);
            if (mb != nullptr && !((mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u)) return false;
            if (batch > 0 && near_kept(kept_xy, pitch, c, x, y, d)) return false;
            return true;
        };
        for (uint32_t i0 = 0; i0 < n; i0 += 4u * blockDim.x) {
            uint64_t k4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
                k4[u] = (i < n) ? __ldg(keys + i) : kDeadKey;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int c;
                if (admit(k4[u], c)) atomicAdd(cstart + c, 1u);
            }
        }
        __syncthreads();
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
                // the first evaluation list: every cell that receives a candidate
                if (v != 0u) klist[atomicAdd(&s_nk, 1u)] = uint16_t(b);
            }
        }
        __syncthreads();
        for (uint32_t i0 = 0; i0 < n; i0 += 4u * blockDim.x) {
            uint64_t k4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * blockDim.x + threadIdx.x;
                k4[u] = (i < n) ? __ldg(keys + i) : kDeadKey;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int c;
                if (admit(k4[u], c)) {
                    binned[atomicAdd(cstart + c, 1u)] = k4[u];
                    atomicMin(best + c, static_cast<unsigned long long>(k4[u]));
                }
            }
        }
        __syncthreads();

        // ---- rounds ----
        for (;;) {
            const int n_k = int(s_nk);
            if (n_k == 0) break;
            // K: a listed cell whose best live candidate outranks the best of the 8 cells around it keeps it
            for (int w = threadIdx.x; w < n_k; w += blockDim.x) {
                const int c = int(klist[w]);
                const uint64_t mine = best[c];
                if (mine == kDeadKey) continue;
                uint64_t nb_min = kDeadKey;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (k != 4) nb_min = min(nb_min, uint64_t(best[c + (k / 3 - 1) * pitch + (k % 3 - 1)]));
                if (mine >= nb_min) continue;
                const uint32_t slot = atomicAdd(&s_kept, 1u);
                if (slot < uint32_t(p.kept_capacity)) kept[slot] = mine;
                kept_xy[c] = uint32_t(mine) ^ xy_xor;
                wlist[atomicAdd(&s_nw, 1u)] = uint16_t(c);
            }
            __syncthreads();
            const int n_w = int(s_nw);
            ++round;
            if (threadIdx.x == 0) s_nk = 0u;   // everyone read it before the K phase
            // nothing was kept: no listed cell is alive, and a live cell off the list lost to a neighbour that has not changed since --
            // following those neighbours ends at a live cell that beats all of its own, which would be listed.  So nothing is alive.
            if (n_w == 0) break;
            __syncthreads();
            // Z: 32 lanes per winner.  Lanes 0..24 put the cells of the 5x5 block that hold candidates and no kept point on the next
            // evaluation list, once each; lanes 0..8 clear the bests of the 3x3 cells (the R phase posts the survivors again).
            for (int w = warp; w < n_w; w += n_warps) {
                const int c = int(wlist[w]);
                if (lane < 25) {
                    const int nb = c + (lane / 5 - 2) * pitch + (lane % 5 - 2);
                    const bool has_candidates = nb >= 1 && cstart[nb] != cstart[nb - 1];
                    if (has_candidates && kept_xy[nb] == kEmptyCell && atomicMax(mark + nb, round) < round) klist[atomicAdd(&s_nk, 1u)] = uint16_t(nb);
                }
                if (lane < 9) best[c + (lane / 3 - 1) * pitch + (lane % 3 - 1)] = kDeadKey;
            }
            __syncthreads();
            // R: one warp per winner over the candidates of its 3x3 cells: three runs of the binned list (cells c-1, c, c+1 of a cell row
            // are neighbours in it), a lane per candidate.  A cell next to two winners is walked twice, to the same effect.
            for (int w = warp; w < n_w; w += n_warps) {
                const int c = int(wlist[w]);
#pragma unroll
                for (int r = -1; r <= 1; ++r) {
                    const int cm = c + r * pitch;
                    const uint32_t e0 = cstart[cm - 1], e1 = cstart[cm], j_end = cstart[cm + 1];   // ends of cells cm-1 and cm inside the run
                    for (uint32_t j = cstart[cm - 2] + lane; j < j_end; j += 32) {
                        const uint64_t key = binned[j];
                        if (key == kDeadKey) continue;
                        const uint32_t xy = uint32_t(key) ^ xy_xor;
                        const int cc = cm - 1 + (j >= e0 ? 1 : 0) + (j >= e1 ? 1 : 0);
                        if (near_kept(kept_xy, pitch, cc, int(xy & 0xFFFFu), int(xy >> 16), d)) binned[j] = kDeadKey;
                        else atomicMin(best + cc, static_cast<unsigned long long>(key));
                    }
                }
            }
            if (threadIdx.x == 0) s_nw = 0u;   // everyone read it before the Z phase; the next K phase appends after the barrier below
            __syncthreads();
        }
        // enough kept points (or every candidate admitted): done.  Otherwise admit the next, four times larger, rank range
        if (admitted >= n || s_kept >= want_kept) break;
        lower = limit;
        prefix_k = (prefix_k > n / 4u) ? n : prefix_k * 4u;
        __syncthreads();
    }

    // ---- order the kept points by rank and cut at the number the reference would have pushed ----
    __syncthreads();
    const uint32_t n_kept = min(s_kept, uint32_t(p.kept_capacity));
    float4 *kp_out = p.keypoints + int64_t(frame) * p.kp_capacity;
    const uint32_t n_out = min(n_kept, want_kept);
    if (n_kept <= uint32_t(SELECT_SORT_SMEM)) {
        for (uint32_t i = threadIdx.x; i < n_kept; i += blockDim.x) s_sort[i] = kept[i];
        __syncthreads();
        if (n_kept > 384u) {
            block_bitonic_sort(s_sort, n_kept);
            for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
                const uint64_t key = s_sort[i];
                const uint32_t xy = uint32_t(key) ^ xy_xor;
                kp_out[i] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);
            }
        } else {
            // rank by counting: kept points are few (at most one per cell), and every thread reads the same word at a time
            for (uint32_t i = threadIdx.x; i < n_kept; i += blockDim.x) {
                const uint64_t key = s_sort[i];
                uint32_t rank = 0u;
                for (uint32_t j = 0; j < n_kept; ++j) rank += (s_sort[j] < key) ? 1u : 0u;
                if (rank < n_out) {
                    const uint32_t xy = uint32_t(key) ^ xy_xor;
                    kp_out[rank] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);  // Vec2(pixel.x(), pixel.y()), :67
                }
            }
        }
    } else {
        block_bitonic_sort(kept, n_kept);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
            const uint64_t key = kept[i];
            const uint32_t xy = uint32_t(key) ^ xy_xor;
            kp_out[i] = make_float4(float(xy & 0xFFFFu), float(xy >> 16), cand_key_response(key), 0.0f);
        }
    }
    if (threadIdx.x == 0) p.kp_counts[frame] = int32_t(n_out);
}

}  // namespace

size_t select_lean_smem_bytes(int cells_x, int cells_y) {   // 24 bytes per cell of the bordered grid, one extra list start
    const size_t n = size_t(cells_x + 2 * LEAN_BORDER) * (cells_y + 2 * LEAN_BORDER);
    return (n * 24 + 4 + 15) & ~size_t(15);
}

bool select_lean_grid_fits(int cells_x, int cells_y) {   // list entries are 16-bit cell indices
    return size_t(cells_x + 2 * LEAN_BORDER) * (cells_y + 2 * LEAN_BORDER) < 65536 && select_lean_smem_bytes(cells_x, cells_y) <= SELECT_LEAN_SMEM_MAX;
}

cudaError_t launch_select_lean(const SelectArgs &args, int threads, cudaStream_t stream) {
    const size_t smem = select_lean_smem_bytes(args.cells_x, args.cells_y);
    cudaError_t e = cudaFuncSetAttribute(select_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    select_lean_kernel<<<args.n_frames, threads, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
