// Kernel 3, event-driven form -- greedy minimum-distance selection (reference
// src/feature_point_detector/feature_point_detector.cpp:54-74, 76-88), same results as fd_select.cu key for key.
//
// Same reformulation as fd_select.cu: cells of side d+1 hold at most one kept point, everything within d of a pixel lies in the
// 3x3 cells around it, and a candidate that outranks every live candidate of those 3x3 cells is kept by the sequential walk too.
// What differs is what a round costs.  The two forms in fd_select.cu touch every live candidate (or every cell) in every round,
// and FAST makes rounds plentiful: its responses are score + a running offset, so along an edge of equal scores the candidates
// are ranked in raster order and only one of them per (d+1)-stride can be decided per round -- 20 to 30 rounds on the synthetic
// frames.  Here the candidates are grouped by cell once and a round only touches what the previous round changed:
//   K  the cells on the evaluation list compare their best live candidate with the best of the 8 cells around them (32-bit compare of
//      the response words, the position words only on a tie); a winner is kept, and the live cells around it go on the rescan list;
//   R  eight lanes per listed cell drop the candidates the fresh points cover and find the cell's new best; the cell and the live
//      cells around it go on the next evaluation list.
// Two barriers per round, work proportional to the number of kept points, and no 64-bit shared-memory atomics (the 64-bit atomicMin
// of fd_select.cu is a compare-and-swap loop in SASS).  With more than SELECT_PREFIX_MIN candidates the rounds run on rank ranges
// exactly as in fd_select.cu.  One CTA per frame; frames whose cell grid does not fit shared memory stay with fd_select.cu.
#include "fd_select_common.cuh"

namespace fdb {

namespace {

constexpr uint32_t kDeadHi = 0xFFFFFFFFu;   // no response maps to this high word (it would be the NaN 0xFFFFFFFF)

struct LeanCells {
    uint32_t *cstart;    // n_cells + 1: after the scatter the candidates of cell c are binned[cstart[c - 1] .. cstart[c])
    uint32_t *best_hi;   // best live candidate of the cell: high (response) and low (position) key words; kDeadHi = none
    uint32_t *best_lo;
    uint32_t *kept_xy;   // the point kept in the cell, (row << 16) | col, or kEmptyCell
    uint32_t *fresh;     // round in which that point was kept
    uint32_t *mark;      // last list stamp of the cell (one entry per cell per list and round)
    uint16_t *klist;     // cells to evaluate
    uint16_t *rlist;     // cells to rescan
};

__global__ void __launch_bounds__(SELECT_LEAN_MAX_THREADS) select_lean_kernel(const SelectArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int frame = blockIdx.x;
    const int d = p.min_distance;
    const int pitch = p.cells_x + 2;
    const int n_cells = pitch * (p.cells_y + 2);
    LeanCells cs;
    cs.cstart = reinterpret_cast<uint32_t *>(smem);
    cs.best_hi = cs.cstart + n_cells + 1;
    cs.best_lo = cs.best_hi + n_cells;
    cs.kept_xy = cs.best_lo + n_cells;
    cs.fresh = cs.kept_xy + n_cells;
    cs.mark = cs.fresh + n_cells;
    cs.klist = reinterpret_cast<uint16_t *>(cs.mark + n_cells);
    cs.rlist = cs.klist + n_cells;
    __shared__ uint32_t s_kept, s_nk[2], s_nr[2], s_admit, s_warp[32];   // list lengths are double-buffered by pass parity
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
        return (cy + 1) * pitch + cx + 1;
    };

    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) {
        cs.kept_xy[i] = kEmptyCell;
        cs.fresh[i] = 0u;
        cs.mark[i] = 0u;
        cs.best_hi[i] = kDeadHi;
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

    const int sub = lane_id() & 7, group = int(threadIdx.x >> 3), n_groups = int(blockDim.x >> 3);
    const uint32_t group_mask = 0xFFu << (lane_id() & 24);
    uint64_t lower = 0ull;   // keys below this were admitted by earlier batches
    uint32_t round = 0u;     // list stamps: 2 * round + 1 (rescan list), 2 * round + 2 (evaluation list)
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
                    if (lane_id() >= o) incl += v;
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

        // ---- admit and group by cell: count, scan, scatter ----
        for (int i = threadIdx.x; i <= n_cells; i += blockDim.x) cs.cstart[i] = 0u;
        if (threadIdx.x == 0) s_nk[0] = s_nk[1] = s_nr[0] = s_nr[1] = 0u;
        __syncthreads();
        // a candidate of this rank range is admitted unless it sits on a masked-out pixel (feature_point_detector.cpp:66) or a point
        // kept by a better-ranked batch covers it
        auto admit = [&](uint64_t key, int &c) {
            if (key < lower || key >= limit) return false;
            const uint32_t xy = uint32_t(key) ^ xy_xor;
            const int x = int(xy & 0xFFFFu), y = int(xy >> 16);
            c = cell_of(xy);
            if (mb != nullptr && !((mb[int64_t(y) * p.mask.words_per_row + (x >> 5)] >> (x & 31)) & 1u)) return false;
            if (batch > 0 && near_kept(cs.kept_xy, pitch, c, x, y, d)) return false;
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
                if (admit(k4[u], c)) atomicAdd(cs.cstart + c, 1u);
            }
        }
        __syncthreads();
        {
            const int n_tot = n_cells + 1;
            const int per = (n_tot + int(blockDim.x) - 1) / int(blockDim.x);
            const int b0 = min(int(threadIdx.x) * per, n_tot), b1 = min(b0 + per, n_tot);
            uint32_t sum = 0u;
            for (int b = b0; b < b1; ++b) sum += cs.cstart[b];
            uint32_t run = block_exclusive_scan(sum, s_warp);
            for (int b = b0; b < b1; ++b) {
                const uint32_t v = cs.cstart[b];
                cs.cstart[b] = run;
                run += v;
                // the first rescan list: every cell that received a candidate (its best is not known yet)
                if (v != 0u) cs.rlist[atomicAdd(&s_nr[(round + 1u) & 1u], 1u)] = uint16_t(b);
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
                if (admit(k4[u], c)) binned[atomicAdd(cs.cstart + c, 1u)] = k4[u];
            }
        }
        __syncthreads();

        // ---- rounds ----
        for (bool first = true;; first = false) {
            ++round;
            const uint32_t par = round & 1u;
            // the list lengths of a pass live in slot `par`; the slot the coming K phase appends to was last read a pass ago
            if (threadIdx.x == 0) s_nr[par ^ 1u] = 0u;
            // R: eight lanes per listed cell drop what the points kept in the previous K phase cover and find the cell's new best
            const int n_r = int(s_nr[par]);
            const uint32_t r_fresh = round - 1u;   // stamp of those points (none in a batch's first pass)
            for (int w = group; w < n_r; w += n_groups) {
                const int c = int(cs.rlist[w]);
                uint64_t best = kDeadKey;
                if (cs.kept_xy[c] == kEmptyCell) {   // a cell that holds a kept point is finished: the point covers the whole cell
                    uint32_t lo0 = 0xFFFFFFFFu, hi0 = 0xFFFFFFFFu, lo1 = 0xFFFFFFFFu, hi1 = 0xFFFFFFFFu, lo2 = 0xFFFFFFFFu, hi2 = 0xFFFFFFFFu;
                    int n_fresh = 0;
                    if (!first) {
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            const int nb = c + (k / 3 - 1) * pitch + (k % 3 - 1);
                            if (cs.fresh[nb] == r_fresh && cs.kept_xy[nb] != kEmptyCell) {
                                const uint32_t q = cs.kept_xy[nb];
                                const int qx = int(q & 0xFFFFu), qy = int(q >> 16);
                                const uint32_t lo = (uint32_t(max(qy - d, 0)) << 16) | uint32_t(max(qx - d, 0));
                                const uint32_t hi = (uint32_t(min(qy + d, 65534)) << 16) | uint32_t(min(qx + d, 65534));
                                if (n_fresh == 0) lo0 = lo, hi0 = hi;
                                else if (n_fresh == 1) lo1 = lo, hi1 = hi;
                                else if (n_fresh == 2) lo2 = lo, hi2 = hi;
                                ++n_fresh;
                            }
                        }
                    }
                    const uint32_t js = cs.cstart[c - 1] + sub, je = cs.cstart[c];
                    if (n_fresh <= 3) {
                        for (uint32_t j = js; j < je; j += 32) {   // four candidates per lane per trip: their loads are in flight together
                            uint64_t key[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) key[u] = (j + 8 * u < je) ? binned[j + 8 * u] : kDeadKey;
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint32_t xy = uint32_t(key[u]) ^ xy_xor;
                                const bool hit = key[u] != kDeadKey && ((__vminu2(__vmaxu2(xy, lo0), hi0) == xy) | (__vminu2(__vmaxu2(xy, lo1), hi1) == xy) |
                                                                        (__vminu2(__vmaxu2(xy, lo2), hi2) == xy));
                                if (hit) binned[j + 8 * u] = kDeadKey;
                                else best = min(best, key[u]);
                            }
                        }
                    } else {
                        for (uint32_t j = js; j < je; j += 8) {
                            const uint64_t key = binned[j];
                            const uint32_t xy = uint32_t(key) ^ xy_xor;
                            if (key != kDeadKey && near_kept(cs.kept_xy, pitch, c, int(xy & 0xFFFFu), int(xy >> 16), d)) binned[j] = kDeadKey;
                            else best = min(best, key);
                        }
                    }
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) best = min(best, __shfl_xor_sync(group_mask, best, o));
                }
                if (sub == 0) {
                    cs.best_hi[c] = uint32_t(best >> 32);
                    cs.best_lo[c] = uint32_t(best);
                }
                // the cell itself (if still alive) and the live cells around it may win now; a stale "alive" of a cell that is being
                // rescanned only costs an evaluation, and a dead cell stays dead
                const uint32_t stamp = 2u * round + 2u;
                for (int k = sub; k < 9; k += 8) {
                    const int nb = c + (k / 3 - 1) * pitch + (k % 3 - 1);
                    const bool alive = (nb == c) ? (best != kDeadKey) : (cs.best_hi[nb] != kDeadHi);
                    if (alive && atomicMax(cs.mark + nb, stamp) < stamp) cs.klist[atomicAdd(&s_nk[par], 1u)] = uint16_t(nb);
                }
            }
            __syncthreads();
            const int n_k = int(s_nk[par]);
            if (n_k == 0) break;
            if (threadIdx.x == 0) s_nk[par ^ 1u] = 0u;   // the next pass's R phase appends to it, after the barrier below
            // K: a listed cell whose best live candidate outranks the best of the 8 cells around it keeps it
            for (int w = threadIdx.x; w < n_k; w += blockDim.x) {
                const int c = int(cs.klist[w]);
                const uint32_t hi = cs.best_hi[c];
                if (hi == kDeadHi) continue;
                const uint32_t lo = cs.best_lo[c];
                uint32_t nb_hi = kDeadHi;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (k != 4) nb_hi = min(nb_hi, cs.best_hi[c + (k / 3 - 1) * pitch + (k % 3 - 1)]);
                bool win = hi < nb_hi;
                if (hi == nb_hi) {   // equal responses: the position words decide
                    win = true;
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int nb = c + (k / 3 - 1) * pitch + (k % 3 - 1);
                        if (k != 4 && cs.best_hi[nb] == hi && cs.best_lo[nb] < lo) win = false;
                    }
                }
                if (!win) continue;
                const uint32_t slot = atomicAdd(&s_kept, 1u);
                if (slot < uint32_t(p.kept_capacity)) kept[slot] = (uint64_t(hi) << 32) | lo;
                cs.kept_xy[c] = lo ^ xy_xor;
                cs.fresh[c] = round;
                const uint32_t stamp = 2u * round + 3u;   // = 2 * (round + 1) + 1: the rescan list of the next pass
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int nb = c + (k / 3 - 1) * pitch + (k % 3 - 1);
                    if ((k == 4 || cs.best_hi[nb] != kDeadHi) && atomicMax(cs.mark + nb, stamp) < stamp) cs.rlist[atomicAdd(&s_nr[par ^ 1u], 1u)] = uint16_t(nb);
                }
            }
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
        } else
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

size_t select_lean_smem_bytes(int cells_x, int cells_y) {   // six words and two list entries per cell, one extra list start
    const size_t n = size_t(cells_x + 2) * (cells_y + 2);
    return (n * 28 + 4 + 15) & ~size_t(15);
}

cudaError_t launch_select_lean(const SelectArgs &args, int threads, cudaStream_t stream) {
    const size_t smem = select_lean_smem_bytes(args.cells_x, args.cells_y);
    cudaError_t e = cudaFuncSetAttribute(select_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    select_lean_kernel<<<args.n_frames, threads, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
