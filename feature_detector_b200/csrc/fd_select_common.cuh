// Device helpers of the selection kernel (fd_select.cu).
#pragma once

#include "fd_kernels.cuh"

namespace fdb {
namespace {

constexpr uint32_t kEmptyCell = 0xFFFFFFFFu;
constexpr uint64_t kDeadKey = 0xFFFFFFFFFFFFFFFFull;

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

// Unordered append of this thread's surviving candidate key to a list (warp-aggregated counter bump).  Whole warps only: every
// caller rounds its loop bounds to warps, which spares the active-mask vote and the leader election of the general form.
__device__ __forceinline__ uint32_t lane_mask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ void list_push(bool keep, uint64_t value, uint64_t *list, uint32_t *counter) {
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m == 0u) return;
    uint32_t base = 0u;
    if (lane_id() == 0) base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) list[base + __popc(m & lane_mask_lt())] = value;
}

// Exclusive prefix sum of one value per thread over the whole CTA (up to 1024 threads).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane_id() >= o) incl += u;
    }
    if (lane_id() == 31) warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t w = (threadIdx.x < (blockDim.x >> 5)) ? warp_sums[threadIdx.x] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane_id() >= o) wi += u;
        }
        warp_sums[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const uint32_t r = warp_sums[threadIdx.x >> 5] + incl - v;
    __syncthreads();   // warp_sums is free again
    return r;
}

constexpr int SELECT_HIST_BITS = 11;   // rank-prefix histogram over the top key bits: sign, exponent, two mantissa bits

// The first histogram bin at which the running count reaches prefix_k: the batch admits the keys below the NEXT bin's first key.
// One warp (BINS / 32 bins per lane); the lane that finds the bin writes the limit and the count of keys below it.
__device__ __forceinline__ void warp_prefix_limit(const uint32_t *hist, uint32_t prefix_k, uint32_t n, uint64_t *out_limit, uint32_t *out_admit) {
    constexpr int BINS = 1 << SELECT_HIST_BITS;
    constexpr int PER = BINS / 32;
    const int lane = lane_id();
    uint32_t mine = 0u;
#pragma unroll 4
    for (int b = 0; b < PER; ++b) mine += hist[lane * PER + b];
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
            run += hist[lane * PER + b];
            if (run >= prefix_k) break;
        }
        const uint32_t bin = uint32_t(lane * PER + b);
        *out_limit = (bin >= uint32_t(BINS - 1)) ? kDeadKey : (uint64_t(bin + 1u) << (64 - SELECT_HIST_BITS));
        *out_admit = (bin >= uint32_t(BINS - 1)) ? n : run;
    }
}

// How many candidates a frame still wants and how large its first rank range is (shared by the selection kernel and the kernels that
// prepare its first range).
__device__ __forceinline__ uint32_t select_want(const SelectArgs &p, int frame) {
    const uint32_t n_pre = p.existing_counts ? uint32_t(p.existing_counts[frame]) : 0u;
    const uint32_t want = (p.needed > n_pre) ? (p.needed - n_pre) : 1u;  // pushed, then tested: at least one
    return min(want, uint32_t(p.kp_capacity));
}
__device__ __forceinline__ uint32_t select_first_range(uint32_t want_kept) { return max(uint32_t(SELECT_PREFIX_FIRST), 8u * want_kept); }


}  // namespace
}  // namespace fdb
