// Shared device/host helpers for the sm_100a kernels behind include/fd_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fdb {

// ---- frame view --------------------------------------------------------------------------------
// Device-resident batch of 8-bit frames.  `data` is 4-byte aligned and `pitch`, `frame_stride` are
// multiples of 4, so every row can be read as aligned 32-bit words (the context re-pitches anything
// that is not).  Words past `cols` inside the pitch may hold garbage: they only ever feed pixels
// outside the image, whose results are discarded.
struct FrameView {
    const uint8_t *data;
    int rows, cols;
    int64_t pitch;         // bytes between rows
    int64_t frame_stride;  // bytes between frames
    int n_frames;
    int words_per_row;     // readable aligned words per row = pitch / 4
};

// ---- row tiling of one large frame -------------------------------------------------
// The bound buffer holds rows [row_offset, row_offset + fv.rows) of an image that is full_rows tall.  Candidates are
// produced only for the local rows [own_lo, own_hi) (the rest is halo) and carry ABSOLUTE row numbers; FAST's running
// offset is indexed by the absolute pixel position, so tiles are seam-free.  Untiled: {0, 0, rows, rows}.
struct TileView {
    int row_offset, own_lo, own_hi, full_rows;
};

// ---- candidate keys ------------------------------------------------------------------------------
// One candidate = one 64-bit key: high word = bitwise complement of the order-preserving integer image
// of the float response, low word = (row << 16) | col.  Sorting keys ASCENDING therefore yields response
// DESCENDING with ties in raster order -- the tie rule this framework fixes where the reference's unstable
// std::sort leaves it open (feature_point_detector.cpp:58).  Frames are limited to 65535 x 65535.
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_cand_key(float response, uint32_t row, uint32_t col) {
    return (uint64_t(~float_to_ordered(response)) << 32) | (row << 16) | col;
}
__host__ __device__ __forceinline__ float cand_key_response(uint64_t key) { return ordered_to_float(~uint32_t(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t cand_key_xy(uint64_t key) { return uint32_t(key); }  // (row << 16) | col

#ifdef __CUDACC__
// ---- small PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t ld_word(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Work items of the candidate kernels (frame x band x strip, one per warp at a time).  A warp's first item is its own --
// interleaved over the CTAs, so that even a single frame's few hundred items land on every SM and no warp starts by queueing on
// one L2 atomic -- and the rest come from a shared counter (zeroed by the host before the launch) rather than a fixed stride:
// an item's cost follows its content, and a fixed assignment leaves a visible tail.
__device__ __forceinline__ int64_t next_work_item(uint32_t *counter, bool first) {
    const uint32_t warps_per_cta = blockDim.x >> 5;
    if (first) return int64_t(blockIdx.x) + int64_t(gridDim.x) * (threadIdx.x >> 5);
    uint32_t next = 0u;
    if (lane_id() == 0) next = atomicAdd(counter, 1u);
    return int64_t(gridDim.x) * warps_per_cta + int64_t(__shfl_sync(0xffffffffu, next, 0));
}

// Bits of mask word w (columns 32w .. 32w+31) that are FAST-interior columns [3, cols-4].
__device__ __forceinline__ uint32_t fast_interior_bits(int w, int cols) {
    const int lo = max(3 - 32 * w, 0), hi = min(cols - 4 - 32 * w, 31);
    if (lo > 31 || hi < 0 || lo > hi) return 0u;
    const uint32_t upto_hi = (hi == 31) ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u);
    return upto_hi & ~((1u << lo) - 1u);
}

// Warp-aggregated append of up to `n_mine` keys per lane into a per-frame slot.  Returns false if the
// slot overflowed (the counter still advances, so the host sees the true demand).
__device__ __forceinline__ uint32_t warp_reserve(uint32_t *counter, uint32_t n_mine) {
    // inclusive scan of n_mine across the warp
    uint32_t incl = n_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane_id() >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (total != 0) {
        if (lane_id() == 31) base = atomicAdd(counter, total);
        base = __shfl_sync(0xffffffffu, base, 31);
    }
    return base + incl - n_mine;
}
// Ascending bitonic network over keys[0, n) (n need not be a power of two): ascending-only comparators, so the
// tail beyond n acts as +inf padding.
__device__ inline void block_bitonic_sort(uint64_t *keys, uint32_t n) {
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

#endif  // __CUDACC__

}  // namespace fdb
