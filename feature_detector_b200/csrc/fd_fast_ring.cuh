// Exact FAST-16 segment test on four adjacent pixels per lane (shared by the dense and the sparse FAST kernels).
// Reference: src/feature_point_detector/feature_point_fast_detector.cpp:11-81.  See fd_fast.cu for the design notes.
#pragma once

#include <cuda_fp16.h>

#include "fd_kernels.cuh"

namespace fdb {
namespace fastring {

// One image row as this lane sees it: packed bytes of columns x-4..x-1 (w0), x..x+3 (w1), x+4..x+7 (w2) and
// the two words that straddle them, xl = columns x-1..x+2 and xr = columns x+3..x+6.
struct Row {
    uint32_t w0, w1, w2, xl, xr;
};

__device__ __forceinline__ void make_row(Row &r, uint32_t w0, uint32_t w1, uint32_t w2) {
    r.w0 = w0;
    r.w1 = w1;
    r.w2 = w2;
    r.xl = __funnelshift_r(w0, w1, 24);
    r.xr = __funnelshift_r(w1, w2, 24);
}

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }

// Pixels (x+S, x+S+1) of row `r` as half2 (1024 + value each).
template <int S>
__device__ __forceinline__ __half2 pair(const Row &r) {
    constexpr uint32_t M = 0x64646464u;
    if (S == -3) return h2(prmt(r.w0, M, 0x4241u));
    if (S == -2) return h2(prmt(r.w0, M, 0x4342u));
    if (S == -1) return h2(prmt(r.xl, M, 0x4140u));
    if (S == 0) return h2(prmt(r.w1, M, 0x4140u));
    if (S == 1) return h2(prmt(r.w1, M, 0x4241u));
    if (S == 2) return h2(prmt(r.w1, M, 0x4342u));
    if (S == 3) return h2(prmt(r.xr, M, 0x4140u));
    if (S == 4) return h2(prmt(r.w2, M, 0x4140u));
    return h2(prmt(r.w2, M, 0x4241u));  // S == 5
}

// Ring masks as fp16 accumulators: value 1024 + (8-bit mask); index 0 = pixels (0,1), 1 = pixels (2,3).
struct Acc {
    __half2 bLo[2], bHi[2], dLo[2], dHi[2];
};

template <int I, int DX>
__device__ __forceinline__ void ring_step(const Row &r, const __half2 (&nhi)[2], const __half2 (&lo)[2], Acc &a) {
    constexpr uint32_t wbits = (0x3C00u + (uint32_t(I & 7) << 10)) * 0x00010001u;  // half2(2^(I&7), 2^(I&7))
    const __half2 w = h2(wbits);
    const __half2 r01 = pair<DX>(r), r23 = pair<DX + 2>(r);
    const __half2 fb0 = __hadd2_sat(r01, nhi[0]), fb1 = __hadd2_sat(r23, nhi[1]);      // ring > centre + diff
    const __half2 fd0 = __hadd2_sat(lo[0], __hneg2(r01)), fd1 = __hadd2_sat(lo[1], __hneg2(r23));  // ring < centre - diff
    if (I < 8) {
        a.bLo[0] = __hfma2(fb0, w, a.bLo[0]);
        a.bLo[1] = __hfma2(fb1, w, a.bLo[1]);
        a.dLo[0] = __hfma2(fd0, w, a.dLo[0]);
        a.dLo[1] = __hfma2(fd1, w, a.dLo[1]);
    } else {
        a.bHi[0] = __hfma2(fb0, w, a.bHi[0]);
        a.bHi[1] = __hfma2(fb1, w, a.bHi[1]);
        a.dHi[0] = __hfma2(fd0, w, a.dHi[0]);
        a.dHi[1] = __hfma2(fd1, w, a.dHi[1]);
    }
}

template <bool PRECHECK>
__device__ __forceinline__ void fast_step(const Row &rm3, const Row &rm2, const Row &rm1, const Row &r0, const Row &rp1, const Row &rp2,
                                          const Row &rp3, __half2 diff2, const uint8_t *__restrict__ lut, int prune, uint32_t &scores_packed) {
    // centre thresholds: nhi = -(centre + diff), lo = centre - diff (the +1024 of every operand cancels)
    const __half2 c01 = pair<0>(r0), c23 = pair<2>(r0);
    const __half2 nhi[2] = {__hneg2(__hadd2(c01, diff2)), __hneg2(__hadd2(c23, diff2))};
    const __half2 lo[2] = {__hsub2(c01, diff2), __hsub2(c23, diff2)};
    const __half2 k1024 = h2(0x64006400u);
    Acc a = {{k1024, k1024}, {k1024, k1024}, {k1024, k1024}, {k1024, k1024}};
    // ring index: {dx, dy} per fast.cpp:7-8 -- 0 top, clockwise
    ring_step<4, 3>(r0, nhi, lo, a);
    ring_step<8, 0>(rp3, nhi, lo, a);
    ring_step<12, -3>(r0, nhi, lo, a);
    uint32_t pass = 0xFFFFFFFFu;
    if (PRECHECK) {
        // closed form of fast.cpp:20-42: right (bit 4), bottom (bit 8), left (bit 12) all brighter or all darker
        uint32_t m[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t bl = u32(a.bLo[q]), bh = u32(a.bHi[q]), dl = u32(a.dLo[q]), dh = u32(a.dHi[q]);
            const uint32_t pb = (bl >> 4) & bh & (bh >> 4) & 0x00010001u;
            const uint32_t pd = (dl >> 4) & dh & (dh >> 4) & 0x00010001u;
            m[q] = (pb | pd) * 0xFFu;  // 0x00FF per passing pixel of the pair
        }
        pass = prmt(m[0], m[1], 0x6420u);
        if (!__any_sync(0xffffffffu, pass != 0u)) {
            scores_packed = 0u;
            return;
        }
    }
    ring_step<0, 0>(rm3, nhi, lo, a);
    if (prune != 0) {
        // Threshold-aware pruning (exact): only scores >= s_min can become candidates in this row group, and a
        // circular run of s_min ring pixels contains floor(s_min / 4) CONSECUTIVE compass positions (0, 4, 8, 12)
        // of the same polarity.  prune = 1: s_min in 4..7 (one compass flag); prune = 2: s_min >= 8 (two adjacent).
        // A warp whose 128 pixels all fail cannot hold a candidate and skips the other 12 ring positions.
        uint32_t ok = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            // compass flags per pixel: bit 0 = pos 0, bit 4 = pos 4, bit 1 = pos 8, bit 5 = pos 12
            const uint32_t xb = (u32(a.bLo[q]) & 0x00110011u) | ((u32(a.bHi[q]) & 0x00110011u) << 1);
            const uint32_t xd = (u32(a.dLo[q]) & 0x00110011u) | ((u32(a.dHi[q]) & 0x00110011u) << 1);
            if (prune == 1) {
                ok |= xb | xd;
            } else {
                ok |= (xb & (xb >> 4)) | (xb & (xb >> 3) & 0x00020002u) | (xb & (xb >> 5) & 0x00010001u);
                ok |= (xd & (xd >> 4)) | (xd & (xd >> 3) & 0x00020002u) | (xd & (xd >> 5) & 0x00010001u);
            }
        }
        if (!__any_sync(0xffffffffu, ok != 0u)) {
            scores_packed = 0u;
            return;
        }
    }
    ring_step<1, 1>(rm3, nhi, lo, a);
    ring_step<2, 2>(rm2, nhi, lo, a);
    ring_step<3, 3>(rm1, nhi, lo, a);
    ring_step<5, 3>(rp1, nhi, lo, a);
    ring_step<6, 2>(rp2, nhi, lo, a);
    ring_step<7, 1>(rp3, nhi, lo, a);
    ring_step<9, -1>(rp3, nhi, lo, a);
    ring_step<10, -2>(rp2, nhi, lo, a);
    ring_step<11, -3>(rp1, nhi, lo, a);
    ring_step<13, -3>(rm1, nhi, lo, a);
    ring_step<14, -2>(rm2, nhi, lo, a);
    ring_step<15, -1>(rm3, nhi, lo, a);

    uint32_t sp = 0u;
    const uint32_t any = u32(a.bLo[0]) | u32(a.bLo[1]) | u32(a.bHi[0]) | u32(a.bHi[1]) | u32(a.dLo[0]) | u32(a.dLo[1]) | u32(a.dHi[0]) | u32(a.dHi[1]);
    if (any != 0x64006400u) {
        // per pixel: 16-bit masks -> longest circular run (fast.cpp:55-78), best of both polarities
        const uint32_t b0 = prmt(u32(a.bLo[0]), u32(a.bHi[0]), 0x7740u) & 0xFFFFu, d0 = prmt(u32(a.dLo[0]), u32(a.dHi[0]), 0x7740u) & 0xFFFFu;
        const uint32_t b1 = prmt(u32(a.bLo[0]), u32(a.bHi[0]), 0x7762u) & 0xFFFFu, d1 = prmt(u32(a.dLo[0]), u32(a.dHi[0]), 0x7762u) & 0xFFFFu;
        const uint32_t b2 = prmt(u32(a.bLo[1]), u32(a.bHi[1]), 0x7740u) & 0xFFFFu, d2 = prmt(u32(a.dLo[1]), u32(a.dHi[1]), 0x7740u) & 0xFFFFu;
        const uint32_t b3 = prmt(u32(a.bLo[1]), u32(a.bHi[1]), 0x7762u) & 0xFFFFu, d3 = prmt(u32(a.dLo[1]), u32(a.dHi[1]), 0x7762u) & 0xFFFFu;
        const uint32_t s0 = max((uint32_t)lut[b0], (uint32_t)lut[d0]);
        const uint32_t s1 = max((uint32_t)lut[b1], (uint32_t)lut[d1]);
        const uint32_t s2 = max((uint32_t)lut[b2], (uint32_t)lut[d2]);
        const uint32_t s3 = max((uint32_t)lut[b3], (uint32_t)lut[d3]);
        sp = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
    scores_packed = sp & pass;
}

}  // namespace fastring
}  // namespace fdb
