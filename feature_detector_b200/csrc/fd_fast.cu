// Kernel 2 -- FAST ring test, score and running-offset response.
// Replaces FeaturePointFastDetector::ComputeResponseOfPixel / ComputeCandidates
// (reference src/feature_point_detector/feature_point_fast_detector.cpp:11-81, 83-98).
//
// Design (sm_100a, no tensor cores: nothing here is a contraction)
//   * a warp owns a 128-pixel-wide column strip of one frame and streams down a band of rows; each lane
//     owns 4 adjacent pixels (one aligned 32-bit word per row) and keeps the 7 rows the Bresenham ring
//     spans in registers, pre-split into 16-bit lanes (even / odd pixels) so that one 32-bit add compares
//     two pixels against the centre +/- threshold: bit 15 of (ring + 0x8000 - diff - 1 - centre) is the
//     "brighter" flag, bit 15 of (0x8000 - diff - 1 + centre - ring) the "darker" flag.  No saturation
//     is needed: the 16-bit lanes give the int32 semantics of fast.cpp:12-14 exactly;
//   * one PRMT with sign replication turns the two flag words (even, odd) into four byte masks, one
//     LOP3 files them under ring bit i: after 16 ring positions each pixel has a 16-bit brighter mask and
//     a 16-bit darker mask;
//   * the score (longest circular run of set bits, 0..16; fast.cpp:55-78) is a 64 KiB shared-memory
//     table lookup per mask, skipped for lanes whose masks are all zero;
//   * the kN >= 12 pre-check (fast.cpp:20-42) is evaluated in its closed form -- right, bottom and left
//     ring pixels all brighter, or all darker -- on ring positions 4, 8, 12 first, and a
//     warp whose 128 pixels all fail skips the other 13 positions;
//   * response = float(score) + offset(k) with the reference's sequentially accumulated float offset
//     reproduced exactly from a piecewise-linear table of its bit pattern;
//   * candidates (response > threshold) are appended to the frame's slot with one atomic per warp.
// HBM traffic: every input byte is read once from DRAM (neighbouring lanes / row bands re-read it from
// L1 / L2); output is the candidate list (+1 B/px when the dense score map is requested).
#include "fd_kernels.cuh"

namespace fdb {

namespace {

struct Row {
    uint32_t E1, O1, O0, E2, SE01, SO01, SE12, SO12;
};

__device__ __forceinline__ void make_row(Row &r, uint32_t w0, uint32_t w1, uint32_t w2) {
    // even pixels -> 16-bit lanes (px0, px2); odd pixels -> (px1, px3)
    const uint32_t e0 = prmt(w0, 0u, 0x4240u);
    r.O0 = prmt(w0, 0u, 0x4341u);
    r.E1 = prmt(w1, 0u, 0x4240u);
    r.O1 = prmt(w1, 0u, 0x4341u);
    r.E2 = prmt(w2, 0u, 0x4240u);
    const uint32_t o2 = prmt(w2, 0u, 0x4341u);
    r.SE01 = __funnelshift_r(e0, r.E1, 16);    // pixels (x-2, x)
    r.SO01 = __funnelshift_r(r.O0, r.O1, 16);  // pixels (x-1, x+1)
    r.SE12 = __funnelshift_r(r.E1, r.E2, 16);  // pixels (x+2, x+4)
    r.SO12 = __funnelshift_r(r.O1, o2, 16);    // pixels (x+3, x+5)
}

// (even-lane word, odd-lane word) of the four pixels at column offset DX in row `r`.
template <int DX>
__device__ __forceinline__ void ring_words(const Row &r, uint32_t &e, uint32_t &o) {
    if (DX == 0) { e = r.E1; o = r.O1; }
    else if (DX == 1) { e = r.O1; o = r.SE12; }
    else if (DX == -1) { e = r.SO01; o = r.E1; }
    else if (DX == 2) { e = r.SE12; o = r.SO12; }
    else if (DX == -2) { e = r.SE01; o = r.SO01; }
    else if (DX == 3) { e = r.SO12; o = r.E2; }
    else { e = r.O0; o = r.SE01; }  // DX == -3
}

struct Acc {
    uint32_t bLo, bHi, dLo, dHi;  // per pixel byte: ring bits 0-7 / 8-15 of the brighter / darker mask
};

template <int I, int DX>
__device__ __forceinline__ void ring_step(const Row &r, uint32_t kbE, uint32_t kbO, uint32_t kdE, uint32_t kdO, Acc &a) {
    uint32_t e, o;
    ring_words<DX>(r, e, o);
    const uint32_t mb = prmt(e + kbE, o + kbO, 0xFBD9u);  // 0xFF per pixel whose ring pixel I is brighter
    const uint32_t md = prmt(kdE - e, kdO - o, 0xFBD9u);  // ... darker
    constexpr uint32_t bit = 0x01010101u << (I & 7);
    if (I < 8) {
        a.bLo |= mb & bit;
        a.dLo |= md & bit;
    } else {
        a.bHi |= mb & bit;
        a.dHi |= md & bit;
    }
}

// Offset table lookup: bit pattern of the reference's running float `offset` for masked-in pixel index k
// (binary search over the <= 64 linear pieces; only candidate-bearing rows get here).
__device__ __forceinline__ uint32_t offset_bits(const OffsetSeg *__restrict__ segs, int n_seg, uint32_t k) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].k_start <= k) lo = mid; else hi = mid - 1;
    }
    return segs[lo].bits_start + (k - segs[lo].k_start) * segs[lo].step;
}

template <bool PRECHECK>
__device__ __forceinline__ void fast_step(const Row &rm3, const Row &rm2, const Row &rm1, const Row &r0, const Row &rp1, const Row &rp2,
                                          const Row &rp3, uint32_t kbias, const uint8_t *__restrict__ lut, uint32_t &scores_packed) {
    // centre thresholds, per 16-bit lane: 0x8000 - (diff + 1) -/+ centre
    const uint32_t kbE = kbias - r0.E1, kbO = kbias - r0.O1;
    const uint32_t kdE = kbias + r0.E1, kdO = kbias + r0.O1;
    Acc a = {0u, 0u, 0u, 0u};
    // ring index: {dx, dy} per fast.cpp:7-8 -- 0 top, clockwise
    ring_step<4, 3>(r0, kbE, kbO, kdE, kdO, a);
    ring_step<8, 0>(rp3, kbE, kbO, kdE, kdO, a);
    ring_step<12, -3>(r0, kbE, kbO, kdE, kdO, a);
    uint32_t pass = 0xFFFFFFFFu;
    if (PRECHECK) {
        // closed form of fast.cpp:20-42: right, bottom, left all brighter or all darker
        const uint32_t pb = (a.bLo >> 4) & a.bHi & (a.bHi >> 4) & 0x01010101u;
        const uint32_t pd = (a.dLo >> 4) & a.dHi & (a.dHi >> 4) & 0x01010101u;
        pass = (pb | pd) * 0xFFu;  // 0xFF per passing pixel
        if (!__any_sync(0xffffffffu, pass != 0u)) {
            scores_packed = 0u;
            return;
        }
    }
    ring_step<0, 0>(rm3, kbE, kbO, kdE, kdO, a);
    ring_step<1, 1>(rm3, kbE, kbO, kdE, kdO, a);
    ring_step<2, 2>(rm2, kbE, kbO, kdE, kdO, a);
    ring_step<3, 3>(rm1, kbE, kbO, kdE, kdO, a);
    ring_step<5, 3>(rp1, kbE, kbO, kdE, kdO, a);
    ring_step<6, 2>(rp2, kbE, kbO, kdE, kdO, a);
    ring_step<7, 1>(rp3, kbE, kbO, kdE, kdO, a);
    ring_step<9, -1>(rp3, kbE, kbO, kdE, kdO, a);
    ring_step<10, -2>(rp2, kbE, kbO, kdE, kdO, a);
    ring_step<11, -3>(rp1, kbE, kbO, kdE, kdO, a);
    ring_step<13, -3>(rm1, kbE, kbO, kdE, kdO, a);
    ring_step<14, -2>(rm2, kbE, kbO, kdE, kdO, a);
    ring_step<15, -1>(rm3, kbE, kbO, kdE, kdO, a);

    uint32_t sp = 0u;
    if ((a.bLo | a.bHi | a.dLo | a.dHi) != 0u) {
        // per pixel: 16-bit masks -> longest circular run (fast.cpp:55-78), best of both polarities
        const uint32_t b0 = prmt(a.bLo, a.bHi, 0x4440u) & 0xFFFFu, d0 = prmt(a.dLo, a.dHi, 0x4440u) & 0xFFFFu;
        const uint32_t b1 = prmt(a.bLo, a.bHi, 0x4451u) & 0xFFFFu, d1 = prmt(a.dLo, a.dHi, 0x4451u) & 0xFFFFu;
        const uint32_t b2 = prmt(a.bLo, a.bHi, 0x4462u) & 0xFFFFu, d2 = prmt(a.dLo, a.dHi, 0x4462u) & 0xFFFFu;
        const uint32_t b3 = prmt(a.bLo, a.bHi, 0x4473u) & 0xFFFFu, d3 = prmt(a.dLo, a.dHi, 0x4473u) & 0xFFFFu;
        const uint32_t s0 = max((uint32_t)lut[b0], (uint32_t)lut[d0]);
        const uint32_t s1 = max((uint32_t)lut[b1], (uint32_t)lut[d1]);
        const uint32_t s2 = max((uint32_t)lut[b2], (uint32_t)lut[d2]);
        const uint32_t s3 = max((uint32_t)lut[b3], (uint32_t)lut[d3]);
        sp = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
    }
    scores_packed = sp & pass;
}

template <bool PRECHECK>
__global__ void __launch_bounds__(FAST_THREADS, 2) fast_kernel(const FastArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *lut = smem;                                               // 65536 B: run-length table
    OffsetSeg *segs = reinterpret_cast<OffsetSeg *>(smem + 65536);     // p.n_seg + 1 entries
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(lut)[i] = __ldg(reinterpret_cast<const uint4 *>(p.lut) + i);
    for (int i = threadIdx.x; i <= p.n_seg; i += blockDim.x) segs[i] = p.segs[i];
    __syncthreads();

    const FrameView &fv = p.fv;
    const int lane = lane_id();
    const int warps_per_block = blockDim.x >> 5;
    const int64_t total_warps = int64_t(gridDim.x) * warps_per_block;
    const int64_t gwarp = int64_t(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
    const uint32_t kbias = (0x8000u - uint32_t(p.diff) - 1u) * 0x00010001u;
    const int inner_cols = fv.cols - 6;

    for (int64_t item = gwarp; item < p.n_items; item += total_warps) {
        // item -> (frame, band, strip); strips fastest so neighbouring warps share L1 lines
        const int strip = int(item % p.n_strips);
        const int64_t t = item / p.n_strips;
        const int band = int(t % p.n_bands);
        const int frame = int(t / p.n_bands);
        const int row_begin = 3 + band * p.band_rows;
        const int row_end = min(row_begin + p.band_rows, fv.rows - 3);
        if (row_begin >= row_end) continue;

        // Rows are read as three aligned words per lane.  Addresses are clamped instead of predicated: a
        // clamped word only ever feeds pixels outside the interior, whose results are masked below.
        const int w = strip * 32 + lane;                                // this lane's word in the row
        const int wl = min(max(w - 1, 0), fv.words_per_row - 1) * 4;
        const int wc = min(w, fv.words_per_row - 1) * 4;
        const int wr = min(w + 1, fv.words_per_row - 1) * 4;
        const uint8_t *fbase = fv.data + int64_t(frame) * fv.frame_stride;
        const int last_row = fv.rows - 1;
#define FD_LOAD_ROW(ROW, W0, W1, W2)                                          \
        {                                                                     \
            const uint8_t *rp__ = fbase + int64_t(min(ROW, last_row)) * fv.pitch; \
            W0 = ld_word(rp__ + wl);                                          \
            W1 = ld_word(rp__ + wc);                                          \
            W2 = ld_word(rp__ + wr);                                          \
        }

        Row rw[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            uint32_t w0, w1, w2;
            FD_LOAD_ROW(row_begin - 3 + i, w0, w1, w2);
            make_row(rw[i], w0, w1, w2);
        }

        const int col0 = 4 * w;
        uint32_t col_ok = 0u;  // 0xFF per interior column [3, cols-4] of this lane
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (col0 + j >= 3 && col0 + j <= fv.cols - 4) col_ok |= 0xFFu << (8 * j);

        // Smallest score that can become a candidate anywhere in this band: the offset only grows with k, so
        // fl(score + offset) <= fl(score + offset at the band's last pixel).  Rows whose scores all stay below it
        // skip the float path entirely (with the demo threshold 10 that is almost every row).
        uint32_t need_add;  // adding it to the packed scores sets bit 7 of every byte whose score >= s_need
        {
            const uint32_t k_hi = uint32_t(row_end - 1 - 3) * uint32_t(inner_cols) + uint32_t(inner_cols - 1);
            const float off_hi = __uint_as_float(offset_bits(segs, p.n_seg, k_hi));
            int s_need = 17;
            for (int sc = 16; sc >= 0; --sc)
                if (__fadd_rn(float(sc), off_hi) > p.thr) s_need = sc;
            need_add = (s_need == 0) ? 0x80808080u : (s_need > 16 ? 0u : (0x80u - uint32_t(s_need)) * 0x01010101u);
        }
        uint8_t *score_base = p.score_map ? p.score_map + int64_t(frame) * fv.rows * fv.cols : nullptr;

        for (int row = row_begin; row < row_end; row += 7) {
            uint32_t spv[7];
            uint32_t hits = 0u;  // warp-uniform: phases with at least one possible candidate
#pragma unroll
            for (int ph = 0; ph < 7; ++ph) {
                const int r = row + ph;
                uint32_t n0, n1, n2;
                FD_LOAD_ROW(r + 4, n0, n1, n2);  // the row that enters the window next step
                uint32_t sp;
                fast_step<PRECHECK>(rw[(ph + 0) % 7], rw[(ph + 1) % 7], rw[(ph + 2) % 7], rw[(ph + 3) % 7], rw[(ph + 4) % 7],
                                    rw[(ph + 5) % 7], rw[(ph + 6) % 7], kbias, lut, sp);
                sp &= col_ok;
                const bool live = r < row_end;  // rows past the band are computed on clamped data and dropped
                if (score_base != nullptr && live && col0 < fv.cols) {
                    uint8_t *dst = score_base + int64_t(r) * fv.cols + col0;
                    if (p.score_aligned && col0 + 3 < fv.cols) {
                        *reinterpret_cast<uint32_t *>(dst) = sp;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col0 + j < fv.cols) dst[j] = uint8_t(sp >> (8 * j));
                    }
                }
                spv[ph] = sp;
                const bool hit = live && ((((sp + need_add) | (need_add & 0x80808080u)) & col_ok & 0x80808080u) != 0u);
                if (__any_sync(0xffffffffu, hit)) hits |= 1u << ph;
                make_row(rw[(ph + 0) % 7], n0, n1, n2);
            }
            // response = score + offset(k), k = index of the pixel among the masked-in interior pixels in raster
            // order (fast.cpp:85-93); candidates are appended with one atomic per warp and row.
            while (hits != 0u) {
                const int ph = __ffs(hits) - 1;
                hits &= hits - 1u;
                uint32_t sp = spv[0];
#pragma unroll
                for (int q = 1; q < 7; ++q) sp = (ph == q) ? spv[q] : sp;
                const int r = row + ph;
                uint32_t mine = 0u;
                float resp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    resp[j] = 0.0f;
                    if ((col_ok >> (8 * j)) & 1u) {
                        const uint32_t k = uint32_t(r - 3) * uint32_t(inner_cols) + uint32_t(col0 + j - 3);
                        const float off = __uint_as_float(offset_bits(segs, p.n_seg, k));
                        const float v = __fadd_rn(float((sp >> (8 * j)) & 0xFFu), off);
                        if (v > p.thr) {
                            resp[j] = v;
                            mine |= 1u << j;
                        }
                    }
                }
                if (__any_sync(0xffffffffu, mine != 0u)) {
                    uint32_t pos = warp_reserve(p.cand_counts + frame, __popc(mine));
                    uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if ((mine >> j) & 1u) {
                            if (pos < p.cand_capacity) slot[pos] = make_cand_key(resp[j], uint32_t(r) * uint32_t(fv.cols) + uint32_t(col0 + j));
                            ++pos;
                        }
                    }
                }
            }
        }
#undef FD_LOAD_ROW
    }
}

}  // namespace

size_t fast_smem_bytes(int n_seg) { return 65536 + sizeof(OffsetSeg) * size_t(n_seg + 1); }

cudaError_t launch_fast(const FastArgs &args, bool precheck, int grid, cudaStream_t stream) {
    const size_t smem = fast_smem_bytes(args.n_seg);
    cudaError_t e;
    if (precheck) {
        e = cudaFuncSetAttribute(fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        fast_kernel<true><<<grid, FAST_THREADS, smem, stream>>>(args);
    } else {
        e = cudaFuncSetAttribute(fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        fast_kernel<false><<<grid, FAST_THREADS, smem, stream>>>(args);
    }
    return cudaGetLastError();
}

}  // namespace fdb
