// Kernel 2 -- FAST ring test, score and running-offset response.
// Replaces FeaturePointFastDetector::ComputeResponseOfPixel / ComputeCandidates
// (reference src/feature_point_detector/feature_point_fast_detector.cpp:11-81, 83-98).
//
// Design (sm_100a, no tensor cores: nothing here is a contraction)
//   * a warp owns a 128-pixel-wide column strip of one frame and streams down a band of rows; each lane
//     owns 4 adjacent pixels (one aligned 32-bit word per row) and keeps the 7 rows the Bresenham ring
//     spans in registers as packed bytes;
//   * the segment test runs on the FMA pipe in packed half precision (pixel values 0..255 and their
//     differences are exact in fp16): one PRMT turns two neighbouring bytes into a half2 (0x6400 | byte is
//     the fp16 number 1024 + byte, and the common 1024 cancels in every difference), one saturating HADD2
//     yields the brighter (or darker) flag of two pixels as exactly 0.0 / 1.0 -- the int32 comparisons of
//     fast.cpp:12-14,44-52 on integers -- and one HFMA2 files it under ring bit i of an accumulator that
//     starts at 1024.0, so the low byte of the accumulator's bit pattern IS the 8-bit ring mask.  The
//     integer pipe (half the FMA pipe's issue rate on this part) only does the byte shuffles;
//   * the score (longest circular run of set bits, 0..16; fast.cpp:55-78) is a 64 KiB shared-memory
//     table lookup per 16-bit mask, skipped for lanes whose masks are all zero;
//   * the kN >= 12 pre-check (fast.cpp:20-42) is evaluated in its closed form -- right, bottom and left
//     ring pixels all brighter, or all darker -- on ring positions 4, 8, 12 first, and a
//     warp whose 128 pixels all fail skips the other 13 positions;
//   * response = float(score) + offset(k) with the reference's sequentially accumulated float offset
//     reproduced exactly from a piecewise-linear table of its bit pattern; rows that
//     cannot reach the threshold (score below the band's minimum useful score) never enter the float path;
//   * candidates (response > threshold) are appended to the frame's slot with one atomic per warp and row.
// HBM traffic: every input byte is read once from DRAM (neighbouring lanes / row bands re-read it from
// L1 / L2); output is the candidate list (+1 B/px when the dense score map is requested).
#include "fd_fast_ring.cuh"

namespace fdb {

namespace {

using namespace fastring;

template <bool PRECHECK, bool SCORE_MAP, bool MASKED>
__global__ void __launch_bounds__(FAST_THREADS, FAST_CTAS_PER_SM) fast_kernel(const FastArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *lut = smem;                                               // 65536 B: run-length table
    OffsetSeg *segs = reinterpret_cast<OffsetSeg *>(smem + 65536);     // p.n_seg + 1 entries (at most FAST_MAX_SEGS)
    // per-warp staging buffer for candidate keys: one global atomic per flush instead of one per row
    uint64_t *stage = reinterpret_cast<uint64_t *>(smem + 65536 + FAST_MAX_SEGS * sizeof(OffsetSeg)) + (threadIdx.x >> 5) * FAST_STAGE_KEYS;
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(lut)[i] = __ldg(reinterpret_cast<const uint4 *>(p.lut) + i);
    for (int i = threadIdx.x; i <= p.n_seg; i += blockDim.x) segs[i] = p.segs[i];
    __syncthreads();

    const FrameView &fv = p.fv;
    const int lane = lane_id();
    const __half2 diff2 = __float2half2_rn(float(p.diff));
    const int inner_cols = fv.cols - 6;
    const int last_row = fv.rows - 1;

    for (bool first = true;; first = false) {   // next_work_item (fd_common.cuh): own first item, then the shared counter
        const int64_t item = next_work_item(p.work_counter, first);
        if (item >= p.n_items) break;
        // item -> (frame, band, strip); strips fastest so neighbouring warps share L1 lines
        const int strip = int(item % p.n_strips);
        const int64_t t = item / p.n_strips;
        const int band = int(t % p.n_bands);
        const int frame = int(t / p.n_bands);
        const int row_begin = p.proc_lo + band * p.band_rows;
        const int row_end = min(row_begin + p.band_rows, p.proc_hi);
        if (row_begin >= row_end) continue;

        // Rows are read as three aligned words per lane through three running pointers.  Addresses are clamped
        // instead of predicated: a clamped word only ever feeds pixels outside the interior, whose results are
        // masked below, and a clamped row only feeds rows past the band's end, which are dropped.
        const int w = strip * 32 + lane;                                // this lane's word in the row
        const uint8_t *fbase = fv.data + int64_t(frame) * fv.frame_stride + int64_t(row_begin - 3) * fv.pitch;
        const uint8_t *pl = fbase + min(max(w - 1, 0), fv.words_per_row - 1) * 4;
        const uint8_t *pc = fbase + min(w, fv.words_per_row - 1) * 4;
        const uint8_t *pr = fbase + min(w + 1, fv.words_per_row - 1) * 4;
        int next_row = row_begin - 3;  // the row the pointers address

        Row rw[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            make_row(rw[i], ld_word(pl), ld_word(pc), ld_word(pr));
            const int64_t adv = (next_row < last_row) ? fv.pitch : 0;
            pl += adv;
            pc += adv;
            pr += adv;
            ++next_row;
        }

        const int col0 = 4 * w;
        uint32_t col_ok = 0u;  // 0xFF per interior column [3, cols-4] of this lane
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (col0 + j >= 3 && col0 + j <= fv.cols - 4) col_ok |= 0xFFu << (8 * j);

        // p.kmin[s] = first pixel index k at which fl(s + offset(k)) > threshold (host-computed from the same table;
        // non-increasing in s).  s_last / s_first = smallest score that makes a candidate at the last / first pixel
        // of the current group of 7 rows; both only move down as k grows.
        int s_last = 17, s_first = 17;
        int seg = 0;                // linear piece of the offset table that holds the first pixel of the current row group
        uint32_t n_staged = 0u;     // warp-uniform fill of the staging buffer
        uint32_t *counter = p.cand_counts + frame;
        uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;
        uint8_t *score_ptr = nullptr;
        if (SCORE_MAP) score_ptr = p.score_map + (int64_t(frame) * fv.rows + row_begin) * fv.cols + col0;
        // Pre-existing features (MASKED): masked-out pixels are neither scored nor counted by the running offset
        // (fast.cpp:88), so the pixel index k comes from the mask's prefix counts instead of the raster formula.
        const int mwpr = p.mask.words_per_row;
        const uint32_t *mbits = nullptr, *mprefix = nullptr, *mrow_base = nullptr;
        uint32_t m_interior = 0u;
        if (MASKED) {
            mbits = p.mask.bits + int64_t(frame) * fv.rows * mwpr + (col0 >> 5);
            mprefix = p.mask.word_prefix + int64_t(frame) * fv.rows * mwpr + (col0 >> 5);
            mrow_base = p.mask.row_base + int64_t(frame) * (fv.rows + 1);
            m_interior = fast_interior_bits(col0 >> 5, fv.cols);
        }
        // k of the first interior pixel of row r (r in [3, rows-3]; rows-3 gives the total)
        auto row_k0 = [&](int r) -> uint32_t { return MASKED ? __ldg(mrow_base + r) : uint32_t(r + p.tile.row_offset - 3) * uint32_t(inner_cols); };
        bool group_empty = false;

        for (int row = row_begin; row < row_end; ++row) {
            // The row loop is deliberately NOT unrolled: the seven window rows shift down by register moves instead of
            // by renaming, which keeps the whole kernel inside the instruction cache (an unrolled-by-7 body does not fit
            // and stalls on instruction fetch).
            if (((row - row_begin) % 7) == 0) {
                const uint32_t k_first = row_k0(row);
                const uint32_t k_end = row_k0(min(row + 7, row_end));  // one past the last pixel index of the group
                group_empty = (k_end == k_first);                      // (only a mask can empty a group)
                if (!group_empty) {
                    const uint32_t k_last = k_end - 1u;
                    while (s_last > 0 && p.kmin[s_last - 1] <= k_last) --s_last;
                    while (s_first > 0 && p.kmin[s_first - 1] <= k_first) --s_first;
                    while (k_first >= segs[seg + 1].k_start) ++seg;  // warp-uniform, monotonic over the band
                }
            }
            // adding need_add to the packed scores sets bit 7 of every byte whose score >= s_last
            const uint32_t need_add = (s_last > 16 || group_empty) ? 0u : (0x80u - uint32_t(s_last)) * 0x01010101u;
            const int prune = SCORE_MAP ? 0 : (s_last >= 8 ? 2 : (s_last >= 4 ? 1 : 0));

            // the row that enters the window next step (row + 4)
            const uint32_t n0 = ld_word(pl), n1 = ld_word(pc), n2 = ld_word(pr);
            const int64_t adv = (next_row < last_row) ? fv.pitch : 0;
            pl += adv;
            pc += adv;
            pr += adv;
            ++next_row;
            uint32_t sp;
            fast_step<PRECHECK>(rw[0], rw[1], rw[2], rw[3], rw[4], rw[5], rw[6], diff2, lut, prune, sp);
            sp &= col_ok;
            if (SCORE_MAP) {
                if (col0 < fv.cols) {
                    if (p.score_aligned && col0 + 3 < fv.cols) {
                        *reinterpret_cast<uint32_t *>(score_ptr) = sp;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col0 + j < fv.cols) score_ptr[j] = uint8_t(sp >> (8 * j));
                    }
                }
                score_ptr += fv.cols;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) rw[i] = rw[i + 1];
            make_row(rw[6], n0, n1, n2);

            // response = score + offset(k), k = index of the pixel among the masked-in interior pixels in raster
            // order (fast.cpp:85-93).  Only rows holding a pixel whose score can reach the threshold get here, and
            // only those pixels do the float work.  Candidates go to the warp's staging buffer.
            uint32_t able = (sp + need_add) & col_ok & 0x80808080u;  // interior pixels with score >= s_last
            uint32_t mword = 0u;
            if (MASKED) {
                if (able != 0u) {
                    mword = __ldg(mbits + int64_t(row) * mwpr);
                    const uint32_t nib = (mword >> (col0 & 31)) & 0xFu;               // mask bits of this lane's 4 pixels
                    able &= ((nib * 0x00204081u) & 0x01010101u) * 0x80u;              // bit j -> bit 7 of byte j
                }
            }
            if (__any_sync(0xffffffffu, able != 0u)) {
                const int r = row;
                uint32_t k_row, k_lo;  // k_lo <= k of every interior pixel of this lane
                if (MASKED) {
                    k_row = row_k0(r);
                    k_lo = (able != 0u) ? k_row + __ldg(mprefix + int64_t(r) * mwpr) : k_row;
                } else {
                    k_row = uint32_t(r + p.tile.row_offset - 3) * uint32_t(inner_cols);
                    k_lo = k_row + uint32_t(max(col0 - 3, 0));
                }
                int sg = seg;
                if (able != 0u)
                    while (k_lo >= segs[sg + 1].k_start) ++sg;
                uint32_t mine = 0u;
                float resp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    resp[j] = 0.0f;
                    if ((able >> (8 * j + 7)) & 1u) {
                        const uint32_t k = MASKED ? k_lo + __popc(mword & m_interior & ((1u << ((col0 & 31) + j)) - 1u))
                                                  : k_row + uint32_t(col0 + j - 3);
                        int sj = sg;
                        while (k >= segs[sj + 1].k_start) ++sj;
                        const float off = __uint_as_float(segs[sj].bits_start + (k - segs[sj].k_start) * segs[sj].step);
                        const float v = __fadd_rn(float((sp >> (8 * j)) & 0xFFu), off);
                        // inside a group whose first and last pixel need the same score the byte test above is already
                        // exact; otherwise the group straddles a threshold crossing and the float test decides
                        if (s_first == s_last || v > p.thr) {
                            resp[j] = v;
                            mine |= 1u << j;
                        }
                    }
                }
                // warp-local append: ballots give every candidate its place without shuffles or atomics
                uint32_t base = n_staged;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (mine >> j) & 1u);
                    if ((mine >> j) & 1u) stage[base + __popc(m & ((1u << lane) - 1u))] = make_cand_key(resp[j], uint32_t(r + p.tile.row_offset), uint32_t(col0 + j));
                    base += __popc(m);
                }
                n_staged = base;
                if (n_staged > FAST_STAGE_KEYS - 128) {  // a row adds at most 128 keys
                    __syncwarp();
                    uint32_t g = 0u;
                    if (lane == 0) g = atomicAdd(counter, n_staged);
                    g = __shfl_sync(0xffffffffu, g, 0);
                    for (uint32_t i = lane; i < n_staged; i += 32)
                        if (g + i < p.cand_capacity) slot[g + i] = stage[i];
                    __syncwarp();
                    n_staged = 0u;
                }
            }
        }
        if (n_staged != 0u) {  // flush what the band left behind
            __syncwarp();
            uint32_t g = 0u;
            if (lane == 0) g = atomicAdd(counter, n_staged);
            g = __shfl_sync(0xffffffffu, g, 0);
            for (uint32_t i = lane; i < n_staged; i += 32)
                if (g + i < p.cand_capacity) slot[g + i] = stage[i];
            __syncwarp();
        }
    }
}

}  // namespace

size_t fast_smem_bytes(int n_seg) {
    (void)n_seg;
    return 65536 + FAST_MAX_SEGS * sizeof(OffsetSeg) + size_t(FAST_THREADS / 32) * FAST_STAGE_KEYS * 8;
}

template <bool PRECHECK, bool SCORE_MAP, bool MASKED>
static cudaError_t launch_fast_t(const FastArgs &args, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(fast_kernel<PRECHECK, SCORE_MAP, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    fast_kernel<PRECHECK, SCORE_MAP, MASKED><<<grid, FAST_THREADS, smem, stream>>>(args);
    return cudaGetLastError();
}

template <bool PRECHECK, bool SCORE_MAP>
static cudaError_t launch_fast_m(const FastArgs &args, int grid, size_t smem, cudaStream_t stream) {
    return args.mask.bits != nullptr ? launch_fast_t<PRECHECK, SCORE_MAP, true>(args, grid, smem, stream)
                                     : launch_fast_t<PRECHECK, SCORE_MAP, false>(args, grid, smem, stream);
}

cudaError_t launch_fast(const FastArgs &args, bool precheck, int grid, cudaStream_t stream) {
    const size_t smem = fast_smem_bytes(args.n_seg);
    const bool sm = args.score_map != nullptr;
    if (precheck) return sm ? launch_fast_m<true, true>(args, grid, smem, stream) : launch_fast_m<true, false>(args, grid, smem, stream);
    return sm ? launch_fast_m<false, true>(args, grid, smem, stream) : launch_fast_m<false, false>(args, grid, smem, stream);
}

}  // namespace fdb
