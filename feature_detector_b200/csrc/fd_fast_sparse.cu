// Kernel 2, sparse form -- FAST candidates when the threshold rules most pixels out before their score is known.
// Replaces FeaturePointFastDetector::ComputeResponseOfPixel / ComputeCandidates
// (reference src/feature_point_detector/feature_point_fast_detector.cpp:11-81, 83-98), same results as the dense
// kernel in fd_fast.cu bit for bit.
//
// A pixel becomes a candidate iff fl(score + offset(k)) > threshold (fast.cpp:89-92), offset(k) being the reference's
// running float.  The host knows, per score s, the first pixel index at which s passes (kmin[s], fd_api.cu), hence the
// smallest useful score s_min for any stretch of rows.  A circular run of s_min >= 4 same-polarity ring pixels must
// contain a compass position (ring index 0, 4, 8 or 12), a run of >= 8 two adjacent ones, and a polarity flag implies
// |ring - centre| > diff.  So the pass over the frame is split in two:
//
//   A (dense, 4 pixels per lane per instruction): |ring - centre| at the four compass positions with VABSDIFF4 on
//     packed bytes (one ALU-pipe instruction for four pixels), tested against the largest power of two <= diff + 1
//     -- exact for the reference's diff = 15, conservative otherwise -- and combined per 32-bit word:
//        s_min in 4..7 : any compass position differs            (ANY)
//        s_min >= 8    : (top | bottom) and (left | right) differ (ADJ)
//        kN >= 12      : right, bottom and left all differ        (PRE; the reference's pre-check, fast.cpp:20-42,
//                                                                   only needs s_min >= 1 because a failed pre-check
//                                                                   scores 0)
//     About 17 instructions per 128 pixels; 2-19 % of the words survive on the synthetic frames.
//   B (sparse, exact): surviving words are queued per warp; every 32 of them are scored by the full half2 ring test of
//     fd_fast_ring.cuh, one word (4 pixels) per lane, and emit candidates exactly as the dense kernel does.
//
// Data movement: each warp owns a 128-pixel column strip and streams down a band of rows through a private 32-row ring
// in shared memory that TMA fills 8 rows at a time (cp.async.bulk.tensor, 3-D map cols x rows x frames, box 160 x 8 x 1,
// out-of-frame bytes zero-filled by the hardware), one group ahead of the row being tested.  Phase A reads five
// conflict-free words per row; phase B gathers its 21 words per lane from the same ring.  HBM traffic is the frame,
// once, plus the candidate keys.
#include <cuda.h>

#include <type_traits>

#include "fd_fast_ring.cuh"

namespace fdb {

namespace {

using namespace fastring;

constexpr int SP_WARPS = FAST_SPARSE_THREADS / 32;
constexpr int SP_RING_ROWS = 4 * FAST_SPARSE_GROUP_ROWS;  // rows resident per warp (power of two)
constexpr int SP_GROUP_ROWS = FAST_SPARSE_GROUP_ROWS;  // rows per TMA box
constexpr int SP_GROUPS = SP_RING_ROWS / SP_GROUP_ROWS;
constexpr int SP_ROW_WORDS = 40;                    // 160-byte box rows: 16-byte halo, 32 strip words, 16-byte halo (TMA box starts must be 16-byte aligned)
constexpr int SP_W0 = 3;                            // ring word that holds the 4 pixels left of lane 0's own word
constexpr int SP_QUEUE = 512;                       // survivor queue entries per warp (power of two, >= 31 + 32 * SP_GROUP_ROWS)
constexpr int SP_STAGE = 192;                       // candidate staging keys per warp (a batch adds at most 128)
constexpr uint32_t SP_GROUP_BYTES = SP_GROUP_ROWS * SP_ROW_WORDS * 4;

struct WarpSmem {
    uint32_t ring[SP_RING_ROWS * SP_ROW_WORDS];     // must stay first: TMA destinations need 128-byte alignment
    uint64_t stage[SP_STAGE];
    uint16_t queue[SP_QUEUE];                       // (local centre row << 5) | lane; bands are at most 2040 rows
    uint64_t bar[SP_GROUPS];
    uint32_t pad[24];                               // keeps sizeof(WarpSmem) a multiple of 128
};
static_assert(SP_GROUP_ROWS >= 6 && SP_QUEUE >= 32 + 32 * SP_GROUP_ROWS, "ring / queue geometry");
static_assert(sizeof(WarpSmem) % 128 == 0, "per-warp shared block must keep the rings 128-byte aligned");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load: box (160 bytes, 8 rows, 1 frame) at (x, y, frame) into shared memory, completing on `bar`.
__device__ __forceinline__ void tma_load_rows(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

enum : int { MODE_ANY = 0, MODE_ADJ = 1 };

// MASKED: pre-existing features (feature_point_detector.cpp:12-16): masked-out pixels are neither scored nor counted by the running
// offset (fast.cpp:88), so phase B takes the pixel index from the mask's prefix counts; phase A and the score bounds of a band
// keep using the raster index, which is never smaller (a larger index only admits lower scores: conservative).
template <bool PRECHECK, bool MASKED>
__global__ void __launch_bounds__(FAST_SPARSE_THREADS, 1) fast_sparse_kernel(const FastArgs p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ OffsetSeg segs[FAST_MAX_SEGS];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // TMA destinations must be 128-byte aligned: round the dynamic window up (the launch reserves the slack)
    uint8_t *smem_al = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_al)[warp];
    for (int i = threadIdx.x; i <= p.n_seg; i += blockDim.x) segs[i] = p.segs[i];
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < SP_GROUPS; ++g) mbar_init(&ws.bar[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const FrameView &fv = p.fv;
    const __half2 diff2 = __float2half2_rn(float(p.diff));
    const int inner_cols = fv.cols - 6;
    const uint32_t hm = p.absdiff_mask;   // per byte: bits at or above the largest power of two <= diff + 1
    uint32_t slot_parity = 0u;            // bit s: parity of the phase ring slot s completes next

    for (bool first = true;; first = false) {   // next_work_item (fd_common.cuh): own first item, then the shared counter
        const int64_t item = next_work_item(p.work_counter, first);
        if (item >= p.n_items) break;
        const int strip = int(item % p.n_strips);
        const int64_t t = item / p.n_strips;
        const int band = int(t % p.n_bands);
        const int frame = int(t / p.n_bands);
        const int row_begin = p.proc_lo + band * p.band_rows;
        const int row_end = min(row_begin + p.band_rows, p.proc_hi);
        const int abs0 = p.tile.row_offset;   // absolute row of local row 0 (row tiles of a larger frame)
        if (row_begin >= row_end) continue;
        const int n_rows = row_end - row_begin;
        // local row index lr <-> image row (row_begin - 3 + lr); centre rows are lr = 3 .. n_rows + 2
        const int n_groups = (n_rows + 6 + SP_GROUP_ROWS - 1) / SP_GROUP_ROWS;
        const int x0 = strip * 128 - 16;  // first byte of the 160-byte box; the hardware requires 16-byte aligned box starts

        auto issue_group = [&](int g) {   // warp-uniform; lane 0 talks to the TMA unit
            __syncwarp();                 // every lane is done reading the slot this overwrites
            if (lane == 0) {
                const int s = g & (SP_GROUPS - 1);
                mbar_expect_tx(&ws.bar[s], SP_GROUP_BYTES);
                tma_load_rows(&ws.ring[s * SP_GROUP_ROWS * SP_ROW_WORDS], &tmap, x0, row_begin - 3 + g * SP_GROUP_ROWS, frame, &ws.bar[s]);
            }
        };
        auto wait_group = [&](int g) {
            const int s = g & (SP_GROUPS - 1);
            mbar_wait(&ws.bar[s], (slot_parity >> s) & 1u);
            slot_parity ^= 1u << s;
        };

        issue_group(0);

        const int col0 = strip * 128 + 4 * lane;
        uint32_t col_ok = 0u;  // 0xFF per interior column [3, cols-4] of this lane
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (col0 + j >= 3 && col0 + j <= fv.cols - 4) col_ok |= 0xFFu << (8 * j);

        // smallest score that can pass anywhere in this band: at its last pixel (kmin is non-increasing in s)
        int s_band = 17;
        {
            const uint32_t k_last = uint32_t(row_end - 1 + abs0 - 3) * uint32_t(inner_cols) + uint32_t(inner_cols - 1);
            while (s_band > 0 && p.kmin[s_band - 1] <= k_last) --s_band;
        }
        const uint32_t need_add = (s_band > 16) ? 0u : (0x80u - uint32_t(s_band)) * 0x01010101u;
        const int mwpr = p.mask.words_per_row;
        const uint32_t *mbits = nullptr, *mprefix = nullptr, *mrow_base = nullptr;
        if (MASKED) {
            mbits = p.mask.bits + int64_t(frame) * fv.rows * mwpr;
            mprefix = p.mask.word_prefix + int64_t(frame) * fv.rows * mwpr;
            mrow_base = p.mask.row_base + int64_t(frame) * (fv.rows + 1);
        }
        int seg_band = 0;
        {
            const uint32_t k_first = MASKED ? __ldg(mrow_base + row_begin) : uint32_t(row_begin + abs0 - 3) * uint32_t(inner_cols);
            while (k_first >= segs[seg_band + 1].k_start) ++seg_band;
        }

        uint32_t qh = 0u, qn = 0u;   // queue head / tail (monotonic counters, warp-uniform)
        uint32_t n_staged = 0u;
        uint32_t *counter = p.cand_counts + frame;
        uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;

        auto flush_stage = [&]() {
            __syncwarp();
            uint32_t g = 0u;
            if (lane == 0) g = atomicAdd(counter, n_staged);
            g = __shfl_sync(0xffffffffu, g, 0);
            for (uint32_t i = lane; i < n_staged; i += 32)
                if (g + i < p.cand_capacity) slot[g + i] = ws.stage[i];
            __syncwarp();
            n_staged = 0u;
        };

        // ---- phase B: exact scores of up to 32 queued words, candidates appended to the staging buffer ----
        auto drain_batch = [&]() {
            __syncwarp();
            const uint32_t idx = qh + uint32_t(lane);
            const bool valid = idx < qn;
            const uint32_t e = ws.queue[(valid ? idx : qh) & (SP_QUEUE - 1)];
            qh = min(qh + 32u, qn);
            const int lc = int(e >> 5), ln = int(e & 31u);
            Row rw[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const uint32_t *rp = &ws.ring[((lc - 3 + i) & (SP_RING_ROWS - 1)) * SP_ROW_WORDS + SP_W0 + ln];
                make_row(rw[i], rp[0], rp[1], rp[2]);
            }
            uint32_t sp;
            fast_step<PRECHECK>(rw[0], rw[1], rw[2], rw[3], rw[4], rw[5], rw[6], diff2, p.lut, 0, sp);
            const int c0 = strip * 128 + 4 * ln;
            uint32_t ok = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j >= 3 && c0 + j <= fv.cols - 4) ok |= 0xFFu << (8 * j);
            if (!valid) ok = 0u;
            sp &= ok;
            uint32_t able = (sp + need_add) & ok & 0x80808080u;  // score >= s_band: may pass somewhere in the band
            uint32_t mword = 0u;
            if (MASKED && able != 0u) {
                mword = __ldg(mbits + int64_t(row_begin - 3 + lc) * mwpr + (c0 >> 5));
                const uint32_t nib = (mword >> (c0 & 31)) & 0xFu;                  // mask bits of this lane's 4 pixels
                able &= ((nib * 0x00204081u) & 0x01010101u) * 0x80u;               // bit j -> bit 7 of byte j
            }
            if (__any_sync(0xffffffffu, able != 0u)) {
                const int r = row_begin - 3 + lc + abs0;   // absolute row (masks are never combined with row tiles: abs0 = 0)
                uint32_t k_row = uint32_t(r - 3) * uint32_t(inner_cols), k_lo = k_row + uint32_t(max(c0 - 3, 0));
                uint32_t m_interior = 0u;
                if (MASKED && able != 0u) {
                    k_row = __ldg(mrow_base + r);
                    k_lo = k_row + __ldg(mprefix + int64_t(r) * mwpr + (c0 >> 5));   // masked-in interior pixels of the row before this mask word
                    m_interior = fast_interior_bits(c0 >> 5, fv.cols);
                }
                int sg = seg_band;
                if (able != 0u)
                    while (k_lo >= segs[sg + 1].k_start) ++sg;
                uint32_t mine = 0u;
                float resp[4];
                // The word's four pixels have consecutive indices, and a segment of the offset table spans thousands: unless a
                // boundary falls inside the word (or a mask thins the indices out) the four offsets are one base and one step.
                const bool one_seg = !MASKED && able != 0u && k_row + uint32_t(c0) < segs[sg + 1].k_start;   // pixel j = 3 is the word's last
                if (one_seg) {
                    const uint32_t step = segs[sg].step;
                    const uint32_t bits0 = segs[sg].bits_start + (k_row + uint32_t(c0 - 3) - segs[sg].k_start) * step;   // (never used for a j left of the interior)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = __fadd_rn(float((sp >> (8 * j)) & 0xFFu), __uint_as_float(bits0 + uint32_t(j) * step));   // fast.cpp:89, one rounding
                        const bool pass = ((able >> (8 * j + 7)) & 1u) && v > p.thr;                                               // fast.cpp:90
                        resp[j] = pass ? v : 0.0f;
                        mine |= pass ? (1u << j) : 0u;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        resp[j] = 0.0f;
                        if ((able >> (8 * j + 7)) & 1u) {
                            const uint32_t k = MASKED ? k_lo + __popc(mword & m_interior & ((1u << ((c0 & 31) + j)) - 1u)) : k_row + uint32_t(c0 + j - 3);
                            int sj = sg;
                            while (k >= segs[sj + 1].k_start) ++sj;
                            const float off = __uint_as_float(segs[sj].bits_start + (k - segs[sj].k_start) * segs[sj].step);
                            const float v = __fadd_rn(float((sp >> (8 * j)) & 0xFFu), off);   // fast.cpp:89, one rounding
                            if (v > p.thr) {                                                    // fast.cpp:90
                                resp[j] = v;
                                mine |= 1u << j;
                            }
                        }
                    }
                }
                uint32_t base = n_staged;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (mine >> j) & 1u);
                    if ((mine >> j) & 1u) ws.stage[base + __popc(m & ((1u << lane) - 1u))] = make_cand_key(resp[j], uint32_t(r), uint32_t(c0 + j));
                    base += __popc(m);
                }
                n_staged = base;
                if (n_staged > SP_STAGE - 128) flush_stage();
            }
        };

        // ---- phase A over the band, one TMA group (8 rows) per step ----
        // Step g tests the 8 centre rows lc = 8 g - 3 + i, whose ring neighbourhood lc - 3 .. lc + 3 lies in ring groups
        // g - 1 and g; group g + 1 is requested at the start of the step and overwrites group g - 3, which queued words
        // of steps <= g - 2 would still need, so those are drained first.  Inside a step every shared-memory address is
        // a group base plus a compile-time offset, and each lane collects its 8 survivor flags in one register.
        const int lc_end = n_rows + 3;
        for (int g = 0;; ++g) {
            // The one place phase B runs (a single copy of its code keeps the kernel inside the instruction cache):
            // full batches, anything step g's TMA request would strand, and at the end of the band whatever is left.
            while (qh < qn) {   // warp-uniform
                if (qn - qh < 32u && g < n_groups) {
                    const int oldest_step = (int(ws.queue[qh & (SP_QUEUE - 1)] >> 5) + 3) / SP_GROUP_ROWS;
                    if (oldest_step >= g - 1 || g + 1 >= n_groups) break;
                }
                drain_batch();
            }
            if (g == n_groups) break;
            if (g + 1 < n_groups) issue_group(g + 1);
            wait_group(g);
            const uint32_t *cur = &ws.ring[(g & (SP_GROUPS - 1)) * SP_GROUP_ROWS * SP_ROW_WORDS + SP_W0 + lane];
            const uint32_t *prev = &ws.ring[((g - 1) & (SP_GROUPS - 1)) * SP_GROUP_ROWS * SP_ROW_WORDS + SP_W0 + lane];
            const int lc0 = SP_GROUP_ROWS * g - 3;
            // ADJ is allowed when score 7 cannot pass anywhere in the step's rows (kmin is non-increasing in the score)
            bool adj = false;
            if (!PRECHECK) {
                const int r_last = row_begin - 3 + abs0 + min(lc0 + SP_GROUP_ROWS - 1, lc_end - 1);
                adj = p.kmin[7] > uint32_t(r_last - 3) * uint32_t(inner_cols) + uint32_t(inner_cols - 1);
            }
            uint32_t bits = 0u;   // bit i: this lane's word of centre row lc0 + i survives
            auto test_rows = [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;   // 0 ANY, 1 ADJ, 2 PRE
#pragma unroll
                for (int i = 0; i < SP_GROUP_ROWS; ++i) {
                    const uint32_t *rc = (i >= 3) ? cur + (i - 3) * SP_ROW_WORDS : prev + (SP_GROUP_ROWS - 3 + i) * SP_ROW_WORDS;
                    const uint32_t *ru = (i >= 6) ? cur + (i - 6) * SP_ROW_WORDS : prev + (SP_GROUP_ROWS - 6 + i) * SP_ROW_WORDS;
                    const uint32_t wl = rc[0], wc = rc[1], wr = rc[2];
                    const uint32_t a4 = __vabsdiffu4(__funnelshift_r(wc, wr, 24), wc);          // ring 4: (row, col + 3)
                    const uint32_t a12 = __vabsdiffu4(__funnelshift_r(wl, wc, 8), wc);          // ring 12: (row, col - 3)
                    const uint32_t a8 = __vabsdiffu4(cur[(i) * SP_ROW_WORDS + 1], wc);          // ring 8: (row + 3, col)
                    bool surv;
                    if (MODE == 2) {
                        // right, bottom and left all differ by more than diff in the same pixel: (x >> s) + 0x7f sets bit 7 per byte
                        const uint32_t f4 = ((a4 & hm) >> p.absdiff_shift) + 0x7F7F7F7Fu;
                        const uint32_t f8 = ((a8 & hm) >> p.absdiff_shift) + 0x7F7F7F7Fu;
                        const uint32_t f12 = ((a12 & hm) >> p.absdiff_shift) + 0x7F7F7F7Fu;
                        surv = (f4 & f8 & f12 & 0x80808080u) != 0u;
                    } else {
                        const uint32_t a0 = __vabsdiffu4(ru[1], wc);                            // ring 0: (row - 3, col)
                        if (MODE == 1) surv = (((a0 | a8) & hm) != 0u) && (((a4 | a12) & hm) != 0u);
                        else surv = ((a0 | a4 | a8 | a12) & hm) != 0u;
                    }
                    if (surv) bits |= 1u << i;
                }
            };
            if (PRECHECK) test_rows(std::integral_constant<int, 2>());
            else if (adj) test_rows(std::integral_constant<int, 1>());
            else test_rows(std::integral_constant<int, 0>());
            // rows outside the band (the halo rows of the first and last step) and lanes outside the interior drop out
            {
                const int lo = max(3 - lc0, 0), hi = min(lc_end - lc0, SP_GROUP_ROWS);   // live centre rows: i in [lo, hi)
                const uint32_t live = (hi > lo) ? (((1u << (hi - lo)) - 1u) << lo) : 0u;
                bits &= (col_ok != 0u) ? live : 0u;
            }
            // bulk append of the step's survivors: exclusive prefix of the per-lane counts, then each lane writes its rows
            if (__any_sync(0xffffffffu, bits != 0u)) {
                const uint32_t cnt = __popc(bits);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                uint32_t pos = qn + incl - cnt;
                qn += __shfl_sync(0xffffffffu, incl, 31);
                while (bits != 0u) {
                    const int i = __ffs(bits) - 1;
                    bits &= bits - 1u;
                    ws.queue[pos & (SP_QUEUE - 1)] = uint16_t((uint32_t(lc0 + i) << 5) | uint32_t(lane));
                    ++pos;
                }
            }
        }
        if (n_staged != 0u) flush_stage();
        __syncwarp();
    }
}

}  // namespace

size_t fast_sparse_smem_bytes() { return size_t(SP_WARPS) * sizeof(WarpSmem) + 128; }

namespace {
template <bool PRECHECK, bool MASKED>
cudaError_t launch_sparse_t(const FastArgs &args, const CUtensorMap &map, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(fast_sparse_kernel<PRECHECK, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    fast_sparse_kernel<PRECHECK, MASKED><<<grid, FAST_SPARSE_THREADS, smem, stream>>>(args, map);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_fast_sparse(const FastArgs &args, const void *tensor_map, bool precheck, int grid, cudaStream_t stream) {
    const size_t smem = fast_sparse_smem_bytes();
    CUtensorMap map;
    memcpy(&map, tensor_map, sizeof(map));
    const bool masked = args.mask.bits != nullptr;
    if (precheck) return masked ? launch_sparse_t<true, true>(args, map, grid, smem, stream) : launch_sparse_t<true, false>(args, map, grid, smem, stream);
    return masked ? launch_sparse_t<false, true>(args, map, grid, smem, stream) : launch_sparse_t<false, false>(args, map, grid, smem, stream);
}

}  // namespace fdb
