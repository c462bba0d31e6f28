// Kernel 1 -- gradients, 3x3 box structure tensor, Harris / Shi-Tomasi response, 4-neighbour NMS and
// candidate compaction, fused in one pass over the frame.
// Replaces ComputeHorizontalGradientSums / ComputeResponseMap / PerformNMSAndExtractCandidates of
// reference src/feature_point_detector/feature_point_harris_detector.cpp:17-137 and
// feature_point_shi_tomas_detector.cpp:17-137 (the reference's gradients are plain central differences
// and its "Shi-Tomas" response is the LARGER eigenvalue -- SURVEY.md D1, D3 -- reproduced as is).
//
// Design (sm_100a; memory-bound stencil by construction, issue-bound in practice -- see DESIGN.md)
//   * a warp owns a 128-column strip and streams down a band of rows; each lane owns 4 adjacent columns.
//     Strips advance by 126 columns so that every output column has both horizontal NMS neighbours
//     inside its own warp (one shuffle each); the strip start is therefore not word aligned and rows are
//     fetched as aligned 32-bit words and funnel-shifted into place;
//   * all window sums are taken in int32 (IADD3): every partial sum is an exact integer below 2^24
//     (SURVEY.md H1), so the order of the reference's float sliding sums is irrelevant.  The products
//     carry a bias chosen so the 9-term sum lands directly on the bit pattern of a float in [2^23, 2^24),
//     which makes int -> float one FADD instead of an I2F;
//   * the response is evaluated with explicit round-to-nearest mul/add/sub/sqrt intrinsics in the
//     reference's association order: the reference binary has no FMA (its CMakeLists.txt:6 has no -march);
//   * a register pipeline three rows deep: pixel row n arrives -> products of row n-1 -> response of row
//     n-2 -> NMS of row n-3.  Nothing but the input frame is read from HBM and nothing but candidates
//     (and, on request, the dense response map) is written.
#include "fd_corner_common.cuh"

namespace fdb {

namespace {

using corner::response_of;

// Bias per product so that nine of them add up to the bit pattern of a float whose value is
// (sum + kSumBase): 9 * kBiasPos = 0x4B000006 -> float(2^23 + 6 + sum)      (sum >= 0)
//                   9 * kBiasMid = 0x4B400008 -> float(1.5 * 2^23 + 8 + sum)  (|sum| < 2^22)
constexpr int32_t kBiasPos = 139810134;
constexpr int32_t kBiasMid = 140276168;
constexpr float kSumBasePos = 8388614.0f;
constexpr float kSumBaseMid = 12582920.0f;

struct PixRow {
    int32_t v[8];  // columns c0-2 .. c0+5
};

struct HRow {
    int32_t xx[4], yy[4], xy[4];  // horizontal 3-sums of the (biased) products at columns c0 .. c0+3
};

__device__ __forceinline__ void extract_row(PixRow &r, uint32_t a, uint32_t b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        r.v[j] = int32_t((a >> (8 * j)) & 0xFFu);
        r.v[4 + j] = int32_t((b >> (8 * j)) & 0xFFu);
    }
}

// Products of gradient row `mid` (needs `up` = row above, `dn` = row below) and their horizontal 3-sums.
__device__ __forceinline__ void product_row(HRow &h, const PixRow &up, const PixRow &mid, const PixRow &dn) {
    int32_t pxx[6], pyy[6], pxy[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {  // column c0-1+j  <->  v index j+1
        const int32_t ix = mid.v[j + 2] - mid.v[j];      // harris.cpp:36
        const int32_t iy = dn.v[j + 1] - up.v[j + 1];    // harris.cpp:37
        pxx[j] = ix * ix + kBiasPos;                     // harris.cpp:38-40
        pyy[j] = iy * iy + kBiasPos;
        pxy[j] = ix * iy + kBiasMid;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                        // harris.cpp:47-62 (horizontal window)
        h.xx[j] = pxx[j] + pxx[j + 1] + pxx[j + 2];
        h.yy[j] = pyy[j] + pyy[j + 1] + pyy[j + 2];
        h.xy[j] = pxy[j] + pxy[j + 1] + pxy[j + 2];
    }
}

constexpr int CORNER_STAGE = 256;  // candidate keys staged per warp (a row adds at most 128)

template <int KIND, bool MASKED>
__global__ void __launch_bounds__(CORNER_THREADS, 2) corner_kernel(const CornerArgs p) {
    __shared__ uint64_t stage_all[CORNER_THREADS / 32][CORNER_STAGE];
    uint64_t *stage = stage_all[threadIdx.x >> 5];
    uint2 *stage2 = reinterpret_cast<uint2 *>(stage);   // little endian: .x = low word of the key, .y = high word
    const FrameView &fv = p.fv;
    const int lane = lane_id();
    const int row_lo = p.resp_lo, row_hi = p.resp_hi;  // rows with a defined response (bound = 2, harris.cpp:90-92)
    const int col_lo = 2, col_hi = fv.cols - 3;

    for (bool first = true;; first = false) {   // next_work_item (fd_common.cuh): own first item, then the shared counter
        const int64_t item = next_work_item(p.work_counter, first);
        if (item >= p.n_items) break;
        const int strip = int(item % p.n_strips);
        const int64_t t = item / p.n_strips;
        const int band = int(t % p.n_bands);
        const int frame = int(t / p.n_bands);
        const int rb = p.cand_lo + band * p.band_rows;
        const int re = min(rb + p.band_rows, p.cand_hi);
        if (rb >= re) continue;

        const int x0 = 1 + CORNER_STRIP_OUT * strip;  // first computed column of the strip
        const int c0 = x0 + 4 * lane;                 // this lane's first column
        const int a0 = c0 - 2;                        // first byte of the lane's 8-byte window
        const int wi = a0 >> 2;                       // floor(a0 / 4): aligned word index (a0 >= -1)
        const int sh = (a0 & 3) * 8;                  // warp-uniform funnel shift
        const uint8_t *fbase = fv.data + int64_t(frame) * fv.frame_stride;
        const bool ok0 = wi >= 0 && wi < fv.words_per_row;
        const bool ok1 = wi + 1 >= 0 && wi + 1 < fv.words_per_row;
        const bool ok2 = wi + 2 >= 0 && wi + 2 < fv.words_per_row;
        auto load_row = [&](int row, uint32_t &a, uint32_t &b) {
            uint32_t w0 = 0u, w1 = 0u, w2 = 0u;
            if (row >= 0 && row < fv.rows) {
                const uint8_t *rp = fbase + int64_t(row) * fv.pitch + 4 * int64_t(wi);
                if (ok0) w0 = ld_word(rp);
                if (ok1) w1 = ld_word(rp + 4);
                if (ok2) w2 = ld_word(rp + 8);
            }
            a = __funnelshift_r(w0, w1, sh);
            b = __funnelshift_r(w1, w2, sh);
        };

        // which of the lane's 4 columns may carry a response / emit a candidate
        bool col_valid[4], col_owned[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            col_valid[j] = (c >= col_lo && c <= col_hi);
            col_owned[j] = col_valid[j] && (c >= x0 + 1) && (c <= x0 + CORNER_STRIP_OUT);
        }

        PixRow px[3];
        HRow hs[3];
        float resp[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) px[i].v[j] = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hs[i].xx[j] = hs[i].yy[j] = hs[i].xy[j] = 0;
                resp[i][j] = 0.0f;
            }
        }
        uint32_t *counter = p.cand_counts + frame;
        uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;
        uint32_t n_staged = 0u;  // warp-uniform fill of the staging buffer: one global atomic per flush, not per row
        auto flush_stage = [&]() {
            __syncwarp();
            uint32_t g = 0u;
            if (lane == 0) g = atomicAdd(counter, n_staged);
            g = __shfl_sync(0xffffffffu, g, 0);
            for (uint32_t i = lane; i < n_staged; i += 32)
                if (g + i < p.cand_capacity) slot[g + i] = stage[i];
            __syncwarp();
            n_staged = 0u;
        };
        float *resp_map = p.response_map ? p.response_map + int64_t(frame) * fv.rows * fv.cols : nullptr;
        // pre-existing features: the response is only evaluated where the mask is set (harris.cpp:94), 0 elsewhere
        const uint32_t *mbits = MASKED ? p.mask.bits + int64_t(frame) * fv.rows * p.mask.words_per_row + (c0 >> 5) : nullptr;

        // Pixel row n arrives at step n.  Slots rotate with period 3: row q lives in slot q mod 3 (shifted
        // so that the unrolled phase index is a compile-time constant).
        const int n_begin = rb - 3, n_end = re + 3;  // exclusive
        uint32_t na, nb;
        load_row(n_begin, na, nb);
        int n = n_begin;
        while (n < n_end) {
#pragma unroll
            for (int ph = 0; ph < 3; ++ph) {
                if (n < n_end) {  // warp-uniform
                    // slot roles this step: cur = ph (row n), prev1 = ph+2 (row n-1), prev2 = ph+1 (row n-2)
                    const int cur = ph % 3, p1 = (ph + 2) % 3, p2 = (ph + 1) % 3;
                    extract_row(px[cur], na, nb);
                    if (n + 1 < n_end) load_row(n + 1, na, nb);  // prefetch
                    // products of row n-1 -> H slot p1 (overwrites row n-4's sums)
                    product_row(hs[p1], px[p2], px[p1], px[cur]);
                    // response of row q = n-2 from H rows n-3 (slot cur), n-2 (slot p2), n-1 (slot p1)
                    const int q = n - 2;
                    const bool q_valid = (q >= row_lo && q <= row_hi);
                    float rq[4];
                    uint32_t mnib = 0xFu;  // mask bits of columns c0 .. c0+3 of row q
                    if (MASKED) {
                        if (q_valid && c0 < fv.cols) {
                            const uint32_t *mp = mbits + int64_t(q) * p.mask.words_per_row;
                            mnib = __funnelshift_r(__ldg(mp), __ldg(mp + 1), c0 & 31) & 0xFu;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int32_t ixx = hs[cur].xx[j] + hs[p2].xx[j] + hs[p1].xx[j];  // harris.cpp:81-88,108-116
                        const int32_t iyy = hs[cur].yy[j] + hs[p2].yy[j] + hs[p1].yy[j];
                        const int32_t ixy = hs[cur].xy[j] + hs[p2].xy[j] + hs[p1].xy[j];
                        const float sxx = __fsub_rn(__int_as_float(ixx), kSumBasePos);
                        const float syy = __fsub_rn(__int_as_float(iyy), kSumBasePos);
                        const float sxy = __fsub_rn(__int_as_float(ixy), kSumBaseMid);
                        rq[j] = (q_valid && col_valid[j] && ((mnib >> j) & 1u)) ? response_of<KIND>(sxx, syy, sxy, p, p.thr) : 0.0f;
                    }
                    if (resp_map != nullptr && q >= rb && q < re) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col_owned[j]) resp_map[int64_t(q) * fv.cols + c0 + j] = rq[j];
                    }
                    // NMS of row m = n-3.  Response slots before this step's store: cur = row n-4, p2 = row n-3,
                    // p1 = row n-5 (dead); rq = row n-2.
                    const int m = n - 3;
                    const float left_in = __shfl_up_sync(0xffffffffu, resp[p2][3], 1);
                    const float right_in = __shfl_down_sync(0xffffffffu, resp[p2][0], 1);
                    if (m >= rb && m < re) {
                        uint32_t mine = 0u;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float v = resp[p2][j];
                            const float l = (j == 0) ? left_in : resp[p2][j - 1];
                            const float r = (j == 3) ? right_in : resp[p2][j + 1];
                            // v == 0 means "at or below threshold" (harris.cpp:130); strict 4-neighbour max (:131-132)
                            if (col_owned[j] && v > p.thr && v > l && v > r && v > resp[cur][j] && v > rq[j]) mine |= 1u << j;
                        }
                        if (__any_sync(0xffffffffu, mine != 0u)) {
                            // key halves written as two 32-bit words: high = ~ordered(response), low = (row << 16) | col
                            const uint32_t lo0 = (uint32_t(m + p.tile.row_offset) << 16) | uint32_t(c0);
                            const uint32_t lt = (1u << lane) - 1u;
                            uint32_t base = n_staged;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const bool on = (mine >> j) & 1u;
                                const uint32_t bm = __ballot_sync(0xffffffffu, on);
                                if (on) {
                                    const uint32_t b = __float_as_uint(resp[p2][j]);
                                    const uint32_t hi = b ^ ~(uint32_t(int32_t(b) >> 31) | 0x80000000u);   // ~float_to_ordered(b)
                                    stage2[base + __popc(bm & lt)] = make_uint2(lo0 + uint32_t(j), hi);
                                }
                                base += __popc(bm);
                            }
                            n_staged = base;
                            if (n_staged > CORNER_STAGE - 128) flush_stage();
                        }
                    }
                    // rotate: response row n-2 replaces row n-5's slot
#pragma unroll
                    for (int j = 0; j < 4; ++j) resp[p1][j] = rq[j];
                    ++n;
                }
            }
        }
        if (n_staged != 0u) flush_stage();
    }
}

}  // namespace

cudaError_t launch_corner(const CornerArgs &args, int grid, cudaStream_t stream) {
    const bool masked = args.mask.bits != nullptr;
    if (args.kind == 0) {
        if (masked) corner_kernel<0, true><<<grid, CORNER_THREADS, 0, stream>>>(args);
        else corner_kernel<0, false><<<grid, CORNER_THREADS, 0, stream>>>(args);
    } else {
        if (masked) corner_kernel<1, true><<<grid, CORNER_THREADS, 0, stream>>>(args);
        else corner_kernel<1, false><<<grid, CORNER_THREADS, 0, stream>>>(args);
    }
    return cudaGetLastError();
}

}  // namespace fdb
