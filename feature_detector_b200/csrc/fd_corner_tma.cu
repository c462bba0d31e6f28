// Kernel 1, TMA form -- gradients, 3x3 box structure tensor, Harris / Shi-Tomasi response, 4-neighbour NMS and candidate
// compaction (reference src/feature_point_detector/feature_point_harris_detector.cpp:17-137,
// feature_point_shi_tomas_detector.cpp:17-137), same results as fd_corner.cu bit for bit.
//
// What differs from the register-streaming kernel:
//   * rows reach the warp through a private shared-memory ring that TMA fills 12 rows at a time (3-D map cols x rows x
//     frames, box 160 x 12 x 1, box starts 16-byte aligned, out-of-frame bytes zero-filled by the hardware), one group in
//     flight ahead of the rows being processed -- no global-load, clamp or address instructions in the row loop, and the load
//     latency is off the critical path;
//   * gradients, products and window sums run in fp32 on the full-rate FMA pipe instead of int32 on the half-rate pipes:
//     a pixel byte b enters as the float 2^23 + b (one PRMT against the constant 0x4B000000), so a difference of two such
//     floats IS the integer gradient, exactly; products (<= 65 025) and 9-term sums (< 2^24) stay exact integers in fp32
//     (SURVEY.md H1), so the sums equal the reference's sliding float sums whatever the order, and no int -> float
//     conversion is left before the response;
//   * the row loop is unrolled over one TMA group (12 rows = 4 turns of the 3-slot register pipeline), so every shared-
//     memory address is a group base plus a compile-time offset.
//   * that exactness is used throughout: neighbouring 3-sums share a pair sum, products that feed one sum ride on an FMA, the
//     column-validity test is the response threshold (+inf outside the interior), threshold and strip ownership are a fifth operand of
//     the 4-neighbour maximum, and the Harris pre-test is one compare of the trace against a threshold bisected on the host;
//   * candidates are staged per lane (entry k of lane l at stage[32 k + l]): no vote, prefix or shared cursor per row, one scan and
//     one reservation in the frame's slot per flush.
// Used when the frames can be described by a tensor map (with or without a pre-existing-feature mask).
#include <cuda.h>

#include "fd_corner_common.cuh"

namespace fdb {

namespace {

using corner::response_of;

constexpr int CT_WARPS = CORNER_TMA_THREADS / 32;
constexpr int CT_GROUP_ROWS = CORNER_TMA_GROUP_ROWS;   // rows per TMA box; a multiple of 3 (the register pipeline's period)
constexpr int CT_SLOTS = 3;                            // ring slots: the group in use, the one before, the one in flight
constexpr int CT_ROW_WORDS = 40;                       // 160-byte box rows
constexpr int CT_LANE_SLOTS = 16;                     // candidate keys each lane can stage between flushes (a row adds at most 4)
constexpr int CT_STAGE = 32 * CT_LANE_SLOTS;           // ... per warp: entry k of lane l sits at stage[32 k + l]
constexpr uint32_t CT_GROUP_BYTES = CT_GROUP_ROWS * CT_ROW_WORDS * 4;
static_assert(CT_GROUP_ROWS % 3 == 0 && CT_GROUP_ROWS >= 3, "group rows must be a multiple of the pipeline period");

struct WarpSmem {
    uint32_t ring[CT_SLOTS * CT_GROUP_ROWS * CT_ROW_WORDS];   // first: TMA destinations need 128-byte alignment
    uint64_t stage[CT_STAGE];
    uint64_t bar[4];
    uint32_t pad[(128 - (CT_SLOTS * CT_GROUP_ROWS * CT_ROW_WORDS * 4 + CT_STAGE * 8 + 32) % 128) % 128 / 4];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "per-warp shared block must keep the rings 128-byte aligned");
static_assert((CT_GROUP_ROWS * CT_ROW_WORDS * 4) % 128 == 0, "every ring slot must start 128-byte aligned");

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
__device__ __forceinline__ void tma_load_rows(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

// One pixel row as this lane sees it: columns c0-2 .. c0+5 as the floats 2^23 + value.
struct MagicRow {
    float v[8];
};
// Horizontal 3-sums of the gradient products at columns c0 .. c0+3 (exact integers in fp32).
struct SumRow {
    float xx[4], yy[4], xy[4];
};

__device__ __forceinline__ float magic(uint32_t word, uint32_t sel) { return __uint_as_float(prmt(word, 0x4B000000u, sel)); }

__device__ __forceinline__ void load_magic_row(MagicRow &r, const uint32_t *row_words, int sh) {
    const uint32_t w0 = row_words[0], w1 = row_words[1], w2 = row_words[2];
    const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh);
    r.v[0] = magic(a, 0x7650u); r.v[1] = magic(a, 0x7651u); r.v[2] = magic(a, 0x7652u); r.v[3] = magic(a, 0x7653u);
    r.v[4] = magic(b, 0x7650u); r.v[5] = magic(b, 0x7651u); r.v[6] = magic(b, 0x7652u); r.v[7] = magic(b, 0x7653u);
}

// Products of gradient row `mid` (row above `up`, row below `dn`) and their horizontal 3-sums (harris.cpp:36-40, 47-62).
// Gradients are integers of magnitude <= 255, so every product and every partial sum is an integer below 2^24: exact in fp32
// whatever the association and whether or not a multiply is fused into the add that consumes it -- the bits are the reference's.
// Per array: neighbouring outputs share a pair sum and the four products used once ride on an FMA (2 mul + 4 fma + 2 add, not 6 + 8).
__device__ __forceinline__ void product_row(SumRow &h, const MagicRow &up, const MagicRow &mid, const MagicRow &dn) {
    float ix[6], iy[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {  // column c0-1+j  <->  v index j+1
        ix[j] = __fsub_rn(mid.v[j + 2], mid.v[j]);       // harris.cpp:36 (the 2^23 of both operands cancels exactly)
        iy[j] = __fsub_rn(dn.v[j + 1], up.v[j + 1]);     // harris.cpp:37
    }
    auto sums = [](const float *a, const float *b, float *out) {   // out[j] = sum over k = j .. j+2 of a[k] * b[k]
        const float p2 = __fmul_rn(a[2], b[2]), p3 = __fmul_rn(a[3], b[3]);
        const float q = __fmaf_rn(a[1], b[1], p2), r = __fmaf_rn(a[4], b[4], p3);
        out[0] = __fmaf_rn(a[0], b[0], q);
        out[1] = __fadd_rn(q, p3);
        out[2] = __fadd_rn(p2, r);
        out[3] = __fmaf_rn(a[5], b[5], r);
    };
    sums(ix, ix, h.xx);
    sums(iy, iy, h.yy);
    sums(ix, iy, h.xy);
}

template <int KIND, bool MASKED>
__global__ void __launch_bounds__(CORNER_TMA_THREADS, 1) corner_tma_kernel(const CornerArgs p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ uint8_t smem_raw[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    uint8_t *smem_al = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_al)[warp];
    uint2 *stage2 = reinterpret_cast<uint2 *>(ws.stage);   // .x = low word of the key, .y = high word
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < CT_SLOTS; ++s) mbar_init(&ws.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const FrameView &fv = p.fv;
    const int col_lo = 2, col_hi = fv.cols - 3;
    uint32_t slot_parity = 0u;   // bit s: parity of the phase ring slot s completes next

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
        const int bx = (x0 - 2) & ~15;                // 16-byte aligned box start (x0 - 2 >= -1, so bx >= -16)
        const int off = (x0 - 2) - bx;                // 0 .. 15: where the strip's first window byte sits in the box row
        const int wofs = (off >> 2) + lane;           // this lane's first ring word
        const int sh = (off & 3) * 8;                 // warp-uniform funnel shift
        // pixel rows rb-3 .. re+2 are needed; local row lr <-> image row rb - 3 + lr
        const int n_steps = re - rb + 6;
        const int n_groups = (n_steps + CT_GROUP_ROWS - 1) / CT_GROUP_ROWS;

        auto issue_group = [&](int g, int s) {
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(&ws.bar[s], CT_GROUP_BYTES);
                tma_load_rows(&ws.ring[s * CT_GROUP_ROWS * CT_ROW_WORDS], &tmap, bx, rb - 3 + g * CT_GROUP_ROWS, frame, &ws.bar[s]);
            }
        };
        auto wait_slot = [&](int s) {
            mbar_wait(&ws.bar[s], (slot_parity >> s) & 1u);
            slot_parity ^= 1u << s;
        };
        issue_group(0, 0);

        bool col_valid[4], col_owned[4];
        float thr_col[4];   // the response threshold per column, +inf where the column has no response (response_of)
        float thr_own[4];   // ... and +inf where the column belongs to the neighbouring strip (the candidate test below)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            col_valid[j] = (c >= col_lo && c <= col_hi);
            col_owned[j] = col_valid[j] && (c >= x0 + 1) && (c <= x0 + CORNER_STRIP_OUT);
            thr_col[j] = col_valid[j] ? p.thr : __int_as_float(0x7f800000);
            thr_own[j] = col_owned[j] ? p.thr : __int_as_float(0x7f800000);
        }

        MagicRow px[3];
        SumRow hs[3];
        float resp[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) px[i].v[j] = 8388608.0f;   // the magic of pixel value 0
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hs[i].xx[j] = hs[i].yy[j] = hs[i].xy[j] = 0.0f;
                resp[i][j] = 0.0f;
            }
        }
        uint32_t *counter = p.cand_counts + frame;
        uint64_t *slot = p.cand_keys + int64_t(frame) * p.cand_capacity;
        // Candidates are staged per LANE (entry k of lane l at stage[32 k + l]): a row's candidates then need no vote, prefix or
        // shared cursor -- a lane that holds one stores it and bumps its own count -- and the warp-wide bookkeeping (one scan, one
        // reservation in the frame's slot) happens once per flush, every dozen rows or so, instead of once per row.
        uint32_t my_staged = 0u;   // this lane's staged keys
        auto flush_stage = [&]() {
            __syncwarp();
            uint32_t incl = my_staged;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t g = 0u;
            if (lane == 0) g = atomicAdd(counter, total);
            g = __shfl_sync(0xffffffffu, g, 0) + incl - my_staged;
            for (uint32_t k = 0; k < my_staged; ++k)
                if (g + k < p.cand_capacity) slot[g + k] = ws.stage[32u * k + uint32_t(lane)];
            __syncwarp();
            my_staged = 0u;
        };
        float *resp_map = p.response_map ? p.response_map + int64_t(frame) * fv.rows * fv.cols : nullptr;
        // pre-existing features: the response is only evaluated where the mask is set (harris.cpp:94), 0 elsewhere
        const uint32_t *mbits = MASKED ? p.mask.bits + int64_t(frame) * fv.rows * p.mask.words_per_row + (c0 >> 5) : nullptr;

        int s_cur = 0;   // ring slot of the group being processed
        for (int g = 0; g < n_groups; ++g) {
            const int s_next = (s_cur == CT_SLOTS - 1) ? 0 : s_cur + 1;
            if (g + 1 < n_groups) issue_group(g + 1, s_next);   // overwrites group g-2: nothing reads it any more
            wait_slot(s_cur);
            const uint32_t *gbase = &ws.ring[s_cur * CT_GROUP_ROWS * CT_ROW_WORDS + wofs];
            const int n0 = rb - 3 + g * CT_GROUP_ROWS;          // image row of the group's first row
            // three rows (one turn of the register pipeline) per trip: small enough to stay in the instruction cache
#pragma unroll 1
            for (int blk = 0; blk < CT_GROUP_ROWS / 3; ++blk) {
            const uint32_t *bbase = gbase + blk * 3 * CT_ROW_WORDS;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int n = n0 + blk * 3 + i;                  // pixel row arriving this step
                if (n < re + 3) {                                // warp-uniform (only the last group is ragged)
                    // slot roles this step: cur = row n, p1 = row n-1, p2 = row n-2 (period 3, compile-time indices)
                    const int cur = i % 3, p1 = (i + 2) % 3, p2 = (i + 1) % 3;
                    load_magic_row(px[cur], bbase + i * CT_ROW_WORDS, sh);
                    // products of row n-1 -> H slot p1 (overwrites row n-4's sums)
                    product_row(hs[p1], px[p2], px[p1], px[cur]);
                    // response of row q = n-2 from H rows n-3 (slot cur), n-2 (slot p2), n-1 (slot p1)
                    const int q = n - 2;
                    const bool q_valid = (q >= p.resp_lo && q <= p.resp_hi);
                    float rq[4];
                    uint32_t mnib = 0xFu;  // mask bits of columns c0 .. c0+3 of row q
                    if (MASKED) {
                        if (q_valid && c0 < fv.cols) {
                            const uint32_t *mp = mbits + int64_t(q) * p.mask.words_per_row;
                            mnib = __funnelshift_r(__ldg(mp), __ldg(mp + 1), c0 & 31) & 0xFu;
                        }
                    }
                    if (q_valid) {   // (warp-uniform: false only for the halo rows at the ends of a band)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float sxx = __fadd_rn(__fadd_rn(hs[cur].xx[j], hs[p2].xx[j]), hs[p1].xx[j]);  // harris.cpp:81-88,108-116
                            const float syy = __fadd_rn(__fadd_rn(hs[cur].yy[j], hs[p2].yy[j]), hs[p1].yy[j]);
                            const float sxy = __fadd_rn(__fadd_rn(hs[cur].xy[j], hs[p2].xy[j]), hs[p1].xy[j]);
                            const float r = response_of<KIND>(sxx, syy, sxy, p, thr_col[j]);
                            rq[j] = (!MASKED || ((mnib >> j) & 1u)) ? r : 0.0f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) rq[j] = 0.0f;
                    }
                    if (resp_map != nullptr && q >= rb && q < re) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col_owned[j]) resp_map[int64_t(q) * fv.cols + c0 + j] = rq[j];
                    }
                    // NMS of row m = n-3.  Response slots before this step's store: cur = row n-4, p2 = row n-3, rq = row n-2.
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
                            // v == 0 means "at or below threshold" (harris.cpp:130); strict 4-neighbour max (:131-132) as one compare against
                            // the largest neighbour (responses are finite, so the four strict compares and this one agree)
                            // ... and the threshold (and whether the column is this strip's at all) is a fifth operand of the same maximum
                            const float big = fmaxf(fmaxf(fmaxf(l, r), resp[cur][j]), fmaxf(rq[j], thr_own[j]));
                            if (v > big) mine |= 1u << j;
                        }
                        if (__any_sync(0xffffffffu, mine != 0u)) {
                            if (mine != 0u) {
                                const uint32_t lo0 = (uint32_t(m + p.tile.row_offset) << 16) | uint32_t(c0);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    if ((mine >> j) & 1u) {
                                        const uint32_t b = __float_as_uint(resp[p2][j]);
                                        const uint32_t hi = b ^ ~(uint32_t(int32_t(b) >> 31) | 0x80000000u);   // ~float_to_ordered(b)
                                        stage2[32u * my_staged + uint32_t(lane)] = make_uint2(lo0 + uint32_t(j), hi);
                                        ++my_staged;
                                    }
                                }
                            }
                            if (__any_sync(0xffffffffu, my_staged > uint32_t(CT_LANE_SLOTS - 4))) flush_stage();
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) resp[p1][j] = rq[j];   // response row n-2 replaces row n-5's slot
                }
            }
            }
            s_cur = s_next;
        }
        if (__any_sync(0xffffffffu, my_staged != 0u)) flush_stage();
        __syncwarp();
    }
}

}  // namespace

size_t corner_tma_smem_bytes() { return size_t(CT_WARPS) * sizeof(WarpSmem) + 128; }

namespace {
template <int KIND, bool MASKED>
cudaError_t launch_corner_tma_t(const CornerArgs &args, const CUtensorMap &map, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(corner_tma_kernel<KIND, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    corner_tma_kernel<KIND, MASKED><<<grid, CORNER_TMA_THREADS, smem, stream>>>(args, map);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_corner_tma(const CornerArgs &args, const void *tensor_map, int grid, cudaStream_t stream) {
    const size_t smem = corner_tma_smem_bytes();
    CUtensorMap map;
    memcpy(&map, tensor_map, sizeof(map));
    const bool masked = args.mask.bits != nullptr;
    if (args.kind == 0) return masked ? launch_corner_tma_t<0, true>(args, map, grid, smem, stream) : launch_corner_tma_t<0, false>(args, map, grid, smem, stream);
    return masked ? launch_corner_tma_t<1, true>(args, map, grid, smem, stream) : launch_corner_tma_t<1, false>(args, map, grid, smem, stream);
}

}  // namespace fdb
