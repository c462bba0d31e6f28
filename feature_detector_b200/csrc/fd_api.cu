// C-ABI layer: context, device buffers, launch configuration.  See include/fd_b200.h for the contract.
// There is deliberately no CPU implementation behind any entry point.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>

#include "../../include/fd_b200.h"
#include "fd_internal.h"
#include "fd_kernels.cuh"

using namespace fdb;

struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    void *raw = nullptr;       // guard mode (FD_B200_GUARD=1): the allocation, red zones included; ptr = raw + GUARD_BYTES
    const char *name = "";
};

struct fd_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    bool guard = false;                // FD_B200_GUARD=1: every context-owned buffer is exactly as large as asked for and sits between
    std::vector<DevBuf *> guarded;     // two red zones that fd_debug_check_guards verifies (compute-sanitizer is not always available)

    FrameView fv = {};
    bool frames_bound = false;
    DevBuf owned_frames;

    DevBuf lut;       // 65536 B
    DevBuf segs;      // OffsetSeg table
    std::vector<OffsetSeg> host_segs;
    int n_seg = 0;
    uint32_t seg_count_covered = 0;

    DevBuf keys, keys_scratch, counts, flags, cells, alive, kept, pre_hist, pre_keys;
    uint32_t cand_capacity = 0;
    bool have_candidates = false, candidates_sorted = false;

    DevBuf kp, kp_counts;
    int kp_capacity = 0;
    bool have_keypoints = false;
    int select_frames = 0;     // frames the last selection covered (the bound frames, or external candidates)
    bool kp_of_bound_frames = false;   // the keypoints were selected on the frames that are bound now (not on row tiles, gathered keys or heat maps)

    DevBuf user_kp, user_counts;
    int user_capacity = 0;

    DevBuf desc;
    int desc_capacity = 0;     // slots per frame of the described set
    int desc_length = 0;       // kLength of the last description
    const int32_t *desc_counts = nullptr;   // device counts of the described set
    bool desc_from_user = false;
    bool have_desc = false;

    DevBuf mask_bits, mask_rowbase, mask_prefix, existing_xy, existing_counts;
    int existing_capacity = 0, existing_frames = 0;
    bool have_existing = false;
    MaskView mask_view = {};

    float *resp_map = nullptr;
    uint8_t *score_map = nullptr;

    bool tiled = false;        // fd_set_tile: the bound frames are a row tile of a taller image
    TileView tile = {};

    // TMA view of the bound frames for the sparse FAST kernel (rebuilt when the binding changes)
    CUtensorMap frame_map, corner_map;   // [0] box 160 x FAST_SPARSE_GROUP_ROWS (sparse FAST), [1] box 160 x CORNER_TMA_GROUP_ROWS
    bool frame_map_valid = false, frame_map_failed = false, corner_map_valid = false, corner_map_failed = false;
    bool force_stream_corner = false;  // FD_B200_CORNER_STREAM=1: testing knob, always take the register-streaming corner kernel
    int items_per_warp = 8;    // FD_B200_ITEMS_PER_WARP: tuning knob, work items each resident warp should get (band height follows)
    uint32_t select_cells_min = SELECT_CELLS_MIN;   // FD_B200_SELECT_CELLS_MIN: testing knob, candidate count above which selection runs its rounds per cell
    bool force_dense_fast = false;  // FD_B200_FAST_DENSE=1: testing knob, always take the dense kernel
    int select_smem_kb = 200;       // FD_B200_SELECT_SMEM_KB: testing knob, the most shared memory a frame's cell state may take when a CTA has its SM to itself
    bool select_prepare = true;     // FD_B200_SELECT_PREPARE=0: testing knob, selection always builds its rank histogram and first range itself

    void *host_stage = nullptr;     // pinned staging block of fd_detect_describe_host
    size_t host_stage_bytes = 0;
    DevBuf result_pack;             // device image of that block when the results are small enough to leave in one copy
    DevBuf nn_desc, nn_user_desc, desc_float, lsd_work, float_slot[2], matches;
    int match_pairs = 0, match_capacity = 0;
    bool have_desc_float = false;
    int nn_channels = 0;
    bool have_nn_desc = false;
    DevBuf lsd_norm, lsd_angle, lsd_keys, lsd_counts, lsd_sorted, lsd_hist, lsd_start, lsd_bucketed, lsd_item_counts, lsd_chunk_sum;
    size_t lsd_hist_zeroed = 0;   // bytes of lsd_hist known to be zero (the scatter kernel restores the zeros it consumes)
    float *lsd_norm_p = nullptr, *lsd_angle_p = nullptr;
    int32_t *lsd_sorted_p = nullptr, *lsd_nvalid_p = nullptr;
    bool have_lsd = false, lsd_sorted_valid = false;
};

namespace {

fd_status fail(fd_context *ctx, fd_status st, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return st;
}

#define FD_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? FD_ERR_OUT_OF_MEMORY : FD_ERR_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                       \
        }                                                                                           \
    } while (0)

#define FD_TRY(expr)                      \
    do {                                  \
        fd_status s__ = (expr);           \
        if (s__ != FD_OK) return s__;     \
    } while (0)

constexpr size_t GUARD_BYTES = 256;      // keeps ptr as aligned as cudaMalloc's result for every access width the kernels use
constexpr int GUARD_PATTERN = 0xA5;

// Guard mode: compare the two red zones of one buffer with the pattern (the caller has synchronised the stream).
fd_status check_red_zones(fd_context *ctx, const DevBuf &b) {
    if (!b.raw) return FD_OK;
    uint8_t zone[2 * GUARD_BYTES];
    const uint8_t *raw = static_cast<const uint8_t *>(b.raw);
    FD_CUDA(ctx, cudaMemcpy(zone, raw, GUARD_BYTES, cudaMemcpyDeviceToHost));
    FD_CUDA(ctx, cudaMemcpy(zone + GUARD_BYTES, raw + GUARD_BYTES + b.bytes, GUARD_BYTES, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 2 * GUARD_BYTES; ++i) {
        if (zone[i] == uint8_t(GUARD_PATTERN)) continue;
        const bool front = i < GUARD_BYTES;
        return fail(ctx, FD_ERR_CUDA, std::string("red zone ") + (front ? "before " : "after ") + b.name + " (" + std::to_string(b.bytes) +
                                          " bytes) overwritten at byte " + (front ? "-" + std::to_string(GUARD_BYTES - i) : "+" + std::to_string(i - GUARD_BYTES)));
    }
    return FD_OK;
}

fd_status reserve_named(fd_context *ctx, DevBuf &b, size_t bytes, const char *name) {
    if (bytes == 0) bytes = 16;
    // guard mode never reuses a larger buffer: the red zone has to start where this call's bytes end
    if (b.ptr != nullptr && (ctx->guard ? bytes == b.bytes : bytes <= b.bytes)) return FD_OK;
    if (b.ptr != nullptr) {
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        FD_TRY(check_red_zones(ctx, b));   // a buffer leaves guard mode's books only after its zones were looked at
        FD_CUDA(ctx, cudaFree(b.raw ? b.raw : b.ptr));
        b.ptr = b.raw = nullptr;
        b.bytes = 0;
    }
    if (ctx->guard) {
        FD_CUDA(ctx, cudaMalloc(&b.raw, bytes + 2 * GUARD_BYTES));
        uint8_t *raw = static_cast<uint8_t *>(b.raw);
        FD_CUDA(ctx, cudaMemsetAsync(raw, GUARD_PATTERN, GUARD_BYTES, ctx->stream));
        FD_CUDA(ctx, cudaMemsetAsync(raw + GUARD_BYTES + bytes, GUARD_PATTERN, GUARD_BYTES, ctx->stream));
        b.ptr = raw + GUARD_BYTES;
        b.name = name;
        if (std::find(ctx->guarded.begin(), ctx->guarded.end(), &b) == ctx->guarded.end()) ctx->guarded.push_back(&b);
    } else {
        FD_CUDA(ctx, cudaMalloc(&b.ptr, bytes));
    }
    b.bytes = bytes;
    return FD_OK;
}
#define reserve(ctx, b, ...) reserve_named(ctx, b, (__VA_ARGS__), #b)

void release(DevBuf &b) {
    if (b.ptr) cudaFree(b.raw ? b.raw : b.ptr);
    b.ptr = b.raw = nullptr;
    b.bytes = 0;
}

// Longest circular run of set bits in a 16-bit ring mask (feature_point_fast_detector.cpp:55-78).
std::vector<uint8_t> build_run_lut() {
    std::vector<uint8_t> lut(65536);
    for (uint32_t m = 0; m < 65536; ++m) {
        if (m == 0xFFFFu) {
            lut[m] = 16;
            continue;
        }
        int best = 0, run = 0;
        for (int i = 0; i < 32; ++i) {
            if ((m >> (i & 15)) & 1u) {
                ++run;
                best = std::max(best, run);
            } else {
                run = 0;
            }
        }
        lut[m] = uint8_t(std::min(best, 16));
    }
    return lut;
}

// Piecewise-linear description of the reference's running offset (fast.cpp:85,93):
//   offset_0 = 1e-5f;  offset_{k+1} = fl(offset_k + 1e-5f).
// Inside one binade every add rounds to the same number of ulps, so the float's BIT PATTERN advances by a
// constant per step; a new piece starts wherever that constant changes.  Built by running the same float
// loop once (volatile keeps the host compiler from using wider intermediates).
std::vector<OffsetSeg> build_offset_table(uint32_t count) {
    std::vector<OffsetSeg> segs;
    volatile float offset = 1e-5f;
    const volatile float inc = 1e-5f;
    auto bits = [](float f) {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return u;
    };
    uint32_t prev_bits = bits(offset);
    segs.push_back({0u, prev_bits, 0u});
    bool step_known = false;
    for (uint32_t k = 1; k < count; ++k) {
        offset = offset + inc;
        const uint32_t b = bits(offset);
        const uint32_t step = b - prev_bits;
        OffsetSeg &cur = segs.back();
        if (!step_known) {
            cur.step = step;
            step_known = true;
        } else if (step != cur.step) {
            // the value at k is no longer on the current line: open a new piece starting AT k
            segs.push_back({k, b, 0u});
            step_known = false;
        }
        prev_bits = b;
    }
    segs.push_back({0xFFFFFFFFu, 0u, 0u});  // sentinel
    return segs;
}

// Offset bit pattern at pixel index k from the host copy of the table.
uint32_t host_offset_bits(const std::vector<OffsetSeg> &segs, uint32_t k) {
    size_t lo = 0, hi = segs.size() - 2;
    while (lo < hi) {
        const size_t mid = (lo + hi + 1) / 2;
        if (segs[mid].k_start <= k) lo = mid; else hi = mid - 1;
    }
    return segs[lo].bits_start + (k - segs[lo].k_start) * segs[lo].step;
}

// kmin[s] = smallest pixel index k in [0, count) with fl(float(s) + offset(k)) > thr, else 0xFFFFFFFF.
// fl(s + offset) is non-decreasing in k because the offset is, so a binary search finds it.
void build_kmin(const std::vector<OffsetSeg> &segs, uint32_t count, float thr, uint32_t kmin[17]) {
    for (int s = 0; s <= 16; ++s) {
        auto passes = [&](uint32_t k) {
            const uint32_t b = host_offset_bits(segs, k);
            float off;
            std::memcpy(&off, &b, 4);
            volatile float v = float(s) + off;  // one rounding, like fast.cpp:89
            return v > thr;
        };
        if (count == 0 || !passes(count - 1)) {
            kmin[s] = 0xFFFFFFFFu;
            continue;
        }
        uint32_t lo = 0, hi = count - 1;  // passes(hi) holds
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (passes(mid)) hi = mid; else lo = mid + 1;
        }
        kmin[s] = lo;
    }
}

fd_status ensure_fast_tables(fd_context *ctx, uint32_t count) {
    if (ctx->lut.ptr == nullptr) {
        const std::vector<uint8_t> lut = build_run_lut();
        FD_TRY(reserve(ctx, ctx->lut, 65536));
        FD_CUDA(ctx, cudaMemcpyAsync(ctx->lut.ptr, lut.data(), 65536, cudaMemcpyHostToDevice, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (count > ctx->seg_count_covered || ctx->segs.ptr == nullptr) {
        const uint32_t want = std::max<uint32_t>(count, 1u << 20);
        const std::vector<OffsetSeg> segs = build_offset_table(want);
        FD_TRY(reserve(ctx, ctx->segs, segs.size() * sizeof(OffsetSeg)));
        FD_CUDA(ctx, cudaMemcpyAsync(ctx->segs.ptr, segs.data(), segs.size() * sizeof(OffsetSeg), cudaMemcpyHostToDevice, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (segs.size() > size_t(FAST_MAX_SEGS)) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "FAST offset table has too many pieces for this frame size");
        ctx->n_seg = int(segs.size()) - 1;
        ctx->host_segs = segs;
        ctx->seg_count_covered = want;
    }
    return FD_OK;
}

// 3-D TMA map (cols x rows x frames, u8) with a 160 x 16 x 1 box for the sparse FAST kernel.  Needs 16-byte aligned base,
// pitch and frame stride; returns false (and the dense kernel is used) when the layout or the driver does not allow it.
bool ensure_frame_map(fd_context *ctx, bool for_corner = false) {
    bool &valid = for_corner ? ctx->corner_map_valid : ctx->frame_map_valid;
    bool &failed = for_corner ? ctx->corner_map_failed : ctx->frame_map_failed;
    CUtensorMap *map = for_corner ? &ctx->corner_map : &ctx->frame_map;
    if (valid) return true;
    if (failed) return false;
    const FrameView &fv = ctx->fv;
    failed = true;
    if (reinterpret_cast<uintptr_t>(fv.data) % 16 != 0 || fv.pitch % 16 != 0 || fv.frame_stride % 16 != 0) return false;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (encode == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return false;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[3] = {cuuint64_t(fv.cols), cuuint64_t(fv.rows), cuuint64_t(fv.n_frames)};
    const cuuint64_t strides[2] = {cuuint64_t(fv.pitch), cuuint64_t(fv.frame_stride)};
    const cuuint32_t box[3] = {160u, cuuint32_t(for_corner ? CORNER_TMA_GROUP_ROWS : FAST_SPARSE_GROUP_ROWS), 1u};  // box starts are 16-byte aligned
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(fv.data), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    failed = false;
    valid = true;
    return true;
}

// harris.cpp:98 tests (trace * trace * 0.21f * inv_cnt2) > thr, i.e. fl(fl(fl(trace * trace) * 0.21f) * inv2) > thr with inv2 = fl(fl(1/9)^2)
// (harris.cpp:71-72): three roundings of a non-negative trace (a sum of squares), each monotone non-decreasing, so the passing traces are
// an upper interval of the floats.  Its lower end is found by bisection over the bit patterns of the non-negative floats (which order like
// the floats); NaN when no trace passes (thr is +inf or NaN: trace >= NaN is false).
float harris_trace_min(float thr) {
    volatile float nine = 9.0f;
    volatile float inv = 1.0f / nine;
    volatile float inv2 = inv * inv;
    auto passes = [&](uint32_t bits) {
        float trace;
        std::memcpy(&trace, &bits, 4);
        volatile float t0 = trace * trace;
        volatile float t1 = t0 * 0.21f;
        volatile float t2 = t1 * inv2;
        return t2 > thr;
    };
    uint32_t lo = 0u, hi = 0x7F800000u;   // +0 .. +inf
    if (!passes(hi)) return std::nanf("");
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (passes(mid)) hi = mid;
        else lo = mid + 1;
    }
    float out;
    std::memcpy(&out, &lo, 4);
    return out;
}

// Split the interior rows into bands so that every resident warp gets several work items.
void plan_bands(const fd_context *ctx, int interior_rows, int n_strips, int n_frames, int warps_per_cta, int ctas_per_sm, int min_band,
                int band_multiple, int &band_rows, int &n_bands, int64_t &n_items, int &grid) {
    grid = ctx->sm_count * ctas_per_sm;
    const int64_t total_warps = int64_t(grid) * warps_per_cta;
    const int64_t base_items = int64_t(n_frames) * n_strips;
    int64_t want_bands = (ctx->items_per_warp * total_warps + base_items - 1) / std::max<int64_t>(base_items, 1);
    int64_t max_bands = std::max(1, interior_rows / min_band);
    // A few frames (the drop-in classes' one frame per call) cannot fill the grid with bands that tall.  What counts there is the
    // latency of one band on one warp -- a warp alone on its scheduler retires an instruction every few cycles -- not the halo rows
    // shorter bands read twice: one item per warp, bands down to four rows.
    if (base_items * max_bands < total_warps) {
        max_bands = std::max(1, interior_rows / 4);
        want_bands = (total_warps + base_items - 1) / std::max<int64_t>(base_items, 1);
    }
    want_bands = std::max<int64_t>(1, std::min<int64_t>(want_bands, max_bands));
    band_rows = int((interior_rows + want_bands - 1) / want_bands);
    band_rows = std::max(band_rows, 1);
    band_rows = (band_rows + band_multiple - 1) / band_multiple * band_multiple;  // kernels that unroll their row loop
    n_bands = (interior_rows + band_rows - 1) / band_rows;
    n_items = base_items * n_bands;
    // Warps take items from a shared counter, so fewer items than warps still spread over the SMs when every SM gets a CTA: one
    // frame's few hundred items finish in the time of one item, not of a CTA's worth of them on a handful of SMs.
    grid = int(std::max<int64_t>(1, std::min<int64_t>(grid, n_items)));
}

fd_status require_frames(fd_context *ctx) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->frames_bound) return fail(ctx, FD_ERR_NOT_READY, "no frames bound: call fd_upload_frames or fd_bind_device_frames first");
    return FD_OK;
}

fd_status check_params(fd_context *ctx, const fd_detect_params *p) {
    if (!p) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "params is null");
    if (p->kind < FD_HARRIS || p->kind > FD_FAST) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "unknown detector kind");
    if (p->fast_min_pixel_diff < 0 || p->fast_min_pixel_diff > 255) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fast_min_pixel_diff out of [0,255]");
    if (p->reserved != 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "reserved must be 0");
    return FD_OK;
}

fd_status run_candidates(fd_context *ctx, const fd_detect_params *p, int cand_capacity) {
    FD_TRY(require_frames(ctx));
    FD_TRY(check_params(ctx, p));
    const FrameView &fv = ctx->fv;
    const int64_t px = int64_t(fv.rows) * fv.cols;
    TileView tile = {0, 0, fv.rows, fv.rows};
    if (ctx->tiled) {
        tile = ctx->tile;
        if (tile.own_lo < 0 || tile.own_hi > fv.rows || tile.own_lo > tile.own_hi || tile.row_offset < 0 || tile.row_offset + fv.rows > tile.full_rows)
            return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_set_tile: the tile does not fit the bound frames");
        if (ctx->have_existing) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "pre-existing feature masks are not supported on row tiles");
    }
    if (tile.full_rows > 65535 || fv.cols > 65535) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "frames are limited to 65535 x 65535");
    const uint32_t cap = cand_capacity > 0 ? uint32_t(std::min<int64_t>(cand_capacity, px)) : uint32_t(px);
    ctx->cand_capacity = cap;
    FD_TRY(reserve(ctx, ctx->keys, size_t(fv.n_frames) * cap * 8));
    FD_TRY(reserve(ctx, ctx->counts, size_t(fv.n_frames) * 4));
    FD_TRY(reserve(ctx, ctx->flags, 16));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->counts.ptr, 0, size_t(fv.n_frames) * 4, ctx->stream));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->flags.ptr, 0, 16, ctx->stream));
    ctx->have_candidates = false;
    ctx->candidates_sorted = false;
    ctx->have_keypoints = false;

    // pre-existing features: rasterise the bit mask for this call's min distance (feature_point_detector.cpp:12-16,90-98)
    MaskView mask = {};
    if (ctx->have_existing) {
        if (ctx->existing_frames != fv.n_frames) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "existing features were set for a different number of frames");
        const int wpr = (fv.cols + 31) / 32 + 1;
        const size_t words = size_t(fv.n_frames) * fv.rows * wpr;
        FD_TRY(reserve(ctx, ctx->mask_bits, words * 4));
        MaskArgs m = {};
        m.rows = fv.rows;
        m.cols = fv.cols;
        m.n_frames = fv.n_frames;
        m.min_distance = p->min_feature_distance;
        m.xy = static_cast<const float *>(ctx->existing_xy.ptr);
        m.counts = static_cast<const int32_t *>(ctx->existing_counts.ptr);
        m.capacity = ctx->existing_capacity;
        m.bits = static_cast<uint32_t *>(ctx->mask_bits.ptr);
        m.words_per_row = wpr;
        if (p->kind == FD_FAST) {
            FD_TRY(reserve(ctx, ctx->mask_prefix, words * 4));
            FD_TRY(reserve(ctx, ctx->mask_rowbase, size_t(fv.n_frames) * (fv.rows + 1) * 4));
            m.word_prefix = static_cast<uint32_t *>(ctx->mask_prefix.ptr);
            m.row_base = static_cast<uint32_t *>(ctx->mask_rowbase.ptr);
        }
        FD_CUDA(ctx, launch_mask(m, ctx->stream));
        ++ctx->launches;
        mask.bits = m.bits;
        mask.word_prefix = m.word_prefix;
        mask.row_base = m.row_base;
        mask.words_per_row = wpr;
    }
    ctx->mask_view = mask;

    if (p->kind == FD_FAST) {
        if (ctx->score_map) FD_CUDA(ctx, cudaMemsetAsync(ctx->score_map, 0, size_t(fv.n_frames) * px, ctx->stream));
        if (fv.rows >= 7 && fv.cols >= 7) {
            const uint64_t interior64 = uint64_t(tile.full_rows - 6) * uint64_t(fv.cols - 6);
            if (interior64 > 0xFFFFFFF0ull) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "FAST: the frame has more interior pixels than the 32-bit pixel index of the running offset holds");
            const uint32_t interior = uint32_t(interior64);
            FD_TRY(ensure_fast_tables(ctx, interior));
            FastArgs a = {};
            a.fv = fv;
            a.diff = p->fast_min_pixel_diff;
            a.thr = p->min_valid_response;
            a.lut = static_cast<const uint8_t *>(ctx->lut.ptr);
            a.segs = static_cast<const OffsetSeg *>(ctx->segs.ptr);
            a.n_seg = ctx->n_seg;
            build_kmin(ctx->host_segs, interior, p->min_valid_response, a.kmin);
            a.cand_keys = static_cast<uint64_t *>(ctx->keys.ptr);
            a.cand_counts = static_cast<uint32_t *>(ctx->counts.ptr);
            a.cand_capacity = cap;
            a.score_map = ctx->score_map;
            a.score_aligned = (fv.cols % 4 == 0) && (reinterpret_cast<uintptr_t>(ctx->score_map) % 4 == 0);
            a.n_strips = (fv.cols + 127) / 128;
            a.mask = mask;
            a.tile = tile;
            a.proc_lo = std::max(3, tile.own_lo);
            a.proc_hi = std::max(a.proc_lo, std::min(std::min(fv.rows - 3, tile.own_hi), tile.full_rows - 3 - tile.row_offset));
            const int proc_rows = a.proc_hi - a.proc_lo;
            // Sparse form: when the threshold leaves s_min >= 4 (>= 1 with the pre-check: a failed pre-check scores 0) at
            // every pixel of the frame, most words are ruled out by the compass test and only the rest are scored.
            const bool precheck = p->fast_n >= 12;
            int shift = 0;
            while ((2 << shift) <= p->fast_min_pixel_diff + 1) ++shift;   // 2^shift = largest power of two <= diff + 1
            a.absdiff_shift = shift;
            a.absdiff_mask = ((0xFFu << shift) & 0xFFu) * 0x01010101u;
            const bool prunable = precheck ? (a.kmin[0] == 0xFFFFFFFFu && shift >= 1) : (a.kmin[3] == 0xFFFFFFFFu);
            const bool sparse = prunable && proc_rows > 0 && !ctx->force_dense_fast && ctx->score_map == nullptr && ensure_frame_map(ctx);
            int grid;
            a.work_counter = static_cast<uint32_t *>(ctx->flags.ptr) + 1;   // flags[1]: zeroed with the overflow flag at the top of this call
            if (sparse) {
                plan_bands(ctx, proc_rows, a.n_strips, fv.n_frames, FAST_SPARSE_THREADS / 32, 1, 58, 1, a.band_rows, a.n_bands, a.n_items, grid);
                if (a.band_rows > 2032) {  // the kernel's queue entries keep the band-local row in 11 bits
                    a.band_rows = 2032;
                    a.n_bands = (proc_rows + a.band_rows - 1) / a.band_rows;
                    a.n_items = int64_t(fv.n_frames) * a.n_strips * a.n_bands;
                }
                FD_CUDA(ctx, launch_fast_sparse(a, &ctx->frame_map, precheck, grid, ctx->stream));
            } else {
                plan_bands(ctx, proc_rows, a.n_strips, fv.n_frames, FAST_THREADS / 32, FAST_CTAS_PER_SM, 14, 1, a.band_rows, a.n_bands, a.n_items, grid);
                FD_CUDA(ctx, launch_fast(a, precheck, grid, ctx->stream));
            }
            ++ctx->launches;
        }
    } else {
        if (ctx->resp_map) FD_CUDA(ctx, cudaMemsetAsync(ctx->resp_map, 0, size_t(fv.n_frames) * px * 4, ctx->stream));
        if (fv.rows >= 5 && fv.cols >= 5) {
            CornerArgs a = {};
            a.fv = fv;
            a.kind = p->kind;
            a.thr = p->min_valid_response;
            a.alpha = p->harris_alpha;
            volatile float nine = 9.0f;
            volatile float inv = 1.0f / nine;                 // harris.cpp:71
            volatile float inv2 = inv * inv;                  // harris.cpp:72
            a.inv_cnt = inv;
            a.inv_cnt2 = inv2;
            a.harris_trace_min = harris_trace_min(a.thr);
            a.cand_keys = static_cast<uint64_t *>(ctx->keys.ptr);
            a.cand_counts = static_cast<uint32_t *>(ctx->counts.ptr);
            a.cand_capacity = cap;
            a.response_map = ctx->resp_map;
            a.n_strips = (fv.cols - 4 + CORNER_STRIP_OUT - 1) / CORNER_STRIP_OUT;
            a.mask = mask;
            a.tile = tile;
            // a response exists where the 5x5 neighbourhood is inside the FULL frame (harris.cpp:90-92) and inside this buffer
            a.resp_lo = std::max(2, 2 - tile.row_offset);
            a.resp_hi = std::min(fv.rows - 3, tile.full_rows - 3 - tile.row_offset);
            a.cand_lo = std::max(a.resp_lo, tile.own_lo);
            a.cand_hi = std::max(a.cand_lo, std::min(a.resp_hi + 1, tile.own_hi));
            int grid;
            a.work_counter = static_cast<uint32_t *>(ctx->flags.ptr) + 2;   // flags[2]: zeroed at the top of this call
            if (!ctx->force_stream_corner && a.cand_hi > a.cand_lo && ensure_frame_map(ctx, true)) {
                plan_bands(ctx, a.cand_hi - a.cand_lo, a.n_strips, fv.n_frames, CORNER_TMA_THREADS / 32, 1, 42, 1, a.band_rows, a.n_bands, a.n_items, grid);
                FD_CUDA(ctx, launch_corner_tma(a, &ctx->corner_map, grid, ctx->stream));
            } else {
                plan_bands(ctx, a.cand_hi - a.cand_lo, a.n_strips, fv.n_frames, CORNER_THREADS / 32, 2, 16, 1, a.band_rows, a.n_bands, a.n_items, grid);
                FD_CUDA(ctx, launch_corner(a, grid, ctx->stream));
            }
            ++ctx->launches;
        }
    }
    ctx->have_candidates = true;
    return FD_OK;
}

// Greedy selection over candidate keys of n_frames frames of rows x cols pixels (the context's own candidates, or keys
// gathered from the row tiles of one frame).
struct SelectExtras {   // row tiles (fd_tiled.cu): see fd_internal.h
    const fd_select_prefilter *first_range = nullptr;   // run on the gathered first rank ranges only, flag what needs more
    const uint8_t *only_flagged = nullptr;               // run for flagged frames only (and never through the preparation kernels)
};

fd_status run_select(fd_context *ctx, const fd_detect_params *p, int rows, int cols, int n_frames, uint64_t *keys, const uint32_t *counts,
                     uint32_t capacity, uint32_t xy_xor = 0u, const SelectExtras *ex = nullptr) {
    struct { int rows, cols, n_frames; } fv = {rows, cols, n_frames};
    const int kp_cap = int(std::max<uint32_t>(1u, std::min<uint32_t>(p->needed_feature_num, 1u << 20)));
    ctx->kp_capacity = kp_cap;
    FD_TRY(reserve(ctx, ctx->kp, size_t(fv.n_frames) * kp_cap * sizeof(float4)));
    FD_TRY(reserve(ctx, ctx->kp_counts, size_t(fv.n_frames) * 4));
    SelectArgs a = {};
    a.rows = fv.rows;
    a.cols = fv.cols;
    a.n_frames = fv.n_frames;
    a.cand_keys = keys;
    a.cand_counts = counts;
    a.cand_capacity = capacity;
    a.min_distance = p->min_feature_distance;
    a.needed = p->needed_feature_num;
    a.cells_min = ctx->select_cells_min;
    a.existing_counts = ctx->have_existing ? static_cast<const int32_t *>(ctx->existing_counts.ptr) : nullptr;
    a.keypoints = static_cast<float4 *>(ctx->kp.ptr);
    a.kp_counts = static_cast<int32_t *>(ctx->kp_counts.ptr);
    a.kp_capacity = kp_cap;
    const int cell = std::max(p->min_feature_distance, 0) + 1;
    a.cells_x = (fv.cols + cell - 1) / cell;
    a.cells_y = (fv.rows + cell - 1) / cell;
    const size_t cell_bytes = select_cell_bytes(a.cells_x, a.cells_y);   // the grid carries a one-cell empty border
    // ceil(2^32 / cell) does not fit 32 bits for cells of one pixel (min distance 0): 0 stands for the identity there
    a.cell_magic = cell == 1 ? 0u : uint32_t(((uint64_t(1) << 32) + cell - 1) / uint64_t(cell));
    // The per-cell state lives in shared memory when it fits beside three other frames' CTAs on the SM -- or, with no more frames than
    // SMs (a CTA has its SM to itself), whenever it fits at all: a 1920x1080 frame at d = 20 has 5 000 cells, 100 KB.
    a.cells_in_smem = cell_bytes <= size_t(fv.n_frames <= ctx->sm_count ? ctx->select_smem_kb : 48) * 1024;
    if (!a.cells_in_smem) {
        FD_TRY(reserve(ctx, ctx->cells, cell_bytes * fv.n_frames));
        a.cell_scratch = static_cast<uint32_t *>(ctx->cells.ptr);
        a.cell_stride = int64_t(cell_bytes);
    }
    a.kept_capacity = a.cells_x * a.cells_y;   // at most one kept point per cell
    a.few_frames = fv.n_frames <= ctx->sm_count;
    FD_TRY(reserve(ctx, ctx->alive, size_t(fv.n_frames) * capacity * 16));
    FD_TRY(reserve(ctx, ctx->kept, size_t(fv.n_frames) * a.kept_capacity * 8));
    a.live_scratch = static_cast<uint64_t *>(ctx->alive.ptr);
    a.kept_keys = static_cast<uint64_t *>(ctx->kept.ptr);
    a.overflow_flag = static_cast<uint32_t *>(ctx->flags.ptr);
    a.mask = ctx->mask_view;
    a.xy_xor = xy_xor;
    // Few frames whose slots admit many candidates (a 3840x2160 frame, the gathered tiles of one): one CTA per frame would stream over all
    // of a frame's keys twice while most SMs idle, so the rank histogram and the first rank range are prepared by many CTAs per frame.
    bool prepare = fv.n_frames < 2 * ctx->sm_count && capacity > uint32_t(SELECT_CELLS_MIN) && ctx->select_prepare;   // (two more launches: not worth it for slots a CTA streams in a few trips)
    if (ex != nullptr) {
        prepare = false;
        a.only_flagged = ex->only_flagged;
        if (ex->first_range != nullptr) {
            a.pre_hist = ex->first_range->hist;
            a.pre_keys = ex->first_range->pre_keys;
            a.pre_counts = ex->first_range->pre_counts;
            a.pre_capacity = ex->first_range->pre_capacity;
            a.need_more = ex->first_range->need_more;
        }
    }
    if (prepare) {
        const size_t hist_bytes = size_t(fv.n_frames) * 2048 * 4;
        FD_TRY(reserve(ctx, ctx->pre_hist, hist_bytes + size_t(fv.n_frames) * 4));
        FD_TRY(reserve(ctx, ctx->pre_keys, size_t(fv.n_frames) * capacity * 8));
        FD_CUDA(ctx, cudaMemsetAsync(ctx->pre_hist.ptr, 0, hist_bytes + size_t(fv.n_frames) * 4, ctx->stream));
        a.pre_hist = static_cast<uint32_t *>(ctx->pre_hist.ptr);
        a.pre_counts = a.pre_hist + size_t(fv.n_frames) * 2048;
        a.pre_keys = static_cast<uint64_t *>(ctx->pre_keys.ptr);
        a.pre_capacity = capacity;
    }
    FD_CUDA(ctx, launch_select(a, ctx->stream));
    ctx->launches += ((a.cand_capacity > a.cells_min) ? 2 : 1) + (prepare ? 2 : 0);   // the per-cell form is launched only when the capacity admits it
    ctx->candidates_sorted = false;  // selection needs no global sort; fd_download_candidates orders its copy
    ctx->have_keypoints = true;
    ctx->select_frames = fv.n_frames;
    ctx->kp_of_bound_frames = false;   // fd_detect sets it after its own selection
    return FD_OK;
}

fd_status check_overflow(fd_context *ctx) {
    if (ctx->flags.ptr == nullptr) return FD_OK;
    uint32_t flag = 0;
    FD_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->flags.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag != 0) return fail(ctx, FD_ERR_CAPACITY, "a frame produced more candidates than cand_capacity; raise it and run again");
    return FD_OK;
}

}  // namespace

extern "C" {

const char *fd_version(void) { return "feature_detector_b200 0.1 (sm_100a)"; }

fd_status fd_create(int device_ordinal, fd_context **out_ctx) {
    if (!out_ctx) return FD_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return FD_ERR_NO_DEVICE;
    if (device_ordinal < 0 || device_ordinal >= n) return FD_ERR_INVALID_ARGUMENT;
    fd_context *ctx = new (std::nothrow) fd_context();
    if (!ctx) return FD_ERR_OUT_OF_MEMORY;
    ctx->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FD_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (const char *env = std::getenv("FD_B200_GUARD")) ctx->guard = (env[0] == '1');
    if (const char *env = std::getenv("FD_B200_FAST_DENSE")) ctx->force_dense_fast = (env[0] == '1');
    if (const char *env = std::getenv("FD_B200_CORNER_STREAM")) ctx->force_stream_corner = (env[0] == '1');
    if (const char *env = std::getenv("FD_B200_ITEMS_PER_WARP")) ctx->items_per_warp = std::max(1, atoi(env));
    if (const char *env = std::getenv("FD_B200_SELECT_PREPARE")) ctx->select_prepare = (env[0] != '0');
    if (const char *env = std::getenv("FD_B200_SELECT_SMEM_KB")) ctx->select_smem_kb = std::max(0, std::min(200, std::atoi(env)));
    if (const char *env = std::getenv("FD_B200_SELECT_CELLS_MIN")) ctx->select_cells_min = uint32_t(std::strtoul(env, nullptr, 10));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    *out_ctx = ctx;
    return FD_OK;
}

fd_status fd_destroy(fd_context *ctx) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->owned_frames, &ctx->lut, &ctx->segs, &ctx->keys, &ctx->keys_scratch, &ctx->counts, &ctx->flags, &ctx->cells, &ctx->alive, &ctx->kept, &ctx->pre_hist, &ctx->pre_keys, &ctx->kp,
                      &ctx->kp_counts, &ctx->user_kp, &ctx->user_counts, &ctx->desc, &ctx->mask_bits, &ctx->mask_rowbase, &ctx->mask_prefix,
                      &ctx->existing_xy, &ctx->existing_counts, &ctx->lsd_norm, &ctx->lsd_angle, &ctx->lsd_keys, &ctx->lsd_counts, &ctx->lsd_sorted, &ctx->lsd_hist, &ctx->lsd_start, &ctx->lsd_bucketed, &ctx->lsd_item_counts, &ctx->lsd_chunk_sum, &ctx->nn_desc, &ctx->nn_user_desc, &ctx->desc_float, &ctx->lsd_work, &ctx->float_slot[0], &ctx->float_slot[1], &ctx->matches, &ctx->result_pack})
        release(*b);
    if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return FD_OK;
}

const char *fd_last_error(const fd_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

fd_status fd_set_stream(fd_context *ctx, void *cuda_stream) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return FD_OK;
}

void *fd_own_stream(const fd_context *ctx) { return ctx ? static_cast<void *>(ctx->own_stream) : nullptr; }

fd_status fd_sync(fd_context *ctx) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->have_candidates) return check_overflow(ctx);
    return FD_OK;
}

uint64_t fd_launch_count(const fd_context *ctx) { return ctx ? ctx->launches : 0; }

fd_status fd_debug_check_guards(fd_context *ctx, int32_t *n_checked) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (n_checked) *n_checked = 0;
    if (!ctx->guard) return fail(ctx, FD_ERR_NOT_READY, "the context was not created with FD_B200_GUARD=1");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (const DevBuf *b : ctx->guarded) {
        if (!b->raw) continue;
        FD_TRY(check_red_zones(ctx, *b));
        if (n_checked) ++*n_checked;
    }
    return FD_OK;
}

fd_status fd_upload_frames(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames) {
    if (!ctx || !host_frames || rows <= 0 || cols <= 0 || n_frames <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_upload_frames: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pitch = (int64_t(cols) + 15) / 16 * 16;
    const int64_t stride = pitch * rows;
    FD_TRY(reserve(ctx, ctx->owned_frames, size_t(stride) * n_frames));
    if (pitch == cols) {
        FD_CUDA(ctx, cudaMemcpyAsync(ctx->owned_frames.ptr, host_frames, size_t(stride) * n_frames, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        FD_CUDA(ctx, cudaMemcpy2DAsync(ctx->owned_frames.ptr, size_t(pitch), host_frames, size_t(cols), size_t(cols), size_t(rows) * n_frames,
                                       cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->fv = FrameView{static_cast<const uint8_t *>(ctx->owned_frames.ptr), rows, cols, pitch, stride, n_frames, int(pitch / 4)};
    ctx->frames_bound = true;
    ctx->frame_map_valid = ctx->frame_map_failed = ctx->corner_map_valid = ctx->corner_map_failed = false;
    ctx->have_candidates = ctx->have_keypoints = ctx->have_desc = ctx->have_lsd = false;
    return FD_OK;
}

fd_status fd_bind_device_frames(fd_context *ctx, const uint8_t *dev_frames, int rows, int cols, int64_t pitch, int64_t frame_stride, int n_frames) {
    if (!ctx || !dev_frames || rows <= 0 || cols <= 0 || n_frames <= 0 || pitch < cols || frame_stride < pitch * rows)
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_bind_device_frames: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool aligned = (reinterpret_cast<uintptr_t>(dev_frames) % 4 == 0) && (pitch % 4 == 0) && (frame_stride % 4 == 0);
    if (aligned) {
        ctx->fv = FrameView{dev_frames, rows, cols, pitch, frame_stride, n_frames, int(pitch / 4)};
    } else {
        // re-pitch so that rows can be read as aligned words
        const int64_t np = (int64_t(cols) + 15) / 16 * 16;
        const int64_t ns = np * rows;
        FD_TRY(reserve(ctx, ctx->owned_frames, size_t(ns) * n_frames));
        for (int f = 0; f < n_frames; ++f)
            FD_CUDA(ctx, cudaMemcpy2DAsync(static_cast<uint8_t *>(ctx->owned_frames.ptr) + ns * f, size_t(np), dev_frames + frame_stride * f,
                                           size_t(pitch), size_t(cols), size_t(rows), cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->fv = FrameView{static_cast<const uint8_t *>(ctx->owned_frames.ptr), rows, cols, np, ns, n_frames, int(np / 4)};
    }
    ctx->frames_bound = true;
    ctx->frame_map_valid = ctx->frame_map_failed = ctx->corner_map_valid = ctx->corner_map_failed = false;
    ctx->have_candidates = ctx->have_keypoints = ctx->have_desc = ctx->have_lsd = false;
    return FD_OK;
}

fd_status fd_set_existing_features(fd_context *ctx, const float *host_xy, const int32_t *host_counts, int capacity, int n_frames) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (n_frames == 0) {
        ctx->have_existing = false;
        return FD_OK;
    }
    if (!host_xy || !host_counts || capacity <= 0 || n_frames < 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_set_existing_features: bad argument");
    for (int f = 0; f < n_frames; ++f)
        if (host_counts[f] < 0 || host_counts[f] > capacity) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "existing feature count exceeds capacity");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(reserve(ctx, ctx->existing_xy, size_t(n_frames) * capacity * 8));
    FD_TRY(reserve(ctx, ctx->existing_counts, size_t(n_frames) * 4));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->existing_xy.ptr, host_xy, size_t(n_frames) * capacity * 8, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->existing_counts.ptr, host_counts, size_t(n_frames) * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffers may be pageable and reused right away
    ctx->existing_capacity = capacity;
    ctx->existing_frames = n_frames;
    ctx->have_existing = true;
    return FD_OK;
}

fd_status fd_set_dense_outputs(fd_context *ctx, float *dev_response_map, uint8_t *dev_fast_score_map) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    ctx->resp_map = dev_response_map;
    ctx->score_map = dev_fast_score_map;
    return FD_OK;
}

fd_status fd_compute_candidates(fd_context *ctx, const fd_detect_params *params, int cand_capacity) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_candidates(ctx, params, cand_capacity);
}

fd_status fd_detect(fd_context *ctx, const fd_detect_params *params, int cand_capacity) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(run_candidates(ctx, params, cand_capacity));
    if (ctx->tiled) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_detect on a row tile: gather the tiles' candidates and call fd_select_candidates");
    FD_TRY(run_select(ctx, params, ctx->fv.rows, ctx->fv.cols, ctx->fv.n_frames, static_cast<uint64_t *>(ctx->keys.ptr),
                      static_cast<const uint32_t *>(ctx->counts.ptr), ctx->cand_capacity));
    ctx->kp_of_bound_frames = true;
    return FD_OK;
}

fd_status fd_set_tile(fd_context *ctx, int row_offset, int own_first_row, int own_row_count, int full_rows) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (full_rows <= 0) {
        ctx->tiled = false;
        return FD_OK;
    }
    if (row_offset < 0 || own_first_row < 0 || own_row_count < 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_set_tile: negative argument");
    ctx->tile = TileView{row_offset, own_first_row, own_first_row + own_row_count, full_rows};
    ctx->tiled = true;
    return FD_OK;
}

fd_status fd_device_candidates(fd_context *ctx, const uint64_t **dev_keys, const uint32_t **dev_counts, uint32_t *capacity) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_candidates) return fail(ctx, FD_ERR_NOT_READY, "no candidates computed");
    if (dev_keys) *dev_keys = static_cast<const uint64_t *>(ctx->keys.ptr);
    if (dev_counts) *dev_counts = static_cast<const uint32_t *>(ctx->counts.ptr);
    if (capacity) *capacity = ctx->cand_capacity;
    return FD_OK;
}

fd_status fd_export_candidates(fd_context *ctx, uint64_t *dev_dst, int64_t dst_capacity, int64_t *host_counts) {
    if (!ctx || !dev_dst || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_candidates) return fail(ctx, FD_ERR_NOT_READY, "no candidates computed");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(check_overflow(ctx));
    const int nf = ctx->fv.n_frames;
    std::vector<uint32_t> counts(static_cast<size_t>(nf));
    FD_CUDA(ctx, cudaMemcpyAsync(counts.data(), ctx->counts.ptr, size_t(nf) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t total = 0;
    for (int f = 0; f < nf; ++f) {
        host_counts[f] = counts[f];
        total += counts[f];
    }
    if (total > dst_capacity) return fail(ctx, FD_ERR_CAPACITY, "fd_export_candidates: destination too small");
    int64_t at = 0;
    for (int f = 0; f < nf; ++f) {
        if (counts[f] == 0) continue;
        FD_CUDA(ctx, cudaMemcpyAsync(dev_dst + at, static_cast<const uint64_t *>(ctx->keys.ptr) + int64_t(f) * ctx->cand_capacity, size_t(counts[f]) * 8,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
        at += counts[f];
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

fd_status fd_select_candidates(fd_context *ctx, const fd_detect_params *params, uint64_t *dev_keys, const uint32_t *dev_counts, uint32_t capacity,
                               int rows, int cols, int n_frames) {
    if (!ctx || !dev_keys || !dev_counts || rows <= 0 || cols <= 0 || n_frames <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_select_candidates: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(check_params(ctx, params));
    if (rows > 65535 || cols > 65535) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "frames are limited to 65535 x 65535");
    FD_TRY(reserve(ctx, ctx->flags, 16));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->flags.ptr, 0, 16, ctx->stream));   // the overflow flag of this selection
    if (ctx->have_existing) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "pre-existing features are not supported with external candidates");
    ctx->mask_view = MaskView{};
    ctx->select_frames = n_frames;
    return run_select(ctx, params, rows, cols, n_frames, dev_keys, dev_counts, capacity);
}

}  // extern "C"

fd_status fd_internal_select_first_range(fd_context *ctx, const fd_detect_params *params, const fd_select_prefilter *pf, int rows, int cols, int n_frames) {
    if (!ctx || !pf || rows <= 0 || cols <= 0 || n_frames <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_internal_select_first_range: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(check_params(ctx, params));
    FD_TRY(reserve(ctx, ctx->flags, 16));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->flags.ptr, 0, 16, ctx->stream));
    ctx->mask_view = MaskView{};
    SelectExtras ex;
    ex.first_range = pf;
    // the full key slots are not read in this mode: the candidate pointer only has to be non-null
    FD_TRY(run_select(ctx, params, rows, cols, n_frames, pf->pre_keys, pf->total_counts, pf->total_capacity, 0u, &ex));
    return FD_OK;
}

fd_status fd_internal_select_flagged(fd_context *ctx, const fd_detect_params *params, uint64_t *dev_keys, const uint32_t *dev_counts, uint32_t capacity,
                                     const uint8_t *flags, int rows, int cols, int n_frames) {
    if (!ctx || !dev_keys || !dev_counts || !flags) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_internal_select_flagged: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(check_params(ctx, params));
    ctx->mask_view = MaskView{};
    SelectExtras ex;
    ex.only_flagged = flags;
    return run_select(ctx, params, rows, cols, n_frames, dev_keys, dev_counts, capacity, 0u, &ex);
}

extern "C" {

fd_status fd_candidate_counts(fd_context *ctx, int32_t *host_counts) {
    if (!ctx || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_candidates) return fail(ctx, FD_ERR_NOT_READY, "no candidates computed");
    FD_CUDA(ctx, cudaMemcpyAsync(host_counts, ctx->counts.ptr, size_t(ctx->fv.n_frames) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

fd_status fd_download_candidates(fd_context *ctx, int frame, fd_candidate *host_cand, int64_t capacity, int64_t *n_out) {
    if (!ctx || !n_out) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_candidates) return fail(ctx, FD_ERR_NOT_READY, "no candidates computed");
    if (frame < 0 || frame >= ctx->fv.n_frames) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "frame out of range");
    FD_TRY(check_overflow(ctx));
    uint32_t n = 0;
    FD_CUDA(ctx, cudaMemcpyAsync(&n, static_cast<uint32_t *>(ctx->counts.ptr) + frame, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    if (host_cand == nullptr) return FD_OK;
    if (int64_t(n) > capacity) return fail(ctx, FD_ERR_CAPACITY, "host candidate buffer too small");
    std::vector<uint64_t> keys(n);
    FD_CUDA(ctx, cudaMemcpyAsync(keys.data(), static_cast<uint64_t *>(ctx->keys.ptr) + int64_t(frame) * ctx->cand_capacity, size_t(n) * 8,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!ctx->candidates_sorted) std::sort(keys.begin(), keys.end());  // presentation order only; selection never takes this path
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t xy = cand_key_xy(keys[i]);
        host_cand[i].response = cand_key_response(keys[i]);
        host_cand[i].x = int32_t(xy & 0xFFFFu);
        host_cand[i].y = int32_t(xy >> 16);
    }
    return FD_OK;
}

fd_status fd_download_keypoints(fd_context *ctx, fd_keypoint *host_kp, int32_t *host_counts, int kp_capacity) {
    if (!ctx || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_keypoints) return fail(ctx, FD_ERR_NOT_READY, "fd_detect has not run");
    const int nf = ctx->select_frames;
    uint32_t overflow = 0;   // read in the same round trip as the results (one synchronisation per call: it is most of a one-frame call's cost)
    if (ctx->flags.ptr != nullptr) FD_CUDA(ctx, cudaMemcpyAsync(&overflow, ctx->flags.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(host_counts, ctx->kp_counts.ptr, size_t(nf) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_kp != nullptr) {
        static_assert(sizeof(fd_keypoint) == sizeof(float4), "fd_keypoint layout");
        if (kp_capacity == ctx->kp_capacity) {
            FD_CUDA(ctx, cudaMemcpyAsync(host_kp, ctx->kp.ptr, size_t(nf) * kp_capacity * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            const int w = std::min(kp_capacity, ctx->kp_capacity);
            FD_CUDA(ctx, cudaMemcpy2DAsync(host_kp, size_t(kp_capacity) * sizeof(float4), ctx->kp.ptr, size_t(ctx->kp_capacity) * sizeof(float4),
                                           size_t(w) * sizeof(float4), size_t(nf), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (overflow != 0) return fail(ctx, FD_ERR_CAPACITY, "a frame produced more candidates than cand_capacity; raise it and run again");
    if (host_kp != nullptr)
        for (int f = 0; f < nf; ++f)
            if (host_counts[f] > kp_capacity) return fail(ctx, FD_ERR_CAPACITY, "host keypoint buffer too small");
    return FD_OK;
}

fd_status fd_device_keypoints(fd_context *ctx, const fd_keypoint **dev_kp, const int32_t **dev_counts, int *kp_capacity) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_keypoints) return fail(ctx, FD_ERR_NOT_READY, "fd_detect has not run");
    if (dev_kp) *dev_kp = static_cast<const fd_keypoint *>(ctx->kp.ptr);
    if (dev_counts) *dev_counts = static_cast<const int32_t *>(ctx->kp_counts.ptr);
    if (kp_capacity) *kp_capacity = ctx->kp_capacity;
    return FD_OK;
}

fd_status fd_sparsify(const float *host_xy, int n, int image_rows, int image_cols, int grid_rows, int grid_cols, uint8_t status_need_filter,
                      uint8_t status_after_filter, uint8_t *status) {
    if (n < 0 || (n > 0 && (!host_xy || !status)) || grid_rows < 2 || grid_cols < 2) return FD_ERR_INVALID_ARGUMENT;
    const int row_div = image_rows / (grid_rows - 1), col_div = image_cols / (grid_cols - 1);  // integer division, feature_point_detector.cpp:34-35
    const float row_step = float(row_div), col_step = float(col_div);
    std::vector<uint8_t> free_cell(size_t(grid_rows) * grid_cols, 1);                          // :36
    for (int i = 0; i < n; ++i) {
        // An image smaller than the grid makes a step 0 and the quotient infinite or NaN; the reference's float -> int cast of that is
        // undefined (INT_MIN on x86, i.e. outside the grid).  Same outcome here, without the undefined cast.
        const float qr = host_xy[2 * i + 1] / row_step, qc = host_xy[2 * i] / col_step;
        const bool in_range = std::fabs(qr) < 2.0e9f && std::fabs(qc) < 2.0e9f;                // false for NaN too
        const int row = in_range ? int(qr) : -1;                                               // :38
        const int col = in_range ? int(qc) : -1;                                               // :39
        if (row < 0 || row > grid_rows - 1 || col < 0 || col > grid_cols - 1) {                // :41-44
            status[i] = status_after_filter;
            continue;
        }
        uint8_t &cell = free_cell[size_t(row) * grid_cols + col];
        if (cell && status[i] == status_need_filter) cell = 0;                                 // :46-47
        else if (!cell && status[i] == status_need_filter) status[i] = status_after_filter;    // :48-49
    }
    return FD_OK;
}

// ---- BRIEF -----------------------------------------------------------------------------------------
static fd_status run_brief(fd_context *ctx, const fd_brief_params *p, const float4 *kp, const int32_t *counts, int capacity) {
    if (p->length < 1 || p->length > 256 || p->half_patch_size < 0 || p->half_patch_size > 64 || p->reserved != 0 ||
        (p->sampling != FD_SAMPLE_BILINEAR && p->sampling != FD_SAMPLE_TRUNCATE))
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "bad BRIEF parameters");
    const FrameView &fv = ctx->fv;
    FD_TRY(reserve(ctx, ctx->desc, size_t(fv.n_frames) * capacity * 32));
    BriefArgs a = {};
    a.fv = fv;
    a.keypoints = kp;
    a.kp_counts = counts;
    a.kp_capacity = capacity;
    a.length = p->length;
    a.half_patch = p->half_patch_size;
    a.sampling = p->sampling;
    a.integral_keypoints = ctx->desc_from_user ? 0 : 1;   // selected keypoints are pixel positions; caller-supplied ones may be fractional
    a.desc = static_cast<uint8_t *>(ctx->desc.ptr);
    FD_CUDA(ctx, launch_brief(a, ctx->stream));
    ++ctx->launches;
    ctx->desc_capacity = capacity;
    ctx->desc_length = p->length;
    ctx->desc_counts = counts;
    ctx->have_desc = true;
    ctx->have_desc_float = false;
    return FD_OK;
}

fd_status fd_describe_selected(fd_context *ctx, const fd_brief_params *params) {
    if (!ctx || !params) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(require_frames(ctx));
    if (!ctx->have_keypoints) return fail(ctx, FD_ERR_NOT_READY, "fd_detect has not run");
    // keypoints selected from gathered row-tile candidates or a heat map live in another frame count / coordinate frame than the bound frames
    if (!ctx->kp_of_bound_frames || ctx->select_frames != ctx->fv.n_frames)
        return fail(ctx, FD_ERR_NOT_READY, "the selected keypoints do not belong to the bound frames: bind the full frames and use fd_describe_points");
    ctx->desc_from_user = false;
    return run_brief(ctx, params, static_cast<const float4 *>(ctx->kp.ptr), static_cast<const int32_t *>(ctx->kp_counts.ptr), ctx->kp_capacity);
}

}  // extern "C"

namespace {
// Results of a small call laid out on the device exactly as the pinned staging block wants them, so that they leave in ONE
// copy instead of four (flag, counts, keypoints, descriptors): what a single-frame caller waits for is copy latency.
// 16-byte units: [flag | counts, padded | n_frames x w keypoints | n_frames x w descriptors].
__global__ void pack_results_kernel(const uint32_t *__restrict__ flag, const int32_t *__restrict__ counts, const uint4 *__restrict__ kp, int kp_stride,
                                    const uint4 *__restrict__ desc, int desc_stride, int w, int n_frames, int count_units, uint4 *__restrict__ out) {
    const int per_kp = n_frames * w;
    const int total = 1 + count_units + per_kp + (desc != nullptr ? 2 * per_kp : 0);
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < total; u += gridDim.x * blockDim.x) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (u == 0) {
            v.x = *flag;
        } else if (u < 1 + count_units) {
            const int i = (u - 1) * 4;
            v.x = i < n_frames ? uint32_t(counts[i]) : 0u;
            v.y = i + 1 < n_frames ? uint32_t(counts[i + 1]) : 0u;
            v.z = i + 2 < n_frames ? uint32_t(counts[i + 2]) : 0u;
            v.w = i + 3 < n_frames ? uint32_t(counts[i + 3]) : 0u;
        } else if (u < 1 + count_units + per_kp) {
            const int k = u - 1 - count_units;
            v = kp[size_t(k / w) * kp_stride + k % w];
        } else {
            const int k = u - 1 - count_units - per_kp;       // two units per descriptor
            const int slot = k >> 1;
            v = desc[(size_t(slot / w) * desc_stride + slot % w) * 2 + (k & 1)];
        }
        out[u] = v;
    }
}
constexpr size_t PACK_RESULTS_MAX = size_t(1) << 20;   // above this the strided copies are bandwidth-, not latency-bound
}  // namespace

extern "C" {

fd_status fd_detect_describe_host(fd_context *ctx, const uint8_t *host_frames, int rows, int cols, int n_frames, const fd_detect_params *det,
                                  const fd_brief_params *brief, int cand_capacity, fd_keypoint *host_kp, int32_t *host_counts, uint8_t *host_desc,
                                  int kp_capacity) {
    if (!ctx || !host_kp || !host_counts || kp_capacity <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_detect_describe_host: bad argument");
    if ((brief == nullptr) != (host_desc == nullptr)) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_detect_describe_host: brief parameters and host_desc go together");
    FD_TRY(fd_upload_frames(ctx, host_frames, rows, cols, n_frames));
    if (ctx->tiled) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_detect_describe_host on a row tile");
    FD_TRY(run_candidates(ctx, det, cand_capacity));
    FD_TRY(run_select(ctx, det, rows, cols, n_frames, static_cast<uint64_t *>(ctx->keys.ptr), static_cast<const uint32_t *>(ctx->counts.ptr), ctx->cand_capacity));
    ctx->kp_of_bound_frames = true;
    if (brief != nullptr) {
        ctx->desc_from_user = false;
        FD_TRY(run_brief(ctx, brief, static_cast<const float4 *>(ctx->kp.ptr), static_cast<const int32_t *>(ctx->kp_counts.ptr), ctx->kp_capacity));
    }
    // every result comes back through one pinned staging block and ONE synchronisation
    const int w = std::min(kp_capacity, ctx->kp_capacity);
    const int count_units = (n_frames + 3) / 4;
    const size_t kp_bytes = size_t(n_frames) * w * sizeof(float4), cnt_bytes = size_t(count_units) * 16, desc_bytes = brief ? size_t(n_frames) * w * 32 : 0;
    const size_t need = 16 + kp_bytes + cnt_bytes + desc_bytes;
    if (need > ctx->host_stage_bytes) {
        if (ctx->host_stage) FD_CUDA(ctx, cudaFreeHost(ctx->host_stage));
        ctx->host_stage = nullptr;
        ctx->host_stage_bytes = 0;
        FD_CUDA(ctx, cudaMallocHost(&ctx->host_stage, need));
        ctx->host_stage_bytes = need;
    }
    uint8_t *st = static_cast<uint8_t *>(ctx->host_stage);
    if (need <= PACK_RESULTS_MAX && w > 0) {
        FD_TRY(reserve(ctx, ctx->result_pack, need));
        const int units = int(need / 16);
        pack_results_kernel<<<std::min((units + 255) / 256, 592), 256, 0, ctx->stream>>>(
            static_cast<const uint32_t *>(ctx->flags.ptr), static_cast<const int32_t *>(ctx->kp_counts.ptr), static_cast<const uint4 *>(ctx->kp.ptr), ctx->kp_capacity,
            brief ? static_cast<const uint4 *>(ctx->desc.ptr) : nullptr, ctx->desc_capacity, w, n_frames, count_units, static_cast<uint4 *>(ctx->result_pack.ptr));
        FD_CUDA(ctx, cudaGetLastError());
        ++ctx->launches;
        FD_CUDA(ctx, cudaMemcpyAsync(st, ctx->result_pack.ptr, need, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        FD_CUDA(ctx, cudaMemcpyAsync(st, ctx->flags.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaMemcpyAsync(st + 16, ctx->kp_counts.ptr, size_t(n_frames) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (w > 0) {
            FD_CUDA(ctx, cudaMemcpy2DAsync(st + 16 + cnt_bytes, size_t(w) * sizeof(float4), ctx->kp.ptr, size_t(ctx->kp_capacity) * sizeof(float4), size_t(w) * sizeof(float4),
                                           size_t(n_frames), cudaMemcpyDeviceToHost, ctx->stream));
            if (brief != nullptr)
                FD_CUDA(ctx, cudaMemcpy2DAsync(st + 16 + cnt_bytes + kp_bytes, size_t(w) * 32, ctx->desc.ptr, size_t(ctx->desc_capacity) * 32, size_t(w) * 32, size_t(n_frames),
                                               cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t overflow;
    std::memcpy(&overflow, st, 4);
    if (overflow != 0) return fail(ctx, FD_ERR_CAPACITY, "a frame produced more candidates than cand_capacity; raise it and run again");
    std::memcpy(host_counts, st + 16, size_t(n_frames) * 4);
    for (int f = 0; f < n_frames; ++f) {
        if (host_counts[f] > kp_capacity) return fail(ctx, FD_ERR_CAPACITY, "host keypoint buffer too small");
        std::memcpy(host_kp + size_t(f) * kp_capacity, st + 16 + cnt_bytes + size_t(f) * w * sizeof(float4), size_t(w) * sizeof(float4));
        if (brief != nullptr) std::memcpy(host_desc + size_t(f) * kp_capacity * 32, st + 16 + cnt_bytes + kp_bytes + size_t(f) * w * 32, size_t(w) * 32);
    }
    return FD_OK;
}

fd_status fd_describe_points(fd_context *ctx, const fd_brief_params *params, const float *host_xy, const int32_t *host_counts, int capacity,
                             int n_frames) {
    if (!ctx || !params || !host_xy || !host_counts || capacity <= 0) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(require_frames(ctx));
    if (n_frames != ctx->fv.n_frames) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "n_frames does not match the bound frames");
    std::vector<float4> staged(size_t(n_frames) * capacity, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int f = 0; f < n_frames; ++f) {
        if (host_counts[f] < 0 || host_counts[f] > capacity) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "keypoint count exceeds capacity");
        for (int i = 0; i < host_counts[f]; ++i) {
            const size_t s = size_t(f) * capacity + i;
            staged[s] = make_float4(host_xy[2 * s], host_xy[2 * s + 1], 0.f, 0.f);
        }
    }
    FD_TRY(reserve(ctx, ctx->user_kp, staged.size() * sizeof(float4)));
    FD_TRY(reserve(ctx, ctx->user_counts, size_t(n_frames) * 4));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->user_kp.ptr, staged.data(), staged.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->user_counts.ptr, host_counts, size_t(n_frames) * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `staged` is pageable host memory going out of scope
    ctx->user_capacity = capacity;
    ctx->desc_from_user = true;
    return run_brief(ctx, params, static_cast<const float4 *>(ctx->user_kp.ptr), static_cast<const int32_t *>(ctx->user_counts.ptr), capacity);
}

fd_status fd_download_descriptors(fd_context *ctx, uint8_t *host_desc, int kp_capacity) {
    if (!ctx || !host_desc) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_desc) return fail(ctx, FD_ERR_NOT_READY, "no descriptors computed");
    const int nf = ctx->fv.n_frames;
    const int w = std::min(kp_capacity, ctx->desc_capacity);
    FD_CUDA(ctx, cudaMemcpy2DAsync(host_desc, size_t(kp_capacity) * 32, ctx->desc.ptr, size_t(ctx->desc_capacity) * 32, size_t(w) * 32, size_t(nf),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

fd_status fd_descriptors_as_float(fd_context *ctx, float *dev_out) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_desc) return fail(ctx, FD_ERR_NOT_READY, "no descriptors computed");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nf = ctx->fv.n_frames;
    const bool own = dev_out == nullptr;
    if (own) {
        FD_TRY(reserve(ctx, ctx->desc_float, size_t(nf) * ctx->desc_capacity * ctx->desc_length * 4));
        dev_out = static_cast<float *>(ctx->desc_float.ptr);
    }
    FD_CUDA(ctx, launch_brief_to_float(static_cast<const uint8_t *>(ctx->desc.ptr), ctx->desc_counts, ctx->desc_capacity, nf, ctx->desc_length, dev_out,
                                       ctx->stream));
    ++ctx->launches;
    if (own) ctx->have_desc_float = true;
    return FD_OK;
}

fd_status fd_download_descriptors_float(fd_context *ctx, float *host_desc, int kp_capacity) {
    if (!ctx || !host_desc) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_desc_float) return fail(ctx, FD_ERR_NOT_READY, "fd_descriptors_as_float has not written to the context's buffer");
    const size_t row = size_t(ctx->desc_length) * 4;
    const int w = std::min(kp_capacity, ctx->desc_capacity);
    FD_CUDA(ctx, cudaMemcpy2DAsync(host_desc, size_t(kp_capacity) * row, ctx->desc_float.ptr, size_t(ctx->desc_capacity) * row, size_t(w) * row,
                                   size_t(ctx->fv.n_frames), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

fd_status fd_device_descriptors(fd_context *ctx, const uint8_t **dev_desc, int *kp_capacity) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_desc) return fail(ctx, FD_ERR_NOT_READY, "no descriptors computed");
    if (dev_desc) *dev_desc = static_cast<const uint8_t *>(ctx->desc.ptr);
    if (kp_capacity) *kp_capacity = ctx->desc_capacity;
    return FD_OK;
}

// ---- Hamming matching (no reference counterpart) ------------------------------------------------------
fd_status fd_match_descriptors(fd_context *ctx, const uint8_t *dev_desc_a, const int32_t *dev_counts_a, int capacity_a, const uint8_t *dev_desc_b,
                               const int32_t *dev_counts_b, int capacity_b, int n_pairs, fd_match *dev_out) {
    if (!ctx || !dev_desc_a || !dev_counts_a || !dev_desc_b || !dev_counts_b || !dev_out || capacity_a <= 0 || capacity_b <= 0 || n_pairs < 0)
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_match_descriptors: bad argument");
    if (size_t(capacity_b) * 32 > 200 * 1024) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_match_descriptors: a train set is limited to 6400 descriptors");
    if ((reinterpret_cast<uintptr_t>(dev_desc_a) | reinterpret_cast<uintptr_t>(dev_desc_b) | reinterpret_cast<uintptr_t>(dev_out)) % 16 != 0)
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_match_descriptors: descriptor sets and the output must be 16-byte aligned");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(fd_match) == sizeof(int4), "fd_match layout");
    MatchArgs a = {};
    a.desc_a = dev_desc_a;
    a.desc_b = dev_desc_b;
    a.counts_a = dev_counts_a;
    a.counts_b = dev_counts_b;
    a.capacity_a = capacity_a;
    a.capacity_b = capacity_b;
    a.n_pairs = n_pairs;
    a.stride_a_sets = a.stride_b_sets = 1;
    a.out = reinterpret_cast<int4 *>(dev_out);
    FD_CUDA(ctx, launch_match(a, ctx->stream));
    ++ctx->launches;
    return FD_OK;
}

fd_status fd_match_consecutive(fd_context *ctx) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_desc) return fail(ctx, FD_ERR_NOT_READY, "no descriptors computed");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int pairs = std::max(ctx->fv.n_frames - 1, 0), cap = ctx->desc_capacity;
    if (size_t(cap) * 32 > 200 * 1024) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_match_consecutive: more than 6400 keypoint slots per frame");
    FD_TRY(reserve(ctx, ctx->matches, size_t(std::max(pairs, 1)) * cap * sizeof(int4)));
    MatchArgs a = {};
    a.desc_a = a.desc_b = static_cast<const uint8_t *>(ctx->desc.ptr);
    a.counts_a = a.counts_b = ctx->desc_counts;
    a.capacity_a = a.capacity_b = cap;
    a.n_pairs = pairs;
    a.stride_a_sets = a.stride_b_sets = 1;
    a.offset_b_sets = 1;
    a.out = static_cast<int4 *>(ctx->matches.ptr);
    FD_CUDA(ctx, launch_match(a, ctx->stream));
    if (pairs > 0) ++ctx->launches;
    ctx->match_pairs = pairs;
    ctx->match_capacity = cap;
    return FD_OK;
}

fd_status fd_download_matches(fd_context *ctx, fd_match *host_matches, int kp_capacity) {
    if (!ctx || !host_matches || kp_capacity <= 0) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->match_capacity == 0) return fail(ctx, FD_ERR_NOT_READY, "fd_match_consecutive has not run");
    if (ctx->match_pairs > 0) {
        const int w = std::min(kp_capacity, ctx->match_capacity);
        FD_CUDA(ctx, cudaMemcpy2DAsync(host_matches, size_t(kp_capacity) * sizeof(fd_match), ctx->matches.ptr, size_t(ctx->match_capacity) * sizeof(fd_match),
                                       size_t(w) * sizeof(fd_match), size_t(ctx->match_pairs), cudaMemcpyDeviceToHost, ctx->stream));
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

// ---- NN detector post-processing (nn_feature_point_detector.cpp:59-72, 128-155, 163-193) ------------------------------
fd_status fd_nn_select_from_heatmap(fd_context *ctx, const float *dev_heatmap, int rows, int cols, int n_frames, const fd_nn_params *params,
                                    int cand_capacity) {
    if (!ctx || !dev_heatmap || !params || rows <= 0 || cols <= 0 || n_frames <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_nn_select_from_heatmap: bad argument");
    if (params->reserved != 0 || params->invalid_boundary < 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_nn_params: invalid_boundary must be >= 0 and reserved 0");
    if (rows > 65535 || cols > 65535) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "maps are limited to 65535 x 65535");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t px = int64_t(rows) * cols;
    const uint32_t cap = cand_capacity > 0 ? uint32_t(std::min<int64_t>(cand_capacity, px)) : uint32_t(px);
    ctx->cand_capacity = cap;
    FD_TRY(reserve(ctx, ctx->keys, size_t(n_frames) * cap * 8));
    FD_TRY(reserve(ctx, ctx->counts, size_t(n_frames) * 4));
    FD_TRY(reserve(ctx, ctx->flags, 16));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->counts.ptr, 0, size_t(n_frames) * 4, ctx->stream));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->flags.ptr, 0, 16, ctx->stream));
    ctx->have_candidates = ctx->have_keypoints = ctx->have_nn_desc = false;
    ctx->candidates_sorted = false;

    // pre-existing features clear their squares from the mask (CreateMask -> UpdateMaskByFeatures, .cpp:68-70, 85-91)
    MaskView mask = {};
    if (ctx->have_existing) {
        if (ctx->existing_frames != n_frames) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "existing features were set for a different number of frames");
        const int wpr = (cols + 31) / 32 + 1;
        FD_TRY(reserve(ctx, ctx->mask_bits, size_t(n_frames) * rows * wpr * 4));
        MaskArgs m = {};
        m.rows = rows;
        m.cols = cols;
        m.n_frames = n_frames;
        m.min_distance = params->min_feature_distance;
        m.xy = static_cast<const float *>(ctx->existing_xy.ptr);
        m.counts = static_cast<const int32_t *>(ctx->existing_counts.ptr);
        m.capacity = ctx->existing_capacity;
        m.bits = static_cast<uint32_t *>(ctx->mask_bits.ptr);
        m.words_per_row = wpr;
        FD_CUDA(ctx, launch_mask(m, ctx->stream));
        ++ctx->launches;
        mask.bits = m.bits;
        mask.words_per_row = wpr;
    }
    ctx->mask_view = mask;

    NnHeatmapArgs a = {};
    a.heatmap = dev_heatmap;
    a.rows = rows;
    a.cols = cols;
    a.n_frames = n_frames;
    a.min_response = params->min_response;
    a.invalid_boundary = params->invalid_boundary;
    a.cand_keys = static_cast<uint64_t *>(ctx->keys.ptr);
    a.cand_counts = static_cast<uint32_t *>(ctx->counts.ptr);
    a.cand_capacity = cap;
    a.work_counter = static_cast<uint32_t *>(ctx->flags.ptr) + 3;   // flags[3]: zeroed above
    FD_CUDA(ctx, launch_nn_heatmap(a, ctx->sm_count, ctx->stream));
    ++ctx->launches;

    fd_detect_params sel = {};
    sel.kind = FD_HARRIS;   // selection only reads the distance and the count
    sel.min_feature_distance = params->min_feature_distance;
    sel.needed_feature_num = params->max_features;
    ctx->select_frames = n_frames;
    return run_select(ctx, &sel, rows, cols, n_frames, a.cand_keys, a.cand_counts, cap, 0xFFFFFFFFu);
}

fd_status fd_upload_floats(fd_context *ctx, int slot, const float *host, size_t count, const float **dev) {
    if (!ctx || !host || !dev || slot < 0 || slot > 1 || count == 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_upload_floats: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(reserve(ctx, ctx->float_slot[slot], count * 4));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->float_slot[slot].ptr, host, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    *dev = static_cast<const float *>(ctx->float_slot[slot].ptr);
    return FD_OK;
}

fd_status fd_nn_sample_descriptors(fd_context *ctx, const float *dev_maps, int channels, int map_rows, int map_cols, float *dev_out) {
    if (!ctx || !dev_maps || channels <= 0 || map_rows <= 0 || map_cols <= 0) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_nn_sample_descriptors: bad argument");
    if (!ctx->have_keypoints) return fail(ctx, FD_ERR_NOT_READY, "no keypoints selected");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nf = ctx->select_frames;
    if (!dev_out) {
        FD_TRY(reserve(ctx, ctx->nn_desc, size_t(nf) * ctx->kp_capacity * channels * 4));
        dev_out = static_cast<float *>(ctx->nn_desc.ptr);
    }
    NnDescriptorArgs a = {};
    a.maps = dev_maps;
    a.channels = channels;
    a.map_rows = map_rows;
    a.map_cols = map_cols;
    a.n_frames = nf;
    a.keypoints = static_cast<const float4 *>(ctx->kp.ptr);
    a.kp_counts = static_cast<const int32_t *>(ctx->kp_counts.ptr);
    a.kp_capacity = ctx->kp_capacity;
    a.out = dev_out;
    FD_CUDA(ctx, launch_nn_descriptors(a, ctx->stream));
    ++ctx->launches;
    ctx->nn_channels = channels;
    ctx->have_nn_desc = (dev_out == ctx->nn_desc.ptr);
    return FD_OK;
}

fd_status fd_nn_sample_descriptors_at(fd_context *ctx, const float *dev_maps, int channels, int map_rows, int map_cols, const float *host_xy,
                                      const int32_t *host_counts, int capacity, int n_frames, float *host_out) {
    if (!ctx || !dev_maps || !host_xy || !host_counts || !host_out || channels <= 0 || map_rows <= 0 || map_cols <= 0 || capacity <= 0 || n_frames <= 0)
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "fd_nn_sample_descriptors_at: bad argument");
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<float4> staged(size_t(n_frames) * capacity, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int f = 0; f < n_frames; ++f) {
        if (host_counts[f] < 0 || host_counts[f] > capacity) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "point count exceeds capacity");
        for (int i = 0; i < host_counts[f]; ++i) {
            const size_t s = size_t(f) * capacity + i;
            staged[s] = make_float4(host_xy[2 * s], host_xy[2 * s + 1], 0.f, 0.f);
        }
    }
    const size_t out_bytes = size_t(n_frames) * capacity * channels * 4;
    FD_TRY(reserve(ctx, ctx->user_kp, staged.size() * sizeof(float4)));
    FD_TRY(reserve(ctx, ctx->user_counts, size_t(n_frames) * 4));
    FD_TRY(reserve(ctx, ctx->nn_user_desc, out_bytes));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->user_kp.ptr, staged.data(), staged.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaMemcpyAsync(ctx->user_counts.ptr, host_counts, size_t(n_frames) * 4, cudaMemcpyHostToDevice, ctx->stream));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->nn_user_desc.ptr, 0, out_bytes, ctx->stream));   // unused slots read as zeros
    NnDescriptorArgs a = {};
    a.maps = dev_maps;
    a.channels = channels;
    a.map_rows = map_rows;
    a.map_cols = map_cols;
    a.n_frames = n_frames;
    a.keypoints = static_cast<const float4 *>(ctx->user_kp.ptr);
    a.kp_counts = static_cast<const int32_t *>(ctx->user_counts.ptr);
    a.kp_capacity = capacity;
    a.out = static_cast<float *>(ctx->nn_user_desc.ptr);
    FD_CUDA(ctx, launch_nn_descriptors(a, ctx->stream));
    ++ctx->launches;
    FD_CUDA(ctx, cudaMemcpyAsync(host_out, ctx->nn_user_desc.ptr, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `staged` is pageable host memory going out of scope
    return FD_OK;
}

fd_status fd_nn_download_descriptors(fd_context *ctx, float *host_desc, int kp_capacity) {
    if (!ctx || !host_desc) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_nn_desc) return fail(ctx, FD_ERR_NOT_READY, "fd_nn_sample_descriptors has not written to the context's buffer");
    const size_t row = size_t(ctx->nn_channels) * 4;
    const int w = std::min(kp_capacity, ctx->kp_capacity);
    FD_CUDA(ctx, cudaMemcpy2DAsync(host_desc, size_t(kp_capacity) * row, ctx->nn_desc.ptr, size_t(ctx->kp_capacity) * row, size_t(w) * row,
                                   size_t(ctx->select_frames), cudaMemcpyDeviceToHost, ctx->stream));
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FD_OK;
}

// ---- LSD field -------------------------------------------------------------------------------------
fd_status fd_lsd_field(fd_context *ctx, const fd_lsd_params *params, float *dev_norm, float *dev_angle, int32_t *dev_sorted_idx, int32_t *dev_n_valid) {
    if (!ctx || !params) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));
    FD_TRY(require_frames(ctx));
    const FrameView &fv = ctx->fv;
    if (fv.rows < 2 || fv.cols < 2) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "LSD needs rows >= 2 and cols >= 2 (feature_line_detector.cpp:14)");
    const size_t px = size_t(fv.rows) * fv.cols;
    if (!dev_norm) {
        FD_TRY(reserve(ctx, ctx->lsd_norm, px * fv.n_frames * 4));
        dev_norm = static_cast<float *>(ctx->lsd_norm.ptr);
    }
    if (!dev_angle) {
        FD_TRY(reserve(ctx, ctx->lsd_angle, px * fv.n_frames * 4));
        dev_angle = static_cast<float *>(ctx->lsd_angle.ptr);
    }
    if ((reinterpret_cast<uintptr_t>(dev_norm) | reinterpret_cast<uintptr_t>(dev_angle)) % 16 != 0)
        return fail(ctx, FD_ERR_INVALID_ARGUMENT, "norm / angle maps must be 16-byte aligned");
    LsdArgs a = {};
    a.fv = fv;
    a.min_norm = params->min_valid_gradient_norm;
    a.norm = dev_norm;
    a.angle = dev_angle;
    const int n_strips = (fv.cols + 127) / 128;
    int grid;
    plan_bands(ctx, fv.rows, n_strips, fv.n_frames, LSD_THREADS / 32, 4, 16, 2, a.band_rows, a.n_bands, a.n_items, grid);
    if (params->want_sorted) {
        FD_TRY(reserve(ctx, ctx->lsd_keys, size_t(a.n_items) * a.band_rows * 128 * 8));
        FD_TRY(reserve(ctx, ctx->lsd_item_counts, size_t(a.n_items) * 4));
        FD_TRY(reserve(ctx, ctx->lsd_chunk_sum, lsd_chunk_sum_bytes(fv.n_frames)));
        FD_TRY(reserve(ctx, ctx->lsd_counts, size_t(fv.n_frames) * 4));
        if (!dev_sorted_idx) {
            FD_TRY(reserve(ctx, ctx->lsd_sorted, px * fv.n_frames * 4));
            dev_sorted_idx = static_cast<int32_t *>(ctx->lsd_sorted.ptr);
        }
        FD_CUDA(ctx, cudaMemsetAsync(ctx->lsd_counts.ptr, 0, size_t(fv.n_frames) * 4, ctx->stream));
        const size_t hist_bytes = size_t(fv.n_frames) * LSD_BINS * 4;
        if (ctx->lsd_hist.bytes < hist_bytes || (ctx->guard && ctx->lsd_hist.bytes != hist_bytes)) ctx->lsd_hist_zeroed = 0;   // about to be reallocated
        FD_TRY(reserve(ctx, ctx->lsd_hist, hist_bytes));
        FD_TRY(reserve(ctx, ctx->lsd_start, hist_bytes));
        FD_TRY(reserve(ctx, ctx->lsd_bucketed, px * fv.n_frames * 8));
        if (ctx->lsd_hist_zeroed < hist_bytes) {
            FD_CUDA(ctx, cudaMemsetAsync(ctx->lsd_hist.ptr, 0, ctx->lsd_hist.bytes, ctx->stream));
            ctx->lsd_hist_zeroed = ctx->lsd_hist.bytes;
        }
        a.seed_keys = static_cast<uint64_t *>(ctx->lsd_keys.ptr);
        a.seed_counts = static_cast<uint32_t *>(ctx->lsd_counts.ptr);
        a.seed_hist = static_cast<uint32_t *>(ctx->lsd_hist.ptr);
        a.item_counts = static_cast<uint32_t *>(ctx->lsd_item_counts.ptr);
    }
    FD_TRY(reserve(ctx, ctx->lsd_work, 16));
    FD_CUDA(ctx, cudaMemsetAsync(ctx->lsd_work.ptr, 0, 16, ctx->stream));
    a.work_counter = static_cast<uint32_t *>(ctx->lsd_work.ptr);
    FD_CUDA(ctx, launch_lsd(a, grid, ctx->stream));
    ++ctx->launches;
    if (params->want_sorted) {
        FD_CUDA(ctx, launch_seed_order(a, static_cast<uint64_t *>(ctx->lsd_bucketed.ptr), static_cast<uint32_t *>(ctx->lsd_start.ptr),
                                       static_cast<uint32_t *>(ctx->lsd_chunk_sum.ptr), dev_sorted_idx, ctx->stream));
        ctx->launches += LSD_SEED_ORDER_LAUNCHES;
        if (dev_n_valid)
            FD_CUDA(ctx, cudaMemcpyAsync(dev_n_valid, ctx->lsd_counts.ptr, size_t(fv.n_frames) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->lsd_norm_p = dev_norm;
    ctx->lsd_angle_p = dev_angle;
    ctx->lsd_sorted_p = params->want_sorted ? dev_sorted_idx : nullptr;
    ctx->lsd_nvalid_p = params->want_sorted ? static_cast<int32_t *>(ctx->lsd_counts.ptr) : nullptr;
    ctx->have_lsd = true;
    ctx->lsd_sorted_valid = params->want_sorted != 0;
    return FD_OK;
}

fd_status fd_lsd_device_outputs(fd_context *ctx, const float **dev_norm, const float **dev_angle, const int32_t **dev_sorted_idx,
                                const int32_t **dev_n_valid) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    if (!ctx->have_lsd) return fail(ctx, FD_ERR_NOT_READY, "fd_lsd_field has not run");
    if (dev_norm) *dev_norm = ctx->lsd_norm_p;
    if (dev_angle) *dev_angle = ctx->lsd_angle_p;
    if (dev_sorted_idx) *dev_sorted_idx = ctx->lsd_sorted_p;
    if (dev_n_valid) *dev_n_valid = ctx->lsd_nvalid_p;
    return FD_OK;
}

fd_status fd_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return FD_ERR_INVALID_ARGUMENT;
    *ptr = nullptr;
    const cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        *ptr = nullptr;
        return e == cudaErrorMemoryAllocation ? FD_ERR_OUT_OF_MEMORY : FD_ERR_CUDA;
    }
    return FD_OK;
}

fd_status fd_host_free(void *ptr) {
    if (ptr == nullptr) return FD_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? FD_OK : FD_ERR_CUDA;
}

fd_status fd_lsd_download(fd_context *ctx, int frame, float *host_norm, float *host_angle, int32_t *host_sorted_idx, int64_t sorted_capacity,
                          int32_t *host_n_valid) {
    if (!ctx) return FD_ERR_INVALID_ARGUMENT;
    FD_CUDA(ctx, cudaSetDevice(ctx->device));   // a process may hold contexts on several devices
    if (!ctx->have_lsd) return fail(ctx, FD_ERR_NOT_READY, "fd_lsd_field has not run");
    const FrameView &fv = ctx->fv;
    if (frame < 0 || frame >= fv.n_frames) return fail(ctx, FD_ERR_INVALID_ARGUMENT, "frame out of range");
    const size_t px = size_t(fv.rows) * fv.cols;
    if (host_norm) FD_CUDA(ctx, cudaMemcpyAsync(host_norm, ctx->lsd_norm_p + px * frame, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (host_angle) FD_CUDA(ctx, cudaMemcpyAsync(host_angle, ctx->lsd_angle_p + px * frame, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    int32_t n = 0;
    if (ctx->lsd_sorted_valid) {
        FD_CUDA(ctx, cudaMemcpyAsync(&n, ctx->lsd_nvalid_p + frame, 4, cudaMemcpyDeviceToHost, ctx->stream));
        FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (host_sorted_idx) {
            if (int64_t(n) > sorted_capacity) return fail(ctx, FD_ERR_CAPACITY, "host sorted-index buffer too small");
            FD_CUDA(ctx, cudaMemcpyAsync(host_sorted_idx, ctx->lsd_sorted_p + px * frame, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    FD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (host_n_valid) *host_n_valid = n;
    return FD_OK;
}

fd_status fd_debug_fast_offset_bits(uint32_t count, uint32_t *out_bits, int32_t *n_segments) {
    if (!out_bits || count == 0) return FD_ERR_INVALID_ARGUMENT;
    const std::vector<OffsetSeg> segs = build_offset_table(count);
    if (n_segments) *n_segments = int32_t(segs.size()) - 1;
    size_t s = 0;
    for (uint32_t k = 0; k < count; ++k) {
        while (k >= segs[s + 1].k_start) ++s;
        out_bits[k] = segs[s].bits_start + (k - segs[s].k_start) * segs[s].step;
    }
    return FD_OK;
}

fd_status fd_debug_harris_trace_min(float min_valid_response, float *out_trace_min) {
    if (!out_trace_min) return FD_ERR_INVALID_ARGUMENT;
    *out_trace_min = harris_trace_min(min_valid_response);
    return FD_OK;
}

fd_status fd_debug_run_length_lut(uint8_t *out_65536) {
    if (!out_65536) return FD_ERR_INVALID_ARGUMENT;
    const std::vector<uint8_t> lut = build_run_lut();
    std::memcpy(out_65536, lut.data(), 65536);
    return FD_OK;
}

}  // extern "C"
