// Row-tiled detection of large frames across the GPUs of one box, behind the C ABI (include/fd_b200.h, fd_tiled_*).
// SURVEY.md 8e, BASELINE.json configs[3]: the dense stages of
//   feature_point_harris_detector.cpp:17-137 / feature_point_shi_tomas_detector.cpp:17-137 / feature_point_fast_detector.cpp:83-98
// shard by rows with a 3-row halo; the greedy selection (feature_point_detector.cpp:54-74) is global per frame and runs once,
// on the first tile's device, over the gathered candidate keys.
//
// One process, one fd_context per tile.  Nothing on this path waits for the host:
//   * every tile keeps rows [buf_lo, buf_hi) = its own rows + halo of each frame; the own rows arrive from the host or from one
//     device, the halo rows travel tile -> tile as peer copies (cudaMemcpy2DAsync between devices: NVLink when peer access is on),
//     ordered by events -- 2 x 3 x cols bytes per interior seam and frame;
//   * fd_compute_candidates runs per tile (fd_set_tile: candidates for the own rows, absolute row numbers, FAST's running offset
//     indexed by the absolute pixel position, so tiles are seam-free);
//   * the selection needs the best-ranked few thousand of a frame's 10^5 - 10^6 candidates, so only those travel: every tile builds the
//     rank histogram of its keys (select_hist_kernel), the root sums the histograms through peer pointers, derives each frame's first
//     rank limit exactly as the selection kernel would (tiled_limits_kernel) and hands the limits back (one small peer copy per tile);
//     the tiles compact the keys below them (select_admit_kernel), one gather kernel on the root packs those first ranges (plain loads
//     over NVLink, counts included -- no count ever visits the host), and the selection runs on them;
//   * a frame that needs more than its first range (or is too small to have one) is flagged on the device; a conditional full gather
//     and a second selection launch -- both enqueued unconditionally, both skipping the frames that are not theirs -- finish it, so
//     the result is exact whatever the data and still nothing waits for the host (FD_B200_TILED_PREFILTER=0: always gather every key);
//   * fd_tiled_compute_candidates (the candidate list itself is the result) gathers every key.
// The same code runs with all tiles on ONE device (device ordinals may repeat), which is how the single-GPU test-suite covers it.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/fd_b200.h"
#include "fd_internal.h"
#include "fd_select_common.cuh"

namespace {

constexpr int HALO = 3;           // gradient 1 + box window 1 + NMS 1 (Harris / Shi-Tomasi); the radius-3 ring (FAST)
constexpr int MAX_TILES = 16;

struct Tile {
    int device = 0;
    fd_context *ctx = nullptr;
    cudaStream_t stream = nullptr;
    int own_lo = 0, own_hi = 0, buf_lo = 0, buf_hi = 0;   // absolute rows
    uint8_t *buf = nullptr;
    size_t buf_bytes = 0;
    cudaEvent_t own_ready = nullptr, halo_done = nullptr, cand_done = nullptr;   // recorded on this tile's stream
    uint64_t *stage_keys = nullptr;     // root-side copies of this tile's key slots / counts when the root cannot read the peer directly
    uint32_t *stage_counts = nullptr;
    size_t stage_bytes = 0;
    bool root_reads_directly = true;
    // the candidates of the last fd_compute_candidates (device pointers, this tile's device or the root's staging copy)
    const uint64_t *cand_keys = nullptr;
    const uint32_t *cand_counts = nullptr;
    uint32_t cand_capacity = 0;
    // first-range prefilter (this tile's device): rank histogram, the limits the root derived, the compacted first range
    uint8_t *pre = nullptr;
    size_t pre_bytes = 0;
    uint32_t *hist = nullptr, *pre_counts = nullptr;
    uint64_t *limits = nullptr, *pre_keys = nullptr;
    uint32_t pre_capacity = 0;
    cudaEvent_t admit_done = nullptr;
    int own_count() const { return own_hi - own_lo; }
    int buf_rows() const { return buf_hi - buf_lo; }
};

struct GatherArgs {
    const uint64_t *keys[MAX_TILES];
    const uint32_t *counts[MAX_TILES];
    uint32_t capacity[MAX_TILES];
    int n_tiles, n_frames;
    uint64_t *dst;
    uint32_t *dst_counts;
    uint32_t dst_capacity;
    uint32_t *overflow_flag;
    const uint8_t *only_flagged;   // optional: frames whose flag is 0 are left alone
};

// Block (f, j): packs frame f's keys of every tile back to back; the blocks of a frame split each tile's keys between them.
__global__ void __launch_bounds__(256) gather_tiles_kernel(const GatherArgs a) {
    const int f = blockIdx.x;
    if (a.only_flagged != nullptr && a.only_flagged[f] == 0) return;
    uint32_t offset = 0u;
    bool overflow = false;
    uint64_t *dst = a.dst + int64_t(f) * a.dst_capacity;
    for (int t = 0; t < a.n_tiles; ++t) {
        const uint32_t count = a.counts[t][f];
        if (count > a.capacity[t]) overflow = true;
        const uint32_t n = min(count, a.capacity[t]);
        const uint64_t *src = a.keys[t] + int64_t(f) * a.capacity[t];
        for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x)
            if (offset + i < a.dst_capacity) dst[offset + i] = src[i];
        offset += n;
    }
    if (offset > a.dst_capacity) overflow = true;
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        a.dst_counts[f] = min(offset, a.dst_capacity);
        if (overflow) atomicExch(a.overflow_flag, 1u);
    }
}


// First-range prefilter, root side: the tiles' rank histograms summed per frame, the candidates counted, and the limit of the first
// rank range derived exactly as select_kernel derives it (same histogram, same helpers), so that the tiles can compact the keys
// below it and only those travel.  limit 0 = the frame is not prefiltered (few candidates, a first range that is the whole frame or
// larger than the gathered slots): it goes the full-gather way.
struct LimitsArgs {
    const uint32_t *hist[MAX_TILES];
    const uint32_t *counts[MAX_TILES];
    uint32_t capacity[MAX_TILES];
    int n_tiles;
    uint32_t needed, kp_capacity, pre_capacity;
    uint32_t *hist_sum;     // n_frames x 2048
    uint64_t *limits;       // n_frames
    uint32_t *totals;       // n_frames
    uint32_t *overflow_flag;
};

__global__ void __launch_bounds__(256) tiled_limits_kernel(const LimitsArgs a) {
    constexpr int BINS = 1 << fdb::SELECT_HIST_BITS;
    __shared__ uint32_t hist[BINS];
    __shared__ uint64_t s_limit;
    __shared__ uint32_t s_admit;
    const int f = blockIdx.x;
    for (int b = threadIdx.x; b < BINS; b += blockDim.x) {
        uint32_t sum = 0u;
        for (int t = 0; t < a.n_tiles; ++t) sum += a.hist[t][int64_t(f) * BINS + b];
        hist[b] = sum;
        a.hist_sum[int64_t(f) * BINS + b] = sum;
    }
    uint32_t total = 0u;
    bool overflow = false;
    for (int t = 0; t < a.n_tiles; ++t) {
        const uint32_t c = a.counts[t][f];
        overflow |= c > a.capacity[t];
        total += min(c, a.capacity[t]);
    }
    if (threadIdx.x == 0) {
        s_limit = 0ull;
        s_admit = 0u;
        a.totals[f] = total;
        if (overflow) atomicExch(a.overflow_flag, 1u);
    }
    __syncthreads();
    const uint32_t want = min(a.needed > 0u ? a.needed : 1u, a.kp_capacity);   // no pre-existing features on tiles
    const uint32_t prefix_k = fdb::select_first_range(want);
    if (total > uint32_t(fdb::SELECT_PREFIX_MIN) && prefix_k < total && threadIdx.x < 32) fdb::warp_prefix_limit(hist, prefix_k, total, &s_limit, &s_admit);
    __syncthreads();
    if (threadIdx.x == 0) a.limits[f] = (s_limit == fdb::kDeadKey || s_admit > a.pre_capacity) ? 0ull : s_limit;
}

}  // namespace

struct fd_tiled {
    std::vector<Tile> tiles;
    std::string err;
    int rows = 0, cols = 0, n_frames = 0;
    int64_t pitch = 0;
    bool have_frames = false, have_candidates = false, have_keypoints = false;
    uint64_t *gathered = nullptr;      // root device
    uint32_t *gathered_counts = nullptr, *flag = nullptr;
    size_t gathered_bytes = 0, counts_bytes = 0;
    uint32_t gathered_capacity = 0;
    uint64_t halo_bytes = 0;
    cudaEvent_t gather_done = nullptr;   // root stream: the root's kernels have read the tiles' key slots and first ranges
    cudaEvent_t limits_ready = nullptr;  // root stream: the limits have been pushed to the tiles
    bool full_gathered = false;          // t->gathered holds every frame's keys of the last candidates
    bool prefilter = true;               // FD_B200_TILED_PREFILTER=0: always gather every key (testing knob)
    uint8_t *root_pre = nullptr;         // root side of the prefilter: hist_sum | limits | totals | need_more | pre_counts | pre_keys
    size_t root_pre_bytes = 0;
    uint32_t *hist_sum = nullptr, *totals = nullptr, *root_pre_counts = nullptr;
    uint64_t *root_limits = nullptr, *root_pre_keys = nullptr;
    uint8_t *need_more = nullptr;
    uint32_t root_pre_capacity = 0;
    Tile &root() { return tiles[0]; }
};

namespace {

fd_status tfail(fd_tiled *t, fd_status st, const std::string &msg) {
    if (t) t->err = msg;
    return st;
}

#define TD_CUDA(t, call)                                                                                                     \
    do {                                                                                                                     \
        cudaError_t e__ = (call);                                                                                            \
        if (e__ != cudaSuccess)                                                                                              \
            return tfail(t, e__ == cudaErrorMemoryAllocation ? FD_ERR_OUT_OF_MEMORY : FD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define TD_FD(t, tile, call)                                                                   \
    do {                                                                                       \
        fd_status s__ = (call);                                                                \
        if (s__ != FD_OK) return tfail(t, s__, std::string(#call) + ": " + fd_last_error((tile).ctx)); \
    } while (0)

// Contiguous blocks of rows, sizes differing by at most one; trailing tiles are empty when there are more tiles than rows.
void plan(fd_tiled *t, int rows) {
    const int n = int(t->tiles.size());
    const int base = rows / n, rem = rows % n;
    int lo = 0;
    for (int k = 0; k < n; ++k) {
        Tile &tl = t->tiles[k];
        const int hi = lo + base + (k < rem ? 1 : 0);
        tl.own_lo = lo;
        tl.own_hi = hi;
        tl.buf_lo = hi > lo ? std::max(0, lo - HALO) : lo;
        tl.buf_hi = hi > lo ? std::min(rows, hi + HALO) : lo;
        lo = hi;
    }
}

fd_status layout(fd_tiled *t, int rows, int cols, int n_frames) {
    if (rows <= 0 || cols <= 0 || n_frames <= 0) return tfail(t, FD_ERR_INVALID_ARGUMENT, "fd_tiled: bad frame geometry");
    if (rows > 65535 || cols > 65535) return tfail(t, FD_ERR_INVALID_ARGUMENT, "frames are limited to 65535 x 65535");
    t->rows = rows;
    t->cols = cols;
    t->n_frames = n_frames;
    t->pitch = (int64_t(cols) + 15) / 16 * 16;
    plan(t, rows);
    bool grow = false;
    for (Tile &tl : t->tiles) grow |= tl.own_count() > 0 && size_t(t->pitch) * tl.buf_rows() * n_frames > tl.buf_bytes;
    if (grow) {   // a buffer about to be freed may still be the source of another tile's halo copy
        for (Tile &tl : t->tiles) {
            TD_CUDA(t, cudaSetDevice(tl.device));
            TD_CUDA(t, cudaStreamSynchronize(tl.stream));
        }
    }
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        const size_t need = size_t(t->pitch) * tl.buf_rows() * n_frames;
        if (need > tl.buf_bytes) {
            TD_CUDA(t, cudaSetDevice(tl.device));
            if (tl.buf) TD_CUDA(t, cudaFree(tl.buf));
            tl.buf = nullptr;
            tl.buf_bytes = 0;
            TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(&tl.buf), need));
            tl.buf_bytes = need;
        }
    }
    t->have_frames = t->have_candidates = t->have_keypoints = false;
    return FD_OK;
}

// Before a tile's own rows are rewritten, the tiles that copied halo rows out of them must be done reading.
fd_status wait_for_halo_readers(fd_tiled *t, int k) {
    Tile &me = t->tiles[k];
    TD_CUDA(t, cudaSetDevice(me.device));
    for (size_t o = 0; o < t->tiles.size(); ++o)
        if (int(o) != k && t->tiles[o].own_count() > 0) TD_CUDA(t, cudaStreamWaitEvent(me.stream, t->tiles[o].halo_done, 0));
    return FD_OK;
}

// The rows of other tiles' own blocks that fall inside this tile's buffer, copied device to device on this tile's stream.
fd_status exchange_halos(fd_tiled *t) {
    t->halo_bytes = 0;
    const int n = int(t->tiles.size());
    for (int k = 0; k < n; ++k) {
        Tile &me = t->tiles[k];
        if (me.own_count() == 0) continue;
        TD_CUDA(t, cudaSetDevice(me.device));
        const int64_t my_stride = t->pitch * me.buf_rows();
        for (int o = 0; o < n; ++o) {
            Tile &src = t->tiles[o];
            if (o == k || src.own_count() == 0) continue;
            const int lo = std::max(src.own_lo, me.buf_lo), hi = std::min(src.own_hi, me.buf_hi);
            if (hi <= lo) continue;
            TD_CUDA(t, cudaStreamWaitEvent(me.stream, src.own_ready, 0));
            const int64_t src_stride = t->pitch * src.buf_rows();
            // rows [lo, hi) of every frame: `hi - lo` contiguous pitched rows per frame, frames strided on both sides
            TD_CUDA(t, cudaMemcpy2DAsync(me.buf + int64_t(lo - me.buf_lo) * t->pitch, size_t(my_stride), src.buf + int64_t(lo - src.buf_lo) * t->pitch,
                                         size_t(src_stride), size_t(hi - lo) * t->pitch, size_t(t->n_frames), cudaMemcpyDefault, me.stream));
            t->halo_bytes += uint64_t(hi - lo) * t->cols * t->n_frames;
        }
        TD_CUDA(t, cudaEventRecord(me.halo_done, me.stream));
    }
    return FD_OK;
}

fd_status bind_tiles(fd_tiled *t) {
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        TD_FD(t, tl, fd_bind_device_frames(tl.ctx, tl.buf, tl.buf_rows(), t->cols, t->pitch, t->pitch * tl.buf_rows(), t->n_frames));
        TD_FD(t, tl, fd_set_tile(tl.ctx, tl.buf_lo, tl.own_lo - tl.buf_lo, tl.own_count(), t->rows));
    }
    t->have_frames = true;
    return FD_OK;
}

constexpr uint32_t PRE_CAPACITY = 65536;   // slots per frame for the gathered first rank ranges (a larger first range goes the full-gather way)

fd_status grow(fd_tiled *t, int device, cudaStream_t stream, uint8_t **buf, size_t *have, size_t need) {
    if (need <= *have) return FD_OK;
    TD_CUDA(t, cudaSetDevice(device));
    TD_CUDA(t, cudaStreamSynchronize(stream));
    if (*buf) TD_CUDA(t, cudaFree(*buf));
    *buf = nullptr;
    *have = 0;
    TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(buf), need));
    *have = need;
    return FD_OK;
}

// fd_compute_candidates on every tile (and, with the prefilter, each tile's rank histogram); cand_done marks the end of both.
fd_status tile_candidates(fd_tiled *t, const fd_detect_params *params, int cand_capacity_per_tile, bool with_hist) {
    if (!t->have_frames) return tfail(t, FD_ERR_NOT_READY, "fd_tiled: no frames distributed");
    if (!params) return tfail(t, FD_ERR_INVALID_ARGUMENT, "params is null");
    Tile &root = t->root();
    constexpr size_t BINS = size_t(1) << fdb::SELECT_HIST_BITS;
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        // the root's kernels of the previous call may still be reading this tile's key slots
        TD_CUDA(t, cudaSetDevice(tl.device));
        if (t->gather_done) TD_CUDA(t, cudaStreamWaitEvent(tl.stream, t->gather_done, 0));
        TD_FD(t, tl, fd_compute_candidates(tl.ctx, params, cand_capacity_per_tile));
        const uint64_t *keys = nullptr;
        const uint32_t *counts = nullptr;
        uint32_t cap = 0;
        TD_FD(t, tl, fd_device_candidates(tl.ctx, &keys, &counts, &cap));
        if (with_hist) {
            tl.pre_capacity = std::min(cap, PRE_CAPACITY);
            const size_t nf = size_t(t->n_frames);
            const size_t need = nf * BINS * 4 + nf * 8 + nf * 4 + nf * tl.pre_capacity * 8 + 64;
            fd_status st = grow(t, tl.device, tl.stream, &tl.pre, &tl.pre_bytes, need);
            if (st != FD_OK) return st;
            tl.pre_keys = reinterpret_cast<uint64_t *>(tl.pre);
            tl.limits = tl.pre_keys + nf * tl.pre_capacity;
            tl.hist = reinterpret_cast<uint32_t *>(tl.limits + nf);
            tl.pre_counts = tl.hist + nf * BINS;
            TD_CUDA(t, cudaSetDevice(tl.device));
            TD_CUDA(t, cudaMemsetAsync(tl.hist, 0, nf * BINS * 4 + nf * 4, tl.stream));   // histogram and first-range counts
            fdb::SelectArgs a = {};
            a.n_frames = t->n_frames;
            a.cand_keys = const_cast<uint64_t *>(keys);
            a.cand_counts = counts;
            a.cand_capacity = cap;
            a.pre_hist = tl.hist;
            a.hist_always = 1;
            TD_CUDA(t, fdb::launch_select_hist(a, tl.stream));
        }
        if (!tl.root_reads_directly) {
            // no peer access from the root: the whole key slots travel as copies (more bytes, same result)
            const size_t kb = size_t(t->n_frames) * cap * 8, cb = size_t(t->n_frames) * 4;
            TD_CUDA(t, cudaSetDevice(root.device));
            if (kb + cb > tl.stage_bytes) {
                TD_CUDA(t, cudaStreamSynchronize(root.stream));
                if (tl.stage_keys) TD_CUDA(t, cudaFree(tl.stage_keys));
                tl.stage_keys = nullptr;
                TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(&tl.stage_keys), kb + cb));
                tl.stage_bytes = kb + cb;
            }
            tl.stage_counts = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(tl.stage_keys) + kb);
            TD_CUDA(t, cudaSetDevice(tl.device));
            TD_CUDA(t, cudaMemcpyPeerAsync(tl.stage_keys, root.device, keys, tl.device, kb, tl.stream));
            TD_CUDA(t, cudaMemcpyPeerAsync(tl.stage_counts, root.device, counts, tl.device, cb, tl.stream));
            keys = tl.stage_keys;
            counts = tl.stage_counts;
        }
        tl.cand_keys = keys;
        tl.cand_counts = counts;
        tl.cand_capacity = cap;
        TD_CUDA(t, cudaSetDevice(tl.device));
        TD_CUDA(t, cudaEventRecord(tl.cand_done, tl.stream));
    }
    // the root-side slots of the full gather (filled by full_gather, for all frames or for the flagged ones)
    uint64_t total_capacity = 0;
    for (Tile &tl : t->tiles)
        if (tl.own_count() > 0) total_capacity += tl.cand_capacity;
    const uint32_t dst_cap = uint32_t(std::min<uint64_t>(total_capacity, uint64_t(t->rows) * t->cols));
    TD_CUDA(t, cudaSetDevice(root.device));
    const size_t need = size_t(t->n_frames) * dst_cap * 8, need_counts = size_t(t->n_frames) * 4;
    if (need > t->gathered_bytes || need_counts > t->counts_bytes) {
        TD_CUDA(t, cudaStreamSynchronize(root.stream));
        if (t->gathered) TD_CUDA(t, cudaFree(t->gathered));
        if (t->gathered_counts) TD_CUDA(t, cudaFree(t->gathered_counts));
        t->gathered = nullptr;
        t->gathered_counts = nullptr;
        TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(&t->gathered), std::max<size_t>(need, 16)));
        TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(&t->gathered_counts), need_counts));
        t->gathered_bytes = need;
        t->counts_bytes = need_counts;
    }
    if (!t->flag) TD_CUDA(t, cudaMalloc(reinterpret_cast<void **>(&t->flag), 16));
    t->gathered_capacity = dst_cap;
    for (Tile &tl : t->tiles)
        if (tl.own_count() > 0) TD_CUDA(t, cudaStreamWaitEvent(root.stream, tl.cand_done, 0));
    TD_CUDA(t, cudaMemsetAsync(t->flag, 0, 16, root.stream));
    t->have_candidates = true;
    t->have_keypoints = false;
    t->full_gathered = false;
    return FD_OK;
}

GatherArgs gather_args(fd_tiled *t, bool first_ranges) {
    GatherArgs g = {};
    g.n_frames = t->n_frames;
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        g.keys[g.n_tiles] = first_ranges ? tl.pre_keys : tl.cand_keys;
        g.counts[g.n_tiles] = first_ranges ? tl.pre_counts : tl.cand_counts;
        g.capacity[g.n_tiles] = first_ranges ? tl.pre_capacity : tl.cand_capacity;
        ++g.n_tiles;
    }
    g.overflow_flag = t->flag;
    return g;
}

fd_status launch_gather(fd_tiled *t, const GatherArgs &g) {
    Tile &root = t->root();
    TD_CUDA(t, cudaSetDevice(root.device));
    // enough blocks per frame to keep the copy near the link / HBM rate whatever the frame count
    const int per_frame = std::max(1, std::min(64, (148 * 8 + t->n_frames - 1) / t->n_frames));
    gather_tiles_kernel<<<dim3(unsigned(t->n_frames), unsigned(per_frame)), 256, 0, root.stream>>>(g);
    TD_CUDA(t, cudaGetLastError());
    return FD_OK;
}

fd_status mark_tiles_read(fd_tiled *t) {   // from here on the tiles may overwrite their key slots / first ranges
    Tile &root = t->root();
    TD_CUDA(t, cudaSetDevice(root.device));
    if (!t->gather_done) TD_CUDA(t, cudaEventCreateWithFlags(&t->gather_done, cudaEventDisableTiming));
    TD_CUDA(t, cudaEventRecord(t->gather_done, root.stream));
    return FD_OK;
}

// Every key of every tile (of the frames whose flag is set, if flags are given), packed per frame on the root.
fd_status full_gather(fd_tiled *t, const uint8_t *only_flagged) {
    GatherArgs g = gather_args(t, false);
    g.dst = t->gathered;
    g.dst_counts = t->gathered_counts;
    g.dst_capacity = t->gathered_capacity;
    g.only_flagged = only_flagged;
    if (g.n_tiles > 0) {
        fd_status st = launch_gather(t, g);
        if (st != FD_OK) return st;
    } else {
        TD_CUDA(t, cudaMemsetAsync(t->gathered_counts, 0, size_t(t->n_frames) * 4, t->root().stream));
    }
    if (only_flagged == nullptr) t->full_gathered = true;
    return FD_OK;
}

// Selection with the first-range prefilter: only the keys below each frame's first rank limit travel to the root; a frame that
// needs more than its first range (or is too small to have one) is flagged on the device and goes the full-gather way afterwards --
// both launches are enqueued unconditionally and skip the frames that are not theirs, so nothing waits for the host.
fd_status prefiltered_select(fd_tiled *t, const fd_detect_params *params) {
    Tile &root = t->root();
    constexpr size_t BINS = size_t(1) << fdb::SELECT_HIST_BITS;
    const size_t nf = size_t(t->n_frames);
    t->root_pre_capacity = uint32_t(std::min<uint64_t>(PRE_CAPACITY, t->gathered_capacity));
    const size_t need = nf * t->root_pre_capacity * 8 + nf * 8 + nf * BINS * 4 + nf * 4 + nf * 4 + nf + 64;
    fd_status st = grow(t, root.device, root.stream, &t->root_pre, &t->root_pre_bytes, need);
    if (st != FD_OK) return st;
    t->root_pre_keys = reinterpret_cast<uint64_t *>(t->root_pre);
    t->root_limits = t->root_pre_keys + nf * t->root_pre_capacity;
    t->hist_sum = reinterpret_cast<uint32_t *>(t->root_limits + nf);
    t->totals = t->hist_sum + nf * BINS;
    t->root_pre_counts = t->totals + nf;
    t->need_more = reinterpret_cast<uint8_t *>(t->root_pre_counts + nf);

    // limits from the summed histograms (the root stream already waits for every tile's candidates and histogram)
    TD_CUDA(t, cudaSetDevice(root.device));
    LimitsArgs la = {};
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        la.hist[la.n_tiles] = tl.hist;
        la.counts[la.n_tiles] = tl.cand_counts;
        la.capacity[la.n_tiles] = tl.cand_capacity;
        ++la.n_tiles;
    }
    la.needed = params->needed_feature_num;
    la.kp_capacity = std::max<uint32_t>(1u, std::min<uint32_t>(params->needed_feature_num, 1u << 20));   // as run_select sizes the keypoint slots
    la.pre_capacity = t->root_pre_capacity;
    la.hist_sum = t->hist_sum;
    la.limits = t->root_limits;
    la.totals = t->totals;
    la.overflow_flag = t->flag;
    tiled_limits_kernel<<<unsigned(t->n_frames), 256, 0, root.stream>>>(la);
    TD_CUDA(t, cudaGetLastError());
    for (Tile &tl : t->tiles)
        if (tl.own_count() > 0) TD_CUDA(t, cudaMemcpyPeerAsync(tl.limits, tl.device, t->root_limits, root.device, nf * 8, root.stream));
    if (!t->limits_ready) TD_CUDA(t, cudaEventCreateWithFlags(&t->limits_ready, cudaEventDisableTiming));
    TD_CUDA(t, cudaEventRecord(t->limits_ready, root.stream));

    // the tiles compact their keys below the limits
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        TD_CUDA(t, cudaSetDevice(tl.device));
        TD_CUDA(t, cudaStreamWaitEvent(tl.stream, t->limits_ready, 0));
        fdb::SelectArgs a = {};
        a.n_frames = t->n_frames;
        a.cand_keys = const_cast<uint64_t *>(tl.cand_keys);
        a.cand_counts = tl.cand_counts;
        a.cand_capacity = tl.cand_capacity;
        a.ext_limits = tl.limits;
        a.pre_keys = tl.pre_keys;
        a.pre_counts = tl.pre_counts;
        a.pre_capacity = tl.pre_capacity;
        TD_CUDA(t, fdb::launch_select_admit(a, tl.stream));
        if (!tl.admit_done) TD_CUDA(t, cudaEventCreateWithFlags(&tl.admit_done, cudaEventDisableTiming));
        TD_CUDA(t, cudaEventRecord(tl.admit_done, tl.stream));
    }

    // the root gathers the first ranges, selects on them, and deals with the frames that needed more
    TD_CUDA(t, cudaSetDevice(root.device));
    for (Tile &tl : t->tiles)
        if (tl.own_count() > 0) TD_CUDA(t, cudaStreamWaitEvent(root.stream, tl.admit_done, 0));
    TD_CUDA(t, cudaMemsetAsync(t->need_more, 0, nf, root.stream));
    GatherArgs g = gather_args(t, true);
    g.dst = t->root_pre_keys;
    g.dst_counts = t->root_pre_counts;
    g.dst_capacity = t->root_pre_capacity;
    st = launch_gather(t, g);
    if (st != FD_OK) return st;
    fd_select_prefilter pf = {};
    pf.total_counts = t->totals;
    pf.total_capacity = t->gathered_capacity;
    pf.hist = t->hist_sum;
    pf.pre_keys = t->root_pre_keys;
    pf.pre_counts = t->root_pre_counts;
    pf.pre_capacity = t->root_pre_capacity;
    pf.need_more = t->need_more;
    TD_FD(t, root, fd_internal_select_first_range(root.ctx, params, &pf, t->rows, t->cols, t->n_frames));
    st = full_gather(t, t->need_more);
    if (st != FD_OK) return st;
    TD_FD(t, root, fd_internal_select_flagged(root.ctx, params, t->gathered, t->totals, t->gathered_capacity, t->need_more, t->rows, t->cols, t->n_frames));
    return mark_tiles_read(t);
}

fd_status ensure_full_gather(fd_tiled *t) {   // the prefiltered detection gathers first ranges only: the callers that read every key pay for the rest
    if (t->full_gathered) return FD_OK;
    fd_status st = full_gather(t, nullptr);
    if (st != FD_OK) return st;
    return mark_tiles_read(t);
}

fd_status check_flag(fd_tiled *t) {
    Tile &root = t->root();
    uint32_t flag = 0;
    TD_CUDA(t, cudaSetDevice(root.device));
    TD_CUDA(t, cudaMemcpyAsync(&flag, t->flag, 4, cudaMemcpyDeviceToHost, root.stream));
    TD_CUDA(t, cudaStreamSynchronize(root.stream));
    if (flag != 0) return tfail(t, FD_ERR_CAPACITY, "a tile produced more candidates than cand_capacity_per_tile; raise it and run again");
    return FD_OK;
}

}  // namespace

extern "C" {

fd_status fd_tiled_create(const int *device_ordinals, int n_tiles, fd_tiled **out) {
    if (!out) return FD_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!device_ordinals || n_tiles <= 0 || n_tiles > MAX_TILES) return FD_ERR_INVALID_ARGUMENT;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return FD_ERR_NO_DEVICE;
    for (int k = 0; k < n_tiles; ++k)
        if (device_ordinals[k] < 0 || device_ordinals[k] >= n_dev) return FD_ERR_INVALID_ARGUMENT;
    fd_tiled *t = new (std::nothrow) fd_tiled();
    if (!t) return FD_ERR_OUT_OF_MEMORY;
    t->tiles.resize(size_t(n_tiles));
    if (const char *env = std::getenv("FD_B200_TILED_PREFILTER")) t->prefilter = (env[0] != '0');
    for (int k = 0; k < n_tiles; ++k) {
        Tile &tl = t->tiles[k];
        tl.device = device_ordinals[k];
        if (fd_create(tl.device, &tl.ctx) != FD_OK) {
            fd_tiled_destroy(t);
            return FD_ERR_CUDA;
        }
        tl.stream = static_cast<cudaStream_t>(fd_own_stream(tl.ctx));
        cudaSetDevice(tl.device);
        if (cudaEventCreateWithFlags(&tl.own_ready, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&tl.halo_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&tl.cand_done, cudaEventDisableTiming) != cudaSuccess) {
            fd_tiled_destroy(t);
            return FD_ERR_CUDA;
        }
        // events that were never recorded count as complete, so the first waits on them fall through
    }
    // peer access: tile <-> neighbouring tiles (halo copies) and root -> every tile (the gather kernel's loads)
    for (int k = 0; k < n_tiles; ++k) {
        for (int o = 0; o < n_tiles; ++o) {
            const int a = t->tiles[k].device, b = t->tiles[o].device;
            if (a == b || !(o == k + 1 || o == k - 1 || k == 0)) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (can) {
                cudaSetDevice(a);
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
            if (k == 0 && !can) t->tiles[o].root_reads_directly = false;
        }
    }
    *out = t;
    return FD_OK;
}

fd_status fd_tiled_destroy(fd_tiled *t) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    for (Tile &tl : t->tiles) {
        cudaSetDevice(tl.device);
        if (tl.stream) cudaStreamSynchronize(tl.stream);
    }
    if (!t->tiles.empty()) {
        cudaSetDevice(t->root().device);
        if (t->gathered) cudaFree(t->gathered);
        if (t->gathered_counts) cudaFree(t->gathered_counts);
        if (t->flag) cudaFree(t->flag);
        if (t->gather_done) cudaEventDestroy(t->gather_done);
        for (Tile &tl : t->tiles)
            if (tl.stage_keys) cudaFree(tl.stage_keys);
        if (t->root_pre) cudaFree(t->root_pre);
        if (t->limits_ready) cudaEventDestroy(t->limits_ready);
    }
    for (Tile &tl : t->tiles) {
        cudaSetDevice(tl.device);
        if (tl.buf) cudaFree(tl.buf);
        if (tl.pre) cudaFree(tl.pre);
        if (tl.admit_done) cudaEventDestroy(tl.admit_done);
        if (tl.own_ready) cudaEventDestroy(tl.own_ready);
        if (tl.halo_done) cudaEventDestroy(tl.halo_done);
        if (tl.cand_done) cudaEventDestroy(tl.cand_done);
        if (tl.ctx) fd_destroy(tl.ctx);
    }
    delete t;
    return FD_OK;
}

const char *fd_tiled_last_error(const fd_tiled *t) { return t ? t->err.c_str() : "null tiled detector"; }

fd_status fd_tiled_upload_frames(fd_tiled *t, const uint8_t *host_frames, int rows, int cols, int n_frames) {
    if (!t || !host_frames) return tfail(t, FD_ERR_INVALID_ARGUMENT, "fd_tiled_upload_frames: bad argument");
    fd_status st = layout(t, rows, cols, n_frames);
    if (st != FD_OK) return st;
    for (size_t k = 0; k < t->tiles.size(); ++k) {
        Tile &tl = t->tiles[k];
        if (tl.own_count() == 0) continue;
        st = wait_for_halo_readers(t, int(k));
        if (st != FD_OK) return st;
        const int64_t stride = t->pitch * tl.buf_rows();
        uint8_t *dst = tl.buf + int64_t(tl.own_lo - tl.buf_lo) * t->pitch;
        const uint8_t *src = host_frames + int64_t(tl.own_lo) * cols;
        if (t->pitch == cols) {   // the own rows of a frame are one contiguous run on both sides
            TD_CUDA(t, cudaMemcpy2DAsync(dst, size_t(stride), src, size_t(rows) * cols, size_t(tl.own_count()) * cols, size_t(n_frames), cudaMemcpyHostToDevice, tl.stream));
        } else {
            for (int f = 0; f < n_frames; ++f)
                TD_CUDA(t, cudaMemcpy2DAsync(dst + stride * f, size_t(t->pitch), src + int64_t(f) * rows * cols, size_t(cols), size_t(cols), size_t(tl.own_count()),
                                             cudaMemcpyHostToDevice, tl.stream));
        }
        TD_CUDA(t, cudaEventRecord(tl.own_ready, tl.stream));
    }
    st = exchange_halos(t);
    if (st != FD_OK) return st;
    // the caller's buffer may be pageable and reused right away
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        TD_CUDA(t, cudaSetDevice(tl.device));
        TD_CUDA(t, cudaStreamSynchronize(tl.stream));
    }
    return bind_tiles(t);
}

fd_status fd_tiled_scatter_device_frames(fd_tiled *t, const uint8_t *dev_frames, int rows, int cols, int64_t pitch, int64_t frame_stride, int n_frames) {
    if (!t || !dev_frames || pitch < cols || frame_stride < pitch * rows) return tfail(t, FD_ERR_INVALID_ARGUMENT, "fd_tiled_scatter_device_frames: bad argument");
    fd_status st = layout(t, rows, cols, n_frames);
    if (st != FD_OK) return st;
    for (size_t k = 0; k < t->tiles.size(); ++k) {
        Tile &tl = t->tiles[k];
        if (tl.own_count() == 0) continue;
        st = wait_for_halo_readers(t, int(k));
        if (st != FD_OK) return st;
        const int64_t stride = t->pitch * tl.buf_rows();
        uint8_t *dst = tl.buf + int64_t(tl.own_lo - tl.buf_lo) * t->pitch;
        for (int f = 0; f < n_frames; ++f)
            TD_CUDA(t, cudaMemcpy2DAsync(dst + stride * f, size_t(t->pitch), dev_frames + frame_stride * f + int64_t(tl.own_lo) * pitch, size_t(pitch), size_t(cols),
                                         size_t(tl.own_count()), cudaMemcpyDefault, tl.stream));
        TD_CUDA(t, cudaEventRecord(tl.own_ready, tl.stream));
    }
    st = exchange_halos(t);
    if (st != FD_OK) return st;
    return bind_tiles(t);
}

fd_status fd_tiled_tile_info(fd_tiled *t, int tile, int *device, int *own_first_row, int *own_row_count, uint8_t **dev_own_rows, int64_t *pitch,
                             int64_t *frame_stride, void **cuda_stream) {
    if (!t || tile < 0 || tile >= int(t->tiles.size())) return tfail(t, FD_ERR_INVALID_ARGUMENT, "fd_tiled_tile_info: bad argument");
    const Tile &tl = t->tiles[size_t(tile)];
    if (device) *device = tl.device;
    if (own_first_row) *own_first_row = tl.own_lo;
    if (own_row_count) *own_row_count = tl.own_count();
    if (dev_own_rows) *dev_own_rows = tl.buf ? tl.buf + int64_t(tl.own_lo - tl.buf_lo) * t->pitch : nullptr;
    if (pitch) *pitch = t->pitch;
    if (frame_stride) *frame_stride = t->pitch * tl.buf_rows();
    if (cuda_stream) *cuda_stream = tl.stream;
    return FD_OK;
}

fd_status fd_tiled_exchange_halos(fd_tiled *t) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    if (!t->have_frames) return tfail(t, FD_ERR_NOT_READY, "fd_tiled: no frames distributed");
    // the own rows are whatever the tiles' streams last wrote there (fd_tiled_tile_info hands out the pointers and streams)
    for (Tile &tl : t->tiles) {
        if (tl.own_count() == 0) continue;
        TD_CUDA(t, cudaSetDevice(tl.device));
        TD_CUDA(t, cudaEventRecord(tl.own_ready, tl.stream));
    }
    return exchange_halos(t);
}

uint64_t fd_tiled_halo_bytes(const fd_tiled *t) { return t ? t->halo_bytes : 0; }

fd_status fd_tiled_compute_candidates(fd_tiled *t, const fd_detect_params *params, int cand_capacity_per_tile) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    fd_status st = tile_candidates(t, params, cand_capacity_per_tile, false);
    if (st != FD_OK) return st;
    st = full_gather(t, nullptr);
    if (st != FD_OK) return st;
    return mark_tiles_read(t);
}

fd_status fd_tiled_detect(fd_tiled *t, const fd_detect_params *params, int cand_capacity_per_tile) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    bool prefilter = t->prefilter;
    for (const Tile &tl : t->tiles) prefilter = prefilter && (tl.own_count() == 0 || tl.root_reads_directly);
    fd_status st = tile_candidates(t, params, cand_capacity_per_tile, prefilter);
    if (st != FD_OK) return st;
    Tile &root = t->root();
    if (prefilter) {
        st = prefiltered_select(t, params);
        if (st != FD_OK) return st;
    } else {
        st = full_gather(t, nullptr);
        if (st != FD_OK) return st;
        st = mark_tiles_read(t);
        if (st != FD_OK) return st;
        TD_FD(t, root, fd_select_candidates(root.ctx, params, t->gathered, t->gathered_counts, t->gathered_capacity, t->rows, t->cols, t->n_frames));
    }
    t->have_keypoints = true;
    return FD_OK;
}

fd_status fd_tiled_sync(fd_tiled *t) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    for (Tile &tl : t->tiles) {
        TD_CUDA(t, cudaSetDevice(tl.device));
        TD_CUDA(t, cudaStreamSynchronize(tl.stream));
    }
    if (t->have_candidates) return check_flag(t);
    return FD_OK;
}

fd_status fd_tiled_candidate_counts(fd_tiled *t, int32_t *host_counts) {
    if (!t || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    if (!t->have_candidates) return tfail(t, FD_ERR_NOT_READY, "fd_tiled: no candidates computed");
    {
        const fd_status st_g = ensure_full_gather(t);
        if (st_g != FD_OK) return st_g;
    }
    fd_status st = check_flag(t);
    if (st != FD_OK) return st;
    Tile &root = t->root();
    TD_CUDA(t, cudaMemcpyAsync(host_counts, t->gathered_counts, size_t(t->n_frames) * 4, cudaMemcpyDeviceToHost, root.stream));
    TD_CUDA(t, cudaStreamSynchronize(root.stream));
    return FD_OK;
}

fd_status fd_tiled_device_candidates(fd_tiled *t, const uint64_t **dev_keys, const uint32_t **dev_counts, uint32_t *capacity, int *device) {
    if (!t) return FD_ERR_INVALID_ARGUMENT;
    if (!t->have_candidates) return tfail(t, FD_ERR_NOT_READY, "fd_tiled: no candidates computed");
    {
        const fd_status st_g = ensure_full_gather(t);
        if (st_g != FD_OK) return st_g;
    }
    if (dev_keys) *dev_keys = t->gathered;
    if (dev_counts) *dev_counts = t->gathered_counts;
    if (capacity) *capacity = t->gathered_capacity;
    if (device) *device = t->root().device;
    return FD_OK;
}

fd_status fd_tiled_download_candidates(fd_tiled *t, int frame, fd_candidate *host_cand, int64_t capacity, int64_t *n_out) {
    if (!t || !n_out) return FD_ERR_INVALID_ARGUMENT;
    if (!t->have_candidates) return tfail(t, FD_ERR_NOT_READY, "fd_tiled: no candidates computed");
    {
        const fd_status st_g = ensure_full_gather(t);
        if (st_g != FD_OK) return st_g;
    }
    if (frame < 0 || frame >= t->n_frames) return tfail(t, FD_ERR_INVALID_ARGUMENT, "frame out of range");
    fd_status st = check_flag(t);
    if (st != FD_OK) return st;
    Tile &root = t->root();
    uint32_t n = 0;
    TD_CUDA(t, cudaMemcpyAsync(&n, t->gathered_counts + frame, 4, cudaMemcpyDeviceToHost, root.stream));
    TD_CUDA(t, cudaStreamSynchronize(root.stream));
    *n_out = n;
    if (!host_cand) return FD_OK;
    if (int64_t(n) > capacity) return tfail(t, FD_ERR_CAPACITY, "host candidate buffer too small");
    std::vector<uint64_t> keys(n);
    TD_CUDA(t, cudaMemcpyAsync(keys.data(), t->gathered + int64_t(frame) * t->gathered_capacity, size_t(n) * 8, cudaMemcpyDeviceToHost, root.stream));
    TD_CUDA(t, cudaStreamSynchronize(root.stream));
    std::sort(keys.begin(), keys.end());   // presentation order: response descending, raster ties
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t xy = fdb::cand_key_xy(keys[i]);
        host_cand[i].response = fdb::cand_key_response(keys[i]);
        host_cand[i].x = int32_t(xy & 0xFFFFu);
        host_cand[i].y = int32_t(xy >> 16);
    }
    return FD_OK;
}

fd_status fd_tiled_download_keypoints(fd_tiled *t, fd_keypoint *host_kp, int32_t *host_counts, int kp_capacity) {
    if (!t || !host_counts) return FD_ERR_INVALID_ARGUMENT;
    if (!t->have_keypoints) return tfail(t, FD_ERR_NOT_READY, "fd_tiled_detect has not run");
    fd_status st = check_flag(t);
    if (st != FD_OK) return st;
    Tile &root = t->root();
    TD_FD(t, root, fd_download_keypoints(root.ctx, host_kp, host_counts, kp_capacity));
    return FD_OK;
}

fd_context *fd_tiled_root_context(fd_tiled *t) { return t ? t->root().ctx : nullptr; }

}  // extern "C"
