// Kernel 4 -- steered BRIEF-256 (intensity-centroid orientation + rotated ORB pattern).
// Replaces BriefDescriptor::ComputeForOneFeature / Descriptor<T>::Compute
// (reference src/feature_descriptor/descriptor_brief.cpp:8-50, descriptor.h:28-40).
//
// One warp per keypoint.  Every float expression is evaluated with explicit round-to-nearest
// intrinsics in the reference's association order (no FMA in the reference binary):
//   * orientation moments m10 = sum dx*I, m01 = sum dy*I over the (2h+1)^2 patch.  For keypoints on
//     integer coordinates (everything a detector emits) the terms are integers below 2^24,
//     so a warp-parallel sum is exact; for fractional coordinates two lanes replay the reference's
//     sequential dx-outer / dy-inner accumulation over the staged samples to stay bit-exact;
//   * bit i compares two samples at pattern points rotated by [c -s; s c]; each lane produces bits
//     lane, lane+32, ... and a ballot packs 32 of them into one output word (bit i -> byte i/8, bit i%8).
// The float-coordinate pixel fetch is the upstream Image::GetPixelValueNoCheck(float, float), which is not
// in the reference tree; FD_SAMPLE_BILINEAR reproduces compat/slam_utility/datatype_image.h (guess G1).
#include "fd_kernels.cuh"

namespace fdb {

namespace {

__device__ const int8_t kBriefPattern[256][4] = {
#include "brief_pattern_256.inc"
};

struct Img {
    const uint8_t *base;
    int rows, cols;
    int64_t pitch;
};

// Byte at flattened index semantics of a contiguous rows*cols buffer: column overflow wraps into the next
// row, anything past the last byte reads the last byte (the reference would read out of bounds there).
__device__ __forceinline__ float fetch(const Img &im, int r, int c) {
    if (c >= im.cols) {
        c -= im.cols;
        r += 1;
    }
    if (r >= im.rows) {
        r = im.rows - 1;
        c = im.cols - 1;
    }
    if (r < 0) {
        r = 0;
        c = 0;
    }
    return float(im.base[int64_t(r) * im.pitch + c]);
}

constexpr int BRIEF_WARPS = 8;
constexpr int BRIEF_MAX_PATCH = 25 * 25;  // half_patch <= 12 on the bit-exact sequential path (fractional keypoints)
// Shared-memory window of the frame around one keypoint: rows iv-19 .. iv+20, and enough aligned words per
// row to cover columns iu-19 .. iu+20.  Every tap of every rotated pattern point (radius <= 18.39, plus the
// +1 bilinear tap) and of an orientation patch with half size <= 18 falls inside it.
constexpr int WIN_ROWS = 40;
constexpr int WIN_WORDS = 12;
constexpr int WIN_PITCH = WIN_WORDS * 4;
constexpr int WIN_REACH = 19;

struct Window {
    const uint8_t *base;  // shared memory
    int r0, c0;           // frame coordinates of window element (0, 0)
};
__device__ __forceinline__ float wfetch(const Window &w, int r, int c) { return float(w.base[(r - w.r0) * WIN_PITCH + (c - w.c0)]); }

template <int SAMPLING, typename Src, typename Fetch>
__device__ __forceinline__ float sample_t(const Src &src, Fetch fetch_fn, float row, float col) {
    const int r0 = int(row), c0 = int(col);
    if (SAMPLING == 1) return fetch_fn(src, r0, c0);
    const float sub_row = __fsub_rn(row, floorf(row));
    const float sub_col = __fsub_rn(col, floorf(col));
    const float inv_sub_row = __fsub_rn(1.0f, sub_row);
    const float inv_sub_col = __fsub_rn(1.0f, sub_col);
    const float t0 = __fmul_rn(__fmul_rn(inv_sub_col, inv_sub_row), fetch_fn(src, r0, c0));
    const float t1 = __fmul_rn(__fmul_rn(sub_col, inv_sub_row), fetch_fn(src, r0, c0 + 1));
    const float t2 = __fmul_rn(__fmul_rn(inv_sub_col, sub_row), fetch_fn(src, r0 + 1, c0));
    const float t3 = __fmul_rn(__fmul_rn(sub_col, sub_row), fetch_fn(src, r0 + 1, c0 + 1));
    return __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3);
}

// The descriptor proper, generic over where pixels come from (shared-memory window or the frame itself).
template <int SAMPLING, typename Src, typename Fetch>
__device__ __forceinline__ void describe(const Src &src, Fetch fetch_fn, const BriefArgs &p, float u, float v, bool alive, float *staged,
                                         uint32_t *out) {
    const int lane = lane_id();
    float m10 = 0.0f, m01 = 0.0f;
    if (alive) {
        const int h = p.half_patch, side = 2 * h + 1, n = side * side;
        const bool integral = (u == floorf(u)) && (v == floorf(v));
        if (integral || n > BRIEF_MAX_PATCH) {
            // exact integer moments (B1): any summation order gives the reference's floats.
            // (Patches above BRIEF_MAX_PATCH with fractional coordinates are summed in this order too;
            //  no reference configuration reaches that case.)
            float s10 = 0.0f, s01 = 0.0f;
            for (int t = lane; t < n; t += 32) {
                const int dx = t / side - h, dy = t % side - h;  // dx outer, dy inner (brief.cpp:22-23)
                const float value = integral ? fetch_fn(src, int(v) + dy, int(u) + dx)
                                             : sample_t<SAMPLING>(src, fetch_fn, __fadd_rn(v, float(dy)), __fadd_rn(u, float(dx)));
                s10 = __fadd_rn(s10, __fmul_rn(float(dx), value));
                s01 = __fadd_rn(s01, __fmul_rn(float(dy), value));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s10 = __fadd_rn(s10, __shfl_xor_sync(0xffffffffu, s10, o));
                s01 = __fadd_rn(s01, __shfl_xor_sync(0xffffffffu, s01, o));
            }
            m10 = s10;
            m01 = s01;
        } else {
            for (int t = lane; t < n; t += 32) {
                const int dx = t / side - h, dy = t % side - h;
                staged[t] = sample_t<SAMPLING>(src, fetch_fn, __fadd_rn(v, float(dy)), __fadd_rn(u, float(dx)));
            }
            __syncwarp();
            if (lane < 2) {  // lane 0 replays m10, lane 1 replays m01, both in the reference's order
                float acc = 0.0f;
                for (int t = 0; t < n; ++t) {
                    const int dx = t / side - h, dy = t % side - h;
                    acc = __fadd_rn(acc, __fmul_rn(float(lane == 0 ? dx : dy), staged[t]));
                }
                m10 = acc;
            }
            m01 = __shfl_sync(0xffffffffu, m10, 1);
            m10 = __shfl_sync(0xffffffffu, m10, 0);
        }
    }
    const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(m01, m01), __fmul_rn(m10, m10)));  // brief.cpp:29
    if (m < 1e-6f) alive = false;                                                      // brief.cpp:30 (kZeroFloat, G2)
    const float sin_t = alive ? __fdiv_rn(m01, m) : 0.0f;                              // brief.cpp:32
    const float cos_t = alive ? __fdiv_rn(m10, m) : 1.0f;                              // brief.cpp:33
    const float neg_sin = -sin_t;

#pragma unroll 2
    for (int g = 0; g < 8; ++g) {
        const int i = g * 32 + lane;
        bool bit = false;
        if (alive && i < p.length) {                                                   // brief.cpp:38-47
            const char4 q = *reinterpret_cast<const char4 *>(kBriefPattern[i]);
            const float px1 = float(q.x), py1 = float(q.y), px2 = float(q.z), py2 = float(q.w);
            const float x1 = __fadd_rn(__fadd_rn(__fmul_rn(cos_t, px1), __fmul_rn(neg_sin, py1)), u);
            const float y1 = __fadd_rn(__fadd_rn(__fmul_rn(sin_t, px1), __fmul_rn(cos_t, py1)), v);
            const float x2 = __fadd_rn(__fadd_rn(__fmul_rn(cos_t, px2), __fmul_rn(neg_sin, py2)), u);
            const float y2 = __fadd_rn(__fadd_rn(__fmul_rn(sin_t, px2), __fmul_rn(cos_t, py2)), v);
            const float value_1 = sample_t<SAMPLING>(src, fetch_fn, y1, x1);
            const float value_2 = sample_t<SAMPLING>(src, fetch_fn, y2, x2);
            bit = value_1 < value_2;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) out[g] = word;
    }
}

template <int SAMPLING>
__global__ void __launch_bounds__(BRIEF_WARPS * 32) brief_kernel(const BriefArgs p) {
    __shared__ __align__(16) uint8_t window[BRIEF_WARPS][WIN_ROWS * WIN_PITCH];
    __shared__ float staged[BRIEF_WARPS][BRIEF_MAX_PATCH];
    const int lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const int64_t slot = int64_t(blockIdx.x) * BRIEF_WARPS + wib;
    const int frame = int(slot / p.kp_capacity);
    const int idx = int(slot % p.kp_capacity);
    if (frame >= p.fv.n_frames) return;
    if (idx >= p.kp_counts[frame]) return;

    const float4 kp = p.keypoints[slot];
    const float u = kp.x, v = kp.y;
    uint32_t *out = reinterpret_cast<uint32_t *>(p.desc + slot * 32);
    Img im = {p.fv.data + int64_t(frame) * p.fv.frame_stride, p.fv.rows, p.fv.cols, p.fv.pitch};

    // brief.cpp:13-17: border rejection leaves an all-zero descriptor
    const float max_bound = fmaxf(19.0f, __fmul_rn(float(p.half_patch), 2.0f));
    const bool alive = !(u < max_bound || u > __fsub_rn(float(im.cols), max_bound) || v < max_bound || v > __fsub_rn(float(im.rows), max_bound));

    // Stage the window in shared memory when it lies wholly inside the frame (always, except for keypoints on the
    // very last admissible rows / columns, whose taps touch the flattened-buffer wrap the reference relies on).
    const int iu = int(u), iv = int(v);
    const int wr0 = iv - WIN_REACH, wc0 = (iu - WIN_REACH) & ~3;
    const bool windowed = alive && p.half_patch <= 18 && wr0 >= 0 && iv + WIN_REACH + 1 < im.rows && iu - WIN_REACH >= 0 &&
                          iu + WIN_REACH + 1 < im.cols;
    if (windowed) {
        const int last_word = p.fv.words_per_row - 1;
        uint32_t *wdst = reinterpret_cast<uint32_t *>(window[wib]);
        for (int t = lane; t < WIN_ROWS * WIN_WORDS; t += 32) {
            const int r = t / WIN_WORDS, cw = t % WIN_WORDS;
            const int gw = min((wc0 >> 2) + cw, last_word);  // words past the row's end are never sampled
            wdst[t] = ld_word(im.base + int64_t(wr0 + r) * im.pitch + 4 * gw);
        }
        __syncwarp();
        const Window win = {window[wib], wr0, wc0};
        describe<SAMPLING>(win, [](const Window &w, int r, int c) { return wfetch(w, r, c); }, p, u, v, alive, staged[wib], out);
    } else {
        describe<SAMPLING>(im, [](const Img &i, int r, int c) { return fetch(i, r, c); }, p, u, v, alive, staged[wib], out);
    }
}

}  // namespace

cudaError_t launch_brief(const BriefArgs &args, cudaStream_t stream) {
    const int64_t slots = int64_t(args.fv.n_frames) * args.kp_capacity;
    if (slots == 0) return cudaSuccess;
    const int grid = int((slots + BRIEF_WARPS - 1) / BRIEF_WARPS);
    if (args.sampling == 1) brief_kernel<1><<<grid, BRIEF_WARPS * 32, 0, stream>>>(args);
    else brief_kernel<0><<<grid, BRIEF_WARPS * 32, 0, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
