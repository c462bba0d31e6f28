// Hamming matching of packed BRIEF descriptors -- the consumer-side step after kernel 4.  The reference has no
// matcher (its only consumer-facing descriptor transform is the +1 / -1 float form of descriptor.h:43-62), so there is no reference
// parity here: the checker is a numpy XOR / popcount restatement (tests/test_gpu_parity.py::test_hamming_matches_vs_numpy).
//
// One CTA per pair of descriptor sets (A: query, B: train).  B's descriptors are staged in shared memory once; a thread keeps one
// query descriptor in eight registers and walks B: eight XOR + POPC per pair, the B words read as two broadcast 128-bit loads.
// Per query: the nearest train descriptor (lowest index among equal distances), its distance and the second-smallest distance
// (for a ratio test).  Queries beyond counts_a and sets with an empty B yield index -1.
#include "fd_kernels.cuh"

namespace fdb {

namespace {

constexpr int MATCH_THREADS = 128;

__global__ void __launch_bounds__(MATCH_THREADS) match_kernel(const MatchArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint4 *train = reinterpret_cast<uint4 *>(smem);   // capacity_b x 2
    const int pair = blockIdx.x;
    const int64_t set_a = int64_t(pair) * p.stride_a_sets, set_b = int64_t(pair) * p.stride_b_sets + p.offset_b_sets;
    const int n_a = min(p.counts_a[set_a], p.capacity_a), n_b = min(p.counts_b[set_b], p.capacity_b);
    const uint4 *src_b = reinterpret_cast<const uint4 *>(p.desc_b + set_b * p.capacity_b * 32);
    for (int i = threadIdx.x; i < 2 * n_b; i += blockDim.x) train[i] = __ldg(src_b + i);
    __syncthreads();
    const uint4 *src_a = reinterpret_cast<const uint4 *>(p.desc_a + set_a * p.capacity_a * 32);
    int4 *out = p.out + int64_t(pair) * p.capacity_a;
    for (int q = threadIdx.x; q < p.capacity_a; q += blockDim.x) {
        int best = 0x7FFFFFFF, second = 0x7FFFFFFF, best_idx = -1;
        if (q < n_a) {
            const uint4 a0 = __ldg(src_a + 2 * q), a1 = __ldg(src_a + 2 * q + 1);
#pragma unroll 4
            for (int t = 0; t < n_b; ++t) {
                const uint4 b0 = train[2 * t], b1 = train[2 * t + 1];
                const int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                                 __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                if (dist < best) {
                    second = best;
                    best = dist;
                    best_idx = t;
                } else if (dist < second) {
                    second = dist;
                }
            }
        }
        out[q] = make_int4(best_idx, best_idx < 0 ? -1 : best, second == 0x7FFFFFFF ? -1 : second, 0);
    }
}

}  // namespace

cudaError_t launch_match(const MatchArgs &args, cudaStream_t stream) {
    if (args.n_pairs <= 0) return cudaSuccess;
    const size_t smem = size_t(args.capacity_b) * 32;
    cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    match_kernel<<<args.n_pairs, MATCH_THREADS, smem, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace fdb
