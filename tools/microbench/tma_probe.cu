// Probe: 3-D tiled TMA load of a u8 frame batch (box 144 x 16 x 1) with negative / out-of-range coordinates, as the sparse
// FAST kernel issues it.  Prints what landed in shared memory.  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, uint8_t *out, int BW) {
    __shared__ __align__(128) uint8_t tile[256 * 16];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BW * 16) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(s32(&bar))
                     : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bar))
        : "memory");
    for (int i = threadIdx.x; i < BW * 16; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
    const int BW = argc > 1 ? atoi(argv[1]) : 144;
    const int X = argc > 2 ? atoi(argv[2]) : 124, Y = argc > 3 ? atoi(argv[3]) : 20, Z = argc > 4 ? atoi(argv[4]) : 1;
    const int L2P = argc > 5 ? atoi(argv[5]) : 2;
    const int cols = 752, rows = 480, nf = 3;
    std::vector<uint8_t> h(size_t(cols) * rows * nf);
    for (size_t i = 0; i < h.size(); ++i) h[i] = uint8_t((i * 7 + i / cols) & 0xFF) | 1;
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, 256 * 16);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(ce), int(qr), fn);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {cols, rows, nf};
    const cuuint64_t strides[2] = {cols, cuuint64_t(cols) * rows};
    const cuuint32_t box[3] = {cuuint32_t(BW), 16, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                 CU_TENSOR_MAP_SWIZZLE_NONE, CUtensorMapL2promotion(L2P), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", int(r));
    if (r != CUDA_SUCCESS) return 1;
    const int cases[][3] = {{X, Y, Z}};
    int bad = 0;
    for (auto &c : cases) {
        probe<<<1, 32>>>(map, c[0], c[1], c[2], o, BW);
        ce = cudaDeviceSynchronize();
        printf("case x=%d y=%d z=%d: %s\n", c[0], c[1], c[2], cudaGetErrorString(ce));
        if (ce != cudaSuccess) return 1;
        std::vector<uint8_t> got(size_t(BW) * 16);
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int mism = 0;
        for (int rr = 0; rr < 16; ++rr)
            for (int cc = 0; cc < BW; ++cc) {
                const int gx = c[0] + cc, gy = c[1] + rr;
                const uint8_t want = (gx < 0 || gx >= cols || gy < 0 || gy >= rows) ? 0 : h[(size_t(c[2]) * rows + gy) * cols + gx];
                mism += (got[rr * BW + cc] != want);
            }
        printf("   mismatches: %d\n", mism);
        bad += mism;
    }
    return bad != 0;
}
