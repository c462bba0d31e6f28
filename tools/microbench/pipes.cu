// Issue-rate microbenchmark for the instruction mixes the FAST / corner kernels are built from (sm_100a).
// Prints warp-instructions per clock per SM sub-partition for each mix.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

#define ITERS 2048
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint32_t a[CHAINS], b = seed * 0x01010101u + threadIdx.x, c = seed ^ 0x5a5a5a5au;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 2654435761u + i * seed;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) a[i] = __vabsdiffu4(a[i], b);
            if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 2) asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(a[i]) : "r"(b));
            if (MODE == 3) asm volatile("add.sat.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 4) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a[i]) : "r"(b));
            if (MODE == 7) { if (i & 1) a[i] = __vabsdiffu4(a[i], b); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 8) { if (i & 1) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 9) { if (i & 1) a[i] = __vabsdiffu4(a[i], b); else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 10) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 11) asm volatile("vimnmx_placeholder_%=: max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 12) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 13) { asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a[i]) : "r"(b), "r"(c)); }
            // integer dot products and 16-bit pairs (candidates for the corner kernels' exact window sums), conversions
            if (MODE == 14) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 15) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 16) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 17) { if (i & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)); }
            if (MODE == 18) a[i] = __vadd2(a[i], b);
            if (MODE == 19) { if (i & 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(b)); else asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }
            if (MODE == 20) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(a[i]));
            if (MODE == 21) asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(a[i]));
            if (MODE == 22) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int per_iter, uint32_t *out, int sms, double mhz) {
    const int grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 3);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = double(grid) * 8 * ITERS * CHAINS * per_iter;
    const double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SMSP (at %.0f MHz nominal)\n", name, ms, warp_instr / clk / (sms * 4), mhz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    uint32_t *out; cudaMalloc(&out, size_t(p.multiProcessorCount) * 8 * 256 * 4);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("VABSDIFF4", 1, out, p.multiProcessorCount, mhz);
    run<1>("LOP3", 1, out, p.multiProcessorCount, mhz);
    run<2>("PRMT", 1, out, p.multiProcessorCount, mhz);
    run<3>("HADD2.SAT", 1, out, p.multiProcessorCount, mhz);
    run<4>("HFMA2", 1, out, p.multiProcessorCount, mhz);
    run<5>("IMAD", 1, out, p.multiProcessorCount, mhz);
    run<6>("SHF", 1, out, p.multiProcessorCount, mhz);
    run<7>("VABSDIFF4+LOP3 (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<8>("HFMA2+LOP3 (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<9>("VABSDIFF4+HFMA2 (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<10>("IADD", 1, out, p.multiProcessorCount, mhz);
    run<11>("VIMNMX.U16x2", 1, out, p.multiProcessorCount, mhz);
    run<12>("FFMA", 1, out, p.multiProcessorCount, mhz);
    run<13>("ISETP+SEL", 2, out, p.multiProcessorCount, mhz);
    run<14>("IDP.4A", 1, out, p.multiProcessorCount, mhz);
    run<15>("IDP.4A+LOP3 (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<16>("IDP.4A+HFMA2 (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<17>("IDP.4A+FFMA (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<18>("VIADD.16x2", 1, out, p.multiProcessorCount, mhz);
    run<19>("FMUL+FADD (1:1)", 1, out, p.multiProcessorCount, mhz);
    run<20>("I2F.U32", 1, out, p.multiProcessorCount, mhz);
    run<21>("F2I.U32", 1, out, p.multiProcessorCount, mhz);
    run<22>("IDP.2A", 1, out, p.multiProcessorCount, mhz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
