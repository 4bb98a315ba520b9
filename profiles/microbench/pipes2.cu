// pipes2.cu -- sustained (hundreds of ms) issue-rate check of the exact instruction pair the half-band
// FIR body is made of: a three-source IADD3 (ALU pipe) feeding an IMAD with an immediate multiplier
// (FMA pipe), 16 independent accumulators per thread.  Reports warp-instructions per clock per SM both
// against clock64() and against wall time, so that clock ramp-up / power capping is visible.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes2 pipes2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t t, d;
    asm("add.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
    asm("add.u32 %0, %1, %2;" : "=r"(d) : "r"(t), "r"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(128, 4) k(uint32_t* out, uint32_t seed, uint32_t zero, int iters, long long* cyc)
{
    uint32_t acc[16], w[32];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = seed + i * 3 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 32; i++) w[i] = seed * 7 + i + threadIdx.x * 5;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                if (MODE == 0) {  // IADD3 (3 sources) -> IMAD imm
                    uint32_t s = add3(w[(r + t) & 31], w[(r + 31 - t) & 31], zero);
                    asm("mad.lo.u32 %0, %1, 5201, %0;" : "+r"(acc[r]) : "r"(s));
                } else if (MODE == 1) {  // IADD3 only
                    acc[r] = add3(acc[r], w[(r + t) & 31], zero);
                    acc[r] = add3(acc[r], w[(r + 31 - t) & 31], zero);
                } else if (MODE == 2) {  // IMAD imm only
                    asm("mad.lo.u32 %0, %1, 5201, %0;" : "+r"(acc[r]) : "r"(w[(r + t) & 31]));
                    asm("mad.lo.u32 %0, %1, -1698, %0;" : "+r"(acc[r]) : "r"(w[(r + 31 - t) & 31]));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] ^= acc[i];
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int iters)
{
    int grid = 148 * ctas_per_sm;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, grid * 128 * 4); cudaMalloc(&cyc, grid * 8);
    k<MODE><<<grid, 128>>>(out, 12345u, 0u, 100, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 128>>>(out, 12345u, 0u, iters, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long* h = (long long*)malloc(grid * 8);
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
    double inst = 512.0 + 16 + 4;  // per iteration per warp
    double warps = ctas_per_sm * 4;
    printf("%-22s warps/SM=%2.0f  ms=%8.3f  IPC/SM by clock64=%.2f  clock64 rate=%.0f MHz  warp-inst/s/SM=%.3e  [%s]\n", name, warps,
           ms, warps * inst * iters / avg, avg / (ms * 1e3), warps * inst * iters / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc); free(h);
}

int main(int argc, char** argv)
{
    int iters = argc > 1 ? atoi(argv[1]) : 200000;
    for (int c : {1, 2, 4}) run<0>("IADD3+IMADimm 1:1", c, iters);
    for (int c : {2, 4}) run<1>("IADD3 3-src", c, iters);
    for (int c : {2, 4}) run<2>("IMAD imm", c, iters);
    return 0;
}
