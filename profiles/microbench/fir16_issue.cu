// fir16_issue.cu -- how fast can one SM issue a half-band FIR body (taps and steering of hb_decimate.cuh) when nothing else is in
// the way (no shared memory, no barriers)?  Runs fir16_compute from registers in a loop and reports
// warp-instructions per cycle per SM for several warps-per-SM settings.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fir16_issue fir16_issue.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../sdrdaemon_b200/csrc/hb_decimate.cuh"
using namespace sdrd; using namespace sdrd::hb;

// the FIR body of K1 in its 16-output form (K1 itself now runs it as 32-output tasks in four groups of 8)
struct Fir16Regs {
    uint32_t w[48]; /* w[j] = O[n0 - 32 + j]: logical entries 16i .. 16i+47 = padded groups i .. i+2 */
    uint32_t e[17]; /* e[j] = E[n0 - 16 + j]: group i+1 and the first entry of group i+2 */
};

SDRD_DEVICE void fir16_compute(const Fir16Regs& f, uint32_t acc0, const Steer st, int (&y)[16])
{
    constexpr int H[16] = SDRD_HB64_TAPS;
    uint32_t acc[16];
#pragma unroll
    for (int r = 0; r < 16; r++) acc[r] = acc0;
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            /* O[n - t] = w[32 + r - t], O[n - 31 + t] = w[1 + r + t] */
            const uint32_t a = f.w[32 + r - t], b = f.w[1 + r + t];
            const uint32_t sum = t < SDRD_HB_FMA_ADD_TAPS ? mad_lo(a, st.one, b) : add3(a, b, st.zero);
            const uint32_t h = H[t] == 32 ? st.k32 : (H[t] == 256 ? st.k256 : (uint32_t)H[t]);
            acc[r] = mad_lo(sum, h, acc[r]);
        }
    }
    /* centre taps E[n0 - 15 + r] = e[1 + r] */
#pragma unroll
    for (int r = 0; r < 16; r++) y[r] = asr32(mad_lo(f.e[1 + r], st.k8192, acc[r]), HB_SHIFT);
}



__global__ void __launch_bounds__(128, 4) k(const uint32_t* in, uint32_t* out, int iters, Steer st, long long* cyc)
{
    Fir16Regs f;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    #pragma unroll
    for (int j = 0; j < 48; j++) f.w[j] = in[(tid + j * 131) & 0xFFFF];
    #pragma unroll
    for (int j = 0; j < 17; j++) f.e[j] = in[(tid + j * 17 + 7) & 0xFFFF];
    long long t0 = clock64();
    uint32_t sink = 0;
    #pragma unroll 3
    for (int it = 0; it < iters; it++) {
        int y[16];
        fir16_compute(f, 0u, st, y);
        // slide the window like consecutive tasks of a stream would: everything is loop-variant.
        // (unrolled by 3 the shifts are register renames)
        #pragma unroll
        for (int r = 0; r < 16; r++) {
            f.w[r] = f.w[r + 16];
            f.w[r + 16] = f.w[r + 32];
            f.w[r + 32] = (uint32_t)y[r];
            f.e[r] = f.e[r + 1] ^ (uint32_t)y[15 - r];
        }
        sink += (uint32_t)y[3];
    }
    long long t1 = clock64();
    out[tid] = sink;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv)
{
    uint32_t *in, *out; long long* cyc;
    cudaMalloc(&in, 65536 * 4); cudaMalloc(&out, 148 * 16 * 128 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    cudaMemset(in, 0x5A, 65536 * 4);
    Steer st = {0, 1, 32, 256, 8192};
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    for (int ctas_per_sm : {1, 2, 3, 4}) {
        int grid = 148 * ctas_per_sm;
        k<<<grid, 128>>>(in, out, 10, st, cyc);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<<<grid, 128>>>(in, out, iters, st, cyc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        long long h[148 * 16]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
        // instructions per iteration per warp: 16 outputs * (16 add + 16 mad + 1 centre + 1 shift) + ~20
        double inst = 562;
        printf("warps/SM=%2d  ms=%.3f  cycles/iter/warp=%.1f  IPC/SM by clock64=%.2f  clock64 rate=%.0f MHz  IPC/SM by wall time at 1965 MHz=%.2f  [%s]\n", ctas_per_sm * 4, ms,
               avg / iters, ctas_per_sm * 4 * inst / (avg / iters), avg / (ms * 1e3), ctas_per_sm * 4 * inst * iters / (ms * 1e-3) / 1965e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
