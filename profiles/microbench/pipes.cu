// Issue-rate micro-benchmark for the integer / logic / FMA pipes of one B200 SM.
// Evidence for DESIGN.md's ALU ceiling: the half-band cascade needs ~64 lane-ops per
// input sample, so the per-SM lane-op rate bounds the decimator before HBM does.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, unsigned long long* cyc)
{
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + threadIdx.x; }
    uint32_t h = seed | 5, g = seed | 9;
    float fa[CHAINS], fh = (float)seed;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) fa[i] = (float)a[i];
    unsigned long long fb[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) fb[i] = ((unsigned long long)a[i] << 32) | b[i];
    unsigned long long fh2 = ((unsigned long long)__float_as_uint(fh) << 32) | __float_as_uint(fh);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) {  // IMAD only
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(h), "r"(g));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(g), "r"(h));
            } else if (MODE == 1) {  // IADD3 only
                asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(h));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(g));
            } else if (MODE == 2) {  // 1:1 IADD + IMAD (the FIR inner pattern)
                uint32_t s;
                asm volatile("add.s32 %0, %1, %2;" : "=r"(s) : "r"(a[i]), "r"(b[i]));
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(s), "r"(h));
            } else if (MODE == 3) {  // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(a[i]) : "r"(h), "r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(b[i]) : "r"(g), "r"(a[i]));
            } else if (MODE == 4) {  // 1:1 LOP3 + IMAD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(a[i]) : "r"(h), "r"(g));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(g), "r"(h));
            } else if (MODE == 5) {  // FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(fa[i]) : "f"(fh));
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(fa[i]) : "f"(fh));
            } else if (MODE == 6) {  // FFMA2 (packed f32x2)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(fb[i]) : "l"(fh2));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(fb[i]) : "l"(fh2));
            } else if (MODE == 7) {  // dp4a
                asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(h));
                asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(a[i]), "r"(g));
            } else if (MODE == 8) {  // PRMT
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(h), "r"(g));
                asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(g), "r"(h));
            } else if (MODE == 9) {  // SHF (funnel shift)
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(h), "r"(g));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(g), "r"(h));
            } else if (MODE == 10) {  // 2 IMAD : 1 IADD
                uint32_t s;
                asm volatile("add.s32 %0, %1, %2;" : "=r"(s) : "r"(a[i]), "r"(b[i]));
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(s), "r"(h));
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(s), "r"(g));
            } else if (MODE == 11) {  // 1 FFMA2 : 1 IADD : 1 IMAD
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(fb[i]) : "l"(fh2));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(h));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(g), "r"(h));
            } else if (MODE == 12) {  // 1 LOP3 : 1 IADD
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(a[i]) : "r"(h), "r"(g));
                asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(g));
            } else if (MODE == 13) {  // IMAD with immediate multiplier
                asm volatile("mad.lo.s32 %0, %0, 5201, %1;" : "+r"(a[i]) : "r"(g));
                asm volatile("mad.lo.s32 %0, %0, -1698, %1;" : "+r"(b[i]) : "r"(h));
            } else if (MODE == 14) {  // 1:1 IADD + IMAD immediate coefficient
                uint32_t s;
                asm volatile("add.s32 %0, %1, %2;" : "=r"(s) : "r"(a[i]), "r"(b[i]));
                asm volatile("mad.lo.s32 %0, %1, 5201, %0;" : "+r"(a[i]) : "r"(s));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc += a[i] ^ b[i] ^ __float_as_uint(fa[i]) ^ (uint32_t)fb[i] ^ (uint32_t)(fb[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// shared-memory read bandwidth: LDS.128, conflict-free
__global__ void __launch_bounds__(1024) lds128(uint32_t* out, unsigned long long* cyc)
{
    __shared__ uint4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = buf[(threadIdx.x + i * 128 + it) & 2047];
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int opsPerInner, int threads)
{
    int nsm = 148, blocks = nsm;
    uint32_t* out; unsigned long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * blocks * 1024);
    cudaMalloc(&cyc, sizeof(unsigned long long) * blocks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += (double)hc[i]; avg /= blocks;
    double laneops = (double)ITERS * CHAINS * opsPerInner * threads;   // per SM
    printf("%-28s threads=%4d  lane-ops/clk/SM=%7.2f  cycles=%.0f  ms=%.3f  eff_clk_MHz=%.0f  chip_Tops=%.2f\n",
           name, threads, laneops / avg, avg, ms, avg / (ms * 1e3), laneops * blocks / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int threads : {256, 512, 1024}) {
        run<0>("IMAD", 2, threads);
        run<13>("IMAD imm", 2, threads);
        run<1>("IADD", 2, threads);
        run<2>("IADD+IMAD 1:1", 2, threads);
        run<14>("IADD+IMAD imm 1:1", 2, threads);
        run<10>("IADD+2 IMAD", 3, threads);
        run<3>("LOP3", 2, threads);
        run<4>("LOP3+IMAD 1:1", 2, threads);
        run<12>("LOP3+IADD 1:1", 2, threads);
        run<5>("FFMA", 2, threads);
        run<6>("FFMA2 (instrs)", 2, threads);
        run<11>("FFMA2+IADD+IMAD", 3, threads);
        run<7>("DP4A", 2, threads);
        run<8>("PRMT", 2, threads);
        run<9>("SHF", 2, threads);
    }
    {
        uint32_t* out; unsigned long long* cyc;
        cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8 * 148);
        for (int threads : {256, 1024}) {
            lds128<<<148, threads>>>(out, cyc); cudaDeviceSynchronize();
            unsigned long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; i++) avg += (double)hc[i]; avg /= 148;
            printf("LDS.128 threads=%d  bytes/clk/SM=%.1f\n", threads, (double)ITERS * 8 * 16 * threads / avg);
        }
    }
    return 0;
}
