/*
 * sdrd_platform.cuh -- the thin layer the kernels are written against.
 *
 * Product build (nvcc, sm_100a): everything below maps 1:1 onto CUDA built-ins and inline PTX
 * (mbarrier + cp.async.bulk, i.e. the TMA 1-D bulk copy; SASS: UBLKCP / SYNCS).
 *
 * Test build (-DSDRD_EMU, plain g++): the SAME kernel sources are compiled for the host and each CTA
 * is run as a group of std::threads with a std::barrier standing in for __syncthreads() and a
 * memcpy standing in for the bulk copy.  This exists only so that the indexing / pipelining logic
 * of the kernels can be checked against the oracle in the GPU-less build container
 * (tests/emu/).  It is TEST INFRASTRUCTURE: the C-ABI library never contains or calls it, and the
 * product path has no CPU fallback.
 */
#pragma once

#include <stdint.h>
#include <stddef.h>

#if defined(SDRD_EMU)
/* ------------------------------------------------------------------------------------------- */
/* host emulation of the handful of CUDA facilities the kernels use                             */
/* ------------------------------------------------------------------------------------------- */
#include <atomic>
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace sdrd_emu {
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
struct Cta {
    std::barrier<>* bar;
    unsigned char* smem;
    Dim3 blockIdx, blockDim, gridDim;
};
extern thread_local Cta* t_cta;
extern thread_local Dim3 t_tid;

template <class Body>
void launch(Dim3 grid, unsigned nthreads, size_t smem_bytes, Body body)
{
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            std::vector<unsigned char> smem(smem_bytes + 256, (unsigned char)0xA5); /* garbage on purpose */
            std::barrier<> bar((std::ptrdiff_t)nthreads);
            Cta cta;
            cta.bar = &bar;
            cta.smem = (unsigned char*)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
            cta.blockIdx = Dim3{bx, by, 1};
            cta.blockDim = Dim3{nthreads, 1, 1};
            cta.gridDim = grid;
            std::vector<std::thread> th;
            for (unsigned t = 0; t < nthreads; t++)
                th.emplace_back([&, t]() {
                    t_cta = &cta;
                    t_tid = Dim3{t, 0, 0};
                    body();
                });
            for (auto& x : th) x.join();
        }
}
} /* namespace sdrd_emu */

#define SDRD_DEVICE static inline
#define SDRD_HD inline
#define SDRD_KERNEL(bounds_threads, bounds_ctas) static void
#define SDRD_RESTRICT __restrict__
#define threadIdx (sdrd_emu::t_tid)
#define blockIdx (sdrd_emu::t_cta->blockIdx)
#define blockDim (sdrd_emu::t_cta->blockDim)
#define gridDim (sdrd_emu::t_cta->gridDim)
#define SDRD_DYN_SMEM(name) unsigned char* name = sdrd_emu::t_cta->smem
#define SDRD_GRID_CONSTANT const

static inline void __syncthreads() { sdrd_emu::t_cta->bar->arrive_and_wait(); }
/* kernels that use it run one warp per CTA, so the CTA barrier stands in for the warp barrier */
#define SDRD_SYNCWARP() sdrd_emu::t_cta->bar->arrive_and_wait()
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicXor(unsigned* a, unsigned v) { return __atomic_fetch_xor(a, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int* a, int v)
{
    int old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    /* PTX prmt, generic mode: nibble bit 3 replicates the sign of the selected byte */
    uint64_t pool = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned b = (unsigned)(pool >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}

namespace sdrd {
/* cp.async (LDGSTS): in the emulation a plain copy, complete at once */
static inline void cp_async4(void* dst_smem, const void* src_gmem) { memcpy(dst_smem, src_gmem, 4); }
static inline void cp_async_commit() {}
static inline void cp_async_wait_all() {}
typedef std::atomic<uint64_t> mbar_t; /* completed-phase counter */
static inline void mbar_init(mbar_t* b, int) { b->store(0, std::memory_order_relaxed); }
static inline void mbar_fence_init() {}
static inline void mbar_arrive_expect_tx(mbar_t*, uint32_t) {}
static inline void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, mbar_t* b)
{
    memcpy(dst_smem, src_gmem, bytes);
    b->fetch_add(1, std::memory_order_release);
}
static inline void mbar_wait(mbar_t* b, uint32_t parity)
{
    while (((uint32_t)b->load(std::memory_order_acquire) & 1u) == parity) std::this_thread::yield();
}
/* TMA 1-D bulk copy shared -> global (bulk async-group): in the emulation a plain copy, complete at once */
static inline void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) { memcpy(dst_gmem, src_smem, bytes); }
static inline void tma_store_commit() {}
static inline void tma_store_wait_read() {}
static inline void fence_proxy_async_smem() {}
/* TMA tensor store of one 32 x 32-word tile held in shared memory in the 128-byte swizzle (16-byte chunk c of row r sits
 * at chunk c ^ (r & 7)) to rows row0 .. row0 + 31 of a global array of 32-word rows: in the emulation a plain copy that
 * undoes the swizzle.  TileMap is the 128-byte descriptor (a CUtensorMap in the CUDA build, the base pointer here). */
struct alignas(64) TileMap { unsigned char opaque[128]; };
static inline void tma_store_tile32(const TileMap* map, const void* smem_tile, int row0)
{
    uint32_t* base;
    memcpy(&base, map->opaque, sizeof base);
    const unsigned char* t = (const unsigned char*)smem_tile;
    for (int r = 0; r < 32; r++)
        for (int c = 0; c < 8; c++) memcpy(base + ((size_t)(row0 + r) * 32 + 4 * c), t + (r * 8 + (c ^ (r & 7))) * 16, 16);
}
static inline uintptr_t smem_addr(const void* p) { return (uintptr_t)p; }
} /* namespace sdrd */

#else
/* ------------------------------------------------------------------------------------------- */
/* CUDA, sm_100a                                                                                */
/* ------------------------------------------------------------------------------------------- */
#include <cuda_runtime.h>

#define SDRD_DEVICE __device__ __forceinline__
#define SDRD_HD __host__ __device__ __forceinline__
#define SDRD_KERNEL(bounds_threads, bounds_ctas) __global__ void __launch_bounds__(bounds_threads, bounds_ctas)
#define SDRD_RESTRICT __restrict__
#define SDRD_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define SDRD_SYNCWARP() __syncwarp()
#define SDRD_GRID_CONSTANT const __grid_constant__ /* a kernel parameter whose address may be taken (the TMA descriptor) */

namespace sdrd {
typedef uint64_t mbar_t;

SDRD_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* cp.async, 4 bytes per thread: global -> shared without passing through registers (SASS: LDGSTS) */
SDRD_DEVICE void cp_async4(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
SDRD_DEVICE void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SDRD_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

SDRD_DEVICE void mbar_init(mbar_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
/* make the initialised barriers visible to the async (TMA) proxy */
SDRD_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

SDRD_DEVICE void mbar_arrive_expect_tx(mbar_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
/* TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP.S.G). */
SDRD_DEVICE void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, mbar_t* b)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(b))
        : "memory");
}
/* TMA 1-D bulk copy shared -> global of this thread's own bytes (SASS: UBLKCP.G.S): the store does not pass
 * through the LSU data pipe.  fence_proxy_async_smem() after the generic-proxy writes of the source, then the copy,
 * tma_store_commit(), and tma_store_wait_read() before the source is written again. */
SDRD_DEVICE void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
/* TMA tensor store (cp.async.bulk.tensor.2d, SASS UTMASTG): one 32 x 32-word tile, held in shared memory in the
 * 128-byte swizzle, to rows row0 .. row0 + 31 of the 2-D tensor the map describes (32 words per row).  Issued by ONE
 * thread; the tile must be 1024-byte aligned.  Completion as for the bulk copy: commit, wait_group.read before the tile is
 * overwritten. */
struct alignas(64) TileMap { unsigned char opaque[128]; };
SDRD_DEVICE void tma_store_tile32(const TileMap* map, const void* smem_tile, int row0)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem_tile)), "r"(0),
                 "r"(row0)
                 : "memory");
}
SDRD_DEVICE uint32_t smem_addr(const void* p) { return smem_u32(p); }
SDRD_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SDRD_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SDRD_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SDRD_DEVICE void mbar_wait(mbar_t* b, uint32_t parity)
{
    uint32_t a = smem_u32(b);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
} /* namespace sdrd */
#endif

namespace sdrd {
/* a + b + c as ONE three-source integer add.  In the CUDA build the two PTX adds are opaque to the
 * optimiser's reassociation and ptxas fuses them into a single IADD3 R, a, b, c (checked in SASS),
 * an instruction only the ALU pipe executes -- see Steer in hb_decimate.cuh. */
SDRD_DEVICE uint32_t add3(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(SDRD_EMU)
    return a + b + c;
#else
    uint32_t t, d;
    asm("add.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
    asm("add.u32 %0, %1, %2;" : "=r"(d) : "r"(t), "r"(c));
    return d;
#endif
}
/* a * b + c kept as a multiply-add (IMAD, FMA pipe) even when b is 1 at run time */
SDRD_DEVICE uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(SDRD_EMU)
    return a * b + c;
#else
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}
/* high word of the unsigned 64-bit product (IMAD.HI.U32, FMA pipe): a right shift that costs the ALU pipe nothing */
SDRD_DEVICE uint32_t mul_hi(uint32_t a, uint32_t b)
{
#if defined(SDRD_EMU)
    return (uint32_t)(((uint64_t)a * b) >> 32);
#else
    uint32_t d;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#endif
}
/* PTX prmt (generic mode) without the selector masking __byte_perm() adds: callers guarantee that
 * bit 3 of every selector nibble is clear.  SASS: one PRMT. */
SDRD_DEVICE uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
#if defined(SDRD_EMU)
    return __byte_perm(lo, hi, sel);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
#endif
}
/* arithmetic shift right of a wrapping 32-bit accumulator */
SDRD_DEVICE int asr32(uint32_t v, int sh) { return ((int)v) >> sh; }
} /* namespace sdrd */
