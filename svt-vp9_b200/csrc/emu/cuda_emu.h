/*
 * cuda_emu.h -- minimal single-block-at-a-time CPU emulation of the CUDA execution model, used
 * ONLY by the "-m 'not gpu'" unit tests (tests/emu/*.cpp: motion estimation, deblocking, mask building,
 * prediction) to exercise the kernel sources on a box without a GPU.  It is not a product path and not a fallback:
 * libsvtvp9_b200.so is built by nvcc from the same .cu sources and never includes this file.
 *
 * Model: a launch runs its blocks one after another; the threads of a block are real OS threads
 * synchronised by barriers, so __syncthreads(), shared-memory atomics and full-mask warp
 * shuffles behave as on the device.  __shared__ variables become function-local statics (one
 * block alive at a time).
 */
#ifndef SVTB200_CUDA_EMU_H
#define SVTB200_CUDA_EMU_H
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct uint3_emu { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
extern thread_local uint3_emu threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern thread_local uint8_t *emu_dyn_smem;
extern thread_local std::barrier<> *emu_block_barrier, *emu_warp_barrier, *emu_upper_half_barrier;
extern thread_local uint64_t *emu_warp_xchg; /* 32 slots per warp */

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

static inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
/* __syncwarp(mask): the whole warp, or -- for the aligned groups of 4 / 8 / 16 lanes the transform kernels synchronise -- a barrier of that group
 * alone (groups of one warp may run through different numbers of them, e.g. only the blocks that reconstruct) */
extern thread_local std::barrier<> **emu_group_barriers; /* [3][8]: groups of 4, 8, 16 lanes of this thread's warp */
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    const int n = __builtin_popcount(mask);
    if (n == 4 || n == 8 || n == 16) { emu_group_barriers[(n == 4 ? 0 : n == 8 ? 1 : 2) * 8 + __builtin_ctz(mask) / n]->arrive_and_wait(); return; }
    emu_warp_barrier->arrive_and_wait();
}
/* named barrier over the upper half of the CTA's threads (bar.sync 1, blockDim/2 on the device) */
static inline void emu_sync_upper_half() { emu_upper_half_barrier->arrive_and_wait(); }
template <class T> static inline T __ldg(const T *p) { return *p; }

static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) { int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255; r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i); }
    return r;
}
static inline uint32_t __vsadu4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) { int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255; r += (uint32_t)(x > y ? x - y : y - x); }
    return r;
}
static inline uint32_t __vavgu4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) { uint32_t x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255; r |= ((x + y + 1) >> 1) << (8 * i); }
    return r;
}
static inline uint32_t __vminu2(uint32_t a, uint32_t b) {
    const uint32_t l = (a & 0xffff) < (b & 0xffff) ? (a & 0xffff) : (b & 0xffff), h = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return l | h << 16;
}
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) {
    const uint32_t l = (a & 0xffff) > (b & 0xffff) ? (a & 0xffff) : (b & 0xffff), h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return l | h << 16;
}
static inline int __dp4a(uint32_t a, uint32_t b, int c) {
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 255) * (int)((b >> (8 * i)) & 255);
    return c;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMin(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicMax(T *p, T v) {
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

/* full-mask shuffles: every lane of the warp must participate */
template <class T> static inline T emu_shfl(T v, int src_lane) {
    const unsigned lane = threadIdx.x & 31;
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    emu_warp_xchg[lane] = bits;
    emu_warp_barrier->arrive_and_wait();
    uint64_t got = emu_warp_xchg[src_lane & 31];
    emu_warp_barrier->arrive_and_wait();
    T r; memcpy(&r, &got, sizeof(T)); return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { int l = (threadIdx.x & 31) + d; return emu_shfl(v, l > 31 ? (int)(threadIdx.x & 31) : l); }
template <class T> static inline T __shfl_sync(unsigned, T v, int l) { return emu_shfl(v, l); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { int l = (int)(threadIdx.x & 31) - d; return emu_shfl(v, l < 0 ? (int)(threadIdx.x & 31) : l); }
/* full-mask votes (every lane of the warp participates; lanes past the block's last thread do not exist) */
static inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned lane = threadIdx.x & 31, first = threadIdx.x & ~31u;
    const unsigned nthreads = blockDim.x * blockDim.y * blockDim.z, lanes = nthreads - first < 32 ? nthreads - first : 32;
    emu_warp_xchg[lane] = pred ? 1 : 0;
    emu_warp_barrier->arrive_and_wait();
    unsigned m = 0;
    for (unsigned i = 0; i < lanes; i++) m |= (unsigned)(emu_warp_xchg[i] & 1) << i;
    emu_warp_barrier->arrive_and_wait();
    return m;
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <class F> static void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    const unsigned nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32;
    std::vector<uint8_t> smem(smem_bytes + 64);
    std::vector<uint64_t> xchg((size_t)nwarps * 32);
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        std::barrier<> bar(nthreads);
        std::barrier<> upper(nthreads >= 2 ? nthreads - nthreads / 2 : 1);
        std::vector<std::unique_ptr<std::barrier<>>> wbars, gbars;
        std::vector<std::barrier<> *> gptr((size_t)nwarps * 24, nullptr);
        for (unsigned w = 0; w < nwarps; w++) {
            wbars.emplace_back(new std::barrier<>(std::min(32u, nthreads - w * 32)));
            for (unsigned si = 0; si < 3; si++)
                for (unsigned gi = 0; gi < 32u / (4u << si); gi++) { gbars.emplace_back(new std::barrier<>(4u << si)); gptr[(size_t)w * 24 + si * 8 + gi] = gbars.back().get(); }
        }
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < nthreads; t++)
            ts.emplace_back([&, t] {
                threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                blockIdx = {bx, by, bz}; blockDim = block; gridDim = grid;
                emu_dyn_smem = (uint8_t *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
                emu_block_barrier = &bar; emu_upper_half_barrier = &upper; emu_warp_barrier = wbars[t / 32].get(); emu_warp_xchg = xchg.data() + (size_t)(t / 32) * 32;
                emu_group_barriers = gptr.data() + (size_t)(t / 32) * 24;
                body();
            });
        for (auto &th : ts) th.join();
    }
}
#define SVTB200_EMU_DEFINE_STATE                                              \
    thread_local uint3_emu threadIdx, blockIdx; thread_local dim3 blockDim, gridDim; \
    thread_local uint8_t *emu_dyn_smem; thread_local std::barrier<> *emu_block_barrier, *emu_warp_barrier, *emu_upper_half_barrier; \
    thread_local uint64_t *emu_warp_xchg; thread_local std::barrier<> **emu_group_barriers;
#endif
