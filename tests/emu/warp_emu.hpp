// warp_emu.hpp -- TEST INFRASTRUCTURE.  Just enough of the CUDA execution model to run the engine's kernels on the host:
// a thread block is blockDim.x host threads, the 32 lanes of a warp meet at a barrier for every shuffle / vote / __syncwarp,
// all threads of the block at __syncthreads; __shared__ variables are plain statics (one block runs at a time, blocks one
// after the other -- every kernel run here hands out its work by tickets and waits only for work with smaller tickets, so a
// block never waits for a later one).  Slow and simple on purpose: it is there to check a kernel's indexing, arithmetic and
// ordering protocol against the CPU oracle where no GPU is at hand, not to stand in for the hardware's memory model.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __grid_constant__
// dynamic shared memory of a kernel: a plain array the harness defines (kernels_emu.cpp)
#define B200_DYNAMIC_SMEM(name) extern uint8_t name[]

struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmuDim3 threadIdx, blockIdx;
inline EmuDim3 gridDim, blockDim;       // set by runGrid
using std::min;
using std::max;
using std::abs;

namespace warp_emu {
constexpr int kMaxWarps = 32;
struct Barrier {
    std::atomic<int> arrived{0};
    std::atomic<unsigned> generation{0};
    void wait(int parties) {
        const unsigned gen = generation.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) == parties - 1) {
            arrived.store(0, std::memory_order_relaxed);
            generation.fetch_add(1, std::memory_order_acq_rel);
        } else {
            while (generation.load(std::memory_order_acquire) == gen) std::this_thread::yield();
        }
    }
};
inline Barrier gWarpBarrier[kMaxWarps], gBlockBarrier;
inline unsigned long long gExchange[kMaxWarps][32];
inline int warpOf() { return (int)(threadIdx.x >> 5); }
inline int laneOf() { return (int)(threadIdx.x & 31); }
inline void warpSync() { gWarpBarrier[warpOf()].wait(32); }

// `body` as a grid of `blocks` blocks of `threads` threads (a multiple of 32), one block after the other
inline unsigned gBlockY = 0, gBlockZ = 0;   // blockIdx.y / .z of the blocks runGrid starts (set by runGrid3, stubs_rt/cuda_runtime.h)
inline void runGrid(unsigned blocks, unsigned threads, const std::function<void()> &body) {
    gridDim.x = blocks;
    blockDim.x = threads;
    for (unsigned b = 0; b < blocks; b++) {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; t++)
            pool.emplace_back([=, &body]() {
                threadIdx.x = t;
                blockIdx.x = b;
                blockIdx.y = gBlockY;
                blockIdx.z = gBlockZ;
                body();
            });
        for (auto &t : pool) t.join();
    }
}
template <typename T> inline unsigned long long pack(T v) {
    static_assert(sizeof(T) <= 8, "shuffles move at most 64 bits");
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    return raw;
}
template <typename T> inline T unpack(unsigned long long raw) {
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}
// every lane of the (full) warp deposits `v`, then reads lane pick(lane)'s value
template <typename T, typename F> inline T exchange(T v, F pick) {
    const int w = warpOf(), l = laneOf();
    gExchange[w][l] = pack(v);
    warpSync();
    const T r = unpack<T>(gExchange[w][pick(l)]);
    warpSync();
    return r;
}
}  // namespace warp_emu

// Warp-level primitives: every lane of the (full) warp must call them together -- true of the kernels run here.
template <typename T> inline T __shfl_sync(unsigned, T v, int srcLane) { return warp_emu::exchange(v, [=](int) { return srcLane & 31; }); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { return warp_emu::exchange(v, [=](int l) { return l >= (int)d ? l - (int)d : l; }); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { return warp_emu::exchange(v, [=](int l) { return l + (int)d < 32 ? l + (int)d : l; }); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return warp_emu::exchange(v, [=](int l) { return (l ^ m) & 31; }); }
inline unsigned __ballot_sync(unsigned, int pred) {
    const int w = warp_emu::warpOf(), l = warp_emu::laneOf();
    warp_emu::gExchange[w][l] = pred ? 1 : 0;
    warp_emu::warpSync();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(warp_emu::gExchange[w][i] & 1) << i;
    warp_emu::warpSync();
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline void __syncwarp() { warp_emu::warpSync(); }
inline void __syncthreads() { warp_emu::gBlockBarrier.wait((int)blockDim.x); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
inline uint4 __ldcg(const uint4 *p) { uint4 v; std::memcpy(&v, p, sizeof v); return v; }
inline uint2 __ldcg(const uint2 *p) { uint2 v; std::memcpy(&v, p, sizeof v); return v; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
// funnel shift right: the low 32 bits of (hi:lo) >> (shift & 31)
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31));
}
// __byte_perm: result byte i = byte (selector nibble i) of the eight bytes {y, x}; selectors 0..7 only
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((src >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
// per-byte unsigned average, rounded up
inline unsigned __vavgu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xFF) + ((b >> (8 * i)) & 0xFF) + 1) >> 1) << (8 * i);
    return r;
}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// ---- host stand-ins for device_ptx.cuh (its include guard is taken, so the kernels pick these up instead) ----
#define B200_DEVICE_PTX_CUH
#include "cuda.h"
namespace b200 {
inline uint32_t ldAcquire(const uint32_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void stRelease(uint32_t *p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
// (a twentieth of real time: 256 yielding host threads on a busy machine are slow, the kernels' 4 s watchdog becomes 80 s)
inline unsigned long long globalTimerNs() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() / 20;
}
// mbarrier / TMA.  The barrier word: bit 0 = parity of the current phase, bits 32.. = bytes still expected in it.  A tile load
// copies at once (elements outside the tensor read as zero, as with OOB_FILL_NONE) and then reports its bytes; the load that
// brings the count to zero completes the phase (release), which is what a parity wait observes (acquire).  What this cannot
// show is a missing wait on a transfer that the hardware would still have in flight -- here it has always landed.
inline uint32_t smemAddr(const void *) { return 0; }
inline void mbarInit(uint64_t *bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_RELEASE); }
inline void fenceMbarInit() {}
inline void fenceProxyAsync() {}
inline void mbarExpectTx(uint64_t *bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes << 32, __ATOMIC_ACQ_REL); }
inline bool mbarTryWait(uint64_t *bar, uint32_t parity) {
    const bool done = (__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1u) != (parity & 1u);
    if (!done) std::this_thread::yield();
    return done;
}
inline void mbarCompleteTx(uint64_t *bar, uint32_t bytes) {
    const uint64_t left = __atomic_sub_fetch(bar, (uint64_t)bytes << 32, __ATOMIC_ACQ_REL);
    if ((left >> 32) == 0) __atomic_fetch_xor(bar, 1ull, __ATOMIC_ACQ_REL);
}
inline void tmaTile(uint8_t *dst, const CUtensorMap *m, const int *c) {
    const uint32_t *b = m->box;
    const uint32_t n1 = m->rank > 1 ? b[1] : 1, n2 = m->rank > 2 ? b[2] : 1, n3 = m->rank > 3 ? b[3] : 1;
    for (uint32_t i3 = 0; i3 < n3; i3++)
        for (uint32_t i2 = 0; i2 < n2; i2++)
            for (uint32_t i1 = 0; i1 < n1; i1++)
                for (uint32_t i0 = 0; i0 < b[0]; i0++) {
                    const long long x[4] = {c[0] + (long long)i0, c[1] + (long long)i1, m->rank > 2 ? c[2] + (long long)i2 : 0,
                                            m->rank > 3 ? c[3] + (long long)i3 : 0};
                    bool in = true;
                    for (uint32_t d = 0; d < m->rank; d++) in = in && x[d] >= 0 && (uint64_t)x[d] < m->dims[d];
                    uint8_t v = 0;
                    if (in) {
                        uint64_t off = (uint64_t)x[0];
                        for (uint32_t d = 1; d < m->rank; d++) off += (uint64_t)x[d] * m->strides[d - 1];
                        v = m->base[off];
                    }
                    *dst++ = v;
                }
}
inline void tmaLoad3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    const int c[4] = {c0, c1, c2, 0};
    tmaTile(static_cast<uint8_t *>(dst), map, c);
    mbarCompleteTx(bar, map->box[0] * map->box[1] * map->box[2]);
}
inline void tmaLoad4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    const int c[4] = {c0, c1, c2, c3};
    tmaTile(static_cast<uint8_t *>(dst), map, c);
    mbarCompleteTx(bar, map->box[0] * map->box[1] * map->box[2] * map->box[3]);
}
// 1-D bulk copies (cp.async.bulk).  Both ends must be 16-byte aligned and the size a multiple of 16, or the hardware faults: checked
// here.  A load lands at once and reports its bytes to the mbarrier.  A store is DEFERRED: it joins the calling thread's open bulk
// group and reads its shared-memory source only when a wait lets the group go (bulkWaitRead<N>: all but the N most recent groups;
// bulkWaitAll: all) -- so a staging buffer that is refilled before the wait that protects it shows up as wrong pels, and stores
// still pending when the thread leaves the kernel are lost (its list dies with it).
struct BulkStoreOp { void *dst; const void *src; uint32_t bytes; };
inline thread_local std::vector<std::vector<BulkStoreOp>> tBulkGroups;
inline thread_local std::vector<BulkStoreOp> tBulkOpen;
inline void bulkCheck(const void *a, const void *b, uint32_t bytes) {
    if (((uintptr_t)a | (uintptr_t)b | bytes) & 15u) { std::fprintf(stderr, "warp_emu: misaligned cp.async.bulk (%p, %p, %u)\n", a, b, bytes); std::abort(); }
}
inline void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    bulkCheck(dst, src, bytes);
    std::memcpy(dst, src, bytes);
    mbarCompleteTx(bar, bytes);
}
inline void bulkStore(void *dst, const void *src, uint32_t bytes) {
    bulkCheck(dst, src, bytes);
    tBulkOpen.push_back(BulkStoreOp{dst, src, bytes});
}
inline void bulkCommit() { tBulkGroups.push_back(std::move(tBulkOpen)); tBulkOpen.clear(); }
inline void bulkDrain(size_t keep) {
    while (tBulkGroups.size() > keep) {
        for (const BulkStoreOp &op : tBulkGroups.front()) std::memcpy(op.dst, op.src, op.bytes);
        tBulkGroups.erase(tBulkGroups.begin());
    }
}
template <int N> inline void bulkWaitRead() { bulkDrain((size_t)N); }
inline void bulkWaitAll() { bulkDrain(0); }
// dp4a.u32.s32: acc + sum of (unsigned byte of pels) x (signed byte of taps)
inline int dp4aUS(uint32_t pels, int taps, int acc) {
    for (int i = 0; i < 4; i++) acc += (int)((pels >> (8 * i)) & 0xFF) * (int)(int8_t)(((uint32_t)taps >> (8 * i)) & 0xFF);
    return acc;
}
// dp2a.lo.s32.s32: acc + lo16(a) x byte0(b) + hi16(a) x byte1(b), all signed
inline int dp2aLoSS(uint32_t a, int b, int acc) {
    return acc + (int)(int16_t)(a & 0xFFFFu) * (int)(int8_t)((uint32_t)b & 0xFF) + (int)(int16_t)(a >> 16) * (int)(int8_t)(((uint32_t)b >> 8) & 0xFF);
}
// cvt.pack.sat.u8.s32 twice: four ints saturated to bytes, p0 in the low byte
inline uint32_t pack4sat(int p0, int p1, int p2, int p3) {
    auto sat = [](int v) { return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
    return sat(p0) | (sat(p1) << 8) | (sat(p2) << 16) | (sat(p3) << 24);
}
}  // namespace b200
