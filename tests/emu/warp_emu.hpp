// warp_emu.hpp -- TEST INFRASTRUCTURE.  Just enough of the CUDA execution model to run a kernel that uses nothing but
// thread / block indices, warp shuffles, warp barriers and plain memory accesses on the host: the 32 lanes of a warp are 32
// host threads that meet at a barrier for every __shfl_sync / __syncwarp.  One warp at a time (warps of such kernels do not
// talk to each other).  Slow and simple on purpose: it is there to check a kernel's indexing and arithmetic against the CPU
// oracle where no GPU is at hand, not to stand in for the hardware's memory model.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local EmuDim3 threadIdx, blockIdx;
inline EmuDim3 gridDim;       // set by the caller before runBlock
using std::min;
using std::max;

namespace warp_emu {
struct Barrier {
    std::atomic<int> arrived{0};
    std::atomic<unsigned> generation{0};
    void wait() {
        const unsigned gen = generation.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) == 31) {
            arrived.store(0, std::memory_order_relaxed);
            generation.fetch_add(1, std::memory_order_acq_rel);
        } else {
            while (generation.load(std::memory_order_acquire) == gen) std::this_thread::yield();
        }
    }
};
inline Barrier gBarrier;
inline unsigned long long gExchange[32];

// run `body` as block `block` of `warpsPerBlock` warps, one warp after the other
inline void runBlock(unsigned block, unsigned warpsPerBlock, const std::function<void()> &body) {
    for (unsigned w = 0; w < warpsPerBlock; w++) {
        std::vector<std::thread> lanes;
        for (unsigned l = 0; l < 32; l++)
            lanes.emplace_back([=, &body]() {
                threadIdx.x = w * 32 + l;
                blockIdx.x = block;
                body();
            });
        for (auto &t : lanes) t.join();
    }
}
}  // namespace warp_emu

// every lane of the (full) warp must call these the same number of times -- true of the kernels run here, whose control flow
// is warp-uniform
template <typename T> inline T __shfl_sync(unsigned, T v, int srcLane) {
    static_assert(sizeof(T) <= 8, "shuffles move at most 64 bits");
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    warp_emu::gExchange[threadIdx.x & 31] = raw;
    warp_emu::gBarrier.wait();
    raw = warp_emu::gExchange[srcLane & 31];
    warp_emu::gBarrier.wait();
    T r;
    std::memcpy(&r, &raw, sizeof(T));
    return r;
}
inline void __syncwarp() { warp_emu::gBarrier.wait(); }
template <typename T> inline T __ldg(const T *p) { return *p; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
// funnel shift right: the low 32 bits of (hi:lo) >> (shift & 31)
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31));
}
