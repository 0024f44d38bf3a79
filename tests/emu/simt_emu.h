// Host emulation of the CUDA SIMT model for tests (tests/test_emu_kernels.py): the SIMT kernel SOURCES of
// mvster_b200/csrc are compiled with g++ and run on CPU threads - one OS thread per CUDA thread of a block, blocks one after
// another - so that kernel logic can be checked against the oracle without a GPU.  Warp collectives (__all_sync,
// __shfl_xor_sync) and __syncthreads are real barriers between those threads.  Test infrastructure only: nothing in the
// product path includes this header, and tcgen05 / TMA kernels are out of its reach.
#pragma once
#define MVSTER_CPU_EMU 1
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

namespace emu {
struct Warp {
    std::barrier<> bar{32};
    int vote[32];
    float f[32];
    explicit Warp(int n) : bar(n) {}
};
struct Ctx { Warp* warp; std::barrier<>* block; int lane; unsigned char* dyn_smem; };
inline thread_local Ctx ctx;
inline unsigned long long pack(float lo, float hi) {
    uint32_t a, b;
    std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4);
    return (unsigned long long)a | ((unsigned long long)b << 32);
}
inline float2 unpack(unsigned long long v) {
    const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32);
    float2 r; std::memcpy(&r.x, &a, 4); std::memcpy(&r.y, &b, 4);
    return r;
}
}  // namespace emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emu::ctx.block->arrive_and_wait(); }
inline void __syncwarp() { emu::ctx.warp->bar.arrive_and_wait(); }
inline bool __all_sync(unsigned, bool p) {
    emu::Warp& w = *emu::ctx.warp;
    w.vote[emu::ctx.lane] = p;
    w.bar.arrive_and_wait();
    bool r = true;
    for (int i = 0; i < 32; ++i) r = r && w.vote[i];
    w.bar.arrive_and_wait();
    return r;
}
inline float __shfl_xor_sync(unsigned, float v, int o) {
    emu::Warp& w = *emu::ctx.warp;
    w.f[emu::ctx.lane] = v;
    w.bar.arrive_and_wait();
    const float r = w.f[emu::ctx.lane ^ o];
    w.bar.arrive_and_wait();
    return r;
}
inline int __shfl_sync(unsigned, int v, int src) {
    emu::Warp& w = *emu::ctx.warp;
    w.vote[emu::ctx.lane] = v;
    w.bar.arrive_and_wait();
    const int r = w.vote[src & 31];
    w.bar.arrive_and_wait();
    return r;
}
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.f / a; }
inline int __float2int_rn(float a) { return (int)std::lrintf(a); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __expf(float a) { return std::exp(a); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v, std::memory_order_relaxed); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }

namespace emu {
// run kernel(args...) over the grid: blocks sequentially, the threads of a block concurrently (block size a multiple of 32).
// The OS threads are created once per launch and walk the blocks together; the per-block barrier objects are fresh for every
// block because a thread that leaves the kernel early drops out of them.
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    const int nthreads = (int)(block.x * block.y * block.z), nwarps = nthreads / 32;
    const long long nblocks = (long long)grid.x * grid.y * grid.z;
    std::vector<unsigned char> smem_store(smem_bytes + 128);
    unsigned char* smem = smem_store.data() + (128 - (reinterpret_cast<uintptr_t>(smem_store.data()) & 127)) % 128;
    std::barrier<> start_bar(nthreads), end_bar(nthreads);
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<Warp>> warps(nwarps);
    auto worker = [&](int t) {
        for (long long b = 0; b < nblocks; ++b) {
            if (t == 0) {
                block_bar.reset(new std::barrier<>(nthreads));
                for (int w = 0; w < nwarps; ++w) warps[w].reset(new Warp(32));
            }
            start_bar.arrive_and_wait();
            threadIdx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            blockIdx = {(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y))};
            blockDim = block; gridDim = grid;
            ctx = {warps[t / 32].get(), block_bar.get(), t % 32, smem};
            body();
            block_bar->arrive_and_drop();          // a thread that has left the kernel no longer takes part in its barriers
            warps[t / 32]->bar.arrive_and_drop();
            end_bar.arrive_and_wait();             // nobody still uses this block's barrier objects
        }
    };
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
}
}  // namespace emu
