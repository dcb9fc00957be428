// cuda_emu.hpp -- a few dozen lines of "CUDA on host threads" for CPU-tier tests of SIMPLE kernels (tests/test_kernel_emulation.py).
//
// TEST INFRASTRUCTURE ONLY (like oracle/): nothing under qradiolink_b200/ includes this, and it is no fallback -- it exists so that
// index arithmetic and accumulation order of kernels that use only plain loads / stores, __syncthreads and warp shuffles can be
// checked against the oracle in the container that has no GPU.  One OS thread per CUDA thread, blocks run one after the other,
// __shared__ becomes a function-local static (shared by the block's threads), dynamic shared memory is one heap buffer per block.
// Not emulated: TMA / mbarrier / inline PTX, cooperative groups, atomics, textures -- kernels using those are out of reach.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emu {
inline thread_local dim3 t_idx, b_idx;
inline dim3 b_dim, g_dim;
inline std::barrier<>* blk_bar = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
inline unsigned shfl_buf[64][32];
inline void* dyn_smem = nullptr;
alignas(128) inline unsigned char smem_arena[256 << 10];     // dynamic shared memory of the running block (static: see qrl_tma_emu.hpp)
inline unsigned lin_tid() { return t_idx.x + b_dim.x * (t_idx.y + b_dim.y * t_idx.z); }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& kernel_call)
{
    const unsigned nt = block.x * block.y * block.z;
    b_dim = block; g_dim = grid;
    if (smem_bytes > sizeof(smem_arena)) std::abort();
    dyn_smem = smem_arena;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        std::barrier<> bar(nt);
        blk_bar = &bar;
        warp_bar.clear();
        for (unsigned w = 0; w < (nt + 31) / 32; w++) warp_bar.emplace_back(new std::barrier<>(std::min(32u, nt - 32 * w)));
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                t_idx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                b_idx = dim3(bx, by, bz);
                kernel_call();
                // a thread that returned early must still release the others at later barriers
                blk_bar->arrive_and_drop();
                warp_bar[t >> 5]->arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}
}  // namespace emu

#define threadIdx (emu::t_idx)
#define blockIdx (emu::b_idx)
#define blockDim (emu::b_dim)
#define gridDim (emu::g_dim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define EMU_DYN_SMEM(type, name) type* name = static_cast<type*>(emu::dyn_smem)

static inline void __syncthreads() { emu::blk_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_bar[emu::lin_tid() >> 5]->arrive_and_wait(); }
static inline unsigned __shfl_up_sync(unsigned, unsigned v, int off)
{
    const unsigned t = emu::lin_tid(), w = t >> 5, l = t & 31;
    emu::shfl_buf[w][l] = v;
    emu::warp_bar[w]->arrive_and_wait();
    const unsigned r = (static_cast<int>(l) >= off) ? emu::shfl_buf[w][l - off] : v;
    emu::warp_bar[w]->arrive_and_wait();
    return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
