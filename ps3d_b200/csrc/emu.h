// TEST-ONLY CPU model of a CUDA thread block (used when compiling the kernel
// sources with plain g++ and -DPS3D_EMU).  Every CUDA thread of a block is a
// ucontext fibre on one OS thread; __syncthreads() yields to the next fibre, so
// a barrier is one round-robin turn.  Blocks run in parallel over OpenMP
// threads.  This exists so that the index arithmetic of the kernels can be
// exercised by `pytest -m "not gpu"` in a container without a GPU; it is not a
// fallback and is never linked into libps3d_cuda.so.
#pragma once

#include <ucontext.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

namespace emu {

struct BlockState {
    std::vector<ucontext_t> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    ucontext_t sched;
    int cur = 0;
    std::vector<char> smem;
    std::vector<double> xchg;          // warp-shuffle exchange slots (one per thread)
    const std::function<void()>* body = nullptr;
    ~BlockState() { for (char* s : stacks) free(s); }
};

extern thread_local BlockState* tl_bs;
extern thread_local dim3 tl_threadIdx, tl_blockIdx, tl_blockDim, tl_gridDim;

void fibre_main();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline void* smem_base() { return tl_bs->smem.data(); }
inline void yield() {
    BlockState* bs = tl_bs;
    swapcontext(&bs->ctx[bs->cur], &bs->sched);
}

}  // namespace emu

#define threadIdx (::emu::tl_threadIdx)
#define blockIdx (::emu::tl_blockIdx)
#define blockDim (::emu::tl_blockDim)
#define gridDim (::emu::tl_gridDim)

static inline void __syncthreads() { ::emu::yield(); }
// Warp shuffles: every thread of the block must execute the same shuffle sequence (true for the kernels
// here); a shuffle is two barrier rounds through a per-thread exchange slot.
static inline int emu_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }
static inline double __shfl_up_sync(unsigned, double v, unsigned delta, int width = 32) {
    ::emu::BlockState* bs = ::emu::tl_bs;
    const int t = emu_tid();
    bs->xchg[t] = v;
    ::emu::yield();
    const int lane = t % width;
    const double r = (lane >= (int)delta) ? bs->xchg[t - delta] : v;
    ::emu::yield();
    return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int mask, int width = 32) {
    ::emu::BlockState* bs = ::emu::tl_bs;
    const int t = emu_tid();
    bs->xchg[t] = v;
    ::emu::yield();
    const int lane = t % width;
    const int src = t - lane + ((lane ^ mask) % width);
    const double r = (src < (int)bs->xchg.size()) ? bs->xchg[src] : v;
    ::emu::yield();
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline void __threadfence_system() {}
static inline double atomicAdd(double* p, double v) {
    double old;
#pragma omp atomic capture
    { old = *p; *p += v; }
    return old;
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

#ifdef PS3D_EMU_IMPL
namespace emu {
thread_local BlockState* tl_bs = nullptr;
thread_local dim3 tl_threadIdx, tl_blockIdx, tl_blockDim, tl_gridDim;

void fibre_main() {
    BlockState* bs = tl_bs;
    (*bs->body)();
    bs = tl_bs;
    bs->done[bs->cur] = 1;
    swapcontext(&bs->ctx[bs->cur], &bs->sched);
}

static void run_block(BlockState& bs, dim3 bidx, dim3 grid, dim3 block, size_t smem,
                      const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    const size_t stack_sz = 256 * 1024;
    if ((int)bs.ctx.size() < nt) {
        size_t old = bs.ctx.size();
        bs.ctx.resize(nt);
        bs.stacks.resize(nt, nullptr);
        for (size_t i = old; i < (size_t)nt; ++i) bs.stacks[i] = (char*)malloc(stack_sz);
    }
    bs.done.assign(nt, 0);
    bs.smem.assign(smem + 64, 0);
    bs.xchg.assign(nt, 0.0);
    bs.body = &body;
    tl_bs = &bs;
    tl_blockIdx = bidx; tl_blockDim = block; tl_gridDim = grid;
    for (int t = 0; t < nt; ++t) {
        getcontext(&bs.ctx[t]);
        bs.ctx[t].uc_stack.ss_sp = bs.stacks[t];
        bs.ctx[t].uc_stack.ss_size = stack_sz;
        bs.ctx[t].uc_link = &bs.sched;
        makecontext(&bs.ctx[t], (void (*)())fibre_main, 0);
    }
    int remaining = nt;
    while (remaining > 0) {
        for (int t = 0; t < nt; ++t) {
            if (bs.done[t]) continue;
            bs.cur = t;
            tl_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            swapcontext(&bs.sched, &bs.ctx[t]);
            if (bs.done[t]) --remaining;
        }
    }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const long nb = (long)grid.x * grid.y * grid.z;
#pragma omp parallel
    {
        BlockState bs;
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < nb; ++b) {
            dim3 bidx((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y),
                      (unsigned)(b / ((long)grid.x * grid.y)));
            run_block(bs, bidx, grid, block, smem, body);
        }
    }
}
}  // namespace emu
#endif
