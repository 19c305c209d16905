// simt.h — TEST HARNESS ONLY.  A minimal single-block SIMT emulator: one fiber (ucontext) per CUDA thread, so that device
// code using threadIdx / blockDim, __syncthreads(), extern __shared__ memory and full-warp shuffles runs unmodified on the
// CPU-only build container.  Semantics implemented (enough for the block-per-reactor N2 solver, pbr_k4b_body):
//   * __syncthreads(): every live thread of the block waits for all the others;
//   * __shfl_sync / __shfl_up_sync / __shfl_xor_sync with the full mask: all 32 lanes of the warp exchange one value
//     (the callers only shuffle in warp-uniform control flow; blockDim.x is a multiple of 32);
//   * blocks of a grid run one after the other (simt_launch), shared memory is one static buffer.
// A thread that waits while no other thread can run (a barrier not reached by everybody) aborts the process with a message
// instead of hanging.  Not a performance model; never part of the product.
#pragma once
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct SimtDim3 { unsigned x, y, z; };

namespace simt {
enum { RUN = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
struct Fiber { ucontext_t ctx; std::vector<char> stack; int state; };
static std::vector<Fiber> fibers;
static ucontext_t sched_ctx;
static int cur = -1, nthreads = 0;
static SimtDim3 block_idx = {0, 0, 0}, block_dim = {1, 1, 1}, grid_dim = {1, 1, 1};
static double slots[64][32];                      // per warp: the values being shuffled
// Order in which runnable threads are resumed inside one scheduling round: 0 ascending, 1 descending, 2 a fresh pseudo-random
// permutation per round.  A data race -- two threads touching the same shared/global location, one of them writing, with no
// barrier or warp synchronisation in between -- makes the result depend on this order; the tests run every block kernel under
// all three and demand bit-identical results (the pool's compute-sanitizer is closed, this is the race check we can run).
static int order_mode = 0;
static unsigned long long order_state = 88172645463325252ull;
static std::vector<int> perm;
static void next_perm() {
    perm.resize(nthreads);
    for (int t = 0; t < nthreads; ++t) perm[t] = order_mode == 1 ? nthreads - 1 - t : t;
    if (order_mode == 2)
        for (int t = nthreads - 1; t > 0; --t) {
            order_state ^= order_state << 13; order_state ^= order_state >> 7; order_state ^= order_state << 17;
            const int j = (int)(order_state % (unsigned long long)(t + 1));
            const int tmp = perm[t]; perm[t] = perm[j]; perm[j] = tmp;
        }
}
static std::function<void()>* body = nullptr;

static void trampoline() { (*body)(); fibers[cur].state = DONE; swapcontext(&fibers[cur].ctx, &sched_ctx); }

static void yield(int state) { fibers[cur].state = state; swapcontext(&fibers[cur].ctx, &sched_ctx); }

static void run_block(std::function<void()> f) {
    body = &f;
    fibers.assign(nthreads, Fiber());
    for (int t = 0; t < nthreads; ++t) {
        Fiber& fb = fibers[t];
        fb.stack.resize(512*1024);
        getcontext(&fb.ctx);
        fb.ctx.uc_stack.ss_sp = fb.stack.data();
        fb.ctx.uc_stack.ss_size = fb.stack.size();
        fb.ctx.uc_link = &sched_ctx;
        makecontext(&fb.ctx, trampoline, 0);
        fb.state = RUN;
    }
    for (;;) {
        bool progressed = false;
        next_perm();
        for (int q = 0; q < nthreads; ++q) {
            const int t = perm[q];
            if (fibers[t].state == RUN) { cur = t; swapcontext(&sched_ctx, &fibers[t].ctx); progressed = true; }
        }
        // release complete warps
        for (int w = 0; w < nthreads/32; ++w) {
            int at = 0;
            for (int l = 0; l < 32; ++l) at += fibers[w*32 + l].state == AT_WARP;
            if (at == 32) { for (int l = 0; l < 32; ++l) fibers[w*32 + l].state = RUN; progressed = true; }
        }
        int at_block = 0, done = 0;
        for (int t = 0; t < nthreads; ++t) { at_block += fibers[t].state == AT_BLOCK; done += fibers[t].state == DONE; }
        if (done == nthreads) break;
        if (at_block && at_block + done == nthreads) { for (auto& fb : fibers) if (fb.state == AT_BLOCK) fb.state = RUN; progressed = true; }
        if (!progressed) { fprintf(stderr, "simt: deadlock (%d at block barrier, %d done of %d threads)\n", at_block, done, nthreads); abort(); }
    }
    cur = -1;
}

// grid of `grid` blocks with `block` threads each (multiple of 32), one block after the other
static void launch(unsigned grid, unsigned block, std::function<void()> f) {
    if (block % 32 || block > 64*32) { fprintf(stderr, "simt: block size %u\n", block); abort(); }
    grid_dim = {grid, 1, 1}; block_dim = {block, 1, 1}; nthreads = (int)block;
    for (unsigned b = 0; b < grid; ++b) { block_idx = {b, 0, 0}; run_block(f); }
}
}  // namespace simt

static inline SimtDim3 simt_thread_idx() { return SimtDim3{(unsigned)simt::cur, 0, 0}; }
#define threadIdx (simt_thread_idx())
#define blockIdx (simt::block_idx)
#define blockDim (simt::block_dim)
#define gridDim (simt::grid_dim)

#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__            /* `extern __shared__ double pbr_smem[]`: the driver defines the array; static shared arrays use PBR_SHARED */
#define PBR_SHARED static     /* one block runs at a time, so a function-local static IS the block's shared array */
#define __launch_bounds__(...)

static inline void __syncthreads() { simt::yield(simt::AT_BLOCK); }

template <class T> static inline T simt_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= sizeof(double), "simt: shuffle of a type wider than 8 bytes");
    const int w = simt::cur >> 5, l = simt::cur & 31;
    double box = 0.0; memcpy(&box, &v, sizeof(T));
    simt::slots[w][l] = box;
    simt::yield(simt::AT_WARP);                     // everybody has deposited
    T r; memcpy(&r, &simt::slots[w][src_lane & 31], sizeof(T));
    simt::yield(simt::AT_WARP);                     // everybody has read: the slots may be reused
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return simt_exchange(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int l = simt::cur & 31;
    const T r = simt_exchange(v, l >= (int)d ? l - (int)d : l);
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return simt_exchange(v, (simt::cur & 31) ^ m); }

static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
