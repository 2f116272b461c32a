// simt.h — a one-warp SIMT emulator for the CPU build of the lane-cooperative kernels
// (planar_coop.cuh): the 32 lanes of a warp run as ucontext coroutines inside one OS thread and
// meet at every warp collective (__shfl_sync, __ballot_sync, __syncwarp ...), so the very same
// device source executes on the CPU with its cross-lane data flow intact.  TEST TOOL ONLY
// (tools/host_emu, tests/test_kernel_host_emulation.py): never part of libdartb.so.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

namespace simt {

struct Dim { unsigned x, y, z; };

struct Warp {
    static constexpr int NL = 32;
    ucontext_t sched;
    ucontext_t ctx[NL];
    char* stack[NL];
    bool done[NL];
    bool waiting[NL];
    uint64_t slot[2][NL];
    int gen;
    int cur;
    unsigned block_id, grid_dim;
    char* shared;
    std::function<void()> body;
};

inline Warp*& W() { static thread_local Warp* w = nullptr; return w; }

inline Dim cur_thread() { return Dim{(unsigned)W()->cur, 0, 0}; }
inline Dim cur_block() { return Dim{W()->block_id, 0, 0}; }
inline Dim cur_blockdim() { return Dim{32, 1, 1}; }
inline Dim cur_griddim() { return Dim{W()->grid_dim, 1, 1}; }
inline void* shared_ptr() { return W()->shared; }

// every live lane of the warp meets here; returns the generation index whose slots were just filled
inline int meet(uint64_t bits) {
    Warp* w = W();
    const int g = w->gen & 1;
    w->slot[g][w->cur] = bits;
    w->waiting[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
    return g;
}

inline void trampoline() {
    Warp* w = W();
    w->body();
    w->done[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

// run `body` as one block of 32 threads (one warp) per block id
inline void launch(unsigned grid, size_t shared_bytes, const std::function<void()>& body) {
    static thread_local Warp* w = nullptr;
    constexpr size_t STK = 1 << 20;
    if (!w) {
        w = new Warp();
        for (int i = 0; i < Warp::NL; i++) w->stack[i] = (char*)malloc(STK);
        w->shared = nullptr;
    }
    W() = w;
    w->body = body;
    w->grid_dim = grid;
    w->shared = (char*)realloc(w->shared, shared_bytes ? shared_bytes : 16);
    for (unsigned b = 0; b < grid; b++) {
        w->block_id = b;
        w->gen = 0;
        memset(w->shared, 0, shared_bytes);
        for (int i = 0; i < Warp::NL; i++) {
            w->done[i] = false;
            w->waiting[i] = false;
            getcontext(&w->ctx[i]);
            w->ctx[i].uc_stack.ss_sp = w->stack[i];
            w->ctx[i].uc_stack.ss_size = STK;
            w->ctx[i].uc_link = &w->sched;
            makecontext(&w->ctx[i], (void (*)())trampoline, 0);
        }
        for (;;) {
            int live = 0;
            for (int i = 0; i < Warp::NL; i++) {
                if (w->done[i] || w->waiting[i]) continue;
                w->cur = i;
                swapcontext(&w->sched, &w->ctx[i]);
            }
            int nwait = 0;
            for (int i = 0; i < Warp::NL; i++) { if (!w->done[i]) live++; if (w->waiting[i]) nwait++; }
            if (live == 0) break;
            if (nwait == live) {  // the collective completes: release everyone into the next generation
                for (int i = 0; i < Warp::NL; i++) w->waiting[i] = false;
                w->gen++;
            } else if (nwait == 0) {
                continue;
            } else {
                // some lanes finished while others wait at a collective: a partial-warp collective;
                // treat the finished lanes as not participating
                for (int i = 0; i < Warp::NL; i++) w->waiting[i] = false;
                w->gen++;
            }
        }
    }
}

template <typename T>
inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T>
inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

#define threadIdx (simt::cur_thread())
#define blockIdx (simt::cur_block())
#define blockDim (simt::cur_blockdim())
#define gridDim (simt::cur_griddim())

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int me = simt::W()->cur;
    const int g = simt::meet(simt::to_bits(v));
    const int s = (me & ~(width - 1)) | (src & (width - 1));
    return simt::from_bits<T>(simt::W()->slot[g][s]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    const int me = simt::W()->cur;
    const int g = simt::meet(simt::to_bits(v));
    const int s = me ^ m;
    (void)width;
    return simt::from_bits<T>(simt::W()->slot[g][s]);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
    const int me = simt::W()->cur;
    const int g = simt::meet(simt::to_bits(v));
    int s = me + d;
    if ((s & ~(width - 1)) != (me & ~(width - 1))) s = me;
    return simt::from_bits<T>(simt::W()->slot[g][s]);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    const int g = simt::meet(pred ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (!simt::W()->done[i] && simt::W()->slot[g][i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::meet(0); }
inline void __syncthreads() { simt::meet(0); }
inline unsigned __activemask() { return 0xffffffffu; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
