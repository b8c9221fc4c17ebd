// cuemu.cpp — fiber scheduler behind cuemu.h (debug scaffolding, see the header).
#include "cuemu.h"

namespace cuemu {

dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

namespace {
enum State { RUNNABLE, AT_BLOCK, AT_WARP, DONE };
struct Fiber {
    ucontext_t ctx;
    State st;
    dim3 tid;
    char* stack;
};
constexpr size_t kStack = 256 * 1024;
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* cur_body = nullptr;
std::vector<unsigned char> warp_slots;   // per warp: 32 x 16 bytes

void trampoline() {
    (*cur_body)();
    fibers[cur].st = DONE;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}

void yield_as(State s) {
    fibers[cur].st = s;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

void block_barrier() { yield_as(AT_BLOCK); }
void warp_barrier() { yield_as(AT_WARP); }
int lane_id() { return cur & 31; }
int warp_id() { return cur >> 5; }
void* warp_slot() { return warp_slots.data() + (size_t)(cur >> 5) * 32 * 16; }

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) { std::fprintf(stderr, "cuemu: bad block size %d\n", nthreads); std::abort(); }
    std::vector<unsigned char> dyn(smem + 64);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    g_blockDim = block;
    g_gridDim = grid;
    cur_body = &body;
    if ((int)fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < fibers.size(); ++i) fibers[i].stack = (char*)std::malloc(kStack);
    }
    warp_slots.assign((size_t)((nthreads + 31) / 32) * 32 * 16, 0);
    const int nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = dim3(bx, by, bz);
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &sched_ctx;
            makecontext(&f.ctx, trampoline, 0);
            f.st = RUNNABLE;
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        }
        for (;;) {
            bool progressed = false, all_done = true;
            for (int t = 0; t < nthreads; ++t) {
                if (fibers[t].st == RUNNABLE) {
                    cur = t;
                    g_threadIdx = fibers[t].tid;
                    swapcontext(&sched_ctx, &fibers[t].ctx);
                    progressed = true;
                }
                if (fibers[t].st != DONE) all_done = false;
            }
            if (all_done) break;
            bool released = false;
            for (int w = 0; w < nwarps; ++w) {      // warp-level barriers
                int lo = w * 32, hi = std::min(nthreads, lo + 32);
                bool any = false, ok = true;
                for (int t = lo; t < hi; ++t) {
                    if (fibers[t].st == AT_WARP) any = true;
                    else if (fibers[t].st != DONE) ok = false;
                }
                if (any && ok) {
                    for (int t = lo; t < hi; ++t) if (fibers[t].st == AT_WARP) fibers[t].st = RUNNABLE;
                    released = true;
                }
            }
            if (!released) {                         // block-level barrier
                bool any = false, ok = true;
                for (int t = 0; t < nthreads; ++t) {
                    if (fibers[t].st == AT_BLOCK) any = true;
                    else if (fibers[t].st != DONE) ok = false;
                }
                if (any && ok) {
                    for (int t = 0; t < nthreads; ++t) if (fibers[t].st == AT_BLOCK) fibers[t].st = RUNNABLE;
                    released = true;
                }
            }
            if (!progressed && !released) {
                std::fprintf(stderr, "cuemu: deadlock (divergent barrier) in block (%u,%u,%u)\n", bx, by, bz);
                std::abort();
            }
        }
    }
    cur = -1;
    g_dyn_smem = nullptr;
}

}  // namespace cuemu
