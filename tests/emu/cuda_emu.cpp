// Fiber scheduler behind tests/emu/cuda_emu.h (test scaffolding only, see that header).
#include "cuda_emu.h"

namespace emu {
uint3 threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;
char* dyn_smem = nullptr;

namespace {
enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber {
    ucontext_t ctx;
    State st;
    uint3 tid;
    int lin;
    char* stack;
};
constexpr size_t kStack = 256 * 1024;
std::vector<Fiber> fibers;
std::vector<uint64_t> slots;      // 32 per warp
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* cur_body = nullptr;

void trampoline() {
    (*cur_body)();
    fibers[cur].st = DONE;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}

void yield(State s) {
    fibers[cur].st = s;
    swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

void barrier() { yield(WAIT_BLOCK); }
void warp_barrier() { yield(WAIT_WARP); }
int lane_id() { return fibers[cur].lin & 31; }
uint64_t* warp_slot(int lane) { return &slots[(size_t)(fibers[cur].lin >> 5) * 32 + lane]; }
int warp_live_lanes() {
    int n = (int)fibers.size();
    int w = fibers[cur].lin >> 5;
    return std::min(32, n - w * 32);
}

static void run_block(int nthreads, const std::function<void()>& body) {
    cur_body = &body;
    slots.assign((size_t)((nthreads + 31) / 32) * 32, 0);
    for (int i = 0; i < nthreads; ++i) {
        Fiber& f = fibers[i];
        f.st = RUNNABLE;
        f.lin = i;
        f.tid.x = i % blockDim_.x;
        f.tid.y = (i / blockDim_.x) % blockDim_.y;
        f.tid.z = i / (blockDim_.x * blockDim_.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
    }
    int done = 0;
    while (done < nthreads) {
        bool progressed = false;
        for (int i = 0; i < nthreads; ++i) {
            if (fibers[i].st != RUNNABLE) continue;
            progressed = true;
            cur = i;
            threadIdx_ = fibers[i].tid;
            swapcontext(&sched_ctx, &fibers[i].ctx);
            if (fibers[i].st == DONE) ++done;
        }
        // release warp barriers whose live lanes have all arrived (exited lanes do not count)
        int nwarps = (nthreads + 31) / 32;
        for (int w = 0; w < nwarps; ++w) {
            int lo = w * 32, hi = std::min(nthreads, lo + 32), waiting = 0, live = 0;
            for (int i = lo; i < hi; ++i) {
                if (fibers[i].st != DONE) ++live;
                if (fibers[i].st == WAIT_WARP) ++waiting;
            }
            if (live > 0 && waiting == live) {
                for (int i = lo; i < hi; ++i)
                    if (fibers[i].st == WAIT_WARP) fibers[i].st = RUNNABLE;
                progressed = true;
            }
        }
        int waiting = 0, live = 0;
        for (int i = 0; i < nthreads; ++i) {
            if (fibers[i].st != DONE) ++live;
            if (fibers[i].st == WAIT_BLOCK) ++waiting;
        }
        if (live > 0 && waiting == live) {
            for (int i = 0; i < nthreads; ++i)
                if (fibers[i].st == WAIT_BLOCK) fibers[i].st = RUNNABLE;
            progressed = true;
        }
        if (!progressed && done < nthreads) {
            std::fprintf(stderr, "cuda_emu: deadlock (divergent barrier?) in block (%u,%u,%u): %d live\n",
                         blockIdx_.x, blockIdx_.y, blockIdx_.z, live);
            std::abort();
        }
    }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    int nthreads = (int)(block.x * block.y * block.z);
    if ((int)fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < fibers.size(); ++i) fibers[i].stack = (char*)std::malloc(kStack);
    }
    std::vector<Fiber> saved;   // keep only nthreads visible for warp_live_lanes()
    if ((int)fibers.size() > nthreads) {
        saved.assign(fibers.begin() + nthreads, fibers.end());
        fibers.resize(nthreads);
    }
    std::vector<char> sm(smem + 16);
    dyn_smem = sm.data();
    blockDim_ = block;
    gridDim_ = grid;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                blockIdx_ = {x, y, z};
                run_block(nthreads, body);
            }
    fibers.insert(fibers.end(), saved.begin(), saved.end());
    dyn_smem = nullptr;
}
}  // namespace emu
