// cuda_shim.cpp -- fiber scheduler + runtime stubs of the test-only CUDA shim
// (see cuda_shim.h: not part of the product).
#include "cuda_shim.h"

#include <stdio.h>
#include <time.h>
#include <ucontext.h>

#include <vector>
#include <unordered_map>
#include <string.h>
#include <stdlib.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace hostsim {

static const size_t STACK = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char *stack;
    bool done;
    uint3 tid;
};

struct Warp {
    uint64_t buf[32];
    unsigned ballot;
    int arrive;
    unsigned gen;
    int alive;
};

static std::vector<Fiber> fibers;
static std::vector<Warp> warps;
static std::vector<char *> stack_pool;
static ucontext_t sched_ctx;
static int cur = -1;
static int n_alive = 0, blk_arrive = 0;
static unsigned blk_gen = 0;
static const std::function<void()> *cur_body = nullptr;
static std::vector<unsigned char> dsm;
static long n_launch = 0;

static void yield_to_sched() { swapcontext(&fibers[cur].ctx, &sched_ctx); }

static void fiber_main()
{
    (*cur_body)();
    Fiber &f = fibers[cur];
    f.done = true;
    n_alive--;
    Warp &w = warps[cur / 32];
    w.alive--;
    // an exited thread no longer takes part in barriers: release waiters if it was the last one missing
    if (n_alive > 0 && blk_arrive == n_alive) { blk_arrive = 0; blk_gen++; }
    if (w.alive > 0 && w.arrive == w.alive) { w.arrive = 0; w.gen++; }
    swapcontext(&f.ctx, &sched_ctx);
}

void sync_block()
{
    unsigned g = blk_gen;
    if (++blk_arrive == n_alive) { blk_arrive = 0; blk_gen++; return; }
    while (blk_gen == g) yield_to_sched();
}

static void warp_barrier(Warp &w)
{
    unsigned g = w.gen;
    if (++w.arrive == w.alive) { w.arrive = 0; w.gen++; return; }
    while (w.gen == g) yield_to_sched();
}

uint64_t warp_exchange(uint64_t v, int src)
{
    Warp &w = warps[cur / 32];
    int lane = cur & 31;
    w.buf[lane] = v;
    warp_barrier(w);
    uint64_t r = (src >= 0 && src < 32 && (size_t)((cur & ~31) + src) < fibers.size()) ? w.buf[src] : v;
    warp_barrier(w);
    return r;
}

unsigned warp_ballot(int pred)
{
    Warp &w = warps[cur / 32];
    int lane = cur & 31;
    if (w.arrive == 0) w.ballot = 0;
    if (pred) w.ballot |= 1u << lane;
    warp_barrier(w);
    unsigned r = w.ballot;
    warp_barrier(w);
    return r;
}

unsigned warp_gather(uint64_t v, uint64_t *out32)
{
    Warp &w = warps[cur / 32];
    int lane = cur & 31;
    if (w.arrive == 0) w.ballot = 0;
    w.buf[lane] = v;
    w.ballot |= 1u << lane;
    warp_barrier(w);
    unsigned live = w.ballot;
    for (int l = 0; l < 32; l++) out32[l] = w.buf[l];
    warp_barrier(w);
    return live;
}

void *dyn_smem() { return dsm.data(); }
long launches() { return n_launch; }

// Out-of-bounds stores: 64 guard bytes behind the dynamic shared memory and behind every device allocation carry a
// pattern that is checked after every launch (the GPU box has no compute-sanitizer; this is the memcheck the kernels get).
static const unsigned char GUARD = 0xA5;
static std::unordered_map<void *, size_t> &allocs()
{
    static std::unordered_map<void *, size_t> m;
    return m;
}
void guard_alloc(void *p, size_t n) { memset((char *)p + n, GUARD, 64); allocs()[p] = n; }
void guard_free(void *p) { allocs().erase(p); }
static void check_guards(const char *what, size_t smem)
{
    for (size_t i = 0; i < 64; i++)
        if (dsm[smem + i] != GUARD) {
            fprintf(stderr, "hostsim: store behind the dynamic shared memory (%zu bytes, offset +%zu) in launch %ld (%s)\n", smem, i, n_launch, what);
            abort();
        }
    for (auto &kv : allocs())
        for (size_t i = 0; i < 64; i++)
            if (((unsigned char *)kv.first)[kv.second + i] != GUARD) {
                fprintf(stderr, "hostsim: store behind a device allocation of %zu bytes (offset +%zu) in launch %ld (%s)\n", kv.second, i, n_launch, what);
                abort();
            }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    n_launch++;
    size_t nt = (size_t)block.x * block.y * block.z;
    if (nt == 0 || (size_t)grid.x * grid.y * grid.z == 0) return;
    if (dsm.size() < smem + 64) dsm.resize(smem + 64);
    memset(dsm.data() + smem, GUARD, 64);
    while (stack_pool.size() < nt) stack_pool.push_back((char *)malloc(STACK));
    gridDim = grid;
    blockDim = block;
    cur_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                fibers.assign(nt, Fiber());
                warps.assign((nt + 31) / 32, Warp());
                n_alive = (int)nt; blk_arrive = 0;
                for (size_t i = 0; i < nt; i++) {
                    Fiber &f = fibers[i];
                    f.done = false;
                    f.stack = stack_pool[i];
                    f.tid.x = (unsigned)(i % block.x);
                    f.tid.y = (unsigned)((i / block.x) % block.y);
                    f.tid.z = (unsigned)(i / ((size_t)block.x * block.y));
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, fiber_main, 0);
                    warps[i / 32].alive++;
                }
                int remaining = (int)nt;
                while (remaining > 0) {
                    int progressed = 0;
                    for (size_t i = 0; i < nt; i++) {
                        if (fibers[i].done) continue;
                        cur = (int)i;
                        threadIdx = fibers[i].tid;
                        swapcontext(&sched_ctx, &fibers[i].ctx);
                        progressed++;
                        if (fibers[i].done) remaining--;
                    }
                    if (!progressed) break;
                }
            }
    cur = -1;
    cur_body = nullptr;
    check_guards("after the last block", smem);
}
} // namespace hostsim

// ---- runtime stubs ----
struct hs_event { double t; };
static double now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = calloc(1, (n ? n : 1) + 64);
    if (!*p) return cudaErrorMemoryAllocation;
    hostsim::guard_alloc(*p, n ? n : 1);
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { if (p) hostsim::guard_free(p); free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (void *)1; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "hostsim"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    p->multiProcessorCount = 4; p->totalGlobalMem = (size_t)8 << 30; p->major = 10; p->minor = 0;
    p->sharedMemPerBlockOptin = 227 * 1024;
    strcpy(p->name, "hostsim");
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new hs_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new hs_event(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; } // launches run synchronously here
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
