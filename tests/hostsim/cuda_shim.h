// cuda_shim.h -- DEVELOPMENT/TEST-ONLY stand-in for the CUDA language extensions
// and the handful of runtime calls the product sources use, so that the SAME
// kernel sources (relagn_b200/csrc/*.cu) can be compiled with g++ and stepped
// through on a machine without a GPU (the build container has none).
//
// This is NOT a CPU fallback of the product: relagn_b200/ never loads the
// library built from it (tests/hostsim/build.py -> tests/hostsim/_build/), the
// shipped librelagn_b200.so is nvcc-built sm_100a code only and fails with
// RG_E_CUDA without a device.  It exists so kernel LOGIC can be debugged against
// the oracle before GPU minutes are spent; performance is irrelevant here.
//
// Execution model: one OS thread; every CUDA thread of a block is a ucontext
// fiber, blocks run one after another.  __syncthreads()/warp shuffles are
// cooperative barriers between the fibers of the running block.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __constant__
#define __restrict__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

namespace hostsim {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void *dyn_smem();
void sync_block();
uint64_t warp_exchange(uint64_t v, int src_lane); // value of lane src_lane (own value if out of range)
unsigned warp_ballot(int pred);
unsigned warp_gather(uint64_t v, uint64_t *out32); // every live lane's value; returns the live-lane mask
long launches();
}

static inline void __syncthreads() { hostsim::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { hostsim::warp_ballot(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

template <class T> static inline T __ldg(const T *p) { return *p; }

template <class T> static inline T hs_shfl(T v, int src)
{
    static_assert(sizeof(T) <= 8, "shfl payload");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    u = hostsim::warp_exchange(u, src);
    T r;
    memcpy(&r, &u, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int w = 32)
{
    int lane = (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
    return hs_shfl(v, (lane & ~(w - 1)) | (src & (w - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int w = 32)
{
    int lane = (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
    (void)w;
    return hs_shfl(v, lane ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int w = 32)
{
    int lane = (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
    int src = lane + (int)d;
    if ((src & ~(w - 1)) != (lane & ~(w - 1))) src = lane;
    return hs_shfl(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int w = 32)
{
    int lane = (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31;
    int src = lane - (int)d;
    if (src < (lane & ~(w - 1))) src = lane;
    return hs_shfl(v, src);
}
static inline unsigned __ballot_sync(unsigned, int p) { return hostsim::warp_ballot(p); }
static inline int __any_sync(unsigned, int p) { return hostsim::warp_ballot(p) != 0; }
static inline int __all_sync(unsigned, int p) { return hostsim::warp_ballot(!p) == 0; }
template <class T> static inline unsigned __match_any_sync(unsigned, T v)
{
    uint64_t u = 0, all[32];
    memcpy(&u, &v, sizeof(T));
    unsigned live = hostsim::warp_gather(u, all), m = 0;
    for (int l = 0; l < 32; l++)
        if (((live >> l) & 1) && all[l] == u) m |= 1u << l;
    return m;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    uint64_t all[32];
    unsigned live = hostsim::warp_gather((uint64_t)(int64_t)v, all);
    int r = v;
    for (int l = 0; l < 32; l++)
        if ((live >> l) & 1) { int x = (int)(int64_t)all[l]; if (x > r) r = x; }
    return r;
}
static inline int __reduce_min_sync(unsigned, int v)
{
    uint64_t all[32];
    unsigned live = hostsim::warp_gather((uint64_t)(int64_t)v, all);
    int r = v;
    for (int l = 0; l < 32; l++)
        if ((live >> l) & 1) { int x = (int)(int64_t)all[l]; if (x < r) r = x; }
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    uint64_t all[32];
    unsigned live = hostsim::warp_gather((uint64_t)v, all);
    unsigned r = v;
    for (int l = 0; l < 32; l++)
        if ((live >> l) & 1) { unsigned x = (unsigned)all[l]; if (x < r) r = x; }
    return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }

// fibers never run concurrently: plain read-modify-write is atomic here
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
static inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
static inline int atomicCAS(int *p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }

static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int __double2int_rd(double x) { return (int)floor(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline long long clock64() { return 0; }

// ---- runtime API subset ----
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct hs_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice,
                      cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; int major, minor; char name[256];
                        size_t sharedMemPerBlockOptin; };

cudaError_t cudaMalloc(void **p, size_t n);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMallocHost((void **)p, n); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaEventCreate(cudaEvent_t *e);
enum { cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
