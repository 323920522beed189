// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal host-side stand-in for the CUDA runtime so that uxsim_b200/csrc/engine.cu can be compiled with
// g++ (-DUX_EMU) in the GPU-less development container and its kernel *logic* exercised under
// ASan/UBSan/valgrind against the oracle (tests/test_emu_parity.py).  Kernels run as nested loops over
// (block, thread); every kernel in engine.cu is written so that its result does not depend on the order in
// which threads execute (that is also what makes it deterministic on the GPU), and the few warp/block
// cooperative helpers have an emulation branch that reproduces the same arithmetic order.
//
// The emulation library is built into tests/emu/_build/ and is never loaded by the uxsim_b200 package:
// uxsim_b200._capi.load_product() only ever opens uxsim_b200/libuxsim_b200.so (the nvcc build), and fails
// if that is missing.  Nothing here is a fallback path.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <cstdio>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static dim3 blockIdx, threadIdx, blockDim, gridDim;
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaGraph_t;
typedef void *cudaGraphExec_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamCaptureModeThreadLocal = 0, cudaStreamNonBlocking = 1 };

static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
// graphs are not emulated: the emulation build launches kernels eagerly (UX_EMU disables capture)

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicCAS(int *p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}

// Launch: run every (block, thread) sequentially.
#define UX_LAUNCH(kernel, grid, block, stream, ...)                                   \
    do {                                                                              \
        dim3 g_ = (grid), b_ = (block);                                               \
        gridDim = g_; blockDim = b_;                                                  \
        for (unsigned bz_ = 0; bz_ < g_.z; bz_++)                                     \
        for (unsigned by_ = 0; by_ < g_.y; by_++)                                     \
            for (unsigned bx_ = 0; bx_ < g_.x; bx_++)                                 \
                for (unsigned ty_ = 0; ty_ < b_.y; ty_++)                             \
                    for (unsigned tx_ = 0; tx_ < b_.x; tx_++) {                       \
                        blockIdx = dim3(bx_, by_, bz_); threadIdx = dim3(tx_, ty_, 0); \
                        kernel(__VA_ARGS__);                                          \
                    }                                                                 \
    } while (0)
