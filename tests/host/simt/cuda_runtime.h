// tests/host/simt/cuda_runtime.h — a stand-in for the CUDA toolkit header that lets g++ build the product's .cu sources for
// the HOST and run their kernels in a lock-step SIMT emulator (tests/host/simt/simt.cpp).  TEST INFRASTRUCTURE ONLY: the
// product is the nvcc build of the same sources; nothing here is shipped, measured or used as a fallback.
//
// What is emulated
//   * every thread of a block is a fibre with its own stack; blocks run one after the other;
//   * warp collectives (__shfl*_sync, __ballot_sync, __any/__all_sync, __reduce_*_sync, __syncwarp) and __syncthreads are
//     rendez-vous points: a lane blocks there until every lane named in the mask has arrived AT THE SAME SOURCE LINE.
//     Lanes of one mask arriving at different lines, a lane that is not in the mask it passes, or a rendez-vous that can
//     never complete (a collective in divergent code — on the GPU: a hang or garbage) abort the run with a diagnostic;
//   * __shared__ variables are block-wide statics, dynamic shared memory a per-launch buffer;
//   * the CUDA runtime calls the library makes (memory, copies, streams, events) are synchronous host operations.
// What is NOT emulated: timing, the memory model (no races between warps can be observed), asynchronous copies.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <functional>
#include <algorithm>

#define __CUDA_ARCH__ 1000
#ifndef GZB_SIMT_EMULATION
#define GZB_SIMT_EMULATION 1
#endif

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __builtin_assume(x) ((void)0)

// ---------------------------------------------------------------- vector types
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct uint3 { uint32_t x, y, z; };
struct dim3 { uint32_t x, y, z; dim3 (uint32_t a = 1, uint32_t b = 1, uint32_t c = 1) : x (a), y (b), z (c) {} };
static inline uint2 make_uint2 (uint32_t x, uint32_t y) { return uint2{ x, y }; }
static inline uint4 make_uint4 (uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{ x, y, z, w }; }

// ---------------------------------------------------------------- the emulator's interface
namespace simt {
struct Thread { uint3 tid, bid; dim3 bdim, gdim; uint32_t or_gen = 0; };
extern thread_local Thread *cur;                                            // the fibre that is running on this OS thread
void  launch (dim3 grid, dim3 block, size_t dyn_smem, const std::function<void ()> &body);
void *dyn_smem ();
enum Kind { K_SHFL_IDX, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_BALLOT, K_REDUCE_ADD, K_REDUCE_OR, K_REDUCE_AND, K_REDUCE_MIN, K_REDUCE_MAX, K_MATCH_ANY, K_SYNCWARP };
uint64_t collective (const char *file, int line, Kind kind, uint32_t mask, uint64_t v, int arg, int width);
void     syncthreads (const char *file, int line);
int      syncthreads_or (const char *file, int line, int pred);
}
#define threadIdx (simt::cur->tid)
#define blockIdx  (simt::cur->bid)
#define blockDim  (simt::cur->bdim)
#define gridDim   (simt::cur->gdim)
#define warpSize  32
#define SIMT_LAUNCH(kern, grid, block, smem, stream, ...) simt::launch (dim3 grid, dim3 block, (size_t)(smem), [&] { kern (__VA_ARGS__); })

// ---------------------------------------------------------------- warp collectives (macros: the call site identifies the rendez-vous)
namespace simt {
template <class T> inline uint64_t to_bits (T v) { uint64_t u = 0; static_assert (sizeof (T) <= 8, "shuffle of a wide type"); memcpy (&u, &v, sizeof (T)); return u; }
template <class T> inline T from_bits (uint64_t u) { T v; memcpy (&v, &u, sizeof (T)); return v; }
template <class T> inline T shfl (const char *f, int l, Kind k, uint32_t mask, T v, int arg, int width = 32) { return from_bits<T> (collective (f, l, k, mask, to_bits (v), arg, width)); }
}
#define __shfl_sync(...)      simt::shfl (__FILE__, __LINE__, simt::K_SHFL_IDX, __VA_ARGS__)
#define __shfl_up_sync(...)   simt::shfl (__FILE__, __LINE__, simt::K_SHFL_UP, __VA_ARGS__)
#define __shfl_down_sync(...) simt::shfl (__FILE__, __LINE__, simt::K_SHFL_DOWN, __VA_ARGS__)
#define __shfl_xor_sync(...)  simt::shfl (__FILE__, __LINE__, simt::K_SHFL_XOR, __VA_ARGS__)
#define __ballot_sync(m, p)   ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_BALLOT, (m), (p) ? 1 : 0, 0, 32))
#define __any_sync(m, p)      (simt::collective (__FILE__, __LINE__, simt::K_BALLOT, (m), (p) ? 1 : 0, 0, 32) != 0)
#define __all_sync(m, p)      (simt::collective (__FILE__, __LINE__, simt::K_BALLOT, (m), (p) ? 0 : 1, 0, 32) == 0)
#define __reduce_add_sync(m, v) ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_REDUCE_ADD, (m), (uint32_t)(v), 0, 32))
#define __reduce_or_sync(m, v)  ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_REDUCE_OR, (m), (uint32_t)(v), 0, 32))
#define __reduce_and_sync(m, v) ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_REDUCE_AND, (m), (uint32_t)(v), 0, 32))
#define __reduce_min_sync(m, v) ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_REDUCE_MIN, (m), (uint32_t)(v), 0, 32))
#define __reduce_max_sync(m, v) ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_REDUCE_MAX, (m), (uint32_t)(v), 0, 32))
#define __match_any_sync(m, v) ((uint32_t)simt::collective (__FILE__, __LINE__, simt::K_MATCH_ANY, (m), simt::to_bits (v), 0, 32))
#define __syncwarp(...)       ((void)simt::collective (__FILE__, __LINE__, simt::K_SYNCWARP, simt::mask_or_full (__VA_ARGS__), 0, 0, 32))
#define __syncthreads()       simt::syncthreads (__FILE__, __LINE__)
#define __syncthreads_or(p)   simt::syncthreads_or (__FILE__, __LINE__, (p) ? 1 : 0)
namespace simt { inline uint32_t mask_or_full () { return 0xffffffffu; } inline uint32_t mask_or_full (uint32_t m) { return m; } }

// ---------------------------------------------------------------- integer / float intrinsics
template <class T> static inline T __ldg (const T *p) { return *p; }
static inline int      __popc (uint32_t x) { return __builtin_popcount (x); }
static inline int      __popcll (uint64_t x) { return __builtin_popcountll (x); }
static inline int      __ffsll (long long x) { return __builtin_ffsll (x); }
static inline int      __clzll (long long x) { return x ? __builtin_clzll ((unsigned long long)x) : 64; }
static inline int      __ffs (int x) { return __builtin_ffs (x); }
static inline int      __clz (int x) { return x ? __builtin_clz ((uint32_t)x) : 32; }
static inline uint32_t __brev (uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline uint32_t __umulhi (uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi (uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __funnelshift_r (uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline uint32_t __funnelshift_rc (uint32_t lo, uint32_t hi, uint32_t s) { s = s > 32 ? 32 : s; return (uint32_t)((((uint64_t)hi << 32) | lo) >> s); }
static inline uint32_t __funnelshift_l (uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline uint32_t __byte_perm (uint32_t x, uint32_t y, uint32_t s)        // PRMT, default mode
{
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t b = (uint32_t)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;                                // replicate the sign
        r |= b << (8 * i);
    }
    return r;
}
static inline uint32_t __vcmpeq4 (uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline uint32_t __vsub4 (uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) r |= ((((a >> (8 * i)) & 0xff) - ((b >> (8 * i)) & 0xff)) & 0xff) << (8 * i); return r; }
static inline uint32_t __vcmpne4 (uint32_t a, uint32_t b) { return ~__vcmpeq4 (a, b); }
static inline uint32_t __vminu4 (uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) { uint32_t x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; r |= (x < y ? x : y) << (8 * i); } return r; }
static inline uint32_t __vmaxu4 (uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) { uint32_t x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; r |= (x > y ? x : y) << (8 * i); } return r; }
static inline uint32_t __vcmpgtu4 (uint32_t a, uint32_t b) { uint32_t r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) > ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline float    __uint_as_float (uint32_t u) { float f; memcpy (&f, &u, 4); return f; }
static inline uint32_t __float_as_uint (float f) { uint32_t u; memcpy (&u, &f, 4); return u; }
static inline long long __double_as_longlong (double d) { long long l; memcpy (&l, &d, 8); return l; }
static inline double   __longlong_as_double (long long l) { double d; memcpy (&d, &l, 8); return d; }
static inline float    __uint2float_rz (uint32_t u) { float f = (float)u; if ((double)f > (double)u) f = nextafterf (f, 0.0f); return f; }
static inline float    __uint2float_ru (uint32_t u) { float f = (float)u; if ((double)f < (double)u) f = nextafterf (f, INFINITY); return f; }
static inline float    __uint2float_rd (uint32_t u) { return __uint2float_rz (u); }
static inline float    __uint2float_rn (uint32_t u) { return (float)u; }
static inline uint32_t __float2uint_rz (float f) { return f <= 0 ? 0u : f >= 4294967296.0f ? 0xffffffffu : (uint32_t)f; }
static inline float    __fmul_rz (float a, float b) { const double d = (double)a * (double)b; float f = (float)d; if (fabs ((double)f) > fabs (d)) f = nextafterf (f, 0.0f); return f; }
static inline float    __fma_rn (float a, float b, float c) { return fmaf (a, b, c); }
static inline double   __ddiv_rn (double a, double b) { return a / b; }
static inline double   __dmul_rn (double a, double b) { volatile double r = a * b; return r; }   // (volatile: never contracted into an FMA)
static inline double   __dadd_rn (double a, double b) { volatile double r = a + b; return r; }
static inline double   __ll2double_rn (long long l) { return (double)l; }
static inline int      __double2int_rz (double d) { return (int)d; }
namespace simt { static inline float rcp_approx (float x) { return 1.0f / x; } }   // rcp.approx.ftz.f32 is within 1 ulp of this; users must not depend on more

// CUDA's overloads of min / max for mixed integer types
static inline uint32_t min (uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max (uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int      min (int a, int b) { return a < b ? a : b; }
static inline int      max (int a, int b) { return a > b ? a : b; }
static inline uint32_t min (uint32_t a, int b) { return min (a, (uint32_t)b); }
static inline uint32_t min (int a, uint32_t b) { return min ((uint32_t)a, b); }
static inline uint32_t max (uint32_t a, int b) { return max (a, (uint32_t)b); }
static inline uint32_t max (int a, uint32_t b) { return max ((uint32_t)a, b); }
static inline uint64_t min (uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t max (uint64_t a, uint64_t b) { return a > b ? a : b; }
static inline unsigned long long min (unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max (unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline long long min (long long a, long long b) { return a < b ? a : b; }
static inline long long max (long long a, long long b) { return a > b ? a : b; }
static inline float    min (float a, float b) { return fminf (a, b); }
static inline float    max (float a, float b) { return fmaxf (a, b); }
static inline double   min (double a, double b) { return fmin (a, b); }
static inline double   max (double a, double b) { return fmax (a, b); }

// ---------------------------------------------------------------- atomics (blocks of one launch may run on several OS threads)
template <class T> static inline T atomicAdd (T *p, T v) { return __atomic_fetch_add (p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicSub (T *p, T v) { return __atomic_fetch_sub (p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicOr (T *p, T v) { return __atomic_fetch_or (p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd (T *p, T v) { return __atomic_fetch_and (p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch (T *p, T v) { return __atomic_exchange_n (p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicMax (T *p, T v) { T o = *p; while (o < v && !__atomic_compare_exchange_n (p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <class T> static inline T atomicMin (T *p, T v) { T o = *p; while (o > v && !__atomic_compare_exchange_n (p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <class T> static inline T atomicCAS (T *p, T c, T v) { __atomic_compare_exchange_n (p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return c; }
static inline uint32_t atomicOr (uint32_t *p, int v) { return atomicOr (p, (uint32_t)v); }
static inline int      atomicOr (int *p, unsigned v) { return atomicOr (p, (int)v); }
static inline void __threadfence () { __atomic_thread_fence (__ATOMIC_SEQ_CST); }
static inline void __threadfence_block () {}

// ---------------------------------------------------------------- the runtime calls the library makes
typedef int cudaError_t;
typedef struct simt_stream_s { int id; } *cudaStream_t;
typedef struct simt_event_s { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorInvalidDeviceFunction = 98, cudaErrorNoKernelImageForDevice = 209 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaStreamDefault = 0, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1, cudaEventDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem, sharedMemPerBlockOptin; };

namespace simt { double now_ms (); void *dev_alloc (size_t n); void dev_free (void *p); }
static inline cudaError_t cudaGetDeviceCount (int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice (int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice (int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties (cudaDeviceProp *p, int)
{ memset (p, 0, sizeof *p); strcpy (p->name, "SIMT emulator (host)"); p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = 8ull << 30; p->sharedMemPerBlockOptin = 227 * 1024; return cudaSuccess; }
static inline cudaError_t cudaMalloc (void **p, size_t n) { *p = simt::dev_alloc (n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc (T **p, size_t n) { return cudaMalloc ((void **)p, n); }
static inline cudaError_t cudaMallocHost (void **p, size_t n) { return cudaMalloc (p, n); }
template <class T> static inline cudaError_t cudaMallocHost (T **p, size_t n) { return cudaMalloc ((void **)p, n); }
static inline cudaError_t cudaFree (void *p) { simt::dev_free (p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost (void *p) { simt::dev_free (p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync (void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { if (n) memmove (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy (void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove (d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync (void *d, int v, size_t n, cudaStream_t = 0) { if (n) memset (d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset (void *d, int v, size_t n) { if (n) memset (d, v, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol (T &sym, const void *s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) { memcpy ((char *)&sym + off, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate (cudaStream_t *s) { *s = new simt_stream_s{ 1 }; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags (cudaStream_t *s, unsigned) { return cudaStreamCreate (s); }
static inline cudaError_t cudaStreamDestroy (cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize () { return cudaSuccess; }
static inline cudaError_t cudaEventCreate (cudaEvent_t *e) { *e = new simt_event_s{ 0 }; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags (cudaEvent_t *e, unsigned) { return cudaEventCreate (e); }
static inline cudaError_t cudaEventDestroy (cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord (cudaEvent_t e, cudaStream_t = 0) { e->t = simt::now_ms (); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime (float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError () { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError () { return cudaSuccess; }
static inline const char *cudaGetErrorString (cudaError_t e) { return e ? "simt: error" : "no error"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute (F, cudaFuncAttribute, int) { return cudaSuccess; }
