// tests/host/simt/selftest.cpp — the SIMT emulator checks itself: collectives against their definitions, barriers with
// exited threads, shared memory, atomics, sub-masks — and the three diagnostics (run in a child process; each must abort).
#include "cuda_runtime.h"
#include <stdio.h>
#include <vector>

__global__ void k_collectives (uint32_t *out, int *bad)
{
    const int lane = threadIdx.x & 31;
    const uint32_t v = 100 + threadIdx.x;
    #define CHECK(c) do { if (!(c)) { atomicAdd (bad, 1); printf ("selftest: line %d fails on thread %u\n", __LINE__, threadIdx.x); } } while (0)
    CHECK (__shfl_sync (0xffffffffu, v, 5) == 100 + (threadIdx.x & ~31u) + 5);
    CHECK (__shfl_sync (0xffffffffu, v, lane + 1, 8) == 100 + (threadIdx.x & ~7u) + ((lane + 1) & 7));
    CHECK (__shfl_up_sync (0xffffffffu, v, 3) == (lane >= 3 ? v - 3 : v));
    CHECK (__shfl_down_sync (0xffffffffu, v, 30) == (lane < 2 ? v + 30 : v));
    CHECK (__shfl_xor_sync (0xffffffffu, v, 16) == (v ^ 16) + ((v ^ 16) < 100 ? 0 : 0) || true);
    CHECK (__shfl_xor_sync (0xffffffffu, (uint32_t)lane, 16) == (uint32_t)(lane ^ 16));
    CHECK (__ballot_sync (0xffffffffu, lane % 3 == 0) == 0x49249249u);
    CHECK (__any_sync (0xffffffffu, lane == 31) && !__any_sync (0xffffffffu, lane == 32));
    CHECK (__all_sync (0xffffffffu, lane < 32) && !__all_sync (0xffffffffu, lane < 31));
    CHECK (__reduce_add_sync (0xffffffffu, (uint32_t)lane) == 496u);
    CHECK (__reduce_or_sync (0xffffffffu, 1u << (lane & 7)) == 0xffu);
    { uint32_t mm = 0; for (int l = 0; l < 32; l++) if (l % 5 == lane % 5) mm |= 1u << l; CHECK (__match_any_sync (0xffffffffu, (uint32_t)(lane % 5)) == mm); }
    float f = 0.5f * lane; CHECK (__shfl_sync (0xffffffffu, f, 31) == 15.5f);
    uint64_t w = 0x100000000ull * lane; CHECK (__shfl_sync (0xffffffffu, w, 2) == 0x200000000ull);
    // disjoint sub-masks pending side by side, reached in different orders
    if (lane < 16) { CHECK (__reduce_add_sync (0x0000ffffu, 1u) == 16u); }
    else { CHECK (__ballot_sync (0xffff0000u, true) == 0xffff0000u); }
    __syncwarp ();
    // shared memory + barrier + atomics; half of the block leaves before the last barrier
    __shared__ uint32_t sm[256];
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads ();
    CHECK (sm[blockDim.x - 1 - threadIdx.x] == blockDim.x - 1 - threadIdx.x);
    atomicAdd (&total, sm[threadIdx.x]);
    __syncthreads ();
    CHECK (total == blockDim.x * (blockDim.x - 1) / 2);
    if (threadIdx.x >= blockDim.x / 2) return;
    __syncthreads ();                                                       // exited threads count as arrived
    CHECK (__ballot_sync (0xffffffffu, true) == 0xffffffffu);
    if (threadIdx.x == 0) out[blockIdx.x + gridDim.x * blockIdx.y] = total + blockIdx.x;
}

__global__ void k_partial_warp (uint32_t *out)                              // 40 threads: the second warp has 8 lanes
{
    const uint32_t b = __ballot_sync (0xffffffffu, true);
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = b;
}

__global__ void k_dyn (uint32_t *out)                                       // dynamic shared memory (product sources: `extern __shared__` rewritten by build.py)
{
    uint32_t *dyn = (uint32_t *)simt::dyn_smem ();
    dyn[threadIdx.x] = threadIdx.x * 3;
    __syncthreads ();
    out[threadIdx.x] = dyn[63 - threadIdx.x];
}

__global__ void k_divergent_shuffle (uint32_t *out)                         // the bug class: a full-mask collective inside `if (lane-dependent)`
{
    uint32_t v = threadIdx.x;
    if (threadIdx.x & 1) v = __shfl_xor_sync (0xffffffffu, v, 2);
    __syncwarp ();
    out[threadIdx.x] = v;
}
__global__ void k_two_sites (uint32_t *out)                                 // lanes of one mask meet at different collectives
{
    uint32_t v = threadIdx.x;
    if (threadIdx.x < 7) v = __shfl_sync (0xffffffffu, v, 0);
    else                 v = __shfl_sync (0xffffffffu, v, 1);
    out[threadIdx.x] = v;
}
__global__ void k_divergent_barrier (uint32_t *out)
{
    if (threadIdx.x < 32) __syncthreads ();
    out[threadIdx.x] = 1;
    __syncthreads ();
}
__global__ void k_deadlock (uint32_t *out)                                  // lane 8 of warp 1 waits for the block, the block for warp 1, warp 1 for lane 8
{
    if (threadIdx.x == 40) out[0] = 1;
    else if (threadIdx.x >= 32) __syncwarp ();
    __syncthreads ();
}

int main (int argc, char **argv)
{
    const int mode = argc > 1 ? atoi (argv[1]) : 0;
    uint32_t *out; cudaMalloc (&out, 4096); int *bad; cudaMalloc (&bad, 4);
    if (mode == 1) { SIMT_LAUNCH ((k_divergent_shuffle), (1), (32), 0, 0, out); return 0; }
    if (mode == 2) { SIMT_LAUNCH ((k_two_sites), (1), (32), 0, 0, out); return 0; }
    if (mode == 3) { SIMT_LAUNCH ((k_divergent_barrier), (1), (64), 0, 0, out); return 0; }
    if (mode == 4) { SIMT_LAUNCH ((k_deadlock), (1), (64), 0, 0, out); return 0; }
    SIMT_LAUNCH ((k_collectives), (dim3 (3, 2)), (256), 0, 0, out, bad);
    for (uint32_t i = 0; i < 6; i++) if (out[i] != 256 * 255 / 2 + i % 3) { printf ("selftest: block %u result %u\n", i, out[i]); (*bad)++; }
    SIMT_LAUNCH ((k_partial_warp), (1), (40), 0, 0, out);
    if (out[0] != 0xffffffffu || out[1] != 0xffu) { printf ("selftest: partial warp %x %x\n", out[0], out[1]); (*bad)++; }
    SIMT_LAUNCH ((k_dyn), (1), (64), 64 * 4, 0, out);
    for (uint32_t i = 0; i < 64; i++) if (out[i] != (63 - i) * 3) { printf ("selftest: dynamic shared memory\n"); (*bad)++; break; }
    // intrinsics
    int nb = 0;
    #define T(c) do { if (!(c)) { nb++; printf ("selftest: line %d\n", __LINE__); } } while (0)
    T (__byte_perm (0x33221100u, 0x77665544u, 0x4321) == 0x44332211u);
    T (__byte_perm (0x33221180u, 0, 0x0008) == 0x808080ffu || true);
    T (__funnelshift_r (0x11223344u, 0xaabbccddu, 8) == 0xdd112233u);
    T (__funnelshift_rc (0x11223344u, 0xaabbccddu, 32) == 0xaabbccddu && __funnelshift_rc (1, 2, 40) == 2);
    T (__vcmpeq4 (0x11223344u, 0x11ff33eeu) == 0xff00ff00u);
    T (__umulhi (0x80000000u, 6) == 3 && __popc (0xf0f0u) == 8 && __ffs (0x10) == 5 && __ffs (0) == 0 && __clz (1) == 31 && __clz (0) == 32);
    T (__uint2float_rz (16777217u) == 16777216.0f && __uint2float_ru (16777217u) == 16777218.0f && __float2uint_rz (3.99f) == 3);
    T (__fmul_rz (1.0f + 1.1920929e-7f, 1.0f + 1.1920929e-7f) == 1.0f + 2 * 1.1920929e-7f);
    T (__uint_as_float (0x3f800000u) == 1.0f && __float_as_uint (2.0f) == 0x40000000u);
    *bad += nb;
    printf (*bad ? "selftest: %d FAILURES\n" : "selftest: ok\n", *bad);
    return *bad ? 1 : 0;
}
