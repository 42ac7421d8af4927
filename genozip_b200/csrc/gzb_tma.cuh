// gzb_tma.cuh — 1-D bulk asynchronous copies (TMA, cp.async.bulk → UBLKCP in SASS) and the mbarriers that signal them.
//
// Used by the row-serial kernels (PBWT, DOMQ staging): a ring of shared-memory slots is filled global → shared by the copy engine
// while the CTA works on the slot before, so the dependent chain of a row never waits for HBM.
//
// A bulk copy needs 16-byte aligned addresses and a size that is a multiple of 16.  A byte range [g, g+n) of arbitrary alignment
// is brought in as its aligned superset, and lands in the slot at the SAME misalignment (slot + (g & 15)), so slot sizes are
// tma_slot_bytes (n).  The superset must lie inside the caller's buffer: tma_superset_ok () says so; rows at the very edges of a
// buffer that fail it are read with ordinary loads.
//
// On the CPU test suite's SIMT emulator (tests/host/simt) a bulk copy is a memcpy at issue time and a wait returns at once.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gzb {

__host__ __device__ __forceinline__ uint32_t tma_slot_bytes (uint32_t n) { return ((n + 15u + 15u) & ~15u) + 16u; }

__device__ __forceinline__ bool tma_superset_ok (const uint8_t *g, uint32_t n, const uint8_t *buf, uint64_t buf_len)
{
    const uintptr_t a = (uintptr_t)g & ~(uintptr_t)15, b = ((uintptr_t)g + n + 15) & ~(uintptr_t)15;
    return a >= (uintptr_t)buf && b <= (uintptr_t)buf + buf_len;
}

#ifndef GZB_SIMT_EMULATION

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t)__cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init ()       // make the initialised barriers visible to the async proxy
{
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
    asm volatile ("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx (uint64_t *bar, uint32_t bytes)
{
    asm volatile ("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity)
{
    asm volatile (
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
// global → shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s (void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(smem_u32 (smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
// shared → global (bulk group); the shared source was written by ordinary stores: fence them to the async proxy first
__device__ __forceinline__ void bulk_s2g (void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile ("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32 (smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_fence_smem_writes () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit ()            { asm volatile ("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read () { asm volatile ("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all ()  { asm volatile ("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

// a hint: bring the line of `p` into L2 (the walker of a latency-bound chain asks for the lines it may need two steps ahead)
__device__ __forceinline__ void prefetch_l2 (const void *p) { asm volatile ("prefetch.global.L2 [%0];" :: "l"(p)); }

#else   // ------------------------------------------------------------------ the emulator: synchronous stand-ins

__device__ __forceinline__ void prefetch_l2 (const void *) {}

__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init () {}
__device__ __forceinline__ void mbar_arrive (uint64_t *) {}
__device__ __forceinline__ void mbar_arrive_expect_tx (uint64_t *, uint32_t) {}
__device__ __forceinline__ void mbar_wait (uint64_t *, uint32_t) {}
__device__ __forceinline__ void bulk_g2s (void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *) { memcpy (smem_dst, gsrc, bytes); }
__device__ __forceinline__ void bulk_s2g (void *gdst, const void *smem_src, uint32_t bytes) { memcpy (gdst, smem_src, bytes); }
__device__ __forceinline__ void bulk_fence_smem_writes () {}
__device__ __forceinline__ void bulk_commit () {}
template <int N> __device__ __forceinline__ void bulk_wait_read () {}
template <int N> __device__ __forceinline__ void bulk_wait_all () {}

#endif

// The emulator has no asynchrony: a copy is complete when its issuing thread has run, and a wait cannot block.  Where the
// issuing thread and the first waiters are not already ordered by a barrier, the kernel places this marker after the issue.
#ifdef GZB_SIMT_EMULATION
  #define GZB_TMA_EMU_ISSUED() __syncthreads ()
#else
  #define GZB_TMA_EMU_ISSUED()
#endif

// One row of a ring: issue (one thread).  `g` may have any alignment; the data lands at slot + (g & 15).
__device__ __forceinline__ void tma_row_issue (uint8_t *slot, const uint8_t *g, uint32_t n, uint64_t *bar)
{
    const uint32_t mis = (uint32_t)((uintptr_t)g & 15), bytes = (mis + n + 15u) & ~15u;
    mbar_arrive_expect_tx (bar, bytes);
    bulk_g2s (slot, g - mis, bytes, bar);
}

} // namespace gzb
