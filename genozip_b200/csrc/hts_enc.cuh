// hts_enc.cuh — host-visible plan structures shared between api.cu and the hts kernels
#pragma once
#include "gzb_internal.cuh"

namespace gzb {

struct EncPlanDev {              // device pointers of one encode batch
    EncSection    *sections;   uint32_t n_sections;
    EncLeaf       *leaves;     uint32_t n_leaves;
    EncLeafDyn    *dyn;
    Tile          *tiles;      uint32_t n_tiles;          // TILE-sized pieces of every leaf input
    Tile          *stripe_tiles; uint32_t n_stripe_tiles; // TILE-sized pieces of every STRIPE section (Tile.leaf = section)
    uint32_t      *rans_list;  uint32_t n_rans;           // rANS leaves: requested order 1 first, then order 0; longest first within each
    uint2         *rans_jobs;  uint32_t n_rans_jobs;      // warp jobs: (first index into rans_list, count <= 8)
    uint32_t      *arith_list; uint32_t n_arith;          // arithmetic leaves, longest first
    uint32_t       n_arith_big, split_min;                // the first n_arith_big of them are at least split_min long: candidates of the split encoder
    SectionResult *results;
    CopySeg       *segs;
    uint8_t       *stripe_hdr;
    uint8_t       *pack_arena;                            // packed output (gzb_compress_sections_packed): sections are appended to this
    unsigned long long pack_cap, *pack_off;               //   buffer, 16-byte aligned; pack_off[n_sections] = offset of each, [n_sections] = total
    Arena          arena;
    int            rans_gpw, arith_lpw, copy_parts;
    bool           any_pack, any_o1;
    cudaEvent_t    ev_chain0, ev_chain1, ev_chain2, ev_arith0, ev_o0, ev_split;   // rANS kernel: chain0..chain1 on the main stream; arithmetic: arith0..chain2 on st2; order-0 arithmetic: .. ev_o0 on st3; split encoder: .. ev_split on st4
    cudaStream_t   st2, st3, st4;
    cudaEvent_t    ev_prof[2];                            // after the split encoder's bucket / model kernels (their durations, for the bench's kernel table)
    uint32_t      *queue;                                 // zeroed work counters of the persistent chain kernels (Q_* below)
    int            sm_count;
    uint64_t       launches;
};

struct DecPlanDev {
    DecSection    *sections;   uint32_t n_sections;
    DecLeaf       *leaves;                                // 4 per section
    uint32_t      *rans_list;  uint32_t n_rans;           // leaf slots of rANS sections, largest section first
    uint2         *rans_jobs;  uint32_t n_rans_jobs;      // warp jobs: (first index into rans_list, count <= 8)
    uint32_t      *arith_list; uint32_t n_arith;
    SectionResult *results;
    Arena          arena;
    int            rans_gpw, arith_lpw, parts;
    cudaEvent_t    ev_chain0, ev_chain1, ev_chain2, ev_arith0, ev_o0, ev_long;
    cudaStream_t   st2, st3, st4;
    uint32_t       n_long_cand, long_min;                 // the first n_long_cand entries of arith_list belong to sections of at least long_min bytes: candidates of k_arith_decode_long
    uint32_t      *queue;
    int            sm_count;
    uint64_t       launches;
};

// The chain kernels are PERSISTENT: a fixed number of CTAs per SM, every warp takes the next leaf of the (longest-first) list from a
// counter until the list is empty.  What is resident on an SM is then a choice, not the block scheduler's "as many as fit": the
// long leaves, taken first, share an issue port with a few other warps instead of fifteen (a chain wants a slot every ~6 cycles).
enum { Q_ARITH = 0, Q_ARITH_O0 = 1, Q_SPLIT_MODEL = 2, Q_ARITH_LONG = 3, Q_WORDS = 64 };
struct ChainTune { int arith_ctas, arith_o0_ctas, run4, split_stream, long_ent; uint32_t long_min; };      // CTAs per SM of the two arithmetic kernels; GZB_AR_* (see chain_tune)
const ChainTune &chain_tune ();
__device__ __forceinline__ uint32_t queue_take (uint32_t *counter, int lane)
{
    uint32_t s = 0;
    if (lane == 0) s = atomicAdd (counter, 1u);
    return __shfl_sync (0xffffffffu, s, 0);
}

void upload_log_tables (const double *l10, const double *l12);
void enc_run (EncPlanDev &P, cudaStream_t st);
void dec_run (DecPlanDev &P, cudaStream_t st);
void launch_rans_encode (EncPlanDev &P, cudaStream_t st);
void launch_rans_decode (DecPlanDev &P, cudaStream_t st);
void launch_arith_encode (EncPlanDev &P, cudaStream_t st);
void launch_arith_decode (DecPlanDev &P, cudaStream_t st);
void launch_arith_encode_o0 (EncPlanDev &P, cudaStream_t st);
void launch_arith_encode_split (EncPlanDev &P, cudaStream_t st);
void launch_arith_decode_o0 (DecPlanDev &P, cudaStream_t st);
void launch_arith_decode_long (DecPlanDev &P, cudaStream_t st);

} // namespace gzb
