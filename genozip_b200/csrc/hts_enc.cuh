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
    cudaEvent_t    ev_chain0, ev_chain1, ev_chain2, ev_arith0, ev_o0;   // rANS kernel: chain0..chain1 on the main stream; arithmetic: arith0..chain2 on st2; order-0 arithmetic: .. ev_o0 on st3
    cudaStream_t   st2, st3;
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
    cudaEvent_t    ev_chain0, ev_chain1, ev_chain2, ev_arith0, ev_o0;
    cudaStream_t   st2, st3;
    uint64_t       launches;
};

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

} // namespace gzb
