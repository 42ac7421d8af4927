// gzb_internal.cuh — shared device-side definitions of libgzb200 (sm_100a).
//
// Vocabulary (follows the reference's domain): a *section* is one call of a simple codec on one Context
// stream (b250 or local); a *leaf* is one entropy-coded block inside a section's container — the whole
// section for RANB/RANb/ARTB/ARTb, or one candidate method of one byte-plane for the STRIPE codecs
// (RANW/RANw/ARTW/ARTw; reference rANS_static4x16pr.c:1165-1227, arith_dynamic.c:636-768).
// The bitstream fixes 4 dependency chains per rANS leaf and 1 per arithmetic leaf, so
// throughput comes from running every leaf of every section of every VBlock of a batch concurrently.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

// Where the lanes of a warp all read a location that one of them (or all, redundantly) stores to a few instructions later,
// every lane must have read before any lane stores: GZB_WARP_READS_DONE () — a __syncwarp () — stands after every such group
// of reads.  (Round 1 compiled it to nothing on the GPU, relying on lock-step issue of a converged warp; the real barrier
// measured within noise on B200, so it is now always there.  -DGZB_READS_DONE_LOCKSTEP restores the old build for an A/B run.)
#if defined(GZB_READS_DONE_LOCKSTEP) && !defined(GZB_SIMT_EMULATION)
  #define GZB_WARP_READS_DONE()
#else
  #define GZB_WARP_READS_DONE() __syncwarp ()
#endif

namespace gzb {

// container flag bits (reference htscodecs/rANS_static4x16.h:38-44, arith_dynamic.h:41-48)
enum : uint32_t { F_ORDER = 1, F_EXT = 4, F_STRIPE = 8, F_NOSZ = 16, F_CAT = 32, F_RLE = 64, F_PACK = 128 };
constexpr uint32_t RANS_L = 1u << 15;          // rANS_word.h:58

enum : uint8_t { CODER_RANS = 0, CODER_ARITH = 1 };

constexpr uint32_t TILE = 32768;               // bytes of leaf input per CTA in the bandwidth-shaped passes

// rANS encoder symbol (reference RansEncSymbol, rANS_word.h:169-180) packed to one 16-byte load
struct __align__(16) EncSym {
    uint32_t x_max;      // renormalise when x >= x_max
    uint32_t rcp;        // fixed-point reciprocal of freq
    uint32_t bias;
    uint32_t cmpl_sh;    // low 16: (1<<bits)-freq ; high 16: reciprocal shift (0..11)
};

// ---------------------------------------------------------------- encode side
// Device-side bump allocator over a host-sized arena: tables whose size depends on the alphabet actually
// present (known only on the device) are carved here.  Exhaustion sets *overflow; the host grows the arena
// and replays the batch (api.cu).
struct Arena {
    uint8_t *base; unsigned long long cap; unsigned long long *cursor; int *overflow;
    __device__ uint8_t *alloc (unsigned long long bytes) const {
        bytes = (bytes + 15ull) & ~15ull;
        unsigned long long off = atomicAdd (cursor, bytes);
        if (off + bytes > cap) { *overflow = 1; return nullptr; }
        return base + off;
    }
};

struct EncLeaf {                 // host-planned, read-only on the device
    const uint8_t *in;           // leaf input (section input, or a byte-plane of it)
    uint8_t  *packbuf;           // PACK output (n+1 bytes) or nullptr
    uint32_t *hist0;             // 256 counters
    uint8_t  *outbuf;            // leaf scratch: frequency table at the front, payload written backwards from the end
    uint32_t  out_cap;
    uint32_t  n;                 // leaf input length
    uint32_t  section;           // owning section
    uint8_t   coder;             // CODER_RANS / CODER_ARITH
    uint8_t   order_req;         // container flags requested for this leaf (order bit, PACK, RLE, NOSZ)
    uint8_t   pad[2];
};
constexpr uint32_t CTXB = 520;   // max encoded bytes of one context's frequencies: 256 x 2-byte varints + slack

struct EncLeafDyn {              // device-written state of a leaf
    const uint8_t *eff_in;       // what the entropy coder consumes (input or packbuf)
    uint32_t *hist1;             // rANS O1: counters [nsym][nsym] by symbol rank (arena)
    EncSym   *symtab;            // rANS: [256] by symbol (O0) or [nsym][nsym] by rank (O1) (arena)
    uint8_t  *ctxbytes;          // rANS O1: nsym x CTXB temp for per-context encoded frequencies (arena)
    uint32_t *models;            // arith: adaptive model memory (arena)
    uint32_t *split_pos;         // arith, split encoder (arith_split.cu): positions grouped by context, in order (arena); nullptr = not split
    uint32_t *split_start;       //   257 offsets into split_pos
    uint2    *split_rec;         //   per symbol: cumFreq | freq << 16, totFreq
    uint32_t eff_n;
    uint32_t hdr_len;            // container header: flags byte, [varint n], [pack meta, varint packed_len]
    uint32_t tab_len;            // bytes of frequency table at outbuf[0..)
    uint32_t payload_len;        // bytes at outbuf[out_cap - payload_len ..)
    uint32_t total_len;          // final container length (hdr + body or hdr + raw copy)
    uint8_t  eff_order;          // 0/1 after the "<8 symbols" rule
    uint8_t  packed;             // 1 = packbuf holds the coder input
    uint8_t  cat;                // 1 = body >= input: stored raw (X_CAT)
    uint8_t  per_byte;           // PACK: symbols per byte (8,4,2, 0 = constant)
    uint16_t nsym;               // distinct symbols in eff_in (O1: including the forced symbol 0)
    uint8_t  shift;              // O1 table bits (10/12)
    uint8_t  pad;
    uint8_t  hdr[272];
    uint8_t  rank[256];          // symbol -> compact index
    uint8_t  code[256];          // PACK: symbol -> code
};

struct EncSection {              // host-planned
    const uint8_t *in;
    uint8_t  *out;
    uint8_t  *planes;            // STRIPE: transposed copy of the input (n bytes) or nullptr
    uint32_t  n;
    uint32_t  out_cap;
    uint32_t  first_leaf;
    uint32_t  n_leaves;
    uint8_t   coder;
    uint8_t   order;             // genozip order byte: 0x01 / 0x19 / 0x81 / 0x99 (codec_htscodecs.c:17-20)
    uint8_t   stripe;            // 1 = STRIPE container (n > 20)
    uint8_t   soft_fail;         // capacity too small → status GZB_SOFT_FAIL, no output
};

struct SectionResult { uint32_t out_len; int32_t status; };

struct Tile { uint32_t leaf; uint32_t off; };               // TILE bytes of a leaf's input starting at off
struct CopySeg { const uint8_t *src; uint8_t *dst; uint32_t len; uint32_t pad; };

// ---------------------------------------------------------------- decode side
struct DecLeaf {                 // device-written by the parse kernel (4 slots per section)
    const uint8_t *body;         // entropy-coded (or raw) body
    uint8_t  *dst;               // where the coder writes: final output, plane buffer or unpack temp
    uint8_t  *fin;               // where unpack writes (== dst when no PACK)
    uint32_t  body_len;
    uint32_t  body_ulen;         // symbols the coder produces
    uint32_t  ulen;              // bytes after unpack
    uint8_t   valid;             // slot in use
    uint8_t   coder;
    uint8_t   order;
    uint8_t   cat;
    uint8_t   rle;
    uint8_t   pack;              // 1 = PACK present
    uint8_t   per_byte;          // PACK symbols per byte (1 = stored verbatim, 0 = constant)
    uint8_t   shift;             // O1 bits
    uint8_t   map[16];
    uint32_t  payload_off;       // rANS: offset of the 4 initial states inside body
    int32_t   err;
    uint2    *lut;               // rANS O0: slot -> { sym | freq<<16 , slot - start }  (arena)
    uint32_t *lut1;              // rANS O1: [ctx row][1<<shift] slot -> row(sym) | (freq-1)<<8 | (slot-start)<<20 (arena)
    uint32_t *models;            // arith model memory (arena)
    uint16_t  nsym;              // arith: max symbol + 1
    uint16_t  nctx;              // rANS O1: number of rows
    uint8_t   ctxrank[256];      // rANS O1: symbol -> row
    uint8_t   symof[256];        // rANS O1: row -> symbol
};

struct DecSection {              // host-planned
    const uint8_t *in;
    uint8_t  *out;
    uint8_t  *planes;            // n bytes: plane buffer for STRIPE
    uint8_t  *tmp;               // n bytes: PACK temp
    uint32_t  in_len;
    uint32_t  n;                 // expected uncompressed length
    uint8_t   coder;
    uint8_t   pad[3];
};

} // namespace gzb
