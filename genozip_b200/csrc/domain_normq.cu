// domain_normq.cu — NORMQ, the fallback quality codec of SAM / BAM (reference src/codec_normq.c), on sm_100a.
//
//   gather       codec_normq_compress before its sub-codec (:43-62): the quality strings of a VBlock copied into one buffer
//                (QUAL.local), a string reversed where its read is reverse-complemented (str_reverse, :55); a SAM line without
//                quality arrives as the one byte ' ' (sam_zip_qual).
//   reconstruct  codec_normq_reconstruct (:85-106) for every line of a VBlock at once: `len` bytes copied (reversed under
//                last_flags.rev_comp) from the stream's cursor — unless the byte at the cursor is ' ': then the line has no
//                quality, ONE byte is consumed and sam_reconstruct_missing_quality writes '*' (src/sam_qual.c:532).
//
// Both are bandwidth-shaped (N read + N written; one warp per line, the descriptors of 32 lines fetched together).  The one serial
// thing is where a line starts in the stream when lines before it had no quality: a quality character is never ' ' (Phred+33
// starts at '!'), so every ' ' in the stream is one such line, in order; the k-th ' ' at stream position S_k belongs to the line i
// with  F(i) = S_k + sum over the earlier missing lines of (len - 1),  F = prefix sums of the lines' lengths — a binary search per
// missing line, one thread per VBlock (files without missing qualities, the rule, never enter it).
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr int NQ_THREADS = 1024;

struct NqVb {
    const uint8_t *txt; const uint64_t *line_off; const uint32_t *line_len; const uint8_t *is_rev;
    uint8_t *local; uint8_t *out; uint8_t *missing;
    unsigned long long *pos;      // [n_lines + 1] position of every line in the stream (gather: = F; reconstruct: after the missing lines' correction)
    unsigned long long *F;        // [n_lines + 1] prefix sums of line_len
    uint32_t *marks;              // reconstruct: positions of the ' ' bytes of the stream, ascending (at most n_lines are looked at)
    uint32_t *info;               // [0] number of ' ' bytes, [1] error, [2..3] the 64-bit sum of the lines' lengths
    unsigned long long local_len, local_cap, out_cap;
    uint32_t n_lines, pad;
};

// F = exclusive prefix sums of the lines' lengths (F[n_lines] = total): one CTA per VBlock
__global__ void __launch_bounds__(NQ_THREADS) k_normq_prefix (const NqVb *vbs)
{
    const NqVb &V = vbs[blockIdx.x];
    __shared__ unsigned long long s_warp[32], s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads ();
    for (uint32_t base = 0; base < V.n_lines; base += NQ_THREADS) {
        const uint32_t i = base + tid;
        const unsigned long long v = i < V.n_lines ? V.line_len[i] : 0;
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads ();
        if (warp == 0) {
            const unsigned long long x = s_warp[lane]; unsigned long long xi = x;
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, xi, o); if (lane >= o) xi += t; }
            s_warp[lane] = xi - x;
        }
        __syncthreads ();
        const unsigned long long ex = s_carry + s_warp[warp] + inc - v;
        if (i < V.n_lines) { V.F[i] = ex; V.pos[i] = ex; }
        __syncthreads ();
        if (tid == NQ_THREADS - 1) s_carry = ex + v;
        __syncthreads ();
    }
    if (tid == 0) { V.F[V.n_lines] = s_carry; V.pos[V.n_lines] = s_carry; V.info[2] = (uint32_t)s_carry; V.info[3] = (uint32_t)(s_carry >> 32); }
}

// one warp per line, 32 lines' descriptors fetched together; DIR 0: text -> stream (gather), 1: stream -> text (reconstruct)
template <int DIR>
__global__ void __launch_bounds__(256) k_normq_lines (const NqVb *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const NqVb &V = vbs[blk_vb[blockIdx.x]];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = blk_first[blockIdx.x] + warp * 32, gl = first + lane;
    uint32_t my_len = 0, my_rev = 0, my_miss = 0; unsigned long long my_a = 0, my_b = 0;
    if (gl < V.n_lines) {
        my_len = V.line_len[gl]; my_rev = V.is_rev ? V.is_rev[gl] : 0;
        if (DIR == 0) { my_a = V.line_off[gl]; my_b = V.pos[gl]; }                          // a: in the text, b: in the stream
        else {
            my_a = V.F[gl]; my_b = V.pos[gl];                                               // a: in the output (len bytes per line), b: in the stream
            my_miss = my_len && my_b < V.local_len && V.local[my_b] == ' ';
            if (V.missing) V.missing[gl] = (uint8_t)my_miss;
        }
    }
    const uint32_t cnt = first < V.n_lines ? min (32u, V.n_lines - first) : 0;
    for (uint32_t t = 0; t < cnt; t++) {
        const uint32_t len = __shfl_sync (0xffffffffu, my_len, t);
        if (!len) continue;
        const uint32_t rev = __shfl_sync (0xffffffffu, my_rev, t), miss = __shfl_sync (0xffffffffu, my_miss, t);
        const unsigned long long a = __shfl_sync (0xffffffffu, my_a, t), b = __shfl_sync (0xffffffffu, my_b, t);
        if (DIR == 0) {
            const uint8_t *src = V.txt + a; uint8_t *dst = V.local + b;
            for (uint32_t i = lane; i < len; i += 32) dst[i] = src[rev ? len - 1 - i : i];
        }
        else {
            uint8_t *dst = V.out + a;
            if (miss) { if (lane == 0) dst[0] = '*'; continue; }                            // sam_reconstruct_missing_quality; the rest of the line's slot is undefined
            const uint8_t *src = V.local + b;
            for (uint32_t i = lane; i < len; i += 32) dst[i] = src[rev ? len - 1 - i : i];
        }
    }
}

// reconstruct, step 1: the ' ' bytes of the stream, ascending (block-wide compaction, 16 bytes per thread and round; only the
// first n_lines of them can be lines)
__global__ void __launch_bounds__(NQ_THREADS) k_normq_marks (const NqVb *vbs)
{
    const NqVb &V = vbs[blockIdx.x];
    __shared__ uint32_t s_warp[32], s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads ();
    for (unsigned long long base = 0; base < V.local_len; base += 16ull * NQ_THREADS) {
        const unsigned long long i0 = base + 16ull * tid;
        uint32_t mask = 0;                                                   // bit j: byte i0 + j is ' '
        for (int j = 0; j < 16; j++) if (i0 + j < V.local_len && V.local[i0 + j] == ' ') mask |= 1u << j;
        const uint32_t mine = __popc (mask);
        uint32_t inc = mine;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads ();
        if (warp == 0) {
            const uint32_t x = s_warp[lane]; uint32_t xi = x;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync (0xffffffffu, xi, o); if (lane >= o) xi += t; }
            s_warp[lane] = xi - x;
        }
        __syncthreads ();
        uint32_t at = s_carry + s_warp[warp] + inc - mine;
        for (uint32_t m = mask; m; m &= m - 1, at++) if (at < V.n_lines) V.marks[at] = (uint32_t)(i0 + (__ffs (m) - 1));
        __syncthreads ();
        if (tid == NQ_THREADS - 1) s_carry = s_carry + s_warp[warp] + inc;
        __syncthreads ();
    }
    if (tid == 0) V.info[0] = s_carry;
}

// reconstruct, step 2 (only VBlocks whose stream has ' ' bytes): which lines they are, then every line's position in the stream
__global__ void __launch_bounds__(NQ_THREADS) k_normq_resolve (const NqVb *vbs)
{
    const NqVb &V = vbs[blockIdx.x];
    const uint32_t n_marks = V.info[0];
    if (!n_marks) {                                                          // no missing quality: the stream is the concatenation of the lines
        if (threadIdx.x == 0 && V.F[V.n_lines] != V.local_len) V.info[1] = 1;
        return;
    }
    __shared__ int s_bad;
    if (threadIdx.x == 0) {
        s_bad = n_marks > V.n_lines;
        unsigned long long D = 0;                                            // bytes the missing lines so far did NOT take: sum of (len - 1)
        for (uint32_t k = 0; k < n_marks && !s_bad; k++) {
            const unsigned long long target = (unsigned long long)V.marks[k] + D;
            uint32_t lo = 0, hi = V.n_lines;                                 // the first line with F >= target (lines of length 0 share an F: take the one that has bytes)
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (V.F[mid] < target) lo = mid + 1; else hi = mid; }
            while (lo < V.n_lines && V.F[lo] == target && V.line_len[lo] == 0) lo++;
            if (lo >= V.n_lines || V.F[lo] != target) { s_bad = 1; break; }
            V.pos[lo] = ~0ull;                                               // marked; turned into positions below
            D += V.line_len[lo] - 1;
        }
    }
    __syncthreads ();
    if (s_bad) { if (threadIdx.x == 0) V.info[1] = 1; return; }
    // position of line i = F(i) - sum over the marked lines before it of (len - 1): a scan of the corrections, serial over 1024-line tiles
    __shared__ unsigned long long s_warp[32], s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads ();
    for (uint32_t base = 0; base < V.n_lines; base += NQ_THREADS) {
        const uint32_t i = base + tid;
        const bool marked = i < V.n_lines && V.pos[i] == ~0ull;
        const unsigned long long v = marked ? V.line_len[i] - 1 : 0;
        unsigned long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads ();
        if (warp == 0) {
            const unsigned long long x = s_warp[lane]; unsigned long long xi = x;
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync (0xffffffffu, xi, o); if (lane >= o) xi += t; }
            s_warp[lane] = xi - x;
        }
        __syncthreads ();
        const unsigned long long before = s_carry + s_warp[warp] + inc - v;
        if (i < V.n_lines) V.pos[i] = V.F[i] - before;
        __syncthreads ();
        if (tid == NQ_THREADS - 1) s_carry = before + v;
        __syncthreads ();
    }
    if (tid == 0 && V.F[V.n_lines] - s_carry != V.local_len) V.info[1] = 1;  // the stream must be used up exactly
}

struct Carver {
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

int normq_run (gzb_engine *e, gzb_normq_vb *vbs, uint32_t n_vbs, uint32_t flags, int dir)
{
    if (!e || (!vbs && n_vbs)) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<NqVb> h (n_vbs);
    std::vector<uint32_t> bvb, bfirst;
    std::vector<uint64_t> total (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_normq_vb &S = vbs[v]; S.status = GZB_OK;
        if ((S.n_lines && (!S.line_len || (dir == 0 && !S.line_off))) || (dir == 0 ? (!S.txt && S.txt_len) : (!S.local && S.local_len))) return GZB_E_BADARG;
        if (!devptr) for (uint32_t i = 0; i < S.n_lines; i++) total[v] += S.line_len[i];
        else total[v] = dir == 0 ? S.local_cap : S.out_cap;                  // (the lengths are on the device: the capacities bound the work)
        if (!devptr && (dir == 0 ? total[v] > S.local_cap : total[v] > S.out_cap)) { e->err = "NORMQ: output capacity too small"; return GZB_E_BADARG; }
        for (uint32_t f = 0; f < S.n_lines; f += 256) { bvb.push_back (v); bfirst.push_back (f); }
    }
    Carver c { nullptr, 0 };
    NqVb *d_vbs = nullptr; uint32_t *d_bvb = nullptr, *d_bfirst = nullptr, *d_info = nullptr;
    const size_t nb4 = bvb.size () * 4, nb4a = (nb4 + 255) & ~(size_t)255, desc_bytes = ((size_t)n_vbs * sizeof (NqVb) + 255) & ~(size_t)255;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<NqVb> (n_vbs); d_bvb = c.take<uint32_t> (bvb.size () + 1); d_bfirst = c.take<uint32_t> (bfirst.size () + 1);
        d_info = c.take<uint32_t> ((size_t)n_vbs * 4);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const gzb_normq_vb &S = vbs[v]; NqVb &D = h[v];
            D.n_lines = S.n_lines; D.local_cap = S.local_cap; D.out_cap = S.out_cap;
            D.local_len = dir == 0 ? 0 : S.local_len;
            D.line_len = devptr ? S.line_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.is_rev = !S.is_rev ? nullptr : devptr ? S.is_rev : c.take<uint8_t> ((size_t)S.n_lines + 1);
            D.pos = c.take<unsigned long long> ((size_t)S.n_lines + 1); D.F = c.take<unsigned long long> ((size_t)S.n_lines + 1);
            D.info = d_info ? d_info + 4 * (size_t)v : nullptr;
            if (dir == 0) {
                D.txt = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
                D.line_off = devptr ? S.line_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
                D.local = devptr ? (uint8_t *)S.local : c.take<uint8_t> (total[v] + 16);
                D.out = nullptr; D.missing = nullptr; D.marks = nullptr;
            }
            else {
                D.txt = nullptr; D.line_off = nullptr;
                D.local = devptr ? (uint8_t *)S.local : c.take<uint8_t> (S.local_len + 16);
                D.out = devptr ? (uint8_t *)S.out : c.take<uint8_t> (total[v] + 16);
                D.missing = !S.missing ? nullptr : devptr ? S.missing : c.take<uint8_t> ((size_t)S.n_lines + 1);
                D.marks = c.take<uint32_t> ((size_t)S.n_lines + 1);
            }
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, desc_bytes + 2 * nb4a + (size_t)n_vbs * 16 + 512); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs && !devptr; v++) {
        const gzb_normq_vb &S = vbs[v]; NqVb &D = h[v];
        if (S.n_lines) {
            CK (cudaMemcpyAsync ((void *)D.line_len, S.line_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            if (S.is_rev) CK (cudaMemcpyAsync ((void *)D.is_rev, S.is_rev, S.n_lines, cudaMemcpyHostToDevice, st));
            if (dir == 0) CK (cudaMemcpyAsync ((void *)D.line_off, S.line_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
        }
        if (dir == 0 && S.txt_len) CK (cudaMemcpyAsync ((void *)D.txt, S.txt, S.txt_len, cudaMemcpyHostToDevice, st));
        if (dir == 1 && S.local_len) CK (cudaMemcpyAsync (D.local, S.local, S.local_len, cudaMemcpyHostToDevice, st));
    }
    memcpy (e->pin, h.data (), (size_t)n_vbs * sizeof (NqVb));               // descriptors through the pinned staging (stage.cu says why)
    CK (cudaMemcpyAsync (d_vbs, e->pin, (size_t)n_vbs * sizeof (NqVb), cudaMemcpyHostToDevice, st));
    if (nb4) {
        memcpy (e->pin + desc_bytes, bvb.data (), nb4); memcpy (e->pin + desc_bytes + nb4a, bfirst.data (), nb4);
        CK (cudaMemcpyAsync (d_bvb, e->pin + desc_bytes, nb4, cudaMemcpyHostToDevice, st));
        CK (cudaMemcpyAsync (d_bfirst, e->pin + desc_bytes + nb4a, nb4, cudaMemcpyHostToDevice, st));
    }
    CK (cudaMemsetAsync (d_info, 0, (size_t)n_vbs * 16, st));
    k_normq_prefix<<<n_vbs, NQ_THREADS, 0, st>>>(d_vbs); e->launches++;
    if (dir == 1) {
        k_normq_marks<<<n_vbs, NQ_THREADS, 0, st>>>(d_vbs);
        k_normq_resolve<<<n_vbs, NQ_THREADS, 0, st>>>(d_vbs);
        e->launches += 2;
    }
    if (!bvb.empty ()) {
        if (dir == 0) k_normq_lines<0><<<(uint32_t)bvb.size (), 256, 0, st>>>(d_vbs, d_bvb, d_bfirst);
        else          k_normq_lines<1><<<(uint32_t)bvb.size (), 256, 0, st>>>(d_vbs, d_bvb, d_bfirst);
        e->launches++;
    }
    uint32_t *info = reinterpret_cast<uint32_t *>(e->pin + desc_bytes + 2 * nb4a);
    CK (cudaMemcpyAsync (info, d_info, (size_t)n_vbs * 16, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    std::vector<unsigned long long> F_end (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) F_end[v] = (unsigned long long)info[4 * v + 2] | ((unsigned long long)info[4 * v + 3] << 32);
    int rc = GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_normq_vb &S = vbs[v];
        if (dir == 1 && info[4 * v + 1]) { S.status = GZB_E_CORRUPT; rc = GZB_E_CORRUPT; e->err = "NORMQ: the stream does not match the lines"; continue; }
        if (dir == 0) {
            S.local_len = F_end[v];
            if (devptr && S.local_len > S.local_cap) { S.status = GZB_E_BADARG; rc = GZB_E_BADARG; e->err = "NORMQ: output capacity too small"; continue; }
            if (!devptr && S.local_len) CK (cudaMemcpyAsync (S.local, h[v].local, S.local_len, cudaMemcpyDeviceToHost, st));
        }
        else if (!devptr) {
            if (F_end[v]) CK (cudaMemcpyAsync (S.out, h[v].out, F_end[v], cudaMemcpyDeviceToHost, st));
            if (S.missing && S.n_lines) CK (cudaMemcpyAsync (S.missing, h[v].missing, S.n_lines, cudaMemcpyDeviceToHost, st));
        }
    }
    CK (cudaStreamSynchronize (st));
    return rc;
}

} // namespace

extern "C" int gzb_normq_gather (gzb_engine *e, gzb_normq_vb *vbs, uint32_t n_vbs, uint32_t flags) { return normq_run (e, vbs, n_vbs, flags, 0); }
extern "C" int gzb_normq_reconstruct (gzb_engine *e, gzb_normq_vb *vbs, uint32_t n_vbs, uint32_t flags) { return normq_run (e, vbs, n_vbs, flags, 1); }
