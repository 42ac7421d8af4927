// domain_homp.cu — the two homopolymer codecs of Ultima data: HOMP for QUAL (reference src/codec_homp.c) and T0 for the t0:Z tag
// (src/codec_t0.c).  Ultima reads carry one quality pattern per homopolymer run of the sequence (symmetric around its middle, and nothing
// but 'I' after the first 'I'), resp. one t0 character per run, so the string is condensed run by run before its sub-codec:
//   HOMP  a run of h > 1 bases whose qualities are a palindrome (and 'I' to the end once 'I' appeared, :152-165): its first (h + 1) / 2
//         qualities, stopping after an 'I' (:167-170); any other run: its first quality | 0x80, then the rest verbatim (:171-176)
//   T0    a run whose t0 characters are all equal: one of them (:84-85); else the first | 0x80, then the rest (:87-92)
// condense  = the first pass of codec_homp_compress (:132-190) / codec_t0_compress (:69-109): every line by its own thread (lines are
//             independent: lengths, a scan, then the bytes), the condensed strings back to back — what the sub-codec then reads line by line
// expand    = codec_homp_reconstruct (:213-276) / codec_t0_reconstruct (:137-179) for every line: how many bytes a line consumes depends on
//             the bytes themselves, so the lines of a VBlock are walked in order by one thread; VBlocks are independent
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr uint8_t TOP_QUAL = 'I';

struct HpVb {
    const uint8_t *txt; const uint64_t *str_off; const uint32_t *str_len; const uint64_t *seq_off;
    uint8_t *local; uint8_t *out; uint8_t *missing;
    uint32_t *new_len;             // condense: the condensed length of every line
    unsigned long long *pos;       // condense: where every line's condensed string starts in local
    uint32_t *info;                // [0] error, [1..2] total condensed / consumed bytes
    unsigned long long local_len, local_cap, out_cap;
    uint32_t n_lines, mode;        // mode 0 HOMP, 1 T0
};

__device__ __forceinline__ uint32_t hp_len_at (const uint8_t *seq, uint32_t len, uint32_t i)      // homopolymer_len (src/strings.h:194-201)
{
    const uint8_t b = seq[i]; uint32_t k = i + 1;
    while (k < len && seq[k] == b) k++;
    return k - i;
}
// is the run s[i .. i + h) condensable
template <int MODE> __device__ __forceinline__ bool hp_condensable (const uint8_t *s, uint32_t i, uint32_t h)
{
    if (MODE == 1) { for (uint32_t k = 1; k < h; k++) if (s[i + k] != s[i]) return false; return true; }        // str_is_monochar
    uint8_t prev = 0;
    for (uint32_t k = 0; k < (h + 1) / 2; k++) {
        const uint8_t a = s[i + k], m = s[i + h - 1 - k];
        if (a != m || (prev == TOP_QUAL && a != TOP_QUAL)) return false;
        prev = a;
    }
    return true;
}
// the condensed form of one line; WRITE = 0 counts only.  A line the reference skips (HOMP: length <= 1, :137; T0: empty, :75) stays as it is.
template <int MODE, int WRITE> __device__ __forceinline__ uint32_t hp_condense_line (const uint8_t *s, const uint8_t *seq, uint32_t len, uint8_t *dst)
{
    if (MODE == 0 ? len <= 1 : len == 0) { if (WRITE) for (uint32_t i = 0; i < len; i++) dst[i] = s[i]; return len; }
    uint32_t n = 0;
    for (uint32_t i = 0; i < len; i++) {
        const uint32_t h = hp_len_at (seq, len, i);
        if (h > 1) {
            if (hp_condensable<MODE> (s, i, h)) {
                if (MODE == 1) { if (WRITE) dst[n] = s[i]; n++; }
                else for (uint32_t k = 0; k < (h + 1) / 2; k++) { const uint8_t c = s[i + k]; if (WRITE) dst[n] = c; n++; if (c == TOP_QUAL) break; }
            }
            else {
                if (WRITE) dst[n] = s[i] | 0x80; n++;
                for (uint32_t k = 1; k < h; k++) { if (WRITE) dst[n] = s[i + k]; n++; }
            }
            i += h - 1;
        }
        else { if (WRITE) dst[n] = s[i]; n++; }
    }
    return n;
}

template <int MODE> __global__ void __launch_bounds__(128) k_hp_lengths (const HpVb *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const HpVb &V = vbs[blk_vb[blockIdx.x]];
    const uint32_t li = blk_first[blockIdx.x] + threadIdx.x;
    if (li >= V.n_lines || threadIdx.x >= 128) return;
    V.new_len[li] = hp_condense_line<MODE, 0> (V.txt + V.str_off[li], V.txt + V.seq_off[li], V.str_len[li], nullptr);
}
// pos = exclusive prefix sums of new_len, one CTA per VBlock
__global__ void __launch_bounds__(1024) k_hp_prefix (const HpVb *vbs)
{
    const HpVb &V = vbs[blockIdx.x];
    __shared__ unsigned long long s_warp[32], s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads ();
    for (uint32_t base = 0; base < V.n_lines; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long v = i < V.n_lines ? V.new_len[i] : 0;
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
        if (i < V.n_lines) V.pos[i] = before;
        __syncthreads ();
        if (tid == 1023) s_carry = before + v;
        __syncthreads ();
    }
    if (tid == 0) { V.info[1] = (uint32_t)s_carry; V.info[2] = (uint32_t)(s_carry >> 32); if (s_carry > V.local_cap) V.info[0] = 2; }
}
// A thread condenses its line into a slot of shared memory; the warp then copies the 32 slots out one after the other, 32 consecutive bytes per
// instruction (a thread storing its line to global memory byte by byte cost 10 x the DRAM reads: every byte a partial sector).  Lines above
// HP_SLOT bytes are written directly.
constexpr uint32_t HP_SLOT = 320;
template <int MODE> __global__ void __launch_bounds__(128) k_hp_write (const HpVb *vbs, const uint32_t *blk_vb, const uint32_t *blk_first)
{
    const HpVb &V = vbs[blk_vb[blockIdx.x]];
    __shared__ uint8_t slots[128][HP_SLOT + 1];          // (an odd stride: 32 lanes storing byte k of their slots hit 32 banks; 320 would be 2)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t li = blk_first[blockIdx.x] + threadIdx.x;
    const bool mine = li < V.n_lines && !V.info[0];
    uint32_t n = 0; unsigned long long pos = 0;
    if (mine) {
        const uint32_t len = V.str_len[li];
        pos = V.pos[li];
        if (len <= HP_SLOT) n = hp_condense_line<MODE, 1> (V.txt + V.str_off[li], V.txt + V.seq_off[li], len, slots[threadIdx.x]);
        else hp_condense_line<MODE, 1> (V.txt + V.str_off[li], V.txt + V.seq_off[li], len, V.local + pos);
    }
    __syncwarp ();
    for (int t = 0; t < 32; t++) {
        const uint32_t tn = __shfl_sync (0xffffffffu, n, t);
        const unsigned long long tp = __shfl_sync (0xffffffffu, pos, t);
        const uint8_t *src = slots[warp * 32 + t];
        for (uint32_t i = lane; i < tn; i += 32) V.local[tp + i] = src[i];
    }
}

// every line of a VBlock, in order (the next line starts where this one stopped reading).  One WARP per VBlock with one lane at work: the walk
// branches on the data at every byte, and 32 VBlocks sharing a warp executed each other's branches (13.8 s for 32 VBlocks of 13.8 MB, measured).
template <int MODE> __global__ void k_hp_expand (const HpVb *vbs, uint32_t n_vbs)
{
    const uint32_t v = blockIdx.x;
    if (v >= n_vbs || threadIdx.x) return;
    const HpVb &V = vbs[v];
    const uint8_t *c = V.local; unsigned long long next = 0, at = 0;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t len = V.str_len[li];
        if (V.missing) V.missing[li] = 0;
        if (!len) continue;
        uint8_t *o = V.out + at; const uint8_t *seq = V.txt + V.seq_off[li];
        at += len;
        if (at > V.out_cap) { V.info[0] = 2; return; }
        if (next >= V.local_len) { V.info[0] = 1; return; }
        if (MODE == 0 && c[next] == ' ') { o[0] = '*'; if (V.missing) V.missing[li] = 1; next++; continue; }   // SAM missing quality (:241-244)
        uint32_t n = 0;
        for (uint32_t i = 0; i < len; i++) {
            const uint32_t h = hp_len_at (seq, len, i);
            if (h > 1) {
                if (next >= V.local_len) { V.info[0] = 1; return; }
                if (c[next] & 0x80) {                                        // non-condensable (:252-256, t0 :157-161)
                    if (next + h > V.local_len) { V.info[0] = 1; return; }
                    o[n++] = c[next++] & 0x7f;
                    for (uint32_t k = 1; k < h; k++) o[n++] = c[next++];
                }
                else if (MODE == 1) { const uint8_t ch = c[next++]; for (uint32_t k = 0; k < h; k++) o[n++] = ch; }   // :163-167
                else {                                                       // :258-266
                    uint8_t prev = 0;
                    for (uint32_t k = 0; k < (h + 1) / 2; k++) {
                        if (prev != TOP_QUAL) { if (next >= V.local_len) { V.info[0] = 1; return; } prev = c[next++]; }
                        o[n++] = prev;
                    }
                    uint32_t m = n - 1 - (h & 1);                            // the mirror starts before the middle element of an odd run
                    for (uint32_t k = 0; k < h / 2; k++) o[n++] = o[m--];
                }
                i += h - 1;
            }
            else { if (next >= V.local_len) { V.info[0] = 1; return; } o[n++] = c[next++]; }
        }
    }
    V.info[1] = (uint32_t)next; V.info[2] = (uint32_t)(next >> 32);
    if (next != V.local_len) V.info[0] = 1;                                  // the stream must be used up exactly
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

int hp_run (gzb_engine *e, gzb_homp_vb *vbs, uint32_t n_vbs, uint32_t flags, int mode, int expand)
{
    if (!e || (!vbs && n_vbs) || mode < 0 || mode > 1) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<HpVb> h (n_vbs);
    std::vector<uint32_t> bvb, bfirst;
    std::vector<uint64_t> total (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_homp_vb &S = vbs[v]; S.status = GZB_OK;
        if ((S.n_lines && (!S.str_len || !S.seq_off || (!expand && !S.str_off))) || (!S.txt && S.txt_len) || (expand && !S.local && S.local_len)) return GZB_E_BADARG;
        if (!devptr) for (uint32_t i = 0; i < S.n_lines; i++) total[v] += S.str_len[i];
        else total[v] = expand ? S.out_cap : S.local_cap;
        if (!devptr && (expand ? total[v] > S.out_cap : total[v] > S.local_cap)) { e->err = "HOMP / T0: output capacity too small"; return GZB_E_BADARG; }
        if (!expand) for (uint32_t f = 0; f < S.n_lines; f += 128) { bvb.push_back (v); bfirst.push_back (f); }
    }
    Carver c { nullptr, 0 };
    HpVb *d_vbs = nullptr; uint32_t *d_bvb = nullptr, *d_bfirst = nullptr, *d_info = nullptr;
    const size_t nb4 = bvb.size () * 4, nb4a = (nb4 + 255) & ~(size_t)255, desc_bytes = ((size_t)n_vbs * sizeof (HpVb) + 255) & ~(size_t)255;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<HpVb> (n_vbs); d_bvb = c.take<uint32_t> (bvb.size () + 1); d_bfirst = c.take<uint32_t> (bfirst.size () + 1);
        d_info = c.take<uint32_t> ((size_t)n_vbs * 4);
        for (uint32_t v = 0; v < n_vbs; v++) {
            const gzb_homp_vb &S = vbs[v]; HpVb &D = h[v];
            D.n_lines = S.n_lines; D.mode = mode; D.local_cap = S.local_cap; D.out_cap = S.out_cap; D.local_len = expand ? S.local_len : 0;
            D.info = d_info ? d_info + 4 * (size_t)v : nullptr;
            D.txt = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.str_len = devptr ? S.str_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.seq_off = devptr ? S.seq_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.str_off = expand ? nullptr : devptr ? S.str_off : c.take<uint64_t> ((size_t)S.n_lines + 1);
            D.new_len = expand ? nullptr : (devptr && S.new_len) ? S.new_len : c.take<uint32_t> ((size_t)S.n_lines + 1);
            D.pos = expand ? nullptr : c.take<unsigned long long> ((size_t)S.n_lines + 1);
            D.local = devptr ? (uint8_t *)S.local : c.take<uint8_t> ((expand ? S.local_len : total[v]) + 16);
            D.out = !expand ? nullptr : devptr ? (uint8_t *)S.out : c.take<uint8_t> (total[v] + 16);
            D.missing = (!expand || !S.missing) ? nullptr : devptr ? S.missing : c.take<uint8_t> ((size_t)S.n_lines + 1);
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, desc_bytes + 2 * nb4a + (size_t)n_vbs * 16 + 512); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs && !devptr; v++) {
        const gzb_homp_vb &S = vbs[v]; HpVb &D = h[v];
        if (S.n_lines) {
            CK (cudaMemcpyAsync ((void *)D.str_len, S.str_len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
            CK (cudaMemcpyAsync ((void *)D.seq_off, S.seq_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
            if (!expand) CK (cudaMemcpyAsync ((void *)D.str_off, S.str_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
        }
        if (S.txt_len) CK (cudaMemcpyAsync ((void *)D.txt, S.txt, S.txt_len, cudaMemcpyHostToDevice, st));
        if (expand && S.local_len) CK (cudaMemcpyAsync (D.local, S.local, S.local_len, cudaMemcpyHostToDevice, st));
    }
    memcpy (e->pin, h.data (), (size_t)n_vbs * sizeof (HpVb));
    CK (cudaMemcpyAsync (d_vbs, e->pin, (size_t)n_vbs * sizeof (HpVb), cudaMemcpyHostToDevice, st));
    if (nb4) {
        memcpy (e->pin + desc_bytes, bvb.data (), nb4); memcpy (e->pin + desc_bytes + nb4a, bfirst.data (), nb4);
        CK (cudaMemcpyAsync (d_bvb, e->pin + desc_bytes, nb4, cudaMemcpyHostToDevice, st));
        CK (cudaMemcpyAsync (d_bfirst, e->pin + desc_bytes + nb4a, nb4, cudaMemcpyHostToDevice, st));
    }
    CK (cudaMemsetAsync (d_info, 0, (size_t)n_vbs * 16, st));
    if (expand) {
        if (mode == 0) k_hp_expand<0><<<n_vbs, 32, 0, st>>>(d_vbs, n_vbs); else k_hp_expand<1><<<n_vbs, 32, 0, st>>>(d_vbs, n_vbs);
        e->launches++;
    }
    else {
        const uint32_t nb = (uint32_t)bvb.size ();
        if (nb) { if (mode == 0) k_hp_lengths<0><<<nb, 128, 0, st>>>(d_vbs, d_bvb, d_bfirst); else k_hp_lengths<1><<<nb, 128, 0, st>>>(d_vbs, d_bvb, d_bfirst); }
        k_hp_prefix<<<n_vbs, 1024, 0, st>>>(d_vbs);
        if (nb) { if (mode == 0) k_hp_write<0><<<nb, 128, 0, st>>>(d_vbs, d_bvb, d_bfirst); else k_hp_write<1><<<nb, 128, 0, st>>>(d_vbs, d_bvb, d_bfirst); }
        e->launches += 3;
    }
    uint32_t *info = reinterpret_cast<uint32_t *>(e->pin + desc_bytes + 2 * nb4a);
    CK (cudaMemcpyAsync (info, d_info, (size_t)n_vbs * 16, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    int rc = GZB_OK;
    for (uint32_t v = 0; v < n_vbs; v++) {
        gzb_homp_vb &S = vbs[v];
        const unsigned long long n = (unsigned long long)info[4 * v + 1] | ((unsigned long long)info[4 * v + 2] << 32);
        if (info[4 * v]) {
            S.status = info[4 * v] == 2 ? GZB_E_BADARG : GZB_E_CORRUPT; rc = S.status;
            e->err = info[4 * v] == 2 ? "HOMP / T0: output capacity too small" : "HOMP / T0: the stream does not match the lines";
            continue;
        }
        if (!expand) {
            S.local_len = n;
            if (!devptr) {
                if (n) CK (cudaMemcpyAsync (S.local, h[v].local, n, cudaMemcpyDeviceToHost, st));
                if (S.new_len && S.n_lines) CK (cudaMemcpyAsync (S.new_len, h[v].new_len, (size_t)S.n_lines * 4, cudaMemcpyDeviceToHost, st));
            }
        }
        else if (!devptr) {
            if (total[v]) CK (cudaMemcpyAsync (S.out, h[v].out, total[v], cudaMemcpyDeviceToHost, st));
            if (S.missing && S.n_lines) CK (cudaMemcpyAsync (S.missing, h[v].missing, S.n_lines, cudaMemcpyDeviceToHost, st));
        }
    }
    CK (cudaStreamSynchronize (st));
    return rc;
}

} // namespace

extern "C" int gzb_homp_condense (gzb_engine *e, gzb_homp_vb *vbs, uint32_t n_vbs, int mode, uint32_t flags) { return hp_run (e, vbs, n_vbs, flags, mode, 0); }
extern "C" int gzb_homp_expand   (gzb_engine *e, gzb_homp_vb *vbs, uint32_t n_vbs, int mode, uint32_t flags) { return hp_run (e, vbs, n_vbs, flags, mode, 1); }
