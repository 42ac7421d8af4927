// domain_vcf.cu — PBWT genotype-matrix transform on sm_100a.
//
// Reference functions replaced (relative to /root/reference/src/codec_pbwt.c):
//   codec_pbwt_compress (:244-287) = per row: permutation update (codec_pbwt_calculate_permutation :110-154), boustrophedon
//   traversal (:265) and run-length encoding into RUNS + FGRC (codec_pbwt_run_len_encode :213-238, _udpate_fgrc :181-210);
//   codec_pbwt_uncompress (:372-402) + pbwt_decode_one_line (:317-369).
//
// Rows are serial (row r's permutation is the stable partition of row r-1's by row r-1's alleles); columns are parallel.
// One CTA walks one VBlock's matrix: the permutation lives in shared memory, each row costs one gather, one boundary
// compaction (block scan) and one stable multi-key partition (one block scan per allele present, usually 2).  The few
// run boundaries of a row are then appended to RUNS/FGRC by one thread with exactly the reference's state machine.
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

using namespace gzb;

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr int PB_THREADS = 1024;
constexpr uint32_t PB_SMEM_W = 16384;          // widest matrix row whose permutation fits in shared memory (2 x 4 B x w + w + w)

__device__ uint32_t pb_block_excl_sum (uint32_t v, uint32_t *sm, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[warp] = inc;
    __syncthreads ();
    if (warp == 0) {
        uint32_t w = sm[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, wi, o); if (lane >= o) wi += t; }
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads ();
    uint32_t r = sm[warp] + inc - v;
    *total = sm[32];
    __syncthreads ();
    return r;
}

// order in which alleles are grouped (:122-130): '0' .. 255, 0 .. 36 (uint8 wrap-around), then . * % - &
__device__ void pb_key_order (const uint8_t *has, uint8_t *order, uint32_t *n_order)
{
    uint32_t no = 0;
    for (uint32_t a = '0'; a != (uint32_t)(('0' + 245) & 0xff); a = (a + 1) & 0xff) if (has[a]) order[no++] = (uint8_t)a;
    const uint8_t pseudo[5] = { '.', '*', '%', '-', '&' };
    for (int i = 0; i < 5; i++) if (has[pseudo[i]]) order[no++] = pseudo[i];
    *n_order = no;
}

// stable partition of perm by the alleles al[] (both in permuted order) into tmp; returns nothing, caller swaps
__device__ void pb_partition (const uint32_t *perm, uint32_t *tmp, const uint8_t *al, uint32_t w,
                              uint8_t *has, uint8_t *order, uint32_t *n_order, uint32_t *sm)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += PB_THREADS) has[i] = 0;
    __syncthreads ();
    for (uint32_t i = tid; i < w; i += PB_THREADS) has[al[i]] = 1;
    __syncthreads ();
    if (tid == 0) pb_key_order (has, order, n_order);
    __syncthreads ();
    const uint32_t C = (w + PB_THREADS - 1) / PB_THREADS, lo = min (w, tid * C), hi = min (w, lo + C);
    uint32_t base = 0;
    for (uint32_t kx = 0; kx < *n_order; kx++) {
        const uint8_t key = order[kx];
        uint32_t cnt = 0;
        for (uint32_t i = lo; i < hi; i++) cnt += al[i] == key;
        uint32_t tot;
        uint32_t pos = base + pb_block_excl_sum (cnt, sm, &tot);
        for (uint32_t i = lo; i < hi; i++) if (al[i] == key) tmp[pos++] = perm[i];
        base += tot;
    }
    __syncthreads ();
}

struct PbEnc {
    const uint8_t *ht; uint32_t n_lines, w;
    uint32_t *runs, *fgrc; uint32_t runs_cap, fgrc_cap;
    uint32_t *result;            // [0] n_runs, [1] n_fgrc, [2] error
    uint32_t *gperm;             // 2*w words of global scratch when w > PB_SMEM_W
};

__global__ void __launch_bounds__(PB_THREADS) k_pbwt_encode (PbEnc P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t w = P.w;
    const bool in_smem = w <= PB_SMEM_W;
    uint32_t *perm = in_smem ? reinterpret_cast<uint32_t *>(smem) : P.gperm;
    uint32_t *tmp  = perm + w;
    uint8_t  *al   = in_smem ? smem + 8 * (size_t)w : reinterpret_cast<uint8_t *>(P.gperm + 2 * (size_t)w);
    uint32_t *bpos = in_smem ? reinterpret_cast<uint32_t *>(smem + 8 * (size_t)w + ((w + 3) & ~3u)) : P.gperm + 2 * (size_t)w + ((w + 3) / 4);
    __shared__ uint32_t sm[33];
    __shared__ uint8_t has[256], order[256];
    __shared__ uint32_t n_order, s_nb;
    __shared__ uint32_t st_nr, st_nf, st_err; __shared__ uint8_t st_allele;
    const int tid = threadIdx.x;
    if (tid == 0) { st_nr = 0; st_nf = 0; st_err = 0; st_allele = 0; }
    for (uint32_t i = tid; i < w; i += PB_THREADS) perm[i] = i;             // first line: identity (:146-148)
    __syncthreads ();

    for (uint32_t r = 0; r < P.n_lines; r++) {
        const uint8_t *line = P.ht + (size_t)r * w;
        for (uint32_t i = tid; i < w; i += PB_THREADS) al[i] = line[perm[i]];   // :151-153
        __syncthreads ();
        // run boundaries of the row in traversal order (even rows forward, odd rows backward, :265)
        const bool back = r & 1;
        const uint8_t carry = st_allele;
        const uint32_t C = (w + PB_THREADS - 1) / PB_THREADS, lo = min (w, tid * C), hi = min (w, lo + C);
        uint32_t cnt = 0;
        for (uint32_t i = lo; i < hi; i++) {
            const uint8_t cur = al[back ? w - 1 - i : i];
            const uint8_t prev = i ? al[back ? w - i : i - 1] : carry;
            cnt += cur != prev;
        }
        uint32_t tot;
        uint32_t pos = pb_block_excl_sum (cnt, sm, &tot);
        for (uint32_t i = lo; i < hi; i++) {
            const uint8_t cur = al[back ? w - 1 - i : i];
            const uint8_t prev = i ? al[back ? w - i : i - 1] : carry;
            if (cur != prev) bpos[pos++] = i;
        }
        if (tid == 0) s_nb = tot;
        __syncthreads ();
        if (tid == 0 && !st_err) {                                          // codec_pbwt_run_len_encode (:213-238), one boundary at a time
            uint32_t nr = st_nr, nf = st_nf; uint8_t run_allele = st_allele;
            uint32_t prev_pos = 0;
            for (uint32_t b = 0; b <= s_nb; b++) {
                const uint32_t p = b < s_nb ? bpos[b] : w;
                if (p > prev_pos && nr) P.runs[nr - 1] += p - prev_pos;       // extend the current run
                if (b == s_nb) break;
                if (nr + 2 > P.runs_cap || nf + 3 > P.fgrc_cap) { st_err = 1; break; }
                const uint8_t done = run_allele;
                run_allele = al[back ? w - 1 - p : p];
                if (done == '0') {                                          // codec_pbwt_udpate_fgrc (:181-210)
                    if (nf && run_allele == (P.fgrc[nf - 1] & 0xff)) {
                        const uint32_t c = (P.fgrc[nf - 1] >> 8) + 1;
                        if (!(c & 0xffffffu)) { st_err = 2; break; }        // reference asserts on 24-bit overflow
                        P.fgrc[nf - 1] = (P.fgrc[nf - 1] & 0xff) | (c << 8);
                    }
                    else P.fgrc[nf++] = run_allele | (1u << 8);
                }
                else if (run_allele != '0') { P.fgrc[nf++] = run_allele | (1u << 8); P.runs[nr++] = 0; }
                P.runs[nr++] = 0;
                prev_pos = p;
            }
            st_nr = nr; st_nf = nf; st_allele = run_allele;
        }
        __syncthreads ();
        if (r + 1 < P.n_lines) {
            pb_partition (perm, tmp, al, w, has, order, &n_order, sm);
            uint32_t *t = perm; perm = tmp; tmp = t;
        }
    }
    if (tid == 0) {
        uint32_t nf = st_nf;
        const uint64_t len = (uint64_t)P.n_lines * w;                       // :274-276
        if (nf + 2 <= P.fgrc_cap) { P.fgrc[nf++] = (uint32_t)(len & 0xffffffffu); P.fgrc[nf++] = (uint32_t)(len >> 32); } else st_err = 1;
        P.result[0] = st_nr; P.result[1] = nf; P.result[2] = st_err;
    }
}

// ---------------------------------------------------------------------------------------------- decode
struct PbDec {
    const uint32_t *runs, *fgrc; uint32_t n_runs, n_fgrc;   // n_fgrc excludes the trailing length words
    uint64_t *cumpos;            // inclusive prefix of runs
    uint32_t *fgcum;             // inclusive prefix of FGRC counts
    uint8_t  *ral;               // allele of run k
    uint8_t  *ht; uint32_t n_lines, w;
    uint32_t *result;            // [2] error
    uint32_t *gperm;
};

__global__ void k_pbwt_prefix (PbDec P)            // single CTA: prefix sums of run lengths and FGRC counts
{
    __shared__ uint32_t sm[33];
    uint64_t acc = 0;
    for (uint32_t base = 0; base < P.n_runs; base += PB_THREADS) {
        uint32_t i = base + threadIdx.x, v = i < P.n_runs ? P.runs[i] : 0, tot;
        uint32_t ex = pb_block_excl_sum (v, sm, &tot);
        if (i < P.n_runs) P.cumpos[i] = acc + ex + v;
        acc += tot;
    }
    uint32_t a2 = 0;
    for (uint32_t base = 0; base < P.n_fgrc; base += PB_THREADS) {
        uint32_t i = base + threadIdx.x, v = i < P.n_fgrc ? (P.fgrc[i] >> 8) : 0, tot;
        uint32_t ex = pb_block_excl_sum (v, sm, &tot);
        if (i < P.n_fgrc) P.fgcum[i] = a2 + ex + v;
        a2 += tot;
    }
}

__global__ void k_pbwt_run_alleles (PbDec P)       // RUNS alternate background('0') / foreground, starting with background (:326)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.n_runs) return;
    if (!(k & 1)) { P.ral[k] = '0'; return; }
    const uint32_t j = k >> 1;                                              // ordinal of this foreground run
    uint32_t lo = 0, hi = P.n_fgrc;                                         // first group with fgcum > j (:353-358)
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (P.fgcum[mid] > j) hi = mid; else lo = mid + 1; }
    P.ral[k] = lo < P.n_fgrc ? (uint8_t)(P.fgrc[lo] & 0xff) : 0;
}

__global__ void __launch_bounds__(PB_THREADS) k_pbwt_decode (PbDec P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t w = P.w;
    const bool in_smem = w <= PB_SMEM_W;
    uint32_t *perm = in_smem ? reinterpret_cast<uint32_t *>(smem) : P.gperm;
    uint32_t *tmp  = perm + w;
    uint8_t  *al   = in_smem ? smem + 8 * (size_t)w : reinterpret_cast<uint8_t *>(P.gperm + 2 * (size_t)w);
    __shared__ uint32_t sm[33];
    __shared__ uint8_t has[256], order[256];
    __shared__ uint32_t n_order, s_klo;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < w; i += PB_THREADS) perm[i] = i;
    if (tid == 0) s_klo = 0;
    __syncthreads ();
    for (uint32_t r = 0; r < P.n_lines; r++) {
        uint8_t *line = P.ht + (size_t)r * w;
        const bool back = r & 1;
        const uint64_t p0 = (uint64_t)r * w;
        const uint32_t klo = s_klo;                                         // first run reaching into this row
        __syncthreads ();
        for (uint32_t i = tid; i < w; i += PB_THREADS) {
            const uint64_t p = p0 + i;
            uint32_t lo = klo, hi = P.n_runs;                               // first run k with cumpos[k] > p (zero-length runs never match)
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (P.cumpos[mid] > p) hi = mid; else lo = mid + 1; }
            const uint8_t a = lo < P.n_runs ? P.ral[lo] : 0;
            if (lo >= P.n_runs) P.result[2] = 3;
            const uint32_t o = back ? w - 1 - i : i;
            al[o] = a; line[perm[o]] = a;                                   // :340-343
            if (i == w - 1) s_klo = lo;
        }
        __syncthreads ();
        if (r + 1 < P.n_lines) {
            pb_partition (perm, tmp, al, w, has, order, &n_order, sm);
            uint32_t *t = perm; perm = tmp; tmp = t;
        }
    }
}

size_t pb_smem_bytes (uint32_t w) { return w <= PB_SMEM_W ? (size_t)8 * w + ((w + 3) & ~3u) + 4 * (size_t)w + 64 : 64; }

struct Carver {
    uint8_t *base; size_t off;
    template <typename T> T *take (size_t count) {
        size_t bytes = (count * sizeof (T) + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

} // namespace

extern "C" int gzb_pbwt_encode (gzb_engine *e, const void *ht, uint32_t n_lines, uint32_t ht_per_line,
                                uint32_t *runs, uint32_t runs_cap, uint32_t *n_runs,
                                uint32_t *fgrc, uint32_t fgrc_cap, uint32_t *n_fgrc, uint32_t flags)
{
    if (!e || !ht || !runs || !fgrc || !n_runs || !n_fgrc || !ht_per_line) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    const uint64_t len = (uint64_t)n_lines * ht_per_line;
    cudaStream_t st = e->stream;
    Carver c { nullptr, 0 };
    PbEnc P; memset (&P, 0, sizeof P);
    uint8_t *d_ht = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        P.result = c.take<uint32_t> (4);
        P.gperm  = c.take<uint32_t> (ht_per_line > PB_SMEM_W ? 4 * (size_t)ht_per_line + 64 : 1);
        if (!devptr) { d_ht = c.take<uint8_t> (len + 16); P.runs = c.take<uint32_t> (runs_cap + 1); P.fgrc = c.take<uint32_t> (fgrc_cap + 1); }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    if (devptr) { d_ht = (uint8_t *)ht; P.runs = runs; P.fgrc = fgrc; }
    else CK (cudaMemcpyAsync (d_ht, ht, len, cudaMemcpyHostToDevice, st));
    P.ht = d_ht; P.n_lines = n_lines; P.w = ht_per_line; P.runs_cap = runs_cap; P.fgrc_cap = fgrc_cap;
    const size_t smem = pb_smem_bytes (ht_per_line);
    CK (cudaFuncSetAttribute (k_pbwt_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t> (smem, 1024)));
    k_pbwt_encode<<<1, PB_THREADS, smem, st>>>(P); e->launches++;
    uint32_t res[4] = {0, 0, 0, 0};
    CK (cudaMemcpyAsync (res, P.result, sizeof res, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    if (res[2]) { e->err = res[2] == 2 ? "PBWT: more than 0xffffff consecutive foreground runs of one allele" : "PBWT: RUNS/FGRC capacity too small"; return GZB_E_BADARG; }
    *n_runs = res[0]; *n_fgrc = res[1];
    if (!devptr) {
        if (res[0]) CK (cudaMemcpyAsync (runs, P.runs, (size_t)res[0] * 4, cudaMemcpyDeviceToHost, st));
        if (res[1]) CK (cudaMemcpyAsync (fgrc, P.fgrc, (size_t)res[1] * 4, cudaMemcpyDeviceToHost, st));
        CK (cudaStreamSynchronize (st));
    }
    return GZB_OK;
}

extern "C" int gzb_pbwt_decode (gzb_engine *e, const uint32_t *runs, uint32_t n_runs, const uint32_t *fgrc, uint32_t n_fgrc,
                                uint32_t n_lines, void *ht, uint64_t ht_cap, uint64_t *ht_len, uint32_t flags)
{
    if (!e || !runs || !fgrc || !ht || !ht_len || n_fgrc < 2 || !n_lines || !n_runs) return GZB_E_BADARG;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    // the matrix length travels in the last two FGRC words (:293-301)
    uint32_t tail[2];
    if (devptr) { CK (cudaMemcpyAsync (tail, fgrc + (n_fgrc - 2), 8, cudaMemcpyDeviceToHost, st)); CK (cudaStreamSynchronize (st)); }
    else { tail[0] = fgrc[n_fgrc - 2]; tail[1] = fgrc[n_fgrc - 1]; }
    const uint64_t len = (uint64_t)tail[0] | ((uint64_t)tail[1] << 32);
    if (!len || len > ht_cap || len % n_lines) { e->err = "PBWT: bad matrix length"; return GZB_E_CORRUPT; }
    const uint32_t w = (uint32_t)(len / n_lines);
    Carver c { nullptr, 0 };
    PbDec P; memset (&P, 0, sizeof P);
    uint32_t *d_runs = nullptr, *d_fgrc = nullptr; uint8_t *d_ht = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        P.result = c.take<uint32_t> (4);
        P.cumpos = c.take<uint64_t> (n_runs + 1); P.fgcum = c.take<uint32_t> (n_fgrc + 1); P.ral = c.take<uint8_t> (n_runs + 16);
        P.gperm  = c.take<uint32_t> (w > PB_SMEM_W ? 4 * (size_t)w + 64 : 1);
        if (!devptr) { d_runs = c.take<uint32_t> (n_runs + 1); d_fgrc = c.take<uint32_t> (n_fgrc + 1); d_ht = c.take<uint8_t> (len + 16); }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    if (devptr) { d_runs = (uint32_t *)runs; d_fgrc = (uint32_t *)fgrc; d_ht = (uint8_t *)ht; }
    else {
        CK (cudaMemcpyAsync (d_runs, runs, (size_t)n_runs * 4, cudaMemcpyHostToDevice, st));
        CK (cudaMemcpyAsync (d_fgrc, fgrc, (size_t)n_fgrc * 4, cudaMemcpyHostToDevice, st));
    }
    P.runs = d_runs; P.fgrc = d_fgrc; P.n_runs = n_runs; P.n_fgrc = n_fgrc - 2; P.ht = d_ht; P.n_lines = n_lines; P.w = w;
    CK (cudaMemsetAsync (P.result, 0, 16, st));
    k_pbwt_prefix<<<1, PB_THREADS, 0, st>>>(P);
    k_pbwt_run_alleles<<<(n_runs + 255) / 256, 256, 0, st>>>(P);
    const size_t smem = pb_smem_bytes (w);
    CK (cudaFuncSetAttribute (k_pbwt_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t> (smem, 1024)));
    k_pbwt_decode<<<1, PB_THREADS, smem, st>>>(P);
    e->launches += 3;
    uint32_t res[4] = {0, 0, 0, 0};
    CK (cudaMemcpyAsync (res, P.result, sizeof res, cudaMemcpyDeviceToHost, st));
    if (!devptr) CK (cudaMemcpyAsync (ht, d_ht, len, cudaMemcpyDeviceToHost, st));
    CK (cudaStreamSynchronize (st));
    if (res[2]) { e->err = "PBWT: runs do not cover the matrix"; return GZB_E_CORRUPT; }
    *ht_len = len;
    return GZB_OK;
}
