// domain_longr.cu — LONGR long-read quality model on sm_100a.
//
// Reference functions replaced (relative to /root/reference/src): codec_longr_compress before its sub-codec
// (codec_longr.c:161-247: codec_longr_calc_channels :138-159, counting sort :193-230, lens :232-240),
// codec_longr_reconstruct for all reads of a VBlock (:270-373), the state machine codec_longr_update_state /
// _alg_init / _alg_init_read (codec_longr_alg.c:108-159).
//
// The channel computation is serial across the whole VBlock (state tables indexed by a 21-bit context are carried
// from read to read): one thread walks one VBlock, many VBlocks run concurrently.  The thread records, per
// base, its channel and its rank inside the channel, which turns the reference's second (also serial) pass — the
// stable counting sort of the qualities by channel — into a fully parallel scatter.
#include <cstring>
#include <vector>
#include <string>
#include "../../include/gzb200.h"
#include "gzb_internal.cuh"
#include "engine.h"

using namespace gzb;

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { e->err = std::string (#call) + ": " + cudaGetErrorString (_e); return GZB_E_CUDA; } } while (0)

namespace {

constexpr uint32_t NCTX = 1u << 21, NCHAN = 1u << 16, NQ9 = 1u << 9;

struct LrVb {
    const uint8_t *txt; const uint64_t *seq_off, *qual_off; const uint32_t *len; const uint8_t *is_rev;
    uint32_t n_lines;
    uint16_t *avg_sums, *err_sums;       // [1<<21] each (codec_longr_alg.c:99-100)
    uint32_t *chan_num;                  // [65536] bases per channel, then reused as next_of_chan
    uint16_t *base_chan; uint32_t *base_rank; uint8_t *base_q;   // per base, in processing order
    uint8_t  *values; uint32_t *lens_be; uint8_t *qual_out;
    uint64_t  total;
    uint8_t   v2b[256];
};

__device__ __forceinline__ uint32_t acgt_code (uint32_t c)                  // _acgt_encode (reference.c:45-58)
{
    switch (c) {
        case 'C': case 'c': case 'Y': case 'y': case 'S': case 's': case 'B': case 'b': return 1;
        case 'G': case 'g': case 'K': case 'k': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
    }
}
__device__ __forceinline__ uint32_t acgt_code_comp (uint32_t c)             // _acgt_encode_comp (reference.c:63-75)
{
    switch (c) {
        case 'A': case 'a': return 3;
        case 'C': case 'c': case 'M': case 'm': return 2;
        case 'G': case 'g': case 'R': case 'r': case 'S': case 's': case 'V': case 'v': return 1;
        default: return 0;
    }
}

// channel word (codec_longr_alg.c:65-95), LSB first: B:12 | difq:4 | qbin:5 | avg:5 | err_c:2
struct LrState { uint16_t *avg, *err; uint32_t *tot; const uint8_t *v2b; uint32_t chan; };

__device__ __forceinline__ void lr_update (LrState &s, uint32_t b, int32_t q1, int32_t q2)   // codec_longr_update_state :108-136
{
    uint32_t c = s.chan;
    const uint32_t nc = c & 0x1fffffu, qn = (c >> 12) & 0x1ffu;
    const int32_t err = q1 - (((int32_t)s.avg[nc] + 8) >> 4);
    s.avg[nc] = (uint16_t)(s.avg[nc] + err);
    const int32_t ae = err < 0 ? -err : err;
    s.err[nc] = (uint16_t)((int32_t)s.err[nc] + ae - (((int32_t)s.err[nc] + 8) >> 4));
    s.tot[qn] = s.tot[qn] + (uint32_t)ae - ((s.tot[qn] + 8u) >> 4);
    const uint32_t B = (((c & 0xfffu) << 2) | b) & 0xfffu;
    const int32_t d = q1 - q2;
    const uint32_t il = d < 0 ? (((uint32_t)(-d)) << 1) - 1 : ((uint32_t)d) << 1;     // INTERLACE (context.h:100)
    const uint32_t difq = il < 15 ? il : 15;
    const uint32_t qbin = s.v2b[q1 & 0xff] & 0x1fu;
    c = (c & ~0x1fffffu) | B | (difq << 12) | (qbin << 16);
    const uint32_t nc2 = c & 0x1fffffu, qn2 = (c >> 12) & 0x1ffu;
    const uint32_t avg = s.v2b[((((int32_t)s.avg[nc2]) + 8) >> 4) & 0xff] & 0x1fu;
    const uint32_t tot = s.tot[qn2];                                        // TOTAL_ERR_SHIFT - AVG_SHIFT = 0
    const uint32_t ae2 = s.err[nc2];
    const uint32_t ec = ae2 < (tot >> 1) ? 0 : ae2 < tot ? 1 : ae2 < (tot << 1) ? 2 : 3;
    s.chan = (c & ~(0x7fu << 21)) | (avg << 21) | (ec << 26);
}

__device__ __forceinline__ void lr_init_read (LrState &s, const uint8_t *seq, uint32_t len, bool rev)   // codec_longr_alg_init_read :148-159
{
    s.chan = 0;
    for (int i = 0; i < 3; i++)
        lr_update (s, rev ? acgt_code_comp ((int)len - 1 - i >= 0 ? seq[len - 1 - i] : 'T') : acgt_code (i < (int)len ? seq[i] : 'A'), 0, 0);
}

// state tables: chan_avgs_sums[n] = qbin(n) << AVG_SHIFT (:138-146); err sums 0; chan counters 0
__global__ void k_longr_init (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.y];
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < NCTX; n += gridDim.x * blockDim.x) {
        V.avg_sums[n] = (uint16_t)(((n >> 16) & 0x1f) << 4);
        V.err_sums[n] = 0;
        if (n < NCHAN) V.chan_num[n] = 0;
    }
}

// one VBlock per CTA; thread 0 walks the reads (codec_longr_calc_channels :138-159 for every line, :185-203)
__global__ void k_longr_channels (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t tot[NQ9];
    __shared__ uint8_t v2b[256];
    for (int i = threadIdx.x; i < (int)NQ9; i += blockDim.x) tot[i] = 0x10101010u;   // memset (.., 1<<4, ..) on uint32 (:142)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) v2b[i] = V.v2b[i];
    __syncthreads ();
    if (threadIdx.x) return;
    LrState s; s.avg = V.avg_sums; s.err = V.err_sums; s.tot = tot; s.v2b = v2b; s.chan = 0;
    uint64_t nb = 0;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t L = V.len[li];
        if (!L) continue;
        const uint8_t *seq = V.txt + V.seq_off[li], *q = V.txt + V.qual_off[li];
        const bool rev = V.is_rev ? V.is_rev[li] : false;
        lr_init_read (s, seq, L, rev);
        int32_t prev = 0;
        for (uint32_t k = 0; k < L; k++, nb++) {
            const uint32_t i = rev ? L - 1 - k : k;
            const uint32_t ch = (s.chan >> 12) & 0xffffu;
            V.base_chan[nb] = (uint16_t)ch;
            V.base_rank[nb] = V.chan_num[ch]++;
            const uint32_t b = rev ? acgt_code_comp (i >= 3 ? seq[i - 3] : 'T') : acgt_code (i + 3 < L ? seq[i + 3] : 'A');
            const int32_t qq = (uint8_t)(q[i] - '!');
            V.base_q[nb] = (uint8_t)qq;
            lr_update (s, b, qq, prev);
            prev = qq;
        }
    }
}

// lens (BGEN32, :237-240) and the exclusive prefix next_of_chan (:193-196); one CTA per VBlock
__global__ void __launch_bounds__(1024) k_longr_prefix (const LrVb *vbs, int write_lens)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t sm[33];
    uint32_t acc = 0;
    for (uint32_t base = 0; base < NCHAN; base += 1024) {
        const uint32_t c = base + threadIdx.x;
        uint32_t v = write_lens ? V.chan_num[c] : __byte_perm (V.lens_be[c], 0, 0x0123);
        if (write_lens) V.lens_be[c] = __byte_perm (v, 0, 0x0123);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sm[warp] = inc;
        __syncthreads ();
        if (warp == 0) { uint32_t w = sm[lane], wi = w; for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync (0xffffffffu, wi, o); if (lane >= o) wi += t; } sm[lane] = wi - w; if (lane == 31) sm[32] = wi; }
        __syncthreads ();
        V.chan_num[c] = acc + sm[warp] + inc - v;
        acc += sm[32];
        __syncthreads ();
    }
}

// stable scatter of the qualities into their channel segments (:205-230)
__global__ void k_longr_scatter (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.y];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V.total; i += (uint64_t)gridDim.x * blockDim.x)
        V.values[V.chan_num[V.base_chan[i]] + V.base_rank[i]] = V.base_q[i];
}

// codec_longr_recon_one_read (:270-296) for every read; one VBlock per CTA, thread 0
__global__ void k_longr_decode (const LrVb *vbs)
{
    const LrVb &V = vbs[blockIdx.x];
    __shared__ uint32_t tot[NQ9];
    __shared__ uint8_t v2b[256];
    for (int i = threadIdx.x; i < (int)NQ9; i += blockDim.x) tot[i] = 0x10101010u;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) v2b[i] = V.v2b[i];
    __syncthreads ();
    if (threadIdx.x) return;
    LrState s; s.avg = V.avg_sums; s.err = V.err_sums; s.tot = tot; s.v2b = v2b; s.chan = 0;
    uint8_t *out = V.qual_out;
    for (uint32_t li = 0; li < V.n_lines; li++) {
        const uint32_t L = V.len[li];
        if (!L) continue;
        const uint8_t *seq = V.txt + V.seq_off[li];
        const bool rev = V.is_rev ? V.is_rev[li] : false;
        lr_init_read (s, seq, L, rev);
        int32_t prev = 0;
        for (uint32_t k = 0; k < L; k++) {
            const uint32_t i = rev ? L - 1 - k : k;
            const uint32_t b = rev ? acgt_code_comp (i >= 3 ? seq[i - 3] : 'T') : acgt_code (i + 3 < L ? seq[i + 3] : 'A');
            const uint32_t ch = (s.chan >> 12) & 0xffffu;
            const int32_t qq = V.values[V.chan_num[ch]++];
            lr_update (s, b, qq, prev);
            prev = qq;
            out[i] = (uint8_t)(qq + '!');
        }
        out += L;
    }
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

int longr_run (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags, bool encode)
{
    if (!e || !vbs) return GZB_E_BADARG;
    if (!n_vbs) return GZB_OK;
    cudaSetDevice (e->device);
    const bool devptr = flags & GZB_DEVICE_PTRS;
    cudaStream_t st = e->stream;
    std::vector<LrVb> h (n_vbs);
    std::vector<uint64_t> total (n_vbs, 0);
    for (uint32_t v = 0; v < n_vbs; v++) {
        if (devptr) total[v] = vbs[v].txt_len;                  // lengths live on the device: bound by the text size
        else for (uint32_t i = 0; i < vbs[v].n_lines; i++) total[v] += vbs[v].len[i];
    }
    Carver c { nullptr, 0 };
    LrVb *d_vbs = nullptr;
    for (int pass = 0; pass < 2; pass++) {
        c.off = 0;
        d_vbs = c.take<LrVb> (n_vbs);
        for (uint32_t v = 0; v < n_vbs; v++) {
            LrVb &D = h[v]; const gzb_longr_vb &S = vbs[v];
            D.n_lines = S.n_lines; D.total = total[v];
            memcpy (D.v2b, S.value_to_bin, 256);
            D.avg_sums = c.take<uint16_t> (NCTX); D.err_sums = c.take<uint16_t> (NCTX); D.chan_num = c.take<uint32_t> (NCHAN);
            D.txt      = devptr ? (const uint8_t *)S.txt : c.take<uint8_t> (S.txt_len + 16);
            D.seq_off  = devptr ? S.seq_off : c.take<uint64_t> (S.n_lines + 1);
            D.qual_off = devptr ? S.qual_off : (encode ? c.take<uint64_t> (S.n_lines + 1) : nullptr);
            D.len      = devptr ? S.len : c.take<uint32_t> (S.n_lines + 1);
            D.is_rev   = S.is_rev ? (devptr ? S.is_rev : c.take<uint8_t> (S.n_lines + 1)) : nullptr;
            D.values   = devptr ? (uint8_t *)S.values : c.take<uint8_t> (total[v] + 16);
            D.lens_be  = devptr ? S.lens_be : c.take<uint32_t> (NCHAN);
            D.qual_out = (!encode) ? (devptr ? (uint8_t *)S.qual_out : c.take<uint8_t> (total[v] + 16)) : nullptr;
            if (encode) { D.base_chan = c.take<uint16_t> (total[v] + 1); D.base_rank = c.take<uint32_t> (total[v] + 1); D.base_q = c.take<uint8_t> (total[v] + 1); }
            else D.base_chan = nullptr, D.base_rank = nullptr, D.base_q = nullptr;
        }
        if (pass == 0) { int rc = engine_reserve (e, c.off, 4096); if (rc) return rc; c.base = e->ws; }
    }
    for (uint32_t v = 0; v < n_vbs; v++) {
        LrVb &D = h[v]; const gzb_longr_vb &S = vbs[v];
        if (!devptr) {
            if (S.txt_len) CK (cudaMemcpyAsync ((void *)D.txt, S.txt, S.txt_len, cudaMemcpyHostToDevice, st));
            if (S.n_lines) {
                CK (cudaMemcpyAsync ((void *)D.seq_off, S.seq_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
                if (encode) CK (cudaMemcpyAsync ((void *)D.qual_off, S.qual_off, (size_t)S.n_lines * 8, cudaMemcpyHostToDevice, st));
                CK (cudaMemcpyAsync ((void *)D.len, S.len, (size_t)S.n_lines * 4, cudaMemcpyHostToDevice, st));
                if (S.is_rev) CK (cudaMemcpyAsync ((void *)D.is_rev, S.is_rev, S.n_lines, cudaMemcpyHostToDevice, st));
            }
            if (!encode) {
                if (total[v]) CK (cudaMemcpyAsync (D.values, S.values, total[v], cudaMemcpyHostToDevice, st));
                CK (cudaMemcpyAsync (D.lens_be, S.lens_be, NCHAN * 4, cudaMemcpyHostToDevice, st));
            }
        }
    }
    CK (cudaMemcpyAsync (d_vbs, h.data (), n_vbs * sizeof (LrVb), cudaMemcpyHostToDevice, st));
    k_longr_init<<<dim3 (256, n_vbs), 256, 0, st>>>(d_vbs);
    if (encode) {
        k_longr_channels<<<n_vbs, 256, 0, st>>>(d_vbs);
        k_longr_prefix<<<n_vbs, 1024, 0, st>>>(d_vbs, 1);
        k_longr_scatter<<<dim3 (512, n_vbs), 256, 0, st>>>(d_vbs);
        e->launches += 4;
    }
    else {
        k_longr_prefix<<<n_vbs, 1024, 0, st>>>(d_vbs, 0);
        k_longr_decode<<<n_vbs, 256, 0, st>>>(d_vbs);
        e->launches += 3;
    }
    if (!devptr)
        for (uint32_t v = 0; v < n_vbs; v++) {
            if (encode) {
                if (total[v]) CK (cudaMemcpyAsync (vbs[v].values, h[v].values, total[v], cudaMemcpyDeviceToHost, st));
                CK (cudaMemcpyAsync (vbs[v].lens_be, h[v].lens_be, NCHAN * 4, cudaMemcpyDeviceToHost, st));
            }
            else if (total[v]) CK (cudaMemcpyAsync (vbs[v].qual_out, h[v].qual_out, total[v], cudaMemcpyDeviceToHost, st));
        }
    CK (cudaStreamSynchronize (st));
    CK (cudaGetLastError ());
    return GZB_OK;
}

} // namespace

extern "C" int gzb_longr_encode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags) { return longr_run (e, vbs, n_vbs, flags, true); }
extern "C" int gzb_longr_decode (gzb_engine *e, gzb_longr_vb *vbs, uint32_t n_vbs, uint32_t flags) { return longr_run (e, vbs, n_vbs, flags, false); }
