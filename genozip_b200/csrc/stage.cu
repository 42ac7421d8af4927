// stage.cu — staging of VBlock text and sections between host and device while kernels of another batch run
// (gzb_vb_stage; the reference moves vb->txt_data / vb->z_data between its I/O and compute threads the same way,
// src/dispatcher.c).
//
// The entropy chains are latency-bound and leave the copy engines idle; the transfers are bandwidth-bound and need no SM.  But a
// copy engine does not share itself fairly between streams: as long as one stream has another copy queued behind the one in
// flight, the engine stays with it, and a copy of 32 KB or more on any other stream waits until that stream runs dry — the few-MB
// descriptor uploads of the batch calls that run meanwhile waited for the WHOLE 10 GB text upload, and their kernels with them
// (measured on B200, tools/probes/copy_probe3.cu: 298 ms for a 1 MB copy beside a 16 GB upload queued in 8 MB pieces four deep;
// copies of 16 KB and less go through the push buffer and are not affected).  So a transfer is fed to its stream ONE piece at a
// time by a feeder thread: the engine idles for a few microseconds between pieces and takes whatever else is waiting (0.6 ms
// worst case for the others, 54.5 GB/s for the bulk transfer: nothing lost against one big copy).  One feeder per direction
// (the link is full duplex), each in order.
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <cuda_runtime.h>
#include "../../include/gzb200.h"
#include "engine.h"

namespace {

constexpr uint64_t PIECE = 32ull << 20;     // bytes per cudaMemcpyAsync
constexpr int      DEPTH = 1;               // pieces queued on the stream at any time (more than one starves every other stream's copies)

struct Job { uint8_t *dst; const uint8_t *src; uint64_t bytes; };

struct Feeder {
    int device; cudaStream_t stream; cudaMemcpyKind kind;
    std::thread th; std::mutex m; std::condition_variable cv_work, cv_done;
    std::deque<Job> q; uint64_t submitted = 0, completed = 0; bool stop = false; cudaError_t err = cudaSuccess;
    cudaEvent_t ev[DEPTH]; 

    void run ()
    {
        cudaSetDevice (device);
        for (auto &e : ev) cudaEventCreateWithFlags (&e, cudaEventDisableTiming);   // (spin-wait on purpose: with cudaEventBlockingSync the wake-up per piece costs a third of the link — 54 -> 35 GB/s measured)
        uint64_t issued = 0;                                                // pieces issued so far; piece i uses event i % DEPTH
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk (m);
                cv_work.wait (lk, [&] { return stop || !q.empty (); });
                if (q.empty ()) break;
                j = q.front (); q.pop_front ();
            }
            for (uint64_t o = 0; o < j.bytes; o += PIECE, issued++) {
                if (issued >= DEPTH) cudaEventSynchronize (ev[issued % DEPTH]);          // the piece DEPTH back is through
                const uint64_t n = j.bytes - o < PIECE ? j.bytes - o : PIECE;
                const cudaError_t r = cudaMemcpyAsync (j.dst + o, j.src + o, n, kind, stream);
                if (r != cudaSuccess && err == cudaSuccess) err = r;
                cudaEventRecord (ev[issued % DEPTH], stream);
            }
            bool last;
            { std::lock_guard<std::mutex> lk (m); last = q.empty (); }
            if (last) {                                                      // nothing else waiting: see this job's tail through, then report
                const cudaError_t r = cudaStreamSynchronize (stream);
                if (r != cudaSuccess && err == cudaSuccess) err = r;
            }
            {
                std::lock_guard<std::mutex> lk (m);
                completed++;                                                 // (a job followed by another is reported when its successor is: the waiters want "all so far")
                if (!last) continue;
            }
            cv_done.notify_all ();
        }
        for (auto &e : ev) cudaEventDestroy (e);
    }

    void start () { th = std::thread ([this] { run (); }); }
    void push (const Job &j) { { std::lock_guard<std::mutex> lk (m); q.push_back (j); submitted++; } cv_work.notify_one (); }
    cudaError_t wait ()
    {
        std::unique_lock<std::mutex> lk (m);
        cv_done.wait (lk, [&] { return completed == submitted && q.empty (); });
        const cudaError_t r = err; err = cudaSuccess;
        return r;
    }
    void finish () { { std::lock_guard<std::mutex> lk (m); stop = true; } cv_work.notify_one (); if (th.joinable ()) th.join (); }
};

struct Stager { Feeder up, down; };

Stager *stager_of (gzb_engine *e)
{
    if (!e->stager) {
        Stager *s = new Stager ();
        s->up.device = s->down.device = e->device;
        s->up.stream = e->stream_copy;    s->up.kind = cudaMemcpyHostToDevice;
        s->down.stream = e->stream_copy2; s->down.kind = cudaMemcpyDeviceToHost;
        s->up.start (); s->down.start ();
        e->stager = s;
        e->stager_free = [] (void *p) { Stager *t = reinterpret_cast<Stager *>(p); t->up.finish (); t->down.finish (); delete t; };
    }
    return reinterpret_cast<Stager *>(e->stager);
}

} // namespace

extern "C" int gzb_stage_upload (gzb_engine *e, void *dst_device, const void *src_host, uint64_t bytes)
{
    if (!e || (bytes && (!dst_device || !src_host))) return GZB_E_BADARG;
    if (bytes) stager_of (e)->up.push (Job { (uint8_t *)dst_device, (const uint8_t *)src_host, bytes });
    return GZB_OK;
}

extern "C" int gzb_stage_fetch (gzb_engine *e, void *dst_host, const void *src_device, uint64_t bytes)
{
    if (!e || (bytes && (!dst_host || !src_device))) return GZB_E_BADARG;
    if (bytes) stager_of (e)->down.push (Job { (uint8_t *)dst_host, (const uint8_t *)src_device, bytes });
    return GZB_OK;
}

extern "C" int gzb_stage_wait (gzb_engine *e, int which)
{
    if (!e) return GZB_E_BADARG;
    if (!e->stager) return GZB_OK;
    Stager *s = stager_of (e);
    cudaError_t r = cudaSuccess;
    if (which != GZB_STAGE_FETCHES) { const cudaError_t x = s->up.wait (); if (x != cudaSuccess) r = x; }
    if (which != GZB_STAGE_UPLOADS) { const cudaError_t x = s->down.wait (); if (x != cudaSuccess) r = x; }
    if (r != cudaSuccess) { e->err = std::string ("staged transfer: ") + cudaGetErrorString (r); return GZB_E_CUDA; }
    return GZB_OK;
}
