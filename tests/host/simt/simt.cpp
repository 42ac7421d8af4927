// tests/host/simt/simt.cpp — the lock-step SIMT emulator behind tests/host/simt/cuda_runtime.h.  TEST INFRASTRUCTURE ONLY.
//
// A block's threads are fibres (own stacks, hand-written x86-64 context switch) run by one OS thread; a fibre runs until it
// reaches a rendez-vous (warp collective or __syncthreads) that is not complete yet, then the scheduler moves on to the next
// lane.  The last lane to arrive computes every lane's result and releases the others.  A full pass over the block in which
// no fibre made progress is a deadlock: the run aborts with the position of every waiting lane — on the GPU the same program
// would hang (or, for a shuffle, read garbage).  Set SIMT_TRACE=1 to print every launch.
#include "cuda_runtime.h"
#include <stdio.h>
#include <time.h>
#include <sys/mman.h>
#include <vector>
#include <string>

extern "C" void simt_switch (void **save_sp, void *load_sp);
asm (R"(
    .text
    .globl simt_switch
    .type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size simt_switch,.-simt_switch
    .section .note.GNU-stack,"",@progbits
    .text
)");

namespace simt {

thread_local Thread *cur = nullptr;

enum State { RUN, WAIT_WARP, WAIT_BLOCK, DONE };

struct Slot {                                                               // one pending collective; disjoint masks may be pending side by side
    uint32_t arrived = 0, mask = 0;
    Kind kind = K_SYNCWARP; const char *file = nullptr; int line = 0;
    uint64_t in[32], out[32]; int arg[32], width[32];
};
struct Warp { Slot slot[32]; uint32_t exited = 0, exist = 0xffffffffu; };   // pending collectives, keyed by (site, kind, mask): a lane can belong to one only

struct Fibre {
    Thread th;
    void *sp = nullptr; char *stack = nullptr;
    State state = RUN;
    int lane = 0, warp = 0;
    const char *file = nullptr; int line = 0;                                // where it waits
};

struct Block {
    std::vector<Fibre> f;
    std::vector<Warp> w;
    const std::function<void ()> *body = nullptr;
    void *sched_sp = nullptr;
    Fibre *running = nullptr;
    uint32_t n_done = 0, n_at_barrier = 0, barrier_gen = 0;
    bool progress = false;
    int or_acc[3] = { 0, 0, 0 }; uint32_t or_gen = 0;                        // __syncthreads_or: three slots in rotation, see syncthreads_or
    const char *bar_file = nullptr; int bar_line = 0;                        // where the pending barrier's first thread waits
    uint32_t sched_rand = 12345;
    std::vector<uint8_t> dyn;
};

static thread_local Block *blk = nullptr;
static const size_t STACK = 256 * 1024;

static thread_local std::vector<char *> *stack_pool = nullptr;

static char *stack_get ()
{
    if (!stack_pool) stack_pool = new std::vector<char *>;
    if (!stack_pool->empty ()) { char *s = stack_pool->back (); stack_pool->pop_back (); return s; }
    void *p = mmap (nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { fprintf (stderr, "simt: cannot map a fibre stack\n"); abort (); }
    return (char *)p;
}

static void yield_to_scheduler ()
{
    Block *b = blk; Fibre *me = b->running;
    simt_switch (&me->sp, b->sched_sp);
    cur = &me->th;                                                          // (back on this fibre)
}

static bool try_complete (Block *b, int warp, Slot &w);

static void fibre_main ()
{
    Block *b = blk; Fibre *me = b->running;
    cur = &me->th;
    (*b->body) ();
    me->state = DONE; b->n_done++; b->progress = true;
    Warp &w = b->w[me->warp];
    w.exited |= 1u << me->lane;                                             // lanes that have exited are not waited for
    for (int i = 0; i < 32; i++) if (w.slot[i].arrived) try_complete (b, me->warp, w.slot[i]);
    if (b->n_at_barrier && b->n_at_barrier + b->n_done == b->f.size ()) {
        b->n_at_barrier = 0;
        for (Fibre &o : b->f) if (o.state == WAIT_BLOCK) o.state = RUN;
    }
    simt_switch (&me->sp, b->sched_sp);
    abort ();                                                               // a finished fibre is never resumed
}

static void die (Block *b, const char *what)
{
    fprintf (stderr, "simt: %s (block %u,%u)\n", what, b->f[0].th.bid.x, b->f[0].th.bid.y);
    static const char *st[] = { "running", "waiting at a warp collective", "waiting at __syncthreads", "exited" };
    std::string last; int first = -1; size_t n = b->f.size ();
    for (size_t i = 0; i <= n; i++) {                                       // compress equal neighbours into ranges
        char buf[512] = "";
        if (i < n) { Fibre &f = b->f[i]; snprintf (buf, sizeof buf, "%s%s%s:%d", st[f.state], f.state == WAIT_WARP || f.state == WAIT_BLOCK ? " " : "",
                                                    f.state == WAIT_WARP || f.state == WAIT_BLOCK ? f.file : "", f.state == WAIT_WARP || f.state == WAIT_BLOCK ? f.line : 0); }
        if (i == n || last != buf) { if (first >= 0) fprintf (stderr, "   threads %d..%d: %s\n", first, (int)i - 1, last.c_str ()); first = (int)i; last = buf; }
    }
    abort ();
}

static uint64_t lane_result (Slot &w, int lane)
{
    const uint32_t m = w.mask;
    switch (w.kind) {
        case K_SHFL_IDX: case K_SHFL_UP: case K_SHFL_DOWN: case K_SHFL_XOR: {
            const int width = w.width[lane], base = lane & ~(width - 1), a = w.arg[lane];
            int src;
            if (w.kind == K_SHFL_IDX) src = base + (a & (width - 1));
            else if (w.kind == K_SHFL_UP) { src = lane - a; if (src < base) src = lane; }
            else if (w.kind == K_SHFL_DOWN) { src = lane + a; if (src >= base + width) src = lane; }
            else { src = lane ^ a; if (src >= base + width) src = lane; }
            if (!(m >> src & 1)) return 0xDEADDEADDEADDEADull;              // reading a lane that is not in the mask: undefined on the GPU
            return w.in[src];
        }
        case K_BALLOT: { uint32_t r = 0; for (int l = 0; l < 32; l++) if ((m >> l & 1) && w.in[l]) r |= 1u << l; return r; }
        case K_MATCH_ANY: { uint32_t r = 0; for (int l = 0; l < 32; l++) if ((m >> l & 1) && (w.arrived >> l & 1) && w.in[l] == w.in[lane]) r |= 1u << l; return r; }
        case K_REDUCE_ADD: { uint32_t r = 0; for (int l = 0; l < 32; l++) if (m >> l & 1) r += (uint32_t)w.in[l]; return r; }
        case K_REDUCE_OR:  { uint32_t r = 0; for (int l = 0; l < 32; l++) if (m >> l & 1) r |= (uint32_t)w.in[l]; return r; }
        case K_REDUCE_AND: { uint32_t r = ~0u; for (int l = 0; l < 32; l++) if (m >> l & 1) r &= (uint32_t)w.in[l]; return r; }
        case K_REDUCE_MIN: { uint32_t r = ~0u; for (int l = 0; l < 32; l++) if (m >> l & 1) r = std::min (r, (uint32_t)w.in[l]); return r; }
        case K_REDUCE_MAX: { uint32_t r = 0; for (int l = 0; l < 32; l++) if (m >> l & 1) r = std::max (r, (uint32_t)w.in[l]); return r; }
        default: return 0;
    }
}

static bool try_complete (Block *b, int warp, Slot &w)
{
    Warp &W = b->w[warp];
    const uint32_t need = w.mask & W.exist & ~W.exited;
    if ((w.arrived & need) != need) return false;
    const uint32_t who = w.arrived;
    for (int l = 0; l < 32; l++) if (who >> l & 1) w.out[l] = lane_result (w, l);
    w.arrived = 0;
    for (int l = 0; l < 32; l++) if (who >> l & 1) { Fibre &o = b->f[(size_t)32 * warp + l]; if (o.state == WAIT_WARP) o.state = RUN; }
    return true;
}

uint64_t collective (const char *file, int line, Kind kind, uint32_t mask, uint64_t v, int arg, int width)
{
    Block *b = blk; Fibre *me = b->running; Warp &W = b->w[me->warp];
    const int lane = me->lane;
    me->file = file; me->line = line;
    if (!(mask >> lane & 1)) { me->state = WAIT_WARP; die (b, "a lane executes a warp collective with a mask that does not name it"); }
    // join the pending rendez-vous of this very collective, or open one.  Lanes of one mask that wait at DIFFERENT collectives
    // never complete either: the deadlock report names both sites.
    Slot *w = nullptr, *spare = nullptr;
    for (int i = 0; i < 32; i++) {
        Slot &o = W.slot[i];
        if (!o.arrived) { if (!spare) spare = &o; continue; }
        if (o.kind == kind && o.mask == mask && o.line == line && !strcmp (o.file, file)) { w = &o; break; }
    }
    if (!w) { w = spare; w->kind = kind; w->mask = mask; w->file = file; w->line = line; }
    w->in[lane] = v; w->arg[lane] = arg; w->width[lane] = width;
    w->arrived |= 1u << lane;
    b->progress = true;
    if (try_complete (b, me->warp, *w)) return w->out[lane];
    me->state = WAIT_WARP;
    while (me->state == WAIT_WARP) yield_to_scheduler ();
    return w->out[lane];
}

void syncthreads (const char *file, int line)
{
    Block *b = blk; Fibre *me = b->running;
    me->file = file; me->line = line;
    b->progress = true;
    if (b->n_at_barrier && (b->bar_line != line || strcmp (b->bar_file, file))) {   // undefined in CUDA: the hardware counts arrivals, whatever the site
        me->state = WAIT_BLOCK;
        die (b, "threads of one block meet at DIFFERENT __syncthreads (a barrier in divergent code)");
    }
    b->bar_file = file; b->bar_line = line;
    if (b->n_at_barrier + 1 + b->n_done == b->f.size ()) {                   // exited threads count as arrived
        b->n_at_barrier = 0;
        for (Fibre &o : b->f) if (o.state == WAIT_BLOCK) o.state = RUN;
        return;
    }
    b->n_at_barrier++;
    me->state = WAIT_BLOCK;
    while (me->state == WAIT_BLOCK) yield_to_scheduler ();
}

// __syncthreads_or: every thread ORs its predicate into the slot of this call, the barrier, everybody reads it.  The slot two
// calls ahead is cleared by the readers: nobody writes to it before the NEXT call's barrier has released, which needs every
// thread to have left this call.
int syncthreads_or (const char *file, int line, int pred)
{
    Block *b = blk; Fibre *me = b->running;
    const uint32_t gen = me->th.or_gen++ % 3;
    b->or_acc[gen] |= pred;
    syncthreads (file, line);
    const int r = b->or_acc[gen];
    b->or_acc[(gen + 2) % 3] = 0;
    return r;
}

void *dyn_smem () { return blk->dyn.data (); }

static const int lane_order = [] { const char *v = getenv ("SIMT_LANE_ORDER"); return !v ? 0 : !strcmp (v, "desc") ? 1 : !strcmp (v, "random") ? 2 : 0; } ();

static void run_block (Block &b, dim3 grid, dim3 block, uint3 bid)
{
    const uint32_t n = block.x * block.y * block.z;
    b.n_done = b.n_at_barrier = 0; b.or_acc[0] = b.or_acc[1] = b.or_acc[2] = 0;
    for (uint32_t t = 0; t < n; t++) {
        Fibre &f = b.f[t];
        f.th.tid = uint3{ t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
        f.th.bid = bid; f.th.bdim = block; f.th.gdim = grid; f.th.or_gen = 0;
        f.lane = t & 31; f.warp = t >> 5; f.state = RUN;
        // a fresh stack: six callee-saved registers, then the address simt_switch returns to
        uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
        void **sp = (void **)(top - 8);                                     // fibre_main sees the alignment of a called function
        *--sp = (void *)fibre_main;
        for (int i = 0; i < 6; i++) *--sp = nullptr;
        f.sp = sp;
    }
    for (uint32_t wi = 0; wi < b.w.size (); wi++) {
        Warp &w = b.w[wi];
        for (Slot &sl : w.slot) sl.arrived = 0;
        const uint32_t in_warp = std::min (32u, n - 32u * wi);
        w.exited = 0; w.exist = in_warp == 32 ? 0xffffffffu : ((1u << in_warp) - 1);      // lanes that do not exist are never waited for
    }
    blk = &b;
    while (b.n_done < n) {
        b.progress = false;
        for (uint32_t wi = 0; wi < b.w.size (); wi++) {                     // a warp keeps the OS thread while any of its lanes can run
            bool again = true;
            while (again) {
                again = false;
                const uint32_t rot = (b.sched_rand = b.sched_rand * 1103515245u + 12345u) >> 16;
                for (uint32_t k = 0; k < 32; k++) {
                    // SIMT_LANE_ORDER=desc / random: lanes are not taken in ascending order.  Code that is correct only because lane 0
                    // (or lane 31) happens to run first between two rendez-vous points shows up under the other orders.
                    const uint32_t l = lane_order == 1 ? 31 - k : lane_order == 2 ? (k * 13 + rot) % 32 : k;   // (13 is odd: a permutation)
                    if (wi * 32 + l >= n) continue;
                    Fibre &f = b.f[wi * 32 + l];
                    if (f.state != RUN) continue;
                    b.running = &f;
                    simt_switch (&b.sched_sp, f.sp);
                    again = true;
                }
            }
        }
        if (!b.progress) die (&b, "deadlock: a rendez-vous can never complete (a collective or barrier inside divergent code?)");
    }
    blk = nullptr; cur = nullptr;
}

void launch (dim3 grid, dim3 block, size_t dyn, const std::function<void ()> &body)
{
    static const bool trace = getenv ("SIMT_TRACE") != nullptr;
    const uint32_t n = block.x * block.y * block.z;
    if (!n || n > 1024 || !grid.x || !grid.y || !grid.z) { fprintf (stderr, "simt: invalid launch configuration grid %u,%u,%u block %u\n", grid.x, grid.y, grid.z, n); abort (); }
    if (blk) { fprintf (stderr, "simt: launch from inside a kernel\n"); abort (); }
    if (trace) fprintf (stderr, "simt: launch grid %u,%u,%u block %u smem %zu\n", grid.x, grid.y, grid.z, n, dyn);
    Block b;
    b.f.resize (n); b.w.resize ((n + 31) / 32); b.body = &body; b.dyn.resize (dyn + 16);
    for (Fibre &f : b.f) f.stack = stack_get ();
    for (uint32_t z = 0; z < grid.z; z++) for (uint32_t y = 0; y < grid.y; y++) for (uint32_t x = 0; x < grid.x; x++)
        run_block (b, grid, block, uint3{ x, y, z });
    for (Fibre &f : b.f) stack_pool->push_back (f.stack);
}

double now_ms () { timespec t; clock_gettime (CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }

void *dev_alloc (size_t n)
{
    if (!n) n = 1;
    void *p = mmap (nullptr, n + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);   // lazily committed, like an arena
    if (p == MAP_FAILED) return nullptr;
    *(size_t *)p = n + 4096;
    return (char *)p + 4096;
}

void dev_free (void *p) { if (p) { char *q = (char *)p - 4096; munmap (q, *(size_t *)q); } }

} // namespace simt
