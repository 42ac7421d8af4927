// (a) up to which size does a pinned H2D copy on another stream NOT wait for a bulk upload in progress?
// (b) does feeding the bulk upload one piece at a time (the engine idles between pieces) let other copies in, and what does it cost?
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>
static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }
int main ()
{
    const size_t N = 16ull << 30;
    char *h, *d, *hs, *ds;
    cudaMallocHost (&h, N); cudaMalloc (&d, N); cudaMallocHost (&hs, 8 << 20); cudaMalloc (&ds, 8 << 20);
    cudaStream_t sa, sb; cudaStreamCreateWithFlags (&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags (&sb, cudaStreamNonBlocking);
    cudaEvent_t ev[8]; for (auto &e : ev) cudaEventCreateWithFlags (&e, cudaEventDisableTiming);
    struct Cfg { size_t piece; int depth; } cfgs[] = { { 8u << 20, 4 }, { 8u << 20, 1 }, { 32u << 20, 1 }, { 64u << 20, 1 }, { 32u << 20, 2 } };
    const size_t sizes[] = { 16 << 10, 32 << 10, 48 << 10, 64 << 10, 96 << 10, 128 << 10, 1 << 20 };
    for (auto c : cfgs) {
        for (size_t sz : sizes) {
            std::atomic<bool> done (false); double t_up = 0;
            std::thread up ([&] { double a = now (); size_t i = 0; for (size_t o = 0; o < N; o += c.piece, i++) { if ((int)i >= c.depth) cudaEventSynchronize (ev[i % c.depth]); cudaMemcpyAsync (d + o, h + o, c.piece, cudaMemcpyHostToDevice, sa); cudaEventRecord (ev[i % c.depth], sa); }
                                cudaStreamSynchronize (sa); t_up = now () - a; done = true; });
            std::this_thread::sleep_for (std::chrono::milliseconds (20));
            double worst = 0, sum = 0; int n = 0;
            while (!done && n < 300) { double a = now (); cudaMemcpyAsync (ds, hs, sz, cudaMemcpyHostToDevice, sb); cudaStreamSynchronize (sb); double t = now () - a; sum += t; if (t > worst) worst = t; n++; }
            up.join ();
            printf ("bulk: %2zu MB pieces, depth %d -> %.1f GB/s | beside it, H2D of %4zu KB: mean %8.3f ms, worst %8.3f ms (%d copies)\n", c.piece >> 20, c.depth, N / t_up / 1e9, sz >> 10, 1e3 * sum / n, 1e3 * worst, n);
        }
    }
    return 0;
}
