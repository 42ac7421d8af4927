// How long does a small H2D copy on its own stream take while a bulk H2D upload is in progress on another stream?
//   bulk modes: one piece | all pieces queued at once | fed with a bounded depth (4 x 8 MB)
//   small copy: 4 KB from pinned memory, 4 KB from pageable memory, and a kernel launch, each followed by a stream synchronize
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <atomic>
#include <vector>
#include <cuda_runtime.h>
__global__ void nop (int *p) { if (p && threadIdx.x == 12345) *p = 1; }
static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }
int main ()
{
    const size_t N = 8ull << 30, PIECE = 8ull << 20;
    char *h, *d, *hs, *ds; cudaMallocHost (&h, N); cudaMalloc (&d, N); cudaMallocHost (&hs, 4096); cudaMalloc (&ds, 1 << 20);
    char *pg = (char *)malloc (4096);
    cudaStream_t sa, sb; cudaStreamCreateWithFlags (&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags (&sb, cudaStreamNonBlocking);
    cudaEvent_t ev[4]; for (auto &e : ev) cudaEventCreateWithFlags (&e, cudaEventDisableTiming);
    for (int mode = 0; mode < 3; mode++) {
        std::atomic<bool> done (false);
        double t0 = now ();
        std::thread up ([&] {
            if (mode == 0) cudaMemcpyAsync (d, h, N, cudaMemcpyHostToDevice, sa);
            else if (mode == 1) for (size_t o = 0; o < N; o += PIECE) cudaMemcpyAsync (d + o, h + o, PIECE, cudaMemcpyHostToDevice, sa);
            else { size_t i = 0; for (size_t o = 0; o < N; o += PIECE, i++) { if (i >= 4) cudaEventSynchronize (ev[i % 4]); cudaMemcpyAsync (d + o, h + o, PIECE, cudaMemcpyHostToDevice, sa); cudaEventRecord (ev[i % 4], sa); } }
            cudaStreamSynchronize (sa); done = true; });
        std::this_thread::sleep_for (std::chrono::milliseconds (20));
        double worst[3] = { 0, 0, 0 }, sum[3] = { 0, 0, 0 }; int n = 0;
        while (!done) {
            double a = now (); cudaMemcpyAsync (ds, hs, 4096, cudaMemcpyHostToDevice, sb); cudaStreamSynchronize (sb);
            double b = now (); cudaMemcpyAsync (ds, pg, 4096, cudaMemcpyHostToDevice, sb); cudaStreamSynchronize (sb);
            double c = now (); nop<<<1, 32, 0, sb>>>(nullptr); cudaStreamSynchronize (sb);
            double e = now ();
            const double v[3] = { b - a, c - b, e - c };
            for (int k = 0; k < 3; k++) { sum[k] += v[k]; if (v[k] > worst[k]) worst[k] = v[k]; }
            n++;
        }
        up.join ();
        printf ("bulk mode %d (%s): upload %.1f ms (%.1f GB/s); beside it %d rounds: pinned 4 KB mean %.3f worst %.3f ms | pageable 4 KB mean %.3f worst %.3f ms | kernel mean %.3f worst %.3f ms\n",
                mode, mode == 0 ? "one piece" : mode == 1 ? "all pieces queued" : "fed, depth 4", 1e3 * (now () - t0), N / (now () - t0) / 1e9, n,
                1e3 * sum[0] / n, 1e3 * worst[0], 1e3 * sum[1] / n, 1e3 * worst[1], 1e3 * sum[2] / n, 1e3 * worst[2]);
    }
    return 0;
}
