// Which operations of a batch call wait for a bulk upload that is in progress on another stream (fed in 8 MB pieces, depth 4)?
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>
__global__ void reader (const uint4 *p, size_t n16, unsigned long long *out)          // bandwidth kernel: reads n16 x 16 bytes
{
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) { const uint4 v = p[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 0x123456789ull) *out = s;
}
__global__ void smem_kernel (unsigned long long *out) { __shared__ unsigned sm[9888]; sm[threadIdx.x] = threadIdx.x; __syncthreads (); if (sm[(threadIdx.x + 1) & 255] == 77777) *out = 1; }
static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }
int main ()
{
    const size_t N = 16ull << 30, PIECE = 8ull << 20, R = 10ull << 30;
    char *h, *d, *big, *hs, *ds; unsigned long long *out;
    cudaMallocHost (&h, N); cudaMalloc (&d, N); cudaMalloc (&big, R); cudaMallocHost (&hs, 8 << 20); cudaMalloc (&ds, 8 << 20); cudaMalloc (&out, 8);
    cudaStream_t sa, sb; cudaStreamCreateWithFlags (&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags (&sb, cudaStreamNonBlocking);
    cudaEvent_t ev[4]; for (auto &e : ev) cudaEventCreateWithFlags (&e, cudaEventDisableTiming);
    auto ops = [&] (int k) {
        switch (k) {
            case 0: cudaMemsetAsync (ds, 0, 4 << 20, sb); break;
            case 1: cudaMemcpyAsync (hs, ds, 4 << 20, cudaMemcpyDeviceToHost, sb); break;
            case 2: cudaMemcpyAsync (ds, hs, 1 << 20, cudaMemcpyHostToDevice, sb); break;
            case 3: reader<<<148 * 8, 256, 0, sb>>>((const uint4 *)big, R / 16, out); break;
            case 4: reader<<<dim3 (4096, 768), 256, 0, sb>>>((const uint4 *)big, R / 16 / 64, out); break;
            case 5: smem_kernel<<<23000, 256, 0, sb>>>(out); break;
        }
        cudaStreamSynchronize (sb);
    };
    const char *names[6] = { "memset 4 MB", "D2H 4 MB pinned", "H2D 1 MB pinned", "read 10 GB, 1184 CTAs", "big grid 4096x768 CTAs", "23000 CTAs x 39 KB smem" };
    for (int k = 0; k < 6; k++) ops (k);
    double alone[6]; for (int k = 0; k < 6; k++) { double a = now (); ops (k); alone[k] = now () - a; }
    for (int k = 0; k < 6; k++) {
        std::atomic<bool> done (false);
        std::thread up ([&] { size_t i = 0; for (size_t o = 0; o < N; o += PIECE, i++) { if (i >= 4) cudaEventSynchronize (ev[i % 4]); cudaMemcpyAsync (d + o, h + o, PIECE, cudaMemcpyHostToDevice, sa); cudaEventRecord (ev[i % 4], sa); }
                            cudaStreamSynchronize (sa); done = true; });
        std::this_thread::sleep_for (std::chrono::milliseconds (20));
        double a = now (); ops (k); double first = now () - a; int n = 1; double sum = first;
        while (!done && n < 200) { a = now (); ops (k); sum += now () - a; n++; }
        up.join ();
        printf ("%-26s alone %8.3f ms | beside the upload: first %8.3f ms, mean of %3d %8.3f ms\n", names[k], 1e3 * alone[k], 1e3 * first, n, 1e3 * sum / n);
    }
    return 0;
}
