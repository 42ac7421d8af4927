// Do the warps of 1-warp CTAs spread over the 4 sub-partitions (schedulers) of an SM, or pile on one?
// Each warp runs an issue-bound loop (8 independent FMA chains: ~1 instruction per cycle for one warp alone on its scheduler).
//   A: 148 x k CTAs of 32 threads      B: 148 x k/4 CTAs of 128 threads      (same number of warps per SM)
// If A takes ~4x B, 1-warp CTAs share one scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin (float *out, int iters, unsigned *slots)
{
    float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    for (int i = 0; i < iters; i++) {
        a0 = fmaf (a0, 1.0001f, 0.5f); a1 = fmaf (a1, 1.0001f, 0.5f); a2 = fmaf (a2, 1.0001f, 0.5f); a3 = fmaf (a3, 1.0001f, 0.5f);
        a4 = fmaf (a4, 1.0001f, 0.5f); a5 = fmaf (a5, 1.0001f, 0.5f); a6 = fmaf (a6, 1.0001f, 0.5f); a7 = fmaf (a7, 1.0001f, 0.5f);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.f) out[0] = a0;
    if ((threadIdx.x & 31) == 0) {
        unsigned w, s; asm ("mov.u32 %0, %%warpid;" : "=r"(w)); asm ("mov.u32 %0, %%smid;" : "=r"(s));
        atomicAdd (&slots[(s * 4 + (w & 3)) % 1024], 1u);
    }
}
int main ()
{
    float *out; unsigned *slots; cudaMalloc (&out, 4); cudaMalloc (&slots, 4096);
    cudaEvent_t e0, e1; cudaEventCreate (&e0); cudaEventCreate (&e1);
    const int iters = 200000;
    for (int k : { 4, 8, 16 }) {
        for (int shape = 0; shape < 3; shape++) {
            const int threads = shape == 0 ? 32 : shape == 1 ? 64 : 128, grid = 148 * k * 32 / threads;
            cudaMemset (slots, 0, 4096);
            spin<<<grid, threads>>>(out, 1000, slots); cudaDeviceSynchronize ();
            cudaMemset (slots, 0, 4096);
            cudaEventRecord (e0); spin<<<grid, threads>>>(out, iters, slots); cudaEventRecord (e1); cudaDeviceSynchronize ();
            float ms; cudaEventElapsedTime (&ms, e0, e1);
            unsigned h[1024]; cudaMemcpy (h, slots, 4096, cudaMemcpyDeviceToHost);
            unsigned q[4] = { 0, 0, 0, 0 }; for (int i = 0; i < 148 * 4; i++) q[i & 3] += h[i];
            printf ("warps/SM %2d  CTA of %3d threads  grid %5d : %8.3f ms   warpid%%4 histogram %u %u %u %u\n", k, threads, grid, ms, q[0], q[1], q[2], q[3]);
        }
    }
    return 0;
}
