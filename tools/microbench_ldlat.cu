// Micro-benchmark: latency of a batch of N independent coalesced warp loads per thread, 64/128 threads per CTA,
// 1-2 CTAs/SM, strided rows like the conv producer (row stride = 9600 floats).
#include <cuda_runtime.h>
#include <stdio.h>
template <int NL>
__global__ void k(const float* x, float* out, long long* tm, int T, int cs) {
    const int tid = threadIdx.x;
    long long b0 = (long long)blockIdx.x * 1000003 % (32 * 128);   // scatter blocks over rows
    float v[NL];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
#pragma unroll
    for (int i = 0; i < NL; ++i) v[i] = __ldg(x + (b0 + (i & 7)) * cs + (blockIdx.x % 30) * 256 + (i >> 3) * 64 + tid);
    float s = 0;
#pragma unroll
    for (int i = 0; i < NL; ++i) s += v[i];
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (s == 1234.5f) out[0] = s;
    if (tid == 0) tm[blockIdx.x] = (long long)(t1 - t0);
}
int main() {
    const int T = 9600, rows = 32 * 128 + 8;
    float* x; float* out; long long* tm;
    cudaMalloc(&x, (size_t)rows * T * 4); cudaMemset(x, 0, (size_t)rows * T * 4);
    cudaMalloc(&out, 4); cudaMalloc(&tm, 8 * 4096);
    for (int blocks : {148, 296, 592, 1184}) {
        for (int rep = 0; rep < 2; ++rep) {
            k<40><<<blocks, 64>>>(x, out, tm, T, T);
            cudaDeviceSynchronize();
        }
        long long h[4096]; cudaMemcpy(h, tm, 8 * blocks, cudaMemcpyDeviceToHost);
        double avg = 0, mx = 0; for (int i = 0; i < blocks; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
        printf("blocks=%d threads=64 loads/thread=40: avg %.2f us max %.2f us\n", blocks, avg / blocks / 1e3, mx / 1e3);
    }
    return 0;
}
