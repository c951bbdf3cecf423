// Microbenchmark: FP32 FMA-pipe peak on this part with scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fma microbench_fma.cu
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    } else {
        u64 p[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        u64 aa, bb;
        asm("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
        asm("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(b));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(aa), "l"(bb));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int blocks) {
    float* out;
    cudaMalloc(&out, blocks * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 100, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double flops = 2.0 * 32 * (double)iters * blocks * 256;
    printf("%s blocks=%d: %.3f ms  %.2f TFLOP/s\n", name, blocks, best, flops / best / 1e9);
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs=%d\n", sms);
    for (int m : {2, 4, 8}) { run<0>("FFMA ", sms * m); run<1>("FFMA2", sms * m); }
    return 0;
}
