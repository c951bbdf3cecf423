// Microbenchmark + descriptor probe for tcgen05.mma kind::tf32 (M = 128) on sm_100a.
//
//  (1) rate : cycles per MMA for shared-memory operand layouts {no swizzle (the r01 layout), 32B, 64B, 128B swizzle} and
//             N in {128, 256}, 148 CTAs issuing back to back -> which layout lets an SS-mode MMA reach its 128x256x8 floor,
//             and the dense TF32 peak of the box (the denominator of the 3xTF32 ceiling).
//  (2) shift: does a swizzled K-major operand tolerate a start address advanced by an arbitrary number of ROWS (the
//             implicit-GEMM tap shift) -- with base_offset = 0 or (start >> 7) & 7 -- and by 32-byte K steps inside the
//             swizzle row?  D = A * B^T is compared with a host reference.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_mma tools/microbench_mma.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// layout_type: 0 none, 6 = 32B, 4 = 64B, 2 = 128B
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Args {
    int mode;        // 0 rate, 1 shift probe
    int layout;      // 0, 6, 4, 2
    int N;
    int nmma;
    int shift_rows;  // B start advanced by this many rows
    int kstep;       // B and A start advanced by kstep * 32 bytes inside the swizzle row
    int use_base_off;
    int vary;        // rate mode: rotate through `vary` different B start rows (like taps)
    int vstep;       // rows between those starts
    int kvary;       // rate mode, swizzled layouts: rotate the 32-byte K step inside the swizzle row over `kvary` values (same rows)
    int avary;       // rate mode: rotate through `avary` different (aligned) A tiles, 4 KB apart (like per-tap weight blocks)
    const float* A;  // [128][KW]   (KW = floats per row of the layout: 8, 16, 32; none: 8)
    const float* B;  // [ROWS_B][KW]
    float* D;        // [128][N]
    long long* cyc;  // per CTA
    int rows_b;
};

__device__ __forceinline__ uint32_t sw_off(uint32_t lin, int layout) {   // byte offset -> swizzled byte offset (absolute pattern)
    const int bits = layout == 2 ? 3 : layout == 4 ? 2 : layout == 6 ? 1 : 0;
    return lin ^ (((lin >> 7) & ((1u << bits) - 1u)) << 4);
}

__global__ void __launch_bounds__(192, 1) k_mma(const Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int SW = a.layout == 2 ? 128 : a.layout == 4 ? 64 : 32;   // row bytes (none: 2 slabs of 16 B rows)
    const int KW = SW / 4;
    // regions: A at 0 (128 rows), B at 32 KB (rows_b rows), barrier + tmem slot at the end
    unsigned char* smA = smem;
    unsigned char* smB = smem + 32 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    // ---- fill operands
    if (a.layout == 0) {
        // r01 layout: two 4-float slabs, rows 16 B apart; slab stride = rows * 16
        for (int i = tid; i < 128 * 8; i += blockDim.x) {
            const int r = i / 8, kk = i % 8;
            reinterpret_cast<float*>(smA + (kk / 4) * (128 * 16) + r * 16)[kk % 4] = a.A ? a.A[r * 8 + kk] : 0.f;
        }
        for (int i = tid; i < a.rows_b * 8; i += blockDim.x) {
            const int r = i / 8, kk = i % 8;
            reinterpret_cast<float*>(smB + (kk / 4) * (a.rows_b * 16) + r * 16)[kk % 4] = a.B ? a.B[r * 8 + kk] : 0.f;
        }
    } else {
        for (int i = tid; i < 128 * KW; i += blockDim.x) {
            const int r = i / KW, kk = i % KW;
            const uint32_t off = sw_off((uint32_t)(r * SW + (kk / 4) * 16), a.layout) + (kk % 4) * 4;
            *reinterpret_cast<float*>(smA + off) = a.A ? a.A[r * KW + kk] : 0.f;
        }
        for (int i = tid; i < a.rows_b * KW; i += blockDim.x) {
            const int r = i / KW, kk = i % KW;
            const uint32_t off = sw_off((uint32_t)(r * SW + (kk / 4) * 16), a.layout) + (kk % 4) * 4;
            *reinterpret_cast<float*>(smB + off) = a.B ? a.B[r * KW + kk] : 0.f;
        }
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t idesc = make_idesc(a.N);
    if (tid == 160) {
        const uint32_t abase = smem_u32(smA), bbase = smem_u32(smB);
        uint64_t adesc, bdesc0;
        uint32_t rowb;
        if (a.layout == 0) {
            adesc = make_desc(abase, 128 * 16, 128, 0, 0);
            bdesc0 = make_desc(bbase, a.rows_b * 16, 128, 0, 0);
            rowb = 16;
        } else {
            const uint32_t sa = abase + a.kstep * 32;
            adesc = make_desc(sa, 16, 8 * SW, a.layout, 0);
            bdesc0 = 0;
            rowb = SW;
        }
        // descriptors of one 8-MMA round are built BEFORE the timed loop and the round is fully unrolled: the issuing
        // thread's own index arithmetic must not be what is measured (a first version computed i % vary in the loop and
        // was issue-bound at ~237 cycles per MMA whatever the operands did)
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int sh = a.shift_rows + (a.vary > 1 ? (j % a.vary) * a.vstep : 0);
            ad[j] = adesc + (a.avary > 1 ? (uint64_t)(((j % a.avary) * 4096) >> 4) : 0);
            if (a.layout == 0) bd[j] = bdesc0 + (uint64_t)sh;
            else {
                const uint32_t sb = bbase + sh * rowb + (a.kvary > 1 ? (j % a.kvary) : a.kstep) * 32;
                bd[j] = make_desc(sb, 16, 8 * SW, a.layout, a.use_base_off ? ((sb >> 7) & 7) : 0);
                if (a.kvary > 1) ad[j] = make_desc(abase + (j % a.kvary) * 32, 16, 8 * SW, a.layout, 0);
            }
        }
        long long t0 = clock64();
        mma_tf32(tmem, ad[0], bd[0], idesc, 0u);
        for (int i = 8; i < a.nmma; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_tf32(tmem, ad[j], bd[j], idesc, 1u);
        }
        mma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
        long long t1 = clock64();
        if (a.cyc) a.cyc[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (a.mode == 1 && warp < 4 && a.D) {
        const int lane = tid & 31;
        for (int c = 0; c < a.N; c += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i) a.D[(size_t)(warp * 32 + lane) * a.N + c + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

static const char* lname(int l) { return l == 0 ? "none" : l == 6 ? "sw32" : l == 4 ? "sw64" : "sw128"; }

int main(int argc, char** argv) {
    const size_t SMEM = 201 * 1024;
    CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("SMs %d, max clock %.0f MHz\n", sms, khz / 1000.0);
    long long* dcyc;
    CK(cudaMalloc(&dcyc, sizeof(long long) * sms));
    // ------------------------------------------------------------ (1) rate
    printf("# rate: cycles per tcgen05.mma kind::tf32 M=128, K=8 (all %d SMs issuing, nmma=4096)\n", sms);
    struct Case { int shift, vary, vstep, avary, kvary; const char* what; };
    const Case cases[] = {
        {0, 1, 0, 1, 1, "fixed descriptors, aligned"},
        {3, 1, 0, 1, 1, "fixed descriptors, B start 3 rows off an 8-row group"},
        {0, 4, 8, 1, 1, "B rotates over 4 starts, 8 rows apart (aligned)"},
        {0, 4, 3, 1, 1, "B rotates over 4 starts, 3 rows apart (the conv tap pattern, dil 3)"},
        {0, 8, 1, 1, 1, "B rotates over 8 starts, 1 row apart (dil 1)"},
        {0, 1, 0, 4, 1, "A rotates over 4 aligned tiles, B fixed aligned"},
        {0, 4, 3, 4, 1, "A rotates (aligned), B rotates 3 rows apart"},
        {0, 1, 0, 1, 4, "A and B step through the 4 K offsets of one swizzle row (same rows; GEMM K loop)"},
        {0, 4, 3, 1, 4, "K offsets rotate AND B start rotates 3 rows apart"},
    };
    for (int layout : {0, 6, 4, 2})
        for (int N : {128, 256})
            for (const Case& cs : cases) {
                Args a; memset(&a, 0, sizeof(a));
                a.mode = 0; a.layout = layout; a.N = N; a.nmma = 4096; a.vary = cs.vary; a.vstep = cs.vstep; a.avary = cs.avary; a.kvary = cs.kvary;
                a.shift_rows = cs.shift; a.rows_b = 320; a.cyc = dcyc;
                a.use_base_off = 0;
                if (layout != 0 && layout != 6 && cs.avary > 1) continue;   // wide-row A tiles do not fit the 32 KB A region 4x
                if (cs.kvary > 1 && layout != 2) continue;                  // K offsets inside a row: 128-byte rows only
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                k_mma<<<sms, 192, SMEM>>>(a);   // warm-up
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                k_mma<<<sms, 192, SMEM>>>(a);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                std::vector<long long> h(sms);
                CK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
                double avg = 0; for (auto v : h) avg += (double)v; avg /= sms;
                const double cpm = avg / a.nmma;
                printf("layout %-5s N %3d : %7.1f cyc/MMA -> %7.1f dense TF32 TFLOP/s at %.0f MHz | %s\n", lname(layout), N, cpm,
                       2.0 * 128 * N * 8 / cpm * sms * (khz * 1e3) / 1e12, khz / 1000.0, cs.what);
            }
    // ------------------------------------------------------------ (2) shift probe
    printf("# shift probe: D = A * B[shift:]^T, max |err| vs host (tf32-exact inputs)\n");
    for (int layout : {0, 6, 4, 2}) {
        const int SW = layout == 2 ? 128 : layout == 4 ? 64 : 32, KW = SW / 4, N = 256, RB = 320;
        std::vector<float> A(128 * KW), B((size_t)RB * KW);
        srand(1);
        auto rnd = [] { return (float)((rand() % 255) - 127) / 16.f; };   // exactly representable in tf32
        for (auto& v : A) v = rnd();
        for (auto& v : B) v = rnd();
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        for (int kstep = 0; kstep < KW / 8; ++kstep)
            for (int shift : {0, 1, 2, 3, 5, 8, 11, 50})
                for (int ubo = 0; ubo < (layout == 0 ? 1 : 2); ++ubo) {
                    Args a; memset(&a, 0, sizeof(a));
                    a.mode = 1; a.layout = layout; a.N = N; a.nmma = 1; a.shift_rows = shift; a.kstep = kstep; a.use_base_off = ubo;
                    a.vary = 1; a.A = dA; a.B = dB; a.D = dD; a.rows_b = RB;
                    CK(cudaMemset(dD, 0, 128 * N * 4));
                    k_mma<<<1, 192, SMEM>>>(a);
                    CK(cudaDeviceSynchronize());
                    std::vector<float> D(128 * N);
                    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                    double maxerr = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < N; ++n) {
                            double ref = 0;
                            for (int kk = 0; kk < 8; ++kk) ref += (double)A[m * KW + kstep * 8 + kk] * (double)B[(size_t)(n + shift) * KW + kstep * 8 + kk];
                            maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
                        }
                    printf("layout %-5s kstep %d shift %2d base_off %s : max err %.3g %s\n", lname(layout), kstep, shift,
                           layout == 0 ? "-" : (ubo ? "(start>>7)&7" : "0"), maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
                }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
