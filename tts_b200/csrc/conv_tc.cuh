// Shared pieces of the tcgen05 (5th-gen tensor core) 3xTF32 implicit-GEMM conv kernels: barrier / descriptor / TMEM helpers and
// the operand-layout description.  The kernel that runs is conv_tc3.cuh; the first two kernel generations (one tile per
// CTA; persistent with M = time) only live on in tools/legacy/ for the stand-alone harness -- the library never launches them.
//
// conv1d as a tcgen05 implicit GEMM with 3xTF32 split precision.
//
//   D[t, co] += sum_ci A[t + k*dil - pad, ci] * W[co, ci, k]        for every tap k
//
// M = 128 time steps (TMEM lanes), N = Cout tile (<= 256, TMEM columns), K = input channels, 8 per MMA.
// Both operands are K-major, SWIZZLE_NONE ("interleave") canonical layouts: a 4-channel slab is a dense
// [rows][4 floats] array (16 B per row), so row r of K-chunk c lives at slab_c + r*16 -- 8-row core matrices are
// contiguous (SBO = 128 B) and the two 16-byte K chunks of one MMA are LBO = slab stride apart.  Because rows are
// uniformly 16 B apart, the activation operand of tap k is THE SAME shared-memory tile with its descriptor start
// address advanced by k*dil rows: one staged window serves all taps (no im2col, no per-tap copies).
//
// fp32 accuracy: x = hi + lo with hi = x & 0xFFFFE000 (exactly representable in TF32) and lo = x - hi (exact in
// fp32; the tensor core keeps its top 11 bits).  D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, accumulated in fp32 in
// TMEM; the dropped lo*lo term is 2^-22 relative.  Activations are split on the way into shared memory (together
// with the fused mask / leaky-ReLU prologue); weights are split once at pack time.
//
// Warp roles (192 threads): warps 0-3 stage activations then run the epilogue (TMEM lane quarter = warp id),
// warp 4 streams per-tap weight blocks with cp.async.bulk + mbarrier transaction counts, warp 5 allocates TMEM and
// one of its lanes issues tcgen05.mma / tcgen05.commit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200tts {
namespace tc {

constexpr int TT = 256;          // time steps per CTA (2 accumulators of 128 lanes)
constexpr int NSLAB = 2;         // 4-channel slabs per activation stage
constexpr int KC = 4 * NSLAB;    // input channels per stage (one MMA k-step per 2 slabs)
constexpr int NB = 4;            // weight ring depth (one tap block each)
constexpr int NA = 4;            // activation ring depth (two producer groups x 2)
constexpr int PGROUP = 64;       // threads per producer group (group g stages chunks c with c%2 == g)
constexpr int MAXIT = 10;        // max (slab,row) items per producer thread per chunk
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 22;

struct TcArgs {
    const float* x; long long x_bs; int x_cs; int Tin;
    const float* xmask; long long xmask_bs; float in_slope;
    const float* w;            // packed [co_tile][chunk][tap]{hi[NSLAB][N][4], lo[NSLAB][N][4]}
    const float* bias;         // [Rows]
    const float* cond; long long cond_bs;
    int Cin, K, dil, pad, Rows, N;   // N = columns per CTA (multiple of 16, <= 256)
    float* y; long long y_bs; int y_cs; int Tout;
    const float* res; long long res_bs; int res_cs;
    const float* ymask; long long ymask_bs;
    float scale; float post_div; int relu; int accum; int mask_post;
    int rows_pad;              // activation slab rows (TT + halo, multiple of 8)
    int* err;                  // device flag: set on a pipeline timeout
    unsigned long long* trace; // optional [grid][16] globaltimer stamps (debug)
    int sleep_ns;              // back-off inside barrier spin loops (0 = none)
    int dbg;                   // debug: 1 = producers only (no MMA / weights / epilogue)
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC_STAMP(slot) do { if (a.trace) a.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int sleep_ns = 0) {
#pragma unroll 1
    for (uint32_t i = 0; i < SPIN_LIMIT; ++i) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
    if (err) { *reinterpret_cast<volatile int*>(err) = 1; __threadfence_system(); }   // mapped host flag (conv1d.cu)
    return false;
}
// non-blocking probe of a phase (mbarrier.test_wait: returns at once, unlike try_wait which may suspend the thread)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle: start address, LBO (between the two 16-byte K chunks), SBO = 128 B (next 8 rows)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version 1 (sm_100)
    return d;                    // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                // c_format = F32
    d |= 2u << 7;                // a_format = TF32
    d |= 2u << 10;               // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(128 >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
}  // namespace b200tts
