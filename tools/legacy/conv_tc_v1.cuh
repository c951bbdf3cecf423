// First-generation tcgen05 conv kernel (one 256-step tile per CTA, M = time).  Kept for the stand-alone harness
// (tools/test_conv_tc.cu) as a cross-check; the library (tts_b200/csrc) does not include or launch it.
#pragma once
#include "../../tts_b200/csrc/conv_tc.cuh"

namespace b200tts {
namespace tc {

static inline size_t smem_bytes(int N, int rows_pad) {
    return (size_t)NA * (2 * NSLAB * rows_pad * 16) + (size_t)NB * (2 * NSLAB * N * 16) + 256;
}

__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, ROWS = a.rows_pad;
    const uint32_t slabA = (uint32_t)ROWS * 16;            // bytes per 4-channel activation slab
    const uint32_t stageA = 2 * NSLAB * slabA;             // hi[NSLAB] + lo[NSLAB]
    const uint32_t slabB = (uint32_t)N * 16;
    const uint32_t stageB = 2 * NSLAB * slabB;
    unsigned char* smA = smem;
    unsigned char* smB = smem + NA * stageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + NB * stageB);
    // bars: a_full[NA], a_empty[NA], b_full[NB], b_empty[NB], acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NB + 1);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int A_FULL = 0, A_EMPTY = NA, B_FULL = 2 * NA, B_EMPTY = 2 * NA + NB, ACC_FULL = 2 * NA + 2 * NB;

    const int b = blockIdx.z, tile_co = blockIdx.y;
    const int q0 = blockIdx.x * TT;
    const int nchunks = (a.Cin + KC - 1) / KC;
    const int K = a.K;
    if (tid == 0) TC_STAMP(11);
    const uint32_t ncols = (uint32_t)(2 * N) <= 32 ? 32u : ((2 * N) <= 64 ? 64u : ((2 * N) <= 128 ? 128u : ((2 * N) <= 256 ? 256u : 512u)));

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(BAR(A_FULL + i), PGROUP); mbar_init(BAR(A_EMPTY + i), 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
        mbar_init(BAR(ACC_FULL), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) TC_STAMP(0);

    if (warp < 4) {
        // ------------------------------------------------------------ activation producers
        const float* xb = a.x + (long long)b * a.x_bs;
        const float* mb = a.xmask ? a.xmask + (long long)b * a.xmask_bs : nullptr;
        const int tin0 = q0 - a.pad;
        const float slope = a.in_slope;
        bool ok = true;
        const int grp = warp >> 1, gtid = tid & (PGROUP - 1);
        const int nitem = NSLAB * ROWS;
        for (int c = grp; c < nchunks && ok; c += 2) {
            const int st = c % NA;
            const int c0 = c * KC;
            // issue every global load of this chunk first (one memory round trip per chunk) ...
            if (gtid == 0 && c == 2) TC_STAMP(12);
            float v[MAXIT][4], mm[MAXIT];
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {     // loads only: nothing here may consume a loaded value
                const int idx = gtid + it * PGROUP;
                const int s = idx / ROWS, r = idx - s * ROWS;       // slab, row (row fastest across lanes)
                const int t = tin0 + r;
                const bool tok = (idx < nitem) && (t >= 0) && (t < a.Tin);
                const int tc = tok ? t : 0;
                mm[it] = (tok && mb) ? __ldg(mb + tc) : 1.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ch = c0 + 4 * s + i;
                    const bool ld = tok && (ch < a.Cin);
                    const float* ptr = xb + (long long)(ld ? ch : 0) * a.x_cs + tc;
                    const float got = __ldg(ptr);                    // always a valid address; select afterwards
                    v[it][i] = ld ? got : 0.f;
                }
            }
            if (gtid == 0 && c == 2) { float sacc = 0.f;
#pragma unroll
                for (int it = 0; it < MAXIT; ++it) sacc += v[it][0] + v[it][1] + v[it][2] + v[it][3];
                if (sacc == 1234.5678f) a.err[0] = 2; TC_STAMP(13); }
            // ... then wait for the stage, apply the prologue, split hi/lo and store
            if (c >= NA && a.dbg != 1) ok = mbar_wait(BAR(A_EMPTY + st), ((c / NA) - 1) & 1, a.err, a.sleep_ns);
            unsigned char* base = smA + st * stageA;
#pragma unroll
            for (int it = 0; it < MAXIT; ++it) {
                const int idx = gtid + it * PGROUP;
                if (idx < nitem) {
                    const int s = idx / ROWS, r = idx - s * ROWS;
                    float4 hi, lo;
                    float* ph = &hi.x; float* pl = &lo.x;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float u = v[it][i] * mm[it];
                        u = u > 0.f ? u : u * slope;
                        const float h = __uint_as_float(__float_as_uint(u) & 0xFFFFE000u);
                        ph[i] = h;
                        pl[i] = u - h;
                    }
                    *reinterpret_cast<float4*>(base + s * slabA + r * 16) = hi;
                    *reinterpret_cast<float4*>(base + (NSLAB + s) * slabA + r * 16) = lo;
                }
            }
            if (gtid == 0 && c == 2) TC_STAMP(14);
            fence_async_smem();
            mbar_arrive(BAR(A_FULL + st));
            if (gtid == 0 && c == 2) TC_STAMP(15);
            if (gtid == 0 && c == grp) TC_STAMP(1 + grp);
        }
        if (gtid == 0) TC_STAMP(3 + grp);
        // ------------------------------------------------------------ epilogue
        if (a.dbg == 1) ok = false;
        if (ok) ok = mbar_wait(BAR(ACC_FULL), 0, a.err, a.sleep_ns);
        tc_fence_after();
        if (tid == 0) TC_STAMP(5);
        if (ok) {
            for (int m = 0; m < TT / 128; ++m) {
                const int t = q0 + m * 128 + warp * 32 + lane;
                const bool tok = t < a.Tout;
                const float mk = (a.ymask && tok) ? __ldg(a.ymask + (long long)b * a.ymask_bs + t) : 1.f;
                for (int cg = 0; cg < N; cg += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(m * N + cg), v);
                    const int r0 = tile_co * N + cg;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = r0 + i;
                        float u = v[i] + ((r < a.Rows) ? a.bias[r] : 0.f);
                        if (a.cond && r < a.Rows) u += __ldg(a.cond + (long long)b * a.cond_bs + r);
                        if (a.relu) u = fmaxf(u, 0.f);
                        v[i] = u;
                    }
                    const int tcl = tok ? t : 0;
                    if (a.res) {
                        float rv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = (r0 + i < a.Rows) ? r0 + i : 0;
                            rv[i] = a.res[(long long)b * a.res_bs + (long long)rr * a.res_cs + tcl];
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += rv[i];
                    }
                    if (a.accum) {
                        float ov[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = (r0 + i < a.Rows) ? r0 + i : 0;
                            ov[i] = a.y[(long long)b * a.y_bs + (long long)rr * a.y_cs + tcl];
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = v[i] * a.scale + ov[i];
                    } else if (a.scale != 1.f) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] *= a.scale;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float u = v[i];
                        if (a.post_div != 1.f) u = u / a.post_div;
                        if (a.mask_post) u *= mk;
                        if (tok && r0 + i < a.Rows) a.y[(long long)b * a.y_bs + (long long)(r0 + i) * a.y_cs + t] = u;
                    }
                }
            }
        }
        tc_fence_before();
        if (tid == 0) TC_STAMP(6);
    } else if (warp == 4) {
        // ------------------------------------------------------------ weight loader (one lane)
        if (lane == 0 && a.dbg != 1) {
            const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.w) + (size_t)tile_co * nchunks * K * stageB;
            bool ok = true;
            const int total = nchunks * K;
            for (int i = 0; i < total && ok; ++i) {
                const int st = i % NB;
                if (i >= NB) ok = mbar_wait(BAR(B_EMPTY + st), ((i / NB) - 1) & 1, a.err, a.sleep_ns);
                if (!ok) break;
                mbar_expect_tx(BAR(B_FULL + st), stageB);
                bulk_g2s(smem_u32(smB + st * stageB), wsrc + (size_t)i * stageB, stageB, BAR(B_FULL + st));
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer (one lane of warp 5)
        if (lane == 0 && a.dbg != 1) {
            const uint32_t idesc = make_idesc(N);
            bool ok = true;
            int i = 0;
            TC_STAMP(7);
            for (int c = 0; c < nchunks && ok; ++c) {
                const int sa = c % NA;
                ok = mbar_wait(BAR(A_FULL + sa), (c / NA) & 1, a.err, a.sleep_ns);
                if (!ok) break;
                tc_fence_after();
                if (c == 0) TC_STAMP(8);
                const uint32_t abase = smem_u32(smA + sa * stageA);
                for (int k = 0; k < K && ok; ++k, ++i) {
                    const int sb = i % NB;
                    ok = mbar_wait(BAR(B_FULL + sb), (i / NB) & 1, a.err, a.sleep_ns);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t bbase = smem_u32(smB + sb * stageB);
#pragma unroll
                    for (int m = 0; m < TT / 128; ++m) {
                        const uint32_t arow = (uint32_t)(m * 128 + k * a.dil) * 16u;
                        const uint32_t dcol = tmem_base + (uint32_t)(m * N);
#pragma unroll
                        for (int s = 0; s < KC / 8; ++s) {
                            const uint64_t a_hi = make_desc(abase + (2 * s) * slabA + arow, slabA);
                            const uint64_t a_lo = make_desc(abase + (NSLAB + 2 * s) * slabA + arow, slabA);
                            const uint64_t b_hi = make_desc(bbase + (2 * s) * slabB, slabB);
                            const uint64_t b_lo = make_desc(bbase + (NSLAB + 2 * s) * slabB, slabB);
                            const uint32_t first = (c == 0 && k == 0 && s == 0) ? 0u : 1u;
                            mma_tf32(dcol, a_lo, b_hi, idesc, first);   // small terms first
                            mma_tf32(dcol, a_hi, b_lo, idesc, 1u);
                            mma_tf32(dcol, a_hi, b_hi, idesc, 1u);
                        }
                    }
                    mma_commit(BAR(B_EMPTY + sb));
                }
                if (ok) mma_commit(BAR(A_EMPTY + sa));
            }
            if (ok) mma_commit(BAR(ACC_FULL));
            TC_STAMP(9);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
        if (lane == 0) TC_STAMP(10);
    }
}

}  // namespace tc
}  // namespace b200tts
