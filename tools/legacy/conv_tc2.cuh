// Second-generation tcgen05 conv kernel (persistent, M = time).  Kept for the stand-alone harness only; the library
// (tts_b200/csrc) does not include or launch it.
// Persistent tcgen05 3xTF32 conv1d (second generation of conv_tc.cuh; same math, same operand layouts).
//
// One CTA per SM loops over (batch, row-tile, time-tile) work items; every pipeline runs continuously across
// tiles:
//   warps 4-7  producers : cp.async 16-byte copies of raw [8 ch][time] windows into a 6-deep ring (five chunks of
//                          global latency in flight, zero-fill at the sequence ends), then a shared->shared pass that
//                          applies the fused prologue (leaky-ReLU), splits hi/lo and writes the K-major slabs
//   warp  8    loader    : per-tap weight blocks by cp.async.bulk + mbarrier transaction bytes (6-deep ring)
//   warp  9    MMA       : one lane issues tcgen05.mma kind::tf32 (3 per k-step: lo*hi, hi*lo, hi*hi) and commits
//   warps 0-3  epilogue  : TMEM -> registers -> global with bias / conditioning / ReLU / residual / accumulate /
//                          MRF mean / mask, or the polyphase interleaved store of the transposed-conv layers.
// The accumulator is double buffered in TMEM (2 x [2 x 128 lanes x N columns]), so the epilogue of tile i overlaps
// the main loop of tile i+1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc_v1.cuh"   // first generation + the shared descriptor / barrier helpers

namespace b200tts {
namespace tc2 {

using namespace tc;       // smem_u32, mbar_*, make_desc, make_idesc, mma_tf32, mma_commit, tmem_ld16, fences

constexpr int TT2 = 256;          // time steps per tile (2 x 128-lane accumulators)
constexpr int KC2 = 8;            // input channels per chunk (2 slabs, one MMA k-step)
constexpr int NRAW = 4;           // raw (cp.async) ring depth
constexpr int NA2 = 3;            // transformed activation stages
constexpr int NB2 = 10;           // weight ring depth; each slot holds `tg` consecutive tap blocks so that one
                                  // bulk copy covers >= ~0.4 us of MMA work even for narrow layers (N = 32 / 64)
constexpr int BSLOT_BYTES = 0;     // measured: grouping taps into bigger slots is slower than a deeper ring of single taps
constexpr int NTHREADS2 = 448;    // warps 0-3 + 10-13 epilogue, 4-7 producers, 8 loader, 9 MMA
constexpr int NPROD = 128;

struct Tc2Args {
    const float* x; long long x_bs; int x_cs; int Tin;
    float in_slope;
    const float* w;            // packed [row_tile][chunk][tap]{hi[2][N][4], lo[2][N][4]}
    const float* bias;
    const float* cond; long long cond_bs;
    int Cin, K, dil, pad, Rows, N;
    float* y; long long y_bs; int y_cs; int Tout;
    int ups;                   // 1, or the polyphase factor of a transposed conv (row r -> channel r/ups, phase r%ups)
    int Tq;                    // GEMM columns in time (= Tout for ups == 1)
    const float* res; long long res_bs; int res_cs;
    const float* ymask; long long ymask_bs;
    float scale; float post_div; int relu; int accum; int mask_post;
    int rows_pad;              // slab rows  (TT2 + halo, multiple of 8)
    int raw_w;                 // raw row width in floats (rows_pad + 4, multiple of 4)
    int B, n_ttiles, n_rtiles;
    int tg;                    // tap blocks per weight-ring slot (taps_per_slot(N))
    int* err;
    unsigned long long* trace;  // optional [grid][32] globaltimer stamps (debug)
};

static inline int taps_per_slot(int N) { const int blk = 4 * N * 16; return (BSLOT_BYTES <= blk) ? 1 : BSLOT_BYTES / blk; }
static inline size_t smem_bytes2(int N, int rows_pad, int raw_w) {
    return (size_t)NRAW * KC2 * raw_w * 4 + (size_t)NA2 * (4 * rows_pad * 16) +
           (size_t)NB2 * taps_per_slot(N) * (4 * N * 16) + 512;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

#define TC2_STAMP(slot) do { if (a.trace) a.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime(); } while (0)

__global__ void __launch_bounds__(NTHREADS2, 1) conv1d_tc2_kernel(const Tc2Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, ROWS = a.rows_pad, RAWW = a.raw_w, K = a.K;
    const uint32_t rawStage = (uint32_t)KC2 * RAWW * 4;
    const uint32_t slabA = (uint32_t)ROWS * 16, stageA = 4 * slabA;     // hi[2] + lo[2]
    const uint32_t slabB = (uint32_t)N * 16, stageB = 4 * slabB;       // one tap block
    const int tg = a.tg;
    const uint32_t slotB = (uint32_t)tg * stageB;
    unsigned char* smRaw = smem;
    unsigned char* smA = smRaw + NRAW * rawStage;
    unsigned char* smB = smA + NA2 * stageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + NB2 * slotB);
    const int A_FULL = 0, A_EMPTY = NA2, B_FULL = 2 * NA2, B_EMPTY = 2 * NA2 + NB2, ACC_FULL = 2 * NA2 + 2 * NB2,
              ACC_EMPTY = ACC_FULL + 2, NBARS = ACC_EMPTY + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int nchunks = (a.Cin + KC2 - 1) / KC2;
    const int tiles_total = a.B * a.n_rtiles * a.n_ttiles;
    const int my_tiles = (tiles_total > (int)blockIdx.x) ? (tiles_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t acc_cols = (uint32_t)(2 * N);                 // per accumulator buffer
    uint32_t ncols = 32;
    while (ncols < 2 * acc_cols) ncols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < NA2; ++i) { mbar_init(BAR(A_FULL + i), NPROD); mbar_init(BAR(A_EMPTY + i), 1); }
        for (int i = 0; i < NB2; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(ACC_FULL + i), 1); mbar_init(BAR(ACC_EMPTY + i), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) TC2_STAMP(0);

    auto decode = [&](int it, int& b, int& rt, int& q0) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int tt = tile % a.n_ttiles, rest = tile / a.n_ttiles;
        rt = rest % a.n_rtiles;
        b = rest / a.n_rtiles;
        q0 = tt * TT2;
    };

    if (warp >= 4 && warp < 8) {
        // ============================================================ producers
        const int ptid = tid - 128;
        const int total = my_tiles * nchunks;
        const int vec_per_row = RAWW / 4;
        const int nvec = KC2 * vec_per_row;
        const float slope = a.in_slope;
        bool ok = true;
        // per-thread work items are the same for every chunk: decode them once (no divisions in the loop)
        constexpr int MAXV = 6, MAXI = 5;          // ceil(8*81/128), ceil(2*320/128)
        int v_off[MAXV], v_ch[MAXV], v_t[MAXV];    // raw smem float offset, channel in chunk, time offset from `tal`
#pragma unroll
        for (int e = 0; e < MAXV; ++e) {
            const int v = ptid + e * NPROD;
            const int ch = v / vec_per_row, j = v - ch * vec_per_row;
            v_ch[e] = (v < nvec) ? ch : -1;
            v_t[e] = 4 * j;
            v_off[e] = ch * RAWW + 4 * j;
        }
        int i_raw[MAXI], i_dst[MAXI];              // raw float offset of channel 0 of the slab, slab byte offset
#pragma unroll
        for (int e = 0; e < MAXI; ++e) {
            const int idx = ptid + e * NPROD;
            const int sl = idx / ROWS, r = idx - sl * ROWS;
            i_raw[e] = (idx < 2 * ROWS) ? (4 * sl) * RAWW + r : -1;
            i_dst[e] = (int)(sl * slabA) + r * 16;
        }
        auto issue = [&](int g) {
            if (g < total) {
                const int it = g / nchunks, c = g - it * nchunks;
                int b, rt, q0;
                decode(it, b, rt, q0);
                const int tal = ((q0 - a.pad) & ~3);                     // 16-byte aligned window start (may be < 0)
                const float* xb = a.x + (long long)b * a.x_bs;
                const uint32_t dst0 = smem_u32(smRaw + (g % NRAW) * rawStage);
#pragma unroll
                for (int e = 0; e < MAXV; ++e) {
                    if (v_ch[e] < 0) continue;
                    const int t = tal + v_t[e];
                    const int cg = c * KC2 + v_ch[e];
                    // t is a multiple of 4, so a vector is either wholly before the sequence start (zero fill),
                    // wholly inside, or cut by its end (partial source size, rest zero-filled by the hardware)
                    int nb = 0;
                    if (cg < a.Cin && t >= 0) nb = 4 * max(0, min(4, a.Tin - t));
                    const int tsafe = (t >= 0 && t < a.Tin) ? t : 0;
                    const float* src = xb + (long long)(cg < a.Cin ? cg : 0) * a.x_cs + tsafe;
                    cp_async16_zfill(dst0 + (uint32_t)v_off[e] * 4u, src, (uint32_t)nb);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int g = 0; g < NRAW - 1; ++g) issue(g);
        for (int g = 0; g < total && ok; ++g) {
            asm volatile("cp.async.wait_group %0;" ::"n"(NRAW - 2) : "memory");
            named_bar_sync(1, NPROD);                                     // everyone's copies of chunk g have landed
            issue(g + NRAW - 1);                                          // refills the stage transformed last iteration
            const int as = g % NA2;
            if (g >= NA2) ok = mbar_wait(BAR(A_EMPTY + as), ((g / NA2) - 1) & 1, a.err, 64);
            if (!ok) break;
            const int it = g / nchunks;
            int b, rt, q0;
            decode(it, b, rt, q0);
            const int tin0 = q0 - a.pad, off = tin0 - (tin0 & ~3);
            const float* raw = reinterpret_cast<const float*>(smRaw + (g % NRAW) * rawStage) + off;
            unsigned char* base = smA + as * stageA;
            float u[MAXI][4];
#pragma unroll
            for (int e = 0; e < MAXI; ++e) {           // all shared loads first ...
                const int o = i_raw[e] < 0 ? 0 : i_raw[e];
#pragma unroll
                for (int i = 0; i < 4; ++i) u[e][i] = raw[o + i * RAWW];
            }
#pragma unroll
            for (int e = 0; e < MAXI; ++e) {           // ... then prologue, hi/lo split and the two 16-byte stores
                if (i_raw[e] < 0) continue;
                float4 hi, lo;
                float* ph = &hi.x; float* pl = &lo.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float w_ = u[e][i];
                    w_ = w_ > 0.f ? w_ : w_ * slope;
                    const float h = __uint_as_float(__float_as_uint(w_) & 0xFFFFE000u);
                    ph[i] = h;
                    pl[i] = w_ - h;
                }
                *reinterpret_cast<float4*>(base + i_dst[e]) = hi;
                *reinterpret_cast<float4*>(base + 2 * slabA + i_dst[e]) = lo;
            }
            fence_async_smem();
            mbar_arrive(BAR(A_FULL + as));
            if (ptid == 0) { if (g == 0) TC2_STAMP(1); if (g == nchunks - 1) TC2_STAMP(2); if (g == 2 * nchunks - 1) TC2_STAMP(3); if (g == 4 * nchunks - 1) TC2_STAMP(4); }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (ptid == 0) TC2_STAMP(5);
    } else if (warp == 8) {
        // ============================================================ weight loader
        if (lane == 0) {
            bool ok = true;
            int gi = 0;                                                    // global slot-fill counter
            const int total = nchunks * K;                                 // tap blocks per tile (contiguous in memory)
            for (int it = 0; it < my_tiles && ok; ++it) {
                int b, rt, q0;
                decode(it, b, rt, q0);
                const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.w) + (size_t)rt * total * stageB;
                for (int j = 0; j < total && ok; j += tg, ++gi) {
                    const int st = gi % NB2;
                    if (gi >= NB2) ok = mbar_wait(BAR(B_EMPTY + st), ((gi / NB2) - 1) & 1, a.err, 64);
                    if (!ok) break;
                    const uint32_t bytes = (uint32_t)min(tg, total - j) * stageB;
                    mbar_expect_tx(BAR(B_FULL + st), bytes);
                    bulk_g2s(smem_u32(smB + st * slotB), wsrc + (size_t)j * stageB, bytes, BAR(B_FULL + st));
                }
            }
        }
    } else if (warp == 9) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(N);
            bool ok = true;
            int g = 0, gi = 0;
            const int total = nchunks * K;
            for (int it = 0; it < my_tiles && ok; ++it) {
                const int buf = it & 1;
                int j = 0;                                                 // tap-block index inside the tile
                if (it >= 2) ok = mbar_wait(BAR(ACC_EMPTY + buf), ((it >> 1) - 1) & 1, a.err);
                if (!ok) break;
                tc_fence_after();
                const uint32_t dbase = tmem_base + (uint32_t)buf * acc_cols;
                for (int c = 0; c < nchunks && ok; ++c, ++g) {
                    const int sa = g % NA2;
                    ok = mbar_wait(BAR(A_FULL + sa), (g / NA2) & 1, a.err);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t abase = smem_u32(smA + sa * stageA);
                    // descriptors differ only in the 14-bit start-address field: build once, then add (bytes >> 4)
                    const uint64_t a_hi0 = make_desc(abase, slabA), a_lo0 = make_desc(abase + 2 * slabA, slabA);
                    for (int k = 0; k < K && ok; ++k, ++j) {
                        const int sb = gi % NB2, within = j % tg;
                        if (within == 0) {
                            ok = mbar_wait(BAR(B_FULL + sb), (gi / NB2) & 1, a.err);
                            if (!ok) break;
                            tc_fence_after();
                        }
                        const uint32_t bbase = smem_u32(smB + sb * slotB) + (uint32_t)within * stageB;
                        const uint64_t b_hi = make_desc(bbase, slabB);
                        const uint64_t b_lo = b_hi + (uint64_t)((2 * slabB) >> 4);
#pragma unroll
                        for (int m = 0; m < TT2 / 128; ++m) {
                            const uint64_t arow = (uint64_t)(m * 128 + k * a.dil);   // rows are 16 B apart: +1 per row
                            const uint64_t a_hi = a_hi0 + arow;
                            const uint64_t a_lo = a_lo0 + arow;
                            const uint32_t dcol = dbase + (uint32_t)(m * N);
                            mma_tf32(dcol, a_lo, b_hi, idesc, (c == 0 && k == 0) ? 0u : 1u);
                            mma_tf32(dcol, a_hi, b_lo, idesc, 1u);
                            mma_tf32(dcol, a_hi, b_hi, idesc, 1u);
                        }
                        if (within == tg - 1 || j == total - 1) { mma_commit(BAR(B_EMPTY + sb)); ++gi; }
                    }
                    if (ok) mma_commit(BAR(A_EMPTY + sa));
                }
                if (ok) mma_commit(BAR(ACC_FULL + buf));
                if (it == 0) TC2_STAMP(8); if (it == 1) TC2_STAMP(9); if (it == 3) TC2_STAMP(10);
            }
            TC2_STAMP(11);
        }
        __syncwarp();
    } else {
        // ============================================================ epilogue (warps 0-3: subtile 0, warps 10-13: subtile 1)
        bool ok = true;
        const int ups = a.ups;
        const int lq = warp & 3;                 // TMEM lane quarter this warp may access
        const int m_own = (warp >= 10) ? 1 : 0;
        for (int it = 0; it < my_tiles && ok; ++it) {
            const int buf = it & 1;
            int b, rt, q0;
            decode(it, b, rt, q0);
            ok = mbar_wait(BAR(ACC_FULL + buf), (it >> 1) & 1, a.err, 128);
            if (!ok) break;
            tc_fence_after();
            if (tid == 0) { if (it == 0) TC2_STAMP(16); if (it == 1) TC2_STAMP(18); if (it == 3) TC2_STAMP(20); }
            const uint32_t dbase = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(lq * 32) << 16);
            for (int m = m_own; m <= m_own; ++m) {
                const int q = q0 + m * 128 + lq * 32 + lane;            // GEMM column in time
                const bool qok = q < a.Tq;
                if (ups == 1) {
                    const int t = q;
                    const int tcl = qok ? t : 0;
                    const float mk = (a.ymask && qok) ? __ldg(a.ymask + (long long)b * a.ymask_bs + t) : 1.f;
                    const float* resb = a.res ? a.res + (long long)b * a.res_bs + tcl : nullptr;
                    const float* oldb = a.y + (long long)b * a.y_bs + tcl;
                    float rv[16], ov[16];
                    // order-enforced software pipeline: the (volatile) loads of group j+1 are issued before the
                    // (volatile) TMEM load of group j, so their latency hides behind this group's work
                    auto prefetch = [&](int cg, float* r_, float* o_) {
                        const int r0 = rt * N + cg;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const long long row = (long long)min(r0 + i, a.Rows - 1);
                            if (resb) asm volatile("ld.global.f32 %0, [%1];" : "=f"(r_[i]) : "l"(resb + row * a.res_cs));
                            if (a.accum) asm volatile("ld.global.f32 %0, [%1];" : "=f"(o_[i]) : "l"(oldb + row * a.y_cs));
                        }
                    };
                    prefetch(0, rv, ov);
                    for (int cg = 0; cg < N; cg += 16) {
                        float v[16], rn[16], on[16];
                        if (cg + 16 < N) prefetch(cg + 16, rn, on);
                        tmem_ld16(dbase + (uint32_t)(m * N + cg), v);
                        const int r0 = rt * N + cg;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int r = min(r0 + i, a.Rows - 1);
                            float u = v[i] + a.bias[r];
                            if (a.cond) u += __ldg(a.cond + (long long)b * a.cond_bs + r);
                            if (a.relu) u = fmaxf(u, 0.f);
                            if (a.res) u += rv[i];
                            u *= a.scale;
                            if (a.accum) u += ov[i];
                            if (a.post_div != 1.f) u = u / a.post_div;
                            if (a.mask_post) u *= mk;
                            if (qok && r0 + i < a.Rows) a.y[(long long)b * a.y_bs + (long long)(r0 + i) * a.y_cs + t] = u;
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) { rv[i] = rn[i]; ov[i] = on[i]; }
                    }
                } else {
                    // polyphase store: columns [co*ups, co*ups+ups) of this lane are `ups` consecutive output samples
                    for (int cg = 0; cg < N; cg += 16) {
                        float v[16];
                        tmem_ld16(dbase + (uint32_t)(m * N + cg), v);
                        const int r0 = rt * N + cg;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int r = min(r0 + i, a.Rows - 1);
                            float u = v[i] + a.bias[r];
                            if (a.relu) u = fmaxf(u, 0.f);
                            v[i] = u;
                        }
                        if (!qok) continue;
                        const long long t0 = (long long)q * ups;
                        if (ups == 8 || ups == 4) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const int r = r0 + i, co = r / ups, ph = r - co * ups;
                                float* dst = a.y + (long long)b * a.y_bs + (long long)co * a.y_cs + t0 + ph;
                                if (r + 3 < a.Rows && t0 + ph + 3 < a.Tout && ((a.y_cs & 3) == 0)) {
                                    *reinterpret_cast<float4*>(dst) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) if (r + e < a.Rows && t0 + ph + e < a.Tout) dst[e] = v[i + e];
                                }
                            }
                        } else if (ups == 2) {
#pragma unroll
                            for (int i = 0; i < 16; i += 2) {
                                const int r = r0 + i, co = r >> 1;
                                float* dst = a.y + (long long)b * a.y_bs + (long long)co * a.y_cs + t0;
                                if (r + 1 < a.Rows && t0 + 1 < a.Tout && ((a.y_cs & 1) == 0)) {
                                    *reinterpret_cast<float2*>(dst) = make_float2(v[i], v[i + 1]);
                                } else {
                                    if (r < a.Rows && t0 < a.Tout) dst[0] = v[i];
                                    if (r + 1 < a.Rows && t0 + 1 < a.Tout) dst[1] = v[i + 1];
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int r = r0 + i, co = r / ups, ph = r - co * ups;
                                if (r < a.Rows && t0 + ph < a.Tout) a.y[(long long)b * a.y_bs + (long long)co * a.y_cs + t0 + ph] = v[i];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(BAR(ACC_EMPTY + buf));
            if (tid == 0) { if (it == 0) TC2_STAMP(17); if (it == 1) TC2_STAMP(19); if (it == 3) TC2_STAMP(21); }
        }
        if (tid == 0) TC2_STAMP(22);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

}  // namespace tc2
}  // namespace b200tts
