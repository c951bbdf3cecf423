// Persistent tcgen05 3xTF32 conv1d, third generation: M = output rows (weights are the A operand), N = 256 time
// steps (the activation window is the B operand), D[row, t] in TMEM (128 lanes x 256 columns, double buffered).
//
// Why this orientation: measured on B200 (profiles/r01_tc_notes.md) an SS-mode tcgen05.mma with M = 128 costs ~135
// cycles whatever N is -- the 128-row A operand streams from shared memory at about one 32-byte row per clock.  With
// the time axis as N = 256 each instruction carries 128 x 256 x 8 MACs in those ~128 cycles instead of 128 x N_cout x 8.
// The activation tile keeps the row-shift property (rows are 16 B apart), so tap k is still just a descriptor
// start-address offset of k*dil rows -- now on the B operand.
//
//   warps 4-7        producers : cp.async raw [8 ch][time] windows (4-deep ring) -> leaky-ReLU + hi/lo split -> K-major slabs
//   warp  8          loader    : per-tap weight blocks {hi,lo}[2 slabs][128 rows][4] by cp.async.bulk (10-deep ring)
//   warp  9          MMA       : one lane issues 3 tcgen05.mma (lo*hi, hi*lo, hi*hi) per tap and chunk, commits to mbarriers
//   warps 0-3,10-13  epilogue  : lane = output row, columns = time: TMEM -> registers -> float4 global stores (two halves
//                                of the 256 columns), residual / accumulate loads as float4 with an order-enforced prefetch
//
// Grouped mode (GRP = 2 / 4) for narrow layers (exactly 64 / 32 output rows, the last two HiFiGAN stages): the 128
// MMA rows are GRP tap-groups x (128/GRP) channels -- row (channel c, group g) carries the weights of taps g, g+GRP,
// g+2*GRP, ... so one instruction stream of ceil(K/GRP) "tap blocks" (B shifted by GRP*dil rows per block) replaces K
// of them and no MMA row is zero padding.  D_g[c, col] then still misses its own g*dil shift; the epilogue applies it
// as a TMEM column offset (one tcgen05.ld per group), and sums the GRP partials, which live in adjacent lanes of one
// warp, with a shuffle reduce-scatter.  Tiles advance by 240 columns so every shifted read stays inside the
// 256-column accumulator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "conv_tc.cuh"   // descriptor / barrier helpers

namespace b200tts {
namespace tc3 {

using namespace tc;       // smem_u32, mbar_*, make_desc, make_idesc, mma_tf32, mma_commit, tmem_ld16, fences

constexpr int TT2 = 256;          // time steps per tile = MMA N
constexpr int MROWS = 128;        // output rows per tile = MMA M (weight rows are zero padded up to it)
constexpr int KC2 = 8;            // input channels per chunk (2 slabs, one MMA k-step)
#ifndef TC3_NRAW
#define TC3_NRAW 4
#endif
#ifndef TC3_NA2
#define TC3_NA2 3
#endif
#ifndef TC3_NB2
#define TC3_NB2 10
#endif
constexpr int NRAW = TC3_NRAW;    // raw (cp.async) ring depth
constexpr int NA2 = TC3_NA2;      // transformed activation stages
constexpr int NB2 = TC3_NB2;      // weight ring depth (one 8 KB tap block per slot)
constexpr int NTHREADS2 = 448;    // warps 0-3 + 10-13 epilogue, 4-7 producers, 8 loader, 9 MMA
constexpr int NPROD = 128;

struct Tc3Args {
    const float* x; long long x_bs; int x_cs; int Tin;
    float in_slope;
    const float* w;            // packed [row_tile][chunk][tap]{hi[2][128][4], lo[2][128][4]}
    const float* bias;
    const float* cond; long long cond_bs;
    int Cin, K, dil, pad, Rows, N;
    float* y; long long y_bs; int y_cs; int Tout;
    int ups;                   // 1, or the polyphase factor of a transposed conv (row r -> channel r/ups, phase r%ups)
    int Tq;                    // GEMM columns in time (= Tout for ups == 1)
    const float* res; long long res_bs; int res_cs;
    const float* ymask; long long ymask_bs;
    float scale; float post_div; int relu; int accum; int mask_post;
    int mask_pre;              // multiply by ymask before the residual / accumulate (coupling `post`)
    int gate;                  // rows are (tanh, sigmoid) pairs: out[r/2] = tanh(v[2p]) * sigmoid(v[2p+1])  (WaveNet)
    int split;                 // > 0: rows < split -> y (accumulate, mask); rows >= split -> y2 (accumulate iff accum2)
    float* y2; long long y2_bs; int y2_cs; int accum2;
    int KJ;                    // tap blocks per chunk (= K, or ceil(K / GRP) in grouped mode)
    int dil_blk;               // B-row shift between tap blocks (= dil, or GRP * dil)
    int tstep;                 // time steps a tile advances (= TT2, or 240 in grouped mode)
    int rows_pad;              // slab rows  (TT2 + halo, multiple of 8)
    int raw_w;                 // raw row width in floats (rows_pad + 4, multiple of 4)
    int B, n_ttiles, n_rtiles;
    int* err;
    unsigned long long* trace;  // optional [grid][32] globaltimer stamps (debug)
    int stage;                  // != 0: wide-layer epilogue goes through per-warp shared tiles (coalesced global access)
    int stage_off;              // byte offset of those tiles in dynamic shared memory (8 warps x 5120 B)
    int dbg;                    // harness-only bottleneck probes: 1 no cp.async, 2 no transform, 4 no epilogue loads, 8 no stores, 16 no MMA
    // ---- ragged batches (null lens: every row spans the full tensor).  Row b only has tiles for GEMM columns below
    // min(Tq, lens[b] * rate_q + need_q) and its input is read as zero from min(Tin, lens[b] * rate_in + need_in) on:
    // padded frames cost nothing, and `need` keeps every sample below lens[b] bit-identical to the full computation
    // (it is the receptive field of the layers that still follow, worked out per launch by the engine).
    const int* lens; int rate_q, need_q, rate_in, need_in;
    int pref_off;               // byte offset in dynamic shared memory of the (B + 1)-entry tile prefix table
};

static inline size_t smem_bytes3(int rows_pad, int raw_w) {
    return (size_t)NRAW * KC2 * raw_w * 4 + (size_t)NA2 * (4 * rows_pad * 16) + (size_t)NB2 * (4 * MROWS * 16) + 512;
}
static inline size_t ragged_table_bytes(int B) { return ((size_t)(B + 1) * sizeof(int) + 15) / 16 * 16; }

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// tcgen05.ld without the wait: issue several, then one tmem_wait_ld() (the loads' latencies overlap)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void tmem_ld_nowait(uint32_t taddr, uint32_t* r) {   // N = 1, 2, 4, 8 or 16 columns
    if constexpr (N == 16) tmem_ld16_nowait(taddr, r);
    else if constexpr (N == 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    else if constexpr (N == 4)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    else if constexpr (N == 2)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
    else
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

#define TC3_STAMP(slot) do { if (a.trace) a.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime(); } while (0)

constexpr int STAGE_BYTES = 8 * 5120;   // staged epilogue: per epilogue warp an output and a residual tile of [32][20] floats
constexpr int TSTEP_GROUPED = 240;   // 15 chunks of 16 columns: leaves room for the (GRP-1)*dil <= 15 column shift

template <int GRP, int DIL, bool STAGED = false>   // DIL > 0: the layer's dilation as a compile-time constant (grouped epilogue);
                                                   // STAGED: wide-layer epilogue through per-warp shared tiles
__device__ __forceinline__ void tc3_body(const Tc3Args& a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ROWS = a.rows_pad, RAWW = a.raw_w, K = a.KJ;
    const uint32_t rawStage = (uint32_t)KC2 * RAWW * 4;
    const uint32_t slabA = (uint32_t)ROWS * 16, stageA = 4 * slabA;     // hi[2] + lo[2]
    const uint32_t slabB = (uint32_t)MROWS * 16, stageB = 4 * slabB;   // one tap block: {hi,lo}[2 slabs][128 rows][16 B]
    unsigned char* smRaw = smem;
    unsigned char* smA = smRaw + NRAW * rawStage;
    unsigned char* smB = smA + NA2 * stageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smB + NB2 * stageB);
    const int A_FULL = 0, A_EMPTY = NA2, B_FULL = 2 * NA2, B_EMPTY = 2 * NA2 + NB2, ACC_FULL = 2 * NA2 + 2 * NB2,
              ACC_EMPTY = ACC_FULL + 2, NBARS = ACC_EMPTY + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int nchunks = (a.Cin + KC2 - 1) / KC2;
    const bool ragged = a.lens != nullptr;
    int* pref = reinterpret_cast<int*>(smem + a.pref_off);    // pref[b] = first tile of row b (ragged only)
    if (ragged) {
        // tiles per row from its own length; exclusive prefix by warp 0 (rows in lane-contiguous chunks).  `lens` was
        // written several launches ago (durations kernel), so reading it before griddepcontrol.wait is safe.
        for (int b = tid; b < a.B; b += NTHREADS2) {
            const long long e = (long long)a.lens[b] * a.rate_q + a.need_q;
            const int ext = (int)(e < (long long)a.Tq ? (e > 0 ? e : 0) : (long long)a.Tq);
            pref[b + 1] = ((ext + a.tstep - 1) / a.tstep) * a.n_rtiles;
        }
        __syncthreads();
        if (warp == 0) {
            const int per = (a.B + 31) / 32, lo = lane * per, hi = min(a.B, lo + per);
            int sum = 0;
            for (int b = lo; b < hi; ++b) sum += pref[b + 1];
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
            int run = incl - sum;
            for (int b = lo; b < hi; ++b) { const int c = pref[b + 1]; pref[b + 1] = run + c; run += c; }
            if (lane == 0) pref[0] = 0;
        }
        __syncthreads();
    }
    const int tiles_total = ragged ? pref[a.B] : a.B * a.n_rtiles * a.n_ttiles;
    const int my_tiles = (tiles_total > (int)blockIdx.x) ? (tiles_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t acc_cols = (uint32_t)TT2;                     // per accumulator buffer
    const uint32_t ncols = 512;

    if (tid == 0) {
        for (int i = 0; i < NA2; ++i) { mbar_init(BAR(A_FULL + i), NPROD / 32); mbar_init(BAR(A_EMPTY + i), 1); }
        for (int i = 0; i < NB2; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(ACC_FULL + i), 1); mbar_init(BAR(ACC_EMPTY + i), 8); }
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
    if (tid == 0) TC3_STAMP(0);
    // Programmatic dependent launch: the next layer's CTAs may take SMs as this grid drains (they park in their own
    // griddepcontrol.wait); every role that touches activations waits for the previous layer here.  The weight loader
    // (warp 8) reads only constants and starts filling its ring at once.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != 8) asm volatile("griddepcontrol.wait;" ::: "memory");

    auto decode = [&](int it, int& b, int& rt, int& q0) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        if (ragged) {
            int lo = 0, hi = a.B;                   // largest b with pref[b] <= tile (rows without tiles are skipped)
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pref[mid] <= tile) lo = mid; else hi = mid; }
            b = lo;
            const int local = tile - pref[b], nt = (pref[b + 1] - pref[b]) / a.n_rtiles;
            rt = local / nt;
            q0 = (local - rt * nt) * a.tstep;
            return;
        }
        const int tt = tile % a.n_ttiles, rest = tile / a.n_ttiles;
        rt = rest % a.n_rtiles;
        b = rest / a.n_rtiles;
        q0 = tt * a.tstep;
    };
    auto input_extent = [&](int b) -> int {       // columns of x[b] that hold data; beyond it the operand is zero
        if (!ragged) return a.Tin;
        const long long e = (long long)a.lens[b] * a.rate_in + a.need_in;
        return (int)(e < (long long)a.Tin ? (e > 0 ? e : 0) : (long long)a.Tin);
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
        // running positions instead of divisions / modulos per chunk (this loop is on the critical path of the K <= 7 layers)
        int iss_it = 0, iss_c = 0, iss_ring = 0, iss_b = 0, iss_q0 = 0, iss_Tin = 0;   // cp.async front: tile, chunk, raw slot
        bool iss_new = true;
        int tr_it = 0, tr_c = 0, tr_ring = 0, tr_q0 = 0, as = 0;                       // transform: tile, chunk, raw slot, A stage
        uint32_t pa_empty = 1;                                    // parity to wait for on A_EMPTY[as]: round 1 -> 0, round 2 -> 1, ...
        bool tr_new = true;
        const bool no_cp = (a.dbg & 1) != 0;
        auto issue = [&](int g) {
            if (g < total && !no_cp) {
                if (iss_new) {
                    int rt_;
                    decode(iss_it, iss_b, rt_, iss_q0);
                    iss_Tin = input_extent(iss_b);
                    iss_new = false;
                }
                const int c = iss_c;
                const int b = iss_b, q0 = iss_q0, Tin_b = iss_Tin;
                const int tal = ((q0 - a.pad) & ~3);                     // 16-byte aligned window start (may be < 0)
                const float* xb = a.x + (long long)b * a.x_bs;
                const uint32_t dst0 = smem_u32(smRaw + iss_ring * rawStage);
#pragma unroll
                for (int e = 0; e < MAXV; ++e) {
                    if (v_ch[e] < 0) continue;
                    const int t = tal + v_t[e];
                    const int cg = c * KC2 + v_ch[e];
                    // t is a multiple of 4, so a vector is either wholly before the sequence start (zero fill),
                    // wholly inside, or cut by its end (partial source size, rest zero-filled by the hardware)
                    int nb = 0;
                    if (cg < a.Cin && t >= 0) nb = 4 * max(0, min(4, Tin_b - t));
                    const int tsafe = (t >= 0 && t < Tin_b) ? t : 0;
                    const float* src = xb + (long long)(cg < a.Cin ? cg : 0) * a.x_cs + tsafe;
                    cp_async16_zfill(dst0 + (uint32_t)v_off[e] * 4u, src, (uint32_t)nb);
                }
                if (++iss_c == nchunks) { iss_c = 0; ++iss_it; iss_new = true; }
            }
            if (++iss_ring == NRAW) iss_ring = 0;
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int g = 0; g < NRAW - 1; ++g) issue(g);
        for (int g = 0; g < total && ok; ++g) {
            asm volatile("cp.async.wait_group %0;" ::"n"(NRAW - 2) : "memory");
            named_bar_sync(1, NPROD);                                     // everyone's copies of chunk g have landed
            issue(g + NRAW - 1);                                          // refills the stage transformed last iteration
            if (g >= NA2) ok = mbar_wait(BAR(A_EMPTY + as), pa_empty, a.err);
            if (!ok) break;
            if (tr_new) {
                int b_, rt_;
                decode(tr_it, b_, rt_, tr_q0);
                tr_new = false;
            }
            const int q0 = tr_q0;
            const int tin0 = q0 - a.pad, off = tin0 - (tin0 & ~3);
            const float* raw = reinterpret_cast<const float*>(smRaw + tr_ring * rawStage) + off;
            unsigned char* base = smA + as * stageA;
            float u[MAXI][4];
            if (!(a.dbg & 2)) {
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
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A_FULL + as));       // one arrival per producer warp
            if (++tr_c == nchunks) { tr_c = 0; ++tr_it; tr_new = true; }
            if (++tr_ring == NRAW) tr_ring = 0;
            if (++as == NA2) { as = 0; pa_empty ^= 1u; }
            if (ptid == 0) { if (g == 0) TC3_STAMP(1); if (g == nchunks - 1) TC3_STAMP(2); if (g == 2 * nchunks - 1) TC3_STAMP(3); if (g == 4 * nchunks - 1) TC3_STAMP(4); }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (ptid == 0) TC3_STAMP(5);
    } else if (warp == 8) {
        // ============================================================ weight loader
        // One lane per ring slot: lane s owns slot s and feeds it with tap blocks s, s + NB2, s + 2*NB2, ... of this CTA's
        // block sequence.  A single thread walking the ring paid its wait -> expect_tx -> bulk-copy chain (~400 cycles) once
        // per 8 KB block -- measured as a 20 B/clk "L2 limit" that was really this thread (r02 ablation: the kernel without
        // MMAs still took 75 % of its time).  Ten independent chains keep ten copies in flight.
        if (lane < NB2) {
            const int total = nchunks * K;                                 // tap blocks per tile (contiguous in memory)
            const long long all = (long long)my_tiles * total;
            const uint32_t full = BAR(B_FULL + lane), empty = BAR(B_EMPTY + lane);
            const uint32_t dst = smem_u32(smB + lane * stageB);
            int it = 0, j = lane;                                          // block gi = it * total + j
            while (j >= total && it < my_tiles) { j -= total; ++it; }
            int cur_it = -1;
            const unsigned char* wsrc = nullptr;
            uint32_t par = 1;                                              // parity of B_EMPTY to wait for: round 1 -> 0, 2 -> 1
            bool ok = true, first = true;
            for (long long gi = lane; gi < all && ok; gi += NB2) {
                if (it != cur_it) {
                    int b_, rt, q0_;
                    decode(it, b_, rt, q0_);
                    wsrc = reinterpret_cast<const unsigned char*>(a.w) + (size_t)rt * total * stageB;
                    cur_it = it;
                }
                if (!first) { ok = mbar_wait(empty, par, a.err); if (!ok) break; }
                mbar_expect_tx(full, stageB);
                bulk_g2s(dst, wsrc + (size_t)j * stageB, stageB, full);
                first = false;
                par ^= 1u;
                j += NB2;
                while (j >= total) { j -= total; ++it; }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ============================================================ MMA issuer
        // ncu (profiles/r02_tc3_issue_loop.md): this ONE thread is what bounds the MMA-heavy layers -- it never waits long
        // on a barrier, its own dependent instruction chain took ~710 cycles per tap against the 384 cycles the three MMAs
        // need on the tensor pipe.  So the loop carries no divisions / modulos / constant-bank loads / descriptor
        // rebuilds: ring positions, parities, barrier addresses and descriptors are running counters, and the next tap's
        // weight barrier is tested (non-blocking) BEFORE the current tap's MMAs are issued so its latency is off the chain.
        // The whole warp runs the loop CONVERGED (every lane computes the same ring positions / descriptors, so the
        // compiler keeps them in uniform registers and feeds UTCHMMA directly -- inside an `if (lane == 0)` every operand
        // went through R2UR and each tcgen05 instruction got its own ELECT loop); one elected lane issues.
        {
            uint32_t leader;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
            // M = 128 rows (weights), N = 256 time steps
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TT2 >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
            const uint64_t xstep = (uint64_t)a.dil_blk;                                   // B rows per tap (16 B each)
            const uint64_t wlo_off = (uint64_t)((2 * slabB) >> 4), wslot = (uint64_t)(stageB >> 4);
            const uint64_t wdesc0 = make_desc(smem_u32(smB), slabB);
            const uint32_t bfull0 = BAR(B_FULL), bempty0 = BAR(B_EMPTY);
            const bool no_mma = (a.dbg & 16) != 0;
            bool ok = true;
            int sa = 0; uint32_t pa = 0;                                                  // activation stage / its parity
            int sb = 0; uint32_t pb = 0;                                                  // weight slot / its parity
            uint64_t wdesc = wdesc0;
            uint32_t bfull = bfull0, bempty = bempty0;
            bool have = false;                                                            // B_FULL[sb] already seen complete
            for (int it = 0; it < my_tiles && ok; ++it) {
                const int buf = it & 1;
                if (it >= 2) ok = mbar_wait(BAR(ACC_EMPTY + buf), ((it >> 1) - 1) & 1, a.err);
                if (!ok) break;
                tc_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)buf * acc_cols;
                uint32_t acc = 0u;
                for (int c = 0; c < nchunks && ok; ++c) {
                    ok = mbar_wait(BAR(A_FULL + sa), pa, a.err);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t abase = smem_u32(smA + sa * stageA);
                    // descriptors differ only in the 14-bit start-address field: build once, then add rows (16 B each)
                    uint64_t xh = make_desc(abase, slabA), xl = make_desc(abase + 2 * slabA, slabA);
#pragma unroll 1
                    for (int k = 0; k < K; ++k) {
                        if (!have) { ok = mbar_wait(bfull, pb, a.err); if (!ok) break; }
                        tc_fence_after();
                        // where the NEXT tap's weights will be; peek at their barrier now (result used next iteration)
                        const bool wrap = (sb == NB2 - 1);
                        const uint32_t nfull = wrap ? bfull0 : bfull + 8u, npb = wrap ? (pb ^ 1u) : pb;
                        const bool have_next = mbar_test(nfull, npb);
                        const uint64_t w_hi = wdesc, w_lo = wdesc + wlo_off;
                        if (leader) {
                            if (!no_mma) {
                                mma_tf32(dcol, w_hi, xl, idesc, acc);                     // small terms first
                                mma_tf32(dcol, w_lo, xh, idesc, 1u);
                                mma_tf32(dcol, w_hi, xh, idesc, 1u);
                            }
                            mma_commit(bempty);
                        }
                        acc = 1u;
                        xh += xstep; xl += xstep;
                        if (wrap) { sb = 0; wdesc = wdesc0; bempty = bempty0; }
                        else { ++sb; wdesc += wslot; bempty += 8u; }
                        bfull = nfull; pb = npb;
                        have = __all_sync(0xffffffffu, have_next);                        // keep the warp's control flow uniform
                    }
                    if (ok && leader) mma_commit(BAR(A_EMPTY + sa));
                    if (++sa == NA2) { sa = 0; pa ^= 1u; }
                }
                if (ok && leader) mma_commit(BAR(ACC_FULL + buf));
                if (lane == 0) { if (it == 0) TC3_STAMP(8); if (it == 1) TC3_STAMP(9); if (it == 3) TC3_STAMP(10); }
            }
            if (lane == 0) TC3_STAMP(11);
        }
        __syncwarp();
    } else {
        // ============================================================ epilogue (lane = output row, columns = time)
        bool ok = true;
        const int ups = a.ups;
        const int lq = warp & 3;                     // TMEM lane quarter this warp may access
        const int half = (warp >= 10) ? 1 : 0;       // warps 0-3: columns [0,128), warps 10-13: [128,256)
        if constexpr (GRP > 1) {
            // ---- grouped epilogue: lane = (channel cc, tap group g); out[c, t] = sum_g D_g[c, t + g*dil]
            // Written for instruction count and a small footprint (profiles/r01_tc_grouped_notes.md): one TMEM window per
            // 16 columns, the per-lane shift folded into the reduce-scatter's selects, residual loads two groups ahead.
            // (Unrolling the eight groups of a tile to prefetch a whole tile ahead thrashed the instruction cache.)
                        constexpr int CPW = 32 / GRP, NV = 16 / GRP;
            const int g = lane & (GRP - 1), cc = lane / GRP;
            const bool g0 = (g & 1) != 0, g1 = (g & 2) != 0;
            const int co = lq * CPW + cc;                                   // Rows == 128 / GRP
            const int coff = (GRP == 4) ? ((g & 1) * 8 + (g >> 1) * 4) : g * 8;   // this lane's columns in a group
            const int cbeg = half ? 128 : 0;
            const bool has_res = a.res != nullptr && !(a.dbg & 4);
            const bool acc_r = a.accum != 0 && !(a.dbg & 4);
            const bool vec_ok = ((a.y_cs & 3) == 0) && ((a.y_bs & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0) &&
                                (!a.res || (((a.res_cs & 3) == 0) && ((a.res_bs & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.res) & 15) == 0)));
            const bool relu = a.relu != 0, do_store = !(a.dbg & 8);
            const float scale = a.scale, post_div = a.post_div;
            bool fast = false;                                            // interior tile: no bounds checks at all
            auto prefetch_any = [&](const float* rr, int q, float* dst) {   // values of one column group of a row -> registers

#pragma unroll
                for (int j = 0; j < NV / 4; ++j) {
                    const int qq = q + 4 * j;
                    if (fast || (vec_ok && qq + 3 < a.Tout)) {
                        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(dst[4 * j]), "=f"(dst[4 * j + 1]), "=f"(dst[4 * j + 2]), "=f"(dst[4 * j + 3]) : "l"(rr + qq));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) asm volatile("ld.global.f32 %0, [%1];" : "=f"(dst[4 * j + e]) : "l"(rr + min(qq + e, a.Tout - 1)));
                    }
                }
            };
            auto prefetch = [&](const float* rr, int q, float* dst) { if (has_res) prefetch_any(rr, q, dst); };
            auto prefetch_acc = [&](const float* rr, int q, float* dst) { if (rr) prefetch_any(rr, q, dst); };
            for (int it = 0; it < my_tiles && ok; ++it) {
                const int buf = it & 1;
                int b, rt, q0;
                decode(it, b, rt, q0);
                const float* rrow = has_res ? a.res + (long long)b * a.res_bs + (long long)co * a.res_cs : nullptr;
                const int cend = half ? a.tstep : min(128, a.tstep);
                fast = vec_ok && (q0 + a.tstep <= a.Tout);
                // Residual / accumulate values are prefetched TWO column groups ahead into three rotating register sets that
                // are never copied: a register move of a pending load's result waits for the load, which made every group
                // pay a full memory latency (r02 timeline: ~1.5 k cycles per 16-column group, the narrow stages' bound).
                float rA[NV], rB[NV], rC[NV], oA[NV], oB[NV], oC[NV];
                float* yrow = a.y + (long long)b * a.y_bs + (long long)co * a.y_cs;
                const float* orow = acc_r ? yrow : nullptr;
                // the first two column groups are requested before waiting for the accumulator
                prefetch(rrow, q0 + cbeg + coff, rA);
                prefetch_acc(orow, q0 + cbeg + coff, oA);
                if (cbeg + 16 < cend) { prefetch(rrow, q0 + cbeg + 16 + coff, rB); prefetch_acc(orow, q0 + cbeg + 16 + coff, oB); }
                ok = mbar_wait(BAR(ACC_FULL + buf), (it >> 1) & 1, a.err);
                if (!ok) break;
                tc_fence_after();
                if (tid == 0) { if (it == 0) TC3_STAMP(16); if (it == 1) TC3_STAMP(18); if (it == 3) TC3_STAMP(20); }
                const uint32_t dlane = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(lq * 32) << 16);
                float bias = a.bias[co];
                if (a.cond) bias += __ldg(a.cond + (long long)b * a.cond_bs + co);
                auto group = [&](int cg, const float* rv, const float* ov, float* rf, float* of) {
                    if (cg + 32 < cend) { prefetch(rrow, q0 + cg + 32 + coff, rf); prefetch_acc(orow, q0 + cg + 32 + coff, of); }
                    {
                    float S[8];
                    const int q = q0 + cg + coff;
                    if constexpr (DIL > 0) {
                        // TMEM reads run at ~64 B/clk per SM: one window [cg, cg + 16 + (GRP-1)*DIL) serves all
                        // groups.  Lane (c, g) needs P[i] = w[i + g*DIL]; the first reduce-scatter step sends
                        // P[i] or P[i+8] and keeps the other, so the g0 half of the shift is folded into those selects
                        constexpr int EXT = (GRP - 1) * DIL;
                        constexpr int EXTN = EXT <= 1 ? 1 : EXT <= 2 ? 2 : EXT <= 4 ? 4 : EXT <= 8 ? 8 : 16;
                        static_assert(EXT <= 16, "grouped epilogue: shift window too wide");
                        uint32_t w[16 + EXTN];
                        tmem_ld_nowait<16>(dlane + (uint32_t)cg, w);
                        tmem_ld_nowait<EXTN>(dlane + (uint32_t)(cg + 16), w + 16);
                        tmem_wait_ld();
                        if constexpr (GRP == 4) {
#pragma unroll
                            for (int i = 0; i < 16 + DIL; ++i) w[i] = g1 ? w[i + 2 * DIL] : w[i];   // t[i] = w[i + 2*g1*DIL]
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float send = __uint_as_float(g0 ? w[i + DIL] : w[i + 8]);
                            const float keep = __uint_as_float(g0 ? w[i + 8 + DIL] : w[i]);
                            S[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                    } else {
                        float P[16];
                        uint32_t l0[16], l1[16];
                        tmem_ld16_nowait(dlane + (uint32_t)cg, l0);
                        tmem_ld16_nowait(dlane + (uint32_t)(cg + a.dil), l1);
                        if constexpr (GRP == 4) {
                            uint32_t l2[16], l3[16];
                            tmem_ld16_nowait(dlane + (uint32_t)(cg + 2 * a.dil), l2);
                            tmem_ld16_nowait(dlane + (uint32_t)(cg + 3 * a.dil), l3);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const uint32_t lo2 = g0 ? l1[i] : l0[i], hi2 = g0 ? l3[i] : l2[i];
                                P[i] = __uint_as_float(g1 ? hi2 : lo2);
                            }
                        } else {
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 16; ++i) P[i] = __uint_as_float(g0 ? l1[i] : l0[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float send = g0 ? P[i] : P[i + 8];
                            S[i] = (g0 ? P[i + 8] : P[i]) + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                    }
                    float R[NV];
                    if constexpr (GRP == 4) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float send = g1 ? S[i] : S[i + 4];
                            R[i] = (g1 ? S[i + 4] : S[i]) + __shfl_xor_sync(0xffffffffu, send, 2);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) R[i] = S[i];
                    }
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        float u = R[i] + bias;
                        if (relu) u = fmaxf(u, 0.f);
                        if (has_res) u += rv[i];
                        u *= scale;
                        if (acc_r) u += ov[i];
                        if (post_div != 1.f) u = u / post_div;
                        R[i] = u;
                    }
                    if (do_store) {
#pragma unroll
                        for (int j = 0; j < NV / 4; ++j) {
                            const int qq = q + 4 * j;
                            if (fast || (vec_ok && qq + 3 < a.Tout)) {
                                *reinterpret_cast<float4*>(yrow + qq) = make_float4(R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) if (qq + e < a.Tout) yrow[qq + e] = R[4 * j + e];
                            }
                        }
                    }
                    }
                };
#pragma unroll 1
                for (int cg = cbeg; cg < cend; cg += 48) {
                    group(cg, rA, oA, rC, oC);
                    if (cg + 16 < cend) group(cg + 16, rB, oB, rA, oA);
                    if (cg + 32 < cend) group(cg + 32, rC, oC, rB, oB);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
                if (tid == 0) { if (it == 0) TC3_STAMP(17); if (it == 1) TC3_STAMP(19); if (it == 3) TC3_STAMP(21); }
            }
            if (tid == 0) TC3_STAMP(22);
        } else
        for (int it = 0; it < my_tiles && ok; ++it) {
            const int buf = it & 1;
            int b, rt, q0;
            decode(it, b, rt, q0);
            ok = mbar_wait(BAR(ACC_FULL + buf), (it >> 1) & 1, a.err);
            if (!ok) break;
            tc_fence_after();
            if (tid == 0) { if (it == 0) TC3_STAMP(16); if (it == 1) TC3_STAMP(18); if (it == 3) TC3_STAMP(20); }
            const int r = rt * MROWS + lq * 32 + lane;             // GEMM row of this lane
            const bool rok = r < a.Rows;
            const int rc = rok ? r : a.Rows - 1;
            const uint32_t dbase = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(lq * 32) << 16) + (uint32_t)(half * 128);
            const int qb = q0 + half * 128;
            float bias = a.bias[rc];
            if (a.cond) bias += __ldg(a.cond + (long long)b * a.cond_bs + rc);
            if (a.gate) {
                // WaveNet gate (wavenet.py:6-13): even lane = tanh argument, odd lane = sigmoid argument of row r/2
                float* yrow = a.y + (long long)b * a.y_bs + (long long)(rc >> 1) * a.y_cs;
                const bool vec_ok = ((a.y_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0);
                const bool even = (lane & 1) == 0;
                for (int cg = 0; cg < 128; cg += 16) {
                    float v[16];
                    tmem_ld16(dbase + (uint32_t)cg, v);
                    const int q = qb + cg;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float u = v[i] + bias;
                        const float act = even ? tanhf(u) : 1.f / (1.f + expf(-u));
                        const float other = __shfl_xor_sync(0xffffffffu, act, 1);
                        v[i] = act * other;
                    }
                    if (rok && even) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int qq = q + 4 * j;
                            if (vec_ok && qq + 3 < a.Tout) *reinterpret_cast<float4*>(yrow + qq) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) if (qq + e < a.Tout) yrow[qq + e] = v[4 * j + e];
                            }
                        }
                    }
                }
            } else if (STAGED && ups == 1 && a.stage && a.split == 0 && qb + 128 <= a.Tout && ((a.y_cs | a.y_bs) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.y) & 15) == 0 &&
                       (!a.res || ((((a.res_cs | a.res_bs) & 3) == 0) && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0))) {
                // ---- staged epilogue (interior tiles of wide layers).  With lane = row every float4 LDG / STG of the
                // direct path touches 32 different cache lines; measured, that L1 wavefront time is not hidden
                // (profiles/r01_tc_grouped_notes.md).  Here each warp transposes its 32 x 16 block through a private
                // shared tile: global accesses are 8 rows x 64 contiguous bytes per instruction (4x fewer wavefronts),
                // the residual arrives by cp.async one column group ahead.
                const int ew = (warp < 4) ? warp : warp - 6;                       // epilogue warp 0..7
                float* tO = reinterpret_cast<float*>(smem + a.stage_off) + ew * 1280;   // [32][20]
                float* tR = tO + 640;
                const int r8 = lane & 7, p4 = lane >> 3;
                const int Rbase = rt * MROWS + lq * 32 + r8;                        // + 8*i
                float* ybase = a.y + (long long)b * a.y_bs + (long long)Rbase * a.y_cs + qb + 4 * p4;
                const float* rbase = (a.res && !(a.dbg & 4)) ? a.res + (long long)b * a.res_bs + (long long)Rbase * a.res_cs + qb + 4 * p4 : nullptr;
                const bool acc_r = a.accum != 0 && !(a.dbg & 4), mpost_r = a.mask_post != 0, do_store = !(a.dbg & 8);
                const float* mrow = a.ymask ? a.ymask + (long long)b * a.ymask_bs : nullptr;
                float* yrow = a.y + (long long)b * a.y_bs + (long long)rc * a.y_cs;   // lane = row view (accumulate loads)
                auto issue_res = [&](int cg) {
                    if (!rbase) return;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (Rbase + 8 * i < a.Rows)
                            cp_async16_zfill(smem_u32(tR + (8 * i + r8) * 20 + 4 * p4), rbase + (long long)(8 * i) * a.res_cs + cg, 16u);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                auto prefetch_acc = [&](int cg, float* o_) {
                    if (!acc_r || !rok) return;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o_[4 * j]), "=f"(o_[4 * j + 1]), "=f"(o_[4 * j + 2]), "=f"(o_[4 * j + 3]) : "l"(yrow + qb + cg + 4 * j));
                };
                issue_res(0);
#pragma unroll 1
                for (int cg = 0; cg < 128; cg += 16) {
                    float v[16], ov[16], rv[16];
                    prefetch_acc(cg, ov);           // accumulate-into-destination (two layers per stage): same-group load
                    if (rbase) {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t4 = *reinterpret_cast<const float4*>(tR + lane * 20 + 4 * j);
                            rv[4 * j] = t4.x; rv[4 * j + 1] = t4.y; rv[4 * j + 2] = t4.z; rv[4 * j + 3] = t4.w;
                        }
                        __syncwarp();
                        if (cg + 16 < 128) issue_res(cg + 16);
                    }
                    tmem_ld16(dbase + (uint32_t)cg, v);
                    const int q = qb + cg;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float u = v[i] + bias;
                        if (a.relu) u = fmaxf(u, 0.f);
                        const float mk = mrow ? __ldg(mrow + q + i) : 1.f;
                        if (a.mask_pre) u *= mk;
                        if (rbase) u += rv[i];
                        u *= a.scale;
                        if (acc_r) u += ov[i];
                        if (a.post_div != 1.f) u = u / a.post_div;
                        if (mpost_r) u *= mk;
                        v[i] = u;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(tO + lane * 20 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 o4 = *reinterpret_cast<const float4*>(tO + (8 * i + r8) * 20 + 4 * p4);
                        if (do_store && Rbase + 8 * i < a.Rows) *reinterpret_cast<float4*>(ybase + (long long)(8 * i) * a.y_cs + cg) = o4;
                    }
                    __syncwarp();
                }
            } else if (ups == 1) {
                float* yrow = a.y + (long long)b * a.y_bs + (long long)rc * a.y_cs;
                bool acc_r = a.accum != 0, mpost_r = a.mask_post != 0;
                if (a.split > 0) {          // WaveNet res/skip rows (wavenet.py:108-113)
                    if (rc < a.split) { acc_r = true; mpost_r = true; }
                    else { yrow = a.y2 + (long long)b * a.y2_bs + (long long)(rc - a.split) * a.y2_cs; acc_r = a.accum2 != 0; mpost_r = false; }
                }
                const float* rrow = a.res ? a.res + (long long)b * a.res_bs + (long long)rc * a.res_cs : nullptr;
                const float* mrow = a.ymask ? a.ymask + (long long)b * a.ymask_bs : nullptr;
                const int ycs_eff = (a.split > 0 && rc >= a.split) ? a.y2_cs : a.y_cs;
                const bool vec_ok = ((ycs_eff & 3) == 0) && (!a.res || (a.res_cs & 3) == 0) &&
                                    ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0) &&
                                    (!rrow || (reinterpret_cast<uintptr_t>(rrow) & 15) == 0);
                // order-enforced software pipeline: the (volatile) loads of group j+1 are issued before the (volatile)
                // TMEM load of group j, into the OTHER of two register sets -- never copied (a move of a pending load's
                // result waits for the load and would serialise the groups).  float4 per lane along time (lane = one row).
                float rA[16], oA[16], rB[16], oB[16];
                auto prefetch = [&](int cg, float* r_, float* o_) {
                    const int q = qb + cg;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int qq = q + 4 * j;
                        if (vec_ok && qq + 3 < a.Tout) {
                            if (rrow) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r_[4 * j]), "=f"(r_[4 * j + 1]), "=f"(r_[4 * j + 2]), "=f"(r_[4 * j + 3]) : "l"(rrow + qq));
                            if (acc_r) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o_[4 * j]), "=f"(o_[4 * j + 1]), "=f"(o_[4 * j + 2]), "=f"(o_[4 * j + 3]) : "l"(yrow + qq));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int qe = min(qq + e, a.Tout - 1);
                                if (rrow) asm volatile("ld.global.f32 %0, [%1];" : "=f"(r_[4 * j + e]) : "l"(rrow + qe));
                                if (acc_r) asm volatile("ld.global.f32 %0, [%1];" : "=f"(o_[4 * j + e]) : "l"(yrow + qe));
                            }
                        }
                    }
                };
                const bool ld_ok = rok && !(a.dbg & 4), st_ok = rok && !(a.dbg & 8);
                auto group = [&](int cg, const float* rv, const float* ov, float* rn, float* on) {
                    float v[16];
                    if (ld_ok && cg + 16 < 128) prefetch(cg + 16, rn, on);
                    tmem_ld16(dbase + (uint32_t)cg, v);
                    const int q = qb + cg;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float u = v[i] + bias;
                        if (a.relu) u = fmaxf(u, 0.f);
                        const float mk = mrow ? __ldg(mrow + min(q + i, a.Tout - 1)) : 1.f;
                        if (a.mask_pre) u *= mk;
                        if (a.res) u += rv[i];
                        u *= a.scale;
                        if (acc_r) u += ov[i];
                        if (a.post_div != 1.f) u = u / a.post_div;
                        if (mpost_r) u *= mk;
                        v[i] = u;
                    }
                    if (st_ok) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int qq = q + 4 * j;
                            if (vec_ok && qq + 3 < a.Tout) {
                                *reinterpret_cast<float4*>(yrow + qq) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) if (qq + e < a.Tout) yrow[qq + e] = v[4 * j + e];
                            }
                        }
                    }
                };
                if (ld_ok) prefetch(0, rA, oA);
#pragma unroll 1
                for (int cg = 0; cg < 128; cg += 32) {
                    group(cg, rA, oA, rB, oB);
                    group(cg + 16, rB, oB, rA, oA);
                }
            } else {
                // polyphase store: row r = co*ups + ph, column q -> y[co][q*ups + ph]; a warp's 32 lanes cover whole
                // groups of `ups` phases, i.e. contiguous runs of `ups` output samples per channel
                const int co = rc / ups, ph = rc - co * ups;
                float* yrow = a.y + (long long)b * a.y_bs + (long long)co * a.y_cs + ph;
                for (int cg = 0; cg < 128; cg += 16) {
                    float v[16];
                    tmem_ld16(dbase + (uint32_t)cg, v);
                    if (!rok) continue;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int q = qb + cg + i;
                        float u = v[i] + bias;
                        if (a.relu) u = fmaxf(u, 0.f);
                        const long long t = (long long)q * ups;
                        if (q < a.Tq && t + ph < a.Tout) yrow[t] = u;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));   // one arrival per epilogue warp
            if (tid == 0) { if (it == 0) TC3_STAMP(17); if (it == 1) TC3_STAMP(19); if (it == 3) TC3_STAMP(21); }
        }
        if (tid == 0) TC3_STAMP(22);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

__global__ void __launch_bounds__(NTHREADS2, 1) conv1d_tc3_kernel(const Tc3Args a) { tc3_body<1, 0>(a); }
__global__ void __launch_bounds__(NTHREADS2, 1) conv1d_tc3s_kernel(const Tc3Args a) { tc3_body<1, 0, true>(a); }
template <int GRP, int DIL>
__global__ void __launch_bounds__(NTHREADS2, 1) conv1d_tc3g_kernel(const Tc3Args a) { tc3_body<GRP, DIL>(a); }

typedef void (*Tc3Kernel)(const Tc3Args);
// grouped kernel for (tap groups, dilation): dilations 1 / 3 / 5 (the HiFiGAN resblocks) are specialised
static inline Tc3Kernel grouped_kernel(int grp, int dil) {
    if (grp == 2) return dil == 1 ? conv1d_tc3g_kernel<2, 1> : dil == 3 ? conv1d_tc3g_kernel<2, 3> : dil == 5 ? conv1d_tc3g_kernel<2, 5> : conv1d_tc3g_kernel<2, 0>;
    return dil == 1 ? conv1d_tc3g_kernel<4, 1> : dil == 3 ? conv1d_tc3g_kernel<4, 3> : dil == 5 ? conv1d_tc3g_kernel<4, 5> : conv1d_tc3g_kernel<4, 0>;
}

}  // namespace tc3
}  // namespace b200tts
