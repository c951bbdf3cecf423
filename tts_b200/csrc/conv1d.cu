// Generic fused conv1d as an FP32 implicit GEMM on the sm_100a FMA pipe (packed FFMA2).
//
// This one kernel carries >99% of the VITS+HiFiGAN inference FLOPs: the HiFiGAN MRF convs
// (reference: TTS/vocoder/models/hifigan_generator.py:84-99,236-265), the polyphase form of its
// ConvTranspose1d upsamplers (:207-218), the WaveNet k5 / 1x1 convs of the flow
// (TTS/tts/layers/generic/wavenet.py:94-115) and the 1x1 / k3 convs of the text encoder
// (TTS/tts/layers/glow_tts/transformer.py:109-121,290-295).
//
// Layout: activations stay in the reference's [B, C, T] layout (T contiguous).  A CTA owns a
// [CO_T rows] x [T_T time] output tile.  Input channels are streamed in chunks of 8 through a
// double-buffered shared-memory window [8][T_T + (K-1)*dil] (coalesced loads along T, prologue
// -- mask, leaky-relu -- applied once on the way in) next to the chunk's weights
// [8][K][CO_T] (cp.async).  Each lane owns TJ time steps strided by 32 (conflict-free LDS.32,
// tap shifts are plain address offsets) and CJ=16 consecutive rows (warp-uniform weight LDS.128
// broadcasts), accumulating row pairs with fma.rn.f32x2 (SASS FFMA2, x as broadcast scalar).
// Bias / conditioning / gate / residual / mask / MRF-accumulate epilogues are fused.
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace b200tts {

// ------------------------------------------------------------------ error plumbing / counters
static thread_local char g_err[1024] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

int upload(float** dst, const float* src, size_t n) {
    *dst = nullptr;
    if (n == 0) return 0;
    B200_CUDA_OK(cudaMalloc((void**)dst, n * sizeof(float)));
    B200_CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------ device helpers
constexpr int CI_C = 8;  // input channels per pipeline stage

typedef unsigned long long u64;

__device__ __forceinline__ void ffma2(u64& d, u64 a, float x) {
    u64 xx;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(xx));
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct ConvKArgs {
    const float* x; long long x_bs; int x_cs; int Tin;
    const float* xmask; long long xmask_bs; float in_slope;
    const float* w; const float* bias; const float* cond; long long cond_bs;
    int Cin, CinPad, K, dil, pad, Rows, ups, Tq;
    float* y; long long y_bs; int y_cs; int Tout;
    const float* res; long long res_bs; int res_cs;
    const float* ymask; long long ymask_bs;
    float* y2; long long y2_bs; int y2_cs; int split;
    float scale; float post_div; int act; int flags;
    int XS;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_TANH) return tanhf(v);
    return v;
}

// one output element through the generic epilogue (see common.cuh)
__device__ __forceinline__ void epilogue_store(const ConvKArgs& a, int b, int r, int q, float v) {
    if (r >= a.Rows) return;
    v += a.bias[r];
    if (a.cond) v += __ldg(a.cond + b * a.cond_bs + r);
    v = apply_act(v, a.act);
    int ch = r, t = q;
    if (a.ups > 1) { ch = r / a.ups; t = q * a.ups + (r - ch * a.ups); }
    if (t >= a.Tout) return;
    const float m = a.ymask ? __ldg(a.ymask + b * a.ymask_bs + t) : 1.f;
    float* yp; bool accum, mask_post;
    if (a.flags & EPI_SPLIT) {
        if (ch < a.split) { yp = a.y + b * a.y_bs + (long long)ch * a.y_cs + t; accum = true; mask_post = true; }
        else { yp = a.y2 + b * a.y2_bs + (long long)(ch - a.split) * a.y2_cs + t; accum = (a.flags & EPI_ACCUM2) != 0; mask_post = false; }
    } else {
        yp = a.y + b * a.y_bs + (long long)ch * a.y_cs + t;
        accum = (a.flags & EPI_ACCUM) != 0;
        mask_post = (a.flags & EPI_MASK_POST) != 0;
    }
    if (a.flags & EPI_MASK_PRE) v *= m;
    if (a.res) v += a.res[b * a.res_bs + (long long)ch * a.res_cs + t];  // may alias y (in-place residual)
    v *= a.scale;
    if (accum) v += *yp;
    if (a.post_div != 1.f) v = v / a.post_div;
    if (mask_post) v *= m;
    *yp = v;
}

template <int CJ, int TJ, int WCO, int WT, int MINB>
__global__ void __launch_bounds__(32 * WCO * WT, MINB) conv1d_kernel(const ConvKArgs a) {
    constexpr int CO_T = CJ * WCO, T_T = 32 * TJ * WT, NT = 32 * WCO * WT;
    extern __shared__ __align__(16) float smem[];
    const int XS = a.XS;
    const int wchunk = CI_C * a.K * CO_T;
    float* xs0 = smem;
    float* ws0 = smem + 2 * CI_C * XS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wco = warp % WCO, wt = warp / WCO;
    const int b = blockIdx.z, tile_co = blockIdx.y;
    const int q0 = blockIdx.x * T_T;
    const int tin0 = q0 - a.pad;
    const float* xb = a.x + b * a.x_bs;
    const float* mb = a.xmask ? a.xmask + b * a.xmask_bs : nullptr;
    const float* wg = a.w + (size_t)tile_co * a.CinPad * a.K * CO_T;
    const int nchunks = a.CinPad / CI_C;
    const float slope = a.in_slope;

    auto load_chunk = [&](int chunk, int buf) {
        const float* src = wg + (size_t)chunk * wchunk;
        float* dst = ws0 + buf * wchunk;
        for (int i = tid * 4; i < wchunk; i += NT * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
        float* xd = xs0 + buf * CI_C * XS;
        const int c0 = chunk * CI_C;
        for (int i = tid; i < XS; i += NT) {
            const int t = tin0 + i;
            const bool tok = (t >= 0) && (t < a.Tin);
            float m = 1.f;
            if (tok && mb) m = __ldg(mb + t);
            float v[CI_C];
#pragma unroll
            for (int ci = 0; ci < CI_C; ++ci) {
                v[ci] = 0.f;
                if (tok && (c0 + ci) < a.Cin) v[ci] = __ldg(xb + (long long)(c0 + ci) * a.x_cs + t);
            }
#pragma unroll
            for (int ci = 0; ci < CI_C; ++ci) {
                float u = v[ci] * m;
                u = u > 0.f ? u : u * slope;
                xd[ci * XS + i] = u;
            }
        }
    };

    u64 acc[CJ / 2][TJ];
#pragma unroll
    for (int p = 0; p < CJ / 2; ++p)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[p][j] = 0ull;

    load_chunk(0, 0);
    cp_async_wait_all();
    __syncthreads();

    const int K = a.K, dil = a.dil;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) load_chunk(ch + 1, buf ^ 1);
        const float* xr = xs0 + buf * CI_C * XS + wt * 32 * TJ + lane;
        const float* wr = ws0 + buf * wchunk + wco * CJ;
#pragma unroll 1
        for (int ci = 0; ci < CI_C; ++ci) {
            const float* xp0 = xr + ci * XS;
            const float* wp0 = wr + ci * K * CO_T;
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                const float* xp = xp0 + k * dil;
                const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(wp0 + k * CO_T);
                float xv[TJ];
#pragma unroll
                for (int j = 0; j < TJ; ++j) xv[j] = xp[32 * j];
                u64 wv[CJ / 2];
#pragma unroll
                for (int p = 0; p < CJ / 4; ++p) {
                    const ulonglong2 t2 = wp[p];
                    wv[2 * p] = t2.x;
                    wv[2 * p + 1] = t2.y;
                }
#pragma unroll
                for (int j = 0; j < TJ; ++j)
#pragma unroll
                    for (int p = 0; p < CJ / 2; ++p) ffma2(acc[p][j], wv[p], xv[j]);
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // ---------------------------------------------------------------- epilogue
    const int row_base = tile_co * CO_T + wco * CJ;
    const int qb = q0 + wt * 32 * TJ + lane;
    if (a.flags & EPI_GATE) {
        // rows (2p, 2p+1) = (tanh half, sigmoid half) of output row row_base/2 + p
#pragma unroll
        for (int p = 0; p < CJ / 2; ++p) {
            const int r0 = row_base + 2 * p;
            if (r0 + 1 >= a.Rows) continue;
            float b0 = a.bias[r0], b1 = a.bias[r0 + 1];
            if (a.cond) { b0 += __ldg(a.cond + b * a.cond_bs + r0); b1 += __ldg(a.cond + b * a.cond_bs + r0 + 1); }
            float* yrow = a.y + b * a.y_bs + (long long)(r0 >> 1) * a.y_cs;
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const int q = qb + 32 * j;
                if (q >= a.Tout) continue;
                float v0, v1;
                unpack2(acc[p][j], v0, v1);
                v0 += b0; v1 += b1;
                yrow[q] = tanhf(v0) * (1.f / (1.f + expf(-v1)));
            }
        }
        return;
    }
#pragma unroll
    for (int p = 0; p < CJ / 2; ++p) {
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const int q = qb + 32 * j;
            if (q >= a.Tq) continue;
            float v0, v1;
            unpack2(acc[p][j], v0, v1);
            epilogue_store(a, b, row_base + 2 * p, q, v0);
            epilogue_store(a, b, row_base + 2 * p + 1, q, v1);
        }
    }
}

// ------------------------------------------------------------------ host: packing
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

void free_conv(ConvLayer& L) {
    if (L.w) cudaFree(L.w);
    if (L.bias) cudaFree(L.bias);
    L.w = L.bias = nullptr;
}

// Wl(r, ci, k): logical weights already expressed as a correlation-form conv with `rows` GEMM rows
static int pack_rows(ConvLayer& L, const std::vector<float>& Wl, const std::vector<float>& bl, int rows, int Cin,
                     int K) {
    L.Rows = rows;
    L.co_tile = rows >= 64 ? 64 : 32;
    L.RowsPad = round_up(rows, L.co_tile);
    L.Cin = Cin;
    L.CinPad = round_up(Cin, CI_C);
    L.K = K;
    const int T = L.co_tile, ntile = L.RowsPad / T;
    std::vector<float> P((size_t)ntile * L.CinPad * K * T, 0.f);
    for (int r = 0; r < rows; ++r) {
        const int tile = r / T, col = r % T;
        for (int ci = 0; ci < Cin; ++ci)
            for (int k = 0; k < K; ++k)
                P[(((size_t)tile * L.CinPad + ci) * K + k) * T + col] = Wl[((size_t)r * Cin + ci) * K + k];
    }
    std::vector<float> bp(L.RowsPad, 0.f);
    for (int r = 0; r < rows; ++r) bp[r] = bl[r];
    if (upload(&L.w, P.data(), P.size())) return 2;
    if (upload(&L.bias, bp.data(), bp.size())) return 2;
    return 0;
}

int pack_conv(ConvLayer& L, const float* w, const float* bias, int Cout, int Cin, int K, int dil, int pad,
              int gate_half, const int* in_perm, const int* out_perm) {
    B200_REQUIRE(w != nullptr && Cout > 0 && Cin > 0 && K > 0, "pack_conv: bad arguments");
    B200_REQUIRE(gate_half == 0 || 2 * gate_half == Cout, "pack_conv: gate_half must be Cout/2");
    std::vector<float> Wl((size_t)Cout * Cin * K), bl(Cout, 0.f);
    for (int r = 0; r < Cout; ++r) {
        int rn = r;
        if (gate_half > 0) rn = (r < gate_half) ? 2 * r : 2 * (r - gate_half) + 1;
        if (out_perm) rn = out_perm[r];
        for (int ci = 0; ci < Cin; ++ci) {
            const int cn = in_perm ? in_perm[ci] : ci;
            for (int k = 0; k < K; ++k) Wl[((size_t)rn * Cin + cn) * K + k] = w[((size_t)r * Cin + ci) * K + k];
        }
        if (bias) bl[rn] = bias[r];
    }
    L.dil = dil;
    L.pad = pad;
    L.ups = 1;
    return pack_rows(L, Wl, bl, Cout, Cin, K);
}

int pack_conv_transpose(ConvLayer& L, const float* w, const float* bias, int Cin, int Cout, int Kt, int s, int p) {
    // y[co, q*s+ph] = sum_ci sum_m x[ci, q-m] * w[ci, co, ph + p + m*s]   (0 <= ph+p+m*s < Kt)
    B200_REQUIRE(w != nullptr && s >= 1 && Kt >= 1, "pack_conv_transpose: bad arguments");
    auto floordiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
    const int m_lo = -floordiv(s - 1 + p, s);  // ceil(-(s-1+p)/s)
    const int m_hi = floordiv(Kt - 1 - p, s);
    const int K = m_hi - m_lo + 1;
    B200_REQUIRE(K >= 1, "pack_conv_transpose: empty tap range");
    const int rows = Cout * s;
    std::vector<float> Wl((size_t)rows * Cin * K, 0.f), bl(rows, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int ph = 0; ph < s; ++ph) {
            const int r = co * s + ph;
            if (bias) bl[r] = bias[co];
            for (int kk = 0; kk < K; ++kk) {
                const int m = m_hi - kk;
                const int kt = ph + p + m * s;
                if (kt < 0 || kt >= Kt) continue;
                for (int ci = 0; ci < Cin; ++ci)
                    Wl[((size_t)r * Cin + ci) * K + kk] = w[((size_t)ci * Cout + co) * Kt + kt];
            }
        }
    L.dil = 1;
    L.pad = m_hi;
    L.ups = s;
    L.tr_kernel = Kt;
    L.tr_pad = p;
    return pack_rows(L, Wl, bl, rows, Cin, K);
}

// ------------------------------------------------------------------ host: launch
template <int CJ, int TJ, int WCO, int WT, int MINB = 2>
static int launch_variant(const ConvKArgs& ka, int B, int RowsPad, cudaStream_t st) {
    constexpr int CO_T = CJ * WCO, T_T = 32 * TJ * WT, NT = 32 * WCO * WT;
    ConvKArgs a = ka;
    a.XS = round_up(T_T + (a.K - 1) * a.dil, 4);
    const size_t smem = (size_t)(2 * CI_C * a.XS + 2 * CI_C * a.K * CO_T) * sizeof(float);
    B200_REQUIRE(smem <= 227 * 1024, "conv1d: K=%d dil=%d needs %zu B of shared memory", a.K, a.dil, smem);
    static bool attr_done = false;
    if (!attr_done) {
        B200_CUDA_OK(cudaFuncSetAttribute(conv1d_kernel<CJ, TJ, WCO, WT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
        attr_done = true;
    }
    dim3 grid((a.Tq + T_T - 1) / T_T, RowsPad / CO_T, B);
    B200_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv1d: grid too large");
    conv1d_kernel<CJ, TJ, WCO, WT, MINB><<<grid, NT, smem, st>>>(a);
    count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_conv(const ConvLayer& L, const ConvIO& io, cudaStream_t st) {
    B200_REQUIRE(L.w && io.x && io.y, "launch_conv: null tensor");
    ConvKArgs a;
    memset(&a, 0, sizeof(a));
    a.x = io.x; a.x_bs = io.x_bs; a.x_cs = io.x_cs; a.Tin = io.Tin;
    a.xmask = io.xmask; a.xmask_bs = io.xmask_bs; a.in_slope = io.in_slope;
    a.w = L.w; a.bias = L.bias; a.cond = io.cond; a.cond_bs = io.cond_bs;
    a.Cin = L.Cin; a.CinPad = L.CinPad; a.K = L.K; a.dil = L.dil; a.pad = L.pad; a.Rows = L.Rows; a.ups = L.ups;
    a.y = io.y; a.y_bs = io.y_bs; a.y_cs = io.y_cs; a.Tout = io.Tout;
    a.Tq = (L.ups > 1) ? (io.Tout + L.ups - 1) / L.ups : io.Tout;
    a.res = io.res; a.res_bs = io.res_bs; a.res_cs = io.res_cs;
    a.ymask = io.ymask; a.ymask_bs = io.ymask_bs;
    a.y2 = io.y2; a.y2_bs = io.y2_bs; a.y2_cs = io.y2_cs; a.split = io.split;
    a.scale = io.scale; a.post_div = io.post_div; a.act = io.act; a.flags = io.flags;
    if (a.Tq <= 0 || io.B <= 0) return 0;
    B200_REQUIRE(!(a.flags & (EPI_MASK_PRE | EPI_MASK_POST | EPI_SPLIT)) || io.ymask,
                 "launch_conv: masked/split epilogue needs ymask");
    B200_REQUIRE(!(a.flags & EPI_SPLIT) || io.y2, "launch_conv: split epilogue needs y2");
    const bool small_t = a.Tq <= 128;
    static int variant = -1;  // developer knob: B200TTS_CONV_VARIANT (0 = default)
    if (variant < 0) { const char* e = getenv("B200TTS_CONV_VARIANT"); variant = e ? atoi(e) : 0; }
    if (L.co_tile == 64) {
        if (small_t) return launch_variant<16, 2, 4, 1>(a, io.B, L.RowsPad, st);
        if (variant == 1) return launch_variant<16, 8, 4, 1, 3>(a, io.B, L.RowsPad, st);
        if (variant == 2) return launch_variant<16, 8, 4, 2, 1>(a, io.B, L.RowsPad, st);
        if (variant == 3) return launch_variant<16, 4, 4, 2, 2>(a, io.B, L.RowsPad, st);
        return launch_variant<16, 8, 4, 1>(a, io.B, L.RowsPad, st);
    }
    if (small_t) return launch_variant<16, 2, 2, 2>(a, io.B, L.RowsPad, st);
    return launch_variant<16, 8, 2, 2>(a, io.B, L.RowsPad, st);
}

}  // namespace b200tts
