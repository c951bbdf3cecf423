// The hand-off kernels either side of the HiFiGAN generator on the reference's public path:
//
//  vocoder_input_kernel : what Synthesizer.tts does between the TTS model and a standalone vocoder
//      (TTS/utils/synthesizer.py:412-429): tts_ap.denormalize -> vocoder_ap.normalize
//      (TTS/utils/audio/processor.py:259-337, all four branches: mean-var, symmetric, asymmetric, clip) ->
//      interpolate_vocoder_input (TTS/vocoder/utils/generic_utils.py:11-29: bilinear, align_corners=False,
//      recompute_scale_factor=True, scale [1, sr_voc/sr_tts]) -> the replicate padding of HifiganGenerator.inference
//      (TTS/vocoder/models/hifigan_generator.py:281) -- fused into ONE pass that writes conv_pre's input with a
//      16-byte aligned row pitch.  The reference does this in numpy on the host, one sentence at a time.
//  absmax / to_int16    : save_wav's peak normalisation, wav * (32767 / max(0.01, max|wav|)) truncated to int16
//      (TTS/utils/audio/numpy_transforms.py:439-441), on the device: the conv_post kernel already folds max|wav| into a
//      device word while it stores the waveform (conv1d.cu), to_int16 scales and converts.
#include "engines.cuh"

namespace b200tts {

namespace {

struct NormParams {          // one AudioProcessor's normalisation settings
    int signal_norm, symmetric_norm, clip_norm, has_scaler;
    float max_norm, min_level_db, ref_level_db;
    const float* mean;       // [C] (mean-var scaler) or null
    const float* scale;      // [C]
};

__device__ __forceinline__ float denorm_one(const NormParams& p, float s, int c) {
    if (!p.signal_norm) return s;
    if (p.has_scaler) return __fadd_rn(__fmul_rn(s, p.scale[c]), p.mean[c]);     // StandardScaler.inverse_transform
    if (p.symmetric_norm) {
        if (p.clip_norm) s = fminf(fmaxf(s, -p.max_norm), p.max_norm);
        // ((S + max_norm) * -min_level_db / (2 * max_norm)) + min_level_db   evaluated left to right like numpy
        s = __fadd_rn(__fdiv_rn(__fmul_rn(__fadd_rn(s, p.max_norm), -p.min_level_db), __fmul_rn(2.f, p.max_norm)), p.min_level_db);
        return __fadd_rn(s, p.ref_level_db);
    }
    if (p.clip_norm) s = fminf(fmaxf(s, 0.f), p.max_norm);
    s = __fadd_rn(__fdiv_rn(__fmul_rn(s, -p.min_level_db), p.max_norm), p.min_level_db);
    return __fadd_rn(s, p.ref_level_db);
}

__device__ __forceinline__ float norm_one(const NormParams& p, float s, int c) {
    if (!p.signal_norm) return s;
    if (p.has_scaler) return __fdiv_rn(__fsub_rn(s, p.mean[c]), p.scale[c]);     // StandardScaler.transform
    s = __fsub_rn(s, p.ref_level_db);
    float n = __fdiv_rn(__fsub_rn(s, p.min_level_db), -p.min_level_db);
    if (p.symmetric_norm) {
        n = __fsub_rn(__fmul_rn(__fmul_rn(2.f, p.max_norm), n), p.max_norm);
        if (p.clip_norm) n = fminf(fmaxf(n, -p.max_norm), p.max_norm);
        return n;
    }
    n = __fmul_rn(p.max_norm, n);
    if (p.clip_norm) n = fminf(fmaxf(n, 0.f), p.max_norm);
    return n;
}

// out[b, c, j] for j in [0, Tmid + 2*pad): source column jj = clamp(j - pad, 0, Tmid - 1) of the interpolated
// spectrogram; interpolation (Tmid != T): src = (jj + 0.5) * (T / Tmid) - 0.5 clamped at 0, neighbours t0, min(t0+1, T-1).
__global__ void vocoder_input_kernel(const float* __restrict__ x, int x_bs, int x_cs, int x_ts, NormParams dn, NormParams nm,
                                     int C, int T, int Tmid, float rscale, int pad, float* __restrict__ y, int Tout,
                                     int y_pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Tout) return;
    const int jj = min(max(j - pad, 0), Tmid - 1);
    int t0 = jj, t1 = jj;
    float l1 = 0.f;
    if (Tmid != T) {
        float src = fmaf(rscale, (float)jj + 0.5f, -0.5f);
        src = src < 0.f ? 0.f : src;
        t0 = (int)src;
        t1 = t0 + ((t0 < T - 1) ? 1 : 0);
        l1 = src - (float)t0;
    }
    const float l0 = 1.f - l1;
    const int b = blockIdx.z;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const float* xr = x + (long long)b * x_bs + (long long)c * x_cs;
        const float v0 = norm_one(nm, denorm_one(dn, xr[(long long)t0 * x_ts], c), c);
        float v = v0;
        if (Tmid != T) {
            const float v1 = norm_one(nm, denorm_one(dn, xr[(long long)t1 * x_ts], c), c);
            v = __fadd_rn(__fmul_rn(l0, v0), __fmul_rn(l1, v1));      // the order upsample_bilinear2d uses: w0*x0 + w1*x1
        }
        y[((long long)b * C + c) * y_pitch + j] = v;
    }
}

__global__ void absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}

// wav * (32767 / max(0.01, peak)) truncated toward zero (numpy astype(int16) of an in-range float)
__global__ void to_int16_kernel(const float* __restrict__ x, long long n, const unsigned* __restrict__ peak_bits,
                                short* __restrict__ out) {
    const float peak = __uint_as_float(*peak_bits);
    const float s = __fdiv_rn(32767.f, fmaxf(0.01f, peak));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (short)__float2int_rz(__fmul_rn(x[i], s));
}

static NormParams to_params(const b200tts_audio_norm& a) {
    NormParams p;
    p.signal_norm = a.signal_norm; p.symmetric_norm = a.symmetric_norm; p.clip_norm = a.clip_norm;
    p.has_scaler = (a.scaler_mean && a.scaler_scale) ? 1 : 0;
    p.max_norm = a.max_norm; p.min_level_db = a.min_level_db; p.ref_level_db = a.ref_level_db;
    p.mean = a.scaler_mean; p.scale = a.scaler_scale;
    return p;
}

}  // namespace

int vocoder_input_len(int T, float scale_factor, int pad) {
    const int Tmid = (scale_factor == 1.f) ? T : (int)floor((double)T * (double)scale_factor);
    return Tmid + 2 * pad;
}

int launch_vocoder_input(const float* x, long long x_bs, int x_cs, int x_ts, int B, int C, int T,
                         const b200tts_audio_norm* denorm, const b200tts_audio_norm* norm, float scale_factor, int pad,
                         float* y, int y_pitch, cudaStream_t st) {
    B200_REQUIRE(x && y && denorm && norm, "vocoder_input: null pointer");
    B200_REQUIRE(scale_factor > 0.f && pad >= 0, "vocoder_input: bad scale_factor / padding");
    if (B == 0 || C == 0 || T == 0) return 0;
    const int Tmid = (scale_factor == 1.f) ? T : (int)floor((double)T * (double)scale_factor);
    B200_REQUIRE(Tmid >= 1, "vocoder_input: scale_factor %f leaves no frames", (double)scale_factor);
    const int Tout = Tmid + 2 * pad;
    B200_REQUIRE(y_pitch >= Tout, "vocoder_input: output pitch %d < %d columns", y_pitch, Tout);
    dim3 grid((Tout + 127) / 128, C < 65535 ? C : 65535, B);
    B200_REQUIRE(B <= 65535, "vocoder_input: batch too large");
    // recompute_scale_factor=True: coordinates use the size ratio, not the requested factor
    const float rscale = (float)((double)T / (double)Tmid);
    vocoder_input_kernel<<<grid, 128, 0, st>>>(x, (int)x_bs, x_cs, x_ts, to_params(*denorm), to_params(*norm), C, T, Tmid,
                                               rscale, pad, y, Tout, y_pitch);
    count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_absmax(const float* x, long long n, unsigned* peak_bits, cudaStream_t st) {
    B200_REQUIRE(peak_bits && (x || n == 0), "absmax: null pointer");
    if (n == 0) return 0;
    const int blocks = (int)std::min<long long>((n + 1023) / 1024, 148 * 8);
    absmax_kernel<<<blocks, 256, 0, st>>>(x, n, peak_bits);
    count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_to_int16(const float* x, long long n, const unsigned* peak_bits, short* out, cudaStream_t st) {
    B200_REQUIRE(peak_bits && (n == 0 || (x && out)), "to_int16: null pointer");
    if (n == 0) return 0;
    const int blocks = (int)std::min<long long>((n + 1023) / 1024, 148 * 8);
    to_int16_kernel<<<blocks, 256, 0, st>>>(x, n, peak_bits, out);
    count_launch();
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200tts
