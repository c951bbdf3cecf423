/* tts_b200 -- C ABI of the B200-native (sm_100a) VITS + HiFiGAN inference hot path.
 *
 * The reference (coqui-ai/TTS v0.22.0) has no FFI for this path: its "operator API" is Python
 * classes + state_dict.  Each entry point below replaces the body of one
 * reference call; the Python mirror in tts_b200/ binds them with ctypes and keeps the reference's
 * class / function signatures.  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (1 = bad argument, 2 = CUDA error);
 *     b200tts_last_error() returns a thread-local message.  Nothing throws across the ABI.
 *   - tensors are raw DEVICE pointers, fp32 unless stated, in the reference's layouts
 *     ([B, C, T] with T contiguous); the caller owns every buffer.
 *   - *_create() takes HOST pointers to fp32 weights in PyTorch layout (weight-norm already
 *     folded: w = g * v / ||v||), packs them for the kernels and uploads them once.
 *   - scratch comes from a caller workspace sized by *_workspace_bytes(); handles are immutable
 *     after creation and may be shared by concurrent streams (each with its own workspace).
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef TTS_B200_H
#define TTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* b200tts_last_error(void);
/* number of kernels launched by this library in this process (bench.py "gpu_launches") */
unsigned long long b200tts_launch_count(void);
int b200tts_version(void);

/* ---- monotonic alignment search ------------------------------------------------------------
 * Replaces maximum_path_c / maximum_path_each, TTS/tts/utils/monotonic_align/core.pyx:11-47
 * (called through TTS/tts/utils/helpers.py:172-194 from Vits.forward_mas, vits.py:919).
 * value [B,Tx,Ty] f32 (NOT modified, unlike the reference's in-place DP); mask [B,Tx,Ty] f32 or
 * NULL (when given, value*mask is formed on load exactly like helpers.py:184); t_x,t_y int32 [B]
 * device pointers; path [B,Tx,Ty] written in full (zeros and ones) as int32 (path_is_f32 = 0,
 * the reference dtype, core.pyx:11) or float32 (path_is_f32 = 1, the dtype helpers.py:194 casts to).
 */
size_t b200tts_mas_workspace_bytes(int B, int Tx, int Ty);
int b200tts_mas(const float* value, const float* mask, const int32_t* t_x, const int32_t* t_y, int B, int Tx,
                int Ty, void* path, int path_is_f32, void* workspace, size_t workspace_bytes, void* stream);

/* ---- HiFiGAN generator -----------------------------------------------------------------------
 * Replaces HifiganGenerator.forward, TTS/vocoder/models/hifigan_generator.py:236-265
 * (ctor arguments :163-178).
 * weights (host pointers, PyTorch layouts), in this order:
 *   conv_pre.w [C0,Cin,7], conv_pre.b [C0]
 *   cond_layer.w [C0,cond,1], cond_layer.b [C0]                      (only if cond_channels > 0)
 *   for each upsample stage s:
 *     ups[s].w [Cs_in, Cs_in/2, k] (ConvTranspose1d layout), ups[s].b
 *     for each resblock kernel j, for each dilation n:
 *       type "1": convs1[n].w, convs1[n].b, convs2[n].w, convs2[n].b
 *       type "2": convs[n].w,  convs[n].b
 *   conv_post.w [Cout, Clast, 7], conv_post.b (NULL pointer when conv_post_bias=False)
 */
typedef struct {
    int in_channels;
    int out_channels;
    int upsample_initial_channel;
    int cond_channels;
    int resblock_type; /* 1 or 2 */
    int num_upsamples;
    int upsample_factors[8];
    int upsample_kernel_sizes[8];
    int num_kernels;
    int resblock_kernel_sizes[8];
    int num_dilations;
    int resblock_dilations[8][8];
} b200tts_hifigan_config;

typedef struct b200tts_hifigan b200tts_hifigan;
int b200tts_hifigan_create(const b200tts_hifigan_config* cfg, const float* const* weights, int num_weights,
                           b200tts_hifigan** out);
void b200tts_hifigan_destroy(b200tts_hifigan* h);
size_t b200tts_hifigan_workspace_bytes(const b200tts_hifigan* h, int B, int T);
int b200tts_hifigan_out_len(const b200tts_hifigan* h, int T);
/* x [B,Cin,T]; g [B,cond,1] or NULL; wav [B,Cout,out_len(T)] */
int b200tts_hifigan_forward(const b200tts_hifigan* h, const float* x, const float* g, int B, int T, float* wav,
                            void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTS_B200_H */
