// extern "C" surface of libtts_b200.so -- see include/tts_b200.h for the contract.
#include <new>

#include "engines.cuh"

using namespace b200tts;

struct b200tts_hifigan { Hifigan impl; };

extern "C" {

const char* b200tts_last_error(void) { return last_error(); }
unsigned long long b200tts_launch_count(void) { return g_launch_count; }
int b200tts_version(void) { return 100; }

size_t b200tts_mas_workspace_bytes(int B, int Tx, int Ty) { return mas_workspace_bytes(B, Tx, Ty); }

int b200tts_mas(const float* value, const float* mask, const int32_t* t_x, const int32_t* t_y, int B, int Tx,
                int Ty, void* path, int path_is_f32, void* workspace, size_t workspace_bytes, void* stream) {
    return mas_forward(value, mask, t_x, t_y, B, Tx, Ty, path, path_is_f32, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

int b200tts_hifigan_create(const b200tts_hifigan_config* cfg, const float* const* weights, int num_weights,
                           b200tts_hifigan** out) {
    if (!cfg || !weights || !out) { set_error("hifigan_create: null argument"); return 1; }
    *out = nullptr;
    b200tts_hifigan* h = new (std::nothrow) b200tts_hifigan();
    if (!h) { set_error("hifigan_create: out of host memory"); return 1; }
    int rc = h->impl.init(*cfg, weights, num_weights);
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
}
void b200tts_hifigan_destroy(b200tts_hifigan* h) { delete h; }
size_t b200tts_hifigan_workspace_bytes(const b200tts_hifigan* h, int B, int T) {
    return h ? h->impl.workspace_bytes(B, T) : 0;
}
int b200tts_hifigan_out_len(const b200tts_hifigan* h, int T) { return h ? h->impl.out_len(T) : 0; }
int b200tts_hifigan_forward(const b200tts_hifigan* h, const float* x, const float* g, int B, int T, float* wav,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) { set_error("hifigan_forward: null handle"); return 1; }
    return h->impl.forward(x, g, B, T, wav, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
