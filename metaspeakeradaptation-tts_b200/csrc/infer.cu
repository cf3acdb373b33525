// Free-running inference (Tacotron2NV.infer) -- C ABI entry points.
#include "common.cuh"
#include "kernels.h"

extern "C" {
size_t msa_infer_workspace_bytes(const msa_handle* h, int B, int L, int max_steps) {
    (void)h; (void)B; (void)L; (void)max_steps;
    return 0;
}
int msa_infer(msa_handle* h, void* ws, size_t ws_bytes, const float* params, const float* bn_stats, const int64_t* tokens,
              const int64_t* token_lengths, const float* speaker_vecs, const int64_t* speaker_ids, const uint8_t* prenet_masks,
              int B, int L, int max_steps, float* mel_post_out, int32_t* mel_lengths_out, float* align_out, int32_t* n_steps_out,
              void* stream) {
    (void)h; (void)ws; (void)ws_bytes; (void)params; (void)bn_stats; (void)tokens; (void)token_lengths; (void)speaker_vecs;
    (void)speaker_ids; (void)prenet_masks; (void)B; (void)L; (void)max_steps; (void)mel_post_out; (void)mel_lengths_out;
    (void)align_out; (void)n_steps_out; (void)stream;
    msa::set_error("msa_infer: not implemented yet");
    return MSA_E_UNSUPPORTED;
}
}
