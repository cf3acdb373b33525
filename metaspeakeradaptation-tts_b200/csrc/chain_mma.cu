// Grouped persistent recurrences with tensor-core gate products -- see kernels.h.
#include "rec_common.cuh"
#include "kernels.h"

namespace msa {

bool chain_mma_supported(const msa_config& cfg, int G, int B, int T, int L, int sm_count, size_t smem_limit) {
    (void)cfg; (void)G; (void)B; (void)T; (void)L; (void)sm_count; (void)smem_limit;
    return false;
}
int launch_lstm_rec_fwd_mma(const LstmRecParams&, int, size_t, cudaStream_t) { set_error("chain_mma: not built"); return MSA_E_UNSUPPORTED; }
int launch_lstm_rec_bwd_mma(const LstmRecBwdParams&, int, size_t, cudaStream_t) { set_error("chain_mma: not built"); return MSA_E_UNSUPPORTED; }
int launch_attn_chain_fwd_mma(const AttnChainParams&, int, size_t, cudaStream_t) { set_error("chain_mma: not built"); return MSA_E_UNSUPPORTED; }
int launch_attn_chain_bwd_mma(const AttnChainBwdParams&, int, size_t, cudaStream_t) { set_error("chain_mma: not built"); return MSA_E_UNSUPPORTED; }

}  // namespace msa
