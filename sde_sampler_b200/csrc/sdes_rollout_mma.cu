// sdes_rollout_mma.cu — persistent rollout kernel, control MLP on tcgen05 tensor cores.
// (placeholder until the tcgen05 engine lands: reports "unsupported" so callers must ask
// for SDES_F_MLP_SIMT explicitly; there is no silent fallback.)
#include "sdes_common.cuh"

namespace sdes {
bool mma_supported(const KParams&) { return false; }
int64_t mma_weight_image_floats(const SdesRolloutDesc&, int) { return 0; }
cudaError_t launch_rollout_mma(const KParams&, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace sdes
