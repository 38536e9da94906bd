// stft_fast.cu — specialised sm_100a STFT kernels (placeholder until the register/shared-memory FFT engine lands).
#include "stft.h"

namespace omb {
bool stft_fast_supported(const StftConfig&, const DeviceInfo&) { return false; }
int stft_fast_prepare(StftPlan&) { return OMB_OK; }
int launch_stft_fast(const StftPlan&, StftKernelArgs&, cudaStream_t) { return fail(OMB_ERR_UNSUPPORTED, "no specialised kernel"); }
}  // namespace omb
