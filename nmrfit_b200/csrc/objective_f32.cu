// Opt-in FP32 objective (placeholder until the FP32 kernel lands).
#include <cuda_runtime.h>
#include "nmrfit_internal.h"

namespace nmrfit {
cudaError_t launch_objective_f32(ObjArgs, const ObjTune&, int, double*, cudaStream_t, cudaEvent_t, cudaEvent_t) {
    return cudaErrorNotSupported;
}
}  // namespace nmrfit
