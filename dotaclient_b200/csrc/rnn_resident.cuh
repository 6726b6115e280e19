// Weight-resident recurrence kernels (W_hh split between registers and shared memory).
#pragma once
#include "dc_common.cuh"

namespace dc_rnn {

inline bool resident_supported(int cell, int H) { (void)cell; (void)H; return false; }

inline int launch_fwd_resident(int, float *, const float *, const float *, float *, float *, int, int, int, cudaStream_t) {
    return DC_EUNSUPPORTED;
}
inline int launch_bwd_resident(int, float *, const float *, const float *, float *, const float *, const float *,
                               const float *, float *, float *, int, int, int, cudaStream_t) {
    return DC_EUNSUPPORTED;
}

}  // namespace dc_rnn
