// Distance epilogues shared by the SIMT and tcgen05 contraction kernels.
#pragma once
#include "common.cuh"

namespace mpreid {

// dot = q.g ; qa/ga = squared norms (euclid), norms (arccos), unused (1-dot)
template <int METRIC>
__device__ __forceinline__ float finish_distance(float dot, float qa, float ga) {
  if (METRIC == MPREID_SQEUCLID) {
    // utils/metrics.py:10-12: (|q|^2 + |g|^2) is rounded to fp32 first, then -2*q.g is accumulated
    return fmaf(-2.0f, dot, qa + ga);
  } else if (METRIC == MPREID_ARCCOS) {
    // utils/metrics.py:16-24: dot * (1 / (|q||g|)), clip to +-(1 - 1e-5), arccos
    float c = dot * (1.0f / (qa * ga));
    c = fminf(fmaxf(c, -0.99999f), 0.99999f);
    return acosf(c);
  } else if (METRIC == MPREID_ONE_MINUS_DOT) {
    return 1.0f - dot;  // processor/processor_uniprompt_stage2.py:466-467
  } else if (METRIC == MPREID_DOT) {
    return dot;         // loss/supcontrast.py:23 (text @ image^T)
  } else {
    // loss/triplet_loss.py:26-30
    return sqrtf(fmaxf(fmaf(-2.0f, dot, qa + ga), 1e-12f));
  }
}

__device__ __forceinline__ float finish_distance_rt(int metric, float dot, float qa, float ga) {
  switch (metric) {
    case MPREID_SQEUCLID: return finish_distance<MPREID_SQEUCLID>(dot, qa, ga);
    case MPREID_ARCCOS: return finish_distance<MPREID_ARCCOS>(dot, qa, ga);
    case MPREID_ONE_MINUS_DOT: return finish_distance<MPREID_ONE_MINUS_DOT>(dot, qa, ga);
    case MPREID_DOT: return finish_distance<MPREID_DOT>(dot, qa, ga);
    default: return finish_distance<MPREID_SQRT_EUCLID>(dot, qa, ga);
  }
}

// float atomic max that is correct for mixed signs (target initialised to -inf)
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
  else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

}  // namespace mpreid
