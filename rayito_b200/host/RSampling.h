// Drop-in header name of the Rayito API.  Sampling (Rng, CorrelatedMultiJitterSampler,
// MIS heuristics, sample warps) runs on the GPU (rayito_b200/csrc/rt_sampling.cuh);
// the host API exposes no samplers.
#ifndef RAYITO_B200_COMPAT_RSAMPLING_H
#define RAYITO_B200_COMPAT_RSAMPLING_H
#include "rayito_b200/math.hpp"
#endif
