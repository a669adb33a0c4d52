// pt_bvh.cuh -- BVH traversal for the big-mesh part of the scene (device side).
#pragma once
#include "pt_device.cuh"

namespace ptb {

// placeholder until the LBVH lands: scenes are fully "loose" (brute force from shared memory)
__device__ __forceinline__ void bvh_closest_hit(const DScene &sc, V3 o, V3 d, Hit &best) {
    (void)sc; (void)o; (void)d; (void)best;
}

}  // namespace ptb
