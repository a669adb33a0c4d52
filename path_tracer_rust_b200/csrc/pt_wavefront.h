// pt_wavefront.h -- host interface of the wavefront integrator (internal)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>

#include "pt_launch.h"

namespace ptb {

constexpr int WF_MAX_BOUNCES = 12;  // = MAX_DEPTH (mod.rs:661): a branch makes at most 12 radiance() calls

struct WfQueue {   // SoA of ray segments: 4 x float4 per entry
    float4 *o = nullptr;  // origin | path id (sample offset * npix + pixel)
    float4 *d = nullptr;  // direction | depth + (branch code << 8)
    float4 *T = nullptr;  // throughput
    float4 *L = nullptr;  // emission sum of the branch so far
    // closest hit so far: initialised with the shared-memory ("loose") part of the scene by the kernel that creates the
    // segment (all lanes converged there), completed by k_wf_trace with the BVH part, consumed by k_wf_shade
    float *hit_t = nullptr;
    int *hit_ref = nullptr;
    unsigned *hit_prio = nullptr;
};

struct WfWorkspace {
    WfQueue q[2];
    float4 *slots = nullptr;  // [4 branches][n_paths] finished branch sums
    int *counters = nullptr;  // per bounce: queue length, fetch cursor
    int2 *overflow = nullptr; // cooperative trace kernel: stack entries beyond the shared-memory part, per sub-warp
    size_t cap_overflow = 0;
    size_t cap_paths = 0;
    // ray reordering between bounces (wf_sort): keys / queue indices, double-buffered for the radix sort, and its scratch
    unsigned *sort_keys[2] = {nullptr, nullptr};
    int *sort_idx[2] = {nullptr, nullptr};
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0, cap_sort = 0;
};

void wf_release(WfWorkspace &w);
cudaError_t wavefront_render(const DScene &sc, const RenderArgs &a, WfWorkspace &w, int sm_count, size_t target_paths, int refill,
                             int descend_min, int coop, int sort_mode, cudaStream_t st,
                             unsigned *launches);

}  // namespace ptb
