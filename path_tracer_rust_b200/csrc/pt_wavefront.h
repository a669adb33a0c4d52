// pt_wavefront.h -- host interface of the wavefront integrator (internal)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>

#include "pt_launch.h"

namespace ptb {

constexpr int WF_MAX_BOUNCES = 12;  // = MAX_DEPTH (mod.rs:661): a branch makes at most 12 radiance() calls
constexpr int WF_CTR_STRIDE = 4;    // ints per bounce in the counter block: front count, back count (one 64-bit word), fetch cursor, -

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
    // Filled from both ends: entries 0, 1, ... are the segments whose [0, hit_t] touches the BVH's root box (the only ones
    // k_wf_trace looks at), entries cap-1, cap-2, ... the others.
    int cap = 0;          // entries allocated
};

struct WfWorkspace {
    WfQueue q[2];
    float4 *slots = nullptr;   // [4 branches][n_paths] finished branch sums
    int *branch_mask = nullptr; // [n_paths] which of the branches 1..3 exist (bit c): set by the split that creates them
    int *counters = nullptr;   // WF_CTR_STRIDE ints per bounce
    size_t cap_paths = 0;
};

struct WfOptions {
    size_t target_paths = 1u << 25;  // paths in flight per batch (measured on B200: 2^23 -> 2^25 = +15 % synthetic, +8 % mesh.json; 2^26 +3 % more)
    int refill = 8;                  // idle lanes of a warp that trigger a refill from the queue
    int top8_nodes = 0;              // nodes of the eight-wide BVH (breadth first) staged in shared memory, <= BVH8_TOP_MAX.  Measured on
                                     // B200: 0 / 73 / 256 / 512 nodes = 236.9 / 235.5 / 235.6 / 237.3 Mpaths/s -- the 96-byte nodes of the top
                                     // levels stay in the L1 by themselves (the four-wide kernel gains 11 % from its copy), so: off
    int refill_wide = 4, trace_threads_wide = 1024;  // the eight-wide kernel's refill threshold and CTA size (measured: 8 / 512 -> 246.5,
                                     // 4 / 512 -> 251.9, 8 / 1024 -> 249.6, 4 / 1024 -> 255.6 Mpaths/s on the synthetic scene)
    int descend_min_wide = 24;       // the same for the eight-wide kernel (12 / 16 / 24: 228 / 234 / 237 Mpaths/s on the synthetic scene)
    int descend_min = 16;            // lanes that must still be descending for the node loop to go on
    int trace_threads = 512;         // CTA size of the trace kernel (256, 512 or 1024): copies of the BVH's top levels per SM
};

void wf_release(WfWorkspace &w);
cudaError_t wavefront_render(const DScene &sc, const RenderArgs &a, WfWorkspace &w, int sm_count, const WfOptions &opt, cudaStream_t st,
                             unsigned *launches);

}  // namespace ptb
