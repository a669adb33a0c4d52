// pt_bvh_build.h -- device BVH construction interface (internal)
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/ptb.h"
#include "pt_device.cuh"

namespace ptb {


struct BvhDevice {
    float4 *nodes = nullptr;  // 8 x float4 per four-wide node
    float4 *tris = nullptr;   // [2n] (A, E1) pairs, [n] E2, [n] shading records, leaf order (triangles and spheres)
    uint4 *nodes8 = nullptr;  // compressed eight-wide nodes (pt_bvh8.h), 6 x uint4 each
    float4 *tris8 = nullptr;  // primitive records in the wide order, same four sections as `tris`
    unsigned n_nodes8 = 0;
    size_t cap_nodes8 = 0, cap_tris8 = 0;
    float4 *top = nullptr;    // 8 x float4 per node: copy of the top levels for the trace kernel's shared memory (BVH_TOP_MAX nodes)
    int *top_count = nullptr;
    unsigned n_nodes = 0, n_tris = 0, n_spheres = 0;
    int max_depth = 0;
    size_t cap_nodes = 0, cap_tris = 0;
};

// decides which objects are traversed through the BVH (in_bvh[k] = 1) and which stay in the shared-memory list
struct BvhOptions {
    double min_tris = 24;     // meshes with fewer triangles stay in the lock-step shared-memory list
    double min_spheres = 48;  // scenes with fewer spheres keep them in the shared-memory list
    int leaf_max = 2;         // primitives per leaf after collapsing small subtrees (1..8); measured best on B200 for the 1.3 M-triangle scene
    int leaf_max_small = 4;   // the same for sets that take the host SAH path (mesh.json, 810 triangles: 4 -> +2 % over 2)
    int sah_max_prims = 16384; // sets up to this size get a binned-SAH topology from the host, larger ones the device LBVH (Karras)
    int wide = -1;            // compressed eight-wide BVH for the wavefront trace kernel (pt_bvh8.h): 1 = build it, 0 = never, -1 = when the
                              // set is too large for the host SAH build (measured: synthetic 1.3 M triangles +7 %, mesh.json 810 triangles -6 %)
    int wide_sah = 2;         // eight-wide collapse by the surface-area cost recurrence (1) or greedily by largest area (0)
    int top_levels = 5;       // four-wide levels copied for the trace kernel's shared memory (0..5); measured on B200: 0 -> 4 levels +6 %, 5 levels (512-thread CTAs) +11 %
#ifdef PTB_EXPERIMENTS
    double pad_scale = 1.0;   // scales the conservative box padding; anything below 1 voids the parity guarantee
#else
    static constexpr double pad_scale = 1.0;
#endif
};
void choose_bvh_objects(const ptb_scene_desc &desc, size_t max_smem_bytes, const BvhOptions &opt, std::vector<char> &in_bvh);
// builds the BVH over the chosen objects on the device and fills the bvh_* fields of `ds`
cudaError_t bvh_build(const ptb_scene_desc &desc, const std::vector<char> &in_bvh, const std::vector<uint32_t> &prio_base, const BvhOptions &opt,
                      BvhDevice &out, DScene &ds, cudaStream_t st, double *build_ms, std::string &err);
void bvh_release(BvhDevice &b);

}  // namespace ptb
