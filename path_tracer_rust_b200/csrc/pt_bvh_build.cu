// pt_bvh_build.cu -- device BVH construction (stub: everything stays in the shared-memory list)
#include "pt_bvh_build.h"
#include "pt_launch.h"

namespace ptb {

void choose_bvh_objects(const ptb_scene_desc &desc, size_t, std::vector<char> &in_bvh) {
    in_bvh.assign(desc.n_objects, 0);
}

cudaError_t bvh_build(const ptb_scene_desc &, const std::vector<char> &, const std::vector<uint32_t> &, BvhDevice &out, DScene &ds,
                      cudaStream_t, double *build_ms, std::string &) {
    out.n_nodes = out.n_tris = out.n_spheres = 0;
    ds.bvh_root = BVH_EMPTY;
    if (build_ms) *build_ms = 0.0;
    return cudaSuccess;
}

void bvh_release(BvhDevice &b) {
    if (b.nodes) cudaFree(b.nodes);
    if (b.tris) cudaFree(b.tris);
    if (b.spheres) cudaFree(b.spheres);
    b = BvhDevice{};
}

}  // namespace ptb
