// scene_io.hpp -- host scene model above the C ABI (see scene_io.cpp)
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptb.h"

namespace ptb {

struct SceneError : std::runtime_error {
    int code;
    SceneError(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

// SceneData (mod.rs:121-125) in the C ABI's own structs, so it can be handed to ptb_upload_scene directly
// what SceneObjectDescriptorType (mod.rs:298-302) needs to be written back: the MeshFile descriptor or the inline mesh's
// serialised bounding box (mod.rs:441-448)
struct ObjectSource {
    int variant = 0;  // 0 Sphere, 1 MeshFile, 2 Mesh
    std::string path;
    float scale = 1.0f;
    std::vector<ptb_triangle> bounding_box;
};

struct HostScene {
    std::string id;
    std::vector<ptb_object> objects;
    std::vector<ObjectSource> sources;
    std::vector<ptb_triangle> triangles;
    ptb_camera camera{};
    ptb_scene_desc desc{};
    void refresh_desc() {
        desc.objects = objects.data();
        desc.n_objects = objects.size();
        desc.triangles = triangles.data();
        desc.n_triangles = triangles.size();
        desc.camera = camera;
    }
};

HostScene load_scene_json(const std::string &json_path, const std::string &base_dir, unsigned flags = 0);
// SceneData::to_descriptor + SceneDescriptor::save (mod.rs:112-150): serde_json::to_string_pretty layout
std::string scene_to_json(const HostScene &scene);
void save_scene_json(const HostScene &scene, const std::string &json_path);
std::string format_f32(float v);
void load_off(const std::string &path, float scale, std::vector<ptb_triangle> &out, unsigned flags = 0);
void mesh_bounding_sphere(const ptb_triangle *tris, size_t n, float centre[3], float *radius);
uint32_t to_int_with_gamma_correction(float x);
void write_ppm(const std::string &path, const float *mean_rgb, int W, int H, uint64_t spp, const std::string &scene_id,
               uint64_t seconds);
uint64_t hash_pixels(const float *rgb, uint64_t n_pixels);

}  // namespace ptb
