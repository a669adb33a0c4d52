//! src/render/ffi.rs -- `extern "C"` binding of libptb.so (include/ptb.h) for the reference crate.
//! UNTESTED IN THIS REPO: the build image has no cargo/rustc.  It is delivered as source for the maintainer;
//! the same entry points are exercised through ctypes by tests/ and through C++ by csrc/render_cli.cpp.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct ptb_ctx { _private: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_triangle { pub a: [f32; 3], pub b: [f32; 3], pub c: [f32; 3] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_object {
    pub kind: i32,            // 0 sphere, 1 mesh
    pub reflect_type: i32,    // 0 Diffuse, 1 Specular, 2 Refract
    pub position: [f32; 3],
    pub color: [f32; 3],
    pub emission: [f32; 3],
    pub radius: f32,
    pub bs_position: [f32; 3],
    pub bs_radius: f32,
    pub tri_begin: u64,
    pub tri_count: u64,
}

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct ptb_camera { pub position: [f32; 3], pub direction: [f32; 3], pub focal_length: f32, pub sensor_width: f32, pub aspect_ratio: f32 }

#[repr(C)]
pub struct ptb_scene_desc {
    pub objects: *const ptb_object, pub n_objects: u64,
    pub triangles: *const ptb_triangle, pub n_triangles: u64,
    pub camera: ptb_camera,
}

pub const PTB_OK: c_int = 0;
pub const PTB_CANCELLED: c_int = 1;
pub const PTB_OUT_MEAN: c_int = 0;
pub const PTB_OUT_SUM: c_int = 1;

pub type ptb_preview_fn = Option<extern "C" fn(user: *mut c_void, mean_rgb: *const f32, width: c_int, height: c_int,
                                               spp_done: u64, spp_total: u64)>;

#[link(name = "ptb")]
extern "C" {
    pub fn ptb_device_count() -> c_int;
    pub fn ptb_create(device_id: c_int, out: *mut *mut ptb_ctx) -> c_int;
    /// one context for several GPUs of the box (include/ptb.h: ptb_create_multi)
    pub fn ptb_create_multi(device_ids: *const c_int, n_devices: c_int, out: *mut *mut ptb_ctx) -> c_int;
    pub fn ptb_destroy(ctx: *mut ptb_ctx);
    pub fn ptb_last_error(ctx: *const ptb_ctx) -> *const c_char;
    pub fn ptb_upload_scene(ctx: *mut ptb_ctx, desc: *const ptb_scene_desc) -> c_int;
    pub fn ptb_render(ctx: *mut ptb_ctx, width: c_int, height: c_int, spp_begin: u64, spp_count: u64, seed: u64, out_kind: c_int,
                      out_rgb: *mut f32, cancel: *const i32, samples_done: *mut u64) -> c_int;
    /// ptb_render + RenderUpdate.image previews through `on_preview` (include/ptb.h: ptb_render_progressive)
    pub fn ptb_render_progressive(ctx: *mut ptb_ctx, width: c_int, height: c_int, spp_begin: u64, spp_count: u64, seed: u64,
                                  out_kind: c_int, out_rgb: *mut f32, cancel: *const i32, samples_done: *mut u64,
                                  preview_interval_ms: f64, on_preview: ptb_preview_fn, user: *mut c_void) -> c_int;
    pub fn ptb_render_device(ctx: *mut ptb_ctx, width: c_int, height: c_int, spp_begin: u64, spp_count: u64, seed: u64,
                             d_sum_rgb: *mut f32, cuda_stream: *mut c_void, cancel: *const i32, samples_done: *mut u64) -> c_int;
    pub fn ptb_primary_hits(ctx: *mut ptb_ctx, width: c_int, height: c_int, obj: *mut i32, tri: *mut i32, t: *mut f32) -> c_int;
    pub fn ptb_intersect(ctx: *mut ptb_ctx, rays6: *const f32, n: u64, obj: *mut i32, tri: *mut i32, t: *mut f32,
                         point3: *mut f32, normal3: *mut f32) -> c_int;
}

/// SceneData (mod.rs:121-125) -> the flat arrays ptb_upload_scene takes.  Lives next to the private Mesh/Triangle types
/// (`mod ffi;` inside src/render/mod.rs), so it can read `mesh.triangles` and `mesh.bounding_sphere`.
pub fn flatten(scene: &super::SceneData) -> (Vec<ptb_object>, Vec<ptb_triangle>, ptb_camera) {
    use super::{ReflectType, SceneObject};
    let mut objs = Vec::with_capacity(scene.objects.len());
    let mut tris: Vec<ptb_triangle> = Vec::new();
    for o in &scene.objects {
        let mut p = ptb_object::default();
        p.position = o.position.to_array();
        p.color = o.material.color.to_array();
        p.emission = o.material.emmission.to_array();
        p.reflect_type = match o.material.reflect_type { ReflectType::Diffuse => 0, ReflectType::Specular => 1, ReflectType::Refract => 2 };
        match &o.type_ {
            SceneObject::Sphere { radius } => { p.kind = 0; p.radius = *radius; }
            SceneObject::Mesh { mesh, file: _ } => {
                p.kind = 1;
                p.bs_position = mesh.bounding_sphere.position.to_array();
                p.bs_radius = mesh.bounding_sphere.radius;
                p.tri_begin = tris.len() as u64;
                p.tri_count = mesh.triangles.len() as u64;
                tris.extend(mesh.triangles.iter().map(|t| ptb_triangle { a: t.a.to_array(), b: t.b.to_array(), c: t.c.to_array() }));
            }
        }
        objs.push(p);
    }
    let c = &scene.camera;
    let cam = ptb_camera { position: c.position.to_array(), direction: c.direction().to_array(), focal_length: c.focal_length,
                           sensor_width: c.sensor_width, aspect_ratio: c.aspect_ratio };
    (objs, tris, cam)
}
