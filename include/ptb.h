/*
 * ptb.h -- C ABI of the B200 path-tracing backend (libptb.so).
 *
 * Drop-in boundary for the reference's rayon render path.  The reference has no FFI today; the
 * narrowest seam is `pub fn render(RenderConfig, &mut Sink<RenderUpdate>, Arc<AtomicBool>) -> RenderDone`
 * (src/render/mod.rs:928-934, sole caller src/main.rs:364).  A Rust maintainer keeps that
 * signature and replaces the body of the pixel loop (mod.rs:998-1024) by the calls declared here;
 * INTEGRATION.md shows the `extern "C"` block and build.rs step.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain C types only; return 0 = PTB_OK, <0 = error (message via ptb_last_error),
 * PTB_CANCELLED (1) = stopped early by the cancel flag.  Nothing throws or aborts across the
 * boundary.  The caller owns every host buffer it passes; the library owns all device memory
 * behind the opaque ptb_ctx.  A ctx is used from one host thread at a time.  ptb_create gives one CUDA device;
 * ptb_create_multi gives a context that drives several devices of one box from this process (samples per pixel split
 * across them, one internal host thread per device, framebuffers summed over NVLink peer memory inside ptb_render).
 * One process per GPU with the caller's own collective works too (ptb_render_device + ptb_peer_reduce_resolve / NCCL).
 * There is NO CPU fallback: every compute entry point fails with PTB_ERR_CUDA without a device.
 */
#ifndef PTB_H
#define PTB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_OK 0
#define PTB_CANCELLED 1
#define PTB_ERR_ARG (-1)
#define PTB_ERR_CUDA (-2)
#define PTB_ERR_IO (-3)
#define PTB_ERR_PARSE (-4)
#define PTB_ERR_STATE (-5)
#define PTB_ERR_LIMIT (-6)

#define PTB_ABI_VERSION 2

typedef struct ptb_ctx ptb_ctx;     /* device context: streams, scene buffers, BVH, framebuffer */
typedef struct ptb_scene ptb_scene; /* host-side scene loaded from scenes/<id>.json (+ OFF meshes) */

/* ---- data model: mirrors SceneData / SceneObjectData / Material / CameraData ------------------- */

enum { PTB_OBJ_SPHERE = 0, PTB_OBJ_MESH = 1 };                         /* SceneObject,  mod.rs:327-335 */
enum { PTB_REFL_DIFFUSE = 0, PTB_REFL_SPECULAR = 1, PTB_REFL_REFRACT = 2 }; /* ReflectType, mod.rs:72-76 */

typedef struct ptb_triangle { /* Triangle{a,b,c}, mod.rs:539-543 (mesh-local, before the +position translate) */
    float a[3], b[3], c[3];
} ptb_triangle;

typedef struct ptb_object { /* SceneObjectData{type_, position, material}, mod.rs:254-258 + Material mod.rs:79-83 */
    int32_t kind;           /* PTB_OBJ_* */
    int32_t reflect_type;   /* PTB_REFL_* */
    float position[3];
    float color[3];
    float emission[3];      /* `emmission` (sic) */
    float radius;           /* Sphere{radius} */
    float bs_position[3];   /* Mesh.bounding_sphere.position (mesh-local; mod.rs:446) */
    float bs_radius;        /* Mesh.bounding_sphere.radius */
    uint64_t tri_begin;     /* Mesh.triangles = triangles[tri_begin .. tri_begin+tri_count) */
    uint64_t tri_count;
} ptb_object;

typedef struct ptb_camera { /* CameraData, mod.rs:163-176; `direction` is used as stored (not renormalised) */
    float position[3];
    float direction[3];
    float focal_length, sensor_width, aspect_ratio;
} ptb_camera;

typedef struct ptb_scene_desc { /* SceneData{objects, camera}, mod.rs:121-125 */
    const ptb_object *objects;
    uint64_t n_objects;
    const ptb_triangle *triangles;
    uint64_t n_triangles;
    ptb_camera camera;
} ptb_scene_desc;

typedef struct ptb_stats {
    uint64_t segments;        /* closest-hit queries (intersect_scene calls, mod.rs:663) of the last render */
    uint64_t samples;         /* pixel-samples of the last render */
    double render_ms;         /* device time of the render kernels of the last render (CUDA events) */
    double upload_ms;         /* last ptb_upload_scene: H2D + flatten */
    double bvh_build_ms;      /* last ptb_upload_scene: device BVH build */
    uint32_t kernel_launches; /* kernels launched by the last render / query call */
    uint32_t n_loose_objects, n_loose_triangles;   /* brute-force (shared-memory) part of the scene */
    uint32_t n_bvh_triangles, n_bvh_spheres, n_bvh_nodes;
    uint64_t bvh_nodes_visited;  /* last render (wavefront integrator): inner BVH nodes fetched */
    uint64_t bvh_prims_tested;   /* last render (wavefront integrator): leaf primitives tested */
} ptb_stats;

/* ---- host-side scene I/O: SceneDescriptor::load + to_data (mod.rs:92-110, 304-318), load_off.rs:8-85 ---- */
int ptb_scene_load_json(const char *json_path, const char *base_dir, ptb_scene **out, char *err, size_t errlen);
/* Opt-in extension, NOT reference behaviour: PTB_LOAD_TRIANGULATE_POLYGONS fan-triangulates OFF faces with more than three
 * vertices (meshes/hdodec.off has pentagons; load_off.rs:73-76 rejects it and so does ptb_scene_load_json). */
#define PTB_LOAD_TRIANGULATE_POLYGONS 1u
int ptb_scene_load_json_ex(const char *json_path, const char *base_dir, uint32_t flags, ptb_scene **out, char *err, size_t errlen);
const ptb_scene_desc *ptb_scene_get_desc(const ptb_scene *scene);
/* SceneData::to_descriptor + SceneDescriptor::save (mod.rs:112-150): writes serde_json::to_string_pretty's exact layout
 * (MeshFile objects keep their path + scale, inline meshes their serialised bounding sphere and box). */
int ptb_scene_save_json(const ptb_scene *scene, const char *json_path);
/* the GUI edits the camera and saves the scene (viewport_tab.rs); same here */
int ptb_scene_set_camera(ptb_scene *scene, const ptb_camera *camera);
const char *ptb_scene_id(const ptb_scene *scene);
void ptb_scene_free(ptb_scene *scene);

/* ---- context ---------------------------------------------------------------------------------- */
int ptb_abi_version(void);
int ptb_device_count(void);
int ptb_create(int device_id, ptb_ctx **out);
/* Multi-GPU inside the library (the seam is still render(), mod.rs:928-934 / main.rs:364 / cmd_render.rs:17-44: one call, all GPUs).
 * The context drives device_ids[0..n_devices): ptb_upload_scene replicates the scene (and builds the BVH) on every device,
 * ptb_set_option applies to all, ptb_render / ptb_render_progressive render global sample indices
 * [spp_begin + g*spp_count/n, spp_begin + (g+1)*spp_count/n) on device g and sum the framebuffers in device order with the
 * fused peer-memory reduce+resolve kernel, ptb_get_stats reports the whole job.  The parity hooks (ptb_primary_hits, ptb_intersect)
 * run on device_ids[0].  n_devices = 1 is the same as ptb_create.  Like every entry point that takes a context, the calls leave
 * the calling thread's current CUDA device at the context's (first) device. */
int ptb_create_multi(const int *device_ids, int n_devices, ptb_ctx **out);
int ptb_device_ids(const ptb_ctx *ctx, int *ids, int cap); /* returns the number of devices the context drives */
void ptb_destroy(ptb_ctx *ctx);
const char *ptb_last_error(const ptb_ctx *ctx); /* ctx may be NULL: last error of the calling thread */

/* Replaces the `&config.scene.objects` borrow of render_pixel (mod.rs:802): flattens to SoA (pre-translated
 * triangles A'=a+pos, E1, E2 with the reference's own roundings), uploads, builds the device BVH. */
int ptb_upload_scene(ptb_ctx *ctx, const ptb_scene_desc *desc);
int ptb_get_stats(const ptb_ctx *ctx, ptb_stats *out);
/* Diagnostic, host only (no GPU, no context): the shared-memory scene stream (float4 records, layout in DESIGN.md section 3) and the
 * per-triangle shading records exactly as ptb_upload_scene builds them when every object stays in the lock-step list.  Call with
 * NULL buffers to get the sizes.  Used by the CPU tests to check the flattening against the oracle without a GPU. */
int ptb_flatten_loose(const ptb_scene_desc *desc, double quad_min_ratio, float *stream, uint64_t stream_cap_floats,
                      uint64_t *stream_floats, float *tris, uint64_t tris_cap_floats, uint64_t *tris_floats);
/* Tuning knobs, effective from the next ptb_upload_scene (results never depend on them, only speed):
 *   "bvh_min_tris"    meshes with at least this many triangles are traversed through the BVH (default 24; a huge value = brute force)
 *   "bvh_min_spheres" scenes with at least this many spheres put them in the BVH (default 48)
 *   "integrator"      0 = auto (wavefront when the scene has a BVH or the frame is small, else megakernel), 1 = megakernel, 2 = wavefront
 *   "wavefront_paths" paths in flight per wavefront batch (default 2^25; the workspace is 0.7 KB per path)
 *   "bvh_wide"        compressed eight-wide BVH for the wavefront trace kernel: 1 = always, 0 = never, -1 = for large sets (default)
 *   "bvh_sah_max_prims" sets up to this size get a binned-SAH topology built on the host instead of the device LBVH (default 16384)
 *   "bvh_leaf_max" (1..8, default 2), "bvh_top_levels" (0..5: four-wide levels the trace kernel keeps in shared memory),
 *   "wf_refill", "wf_descend_min", "wf_trace_threads" (256 / 512 / 1024), "wf_top8_nodes" (eight-wide nodes staged in shared memory,
 *   default 0), "bvh_wide_sah" (eight-wide collapse by the surface-area cost recurrence, default on; 0 = greedy): traversal tuning,
 *   see DESIGN.md
 *   "quad_min_ratio"  a two-triangle mesh whose bounding-sphere radius is at least this fraction of the scene diagonal is tested
 *                     without the per-mesh warp vote (default 0.125; 0 = every two-triangle mesh, a huge value = none)
 *   "regen_batch"     lanes that must be waiting for a camera ray before the ray-generation code runs (default 24) */
int ptb_set_option(ptb_ctx *ctx, const char *key, double value);
/* Device self-test: the kernels' own correctly-rounded reciprocal (MUFU.RCP + 2 FFMA, no range check) is compared with
 * __frcp_rn over every float whose exponent field is in [1, 252]; *mismatches must come back 0. */
int ptb_selftest(ptb_ctx *ctx, uint64_t *mismatches);

/* ---- the hot path: replaces the rayon loop + render_pixel + radiance (mod.rs:1001-1024, 794-857, 662-792) --
 * Renders global sample indices [spp_begin, spp_begin+spp_count) of every pixel.  `out_rgb` is W*H*3 fp32 in the
 * reference's buffer order (index i <-> x = i % W, y = H-1 - i / W, mod.rs:805-806):
 *   PTB_OUT_MEAN : sum / spp_count clamped to [0,1]  == Image.pixels (mod.rs:849-856)
 *   PTB_OUT_SUM  : raw fp32 radiance sum (for spp-sharded multi-GPU / progressive use; add, then resolve)
 * `cancel` (may be NULL) is polled between launches like stop_render (mod.rs:1003); `samples_done` (may be NULL)
 * receives pixel-samples finished so far (processed_pixel_count analogue, mod.rs:850).  When either is given the work per
 * launch is capped (about 0.1-0.2 s), so a cancel takes effect and progress moves at that granularity. */
enum { PTB_OUT_MEAN = 0, PTB_OUT_SUM = 1 };
int ptb_render(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed,
               int out_kind, float *out_rgb, const volatile int32_t *cancel, volatile uint64_t *samples_done);

/* RenderUpdate{progress, image} (mod.rs:882-886, sent about every 500 ms by mod.rs:965-982 and drawn by views/render_tab.rs:278-297):
 * ptb_render_progressive is ptb_render that also hands the caller the image so far.  Between launches, once `preview_interval_ms`
 * have passed since the last preview, the partial sum is resolved (mean of the samples finished so far, clamped) and `on_preview`
 * is called on the rendering thread with a library-owned host buffer that is valid during the call only; spp_done / spp_total is
 * the progress.  The final image arrives through out_rgb exactly as with ptb_render (bit-identical to an unpreviewed render).
 * On a multi-GPU context the previews show device 0's share of the samples. */
typedef void (*ptb_preview_fn)(void *user, const float *mean_rgb, int width, int height, uint64_t spp_done, uint64_t spp_total);
int ptb_render_progressive(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed,
                           int out_kind, float *out_rgb, const volatile int32_t *cancel, volatile uint64_t *samples_done,
                           double preview_interval_ms, ptb_preview_fn on_preview, void *user);

/* Same, accumulating into a caller-provided DEVICE sum framebuffer (W*H*3 fp32, += in sample order) on
 * `cuda_stream` (a cudaStream_t; NULL = the default stream).  Asynchronous unless cancel/samples_done is given.  The kernels use
 * per-context state (queues, counters, BVH, scene buffers); the library orders the next render / upload on this context after this
 * one with an event, whatever stream it runs on.  The CALLER's buffer (d_sum_rgb) is the caller's to synchronise. */
int ptb_render_device(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed,
                      float *d_sum_rgb, void *cuda_stream, const volatile int32_t *cancel,
                      volatile uint64_t *samples_done);
/* radiance / spp then clamp (mod.rs:849-856) on device buffers; d_mean may alias d_sum */
int ptb_resolve_device(ptb_ctx *ctx, const float *d_sum_rgb, uint64_t n_floats, uint64_t spp_total, float *d_mean_rgb,
                       void *cuda_stream);

/* ---- multi-GPU reduce without a library collective on the data path --------------------------------------------
 * One process per GPU.  Every rank renders its share of the samples into a sum framebuffer obtained from ptb_device_alloc,
 * exports it (CUDA IPC), opens the other ranks' buffers, and after a barrier reduces + resolves ITS slice of the image straight
 * out of peer memory over NVLink into rank 0's output buffer: one kernel does the sum over ranks (fixed rank order, so the image
 * is deterministic), the division by spp, the clamp (mod.rs:849-856) and the P2P store.  See path_tracer_rust_b200/distributed.py. */
int ptb_device_alloc(ptb_ctx *ctx, uint64_t n_bytes, void **d_ptr);
int ptb_device_free(ptb_ctx *ctx, void *d_ptr);
int ptb_device_memset(ptb_ctx *ctx, void *d_ptr, int value, uint64_t n_bytes, void *cuda_stream);
int ptb_device_to_host(ptb_ctx *ctx, void *host_dst, const void *d_src, uint64_t n_bytes, void *cuda_stream); /* blocking */
int ptb_device_sync(ptb_ctx *ctx, void *cuda_stream);
int ptb_ipc_export(ptb_ctx *ctx, const void *d_ptr, unsigned char handle64[64]);
int ptb_ipc_open(ptb_ctx *ctx, const unsigned char handle64[64], void **d_ptr);
int ptb_ipc_close(ptb_ctx *ctx, void *d_ptr);
/* d_dst[i] = clamp((sum_g d_peer_sums[g][i]) / spp_total, 0, 1) for i in [first_float, first_float + n_floats) */
int ptb_peer_reduce_resolve(ptb_ctx *ctx, const float *const *d_peer_sums, int n_peers, uint64_t first_float, uint64_t n_floats,
                            uint64_t spp_total, float *d_dst, void *cuda_stream);

/* ---- parity hooks ----------------------------------------------------------------------------- */
/* intersect_scene (mod.rs:631-659) for the deterministic centre ray of every pixel (xsub=ysub=xfilter=yfilter=0
 * in mod.rs:833-838).  obj = object index or -1, tri = triangle index inside the mesh or -1, t = distance. */
int ptb_primary_hits(ptb_ctx *ctx, int width, int height, int32_t *obj, int32_t *tri, float *t);
/* intersect_scene for n arbitrary rays (origin xyz, direction xyz).  point/normal may be NULL. */
int ptb_intersect(ptb_ctx *ctx, const float *rays6, uint64_t n, int32_t *obj, int32_t *tri, float *t, float *point3,
                  float *normal3);

/* ---- output resolve: mod.rs:57-63 (gamma), :1042-1076 (P3 PPM, reversed pixel order), :916-926 (hash) ---- */
uint32_t ptb_to_int_with_gamma_correction(float x);
int ptb_write_ppm(const char *path, const float *mean_rgb, int width, int height, uint64_t spp, const char *scene_id,
                  uint64_t seconds);
uint64_t ptb_hash_pixels(const float *rgb, uint64_t n_pixels);

#ifdef __cplusplus
}
#endif
#endif /* PTB_H */
