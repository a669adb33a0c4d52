/* pt_oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE, see pt_oracle.c header). */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pto_scene pto_scene;

enum { PTO_RNG_PHILOX = 0, PTO_RNG_SEQ = 1, PTO_RNG_MOCK = 2 };
enum { PTO_SINCOS_DET = 0, PTO_SINCOS_LIBM = 1 };
enum { PTO_ACCUM_FORWARD = 0, PTO_ACCUM_RECURSIVE = 1 };

typedef struct {
    int32_t rng_mode;    /* PTO_RNG_*    */
    int32_t sincos_mode; /* PTO_SINCOS_* */
    int32_t accum_mode;  /* PTO_ACCUM_*  */
    uint64_t seed;
    int32_t threads;     /* worker threads (>=1) */
    int32_t shuffle;     /* shuffle the pixel list like mod.rs:1021-1022 (scheduling only) */
} pto_render_cfg;

pto_scene *pto_scene_load(const char *json_path, const char *base_dir, char *err, int errlen);
/* fan_polygons != 0: opt-in fan triangulation of n-gon faces (NOT reference behaviour; load_off.rs:73-76 rejects them) */
pto_scene *pto_scene_load_ex(const char *json_path, const char *base_dir, int fan_polygons, char *err, int errlen);
void pto_scene_free(pto_scene *sc);
int pto_scene_counts(const pto_scene *sc, int *nobjs, int *nspheres, int *nmeshes, int *ntris);
const char *pto_scene_id(const pto_scene *sc);
int pto_scene_mesh_bounds(const pto_scene *sc, int obj, float *pos3, float *radius);

/* sum framebuffer (W*H*3 fp32, reference index order) += radiance of samples [spp_begin, spp_begin+spp_count) */
int pto_render(const pto_scene *sc, int W, int H, uint64_t spp_begin, uint64_t spp_count, const pto_render_cfg *cfg,
               float *sum_rgb, uint64_t *stats4 /* segments, sphere tests, gate tests, triangle tests */);
int pto_render_region(const pto_scene *sc, int W, int H, uint32_t pixel_begin, uint32_t pixel_count,
                      uint64_t spp_begin, uint64_t spp_count, const pto_render_cfg *cfg, float *sum_rgb,
                      uint64_t *stats4);
void pto_mock_reset(void);
void pto_resolve(const float *sum_rgb, size_t n_floats, uint64_t spp, float *mean_rgb);
uint32_t pto_to_int_with_gamma_correction(float x);
int pto_write_ppm(const char *path, const float *mean_rgb, int W, int H, uint64_t spp, const char *scene_id,
                  uint64_t seconds);

int pto_intersect(const pto_scene *sc, const float *rays6, int n, int32_t *obj, int32_t *tri, float *t, float *point3,
                  float *normal3);
int pto_primary_rays(const pto_scene *sc, int W, int H, float *rays6);
int pto_primary_hits(const pto_scene *sc, int W, int H, int32_t *obj, int32_t *tri, float *t);
void pto_camera_frame(const pto_scene *sc, float *out12 /* lens_center, su, sv, sensor_origin */);
int pto_radiance_mean(const pto_scene *sc, const float *ray6, uint64_t n, const pto_render_cfg *cfg, float *mean3);

void pto_philox4x32_10(const uint32_t *ctr4, const uint32_t *key2, uint32_t *out4);
void pto_sincos_det(float x, float *s, float *c);
float pto_vec_dot(const float *a, const float *b);
float pto_vec_length(const float *a);
void pto_vec_cross(const float *a, const float *b, float *o);
void pto_vec_normalize(const float *a, float *o);
void pto_vec_divs(const float *a, float s, float *o);

#ifdef __cplusplus
}
#endif
#endif
