/*
 * pt_oracle.c -- CPU ORACLE for the per-pixel Monte-Carlo radiance loop.
 *
 * >>> TEST INFRASTRUCTURE. NOT PRODUCT CODE. <<<
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / reported CPU baseline.  The product path
 * (path_tracer_rust_b200/csrc, libptb.so) never links, imports or calls anything in oracle/.
 *
 * What it is: a strict-fp32, un-fused (-ffp-contract=off) plain-C restatement of the reference's
 * algorithm, function by function.  The reference (Rust) cannot be compiled in this image (no
 * cargo/rustc), so oracle/_ref does not exist; every function below cites the reference
 * file:line it follows (paths relative to /root/reference).
 *
 * Parity pinning:
 *   PINNED by the reference's own tests (src/render/test.rs), checked in tests/test_oracle_kat.py:
 *     - glam Vec3 op results (test.rs:3-27), gamma->u8 (test.rs:29-35),
 *     - four exact sphere-hit known answers (test.rs:43-144),
 *     - diffuse transport mean (test.rs:146-183; analytic 50/144).
 *   PARITY UNPINNED in the reference (no golden vector exists there): triangle / mesh
 *   intersection, refraction, camera rays, OFF / JSON parsing, RNG output, whole images.  For
 *   those this file *is* the definition, justified line by line against src/render/mod.rs.
 *
 * Third-party arithmetic restated here (crates are not vendored under /root/reference):
 *   glam 0.30.8 scalar Vec3:  dot = (x*x' + y*y') + z*z';  cross = (y*z' - y'*z, z*x' - z'*x,
 *     x*y' - x'*y);  length = sqrt(dot(v,v));  normalize = v * (1.0/length);  Vec3/f32 divides
 *     each component;  rand 0.8.5 Standard f32 = (next_u32() >> 8) * 2^-24.
 *
 * Modes that are NOT in the reference but are needed to compare against a GPU:
 *   - RNG "philox": counter based Philox4x32-10, key = seed, counter = (pixel, sample_lo,
 *     sample_hi, event); event 0 is the camera sample (slots 0,1 = r1,r2); the radiance() call at
 *     depth new_depth (1..12) of path-tree branch `code` has event (code << 4) | new_depth, where
 *     code (0..3) says which of the two deterministic refraction splits (mod.rs:775-786, depth 1
 *     and 2) were left through the transmitted child (bit 0 / bit 1).  Slot 0 = Russian roulette,
 *     1,2 = diffuse r1,r2, 3 = refraction choice.  Events are therefore independent of evaluation
 *     order, which lets a wavefront GPU integrator trace both children of a split concurrently.
 *     The reference draws from an OS-seeded thread-local ChaCha12 (mod.rs:53) and is not
 *     reproducible, so only the distribution can be matched.
 *   - "det" sin/cos: a fixed fp32 polynomial (Cephes sinf/cosf form) evaluated with un-fused ops so
 *     CPU and GPU agree bit for bit; "libm" is the reference's behaviour (f32::sin/cos -> sinf/cosf).
 *   - "forward" accumulation: iterative throughput form of radiance() with a 2-entry stack, the
 *     bit-exact twin of the CUDA integrators.  Every branch of the path tree (code 0..3) sums its own
 *     emission terms; the sample's radiance is ((L0 + L1) + L2) + L3.  "recursive" is the reference's form.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pt_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* glam 0.30.8 Vec3 (scalar) op orders; pinned by test.rs:3-27                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v_add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v_sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v_mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v_scale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 v_divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float v_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 v_cross(v3 a, v3 b) {
    return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float v_length(v3 a) { return sqrtf(v_dot(a, a)); }
static inline v3 v_normalize(v3 a) { return v_scale(a, 1.0f / v_length(a)); }

/* ------------------------------------------------------------------------------------------ */
/* data model: mod.rs:65-83 (Ray, Material), :254 (SceneObjectData), :441 (Mesh), :539 Triangle */
/* ------------------------------------------------------------------------------------------ */
typedef struct { v3 o, d; } ray_t;
typedef struct { v3 a, b, c; } tri_t;
enum { REFL_DIFFUSE = 0, REFL_SPECULAR = 1, REFL_REFRACT = 2 };
enum { OBJ_SPHERE = 0, OBJ_MESH = 1 };

typedef struct {
    int type;
    v3 position;
    v3 color, emission;
    int refl;
    float radius;          /* sphere */
    tri_t *tris;           /* mesh */
    int ntris;
    v3 bs_pos;             /* mesh.bounding_sphere (read from JSON for inline meshes, mod.rs:446) */
    float bs_radius;
} obj_t;

struct pto_scene {
    char id[128];
    obj_t *objs;
    int nobjs;
    /* CameraData mod.rs:163-176 */
    v3 cam_pos, cam_dir;
    float focal_length, sensor_width, aspect_ratio;
};

typedef struct { float t; v3 x, n; int tri; } hit_t;

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11); KAT checked in tests/test_oracle_kat.py                */
/* ------------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void pto_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    philox4x32_10(ctr, key, out);
}
/* rand 0.8.5 Standard<f32>: 24 random mantissa bits, [0,1) */
static inline float u32_to_unit(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

/* ------------------------------------------------------------------------------------------ */
/* deterministic fp32 sin/cos for x in [0, 2*pi] (Cephes sinf/cosf polynomials, un-fused)        */
/* ------------------------------------------------------------------------------------------ */
static void sincos_det(float x, float *s_out, float *c_out) {
    int j = (int)(x * 1.27323954473516f);        /* x * 4/pi, truncated */
    j = (j + 1) & ~1;                            /* even octant index 0..8 */
    float y = (float)j;
    float z = ((x - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float zz = z * z;
    float ps = ((-1.9515295891e-4f * zz + 8.3321608736e-3f) * zz - 1.6666654611e-1f) * zz * z + z;
    float pc = ((2.443315711809948e-5f * zz - 1.388731625493765e-3f) * zz + 4.166664568298827e-2f) * zz * zz
               - 0.5f * zz + 1.0f;
    int q = (j >> 1) & 3;
    float s = (q & 1) ? pc : ps;
    float c = (q & 1) ? ps : pc;
    if (q == 2 || q == 3) s = -s;
    if (q == 1 || q == 2) c = -c;
    *s_out = s; *c_out = c;
}
void pto_sincos_det(float x, float *s, float *c) { sincos_det(x, s, c); }

/* ------------------------------------------------------------------------------------------ */
/* sampler: rand01() of mod.rs:48-55 in three flavours                                           */
/* ------------------------------------------------------------------------------------------ */
static const float MOCK_RANDOMS[9] = { /* mod.rs:33-43, narrowed to f32 as the Rust literals are */
    0.75902418061906407f, 0.023879213030728041f, 0.21016190197770457f,
    0.78814922184253244f, 0.56819568237964491f,  0.7689823904006352f,
    0.16910304067812287f, 0.54519597695203492f,  0.63614169009490062f};
static atomic_size_t g_mock_index; /* mod.rs:45 */

typedef struct {
    const struct pto_scene *sc;
    int rng_mode, sincos_mode;
    uint32_t key[2];
    uint32_t ctr_pixel, ctr_slo, ctr_shi, event;
    float blk[4];
    uint64_t seq_state, seq_inc;
    uint64_t n_segments, n_sphere, n_gate, n_tri;
} sampler_t;

static inline uint32_t pcg32_next(sampler_t *S) {
    uint64_t old = S->seq_state;
    S->seq_state = old * 6364136223846793005ULL + S->seq_inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static void sampler_seed_pixel(sampler_t *S, uint64_t seed, uint32_t pixel) {
    S->key[0] = (uint32_t)seed; S->key[1] = (uint32_t)(seed >> 32);
    S->ctr_pixel = pixel;
    S->seq_state = splitmix64(seed ^ splitmix64(pixel));
    S->seq_inc = (splitmix64(S->seq_state) << 1) | 1u;
    pcg32_next(S);
}
/* start event `ev` of sample (pixel, s): in philox mode this fixes the four addressed draws */
static inline void sampler_event(sampler_t *S, uint32_t ev) {
    S->event = ev;
    if (S->rng_mode == PTO_RNG_PHILOX) {
        uint32_t ctr[4] = {S->ctr_pixel, S->ctr_slo, S->ctr_shi, ev}, out[4];
        philox4x32_10(ctr, S->key, out);
        for (int i = 0; i < 4; ++i) S->blk[i] = u32_to_unit(out[i]);
    }
}
static inline float rand01(sampler_t *S, int slot) {
    switch (S->rng_mode) {
    case PTO_RNG_PHILOX: return S->blk[slot];
    case PTO_RNG_MOCK: { /* mod.rs:49-51 */
        size_t i = atomic_fetch_add_explicit(&g_mock_index, 1, memory_order_relaxed) % 9;
        return MOCK_RANDOMS[i];
    }
    default: return u32_to_unit(pcg32_next(S)); /* sequential thread-local-like stream, mod.rs:53 */
    }
}
static inline void sincos_sel(const sampler_t *S, float x, float *s, float *c) {
    if (S->sincos_mode == PTO_SINCOS_DET) sincos_det(x, s, c);
    else { *s = sinf(x); *c = cosf(x); }  /* f32::sin / f32::cos -> platform libm */
}

/* ------------------------------------------------------------------------------------------ */
/* intersect_sphere  mod.rs:412-438 (pinned by test.rs:43-144)                                   */
/* ------------------------------------------------------------------------------------------ */
static int intersect_sphere(v3 position, float radius, const ray_t *ray, hit_t *h) {
    v3 op = v_sub(position, ray->o);
    const float eps = 1e-4f;
    float b = v_dot(op, ray->d);
    float det = b * b - v_dot(op, op) + radius * radius;
    if (det < 0.0f) return 0;
    det = sqrtf(det);
    float t;
    if (b - det >= eps) t = b - det;
    else if (b + det >= eps) t = b + det;
    else return 0;
    v3 xmin = v_add(ray->o, v_scale(ray->d, t));
    v3 nmin = v_normalize(v_sub(xmin, position));
    h->t = t; h->x = xmin; h->n = nmin; h->tri = -1;
    return 1;
}

/* Triangle::transformed + Triangle::intersect  mod.rs:546-615 (USE_CULLING=false, mod.rs:28) */
static int intersect_triangles(const ray_t *ray, v3 offset, const tri_t *tris, int ntris, hit_t *h,
                               sampler_t *S) {
    int have = 0;
    for (int i = 0; i < ntris; ++i) {
        if (S) S->n_tri++;
        v3 a = v_add(tris[i].a, offset), b = v_add(tris[i].b, offset), c = v_add(tris[i].c, offset);
        v3 va_vb = v_sub(b, a);
        v3 va_vc = v_sub(c, a);
        v3 pvec = v_cross(ray->d, va_vc);
        float determinant = v_dot(va_vb, pvec);
        if (fabsf(determinant) < 1e-4f) continue;
        float inv_determinant = 1.0f / determinant;
        v3 tvec = v_sub(ray->o, a);
        float u = v_dot(tvec, pvec) * inv_determinant;
        if (u < 0.0f || u > 1.0f) continue;
        v3 qvec = v_cross(tvec, va_vb);
        float v = v_dot(ray->d, qvec) * inv_determinant;
        if (v < 0.0f || (u + v) > 1.0f) continue;
        float distance = v_dot(va_vc, qvec) * inv_determinant;
        if (distance <= 0.0f) continue;
        if (!have || distance < h->t) {
            h->t = distance;
            h->x = v_add(ray->o, v_scale(ray->d, distance));
            h->n = v_normalize(v_cross(va_vb, va_vc));
            h->tri = i;
            have = 1;
        }
    }
    return have;
}

/* SceneObjectData::intersect  mod.rs:260-280 */
static int intersect_object(const obj_t *o, const ray_t *ray, hit_t *h, sampler_t *S) {
    if (o->type == OBJ_SPHERE) {
        if (S) S->n_sphere++;
        return intersect_sphere(o->position, o->radius, ray, h);
    }
    hit_t gate;
    if (S) S->n_gate++;
    if (intersect_sphere(v_add(o->bs_pos, o->position), o->bs_radius, ray, &gate))
        return intersect_triangles(ray, o->position, o->tris, o->ntris, h, S);
    return 0;
}

/* intersect_scene  mod.rs:631-659: reverse order, strict < */
static int intersect_scene(const struct pto_scene *sc, const ray_t *ray, hit_t *best, sampler_t *S) {
    int best_id = -1;
    for (int i = sc->nobjs - 1; i >= 0; --i) {
        hit_t h;
        if (!intersect_object(&sc->objs[i], ray, &h, S)) continue;
        if (best_id < 0 || h.t < best->t) { *best = h; best_id = i; }
    }
    return best_id;
}

/* ------------------------------------------------------------------------------------------ */
/* material arms shared by both accumulation forms (mod.rs:687-758)                              */
/* ------------------------------------------------------------------------------------------ */
#define PI_F 3.141592653589793f /* mod.rs:29, narrows to 3.1415927 */

static v3 diffuse_dir(sampler_t *S, v3 nl) { /* mod.rs:691-704 */
    float r1 = 2.0f * PI_F * rand01(S, 1);
    float r2 = rand01(S, 2);
    float r2s = sqrtf(r2);
    v3 w = nl;
    v3 u = v_normalize(v_cross(fabsf(w.x) > 0.1f ? V(0.0f, 1.0f, 0.0f) : V(1.0f, 0.0f, 0.0f), w));
    v3 v = v_cross(w, u);
    float sn, cs;
    sincos_sel(S, r1, &sn, &cs);
    v3 d = v_add(v_add(v_scale(v_scale(u, cs), r2s), v_scale(v_scale(v, sn), r2s)),
                 v_scale(w, sqrtf(1.0f - r2)));
    return v_normalize(d);
}
static v3 reflect_dir(v3 d, v3 n) { /* mod.rs:722-723: d - n*2.0*n.dot(d) */
    return v_sub(d, v_scale(v_scale(n, 2.0f), v_dot(n, d)));
}
/* refraction terms, mod.rs:736-758.  returns 0 on total internal reflection */
typedef struct { v3 tdir; float re, tr, p, rp, tp; } fresnel_t;
static int refract_terms(v3 d, v3 n, v3 nl, fresnel_t *f) {
    int into = v_dot(n, nl) > 0.0f;
    const float nc = 1.0f, nt = 1.5f;
    float nnt = into ? nc / nt : nt / nc;
    float ddn = v_dot(d, nl);
    float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
    if (cos2t < 0.0f) return 0;
    f->tdir = v_normalize(v_sub(v_scale(d, nnt),
                                v_scale(n, (into ? 1.0f : -1.0f) * (ddn * nnt + sqrtf(cos2t)))));
    float a = nt - nc, b = nt + nc;
    float r0 = a * a / (b * b);
    float c = 1.0f - (into ? -ddn : v_dot(f->tdir, n));
    float c2 = c * c;
    f->re = r0 + (1.0f - r0) * (c * (c2 * c2)); /* powi(5): x*x, (x^2)^2, *x */
    f->tr = 1.0f - f->re;
    f->p = 0.25f + 0.5f * f->re;
    f->rp = f->re / f->p;
    f->tp = f->tr / (1.0f - f->p);
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* radiance  mod.rs:661-792  (reference form: recursive)                                         */
/* ------------------------------------------------------------------------------------------ */
#define MAX_DEPTH 12
static v3 radiance_rec(const ray_t *ray, int depth, int code, sampler_t *S) {
    hit_t hit;
    S->n_segments++;
    sampler_event(S, ((uint32_t)code << 4) | (uint32_t)(depth + 1));
    int id = intersect_scene(S->sc, ray, &hit, S);
    if (id < 0) return V(0.0f, 0.0f, 0.0f);
    const obj_t *o = &S->sc->objs[id];
    v3 color = o->color;
    float max_reflection = fmaxf(color.x, fmaxf(color.y, color.z));
    v3 nl = v_dot(hit.n, ray->d) < 0.0f ? hit.n : v_scale(hit.n, -1.0f);
    int new_depth = depth + 1;
    if (new_depth > 5) { /* mod.rs:677-683: rand01() is drawn before the depth test */
        if (rand01(S, 0) < max_reflection && new_depth < MAX_DEPTH) color = v_scale(color, 1.0f / max_reflection);
        else return o->emission;
    }
    v3 inner;
    switch (o->refl) {
    case REFL_DIFFUSE: {
        ray_t r = {hit.x, diffuse_dir(S, nl)};
        inner = v_mul(color, radiance_rec(&r, new_depth, code, S));
    } break;
    case REFL_SPECULAR: {
        ray_t r = {hit.x, reflect_dir(ray->d, hit.n)};
        inner = v_mul(color, radiance_rec(&r, new_depth, code, S));
    } break;
    default: {
        ray_t refl = {hit.x, reflect_dir(ray->d, hit.n)};
        fresnel_t f;
        if (!refract_terms(ray->d, hit.n, nl, &f)) {
            inner = v_mul(color, radiance_rec(&refl, new_depth, code, S));
        } else {
            ray_t tr = {hit.x, f.tdir};
            if (new_depth > 2) {
                if (rand01(S, 3) < f.p) inner = v_scale(v_mul(color, radiance_rec(&refl, new_depth, code, S)), f.rp);
                else inner = v_scale(v_mul(color, radiance_rec(&tr, new_depth, code, S)), f.tp);
            } else { /* mod.rs:776-785: reflection child evaluated first */
                v3 a = v_scale(radiance_rec(&refl, new_depth, code, S), f.re);
                v3 b = v_scale(radiance_rec(&tr, new_depth, code | (1 << (new_depth - 1)), S), f.tr);
                inner = v_mul(color, v_add(a, b));
            }
        }
    } break;
    }
    return v_add(o->emission, inner);
}

/* forward (throughput) form: bit-exact twin of the CUDA integrator; same paths, same draws,
 * different association of the colour products (differs from radiance_rec by rounding only). */
static v3 radiance_fwd(const ray_t *ray0, sampler_t *S) {
    struct { ray_t r; v3 T; int depth, code; } stack[2];
    int sp = 0;
    v3 Lc[4] = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
    v3 T = V(1.0f, 1.0f, 1.0f);
    ray_t ray = *ray0;
    int depth = 0, code = 0;
    for (;;) {
        hit_t hit;
        S->n_segments++;
        sampler_event(S, ((uint32_t)code << 4) | (uint32_t)(depth + 1));
        int id = intersect_scene(S->sc, &ray, &hit, S);
        if (id >= 0) {
            const obj_t *o = &S->sc->objs[id];
            v3 color = o->color;
            float max_reflection = fmaxf(color.x, fmaxf(color.y, color.z));
            v3 nl = v_dot(hit.n, ray.d) < 0.0f ? hit.n : v_scale(hit.n, -1.0f);
            int new_depth = depth + 1;
            int alive = 1;
            if (new_depth > 5) {
                if (rand01(S, 0) < max_reflection && new_depth < MAX_DEPTH) color = v_scale(color, 1.0f / max_reflection);
                else alive = 0;
            }
            Lc[code] = v_add(Lc[code], v_mul(T, o->emission));
            if (alive) {
                v3 Tc = v_mul(T, color);
                if (o->refl == REFL_DIFFUSE) {
                    ray.d = diffuse_dir(S, nl); ray.o = hit.x; T = Tc; depth = new_depth;
                    continue;
                } else if (o->refl == REFL_SPECULAR) {
                    ray.d = reflect_dir(ray.d, hit.n); ray.o = hit.x; T = Tc; depth = new_depth;
                    continue;
                } else {
                    v3 rd = reflect_dir(ray.d, hit.n);
                    fresnel_t f;
                    if (!refract_terms(ray.d, hit.n, nl, &f)) {
                        T = Tc; ray.d = rd;
                    } else if (new_depth > 2) {
                        if (rand01(S, 3) < f.p) { T = v_scale(Tc, f.rp); ray.d = rd; }
                        else { T = v_scale(Tc, f.tp); ray.d = f.tdir; }
                    } else {
                        stack[sp].r.o = hit.x; stack[sp].r.d = f.tdir;
                        stack[sp].T = v_scale(Tc, f.tr); stack[sp].depth = new_depth;
                        stack[sp].code = code | (1 << (new_depth - 1)); sp++;
                        T = v_scale(Tc, f.re); ray.d = rd;
                    }
                    ray.o = hit.x; depth = new_depth;
                    continue;
                }
            }
        }
        if (sp == 0) break;
        sp--; ray = stack[sp].r; T = stack[sp].T; depth = stack[sp].depth; code = stack[sp].code;
    }
    return v_add(v_add(v_add(Lc[0], Lc[1]), Lc[2]), Lc[3]);
}

/* ------------------------------------------------------------------------------------------ */
/* camera  mod.rs:211-232 and the per-sample ray of render_pixel  mod.rs:805-843                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { v3 lens_center, su, sv, sensor_origin; } cam_frame_t;
static cam_frame_t camera_frame(const struct pto_scene *sc) {
    cam_frame_t f;
    v3 dir = sc->cam_dir;
    f.sensor_origin = sc->cam_pos;
    f.lens_center = v_add(sc->cam_pos, v_scale(dir, sc->focal_length));
    v3 su = v_normalize(v_cross(dir, fabsf(dir.y) < 0.9f ? V(0.0f, 1.0f, 0.0f) : V(0.0f, 0.0f, 1.0f)));
    v3 sv = v_cross(su, dir);
    float sensor_height = sc->sensor_width / sc->aspect_ratio;
    f.su = v_scale(su, sc->sensor_width);
    f.sv = v_scale(sv, sensor_height);
    return f;
}
static inline ray_t camera_ray(const cam_frame_t *f, int W, int H, int x, int y, float xsub, float ysub,
                               float xfilter, float yfilter) {
    float sx = ((float)x + 0.5f * (0.5f + xsub + xfilter)) / (float)W - 0.5f;
    float sy = ((float)y + 0.5f * (0.5f + ysub + yfilter)) / (float)H - 0.5f;
    v3 sensor_pos = v_add(v_add(f->sensor_origin, v_scale(f->su, sx)), v_scale(f->sv, sy));
    ray_t r;
    r.d = v_normalize(v_sub(f->lens_center, sensor_pos));
    r.o = f->lens_center;
    return r;
}
static inline float tent(float r) { /* mod.rs:820-830 */
    return r < 1.0f ? sqrtf(r) - 1.0f : 1.0f - sqrtf(2.0f - r);
}

/* render_pixel  mod.rs:794-857, without the final /spp and clamp (see pto_resolve) */
static v3 render_pixel_sum(const struct pto_scene *sc, const cam_frame_t *cf, int W, int H, uint32_t pixel_index,
                           uint64_t spp_begin, uint64_t spp_count, v3 acc, const pto_render_cfg *cfg, sampler_t *S) {
    (void)sc;
    int y = H - 1 - (int)(pixel_index / (uint32_t)W);
    int x = (int)(pixel_index % (uint32_t)W);
    sampler_seed_pixel(S, cfg->seed, pixel_index);
    for (uint64_t s = spp_begin; s < spp_begin + spp_count; ++s) {
        float ysub = (float)((s / 2) % 2);
        float xsub = (float)(s % 2);
        S->ctr_slo = (uint32_t)s; S->ctr_shi = (uint32_t)(s >> 32);
        sampler_event(S, 0);
        float r1 = 2.0f * rand01(S, 0);
        float r2 = 2.0f * rand01(S, 1);
        ray_t ray = camera_ray(cf, W, H, x, y, xsub, ysub, tent(r1), tent(r2));
        v3 rad = cfg->accum_mode == PTO_ACCUM_FORWARD ? radiance_fwd(&ray, S) : radiance_rec(&ray, 0, 0, S);
        acc = v_add(acc, rad);
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------ */
/* frame driver: render()  mod.rs:928-1024 (shuffled pixel list, dynamic scheduling)             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const struct pto_scene *sc;
    cam_frame_t cf;
    int W, H;
    uint64_t spp_begin, spp_count;
    const pto_render_cfg *cfg;
    const uint32_t *order;
    uint32_t npix, p0;
    atomic_uint next;
    float *sum;
    atomic_ullong stats[4];
} job_t;

static void *render_worker(void *arg) {
    job_t *J = (job_t *)arg;
    sampler_t S;
    memset(&S, 0, sizeof S);
    S.sc = J->sc; S.rng_mode = J->cfg->rng_mode; S.sincos_mode = J->cfg->sincos_mode;
    const uint32_t chunk = 16;
    for (;;) {
        uint32_t b = atomic_fetch_add(&J->next, chunk);
        if (b >= J->npix) break;
        uint32_t e = b + chunk < J->npix ? b + chunk : J->npix;
        for (uint32_t k = b; k < e; ++k) {
            uint32_t p = J->order ? J->order[k] : J->p0 + k;
            float *px = J->sum + 3 * (size_t)p;
            v3 acc = render_pixel_sum(J->sc, &J->cf, J->W, J->H, p, J->spp_begin, J->spp_count, V(px[0], px[1], px[2]),
                                      J->cfg, &S);
            px[0] = acc.x; px[1] = acc.y; px[2] = acc.z;
        }
    }
    atomic_fetch_add(&J->stats[0], S.n_segments);
    atomic_fetch_add(&J->stats[1], S.n_sphere);
    atomic_fetch_add(&J->stats[2], S.n_gate);
    atomic_fetch_add(&J->stats[3], S.n_tri);
    return NULL;
}

int pto_render(const pto_scene *sc, int W, int H, uint64_t spp_begin, uint64_t spp_count, const pto_render_cfg *cfg,
               float *sum_rgb, uint64_t *stats4) {
    return pto_render_region(sc, W, H, 0, (uint32_t)W * (uint32_t)H, spp_begin, spp_count, cfg, sum_rgb, stats4);
}

int pto_render_region(const pto_scene *sc, int W, int H, uint32_t pixel_begin, uint32_t pixel_count,
                      uint64_t spp_begin, uint64_t spp_count, const pto_render_cfg *cfg, float *sum_rgb,
                      uint64_t *stats4) {
    if (!sc || W <= 0 || H <= 0 || !sum_rgb || !cfg) return -1;
    if ((uint64_t)pixel_begin + pixel_count > (uint64_t)W * (uint64_t)H) return -1;
    job_t J;
    memset(&J, 0, sizeof J);
    J.sc = sc; J.cf = camera_frame(sc); J.W = W; J.H = H;
    J.spp_begin = spp_begin; J.spp_count = spp_count; J.cfg = cfg; J.sum = sum_rgb;
    J.npix = pixel_count; J.p0 = pixel_begin;
    int nthreads = cfg->threads > 0 ? cfg->threads : 1;
    uint32_t *order = NULL;
    if (cfg->rng_mode == PTO_RNG_MOCK) {
        nthreads = 1; /* mod.rs:1017-1018: serial, natural order */
    } else if (cfg->shuffle) { /* mod.rs:1021-1022 */
        order = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)pixel_count);
        for (uint32_t i = 0; i < pixel_count; ++i) order[i] = pixel_begin + i;
        uint64_t st = splitmix64(cfg->seed ^ 0x5bd1e995u);
        for (uint32_t i = pixel_count; i > 1; --i) {
            st = splitmix64(st);
            uint32_t j = (uint32_t)(st % i);
            uint32_t tmp = order[i - 1]; order[i - 1] = order[j]; order[j] = tmp;
        }
        J.order = order;
    }
    atomic_init(&J.next, 0);
    for (int i = 0; i < 4; ++i) atomic_init(&J.stats[i], 0);
    if (nthreads == 1) {
        render_worker(&J);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, render_worker, &J);
        for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    free(order);
    if (stats4) for (int i = 0; i < 4; ++i) stats4[i] = atomic_load(&J.stats[i]);
    return 0;
}

void pto_mock_reset(void) { atomic_store(&g_mock_index, 0); }

/* mod.rs:849-856: radiance/spp then clamp each channel to [0,1] */
void pto_resolve(const float *sum_rgb, size_t n_floats, uint64_t spp, float *mean_rgb) {
    float d = (float)spp;
    for (size_t i = 0; i < n_floats; ++i) {
        float v = sum_rgb[i] / d;
        mean_rgb[i] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    }
}
/* mod.rs:57-63, pinned by test.rs:29-35 */
uint32_t pto_to_int_with_gamma_correction(float x) {
    float c = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    float g = powf(c, 1.0f / 2.2f);
    return (uint32_t)(255.0f * g + 0.5f);
}

/* PPM writer mod.rs:1042-1076 (P3, two comment lines, pixels in reverse index order) */
int pto_write_ppm(const char *path, const float *mean_rgb, int W, int H, uint64_t spp, const char *scene_id,
                  uint64_t seconds) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "P3\n# samplesPerPixel: %llu, resolution_y: %d, scene_id: %s\n", (unsigned long long)spp, H, scene_id);
    fprintf(f, "# rendering time: %llu s\n%d %d\n%d\n", (unsigned long long)seconds, W, H, 255);
    for (long i = (long)W * H - 1; i >= 0; --i)
        fprintf(f, "%u %u %u ", pto_to_int_with_gamma_correction(mean_rgb[3 * i]),
                pto_to_int_with_gamma_correction(mean_rgb[3 * i + 1]), pto_to_int_with_gamma_correction(mean_rgb[3 * i + 2]));
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* parity hooks: arbitrary rays, deterministic primary rays, single-ray radiance                 */
/* ------------------------------------------------------------------------------------------ */
int pto_intersect(const pto_scene *sc, const float *rays, int n, int32_t *obj, int32_t *tri, float *t, float *point,
                  float *normal) {
    for (int i = 0; i < n; ++i) {
        ray_t r = {V(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), V(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])};
        hit_t h;
        int id = intersect_scene(sc, &r, &h, NULL);
        obj[i] = id;
        if (id < 0) { h.t = 0.0f; h.x = V(0, 0, 0); h.n = V(0, 0, 0); h.tri = -1; }
        if (tri) tri[i] = h.tri;
        if (t) t[i] = h.t;
        if (point) { point[3 * i] = h.x.x; point[3 * i + 1] = h.x.y; point[3 * i + 2] = h.x.z; }
        if (normal) { normal[3 * i] = h.n.x; normal[3 * i + 1] = h.n.y; normal[3 * i + 2] = h.n.z; }
    }
    return 0;
}

/* centre rays: xsub = ysub = xfilter = yfilter = 0 in mod.rs:833-838 */
int pto_primary_rays(const pto_scene *sc, int W, int H, float *rays) {
    cam_frame_t cf = camera_frame(sc);
    for (int p = 0; p < W * H; ++p) {
        int y = H - 1 - p / W, x = p % W;
        ray_t r = camera_ray(&cf, W, H, x, y, 0.0f, 0.0f, 0.0f, 0.0f);
        float *o = rays + 6 * (size_t)p;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
    }
    return 0;
}
int pto_primary_hits(const pto_scene *sc, int W, int H, int32_t *obj, int32_t *tri, float *t) {
    cam_frame_t cf = camera_frame(sc);
    for (int p = 0; p < W * H; ++p) {
        int y = H - 1 - p / W, x = p % W;
        ray_t r = camera_ray(&cf, W, H, x, y, 0.0f, 0.0f, 0.0f, 0.0f);
        hit_t h;
        int id = intersect_scene(sc, &r, &h, NULL);
        obj[p] = id; tri[p] = id < 0 ? -1 : h.tri; t[p] = id < 0 ? 0.0f : h.t;
    }
    return 0;
}
void pto_camera_frame(const pto_scene *sc, float *out12) {
    cam_frame_t f = camera_frame(sc);
    v3 a[4] = {f.lens_center, f.su, f.sv, f.sensor_origin};
    for (int i = 0; i < 4; ++i) { out12[3 * i] = a[i].x; out12[3 * i + 1] = a[i].y; out12[3 * i + 2] = a[i].z; }
}

/* test.rs:146-183: mean of n radiance() samples along one fixed ray */
int pto_radiance_mean(const pto_scene *sc, const float *ray6, uint64_t n, const pto_render_cfg *cfg, float *mean3) {
    sampler_t S;
    memset(&S, 0, sizeof S);
    S.sc = sc; S.rng_mode = cfg->rng_mode; S.sincos_mode = cfg->sincos_mode;
    sampler_seed_pixel(&S, cfg->seed, 0);
    ray_t r = {V(ray6[0], ray6[1], ray6[2]), V(ray6[3], ray6[4], ray6[5])};
    /* accumulate in double here: this is a statistical check of the estimator, not of the frame sum */
    double acc[3] = {0, 0, 0};
    for (uint64_t s = 0; s < n; ++s) {
        S.ctr_slo = (uint32_t)s; S.ctr_shi = (uint32_t)(s >> 32);
        S.event = 0;
        v3 v = cfg->accum_mode == PTO_ACCUM_FORWARD ? radiance_fwd(&r, &S) : radiance_rec(&r, 0, 0, &S);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z;
    }
    for (int i = 0; i < 3; ++i) mean3[i] = (float)(acc[i] / (double)n);
    return 0;
}

/* glam op probes for test.rs:3-27 */
float pto_vec_dot(const float *a, const float *b) { return v_dot(V(a[0], a[1], a[2]), V(b[0], b[1], b[2])); }
float pto_vec_length(const float *a) { return v_length(V(a[0], a[1], a[2])); }
void pto_vec_cross(const float *a, const float *b, float *o) {
    v3 r = v_cross(V(a[0], a[1], a[2]), V(b[0], b[1], b[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void pto_vec_normalize(const float *a, float *o) {
    v3 r = v_normalize(V(a[0], a[1], a[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void pto_vec_divs(const float *a, float s, float *o) {
    v3 r = v_divs(V(a[0], a[1], a[2]), s); o[0] = r.x; o[1] = r.y; o[2] = r.z;
}

/* ------------------------------------------------------------------------------------------ */
/* minimal JSON reader (serde_json layout of mod.rs:85-90, 236-241, 297-302)                     */
/* ------------------------------------------------------------------------------------------ */
typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } jtype;
typedef struct jnode {
    jtype t;
    double num;
    char *str;
    struct jnode **items;
    char **keys;
    int n, cap;
} jnode;

typedef struct { const char *p, *end; char err[160]; } jparser;

static void j_free(jnode *n) {
    if (!n) return;
    for (int i = 0; i < n->n; ++i) { j_free(n->items[i]); if (n->keys) free(n->keys[i]); }
    free(n->items); free(n->keys); free(n->str); free(n);
}
static void j_ws(jparser *P) { while (P->p < P->end && isspace((unsigned char)*P->p)) P->p++; }
static jnode *j_new(jtype t) { jnode *n = (jnode *)calloc(1, sizeof *n); n->t = t; return n; }
static void j_push(jnode *a, jnode *v, char *key) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 8;
        a->items = (jnode **)realloc(a->items, sizeof(jnode *) * (size_t)a->cap);
        if (a->t == J_OBJ) a->keys = (char **)realloc(a->keys, sizeof(char *) * (size_t)a->cap);
    }
    a->items[a->n] = v;
    if (a->t == J_OBJ) a->keys[a->n] = key;
    a->n++;
}
static char *j_string(jparser *P) {
    if (*P->p != '"') { snprintf(P->err, sizeof P->err, "expected string"); return NULL; }
    P->p++;
    size_t cap = 32, len = 0;
    char *s = (char *)malloc(cap);
    while (P->p < P->end && *P->p != '"') {
        char c = *P->p++;
        if (c == '\\' && P->p < P->end) {
            char e = *P->p++;
            switch (e) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break;
                         case 'b': c = '\b'; break; case 'f': c = '\f'; break; default: c = e; }
        }
        if (len + 2 > cap) { cap *= 2; s = (char *)realloc(s, cap); }
        s[len++] = c;
    }
    if (P->p >= P->end) { free(s); snprintf(P->err, sizeof P->err, "unterminated string"); return NULL; }
    P->p++;
    s[len] = 0;
    return s;
}
static jnode *j_value(jparser *P) {
    j_ws(P);
    if (P->p >= P->end) { snprintf(P->err, sizeof P->err, "unexpected end"); return NULL; }
    char c = *P->p;
    if (c == '{' || c == '[') {
        int is_obj = c == '{';
        char close = is_obj ? '}' : ']';
        jnode *n = j_new(is_obj ? J_OBJ : J_ARR);
        P->p++;
        j_ws(P);
        if (P->p < P->end && *P->p == close) { P->p++; return n; }
        for (;;) {
            char *key = NULL;
            j_ws(P);
            if (is_obj) {
                key = j_string(P);
                if (!key) { j_free(n); return NULL; }
                j_ws(P);
                if (P->p >= P->end || *P->p != ':') { free(key); j_free(n); snprintf(P->err, sizeof P->err, "expected ':'"); return NULL; }
                P->p++;
            }
            jnode *v = j_value(P);
            if (!v) { free(key); j_free(n); return NULL; }
            j_push(n, v, key);
            j_ws(P);
            if (P->p < P->end && *P->p == ',') { P->p++; continue; }
            if (P->p < P->end && *P->p == close) { P->p++; return n; }
            j_free(n);
            snprintf(P->err, sizeof P->err, "expected ',' or '%c'", close);
            return NULL;
        }
    }
    if (c == '"') { char *s = j_string(P); if (!s) return NULL; jnode *n = j_new(J_STR); n->str = s; return n; }
    if (!strncmp(P->p, "null", 4)) { P->p += 4; return j_new(J_NULL); }
    if (!strncmp(P->p, "true", 4)) { P->p += 4; jnode *n = j_new(J_BOOL); n->num = 1; return n; }
    if (!strncmp(P->p, "false", 5)) { P->p += 5; return j_new(J_BOOL); }
    char *endp = NULL;
    double d = strtod(P->p, &endp); /* serde_json parses f64 ... */
    if (endp == P->p) { snprintf(P->err, sizeof P->err, "bad token at '%.12s'", P->p); return NULL; }
    P->p = endp;
    jnode *n = j_new(J_NUM);
    n->num = d;
    return n;
}
static const jnode *j_get(const jnode *o, const char *key) {
    if (!o || o->t != J_OBJ) return NULL;
    for (int i = 0; i < o->n; ++i) if (!strcmp(o->keys[i], key)) return o->items[i];
    return NULL;
}
/* ... then narrows to f32 (`as f32`) */
static int j_f32(const jnode *n, float *out) { if (!n || n->t != J_NUM) return 0; *out = (float)n->num; return 1; }
static int j_v3(const jnode *n, v3 *out) {
    if (!n || n->t != J_ARR || n->n != 3) return 0;
    return j_f32(n->items[0], &out->x) && j_f32(n->items[1], &out->y) && j_f32(n->items[2], &out->z);
}

/* ------------------------------------------------------------------------------------------ */
/* Mesh::new  mod.rs:450-499 (bounding sphere centre = min + max*0.5, sic)                       */
/* ------------------------------------------------------------------------------------------ */
static void mesh_bounds(const tri_t *tris, int n, v3 *bs_pos, float *bs_radius) {
    v3 mn = V(INFINITY, INFINITY, INFINITY), mx = V(-INFINITY, -INFINITY, -INFINITY);
    for (int i = 0; i < n; ++i) {
        const v3 *vs[3] = {&tris[i].a, &tris[i].b, &tris[i].c};
        for (int k = 0; k < 3; ++k) {
            const v3 *p = vs[k];
            if (p->x < mn.x) mn.x = p->x;
            if (p->y < mn.y) mn.y = p->y;
            if (p->z < mn.z) mn.z = p->z;
            if (p->x > mx.x) mx.x = p->x;
            if (p->y > mx.y) mx.y = p->y;
            if (p->z > mx.z) mx.z = p->z;
        }
    }
    v3 c = V(mn.x + mx.x * 0.5f, mn.y + mx.y * 0.5f, mn.z + mx.z * 0.5f);
    float r0 = v_length(v_sub(mn, c)), r1 = v_length(v_sub(mx, c));
    *bs_pos = c;
    *bs_radius = r1 >= r0 ? r1 : r0; /* max_by keeps the last maximum */
}

/* load_off  load_off.rs:8-85 */
static int off_next_line(FILE *f, char *buf, size_t n) { /* skips blank and '#' lines after trim (load_off.rs:12-20) */
    for (;;) {
        if (!fgets(buf, (int)n, f)) return 0;
        char *s = buf;
        while (*s && isspace((unsigned char)*s)) s++;
        size_t len = strlen(s);
        while (len && isspace((unsigned char)s[len - 1])) s[--len] = 0;
        if (len == 0 || s[0] == '#') continue;
        memmove(buf, s, len + 1);
        return 1;
    }
}
static int parse_usize_tok(const char *tok, long *out) { /* str::parse::<usize>: digits only, optional '+' */
    const char *p = tok;
    if (*p == '+') p++;
    if (!*p) return 0;
    long v = 0;
    for (; *p; ++p) { if (!isdigit((unsigned char)*p)) return 0; v = v * 10 + (*p - '0'); }
    *out = v;
    return 1;
}
static int load_off(const char *path, float scale, int fan_polygons, tri_t **tris_out, int *ntris_out, char *err, int errlen) {
    FILE *f = fopen(path, "r");
    if (!f) { snprintf(err, (size_t)errlen, "cannot open %s", path); return -1; }
    char line[1024];
    int rc = -1;
    v3 *verts = NULL;
    tri_t *tris = NULL;
    if (!off_next_line(f, line, sizeof line) || strcmp(line, "OFF")) { snprintf(err, (size_t)errlen, "Invalid header"); goto done; }
    if (!off_next_line(f, line, sizeof line)) { snprintf(err, (size_t)errlen, "Invalid element counts"); goto done; }
    long counts[3]; int nc = 0, ok = 1;
    for (char *tok = strtok(line, " \t\r\n"); tok; tok = strtok(NULL, " \t\r\n")) {
        long v;
        if (nc < 3) { if (!parse_usize_tok(tok, &v)) ok = 0; else counts[nc] = v; }
        nc++;
    }
    if (nc != 3 || !ok) { snprintf(err, (size_t)errlen, "Invalid element counts"); goto done; }
    long nv = counts[0], nf = counts[1];
    verts = (v3 *)malloc(sizeof(v3) * (size_t)(nv > 0 ? nv : 1));
    for (long i = 0; i < nv; ++i) {
        if (!off_next_line(f, line, sizeof line)) { snprintf(err, (size_t)errlen, "unexpected EOF in vertices"); goto done; }
        float c[3]; int n = 0; ok = 1;
        for (char *tok = strtok(line, " \t\r\n"); tok; tok = strtok(NULL, " \t\r\n")) {
            if (n < 3) { char *e; c[n] = strtof(tok, &e); if (e == tok || *e) ok = 0; } /* str::parse::<f32>, correctly rounded */
            n++;
        }
        if (n != 3 || !ok) { snprintf(err, (size_t)errlen, "Invalid vertex coordinates"); goto done; }
        verts[i] = v_scale(V(c[0], c[1], c[2]), scale); /* load_off.rs:52 */
    }
    long cap = nf > 0 ? nf : 1, nt = 0;
    tris = (tri_t *)malloc(sizeof(tri_t) * (size_t)cap);
    for (long i = 0; i < nf; ++i) {
        if (!off_next_line(f, line, sizeof line)) { snprintf(err, (size_t)errlen, "unexpected EOF in faces"); goto done; }
        char copy[1024];
        snprintf(copy, sizeof copy, "%s", line);
        long idx[64]; int n = 0; ok = 1;
        for (char *tok = strtok(line, " \t\r\n"); tok; tok = strtok(NULL, " \t\r\n")) {
            if (n < 64 && !parse_usize_tok(tok, &idx[n])) { if (n < 4) ok = 0; else idx[n] = -1; } /* first four tokens are unwrap()ed */
            n++;
        }
        if (n >= 4 && ok && idx[0] > 3 && fan_polygons) { /* opt-in, not reference behaviour: fan (v0, v_i, v_i+1) */
            long cnt = idx[0];
            int good = cnt < 63 && n >= cnt + 1;
            for (long k = 1; good && k <= cnt; ++k) good = idx[k] >= 0 && idx[k] < nv;
            if (!good) { snprintf(err, (size_t)errlen, "Invalid face: %.100s", copy); goto done; }
            for (long k = 2; k < cnt; ++k) {
                if (nt == cap) { cap *= 2; tris = (tri_t *)realloc(tris, sizeof(tri_t) * (size_t)cap); }
                tris[nt].a = verts[idx[1]]; tris[nt].b = verts[idx[k]]; tris[nt].c = verts[idx[k + 1]]; nt++;
            }
            continue;
        }
        if (n < 4 || !ok || idx[0] != 3 || idx[1] >= nv || idx[2] >= nv || idx[3] >= nv) {
            snprintf(err, (size_t)errlen, "Invalid face: %.100s", copy);
            goto done;
        }
        if (nt == cap) { cap *= 2; tris = (tri_t *)realloc(tris, sizeof(tri_t) * (size_t)cap); }
        tris[nt].a = verts[idx[1]]; tris[nt].b = verts[idx[2]]; tris[nt].c = verts[idx[3]]; nt++;
    }
    *tris_out = tris; *ntris_out = (int)nt; tris = NULL; rc = 0;
done:
    free(verts); free(tris); fclose(f);
    return rc;
}

/* SceneDescriptor::load + to_data  mod.rs:92-110, 304-318 */
void pto_scene_free(pto_scene *sc) {
    if (!sc) return;
    for (int i = 0; i < sc->nobjs; ++i) free(sc->objs[i].tris);
    free(sc->objs); free(sc);
}
#define FAIL(...) do { snprintf(err, (size_t)errlen, __VA_ARGS__); goto fail; } while (0)
pto_scene *pto_scene_load(const char *json_path, const char *base_dir, char *err, int errlen) {
    return pto_scene_load_ex(json_path, base_dir, 0, err, errlen);
}
pto_scene *pto_scene_load_ex(const char *json_path, const char *base_dir, int fan_polygons, char *err, int errlen) {
    char dummy[8];
    if (!err) { err = dummy; errlen = sizeof dummy; }
    err[0] = 0;
    FILE *f = fopen(json_path, "rb");
    if (!f) { snprintf(err, (size_t)errlen, "cannot open %s", json_path); return NULL; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char *text = (char *)malloc((size_t)sz + 1);
    if (fread(text, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(text); snprintf(err, (size_t)errlen, "read error"); return NULL; }
    fclose(f);
    text[sz] = 0;
    jparser P = {text, text + sz, {0}};
    jnode *root = j_value(&P);
    pto_scene *sc = NULL;
    if (!root) { snprintf(err, (size_t)errlen, "json: %s", P.err); free(text); return NULL; }
    sc = (pto_scene *)calloc(1, sizeof *sc);
    const jnode *id = j_get(root, "id"), *objs = j_get(root, "objects"), *cam = j_get(root, "camera");
    if (!id || id->t != J_STR) FAIL("missing field `id`");
    snprintf(sc->id, sizeof sc->id, "%s", id->str);
    if (!objs || objs->t != J_ARR) FAIL("missing field `objects`");
    if (!cam || cam->t != J_OBJ) FAIL("missing field `camera`");
    if (!j_v3(j_get(cam, "position"), &sc->cam_pos) || !j_v3(j_get(cam, "direction"), &sc->cam_dir) ||
        !j_f32(j_get(cam, "focal_length"), &sc->focal_length) || !j_f32(j_get(cam, "sensor_width"), &sc->sensor_width) ||
        !j_f32(j_get(cam, "aspect_ratio"), &sc->aspect_ratio))
        FAIL("bad camera");
    sc->objs = (obj_t *)calloc((size_t)(objs->n ? objs->n : 1), sizeof(obj_t));
    for (int i = 0; i < objs->n; ++i) {
        const jnode *jo = objs->items[i];
        obj_t *o = &sc->objs[i];
        sc->nobjs = i + 1;
        const jnode *ty = j_get(jo, "type_"), *mat = j_get(jo, "material");
        if (!ty || ty->t != J_OBJ || ty->n != 1) FAIL("object %d: bad type_", i);
        if (!j_v3(j_get(jo, "position"), &o->position)) FAIL("object %d: bad position", i);
        if (!mat || !j_v3(j_get(mat, "color"), &o->color) || !j_v3(j_get(mat, "emmission"), &o->emission))
            FAIL("object %d: bad material", i);
        const jnode *rt = j_get(mat, "reflect_type");
        if (!rt || rt->t != J_STR) FAIL("object %d: bad reflect_type", i);
        if (!strcmp(rt->str, "Diffuse")) o->refl = REFL_DIFFUSE;
        else if (!strcmp(rt->str, "Specular")) o->refl = REFL_SPECULAR;
        else if (!strcmp(rt->str, "Refract")) o->refl = REFL_REFRACT;
        else FAIL("object %d: unknown variant `%s`", i, rt->str);
        const char *kind = ty->keys[0];
        const jnode *body = ty->items[0];
        if (!strcmp(kind, "Sphere")) {
            o->type = OBJ_SPHERE;
            if (!j_f32(j_get(body, "radius"), &o->radius)) FAIL("object %d: bad radius", i);
        } else if (!strcmp(kind, "MeshFile")) {
            o->type = OBJ_MESH;
            const jnode *jp = j_get(body, "path");
            float scale;
            if (!jp || jp->t != J_STR || !j_f32(j_get(body, "scale"), &scale)) FAIL("object %d: bad MeshFile", i);
            char path[1024];
            if (jp->str[0] == '/' || !base_dir || !base_dir[0]) snprintf(path, sizeof path, "%s", jp->str);
            else snprintf(path, sizeof path, "%s/%s", base_dir, jp->str);
            char e2[200];
            if (load_off(path, scale, fan_polygons, &o->tris, &o->ntris, e2, sizeof e2)) FAIL("object %d: %s", i, e2);
            mesh_bounds(o->tris, o->ntris, &o->bs_pos, &o->bs_radius); /* Mesh::new, load_off.rs:84 */
        } else if (!strcmp(kind, "Mesh")) {
            o->type = OBJ_MESH;
            const jnode *jt = j_get(body, "triangles"), *bs = j_get(body, "bounding_sphere");
            if (!jt || jt->t != J_ARR) FAIL("object %d: bad triangles", i);
            if (!j_get(body, "bounding_box")) FAIL("object %d: missing field `bounding_box`", i);
            /* bounding sphere is deserialised, not recomputed (mod.rs:441-448) */
            if (!bs || !j_v3(j_get(bs, "position"), &o->bs_pos) || !j_f32(j_get(bs, "radius"), &o->bs_radius))
                FAIL("object %d: bad bounding_sphere", i);
            o->ntris = jt->n;
            o->tris = (tri_t *)malloc(sizeof(tri_t) * (size_t)(jt->n ? jt->n : 1));
            for (int k = 0; k < jt->n; ++k) {
                const jnode *t = jt->items[k];
                if (!j_v3(j_get(t, "a"), &o->tris[k].a) || !j_v3(j_get(t, "b"), &o->tris[k].b) || !j_v3(j_get(t, "c"), &o->tris[k].c))
                    FAIL("object %d: bad triangle %d", i, k);
            }
        } else FAIL("object %d: unknown variant `%s`", i, kind);
    }
    j_free(root); free(text);
    return sc;
fail:
    j_free(root); free(text); pto_scene_free(sc);
    return NULL;
}

int pto_scene_counts(const pto_scene *sc, int *nobjs, int *nspheres, int *nmeshes, int *ntris) {
    int s = 0, m = 0, t = 0;
    for (int i = 0; i < sc->nobjs; ++i) {
        if (sc->objs[i].type == OBJ_SPHERE) s++; else { m++; t += sc->objs[i].ntris; }
    }
    if (nobjs) *nobjs = sc->nobjs;
    if (nspheres) *nspheres = s;
    if (nmeshes) *nmeshes = m;
    if (ntris) *ntris = t;
    return 0;
}
const char *pto_scene_id(const pto_scene *sc) { return sc->id; }
int pto_scene_mesh_bounds(const pto_scene *sc, int obj, float *pos3, float *radius) {
    if (obj < 0 || obj >= sc->nobjs || sc->objs[obj].type != OBJ_MESH) return -1;
    pos3[0] = sc->objs[obj].bs_pos.x; pos3[1] = sc->objs[obj].bs_pos.y; pos3[2] = sc->objs[obj].bs_pos.z;
    *radius = sc->objs[obj].bs_radius;
    return 0;
}

#ifdef PTO_MAIN
/* CLI in the shape of the reference's dead cmd_render.rs:17-44:  render <spp> <res_y> <scene-id> [threads] */
int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <spp> <res_y> <scene.json> [threads] [out.ppm]\n", argv[0]); return 1; }
    uint64_t spp = strtoull(argv[1], NULL, 10);
    int H = atoi(argv[2]), W = H * 3 / 2; /* main.rs:176 */
    char err[256];
    pto_scene *sc = pto_scene_load(argv[3], ".", err, sizeof err);
    if (!sc) { fprintf(stderr, "%s\n", err); return 1; }
    pto_render_cfg cfg = {PTO_RNG_SEQ, PTO_SINCOS_LIBM, PTO_ACCUM_RECURSIVE, (uint64_t)time(NULL), argc > 4 ? atoi(argv[4]) : 1, 1};
    float *sum = (float *)calloc((size_t)W * H * 3, sizeof(float));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    uint64_t st[4];
    pto_render(sc, W, H, 0, spp, &cfg, sum, st);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    pto_resolve(sum, (size_t)W * H * 3, spp, sum);
    printf("scene %s %dx%d spp %llu: %.3f s, %.3f Mpaths/s, %.3f Mseg/s\n", sc->id, W, H, (unsigned long long)spp, sec,
           (double)W * H * (double)spp / sec * 1e-6, (double)st[0] / sec * 1e-6);
    pto_write_ppm(argc > 5 ? argv[5] : "oracle_latest.ppm", sum, W, H, spp, sc->id, (uint64_t)sec);
    free(sum);
    pto_scene_free(sc);
    return 0;
}
#endif
