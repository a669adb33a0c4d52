// pt_api.cu -- the C ABI (include/ptb.h): context, scene flatten + upload, render driver, parity hooks.
//
// Replaces the render thread of render() (src/render/mod.rs:984-1026): instead of shuffling a pixel list into a
// rayon pool it launches the persistent integrator kernel in sample batches, polling the cancel flag between
// launches (mod.rs:1003) and publishing progress (mod.rs:850).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptb.h"
#include "pt_bvh.cuh"
#include "pt_bvh_build.h"
#include "pt_launch.h"
#include "pt_wavefront.h"
#include "scene_io.hpp"

using namespace ptb;

namespace {

thread_local std::string g_thread_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t resize(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t st) {
        cudaError_t e = resize(std::max<size_t>(h.size(), 1));
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
};

inline float4 f4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline float ibits(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float ubits(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline V3 v3(const float *p) { return mk3(p[0], p[1], p[2]); }

}  // namespace

struct ptb_scene {
    HostScene host;
};

struct ptb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_done = nullptr;  // end of the last render on whatever stream it ran: the next user of the context's buffers waits for it
    bool render_in_flight = false;
    std::string err;
    bool has_scene = false;
    DScene ds{};
    DevBuf<float4> loose_obj, loose_tri, obj_gate, mat_color, mat_emis;
    BvhDevice bvh;
    BvhOptions bvh_opt;
    WfWorkspace wf;
    int integrator = 0;             // 0 = auto, 1 = megakernel (one lane per pixel), 2 = wavefront, 3 = sample-parallel megakernel
    WfOptions wf_opt;               // wavefront integrator tuning (paths in flight, refill threshold, trace CTA size)
    int regen_batch = REGEN_BATCH;
    double quad_min_ratio = 0.125;  // one-pair meshes with gate radius >= this x scene diagonal are tested without the warp vote
    DevBuf<float> fb, scratch_f, preview;
    DevBuf<float4> sample_L;        // sample-parallel megakernel: per-sample radiance of one launch
    DevBuf<int> check_word;         // PTB_CHECK build: first bounds violation seen by a kernel (0 = none)
    float *preview_host = nullptr;  // pinned staging buffer of the progressive previews
    size_t preview_host_floats = 0;
    // multi-GPU context (ptb_create_multi): this context drives device_ids[0], `members` the other devices
    std::vector<ptb_ctx *> members;
    std::vector<float *> peer_stage;  // without peer access: copies of the members' sum framebuffers on this device
    size_t peer_stage_floats = 0;
    bool peer_access = false;
    DevBuf<int> scratch_i;
    DevBuf<int> tile_counter;
    DevBuf<unsigned long long> seg_counter;
    ptb_stats stats{};
    size_t max_smem_optin = 0;
    // an asynchronous ptb_render_device leaves its counters on the device until ptb_get_stats asks for them
    mutable bool stats_pending = false;
    mutable cudaStream_t pending_stream = nullptr;
};

namespace {

int fail(ptb_ctx *ctx, int code, const std::string &msg) {
    g_thread_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}
int cuda_fail(ptb_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, PTB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define CU(ctx, expr)                                             \
    do {                                                          \
        cudaError_t e__ = (expr);                                 \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr); \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host-side scene I/O
// ------------------------------------------------------------------------------------------------
extern "C" int ptb_scene_load_json(const char *json_path, const char *base_dir, ptb_scene **out, char *err, size_t errlen) {
    return ptb_scene_load_json_ex(json_path, base_dir, 0u, out, err, errlen);
}
extern "C" int ptb_scene_load_json_ex(const char *json_path, const char *base_dir, uint32_t flags, ptb_scene **out, char *err,
                                      size_t errlen) {
    if (err && errlen) err[0] = 0;
    if (!json_path || !out) return fail(nullptr, PTB_ERR_ARG, "ptb_scene_load_json: null argument");
    try {
        auto *s = new ptb_scene;
        s->host = load_scene_json(json_path, base_dir ? base_dir : "", flags);
        s->host.refresh_desc();
        *out = s;
        return PTB_OK;
    } catch (const SceneError &e) {
        if (err && errlen) std::snprintf(err, errlen, "%s", e.what());
        return fail(nullptr, e.code, e.what());
    } catch (const std::exception &e) {
        if (err && errlen) std::snprintf(err, errlen, "%s", e.what());
        return fail(nullptr, PTB_ERR_PARSE, e.what());
    }
}
extern "C" int ptb_scene_save_json(const ptb_scene *scene, const char *json_path) {
    if (!scene || !json_path) return fail(nullptr, PTB_ERR_ARG, "ptb_scene_save_json: null argument");
    try {
        save_scene_json(scene->host, json_path);
        return PTB_OK;
    } catch (const SceneError &e) {
        return fail(nullptr, e.code, e.what());
    }
}
extern "C" int ptb_scene_set_camera(ptb_scene *scene, const ptb_camera *camera) {
    if (!scene || !camera) return fail(nullptr, PTB_ERR_ARG, "ptb_scene_set_camera: null argument");
    scene->host.camera = *camera;
    scene->host.refresh_desc();
    return PTB_OK;
}
extern "C" const ptb_scene_desc *ptb_scene_get_desc(const ptb_scene *scene) { return scene ? &scene->host.desc : nullptr; }
extern "C" const char *ptb_scene_id(const ptb_scene *scene) { return scene ? scene->host.id.c_str() : ""; }
extern "C" void ptb_scene_free(ptb_scene *scene) { delete scene; }

extern "C" uint32_t ptb_to_int_with_gamma_correction(float x) { return to_int_with_gamma_correction(x); }
extern "C" int ptb_write_ppm(const char *path, const float *mean_rgb, int width, int height, uint64_t spp, const char *scene_id,
                             uint64_t seconds) {
    if (!path || !mean_rgb || width <= 0 || height <= 0) return fail(nullptr, PTB_ERR_ARG, "ptb_write_ppm: bad argument");
    try {
        write_ppm(path, mean_rgb, width, height, spp, scene_id ? scene_id : "", seconds);
        return PTB_OK;
    } catch (const SceneError &e) {
        return fail(nullptr, e.code, e.what());
    }
}
extern "C" uint64_t ptb_hash_pixels(const float *rgb, uint64_t n_pixels) { return rgb ? hash_pixels(rgb, n_pixels) : 0; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int ptb_abi_version(void) { return PTB_ABI_VERSION; }

extern "C" int ptb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *ptb_last_error(const ptb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_thread_error.c_str(); }

extern "C" void ptb_destroy(ptb_ctx *ctx) {
    if (!ctx) return;
    for (ptb_ctx *m : ctx->members) ptb_destroy(m);
    ctx->members.clear();
    cudaSetDevice(ctx->device);
    for (float *p : ctx->peer_stage) cudaFree(p);
    if (ctx->preview_host) cudaFreeHost(ctx->preview_host);
    ctx->preview.release(); ctx->check_word.release(); ctx->sample_L.release();
    ctx->loose_obj.release(); ctx->loose_tri.release(); ctx->obj_gate.release(); ctx->mat_color.release();
    ctx->mat_emis.release(); ctx->fb.release(); ctx->scratch_f.release(); ctx->scratch_i.release();
    ctx->tile_counter.release(); ctx->seg_counter.release();
    bvh_release(ctx->bvh);
    wf_release(ctx->wf);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int ptb_create(int device_id, ptb_ctx **out) {
    if (!out) return fail(nullptr, PTB_ERR_ARG, "ptb_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, PTB_ERR_CUDA, "ptb_create: no CUDA device (this backend has no CPU fallback)");
    }
    if (device_id < 0 || device_id >= n) return fail(nullptr, PTB_ERR_ARG, "ptb_create: device_id out of range");
    auto *ctx = new ptb_ctx;
    ctx->device = device_id;
    auto bail = [&](cudaError_t ce, const char *what) {
        int rc = cuda_fail(nullptr, ce, what);
        ptb_destroy(ctx);
        return rc;
    };
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        ptb_destroy(ctx);
        return fail(nullptr, PTB_ERR_CUDA, std::string("ptb_create: built for sm_100a only, device is ") + prop.name);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = ctx->tile_counter.resize(1)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = ctx->seg_counter.resize(4)) != cudaSuccess) return bail(e, "cudaMalloc");  // segments, BVH nodes, BVH prims, -
    if ((e = ctx->check_word.resize(1)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemsetAsync(ctx->check_word.p, 0, sizeof(int), ctx->stream)) != cudaSuccess) return bail(e, "cudaMemset");
    // parity self-test: the kernels must have been built with --fmad=false
    if ((e = ctx->scratch_f.resize(4)) != cudaSuccess) return bail(e, "cudaMalloc");
    const float a = 1.0f + 1.0f / 8192.0f, c = -(1.0f + 1.0f / 4096.0f);
    if ((e = launch_contraction_probe(a, a, c, ctx->scratch_f.p, ctx->stream)) != cudaSuccess) return bail(e, "probe launch");
    float probe[3] = {1.0f, 1.0f, 1.0f};
    if ((e = cudaMemcpyAsync(probe, ctx->scratch_f.p, 12, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bail(e, "probe copy");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "probe sync");
    if (probe[0] != 0.0f || probe[1] != 0.0f || probe[2] != 0.0f) {
        ptb_destroy(ctx);
        return fail(nullptr, PTB_ERR_STATE, "ptb_create: kernels were built with FMA contraction; rebuild with --fmad=false");
    }
    *out = ctx;
    return PTB_OK;
}

// ------------------------------------------------------------------------------------------------
// scene flatten (pure host code) + upload
// ------------------------------------------------------------------------------------------------
namespace {

const char *validate_desc(const ptb_scene_desc *desc, int &code) {  // nullptr = fine
    code = PTB_ERR_ARG;
    if (desc->n_objects && !desc->objects) return "objects is null";
    if (desc->n_objects > (1u << 28) || desc->n_triangles > (1u << 28)) { code = PTB_ERR_LIMIT; return "more than 2^28 objects or triangles"; }
    if (desc->n_triangles && !desc->triangles) return "triangles is null";
    for (size_t i = 0; i < desc->n_objects; ++i) {
        const ptb_object &o = desc->objects[i];
        if (o.kind != PTB_OBJ_SPHERE && o.kind != PTB_OBJ_MESH) return "bad object kind";
        if (o.reflect_type < 0 || o.reflect_type > 2) return "bad reflect_type";
        if (o.kind == PTB_OBJ_MESH && (o.tri_begin > desc->n_triangles || o.tri_count > desc->n_triangles - o.tri_begin))
            return "triangle range out of bounds";
    }
    return nullptr;
}

// prio = rank in the reference's scan order: objects last-to-first (mod.rs:637), triangles first-to-last (mod.rs:558)
bool scan_priorities(const ptb_scene_desc &desc, std::vector<uint32_t> &prio_base) {
    prio_base.assign(desc.n_objects, 0);
    uint64_t run = 0;
    for (size_t k = desc.n_objects; k-- > 0;) {
        prio_base[k] = static_cast<uint32_t>(run);
        run += desc.objects[k].kind == PTB_OBJ_SPHERE ? 1 : desc.objects[k].tri_count;
    }
    return run < 0xffffffffull;
}

float4 gate_sphere(const ptb_object &o) {
    if (o.kind != PTB_OBJ_MESH) return f4(0, 0, 0, 0);
    const V3 c = v3(o.bs_position) + v3(o.position);           // mod.rs:268
    return f4(c.x, c.y, c.z, o.bs_radius * o.bs_radius);        // radius.powi(2), mod.rs:416
}

double scene_diagonal(const ptb_scene_desc &desc) {  // diagonal of the box around every primitive
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    auto grow = [&](double x, double y, double z, double r) {
        const double p[3] = {x, y, z};
        for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], p[c] - r); hi[c] = std::max(hi[c], p[c] + r); }
    };
    for (size_t k = 0; k < desc.n_objects; ++k) {
        const ptb_object &o = desc.objects[k];
        if (o.kind == PTB_OBJ_SPHERE) grow(o.position[0], o.position[1], o.position[2], o.radius);
        else
            for (uint64_t j = 0; j < o.tri_count; ++j) {
                const ptb_triangle &t = desc.triangles[o.tri_begin + j];
                grow(t.a[0] + o.position[0], t.a[1] + o.position[1], t.a[2] + o.position[2], 0);
                grow(t.b[0] + o.position[0], t.b[1] + o.position[1], t.b[2] + o.position[2], 0);
                grow(t.c[0] + o.position[0], t.c[1] + o.position[1], t.c[2] + o.position[2], 0);
            }
    }
    if (hi[0] < lo[0]) return 0.0;
    return std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
}

struct LooseFlat {
    std::vector<float4> obj, tri;  // the object stream and the triangle records (layout: pt_device.cuh, DScene)
    uint32_t n_objects = 0, n_real_tris = 0;
};

// the objects with in_bvh[k] == 0, in the reference's scan order
void flatten_loose(const ptb_scene_desc &desc, const std::vector<char> &in_bvh, const std::vector<uint32_t> &prio_base,
                   double quad_min_ratio, LooseFlat &out) {
    const double scene_diag = scene_diagonal(desc);
    std::vector<float4> &lobj = out.obj, &ltri = out.tri;
    for (size_t k = desc.n_objects; k-- > 0;) {  // reverse index order = the reference's scan order
        const ptb_object &o = desc.objects[k];
        if (in_bvh[k]) continue;
        // A mesh without triangles can never be hit (mod.rs:558 loops zero times).  It must not enter the stream: a triangle
        // count of 0 in a mesh header is the marker of a vote-free wall quad (closest_hit_loose), which would then read the
        // following records as a pair and lose its place in the stream.
        if (o.kind == PTB_OBJ_MESH && o.tri_count == 0) continue;
        out.n_objects++;
        if (o.kind == PTB_OBJ_SPHERE) {
            const int32_t self = static_cast<int32_t>(lobj.size());
            lobj.push_back(f4(o.position[0], o.position[1], o.position[2], o.radius * o.radius));
            lobj.push_back(f4(ibits(KIND_SPHERE), ubits(prio_base[k]), ibits(self), ibits(static_cast<int32_t>(k))));
        } else {
            // gate shortcut (sphere_gate): usable for 0.4 <= r <= 1000, see the derivation there
            const float r2_inside = (o.bs_radius >= 0.4f && o.bs_radius <= 1000.0f) ? (o.bs_radius * o.bs_radius) * 0.999f : -1.0f;
            const uint64_t n_padded = (o.tri_count + 1) & ~1ull;
            const int32_t k_begin = static_cast<int32_t>(ltri.size() / 2);
            lobj.push_back(gate_sphere(o));
            // a one-pair mesh whose gate sphere is large against the scene skips the per-mesh warp vote (closest_hit_loose)
            const bool always = n_padded == 2 && static_cast<double>(o.bs_radius) >= quad_min_ratio * scene_diag;
            lobj.push_back(f4(r2_inside, ibits(k_begin), ibits(always ? 0 : static_cast<int32_t>(n_padded)),
                              ibits(static_cast<int32_t>(2 + 5 * (n_padded / 2)))));
            const V3 off = v3(o.position);
            // per triangle: A' | E1 | E2 | prio for the pair records, unit normal | ids for the triangle records
            std::vector<V3> ta(n_padded, mk3(0, 0, 0)), te1(n_padded, mk3(0, 0, 0)), te2(n_padded, mk3(0, 0, 0));
            std::vector<uint32_t> tp(n_padded, 0xffffffffu);  // (odd count: a null triangle pads the pair, det = 0, always rejected, mod.rs:571)
            for (uint64_t j = 0; j < n_padded; ++j) {
                if (j < o.tri_count) {
                    const ptb_triangle &t = desc.triangles[o.tri_begin + j];
                    // Triangle::transformed then the edge vectors, exactly as per ray in mod.rs:559-561
                    const V3 a = v3(t.a) + off, b = v3(t.b) + off, c = v3(t.c) + off;
                    ta[j] = a; te1[j] = b - a; te2[j] = c - a;
                    tp[j] = prio_base[k] + static_cast<uint32_t>(j);
                    out.n_real_tris++;
                }
                const V3 nrm = normalize(cross(te1[j], te2[j]));  // mod.rs:605, the same fp32 operations the reference does per hit
                ltri.push_back(f4(nrm.x, nrm.y, nrm.z, ibits(static_cast<int32_t>(k))));
                ltri.push_back(f4(ibits(j < o.tri_count ? static_cast<int32_t>(j) : -1), 0, 0, 0));
            }
            for (uint64_t j = 0; j < n_padded; j += 2) {  // pair records for the packed tests (pt_device.cuh: triangle_pair_hit)
                const V3 &a0 = ta[j], &a1 = ta[j + 1], &p0 = te1[j], &p1 = te1[j + 1], &q0 = te2[j], &q1 = te2[j + 1];
                lobj.push_back(f4(a0.x, a1.x, a0.y, a1.y));
                lobj.push_back(f4(a0.z, a1.z, p0.x, p1.x));
                lobj.push_back(f4(p0.y, p1.y, p0.z, p1.z));
                lobj.push_back(f4(q0.x, q1.x, q0.y, q1.y));
                lobj.push_back(f4(q0.z, q1.z, ubits(tp[j]), ubits(tp[j + 1])));
            }
        }
    }
    lobj.push_back(f4(0, 0, 0, 0));
    lobj.push_back(f4(ibits(KIND_END), 0, 0, 0));
}

}  // namespace

// Diagnostic, host only (no GPU, no context): the shared-memory scene stream and triangle records exactly as ptb_upload_scene
// builds them when every object stays in the lock-step list.  Lets the layout be checked on a machine without a GPU.
extern "C" int ptb_flatten_loose(const ptb_scene_desc *desc, double quad_min_ratio, float *stream, uint64_t stream_cap_floats,
                                 uint64_t *stream_floats, float *tris, uint64_t tris_cap_floats, uint64_t *tris_floats) {
    if (!desc || !stream_floats || !tris_floats) return fail(nullptr, PTB_ERR_ARG, "ptb_flatten_loose: null argument");
    int code = PTB_OK;
    if (const char *why = validate_desc(desc, code)) return fail(nullptr, code, std::string("ptb_flatten_loose: ") + why);
    std::vector<uint32_t> prio_base;
    if (!scan_priorities(*desc, prio_base)) return fail(nullptr, PTB_ERR_LIMIT, "ptb_flatten_loose: too many primitives");
    LooseFlat flat;
    flatten_loose(*desc, std::vector<char>(desc->n_objects, 0), prio_base, quad_min_ratio, flat);
    *stream_floats = 4 * flat.obj.size();
    *tris_floats = 4 * flat.tri.size();
    if (stream && stream_cap_floats >= *stream_floats && !flat.obj.empty()) std::memcpy(stream, flat.obj.data(), *stream_floats * sizeof(float));
    if (tris && tris_cap_floats >= *tris_floats && !flat.tri.empty()) std::memcpy(tris, flat.tri.data(), *tris_floats * sizeof(float));
    return PTB_OK;
}

static int upload_scene_one(ptb_ctx *ctx, const ptb_scene_desc *desc) {
    int code = PTB_OK;
    if (const char *why = validate_desc(desc, code)) return fail(ctx, code, std::string("ptb_upload_scene: ") + why);
    CU(ctx, cudaSetDevice(ctx->device));
    // an asynchronous ptb_render_device on a caller's stream may still be reading the scene buffers this call rewrites
    if (ctx->render_in_flight) { CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_done, 0)); CU(ctx, cudaEventSynchronize(ctx->ev_done)); ctx->render_in_flight = false; }
    const double t0 = now_ms();
    const size_t nobj = desc->n_objects;

    std::vector<uint32_t> prio_base;
    if (!scan_priorities(*desc, prio_base)) return fail(ctx, PTB_ERR_LIMIT, "ptb_upload_scene: too many primitives");

    std::vector<float4> gate(nobj), mcol(nobj), memi(nobj);
    for (size_t i = 0; i < nobj; ++i) {
        const ptb_object &o = desc->objects[i];
        gate[i] = gate_sphere(o);
        mcol[i] = f4(o.color[0], o.color[1], o.color[2], ibits(o.reflect_type));
        const bool emits = o.emission[0] != 0.0f || o.emission[1] != 0.0f || o.emission[2] != 0.0f;
        memi[i] = f4(o.emission[0], o.emission[1], o.emission[2], ibits(emits ? 1 : 0));
    }

    // split: which objects go to the BVH, which stay in the lock-step shared-memory list
    std::vector<char> in_bvh(nobj, 0);
    choose_bvh_objects(*desc, ctx->max_smem_optin, ctx->bvh_opt, in_bvh);

    LooseFlat flat;
    flatten_loose(*desc, in_bvh, prio_base, ctx->quad_min_ratio, flat);
    const std::vector<float4> &lobj = flat.obj, &ltri = flat.tri;
    const uint32_t n_real_loose_tris = flat.n_real_tris, n_loose_objects = flat.n_objects;
    if (lobj.size() >= (1u << 28)) return fail(ctx, PTB_ERR_LIMIT, "ptb_upload_scene: loose primitive list too long");
    const size_t loose_bytes = (lobj.size() + ltri.size()) * sizeof(float4);
    if (loose_bytes > ctx->max_smem_optin)
        return fail(ctx, PTB_ERR_LIMIT, "ptb_upload_scene: loose primitive list does not fit in shared memory");

    CU(ctx, ctx->loose_obj.upload(lobj, ctx->stream));
    CU(ctx, ctx->loose_tri.upload(ltri, ctx->stream));
    CU(ctx, ctx->obj_gate.upload(gate, ctx->stream));
    CU(ctx, ctx->mat_color.upload(mcol, ctx->stream));
    CU(ctx, ctx->mat_emis.upload(memi, ctx->stream));

    DScene &ds = ctx->ds;
    ds = DScene{};
    ds.loose_obj = ctx->loose_obj.p; ds.loose_tri = ctx->loose_tri.p;
    ds.n_loose_f4 = static_cast<int>(lobj.size()); ds.n_loose_obj = static_cast<int>(n_loose_objects);
    ds.n_loose_tri = static_cast<int>(ltri.size() / 2);
    ds.obj_gate = ctx->obj_gate.p; ds.mat_color = ctx->mat_color.p; ds.mat_emis = ctx->mat_emis.p;
    ds.n_obj = static_cast<int>(nobj);
    ds.bvh_root = BVH_EMPTY_REF;
    ds.check = ctx->check_word.p;

    // camera frame, once per scene like render() (mod.rs:998-999; CameraData mod.rs:211-232)
    {
        const ptb_camera &c = desc->camera;
        const V3 dir = v3(c.direction), pos = v3(c.position);
        ds.sensor_origin = pos;
        ds.lens_center = pos + dir * c.focal_length;
        const V3 su = normalize(cross(dir, std::fabs(dir.y) < 0.9f ? mk3(0.f, 1.f, 0.f) : mk3(0.f, 0.f, 1.f)));
        const V3 sv = cross(su, dir);
        const float sensor_height = c.sensor_width / c.aspect_ratio;
        ds.su = su * c.sensor_width;
        ds.sv = sv * sensor_height;
    }

    const double t1 = now_ms();
    double bvh_ms = 0.0;
    ctx->stats.n_bvh_triangles = ctx->stats.n_bvh_spheres = ctx->stats.n_bvh_nodes = 0;
    {
        std::string berr;
        cudaError_t e = bvh_build(*desc, in_bvh, prio_base, ctx->bvh_opt, ctx->bvh, ds, ctx->stream, &bvh_ms, berr);
        if (e != cudaSuccess) return cuda_fail(ctx, e, berr.empty() ? "bvh_build" : berr.c_str());
        ctx->stats.n_bvh_triangles = ctx->bvh.n_tris;
        ctx->stats.n_bvh_spheres = ctx->bvh.n_spheres;
        ctx->stats.n_bvh_nodes = ctx->bvh.n_nodes;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.upload_ms = (t1 - t0) + (now_ms() - t1 - bvh_ms > 0 ? now_ms() - t1 - bvh_ms : 0.0);
    ctx->stats.bvh_build_ms = bvh_ms;
    ctx->stats.n_loose_objects = static_cast<uint32_t>(ds.n_loose_obj);
    ctx->stats.n_loose_triangles = n_real_loose_tris;
    ctx->has_scene = true;
    return PTB_OK;
}

// runs fn(member) for every device of a (multi-GPU) context, one host thread per device; returns the first error
template <typename F>
static int for_each_device(ptb_ctx *ctx, F fn) {
    if (ctx->members.empty()) return fn(ctx, 0);
    const size_t G = ctx->members.size() + 1;
    std::vector<int> rc(G, PTB_OK);
    std::vector<std::thread> th;
    for (size_t g = 1; g < G; ++g) th.emplace_back([&, g] { rc[g] = fn(ctx->members[g - 1], (int)g); });
    rc[0] = fn(ctx, 0);
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; ++g)
        if (rc[g] < 0) {
            if (g > 0) ctx->err = "device " + std::to_string(ctx->members[g - 1]->device) + ": " + ctx->members[g - 1]->err;
            g_thread_error = ctx->err;
            return rc[g];
        }
    return PTB_OK;
}

extern "C" int ptb_upload_scene(ptb_ctx *ctx, const ptb_scene_desc *desc) {
    if (!ctx || !desc) return fail(ctx, PTB_ERR_ARG, "ptb_upload_scene: null argument");
    return for_each_device(ctx, [&](ptb_ctx *m, int) { return upload_scene_one(m, desc); });  // the scene is replicated
}

extern "C" int ptb_selftest(ptb_ctx *ctx, uint64_t *mismatches) {
    if (!ctx || !mismatches) return fail(ctx, PTB_ERR_ARG, "ptb_selftest: null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(ctx->seg_counter.p, 0, sizeof(unsigned long long), ctx->stream));
    CU(ctx, launch_rcp_selftest(ctx->seg_counter.p, ctx->sm_count, ctx->stream));
    unsigned long long bad = 0;
    CU(ctx, cudaMemcpyAsync(&bad, ctx->seg_counter.p, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *mismatches = bad;
    return PTB_OK;
}

extern "C" int ptb_set_option(ptb_ctx *ctx, const char *key, double value) {
    if (!ctx || !key) return fail(ctx, PTB_ERR_ARG, "ptb_set_option: null argument");
    const std::string k = key;
    for (ptb_ctx *m : ctx->members) {
        const int rc = ptb_set_option(m, key, value);
        if (rc != PTB_OK) return fail(ctx, rc, m->err);
    }
    if (k == "bvh_min_tris") ctx->bvh_opt.min_tris = value;
    else if (k == "bvh_min_spheres") ctx->bvh_opt.min_spheres = value;
    else if (k == "bvh_leaf_max") ctx->bvh_opt.leaf_max = ctx->bvh_opt.leaf_max_small = (int)value;
    else if (k == "bvh_sah_max_prims") ctx->bvh_opt.sah_max_prims = (int)std::max(0.0, std::min(1e9, value));
    else if (k == "bvh_wide") ctx->bvh_opt.wide = (int)value;
    else if (k == "bvh_wide_sah") ctx->bvh_opt.wide_sah = (int)value;
    else if (k == "bvh_top_levels") ctx->bvh_opt.top_levels = std::max(0, std::min(5, (int)value));
    else if (k == "wf_refill") ctx->wf_opt.refill = ctx->wf_opt.refill_wide = (int)value;
    else if (k == "wf_descend_min") ctx->wf_opt.descend_min = ctx->wf_opt.descend_min_wide = (int)value;
    else if (k == "wf_trace_threads") ctx->wf_opt.trace_threads = ctx->wf_opt.trace_threads_wide = (int)value;
    else if (k == "wf_top8_nodes") ctx->wf_opt.top8_nodes = (int)value;
    else if (k == "quad_min_ratio") ctx->quad_min_ratio = value;
    else if (k == "regen_batch") ctx->regen_batch = std::max(1, std::min(32, (int)value));
    else if (k == "integrator") ctx->integrator = (int)value;
    else if (k == "wavefront_paths") ctx->wf_opt.target_paths = (size_t)std::max(1024.0, value);
#ifdef PTB_EXPERIMENTS  // never in the release library: a value below 1 voids the parity guarantee (tools/pad_check.py)
    else if (k == "bvh_pad_scale_UNSAFE") ctx->bvh_opt.pad_scale = value;
#endif
    else return fail(ctx, PTB_ERR_ARG, "ptb_set_option: unknown key " + k);
    return PTB_OK;
}

static int finish_pending_stats(const ptb_ctx *ctx) {
    if (!ctx->stats_pending) return PTB_OK;
    ptb_ctx *m = const_cast<ptb_ctx *>(ctx);
    CU(m, cudaSetDevice(ctx->device));
    CU(m, cudaStreamSynchronize(ctx->pending_stream));
    float ms = 0.f;
    CU(m, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    unsigned long long cnt[4] = {0, 0, 0, 0};
    CU(m, cudaMemcpy(cnt, ctx->seg_counter.p, sizeof cnt, cudaMemcpyDeviceToHost));
    m->stats.render_ms = ms;
    m->stats.segments = cnt[0];
    m->stats.bvh_nodes_visited = cnt[1];
    m->stats.bvh_prims_tested = cnt[2];
    ctx->stats_pending = false;
    return PTB_OK;
}

extern "C" int ptb_get_stats(const ptb_ctx *ctx, ptb_stats *out) {
    if (!ctx || !out) return PTB_ERR_ARG;
    int rc = finish_pending_stats(ctx);
    if (rc != PTB_OK) return rc;
    *out = ctx->stats;
    // a multi-GPU context reports the whole job: work summed over its devices, render_ms of the slowest one
    for (const ptb_ctx *m : ctx->members) {
        if ((rc = finish_pending_stats(m)) != PTB_OK) return rc;
        out->segments += m->stats.segments;
        out->samples += m->stats.samples;
        out->kernel_launches += m->stats.kernel_launches;
        out->bvh_nodes_visited += m->stats.bvh_nodes_visited;
        out->bvh_prims_tested += m->stats.bvh_prims_tested;
        out->render_ms = std::max(out->render_ms, m->stats.render_ms);
    }
    // (every entry point leaves the calling thread on the context's first device, also after it has visited the members)
    if (!ctx->members.empty()) cudaSetDevice(ctx->device);
    return PTB_OK;
}

// ------------------------------------------------------------------------------------------------
// render
// ------------------------------------------------------------------------------------------------
namespace {

struct Preview {  // progressive previews (RenderUpdate.image, mod.rs:965-982): resolved partial mean handed to a callback
    ptb_preview_fn fn = nullptr;
    void *user = nullptr;
    double interval_ms = 500.0;
};

#ifdef PTB_CHECK
// turns a bounds violation recorded by a kernel into an error (the stream must be idle)
int check_word_result(ptb_ctx *ctx) {
    int word = 0;
    CU(ctx, cudaMemcpy(&word, ctx->check_word.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (word != 0) {
        cudaMemset(ctx->check_word.p, 0, sizeof(int));
        return fail(ctx, PTB_ERR_STATE, "PTB_CHECK: a kernel saw an out-of-bounds index, code " + std::to_string(word) +
                                            " (1 traversal stack, 2 queue append, 3 slot / branch mask, 5 ray index, 7 node, 8 primitive)");
    }
    return PTB_OK;
}
#endif

// fresh_frame: the buffer's content is undefined and counts as zero; the first batch overwrites it instead of accumulating
int render_device_impl(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed, float *d_sum_rgb,
                       void *cuda_stream, const volatile int32_t *cancel, volatile uint64_t *samples_done, bool fresh_frame,
                       const Preview *preview = nullptr) {
    if (!ctx) return fail(nullptr, PTB_ERR_ARG, "ptb_render_device: ctx is null");
    if (!ctx->has_scene) return fail(ctx, PTB_ERR_STATE, "ptb_render_device: no scene uploaded");
    if (width <= 0 || height <= 0 || !d_sum_rgb) return fail(ctx, PTB_ERR_ARG, "ptb_render_device: bad argument");
    if (static_cast<uint64_t>(width) * static_cast<uint64_t>(height) > (1ull << 31) - 1)
        return fail(ctx, PTB_ERR_LIMIT, "ptb_render_device: more than 2^31-1 pixels");
    if (spp_begin + spp_count < spp_begin) return fail(ctx, PTB_ERR_ARG, "ptb_render_device: sample range overflows");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);  // NULL = the legacy default stream, as in CUDA
    const bool with_preview = preview && preview->fn;
    const bool interactive = cancel != nullptr || samples_done != nullptr || with_preview;

    RenderArgs a{};
    a.width = width; a.height = height; a.seed = seed; a.sum_rgb = d_sum_rgb;
    philox_round_keys(seed, a.rk);
    a.tile_counter = ctx->tile_counter.p; a.segment_counter = ctx->seg_counter.p;
    a.regen_batch = ctx->regen_batch;
    a.tiles_x = (width + TILE_W - 1) / TILE_W;
    a.n_tiles = a.tiles_x * ((height + TILE_H - 1) / TILE_H);

    ctx->stats.kernel_launches = 0;
    ctx->stats.samples = 0;
    // the kernels use per-context state (queues, counters, workspace): order this render after the previous one even when the
    // caller switched streams in between
    if (ctx->render_in_flight) CU(ctx, cudaStreamWaitEvent(st, ctx->ev_done, 0));
    CU(ctx, cudaMemsetAsync(ctx->seg_counter.p, 0, 4 * sizeof(unsigned long long), st));
    CU(ctx, cudaEventRecord(ctx->ev0, st));
    const uint64_t npix = static_cast<uint64_t>(width) * static_cast<uint64_t>(height);
    // auto: the wavefront integrator when the scene has a BVH; the megakernel otherwise -- one lane per pixel when the frame gives
    // every resident warp at least two pixel tiles, sample-parallel (a lane takes a chunk of a pixel's samples) for smaller frames
    // such as the reference's default 450x300 (mod.rs:872-879)
    const bool has_bvh = ctx->ds.bvh_root != BVH_EMPTY_REF;
    const bool small_image = a.n_tiles < 2 * ctx->sm_count * (RENDER_MIN_BLOCKS * RENDER_THREADS / 32);
    const bool wavefront = ctx->integrator == 2 || (ctx->integrator == 0 && has_bvh);
    const bool sample_parallel = !wavefront && (ctx->integrator == 3 || (ctx->integrator == 0 && small_image));
    // Work per launch.  The cancel flag is polled and progress published between launches only (the reference polls per pixel,
    // mod.rs:1003), so a caller that watches either gets launches of roughly 0.1-0.2 s: 2^28 samples on shared-memory scenes
    // (~2 Gpaths/s), 2^25 through a BVH; otherwise 2^31 samples per launch keep the launch count low.
    const uint64_t per_launch = !interactive ? (1ull << 31) : (has_bvh ? (1ull << 25) : (1ull << 28));
    uint64_t batch = std::max<uint64_t>(1, per_launch / npix);
    if (wavefront)  // whole wavefront batches: a launch smaller than the paths in flight would leave the queues short
        batch = std::max<uint64_t>(batch, std::max<uint64_t>(1, ctx->wf_opt.target_paths / npix));
    if (sample_parallel) {  // the per-sample radiance buffer bounds a launch: 2^25 samples (512 MB)
        batch = std::max<uint64_t>(1, std::min<uint64_t>(batch, (1ull << 25) / npix));
        CU(ctx, ctx->sample_L.resize(std::min<uint64_t>(batch, spp_count) * npix));
    }
    uint64_t done = 0;
    int rc = PTB_OK;
    double last_preview = now_ms();
    while (done < spp_count) {
        if (cancel && *cancel) { rc = PTB_CANCELLED; break; }
        const uint64_t n = std::min(batch, spp_count - done);
        a.spp_begin = spp_begin + done;
        a.spp_count = n;
        a.fb_zero = (fresh_frame && done == 0) ? 1 : 0;
        if (wavefront) {
            CU(ctx, wavefront_render(ctx->ds, a, ctx->wf, ctx->sm_count, ctx->wf_opt, st, &ctx->stats.kernel_launches));
        } else {
            a.sample_L = nullptr;
            if (sample_parallel) {
                // chunks of 8 samples, fewer when that would leave less than ~8 slots per resident lane
                const uint64_t lanes = (uint64_t)ctx->sm_count * RENDER_MIN_BLOCKS * RENDER_THREADS;
                uint64_t chunk = 8;
                while (chunk > 1 && npix * ((n + chunk - 1) / chunk) < 8 * lanes) chunk /= 2;
                a.sample_L = ctx->sample_L.p;
                a.sp_chunk = (int)chunk;
                a.sp_n_chunks = (int)((n + chunk - 1) / chunk);
                if ((uint64_t)a.n_tiles * 32ull * (uint64_t)a.sp_n_chunks >= (1ull << 31))
                    return fail(ctx, PTB_ERR_LIMIT, "ptb_render_device: too many sample chunks for one launch");
            }
            CU(ctx, cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(int), st));
            CU(ctx, launch_render(ctx->ds, a, ctx->sm_count, st));
            ctx->stats.kernel_launches += sample_parallel ? 2 : 1;
        }
        done += n;
        if (interactive) {
            CU(ctx, cudaStreamSynchronize(st));
            if (samples_done) *samples_done = done * npix;
            if (with_preview && done < spp_count && now_ms() - last_preview >= preview->interval_ms) {
                // RenderUpdate.image: the mean of the samples finished so far, clamped like the final image (mod.rs:849-856)
                const size_t nfl = npix * 3;
                CU(ctx, ctx->preview.resize(nfl));
                if (ctx->preview_host_floats < nfl) {
                    if (ctx->preview_host) cudaFreeHost(ctx->preview_host);
                    ctx->preview_host = nullptr; ctx->preview_host_floats = 0;
                    CU(ctx, cudaMallocHost(reinterpret_cast<void **>(&ctx->preview_host), nfl * sizeof(float)));
                    ctx->preview_host_floats = nfl;
                }
                CU(ctx, launch_resolve(d_sum_rgb, nfl, done, ctx->preview.p, ctx->sm_count, st));
                CU(ctx, cudaMemcpyAsync(ctx->preview_host, ctx->preview.p, nfl * sizeof(float), cudaMemcpyDeviceToHost, st));
                CU(ctx, cudaStreamSynchronize(st));
                ctx->stats.kernel_launches++;
                preview->fn(preview->user, ctx->preview_host, width, height, done, spp_count);
                last_preview = now_ms();
            }
        }
    }
    if (fresh_frame && done == 0)  // nothing was rendered (no samples asked for, or cancelled at once): the frame is black
        CU(ctx, cudaMemsetAsync(d_sum_rgb, 0, npix * 3 * sizeof(float), st));
    CU(ctx, cudaEventRecord(ctx->ev1, st));
    CU(ctx, cudaEventRecord(ctx->ev_done, st));
    ctx->render_in_flight = true;
    ctx->stats.samples = done * npix;
    ctx->stats_pending = true;
    ctx->pending_stream = st;
    if (interactive) {
        const int frc = finish_pending_stats(ctx);
        if (frc != PTB_OK) return frc;
    }
#ifdef PTB_CHECK
    CU(ctx, cudaStreamSynchronize(st));
    if (const int crc = check_word_result(ctx)) return crc;
#endif
    return rc;
}

// one frame on one context: render into ctx->fb, resolve, copy to the host
int render_one(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed, int out_kind, float *out_rgb,
               const volatile int32_t *cancel, volatile uint64_t *samples_done, const Preview *preview) {
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t nfl = static_cast<size_t>(width) * static_cast<size_t>(height) * 3;
    CU(ctx, ctx->fb.resize(nfl));
    // (no progress pointer is invented here: a caller that passes neither cancel nor samples_done gets no per-launch sync)
    int rc = render_device_impl(ctx, width, height, spp_begin, spp_count, seed, ctx->fb.p, ctx->stream, cancel, samples_done,
                                /*fresh_frame=*/true, preview);
    if (rc < 0) return rc;
    const uint64_t spp_done = ctx->stats.samples / (static_cast<uint64_t>(width) * static_cast<uint64_t>(height));
    if (out_kind == PTB_OUT_MEAN && spp_done > 0) {
        CU(ctx, launch_resolve(ctx->fb.p, nfl, rc == PTB_CANCELLED ? spp_done : spp_count, ctx->fb.p, ctx->sm_count, ctx->stream));
        ctx->stats.kernel_launches++;
    }
    CU(ctx, cudaMemcpyAsync(out_rgb, ctx->fb.p, nfl * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// One frame on a multi-GPU context: device g renders the global sample indices [g*spp/G, (g+1)*spp/G) of every pixel
// into its own unclamped sum framebuffer, one host thread per device; then every device reduces ITS slice of the image straight
// out of the peers' framebuffers over NVLink (k_peer_reduce_resolve: fixed rank order ((fb0+fb1)+fb2)+..., so the image is
// deterministic), resolves it and stores it into device 0's buffer, which is copied to the host.
int render_multi(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed, int out_kind, float *out_rgb,
                 const volatile int32_t *cancel, volatile uint64_t *samples_done, const Preview *preview) {
    const size_t G = ctx->members.size() + 1;
    const uint64_t npix = static_cast<uint64_t>(width) * static_cast<uint64_t>(height);
    const size_t nfl = npix * 3;
    std::vector<ptb_ctx *> dev(G);
    dev[0] = ctx;
    for (size_t g = 1; g < G; ++g) dev[g] = ctx->members[g - 1];
    for (ptb_ctx *m : dev)
        if (!m->has_scene) return fail(ctx, PTB_ERR_STATE, "ptb_render: no scene uploaded");
    std::unique_ptr<volatile uint64_t[]> progress(new volatile uint64_t[G]);
    for (size_t g = 0; g < G; ++g) progress[g] = 0;
    std::vector<int> rc(G, PTB_OK);
    std::atomic<size_t> running{G};
    const bool watch = cancel != nullptr || samples_done != nullptr || (preview && preview->fn);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            ptb_ctx *m = dev[g];
            const uint64_t b = spp_count * g / G, e = spp_count * (g + 1) / G;  // the split of distributed.py: shard_samples
            cudaError_t ce = cudaSetDevice(m->device);
            if (ce == cudaSuccess) ce = m->fb.resize(nfl);
            if (ce != cudaSuccess) rc[g] = cuda_fail(m, ce, "ptb_render (multi): framebuffer");
            else {
                // previews come from device 0's share alone: an unbiased partial mean of the same image
                rc[g] = render_device_impl(m, width, height, spp_begin + b, e - b, seed, m->fb.p, m->stream, cancel,
                                           watch ? &progress[g] : nullptr, /*fresh_frame=*/true, g == 0 ? preview : nullptr);
                if (rc[g] >= 0 && cudaStreamSynchronize(m->stream) != cudaSuccess) rc[g] = fail(m, PTB_ERR_CUDA, "ptb_render (multi): sync");
            }
            running.fetch_sub(1);
        });
    while (running.load() > 0) {  // the calling thread publishes the job's progress (processed_pixel_count analogue, mod.rs:850)
        if (samples_done) {
            uint64_t sum = 0;
            for (size_t g = 0; g < G; ++g) sum += progress[g];
            *samples_done = sum;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(watch ? 5 : 1));
    }
    for (auto &t : th) t.join();
    int result = PTB_OK;
    uint64_t samples = 0;
    for (size_t g = 0; g < G; ++g) {
        if (rc[g] < 0) {
            ctx->err = "device " + std::to_string(dev[g]->device) + ": " + dev[g]->err;
            g_thread_error = ctx->err;
            return rc[g];
        }
        if (rc[g] == PTB_CANCELLED) result = PTB_CANCELLED;
        samples += dev[g]->stats.samples;
    }
    if (samples_done) *samples_done = samples;
    const uint64_t spp_done = samples / npix;  // (a cancelled frame: every device stopped after whole launches)
    const bool resolve = out_kind == PTB_OUT_MEAN && spp_done > 0;
    const uint64_t divisor = resolve ? (result == PTB_CANCELLED ? spp_done : spp_count) : 0;  // 0 = raw sum
    PeerPtrs pp{};
    if (ctx->peer_access) {
        for (size_t g = 0; g < G; ++g) pp.p[g] = dev[g]->fb.p;
        for (size_t g = 0; g < G; ++g) {  // slice g of the image is reduced by device g, written into device 0's buffer (P2P store)
            const uint64_t f0 = (nfl * g / G) & ~3ull, f1 = g + 1 == G ? nfl : ((nfl * (g + 1) / G) & ~3ull);
            CU(ctx, cudaSetDevice(dev[g]->device));
            CU(ctx, launch_peer_reduce_resolve(pp, (int)G, f0, f1 - f0, divisor, ctx->fb.p, dev[g]->sm_count, dev[g]->stream));
            dev[g]->stats.kernel_launches++;
        }
        for (size_t g = 0; g < G; ++g) {
            CU(ctx, cudaSetDevice(dev[g]->device));
            CU(ctx, cudaStreamSynchronize(dev[g]->stream));
        }
        CU(ctx, cudaSetDevice(ctx->device));
    } else {  // no peer access between these devices: stage the members' framebuffers on device 0 and reduce there
        CU(ctx, cudaSetDevice(ctx->device));
        if (ctx->peer_stage.size() < G - 1 || ctx->peer_stage_floats < nfl) {
            for (float *p : ctx->peer_stage) cudaFree(p);
            ctx->peer_stage.assign(G - 1, nullptr);
            ctx->peer_stage_floats = 0;
            for (size_t g = 1; g < G; ++g) CU(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->peer_stage[g - 1]), nfl * sizeof(float)));
            ctx->peer_stage_floats = nfl;
        }
        pp.p[0] = ctx->fb.p;
        for (size_t g = 1; g < G; ++g) {
            CU(ctx, cudaMemcpyPeerAsync(ctx->peer_stage[g - 1], ctx->device, dev[g]->fb.p, dev[g]->device, nfl * sizeof(float), ctx->stream));
            pp.p[g] = ctx->peer_stage[g - 1];
        }
        CU(ctx, launch_peer_reduce_resolve(pp, (int)G, 0, nfl, divisor, ctx->fb.p, ctx->sm_count, ctx->stream));
        ctx->stats.kernel_launches++;
    }
    CU(ctx, cudaMemcpyAsync(out_rgb, ctx->fb.p, nfl * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return result;
}

int render_any(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed, int out_kind, float *out_rgb,
               const volatile int32_t *cancel, volatile uint64_t *samples_done, const Preview *preview, const char *who) {
    if (!ctx) return fail(nullptr, PTB_ERR_ARG, std::string(who) + ": ctx is null");
    if (!out_rgb || width <= 0 || height <= 0) return fail(ctx, PTB_ERR_ARG, std::string(who) + ": bad argument");
    if (out_kind != PTB_OUT_MEAN && out_kind != PTB_OUT_SUM) return fail(ctx, PTB_ERR_ARG, std::string(who) + ": bad out_kind");
    if (out_kind == PTB_OUT_MEAN && spp_count == 0) return fail(ctx, PTB_ERR_ARG, std::string(who) + ": spp_count is 0");
    if (!ctx->has_scene) return fail(ctx, PTB_ERR_STATE, std::string(who) + ": no scene uploaded");
    if (spp_begin + spp_count < spp_begin) return fail(ctx, PTB_ERR_ARG, std::string(who) + ": sample range overflows");
    if (ctx->members.empty()) return render_one(ctx, width, height, spp_begin, spp_count, seed, out_kind, out_rgb, cancel, samples_done, preview);
    return render_multi(ctx, width, height, spp_begin, spp_count, seed, out_kind, out_rgb, cancel, samples_done, preview);
}

}  // namespace

extern "C" int ptb_render_device(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed,
                                 float *d_sum_rgb, void *cuda_stream, const volatile int32_t *cancel,
                                 volatile uint64_t *samples_done) {
    if (ctx && !ctx->members.empty())
        return fail(ctx, PTB_ERR_STATE, "ptb_render_device: a multi-GPU context renders through ptb_render / ptb_render_progressive");
    return render_device_impl(ctx, width, height, spp_begin, spp_count, seed, d_sum_rgb, cuda_stream, cancel, samples_done, false);
}

extern "C" int ptb_resolve_device(ptb_ctx *ctx, const float *d_sum_rgb, uint64_t n_floats, uint64_t spp_total, float *d_mean_rgb,
                                  void *cuda_stream) {
    if (!ctx || !d_sum_rgb || !d_mean_rgb || spp_total == 0) return fail(ctx, PTB_ERR_ARG, "ptb_resolve_device: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);  // NULL = the legacy default stream, as in CUDA
    CU(ctx, launch_resolve(d_sum_rgb, n_floats, spp_total, d_mean_rgb, ctx->sm_count, st));
    return PTB_OK;
}

extern "C" int ptb_render(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed, int out_kind,
                          float *out_rgb, const volatile int32_t *cancel, volatile uint64_t *samples_done) {
    return render_any(ctx, width, height, spp_begin, spp_count, seed, out_kind, out_rgb, cancel, samples_done, nullptr, "ptb_render");
}

extern "C" int ptb_render_progressive(ptb_ctx *ctx, int width, int height, uint64_t spp_begin, uint64_t spp_count, uint64_t seed,
                                      int out_kind, float *out_rgb, const volatile int32_t *cancel, volatile uint64_t *samples_done,
                                      double preview_interval_ms, ptb_preview_fn on_preview, void *user) {
    Preview pv;
    pv.fn = on_preview; pv.user = user; pv.interval_ms = preview_interval_ms > 0.0 ? preview_interval_ms : 0.0;
    return render_any(ctx, width, height, spp_begin, spp_count, seed, out_kind, out_rgb, cancel, samples_done, &pv,
                      "ptb_render_progressive");
}

extern "C" int ptb_create_multi(const int *device_ids, int n_devices, ptb_ctx **out) {
    if (!out) return fail(nullptr, PTB_ERR_ARG, "ptb_create_multi: out is null");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > MAX_PEERS) return fail(nullptr, PTB_ERR_ARG, "ptb_create_multi: 1..16 devices");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (device_ids[i] == device_ids[j]) return fail(nullptr, PTB_ERR_ARG, "ptb_create_multi: a device is listed twice");
    ptb_ctx *ctx = nullptr;
    int rc = ptb_create(device_ids[0], &ctx);
    if (rc != PTB_OK) return rc;
    for (int i = 1; i < n_devices; ++i) {
        ptb_ctx *m = nullptr;
        if ((rc = ptb_create(device_ids[i], &m)) != PTB_OK) { ptb_destroy(ctx); return rc; }
        ctx->members.push_back(m);
    }
    // peer access between every pair: the reduce kernel of device g loads from all framebuffers and stores into device 0's
    bool all = n_devices > 1;
    for (int i = 0; i < n_devices && all; ++i)
        for (int j = 0; j < n_devices && all; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]) != cudaSuccess || !can) all = false;
        }
    if (all)
        for (int i = 0; i < n_devices; ++i) {
            cudaSetDevice(device_ids[i]);
            for (int j = 0; j < n_devices; ++j) {
                if (i == j) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) all = false;
                cudaGetLastError();
            }
        }
    ctx->peer_access = all;
    cudaSetDevice(device_ids[0]);
    *out = ctx;
    return PTB_OK;
}

extern "C" int ptb_device_ids(const ptb_ctx *ctx, int *ids, int cap) {
    if (!ctx) return 0;
    const int n = 1 + (int)ctx->members.size();
    if (ids)
        for (int g = 0; g < n && g < cap; ++g) ids[g] = g == 0 ? ctx->device : ctx->members[g - 1]->device;
    return n;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: device buffers, CUDA IPC, peer reduce + resolve
// ------------------------------------------------------------------------------------------------
extern "C" int ptb_device_alloc(ptb_ctx *ctx, uint64_t n_bytes, void **d_ptr) {
    if (!ctx || !d_ptr || n_bytes == 0) return fail(ctx, PTB_ERR_ARG, "ptb_device_alloc: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMalloc(d_ptr, n_bytes));
    return PTB_OK;
}
extern "C" int ptb_device_free(ptb_ctx *ctx, void *d_ptr) {
    if (!ctx) return fail(ctx, PTB_ERR_ARG, "ptb_device_free: ctx is null");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaFree(d_ptr));
    return PTB_OK;
}
extern "C" int ptb_device_memset(ptb_ctx *ctx, void *d_ptr, int value, uint64_t n_bytes, void *cuda_stream) {
    if (!ctx || !d_ptr) return fail(ctx, PTB_ERR_ARG, "ptb_device_memset: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(d_ptr, value, n_bytes, static_cast<cudaStream_t>(cuda_stream)));
    return PTB_OK;
}
extern "C" int ptb_device_to_host(ptb_ctx *ctx, void *host_dst, const void *d_src, uint64_t n_bytes, void *cuda_stream) {
    if (!ctx || !host_dst || !d_src) return fail(ctx, PTB_ERR_ARG, "ptb_device_to_host: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(host_dst, d_src, n_bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(cuda_stream)));
    CU(ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    return PTB_OK;
}
extern "C" int ptb_device_sync(ptb_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(ctx, PTB_ERR_ARG, "ptb_device_sync: ctx is null");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    return PTB_OK;
}
extern "C" int ptb_ipc_export(ptb_ctx *ctx, const void *d_ptr, unsigned char handle64[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!ctx || !d_ptr || !handle64) return fail(ctx, PTB_ERR_ARG, "ptb_ipc_export: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    std::memcpy(handle64, &h, 64);
    return PTB_OK;
}
extern "C" int ptb_ipc_open(ptb_ctx *ctx, const unsigned char handle64[64], void **d_ptr) {
    if (!ctx || !d_ptr || !handle64) return fail(ctx, PTB_ERR_ARG, "ptb_ipc_open: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    CU(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PTB_OK;
}
extern "C" int ptb_ipc_close(ptb_ctx *ctx, void *d_ptr) {
    if (!ctx || !d_ptr) return fail(ctx, PTB_ERR_ARG, "ptb_ipc_close: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaIpcCloseMemHandle(d_ptr));
    return PTB_OK;
}
extern "C" int ptb_peer_reduce_resolve(ptb_ctx *ctx, const float *const *d_peer_sums, int n_peers, uint64_t first_float,
                                       uint64_t n_floats, uint64_t spp_total, float *d_dst, void *cuda_stream) {
    if (!ctx || !d_peer_sums || !d_dst || spp_total == 0) return fail(ctx, PTB_ERR_ARG, "ptb_peer_reduce_resolve: bad argument");
    if (n_peers < 1 || n_peers > MAX_PEERS) return fail(ctx, PTB_ERR_LIMIT, "ptb_peer_reduce_resolve: 1..16 peers");
    CU(ctx, cudaSetDevice(ctx->device));
    PeerPtrs pp{};
    for (int g = 0; g < n_peers; ++g) {
        if (!d_peer_sums[g]) return fail(ctx, PTB_ERR_ARG, "ptb_peer_reduce_resolve: null peer pointer");
        pp.p[g] = d_peer_sums[g];
    }
    CU(ctx, launch_peer_reduce_resolve(pp, n_peers, first_float, n_floats, spp_total, d_dst, ctx->sm_count, static_cast<cudaStream_t>(cuda_stream)));
    return PTB_OK;
}

// ------------------------------------------------------------------------------------------------
// parity hooks
// ------------------------------------------------------------------------------------------------
static int run_intersect(ptb_ctx *ctx, const float *rays6, uint64_t n, int pw, int ph, int32_t *obj, int32_t *tri, float *t,
                         float *point3, float *normal3) {
    if (!ctx->has_scene) return fail(ctx, PTB_ERR_STATE, "no scene uploaded");
    if (!obj) return fail(ctx, PTB_ERR_ARG, "obj output is null");
    if (n == 0) return PTB_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t fl = (rays6 ? 6 * n : 0) + n + (point3 ? 3 * n : 0) + (normal3 ? 3 * n : 0);
    CU(ctx, ctx->scratch_f.resize(fl + 4));
    CU(ctx, ctx->scratch_i.resize(2 * n));
    float *d_rays = nullptr, *d_t = ctx->scratch_f.p, *d_point = nullptr, *d_normal = nullptr;
    size_t off = n;
    if (rays6) { d_rays = ctx->scratch_f.p + off; off += 6 * n; }
    if (point3) { d_point = ctx->scratch_f.p + off; off += 3 * n; }
    if (normal3) { d_normal = ctx->scratch_f.p + off; off += 3 * n; }
    int *d_obj = ctx->scratch_i.p, *d_tri = ctx->scratch_i.p + n;
    cudaStream_t st = ctx->stream;
    if (rays6) CU(ctx, cudaMemcpyAsync(d_rays, rays6, 6 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(ctx, launch_intersect(ctx->ds, d_rays, n, pw, ph, d_obj, d_tri, d_t, d_point, d_normal, ctx->sm_count, st));
    ctx->stats.kernel_launches = 1;
    CU(ctx, cudaMemcpyAsync(obj, d_obj, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (tri) CU(ctx, cudaMemcpyAsync(tri, d_tri, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (t) CU(ctx, cudaMemcpyAsync(t, d_t, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (point3) CU(ctx, cudaMemcpyAsync(point3, d_point, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (normal3) CU(ctx, cudaMemcpyAsync(normal3, d_normal, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return PTB_OK;
}

extern "C" int ptb_primary_hits(ptb_ctx *ctx, int width, int height, int32_t *obj, int32_t *tri, float *t) {
    if (!ctx) return fail(nullptr, PTB_ERR_ARG, "ptb_primary_hits: ctx is null");
    if (width <= 0 || height <= 0) return fail(ctx, PTB_ERR_ARG, "ptb_primary_hits: bad resolution");
    return run_intersect(ctx, nullptr, static_cast<uint64_t>(width) * static_cast<uint64_t>(height), width, height, obj, tri, t,
                         nullptr, nullptr);
}

extern "C" int ptb_intersect(ptb_ctx *ctx, const float *rays6, uint64_t n, int32_t *obj, int32_t *tri, float *t, float *point3,
                             float *normal3) {
    if (!ctx) return fail(nullptr, PTB_ERR_ARG, "ptb_intersect: ctx is null");
    if (n && !rays6) return fail(ctx, PTB_ERR_ARG, "ptb_intersect: rays is null");
    return run_intersect(ctx, rays6, n, 0, 0, obj, tri, t, point3, normal3);
}
