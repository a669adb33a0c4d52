// pt_scene_dev.cuh -- closest hit over the flattened scene (shared-memory loose list + BVH) and the material arms of
// radiance(), shared by the megakernel (pt_kernels.cu) and the wavefront integrator (pt_wavefront.cu).
#pragma once
#include "pt_bvh.cuh"
#include "pt_device.cuh"

namespace ptb {

constexpr float PI_F = 3.141592653589793f;  // mod.rs:29
constexpr int MAX_DEPTH = 12;               // mod.rs:661

// ---------------------------------------------------------------------------------------------
// closest hit over the shared-memory ("loose") object stream, in the reference's scan order.  Every lane named in `vmask` runs
// this in lock step (broadcast LDS.128).  Lanes that only keep the warp converged ("passengers") pass any finite ray: their
// result is never read, and they can at worst make the warp scan a mesh that no live lane's gate let through.
// ---------------------------------------------------------------------------------------------
// `unit_dir`: the lane's direction has unit length (every ray the integrator generates).  The gate's origin-inside shortcut
// (sphere_gate, r2_inside) is derived for |d| = 1; a caller-supplied ray of another length (ptb_intersect) gets the exact gate.
__device__ __forceinline__ void closest_hit_loose(const float4 *__restrict__ s_obj, V3 o, V3 d, unsigned vmask, Hit &best,
                                                  bool unit_dir = true) {
    const float4 *rec = s_obj;
    for (;;) {
        const float4 sph = rec[0];
        const float4 mb = rec[1];
        const int kind = __float_as_int(mb.x);
        if (kind == KIND_SPHERE) {
            const float t = sphere_t(xyz(sph), sph.w, o, d);
            if (t >= 0.0f && t < best.t) {
                best.t = t;
                best.prio = (uint32_t)__float_as_int(mb.y);
                best.ref = REF_SPHERE_BIT | __float_as_int(mb.z);
            }
            rec += 2;
            continue;
        }
        if (kind == KIND_END) break;
        // mesh: bounding-sphere gate first (mod.rs:267-277)
        const bool pass = sphere_gate(xyz(sph), sph.w, o, d, unit_dir ? mb.x : -1.0f);
        const int n_tri = __float_as_int(mb.z);
        if (n_tri == 0) {
            // one pair (a wall quad) whose gate sphere is large against the scene, marked by the host: some lane nearly always
            // passes, so the pair is tested without asking the warp first -- no vote, no loop; `pass` still decides acceptance
            const int k = __float_as_int(mb.y);
            const float4 *pr = rec + 2;
            float ta, tb;
            bool ha, hb;
            triangle_pair_hit(pr, o, d, ha, ta, hb, tb);
            if (pass && ha && ta < best.t) { best.t = ta; best.prio = (uint32_t)__float_as_int(pr[4].z); best.ref = k; }
            if (pass && hb && tb < best.t) { best.t = tb; best.prio = (uint32_t)__float_as_int(pr[4].w); best.ref = k + 1; }
            rec += 7;
            continue;
        }
        if (__any_sync(vmask, pass)) {  // skip the triangle scan if no lane passes
            int k = __float_as_int(mb.y);  // even: meshes are padded to whole pairs
            const int k1 = k + n_tri;
            const float4 *pr = rec + 2;
            for (; k < k1; k += 2, pr += 5) {  // two triangles per trip, packed multiplies
                float ta, tb;
                bool ha, hb;
                triangle_pair_hit(pr, o, d, ha, ta, hb, tb);
                if (pass && ha && ta < best.t) { best.t = ta; best.prio = (uint32_t)__float_as_int(pr[4].z); best.ref = k; }
                if (pass && hb && tb < best.t) { best.t = tb; best.prio = (uint32_t)__float_as_int(pr[4].w); best.ref = k + 1; }
            }
        }
        rec += __float_as_int(mb.w);
    }
}

// resolves the winning primitive into (object id, triangle id, hit point, geometric normal)
__device__ __forceinline__ void finish_hit(const DScene &sc, const float4 *__restrict__ s_obj, const float4 *__restrict__ s_tri,
                                           const Hit &h, V3 o, V3 d, int &obj, int &tri, V3 &x, V3 &n) {
    x = o + d * h.t;  // mod.rs:430 / :604
    const int k = h.ref & REF_INDEX_MASK;
    if (h.ref & REF_BVH_BIT) {
        const bool wide = (h.ref & REF_WIDE_BIT) != 0;
        const float4 F = __ldg(wide ? &sc.bvh8_fin[k] : &sc.bvh_fin[k]);
        obj = __float_as_int(F.w);
        if (h.ref & REF_SPHERE_BIT) { tri = -1; n = normalize(x - xyz(F)); }
        else { tri = __float_as_int(__ldg(wide ? &sc.bvh8_tri[2 * k + 1] : &sc.bvh_tri[2 * k + 1]).w); n = xyz(F); }
    } else if (h.ref & REF_SPHERE_BIT) {
        const float4 sph = s_obj[k], mb = s_obj[k + 1];
        obj = __float_as_int(mb.w);
        tri = -1;
        n = normalize(x - xyz(sph));  // mod.rs:431
    } else {
        const float4 N = s_tri[2 * k];
        obj = __float_as_int(N.w);
        tri = __float_as_int(s_tri[2 * k + 1].x);
        n = xyz(N);  // normalize(cross(e1, e2)) of mod.rs:605, evaluated once per triangle on the host with the same operations
    }
}

// `live` lanes get their closest hit; all lanes of `vmask` must call (the loose scan votes)
template <bool HAS_BVH>
__device__ __forceinline__ Hit closest_hit(const DScene &sc, const float4 *__restrict__ s_obj, V3 o, V3 d, unsigned vmask, bool live,
                                           bool unit_dir = true) {
    Hit best;
    best.t = __int_as_float(0x7f800000);
    best.prio = PRIO_NONE;
    best.ref = REF_NONE;
    closest_hit_loose(s_obj, o, d, vmask, best, unit_dir);
    if (HAS_BVH) {
        if (live) bvh_closest_hit(sc, o, d, best);
    }
    return best;
}

__device__ __forceinline__ void stage_loose(const DScene &sc, float4 *smem, const float4 *&s_obj, const float4 *&s_tri) {
    const int n0 = sc.n_loose_f4, n1 = 2 * sc.n_loose_tri;
    for (int i = threadIdx.x; i < n0; i += blockDim.x) smem[i] = __ldg(&sc.loose_obj[i]);
    for (int i = threadIdx.x; i < n1; i += blockDim.x) smem[n0 + i] = __ldg(&sc.loose_tri[i]);
    __syncthreads();
    s_obj = smem;
    s_tri = smem + n0;
}
__host__ __device__ __forceinline__ size_t loose_smem_bytes(const DScene &sc) {
    return sizeof(float4) * ((size_t)sc.n_loose_f4 + 2ull * (size_t)sc.n_loose_tri);
}


// ---------------------------------------------------------------------------------------------
// one radiance() call after the closest hit is known (mod.rs:665-790), in throughput form.
// ---------------------------------------------------------------------------------------------
struct ShadeOut {
    bool cont;      // the branch continues with (d, T)
    bool split;     // a transmitted child (child_d, child_T) is spawned (deterministic split, mod.rs:775-786)
    bool emits;     // emit = T * emission must be added to the branch's sum
    V3 d, T, child_d, child_T, emit;
};

__device__ __forceinline__ void shade_hit(const DScene &sc, int obj, V3 n, V3 d_in, V3 T_in, int new_depth, const uint32_t rnd[4],
                                          ShadeOut &out) {
    const float4 mc = __ldg(&sc.mat_color[obj]);
    const float4 me = __ldg(&sc.mat_emis[obj]);
    const int refl = __float_as_int(mc.w);
    V3 color = xyz(mc);
    const float max_reflection = fmaxf(color.x, fmaxf(color.y, color.z));
    const V3 nl = dot(n, d_in) < 0.0f ? n : n * -1.0f;
    bool alive = true;
    if (new_depth > 5) {  // Russian roulette, mod.rs:677-683 (rand01 is drawn before the depth test)
        if (u32_to_unit(rnd[0]) < max_reflection && new_depth < MAX_DEPTH) color = color * PTB_RCP(max_reflection);
        else alive = false;
    }
    out.emits = __float_as_int(me.w) != 0;
    out.emit = T_in * xyz(me);
    out.cont = alive;
    out.split = false;
    out.d = d_in; out.T = T_in; out.child_d = d_in; out.child_T = T_in;
    if (alive) {
        const V3 Tc = T_in * color;
        // The diffuse arm and the refraction arm both end in a normalize(); the lanes of both run it together (`pre`, `norm`),
        // so the few glass lanes of a warp do not make it issue a second copy.  Every lane still performs exactly the
        // operations of its own arm, in the reference's order.
        V3 pre = d_in, rd = d_in;
        bool norm = false, refr = false, into = false;
        float ddn = 0.0f;
        if (refl == 0) {  // Diffuse, mod.rs:687-715
            const float r1 = 2.0f * PI_F * u32_to_unit(rnd[1]);
            const float r2 = u32_to_unit(rnd[2]);
            const float r2s = PTB_SQRT(r2);
            const V3 w = nl;
            const V3 u = normalize(cross(fabsf(w.x) > 0.1f ? mk3(0.f, 1.f, 0.f) : mk3(1.f, 0.f, 0.f), w));
            const V3 v = cross(w, u);
            float sn, cs;
            sincos_det(r1, sn, cs);
            pre = u * cs * r2s + v * sn * r2s + w * PTB_SQRT(1.0f - r2);
            norm = true;
            out.T = Tc;
        } else {
            rd = d_in - n * 2.0f * dot(n, d_in);  // mod.rs:722-723
            out.d = rd; out.T = Tc;               // Specular; also refraction's total internal reflection (mod.rs:743-744)
            if (refl == 2) {                      // Refract, mod.rs:729-788
                into = dot(n, nl) > 0.0f;
                constexpr float NNT_IN = 1.0f / 1.5f, NNT_OUT = 1.5f / 1.0f;  // IEEE divisions folded by the compiler (mod.rs:739)
                const float nnt = into ? NNT_IN : NNT_OUT;
                ddn = dot(d_in, nl);
                const float cos2t = 1.0f - nnt * nnt * (1.0f - ddn * ddn);
                if (!(cos2t < 0.0f)) {
                    pre = d_in * nnt - n * ((into ? 1.0f : -1.0f) * (ddn * nnt + PTB_SQRT(cos2t)));
                    norm = true; refr = true;
                }
            }
        }
        if (norm) {
            const V3 nd = normalize(pre);
            if (!refr) out.d = nd;
            else {
                const V3 tdir = nd;
                constexpr float r0 = (0.5f * 0.5f) / (2.5f * 2.5f);  // (nt-nc)^2 / (nt+nc)^2, folded (mod.rs:750-752)
                const float c = 1.0f - (into ? -ddn : dot(tdir, n));
                const float c2 = c * c;
                const float re = r0 + (1.0f - r0) * (c * (c2 * c2));
                const float tr = 1.0f - re;
                const float p = 0.25f + 0.5f * re;
                if (new_depth > 2) {  // Russian roulette between the two children: one division for whichever is taken
                    const bool reflect = u32_to_unit(rnd[3]) < p;
                    const float num = reflect ? re : tr, den = reflect ? p : 1.0f - p;
                    out.T = Tc * PTB_DIV(num, den);
                    out.d = reflect ? rd : tdir;
                } else {  // deterministic two-way split (mod.rs:776-785)
                    out.split = true;
                    out.child_d = tdir; out.child_T = Tc * tr;
                    out.T = Tc * re; out.d = rd;
                }
            }
        }
    }
}

}  // namespace ptb
