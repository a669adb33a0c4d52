// pt_bvh.cuh -- BVH traversal for the big-mesh / many-sphere part of the scene (device side).
//
// The reference is brute force (src/render/mod.rs:631-659 scans every object, :558 every triangle).  A BVH only
// changes WHICH primitives get tested, never the arithmetic of a test, so the closest hit stays bit-identical as
// long as (1) no primitive the reference would accept is culled and (2) ties are broken like the reference's scan.
//   (1) every box is padded at build time by a bound on how far the reference's fp32 Moeller-Trumbore / sphere
//       test can place an accepted hit from the true primitive (pt_bvh_build.cu: triangle_pad / sphere_extent),
//       and a node is skipped only if its padded slab interval lies strictly beyond the current best t;
//   (2) `prio` = rank in the reference's scan order; at equal t the lower prio wins.
// Mesh hits additionally need the mesh's bounding-sphere gate to pass (mod.rs:267-277); it is evaluated lazily, only
// when a triangle would become the best hit, and cached per object.
//
// Node = 64 bytes = 4 x float4: both children's boxes + two child references.
//   n0 = (c0.lo.x, c0.lo.y, c0.lo.z, c0.hi.x)   n1 = (c0.hi.y, c0.hi.z, c1.lo.x, c1.lo.y)
//   n2 = (c1.lo.z, c1.hi.x, c1.hi.y, c1.hi.z)   n3 = (ref0, ref1, -, -) as int bits
// ref >= 0: inner node index;  ref < 0: leaf, ~ref = (first_prim << 3) | (count - 1), prims contiguous in bvh_tri.
// Primitive record = 3 x float4 (same as the shared-memory triangle record); a sphere is stored as
//   (centre | obj), (radius^2, 0, 0 | -1), (0, 0, 0 | prio).
#pragma once
#include "pt_device.cuh"

namespace ptb {

constexpr int BVH_STACK = 64;
constexpr int BVH_EMPTY_REF = (int)0x80000000;

__device__ __forceinline__ float safe_rcp_dir(float d) {
    // a zero (or denormal) direction component would give inf * 0 = NaN in the slab test; 1e-20 keeps it finite
    const float a = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
    return __frcp_rn(a);
}

// returns entry distance of the padded box, or a negative number if [0, tmax] misses it
__device__ __forceinline__ bool slab(float lx, float ly, float lz, float hx, float hy, float hz, V3 id, V3 ood, float tmax,
                                     float &t_in) {
    const float x0 = __fmaf_rn(lx, id.x, -ood.x), x1 = __fmaf_rn(hx, id.x, -ood.x);
    const float y0 = __fmaf_rn(ly, id.y, -ood.y), y1 = __fmaf_rn(hy, id.y, -ood.y);
    const float z0 = __fmaf_rn(lz, id.z, -ood.z), z1 = __fmaf_rn(hz, id.z, -ood.z);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    t_in = tn;
    return tn <= tf;
}

// While-while traversal (Aila & Laine 2009): every lane first descends through inner nodes until it holds a leaf (or is
// done), then all lanes that hold a leaf test its primitives together; the leaf code, which is the longest, then runs
// with many lanes instead of one or two.  The stack stores the entry distance of a postponed child so that it can be
// dropped at pop time once a closer hit is known (strictly farther only: ties must still be visited for the prio rule).
__device__ __forceinline__ void bvh_closest_hit(const DScene &sc, V3 o, V3 d, Hit &best) {
    int cur = sc.bvh_root;
    if (cur == BVH_EMPTY_REF) return;
    const V3 id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    const V3 ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
    int stack_ref[BVH_STACK];
    float stack_t[BVH_STACK];
    int sp = 0;
    int gate_obj = -1;
    bool gate_pass = false;
#define PTB_BVH_POP()                                                    \
    do {                                                                 \
        cur = BVH_EMPTY_REF;                                             \
        while (sp > 0) {                                                 \
            --sp;                                                        \
            if (stack_t[sp] <= best.t) { cur = stack_ref[sp]; break; }   \
        }                                                                \
    } while (0)
    while (cur != BVH_EMPTY_REF) {
        while (cur >= 0) {
            const float4 *n = sc.bvh_nodes + 4 * (size_t)cur;
            const float4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
            float t0, t1;
            const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, id, ood, best.t, t0);
            const bool h1 = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, id, ood, best.t, t1);
            const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool first0 = t0 <= t1;
                cur = first0 ? r0 : r1;
                stack_ref[sp] = first0 ? r1 : r0;
                stack_t[sp] = first0 ? t1 : t0;
                sp++;
            } else if (h0) cur = r0;
            else if (h1) cur = r1;
            else PTB_BVH_POP();
        }
        if (cur != BVH_EMPTY_REF) {
            const int code = ~cur;
            const int first = code >> 3, count = (code & 7) + 1;
            for (int k = first; k < first + count; ++k) {
                const float4 A = __ldg(&sc.bvh_tri[3 * k]), E1 = __ldg(&sc.bvh_tri[3 * k + 1]), E2 = __ldg(&sc.bvh_tri[3 * k + 2]);
                const bool is_sphere = __float_as_int(E1.w) < 0;
                float tt;
                if (is_sphere) tt = sphere_t(xyz(A), E1.x, o, d);
                else tt = triangle_t(xyz(A), xyz(E1), xyz(E2), o, d);
                const uint32_t prio = (uint32_t)__float_as_int(E2.w);
                if (tt > 0.0f && (tt < best.t || (tt == best.t && prio < best.prio))) {
                    bool ok = true;
                    if (!is_sphere) {  // mesh gate (mod.rs:267-277), evaluated lazily and cached per object
                        const int obj = __float_as_int(A.w);
                        if (obj != gate_obj) {
                            const float4 g = __ldg(&sc.obj_gate[obj]);
                            gate_pass = sphere_gate(xyz(g), g.w, o, d);
                            gate_obj = obj;
                        }
                        ok = gate_pass;
                    }
                    if (ok) {
                        best.t = tt;
                        best.prio = prio;
                        best.ref = REF_BVH_BIT | (is_sphere ? REF_SPHERE_BIT : 0) | k;
                    }
                }
            }
            PTB_BVH_POP();
        }
    }
#undef PTB_BVH_POP
}

}  // namespace ptb
