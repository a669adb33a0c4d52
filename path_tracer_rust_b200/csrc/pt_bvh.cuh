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
// Node = 128 bytes = 8 x float4 (one cache line), four children, child-major:
//   [c]     = (lo.x, lo.y, lo.z, hi.x) of child c        [4 + c] = (hi.y, hi.z, ref as int bits, -) of child c
// A single lane fetches it with four 256-bit loads; a sub-warp of four lanes (cooperative trace kernel) with two coalesced
// 64-byte accesses, lane c taking child c.
// A node has two to four children, packed from slot 0; an unused slot carries ref = BVH_EMPTY_REF.
// ref >= 0: inner node index;  ref < 0: leaf, ~ref = (first_prim << 3) | (count - 1), prims contiguous in bvh_tri.
// Primitive record = (A | obj), (E1 | tri) in bvh_tri (32 bytes, one 256-bit load) + (E2 | prio) in bvh_e2; a sphere is
// stored as (centre | obj), (radius^2, 0, 0 | -1), (0, 0, 0 | prio).
#pragma once
#include "pt_device.cuh"

namespace ptb {

constexpr int BVH_STACK = 96;
constexpr int BVH_EMPTY_REF = (int)0x80000000;
constexpr int BVH_TOP_BIT = 1 << 30;  // inner-node ref that indexes DScene::bvh_top (the copy of the top levels) instead of bvh_nodes
constexpr int BVH_TOP_MAX = 341;      // nodes of five complete four-wide levels

__device__ __forceinline__ float safe_rcp_dir(float d) {
    // a zero (or denormal) direction component would give inf * 0 = NaN in the slab test; 1e-20 keeps it finite
    const float a = fabsf(d) < 1e-20f ? copysignf(1e-20f, d) : d;
    return rcp_rn_normal(a);  // |a| in [1e-20, 1]: the normal range, where this is __frcp_rn bit for bit (ptb_selftest)
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

// entry distance of [0, tmax] into the padded box, +inf if it misses
__device__ __forceinline__ float slab_t(float lx, float ly, float lz, float hx, float hy, float hz, V3 id, V3 ood, float tmax) {
    float t;
    return slab(lx, ly, lz, hx, hy, hz, id, ood, tmax, t) ? t : __int_as_float(0x7f800000);
}
__device__ __forceinline__ void cswap(float &ta, int &ra, float &tb, int &rb) {
    const bool s = tb < ta;
    const float t = s ? tb : ta; tb = s ? ta : tb; ta = t;
    const int r = s ? rb : ra; rb = s ? ra : rb; ra = r;
}

// The traversal stack holds (child ref, entry distance) pairs.  PTB_STK_LD(i) reads entry i, PTB_STK_ST(i, v) writes it; the
// includer defines both before using the macros below (the wavefront trace kernel keeps the first entries in shared memory
// and addresses the two memories explicitly: a pointer select would turn every access into a generic LD/ST).
#define PTB_BVH_POP()                                                    \
    do {                                                                 \
        cur = BVH_EMPTY_REF;                                             \
        while (sp > 0) {                                                 \
            --sp;                                                        \
            const int2 e_ = PTB_STK_LD(sp);                              \
            if (__int_as_float(e_.y) <= best.t) { cur = e_.x; break; }   \
        }                                                                \
    } while (0)
#define PTB_BVH_PUSH(ref_, t_)                                           \
    do {                                                                 \
        if (PTB_CHECKED(sp < BVH_STACK, PTB_CHK_STACK, sc.check)) {      \
            PTB_STK_ST(sp, make_int2((ref_), __float_as_int(t_)));       \
            sp++;                                                        \
        }                                                                \
    } while (0)

// One step through a four-wide inner node: test the four child boxes, continue with the nearest one that is hit and
// postpone the others (far to near, so the nearer is popped first) together with their entry distances.
// PTB_BVH_NODE_TEST expects the node in a01_, a23_, b01_, b23_ (children 0|1, 2|3: lo.xyz hi.x, then hi.yz ref -).
#define PTB_BVH_NODE_TEST()                                                                                      \
    do {                                                                                                         \
        float t0_ = slab_t(a01_.a.x, a01_.a.y, a01_.a.z, a01_.a.w, b01_.a.x, b01_.a.y, id, ood, best.t);         \
        float t1_ = slab_t(a01_.b.x, a01_.b.y, a01_.b.z, a01_.b.w, b01_.b.x, b01_.b.y, id, ood, best.t);         \
        float t2_ = slab_t(a23_.a.x, a23_.a.y, a23_.a.z, a23_.a.w, b23_.a.x, b23_.a.y, id, ood, best.t);         \
        float t3_ = slab_t(a23_.b.x, a23_.b.y, a23_.b.z, a23_.b.w, b23_.b.x, b23_.b.y, id, ood, best.t);         \
        int r0_ = __float_as_int(b01_.a.z), r1_ = __float_as_int(b01_.b.z), r2_ = __float_as_int(b23_.a.z), r3_ = __float_as_int(b23_.b.z); \
        const float inf_ = __int_as_float(0x7f800000);                                                           \
        /* a node has two to four children; the slab test cannot reject an empty slot by itself */               \
        if (r2_ == BVH_EMPTY_REF) t2_ = inf_;                                                                    \
        if (r3_ == BVH_EMPTY_REF) t3_ = inf_;                                                                    \
        cswap(t0_, r0_, t1_, r1_); cswap(t2_, r2_, t3_, r3_); cswap(t0_, r0_, t2_, r2_);                         \
        cswap(t1_, r1_, t3_, r3_); cswap(t1_, r1_, t2_, r2_);                                                    \
        if (t0_ < inf_) {                                                                                        \
            cur = r0_;                                                                                           \
            if (t3_ < inf_) PTB_BVH_PUSH(r3_, t3_);                                                              \
            if (t2_ < inf_) PTB_BVH_PUSH(r2_, t2_);                                                              \
            if (t1_ < inf_) PTB_BVH_PUSH(r1_, t1_);                                                              \
        } else PTB_BVH_POP();                                                                                    \
    } while (0)
#define PTB_BVH_NODE_STEP()                                                                                      \
    do {                                                                                                         \
        const float4 *nd_ = sc.bvh_nodes + 8 * (size_t)cur;                                                      \
        (void)PTB_CHECKED(cur < sc.n_bvh_nodes, PTB_CHK_NODE, sc.check);                                             \
        const F8 a01_ = ld256(nd_), a23_ = ld256(nd_ + 2), b01_ = ld256(nd_ + 4), b23_ = ld256(nd_ + 6);         \
        PTB_BVH_NODE_TEST();                                                                                     \
    } while (0)

// Tests the primitives of the leaf `cur` (reference arithmetic, prio tie-break, lazy mesh gate), then pops.
#define PTB_BVH_LEAF()                                                                                           \
    do {                                                                                                         \
        const int code_ = ~cur;                                                                                  \
        const int first_ = code_ >> 3, count_ = (code_ & 7) + 1;                                                 \
        (void)PTB_CHECKED(first_ + count_ <= sc.n_bvh_prims, PTB_CHK_PRIM, sc.check);                                \
        for (int k = first_; k < first_ + count_; ++k) {                                                         \
            const F8 ae_ = ld256(sc.bvh_tri + 2 * (size_t)k);                                                    \
            const float4 A = ae_.a, E1 = ae_.b, E2 = __ldg(&sc.bvh_e2[k]);                                      \
            const bool is_sphere = __float_as_int(E1.w) < 0;                                                     \
            float tt;                                                                                            \
            if (is_sphere) tt = sphere_t(xyz(A), E1.x, o, d);                                                    \
            else tt = triangle_t(xyz(A), xyz(E1), xyz(E2), o, d);                                                \
            const uint32_t prio = (uint32_t)__float_as_int(E2.w);                                                \
            if (tt > 0.0f && (tt < best.t || (tt == best.t && prio < best.prio))) {                              \
                bool ok = true;                                                                                  \
                if (!is_sphere) { /* mesh gate (mod.rs:267-277), evaluated lazily and cached per object */       \
                    const int obj = __float_as_int(A.w);                                                         \
                    if (obj != gate_obj) {                                                                       \
                        const float4 g = __ldg(&sc.obj_gate[obj]);                                               \
                        gate_pass = sphere_gate(xyz(g), g.w, o, d);                                              \
                        gate_obj = obj;                                                                          \
                    }                                                                                            \
                    ok = gate_pass;                                                                              \
                }                                                                                                \
                if (ok) {                                                                                        \
                    best.t = tt; best.prio = prio;                                                               \
                    best.ref = REF_BVH_BIT | (is_sphere ? REF_SPHERE_BIT : 0) | k;                               \
                }                                                                                                \
            }                                                                                                    \
        }                                                                                                        \
        PTB_BVH_POP();                                                                                           \
    } while (0)

// While-while traversal (Aila & Laine 2009): every lane first descends through inner nodes until it holds a leaf (or is
// done), then all lanes that hold a leaf test its primitives together.  The stack stores the entry distance of a postponed
// child so that it can be dropped at pop time once a closer hit is known (strictly farther only: ties must still be
// visited for the prio rule).
__device__ __forceinline__ void bvh_closest_hit(const DScene &sc, V3 o, V3 d, Hit &best) {
    int cur = sc.bvh_root;
    if (cur == BVH_EMPTY_REF) return;
    const V3 id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    const V3 ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
    int2 stack_[BVH_STACK];
#define PTB_STK_LD(i) stack_[i]
#define PTB_STK_ST(i, v) stack_[i] = (v)
    int sp = 0;
    int gate_obj = -1;
    bool gate_pass = false;
    while (cur != BVH_EMPTY_REF) {
        while (cur >= 0) PTB_BVH_NODE_STEP();
        if (cur != BVH_EMPTY_REF) PTB_BVH_LEAF();
    }
#undef PTB_STK_LD
#undef PTB_STK_ST
}

}  // namespace ptb
