// pt_wavefront.cu -- wavefront integrator for BVH scenes (big meshes / many spheres) and for frames too small to fill the
// megakernel's persistent warps.
//
// Same semantics as the megakernel (pt_kernels.cu) and as the oracle's forward twin: identical Philox events, identical
// arithmetic, identical per-pixel summation order, hence a bit-identical framebuffer.  What changes is the schedule.
// In a megakernel every lane walks its own BVH path and the warp runs at the pace of its slowest lane (measured: 6.5 of
// 32 lanes active).  Here a batch of K samples of every pixel is in flight as a queue of ray segments in HBM and each
// bounce is two kernels:
//   k_wf_trace  persistent warps; a lane whose traversal ends writes its hit and is refilled with the next ray of the trace
//               list as soon as `refill` lanes of the warp are idle (Aila & Laine 2009), so traversal runs with mostly full warps;
//   k_wf_shade  one thread per segment: Philox block, hit point / normal, material arm (shade_hit), closest hit of the NEXT
//               segment over the shared-memory list, append of the continuation (and of the transmitted child of a
//               deterministic refraction split) to the next queue with one warp-aggregated atomic; a finished branch stores
//               its emission sum in its slot.
// Only segments that can reach BVH geometry are traced: generate / shade test the segment [0, t_loose] against the BVH's root
// box (the slab test the traversal itself would start with; monotone rounding makes "misses the root box" imply "misses every
// child box").  The queue is filled from both ends: segments to trace take indices 0, 1, 2, ... (the trace kernel reads them
// as one dense, coalesced range), the others cap-1, cap-2, ...; one 64-bit atomic per warp advances both counts.  A scene
// without a BVH launches no trace kernel at all.
// Branches of a path tree are independent queue entries: event ids depend only on (branch code, depth) and every branch
// sums its own emission, so nothing depends on the order in which they are processed.  After MAX_DEPTH bounces
// k_wf_accumulate adds ((L0+L1)+L2)+L3 of samples s0..s0+K-1 to the pixel sum in sample order (mod.rs:846); which of the
// branches 1..3 exist is recorded per path by the split that creates them (branch_mask), so slots are neither cleared nor
// read for branches that never existed.
// Queue entries are 4 x float4 (o|path, d|depth+code, T, L) in SoA arrays: coalesced 16-byte loads and stores.
#include "pt_bvh8.cuh"
#include "pt_bvh8.h"
#include "pt_launch.h"
#include "pt_scene_dev.cuh"
#include "pt_wavefront.h"

namespace ptb {

namespace {

constexpr int WF_CHUNK = 512;  // most rays a warp takes from the queue per global atomic
constexpr int WF_SSTACK = 12;  // traversal-stack entries per lane kept in shared memory (deeper entries go to local memory)

// closest hit of a new segment over the shared-memory list, and whether the segment can reach BVH geometry at all
// (called by the lanes of `amask` only: they are all live)
__device__ __forceinline__ bool loose_hit(const DScene &sc, const float4 *s_obj, V3 o, V3 d, unsigned amask, Hit &best) {
    best.t = __int_as_float(0x7f800000); best.prio = PRIO_NONE; best.ref = REF_NONE;
    closest_hit_loose(s_obj, o, d, amask, best);
    if (sc.bvh_root == BVH_EMPTY_REF) return true;  // (uniform) no BVH, no trace kernel: everything queues up from the front
    const V3 id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    const V3 ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
    float t_in;
    return slab(sc.bvh_lo.x, sc.bvh_lo.y, sc.bvh_lo.z, sc.bvh_hi.x, sc.bvh_hi.y, sc.bvh_hi.z, id, ood, best.t, t_in);
}

// Two-ended append, one atomic per warp (all 32 lanes call).  A lane adds up to two entries (a, b), each either to the front
// (segments the trace kernel must see) or to the back of the queue.  ctr[0] = front count, ctr[1] = back count (one aligned
// 64-bit word).  Returns the queue indices through ia / ib (-1: nothing to add, or the PTB_CHECK build refused it).
__device__ __forceinline__ void append2(const DScene &sc, const WfQueue &q, int *__restrict__ ctr, bool has_a, bool front_a, bool has_b,
                                        bool front_b, unsigned lt_mask, int &ia, int &ib) {
    const unsigned fa = __ballot_sync(0xffffffffu, has_a && front_a), ba = __ballot_sync(0xffffffffu, has_a && !front_a);
    const unsigned fb = __ballot_sync(0xffffffffu, has_b && front_b), bb = __ballot_sync(0xffffffffu, has_b && !front_b);
    ia = ib = -1;
    if ((fa | ba | fb | bb) == 0u) return;
    const unsigned n_front = __popc(fa) + __popc(fb), n_back = __popc(ba) + __popc(bb);
    unsigned long long old = 0;
    if ((threadIdx.x & 31) == 0) old = atomicAdd(reinterpret_cast<unsigned long long *>(ctr), ((unsigned long long)n_back << 32) | n_front);
    old = __shfl_sync(0xffffffffu, old, 0);
    const int f0 = (int)(unsigned)old, b0 = (int)(unsigned)(old >> 32);
    if (!PTB_CHECKED((long long)f0 + n_front + b0 + n_back <= (long long)q.cap, PTB_CHK_QUEUE, sc.check)) return;
    if (has_a) ia = front_a ? f0 + __popc(fa & lt_mask) : q.cap - 1 - (b0 + __popc(ba & lt_mask));
    if (has_b) ib = front_b ? f0 + __popc(fa) + __popc(fb & lt_mask) : q.cap - 1 - (b0 + __popc(ba) + __popc(bb & lt_mask));
}

__global__ void __launch_bounds__(256) k_wf_generate(const DScene sc, int W, int H, unsigned npix, unsigned long long s0, unsigned K,
                                                     unsigned long long seed, WfQueue q, int *__restrict__ branch_mask,
                                                     int *__restrict__ ctr) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);
    const unsigned n = npix * K;  // < 2^29 (checked by the host)
    const unsigned stride = gridDim.x * blockDim.x;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const unsigned n_round = (n + 31u) & ~31u;
    const unsigned lt_mask = (1u << (threadIdx.x & 31)) - 1u;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
        const bool valid = p < n;
        const unsigned amask = __ballot_sync(0xffffffffu, valid);
        V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);
        Hit best;
        bool front = false;
        if (valid) {
            const uint32_t pixel = p % npix;
            const unsigned long long s = s0 + p / npix;
            const int row = (int)(pixel / (uint32_t)W), px = (int)(pixel % (uint32_t)W);
            uint32_t rnd[4];
            philox4x32_10(pixel, (uint32_t)s, (uint32_t)(s >> 32), 0u, k0, k1, rnd);  // camera sample: event 0, slots 0,1
            const float ysub = (float)((s / 2) % 2), xsub = (float)(s % 2);
            const float r1 = 2.0f * u32_to_unit(rnd[0]);
            const float r2 = 2.0f * u32_to_unit(rnd[1]);
            camera_ray(sc, W, H, px, H - 1 - row, xsub, ysub, tent(r1), tent(r2), o, d);
            front = loose_hit(sc, s_obj, o, d, amask, best);
            __stcs(&branch_mask[p], 0);
        }
        int j, unused;
        append2(sc, q, ctr, valid, front, false, false, lt_mask, j, unused);
        if (j >= 0) {  // streaming stores: queue data is written once and read once, it must not push the BVH out of the L2
            __stcs(&q.o[j], make_float4(o.x, o.y, o.z, __int_as_float((int)p)));
            __stcs(&q.d[j], make_float4(d.x, d.y, d.z, __int_as_float(0)));
            __stcs(&q.T[j], make_float4(1.f, 1.f, 1.f, 0.f));
            __stcs(&q.L[j], make_float4(0.f, 0.f, 0.f, 0.f));
            __stcs(&q.hit_t[j], best.t); __stcs(&q.hit_ref[j], best.ref); __stcs(&q.hit_prio[j], best.prio);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// closest BVH hit of every segment on the trace list
// (measured in round 1: 5 or 6 CTAs per SM with the spills that takes, prefetching the leaf while a lane waits, four lanes per ray
//  and tracing in (octant, Morton cell) order are all slower or neutral: profiles/r01g, r01j; tools/experiments/)
// Shared memory: [copy of the top levels of the BVH (sc.n_bvh_top nodes, 144-byte pitch)] [traversal stacks]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TOP_PITCH = 9;  // float4 per shared-memory node: 128 bytes of node + 16 bytes that spread the nodes over the banks

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_wf_trace(const DScene sc, const WfQueue q, const int *__restrict__ ctr,
                                                                     int *__restrict__ fetch_ptr, unsigned long long *__restrict__ counters,
                                                                     const int wf_refill, const int wf_descend_min) {
    extern __shared__ float4 smem[];
    float4 *const s_top = smem;
    int2 *const s_stack = reinterpret_cast<int2 *>(smem + TOP_PITCH * sc.n_bvh_top);
    for (int i = threadIdx.x; i < 8 * sc.n_bvh_top; i += THREADS) s_top[(i >> 3) * TOP_PITCH + (i & 7)] = __ldg(&sc.bvh_top[i]);
    if (sc.n_bvh_top) __syncthreads();
    const int n = ctr[0];  // the front part of the queue: the segments that can reach BVH geometry
    // Rays are handed out in chunks (one same-address atomic per chunk instead of one per refill).  The chunk shrinks with the
    // queue so that every warp of the grid gets about four of them: with a fixed 512 a late bounce (or a scene where few segments
    // reach the BVH) would be traced by a handful of warps while the rest of the GPU idles (measured: 0.8-1.0 ms per launch
    // whatever the queue length, profiles/r02c).
    const int chunk = max(32, min(WF_CHUNK, (n / (int)(gridDim.x * (THREADS / 32) * 4)) & ~31));
    const int root_ref = sc.n_bvh_top ? BVH_TOP_BIT : sc.bvh_root;  // top node 0 is the root's copy
    unsigned n_nodes = 0, n_prims = 0;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    bool busy = false, exhausted = false;
    int w_next = 0, w_end = 0;  // warp-uniform
    int ray_idx = 0;
    V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f), id = mk3(1.f, 1.f, 1.f), ood = mk3(0.f, 0.f, 0.f);
    Hit best;
    best.t = 0.f; best.prio = PRIO_NONE; best.ref = REF_NONE;
    int cur = BVH_EMPTY_REF, sp = 0, gate_obj = -1;
    bool gate_pass = false;
    // traversal stack: the hot top of it lives in shared memory (entry-major, conflict-free 8-byte accesses), so a pop
    // costs a fixed ~30 cycles instead of a local-memory load that competes with node data for the L1
    int2 l_stack[BVH_STACK - WF_SSTACK];
#define PTB_STK_LD(i) ((i) < WF_SSTACK ? s_stack[(i) * THREADS + threadIdx.x] : l_stack[(i) - WF_SSTACK])
#define PTB_STK_ST(i, v)                                                        \
    do {                                                                        \
        if ((i) < WF_SSTACK) s_stack[(i) * THREADS + threadIdx.x] = (v);        \
        else l_stack[(i) - WF_SSTACK] = (v);                                    \
    } while (0)

    for (;;) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, busy);
        const int n_idle = 32 - __popc(busy_mask);
        if (!exhausted && (n_idle >= wf_refill || busy_mask == 0u)) {
            // the warp owns a private chunk [w_next, w_end) of the queue and only touches the global cursor when it runs dry
            if (w_next >= w_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fetch_ptr, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                w_next = base;
                w_end = min(base + chunk, n);
                if (base >= n) exhausted = true;
            }
            const int idx = w_next + __popc(~busy_mask & lt_mask);
            const bool got = !busy && idx < w_end;
            w_next = min(w_next + n_idle, w_end);
            if (got) {
                const int r = idx;
                if (PTB_CHECKED(r >= 0 && r < q.cap, PTB_CHK_RAY, sc.check)) {
                    const float4 qo = __ldcs(&q.o[r]), qd = __ldcs(&q.d[r]);  // streaming: keep the L2 for the BVH
                    o = mk3(qo.x, qo.y, qo.z); d = mk3(qd.x, qd.y, qd.z);
                    ray_idx = r;
                    best.t = __ldcs(&q.hit_t[r]); best.ref = __ldcs(&q.hit_ref[r]); best.prio = __ldcs(&q.hit_prio[r]);
                    id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                    ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
                    cur = root_ref; sp = 0; gate_obj = -1;
                    busy = true;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        if (busy) {
            while (cur >= 0) {  // descend until this lane holds a leaf or is done
                F8 a01_, a23_, b01_, b23_;
                if (cur & BVH_TOP_BIT) {  // one of the top levels: shared-memory copy
                    const float4 *nd_ = s_top + TOP_PITCH * (cur & (BVH_TOP_BIT - 1));
                    a01_.a = nd_[0]; a01_.b = nd_[1]; a23_.a = nd_[2]; a23_.b = nd_[3];
                    b01_.a = nd_[4]; b01_.b = nd_[5]; b23_.a = nd_[6]; b23_.b = nd_[7];
                } else {
                    const float4 *nd_ = sc.bvh_nodes + 8 * (size_t)cur;
                    (void)PTB_CHECKED(cur < sc.n_bvh_nodes, PTB_CHK_NODE, sc.check);
                    a01_ = ld256(nd_); a23_ = ld256(nd_ + 2); b01_ = ld256(nd_ + 4); b23_ = ld256(nd_ + 6);
                }
                PTB_BVH_NODE_TEST();
                n_nodes++;
                if (__popc(__activemask()) < wf_descend_min) break;  // let the lanes that hold a leaf get on with it
            }
            if (cur < 0 && cur != BVH_EMPTY_REF) {
                n_prims += ((~cur) & 7) + 1;
                PTB_BVH_LEAF();
            }
            if (cur == BVH_EMPTY_REF) {
                __stcs(&q.hit_t[ray_idx], best.t);
                __stcs(&q.hit_ref[ray_idx], best.ref);
                busy = false;
            }
        }
    }
#undef PTB_STK_LD
#undef PTB_STK_ST
    for (int off = 16; off > 0; off >>= 1) {
        n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
        n_prims += __shfl_down_sync(0xffffffffu, n_prims, off);
    }
    if (lane == 0 && (n_nodes | n_prims)) {
        atomicAdd(&counters[1], (unsigned long long)n_nodes);
        atomicAdd(&counters[2], (unsigned long long)n_prims);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same job through the compressed eight-wide BVH (pt_bvh8.h): one node fetch tests eight quantised child boxes, the children
// are visited in the fixed order of the ray's direction octant (nothing is sorted) and ONE stack entry -- (child base, hit mask)
// -- stands for all postponed children of a node.  Primitives of leaf children that are hit are tested right after the node.
// Shared memory: [the first nodes in breadth-first order (<= BVH8_TOP_MAX, 112-byte pitch)] [traversal stacks]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TOP8_PITCH = 7;  // uint4 per shared-memory node: 80 bytes of node + padding that spreads diverged lanes over the banks

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_wf_trace8(const DScene sc, const WfQueue q, const int *__restrict__ ctr,
                                                                      int *__restrict__ fetch_ptr, unsigned long long *__restrict__ counters,
                                                                      const int wf_refill, const int wf_descend_min) {
    extern __shared__ float4 smem[];
    uint4 *const s_top = reinterpret_cast<uint4 *>(smem);
    const int n_top = sc.n_bvh8_top;
    int2 *const s_stack = reinterpret_cast<int2 *>(s_top + TOP8_PITCH * n_top);
    for (int i = threadIdx.x; i < 5 * n_top; i += THREADS) s_top[(i / 5) * TOP8_PITCH + (i % 5)] = __ldg(&sc.bvh8_nodes[(i / 5) * BVH8_NODE_F4 + (i % 5)]);
    __syncthreads();
    const int n = ctr[0];  // the front part of the queue: the segments that can reach BVH geometry
    const int chunk = max(32, min(WF_CHUNK, (n / (int)(gridDim.x * (THREADS / 32) * 4)) & ~31));
    unsigned n_nodes = 0, n_prims = 0;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    bool busy = false, exhausted = false;
    int w_next = 0, w_end = 0;  // warp-uniform
    int ray_idx = 0;
    V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f), id = mk3(1.f, 1.f, 1.f);
    Hit best;
    best.t = 0.f; best.prio = PRIO_NONE; best.ref = REF_NONE;
    int sp = 0, gate_obj = -1;
    bool gate_pass = false;
    unsigned ng_base = 0, ng_mask = 0, tg_base = 0, tg_mask = 0, octinv = 0;  // current node group / primitive group
    int2 l_stack[BVH_STACK - WF_SSTACK];
#define PTB_STK_LD(i) ((i) < WF_SSTACK ? s_stack[(i) * THREADS + threadIdx.x] : l_stack[(i) - WF_SSTACK])
#define PTB_STK_ST(i, v)                                                        \
    do {                                                                        \
        if ((i) < WF_SSTACK) s_stack[(i) * THREADS + threadIdx.x] = (v);        \
        else if (PTB_CHECKED((i) < BVH_STACK, PTB_CHK_STACK, sc.check)) l_stack[(i) - WF_SSTACK] = (v); \
    } while (0)

    for (;;) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, busy);
        const int n_idle = 32 - __popc(busy_mask);
        if (!exhausted && (n_idle >= wf_refill || busy_mask == 0u)) {
            if (w_next >= w_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fetch_ptr, chunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                w_next = base;
                w_end = min(base + chunk, n);
                if (base >= n) exhausted = true;
            }
            const int idx = w_next + __popc(~busy_mask & lt_mask);
            const bool got = !busy && idx < w_end;
            w_next = min(w_next + n_idle, w_end);
            if (got) {
                const int r = idx;
                if (PTB_CHECKED(r >= 0 && r < q.cap, PTB_CHK_RAY, sc.check)) {
                    const float4 qo = __ldcs(&q.o[r]), qd = __ldcs(&q.d[r]);  // streaming: keep the L2 for the BVH
                    o = mk3(qo.x, qo.y, qo.z); d = mk3(qd.x, qd.y, qd.z);
                    ray_idx = r;
                    best.t = __ldcs(&q.hit_t[r]); best.ref = __ldcs(&q.hit_ref[r]); best.prio = __ldcs(&q.hit_prio[r]);
                    id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                    octinv = bvh8_octinv(id.x, id.y, id.z);
                    ng_base = 0u; ng_mask = 0x80000000u;  // "the root is a hit inner child of a group that starts at node 0"
                    tg_mask = 0u; sp = 0; gate_obj = -1;
                    busy = true;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        if (busy) {
            for (;;) {  // node phase: visit postponed children until primitives turn up or the ray is done
                if (ng_mask <= 0x00ffffffu) {  // the current group has no inner child left: take the next group
                    if (sp == 0) break;
                    --sp;
                    const int2 e_ = PTB_STK_LD(sp);
                    ng_base = (unsigned)e_.x; ng_mask = (unsigned)e_.y;
                }
                const int bit = 31 - __clz((int)ng_mask);  // first in the octant's visiting order
                ng_mask &= ~(1u << bit);
                const unsigned slot = (unsigned)(bit - 24) ^ octinv;
                const unsigned node = ng_base + (unsigned)__popc(ng_mask & 0xffu & ((1u << slot) - 1u));
                if (ng_mask > 0x00ffffffu) { PTB_STK_ST(sp, make_int2((int)ng_base, (int)ng_mask)); sp++; }
                // five 16-byte words from the shared-memory copy or from global memory, into the same registers either way
                const bool in_top = (int)node < n_top;
                (void)PTB_CHECKED((int)node < sc.n_bvh8_nodes, PTB_CHK_NODE, sc.check);
                const uint4 *ps_ = s_top + TOP8_PITCH * (in_top ? node : 0u);
                const uint4 *pg_ = sc.bvh8_nodes + (size_t)BVH8_NODE_F4 * (in_top ? 0u : node);
                const uint4 a0 = in_top ? ps_[0] : __ldg(pg_), a1 = in_top ? ps_[1] : __ldg(pg_ + 1), a2 = in_top ? ps_[2] : __ldg(pg_ + 2),
                            a3 = in_top ? ps_[3] : __ldg(pg_ + 3), a4 = in_top ? ps_[4] : __ldg(pg_ + 4);
                Bvh8Node nd;
                nd.w[0] = a0.x; nd.w[1] = a0.y; nd.w[2] = a0.z; nd.w[3] = a0.w; nd.w[4] = a1.x; nd.w[5] = a1.y; nd.w[6] = a1.z; nd.w[7] = a1.w;
                nd.w[8] = a2.x; nd.w[9] = a2.y; nd.w[10] = a2.z; nd.w[11] = a2.w; nd.w[12] = a3.x; nd.w[13] = a3.y; nd.w[14] = a3.z; nd.w[15] = a3.w;
                nd.w[16] = a4.x; nd.w[17] = a4.y; nd.w[18] = a4.z; nd.w[19] = a4.w;
                n_nodes++;
                const unsigned hits = bvh8_node_hits(nd, o.x, o.y, o.z, id.x, id.y, id.z, best.t, octinv, sc.bvh8_magic);
                ng_base = nd.w[4]; ng_mask = (hits & 0xff000000u) | (nd.w[3] >> 24);
                tg_base = nd.w[5]; tg_mask = hits & 0x00ffffffu;
                if (tg_mask) break;
                if (__popc(__activemask()) < wf_descend_min) break;  // let the lanes that hold primitives get on with it
            }
            while (tg_mask) {  // primitives of the leaf children that were hit (reference arithmetic, prio tie-break, lazy gate)
                const int b_ = 31 - __clz((int)tg_mask);
                tg_mask &= ~(1u << b_);
                const int k = (int)tg_base + b_;
                n_prims++;
                if (!PTB_CHECKED(k < sc.n_bvh_prims, PTB_CHK_PRIM, sc.check)) continue;
                const F8 ae_ = ld256(sc.bvh8_tri + 2 * (size_t)k);
                const float4 A = ae_.a, E1 = ae_.b, E2 = __ldg(&sc.bvh8_e2[k]);
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
                        best.t = tt; best.prio = prio;
                        best.ref = REF_BVH_BIT | REF_WIDE_BIT | (is_sphere ? REF_SPHERE_BIT : 0) | k;
                    }
                }
            }
            if (ng_mask <= 0x00ffffffu && sp == 0) {  // nothing left to visit
                __stcs(&q.hit_t[ray_idx], best.t);
                __stcs(&q.hit_ref[ray_idx], best.ref);
                busy = false;
            }
        }
    }
#undef PTB_STK_LD
#undef PTB_STK_ST
    for (int off = 16; off > 0; off >>= 1) {
        n_nodes += __shfl_down_sync(0xffffffffu, n_nodes, off);
        n_prims += __shfl_down_sync(0xffffffffu, n_prims, off);
    }
    if (lane == 0 && (n_nodes | n_prims)) {
        atomicAdd(&counters[1], (unsigned long long)n_nodes);
        atomicAdd(&counters[2], (unsigned long long)n_prims);
    }
}

// material arm of every queued segment; appends the next bounce
// (four CTAs per SM: 64 registers with ~100 bytes of spills beat 92 registers at two CTAs -- mesh.json 1080p 996 -> 1164 Mpaths/s, the
//  kernel waits on its queue loads and needs the warps; three CTAs: 1124)
__global__ void __launch_bounds__(256, 4) k_wf_shade(const DScene sc, const WfQueue q, const int *__restrict__ ctr, WfQueue nq,
                                                  int *__restrict__ nctr, float4 *__restrict__ slots, int *__restrict__ branch_mask,
                                                  unsigned n_paths, unsigned npix, unsigned long long s0, unsigned long long seed,
                                                  unsigned long long *__restrict__ segment_counter) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);
    const int n_front = ctr[0], n = n_front + ctr[1];  // entries 0 .. n_front-1 and cap-1 .. cap-n_back
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(segment_counter, (unsigned long long)n);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int stride = gridDim.x * blockDim.x;
    const int n_round = (n + 31) & ~31;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_round; it += stride) {
        const bool valid = it < n;
        const int i = it < n_front ? it : q.cap - 1 - (it - n_front);
        int n_out = 0;
        float4 c_o, c_d, c_T, c_L, k_o, k_d, k_T;  // continuation and (optional) transmitted child
        if (valid && PTB_CHECKED(i >= 0 && i < q.cap, PTB_CHK_RAY, sc.check)) {
            const float4 qo = __ldcs(&q.o[i]), qd = __ldcs(&q.d[i]), qT = __ldcs(&q.T[i]), qL = __ldcs(&q.L[i]);
            const int path = __float_as_int(qo.w), dc = __float_as_int(qd.w);
            const int depth = dc & 0xff, code = dc >> 8;
            const V3 o = mk3(qo.x, qo.y, qo.z), d = mk3(qd.x, qd.y, qd.z), T = mk3(qT.x, qT.y, qT.z);
            V3 L = mk3(qL.x, qL.y, qL.z);
            Hit h;
            h.t = __ldcs(&q.hit_t[i]); h.ref = __ldcs(&q.hit_ref[i]); h.prio = 0;
            bool cont = false;
            if (h.ref != REF_NONE) {
                const uint32_t pixel = (uint32_t)((unsigned)path % npix);
                const unsigned long long s = s0 + (unsigned)path / npix;
                uint32_t rnd[4];
                philox4x32_10(pixel, (uint32_t)s, (uint32_t)(s >> 32), ((uint32_t)code << 4) | (uint32_t)(depth + 1), k0, k1, rnd);
                int obj, tri;
                V3 x, nn;
                finish_hit(sc, s_obj, s_tri, h, o, d, obj, tri, x, nn);
                const int new_depth = depth + 1;
                ShadeOut so;
                shade_hit(sc, obj, nn, d, T, new_depth, rnd, so);
                if (so.emits) L = L + so.emit;
                if (so.cont) {
                    cont = true;
                    n_out = 1;
                    c_o = make_float4(x.x, x.y, x.z, qo.w);
                    c_d = make_float4(so.d.x, so.d.y, so.d.z, __int_as_float(new_depth | (code << 8)));
                    c_T = make_float4(so.T.x, so.T.y, so.T.z, 0.f);
                    c_L = make_float4(L.x, L.y, L.z, 0.f);
                    if (so.split) {
                        n_out = 2;
                        const int child = code | (1 << (new_depth - 1));
                        k_o = c_o;
                        k_d = make_float4(so.child_d.x, so.child_d.y, so.child_d.z, __int_as_float(new_depth | (child << 8)));
                        k_T = make_float4(so.child_T.x, so.child_T.y, so.child_T.z, 0.f);
                        if (PTB_CHECKED((unsigned)path < n_paths && child >= 1 && child <= 3, PTB_CHK_SLOT, sc.check))
                            atomicOr(&branch_mask[path], 1 << child);  // branch `child` of this path now exists
                    }
                }
            }
            if (!cont && PTB_CHECKED((unsigned)path < n_paths && (unsigned)code < 4u, PTB_CHK_SLOT, sc.check))
                __stcs(&slots[(size_t)code * n_paths + (size_t)path], make_float4(L.x, L.y, L.z, 0.f));  // branch finished
        }
        // closest hit of the new segments over the shared-memory list (so the trace kernel only does BVH work), then one atomic
        // per warp for everything the warp appends
        const unsigned m1 = __ballot_sync(0xffffffffu, n_out >= 1), m2 = __ballot_sync(0xffffffffu, n_out == 2);
        if ((m1 | m2) == 0u) continue;
        Hit h1, h2;
        bool f1 = false, f2 = false;
        if (n_out >= 1) f1 = loose_hit(sc, s_obj, mk3(c_o.x, c_o.y, c_o.z), mk3(c_d.x, c_d.y, c_d.z), m1, h1);
        if (m2 != 0u && n_out == 2) f2 = loose_hit(sc, s_obj, mk3(k_o.x, k_o.y, k_o.z), mk3(k_d.x, k_d.y, k_d.z), m2, h2);
        int j1, j2;
        append2(sc, nq, nctr, n_out >= 1, f1, n_out == 2, f2, lt_mask, j1, j2);
        if (j1 >= 0) {  // (streaming stores, like every queue access: the L2 is for the BVH)
            __stcs(&nq.o[j1], c_o); __stcs(&nq.d[j1], c_d); __stcs(&nq.T[j1], c_T); __stcs(&nq.L[j1], c_L);
            __stcs(&nq.hit_t[j1], h1.t); __stcs(&nq.hit_ref[j1], h1.ref); __stcs(&nq.hit_prio[j1], h1.prio);
        }
        if (j2 >= 0) {
            __stcs(&nq.o[j2], k_o); __stcs(&nq.d[j2], k_d); __stcs(&nq.T[j2], k_T); __stcs(&nq.L[j2], make_float4(0.f, 0.f, 0.f, 0.f));
            __stcs(&nq.hit_t[j2], h2.t); __stcs(&nq.hit_ref[j2], h2.ref); __stcs(&nq.hit_prio[j2], h2.prio);
        }
    }
}

// radiance_v += radiance(sample) for the K samples of the batch, in sample order (mod.rs:846)
__global__ void __launch_bounds__(256) k_wf_accumulate(const float4 *__restrict__ slots, const int *__restrict__ branch_mask,
                                                       unsigned n_paths, unsigned npix, unsigned K, float *__restrict__ sum_rgb,
                                                       const int fb_zero) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += stride) {
        float *fb = sum_rgb + 3ull * pixel;
        V3 acc = fb_zero ? mk3(0.f, 0.f, 0.f) : mk3(fb[0], fb[1], fb[2]);
        for (unsigned k = 0; k < K; ++k) {
            const size_t p = (size_t)k * npix + pixel;
            const int mask = __ldcs(&branch_mask[p]);
            const float4 a = __ldcs(&slots[p]);
            V3 L = mk3(a.x, a.y, a.z);
            if (mask) {  // the path split: ((L0 + L1) + L2) + L3, absent branches count as zero
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 b = (mask & 2) ? __ldcs(&slots[n_paths + p]) : z;
                const float4 c = (mask & 4) ? __ldcs(&slots[2 * (size_t)n_paths + p]) : z;
                const float4 e = (mask & 8) ? __ldcs(&slots[3 * (size_t)n_paths + p]) : z;
                L = ((L + mk3(b.x, b.y, b.z)) + mk3(c.x, c.y, c.z)) + mk3(e.x, e.y, e.z);
            }
            acc = acc + L;
        }
        fb[0] = acc.x; fb[1] = acc.y; fb[2] = acc.z;
    }
}

template <typename T>
cudaError_t wf_alloc(T *&p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(&p), n * sizeof(T)); }

constexpr size_t WF_CTR_INTS = (size_t)WF_CTR_STRIDE * (WF_MAX_BOUNCES + 2);

}  // namespace

void wf_release(WfWorkspace &w) {
    for (int b = 0; b < 2; ++b) {
        WfQueue &q = w.q[b];
        cudaFree(q.o); cudaFree(q.d); cudaFree(q.T); cudaFree(q.L);
        cudaFree(q.hit_t); cudaFree(q.hit_ref); cudaFree(q.hit_prio);
    }
    cudaFree(w.slots); cudaFree(w.branch_mask); cudaFree(w.counters);
    w = WfWorkspace{};
}

static cudaError_t wf_reserve(WfWorkspace &w, size_t n_paths) {
    if (n_paths <= w.cap_paths) return cudaSuccess;
    wf_release(w);
    const size_t cap = 4 * n_paths;  // every path can split twice (mod.rs:775-786): at most 4 live branches
    cudaError_t e;
    for (int b = 0; b < 2; ++b) {
        WfQueue &q = w.q[b];
        if ((e = wf_alloc(q.o, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.d, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.T, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.L, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.hit_t, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.hit_ref, cap)) != cudaSuccess) return e;
        if ((e = wf_alloc(q.hit_prio, cap)) != cudaSuccess) return e;
        q.cap = (int)cap;
    }
    if ((e = wf_alloc(w.slots, 4 * n_paths)) != cudaSuccess) return e;
    if ((e = wf_alloc(w.branch_mask, n_paths)) != cudaSuccess) return e;
    if ((e = wf_alloc(w.counters, WF_CTR_INTS)) != cudaSuccess) return e;
    w.cap_paths = n_paths;
    return cudaSuccess;
}

namespace {

struct TraceLaunch {
    void (*kern)(const DScene, const WfQueue, const int *, int *, unsigned long long *, int, int) = nullptr;
    int threads = 256, blocks = 0;
    size_t smem = 0;
};

template <int THREADS>
cudaError_t trace_config(const DScene &sc, int sm_count, TraceLaunch &t) {
    const bool wide = sc.n_bvh8_nodes > 0;
    t.kern = wide ? k_wf_trace8<THREADS> : k_wf_trace<THREADS>;
    t.threads = THREADS;
    t.smem = (wide ? sizeof(uint4) * TOP8_PITCH * (size_t)sc.n_bvh8_top : sizeof(float4) * TOP_PITCH * (size_t)sc.n_bvh_top) +
             sizeof(int2) * WF_SSTACK * THREADS;
    cudaError_t e = cudaFuncSetAttribute(t.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, t.kern, THREADS, t.smem)) != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    t.blocks = sm_count * per_sm;  // persistent grid: a whole number of resident CTAs per SM
    return cudaSuccess;
}

}  // namespace

// renders samples [a.spp_begin, a.spp_begin + a.spp_count) of every pixel into a.sum_rgb; adds the kernels launched to *launches
cudaError_t wavefront_render(const DScene &scene, const RenderArgs &a, WfWorkspace &w, int sm_count, const WfOptions &opt, cudaStream_t st,
                             unsigned *launches) {
    DScene sc = scene;
    sc.n_bvh8_top = std::max(0, std::min(std::min(sc.n_bvh8_nodes, BVH8_TOP_MAX), opt.top8_nodes));
    const unsigned npix = (unsigned)a.width * (unsigned)a.height;
    unsigned K = (unsigned)std::max<size_t>(1, opt.target_paths / npix);
    if ((unsigned long long)K > a.spp_count) K = (unsigned)a.spp_count;
    if (K == 0) return cudaSuccess;
    const size_t n_paths = (size_t)npix * K;
    if (n_paths >= (1ull << 29)) return cudaErrorInvalidValue;  // queue indices are 32-bit ints, 4 branches per path
    cudaError_t e = wf_reserve(w, n_paths);
    if (e != cudaSuccess) return e;
    const size_t smem = loose_smem_bytes(sc);
    if ((e = cudaFuncSetAttribute(k_wf_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_wf_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    const bool has_bvh = sc.bvh_root != BVH_EMPTY_REF;
    TraceLaunch tl;
    if (has_bvh) {
        // the CTA size trades copies of the top levels per SM against scheduling granularity (opt.trace_threads: 256, 512 or 1024)
        const int threads = sc.n_bvh8_nodes > 0 ? opt.trace_threads_wide : opt.trace_threads;
        if (threads >= 1024) e = trace_config<1024>(sc, sm_count, tl);
        else if (threads >= 512) e = trace_config<512>(sc, sm_count, tl);
        else e = trace_config<256>(sc, sm_count, tl);
        if (e != cudaSuccess) return e;
    }
    auto blocks_for = [&](size_t n) { return (unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)sm_count * 8)); };

    unsigned long long done = 0;
    while (done < a.spp_count) {
        const unsigned k_now = (unsigned)std::min<unsigned long long>(K, a.spp_count - done);
        const size_t paths_now = (size_t)npix * k_now;
        const unsigned long long s0 = a.spp_begin + done;
        const unsigned wide = blocks_for(paths_now);
        // per bounce b: ctr[4b] = entries at the front of its queue (to trace), [4b+1] = entries at its back, [4b+2] = the trace kernel's fetch cursor
        if ((e = cudaMemsetAsync(w.counters, 0, WF_CTR_INTS * sizeof(int), st)) != cudaSuccess) return e;
        k_wf_generate<<<wide, 256, smem, st>>>(sc, a.width, a.height, npix, s0, k_now, a.seed, w.q[0], w.branch_mask, w.counters);
        (*launches)++;
        for (int b = 0; b < WF_MAX_BOUNCES; ++b) {
            const WfQueue &cur = w.q[b & 1], &nxt = w.q[(b + 1) & 1];
            int *ctr = w.counters + WF_CTR_STRIDE * b;
            if (has_bvh) {
                const bool wide = sc.n_bvh8_nodes > 0;
                tl.kern<<<tl.blocks, tl.threads, tl.smem, st>>>(sc, cur, ctr, ctr + 2, a.segment_counter, wide ? opt.refill_wide : opt.refill,
                                                                wide ? opt.descend_min_wide : opt.descend_min);
                (*launches)++;
            }
            k_wf_shade<<<wide, 256, smem, st>>>(sc, cur, ctr, nxt, ctr + WF_CTR_STRIDE, w.slots, w.branch_mask, (unsigned)n_paths, npix, s0,
                                                a.seed, a.segment_counter);
            (*launches)++;
        }
        k_wf_accumulate<<<wide, 256, 0, st>>>(w.slots, w.branch_mask, (unsigned)n_paths, npix, k_now, a.sum_rgb,
                                              (a.fb_zero && done == 0) ? 1 : 0);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        done += k_now;
    }
    return cudaSuccess;
}

}  // namespace ptb
