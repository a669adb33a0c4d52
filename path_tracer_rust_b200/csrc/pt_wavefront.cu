// pt_wavefront.cu -- wavefront integrator for BVH scenes (big meshes / many spheres).
//
// Same semantics as the megakernel (pt_kernels.cu) and as the oracle's forward twin: identical Philox events, identical
// arithmetic, identical per-pixel summation order, hence a bit-identical framebuffer.  What changes is the schedule.
// In a megakernel every lane walks its own BVH path and the warp runs at the pace of its slowest lane (measured: 6.5 of
// 32 lanes active).  Here a batch of K samples of every pixel is in flight as a queue of ray segments in HBM and each
// bounce is two kernels:
//   k_wf_trace  persistent warps; a lane whose traversal ends writes its hit and is refilled with the next ray of the queue
//               as soon as REFILL lanes of the warp are idle (Aila & Laine 2009), so traversal runs with mostly full warps;
//   k_wf_shade  one thread per segment: Philox block, hit point / normal, material arm (shade_hit), appends the continuation
//               (and the transmitted child of a deterministic refraction split) to the next queue with one warp-aggregated
//               atomic; a finished branch stores its emission sum in its slot.
// Branches of a path tree are independent queue entries: event ids depend only on (branch code, depth) and every branch
// sums its own emission, so nothing depends on the order in which they are processed.  After MAX_DEPTH bounces
// k_wf_accumulate adds ((L0+L1)+L2)+L3 of samples s0..s0+K-1 to the pixel sum in sample order (mod.rs:846).
// Queue entries are 4 x float4 (o|path, d|depth+code, T, L) in SoA arrays: coalesced 16-byte loads and stores.
#include <cub/cub.cuh>

#include "pt_launch.h"
#include "pt_scene_dev.cuh"
#include "pt_wavefront.h"

namespace ptb {

namespace {

constexpr int WF_THREADS = 256;
constexpr int WF_CHUNK = 512;  // rays a warp takes from the queue per global atomic
constexpr int WF_SSTACK = 12;  // traversal-stack entries per lane kept in shared memory (deeper entries go to local memory)

// (called by the lanes of `amask` only: they are all live)
__device__ __forceinline__ void store_loose_hit(const WfQueue &q, size_t j, const float4 *s_obj, V3 o, V3 d, unsigned amask) {
    Hit best;
    best.t = __int_as_float(0x7f800000); best.prio = PRIO_NONE; best.ref = REF_NONE;
    closest_hit_loose(s_obj, o, d, amask, best);
    q.hit_t[j] = best.t; q.hit_ref[j] = best.ref; q.hit_prio[j] = best.prio;
}

__global__ void __launch_bounds__(256) k_wf_generate(const DScene sc, int W, int H, unsigned npix, unsigned long long s0, unsigned K,
                                                     unsigned long long seed, WfQueue q) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);
    const unsigned long long n = (unsigned long long)npix * K;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const unsigned long long n_round = (n + 31ull) & ~31ull;
    for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
        const unsigned amask = __ballot_sync(0xffffffffu, p < n);
        if (p >= n) continue;
        const uint32_t pixel = (uint32_t)(p % npix);
        const unsigned long long s = s0 + p / npix;
        const int row = (int)(pixel / (uint32_t)W), px = (int)(pixel % (uint32_t)W);
        uint32_t rnd[4];
        philox4x32_10(pixel, (uint32_t)s, (uint32_t)(s >> 32), 0u, k0, k1, rnd);  // camera sample: event 0, slots 0,1
        const float ysub = (float)((s / 2) % 2), xsub = (float)(s % 2);
        const float r1 = 2.0f * u32_to_unit(rnd[0]);
        const float r2 = 2.0f * u32_to_unit(rnd[1]);
        V3 o, d;
        camera_ray(sc, W, H, px, H - 1 - row, xsub, ysub, tent(r1), tent(r2), o, d);
        q.o[p] = make_float4(o.x, o.y, o.z, __int_as_float((int)p));
        q.d[p] = make_float4(d.x, d.y, d.z, __int_as_float(0));
        q.T[p] = make_float4(1.f, 1.f, 1.f, 0.f);
        q.L[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        store_loose_hit(q, p, s_obj, o, d, amask);
    }
}

// closest hit of every queued segment
// (measured: 5 or 6 CTAs per SM with the spills that takes, or prefetching the leaf while a lane waits, are all slower or neutral)
// `order` (optional): the queue entries in the order they should be traced (k_wf_ray_keys + radix sort); results go back to
// the entry itself, so nothing downstream sees the order
__global__ void __launch_bounds__(WF_THREADS, 4) k_wf_trace(const DScene sc, const WfQueue q, const int *__restrict__ n_rays_ptr,
                                                             int *__restrict__ fetch_ptr, unsigned long long *__restrict__ counters,
                                                             const int wf_refill, const int wf_descend_min,
                                                             const int *__restrict__ order) {
    const int n = *n_rays_ptr;
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
    __shared__ int2 s_stack[WF_SSTACK * WF_THREADS];
    int2 l_stack[BVH_STACK - WF_SSTACK];
#define PTB_STK(i) (*((i) < WF_SSTACK ? &s_stack[(i) * WF_THREADS + threadIdx.x] : &l_stack[(i) - WF_SSTACK]))

    for (;;) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, busy);
        const int n_idle = 32 - __popc(busy_mask);
        if (!exhausted && (n_idle >= wf_refill || busy_mask == 0u)) {
            // the warp owns a private chunk [w_next, w_end) of the queue and only touches the global cursor when it runs dry:
            // one same-address atomic per WF_CHUNK rays instead of one per refill
            if (w_next >= w_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fetch_ptr, WF_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                w_next = base;
                w_end = min(base + WF_CHUNK, n);
                if (base >= n) exhausted = true;
            }
            const int idx = w_next + __popc(~busy_mask & lt_mask);
            const bool got = !busy && idx < w_end;
            w_next = min(w_next + n_idle, w_end);
            if (got) {
                const int r = order ? __ldcs(&order[idx]) : idx;
                const float4 qo = __ldcs(&q.o[r]), qd = __ldcs(&q.d[r]);  // streaming: keep the L2 for the BVH
                o = mk3(qo.x, qo.y, qo.z); d = mk3(qd.x, qd.y, qd.z);
                ray_idx = r;
                best.t = __ldcs(&q.hit_t[r]); best.ref = __ldcs(&q.hit_ref[r]); best.prio = __ldcs(&q.hit_prio[r]);
                id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
                cur = sc.bvh_root; sp = 0; gate_obj = -1;
                busy = true;
            }
        }
        if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        if (busy) {
            while (cur >= 0) {  // descend until this lane holds a leaf or is done
                PTB_BVH_NODE_STEP();
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
#undef PTB_STK
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
// Cooperative trace kernel: FOUR LANES PER RAY, eight rays per warp.
//
// Why: with one lane per ray every diverged lane pays one L1 data-pipe wavefront per 16 bytes it loads (a 128-byte node = 8),
// and ncu shows that pipe 90 % busy (profiles/r01d).  Here lane c of a sub-warp owns child c of the four-wide node: the four
// lanes fetch the node with two coalesced 64-byte accesses (2 wavefronts instead of 8), each runs ONE slab test, the entry
// distances are exchanged with three shuffles, every lane computes its child's rank, the nearest child becomes the next node and
// the other hit children are pushed far-to-near, in parallel, onto the sub-warp's stack in shared memory.  In a leaf, lane c
// tests primitive c (a leaf holds at most four) and the best candidate is found with two shuffle steps.  Only eight rays share
// an instruction stream, so far fewer lanes wait for the slowest ray of the warp.
// Arithmetic per primitive, prio tie-break and the lazy mesh gate are those of the one-lane-per-ray traversal: same hits.
//
// MEASURED (B200, synthetic scene, 1080p x 8 spp): bit-identical images, 2 wavefronts per node visit as designed, but 75 vs 175
// Mpaths/s: with eight instead of ~26 rays per warp and the same 32 warps per SM there are three times fewer independent rays
// in flight, and the traversal is latency bound.  Kept behind the option "wf_coop" (default off) as a tested experiment.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WF_CSTACK = 32;  // shared-memory stack entries per ray; deeper entries go to a global overflow area

__global__ void __launch_bounds__(WF_THREADS, 4) k_wf_trace_coop(const DScene sc, const WfQueue q, const int *__restrict__ n_rays_ptr,
                                                                  int *__restrict__ fetch_ptr, unsigned long long *__restrict__ counters,
                                                                  int2 *__restrict__ overflow, const int refill_groups) {
    __shared__ int2 s_stack[WF_CSTACK * (WF_THREADS / 4)];
    const int n = *n_rays_ptr;
    const int lane = threadIdx.x & 31, sub = lane & 3, gbase = lane & ~3;
    const unsigned gmask = 0xFu << gbase;
    const int group_in_block = threadIdx.x >> 2;
    int2 *const ovf = overflow + ((size_t)blockIdx.x * (WF_THREADS / 4) + group_in_block) * BVH_STACK;
#define PTB_CSTK(i) (*((i) < WF_CSTACK ? &s_stack[(i) * (WF_THREADS / 4) + group_in_block] : &ovf[(i) - WF_CSTACK]))
    const float inf = __int_as_float(0x7f800000);
    unsigned n_nodes = 0, n_prims = 0;

    bool busy = false, exhausted = false;  // busy / cur / sp / best are uniform within a sub-warp
    int w_next = 0, w_end = 0;
    int ray_idx = 0;
    V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f), id = mk3(1.f, 1.f, 1.f), ood = mk3(0.f, 0.f, 0.f);
    Hit best;
    best.t = 0.f; best.prio = PRIO_NONE; best.ref = REF_NONE;
    int cur = BVH_EMPTY_REF, sp = 0, gate_obj = -1;
    bool gate_pass = false;

#define PTB_CPOP()                                                          \
    do {                                                                    \
        cur = BVH_EMPTY_REF;                                                \
        while (sp > 0) {                                                    \
            --sp;                                                           \
            const int2 e_ = PTB_CSTK(sp);                                   \
            if (__int_as_float(e_.y) <= best.t) { cur = e_.x; break; }      \
        }                                                                   \
    } while (0)

    for (;;) {
        const unsigned busy_mask = __ballot_sync(0xffffffffu, busy);         // four equal bits per sub-warp
        const int n_idle = (32 - __popc(busy_mask)) >> 2;                    // idle sub-warps
        if (!exhausted && (n_idle >= refill_groups || busy_mask == 0u)) {
            if (w_next >= w_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fetch_ptr, WF_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                w_next = base;
                w_end = min(base + WF_CHUNK, n);
                if (base >= n) exhausted = true;
            }
            // rank of this sub-warp among the idle ones: idle sub-warps below it (count one bit per sub-warp)
            const unsigned idle_leaders = ~busy_mask & 0x11111111u;
            const int idx = w_next + __popc(idle_leaders & ((1u << gbase) - 1u));
            const bool got = !busy && idx < w_end;
            w_next = min(w_next + n_idle, w_end);
            if (got) {
                const float4 qo = __ldcs(&q.o[idx]), qd = __ldcs(&q.d[idx]);
                o = mk3(qo.x, qo.y, qo.z); d = mk3(qd.x, qd.y, qd.z);
                ray_idx = idx;
                best.t = __ldcs(&q.hit_t[idx]); best.ref = __ldcs(&q.hit_ref[idx]); best.prio = __ldcs(&q.hit_prio[idx]);
                id = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                ood = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
                cur = sc.bvh_root; sp = 0; gate_obj = -1;
                busy = true;
            }
        }
        if (__ballot_sync(0xffffffffu, busy) == 0u) break;
        if (busy) {
            while (cur >= 0) {  // ---- inner node: lane `sub` owns child `sub`
                const float4 *nd = sc.bvh_nodes + 8 * (size_t)cur;
                const float4 ca = __ldg(nd + sub), cb = __ldg(nd + 4 + sub);
                const int ref = __float_as_int(cb.z);
                float t_in;
                const bool hit = ref != BVH_EMPTY_REF && slab(ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, id, ood, best.t, t_in);
                const float t = hit ? t_in : inf;
                // rank of my child among the hit children (ties by child index)
                int rank = 0;
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    const float tj = __shfl_sync(gmask, t, gbase + ((sub + j) & 3));
                    const int cj = (sub + j) & 3;
                    rank += (tj < t || (tj == t && cj < sub)) ? 1 : 0;
                }
                const unsigned hits = (__ballot_sync(gmask, hit) >> gbase) & 0xFu;
                const int n_hit = __popc(hits);
                if (sub == 0) n_nodes++;
                if (n_hit == 0) {
                    PTB_CPOP();
                } else {
                    // the nearest child is followed, the others are postponed far-to-near (the nearest of them ends on top)
                    if (hit && rank > 0) PTB_CSTK(sp + (n_hit - 1 - rank)) = make_int2(ref, __float_as_int(t));
                    const unsigned first = __ballot_sync(gmask, hit && rank == 0);
                    cur = __shfl_sync(gmask, ref, __ffs(first) - 1);
                    sp += n_hit - 1;
                    __syncwarp(gmask);  // the pushes must be visible to the sub-warp's later pops
                }
                if ((__popc(__activemask()) >> 2) < 3) break;  // few rays still descending: let the leaf holders get on with it
            }
            if (cur < 0 && cur != BVH_EMPTY_REF) {  // ---- leaf: lane `sub` owns primitive `sub`
                const int code = ~cur;
                const int first = code >> 3, count = (code & 7) + 1;
                if (sub == 0) n_prims += count;
                float ct = inf;
                uint32_t cprio = PRIO_NONE;
                int cref = REF_NONE;
                for (int k0 = 0; k0 < count; k0 += 4) {  // leaves hold at most four primitives by construction; stay general
                    const int k = first + k0 + sub;
                    if (k0 + sub < count) {
                        const float4 A = __ldg(&sc.bvh_tri[2 * (size_t)k]), E1 = __ldg(&sc.bvh_tri[2 * (size_t)k + 1]), E2 = __ldg(&sc.bvh_e2[k]);
                        const bool is_sphere = __float_as_int(E1.w) < 0;
                        float tt;
                        if (is_sphere) tt = sphere_t(xyz(A), E1.x, o, d);
                        else tt = triangle_t(xyz(A), xyz(E1), xyz(E2), o, d);
                        const uint32_t prio = (uint32_t)__float_as_int(E2.w);
                        const bool beats_best = tt < best.t || (tt == best.t && prio < best.prio);
                        const bool beats_mine = tt < ct || (tt == ct && prio < cprio);
                        if (tt > 0.0f && beats_best && beats_mine) {
                            bool ok = true;
                            if (!is_sphere) {  // mesh gate (mod.rs:267-277), lazily, cached per lane and object
                                const int obj = __float_as_int(A.w);
                                if (obj != gate_obj) {
                                    const float4 g = __ldg(&sc.obj_gate[obj]);
                                    gate_pass = sphere_gate(xyz(g), g.w, o, d);
                                    gate_obj = obj;
                                }
                                ok = gate_pass;
                            }
                            if (ok) { ct = tt; cprio = prio; cref = REF_BVH_BIT | (is_sphere ? REF_SPHERE_BIT : 0) | k; }
                        }
                    }
                }
                // best candidate of the sub-warp (t, then prio), two butterfly steps
#pragma unroll
                for (int m = 1; m <= 2; m <<= 1) {
                    const float ot = __shfl_xor_sync(gmask, ct, m);
                    const uint32_t op = __shfl_xor_sync(gmask, cprio, m);
                    const int orf = __shfl_xor_sync(gmask, cref, m);
                    if (ot < ct || (ot == ct && op < cprio)) { ct = ot; cprio = op; cref = orf; }
                }
                if (cref != REF_NONE) { best.t = ct; best.prio = cprio; best.ref = cref; }
                PTB_CPOP();
            }
            if (cur == BVH_EMPTY_REF) {
                if (sub == 0) {
                    __stcs(&q.hit_t[ray_idx], best.t);
                    __stcs(&q.hit_ref[ray_idx], best.ref);
                }
                busy = false;
            }
        }
    }
#undef PTB_CPOP
#undef PTB_CSTK
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
__global__ void __launch_bounds__(256) k_wf_shade(const DScene sc, const WfQueue q, const int *__restrict__ n_rays_ptr, WfQueue nq,
                                                  int *__restrict__ n_next_ptr, float4 *__restrict__ slots, unsigned long long n_paths,
                                                  unsigned npix, unsigned long long s0, unsigned long long seed,
                                                  unsigned long long *__restrict__ segment_counter) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);
    const int n = *n_rays_ptr;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(segment_counter, (unsigned long long)n);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int stride = gridDim.x * blockDim.x;
    const int n_round = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        int n_out = 0;
        float4 c_o, c_d, c_T, c_L, k_o, k_d, k_T;  // continuation and (optional) transmitted child
        if (valid) {
            const float4 qo = __ldg(&q.o[i]), qd = __ldg(&q.d[i]), qT = __ldg(&q.T[i]), qL = __ldg(&q.L[i]);
            const int path = __float_as_int(qo.w), dc = __float_as_int(qd.w);
            const int depth = dc & 0xff, code = dc >> 8;
            const V3 o = mk3(qo.x, qo.y, qo.z), d = mk3(qd.x, qd.y, qd.z), T = mk3(qT.x, qT.y, qT.z);
            V3 L = mk3(qL.x, qL.y, qL.z);
            Hit h;
            h.t = q.hit_t[i]; h.ref = q.hit_ref[i]; h.prio = 0;
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
                    }
                }
            }
            if (!cont) slots[(size_t)code * n_paths + (size_t)path] = make_float4(L.x, L.y, L.z, 0.f);  // branch finished
        }
        // one atomic per warp for all appended entries
        const unsigned m1 = __ballot_sync(0xffffffffu, n_out >= 1), m2 = __ballot_sync(0xffffffffu, n_out == 2);
        const int total = __popc(m1) + __popc(m2);
        if (total) {
            int base = 0;
            if (lane == 0) base = atomicAdd(n_next_ptr, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (n_out >= 1) {
                const int j = base + __popc(m1 & lt_mask);
                nq.o[j] = c_o; nq.d[j] = c_d; nq.T[j] = c_T; nq.L[j] = c_L;
                store_loose_hit(nq, j, s_obj, mk3(c_o.x, c_o.y, c_o.z), mk3(c_d.x, c_d.y, c_d.z), m1);
            }
            if (n_out == 2) {
                const int j = base + __popc(m1) + __popc(m2 & lt_mask);
                nq.o[j] = k_o; nq.d[j] = k_d; nq.T[j] = k_T; nq.L[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                store_loose_hit(nq, j, s_obj, mk3(k_o.x, k_o.y, k_o.z), mk3(k_d.x, k_d.y, k_d.z), m2);
            }
        }
    }
}

// Ray reordering between bounces (option wf_sort, default off).  Sort key of a queued ray: rays that start close together and
// point into the same octant visit the same BVH nodes, so a warp tracing 32 of them could share its node fetches.
// 21-bit Morton code of the origin's cell in a 128^3 grid over the scene box + 3 bits of direction signs.  Only the ORDER in
// which rays are traced depends on it: hits, shading and the per-path sums do not (bit-identical images, tested).
// MEASURED (B200, profiles/r01j_ray_sort_experiment.md): synthetic 1.31 M triangles 4K: 179 -> 153 (octant-major) / 163 (cell-major)
// Mpaths/s; mesh.json 1080p: 707 -> 343 / 357.  Diffuse bounces inside one octant still diverge within a few levels of a deep BVH,
// so the traversal gains a few per cent while key + radix sort + length read-back + scattered ray fetch cost 0.7-1 ms per bounce.
__device__ __forceinline__ unsigned spread7(unsigned v) {  // bit i -> bit 3i (i < 10)
    v = (v ^ (v << 16)) & 0xff0000ffu;
    v = (v ^ (v << 8)) & 0x0300f00fu;
    v = (v ^ (v << 4)) & 0x030c30c3u;
    v = (v ^ (v << 2)) & 0x09249249u;
    return v;
}
__global__ void __launch_bounds__(256) k_wf_ray_keys(const DScene sc, const WfQueue q, int n, int mode, unsigned *__restrict__ keys,
                                                     int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = __ldg(&q.o[i]), d = __ldg(&q.d[i]);
    const float fx = fminf(fmaxf((o.x - sc.world_lo.x) * sc.world_inv.x, 0.f), 1.f);
    const float fy = fminf(fmaxf((o.y - sc.world_lo.y) * sc.world_inv.y, 0.f), 1.f);
    const float fz = fminf(fmaxf((o.z - sc.world_lo.z) * sc.world_inv.z, 0.f), 1.f);
    const unsigned cx = min((unsigned)(fx * 128.f), 127u), cy = min((unsigned)(fy * 128.f), 127u), cz = min((unsigned)(fz * 128.f), 127u);
    const unsigned cell = spread7(cx) | (spread7(cy) << 1) | (spread7(cz) << 2);
    const unsigned oct = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
    keys[i] = mode == 2 ? ((cell << 3) | oct) : ((oct << 21) | cell);
    idx[i] = i;
}

// radiance_v += radiance(sample) for the K samples of the batch, in sample order (mod.rs:846)
__global__ void __launch_bounds__(256) k_wf_accumulate(const float4 *__restrict__ slots, unsigned long long n_paths, unsigned npix,
                                                       unsigned K, float *__restrict__ sum_rgb, const int fb_zero) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += stride) {
        float *fb = sum_rgb + 3ull * pixel;
        V3 acc = fb_zero ? mk3(0.f, 0.f, 0.f) : mk3(fb[0], fb[1], fb[2]);
        for (unsigned k = 0; k < K; ++k) {
            const size_t p = (size_t)k * npix + pixel;
            const float4 a = slots[p], b = slots[n_paths + p], c = slots[2 * n_paths + p], e = slots[3 * n_paths + p];
            const V3 L = ((mk3(a.x, a.y, a.z) + mk3(b.x, b.y, b.z)) + mk3(c.x, c.y, c.z)) + mk3(e.x, e.y, e.z);
            acc = acc + L;
        }
        fb[0] = acc.x; fb[1] = acc.y; fb[2] = acc.z;
    }
}

size_t loose_smem(const DScene &sc) { return loose_smem_bytes(sc); }

}  // namespace

void wf_release(WfWorkspace &w) {
    for (int b = 0; b < 2; ++b) {
        if (w.q[b].o) cudaFree(w.q[b].o);
        if (w.q[b].d) cudaFree(w.q[b].d);
        if (w.q[b].T) cudaFree(w.q[b].T);
        if (w.q[b].L) cudaFree(w.q[b].L);
        if (w.q[b].hit_t) cudaFree(w.q[b].hit_t);
        if (w.q[b].hit_ref) cudaFree(w.q[b].hit_ref);
        if (w.q[b].hit_prio) cudaFree(w.q[b].hit_prio);
    }
    if (w.slots) cudaFree(w.slots);
    if (w.counters) cudaFree(w.counters);
    if (w.overflow) cudaFree(w.overflow);
    for (int b = 0; b < 2; ++b) {
        if (w.sort_keys[b]) cudaFree(w.sort_keys[b]);
        if (w.sort_idx[b]) cudaFree(w.sort_idx[b]);
    }
    if (w.sort_tmp) cudaFree(w.sort_tmp);
    w = WfWorkspace{};
}

static cudaError_t wf_reserve(WfWorkspace &w, size_t n_paths) {
    if (n_paths <= w.cap_paths) return cudaSuccess;
    wf_release(w);
    const size_t cap = 4 * n_paths;  // every path can split twice (mod.rs:775-786): at most 4 live branches
    cudaError_t e;
    for (int b = 0; b < 2; ++b) {
        if ((e = cudaMalloc((void **)&w.q[b].o, cap * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].d, cap * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].T, cap * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].L, cap * sizeof(float4))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].hit_t, cap * sizeof(float))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].hit_ref, cap * sizeof(int))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&w.q[b].hit_prio, cap * sizeof(unsigned))) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc((void **)&w.slots, 4 * n_paths * sizeof(float4))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&w.counters, 2 * (WF_MAX_BOUNCES + 2) * sizeof(int))) != cudaSuccess) return e;
    w.cap_paths = n_paths;
    return cudaSuccess;
}

// renders samples [a.spp_begin, a.spp_begin + a.spp_count) of every pixel into a.sum_rgb; returns the number of kernels launched
cudaError_t wavefront_render(const DScene &sc, const RenderArgs &a, WfWorkspace &w, int sm_count, size_t target_paths, int refill,
                             int descend_min, int coop, int sort_mode, cudaStream_t st,
                             unsigned *launches) {
    const unsigned npix = (unsigned)a.width * (unsigned)a.height;
    unsigned K = (unsigned)std::max<size_t>(1, target_paths / npix);
    if ((unsigned long long)K > a.spp_count) K = (unsigned)a.spp_count;
    if (K == 0) return cudaSuccess;
    const size_t n_paths = (size_t)npix * K;
    if (n_paths > (1ull << 29)) return cudaErrorInvalidValue;  // queue indices are 32-bit ints, 4 branches per path
    cudaError_t e = wf_reserve(w, n_paths);
    if (e != cudaSuccess) return e;
    const size_t smem = loose_smem(sc);
    if ((e = cudaFuncSetAttribute(k_wf_generate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_wf_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    int trace_per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&trace_per_sm, k_wf_trace, WF_THREADS, 0)) != cudaSuccess) return e;
    if (trace_per_sm < 1) return cudaErrorLaunchOutOfResources;
    int coop_per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&coop_per_sm, k_wf_trace_coop, WF_THREADS, 0)) != cudaSuccess) return e;
    if (coop_per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int trace_blocks = sm_count * trace_per_sm, coop_blocks = sm_count * coop_per_sm, wide_blocks = sm_count * 8;
    if (coop) {
        const size_t need = (size_t)coop_blocks * (WF_THREADS / 4) * BVH_STACK;
        if (w.cap_overflow < need) {
            if (w.overflow) cudaFree(w.overflow);
            w.overflow = nullptr; w.cap_overflow = 0;
            if ((e = cudaMalloc((void **)&w.overflow, need * sizeof(int2))) != cudaSuccess) return e;
            w.cap_overflow = need;
        }
    }

    // ray reordering only pays where rays walk a BVH, and the cooperative kernel has its own fetch logic
    const bool sorting = sort_mode != 0 && !coop && sc.bvh_root != BVH_EMPTY_REF;
    if (sorting && w.cap_sort < 4 * n_paths) {
        for (int b = 0; b < 2; ++b) {
            if (w.sort_keys[b]) cudaFree(w.sort_keys[b]);
            if (w.sort_idx[b]) cudaFree(w.sort_idx[b]);
            w.sort_keys[b] = nullptr; w.sort_idx[b] = nullptr;
        }
        if (w.sort_tmp) cudaFree(w.sort_tmp);
        w.sort_tmp = nullptr; w.cap_sort = 0;
        const size_t cap = 4 * n_paths;
        for (int b = 0; b < 2; ++b) {
            if ((e = cudaMalloc((void **)&w.sort_keys[b], cap * sizeof(unsigned))) != cudaSuccess) return e;
            if ((e = cudaMalloc((void **)&w.sort_idx[b], cap * sizeof(int))) != cudaSuccess) return e;
        }
        w.sort_tmp_bytes = 0;
        if ((e = cub::DeviceRadixSort::SortPairs(nullptr, w.sort_tmp_bytes, w.sort_keys[0], w.sort_keys[1], w.sort_idx[0], w.sort_idx[1],
                                                 (long long)cap, 0, 24, st)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&w.sort_tmp, std::max<size_t>(w.sort_tmp_bytes, 16))) != cudaSuccess) return e;
        w.cap_sort = cap;
    }

    unsigned long long done = 0;
    while (done < a.spp_count) {
        const unsigned k_now = (unsigned)std::min<unsigned long long>(K, a.spp_count - done);
        const size_t paths_now = (size_t)npix * k_now;
        const unsigned long long s0 = a.spp_begin + done;
        // counters[2*b] = entries of bounce b's queue, counters[2*b+1] = its fetch cursor
        if ((e = cudaMemsetAsync(w.counters, 0, 2 * (WF_MAX_BOUNCES + 2) * sizeof(int), st)) != cudaSuccess) return e;
        const int first = (int)paths_now;
        if ((e = cudaMemcpyAsync(w.counters, &first, sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(w.slots, 0, 4 * n_paths * sizeof(float4), st)) != cudaSuccess) return e;
        k_wf_generate<<<wide_blocks, 256, smem, st>>>(sc, a.width, a.height, npix, s0, k_now, a.seed, w.q[0]);
        (*launches)++;
        const int *order = nullptr;  // bounce 0: camera rays, generated in pixel order, are coherent as they are
        for (int b = 0; b < WF_MAX_BOUNCES; ++b) {
            const WfQueue &cur = w.q[b & 1], &nxt = w.q[(b + 1) & 1];
            if (coop)
                k_wf_trace_coop<<<coop_blocks, WF_THREADS, 0, st>>>(sc, cur, w.counters + 2 * b, w.counters + 2 * b + 1, a.segment_counter,
                                                                     w.overflow, 2);
            else
                k_wf_trace<<<trace_blocks, WF_THREADS, 0, st>>>(sc, cur, w.counters + 2 * b, w.counters + 2 * b + 1, a.segment_counter, refill,
                                                                descend_min, order);
            k_wf_shade<<<wide_blocks, 256, smem, st>>>(sc, cur, w.counters + 2 * b, nxt, w.counters + 2 * (b + 1), w.slots,
                                                       n_paths, npix, s0, a.seed, a.segment_counter);
            *launches += 2;
            if (sorting && b + 1 < WF_MAX_BOUNCES) {
                int n_next = 0;  // the sort needs the queue length on the host: one 4-byte read-back per bounce
                if ((e = cudaMemcpyAsync(&n_next, w.counters + 2 * (b + 1), sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
                if (n_next <= 0) break;  // every branch has ended: the remaining bounces would be empty launches
                k_wf_ray_keys<<<(n_next + 255) / 256, 256, 0, st>>>(sc, nxt, n_next, sort_mode, w.sort_keys[0], w.sort_idx[0]);
                (*launches)++;
                size_t tmp = w.sort_tmp_bytes;
                if ((e = cub::DeviceRadixSort::SortPairs(w.sort_tmp, tmp, w.sort_keys[0], w.sort_keys[1], w.sort_idx[0], w.sort_idx[1],
                                                         (long long)n_next, 0, 24, st)) != cudaSuccess) return e;
                order = w.sort_idx[1];
            }
        }
        k_wf_accumulate<<<wide_blocks, 256, 0, st>>>(w.slots, n_paths, npix, k_now, a.sum_rgb, (a.fb_zero && done == 0) ? 1 : 0);
        (*launches)++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        done += k_now;
    }
    return cudaSuccess;
}

}  // namespace ptb
