// pt_kernels.cu -- sm_100a kernels of the radiance loop: closest hit, integrator megakernel, resolve.
//
// Replaces (reference, src/render/mod.rs): the rayon pixel loop :1001-1024, render_pixel :794-857,
// radiance :661-792, intersect_scene :631-659, intersect_sphere :412-438, Triangle::intersect :554-615.
//
// Integrator design: persistent warps pull 8x4 pixel tiles from a global counter; one lane owns one
// pixel and walks its samples in order (so the per-pixel fp32 sum has the reference's sequential
// order and the image is deterministic).  A lane that finishes a path regenerates the next sample of
// its pixel immediately, so lanes stay busy until the pixel's sample budget is used up.  The
// "loose" part of the scene (few, large primitives: spheres, wall quads) is staged in shared memory
// and scanned by all lanes in lock-step (broadcast LDS.128, no divergence); big meshes live in a
// BVH (pt_bvh.cuh).  No tensor cores: there is no dense contraction in this workload.
#include <algorithm>

#include "pt_launch.h"
#include "pt_scene_dev.cuh"

namespace ptb {

// ---------------------------------------------------------------------------------------------
// parity hooks: arbitrary rays / deterministic primary rays
// ---------------------------------------------------------------------------------------------
template <bool HAS_BVH>
__global__ void __launch_bounds__(256) k_intersect(const DScene sc, const float *__restrict__ rays, unsigned long long n,
                                                   int primary_w, int primary_h, int *__restrict__ obj_out,
                                                   int *__restrict__ tri_out, float *__restrict__ t_out,
                                                   float *__restrict__ point_out, float *__restrict__ normal_out) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_round = (n + 31ull) & ~31ull;  // whole warps stay converged for the votes
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f);
        if (valid) {
            if (rays) {
                o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
                d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
            } else {  // centre ray of pixel i: xsub = ysub = xfilter = yfilter = 0 (mod.rs:805-843)
                const int row = (int)(i / (unsigned)primary_w), x = (int)(i % (unsigned)primary_w);
                camera_ray(sc, primary_w, primary_h, x, primary_h - 1 - row, 0.f, 0.f, 0.f, 0.f, o, d);
            }
        }
        // caller rays need not be normalised (the reference's intersect_scene takes any direction): the gate shortcut that
        // assumes |d| = 1 is switched off for the lanes whose direction is not unit length to 1e-3
        const bool unit_dir = fabsf(dot(d, d) - 1.0f) <= 1e-3f;
        const Hit h = closest_hit<HAS_BVH>(sc, s_obj, o, d, 0xffffffffu, valid, unit_dir);
        if (valid) {
            int obj = -1, tri = -1;
            V3 x = mk3(0.f, 0.f, 0.f), nn = mk3(0.f, 0.f, 0.f);
            float t = 0.f;
            if (h.ref != REF_NONE) { finish_hit(sc, s_obj, s_tri, h, o, d, obj, tri, x, nn); t = h.t; }
            obj_out[i] = obj;
            if (tri_out) tri_out[i] = tri;
            if (t_out) t_out[i] = t;
            if (point_out) { point_out[3 * i] = x.x; point_out[3 * i + 1] = x.y; point_out[3 * i + 2] = x.z; }
            if (normal_out) { normal_out[3 * i] = nn.x; normal_out[3 * i + 1] = nn.y; normal_out[3 * i + 2] = nn.z; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the integrator megakernel
// ---------------------------------------------------------------------------------------------
struct PathStackEntry { V3 o, d, T; int depth, code; };

// SAMPLE_PARALLEL (frames with too few pixels to fill the GPU, e.g. the reference's default 450x300): a work slot is a chunk of
// a.sp_chunk consecutive samples of one pixel instead of the whole pixel, so every lane of every SM has work; a finished sample's
// radiance goes to a.sample_L[sample - spp_begin][pixel] and k_accumulate_samples adds them to the pixel in sample order
// afterwards -- the same sequential fp32 sum (mod.rs:846), hence the same bits.
template <bool HAS_BVH, bool SAMPLE_PARALLEL>
__global__ void __launch_bounds__(RENDER_THREADS, RENDER_MIN_BLOCKS) k_render(const DScene sc, const RenderArgs a) {
    extern __shared__ float4 smem[];
    const float4 *s_obj, *s_tri;
    stage_loose(sc, smem, s_obj, s_tri);

    const int lane = threadIdx.x & 31;
    const int W = a.width, H = a.height;
    unsigned long long n_segments = 0;

    // Work distribution: pixels are enumerated tile-major (slot = tile * 32 + position inside the 8x4 tile) and handed out
    // lane by lane from one global counter.  A lane that has used up its pixel's sample budget stores the sum and takes the next
    // slot at once, so no lane waits for the slowest pixel of a tile.  One lane still owns one pixel at a time and walks its
    // samples in order: the per-pixel fp32 sum keeps the reference's order (mod.rs:846).
    const unsigned n_pixel_slots = (unsigned)a.n_tiles * 32u;
    const unsigned n_slots = SAMPLE_PARALLEL ? n_pixel_slots * (unsigned)a.sp_n_chunks : n_pixel_slots;  // (< 2^31, checked by the host)
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long s_last = a.spp_begin + a.spp_count;
    unsigned long long s_end = s_last;  // SAMPLE_PARALLEL: end of the lane's current chunk
    bool have_pixel = false, retired = false;
    uint32_t pixel = 0;
    int px = 0, y = 0;
    float *fb = a.sum_rgb;
    V3 acc = mk3(0.f, 0.f, 0.f);
    // Camera rays are generated in batches: every lane keeps one spare primary direction, a lane that finishes a path
    // just swaps it in, and the (long, otherwise badly diverged) ray-generation code runs only when REGEN_BATCH lanes
    // need a new spare or some lane would have to idle.  Per-lane order of samples and of `acc += L` is unchanged.
    unsigned long long s_next = a.spp_begin, s = a.spp_begin;
    bool has_path = false, spare_ok = false;
    V3 spare_d = mk3(0.f, 0.f, 1.f);
    V3 o = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 1.f), T = mk3(1.f, 1.f, 1.f), L = mk3(0.f, 0.f, 0.f);
    int depth = 0, sp = 0, code = 0;      // code: which refraction splits were left through the transmitted child
    unsigned chain_mask = 1u;             // branches of the path tree that exist for this sample
    uint32_t nseg = 0;
    PathStackEntry stk[2];
    V3 Lc[4];                             // finished branches' emission sums; radiance = ((L0+L1)+L2)+L3

    for (;;) {
        // a pixel whose samples are all done (no path in flight, no spare ray, budget used up) is written back
        if (have_pixel && !has_path && !spare_ok && s_next >= s_end) {
            if (!SAMPLE_PARALLEL) { fb[0] = acc.x; fb[1] = acc.y; fb[2] = acc.z; }
            have_pixel = false;
        }
        const unsigned want_mask = __ballot_sync(0xffffffffu, !have_pixel && !retired);
        if (want_mask) {
            const int leader = __ffs(want_mask) - 1;
            unsigned base = 0;
            if (lane == leader) base = (unsigned)atomicAdd(a.tile_counter, __popc(want_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!have_pixel && !retired) {
                const unsigned slot_all = base + __popc(want_mask & lt_mask);
                if (slot_all >= n_slots) retired = true;
                else {
                    // chunk-major: the 32 lanes of a warp take the same chunk of 32 neighbouring pixels
                    const unsigned chunk = SAMPLE_PARALLEL ? slot_all / n_pixel_slots : 0u;
                    const unsigned slot = SAMPLE_PARALLEL ? slot_all % n_pixel_slots : slot_all;
                    const int tile = (int)(slot >> 5), pos = (int)(slot & 31u);
                    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
                    px = tx * TILE_W + (pos & (TILE_W - 1));
                    const int row = ty * TILE_H + (pos / TILE_W);
                    if (px < W && row < H && a.spp_count > 0) {  // (off-image slots of edge tiles are simply skipped)
                        pixel = (uint32_t)row * (uint32_t)W + (uint32_t)px;
                        y = H - 1 - row;  // mod.rs:805
                        if (SAMPLE_PARALLEL) {
                            s_next = a.spp_begin + (unsigned long long)chunk * (unsigned)a.sp_chunk;
                            s_end = s_next + (unsigned)a.sp_chunk < s_last ? s_next + (unsigned)a.sp_chunk : s_last;
                        } else {
                            fb = a.sum_rgb + 3ull * pixel;
                            acc = a.fb_zero ? mk3(0.f, 0.f, 0.f) : mk3(fb[0], fb[1], fb[2]);
                            s_next = a.spp_begin;
                        }
                        have_pixel = true;
                    }
                }
            }
        }
        const bool need = have_pixel && !spare_ok && s_next < s_end;
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        const bool starving = __any_sync(0xffffffffu, need && !has_path);
        if (starving || __popc(need_mask) >= a.regen_batch) {
            if (need) {  // camera sample: event 0, slots 0,1 (mod.rs:814-843)
                uint32_t rnd[4];
                philox4x32_10(pixel, (uint32_t)s_next, (uint32_t)(s_next >> 32), 0u, a.rk, rnd);
                const float ysub = (float)((s_next / 2) % 2), xsub = (float)(s_next % 2);
                const float r1 = 2.0f * u32_to_unit(rnd[0]);
                const float r2 = 2.0f * u32_to_unit(rnd[1]);
                V3 o_unused;
                camera_ray(sc, W, H, px, y, xsub, ysub, tent(r1), tent(r2), o_unused, spare_d);
                spare_ok = true;
                s_next++;
            }
        }
        if (!has_path && spare_ok) {  // start sample s = s_next - 1
            o = sc.lens_center; d = spare_d;
            T = mk3(1.f, 1.f, 1.f); L = mk3(0.f, 0.f, 0.f);
            depth = 0; sp = 0; code = 0; chain_mask = 1u;
            s = s_next - 1;
            spare_ok = false;
            has_path = true;
        }
        const unsigned amask = __ballot_sync(0xffffffffu, has_path);
        if (amask == 0u) {
            if (__ballot_sync(0xffffffffu, have_pixel || !retired) == 0u) break;  // every lane is out of pixels
            continue;
        }
        // ---- one radiance() call (mod.rs:662): event (code<<4 | new_depth); slots 0 = RR, 1,2 = diffuse, 3 = refraction.
        // The whole warp walks the object stream (lanes without a path are passengers: full-mask votes, no divergence).
        const Hit h = closest_hit<HAS_BVH>(sc, s_obj, o, d, 0xffffffffu, has_path);
        if (has_path) {
            nseg++;
            bool cont = false;
            if (h.ref != REF_NONE) {
                // the event's random numbers are only needed once something is hit (an open scene's rays mostly leave: the
                // sphere scenes spend 13 % of their instructions here otherwise)
                uint32_t rnd[4];
                philox4x32_10(pixel, (uint32_t)s, (uint32_t)(s >> 32), ((uint32_t)code << 4) | (uint32_t)(depth + 1), a.rk, rnd);
                int obj, tri;
                V3 x, n;
                finish_hit(sc, s_obj, s_tri, h, o, d, obj, tri, x, n);
                const int new_depth = depth + 1;
                ShadeOut so;
                shade_hit(sc, obj, n, d, T, new_depth, rnd, so);
                if (so.emits) L = L + so.emit;
                if (so.cont) {
                    if (so.split) {  // the reflected child continues this branch, the transmitted one is postponed
                        const int child = code | (1 << (new_depth - 1));
                        stk[sp].o = x; stk[sp].d = so.child_d; stk[sp].T = so.child_T; stk[sp].depth = new_depth;
                        stk[sp].code = child;
                        chain_mask |= 1u << child;
                        sp++;
                    }
                    o = x; d = so.d; T = so.T;
                    depth = new_depth;
                    cont = true;
                }
            }
            if (!cont) {
                if (sp > 0) {  // this branch is finished, continue with a postponed transmitted child
                    Lc[code] = L;
                    sp--;
                    o = stk[sp].o; d = stk[sp].d; T = stk[sp].T; depth = stk[sp].depth; code = stk[sp].code;
                    L = mk3(0.f, 0.f, 0.f);
                } else {  // sample finished: radiance_v += radiance (mod.rs:846)
                    if (chain_mask != 1u) {
                        Lc[code] = L;
                        const V3 z = mk3(0.f, 0.f, 0.f);
                        L = ((Lc[0] + ((chain_mask & 2u) ? Lc[1] : z)) + ((chain_mask & 4u) ? Lc[2] : z)) + ((chain_mask & 8u) ? Lc[3] : z);
                    }
                    if (SAMPLE_PARALLEL)
                        a.sample_L[(size_t)(s - a.spp_begin) * ((size_t)W * (size_t)H) + pixel] = make_float4(L.x, L.y, L.z, 0.f);
                    else acc = acc + L;
                    has_path = false;
                }
            }
        }
    }
    n_segments += nseg;
    // one atomic per warp
    for (int off = 16; off > 0; off >>= 1) n_segments += __shfl_down_sync(0xffffffffu, n_segments, off);
    if (lane == 0 && n_segments) atomicAdd(a.segment_counter, n_segments);
}

// radiance_v += radiance(sample) in sample order (mod.rs:846) for the K samples a SAMPLE_PARALLEL launch left in sample_L
__global__ void __launch_bounds__(256) k_accumulate_samples(const float4 *__restrict__ sample_L, unsigned npix, unsigned K,
                                                            float *__restrict__ sum_rgb, const int fb_zero) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += stride) {
        float *fb = sum_rgb + 3ull * pixel;
        V3 acc = fb_zero ? mk3(0.f, 0.f, 0.f) : mk3(fb[0], fb[1], fb[2]);
        for (unsigned k = 0; k < K; ++k) {
            const float4 v = __ldcs(&sample_L[(size_t)k * npix + pixel]);
            acc = acc + mk3(v.x, v.y, v.z);
        }
        fb[0] = acc.x; fb[1] = acc.y; fb[2] = acc.z;
    }
}

// radiance / spp, clamp to [0,1] (mod.rs:849-856)
__global__ void k_resolve(const float *__restrict__ sum, unsigned long long n, float spp, float *__restrict__ mean) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = PTB_DIV(sum[i], spp);
        mean[i] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

// Multi-GPU: sum the ranks' partial framebuffers (peer memory, NVLink loads), resolve, store (possibly into a peer's buffer).
// The order of the adds is the rank order, whatever the interconnect does: ((fb0 + fb1) + fb2) + ...
// spp = 0: store the raw sum (PTB_OUT_SUM of a multi-GPU context) instead of the clamped mean.
__device__ __forceinline__ float resolve1(float s, float spp) { return spp > 0.0f ? fminf(fmaxf(PTB_DIV(s, spp), 0.0f), 1.0f) : s; }
__global__ void __launch_bounds__(256) k_peer_reduce_resolve(const PeerPtrs peers, int n_peers, unsigned long long first,
                                                             unsigned long long n, float spp, float *__restrict__ dst) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((first & 3ull) == 0) {  // 16-byte path
        const unsigned long long n4 = n / 4;
        for (unsigned long long i = tid; i < n4; i += stride) {
            const unsigned long long at = first / 4 + i;
            float4 s = reinterpret_cast<const float4 *>(peers.p[0])[at];
            for (int g = 1; g < n_peers; ++g) {
                const float4 v = reinterpret_cast<const float4 *>(peers.p[g])[at];
                s.x = s.x + v.x; s.y = s.y + v.y; s.z = s.z + v.z; s.w = s.w + v.w;
            }
            float4 r;
            r.x = resolve1(s.x, spp); r.y = resolve1(s.y, spp); r.z = resolve1(s.z, spp); r.w = resolve1(s.w, spp);
            reinterpret_cast<float4 *>(dst)[at] = r;
        }
        for (unsigned long long i = n4 * 4 + tid; i < n; i += stride) {
            float s = peers.p[0][first + i];
            for (int g = 1; g < n_peers; ++g) s = s + peers.p[g][first + i];
            dst[first + i] = resolve1(s, spp);
        }
    } else {
        for (unsigned long long i = tid; i < n; i += stride) {
            float s = peers.p[0][first + i];
            for (int g = 1; g < n_peers; ++g) s = s + peers.p[g][first + i];
            dst[first + i] = resolve1(s, spp);
        }
    }
}

// exhaustive check of rcp_rn_normal against __frcp_rn over every float whose exponent field is in [1, 252]
__global__ void k_rcp_selftest(unsigned long long *mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const uint32_t bits = (uint32_t)i, e = (bits >> 23) & 0xffu;
        if (e < 1u || e > 252u) continue;
        const float x = __uint_as_float(bits);
        if (__float_as_uint(rcp_rn_normal(x)) != __float_as_uint(__frcp_rn(x))) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// a*b+c with operands chosen so that a fused multiply-add gives a different answer
__global__ void k_contraction_probe(float a, float b, float c, float *out) {
    out[0] = a * b + c;                                             // scalar multiply then add
    const float2 p = __fmul2_rn(make_float2(a, a), make_float2(b, b));  // packed multiply (FMUL2) then scalar adds
    out[1] = p.x + c;
    out[2] = p.y + c;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------

cudaError_t launch_rcp_selftest(unsigned long long *d_mismatches, int sm_count, cudaStream_t st) {
    k_rcp_selftest<<<sm_count * 8, 256, 0, st>>>(d_mismatches);
    return cudaGetLastError();
}

cudaError_t launch_contraction_probe(float a, float b, float c, float *d_out, cudaStream_t st) {
    k_contraction_probe<<<1, 1, 0, st>>>(a, b, c, d_out);
    return cudaGetLastError();
}

cudaError_t launch_intersect(const DScene &sc, const float *d_rays, unsigned long long n, int pw, int ph, int *d_obj, int *d_tri,
                             float *d_t, float *d_point, float *d_normal, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t smem = loose_smem_bytes(sc);
    const bool bvh = sc.bvh_root != BVH_EMPTY_REF;
    auto kern = bvh ? k_intersect<true> : k_intersect<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    unsigned long long blocks = (n + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, 256, smem, st>>>(sc, d_rays, n, pw, ph, d_obj, d_tri, d_t, d_point, d_normal);
    return cudaGetLastError();
}

cudaError_t launch_render(const DScene &sc, const RenderArgs &a, int sm_count, cudaStream_t st) {
    const size_t smem = loose_smem_bytes(sc);
    const bool bvh = sc.bvh_root != BVH_EMPTY_REF;
    const bool sp = a.sample_L != nullptr;
    auto kern = sp ? (bvh ? k_render<true, true> : k_render<false, true>) : (bvh ? k_render<true, false> : k_render<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RENDER_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    // persistent grid: a whole number of resident CTAs per SM (148 SMs on B200)
    long long blocks = (long long)sm_count * per_sm;
    const long long need = ((long long)a.n_tiles * (sp ? a.sp_n_chunks : 1) + RENDER_THREADS / 32 - 1) / (RENDER_THREADS / 32);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, RENDER_THREADS, smem, st>>>(sc, a);
    if (sp) {  // the K = spp_count samples of every pixel, added in sample order
        const unsigned npix = (unsigned)a.width * (unsigned)a.height;
        const unsigned ablocks = (unsigned)std::max<long long>(1, std::min<long long>((npix + 255) / 256, (long long)sm_count * 8));
        k_accumulate_samples<<<ablocks, 256, 0, st>>>(a.sample_L, npix, (unsigned)a.spp_count, a.sum_rgb, a.fb_zero);
    }
    return cudaGetLastError();
}

cudaError_t launch_peer_reduce_resolve(const PeerPtrs &peers, int n_peers, unsigned long long first, unsigned long long n,
                                       unsigned long long spp, float *d_dst, int sm_count, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n / 4 + 255) / 256 + 1;
    const unsigned long long cap = (unsigned long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_peer_reduce_resolve<<<(unsigned)blocks, 256, 0, st>>>(peers, n_peers, first, n, (float)spp, d_dst);
    return cudaGetLastError();
}

cudaError_t launch_resolve(const float *d_sum, unsigned long long n, unsigned long long spp, float *d_mean, int sm_count,
                           cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_resolve<<<(unsigned)blocks, 256, 0, st>>>(d_sum, n, (float)spp, d_mean);
    return cudaGetLastError();
}

}  // namespace ptb
