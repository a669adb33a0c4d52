// pt_bvh_build.cu -- LBVH construction on the device (Karras 2012): Morton codes -> radix sort (CUB) -> hierarchy ->
// bottom-up refit -> collapse small subtrees into multi-primitive leaves -> collapse pairs of levels into 128-byte
// four-child nodes.
//
// New with respect to the reference (which is brute force, src/render/mod.rs:631-659 / :554-615).  The hierarchy only
// selects which primitives are tested; the tests themselves are the reference's arithmetic (pt_device.cuh), and the
// boxes are padded so no hit the reference's fp32 test would accept can be culled (triangle_pad / sphere_extent below).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "pt_bvh.cuh"
#include "pt_bvh8.h"
#include "pt_bvh_build.h"
#include "pt_launch.h"

namespace ptb {

namespace {

constexpr float UNIT_ROUNDOFF = 5.9604645e-8f;  // 2^-24

inline float4 f4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline float ibits(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float ubits(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline V3 v3(const float *p) { return mk3(p[0], p[1], p[2]); }

// How far (in world units) the point o + d*t_c of a hit ACCEPTED by the reference's fp32 Moeller-Trumbore test
// (mod.rs:560-593) can lie from the true triangle.  First-order forward error analysis with unit roundoff u:
//   |d det| <= 7u|e1||e2|,  |d(tv.p)| <= 8u|tv||e2|,  |d(d.q)| <= 8u|tv||e1|,  |d(e2.q)| <= 8u|tv||e1||e2|,
// the accepted |det| is >= 1e-4 (mod.rs:571), |tv| and t are bounded by D (distance bound between any ray origin and
// the scene), so  |du||e1| + |dv||e2| + |dt|  <=  u |e1||e2| (31 D + 7 L) / 1e-4 + 2u (L + D),  L = |e1| + |e2|.
// A safety factor 2 covers the second-order terms; the last term covers the slab test's own fp32 rounding.
inline float triangle_pad(V3 e1, V3 e2, float D, float coord_max) {
    const double l1 = std::sqrt((double)e1.x * e1.x + (double)e1.y * e1.y + (double)e1.z * e1.z);
    const double l2 = std::sqrt((double)e2.x * e2.x + (double)e2.y * e2.y + (double)e2.z * e2.z);
    const double u = UNIT_ROUNDOFF, L = l1 + l2;
    const double grazing = u * l1 * l2 * (31.0 * D + 7.0 * L) / 1e-4;
    const double pad = 2.0 * (grazing + 2.0 * u * (L + D)) + 16.0 * u * (coord_max + D);
    return (float)pad;
}
// Sphere (mod.rs:412-438): det = b^2 - |op|^2 + r^2 is accepted when >= 0; |d det| <= 12u|op|^2 + u r^2, and the
// accepted point satisfies |P - c|^2 = r^2 + d det, so half extent = sqrt(r^2 + 2 * 13u D^2) (+ slab rounding).
inline float sphere_extent(float r, float D, float coord_max) {
    const double u = UNIT_ROUNDOFF;
    return (float)(std::sqrt((double)r * r + 2.0 * 13.0 * u * ((double)D * D + (double)r * r)) + 16.0 * u * (coord_max + D));
}

// ---------------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_prim_boxes(const float4 *__restrict__ recs, const float *__restrict__ pads, int n, float4 *__restrict__ blo,
                             float4 *__restrict__ bhi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 A = recs[3 * i], E1 = recs[3 * i + 1], E2 = recs[3 * i + 2];
    const float p = pads[i];
    float3 lo, hi;
    if (__float_as_int(E1.w) < 0) {  // sphere: pads[] holds the half extent
        lo = make_float3(A.x - p, A.y - p, A.z - p);
        hi = make_float3(A.x + p, A.y + p, A.z + p);
    } else {
        const float bx = A.x + E1.x, by = A.y + E1.y, bz = A.z + E1.z;
        const float cx = A.x + E2.x, cy = A.y + E2.y, cz = A.z + E2.z;
        lo = make_float3(fminf(A.x, fminf(bx, cx)) - p, fminf(A.y, fminf(by, cy)) - p, fminf(A.z, fminf(bz, cz)) - p);
        hi = make_float3(fmaxf(A.x, fmaxf(bx, cx)) + p, fmaxf(A.y, fmaxf(by, cy)) + p, fmaxf(A.z, fmaxf(bz, cz)) + p);
    }
    blo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    bhi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n, float3 cmin, float3 cscale,
                         unsigned long long *__restrict__ keys, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float cx = 0.5f * (blo[i].x + bhi[i].x), cy = 0.5f * (blo[i].y + bhi[i].y), cz = 0.5f * (blo[i].z + bhi[i].z);
    const float fx = fminf(fmaxf((cx - cmin.x) * cscale.x, 0.f), 2097151.f);
    const float fy = fminf(fmaxf((cy - cmin.y) * cscale.y, 0.f), 2097151.f);
    const float fz = fminf(fmaxf((cz - cmin.z) * cscale.z, 0.f), 2097151.f);
    keys[i] = spread21((unsigned long long)fx) << 2 | spread21((unsigned long long)fy) << 1 | spread21((unsigned long long)fz);
    idx[i] = i;
}

__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// one thread per internal node (Karras 2012, fig. 4); child ref >= 0: internal node, < 0: ~leaf index
__global__ void k_hierarchy(const unsigned long long *__restrict__ keys, int n, int *__restrict__ left, int *__restrict__ right,
                            int *__restrict__ first, int *__restrict__ last, int *__restrict__ parent_inner,
                            int *__restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int l_ref = (lo == gamma) ? ~gamma : gamma;
    const int r_ref = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = l_ref; right[i] = r_ref; first[i] = lo; last[i] = hi;
    if (l_ref >= 0) parent_inner[l_ref] = i; else parent_leaf[~l_ref] = i;
    if (r_ref >= 0) parent_inner[r_ref] = i; else parent_leaf[~r_ref] = i;
    if (i == 0) parent_inner[0] = -1;
}

// bottom-up: the second thread to arrive at a node merges its children's boxes (children are complete by then)
__global__ void k_refit(const int *__restrict__ idx, const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                        const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent_inner,
                        const int *__restrict__ parent_leaf, int *__restrict__ flags, float4 *__restrict__ nlo,
                        float4 *__restrict__ nhi, int *__restrict__ depth_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent_leaf[i];
    int height = 1;
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flags[node], 1) == 0) return;
        __threadfence();
        const int l = left[node], r = right[node];
        float4 alo, ahi, clo, chi;
        if (l >= 0) { alo = __ldcg(&nlo[l]); ahi = __ldcg(&nhi[l]); } else { alo = blo[idx[~l]]; ahi = bhi[idx[~l]]; }
        if (r >= 0) { clo = __ldcg(&nlo[r]); chi = __ldcg(&nhi[r]); } else { clo = blo[idx[~r]]; chi = bhi[idx[~r]]; }
        // w carries the subtree height so the host can check it against the traversal stack
        const float h = fmaxf(l >= 0 ? alo.w : 0.f, r >= 0 ? clo.w : 0.f) + 1.f;
        __stcg(&nlo[node], make_float4(fminf(alo.x, clo.x), fminf(alo.y, clo.y), fminf(alo.z, clo.z), h));
        __stcg(&nhi[node], make_float4(fmaxf(ahi.x, chi.x), fmaxf(ahi.y, chi.y), fmaxf(ahi.z, chi.z), 0.f));
        height = (int)h;
        node = parent_inner[node];
    }
    atomicMax(depth_out, height);
}

__global__ void k_alive(const int *__restrict__ first, const int *__restrict__ last, int n_inner, int leaf_max, int *__restrict__ alive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
    alive[i] = (last[i] - first[i] + 1) > leaf_max ? 1 : 0;
}

__device__ __forceinline__ void child_ref_and_box(int c, const int *alive, const int *new_index, const int *first, const int *last,
                                                  const int *idx, const float4 *blo, const float4 *bhi, const float4 *nlo,
                                                  const float4 *nhi, int &ref, float4 &lo, float4 &hi) {
    if (c >= 0) {
        lo = nlo[c]; hi = nhi[c];
        if (alive[c]) ref = new_index[c];
        else ref = ~((first[c] << 3) | (last[c] - first[c]));
    } else {
        const int k = ~c;
        lo = blo[idx[k]]; hi = bhi[idx[k]];
        ref = ~((k << 3) | 0);
    }
}

// A 4-wide node is a binary node at even depth together with its (alive) children: up to four grandchildren.
__global__ void k_select4(int n_inner, const int *__restrict__ alive, const int *__restrict__ parent_inner, int *__restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner) return;
    int depth = 0;
    for (int p = parent_inner[i]; p >= 0; p = parent_inner[p]) depth++;
    sel[i] = (alive[i] && (depth & 1) == 0) ? 1 : 0;
}

__global__ void k_emit_nodes4(int n_inner, const int *__restrict__ alive, const int *__restrict__ sel, const int *__restrict__ new4,
                              const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ first,
                              const int *__restrict__ last, const int *__restrict__ idx, const float4 *__restrict__ blo,
                              const float4 *__restrict__ bhi, const float4 *__restrict__ nlo, const float4 *__restrict__ nhi,
                              float4 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner || !sel[i]) return;
    const float inf = __int_as_float(0x7f800000);
    int ref[4] = {BVH_EMPTY_REF, BVH_EMPTY_REF, BVH_EMPTY_REF, BVH_EMPTY_REF};
    float4 lo[4], hi[4];
    for (int k = 0; k < 4; ++k) { lo[k] = make_float4(inf, inf, inf, 0.f); hi[k] = make_float4(-inf, -inf, -inf, 0.f); }
    int m = 0;
    const int kids[2] = {left[i], right[i]};
    for (int c = 0; c < 2; ++c) {
        const int k = kids[c];
        if (k >= 0 && alive[k]) {  // expand: its two children become children of the wide node
            child_ref_and_box(left[k], alive, new4, first, last, idx, blo, bhi, nlo, nhi, ref[m], lo[m], hi[m]); m++;
            child_ref_and_box(right[k], alive, new4, first, last, idx, blo, bhi, nlo, nhi, ref[m], lo[m], hi[m]); m++;
        } else {
            child_ref_and_box(k, alive, new4, first, last, idx, blo, bhi, nlo, nhi, ref[m], lo[m], hi[m]); m++;
        }
    }
    float4 *o = out + 8 * (size_t)new4[i];
    for (int c = 0; c < 4; ++c) {  // child-major: a sub-warp of four lanes fetches a node with two coalesced 64-byte accesses
        o[c] = make_float4(lo[c].x, lo[c].y, lo[c].z, hi[c].x);
        o[4 + c] = make_float4(hi[c].y, hi[c].z, __int_as_float(ref[c]), 0.f);
    }
}

// leaf-order primitive records; out_fin is what shading needs of the winner: (unit normal | obj) for a triangle --
// normalize(cross(e1, e2)) of mod.rs:605 with the device's un-fused fp32 operations, once per triangle instead of per hit --
// and (centre | obj) for a sphere
__global__ void k_gather_prims(const float4 *__restrict__ recs, const int *__restrict__ idx, int n, float4 *__restrict__ out_ae,
                               float4 *__restrict__ out_e2, float4 *__restrict__ out_fin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = idx[i];
    const float4 A = recs[3 * s], E1 = recs[3 * s + 1], E2 = recs[3 * s + 2];
    out_ae[2 * i] = A; out_ae[2 * i + 1] = E1; out_e2[i] = E2;
    if (__float_as_int(E1.w) < 0) out_fin[i] = A;
    else {
        const V3 nrm = normalize(cross(xyz(E1), xyz(E2)));
        out_fin[i] = make_float4(nrm.x, nrm.y, nrm.z, A.w);
    }
}

// Copies the top `levels` four-wide levels (breadth first from node 0, at most `cap` nodes) into `top`; a child ref that points to
// another copied node becomes BVH_TOP_BIT | its index in `top`.  One CTA; the order inside a level does not matter.
__global__ void __launch_bounds__(256) k_top_levels(const float4 *__restrict__ nodes, int levels, int cap, float4 *__restrict__ top,
                                                    int *__restrict__ n_top) {
    __shared__ int s_src[2][256], s_dst[2][256], s_n[2], s_total;
    if (threadIdx.x == 0) { s_src[0][0] = 0; s_dst[0][0] = 0; s_n[0] = 1; s_n[1] = 0; s_total = 1; }
    __syncthreads();
    for (int lvl = 0; lvl < levels; ++lvl) {
        const int c = lvl & 1, nx = c ^ 1;
        const int n_cur = s_n[c];
        __syncthreads();
        if (threadIdx.x == 0) s_n[nx] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n_cur; i += blockDim.x) {
            const int src = s_src[c][i], dst = s_dst[c][i];
            float4 rec[8];
            for (int k = 0; k < 8; ++k) rec[k] = nodes[8 * (size_t)src + k];
            for (int k = 0; k < 4; ++k) {
                const int ref = __float_as_int(rec[4 + k].z);
                if (ref >= 0 && lvl + 1 < levels) {
                    const int slot = atomicAdd(&s_total, 1);
                    if (slot < cap) {
                        const int j = atomicAdd(&s_n[nx], 1);  // (a level of a four-wide tree below depth 5 has at most 256 nodes)
                        s_src[nx][j] = ref; s_dst[nx][j] = slot;
                        rec[4 + k].z = __int_as_float(BVH_TOP_BIT | slot);
                    } else atomicSub(&s_total, 1);  // no room: the child stays a reference into bvh_nodes
                }
            }
            for (int k = 0; k < 8; ++k) top[8 * (size_t)dst + k] = rec[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_top = min(s_total, cap);
}

// ---------------------------------------------------------------------------------------------
// Binary hierarchy by binned SAH on the host, for small primitive sets (a few thousand: scenes/mesh.json has 810 triangles).
// Same output convention as k_hierarchy (Karras): n - 1 internal nodes, node 0 is the root, child ref >= 0 = internal node,
// < 0 = ~leaf index, a node covers the leaves [first, last] of the order `idx`; the device pipeline (refit, collapse to
// four-wide nodes) is the same for both builders.  Only the TOPOLOGY comes from here -- boxes are refitted on the device from
// the padded primitive boxes -- so nothing about the result depends on it, only how many nodes and primitives a ray visits
// (measured on mesh.json 1080p: see DESIGN.md section 4).
// ---------------------------------------------------------------------------------------------
struct HostBox { float lo[3], hi[3]; };
struct SahOut { std::vector<int> idx, left, right, first, last, parent_inner, parent_leaf; };

void sah_hierarchy(const std::vector<HostBox> &box, SahOut &o) {
    const int n = (int)box.size();
    o.idx.resize(n);
    for (int i = 0; i < n; ++i) o.idx[i] = i;
    const int ni = std::max(n - 1, 1);
    o.left.assign(ni, 0); o.right.assign(ni, 0); o.first.assign(ni, 0); o.last.assign(ni, 0);
    o.parent_inner.assign(std::max(n, 1), -1); o.parent_leaf.assign(std::max(n, 1), -1);
    if (n < 2) return;
    auto area = [](const float *lo, const float *hi) {
        const double dx = std::max(0.f, hi[0] - lo[0]), dy = std::max(0.f, hi[1] - lo[1]), dz = std::max(0.f, hi[2] - lo[2]);
        return dx * dy + dy * dz + dz * dx;
    };
    struct Job { int node, begin, end, depth; };
    std::vector<Job> jobs;
    jobs.push_back({0, 0, n, 0});
    int next_node = 1;
    constexpr int NB = 16;
    while (!jobs.empty()) {
        const Job j = jobs.back();
        jobs.pop_back();
        const int cnt = j.end - j.begin;
        o.first[j.node] = j.begin; o.last[j.node] = j.end - 1;
        float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
        for (int k = j.begin; k < j.end; ++k) {
            const HostBox &b = box[o.idx[k]];
            for (int a = 0; a < 3; ++a) { const float c = 0.5f * (b.lo[a] + b.hi[a]); clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c); }
        }
        int best_axis = -1, best_bin = 0;
        double best_cost = 1e300;
        for (int a = 0; a < 3; ++a) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.f)) continue;
            int bc[NB] = {0};
            float blo[NB][3], bhi[NB][3];
            for (int q = 0; q < NB; ++q) for (int c = 0; c < 3; ++c) { blo[q][c] = 3e38f; bhi[q][c] = -3e38f; }
            const float scale = NB / ext;
            for (int k = j.begin; k < j.end; ++k) {
                const HostBox &b = box[o.idx[k]];
                const int q = std::min(NB - 1, std::max(0, (int)((0.5f * (b.lo[a] + b.hi[a]) - clo[a]) * scale)));
                bc[q]++;
                for (int c = 0; c < 3; ++c) { blo[q][c] = std::min(blo[q][c], b.lo[c]); bhi[q][c] = std::max(bhi[q][c], b.hi[c]); }
            }
            double la[NB], ra[NB];
            int lc[NB], rc[NB];
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
            int c0 = 0;
            for (int q = 0; q < NB; ++q) {
                c0 += bc[q];
                for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], blo[q][c]); hi[c] = std::max(hi[c], bhi[q][c]); }
                lc[q] = c0; la[q] = c0 ? area(lo, hi) : 0.0;
            }
            for (int c = 0; c < 3; ++c) { lo[c] = 3e38f; hi[c] = -3e38f; }
            c0 = 0;
            for (int q = NB - 1; q >= 0; --q) {
                c0 += bc[q];
                for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], blo[q][c]); hi[c] = std::max(hi[c], bhi[q][c]); }
                rc[q] = c0; ra[q] = c0 ? area(lo, hi) : 0.0;
            }
            for (int q = 0; q + 1 < NB; ++q) {  // split between bin q and q + 1
                if (lc[q] == 0 || rc[q + 1] == 0) continue;
                const double cost = la[q] * lc[q] + ra[q + 1] * rc[q + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = q; }
            }
        }
        int mid;
        // all centroids coincide: split the list in the middle; likewise below depth 40, so that the tree stays within the traversal
        // stack whatever the input (40 + log2(16384) binary levels at most)
        if (best_axis < 0 || j.depth >= 40) mid = j.begin + cnt / 2;
        else {
            const float ext = chi[best_axis] - clo[best_axis], scale = NB / ext;
            auto it = std::partition(o.idx.begin() + j.begin, o.idx.begin() + j.end, [&](int p) {
                const HostBox &b = box[p];
                const int q = std::min(NB - 1, std::max(0, (int)((0.5f * (b.lo[best_axis] + b.hi[best_axis]) - clo[best_axis]) * scale)));
                return q <= best_bin;
            });
            mid = (int)(it - o.idx.begin());
            if (mid == j.begin || mid == j.end) mid = j.begin + cnt / 2;
        }
        auto child = [&](int b, int e) {
            if (e - b == 1) { o.parent_leaf[b] = j.node; return ~b; }
            const int id = next_node++;
            o.parent_inner[id] = j.node;
            jobs.push_back({id, b, e, j.depth + 1});
            return id;
        };
        o.left[j.node] = child(j.begin, mid);
        o.right[j.node] = child(mid, j.end);
    }
    o.parent_inner[0] = -1;
}

// padded primitive boxes in leaf order (for the host-side wide collapse); wide-order index map = sorted index of the leaf position
__global__ void k_leaf_boxes(const int *__restrict__ idx, const float4 *__restrict__ blo, const float4 *__restrict__ bhi, int n,
                             float4 *__restrict__ lo, float4 *__restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lo[i] = blo[idx[i]]; hi[i] = bhi[idx[i]];
}
__global__ void k_compose_order(const int *__restrict__ idx, const int *__restrict__ order, int n, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = idx[order[i]];
}

template <typename T>
cudaError_t dev_alloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), std::max<size_t>(n, 1) * sizeof(T)); }

}  // namespace

// ---------------------------------------------------------------------------------------------
void choose_bvh_objects(const ptb_scene_desc &desc, size_t max_smem_bytes, const BvhOptions &opt, std::vector<char> &in_bvh) {
    const size_t n = desc.n_objects;
    in_bvh.assign(n, 0);
    uint64_t n_spheres = 0;
    for (size_t k = 0; k < n; ++k) n_spheres += desc.objects[k].kind == PTB_OBJ_SPHERE;
    for (size_t k = 0; k < n; ++k) {
        const ptb_object &o = desc.objects[k];
        if (o.kind == PTB_OBJ_MESH) in_bvh[k] = (double)o.tri_count >= opt.min_tris;
        else in_bvh[k] = (double)n_spheres >= opt.min_spheres;
    }
    // the lock-step list must fit comfortably in shared memory (two CTAs per SM): move the largest meshes out first
    auto loose_bytes = [&]() {
        size_t b = 0;
        for (size_t k = 0; k < n; ++k)
            if (!in_bvh[k] && !(desc.objects[k].kind == PTB_OBJ_MESH && desc.objects[k].tri_count == 0))  // (empty meshes are dropped)
                b += 32 + (desc.objects[k].kind == PTB_OBJ_MESH ? 88 * (desc.objects[k].tri_count + 1) : 0);
        return b;
    };
    // (a caller that switches the BVH off for testing gets the whole opt-in shared memory of a CTA instead)
    const size_t budget = opt.min_tris > 1e9 ? max_smem_bytes - 1024 : std::min<size_t>(max_smem_bytes, 96 * 1024);
    while (loose_bytes() > budget) {
        size_t big = n;
        uint64_t big_n = 0;
        for (size_t k = 0; k < n; ++k)
            if (!in_bvh[k] && desc.objects[k].kind == PTB_OBJ_MESH && desc.objects[k].tri_count >= big_n) { big = k; big_n = desc.objects[k].tri_count; }
        if (big == n || big_n == 0) {  // only spheres / empty meshes left: move all spheres
            bool moved = false;
            for (size_t k = 0; k < n; ++k)
                if (!in_bvh[k] && desc.objects[k].kind == PTB_OBJ_SPHERE) { in_bvh[k] = 1; moved = true; }
            if (!moved) break;
        } else in_bvh[big] = 1;
    }
}

void bvh_release(BvhDevice &b) {
    if (b.nodes) cudaFree(b.nodes);
    if (b.tris) cudaFree(b.tris);
    if (b.top) cudaFree(b.top);
    if (b.nodes8) cudaFree(b.nodes8);
    if (b.tris8) cudaFree(b.tris8);
    if (b.top_count) cudaFree(b.top_count);
    b = BvhDevice{};
}

#define BV(expr)                                   \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) { err = #expr; rc = e__; goto done; } \
    } while (0)

cudaError_t bvh_build(const ptb_scene_desc &desc, const std::vector<char> &in_bvh, const std::vector<uint32_t> &prio_base, const BvhOptions &opt,
                      BvhDevice &out, DScene &ds, cudaStream_t st, double *build_ms, std::string &err) {
    out.n_nodes = out.n_tris = out.n_spheres = 0;
    out.max_depth = 0;
    ds.bvh_root = BVH_EMPTY_REF;
    ds.bvh_nodes = nullptr; ds.bvh_tri = nullptr; ds.bvh_e2 = nullptr; ds.bvh_fin = nullptr; ds.n_bvh_nodes = 0;
    ds.bvh_top = nullptr; ds.n_bvh_top = 0; ds.n_bvh_prims = 0;
    ds.bvh8_nodes = nullptr; ds.n_bvh8_nodes = 0; ds.bvh8_tri = nullptr; ds.bvh8_e2 = nullptr; ds.bvh8_fin = nullptr;
    out.n_nodes8 = 0;
    ds.bvh_lo = mk3(0.f, 0.f, 0.f); ds.bvh_hi = mk3(0.f, 0.f, 0.f);
    if (build_ms) *build_ms = 0.0;

    // ---- host: distance bound D, primitive records, pads ----------------------------------------------------------------
    // D bounds |o - a| and t for every ray the integrator can generate: origins are the lens centre or surface points.
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    auto grow = [&](double x, double y, double z, double r) {
        const double p[3] = {x, y, z};
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k] - r); hi[k] = std::max(hi[k], p[k] + r); }
    };
    size_t n_prims = 0;
    for (size_t k = 0; k < desc.n_objects; ++k) {
        const ptb_object &o = desc.objects[k];
        if (o.kind == PTB_OBJ_SPHERE) { grow(o.position[0], o.position[1], o.position[2], o.radius); n_prims += in_bvh[k] ? 1 : 0; }
        else {
            for (uint64_t j = 0; j < o.tri_count; ++j) {
                const ptb_triangle &t = desc.triangles[o.tri_begin + j];
                grow(t.a[0] + o.position[0], t.a[1] + o.position[1], t.a[2] + o.position[2], 0);
                grow(t.b[0] + o.position[0], t.b[1] + o.position[1], t.b[2] + o.position[2], 0);
                grow(t.c[0] + o.position[0], t.c[1] + o.position[1], t.c[2] + o.position[2], 0);
            }
            n_prims += in_bvh[k] ? o.tri_count : 0;
        }
    }
    if (n_prims == 0) return cudaSuccess;
    {
        const ptb_camera &c = desc.camera;
        grow(c.position[0] + c.direction[0] * c.focal_length, c.position[1] + c.direction[1] * c.focal_length,
             c.position[2] + c.direction[2] * c.focal_length, 0);
    }
    const double diag = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
    double coord_max = 0;
    for (int k = 0; k < 3; ++k) coord_max = std::max(coord_max, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
    const float D = (float)(1.5 * diag);  // 1.5x: hit points are themselves computed with rounding and may sit just outside

    std::vector<float4> recs;
    std::vector<float> pads;
    recs.reserve(3 * n_prims);
    pads.reserve(n_prims);
    unsigned n_tris = 0, n_sph = 0;
    double cl[3] = {1e300, 1e300, 1e300}, ch[3] = {-1e300, -1e300, -1e300};
    auto centroid = [&](double x, double y, double z) {
        const double p[3] = {x, y, z};
        for (int k = 0; k < 3; ++k) { cl[k] = std::min(cl[k], p[k]); ch[k] = std::max(ch[k], p[k]); }
    };
    for (size_t k = 0; k < desc.n_objects; ++k) {
        if (!in_bvh[k]) continue;
        const ptb_object &o = desc.objects[k];
        if (o.kind == PTB_OBJ_SPHERE) {
            recs.push_back(f4(o.position[0], o.position[1], o.position[2], ibits((int32_t)k)));
            recs.push_back(f4(o.radius * o.radius, 0.f, 0.f, ibits(-1)));  // r^2 as in mod.rs:416
            recs.push_back(f4(0.f, 0.f, 0.f, ubits(prio_base[k])));
            pads.push_back(o.radius + (float)opt.pad_scale * (sphere_extent(o.radius, D, (float)coord_max) - o.radius));
            centroid(o.position[0], o.position[1], o.position[2]);
            n_sph++;
        } else {
            const V3 off = v3(o.position);
            for (uint64_t j = 0; j < o.tri_count; ++j) {
                const ptb_triangle &t = desc.triangles[o.tri_begin + j];
                const V3 a = v3(t.a) + off, b = v3(t.b) + off, c = v3(t.c) + off;  // mod.rs:559
                const V3 e1 = b - a, e2 = c - a;                                    // mod.rs:560-561
                recs.push_back(f4(a.x, a.y, a.z, ibits((int32_t)k)));
                recs.push_back(f4(e1.x, e1.y, e1.z, ibits((int32_t)j)));
                recs.push_back(f4(e2.x, e2.y, e2.z, ubits(prio_base[k] + (uint32_t)j)));
                pads.push_back((float)opt.pad_scale * triangle_pad(e1, e2, D, (float)coord_max));
                centroid(a.x + (e1.x + e2.x) / 3.0, a.y + (e1.y + e2.y) / 3.0, a.z + (e1.z + e2.z) / 3.0);
                n_tris++;
            }
        }
    }
    const int n = (int)n_prims;

    cudaError_t rc = cudaSuccess;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float4 *d_recs = nullptr, *d_blo = nullptr, *d_bhi = nullptr, *d_nlo = nullptr, *d_nhi = nullptr;
    float *d_pads = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    int *d_idx = nullptr, *d_idx2 = nullptr, *d_left = nullptr, *d_right = nullptr, *d_first = nullptr, *d_last = nullptr;
    int *d_pin = nullptr, *d_pleaf = nullptr, *d_flags = nullptr, *d_alive = nullptr, *d_new = nullptr, *d_depth = nullptr, *d_sel = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    const int T = 256, B = (n + T - 1) / T;
    const int n_inner = n - 1;
    int h_depth = 0, n_alive = 0, n_top = 0;
    float4 h_root[2];  // padded box around everything in the BVH

    BV(cudaEventCreate(&ev0)); BV(cudaEventCreate(&ev1));
    BV(dev_alloc(&d_recs, 3 * (size_t)n)); BV(dev_alloc(&d_pads, (size_t)n));
    BV(dev_alloc(&d_blo, (size_t)n)); BV(dev_alloc(&d_bhi, (size_t)n));
    BV(dev_alloc(&d_keys, (size_t)n)); BV(dev_alloc(&d_keys2, (size_t)n));
    BV(dev_alloc(&d_idx, (size_t)n)); BV(dev_alloc(&d_idx2, (size_t)n));
    BV(dev_alloc(&d_left, (size_t)n)); BV(dev_alloc(&d_right, (size_t)n)); BV(dev_alloc(&d_first, (size_t)n)); BV(dev_alloc(&d_last, (size_t)n));
    BV(dev_alloc(&d_pin, (size_t)n)); BV(dev_alloc(&d_pleaf, (size_t)n)); BV(dev_alloc(&d_flags, (size_t)n));
    BV(dev_alloc(&d_alive, (size_t)n)); BV(dev_alloc(&d_new, (size_t)n)); BV(dev_alloc(&d_depth, 1)); BV(dev_alloc(&d_sel, (size_t)n));
    BV(dev_alloc(&d_nlo, (size_t)n)); BV(dev_alloc(&d_nhi, (size_t)n));
    BV(cudaMemcpyAsync(d_recs, recs.data(), recs.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
    BV(cudaMemcpyAsync(d_pads, pads.data(), pads.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    BV(cudaEventRecord(ev0, st));

    // output arrays (kept): primitive records in leaf order
    if (out.cap_tris < (size_t)n) {
        if (out.tris) cudaFree(out.tris);
        out.tris = nullptr; out.cap_tris = 0;
        BV(dev_alloc(&out.tris, 4 * (size_t)n));
        out.cap_tris = (size_t)n;
    }

    k_prim_boxes<<<B, T, 0, st>>>(d_recs, d_pads, n, d_blo, d_bhi);
    BV(cudaGetLastError());
    if (n > 1 && n <= opt.sah_max_prims) {
        // small set: binned-SAH topology from the host (boxes here are the unpadded ones, good enough to choose splits)
        std::vector<HostBox> hb((size_t)n);
        for (int i = 0; i < n; ++i) {
            const float4 A = recs[3 * (size_t)i], E1 = recs[3 * (size_t)i + 1], E2 = recs[3 * (size_t)i + 2];
            int32_t tag;
            std::memcpy(&tag, &E1.w, 4);
            const float a[3] = {A.x, A.y, A.z};
            if (tag < 0) {
                const float r = std::sqrt(E1.x);
                for (int c = 0; c < 3; ++c) { hb[i].lo[c] = a[c] - r; hb[i].hi[c] = a[c] + r; }
            } else {
                const float e1[3] = {E1.x, E1.y, E1.z}, e2[3] = {E2.x, E2.y, E2.z};
                for (int c = 0; c < 3; ++c) {
                    hb[i].lo[c] = std::min(a[c], std::min(a[c] + e1[c], a[c] + e2[c]));
                    hb[i].hi[c] = std::max(a[c], std::max(a[c] + e1[c], a[c] + e2[c]));
                }
            }
        }
        SahOut so;
        sah_hierarchy(hb, so);
        const size_t bi = sizeof(int) * (size_t)n_inner, bl = sizeof(int) * (size_t)n;
        BV(cudaMemcpyAsync(d_idx2, so.idx.data(), bl, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_left, so.left.data(), bi, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_right, so.right.data(), bi, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_first, so.first.data(), bi, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_last, so.last.data(), bi, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_pin, so.parent_inner.data(), bl, cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(d_pleaf, so.parent_leaf.data(), bl, cudaMemcpyHostToDevice, st));
        BV(cudaStreamSynchronize(st));  // (the host vectors go out of scope)
        BV(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_alive, d_new, n_inner, st));
        BV(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    } else if (n > 1) {
        float3 cmin = make_float3((float)cl[0], (float)cl[1], (float)cl[2]);
        auto sc = [&](int k) { const double e = ch[k] - cl[k]; return (float)(e > 0 ? 2097151.0 / e : 0.0); };
        float3 cscale = make_float3(sc(0), sc(1), sc(2));
        k_morton<<<B, T, 0, st>>>(d_blo, d_bhi, n, cmin, cscale, d_keys, d_idx);
        BV(cudaGetLastError());
        BV(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, n, 0, 63, st));
        BV(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_alive, d_new, n_inner, st));
        tmp_bytes = std::max(tmp_bytes, tmp2);
        BV(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
        BV(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, n, 0, 63, st));
        k_hierarchy<<<(n_inner + T - 1) / T, T, 0, st>>>(d_keys2, n, d_left, d_right, d_first, d_last, d_pin, d_pleaf);
        BV(cudaGetLastError());
    }
    if (n > 1) {
        BV(cudaMemsetAsync(d_flags, 0, sizeof(int) * (size_t)n, st));
        BV(cudaMemsetAsync(d_depth, 0, sizeof(int), st));
        k_refit<<<B, T, 0, st>>>(d_idx2, d_blo, d_bhi, n, d_left, d_right, d_pin, d_pleaf, d_flags, d_nlo, d_nhi, d_depth);
        BV(cudaGetLastError());
        k_alive<<<(n_inner + T - 1) / T, T, 0, st>>>(d_first, d_last, n_inner, std::min(8, std::max(1, n <= opt.sah_max_prims ? opt.leaf_max_small : opt.leaf_max)), d_alive);
        BV(cudaGetLastError());
        k_select4<<<(n_inner + T - 1) / T, T, 0, st>>>(n_inner, d_alive, d_pin, d_sel);
        BV(cudaGetLastError());
        BV(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_sel, d_new, n_inner, st));
        int last_alive = 0, last_new = 0;
        BV(cudaMemcpyAsync(&last_alive, d_sel + (n_inner - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        BV(cudaMemcpyAsync(&last_new, d_new + (n_inner - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        BV(cudaMemcpyAsync(&h_depth, d_depth, sizeof(int), cudaMemcpyDeviceToHost, st));
        BV(cudaMemcpyAsync(&h_root[0], d_nlo, sizeof(float4), cudaMemcpyDeviceToHost, st));  // Karras' node 0 is the root
        BV(cudaMemcpyAsync(&h_root[1], d_nhi, sizeof(float4), cudaMemcpyDeviceToHost, st));
        BV(cudaStreamSynchronize(st));
        n_alive = last_alive + last_new;
        // a four-wide step pushes at most three entries per pair of binary levels
        if (3 * (h_depth / 2 + 1) > BVH_STACK) { err = "BVH deeper than the traversal stack"; rc = cudaErrorInvalidValue; goto done; }
    } else {
        BV(cudaMemcpyAsync(d_idx2, d_idx, 0, cudaMemcpyDeviceToDevice, st));
        int zero = 0;
        BV(cudaMemcpyAsync(d_idx2, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
        BV(cudaMemcpyAsync(&h_root[0], d_blo, sizeof(float4), cudaMemcpyDeviceToHost, st));
        BV(cudaMemcpyAsync(&h_root[1], d_bhi, sizeof(float4), cudaMemcpyDeviceToHost, st));
        BV(cudaStreamSynchronize(st));
    }
    if (n_alive > 0) {
        if (out.cap_nodes < (size_t)n_alive) {
            if (out.nodes) cudaFree(out.nodes);
            out.nodes = nullptr; out.cap_nodes = 0;
            BV(dev_alloc(&out.nodes, 8 * (size_t)n_alive));
            out.cap_nodes = (size_t)n_alive;
        }
        k_emit_nodes4<<<(n_inner + T - 1) / T, T, 0, st>>>(n_inner, d_alive, d_sel, d_new, d_left, d_right, d_first, d_last, d_idx2, d_blo,
                                                            d_bhi, d_nlo, d_nhi, out.nodes);
        BV(cudaGetLastError());
        ds.bvh_root = 0;  // Karras' internal node 0 is the root and is always alive here
        if (opt.top_levels > 0) {
            if (!out.top) { BV(dev_alloc(&out.top, 8 * (size_t)BVH_TOP_MAX)); BV(dev_alloc(&out.top_count, 1)); }
            k_top_levels<<<1, 256, 0, st>>>(out.nodes, std::min(opt.top_levels, 5), BVH_TOP_MAX, out.top, out.top_count);
            BV(cudaGetLastError());
            BV(cudaMemcpyAsync(&n_top, out.top_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        }
    } else {
        ds.bvh_root = ~((0 << 3) | (n - 1));  // the whole set fits one leaf
    }
    k_gather_prims<<<B, T, 0, st>>>(d_recs, d_idx2, n, out.tris, out.tris + 2 * (size_t)n, out.tris + 3 * (size_t)n);
    BV(cudaGetLastError());
    if ((opt.wide == 1 || (opt.wide < 0 && n > opt.sah_max_prims)) && n > 1) {
        // ---- compressed eight-wide BVH over the same binary hierarchy: collapse on the host (pt_bvh8_build.cpp), upload
        float4 *d_llo = nullptr, *d_lhi = nullptr;
        int *d_order = nullptr;
        std::vector<int> h_left((size_t)n_inner), h_right((size_t)n_inner), h_first((size_t)n_inner), h_last((size_t)n_inner);
        std::vector<float4> h_nlo((size_t)n_inner), h_nhi((size_t)n_inner), h_llo((size_t)n), h_lhi((size_t)n);
        cudaError_t we = cudaSuccess;
        auto W = [&](cudaError_t e) { if (we == cudaSuccess) we = e; return e == cudaSuccess; };
        W(dev_alloc(&d_llo, (size_t)n)); W(dev_alloc(&d_lhi, (size_t)n)); W(dev_alloc(&d_order, (size_t)n));
        if (we == cudaSuccess) {
            k_leaf_boxes<<<B, T, 0, st>>>(d_idx2, d_blo, d_bhi, n, d_llo, d_lhi);
            const size_t bi = sizeof(int) * (size_t)n_inner, bf = sizeof(float4) * (size_t)n_inner;
            W(cudaMemcpyAsync(h_left.data(), d_left, bi, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_right.data(), d_right, bi, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_first.data(), d_first, bi, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_last.data(), d_last, bi, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_nlo.data(), d_nlo, bf, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_nhi.data(), d_nhi, bf, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_llo.data(), d_llo, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, st));
            W(cudaMemcpyAsync(h_lhi.data(), d_lhi, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, st));
            W(cudaStreamSynchronize(st));
        }
        if (we == cudaSuccess) {
            std::vector<Bvh8Box> nb((size_t)n_inner), lb((size_t)n);
            for (int i = 0; i < n_inner; ++i) nb[i] = Bvh8Box{{h_nlo[i].x, h_nlo[i].y, h_nlo[i].z}, {h_nhi[i].x, h_nhi[i].y, h_nhi[i].z}};
            for (int i = 0; i < n; ++i) lb[i] = Bvh8Box{{h_llo[i].x, h_llo[i].y, h_llo[i].z}, {h_lhi[i].x, h_lhi[i].y, h_lhi[i].z}};
            std::vector<uint32_t> wn;
            std::vector<int> order;
            const int wdepth = bvh8_collapse(n, h_left.data(), h_right.data(), h_first.data(), h_last.data(), nb.data(), lb.data(),
                                             (double)D + coord_max, wn, order, opt.wide_sah);
            if (wdepth > 0 && wdepth + 2 <= BVH_STACK && (int)order.size() == n) {
                const size_t nw = wn.size() / 24;
                if (out.cap_nodes8 < nw) {
                    if (out.nodes8) cudaFree(out.nodes8);
                    out.nodes8 = nullptr; out.cap_nodes8 = 0;
                    if (W(dev_alloc(&out.nodes8, 6 * nw))) out.cap_nodes8 = nw;
                }
                if (out.cap_tris8 < (size_t)n) {
                    if (out.tris8) cudaFree(out.tris8);
                    out.tris8 = nullptr; out.cap_tris8 = 0;
                    if (W(dev_alloc(&out.tris8, 4 * (size_t)n))) out.cap_tris8 = (size_t)n;
                }
                if (we == cudaSuccess) {
                    W(cudaMemcpyAsync(out.nodes8, wn.data(), wn.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                    W(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
                    k_compose_order<<<B, T, 0, st>>>(d_idx2, d_order, n, d_idx);  // (d_idx is free again: the sort has consumed it)
                    k_gather_prims<<<B, T, 0, st>>>(d_recs, d_idx, n, out.tris8, out.tris8 + 2 * (size_t)n, out.tris8 + 3 * (size_t)n);
                    W(cudaGetLastError());
                    W(cudaStreamSynchronize(st));  // (the host vectors go out of scope)
                    if (we == cudaSuccess) out.n_nodes8 = (unsigned)nw;
                }
            }
        }
        cudaFree(d_llo); cudaFree(d_lhi); cudaFree(d_order);
        if (we != cudaSuccess) { err = "wide BVH build"; rc = we; goto done; }
    }
    BV(cudaEventRecord(ev1, st));
    BV(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        BV(cudaEventElapsedTime(&ms, ev0, ev1));
        if (build_ms) *build_ms = ms;
    }
    out.n_nodes = (unsigned)n_alive; out.n_tris = n_tris; out.n_spheres = n_sph; out.max_depth = h_depth;
    ds.bvh_nodes = out.nodes; ds.bvh_tri = out.tris; ds.bvh_e2 = out.tris + 2 * (size_t)n; ds.bvh_fin = out.tris + 3 * (size_t)n;
    ds.n_bvh_nodes = n_alive; ds.n_bvh_prims = n;
    ds.bvh_top = out.top; ds.n_bvh_top = n_top;
    if (out.n_nodes8 > 0) {
        ds.bvh8_nodes = out.nodes8; ds.n_bvh8_nodes = (int)out.n_nodes8; ds.bvh8_magic = 0x4B000000u;
        ds.bvh8_tri = out.tris8; ds.bvh8_e2 = out.tris8 + 2 * (size_t)n; ds.bvh8_fin = out.tris8 + 3 * (size_t)n;
    }
    ds.bvh_lo = mk3(h_root[0].x, h_root[0].y, h_root[0].z); ds.bvh_hi = mk3(h_root[1].x, h_root[1].y, h_root[1].z);

done:
    cudaFree(d_recs); cudaFree(d_pads); cudaFree(d_blo); cudaFree(d_bhi); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_idx);
    cudaFree(d_idx2); cudaFree(d_left); cudaFree(d_right); cudaFree(d_first); cudaFree(d_last); cudaFree(d_pin); cudaFree(d_pleaf);
    cudaFree(d_flags); cudaFree(d_sel); cudaFree(d_alive); cudaFree(d_new); cudaFree(d_depth); cudaFree(d_nlo); cudaFree(d_nhi); cudaFree(d_tmp);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (rc != cudaSuccess) ds.bvh_root = BVH_EMPTY_REF;
    return rc;
}

}  // namespace ptb
