// CPU check of the compressed eight-wide BVH (pt_bvh8.h): builds it over random boxes, walks it with the SAME node test the CUDA
// kernel uses (pt_bvh8.cuh is host + device) and the same group / stack logic, and verifies against an exact (double) slab test
// that every primitive whose box the ray segment touches is reported -- i.e. nothing the brute-force scan could hit is culled.
// Built and run by tests/test_bvh8_cpu.py:  g++ -O2 -std=c++17 -ffp-contract=off bvh8_check.cpp ../../path_tracer_rust_b200/csrc/pt_bvh8_build.cpp
#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../path_tracer_rust_b200/csrc/pt_bvh8.cuh"
#include "../../path_tracer_rust_b200/csrc/pt_bvh8.h"

using namespace ptb;

struct Tree { std::vector<int> left, right, first, last; std::vector<Bvh8Box> node_box; };

static Bvh8Box merge(const Bvh8Box &a, const Bvh8Box &b) {
    Bvh8Box r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
    return r;
}

// median split on the widest centroid axis; Karras convention (node 0 = root, child >= 0 inner, < 0 ~leaf position)
static int build(Tree &t, std::vector<Bvh8Box> &leaf, int b, int e, int &next) {
    if (e - b == 1) return ~b;
    const int id = next++;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = b; i < e; ++i)
        for (int k = 0; k < 3; ++k) { const float c = 0.5f * (leaf[i].lo[k] + leaf[i].hi[k]); lo[k] = std::min(lo[k], c); hi[k] = std::max(hi[k], c); }
    int ax = 0;
    for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[ax] - lo[ax]) ax = k;
    const int mid = (b + e) / 2;
    std::nth_element(leaf.begin() + b, leaf.begin() + mid, leaf.begin() + e,
                     [ax](const Bvh8Box &x, const Bvh8Box &y) { return x.lo[ax] + x.hi[ax] < y.lo[ax] + y.hi[ax]; });
    t.first[id] = b; t.last[id] = e - 1;
    const int l = build(t, leaf, b, mid, next), r = build(t, leaf, mid, e, next);
    t.left[id] = l; t.right[id] = r;
    t.node_box[id] = merge(l < 0 ? leaf[~l] : t.node_box[l], r < 0 ? leaf[~r] : t.node_box[r]);
    return id;
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 20000, n_rays = argc > 2 ? std::atoi(argv[2]) : 20000;
    std::mt19937 rng(12345u + (unsigned)n);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<Bvh8Box> leaf(n);
    for (int i = 0; i < n; ++i) {  // a dense "mesh" ball of tiny boxes, scattered small boxes, a few flat and a few huge ones
        const float kind = U(rng);
        float c[3], h[3];
        if (kind < 0.6f) { for (int k = 0; k < 3; ++k) { c[k] = -5.f + 10.f * U(rng); h[k] = 0.002f + 0.03f * U(rng); } c[1] -= 12.f; }
        else if (kind < 0.95f) { for (int k = 0; k < 3; ++k) { c[k] = -50.f + 100.f * U(rng); h[k] = 0.1f + 0.7f * U(rng); } }
        else if (kind < 0.98f) { for (int k = 0; k < 3; ++k) { c[k] = -50.f + 100.f * U(rng); h[k] = 2.f * U(rng); } h[(int)(3 * U(rng)) % 3] = 0.f; }
        else { for (int k = 0; k < 3; ++k) { c[k] = -20.f + 40.f * U(rng); h[k] = 5.f + 30.f * U(rng); } }
        for (int k = 0; k < 3; ++k) { leaf[i].lo[k] = c[k] - h[k]; leaf[i].hi[k] = c[k] + h[k]; }
    }
    Tree t;
    t.left.assign(std::max(n - 1, 1), 0); t.right = t.first = t.last = t.left; t.node_box.resize(std::max(n - 1, 1));
    int next = 0;
    if (n > 1) build(t, leaf, 0, n, next);
    std::vector<uint32_t> nodes;
    std::vector<int> order;
    const double d_bound = 400.0;
    const int depth = bvh8_collapse(n, t.left.data(), t.right.data(), t.first.data(), t.last.data(), t.node_box.data(), leaf.data(), d_bound,
                                    nodes, order);
    if (depth <= 0) { std::printf("collapse failed\n"); return 2; }
    const size_t n_wide = nodes.size() / 24;
    // every primitive exactly once in the wide order
    std::vector<int> seen(n, 0);
    for (int k : order) seen[k]++;
    for (int i = 0; i < n; ++i) if (seen[i] != 1) { std::printf("primitive %d appears %d times\n", i, seen[i]); return 3; }

    long long missed = 0, touched = 0, visited_total = 0, nodes_total = 0;
    std::vector<char> mark(n);
    for (int r = 0; r < n_rays; ++r) {
        float o[3], d[3];
        for (int k = 0; k < 3; ++k) { o[k] = -60.f + 120.f * U(rng); d[k] = -1.f + 2.f * U(rng); }
        if (r % 7 == 0) d[(r / 7) % 3] = 0.f;                       // axis-parallel components
        if (r % 11 == 0) { o[0] = 0.f; o[1] = -12.f; o[2] = 0.f; }  // from inside the dense ball
        if (r % 5 == 0) {  // grazing: aimed at a corner / edge point of some box, from near or far
            const Bvh8Box &b = leaf[(size_t)(U(rng) * n) % n];
            float tgt[3];
            for (int k = 0; k < 3; ++k) {
                const float u = U(rng);
                tgt[k] = u < 0.4f ? b.lo[k] : (u < 0.8f ? b.hi[k] : b.lo[k] + (b.hi[k] - b.lo[k]) * U(rng));
                tgt[k] += (U(rng) - 0.5f) * 2e-6f * (1.0f + std::fabs(tgt[k]));
            }
            if (r % 10 == 0) for (int k = 0; k < 3; ++k) o[k] = tgt[k] + (U(rng) - 0.5f) * 0.02f;   // origin almost on the box
            for (int k = 0; k < 3; ++k) d[k] = tgt[k] - o[k];
        }
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (len == 0.f) continue;
        for (int k = 0; k < 3; ++k) d[k] /= len;
        const float tmax = r % 3 == 0 ? INFINITY : 5.f + 150.f * U(rng);
        float id[3];
        for (int k = 0; k < 3; ++k) { const float a = std::fabs(d[k]) < 1e-20f ? std::copysign(1e-20f, d[k]) : d[k]; id[k] = 1.0f / a; }
        const uint32_t octinv = bvh8_octinv(id[0], id[1], id[2]);
        std::fill(mark.begin(), mark.end(), 0);
        // the kernel's traversal: current node group (base, hits << 24 | imask), one stack entry per node
        uint32_t ng_base = 0, ng_mask = 0x80000000u;
        std::vector<std::pair<uint32_t, uint32_t>> stack;
        for (;;) {
            if (ng_mask <= 0x00ffffffu) {
                if (stack.empty()) break;
                ng_base = stack.back().first; ng_mask = stack.back().second;
                stack.pop_back();
            }
            const int bit = 31 - __builtin_clz(ng_mask);
            ng_mask &= ~(1u << bit);
            const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
            const uint32_t node = ng_base + (uint32_t)__builtin_popcount(ng_mask & 0xffu & ((1u << slot) - 1u));
            if (ng_mask > 0x00ffffffu) stack.push_back({ng_base, ng_mask});
            if (node >= n_wide) { std::printf("node index %u out of range\n", node); return 4; }
            Bvh8Node nd;
            std::memcpy(nd.w, &nodes[24 * (size_t)node], 80);
            nodes_total++;
            const uint32_t hits = bvh8_node_hits(nd, o[0], o[1], o[2], id[0], id[1], id[2], tmax, octinv, 0x4B000000u);
            ng_base = nd.w[4];
            ng_mask = (hits & 0xff000000u) | (nd.w[3] >> 24);
            uint32_t tg = hits & 0x00ffffffu;
            while (tg) {
                const int b = 31 - __builtin_clz(tg);
                tg &= ~(1u << b);
                const size_t pos = nd.w[5] + (uint32_t)b;
                if (pos >= order.size()) { std::printf("primitive position out of range\n"); return 5; }
                mark[order[pos]] = 1;
                visited_total++;
            }
        }
        for (int i = 0; i < n; ++i) {  // exact slab test in double
            double tn = 0.0, tf = tmax;
            bool hit = true;
            for (int k = 0; k < 3 && hit; ++k) {
                if (d[k] == 0.f) { if (o[k] < leaf[i].lo[k] || o[k] > leaf[i].hi[k]) hit = false; continue; }
                double a = ((double)leaf[i].lo[k] - o[k]) / d[k], b = ((double)leaf[i].hi[k] - o[k]) / d[k];
                if (a > b) std::swap(a, b);
                tn = std::max(tn, a); tf = std::min(tf, b);
                if (tn > tf) hit = false;
            }
            if (hit) { touched++; if (!mark[i]) missed++; }
        }
    }
    std::printf("BVH8_CHECK prims %d wide_nodes %zu depth %d rays %d touched %lld missed %lld visited_per_ray %.1f nodes_per_ray %.1f\n", n, n_wide, depth,
                n_rays, touched, missed, (double)visited_total / n_rays, (double)nodes_total / n_rays);
    return missed == 0 ? 0 : 1;
}
