// pt_bvh8_build.cpp -- host side of the compressed eight-wide BVH (layout and parity argument: pt_bvh8.h).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>

#include "pt_bvh8.h"

namespace ptb {

namespace {

inline double area(const Bvh8Box &b) {
    const double dx = std::max(0.f, b.hi[0] - b.lo[0]), dy = std::max(0.f, b.hi[1] - b.lo[1]), dz = std::max(0.f, b.hi[2] - b.lo[2]);
    return dx * dy + dy * dz + dz * dx;
}

}  // namespace

int bvh8_collapse(int n, const int *left, const int *right, const int *first, const int *last, const Bvh8Box *node_box,
                  const Bvh8Box *leaf_box, double d_bound, std::vector<uint32_t> &nodes, std::vector<int> &order, int sah_collapse) {
    nodes.clear();
    order.clear();
    if (n <= 0) return 0;
    // child ref: >= 0 binary inner node, < 0 ~leaf position
    auto ref_first = [&](int r) { return r < 0 ? ~r : first[r]; };
    auto ref_last = [&](int r) { return r < 0 ? ~r : last[r]; };
    auto ref_size = [&](int r) { return ref_last(r) - ref_first(r) + 1; };
    auto ref_box = [&](int r) -> const Bvh8Box & { return r < 0 ? leaf_box[~r] : node_box[r]; };
    auto is_leaf = [&](int r) { return ref_size(r) <= BVH8_LEAF_MAX; };

    // ---- pass 0 (optional): which binary subtrees become the children of a wide node, by the cost recurrence of Ylitie et al. 3.1:
    //   C(n, 1) = min( leaf cost, A_n c_node + D(n, 8) )          one root: a leaf child, or a wide node whose children are a forest of <= 8
    //   C(n, i) = min( D(n, i), C(n, i - 1) )   i = 2..7          a forest of at most i roots covering n's primitives, no node at n
    //   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
    // with A the surface area; a leaf child costs A c_prim per primitive (one primitive per leaf child here).
    constexpr float C_NODE = 1.0f, INF = 3e38f;
    const float C_PRIM = 0.25f * (float)std::max(1, sah_collapse);  // (sah_collapse doubles as the weight in quarters: 2 = 0.5)
    struct Dp { float c[8]; uint8_t split[9]; uint8_t internal; };  // c[i], i = 1..7; split[j]: the k of D(n, j), 0 = "use C(n, j-1)"; internal: C(n,1) is a wide node
    std::vector<Dp> dp;
    auto cost_of = [&](int r, int i) -> float {  // C(r, i) for a child ref
        if (r < 0) return (float)area(leaf_box[~r]) * C_PRIM;
        return dp[r].c[i];
    };
    if (sah_collapse && n > 1) {
        dp.resize((size_t)n - 1);
        // children before parents.  The tree is cut a few levels below the root: the subtrees below the cut are independent and
        // are solved by one thread each (reverse pre-order inside a subtree visits children before parents), the few nodes above
        // the cut afterwards.
        auto solve = [&](int b) {
            const int l = left[b], r = right[b];
            Dp &d = dp[b];
            float dist[9];
            dist[0] = dist[1] = INF;
            d.split[0] = d.split[1] = 0;
            for (int j = 2; j <= 8; ++j) {
                float best = INF;
                int bk = 1;
                for (int k = 1; k < j; ++k) {
                    if (k > 7 || j - k > 7) continue;
                    const float v = cost_of(l, k) + cost_of(r, j - k);
                    if (v < best) { best = v; bk = k; }
                }
                dist[j] = best;
                d.split[j] = (uint8_t)bk;
            }
            const float a_n = (float)area(node_box[b]);
            const float c_int = a_n * C_NODE + dist[8];
            const float c_leaf = is_leaf(b) ? a_n * C_PRIM * (float)ref_size(b) : INF;
            d.internal = c_int < c_leaf ? 1 : 0;
            d.c[0] = INF;
            d.c[1] = std::min(c_int, c_leaf);
            for (int i = 2; i <= 7; ++i) {
                if (dist[i] < d.c[i - 1]) d.c[i] = dist[i];
                else { d.c[i] = d.c[i - 1]; d.split[i] = 0; }  // fewer roots are at least as good
            }
        };
        auto solve_subtree = [&](int root) {
            std::vector<int> walk, st;
            st.push_back(root);
            while (!st.empty()) {
                const int b = st.back();
                st.pop_back();
                walk.push_back(b);
                if (left[b] >= 0) st.push_back(left[b]);
                if (right[b] >= 0) st.push_back(right[b]);
            }
            for (size_t w = walk.size(); w-- > 0;) solve(walk[w]);
        };
        std::vector<int> top, cut;  // nodes above the cut (pre-order), roots of the subtrees below it
        {
            std::vector<std::pair<int, int>> st;
            st.push_back({0, 0});
            while (!st.empty()) {
                const auto [b, depth] = st.back();
                st.pop_back();
                if (depth >= 6) { cut.push_back(b); continue; }
                top.push_back(b);
                if (left[b] >= 0) st.push_back({left[b], depth + 1});
                if (right[b] >= 0) st.push_back({right[b], depth + 1});
            }
        }
        {
            const unsigned hw_dp = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
            if (cut.size() < 2 || hw_dp == 1 || n < 20000) for (int c : cut) solve_subtree(c);
            else {
                std::vector<std::thread> th;
                for (unsigned t = 0; t < hw_dp; ++t)
                    th.emplace_back([&, t] { for (size_t i = t; i < cut.size(); i += hw_dp) solve_subtree(cut[i]); });
                for (auto &x : th) x.join();
            }
        }
        for (size_t w = top.size(); w-- > 0;) solve(top[w]);
    }
    // the children of the wide node that stands for binary node `b`, following the recurrence's decisions
    struct Collect {
        const int *left, *right;
        const std::vector<Dp> *dp;
        int *c;
        int nc;
        void run(int m, int budget) {
            if (m < 0 || budget <= 1) { c[nc++] = m; return; }  // a root of the forest: leaf child or wide node (decided when it is visited)
            const uint8_t k = (*dp)[m].split[budget];
            if (k == 0) { run(m, budget - 1); return; }
            run(left[m], k);
            run(right[m], budget - k);
        }
    };

    // ---- pass 1 (sequential, breadth first): children of every wide node, their slots, the numbering of nodes and primitives
    struct Wide {
        int bin;          // the binary subtree this wide node stands for (-1: the single-primitive set)
        int depth;
        int child[8];     // per slot: child ref, or INT_MIN for an empty slot
        uint32_t child_base, prim_base;
    };
    constexpr int EMPTY = INT32_MIN;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto parallel_for = [&](size_t begin, size_t end, const std::function<void(size_t, size_t)> &fn) {
        const size_t cnt = end - begin;
        if (cnt < 4096 || hw == 1) { fn(begin, end); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < hw; ++t) th.emplace_back(fn, begin + cnt * t / hw, begin + cnt * (t + 1) / hw);
        for (auto &x : th) x.join();
    };
    std::vector<Wide> wide;
    wide.reserve((size_t)n / 3 + 16);
    wide.push_back(Wide{n == 1 ? -1 : 0, 1, {EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY}, 0u, 0u});
    order.assign((size_t)n, 0);
    size_t n_order = 0;
    int max_depth = 1;
    // level by level (a level's nodes are consecutive): children and slots in parallel, then the numbering by a prefix sum, then
    // the next level's nodes and the primitive order in parallel (the walk is bound by cache misses on the binary nodes' boxes)
    std::vector<uint32_t> cnt_inner, cnt_prim;
    for (size_t lb = 0, le = 1; lb < le;) {
        if (le >= (1u << 28)) return 0;
        max_depth = wide[lb].depth;
        cnt_inner.assign(le - lb + 1, 0u);
        cnt_prim.assign(le - lb + 1, 0u);
        parallel_for(lb, le, [&](size_t w0, size_t w1) {
            for (size_t w = w0; w < w1; ++w) {
                const int bin = wide[w].bin;
                int c[8];
                int nc = 0;
                const Bvh8Box &nb = bin < 0 ? leaf_box[0] : node_box[bin];
                if (bin < 0) c[nc++] = ~0;  // the whole set is one primitive
                else if (is_leaf(bin)) c[nc++] = bin;  // (the root itself is small enough to be one leaf child)
                else if (!dp.empty()) {
                    Collect col{left, right, &dp, c, 0};
                    const uint8_t k = dp[bin].split[8];
                    col.run(left[bin], k);
                    col.run(right[bin], 8 - k);
                    nc = col.nc;
                } else {
                    c[nc++] = left[bin];
                    c[nc++] = right[bin];
                    float ar[8];
                    ar[0] = is_leaf(c[0]) ? -1.f : (float)area(ref_box(c[0]));
                    ar[1] = is_leaf(c[1]) ? -1.f : (float)area(ref_box(c[1]));
                    while (nc < 8) {  // open the inner child with the largest surface area until the node is full
                        int best = -1;
                        float best_a = -1.f;
                        for (int i = 0; i < nc; ++i)
                            if (ar[i] > best_a) { best_a = ar[i]; best = i; }
                        if (best < 0 || best_a < 0.f) break;
                        const int b = c[best];
                        c[best] = left[b];
                        ar[best] = is_leaf(c[best]) ? -1.f : (float)area(ref_box(c[best]));
                        c[nc] = right[b];
                        ar[nc] = is_leaf(c[nc]) ? -1.f : (float)area(ref_box(c[nc]));
                        nc++;
                    }
                }
                // slots: the child towards (+x, +y, +z) bits of the node's centre; greedy on the projection of the centroid offset
                float off[8][3];
                for (int i = 0; i < nc; ++i) {
                    const Bvh8Box &b = ref_box(c[i]);
                    for (int a = 0; a < 3; ++a) off[i][a] = 0.5f * (b.lo[a] + b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
                }
                int slot_child[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
                unsigned done_children = 0;
                for (int k = 0; k < nc; ++k) {
                    int bi = -1, bs = -1;
                    float bv = -3e38f;
                    for (int i = 0; i < nc; ++i) {
                        if (done_children >> i & 1u) continue;
                        for (int sl = 0; sl < 8; ++sl) {
                            if (slot_child[sl] >= 0) continue;
                            const float v = (sl & 1 ? off[i][0] : -off[i][0]) + (sl & 2 ? off[i][1] : -off[i][1]) + (sl & 4 ? off[i][2] : -off[i][2]);
                            if (v > bv) { bv = v; bi = i; bs = sl; }
                        }
                    }
                    slot_child[bs] = bi;
                    done_children |= 1u << bi;
                }
                uint32_t ni = 0, np = 0;
                for (int sl = 0; sl < 8; ++sl) {
                    const int r = slot_child[sl] < 0 ? EMPTY : c[slot_child[sl]];
                    wide[w].child[sl] = r;
                    if (r == EMPTY) continue;
                    if (is_leaf(r)) np += (uint32_t)ref_size(r); else ni++;
                }
                cnt_inner[w - lb] = ni;
                cnt_prim[w - lb] = np;
            }
        });
        // numbering: inner children consecutive in slot order, primitives of leaf children consecutive in slot order
        size_t next_node = le, next_prim = n_order;
        for (size_t w = lb; w < le; ++w) {
            wide[w].child_base = (uint32_t)next_node;
            wide[w].prim_base = (uint32_t)next_prim;
            next_node += cnt_inner[w - lb];
            next_prim += cnt_prim[w - lb];
        }
        if (next_prim > (size_t)n) return 0;
        wide.resize(next_node, Wide{0, 0, {EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY, EMPTY}, 0u, 0u});
        parallel_for(lb, le, [&](size_t w0, size_t w1) {
            for (size_t w = w0; w < w1; ++w) {
                size_t at_node = wide[w].child_base, at_prim = wide[w].prim_base;
                for (int sl = 0; sl < 8; ++sl) {
                    const int r = wide[w].child[sl];
                    if (r == EMPTY) continue;
                    if (is_leaf(r)) {
                        for (int k = ref_first(r); k <= ref_last(r); ++k) order[at_prim++] = k;
                    } else {
                        wide[at_node].bin = r;
                        wide[at_node].depth = wide[w].depth + 1;
                        at_node++;
                    }
                }
            }
        });
        n_order = next_prim;
        lb = le;
        le = next_node;
    }
    order.resize(n_order);

    // ---- pass 2 (parallel): grid of every node, quantised child boxes, meta bytes
    // the smallest grid step that keeps the traversal's decode error below one step (pt_bvh8.h): 8 u D
    const double min_step = 8.0 * 5.9604645e-8 * std::max(d_bound, 1e-30);
    const int min_e = (int)std::ceil(std::log2(min_step));
    nodes.assign(24 * wide.size(), 0u);
    auto pack = [&](size_t w0, size_t w1) {
        for (size_t w = w0; w < w1; ++w) {
            const Wide &W = wide[w];
            const Bvh8Box &nb = W.bin < 0 ? leaf_box[0] : node_box[W.bin];
            float p[3];
            int e[3];
            double inv_step[3];
            for (int a = 0; a < 3; ++a) {
                // origin one step below the node's lower corner, steps 2^e with (extent + 2 steps) <= 252 steps and step >= min_step
                const double ext = std::max(0.0, (double)nb.hi[a] - (double)nb.lo[a]);
                int ex = min_e;
                if (ext > 0.0) {
                    int fe;
                    std::frexp(ext / 250.0, &fe);  // ext / 250 = m * 2^fe, 0.5 <= m < 1: 2^fe >= ext / 250
                    ex = std::max(min_e, fe);
                }
                ex = std::max(-100, std::min(100, ex));
                e[a] = ex;
                const double step = std::ldexp(1.0, ex);
                inv_step[a] = std::ldexp(1.0, -ex);
                const double pd = (double)nb.lo[a] - step;
                float pf = (float)pd;
                if ((double)pf > pd) pf = std::nextafterf(pf, -INFINITY);  // rounded DOWN
                p[a] = pf;
            }
            uint32_t imask = 0;
            uint8_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[6][8];
            // an empty slot gets an inverted box (lower planes at 255, upper planes at 0): no ray can hit it, whatever its direction,
            // so the node test needs no "is this slot used" decision
            for (int sl = 0; sl < 8; ++sl)
                for (int a = 0; a < 3; ++a) { q[a][sl] = 255; q[3 + a][sl] = 0; }
            int prim_off = 0;
            for (int sl = 0; sl < 8; ++sl) {
                const int r = W.child[sl];
                if (r == EMPTY) continue;
                const Bvh8Box &b = ref_box(r);
                for (int a = 0; a < 3; ++a) {
#ifdef PTB_BVH8_NO_MARGIN  // negative control of tests/test_bvh8_cpu.py only: without the extra step the traversal must lose hits
                    const double margin = 0.0;
#else
                    const double margin = 1.0;  // one step for the traversal's decode error (pt_bvh8.h)
#endif
                    const double lo = std::floor(((double)b.lo[a] - (double)p[a]) * inv_step[a]) - margin;
                    const double hi = std::ceil(((double)b.hi[a] - (double)p[a]) * inv_step[a]) + margin;
                    q[a][sl] = (uint8_t)std::max(0.0, std::min(255.0, lo));
                    q[3 + a][sl] = (uint8_t)std::max(0.0, std::min(255.0, hi));
                }
                if (is_leaf(r)) {
                    const int cnt = ref_size(r);
                    meta[sl] = (uint8_t)((((1u << cnt) - 1u) << 5) | (uint32_t)prim_off);
                    prim_off += cnt;
                } else {
                    imask |= 1u << sl;
                    meta[sl] = (uint8_t)(0x20u | (24u + (uint32_t)sl));
                }
            }
            uint32_t *rec = &nodes[24 * w];
            std::memcpy(&rec[0], &p[0], 4); std::memcpy(&rec[1], &p[1], 4); std::memcpy(&rec[2], &p[2], 4);
            rec[3] = (uint32_t)(e[0] + 127) | (uint32_t)(e[1] + 127) << 8 | (uint32_t)(e[2] + 127) << 16 | imask << 24;
            rec[4] = W.child_base; rec[5] = W.prim_base;
            std::memcpy(&rec[6], &meta[0], 4); std::memcpy(&rec[7], &meta[4], 4);
            for (int k = 0; k < 6; ++k) { std::memcpy(&rec[8 + 2 * k], &q[k][0], 4); std::memcpy(&rec[9 + 2 * k], &q[k][4], 4); }
        }
    };
    parallel_for(0, wide.size(), pack);
    return max_depth;
}

}  // namespace ptb
