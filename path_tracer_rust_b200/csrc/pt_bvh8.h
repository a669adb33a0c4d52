// pt_bvh8.h -- compressed eight-wide BVH for the wavefront trace kernel (after Ylitie, Karras, Laine 2017, "Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs"), rebuilt for this backend's parity rules.
//
// Why: the four-wide traversal spends half its instructions in the four-box test + sort with 17 of 32 lanes and a fifth in branchy
// push / pop sequences (profiles/r02e_*).  An eight-wide node halves the number of dependent node fetches per ray, child boxes
// quantised to 8 bits against the node's own frame make a node 96 bytes for eight children (the four-wide node: 128 bytes for
// four), children are visited in a fixed order chosen by the ray's direction octant, so nothing is sorted and ONE stack entry
// (child base index + hit mask) stands for all postponed children of a node.
//
// Parity: a BVH only selects which primitives get tested.  Child boxes are the padded boxes of the four-wide builder (pt_bvh_build.cu:
// triangle_pad / sphere_extent), rounded OUTWARD to the node's grid and widened by one more step; the decode error of the traversal
// is below one step by construction (bvh8_collapse: the step is at least 8 u D), so no primitive the reference would accept is culled.
//
// Node = 96 bytes = 6 x uint4/float4:
//   [0] float px, py, pz (grid origin), uint32 ex | ey << 8 | ez << 16 | imask << 24   (biased exponents of the grid steps;
//       imask: bit i = slot i holds an inner node)
//   [1] uint32 child_base (index of the first inner child: inner children are consecutive, in slot order), uint32 prim_base
//       (first primitive of this node's leaf children in the wide primitive order), uint32 meta[0..3], uint32 meta[4..7]
//       meta byte of slot i: 0 = empty; inner: 0x20 | (24 + i); leaf: (unary count 1 -> 001, 2 -> 011, 3 -> 111) << 5 | offset (0..23)
//   [2] qlo_x[8] qlo_y[8]   [3] qlo_z[8] qhi_x[8]   [4] qhi_y[8] qhi_z[8]   (bytes; box = origin + q * 2^e)     [5] unused
// Slot i of a node is the child that lies towards (bit0: +x, bit1: +y, bit2: +z) of the node's centre (greedy assignment), so a ray
// visits the slots in the order of decreasing (slot XOR octinv), octinv bit a = (d[a] >= 0).
#pragma once
#include <stdint.h>

#include <vector>

namespace ptb {

constexpr int BVH8_NODE_F4 = 6;       // float4 per node
constexpr int BVH8_LEAF_MAX = 1;      // primitives per leaf child (measured: 1 -> 234, 2 -> 224, 3 -> 215 Mpaths/s on the synthetic scene)
constexpr int BVH8_TOP_MAX = 1024;    // nodes (breadth first from the root) the trace kernel keeps in shared memory

struct Bvh8Box { float lo[3], hi[3]; };

// Host: collapses a binary hierarchy (Karras convention: n - 1 inner nodes, node 0 = root, child ref >= 0 inner node, < 0 ~leaf; an
// inner node covers the leaves [first, last] of the leaf order) into breadth-first numbered eight-wide nodes.
//   node_box[i]   padded box of binary inner node i          leaf_box[k]  padded box of the primitive at leaf position k
//   d_bound       bound on |ray origin - anything| + coordinates (the D + coord_max of the padding)
//   out_nodes     6 x 4 x uint32 per wide node               out_order    wide primitive order: position -> leaf position k
// Returns the depth of the wide tree (0 = failed: more than 2^28 nodes).
//   sah_collapse  > 0: the children of every wide node are chosen by the surface-area cost recurrence of Ylitie et al. (section 3.1,
//                 dynamic programme over the binary tree) with a primitive weighing sah_collapse / 4 node visits; 0: greedily, opening
//                 the child with the largest surface area
int bvh8_collapse(int n_prims, const int *left, const int *right, const int *first, const int *last, const Bvh8Box *node_box,
                  const Bvh8Box *leaf_box, double d_bound, std::vector<uint32_t> &out_nodes, std::vector<int> &out_order,
                  int sah_collapse = 2);

}  // namespace ptb
