// pt_bvh8.cuh -- node test of the compressed eight-wide BVH (layout: pt_bvh8.h).  Host + device: the same function is walked on the
// CPU by tests/cpp/bvh8_check.cpp against an exact box test, without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "pt_bvh8.h"

#ifdef __CUDACC__
#define PTB_HD __host__ __device__ __forceinline__
#else
#define PTB_HD inline
#endif

namespace ptb {

struct Bvh8Node {  // the 80 bytes of a node that the traversal reads, as five 16-byte words
    uint32_t w[20];
};

PTB_HD float bvh8_bits_to_float(uint32_t b) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// byte k of `word` as the float 2^23 + byte (exact); the 2^23 is folded into the ray's origin term by the caller
template <int K>
PTB_HD float bvh8_byte_as_biased_float(uint32_t word, uint32_t magic /* 0x4B000000, in a register */) {
#ifdef __CUDA_ARCH__
    // one PRMT with the selector as an immediate: SASS PRMT takes a single immediate, and ptxas gives that slot to a constant
    // 0x4B000000 and re-materialises the selector into a register for every use -- so the caller hides the constant in a register
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(magic), "n"(0x7440 | K));
    return __uint_as_float(r);
#else
    (void)magic;
    const int k = K;
    return bvh8_bits_to_float(0x4B000000u | ((word >> (8 * k)) & 0xffu));
#endif
}

// Tests the eight child boxes of a node against the segment [0, tmax] of the ray (origin o, reciprocal direction id).
// Returns the hit mask: bit 24 + (slot XOR octinv) for an inner child that is hit (so the highest set bit is the child to visit first),
// bits offset .. offset + count - 1 for the primitives of a leaf child that is hit.
// Error budget (pt_bvh8.h): the boxes carry one grid step of margin on every side and a step is at least 8 u D, the computed slab
// distances are off by less than one step.
PTB_HD uint32_t bvh8_node_hits(const Bvh8Node &n, float ox, float oy, float oz, float idx, float idy, float idz, float tmax, uint32_t octinv,
                               const uint32_t magic /* 0x4B000000 from a place the compiler cannot see through: a kernel parameter */) {
    const float px = bvh8_bits_to_float(n.w[0]), py = bvh8_bits_to_float(n.w[1]), pz = bvh8_bits_to_float(n.w[2]);
    const uint32_t e = n.w[3];
    // reciprocal direction in grid steps, and the origin term with the 2^23 of the byte conversion folded in
    const float ax = idx * bvh8_bits_to_float((e & 0xffu) << 23);
    const float ay = idy * bvh8_bits_to_float(((e >> 8) & 0xffu) << 23);
    const float az = idz * bvh8_bits_to_float(((e >> 16) & 0xffu) << 23);
    const float bx = fmaf(-8388608.0f, ax, (px - ox) * idx);
    const float by = fmaf(-8388608.0f, ay, (py - oy) * idy);
    const float bz = fmaf(-8388608.0f, az, (pz - oz) * idz);
    // near / far plane bytes by the sign of the reciprocal direction (w[8..9] qlo_x, [10..11] qlo_y, [12..13] qlo_z, [14..19] qhi)
    const bool xp = idx >= 0.0f, yp = idy >= 0.0f, zp = idz >= 0.0f;
    uint32_t hits = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t nx = xp ? n.w[8 + half] : n.w[14 + half], fx = xp ? n.w[14 + half] : n.w[8 + half];
        const uint32_t ny = yp ? n.w[10 + half] : n.w[16 + half], fy = yp ? n.w[16 + half] : n.w[10 + half];
        const uint32_t nz = zp ? n.w[12 + half] : n.w[18 + half], fz = zp ? n.w[18 + half] : n.w[12 + half];
        const uint32_t meta4 = n.w[6 + half];
        const uint32_t bits4 = (meta4 >> 5) & 0x07070707u, at4 = meta4 & 0x1f1f1f1fu;  // per child: unary primitive count (1 = inner), bit position
#define PTB_BVH8_CHILD(k)                                                                                                  \
        {                                                                                                                  \
            const float t0x = fmaf(bvh8_byte_as_biased_float<k>(nx, magic), ax, bx), t1x = fmaf(bvh8_byte_as_biased_float<k>(fx, magic), ax, bx); \
            const float t0y = fmaf(bvh8_byte_as_biased_float<k>(ny, magic), ay, by), t1y = fmaf(bvh8_byte_as_biased_float<k>(fy, magic), ay, by); \
            const float t0z = fmaf(bvh8_byte_as_biased_float<k>(nz, magic), az, bz), t1z = fmaf(bvh8_byte_as_biased_float<k>(fz, magic), az, bz); \
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));                                                     \
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));                                                     \
            /* branch-free: an empty slot's box is inverted and never hit; an inner child sets bit 24 + slot for now;   */ \
            /* with one primitive per leaf child every used slot contributes exactly one bit                            */ \
            if (BVH8_LEAF_MAX == 1) hits |= (tn <= tf ? 1u : 0u) << ((at4 >> (8 * k)) & 0xffu);                            \
            else hits |= (tn <= tf ? (bits4 >> (8 * k)) & 0xffu : 0u) << ((at4 >> (8 * k)) & 0xffu);                       \
        }
        PTB_BVH8_CHILD(0) PTB_BVH8_CHILD(1) PTB_BVH8_CHILD(2) PTB_BVH8_CHILD(3)
#undef PTB_BVH8_CHILD
    }
    // inner children: from slot order to visiting order, bit (slot XOR octinv) -- an XOR permutation of the top byte is three
    // conditional butterfly stages
    uint32_t top = hits >> 24;
    if (octinv & 1u) top = ((top & 0xaau) >> 1) | ((top & 0x55u) << 1);
    if (octinv & 2u) top = ((top & 0xccu) >> 2) | ((top & 0x33u) << 2);
    if (octinv & 4u) top = ((top & 0xf0u) >> 4) | ((top & 0x0fu) << 4);
    hits = (hits & 0x00ffffffu) | (top << 24);
    return hits;
}

// octinv: bit a set when the ray moves towards +a (children on the -a side come first)
PTB_HD uint32_t bvh8_octinv(float idx, float idy, float idz) {
    return (idx >= 0.0f ? 1u : 0u) | (idy >= 0.0f ? 2u : 0u) | (idz >= 0.0f ? 4u : 0u);
}

}  // namespace ptb
