// pt_launch.h -- host-visible launch interface of pt_kernels.cu / pt_bvh_build.cu (internal, not part of the C ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_device.cuh"

namespace ptb {

constexpr int TILE_W = 8, TILE_H = 4;  // one warp = one 8x4 pixel tile
constexpr int RENDER_THREADS = 256;
constexpr int RENDER_MIN_BLOCKS = 3;
constexpr int REGEN_BATCH = 24;        // lanes that must be waiting for a camera ray before the ray-gen code runs

struct RenderArgs {
    int width, height;
    unsigned long long spp_begin, spp_count;
    unsigned long long seed;
    PhiloxKeys rk;                         // round keys of `seed` (philox_round_keys)
    float *sum_rgb;                        // W*H*3 fp32 running sum, reference index order
    int *tile_counter;                     // zeroed before each launch
    int n_tiles, tiles_x;
    int fb_zero;                           // 1: sum_rgb is to be treated as all zeros (first batch of a fresh frame: no load, no memset)
    int regen_batch;                       // lanes that must be waiting for a camera ray before the ray-gen code runs
    unsigned long long *segment_counter;   // [0] += closest-hit queries, [1] += BVH nodes fetched, [2] += BVH primitives tested
    // sample-parallel megakernel (small frames): non-null = work slots are chunks of sp_chunk samples, per-sample radiance goes to
    // sample_L[sample - spp_begin][pixel] and is added to sum_rgb in sample order by k_accumulate_samples (same launch call)
    float4 *sample_L;
    int sp_chunk, sp_n_chunks;
};

cudaError_t launch_rcp_selftest(unsigned long long *d_mismatches, int sm_count, cudaStream_t st);
cudaError_t launch_contraction_probe(float a, float b, float c, float *d_out, cudaStream_t st);
cudaError_t launch_intersect(const DScene &sc, const float *d_rays, unsigned long long n, int pw, int ph, int *d_obj, int *d_tri,
                             float *d_t, float *d_point, float *d_normal, int sm_count, cudaStream_t st);
cudaError_t launch_render(const DScene &sc, const RenderArgs &a, int sm_count, cudaStream_t st);
constexpr int MAX_PEERS = 16;
struct PeerPtrs { const float *p[MAX_PEERS]; };
cudaError_t launch_peer_reduce_resolve(const PeerPtrs &peers, int n_peers, unsigned long long first, unsigned long long n,
                                       unsigned long long spp, float *d_dst, int sm_count, cudaStream_t st);
cudaError_t launch_resolve(const float *d_sum, unsigned long long n, unsigned long long spp, float *d_mean, int sm_count,
                           cudaStream_t st);

}  // namespace ptb
