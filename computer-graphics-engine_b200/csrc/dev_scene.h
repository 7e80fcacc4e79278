// dev_scene.h — layout of the flattened scene + BVHs resident in HBM (all arrays 16-byte aligned float4 rows so
// every fetch on the traversal path is one LDG.128).
//
// Two trees index the same primitives:
//   * the REFERENCE-ORDER tree (nodes/tris): node for node the reference's median-split tree, walked literally by
//     CGE_TRAVERSAL_REFERENCE (exhaustive, right-first, exact libIntersect box arithmetic);
//   * the FAST tree (fnodes/ftris): binned-SAH, <= 4 primitives per leaf, walked near-first with t culling by
//     CGE_TRAVERSAL_FAST (bvh_sah.h explains why the result is the same).
//
//   nodes / fnodes   4 x float4 per INNER node (64 B).  A node carries the boxes of BOTH children, so one visit = one
//             64-byte fetch = two box tests, as in the reference's inner-node branch
//             (reference src/bounding_volume_hierarchy.cpp:331-355):
//               q0 = L.lower.xyz, L.upper.x     q1 = L.upper.yz, R.lower.xy
//               q2 = R.lower.z, R.upper.xyz     q3 = child references (bit patterns)
//             reference-order tree: q3 = { left ref, right ref, left count, right count }; count == 0 -> ref is an
//             inner-node index, count > 0 -> leaf holding primitives [ref, ref+count) of `tris`.
//             fast tree: q3 = { left, right, eL, eR }: child references packed as in bvh_sah.h (bit 31 leaf, bits 28..30
//             count-1) and the largest extent (max over the axes of upper - lower) of the left / right child box.
//   qnodes    2 x uint4 per inner node of the fast tree (32 B = one sector): { Lx, Ly, Lz, left } { Rx, Ry, Rz, right }, each
//             box word = lower | upper << 16 of one axis.  A 16-bit value is 0x8000 | q with q on a 15-bit grid over the
//             (padded) scene bounds, so that ONE byte permute with the constant 0x3F000000 turns it into the float
//             m = 1 + q / 32768 and the plane is qlo + m * qext; the slab test folds that affine map into the per-ray FFMA
//             constants (trace.cuh).  Boxes are rounded outwards by at least one grid step, which also covers the rounding
//             of that arithmetic.  Shadow rays only need a conservative gate and are bound by L1 traffic (4 x LDG.128 per
//             visit with up to 32 different nodes per warp): half the bytes per visit.
//   tris / ftris     6 x float4 per primitive (96 B), in the LEAF ORDER of the respective tree:
//               r0 = n.xyz, D                 plane of trianglePlane (libIntersect I1), bit-identical: same ops, no FMA
//               r1 = v0.xyz, e0.x             e0 = cross(v2 - v0, n)   first edge test of pointInTriangle (I3)
//               r2 = e0.yz, v1.xy
//               r3 = v1.z, e1.xyz             e1 = cross(v0 - v1, n)
//               r4 = v2.xyz, e2.x             e2 = cross(v1 - v2, n)
//               r5 = e2.yz, bits{rank}, bits{global primitive id | sphere << 31}
//             rank = position of the primitive in the reference's exhaustive right-first DFS visit order; among
//             equal-t hits the reference keeps the LAST visited one (:288-290), i.e. the largest rank.
//             sphere records: r1 = center.xyz, radius and r0 = NaN so the plane test rejects.
//   shade     5 x float4 per primitive (80 B), indexed by GLOBAL primitive id, touched once per closest hit:
//               s0 = n0.xyz, uv0.x   s1 = n1.xyz, uv0.y   s2 = n2.xyz, uv1.x   s3 = uv1.y, uv2.xy, 0
//               s4 = bits{ material id, 0, 0, 0 }
//   materials 3 x float4 per material: kd.xyz, shininess | ks.xyz, transparency | bits{texture id,0,0,0}
//   textures  int4 per texture: width, height, texel offset, 0;  texels: packed float RGB
//   lights    24 floats per light: bits{type}, v[21], 0, 0   (cge_light_desc, include/cge.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cge {

constexpr int kNodeRows = 4;
constexpr int kTriRows = 6;
constexpr int kShadeRows = 5;
constexpr int kMaterialRows = 3;
constexpr int kLightFloats = 24;
constexpr int kMaxRayDepth = 15;
constexpr int kTileW = 8, kTileH = 4; // one warp = one 8x4 pixel tile
constexpr uint32_t kSphereBit = 0x80000000u;

struct DevScene {
    const float4* nodes;
    const float4* tris;
    const float4* fnodes;
    const float4* onodes; // eight copies of fnodes for the shadow rays, copy o (bit 0 / 1 / 2 = the ray moves towards -x / -y / -z) at
                          // node index o * n_fnodes: q0 = L.near.xyz, L.far.x  q1 = L.far.yz, R.near.xy  q2 = R.near.z, R.far.xyz with
                          // near = (direction negative on that axis ? upper : lower); q3 as in fnodes, but INNER child references
                          // are offset by o * n_fnodes, so a walk that starts in copy o stays in it
    uint32_t n_fnodes;    // inner nodes of the fast tree
    const uint4* qnodes; // the FAST tree again, 32 B per inner node, boxes on a 15-bit grid (below): walked by shadow rays
    const float4* f4nodes; // the FAST tree collapsed to 4 children per node, 8 x float4 (128 B) per node, walked by the shadow rays:
                           //   rows 0..5 = lower.x[4], lower.y[4], lower.z[4], upper.x[4], upper.y[4], upper.z[4] of the children,
                           //   row 6 = the 4 child references (packed as in fnodes; 0x7fffffff = no child), row 7 unused
    uint32_t f4root;       // packed reference of the 4-wide tree's root
    const float4* ftris;
    const float4* shade;
    const float4* materials;
    const int4* textures;
    const float* texels;
    const float* lights;
    uint32_t n_lights;
    uint32_t n_prims;
    uint32_t root_ref, root_count; // reference-order tree root, same encoding as a child slot
    uint32_t froot;                // fast tree root (packed)
    uint32_t has_spheres;
    // FAST traversal of scenes with spheres / without enableAccelStructure (trace.cuh sphere_pass_*): the fast tree holds the
    // triangles only; the (few) spheres are tested one by one, each behind the chain of REFERENCE-tree boxes the reference's
    // traversal tests before it reaches the sphere's leaf (the archive's sphere test assumes a unit direction, so for shadow rays
    // WHETHER it is called decides the result)
    const float4* sph_rows;      // kTriRows rows per sphere, as in `tris` (r1 = centre, radius; r2.x = bits{position in the reference's
                                 // primitive vector}; r5.z = bits{visit rank}; r5.w = bits{global id | sphere bit})
    const float4* sph_boxes;     // 2 rows per box: lower.xyz, upper.xyz
    const uint32_t* sph_box_off; // n_sph + 1 offsets into sph_boxes (in boxes)
    const uint32_t* fpos;        // fast-tree primitive index -> position in the reference's primitive vector (the tie rank when
                                 // enableAccelStructure is off: the reference then loops over that vector, :303-305)
    uint32_t n_sph, n_ftris;
    uint32_t noaccel;            // this frame has enableAccelStructure off
    float qlo[3], qext[3]; // quantisation grid of qnodes: coordinate = qlo + m * qext, m in [1, 2)
    uint32_t cull_zero_shading; // FAST traversal: skip the shadow ray of a light sample with n.l <= 0 (shade.cuh shading_is_zero)
};

struct DevCamera {
    float ox, oy, oz;
    float qw, qx, qy, qz;
    float half_w, half_h;
};

struct DevParams;
__host__ __device__ inline uint32_t part_tile_of(uint32_t unit, uint32_t index, uint32_t count, uint32_t k)
{
    const uint32_t j = k / unit;
    return (index + j * count) * unit + (k - j * unit); // entry k of the part's tile list -> tile id
}
__host__ __device__ inline uint32_t part_entry_of(uint32_t unit, uint32_t index, uint32_t count, uint32_t tile)
{
    const uint32_t u = tile / unit;
    return (u - index) / count * unit + (tile - u * unit); // inverse of part_tile_of for a tile of this part
}

struct DevParams {
    int32_t width, height;
    uint32_t features;
    int32_t ray_depth;
    int32_t segment_samples, parallelogram_samples;
    float segment_recip, parallelogram_recip; // 1 / samples when that is exact (power of two), else 0 (shade.cuh div_by_count)
    uint32_t seed;
    uint32_t draws_per_hit;       // rand() draws one computeLightContribution call makes (0 -> deterministic fold)
    uint32_t shadow_rays_per_hit; // shadow rays one computeLightContribution call traces
    uint32_t samples_per_hit;     // shading evaluations per call (point: 1, segment: N, parallelogram: N*N)
    uint32_t part_index, part_count;
    // the image partition deals UNITS of part_unit consecutive tiles (row-major tile numbering): unit u belongs to part u % part_count.
    // 1 = single 8x4 tiles (cge_render's documented partition), n_tiles_x = whole tile rows (cge_render_distributed: a rank's pixels
    // are then runs of 4 complete image rows, contiguous in the frame)
    uint32_t part_unit;
    // output layout.  0: Screen::pixels() order over the whole frame, index (H-1-y)*W + x (src/screen.cpp:45).  n > 0 (a rank of
    // cge_render_distributed; part_unit = n_tiles_x, n = the part's tile rows): COMPACT - only the part's own image rows, tile row
    // by tile row in frame order (render_kernels.cuh out_pixel_index), so that the part's share of the frame is one strided copy
    uint32_t compact_units;
    // this launch renders entries [tile_first, tile_first + tile_count) of the partition's tile list (entry k = tile
    // part_index + k * part_count); the whole list unless the frame is rendered in bands (cge_api.cu cge_render)
    uint32_t tile_first, tile_count;
    // dynamic tile dealing across GPUs (cge_render_distributed with CGE_FLAG_DYNAMIC_TILES): when set, the launch renders the chunk a
    // grant kernel took from the frame's shared counter just before it - { tile_first, tile_count, part_index } in device memory
    // (tile_count = 0: the pool was empty, the launch has nothing to do).  The fields above then only size the launch.
    const uint32_t* grant;
    uint32_t n_tiles_x, n_tiles_y;
    uint32_t levels;          // ray_depth + 1 when recursive, else 1
    uint32_t units_per_lane;  // direct-lighting evaluations a pixel can need: levels (fold) or 2^levels - 1
    uint32_t debug_cycles;    // CGE_DEV_FLAG_DEBUG_CYCLES
    uint32_t shade_mode;      // wavefront: 1 = wf_shade_kernel<false> traces the shadow rays itself, 2 = visibility bytes (wavefront.cuh)
    uint32_t aa_side;              // raysPerPixelSide when extra.enableMultipleRaysPerPixel is set, else 0
    uint32_t chain_unsplit;        // host-side hint: keep the chain stage in one kernel (cge_render_distributed with the peer frame)
    uint32_t vis_cull;             // wavefront: the light-hull pre-pass ran, the shadow-ray kernel traces the hard lists (wavefront.cuh)
    uint32_t cull_budget;          // inner nodes one hull walk of the pre-pass may visit
    uint32_t packet_budget, packet_leaf_cost; // shadow_packet.cuh: node visits a hull walk may spend, and what a leaf counts for
    float packet_fat;              // shadow_packet.cuh: a child box is deferred to the per-ray phase when the packet's hull is wider
                                   // than (the box's largest extent) / packet_fat where it enters the box
};

} // namespace cge
