// bvh_sah.h — the traversal tree used by CGE_TRAVERSAL_FAST.
//
// The reference's closest hit is the global minimum-t hit over every primitive whose leaf the ray reaches; its own
// tree never culls by t (reference src/bounding_volume_hierarchy.cpp:334-352), so the RESULT does not depend on the
// tree shape except through (a) equal-t ties, resolved by the reference visit order, and (b) rays that graze a box
// face within rounding.  (a) is carried explicitly as a per-primitive visit rank; (b) is made one-sided by testing
// boxes with a small conservative slack.  That leaves the fast path free to walk a much better tree than the
// reference's median-split / MAX_DEPTH-16 tree (27 triangles per leaf on the dragon): a binned-SAH binary BVH with
// at most 4 primitives per leaf, built once at scene creation - on the GPU (bvh_sah_gpu.cu), with the host builder
// (bvh_sah.cpp) as its checker and fallback; sah_split.h specifies the tree both must produce.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cge.h"

namespace cge {

// packed child reference: bit 31 = leaf; leaf: bits 28..30 = count-1, bits 0..27 = first primitive (fast leaf order);
// inner: node index.
constexpr uint32_t kFastLeafBit = 0x80000000u;
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t fast_leaf_ref(uint32_t first, uint32_t count) { return kFastLeafBit | ((count - 1u) << 28) | first; }

struct FastNode { // 64 bytes, same row layout as the reference-order nodes (dev_scene.h)
    float l_lo[3], l_hi[3];
    float r_lo[3], r_hi[3];
    uint32_t left, right;
};

struct FastBvh {
    std::vector<FastNode> nodes;
    std::vector<uint32_t> prim_order; // fast leaf order -> global primitive id
    uint32_t root = 0;                // packed ref (a single-leaf scene has no inner node)
    uint32_t depth = 0;
    uint32_t n_leaves = 0;
};

// Host build (bvh_sah.cpp): the checker of the GPU builder and the fallback for scenes it does not take.
bool build_sah_bvh(const cge_scene_desc& desc, FastBvh& out);

// GPU build (bvh_sah_gpu.cu) on the current CUDA device: the same tree, node for node (sah_split.h is the shared spec).
// sah_gpu_supported: triangles only, finite vertex coordinates.  build_ms: device time of the build kernels.
bool sah_gpu_supported(const cge_scene_desc& desc);
bool build_sah_bvh_gpu(const cge_scene_desc& desc, FastBvh& out, float* build_ms, std::string* err);

} // namespace cge
