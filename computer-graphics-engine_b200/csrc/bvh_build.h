// bvh_build.h — host-side rebuild of the reference's bounding volume hierarchy, node for node.
//
// The reference keeps `nodes`, `primitives` and `root` private (reference src/bounding_volume_hierarchy.h:83-98)
// and BvhInterface hides the implementation pointer (src/bvh_interface.h:48), so a drop-in cannot read the tree
// out of a BvhInterface without touching a do-not-touch header.  Because ties between equal-t hits are resolved
// by visit order (src/bounding_volume_hierarchy.cpp:288-290,354-355), primary-hit parity needs the SAME primitive
// permutation and the SAME boxes.  This builder reproduces createBVH (src/bounding_volume_hierarchy.cpp:130-147)
// with splitStandard (:74-78): initial order = mesh triangles then spheres (:158-172), centroid = (a+b+c)/3 (:70-72),
// median split with std::nth_element on axis depth%3 using the comparator of bounding_volume_hierarchy.h:46-50,
// leaf when depth+1 == MAX_DEPTH(16) or a single primitive (:136), children pushed before the parent (:142-146).
// std::nth_element is called on the same libstdc++ with the same comparison results, hence the same permutation.
#pragma once
#include <cstdint>
#include <vector>

#include "cge.h"

namespace cge {

struct HostBvh {
    std::vector<cge_bvh_node> nodes;   // reference node order (post-order, root last)
    std::vector<uint32_t> prim_order;  // leaf order -> global primitive id
    uint32_t root = 0;
    uint32_t n_levels = 0;
    uint32_t n_leaves = 0;
    uint32_t max_leaf_prims = 0;
};

// Global primitive id: triangles of mesh 0, mesh 1, ... in Mesh::triangles order, then spheres.
// Returns false (and leaves out empty) when the scene has no primitives (the reference throws there:
// getBoundingBox(...).value() on an empty optional, src/bounding_volume_hierarchy.cpp:132).
bool build_reference_bvh(const cge_scene_desc& desc, HostBvh& out);

// Validate a caller-supplied tree (indices in range, leaves partition [0,N)) and derive levels/leaves.
bool adopt_bvh(const cge_scene_desc& desc, HostBvh& out);

} // namespace cge
