// bvh_build.cpp — see bvh_build.h.  Host only; compiled with -ffp-contract=off.
#include "bvh_build.h"

#include <algorithm>
#include <cfloat>
#include <thread>

namespace cge {
namespace {

struct Prim {
    uint32_t id;      // global primitive id
    float center[3];  // centroid used by the split comparators
    float lo[3], hi[3];
};

// nodes a subtree over n primitives at `depth` will have: the shape of the median-split tree depends on n and depth only
uint32_t subtree_nodes(uint32_t n, uint32_t depth)
{
    if (depth + 1 == 16 || n == 1)
        return 1;
    return 1 + subtree_nodes(n / 2, depth + 1) + subtree_nodes(n - n / 2, depth + 1);
}

struct Builder {
    std::vector<Prim> prims;
    HostBvh* out;

    // Builds the subtree over [beg, end) into nodes [base, base + subtree_nodes) in the reference's order (children before the
    // parent, left subtree first: src/bounding_volume_hierarchy.cpp:142-146 pushes both children, then the node) and returns
    // the subtree root's index.  The two halves of a split touch disjoint ranges of `prims` and of `nodes`, so the top
    // kParallelDepth levels hand their left half to another thread: same calls of std::nth_element on the same data as the
    // sequential recursion, hence the same permutation.
    static constexpr uint32_t kParallelDepth = 4;
    uint32_t create(uint32_t beg, uint32_t end, uint32_t depth, uint32_t base)
    {
        cge_bvh_node node {};
        node.depth = depth;
        node.beg = beg;
        node.end = end;
        const uint32_t self = base + subtree_nodes(end - beg, depth) - 1;
        if (depth + 1 == 16 || beg + 1 == end) {
            for (int k = 0; k < 3; k++) {
                node.lower[k] = prims[beg].lo[k];
                node.upper[k] = prims[beg].hi[k];
            }
            for (uint32_t i = beg + 1; i < end; i++)
                for (int k = 0; k < 3; k++) {
                    node.lower[k] = std::min(node.lower[k], prims[i].lo[k]);
                    node.upper[k] = std::max(node.upper[k], prims[i].hi[k]);
                }
            node.is_leaf = 1;
            out->nodes[self] = node;
            return self;
        }
        const uint32_t mid = beg + (end - beg) / 2;
        const int axis = int(depth % 3);
        std::nth_element(prims.begin() + beg, prims.begin() + mid, prims.begin() + end,
            [axis](const Prim& a, const Prim& b) { return a.center[axis] < b.center[axis]; });
        const uint32_t leftNodes = subtree_nodes(mid - beg, depth + 1);
        uint32_t left = 0, right = 0;
        if (depth < kParallelDepth && end - beg > 4096) {
            std::thread worker([&]() { left = create(beg, mid, depth + 1, base); });
            right = create(mid, end, depth + 1, base + leftNodes);
            worker.join();
        } else {
            left = create(beg, mid, depth + 1, base);
            right = create(mid, end, depth + 1, base + leftNodes);
        }
        // min/max are exact, so the union of the children's boxes equals the box over all primitives of the range
        for (int k = 0; k < 3; k++) {
            node.lower[k] = std::min(out->nodes[left].lower[k], out->nodes[right].lower[k]);
            node.upper[k] = std::max(out->nodes[left].upper[k], out->nodes[right].upper[k]);
        }
        node.is_leaf = 0;
        node.left = left;
        node.right = right;
        out->nodes[self] = node;
        return self;
    }
};

} // namespace

bool build_reference_bvh(const cge_scene_desc& d, HostBvh& out)
{
    out = HostBvh {};
    Builder b;
    b.out = &out;
    b.prims.reserve(size_t(d.n_triangles) + d.n_spheres);
    uint32_t gid = 0;
    for (uint32_t m = 0; m < d.n_meshes; m++) {
        const cge_mesh_desc& md = d.meshes[m];
        const cge_vertex* verts = d.vertices + md.vertex_offset;
        for (uint32_t t = 0; t < md.triangle_count; t++, gid++) {
            const uint32_t* idx = d.triangles + 3 * size_t(md.triangle_offset + t);
            const float* p1 = verts[idx[0]].position;
            const float* p2 = verts[idx[1]].position;
            const float* p3 = verts[idx[2]].position;
            Prim p;
            p.id = gid;
            for (int k = 0; k < 3; k++) {
                p.center[k] = ((p1[k] + p2[k]) + p3[k]) / 3.0f; // triangleCenter, bounding_volume_hierarchy.cpp:70-72
                p.lo[k] = std::min({ p1[k], p2[k], p3[k] });
                p.hi[k] = std::max({ p1[k], p2[k], p3[k] });
            }
            b.prims.push_back(p);
        }
    }
    for (uint32_t s = 0; s < d.n_spheres; s++, gid++) {
        const cge_sphere_desc& sd = d.spheres[s];
        Prim p;
        p.id = gid;
        for (int k = 0; k < 3; k++) {
            p.center[k] = sd.center[k];
            p.lo[k] = sd.center[k] - sd.radius;
            p.hi[k] = sd.center[k] + sd.radius;
        }
        b.prims.push_back(p);
    }
    if (b.prims.empty())
        return false;
    out.nodes.resize(subtree_nodes(uint32_t(b.prims.size()), 0));
    out.root = b.create(0, uint32_t(b.prims.size()), 0, 0);
    for (const cge_bvh_node& n : out.nodes) {
        out.n_levels = std::max(out.n_levels, n.depth + 1);
        if (n.is_leaf) {
            out.n_leaves++;
            out.max_leaf_prims = std::max(out.max_leaf_prims, n.end - n.beg);
        }
    }
    out.prim_order.resize(b.prims.size());
    for (size_t i = 0; i < b.prims.size(); i++)
        out.prim_order[i] = b.prims[i].id;
    return true;
}

bool adopt_bvh(const cge_scene_desc& d, HostBvh& out)
{
    out = HostBvh {};
    const uint32_t n_prims = d.n_triangles + d.n_spheres;
    if (!d.bvh_nodes || !d.bvh_prim_order || d.n_bvh_nodes == 0 || d.bvh_root >= d.n_bvh_nodes)
        return false;
    out.nodes.assign(d.bvh_nodes, d.bvh_nodes + d.n_bvh_nodes);
    out.prim_order.assign(d.bvh_prim_order, d.bvh_prim_order + n_prims);
    out.root = d.bvh_root;
    std::vector<uint8_t> seen(n_prims, 0);
    for (uint32_t i = 0; i < n_prims; i++) {
        if (out.prim_order[i] >= n_prims || seen[out.prim_order[i]])
            return false;
        seen[out.prim_order[i]] = 1;
    }
    // Walk the tree from the root (iteratively: the input is untrusted).  Every node may be reached at most once (no cycles, no
    // shared subtrees, hence at most n_bvh_nodes visits), the children of an inner node must split its primitive range exactly
    // ([beg, mid) left, [mid, end) right - the reference's createBVH, src/bounding_volume_hierarchy.cpp:138-146), the root must
    // cover [0, n_prims): the leaves reachable from the root then tile [0, n_prims) without overlap.  The depth that sizes the
    // device traversal stack is the one measured here, not the (untrusted) depth field.
    const auto& root = out.nodes[out.root];
    if (root.beg != 0 || root.end != n_prims)
        return false;
    std::vector<uint8_t> visited(out.nodes.size(), 0);
    std::vector<std::pair<uint32_t, uint32_t>> todo { { out.root, 0u } }; // (node, true depth)
    while (!todo.empty()) {
        const auto [ni, depth] = todo.back();
        todo.pop_back();
        if (visited[ni])
            return false;
        visited[ni] = 1;
        const auto& n = out.nodes[ni];
        if (n.beg >= n.end || n.end > n_prims)
            return false;
        out.n_levels = std::max(out.n_levels, depth + 1);
        if (n.is_leaf) {
            out.n_leaves++;
            out.max_leaf_prims = std::max(out.max_leaf_prims, n.end - n.beg);
            continue;
        }
        if (n.left >= d.n_bvh_nodes || n.right >= d.n_bvh_nodes || n.left == n.right)
            return false;
        const auto &l = out.nodes[n.left], &r = out.nodes[n.right];
        if (l.beg != n.beg || l.end != r.beg || r.end != n.end)
            return false;
        todo.push_back({ n.left, depth + 1 });
        todo.push_back({ n.right, depth + 1 });
    }
    return true;
}

} // namespace cge
