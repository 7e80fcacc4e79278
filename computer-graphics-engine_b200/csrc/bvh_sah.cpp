// bvh_sah.cpp — binned surface-area-heuristic BVH build (host), see bvh_sah.h.
#include "bvh_sah.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace cge {
namespace {

#ifndef CGE_SAH_MAX_LEAF
#define CGE_SAH_MAX_LEAF 4
#endif
#ifndef CGE_SAH_CT
#define CGE_SAH_CT 1.0f
#endif
constexpr int kBins = 16;
constexpr uint32_t kMaxLeaf = CGE_SAH_MAX_LEAF;   // <= 8 (3 bits in the packed leaf reference)
constexpr uint32_t kMaxDepth = 56;
constexpr float kTraversalCost = CGE_SAH_CT; // cost of one inner-node visit relative to one primitive test

struct Box {
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    void grow(const float* l, const float* h)
    {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], l[k]);
            hi[k] = std::max(hi[k], h[k]);
        }
    }
    void grow(const Box& b) { grow(b.lo, b.hi); }
    float area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f))
            return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct Prim {
    uint32_t id;
    float c[3];
    float lo[3], hi[3];
};

struct Builder {
    std::vector<Prim> prims;
    FastBvh* out;

    uint32_t make_leaf(uint32_t beg, uint32_t end)
    {
        out->n_leaves++;
        return fast_leaf_ref(beg, end - beg);
    }

    // returns the packed reference of the subtree over [beg,end) and its bounds
    uint32_t build(uint32_t beg, uint32_t end, uint32_t depth, Box& bounds)
    {
        out->depth = std::max(out->depth, depth + 1);
        Box cb; // centroid bounds
        bounds = Box {};
        for (uint32_t i = beg; i < end; i++) {
            bounds.grow(prims[i].lo, prims[i].hi);
            cb.grow(prims[i].c, prims[i].c);
        }
        const uint32_t n = end - beg;
        if (n == 1)
            return make_leaf(beg, end);

        uint32_t mid = 0;
        bool haveSplit = false;
        if (depth < kMaxDepth) {
            float bestCost = FLT_MAX;
            int bestAxis = -1, bestBin = -1;
            for (int axis = 0; axis < 3; axis++) {
                const float ext = cb.hi[axis] - cb.lo[axis];
                if (!(ext > 0.0f))
                    continue;
                const float scale = float(kBins) / ext;
                Box binBox[kBins];
                uint32_t binCount[kBins] = {};
                for (uint32_t i = beg; i < end; i++) {
                    int b = int((prims[i].c[axis] - cb.lo[axis]) * scale);
                    b = std::min(std::max(b, 0), kBins - 1);
                    binBox[b].grow(prims[i].lo, prims[i].hi);
                    binCount[b]++;
                }
                float rightArea[kBins];
                uint32_t rightCount[kBins];
                Box acc;
                uint32_t cnt = 0;
                for (int b = kBins - 1; b > 0; b--) {
                    acc.grow(binBox[b]);
                    cnt += binCount[b];
                    rightArea[b] = acc.area();
                    rightCount[b] = cnt;
                }
                acc = Box {};
                cnt = 0;
                for (int b = 0; b < kBins - 1; b++) {
                    acc.grow(binBox[b]);
                    cnt += binCount[b];
                    if (cnt == 0 || rightCount[b + 1] == 0)
                        continue;
                    const float cost = acc.area() * float(cnt) + rightArea[b + 1] * float(rightCount[b + 1]);
                    if (cost < bestCost) {
                        bestCost = cost;
                        bestAxis = axis;
                        bestBin = b;
                    }
                }
            }
            if (bestAxis >= 0) {
                const float parentArea = std::max(bounds.area(), 1e-30f);
                const float splitCost = kTraversalCost + bestCost / parentArea;
                if (n <= kMaxLeaf && float(n) <= splitCost)
                    return make_leaf(beg, end);
                const float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
                const float scale = float(kBins) / ext;
                const float lo = cb.lo[bestAxis];
                auto it = std::partition(prims.begin() + beg, prims.begin() + end, [&](const Prim& p) {
                    int b = int((p.c[bestAxis] - lo) * scale);
                    b = std::min(std::max(b, 0), kBins - 1);
                    return b <= bestBin;
                });
                mid = uint32_t(it - prims.begin());
                haveSplit = mid > beg && mid < end;
            }
        }
        if (!haveSplit) {
            if (n <= kMaxLeaf)
                return make_leaf(beg, end);
            // coincident centroids or depth cap: balanced split on the widest axis of the primitive bounds
            int axis = 0;
            for (int k = 1; k < 3; k++)
                if (bounds.hi[k] - bounds.lo[k] > bounds.hi[axis] - bounds.lo[axis])
                    axis = k;
            mid = beg + n / 2;
            std::nth_element(prims.begin() + beg, prims.begin() + mid, prims.begin() + end,
                [axis](const Prim& a, const Prim& b) { return a.c[axis] < b.c[axis]; });
        }
        const uint32_t self = uint32_t(out->nodes.size());
        out->nodes.emplace_back();
        Box lb, rb;
        const uint32_t l = build(beg, mid, depth + 1, lb);
        const uint32_t r = build(mid, end, depth + 1, rb);
        FastNode& nd = out->nodes[self];
        for (int k = 0; k < 3; k++) {
            nd.l_lo[k] = lb.lo[k], nd.l_hi[k] = lb.hi[k];
            nd.r_lo[k] = rb.lo[k], nd.r_hi[k] = rb.hi[k];
        }
        nd.left = l;
        nd.right = r;
        return self;
    }
};

} // namespace

bool build_sah_bvh(const cge_scene_desc& d, FastBvh& out)
{
    out = FastBvh {};
    Builder b;
    b.out = &out;
    b.prims.reserve(size_t(d.n_triangles) + d.n_spheres);
    uint32_t gid = 0;
    for (uint32_t m = 0; m < d.n_meshes; m++) {
        const cge_mesh_desc& md = d.meshes[m];
        const cge_vertex* verts = d.vertices + md.vertex_offset;
        for (uint32_t t = 0; t < md.triangle_count; t++, gid++) {
            const uint32_t* idx = d.triangles + 3 * size_t(md.triangle_offset + t);
            Prim p;
            p.id = gid;
            for (int k = 0; k < 3; k++) {
                const float a = verts[idx[0]].position[k], bb = verts[idx[1]].position[k], c = verts[idx[2]].position[k];
                p.lo[k] = std::min({ a, bb, c });
                p.hi[k] = std::max({ a, bb, c });
                p.c[k] = 0.5f * (p.lo[k] + p.hi[k]);
            }
            b.prims.push_back(p);
        }
    }
    for (uint32_t s = 0; s < d.n_spheres; s++, gid++) {
        const cge_sphere_desc& sd = d.spheres[s];
        Prim p;
        p.id = gid;
        const float r = std::fabs(sd.radius);
        for (int k = 0; k < 3; k++) {
            p.c[k] = sd.center[k];
            p.lo[k] = sd.center[k] - r;
            p.hi[k] = sd.center[k] + r;
        }
        b.prims.push_back(p);
    }
    if (b.prims.empty())
        return false;
    // primitives with NaN coordinates can never be hit; keep them out of the bounds but in the order
    out.nodes.reserve(b.prims.size());
    Box root;
    out.root = b.build(0, uint32_t(b.prims.size()), 0, root);
    out.prim_order.resize(b.prims.size());
    for (size_t i = 0; i < b.prims.size(); i++)
        out.prim_order[i] = b.prims[i].id;
    return true;
}

} // namespace cge
