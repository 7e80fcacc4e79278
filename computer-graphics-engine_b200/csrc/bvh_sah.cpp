// bvh_sah.cpp — binned surface-area-heuristic BVH build on the host: the reference implementation of the tree specified in
// sah_split.h.  The GPU builder (bvh_sah_gpu.cu) produces the same tree and is the one cge_scene_create uses; this one is
// its checker (tests/test_gpu_bvh_build.py compares them node for node) and the fallback for scenes with non-finite vertex
// coordinates.
#include "bvh_sah.h"

#include <algorithm>
#include <cmath>

#include "sah_split.h"

namespace cge {
namespace {

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
    uint32_t build(uint32_t beg, uint32_t end, uint32_t depth, SahBox& bounds)
    {
        out->depth = std::max(out->depth, depth + 1);
        SahBox cb = sah_empty_box(); // centroid bounds
        bounds = sah_empty_box();
        for (uint32_t i = beg; i < end; i++) {
            sah_grow(bounds, prims[i].lo, prims[i].hi);
            sah_grow(cb, prims[i].c, prims[i].c);
        }
        const uint32_t n = end - beg;
        SahBinsHost bins;
        if (n > 1 && depth < kSahMaxDepth)
            for (int axis = 0; axis < 3; axis++) {
                const float ext = cb.hi[axis] - cb.lo[axis];
                if (!(ext > 0.0f))
                    continue;
                const float scale = sah_bin_scale(ext);
                for (int b = 0; b < kSahBins; b++) {
                    bins.b[axis][b] = sah_empty_box();
                    bins.n[axis][b] = 0;
                }
                for (uint32_t i = beg; i < end; i++) {
                    const int b = sah_bin_of(prims[i].c[axis], cb.lo[axis], scale);
                    sah_grow(bins.b[axis][b], prims[i].lo, prims[i].hi);
                    bins.n[axis][b]++;
                }
            }
        const SahDecision d = sah_decide(n, depth, bounds, cb, bins);
        if (d.kind == kSahLeaf)
            return make_leaf(beg, end);
        uint32_t mid = beg + n / 2; // kSahSplitMiddle: cut the range in the middle, order untouched
        if (d.kind == kSahSplitBin) {
            auto it = std::stable_partition(prims.begin() + beg, prims.begin() + end,
                [&](const Prim& p) { return sah_bin_of(p.c[d.axis], d.lo, d.scale) <= d.bin; });
            mid = uint32_t(it - prims.begin());
        }
        const uint32_t self = uint32_t(out->nodes.size());
        out->nodes.emplace_back();
        SahBox lb, rb;
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
            sah_triangle_bounds(verts[idx[0]].position, verts[idx[1]].position, verts[idx[2]].position, p.lo, p.hi, p.c);
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
    out.nodes.reserve(b.prims.size());
    SahBox root;
    out.root = b.build(0, uint32_t(b.prims.size()), 0, root);
    out.prim_order.resize(b.prims.size());
    for (size_t i = 0; i < b.prims.size(); i++)
        out.prim_order[i] = b.prims[i].id;
    return true;
}

} // namespace cge
