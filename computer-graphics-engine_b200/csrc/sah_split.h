// sah_split.h — the binned-SAH split rule of the FAST traversal tree, shared by the host builder (bvh_sah.cpp) and the GPU
// builder (bvh_sah_gpu.cu) so that both produce the SAME tree, node for node and bit for bit.
//
// The tree is specified independently of the order in which nodes are processed:
//   * a node owns a contiguous range of the primitive permutation; its bounds are the union of its primitives' boxes, its
//     centroid bounds the union of their centroids (min / max under the total order of sah_key, so -0 < +0 and the result
//     does not depend on the order of accumulation);
//   * one primitive -> leaf.  Otherwise every axis with a positive centroid extent is binned into kSahBins bins,
//     bin = clamp(int((c - cb.lo) * (kSahBins / extent))), and the split (axis, bin) of minimum
//     area(left) * count(left) + area(right) * count(right) is taken, axes in order x, y, z, bins ascending, first minimum
//     wins; with <= kSahMaxLeaf primitives the node becomes a leaf when that is not more expensive than splitting;
//   * the split is a STABLE partition (primitives with bin <= split keep their relative order, then the others);
//   * no axis with a positive extent (coincident centroids) or depth >= kSahMaxDepth: leaf if <= kSahMaxLeaf primitives, else
//     the range is cut in the middle without reordering;
//   * inner nodes are numbered in depth-first pre-order (node, left subtree, right subtree).
// All arithmetic is fp32 without contraction (host: -ffp-contract=off, device: -fmad=false).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CGE_SAH_HD __host__ __device__ __forceinline__
#else
#define CGE_SAH_HD inline
#endif

#ifndef CGE_SAH_MAX_LEAF
#define CGE_SAH_MAX_LEAF 4
#endif
#ifndef CGE_SAH_CT
#define CGE_SAH_CT 1.0f
#endif

namespace cge {

constexpr int kSahBins = 16;
constexpr uint32_t kSahMaxLeaf = CGE_SAH_MAX_LEAF; // <= 8 (3 bits in the packed leaf reference)
constexpr uint32_t kSahMaxDepth = 56;
constexpr float kSahTraversalCost = CGE_SAH_CT; // cost of one inner-node visit relative to one primitive test
constexpr float kSahFltMax = 3.402823466e+38f;

// Monotone map float -> uint32 (total order: -inf < ... < -0 < +0 < ... < +inf): min / max of floats become min / max of
// unsigned integers, which is what the GPU builder's atomicMin / atomicMax need and what makes ties (+-0) order independent.
CGE_SAH_HD uint32_t sah_key(float f)
{
    union {
        float f;
        uint32_t u;
    } v;
    v.f = f;
    return (v.u & 0x80000000u) ? ~v.u : (v.u | 0x80000000u);
}
CGE_SAH_HD float sah_unkey(uint32_t k)
{
    union {
        float f;
        uint32_t u;
    } v;
    v.u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return v.f;
}
CGE_SAH_HD float sah_min(float a, float b) { return sah_key(b) < sah_key(a) ? b : a; }
CGE_SAH_HD float sah_max(float a, float b) { return sah_key(b) > sah_key(a) ? b : a; }

struct SahBox {
    float lo[3], hi[3];
};
CGE_SAH_HD SahBox sah_empty_box()
{
    SahBox b;
    for (int k = 0; k < 3; k++) {
        b.lo[k] = kSahFltMax;
        b.hi[k] = -kSahFltMax;
    }
    return b;
}
CGE_SAH_HD void sah_grow(SahBox& b, const float* lo, const float* hi)
{
    for (int k = 0; k < 3; k++) {
        b.lo[k] = sah_min(b.lo[k], lo[k]);
        b.hi[k] = sah_max(b.hi[k], hi[k]);
    }
}
CGE_SAH_HD float sah_area(const SahBox& b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f))
        return 0.0f;
    return 2.0f * ((dx * dy + dy * dz) + dz * dx);
}

// primitive box and centroid from three vertex positions
CGE_SAH_HD void sah_triangle_bounds(const float* a, const float* b, const float* c, float* lo, float* hi, float* centroid)
{
    for (int k = 0; k < 3; k++) {
        lo[k] = sah_min(sah_min(a[k], b[k]), c[k]);
        hi[k] = sah_max(sah_max(a[k], b[k]), c[k]);
        centroid[k] = 0.5f * (lo[k] + hi[k]);
    }
}

CGE_SAH_HD float sah_bin_scale(float extent) { return float(kSahBins) / extent; }
CGE_SAH_HD int sah_bin_of(float c, float cbLo, float scale)
{
    int b = int((c - cbLo) * scale);
    b = b < 0 ? 0 : b;
    return b > kSahBins - 1 ? kSahBins - 1 : b;
}

// Bins are passed as any type with   SahBox box(int axis, int bin) const   and   uint32_t count(int axis, int bin) const:
// per axis, the box of the primitive bounds and the primitive count of every bin (host: plain arrays, SahBinsHost below;
// GPU: the atomically accumulated keys in global memory).
struct SahBinsHost {
    SahBox b[3][kSahBins];
    uint32_t n[3][kSahBins];
    CGE_SAH_HD SahBox box(int axis, int bin) const { return b[axis][bin]; }
    CGE_SAH_HD uint32_t count(int axis, int bin) const { return n[axis][bin]; }
};

enum { kSahLeaf = 0, kSahSplitBin = 1, kSahSplitMiddle = 2 };
struct SahDecision {
    int kind;    // kSahLeaf / kSahSplitBin / kSahSplitMiddle
    int axis;    // kSahSplitBin: primitives with sah_bin_of(c[axis], lo, scale) <= bin go left
    int bin;
    float lo, scale;
};

// n primitives, node bounds, centroid bounds, bins of every axis with a positive centroid extent (other axes are ignored)
template <typename Bins>
CGE_SAH_HD SahDecision sah_decide(uint32_t n, uint32_t depth, const SahBox& bounds, const SahBox& cb, const Bins& bins)
{
    SahDecision d;
    d.kind = kSahLeaf;
    d.axis = d.bin = -1;
    d.lo = d.scale = 0.0f;
    if (n <= 1)
        return d;
    if (depth < kSahMaxDepth) {
        float bestCost = kSahFltMax;
        for (int axis = 0; axis < 3; axis++) {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.0f))
                continue;
            float rightArea[kSahBins];
            uint32_t rightCount[kSahBins];
            SahBox acc = sah_empty_box();
            uint32_t cnt = 0;
            for (int b = kSahBins - 1; b > 0; b--) {
                const uint32_t c = bins.count(axis, b);
                if (c) {
                    const SahBox bb = bins.box(axis, b);
                    sah_grow(acc, bb.lo, bb.hi);
                }
                cnt += c;
                rightArea[b] = sah_area(acc);
                rightCount[b] = cnt;
            }
            acc = sah_empty_box();
            cnt = 0;
            for (int b = 0; b < kSahBins - 1; b++) {
                const uint32_t c = bins.count(axis, b);
                if (c) {
                    const SahBox bb = bins.box(axis, b);
                    sah_grow(acc, bb.lo, bb.hi);
                }
                cnt += c;
                if (cnt == 0 || rightCount[b + 1] == 0)
                    continue;
                const float cost = sah_area(acc) * float(cnt) + rightArea[b + 1] * float(rightCount[b + 1]);
                if (cost < bestCost) {
                    bestCost = cost;
                    d.axis = axis;
                    d.bin = b;
                }
            }
        }
        if (d.axis >= 0) {
            const float area = sah_area(bounds);
            const float parentArea = area > 1e-30f ? area : 1e-30f;
            const float splitCost = kSahTraversalCost + bestCost / parentArea;
            if (n <= kSahMaxLeaf && float(n) <= splitCost)
                return d; // leaf
            d.kind = kSahSplitBin;
            d.lo = cb.lo[d.axis];
            d.scale = sah_bin_scale(cb.hi[d.axis] - cb.lo[d.axis]);
            return d;
        }
    }
    d.axis = d.bin = -1;
    d.kind = n <= kSahMaxLeaf ? kSahLeaf : kSahSplitMiddle;
    return d;
}

} // namespace cge
