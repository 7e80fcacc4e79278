// trace.cuh — BVH traversal on the device.
//
//   trace_reference : BoundingVolumeHierarchy::intersect's traversal reproduced literally
//                     (reference src/bounding_volume_hierarchy.cpp:312-361 + getIntersecting :272-293): exhaustive DFS
//                     of the reference's own tree, both child boxes tested with ray.t = FLT_MAX through the exact
//                     libIntersect box arithmetic (I5), left pushed then right pushed (right popped first), leaf
//                     primitives ascending, every accepted primitive overwrites the winner, NO culling by t.
//   trace_fast      : the same answer from the SAH tree (bvh_sah.h): near child first, boxes culled against the
//                     best t so far, slab test with a precomputed reciprocal direction and a conservative slack,
//                     equal-t ties resolved by the reference visit rank, optional any-hit exit for shadow rays.
// Both use the SAME triangle arithmetic: the archive's plane test + three inclusive edge tests (I2-I4) on the
// precomputed rows, no FMA, IEEE division.
#pragma once
#include "dev_scene.h"
#include "intersect.cuh"

namespace cge {

constexpr int kRefStackSize = 40;
constexpr int kFastStackSize = 64;

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

#ifndef CGE_OCTANT_NODES
#define CGE_OCTANT_NODES 1 // shadow rays walk one of EIGHT copies of the fast tree's inner nodes, chosen once per ray by the signs of its
                           // direction: in copy o every slab's near plane is stored where the walk reads "near" (dev_scene.h onodes),
                           // and inner child references already point into copy o.  Removes 12 FSEL + 3 FSETP of ~66 instructions per
                           // visit; costs 8 x 64 B per inner node of HBM (C5: 8 x 24 MB)
#endif
#ifndef CGE_NODE_LD256
#define CGE_NODE_LD256 1 // 1: the shadow walk fetches a 64-byte node with two 256-bit loads (LDG.E.256, new on sm_100) instead of
                         // 3 x 128 + 1 x 64 bit: half the load instructions and L1 requests per visit
#endif
// 32 bytes (two float4 rows) in one 256-bit read-only load; p must be 32-byte aligned
__device__ __forceinline__ void ldg8(const float4* p, float4& a, float4& b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

struct Hit {
    float t;
    int prim;      // index into the tree's leaf-ordered primitive array (-1: miss)
    unsigned gid;  // global primitive id | kSphereBit
};

// One triangle candidate against the current best.  Returns true when the archive would accept it
// (0 <= t <= best and inside all three edges).  r5 is returned for rank / id.
// The part of the test after the plane numerator num = D - dot(o, n) (n = the plane normal of row 0): shared by the single-ray
// test below and by the packet walk of shadow_packet.cuh, whose rays have one origin and therefore one numerator per triangle.
__device__ __forceinline__ bool triangle_rows_tail(const float4* __restrict__ tr, const vec3 n, const float num, const vec3 o, const vec3 d,
    float best, float& tOut, float4& r5)
{
    const float den = dot(d, n);
    // IEEE division gives sign(num) xor sign(den); when they differ and the quotient cannot underflow to -0 the
    // archive's `t >= 0` is false, so the (10-instruction) division can be skipped without changing any decision.
    if ((__float_as_uint(num) ^ __float_as_uint(den)) >> 31 && fabsf(num) >= 1e-30f && fabsf(den) <= 1e6f)
        return false;
    const float t = fdiv(num, den); // I2
    if (!(t >= 0.0f))
        return false;
    if (!(best >= t))
        return false;
    const vec3 p = d * t + o;
    const float4 r1 = ldg4(tr + 1);
    const float4 r2 = ldg4(tr + 2);
    if (!(dot(v3(r1.w, r2.x, r2.y), p - v3(r1.x, r1.y, r1.z)) >= 0.0f)) // I3, archive order, short-circuit
        return false;
    const float4 r3 = ldg4(tr + 3);
    if (!(dot(v3(r3.y, r3.z, r3.w), p - v3(r2.z, r2.w, r3.x)) >= 0.0f))
        return false;
    const float4 r4 = ldg4(tr + 4);
    r5 = ldg4(tr + 5);
    if (!(dot(v3(r4.w, r5.x, r5.y), p - v3(r4.x, r4.y, r4.z)) >= 0.0f))
        return false;
    tOut = t;
    return true;
}

#ifndef CGE_TRI_LD256
#define CGE_TRI_LD256 0 // 1: a triangle's six rows are fetched as three 256-bit loads (rows 0-1, 2-3, 4-5; the record is 96 bytes and
                        // 32-byte aligned) instead of up to six 128-bit ones.  Measured on B200: slower everywhere (C5
                        // frame 13.64 -> 13.79 ms, C4 0.85 -> 0.905 ms): most tests end after the plane row, the second 16 bytes are
                        // fetched for nothing.  Off.
#endif
__device__ __forceinline__ bool triangle_rows_hit(const float4* __restrict__ tr, const vec3 o, const vec3 d, float best, float& tOut,
    float4& r5)
{
#if CGE_TRI_LD256
    float4 r0, r1;
    ldg8(tr, r0, r1);
    const vec3 n = v3(r0.x, r0.y, r0.z);
    const float num = fsub(r0.w, dot(o, n)), den = dot(d, n);
    if ((__float_as_uint(num) ^ __float_as_uint(den)) >> 31 && fabsf(num) >= 1e-30f && fabsf(den) <= 1e6f)
        return false; // (see triangle_rows_tail)
    const float t = fdiv(num, den); // I2
    if (!(t >= 0.0f))
        return false;
    if (!(best >= t))
        return false;
    const vec3 p = d * t + o;
    float4 r2, r3;
    ldg8(tr + 2, r2, r3);
    if (!(dot(v3(r1.w, r2.x, r2.y), p - v3(r1.x, r1.y, r1.z)) >= 0.0f)) // I3, archive order, short-circuit
        return false;
    if (!(dot(v3(r3.y, r3.z, r3.w), p - v3(r2.z, r2.w, r3.x)) >= 0.0f))
        return false;
    float4 r4;
    ldg8(tr + 4, r4, r5);
    if (!(dot(v3(r4.w, r5.x, r5.y), p - v3(r4.x, r4.y, r4.z)) >= 0.0f))
        return false;
    tOut = t;
    return true;
#else
    const float4 r0 = ldg4(tr);
    const vec3 n = v3(r0.x, r0.y, r0.z);
    return triangle_rows_tail(tr, n, fsub(r0.w, dot(o, n)), o, d, best, tOut, r5);
#endif
}

template <bool kSpheres, bool kCount>
__device__ Hit trace_reference(const DevScene& s, const vec3 o, const vec3 d, float tmax, unsigned& nbox, unsigned& ntri)
{
    Hit h { tmax, -1, 0u };
    if (s.n_prims == 0)
        return h;
    uint2 stack[kRefStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(s.root_ref, s.root_count);
    while (sp > 0) {
        const uint2 e = stack[--sp];
        if (e.y > 0) {
            for (unsigned i = e.x; i < e.x + e.y; i++) {
                const float4* tr = s.tris + size_t(i) * kTriRows;
                if (kCount)
                    ntri++;
                if (kSpheres) {
                    const float4 r5 = ldg4(tr + 5);
                    if (__float_as_uint(r5.w) & kSphereBit) {
                        const float4 r1 = ldg4(tr + 1);
                        Ray ray { o, d, h.t };
                        if (intersect_sphere(v3(r1.x, r1.y, r1.z), r1.w, ray, nullptr)) { // strict t < ray.t (I6)
                            h.t = ray.t;
                            h.prim = int(i);
                            h.gid = __float_as_uint(r5.w);
                        }
                        continue;
                    }
                }
                float t;
                float4 r5;
                if (triangle_rows_hit(tr, o, d, h.t, t, r5)) {
                    h.t = t;
                    h.prim = int(i);
                    h.gid = __float_as_uint(r5.w);
                }
            }
        } else {
            const float4* nd = s.nodes + size_t(e.x) * kNodeRows;
            const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
            if (kCount)
                nbox += 2;
            Ray ray { o, d, FLT_MAX };
            const bool hitL = intersect_aabb(v3(q0.x, q0.y, q0.z), v3(q0.w, q1.x, q1.y), ray);
            ray.t = FLT_MAX;
            const bool hitR = intersect_aabb(v3(q1.z, q1.w, q2.x), v3(q2.y, q2.z, q2.w), ray);
            if (hitL)
                stack[sp++] = make_uint2(__float_as_uint(q3.x), __float_as_uint(q3.z));
            if (hitR)
                stack[sp++] = make_uint2(__float_as_uint(q3.y), __float_as_uint(q3.w)); // popped first
        }
    }
    return h;
}

// max / min of three: one FMNMX3 on sm_100 instead of two FMNMX
__device__ __forceinline__ float max3(float a, float b, float c)
{
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- slab test with one FFMA per plane ------------------------------------------------------------------------------
// t(b) = (b - o) / d is evaluated as fma(b, inv, c) with inv = 1/d and c = -(o * inv) precomputed per ray: 6 FFMA per box instead
// of 6 FADD + 6 FMUL, and because the sign of inv tells which plane of a slab is the near one, the 6 min/max that sort
// t(lo), t(hi) become 6 selects.  The price is cancellation: fma(b, inv, c) differs from the exact quotient by up to
// 2^-23 |t| + 2^-24 |o * inv|, and the second term is not small when the ray is almost parallel to a slab (|inv| huge).
// It is absorbed PER AXIS into the constants - near planes use c - s, far planes c + s with s = 2^-22 |o * inv| - so that an
// axis with a huge |inv| widens only its own interval (by a few ulps of the plane coordinate, where it cannot matter) and
// the others keep their precision.  A global slack (tried first) let such rays into almost every nearby
// box.  As before the result only has to be conservative: it gates which leaves are reached, never a hit decision.
// d == 0 (or so small that 1/d overflows) uses inv = 1e18: t(b) is then +-huge outside the slab and ~0 inside it.
struct SlabRay {
    vec3 inv, cNear, cFar;
    bool nx, ny, nz; // inv < 0: the upper plane is the near one
};
__device__ __forceinline__ SlabRay slab_ray(const vec3 o, const vec3 d)
{
    SlabRay r;
    auto recip = [](float v) { return fabsf(v) > 1e-18f ? fdiv(1.0f, v) : 1e18f; };
    r.inv = v3(recip(d.x), recip(d.y), recip(d.z));
    const vec3 c = v3(-fmul(o.x, r.inv.x), -fmul(o.y, r.inv.y), -fmul(o.z, r.inv.z));
    const vec3 s = v3(fabsf(c.x) * 2.384185791015625e-07f, fabsf(c.y) * 2.384185791015625e-07f, fabsf(c.z) * 2.384185791015625e-07f);
    r.cNear = v3(c.x - s.x, c.y - s.y, c.z - s.z);
    r.cFar = v3(c.x + s.x, c.y + s.y, c.z + s.z);
    r.nx = r.inv.x < 0.0f, r.ny = r.inv.y < 0.0f, r.nz = r.inv.z < 0.0f;
    return r;
}
// entry / exit distance of the ray through the box [lo, hi]; the box is hit within [0, bound] iff
// ent <= ext * 1.000002f && ent <= bound (the multiplicative slack covers the relative part of the error)
__device__ __forceinline__ void slab_box(const SlabRay& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float& ent,
    float& ext)
{
    const float tnx = __fmaf_rn(r.nx ? hix : lox, r.inv.x, r.cNear.x), tfx = __fmaf_rn(r.nx ? lox : hix, r.inv.x, r.cFar.x);
    const float tny = __fmaf_rn(r.ny ? hiy : loy, r.inv.y, r.cNear.y), tfy = __fmaf_rn(r.ny ? loy : hiy, r.inv.y, r.cFar.y);
    const float tnz = __fmaf_rn(r.nz ? hiz : loz, r.inv.z, r.cNear.z), tfz = __fmaf_rn(r.nz ? loz : hiz, r.inv.z, r.cFar.z);
    ent = fmaxf(max3(tnx, tny, tnz), 0.0f);
    ext = min3(tfx, tfy, tfz);
}

// the same test on a node of the octant-sorted copies (dev_scene.h onodes): the near / far plane of every slab was picked when the
// copy was written, so the 12 selects (and the 3 sign predicates) of slab_box are gone
__device__ __forceinline__ void slab_box_sorted(const SlabRay& r, float nx, float ny, float nz, float fx, float fy, float fz, float& ent,
    float& ext)
{
    const float tnx = __fmaf_rn(nx, r.inv.x, r.cNear.x), tfx = __fmaf_rn(fx, r.inv.x, r.cFar.x);
    const float tny = __fmaf_rn(ny, r.inv.y, r.cNear.y), tfy = __fmaf_rn(fy, r.inv.y, r.cFar.y);
    const float tnz = __fmaf_rn(nz, r.inv.z, r.cNear.z), tfz = __fmaf_rn(fz, r.inv.z, r.cFar.z);
    ent = fmaxf(max3(tnx, tny, tnz), 0.0f);
    ext = min3(tfx, tfy, tfz);
}

// ---- spheres beside the fast tree (dev_scene.h) ------------------------------------------------------------------------------------
// Would the reference's traversal reach sphere i's leaf?  It tests the box of every node on the way from the root's child down
// to the leaf with ray.t = FLT_MAX (src/bounding_volume_hierarchy.cpp:334-352); without enableAccelStructure it tests no box.
__device__ __forceinline__ bool sphere_reached(const DevScene& s, unsigned i, const vec3 o, const vec3 d)
{
    if (s.noaccel)
        return true;
    for (unsigned b = __ldg(s.sph_box_off + i); b < __ldg(s.sph_box_off + i + 1); b++) {
        const float4 lo = ldg4(s.sph_boxes + 2 * size_t(b)), hi = ldg4(s.sph_boxes + 2 * size_t(b) + 1);
        Ray ray { o, d, FLT_MAX };
        if (!intersect_aabb(v3(lo.x, lo.y, lo.z), v3(hi.x, hi.y, hi.z), ray))
            return false;
    }
    return true;
}
// Closest hit: merge the spheres into the triangle result h.  A sphere is accepted only with t STRICTLY below the ray's current
// t (I6), a triangle with t <= (I4): whatever the visit order, a sphere wins exactly when its t is strictly below every triangle's,
// and among spheres of equal t the first visited one (lowest rank) stays.
__device__ __forceinline__ void sphere_pass_closest(const DevScene& s, const vec3 o, const vec3 d, Hit& h)
{
    float bestT = h.t;
    unsigned bestKey = 0xffffffffu;
    int best = -1;
    for (unsigned i = 0; i < s.n_sph; i++) {
        const float4* tr = s.sph_rows + size_t(i) * kTriRows;
        const float4 r1 = ldg4(tr + 1);
        Ray ray { o, d, h.t };
        if (!intersect_sphere(v3(r1.x, r1.y, r1.z), r1.w, ray, nullptr)) // rejects t >= h.t: the triangle (or tmax) keeps ties
            continue;
        if (!sphere_reached(s, i, o, d)) // (the cheap test first: most rays miss the sphere and never need the box chain)
            continue;
        const unsigned key = s.noaccel ? __float_as_uint(ldg4(tr + 2).x) : __float_as_uint(ldg4(tr + 5).z);
        if (ray.t < bestT || (ray.t == bestT && best >= 0 && key < bestKey)) {
            bestT = ray.t;
            bestKey = key;
            best = int(i);
        }
    }
    if (best >= 0) {
        h.t = bestT;
        h.prim = best;
        h.gid = __float_as_uint(ldg4(s.sph_rows + size_t(best) * kTriRows + 5).w);
    }
}
// Any hit with 0 <= t < tmax on a sphere the reference would test?
__device__ __forceinline__ bool sphere_pass_shadow(const DevScene& s, const vec3 o, const vec3 d, float tmax)
{
    for (unsigned i = 0; i < s.n_sph; i++) {
        const float4 r1 = ldg4(s.sph_rows + size_t(i) * kTriRows + 1);
        Ray ray { o, d, tmax };
        if (intersect_sphere(v3(r1.x, r1.y, r1.z), r1.w, ray, nullptr) && sphere_reached(s, i, o, d))
            return true;
    }
    return false;
}
// rows of a hit primitive: the fast tree's triangle rows, or the sphere rows
__device__ __forceinline__ const float4* fast_hit_rows(const DevScene& s, const Hit& h)
{
    return ((h.gid & kSphereBit) ? s.sph_rows : s.ftris) + size_t(h.prim) * kTriRows;
}
constexpr int kSphereBlocker = -2; // trace_shadow: blocked by a sphere (no triangle to remember)

// Fast tree (triangles) + the sphere pass.
#ifndef CGE_CLOSEST_OCTANT
#define CGE_CLOSEST_OCTANT 0 // closest-hit rays (camera, reflection) on the octant-sorted node copies too: measured within noise
                             // (C5 chain stage 1.31 -> 1.25 ms, C4 0.96 -> 0.94 ms in a build that was otherwise slower): off
#endif
#ifndef CGE_CLOSEST_LD256
#define CGE_CLOSEST_LD256 0 // closest-hit rays fetch a node with two 256-bit loads: no change measured, off
#endif
template <bool kAnyHit, bool kCount = false>
__device__ Hit trace_fast(const DevScene& s, const vec3 o, const vec3 d, float tmax, unsigned* nbox = nullptr, unsigned* ntri = nullptr)
{
    Hit h { tmax, -1, 0u };
    if (s.n_ftris == 0) {
        if (s.n_sph) {
            if (!kAnyHit)
                sphere_pass_closest(s, o, d, h);
            else if (sphere_pass_shadow(s, o, d, tmax))
                h.prim = 0, h.gid = kSphereBit;
        }
        return h;
    }
    unsigned bestRank = 0;
    // Reciprocal direction for the slab test.  For an axis with d == 0 the archive's box function substitutes the
    // constants [FLT_MIN, FLT_MAX] whatever the origin, i.e. it never rejects on that axis.
    // That quirk only ADDS box visits: a triangle can be hit only where the ray really passes, so every accepted hit
    // lies inside the geometric slabs of all its ancestors' boxes.  The fast tree therefore uses ordinary slab
    // semantics; a huge finite reciprocal (not inf) keeps 0 * inv == 0 instead of NaN when the origin lies exactly
    // on a box face.
    const SlabRay sr = slab_ray(o, d); // one FFMA per slab plane, per-axis slack (below)

    uint2 stack[kFastStackSize]; // (child ref, entry distance bits): re-culled against the best t when popped
    int sp = 0;
    constexpr unsigned kDone = 0x7fffffffu; // not a valid inner-node index
#if CGE_OCTANT_NODES && CGE_CLOSEST_OCTANT
    // the copy of the inner nodes sorted for this ray's direction signs (dev_scene.h onodes)
    unsigned cur = s.froot < kDone ? s.froot + ((sr.nx ? 1u : 0u) | (sr.ny ? 2u : 0u) | (sr.nz ? 4u : 0u)) * s.n_fnodes : s.froot;
#else
    unsigned cur = s.froot;
#endif
    auto pop = [&]() -> unsigned {
        while (sp > 0) {
            const uint2 e = stack[--sp];
            if (__uint_as_float(e.y) > h.t + fmaxf(fabsf(h.t), 1.0f) * 1e-4f)
                continue;
            return e.x;
        }
        return kDone;
    };
    // "while-while" traversal: all lanes of a warp walk inner nodes together, then all process their leaves together,
    // so the two code paths are not interleaved lane by lane.
    while (cur != kDone) {
        while (cur < kDone) { // inner node (leaf references have bit 31 set)
#if CGE_OCTANT_NODES && CGE_CLOSEST_OCTANT
            const float4* nd = s.onodes + size_t(cur) * kNodeRows;
#else
            const float4* nd = s.fnodes + size_t(cur) * kNodeRows;
#endif
#if CGE_CLOSEST_LD256
            float4 q0, q1, q2, q3;
            ldg8(nd + 2, q2, q3);
            ldg8(nd, q0, q1);
#else
            const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
#endif
            if (kCount)
                *nbox += 2;
            // a box is skipped only if it starts clearly beyond the best hit so far; the slab test carries a small
            // multiplicative slack so that rounding can only ADD visits relative to the exact arithmetic
            const float bound = h.t + fmaxf(fabsf(h.t), 1.0f) * 1e-4f;
            float entL, extL, entR, extR;
#if CGE_OCTANT_NODES && CGE_CLOSEST_OCTANT
            slab_box_sorted(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
            slab_box_sorted(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
#else
            slab_box(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
            slab_box(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
#endif
            const bool hitL = entL <= extL * 1.000002f && entL <= bound;
            const bool hitR = entR <= extR * 1.000002f && entR <= bound;
            const unsigned cl = __float_as_uint(q3.x), cr = __float_as_uint(q3.y);
            const bool leftFirst = hitL && (!hitR || entL <= entR);
            if (hitL && hitR)
                stack[sp++] = leftFirst ? make_uint2(cr, __float_as_uint(entR)) : make_uint2(cl, __float_as_uint(entL));
            if (hitL || hitR)
                cur = leftFirst ? cl : cr;
            else
                cur = pop();
        }
        if (cur == kDone)
            break;
        const unsigned first = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
        if (kCount)
            *ntri += count;
        for (unsigned i = first; i < first + count; i++) {
            float t;
            float4 r5;
            if (!triangle_rows_hit(s.ftris + size_t(i) * kTriRows, o, d, h.t, t, r5))
                continue;
            const unsigned rank = s.noaccel ? __ldg(s.fpos + i) : __float_as_uint(r5.z);
            if (t == h.t && h.prim >= 0 && rank < bestRank)
                continue; // an equal-t triangle the reference visits later is already held
            h.t = t;
            h.prim = int(i);
            h.gid = __float_as_uint(r5.w);
            bestRank = rank;
            if (kAnyHit)
                return h;
        }
        cur = pop();
    }
    if (s.n_sph) {
        if (!kAnyHit)
            sphere_pass_closest(s, o, d, h);
        else if (h.prim < 0 && sphere_pass_shadow(s, o, d, tmax))
            h.prim = 0, h.gid = kSphereBit;
    }
    return h;
}

#ifndef CGE_SHADOW_FAR_FIRST
#define CGE_SHADOW_FAR_FIRST 1 // shadow rays enter the child that starts FARTHER along the ray first, i.e. they search from the light's
                               // side: a blocked ray then meets its blocker before it has worked through the boxes crowded around
                               // its own origin (any-hit is order-free, frames are bit-identical).  Measured on B200, shadow pass near-
                               // first / far-first: C5 12.64 / 12.07 ms, one rank's 1/8 share 2.03 / 1.75 ms, C3 0.82 / 0.86 ms
#endif
#ifndef CGE_PREFETCH
#define CGE_PREFETCH 0 // 1: prefetch the far child when pushed, 2: both children on arrival - both measured slower
#endif
// pull the record a child reference points at (inner node: 64 B; leaf: its first triangle's plane row) towards L1
__device__ __forceinline__ void prefetch_child(const DevScene& s, unsigned ref)
{
    const void* p = ref < 0x7fffffffu ? static_cast<const void*>(s.fnodes + size_t(ref) * kNodeRows)
                                      : static_cast<const void*>(s.ftris + size_t(ref & 0x0fffffffu) * kTriRows);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ---- quantised nodes (dev_scene.h qnodes) ---------------------------------------------------------------------------------
// plane = qlo + m * qext  =>  t(plane) = (plane - o) * inv = fma(m, qext * inv, (qlo - o) * inv): the dequantisation costs nothing
// beyond the byte permute that builds m.  The constants carry no slack: every box was rounded outwards by >= 1 grid step
// (qext / 32768), two orders of magnitude more than the rounding of this expression (<= ~4e-7 qext |inv| for an origin inside
// the scene bounds, where every shadow ray starts).
#ifndef CGE_QNODES
#define CGE_QNODES 0 // 1: shadow rays walk the 32-byte quantised nodes.  Measured on B200: the 15-bit grid inflates the
                     // leaf-level boxes of the 868K-triangle scene enough to cost more visits than the halved traffic saves
                     // (C5 shadow pass 13.8 -> 14.5 ms; C3 0.84 -> 0.81 ms): off by default, the float nodes are walked.
#endif
struct SlabRayQ {
    vec3 a, c;                 // t = fma(m, a, c) per axis
    unsigned nearSel[3], farSel[3]; // byte-permute selectors: which half of an axis word is the near / far plane
};
__device__ __forceinline__ SlabRayQ slab_ray_q(const DevScene& s, const vec3 o, const vec3 d)
{
    SlabRayQ r;
    auto recip = [](float v) { return fabsf(v) > 1e-18f ? fdiv(1.0f, v) : 1e18f; };
    const vec3 inv = v3(recip(d.x), recip(d.y), recip(d.z));
    r.a = v3(fmul(s.qext[0], inv.x), fmul(s.qext[1], inv.y), fmul(s.qext[2], inv.z));
    r.c = v3(fmul(fsub(s.qlo[0], o.x), inv.x), fmul(fsub(s.qlo[1], o.y), inv.y), fmul(fsub(s.qlo[2], o.z), inv.z));
    // result bytes (3..0) = { 0x3F, half.hi, half.lo, 0x00 } from { K = 0x3F000000 : word }: lower half 0x7104, upper half 0x7324
    const float iv[3] = { inv.x, inv.y, inv.z };
    for (int k = 0; k < 3; k++) {
        r.nearSel[k] = iv[k] < 0.0f ? 0x7324u : 0x7104u;
        r.farSel[k] = iv[k] < 0.0f ? 0x7104u : 0x7324u;
    }
    return r;
}
__device__ __forceinline__ void slab_box_q(const SlabRayQ& r, unsigned wx, unsigned wy, unsigned wz, float& ent, float& ext)
{
    constexpr unsigned K = 0x3F000000u;
    const float tnx = __fmaf_rn(__uint_as_float(__byte_perm(wx, K, r.nearSel[0])), r.a.x, r.c.x);
    const float tfx = __fmaf_rn(__uint_as_float(__byte_perm(wx, K, r.farSel[0])), r.a.x, r.c.x);
    const float tny = __fmaf_rn(__uint_as_float(__byte_perm(wy, K, r.nearSel[1])), r.a.y, r.c.y);
    const float tfy = __fmaf_rn(__uint_as_float(__byte_perm(wy, K, r.farSel[1])), r.a.y, r.c.y);
    const float tnz = __fmaf_rn(__uint_as_float(__byte_perm(wz, K, r.nearSel[2])), r.a.z, r.c.z);
    const float tfz = __fmaf_rn(__uint_as_float(__byte_perm(wz, K, r.farSel[2])), r.a.z, r.c.z);
    ent = fmaxf(max3(tnx, tny, tnz), 0.0f);
    ext = min3(tfx, tfy, tfz);
}

#ifndef CGE_VIS_SHORT_STACK
#define CGE_VIS_SHORT_STACK 0 // > 0: the shadow-ray kernel keeps that many stack entries per lane in shared memory (SharedStack
                              // below) instead of the per-thread local array.  Measured on B200, shadow pass of
                              // C5 / a 1/8 share / C3: local array 12.07 / 1.75 / 0.86 ms; 12 shared entries 14.99 / 2.14 / 0.96 ms
                              // (10 CTAs per SM, 48 registers: 13.39 / 1.98 / 0.88; 16 entries: 13.41 / 1.93 / 0.88): the index
                              // arithmetic and the overflow test cost more than the L1 transactions they save.  Off.
#endif
// ---- traversal stacks ------------------------------------------------------------------------------------------------------
// LocalStack: a per-thread array (local memory: L1-resident, but every push / pop is an L1 transaction of its own).
// SharedStack: the first kShort entries of every lane live in shared memory, lane-interleaved (entry e of thread t at
// [e][t]: a warp's push or pop is one conflict-free access), deeper entries spill to a small local array.  The near-first
// (or far-first) walk of a binary tree holds at most one entry per level; the any-hit walks of the shadow pass rarely hold
// more than a handful, so with kShort = 12 the local part is practically never touched (ncu: local sectors ~ 0).
struct LocalStack {
    unsigned v[kFastStackSize];
    int sp = 0;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(unsigned x) { v[sp++] = x; }
    __device__ __forceinline__ bool pop(unsigned& x)
    {
        if (sp == 0)
            return false;
        x = v[--sp];
        return true;
    }
    __device__ __forceinline__ unsigned pop_or(unsigned dflt) { return sp > 0 ? v[--sp] : dflt; }
};
template <int kShort>
struct SharedStack {
    unsigned (*s)[128]; // [kShort][128] in shared memory, one column per thread of the 128-thread CTA
    unsigned tid;
    unsigned ovf[kFastStackSize - kShort];
    int sp = 0;
    __device__ __forceinline__ SharedStack(unsigned (*base)[128], unsigned t)
        : s(base)
        , tid(t)
    {
    }
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(unsigned x)
    {
        if (sp < kShort)
            s[sp][tid] = x;
        else
            ovf[sp - kShort] = x;
        sp++;
    }
    __device__ __forceinline__ bool pop(unsigned& x)
    {
        if (sp == 0)
            return false;
        sp--;
        x = sp < kShort ? s[sp][tid] : ovf[sp - kShort];
        return true;
    }
    __device__ __forceinline__ unsigned pop_or(unsigned dflt)
    {
        unsigned x = dflt;
        pop(x);
        return x;
    }
};

// Shadow rays (src/light.cpp:60-72: closest hit with ray.t = 1 used as a boolean): is ANY triangle accepted with 0 <= t <= 1?
// Lean specialisation of trace_fast<true>: the bound is the constant 1, so the stack needs no entry distances and no
// re-culling, and there is no tie bookkeeping.  Returns the blocking triangle (index into ftris), kSphereBlocker, or -1 (visible).
// kCountVisits: *visits receives the number of inner nodes the ray visited (wf_vis_regroup_kernel ranks a warp's hits by it)
// stk is used with CGE_VIS_SHORT_STACK > 0 only: otherwise the stack is a plain local array of this function (a stack OBJECT
// that outlives the call - even the trivial LocalStack - measured 5 % slower: 12.72 vs 12.12 ms on the C5 shadow pass).
template <bool kCountVisits, typename Stack>
__device__ __forceinline__ int trace_shadow_on(Stack& stk, const DevScene& s, const vec3 o, const vec3 d, unsigned* visits = nullptr)
{
    if (s.n_ftris == 0)
        return s.n_sph && sphere_pass_shadow(s, o, d, 1.0f) ? kSphereBlocker : -1;
#if CGE_QNODES
    const SlabRayQ sr = slab_ray_q(s, o, d);
#else
    const SlabRay sr = slab_ray(o, d);
#endif
#if CGE_VIS_SHORT_STACK == 0
    unsigned rawStack[kFastStackSize];
    int rawSp = 0;
#else
    stk.reset();
#endif
    constexpr unsigned kDone = 0x7fffffffu;
#if CGE_OCTANT_NODES && !CGE_QNODES
    unsigned cur = s.froot < kDone ? s.froot + ((sr.nx ? 1u : 0u) | (sr.ny ? 2u : 0u) | (sr.nz ? 4u : 0u)) * s.n_fnodes : s.froot;
#else
    unsigned cur = s.froot;
#endif
    while (cur != kDone) {
        while (cur < kDone) {
            if (kCountVisits)
                (*visits)++;
            float entL, extL, entR, extR;
#if CGE_QNODES
            const uint4* nd = s.qnodes + size_t(cur) * 2;
#if CGE_NODE_LD256
            uint4 q0, q1;
            asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w), "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w)
                : "l"(nd));
#else
            const uint4 q0 = __ldg(nd), q1 = __ldg(nd + 1);
#endif
            slab_box_q(sr, q0.x, q0.y, q0.z, entL, extL);
            slab_box_q(sr, q1.x, q1.y, q1.z, entR, extR);
            const unsigned cl = q0.w, cr = q1.w;
#else
#if CGE_OCTANT_NODES
            const float4* nd = s.onodes + size_t(cur) * kNodeRows;
#else
            const float4* nd = s.fnodes + size_t(cur) * kNodeRows;
#endif
#if CGE_NODE_LD256
            float4 q0, q1, q2, q3;
            ldg8(nd + 2, q2, q3);
            ldg8(nd, q0, q1);
#else
            const float4 q3 = ldg4(nd + 3);
            const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2);
#endif
#if CGE_PREFETCH == 2
            prefetch_child(s, __float_as_uint(q3.x));
            prefetch_child(s, __float_as_uint(q3.y));
#endif
#if CGE_OCTANT_NODES
            slab_box_sorted(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
            slab_box_sorted(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
#else
            slab_box(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
            slab_box(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
#endif
            const unsigned cl = __float_as_uint(q3.x), cr = __float_as_uint(q3.y);
#endif
            const bool hitL = entL <= extL * 1.000002f && entL <= 1.0001f;
            const bool hitR = entR <= extR * 1.000002f && entR <= 1.0001f;
#if CGE_SHADOW_FAR_FIRST
            const bool leftFirst = hitL && (!hitR || entL >= entR);
#else
            const bool leftFirst = hitL && (!hitR || entL <= entR);
#endif
            if (hitL && hitR) {
                const unsigned far = leftFirst ? cr : cl;
#if CGE_VIS_SHORT_STACK == 0
                rawStack[rawSp++] = far;
#else
                stk.push(far);
#endif
#if CGE_PREFETCH == 1
                prefetch_child(s, far);
#endif
            }
            if (hitL || hitR)
                cur = leftFirst ? cl : cr;
            else
#if CGE_VIS_SHORT_STACK == 0
                cur = rawSp > 0 ? rawStack[--rawSp] : kDone;
#else
                cur = stk.pop_or(kDone);
#endif
        }
        if (cur == kDone)
            break;
        const unsigned first = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
        for (unsigned i = first; i < first + count; i++) {
            float t;
            float4 r5;
            if (triangle_rows_hit(s.ftris + size_t(i) * kTriRows, o, d, 1.0f, t, r5))
                return int(i);
        }
#if CGE_VIS_SHORT_STACK == 0
        cur = rawSp > 0 ? rawStack[--rawSp] : kDone;
#else
        cur = stk.pop_or(kDone);
#endif
    }
    return s.n_sph && sphere_pass_shadow(s, o, d, 1.0f) ? kSphereBlocker : -1;
}

#ifndef CGE_SHADOW_BVH4
#define CGE_SHADOW_BVH4 0 // 1: shadow rays walk the 4-wide collapse of the fast tree (dev_scene.h f4nodes): half the dependent node
                          // fetches per ray for about the same box tests.  Measured on B200, C5 shadow pass: binary
                          // tree 12.8 ms (same build), 4-wide 19.6 / 16.2 / 15.1 ms at 12 / 10 / 8 CTAs per SM (40 / 48 / 64
                          // registers: 436 / 348 / 140 bytes of spills): the 28 box values of a node do not fit the register
                          // budget the latency-bound walk needs.  Off.
#endif
// The same any-hit walk over the 4-wide tree.  One visit = one 112-byte node = four box tests; the near / far plane of every slab
// is picked by LOADING the right row (the ray's sign bits choose row offsets once per ray) instead of selecting per box.  Among
// the children hit the one that starts farthest along the ray is entered first (CGE_SHADOW_FAR_FIRST), the others are pushed.
template <bool kCountVisits, typename Stack>
__device__ __forceinline__ int trace_shadow4_on(Stack& stk, const DevScene& s, const vec3 o, const vec3 d, unsigned* visits = nullptr)
{
    if (s.n_ftris == 0)
        return s.n_sph && sphere_pass_shadow(s, o, d, 1.0f) ? kSphereBlocker : -1;
    const SlabRay sr = slab_ray(o, d);
    const int nearX = sr.nx ? 3 : 0, farX = 3 - nearX, nearY = sr.ny ? 4 : 1, farY = 5 - nearY, nearZ = sr.nz ? 5 : 2, farZ = 7 - nearZ;
    stk.reset();
    constexpr unsigned kDone = 0x7fffffffu;
    unsigned cur = s.f4root;
    while (cur != kDone) {
        while (cur < kDone) {
            if (kCountVisits)
                (*visits)++;
            const float4* nd = s.f4nodes + size_t(cur) * 8;
            const float4 rf = ldg4(nd + 6);
            const float4 nx = ldg4(nd + nearX), ny = ldg4(nd + nearY), nz = ldg4(nd + nearZ);
            const float4 fx = ldg4(nd + farX), fy = ldg4(nd + farY), fz = ldg4(nd + farZ);
            const unsigned ref[4] = { __float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z), __float_as_uint(rf.w) };
            const float nxs[4] = { nx.x, nx.y, nx.z, nx.w }, nys[4] = { ny.x, ny.y, ny.z, ny.w }, nzs[4] = { nz.x, nz.y, nz.z, nz.w };
            const float fxs[4] = { fx.x, fx.y, fx.z, fx.w }, fys[4] = { fy.x, fy.y, fy.z, fy.w }, fzs[4] = { fz.x, fz.y, fz.z, fz.w };
            float ent[4];
            bool hit[4];
            int first = -1;
            float firstEnt = -1.0f;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float tn = max3(__fmaf_rn(nxs[c], sr.inv.x, sr.cNear.x), __fmaf_rn(nys[c], sr.inv.y, sr.cNear.y),
                    __fmaf_rn(nzs[c], sr.inv.z, sr.cNear.z));
                const float tf = min3(__fmaf_rn(fxs[c], sr.inv.x, sr.cFar.x), __fmaf_rn(fys[c], sr.inv.y, sr.cFar.y),
                    __fmaf_rn(fzs[c], sr.inv.z, sr.cFar.z));
                ent[c] = fmaxf(tn, 0.0f);
                hit[c] = ref[c] != kDone && ent[c] <= tf * 1.000002f && ent[c] <= 1.0001f;
#if CGE_SHADOW_FAR_FIRST
                if (hit[c] && ent[c] > firstEnt)
#else
                if (hit[c] && (first < 0 || ent[c] < firstEnt))
#endif
                    first = c, firstEnt = ent[c];
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (hit[c] && c != first)
                    stk.push(ref[c]);
            if (first >= 0)
                cur = ref[first];
            else if (!stk.pop(cur))
                cur = kDone;
        }
        if (cur == kDone)
            break;
        const unsigned firstTri = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
        for (unsigned i = firstTri; i < firstTri + count; i++) {
            float t;
            float4 r5;
            if (triangle_rows_hit(s.ftris + size_t(i) * kTriRows, o, d, 1.0f, t, r5))
                return int(i);
        }
        if (!stk.pop(cur))
            cur = kDone;
    }
    return s.n_sph && sphere_pass_shadow(s, o, d, 1.0f) ? kSphereBlocker : -1;
}

template <bool kCountVisits = false>
__device__ __forceinline__ int trace_shadow(const DevScene& s, const vec3 o, const vec3 d, unsigned* visits = nullptr)
{
    LocalStack stk;
#if CGE_SHADOW_BVH4
    return trace_shadow4_on<kCountVisits>(stk, s, o, d, visits);
#else
    return trace_shadow_on<kCountVisits>(stk, s, o, d, visits);
#endif
}

} // namespace cge
