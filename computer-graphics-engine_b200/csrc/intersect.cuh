// intersect.cuh — the six functions of the reference's prebuilt libIntersect (declared in reference
// src/intersect.h:5-16, object code prebuilt/libIntersect_linux_x64.a; behaviour decoded in SURVEY.md Appendix A),
// written for sm_100a with the same scalar fp32 operation order, IEEE div/sqrt and no FMA.
// NOT Möller–Trumbore: plane test + three inclusive edge tests.  Comparisons are arranged so that NaN takes the
// reject branch exactly like the archive's comiss/jb sequences.
#pragma once
#include <float.h>

#include "exact_math.cuh"

namespace cge {

struct Ray {
    vec3 o, d;
    float t;
};
struct Plane {
    float D;
    vec3 n;
};

// I1  trianglePlane (src/intersect.h:10): n = normalize(cross(v1-v0, v2-v0)); D = dot(n, v0)
CGE_HD Plane triangle_plane(vec3 v0, vec3 v1, vec3 v2)
{
    Plane p;
    p.n = normalize(cross(v1 - v0, v2 - v0));
    p.D = dot(p.n, v0);
    return p;
}

// I2  intersectRayWithPlane (src/intersect.h:5): t = (D - dot(o,n)) / dot(d,n); accept 0 <= t <= ray.t
CGE_HD bool intersect_plane(const Plane& pl, Ray& ray)
{
    const float t = fdiv(fsub(pl.D, dot(ray.o, pl.n)), dot(ray.d, pl.n));
    if (!(t >= 0.0f))
        return false;
    if (!(ray.t >= t))
        return false;
    ray.t = t;
    return true;
}

// I3  pointInTriangle (src/intersect.h:8): three inclusive edge tests, short-circuit left to right
CGE_HD bool point_in_triangle(vec3 v0, vec3 v1, vec3 v2, vec3 n, vec3 p)
{
    if (!(dot(cross(v2 - v0, n), p - v0) >= 0.0f))
        return false;
    if (!(dot(cross(v0 - v1, n), p - v1) >= 0.0f))
        return false;
    return dot(cross(v1 - v2, n), p - v2) >= 0.0f;
}

// I4  intersectRayWithTriangle (src/intersect.h:12).  hitInfo is untouched on success and restored on failure,
// so only ray.t is observable.
CGE_HD bool intersect_triangle(vec3 v0, vec3 v1, vec3 v2, Ray& ray)
{
    const float tOld = ray.t;
    const Plane pl = triangle_plane(v0, v1, v2);
    if (intersect_plane(pl, ray)) {
        const vec3 p = ray.d * ray.t + ray.o;
        if (point_in_triangle(v0, v1, v2, pl.n, p))
            return true;
    }
    ray.t = tOld;
    return false;
}

// Ray-independent part of I4, precomputed once per triangle with identical bits (last
// paragraph): n, D and the three edge vectors.
struct TriPre {
    vec3 n;
    float D;
    vec3 e0, e1, e2;
};
CGE_HD TriPre triangle_precompute(vec3 v0, vec3 v1, vec3 v2)
{
    TriPre t;
    const Plane pl = triangle_plane(v0, v1, v2);
    t.n = pl.n;
    t.D = pl.D;
    t.e0 = cross(v2 - v0, pl.n);
    t.e1 = cross(v0 - v1, pl.n);
    t.e2 = cross(v1 - v2, pl.n);
    return t;
}

// std::min / std::max argument semantics: min(a,b) = (b < a) ? b : a ; max(a,b) = (a < b) ? b : a
CGE_HD float std_min(float a, float b) { return (b < a) ? b : a; }
CGE_HD float std_max(float a, float b) { return (a < b) ? b : a; }

// I5  intersectRayWithShape(AxisAlignedBox, Ray) (src/intersect.h:16).
// `entry` (optional) receives max(raw tin, 0): the parametric distance at which the ray enters the box, used
// only by the ordered traversal for near-first ordering and conservative culling — never for the boolean.
CGE_HD bool intersect_aabb(vec3 lower, vec3 upper, Ray& ray, float* entry = nullptr)
{
    float txU, txL, tyU, tyL, tzU, tzL;
    if (ray.d.x != 0.0f) {
        txU = fdiv(fsub(upper.x, ray.o.x), ray.d.x);
        txL = fdiv(fsub(lower.x, ray.o.x), ray.d.x);
    } else {
        txU = FLT_MAX;
        txL = FLT_MIN; // sic: smallest positive normal, not -FLT_MAX
    }
    if (ray.d.y != 0.0f) {
        tyU = fdiv(fsub(upper.y, ray.o.y), ray.d.y);
        tyL = fdiv(fsub(lower.y, ray.o.y), ray.d.y);
    } else {
        tyU = FLT_MAX;
        tyL = FLT_MIN;
    }
    if (ray.d.z != 0.0f) {
        tzU = fdiv(fsub(upper.z, ray.o.z), ray.d.z);
        tzL = fdiv(fsub(lower.z, ray.o.z), ray.d.z);
    } else {
        tzU = FLT_MAX;
        tzL = FLT_MIN;
    }
    float tin = std_max(std_max(std_min(txL, txU), std_min(tyL, tyU)), std_min(tzL, tzU));
    float tout = std_min(std_min(std_max(txL, txU), std_max(tyL, tyU)), std_max(tzL, tzU));
    if (entry)
        *entry = tin < 0.0f ? 0.0f : tin;
    if (tin < 0.0f) {
        if (!(tout > 0.0f))
            return false;
        tin = tout;
        tout = FLT_MAX;
    }
    if (tin > tout || tin < 0.0f || tin > ray.t)
        return false;
    ray.t = tin;
    return true;
}

// I6  intersectRayWithShape(Sphere, Ray, HitInfo) (src/intersect.h:14): assumes |d| == 1, strict t < ray.t.
CGE_HD bool intersect_sphere(vec3 c, float r, Ray& ray, vec3* normal)
{
    const vec3 oc = ray.o - c;
    float B = dot(ray.d, oc);
    B = fadd(B, B);
    const float C = fsub(dot(oc, oc), fmul(r, r));
    const float disc = fsub(fmul(B, B), fmul(C, 4.0f));
    if (disc < 0.0f)
        return false;
    float t0, t1;
    if (disc == 0.0f) {
        t0 = t1 = fmul(0.5f, -B);
    } else {
        const float q = fsqrt(disc);
        t1 = fmul(fsub(q, B), 0.5f);
        t0 = fmul(fsub(-B, q), 0.5f);
    }
    if (t0 < 0.0f)
        t0 = FLT_MAX;
    if (t1 < 0.0f)
        t1 = FLT_MAX;
    const float t = std_min(t0, t1);
    if (t < 0.0f || t >= ray.t)
        return false;
    if (normal)
        *normal = normalize((ray.o + ray.d * t) - c);
    ray.t = t;
    return true;
}

} // namespace cge
