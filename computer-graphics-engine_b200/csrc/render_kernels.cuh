// render_kernels.cuh — ray generation, BVH traversal, hit resolution, Phong + shadows and the reflection
// recursion of the reference's per-pixel path, as sm_100a device code.
//
// Reference functions restated here (file:line under /root/reference):
//   Trackball::generateRay                 framework/src/trackball.cpp:101-110
//   BoundingVolumeHierarchy::intersect     src/bounding_volume_hierarchy.cpp:299-427   (+ getIntersecting :272-293)
//   computeBarycentricCoord / interpolate* src/interpolate.cpp:4-28
//   acquireTexel (nearest)                 src/texture.cpp:15-27
//   computeShading / computeReflectionRay  src/shading.cpp:7-62
//   sample*Light / testVisibilityLightSample / computeLightContribution   src/light.cpp:19-164
//   recursiveRayTrace / getFinalColor      src/render.cpp:27-155
//   renderRayTracing pixel loop            src/render.cpp:273-329, Screen::setPixel src/screen.cpp:41-47
#pragma once
#include "dev_scene.h"
#include "intersect.cuh"
#include "sampler.h"
#include "cge.h"

namespace cge {

constexpr int kStackSize = 40;
constexpr int kMaxLevels = kMaxRayDepth + 1;

struct Counters {
    unsigned long long primary, bounce, shadow, reference, box, tri;
};

struct Material {
    vec3 kd, ks;
    float shininess, transparency;
};
struct HitRec {
    Ray ray; // incoming ray with t at the hit
    vec3 normal;
    Material m;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// -----------------------------------------------------------------------------------------------------------
// BVH traversal.  kFast = false reproduces the reference visit order literally (exhaustive DFS, right child
// popped first, no t culling, last accepted primitive wins).  kFast = true walks the same tree near-child-first,
// culls boxes that start beyond the current best t (with a safety margin) and resolves equal-t hits by the
// precomputed visit rank, which yields the same winner; kAnyHit additionally stops at the first accepted hit
// (shadow rays only use the boolean, src/light.cpp:61-72).
// -----------------------------------------------------------------------------------------------------------
template <bool kFast, bool kSpheres, bool kAnyHit, bool kCount>
__device__ int trace(const DevScene& s, const vec3 o, const vec3 d, float& t_io, unsigned& nbox, unsigned& ntri)
{
    if (s.n_prims == 0)
        return -1;
    float best = t_io;
    int bestPrim = -1;
    unsigned bestRank = 0;
    bool bestSphere = false;
    (void)bestRank;
    (void)bestSphere;

    uint2 stack[kStackSize];
    int sp = 0;
    stack[sp++] = make_uint2(s.root_ref, s.root_count);

    while (sp > 0) {
        const uint2 e = stack[--sp];
        if (e.y > 0) {
            // ---- leaf: getIntersecting over primitives [ref, ref+count), ascending ----
            for (unsigned i = e.x; i < e.x + e.y; i++) {
                const float4* tr = s.tris + size_t(i) * kTriRows;
                if (kSpheres) {
                    const float4 r5 = ldg4(tr + 5);
                    if (__float_as_uint(r5.w) == 1u) {
                        const float4 r1 = ldg4(tr + 1);
                        Ray ray { o, d, best };
                        if (kCount)
                            ntri++;
                        if (kFast) {
                            // strict '<' in the archive: an equal-t sphere never replaces the current hit, but a
                            // sphere visited earlier (smaller rank) would have been there first.
                            ray.t = FLT_MAX;
                            if (intersect_sphere(v3(r1.x, r1.y, r1.z), r1.w, ray, nullptr)) {
                                const unsigned rank = __float_as_uint(r5.z);
                                const bool take = ray.t < best || (ray.t == best && bestPrim >= 0 && bestSphere && rank < bestRank);
                                if (take && ray.t < t_io) {
                                    best = ray.t;
                                    bestPrim = int(i);
                                    bestRank = rank;
                                    bestSphere = true;
                                    if (kAnyHit) {
                                        t_io = best;
                                        return bestPrim;
                                    }
                                }
                            }
                        } else if (intersect_sphere(v3(r1.x, r1.y, r1.z), r1.w, ray, nullptr)) {
                            best = ray.t;
                            bestPrim = int(i);
                        }
                        continue;
                    }
                }
                if (kCount)
                    ntri++;
                // I2 with the precomputed plane (r0 = n, D)
                const float4 r0 = ldg4(tr);
                const vec3 n = v3(r0.x, r0.y, r0.z);
                const float t = fdiv(fsub(r0.w, dot(o, n)), dot(d, n));
                if (!(t >= 0.0f))
                    continue;
                if (!(best >= t))
                    continue;
                const vec3 p = d * t + o;
                // I3 with the precomputed edge vectors, short-circuit in the archive's order
                const float4 r1 = ldg4(tr + 1);
                const float4 r2 = ldg4(tr + 2);
                if (!(dot(v3(r1.w, r2.x, r2.y), p - v3(r1.x, r1.y, r1.z)) >= 0.0f))
                    continue;
                const float4 r3 = ldg4(tr + 3);
                if (!(dot(v3(r3.y, r3.z, r3.w), p - v3(r2.z, r2.w, r3.x)) >= 0.0f))
                    continue;
                const float4 r4 = ldg4(tr + 4);
                const float4 r5 = ldg4(tr + 5);
                if (!(dot(v3(r4.w, r5.x, r5.y), p - v3(r4.x, r4.y, r4.z)) >= 0.0f))
                    continue;
                if (kFast) {
                    const unsigned rank = __float_as_uint(r5.z);
                    if (t == best && bestPrim >= 0 && !bestSphere && rank < bestRank)
                        continue; // an equal-t triangle the reference would have visited later is already held
                    bestRank = rank;
                    bestSphere = false;
                }
                best = t;
                bestPrim = int(i);
                if (kFast && kAnyHit) {
                    t_io = best;
                    return bestPrim;
                }
            }
        } else {
            // ---- inner node: both child boxes tested with ray.t forced to FLT_MAX (reference :334-352) ----
            const float4* nd = s.nodes + size_t(e.x) * kNodeRows;
            const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2), q3 = ldg4(nd + 3);
            if (kCount)
                nbox += 2;
            Ray ray { o, d, FLT_MAX };
            float entL, entR;
            const bool hitL = intersect_aabb(v3(q0.x, q0.y, q0.z), v3(q0.w, q1.x, q1.y), ray, &entL);
            ray.t = FLT_MAX;
            const bool hitR = intersect_aabb(v3(q1.z, q1.w, q2.x), v3(q2.y, q2.z, q2.w), ray, &entR);
            const uint2 cl = make_uint2(__float_as_uint(q3.x), __float_as_uint(q3.z));
            const uint2 cr = make_uint2(__float_as_uint(q3.y), __float_as_uint(q3.w));
            if (!kFast) {
                if (hitL)
                    stack[sp++] = cl;
                if (hitR)
                    stack[sp++] = cr; // popped first
            } else {
                // conservative culling: a box is skipped only if it starts clearly beyond the best hit so far
                const float bound = best + fmaxf(fabsf(best), 1.0f) * 1e-4f;
                const bool goL = hitL && !(entL > bound);
                const bool goR = hitR && !(entR > bound);
                if (goL && goR) {
                    if (entL <= entR) {
                        stack[sp++] = cr;
                        stack[sp++] = cl;
                    } else {
                        stack[sp++] = cl;
                        stack[sp++] = cr;
                    }
                } else if (goL) {
                    stack[sp++] = cl;
                } else if (goR) {
                    stack[sp++] = cr;
                }
            }
        }
    }
    t_io = best;
    return bestPrim;
}

// -----------------------------------------------------------------------------------------------------------
// Hit resolution (reference src/bounding_volume_hierarchy.cpp:365-426)
// -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ vec3 barycentric(vec3 v0, vec3 v1, vec3 v2, vec3 p)
{
    // src/interpolate.cpp:4-17 (Ericson / Cramer)
    const vec3 a = v1 - v0, b = v2 - v0, c = p - v0;
    const float d00 = dot(a, a), d01 = dot(a, b), d11 = dot(b, b), d20 = dot(c, a), d21 = dot(c, b);
    const float denom = fsub(fmul(d00, d11), fmul(d01, d01));
    const float v = fdiv(fsub(fmul(d11, d20), fmul(d01, d21)), denom);
    const float w = fdiv(fsub(fmul(d00, d21), fmul(d01, d20)), denom);
    const float u = fsub(fsub(1.0f, v), w);
    return v3(u, v, w);
}

__device__ __forceinline__ vec3 acquire_texel(const DevScene& s, int tex, vec2 uv)
{
    // src/texture.cpp:19-27 (nearest): i = int(max(u*W, 0)), j = int(max((1-v)*H, 0)), clamped to W-1 / H-1
    const int4 td = __ldg(s.textures + tex);
    const float fx = fmul(uv.x, float(td.x));
    const float fy = fmul(fsub(1.0f, uv.y), float(td.y));
    int i = __float2int_rz(std_max(fx, 0.0f)); // NaN / overflow are UB in the reference; saturating here
    int j = __float2int_rz(std_max(fy, 0.0f));
    i = min(i, td.x - 1);
    j = min(j, td.y - 1);
    i = max(i, 0);
    j = max(j, 0);
    const float* px = s.texels + (size_t(unsigned(td.z)) + size_t(j) * size_t(td.x) + size_t(i)) * 3;
    return v3(__ldg(px), __ldg(px + 1), __ldg(px + 2));
}

template <bool kSpheres>
__device__ void resolve_hit(const DevScene& s, unsigned features, int prim, const Ray& ray, HitRec& rec, int& gid)
{
    const float4* tr = s.tris + size_t(prim) * kTriRows;
    const float4* sh = s.shade + size_t(prim) * kShadeRows;
    const float4 s4 = ldg4(sh + 4);
    const unsigned mid = __float_as_uint(s4.x);
    gid = int(__float_as_uint(s4.y));
    const float4 m0 = ldg4(s.materials + size_t(mid) * kMaterialRows);
    const float4 m1 = ldg4(s.materials + size_t(mid) * kMaterialRows + 1);
    const float4 m2 = ldg4(s.materials + size_t(mid) * kMaterialRows + 2);
    rec.ray = ray;
    rec.m.kd = v3(m0.x, m0.y, m0.z);
    rec.m.shininess = m0.w;
    rec.m.ks = v3(m1.x, m1.y, m1.z);
    rec.m.transparency = m1.w;
    const int tex = int(__float_as_uint(m2.x));

    bool sphere = false;
    if (kSpheres)
        sphere = __float_as_uint(ldg4(tr + 5).w) == 1u;
    if (sphere) {
        const float4 r1 = ldg4(tr + 1);
        const vec3 p = ray.o + ray.d * ray.t;
        rec.normal = normalize(p - v3(r1.x, r1.y, r1.z));
        return; // the reference returns the sphere's material untouched (:421-423)
    }
    const bool interp = features & CGE_FEAT_NORMAL_INTERP;
    const bool textured = (features & CGE_FEAT_TEXTURE_MAPPING) && tex >= 0;
    if (!interp && !textured) {
        const float4 r0 = ldg4(tr);
        rec.normal = v3(r0.x, r0.y, r0.z); // == normalize(cross(v2-v1, v3-v1)) bit for bit (:395-397 vs I1)
        return;
    }
    const float4 r1 = ldg4(tr + 1), r2 = ldg4(tr + 2), r3 = ldg4(tr + 3), r4 = ldg4(tr + 4);
    const vec3 v0 = v3(r1.x, r1.y, r1.z), v1 = v3(r2.z, r2.w, r3.x), v2 = v3(r4.x, r4.y, r4.z);
    // o + d*t and t*d + o are the same bits (IEEE add/mul commute), so one barycentric serves both uses
    const vec3 bary = barycentric(v0, v1, v2, ray.o + ray.d * ray.t);
    const float4 s0 = ldg4(sh), s1 = ldg4(sh + 1), s2 = ldg4(sh + 2), s3 = ldg4(sh + 3);
    if (interp) {
        // interpolateNormal src/interpolate.cpp:19-23, then flipped toward the viewer (:383-387)
        const vec3 n0 = v3(s0.x, s0.y, s0.z), n1 = v3(s1.x, s1.y, s1.z), n2 = v3(s2.x, s2.y, s2.z);
        vec3 nn = normalize(((n0 * bary.x + n1 * bary.y) + n2 * bary.z) / 3.0f);
        if (dot(nn, ray.d) > 0.0f)
            nn = -nn;
        rec.normal = nn;
    } else {
        const float4 r0 = ldg4(tr);
        rec.normal = v3(r0.x, r0.y, r0.z);
    }
    if (textured) {
        // interpolateTexCoord src/interpolate.cpp:25-28
        const vec2 t0 { s0.w, s1.w }, t1 { s2.w, s3.x }, t2 { s3.y, s3.z };
        const vec2 uv = (bary.x * t0 + bary.y * t1) + bary.z * t2;
        rec.m.kd = acquire_texel(s, tex, uv);
    }
}

// -----------------------------------------------------------------------------------------------------------
// Shading (src/shading.cpp) and lights (src/light.cpp)
// -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ vec3 compute_shading(vec3 lightPos, vec3 lightColor, const HitRec& h)
{
    const vec3 n = normalize(h.normal);
    const vec3 l = normalize(lightPos - (h.ray.d * h.ray.t + h.ray.o));
    const float nl = dot(n, l);
    float dd = nl;
    if (dd < 0.0f)
        dd = 0.0f;
    const vec3 diffuse = (h.m.kd * lightColor) * dd;
    const vec3 cam = normalize(h.ray.d);
    float sp = 0.0f;
    if (nl > 0.0f && dot(n, cam) > 0.0f) {
        const vec3 refl = (fmul(2.0f, dot(l, n)) * n) - l;
        sp = powf(dot(cam, refl), h.m.shininess); // negative base & non-integer exponent -> NaN, as on the CPU
    }
    const vec3 specular = (h.m.ks * lightColor) * sp;
    return diffuse + specular;
}

__device__ __forceinline__ bool reflection_ray(const HitRec& h, Ray& out)
{
    // src/shading.cpp:40-62; returns false for the ks == 0 sentinel
    if (h.m.ks.x == 0.0f && h.m.ks.y == 0.0f && h.m.ks.z == 0.0f)
        return false;
    const vec3 point = h.ray.t * h.ray.d + h.ray.o;
    const vec3 n = normalize(h.normal);
    const vec3 r = normalize(-h.ray.d);
    out.d = normalize((fmul(2.0f, dot(n, r)) * n) - r);
    out.o = point + 0.00001f * n;
    out.t = FLT_MAX;
    return true;
}

__device__ __forceinline__ float rand01(const DevParams& p, unsigned pixel, unsigned& ctr)
{
    // (float)rand() / RAND_MAX : int -> float conversion, RAND_MAX (2^31-1) converts to 2^31
    const unsigned r = cge_hash_sample(p.seed, pixel, ctr++);
    return fdiv(float(int(r)), 2147483648.0f);
}

template <bool kFast, bool kSpheres, bool kCount>
struct Shader {
    const DevScene& s;
    const DevParams& p;
    unsigned nbox = 0, ntri = 0;
    unsigned long long nshadow = 0;

    __device__ Shader(const DevScene& s_, const DevParams& p_)
        : s(s_)
        , p(p_)
    {
    }

    // testVisibilityLightSample (src/light.cpp:49-73) with the origin `sp` hoisted (it only depends on the hit)
    __device__ float visibility(vec3 sp, vec3 samplePos)
    {
        float t = 1.0f;
        nshadow++;
        const int prim = trace<kFast, kSpheres, true, kCount>(s, sp, samplePos - sp, t, nbox, ntri);
        return prim >= 0 ? 0.0f : 1.0f;
    }

    // computeLightContribution (src/light.cpp:108-164)
    __device__ vec3 direct(const HitRec& h, unsigned pixel, unsigned ctr)
    {
        if (!(p.features & CGE_FEAT_SHADING))
            return h.m.kd;
        const bool hard = p.features & CGE_FEAT_HARD_SHADOW, soft = p.features & CGE_FEAT_SOFT_SHADOW;
        // shadow-ray origin: ray.t *= length(d); d = normalize(d); p = o + d*(t - 1e-5)
        const float tl = fmul(h.ray.t, length(h.ray.d));
        const vec3 dn = normalize(h.ray.d);
        const vec3 sp = h.ray.o + dn * fsub(tl, 0.00001f);
        vec3 result = v3(0.0f);
        for (unsigned li = 0; li < s.n_lights; li++) {
            const float* L = s.lights + size_t(li) * kLightFloats;
            const unsigned type = __float_as_uint(__ldg(L));
            auto ld3 = [&](int k) { return v3(__ldg(L + 1 + k), __ldg(L + 2 + k), __ldg(L + 3 + k)); };
            if (type == CGE_LIGHT_POINT) {
                const vec3 pos = ld3(0), col = ld3(3);
                const vec3 c = compute_shading(pos, col, h);
                float vis = 1.0f;
                if (hard)
                    vis = visibility(sp, pos);
                result = result + c * vis;
            } else if (type == CGE_LIGHT_SEGMENT) {
                if (!soft)
                    continue;
                const vec3 e0 = ld3(0), e1 = ld3(3), c0 = ld3(6), c1 = ld3(9);
                vec3 color = v3(0.0f);
                const float n = float(p.segment_samples);
                for (int i = 0; float(i) < n; i++) {
                    const float r = rand01(p, pixel, ctr);
                    const float w = fdiv(fadd(float(i), r), n);
                    const vec3 pos = (e1 - e0) * w + e0;
                    const vec3 col = w * c1 + fsub(1.0f, w) * c0;
                    const float vis = visibility(sp, pos);
                    color = color + compute_shading(pos, col, h) * vis;
                }
                result = result + color / n;
            } else {
                if (!soft)
                    continue;
                const vec3 v0 = ld3(0), e01 = ld3(3), e02 = ld3(6), c0 = ld3(9), c1 = ld3(12), c2 = ld3(15), c3 = ld3(18);
                vec3 color = v3(0.0f);
                const float n = float(p.parallelogram_samples);
                for (int i = 0; float(i) < n; i++) {
                    for (int k = 0; float(k) < n; k++) {
                        const float hr = rand01(p, pixel, ctr);
                        const float vr = rand01(p, pixel, ctr);
                        const float hw = fdiv(fadd(float(i), hr), n);
                        const float vw = fdiv(fadd(float(k), vr), n);
                        const vec3 pos = (v0 + hw * e01) + vw * e02;
                        const vec3 bottom = hw * c1 + fsub(1.0f, hw) * c0;
                        const vec3 top = hw * c3 + fsub(1.0f, hw) * c2;
                        const vec3 col = vw * top + fsub(1.0f, vw) * bottom;
                        const float vis = visibility(sp, pos);
                        color = color + compute_shading(pos, col, h) * vis;
                    }
                }
                result = result + color / fmul(n, n);
            }
        }
        return result;
    }

    // getFinalColor(scene, bvh, ray, features, depth) (src/render.cpp:27-155).  The reference traces the mirror
    // direction twice per hit and adds both results un-attenuated (:100 and :118); both copies follow the same
    // geometric chain, so the chain is traced ONCE and the 2-ary recursion is evaluated over it:
    //   no random draws  ->  L_k = (direct_k + L_{k+1}) + L_{k+1}   folded from the tail,
    //   soft shadows     ->  every copy draws its own jitter, so the 2^k direct terms of level k are evaluated in
    //                        the reference's depth-first order with a running draw counter.
    __device__ vec3 final_color(Ray ray, unsigned pixel, int& primId, Counters& cnt)
    {
        HitRec recs[kMaxLevels];
        vec3 directs[kMaxLevels];
        const bool fold = p.draws_per_hit == 0;
        const bool recursive = p.features & CGE_FEAT_RECURSIVE;
        int n = 0;
        bool missEnd = false;
        primId = -1;
        for (int level = 0;; level++) {
            float t = ray.t;
            const int prim = trace<kFast, kSpheres, false, kCount>(s, ray.o, ray.d, t, nbox, ntri);
            if (level == 0)
                cnt.primary++;
            else
                cnt.bounce++;
            if (prim < 0) {
                missEnd = true;
                break;
            }
            ray.t = t;
            int gid;
            HitRec& h = recs[fold ? 0 : level];
            resolve_hit<kSpheres>(s, p.features, prim, ray, h, gid);
            if (level == 0)
                primId = gid;
            if (fold)
                directs[level] = direct(h, pixel, 0);
            n = level + 1;
            if (!recursive || level >= p.ray_depth)
                break;
            Ray next;
            if (!reflection_ray(h, next))
                break;
            ray = next;
        }
        // reference-equivalent BvhInterface::intersect calls: level k is visited 2^k times
        {
            unsigned long long calls = 0;
            for (int k = 0; k < n; k++)
                calls += (1ull << k) * (1ull + p.shadow_rays_per_hit);
            if (missEnd)
                calls += 1ull << n;
            cnt.reference += calls;
        }
        if (n == 0)
            return v3(0.0f);
        if (fold) {
            vec3 val = directs[n - 1];
            if (missEnd)
                val = (val + v3(0.0f)) + v3(0.0f);
            for (int k = n - 2; k >= 0; k--)
                val = (directs[k] + val) + val;
            return val;
        }
        vec3 acc[kMaxLevels];
        unsigned char state[kMaxLevels];
        unsigned ctr = 0;
        int level = 0;
        acc[0] = direct(recs[0], pixel, ctr);
        ctr += p.draws_per_hit;
        state[0] = 0;
        for (;;) {
            const bool spawned = (level < n - 1) || missEnd;
            if (!spawned || state[level] == 2) {
                const vec3 v = acc[level];
                if (level == 0)
                    return v;
                level--;
                acc[level] = acc[level] + v;
                state[level]++;
            } else if (level + 1 < n) {
                level++;
                acc[level] = direct(recs[level], pixel, ctr);
                ctr += p.draws_per_hit;
                state[level] = 0;
            } else {
                acc[level] = acc[level] + v3(0.0f); // the reflected copy missed (src/render.cpp:148)
                state[level]++;
            }
        }
    }
};

__device__ __forceinline__ Ray generate_ray(const DevCamera& c, int x, int y, int W, int H)
{
    // src/render.cpp:286-289 (pixel CORNER, y up) + framework/src/trackball.cpp:101-110
    const float px = fsub(fmul(fdiv(float(x), float(W)), 2.0f), 1.0f);
    const float py = fsub(fmul(fdiv(float(y), float(H)), 2.0f), 1.0f);
    const vec3 camDir = normalize(v3(fmul(-px, c.half_w), fmul(py, c.half_h), 1.0f));
    Ray r;
    r.o = v3(c.ox, c.oy, c.oz);
    r.d = quat_rotate(c.qw, v3(c.qx, c.qy, c.qz), camDir);
    r.t = FLT_MAX;
    return r;
}

__device__ __forceinline__ void flush_counters(Counters& c, unsigned nbox, unsigned ntri, unsigned long long nshadow, Counters* g)
{
    // warp-aggregate, one atomic per warp per counter
    c.shadow += nshadow;
    c.box += nbox;
    c.tri += ntri;
    unsigned long long* src = reinterpret_cast<unsigned long long*>(&c);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(g);
    for (int k = 0; k < 6; k++) {
        unsigned long long v = src[k];
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v)
            atomicAdd(dst + k, v);
    }
}

// One warp = one 8x4 pixel tile; persistent CTAs pull tiles from a global counter (dynamic assignment, because
// the cost per tile is wildly non-uniform: a miss is 1 ray, a mirror region 2^depth).  For multi-GPU runs the
// tile list is interleaved across ranks: this launch renders tiles  part_index + k * part_count.
template <bool kFast, bool kSpheres, bool kCount>
__global__ void __launch_bounds__(128) render_kernel(DevScene s, DevCamera cam, DevParams p, float* __restrict__ rgb,
    int* __restrict__ ids, unsigned* __restrict__ tileCounter, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned nTiles = p.n_tiles_x * p.n_tiles_y;
    const unsigned myTiles = (nTiles > p.part_index) ? (nTiles - p.part_index + p.part_count - 1) / p.part_count : 0;
    Shader<kFast, kSpheres, kCount> sh(s, p);
    Counters cnt {};
    for (;;) {
        unsigned k = 0;
        if (lane == 0)
            k = atomicAdd(tileCounter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= myTiles)
            break;
        const unsigned tile = p.part_index + k * p.part_count;
        const int tx = int(tile % p.n_tiles_x), ty = int(tile / p.n_tiles_x);
        const int x = tx * kTileW + int(lane % kTileW);
        const int y = ty * kTileH + int(lane / kTileW);
        if (x < p.width && y < p.height) {
            const Ray ray = generate_ray(cam, x, y, p.width, p.height);
            int primId;
            const vec3 c = sh.final_color(ray, unsigned(y) * unsigned(p.width) + unsigned(x), primId, cnt);
            const size_t idx = size_t(p.height - 1 - y) * size_t(p.width) + size_t(x); // Screen::setPixel y flip
            rgb[idx * 3 + 0] = c.x;
            rgb[idx * 3 + 1] = c.y;
            rgb[idx * 3 + 2] = c.z;
            if (ids)
                ids[idx] = primId;
        }
    }
    flush_counters(cnt, sh.nbox, sh.ntri, sh.nshadow, gcnt);
}

// getFinalColor for caller-supplied rays (debug-ray callers, reference src/main.cpp:398,401, and unit tests)
template <bool kFast, bool kSpheres, bool kCount>
__global__ void __launch_bounds__(128) trace_rays_kernel(DevScene s, DevParams p, const float* __restrict__ rays7, unsigned n,
    float* __restrict__ rgb, int* __restrict__ ids, Counters* __restrict__ gcnt)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    Shader<kFast, kSpheres, kCount> sh(s, p);
    Counters cnt {};
    if (i < n) {
        const float* q = rays7 + size_t(i) * 7;
        Ray ray { v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), q[6] };
        int primId;
        const vec3 c = sh.final_color(ray, i, primId, cnt);
        rgb[i * 3 + 0] = c.x;
        rgb[i * 3 + 1] = c.y;
        rgb[i * 3 + 2] = c.z;
        if (ids)
            ids[i] = primId;
    }
    flush_counters(cnt, sh.nbox, sh.ntri, sh.nshadow, gcnt);
}

} // namespace cge
