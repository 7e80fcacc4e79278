// shade.cuh — hit resolution, Phong shading, mirror ray, light sampling: the reference's per-hit arithmetic with the
// same fp32 operation order (no FMA).
//
//   resolve_hit        reference src/bounding_volume_hierarchy.cpp:365-426
//   barycentric        src/interpolate.cpp:4-17          interpolateNormal :19-23    interpolateTexCoord :25-28
//   acquire_texel      src/texture.cpp:15-27 (nearest)
//   compute_shading    src/shading.cpp:7-37              reflection_ray  src/shading.cpp:40-62
//   sample_light_task  src/light.cpp:19-45 (sampleSegmentLight / sampleParallelogramLight)
//   shadow_origin      src/light.cpp:54-58 (testVisibilityLightSample's offset origin)
#pragma once
#include "cge.h"
#include "sampler.h"
#include "trace.cuh"

namespace cge {

struct Material {
    vec3 kd, ks;
    float shininess;
};
struct HitRec {
    Ray ray; // incoming ray with t at the hit
    vec3 normal;
    Material m;
};
constexpr int kRecFloats = 17; // o3 d3 t n3 kd3 ks3 shininess

__device__ __forceinline__ vec3 barycentric(vec3 v0, vec3 v1, vec3 v2, vec3 p)
{
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
    // i = int(max(u*W, 0)), j = int(max((1-v)*H, 0)), clamped to W-1 / H-1
    const int4 td = __ldg(s.textures + tex);
    const float fx = fmul(uv.x, float(td.x));
    const float fy = fmul(fsub(1.0f, uv.y), float(td.y));
    int i = __float2int_rz(std_max(fx, 0.0f)); // NaN / overflow are UB in the reference; saturating here
    int j = __float2int_rz(std_max(fy, 0.0f));
    i = max(min(i, td.x - 1), 0);
    j = max(min(j, td.y - 1), 0);
    const float* px = s.texels + (size_t(unsigned(td.z)) + size_t(j) * size_t(td.x) + size_t(i)) * 3;
    return v3(__ldg(px), __ldg(px + 1), __ldg(px + 2));
}

// tr = the hit primitive's rows in whichever tree found it; gid indexes the shading records.
__device__ void resolve_hit(const DevScene& s, unsigned features, const float4* __restrict__ tr, unsigned gidBits, const Ray& ray,
    HitRec& rec)
{
    const unsigned gid = gidBits & ~kSphereBit;
    const float4* sh = s.shade + size_t(gid) * kShadeRows;
    const unsigned mid = __float_as_uint(ldg4(sh + 4).x);
    const float4 m0 = ldg4(s.materials + size_t(mid) * kMaterialRows);
    const float4 m1 = ldg4(s.materials + size_t(mid) * kMaterialRows + 1);
    const float4 m2 = ldg4(s.materials + size_t(mid) * kMaterialRows + 2);
    rec.ray = ray;
    rec.m.kd = v3(m0.x, m0.y, m0.z);
    rec.m.shininess = m0.w;
    rec.m.ks = v3(m1.x, m1.y, m1.z);
    const int tex = int(__float_as_uint(m2.x));

    if (gidBits & kSphereBit) {
        const float4 r1 = ldg4(tr + 1);
        const vec3 p = ray.o + ray.d * ray.t;
        rec.normal = normalize(p - v3(r1.x, r1.y, r1.z));
        return; // the sphere's material is returned untouched (:421-423)
    }
    const bool interp = features & CGE_FEAT_NORMAL_INTERP;
    const bool textured = (features & CGE_FEAT_TEXTURE_MAPPING) && tex >= 0;
    const float4 r0 = ldg4(tr);
    rec.normal = v3(r0.x, r0.y, r0.z); // == normalize(cross(v2-v1, v3-v1)) bit for bit (:395-397 vs I1), unflipped
    if (!interp && !textured)
        return;
    const float4 r1 = ldg4(tr + 1), r2 = ldg4(tr + 2), r3 = ldg4(tr + 3), r4 = ldg4(tr + 4);
    const vec3 v0 = v3(r1.x, r1.y, r1.z), v1 = v3(r2.z, r2.w, r3.x), v2 = v3(r4.x, r4.y, r4.z);
    // o + d*t and t*d + o are the same bits (IEEE add/mul commute), so one barycentric serves both uses
    const vec3 bary = barycentric(v0, v1, v2, ray.o + ray.d * ray.t);
    const float4 s0 = ldg4(sh), s1 = ldg4(sh + 1), s2 = ldg4(sh + 2), s3 = ldg4(sh + 3);
    if (interp) {
        const vec3 n0 = v3(s0.x, s0.y, s0.z), n1 = v3(s1.x, s1.y, s1.z), n2 = v3(s2.x, s2.y, s2.z);
        vec3 nn = normalize(((n0 * bary.x + n1 * bary.y) + n2 * bary.z) / 3.0f);
        if (dot(nn, ray.d) > 0.0f) // flipped toward the viewer (:383-387)
            nn = -nn;
        rec.normal = nn;
    }
    if (textured) {
        const vec2 t0 { s0.w, s1.w }, t1 { s2.w, s3.x }, t2 { s3.y, s3.z };
        const vec2 uv = (bary.x * t0 + bary.y * t1) + bary.z * t2;
        rec.m.kd = acquire_texel(s, tex, uv);
    }
}

// The two ray-independent operands of computeShading (src/shading.cpp:12-17): the unit normal and the hit point.
struct ShadeFrame {
    vec3 n, p;
};
__device__ __forceinline__ ShadeFrame shade_frame(const HitRec& h)
{
    return ShadeFrame { normalize(h.normal), h.ray.d * h.ray.t + h.ray.o };
}
__device__ __forceinline__ float shade_n_dot_l(const ShadeFrame& f, vec3 lightPos)
{
    return dot(f.n, normalize(lightPos - f.p));
}

// A light sample whose Phong term is exactly zero needs no visibility test.  The reference multiplies the term by the 0/1
// visibility and adds it (src/light.cpp: `color += shading * visibility`); when n.l <= 0 computeShading returns
// (kd*Lc)*0 + (ks*Lc)*0 (src/shading.cpp:19-34: the specular branch needs n.l > 0), which is +-0 for finite products, so the
// sum is the same bits whatever the shadow ray would have said.  DevScene::cull_zero_shading is set by the host only when
// every kd / ks / texel / light colour is finite and small enough that those products cannot overflow; n.l is evaluated by
// the same instructions as in compute_shading below (same inputs, no FMA), so the two always agree.  Only the FAST
// traversal uses this (the literal traversal keeps the reference's ray and test counts).
__device__ __forceinline__ bool shading_is_zero(const DevScene& s, const ShadeFrame& f, vec3 lightPos)
{
    return s.cull_zero_shading && shade_n_dot_l(f, lightPos) <= 0.0f;
}

__device__ __forceinline__ vec3 compute_shading(vec3 lightPos, vec3 lightColor, const HitRec& h)
{
    const ShadeFrame f = shade_frame(h);
    const vec3 n = f.n;
    const vec3 l = normalize(lightPos - f.p);
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
    if (h.m.ks.x == 0.0f && h.m.ks.y == 0.0f && h.m.ks.z == 0.0f)
        return false; // the reference's all-zero sentinel ray: no reflection
    const vec3 point = h.ray.t * h.ray.d + h.ray.o;
    const vec3 n = normalize(h.normal);
    const vec3 r = normalize(-h.ray.d);
    out.d = normalize((fmul(2.0f, dot(n, r)) * n) - r);
    out.o = point + 0.00001f * n;
    out.t = FLT_MAX;
    return true;
}

__device__ __forceinline__ vec3 shadow_origin(const HitRec& h)
{
    // ray.t *= length(d); d = normalize(d); p = o + d*(t - 1e-5)
    const float tl = fmul(h.ray.t, length(h.ray.d));
    const vec3 dn = normalize(h.ray.d);
    return h.ray.o + dn * fsub(tl, 0.00001f);
}

__device__ __forceinline__ float rand01(unsigned seed, unsigned pixel, unsigned ctr)
{
    // (float)rand() / RAND_MAX : int -> float conversion, RAND_MAX (2^31-1) converts to 2^31.  Dividing by a power of two is
    // exact, so the product with 2^-31 is the same float as the IEEE quotient (one FMUL instead of a ~10-instruction division;
    // the quotient is an integer < 2^31 scaled down, never subnormal).
    return fmul(float(int(cge_hash_sample(seed, pixel, ctr))), 4.656612873077392578125e-10f);
}

// x / float(n) as the reference computes it.  For a power-of-two sample count (the common 2 / 4 / 8 / 16) the quotient equals the
// product with the exactly representable reciprocal, which the host put into DevParams (0 = not a power of two): one FMUL
// instead of an IEEE division; the branch is uniform.
__device__ __forceinline__ float div_by_count(float x, int n, float exactRecip)
{
    return exactRecip != 0.0f ? fmul(x, exactRecip) : fdiv(x, float(n));
}

struct LightSample {
    vec3 pos, col;
    bool shadowed; // whether the reference tests visibility for this sample
};

// Sample `si` of light `L` (24-float record).  ctr = draw index of this light's first rand() call.
__device__ __forceinline__ LightSample sample_light(const float* __restrict__ L, unsigned type, int si, const DevParams& p,
    unsigned pixel, unsigned ctr)
{
    auto ld3 = [&](int k) { return v3(__ldg(L + 1 + k), __ldg(L + 2 + k), __ldg(L + 3 + k)); };
    LightSample out;
    if (type == CGE_LIGHT_POINT) {
        out.pos = ld3(0);
        out.col = ld3(3);
        out.shadowed = p.features & CGE_FEAT_HARD_SHADOW;
    } else if (type == CGE_LIGHT_SEGMENT) {
        const vec3 e0 = ld3(0), e1 = ld3(3), c0 = ld3(6), c1 = ld3(9);
        const float r = rand01(p.seed, pixel, ctr + unsigned(si));
        const float w = div_by_count(fadd(float(si), r), p.segment_samples, p.segment_recip);
        out.pos = (e1 - e0) * w + e0;
        out.col = w * c1 + fsub(1.0f, w) * c0;
        out.shadowed = true;
    } else {
        const vec3 v0 = ld3(0), e01 = ld3(3), e02 = ld3(6), c0 = ld3(9), c1 = ld3(12), c2 = ld3(15), c3 = ld3(18);
        const int ns = p.parallelogram_samples;
        const int i = si / ns, k = si % ns; // i (edge01) outer, k (edge02) inner; horizontal draw first
        const float hr = rand01(p.seed, pixel, ctr + 2u * unsigned(si));
        const float vr = rand01(p.seed, pixel, ctr + 2u * unsigned(si) + 1u);
        const float hw = div_by_count(fadd(float(i), hr), ns, p.parallelogram_recip);
        const float vw = div_by_count(fadd(float(k), vr), ns, p.parallelogram_recip);
        out.pos = (v0 + hw * e01) + vw * e02;
        const vec3 bottom = hw * c1 + fsub(1.0f, hw) * c0;
        const vec3 top = hw * c3 + fsub(1.0f, hw) * c2;
        out.col = vw * top + fsub(1.0f, vw) * bottom;
        out.shadowed = true;
    }
    return out;
}

// samples / rand() draws a light contributes to one computeLightContribution call (src/light.cpp:114-156)
__device__ __forceinline__ void light_counts(unsigned type, const DevParams& p, unsigned& samples, unsigned& draws)
{
    const bool soft = p.features & CGE_FEAT_SOFT_SHADOW;
    if (type == CGE_LIGHT_POINT) {
        samples = 1;
        draws = 0;
    } else if (type == CGE_LIGHT_SEGMENT) {
        samples = soft ? unsigned(p.segment_samples) : 0u;
        draws = samples;
    } else {
        samples = soft ? unsigned(p.parallelogram_samples * p.parallelogram_samples) : 0u;
        draws = 2u * samples;
    }
}

// Trackball::generateRay(pixel) (framework/src/trackball.cpp:101-110) for a position in normalised device coordinates
__device__ __forceinline__ Ray generate_ray_ndc(const DevCamera& c, float px, float py)
{
    const vec3 camDir = normalize(v3(fmul(-px, c.half_w), fmul(py, c.half_h), 1.0f));
    Ray r;
    r.o = v3(c.ox, c.oy, c.oz);
    r.d = quat_rotate(c.qw, v3(c.qx, c.qy, c.qz), camDir);
    r.t = FLT_MAX;
    return r;
}

__device__ __forceinline__ Ray generate_ray(const DevCamera& c, int x, int y, int W, int H)
{
    // src/render.cpp:286-289 (pixel CORNER, y up)
    const float px = fsub(fmul(fdiv(float(x), float(W)), 2.0f), 1.0f);
    const float py = fsub(fmul(fdiv(float(y), float(H)), 2.0f), 1.0f);
    return generate_ray_ndc(c, px, py);
}

// getRaySamples (src/render.cpp:211-227): the n x n jittered camera rays of one pixel, in the reference's order (i over x
// outer, j over y inner).  g++ evaluates the arguments of `glm::vec2(xJitter(gen), yJitter(gen))` right to left, so the
// FIRST draw of a sample is its y jitter (pinned against the compiled reference, tests/test_oracle_pins.py).
struct PixelSampler {
    CgeMt19937Head mt;
    float px, py, bx, by; // pixel corner and the size of one stratum (pixelBox) in NDC
    __device__ PixelSampler(const DevParams& p, int x, int y)
    {
        px = py = bx = by = 0.0f;
        if (!p.aa_side) // feature off: nothing to set up (the generator's warm-up is 397 integer steps)
            return;
        mt = CgeMt19937Head(cge_aa_seed(p.seed, unsigned(y) * unsigned(p.width) + unsigned(x)));
        const float n = float(p.aa_side);
        px = fsub(fmul(fdiv(float(x), float(p.width)), 2.0f), 1.0f);
        py = fsub(fmul(fdiv(float(y), float(p.height)), 2.0f), 1.0f);
        bx = fdiv(fmul(fdiv(1.0f, float(p.width)), 2.0f), n);
        by = fdiv(fmul(fdiv(1.0f, float(p.height)), 2.0f), n);
    }
    // std::uniform_real_distribution<float>(0, b)(gen) as libstdc++ evaluates it: generate_canonical * (b - a) + a
    __device__ float uniform(float b)
    {
        float u = fdiv(__uint2float_rn(mt.next()), 4294967296.0f);
        if (u >= 1.0f)
            u = 0.99999994f; // nextafter(1, 0)
        return fadd(fmul(u, fsub(b, 0.0f)), 0.0f);
    }
    __device__ Ray ray(const DevCamera& c, int i, int j)
    {
        const float jy = uniform(by);
        const float jx = uniform(bx);
        return generate_ray_ndc(c, fadd(fadd(px, fmul(float(i), bx)), jx), fadd(fadd(py, fmul(float(j), by)), jy));
    }
};

} // namespace cge
