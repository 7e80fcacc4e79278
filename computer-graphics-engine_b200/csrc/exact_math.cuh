// exact_math.cuh — fp32 vector math with the reference's operation order and NO fused multiply-add.
//
// The reference engine is built without -march/-ffast-math, so every glm expression is evaluated as separate
// IEEE round-to-nearest mul / add / sub / div / sqrt; the prebuilt libIntersect archive is
// scalar SSE with the same property.  Hit/miss decisions flip if a compiler contracts a*b+c into an FMA, so on
// the device every operation goes through the __f*_rn intrinsics (which ptxas never contracts) and the
// translation unit is additionally compiled with -fmad=false; on the host the same functions are used to
// precompute ray-independent triangle quantities and are compiled with -ffp-contract=off.
//
// Operation orders reproduced (reference framework/third_party/glm/glm):
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z                 detail/func_geometric.inl:48-55
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)   :68-79
//   normalize(v)  = v * (1.0f / sqrt(dot(v,v)))                   :82-90, detail/func_exponential.inl:134-139
//   length(v)     = sqrt(dot(v,v))
//   quat * vec3   = v + ((uv * q.w) + uuv) * 2                    detail/type_quat.inl:347-354
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define CGE_HD __host__ __device__ __forceinline__

namespace cge {

CGE_HD float fmul(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
CGE_HD float fadd(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
CGE_HD float fsub(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
CGE_HD float fdiv(float a, float b)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
CGE_HD float fsqrt(float a)
{
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

struct vec3 {
    float x, y, z;
};
struct vec2 {
    float x, y;
};

CGE_HD vec3 v3(float x, float y, float z) { return vec3 { x, y, z }; }
CGE_HD vec3 v3(float s) { return vec3 { s, s, s }; }
CGE_HD vec3 operator+(vec3 a, vec3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
CGE_HD vec3 operator-(vec3 a, vec3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
CGE_HD vec3 operator*(vec3 a, vec3 b) { return v3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
CGE_HD vec3 operator*(vec3 a, float s) { return v3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
CGE_HD vec3 operator*(float s, vec3 a) { return v3(fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)); }
CGE_HD vec3 operator/(vec3 a, float s) { return v3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
CGE_HD vec3 operator-(vec3 a) { return v3(-a.x, -a.y, -a.z); }
CGE_HD vec2 operator+(vec2 a, vec2 b) { return vec2 { fadd(a.x, b.x), fadd(a.y, b.y) }; }
CGE_HD vec2 operator*(float s, vec2 a) { return vec2 { fmul(s, a.x), fmul(s, a.y) }; }

CGE_HD float dot(vec3 a, vec3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
CGE_HD vec3 cross(vec3 x, vec3 y)
{
    return v3(fsub(fmul(x.y, y.z), fmul(y.y, x.z)), fsub(fmul(x.z, y.x), fmul(y.z, x.x)), fsub(fmul(x.x, y.y), fmul(y.x, x.y)));
}
CGE_HD float length(vec3 v) { return fsqrt(dot(v, v)); }
CGE_HD vec3 normalize(vec3 v) { return v * fdiv(1.0f, fsqrt(dot(v, v))); }

// glm::quat * vec3 with q = (w, x, y, z)
CGE_HD vec3 quat_rotate(float qw, vec3 qv, vec3 v)
{
    const vec3 uv = cross(qv, v);
    const vec3 uuv = cross(qv, uv);
    return v + ((uv * qw) + uuv) * 2.0f;
}

} // namespace cge
