// sampler.h — the deterministic replacement for the reference's global rand() on the soft-shadow path.
//
// The reference jitters its stratified light samples with `(float)rand() / RAND_MAX` (reference
// src/light.cpp:21,32-33).  rand() is global and order dependent, so under OpenMP the reference image is not
// reproducible.  Both sides therefore use the SAME stateless stream instead:
//
//     rand()  :=  cge_hash_sample(seed, pixel, k)        k = 0,1,2,... counts the draws made for this pixel in
//                                                         the reference's own call order
//
// pixel = y * width + x in the reference's pixel coordinates (y up, src/render.cpp:280-289).  Within a pixel
// the reference's call order is the depth-first pre-order of its 2-ary reflection recursion (direct lighting of
// a hit, then the first reflection copy's subtree, then the second copy's, src/render.cpp:33,100,118); inside one
// computeLightContribution call, lights in scene order, samples in loop order, parallelogram: horizontal draw then
// vertical draw (src/light.cpp:32-33,145-154).  On the reference side the stream is installed with
// `ld --wrap=rand` (oracle/ref/ref_api.cpp) without touching reference sources.
//
// The value is a 31-bit integer like glibc's rand(); the caller divides float(r) by float(RAND_MAX) = 2^31
// exactly as the reference expression does.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CGE_SAMPLER_HD __host__ __device__ __forceinline__
#else
#define CGE_SAMPLER_HD inline
#endif

CGE_SAMPLER_HD uint32_t cge_hash_sample(uint32_t seed, uint32_t pixel, uint32_t counter)
{
    uint32_t h = seed ^ (pixel * 0x9E3779B1u);
    h ^= counter * 0x85EBCA77u;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h >> 1;
}
