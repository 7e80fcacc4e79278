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

// ---- multi-sample anti-aliasing (reference getRaySamples, src/render.cpp:211-227) -------------------------------------
// The reference seeds a fresh std::mt19937 per pixel from std::random_device and draws its jitter through
// std::uniform_real_distribution<float>.  random_device is the only non-deterministic piece; both sides replace its value
// by cge_aa_seed(seed, pixel) (reference side: ld --wrap of std::random_device::_M_getval, oracle/ref/ref_api.cpp) and then
// run the SAME generator: MT19937 (init_genrand, twist, tempering) and libstdc++'s generate_canonical<float, 24>
// (one 32-bit draw: float(u32) / 2^32, clamped below 1).
CGE_SAMPLER_HD uint32_t cge_aa_seed(uint32_t seed, uint32_t pixel) { return cge_hash_sample(seed ^ 0x52444556u, pixel, 0u); }

// The first 227 outputs of std::mt19937(seed) without the 624-word state: output i is the tempered
//   s[i + 397] ^ twist(s[i], s[i + 1])      with s[] the init_genrand sequence s[j] = 1812433253 * (s[j-1] ^ (s[j-1] >> 30)) + j,
// because for i < 227 the twist only reads words the first regeneration has not yet overwritten.  Two running copies of
// the init recurrence (at index i and at index i + 397) are all the state needed: 2 * 10 * 10 = 200 draws suffice for the
// largest raysPerPixelSide the reference GUI allows.
struct CgeMt19937Head {
    uint32_t lo, hi, i; // s[i], s[i + 397]
    CGE_SAMPLER_HD CgeMt19937Head()
        : lo(0)
        , hi(0)
        , i(0)
    {
    }
    CGE_SAMPLER_HD explicit CgeMt19937Head(uint32_t seed)
        : lo(seed)
        , hi(seed)
        , i(0)
    {
        for (uint32_t j = 1; j <= 397u; j++)
            hi = 1812433253u * (hi ^ (hi >> 30)) + j;
    }
    CGE_SAMPLER_HD uint32_t next()
    {
        const uint32_t lo1 = 1812433253u * (lo ^ (lo >> 30)) + (i + 1u);
        const uint32_t y = (lo & 0x80000000u) | (lo1 & 0x7fffffffu);
        uint32_t v = hi ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        lo = lo1;
        hi = 1812433253u * (hi ^ (hi >> 30)) + (i + 398u);
        i++;
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        v ^= v >> 18;
        return v;
    }
};
constexpr int kCgeMaxRaysPerPixelSide = 10; // 2 * n * n draws must stay below 227
