// render_kernels.cuh — the one-thread-per-pixel render kernel.
//
//   render_kernel       one thread = one pixel, everything sequential per thread: the literal restatement of
//                       renderRayTracing -> getFinalColor -> recursiveRayTrace (reference src/render.cpp:27-155,
//                       273-329).  Runs CGE_TRAVERSAL_REFERENCE (validation mode) and, over the fast tree, every
//                       CGE_TRAVERSAL_FAST frame without area lights (bounded cost per pixel, one launch); frames with
//                       area lights go through the wavefront pipeline (wavefront.cuh).
//                       Persistent CTAs pull 8x4 tiles from a global counter (cost per tile is wildly non-uniform).
//   trace_rays_kernel   getFinalColor for caller-supplied rays.
#pragma once
#include "shade.cuh"

namespace cge {

constexpr int kMaxLevels = kMaxRayDepth + 1;

// minimum resident 128-thread CTAs per SM the kernels are compiled for (register budget = 65536 / (128 * N));
// chosen from the A/B runs recorded in profiles/ (see DESIGN.md "Occupancy").
#ifndef CGE_MINB_THREAD
#define CGE_MINB_THREAD 8
#endif

struct Counters {
    unsigned long long primary, bounce, shadow, reference, box, tri, reference_shadow;
};

__device__ __forceinline__ void flush_counters(const Counters& c, Counters* g)
{
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&c);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(g);
    for (int k = 0; k < 7; k++) {
        unsigned long long v = src[k];
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v)
            atomicAdd(dst + k, v);
    }
}

// reference-equivalent BvhInterface::intersect calls of one pixel: level k is visited 2^k times
__device__ __forceinline__ void reference_calls(Counters& c, int n, bool missEnd, unsigned shadowPerHit)
{
    unsigned long long calls = 0;
    for (int k = 0; k < n; k++)
        calls += (1ull << k) * (1ull + shadowPerHit);
    if (missEnd)
        calls += 1ull << n;
    c.reference += calls;
    c.reference_shadow += ((1ull << n) - 1ull) * shadowPerHit;
}

// -----------------------------------------------------------------------------------------------------------------
// per-thread kernel
// -----------------------------------------------------------------------------------------------------------------
template <bool kFast, bool kSpheres, bool kCount>
struct PixelTracer {
    const DevScene& s;
    const DevParams& p;
    Counters cnt {};
    unsigned nbox = 0, ntri = 0;

    __device__ PixelTracer(const DevScene& s_, const DevParams& p_)
        : s(s_)
        , p(p_)
    {
    }

    __device__ Hit closest(const Ray& r)
    {
        if (kFast)
            return trace_fast<false, kCount>(s, r.o, r.d, r.t, &nbox, &ntri);
        return trace_reference<kSpheres, kCount>(s, r.o, r.d, r.t, nbox, ntri);
    }
    __device__ bool occluded(vec3 o, vec3 d)
    {
        cnt.shadow++;
        if (kFast)
            return kCount ? trace_fast<true, kCount>(s, o, d, 1.0f, &nbox, &ntri).prim >= 0 : trace_shadow(s, o, d) != -1;
        return trace_reference<kSpheres, kCount>(s, o, d, 1.0f, nbox, ntri).prim >= 0;
    }
    __device__ const float4* rows(const Hit& h) const { return kFast ? fast_hit_rows(s, h) : s.tris + size_t(h.prim) * kTriRows; }

    // computeLightContribution (src/light.cpp:108-164)
    __device__ vec3 direct(const HitRec& h, unsigned pixel, unsigned ctr)
    {
        if (!(p.features & CGE_FEAT_SHADING))
            return h.m.kd;
        const vec3 sp = shadow_origin(h);
        const ShadeFrame frame = shade_frame(h);
        // kFast only: a sample whose Phong term is exactly zero needs no shadow ray (shade.cuh shading_is_zero); the literal
        // traversal traces every ray the reference traces so that its ray / box / triangle counters stay comparable
        auto needed = [&](const LightSample& ls) { return !(kFast && shading_is_zero(s, frame, ls.pos)); };
        vec3 result = v3(0.0f);
        for (unsigned li = 0; li < s.n_lights; li++) {
            const float* L = s.lights + size_t(li) * kLightFloats;
            const unsigned type = __float_as_uint(__ldg(L));
            unsigned samples, draws;
            light_counts(type, p, samples, draws);
            if (type == CGE_LIGHT_POINT) {
                const LightSample ls = sample_light(L, type, 0, p, pixel, ctr);
                const vec3 c = compute_shading(ls.pos, ls.col, h);
                float vis = 1.0f;
                if (ls.shadowed && needed(ls))
                    vis = occluded(sp, ls.pos - sp) ? 0.0f : 1.0f;
                result = result + c * vis;
            } else if (samples) {
                vec3 color = v3(0.0f);
                for (unsigned si = 0; si < samples; si++) {
                    const LightSample ls = sample_light(L, type, int(si), p, pixel, ctr);
                    const float vis = needed(ls) && occluded(sp, ls.pos - sp) ? 0.0f : 1.0f;
                    color = color + compute_shading(ls.pos, ls.col, h) * vis;
                }
                // sampleSize, or sampleSizeA * sampleSizeB, as floats (src/light.cpp:137,155)
                const float denom = type == CGE_LIGHT_SEGMENT ? float(p.segment_samples)
                                                              : fmul(float(p.parallelogram_samples), float(p.parallelogram_samples));
                result = result + color / denom;
            }
            ctr += draws;
        }
        return result;
    }

    // getFinalColor (src/render.cpp:27-155).  The reference traces the mirror direction twice per hit and adds both
    // results un-attenuated (:100 and :118); both copies follow the same geometric chain, so the chain is traced
    // once and the 2-ary recursion is evaluated over it:
    //   no random draws  ->  L_k = (direct_k + L_{k+1}) + L_{k+1}, folded from the tail;
    //   soft shadows     ->  every copy draws its own jitter, so the 2^k direct terms of level k are evaluated in
    //                        the reference's depth-first order with a running draw counter.
    // ctr: index of the pixel's next rand() draw; a pixel with several camera rays (anti-aliasing) keeps counting across them
    __device__ vec3 final_color(Ray ray, unsigned pixel, int& primId, unsigned& ctr)
    {
        HitRec recs[kMaxLevels];
        vec3 directs[kMaxLevels];
        const bool fold = p.draws_per_hit == 0;
        const bool recursive = p.features & CGE_FEAT_RECURSIVE;
        int n = 0;
        bool missEnd = false;
        primId = -1;
        for (int level = 0;; level++) {
            const Hit h = closest(ray);
            if (level == 0)
                cnt.primary++;
            else
                cnt.bounce++;
            if (h.prim < 0) {
                missEnd = true;
                break;
            }
            ray.t = h.t;
            HitRec& rec = recs[fold ? 0 : level];
            resolve_hit(s, p.features, rows(h), h.gid, ray, rec);
            if (level == 0)
                primId = int(h.gid & ~kSphereBit);
            if (fold)
                directs[level] = direct(rec, pixel, 0);
            n = level + 1;
            if (!recursive || level >= p.ray_depth)
                break;
            Ray next;
            if (!reflection_ray(rec, next))
                break;
            ray = next;
        }
        reference_calls(cnt, n, missEnd, p.shadow_rays_per_hit);
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

    __device__ void finish(Counters* g)
    {
        cnt.box += nbox;
        cnt.tri += ntri;
        flush_counters(cnt, g);
    }
};

// kOut: the tile's position in this launch's list (its pixels are [32 kOut, 32 kOut + 32) of every per-launch buffer)
__device__ __forceinline__ bool next_tile(const DevParams& p, unsigned* tileCounter, unsigned lane, int& x, int& y, unsigned* kOut = nullptr)
{
    unsigned k = 0;
    if (lane == 0)
        k = atomicAdd(tileCounter, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    unsigned first = p.tile_first, count = p.tile_count, part = p.part_index;
    if (p.grant) // a chunk of some rank's tile list, dealt to this launch by the frame's shared counter (dev_scene.h)
        first = __ldg(p.grant), count = __ldg(p.grant + 1), part = __ldg(p.grant + 2);
    if (k >= count)
        return false;
    if (kOut)
        *kOut = k;
    // multi-GPU: the tile list is interleaved across ranks in units of part_unit tiles (dev_scene.h part_tile_of)
    const unsigned tile = part_tile_of(p.part_unit, part, p.part_count, first + k);
    x = int(tile % p.n_tiles_x) * kTileW + int(lane % kTileW);
    y = int(tile / p.n_tiles_x) * kTileH + int(lane / kTileW);
    return true;
}

// Where pixel (x, y) of the reference's coordinates (y up) goes in the output buffers.  Full layout: Screen::setPixel's y flip
// (src/screen.cpp:45).  Compact layout (dev_scene.h compact_units): tile row u = y / 4 is the part's j-th unit, j = (u - part_index)
// / part_count; its up to 4 image rows form block (compact_units - 1 - j) of 4 x W pixels, blocks and the rows inside a block in
// frame order (top first) - a strided copy with a positive pitch puts them into the frame.
__host__ __device__ inline size_t out_pixel_index(const DevParams& p, int x, int y)
{
    size_t row = size_t(p.height - 1 - y);
    if (p.compact_units) {
        const unsigned u = unsigned(y) / unsigned(kTileH), j = (u - p.part_index) / p.part_count;
        const int top = p.height - (int(u + 1) * kTileH < p.height ? int(u + 1) * kTileH : p.height); // frame row of the block's first row
        row = size_t(p.compact_units - 1u - j) * kTileH + (row - size_t(top));
    }
    return row * size_t(p.width) + size_t(x);
}

__device__ __forceinline__ void store_pixel(const DevParams& p, float* __restrict__ rgb, int* __restrict__ ids, int x, int y, vec3 c,
    int primId)
{
    const size_t idx = out_pixel_index(p, x, y);
    rgb[idx * 3 + 0] = c.x;
    rgb[idx * 3 + 1] = c.y;
    rgb[idx * 3 + 2] = c.z;
    if (ids)
        ids[idx] = primId;
}

template <bool kFast, bool kSpheres, bool kCount>
__global__ void __launch_bounds__(128, CGE_MINB_THREAD) render_kernel(DevScene s, DevCamera cam, DevParams p, float* __restrict__ rgb,
    int* __restrict__ ids, unsigned* __restrict__ tileCounter, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    PixelTracer<kFast, kSpheres, kCount> pt(s, p);
    int x, y;
    while (next_tile(p, tileCounter, lane, x, y)) {
        if (x < p.width && y < p.height) {
            const long long t0 = p.debug_cycles ? clock64() : 0;
            const unsigned b0 = pt.nbox;
            const Ray ray = generate_ray(cam, x, y, p.width, p.height);
            const unsigned pixel = unsigned(y) * unsigned(p.width) + unsigned(x);
            int primId;
            unsigned ctr = 0;
            vec3 c;
            if (p.aa_side) {
                // extra.enableMultipleRaysPerPixel (src/render.cpp:295-303,322): n*n jittered rays summed in order,
                // color /= n*n; colorSum += color; finalColor = colorSum / float(weight = 1)
                PixelSampler ps(p, x, y);
                vec3 color = v3(0.0f);
                for (int i = 0; i < int(p.aa_side); i++)
                    for (int j = 0; j < int(p.aa_side); j++) {
                        int subId;
                        color = color + pt.final_color(ps.ray(cam, i, j), pixel, subId, ctr);
                    }
                color = color / float(int(p.aa_side * p.aa_side));
                c = (v3(0.0f) + color) / 1.0f;
                primId = -1;
                if (ids) { // the id map stays that of the un-jittered pixel-corner ray (not a ray the reference traces: not counted)
                    const Hit h = pt.closest(ray);
                    if (h.prim >= 0)
                        primId = int(h.gid & ~kSphereBit);
                }
            } else {
                c = pt.final_color(ray, pixel, primId, ctr);
            }
            if (p.debug_cycles) // development aid: per-pixel cost map in place of the primitive ids
                primId = kCount ? int(pt.nbox - b0) : int((clock64() - t0) >> 4);
            store_pixel(p, rgb, ids, x, y, c, primId);
        }
    }
    pt.finish(gcnt);
}

// getFinalColor for caller-supplied rays (debug-ray callers, reference src/main.cpp:398,401, and unit tests)
template <bool kFast, bool kSpheres, bool kCount>
__global__ void __launch_bounds__(128) trace_rays_kernel(DevScene s, DevParams p, const float* __restrict__ rays7, unsigned n,
    float* __restrict__ rgb, int* __restrict__ ids, Counters* __restrict__ gcnt)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    PixelTracer<kFast, kSpheres, kCount> pt(s, p);
    if (i < n) {
        const float* q = rays7 + size_t(i) * 7;
        Ray ray { v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), q[6] };
        int primId;
        unsigned ctr = 0;
        const vec3 c = pt.final_color(ray, i, primId, ctr);
        rgb[i * 3 + 0] = c.x;
        rgb[i * 3 + 1] = c.y;
        rgb[i * 3 + 2] = c.z;
        if (ids)
            ids[i] = primId;
    }
    pt.finish(gcnt);
}

} // namespace cge
