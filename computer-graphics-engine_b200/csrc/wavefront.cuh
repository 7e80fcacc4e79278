// wavefront.cuh — the production pipeline of CGE_TRAVERSAL_FAST: kernels over compact queues in HBM.
//
// Why: in the one-thread-per-pixel kernel a pixel is 1 ray (a miss) or up to (levels + 16 * (2^levels - 1)) rays (a
// mirror chain with soft shadows); measured on config C5 the slowest 8x4 tile costs 350x the median tile and the last
// tiles alone keep the kernel alive for a third of its run time (profiles/, DESIGN.md "Kernels").  The wavefront
// splits the frame into uniform work items and compacts them:
//
//   wf_chain_kernel  one lane = one pixel: generate the camera ray, trace the mirror chain (closest hits only, <= levels
//                    rays).  Every hit is appended to the queue of its recursion level: the lanes of a warp that hit are
//                    counted with __ballot_sync, ONE atomicAdd per warp reserves a contiguous run of queue slots, and
//                    each lane writes its hit record (SoA, coalesced) at base + __popc(ballot & lanes_below).
//                    Consecutive queue slots are therefore neighbouring pixels.
//   wf_vis_cull_kernel / wf_vis_regroup_kernel   the shadow stage: every shadow ray of the frame as one visibility byte.  The
//                    light-hull pre-pass settles, with ONE conservative walk per hit, the hits none of whose rays can be blocked
//                    and compacts the others into per-level hard lists; the per-ray kernel traces those (both further down).
//   wf_shade_kernel  one lane = one direct-lighting evaluation (pixel, level, reflection copy) = one
//                    computeLightContribution call of the reference: all its shadow rays, in the reference's sample
//                    order.  Work items are numbered copy-major inside a level, so a warp holds 32 neighbouring pixels
//                    evaluating the SAME sample sequence: coherent rays, equal trip counts, full warps.
//   wf_fold_kernel   one lane = one pixel with a primary hit: folds the 2-ary reflection recursion from the stored
//                    direct terms (reference src/render.cpp:100,118: Lo = (direct + R1) + R2) and writes the pixel.
//
// Queue layout (cap = pixels covered by this launch):
//   rec   [level][kWaveRecFloats][cap] float   hit record + shadow-ray origin, structure of arrays
//   vis   [ray index] u8                   ray index = level offset + ((copy * S + sample) * count[level]) + slot
//   meta  [level][cap] uint2                .x = pixel index (y*W + x, reference coordinates), .y = n | missEnd << 8
//   next  [level][cap] uint                 queue slot of the same pixel at level+1 (valid while level+1 < n)
//   dir   block of level k at dirOff(k): [copy][3][cap] float
#pragma once
#include "render_kernels.cuh"

namespace cge {

constexpr int kWaveRecFloats = kRecFloats + 3; // hit record + the shadow-ray origin of src/light.cpp:54-58 (hoisted)

// meta word layout.  .x = pixel index (y * W + x, reference coordinates) | sub-ray << 24: with extra.enableMultipleRaysPerPixel
// a pixel has n x n camera rays, each a chain of its own in the queues (the host takes this path only for frames of at most
// 2^24 pixels).  .y = hit levels of the chain | missEnd << 8 | units << 9, units = direct-lighting evaluations made by the
// pixel's EARLIER camera rays: the reference's rand() counter keeps running across them (src/render.cpp:297-299).
constexpr unsigned kMetaPixelMask = 0x00ffffffu;
__device__ __forceinline__ unsigned wf_pixel(const uint2& m) { return m.x & kMetaPixelMask; }
__device__ __forceinline__ unsigned wf_sub_ray(const uint2& m) { return m.x >> 24; }

struct WaveBuffers {
    float* rec;
    uint2* meta;
    unsigned* next;
    float* dir;
    unsigned char* vis; // one byte per (level, copy, sample, slot): 1 = light sample visible
    unsigned* counts;   // [0..15] queue length per level, [16] chain tile counter, [17] shade chunk counter, [18] / [19] shadow-ray chunk
                        // counters (level 0 / deeper levels when the chain stage is split), [20] continuation-queue length, [21] its
                        // chunk counter
    unsigned* hard;     // [level][cap] queue slots of the hits the light-hull pre-pass could not resolve (wf_vis_cull_kernel); their
                        // number per level is counts[32 + level]; counts[22] / [23] are that kernel's chunk counters
    unsigned* cont;     // [cap] level-0 queue slots of the hits whose chain goes on (wf_primary_kernel -> wf_continue_kernel)
    float* sub;         // multiple rays per pixel: [3][launch pixel][sub-ray] colour of every camera ray, summed by wf_resolve_kernel
    unsigned cap;
};

// Shadow rays of this launch (see cge_api.cu launch_render): true = traced by wf_vis_regroup_kernel into one visibility byte per
// ray, wf_shade_kernel<true> shades from the bytes; false = wf_shade_kernel<false> traces them itself (frames with fewer than 8
// samples per evaluation, e.g. point lights forced through the wavefront, or visibility bytes beyond 4 GB).
__device__ __forceinline__ bool wf_use_visibility_bytes(const DevParams& p, const WaveBuffers&) { return p.shade_mode != 1; }

__device__ __forceinline__ size_t wf_dir_off(const DevParams& p, unsigned cap, unsigned k)
{
    const unsigned unitsBefore = p.draws_per_hit == 0 ? k : ((1u << k) - 1u);
    return size_t(unitsBefore) * 3u * cap;
}

#ifndef CGE_MINB_CHAIN
#define CGE_MINB_CHAIN 6
#endif
// kSubRays: extra.enableMultipleRaysPerPixel (a separate instantiation so that the common case does not carry the sampler's state)
template <bool kSubRays>
__global__ void __launch_bounds__(128, kSubRays ? CGE_MINB_CHAIN : 8) wf_chain_kernel(DevScene s, DevCamera cam, DevParams p, WaveBuffers wb, float* __restrict__ rgb,
    int* __restrict__ ids, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const bool recursive = p.features & CGE_FEAT_RECURSIVE;
    Counters cnt {};
    int x, y;
    unsigned tileK = 0;
    const unsigned aa = kSubRays ? p.aa_side : 0u;
    const unsigned nSub = kSubRays ? aa * aa : 1u;
    while (next_tile(p, wb.counts + 16, lane, x, y, &tileK)) {
        const bool live = x < p.width && y < p.height;
        const unsigned pixel = unsigned(y) * unsigned(p.width) + unsigned(x);
        DevParams sp = p;
        if (!kSubRays)
            sp.aa_side = 0;
        PixelSampler ps(sp, live ? x : 0, live ? y : 0);
        unsigned unitsBefore = 0; // direct-lighting evaluations of this pixel's earlier camera rays
        if (kSubRays && live && ids) { // the id map stays that of the un-jittered pixel-corner ray (not a reference ray: not counted)
            const Ray c = generate_ray(cam, x, y, p.width, p.height);
            const Hit h = trace_fast<false>(s, c.o, c.d, c.t);
            ids[out_pixel_index(p, x, y)] = h.prim >= 0 ? int(h.gid & ~kSphereBit) : -1;
        }
        for (unsigned sub = 0; sub < nSub; sub++) {
            Ray ray {};
            if (live)
                ray = kSubRays ? ps.ray(cam, int(sub / aa), int(sub % aa)) : generate_ray(cam, x, y, p.width, p.height);
            bool alive = live;
            bool missEnd = false;
            int n = 0;
            unsigned slots[kMaxLevels];
            for (int level = 0; __any_sync(0xffffffffu, alive); level++) {
                bool hit = false;
                Hit h {};
                if (alive) {
                    h = trace_fast<false>(s, ray.o, ray.d, ray.t);
                    if (level == 0)
                        cnt.primary++;
                    else
                        cnt.bounce++;
                    hit = h.prim >= 0;
                    if (!hit) {
                        missEnd = true;
                        alive = false;
                    }
                }
                // warp-aggregated queue append: one atomic per warp and level
                const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                unsigned base = 0;
                if (lane == 0 && ballot)
                    base = atomicAdd(wb.counts + level, unsigned(__popc(ballot)));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) {
                    const unsigned slot = base + unsigned(__popc(ballot & below));
                    slots[level] = slot;
                    ray.t = h.t;
                    HitRec r;
                    resolve_hit(s, p.features, fast_hit_rows(s, h), h.gid, ray, r);
                    float* b = wb.rec + (size_t(level) * kWaveRecFloats) * wb.cap + slot;
                    const size_t c = wb.cap;
                    b[0 * c] = r.ray.o.x, b[1 * c] = r.ray.o.y, b[2 * c] = r.ray.o.z;
                    b[3 * c] = r.ray.d.x, b[4 * c] = r.ray.d.y, b[5 * c] = r.ray.d.z;
                    b[6 * c] = r.ray.t;
                    b[7 * c] = r.normal.x, b[8 * c] = r.normal.y, b[9 * c] = r.normal.z;
                    b[10 * c] = r.m.kd.x, b[11 * c] = r.m.kd.y, b[12 * c] = r.m.kd.z;
                    b[13 * c] = r.m.ks.x, b[14 * c] = r.m.ks.y, b[15 * c] = r.m.ks.z;
                    b[16 * c] = r.m.shininess;
                    const vec3 so = shadow_origin(r);
                    b[17 * c] = so.x, b[18 * c] = so.y, b[19 * c] = so.z;
                    if (level > 0)
                        wb.next[size_t(level - 1) * wb.cap + slots[level - 1]] = slot;
                    else if (ids && !kSubRays)
                        ids[out_pixel_index(p, x, y)] = int(h.gid & ~kSphereBit);
                    n = level + 1;
                    Ray nextRay;
                    if (!recursive || level >= p.ray_depth || !reflection_ray(r, nextRay))
                        alive = false;
                    else
                        ray = nextRay;
                }
            }
            if (live) {
                reference_calls(cnt, n, missEnd, p.shadow_rays_per_hit);
                const uint2 m = make_uint2(pixel | (sub << 24), unsigned(n) | (missEnd ? 256u : 0u) | (unitsBefore << 9));
                for (int k = 0; k < n; k++)
                    wb.meta[size_t(k) * wb.cap + slots[k]] = m;
                if (kSubRays)
                    unitsBefore += p.draws_per_hit ? (1u << n) - 1u : 0u;
                if (n == 0) {
                    if (!kSubRays) {
                        store_pixel(p, rgb, ids, x, y, v3(0.0f), -1); // primary miss: black (reference src/render.cpp:148)
                    } else { // this camera ray adds vec3(0) to the pixel's sum
                        const size_t at = (size_t(tileK) * 32u + lane) * nSub + sub, plane = size_t(p.tile_count) * 32u * nSub;
                        wb.sub[at] = 0.0f, wb.sub[plane + at] = 0.0f, wb.sub[2 * plane + at] = 0.0f;
                    }
                }
            }
        }
    }
    flush_counters(cnt, gcnt);
}

// ---------------------------------------------------------------------------------------------------------------------
// The chain stage in two kernels, so that its tail can hide under the shadow pass (frames without multiple rays per pixel and
// with recursion on; cge_api.cu launch_render).  The tail of wf_chain_kernel is a handful of pixels whose mirror chains graze the
// model - up to 4 dependent closest-hit rays of several hundred node visits each, ~0.4 ms that no partition of the frame shortens
// (on a 1/8 share the schedulers idle 60 % of the kernel, profiles/r02_wf_chain_part8_c5.txt).  The camera rays have no such
// tail, and level 0 holds most of the frame's direct-lighting evaluations:
//   wf_primary_kernel   camera rays only; hits go to the level-0 queue, hits that reflect additionally to the continuation
//                       queue.  When it has finished the level-0 queue is final and its shadow rays start (own stream).
//   wf_continue_kernel  one lane = the rest of one pixel's chain (levels 1 ..), 32 reflecting pixels per warp.  Runs beside the
//                       level-0 shadow pass; the deeper levels' shadow rays follow it.
// Both write exactly the records, links and tags wf_chain_kernel writes (the slot ORDER differs, which nothing depends on).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wf_write_record(const WaveBuffers& wb, unsigned level, unsigned slot, const HitRec& r)
{
    float* b = wb.rec + (size_t(level) * kWaveRecFloats) * wb.cap + slot;
    const size_t c = wb.cap;
    b[0 * c] = r.ray.o.x, b[1 * c] = r.ray.o.y, b[2 * c] = r.ray.o.z;
    b[3 * c] = r.ray.d.x, b[4 * c] = r.ray.d.y, b[5 * c] = r.ray.d.z;
    b[6 * c] = r.ray.t;
    b[7 * c] = r.normal.x, b[8 * c] = r.normal.y, b[9 * c] = r.normal.z;
    b[10 * c] = r.m.kd.x, b[11 * c] = r.m.kd.y, b[12 * c] = r.m.kd.z;
    b[13 * c] = r.m.ks.x, b[14 * c] = r.m.ks.y, b[15 * c] = r.m.ks.z;
    b[16 * c] = r.m.shininess;
    const vec3 so = shadow_origin(r);
    b[17 * c] = so.x, b[18 * c] = so.y, b[19 * c] = so.z;
}

__global__ void __launch_bounds__(128, 8) wf_primary_kernel(DevScene s, DevCamera cam, DevParams p, WaveBuffers wb, float* __restrict__ rgb,
    int* __restrict__ ids, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const bool recursive = (p.features & CGE_FEAT_RECURSIVE) && p.ray_depth > 0;
    Counters cnt {};
    int x, y;
    while (next_tile(p, wb.counts + 16, lane, x, y)) {
        const bool live = x < p.width && y < p.height;
        const unsigned pixel = unsigned(y) * unsigned(p.width) + unsigned(x);
        Ray ray {};
        Hit h {};
        bool hit = false;
        if (live) {
            ray = generate_ray(cam, x, y, p.width, p.height);
            h = trace_fast<false>(s, ray.o, ray.d, ray.t);
            cnt.primary++;
            hit = h.prim >= 0;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        unsigned base = 0;
        if (lane == 0 && ballot)
            base = atomicAdd(wb.counts + 0, unsigned(__popc(ballot)));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned slot = base + unsigned(__popc(ballot & below));
        bool goesOn = false;
        if (hit) {
            ray.t = h.t;
            HitRec r;
            resolve_hit(s, p.features, fast_hit_rows(s, h), h.gid, ray, r);
            wf_write_record(wb, 0, slot, r);
            if (ids)
                ids[out_pixel_index(p, x, y)] = int(h.gid & ~kSphereBit);
            goesOn = recursive && !(r.m.ks.x == 0.0f && r.m.ks.y == 0.0f && r.m.ks.z == 0.0f); // computeReflectionRay's sentinel
            // the tag (.y) of a chain that goes on is written by wf_continue_kernel when the chain has ended
            wb.meta[slot] = make_uint2(pixel, goesOn ? 0u : 1u);
        }
        const unsigned more = __ballot_sync(0xffffffffu, goesOn);
        unsigned cbase = 0;
        if (lane == 0 && more)
            cbase = atomicAdd(wb.counts + 20, unsigned(__popc(more)));
        cbase = __shfl_sync(0xffffffffu, cbase, 0);
        if (goesOn)
            wb.cont[cbase + unsigned(__popc(more & below))] = slot;
        else if (live) {
            reference_calls(cnt, hit ? 1 : 0, !hit, p.shadow_rays_per_hit);
            if (!hit)
                store_pixel(p, rgb, ids, x, y, v3(0.0f), -1); // primary miss: black (reference src/render.cpp:148)
        }
    }
    flush_counters(cnt, gcnt);
}

__global__ void __launch_bounds__(128, 8) wf_continue_kernel(DevScene s, DevParams p, WaveBuffers wb, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const unsigned total = wb.counts[20];
    Counters cnt {};
    for (;;) {
        unsigned chunk = 0;
        if (lane == 0)
            chunk = atomicAdd(wb.counts + 21, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if ((unsigned long long)chunk * 32ull >= total)
            break;
        const unsigned i = chunk * 32u + lane;
        const bool live = i < total;
        unsigned slots[kMaxLevels];
        unsigned pixel = 0;
        Ray ray {};
        if (live) {
            const unsigned e0 = wb.cont[i];
            slots[0] = e0;
            pixel = wb.meta[e0].x;
            // the reflected ray of the level-0 hit, from its record (computeReflectionRay, src/shading.cpp:40-62)
            const float* b = wb.rec + e0;
            const size_t c = wb.cap;
            HitRec prev;
            prev.ray.o = v3(b[0 * c], b[1 * c], b[2 * c]);
            prev.ray.d = v3(b[3 * c], b[4 * c], b[5 * c]);
            prev.ray.t = b[6 * c];
            prev.normal = v3(b[7 * c], b[8 * c], b[9 * c]);
            prev.m.ks = v3(b[13 * c], b[14 * c], b[15 * c]);
            reflection_ray(prev, ray); // ks != 0 was checked when the entry was queued
        }
        bool alive = live, missEnd = false;
        int n = 1;
        for (int level = 1; __any_sync(0xffffffffu, alive); level++) {
            bool hit = false;
            Hit h {};
            if (alive) {
                h = trace_fast<false>(s, ray.o, ray.d, ray.t);
                cnt.bounce++;
                hit = h.prim >= 0;
                if (!hit) {
                    missEnd = true;
                    alive = false;
                }
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            unsigned base = 0;
            if (lane == 0 && ballot)
                base = atomicAdd(wb.counts + level, unsigned(__popc(ballot)));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned slot = base + unsigned(__popc(ballot & below));
                slots[level] = slot;
                ray.t = h.t;
                HitRec r;
                resolve_hit(s, p.features, fast_hit_rows(s, h), h.gid, ray, r);
                wf_write_record(wb, unsigned(level), slot, r);
                wb.next[size_t(level - 1) * wb.cap + slots[level - 1]] = slot;
                n = level + 1;
                Ray nextRay;
                if (level >= p.ray_depth || !reflection_ray(r, nextRay))
                    alive = false;
                else
                    ray = nextRay;
            }
        }
        if (live) {
            reference_calls(cnt, n, missEnd, p.shadow_rays_per_hit);
            const uint2 m = make_uint2(pixel, unsigned(n) | (missEnd ? 256u : 0u));
            wb.meta[slots[0]].y = m.y; // (32-bit store: the level-0 shadow pass may be reading .x of this entry)
            for (int k = 1; k < n; k++)
                wb.meta[size_t(k) * wb.cap + slots[k]] = m;
        }
    }
    flush_counters(cnt, gcnt);
}

// computeLightContribution for one (pixel, level, copy): same arithmetic and order as PixelTracer::direct.
// kLookup: visibilities were traced by wf_vis_regroup_kernel; vis points at this evaluation's sample 0, stride = queue length.
template <bool kLookup>
__device__ __forceinline__ vec3 wf_direct(const DevScene& s, const DevParams& p, const HitRec& h, unsigned pixel, unsigned ctr,
    unsigned long long& nshadow, const unsigned char* __restrict__ vis, size_t visStride)
{
    const vec3 sp = kLookup ? v3(0.0f) : shadow_origin(h);
    const ShadeFrame frame = shade_frame(h);
    vec3 result = v3(0.0f);
    unsigned sg = 0; // running sample index over all lights
    for (unsigned li = 0; li < s.n_lights; li++) {
        const float* L = s.lights + size_t(li) * kLightFloats;
        const unsigned type = __float_as_uint(__ldg(L));
        unsigned samples, draws;
        light_counts(type, p, samples, draws);
        if (type == CGE_LIGHT_POINT) {
            const LightSample ls = sample_light(L, type, 0, p, pixel, ctr);
            const vec3 c = compute_shading(ls.pos, ls.col, h);
            float v = 1.0f;
            if (kLookup) {
                v = vis[size_t(sg) * visStride] ? 1.0f : 0.0f;
            } else if (ls.shadowed && !shading_is_zero(s, frame, ls.pos)) {
                nshadow++;
                v = trace_shadow(s, sp, ls.pos - sp) != -1 ? 0.0f : 1.0f;
            }
            result = result + c * v;
        } else if (samples) {
            vec3 color = v3(0.0f);
            for (unsigned si = 0; si < samples; si++) {
                const LightSample ls = sample_light(L, type, int(si), p, pixel, ctr);
                float v;
                if (kLookup) {
                    v = vis[size_t(sg + si) * visStride] ? 1.0f : 0.0f;
                } else if (shading_is_zero(s, frame, ls.pos)) {
                    v = 1.0f;
                } else {
                    nshadow++;
                    v = trace_shadow(s, sp, ls.pos - sp) != -1 ? 0.0f : 1.0f;
                }
                color = color + compute_shading(ls.pos, ls.col, h) * v;
            }
            const float denom = type == CGE_LIGHT_SEGMENT ? float(p.segment_samples)
                                                          : fmul(float(p.parallelogram_samples), float(p.parallelogram_samples));
            result = result + color / denom;
        }
        ctr += draws;
        sg += samples;
    }
    return result;
}

// first rand() draw index of the direct-lighting evaluation (level k, copy `path`) of a chain with n hit levels, in the
// reference's depth-first order over its 2-ary recursion:  index = sum_{i=1..k} (1 + b_i * (2^(n-i) - 1)),
// b_i = i-th copy choice on the way down (src/render.cpp:33,100,118)
__device__ __forceinline__ unsigned wf_draw_base(const DevParams& p, unsigned k, unsigned path, unsigned nChain)
{
    if (p.draws_per_hit == 0)
        return 0;
    unsigned idx = 0;
    for (unsigned i = 1; i <= k; i++) {
        const unsigned b = (path >> (k - i)) & 1u;
        idx += 1u + b * ((1u << (nChain - i)) - 1u);
    }
    return idx * p.draws_per_hit;
}
// ... plus the draws of the pixel's earlier camera rays (meta .y bits 9..31, zero without multiple rays per pixel)
__device__ __forceinline__ unsigned wf_unit_ctr(const DevParams& p, unsigned k, unsigned path, const uint2& m)
{
    return wf_draw_base(p, k, path, m.y & 255u) + (m.y >> 9) * p.draws_per_hit;
}

// hit-record fields the zero-shading cull needs (shade.cuh shading_is_zero): incoming ray and normal of queue slot e, level k
__device__ __forceinline__ ShadeFrame wf_load_frame(const WaveBuffers& wb, unsigned k, unsigned e)
{
    const float* b = wb.rec + (size_t(k) * kWaveRecFloats) * wb.cap + e;
    const size_t c = wb.cap;
    HitRec h;
    h.ray.o = v3(b[0 * c], b[1 * c], b[2 * c]);
    h.ray.d = v3(b[3 * c], b[4 * c], b[5 * c]);
    h.ray.t = b[6 * c];
    h.normal = v3(b[7 * c], b[8 * c], b[9 * c]);
    return shade_frame(h);
}

// sample sg of one computeLightContribution call -> (light record, sample within the light, first draw of that light)
__device__ __forceinline__ LightSample wf_sample(const DevScene& s, const DevParams& p, unsigned sg, unsigned pixel, unsigned ctrBase)
{
    unsigned li = 0, si = sg, samples = 0, draws = 0, type = 0, ctr = ctrBase;
    const float* L = s.lights;
    for (;; li++) {
        L = s.lights + size_t(li) * kLightFloats;
        type = __float_as_uint(__ldg(L));
        light_counts(type, p, samples, draws);
        if (si < samples)
            break;
        si -= samples;
        ctr += draws;
    }
    return sample_light(L, type, int(si), p, pixel, ctr);
}

// ---------------------------------------------------------------------------------------------------------------------
// wf_vis_cull_kernel: the light-hull pre-pass of the shadow stage.  One lane = one hit (level, queue slot), whatever its copies
// and samples: EVERY shadow ray the hit can need leaves the same point o (src/light.cpp:54-58) and ends inside the convex hull
// of its lights' corner points (src/light.cpp:19-45: position = (v0 + hw * e01) + vw * e02 with hw, vw in [0, 1]; every float
// operation is monotone, so each coordinate of every sample lies between the values at the corners).  The lane walks the fast tree
// ONCE with that hull - the per-axis direction box [dmin, dmax] of (corner - o), evaluated per node exactly like a ray's slab test -
// near child first, and looks at every triangle of every leaf the hull reaches:
//   * a triangle none of the hull's rays can be accepted by is CLEAR (below: sign / magnitude of the plane parameter, then the
//     four side planes of the pyramid (o, light corners) pushed outwards by a margin that covers the rounding of the archive's
//     test);
//   * if every triangle met is clear and the stack runs empty, no ray of the hit can be blocked: all its visibility bytes are 1
//     and it is finished - one walk of ~25 nodes instead of copies x samples walks;
//   * the first triangle that is not clear (or an exhausted budget) sends the hit to the HARD list of its level, which is what
//     wf_vis_regroup_kernel then traces ray by ray, exactly as before.
// The pre-pass only ever answers "certainly visible"; every decision it does not make is made by the unchanged per-ray code, so
// frames are bit-identical with and without it (CGE_VIS_CULL=0).  Scenes with spheres skip it (their test is not a plane test).
// ---------------------------------------------------------------------------------------------------------------------
// Per-axis constants of the hull test.  Constraint A: o + t * dmin <= hi, constraint B: o + t * dmax >= lo.
//   all directions positive : B is the lower bound (entry), A the upper bound (exit)
//   all directions negative : A is the lower bound, B the upper bound
//   mixed signs             : A and B are both lower bounds, the axis has no upper bound
// v1 = fma(sel ? hi : lo, k1, c1) is always a lower bound; v2 = fma(sel ? lo : hi, k2, c2) is an upper bound, or with `mixed`
// a second lower bound.  The constants carry the cancellation slack of trace.cuh slab_ray, doubled: c -+ 2^-21 |c| moves each
// plane ~8 ulps of the origin's coordinate outwards.
struct HullAxis {
    float k1, c1, k2, c2;
    bool sel, mixed;
};
CGE_HD HullAxis hull_axis(float o, float dmin, float dmax)
{
    auto recip = [](float v, float tiny) { return fabsf(v) > 1e-18f ? fdiv(1.0f, v) : tiny; };
    HullAxis h;
    const bool pos = dmin > 0.0f, neg = dmax < 0.0f;
    const float iA = recip(dmin, pos ? 1e18f : -1e18f); // dmin == 0: no ray moves towards -axis: hi < o rejects
    const float iB = recip(dmax, neg ? -1e18f : 1e18f);
    const float cA = -fmul(o, iA), cB = -fmul(o, iB);
    const float sA = fabsf(cA) * 4.76837158203125e-07f, sB = fabsf(cB) * 4.76837158203125e-07f;
    h.sel = !pos; // v1 reads hi (constraint A) unless all directions are positive
    h.mixed = !pos && !neg;
    if (pos) {
        h.k1 = iB, h.c1 = cB - sB; // lower
        h.k2 = iA, h.c2 = cA + sA; // upper
    } else {
        h.k1 = iA, h.c1 = cA - sA;                           // lower
        h.k2 = iB, h.c2 = h.mixed ? cB - sB : cB + sB; // second lower bound, or the upper bound
    }
    return h;
}

// The hull against one box: entry / exit parameter of the set { o + t * d : d in the hull's direction box } (entry clamped to 0); some
// ray of the hull may pass through the box within t in [0, 1] iff hull_box_hit(ent, ext).  Host + device: the CPU suite checks the
// test's conservativeness through cge_hull_box_host (tests/test_hull_clear.py).
struct HullWalk {
    HullAxis hx, hy, hz;
    bool anyMixed;
};
CGE_HD HullWalk hull_walk(const vec3 o, const vec3 dmin, const vec3 dmax)
{
    HullWalk w;
    w.hx = hull_axis(o.x, dmin.x, dmax.x), w.hy = hull_axis(o.y, dmin.y, dmax.y), w.hz = hull_axis(o.z, dmin.z, dmax.z);
    w.anyMixed = w.hx.mixed || w.hy.mixed || w.hz.mixed;
    return w;
}
CGE_HD float hull_fma(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
CGE_HD float hull_max3(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return max3(a, b, c);
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
CGE_HD float hull_min3(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return min3(a, b, c);
#else
    return fminf(fminf(a, b), c);
#endif
}
CGE_HD void hull_box(const HullWalk& w, float lox, float loy, float loz, float hix, float hiy, float hiz, float& ent, float& ext)
{
    const float kInf = __builtin_huge_valf();
    const HullAxis &hx = w.hx, &hy = w.hy, &hz = w.hz;
    const float ax = hull_fma(hx.sel ? hix : lox, hx.k1, hx.c1), bx = hull_fma(hx.sel ? lox : hix, hx.k2, hx.c2);
    const float ay = hull_fma(hy.sel ? hiy : loy, hy.k1, hy.c1), by = hull_fma(hy.sel ? loy : hiy, hy.k2, hy.c2);
    const float az = hull_fma(hz.sel ? hiz : loz, hz.k1, hz.c1), bz = hull_fma(hz.sel ? loz : hiz, hz.k2, hz.c2);
    ent = fmaxf(hull_max3(ax, ay, az), 0.0f);
    if (!w.anyMixed) {
        ext = hull_min3(bx, by, bz);
    } else {
        ent = fmaxf(ent, hull_max3(hx.mixed ? bx : 0.0f, hy.mixed ? by : 0.0f, hz.mixed ? bz : 0.0f));
        ext = hull_min3(hx.mixed ? kInf : bx, hy.mixed ? kInf : by, hz.mixed ? kInf : bz);
    }
}
CGE_HD bool hull_box_hit(float ent, float ext) { return ent <= ext * 1.000008f && ent <= 1.0001f; }

// The points between which every shadowed sample of light L lies (n = 1, 2 or 4; 0: the light casts no shadow ray)
__device__ __forceinline__ unsigned light_corners(const float* __restrict__ L, const DevParams& p, vec3 c[4])
{
    auto ld3 = [&](int k) { return v3(__ldg(L + 1 + k), __ldg(L + 2 + k), __ldg(L + 3 + k)); };
    const unsigned type = __float_as_uint(__ldg(L));
    unsigned samples, draws;
    light_counts(type, p, samples, draws);
    if (type == CGE_LIGHT_POINT) {
        c[0] = ld3(0);
        return (p.features & CGE_FEAT_HARD_SHADOW) ? 1u : 0u;
    }
    if (samples == 0)
        return 0;
    if (type == CGE_LIGHT_SEGMENT) {
        const vec3 e0 = ld3(0), e1 = ld3(3);
        c[0] = e0;                     // w = 0: (e1 - e0) * 0 + e0
        c[1] = (e1 - e0) * 1.0f + e0;  // w = 1, evaluated like sample_light
        return 2;
    }
    const vec3 v0 = ld3(0), e01 = ld3(3), e02 = ld3(6);
    c[0] = v0;
    c[1] = v0 + e01;
    c[2] = (v0 + e01) + e02;
    c[3] = v0 + e02;
    return 4;
}

// a row of a triangle record: read-only cache path on the device, a plain load in the host-side test entry (cge_hull_clear_host)
CGE_HD float4 hull_row(const float4* p)
{
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
// per-axis bounds of the hull's ray directions (corner - o over all corners of all lights) and what the clear test derives from them
struct HullDirs {
    vec3 dmin, dmax;
};
CGE_HD HullDirs hull_dirs_empty()
{
    const float inf = __builtin_huge_valf();
    return HullDirs { v3(inf), v3(-inf) };
}
CGE_HD void hull_dirs_add(HullDirs& h, const vec3 d)
{
    h.dmin = v3(fminf(h.dmin.x, d.x), fminf(h.dmin.y, d.y), fminf(h.dmin.z, d.z));
    h.dmax = v3(fmaxf(h.dmax.x, d.x), fmaxf(h.dmax.y, d.y), fmaxf(h.dmax.z, d.z));
}
// an upper bound of |d| over the hull
CGE_HD float hull_dirs_length(const HullDirs& h)
{
    const float mx = fmaxf(fabsf(h.dmin.x), fabsf(h.dmax.x)), my = fmaxf(fabsf(h.dmin.y), fabsf(h.dmax.y)), mz = fmaxf(fabsf(h.dmin.z), fabsf(h.dmax.z));
    return fsqrt(mx * mx + my * my + mz * mz) * 1.0001f;
}
CGE_HD float max_abs3(const vec3 v) { return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fabsf(v.z)); }

// Can NO ray o + t * d, t in [0, 1], d inside the hull, be accepted by triangle tr?  (true = certainly not)
//   a[4] = light corners - o (pyramid edges; used when `pyramid`), dmin / dmax = their per-axis bounds, dLen >= |d| of every ray,
//   mag = coordinate magnitude of the configuration (sets the absolute rounding scale)
CGE_HD bool cull_triangle_clear(const float4* __restrict__ tr, const vec3 o, const vec3 dmin, const vec3 dmax, float dLen,
    bool pyramid, const vec3* a, float mag)
{
    const float4 r0 = hull_row(tr);
    const vec3 n = v3(r0.x, r0.y, r0.z);
    const float num = fsub(r0.w, dot(o, n)); // I2's numerator, the archive's own operations: its sign is the archive's sign
    // range of the denominators dot(d, n) over the hull, widened by the rounding of a three-term product sum
    const float ex0 = n.x * dmin.x, ex1 = n.x * dmax.x, ey0 = n.y * dmin.y, ey1 = n.y * dmax.y, ez0 = n.z * dmin.z, ez1 = n.z * dmax.z;
    const float denEps = (fmaxf(fabsf(ex0), fabsf(ex1)) + fmaxf(fabsf(ey0), fabsf(ey1)) + fmaxf(fabsf(ez0), fabsf(ez1))) * 1e-6f;
    const float denLo = fminf(ex0, ex1) + fminf(ey0, ey1) + fminf(ez0, ez1) - denEps;
    const float denHi = fmaxf(ex0, ex1) + fmaxf(ey0, ey1) + fmaxf(ez0, ez1) + denEps;
    // 0 <= num / den <= 1 needs a den with num's sign and |den| >= |num| (a NaN plane fails both comparisons: never accepted)
    const bool viaPos = num >= 0.0f && denHi * 1.000001f >= num, viaNeg = num <= 0.0f && denLo * 1.000001f <= num;
    if (!viaPos && !viaNeg)
        return true;
    if (!pyramid)
        return false;
    // The rays that can reach the plane meet it under an angle whose cosine is at least cosMin; the archive's hit point then lies
    // within ~(few ulp of the coordinates) / cosMin of the exact one.  A hull that contains rays parallel to the plane cannot be
    // judged by distances at all.
    const float denMin = viaPos && viaNeg ? 0.0f : viaPos ? fmaxf(denLo, 0.0f) : fmaxf(-denHi, 0.0f);
    const float nLen = fsqrt(dot(n, n));
    const float cosMin = denMin / (dLen * nLen);
    if (!(cosMin > 1e-3f))
        return false;
    const float margin = mag * 4e-6f / fminf(cosMin, 1.0f);
    const float4 r1 = hull_row(tr + 1), r2 = hull_row(tr + 2), r3 = hull_row(tr + 3), r4 = hull_row(tr + 4);
    const vec3 w0 = v3(r1.x, r1.y, r1.z) - o, w1 = v3(r2.z, r2.w, r3.x) - o, w2 = v3(r4.x, r4.y, r4.z) - o;
    const vec3 m = (a[0] + a[1]) + (a[2] + a[3]); // a direction inside the pyramid
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) {
        vec3 pn = cross(a[i], a[(i + 1) & 3]);
        const float inside = dot(pn, m), len = fsqrt(dot(pn, pn));
        if (!(fabsf(inside) > 1e-3f * len * fsqrt(dot(m, m))))
            continue; // degenerate side (the origin lies in the light's plane): no judgement from this plane
        const float sgn = inside > 0.0f ? -1.0f : 1.0f, thr = margin * len;
        if (sgn * dot(pn, w0) > thr && sgn * dot(pn, w1) > thr && sgn * dot(pn, w2) > thr)
            return true; // the whole triangle lies outside this side of the pyramid, by more than the margin
    }
    return false;
}

#ifndef CGE_MINB_CULL
#define CGE_MINB_CULL 8
#endif
__global__ void __launch_bounds__(128, CGE_MINB_CULL) wf_vis_cull_kernel(DevScene s, DevParams p, WaveBuffers wb,
    unsigned levelBegin, unsigned levelEnd, unsigned counterIdx)
{
    const unsigned lane = threadIdx.x & 31, below = (1u << lane) - 1u;
    const bool fold = p.draws_per_hit == 0;
    const unsigned S = p.samples_per_hit;
    unsigned cum[kMaxLevels + 1], hitsBefore[kMaxLevels + 1]; // direct-lighting evaluations before level k; hits of this launch before it
    cum[0] = 0;
    for (unsigned k = 0; k < p.levels; k++)
        cum[k + 1] = cum[k] + wb.counts[k] * (fold ? 1u : (1u << k));
    hitsBefore[levelBegin] = 0;
    for (unsigned k = levelBegin; k < levelEnd; k++)
        hitsBefore[k + 1] = hitsBefore[k] + wb.counts[k];
    const unsigned total = hitsBefore[levelEnd];
    // one parallelogram light: the pyramid (o, corners) is known and its side planes can clear triangles; otherwise only the
    // direction box of all lights' corners is used
    const bool pyramid = s.n_lights == 1 && __float_as_uint(__ldg(s.lights)) == CGE_LIGHT_PARALLELOGRAM;
    float lightMag = 0.0f;
    for (unsigned li = 0; li < s.n_lights; li++) {
        vec3 c[4];
        const unsigned nc = light_corners(s.lights + size_t(li) * kLightFloats, p, c);
        for (unsigned j = 0; j < nc; j++)
            lightMag = fmaxf(lightMag, max_abs3(c[j]));
    }
    constexpr unsigned kDone = 0x7fffffffu;
    unsigned long long nResolved = 0;
    for (;;) {
        unsigned chunk = 0;
        if (lane == 0)
            chunk = atomicAdd(wb.counts + counterIdx, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        const unsigned first = chunk * 32u;
        if (first >= total)
            break;
        const unsigned item = first + lane;
        bool hardHit = false;
        unsigned k = levelBegin, e = 0;
        if (item < total) {
            while (item >= hitsBefore[k + 1])
                k++;
            e = item - hitsBefore[k];
            const float* b = wb.rec + (size_t(k) * kWaveRecFloats + 17) * wb.cap + e;
            const vec3 o = v3(b[0], b[wb.cap], b[2 * size_t(wb.cap)]);
            // the hull: per-axis bounds of (corner - o) over all lights; for the single parallelogram also the four edges
            vec3 a[4];
            HullDirs dirs = hull_dirs_empty();
            unsigned corners = 0;
            for (unsigned li = 0; li < s.n_lights; li++) {
                vec3 c[4];
                const unsigned nc = light_corners(s.lights + size_t(li) * kLightFloats, p, c);
                for (unsigned j = 0; j < nc; j++) {
                    const vec3 d = c[j] - o;
                    if (pyramid)
                        a[j] = d;
                    hull_dirs_add(dirs, d);
                }
                corners += nc;
            }
            const vec3 dmin = dirs.dmin, dmax = dirs.dmax;
            const bool finite = fabsf(dmin.x) <= 3e38f && fabsf(dmin.y) <= 3e38f && fabsf(dmin.z) <= 3e38f && fabsf(dmax.x) <= 3e38f
                && fabsf(dmax.y) <= 3e38f && fabsf(dmax.z) <= 3e38f;
            if (corners == 0 || s.n_ftris == 0) {
                hardHit = false; // nothing casts or nothing blocks a shadow ray: every byte is 1
            } else if (!finite) {
                hardHit = true;
            } else {
                const HullWalk walk = hull_walk(o, dmin, dmax);
                const float dLen = hull_dirs_length(dirs);
                const float mag = max_abs3(o) + lightMag;
                unsigned stack[kFastStackSize];
                int sp = 0;
                unsigned cur = s.froot;
                unsigned budget = p.cull_budget; // inner nodes this walk may visit before it gives the hit up
                while (cur != kDone) {
                    while (cur < kDone) {
                        if (budget-- == 0) {
                            hardHit = true;
                            break;
                        }
                        const float4* nd = s.fnodes + size_t(cur) * kNodeRows;
                        float4 q0, q1, q2, q3;
#if CGE_NODE_LD256
                        ldg8(nd + 2, q2, q3);
                        ldg8(nd, q0, q1);
#else
                        q3 = ldg4(nd + 3), q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2);
#endif
                        float entL, extL, entR, extR;
                        hull_box(walk, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
                        hull_box(walk, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
                        const unsigned cl = __float_as_uint(q3.x), cr = __float_as_uint(q3.y);
                        const bool hitL = hull_box_hit(entL, extL), hitR = hull_box_hit(entR, extR);
                        const bool leftFirst = hitL && (!hitR || entL <= entR);
                        if (hitL && hitR)
                            stack[sp++] = leftFirst ? cr : cl;
                        if (hitL || hitR)
                            cur = leftFirst ? cl : cr;
                        else
                            cur = sp > 0 ? stack[--sp] : kDone;
                    }
                    if (hardHit || cur == kDone)
                        break;
                    const unsigned firstTri = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
                    for (unsigned i = firstTri; i < firstTri + count; i++)
                        if (!cull_triangle_clear(s.ftris + size_t(i) * kTriRows, o, dmin, dmax, dLen, pyramid, a, mag)) {
                            hardHit = true;
                            break;
                        }
                    if (hardHit)
                        break;
                    cur = sp > 0 ? stack[--sp] : kDone;
                }
            }
            if (!hardHit) { // every visibility byte of the hit (all copies, all samples) is 1
                const unsigned cnt = wb.counts[k], copies = fold ? 1u : (1u << k);
                unsigned char* v = wb.vis + size_t(cum[k]) * S + e;
                for (unsigned j = 0; j < copies * S; j++)
                    v[size_t(j) * cnt] = 1;
                nResolved += copies * S;
            }
        }
        // the hard hits of this warp keep their order: one atomic per warp and level reserves their run of the list
        for (unsigned kk = levelBegin; kk < levelEnd; kk++) {
            const unsigned ballot = __ballot_sync(0xffffffffu, hardHit && k == kk);
            if (!ballot)
                continue;
            unsigned base = 0;
            if (lane == 0)
                base = atomicAdd(wb.counts + 32 + kk, unsigned(__popc(ballot)));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hardHit && k == kk)
                wb.hard[size_t(kk) * wb.cap + base + unsigned(__popc(ballot & below))] = e;
        }
    }
    // light samples settled without a ray: a 64-bit counter in counts[48..49] (cge_stats::shadow_samples_culled)
    for (int off = 16; off > 0; off >>= 1)
        nResolved += __shfl_down_sync(0xffffffffu, nResolved, off);
    if (lane == 0 && nResolved)
        atomicAdd(reinterpret_cast<unsigned long long*>(wb.counts + 48), nResolved);
}

constexpr int kVisShortStack = CGE_VIS_SHORT_STACK > 0 ? CGE_VIS_SHORT_STACK : 1;
#ifndef CGE_MINB_VIS
#define CGE_MINB_VIS 12 // resident 128-thread CTAs per SM the shadow-ray kernel is compiled for (A/B in DESIGN.md 5.5)
#endif
// ---------------------------------------------------------------------------------------------------------------------
// wf_vis_regroup_kernel: every shadow ray of the frame, any-hit, one visibility byte per ray.  A work item is 32 neighbouring
// hits (one per lane) x kGroup consecutive light samples, numbered level -> copy -> sample group -> queue slot, so a warp traces
// the same samples for 32 neighbouring pixels.  The lanes of a warp trade hits between the samples: a lane's kGroup rays are all
// long or all short (they start at the same point and end on the same small light), so with one hit per lane for all kGroup
// steps (the form this kernel replaced) every step of a warp lasts as long as the warp's longest hit.  Here step 0 is the same (every lane traces the first sample of its own hit, counting the nodes
// visited); then the 32 hits are ranked by that count and the remaining 32 x (kGroup - 1) rays are dealt out rank by rank:
// the second step gets all remaining samples of the longest third of the hits, the last step those of the shortest third.
// The hits are still the same 32 neighbouring pixels (the coherence of the node fetches is untouched), a hit's data travel
// between lanes by shuffle, every ray is traced exactly once with the same arithmetic into the same visibility byte.
// ---------------------------------------------------------------------------------------------------------------------
struct RegroupHit {
    unsigned k, e, path, g, pixel, ctrBase, cnt;
    vec3 o;
    ShadeFrame f;
    int occluder;
    unsigned valid;
};
__device__ __forceinline__ RegroupHit regroup_fetch(const RegroupHit& h, unsigned src)
{
    const unsigned full = 0xffffffffu;
    RegroupHit r;
    r.k = __shfl_sync(full, h.k, src), r.e = __shfl_sync(full, h.e, src), r.path = __shfl_sync(full, h.path, src);
    r.g = __shfl_sync(full, h.g, src), r.pixel = __shfl_sync(full, h.pixel, src), r.ctrBase = __shfl_sync(full, h.ctrBase, src);
    r.cnt = __shfl_sync(full, h.cnt, src);
    r.o = v3(__shfl_sync(full, h.o.x, src), __shfl_sync(full, h.o.y, src), __shfl_sync(full, h.o.z, src));
    r.f.n = v3(__shfl_sync(full, h.f.n.x, src), __shfl_sync(full, h.f.n.y, src), __shfl_sync(full, h.f.n.z, src));
    r.f.p = v3(__shfl_sync(full, h.f.p.x, src), __shfl_sync(full, h.f.p.y, src), __shfl_sync(full, h.f.p.z, src));
    r.occluder = __shfl_sync(full, h.occluder, src);
    r.valid = __shfl_sync(full, h.valid, src);
    return r;
}

// Samples per lane of the launch: 8 when it has many direct-lighting evaluations (one unranked step in eight instead of one in
// four: C5 shadow pass 13.11 -> 12.58 ms), 4 when it is small (the coarser items leave a tail: a 1/8 share of C5 1.95 -> 2.47 ms,
// C3 0.80 -> 0.89 ms with 8).  Both instantiations are launched; the one not selected returns at once (the queue lengths only
// exist on the device).
constexpr unsigned kRegroupWideUnits = 1250000u; // 20 M shadow rays with 16 samples per evaluation
__device__ __forceinline__ unsigned wf_regroup_samples_per_lane(const DevParams& p, unsigned units)
{
    return units >= kRegroupWideUnits && p.samples_per_hit >= 8 ? 8u : 4u;
}

// levelBegin / levelEnd: the recursion levels this launch covers (the whole frame, or level 0 and the deeper levels in separate
// launches when the chain stage is split); counterIdx: its chunk counter in wb.counts
template <unsigned kGroup>
__global__ void __launch_bounds__(128, CGE_MINB_VIS) wf_vis_regroup_kernel(DevScene s, DevParams p, WaveBuffers wb, Counters* __restrict__ gcnt,
    unsigned levelBegin, unsigned levelEnd, unsigned counterIdx)
{
    __shared__ unsigned char laneOfRank[4][32];
#if CGE_VIS_SHORT_STACK > 0
    __shared__ unsigned stackMem[kVisShortStack][128]; // the traversal stacks (trace.cuh SharedStack)
    SharedStack<kVisShortStack> stk(stackMem, threadIdx.x);
#else
    LocalStack stk;
#endif
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool fold = p.draws_per_hit == 0;
    const unsigned S = p.samples_per_hit;
    const unsigned groups = (S + kGroup - 1) / kGroup;
    // cum: direct-lighting evaluations before level k (the layout of the visibility bytes); work: the evaluations this kernel
    // traces - the same, or with the light-hull pre-pass (p.vis_cull) those of the hard hits only (wb.hard, counts[32 + k])
    const bool culled = p.vis_cull != 0;
    unsigned cum[kMaxLevels + 1], work[kMaxLevels + 1];
    cum[0] = work[0] = 0;
    for (unsigned k = 0; k < p.levels; k++) {
        cum[k + 1] = cum[k] + wb.counts[k] * (fold ? 1u : (1u << k));
        work[k + 1] = work[k] + wb.counts[culled ? 32 + k : k] * (fold ? 1u : (1u << k));
    }
    const unsigned long long itemBase = (unsigned long long)work[levelBegin] * groups;
    const unsigned long long total = (unsigned long long)(work[levelEnd] - work[levelBegin]) * groups;
    unsigned long long nshadow = 0;
    if (!wf_use_visibility_bytes(p, wb) || wf_regroup_samples_per_lane(p, work[levelEnd] - work[levelBegin]) != kGroup)
        return;
    // one shadow ray of hit h, sample sg; visits (optional) counts the nodes it walked; updates h.occluder
    auto trace_sample = [&](RegroupHit& h, unsigned sg, unsigned* visits) {
        const LightSample ls = wf_sample(s, p, sg, h.pixel, h.ctrBase);
        unsigned char v = 1;
        if (ls.shadowed && !shading_is_zero(s, h.f, ls.pos)) {
            nshadow++;
            const vec3 d = ls.pos - h.o;
            float t;
            float4 r5;
            if (h.occluder >= 0 && triangle_rows_hit(s.ftris + size_t(h.occluder) * kTriRows, h.o, d, 1.0f, t, r5)) {
                v = 0;
            } else {
#if CGE_SHADOW_BVH4
                const int blocker = visits ? trace_shadow4_on<true>(stk, s, h.o, d, visits) : trace_shadow4_on<false>(stk, s, h.o, d);
#else
                const int blocker = visits ? trace_shadow_on<true>(stk, s, h.o, d, visits) : trace_shadow_on<false>(stk, s, h.o, d);
#endif
                if (blocker != -1) {
                    v = 0;
                    if (blocker >= 0)
                        h.occluder = blocker;
                }
            }
        }
        wb.vis[size_t(cum[h.k]) * S + size_t(h.path) * S * h.cnt + h.e + size_t(sg) * h.cnt] = v;
    };
    for (;;) {
        unsigned chunk = 0;
        if (lane == 0)
            chunk = atomicAdd(wb.counts + counterIdx, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        const unsigned long long first = (unsigned long long)chunk * 32ull;
        if (first >= total)
            break;
        const unsigned long long item = itemBase + first + lane;
        RegroupHit own {};
        own.occluder = -1;
        own.valid = first + lane < total ? 1u : 0u;
        unsigned visits = 0;
        if (own.valid) {
            unsigned k = 0;
            while (item >= (unsigned long long)work[k + 1] * groups)
                k++;
            const unsigned inLevel = unsigned(item - (unsigned long long)work[k] * groups), cnt = wb.counts[culled ? 32 + k : k];
            const unsigned block = inLevel / cnt, idx = inLevel - block * cnt;
            own.k = k, own.cnt = wb.counts[k], own.e = culled ? wb.hard[size_t(k) * wb.cap + idx] : idx;
            own.path = block / groups, own.g = block - own.path * groups;
            const uint2 m = wb.meta[size_t(k) * wb.cap + own.e];
            own.pixel = wf_pixel(m);
            own.ctrBase = wf_unit_ctr(p, k, own.path, m);
            const float* b = wb.rec + (size_t(k) * kWaveRecFloats + 17) * wb.cap + own.e;
            own.o = v3(b[0], b[wb.cap], b[2 * size_t(wb.cap)]);
            if (s.cull_zero_shading)
                own.f = wf_load_frame(wb, k, own.e);
            trace_sample(own, own.g * kGroup, &visits); // step 0: the hit's first sample of this group (always < S)
        }
        // rank the hits by the length of that ray, longest first (ties by lane, so the ranks are a permutation)
        const unsigned key = visits * 32u + (31u - lane);
        unsigned rank = 0;
        for (unsigned l = 0; l < 32; l++)
            rank += __shfl_sync(0xffffffffu, key, l) > key ? 1u : 0u;
        __syncwarp();
        laneOfRank[warp][rank] = (unsigned char)lane;
        __syncwarp();
        for (unsigned step = 1; step < kGroup; step++) {
            const unsigned w = (step - 1u) * 32u + lane; // the w-th remaining ray in rank order
            const unsigned src = laneOfRank[warp][w / (kGroup - 1u)];
            RegroupHit h = regroup_fetch(own, src);
            const unsigned sg = h.g * kGroup + 1u + w % (kGroup - 1u);
            if (h.valid && sg < S)
                trace_sample(h, sg, nullptr);
        }
    }
    Counters cnt {};
    cnt.shadow = nshadow;
    flush_counters(cnt, gcnt);
}

template <bool kLookup>
__global__ void __launch_bounds__(128, 8) wf_shade_kernel(DevScene s, DevParams p, WaveBuffers wb, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const bool fold = p.draws_per_hit == 0;
    // work items: level-major, inside a level copy-major, inside a copy queue order
    unsigned cum[kMaxLevels + 1];
    cum[0] = 0;
    for (unsigned k = 0; k < p.levels; k++)
        cum[k + 1] = cum[k] + wb.counts[k] * (fold ? 1u : (1u << k));
    const unsigned total = cum[p.levels];
    unsigned long long nshadow = 0;
    const unsigned S = p.samples_per_hit;
    if (wf_use_visibility_bytes(p, wb) != kLookup)
        return; // the other instantiation handles this launch
    for (;;) {
        unsigned chunk = 0;
        if (lane == 0)
            chunk = atomicAdd(wb.counts + 17, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        // deepest levels first: their items are the most numerous copies of the same hits
        const unsigned long long first = (unsigned long long)chunk * 32ull;
        if (first >= total)
            break;
        const unsigned g = total - 1u - unsigned(first) - lane; // reversed numbering, lane-consecutive
        if (unsigned(first) + lane >= total)
            continue;
        unsigned k = 0;
        while (g >= cum[k + 1])
            k++;
        const unsigned inLevel = g - cum[k];
        const unsigned cnt = wb.counts[k];
        const unsigned path = inLevel / cnt, e = inLevel - path * cnt;
        const uint2 m = wb.meta[size_t(k) * wb.cap + e];
        const unsigned ctr = wf_unit_ctr(p, k, path, m);
        const float* b = wb.rec + (size_t(k) * kWaveRecFloats) * wb.cap + e;
        const size_t c = wb.cap;
        HitRec h;
        h.ray.o = v3(b[0 * c], b[1 * c], b[2 * c]);
        h.ray.d = v3(b[3 * c], b[4 * c], b[5 * c]);
        h.ray.t = b[6 * c];
        h.normal = v3(b[7 * c], b[8 * c], b[9 * c]);
        h.m.kd = v3(b[10 * c], b[11 * c], b[12 * c]);
        h.m.ks = v3(b[13 * c], b[14 * c], b[15 * c]);
        h.m.shininess = b[16 * c];
        // ray index of this evaluation's sample 0 (see wf_vis_regroup_kernel); consecutive samples are cnt bytes apart
        const unsigned char* vis = kLookup ? wb.vis + (size_t(cum[k]) * S + size_t(path) * S * cnt + e) : nullptr;
        const vec3 d = wf_direct<kLookup>(s, p, h, wf_pixel(m), ctr, nshadow, vis, cnt);
        float* out = wb.dir + wf_dir_off(p, wb.cap, k) + (size_t(path) * 3u) * wb.cap + e;
        out[0] = d.x;
        out[c] = d.y;
        out[2 * c] = d.z;
    }
    Counters cnt {};
    cnt.shadow = nshadow;
    flush_counters(cnt, gcnt);
}

__global__ void __launch_bounds__(128) wf_fold_kernel(DevParams p, WaveBuffers wb, float* __restrict__ rgb)
{
    const unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (e0 >= wb.counts[0])
        return;
    const bool fold = p.draws_per_hit == 0;
    const uint2 m = wb.meta[e0];
    const int n = int(m.y & 255u);
    const bool missEnd = (m.y & 256u) != 0;
    unsigned slot[kMaxLevels];
    slot[0] = e0;
    for (int k = 1; k < n; k++)
        slot[k] = wb.next[size_t(k - 1) * wb.cap + slot[k - 1]];
    const size_t c = wb.cap;
    auto dirAt = [&](unsigned k, unsigned path) {
        const float* d = wb.dir + wf_dir_off(p, wb.cap, k) + (size_t(path) * 3u) * c + slot[k];
        return v3(d[0], d[c], d[2 * c]);
    };
    vec3 out;
    if (fold) {
        vec3 val = dirAt(unsigned(n - 1), 0);
        if (missEnd)
            val = (val + v3(0.0f)) + v3(0.0f);
        for (int k = n - 2; k >= 0; k--)
            val = (dirAt(unsigned(k), 0) + val) + val;
        out = val;
    } else {
        vec3 acc[kMaxLevels];
        unsigned char state[kMaxLevels];
        int level = 0;
        unsigned path = 0;
        acc[0] = dirAt(0, 0);
        state[0] = 0;
        for (;;) {
            const bool spawned = (level < n - 1) || missEnd;
            if (!spawned || state[level] == 2) {
                const vec3 v = acc[level];
                if (level == 0) {
                    out = v;
                    break;
                }
                level--;
                path >>= 1;
                acc[level] = acc[level] + v;
                state[level]++;
            } else if (level + 1 < n) {
                path = path * 2u + state[level];
                level++;
                acc[level] = dirAt(unsigned(level), path);
                state[level] = 0;
            } else {
                acc[level] = acc[level] + v3(0.0f); // the reflected copy missed (src/render.cpp:148)
                state[level]++;
            }
        }
    }
    const unsigned pixel = wf_pixel(m);
    const unsigned px = pixel % unsigned(p.width), py = pixel / unsigned(p.width);
    if (p.aa_side) { // one of the pixel's n x n camera rays: parked for wf_resolve_kernel, which adds them in the reference's order
        const unsigned nSub = p.aa_side * p.aa_side;
        const unsigned tile = (py / kTileH) * p.n_tiles_x + px / kTileW;
        const unsigned k = part_entry_of(p.part_unit, p.part_index, p.part_count, tile) - p.tile_first; // position of the tile in this launch's list
        const size_t at = (size_t(k) * 32u + (py % kTileH) * kTileW + px % kTileW) * nSub + wf_sub_ray(m);
        const size_t plane = size_t(p.tile_count) * 32u * nSub;
        wb.sub[at] = out.x, wb.sub[plane + at] = out.y, wb.sub[2 * plane + at] = out.z;
        return;
    }
    const size_t idx = out_pixel_index(p, int(px), int(py));
    rgb[idx * 3 + 0] = out.x;
    rgb[idx * 3 + 1] = out.y;
    rgb[idx * 3 + 2] = out.z;
}

// extra.enableMultipleRaysPerPixel (src/render.cpp:295-303,322): color = sum of the camera rays' colours in their order,
// color /= n * n, colorSum = 0 + color, finalColor = colorSum / float(1)
__global__ void __launch_bounds__(128) wf_resolve_kernel(DevParams p, WaveBuffers wb, float* __restrict__ rgb)
{
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned k = g / 32u, lane = g % 32u;
    if (k >= p.tile_count)
        return;
    const unsigned tile = part_tile_of(p.part_unit, p.part_index, p.part_count, p.tile_first + k);
    const int x = int(tile % p.n_tiles_x) * kTileW + int(lane % kTileW), y = int(tile / p.n_tiles_x) * kTileH + int(lane / kTileW);
    if (x >= p.width || y >= p.height)
        return;
    const unsigned nSub = p.aa_side * p.aa_side;
    const size_t plane = size_t(p.tile_count) * 32u * nSub;
    const float* q = wb.sub + size_t(g) * nSub;
    vec3 color = v3(0.0f);
    for (unsigned sub = 0; sub < nSub; sub++)
        color = color + v3(q[sub], q[plane + sub], q[2 * plane + sub]);
    color = color / float(int(nSub));
    const vec3 out = (v3(0.0f) + color) / 1.0f;
    const size_t idx = out_pixel_index(p, x, y);
    rgb[idx * 3 + 0] = out.x;
    rgb[idx * 3 + 1] = out.y;
    rgb[idx * 3 + 2] = out.z;
}

} // namespace cge
