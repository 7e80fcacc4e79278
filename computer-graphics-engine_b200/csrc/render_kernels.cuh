// render_kernels.cuh — the two render kernels.
//
//   render_kernel       one thread = one pixel, everything sequential per thread: the literal restatement of
//                       renderRayTracing -> getFinalColor -> recursiveRayTrace (reference src/render.cpp:27-155,
//                       273-329).  Used for CGE_TRAVERSAL_REFERENCE (validation mode), for scenes with spheres, and
//                       as the fallback of the cooperative kernel.
//   render_coop_kernel  the production kernel (CGE_TRAVERSAL_FAST): one warp = one 8x4 pixel tile, three phases:
//                         A  every lane traces its pixel's mirror chain (closest hits) and parks one hit record per
//                            level in shared memory;
//                         B  the warp's direct-lighting work — every (pixel, level, reflection copy, light, sample) —
//                            is enumerated with a warp prefix sum and dealt out to ALL 32 lanes, consecutive lanes
//                            taking consecutive samples of the same hit (coherent shadow rays); per-sample terms are
//                            summed in the reference's order through warp shuffles;
//                         C  every lane folds its pixel's 2-ary reflection recursion from the parked terms.
//                       Persistent CTAs pull tiles from a global counter (cost per tile is wildly non-uniform).
#pragma once
#include "shade.cuh"

namespace cge {

constexpr int kMaxLevels = kMaxRayDepth + 1;

// minimum resident 128-thread CTAs per SM the kernels are compiled for (register budget = 65536 / (128 * N));
// chosen from the A/B runs recorded in profiles/ (see DESIGN.md "Occupancy").
#ifndef CGE_MINB_THREAD
#define CGE_MINB_THREAD 8
#endif
#ifndef CGE_MINB_COOP
#define CGE_MINB_COOP 3
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
            return kCount ? trace_fast<true, kCount>(s, o, d, 1.0f, &nbox, &ntri).prim >= 0 : trace_shadow(s, o, d) >= 0;
        return trace_reference<kSpheres, kCount>(s, o, d, 1.0f, nbox, ntri).prim >= 0;
    }
    __device__ const float4* rows(const Hit& h) const { return (kFast ? s.ftris : s.tris) + size_t(h.prim) * kTriRows; }

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
    if (k >= p.tile_count)
        return false;
    if (kOut)
        *kOut = k;
    // multi-GPU: the tile list is interleaved across ranks, entry k of this rank's list is tile part_index + k * part_count
    const unsigned tile = p.part_index + (p.tile_first + k) * p.part_count;
    x = int(tile % p.n_tiles_x) * kTileW + int(lane % kTileW);
    y = int(tile / p.n_tiles_x) * kTileH + int(lane / kTileW);
    return true;
}

__device__ __forceinline__ void store_pixel(const DevParams& p, float* __restrict__ rgb, int* __restrict__ ids, int x, int y, vec3 c,
    int primId)
{
    const size_t idx = size_t(p.height - 1 - y) * size_t(p.width) + size_t(x); // Screen::setPixel y flip (src/screen.cpp:45)
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

// -----------------------------------------------------------------------------------------------------------------
// cooperative kernel (fast tree, triangles only, shading on, 1 <= samples_per_hit <= 32)
// -----------------------------------------------------------------------------------------------------------------
// Shared memory per warp (floats):  rec[levels][kRecFloats][32] | dir[units_per_lane][3][32] | pref[33]
__host__ __device__ inline unsigned coop_warp_floats(unsigned levels, unsigned units)
{
    return levels * kRecFloats * 32u + units * 3u * 32u + 40u;
}

__device__ __forceinline__ void rec_store(float* rec, unsigned level, unsigned lane, const HitRec& h)
{
    float* b = rec + size_t(level) * kRecFloats * 32 + lane;
    b[0 * 32] = h.ray.o.x, b[1 * 32] = h.ray.o.y, b[2 * 32] = h.ray.o.z;
    b[3 * 32] = h.ray.d.x, b[4 * 32] = h.ray.d.y, b[5 * 32] = h.ray.d.z;
    b[6 * 32] = h.ray.t;
    b[7 * 32] = h.normal.x, b[8 * 32] = h.normal.y, b[9 * 32] = h.normal.z;
    b[10 * 32] = h.m.kd.x, b[11 * 32] = h.m.kd.y, b[12 * 32] = h.m.kd.z;
    b[13 * 32] = h.m.ks.x, b[14 * 32] = h.m.ks.y, b[15 * 32] = h.m.ks.z;
    b[16 * 32] = h.m.shininess;
}
__device__ __forceinline__ HitRec rec_load(const float* rec, unsigned level, unsigned lane)
{
    const float* b = rec + size_t(level) * kRecFloats * 32 + lane;
    HitRec h;
    h.ray.o = v3(b[0 * 32], b[1 * 32], b[2 * 32]);
    h.ray.d = v3(b[3 * 32], b[4 * 32], b[5 * 32]);
    h.ray.t = b[6 * 32];
    h.normal = v3(b[7 * 32], b[8 * 32], b[9 * 32]);
    h.m.kd = v3(b[10 * 32], b[11 * 32], b[12 * 32]);
    h.m.ks = v3(b[13 * 32], b[14 * 32], b[15 * 32]);
    h.m.shininess = b[16 * 32];
    return h;
}

__global__ void __launch_bounds__(128, CGE_MINB_COOP) render_coop_kernel(DevScene s, DevCamera cam, DevParams p, float* __restrict__ rgb,
    int* __restrict__ ids, unsigned* __restrict__ tileCounter, Counters* __restrict__ gcnt)
{
    extern __shared__ float smem[];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* rec = smem + size_t(warp) * coop_warp_floats(p.levels, p.units_per_lane);
    float* dir = rec + size_t(p.levels) * kRecFloats * 32;
    unsigned* pref = reinterpret_cast<unsigned*>(dir + size_t(p.units_per_lane) * 3 * 32);

    const bool fold = p.draws_per_hit == 0;
    const bool recursive = p.features & CGE_FEAT_RECURSIVE;
    const unsigned S = p.samples_per_hit; // 1..32, warp-uniform
    const unsigned P = 32u / S;           // hits shaded concurrently by one warp
    Counters cnt {};
    int x, y;
    while (next_tile(p, tileCounter, lane, x, y)) {
        const bool live = x < p.width && y < p.height;
        // ---- phase A: the pixel's mirror chain ----------------------------------------------------------------------
        int n = 0, primId = -1;
        bool missEnd = false;
        if (live) {
            Ray ray = generate_ray(cam, x, y, p.width, p.height);
            for (int level = 0;; level++) {
                const Hit h = trace_fast<false>(s, ray.o, ray.d, ray.t);
                if (level == 0)
                    cnt.primary++;
                else
                    cnt.bounce++;
                if (h.prim < 0) {
                    missEnd = true;
                    break;
                }
                ray.t = h.t;
                HitRec r;
                resolve_hit(s, p.features, s.ftris + size_t(h.prim) * kTriRows, h.gid, ray, r);
                rec_store(rec, unsigned(level), lane, r);
                if (level == 0)
                    primId = int(h.gid);
                n = level + 1;
                if (!recursive || level >= p.ray_depth)
                    break;
                Ray next;
                if (!reflection_ray(r, next))
                    break;
                ray = next;
            }
            reference_calls(cnt, n, missEnd, p.shadow_rays_per_hit);
        }
        // ---- phase B: direct lighting of every (pixel, level, copy), dealt out to all lanes --------------------------
        const unsigned myUnits = fold ? unsigned(n) : ((1u << n) - 1u);
        unsigned incl = myUnits;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= unsigned(off))
                incl += v;
        }
        pref[lane + 1] = incl;
        if (lane == 0)
            pref[0] = 0;
        __syncwarp();
        const unsigned U = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned g = lane / S, j = lane % S;
        for (unsigned base = 0; base < U; base += P) {
            const unsigned u = base + g;
            const bool active = g < P && u < U;
            vec3 term = v3(0.0f);
            unsigned src = 0, ul = 0;
            if (active) {
                // owner lane of unit u: pref[src] <= u < pref[src+1]
                unsigned lo = 0, hi = 32;
                while (hi - lo > 1) {
                    const unsigned mid = (lo + hi) >> 1;
                    if (pref[mid] <= u)
                        lo = mid;
                    else
                        hi = mid;
                }
                src = lo;
                ul = u - pref[src];
                // (level k, copy path) of the unit and its first draw index in the reference's depth-first order:
                // index = sum_{i=1..k} (1 + b_i * (2^(nSrc-i) - 1)),  b_i = i-th copy choice on the way down
                unsigned k = ul, path = 0, ctr = 0;
                if (!fold) {
                    k = 31u - unsigned(__clz(int(ul + 1u)));
                    path = ul + 1u - (1u << k);
                    const unsigned nSrc = 32u - unsigned(__clz(int(pref[src + 1] - pref[src]))); // units = 2^n - 1
                    unsigned idx = 0;
                    for (unsigned i = 1; i <= k; i++) {
                        const unsigned b = (path >> (k - i)) & 1u;
                        idx += 1u + b * ((1u << (nSrc - i)) - 1u);
                    }
                    ctr = idx * p.draws_per_hit;
                }
                const HitRec h = rec_load(rec, k, src);
                // task j -> (light, sample)
                unsigned li = 0, si = j, samples = 0, draws = 0, type = 0;
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
                // pixel id of the owner lane = this tile's pixel of lane `src`
                const int sx = x - int(lane % kTileW) + int(src % kTileW), sy = y - int(lane / kTileW) + int(src / kTileW);
                const unsigned spx = unsigned(sy) * unsigned(p.width) + unsigned(sx);
                const LightSample ls = sample_light(L, type, int(si), p, spx, ctr);
                float vis = 1.0f;
                if (ls.shadowed && !shading_is_zero(s, shade_frame(h), ls.pos)) {
                    const vec3 so = shadow_origin(h);
                    cnt.shadow++;
                    vis = trace_shadow(s, so, ls.pos - so) >= 0 ? 0.0f : 1.0f;
                }
                term = compute_shading(ls.pos, ls.col, h) * vis;
            }
            // ordered reduction: the group leader adds the terms exactly as computeLightContribution does
            vec3 result = v3(0.0f);
            unsigned idx = 0;
            const unsigned gbase = (g < P ? g : 0u) * S;
            for (unsigned li = 0; li < s.n_lights; li++) {
                const unsigned type = __float_as_uint(__ldg(s.lights + size_t(li) * kLightFloats));
                unsigned samples, draws;
                light_counts(type, p, samples, draws);
                if (samples == 0)
                    continue;
                if (type == CGE_LIGHT_POINT) {
                    const unsigned from = gbase + idx;
                    result = result + v3(__shfl_sync(0xffffffffu, term.x, from), __shfl_sync(0xffffffffu, term.y, from),
                                          __shfl_sync(0xffffffffu, term.z, from));
                } else {
                    vec3 color = v3(0.0f);
                    for (unsigned q = 0; q < samples; q++) {
                        const unsigned from = gbase + idx + q;
                        color = color + v3(__shfl_sync(0xffffffffu, term.x, from), __shfl_sync(0xffffffffu, term.y, from),
                                            __shfl_sync(0xffffffffu, term.z, from));
                    }
                    const float denom = type == CGE_LIGHT_SEGMENT ? float(p.segment_samples)
                                                                  : fmul(float(p.parallelogram_samples), float(p.parallelogram_samples));
                    result = result + color / denom;
                }
                idx += samples;
            }
            if (active && j == 0) {
                float* d = dir + size_t(ul) * 3 * 32 + src;
                d[0] = result.x, d[32] = result.y, d[64] = result.z;
            }
        }
        __syncwarp();
        // ---- phase C: fold the 2-ary recursion of this lane's pixel -------------------------------------------------
        if (live) {
            vec3 out = v3(0.0f);
            auto dirAt = [&](unsigned unit) {
                const float* d = dir + size_t(unit) * 3 * 32 + lane;
                return v3(d[0], d[32], d[64]);
            };
            if (n > 0) {
                if (fold) {
                    vec3 val = dirAt(unsigned(n - 1));
                    if (missEnd)
                        val = (val + v3(0.0f)) + v3(0.0f);
                    for (int k = n - 2; k >= 0; k--)
                        val = (dirAt(unsigned(k)) + val) + val;
                    out = val;
                } else {
                    vec3 acc[kMaxLevels];
                    unsigned char state[kMaxLevels];
                    int level = 0;
                    unsigned path = 0;
                    acc[0] = dirAt(0);
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
                            acc[level] = dirAt((1u << level) - 1u + path);
                            state[level] = 0;
                        } else {
                            acc[level] = acc[level] + v3(0.0f);
                            state[level]++;
                        }
                    }
                }
            }
            store_pixel(p, rgb, ids, x, y, out, primId);
        }
        __syncwarp();
    }
    flush_counters(cnt, gcnt);
}

} // namespace cge
