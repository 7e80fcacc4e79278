// wavefront.cuh — the production pipeline of CGE_TRAVERSAL_FAST: three kernels over compact queues in HBM.
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
//   wf_shade_kernel  one lane = one direct-lighting evaluation (pixel, level, reflection copy) = one
//                    computeLightContribution call of the reference: all its shadow rays, in the reference's sample
//                    order.  Work items are numbered copy-major inside a level, so a warp holds 32 neighbouring pixels
//                    evaluating the SAME sample sequence: coherent rays, equal trip counts, full warps.
//   wf_fold_kernel   one lane = one pixel with a primary hit: folds the 2-ary reflection recursion from the stored
//                    direct terms (reference src/render.cpp:100,118: Lo = (direct + R1) + R2) and writes the pixel.
//
// Queue layout (cap = pixels covered by this launch):
//   rec   [level][kRecFloats][cap] float   hit record, structure of arrays
//   meta  [level][cap] uint2                .x = pixel index (y*W + x, reference coordinates), .y = n | missEnd << 8
//   next  [level][cap] uint                 queue slot of the same pixel at level+1 (valid while level+1 < n)
//   dir   block of level k at dirOff(k): [copy][3][cap] float
#pragma once
#include "render_kernels.cuh"

namespace cge {

struct WaveBuffers {
    float* rec;
    uint2* meta;
    unsigned* next;
    float* dir;
    unsigned* counts;   // [0..15] queue length per level, [16] chain tile counter, [17] shade chunk counter
    unsigned cap;
};

__device__ __forceinline__ size_t wf_dir_off(const DevParams& p, unsigned cap, unsigned k)
{
    const unsigned unitsBefore = p.draws_per_hit == 0 ? k : ((1u << k) - 1u);
    return size_t(unitsBefore) * 3u * cap;
}

__global__ void __launch_bounds__(128, 6) wf_chain_kernel(DevScene s, DevCamera cam, DevParams p, WaveBuffers wb, float* __restrict__ rgb,
    int* __restrict__ ids, Counters* __restrict__ gcnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned below = (1u << lane) - 1u;
    const bool recursive = p.features & CGE_FEAT_RECURSIVE;
    Counters cnt {};
    int x, y;
    while (next_tile(p, wb.counts + 16, lane, x, y)) {
        const bool live = x < p.width && y < p.height;
        const unsigned pixel = unsigned(y) * unsigned(p.width) + unsigned(x);
        Ray ray {};
        if (live)
            ray = generate_ray(cam, x, y, p.width, p.height);
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
                resolve_hit(s, p.features, s.ftris + size_t(h.prim) * kTriRows, h.gid, ray, r);
                float* b = wb.rec + (size_t(level) * kRecFloats) * wb.cap + slot;
                const size_t c = wb.cap;
                b[0 * c] = r.ray.o.x, b[1 * c] = r.ray.o.y, b[2 * c] = r.ray.o.z;
                b[3 * c] = r.ray.d.x, b[4 * c] = r.ray.d.y, b[5 * c] = r.ray.d.z;
                b[6 * c] = r.ray.t;
                b[7 * c] = r.normal.x, b[8 * c] = r.normal.y, b[9 * c] = r.normal.z;
                b[10 * c] = r.m.kd.x, b[11 * c] = r.m.kd.y, b[12 * c] = r.m.kd.z;
                b[13 * c] = r.m.ks.x, b[14 * c] = r.m.ks.y, b[15 * c] = r.m.ks.z;
                b[16 * c] = r.m.shininess;
                if (level > 0)
                    wb.next[size_t(level - 1) * wb.cap + slots[level - 1]] = slot;
                else if (ids)
                    ids[size_t(p.height - 1 - y) * size_t(p.width) + size_t(x)] = int(h.gid);
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
            const unsigned tag = unsigned(n) | (missEnd ? 256u : 0u);
            for (int k = 0; k < n; k++)
                wb.meta[size_t(k) * wb.cap + slots[k]] = make_uint2(pixel, tag);
            if (n == 0)
                store_pixel(p, rgb, ids, x, y, v3(0.0f), -1); // primary miss: black (reference src/render.cpp:148)
        }
    }
    flush_counters(cnt, gcnt);
}

// computeLightContribution for one (pixel, level, copy): same arithmetic and order as PixelTracer::direct
__device__ __forceinline__ vec3 wf_direct(const DevScene& s, const DevParams& p, const HitRec& h, unsigned pixel, unsigned ctr,
    unsigned long long& nshadow)
{
    const vec3 sp = shadow_origin(h);
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
            if (ls.shadowed) {
                nshadow++;
                vis = trace_fast<true>(s, sp, ls.pos - sp, 1.0f).prim >= 0 ? 0.0f : 1.0f;
            }
            result = result + c * vis;
        } else if (samples) {
            vec3 color = v3(0.0f);
            for (unsigned si = 0; si < samples; si++) {
                const LightSample ls = sample_light(L, type, int(si), p, pixel, ctr);
                nshadow++;
                const float vis = trace_fast<true>(s, sp, ls.pos - sp, 1.0f).prim >= 0 ? 0.0f : 1.0f;
                color = color + compute_shading(ls.pos, ls.col, h) * vis;
            }
            const float denom = type == CGE_LIGHT_SEGMENT ? float(p.segment_samples)
                                                          : fmul(float(p.parallelogram_samples), float(p.parallelogram_samples));
            result = result + color / denom;
        }
        ctr += draws;
    }
    return result;
}

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
        const unsigned nChain = m.y & 255u;
        unsigned ctr = 0;
        if (!fold) {
            // first draw index of this evaluation in the reference's depth-first order over the 2-ary recursion:
            // index = sum_{i=1..k} (1 + b_i * (2^(n-i) - 1)),  b_i = i-th copy choice on the way down
            unsigned idx = 0;
            for (unsigned i = 1; i <= k; i++) {
                const unsigned b = (path >> (k - i)) & 1u;
                idx += 1u + b * ((1u << (nChain - i)) - 1u);
            }
            ctr = idx * p.draws_per_hit;
        }
        const float* b = wb.rec + (size_t(k) * kRecFloats) * wb.cap + e;
        const size_t c = wb.cap;
        HitRec h;
        h.ray.o = v3(b[0 * c], b[1 * c], b[2 * c]);
        h.ray.d = v3(b[3 * c], b[4 * c], b[5 * c]);
        h.ray.t = b[6 * c];
        h.normal = v3(b[7 * c], b[8 * c], b[9 * c]);
        h.m.kd = v3(b[10 * c], b[11 * c], b[12 * c]);
        h.m.ks = v3(b[13 * c], b[14 * c], b[15 * c]);
        h.m.shininess = b[16 * c];
        const vec3 d = wf_direct(s, p, h, m.x, ctr, nshadow);
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
    const unsigned px = m.x % unsigned(p.width), py = m.x / unsigned(p.width);
    const size_t idx = size_t(p.height - 1 - int(py)) * size_t(p.width) + size_t(px);
    rgb[idx * 3 + 0] = out.x;
    rgb[idx * 3 + 1] = out.y;
    rgb[idx * 3 + 2] = out.z;
}

} // namespace cge
