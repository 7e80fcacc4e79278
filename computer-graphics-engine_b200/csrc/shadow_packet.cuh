// shadow_packet.cuh — the shadow pass of the wavefront pipeline: ONE tree walk per lane for a packet of light samples.
//
// The shadow rays of one computeLightContribution call (reference src/light.cpp:108-164) leave the same point (the offset hit
// point of src/light.cpp:54-58) and end on one small light, and testVisibilityLightSample (src/light.cpp:49-73) only asks
// whether ANY primitive is accepted with 0 <= t <= 1.  The kernels this one replaces walked the tree once per ray: 13 inner
// nodes per ray, every lane re-descending the ~24 levels around its own origin for each of its samples.  Here a
// lane owns a packet of up to kPacket samples of one light for one hit and walks the tree ONCE:
//
//   * inner nodes are tested against the packet's HULL: the pyramid { o + t * d : t in [0, 1], d in [dmin, dmax] } with
//     dmin / dmax the per-axis bounds of the packet's ray directions.  Per axis that is two linear constraints on t,
//         o + t * dmin <= hi   and   o + t * dmax >= lo ,
//     each a lower or an upper bound of t depending on the sign of dmin / dmax, evaluated as one FFMA per plane exactly like
//     the single-ray slab test of trace.cuh (an axis whose directions have both signs yields two lower bounds and no upper one).
//     The test is conservative - it can only ADD visits - and gates which leaves are reached, never a hit decision;
//   * at a leaf every still-undecided ray of the packet is tested against every triangle with the unchanged libIntersect
//     arithmetic (I2-I4, trace.cuh triangle_rows_hit): the plane numerator D - dot(o, n) is the same for all rays of the packet
//     and evaluated once, and a triangle whose denominator range over the hull cannot give 0 <= t <= 1 is skipped for the whole
//     packet.  A ray leaves the packet when a triangle accepts it; the walk ends when no ray is left or the stack is empty.
//
// Any-hit is order-free, every ray meets (at least) the leaves it would have met alone, and its accept decisions are made by
// the same instructions on the same operands: the visibility bytes are bit-identical to the per-ray kernels'.
// The traversal stack lives in shared memory (kShort entries per lane, lane-interleaved so that a push or pop of a warp is one
// conflict-free access); deeper pushes spill to a small local array (never observed on the configs; the near-first walk of a
// binary tree holds at most one entry per level).  The packet's directions are parked in shared memory as well (exact bits
// are needed at the leaves; 12 bytes per ray).
#pragma once
#include "wavefront.cuh"

namespace cge {

#ifndef CGE_PACKET_MINB16
#define CGE_PACKET_MINB16 6
#endif
#ifndef CGE_PACKET_MINB8
#define CGE_PACKET_MINB8 8
#endif
#ifndef CGE_PACKET_MINB4
#define CGE_PACKET_MINB4 10
#endif

template <unsigned kPacket>
struct PacketCfg {
    static constexpr unsigned kShort = kPacket >= 16 ? 12u : 16u; // shared-memory stack entries per lane
    static constexpr unsigned kMinBlocks = kPacket >= 16 ? CGE_PACKET_MINB16 : kPacket >= 8 ? CGE_PACKET_MINB8 : CGE_PACKET_MINB4;
};

// groups (packets) one direct-lighting evaluation is split into: every light's samples in chunks of kPacket, never across lights
template <unsigned kPacket>
__device__ __forceinline__ unsigned packet_groups(const DevScene& s, const DevParams& p)
{
    unsigned groups = 0;
    for (unsigned li = 0; li < s.n_lights; li++) {
        unsigned samples, draws;
        light_counts(__float_as_uint(__ldg(s.lights + size_t(li) * kLightFloats)), p, samples, draws);
        groups += (samples + kPacket - 1) / kPacket;
    }
    return groups;
}

// (HullAxis / hull_axis: wavefront.cuh, shared with the light-hull pre-pass)

template <unsigned kPacket>
__global__ void __launch_bounds__(128, PacketCfg<kPacket>::kMinBlocks) wf_vis_packet_kernel(DevScene s, DevParams p, WaveBuffers wb,
    Counters* __restrict__ gcnt)
{
    constexpr unsigned kShort = PacketCfg<kPacket>::kShort;
    constexpr unsigned kOverflow = kFastStackSize - kShort;
    __shared__ float sdir[kPacket * 3][128];          // the packet's ray directions (exact bits are needed at the leaves)
    __shared__ unsigned sstk[kShort][128];            // traversal stack, lane-interleaved
    __shared__ float sorg[3][128];                    // phase 2: the packet's origin ...
    __shared__ unsigned svis[3][128];                 // ... and where its visibility bytes go (address lo / hi, stride)
    __shared__ unsigned short sitem[4][32 * kPacket]; // phase 2: the warp's undecided rays, (lane << 8 | ray)
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wbase = tid & ~31u;
    const unsigned below = (1u << lane) - 1u;
    const bool fold = p.draws_per_hit == 0;
    const unsigned S = p.samples_per_hit;
    const unsigned groups = packet_groups<kPacket>(s, p);
    unsigned cum[kMaxLevels + 1]; // in units (direct-lighting evaluations)
    cum[0] = 0;
    for (unsigned k = 0; k < p.levels; k++)
        cum[k + 1] = cum[k] + wb.counts[k] * (fold ? 1u : (1u << k));
    const unsigned long long total = (unsigned long long)cum[p.levels] * groups;
    unsigned long long nshadow = 0;
    constexpr unsigned kDone = 0x7fffffffu;
    constexpr float kInf = __builtin_huge_valf();
#ifdef CGE_PACKET_STATS
    unsigned long long stHull = 0, stRay = 0, stTri = 0, stDef = 0, stFull = 0, stTri2 = 0;
#define PSTAT(x) x
#else
#define PSTAT(x)
#endif

    // the traversal stack of this lane: kShort entries in shared memory, the (rare) rest in local memory
    unsigned overflow[kOverflow];
    unsigned sp = 0;
    auto push = [&](unsigned v) {
        if (sp < kShort)
            sstk[sp][tid] = v;
        else
            overflow[sp - kShort] = v;
        sp++;
    };
    auto pop = [&]() -> unsigned {
        if (sp == 0)
            return kDone;
        sp--;
        return sp < kShort ? sstk[sp][tid] : overflow[sp - kShort];
    };

    for (;;) {
        unsigned chunk = 0;
        if (lane == 0)
            chunk = atomicAdd(wb.counts + 18, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        const unsigned long long first = (unsigned long long)chunk * 32ull;
        if (first >= total)
            break;
        const unsigned long long item = first + lane;
        unsigned undecided = 0; // rays of this lane's packet not yet known to be blocked
        if (item < total) {
            // item -> (level k, copy, packet, slot): level-major, then copy, then packet, then queue order (32 neighbouring hits per warp)
            unsigned k = 0;
            while (item >= (unsigned long long)cum[k + 1] * groups)
                k++;
            const unsigned inLevel = unsigned(item - (unsigned long long)cum[k] * groups), cnt = wb.counts[k];
            const unsigned block = inLevel / cnt, e = inLevel - block * cnt;
            const unsigned path = block / groups;
            unsigned g = block - path * groups;
            const uint2 m = wb.meta[size_t(k) * wb.cap + e];
            const unsigned pixel = wf_pixel(m);
            unsigned ctr = wf_unit_ctr(p, k, path, m);
            // packet g -> light, its first sample and sample count
            const float* L = s.lights;
            unsigned type = 0, sBase = 0, siBeg = 0, siEnd = 0;
            for (unsigned li = 0;; li++) {
                L = s.lights + size_t(li) * kLightFloats;
                type = __float_as_uint(__ldg(L));
                unsigned samples, draws;
                light_counts(type, p, samples, draws);
                const unsigned ng = (samples + kPacket - 1) / kPacket;
                if (g < ng) {
                    siBeg = g * kPacket;
                    siEnd = min(samples, siBeg + kPacket);
                    break;
                }
                g -= ng;
                sBase += samples;
                ctr += draws;
            }
            const float* b = wb.rec + (size_t(k) * kWaveRecFloats + 17) * wb.cap + e;
            const vec3 o = v3(b[0], b[wb.cap], b[2 * size_t(wb.cap)]);
            ShadeFrame frame {};
            if (s.cull_zero_shading)
                frame = wf_load_frame(wb, k, e);
            unsigned char* visOut = wb.vis + (size_t(cum[k]) * S + size_t(path) * S * cnt + e) + size_t(sBase + siBeg) * cnt;

            // ---- the packet: directions of the samples that need a ray, their per-axis bounds ---------------------------------
            unsigned active = 0;
            vec3 dmin = v3(kInf), dmax = v3(-kInf);
            for (unsigned i = 0; i < siEnd - siBeg; i++) {
                const LightSample ls = sample_light(L, type, int(siBeg + i), p, pixel, ctr);
                if (!ls.shadowed || shading_is_zero(s, frame, ls.pos))
                    continue; // the reference does not test this sample, or its term is exactly zero: visible
                const vec3 d = ls.pos - o;
                nshadow++;
                if (!(fabsf(d.x) <= 3e38f && fabsf(d.y) <= 3e38f && fabsf(d.z) <= 3e38f))
                    continue; // a non-finite direction is never accepted by the archive's test (NaN fails every comparison): visible
                sdir[i * 3 + 0][tid] = d.x, sdir[i * 3 + 1][tid] = d.y, sdir[i * 3 + 2][tid] = d.z;
                dmin = v3(fminf(dmin.x, d.x), fminf(dmin.y, d.y), fminf(dmin.z, d.z));
                dmax = v3(fmaxf(dmax.x, d.x), fmaxf(dmax.y, d.y), fmaxf(dmax.z, d.z));
                active |= 1u << i;
            }
            undecided = s.n_prims ? active : 0u;
            bool resolved = true; // phase 1 saw every leaf the packet's rays can reach
            PSTAT(stTri2 += active ? 1u : 0u);
            if (undecided) {
                // ---- phase 1: the hull walks the tree while it is thin -------------------------------------------------------
                const HullAxis hx = hull_axis(o.x, dmin.x, dmax.x), hy = hull_axis(o.y, dmin.y, dmax.y), hz = hull_axis(o.z, dmin.z, dmax.z);
                const bool anyMixed = hx.mixed || hy.mixed || hz.mixed;
                // width the hull gains per unit of t, scaled by the threshold: a child box entered at t is "fat" when
                // t * spread > (largest extent of the box): the rays of the packet are then further apart than the box is wide,
                // the hull would visit many more nodes than any one ray, and the packet is handed to phase 2
                const float spread = fmaxf(fmaxf(dmax.x - dmin.x, dmax.y - dmin.y), dmax.z - dmin.z) * p.packet_fat;
                // denominators dot(d, n) of the packet lie within +- this of the interval evaluated from dmin / dmax
                const float dScale = fmaxf(fmaxf(fmaxf(fabsf(dmin.x), fabsf(dmax.x)), fmaxf(fabsf(dmin.y), fabsf(dmax.y))),
                    fmaxf(fabsf(dmin.z), fabsf(dmax.z)));
                const float denEps = dScale * 1e-5f;
                // hull against one child box: entry / exit parameter (entry clamped to 0)
                auto hull_box = [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float& ent, float& ext) {
                    const float ax = __fmaf_rn(hx.sel ? hix : lox, hx.k1, hx.c1), bx = __fmaf_rn(hx.sel ? lox : hix, hx.k2, hx.c2);
                    const float ay = __fmaf_rn(hy.sel ? hiy : loy, hy.k1, hy.c1), by = __fmaf_rn(hy.sel ? loy : hiy, hy.k2, hy.c2);
                    const float az = __fmaf_rn(hz.sel ? hiz : loz, hz.k1, hz.c1), bz = __fmaf_rn(hz.sel ? loz : hiz, hz.k2, hz.c2);
                    ent = fmaxf(max3(ax, ay, az), 0.0f);
                    if (!anyMixed) {
                        ext = min3(bx, by, bz);
                    } else {
                        ent = fmaxf(ent, max3(hx.mixed ? bx : 0.0f, hy.mixed ? by : 0.0f, hz.mixed ? bz : 0.0f));
                        ext = min3(hx.mixed ? kInf : bx, hy.mixed ? kInf : by, hz.mixed ? kInf : bz);
                    }
                };
                sp = 0;
                unsigned cur = s.froot;
                unsigned budget = p.packet_budget; // inner nodes + leaves this walk may still visit before it gives the packet up
                while (cur != kDone) {
                    while (cur < kDone) {
                        const float4* nd = s.fnodes + size_t(cur) * kNodeRows;
                        const float4 q3 = ldg4(nd + 3);
                        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2);
                        float entL, extL, entR, extR;
                        PSTAT(stHull++);
                        if (budget-- == 0) {
                            resolved = false;
                            cur = kDone;
                            break;
                        }
                        hull_box(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
                        hull_box(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
                        const unsigned cl = __float_as_uint(q3.x), cr = __float_as_uint(q3.y);
                        const bool hitL = entL <= extL * 1.000008f && entL <= 1.0001f;
                        const bool hitR = entR <= extR * 1.000008f && entR <= 1.0001f;
                        // an inner child the hull enters where it is fat (q3.z / q3.w: largest extent of the child's box)
                        if ((hitL && cl < kDone && entL * spread > q3.z) || (hitR && cr < kDone && entR * spread > q3.w)) {
                            resolved = false;
                            cur = kDone;
                            break;
                        }
                        const bool leftFirst = hitL && (!hitR || entL <= entR);
                        if (hitL && hitR)
                            push(leftFirst ? cr : cl);
                        if (hitL || hitR)
                            cur = leftFirst ? cl : cr;
                        else
                            cur = pop();
                    }
                    if (cur == kDone)
                        break;
                    const unsigned firstTri = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
                    if (budget < p.packet_leaf_cost) {
                        resolved = false;
                        break;
                    }
                    budget -= p.packet_leaf_cost;
                    for (unsigned i = firstTri; i < firstTri + count && undecided; i++) {
                        const float4* tr = s.ftris + size_t(i) * kTriRows;
                        const float4 r0 = ldg4(tr);
                        const vec3 n = v3(r0.x, r0.y, r0.z);
                        const float num = fsub(r0.w, dot(o, n)); // I2's numerator: the same for every ray of the packet
                        // range of the denominators dot(d, n) over the hull; 0 <= num / den <= 1 needs a den with num's sign and
                        // |den| >= |num| (a NaN plane - a sphere record - fails both comparisons)
                        const float ex0 = n.x * dmin.x, ex1 = n.x * dmax.x, ey0 = n.y * dmin.y, ey1 = n.y * dmax.y, ez0 = n.z * dmin.z,
                                    ez1 = n.z * dmax.z;
                        const float denLo = fminf(ex0, ex1) + fminf(ey0, ey1) + fminf(ez0, ez1) - denEps;
                        const float denHi = fmaxf(ex0, ex1) + fmaxf(ey0, ey1) + fmaxf(ez0, ez1) + denEps;
                        if (!((num >= 0.0f && denHi >= num) || (num <= 0.0f && denLo <= num)))
                            continue;
                        for (unsigned rest = undecided; rest; rest &= rest - 1u) {
                            const unsigned j = unsigned(__ffs(int(rest))) - 1u;
                            const vec3 d = v3(sdir[j * 3 + 0][tid], sdir[j * 3 + 1][tid], sdir[j * 3 + 2][tid]);
                            float t;
                            float4 r5;
                            PSTAT(stTri++);
                            if (triangle_rows_tail(tr, n, num, o, d, 1.0f, t, r5))
                                undecided &= ~(1u << j);
                        }
                    }
                    cur = undecided ? pop() : kDone;
                }
            }
            // visibility so far: blocked rays 0, everything else 1 (phase 2 clears the bytes of the rays it finds blocked)
            const unsigned blocked = s.n_prims ? active & ~undecided : 0u;
            for (unsigned i = 0; i < siEnd - siBeg; i++)
                visOut[size_t(i) * cnt] = (blocked >> i) & 1u ? 0 : 1;
            if (resolved)
                undecided = 0;
            PSTAT(stFull += undecided ? 1ull : 0ull);
            if (undecided) { // publish what phase 2 needs to trace this packet's rays from any lane
                sorg[0][tid] = o.x, sorg[1][tid] = o.y, sorg[2][tid] = o.z;
                const unsigned long long va = reinterpret_cast<unsigned long long>(visOut);
                svis[0][tid] = unsigned(va), svis[1][tid] = unsigned(va >> 32), svis[2][tid] = cnt;
            }
        }
        // ---- phase 2: the warp's undecided rays, dealt evenly to its 32 lanes, each walking the tree alone ----------------------
        // (rays of one packet stay together in one lane or in neighbouring lanes: the blocker of the previous ray is tried first)
        const unsigned nMine = unsigned(__popc(undecided));
        unsigned offset = nMine; // inclusive prefix sum over the lanes
        for (unsigned d = 1; d < 32; d <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, offset, d);
            if (lane >= d)
                offset += v;
        }
        const unsigned totalRays = __shfl_sync(0xffffffffu, offset, 31);
        if (totalRays == 0)
            continue;
        offset -= nMine;
        for (unsigned rest = undecided; rest; rest &= rest - 1u)
            sitem[warp][offset++] = (unsigned short)((lane << 8) | (unsigned(__ffs(int(rest))) - 1u));
        __syncwarp();
        const unsigned per = (totalRays + 31u) / 32u;
        int occluder = -1;
        unsigned occLane = 32;
        for (unsigned q = lane * per; q < min(totalRays, (lane + 1u) * per); q++) {
            const unsigned it = sitem[warp][q], src = wbase + (it >> 8), j = it & 255u;
            const vec3 o = v3(sorg[0][src], sorg[1][src], sorg[2][src]);
            const vec3 d = v3(sdir[j * 3 + 0][src], sdir[j * 3 + 1][src], sdir[j * 3 + 2][src]);
            float t;
            float4 r5;
            bool hit = false;
            // any accepted triangle proves occlusion: first the one that blocked the previous ray of the same packet
            if (occluder >= 0 && occLane == (it >> 8) && triangle_rows_hit(s.ftris + size_t(occluder) * kTriRows, o, d, 1.0f, t, r5)) {
                hit = true;
            } else {
                const SlabRay sr = slab_ray(o, d);
                sp = 0;
                unsigned cur = s.froot;
                while (cur != kDone) {
                    while (cur < kDone) {
                        const float4* nd = s.fnodes + size_t(cur) * kNodeRows;
                        const float4 q3 = ldg4(nd + 3);
                        const float4 q0 = ldg4(nd), q1 = ldg4(nd + 1), q2 = ldg4(nd + 2);
                        float entL, extL, entR, extR;
                        PSTAT(stRay++);
                        slab_box(sr, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, entL, extL);
                        slab_box(sr, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, entR, extR);
                        const unsigned cl = __float_as_uint(q3.x), cr = __float_as_uint(q3.y);
                        const bool hitL = entL <= extL * 1.000002f && entL <= 1.0001f;
                        const bool hitR = entR <= extR * 1.000002f && entR <= 1.0001f;
                        const bool leftFirst = hitL && (!hitR || entL <= entR);
                        if (hitL && hitR)
                            push(leftFirst ? cr : cl);
                        if (hitL || hitR)
                            cur = leftFirst ? cl : cr;
                        else
                            cur = pop();
                    }
                    if (cur == kDone)
                        break;
                    const unsigned firstTri = cur & 0x0fffffffu, count = ((cur >> 28) & 7u) + 1u;
                    for (unsigned i = firstTri; i < firstTri + count; i++)
                        if (triangle_rows_hit(s.ftris + size_t(i) * kTriRows, o, d, 1.0f, t, r5)) {
                            hit = true;
                            occluder = int(i);
                            occLane = it >> 8;
                            break;
                        }
                    cur = hit ? kDone : pop();
                }
            }
            if (hit) {
                unsigned char* vis = reinterpret_cast<unsigned char*>((unsigned long long)svis[0][src] | ((unsigned long long)svis[1][src] << 32));
                vis[size_t(j) * svis[2][src]] = 0;
            }
        }
        __syncwarp(); // the shared records are rewritten by the next chunk
    }
    Counters cnt {};
    cnt.shadow = nshadow;
#ifdef CGE_PACKET_STATS // development build: hull node visits, per-ray node visits, packet triangle tests, packets, packets handed to phase 2
    cnt.box = stHull, cnt.tri = stRay, cnt.primary = stTri, cnt.bounce = stTri2, cnt.reference = stDef, cnt.reference_shadow = stFull;
#endif
    flush_counters(cnt, gcnt);
}

} // namespace cge
