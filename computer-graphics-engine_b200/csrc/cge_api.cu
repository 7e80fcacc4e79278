// cge_api.cu — implementation of the C ABI in include/cge.h: scene flattening + upload, kernel dispatch,
// known-answer entry points, and the multi-GPU tile gather.  No CPU fallback exists anywhere in this file: every
// compute entry point launches CUDA kernels or returns CGE_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "bvh_build.h"
#include "bvh_sah.h"
#include "cge.h"
#include "dev_scene.h"
#include "nccl_min.h"
#include "render_kernels.cuh"
#include "wavefront.cuh"
#ifndef CGE_EXPERIMENTS
#define CGE_EXPERIMENTS 0 // 1: compile the measured-and-rejected variants kept for the record (shadow_packet.cuh) behind their switches
#endif
#if CGE_EXPERIMENTS
#include "shadow_packet.cuh"
#endif


using namespace cge;

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
#define CGE_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return fail(CGE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                 \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t upload(const std::vector<T>& h)
    {
        release();
        n = h.size();
        if (n == 0) {
            // keep a valid non-null pointer so kernels can form addresses
            cudaError_t e = cudaMalloc(&p, sizeof(T) > 16 ? sizeof(T) : 16);
            return e;
        }
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess)
            return e;
        return cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

constexpr unsigned kMaxBands = 8;

struct Scratch {
    float* rgb = nullptr;
    int* ids = nullptr;
    size_t pixels = 0;
    unsigned* tileCounter = nullptr; // [0] tile counter
    Counters* counters = nullptr;
    float* gatherRgb = nullptr; // rank 0 only: receive staging for the other ranks' packed tiles
    int* gatherIds = nullptr;
    size_t gatherPixels = 0;
    // wavefront queues (wavefront.cuh), grown on demand and kept for the next frame
    WaveBuffers wave {};
    size_t waveRecFloats = 0, waveMeta = 0, waveNext = 0, waveDirFloats = 0, waveVis = 0, waveSub = 0, waveCont = 0, waveHard = 0;
    // split chain stage (wavefront.cuh wf_primary_kernel / wf_continue_kernel): the level-0 shadow pass runs on auxStream beside
    // the continuation kernel
    cudaStream_t auxStream = nullptr;
    cudaEvent_t evPrimary = nullptr, evVis0 = nullptr;
    float* bloomTmp = nullptr; // thresholded copy of the frame (renderBloomFilter's screenThreshold)
    size_t bloomPixels = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaEvent_t stage[5] = {}; // wavefront stage boundaries: chain | shadow rays | shading | fold |
    bool staged = false;       // stage[] recorded by the last launch
    unsigned* grant = nullptr;    // { tile_first, tile_count, part_index } of a dynamically dealt chunk (grant_kernel)
    cudaEvent_t evCull = nullptr; // behind the light-hull pre-pass, inside the shadow-ray stage (cge_stats::vis_cull_ms)
    bool cullTimed = false;
    // a frame rendered in concurrent bands (cge_render) uses one Scratch per band: kernels done / output copied
    cudaEvent_t bandDone = nullptr, copyDone = nullptr;
    // band b runs on a stream of priority (greatest - b), so that earlier bands win the SMs and finish (and leave for the
    // host) first while later bands fill the SMs they leave idle; created on first use, `stream` points at one of them or at
    // `baseStream` for the duration of a call
    cudaStream_t baseStream = nullptr, prioStream[kMaxBands] = {};
    bool busy = false;
};
} // namespace

// One immutable version of the scene's light list.  A render takes a reference (shared_ptr) under the scene lock when it starts
// and keeps it until its stream has drained, so cge_scene_update_lights never overwrites or frees a buffer a frame in flight
// reads: it fills a version nobody holds (or a new one) and swaps it in.
struct LightSet {
    std::vector<cge_light_desc> host;
    float* dev = nullptr;
    size_t capFloats = 0;
    uint32_t n = 0;
    ~LightSet()
    {
        if (dev)
            cudaFree(dev);
    }
};
using LightsRef = std::shared_ptr<const LightSet>;

struct cge_scene {
    int device = 0;
    int sm_count = 0;
    DevScene dev {};
    DevBuf<float4> nodes, tris, fnodes, f4nodes, onodes, ftris, shade, materials, sph_rows, sph_boxes;
    DevBuf<uint32_t> sph_box_off, fpos;
    DevBuf<uint4> qnodes;
    FastBvh fast;
    bool fast_built_on_gpu = false;
    float fast_build_ms = 0.0f; // device time of the GPU SAH build
    DevBuf<int4> textures;
    DevBuf<float> texels;
    std::shared_ptr<LightSet> lights;                   // the current version (guarded by mu)
    std::vector<std::shared_ptr<LightSet>> light_pool; // every version allocated so far; one with use_count() == 1 is free
    cudaStream_t light_stream = nullptr;
    HostBvh bvh;
    uint32_t n_triangles = 0, n_spheres = 0;
    uint32_t accel_root_ref = 0, accel_root_count = 0;
    bool any_transparent = false;
    bool colours_bounded = true; // every kd / ks / texel is finite and <= kColourBound in magnitude (zero-shading cull)
    std::mutex mu;
    std::vector<Scratch*> pool;
};

struct cge_comm {
    int rank = 0, n_ranks = 1, device = 0;
    void* nccl = nullptr; // ncclComm_t
    struct HostFrame {
        void* ptr;
        size_t bytes;
    };
    std::vector<HostFrame> host_frames; // cge_comm_host_frame mappings, released with the communicator
    std::vector<HostFrame> peer_frames; // cge_comm_peer_frame: rank 0's device frames (owned there, IPC mappings elsewhere); each is
                                        // followed by the tile counter of CGE_FLAG_DYNAMIC_TILES (at bytes rounded up to 256)
    std::vector<uint32_t> peer_grants;  // grants taken from that counter so far, per frame (the same number on every rank)
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
namespace {

inline float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
inline float bitsf(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

void host_light_counts(const std::vector<cge_light_desc>& lights, const cge_params& p, uint32_t& draws, uint32_t& shadows,
    uint32_t& samples)
{
    draws = 0;
    shadows = 0;
    samples = 0;
    if (!(p.features & CGE_FEAT_SHADING))
        return;
    const bool hard = p.features & CGE_FEAT_HARD_SHADOW, soft = p.features & CGE_FEAT_SOFT_SHADOW;
    for (const auto& l : lights) {
        if (l.type == CGE_LIGHT_POINT) {
            samples += 1;
            if (hard)
                shadows += 1;
        } else if (l.type == CGE_LIGHT_SEGMENT) {
            if (soft) {
                const uint32_t n = uint32_t(std::max(p.segment_samples, 0));
                draws += n;
                shadows += n;
                samples += n;
            }
        } else if (soft) {
            const uint32_t n = uint32_t(std::max(p.parallelogram_samples, 0));
            draws += 2 * n * n;
            shadows += n * n;
            samples += n * n;
        }
    }
}

std::vector<float> pack_lights(const cge_light_desc* lights, uint32_t n)
{
    std::vector<float> out(size_t(n) * kLightFloats, 0.0f);
    for (uint32_t i = 0; i < n; i++) {
        out[size_t(i) * kLightFloats] = bitsf(lights[i].type);
        std::memcpy(&out[size_t(i) * kLightFloats + 1], lights[i].v, sizeof(float) * 21);
    }
    return out;
}

int acquire_scratch(cge_scene* sc, size_t pixels, bool wantIds, size_t gatherPixels, Scratch** out)
{
    Scratch* s = nullptr;
    {
        std::lock_guard<std::mutex> lk(sc->mu);
        for (auto* c : sc->pool)
            if (!c->busy) {
                s = c;
                break;
            }
        if (!s) {
            s = new Scratch();
            sc->pool.push_back(s);
        }
        s->busy = true;
    }
    *out = s;
    if (!s->stream) {
        CGE_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->baseStream = s->stream;
        CGE_CUDA(cudaEventCreate(&s->ev0));
        CGE_CUDA(cudaEventCreate(&s->ev1));
        CGE_CUDA(cudaEventCreate(&s->ev2));
        for (auto& ev : s->stage)
            CGE_CUDA(cudaEventCreate(&ev));
        CGE_CUDA(cudaEventCreate(&s->evCull));
        CGE_CUDA(cudaEventCreateWithFlags(&s->bandDone, cudaEventDisableTiming));
        CGE_CUDA(cudaEventCreateWithFlags(&s->copyDone, cudaEventDisableTiming));
        CGE_CUDA(cudaMalloc(&s->tileCounter, 64));
        CGE_CUDA(cudaMalloc(&s->counters, sizeof(Counters)));
    }
    if (s->pixels < pixels || (wantIds && !s->ids)) {
        if (s->rgb)
            cudaFree(s->rgb);
        if (s->ids)
            cudaFree(s->ids);
        s->rgb = nullptr;
        s->ids = nullptr;
        s->pixels = 0;
        CGE_CUDA(cudaMalloc(&s->rgb, pixels * 3 * sizeof(float)));
        CGE_CUDA(cudaMalloc(&s->ids, pixels * sizeof(int)));
        s->pixels = pixels;
    }
    if (gatherPixels > s->gatherPixels) {
        if (s->gatherRgb)
            cudaFree(s->gatherRgb);
        if (s->gatherIds)
            cudaFree(s->gatherIds);
        s->gatherRgb = nullptr;
        s->gatherIds = nullptr;
        s->gatherPixels = 0;
        CGE_CUDA(cudaMalloc(&s->gatherRgb, gatherPixels * 3 * sizeof(float)));
        CGE_CUDA(cudaMalloc(&s->gatherIds, gatherPixels * sizeof(int)));
        s->gatherPixels = gatherPixels;
    }
    return CGE_OK;
}

void release_scratch(cge_scene* sc, Scratch* s)
{
    if (!s)
        return;
    if (s->baseStream)
        s->stream = s->baseStream;
    std::lock_guard<std::mutex> lk(sc->mu);
    s->busy = false;
}

int validate_params(const cge_scene* sc, const cge_params* p)
{
    if (!sc || !p)
        return fail(CGE_ERR_INVALID_ARG, "null scene or params");
    if (p->width <= 0 || p->height <= 0 || int64_t(p->width) * p->height > (int64_t(1) << 30))
        return fail(CGE_ERR_INVALID_ARG, "bad resolution");
    if (p->features & CGE_FEAT_EXTRA_MASK & ~CGE_FEAT_EXTRA_SUPPORTED)
        return fail(CGE_ERR_UNSUPPORTED,
            "of the ExtraFeatures (reference src/common.h:54-65) only enableBloomEffect and enableMultipleRaysPerPixel are implemented");
    if ((p->features & CGE_FEAT_MULTIPLE_RAYS_PER_PIXEL)
        && (p->rays_per_pixel_side < 1 || p->rays_per_pixel_side > kCgeMaxRaysPerPixelSide))
        return fail(CGE_ERR_INVALID_ARG, "rays_per_pixel_side must be in [1, 10] (the reference GUI range, src/main.cpp:195)");
    if (p->features & ~(0x7fu | CGE_FEAT_EXTRA_MASK))
        return fail(CGE_ERR_INVALID_ARG, "unknown feature bits");
    if (p->ray_depth < 0 || p->ray_depth > kMaxRayDepth)
        return fail(CGE_ERR_INVALID_ARG, "ray_depth must be in [0, 15]");
    if (p->segment_samples < 0 || p->parallelogram_samples < 0 || p->segment_samples > 4096 || p->parallelogram_samples > 256)
        return fail(CGE_ERR_INVALID_ARG, "bad sample counts");
    if (p->sampler != CGE_SAMPLER_HASH)
        return fail(CGE_ERR_UNSUPPORTED, "only CGE_SAMPLER_HASH is implemented");
    if (p->traversal > CGE_TRAVERSAL_FAST)
        return fail(CGE_ERR_INVALID_ARG, "bad traversal mode");
    if ((p->features & CGE_FEAT_RECURSIVE) && sc->any_transparent)
        return fail(CGE_ERR_UNSUPPORTED,
            "a material has transparency != 1: with enableRecursive the reference recurses without bound "
            "(src/render.cpp:122-130), there is no result to reproduce");
    return CGE_OK;
}

LightsRef lights_of(cge_scene* sc)
{
    std::lock_guard<std::mutex> lk(sc->mu);
    return sc->lights;
}

// Install a new light list: reuse a version no render holds, or allocate one; the upload has completed when this returns.
int set_lights(cge_scene* sc, const cge_light_desc* lights, uint32_t n)
{
    const std::vector<float> packed = pack_lights(lights, n);
    std::lock_guard<std::mutex> lk(sc->mu);
    std::shared_ptr<LightSet> ls;
    for (auto& c : sc->light_pool)
        if (c.use_count() == 1 && c->capFloats >= packed.size()) {
            ls = c;
            break;
        }
    if (!ls) {
        ls = std::make_shared<LightSet>();
        ls->capFloats = std::max<size_t>(packed.size(), 4 * kLightFloats);
        CGE_CUDA(cudaMalloc(&ls->dev, ls->capFloats * sizeof(float)));
        sc->light_pool.push_back(ls);
    }
    if (!sc->light_stream)
        CGE_CUDA(cudaStreamCreateWithFlags(&sc->light_stream, cudaStreamNonBlocking));
    if (!packed.empty()) {
        CGE_CUDA(cudaMemcpyAsync(ls->dev, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice, sc->light_stream));
        CGE_CUDA(cudaStreamSynchronize(sc->light_stream));
    }
    ls->host.assign(lights, lights + n);
    ls->n = n;
    sc->lights = ls;
    return CGE_OK;
}

// tiles part `rank` of `nRanks` renders: its units (dev_scene.h part_tile_of) are rank, rank + nRanks, ...; the frame's last unit
// may be short
unsigned tiles_of(const DevParams& d, unsigned rank, unsigned nRanks)
{
    const unsigned nTiles = d.n_tiles_x * d.n_tiles_y, unit = std::max(d.part_unit, 1u);
    const unsigned nUnits = (nTiles + unit - 1) / unit;
    if (nUnits <= rank)
        return 0;
    const unsigned mine = (nUnits - rank + nRanks - 1) / nRanks;
    const unsigned shortBy = (nUnits - 1) % nRanks == rank ? nUnits * unit - nTiles : 0;
    return mine * unit - shortBy;
}

DevParams make_dev_params(const cge_scene* sc, const cge_params& p, const LightSet& ls)
{
    DevParams d {};
    d.width = p.width;
    d.height = p.height;
    d.features = p.features;
    d.ray_depth = p.ray_depth;
    d.segment_samples = p.segment_samples;
    d.parallelogram_samples = p.parallelogram_samples;
    auto exact_recip = [](int n) { return n > 0 && (n & (n - 1)) == 0 ? 1.0f / float(n) : 0.0f; };
    d.segment_recip = exact_recip(p.segment_samples);
    d.parallelogram_recip = exact_recip(p.parallelogram_samples);
    d.seed = p.seed;
    host_light_counts(ls.host, p, d.draws_per_hit, d.shadow_rays_per_hit, d.samples_per_hit);
    d.levels = (p.features & CGE_FEAT_RECURSIVE) ? uint32_t(p.ray_depth) + 1u : 1u;
    d.units_per_lane = d.draws_per_hit == 0 ? d.levels : ((1u << d.levels) - 1u);
    d.debug_cycles = (p.flags & CGE_DEV_FLAG_DEBUG_CYCLES) ? 1u : 0u;
    d.part_index = p.part_count > 1 ? p.part_index : 0;
    d.part_count = p.part_count > 1 ? p.part_count : 1;
    d.n_tiles_x = uint32_t((p.width + kTileW - 1) / kTileW);
    d.n_tiles_y = uint32_t((p.height + kTileH - 1) / kTileH);
    d.aa_side = (p.features & CGE_FEAT_MULTIPLE_RAYS_PER_PIXEL) ? uint32_t(p.rays_per_pixel_side) : 0u;
        d.part_unit = (d.part_count > 1 && (p.flags & CGE_FLAG_PARTITION_TILE_ROWS)) ? d.n_tiles_x : 1u;
    d.tile_first = 0;
    d.tile_count = tiles_of(d, d.part_index, d.part_count);
    return d;
}

// Boxes of the fast tree on the 15-bit grid of dev_scene.h: plane = lo + m * ext with m = 1 + q / 32768.  The grid spans the
// scene bounds padded by 1/256 of their extent; lower planes are rounded down and upper planes up, then moved one more step
// outwards (the margin that absorbs the device-side rounding, trace.cuh).  Evaluated in double against the float lo / ext the
// device will use.
[[maybe_unused]] void quantise_fast_nodes(const FastBvh& fb, std::vector<uint4>& out, float lo[3], float ext[3])
{
    double smin[3] = { 1e300, 1e300, 1e300 }, smax[3] = { -1e300, -1e300, -1e300 };
    for (const FastNode& n : fb.nodes)
        for (int k = 0; k < 3; k++)
            for (float v : { n.l_lo[k], n.r_lo[k], n.l_hi[k], n.r_hi[k] })
                if (std::isfinite(v)) { // a box around non-finite vertices gets the whole grid below
                    smin[k] = std::min(smin[k], double(v));
                    smax[k] = std::max(smax[k], double(v));
                }
    for (int k = 0; k < 3; k++)
        if (smin[k] > smax[k])
            smin[k] = smax[k] = 0.0;
    for (int k = 0; k < 3; k++) {
        const double e = std::max(smax[k] - smin[k], 1e-6 * (std::fabs(smin[k]) + std::fabs(smax[k]) + 1.0));
        ext[k] = float(e * (1.0 + 2.0 / 256.0));
        lo[k] = float(smin[k] - e / 256.0 - double(ext[k])); // grid point m = 1 sits e / 256 below the scene minimum
    }
    auto grid = [&](int k, double v, bool upper) {
        if (!std::isfinite(v))
            return uint32_t(0x8000u | (upper ? 32767u : 0u));
        const double x = ((v - double(lo[k])) / double(ext[k]) - 1.0) * 32768.0;
        long q = upper ? long(std::ceil(x)) + 1 : long(std::floor(x)) - 1;
        q = std::min(std::max(q, 0l), 32767l);
        return uint32_t(0x8000u | uint32_t(q));
    };
    out.resize(fb.nodes.size() * 2);
    for (size_t i = 0; i < fb.nodes.size(); i++) {
        const FastNode& n = fb.nodes[i];
        uint32_t l[3], r[3];
        for (int k = 0; k < 3; k++) {
            l[k] = grid(k, n.l_lo[k], false) | (grid(k, n.l_hi[k], true) << 16);
            r[k] = grid(k, n.r_lo[k], false) | (grid(k, n.r_hi[k], true) << 16);
        }
        out[2 * i] = make_uint4(l[0], l[1], l[2], n.left);
        out[2 * i + 1] = make_uint4(r[0], r[1], r[2], n.right);
    }
}

// Magnitude below which a product of a material colour and a light colour cannot overflow: the zero-shading cull
// (shade.cuh shading_is_zero) relies on (kd * Lc) * 0 and (ks * Lc) * 0 being +-0, not NaN.
constexpr float kColourBound = 1e18f;
bool bounded(const float* v, size_t n)
{
    for (size_t i = 0; i < n; i++)
        if (!(std::fabs(v[i]) <= kColourBound))
            return false;
    return true;
}
bool light_colours_bounded(const std::vector<cge_light_desc>& lights)
{
    for (const auto& l : lights) {
        // colour members per type (include/cge.h cge_light_desc); interpolated sample colours stay within 2x the corner colours
        const int from = l.type == CGE_LIGHT_POINT ? 3 : l.type == CGE_LIGHT_SEGMENT ? 6 : 9;
        const int to = l.type == CGE_LIGHT_POINT ? 6 : l.type == CGE_LIGHT_SEGMENT ? 12 : 21;
        if (!bounded(l.v + from, size_t(to - from)))
            return false;
    }
    return true;
}

// development aid: integer tunables can be overridden from the environment for A/B sweeps (tools/sweep_vis.py)
int env_int(const char* name, int fallback)
{
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : fallback;
}

DevScene scene_for(const cge_scene* sc, const cge_params& p, const LightSet& ls)
{
    DevScene d = sc->dev;
    d.lights = ls.dev;
    d.n_lights = ls.n;
    d.cull_zero_shading = sc->colours_bounded && light_colours_bounded(ls.host) && env_int("CGE_ZERO_SHADING_CULL", 1);
    d.noaccel = (p.features & CGE_FEAT_ACCEL_STRUCTURE) ? 0u : 1u;
    if (p.features & CGE_FEAT_ACCEL_STRUCTURE) {
        d.root_ref = sc->accel_root_ref;
        d.root_count = sc->accel_root_count;
    } else {
        // !enableAccelStructure: one getIntersecting over the whole (permuted) primitive vector
        // (reference src/bounding_volume_hierarchy.cpp:303-305)
        d.root_ref = 0;
        d.root_count = d.n_prims;
    }
    return d;
}

// Output stage of Screen::writeBitmapToFile (reference src/screen.cpp:49-60): glm::clamp(color, 0, 1) -> vec4(c, 1) * 255 ->
// glm::u8vec4 (float -> u8 truncation).  glm::clamp = min(max(x, 0), 1) lets NaN through and the x86 conversion of NaN to an
// integer yields 0 in the low byte; reproduced as 0.
__global__ void pack_rgba8_kernel(const float* __restrict__ rgb, uchar4* __restrict__ out, size_t pixels)
{
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= pixels)
        return;
    auto conv = [](float v) -> unsigned char {
        if (v != v)
            return 0;
        v = fminf(fmaxf(v, 0.0f), 1.0f);
        return (unsigned char)__float2int_rz(__fmul_rn(v, 255.0f));
    };
    out[i] = make_uchar4(conv(rgb[i * 3]), conv(rgb[i * 3 + 1]), conv(rgb[i * 3 + 2]), 255);
}

// ---- extra.enableBloomEffect: renderBloomFilter (reference src/render.cpp:158-196) ---------------------------------
// weightsGaussian(sigma) (src/render.cpp:198-210) with the reference's types: the exponent is an int divided by a float,
// exp is the DOUBLE function (the reference object file imports `exp`, not `expf`), the normalisation 2 * 3.1415 * sigma^2
// is double, each weight is rounded to float, the float sum normalises.  Indexed [k + 1][j + 1] like the glm::mat3.
struct BloomWeights {
    float w[3][3];
};
BloomWeights weights_gaussian(float sigma)
{
    BloomWeights a {};
    float sum = 0.0f;
    for (int i = -1; i < 2; i++)
        for (int k = -1; k < 2; k++) {
            const float weight = float(std::exp(double(-(i * i + k * k) / (2 * sigma * sigma))) / (2 * 3.1415 * sigma * sigma));
            a.w[i + 1][k + 1] = weight;
            sum += weight;
        }
    for (auto& col : a.w)
        for (float& v : col)
            v = v / sum;
    return a;
}

// pass 1 (:165-170): pixels darker than the threshold become black.  brightness is evaluated in DOUBLE (the literals
// 0.2126 / 0.7152 / 0.0722 are doubles) and rounded to float before the float comparison; NaN brightness keeps the pixel.
__global__ void bloom_threshold_kernel(const float* __restrict__ rgb, float* __restrict__ thr, size_t pixels, float threshold)
{
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= pixels)
        return;
    const float r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
    const double br = __dadd_rn(__dadd_rn(__dmul_rn(0.2126, double(r)), __dmul_rn(0.7152, double(g))), __dmul_rn(0.0722, double(b)));
    const bool dark = __double2float_rn(br) < threshold;
    thr[i * 3] = dark ? 0.0f : r;
    thr[i * 3 + 1] = dark ? 0.0f : g;
    thr[i * 3 + 2] = dark ? 0.0f : b;
}

// pass 2 (:171-195): 3x3 Gaussian of the thresholded image added to every pixel EXCEPT the last column (x = W-1) and the
// bottom screen row (y = H-1 in reference coordinates), which the reference's loop bounds leave untouched.  k (x offset)
// outer, j (y offset) inner, products summed in that order.  In place: a pixel's own value is the only thing read from rgb.
__global__ void bloom_apply_kernel(float* __restrict__ rgb, const float* __restrict__ thr, int W, int H, BloomWeights wg, float scalar,
    int debugOption)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W - 1 || y >= H - 1)
        return;
    auto indexAt = [&](int xx, int yy) { return size_t(H - 1 - yy) * size_t(W) + size_t(xx); }; // src/screen.cpp:45
    vec3 sum = v3(0.0f);
    for (int k = -1; k < 2; k++)
        for (int j = -1; j < 2; j++) {
            if (x + k < 0 || x + k > W - 1 || y + j < 0 || y + j > H - 1)
                continue;
            const float* t = thr + indexAt(x + k, y + j) * 3;
            sum = sum + v3(t[0], t[1], t[2]) * wg.w[k + 1][j + 1];
        }
    float* px = rgb + indexAt(x, y) * 3;
    const vec3 bloom = sum * scalar;
    vec3 out = v3(px[0], px[1], px[2]);
    if (debugOption == 0)
        out = out + bloom;
    else if (debugOption == 1)
        out = bloom;
    px[0] = out.x, px[1] = out.y, px[2] = out.z;
}

// Which kernels a call runs.
//   fast     : CGE_TRAVERSAL_FAST over the SAH tree of the triangles, with the scene's spheres tested beside it (the archive's
//              sphere test assumes a unit direction, so with shadow rays its result depends on which boxes the REFERENCE tree
//              lets through: each sphere carries that box chain, trace.cuh sphere_reached); else the literal traversal of the
//              reference-order tree.
//   count    : box / triangle test counters (CGE_FLAG_COUNT_TESTS; per-thread kernel).
//   wave     : the wavefront pipeline (wavefront.cuh) instead of the per-thread kernel.
struct Variant {
    bool fast, spheres, count, wave;
    size_t waveBytes;
};

constexpr unsigned kMaxFastSpheres = 64;
constexpr size_t kWaveScratchLimit = size_t(24) << 30; // queues larger than this fall back to the per-thread kernel

struct WaveSizes {
    size_t recFloats, meta, next, dirFloats, vis;
    size_t bytes() const { return recFloats * 4 + meta * 8 + next * 4 + dirFloats * 4 + vis; } // (the sub-ray colours are small beside these)
};
// camera rays per pixel: with extra.enableMultipleRaysPerPixel every one of them is a chain of its own in the queues
size_t sub_rays(const DevParams& dp) { return dp.aa_side ? size_t(dp.aa_side) * dp.aa_side : 1; }

// cap = queue capacity per level = camera rays of the launch
WaveSizes wave_sizes(const DevParams& dp, size_t cap)
{
    WaveSizes w;
    w.vis = size_t(dp.units_per_lane) * std::max<size_t>(dp.samples_per_hit, 1) * cap; // one byte per light sample
    w.recFloats = size_t(dp.levels) * kWaveRecFloats * cap;
    w.meta = size_t(dp.levels) * cap;
    w.next = size_t(dp.levels) * cap;
    w.dirFloats = size_t(dp.units_per_lane) * 3 * cap;
    return w;
}

Variant choose_variant(const DevScene& ds, const cge_params& p, const DevParams& dp)
{
    Variant v {};
    v.spheres = ds.has_spheres != 0;
    // The fast tree serves every CGE_TRAVERSAL_FAST frame: spheres are tested beside it behind the reference tree's box chain,
    // and a frame without enableAccelStructure is the same minimum-t answer with the tie rank of the reference's primitive
    // vector (trace.cuh).  More spheres than a linear pass should carry fall back to the literal traversal.
    v.fast = p.traversal == CGE_TRAVERSAL_FAST && ds.n_sph <= kMaxFastSpheres;
    v.count = (p.flags & CGE_FLAG_COUNT_TESTS) != 0;
    const size_t cap = size_t(dp.tile_count) * 32 * sub_rays(dp);
    v.waveBytes = wave_sizes(dp, std::max<size_t>(cap, 1)).bytes();
    // The wavefront pays off when a pixel's cost is wildly non-uniform, i.e. with area lights (16+ shadow rays per evaluation,
    // 2^k evaluations at level k).  Point-light frames (<= 1 shadow ray per light and hit, recursion folded) are bounded per
    // pixel and launch-latency sensitive: there the single per-thread kernel is faster.
    const bool areaLights = dp.draws_per_hit > 0 || (p.flags & CGE_DEV_FLAG_WAVEFRONT);
    // several camera rays per pixel: the queues address a pixel with 24 bits (wavefront.cuh meta layout), larger frames take the
    // per-thread kernel
    const bool aaFits = !dp.aa_side || size_t(p.width) * size_t(p.height) <= (size_t(1) << 24);
    v.wave = v.fast && !v.count && areaLights && aaFits && (p.features & CGE_FEAT_SHADING)
        && !(p.flags & (CGE_DEV_FLAG_PER_THREAD | CGE_DEV_FLAG_DEBUG_CYCLES)) && v.waveBytes <= kWaveScratchLimit && cap > 0;
    return v;
}

template <typename F>
void dispatch(const Variant& v, F&& f)
{
    if (v.fast && v.count)
        f(std::true_type {}, std::false_type {}, std::true_type {});
    else if (v.fast)
        f(std::true_type {}, std::false_type {}, std::false_type {});
    else if (v.spheres && v.count)
        f(std::false_type {}, std::true_type {}, std::true_type {});
    else if (v.spheres)
        f(std::false_type {}, std::true_type {}, std::false_type {});
    else if (v.count)
        f(std::false_type {}, std::false_type {}, std::true_type {});
    else
        f(std::false_type {}, std::false_type {}, std::false_type {});
}

// packed-tile <-> screen layout.  A rank's k-th tile occupies pixels [32k, 32k+32) of its packed buffer.
__global__ void pack_tiles_kernel(const float* __restrict__ rgb, const int* __restrict__ ids, float* __restrict__ outRgb,
    int* __restrict__ outIds, DevParams p, unsigned myTiles)
{
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned k = g / 32, lane = g % 32;
    if (k >= myTiles)
        return;
    const unsigned tile = part_tile_of(p.part_unit, p.part_index, p.part_count, k);
    const int x = int(tile % p.n_tiles_x) * kTileW + int(lane % kTileW);
    const int y = int(tile / p.n_tiles_x) * kTileH + int(lane / kTileW);
    float r = 0.f, gg = 0.f, b = 0.f;
    int id = -1;
    if (x < p.width && y < p.height) {
        const size_t idx = size_t(p.height - 1 - y) * size_t(p.width) + size_t(x);
        r = rgb[idx * 3], gg = rgb[idx * 3 + 1], b = rgb[idx * 3 + 2];
        if (ids)
            id = ids[idx];
    }
    outRgb[size_t(g) * 3] = r;
    outRgb[size_t(g) * 3 + 1] = gg;
    outRgb[size_t(g) * 3 + 2] = b;
    if (outIds)
        outIds[g] = id;
}

// Rank 0 of cge_render_distributed: the other ranks' compact rows (dev_scene.h compact_units), received back to back in rank
// order, scattered into the frame in ONE launch.  Rank r's block b holds tile row u = r + (units[r] - 1 - b) * n_ranks.
struct UnpackRows {
    unsigned n_ranks;
    unsigned offset[16], units[16]; // first block and block count of rank r in the staging buffer (cge_comm_create: <= 16 ranks)
};
__global__ void unpack_rows_kernel(const float* __restrict__ inRgb, const int* __restrict__ inIds, float* __restrict__ rgb,
    int* __restrict__ ids, int W, int H, UnpackRows up, size_t total)
{
    const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= total)
        return;
    const size_t blockPixels = size_t(kTileH) * size_t(W);
    const unsigned blk = unsigned(g / blockPixels), within = unsigned(g - size_t(blk) * blockPixels);
    unsigned r = 1;
    while (r + 1 < up.n_ranks && blk >= up.offset[r + 1])
        r++;
    const unsigned b = blk - up.offset[r], u = r + (up.units[r] - 1u - b) * up.n_ranks;
    const int rows = min(int(u + 1) * kTileH, H) - int(u) * kTileH, top = H - min(int(u + 1) * kTileH, H);
    const int row = int(within / unsigned(W)), x = int(within % unsigned(W));
    if (row >= rows)
        return; // the frame's topmost tile row can be short
    const size_t idx = size_t(top + row) * size_t(W) + size_t(x);
    rgb[idx * 3] = inRgb[g * 3], rgb[idx * 3 + 1] = inRgb[g * 3 + 1], rgb[idx * 3 + 2] = inRgb[g * 3 + 2];
    if (ids && inIds)
        ids[idx] = inIds[g];
}

int launch_render(cge_scene* sc, const LightSet& ls, Scratch* s, const cge_camera* cam, const cge_params* p, const DevParams& dp,
    float* rgbDev, int* idsDev, uint32_t* launches)
{
    const DevScene ds = scene_for(sc, *p, ls);
    DevCamera dc { cam->origin[0], cam->origin[1], cam->origin[2], cam->quat[0], cam->quat[1], cam->quat[2], cam->quat[3],
        cam->half_width, cam->half_height };
    CGE_CUDA(cudaMemsetAsync(s->tileCounter, 0, 64, s->stream));
    CGE_CUDA(cudaMemsetAsync(s->counters, 0, sizeof(Counters), s->stream));
    s->staged = false;
    s->cullTimed = false;
    const Variant v = choose_variant(ds, *p, dp);
    cudaEvent_t* stage = s->stage;
    const unsigned myTiles = dp.tile_count;
    cudaError_t err = cudaSuccess;
    auto grid_for = [&](int perSm) {
        // persistent CTAs: a multiple of the SM count, never more CTAs than there are 4-tile batches
        unsigned grid = unsigned(sc->sm_count * std::max(perSm, 1));
        return std::max(1u, std::min(grid, (myTiles + 3) / 4));
    };
    if (v.wave) {
        const size_t cap = size_t(myTiles) * 32 * sub_rays(dp);
        const WaveSizes ws = wave_sizes(dp, cap);
        auto grow = [&](auto*& ptr, size_t& have, size_t need, size_t elem) -> cudaError_t {
            if (have >= need)
                return cudaSuccess;
            if (ptr)
                cudaFree(ptr);
            ptr = nullptr;
            have = 0;
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), need * elem);
            if (e == cudaSuccess)
                have = need;
            return e;
        };
        err = grow(s->wave.rec, s->waveRecFloats, ws.recFloats, sizeof(float));
        if (err == cudaSuccess)
            err = grow(s->wave.meta, s->waveMeta, ws.meta, sizeof(uint2));
        if (err == cudaSuccess)
            err = grow(s->wave.next, s->waveNext, ws.next, sizeof(unsigned));
        if (err == cudaSuccess)
            err = grow(s->wave.dir, s->waveDirFloats, ws.dirFloats, sizeof(float));
        if (err == cudaSuccess && dp.aa_side)
            err = grow(s->wave.sub, s->waveSub, cap * 3, sizeof(float));
        // Shadow rays: traced by wf_vis_regroup_kernel into one visibility byte per ray and shaded from the bytes by
        // wf_shade_kernel<true>; launches with fewer than 8 samples per evaluation (or visibility bytes beyond 4 GB) let
        // wf_shade_kernel<false> trace them itself.
        const bool visFits = dp.samples_per_hit >= 8 && ws.vis < (size_t(1) << 32);
        DevParams wp = dp;
        wp.packet_fat = float(env_int("CGE_PACKET_FAT_PCT", 100)) * 0.01f;
        wp.packet_budget = uint32_t(env_int("CGE_PACKET_BUDGET", 48));
        wp.packet_leaf_cost = uint32_t(env_int("CGE_PACKET_LEAF_COST", 4));
        wp.shade_mode = visFits ? 2 : 1;
        // The light-hull pre-pass (wavefront.cuh wf_vis_cull_kernel): hits none of whose shadow rays can be blocked are settled by one
        // conservative walk; the shadow-ray kernel traces the rest.  Scenes with spheres keep every hit (a sphere is not judged by
        // plane tests).  CGE_VIS_CULL=0 switches it off (A/B, tests: the frame is the same bit for bit).
        // Pays for its extra launch on large launches only (measured on B200, DESIGN.md 5.10: C5 frame 14.20 -> 13.64 ms, but a
        // 1/8 share 2.38 -> 2.47 ms and C3 1.04 -> 1.08 ms): on from 1.5 Mpixel per launch, CGE_VIS_CULL=0 / 1 forces it.
        const bool smallLaunch = size_t(myTiles) * 32 < (size_t(3) << 19);
        wp.vis_cull = visFits && ds.n_sph == 0 && ds.n_ftris > 0 && env_int("CGE_VIS_CULL", smallLaunch ? 0 : 1) ? 1u : 0u;
        wp.cull_budget = uint32_t(env_int("CGE_CULL_BUDGET", 96));
        if (err == cudaSuccess && visFits)
            err = grow(s->wave.vis, s->waveVis, ws.vis, 1);
        if (err == cudaSuccess && wp.vis_cull)
            err = grow(s->wave.hard, s->waveHard, ws.meta, sizeof(unsigned));
        if (err == cudaSuccess && !s->wave.counts)
            err = cudaMalloc(reinterpret_cast<void**>(&s->wave.counts), 64 * sizeof(unsigned));
        if (err == cudaSuccess)
            err = cudaMemsetAsync(s->wave.counts, 0, 64 * sizeof(unsigned), s->stream);
        s->wave.cap = unsigned(cap);
        int perSm = 0;
        // The chain stage in two kernels (wavefront.cuh): camera rays, then the rest of the chains BESIDE the level-0 shadow pass,
        // which hides the chain tail.  Frames with recursion and one camera ray per pixel; CGE_CHAIN_SPLIT=0 keeps the one kernel.
        // Measured on B200: a 1/8 share of C5 (one rank of eight) 2.59 -> 2.52 ms; the whole frame as one pipeline
        // 15.57 -> 15.67 ms and as four concurrent bands 15.33 -> 15.91 ms (bands already fill the tails, and the two shadow
        // launches each end in a tail of their own): on for launches below the band threshold only.
        // (not when the pixels go to rank 0's frame over NVLink: on 8 GPUs the unsplit steps of the same run took 2.67 ms against
        //  2.92 ms, on 2 GPUs at 1 Mpixel per rank 2.61 against 2.63 ms - DESIGN.md 6)
        const bool split = visFits && !dp.aa_side && wp.levels > 1 && env_int("CGE_CHAIN_SPLIT", smallLaunch && !dp.chain_unsplit ? 1 : 0);
        auto launchVis = [&](cudaStream_t st, unsigned levelBegin, unsigned levelEnd, unsigned counterIdx) {
            if (wp.vis_cull) { // (chunk counters 22 / 23 beside the shadow-ray kernel's 18 / 19)
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, wf_vis_cull_kernel, 128, 0);
                if (err != cudaSuccess)
                    return;
                wf_vis_cull_kernel<<<unsigned(sc->sm_count * std::max(perSm, 1)), 128, 0, st>>>(ds, wp, s->wave, levelBegin, levelEnd, counterIdx + 4);
                err = cudaGetLastError();
                *launches += 1;
                if (err != cudaSuccess)
                    return;
                if (st == s->stream && levelBegin == 0 && levelEnd == wp.levels) { // (the whole stage on the main stream: timed)
                    cudaEventRecord(s->evCull, st);
                    s->cullTimed = true;
                }
            }
            auto go = [&](auto kern) {
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 128, 0);
                if (err == cudaSuccess) {
                    kern<<<unsigned(sc->sm_count * std::max(perSm, 1)), 128, 0, st>>>(ds, wp, s->wave, s->counters, levelBegin, levelEnd, counterIdx);
                    err = cudaGetLastError();
                    *launches += 1;
                }
            };
            // 4 or 8 samples per lane, decided on the device from the queue lengths (they only exist there): both
            // instantiations are launched and the one not selected returns at once
            go(wf_vis_regroup_kernel<4>);
            if (err == cudaSuccess)
                go(wf_vis_regroup_kernel<8>);
        };
        cudaEventRecord(stage[0], s->stream);
        if (err == cudaSuccess && split) {
            if (!s->auxStream) {
                int least = 0, greatest = 0;
                cudaDeviceGetStreamPriorityRange(&least, &greatest);
                err = cudaStreamCreateWithPriority(&s->auxStream, cudaStreamNonBlocking, greatest);
                if (err == cudaSuccess)
                    err = cudaEventCreateWithFlags(&s->evPrimary, cudaEventDisableTiming);
                if (err == cudaSuccess)
                    err = cudaEventCreateWithFlags(&s->evVis0, cudaEventDisableTiming);
            }
            if (err == cudaSuccess)
                err = grow(s->wave.cont, s->waveCont, cap, sizeof(unsigned));
            if (err == cudaSuccess)
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, wf_primary_kernel, 128, 0);
            if (err == cudaSuccess) {
                wf_primary_kernel<<<grid_for(perSm), 128, 0, s->stream>>>(ds, dc, wp, s->wave, rgbDev, idsDev, s->counters);
                err = cudaGetLastError();
                cudaEventRecord(s->evPrimary, s->stream);
            }
            if (err == cudaSuccess) { // level 0 is final: its shadow rays start on the second stream
                cudaStreamWaitEvent(s->auxStream, s->evPrimary, 0);
                launchVis(s->auxStream, 0, 1, 18);
                cudaEventRecord(s->evVis0, s->auxStream);
            }
            if (err == cudaSuccess)
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, wf_continue_kernel, 128, 0);
            if (err == cudaSuccess) {
                wf_continue_kernel<<<unsigned(sc->sm_count * std::max(perSm, 1)), 128, 0, s->stream>>>(ds, wp, s->wave, s->counters);
                err = cudaGetLastError();
                *launches += 1;
            }
            cudaEventRecord(stage[1], s->stream);
            if (err == cudaSuccess)
                launchVis(s->stream, 1, wp.levels, 19);
            cudaStreamWaitEvent(s->stream, s->evVis0, 0);
        } else {
            if (err == cudaSuccess)
                err = dp.aa_side ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, wf_chain_kernel<true>, 128, 0)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, wf_chain_kernel<false>, 128, 0);
            if (err == cudaSuccess) {
                if (dp.aa_side)
                    wf_chain_kernel<true><<<grid_for(perSm), 128, 0, s->stream>>>(ds, dc, wp, s->wave, rgbDev, idsDev, s->counters);
                else
                    wf_chain_kernel<false><<<grid_for(perSm), 128, 0, s->stream>>>(ds, dc, wp, s->wave, rgbDev, idsDev, s->counters);
                err = cudaGetLastError();
            }
            cudaEventRecord(stage[1], s->stream);
            if (err == cudaSuccess && visFits) {
#if CGE_EXPERIMENTS
                const int packet = env_int("CGE_PACKET", 0); // the packet (hull) shadow walk: measured slower
                auto goPacket = [&](auto kern) {
                    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 128, 0);
                    if (err == cudaSuccess) {
                        kern<<<unsigned(sc->sm_count * std::max(perSm, 1)), 128, 0, s->stream>>>(ds, wp, s->wave, s->counters);
                        err = cudaGetLastError();
                        *launches += 1;
                    }
                };
                if (packet >= 16)
                    goPacket(wf_vis_packet_kernel<16>);
                else if (packet >= 8)
                    goPacket(wf_vis_packet_kernel<8>);
                else if (packet > 0)
                    goPacket(wf_vis_packet_kernel<4>);
                else
#endif
                    launchVis(s->stream, 0, wp.levels, 18);
            }
        }
        cudaEventRecord(stage[2], s->stream);
        if (err == cudaSuccess) {
            auto kern = visFits ? wf_shade_kernel<true> : wf_shade_kernel<false>;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 128, 0);
            if (err == cudaSuccess) {
                kern<<<unsigned(sc->sm_count * std::max(perSm, 1)), 128, 0, s->stream>>>(ds, wp, s->wave, s->counters);
                err = cudaGetLastError();
                *launches += 1;
            }
        }
        cudaEventRecord(stage[3], s->stream);
        if (err == cudaSuccess) {
            wf_fold_kernel<<<unsigned((cap + 127) / 128), 128, 0, s->stream>>>(wp, s->wave, rgbDev);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess && dp.aa_side) {
            wf_resolve_kernel<<<unsigned((size_t(myTiles) * 32 + 127) / 128), 128, 0, s->stream>>>(wp, s->wave, rgbDev);
            err = cudaGetLastError();
            *launches += 1;
        }
        cudaEventRecord(stage[4], s->stream);
        s->staged = true;
        *launches += 1; // chain + fold (the caller adds one)
    } else {
        dispatch(v, [&](auto kFast, auto kSpheres, auto kCount) {
            auto kern = render_kernel<decltype(kFast)::value, decltype(kSpheres)::value, decltype(kCount)::value>;
            int perSm = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 128, 0);
            if (err != cudaSuccess)
                return;
            kern<<<grid_for(perSm), 128, 0, s->stream>>>(ds, dc, dp, rgbDev, idsDev, s->tileCounter, s->counters);
            err = cudaGetLastError();
        });
    }
    if (err != cudaSuccess)
        return fail(CGE_ERR_CUDA, std::string("render_kernel launch: ") + cudaGetErrorString(err));
    *launches += 1;
    return CGE_OK;
}

// extra.enableBloomEffect on a complete frame resident on this GPU (src/render.cpp:326-328)
int apply_bloom(Scratch* s, const cge_params& p, float* frame, uint32_t* launches)
{
    const size_t pixels = size_t(p.width) * size_t(p.height);
    if (s->bloomPixels < pixels) {
        if (s->bloomTmp)
            cudaFree(s->bloomTmp);
        s->bloomTmp = nullptr;
        s->bloomPixels = 0;
        CGE_CUDA(cudaMalloc(&s->bloomTmp, pixels * 3 * sizeof(float)));
        s->bloomPixels = pixels;
    }
    bloom_threshold_kernel<<<unsigned((pixels + 255) / 256), 256, 0, s->stream>>>(frame, s->bloomTmp, pixels, p.bloom_threshold);
    const dim3 block(32, 8), grid(unsigned((p.width + 31) / 32), unsigned((p.height + 7) / 8));
    bloom_apply_kernel<<<grid, block, 0, s->stream>>>(frame, s->bloomTmp, p.width, p.height, weights_gaussian(1.0f), p.bloom_scalar,
        p.bloom_debug_option);
    CGE_CUDA(cudaGetLastError());
    *launches += 2;
    return CGE_OK;
}

// Number of concurrent bands of a launch (a frame in cge_render, a rank's partition in cge_render_distributed).  Measured on
// B200 (tools/sweep_bands.py): the wavefront pipeline gains from ~1 Mpixel up (its stages have long tails), the
// single per-thread kernel only through the overlapped copy of a host-output frame; small launches are launch-latency bound
// and stay in one piece.
unsigned nBandsFor(const cge_scene* sc, const LightSet& ls, const cge_params& p, const DevParams& dp, bool hostCopy)
{
    const size_t pixels = size_t(dp.tile_count) * 32;
    unsigned n = 1;
    const bool wave = choose_variant(scene_for(sc, p, ls), p, dp).wave;
    if (wave)
        // bands of at least 1.5 Mpixel, the size from which the light-hull pre-pass of the shadow stage pays (launch_render): a
        // 4 Mpixel share (C5 on 2 GPUs) measured 7.58 ms as 4 bands without it, 7.15 ms as 2 bands with it; a 1 Mpixel share (8 GPUs)
        // is faster in one piece
        n = pixels >= (size_t(6) << 20) ? 4 : pixels >= (size_t(3) << 20) ? 2 : 1;
    else if (hostCopy)
        n = pixels >= (size_t(2) << 20) ? 4 : 1;
    n = unsigned(std::max(env_int("CGE_BANDS", int(n)), 1));
    return std::min<unsigned>({ n, kMaxBands, std::max(dp.n_tiles_y, 1u), std::max(dp.tile_count, 1u) });
}

int use_band_stream(Scratch* s, unsigned band)
{
    int least = 0, greatest = 0;
    CGE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest)); // numerically lower = higher priority
    const int prio = std::min(greatest + int(band), least);
    if (!s->prioStream[band])
        CGE_CUDA(cudaStreamCreateWithPriority(&s->prioStream[band], cudaStreamNonBlocking, prio));
    s->stream = s->prioStream[band];
    return CGE_OK;
}

// One Scratch (stream, queues, counters) per band; bands[0] is the caller's.  Returns with rc != OK after releasing the helpers.
int acquire_bands(cge_scene* sc, Scratch* primary, unsigned nBands, std::vector<Scratch*>& bands)
{
    bands.assign(1, primary);
    for (unsigned b = 1; b < nBands; b++) {
        Scratch* h = nullptr;
        const int rc = acquire_scratch(sc, 1, false, 0, &h);
        if (h)
            bands.push_back(h);
        if (rc != CGE_OK) {
            for (unsigned k = 1; k < bands.size(); k++)
                release_scratch(sc, bands[k]);
            bands.resize(1);
            return rc;
        }
    }
    return CGE_OK;
}

// Render entries [first[b], first[b] + count[b]) of the launch's tile list as concurrent pipelines, band b on bands[b]'s
// stream.  The persistent kernels of one band drain while another band's kernels fill the freed SMs, which hides the tail
// every stage otherwise leaves (the slowest 8x4 tile of the chain kernel alone runs 0.8 ms on config C5).  staggered: band b
// gets stream priority (greatest - b), so that the bands finish one after the other and `after(b, scratch)` - which must
// record scratch->bandDone when the band's kernels are queued - can start the band's copy while later bands are traced;
// otherwise all bands share the top priority, which hides the tails best.  On return the primary stream has waited for the
// kernels of every band; ev0 must already be recorded on it.
template <typename After>
int launch_bands(cge_scene* sc, const LightSet& ls, const std::vector<Scratch*>& bands, const std::vector<uint2>& ranges, bool staggered, const cge_camera* cam,
    const cge_params* p, const DevParams& dp, float* rgbDev, int* idsDev, uint32_t* launches, After&& after)
{
    Scratch* s = bands[0];
    for (unsigned b = 0; b < bands.size(); b++) {
        Scratch* sb = bands[b];
        if (b) {
            const int rc = use_band_stream(sb, staggered ? b : 0);
            if (rc != CGE_OK)
                return rc;
            cudaStreamWaitEvent(sb->stream, s->ev0, 0);
        }
        DevParams bp = dp;
        bp.tile_first = ranges[b].x;
        bp.tile_count = ranges[b].y;
        const int rc = launch_render(sc, ls, sb, cam, p, bp, rgbDev, idsDev, launches);
        if (rc != CGE_OK)
            return rc;
        after(b, sb);
    }
    for (unsigned b = 1; b < bands.size(); b++)
        cudaStreamWaitEvent(s->stream, bands[b]->bandDone, 0);
    return CGE_OK;
}

// bands: the Scratch of every band of the frame, the primary one (which holds ev0..ev2) first
int fill_stats(const std::vector<Scratch*>& bands, cge_stats* st, uint32_t launches)
{
    if (!st)
        return CGE_OK;
    Scratch* s = bands[0];
    std::memset(st, 0, sizeof(*st));
    float ms = 0.f;
    for (Scratch* b : bands) {
        Counters c {};
        CGE_CUDA(cudaMemcpyAsync(&c, b->counters, sizeof(c), cudaMemcpyDeviceToHost, b->stream));
        CGE_CUDA(cudaStreamSynchronize(b->stream));
        st->primary_rays += c.primary;
        st->bounce_rays += c.bounce;
        st->shadow_rays += c.shadow;
        st->reference_rays += c.reference;
        st->box_tests += c.box;
        st->tri_tests += c.tri;
        st->reference_shadow_rays += c.reference_shadow;
        if (b->staged && b->wave.counts) { // the light-hull pre-pass counts in the pipeline's own counter block (wavefront.cuh)
            unsigned long long culled = 0;
            CGE_CUDA(cudaMemcpyAsync(&culled, b->wave.counts + 48, sizeof(culled), cudaMemcpyDeviceToHost, b->stream));
            CGE_CUDA(cudaStreamSynchronize(b->stream));
            st->shadow_samples_culled += culled;
        }
        if (b->staged && b->cullTimed) {
            CGE_CUDA(cudaEventElapsedTime(&ms, b->stage[1], b->evCull));
            st->vis_cull_ms += ms;
        }
        if (b->staged) // summed over the bands (which overlap in time when there are several)
            for (int k = 0; k < 4; k++) {
                CGE_CUDA(cudaEventElapsedTime(&ms, b->stage[k], b->stage[k + 1]));
                st->stage_ms[k] += ms;
            }
    }
    CGE_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    st->kernel_ms = ms;
    CGE_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev2));
    st->total_ms = ms;
    st->kernel_launches = launches;
    return CGE_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
// entry points
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int cge_abi_version(void) { return CGE_ABI_VERSION; }
const char* cge_last_error(void) { return g_err.c_str(); }

int cge_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cge_camera_from_trackball(float fovy, float aspect, const float look_at[3], float dist, const float rot[3], cge_camera* out)
{
    if (!look_at || !rot || !out)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    // Trackball ctor (reference framework/src/trackball.cpp:23-31)
    out->half_height = std::tan(fovy / 2.0f);
    out->half_width = aspect * out->half_height;
    // glm::quat(eulerAngles) (glm/detail/type_quat.inl:208-217)
    const float cx = std::cos(rot[0] * 0.5f), cy = std::cos(rot[1] * 0.5f), cz = std::cos(rot[2] * 0.5f);
    const float sx = std::sin(rot[0] * 0.5f), sy = std::sin(rot[1] * 0.5f), sz = std::sin(rot[2] * 0.5f);
    const float w = cx * cy * cz + sx * sy * sz;
    const float x = sx * cy * cz - cx * sy * sz;
    const float y = cx * sy * cz + sx * cy * sz;
    const float z = cx * cy * sz - sx * sy * cz;
    out->quat[0] = w, out->quat[1] = x, out->quat[2] = y, out->quat[3] = z;
    // position() = lookAt + q * vec3(0, 0, -dist) (trackball.cpp:71-74)
    const vec3 p = quat_rotate(w, v3(x, y, z), v3(0.0f, 0.0f, -dist));
    out->origin[0] = look_at[0] + p.x;
    out->origin[1] = look_at[1] + p.y;
    out->origin[2] = look_at[2] + p.z;
    return CGE_OK;
}

int cge_scene_create(const cge_scene_desc* d, int device, cge_scene** out)
{
    if (!d || !out)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if ((d->n_meshes && !d->meshes) || (d->n_vertices && !d->vertices) || (d->n_triangles && !d->triangles)
        || (d->n_spheres && !d->spheres) || (d->n_lights && !d->lights) || (d->n_textures && !d->textures)
        || (d->n_texels && !d->texels))
        return fail(CGE_ERR_INVALID_ARG, "count without array");
    // ---- validate indices -------------------------------------------------------------------------------
    uint64_t triSum = 0;
    for (uint32_t m = 0; m < d->n_meshes; m++) {
        const auto& md = d->meshes[m];
        if (uint64_t(md.vertex_offset) + md.vertex_count > d->n_vertices || uint64_t(md.triangle_offset) + md.triangle_count > d->n_triangles)
            return fail(CGE_ERR_INVALID_ARG, "mesh range out of bounds");
        if (md.triangle_offset != triSum)
            return fail(CGE_ERR_INVALID_ARG, "meshes must list their triangles contiguously in mesh order");
        triSum += md.triangle_count;
        if (md.texture_id >= int32_t(d->n_textures))
            return fail(CGE_ERR_INVALID_ARG, "texture id out of range");
        for (uint32_t t = 0; t < md.triangle_count; t++)
            for (int k = 0; k < 3; k++)
                if (d->triangles[3 * size_t(md.triangle_offset + t) + k] >= md.vertex_count)
                    return fail(CGE_ERR_INVALID_ARG, "vertex index out of range");
    }
    if (triSum != d->n_triangles)
        return fail(CGE_ERR_INVALID_ARG, "triangles not covered by meshes");
    for (uint32_t t = 0; t < d->n_textures; t++) {
        const auto& td = d->textures[t];
        if (td.width <= 0 || td.height <= 0 || td.texel_offset + uint64_t(td.width) * uint64_t(td.height) > d->n_texels
            || td.texel_offset > 0x7fffffffull)
            return fail(CGE_ERR_INVALID_ARG, "texture out of bounds");
    }
    for (uint32_t l = 0; l < d->n_lights; l++)
        if (d->lights[l].type > CGE_LIGHT_PARALLELOGRAM)
            return fail(CGE_ERR_INVALID_ARG, "bad light type");

    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
        cudaGetLastError();
        return fail(CGE_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= nDev)
        return fail(CGE_ERR_INVALID_ARG, "bad device ordinal");
    CGE_CUDA(cudaSetDevice(device));

    auto* sc = new cge_scene();
    sc->device = device;
    cudaDeviceGetAttribute(&sc->sm_count, cudaDevAttrMultiProcessorCount, device);
    sc->n_triangles = d->n_triangles;
    sc->n_spheres = d->n_spheres;
    const uint32_t nPrims = d->n_triangles + d->n_spheres;

    // ---- BVH (reference order) ----------------------------------------------------------------------------
    bool haveBvh = false;
    if (nPrims) {
        haveBvh = d->bvh_nodes ? adopt_bvh(*d, sc->bvh) : build_reference_bvh(*d, sc->bvh);
        if (!haveBvh) {
            delete sc;
            return fail(CGE_ERR_INVALID_ARG, "supplied BVH is inconsistent with the scene");
        }
        if (sc->bvh.n_levels > uint32_t(kRefStackSize - 4)) {
            delete sc;
            return fail(CGE_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
        }
    }

    // ---- global primitive tables ------------------------------------------------------------------------------
    std::vector<uint32_t> triMesh(d->n_triangles);
    for (uint32_t m = 0; m < d->n_meshes; m++)
        for (uint32_t t = 0; t < d->meshes[m].triangle_count; t++)
            triMesh[d->meshes[m].triangle_offset + t] = m;

    // visit rank = position in the reference's exhaustive right-first DFS (src/bounding_volume_hierarchy.cpp:312-361)
    std::vector<uint32_t> rank(nPrims, 0);
    if (haveBvh) {
        std::vector<uint32_t> st { sc->bvh.root };
        uint32_t r = 0;
        while (!st.empty()) {
            const uint32_t ni = st.back();
            st.pop_back();
            const auto& n = sc->bvh.nodes[ni];
            if (n.is_leaf) {
                for (uint32_t i = n.beg; i < n.end; i++)
                    rank[i] = r++;
            } else {
                st.push_back(n.left);
                st.push_back(n.right);
            }
        }
    }

    // rank is indexed by reference leaf position; re-index by global primitive id
    std::vector<uint32_t> rankOf(nPrims, 0);
    for (uint32_t i = 0; i < nPrims; i++)
        rankOf[sc->bvh.prim_order[i]] = rank[i];

    // per-primitive records by GLOBAL id (the ray-independent part of libIntersect's triangle test, precomputed on the
    // host with the same operation order and no FMA), then laid out in each tree's leaf order
    std::vector<float4> recs(size_t(nPrims) * kTriRows), shade(size_t(nPrims) * kShadeRows);
    const float qnan = std::nanf("");
    for (uint32_t gid = 0; gid < nPrims; gid++) {
        float4* tr = &recs[size_t(gid) * kTriRows];
        float4* sh = &shade[size_t(gid) * kShadeRows];
        if (gid < d->n_triangles) {
            const uint32_t m = triMesh[gid];
            const auto& md = d->meshes[m];
            const uint32_t* idx = d->triangles + 3 * size_t(gid);
            const cge_vertex& a = d->vertices[md.vertex_offset + idx[0]];
            const cge_vertex& b = d->vertices[md.vertex_offset + idx[1]];
            const cge_vertex& c = d->vertices[md.vertex_offset + idx[2]];
            const vec3 v0 = v3(a.position[0], a.position[1], a.position[2]);
            const vec3 v1 = v3(b.position[0], b.position[1], b.position[2]);
            const vec3 v2 = v3(c.position[0], c.position[1], c.position[2]);
            const TriPre t = triangle_precompute(v0, v1, v2);
            tr[0] = f4(t.n.x, t.n.y, t.n.z, t.D);
            tr[1] = f4(v0.x, v0.y, v0.z, t.e0.x);
            tr[2] = f4(t.e0.y, t.e0.z, v1.x, v1.y);
            tr[3] = f4(v1.z, t.e1.x, t.e1.y, t.e1.z);
            tr[4] = f4(v2.x, v2.y, v2.z, t.e2.x);
            tr[5] = f4(t.e2.y, t.e2.z, bitsf(rankOf[gid]), bitsf(gid));
            sh[0] = f4(a.normal[0], a.normal[1], a.normal[2], a.texcoord[0]);
            sh[1] = f4(b.normal[0], b.normal[1], b.normal[2], a.texcoord[1]);
            sh[2] = f4(c.normal[0], c.normal[1], c.normal[2], b.texcoord[0]);
            sh[3] = f4(b.texcoord[1], c.texcoord[0], c.texcoord[1], 0.0f);
            sh[4] = f4(bitsf(m), 0.0f, 0.0f, 0.0f);
        } else {
            const auto& sd = d->spheres[gid - d->n_triangles];
            tr[0] = f4(qnan, qnan, qnan, qnan);
            tr[1] = f4(sd.center[0], sd.center[1], sd.center[2], sd.radius);
            tr[2] = tr[3] = tr[4] = f4(qnan, qnan, qnan, qnan);
            tr[5] = f4(qnan, qnan, bitsf(rankOf[gid]), bitsf(gid | kSphereBit));
            sh[0] = sh[1] = sh[2] = sh[3] = f4(0, 0, 0, 0);
            sh[4] = f4(bitsf(d->n_meshes + (gid - d->n_triangles)), 0.0f, 0.0f, 0.0f);
        }
    }
    auto in_order = [&](const std::vector<uint32_t>& order) {
        std::vector<float4> out(order.size() * kTriRows);
        for (size_t i = 0; i < order.size(); i++)
            std::memcpy(&out[i * kTriRows], &recs[size_t(order[i]) * kTriRows], sizeof(float4) * kTriRows);
        return out;
    };
    std::vector<float4> tris = haveBvh ? in_order(sc->bvh.prim_order) : std::vector<float4>();

    // ---- fast tree (binned SAH, <= 4 primitives per leaf) for CGE_TRAVERSAL_FAST: over the TRIANGLES; spheres are kept beside
    //      it (below) -------------------------------------------------------------------------------------------------------
    std::vector<float4> ftris, fnodes;
    std::vector<uint32_t> fpos;
    cge_scene_desc triDesc = *d;
    triDesc.n_spheres = 0;
    triDesc.spheres = nullptr;
    triDesc.bvh_nodes = nullptr;
    triDesc.bvh_prim_order = nullptr;
    triDesc.n_bvh_nodes = 0;
    if (haveBvh && d->n_triangles > 0) {
        // built on the GPU (bvh_sah_gpu.cu); the host builder produces the identical tree and serves scenes with non-finite
        // coordinates.  CGE_SAH_BUILD=host forces it (A/B timing, tests).
        const char* forced = std::getenv("CGE_SAH_BUILD");
        const bool onGpu = sah_gpu_supported(triDesc) && !(forced && std::string(forced) == "host");
        std::string buildErr;
        const bool ok = onGpu ? build_sah_bvh_gpu(triDesc, sc->fast, &sc->fast_build_ms, &buildErr) : build_sah_bvh(triDesc, sc->fast);
        sc->fast_built_on_gpu = onGpu;
        if (!ok || sc->fast.depth > uint32_t(kFastStackSize - 2)) {
            delete sc;
            return fail(ok ? CGE_ERR_UNSUPPORTED : CGE_ERR_CUDA, buildErr.empty() ? "fast BVH build failed" : buildErr);
        }
        ftris = in_order(sc->fast.prim_order);
        // position of every fast-tree primitive in the reference's primitive vector: the tie rank without enableAccelStructure
        std::vector<uint32_t> posOf(nPrims, 0);
        for (uint32_t i = 0; i < nPrims; i++)
            posOf[sc->bvh.prim_order[i]] = i;
        fpos.resize(sc->fast.prim_order.size());
        for (size_t i = 0; i < fpos.size(); i++)
            fpos[i] = posOf[sc->fast.prim_order[i]];
        fnodes.resize(sc->fast.nodes.size() * kNodeRows);
        for (size_t i = 0; i < sc->fast.nodes.size(); i++) {
            const FastNode& n = sc->fast.nodes[i];
            float4* q = &fnodes[i * kNodeRows];
            q[0] = f4(n.l_lo[0], n.l_lo[1], n.l_lo[2], n.l_hi[0]);
            q[1] = f4(n.l_hi[1], n.l_hi[2], n.r_lo[0], n.r_lo[1]);
            q[2] = f4(n.r_lo[2], n.r_hi[0], n.r_hi[1], n.r_hi[2]);
            auto extent = [](const float* lo, const float* hi) { return std::max({ hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2] }); };
            q[3] = f4(bitsf(n.left), bitsf(n.right), extent(n.l_lo, n.l_hi), extent(n.r_lo, n.r_hi));
        }
    }

    // ---- eight octant-sorted copies of the inner nodes for the shadow rays (dev_scene.h onodes, trace.cuh CGE_OCTANT_NODES) ------
    std::vector<float4> onodes;
    if (CGE_OCTANT_NODES && sc->fast.nodes.size() >= (size_t(1) << 27)) { // (inner references of copy 7 must stay below 2^31 - 1)
        delete sc;
        return fail(CGE_ERR_UNSUPPORTED, "fast BVH has too many inner nodes");
    }
    if (CGE_OCTANT_NODES && !sc->fast.nodes.empty()) {
        const size_t nn = sc->fast.nodes.size();
        onodes.resize(8 * nn * kNodeRows);
        for (uint32_t o = 0; o < 8; o++)
            for (size_t i = 0; i < nn; i++) {
                const FastNode& n = sc->fast.nodes[i];
                float ln[3], lf[3], rn[3], rf[3];
                for (int a = 0; a < 3; a++) {
                    const bool neg = (o >> a) & 1u;
                    ln[a] = neg ? n.l_hi[a] : n.l_lo[a], lf[a] = neg ? n.l_lo[a] : n.l_hi[a];
                    rn[a] = neg ? n.r_hi[a] : n.r_lo[a], rf[a] = neg ? n.r_lo[a] : n.r_hi[a];
                }
                auto ref = [&](uint32_t r) { return (r & kFastLeafBit) ? r : r + o * uint32_t(nn); };
                float4* q = &onodes[(size_t(o) * nn + i) * kNodeRows];
                q[0] = f4(ln[0], ln[1], ln[2], lf[0]);
                q[1] = f4(lf[1], lf[2], rn[0], rn[1]);
                q[2] = f4(rn[2], rf[0], rf[1], rf[2]);
                q[3] = f4(bitsf(ref(n.left)), bitsf(ref(n.right)), fnodes[i * kNodeRows + 3].z, fnodes[i * kNodeRows + 3].w);
            }
    }

    // ---- the fast tree collapsed to 4 children per node (dev_scene.h f4nodes), for the shadow rays: every node takes in the
    //      children of its larger (by surface area) inner children until it has four -----------------------------------------
    std::vector<float4> f4nodes;
    uint32_t f4root = sc->fast.root;
    if (CGE_SHADOW_BVH4 && !sc->fast.nodes.empty()) { // (measured slower than the binary walk, DESIGN.md 5.9: compiled out by default)
        struct Cand {
            float lo[3], hi[3];
            uint32_t ref;
        };
        auto area = [](const Cand& c) {
            const float ex = c.hi[0] - c.lo[0], ey = c.hi[1] - c.lo[1], ez = c.hi[2] - c.lo[2];
            return ex * ey + ey * ez + ez * ex;
        };
        auto childrenOf = [&](uint32_t inner, Cand& l, Cand& r) {
            const FastNode& n = sc->fast.nodes[inner];
            std::memcpy(l.lo, n.l_lo, 12), std::memcpy(l.hi, n.l_hi, 12), l.ref = n.left;
            std::memcpy(r.lo, n.r_lo, 12), std::memcpy(r.hi, n.r_hi, 12), r.ref = n.right;
        };
        // iterative (the tree can be 60 levels deep and has ~n/2 nodes): a work list of (binary inner node, slot of its 4-wide node)
        std::vector<std::pair<uint32_t, uint32_t>> todo;
        auto newNode = [&]() {
            f4nodes.resize(f4nodes.size() + 8, f4(0, 0, 0, 0));
            return uint32_t(f4nodes.size() / 8 - 1);
        };
        f4root = newNode();
        todo.push_back({ sc->fast.root, f4root });
        while (!todo.empty()) {
            const auto [inner, slot] = todo.back();
            todo.pop_back();
            Cand c[4];
            int n = 2;
            childrenOf(inner, c[0], c[1]);
            while (n < 4) {
                int pick = -1;
                for (int k = 0; k < n; k++)
                    if (!(c[k].ref & kFastLeafBit) && (pick < 0 || area(c[k]) > area(c[pick])))
                        pick = k;
                if (pick < 0)
                    break;
                Cand l, r;
                childrenOf(c[pick].ref, l, r);
                c[pick] = l;
                c[n++] = r;
            }
            uint32_t refs[4];
            for (int k = 0; k < 4; k++) {
                if (k >= n) {
                    refs[k] = 0x7fffffffu;
                    for (int a = 0; a < 3; a++)
                        c[k].lo[a] = 3e38f, c[k].hi[a] = -3e38f;
                } else if (c[k].ref & kFastLeafBit) {
                    refs[k] = c[k].ref;
                } else {
                    refs[k] = newNode();
                    todo.push_back({ c[k].ref, refs[k] });
                }
            }
            float4* q = &f4nodes[size_t(slot) * 8];
            for (int a = 0; a < 3; a++) {
                q[a] = f4(c[0].lo[a], c[1].lo[a], c[2].lo[a], c[3].lo[a]);
                q[3 + a] = f4(c[0].hi[a], c[1].hi[a], c[2].hi[a], c[3].hi[a]);
            }
            q[6] = f4(bitsf(refs[0]), bitsf(refs[1]), bitsf(refs[2]), bitsf(refs[3]));
        }
    }

    // ---- the fast tree once more with quantised boxes (dev_scene.h qnodes), for the shadow rays ----------------------------
    std::vector<uint4> qnodes;
    float qlo[3] = { 0, 0, 0 }, qext[3] = { 1, 1, 1 };
#if CGE_QNODES
    if (!sc->fast.nodes.empty())
        quantise_fast_nodes(sc->fast, qnodes, qlo, qext);
#endif

    // ---- spheres beside the fast tree: their rows and, per sphere, the boxes the reference's traversal tests on its way to the
    //      sphere's leaf - the nodes below the root on the path to it (the root's own box is never tested, :312-361) --------------
    std::vector<float4> sphRows, sphBoxes;
    std::vector<uint32_t> sphBoxOff { 0 };
    if (haveBvh && d->n_spheres) {
        std::vector<uint32_t> posOf(nPrims, 0);
        for (uint32_t i = 0; i < nPrims; i++)
            posOf[sc->bvh.prim_order[i]] = i;
        for (uint32_t k = 0; k < d->n_spheres; k++) {
            const uint32_t gid = d->n_triangles + k, pos = posOf[gid];
            const float4* rec = &recs[size_t(gid) * kTriRows];
            sphRows.insert(sphRows.end(), rec, rec + kTriRows);
            sphRows[sphRows.size() - kTriRows + 2].x = bitsf(pos);
            uint32_t ni = sc->bvh.root;
            while (!sc->bvh.nodes[ni].is_leaf) {
                const auto& n = sc->bvh.nodes[ni];
                ni = pos < sc->bvh.nodes[n.left].end ? n.left : n.right; // children split [beg, mid) | [mid, end)
                const auto& c = sc->bvh.nodes[ni];
                sphBoxes.push_back(f4(c.lower[0], c.lower[1], c.lower[2], 0.0f));
                sphBoxes.push_back(f4(c.upper[0], c.upper[1], c.upper[2], 0.0f));
            }
            sphBoxOff.push_back(uint32_t(sphBoxes.size() / 2));
        }
    }

    // ---- inner nodes: each carries both child boxes ----------------------------------------------------------
    std::vector<float4> nodes;
    if (haveBvh) {
        std::vector<uint32_t> innerIndex(sc->bvh.nodes.size(), 0xffffffffu);
        uint32_t nInner = 0;
        for (size_t i = 0; i < sc->bvh.nodes.size(); i++)
            if (!sc->bvh.nodes[i].is_leaf)
                innerIndex[i] = nInner++;
        nodes.resize(size_t(nInner) * kNodeRows);
        auto childRef = [&](uint32_t ni, uint32_t& ref, uint32_t& cnt) {
            const auto& n = sc->bvh.nodes[ni];
            if (n.is_leaf) {
                ref = n.beg;
                cnt = n.end - n.beg;
            } else {
                ref = innerIndex[ni];
                cnt = 0;
            }
        };
        for (size_t i = 0; i < sc->bvh.nodes.size(); i++) {
            const auto& n = sc->bvh.nodes[i];
            if (n.is_leaf)
                continue;
            const auto& L = sc->bvh.nodes[n.left];
            const auto& R = sc->bvh.nodes[n.right];
            uint32_t lr, lc, rr, rc;
            childRef(n.left, lr, lc);
            childRef(n.right, rr, rc);
            float4* q = &nodes[size_t(innerIndex[i]) * kNodeRows];
            q[0] = f4(L.lower[0], L.lower[1], L.lower[2], L.upper[0]);
            q[1] = f4(L.upper[1], L.upper[2], R.lower[0], R.lower[1]);
            q[2] = f4(R.lower[2], R.upper[0], R.upper[1], R.upper[2]);
            q[3] = f4(bitsf(lr), bitsf(rr), bitsf(lc), bitsf(rc));
        }
        childRef(sc->bvh.root, sc->accel_root_ref, sc->accel_root_count);
    }

    // ---- materials: meshes then spheres -----------------------------------------------------------------------
    std::vector<float4> mats(size_t(d->n_meshes + d->n_spheres) * kMaterialRows);
    for (uint32_t m = 0; m < d->n_meshes; m++) {
        const auto& md = d->meshes[m];
        mats[size_t(m) * 3 + 0] = f4(md.kd[0], md.kd[1], md.kd[2], md.shininess);
        mats[size_t(m) * 3 + 1] = f4(md.ks[0], md.ks[1], md.ks[2], md.transparency);
        mats[size_t(m) * 3 + 2] = f4(bitsf(uint32_t(md.texture_id)), 0, 0, 0);
        if (md.triangle_count && md.transparency != 1.0f)
            sc->any_transparent = true;
    }
    for (uint32_t s = 0; s < d->n_spheres; s++) {
        const auto& sd = d->spheres[s];
        const size_t m = d->n_meshes + s;
        mats[m * 3 + 0] = f4(sd.kd[0], sd.kd[1], sd.kd[2], sd.shininess);
        mats[m * 3 + 1] = f4(sd.ks[0], sd.ks[1], sd.ks[2], sd.transparency);
        mats[m * 3 + 2] = f4(bitsf(uint32_t(-1)), 0, 0, 0); // sphere hits never sample a texture (reference :421-423)
        if (sd.transparency != 1.0f)
            sc->any_transparent = true;
    }
    std::vector<int4> texs(d->n_textures);
    for (uint32_t t = 0; t < d->n_textures; t++)
        texs[t] = make_int4(d->textures[t].width, d->textures[t].height, int(d->textures[t].texel_offset), 0);
    std::vector<float> texels(d->texels, d->texels + size_t(d->n_texels) * 3);
    sc->colours_bounded = bounded(texels.data(), texels.size());
    for (size_t m = 0; m < mats.size(); m += 3) // kd.xyz and ks.xyz (shininess and transparency are not colours)
        sc->colours_bounded = sc->colours_bounded && bounded(&mats[m].x, 3) && bounded(&mats[m + 1].x, 3);
    cudaError_t e = cudaSuccess;
    auto up = [&](auto& buf, const auto& host) {
        if (e == cudaSuccess)
            e = buf.upload(host);
    };
    up(sc->nodes, nodes);
    up(sc->tris, tris);
    up(sc->fnodes, fnodes);
    up(sc->f4nodes, f4nodes);
    up(sc->onodes, onodes);
    up(sc->qnodes, qnodes);
    up(sc->ftris, ftris);
    up(sc->sph_rows, sphRows);
    up(sc->sph_boxes, sphBoxes);
    up(sc->sph_box_off, sphBoxOff);
    up(sc->fpos, fpos);
    up(sc->shade, shade);
    up(sc->materials, mats);
    up(sc->textures, texs);
    up(sc->texels, texels);
    if (e != cudaSuccess) {
        std::string msg = std::string("scene upload: ") + cudaGetErrorString(e);
        cge_scene_destroy(sc);
        return fail(e == cudaErrorMemoryAllocation ? CGE_ERR_NOMEM : CGE_ERR_CUDA, msg);
    }
    sc->dev.nodes = sc->nodes.p;
    sc->dev.tris = sc->tris.p;
    sc->dev.fnodes = sc->fnodes.p;
    sc->dev.f4nodes = sc->f4nodes.p;
    sc->dev.onodes = sc->onodes.p;
    sc->dev.n_fnodes = uint32_t(sc->fast.nodes.size());
    sc->dev.f4root = f4root;
    sc->dev.qnodes = sc->qnodes.p;
    for (int k = 0; k < 3; k++) {
        sc->dev.qlo[k] = qlo[k];
        sc->dev.qext[k] = qext[k];
    }
    sc->dev.ftris = sc->ftris.p;
    sc->dev.n_ftris = uint32_t(ftris.size() / kTriRows);
    sc->dev.sph_rows = sc->sph_rows.p;
    sc->dev.sph_boxes = sc->sph_boxes.p;
    sc->dev.sph_box_off = sc->sph_box_off.p;
    sc->dev.fpos = sc->fpos.p;
    sc->dev.n_sph = uint32_t(sphRows.size() / kTriRows);
    sc->dev.froot = sc->fast.root;
    sc->dev.shade = sc->shade.p;
    sc->dev.materials = sc->materials.p;
    sc->dev.textures = sc->textures.p;
    sc->dev.texels = sc->texels.p;
    sc->dev.lights = nullptr; // per frame: scene_for() fills in the light-list version the frame holds
    sc->dev.n_lights = 0;
    sc->dev.n_prims = nPrims;
    sc->dev.has_spheres = d->n_spheres ? 1u : 0u;
    const int lrc = set_lights(sc, d->lights, d->n_lights);
    if (lrc != CGE_OK) {
        const std::string msg = g_err;
        cge_scene_destroy(sc);
        return fail(lrc, msg);
    }
    *out = sc;
    return CGE_OK;
}

int cge_scene_update_lights(cge_scene* sc, const cge_light_desc* lights, uint32_t n)
{
    if (!sc || (n && !lights))
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    for (uint32_t l = 0; l < n; l++)
        if (lights[l].type > CGE_LIGHT_PARALLELOGRAM)
            return fail(CGE_ERR_INVALID_ARG, "bad light type");
    CGE_CUDA(cudaSetDevice(sc->device));
    // never touches a buffer a frame in flight reads: set_lights fills a version no render holds and swaps it in
    return set_lights(sc, lights, n);
}

int cge_scene_destroy(cge_scene* sc)
{
    if (!sc)
        return CGE_OK;
    cudaSetDevice(sc->device);
    for (auto* s : sc->pool) {
        if (s->stream)
            cudaStreamSynchronize(s->stream);
        cudaFree(s->rgb);
        cudaFree(s->ids);
        cudaFree(s->tileCounter);
        cudaFree(s->counters);
        cudaFree(s->gatherRgb);
        cudaFree(s->gatherIds);
        cudaFree(s->wave.rec);
        cudaFree(s->wave.meta);
        cudaFree(s->wave.next);
        cudaFree(s->wave.dir);
        cudaFree(s->wave.vis);
        cudaFree(s->wave.sub);
        cudaFree(s->wave.cont);
        cudaFree(s->wave.hard);
        if (s->auxStream)
            cudaStreamDestroy(s->auxStream);
        if (s->evPrimary)
            cudaEventDestroy(s->evPrimary);
        if (s->evVis0)
            cudaEventDestroy(s->evVis0);
        cudaFree(s->bloomTmp);
        cudaFree(s->wave.counts);
        if (s->ev0)
            cudaEventDestroy(s->ev0);
        if (s->ev1)
            cudaEventDestroy(s->ev1);
        if (s->ev2)
            cudaEventDestroy(s->ev2);
        for (auto& ev : s->stage)
            if (ev)
                cudaEventDestroy(ev);
        for (auto& ps : s->prioStream)
            if (ps)
                cudaStreamDestroy(ps);
        if (s->evCull)
            cudaEventDestroy(s->evCull);
        if (s->grant)
            cudaFree(s->grant);
        if (s->bandDone)
            cudaEventDestroy(s->bandDone);
        if (s->copyDone)
            cudaEventDestroy(s->copyDone);
        if (s->stream)
            cudaStreamDestroy(s->stream);
        delete s;
    }
    sc->nodes.release();
    sc->tris.release();
    sc->fnodes.release();
    sc->f4nodes.release();
    sc->onodes.release();
    sc->qnodes.release();
    sc->ftris.release();
    sc->sph_rows.release();
    sc->sph_boxes.release();
    sc->sph_box_off.release();
    sc->fpos.release();
    sc->shade.release();
    sc->materials.release();
    sc->textures.release();
    sc->texels.release();
    sc->lights.reset();
    sc->light_pool.clear();
    if (sc->light_stream)
        cudaStreamDestroy(sc->light_stream);
    delete sc;
    return CGE_OK;
}

int cge_scene_bvh_info(const cge_scene* sc, uint32_t* nNodes, uint32_t* nLevels, uint32_t* nLeaves, uint32_t* maxLeaf)
{
    if (!sc)
        return fail(CGE_ERR_INVALID_ARG, "null scene");
    if (nNodes)
        *nNodes = uint32_t(sc->bvh.nodes.size());
    if (nLevels)
        *nLevels = sc->bvh.n_levels;
    if (nLeaves)
        *nLeaves = sc->bvh.n_leaves;
    if (maxLeaf)
        *maxLeaf = sc->bvh.max_leaf_prims;
    return CGE_OK;
}

int cge_bvh_build_reference_order(const cge_scene_desc* d, cge_bvh_node* nodesOut, uint32_t* nNodesInOut, uint32_t* orderOut,
    uint32_t* rootOut, uint32_t* nLevelsOut, uint32_t* nLeavesOut)
{
    if (!d || !nNodesInOut)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    if ((d->n_meshes && !d->meshes) || (d->n_vertices && !d->vertices) || (d->n_triangles && !d->triangles) || (d->n_spheres && !d->spheres))
        return fail(CGE_ERR_INVALID_ARG, "count without array");
    HostBvh bvh;
    if (!build_reference_bvh(*d, bvh)) { // no primitives
        *nNodesInOut = 0;
        return CGE_OK;
    }
    if (nodesOut && *nNodesInOut < bvh.nodes.size())
        return fail(CGE_ERR_INVALID_ARG, "nodes_out too small");
    if (nodesOut)
        std::memcpy(nodesOut, bvh.nodes.data(), bvh.nodes.size() * sizeof(cge_bvh_node));
    if (orderOut)
        std::memcpy(orderOut, bvh.prim_order.data(), bvh.prim_order.size() * sizeof(uint32_t));
    *nNodesInOut = uint32_t(bvh.nodes.size());
    if (rootOut)
        *rootOut = bvh.root;
    if (nLevelsOut)
        *nLevelsOut = bvh.n_levels;
    if (nLeavesOut)
        *nLeavesOut = bvh.n_leaves;
    return CGE_OK;
}

// Host evaluation of PixelSampler (shade.cuh): same generator (sampler.h), same fp32 operations in the same order; this TU is
// compiled with -ffp-contract=off so the host compiler cannot fuse them either.
int cge_bvh_validate(const cge_scene_desc* d, uint32_t* nLevelsOut, uint32_t* nLeavesOut)
{
    if (!d)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    HostBvh bvh;
    if (!adopt_bvh(*d, bvh))
        return fail(CGE_ERR_INVALID_ARG, "supplied BVH is inconsistent with the scene");
    if (nLevelsOut)
        *nLevelsOut = bvh.n_levels;
    if (nLeavesOut)
        *nLeavesOut = bvh.n_leaves;
    return CGE_OK;
}

int cge_ray_sample_positions(int32_t W, int32_t H, int32_t x, int32_t y, int32_t n, uint32_t seed, float* ndcOut)
{
    if (W <= 0 || H <= 0 || x < 0 || y < 0 || x >= W || y >= H || n < 1 || n > kCgeMaxRaysPerPixelSide || !ndcOut)
        return fail(CGE_ERR_INVALID_ARG, "bad pixel, frame size or rays_per_pixel_side");
    CgeMt19937Head mt(cge_aa_seed(seed, uint32_t(y) * uint32_t(W) + uint32_t(x)));
    const float px = float(x) / float(W) * 2.0f - 1.0f, py = float(y) / float(H) * 2.0f - 1.0f;
    const float bx = (1.0f / float(W) * 2.0f) / float(n), by = (1.0f / float(H) * 2.0f) / float(n);
    auto uniform = [&](float b) {
        float u = float(mt.next()) / 4294967296.0f;
        if (u >= 1.0f)
            u = 0.99999994f;
        return u * (b - 0.0f) + 0.0f;
    };
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const float jy = uniform(by);
            const float jx = uniform(bx);
            *ndcOut++ = (px + float(i) * bx) + jx;
            *ndcOut++ = (py + float(j) * by) + jy;
        }
    return CGE_OK;
}

int cge_bloom_weights(float sigma, float* out9)
{
    if (!out9)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    const BloomWeights w = weights_gaussian(sigma);
    std::memcpy(out9, w.w, sizeof(w.w));
    return CGE_OK;
}

// The light-hull pre-pass's triangle test on the host (the same function the kernel calls, wavefront.cuh cull_triangle_clear, with the
// kernel's set-up for ONE parallelogram light): per case a hit point o, a light (v0, edge01, edge02) and a triangle.  clear_out = 1
// means "no ray from o to any point of the light can be accepted by this triangle"; the CPU tests check that against the
// reference's own intersection routine on sampled rays (tests/test_hull_clear.py).  No GPU involved.
int cge_hull_clear_host(const float* o3, const float* light9, const float* tri9, uint32_t n, int32_t* clearOut)
{
    if (!o3 || !light9 || !tri9 || !clearOut)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    for (uint32_t i = 0; i < n; i++) {
        const float *po = o3 + 3 * size_t(i), *pl = light9 + 9 * size_t(i), *pt = tri9 + 9 * size_t(i);
        const vec3 o = v3(po[0], po[1], po[2]);
        const vec3 v0 = v3(pl[0], pl[1], pl[2]), e01 = v3(pl[3], pl[4], pl[5]), e02 = v3(pl[6], pl[7], pl[8]);
        const vec3 c[4] = { v0, v0 + e01, (v0 + e01) + e02, v0 + e02 }; // light_corners (wavefront.cuh), evaluated like sample_light
        vec3 a[4];
        HullDirs dirs = hull_dirs_empty();
        float lightMag = 0.0f;
        for (int j = 0; j < 4; j++) {
            a[j] = c[j] - o;
            hull_dirs_add(dirs, a[j]);
            lightMag = fmaxf(lightMag, max_abs3(c[j]));
        }
        const vec3 t0 = v3(pt[0], pt[1], pt[2]), t1 = v3(pt[3], pt[4], pt[5]), t2 = v3(pt[6], pt[7], pt[8]);
        const TriPre t = triangle_precompute(t0, t1, t2);
        const float4 rows[kTriRows] = { f4(t.n.x, t.n.y, t.n.z, t.D), f4(t0.x, t0.y, t0.z, t.e0.x), f4(t.e0.y, t.e0.z, t1.x, t1.y),
            f4(t1.z, t.e1.x, t.e1.y, t.e1.z), f4(t2.x, t2.y, t2.z, t.e2.x), f4(t.e2.y, t.e2.z, 0.0f, 0.0f) };
        clearOut[i] = cull_triangle_clear(rows, o, dirs.dmin, dirs.dmax, hull_dirs_length(dirs), true, a, max_abs3(o) + lightMag) ? 1 : 0;
    }
    return CGE_OK;
}

// ... and its box test (wavefront.cuh hull_box / hull_box_hit, with the kernel's set-up): hit_out = 0 claims that no ray from o to any
// point of the light passes through the box [lo, hi] within its length.
int cge_hull_box_host(const float* o3, const float* light9, const float* box6, uint32_t n, int32_t* hitOut)
{
    if (!o3 || !light9 || !box6 || !hitOut)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    for (uint32_t i = 0; i < n; i++) {
        const float *po = o3 + 3 * size_t(i), *pl = light9 + 9 * size_t(i), *pb = box6 + 6 * size_t(i);
        const vec3 o = v3(po[0], po[1], po[2]);
        const vec3 v0 = v3(pl[0], pl[1], pl[2]), e01 = v3(pl[3], pl[4], pl[5]), e02 = v3(pl[6], pl[7], pl[8]);
        const vec3 c[4] = { v0, v0 + e01, (v0 + e01) + e02, v0 + e02 };
        HullDirs dirs = hull_dirs_empty();
        for (int j = 0; j < 4; j++)
            hull_dirs_add(dirs, c[j] - o);
        const HullWalk walk = hull_walk(o, dirs.dmin, dirs.dmax);
        float ent, ext;
        hull_box(walk, pb[0], pb[1], pb[2], pb[3], pb[4], pb[5], ent, ext);
        hitOut[i] = hull_box_hit(ent, ext) ? 1 : 0;
    }
    return CGE_OK;
}

// The FAST traversal tree alone, by either builder (parity test of the GPU builder against the host builder).
int cge_fast_bvh_build(const cge_scene_desc* d, int onGpu, int device, cge_fast_node* nodesOut, uint32_t* nNodesInOut, uint32_t* orderOut,
    uint32_t* rootOut, uint32_t* depthOut, uint32_t* nLeavesOut, float* buildMsOut)
{
    if (!d || !nNodesInOut || (d->n_triangles && (!d->triangles || !d->vertices || !d->meshes)))
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cge_fast_node) == sizeof(FastNode), "cge_fast_node mirrors FastNode");
    FastBvh fb;
    float ms = 0.0f;
    if (onGpu) {
        int nDev = 0;
        if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
            cudaGetLastError();
            return fail(CGE_ERR_CUDA, "no CUDA device (this library has no CPU path)");
        }
        if (!sah_gpu_supported(*d))
            return fail(CGE_ERR_UNSUPPORTED, "the GPU builder takes triangle scenes with finite coordinates");
        CGE_CUDA(cudaSetDevice(device));
        std::string err;
        if (!build_sah_bvh_gpu(*d, fb, &ms, &err))
            return fail(CGE_ERR_CUDA, err);
    } else if (!build_sah_bvh(*d, fb)) {
        return fail(CGE_ERR_INVALID_ARG, "scene has no primitives");
    }
    const uint32_t room = *nNodesInOut;
    *nNodesInOut = uint32_t(fb.nodes.size());
    if (rootOut)
        *rootOut = fb.root;
    if (depthOut)
        *depthOut = fb.depth;
    if (nLeavesOut)
        *nLeavesOut = fb.n_leaves;
    if (buildMsOut)
        *buildMsOut = ms;
    if (nodesOut) {
        if (room < fb.nodes.size())
            return fail(CGE_ERR_INVALID_ARG, "nodes_out too small");
        std::memcpy(nodesOut, fb.nodes.data(), fb.nodes.size() * sizeof(FastNode));
    }
    if (orderOut)
        std::memcpy(orderOut, fb.prim_order.data(), fb.prim_order.size() * sizeof(uint32_t));
    return CGE_OK;
}

int cge_scene_bvh_export(const cge_scene* sc, cge_bvh_node* nodesOut, uint32_t* orderOut)
{
    if (!sc)
        return fail(CGE_ERR_INVALID_ARG, "null scene");
    if (nodesOut && !sc->bvh.nodes.empty())
        std::memcpy(nodesOut, sc->bvh.nodes.data(), sc->bvh.nodes.size() * sizeof(cge_bvh_node));
    if (orderOut && !sc->bvh.prim_order.empty())
        std::memcpy(orderOut, sc->bvh.prim_order.data(), sc->bvh.prim_order.size() * sizeof(uint32_t));
    return CGE_OK;
}

int cge_render(cge_scene* sc, const cge_camera* cam, const cge_params* p, float* rgbOut, int32_t* idsOut, cge_stats* st)
{
    int rc = validate_params(sc, p);
    if (rc)
        return rc;
    if (!cam || !rgbOut)
        return fail(CGE_ERR_INVALID_ARG, "null camera or output");
    const bool wantIds = (p->flags & CGE_FLAG_WANT_PRIM_IDS) && idsOut;
    const bool rgba8 = p->flags & CGE_FLAG_OUTPUT_RGBA8;
    const bool devOut = (p->flags & CGE_FLAG_RGB_DEVICE_PTR) && !rgba8;
    if (rgba8 && ((p->flags & CGE_FLAG_RGB_DEVICE_PTR) || p->part_count > 1))
        return fail(CGE_ERR_UNSUPPORTED, "CGE_FLAG_OUTPUT_RGBA8 writes a whole host frame");
    if ((p->features & CGE_FEAT_BLOOM_EFFECT) && p->part_count > 1)
        return fail(CGE_ERR_UNSUPPORTED,
            "the bloom filter reads neighbouring pixels: render the whole frame, or use cge_render_distributed");
    CGE_CUDA(cudaSetDevice(sc->device));
    const size_t pixels = size_t(p->width) * size_t(p->height);
    Scratch* s = nullptr;
    std::vector<float> partRgb; // a partition rendered to a host frame: its packed tiles (below)
    std::vector<int> partIds;
    const LightsRef lights = lights_of(sc); // this frame's version of the light list, held until the stream has drained
    const LightSet& ls = *lights;
    const DevParams dp = make_dev_params(sc, *p, ls);
    const bool partToHost = p->part_count > 1 && !devOut;
    const size_t partPixels = partToHost ? size_t(dp.tile_count) * 32 : 0;
    // rgba8: the float frame lives in the first 3/4 of a doubled scratch frame, the packed bytes in the ids buffer
    rc = acquire_scratch(sc, devOut ? 1 : pixels, wantIds || rgba8, partPixels, &s);
    if (rc) {
        release_scratch(sc, s);
        return rc;
    }
    float* rgbDev = devOut ? rgbOut : s->rgb;
    int* idsDev = wantIds ? (devOut ? idsOut : s->ids) : nullptr;
    uint32_t launches = 0;
    // A large frame is rendered as several horizontal bands of tile rows (launch_bands); a band that is finished travels to
    // the host while the others are still traced.  Not with bloom (it needs the complete frame), not when the ids share their
    // buffer with the packed pixels, and a partition (part_count > 1) is left to cge_render_distributed.
    unsigned nBands = 1;
    if (dp.part_count <= 1 && !(p->features & CGE_FEAT_BLOOM_EFFECT) && !(rgba8 && wantIds))
        nBands = nBandsFor(sc, ls, *p, dp, !devOut);
    std::vector<Scratch*> bands;
    if (acquire_bands(sc, s, nBands, bands) != CGE_OK) { // no room for the helpers' queues: one pipeline is always possible
        cudaGetLastError();
        nBands = 1;
    }
    if (nBands > 1)
        rc = use_band_stream(s, 0);
    if (rc != CGE_OK) {
        for (Scratch* b : bands)
            release_scratch(sc, b);
        return rc;
    }
    cudaEventRecord(s->ev0, s->stream);
    if (nBands > 1) {
        std::vector<uint2> ranges;
        std::vector<unsigned> rows; // tile-row boundaries of the bands
        for (unsigned b = 0; b <= nBands; b++)
            rows.push_back(unsigned(uint64_t(dp.n_tiles_y) * b / nBands));
        for (unsigned b = 0; b < nBands; b++)
            ranges.push_back(make_uint2(rows[b] * dp.n_tiles_x, (rows[b + 1] - rows[b]) * dp.n_tiles_x));
        // frame rows of band b: tile rows [r0, r1) = reference rows y in [4 r0, min(4 r1, H)) = rows [H - yEnd, H - yBeg) (y flip)
        auto span = [&](unsigned b, size_t& first, size_t& count) {
            const int yBeg = int(rows[b]) * kTileH, yEnd = std::min(int(rows[b + 1]) * kTileH, p->height);
            first = size_t(p->height - yEnd) * size_t(p->width);
            count = size_t(yEnd - yBeg) * size_t(p->width);
        };
        rc = launch_bands(sc, ls, bands, ranges, !devOut, cam, p, dp, rgbDev, idsDev, &launches, [&](unsigned b, Scratch* sb) {
            if (rgba8) {
                size_t first, count;
                span(b, first, count);
                pack_rgba8_kernel<<<unsigned((count + 255) / 256), 256, 0, sb->stream>>>(s->rgb + first * 3,
                    reinterpret_cast<uchar4*>(s->ids) + first, count);
                launches++;
            }
            cudaEventRecord(sb->bandDone, sb->stream);
        });
        cudaEventRecord(s->ev1, s->stream); // every band's kernels (launch_bands made this stream wait for them)
        // The copies are queued only now that every band's kernels are: into pageable host memory cudaMemcpyAsync blocks the
        // calling thread until the copy is done, which would otherwise hold back the launch of the next band.
        for (unsigned b = 0; b < nBands && rc == CGE_OK; b++) {
            Scratch* sb = bands[b];
            size_t first, count;
            span(b, first, count);
            if (rgba8) {
                cudaMemcpyAsync(reinterpret_cast<uchar4*>(rgbOut) + first, reinterpret_cast<uchar4*>(s->ids) + first, count * 4,
                    cudaMemcpyDeviceToHost, sb->stream);
            } else if (!devOut) {
                cudaMemcpyAsync(rgbOut + first * 3, s->rgb + first * 3, count * 3 * sizeof(float), cudaMemcpyDeviceToHost, sb->stream);
                if (wantIds)
                    cudaMemcpyAsync(idsOut + first, s->ids + first, count * sizeof(int), cudaMemcpyDeviceToHost, sb->stream);
            }
            cudaEventRecord(sb->copyDone, sb->stream);
        }
        for (unsigned b = 1; b < bands.size(); b++)
            cudaStreamWaitEvent(s->stream, bands[b]->copyDone, 0);
    } else {
        rc = launch_render(sc, ls, s, cam, p, dp, rgbDev, idsDev, &launches);
        if (rc == CGE_OK && (p->features & CGE_FEAT_BLOOM_EFFECT))
            rc = apply_bloom(s, *p, rgbDev, &launches);
        cudaEventRecord(s->ev1, s->stream);
    }
    if (nBands > 1) {
        // every band is already on its way
    } else if (rc == CGE_OK && rgba8) {
        // ids (if wanted) leave first, then their buffer is reused for the packed pixels (4 bytes per pixel either way)
        if (wantIds)
            cudaMemcpyAsync(idsOut, s->ids, pixels * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
        pack_rgba8_kernel<<<unsigned((pixels + 255) / 256), 256, 0, s->stream>>>(s->rgb, reinterpret_cast<uchar4*>(s->ids), pixels);
        launches++;
        cudaMemcpyAsync(rgbOut, s->ids, pixels * 4, cudaMemcpyDeviceToHost, s->stream);
    } else if (rc == CGE_OK && !devOut) {
        if (dp.part_count <= 1) {
            cudaMemcpyAsync(rgbOut, s->rgb, pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
            if (wantIds)
                cudaMemcpyAsync(idsOut, s->ids, pixels * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
        } else {
            // partial frame (one partition's tiles, the rest of rgb_out stays untouched): pack the tiles contiguously with the
            // gather kernel of the multi-GPU path, ONE copy to the host, scattered into the frame below once the stream drained
            if (dp.tile_count) {
                pack_tiles_kernel<<<(dp.tile_count * 32 + 255) / 256, 256, 0, s->stream>>>(s->rgb, wantIds ? s->ids : nullptr, s->gatherRgb,
                    wantIds ? s->gatherIds : nullptr, dp, dp.tile_count);
                launches++;
                partRgb.resize(size_t(dp.tile_count) * 32 * 3);
                cudaMemcpyAsync(partRgb.data(), s->gatherRgb, partRgb.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
                if (wantIds) {
                    partIds.resize(size_t(dp.tile_count) * 32);
                    cudaMemcpyAsync(partIds.data(), s->gatherIds, partIds.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
                }
            }
        }
    }
    cudaEventRecord(s->ev2, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (rc == CGE_OK && e != cudaSuccess)
        rc = fail(CGE_ERR_CUDA, std::string("render: ") + cudaGetErrorString(e));
    if (rc == CGE_OK && !partRgb.empty()) // packed tile k, pixel j (row-major inside the 8x4 tile) -> Screen layout
        for (unsigned k = 0; k < dp.tile_count; k++) {
            const unsigned tile = part_tile_of(dp.part_unit, dp.part_index, dp.part_count, k);
            const int x0 = int(tile % dp.n_tiles_x) * kTileW, y0 = int(tile / dp.n_tiles_x) * kTileH;
            const int w = std::min(kTileW, p->width - x0);
            for (int y = y0; y < std::min(y0 + kTileH, p->height); y++) {
                const size_t idx = size_t(p->height - 1 - y) * size_t(p->width) + size_t(x0);
                const size_t src = size_t(k) * 32 + size_t(y - y0) * kTileW;
                std::memcpy(rgbOut + idx * 3, partRgb.data() + src * 3, size_t(w) * 3 * sizeof(float));
                if (wantIds)
                    std::memcpy(idsOut + idx, partIds.data() + src, size_t(w) * sizeof(int));
            }
        }
    for (unsigned b = 1; b < bands.size(); b++) {
        const cudaError_t eb = cudaStreamSynchronize(bands[b]->stream);
        if (rc == CGE_OK && eb != cudaSuccess)
            rc = fail(CGE_ERR_CUDA, std::string("render: ") + cudaGetErrorString(eb));
    }
    if (rc == CGE_OK)
        rc = fill_stats(bands, st, launches);
    for (Scratch* b : bands)
        release_scratch(sc, b);
    return rc;
}

int cge_trace_rays(cge_scene* sc, const float* rays7, uint32_t n, const cge_params* pIn, float* rgbOut, int32_t* idsOut)
{
    if (!pIn)
        return fail(CGE_ERR_INVALID_ARG, "null params");
    cge_params p = *pIn;
    p.width = p.width > 0 ? p.width : 1;
    p.height = p.height > 0 ? p.height : 1;
    int rc = validate_params(sc, &p);
    if (rc)
        return rc;
    if (!rays7 || !rgbOut)
        return fail(CGE_ERR_INVALID_ARG, "null rays or output");
    if (n == 0)
        return CGE_OK;
    CGE_CUDA(cudaSetDevice(sc->device));
    Scratch* s = nullptr;
    rc = acquire_scratch(sc, n, true, 0, &s);
    if (rc) {
        release_scratch(sc, s);
        return rc;
    }
    float* dRays = nullptr;
    cudaError_t e = cudaMalloc(&dRays, size_t(n) * 7 * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(dRays, rays7, size_t(n) * 7 * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(s->counters, 0, sizeof(Counters), s->stream);
    const LightsRef lights = lights_of(sc);
    const DevParams dp = make_dev_params(sc, p, *lights);
    const DevScene ds = scene_for(sc, p, *lights);
    Variant v = choose_variant(ds, p, dp);
    v.count = false;
    if (e == cudaSuccess) {
        dispatch(v, [&](auto kFast, auto kSpheres, auto kCount) {
            trace_rays_kernel<decltype(kFast)::value, decltype(kSpheres)::value, decltype(kCount)::value>
                <<<(n + 127) / 128, 128, 0, s->stream>>>(ds, dp, dRays, n, s->rgb, idsOut ? s->ids : nullptr, s->counters);
            e = cudaGetLastError();
        });
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(rgbOut, s->rgb, size_t(n) * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && idsOut)
        e = cudaMemcpyAsync(idsOut, s->ids, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    cudaError_t e2 = cudaStreamSynchronize(s->stream);
    cudaFree(dRays);
    release_scratch(sc, s);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(CGE_ERR_CUDA, std::string("trace_rays: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    return CGE_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// KAT kernels: the six libIntersect functions evaluated on the device, one case per thread
// ---------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ Ray load_ray(const float* q) { return Ray { v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), q[6] }; }

__global__ void kat_triangle_kernel(const float* v, float* ray7, int* hit, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = v + size_t(i) * 9;
    Ray r = load_ray(ray7 + size_t(i) * 7);
    hit[i] = intersect_triangle(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), v3(a[6], a[7], a[8]), r) ? 1 : 0;
    ray7[size_t(i) * 7 + 6] = r.t;
}
// Same test through the precomputed-plane path used by the traversal kernel (must agree bit for bit with I4).
__global__ void kat_triangle_pre_kernel(const float* v, float* ray7, int* hit, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = v + size_t(i) * 9;
    const vec3 v0 = v3(a[0], a[1], a[2]), v1 = v3(a[3], a[4], a[5]), v2 = v3(a[6], a[7], a[8]);
    Ray r = load_ray(ray7 + size_t(i) * 7);
    const TriPre t = triangle_precompute(v0, v1, v2);
    const float tt = fdiv(fsub(t.D, dot(r.o, t.n)), dot(r.d, t.n));
    bool ok = (tt >= 0.0f) && (r.t >= tt);
    if (ok) {
        const vec3 p = r.d * tt + r.o;
        ok = dot(t.e0, p - v0) >= 0.0f && dot(t.e1, p - v1) >= 0.0f && dot(t.e2, p - v2) >= 0.0f;
    }
    hit[i] = ok ? 1 : 0;
    if (ok)
        ray7[size_t(i) * 7 + 6] = tt;
}
__global__ void kat_aabb_kernel(const float* b, float* ray7, int* hit, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = b + size_t(i) * 6;
    Ray r = load_ray(ray7 + size_t(i) * 7);
    hit[i] = intersect_aabb(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), r) ? 1 : 0;
    ray7[size_t(i) * 7 + 6] = r.t;
}
__global__ void kat_sphere_kernel(const float* sp, float* ray7, float* nrm, int* hit, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = sp + size_t(i) * 4;
    Ray r = load_ray(ray7 + size_t(i) * 7);
    vec3 nn = v3(0.0f);
    hit[i] = intersect_sphere(v3(a[0], a[1], a[2]), a[3], r, &nn) ? 1 : 0;
    ray7[size_t(i) * 7 + 6] = r.t;
    nrm[size_t(i) * 3] = nn.x, nrm[size_t(i) * 3 + 1] = nn.y, nrm[size_t(i) * 3 + 2] = nn.z;
}
__global__ void kat_plane_kernel(const float* pl, float* ray7, int* hit, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = pl + size_t(i) * 4;
    Ray r = load_ray(ray7 + size_t(i) * 7);
    Plane p { a[0], v3(a[1], a[2], a[3]) };
    hit[i] = intersect_plane(p, r) ? 1 : 0;
    ray7[size_t(i) * 7 + 6] = r.t;
}
__global__ void kat_triangle_plane_kernel(const float* v, float* out4, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = v + size_t(i) * 9;
    const Plane p = triangle_plane(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), v3(a[6], a[7], a[8]));
    out4[size_t(i) * 4] = p.D, out4[size_t(i) * 4 + 1] = p.n.x, out4[size_t(i) * 4 + 2] = p.n.y, out4[size_t(i) * 4 + 3] = p.n.z;
}
__global__ void kat_pit_kernel(const float* v, const float* nrm, const float* pt, int* inside, unsigned n)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* a = v + size_t(i) * 9;
    inside[i] = point_in_triangle(v3(a[0], a[1], a[2]), v3(a[3], a[4], a[5]), v3(a[6], a[7], a[8]),
                    v3(nrm[size_t(i) * 3], nrm[size_t(i) * 3 + 1], nrm[size_t(i) * 3 + 2]),
                    v3(pt[size_t(i) * 3], pt[size_t(i) * 3 + 1], pt[size_t(i) * 3 + 2]))
        ? 1
        : 0;
}

struct KatBufs {
    std::vector<void*> dev;
    ~KatBufs()
    {
        for (void* p : dev)
            cudaFree(p);
    }
    template <typename T>
    T* in(const T* host, size_t count, cudaError_t& e)
    {
        T* d = nullptr;
        if (e == cudaSuccess)
            e = cudaMalloc(&d, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) {
            dev.push_back(d);
            if (host)
                e = cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice);
        }
        return d;
    }
};

int kat_begin(int device)
{
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
        cudaGetLastError();
        return fail(CGE_ERR_CUDA, "no CUDA device (the KAT entry points run the device functions; there is no CPU path)");
    }
    if (device < 0 || device >= nDev)
        return fail(CGE_ERR_INVALID_ARG, "bad device ordinal");
    CGE_CUDA(cudaSetDevice(device));
    return CGE_OK;
}
int kat_end(cudaError_t e)
{
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(CGE_ERR_CUDA, std::string("kat: ") + cudaGetErrorString(e));
    return CGE_OK;
}
} // namespace

extern "C" {

int cge_kat_triangle(const float* v, float* ray7, int32_t* hit, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(v, size_t(n) * 9, e);
    float* dr = b.in(ray7, size_t(n) * 7, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_triangle_kernel<<<(n + 255) / 256, 256>>>(dv, dr, dh, n);
        e = cudaMemcpy(ray7, dr, size_t(n) * 7 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(hit, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
// variant through the precomputed rows (not in the reference API; used to prove the precompute is exact)
int cge_kat_triangle_precomputed(const float* v, float* ray7, int32_t* hit, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(v, size_t(n) * 9, e);
    float* dr = b.in(ray7, size_t(n) * 7, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_triangle_pre_kernel<<<(n + 255) / 256, 256>>>(dv, dr, dh, n);
        e = cudaMemcpy(ray7, dr, size_t(n) * 7 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(hit, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
int cge_kat_aabb(const float* box, float* ray7, int32_t* hit, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(box, size_t(n) * 6, e);
    float* dr = b.in(ray7, size_t(n) * 7, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_aabb_kernel<<<(n + 255) / 256, 256>>>(dv, dr, dh, n);
        e = cudaMemcpy(ray7, dr, size_t(n) * 7 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(hit, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
int cge_kat_sphere(const float* sp, float* ray7, float* nrm, int32_t* hit, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(sp, size_t(n) * 4, e);
    float* dr = b.in(ray7, size_t(n) * 7, e);
    float* dn = b.in<float>(nullptr, size_t(n) * 3, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_sphere_kernel<<<(n + 255) / 256, 256>>>(dv, dr, dn, dh, n);
        e = cudaMemcpy(ray7, dr, size_t(n) * 7 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(nrm, dn, size_t(n) * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(hit, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
int cge_kat_plane(const float* pl, float* ray7, int32_t* hit, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(pl, size_t(n) * 4, e);
    float* dr = b.in(ray7, size_t(n) * 7, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_plane_kernel<<<(n + 255) / 256, 256>>>(dv, dr, dh, n);
        e = cudaMemcpy(ray7, dr, size_t(n) * 7 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(hit, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
int cge_kat_triangle_plane(const float* v, float* out4, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(v, size_t(n) * 9, e);
    float* dout = b.in<float>(nullptr, size_t(n) * 4, e);
    if (e == cudaSuccess) {
        kat_triangle_plane_kernel<<<(n + 255) / 256, 256>>>(dv, dout, n);
        e = cudaMemcpy(out4, dout, size_t(n) * 4 * sizeof(float), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}
int cge_kat_point_in_triangle(const float* v, const float* nrm, const float* pt, int32_t* inside, uint32_t n, int device)
{
    int rc = kat_begin(device);
    if (rc || n == 0)
        return rc;
    KatBufs b;
    cudaError_t e = cudaSuccess;
    float* dv = b.in(v, size_t(n) * 9, e);
    float* dn = b.in(nrm, size_t(n) * 3, e);
    float* dp = b.in(pt, size_t(n) * 3, e);
    int* dh = b.in<int>(nullptr, n, e);
    if (e == cudaSuccess) {
        kat_pit_kernel<<<(n + 255) / 256, 256>>>(dv, dn, dp, dh, n);
        e = cudaMemcpy(inside, dh, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost);
    }
    return kat_end(e);
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU: NCCL (loaded lazily with dlopen so that the library itself has no link-time NCCL dependency)
// ---------------------------------------------------------------------------------------------------------------
int cge_comm_unique_id(uint8_t idOut[CGE_UNIQUE_ID_BYTES])
{
    const NcclApi* api = nccl_api();
    if (!api)
        return fail(CGE_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess)
        return fail(CGE_ERR_NCCL, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == CGE_UNIQUE_ID_BYTES, "unique id size");
    std::memcpy(idOut, &id, sizeof(id));
    return CGE_OK;
}

int cge_comm_create(const uint8_t idIn[CGE_UNIQUE_ID_BYTES], int rank, int nRanks, int device, cge_comm** out)
{
    if (!idIn || !out || rank < 0 || nRanks < 1 || nRanks > 16 || rank >= nRanks)
        return fail(CGE_ERR_INVALID_ARG, "bad comm arguments (1..16 ranks)");
    const NcclApi* api = nccl_api();
    if (!api)
        return fail(CGE_ERR_NCCL, "libnccl.so.2 not found");
    CGE_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, idIn, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = api->CommInitRank(&comm, nRanks, id, rank);
    if (r != ncclSuccess)
        return fail(CGE_ERR_NCCL, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
    auto* c = new cge_comm();
    c->rank = rank;
    c->n_ranks = nRanks;
    c->device = device;
    c->nccl = comm;
    *out = c;
    return CGE_OK;
}

int cge_comm_destroy(cge_comm* c)
{
    if (!c)
        return CGE_OK;
    const NcclApi* api = nccl_api();
    for (const auto& f : c->peer_frames) {
        if (c->rank == 0)
            cudaFree(f.ptr);
        else
            cudaIpcCloseMemHandle(f.ptr);
    }
    for (const auto& f : c->host_frames) {
        cudaHostUnregister(f.ptr);
        munmap(f.ptr, f.bytes);
    }
    if (api && c->nccl)
        api->CommDestroy(static_cast<ncclComm_t>(c->nccl));
    delete c;
    return CGE_OK;
}

// A frame in host memory that every rank of the communicator maps (POSIX shared memory, page-locked in each process): with
// CGE_FLAG_SHARED_HOST_FRAME each rank copies the image rows it rendered straight into it over its own PCIe link.  Collective.
int cge_comm_host_frame(cge_comm* comm, uint64_t bytes, void** out)
{
    if (!comm || !out || bytes == 0)
        return fail(CGE_ERR_INVALID_ARG, "bad host frame arguments");
    const NcclApi* api = nccl_api();
    if (!api)
        return fail(CGE_ERR_NCCL, "libnccl.so.2 not found");
    CGE_CUDA(cudaSetDevice(comm->device));
    ncclComm_t nc = static_cast<ncclComm_t>(comm->nccl);
    char name[64] = {};
    static std::atomic<unsigned> serial { 0 };
    if (comm->rank == 0)
        std::snprintf(name, sizeof name, "/cge_frame_%d_%u", int(getpid()), serial.fetch_add(1));
    char* dName = nullptr;
    int* dOk = nullptr;
    CGE_CUDA(cudaMalloc(&dName, sizeof name));
    CGE_CUDA(cudaMalloc(&dOk, sizeof(int)));
    void* ptr = MAP_FAILED;
    int fd = -1;
    if (comm->rank == 0) { // the creator sizes the segment before anyone else opens it
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd >= 0 && ftruncate(fd, off_t(bytes)) != 0) {
            close(fd);
            fd = -1;
        }
    }
    cudaMemcpy(dName, name, sizeof name, cudaMemcpyHostToDevice);
    ncclResult_t nr = api->Broadcast(dName, dName, sizeof name, ncclChar, 0, nc, nullptr);
    cudaError_t ce = cudaMemcpy(name, dName, sizeof name, cudaMemcpyDeviceToHost); // (legacy stream: ordered after the broadcast)
    if (comm->rank != 0 && nr == ncclSuccess && ce == cudaSuccess && name[0])
        fd = shm_open(name, O_RDWR, 0600);
    if (fd >= 0) {
        ptr = mmap(nullptr, size_t(bytes), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
    }
    bool registered = false;
    if (ptr != MAP_FAILED)
        registered = cudaHostRegister(ptr, size_t(bytes), cudaHostRegisterPortable) == cudaSuccess;
    // every rank learns whether every rank succeeded; only then is the name removed (the mappings keep the memory alive)
    int okHere = (ptr != MAP_FAILED && registered) ? 1 : 0, okAll = 0;
    cudaMemcpy(dOk, &okHere, sizeof(int), cudaMemcpyHostToDevice);
    if (nr == ncclSuccess)
        nr = api->AllReduce(dOk, dOk, 1, ncclInt, ncclMin, nc, nullptr);
    cudaMemcpy(&okAll, dOk, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dName);
    cudaFree(dOk);
    if (comm->rank == 0 && name[0])
        shm_unlink(name);
    if (nr != ncclSuccess || !okAll) {
        cudaGetLastError();
        if (registered)
            cudaHostUnregister(ptr);
        if (ptr != MAP_FAILED)
            munmap(ptr, size_t(bytes));
        return fail(nr != ncclSuccess ? CGE_ERR_NCCL : CGE_ERR_NOMEM, "could not map the shared host frame on every rank");
    }
    comm->host_frames.push_back({ ptr, size_t(bytes) });
    *out = ptr;
    return CGE_OK;
}

namespace {
inline size_t peer_counter_offset(size_t bytes) { return (bytes + 255) / 256 * 256; }

// Dynamic tile dealing (CGE_FLAG_DYNAMIC_TILES): the head of a pipeline takes the next chunk of the frame's pool from the counter
// beside rank 0's frame - a system-scope atomic over NVLink - and leaves it where the pipeline's kernels read their tile range
// (DevParams::grant).  Chunk c belongs to the tile list of rank c % nRanks: chunk c / nRanks of the first pool rows of that list.
__global__ void grant_kernel(unsigned* counter, unsigned base, unsigned nRanks, unsigned chunksPerOwner, unsigned chunkRows, unsigned poolPct,
    unsigned nTilesX, unsigned nTilesY, unsigned* grant)
{
    const unsigned idx = atomicAdd_system(counter, 1u) - base;
    unsigned first = 0, count = 0, owner = 0;
    if (idx < nRanks * chunksPerOwner) {
        owner = idx % nRanks;
        const unsigned j = idx / nRanks;
        const unsigned rows = nTilesY > owner ? (nTilesY - owner + nRanks - 1) / nRanks : 0u;
        const unsigned pool = rows * poolPct / 100u;
        if (j * chunkRows < pool) {
            first = j * chunkRows * nTilesX;
            count = min(chunkRows, pool - j * chunkRows) * nTilesX;
        }
    }
    grant[0] = first, grant[1] = count, grant[2] = owner;
}
} // namespace

int cge_comm_peer_frame(cge_comm* comm, uint64_t bytes, void** out)
{
    if (!comm || !out || bytes == 0)
        return fail(CGE_ERR_INVALID_ARG, "bad peer frame arguments");
    const NcclApi* api = nccl_api();
    if (!api)
        return fail(CGE_ERR_NCCL, "libnccl.so.2 not found");
    CGE_CUDA(cudaSetDevice(comm->device));
    ncclComm_t nc = static_cast<ncclComm_t>(comm->nccl);
    cudaIpcMemHandle_t handle {};
    void* ptr = nullptr;
    bool okHere = true;
    if (comm->rank == 0) // the frame lives on rank 0's GPU; its handle travels over the communicator
        okHere = cudaMalloc(&ptr, peer_counter_offset(size_t(bytes)) + 256) == cudaSuccess
            && cudaMemset(static_cast<char*>(ptr) + peer_counter_offset(size_t(bytes)), 0, 256) == cudaSuccess
            && cudaIpcGetMemHandle(&handle, ptr) == cudaSuccess;
    char* dHandle = nullptr;
    int* dOk = nullptr;
    CGE_CUDA(cudaMalloc(&dHandle, sizeof handle));
    CGE_CUDA(cudaMalloc(&dOk, sizeof(int)));
    cudaMemcpy(dHandle, &handle, sizeof handle, cudaMemcpyHostToDevice);
    ncclResult_t nr = comm->n_ranks > 1 ? api->Broadcast(dHandle, dHandle, sizeof handle, ncclChar, 0, nc, nullptr) : ncclSuccess;
    cudaError_t ce = cudaMemcpy(&handle, dHandle, sizeof handle, cudaMemcpyDeviceToHost); // (legacy stream: ordered after the broadcast)
    if (comm->rank != 0)
        okHere = nr == ncclSuccess && ce == cudaSuccess && cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    int ok = okHere ? 1 : 0, okAll = 0;
    cudaMemcpy(dOk, &ok, sizeof(int), cudaMemcpyHostToDevice);
    if (nr == ncclSuccess && comm->n_ranks > 1)
        nr = api->AllReduce(dOk, dOk, 1, ncclInt, ncclMin, nc, nullptr);
    cudaMemcpy(&okAll, dOk, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dHandle);
    cudaFree(dOk);
    if (nr != ncclSuccess || !okAll) {
        cudaGetLastError();
        if (ptr && okHere) {
            if (comm->rank == 0)
                cudaFree(ptr);
            else
                cudaIpcCloseMemHandle(ptr);
        }
        return fail(nr != ncclSuccess ? CGE_ERR_NCCL : CGE_ERR_CUDA, "could not map rank 0's device frame on every rank (CUDA IPC / peer access)");
    }
    comm->peer_frames.push_back({ ptr, size_t(bytes) });
    comm->peer_grants.push_back(0u);
    *out = ptr;
    return CGE_OK;
}

int cge_render_distributed(cge_scene* sc, cge_comm* comm, const cge_camera* cam, const cge_params* pIn, float* rgbOut,
    int32_t* idsOut, cge_stats* st)
{
    if (!comm || !pIn)
        return fail(CGE_ERR_INVALID_ARG, "null comm or params");
    cge_params p = *pIn;
    p.part_index = uint32_t(comm->rank);
    p.part_count = uint32_t(comm->n_ranks);
    p.flags |= CGE_FLAG_PARTITION_TILE_ROWS; // a rank's share = runs of 4 complete image rows, contiguous in the frame
    int rc = validate_params(sc, &p);
    if (rc)
        return rc;
    const bool wantIds = (p.flags & CGE_FLAG_WANT_PRIM_IDS) != 0;
    const bool rgba8 = p.flags & CGE_FLAG_OUTPUT_RGBA8;
    const bool devOut = (p.flags & CGE_FLAG_RGB_DEVICE_PTR) || (p.flags & CGE_FLAG_PEER_FRAME);
    // every rank writes its rows into the shared host frame itself (not with bloom: it needs the gathered frame on one GPU)
    const bool shared = (p.flags & CGE_FLAG_SHARED_HOST_FRAME) && comm->n_ranks > 1 && !(p.features & CGE_FEAT_BLOOM_EFFECT);
    // every rank's kernels store straight into rank 0's device frame (cge_comm_peer_frame)
    const bool peer = (p.flags & CGE_FLAG_PEER_FRAME) && comm->n_ranks > 1;
    if (!cam || ((comm->rank == 0 || shared || peer) && !rgbOut))
        return fail(CGE_ERR_INVALID_ARG, "null camera or output");
    if ((p.flags & CGE_FLAG_PEER_FRAME) && (wantIds || rgba8 || (p.flags & CGE_FLAG_SHARED_HOST_FRAME)))
        return fail(CGE_ERR_UNSUPPORTED, "CGE_FLAG_PEER_FRAME delivers the float frame on rank 0's GPU (no ids, no RGBA8, no host frame)");
    size_t peerIdx = 0;
    if ((p.flags & CGE_FLAG_DYNAMIC_TILES) && !(p.flags & CGE_FLAG_PEER_FRAME))
        return fail(CGE_ERR_UNSUPPORTED, "CGE_FLAG_DYNAMIC_TILES needs CGE_FLAG_PEER_FRAME (dealt chunks have no fixed place in a rank's buffer)");
    if (p.flags & CGE_FLAG_PEER_FRAME) {
        bool known = false;
        for (size_t i = 0; i < comm->peer_frames.size(); i++) {
            const auto& f = comm->peer_frames[i];
            if (f.ptr == rgbOut && f.bytes >= size_t(p.width) * size_t(p.height) * 12)
                known = true, peerIdx = i;
        }
        if (!known)
            return fail(CGE_ERR_INVALID_ARG, "CGE_FLAG_PEER_FRAME: rgb_out must be the pointer cge_comm_peer_frame returned on this rank");
    }
    if (sc->device != comm->device)
        return fail(CGE_ERR_INVALID_ARG, "scene and communicator live on different devices");
    if ((rgba8 || (p.flags & CGE_FLAG_SHARED_HOST_FRAME)) && devOut)
        return fail(CGE_ERR_UNSUPPORTED, "CGE_FLAG_OUTPUT_RGBA8 / CGE_FLAG_SHARED_HOST_FRAME write a host frame");
    if (shared && rgba8)
        return fail(CGE_ERR_UNSUPPORTED, "CGE_FLAG_SHARED_HOST_FRAME carries the float frame");
    if (shared) {
        bool known = false;
        for (const auto& f : comm->host_frames)
            known = known || (f.ptr == rgbOut && f.bytes >= size_t(p.width) * size_t(p.height) * 12);
        if (!known)
            return fail(CGE_ERR_INVALID_ARG, "CGE_FLAG_SHARED_HOST_FRAME: rgb_out must be a frame cge_comm_host_frame returned");
    }
    const NcclApi* api = nccl_api();
    CGE_CUDA(cudaSetDevice(sc->device));
    const size_t W = size_t(p.width), pixels = W * size_t(p.height);
    const LightsRef lights = lights_of(sc);
    const LightSet& ls = *lights;
    DevParams dp = make_dev_params(sc, p, ls);
    const unsigned R = unsigned(comm->n_ranks), rank = unsigned(comm->rank);
    // a rank's share: units = tile rows rank, rank + R, ...; COMPACT layout (dev_scene.h): units(r) blocks of 4 x W pixels
    auto unitsOf = [&](unsigned r) { return dp.n_tiles_y > r ? (dp.n_tiles_y - r + R - 1) / R : 0u; };
    const unsigned myUnits = unitsOf(rank);
    const bool compact = R > 1 && (rank != 0 || shared) && !peer; // rank 0 renders straight into the frame it gathers
    dp.compact_units = compact ? std::max(myUnits, 1u) : 0u;
    const size_t blockPixels = size_t(kTileH) * W;
    size_t othersPixels = 0;
    UnpackRows up {};
    if (rank == 0 && !shared && !peer)
        for (unsigned r = 1; r < R; r++) {
            up.offset[r] = unsigned(othersPixels / blockPixels);
            up.units[r] = unitsOf(r);
            othersPixels += size_t(unitsOf(r)) * blockPixels;
        }
    // ---- dynamic tile dealing: the first poolPct % of every rank's tile rows form a pool of chunks that is dealt by a counter in
    // rank 0's memory to whichever rank gets there first; each rank renders the rest of its rows (the static part) as before, then
    // K pipelines that each take one chunk (or nothing) ------------------------------------------------------------------------
    dp.chain_unsplit = peer ? 1u : 0u;
    const bool dynamic = peer && (p.flags & CGE_FLAG_DYNAMIC_TILES) && !dp.aa_side;
    const unsigned poolPct = dynamic ? unsigned(std::min(std::max(env_int("CGE_DYNAMIC_POOL_PCT", 25), 1), 90)) : 0u;
    const unsigned chunksPerOwner = unsigned(std::min(std::max(env_int("CGE_DYNAMIC_CHUNKS", 2), 1), 8));
    const unsigned kGrants = chunksPerOwner + 1; // pipelines per rank: N * K grants >= N * chunksPerOwner chunks, every chunk is taken
    const unsigned poolRows = unitsOf(rank) * poolPct / 100u, maxPoolRows = unitsOf(0) * poolPct / 100u;
    const unsigned chunkRows = std::max(1u, (maxPoolRows + chunksPerOwner - 1) / chunksPerOwner);
    if (dynamic) {
        dp.tile_first = poolRows * dp.n_tiles_x;
        dp.tile_count -= std::min(dp.tile_count, poolRows * dp.n_tiles_x);
    }
    Scratch* s = nullptr;
    rc = acquire_scratch(sc, compact ? std::max<size_t>(size_t(myUnits) * blockPixels, 1) : peer ? size_t(1) : pixels, true, std::max<size_t>(othersPixels, 1), &s);
    if (rc) {
        release_scratch(sc, s);
        return rc;
    }
    ncclComm_t nc = static_cast<ncclComm_t>(comm->nccl);
    uint32_t launches = 0;
    // every rank renders its share as concurrent bands of its tile list when the share is large enough (launch_bands: the
    // stage tails do not shrink with the partition, so they weigh more the more GPUs share the frame)
    unsigned nBands = nBandsFor(sc, ls, p, dp, false);
    std::vector<Scratch*> bands;
    if (acquire_bands(sc, s, nBands, bands) != CGE_OK) {
        cudaGetLastError();
        nBands = 1;
    }
    if (nBands > 1)
        rc = use_band_stream(s, 0);
    if (rc != CGE_OK) {
        for (Scratch* b : bands)
            release_scratch(sc, b);
        return rc;
    }
    cudaEventRecord(s->ev0, s->stream);
    float* frame = ((rank == 0 && devOut) || peer) ? rgbOut : s->rgb;
    int* frameIds = wantIds ? ((rank == 0 && devOut && idsOut) ? idsOut : s->ids) : nullptr;
    if (nBands > 1) {
        std::vector<uint2> ranges;
        for (unsigned b = 0; b < nBands; b++) {
            const unsigned f0 = unsigned(uint64_t(dp.tile_count) * b / nBands), f1 = unsigned(uint64_t(dp.tile_count) * (b + 1) / nBands);
            ranges.push_back(make_uint2(dp.tile_first + f0, f1 - f0));
        }
        rc = launch_bands(sc, ls, bands, ranges, false, cam, &p, dp, frame, frameIds, &launches,
            [&](unsigned, Scratch* sb) { cudaEventRecord(sb->bandDone, sb->stream); });
    } else {
        rc = launch_render(sc, ls, s, cam, &p, dp, frame, frameIds, &launches);
    }
    std::vector<Scratch*> dyn; // one Scratch (queues, counters, grant) per dynamic pipeline; two streams, alternating
    if (rc == CGE_OK && dynamic) {
        unsigned* counter = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(rgbOut) + peer_counter_offset(comm->peer_frames[peerIdx].bytes));
        const unsigned base = comm->peer_grants[peerIdx];
        comm->peer_grants[peerIdx] += R * kGrants; // (unsigned wrap-around is fine: the kernel subtracts modulo 2^32)
        cudaEventRecord(s->copyDone, s->stream);   // the static part of this rank is queued: the dealt chunks follow it
        DevParams gp = dp;
        gp.tile_first = 0;
        gp.tile_count = chunkRows * dp.n_tiles_x; // sizes the launch; the grant says what is rendered
        cudaStream_t dynStream[2] = { nullptr, nullptr };
        for (unsigned g = 0; g < kGrants && rc == CGE_OK; g++) {
            Scratch* sd = nullptr;
            rc = acquire_scratch(sc, 1, false, 0, &sd);
            if (sd)
                dyn.push_back(sd);
            if (rc != CGE_OK)
                break;
            if (g < 2) {
                // a dealt chunk starts when this rank's own rows have left their shadow stage (the chunk's chain stage then runs beside
                // their shading and fold stages); frames of the single per-thread kernel: when that kernel has finished
                dynStream[g] = sd->baseStream;
                bool staged = true;
                for (Scratch* b : bands)
                    staged = staged && b->staged;
                if (staged)
                    for (Scratch* b : bands)
                        cudaStreamWaitEvent(dynStream[g], b->stage[2], 0);
                else
                    cudaStreamWaitEvent(dynStream[g], s->copyDone, 0);
            }
            sd->stream = dynStream[g % 2];
            if (!sd->grant)
                rc = cudaMalloc(&sd->grant, 16) == cudaSuccess ? CGE_OK : fail(CGE_ERR_CUDA, "grant buffer");
            if (rc != CGE_OK)
                break;
            grant_kernel<<<1, 1, 0, sd->stream>>>(counter, base, R, chunksPerOwner, chunkRows, poolPct, dp.n_tiles_x, dp.n_tiles_y, sd->grant);
            launches++;
            gp.grant = sd->grant;
            rc = launch_render(sc, ls, sd, cam, &p, gp, frame, nullptr, &launches);
        }
        for (unsigned g = 0; g < 2 && g < dyn.size(); g++) {
            cudaEventRecord(dyn[g]->bandDone, dynStream[g]); // (recorded after the last pipeline queued on that stream)
            cudaStreamWaitEvent(s->stream, dyn[g]->bandDone, 0);
        }
    }
    cudaEventRecord(s->ev1, s->stream);
    ncclResult_t nr = ncclSuccess;
    if (rc == CGE_OK && shared) {
        // ---- each rank's rows leave for the shared host frame over its own PCIe link: ONE strided copy -----------------------
        // block b of the compact buffer holds tile row u = rank + (myUnits - 1 - b) * R, frame rows [H - min(4u + 4, H), H - 4u);
        // only the frame's topmost tile row can be short (H not a multiple of 4): it is block 0 of the rank that owns it
        auto copyOut = [&](char* dst, const char* src, size_t px) { // px = bytes per pixel
            unsigned b0 = 0;
            const unsigned uTop = rank + (myUnits - 1) * R;
            const int topRows = p.height - int(uTop) * kTileH; // rows of this rank's topmost tile row
            if (myUnits && topRows < kTileH) {
                cudaMemcpyAsync(dst + size_t(p.height - int(uTop) * kTileH - topRows) * W * px, src, size_t(topRows) * W * px,
                    cudaMemcpyDeviceToHost, s->stream);
                b0 = 1;
            }
            if (myUnits > b0) {
                const unsigned uFirst = rank + (myUnits - 1 - b0) * R; // tile row of block b0
                const size_t top = size_t(p.height - int(uFirst + 1) * kTileH);
                cudaMemcpy2DAsync(dst + top * W * px, size_t(R) * blockPixels * px, src + size_t(b0) * blockPixels * px, blockPixels * px,
                    blockPixels * px, myUnits - b0, cudaMemcpyDeviceToHost, s->stream);
            }
        };
        copyOut(reinterpret_cast<char*>(rgbOut), reinterpret_cast<const char*>(s->rgb), 12);
        if (wantIds && idsOut)
            copyOut(reinterpret_cast<char*>(idsOut), reinterpret_cast<const char*>(s->ids), 4);
        // the frame is complete when every rank's copy has landed: a 4-byte all-reduce behind the copies on every stream
        nr = api->AllReduce(s->tileCounter, s->tileCounter, 1, ncclInt, ncclSum, nc, s->stream);
    } else if (rc == CGE_OK && peer) {
        // ---- the pixels are already in rank 0's frame: every rank's kernels stored them there over NVLink.  Kernel completion
        // makes a rank's stores visible; the 4-byte all-reduce behind the kernels on every stream completes on rank 0 only after
        // every rank has got there ------------------------------------------------------------------------------------------------
        nr = api->AllReduce(s->tileCounter, s->tileCounter, 1, ncclInt, ncclSum, nc, s->stream);
    } else if (rc == CGE_OK && R > 1) {
        // ---- gather on rank 0: the other ranks send their compact rows as they are, one unpack launch scatters all of them ----
        const size_t myPixels = size_t(myUnits) * blockPixels;
        if (rank != 0) {
            if (myPixels) {
                nr = api->GroupStart();
                if (nr == ncclSuccess)
                    nr = api->Send(s->rgb, myPixels * 3, ncclFloat, 0, nc, s->stream);
                if (nr == ncclSuccess && wantIds)
                    nr = api->Send(s->ids, myPixels, ncclInt, 0, nc, s->stream);
                if (nr == ncclSuccess)
                    nr = api->GroupEnd();
            }
        } else {
            nr = api->GroupStart();
            for (unsigned r = 1; r < R && nr == ncclSuccess; r++) {
                const size_t n = size_t(up.units[r]) * blockPixels, at = size_t(up.offset[r]) * blockPixels;
                if (!n)
                    continue;
                nr = api->Recv(s->gatherRgb + at * 3, n * 3, ncclFloat, int(r), nc, s->stream);
                if (nr == ncclSuccess && wantIds)
                    nr = api->Recv(s->gatherIds + at, n, ncclInt, int(r), nc, s->stream);
            }
            if (nr == ncclSuccess)
                nr = api->GroupEnd();
            if (nr == ncclSuccess && othersPixels) {
                up.n_ranks = R;
                unpack_rows_kernel<<<unsigned((othersPixels + 255) / 256), 256, 0, s->stream>>>(s->gatherRgb, wantIds ? s->gatherIds : nullptr,
                    frame, frameIds, p.width, p.height, up, othersPixels);
                launches++;
            }
        }
    }
    if (rc == CGE_OK && nr != ncclSuccess)
        rc = fail(CGE_ERR_NCCL, std::string("nccl gather: ") + api->GetErrorString(nr));
    // the bloom filter is the one cross-pixel step of the path: it runs on rank 0 once the gathered frame is complete
    if (rc == CGE_OK && rank == 0 && (p.features & CGE_FEAT_BLOOM_EFFECT))
        rc = apply_bloom(s, p, frame, &launches);
    if (rc == CGE_OK && rank == 0 && rgba8) {
        // the output stage of Screen::writeBitmapToFile on the gathered frame: the ids (if wanted) leave first, then their buffer
        // holds the packed pixels (4 bytes per pixel either way), as in cge_render
        if (wantIds && idsOut)
            cudaMemcpyAsync(idsOut, s->ids, pixels * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
        pack_rgba8_kernel<<<unsigned((pixels + 255) / 256), 256, 0, s->stream>>>(s->rgb, reinterpret_cast<uchar4*>(s->ids), pixels);
        launches++;
        cudaMemcpyAsync(rgbOut, s->ids, pixels * 4, cudaMemcpyDeviceToHost, s->stream);
    } else if (rc == CGE_OK && rank == 0 && !devOut && !shared) {
        cudaMemcpyAsync(rgbOut, s->rgb, pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
        if (wantIds && idsOut)
            cudaMemcpyAsync(idsOut, s->ids, pixels * sizeof(int), cudaMemcpyDeviceToHost, s->stream);
    }
    cudaEventRecord(s->ev2, s->stream);
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (rc == CGE_OK && e != cudaSuccess)
        rc = fail(CGE_ERR_CUDA, std::string("render_distributed: ") + cudaGetErrorString(e));
    for (unsigned b = 1; b < bands.size(); b++)
        cudaStreamSynchronize(bands[b]->stream);
    for (Scratch* sd : dyn) {
        cudaStreamSynchronize(sd->stream);
        sd->stream = sd->baseStream;
    }
    if (rc == CGE_OK) {
        std::vector<Scratch*> all = bands;
        all.insert(all.end(), dyn.begin(), dyn.end());
        rc = fill_stats(all, st, launches);
    }
    for (Scratch* sd : dyn)
        release_scratch(sc, sd);
    for (Scratch* b : bands)
        release_scratch(sc, b);
    return rc;
}

// pinned host buffers for callers that want the D2H copy at full PCIe speed
int cge_host_alloc(void** out, uint64_t bytes)
{
    if (!out)
        return fail(CGE_ERR_INVALID_ARG, "null argument");
    CGE_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return CGE_OK;
}
int cge_host_free(void* p)
{
    if (p)
        CGE_CUDA(cudaFreeHost(p));
    return CGE_OK;
}

} // extern "C"
