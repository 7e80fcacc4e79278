// render_gpu.cpp — the reference-side binding of libcge.so: the file a maintainer ADDS to the reference engine
// (src/render_gpu.cpp in its tree) to put the GPU path behind the engine's own seam
//
//     void renderRayTracing(const Scene&, const Trackball&, const BvhInterface&, Screen&, const Features&);   // src/render.h:32
//
// It compiles against the reference's UNMODIFIED headers (src/render.h, scene.h, screen.h, light.h, common.h,
// framework/trackball.h, framework/window.h) plus include/cge.h, and is built and run by this repo's tests exactly that way
// (oracle/ref/build_ref.sh -> oracle/_ref/libcge_ref_gpu.so, tests/test_gpu_reference_shim.py): the reference's own Scene,
// Trackball and Screen objects go in, the image in Screen::pixels() comes from the GPU.
//
// The former body of renderRayTracing (src/render.cpp:273-329) is kept as renderRayTracingCPU and is what this function falls
// back to when libcge.so refuses the call (no CUDA device, an ExtraFeatures flag it does not implement, transparency with
// recursion): the fallback is the reference's own code on the reference's side of the boundary.  In the reference's build that
// is a one-word rename in src/render.cpp; the test build does the rename on a copy of the object file (objcopy
// --redefine-sym), so no reference source is edited.
#include "render.h"
#include "light.h"   // segmentLightSamples, parallelogramLightDirectionSamples (src/light.cpp:12-13)
#include "scene.h"
#include "screen.h"
#include <framework/trackball.h>
#include <framework/window.h>

#include <cge.h>

#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

void renderRayTracingCPU(const Scene& scene, const Trackball& camera, const BvhInterface& bvh, Screen& screen, const Features& features);

namespace {

// Trackball keeps the field of view and its Window private (framework/include/framework/trackball.h:49-54) and the header is
// not ours to edit.  Explicit template instantiation may name private members, which gives read access without touching it.
// A maintainer would rather add `float fovy() const` and `float aspectRatio() const`.
template <typename Tag, typename Tag::type Member>
struct PrivateMember {
    friend typename Tag::type memberOf(Tag) { return Member; }
};
struct FovyTag {
    using type = float Trackball::*;
    friend type memberOf(FovyTag);
};
struct WindowTag {
    using type = const Window* Trackball::*;
    friend type memberOf(WindowTag);
};
template struct PrivateMember<FovyTag, &Trackball::m_fovy>;
template struct PrivateMember<WindowTag, &Trackball::m_pWindow>;

uint32_t featureBits(const Features& f)
{
    uint32_t b = 0;
    b |= f.enableShading ? CGE_FEAT_SHADING : 0;
    b |= f.enableRecursive ? CGE_FEAT_RECURSIVE : 0;
    b |= f.enableHardShadow ? CGE_FEAT_HARD_SHADOW : 0;
    b |= f.enableSoftShadow ? CGE_FEAT_SOFT_SHADOW : 0;
    b |= f.enableNormalInterp ? CGE_FEAT_NORMAL_INTERP : 0;
    b |= f.enableTextureMapping ? CGE_FEAT_TEXTURE_MAPPING : 0;
    b |= f.enableAccelStructure ? CGE_FEAT_ACCEL_STRUCTURE : 0;
    // ExtraFeatures (src/common.h:54-65) in declaration order from bit 16: bloom (19) and multiple rays per pixel (22) run on
    // the GPU, any other one makes cge_render answer CGE_ERR_UNSUPPORTED and the frame takes the CPU path below
    const bool extra[10] = { f.extra.enableEnvironmentMapping, f.extra.enableBvhSahBinning, f.extra.enableMotionBlur,
        f.extra.enableBloomEffect, f.extra.enableBilinearTextureFiltering, f.extra.enableMipmapTextureFiltering,
        f.extra.enableMultipleRaysPerPixel, f.extra.enableGlossyReflection, f.extra.enableTransparency, f.extra.enableDepthOfField };
    for (int i = 0; i < 10; i++)
        b |= extra[i] ? (1u << (16 + i)) : 0u;
    return b;
}

// std::variant<PointLight, SegmentLight, ParallelogramLight> (src/scene.h:32) -> tagged POD, members in declaration order
std::vector<cge_light_desc> lightsOf(const Scene& scene)
{
    static_assert(sizeof(PointLight) == 24 && sizeof(SegmentLight) == 48 && sizeof(ParallelogramLight) == 84, "light layouts");
    std::vector<cge_light_desc> out;
    for (const auto& l : scene.lights) {
        cge_light_desc d {};
        if (const auto* p = std::get_if<PointLight>(&l)) {
            d.type = CGE_LIGHT_POINT;
            std::memcpy(d.v, p, sizeof *p);
        } else if (const auto* s = std::get_if<SegmentLight>(&l)) {
            d.type = CGE_LIGHT_SEGMENT;
            std::memcpy(d.v, s, sizeof *s);
        } else {
            const auto* q = std::get_if<ParallelogramLight>(&l);
            d.type = CGE_LIGHT_PARALLELOGRAM;
            std::memcpy(d.v, q, sizeof *q);
        }
        out.push_back(d);
    }
    return out;
}

void putMaterial(const Material& m, float kd[3], float ks[3], float& shininess, float& transparency)
{
    std::memcpy(kd, &m.kd.x, 12);
    std::memcpy(ks, &m.ks.x, 12);
    shininess = m.shininess;
    transparency = m.transparency;
}

// One device scene per Scene object, rebuilt when its geometry changes - the reference rebuilds its BvhInterface in the same
// situations (src/main.cpp:181-187: scene switch).  The light list is NOT part of the key: the GUI edits it every frame
// (src/main.cpp:290-368) and it is re-sent with cge_scene_update_lights before every frame.
struct Entry {
    cge_scene* handle = nullptr;
    size_t nVertices = 0, nTriangles = 0, nSpheres = 0;
    std::vector<cge_light_desc> lights; // the list last uploaded: unchanged lists are not uploaded again
};
std::mutex g_mu;
std::unordered_map<const Scene*, Entry> g_scenes;

Entry* deviceScene(const Scene& scene)
{
    static_assert(sizeof(Vertex) == sizeof(cge_vertex), "Vertex is the 32-byte record cge_vertex mirrors");
    size_t nV = 0, nT = 0;
    for (const Mesh& m : scene.meshes)
        nV += m.vertices.size(), nT += m.triangles.size();
    if (auto it = g_scenes.find(&scene); it != g_scenes.end()) {
        Entry& e = it->second;
        if (e.nVertices == nV && e.nTriangles == nT && e.nSpheres == scene.spheres.size())
            return &e;
        cge_scene_destroy(e.handle);
        g_scenes.erase(it);
    }
    std::vector<cge_mesh_desc> meshes;
    std::vector<cge_vertex> vertices;
    std::vector<uint32_t> triangles;
    std::vector<cge_sphere_desc> spheres;
    std::vector<cge_texture_desc> textures;
    std::vector<float> texels;
    std::vector<const Image*> seen;
    auto textureId = [&](const std::shared_ptr<Image>& img) -> int32_t { // Material::kdTexture (framework/mesh.h:33)
        if (!img)
            return -1;
        for (size_t i = 0; i < seen.size(); i++)
            if (seen[i] == img.get())
                return int32_t(i);
        cge_texture_desc td {};
        td.width = img->width, td.height = img->height, td.texel_offset = texels.size() / 3;
        const size_t at = texels.size();
        texels.resize(at + img->pixels.size() * 3);
        std::memcpy(texels.data() + at, img->pixels.data(), img->pixels.size() * 12);
        textures.push_back(td);
        seen.push_back(img.get());
        return int32_t(seen.size() - 1);
    };
    vertices.reserve(nV), triangles.reserve(nT * 3);
    for (const Mesh& m : scene.meshes) {
        cge_mesh_desc md {};
        md.vertex_offset = uint32_t(vertices.size()), md.vertex_count = uint32_t(m.vertices.size());
        md.triangle_offset = uint32_t(triangles.size() / 3), md.triangle_count = uint32_t(m.triangles.size());
        putMaterial(m.material, md.kd, md.ks, md.shininess, md.transparency);
        md.texture_id = textureId(m.material.kdTexture);
        meshes.push_back(md);
        const size_t v0 = vertices.size();
        vertices.resize(v0 + m.vertices.size());
        std::memcpy(vertices.data() + v0, m.vertices.data(), m.vertices.size() * sizeof(Vertex));
        for (const glm::uvec3& t : m.triangles)
            triangles.insert(triangles.end(), { t.x, t.y, t.z });
    }
    for (const Sphere& s : scene.spheres) {
        cge_sphere_desc sd {};
        std::memcpy(sd.center, &s.center.x, 12);
        sd.radius = s.radius;
        putMaterial(s.material, sd.kd, sd.ks, sd.shininess, sd.transparency);
        sd.texture_id = textureId(s.material.kdTexture);
        spheres.push_back(sd);
    }
    Entry e;
    e.lights = lightsOf(scene);
    cge_scene_desc d {};
    d.n_meshes = uint32_t(meshes.size()), d.n_vertices = uint32_t(vertices.size()), d.n_triangles = uint32_t(triangles.size() / 3);
    d.n_spheres = uint32_t(spheres.size()), d.n_lights = uint32_t(e.lights.size()), d.n_textures = uint32_t(textures.size());
    d.n_texels = texels.size() / 3;
    d.meshes = meshes.data(), d.vertices = vertices.data(), d.triangles = triangles.data(), d.spheres = spheres.data();
    d.lights = e.lights.data(), d.textures = textures.data(), d.texels = texels.data();
    if (cge_scene_create(&d, /*device*/ 0, &e.handle) != CGE_OK)
        return nullptr;
    e.nVertices = nV, e.nTriangles = nT, e.nSpheres = scene.spheres.size();
    return &(g_scenes[&scene] = std::move(e));
}

uint32_t g_seed = 0; // seed of the stateless sampler that stands in for rand() (csrc/sampler.h)

} // namespace

// the sampler seed is the one knob the reference has no global for (it calls the process-wide rand(), src/light.cpp:21,32-33)
extern "C" void cgeSetSamplerSeed(uint32_t seed) { g_seed = seed; }

void renderRayTracing(const Scene& scene, const Trackball& camera, const BvhInterface& bvh, Screen& screen, const Features& features)
{
    cge_params p {};
    p.width = screen.resolution().x, p.height = screen.resolution().y;
    p.features = featureBits(features);
    p.ray_depth = 5; // the literal at src/render.cpp:298,308,318
    p.segment_samples = segmentLightSamples;
    p.parallelogram_samples = parallelogramLightDirectionSamples;
    p.sampler = CGE_SAMPLER_HASH, p.seed = g_seed, p.traversal = CGE_TRAVERSAL_FAST;
    p.rays_per_pixel_side = raysPerPixelSide;                                                         // src/render.cpp:14
    p.bloom_scalar = bloomScalar, p.bloom_threshold = bloomThreshold, p.bloom_debug_option = bloomDebugOption; // :19-21
    const glm::vec3 look = camera.lookAt(), rot = camera.rotationEulerAngles();
    const float fovy = camera.*memberOf(FovyTag {});
    const float aspect = (camera.*memberOf(WindowTag {}))->getAspectRatio();
    cge_camera cam {};
    bool ok = cge_camera_from_trackball(fovy, aspect, &look.x, camera.distanceFromLookAt(), &rot.x, &cam) == CGE_OK;
    cge_scene* dev = nullptr;
    if (ok) {
        // the map and the light upload are serialised; the frames themselves run concurrently (cge_render is re-entrant on one
        // scene, as src/main.cpp:514-528 needs: one thread per camera)
        std::lock_guard<std::mutex> lk(g_mu);
        Entry* e = deviceScene(scene);
        ok = e != nullptr;
        if (ok) {
            dev = e->handle;
            std::vector<cge_light_desc> lights = lightsOf(scene);
            const bool same = lights.size() == e->lights.size()
                && (lights.empty() || std::memcmp(lights.data(), e->lights.data(), lights.size() * sizeof(cge_light_desc)) == 0);
            if (!same) {
                ok = cge_scene_update_lights(dev, lights.data(), uint32_t(lights.size())) == CGE_OK;
                e->lights = std::move(lights);
            }
        }
    }
    // Screen::pixels() is W*H packed glm::vec3, row 0 = top (src/screen.cpp:41-47,119-122): exactly cge_render's output layout
    if (ok)
        ok = cge_render(dev, &cam, &p, &screen.pixels()[0].x, nullptr, nullptr) == CGE_OK;
    if (!ok)
        renderRayTracingCPU(scene, camera, bvh, screen, features);
}
