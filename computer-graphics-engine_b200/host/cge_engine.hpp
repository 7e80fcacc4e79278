// cge_engine.hpp — host-side C++ mirror of the reference engine's interface for the ray-tracing path.
//
// The reference's toolchain (C++20 + glm + its framework) is not part of this repository, so the layer above the C ABI
// is written here with the SAME names, argument meaning and error behaviour as the reference declarations it stands in
// for; a maintainer of the reference keeps their own headers and only swaps the body of renderRayTracing
// (INTEGRATION.md shows that shim).  Mirrored declarations (reference file:line):
//   Vertex / Material / Mesh                 framework/include/framework/mesh.h:14-43
//   Image                                    framework/include/framework/image.h:13-20
//   Ray                                      framework/include/framework/ray.h:9-13
//   HitInfo / Sphere / *Light / Features     src/common.h:14-77
//   Scene                                    src/scene.h:28-33
//   Screen                                   src/screen.h:10-33 (setPixel y flip src/screen.cpp:41-47)
//   Trackball                                framework/include/framework/trackball.h:15-62 (camera part only; the
//                                            reference couples it to a GLFW Window for the aspect ratio, here the
//                                            aspect ratio is a constructor argument)
//   BvhInterface                             src/bvh_interface.h:12-49
//   renderRayTracing / getFinalColor         src/render.h:32,35
//   segmentLightSamples / parallelogramLightDirectionSamples   src/light.h:9-10 (mutable globals in the reference)
// Everything computes on the GPU through include/cge.h; failures throw std::runtime_error carrying cge_last_error()
// (the reference itself signals load errors with `throw std::exception()`, framework/src/mesh.cpp:54-57).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "cge.h"
#include "cge_scene_file.h"

namespace cge_engine {

struct vec2 {
    float x = 0, y = 0;
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct ivec2 {
    int x = 0, y = 0;
};
struct uvec3 {
    unsigned x = 0, y = 0, z = 0;
};

struct Vertex {
    vec3 position, normal;
    vec2 texCoord;
};
struct Image {
    int width = 0, height = 0;
    std::vector<vec3> pixels;
};
struct Material {
    vec3 kd;
    vec3 ks { 0.0f };
    float shininess { 1.0f };
    float transparency { 1.0f };
    std::shared_ptr<Image> kdTexture;
};
struct Mesh {
    std::vector<Vertex> vertices;
    std::vector<uvec3> triangles;
    Material material;
};
struct Ray {
    vec3 origin { 0.0f };
    vec3 direction { 0.0f, 0.0f, -1.0f };
    float t { std::numeric_limits<float>::max() };
};
struct Sphere {
    vec3 center { 0.0f };
    float radius = 1.0f;
    Material material;
};
struct PointLight {
    vec3 position, color;
};
struct SegmentLight {
    vec3 endpoint0, endpoint1, color0, color1;
};
struct ParallelogramLight {
    vec3 v0, edge01, edge02, color0, color1, color2, color3;
};
struct ExtraFeatures {
    bool enableEnvironmentMapping = false, enableBvhSahBinning = false, enableMotionBlur = false, enableBloomEffect = false,
         enableBilinearTextureFiltering = false, enableMipmapTextureFiltering = false, enableMultipleRaysPerPixel = false,
         enableGlossyReflection = false, enableTransparency = false, enableDepthOfField = false;
};
struct Features {
    bool enableShading = false, enableRecursive = false, enableHardShadow = false, enableSoftShadow = false,
         enableNormalInterp = false, enableTextureMapping = false, enableAccelStructure = false;
    ExtraFeatures extra = {};
};
struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<std::variant<PointLight, SegmentLight, ParallelogramLight>> lights;
};

// the reference's tunables are mutable globals (src/light.cpp:12-13); kept as such on the host side, passed by value below
inline int segmentLightSamples = 25;
inline int parallelogramLightDirectionSamples = 5;
// likewise the globals of the two implemented ExtraFeatures (src/render.cpp:14,19-21; extern in src/render.h:10-13,24)
inline int raysPerPixelSide = 3;
inline float bloomScalar = .3f;
inline float bloomThreshold = .4f;
inline int bloomDebugOption = 0;
inline uint32_t samplerSeed = 0; // seed of the hash sampler that stands in for rand() / std::random_device (csrc/sampler.h)

inline void check(int rc, const char* what)
{
    if (rc != CGE_OK)
        throw std::runtime_error(std::string(what) + ": " + cge_last_error());
}

inline uint32_t featureBits(const Features& f)
{
    uint32_t b = 0;
    b |= f.enableShading ? CGE_FEAT_SHADING : 0;
    b |= f.enableRecursive ? CGE_FEAT_RECURSIVE : 0;
    b |= f.enableHardShadow ? CGE_FEAT_HARD_SHADOW : 0;
    b |= f.enableSoftShadow ? CGE_FEAT_SOFT_SHADOW : 0;
    b |= f.enableNormalInterp ? CGE_FEAT_NORMAL_INTERP : 0;
    b |= f.enableTextureMapping ? CGE_FEAT_TEXTURE_MAPPING : 0;
    b |= f.enableAccelStructure ? CGE_FEAT_ACCEL_STRUCTURE : 0;
    const bool* e = reinterpret_cast<const bool*>(&f.extra);
    for (int i = 0; i < 10; i++)
        b |= e[i] ? (1u << (16 + i)) : 0; // bloom and multiple rays per pixel are implemented; the library refuses the others
                                          // (CGE_ERR_UNSUPPORTED), never silently ignores them
    return b;
}

class Screen {
public:
    explicit Screen(const ivec2& resolution, bool /*presentable*/ = false)
        : m_resolution(resolution)
        , m_textureData(size_t(resolution.x) * size_t(resolution.y), vec3(0.0f))
    {
    }
    void clear(const vec3& color) { std::fill(m_textureData.begin(), m_textureData.end(), color); }
    void setPixel(int x, int y, const vec3& color) { m_textureData[size_t(indexAt(x, y))] = color; }
    [[nodiscard]] ivec2 resolution() const { return m_resolution; }
    [[nodiscard]] int indexAt(int x, int y) const { return (m_resolution.y - 1 - y) * m_resolution.x + x; }
    [[nodiscard]] const std::vector<vec3>& pixels() const { return m_textureData; }
    [[nodiscard]] std::vector<vec3>& pixels() { return m_textureData; }
    // void writeBitmapToFile(const std::filesystem::path&) (src/screen.cpp:49-60): clamp -> vec4(c, 1) * 255 -> u8vec4 on the
    // host, then the same file stb's stbi_write_bmp(path, w, h, 4, data) writes (writeBmpRgba below).
    inline void writeBitmapToFile(const std::string& filePath) const;

private:
    ivec2 m_resolution;
    std::vector<vec3> m_textureData;
};

// The file stb_image_write produces for comp = 4 (the reference's writer, framework/third_party/stb, stbi_write_bmp_core):
// 14-byte file header + 108-byte BITMAPV4HEADER (32 bpp, BI_BITFIELDS, masks R 00ff0000 G 0000ff00 B 000000ff A ff000000, all
// other fields zero), rows bottom-up, pixels as B G R A.  rgba: width * height * 4 bytes, row 0 = top (Screen::pixels() order).
inline void writeBmpRgba(const std::string& path, int width, int height, const uint8_t* rgba)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f)
        throw std::runtime_error("cannot write " + path);
    auto u16 = [&](uint32_t v) { const uint8_t b[2] = { uint8_t(v), uint8_t(v >> 8) }; std::fwrite(b, 1, 2, f); };
    auto u32 = [&](uint32_t v) { const uint8_t b[4] = { uint8_t(v), uint8_t(v >> 8), uint8_t(v >> 16), uint8_t(v >> 24) }; std::fwrite(b, 1, 4, f); };
    std::fputc('B', f), std::fputc('M', f);
    u32(14u + 108u + uint32_t(width) * uint32_t(height) * 4u), u16(0), u16(0), u32(14 + 108);
    u32(108), u32(uint32_t(width)), u32(uint32_t(height)), u16(1), u16(32), u32(3);
    for (int k = 0; k < 5; k++)
        u32(0); // image size, x / y pixels per metre, colours used, important colours
    u32(0x00ff0000u), u32(0x0000ff00u), u32(0x000000ffu), u32(0xff000000u);
    for (int k = 0; k < 13; k++)
        u32(0); // colour-space type, 9 endpoint words, 3 gamma words
    std::vector<uint8_t> row(size_t(width) * 4);
    for (int y = height - 1; y >= 0; y--) {
        const uint8_t* src = rgba + size_t(y) * size_t(width) * 4;
        for (int x = 0; x < width; x++) {
            row[4 * x + 0] = src[4 * x + 2], row[4 * x + 1] = src[4 * x + 1], row[4 * x + 2] = src[4 * x + 0], row[4 * x + 3] = src[4 * x + 3];
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
}

inline void Screen::writeBitmapToFile(const std::string& filePath) const
{
    std::vector<uint8_t> bytes(m_textureData.size() * 4);
    auto conv = [](float v) -> uint8_t { // glm::clamp = min(max(x, 0), 1) lets NaN through; NaN -> u8 gives 0 on x86
        if (v != v)
            return 0;
        v = std::fmin(std::fmax(v, 0.0f), 1.0f);
        return uint8_t(v * 255.0f);
    };
    for (size_t i = 0; i < m_textureData.size(); i++) {
        bytes[4 * i] = conv(m_textureData[i].x), bytes[4 * i + 1] = conv(m_textureData[i].y), bytes[4 * i + 2] = conv(m_textureData[i].z);
        bytes[4 * i + 3] = 255;
    }
    writeBmpRgba(filePath, m_resolution.x, m_resolution.y, bytes.data());
}

class Trackball {
public:
    // fovy in radians (as in the reference); aspect = width / height of the window the reference would query
    Trackball(float aspect, float fovy, float distanceFromLookAt = 4.0f, float rotationX = 0.0f, float rotationY = 0.0f)
        : m_aspect(aspect)
        , m_fovy(fovy)
        , m_lookAt(0.0f)
        , m_distanceFromLookAt(distanceFromLookAt)
        , m_rotationEulerAngles(rotationX, rotationY, 0.0f)
    {
    }
    void setCamera(const vec3 lookAt, const vec3 rotations, const float dist)
    {
        m_lookAt = lookAt;
        m_rotationEulerAngles = rotations;
        m_distanceFromLookAt = dist;
    }
    [[nodiscard]] cge_camera camera() const
    {
        cge_camera c;
        const float look[3] = { m_lookAt.x, m_lookAt.y, m_lookAt.z };
        const float rot[3] = { m_rotationEulerAngles.x, m_rotationEulerAngles.y, m_rotationEulerAngles.z };
        check(cge_camera_from_trackball(m_fovy, m_aspect, look, m_distanceFromLookAt, rot, &c), "cge_camera_from_trackball");
        return c;
    }
    [[nodiscard]] vec3 position() const
    {
        const cge_camera c = camera();
        return { c.origin[0], c.origin[1], c.origin[2] };
    }

private:
    float m_aspect, m_fovy;
    vec3 m_lookAt;
    float m_distanceFromLookAt;
    vec3 m_rotationEulerAngles;
};

// Flatten a Scene into the arrays of cge_scene_desc (what the reference-side shim does).
struct FlatScene {
    std::vector<cge_mesh_desc> meshes;
    std::vector<cge_vertex> vertices;
    std::vector<uint32_t> triangles;
    std::vector<cge_sphere_desc> spheres;
    std::vector<cge_light_desc> lights;
    std::vector<cge_texture_desc> textures;
    std::vector<float> texels;
    cge_scene_desc desc() const
    {
        cge_scene_desc d {};
        d.n_meshes = uint32_t(meshes.size()), d.n_vertices = uint32_t(vertices.size()), d.n_triangles = uint32_t(triangles.size() / 3);
        d.n_spheres = uint32_t(spheres.size()), d.n_lights = uint32_t(lights.size()), d.n_textures = uint32_t(textures.size());
        d.n_texels = texels.size() / 3;
        d.meshes = meshes.data(), d.vertices = vertices.data(), d.triangles = triangles.data(), d.spheres = spheres.data();
        d.lights = lights.data(), d.textures = textures.data(), d.texels = texels.data();
        return d;
    }
};

inline std::vector<cge_light_desc> flattenLights(const Scene& scene)
{
    std::vector<cge_light_desc> out;
    for (const auto& l : scene.lights) {
        cge_light_desc ld {};
        if (std::holds_alternative<PointLight>(l)) {
            ld.type = CGE_LIGHT_POINT;
            std::memcpy(ld.v, &std::get<PointLight>(l), sizeof(PointLight));
        } else if (std::holds_alternative<SegmentLight>(l)) {
            ld.type = CGE_LIGHT_SEGMENT;
            std::memcpy(ld.v, &std::get<SegmentLight>(l), sizeof(SegmentLight));
        } else {
            ld.type = CGE_LIGHT_PARALLELOGRAM;
            std::memcpy(ld.v, &std::get<ParallelogramLight>(l), sizeof(ParallelogramLight));
        }
        out.push_back(ld);
    }
    return out;
}

inline FlatScene flatten(const Scene& scene)
{
    static_assert(sizeof(Vertex) == sizeof(cge_vertex) && sizeof(PointLight) == 24 && sizeof(ParallelogramLight) == 84);
    FlatScene f;
    std::vector<const Image*> seen;
    auto texId = [&](const std::shared_ptr<Image>& img) -> int32_t {
        if (!img)
            return -1;
        for (size_t i = 0; i < seen.size(); i++)
            if (seen[i] == img.get())
                return int32_t(i);
        cge_texture_desc td {};
        td.width = img->width, td.height = img->height, td.texel_offset = f.texels.size() / 3;
        for (const auto& p : img->pixels)
            f.texels.insert(f.texels.end(), { p.x, p.y, p.z });
        f.textures.push_back(td);
        seen.push_back(img.get());
        return int32_t(seen.size() - 1);
    };
    for (const auto& mesh : scene.meshes) {
        cge_mesh_desc md {};
        md.vertex_offset = uint32_t(f.vertices.size()), md.vertex_count = uint32_t(mesh.vertices.size());
        md.triangle_offset = uint32_t(f.triangles.size() / 3), md.triangle_count = uint32_t(mesh.triangles.size());
        md.kd[0] = mesh.material.kd.x, md.kd[1] = mesh.material.kd.y, md.kd[2] = mesh.material.kd.z;
        md.ks[0] = mesh.material.ks.x, md.ks[1] = mesh.material.ks.y, md.ks[2] = mesh.material.ks.z;
        md.shininess = mesh.material.shininess, md.transparency = mesh.material.transparency;
        md.texture_id = texId(mesh.material.kdTexture);
        f.meshes.push_back(md);
        for (const auto& v : mesh.vertices) {
            cge_vertex cv;
            std::memcpy(&cv, &v, sizeof(cv));
            f.vertices.push_back(cv);
        }
        for (const auto& t : mesh.triangles)
            f.triangles.insert(f.triangles.end(), { t.x, t.y, t.z });
    }
    for (const auto& s : scene.spheres) {
        cge_sphere_desc sd {};
        sd.center[0] = s.center.x, sd.center[1] = s.center.y, sd.center[2] = s.center.z, sd.radius = s.radius;
        sd.kd[0] = s.material.kd.x, sd.kd[1] = s.material.kd.y, sd.kd[2] = s.material.kd.z;
        sd.ks[0] = s.material.ks.x, sd.ks[1] = s.material.ks.y, sd.ks[2] = s.material.ks.z;
        sd.shininess = s.material.shininess, sd.transparency = s.material.transparency, sd.texture_id = texId(s.material.kdTexture);
        f.spheres.push_back(sd);
    }
    f.lights = flattenLights(scene);
    return f;
}

// BvhInterface(Scene*, Features): builds the acceleration structure — here: flattens the scene, rebuilds the
// reference-order tree + the fast tree and uploads everything to HBM (cge_scene_create).
class BvhInterface {
public:
    BvhInterface(Scene* pScene, const Features& /*features*/, int device = 0)
    {
        const FlatScene f = flatten(*pScene);
        const cge_scene_desc d = f.desc();
        check(cge_scene_create(&d, device, &m_scene), "cge_scene_create");
    }
    ~BvhInterface() { cge_scene_destroy(m_scene); }
    BvhInterface(const BvhInterface&) = delete;
    BvhInterface& operator=(const BvhInterface&) = delete;
    [[nodiscard]] int numLevels() const
    {
        uint32_t n = 0;
        cge_scene_bvh_info(m_scene, nullptr, &n, nullptr, nullptr);
        return int(n);
    }
    [[nodiscard]] int numLeaves() const
    {
        uint32_t n = 0;
        cge_scene_bvh_info(m_scene, nullptr, nullptr, &n, nullptr);
        return int(n);
    }
    void setRecursionLevel(int) const {}      // debug-draw state in the reference; no effect on the image
    void setDebugRecursionLevel(int) const {}
    [[nodiscard]] cge_scene* handle() const { return m_scene; }

private:
    cge_scene* m_scene = nullptr;
};

inline cge_params makeParams(const ivec2& res, const Features& features, int rayDepth)
{
    cge_params p {};
    p.width = res.x, p.height = res.y;
    p.features = featureBits(features);
    p.ray_depth = rayDepth;
    p.segment_samples = segmentLightSamples;
    p.parallelogram_samples = parallelogramLightDirectionSamples;
    p.sampler = CGE_SAMPLER_HASH;
    p.seed = samplerSeed;
    p.traversal = CGE_TRAVERSAL_FAST;
    p.rays_per_pixel_side = raysPerPixelSide;
    p.bloom_scalar = bloomScalar;
    p.bloom_threshold = bloomThreshold;
    p.bloom_debug_option = bloomDebugOption;
    return p;
}

// void renderRayTracing(const Scene&, const Trackball&, const BvhInterface&, Screen&, const Features&)   (src/render.h:32)
// The light list is re-sent every call because the reference's GUI edits scene.lights between frames
// (src/main.cpp:290-368); geometry lives in the BvhInterface.  rayDepth: the reference passes the literal 5.
inline void renderRayTracing(const Scene& scene, const Trackball& camera, const BvhInterface& bvh, Screen& screen, const Features& features,
    int rayDepth = 5, cge_stats* stats = nullptr)
{
    const std::vector<cge_light_desc> lights = flattenLights(scene);
    check(cge_scene_update_lights(bvh.handle(), lights.data(), uint32_t(lights.size())), "cge_scene_update_lights");
    const cge_camera cam = camera.camera();
    const cge_params p = makeParams(screen.resolution(), features, rayDepth);
    static_assert(sizeof(vec3) == 12);
    check(cge_render(bvh.handle(), &cam, &p, reinterpret_cast<float*>(screen.pixels().data()), nullptr, stats), "cge_render");
}

// renderRayTracing followed by Screen::writeBitmapToFile as the CLI does (src/main.cpp:520-524), with the clamp -> u8x4
// conversion done on the GPU before the read-back (CGE_FLAG_OUTPUT_RGBA8: 4 instead of 12 bytes per pixel cross PCIe).
inline void renderRayTracingToBitmap(const Scene& scene, const Trackball& camera, const BvhInterface& bvh, const ivec2& resolution,
    const Features& features, const std::string& filePath, int rayDepth = 5, cge_stats* stats = nullptr)
{
    const std::vector<cge_light_desc> lights = flattenLights(scene);
    check(cge_scene_update_lights(bvh.handle(), lights.data(), uint32_t(lights.size())), "cge_scene_update_lights");
    const cge_camera cam = camera.camera();
    cge_params p = makeParams(resolution, features, rayDepth);
    p.flags |= CGE_FLAG_OUTPUT_RGBA8;
    std::vector<uint8_t> rgba(size_t(resolution.x) * size_t(resolution.y) * 4);
    check(cge_render(bvh.handle(), &cam, &p, reinterpret_cast<float*>(rgba.data()), nullptr, stats), "cge_render");
    writeBmpRgba(filePath, resolution.x, resolution.y, rgba.data());
}

// glm::vec3 getFinalColor(const Scene&, const BvhInterface&, Ray, const Features&, int rayDepth = 0)     (src/render.h:35)
inline vec3 getFinalColor(const Scene& scene, const BvhInterface& bvh, Ray ray, const Features& features, int rayDepth = 0)
{
    const std::vector<cge_light_desc> lights = flattenLights(scene);
    check(cge_scene_update_lights(bvh.handle(), lights.data(), uint32_t(lights.size())), "cge_scene_update_lights");
    const cge_params p = makeParams({ 1, 1 }, features, rayDepth);
    const float r7[7] = { ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z, ray.t };
    vec3 out;
    check(cge_trace_rays(bvh.handle(), r7, 1, &p, &out.x, nullptr), "cge_trace_rays");
    return out;
}

// Scene from a flat scene file (stands in for loadScenePrebuilt / loadSceneFromFile, src/scene.cpp:5-103, whose OBJ /
// PNG parsing is host I/O outside the hot path).
inline Scene loadFlatScene(const std::string& path)
{
    // the file is user input: it is closed on every path, every section must be there in full, and every offset / count / id that
    // is used as an index below is checked against the header's counts first
    std::unique_ptr<FILE, int (*)(FILE*)> file(std::fopen(path.c_str(), "rb"), &std::fclose);
    FILE* fp = file.get();
    if (!fp)
        throw std::runtime_error("File " + path + " does not exist.");
    cge_scene_file_header h;
    if (std::fread(&h, sizeof(h), 1, fp) != 1 || std::memcmp(h.magic, CGE_SCENE_FILE_MAGIC, 8) != 0)
        throw std::runtime_error("Failed to load scene " + path + ": not a flat scene file");
    std::fseek(fp, 0, SEEK_END);
    const size_t fileBytes = size_t(std::ftell(fp));
    std::fseek(fp, long(sizeof(h)), SEEK_SET);
    auto rd = [&](auto& vec, size_t count) {
        if (count > fileBytes / sizeof(vec[0])) // (before resize: a corrupt count must not become a huge allocation)
            throw std::runtime_error("Failed to load scene " + path + ": truncated scene file");
        vec.resize(count);
        if (count && std::fread(vec.data(), sizeof(vec[0]), count, fp) != count)
            throw std::runtime_error("Failed to load scene " + path + ": truncated scene file");
    };
    std::vector<cge_mesh_desc> meshes;
    std::vector<cge_vertex> vertices;
    std::vector<uint32_t> tris;
    std::vector<cge_sphere_desc> spheres;
    std::vector<cge_light_desc> lights;
    std::vector<cge_texture_desc> textures;
    std::vector<float> texels;
    rd(meshes, h.n_meshes), rd(vertices, h.n_vertices), rd(tris, size_t(h.n_triangles) * 3), rd(spheres, h.n_spheres);
    rd(lights, h.n_lights), rd(textures, h.n_textures), rd(texels, size_t(h.n_texels) * 3);
    auto bad = [&](const char* what) { return std::runtime_error("Failed to load scene " + path + ": " + what + " out of range"); };
    for (const auto& td : textures)
        if (td.width < 0 || td.height < 0 || td.texel_offset > h.n_texels || uint64_t(td.width) * uint64_t(td.height) > h.n_texels - td.texel_offset)
            throw bad("texture");
    for (const auto& md : meshes) {
        if (md.vertex_offset > h.n_vertices || md.vertex_count > h.n_vertices - md.vertex_offset)
            throw bad("mesh vertices");
        if (md.triangle_offset > h.n_triangles || md.triangle_count > h.n_triangles - md.triangle_offset)
            throw bad("mesh triangles");
        if (md.texture_id >= int32_t(h.n_textures))
            throw bad("mesh texture id");
        for (uint32_t t = 0; t < md.triangle_count * 3; t++)
            if (tris[3 * size_t(md.triangle_offset) + t] >= md.vertex_count)
                throw bad("triangle index");
    }
    for (const auto& sd : spheres)
        if (sd.texture_id >= int32_t(h.n_textures))
            throw bad("sphere texture id");
    for (const auto& ld : lights)
        if (ld.type != CGE_LIGHT_POINT && ld.type != CGE_LIGHT_SEGMENT && ld.type != CGE_LIGHT_PARALLELOGRAM)
            throw bad("light type");
    Scene sc;
    std::vector<std::shared_ptr<Image>> images;
    for (const auto& td : textures) {
        auto img = std::make_shared<Image>();
        img->width = td.width, img->height = td.height;
        img->pixels.resize(size_t(td.width) * size_t(td.height));
        std::memcpy(static_cast<void*>(img->pixels.data()), texels.data() + td.texel_offset * 3, img->pixels.size() * sizeof(vec3));
        images.push_back(img);
    }
    auto mat = [&](const float* kd, const float* ks, float sh, float tr, int32_t tex) {
        Material m;
        m.kd = { kd[0], kd[1], kd[2] }, m.ks = { ks[0], ks[1], ks[2] }, m.shininess = sh, m.transparency = tr;
        if (tex >= 0)
            m.kdTexture = images[size_t(tex)];
        return m;
    };
    for (const auto& md : meshes) {
        Mesh m;
        m.vertices.resize(md.vertex_count);
        std::memcpy(static_cast<void*>(m.vertices.data()), vertices.data() + md.vertex_offset, size_t(md.vertex_count) * sizeof(Vertex));
        for (uint32_t t = 0; t < md.triangle_count; t++) {
            const uint32_t* q = tris.data() + 3 * size_t(md.triangle_offset + t);
            m.triangles.push_back({ q[0], q[1], q[2] });
        }
        m.material = mat(md.kd, md.ks, md.shininess, md.transparency, md.texture_id);
        sc.meshes.push_back(std::move(m));
    }
    for (const auto& sd : spheres)
        sc.spheres.push_back({ { sd.center[0], sd.center[1], sd.center[2] }, sd.radius, mat(sd.kd, sd.ks, sd.shininess, sd.transparency, sd.texture_id) });
    for (const auto& ld : lights) {
        if (ld.type == CGE_LIGHT_POINT) {
            PointLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            sc.lights.emplace_back(l);
        } else if (ld.type == CGE_LIGHT_SEGMENT) {
            SegmentLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            sc.lights.emplace_back(l);
        } else {
            ParallelogramLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            sc.lights.emplace_back(l);
        }
    }
    return sc;
}

} // namespace cge_engine
