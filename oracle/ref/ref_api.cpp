// TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
//
// extern "C" harness around the UNMODIFIED reference engine (sources compiled where they lie under
// /root/reference by oracle/ref/build_ref.sh, linked with prebuilt/libIntersect_linux_x64.a).  It gives the
// tests, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs:
//   * the six libIntersect functions (src/intersect.h:5-16) as batch known-answer generators,
//   * the reference Scene / BvhInterface built from a flat scene file (include/cge_scene_file.h),
//   * the reference renderer: either renderRayTracing itself (src/render.cpp:273, depth literal 5) or the
//     identical pixel loop calling getFinalColor(scene,bvh,ray,features,depth) (src/render.h:35) so that
//     the depth can be chosen, plus primary-hit primitive ids and ray/box/triangle counters obtained with
//     `ld --wrap` (no reference source is edited),
//   * scene export: the reference's own loaders -> flat scene files (only where /root/reference exists).
//
// Soft-shadow sampling: light.cpp calls the global rand() (src/light.cpp:21,32-33).  With --wrap=rand the
// harness substitutes hash(seed, pixel, per-pixel draw counter) (sampler mode 1) — order independent across
// pixels, so OpenMP stays deterministic and the CUDA path can reproduce it; mode 0 is glibc rand().
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <span>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>
#include <omp.h>

#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include <framework/mesh.h>
#include <framework/ray.h>
#include <framework/image.h>
#include <framework/window.h>

// Read-only access to private members (BVH nodes / primitive order, camera constants) for EXPORT only.
// Layout is unaffected by access specifiers; no reference file is modified.
#define private public
#include <framework/trackball.h>
#include "bounding_volume_hierarchy.h"
#include "bvh_interface.h"
#undef private

#include "common.h"
#include "intersect.h"
#include "interpolate.h"
#include "light.h"
#include "render.h"
#include "scene.h"
#include "screen.h"
#include "shading.h"
#include "texture.h"

#include "cge_scene_file.h"

// defined in src/render.cpp:158 with external linkage, not declared in render.h
void renderBloomFilter(Screen& screen, const Features& features);

// ------------------------------------------------------------------------------------------------------
// ld --wrap interposers
// ------------------------------------------------------------------------------------------------------
namespace {
struct Tls {
    uint64_t rays = 0, boxes = 0, tris = 0, spheres = 0;
    const glm::vec3* lastV[3] = { nullptr, nullptr, nullptr };
    const Sphere* lastSphere = nullptr;
    uint32_t pixel = 0;
    uint32_t counter = 0;
};
thread_local Tls tls;
int g_samplerMode = 0; // 0 = glibc rand, 1 = hash
uint32_t g_seed = 0;
std::atomic<uint64_t> g_rays { 0 }, g_boxes { 0 }, g_tris { 0 }, g_spheres { 0 };

inline uint32_t hashSample(uint32_t seed, uint32_t pixel, uint32_t counter)
{
    // Must stay in sync with cge_hash_sample in computer-graphics-engine_b200/csrc/sampler.h and
    // oracle/cge_oracle.cpp (documented in DESIGN.md "sampler").
    uint32_t h = seed ^ (pixel * 0x9E3779B1u);
    h ^= counter * 0x85EBCA77u;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h >> 1; // 31 bits: 0 .. RAND_MAX
}
} // namespace

extern "C" {
int __real_rand(void);
int __wrap_rand(void)
{
    if (g_samplerMode == 0)
        return __real_rand();
    return int(hashSample(g_seed, tls.pixel, tls.counter++));
}
}

// getRaySamples (src/render.cpp:212-214) seeds a fresh std::mt19937 per pixel from std::random_device, which is
// non-deterministic.  The only out-of-line piece is std::random_device::_M_getval() (libstdc++.so): wrapped, it returns a
// hash of (seed, pixel), so the reference's own mt19937 + uniform_real_distribution run on a reproducible seed.
// Must stay in sync with cge_aa_seed in computer-graphics-engine_b200/csrc/sampler.h.
extern "C" unsigned int __real__ZNSt13random_device9_M_getvalEv(void*);
extern "C" unsigned int __wrap__ZNSt13random_device9_M_getvalEv(void* self)
{
    if (g_samplerMode == 0)
        return __real__ZNSt13random_device9_M_getvalEv(self);
    return hashSample(g_seed ^ 0x52444556u, tls.pixel, 0u);
}

#ifndef CGE_REF_NO_COUNTERS
// Mangled names: see SURVEY.md Appendix B.
extern "C" bool __real__Z24intersectRayWithTriangleRKN3glm3vecILi3EfLNS_9qualifierE0EEES4_S4_R3RayR7HitInfo(
    const glm::vec3&, const glm::vec3&, const glm::vec3&, Ray&, HitInfo&);
extern "C" bool __wrap__Z24intersectRayWithTriangleRKN3glm3vecILi3EfLNS_9qualifierE0EEES4_S4_R3RayR7HitInfo(
    const glm::vec3& v0, const glm::vec3& v1, const glm::vec3& v2, Ray& ray, HitInfo& hitInfo)
{
    tls.tris++;
    bool hit = __real__Z24intersectRayWithTriangleRKN3glm3vecILi3EfLNS_9qualifierE0EEES4_S4_R3RayR7HitInfo(v0, v1, v2, ray, hitInfo);
    if (hit) {
        tls.lastV[0] = &v0;
        tls.lastV[1] = &v1;
        tls.lastV[2] = &v2;
        tls.lastSphere = nullptr;
    }
    return hit;
}
extern "C" bool __real__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(const AxisAlignedBox&, Ray&);
extern "C" bool __wrap__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(const AxisAlignedBox& box, Ray& ray)
{
    tls.boxes++;
    return __real__Z21intersectRayWithShapeRK14AxisAlignedBoxR3Ray(box, ray);
}
extern "C" bool __real__Z21intersectRayWithShapeRK6SphereR3RayR7HitInfo(const Sphere&, Ray&, HitInfo&);
extern "C" bool __wrap__Z21intersectRayWithShapeRK6SphereR3RayR7HitInfo(const Sphere& s, Ray& ray, HitInfo& hitInfo)
{
    tls.spheres++;
    bool hit = __real__Z21intersectRayWithShapeRK6SphereR3RayR7HitInfo(s, ray, hitInfo);
    if (hit) {
        tls.lastSphere = &s;
        tls.lastV[0] = nullptr;
    }
    return hit;
}
extern "C" bool __real__ZNK12BvhInterface9intersectER3RayR7HitInfoRK8Features(const BvhInterface*, Ray&, HitInfo&, const Features&);
extern "C" bool __wrap__ZNK12BvhInterface9intersectER3RayR7HitInfoRK8Features(const BvhInterface* self, Ray& ray, HitInfo& hitInfo, const Features& f)
{
    tls.rays++;
    return __real__ZNK12BvhInterface9intersectER3RayR7HitInfoRK8Features(self, ray, hitInfo, f);
}
#endif

// ------------------------------------------------------------------------------------------------------
// Scene handle
// ------------------------------------------------------------------------------------------------------
namespace {
struct RefScene {
    Scene scene;
    std::vector<size_t> meshTriBase; // global primitive id of triangle 0 of each mesh
    size_t nTriangles = 0;
    // (mesh, i0, i1, i2) -> global triangle id (last one wins for duplicate triangles, matching visit order is
    // impossible to know here; duplicates do not occur in the fixtures)
    std::vector<std::map<std::array<uint32_t, 3>, uint32_t>> triLookup;
};
struct RefBvh {
    std::unique_ptr<BvhInterface> bvh;
    Features features;
};

Features featuresFromBits(uint32_t bits)
{
    Features f;
    f.enableShading = bits & CGE_FEAT_SHADING;
    f.enableRecursive = bits & CGE_FEAT_RECURSIVE;
    f.enableHardShadow = bits & CGE_FEAT_HARD_SHADOW;
    f.enableSoftShadow = bits & CGE_FEAT_SOFT_SHADOW;
    f.enableNormalInterp = bits & CGE_FEAT_NORMAL_INTERP;
    f.enableTextureMapping = bits & CGE_FEAT_TEXTURE_MAPPING;
    f.enableAccelStructure = bits & CGE_FEAT_ACCEL_STRUCTURE;
    // ExtraFeatures are outside the hot-path scope but the harness can still set them (bits 16..25) so that
    // a later round can pin them too.
    bool* extra = reinterpret_cast<bool*>(&f.extra);
    for (int i = 0; i < 10; i++)
        extra[i] = (bits >> (16 + i)) & 1u;
    return f;
}

// Image has no default constructor (framework/include/framework/image.h:13-20): build one in raw storage.
std::shared_ptr<Image> makeImage(int w, int h, const float* rgb)
{
    void* raw = ::operator new(sizeof(Image));
    Image* img = static_cast<Image*>(raw);
    img->width = w;
    img->height = h;
    new (&img->pixels) std::vector<glm::vec3>();
    img->pixels.resize(size_t(w) * size_t(h));
    std::memcpy(img->pixels.data(), rgb, size_t(w) * size_t(h) * 3 * sizeof(float));
    return std::shared_ptr<Image>(img, [](Image* p) {
        p->pixels.~vector();
        ::operator delete(static_cast<void*>(p));
    });
}

void finishScene(RefScene& rs)
{
    rs.meshTriBase.clear();
    rs.triLookup.clear();
    size_t base = 0;
    for (const auto& mesh : rs.scene.meshes) {
        rs.meshTriBase.push_back(base);
        std::map<std::array<uint32_t, 3>, uint32_t> lk;
        for (size_t t = 0; t < mesh.triangles.size(); t++) {
            const auto& tri = mesh.triangles[t];
            lk[{ tri.x, tri.y, tri.z }] = uint32_t(base + t);
        }
        rs.triLookup.push_back(std::move(lk));
        base += mesh.triangles.size();
    }
    rs.nTriangles = base;
}

int32_t lastPrimitiveId(const RefScene& rs)
{
    if (tls.lastSphere) {
        return int32_t(rs.nTriangles + size_t(tls.lastSphere - rs.scene.spheres.data()));
    }
    if (!tls.lastV[0])
        return -1;
    for (size_t m = 0; m < rs.scene.meshes.size(); m++) {
        const auto& verts = rs.scene.meshes[m].vertices;
        if (verts.empty())
            continue;
        const char* lo = reinterpret_cast<const char*>(verts.data());
        const char* hi = lo + verts.size() * sizeof(Vertex);
        const char* p = reinterpret_cast<const char*>(tls.lastV[0]);
        if (p >= lo && p < hi) {
            std::array<uint32_t, 3> key;
            for (int k = 0; k < 3; k++)
                key[k] = uint32_t((reinterpret_cast<const char*>(tls.lastV[k]) - lo) / sizeof(Vertex));
            auto it = rs.triLookup[m].find(key);
            return it == rs.triLookup[m].end() ? -2 : int32_t(it->second);
        }
    }
    return -2;
}

bool writeFlat(const RefScene& rs, const BoundingVolumeHierarchy* bvh, const char* path)
{
    const Scene& sc = rs.scene;
    cge_scene_file_header h {};
    std::memcpy(h.magic, CGE_SCENE_FILE_MAGIC, 8);
    std::vector<cge_mesh_desc> meshes;
    std::vector<cge_vertex> vertices;
    std::vector<uint32_t> tris;
    std::vector<cge_sphere_desc> spheres;
    std::vector<cge_light_desc> lights;
    std::vector<cge_texture_desc> textures;
    std::vector<float> texels;
    std::map<const Image*, int32_t> texIds;
    auto texId = [&](const std::shared_ptr<Image>& img) -> int32_t {
        if (!img)
            return -1;
        auto it = texIds.find(img.get());
        if (it != texIds.end())
            return it->second;
        cge_texture_desc td {};
        td.width = img->width;
        td.height = img->height;
        td.texel_offset = texels.size() / 3;
        for (const auto& p : img->pixels) {
            texels.push_back(p.x);
            texels.push_back(p.y);
            texels.push_back(p.z);
        }
        int32_t id = int32_t(textures.size());
        textures.push_back(td);
        texIds[img.get()] = id;
        return id;
    };
    for (const auto& mesh : sc.meshes) {
        cge_mesh_desc md {};
        md.vertex_offset = uint32_t(vertices.size());
        md.vertex_count = uint32_t(mesh.vertices.size());
        md.triangle_offset = uint32_t(tris.size() / 3);
        md.triangle_count = uint32_t(mesh.triangles.size());
        for (int k = 0; k < 3; k++) {
            md.kd[k] = mesh.material.kd[k];
            md.ks[k] = mesh.material.ks[k];
        }
        md.shininess = mesh.material.shininess;
        md.transparency = mesh.material.transparency;
        md.texture_id = texId(mesh.material.kdTexture);
        meshes.push_back(md);
        static_assert(sizeof(Vertex) == sizeof(cge_vertex));
        for (const auto& v : mesh.vertices) {
            cge_vertex cv;
            std::memcpy(&cv, &v, sizeof(cv));
            vertices.push_back(cv);
        }
        for (const auto& t : mesh.triangles) {
            tris.push_back(t.x);
            tris.push_back(t.y);
            tris.push_back(t.z);
        }
    }
    for (const auto& s : sc.spheres) {
        cge_sphere_desc sd {};
        for (int k = 0; k < 3; k++) {
            sd.center[k] = s.center[k];
            sd.kd[k] = s.material.kd[k];
            sd.ks[k] = s.material.ks[k];
        }
        sd.radius = s.radius;
        sd.shininess = s.material.shininess;
        sd.transparency = s.material.transparency;
        sd.texture_id = texId(s.material.kdTexture);
        spheres.push_back(sd);
    }
    for (const auto& l : sc.lights) {
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
        lights.push_back(ld);
    }
    std::vector<cge_bvh_node> nodes;
    std::vector<uint32_t> order;
    if (bvh) {
        for (const auto& n : bvh->nodes) {
            cge_bvh_node bn {};
            for (int k = 0; k < 3; k++) {
                bn.lower[k] = n.aabb.lower[k];
                bn.upper[k] = n.aabb.upper[k];
            }
            bn.is_leaf = uint32_t(n.data[0]);
            bn.depth = uint32_t(n.data[1]);
            bn.beg = uint32_t(n.data[2]);
            bn.end = uint32_t(n.data[3]);
            if (!bn.is_leaf) {
                bn.left = uint32_t(n.data[4]);
                bn.right = uint32_t(n.data[5]);
            }
            nodes.push_back(bn);
        }
        for (const auto& p : bvh->primitives) {
            if (std::holds_alternative<TrianglePrim>(p.p)) {
                const auto& t = std::get<TrianglePrim>(p.p);
                auto it = rs.triLookup[t.meshIdx].find({ uint32_t(t.v1), uint32_t(t.v2), uint32_t(t.v3) });
                order.push_back(it->second);
            } else {
                order.push_back(uint32_t(rs.nTriangles + std::get<SpherePrim>(p.p).sphereIdx));
            }
        }
        h.n_bvh_nodes = uint32_t(nodes.size());
        h.bvh_root = uint32_t(bvh->root);
    }
    h.n_meshes = uint32_t(meshes.size());
    h.n_vertices = uint32_t(vertices.size());
    h.n_triangles = uint32_t(tris.size() / 3);
    h.n_spheres = uint32_t(spheres.size());
    h.n_lights = uint32_t(lights.size());
    h.n_textures = uint32_t(textures.size());
    h.n_texels = texels.size() / 3;
    FILE* f = std::fopen(path, "wb");
    if (!f)
        return false;
    auto put = [&](const void* p, size_t bytes) { if (bytes) std::fwrite(p, 1, bytes, f); };
    put(&h, sizeof(h));
    put(meshes.data(), meshes.size() * sizeof(cge_mesh_desc));
    put(vertices.data(), vertices.size() * sizeof(cge_vertex));
    put(tris.data(), tris.size() * sizeof(uint32_t));
    put(spheres.data(), spheres.size() * sizeof(cge_sphere_desc));
    put(lights.data(), lights.size() * sizeof(cge_light_desc));
    put(textures.data(), textures.size() * sizeof(cge_texture_desc));
    put(texels.data(), texels.size() * sizeof(float));
    put(nodes.data(), nodes.size() * sizeof(cge_bvh_node));
    put(order.data(), order.size() * sizeof(uint32_t));
    std::fclose(f);
    return true;
}
} // namespace

extern "C" {

// ---- I1-I6 batch KATs straight from the prebuilt archive -------------------------------------------------
void ref_kat_triangle(const float* v, float* ray7, int32_t* hit, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        glm::vec3 v0(v[9 * i + 0], v[9 * i + 1], v[9 * i + 2]), v1(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]), v2(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]);
        Ray r { { ray7[7 * i], ray7[7 * i + 1], ray7[7 * i + 2] }, { ray7[7 * i + 3], ray7[7 * i + 4], ray7[7 * i + 5] }, ray7[7 * i + 6] };
        HitInfo h {};
        hit[i] = intersectRayWithTriangle(v0, v1, v2, r, h) ? 1 : 0;
        ray7[7 * i + 6] = r.t;
    }
}
void ref_kat_aabb(const float* b, float* ray7, int32_t* hit, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        AxisAlignedBox box { { b[6 * i], b[6 * i + 1], b[6 * i + 2] }, { b[6 * i + 3], b[6 * i + 4], b[6 * i + 5] } };
        Ray r { { ray7[7 * i], ray7[7 * i + 1], ray7[7 * i + 2] }, { ray7[7 * i + 3], ray7[7 * i + 4], ray7[7 * i + 5] }, ray7[7 * i + 6] };
        hit[i] = intersectRayWithShape(box, r) ? 1 : 0;
        ray7[7 * i + 6] = r.t;
    }
}
void ref_kat_sphere(const float* s, float* ray7, float* normal, int32_t* hit, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        Sphere sp { { s[4 * i], s[4 * i + 1], s[4 * i + 2] }, s[4 * i + 3], Material { glm::vec3(0.5f) } };
        Ray r { { ray7[7 * i], ray7[7 * i + 1], ray7[7 * i + 2] }, { ray7[7 * i + 3], ray7[7 * i + 4], ray7[7 * i + 5] }, ray7[7 * i + 6] };
        HitInfo h {};
        h.normal = glm::vec3(0.0f);
        hit[i] = intersectRayWithShape(sp, r, h) ? 1 : 0;
        ray7[7 * i + 6] = r.t;
        normal[3 * i] = h.normal.x;
        normal[3 * i + 1] = h.normal.y;
        normal[3 * i + 2] = h.normal.z;
    }
}
void ref_kat_plane(const float* pl, float* ray7, int32_t* hit, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        Plane p { pl[4 * i], { pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3] } };
        Ray r { { ray7[7 * i], ray7[7 * i + 1], ray7[7 * i + 2] }, { ray7[7 * i + 3], ray7[7 * i + 4], ray7[7 * i + 5] }, ray7[7 * i + 6] };
        hit[i] = intersectRayWithPlane(p, r) ? 1 : 0;
        ray7[7 * i + 6] = r.t;
    }
}
void ref_kat_triangle_plane(const float* v, float* out4, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        glm::vec3 v0(v[9 * i + 0], v[9 * i + 1], v[9 * i + 2]), v1(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]), v2(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]);
        Plane p = trianglePlane(v0, v1, v2);
        out4[4 * i] = p.D;
        out4[4 * i + 1] = p.normal.x;
        out4[4 * i + 2] = p.normal.y;
        out4[4 * i + 3] = p.normal.z;
    }
}
void ref_kat_point_in_triangle(const float* v, const float* nrm, const float* p, int32_t* inside, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        glm::vec3 v0(v[9 * i + 0], v[9 * i + 1], v[9 * i + 2]), v1(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]), v2(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]);
        glm::vec3 nn(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), pp(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        inside[i] = pointInTriangle(v0, v1, v2, nn, pp) ? 1 : 0;
    }
}

// ---- S1-S5 KATs (reference source functions, called directly) --------------------------------------------
void ref_kat_barycentric(const float* v, const float* p, float* out3, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        glm::vec3 v0(v[9 * i + 0], v[9 * i + 1], v[9 * i + 2]), v1(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]), v2(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]);
        glm::vec3 b = computeBarycentricCoord(v0, v1, v2, glm::vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        out3[3 * i] = b.x;
        out3[3 * i + 1] = b.y;
        out3[3 * i + 2] = b.z;
    }
}
// in: lightPos[3] lightColor[3] rayO[3] rayD[3] t normal[3] kd[3] ks[3] shininess  = 23 floats per case
void ref_kat_shading(const float* in23, float* out3, uint32_t n)
{
    Features f;
    f.enableShading = true;
    for (uint32_t i = 0; i < n; i++) {
        const float* a = in23 + 23 * i;
        Ray r { { a[6], a[7], a[8] }, { a[9], a[10], a[11] }, a[12] };
        HitInfo h {};
        h.normal = glm::vec3(a[13], a[14], a[15]);
        h.material.kd = glm::vec3(a[16], a[17], a[18]);
        h.material.ks = glm::vec3(a[19], a[20], a[21]);
        h.material.shininess = a[22];
        glm::vec3 c = computeShading(glm::vec3(a[0], a[1], a[2]), glm::vec3(a[3], a[4], a[5]), f, r, h);
        out3[3 * i] = c.x;
        out3[3 * i + 1] = c.y;
        out3[3 * i + 2] = c.z;
    }
}
// in: rayO[3] rayD[3] t normal[3] ks[3] = 13 floats; out: o[3] d[3] t
void ref_kat_reflection(const float* in13, float* out7, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) {
        const float* a = in13 + 13 * i;
        Ray r { { a[0], a[1], a[2] }, { a[3], a[4], a[5] }, a[6] };
        HitInfo h {};
        h.normal = glm::vec3(a[7], a[8], a[9]);
        h.material.ks = glm::vec3(a[10], a[11], a[12]);
        Ray o = computeReflectionRay(r, h);
        float* q = out7 + 7 * i;
        q[0] = o.origin.x, q[1] = o.origin.y, q[2] = o.origin.z;
        q[3] = o.direction.x, q[4] = o.direction.y, q[5] = o.direction.z;
        q[6] = o.t;
    }
}

// ---- camera constants exactly as Trackball computes them -------------------------------------------------
void ref_camera(float fovy, int width, int height, const float* lookAt, float dist, const float* rot, cge_camera* out)
{
    Window window { "oracle", glm::ivec2(width, height), OpenGLVersion::GL2, false };
    Trackball cam { &window, fovy, dist };
    cam.setCamera(glm::vec3(lookAt[0], lookAt[1], lookAt[2]), glm::vec3(rot[0], rot[1], rot[2]), dist);
    glm::vec3 p = cam.position();
    glm::quat q(cam.m_rotationEulerAngles);
    out->origin[0] = p.x, out->origin[1] = p.y, out->origin[2] = p.z;
    out->quat[0] = q.w, out->quat[1] = q.x, out->quat[2] = q.y, out->quat[3] = q.z;
    out->half_width = cam.m_halfScreenSpaceWidth;
    out->half_height = cam.m_halfScreenSpaceHeight;
}
// rays for given ndc positions (unit test of ray generation)
void ref_generate_rays(float fovy, int width, int height, const float* lookAt, float dist, const float* rot,
                       const float* ndc2, float* ray7, uint32_t n)
{
    Window window { "oracle", glm::ivec2(width, height), OpenGLVersion::GL2, false };
    Trackball cam { &window, fovy, dist };
    cam.setCamera(glm::vec3(lookAt[0], lookAt[1], lookAt[2]), glm::vec3(rot[0], rot[1], rot[2]), dist);
    for (uint32_t i = 0; i < n; i++) {
        Ray r = cam.generateRay(glm::vec2(ndc2[2 * i], ndc2[2 * i + 1]));
        float* q = ray7 + 7 * i;
        q[0] = r.origin.x, q[1] = r.origin.y, q[2] = r.origin.z;
        q[3] = r.direction.x, q[4] = r.direction.y, q[5] = r.direction.z;
        q[6] = r.t;
    }
}

// ---- scenes ------------------------------------------------------------------------------------------------
void* ref_scene_load_flat(const char* path)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return nullptr;
    cge_scene_file_header h;
    if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, CGE_SCENE_FILE_MAGIC, 8) != 0) {
        std::fclose(f);
        return nullptr;
    }
    auto rd = [&](auto& vec, size_t count) {
        vec.resize(count);
        if (count)
            (void)!std::fread(vec.data(), sizeof(vec[0]), count, f);
    };
    std::vector<cge_mesh_desc> meshes;
    std::vector<cge_vertex> vertices;
    std::vector<uint32_t> tris;
    std::vector<cge_sphere_desc> spheres;
    std::vector<cge_light_desc> lights;
    std::vector<cge_texture_desc> textures;
    std::vector<float> texels;
    rd(meshes, h.n_meshes);
    rd(vertices, h.n_vertices);
    rd(tris, size_t(h.n_triangles) * 3);
    rd(spheres, h.n_spheres);
    rd(lights, h.n_lights);
    rd(textures, h.n_textures);
    rd(texels, size_t(h.n_texels) * 3);
    std::fclose(f);

    auto* rs = new RefScene();
    std::vector<std::shared_ptr<Image>> images;
    for (const auto& td : textures)
        images.push_back(makeImage(td.width, td.height, texels.data() + td.texel_offset * 3));
    auto material = [&](const float* kd, const float* ks, float sh, float tr, int32_t tex) {
        Material m;
        m.kd = glm::vec3(kd[0], kd[1], kd[2]);
        m.ks = glm::vec3(ks[0], ks[1], ks[2]);
        m.shininess = sh;
        m.transparency = tr;
        if (tex >= 0)
            m.kdTexture = images[size_t(tex)];
        return m;
    };
    for (const auto& md : meshes) {
        Mesh mesh;
        mesh.vertices.resize(md.vertex_count);
        std::memcpy(mesh.vertices.data(), vertices.data() + md.vertex_offset, size_t(md.vertex_count) * sizeof(Vertex));
        mesh.triangles.resize(md.triangle_count);
        for (uint32_t t = 0; t < md.triangle_count; t++) {
            const uint32_t* q = tris.data() + 3 * size_t(md.triangle_offset + t);
            mesh.triangles[t] = glm::uvec3(q[0], q[1], q[2]);
        }
        mesh.material = material(md.kd, md.ks, md.shininess, md.transparency, md.texture_id);
        rs->scene.meshes.push_back(std::move(mesh));
    }
    for (const auto& sd : spheres)
        rs->scene.spheres.push_back(Sphere { glm::vec3(sd.center[0], sd.center[1], sd.center[2]), sd.radius,
            material(sd.kd, sd.ks, sd.shininess, sd.transparency, sd.texture_id) });
    for (const auto& ld : lights) {
        if (ld.type == CGE_LIGHT_POINT) {
            PointLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            rs->scene.lights.emplace_back(l);
        } else if (ld.type == CGE_LIGHT_SEGMENT) {
            SegmentLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            rs->scene.lights.emplace_back(l);
        } else {
            ParallelogramLight l;
            std::memcpy(&l, ld.v, sizeof(l));
            rs->scene.lights.emplace_back(l);
        }
    }
    rs->scene.type = Custom;
    finishScene(*rs);
    return rs;
}

void ref_scene_free(void* s) { delete static_cast<RefScene*>(s); }

// Reference loaders -> flat file.  Only works where the reference data directory exists (build container).
// kind 0: loadScenePrebuilt(SceneType(arg)) from dataDir;  kind 1: loadMesh(path, normalize=arg) only (no lights).
int ref_scene_export(int kind, int arg, const char* pathOrDataDir, uint32_t featureBitsForBvh, int withBvh, const char* outPath)
{
    try {
        RefScene rs;
        if (kind == 0) {
            rs.scene = loadScenePrebuilt(SceneType(arg), pathOrDataDir);
        } else {
            auto sub = loadMesh(pathOrDataDir, arg != 0);
            std::move(sub.begin(), sub.end(), std::back_inserter(rs.scene.meshes));
            rs.scene.type = Custom;
        }
        finishScene(rs);
        std::unique_ptr<BvhInterface> bvh;
        if (withBvh)
            bvh = std::make_unique<BvhInterface>(&rs.scene, featuresFromBits(featureBitsForBvh));
        return writeFlat(rs, bvh ? bvh->m_impl : nullptr, outPath) ? 0 : 2;
    } catch (...) {
        return 1;
    }
}

// Re-write a loaded flat scene together with the reference-built BVH (used to pin the BVH builder).
int ref_scene_write_with_bvh(void* scene, uint32_t featureBits, const char* outPath)
{
    auto* rs = static_cast<RefScene*>(scene);
    BvhInterface bvh(&rs->scene, featuresFromBits(featureBits));
    return writeFlat(*rs, bvh.m_impl, outPath) ? 0 : 2;
}

void* ref_bvh_build(void* scene, uint32_t featureBits)
{
    auto* rs = static_cast<RefScene*>(scene);
    auto* b = new RefBvh();
    b->features = featuresFromBits(featureBits);
    b->bvh = std::make_unique<BvhInterface>(&rs->scene, b->features);
    return b;
}
void ref_bvh_free(void* b) { delete static_cast<RefBvh*>(b); }
void ref_bvh_info(void* b, int32_t* nNodes, int32_t* nLevels, int32_t* nLeaves)
{
    auto* rb = static_cast<RefBvh*>(b);
    *nNodes = int32_t(rb->bvh->m_impl->nodes.size());
    *nLevels = rb->bvh->numLevels();
    *nLeaves = rb->bvh->numLeaves();
}

#ifdef CGE_REF_GPU_SHIM
extern "C" void cgeSetSamplerSeed(uint32_t seed); // computer-graphics-engine_b200/host/render_gpu.cpp
#endif

struct ref_render_params {
    int32_t width, height;
    uint32_t features;
    int32_t ray_depth;
    int32_t segment_samples, parallelogram_samples;
    uint32_t sampler; // 0 glibc rand(), 1 hash
    uint32_t seed;
    int32_t threads;                // OMP threads (<=0: all)
    int32_t use_render_ray_tracing; // 1: call the reference's own renderRayTracing (depth literal 5)
    int32_t want_ids;
    float fovy;        // radians
    float look_at[3];
    float dist;
    float rotation[3]; // radians
    // restrict to a pixel window [x0,x1) x [y0,y1) in reference coordinates (y up); all-zero = full frame.
    int32_t x0, y0, x1, y1;
    // bounded sample for timing: render only rows y with (y - y0) % y_stride == 0 (0/1 = every row).
    int32_t y_stride;
    // ExtraFeatures globals (src/render.cpp:14,19-21), applied when the corresponding feature bit (16 + index in ExtraFeatures) is set
    int32_t rays_per_pixel_side;
    float bloom_scalar, bloom_threshold;
    int32_t bloom_debug_option;
};
struct ref_render_stats {
    uint64_t rays, box_tests, tri_tests, sphere_tests;
    double ms;
};

// rgb: W*H*3 floats in Screen::pixels() order (row 0 = top).  ids: W*H int32, same order (optional).
int ref_render(void* scene, void* bvhHandle, const ref_render_params* p, float* rgb, int32_t* ids, ref_render_stats* stats)
{
    auto* rs = static_cast<RefScene*>(scene);
    auto* rb = static_cast<RefBvh*>(bvhHandle);
    const Features features = featuresFromBits(p->features);
    segmentLightSamples = p->segment_samples;
    parallelogramLightDirectionSamples = p->parallelogram_samples;
    g_samplerMode = int(p->sampler);
    g_seed = p->seed;
#ifdef CGE_REF_GPU_SHIM
    cgeSetSamplerSeed(p->seed); // libcge_ref_gpu.so: renderRayTracing is the GPU shim; its stateless sampler gets the same seed
#endif
    if (p->rays_per_pixel_side > 0)
        raysPerPixelSide = p->rays_per_pixel_side;
    if (features.extra.enableBloomEffect) {
        bloomScalar = p->bloom_scalar;
        bloomThreshold = p->bloom_threshold;
        bloomDebugOption = p->bloom_debug_option;
    }
    if (p->threads > 0)
        omp_set_num_threads(p->threads);
    else
        omp_set_num_threads(omp_get_num_procs());

    const int W = p->width, H = p->height;
    Window window { "oracle", glm::ivec2(W, H), OpenGLVersion::GL2, false };
    Trackball camera { &window, p->fovy, p->dist };
    camera.setCamera(glm::vec3(p->look_at[0], p->look_at[1], p->look_at[2]),
        glm::vec3(p->rotation[0], p->rotation[1], p->rotation[2]), p->dist);
    Screen screen { glm::ivec2(W, H), false };
    const BvhInterface& bvh = *rb->bvh;
    g_rays = g_boxes = g_tris = g_spheres = 0;

    int x0 = p->x0, y0 = p->y0, x1 = p->x1, y1 = p->y1;
    if (x0 == 0 && y0 == 0 && x1 == 0 && y1 == 0) {
        x1 = W;
        y1 = H;
    }

    const int ystride = p->y_stride > 1 ? p->y_stride : 1;
    const int nRows = (y1 - y0 + ystride - 1) / ystride;
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (p->use_render_ray_tracing) {
        renderRayTracing(rs->scene, camera, bvh, screen, features);
    } else {
        // Same loop shape as src/render.cpp:277-289,316-323 with the depth as a parameter.
#pragma omp parallel
        {
            tls = Tls {};
#pragma omp for schedule(guided)
            for (int row = 0; row < nRows; row++) {
                const int y = y0 + row * ystride;
                for (int x = x0; x != x1; x++) {
                    const glm::vec2 normalizedPixelPos {
                        float(x) / float(W) * 2.0f - 1.0f,
                        float(y) / float(H) * 2.0f - 1.0f
                    };
                    tls.pixel = uint32_t(y) * uint32_t(W) + uint32_t(x);
                    tls.counter = 0;
                    if (features.extra.enableMultipleRaysPerPixel) {
                        // src/render.cpp:290-303,322 with the reference's own getRaySamples
                        const glm::vec2 pixelSize { 1 / float(W) * 2.f, 1 / float(H) * 2.f };
                        auto colorSum = glm::vec3(0.f);
                        size_t weight = 0;
                        auto color = glm::vec3(0.f);
                        for (auto& ray : getRaySamples(normalizedPixelPos, pixelSize, camera, raysPerPixelSide))
                            color += getFinalColor(rs->scene, bvh, ray, features, p->ray_depth);
                        color /= raysPerPixelSide * raysPerPixelSide;
                        colorSum += color;
                        weight++;
                        screen.setPixel(x, y, colorSum / float(weight));
                        continue;
                    }
                    const Ray cameraRay = camera.generateRay(normalizedPixelPos);
                    screen.setPixel(x, y, getFinalColor(rs->scene, bvh, cameraRay, features, p->ray_depth));
                }
            }
            g_rays += tls.rays;
            g_boxes += tls.boxes;
            g_tris += tls.tris;
            g_spheres += tls.spheres;
        }
    }
    if (!p->use_render_ray_tracing && features.extra.enableBloomEffect)
        renderBloomFilter(screen, features); // src/render.cpp:326-328 (renderRayTracing applies it itself)
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (stats) {
        stats->ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->rays = g_rays;
        stats->box_tests = g_boxes;
        stats->tri_tests = g_tris;
        stats->sphere_tests = g_spheres;
    }
    if (rgb)
        std::memcpy(rgb, screen.pixels().data(), size_t(W) * size_t(H) * 3 * sizeof(float));

    if (ids && p->want_ids) {
#ifdef CGE_REF_NO_COUNTERS
        return 3;
#else
#pragma omp parallel for schedule(guided)
        for (int row = 0; row < nRows; row++) {
            const int y = y0 + row * ystride;
            for (int x = x0; x != x1; x++) {
                const glm::vec2 ndc { float(x) / float(W) * 2.0f - 1.0f, float(y) / float(H) * 2.0f - 1.0f };
                Ray ray = camera.generateRay(ndc);
                HitInfo hi;
                tls.lastV[0] = nullptr;
                tls.lastSphere = nullptr;
                const bool hit = bvh.intersect(ray, hi, features);
                ids[size_t(H - 1 - y) * size_t(W) + size_t(x)] = hit ? lastPrimitiveId(*rs) : -1;
            }
        }
#endif
    }
    return 0;
}

// getFinalColor for explicit rays (src/render.h:35); sampler pixel id = ray index.
int ref_trace_rays(void* scene, void* bvhHandle, const float* ray7, uint32_t n, const ref_render_params* p, float* rgb, int32_t* ids)
{
    auto* rs = static_cast<RefScene*>(scene);
    auto* rb = static_cast<RefBvh*>(bvhHandle);
    const Features features = featuresFromBits(p->features);
    segmentLightSamples = p->segment_samples;
    parallelogramLightDirectionSamples = p->parallelogram_samples;
    g_samplerMode = int(p->sampler);
    g_seed = p->seed;
    for (uint32_t i = 0; i < n; i++) {
        const float* q = ray7 + 7 * i;
        Ray r { { q[0], q[1], q[2] }, { q[3], q[4], q[5] }, q[6] };
        tls.pixel = i;
        tls.counter = 0;
        glm::vec3 c = getFinalColor(rs->scene, *rb->bvh, r, features, p->ray_depth);
        rgb[3 * i] = c.x, rgb[3 * i + 1] = c.y, rgb[3 * i + 2] = c.z;
        if (ids) {
#ifndef CGE_REF_NO_COUNTERS
            Ray r2 { { q[0], q[1], q[2] }, { q[3], q[4], q[5] }, q[6] };
            HitInfo hi;
            tls.lastV[0] = nullptr;
            tls.lastSphere = nullptr;
            ids[i] = rb->bvh->intersect(r2, hi, features) ? lastPrimitiveId(*rs) : -1;
#else
            ids[i] = -3;
#endif
        }
    }
    return 0;
}

// BvhInterface::intersect for explicit rays: out per ray = hit, t, normal[3], kd[3], ks[3], shininess, transparency
int ref_intersect_rays(void* scene, void* bvhHandle, const float* ray7, uint32_t n, uint32_t featureBits,
                       int32_t* hit, float* t, float* normal3, float* material8, int32_t* ids)
{
    auto* rs = static_cast<RefScene*>(scene);
    auto* rb = static_cast<RefBvh*>(bvhHandle);
    const Features features = featuresFromBits(featureBits);
    for (uint32_t i = 0; i < n; i++) {
        const float* q = ray7 + 7 * i;
        Ray r { { q[0], q[1], q[2] }, { q[3], q[4], q[5] }, q[6] };
        HitInfo hi;
        hi.normal = glm::vec3(0.0f);
        tls.lastV[0] = nullptr;
        tls.lastSphere = nullptr;
        const bool h = rb->bvh->intersect(r, hi, features);
        hit[i] = h ? 1 : 0;
        t[i] = r.t;
        if (h) {
            for (int k = 0; k < 3; k++) {
                normal3[3 * i + k] = hi.normal[k];
                material8[8 * i + k] = hi.material.kd[k];
                material8[8 * i + 3 + k] = hi.material.ks[k];
            }
            material8[8 * i + 6] = hi.material.shininess;
            material8[8 * i + 7] = hi.material.transparency;
        }
        if (ids) {
#ifndef CGE_REF_NO_COUNTERS
            ids[i] = h ? lastPrimitiveId(*rs) : -1;
#else
            ids[i] = -3;
#endif
        }
    }
    return 0;
}

// Screen::writeBitmapToFile (src/screen.cpp:49-60) applied to a caller-supplied float frame (Screen::pixels() order):
// the reference's own clamp -> *255 -> u8x4 conversion and BMP writer, for the output-stage parity test.
int ref_write_bmp(const float* rgb, int width, int height, const char* path)
{
    Screen screen { glm::ivec2(width, height), false };
    std::memcpy(screen.pixels().data(), rgb, size_t(width) * size_t(height) * 3 * sizeof(float));
    screen.writeBitmapToFile(path);
    return 0;
}

int ref_has_counters(void)
{
#ifdef CGE_REF_NO_COUNTERS
    return 0;
#else
    return 1;
#endif
}
// weightsGaussian(sigma) (src/render.cpp:198-210), column-major 3x3 as glm stores it
void ref_weights_gaussian(float sigma, float* out9)
{
    const glm::mat3 m = weightsGaussian(sigma);
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            out9[c * 3 + r] = m[c][r];
}

// getRaySamples (src/render.cpp:211-227) for one pixel with the hash-seeded std::mt19937: n*n rays as (origin, direction)
int ref_ray_samples(const ref_render_params* p, int x, int y, float* ray6)
{
    g_samplerMode = 1;
    g_seed = p->seed;
    const int W = p->width, H = p->height;
    Window window { "oracle", glm::ivec2(W, H), OpenGLVersion::GL2, false };
    Trackball camera { &window, p->fovy, p->dist };
    camera.setCamera(glm::vec3(p->look_at[0], p->look_at[1], p->look_at[2]),
        glm::vec3(p->rotation[0], p->rotation[1], p->rotation[2]), p->dist);
    const glm::vec2 pos { float(x) / float(W) * 2.0f - 1.0f, float(y) / float(H) * 2.0f - 1.0f };
    const glm::vec2 pixelSize { 1 / float(W) * 2.f, 1 / float(H) * 2.f };
    tls.pixel = uint32_t(y) * uint32_t(W) + uint32_t(x);
    tls.counter = 0;
    int k = 0;
    for (auto& r : getRaySamples(pos, pixelSize, camera, p->rays_per_pixel_side)) {
        const float v[6] = { r.origin.x, r.origin.y, r.origin.z, r.direction.x, r.direction.y, r.direction.z };
        std::memcpy(ray6 + 6 * k++, v, sizeof(v));
    }
    return k;
}

int ref_num_procs(void) { return omp_get_num_procs(); }

} // extern "C"
