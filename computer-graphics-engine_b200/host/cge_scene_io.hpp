// cge_scene_io.hpp — the reference's scene loaders on the host side of the drop-in interface.
//
//   loadMesh(file, centerAndNormalize)        reference framework/src/mesh.cpp:52-176: OBJ + MTL -> one Mesh per (shape, material run),
//                                             vertices de-duplicated in order of first use, optional centring / scaling to the unit sphere
//   Image loadImage(file)                     framework/src/image.cpp:12-35: 8-bit RGB texels / 255.0f (PNG, non-interlaced)
//   loadScenePrebuilt(type, dataDir)          src/scene.cpp:5-92: the ten built-in scenes with their lights
//   loadSceneFromFile(file, lights)           src/scene.cpp:94-103
//   deserializeSceneType(name)                src/config.cpp:404-431
//   saveFlatScene(flat, file)                 the flat scene file of include/cge_scene_file.h (what loadFlatScene reads)
//
// Host I/O only - nothing here is on the ray-tracing path.  The point of these functions is that the CLI (cge_cli.cpp) consumes the
// reference's own config + data directory, and that the scene they produce is the reference's scene BIT FOR BIT (the renderer's
// parity is stated against it): tests/test_scene_io.py compares every built-in scene with the fixture the unmodified reference
// loaders exported.  Two third-party behaviours are therefore restated rather than approximated (the reference vendors
// tinyobjloader 2.0.0rc and stb_image; neither is copied here):
//   * OBJ / MTL numbers are read the way tinyobjloader's parser reads them - digits accumulated into a double one by one, the
//     fraction through powers of ten, a decimal exponent applied as 5^e * 2^e - and then narrowed to float; strtod() rounds
//     correctly and would differ from it in the last bit of an occasional coordinate;
//   * quads are split along their shorter diagonal (0-2 if strictly shorter, else 1-3); polygons with more than four corners are
//     refused (the reference's data directory has none; ear clipping is not restated).
// PNG decoding uses zlib for the inflate step.
#pragma once
#include "cge_engine.hpp"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace cge_engine {

enum SceneType { SingleTriangle, Cube, CubeTextured, CornellBox, CornellBoxParallelogramLight, Monkey, Teapot, Dragon, Spheres, Custom };

namespace io_detail {

inline bool isDigit(char c) { return c >= '0' && c <= '9'; }

// One number of an OBJ / MTL line, [s, end).  Leaves *out alone when the text is not a number (the caller's default stands).
inline void parseDouble(const char* s, const char* end, double* out)
{
    if (s >= end)
        return;
    const char* c = s;
    double sign = 1.0, mantissa = 0.0;
    bool leadingDot = false;
    if (*c == '+' || *c == '-') {
        sign = *c == '-' ? -1.0 : 1.0;
        c++;
        if (c != end && *c == '.')
            leadingDot = true;
    } else if (*c == '.') {
        leadingDot = true;
    } else if (!isDigit(*c)) {
        return;
    }
    int read = 0;
    if (!leadingDot) {
        while (c != end && isDigit(*c)) {
            mantissa = mantissa * 10 + double(*c - '0');
            c++, read++;
        }
        if (read == 0)
            return;
    }
    int exponent = 0;
    auto assemble = [&]() { *out = sign * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa); };
    if (c == end)
        return assemble();
    if (*c == '.') {
        c++;
        read = 1;
        static const double lut[] = { 1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001 };
        while (c != end && isDigit(*c)) {
            mantissa += double(*c - '0') * (read < 8 ? lut[read] : std::pow(10.0, -read));
            read++, c++;
        }
    } else if (*c != 'e' && *c != 'E') {
        return assemble();
    }
    if (c == end)
        return assemble();
    if (*c == 'e' || *c == 'E') {
        c++;
        bool negative = false;
        if (c != end && (*c == '+' || *c == '-')) {
            negative = *c == '-';
            c++;
        } else if (c == end || !isDigit(*c)) {
            return; // an empty exponent is not a number
        }
        read = 0;
        while (c != end && isDigit(*c)) {
            if (exponent > 2147483647 / 10)
                return;
            exponent = exponent * 10 + (*c - '0');
            c++, read++;
        }
        if (negative)
            exponent = -exponent;
        if (read == 0)
            return;
    }
    assemble();
}

struct Cursor {
    const char* p;
    void skipSpace() { p += std::strspn(p, " \t"); }
    float real(double dflt = 0.0)
    {
        skipSpace();
        const char* end = p + std::strcspn(p, " \t\r");
        double v = dflt;
        parseDouble(p, end, &v);
        p = end;
        return float(v);
    }
    std::string word()
    {
        skipSpace();
        const size_t n = std::strcspn(p, " \t\r");
        std::string s(p, n);
        p += n;
        return s;
    }
    bool atEnd()
    {
        skipSpace();
        return *p == '\0' || *p == '\r' || *p == '\n';
    }
};

struct ObjMaterial {
    float diffuse[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 };
    float shininess = 1.0f, dissolve = 1.0f;
    std::string diffuseTexname;
};
struct ObjIndex {
    int v = -1, vt = -1, vn = -1;
};
struct ObjShape {
    std::vector<ObjIndex> indices; // three per triangle
    std::vector<int> materialIds;  // one per triangle
};
struct ObjFile {
    std::vector<float> v, vn, vt;
    std::vector<ObjShape> shapes;
    std::vector<ObjMaterial> materials;
};

inline bool getLine(std::istream& in, std::string& line)
{
    line.clear();
    if (!std::getline(in, line))
        return !line.empty();
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n'))
        line.pop_back();
    return true;
}

inline void loadMtl(const std::string& path, std::map<std::string, int>& byName, std::vector<ObjMaterial>& out)
{
    std::ifstream in(path);
    if (!in)
        return; // (the OBJ loader only warns about a missing material library: the faces keep material id -1)
    ObjMaterial cur;
    std::string name;
    bool haveD = false;  // reset by every newmtl
    bool haveKd = false; // NOT reset by newmtl (the library the reference uses keeps it for the whole file)
    auto flush = [&]() {
        if (!name.empty()) {
            byName.insert({ name, int(out.size()) }); // the first material of a name wins
            out.push_back(cur);
        }
    };
    std::string line;
    while (getLine(in, line)) {
        Cursor c { line.c_str() };
        c.skipSpace();
        const char* t = c.p;
        if (*t == '\0' || *t == '#')
            continue;
        auto is = [&](const char* key) {
            const size_t n = std::strlen(key);
            return std::strncmp(t, key, n) == 0 && (t[n] == ' ' || t[n] == '\t');
        };
        if (is("newmtl")) {
            flush();
            cur = ObjMaterial();
            haveD = false;
            c.p = t + 7;
            name = c.word();
        } else if (is("Kd")) {
            c.p = t + 2;
            for (float& f : cur.diffuse)
                f = c.real();
            haveKd = true;
        } else if (is("Ks")) {
            c.p = t + 2;
            for (float& f : cur.specular)
                f = c.real();
        } else if (is("Ns")) {
            c.p = t + 2;
            cur.shininess = c.real();
        } else if (is("d")) {
            c.p = t + 1;
            cur.dissolve = c.real();
            haveD = true;
        } else if (is("Tr")) {
            c.p = t + 2;
            if (!haveD) // `d` wins over `Tr`
                cur.dissolve = 1.0f - c.real();
        } else if (is("map_Kd")) {
            c.p = t + 6;
            c.skipSpace();
            if (*c.p == '-')
                throw std::runtime_error("map_Kd texture options are not supported: " + path);
            cur.diffuseTexname = c.word();
            if (!haveKd) // a diffuse texture without any Kd so far: the loader's default grey
                cur.diffuse[0] = cur.diffuse[1] = cur.diffuse[2] = float(0.6);
        }
    }
    flush();
}

inline int fixIndex(int idx, int n, const std::string& path)
{
    if (idx > 0)
        return idx - 1;
    if (idx == 0)
        throw std::runtime_error("OBJ index 0 in " + path);
    return n + idx; // relative to the elements read so far
}

inline ObjFile loadObj(const std::string& path)
{
    std::ifstream in(path);
    if (!in)
        throw std::runtime_error("File " + path + " does not exist.");
    const size_t slash = path.find_last_of('/');
    const std::string baseDir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
    ObjFile obj;
    std::map<std::string, int> materialByName;
    struct Face {
        std::vector<ObjIndex> corners;
    };
    std::vector<Face> pending; // faces since the last flush (they all carry the material that was current when the flush happens)
    ObjShape shape;
    int material = -1;
    // faces waiting -> triangles of `shape`, with the material current at this moment
    auto flushFaces = [&]() {
        for (const Face& f : pending) {
            const size_t n = f.corners.size();
            if (n < 3)
                continue;
            auto emit = [&](int a, int b, int c) {
                shape.indices.push_back(f.corners[size_t(a)]);
                shape.indices.push_back(f.corners[size_t(b)]);
                shape.indices.push_back(f.corners[size_t(c)]);
                shape.materialIds.push_back(material);
            };
            if (n == 3) {
                emit(0, 1, 2);
            } else if (n == 4) {
                float p[4][3];
                bool valid = true;
                for (int k = 0; k < 4; k++) {
                    const size_t vi = size_t(f.corners[size_t(k)].v);
                    if (3 * vi + 2 >= obj.v.size()) {
                        valid = false;
                        break;
                    }
                    for (int a = 0; a < 3; a++)
                        p[k][a] = obj.v[3 * vi + size_t(a)];
                }
                if (!valid)
                    continue;
                const float e02x = p[2][0] - p[0][0], e02y = p[2][1] - p[0][1], e02z = p[2][2] - p[0][2];
                const float e13x = p[3][0] - p[1][0], e13y = p[3][1] - p[1][1], e13z = p[3][2] - p[1][2];
                const float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z, sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
                if (sqr02 < sqr13) {
                    emit(0, 1, 2);
                    emit(0, 2, 3);
                } else {
                    emit(0, 1, 3);
                    emit(1, 2, 3);
                }
            } else {
                throw std::runtime_error("polygons with more than four corners are not supported: " + path);
            }
        }
        pending.clear();
    };
    std::string line;
    while (getLine(in, line)) {
        Cursor c { line.c_str() };
        c.skipSpace();
        const char* t = c.p;
        if (*t == '\0' || *t == '#')
            continue;
        const bool sp1 = t[1] == ' ' || t[1] == '\t', sp2 = t[2] == ' ' || t[2] == '\t';
        if (t[0] == 'v' && sp1) {
            c.p = t + 2;
            for (int a = 0; a < 3; a++)
                obj.v.push_back(c.real());
        } else if (t[0] == 'v' && t[1] == 'n' && sp2) {
            c.p = t + 3;
            for (int a = 0; a < 3; a++)
                obj.vn.push_back(c.real());
        } else if (t[0] == 'v' && t[1] == 't' && sp2) {
            c.p = t + 3;
            for (int a = 0; a < 2; a++)
                obj.vt.push_back(c.real());
        } else if (t[0] == 'f' && sp1) {
            c.p = t + 2;
            Face face;
            const int nv = int(obj.v.size() / 3), nvn = int(obj.vn.size() / 3), nvt = int(obj.vt.size() / 2);
            while (!c.atEnd()) {
                // i, i/j, i//k or i/j/k
                ObjIndex idx;
                idx.v = fixIndex(std::atoi(c.p), nv, path);
                c.p += std::strcspn(c.p, "/ \t\r");
                if (*c.p == '/') {
                    c.p++;
                    if (*c.p == '/') {
                        c.p++;
                        idx.vn = fixIndex(std::atoi(c.p), nvn, path);
                        c.p += std::strcspn(c.p, "/ \t\r");
                    } else {
                        idx.vt = fixIndex(std::atoi(c.p), nvt, path);
                        c.p += std::strcspn(c.p, "/ \t\r");
                        if (*c.p == '/') {
                            c.p++;
                            idx.vn = fixIndex(std::atoi(c.p), nvn, path);
                            c.p += std::strcspn(c.p, "/ \t\r");
                        }
                    }
                }
                face.corners.push_back(idx);
            }
            pending.push_back(std::move(face));
        } else if (std::strncmp(t, "usemtl", 6) == 0) {
            c.p = t + 6;
            const std::string name = c.word();
            const auto it = materialByName.find(name);
            const int next = it == materialByName.end() ? -1 : it->second;
            if (next != material) { // the faces read so far keep the material they were read under
                flushFaces();
                material = next;
            }
        } else if (std::strncmp(t, "mtllib", 6) == 0 && (t[6] == ' ' || t[6] == '\t')) {
            c.p = t + 7;
            while (!c.atEnd()) {
                const size_t before = obj.materials.size();
                const std::string name = c.word();
                std::ifstream probe(baseDir + name);
                if (probe) {
                    loadMtl(baseDir + name, materialByName, obj.materials);
                    (void)before;
                    break; // the first library that can be opened is the one that is read
                }
            }
        } else if ((t[0] == 'g' || t[0] == 'o') && sp1) {
            flushFaces();
            if (!shape.indices.empty())
                obj.shapes.push_back(std::move(shape));
            shape = ObjShape(); // (the current material carries over into the next group)
        }
    }
    flushFaces();
    if (!shape.indices.empty())
        obj.shapes.push_back(std::move(shape));
    return obj;
}

struct VertexKey {
    float f[8];
    bool operator==(const VertexKey& o) const
    {
        for (int i = 0; i < 8; i++)
            if (!(f[i] == o.f[i])) // value equality, as the reference's defaulted operator== (so -0 == +0, NaN != NaN)
                return false;
        return true;
    }
};
struct VertexKeyHash {
    size_t operator()(const VertexKey& k) const
    {
        size_t seed = 0;
        for (float v : k.f)
            seed ^= std::hash<float>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};

inline float length3(const vec3& v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

// ---- PNG -----------------------------------------------------------------------------------------------------------------------------
inline uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

} // namespace io_detail

// 8-bit RGB texels / 255.0f, rows top to bottom (reference framework/src/image.cpp:12-35: stbi_load(..., STBI_rgb)).  PNG only
// (what the reference's data directory holds): grey, RGB, palette, grey + alpha, RGBA; 1 - 16 bits; not interlaced.
inline Image loadImage(const std::string& path)
{
    using namespace io_detail;
    std::ifstream in(path, std::ios::binary);
    if (!in)
        throw std::runtime_error("Texture file " + path + " does not exists!");
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0)
        throw std::runtime_error("Failed to read texture " + path + ": not a PNG file");
    uint32_t w = 0, h = 0;
    int depth = 0, colour = 0, interlace = 0;
    std::vector<unsigned char> palette, idat;
    for (size_t at = 8; at + 12 <= file.size();) {
        const uint32_t n = be32(&file[at]);
        const unsigned char* type = &file[at + 4];
        const unsigned char* data = &file[at + 8];
        if (at + 12 + n > file.size())
            throw std::runtime_error("Failed to read texture " + path + ": truncated chunk");
        if (std::memcmp(type, "IHDR", 4) == 0 && n >= 13) {
            w = be32(data), h = be32(data + 4);
            depth = data[8], colour = data[9], interlace = data[12];
        } else if (std::memcmp(type, "PLTE", 4) == 0) {
            palette.assign(data, data + n);
        } else if (std::memcmp(type, "IDAT", 4) == 0) {
            idat.insert(idat.end(), data, data + n);
        } else if (std::memcmp(type, "IEND", 4) == 0) {
            break;
        }
        at += 12 + size_t(n);
    }
    const int channels = colour == 0 ? 1 : colour == 2 ? 3 : colour == 3 ? 1 : colour == 4 ? 2 : colour == 6 ? 4 : 0;
    if (!w || !h || !channels || interlace || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16))
        throw std::runtime_error("Failed to read texture " + path + ": unsupported PNG variant");
    const size_t bitsPerPixel = size_t(channels) * size_t(depth), stride = (size_t(w) * bitsPerPixel + 7) / 8;
    const size_t bpp = std::max<size_t>(1, bitsPerPixel / 8); // filter distance in bytes
    std::vector<unsigned char> raw((stride + 1) * size_t(h));
    uLongf rawLen = uLongf(raw.size());
    if (uncompress(raw.data(), &rawLen, idat.data(), uLong(idat.size())) != Z_OK || rawLen != raw.size())
        throw std::runtime_error("Failed to read texture " + path + ": bad image data");
    // undo the per-row filters in place
    std::vector<unsigned char> zero(stride, 0);
    for (size_t y = 0; y < h; y++) {
        unsigned char* row = &raw[y * (stride + 1) + 1];
        const unsigned char* up = y ? &raw[(y - 1) * (stride + 1) + 1] : zero.data();
        const int filter = raw[y * (stride + 1)];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? row[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
            int pred = 0;
            switch (filter) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) / 2; break;
            case 4: {
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            } break;
            default: throw std::runtime_error("Failed to read texture " + path + ": bad filter");
            }
            row[x] = (unsigned char)(row[x] + pred);
        }
    }
    Image img;
    img.width = int(w), img.height = int(h);
    img.pixels.reserve(size_t(w) * h);
    // one sample as 8 bits: 16-bit samples keep their high byte; grey samples of fewer bits are stretched to 0..255; palette
    // indices stay indices
    auto sample = [&](const unsigned char* row, size_t index) -> unsigned {
        if (depth == 8)
            return row[index];
        if (depth == 16)
            return row[2 * index];
        const size_t bit = index * size_t(depth);
        const unsigned v = (row[bit / 8] >> (8 - depth - int(bit % 8))) & ((1u << depth) - 1u);
        return colour == 3 ? v : v * (depth == 1 ? 255u : depth == 2 ? 85u : 17u);
    };
    for (size_t y = 0; y < h; y++) {
        const unsigned char* row = &raw[y * (stride + 1) + 1];
        for (size_t x = 0; x < w; x++) {
            unsigned r, g, b;
            if (colour == 3) {
                const unsigned i = sample(row, x);
                if (3 * size_t(i) + 2 >= palette.size())
                    throw std::runtime_error("Failed to read texture " + path + ": palette index out of range");
                r = palette[3 * i], g = palette[3 * i + 1], b = palette[3 * i + 2];
            } else if (channels <= 2) {
                r = g = b = sample(row, x * size_t(channels));
            } else {
                r = sample(row, x * size_t(channels)), g = sample(row, x * size_t(channels) + 1), b = sample(row, x * size_t(channels) + 2);
            }
            img.pixels.emplace_back(float(r) / 255.0f, float(g) / 255.0f, float(b) / 255.0f);
        }
    }
    return img;
}

inline std::vector<Mesh> loadMesh(const std::string& file, bool centerAndNormalize = false)
{
    using namespace io_detail;
    const ObjFile obj = loadObj(file);
    const size_t slash = file.find_last_of('/');
    const std::string baseDir = slash == std::string::npos ? std::string() : file.substr(0, slash + 1);
    std::vector<Mesh> out;
    auto pos = [&](int vi) { return vec3(obj.v[3 * size_t(vi)], obj.v[3 * size_t(vi) + 1], obj.v[3 * size_t(vi) + 2]); };
    for (const ObjShape& shape : obj.shapes) {
        const size_t nTris = shape.indices.size() / 3;
        size_t start = 0;
        int prevMaterial = shape.materialIds[0];
        // one Mesh per run of equal material ids; the last triangle of a shape always closes the current run, even when its own
        // material differs (src: framework/src/mesh.cpp:78-85)
        for (size_t end = 0; end < nTris; ++end) {
            if (end == nTris - 1)
                ++end;
            else if (shape.materialIds[end] == prevMaterial)
                continue;
            else
                prevMaterial = shape.materialIds[end];
            Mesh mesh;
            std::unordered_map<VertexKey, uint32_t, VertexKeyHash> cache;
            for (size_t i = start * 3; i != end * 3; i += 3) {
                const vec3 v0 = pos(shape.indices[i].v), v1 = pos(shape.indices[i + 1].v), v2 = pos(shape.indices[i + 2].v);
                // glm::normalize(glm::cross(v1 - v0, v2 - v0)): cross, then * (1 / sqrt(dot))
                const vec3 a(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z), b(v2.x - v0.x, v2.y - v0.y, v2.z - v0.z);
                const vec3 cr(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
                const float inv = 1.0f / std::sqrt((cr.x * cr.x + cr.y * cr.y) + cr.z * cr.z);
                const vec3 geometric(cr.x * inv, cr.y * inv, cr.z * inv);
                uvec3 tri;
                unsigned* t3[3] = { &tri.x, &tri.y, &tri.z };
                for (int j = 0; j < 3; j++) {
                    const ObjIndex& idx = shape.indices[i + size_t(j)];
                    Vertex vertex;
                    vertex.position = pos(idx.v);
                    if (idx.vn != -1 && !obj.vn.empty())
                        vertex.normal = vec3(obj.vn[3 * size_t(idx.vn)], obj.vn[3 * size_t(idx.vn) + 1], obj.vn[3 * size_t(idx.vn) + 2]);
                    else
                        vertex.normal = geometric;
                    if (idx.vt != -1 && !obj.vt.empty())
                        vertex.texCoord = vec2 { obj.vt[2 * size_t(idx.vt)], obj.vt[2 * size_t(idx.vt) + 1] };
                    VertexKey key;
                    std::memcpy(key.f, &vertex, sizeof key.f);
                    const auto it = cache.find(key);
                    if (it != cache.end()) {
                        *t3[j] = it->second;
                    } else {
                        *t3[j] = uint32_t(mesh.vertices.size());
                        cache.emplace(key, *t3[j]);
                        mesh.vertices.push_back(vertex);
                    }
                }
                mesh.triangles.push_back(tri);
            }
            const int materialId = shape.materialIds[start];
            if (materialId == -1) {
                mesh.material.kd = vec3(1.0f);
                mesh.material.ks = vec3(0.0f);
                mesh.material.shininess = 1.0f;
            } else {
                const ObjMaterial& m = obj.materials[size_t(materialId)];
                mesh.material.kd = vec3(m.diffuse[0], m.diffuse[1], m.diffuse[2]);
                if (!m.diffuseTexname.empty())
                    mesh.material.kdTexture = std::make_shared<Image>(loadImage(baseDir + m.diffuseTexname));
                mesh.material.ks = vec3(m.specular[0], m.specular[1], m.specular[2]);
                mesh.material.shininess = m.shininess;
                mesh.material.transparency = m.dissolve;
            }
            out.push_back(std::move(mesh));
            start = end;
        }
    }
    if (centerAndNormalize) { // framework/src/mesh.cpp:150-176: mean of all vertex positions, largest distance from it
        vec3 sum(0.0f);
        size_t n = 0;
        for (const Mesh& m : out)
            for (const Vertex& v : m.vertices)
                sum = vec3(sum.x + v.position.x, sum.y + v.position.y, sum.z + v.position.z), n++;
        const float fn = float(n);
        const vec3 center(sum.x / fn, sum.y / fn, sum.z / fn);
        float maxD = 0.0f;
        for (const Mesh& m : out)
            for (const Vertex& v : m.vertices) {
                const vec3 d(v.position.x - center.x, v.position.y - center.y, v.position.z - center.z);
                maxD = std::max(std::sqrt((d.x * d.x + d.y * d.y) + d.z * d.z), maxD);
            }
        for (Mesh& m : out)
            for (Vertex& v : m.vertices)
                v.position = vec3((v.position.x - center.x) / maxD, (v.position.y - center.y) / maxD, (v.position.z - center.z) / maxD);
    }
    return out;
}

inline std::optional<SceneType> deserializeSceneType(const std::string& name)
{
    std::string s;
    for (char c : name)
        s.push_back(char(std::tolower((unsigned char)c)));
    static const std::pair<const char*, SceneType> names[] = {
        { "single_triangle", SingleTriangle }, { "singletriangle", SingleTriangle }, { "single-triangle", SingleTriangle },
        { "cube", Cube }, { "cube-textured", CubeTextured }, { "cube_textured", CubeTextured }, { "cubetextured", CubeTextured },
        { "cornell_box", CornellBox }, { "cornellbox", CornellBox }, { "cornell-box", CornellBox },
        { "cornell_box_parallelogram_light", CornellBoxParallelogramLight }, { "cornellboxparallelogramlight", CornellBoxParallelogramLight },
        { "cornell-box-parallelogram-light", CornellBoxParallelogramLight }, { "monkey", Monkey }, { "teapot", Teapot },
        { "dragon", Dragon }, { "spheres", Spheres }, { "custom", Custom },
    };
    for (const auto& [text, type] : names)
        if (s == text)
            return type;
    return std::nullopt;
}

// The built-in scenes with their lights (reference src/scene.cpp:5-92).
inline Scene loadScenePrebuilt(SceneType type, const std::string& dataDir)
{
    const std::string dir = dataDir.empty() || dataDir.back() == '/' ? dataDir : dataDir + "/";
    Scene scene;
    auto add = [&](const char* file, bool normalize) {
        auto meshes = loadMesh(dir + file, normalize);
        for (auto& m : meshes)
            scene.meshes.push_back(std::move(m));
    };
    auto sphere = [](vec3 c, float r, vec3 kd) {
        Sphere s;
        s.center = c, s.radius = r, s.material.kd = kd;
        return s;
    };
    switch (type) {
    case SingleTriangle:
        add("triangle.obj", false);
        scene.meshes[0].material.kd = vec3(1.0f);
        scene.lights.emplace_back(PointLight { vec3(-1, 1, -1), vec3(1) });
        break;
    case Cube:
        add("cube.obj", false);
        scene.lights.emplace_back(SegmentLight { vec3(1.5f, 0.5f, -0.6f), vec3(-1, 0.5f, -0.5f), vec3(0.9f, 0.2f, 0.1f), vec3(0.2f, 1, 0.3f) });
        break;
    case CubeTextured:
        add("cube-textured.obj", false);
        scene.lights.emplace_back(PointLight { vec3(-1.0f, 1.5f, -1.0f), vec3(1) });
        break;
    case CornellBox:
        add("CornellBox-Mirror-Rotated.obj", true);
        scene.lights.emplace_back(PointLight { vec3(0, 0.58f, 0), vec3(1) });
        break;
    case CornellBoxParallelogramLight:
        add("CornellBox-Mirror-Rotated.obj", true);
        scene.lights.emplace_back(ParallelogramLight { vec3(-0.2f, 0.5f, 0), vec3(0.4f, 0, 0), vec3(0.0f, 0.0f, 0.4f), vec3(1, 0, 0), vec3(0, 1, 0),
            vec3(0, 0, 1), vec3(0, 1, 1) });
        break;
    case Monkey:
        add("monkey.obj", true);
        scene.lights.emplace_back(PointLight { vec3(-1, 1, -1), vec3(1) });
        scene.lights.emplace_back(PointLight { vec3(1, -1, -1), vec3(1) });
        break;
    case Teapot:
        add("teapot.obj", true);
        scene.lights.emplace_back(PointLight { vec3(-1, 1, -1), vec3(1) });
        break;
    case Dragon:
        add("dragon.obj", true);
        scene.lights.emplace_back(PointLight { vec3(-1, 1, -1), vec3(1) });
        break;
    case Spheres:
        scene.spheres.push_back(sphere(vec3(3.0f, -2.0f, 10.2f), 1.0f, vec3(0.8f, 0.2f, 0.2f)));
        scene.spheres.push_back(sphere(vec3(-2.0f, 2.0f, 4.0f), 2.0f, vec3(0.6f, 0.8f, 0.2f)));
        scene.spheres.push_back(sphere(vec3(0.0f, 0.0f, 6.0f), 0.75f, vec3(0.2f, 0.2f, 0.8f)));
        scene.lights.emplace_back(PointLight { vec3(3, 0, 3), vec3(15) });
        break;
    case Custom:
        add("custom.obj", false);
        scene.lights.emplace_back(PointLight { vec3(-1, 1, -1), vec3(1) });
        break;
    }
    return scene;
}

inline Scene loadSceneFromFile(const std::string& path, const std::vector<std::variant<PointLight, SegmentLight, ParallelogramLight>>& lights)
{
    Scene scene;
    scene.lights = lights;
    auto meshes = loadMesh(path);
    for (auto& m : meshes)
        scene.meshes.push_back(std::move(m));
    return scene;
}

// The flat scene file of include/cge_scene_file.h (what loadFlatScene reads), without a stored tree.
inline void saveFlatScene(const FlatScene& f, const std::string& path)
{
    FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp)
        throw std::runtime_error("cannot write " + path);
    struct Header {
        char magic[8];
        uint32_t n_meshes, n_vertices, n_triangles, n_spheres, n_lights, n_textures;
        uint64_t n_texels;
        uint32_t n_bvh_nodes, bvh_root, reserved[2];
    } h {};
    static_assert(sizeof(Header) == 56, "flat scene header");
    std::memcpy(h.magic, "CGESCN01", 8);
    h.n_meshes = uint32_t(f.meshes.size()), h.n_vertices = uint32_t(f.vertices.size()), h.n_triangles = uint32_t(f.triangles.size() / 3);
    h.n_spheres = uint32_t(f.spheres.size()), h.n_lights = uint32_t(f.lights.size()), h.n_textures = uint32_t(f.textures.size());
    h.n_texels = f.texels.size() / 3;
    bool ok = std::fwrite(&h, sizeof h, 1, fp) == 1;
    auto put = [&](const void* p, size_t bytes) { ok = ok && (bytes == 0 || std::fwrite(p, 1, bytes, fp) == bytes); };
    put(f.meshes.data(), f.meshes.size() * sizeof(cge_mesh_desc));
    put(f.vertices.data(), f.vertices.size() * sizeof(cge_vertex));
    put(f.triangles.data(), f.triangles.size() * sizeof(uint32_t));
    put(f.spheres.data(), f.spheres.size() * sizeof(cge_sphere_desc));
    put(f.lights.data(), f.lights.size() * sizeof(cge_light_desc));
    put(f.textures.data(), f.textures.size() * sizeof(cge_texture_desc));
    put(f.texels.data(), f.texels.size() * sizeof(float));
    ok = std::fclose(fp) == 0 && ok;
    if (!ok)
        throw std::runtime_error("short write: " + path);
}

} // namespace cge_engine
