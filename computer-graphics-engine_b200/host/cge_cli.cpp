// cge_cli — headless command-line rendering on the GPU path, in the shape of the reference's CLI mode
// (reference src/main.cpp:478-535: read the TOML config, load the scene, build the BvhInterface, render every camera from its
// own thread, write one BMP per camera) and of its config reader (src/config.cpp:193-374), on top of the mirrored interface
// (cge_engine.hpp).  SURVEY.md 8(f) N4.
//
//     cge_cli config.toml
//
// Keys, with the reference's names and defaults:
//     window_size = [800, 800]        data_path = "..."       scene = "cornell.cges"      output_dir = "out"
//     [features]        enable_shading / enable_recursive / enable_hard_shadow / enable_normal_interp /
//                       enable_texture_mapping / enable_accel_structure = false
//     [features.extra]  enable_bloom_effect / enable_multiple_rays_per_pixel = false   (any other extra flag set to true is
//                       refused by the library, CGE_ERR_UNSUPPORTED, exactly like through the C ABI)
//     [[cameras]]       field_of_view = 50.0, distance_from_look_at = 3.0, look_at = [0,0,0], rotation = [20,20,0]
//     [[lights]]        type = "point" (position, color) | "segment" (endpoints, colors) | "parallelogram" (corner, edges, colors)
// The keys the reference's reader lacks although the renderer has the switch (it leaves them to the GUI):
//     [features] enable_soft_shadow = false
//     [render]   ray_depth = 5, segment_light_samples = 25, parallelogram_light_samples = 5, rays_per_pixel_side = 3,
//                bloom_scalar = 0.3, bloom_threshold = 0.4, bloom_debug_option = 0, seed = 0, timestamp = true
// The scene is what the reference's config names (src/config.cpp:216-235): a built-in scene (number or name: its OBJ / MTL / PNG
// files are read from data_path by host/cge_scene_io.hpp, the loaders' bit-identical mirror) or an OBJ file in data_path - or a
// flat scene file (include/cge_scene_file.h).  Lights from the config replace a file's, as loadSceneFromFile does
// (src/scene.cpp:94-103); a built-in scene keeps its own.
// Only a TOML subset is read: tables, arrays of tables, booleans, numbers, basic strings, (nested, multi-line) arrays.
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "cge_engine.hpp"
#include "cge_scene_io.hpp"

using namespace cge_engine;

namespace {

// ---- TOML ------------------------------------------------------------------------------------------------------------------
// Tables, dotted keys, arrays of tables, inline tables, arrays over several lines, basic / literal / multi-line strings, booleans,
// integers (decimal, hex, octal, binary, underscores) and floats (exponents, inf, nan), comments.  Not read: dates and times (no
// key of the reference's config is one).
struct Value {
    enum Kind { None, Bool, Number, String, Array, Table } kind = None;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> tab;

    const Value* find(const std::string& key) const
    {
        auto it = tab.find(key);
        return it == tab.end() ? nullptr : &it->second;
    }
};

struct Parser {
    const std::string& t;
    size_t i = 0;
    int line = 1;
    explicit Parser(const std::string& text)
        : t(text)
    {
    }
    [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("config line " + std::to_string(line) + ": " + what); }
    void skipSpace(bool newlines)
    {
        while (i < t.size()) {
            if (t[i] == '#') {
                while (i < t.size() && t[i] != '\n')
                    i++;
            } else if (t[i] == '\n') {
                if (!newlines)
                    return;
                line++, i++;
            } else if (std::isspace(static_cast<unsigned char>(t[i]))) {
                i++;
            } else {
                return;
            }
        }
    }
    std::string key()
    {
        std::string k;
        if (i < t.size() && t[i] == '"') {
            for (i++; i < t.size() && t[i] != '"'; i++)
                k += t[i];
            i++;
        } else {
            while (i < t.size() && (std::isalnum(static_cast<unsigned char>(t[i])) || t[i] == '_' || t[i] == '-'))
                k += t[i++];
        }
        if (k.empty())
            fail("expected a key");
        return k;
    }
    std::vector<std::string> dottedKey()
    {
        std::vector<std::string> path { key() };
        for (skipSpace(false); i < t.size() && t[i] == '.'; skipSpace(false)) {
            i++;
            skipSpace(false);
            path.push_back(key());
        }
        return path;
    }
    Value value()
    {
        Value v;
        skipSpace(false);
        if (i >= t.size())
            fail("expected a value");
        if (t[i] == '[') {
            v.kind = Value::Array;
            i++;
            for (;;) {
                skipSpace(true);
                if (i < t.size() && t[i] == ']') {
                    i++;
                    break;
                }
                v.arr.push_back(value());
                skipSpace(true);
                if (i < t.size() && t[i] == ',')
                    i++;
                else if (i < t.size() && t[i] == ']') {
                    i++;
                    break;
                } else
                    fail("expected ',' or ']' in array");
            }
        } else if (t[i] == '{') { // inline table: { key = value, dotted.key = value }
            v.kind = Value::Table;
            i++;
            for (;;) {
                skipSpace(false);
                if (i < t.size() && t[i] == '}') {
                    i++;
                    break;
                }
                const std::vector<std::string> path = dottedKey();
                skipSpace(false);
                if (i >= t.size() || t[i] != '=')
                    fail("expected '=' in inline table");
                i++;
                Value& owner = descend(v, path, path.size() - 1);
                owner.tab[path.back()] = value();
                skipSpace(false);
                if (i < t.size() && t[i] == ',')
                    i++;
                else if (i < t.size() && t[i] == '}') {
                    i++;
                    break;
                } else
                    fail("expected ',' or '}' in inline table");
            }
        } else if (t[i] == '"' || t[i] == '\'') {
            // basic ("...", escapes) and literal ('...', verbatim) strings, single-line or multi-line (three quotes: a newline right
            // after the opening quotes is dropped, a backslash at the end of a line of a basic string joins the lines)
            v.kind = Value::String;
            const char q = t[i];
            const bool multi = t.compare(i, 3, std::string(3, q)) == 0;
            i += multi ? 3 : 1;
            if (multi && i < t.size() && t[i] == '\n')
                line++, i++;
            for (;;) {
                if (i >= t.size() || (!multi && t[i] == '\n'))
                    fail("unterminated string");
                if (multi ? t.compare(i, 3, std::string(3, q)) == 0 : t[i] == q) {
                    i += multi ? 3 : 1;
                    break;
                }
                if (q == '"' && t[i] == '\\' && i + 1 < t.size()) {
                    const char e = t[++i];
                    i++;
                    if (e == 'n')
                        v.str += '\n';
                    else if (e == 't')
                        v.str += '\t';
                    else if (e == 'r')
                        v.str += '\r';
                    else if (e == '\n') { // line-ending backslash: skip the newline and the next line's indentation
                        line++;
                        while (i < t.size() && std::isspace(static_cast<unsigned char>(t[i]))) {
                            if (t[i] == '\n')
                                line++;
                            i++;
                        }
                    } else
                        v.str += e; // \" \\ and anything else: the character itself
                    continue;
                }
                if (t[i] == '\n')
                    line++;
                v.str += t[i++];
            }
        } else if (t.compare(i, 4, "true") == 0) {
            v.kind = Value::Bool, v.b = true, i += 4;
        } else if (t.compare(i, 5, "false") == 0) {
            v.kind = Value::Bool, v.b = false, i += 5;
        } else {
            // integers and floats: optional sign, 0x / 0o / 0b prefixes, underscores between digits, exponents, inf, nan
            std::string tok;
            size_t j = i;
            while (j < t.size() && (std::isalnum(static_cast<unsigned char>(t[j])) || t[j] == '+' || t[j] == '-' || t[j] == '.' || t[j] == '_'))
                if (t[j++] != '_')
                    tok += t[j - 1];
            const size_t signLen = !tok.empty() && (tok[0] == '+' || tok[0] == '-') ? 1 : 0;
            const std::string body = tok.substr(signLen);
            const double sign = signLen && tok[0] == '-' ? -1.0 : 1.0;
            if (body == "inf") {
                v.num = sign * std::numeric_limits<double>::infinity();
            } else if (body == "nan") {
                v.num = std::numeric_limits<double>::quiet_NaN();
            } else if (body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'o' || body[1] == 'b')) {
                char* end = nullptr;
                v.num = sign * double(std::strtoull(body.c_str() + 2, &end, body[1] == 'x' ? 16 : body[1] == 'o' ? 8 : 2));
                if (*end)
                    fail("unsupported value");
            } else {
                char* end = nullptr;
                v.num = std::strtod(tok.c_str(), &end);
                if (tok.empty() || *end)
                    fail("unsupported value");
            }
            v.kind = Value::Number;
            i = j;
        }
        return v;
    }
    static Value& descend(Value& root, const std::vector<std::string>& path, size_t upto)
    {
        Value* cur = &root;
        for (size_t k = 0; k < upto; k++) {
            Value& next = cur->tab[path[k]];
            if (next.kind == Value::None)
                next.kind = Value::Table;
            cur = next.kind == Value::Array ? &next.arr.back() : &next; // [a.b] below [[a]] addresses its last element
        }
        return *cur;
    }
    Value parse()
    {
        Value root;
        root.kind = Value::Table;
        Value* cur = &root;
        for (skipSpace(true); i < t.size(); skipSpace(true)) {
            if (t[i] == '[') {
                const bool arrayOfTables = i + 1 < t.size() && t[i + 1] == '[';
                i += arrayOfTables ? 2 : 1;
                skipSpace(false);
                const std::vector<std::string> path = dottedKey();
                skipSpace(false);
                if (t.compare(i, arrayOfTables ? 2 : 1, arrayOfTables ? "]]" : "]") != 0)
                    fail("unterminated table header");
                i += arrayOfTables ? 2 : 1;
                Value& parent = descend(root, path, path.size() - 1);
                Value& slot = parent.tab[path.back()];
                if (arrayOfTables) {
                    slot.kind = Value::Array;
                    slot.arr.emplace_back();
                    slot.arr.back().kind = Value::Table;
                    cur = &slot.arr.back();
                } else {
                    if (slot.kind == Value::None)
                        slot.kind = Value::Table;
                    cur = &slot;
                }
            } else {
                const std::vector<std::string> path = dottedKey();
                skipSpace(false);
                if (i >= t.size() || t[i] != '=')
                    fail("expected '='");
                i++;
                Value& owner = descend(*cur, path, path.size() - 1);
                owner.tab[path.back()] = value();
            }
        }
        return root;
    }
};

// ---- typed access with the reference's defaults ---------------------------------------------------------------------------
const Value* at(const Value& v, std::initializer_list<const char*> path)
{
    const Value* cur = &v;
    for (const char* k : path) {
        if (!cur || cur->kind != Value::Table)
            return nullptr;
        cur = cur->find(k);
    }
    return cur;
}
bool getBool(const Value& v, std::initializer_list<const char*> path, bool fallback)
{
    const Value* x = at(v, path);
    return x && x->kind == Value::Bool ? x->b : fallback;
}
double getNumber(const Value& v, std::initializer_list<const char*> path, double fallback)
{
    const Value* x = at(v, path);
    return x && x->kind == Value::Number ? x->num : fallback;
}
std::string getString(const Value& v, std::initializer_list<const char*> path, const std::string& fallback)
{
    const Value* x = at(v, path);
    return x && x->kind == Value::String ? x->str : fallback;
}
vec3 toVec3(const Value* x, vec3 fallback)
{
    if (!x || x->kind != Value::Array || x->arr.size() < 3)
        return fallback;
    return vec3(float(x->arr[0].num), float(x->arr[1].num), float(x->arr[2].num));
}
vec3 vec3At(const Value& table, const char* key, size_t index, vec3 fallback)
{
    const Value* a = table.find(key);
    if (!a || a->kind != Value::Array || a->arr.size() <= index)
        return fallback;
    return toVec3(&a->arr[index], fallback);
}

struct CameraConfig { // src/config.h:16-21
    float fieldOfView = 50.0f, distanceFromLookAt = 3.0f;
    vec3 lookAt { 0.0f }, rotation { 20.0f, 20.0f, 0.0f };
};

std::string joinPath(const std::string& dir, const std::string& name)
{
    if (dir.empty() || name.empty() || name[0] == '/')
        return name;
    return dir.back() == '/' ? dir + name : dir + "/" + name;
}

} // namespace

int main(int argc, char** argv)
{
    // --print-config: read the file, print what was understood (the reference prints its Config the same way before it
    // renders, src/main.cpp:480 / src/config.cpp:68-123) and stop before any GPU work
    const bool printOnly = argc == 3 && std::string(argv[2]) == "--print-config";
    if (argc != 2 && !printOnly) {
        std::fprintf(stderr, "usage: %s config.toml [--print-config]\n", argv[0]);
        return 2;
    }
    try {
        std::ifstream in(argv[1]);
        if (!in)
            throw std::runtime_error(std::string("cannot read ") + argv[1]);
        std::stringstream buf;
        buf << in.rdbuf();
        const std::string text = buf.str();
        const Value cfg = Parser(text).parse();

        // ---- src/config.cpp:193-374 ---------------------------------------------------------------------------------------
        ivec2 windowSize { 800, 800 };
        if (const Value* ws = cfg.find("window_size"); ws && ws->kind == Value::Array && ws->arr.size() >= 2)
            windowSize = { int(ws->arr[0].num), int(ws->arr[1].num) };
        const std::string dataPath = getString(cfg, { "data_path" }, ".");
        // scene (src/config.cpp:216-235): a SceneType number, one of the built-in scenes' names, or a file in data_path - an OBJ
        // file as in the reference, or a flat scene file (include/cge_scene_file.h)
        std::string sceneName = getString(cfg, { "scene" }, "none");
        std::optional<SceneType> sceneType;
        if (const Value* sv = cfg.find("scene"); sv && sv->kind == Value::Number)
            sceneType = SceneType(int(sv->num));
        else
            sceneType = deserializeSceneType(sceneName);
        static const char* const kSceneNames[] = { "single_triangle", "cube", "cube_textured", "cornell_box", "cornell_box_parallelogram_light",
            "monkey", "teapot", "dragon", "spheres", "custom" };
        if (sceneType && (int(*sceneType) < 0 || int(*sceneType) > int(Custom)))
            throw std::runtime_error("scene number out of range");
        if (sceneType)
            sceneName = kSceneNames[int(*sceneType)]; // (serialize(), src/config.cpp:376-402: names the output bitmaps)
        const std::string outputDir = getString(cfg, { "output_dir" }, ".");
        Features features;
        features.enableShading = getBool(cfg, { "features", "enable_shading" }, false);
        features.enableRecursive = getBool(cfg, { "features", "enable_recursive" }, false);
        features.enableHardShadow = getBool(cfg, { "features", "enable_hard_shadow" }, false);
        features.enableSoftShadow = getBool(cfg, { "features", "enable_soft_shadow" }, false);
        features.enableNormalInterp = getBool(cfg, { "features", "enable_normal_interp" }, false);
        features.enableTextureMapping = getBool(cfg, { "features", "enable_texture_mapping" }, false);
        features.enableAccelStructure = getBool(cfg, { "features", "enable_accel_structure" }, false);
        features.extra.enableBloomEffect = getBool(cfg, { "features", "extra", "enable_bloom_effect" }, false);
        features.extra.enableMultipleRaysPerPixel = getBool(cfg, { "features", "extra", "enable_multiple_rays_per_pixel" }, false);
        features.extra.enableMotionBlur = getBool(cfg, { "features", "extra", "enable_motion_blur" }, false);
        features.extra.enableDepthOfField = getBool(cfg, { "features", "extra", "enable_depth_of_field" }, false);
        features.extra.enableGlossyReflection = getBool(cfg, { "features", "extra", "enable_glossy_reflection" }, false);
        features.extra.enableEnvironmentMapping = getBool(cfg, { "features", "extra", "enable_environment_mapping" }, false);
        features.extra.enableBilinearTextureFiltering = getBool(cfg, { "features", "extra", "enable_bilinear_texture_filtering" }, false);
        features.extra.enableMipmapTextureFiltering = getBool(cfg, { "features", "extra", "enable_mipmap_texture_filtering" }, false);

        std::vector<CameraConfig> cameras;
        if (const Value* cams = cfg.find("cameras"); cams && cams->kind == Value::Array)
            for (const Value& c : cams->arr) {
                CameraConfig cc;
                cc.fieldOfView = float(getNumber(c, { "field_of_view" }, 50.0));
                cc.distanceFromLookAt = float(getNumber(c, { "distance_from_look_at" }, 3.0));
                cc.lookAt = toVec3(c.find("look_at"), vec3(0.0f));
                cc.rotation = toVec3(c.find("rotation"), vec3(20.0f, 20.0f, 0.0f));
                cameras.push_back(cc);
            }

        // the renderer's globals (src/light.cpp:12-13, src/render.cpp:14,19-21) and the depth literal (src/render.cpp:318)
        const int rayDepth = int(getNumber(cfg, { "render", "ray_depth" }, 5));
        segmentLightSamples = int(getNumber(cfg, { "render", "segment_light_samples" }, 25));
        parallelogramLightDirectionSamples = int(getNumber(cfg, { "render", "parallelogram_light_samples" }, 5));
        raysPerPixelSide = int(getNumber(cfg, { "render", "rays_per_pixel_side" }, 3));
        bloomScalar = float(getNumber(cfg, { "render", "bloom_scalar" }, 0.3));
        bloomThreshold = float(getNumber(cfg, { "render", "bloom_threshold" }, 0.4));
        bloomDebugOption = int(getNumber(cfg, { "render", "bloom_debug_option" }, 0));
        samplerSeed = uint32_t(getNumber(cfg, { "render", "seed" }, 0));
        const bool timestamp = getBool(cfg, { "render", "timestamp" }, true);

        if (printOnly) {
            std::printf("window_size %d %d\nscene %s%s\ndata_path %s\noutput_dir %s\n", windowSize.x, windowSize.y, sceneName.c_str(),
                sceneType ? " (built-in)" : "", dataPath.c_str(), outputDir.c_str());
            std::printf("features shading %d recursive %d hard_shadow %d soft_shadow %d normal_interp %d texture_mapping %d accel_structure %d\n",
                features.enableShading, features.enableRecursive, features.enableHardShadow, features.enableSoftShadow,
                features.enableNormalInterp, features.enableTextureMapping, features.enableAccelStructure);
            std::printf("extra bits 0x%x\n", featureBits(features) >> 16);
            std::printf("render ray_depth %d segment_light_samples %d parallelogram_light_samples %d rays_per_pixel_side %d bloom %g %g %d seed %u\n",
                rayDepth, segmentLightSamples, parallelogramLightDirectionSamples, raysPerPixelSide, bloomScalar, bloomThreshold,
                bloomDebugOption, samplerSeed);
            for (const CameraConfig& c : cameras)
                std::printf("camera fov %g dist %g look_at %g %g %g rotation %g %g %g\n", c.fieldOfView, c.distanceFromLookAt, c.lookAt.x,
                    c.lookAt.y, c.lookAt.z, c.rotation.x, c.rotation.y, c.rotation.z);
            if (const Value* lights = cfg.find("lights"); lights && lights->kind == Value::Array)
                for (const Value& l : lights->arr)
                    std::printf("light %s\n", getString(l, { "type" }, "none").c_str());
            return 0;
        }

        // ---- src/main.cpp:478-535 -----------------------------------------------------------------------------------------
        // lights of the config (src/config.cpp:336-372); they replace a scene FILE's lights (loadSceneFromFile, src/scene.cpp:94-103), a
        // built-in scene keeps its own (src/main.cpp:491-500)
        std::vector<std::variant<PointLight, SegmentLight, ParallelogramLight>> cfgLights;
        bool haveCfgLights = false;
        if (const Value* lights = cfg.find("lights"); lights && lights->kind == Value::Array) {
            haveCfgLights = true;
            for (const Value& l : lights->arr) {
                const std::string type = getString(l, { "type" }, "none");
                const vec3 zero(0.0f);
                if (type == "point") {
                    cfgLights.emplace_back(PointLight { toVec3(l.find("position"), zero), toVec3(l.find("color"), zero) });
                } else if (type == "segment") {
                    cfgLights.emplace_back(SegmentLight { vec3At(l, "endpoints", 0, zero), vec3At(l, "endpoints", 1, zero),
                        vec3At(l, "colors", 0, zero), vec3At(l, "colors", 1, zero) });
                } else if (type == "parallelogram") {
                    cfgLights.emplace_back(ParallelogramLight { toVec3(l.find("corner"), zero), vec3At(l, "edges", 0, zero),
                        vec3At(l, "edges", 1, zero), vec3At(l, "colors", 0, zero), vec3At(l, "colors", 1, zero), vec3At(l, "colors", 2, zero),
                        vec3At(l, "colors", 3, zero) });
                } else {
                    std::fprintf(stderr, "Unknown light type: %s -- Skip\n", type.c_str());
                }
            }
        }
        Scene scene;
        const bool isObj = sceneName.size() > 4 && sceneName.compare(sceneName.size() - 4, 4, ".obj") == 0;
        if (sceneType) {
            scene = loadScenePrebuilt(*sceneType, dataPath);
        } else if (isObj) {
            scene = loadSceneFromFile(joinPath(dataPath, sceneName), cfgLights);
        } else {
            scene = loadFlatScene(joinPath(dataPath, sceneName));
            if (haveCfgLights)
                scene.lights = cfgLights;
            else
                std::fprintf(stderr, "WARN: No lights found in config file, keeping the scene file's.\n");
        }
        std::string stem = sceneName.substr(sceneName.find_last_of('/') + 1);
        stem = stem.substr(0, stem.find_last_of('.'));

        BvhInterface bvh { &scene, features };
        if (std::system(("mkdir -p '" + outputDir + "'").c_str()) != 0)
            throw std::runtime_error("cannot create " + outputDir);
        const auto start = std::chrono::high_resolution_clock::now();
        char stamp[64] = "";
        if (timestamp) {
            const std::time_t now = std::time(nullptr);
            std::strftime(stamp, sizeof(stamp), "_%Y-%m-%d-%H:%M:%S", std::localtime(&now));
        }
        const float deg = 0.01745329251994329576923690768489f; // glm::radians
        std::vector<std::thread> workers;
        std::vector<std::string> errors(cameras.size());
        for (size_t i = 0; i < cameras.size(); i++)
            workers.emplace_back([&, i]() {
                try {
                    const CameraConfig& cc = cameras[i];
                    Trackball camera { float(windowSize.x) / float(windowSize.y), cc.fieldOfView * deg, cc.distanceFromLookAt };
                    camera.setCamera(cc.lookAt, vec3(cc.rotation.x * deg, cc.rotation.y * deg, cc.rotation.z * deg), cc.distanceFromLookAt);
                    const std::string path = joinPath(outputDir, stem + stamp + "_cam_" + std::to_string(i) + ".bmp");
                    renderRayTracingToBitmap(scene, camera, bvh, windowSize, features, path, rayDepth);
                    std::printf("Image %zu saved to %s\n", i, path.c_str());
                } catch (const std::exception& e) {
                    errors[i] = e.what();
                }
            });
        for (auto& w : workers)
            w.join();
        for (const std::string& e : errors)
            if (!e.empty())
                throw std::runtime_error(e);
        const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start).count();
        std::printf("Rendering took %lld ms, %zu images rendered.\n", static_cast<long long>(ms), cameras.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
