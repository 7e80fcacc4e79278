// Test entry for host/cge_scene_io.hpp: load a scene with the mirrored reference loaders, flatten it and write the flat scene file.
//   scene_export prebuilt <SceneType number or name> <dataDir> <out.cges>
//   scene_export obj <file.obj> <centerAndNormalize 0|1> <out.cges>
#include "cge_scene_io.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>

using namespace cge_engine;

int main(int argc, char** argv)
{
    try {
        if (argc == 5 && std::string(argv[1]) == "prebuilt") {
            const std::string what = argv[2];
            SceneType type;
            if (!what.empty() && std::isdigit((unsigned char)what[0])) {
                type = SceneType(std::atoi(what.c_str()));
            } else {
                const auto t = deserializeSceneType(what);
                if (!t) {
                    std::fprintf(stderr, "unknown scene %s\n", what.c_str());
                    return 2;
                }
                type = *t;
            }
            saveFlatScene(flatten(loadScenePrebuilt(type, argv[3])), argv[4]);
            return 0;
        }
        if (argc == 5 && std::string(argv[1]) == "obj") {
            Scene scene;
            auto meshes = loadMesh(argv[2], std::atoi(argv[3]) != 0);
            for (auto& m : meshes)
                scene.meshes.push_back(std::move(m));
            saveFlatScene(flatten(scene), argv[4]);
            return 0;
        }
        if (argc == 3 && std::string(argv[1]) == "check") { // loadFlatScene on a (possibly corrupt) flat scene file
            const Scene s = loadFlatScene(argv[2]);
            std::printf("ok %zu meshes %zu spheres %zu lights\n", s.meshes.size(), s.spheres.size(), s.lights.size());
            return 0;
        }
        std::fprintf(stderr, "usage: scene_export prebuilt <type> <dataDir> <out> | obj <file> <0|1> <out> | check <file.cges>\n");
        return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "scene_export: %s\n", e.what());
        return 1;
    }
}
