// Headless render entry written against the mirrored reference interface (computer-graphics-engine_b200/host/
// cge_engine.hpp): the shape of reference src/main.cpp:478-535 (CLI mode) without GLFW.  Used by
// tests/test_gpu_dropin.py:   drop_in_main scene.cges W H fov_deg dist rotx_deg roty_deg featureBits depth paraSamples out.raw
#include <cstdio>
#include <cstdlib>
#include <string>

#include "cge_engine.hpp"

using namespace cge_engine;

int main(int argc, char** argv)
{
    if (argc < 12) {
        std::fprintf(stderr, "usage: %s scene W H fov dist rotx roty features depth paraSamples out.raw\n", argv[0]);
        return 2;
    }
    try {
        Scene scene = loadFlatScene(argv[1]);
        const ivec2 res { std::atoi(argv[2]), std::atoi(argv[3]) };
        const float deg = 0.01745329251994329576923690768489f; // glm::radians
        const uint32_t bits = uint32_t(std::strtoul(argv[8], nullptr, 0));
        Features features;
        features.enableShading = bits & 1, features.enableRecursive = bits & 2, features.enableHardShadow = bits & 4;
        features.enableSoftShadow = bits & 8, features.enableNormalInterp = bits & 16, features.enableTextureMapping = bits & 32;
        features.enableAccelStructure = bits & 64;
        parallelogramLightDirectionSamples = std::atoi(argv[10]);
        BvhInterface bvh { &scene, features };
        Trackball camera { float(res.x) / float(res.y), float(std::atof(argv[4])) * deg, float(std::atof(argv[5])) };
        camera.setCamera(vec3(0.0f), vec3(float(std::atof(argv[6])) * deg, float(std::atof(argv[7])) * deg, 0.0f), float(std::atof(argv[5])));
        Screen screen { res, false };
        cge_stats st {};
        renderRayTracing(scene, camera, bvh, screen, features, std::atoi(argv[9]), &st);
        // single debug ray through the screen centre, like the reference's ray debugger (src/main.cpp:398)
        Ray ray;
        ray.origin = camera.position();
        const vec3 c = getFinalColor(scene, bvh, ray, features, std::atoi(argv[9]));
        FILE* f = std::fopen(argv[11], "wb");
        std::fwrite(screen.pixels().data(), sizeof(vec3), screen.pixels().size(), f);
        std::fclose(f);
        std::printf("levels %d leaves %d kernel_ms %.3f rays %llu debug_ray %g %g %g\n", bvh.numLevels(), bvh.numLeaves(), st.kernel_ms,
            (unsigned long long)(st.primary_rays + st.bounce_rays + st.shadow_rays), c.x, c.y, c.z);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
