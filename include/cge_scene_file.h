/*
 * cge_scene_file.h — on-disk form of a cge_scene_desc (".cges" files).
 *
 * A flat scene file is nothing more than the arrays of cge_scene_desc (include/cge.h) written back to back
 * after a fixed header, little-endian, no padding between arrays.  It exists so that the reference engine's
 * own loaders (tinyobj + centerAndScaleToUnitMesh + stb_image, framework/src/mesh.cpp:52-176,
 * framework/src/image.cpp:12-35) run ONCE, in the build container, and the exact Scene they produced
 * (src/scene.h:28-33) can then be fed bit-for-bit to the CUDA path, to the CPU oracle and to the reference
 * harness on a machine that has no reference checkout.
 *
 * Layout:  cge_scene_file_header
 *          cge_mesh_desc   [n_meshes]
 *          cge_vertex      [n_vertices]
 *          uint32_t        [3 * n_triangles]
 *          cge_sphere_desc [n_spheres]
 *          cge_light_desc  [n_lights]
 *          cge_texture_desc[n_textures]
 *          float           [3 * n_texels]
 *          cge_bvh_node    [n_bvh_nodes]                     (optional: the reference's own tree, for tests)
 *          uint32_t        [n_triangles + n_spheres]         (only when n_bvh_nodes != 0: leaf order)
 */
#ifndef CGE_SCENE_FILE_H_
#define CGE_SCENE_FILE_H_

#include "cge.h"

#define CGE_SCENE_FILE_MAGIC "CGESCN01"

typedef struct cge_scene_file_header {
    char magic[8];
    uint32_t n_meshes, n_vertices, n_triangles, n_spheres, n_lights, n_textures;
    uint64_t n_texels;
    uint32_t n_bvh_nodes, bvh_root;
    uint32_t reserved[2];
} cge_scene_file_header;

#endif
