/*
 * cge.h — C ABI of the B200-native ray-tracing hot path ("cge" = Computer-Graphics-Engine).
 *
 * This is the only boundary between host code and the CUDA implementation.  It replaces, for the
 * reference engine (Anton-Kalpakchiev/Computer-Graphics-Engine), the body of
 *
 *     void renderRayTracing(const Scene&, const Trackball&, const BvhInterface&, Screen&, const Features&)
 *                                                                        (reference src/render.h:32, src/render.cpp:273-329)
 *
 * and everything that call reaches: getFinalColor / recursiveRayTrace (src/render.cpp:27-155),
 * BoundingVolumeHierarchy::intersect (src/bounding_volume_hierarchy.cpp:299-427), the prebuilt
 * libIntersect functions (src/intersect.h:5-16), computeLightContribution / testVisibilityLightSample /
 * sample*Light (src/light.cpp:19-164), computeShading / computeReflectionRay (src/shading.cpp:7-62),
 * computeBarycentricCoord / interpolateNormal / interpolateTexCoord (src/interpolate.cpp:4-28),
 * acquireTexel nearest (src/texture.cpp:15-27) and Trackball::generateRay (framework/src/trackball.cpp:101-110).
 *
 * Rules of the boundary: extern "C", POD only, plain pointers and sizes, int status codes, no exceptions,
 * no STL / glm / torch types.  Host arrays passed to cge_scene_create are BORROWED for the duration of the
 * call and copied to HBM; the library owns all device memory.  There is no CPU fallback: if no CUDA device
 * is usable every entry point returns CGE_ERR_CUDA.
 */
#ifndef CGE_H_
#define CGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGE_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------------------------------------ */
enum {
    CGE_OK = 0,
    CGE_ERR_INVALID_ARG = 1,  /* null pointer, bad size, index out of range                                */
    CGE_ERR_UNSUPPORTED = 2,  /* an ExtraFeatures flag (src/common.h:54-65) or a non-terminating material  */
    CGE_ERR_CUDA = 3,         /* CUDA runtime error or no device; text in cge_last_error()                 */
    CGE_ERR_NCCL = 4,         /* NCCL error in the multi-GPU gather                                        */
    CGE_ERR_NOMEM = 5
};

/* ---- Features (reference src/common.h:67-77), one bit per bool, same order ---------------------------- */
enum {
    CGE_FEAT_SHADING = 1u << 0,         /* enableShading          */
    CGE_FEAT_RECURSIVE = 1u << 1,       /* enableRecursive        */
    CGE_FEAT_HARD_SHADOW = 1u << 2,     /* enableHardShadow       */
    CGE_FEAT_SOFT_SHADOW = 1u << 3,     /* enableSoftShadow       */
    CGE_FEAT_NORMAL_INTERP = 1u << 4,   /* enableNormalInterp     */
    CGE_FEAT_TEXTURE_MAPPING = 1u << 5, /* enableTextureMapping   */
    CGE_FEAT_ACCEL_STRUCTURE = 1u << 6, /* enableAccelStructure   */
    /* ExtraFeatures (src/common.h:54-65) occupy bits 16..25 in declaration order.  Two are implemented (SURVEY.md 8f N2:
       the ones that are per-pixel multipliers of the same kernels); any OTHER extra bit makes cge_render return
       CGE_ERR_UNSUPPORTED instead of silently rendering something else. */
    CGE_FEAT_BLOOM_EFFECT = 1u << 19,            /* extra.enableBloomEffect           (src/render.cpp:158-210,326-328) */
    CGE_FEAT_MULTIPLE_RAYS_PER_PIXEL = 1u << 22, /* extra.enableMultipleRaysPerPixel  (src/render.cpp:211-227,295-303) */
    CGE_FEAT_EXTRA_SUPPORTED = (1u << 19) | (1u << 22),
    CGE_FEAT_EXTRA_MASK = 0x03ff0000u
};

/* ---- scene description (what Scene, src/scene.h:28-33, flattens to) ---------------------------------- */

/* Same 32-byte layout as the reference Vertex (framework/include/framework/mesh.h:14-20). */
typedef struct cge_vertex {
    float position[3];
    float normal[3];
    float texcoord[2];
} cge_vertex;

/* One per Scene::meshes entry (mesh.h:36-43) with its Material (mesh.h:22-34). */
typedef struct cge_mesh_desc {
    uint32_t vertex_offset;   /* first vertex of this mesh in cge_scene_desc::vertices   */
    uint32_t vertex_count;
    uint32_t triangle_offset; /* first triangle of this mesh in cge_scene_desc::triangles */
    uint32_t triangle_count;
    float kd[3];
    float ks[3];
    float shininess;
    float transparency;
    int32_t texture_id; /* index into textures[], -1 when Material::kdTexture is empty */
    uint32_t reserved[3];
} cge_mesh_desc;

/* Sphere (src/common.h:31-35). */
typedef struct cge_sphere_desc {
    float center[3];
    float radius;
    float kd[3];
    float ks[3];
    float shininess;
    float transparency;
    int32_t texture_id;
    uint32_t reserved;
} cge_sphere_desc;

enum { CGE_LIGHT_POINT = 0, CGE_LIGHT_SEGMENT = 1, CGE_LIGHT_PARALLELOGRAM = 2 };

/* Tagged union of PointLight / SegmentLight / ParallelogramLight (src/common.h:37-52); v[] holds the
 * members in declaration order:  point: position,color (6) · segment: endpoint0,endpoint1,color0,color1 (12)
 * · parallelogram: v0,edge01,edge02,color0,color1,color2,color3 (21). */
typedef struct cge_light_desc {
    uint32_t type;
    float v[21];
} cge_light_desc;

/* Image (framework/include/framework/image.h:13-20): float RGB texels in [0,1], row-major, row 0 first. */
typedef struct cge_texture_desc {
    int32_t width;
    int32_t height;
    uint64_t texel_offset; /* in texels (3 floats each) into cge_scene_desc::texels */
} cge_texture_desc;

/* Optional caller-supplied BVH in the reference's own node order (src/bounding_volume_hierarchy.h:31-41):
 * when bvh_nodes == NULL the library rebuilds the identical tree itself (median split, std::nth_element on
 * centroid axis depth%3, MAX_DEPTH 16 — src/bounding_volume_hierarchy.cpp:74-78,130-147). */
typedef struct cge_bvh_node {
    float lower[3];
    float upper[3];
    uint32_t is_leaf;
    uint32_t depth;
    uint32_t beg, end;    /* primitive range in bvh_prim_order */
    uint32_t left, right; /* child node indices (inner nodes only) */
} cge_bvh_node;

typedef struct cge_scene_desc {
    uint32_t n_meshes, n_vertices, n_triangles, n_spheres, n_lights, n_textures;
    uint64_t n_texels;
    const cge_mesh_desc* meshes;
    const cge_vertex* vertices;
    const uint32_t* triangles; /* 3 mesh-local vertex indices per triangle (Mesh::triangles) */
    const cge_sphere_desc* spheres;
    const cge_light_desc* lights;
    const cge_texture_desc* textures;
    const float* texels;
    /* optional reference-built BVH; primitive ids are "global primitive ids": triangles of mesh 0, mesh 1, …
       in Mesh::triangles order, then spheres (src/bounding_volume_hierarchy.cpp:158-172). */
    uint32_t n_bvh_nodes;
    uint32_t bvh_root;
    const cge_bvh_node* bvh_nodes;
    const uint32_t* bvh_prim_order; /* n_triangles + n_spheres global primitive ids in leaf order */
} cge_scene_desc;

/* ---- camera: the ray-independent part of Trackball::generateRay (framework/src/trackball.cpp:101-110) -- */
typedef struct cge_camera {
    float origin[3]; /* Trackball::position()                                  (trackball.cpp:71-74) */
    float quat[4];   /* glm::quat(m_rotationEulerAngles) as w,x,y,z            (type_quat.inl:208-217) */
    float half_width;  /* m_halfScreenSpaceWidth  = aspect * tan(fovy/2)       (trackball.cpp:26-27) */
    float half_height; /* m_halfScreenSpaceHeight = tan(fovy/2) */
} cge_camera;

/* Convenience: fill a cge_camera exactly as the reference Trackball would (same float op order).
 * fovy and rotation in radians; aspect = float(W)/float(H) (framework/src/window.cpp:379-384). */
int cge_camera_from_trackball(float fovy, float aspect, const float look_at[3], float dist,
                              const float rotation_euler[3], cge_camera* out);

/* ---- render parameters: every global the reference reads on this path becomes a field ------------------ */
enum {
    CGE_TRAVERSAL_REFERENCE = 0, /* exhaustive DFS, right child first, no t culling: visit order and tie
                                    rule of src/bounding_volume_hierarchy.cpp:312-361 reproduced literally */
    CGE_TRAVERSAL_FAST = 1       /* binned-SAH tree (<= 4 primitives per leaf), near child first, conservative t culling,
                                    shadow rays stop at the first blocker; equal-t winner chosen by the reference's
                                    visit rank; with area lights a wavefront pipeline over warp-compacted hit queues, otherwise
                                    one thread per pixel.  Spheres (up to 64) are tested beside the tree, each behind the chain
                                    of reference-tree boxes that decides whether the reference calls its sphere test; frames
                                    without enableAccelStructure walk the same tree with the tie rank of the reference's
                                    primitive vector.  Scenes with more spheres take the literal traversal. */
};
enum {
    CGE_SAMPLER_HASH = 0 /* rand() replaced by hash(seed, pixel, draw index) — see DESIGN.md "sampler" */
};
enum {
    CGE_FLAG_WANT_PRIM_IDS = 1u << 0,  /* fill prim_id_out */
    CGE_FLAG_RGB_DEVICE_PTR = 1u << 1, /* rgb_out / prim_id_out are device pointers on the calling GPU: no D2H copy (a caller that
                                          keeps the frame in HBM, or kernel-only timing) */
    CGE_FLAG_COUNT_TESTS = 1u << 2,    /* fill cge_stats::box_tests / tri_tests (with CGE_TRAVERSAL_REFERENCE they equal the
                                          reference's own intersectRayWithShape / intersectRayWithTriangle call counts) */
    CGE_FLAG_OUTPUT_RGBA8 = 1u << 3,   /* rgb_out is a uint8_t[W*H*4] RGBA host buffer: the frame goes through the output stage of
                                          Screen::writeBitmapToFile (src/screen.cpp:49-60: clamp to [0,1], *255, truncate, alpha 255;
                                          NaN -> 0) on the GPU, so the D2H copy is 4 instead of 12 bytes per pixel */
    CGE_FLAG_PARTITION_TILE_ROWS = 1u << 4, /* part_index / part_count deal whole tile rows (runs of 4 complete image rows) instead
                                          of single 8x4 tiles; cge_render_distributed always partitions this way */
    CGE_FLAG_SHARED_HOST_FRAME = 1u << 5,   /* cge_render_distributed: rgb_out is the frame cge_comm_host_frame returned (the same
                                          on every rank); each rank copies the rows it rendered straight into it over its own
                                          PCIe link instead of rank 0 gathering the frame and copying all of it */
    CGE_FLAG_PEER_FRAME = 1u << 6,          /* cge_render_distributed: rgb_out is the pointer cge_comm_peer_frame returned on this rank
                                          (rank 0's device frame, mapped into every rank): the kernels of every rank store their
                                          pixels straight into it over NVLink - no gather, no staging buffer, no unpack */
    CGE_FLAG_DYNAMIC_TILES = 1u << 7        /* cge_render_distributed with CGE_FLAG_PEER_FRAME: part of every rank's tile rows forms a pool
                                          of chunks dealt at run time - by an atomic counter beside rank 0's frame, over NVLink - to
                                          whichever GPU has finished its own rows first (the reference deals its rows the same way:
                                          src/render.cpp:277-280, schedule(guided)) */
};
/* Development switches (A/B measurements and the tests that prove both production pipelines bit-identical); not needed by a
 * caller, every setting renders the same frame. */
enum {
    CGE_DEV_FLAG_PER_THREAD = 1u << 16,  /* CGE_TRAVERSAL_FAST: one thread per pixel even for area-light frames (default there: wavefront) */
    CGE_DEV_FLAG_WAVEFRONT = 1u << 17,   /* CGE_TRAVERSAL_FAST: wavefront pipeline even for point-light frames (default there: per thread) */
    CGE_DEV_FLAG_DEBUG_CYCLES = 1u << 18 /* prim_id_out receives each pixel's cost in SM cycles >> 4 (implies the per-thread kernel) */
};

typedef struct cge_params {
    int32_t width, height;      /* Screen::resolution()                       (src/screen.h:18)        */
    uint32_t features;          /* CGE_FEAT_* bits                                                       */
    int32_t ray_depth;          /* rayDepth of getFinalColor; the reference passes the literal 5
                                   (src/render.cpp:298,308,318)                                          */
    int32_t segment_samples;    /* segmentLightSamples               = 25    (src/light.cpp:12)         */
    int32_t parallelogram_samples; /* parallelogramLightDirectionSamples = 5 (src/light.cpp:13)         */
    uint32_t sampler;           /* CGE_SAMPLER_*                                                          */
    uint32_t seed;
    uint32_t traversal;         /* CGE_TRAVERSAL_*                                                        */
    uint32_t flags;             /* CGE_FLAG_*                                                             */
    /* image partition for multi-GPU: this call renders only tiles t with t % part_count == part_index
       (tiles are 8x4 pixels, numbered row-major).  part_count <= 1 renders the whole frame. */
    uint32_t part_index, part_count;
    /* the globals the two implemented ExtraFeatures read; ignored unless the feature bit is set */
    int32_t rays_per_pixel_side;  /* raysPerPixelSide = 3   (src/render.cpp:14; GUI range 1..10, src/main.cpp:195)   */
    float bloom_scalar;           /* bloomScalar = .3f      (src/render.cpp:19)                                      */
    float bloom_threshold;        /* bloomThreshold = .4f   (src/render.cpp:20)                                      */
    int32_t bloom_debug_option;   /* bloomDebugOption = 0   (src/render.cpp:21): 0 image + bloom, 1 bloom only, 2 image */
} cge_params;

typedef struct cge_stats {
    uint64_t primary_rays;    /* camera rays traced                                   */
    uint64_t bounce_rays;     /* unique reflection rays traced on the GPU             */
    uint64_t shadow_rays;     /* unique shadow rays traced on the GPU                 */
    uint64_t reference_rays;  /* BvhInterface::intersect calls the reference would have made for this frame
                                 (duplicate reflection subtrees counted, src/render.cpp:100,118) */
    uint64_t box_tests, tri_tests; /* only filled with CGE_FLAG_COUNT_TESTS */
    uint64_t reference_shadow_rays; /* the part of reference_rays made from testVisibilityLightSample (src/light.cpp:61) */
    float kernel_ms;          /* device time of the render kernels (CUDA events on the call's stream) */
    float total_ms;           /* device time including uploads of camera/params and the D2H copy     */
    uint32_t kernel_launches; /* kernels launched by this call                                        */
    float stage_ms[4];        /* wavefront pipeline: wf_chain / wf_visibility / wf_shade / wf_fold device times */
    uint64_t shadow_samples_culled; /* light samples (shadow rays and samples that need none) of hits the light-hull pre-pass proved
                                       unoccluded: settled without a ray, not counted in shadow_rays (ABI 4) */
    float vis_cull_ms;        /* the part of stage_ms[1] spent in that pre-pass (0 when it did not run or ran beside the chain stage) */
} cge_stats;

typedef struct cge_scene cge_scene; /* opaque: device-resident flattened scene + BVH on ONE GPU */

/* ---- development switches (environment variables read at call time; A/B measurements and tests only, not part of the ABI:
 *      every setting produces the same frame bit for bit) -----------------------------------------------------------------
 *   CGE_ZERO_SHADING_CULL=0   trace the shadow ray of light samples whose Phong term is exactly zero as well
 *   CGE_VIS_CULL=0 | 1        wavefront pipeline without / with the light-hull pre-pass whatever the launch size (default: from
 *                             1.5 Mpixel per launch); CGE_CULL_BUDGET=n: inner nodes one hull walk may visit (96)
 *   CGE_CHAIN_SPLIT=0 | 1     chain stage as one kernel / as camera rays + continuation beside the level-0 shadow rays
 *   CGE_DYNAMIC_POOL_PCT=n, CGE_DYNAMIC_CHUNKS=n   CGE_FLAG_DYNAMIC_TILES: share of every rank's tile rows in the pool (25), chunks per rank (2)
 *   CGE_BANDS=n               number of concurrent bands a frame / rank partition is rendered in (1 = one pipeline)
 *   CGE_REGROUP=0             shadow pass without the in-warp regrouping (wf_vis_grouped_kernel)
 *   CGE_SAH_BUILD=host        build the FAST traversal tree with the host builder instead of the GPU builder
 *   CGE_TIMING=1              print the GPU tree builder's phases on stderr */

/* ---- entry points ------------------------------------------------------------------------------------- */

int cge_abi_version(void);
const char* cge_last_error(void); /* thread-local message of the last failing call on this thread */

/* Number of usable CUDA devices (0 => every other call fails with CGE_ERR_CUDA). */
int cge_device_count(void);

/* Flatten + upload.  device = CUDA ordinal.  Builds the reference-order BVH unless desc->bvh_nodes is given. */
int cge_scene_create(const cge_scene_desc* desc, int device, cge_scene** out);
/* GUI edits lights every frame (reference src/main.cpp:290-368): replace the light list only. */
int cge_scene_update_lights(cge_scene* scene, const cge_light_desc* lights, uint32_t n_lights);
int cge_scene_destroy(cge_scene* scene);

/* Introspection used by BvhInterface::numLevels / numLeaves (src/bvh_interface.h:20,24) and by the tests. */
int cge_scene_bvh_info(const cge_scene* scene, uint32_t* n_nodes, uint32_t* n_levels, uint32_t* n_leaves,
                       uint32_t* max_leaf_prims);
/* Host-only (no GPU needed): rebuild the reference's tree for a scene description — BoundingVolumeHierarchy's constructor
 * (src/bounding_volume_hierarchy.cpp:149-194) without the device upload.  nodes_out has room for *n_nodes_inout nodes; on
 * return *n_nodes_inout is the node count (CGE_ERR_INVALID_ARG if the room was too small).  prim_order_out: n_triangles +
 * n_spheres entries. */
int cge_bvh_build_reference_order(const cge_scene_desc* desc, cge_bvh_node* nodes_out, uint32_t* n_nodes_inout,
                                  uint32_t* prim_order_out, uint32_t* root_out, uint32_t* n_levels_out, uint32_t* n_leaves_out);
/* Host-only (no GPU needed): the check cge_scene_create applies to a caller-supplied tree (desc->bvh_nodes, bvh_prim_order,
 * bvh_root).  The input is untrusted (flat scene files carry stored trees): bvh_prim_order must be a permutation, the nodes
 * reachable from the root must form a tree (each reached once: no cycles, no shared subtrees), the children of an inner node
 * must split its primitive range, the root must cover every primitive.  The depth reported - and used to size the device
 * traversal stack - is the one measured by the walk, not the depth field.  CGE_OK or CGE_ERR_INVALID_ARG. */
int cge_bvh_validate(const cge_scene_desc* desc, uint32_t* n_levels_out, uint32_t* n_leaves_out);
/* The tree CGE_TRAVERSAL_FAST walks (binned SAH, <= 4 primitives per leaf; csrc/sah_split.h specifies it), built by the GPU
 * builder (on_gpu != 0, the one cge_scene_create uses) or by the host builder (on_gpu == 0, no GPU needed).  Both must return
 * the same tree bit for bit.  Inner nodes in depth-first pre-order; child reference: bit 31 = leaf (bits 28..30 = count - 1,
 * bits 0..27 = first primitive in prim_order_out), else inner-node index.  nodes_out has room for *n_nodes_inout nodes (at
 * most n_triangles); either output pointer may be NULL.  build_ms_out: device time of the GPU build kernels. */
typedef struct cge_fast_node {
    float left_lower[3], left_upper[3];
    float right_lower[3], right_upper[3];
    uint32_t left, right;
} cge_fast_node;
int cge_fast_bvh_build(const cge_scene_desc* desc, int on_gpu, int device, cge_fast_node* nodes_out, uint32_t* n_nodes_inout,
                       uint32_t* prim_order_out, uint32_t* root_out, uint32_t* depth_out, uint32_t* n_leaves_out,
                       float* build_ms_out);
/* Host-only (no GPU needed): the host-side arithmetic of the two implemented ExtraFeatures, exposed for parity tests.
 * cge_ray_sample_positions: the n*n sample positions of getRaySamples (src/render.cpp:211-227) for pixel (x, y) of a
 * width x height frame, in normalised device coordinates (what the reference passes to Trackball::generateRay), drawn from
 * the hash-seeded MT19937 stream the device code uses (csrc/sampler.h).  ndc_out: n*n (x, y) pairs.
 * cge_bloom_weights: weightsGaussian(sigma) (src/render.cpp:198-210) as 9 floats, index [k + 1][j + 1]. */
int cge_ray_sample_positions(int32_t width, int32_t height, int32_t x, int32_t y, int32_t rays_per_pixel_side, uint32_t seed,
                             float* ndc_out);
int cge_bloom_weights(float sigma, float* weights9_out);
/* cge_hull_clear_host: the triangle test of the light-hull pre-pass (the function the shadow stage's pre-pass kernel calls), evaluated on the
 * host for n cases of (hit point o[3], parallelogram light v0 / edge01 / edge02 [9], triangle [9]); clear_out[i] = 1 claims that no ray from
 * o to any point of the light is accepted by the triangle.  Parity-test entry: the CPU suite checks the claim against the reference's
 * own intersectRayWithTriangle on sampled rays. */
int cge_hull_clear_host(const float* o3, const float* light9, const float* tri9, uint32_t n, int32_t* clear_out);
/* ... and the pre-pass's box test, same set-up: box6 = lower.xyz, upper.xyz; hit_out[i] = 0 claims that no ray from o to any point of
 * the light passes through the box within its length. */
int cge_hull_box_host(const float* o3, const float* light9, const float* box6, uint32_t n, int32_t* hit_out);
/* Copy out the BVH the library built (for parity tests against the reference's tree). Either pointer may be NULL. */
int cge_scene_bvh_export(const cge_scene* scene, cge_bvh_node* nodes_out, uint32_t* prim_order_out);

/* Render one frame (or this rank's tile subset).  rgb_out: width*height*3 floats in Screen::pixels() order —
 * row 0 = TOP of the image, index (H-1-y)*W + x (src/screen.cpp:41-47).  prim_id_out (optional, needs
 * CGE_FLAG_WANT_PRIM_IDS): width*height int32 global primitive id of the primary hit, -1 on miss, same pixel order.
 * Pixels outside this call's partition are left untouched.  Thread-safe on one cge_scene: every call uses its own
 * stream and scratch (reference src/main.cpp:514-528 renders several cameras concurrently). */
int cge_render(cge_scene* scene, const cge_camera* camera, const cge_params* params, float* rgb_out,
               int32_t* prim_id_out, cge_stats* stats_out);

/* Single-ray entry mirroring getFinalColor(scene,bvh,ray,features,depth) (src/render.h:35) for the debug-ray
 * callers (src/main.cpp:398,401): n rays in (origin[3],direction[3],t) records, n RGB out.  pixel ids for the
 * sampler are 0..n-1. */
int cge_trace_rays(cge_scene* scene, const float* rays7, uint32_t n, const cge_params* params, float* rgb_out,
                   int32_t* prim_id_out);

/* ---- device-function KATs: the six libIntersect functions (src/intersect.h:5-16) evaluated ON THE GPU ---- */
/* in/out arrays are host pointers; n cases each.  ray7 = origin[3], direction[3], t (t updated in place).  */
int cge_kat_triangle(const float* v0v1v2 /*9n*/, float* ray7 /*7n*/, int32_t* hit_out, uint32_t n, int device);
/* the same triangle test evaluated through the precomputed plane/edge rows the traversal kernel reads (not a
 * reference function: proves that hoisting the ray-independent part of I4 is bit-exact) */
int cge_kat_triangle_precomputed(const float* v0v1v2 /*9n*/, float* ray7 /*7n*/, int32_t* hit_out, uint32_t n, int device);
int cge_kat_aabb(const float* lower_upper /*6n*/, float* ray7, int32_t* hit_out, uint32_t n, int device);
int cge_kat_sphere(const float* center_radius /*4n*/, float* ray7, float* normal_out /*3n*/, int32_t* hit_out,
                   uint32_t n, int device);
int cge_kat_plane(const float* D_normal /*4n*/, float* ray7, int32_t* hit_out, uint32_t n, int device);
int cge_kat_triangle_plane(const float* v0v1v2 /*9n*/, float* D_normal_out /*4n*/, uint32_t n, int device);
int cge_kat_point_in_triangle(const float* v0v1v2 /*9n*/, const float* normal /*3n*/, const float* p /*3n*/,
                              int32_t* inside_out, uint32_t n, int device);

/* ---- multi-GPU: one process per GPU; the framebuffer tiles are gathered to rank 0 over NCCL -------------- */
typedef struct cge_comm cge_comm; /* opaque NCCL communicator wrapper */
#define CGE_UNIQUE_ID_BYTES 128
int cge_comm_unique_id(uint8_t id_out[CGE_UNIQUE_ID_BYTES]); /* rank 0 calls, broadcasts by any host means */
int cge_comm_create(const uint8_t id[CGE_UNIQUE_ID_BYTES], int rank, int n_ranks, int device, cge_comm** out);
int cge_comm_destroy(cge_comm* comm);
/* Render this rank's share of the frame - tile rows rank, rank + n_ranks, ... (part_index / part_count are overwritten with
 * rank / n_ranks, CGE_FLAG_PARTITION_TILE_ROWS is implied) - and deliver the frame on rank 0:
 *   default                      the other ranks render into a compact buffer that holds just their rows and ncclSend it as it is;
 *                                rank 0 renders into the frame, receives, scatters every rank's rows in one launch and (unless
 *                                CGE_FLAG_RGB_DEVICE_PTR) copies the frame to rgb_out.  rgb_out / prim_id_out are only used on rank 0.
 *   CGE_FLAG_SHARED_HOST_FRAME   rgb_out (and prim_id_out) are frames from cge_comm_host_frame, the same memory on every rank: each
 *                                rank copies the rows it rendered straight into it over its own PCIe link (one strided copy), and a
 *                                4-byte all-reduce tells every rank when the frame is complete.  No device-side gather at all.
 *   CGE_FLAG_PEER_FRAME          rgb_out is this rank's pointer from cge_comm_peer_frame: ONE device frame on rank 0 that every rank's
 *                                render kernels store into directly (peer memory over NVLink / NVSwitch); a 4-byte all-reduce behind
 *                                the kernels tells every rank when the frame is complete on rank 0.  Float frame, no prim_id_out.
 * The image does not depend on the number of ranks (bit-identical to cge_render). */
int cge_render_distributed(cge_scene* scene, cge_comm* comm, const cge_camera* camera, const cge_params* params,
                           float* rgb_out, int32_t* prim_id_out, cge_stats* stats_out);
/* Collective over the communicator: `bytes` of host memory mapped by every rank (POSIX shared memory, page-locked in each
 * process), for CGE_FLAG_SHARED_HOST_FRAME.  *out is this process's address of it.  Released by cge_comm_destroy. */
int cge_comm_host_frame(cge_comm* comm, uint64_t bytes, void** out);
/* Collective over the communicator: `bytes` of device memory on rank 0 that every rank can address (CUDA IPC; the ranks are
 * separate processes on one node with peer access to rank 0's GPU), for CGE_FLAG_PEER_FRAME.  *out is this process's device
 * pointer to it: on rank 0 the frame itself (readable like any device buffer once cge_render_distributed has returned), on the
 * other ranks a mapping of it.  Released by cge_comm_destroy. */
int cge_comm_peer_frame(cge_comm* comm, uint64_t bytes, void** out);

/* pinned host memory for rgb_out / prim_id_out so the D2H copy runs at full PCIe speed (optional) */
int cge_host_alloc(void** out, uint64_t bytes);
int cge_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CGE_H_ */
