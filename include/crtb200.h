/* crtb200.h -- C ABI of the B200-native renderer core (libcrtb200.so).
 *
 * Drop-in boundary for the per-pixel hot path of the Chaos course ray tracer
 * (reference: SourceCode/src/RayTracer.cpp).  The reference has no FFI; its boundary is the C++ member
 *     std::vector<std::vector<Color>> RayTracer::render(const std::string&, RenderOptions)
 *         (include/tracer/RayTracer.h:100, src/RayTracer.cpp:204-298)
 * constructed by  explicit RayTracer(Scene&)  (RayTracer.h:97, RayTracer.cpp:45-51), camera mutated through
 * Camera& setCamera() (RayTracer.h:99).  Every entry point below names the reference interface it replaces.
 * INTEGRATION.md shows the binding a maintainer adds to RayTracer.cpp.
 *
 * Conventions: plain C, plain pointers + sizes, no torch / C++ types.  Every function returns 0 on success or a
 * negative crtb200_status; crtb200_last_error_ctx(ctx) returns the message of the last failed call on that context
 * (crtb200_last_error() the last one of the calling thread, for calls that have no context yet).
 * Host pointers handed to upload/render are borrowed for the duration of the call only.  One context = one caller
 * thread at a time (the reference's render() is not re-entrant either, RayTracer.cpp:205-206); a context drives one
 * GPU (crtb200_create) or several (crtb200_create_multi: scene replicated, every frame split by tiles).
 * There is NO CPU fallback: every call fails with CRTB200_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef CRTB200_H
#define CRTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRTB200_ABI_VERSION 2u
#define CRTB200_INVALID 0xFFFFFFFFu /* = INVALID_INDEX, include/tree/KDTree.h:11 */

typedef enum crtb200_status {
  CRTB200_OK = 0,
  CRTB200_ERR_ARG = -1,    /* null / inconsistent argument                           */
  CRTB200_ERR_CUDA = -2,   /* CUDA runtime error or no usable device                 */
  CRTB200_ERR_STATE = -3,  /* call order (render before upload, ...)                 */
  CRTB200_ERR_MEMORY = -4, /* the frame does not fit the configured device budget    */
  CRTB200_ERR_SCENE = -5   /* scene arrays fail validation (index out of range, ...) */
} crtb200_status;

/* enum MaterialType { Diffuse, Reflective, Constant, Refractive }  include/tracer/Material.h:7 */
enum { CRTB200_MAT_DIFFUSE = 0, CRTB200_MAT_REFLECTIVE = 1, CRTB200_MAT_CONSTANT = 2, CRTB200_MAT_REFRACTIVE = 3 };
/* texture classes of include/tracer/Texture.h:18-59 */
enum { CRTB200_TEX_ALBEDO = 0, CRTB200_TEX_EDGES = 1, CRTB200_TEX_CHECKER = 2, CRTB200_TEX_BITMAP = 3 };
/* enum RayType { PrimaryRay, ShadowRay, ReflectionRay, RefractionRay, DiffuseRay }  include/tracer/Ray.h:14 */
enum { CRTB200_RAY_PRIMARY = 0, CRTB200_RAY_SHADOW = 1, CRTB200_RAY_REFLECTION = 2, CRTB200_RAY_REFRACTION = 3 };

/* One KD-tree node in the REFERENCE's numbering: KDTree<T>::TreeNode (include/tree/KDTree.h:16-21), nodes appended
 * in DFS pre-order by createNode (KDTree.h:33-38), root = 0 of its tree.  A node is a leaf iff leaf_count > 0
 * (KDTree.cpp:58 tests !indexes.empty()).  child indices are relative to the tree's first node. */
typedef struct crtb200_kdnode {
  float box_min[3];
  float box_max[3];
  uint32_t child[2];   /* CRTB200_INVALID when absent (KDTree.cpp:35,41)                        */
  uint32_t leaf_start; /* into the tree's leaf-reference array, relative to the tree's first ref */
  uint32_t leaf_count;
} crtb200_kdnode;

/* Mesh (include/tracer/Scene.h:25-38) + its TriangleKDTree (include/tree/AccelerationStructure.h:6-13). */
typedef struct crtb200_mesh {
  uint32_t material;       /* index into materials                                              */
  uint32_t first_triangle; /* the mesh's triangles are [first_triangle, first_triangle + n)     */
  uint32_t n_triangles;
  uint32_t first_vertex; /* informational; triangle_vertex holds GLOBAL vertex ids             */
  uint32_t n_vertices;
  uint32_t first_node; /* the mesh tree's nodes are mesh_nodes[first_node .. first_node+n_nodes) */
  uint32_t n_nodes;
  uint32_t first_leaf_ref; /* mesh_leaf_refs[first_leaf_ref ..) hold MESH-LOCAL triangle indices  */
  uint32_t n_leaf_refs;
} crtb200_mesh;

/* Material (include/tracer/Material.h:9-30). texture = CRTB200_INVALID for the non-USE_TEXTURES flavour. */
typedef struct crtb200_material {
  uint32_t type;
  uint32_t smooth_shading;
  uint32_t texture;
  float albedo[3];
  float ior;
} crtb200_material;

/* Texture classes (include/tracer/Texture.h:18-59, src/Texture.cpp:14-72).
 *   ALBEDO : color_a                         EDGES  : color_a = inner, color_b = edge, scalar = width
 *   CHECKER: color_a / color_b, scalar = square size
 *   BITMAP : width x height float RGB texels at texels[3*texel_offset ..] (already /255, Texture.cpp:55-59) */
typedef struct crtb200_texture {
  uint32_t kind;
  float color_a[3];
  float color_b[3];
  float scalar;
  uint32_t width, height;
  uint64_t texel_offset;
} crtb200_texture;

/* struct Light { Vector position; unsigned intentsity; }  include/tracer/Scene.h:20-23 */
typedef struct crtb200_light {
  float position[3];
  uint32_t intensity;
} crtb200_light;

/* Everything the hot path reads: Scene (Scene.h:50-69), the scene AccelerationStructure (RayTracer.h:65) flattened.
 * All arrays are host memory; indices are validated on upload. */
typedef struct crtb200_scene {
  uint32_t abi_version; /* CRTB200_ABI_VERSION */
  uint32_t width, height; /* sceneSettings.image */
  float background[3];    /* sceneSettings.sceneBackgroundColor */

  uint32_t n_vertices;
  const float *vertex_position; /* 3 per vertex: Vertex::position                                         */
  const float *vertex_normal;   /* 3 per vertex: Vertex::normal after Mesh::Mesh (Scene.cpp:5-30)         */
  const float *vertex_uv;       /* 3 per vertex (Vertex::UV) or NULL                                      */

  uint32_t n_triangles;
  const uint32_t *triangle_vertex; /* 3 global vertex ids per triangle, in Mesh::triangles order            */
  const float *triangle_normal;    /* 3 per triangle: Triangle::normal (Triangle.cpp:13-16), as computed by  */
                                   /* the host front end in binary32 without FMA                            */
  uint32_t n_meshes;
  const crtb200_mesh *meshes;
  uint32_t n_materials;
  const crtb200_material *materials;
  uint32_t n_textures;
  const crtb200_texture *textures;
  uint64_t n_texels;
  const float *texels;
  uint32_t n_lights;
  const crtb200_light *lights;

  uint32_t n_mesh_nodes;
  const crtb200_kdnode *mesh_nodes;
  uint32_t n_mesh_leaf_refs;
  const uint32_t *mesh_leaf_refs;
  uint32_t n_top_nodes; /* ObjectKDTree (AccelerationStructure.h:21-33): leaves reference mesh indices */
  const crtb200_kdnode *top_nodes;
  uint32_t n_top_leaf_refs;
  const uint32_t *top_leaf_refs;
} crtb200_scene;

/* Camera state read by getRay (RayTracer.cpp:61-80): position + row-major 3x3 (Camera.h:7-8). */
typedef struct crtb200_camera {
  float position[3];
  float rotation[9];
} crtb200_camera;

/* One rectangle handed to renderRectangle(row, col, w, h)  (RayTracer.cpp:82-85). */
typedef struct crtb200_rect {
  uint32_t row, col, width, height;
} crtb200_rect;

/* RenderOptions (RayTracer.h:25-50) minus the GI members (out of scope: clock-seeded RNG) + the rectangle list
 * the reference's schedulers would produce (RayTracer.cpp:114-202).  n_rects = 0 renders the whole image. */
typedef struct crtb200_options {
  uint32_t max_depth;    /* MAX_DEPTH, default 5           */
  float shadow_bias;     /* SHADOW_BIAS, default 1e-4f     */
  float reflection_bias; /* REFLECTION_BIAS, default 1e-4f */
  float refraction_bias; /* REFRACTION_BIAS, default 1e-4f */
  uint32_t n_rects;
  const crtb200_rect *rects;
  uint32_t traversal; /* 0 = DEFAULT: the reference's walk (KDTree.cpp:48-87, 127-166) minus the subtrees that provably  */
                      /*     cannot contribute (conservative culling with a per-mesh margin), long    */
                      /*     walks finished by a group of lanes (tail hand-off), shadow rays whose light */
                      /*     term is exactly zero answered without a walk (3.9).  Results are identical to the        */
                      /*     reference's: hit ids, t, float RGB, ray counts.                                          */
                      /* 1 = the reference's literal itinerary: every node whose box passes is visited, nothing is    */
                      /*     skipped or reordered (what count_work = 1 counts; the round-1 default).                  */
  uint32_t count_work; /* 0 = off; 1 = count node / triangle tests under the reference's visit-all rule (shadow  */
                       /* early termination disabled; same pixels) -- the figure the roofline arithmetic uses;  */
                       /* 2 = count the tests the production kernels really perform                             */
  /* tile sharding (multi-GPU): render only tile blocks b with b % shard_count == shard_index; 0/1 = all */
  uint32_t shard_index, shard_count;
  /* crtb200_render_device with shard_count > 1: 0 = write the shard's pixels compactly (a slab, for an NCCL gather +    */
  /* crtb200_assemble_shards); 1 = write them at their place in a FULL frame -- d_rgb_out / d_rgb8_out then usually    */
  /* point at the gathering rank's frame, mapped with crtb200_ipc_open: the store kernel does the transfer (NVLink)   */
  uint32_t shard_full_frame;
} crtb200_options;

/* Primary-ray closest hit, the record the parity gate compares (mesh index, mesh-local triangle index, t). */
typedef struct crtb200_hit {
  int32_t mesh; /* -1 = miss */
  int32_t triangle;
  float t;
} crtb200_hit;

typedef struct crtb200_stats {
  uint64_t rays_primary, rays_shadow, rays_reflection, rays_refraction; /* traced rays, SURVEY 8(d) definition */
  /* AABB / triangle tests per options.count_work, split by kernel: closest-hit (K2) and shadow (K3) */
  uint64_t node_tests_closest, triangle_tests_closest;
  uint64_t node_tests_shadow, triangle_tests_shadow;
  double device_ms;  /* CUDA-event time of the whole frame on the launching stream                          */
  double closest_ms; /* of which: closest-hit traversal launches (K2, all levels)                            */
  double shadow_ms;  /* of which: shadow any-hit + accumulate launches (K3)                                  */
  double total_ms;   /* host wall time of the call, copies included                                          */
  uint32_t kernel_launches;
  uint32_t levels;
  /* tail hand-off: walks finished one warp per ray by k_coop, and the part of closest_ms / shadow_ms  */
  /* spent there                                                                                                      */
  uint64_t handoff_closest, handoff_shadow;
  double coop_closest_ms, coop_shadow_ms;
  /* shadow rays (included in rays_shadow) that were answered without a walk because their light term is exactly zero  */
  /* whatever the visibility: surface turned away from the light, RayTracer.cpp:313-327                */
  uint64_t shadow_rays_zero_term;
} crtb200_stats;

typedef struct crtb200_ctx crtb200_ctx;

/* library / device */
uint32_t crtb200_abi_version(void);
const char *crtb200_last_error(void);
int crtb200_device_count(int *count);

const char *crtb200_last_error_ctx(const crtb200_ctx *ctx);

/* replaces: RayTracer::RayTracer(Scene&) resource acquisition (RayTracer.cpp:45-51) */
int crtb200_create(int device, crtb200_ctx **out);
/* The same on n GPUs of one node (SURVEY 8(b) "Proposed exports", 8(e)): replaces the std::thread bucket pool of
 * RayTracer::renderBucketsThreadpool (RayTracer.cpp:141-158) at device scale.  The context behaves like a single-GPU
 * one -- same calls, same results bit for bit -- but crtb200_upload_scene replicates the scene on every GPU and every
 * frame's 8x4-pixel tiles are dealt round-robin over the GPUs; each GPU stores its pixels straight into the frame on
 * device_ids[0] through peer-mapped memory (NVLink), from where the frame is returned.  A device id may repeat (several
 * tile shards on one GPU; used by the single-GPU tests). */
int crtb200_create_multi(const int *device_ids, int n, crtb200_ctx **out);
int crtb200_device_list(const crtb200_ctx *ctx, int *device_ids, int capacity, int *count);
int crtb200_destroy(crtb200_ctx *ctx);
/* device budget in bytes for the per-frame ray queues (default 16 GiB); frames that need more are chunked */
int crtb200_set_queue_budget(crtb200_ctx *ctx, uint64_t bytes);
/* chunks of a frame rendered concurrently on separate streams (default 2; 1 = strictly sequential kernels, which is
 * what the per-kernel timers closest_ms / shadow_ms of crtb200_stats require -- they read 0 otherwise) */
int crtb200_set_concurrency(crtb200_ctx *ctx, uint32_t chunks_in_flight);

/* replaces: the data RayTracer keeps in `scene`, `boundingBox`, `accelerationStructure` (RayTracer.h:64-67).
 * Copies + re-lays-out everything to device SoA during the call (host flattener H1). */
int crtb200_upload_scene(crtb200_ctx *ctx, const crtb200_scene *scene);

/* replaces: RayTracer::render's scheduler switch + renderRectangle pixel loop (RayTracer.cpp:82-112,209-286).
 * rgb_out: H*W*3 float32 row-major, host memory, = the colorBuffer render() returns (RayTracer.cpp:297).  Pixels not
 * covered by options->rects keep the value of the previous render of this context (initially 0), like colorBuffer.
 * rgb8_out (optional): H*W*3 uint8, PPMColor quantisation (Color.cpp:12-16).
 * hits_out (optional): H*W primary-ray closest hits.   stats (optional). */
int crtb200_render(crtb200_ctx *ctx, const crtb200_camera *camera, const crtb200_options *options, float *rgb_out,
                   uint8_t *rgb8_out, crtb200_hit *hits_out, crtb200_stats *stats);

/* replaces: the animation loop body app/animation.cpp:24-38 (setCamera + render per frame), batched.
 * rgb_out / rgb8_out: n_frames consecutive frames. */
int crtb200_render_frames(crtb200_ctx *ctx, const crtb200_camera *cameras, uint32_t n_frames,
                          const crtb200_options *options, float *rgb_out, uint8_t *rgb8_out, crtb200_stats *stats);

/* Device-resident variant for callers that own device memory and a stream (multi-GPU gather, benchmarks):
 * d_rgb_out / d_rgb8_out are DEVICE pointers (either may be NULL), stream is a cudaStream_t (NULL = default).
 * Asynchronous: returns after enqueueing.  With tile sharding the output is the full-frame layout; pixels of
 * other shards are left untouched.  Stats are available through crtb200_last_stats after synchronisation. */
int crtb200_render_device(crtb200_ctx *ctx, const crtb200_camera *camera, const crtb200_options *options,
                          float *d_rgb_out, uint8_t *d_rgb8_out, void *stream);
int crtb200_last_stats(crtb200_ctx *ctx, crtb200_stats *stats);

/* Tile sharding for multi-GPU (one context per GPU, scene replicated).  8x4-pixel tiles are dealt
 * round-robin: tile t belongs to shard t % shard_count.  With options->shard_count > 1 crtb200_render_device writes the
 * shard's pixels COMPACTLY into d_rgb_out (crtb200_shard_items(ctx, shard_count) x 3 floats, equal for all shards so the
 * slabs can be all-gathered / gathered with NCCL); d_rgb8_out must be NULL.  crtb200_assemble_shards scatters
 * shard_count consecutive slabs back into a full frame (and/or PPMColor bytes) on the gathering rank. */
int crtb200_shard_items(crtb200_ctx *ctx, uint32_t shard_count, uint32_t *items);
/* One process per GPU without a gather: the gathering rank exports its frame buffer (cudaIpcGetMemHandle, 64 bytes to
 * ship to the other ranks by any means), the others map it and render their shard straight into it with
 * options.shard_full_frame = 1.  Completion still needs a barrier between the ranks' streams (e.g. one tiny NCCL
 * all-reduce).  d_ptr must be the base of a cudaMalloc allocation of the exporting process. */
#define CRTB200_IPC_HANDLE_BYTES 64
int crtb200_ipc_alloc(int device, size_t bytes, void **d_ptr_out); /* a whole cudaMalloc allocation, exportable */
int crtb200_ipc_free(int device, void *d_ptr);
int crtb200_ipc_export(void *d_ptr, uint8_t handle_out[CRTB200_IPC_HANDLE_BYTES]);
int crtb200_ipc_open(int device, const uint8_t handle[CRTB200_IPC_HANDLE_BYTES], void **d_ptr_out);
int crtb200_ipc_close(int device, void *d_ptr);
int crtb200_assemble_shards(crtb200_ctx *ctx, const float *d_slabs, uint32_t shard_count, float *d_rgb_out,
                            uint8_t *d_rgb8_out, void *stream);

/* replaces: RayTracer::getRay + the re-normalisation of shootRay (RayTracer.cpp:61-80,420), exposed so the
 * parity tests can compare primary rays bit for bit.  rays_out: H*W*6 float32 (origin xyz, direction xyz), host. */
int crtb200_generate_rays(crtb200_ctx *ctx, const crtb200_camera *camera, float *rays_out);

/* replaces: RayTracer::trace (RayTracer.cpp:453-458) / RayTracer::hasIntersection (RayTracer.cpp:507-518) for caller-
 * supplied rays.  rays: n*6 float32 host; ray_type: CRTB200_RAY_*; for shadow queries max_distance holds n floats
 * and occluded_out n bytes; for closest-hit queries hits_out holds n records. */
int crtb200_trace_rays(crtb200_ctx *ctx, const float *rays, uint32_t n, uint32_t ray_type, uint32_t traversal,
                       const float *max_distance, crtb200_hit *hits_out, uint8_t *occluded_out);

/* Test hook: evaluates the device's (1 - cos)^5 routine (csrc/crt_powf5.h), the stand-in for glibc's
 * std::powf(x, 5) of RayTracer.cpp:407, on n host floats -- so the parity suite can check it bit for bit against libm. */
int crtb200_debug_powf5(crtb200_ctx *ctx, const float *x, uint32_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif /* CRTB200_H */
